# every bench workload once (short): C3 with the robustness cases, C2, C4, C5
mkdir -p gpurun_out
T=$1
for W in C3 C2 C4 C5; do
  X=""; if [ $W != C3 ]; then X="--no-cpu-baseline"; fi
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_$W.json 2> gpurun_out/${T}_$W.err; echo "$W rc=$?"; tail -2 gpurun_out/${T}_$W.err
done
python - "$T" <<'PY'
import json,sys
for w in ('C3','C2','C4','C5'):
    try:
        d=json.load(open('gpurun_out/%s_%s.json'%(sys.argv[1],w)))
    except Exception as e:
        print(w,'no json',e); continue
    keys=('ms_per_step','value','unit')
    print(w,{k:d.get(k) for k in keys},'e2e',d.get('e2e'),'\n   roofline',d.get('roofline_binding') or d.get('roofline'),'\n   extra',{k:d[k] for k in d if k in ('e2e_compact','e2e_resident','robustness','stages_ms','c4','c5','check')})
PY

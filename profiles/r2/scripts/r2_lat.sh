# where the generation's time goes besides K2: stage events at P = 500 and at one rank's share of 8 GPUs, K0 phase clocks, K2 timeline
mkdir -p gpurun_out
T=$1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-cases > gpurun_out/${T}_c3.json 2> gpurun_out/${T}_c3.err; echo "c3 rc=$?"
python bench.py --workload C3s8 --steps 10 --warmup 3 --no-cpu-baseline --no-cases > gpurun_out/${T}_c3s8.json 2> gpurun_out/${T}_c3s8.err; echo "c3s8 rc=$?"
SCB_B200_LIB=scbonita_b200/libscbonita_trace.so python profiles/r2/scripts/trace_k0.py C3 > gpurun_out/${T}_k0.txt 2>&1; echo "k0 rc=$?"
cat gpurun_out/${T}_k0.txt
for W in C3s8 C3; do
SCB_B200_LIB=scbonita_b200/libscbonita_trace.so python profiles/r2/scripts/trace_events.py $W > gpurun_out/${T}_ev_$W.txt 2>&1; echo "ev $W rc=$?"
cat gpurun_out/${T}_ev_$W.txt
done
python - "$T" <<'PY'
import json,sys
for w in ('c3','c3s8'):
    d=json.load(open('gpurun_out/%s_%s.json'%(sys.argv[1],w)))
    print(w,'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'stages',d['stages_ms'],'plain',{k:round(v,4) for k,v in d['stats'].items() if k.endswith('_ms')})
PY

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_steady.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_steady.log
B="timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
for o in "steady=1" "steady=0" "steady=1,warps=8" "steady=1,warps=12" "steady=1,warps=16" "steady=1,warps=24" "steady=1,check_every=4" "steady=1,check_every=8"; do
  args=""; for kv in $(echo $o | tr ',' ' '); do args="$args --opt $kv"; done
  $B $args > "gpurun_out/st_$o.json" 2>> gpurun_out/st.err
  python - "gpurun_out/st_$o.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'smem frac', round(d['roofline_binding']['frac'],3), d['roofline_binding'].get('peak'), 'steps', d['stats']['steps_executed'], 'rows', d['stats'].get('state_rows_moved'), 'chk', d['fitness_checksum'])
PY
done

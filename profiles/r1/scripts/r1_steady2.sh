mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
for o in "steady=1" "steady=0" "steady=1,warps=16" "steady=1,warps=24" "steady=1,check_every=8" "steady=1,check_every=8,warps=16"; do
  args=""; for kv in $(echo $o | tr ',' ' '); do args="$args --opt $kv"; done
  $B $args > "gpurun_out/st2_$o.json" 2>> gpurun_out/st2.err
  python - "gpurun_out/st2_$o.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'smem frac', round(d['roofline_binding']['frac'],3), 'steps', d['stats']['steps_executed'], 'rows', d['stats'].get('state_rows_moved'), 'chk', d['fitness_checksum'])
PY
done

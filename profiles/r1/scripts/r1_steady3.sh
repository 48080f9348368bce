mkdir -p gpurun_out
run() { name=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 "$@" > "gpurun_out/st3_$name.json" 2>> gpurun_out/st3.err
  python - "gpurun_out/st3_$name.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'e2ec', round(d['e2e_compact']['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'smem frac', round(d['roofline_binding']['frac'],3), 'steps', d['stats']['steps_executed'], 'rows', d['stats'].get('state_rows_moved'), 'chk', d['fitness_checksum'])
PY
}
run s1
run s0 --opt steady=0
run s1_w16 --opt warps=16
run s1_w20chk8gap6 --opt check_every=8 --opt snap_min_gap=6
run s1_chk8 --opt check_every=8
run s1_chk10 --opt check_every=10
run s1_gap4 --opt snap_min_gap=4

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_stage.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_stage.log
run() { name=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 "$@" > "gpurun_out/st4_$name.json" 2>> gpurun_out/st4.err
  python - "gpurun_out/st4_$name.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'e2ec', round(d['e2e_compact']['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'kernel_ms', round(d['stats']['kernel_ms'],3), 'smem frac', round(d['roofline_binding']['frac'],3), 'steps', d['stats']['steps_executed'], 'rows', d['stats'].get('state_rows_moved'), 'chk', d['fitness_checksum'])
PY
}
run s1
run s0 --opt steady=0
run s1_w16 --opt warps=16
run s1_chk8 --opt check_every=8
run s1_chk10 --opt check_every=10

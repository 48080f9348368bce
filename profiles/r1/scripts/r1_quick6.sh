mkdir -p gpurun_out
run() { name=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 "$@" > "gpurun_out/q6_$name.json" 2>> gpurun_out/q6.err
  python - "gpurun_out/q6_$name.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'kernel_ms', round(d['stats']['kernel_ms'],3), 'steps', d['stats']['steps_executed'], 'chk', d['fitness_checksum'])
PY
}
run def
run w16 --opt warps=16
run chk8 --opt check_every=8
run w16chk8 --opt warps=16 --opt check_every=8

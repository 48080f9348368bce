mkdir -p gpurun_out
run() { name=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 "$@" > "gpurun_out/q9_$name.json" 2>> gpurun_out/q9.err
  python - "gpurun_out/q9_$name.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'steps', d['stats']['steps_executed'], 'resolver', d['stats']['resolver_chunks'], 'chk', d['fitness_checksum'])
PY
}
run def
run gap12 --opt snap_min_gap=12
run gap16 --opt snap_min_gap=16
run gap5 --opt snap_min_gap=5
run s40_70 --opt snap_lo=1099511627776 --opt snap_hi=64
run s32_66 --opt snap_lo=4294967296 --opt snap_hi=4
run chk10 --opt check_every=10
run chk12 --opt check_every=12

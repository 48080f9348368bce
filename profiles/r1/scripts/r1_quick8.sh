mkdir -p gpurun_out
run() { name=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 "$@" > "gpurun_out/q8_$name.json" 2>> gpurun_out/q8.err
  python - "gpurun_out/q8_$name.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'steps', d['stats']['steps_executed'], 'chk', d['fitness_checksum'])
PY
}
run from4
run from3 --opt check_from=3
run from2 --opt check_from=2
run from1 --opt check_from=1
run none --opt check_from=-1
run from3_chk6 --opt check_from=3 --opt check_every=6
run from2_chk6 --opt check_from=2 --opt check_every=6

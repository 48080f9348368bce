mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/q3_$name.json 2>> gpurun_out/q3.err; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytestq3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytestq3.log
run default
run old_sched --opt check_every=4 --opt snap_min_gap=1 --opt snap_lo=$((1<<32)) --opt snap_hi=1
run gap4 --opt snap_min_gap=4
run gap12 --opt snap_min_gap=12
run chk4 --opt check_every=4
run chk8 --opt check_every=8
run warps24 --opt warps=24
run warps16 --opt warps=16
tail -n 3 gpurun_out/pytestq3.log

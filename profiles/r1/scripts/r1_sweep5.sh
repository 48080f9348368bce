mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/s5_$name.json 2>> gpurun_out/s5.err; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest5.log
run default
run snap36_66 --opt snap_lo=$((1<<36)) --opt snap_hi=4
run plain --opt early_exit=0 --opt snapshots=0
run c2 --workload C2
tail -n 3 gpurun_out/pytest5.log

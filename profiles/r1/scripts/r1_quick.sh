mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/q_$name.json 2>> gpurun_out/q.err; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytestq.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytestq.log
run default
run plain --opt early_exit=0 --opt snapshots=0
run warps20 --opt warps=20
run warps16 --opt warps=16
run c2 --workload C2
tail -n 3 gpurun_out/pytestq.log

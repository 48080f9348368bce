mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_tmem.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tmem.log
B="timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/tm_$name.json 2>> gpurun_out/tm.err; }
run tmem1
run tmem0 --opt tmem=0
run tmem1_gap1 --opt snap_min_gap=1
run tmem1_gap4 --opt snap_min_gap=4
run tmem1_c2 --workload C2
tail -n 3 gpurun_out/pytest_tmem.log

mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/s4_$name.json 2>> gpurun_out/s4.err; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest4.log
run default
run warps16 --opt warps=16
run warps20 --opt warps=20
run snap36_66 --opt snap_lo=$((1<<36)) --opt snap_hi=4
run chk6 --opt check_every=6
run plain --opt early_exit=0 --opt snapshots=0
run c2 --workload C2
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -n 3 gpurun_out/pytest4.log

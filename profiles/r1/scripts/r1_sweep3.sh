mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/s3_$name.json 2>> gpurun_out/s3.err; }
run default
run warps16 --opt warps=16
run warps20 --opt warps=20
run snap36_66 --opt snap_lo=$((1<<36)) --opt snap_hi=4
run snap16_32_64 --opt snap_lo=$(((1<<16)|(1<<32))) --opt snap_hi=1
run chk3 --opt check_every=3
run chk2 --opt check_every=2
run chk6 --opt check_every=6
run plain --opt early_exit=0 --opt snapshots=0
run plain16 --opt early_exit=0 --opt snapshots=0 --opt warps=16
run earlyonly --opt snapshots=0
run c2 --workload C2

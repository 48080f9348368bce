mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/q2_$name.json 2>> gpurun_out/q2.err; }
run default
run plain --opt early_exit=0 --opt snapshots=0
run warps20 --opt warps=20
run plain20 --opt early_exit=0 --opt snapshots=0 --opt warps=20
run plain16 --opt early_exit=0 --opt snapshots=0 --opt warps=16

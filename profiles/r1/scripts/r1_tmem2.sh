mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/tm2_$name.json 2>> gpurun_out/tm2.err; }
run tmem1
run tmem0 --opt tmem=0
run tmem1_gap4 --opt snap_min_gap=4
run tmem1_gap1 --opt snap_min_gap=1
run tmem1_w24 --opt warps=24
run tmem1_w16 --opt warps=16

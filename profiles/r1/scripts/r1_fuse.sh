mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fuse_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/fuse_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_fuse.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fuse.log
B="timeout 120 python bench.py --no-cpu-baseline --steps 8 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/fu_$name.json 2>> gpurun_out/fu.err; }
run fuse1
run fuse0 --opt fuse=0
run fuse1b
run fuse0b --opt fuse=0
tail -n 2 gpurun_out/fuse_smoke.log; tail -n 3 gpurun_out/pytest_fuse.log

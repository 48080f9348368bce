mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/tma_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/tma_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_tma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tma.log
B="timeout 120 python bench.py --no-cpu-baseline --steps 8 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/tma_$name.json 2>> gpurun_out/tma.err; }
run tma1
run tma0 --opt tma=0
run tma1b
run tma0b --opt tma=0
tail -n 2 gpurun_out/tma_smoke.log | cut -c1-150; tail -n 3 gpurun_out/pytest_tma.log

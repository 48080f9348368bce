mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_compact.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_compact.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/cp_bench.json 2> gpurun_out/cp_bench.err; echo "bench rc=$?"; cat gpurun_out/cp_bench.json
timeout 300 python bench.py --workload C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/cp_bench_c2.json 2>> gpurun_out/cp_bench.err; cat gpurun_out/cp_bench_c2.json

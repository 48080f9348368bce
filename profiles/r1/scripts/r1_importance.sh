mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/imp_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_imp.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_imp.log
timeout 300 python profiles/scripts/importance_timing.py 200 50000 40 > gpurun_out/imp_c3.json 2> gpurun_out/imp.err; cat gpurun_out/imp_c3.json
timeout 300 python profiles/scripts/importance_timing.py 60 10000 60 > gpurun_out/imp_c2.json 2>> gpurun_out/imp.err; cat gpurun_out/imp_c2.json
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/imp_bench.json 2> gpurun_out/imp_bench.err; cat gpurun_out/imp_bench.json

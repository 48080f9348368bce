mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c5.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_c5.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/c5_bench.json'))
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'e2ec', round(d['e2e_compact']['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'smem frac', round(d['roofline_binding']['frac'],3), 'chk', d['fitness_checksum'], d['cpu_baseline'])
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/c5_launches.csv $CMD > gpurun_out/c5_ncu_launches.log 2>&1
grep -E "scb_k" gpurun_out/c5_launches.csv | awk -F'","' '{print $5, $NF}' | head -8

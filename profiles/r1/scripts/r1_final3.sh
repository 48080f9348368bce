mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/f3_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f3_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/f3_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f3_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/f3_bench.json 2> gpurun_out/f3_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f3_ref.json 2>> gpurun_out/f3_bench.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/f3_bench.json'))
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'e2ec', round(d['e2e_compact']['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'smem frac', round(d['roofline_binding']['frac'],3), 'rows', d['stats']['state_rows_moved'], 'chk', d['fitness_checksum'])
print(open('gpurun_out/f3_ref.json').read()[:300])
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/f3_launches.csv $CMD > gpurun_out/f3_ncu_launches.log 2>&1
grep -E "scb_k" gpurun_out/f3_launches.csv | awk -F'","' '{print $5, $NF}' | head -4
CMD2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt fuse=0"
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 1 -c 1 -f -o gpurun_out/f3_prof_sim $CMD2 > gpurun_out/f3_ncu_full.log 2>&1
echo "ncu full rc=$?"

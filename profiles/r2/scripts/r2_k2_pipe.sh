# round 2, step 1: software-pipelined K2 step -- parity on the GPU, bench A/B against round 1 (3.26 ms / K2 2.785 ms), ncu
mkdir -p gpurun_out
T=$1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
for V in "" "--opt warps=16" "--opt warps=12"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $V > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench [$V] rc=$?"
  python - "$T" <<'PY'
import json,sys
d=json.load(open('gpurun_out/%s_bench.json'%sys.argv[1]))
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'smem frac', round(d['roofline_binding']['frac'],3), 'rows', d['stats']['state_rows_moved'], 'steps', d['stats']['steps_executed'], 'chk', d['fitness_checksum'], d['config']['geometry'])
PY
  cp gpurun_out/${T}_bench.json "gpurun_out/${T}_bench$(echo $V | tr -d ' =-').json"
done
CMD2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt fuse=0"
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 1 -c 1 -f -o gpurun_out/${T}_prof_sim $CMD2 > gpurun_out/${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"

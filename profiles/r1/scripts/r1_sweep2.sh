mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest2.log
for o in "warps=32" "warps=24" "warps=16" "early_exit=0,snapshots=0" "early_exit=0,snapshots=0,warps=16" "period_search=16" "check_every=8" "words=2"; do
  args=""; for kv in $(echo $o | tr ',' ' '); do args="$args --opt $kv"; done
  $B $args > gpurun_out/s2_$o.json 2>> gpurun_out/s2.err
done
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_for_ncu2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 1 -c 1 -o gpurun_out/prof_sim2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest2.log

mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3"
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
for o in "warps=32" "warps=24" "warps=8" "words=2" "words=2,warps=32" "check_every=2" "check_every=8" "period_search=16" "early_exit=0,snapshots=0" "early_exit=0,snapshots=0,warps=32" "early_exit=0,snapshots=0,words=2,warps=32"; do
  args=""; for kv in $(echo $o | tr ',' ' '); do args="$args --opt $kv"; done
  $B $args > gpurun_out/sweep_$o.json 2>> gpurun_out/sweep.err
done
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_for_ncu2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 1 -c 1 -o gpurun_out/prof_sim python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out

mkdir -p gpurun_out
CMD2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt fuse=0"
timeout 300 $CMD2 > gpurun_out/n7_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 1 -c 1 -f -o gpurun_out/n7_prof_sim $CMD2 > gpurun_out/n7_ncu_full.log 2>&1
echo "full rc=$?"

mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ncu_launches.log 2>&1
CMD2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt fuse=0"
timeout 300 $CMD2 > gpurun_out/final_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 1 -c 1 -o gpurun_out/final_prof_sim $CMD2 > gpurun_out/final_ncu_full.log 2>&1

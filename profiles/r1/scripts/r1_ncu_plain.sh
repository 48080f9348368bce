mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --opt early_exit=0 --opt snapshots=0"
timeout 300 $CMD > gpurun_out/plain_run.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 0 -c 1 -o gpurun_out/prof_plain $CMD > gpurun_out/ncu_plain.log 2>&1
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD2 > gpurun_out/def_run.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 0 -c 1 -o gpurun_out/prof_def $CMD2 > gpurun_out/ncu_def.log 2>&1
ls -la gpurun_out/*.ncu-rep

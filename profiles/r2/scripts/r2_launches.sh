mkdir -p gpurun_out
T=$1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline $2"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launches.log 2>&1
grep -E "scb_k" gpurun_out/${T}_launches.csv | awk -F'","' '{print $5, $(NF)}' | head -12

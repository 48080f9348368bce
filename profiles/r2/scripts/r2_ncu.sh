# round 2 evidence: launch list of the bench command, then one full capture of K2 (plain launch, resolver not fused: K2's own
# duration and traffic) and of K0.  Run only after the same bench command exited 0 without ncu.
mkdir -p gpurun_out
T=$1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cases"
$CMD > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launches.log 2>&1; echo "launch list rc=$?"
CMD2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cases --opt graph=0 --opt fuse=0"
$CMD2 > gpurun_out/${T}_plain2.json 2> gpurun_out/${T}_plain2.err; echo "plain2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 2 -c 1 -f -o gpurun_out/${T}_k2 $CMD2 > gpurun_out/${T}_ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:scb_k_compile -s 2 -c 1 -f -o gpurun_out/${T}_k0 $CMD2 > gpurun_out/${T}_ncu_k0.log 2>&1; echo "ncu k0 rc=$?"
CMD3="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cases --opt graph=0"
ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 2 -c 1 -f -o gpurun_out/${T}_k2_fused $CMD3 > gpurun_out/${T}_ncu_k2f.log 2>&1; echo "ncu k2 fused rc=$?"
ls -la gpurun_out/${T}_*

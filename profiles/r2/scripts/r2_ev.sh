mkdir -p gpurun_out
T=$1; shift
for W in "$@"; do
SCB_B200_LIB=scbonita_b200/libscbonita_trace.so python profiles/r2/scripts/trace_events.py $W > gpurun_out/${T}_ev_$W.txt 2>&1; echo "ev $W rc=$?"
cat gpurun_out/${T}_ev_$W.txt
done

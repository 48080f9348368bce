mkdir -p gpurun_out
T=$1
SCB_B200_LIB=scbonita_b200/libscbonita_trace.so python profiles/r2/scripts/trace_steps.py C3 > gpurun_out/${T}_steps_c3.txt 2>&1; echo rc=$?
head -70 gpurun_out/${T}_steps_c3.txt | cut -c1-230
SCB_B200_LIB=scbonita_b200/libscbonita_trace.so python profiles/r2/scripts/trace_events.py C2 > gpurun_out/${T}_ev_C2.txt 2>&1; cat gpurun_out/${T}_ev_C2.txt | cut -c1-600

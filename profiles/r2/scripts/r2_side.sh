# side workloads once: level-0 shim timing, C5 attractor analysis, C2
mkdir -p gpurun_out
T=$1
for W in L0 C5 C2; do
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-cases > gpurun_out/${T}_$W.json 2> gpurun_out/${T}_$W.err; echo "$W rc=$?"; tail -2 gpurun_out/${T}_$W.err
  head -c 3000 gpurun_out/${T}_$W.json; echo
done
timeout 600 python -m pytest tests -x -q -m gpu -k "attractor or hamming or general or update or resident or beyond" 2>&1 | tail -3

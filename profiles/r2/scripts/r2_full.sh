# the whole GPU suite + the default bench + the reference arm (what the driver runs at round end)
mkdir -p gpurun_out
T=$1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
python - "$T" <<'PY'
import json,sys
d=json.load(open('gpurun_out/%s_bench.json'%sys.argv[1]))
print('ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'stages',d.get('stages_ms'),'roof',d['roofline'],'cpu',d.get('cpu_baseline'),d.get('cpu_baseline_1thread'))
PY

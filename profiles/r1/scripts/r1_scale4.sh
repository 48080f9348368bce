mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sc4_weak_n1.json 2> gpurun_out/sc4.err
for n in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2968$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sc4_weak_n$n.json 2>> gpurun_out/sc4.err
done
python - <<'PY'
import json
a=json.load(open('gpurun_out/sc4_weak_n1.json'))
for n in (2,4):
    b=json.load(open('gpurun_out/sc4_weak_n%d.json'%n))
    print('n1', round(a['ms_per_step'],3), 'n%d'%n, round(b['ms_per_step'],3), b['value'], 'eff', round(b['value']/(n*a['value']),4))
PY

mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/sc3_gpus.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sc3_weak_n1.json 2> gpurun_out/sc3.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sc3_weak_n8.json 2>> gpurun_out/sc3.err
python - <<'PY'
import json
a=json.load(open('gpurun_out/sc3_weak_n1.json')); b=json.load(open('gpurun_out/sc3_weak_n8.json'))
print('n1', a['ms_per_step'], a['value'], 'n8', b['ms_per_step'], b['value'], 'eff', b['value']/(8*a['value']))
PY

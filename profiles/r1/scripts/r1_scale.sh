# usage: bash profiles/scripts/r1_scale.sh N   (N = number of GPUs of the box)
N=${1:-1}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale_gpus.txt
if [ "$N" = "1" ]; then
  timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
else
  for n in 1 2 4 8; do
    if [ $n -le $N ]; then
      if [ $n = 1 ]; then timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
      else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
    fi
  done
fi

mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/sc_weak_n1.json 2> gpurun_out/sc.err
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 296$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sc_weak_n$n.json 2>> gpurun_out/sc.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 297$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --scaling strong > gpurun_out/sc_strong_n$n.json 2>> gpurun_out/sc.err
done

mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest6.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/m_n1.json 2> gpurun_out/m_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/m_n2.json 2> gpurun_out/m_n2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/m_ref.json 2> gpurun_out/m_ref.err
tail -n 4 gpurun_out/pytest6.log; tail -n 3 gpurun_out/m_n2.err

# strong-scaling check on N GPUs of one box: arg 1 = tag, arg 2 = N, rest = extra bench flags
mkdir -p gpurun_out
T=$1; N=$2; shift; shift
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-cases > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err; echo "n1 rc=$?"
for G in peer nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-cases --gather $G "$@" > gpurun_out/${T}_n${N}_$G.json 2> gpurun_out/${T}_n${N}_$G.err; echo "n$N $G rc=$?"; tail -3 gpurun_out/${T}_n${N}_$G.err
done
python - "$T" "$N" <<'PY'
import json,sys
T,N=sys.argv[1],sys.argv[2]
d1=json.load(open('gpurun_out/%s_n1.json'%T))
print('N=1 ms',round(d1['ms_per_step'],3),'e2e',round(d1['e2e']['ms_per_step'],3))
for G in ('peer','nccl'):
    try:
        d=json.load(open('gpurun_out/%s_n%s_%s.json'%(T,N,G)))
        print('N=%s %s ms'%(N,G),round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'eff',round(d['value']/d1['value']/int(N),3),'per_step',d['per_step_ms'],'check',d.get('multi_gpu_check'),d.get('shard_balance'))
    except Exception as e: print(G,'failed',e)
PY

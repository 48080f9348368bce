import sys
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np
from common import *
from scbonita_b200 import _lib
pr = Problem.synthetic(200, 50000, 500)
outs={}
for name,opts in (("s0",dict(steady=0)),("s1",dict(steady=1)),("s1_notmem",dict(steady=1,tmem=0)),("s1_nosnap",dict(steady=1,snapshots=0)),("s1_noearly",dict(steady=1,early_exit=0,snapshots=0)),("s1_nofuse",dict(steady=1,fuse=0)),("s1_notma",dict(steady=1,tma=0))):
    with pr.simulator(CUDA_LIB, **opts) as sim:
        outs[name] = sim.evaluate_population(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai, mode=1)
        print(name, int(outs[name].sum()), sim.last_stats['steps_executed'], flush=True)
base=outs["s0"]
for k,v in outs.items():
    bad=np.argwhere(v!=base)
    print(k, "differing (ind,node):", len(bad), bad[:12].tolist(), (v-base)[tuple(bad[:12].T)].tolist() if len(bad) else "")

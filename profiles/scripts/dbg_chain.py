import sys
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np
from common import *
lib = CUDA_LIB if len(sys.argv) < 2 or sys.argv[1] == "cuda" else EMU_LIB
for cells in (700, 4096, 9000):
    pr = ring_problem([5, 9], 40, cells)
    ref = pr.oracle()
    for opts in (dict(), dict(early_exit=0, snapshots=0), dict(steady=0), dict(max_ctas=1), dict(words=1), dict(words=2), dict(warps=4),
                 dict(tmem=0), dict(tma=0), dict(fuse=0), dict(tma=0, tmem=0, fuse=0, early_exit=0, snapshots=0)):
        with pr.simulator(lib, **opts) as sim:
            out = sim.evaluate_population(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai, mode=2)[0]
            bad = np.nonzero(out != ref["cellwise"][0])[0]
            words = sorted(set((bad // 32).tolist()))
            print(cells, opts, "ok" if len(bad) == 0 else ("bad cells", len(bad), "words", words[:40], "diff", (out - ref["cellwise"][0])[bad][:10].tolist()),
                  sim.info()["words_per_lane"], sim.info()["warps"], sim.last_stats["steps_executed"], sim.last_stats["resolver_chunks"], flush=True)

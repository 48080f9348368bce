import sys
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np
from common import *
lib = CUDA_LIB if len(sys.argv) < 2 or sys.argv[1] == "cuda" else EMU_LIB
keys = {0: "fitness", 1: "nodewise", 2: "cellwise"}
for cells in (700, 9000):
    pr = ring_problem([5, 9], 40, cells)
    ref = pr.oracle()
    for opts in (dict(), dict(steady=0), dict(early_exit=0, snapshots=0), dict(fuse=0)):
        for seq in ((0,), (1,), (2,), (0, 1, 2), (2, 0), (1, 1)):
            with pr.simulator(lib, **opts) as sim:
                res = []
                for mode in seq:
                    out = sim.evaluate_population(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai, mode=mode)
                    ok = np.array_equal(out, ref[keys[mode]])
                    extra = ""
                    if not ok and mode == 1:
                        bad = np.nonzero(out[0] != ref["nodewise"][0])[0]
                        extra = " nodes %s diff %s" % (bad[:12].tolist(), (out[0] - ref["nodewise"][0])[bad][:12].tolist())
                    if not ok and mode == 0:
                        extra = " diff %d" % int(out[0] - ref["fitness"][0])
                    res.append(("ok" if ok else "BAD") + extra)
                print(cells, opts, seq, res, flush=True)

import sys, ctypes
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np
from common import *
from scbonita_b200 import _lib
which = sys.argv[1]
lib_path = EMU_LIB if which == "emu" else CUDA_LIB
pr = ring_problem([5, 9], 40, 700)
ref = pr.oracle()
with pr.simulator(lib_path) as sim:
    out = sim.evaluate_population(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai, mode=1)
    print("nodewise ok:", np.array_equal(out, ref["nodewise"]), np.argwhere(out != ref["nodewise"])[:10].tolist())
    n = pr.n
    d1 = np.zeros((n, 4), np.uint32); d2 = np.zeros((n, 4), np.uint32)
    c1 = np.zeros(4, np.int32); c2 = np.zeros(4, np.int32); dep = np.zeros(1, np.int32)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    sim.lib.scb_debug_programs.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5
    rc = sim.lib.scb_debug_programs(sim._ctx, 0, P(d1), P(c1), P(d2), P(c2), P(dep))
    print("rc", rc, "counts", c1.tolist(), "counts2", c2.tolist(), "depth", dep.tolist())
    np.save("gpurun_out/prog_%s.npy" % which, np.concatenate([d1.ravel(), d2.ravel(), c1.view(np.uint32), c2.view(np.uint32), dep.view(np.uint32)]))
    for j in range(n):
        print(j, [int(x) for x in d2[j][:3]], int(d2[j][3] & 0xFFFFFF), hex(int(d2[j][3] >> 24)))

"""Phase clocks of K0's CTA 0 (measurement build, see trace_steps.py): SM clocks between the marks of scb_k_compile."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from scbonita_b200 import _lib, simulator
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
pr, bm = bench.build_problem(wl)
model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv, pr.node_positions, pr.net.ind_len)
sim = simulator.Simulator(bm, model)
sim.set_option(_lib.OPT_GRAPH, 0)
for _ in range(2):
    sim.evaluate_population(pr.individuals, pr.knockouts, pr.knockins, pr.pop_and_nodes, pr.pop_and_inv)
tr = np.zeros(100 * 24 * 4, np.uint32); fl = np.zeros(204, np.uint32)
sim.lib.scb_debug_trace.argtypes = [ctypes.c_void_p] * 3
sim.lib.scb_debug_trace(sim._ctx, ctypes.c_void_p(tr.ctypes.data), ctypes.c_void_p(fl.ctypes.data))
t = tr[:40].astype(np.int64)
d = lambda a, b: int((t[b] - t[a]) & 0xffffffff)
print("threads", t[7], "depth at the fixed point", t[6])
print("stage tables", d(0, 1), "compile nodes", d(1, 2), "sort full", d(2, 3), "layout full", d(3, 4), "total", d(0, 5))
prev = 4
for slot in range(5):
    a, b, c = 8 + 4 * slot, 9 + 4 * slot, 10 + 4 * slot
    if t[a]:
        print("program", slot, "propagate+restrict since previous", d(prev, a), "sort", d(a, b), "layout", d(b, c))
        prev = c
print("tail", d(prev, 5))
print(sim.last_stats)

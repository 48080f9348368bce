"""Step timeline of one tile (measurement build: make -C scbonita_b200/csrc EXTRA=-DSCB_TRACE OUT=../libscbonita_trace.so
OBJDIR=_obj_trace; run with SCB_B200_LIB=scbonita_b200/libscbonita_trace.so).  Prints, per step of CTA 0's first tile,
how long the slowest and the average warp took and how long the step lasted from the first start to the last end."""
import ctypes, json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from scbonita_b200 import _lib, simulator, synth
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
pr, bm = bench.build_problem(wl)
model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv, pr.node_positions, pr.net.ind_len)
sim = simulator.Simulator(bm, model)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    sim.set_option(getattr(_lib, "OPT_" + k.upper()), int(v))
for _ in range(2):
    sim.evaluate_population(pr.individuals, pr.knockouts, pr.knockins, pr.pop_and_nodes, pr.pop_and_inv)
tr = np.zeros(100 * 24 * 4, np.uint32); fl = np.zeros(204, np.uint32)
sim.lib.scb_debug_trace.argtypes = [ctypes.c_void_p] * 3
sim.lib.scb_debug_trace(sim._ctx, ctypes.c_void_p(tr.ctypes.data), ctypes.c_void_p(fl.ctypes.data))
tr = tr.reshape(100, 24, 4).astype(np.int64)
S = sim.info()["warps"]
prev_end = None
rows = []
for t in range(1, 100):
    a, b = tr[t, :S, 0], tr[t, :S, 1]
    if not a.any():
        continue
    d = (b - a) & 0xffffffff
    start, end = a.min(), b.max()
    gap = None if prev_end is None else int((start - prev_end) & 0xffffffff)
    rows.append(dict(step=t, flags=int(fl[t]) & 0xffff, rows=int(fl[t]) >> 16, warp_min=int(d.min()), warp_mean=float(d.mean()), warp_max=int(d.max()),
                     span=int((end - start) & 0xffffffff), gap_from_prev=gap))
    if tr[t, 0, 2]:
        rows[-1]["after_bar1"] = int((tr[t, :S, 2].max() - end) & 0xffffffff)
        rows[-1]["after_bar2"] = int((tr[t, :S, 3].max() - end) & 0xffffffff)
    prev_end = end
for r in rows:
    print(r)
print('exit-step histogram', {t: int(fl[100 + t]) for t in range(100) if fl[100 + t]})
print('exits: E2 %d  E3 %d  all-proven %d  to-resolver %d' % tuple(int(x) for x in fl[200:204]))
print(json.dumps(dict(info=sim.info(), stats=sim.last_stats)))

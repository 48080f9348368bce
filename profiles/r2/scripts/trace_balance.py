"""Per-warp work against per-warp time in the plain steps of one tile (measurement build, see trace_steps.py): which
static quantity of a warp's piece of the program (rows, nodes, runs, one-node-loop nodes) predicts how long the warp
takes -- input for the cost model K0 cuts the pieces with."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from scbonita_b200 import _lib, simulator
import bench

UNROLL2 = {0xF0, 0x0F, 0xC0, 0xFC}
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
pr, bm = bench.build_problem(wl)
model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv, pr.node_positions, pr.net.ind_len)
sim = simulator.Simulator(bm, model)
for _ in range(2):
    sim.evaluate_population(pr.individuals, pr.knockouts, pr.knockins, pr.pop_and_nodes, pr.pop_and_inv)
tr = np.zeros(100 * 24 * 4, np.uint32); fl = np.zeros(204, np.uint32)
sim.lib.scb_debug_trace.argtypes = [ctypes.c_void_p] * 3
sim.lib.scb_debug_trace(sim._ctx, ctypes.c_void_p(tr.ctypes.data), ctypes.c_void_p(fl.ctypes.data))
tr = tr.reshape(100, 24, 4).astype(np.int64)
S = sim.info()["warps"]
# CTA 0's first tile is ticket 0: individual 0 (no order set), tile 0
X, Y = [], []
for k in range(5):
    pg = sim.programs(0, k)
    frm = pg["from_k"]
    if not frm:
        continue
    nxt = [sim.programs(0, j)["from_k"] for j in range(k + 1, 5)]
    nxt = [f for f in nxt if f > frm]
    last = (min(nxt) if nxt else 24) - 1
    d = pg["descs_k"]; c3, c2, c1, c0 = (int(x) for x in pg["counts_k"])
    e1 = c3 + c2 + c1
    d = d[:e1]
    cum, ar, run, lut = d[:, 0] >> 16, (d[:, 1] >> 16) & 3, d[:, 2] >> 16, d[:, 3] >> 24
    total = 4 * c3 + 3 * c2 + 2 * c1
    warp = (cum.astype(np.int64) * S) // total
    steps = [t for t in range(frm, last + 1) if tr[t, :S, 0].any() and (fl[t] & 0xffff) == 0]
    if not steps:
        continue
    times = np.mean([(tr[t, :S, 1] - tr[t, :S, 0]) & 0xffffffff for t in steps], axis=0)
    for w in range(S):
        m = warp == w
        nodes = int(m.sum()); rows = int((ar[m] + 1).sum())
        runs = len(set(zip(run[m].tolist(), lut[m].tolist())))
        one = int(sum(1 for v in lut[m] if int(v) not in UNROLL2))
        X.append([rows, nodes, runs, one, 1.0]); Y.append(times[w])
        print("program", k, "steps", steps[0], "-", steps[-1], "warp", w, "rows", rows, "nodes", nodes, "runs", runs, "one-node-loop nodes", one, "clk %.0f" % times[w])
X, Y = np.array(X, float), np.array(Y, float)
for cols, names in (((0, 4), "rows"), ((0, 2, 4), "rows, runs"), ((0, 1, 2, 4), "rows, nodes, runs"), ((0, 1, 2, 3, 4), "rows, nodes, runs, one-node")):
    coef, res, *_ = np.linalg.lstsq(X[:, cols], Y, rcond=None)
    pred = X[:, cols] @ coef
    print("fit clk ~ %s + 1:" % names, np.round(coef, 1), "rms error %.0f of mean %.0f" % (np.sqrt(np.mean((pred - Y) ** 2)), Y.mean()))

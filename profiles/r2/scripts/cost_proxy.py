"""How well do quantities K0 knows at compile time predict an individual's simulation cost (for ordering a population
that has never been scored)?  Spearman correlation of the measured cost with rows of the fixed-point program, rows of
the full program, propagation depth."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from scbonita_b200 import _lib, simulator
import bench

def spearman(a, b):
    ra, rb = np.argsort(np.argsort(a)), np.argsort(np.argsort(b))
    return float(np.corrcoef(ra, rb)[0, 1])

for seed_shift in (0, 1000):
    wl = bench.WORKLOADS["C3"]
    pr, bm = bench.build_problem(wl, seed_shift=seed_shift)
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv, pr.node_positions, pr.net.ind_len)
    sim = simulator.Simulator(bm, model)
    sim.evaluate_population(pr.individuals, pr.knockouts, pr.knockins, pr.pop_and_nodes, pr.pop_and_inv)
    total, mx = sim.costs(tile_max=True)
    P = len(total)
    rows_fp, rows_full, depth, live = np.zeros(P), np.zeros(P), np.zeros(P), np.zeros(P)
    for p in range(P):
        pg = sim.programs(p, 4)
        c3, c2, c1, c0 = (int(x) for x in pg["counts_k"])
        rows_fp[p] = 4 * c3 + 3 * c2 + 2 * c1; live[p] = c3 + c2 + c1
        f3, f2, f1, f0 = (int(x) for x in pg["counts"])
        rows_full[p] = 4 * f3 + 3 * f2 + 2 * f1
        depth[p] = pg["from_k"]
    print("population seed shift", seed_shift)
    for name, x in (("rows of the fixed-point program", rows_fp), ("live nodes at the fixed point", live), ("rows of the full program", rows_full),
                    ("fixed-point depth", depth), ("rows_fp * depth", rows_fp * depth)):
        print("  %-34s spearman vs total cost %.3f, vs costliest tile %.3f" % (name, spearman(x, total), spearman(x, mx)))
    sim.close()

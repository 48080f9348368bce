"""Times the SURVEY 8(f) rows next to the GA generation on one GPU (results are parity-checked in tests/, this is
the measurement): N3 local search of every node in one launch, N2 context creation from CSR vs dense, N1 batched
importance scores, and BASELINE config C5 (attractor windows + Hamming decider, 100k cells x 300 nodes).
Prints one JSON object.  Usage: python profiles/scripts/next_rows_timing.py [small]"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from scbonita_b200 import simulator, synth  # noqa: E402


def problem(n, cells):
    pr = synth.make_problem(n, cells, 1)
    pos = np.asarray(pr.node_positions)
    bm = np.zeros((int(pos.max()) + 1, cells), np.uint8)
    bm[pos] = pr.bin_rows
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv, pos, pr.net.ind_len)
    return pr, bm, model


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def main():
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    n, cells = (40, 3000) if small else (200, 50000)
    res = {"workload": "%d nodes x %d cells" % (n, cells)}
    pr, bm, model = problem(n, cells)
    ind = pr.individuals[0]
    # N2: context creation, dense uint8 vs scipy CSR (what the reference stores)
    t_dense, _ = timed(lambda: simulator.Simulator(bm, model).close())
    csr = sp.csr_matrix(bm.astype(np.float64))
    t_csr, _ = timed(lambda: simulator.Simulator(csr, model).close())
    res["N2_ctx_create_ms"] = {"dense_uint8": t_dense * 1e3, "csr": t_csr * 1e3, "nnz": int(csr.nnz)}
    with simulator.Simulator(bm, model) as sim:
        # N3: all 2^k - 1 variants of every node (the loop of singleCell.__scoreNodes)
        t_ls, out = timed(lambda: sim.local_search(ind, [0], [0]), reps=2)
        variants = sum(len(o[3]) for o in out if not isinstance(o[3], float))
        res["N3_local_search_all_nodes"] = {"ms": t_ls * 1e3, "variants": int(variants),
                                            "cre": float(variants) * cells * n * 99,
                                            "cre_per_s": float(variants) * cells * n * 99 / t_ls}
        # N1: importance scores of every node
        t_is, sc = timed(lambda: sim.importance_scores(ind, [[i] for i in range(n)], strict=False))
        res["N1_importance_all_nodes"] = {"ms": t_is * 1e3, "items": n, "nan_items": int(np.isnan(sc).sum())}
    # C5: attractor analysis
    n5, c5 = (60, 5000) if small else (300, 100000)
    pr5, bm5, model5 = problem(n5, c5)
    with simulator.Simulator(bm5, model5) as sim:
        t_at, attrs = timed(lambda: sim.attractor_set(pr5.individuals[0]), reps=1)
        t_traj, _ = timed(lambda: sim.attractors(pr5.individuals[0], semantics="python", max_states=10), reps=2)
        t_c, _ = timed(lambda: sim.attractors(pr5.individuals[0], semantics="c", max_states=1), reps=2)
        if len(attrs) == 0:
            attrs = np.ones((1, n5), np.int32)
        attrs = attrs[:2000]
        t_h, _ = timed(lambda: sim.hamming_argmin(attrs, want_dist=False), reps=3)
        res["C5_attractors"] = {"workload": "%d nodes x %d cells" % (n5, c5),
                                "python_semantics_windows_ms": t_traj * 1e3, "c_semantics_windows_ms": t_c * 1e3,
                                "attractor_set_incl_host_dedupe_ms": t_at * 1e3, "unique_attractors_used": int(len(attrs)),
                                "hamming_argmin_ms": t_h * 1e3,
                                "bit_compares_per_s": float(c5) * len(attrs) * n5 / t_h}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

"""Times the batched importance score (scb_importance_scores: all N knock-out/knock-in pairs of
ruleMaker.__calcImportance in one simulator launch) against N single scb_importance_score calls on the same
context, checks that both give the same doubles, and prints one JSON line.  Usage:
    python profiles/scripts/importance_timing.py [nodes cells [single_calls]]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from scbonita_b200 import simulator, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
    singles = int(sys.argv[3]) if len(sys.argv) > 3 else min(n, 40)
    pr = synth.make_problem(n, cells, 1, knock=False)
    pos = np.asarray(pr.node_positions)
    bm = np.zeros((int(pos.max()) + 1, cells), np.int32)
    bm[pos] = pr.bin_rows
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_and_nodes[0], pr.pop_and_inv[0], pos, pr.net.ind_len)
    ind = pr.individuals[0]
    lists = [[i] for i in range(n)]
    with simulator.Simulator(bm, model) as sim:
        sim.importance_scores(ind, lists, strict=False)                       # warm-up
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            got = sim.importance_scores(ind, lists, strict=False)
        t_batch = (time.perf_counter() - t0) / reps
        kernel_ms = sim.last_stats.get("kernel_ms") if hasattr(sim, "last_stats") and sim.last_stats else None
        t0 = time.perf_counter()
        one = []
        for i in range(singles):
            try:
                one.append(sim.evaluate_by_node(ind, [i], [i], importanceScores=True)[0])
            except Exception:
                one.append(float("nan"))
        t_single = (time.perf_counter() - t0) / singles
    one = np.asarray(one)
    same = bool(np.array_equal(np.isnan(one), np.isnan(got[:singles])) and
                one[~np.isnan(one)].tobytes() == got[:singles][~np.isnan(got[:singles])].tobytes())
    # one importanceScore call simulates 2 passes x cells x nodes x 99 steps (simulator.c:585-720)
    cre = 2.0 * cells * n * 99
    print(json.dumps({"workload": "importance scores, %d nodes x %d cells, %d items" % (n, cells, n),
                      "batched_ms_total": t_batch * 1e3, "batched_ms_per_item": t_batch * 1e3 / n,
                      "single_call_ms": t_single * 1e3, "speedup_vs_single_calls": t_single * n / t_batch,
                      "items_nan": int(np.isnan(got).sum()), "single_matches_batched": same,
                      "reference_cre_per_item": cre, "batched_cre_per_s": cre * n / t_batch,
                      "scores_head": [None if np.isnan(x) else float(x) for x in got[:4]], "kernel_ms": kernel_ms}))


if __name__ == "__main__":
    main()

"""BASELINE.json-size runs on the GPU: config C3 (50 000 cells x 200 nodes x 500 individuals) against the
bit-packed oracle in full, plus size-independent properties; config C2; config C5's decider stage."""
import numpy as np
import pytest

import oracle
from common import Problem
from scbonita_b200 import simulator

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    return Problem.synthetic(200, 50_000, 500)


def test_c3_full_population_matches_oracle(cuda_lib, c3):
    want = c3.oracle()          # ~15-25 s on the host cores
    with c3.simulator(cuda_lib) as sim:
        fit = sim.evaluate_population(c3.individuals, c3.ko, c3.ki, c3.pop_an, c3.pop_ai, mode=simulator.MODE_SUM)
        assert np.array_equal(fit, want["fitness"])
        assert sim.last_stats["not_found_cells"] == int(want["not_found"].sum())
        nw = sim.evaluate_population(c3.individuals, c3.ko, c3.ki, c3.pop_an, c3.pop_ai, mode=simulator.MODE_NODEWISE)
        assert np.array_equal(nw, want["nodewise"])
        # size-independent property: both reductions of the same error bits agree
        assert np.array_equal(nw.sum(axis=1), fit)


def test_c3_cellwise_slice_and_schedule_independence(cuda_lib, c3):
    """Cellwise vectors (50 000 ints per individual) for a slice of the population, with and without the
    early-exit / snapshot machinery: identical integers, and their sums equal the SUM-mode fitness."""
    want = c3.oracle()
    sl = slice(40, 72)
    args = (c3.individuals[sl], c3.ko, c3.ki, c3.pop_an[sl], c3.pop_ai[sl])
    with c3.simulator(cuda_lib) as sim:
        cw = sim.evaluate_population(*args, mode=simulator.MODE_CELLWISE)
    with c3.simulator(cuda_lib, early_exit=0, snapshots=0) as sim:
        cw_plain = sim.evaluate_population(*args, mode=simulator.MODE_CELLWISE)
    assert np.array_equal(cw, want["cellwise"][sl])
    assert np.array_equal(cw, cw_plain)
    assert np.array_equal(cw.sum(axis=1), want["fitness"][sl])


def test_c2_matches_oracle(cuda_lib):
    pr = Problem.synthetic(60, 10_000, 100)
    with pr.simulator(cuda_lib) as sim:
        pr.check(sim)


def test_c5_decider_stage(cuda_lib):
    """100 000 cells x 300 nodes mapped to attractors found on the device."""
    pr = Problem.synthetic(300, 100_000, 1, knock=False)
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_an[0], pr.pop_ai[0], pr.node_positions, pr.net.ind_len)
    with simulator.Simulator(pr.bm, model, lib_path=cuda_lib) as sim:
        attrs = sim.attractor_set(pr.individuals[0])
        assert len(attrs) >= 1
        attrs = attrs[:64]
        dist, dec, dd = sim.hamming_argmin(attrs)
    wd, wdec, wdd = oracle.hamming_argmin(attrs, pr.bm, pr.node_positions, 100_000)
    assert np.array_equal(dist, wd) and np.array_equal(dec, wdec) and np.array_equal(dd, wdd)
    # property: the decider is a first minimum of its row
    assert np.array_equal(dist[np.arange(len(dec)), dec], dist.min(axis=1))


def test_c5_trajectories_windows_and_unique_states(cuda_lib):
    """K4 at C5 size (100 000 cells x 300 nodes, Python semantics): every window and every window state against the
    restated singleCell.__cluster, and the device-side dedupe against the first-seen order of those states."""
    from common import oracle_model
    C = 100_000
    pr = Problem.synthetic(300, C, 1, knock=False)
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_an[0], pr.pop_ai[0], pr.node_positions, pr.net.ind_len)
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
    m.and_nodes, m.and_inv = pr.pop_an[0], pr.pop_ai[0]
    with simulator.Simulator(pr.bm, model, lib_path=cuda_lib) as sim:
        res, nst, bits = sim.attractors(pr.individuals[0], semantics="python", max_states=10)
        uniq = sim.attractor_set(pr.individuals[0], remove_zero=False)
    seen, order = set(), []
    step = 20_000                                               # the restatement's int32 trajectory: 240 MB per slice
    for lo in range(0, C, step):
        vals, wres, _ = oracle.py_cluster(m, pr.individuals[0], np.ascontiguousarray(pr.bm[:, lo:lo + step]), step)
        assert np.array_equal(res[lo:lo + step], wres)
        first = np.where(wres[:, 0] >= 0, wres[:, 0], 9)
        last = np.where(wres[:, 0] >= 0, wres[:, 1], 9)
        assert np.array_equal(nst[lo:lo + step], last - first + 1)
        for s in range(10):                                     # state s of every window that has one
            has = first + s <= last
            idx = np.nonzero(has)[0]
            assert np.array_equal(bits[lo + idx, s], vals[idx, first[idx] + s].astype(np.uint8))
        packed = np.packbits(vals.astype(np.uint8), axis=2)
        for c in range(step):
            for r in range(first[c], last[c] + 1):
                k = packed[c, r].tobytes()
                if k not in seen:
                    seen.add(k); order.append((lo + c, r - first[c]))
    assert uniq.shape[0] == len(order)
    want = np.stack([bits[c, s] for c, s in order]).astype(np.int32)
    assert np.array_equal(uniq, want)


def test_two_gpu_sharding_nccl(cuda_lib):
    """individuals sharded over 2 ranks + NCCL all-gather (skipped on a one-GPU box; the gloo test covers the logic)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import subprocess
    import sys
    from common import ROOT
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tests", "nccl_worker.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_c4_independent_pathways(cuda_lib):
    """BASELINE config 4 in miniature: the same ~60 000 cells scored against several KEGG-shaped pathways of
    different sizes, one context per pathway (independent work: pathways map to GPUs round-robin, no collective)."""
    for n in (47, 60, 100, 159, 200, 300):
        pr = Problem.synthetic(n, 60_000, 24)
        with pr.simulator(cuda_lib) as sim:
            fit = sim.evaluate_population(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai, mode=simulator.MODE_SUM)
        assert np.array_equal(fit, pr.oracle()["fitness"]), n

"""The oracle (CPU restatement) against the UNMODIFIED reference C and the committed golden vectors."""
import os

import numpy as np
import pytest

import oracle
from common import GOLDEN, Problem, oracle_model, ring_problem

HAVE_REF = oracle.ref_available(15000)


def _ref_eval(pr, p, local_search):
    ref = oracle.RefSimulator(pr.n_cells)
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
    bm = ref.dense(pr.bin_rows.astype(np.int32), pr.node_positions)
    an = None if pr.pop_an is None else pr.pop_an[p]
    ai = None if pr.pop_ai is None else pr.pop_ai[p]
    out = ref.sync_bool(m, pr.individuals[p], bm, pr.n_cells, local_search, and_nodes=an, and_inv=ai)
    return out[:pr.n] if local_search else out[:pr.n_cells]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n,cells,pop,knock", [(60, 700, 4, True), (47, 100, 3, False), (200, 150, 2, True), (9, 33, 5, True)])
def test_restatements_match_reference_build(n, cells, pop, knock):
    pr = Problem.synthetic(n, cells, pop, knock=knock)
    o = pr.oracle()
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
    for p in range(pop):
        cell_ref, node_ref = _ref_eval(pr, p, 0), _ref_eval(pr, p, 1)
        assert np.array_equal(o["cellwise"][p], cell_ref)
        assert np.array_equal(o["nodewise"][p], node_ref)
        assert o["fitness"][p] == cell_ref.sum()
        e, _ = oracle.sync_bool(m, pr.individuals[p], pr.bm, cells, 0, and_nodes=pr.pop_an[p], and_inv=pr.pop_ai[p])
        assert np.array_equal(e, cell_ref)
        e, _ = oracle.sync_bool(m, pr.individuals[p], pr.bm, cells, 1, and_nodes=pr.pop_an[p], and_inv=pr.pop_ai[p])
        assert np.array_equal(e, node_ref)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("rings,chain,cells", [([60], 0, 300), ([45], 12, 1500), ([49], 0, 100), ([50], 0, 100), ([3, 5], 4, 400)])
def test_not_found_fill_forward_matches_reference(rings, chain, cells):
    """cells whose state 99 repeats no earlier state inherit the previous cell's error."""
    pr = ring_problem(rings, chain, cells)
    o = pr.oracle()
    assert np.array_equal(o["cellwise"][0], _ref_eval(pr, 0, 0))
    assert np.array_equal(o["nodewise"][0], _ref_eval(pr, 0, 1))


def test_golden_vectors():
    """Vectors generated from the unmodified reference build by tests/golden/make_golden.py."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.startswith("sim_") and f.endswith(".npz"))
    assert files, "no golden vectors committed"
    for f in files:
        g = np.load(os.path.join(GOLDEN, f))
        m = oracle.Model(g["and_len"], g["parse"], g["and_nodes"], g["and_inv"], g["node_positions"],
                         int(g["ind_len"]), g["knockouts"], g["knockins"])
        C = int(g["n_cells"])
        bm = np.zeros((int(g["node_positions"].max()) + 1, C), np.int32)
        bm[g["node_positions"]] = np.unpackbits(g["bin_rows"], axis=1)[:, :C]
        per_ind = g["pop_and_nodes"].ndim == 4
        r = oracle.packed_population(m, g["individuals"], bm, C, and_nodes=g["pop_and_nodes"] if per_ind else None,
                                     and_inv=g["pop_and_inv"] if per_ind else None, want_cellwise=True)
        assert np.array_equal(r["cellwise"], g["ref_cellwise"]), f
        assert np.array_equal(r["nodewise"], g["ref_nodewise"]), f


def test_hamming_known_answer_from_reference_package():
    """The reference's packaged attractorDistance.csv (100 cells x 51 attractors) is a bit-exact KAT for
    the distance / decider stage; committed as tests/golden/hamming_kat.npz."""
    g = np.load(os.path.join(GOLDEN, "hamming_kat.npz"))
    dist, dec, dd = oracle.hamming_argmin(g["attractors"], g["bin_mat"], g["node_positions"], int(g["n_cells"]))
    assert np.array_equal(dist, g["dist"])
    assert np.array_equal(dec, g["decider"])
    assert np.array_equal(dd, g["decider_distance"])


# ---- importanceScore and cluster restatements against the unmodified reference build -------------------------
def _knock_model(pr, ko_nodes, ki_nodes, p=0):
    ko = np.zeros(pr.n, np.int32); ko[list(ko_nodes)] = 1
    ki = np.zeros(pr.n, np.int32); ki[list(ki_nodes)] = 1
    m = oracle_model(pr.net, pr.node_positions, ko, ki)
    if pr.pop_an is not None:
        m.and_nodes, m.and_inv = oracle._i32(pr.pop_an[p]), oracle._i32(pr.pop_ai[p])
    return m


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n,cells,seed", [(12, 30, None), (20, 200, 4242), (47, 100, 7), (61, 300, 11), (9, 64, 99)])
def test_importance_restatement_matches_reference_build(n, cells, seed):
    """scbo_importance_score (simulator.c:555-755 restated) against the reference's own importanceScore: every
    double bit-identical, for knock-out == knock-in lists (what __calcImportance passes, ruleMaker.py:1285-1300)
    and for distinct ones.  Cases where a cell has no attractor window are skipped on the reference side: the
    reference divides by zero there (SIGFPE, :647); the restatement reports them (ValueError)."""
    pr = Problem.synthetic(n, cells, 1, knock=False, net_seed=seed)
    ref = oracle.RefSimulator(cells)
    bm = ref.dense(pr.bin_rows.astype(np.int32), pr.node_positions)
    rng = np.random.default_rng(5 if seed is None else seed)
    cases = [([i], [i]) for i in range(min(n, 6))] + [([1 % n], [n - 1]), ([], [])]
    cases += [(list(rng.choice(n, 2, replace=False)), list(rng.choice(n, 3, replace=False))) for _ in range(3)]
    compared = 0
    for ko, ki in cases:
        m = _knock_model(pr, ko, ki)
        try:
            want = oracle.importance_score(m, pr.individuals[0], pr.bm, cells)
        except ValueError:
            continue
        got = ref.importance_score(m, pr.individuals[0], bm, cells)
        assert np.float64(got).tobytes() == np.float64(want).tobytes(), (ko, ki, got, want)
        compared += 1
    assert compared >= 3


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("rings,cells", [([40, 5], 60), ([3, 5], 80)])
def test_importance_restatement_on_rings(rings, cells):
    """Long cycles: windows of length > 1 add (int)(state / length) = 0 (:647)."""
    pr = ring_problem(rings, 3, cells)
    ref = oracle.RefSimulator(cells)
    bm = ref.dense(pr.bin_rows.astype(np.int32), pr.node_positions)
    compared = 0
    for node in range(0, pr.n, 5):
        m = _knock_model(pr, [node], [node])
        try:
            want = oracle.importance_score(m, pr.individuals[0], pr.bm, cells)
        except ValueError:
            continue
        got = ref.importance_score(m, pr.individuals[0], bm, cells)
        assert np.float64(got).tobytes() == np.float64(want).tobytes(), (node, got, want)
        compared += 1
    assert compared >= 1


def _cluster_cases(pr):
    """knock lists for `cluster`: none, knock-in only, knock-outs (the branch with the column slip, :518)."""
    n = pr.n
    return [([], []), ([], [n // 2]), ([2 % n], []), ([0, n - 1], [1 % n]), ([n // 3, (2 * n) // 3], [])]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n,cells,seed,identity", [(20, 40, None, False), (33, 25, 31, False), (12, 30, 8, True), (47, 20, 3, False)])
def test_cluster_restatement_matches_reference_build(n, cells, seed, identity):
    """scbo_cluster_literal against the reference's `cluster` (simulator.c:487-552): the attractor window AND the
    whole scratch the call leaves behind, with knock-outs on permuted nodePositions (where :518 records the
    knocked node in column i instead of nodePositions[i]) and on a scratch that is not zero."""
    pr = Problem.synthetic(n, cells, 1, knock=False, net_seed=seed)
    if identity:
        pr = Problem(pr.net, pr.bin_rows, np.arange(n, dtype=np.int32), pr.individuals, pr.pop_an, pr.pop_ai)
    ref = oracle.RefSimulator(cells)
    bm = ref.dense(pr.bin_rows.astype(np.int32), pr.node_positions)
    stride = max(int(pr.node_positions.max()) + 1, n)
    rng = np.random.default_rng(17)
    dirty = rng.integers(0, 2, (100, stride)).astype(np.int32)
    for ko, ki in _cluster_cases(pr):
        m = _knock_model(pr, ko, ki)
        for cell in (0, cells // 2, cells - 1):
            for scratch in (None, dirty):
                vals, res = ref.cluster(m, pr.individuals[0], bm, cell, sim_data=scratch)
                sd, wres = oracle.c_cluster_literal(m, pr.individuals[0], pr.bm, cell, sim_data=scratch, sim_stride=stride)
                assert tuple(res) == tuple(wres), (ko, ki, cell, res, wres)
                assert np.array_equal(vals[:, :stride], sd), (ko, ki, cell)
                assert not vals[:, stride:].any()
        if not ko:
            # without knock-outs the scratch holds the true trajectory: the node-order restatement agrees too
            traj, tres = oracle.c_cluster(m, pr.individuals[0], pr.bm, 0)
            vals, res = ref.cluster(m, pr.individuals[0], bm, 0)
            assert tuple(res) == tuple(tres) and np.array_equal(vals[:, pr.node_positions], traj)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
def test_cluster_restatement_on_rings():
    pr = ring_problem([7, 4], 3, 50)
    ref = oracle.RefSimulator(50)
    bm = ref.dense(pr.bin_rows.astype(np.int32), pr.node_positions)
    stride = max(int(pr.node_positions.max()) + 1, pr.n)
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)       # the chain head is knocked out
    for cell in range(0, 50, 7):
        vals, res = ref.cluster(m, pr.individuals[0], bm, cell)
        sd, wres = oracle.c_cluster_literal(m, pr.individuals[0], pr.bm, cell, sim_stride=stride)
        assert tuple(res) == tuple(wres) and np.array_equal(vals[:, :stride], sd), cell


def test_importance_and_cluster_golden_vectors():
    """imp_*.npz / cluster_*.npz: outputs of the unmodified reference build (tests/golden/make_golden.py)."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.startswith(("imp_", "cluster_")) and f.endswith(".npz"))
    assert any(f.startswith("imp_") for f in files) and any(f.startswith("cluster_") for f in files)
    for f in files:
        g = np.load(os.path.join(GOLDEN, f))
        C = int(g["n_cells"])
        bm = np.zeros((int(g["node_positions"].max()) + 1, C), np.int32)
        bm[g["node_positions"]] = np.unpackbits(g["bin_rows"], axis=1)[:, :C]
        for k in range(len(g["knockouts"])):
            m = oracle.Model(g["and_len"], g["parse"], g["and_nodes"], g["and_inv"], g["node_positions"],
                             int(g["ind_len"]), g["knockouts"][k], g["knockins"][k])
            if f.startswith("imp_"):
                got = oracle.importance_score(m, g["individual"], bm, C)
                assert np.float64(got).tobytes() == g["ref_scores"][k].tobytes(), (f, k)
            else:
                stride = int(g["stride"])
                for ci, cell in enumerate(g["cells"]):
                    sd, res = oracle.c_cluster_literal(m, g["individual"], bm, int(cell), sim_stride=stride)
                    assert tuple(res) == tuple(g["ref_res"][k, ci]), (f, k, cell)
                    assert np.array_equal(sd.astype(np.int8), g["ref_sim"][k, ci]), (f, k, cell)

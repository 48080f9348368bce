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

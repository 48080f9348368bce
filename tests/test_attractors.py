"""Attractor companion: trajectories + windows (K4) against the oracle and against vectors produced by the
reference's own Python (singleCell.__cluster) and C (`cluster`)."""
import ctypes
import os

import numpy as np
import pytest

import oracle
from common import GOLDEN, Problem, oracle_model, ring_problem, tables
from scbonita_b200 import _lib, simulator

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def backend(request):
    if request.param == "emu":
        return "emu", request.getfixturevalue("emu_lib")
    return "cuda", request.getfixturevalue("cuda_lib")


def _sim(pr, backend):
    sim = simulator.Simulator(pr.bm, tables(pr.net, pr.node_positions), lib_path=backend[1])
    if backend[0] == "emu":
        sim.set_option(_lib.OPT_MAX_CTAS, 3)
    return sim


@pytest.mark.parametrize("name", ["pycluster_n24.npz", "pycluster_n40.npz"])
def test_python_semantics_against_reference_python(backend, name):
    """Golden vectors come from the reference's own singleCell.__cluster (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, name))
    C = int(g["n_cells"]); pos = g["node_positions"]; n = len(pos)
    bm = np.zeros((n, C), np.int32)
    bm[pos] = np.unpackbits(g["bin_rows"], axis=1)[:, :C]
    model = simulator.ModelTables(g["and_len"], g["parse"], g["and_nodes"], g["and_inv"], pos, int(g["ind_len"]))
    with simulator.Simulator(bm, model, lib_path=backend[1]) as sim:
        res, nst, bits = sim.attractors(g["individual"], g["knockouts"], g["knockins"], "python", 10)
    assert np.array_equal(res, g["ref_res"])
    vals = g["ref_vals"]
    for c in range(C):
        r0, r1 = g["ref_res"][c]
        rows = list(range(r0, r1 + 1)) if r0 >= 0 else [9]          # range(-1, 0) -> vals[-1]
        assert nst[c] == len(rows)
        assert np.array_equal(bits[c, :len(rows)], vals[c, rows].astype(np.uint8)), c


@pytest.mark.parametrize("n,cells", [(30, 300), (61, 1500), (7, 40)])
def test_python_semantics_against_oracle(backend, n, cells):
    pr = Problem.synthetic(n, cells, 1, knock=False)
    ki = np.zeros(n, np.int32); ki[n // 3] = 1
    ko = np.zeros(n, np.int32); ko[min(n - 1, 2)] = 1
    m = oracle_model(pr.net, pr.node_positions, ko, ki)
    m.and_nodes, m.and_inv = pr.pop_an[0], pr.pop_ai[0]
    vals, wres, _ = oracle.py_cluster(m, pr.individuals[0], pr.bm, cells)
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_an[0], pr.pop_ai[0], pr.node_positions, pr.net.ind_len)
    sim = simulator.Simulator(pr.bm, model, lib_path=backend[1])
    with sim:
        res, nst, bits = sim.attractors(pr.individuals[0], ko, ki, "python", 10)
    assert np.array_equal(res, wres)
    for c in range(cells):
        r0, r1 = wres[c]
        rows = list(range(r0, r1 + 1)) if r0 >= 0 else [9]
        assert nst[c] == len(rows)
        assert np.array_equal(bits[c, :len(rows)], vals[c, rows].astype(np.uint8)), c


@pytest.mark.parametrize("n,cells", [(30, 64), (12, 100)])
def test_c_semantics_against_oracle(backend, n, cells):
    """C `cluster` (simulator.c:487-552): window (99 - period, 99) or (0, 0); states [res0, res1)."""
    pr = Problem.synthetic(n, cells, 1, knock=False)
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
    m.and_nodes, m.and_inv = pr.pop_an[0], pr.pop_ai[0]
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_an[0], pr.pop_ai[0], pr.node_positions, pr.net.ind_len)
    with simulator.Simulator(pr.bm, model, lib_path=backend[1]) as sim:
        res, nst, bits = sim.attractors(pr.individuals[0], pr.ko, pr.ki, "c", 12)
    for c in range(cells):
        traj, wres = oracle.c_cluster(m, pr.individuals[0], pr.bm, c)
        assert tuple(res[c]) == tuple(wres), c
        k = int(wres[1] - wres[0])
        assert nst[c] == k
        kk = min(k, 12)
        assert np.array_equal(bits[c, :kk], traj[wres[0]:wres[0] + kk].astype(np.uint8)), c


def test_legacy_cluster_symbol(backend):
    """The reference's `cluster` argument list (simulator.c:487), one cell per call, trajectory left in simData."""
    pr = Problem.synthetic(20, 40, 1, knock=False)
    lib = _lib.load(backend[1])
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
    m.and_nodes, m.and_inv = pr.pop_an[0], pr.pop_ai[0]
    stride, nstride = 64, 32
    lib.scb_set_legacy_geometry(nstride, stride)
    bm = np.zeros((pr.bm.shape[0], stride), np.int32)
    bm[:, :40] = pr.bm
    V = ctypes.c_void_p
    arr = [np.ascontiguousarray(a, np.int32) for a in (pr.individuals[0], pr.net.and_len, pr.net.parse, pr.pop_an[0],
                                                       pr.pop_ai[0], pr.ko, pr.ki, pr.node_positions)]
    ind, al, parse, an, ai, ko, ki, pos = arr
    fn = lib.cluster
    fn.restype = None
    fn.argtypes = None
    for cell in (0, 7, 39):
        vals = np.zeros((100, nstride), np.int32)
        res = np.full(2, 2, np.int32)
        fn(V(vals.ctypes.data), V(res.ctypes.data), V(cell), V(ind.ctypes.data), V(len(ind)), V(pr.n), V(al.ctypes.data),
           V(parse.ctypes.data), V(an.ctypes.data), V(ai.ctypes.data), V(100), V(ko.ctypes.data), V(ki.ctypes.data),
           V(bm.ctypes.data), V(pos.ctypes.data))
        traj, wres = oracle.c_cluster(m, pr.individuals[0], pr.bm, cell)
        assert tuple(res) == tuple(wres)
        assert np.array_equal(vals[:, pr.node_positions], traj)
    lib.scb_set_legacy_geometry(20000, 15000)


def test_attractor_set_then_decider(backend):
    """assignAttractors end to end on the device: unique attractor states, then distance + first-min decider."""
    pr = Problem.synthetic(25, 400, 1, knock=False)
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_an[0], pr.pop_ai[0], pr.node_positions, pr.net.ind_len)
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
    m.and_nodes, m.and_inv = pr.pop_an[0], pr.pop_ai[0]
    vals, wres, _ = oracle.py_cluster(m, pr.individuals[0], pr.bm, 400)
    want = set()
    for c in range(400):
        r0, r1 = wres[c]
        for r in (range(r0, r1 + 1) if r0 >= 0 else [9]):
            want.add(tuple(int(x) for x in vals[c, r]))
    want = {a for a in want if sum(a) > 0}
    with simulator.Simulator(pr.bm, model, lib_path=backend[1]) as sim:
        attrs = sim.attractor_set(pr.individuals[0])
        assert {tuple(int(x) for x in a) for a in attrs} == want
        dist, dec, dd = sim.hamming_argmin(attrs)
        # the same states in first-seen order (cell-major, then window order), deduplicated on the device
        seen, order = set(), []
        for c in range(400):
            r0, r1 = wres[c]
            for r in (range(r0, r1 + 1) if r0 >= 0 else [9]):
                a = tuple(int(x) for x in vals[c, r])
                if a not in seen and sum(a) > 0:
                    seen.add(a); order.append(a)
        assert [tuple(int(x) for x in a) for a in attrs] == order
        # and as bit rows, which the decider takes without unpacking
        packed = sim.attractor_set(pr.individuals[0], packed=True)
        assert packed.dtype == np.uint32 and packed.shape == (len(order), 1)
        _, dec_p, dd_p = sim.hamming_argmin(packed, want_dist=False)
        assert np.array_equal(dec_p, dec) and np.array_equal(dd_p, dd)
    wd, wdec, wdd = oracle.hamming_argmin(attrs, pr.bm, pr.node_positions, 400)
    assert np.array_equal(dist, wd) and np.array_equal(dec, wdec) and np.array_equal(dd, wdd)


def test_packed_attractors_wider_than_one_group(backend):
    """150 nodes = 5 words per state, padded to 8 on the device (the instantiated widths are 4, 8, 12, 16, 24, 32)."""
    pr = Problem.synthetic(150, 300, 1, knock=False)
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_an[0], pr.pop_ai[0], pr.node_positions, pr.net.ind_len)
    with simulator.Simulator(pr.bm, model, lib_path=backend[1]) as sim:
        attrs = sim.attractor_set(pr.individuals[0])
        packed = sim.attractor_set(pr.individuals[0], packed=True)
        assert packed.shape == (attrs.shape[0], 5)
        _, dec, dd = sim.hamming_argmin(attrs, want_dist=False)
        _, dec_p, dd_p = sim.hamming_argmin(packed, want_dist=False)
        assert np.array_equal(dec_p, dec) and np.array_equal(dd_p, dd)
    _, wdec, wdd = oracle.hamming_argmin(attrs, pr.bm, pr.node_positions, 300, want_dist=False)
    assert np.array_equal(dec, wdec) and np.array_equal(dd, wdd)


@pytest.mark.parametrize("n,cells,seed", [(12, 30, None), (20, 200, 4242), (9, 64, 99)])
def test_legacy_importanceScore_symbol(backend, n, cells, seed):
    """importanceScore (simulator.c:555-755): double result bit-identical to the oracle restatement.  The
    restatement is pinned against the unmodified reference build in tests/test_oracle.py
    (test_importance_restatement_matches_reference_build) and the reference's own numbers travel to the GPU box
    as tests/golden/imp_*.npz (test_importance_golden_from_reference_build below)."""
    pr = Problem.synthetic(n, cells, 1, knock=False, net_seed=seed)
    lib = _lib.load(backend[1])
    stride = 256
    lib.scb_set_legacy_geometry(64, stride)
    bm = np.zeros((pr.bm.shape[0], stride), np.int32)
    bm[:, :cells] = pr.bm
    ko = np.zeros(n, np.int32); ko[1 % n] = 1
    ki = np.zeros(n, np.int32); ki[(n - 1)] = 1
    m = oracle_model(pr.net, pr.node_positions, ko, ki)
    m.and_nodes, m.and_inv = pr.pop_an[0], pr.pop_ai[0]
    try:
        want = oracle.importance_score(m, pr.individuals[0], pr.bm, cells)
    except ValueError:
        want = float("nan")
    V = ctypes.c_void_p
    arr = [np.ascontiguousarray(a, np.int32) for a in (pr.individuals[0], pr.net.and_len, pr.net.parse, pr.pop_an[0],
                                                       pr.pop_ai[0], ko, ki, pr.node_positions)]
    ind, al, parse, an, ai, ko_, ki_, pos = arr
    score = np.zeros(1, np.float64)
    fn = lib.importanceScore
    fn.restype = None
    fn.argtypes = None
    fn(V(0), V(ind.ctypes.data), V(len(ind)), V(n), V(al.ctypes.data), V(parse.ctypes.data), V(an.ctypes.data),
       V(ai.ctypes.data), V(100), V(ko_.ctypes.data), V(ki_.ctypes.data), V(cells), V(bm.ctypes.data), V(pos.ctypes.data),
       V(score.ctypes.data))
    lib.scb_set_legacy_geometry(20000, 15000)
    if np.isnan(want):
        assert np.isnan(score[0])
    else:
        assert score[0].tobytes() == np.float64(want).tobytes(), (score[0], want)


def _oracle_importance(pr, an, ai, individual, ko, ki):
    m = oracle_model(pr.net, pr.node_positions, np.asarray(ko, np.int32), np.asarray(ki, np.int32))
    m.and_nodes, m.and_inv = an, ai
    try:
        return oracle.importance_score(m, individual, pr.bm, pr.n_cells)
    except ValueError:
        return float("nan")


def _same_doubles(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.array_equal(np.isnan(got), np.isnan(want)) and got[~np.isnan(got)].tobytes() == want[~np.isnan(want)].tobytes()


@pytest.mark.parametrize("n,cells,seed", [(20, 200, 4242), (61, 1500, None), (33, 4200, 5)])
def test_importance_scores_all_nodes_one_launch(backend, n, cells, seed):
    """The loop over nodes of ruleMaker.__calcImportance (ruleMaker.py:1285-1300) as one batched call: every
    score bit-identical to the oracle's importanceScore for KOlist = KIlist = [node]."""
    pr = Problem.synthetic(n, cells, 1, knock=False, net_seed=seed)
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_an[0], pr.pop_ai[0], pr.node_positions, pr.net.ind_len)
    want = []
    for node in range(n):
        k = np.zeros(n, np.int32); k[node] = 1
        want.append(_oracle_importance(pr, pr.pop_an[0], pr.pop_ai[0], pr.individuals[0], k, k))
    with simulator.Simulator(pr.bm, model, lib_path=backend[1]) as sim:
        if backend[0] == "emu":
            sim.set_option(_lib.OPT_MAX_CTAS, 3)
        got = sim.importance_scores(pr.individuals[0], [[i] for i in range(n)], strict=False)
        one = sim.evaluate_by_node(pr.individuals[0], [3], [3], importanceScores=True) if not np.isnan(want[3]) else None
    assert _same_doubles(got, want), (got, want)
    assert not np.isnan(np.asarray(want)).all()
    if one is not None:
        assert np.float64(one[0]).tobytes() == np.float64(want[3]).tobytes()


def test_importance_scores_distinct_knockin_arrays(backend):
    """Knock-in arrays that differ from the knock-out arrays: two simulated passes per item (simulator.c:674, :717),
    non-0/1 knock values count as set (:604 tests for non-zero)."""
    n, cells = 26, 900
    pr = Problem.synthetic(n, cells, 1, knock=False, net_seed=77)
    rng = np.random.default_rng(3)
    items = 9
    kos = (rng.random((items, n)) < 0.1).astype(np.int32) * rng.integers(1, 4, (items, n)).astype(np.int32)
    kis = (rng.random((items, n)) < 0.1).astype(np.int32)
    kis[4] = (kos[4] != 0)                                       # one item whose passes coincide
    want = [_oracle_importance(pr, pr.pop_an[0], pr.pop_ai[0], pr.individuals[0], kos[k], kis[k]) for k in range(items)]
    lib = _lib.load(backend[1])
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_an[0], pr.pop_ai[0], pr.node_positions, pr.net.ind_len)
    with simulator.Simulator(pr.bm, model, lib_path=backend[1]) as sim:
        if backend[0] == "emu":
            sim.set_option(_lib.OPT_MAX_CTAS, 3)
        ind = np.ascontiguousarray(pr.individuals[0], np.int32)
        got = np.zeros(items, np.float64)
        P = lambda a: ctypes.c_void_p(a.ctypes.data)
        m = sim.model
        rc = lib.scb_importance_scores(sim._ctx, P(ind), m.ind_len, P(m.and_len_list), P(m.individual_parse), P(m.and_nodes),
                                       P(m.and_node_invert_list), items, P(kos), P(kis), P(got))
    assert rc in (_lib.OK, _lib.ERR_INVALID)
    assert (rc == _lib.ERR_INVALID) == bool(np.isnan(np.asarray(want)).any())
    assert _same_doubles(got, want), (got, want)


@pytest.mark.parametrize("rings,cells", [([40, 5], 1300), ([60, 7], 700)])
def test_importance_scores_long_cycles(backend, rings, cells):
    if backend[0] == "emu":
        cells = 200 if cells > 1000 else 90            # one chunk under the host-thread emulator: same code paths, a tenth of the time
    """Rings of period 80 (found only by the exact resolver: window length 80 adds (int)(state / 80) = 0) and of
    period 120 (no attractor window in 100 rows: the reference divides by zero -> NaN for the items that leave
    the ring alive, exact scores for the items that knock it)."""
    pr = ring_problem(rings, 3, cells)
    n = pr.n
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv, pr.node_positions, pr.net.ind_len)
    want = []
    for node in range(n):
        k = np.zeros(n, np.int32); k[node] = 1
        want.append(_oracle_importance(pr, pr.net.and_nodes, pr.net.and_inv, pr.individuals[0], k, k))
    with simulator.Simulator(pr.bm, model, lib_path=backend[1]) as sim:
        if backend[0] == "emu":
            sim.set_option(_lib.OPT_MAX_CTAS, 3)
        got = sim.importance_scores(pr.individuals[0], [[i] for i in range(n)], strict=False)
        if np.isnan(np.asarray(want)).any():
            with pytest.raises(_lib.ScbError):
                sim.importance_scores(pr.individuals[0], [[i] for i in range(n)])
    assert _same_doubles(got, want), (got, want)
    assert (~np.isnan(np.asarray(want))).any()


# ---- golden vectors produced by the UNMODIFIED reference build (tests/golden/make_golden.py) -------------------
def _golden_problem(g):
    C = int(g["n_cells"]); pos = g["node_positions"]
    bm = np.zeros((int(pos.max()) + 1, C), np.int32)
    bm[pos] = np.unpackbits(g["bin_rows"], axis=1)[:, :C]
    model = simulator.ModelTables(g["and_len"], g["parse"], g["and_nodes"], g["and_inv"], pos, int(g["ind_len"]))
    return C, bm, model


@pytest.mark.parametrize("name", ["imp_n33.npz", "imp_ring.npz"])
def test_importance_golden_from_reference_build(backend, name):
    """importanceScore of the reference itself (oracle/_ref, simulator.c:555-755): every double bit-identical,
    through the batched entry point and, for a few items, through the legacy 15-argument symbol."""
    g = np.load(os.path.join(GOLDEN, name))
    C, bm, model = _golden_problem(g)
    kos, kis, want = g["knockouts"], g["knockins"], g["ref_scores"]
    lib = _lib.load(backend[1])
    with simulator.Simulator(bm, model, lib_path=backend[1]) as sim:
        if backend[0] == "emu":
            sim.set_option(_lib.OPT_MAX_CTAS, 3)
        ind = np.ascontiguousarray(g["individual"], np.int32)
        got = np.zeros(len(want), np.float64)
        P = lambda a: ctypes.c_void_p(a.ctypes.data)
        m = sim.model
        kos_c, kis_c = np.ascontiguousarray(kos, np.int32), np.ascontiguousarray(kis, np.int32)
        rc = lib.scb_importance_scores(sim._ctx, P(ind), m.ind_len, P(m.and_len_list), P(m.individual_parse), P(m.and_nodes),
                                       P(m.and_node_invert_list), len(want), P(kos_c), P(kis_c), P(got))
    assert rc == _lib.OK
    assert got.tobytes() == want.tobytes(), (got, want)
    # the drop-in symbol, reference argument list
    stride = 512
    lib.scb_set_legacy_geometry(64, stride)
    wide = np.zeros((bm.shape[0], stride), np.int32)
    wide[:, :C] = bm
    V = ctypes.c_void_p
    fn = lib.importanceScore
    fn.restype = None
    fn.argtypes = None
    al, parse, an, ai, pos = [np.ascontiguousarray(a, np.int32) for a in (g["and_len"], g["parse"], g["and_nodes"], g["and_inv"], g["node_positions"])]
    for k in (0, len(want) // 2, len(want) - 1):
        score = np.zeros(1, np.float64)
        ko, ki = np.ascontiguousarray(kos[k]), np.ascontiguousarray(kis[k])
        fn(V(0), V(ind.ctypes.data), V(len(ind)), V(model.n), V(al.ctypes.data), V(parse.ctypes.data), V(an.ctypes.data),
           V(ai.ctypes.data), V(100), V(ko.ctypes.data), V(ki.ctypes.data), V(C), V(wide.ctypes.data), V(pos.ctypes.data),
           V(score.ctypes.data))
        assert score.tobytes() == want[k:k + 1].tobytes(), (k, score, want[k])
    lib.scb_set_legacy_geometry(20000, 15000)


@pytest.mark.parametrize("name", ["cluster_n33.npz", "cluster_ring.npz"])
def test_cluster_golden_from_reference_build(backend, name):
    """The reference's own `cluster` (oracle/_ref, simulator.c:487-552): attractor window and the scratch the call
    leaves behind, including knock-outs on permuted nodePositions, where the reference records the knocked node in
    column i instead of column nodePositions[i] (:518)."""
    g = np.load(os.path.join(GOLDEN, name))
    C, bm, model = _golden_problem(g)
    stride, cstride = int(g["stride"]), 512
    lib = _lib.load(backend[1])
    lib.scb_set_legacy_geometry(stride, cstride)
    wide = np.zeros((bm.shape[0], cstride), np.int32)
    wide[:, :C] = bm
    V = ctypes.c_void_p
    fn = lib.cluster
    fn.restype = None
    fn.argtypes = None
    ind, al, parse, an, ai, pos = [np.ascontiguousarray(a, np.int32) for a in (g["individual"], g["and_len"], g["parse"], g["and_nodes"],
                                                                              g["and_inv"], g["node_positions"])]
    try:
        for k in range(len(g["knockouts"])):
            ko, ki = np.ascontiguousarray(g["knockouts"][k]), np.ascontiguousarray(g["knockins"][k])
            for ci, cell in enumerate(g["cells"]):
                vals = np.zeros((100, stride), np.int32)
                res = np.full(2, 2, np.int32)
                fn(V(vals.ctypes.data), V(res.ctypes.data), V(int(cell)), V(ind.ctypes.data), V(len(ind)), V(model.n), V(al.ctypes.data),
                   V(parse.ctypes.data), V(an.ctypes.data), V(ai.ctypes.data), V(100), V(ko.ctypes.data), V(ki.ctypes.data),
                   V(wide.ctypes.data), V(pos.ctypes.data))
                assert tuple(res) == tuple(g["ref_res"][k, ci]), (k, cell, res, g["ref_res"][k, ci])
                assert np.array_equal(vals.astype(np.int8), g["ref_sim"][k, ci]), (k, cell)
    finally:
        lib.scb_set_legacy_geometry(20000, 15000)


def test_cluster_symbol_with_dirty_scratch(backend):
    """A knocked node on permuted positions leaves its own column unwritten (simulator.c:518): what the caller's
    scratch held there shows through, in the window search too.  Against the literal restatement, which
    tests/test_oracle.py pins against the reference build for exactly this case."""
    pr = Problem.synthetic(20, 40, 1, knock=False)
    n = pr.n
    stride = max(int(pr.node_positions.max()) + 1, n)
    ko = np.zeros(n, np.int32); ko[[2, 11]] = 1
    ki = np.zeros(n, np.int32); ki[5] = 1
    m = oracle_model(pr.net, pr.node_positions, ko, ki)
    m.and_nodes, m.and_inv = pr.pop_an[0], pr.pop_ai[0]
    rng = np.random.default_rng(17)
    dirty = rng.integers(0, 2, (100, stride)).astype(np.int32)
    lib = _lib.load(backend[1])
    lib.scb_set_legacy_geometry(stride, 64)
    wide = np.zeros((pr.bm.shape[0], 64), np.int32)
    wide[:, :40] = pr.bm
    V = ctypes.c_void_p
    fn = lib.cluster
    fn.restype = None
    fn.argtypes = None
    ind, al, parse, an, ai, pos = [np.ascontiguousarray(a, np.int32) for a in (pr.individuals[0], pr.net.and_len, pr.net.parse, pr.pop_an[0],
                                                                              pr.pop_ai[0], pr.node_positions)]
    try:
        for cell in (0, 13, 39):
            vals = dirty.copy()
            res = np.full(2, 2, np.int32)
            fn(V(vals.ctypes.data), V(res.ctypes.data), V(cell), V(ind.ctypes.data), V(len(ind)), V(n), V(al.ctypes.data),
               V(parse.ctypes.data), V(an.ctypes.data), V(ai.ctypes.data), V(100), V(ko.ctypes.data), V(ki.ctypes.data),
               V(wide.ctypes.data), V(pos.ctypes.data))
            sd, wres = oracle.c_cluster_literal(m, pr.individuals[0], pr.bm, cell, sim_data=dirty, sim_stride=stride)
            assert tuple(res) == tuple(wres), (cell, res, wres)
            assert np.array_equal(vals, sd), cell
    finally:
        lib.scb_set_legacy_geometry(20000, 15000)

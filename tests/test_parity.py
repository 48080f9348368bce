"""Parity of the CUDA path with the oracle, bit for bit.

Every test runs against two builds of the SAME sources (scbonita_b200/csrc):
  * ``emu``  -- g++ + the host-thread emulator (tests/emu), on CPU, small sizes;
  * ``cuda`` -- the product library on a B200 (marked ``gpu``), through the C-ABI, larger sizes.
The oracle (oracle/) is only the checker.
"""
import ctypes
import os

import numpy as np
import pytest

import oracle
from common import EMU_LIB, GOLDEN, Problem, oracle_model, ring_problem, tables
from scbonita_b200 import _lib, simulator

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def backend(request):
    if request.param == "emu":
        return "emu", request.getfixturevalue("emu_lib")
    return "cuda", request.getfixturevalue("cuda_lib")


def big(backend, small, large):
    return large if backend[0] == "cuda" else small


# ---- seeded synthetic shapes ------------------------------------------------------------------
@pytest.mark.parametrize("n,cells,pop", [(20, 200, 3), (60, 1100, 4), (200, 1300, 3), (230, 500, 2), (460, 200, 2),
                                        (5, 40, 4), (1, 70, 2)])
def test_synthetic_matches_oracle(backend, n, cells, pop):
    cells = big(backend, cells, cells * 9 + 17)
    pop = big(backend, pop, pop * 4)
    pr = Problem.synthetic(n, cells, pop)
    with pr.simulator(backend[1]) as sim:
        pr.check(sim)


@pytest.mark.parametrize("opts", [dict(early_exit=0, snapshots=0), dict(early_exit=1, snapshots=0),
                                  dict(early_exit=0, snapshots=1), dict(words=1), dict(words=2, check_every=3),
                                  dict(warps=4), dict(period_search=24), dict(check_every=1), dict(snap_min_gap=1, check_every=4),
                                  dict(snap_min_gap=20, snap_lo=1 << 20, snap_hi=0), dict(fuse=0), dict(fuse=0, early_exit=0, snapshots=0),
                                  dict(steady=0), dict(steady=1, snap_lo=(1 << 4) | (1 << 30), snap_hi=0, snap_min_gap=2),
                                  dict(steady=1, tmem=0), dict(steady=0, tmem=0), dict(steady=1, early_exit=0, snapshots=0),
                                  dict(check_from=-1), dict(check_from=0, check_every=3), dict(check_from=4)])
def test_every_schedule_gives_identical_integers(backend, opts):
    """Early exit, snapshot resolution and tile geometry are optimisations: results must not move."""
    pr = Problem.synthetic(60, big(backend, 2100, 21000), big(backend, 5, 24))
    with pr.simulator(backend[1], **opts) as sim:
        pr.check(sim)


def test_no_knockouts(backend):
    pr = Problem.synthetic(47, big(backend, 100, 5000), 4, knock=False)
    with pr.simulator(backend[1]) as sim:
        pr.check(sim)


def test_shared_tables(backend):
    """tablesPerIndividual = 0: one [N][7][3] model for the whole population."""
    pr = Problem.synthetic(60, big(backend, 600, 6000), 5)
    pr.pop_an = pr.pop_ai = None
    with pr.simulator(backend[1]) as sim:
        pr.check(sim)


def test_uint8_matrix(backend):
    pr = Problem.synthetic(30, 777, 3)
    sim = simulator.Simulator(pr.bm.astype(np.uint8), tables(pr.net, pr.node_positions), lib_path=backend[1])
    if backend[0] == "emu":
        sim.set_option(_lib.OPT_MAX_CTAS, 2)
    with sim:
        pr.check(sim)


# ---- golden vectors from the unmodified reference ---------------------------------------------
@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLDEN) if f.startswith("sim_")))
def test_golden_vectors(backend, name):
    g = np.load(os.path.join(GOLDEN, name))
    C = int(g["n_cells"])
    pos = g["node_positions"]
    bm = np.zeros((int(pos.max()) + 1, C), np.int32)
    bm[pos] = np.unpackbits(g["bin_rows"], axis=1)[:, :C]
    model = simulator.ModelTables(g["and_len"], g["parse"], g["and_nodes"], g["and_inv"], pos, int(g["ind_len"]))
    per_ind = g["pop_and_nodes"].ndim == 4
    an = g["pop_and_nodes"] if per_ind else None
    ai = g["pop_and_inv"] if per_ind else None
    sim = simulator.Simulator(bm, model, lib_path=backend[1])
    if backend[0] == "emu":
        sim.set_option(_lib.OPT_MAX_CTAS, 3)
    with sim:
        cw = sim.evaluate_population(g["individuals"], g["knockouts"], g["knockins"], an, ai, mode=simulator.MODE_CELLWISE)
        nw = sim.evaluate_population(g["individuals"], g["knockouts"], g["knockins"], an, ai, mode=simulator.MODE_NODEWISE)
        fit = sim.evaluate_population(g["individuals"], g["knockouts"], g["knockins"], an, ai, mode=simulator.MODE_SUM)
    assert np.array_equal(cw, g["ref_cellwise"])
    assert np.array_equal(nw, g["ref_nodewise"])
    assert np.array_equal(fit, g["ref_cellwise"].sum(axis=1))


# ---- attractor-found semantics ---------------------------------------------------
@pytest.mark.parametrize("rings,chain,cells", [([60], 0, 300), ([45], 12, 2500), ([3, 5], 4, 1500), ([49], 0, 200),
                                               ([50], 0, 200), ([7, 11, 13], 20, 3000), ([2, 3, 4], 0, 1100)])
@pytest.mark.parametrize("opts", [dict(), dict(early_exit=0, snapshots=0), dict(words=1), dict(fuse=0)])
def test_cycles_not_found_and_fill_forward(backend, rings, chain, cells, opts):
    pr = ring_problem(rings, chain, cells)
    with pr.simulator(backend[1], **opts) as sim:
        st = pr.check(sim)
    assert st["undefined_leading"] == 0


def test_leading_not_found_cells_are_counted_and_defined_as_zero(backend):
    """The reference reads uninitialised stack for not-found cells before the first found one; the oracle
    and the library both define them as 0 and count them."""
    pr = ring_problem([60], 0, 2300, gate_density=0.97, lead_found=False, seed=3)
    o = pr.oracle()
    with pr.simulator(backend[1]) as sim:
        st = pr.check(sim)
    first_found = int(np.argmax(pr.bin_rows[0] == 0)) if (pr.bin_rows[0] == 0).any() else pr.n_cells
    assert st["undefined_leading"] >= 1
    assert st["undefined_leading"] <= first_found
    assert int(o["not_found"][0]) == st["not_found_cells"]


def test_fill_forward_across_chunk_boundaries(backend):
    """Runs of not-found cells longer than a 1024-cell chunk: the value comes from an earlier chunk."""
    pr = ring_problem([55], 0, 4000, gate_density=0.5, seed=11)
    pr.bin_rows[0, 900:3300] = 1          # gate on: 2400 consecutive not-found cells
    pr.bm[pr.node_positions[0]] = pr.bin_rows[0]
    with pr.simulator(backend[1]) as sim:
        pr.check(sim)


# ---- reference quirks, hand-built tables ------------------------------------------------------------
def _random_tables(rng, n, max_distinct=3, weird_flags=True):
    """Arbitrary reference-format tables (not only what ruleMaker generates): random term counts, repeated
    inputs, flags outside {0,1} (constant-TRUE literals, Q3), zero-input nodes (Q2)."""
    class Net:
        pass
    net = Net()
    net.n = n
    net.and_len = np.zeros(n, np.int32)
    net.parse = np.zeros(n + 1, np.int32)
    net.and_nodes = np.zeros((n, 7, 3), np.int32)
    net.and_inv = np.zeros((n, 7, 3), np.int32)
    cnt = 0
    for i in range(n):
        net.parse[i] = cnt
        kind = rng.integers(0, 10)
        if kind == 0:
            net.and_len[i] = 0
            net.and_nodes[i, 0] = [rng.integers(0, n), 0, 0]
        elif kind <= 2:
            net.and_len[i] = 1
            net.and_nodes[i, 0] = [rng.integers(0, n), -1, -1]
            net.and_inv[i, 0] = [rng.choice([0, 1, 1, 0, 5, -1]) if weird_flags else rng.integers(0, 2), -1, -1]
            cnt += 1
        else:
            nt = int(rng.integers(2, 8))
            pool = rng.choice(n, size=min(n, max_distinct), replace=False)
            net.and_len[i] = nt
            for a in range(nt):
                k = int(rng.integers(1, 4))
                ins = list(rng.choice(pool, size=k)) + [-1] * (3 - k)
                fl = [int(rng.choice([0, 1, 0, 1, 7, -3])) if weird_flags else int(rng.integers(0, 2)) for _ in range(k)] + [-1] * (3 - k)
                net.and_nodes[i, a] = ins
                net.and_inv[i, a] = fl
            cnt += nt
    net.parse[n] = cnt
    net.ind_len = cnt
    return net


@pytest.mark.parametrize("seed", range(6))
def test_arbitrary_tables(backend, seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(3, 40))
    net = _random_tables(rng, n)
    cells = big(backend, 300, 3000)
    pop = 4
    rows = (rng.random((n, cells)) < 0.4).astype(np.uint8)
    inds = rng.integers(0, 2, size=(pop, max(net.ind_len, 1))).astype(np.int32)[:, :net.ind_len]
    for i in range(n):          # at least one selected term per multi-term node (Q9 contract)
        s, e = net.parse[i], net.parse[i + 1]
        if net.and_len[i] >= 2:
            inds[:, s] |= (inds[:, s:e].sum(axis=1) == 0)
    ko = (rng.random(n) < 0.1).astype(np.int32) * rng.choice([1, 2], size=n)      # 2 does not knock (Q5)
    ki = (rng.random(n) < 0.1).astype(np.int32) * rng.choice([1, 3], size=n)
    pr = Problem(net, rows, rng.permutation(n + 5)[:n].astype(np.int32), inds, ko=ko.astype(np.int32), ki=ki.astype(np.int32))
    with pr.simulator(backend[1]) as sim:
        pr.check(sim)


def test_empty_rule_is_defined_as_zero_and_counted(backend):
    """Q9: a multi-term node with no selected term reads uninitialised memory in the reference."""
    rng = np.random.default_rng(5)
    net = _random_tables(rng, 12, weird_flags=False)
    rows = (rng.random((12, 100)) < 0.5).astype(np.uint8)
    inds = np.zeros((2, net.ind_len), np.int32)
    inds[1] = 1
    pr = Problem(net, rows, np.arange(12, dtype=np.int32), inds)
    with pr.simulator(backend[1]) as sim:
        st = pr.check(sim)
    assert st["empty_rule_nodes"] == int((net.and_len >= 2).sum())


# ---- error behaviour ------------------------------------------------------------------------------------
def test_more_than_three_distinct_inputs_go_through_the_general_simulator(backend):
    """Any [7][3] table is legal input of scSyncBool.  A rule over more than three distinct state bits cannot be one
    LOP3: the population is scored again, exactly, by the general simulator (scb_k_generic)."""
    rng = np.random.default_rng(9)
    net = _random_tables(rng, 12, max_distinct=5, weird_flags=False)
    rows = (rng.random((12, 1500)) < 0.5).astype(np.uint8)
    inds = (rng.random((3, net.ind_len)) < 0.6).astype(np.int32)
    inds[0] = 1
    pr = Problem(net, rows, np.arange(12, dtype=np.int32), inds)
    with pr.simulator(backend[1]) as sim:
        pr.check(sim)
        assert sim.last_stats["unsupported_nodes"] == 0 and sim.last_stats["tiles"] == 3 * 2     # 3 individuals x 2 chunks, generic


@pytest.mark.parametrize("n,cells,pop", [(20, 200, 3), (60, 2100, 4), (230, 1100, 2)])
def test_general_simulator_matches_oracle(backend, n, cells, pop):
    """SCB_OPT_GENERIC: ordinary populations through the general simulator, all three output shapes."""
    pr = Problem.synthetic(n, cells, pop)
    with pr.simulator(backend[1], generic=1) as sim:
        pr.check(sim)
        assert sim.last_stats["tiles"] == pop * -(-cells // 1024)


@pytest.mark.parametrize("rings,chain,cells", [([60], 0, 300), ([45], 12, 2500), ([3, 5], 4, 1500)])
def test_general_simulator_cycles_not_found_and_fill_forward(backend, rings, chain, cells):
    pr = ring_problem(rings, chain, cells)
    with pr.simulator(backend[1], generic=1) as sim:
        pr.check(sim)


def test_general_simulator_compact_population(backend):
    pr = Problem.synthetic(40, 1300, 3)
    up, upinv = zip(*(simulator.upstream_from_tables(pr.net.and_len, pr.pop_an[i], pr.pop_ai[i]) for i in range(3)))
    with pr.simulator(backend[1], generic=1) as sim:
        sim.upload_compact(pr.individuals, np.array(up, np.int32), np.array(upinv, np.int32), pr.ko, pr.ki)
        sim.run()
        assert np.array_equal(sim.fetch(), pr.oracle()["fitness"])


def test_bad_input_index_is_refused(backend):
    pr = Problem.synthetic(10, 64, 1)
    pr.net.and_nodes = pr.net.and_nodes.copy()
    i = int(np.argmax(pr.net.and_len >= 1))
    pr.net.and_nodes[i, 0, 0] = 10
    pr.pop_an = pr.pop_ai = None
    with pr.simulator(backend[1]) as sim:
        with pytest.raises(simulator.ScbError) as ei:
            sim.evaluate_population(pr.individuals, pr.ko * 0, pr.ki * 0)
        assert ei.value.code == _lib.ERR_INVALID


def test_non_binary_matrix_is_refused(backend):
    pr = Problem.synthetic(8, 64, 1)
    bm = pr.bm.copy()
    bm[pr.node_positions[3], 5] = 2
    with pytest.raises(simulator.ScbError) as ei:
        simulator.Simulator(bm, tables(pr.net, pr.node_positions), lib_path=backend[1])
    assert ei.value.code == _lib.ERR_INVALID


# ---- callers: local search, per-individual entry point, legacy symbol -------------------------------------
def test_evaluate_by_node_mirrors_reference_returns(backend):
    pr = Problem.synthetic(30, 400, 2, knock=True)
    pr.pop_an = pr.pop_ai = None
    o = pr.oracle()
    with pr.simulator(backend[1]) as sim:
        # KOlist = [0]*N used as indices -> node 0 knocked out and in (Q4)
        assert sim.evaluate_by_node(pr.individuals[0].tolist(), [0] * pr.n, [0] * pr.n) == [int(o["fitness"][0])]
        assert sim.evaluate_by_node(pr.individuals[1].tolist(), [0] * pr.n, [0] * pr.n, localSearch=True) == o["nodewise"][1].tolist()


def test_check_node_possibilities(backend):
    pr = Problem.synthetic(big(backend, 8, 25), big(backend, 100, 4000), 1)
    pr.pop_an = pr.pop_ai = None
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
    with pr.simulator(backend[1]) as sim:
        for node in range(pr.n):
            start, end = int(pr.net.parse[node]), int(pr.net.parse[node + 1])
            truth, equivs, minny, errs = sim.check_node_possibilities(node, pr.individuals[0].tolist(), [0] * pr.n, [0] * pr.n)
            if end == start:
                assert errs == 0.0
                continue
            want = []
            for v in range(1, 2 ** (end - start)):
                ind = pr.individuals[0].copy()
                ind[start:end] = [(v >> j) & 1 for j in range(end - start)]          # ruleMaker.__bitList
                e, _ = oracle.sync_bool(m, ind, pr.bm, pr.n_cells, 1)
                want.append(int(e[node]))
            assert errs == want
            assert minny == min(want)
            assert equivs == [[(v >> j) & 1 for j in range(end - start)] for v in range(1, 2 ** (end - start)) if want[v - 1] == min(want)]
            assert truth == equivs[0]


def test_local_search_all_nodes_in_one_launch(backend):
    pr = Problem.synthetic(big(backend, 8, 25), big(backend, 100, 4000), 1)
    pr.pop_an = pr.pop_ai = None
    with pr.simulator(backend[1]) as sim:
        every = sim.local_search(pr.individuals[0].tolist(), [0] * pr.n, [0] * pr.n)
        for node in range(pr.n):
            assert every[node] == sim.check_node_possibilities(node, pr.individuals[0].tolist(), [0] * pr.n, [0] * pr.n)


def test_legacy_scSyncBool_symbol(backend):
    """The reference's own calling convention: 17 x c_void_p, ints smuggled as pointers (ruleMaker.py:1117-1219)."""
    pr = Problem.synthetic(20, 150, 1)
    lib = _lib.load(backend[1])
    o = pr.oracle()
    stride = 256
    lib.scb_set_legacy_geometry(20000, stride)
    bm = np.zeros((pr.bm.shape[0], stride), np.int32)
    bm[:, :150] = pr.bm
    V = ctypes.c_void_p
    vals = np.zeros((100, 64), np.int32)
    arr = [np.ascontiguousarray(a, np.int32) for a in (pr.individuals[0], pr.net.and_len, pr.net.parse, pr.pop_an[0],
                                                       pr.pop_ai[0], pr.ko, pr.ki, pr.node_positions)]
    ind, al, parse, an, ai, ko, ki, pos = arr
    fn = lib.scSyncBool
    fn.restype = None
    fn.argtypes = None
    for local, want in ((0, o["cellwise"][0]), (1, o["nodewise"][0])):
        errors = np.full(max(pr.n, 150) + 3, 7 if local else -1, np.int32)
        fn(V(vals.ctypes.data), V(ind.ctypes.data), V(len(ind)), V(pr.n), V(al.ctypes.data), V(parse.ctypes.data),
           V(an.ctypes.data), V(ai.ctypes.data), V(100), V(ko.ctypes.data), V(ki.ctypes.data), V(150),
           V(bm.ctypes.data), V(pos.ctypes.data), V(errors.ctypes.data), V(local), V(0))
        if local:
            assert np.array_equal(errors[:pr.n], want + 7)           # accumulates (simulator.c:477-479)
            assert (errors[pr.n:] == 7).all()
        else:
            assert np.array_equal(errors[:150], want)                # assigns (simulator.c:481)
            assert (errors[150:] == -1).all()
    lib.scb_set_legacy_geometry(20000, 15000)


# ---- attractor companion ----------------------------------------------------------------------------------
def test_hamming_matches_oracle_and_reference_kat(backend):
    g = np.load(os.path.join(GOLDEN, "hamming_kat.npz"))
    n = g["attractors"].shape[1]
    model = simulator.ModelTables(np.zeros(n), np.zeros(n + 1), np.zeros((n, 7, 3)), np.zeros((n, 7, 3)), g["node_positions"], 0)
    with simulator.Simulator(g["bin_mat"], model, n_cells=int(g["n_cells"]), lib_path=backend[1]) as sim:
        dist, dec, dd = sim.hamming_argmin(g["attractors"])
    assert np.array_equal(dist, g["dist"])
    assert np.array_equal(dec, g["decider"])
    assert np.array_equal(dd, g["decider_distance"])


@pytest.mark.parametrize("n,cells,A", [(47, 333, 7), (300, 1000, 40), (33, 64, 1)])
def test_hamming_random(backend, n, cells, A):
    rng = np.random.default_rng(n + A)
    cells = big(backend, cells, cells * 20)
    rows = (rng.random((n, cells)) < 0.3).astype(np.int32)
    pos = rng.permutation(n).astype(np.int32)
    bm = np.zeros((n, cells), np.int32)
    bm[pos] = rows
    attr = (rng.random((A, n)) < 0.4).astype(np.int32)
    if A > 3:
        attr[3] = attr[1]             # ties: first minimum wins (pandas idxmin)
    model = simulator.ModelTables(np.zeros(n), np.zeros(n + 1), np.zeros((n, 7, 3)), np.zeros((n, 7, 3)), pos, 0)
    with simulator.Simulator(bm, model, lib_path=backend[1]) as sim:
        dist, dec, dd = sim.hamming_argmin(attr)
    wd, wdec, wdd = oracle.hamming_argmin(attr, bm, pos, cells)
    assert np.array_equal(dist, wd) and np.array_equal(dec, wdec) and np.array_equal(dd, wdd)


def test_csr_matrix_ingest(backend):
    """the context is built straight from scipy CSR (how the reference stores binMat)."""
    import scipy.sparse as sp
    pr = Problem.synthetic(40, big(backend, 1500, 15000), 3)
    big_rows = np.zeros((pr.bm.shape[0] + 7, pr.n_cells + 100), np.float64)      # extra genes / extra columns
    big_rows[:pr.bm.shape[0], :pr.n_cells] = pr.bm
    big_rows[:, pr.n_cells:] = 1.0                                               # beyond lenSamples: ignored
    csr = sp.csr_matrix(big_rows)
    sim = simulator.Simulator(csr, tables(pr.net, pr.node_positions), n_cells=pr.n_cells, lib_path=backend[1])
    if backend[0] == "emu":
        sim.set_option(_lib.OPT_MAX_CTAS, 2)
    with sim:
        pr.check(sim)
    bad = sp.csr_matrix(np.where(big_rows > 0, 2.0, 0.0))
    with pytest.raises(simulator.ScbError) as ei:
        simulator.Simulator(bad, tables(pr.net, pr.node_positions), n_cells=pr.n_cells, lib_path=backend[1])
    assert ei.value.code == _lib.ERR_INVALID


def test_legacy_symbol_is_reentrant(backend):
    """The reference's parallel local search calls scSyncBool from a ThreadPool (singleCell.py:487-497; ctypes
    releases the GIL): concurrent calls must not interfere."""
    from multiprocessing.pool import ThreadPool
    pr = Problem.synthetic(20, 150, 6)
    lib = _lib.load(backend[1])
    o = pr.oracle()
    stride = 256
    lib.scb_set_legacy_geometry(20000, stride)
    bm = np.zeros((pr.bm.shape[0], stride), np.int32)
    bm[:, :150] = pr.bm
    V = ctypes.c_void_p
    al, parse, ko, ki, pos = [np.ascontiguousarray(a, np.int32) for a in (pr.net.and_len, pr.net.parse, pr.ko, pr.ki, pr.node_positions)]
    fn = lib.scSyncBool
    fn.restype = None
    fn.argtypes = None

    def one(p):
        ind, an, ai = [np.ascontiguousarray(a, np.int32) for a in (pr.individuals[p], pr.pop_an[p], pr.pop_ai[p])]
        errors = np.zeros(pr.n + 150, np.int32)
        vals = np.zeros((100, 64), np.int32)
        fn(V(vals.ctypes.data), V(ind.ctypes.data), V(len(ind)), V(pr.n), V(al.ctypes.data), V(parse.ctypes.data),
           V(an.ctypes.data), V(ai.ctypes.data), V(100), V(ko.ctypes.data), V(ki.ctypes.data), V(150),
           V(bm.ctypes.data), V(pos.ctypes.data), V(errors.ctypes.data), V(1), V(0))
        return errors[:pr.n].copy()

    with ThreadPool(4) as pool:
        got = pool.map(one, list(range(6)) * 2)
    lib.scb_set_legacy_geometry(20000, 15000)
    for p, e in zip(list(range(6)) * 2, got):
        assert np.array_equal(e, o["nodewise"][p]), p


def test_network_too_large_for_on_chip_state_uses_the_general_simulator(backend):
    """N beyond the shared-memory state buffers (the reference allows NODE = 20000): populations are scored by the
    general simulator, state in global memory; the other entry points refuse."""
    rng = np.random.default_rng(2)
    n = 800 if backend[0] == "emu" else 1500
    net = _random_tables(rng, n, weird_flags=False)
    rows = (rng.random((n, 70 if backend[0] == "emu" else 200)) < 0.5).astype(np.uint8)
    inds = (rng.random((2, net.ind_len)) < 0.7).astype(np.int32)
    pr = Problem(net, rows, np.arange(n, dtype=np.int32), inds)
    with pr.simulator(backend[1]) as sim:
        pr.check(sim, modes=(0, 2) if backend[0] == "emu" else (0, 1, 2))
        with pytest.raises(simulator.ScbError) as ei:
            sim.importance_scores(pr.individuals[0], [[0]])
        assert ei.value.code == _lib.ERR_UNSUPPORTED


# ---- compact population ------------------------------------------------------------
def test_upstream_triples_rebuild_the_reference_tables():
    """upstream_from_tables + the device-side expansion rule (scb_expand_terms, restated here in numpy) give back
    exactly the [7][3] tables the reference marshals (ruleMaker.py:188-200, :337-360)."""
    import itertools
    pr = Problem.synthetic(80, 10, 6)
    up, ui = simulator.upstream_from_tables(pr.net.and_len, pr.pop_an, pr.pop_ai)
    assert up.shape == (6, 80, 3)
    for p in range(6):
        for i in range(80):
            ins = [int(x) for x in up[p, i] if x >= 0]
            fl = [int(x) for x in ui[p, i][:len(ins)]]
            terms = [[(a, f) for (a, f) in t if a != "empty"] for t in
                     itertools.product(*[((a, f), ("empty", 0)) for a, f in zip(ins, fl)])]
            terms = [t for t in terms if t]
            an = np.zeros((7, 3), np.int32); ai = np.zeros((7, 3), np.int32)
            for m, t in enumerate(terms):
                an[m] = [a for a, _ in t] + [-1] * (3 - len(t))
                ai[m] = [f for _, f in t] + [-1] * (3 - len(t))
            assert np.array_equal(an, pr.pop_an[p, i]) and np.array_equal(ai, pr.pop_ai[p, i]), (p, i)


@pytest.mark.parametrize("n,cells,pop,shared", [(60, 700, 6, False), (200, 1300, 3, False), (33, 400, 5, True)])
def test_compact_population_matches_oracle(backend, n, cells, pop, shared):
    """The population handed over as bit-strings + upstream triples (24 instead of 168 bytes per node): identical
    integers in every output shape."""
    cells = big(backend, cells, cells * 9 + 5)
    pop = big(backend, pop, pop * 6)
    pr = Problem.synthetic(n, cells, pop)
    if shared:
        pr.pop_an = pr.pop_ai = None
        up, ui = simulator.upstream_from_tables(pr.net.and_len, pr.net.and_nodes, pr.net.and_inv)
    else:
        up, ui = simulator.upstream_from_tables(pr.net.and_len, pr.pop_an, pr.pop_ai)
    ref = pr.oracle()
    with pr.simulator(backend[1]) as sim:
        for mode, key in ((0, "fitness"), (1, "nodewise"), (2, "cellwise")):
            sim.upload_compact(pr.individuals, up, ui, pr.ko, pr.ki)
            sim.run(mode)
            assert np.array_equal(sim.fetch(mode), ref[key]), key


def test_compact_population_rejects_inconsistent_triples(backend):
    pr = Problem.synthetic(30, 100, 2)
    up, ui = simulator.upstream_from_tables(pr.net.and_len, pr.pop_an, pr.pop_ai)
    i = int(np.argmax(pr.net.and_len == 7))
    up[1, i, 2] = -1                                   # two inputs where andLenList says seven terms
    with pr.simulator(backend[1]) as sim:
        sim.upload_compact(pr.individuals, up, ui, pr.ko, pr.ki)
        sim.run(0)
        with pytest.raises(simulator.ScbError) as e:
            sim.fetch(0)
        assert e.value.code == _lib.ERR_INVALID


# ---- steady program (constant propagation in K0) ---------------------------------------------------
@pytest.mark.parametrize("chain", [1, 2, 3, 4, 5, 6, 8, 9, 12, 13, 23, 24, 25, 40])
@pytest.mark.parametrize("opts", [dict(), dict(steady=0)])
def test_constants_spreading_along_chains(backend, chain, opts):
    """A knocked node feeding a chain of copies: node j of the chain is constant from step j + 1 on, so the
    constant-propagation depth is the chain length -- around every depth at which K0 emits a reduced program
    (2, 5, 8, 12, the fixed point) and below, at and above the cap (SCB_STEADY_MAX_DEPTH = 24).  A reduced program
    may only compute a step once both state buffers hold its constants."""
    pr = ring_problem([5, 9], chain, big(backend, 700, 9000))
    with pr.simulator(backend[1], **opts) as sim:
        pr.check(sim)


def test_steady_program_moves_fewer_rows(backend):
    """The always-knocked node 0 and its zero-input copies erase AND-terms: the steady program must move clearly
    fewer state rows through shared memory for the same, identical, result."""
    pr = Problem.synthetic(120, big(backend, 1500, 20000), 6)
    rows = {}
    for steady in (0, 1):
        with pr.simulator(backend[1], steady=steady) as sim:
            st = pr.check(sim, modes=(0,))
            rows[steady] = st["state_rows_moved"]
    assert 0 < rows[1] < 0.8 * rows[0], rows


def test_every_truth_table(backend):
    """Every Boolean function of three inputs as a rule (a DNF of up to seven minterms, each a 3-literal AND-term
    with inversion flags), so that every entry of the LOP3 jump table -- and, through the functions that ignore an
    input, of the two-input and one-input paths -- is compared with the oracle's term-by-term evaluation."""
    luts = list(range(1, 255))                       # 0 and 255 are constants (8 minterms do not fit 7 terms)
    luts = [l for l in luts if bin(l).count("1") <= 7]
    n = 3 + len(luts)
    cells = big(backend, 300, 5000)
    net = type("Net", (), {})()
    net.n = n
    net.and_len = np.zeros(n, np.int32)
    net.parse = np.zeros(n + 1, np.int32)
    net.and_nodes = np.zeros((n, 7, 3), np.int32)
    net.and_inv = np.zeros((n, 7, 3), np.int32)
    bits = []
    for i in range(3):                               # free inputs: x <- x
        net.and_len[i] = 1
        net.and_nodes[i, 0] = [i, -1, -1]
        net.and_inv[i, 0] = [0, -1, -1]
        net.parse[i] = len(bits)
        bits.append(1)
    for k, lut in enumerate(luts):
        i = 3 + k
        net.and_len[i] = 7
        net.parse[i] = len(bits)
        terms = [m for m in range(8) if (lut >> m) & 1]
        for a in range(7):
            if a < len(terms):
                m = terms[a]
                net.and_nodes[i, a] = [0, 1, 2]
                net.and_inv[i, a] = [1 - ((m >> 2) & 1), 1 - ((m >> 1) & 1), 1 - (m & 1)]   # literal true <=> x == bit
                bits.append(1)
            else:
                net.and_nodes[i, a] = [0, 1, 2]
                net.and_inv[i, a] = [0, 0, 0]
                bits.append(0)
    net.parse[n] = len(bits)
    net.ind_len = len(bits)
    rng = np.random.default_rng(11)
    rows = (rng.random((n, cells)) < 0.5).astype(np.uint8)
    pr = Problem(net, rows, rng.permutation(n).astype(np.int32), np.asarray(bits, np.int32).reshape(1, -1))
    with pr.simulator(backend[1]) as sim:
        pr.check(sim)


@pytest.mark.parametrize("chain,want_from", [(7, [4, 7, 0, 0, 9]), (1, [0, 0, 0, 0, 3]), (14, [4, 7, 10, 14, 16]), (40, [4, 7, 10, 14, 26])])
def test_reduced_program_schedule(backend, chain, want_from):
    """K0's reduced programs on a chain fed by a knocked node: chain node j is constant from S_(j+1) on, so c(t) is the
    first t chain nodes.  Programs are emitted at depths 2, 5, 8, 12 (while the propagation still runs) and at the
    fixed point (or the cap of 24); program of depth D may compute steps >= D + 2; the fixed-point program has lost
    exactly the constant chain nodes."""
    pr = ring_problem([5], chain, 300)
    with pr.simulator(backend[1]) as sim:
        pr.check(sim, modes=(0,))
        got = [sim.programs(0, k) for k in range(5)]
    assert [g["from_k"] for g in got] == want_from
    full = got[0]["counts"]
    assert int(full.sum()) == pr.n and int(full[3]) == 1                       # the knocked chain head is the only constant rule
    depth = want_from[4] - 2
    assert int(got[4]["counts_k"][3]) == min(chain, depth)                     # constants of the last program = c(depth)
    for k, g in enumerate(got):
        if g["from_k"]:
            assert int(g["counts_k"].sum()) == pr.n
            assert int(g["counts_k"][3]) == min(chain, g["from_k"] - 2)


# ---- advisor findings of round 1 ---------------------------------------------------------------------------
def test_shared_context_from_many_threads(backend):
    """One context driven from a ThreadPool (what replacing __checkNodePossibilities inside the reference's parallel
    local search does, singleCell.py:490-492; ctypes releases the GIL): every composite call holds the context's lock
    across upload + run + fetch, so concurrent populations cannot interleave."""
    from multiprocessing.pool import ThreadPool
    pr = Problem.synthetic(24, big(backend, 300, 3000), 8)
    o = pr.oracle()
    with pr.simulator(backend[1]) as sim:
        def one(k):
            p = k % 8
            got = sim.evaluate_population(pr.individuals[p:p + 1], pr.ko, pr.ki, pr.pop_an[p:p + 1], pr.pop_ai[p:p + 1],
                                          mode=1 + (k % 2))
            want = o["nodewise"][p] if k % 2 == 0 else o["cellwise"][p]
            return np.array_equal(got[0], want)
        with ThreadPool(6) as pool:
            ok = pool.map(one, range(32))
    assert all(ok), ok


def test_duplicate_node_positions_are_refused(backend):
    """Two nodes on one gene row alias one simData column in the reference (the later node overwrites the earlier one
    in the attractor search and the error); the packed kernels keep nodes independent, so such input is refused."""
    pr = Problem.synthetic(12, 64, 1)
    pos = pr.node_positions.copy()
    pos[3] = pos[7]
    with pytest.raises(simulator.ScbError) as ei:
        simulator.Simulator(pr.bm, tables(pr.net, pos), lib_path=backend[1])
    assert ei.value.code == _lib.ERR_UNSUPPORTED and "twice" in str(ei.value)


def test_result_pointer_needs_a_matching_run(backend):
    pr = Problem.synthetic(12, 64, 2)
    with pr.simulator(backend[1]) as sim:
        sim.upload(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai)
        with pytest.raises(simulator.ScbError):
            sim.result_device_pointer(simulator.MODE_SUM)                 # nothing has run yet
        sim.run(simulator.MODE_SUM)
        sim.result_device_pointer(simulator.MODE_SUM)
        with pytest.raises(simulator.ScbError):
            sim.result_device_pointer(simulator.MODE_NODEWISE)            # not what the last run produced


_LEGACY_FAIL_SCRIPT = r"""
import ctypes, sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from common import Problem
from scbonita_b200 import _lib
pr = Problem.synthetic(12, 40, 1)
lib = _lib.load(%(lib)r)
lib.scb_set_legacy_geometry(64, 64)
bm = np.zeros((pr.bm.shape[0], 64), np.int32); bm[:, :40] = pr.bm
V = ctypes.c_void_p
ind, al, parse, an, ai, ko, ki, pos = [np.ascontiguousarray(a, np.int32) for a in (pr.individuals[0], pr.net.and_len, pr.net.parse,
                                       pr.pop_an[0], pr.pop_ai[0], pr.ko, pr.ki, pr.node_positions)]
fn = lib.scSyncBool; fn.restype = None; fn.argtypes = None
for local in (0, 1):
    errors = np.zeros(64, np.int32)
    fn(V(0), V(ind.ctypes.data), V(len(ind)), V(pr.n), V(al.ctypes.data), V(parse.ctypes.data), V(an.ctypes.data), V(ai.ctypes.data),
       V(50), V(ko.ctypes.data), V(ki.ctypes.data), V(40), V(bm.ctypes.data), V(pos.ctypes.data), V(errors.ctypes.data), V(local), V(0))
    n = pr.n if local else 40
    print("SUM", int(errors[:n].astype(np.int64).sum()), int(errors[:n].min()))
"""


@pytest.mark.parametrize("how", ["abort", "sentinel"])
def test_legacy_symbol_failure_cannot_look_like_a_result(backend, how):
    """A failed drop-in scSyncBool (here: simSteps != 100) must never leave `errors` as the caller's zeros -- the
    unmodified Python would score sum(errors) == 0 as a perfect individual.  Default: message + abort();
    SCB_LEGACY_ON_ERROR=sentinel: every entry a value no evaluation can reach."""
    import subprocess
    import sys as _sys
    env = dict(os.environ)
    if how == "sentinel":
        env["SCB_LEGACY_ON_ERROR"] = "sentinel"
    else:
        env.pop("SCB_LEGACY_ON_ERROR", None)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, "-c", _LEGACY_FAIL_SCRIPT % dict(root=root, lib=backend[1])], env=env,
                       capture_output=True, text=True, timeout=600)
    assert "scSyncBool failed" in r.stderr and "simSteps" in r.stderr, r.stderr
    if how == "abort":
        assert r.returncode != 0 and "SUM" not in r.stdout
    else:
        assert r.returncode == 0, r.stderr
        sums = [line.split() for line in r.stdout.splitlines() if line.startswith("SUM")]
        assert len(sums) == 2
        for _, total, smallest in sums:
            assert int(total) > 10 ** 9 and int(smallest) > 12 * 40         # far beyond any reachable error


# ---- resident population -------------------------------------------------------------------------
@pytest.mark.parametrize("compact", [False, True])
def test_update_replaces_and_scores_only_the_given_slots(backend, compact):
    pr = Problem.synthetic(60, big(backend, 700, 9000), 8)
    o = pr.oracle()
    with pr.simulator(backend[1]) as sim:
        if compact:
            up, upinv = zip(*(simulator.upstream_from_tables(pr.net.and_len, pr.pop_an[i], pr.pop_ai[i]) for i in range(8)))
            up, upinv = np.array(up, np.int32), np.array(upinv, np.int32)
            sim.upload_compact(pr.individuals, up, upinv, pr.ko, pr.ki)
        else:
            sim.upload(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai)
        sim.run()
        assert np.array_equal(sim.fetch(), o["fitness"])
        tiles_full = sim.last_stats["tiles"]
        slots = np.array([6, 1, 3], np.int32)
        src = np.array([0, 2, 7])
        new_inds = pr.individuals.copy()
        new_an, new_ai = pr.pop_an.copy(), pr.pop_ai.copy()
        new_inds[slots], new_an[slots], new_ai[slots] = pr.individuals[src], pr.pop_an[src], pr.pop_ai[src]
        new_inds[6] = 1 - new_inds[6]                                             # and one rule set nobody has seen
        m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
        want = oracle.packed_population(m, new_inds, pr.bm, pr.n_cells, and_nodes=new_an, and_inv=new_ai)["fitness"]
        if compact:
            sim.update(slots, new_inds[slots], up[src], upinv[src])
        else:
            sim.update(slots, new_inds[slots], new_an[slots], new_ai[slots])
        sim.run()
        got = sim.fetch()
        assert np.array_equal(got, want)
        assert sim.last_stats["tiles"] * 8 == 3 * tiles_full                      # 3 of 8 individuals were simulated
        with pytest.raises(simulator.ScbError):
            sim.update([1, 1], new_inds[:2], new_an[:2], new_ai[:2])              # a slot twice
        with pytest.raises(simulator.ScbError):
            sim.update([8], new_inds[:1], new_an[:1], new_ai[:1])                 # outside the population
        sim.update([0], new_inds[:1], (up if compact else new_an)[:1], (upinv if compact else new_ai)[:1])
        with pytest.raises(simulator.ScbError):
            sim.run(simulator.MODE_CELLWISE)                                      # only the sums are kept current
        sim.run()
        assert np.array_equal(sim.fetch(), want)



def test_legacy_symbol_with_a_network_beyond_shared_memory(backend):
    """scSyncBool on a 900-node network with rules over up to five distinct inputs: both reasons for the general path."""
    rng = np.random.default_rng(11)
    n, cells = 900, 150
    net = _random_tables(rng, n, max_distinct=5, weird_flags=False)
    rows = (rng.random((n, cells)) < 0.5).astype(np.uint8)
    ind = (rng.random((1, net.ind_len)) < 0.7).astype(np.int32)
    pr = Problem(net, rows, np.arange(n, dtype=np.int32), ind)
    o = pr.oracle()
    lib = _lib.load(backend[1])
    stride = 256
    lib.scb_set_legacy_geometry(20000, stride)
    try:
        bm = np.zeros((n, stride), np.int32)
        bm[:, :cells] = pr.bm
        V = ctypes.c_void_p
        vals = np.zeros((100, n), np.int32)
        arr = [np.ascontiguousarray(a, np.int32) for a in (pr.individuals[0], pr.net.and_len, pr.net.parse, pr.net.and_nodes,
                                                           pr.net.and_inv, pr.ko, pr.ki, pr.node_positions)]
        ind0, al, parse, an, ai, ko, ki, pos = arr
        fn = lib.scSyncBool
        fn.restype, fn.argtypes = None, None
        errors = np.full(cells + 1, -1, np.int32)
        fn(V(vals.ctypes.data), V(ind0.ctypes.data), V(len(ind0)), V(n), V(al.ctypes.data), V(parse.ctypes.data),
           V(an.ctypes.data), V(ai.ctypes.data), V(100), V(ko.ctypes.data), V(ki.ctypes.data), V(cells),
           V(bm.ctypes.data), V(pos.ctypes.data), V(errors.ctypes.data), V(0), V(0))
        assert np.array_equal(errors[:cells], o["cellwise"][0]) and errors[cells] == -1
    finally:
        lib.scb_set_legacy_geometry(20000, 15000)


def test_q8_extra_accumulation_from_1000_nodes(backend):
    """SURVEY Q8 (simulator.c:457-468): with nodeNum >= 1000 the threshold 0.001 * nodeNum reaches 1 and the first
    window state with a new-minimum error count at or below it adds its error vector to the per-node sums once more.
    A 1200-node network of self-copies with two negated copies and one self-negating node (period 2): cells with one
    wrong node are counted twice, cells with two are not, and the order of the two window states matters."""
    class Net:
        pass
    n, cells = 1200, 96
    net = Net()
    net.n = n
    net.and_len = np.zeros(n, np.int32)
    net.parse = np.zeros(n + 1, np.int32)
    net.and_nodes = np.zeros((n, 7, 3), np.int32)
    net.and_inv = np.zeros((n, 7, 3), np.int32)
    net.and_nodes[:, 0, 0] = np.arange(n)                      # andLen 0: a copy of andNodes[i][0][0] = itself (Q2)
    for node, src in ((1, 2), (3, 4), (5, 5)):                 # node <- NOT src; node 5 negates itself: period 2
        net.and_len[node] = 1
        net.and_nodes[node, 0] = [src, -1, -1]
        net.and_inv[node, 0] = [1, -1, -1]
    net.parse[:] = np.concatenate([[0], np.cumsum(net.and_len)])
    net.ind_len = int(net.parse[n])
    rng = np.random.default_rng(5)
    rows = (rng.random((n, cells)) < 0.5).astype(np.uint8)
    ind = np.ones((1, max(net.ind_len, 1)), np.int32)[:, :net.ind_len] if net.ind_len else np.zeros((1, 0), np.int32)
    pr = Problem(net, rows, np.arange(n, dtype=np.int32), ind, ko=np.zeros(n, np.int32), ki=np.zeros(n, np.int32))
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
    want, _ = oracle.sync_bool(m, pr.individuals[0], pr.bm, cells, 1)          # the scalar restatement implements Q8
    plain = rows[1] == rows[2]
    assert want[1] > int(plain.sum())                                           # the case really exercises the quirk
    with pr.simulator(backend[1]) as sim:
        got = sim.evaluate_population(pr.individuals, pr.ko, pr.ki, None, None, mode=simulator.MODE_NODEWISE)
        assert np.array_equal(got[0], want)
        cw = sim.evaluate_population(pr.individuals, pr.ko, pr.ki, None, None, mode=simulator.MODE_CELLWISE)
        wc, _ = oracle.sync_bool(m, pr.individuals[0], pr.bm, cells, 0)
        assert np.array_equal(cw[0][:cells], wc[:cells])


@pytest.mark.gpu
def test_population_beyond_the_device_side_ordering(cuda_lib):
    """More individuals than the one-CTA rank sort orders (SCB_ORDER_MAX_P = 4096): index order, same integers; and the
    same population with the predicted-cost order switched off and with an explicit order."""
    pr = Problem.synthetic(16, 700, 4200)
    want = pr.oracle()["fitness"]
    with pr.simulator(cuda_lib) as sim:
        assert np.array_equal(sim.evaluate_population(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai), want)
    sl = slice(0, 3000)
    with pr.simulator(cuda_lib) as sim:
        a = sim.evaluate_population(pr.individuals[sl], pr.ko, pr.ki, pr.pop_an[sl], pr.pop_ai[sl])          # ordered by K0's prediction
        sim.set_option(_lib.OPT_AUTO_ORDER, 0)
        b = sim.evaluate_population(pr.individuals[sl], pr.ko, pr.ki, pr.pop_an[sl], pr.pop_ai[sl])          # index order
        sim.set_order(sim.cost_order())
        c = sim.evaluate_population(pr.individuals[sl], pr.ko, pr.ki, pr.pop_an[sl], pr.pop_ai[sl])          # measured order
    assert np.array_equal(a, want[sl]) and np.array_equal(b, want[sl]) and np.array_equal(c, want[sl])

"""Generates the golden vectors under tests/golden/ (run in the build container, where /root/reference and
oracle/_ref exist):   python tests/golden/make_golden.py

* sim_*.npz -- seeded inputs in the reference's int32 layouts plus the outputs of the UNMODIFIED reference
  C (oracle/_ref/simulator.so driven with the reference's own ctypes convention): errors[] of scSyncBool
  for localSearch = 0 (cellwise) and 1 (nodewise), for every individual.
* imp_*.npz / cluster_*.npz -- the same for the reference's importanceScore (one double per knock-out / knock-in
  pair) and `cluster` (attractor window + the scratch the call leaves behind, incl. knock-outs on permuted
  nodePositions, simulator.c:518).
* hamming_kat.npz -- the reference's own packaged known answer for the distance / decider stage:
  src/scBONITA/data/hsa00010.graphml_processed.graphml_attractorDistance.csv (100 cells x 51 attractors +
  deciderDistance + decider) together with the pickled binarized matrix it was computed from.  The script
  asserts that sum|attractor - cell| recomputed from the matrix reproduces the CSV before writing it.
"""
import ast
import csv
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from common import Problem, oracle_model, ring_problem  # noqa: E402

REF_DATA = "/root/reference/src/scBONITA/data"


def ref_outputs(pr):
    ref = oracle.RefSimulator(pr.n_cells)
    m = oracle_model(pr.net, pr.node_positions, pr.ko, pr.ki)
    bm = ref.dense(pr.bin_rows.astype(np.int32), pr.node_positions)
    cw, nw = [], []
    for p in range(pr.individuals.shape[0]):
        an = None if pr.pop_an is None else pr.pop_an[p]
        ai = None if pr.pop_ai is None else pr.pop_ai[p]
        cw.append(ref.sync_bool(m, pr.individuals[p], bm, pr.n_cells, 0, and_nodes=an, and_inv=ai)[:pr.n_cells].copy())
        nw.append(ref.sync_bool(m, pr.individuals[p], bm, pr.n_cells, 1, and_nodes=an, and_inv=ai)[:pr.n].copy())
    return np.array(cw, np.int32), np.array(nw, np.int32)


def save_sim(name, pr):
    cw, nw = ref_outputs(pr)
    per_ind = pr.pop_an is not None
    np.savez_compressed(
        os.path.join(HERE, name), and_len=pr.net.and_len, parse=pr.net.parse, and_nodes=pr.net.and_nodes,
        and_inv=pr.net.and_inv, node_positions=pr.node_positions, ind_len=pr.net.ind_len, knockouts=pr.ko,
        knockins=pr.ki, n_cells=pr.n_cells, bin_rows=np.packbits(pr.bin_rows, axis=1), individuals=pr.individuals,
        pop_and_nodes=pr.pop_an if per_ind else np.zeros(1, np.int32),
        pop_and_inv=pr.pop_ai if per_ind else np.zeros(1, np.int32), ref_cellwise=cw, ref_nodewise=nw)
    print(name, "fitness", cw.sum(axis=1))


def hamming_kat():
    import networkx as nx

    class _U(pickle.Unpickler):
        def find_class(self, module, name):
            if module in ("singleCell", "ruleMaker", "__main__"):
                return type(name, (object,), {})
            return super().find_class(module, name)

    sc = _U(open(os.path.join(REF_DATA, "trainingData.csvscTest.pickle"), "rb")).load()
    genes = [str(g) for g in sc.geneList]
    n_cells = len(sc.sampleList)
    graph = nx.read_graphml(os.path.join(REF_DATA, "hsa00010.graphml_processed.graphml"))
    node_list = [n for n in graph.nodes() if n in genes]           # ruleMaker.__init__ keeps graph order
    pos = np.array([genes.index(n) for n in node_list], np.int32)
    dense = np.asarray(sc.binMat[: len(genes), :n_cells].todense()).astype(np.int32)
    with open(os.path.join(REF_DATA, "hsa00010.graphml_processed.graphml_attractorDistance.csv")) as f:
        rows = list(csv.reader(f))
    header = rows[0][1:]
    attractors = np.array([ast.literal_eval(h) for h in header[:-2]], np.int32)
    body = np.array([[int(float(x)) for x in r[1:]] for r in rows[1:]], np.int64)
    dist, dd, dec = body[:, :-2].astype(np.int32), body[:, -2].astype(np.int32), body[:, -1]
    assert attractors.shape[1] == len(node_list), (attractors.shape, len(node_list))
    mine = np.abs(attractors[None, :, :] - dense[pos].T[:, None, :]).sum(axis=2)
    assert np.array_equal(mine, dist), "packaged distances do not reproduce"
    assert np.array_equal(mine.min(axis=1), dd)
    # `decider` holds the column label (attractor string) or index depending on pandas; derive first-min index
    decider = mine.argmin(axis=1).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "hamming_kat.npz"), attractors=attractors, bin_mat=dense,
                        node_positions=pos, n_cells=n_cells, dist=dist, decider=decider, decider_distance=dd)
    print("hamming_kat: %d cells x %d attractors reproduce the packaged CSV" % dist.shape)
    return rows




# ---------------------------------------------------------------------------------------------
# importanceScore (simulator.c:555-755) and cluster (simulator.c:487-552) from the unmodified reference build
# ---------------------------------------------------------------------------------------------
def _knock_rows(n, lists):
    out = np.zeros((len(lists), n), np.int32)
    for k, nodes in enumerate(lists):
        out[k, list(nodes)] = 1
    return out


def save_importance(name, pr, knock_lists):
    """One reference importanceScore call per (knock-out list, knock-in list).  Items for which a cell has no
    attractor window are left out: the reference divides by zero there (SIGFPE, simulator.c:647)."""
    ref = oracle.RefSimulator(pr.n_cells)
    bm = ref.dense(pr.bin_rows.astype(np.int32), pr.node_positions)
    an = pr.net.and_nodes if pr.pop_an is None else pr.pop_an[0]
    ai = pr.net.and_inv if pr.pop_ai is None else pr.pop_ai[0]
    kos, kis, scores = [], [], []
    for ko_nodes, ki_nodes in knock_lists:
        ko, ki = _knock_rows(pr.n, [ko_nodes])[0], _knock_rows(pr.n, [ki_nodes])[0]
        m = oracle.Model(pr.net.and_len, pr.net.parse, an, ai, pr.node_positions, pr.net.ind_len, ko, ki)
        try:
            oracle.importance_score(m, pr.individuals[0], pr.bm, pr.n_cells)
        except ValueError:
            continue
        kos.append(ko); kis.append(ki)
        scores.append(ref.importance_score(m, pr.individuals[0], bm, pr.n_cells))
    np.savez_compressed(os.path.join(HERE, name), and_len=pr.net.and_len, parse=pr.net.parse, and_nodes=an, and_inv=ai,
                        node_positions=pr.node_positions, ind_len=pr.net.ind_len, n_cells=pr.n_cells,
                        bin_rows=np.packbits(pr.bin_rows, axis=1), individual=pr.individuals[0],
                        knockouts=np.array(kos, np.int32), knockins=np.array(kis, np.int32),
                        ref_scores=np.array(scores, np.float64))
    print(name, "%d items" % len(scores), scores[:4])


def save_cluster(name, pr, knock_lists, cells):
    """The reference's `cluster` for a few cells per knock list, on a zeroed scratch like the reference's Python
    passes: the attractor window and the first `stride` columns of the scratch it leaves behind."""
    ref = oracle.RefSimulator(pr.n_cells)
    bm = ref.dense(pr.bin_rows.astype(np.int32), pr.node_positions)
    an = pr.net.and_nodes if pr.pop_an is None else pr.pop_an[0]
    ai = pr.net.and_inv if pr.pop_ai is None else pr.pop_ai[0]
    stride = max(int(pr.node_positions.max()) + 1, pr.n)
    kos = _knock_rows(pr.n, [k for k, _ in knock_lists])
    kis = _knock_rows(pr.n, [k for _, k in knock_lists])
    sims = np.zeros((len(knock_lists), len(cells), 100, stride), np.int8)
    ress = np.zeros((len(knock_lists), len(cells), 2), np.int32)
    for k in range(len(knock_lists)):
        m = oracle.Model(pr.net.and_len, pr.net.parse, an, ai, pr.node_positions, pr.net.ind_len, kos[k], kis[k])
        for ci, cell in enumerate(cells):
            vals, res = ref.cluster(m, pr.individuals[0], bm, int(cell))
            assert not vals[:, stride:].any()
            sims[k, ci] = vals[:, :stride]
            ress[k, ci] = res
    np.savez_compressed(os.path.join(HERE, name), and_len=pr.net.and_len, parse=pr.net.parse, and_nodes=an, and_inv=ai,
                        node_positions=pr.node_positions, ind_len=pr.net.ind_len, n_cells=pr.n_cells,
                        bin_rows=np.packbits(pr.bin_rows, axis=1), individual=pr.individuals[0],
                        knockouts=kos, knockins=kis, cells=np.asarray(cells, np.int32), stride=stride,
                        ref_sim=sims, ref_res=ress)
    print(name, "windows", sorted({tuple(r) for r in ress.reshape(-1, 2)})[:8])


# ---------------------------------------------------------------------------------------------
# Python attractor path (singleCell.__cluster, singleCell.py:990-1059) run from the reference itself
# ---------------------------------------------------------------------------------------------
def import_reference_single_cell():
    """Imports /root/reference/src/scBONITA/singleCell.py with stand-ins for the plotting / GA / KEGG
    packages that are not installed here (none of them takes part in __cluster)."""
    from unittest.mock import MagicMock
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import deap_standin
    deap_standin.install()                       # functional stand-in for the GA containers (deap is not installed)
    for name in ("seaborn",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "statsmodels", "statsmodels.stats",
                 "statsmodels.stats.multitest", "bioservices", "bioservices.kegg", "bs4", "requests"):
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, "/root/reference/src/scBONITA")
    import singleCell as ref_sc
    return ref_sc


def ragged(net, pop_an, pop_ai):
    """andNodeList / andNodeInvertList as ruleMaker keeps them (ragged lists, bools)."""
    an, ai = [], []
    for i in range(net.n):
        k = int(net.and_len[i])
        an.append([[int(x) for x in pop_an[i, a] if x > -1] for a in range(k)])
        ai.append([[bool(f) for x, f in zip(pop_an[i, a], pop_ai[i, a]) if x > -1] for a in range(k)])
    return an, ai


def save_py_cluster(name, n, cells, seed):
    from scbonita_b200 import synth
    ref_sc = import_reference_single_cell()
    pr = synth.make_problem(n, cells, 1, knock=False, net_seed=seed, cell_seed=seed + 1, pop_seed=seed + 2)
    bm = np.zeros((n, cells), np.int32)
    bm[pr.node_positions] = pr.bin_rows

    stub = object.__new__(ref_sc.singleCell)          # the real class, without running its __init__
    stub.maxNodes = n + 5
    stub.nodeList = list(range(n))
    an, ai = ragged(pr.net, pr.pop_and_nodes[0], pr.pop_and_inv[0])
    ko = np.zeros(n, np.intc)
    ki = np.zeros(n, np.intc)
    ki[n // 2] = 1
    all_vals, all_res = [], []
    for c in range(cells):
        vals = np.full((10, n), 0, dtype=np.intc)
        res = np.full(2, 2, dtype=np.intc)
        res = ref_sc.singleCell._singleCell__cluster(stub, vals, res, c, pr.individuals[0], pr.net.ind_len, n,
                                                     pr.net.and_len, pr.net.parse, an, ai, 10, ko, ki, bm,
                                                     list(pr.node_positions))
        all_vals.append(vals.copy())
        all_res.append([int(res[0]), int(res[1])])
    np.savez_compressed(os.path.join(HERE, name), and_len=pr.net.and_len, parse=pr.net.parse,
                        and_nodes=pr.pop_and_nodes[0], and_inv=pr.pop_and_inv[0], node_positions=pr.node_positions,
                        ind_len=pr.net.ind_len, knockouts=ko, knockins=ki, n_cells=cells,
                        bin_rows=np.packbits(pr.bin_rows, axis=1), individual=pr.individuals[0],
                        ref_vals=np.array(all_vals, np.int8), ref_res=np.array(all_res, np.int32))
    print(name, "res histogram", {tuple(r): all_res.count(r) for r in map(list, set(map(tuple, all_res)))})


if __name__ == "__main__":
    oracle.build(ref=True)
    save_sim("sim_kegg60.npz", Problem.synthetic(60, 700, 4))
    save_sim("sim_hsa47_noknock.npz", Problem.synthetic(47, 100, 3, knock=False))
    save_sim("sim_n200.npz", Problem.synthetic(200, 130, 3))
    save_sim("sim_ring45_chain12.npz", ring_problem([45], 12, 1500))
    save_sim("sim_ring60_undefined_lead.npz", ring_problem([60], 0, 200, lead_found=True))
    hamming_kat()
    pr = Problem.synthetic(33, 400, 1, knock=False, net_seed=31)
    save_importance("imp_n33.npz", pr, [([i], [i]) for i in range(33)] + [([1], [32]), ([], []), ([3, 9], [4, 5, 6])])
    save_importance("imp_ring.npz", ring_problem([40, 5], 3, 300), [([i], [i]) for i in range(0, 49, 3)])
    save_cluster("cluster_n33.npz", pr, [([], []), ([], [16]), ([2], []), ([0, 32], [1]), ([11, 22], [])], [0, 7, 200, 399])
    save_cluster("cluster_ring.npz", ring_problem([7, 4], 3, 50), [([], []), ([12], []), ([1, 9], [3])], [0, 7, 21, 49])
    save_py_cluster("pycluster_n24.npz", 24, 70, 501)
    save_py_cluster("pycluster_n40.npz", 40, 50, 777)
    save_ga_trace()


# ---------------------------------------------------------------------------------------------
# C1 on the GPU box: a trace of what the UNMODIFIED reference GA hands to scSyncBool
# ---------------------------------------------------------------------------------------------
def save_ga_trace(name="ga_trace_hsa00010.npz", seed=20260, pop=12, lambd=12, mu=6, gens=3):
    """Runs ruleMaker.__eaMuPlusLambdaAdaptive (unmodified, packaged 47-gene x 100-cell example, deap stand-in) with
    ./simulator.so = the unmodified reference C, and records every scSyncBool call at the C boundary: the bit-string,
    the int32 [N][7][3] tables exactly as the C code reads them (the shipped wrapper passes andNodeInvertList from a
    dtype=object array, SURVEY Q3: the C side sees pointer halves, kept here byte for byte), and the errors it wrote.
    The GPU box replays the trace through the CUDA library: same numbers, same logbook."""
    import copy
    import ctypes
    import random
    import tempfile
    import warnings
    import networkx as nx
    import scipy.sparse as sp
    import_reference_single_cell()
    ref_dir = "/root/reference/src/scBONITA"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sc = pickle.load(open(os.path.join(ref_dir, "data", "trainingData.csvscTest.pickle"), "rb"))
        net = nx.read_graphml(os.path.join(ref_dir, "data", "hsa00010.graphml_processed.graphml"))
        sc._singleCell__inherit(net, removeSelfEdges=False, restrictIncomingEdges=True, maxIncomingEdges=3, groundTruth=False)
    sc._ruleMaker__updateCpointers()
    sc._singleCell__setupEmptyKOKI()
    rows = max(sc.nodePositions) + 1
    n_cells = len(sc.sampleList)
    dense = sc.binMat[:rows, :n_cells].toarray()
    bm = np.zeros((rows, 15000))
    bm[:, :dense.shape[1]] = dense
    sc.binMat = sp.csr_matrix(bm)
    real = ctypes.cdll.LoadLibrary(os.path.join(ROOT, "oracle", "_ref", "simulator.so"))
    N = len(sc.nodeList)
    calls = []

    def grab(ptr, count):
        return np.frombuffer(ctypes.string_at(ptr.value, count * 4), dtype=np.int32).copy()

    def recording_scSyncBool(*a):
        ind_len, node_num, n_samp, local = (a[2].value or 0), (a[3].value or 0), (a[11].value or 0), (a[15].value or 0)
        rec = dict(individual=grab(a[1], ind_len), and_len=grab(a[4], node_num), parse=grab(a[5], node_num),
                   and_nodes=grab(a[6], node_num * 21), and_inv=grab(a[7], node_num * 21), knockouts=grab(a[9], node_num),
                   knockins=grab(a[10], node_num), node_positions=grab(a[13], node_num), local=local, n_cells=n_samp)
        real.scSyncBool(*a)
        rec["errors"] = grab(a[14], node_num if local else n_samp)
        calls.append(rec)

    class Proxy:
        scSyncBool = staticmethod(recording_scSyncBool)
        importanceScore = real.importanceScore

    model = copy.deepcopy(sc)
    model._ruleMaker__makeToolBox(None)
    model.params.popSize, model.params.lambd, model.params.mu, model.params.generations = pop, lambd, mu, gens
    random.seed(seed)
    np.random.seed(seed)
    orig = ctypes.cdll.LoadLibrary
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        ctypes.cdll.LoadLibrary = lambda path: Proxy if "simulator.so" in str(path) else orig(path)
        try:
            population, logbook = model._ruleMaker__eaMuPlusLambdaAdaptive(None, "graph", verbose=False)
        finally:
            ctypes.cdll.LoadLibrary = orig
            os.chdir(cwd)
    assert all(c["local"] == 0 and c["n_cells"] == n_cells for c in calls)
    for key in ("and_len", "parse", "knockouts", "knockins", "node_positions"):
        assert all(np.array_equal(c[key], calls[0][key]) for c in calls), key
    nevals = [int(r["nevals"]) for r in logbook]
    assert sum(nevals) == len(calls)
    pos = calls[0]["node_positions"]
    np.savez_compressed(
        os.path.join(HERE, name), n_cells=n_cells, ind_len=len(calls[0]["individual"]), and_len=calls[0]["and_len"],
        parse=calls[0]["parse"], knockouts=calls[0]["knockouts"], knockins=calls[0]["knockins"], node_positions=pos,
        bin_rows=np.packbits(dense[pos].astype(np.uint8), axis=1),
        individuals=np.array([c["individual"] for c in calls], np.int32),
        and_nodes=np.array([c["and_nodes"].reshape(N, 7, 3) for c in calls], np.int32),
        and_inv=np.array([c["and_inv"].reshape(N, 7, 3) for c in calls], np.int32),
        fitness=np.array([int(c["errors"].astype(np.int64).sum()) for c in calls], np.int64),
        cellwise=np.array([c["errors"] for c in calls], np.int32),
        nevals=np.array(nevals, np.int32),
        logbook=np.array([(int(r["gen"]), int(r["nevals"]), float(r["avg"]), float(r["std"]), float(r["min"]), float(r["max"])) for r in logbook], np.float64),
        final_fitness=np.array([float(ind.fitness.values[0]) for ind in population], np.float64))
    print(name, "%d scSyncBool calls" % len(calls), "nevals", nevals, "min per generation", [float(r["min"]) for r in logbook])

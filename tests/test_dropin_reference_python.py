"""Level-0 drop-in: the reference's own, UNMODIFIED Python marshaller
``ruleMaker.__evaluateByNode`` (ruleMaker.py:1077-1221) is run on the packaged 47-gene x 100-cell example with
``cFunction`` bound once to the reference C (oracle/_ref/simulator.so) and once to this library's exported
``scSyncBool`` / ``importanceScore`` (host-thread emulator build on CPU).  Every returned number must agree --
including the effect of the wrapper's ``dtype=object`` inversion table, which the shim
reinterprets byte for byte like the C code does.

Needs /root/reference (build container only); skipped elsewhere."""
import ctypes
import os
import pickle
import sys
import warnings

import numpy as np
import pytest

from common import EMU_LIB, ROOT

REF = "/root/reference/src/scBONITA"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ruleMaker.py")) or
                                not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "simulator.so")),
                                reason="needs the reference tree and oracle/_ref")


@pytest.fixture(scope="module")
def model():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import import_reference_single_cell
    import_reference_single_cell()
    import networkx as nx
    import scipy.sparse as sp
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sc = pickle.load(open(os.path.join(REF, "data", "trainingData.csvscTest.pickle"), "rb"))
        net = nx.read_graphml(os.path.join(REF, "data", "hsa00010.graphml_processed.graphml"))
        sc._singleCell__inherit(net, removeSelfEdges=False, restrictIncomingEdges=True, maxIncomingEdges=3,
                                groundTruth=False)
    sc._ruleMaker__updateCpointers()
    sc._singleCell__setupEmptyKOKI()
    # keep only the gene rows the network uses; the row stride must be the reference's compile-time CELL
    rows = max(sc.nodePositions) + 1
    dense = sc.binMat[:rows, :len(sc.sampleList)].toarray()
    bm = np.zeros((rows, 15000))
    bm[:, :dense.shape[1]] = dense
    sc.binMat = sp.csr_matrix(bm)
    out2 = pickle.load(open(os.path.join(REF, "data", "hsa00010.graphml_processed.graphml_out2.pickle"), "rb"))
    return sc, [int(x) for x in out2]


def test_unmodified_reference_marshaller_gets_identical_numbers(model, emu_lib):
    sc, ind = model
    ref = ctypes.cdll.LoadLibrary(os.path.join(ROOT, "oracle", "_ref", "simulator.so"))
    ours = ctypes.cdll.LoadLibrary(emu_lib)
    rng = np.random.default_rng(1)
    individuals = [ind]
    for _ in range(2):                                    # two more rule sets, >= 1 bit per node kept
        alt = list(ind)
        for k in rng.choice(len(alt), size=25, replace=False):
            alt[k] = 1
        individuals.append(alt)
    for cand in individuals:
        want_ga = sc._ruleMaker__evaluateByNode(cand, sc.knockoutLists, sc.knockinLists, ref.scSyncBool)
        got_ga = sc._ruleMaker__evaluateByNode(cand, sc.knockoutLists, sc.knockinLists, ours.scSyncBool)
        assert got_ga == want_ga
        want_ls = sc._ruleMaker__evaluateByNode(cand, sc.knockoutLists, sc.knockinLists, ref.scSyncBool, localSearch=True)
        got_ls = sc._ruleMaker__evaluateByNode(cand, sc.knockoutLists, sc.knockinLists, ours.scSyncBool, localSearch=True)
        assert got_ls == want_ls
    assert sc._ruleMaker__evaluateByNode(ind, sc.knockoutLists, sc.knockinLists, ours.scSyncBool) == [3017]   # SURVEY Q3 probe


def test_unmodified_reference_local_search_node(model, emu_lib):
    """__checkNodePossibilities (ruleMaker.py:1235-1266) for one 3-bit node through both libraries."""
    sc, ind = model
    ref = ctypes.cdll.LoadLibrary(os.path.join(ROOT, "oracle", "_ref", "simulator.so"))
    ours = ctypes.cdll.LoadLibrary(emu_lib)
    node = next(i for i in range(len(sc.nodeList)) if sc._ruleMaker__findEnd(i) - sc.individualParse[i] == 3)
    want = sc._ruleMaker__checkNodePossibilities(node, ind, sc.knockoutLists, sc.knockinLists, ref.scSyncBool)
    got = sc._ruleMaker__checkNodePossibilities(node, ind, sc.knockoutLists, sc.knockinLists, ours.scSyncBool)
    assert got == want


def test_unmodified_reference_importance_call(model, emu_lib):
    sc, ind = model
    ref = ctypes.cdll.LoadLibrary(os.path.join(ROOT, "oracle", "_ref", "simulator.so"))
    ours = ctypes.cdll.LoadLibrary(emu_lib)
    want = sc._ruleMaker__evaluateByNode(ind, [3], [], ref.importanceScore, importanceScores=True)
    got = sc._ruleMaker__evaluateByNode(ind, [3], [], ours.importanceScore, importanceScores=True)
    assert np.float64(got[0]).tobytes() == np.float64(want[0]).tobytes(), (got, want)


def _run_reference_ga(sc, so_path, tmp_path, seed):
    """ruleMaker.__eaMuPlusLambdaAdaptive loads "./simulator.so" relative to the working directory
    (ruleMaker.py:979): the drop-in is literally a file swap."""
    import random
    import copy
    workdir = tmp_path / ("run_" + os.path.basename(so_path))
    workdir.mkdir()
    os.symlink(so_path, workdir / "simulator.so")
    model = copy.deepcopy(sc)
    model._ruleMaker__makeToolBox(None)
    model.params.popSize, model.params.lambd, model.params.mu, model.params.generations = 8, 8, 4, 2
    random.seed(seed)
    np.random.seed(seed)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        population, logbook = model._ruleMaker__eaMuPlusLambdaAdaptive(None, "graph", verbose=False)
    finally:
        os.chdir(cwd)
    return [list(ind[1]) for ind in population], [list(ind.fitness.values) for ind in population], \
           [(rec["gen"], rec["nevals"], float(rec["min"]), float(rec["max"])) for rec in logbook]


def test_unmodified_reference_ga_run_is_reproduced(model, emu_lib, tmp_path, capsys):
    """The whole GA driver (population init, adaptive mutation, NSGA-II selection, logbook), unmodified, with only
    ./simulator.so swapped: same final population, same fitness values, same logbook."""
    sc, _ = model
    if not hasattr(sc, "params") or not hasattr(sc.params, "popSize"):
        pytest.skip("reference model has no GA parameters")
    want = _run_reference_ga(sc, os.path.join(ROOT, "oracle", "_ref", "simulator.so"), tmp_path, 20260)
    got = _run_reference_ga(sc, emu_lib, tmp_path, 20260)
    capsys.readouterr()
    assert got[2] == want[2]          # logbook: gen, nevals, min, max
    assert got[1] == want[1]          # fitness of the survivors
    assert got[0] == want[0]          # their rule bit-strings

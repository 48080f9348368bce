"""Config C1 on the GPU box: replay of an UNMODIFIED reference GA run.

``tests/golden/ga_trace_hsa00010.npz`` (made by tests/golden/make_golden.py::save_ga_trace in the build container)
holds every scSyncBool call ``ruleMaker.__eaMuPlusLambdaAdaptive`` (ruleMaker.py:925-1022) issued on the packaged
47-gene x 100-cell example with ./simulator.so = the unmodified reference C: the bit-strings, the int32 tables
exactly as the C code read them (the shipped wrapper's dtype=object inversion table included), the errors
the reference wrote, and the logbook.  The reference tree cannot travel, the trace can: the CUDA library must give
the same integers call for call (legacy symbol), generation for generation (population API), and when only the
changed individuals of a resident population are re-scored (delta API) -- hence the same logbook.
"""
import ctypes
import os

import numpy as np
import pytest

from common import EMU_LIB, GOLDEN, dense_matrix
from scbonita_b200 import _lib, simulator

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def backend(request):
    if request.param == "emu":
        return "emu", request.getfixturevalue("emu_lib")
    return "cuda", request.getfixturevalue("cuda_lib")


@pytest.fixture(scope="module")
def trace():
    z = np.load(os.path.join(GOLDEN, "ga_trace_hsa00010.npz"))
    t = {k: z[k] for k in z.files}
    t["n_cells"], t["ind_len"] = int(t["n_cells"]), int(t["ind_len"])
    t["bin_rows"] = np.unpackbits(t["bin_rows"], axis=1)[:, :t["n_cells"]].astype(np.int32)
    return t


def _simulator(t, lib_path):
    model = simulator.ModelTables(t["and_len"], t["parse"], t["and_nodes"][0], t["and_inv"][0], t["node_positions"],
                                  t["ind_len"])
    sim = simulator.Simulator(dense_matrix(t["bin_rows"], t["node_positions"]), model, lib_path=lib_path)
    if lib_path == EMU_LIB:
        sim.set_option(_lib.OPT_MAX_CTAS, 3)
    return sim


def test_trace_is_a_real_ga_run(trace):
    """Sanity of the fixture itself: mu+lambda shape, reference sums, and the logbook rows the GA printed."""
    t = trace
    assert t["individuals"].shape[0] == t["nevals"].sum() == t["fitness"].shape[0]
    assert np.array_equal(t["cellwise"].astype(np.int64).sum(axis=1), t["fitness"])
    first = t["fitness"][:t["nevals"][0]].astype(np.float64)
    gen, nevals, avg, std, lo, hi = t["logbook"][0]
    assert (gen, nevals) == (0, t["nevals"][0])
    assert (avg, std, lo, hi) == (first.mean(), first.std(), first.min(), first.max())
    assert len(np.unique(t["individuals"], axis=0)) > t["nevals"][0]      # the GA really varied the rule sets


def test_population_api_reproduces_the_logbook(backend, trace):
    """One population call per generation: fitness and cellwise errors identical to what the reference C returned to the
    unmodified GA, so every statistic the GA logs (ruleMaker.py:958-962) is identical."""
    t = trace
    with _simulator(t, backend[1]) as sim:
        lo = 0
        seen = []
        for g, n in enumerate(t["nevals"]):
            sl = slice(lo, lo + int(n))
            fit = sim.evaluate_population(t["individuals"][sl], t["knockouts"], t["knockins"], t["and_nodes"][sl],
                                          t["and_inv"][sl], mode=simulator.MODE_SUM)
            cw = sim.evaluate_population(t["individuals"][sl], t["knockouts"], t["knockins"], t["and_nodes"][sl],
                                         t["and_inv"][sl], mode=simulator.MODE_CELLWISE)
            assert np.array_equal(fit, t["fitness"][sl]), g
            assert np.array_equal(cw[:, :t["n_cells"]], t["cellwise"][sl]), g
            seen.extend(int(f) for f in fit)
            lo += int(n)
        first = np.array(seen[:t["nevals"][0]], np.float64)
        assert tuple(t["logbook"][0][2:]) == (first.mean(), first.std(), first.min(), first.max())
        # mu+lambda keeps the best: the logged minimum of every generation is the best value evaluated so far
        best = np.minimum.accumulate([min(seen[a:b]) for a, b in zip(np.cumsum(t["nevals"]) - t["nevals"], np.cumsum(t["nevals"]))])
        assert np.array_equal(best.astype(np.float64), t["logbook"][:, 4])


def test_legacy_symbol_replays_every_call(backend, trace):
    """The exact 17 x c_void_p call (ruleMaker.py:1200-1219) with the exact bytes the reference C saw."""
    t = trace
    lib = _lib.load(backend[1])
    stride = 128
    lib.scb_set_legacy_geometry(20000, stride)
    try:
        bm = np.zeros((int(t["node_positions"].max()) + 1, stride), np.int32)
        bm[:, :t["n_cells"]] = dense_matrix(t["bin_rows"], t["node_positions"])
        N = len(t["and_len"])
        V = ctypes.c_void_p
        fn = lib.scSyncBool
        fn.restype, fn.argtypes = None, None
        vals = np.zeros((100, N), np.int32)
        al, parse, ko, ki, pos = (np.ascontiguousarray(t[k], np.int32) for k in ("and_len", "parse", "knockouts", "knockins", "node_positions"))
        step = 1 if backend[0] == "cuda" else 6
        for c in range(0, t["individuals"].shape[0], step):
            ind = np.ascontiguousarray(t["individuals"][c], np.int32)
            an = np.ascontiguousarray(t["and_nodes"][c], np.int32)
            ai = np.ascontiguousarray(t["and_inv"][c], np.int32)
            errors = np.full(t["n_cells"] + 2, -5, np.int32)
            fn(V(vals.ctypes.data), V(ind.ctypes.data), V(len(ind)), V(N), V(al.ctypes.data), V(parse.ctypes.data),
               V(an.ctypes.data), V(ai.ctypes.data), V(100), V(ko.ctypes.data), V(ki.ctypes.data), V(t["n_cells"]),
               V(bm.ctypes.data), V(pos.ctypes.data), V(errors.ctypes.data), V(0), V(0))
            assert np.array_equal(errors[:t["n_cells"]], t["cellwise"][c]), c
            assert (errors[t["n_cells"]:] == -5).all()
    finally:
        lib.scb_set_legacy_geometry(20000, 15000)


def test_resident_population_scores_only_the_offspring(backend, trace):
    """the population stays on the device, every generation only the offspring rows are replaced
    (scb_pop_update) and only they are simulated; the parents keep the sums of the run that scored them."""
    t = trace
    n0 = int(t["nevals"][0])
    with _simulator(t, backend[1]) as sim:
        # slots [0, n0): the first generation; slots [n0, n0 + lambda): room for one generation of offspring
        lam = int(t["nevals"][1:].max())
        first = slice(0, n0)
        filler = np.arange(lam) % n0
        sim.upload(np.concatenate([t["individuals"][first], t["individuals"][filler]]), t["knockouts"], t["knockins"],
                   np.concatenate([t["and_nodes"][first], t["and_nodes"][filler]]),
                   np.concatenate([t["and_inv"][first], t["and_inv"][filler]]))
        sim.run()
        fit = sim.fetch()
        assert np.array_equal(fit[:n0], t["fitness"][first])
        assert np.array_equal(fit[n0:], t["fitness"][filler])
        lo = n0
        for g in range(1, len(t["nevals"])):
            n = int(t["nevals"][g])
            sl = slice(lo, lo + n)
            slots = n0 + np.arange(n)
            before = sim.fetch().copy()
            sim.update(slots, t["individuals"][sl], t["and_nodes"][sl], t["and_inv"][sl])
            sim.run()
            fit = sim.fetch()
            assert np.array_equal(fit[slots], t["fitness"][sl]), g
            keep = np.setdiff1d(np.arange(n0 + lam), slots)
            assert np.array_equal(fit[keep], before[keep]), g                  # untouched slots keep their sums
            lo += n
        # an empty update: nothing to score, everything stays
        sim.update(np.zeros(0, np.int32), np.zeros((0, t["ind_len"]), np.int32), np.zeros((0,) + t["and_nodes"].shape[1:], np.int32),
                   np.zeros((0,) + t["and_inv"].shape[1:], np.int32))
        sim.run()
        assert np.array_equal(sim.fetch(), fit)
        # a plain run afterwards scores everything again
        sim.run()
        assert np.array_equal(sim.fetch(), fit)

"""Host-side mirror of the reference's simulator boundary, on top of the C-ABI.

What it mirrors (file:line under /root/reference/src/scBONITA/):

* ``ruleMaker.__evaluateByNode(individual, KOlist, KIlist, cFunction, localSearch, importanceScores)``
  (ruleMaker.py:1077-1221) -> :meth:`Simulator.evaluate_by_node` -- same argument meaning and
  return shapes (``[sum(errors)]`` for the GA, ``errors[:nodeNum]`` for local search).
* the two list comprehensions that score a GA population one C call at a time
  (ruleMaker.py:982-985, 1030-1033) -> :meth:`Simulator.evaluate_population` (one launch).
* ``ruleMaker.__checkNodePossibilities`` (ruleMaker.py:1235-1266) ->
  :meth:`Simulator.check_node_possibilities` -- same 4-tuple.
* the distance / decider stage of ``singleCell.assignAttractors`` (singleCell.py:1211-1238) ->
  :meth:`Simulator.hamming_argmin`.

All arithmetic happens in the CUDA library; this module only marshals numpy arrays.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import MODE_CELLWISE, MODE_NODEWISE, MODE_SUM, ScbError, Stats  # noqa: F401


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


class ModelTables:
    """The int32 tables of one network model exactly as ``__evaluateByNode`` builds them
    (ruleMaker.py:1089-1129): andLenList, individualParse, andNodes[N][7][3],
    andNodeInvertList[N][7][3], nodePositions; ``ind_len`` = len(individual) (``model.size``)."""

    def __init__(self, and_len_list, individual_parse, and_nodes, and_node_invert_list, node_positions, ind_len):
        self.and_len_list = _i32(and_len_list)
        self.n = int(self.and_len_list.shape[0])
        self.individual_parse = _i32(individual_parse)
        if self.individual_parse.shape[0] < self.n:
            raise ValueError("individualParse needs at least nodeNum entries")
        self.and_nodes = _i32(and_nodes).reshape(self.n, 7, 3)
        self.and_node_invert_list = _i32(and_node_invert_list).reshape(self.n, 7, 3)
        self.node_positions = _i32(node_positions)
        self.ind_len = int(ind_len)


class Simulator:
    """One (binarized dataset, network) pair on one GPU: packs the matrix once (K1) and scores
    rule sets against it.  ``bin_mat`` is the dense gene-major 0/1 matrix (int32 as the
    reference passes it, or uint8); only rows ``node_positions`` are used."""

    def __init__(self, bin_mat, model, n_cells=None, device=0, lib_path=None):
        self.lib = _lib.load(lib_path)
        self.model = model
        if hasattr(bin_mat, "indptr") and hasattr(bin_mat, "indices"):
            # scipy.sparse CSR, the form the reference keeps binMat in: no dense detour
            csr = bin_mat.tocsr() if hasattr(bin_mat, "tocsr") else bin_mat
            self.n_cells = int(csr.shape[1] if n_cells is None else n_cells)
            if int(model.node_positions.max(initial=0)) >= csr.shape[0]:
                raise ValueError("nodePositions exceed the rows of bin_mat")
            indptr, indices = _i32(csr.indptr), _i32(csr.indices)
            data = np.ascontiguousarray(csr.data, dtype=np.float64)
            self._ctx = ctypes.c_void_p()
            rc = self.lib.scb_ctx_create_csr(ctypes.byref(self._ctx), int(device), model.n, self.n_cells, _ptr(indptr),
                                             _ptr(indices), _ptr(data), int(csr.shape[0]), _ptr(model.node_positions))
            _lib.check(self.lib, rc)
            self.last_stats = None
            return
        bm = np.asarray(bin_mat)
        if bm.dtype == np.uint8 or bm.dtype == np.bool_:
            bm = np.ascontiguousarray(bm, dtype=np.uint8)
            elem = 1
        else:
            bm = _i32(bm)
            elem = 4
        if bm.ndim != 2:
            raise ValueError("bin_mat must be 2-D (genes x cells)")
        self.n_cells = int(bm.shape[1] if n_cells is None else n_cells)
        if int(model.node_positions.max(initial=0)) >= bm.shape[0]:
            raise ValueError("nodePositions exceed the rows of bin_mat")
        self._ctx = ctypes.c_void_p()
        rc = self.lib.scb_ctx_create(ctypes.byref(self._ctx), int(device), model.n, self.n_cells, _ptr(bm), elem,
                                     int(bm.shape[1]), _ptr(model.node_positions))
        _lib.check(self.lib, rc)
        self.last_stats = None

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.scb_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_option(self, key, value):
        _lib.check(self.lib, self.lib.scb_ctx_set_option(self._ctx, int(key), int(value)))

    def info(self):
        v = [ctypes.c_int() for _ in range(6)]
        _lib.check(self.lib, self.lib.scb_ctx_info(self._ctx, *[ctypes.byref(x) for x in v]))
        keys = ("nodes", "cells", "words_per_lane", "warps", "ctas", "smem_bytes")
        return dict(zip(keys, (x.value for x in v)))

    # -- population scoring -------------------------------------------------------------------
    def _knock(self, ko_list, ki_list):
        """ruleMaker.py:1089-1095: both arrays are filled from *KOlist*, used as indices."""
        n = self.model.n
        knockouts = np.zeros(n, dtype=np.int32)
        knockins = np.zeros(n, dtype=np.int32)
        for k in ko_list:
            knockouts[k] = 1
        for k in ko_list:
            knockins[k] = 1
        return knockouts, knockins

    def _tables(self, individuals, and_nodes, and_inv):
        m = self.model
        inds = _i32(individuals)
        if inds.ndim == 1:
            inds = inds.reshape(1, -1)
        if inds.shape[1] != m.ind_len:
            raise ValueError("individuals must have %d bits" % m.ind_len)
        an = m.and_nodes if and_nodes is None else _i32(and_nodes)
        ai = m.and_node_invert_list if and_inv is None else _i32(and_inv)
        per_ind = int(an.ndim == 4)
        if per_ind and (an.shape[0] != inds.shape[0] or ai.shape != an.shape):
            raise ValueError("per-individual tables must be [P][N][7][3]")
        return inds, an, ai, per_ind

    def upload(self, individuals, knockouts=None, knockins=None, and_nodes=None, and_inv=None):
        """H2D of one population (async on the context's stream)."""
        m = self.model
        inds, an, ai, per_ind = self._tables(individuals, and_nodes, and_inv)
        ko = _i32(knockouts if knockouts is not None else np.zeros(m.n))
        ki = _i32(knockins if knockins is not None else np.zeros(m.n))
        self._keep = (inds, an, ai, ko, ki)          # keep the host arrays alive until the copy is done
        self._P = inds.shape[0]
        rc = self.lib.scb_pop_upload(self._ctx, inds.shape[0], _ptr(inds), m.ind_len, _ptr(m.and_len_list),
                                     _ptr(m.individual_parse), _ptr(an), _ptr(ai), per_ind, _ptr(ko), _ptr(ki))
        _lib.check(self.lib, rc)

    def upload_compact(self, individuals, upstream, upstream_inv, knockouts=None, knockins=None):
        """H2D of one population in the compact form: ``upstream`` / ``upstream_inv`` are
        [P][N][3] (every individual its own wiring, what ``__update_upstream`` changes, ruleMaker.py:300-335) or
        one shared [N][3]; see :func:`upstream_from_tables`.  Then ``run`` / ``fetch`` as usual."""
        m = self.model
        inds = _i32(individuals)
        if inds.ndim == 1:
            inds = inds.reshape(1, -1)
        if inds.shape[1] != m.ind_len:
            raise ValueError("individuals must have %d bits" % m.ind_len)
        up, ui = _i32(upstream), _i32(upstream_inv)
        per_ind = int(up.ndim == 3)
        if up.shape != ui.shape or up.shape[-2:] != (m.n, 3) or (per_ind and up.shape[0] != inds.shape[0]):
            raise ValueError("upstream tables must be [P][N][3] or [N][3]")
        ko = _i32(knockouts if knockouts is not None else np.zeros(m.n))
        ki = _i32(knockins if knockins is not None else np.zeros(m.n))
        self._keep = (inds, up, ui, ko, ki)
        self._P = inds.shape[0]
        rc = self.lib.scb_pop_upload_compact(self._ctx, inds.shape[0], _ptr(inds), m.ind_len, _ptr(m.and_len_list),
                                             _ptr(m.individual_parse), _ptr(up), _ptr(ui), per_ind, _ptr(ko), _ptr(ki))
        _lib.check(self.lib, rc)

    def update(self, slots, individuals, and_nodes=None, and_inv=None):
        """Resident population (scb_pop_update): individuals ``slots`` of the uploaded population are replaced in place
        and the next ``run()`` scores only them; ``fetch()`` keeps returning all P sums."""
        sl = _i32(slots).reshape(-1)
        n = int(sl.shape[0])
        inds = _i32(individuals).reshape(n, -1) if n else np.zeros((0, 1), np.int32)
        an = _i32(and_nodes) if and_nodes is not None else None
        ai = _i32(and_inv) if and_inv is not None else None
        rc = self.lib.scb_pop_update(self._ctx, n, _ptr(sl), _ptr(inds), _ptr(an) if an is not None else None,
                                     _ptr(ai) if ai is not None else None)
        _lib.check(self.lib, rc)

    def run(self, mode=MODE_SUM):
        _lib.check(self.lib, self.lib.scb_pop_run(self._ctx, int(mode)))

    def _out_array(self, mode, P):
        if mode == MODE_SUM:
            return np.zeros(P, dtype=np.int64)
        if mode == MODE_NODEWISE:
            return np.zeros((P, self.model.n), dtype=np.int32)
        return np.zeros((P, self.n_cells), dtype=np.int32)

    def fetch(self, mode=MODE_SUM, out=None):
        out = self._out_array(mode, self._P) if out is None else out
        st = Stats()
        rc = self.lib.scb_pop_fetch(self._ctx, int(mode), _ptr(out), ctypes.byref(st))
        self.last_stats = st.as_dict()
        _lib.check(self.lib, rc)
        return out

    # -- cost-ordered work list, peer-memory fitness exchange --------------------------------
    def costs(self, tile_max=False):
        """uint32[P]: cost of every individual of the last run in state-row units (scb_pop_costs); with ``tile_max``
        also the cost of its costliest tile."""
        out = np.zeros(self._P, dtype=np.uint32)
        if not tile_max:
            _lib.check(self.lib, self.lib.scb_pop_costs(self._ctx, _ptr(out), None))
            return out
        mx = np.zeros(self._P, dtype=np.uint32)
        _lib.check(self.lib, self.lib.scb_pop_costs(self._ctx, _ptr(out), _ptr(mx)))
        return out, mx

    def cost_order(self):
        """The order to hand to :meth:`set_order` after a run: individuals with the costliest tiles first."""
        total, mx = self.costs(tile_max=True)
        return np.lexsort((-total.astype(np.int64), -mx.astype(np.int64))).astype(np.int32)

    def set_order(self, order):
        """Hand the next runs their individuals costliest first (``np.argsort(-costs)``); None: index order."""
        if order is None:
            _lib.check(self.lib, self.lib.scb_pop_set_order(self._ctx, None, 0))
            return
        o = _i32(order)
        _lib.check(self.lib, self.lib.scb_pop_set_order(self._ctx, _ptr(o), int(o.shape[0])))

    def gather_init(self, world, rank, width):
        """Allocates this rank's gather buffers; returns the 128 bytes of CUDA IPC handles the other ranks need."""
        h = np.zeros(128, dtype=np.uint8)
        _lib.check(self.lib, self.lib.scb_gather_init(self._ctx, int(world), int(rank), int(width), _ptr(h)))
        self._gather = (int(world), int(width))
        return h

    def gather_connect(self, all_handles):
        h = np.ascontiguousarray(all_handles, dtype=np.uint8)
        _lib.check(self.lib, self.lib.scb_gather_connect(self._ctx, _ptr(h)))

    def gather_result(self, out=None):
        world, width = self._gather
        out = np.zeros(world * width, dtype=np.int64) if out is None else out
        _lib.check(self.lib, self.lib.scb_gather_result(self._ctx, _ptr(out)))
        return out

    def gather_close(self):
        _lib.check(self.lib, self.lib.scb_gather_close(self._ctx))

    def sync(self):
        _lib.check(self.lib, self.lib.scb_ctx_sync(self._ctx))

    def event_record(self, slot):
        _lib.check(self.lib, self.lib.scb_ctx_event_record(self._ctx, int(slot)))

    def event_elapsed_ms(self, slot0, slot1):
        ms = ctypes.c_float()
        _lib.check(self.lib, self.lib.scb_ctx_event_elapsed_ms(self._ctx, int(slot0), int(slot1), ctypes.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        _lib.check(self.lib, self.lib.scb_ctx_flush_l2(self._ctx))

    def result_device_pointer(self, mode=MODE_SUM):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(self.lib, self.lib.scb_pop_result_dev(self._ctx, int(mode), ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def stream_pointer(self):
        p = ctypes.c_void_p()
        _lib.check(self.lib, self.lib.scb_ctx_stream(self._ctx, ctypes.byref(p)))
        return p.value or 0

    def programs(self, p=0, k=4):
        """Diagnostics: the compiled full program of individual ``p`` of the last run and its reduced program ``k``
        (0..4; 4 = constants propagated to the fixed point).  Returns a dict with ``descs`` / ``descs_k``
        (uint32[N][4]), ``counts`` / ``counts_k`` (nodes of arity 3, 2, 1, 0) and ``from_k`` (first step the reduced
        program may compute, 0 if it was not emitted)."""
        n = self.model.n
        d1, dk = np.zeros((n, 4), np.uint32), np.zeros((n, 4), np.uint32)
        c1, ck, fk = np.zeros(4, np.int32), np.zeros(4, np.int32), np.zeros(1, np.int32)
        _lib.check(self.lib, self.lib.scb_debug_programs(self._ctx, int(p), int(k), _ptr(d1), _ptr(c1), _ptr(dk), _ptr(ck), _ptr(fk)))
        return {"descs": d1, "counts": c1, "descs_k": dk, "counts_k": ck, "from_k": int(fk[0])}

    def microbench(self, kind):
        """Measured peak of one on-chip resource with every SM busy: "lop3" / "popc" in thread-level ops per
        second, "lds" in bytes per second of conflict-free 128-bit shared-memory loads (roofline denominators)."""
        v = ctypes.c_double()
        _lib.check(self.lib, self.lib.scb_microbench(self._ctx, {"lop3": 0, "popc": 1, "lds": 2}[kind], ctypes.byref(v)))
        return v.value

    def evaluate_population(self, individuals, knockouts=None, knockins=None, and_nodes=None, and_inv=None,
                            mode=MODE_SUM):
        """Score P individuals in one launch.  Returns int64[P] (SUM), int32[P][N] (NODEWISE) or
        int32[P][C] (CELLWISE).  ``and_nodes`` / ``and_inv`` may be [P][N][7][3] (every GA individual owns
        a model copy, ruleMaker.py:650) or None for the context's shared model."""
        m = self.model
        inds, an, ai, per_ind = self._tables(individuals, and_nodes, and_inv)
        ko = _i32(knockouts if knockouts is not None else np.zeros(m.n))
        ki = _i32(knockins if knockins is not None else np.zeros(m.n))
        out = self._out_array(mode, inds.shape[0])
        st = Stats()
        # ONE library call (upload + run + fetch under the context's lock): safe from a ThreadPool, which is how the
        # reference drives its parallel local search (singleCell.py:490-492); upload()/run()/fetch() used separately
        # are for a single driving thread (bench.py's timed loops)
        rc = self.lib.scb_eval_population(self._ctx, inds.shape[0], _ptr(inds), m.ind_len, _ptr(m.and_len_list),
                                          _ptr(m.individual_parse), _ptr(an), _ptr(ai), per_ind, _ptr(ko), _ptr(ki),
                                          int(mode), _ptr(out), ctypes.byref(st))
        self.last_stats = st.as_dict()
        self._P = inds.shape[0]
        _lib.check(self.lib, rc)
        return out

    # -- the reference's per-individual entry point ------------------------------------------
    def evaluate_by_node(self, individual, KOlist, KIlist, cFunction=None, localSearch=False,
                         importanceScores=False, and_nodes=None, and_inv=None):
        """ruleMaker.__evaluateByNode (ruleMaker.py:1077-1221).  ``cFunction`` is accepted and
        ignored (the reference passes the ctypes symbol).  Returns ``errors[:nodeNum]`` as a
        list when ``localSearch`` else ``[sum(errors)]``."""
        knockouts, knockins = self._knock(KOlist, KIlist)
        if importanceScores and not localSearch:
            # ruleMaker.py:1172-1194: cFunction is importanceScore, the result a one-element float list
            m = self.model
            ind = _i32(individual)
            score = np.zeros(1, dtype=np.float64)
            rc = self.lib.scb_importance_score(self._ctx, _ptr(ind), m.ind_len, _ptr(m.and_len_list),
                                               _ptr(m.individual_parse), _ptr(m.and_nodes),
                                               _ptr(m.and_node_invert_list), _ptr(knockouts), _ptr(knockins), _ptr(score))
            _lib.check(self.lib, rc)
            return score.tolist()
        if localSearch:
            out = self.evaluate_population(individual, knockouts, knockins, and_nodes, and_inv, MODE_NODEWISE)
            return out[0].tolist()
        out = self.evaluate_population(individual, knockouts, knockins, and_nodes, and_inv, MODE_SUM)
        return [int(out[0])]

    def importance_scores(self, individual, KOlists, KIlists=None, strict=True):
        """The loop over nodes of ruleMaker.__calcImportance (ruleMaker.py:1285-1300) in one launch: one
        ``__evaluateByNode(individual, KOlist, KIlist, importanceScore, importanceScores=True)`` result per
        (KOlist, KIlist) pair, each bit-identical to the single call.  ``KIlists`` defaults to ``KOlists``
        (what ``__calcImportance`` passes: ``[node], [node]``).  Returns float64[len(KOlists)].  Where the reference
        divides by zero (a cell without an attractor window) the score is NaN: an error when ``strict``."""
        m = self.model
        KIlists = KOlists if KIlists is None else KIlists
        if len(KOlists) != len(KIlists) or len(KOlists) == 0:
            raise ValueError("KOlists and KIlists must be non-empty and of equal length")
        kos = np.zeros((len(KOlists), m.n), dtype=np.int32)
        kis = np.zeros((len(KOlists), m.n), dtype=np.int32)
        for k, (kol, kil) in enumerate(zip(KOlists, KIlists)):
            kos[k], kis[k] = self._knock(kol, kil)
        ind = _i32(individual)
        scores = np.zeros(len(KOlists), dtype=np.float64)
        rc = self.lib.scb_importance_scores(self._ctx, _ptr(ind), m.ind_len, _ptr(m.and_len_list),
                                            _ptr(m.individual_parse), _ptr(m.and_nodes), _ptr(m.and_node_invert_list),
                                            len(KOlists), _ptr(kos), _ptr(kis), _ptr(scores))
        if not (rc == _lib.ERR_INVALID and not strict and np.isnan(scores).any()):
            _lib.check(self.lib, rc)
        return scores

    def check_node_possibilities(self, node, indy, KOlist, KIlist, scSyncBoolC=None):
        """ruleMaker.__checkNodePossibilities (ruleMaker.py:1235-1266): all 2^k - 1 rule bit-strings of
        ``node`` in one launch.  Returns ``(truth, equivs, minny, indErrors)``."""
        m = self.model
        start = int(m.individual_parse[node])
        end = m.ind_len if node == m.n - 1 else int(m.individual_parse[node + 1])
        truth = list(indy[start:end])
        equivs = [truth]
        if end - start == 0:
            return truth, equivs, equivs, 0.0
        knockouts, knockins = self._knock(KOlist, KIlist)
        ind = _i32(indy)
        nv = 2 ** (end - start) - 1
        errs = np.zeros(nv, dtype=np.int32)
        got = ctypes.c_int()
        rc = self.lib.scb_eval_local_search(self._ctx, _ptr(ind), m.ind_len, _ptr(m.and_len_list),
                                            _ptr(m.individual_parse), _ptr(m.and_nodes),
                                            _ptr(m.and_node_invert_list), _ptr(knockouts), _ptr(knockins),
                                            int(node), _ptr(errs), nv, ctypes.byref(got))
        _lib.check(self.lib, rc)
        ind_errors = errs.tolist()
        ind_options = [[(i >> j) & 1 for j in range(end - start)] for i in range(1, nv + 1)]   # __bitList
        minny = min(ind_errors)
        equivs = [ind_options[i] for i in range(nv) if ind_errors[i] <= minny]
        return equivs[0], equivs, minny, ind_errors

    def local_search(self, indy, KOlist, KIlist):
        """The loop over nodes of singleCell.__scoreNodes (singleCell.py:478-510) in one launch: returns, per node,
        the same 4-tuple ``__checkNodePossibilities`` returns."""
        m = self.model
        knockouts, knockins = self._knock(KOlist, KIlist)
        ind = _i32(indy)
        ends = [m.ind_len if i == m.n - 1 else int(m.individual_parse[i + 1]) for i in range(m.n)]
        ks = [ends[i] - int(m.individual_parse[i]) for i in range(m.n)]
        total = sum((1 << k) - 1 for k in ks if k > 0)
        errs = np.zeros(max(total, 1), dtype=np.int32)
        offs = np.zeros(m.n + 1, dtype=np.int32)
        rc = self.lib.scb_eval_local_search_all(self._ctx, _ptr(ind), m.ind_len, _ptr(m.and_len_list),
                                                _ptr(m.individual_parse), _ptr(m.and_nodes),
                                                _ptr(m.and_node_invert_list), _ptr(knockouts), _ptr(knockins),
                                                _ptr(errs), int(errs.shape[0]), _ptr(offs))
        _lib.check(self.lib, rc)
        out = []
        for i in range(m.n):
            start, k = int(m.individual_parse[i]), ks[i]
            truth = list(indy[start:start + k])
            if k == 0:
                out.append((truth, [truth], [truth], 0.0))
                continue
            ind_errors = errs[offs[i]:offs[i + 1]].tolist()
            options = [[(v >> j) & 1 for j in range(k)] for v in range(1, (1 << k))]
            minny = min(ind_errors)
            equivs = [options[v] for v in range(len(options)) if ind_errors[v] <= minny]
            out.append((equivs[0], equivs, minny, ind_errors))
        return out

    # -- attractor companion ------------------------------------------------------------------
    def _attractors_packed(self, individual, knockouts, knockins, semantics, max_states):
        m = self.model
        sem = {"c": 0, "python": 1}[semantics]
        ind = _i32(individual)
        ko = _i32(knockouts if knockouts is not None else np.zeros(m.n))
        ki = _i32(knockins if knockins is not None else np.zeros(m.n))
        nw = (m.n + 31) // 32
        res = np.zeros((self.n_cells, 2), dtype=np.int32)
        nst = np.zeros(self.n_cells, dtype=np.int32)
        packed = np.zeros((self.n_cells, max(max_states, 1), nw), dtype=np.uint32)
        rc = self.lib.scb_attractors(self._ctx, _ptr(ind), m.ind_len, _ptr(m.and_len_list), _ptr(m.individual_parse),
                                     _ptr(m.and_nodes), _ptr(m.and_node_invert_list), _ptr(ko), _ptr(ki), sem,
                                     int(max_states), _ptr(res), _ptr(nst), _ptr(packed))
        _lib.check(self.lib, rc)
        return res, nst, packed

    def attractors(self, individual, knockouts=None, knockins=None, semantics="python", max_states=10):
        """Per-cell attractor window and its states (singleCell.py:1162-1192 runs this one cell at a time in
        pure Python).  ``semantics``: "python" = singleCell.__cluster with simSteps 10 (rows res[0]..res[1]
        inclusive; [-1,-1] yields the last row), "c" = simulator.c cluster with simSteps 100 (rows
        [res[0], res[1])).  Returns (res[C][2], n_states[C], states[C][max_states][N] as 0/1 uint8)."""
        res, nst, packed = self._attractors_packed(individual, knockouts, knockins, semantics, max_states)
        bits = np.unpackbits(packed.view(np.uint8), axis=-1, bitorder="little")[..., :self.model.n]
        return res, nst, bits

    def attractor_set(self, individual, knockouts=None, knockins=None, remove_zero=True, packed=False):
        """The unique attractor states of all cells, as assignAttractors collects them
        (singleCell.py:1189-1207), in first-seen order (the reference's order is Python set order).  Deduplicated on
        the device (scb_attractor_states); ``packed=True`` returns the bit rows uint32[A][ceil(N/32)] as
        :meth:`hamming_argmin` also accepts them, else 0/1 int32[A][N]."""
        m = self.model
        n = m.n
        ind = _i32(individual)
        ko = _i32(knockouts if knockouts is not None else np.zeros(n))
        ki = _i32(knockins if knockins is not None else np.zeros(n))
        nw = (n + 31) // 32
        cap = 1 << 14
        while True:
            uniq = np.zeros((cap, nw), dtype=np.uint32)
            cnt = ctypes.c_int(0)
            rc = self.lib.scb_attractor_states(self._ctx, _ptr(ind), m.ind_len, _ptr(m.and_len_list), _ptr(m.individual_parse),
                                               _ptr(m.and_nodes), _ptr(m.and_node_invert_list), _ptr(ko), _ptr(ki), 1, 10,
                                               _ptr(uniq), cap, ctypes.byref(cnt))
            _lib.check(self.lib, rc)
            if cnt.value <= cap:
                break
            cap = cnt.value
        uniq = uniq[:cnt.value]
        if remove_zero:
            uniq = uniq[uniq.any(axis=1)]
        if packed:
            return np.ascontiguousarray(uniq)
        if uniq.shape[0] == 0:
            return np.zeros((0, n), dtype=np.int32)
        out = np.unpackbits(np.ascontiguousarray(uniq).view(np.uint8), axis=-1, bitorder="little")[:, :n].astype(np.int32)
        return np.ascontiguousarray(out).reshape(-1, n)

    def hamming_argmin(self, attractors, want_dist=True):
        """singleCell.py:1211-1238: returns (dist[C][A] or None, decider[C], deciderDistance[C]).  ``attractors``:
        0/1 values [A][nodeNum], or (dtype uint32) bit rows [A][ceil(nodeNum/32)] as ``attractor_set(packed=True)``
        returns them."""
        raw = np.asarray(attractors)
        nw = (self.model.n + 31) // 32
        is_packed = raw.dtype == np.uint32 and raw.ndim == 2 and raw.shape[1] == nw and nw != self.model.n
        at = np.ascontiguousarray(raw) if is_packed else _i32(raw)
        if at.ndim != 2 or at.shape[1] != (nw if is_packed else self.model.n):
            raise ValueError("attractors must be [A][nodeNum]")
        A = at.shape[0]
        dist = np.zeros((self.n_cells, A), dtype=np.int32) if want_dist else None
        dec = np.zeros(self.n_cells, dtype=np.int32)
        dd = np.zeros(self.n_cells, dtype=np.int32)
        fn = self.lib.scb_hamming_argmin_packed if is_packed else self.lib.scb_hamming_argmin
        rc = fn(self._ctx, _ptr(at), A, _ptr(dist), _ptr(dec), _ptr(dd))
        _lib.check(self.lib, rc)
        return dist, dec, dd


def upstream_from_tables(and_len_list, and_nodes, and_inv):
    """The compact form of reference-format term tables: each node's upstream nodes and inversion flags,
    int32[...][N][3] padded with -1.  The first AND-term of a node lists all of its inputs
    (ruleMaker.py:188-200: itertools.product starts with the full set), so the triple is that term cut to
    log2(andLen + 1) entries."""
    al = np.asarray(and_len_list, dtype=np.int64)
    an, ai = np.asarray(and_nodes, dtype=np.int32), np.asarray(and_inv, dtype=np.int32)
    k = np.round(np.log2(al + 1)).astype(np.int64)                       # 0, 1, 3, 7 terms -> 0, 1, 2, 3 inputs
    if not np.array_equal((1 << k) - 1, al):
        raise ValueError("andLenList entries must be 0, 1, 3 or 7 for the compact form")
    keep = np.arange(3)[None, :] < k[:, None]                            # [N][3]
    up = np.where(keep, an[..., 0, :], -1).astype(np.int32)
    ui = np.where(keep, ai[..., 0, :], -1).astype(np.int32)
    return np.ascontiguousarray(up), np.ascontiguousarray(ui)


class PinnedArray:
    """A numpy array in pinned host memory handed out by the library (scb_host_alloc), so that uploads are
    truly asynchronous.  Keep the object alive while ``.array`` is in use; ``close()`` frees it."""

    def __init__(self, shape, dtype, lib_path=None):
        self.lib = _lib.load(lib_path)
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        self._ptr = ctypes.c_void_p()
        _lib.check(self.lib, self.lib.scb_host_alloc(ctypes.byref(self._ptr), max(count * dtype.itemsize, 1)))
        buf = (ctypes.c_char * max(count * dtype.itemsize, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def close(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            self.lib.scb_host_free(self._ptr)
            self._ptr = None

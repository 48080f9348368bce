"""oracle -- CPU checkers for the scBONITA fitness hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (scbonita_b200/) never does.

Two things live here:

* ``liboracle`` -- our CPU restatements (oracle/scb_oracle.c scalar, scb_oracle_packed.c
  bit-packed), built by ``oracle/Makefile`` into ``oracle/_build/``.
* ``RefSimulator`` -- the UNMODIFIED reference ``simulator.c`` compiled into ``oracle/_ref/``
  (``make -C oracle ref``; needs /root/reference, so only in the build container -- the
  prebuilt .so files travel to the GPU box) and driven through ctypes with the reference's own
  calling convention (every argument a ``c_void_p``, ruleMaker.py:1117-1219).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
BUILD_DIR = os.path.join(_HERE, "_build")
REF_SRC = "/root/reference/src/scBONITA/simulator.c"

REF_NODE = 20000            # simulator.c:7
REF_STEP = 100              # simulator.c:6
REF_CELL_VARIANTS = {15000: "simulator.so", 65536: "simulator_cell65536.so",
                     131072: "simulator_cell131072.so"}


def build(ref=True, quiet=True):
    """Build the restatements and, when the reference tree is present, oracle/_ref."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", _HERE, "all"], stdout=out)
    if ref and os.path.exists(REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=out)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Diag(ctypes.Structure):
    _fields_ = [("not_found_cells", ctypes.c_int64), ("undefined_leading", ctypes.c_int64),
                ("empty_rule_evals", ctypes.c_int64)]


_lib = None


def liboracle():
    global _lib
    if _lib is None:
        path = os.path.join(BUILD_DIR, "libscb_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _lib = ctypes.CDLL(path)
        for name in ("scbo_sync_bool", "scbo_cluster", "scbo_cluster_literal", "scbo_py_cluster", "scbo_hamming_argmin",
                     "scbo_importance_score", "scbo_packed_population"):
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


class Model:
    """The int32 tables the hot path consumes (what ruleMaker.__evaluateByNode marshals,
    ruleMaker.py:1089-1129), for one network.  and_nodes / and_inv: [N][7][3]."""

    def __init__(self, and_len, parse, and_nodes, and_inv, node_positions, ind_len,
                 knockouts=None, knockins=None):
        self.and_len = _i32(and_len)
        self.n = int(self.and_len.shape[0])
        self.parse = _i32(parse)
        self.and_nodes = _i32(and_nodes).reshape(self.n, 7, 3)
        self.and_inv = _i32(and_inv).reshape(self.n, 7, 3)
        self.node_positions = _i32(node_positions)
        self.ind_len = int(ind_len)
        self.knockouts = _i32(knockouts if knockouts is not None else np.zeros(self.n))
        self.knockins = _i32(knockins if knockins is not None else np.zeros(self.n))


def sync_bool(model, individual, bin_mat, n_cells, local_search, sim_steps=100,
              and_nodes=None, and_inv=None, want_found=False):
    """Scalar restatement of scSyncBool. bin_mat: int32 2-D (rows >= max(node_positions)+1).
    Returns (errors, diag[, found])."""
    lib = liboracle()
    bm = _i32(bin_mat)
    ind = _i32(individual)
    an = _i32(and_nodes if and_nodes is not None else model.and_nodes)
    ai = _i32(and_inv if and_inv is not None else model.and_inv)
    errors = np.zeros(model.n if local_search else n_cells, dtype=np.int32)
    found = np.zeros(n_cells, dtype=np.int8) if want_found else None
    d = Diag()
    rc = lib.scbo_sync_bool(_p(ind), ctypes.c_int(len(ind)), ctypes.c_int(model.n),
                            _p(model.and_len), _p(model.parse), _p(an), _p(ai),
                            ctypes.c_int(sim_steps), _p(model.knockouts), _p(model.knockins),
                            ctypes.c_int(n_cells), _p(bm), ctypes.c_int64(bm.shape[1]),
                            _p(model.node_positions), _p(errors), ctypes.c_int(int(local_search)),
                            None, ctypes.c_int64(0), _p(found), ctypes.byref(d))
    if rc:
        raise ValueError("scbo_sync_bool: contract violation (rc=%d)" % rc)
    return (errors, d, found) if want_found else (errors, d)


def packed_population(model, individuals, bin_mat, n_cells, and_nodes=None, and_inv=None,
                      want_nodewise=True, want_cellwise=False, want_mu_lambda=False, threads=0):
    """Bit-packed restatement over a population.  and_nodes/and_inv: None (model's, shared),
    [N][7][3] shared, or [P][N][7][3] per individual."""
    lib = liboracle()
    bm = _i32(bin_mat)
    inds = _i32(individuals).reshape(-1, model.ind_len)
    P = inds.shape[0]
    an = _i32(and_nodes if and_nodes is not None else model.and_nodes)
    ai = _i32(and_inv if and_inv is not None else model.and_inv)
    per_ind = int(an.ndim == 4)
    fitness = np.zeros(P, dtype=np.int64)
    nodewise = np.zeros((P, model.n), dtype=np.int32) if want_nodewise else None
    cellwise = np.zeros((P, n_cells), dtype=np.int32) if want_cellwise else None
    not_found = np.zeros(P, dtype=np.int64)
    mulam = np.zeros((P, n_cells, 2), dtype=np.int32) if want_mu_lambda else None
    rc = lib.scbo_packed_population(ctypes.c_int(P), _p(inds), ctypes.c_int(model.ind_len),
                                    ctypes.c_int(model.n), _p(model.and_len), _p(model.parse),
                                    _p(an), _p(ai), ctypes.c_int(per_ind),
                                    _p(model.knockouts), _p(model.knockins),
                                    ctypes.c_int(n_cells), _p(bm), ctypes.c_int64(bm.shape[1]),
                                    _p(model.node_positions), _p(fitness), _p(nodewise),
                                    _p(cellwise), _p(not_found), _p(mulam), ctypes.c_int(threads))
    if rc:
        raise ValueError("scbo_packed_population: contract violation (rc=%d)" % rc)
    return dict(fitness=fitness, nodewise=nodewise, cellwise=cellwise, not_found=not_found,
                mu_lambda=mulam)


def py_cluster(model, individual, bin_mat, n_cells, sim_steps=10):
    """singleCell.__cluster restated: returns (vals[C][steps][N], res[C][2])."""
    lib = liboracle()
    bm = _i32(bin_mat)
    ind = _i32(individual)
    vals = np.zeros((n_cells, sim_steps, model.n), dtype=np.int32)
    res = np.zeros((n_cells, 2), dtype=np.int32)
    d = Diag()
    rc = lib.scbo_py_cluster(_p(ind), ctypes.c_int(len(ind)), ctypes.c_int(model.n),
                             _p(model.and_len), _p(model.parse), _p(model.and_nodes),
                             _p(model.and_inv), ctypes.c_int(sim_steps), _p(model.knockouts),
                             _p(model.knockins), ctypes.c_int(n_cells), _p(bm),
                             ctypes.c_int64(bm.shape[1]), _p(model.node_positions),
                             _p(vals), _p(res), ctypes.byref(d))
    if rc:
        raise ValueError("scbo_py_cluster: rc=%d" % rc)
    return vals, res, d


def c_cluster(model, individual, bin_mat, cell, sim_steps=100):
    """C `cluster` restated for one cell: (traj[100][N], res[2])."""
    lib = liboracle()
    bm = _i32(bin_mat)
    ind = _i32(individual)
    traj = np.zeros((REF_STEP, model.n), dtype=np.int32)
    res = np.zeros(2, dtype=np.int32)
    d = Diag()
    rc = lib.scbo_cluster(ctypes.c_int(cell), _p(ind), ctypes.c_int(len(ind)), ctypes.c_int(model.n),
                          _p(model.and_len), _p(model.parse), _p(model.and_nodes), _p(model.and_inv),
                          ctypes.c_int(sim_steps), _p(model.knockouts), _p(model.knockins), _p(bm),
                          ctypes.c_int64(bm.shape[1]), _p(model.node_positions), _p(traj), _p(res),
                          ctypes.byref(d))
    if rc:
        raise ValueError("scbo_cluster: rc=%d" % rc)
    return traj, res


def c_cluster_literal(model, individual, bin_mat, cell, sim_data=None, sim_stride=None, sim_steps=100):
    """C `cluster` restated LITERALLY on a scratch like the caller's ``simData`` (incl. the knock-out branch's
    column slip, simulator.c:518).  ``sim_data``: int32[100][sim_stride] scratch (zeros when None; the reference's
    Python passes zeros).  Returns (sim_data after the call, res[2])."""
    lib = liboracle()
    bm = _i32(bin_mat)
    ind = _i32(individual)
    if sim_stride is None:
        sim_stride = max(int(model.node_positions.max()) + 1, model.n)
    sd = np.zeros((REF_STEP, sim_stride), dtype=np.int32) if sim_data is None else np.array(sim_data, dtype=np.int32, order="C")
    assert sd.shape == (REF_STEP, sim_stride)
    res = np.zeros(2, dtype=np.int32)
    d = Diag()
    rc = lib.scbo_cluster_literal(ctypes.c_int(cell), _p(ind), ctypes.c_int(len(ind)), ctypes.c_int(model.n),
                                  _p(model.and_len), _p(model.parse), _p(model.and_nodes), _p(model.and_inv),
                                  ctypes.c_int(sim_steps), _p(model.knockouts), _p(model.knockins), _p(bm),
                                  ctypes.c_int64(bm.shape[1]), _p(model.node_positions), _p(sd),
                                  ctypes.c_int64(sim_stride), _p(res), ctypes.byref(d))
    if rc:
        raise ValueError("scbo_cluster_literal: rc=%d" % rc)
    return sd, res


def hamming_argmin(attractors, bin_mat, node_positions, n_cells, want_dist=True):
    lib = liboracle()
    at = _i32(attractors)
    A, n = at.shape
    bm = _i32(bin_mat)
    pos = _i32(node_positions)
    dist = np.zeros((n_cells, A), dtype=np.int32) if want_dist else None
    dec = np.zeros(n_cells, dtype=np.int32)
    dd = np.zeros(n_cells, dtype=np.int32)
    lib.scbo_hamming_argmin(_p(at), ctypes.c_int(A), ctypes.c_int(n), ctypes.c_int(n_cells), _p(bm),
                            ctypes.c_int64(bm.shape[1]), _p(pos), _p(dist), _p(dec), _p(dd))
    return dist, dec, dd


def importance_score(model, individual, bin_mat, n_cells, sim_steps=100):
    lib = liboracle()
    bm = _i32(bin_mat)
    ind = _i32(individual)
    score = np.zeros(1, dtype=np.float64)
    rc = lib.scbo_importance_score(_p(ind), ctypes.c_int(len(ind)), ctypes.c_int(model.n),
                                   _p(model.and_len), _p(model.parse), _p(model.and_nodes),
                                   _p(model.and_inv), ctypes.c_int(sim_steps), _p(model.knockouts),
                                   _p(model.knockins), ctypes.c_int(n_cells), _p(bm),
                                   ctypes.c_int64(bm.shape[1]), _p(model.node_positions), _p(score))
    if rc:
        raise ValueError("scbo_importance_score: rc=%d (-2: attractor not found -> reference SIGFPE)" % rc)
    return float(score[0])


# ------------------------------------------------------------------------------------------
# the UNMODIFIED reference, through its own ctypes convention
# ------------------------------------------------------------------------------------------
def ref_available(cells=15000):
    return any(os.path.exists(os.path.join(REF_DIR, f)) for c, f in REF_CELL_VARIANTS.items()
               if c >= cells)


class RefSimulator:
    """Loads oracle/_ref/simulator*.so (smallest CELL variant >= n_cells) and calls the
    reference symbols exactly like ruleMaker.__evaluateByNode does (ruleMaker.py:1117-1219):
    all 17 arguments as c_void_p, ints smuggled as pointer values, binMat dense int32 with the
    compiled-in row stride CELL, simData a zeroed int32[100][NODE]."""

    def __init__(self, n_cells):
        cell = min(c for c in REF_CELL_VARIANTS if c >= n_cells)
        path = os.path.join(REF_DIR, REF_CELL_VARIANTS[cell])
        if not os.path.exists(path):
            if os.path.exists(REF_SRC):
                build(ref=True)
            else:
                raise FileNotFoundError(path + " (build it in the container: make -C oracle ref)")
        self.cell = cell
        self.lib = ctypes.cdll.LoadLibrary(path)

    def dense(self, bin_rows, node_positions):
        """Place rows (len(node_positions) x C, row k = gene node_positions[k]) into a
        [max(pos)+1][CELL] int32 matrix -- the only rows the C code reads."""
        pos = np.asarray(node_positions)
        C = bin_rows.shape[1]
        bm = np.zeros((int(pos.max()) + 1, self.cell), dtype=np.int32)
        bm[pos, :C] = bin_rows
        return bm

    def sync_bool(self, model, individual, bin_mat, n_cells, local_search, sim_steps=100,
                  and_nodes=None, and_inv=None):
        """bin_mat: int32 2-D with row stride == self.cell (use .dense())."""
        assert bin_mat.dtype == np.int32 and bin_mat.shape[1] == self.cell and bin_mat.flags.c_contiguous
        ind = _i32(individual)
        an = _i32(and_nodes if and_nodes is not None else model.and_nodes)
        ai = _i32(and_inv if and_inv is not None else model.and_inv)
        vals = np.zeros((REF_STEP, REF_NODE), dtype=np.int32)
        errors = np.zeros(max(model.n, 1) if local_search else max(n_cells, 1), dtype=np.int32)
        V = ctypes.c_void_p
        self.lib.scSyncBool(V(vals.ctypes.data), V(ind.ctypes.data), V(len(ind)), V(model.n),
                            V(model.and_len.ctypes.data), V(model.parse.ctypes.data),
                            V(an.ctypes.data), V(ai.ctypes.data), V(sim_steps),
                            V(model.knockouts.ctypes.data), V(model.knockins.ctypes.data),
                            V(n_cells), V(bin_mat.ctypes.data), V(model.node_positions.ctypes.data),
                            V(errors.ctypes.data), V(int(local_search)), V(0))
        return errors

    def importance_score(self, model, individual, bin_mat, n_cells, sim_steps=100):
        ind = _i32(individual)
        vals = np.zeros((REF_STEP, REF_NODE), dtype=np.int32)
        score = np.zeros(1, dtype=np.float64)
        V = ctypes.c_void_p
        self.lib.importanceScore(V(vals.ctypes.data), V(ind.ctypes.data), V(len(ind)), V(model.n),
                                 V(model.and_len.ctypes.data), V(model.parse.ctypes.data),
                                 V(model.and_nodes.ctypes.data), V(model.and_inv.ctypes.data),
                                 V(sim_steps), V(model.knockouts.ctypes.data),
                                 V(model.knockins.ctypes.data), V(n_cells), V(bin_mat.ctypes.data),
                                 V(model.node_positions.ctypes.data), V(score.ctypes.data))
        return float(score[0])

    def cluster(self, model, individual, bin_mat, cell, sim_steps=100, sim_data=None):
        """sim_data: optional int32[100][k] initial scratch contents (copied into the first k columns)."""
        ind = _i32(individual)
        vals = np.zeros((REF_STEP, REF_NODE), dtype=np.int32)
        if sim_data is not None:
            vals[:, :np.shape(sim_data)[1]] = sim_data
        res = np.full(2, 2, dtype=np.int32)
        V = ctypes.c_void_p
        self.lib.cluster(V(vals.ctypes.data), V(res.ctypes.data), V(cell), V(ind.ctypes.data),
                         V(len(ind)), V(model.n), V(model.and_len.ctypes.data),
                         V(model.parse.ctypes.data), V(model.and_nodes.ctypes.data),
                         V(model.and_inv.ctypes.data), V(sim_steps), V(model.knockouts.ctypes.data),
                         V(model.knockins.ctypes.data), V(bin_mat.ctypes.data),
                         V(model.node_positions.ctypes.data))
        return vals, res

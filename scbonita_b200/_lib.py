"""ctypes binding of the C-ABI in include/scbonita_b200.h.

The product library is the in-tree ``scbonita_b200/libscbonita_b200.so`` (built by
``__graft_entry__.build()`` / ``make -C scbonita_b200/csrc``).  There is no CPU fallback: if
the library is missing, or no CUDA device is usable, calls raise.  ``load(path)`` accepts an
explicit path only so that the CPU test-suite can point the *same* binding at the host-thread
emulator build of the kernels under tests/emu/ (test infrastructure, never used by default).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCB_B200_LIB points the binding at another build of the same library (A/B measurements of kernel variants)
DEFAULT_PATH = os.environ.get("SCB_B200_LIB") or os.path.join(_HERE, "libscbonita_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = 0, -1, -2, -3, -4
MODE_SUM, MODE_NODEWISE, MODE_CELLWISE = 0, 1, 2
OPT_EARLY_EXIT, OPT_SNAPSHOTS, OPT_WARPS, OPT_MAX_CTAS, OPT_WORDS, OPT_CHECK_EVERY, OPT_PERIOD_SEARCH = 1, 2, 3, 4, 5, 6, 7
OPT_SNAP_LO, OPT_SNAP_HI, OPT_SNAP_MIN_GAP, OPT_TMEM, OPT_FUSE, OPT_TMA, OPT_STEADY, OPT_CHECK_FROM, OPT_GRAPH = 8, 9, 10, 11, 12, 13, 14, 15, 16
OPT_GENERIC = 17
OPT_RESOLVE_GROUP = 18
OPT_AUTO_ORDER = 19

# every symbol include/scbonita_b200.h declares (tests check that the library exports them all)
SYMBOLS = (
    "scb_last_error", "scb_version", "scb_device_count",
    "scb_ctx_create", "scb_ctx_create_csr", "scb_ctx_destroy", "scb_ctx_set_option", "scb_ctx_info",
    "scb_eval_population", "scb_pop_upload", "scb_pop_upload_compact", "scb_pop_run", "scb_pop_fetch", "scb_pop_result_dev",
    "scb_pop_costs", "scb_pop_set_order", "scb_pop_update", "scb_gather_init", "scb_gather_connect", "scb_gather_result", "scb_gather_close",
    "scb_ctx_stream", "scb_ctx_sync", "scb_ctx_event_record", "scb_ctx_event_elapsed_ms",
    "scb_ctx_flush_l2", "scb_host_alloc", "scb_host_free",
    "scb_eval_local_search", "scb_eval_local_search_all", "scb_hamming_argmin", "scb_attractors", "scb_attractor_states", "scb_hamming_argmin_packed", "scb_importance_score",
    "scb_importance_scores", "scb_microbench", "scb_debug_programs",
    "scb_set_legacy_geometry", "scSyncBool", "importanceScore", "cluster",
)


class Stats(ctypes.Structure):
    _fields_ = [("not_found_cells", ctypes.c_int64), ("undefined_leading", ctypes.c_int64),
                ("empty_rule_nodes", ctypes.c_int64), ("unsupported_nodes", ctypes.c_int64),
                ("bad_index_nodes", ctypes.c_int64), ("term_overflow_nodes", ctypes.c_int64),
                ("resolver_chunks", ctypes.c_int64), ("steps_executed", ctypes.c_int64),
                ("tiles", ctypes.c_int64), ("kernel_ms", ctypes.c_float), ("sim_ms", ctypes.c_float),
                ("cyc_prologue", ctypes.c_int64), ("cyc_steps", ctypes.c_int64), ("cyc_epilogue", ctypes.c_int64),
                ("state_rows_moved", ctypes.c_int64), ("clear_ms", ctypes.c_float), ("compile_ms", ctypes.c_float),
                ("resolve_ms", ctypes.c_float), ("fill_ms", ctypes.c_float), ("gather_ms", ctypes.c_float),
                ("reserved_", ctypes.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ScbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("scbonita_b200 error %d: %s" % (code, message))
        self.code = code


_libs = {}


def load(path=None):
    """Load (once per path) and type the library.  Raises if it is not there."""
    path = os.path.abspath(path or DEFAULT_PATH)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C scbonita_b200/csrc); scbonita_b200 has no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    vp, i32p = ctypes.c_void_p, ctypes.c_void_p
    lib.scb_last_error.restype = ctypes.c_char_p
    lib.scb_last_error.argtypes = []
    lib.scb_version.restype = ctypes.c_int
    lib.scb_device_count.restype = ctypes.c_int
    lib.scb_ctx_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int,
                                   ctypes.c_int64, i32p]
    lib.scb_ctx_create_csr.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p, vp,
                                       ctypes.c_int, i32p]
    lib.scb_ctx_create_csr.restype = ctypes.c_int
    lib.scb_ctx_destroy.argtypes = [vp]
    lib.scb_ctx_set_option.argtypes = [vp, ctypes.c_int, ctypes.c_int64]
    lib.scb_ctx_info.argtypes = [vp] + [ctypes.POINTER(ctypes.c_int)] * 6
    tables = [i32p, ctypes.c_int, i32p, i32p, i32p, i32p, ctypes.c_int, i32p, i32p]
    lib.scb_eval_population.argtypes = [vp, ctypes.c_int] + tables + [ctypes.c_int, vp, ctypes.POINTER(Stats)]
    lib.scb_pop_upload.argtypes = [vp, ctypes.c_int] + tables
    lib.scb_pop_upload_compact.argtypes = [vp, ctypes.c_int] + tables
    lib.scb_pop_upload_compact.restype = ctypes.c_int
    lib.scb_pop_run.argtypes = [vp, ctypes.c_int]
    lib.scb_pop_fetch.argtypes = [vp, ctypes.c_int, vp, ctypes.POINTER(Stats)]
    lib.scb_pop_result_dev.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int64)]
    lib.scb_pop_costs.argtypes = [vp, vp, vp]
    lib.scb_pop_update.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
    lib.scb_pop_set_order.argtypes = [vp, vp, ctypes.c_int]
    lib.scb_gather_init.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
    lib.scb_gather_connect.argtypes = [vp, vp]
    lib.scb_gather_result.argtypes = [vp, vp]
    lib.scb_gather_close.argtypes = [vp]
    for name in ("scb_pop_costs", "scb_pop_set_order", "scb_pop_update", "scb_gather_init", "scb_gather_connect", "scb_gather_result", "scb_gather_close"):
        getattr(lib, name).restype = ctypes.c_int
    lib.scb_ctx_stream.argtypes = [vp, ctypes.POINTER(vp)]
    lib.scb_ctx_sync.argtypes = [vp]
    lib.scb_ctx_event_record.argtypes = [vp, ctypes.c_int]
    lib.scb_ctx_event_elapsed_ms.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    lib.scb_ctx_flush_l2.argtypes = [vp]
    lib.scb_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_int64]
    lib.scb_host_free.argtypes = [vp]
    lib.scb_eval_local_search.argtypes = [vp, i32p, ctypes.c_int, i32p, i32p, i32p, i32p, i32p, i32p, ctypes.c_int,
                                          i32p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.scb_eval_local_search_all.argtypes = [vp, i32p, ctypes.c_int, i32p, i32p, i32p, i32p, i32p, i32p, i32p,
                                              ctypes.c_int, i32p]
    lib.scb_eval_local_search_all.restype = ctypes.c_int
    lib.scb_hamming_argmin.argtypes = [vp, i32p, ctypes.c_int, i32p, i32p, i32p]
    lib.scb_hamming_argmin_packed.argtypes = [vp, vp, ctypes.c_int, i32p, i32p, i32p]
    lib.scb_attractor_states.argtypes = [vp, i32p, ctypes.c_int, i32p, i32p, i32p, i32p, i32p, i32p, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.scb_attractors.argtypes = [vp, i32p, ctypes.c_int, i32p, i32p, i32p, i32p, i32p, i32p, ctypes.c_int,
                                   ctypes.c_int, i32p, i32p, vp]
    lib.scb_importance_score.argtypes = [vp, i32p, ctypes.c_int, i32p, i32p, i32p, i32p, i32p, i32p, vp]
    lib.scb_importance_score.restype = ctypes.c_int
    lib.scb_importance_scores.argtypes = [vp, i32p, ctypes.c_int, i32p, i32p, i32p, i32p, ctypes.c_int, i32p, i32p, vp]
    lib.scb_importance_scores.restype = ctypes.c_int
    lib.scb_debug_programs.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]
    lib.scb_debug_programs.restype = ctypes.c_int
    lib.scb_microbench.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.scb_microbench.restype = ctypes.c_int
    lib.scb_set_legacy_geometry.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.scb_set_legacy_geometry.restype = None
    for name in ("scb_ctx_create", "scb_ctx_create_csr", "scb_ctx_destroy", "scb_ctx_set_option", "scb_ctx_info", "scb_eval_population",
                 "scb_pop_upload", "scb_pop_run", "scb_pop_fetch", "scb_pop_result_dev", "scb_ctx_stream",
                 "scb_ctx_sync", "scb_ctx_event_record", "scb_ctx_event_elapsed_ms", "scb_ctx_flush_l2", "scb_host_alloc",
                 "scb_host_free", "scb_eval_local_search",
                 "scb_hamming_argmin", "scb_attractors", "scb_attractor_states", "scb_hamming_argmin_packed"):
        getattr(lib, name).restype = ctypes.c_int
    _libs[path] = lib
    return lib


def check(lib, rc):
    if rc != OK:
        raise ScbError(rc, (lib.scb_last_error() or b"").decode("utf-8", "replace"))

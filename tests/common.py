"""Shared helpers of the test-suite: seeded problems, adversarial networks, oracle wrappers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle  # noqa: E402  (test infrastructure: the checker)
from scbonita_b200 import simulator, synth  # noqa: E402

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libscbonita_emu.so")
CUDA_LIB = os.path.join(ROOT, "scbonita_b200", "libscbonita_b200.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def dense_matrix(bin_rows, node_positions):
    """[genes][cells] int32 with gene row node_positions[i] = bin_rows[i]."""
    pos = np.asarray(node_positions)
    bm = np.zeros((int(pos.max()) + 1, bin_rows.shape[1]), dtype=np.int32)
    bm[pos] = bin_rows
    return bm


def oracle_model(net, node_positions, ko, ki):
    return oracle.Model(net.and_len, net.parse, net.and_nodes, net.and_inv, node_positions, net.ind_len, ko, ki)


def tables(net, node_positions):
    return simulator.ModelTables(net.and_len, net.parse, net.and_nodes, net.and_inv, node_positions, net.ind_len)


class Problem:
    """A synthetic problem plus the oracle's answers for it."""

    def __init__(self, net, bin_rows, node_positions, individuals, pop_an=None, pop_ai=None, ko=None, ki=None):
        self.net = net
        self.n = net.n
        self.bin_rows = bin_rows
        self.node_positions = np.asarray(node_positions, dtype=np.int32)
        self.n_cells = bin_rows.shape[1]
        self.individuals = np.ascontiguousarray(individuals, dtype=np.int32).reshape(-1, net.ind_len)
        self.pop_an, self.pop_ai = pop_an, pop_ai
        self.ko = np.zeros(net.n, np.int32) if ko is None else np.asarray(ko, np.int32)
        self.ki = np.zeros(net.n, np.int32) if ki is None else np.asarray(ki, np.int32)
        self.bm = dense_matrix(bin_rows, self.node_positions)
        self._oracle = None

    @classmethod
    def synthetic(cls, n, cells, pop, knock=True, **kw):
        pr = synth.make_problem(n, cells, pop, knock=knock, **kw)
        return cls(pr.net, pr.bin_rows, pr.node_positions, pr.individuals, pr.pop_and_nodes, pr.pop_and_inv,
                   pr.knockouts, pr.knockins)

    def oracle(self):
        if self._oracle is None:
            m = oracle_model(self.net, self.node_positions, self.ko, self.ki)
            self._oracle = oracle.packed_population(m, self.individuals, self.bm, self.n_cells,
                                                    and_nodes=self.pop_an, and_inv=self.pop_ai, want_cellwise=True)
        return self._oracle

    def simulator(self, lib_path, **options):
        from scbonita_b200 import _lib
        sim = simulator.Simulator(self.bm, tables(self.net, self.node_positions), lib_path=lib_path)
        if lib_path == EMU_LIB:
            sim.set_option(_lib.OPT_MAX_CTAS, options.pop("max_ctas", 3))
        for k, v in options.items():
            sim.set_option(getattr(_lib, "OPT_" + k.upper()), v)
        return sim

    def check(self, sim, modes=(0, 1, 2)):
        """Run all output shapes and compare with the oracle bit for bit."""
        ref = self.oracle()
        keys = {0: "fitness", 1: "nodewise", 2: "cellwise"}
        for mode in modes:
            out = sim.evaluate_population(self.individuals, self.ko, self.ki, self.pop_an, self.pop_ai, mode=mode)
            want = ref[keys[mode]]
            assert out.shape == want.shape, (mode, out.shape, want.shape)
            if not np.array_equal(out, want):
                bad = np.argwhere(out != want)
                raise AssertionError("mode %s differs at %d places, first %s: got %s want %s (stats %s)" % (
                    keys[mode], len(bad), bad[0], out[tuple(bad[0])], want[tuple(bad[0])], sim.last_stats))
            assert sim.last_stats["not_found_cells"] == int(ref["not_found"].sum()), sim.last_stats
        return sim.last_stats


# ---- adversarial networks -----------------------------------------------------------------
class _Net:
    pass


def gated_ring_network(ring_lens, chain_len=0, inverter=True):
    """Rings of copy nodes with one optional inverter each, every ring node ANDed with a gate node g
    (g <- g), plus a decaying chain (c_0 <- 0 via knock-out, c_j <- c_{j-1}).

    For a cell with g = 1 a ring of n nodes rotates with period 2n (n without inverter) after a transient
    set by the chain; with g = 0 everything dies to a fixed point.  So "attractor found" (transient +
    period <= 99) varies cell by cell -- the reference's fill-forward of not-found cells and
    the exact resolver get exercised.  Returns (net, individual, knockouts)."""
    nodes = []          # (kind, args)
    g = 0
    nodes.append(("copy", g))
    for n in ring_lens:
        base = len(nodes)
        for j in range(n):
            pred = base + (j - 1) % n
            nodes.append(("and2", pred, g, 1 if (inverter and j == 0) else 0))
    chain_base = len(nodes)
    for j in range(chain_len):
        nodes.append(("ko",) if j == 0 else ("copy", chain_base + j - 1))
    N = len(nodes)
    net = _Net()
    net.n = N
    net.and_len = np.zeros(N, np.int32)
    net.parse = np.zeros(N + 1, np.int32)
    net.and_nodes = np.zeros((N, 7, 3), np.int32)
    net.and_inv = np.zeros((N, 7, 3), np.int32)
    bits, ko = [], np.zeros(N, np.int32)
    for i, nd in enumerate(nodes):
        net.parse[i] = len(bits)
        if nd[0] == "copy":
            net.and_len[i] = 1
            net.and_nodes[i, 0] = [nd[1], -1, -1]
            net.and_inv[i, 0] = [0, -1, -1]
            bits.append(1)
        elif nd[0] == "ko":
            net.and_len[i] = 1
            net.and_nodes[i, 0] = [i, -1, -1]
            net.and_inv[i, 0] = [0, -1, -1]
            ko[i] = 1
            bits.append(1)
        else:
            _, pred, gate, inv = nd
            net.and_len[i] = 3                      # terms [pred&g, pred, g]; only the first selected
            net.and_nodes[i, 0] = [pred, gate, -1]
            net.and_inv[i, 0] = [inv, 0, -1]
            net.and_nodes[i, 1] = [pred, -1, -1]
            net.and_inv[i, 1] = [inv, -1, -1]
            net.and_nodes[i, 2] = [gate, -1, -1]
            net.and_inv[i, 2] = [0, -1, -1]
            bits.extend([1, 0, 0])
    net.parse[N] = len(bits)
    net.ind_len = len(bits)
    return net, np.asarray(bits, np.int32), ko


def ring_problem(ring_lens, chain_len, cells, seed=7, gate_density=0.5, inverter=True, lead_found=True):
    net, ind, ko = gated_ring_network(ring_lens, chain_len, inverter)
    rng = np.random.default_rng(seed)
    rows = (rng.random((net.n, cells)) < 0.5).astype(np.uint8)
    rows[0] = (rng.random(cells) < gate_density).astype(np.uint8)
    if lead_found:
        rows[0, 0] = 0        # the first cell reaches a fixed point, so no fill-forward value is undefined
    pos = rng.permutation(net.n).astype(np.int32)
    return Problem(net, rows, pos, ind.reshape(1, -1), ko=ko, ki=np.zeros(net.n, np.int32))

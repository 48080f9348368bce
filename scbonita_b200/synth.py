"""Seeded synthetic inputs of the shapes BASELINE.json names.

Everything here produces the *reference-format* int32 tables that
``ruleMaker.__evaluateByNode`` marshals for the C simulator (ruleMaker.py:1089-1129), so the
same arrays can be fed to the reference build, the CPU restatement and the CUDA path.

* network  (seed 1234 + N): in-degree per node ~ {0: .106, 1: .128, 2: .362, 3: .404}
  (the term-count histogram of the packaged hsa00010 model), distinct predecessors uniform
  over nodes (self edges allowed, ``removeSelfEdges=False``), inhibitory flag per edge
  Bernoulli(0.10); AND-terms enumerated in ``itertools.product`` order
  ``[abc, ab, ac, a, bc, b, c]`` (ruleMaker.py:188-200) and padded with -1 / ``[0,0,0]``
  exactly like ``__updateCpointers`` (ruleMaker.py:337-360).
* cells    (seed 5678): ``binMat[g][c]`` ~ Bernoulli(0.21) i.i.d.; ``nodePositions`` a random
  permutation of the gene rows.
* population (seed 91011): bits per multi-term node uniform, at most five and at least one
  set (``__genBits``, ruleMaker.py:627-650); nodes with spare candidate predecessors are
  rewired per individual with probability 0.25 (``__mutFlipBitAdapt``, ruleMaker.py:588-611),
  which makes the connectivity tables per-individual.
"""
from dataclasses import dataclass, field
from itertools import product

import numpy as np

IN_DEGREE_P = {0: 0.106, 1: 0.128, 2: 0.362, 3: 0.404}
MAX_TERMS = 7
MAX_IN = 3


def enumerate_terms(preds, flags):
    """AND-term enumeration of ruleMaker.py:188-200 for up to three predecessors.
    Returns (terms, term_flags): lists of lists, product order, empty term dropped."""
    terms, tflags = [], []
    pairs = [((p, f), None) for p, f in zip(preds, flags)]
    for combo in product(*pairs):
        chosen = [x for x in combo if x is not None]
        if chosen:
            terms.append([p for p, _ in chosen])
            tflags.append([f for _, f in chosen])
    return terms, tflags


def pad_tables(terms, tflags):
    """__updateCpointers (ruleMaker.py:337-360) for one node: [7][3] int32 pair."""
    an = np.zeros((MAX_TERMS, MAX_IN), dtype=np.int32)
    ai = np.zeros((MAX_TERMS, MAX_IN), dtype=np.int32)
    for a, (t, f) in enumerate(zip(terms, tflags)):
        an[a] = t + [-1] * (MAX_IN - len(t))
        ai[a] = [int(x) for x in f] + [-1] * (MAX_IN - len(f))
    return an, ai


@dataclass
class SynthNetwork:
    n: int
    and_len: np.ndarray          # [N]
    parse: np.ndarray            # [N+1] (the reference passes this whole list, ruleMaker.py:1103)
    ind_len: int
    and_nodes: np.ndarray        # [N][7][3]
    and_inv: np.ndarray          # [N][7][3]
    preds: list = field(default_factory=list)        # used predecessors per node
    candidates: list = field(default_factory=list)   # all candidate predecessors per node
    cand_flags: list = field(default_factory=list)   # inhibitory flag per candidate edge


def make_network(n, seed=None, inhibit_p=0.10, spare_p=0.30):
    rng = np.random.default_rng(1234 + n if seed is None else seed)
    ks = rng.choice(list(IN_DEGREE_P), size=n, p=list(IN_DEGREE_P.values()))
    ks = np.minimum(ks, n)
    and_len = np.zeros(n, dtype=np.int32)
    parse = np.zeros(n + 1, dtype=np.int32)
    and_nodes = np.zeros((n, MAX_TERMS, MAX_IN), dtype=np.int32)
    and_inv = np.zeros((n, MAX_TERMS, MAX_IN), dtype=np.int32)
    preds, cands, cflags = [], [], []
    counter = 0
    for i in range(n):
        k = int(ks[i])
        extra = 0
        if k == 3 and n > 6 and rng.random() < spare_p:
            extra = int(rng.integers(1, 4))
        cand = [int(x) for x in rng.choice(n, size=k + extra, replace=False)] if k else []
        flags = [bool(rng.random() < inhibit_p) for _ in cand]
        terms, tflags = enumerate_terms(cand[:k], flags[:k])
        and_nodes[i], and_inv[i] = pad_tables(terms, tflags)
        and_len[i] = len(terms)
        parse[i] = counter
        counter += len(terms)
        preds.append(cand[:k]); cands.append(cand); cflags.append(flags)
    parse[n] = counter
    return SynthNetwork(n, and_len, parse, counter, and_nodes, and_inv, preds, cands, cflags)


def make_cells(n_genes, n_cells, density=0.21, seed=5678):
    """Returns (bin_rows uint8 [n_genes][n_cells], node_positions int32 [n_genes])."""
    rng = np.random.default_rng(seed)
    bin_rows = (rng.random((n_genes, n_cells)) < density).astype(np.uint8)
    pos = rng.permutation(n_genes).astype(np.int32)
    return bin_rows, pos


def make_population(net, pop, seed=91011, rewire_p=0.25):
    """Returns (individuals int32 [P][L], and_nodes [P][N][7][3], and_inv [P][N][7][3]).
    The per-individual tables equal the network's except on rewired nodes."""
    rng = np.random.default_rng(seed)
    L = max(net.ind_len, 1)
    inds = rng.integers(0, 2, size=(pop, L)).astype(np.int32)
    if net.ind_len == 0:
        inds = np.zeros((pop, 0), dtype=np.int32)
    an = np.broadcast_to(net.and_nodes, (pop,) + net.and_nodes.shape).copy()
    ai = np.broadcast_to(net.and_inv, (pop,) + net.and_inv.shape).copy()
    for i in range(net.n):
        s, e = int(net.parse[i]), int(net.parse[i + 1])
        w = e - s
        if w == 1:
            inds[:, s] = 1
        elif w > 1:
            for p in range(pop):
                bits = inds[p, s:e]
                while bits.sum() > 5:                      # ruleMaker.py:641-644
                    bits[int(rng.random() * w)] = 0
                if bits.sum() == 0:                        # ruleMaker.py:645-647
                    bits[int(rng.random() * w)] = 1
        if len(net.candidates[i]) > 3:
            for p in np.nonzero(rng.random(pop) < rewire_p)[0]:
                pick = rng.choice(len(net.candidates[i]), size=3, replace=False)
                terms, tflags = enumerate_terms([net.candidates[i][j] for j in pick],
                                                [net.cand_flags[i][j] for j in pick])
                an[p, i], ai[p, i] = pad_tables(terms, tflags)
    return inds, an, ai


def default_knock(n):
    """Q4: the shipped GA / local search always passes KO = KI = [1, 0, 0, ...]
    (singleCell.py:440-443 + ruleMaker.py:1092-1095)."""
    k = np.zeros(n, dtype=np.int32)
    if n:
        k[0] = 1
    return k, k.copy()


@dataclass
class SynthProblem:
    net: SynthNetwork
    bin_rows: np.ndarray         # uint8 [N][C], gene-major
    node_positions: np.ndarray   # int32 [N]
    individuals: np.ndarray      # int32 [P][L]
    pop_and_nodes: np.ndarray    # int32 [P][N][7][3]
    pop_and_inv: np.ndarray      # int32 [P][N][7][3]
    knockouts: np.ndarray
    knockins: np.ndarray

    @property
    def n_cells(self):
        return self.bin_rows.shape[1]


CONFIGS = {
    # name: (nodes, cells, population)   -- BASELINE.json configs[1], [2], [4]
    "C2": (60, 10_000, 100),
    "C3": (200, 50_000, 500),
    "C5": (300, 100_000, 1),
}


def make_problem(n, cells, pop, knock=True, net_seed=None, cell_seed=5678, pop_seed=91011):
    net = make_network(n, seed=net_seed)
    bin_rows, pos = make_cells(n, cells, seed=cell_seed)
    inds, an, ai = make_population(net, pop, seed=pop_seed)
    ko, ki = default_knock(n) if knock else (np.zeros(n, np.int32), np.zeros(n, np.int32))
    return SynthProblem(net, bin_rows, pos, inds, an, ai, ko, ki)


def make_gated_rings(n, cells, pop, ring_lens=(45, 50, 57), chain_len=24, seed=7, gate_density=0.5):
    """An adversarial network for the early-exit machinery: rings of copy nodes with one inverter each, every ring
    node ANDed with a gate node g (g <- g), a decaying chain fed by a knocked node, free copy nodes up to `n`.
    A cell with g = 1 puts ring r on a cycle of period 2 * ring_lens[r] after a transient set by the chain; with
    g = 0 everything dies to a fixed point.  Periods beyond 35 defeat the in-flight repeat proofs, periods beyond
    99 make the cell "attractor not found" (SURVEY Q7): every tile runs all 99 steps and goes to the exact
    resolver.  Returns a SynthProblem whose `pop` individuals are identical (one rule set per node)."""
    nodes = [("copy", 0)]
    for ln in ring_lens:
        base = len(nodes)
        for j in range(ln):
            nodes.append(("and2", base + (j - 1) % ln, 0, 1 if j == 0 else 0))
    chain_base = len(nodes)
    for j in range(chain_len):
        nodes.append(("ko",) if j == 0 else ("copy", chain_base + j - 1))
    while len(nodes) < n:
        nodes.append(("copy", len(nodes) - 1))
    if len(nodes) != n:
        raise ValueError("rings + chain need %d nodes, n = %d" % (len(nodes), n))
    and_len = np.zeros(n, np.int32)
    parse = np.zeros(n + 1, np.int32)
    and_nodes = np.zeros((n, MAX_TERMS, MAX_IN), np.int32)
    and_inv = np.zeros((n, MAX_TERMS, MAX_IN), np.int32)
    bits, ko = [], np.zeros(n, np.int32)
    for i, nd in enumerate(nodes):
        parse[i] = len(bits)
        if nd[0] in ("copy", "ko"):
            and_len[i] = 1
            and_nodes[i, 0] = [nd[1] if nd[0] == "copy" else i, -1, -1]
            and_inv[i, 0] = [0, -1, -1]
            ko[i] = 1 if nd[0] == "ko" else 0
            bits.append(1)
        else:
            _, pred, gate, inv = nd
            and_len[i] = 3                               # terms [pred & g, pred, g]; only the first is selected
            and_nodes[i, 0] = [pred, gate, -1]; and_inv[i, 0] = [inv, 0, -1]
            and_nodes[i, 1] = [pred, -1, -1]; and_inv[i, 1] = [inv, -1, -1]
            and_nodes[i, 2] = [gate, -1, -1]; and_inv[i, 2] = [0, -1, -1]
            bits.extend([1, 0, 0])
    parse[n] = len(bits)
    net = SynthNetwork(n, and_len, parse, len(bits), and_nodes, and_inv)
    rng = np.random.default_rng(seed)
    rows = (rng.random((n, cells)) < 0.5).astype(np.uint8)
    rows[0] = (rng.random(cells) < gate_density).astype(np.uint8)
    rows[0, 0] = 0                                       # the first cell reaches a fixed point: no undefined fill-forward
    pos = rng.permutation(n).astype(np.int32)
    inds = np.tile(np.asarray(bits, np.int32), (pop, 1))
    an = np.broadcast_to(and_nodes, (pop,) + and_nodes.shape).copy()
    ai = np.broadcast_to(and_inv, (pop,) + and_inv.shape).copy()
    return SynthProblem(net, rows, pos, inds, an, ai, ko, np.zeros(n, np.int32))

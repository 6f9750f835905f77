"""
TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's random-walk transition rule.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product path never does.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference
(`/root/reference/shallow_encoders/graph/random_walk_generator.py`) under a replayed uniform
stream and checks string-for-string equality with `walk()` below before writing
`tests/golden/walks_*.npz`; `tests/test_oracle_golden.py` re-checks this module against those
fixtures on every run.

What is restated (reference file:line):
  * DeepWalk.walk                      graph/random_walk_generator.py:61-72
  * Node2Vec.walk (CODE rule: x==t -> 1/p ; t in N(x) -> 1/q ; else 1 ; first step unbiased)
                                       graph/random_walk_generator.py:94-119
  * edge weights (1 unless every edge carries `weight`)
                                       graph/random_walk_generator.py:44-53
  * random.choices(population, weights, k=1) = inverse CDF with ONE uniform per step
                                       CPython 3.12.3 Lib/random.py `Random.choices`
                                       (third-party, not under /root/reference): cum = accumulate(w);
                                       total = cum[-1] + 0.0; idx = bisect_right(cum, u*total, 0, n-1)
  * builtin sum() of mixed int/float   CPython 3.12.3 Python/bltinmodule.c `builtin_sum_impl`:
                                       leading ints exact, first float added plainly, later floats
                                       Neumaier-compensated, later ints added plainly, final += c.
"""
from bisect import bisect_right
from itertools import accumulate
from math import isfinite
from typing import List, Optional, Sequence

import numpy as np

RULE_REFERENCE = 0  # what the reference CODE does (distance-1 -> 1/q)
RULE_PAPER = 1      # node2vec paper / reference README (distance-1 -> 1, distance-2 -> 1/q)


class OracleGraph:
    """Adjacency in the reference's CDF order (= networkx adjacency iteration order).

    adj[v]  : list of neighbour ids, in CDF order
    wts[v]  : list of edge weights (python int or float, type preserved) or None if the graph
              is "unweighted" in the reference's sense (random_walk_generator.py:46)
    names   : node names; id = lexicographic rank of the name (torch_dataset.py:99-110 ordering)
    """

    def __init__(self, adj: List[List[int]], wts: Optional[List[List]] = None, names: Optional[List[str]] = None):
        self.adj = adj
        self.wts = wts
        self.names = names if names is not None else [f'n{i:07d}' for i in range(len(adj))]
        self.nbrset = [set(a) for a in adj]

    @property
    def n_nodes(self) -> int:
        return len(self.adj)

    @staticmethod
    def from_networkx(graph) -> 'OracleGraph':
        import networkx as nx
        names = sorted(str(n) for n in graph.nodes)
        assert names == sorted(n.lower() for n in names), 'node names must be lower-case (tokenizer lowercases)'
        idx = {n: i for i, n in enumerate(names)}
        weighted = nx.is_weighted(graph)
        adj, wts = [], ([] if weighted else None)
        for n in names:
            nb = list(graph.neighbors(n))
            adj.append([idx[str(x)] for x in nb])
            if weighted:
                wts.append([graph[n][x]['weight'] for x in nb])
        return OracleGraph(adj, wts, names)

    def to_csr(self):
        """(rowptr int64[n+1], col int32[nnz], w float64[nnz] | None, w_is_int bool)"""
        deg = np.array([len(a) for a in self.adj], dtype=np.int64)
        rowptr = np.zeros(len(self.adj) + 1, dtype=np.int64)
        np.cumsum(deg, out=rowptr[1:])
        col = np.array([x for a in self.adj for x in a], dtype=np.int32)
        if self.wts is None:
            return rowptr, col, None, True
        flat = [x for a in self.wts for x in a]
        w_is_int = all(isinstance(x, (int, np.integer)) for x in flat)
        return rowptr, col, np.array(flat, dtype=np.float64), w_is_int


def py312_sum(values: Sequence):
    """Explicit restatement of CPython 3.12 `sum(values)` for a list of python ints/floats."""
    n = len(values)
    i = 0
    i_result = 0
    while i < n and isinstance(values[i], int):
        i_result += values[i]
        i += 1
    if i == n:
        return i_result
    f_result = float(i_result) + values[i]
    i += 1
    c = 0.0
    while i < n:
        x = values[i]
        if isinstance(x, float):
            t = f_result + x
            if abs(f_result) >= abs(x):
                c += (f_result - t) + x
            else:
                c += (x - t) + f_result
            f_result = t
        else:
            f_result += float(x)
        i += 1
    if c and isfinite(c):
        f_result += c
    return f_result


def transition_weights(g: OracleGraph, prev: Optional[int], node: int, p: float, q: float,
                       node2vec: bool, rule: int = RULE_REFERENCE) -> List:
    """Unnormalised weights of `node`'s neighbours, python types preserved
    (random_walk_generator.py:100-108; DeepWalk: :50-53)."""
    nbrs = g.adj[node]
    w = list(g.wts[node]) if g.wts is not None else [1 for _ in nbrs]
    if not node2vec:
        return w
    for i, x in enumerate(nbrs):
        if prev is not None and x == prev:
            w[i] *= 1 / p
            continue
        if rule == RULE_REFERENCE:
            if prev is not None and prev in g.nbrset[x]:
                w[i] *= 1 / q
        else:
            if prev is not None and prev not in g.nbrset[x]:
                w[i] *= 1 / q
    return w


def choose(weights: Sequence, u: float) -> int:
    """Index picked by `random.choices(pop, weights=[w/sum(w)], k=1)` when random() returns u."""
    s = py312_sum(list(weights))
    nw = [x / s for x in weights]
    cum = list(accumulate(nw))
    total = cum[-1] + 0.0
    return bisect_right(cum, u * total, 0, len(nw) - 1)


def walk(g: OracleGraph, start: int, length: int, uniforms: Sequence[float], p: float = 1.0, q: float = 1.0,
         node2vec: bool = False, rule: int = RULE_REFERENCE) -> List[int]:
    """One walk of `length` NODES (length-1 transitions, random_walk_generator.py:64,:98),
    consuming exactly one uniform per transition."""
    out = [start]
    prev, node = None, start
    k = 0
    while len(out) < length:
        w = transition_weights(g, prev, node, p, q, node2vec, rule)
        child = g.adj[node][choose(w, uniforms[k])]
        k += 1
        out.append(child)
        prev, node = node, child
    return out


def walks(g: OracleGraph, starts: Sequence[int], length: int, uniforms: np.ndarray, p: float = 1.0, q: float = 1.0,
          node2vec: bool = False, rule: int = RULE_REFERENCE) -> np.ndarray:
    """Batch form: uniforms[(n_walks, length-1)] float64 -> int32[(n_walks, length)]."""
    uniforms = np.asarray(uniforms, dtype=np.float64).reshape(len(starts), max(length - 1, 0))
    out = np.empty((len(starts), length), dtype=np.int32)
    for i, s in enumerate(starts):
        out[i] = walk(g, int(s), length, [float(x) for x in uniforms[i]], p, q, node2vec, rule)
    return out


def transition_probabilities(g: OracleGraph, prev: Optional[int], node: int, p: float, q: float,
                             node2vec: bool, rule: int = RULE_REFERENCE) -> np.ndarray:
    """Exact next-node distribution over g.adj[node] (for the chi-square test of the rejection sampler)."""
    w = np.array([float(x) for x in transition_weights(g, prev, node, p, q, node2vec, rule)], dtype=np.float64)
    return w / w.sum()

"""
TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's link-prediction feature path.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this module.

Parity status: PINNED.  `oracle/make_golden.py` checks `edge_embeddings` against the unmodified reference's
`shallow_encoders.graph.edge_operators` (imported from /root/reference) before writing tests/golden/edge_ops.npz.

Restated (reference file:line):
  * average / hadamard / weighted_l1 / weighted_l2     shallow_encoders/graph/edge_operators.py:10-64
  * create_edge_embeddings                              tools/graph_model_downstream_classification.py:203-224
        stack(op(E[s], E[e]) for (s, e) in edges)
  * sample_negative_edges                               tools/graph_model_downstream_classification.py:170-200
        node uniform over V; partner uniform over V \\ N(node) (the node itself is a non-neighbour); nodes without any
        non-neighbour are re-drawn.  `negative_edge_probabilities` gives the exact pair distribution for chi-square tests.
"""
from typing import List, Sequence, Tuple

import numpy as np

OPERATORS = {
    'average': lambda lhs, rhs: (lhs + rhs) / 2,
    'hadamard': lambda lhs, rhs: lhs * rhs,
    'weighted_l1': lambda lhs, rhs: np.abs(lhs - rhs),
    'weighted_l2': lambda lhs, rhs: (lhs - rhs) ** 2,
}


def edge_embeddings(node_embeddings: np.ndarray, edges: Sequence[Tuple[int, int]], op: str) -> np.ndarray:
    f = OPERATORS[op.lower()]
    return np.stack([f(node_embeddings[s, :], node_embeddings[e, :]) for s, e in edges])


def negative_edge_probabilities(adj: List[Sequence[int]]) -> np.ndarray:
    """P[node, partner] of one draw of sample_negative_edges on adjacency lists `adj`."""
    n = len(adj)
    feasible = [i for i in range(n) if len(set(adj[i])) < n]
    p = np.zeros((n, n))
    for i in feasible:
        non = sorted(set(range(n)) - set(adj[i]))
        p[i, non] = 1.0 / (len(feasible) * len(non))
    return p

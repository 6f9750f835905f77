"""
Edge embedding operators with the reference's interface (shallow_encoders/graph/edge_operators.py:10-90):
    vector(edge(n1, n2)) = f(vector(n1), vector(n2))
`edge_operator_factory(name)` returns a callable on two CUDA float32 tensors of equal shape; the arithmetic runs in the
`se_edge_op` kernel.  Batched link-prediction features straight from an embedding table and an edge list are
`edge_embeddings` (`se_edge_features`: the gather and the operator fused, nothing but the result materialised).
"""
from typing import Callable

import torch

from shallow_encoders import _native as nat

EdgeOperator = Callable[[torch.Tensor, torch.Tensor], torch.Tensor]


def average(lhs: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
    """(lhs + rhs) / 2  (reference :10-21)."""
    return nat.edge_op(lhs.contiguous(), rhs.contiguous(), 'average')


def hadamard(lhs: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
    """lhs * rhs  (reference :24-35)."""
    return nat.edge_op(lhs.contiguous(), rhs.contiguous(), 'hadamard')


def weighted_l1(lhs: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
    """|lhs - rhs|  (reference :38-49)."""
    return nat.edge_op(lhs.contiguous(), rhs.contiguous(), 'weighted_l1')


def weighted_l2(lhs: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
    """(lhs - rhs) ** 2  (reference :52-63)."""
    return nat.edge_op(lhs.contiguous(), rhs.contiguous(), 'weighted_l2')


def edge_operator_factory(name: str) -> EdgeOperator:
    """Operator by name, validated like the reference (:69-90)."""
    name = name.lower()
    operators = {'average': average, 'hadamard': hadamard, 'weighted_l1': weighted_l1, 'weighted_l2': weighted_l2}
    assert name in operators, f'Operator "{name}" is not supported. Available: {list(operators.keys())}'
    return operators[name]


def edge_embeddings(node_embeddings: torch.Tensor, src_rows: torch.Tensor, dst_rows: torch.Tensor, edge_operator_name: str) -> torch.Tensor:
    """create_edge_embeddings (tools/graph_model_downstream_classification.py:203-224) for a whole edge list in one launch:
    float32 [n_edges, emb] on the device; rows index `node_embeddings` (table rows, '<unk>' included)."""
    assert edge_operator_name.lower() in nat.EDGE_OPS, \
        f'Operator "{edge_operator_name}" is not supported. Available: {list(nat.EDGE_OPS.keys())}'
    return nat.edge_features(node_embeddings.contiguous(), src_rows.to(node_embeddings.device, torch.int64).contiguous(),
                             dst_rows.to(node_embeddings.device, torch.int64).contiguous(), edge_operator_name)

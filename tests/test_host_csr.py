"""CPU: CSR construction mirrors the reference's node ordering / weight semantics."""
import networkx as nx
import numpy as np

from helpers import random_csr
from shallow_encoders.graph.csr import CSRGraph


def test_from_networkx_karate_order_and_weights():
    g = nx.relabel_nodes(nx.karate_club_graph(), {i: f'n{i + 1:02d}' for i in range(34)})
    csr = CSRGraph.from_networkx(g, device='cpu')
    assert csr.n_nodes == 34 and csr.nnz == 156 and csr.weighted and csr.w_is_int and csr.symmetric
    assert csr.names[0] == 'n01' and csr.names[-1] == 'n34'
    rowptr, col, w = csr.rowptr.numpy(), csr.col.numpy(), csr.w.numpy()
    for i, name in enumerate(csr.names):
        nb = list(g.neighbors(name))
        assert [csr.names[c] for c in col[rowptr[i]:rowptr[i + 1]]] == nb
        assert [g[name][x]['weight'] for x in nb] == list(w[rowptr[i]:rowptr[i + 1]])
    # wcdf is the per-row prefix sum in sorted-neighbour order
    cs, wc = csr.col_sorted.numpy(), csr.wcdf.numpy()
    for i, name in enumerate(csr.names):
        row = cs[rowptr[i]:rowptr[i + 1]]
        assert list(row) == sorted(row)
        want = np.cumsum([g[name][csr.names[c]]['weight'] for c in row])
        np.testing.assert_allclose(wc[rowptr[i]:rowptr[i + 1]], want)


def test_unsorted_rows_get_a_sorted_copy():
    rowptr, col = random_csr(200, 600, 3)
    csr = CSRGraph.from_arrays(rowptr, col, device='cpu')
    assert csr.col_sorted is not csr.col and not csr.weighted
    rowptr2, col2 = random_csr(200, 600, 3, sort_rows=True)
    csr2 = CSRGraph.from_arrays(rowptr2, col2, device='cpu')
    assert csr2.col_sorted is csr2.col
    assert np.array_equal(csr.col_sorted.numpy(), csr2.col.numpy())


def test_from_edges_device_refuses_host_tensors():
    """Device ingest is a CUDA path (csrc/ingest.cu); host edge lists go through from_arrays / from_networkx."""
    import pytest
    import torch
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        CSRGraph.from_edges_device(torch.tensor([0, 1]), torch.tensor([1, 0]), 2)

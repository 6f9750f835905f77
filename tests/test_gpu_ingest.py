"""Graph ingest on the device (csrc/ingest.cu, se_csr_build): edge list -> CSR with networkx's simple-graph semantics, checked against
networkx itself and against the walk oracle's CSR on non-toy graphs (duplicates, self loops, weights, a hub longer than the
shared-memory sort, unsorted input), and the resulting graph walks bit-exactly like the oracle's."""
import numpy as np
import pytest
import torch

from helpers import cuda_device
from oracle import walk_oracle
from shallow_encoders import _native as nat
from shallow_encoders.graph.csr import CSRGraph

pytestmark = pytest.mark.gpu


def _reference_csr(src, dst, w, n, symmetrize=True):
    """networkx semantics in numpy: self loops dropped, both directions stored, LAST occurrence of a duplicate wins, rows ascending."""
    last = {}
    for e, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        if u == v:
            continue
        last[(u, v)] = e
        if symmetrize:
            last[(v, u)] = e
    keys = sorted(last)
    rows = np.array([k[0] for k in keys], dtype=np.int64)
    col = np.array([k[1] for k in keys], dtype=np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=rowptr[1:])
    wv = np.array([w[last[k]] for k in keys], dtype=np.float64) if w is not None else None
    return rowptr, col, wv


@pytest.mark.parametrize('n,m,weighted,hub', [(4, 7, False, 0), (50_000, 400_000, False, 3000), (20_000, 150_000, True, 2500), (1000, 0, False, 0),
                                             (3000, 40_000, True, 0)])
def test_csr_build_matches_networkx_semantics(n, m, weighted, hub):
    dev = cuda_device()
    rng = np.random.default_rng(n + m)
    if n == 4:
        src = np.array([0, 1, 1, 2, 3, 3, 0], dtype=np.int32); dst = np.array([1, 0, 2, 2, 0, 0, 1], dtype=np.int32)     # round 1's toy case
    else:
        src = rng.integers(0, n, m).astype(np.int32); dst = rng.integers(0, n, m).astype(np.int32)
        if m:
            dup = rng.integers(0, m, m // 10)                                    # repeated edges (in both orientations) and self loops
            src = np.concatenate([src, src[dup], dst[dup[: len(dup) // 2]], np.arange(50, dtype=np.int32)])
            dst = np.concatenate([dst, dst[dup], src[dup[: len(dup) // 2]], np.arange(50, dtype=np.int32)])
        if hub:
            nb = rng.permutation(n - 1)[:hub].astype(np.int32) + 1               # node 0 gets a row longer than the shared-memory sort (2048)
            src = np.concatenate([src, np.zeros(hub, dtype=np.int32)]); dst = np.concatenate([dst, nb])
        perm = rng.permutation(len(src))
        src, dst = src[perm], dst[perm]
    w = rng.integers(1, 9, len(src)).astype(np.float64) if weighted else None
    rowptr, col, wv, wcdf, max_deg, skipped = nat.csr_build(torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev), n,
                                                            torch.from_numpy(w).to(dev) if weighted else None, True)
    want_ptr, want_col, want_w = _reference_csr(src, dst, w, n)
    assert skipped == 0 and np.array_equal(rowptr.cpu().numpy(), want_ptr) and np.array_equal(col.cpu().numpy(), want_col)
    assert max_deg == (np.diff(want_ptr).max() if n else 0)
    if weighted:
        assert np.array_equal(wv.cpu().numpy(), want_w)
        cs = np.cumsum(want_w)
        base = np.concatenate([[0.0], cs])[want_ptr[:-1]]
        np.testing.assert_allclose(wcdf.cpu().numpy(), (cs - np.repeat(base, np.diff(want_ptr))).astype(np.float32), rtol=1e-6)
    if n == 4:
        assert rowptr.tolist() == [0, 2, 4, 5, 6] and col.tolist() == [1, 3, 0, 2, 1, 0]
    if hub:
        assert max_deg >= hub


def test_csr_build_counts_invalid_endpoints_and_equals_networkx_on_a_named_graph(tmp_path):
    dev = cuda_device()
    import networkx as nx
    src = torch.tensor([0, 5, 2, -1], dtype=torch.int32, device=dev); dst = torch.tensor([1, 1, 9, 0], dtype=torch.int32, device=dev)
    *_, skipped = nat.csr_build(src, dst, 6, None, True)
    assert skipped == 2
    with pytest.raises(IndexError):
        CSRGraph.from_edges_device(src, dst, 6)
    # an edge file in the reference's cora.cites layout -> the same graph networkx builds from it (graph/datasets.py:199-200)
    rng = np.random.default_rng(1)
    ids = rng.permutation(100000)[:800]
    pairs = [(int(ids[a]), int(ids[b])) for a, b in rng.integers(0, 800, (3000, 2)) if a != b]
    path = tmp_path / 'cora.cites'
    path.write_text('\n'.join(f'{a}\t{b}' for a, b in pairs) + '\n')
    csr = CSRGraph.from_edge_file(str(path), device=dev)
    g = nx.Graph()
    g.add_edges_from((f'n{a}', f'n{b}') for a, b in pairs)
    ref = CSRGraph.from_networkx(g, device=dev)
    assert csr.names == ref.names and torch.equal(csr.rowptr, ref.rowptr) and torch.equal(csr.col_sorted, ref.col_sorted)
    assert csr.symmetric and not csr.weighted and csr.max_degree == ref.max_degree


def test_walks_on_a_device_built_graph_are_bit_exact_against_the_oracle():
    """The CSR that comes out of se_csr_build feeds the exact walk kernel: same walks as the python oracle on the same (sorted) adjacency,
    weighted, node2vec with the reference's code rule."""
    dev = cuda_device()
    rng = np.random.default_rng(9)
    n, m = 3000, 30000
    src = rng.integers(0, n, m).astype(np.int32); dst = rng.integers(0, n, m).astype(np.int32)
    ring = np.arange(n, dtype=np.int32)
    src = np.concatenate([src, ring]); dst = np.concatenate([dst, np.roll(ring, 1)])          # no isolated node
    w = rng.integers(1, 6, len(src)).astype(np.float64)
    csr = CSRGraph.from_edges_device(torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev), n, True, torch.from_numpy(w).to(dev))
    assert csr.weighted and csr.w_is_int and csr.col is csr.col_sorted
    rowptr, col, wv = csr.rowptr.cpu().numpy(), csr.col.cpu().numpy(), csr.w.cpu().numpy()
    adj = [col[rowptr[i]:rowptr[i + 1]].tolist() for i in range(n)]
    wts = [[int(x) for x in wv[rowptr[i]:rowptr[i + 1]]] for i in range(n)]
    og = walk_oracle.OracleGraph(adj, wts, None)
    starts = rng.integers(0, n, 400).astype(np.int32)
    uni = rng.random((400, 11))
    got = nat.walk_exact(csr, torch.from_numpy(starts).to(dev), 12, 0.5, 2.0, True, 0, torch.from_numpy(uni).to(dev))
    want = walk_oracle.walks(og, starts, 12, uni, 0.5, 2.0, node2vec=True)
    assert np.array_equal(got.cpu().numpy(), want)

import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
WALK_CASES = sorted(os.path.basename(p)[len('walks_'):-len('.npz')] for p in glob.glob(os.path.join(GOLDEN, 'walks_*.npz')))
SGNS_CASES = sorted(os.path.basename(p)[len('sgns_'):-len('.npz')] for p in glob.glob(os.path.join(GOLDEN, 'sgns_*.npz')))


def cuda_device():
    """GPU tests never skip: a missing device or library is a failure."""
    import torch
    assert torch.cuda.is_available(), 'gpu-marked test run without a CUDA device'
    from shallow_encoders import _native
    _native.load()
    return torch.device('cuda:0')


def oracle_graph_from_csr(rowptr, col, w=None, w_is_int=True):
    from oracle import walk_oracle
    n = len(rowptr) - 1
    adj = [list(map(int, col[rowptr[i]:rowptr[i + 1]])) for i in range(n)]
    wts = None
    if w is not None:
        conv = int if w_is_int else float
        wts = [[conv(x) for x in w[rowptr[i]:rowptr[i + 1]]] for i in range(n)]
    return walk_oracle.OracleGraph(adj, wts)


def random_csr(n, m, seed, sort_rows=False):
    """Connected-ish undirected simple graph as CSR arrays with shuffled (unsorted) rows."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n)
    a = np.concatenate([perm[:-1], rng.integers(0, n, m)])
    b = np.concatenate([perm[1:], rng.integers(0, n, m)])
    keep = a != b
    a, b = a[keep], b[keep]
    key = np.unique(np.minimum(a, b).astype(np.int64) * n + np.maximum(a, b))
    a, b = key // n, key % n
    src, dst = np.concatenate([a, b]), np.concatenate([b, a])
    order = np.lexsort((dst if sort_rows else rng.random(len(src)), src))
    src, dst = src[order], dst[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=rowptr[1:])
    return rowptr, dst.astype(np.int32)

"""
CSR graph resident in HBM -- the layout the walk kernels read.

The reference keeps a networkx adjacency dict and rebuilds python lists on every step
(graph/random_walk_generator.py:41-48).  Here the graph is converted once:

  rowptr      int64 [n+1]
  col         int32 [nnz]   neighbours in the reference's CDF order (networkx adjacency order) -- exact mode only
  col_sorted  int32 [nnz]   each row ascending (aliases `col` when the adjacency is already sorted) -- membership
                            tests and the production sampler
  w           float64 [nnz] edge weights aligned with `col`, or None when the graph is unweighted in the
                            reference's sense (nx.is_weighted, :46)
  wcdf        float32 [nnz] per-row inclusive prefix sums of the weights aligned with `col_sorted`, or None
  node id     lexicographic rank of the lower-cased node name, so embedding row = id + 1
              (word2vec/dataloader/torch_dataset.py:99-110 puts '<unk>' at row 0)
"""
from typing import List, Optional

import numpy as np
import torch


class CSRGraph:
    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, col_sorted: torch.Tensor, w: Optional[torch.Tensor],
                 wcdf: Optional[torch.Tensor], w_is_int: bool, symmetric: bool, names: Optional[List[str]] = None,
                 max_degree: Optional[int] = None):
        self.rowptr, self.col, self.col_sorted, self.w, self.wcdf = rowptr, col, col_sorted, w, wcdf
        self.w_is_int = bool(w_is_int)
        self.symmetric = bool(symmetric)
        self.names = names
        self.n_nodes = rowptr.numel() - 1
        self.nnz = col.numel()
        if max_degree is None:
            max_degree = int((rowptr[1:] - rowptr[:-1]).max().item()) if self.n_nodes > 0 else 0
        self.max_degree = max_degree

    @property
    def device(self):
        return self.rowptr.device

    def __len__(self) -> int:
        return self.n_nodes

    @property
    def weighted(self) -> bool:
        return self.w is not None

    def nbytes(self) -> int:
        tensors = {id(t): t for t in (self.rowptr, self.col, self.col_sorted, self.w, self.wcdf) if t is not None}
        return sum(t.numel() * t.element_size() for t in tensors.values())

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def from_arrays(rowptr: np.ndarray, col: np.ndarray, w: Optional[np.ndarray] = None, w_is_int: bool = True,
                    symmetric: bool = True, names: Optional[List[str]] = None, device='cuda') -> 'CSRGraph':
        """From host CSR arrays in CDF order (this is a set-up step: per-row sort + prefix sums in numpy)."""
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        n = len(rowptr) - 1
        deg = np.diff(rowptr)
        row_of = np.repeat(np.arange(n, dtype=np.int64), deg)
        order = np.lexsort((col, row_of))                 # row-major, ascending neighbour id inside a row
        col_sorted = col[order]
        already_sorted = bool(np.array_equal(col_sorted, col))
        wcdf = None
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.float64)
            ws = w[order]
            cs = np.cumsum(ws)
            row_base = np.concatenate([[0.0], cs])[rowptr[:-1]]
            wcdf = (cs - np.repeat(row_base, deg)).astype(np.float32)
        dev = torch.device(device)
        t_col = torch.from_numpy(col).to(dev)
        t_sorted = t_col if already_sorted else torch.from_numpy(col_sorted).to(dev)
        return CSRGraph(torch.from_numpy(rowptr).to(dev), t_col, t_sorted,
                        torch.from_numpy(w).to(dev) if w is not None else None,
                        torch.from_numpy(wcdf).to(dev) if wcdf is not None else None,
                        w_is_int, symmetric, names, int(deg.max()) if n > 0 else 0)

    @staticmethod
    def from_networkx(graph, device='cuda') -> 'CSRGraph':
        """networkx.Graph -> CSR with the reference's neighbour order and weight semantics."""
        import networkx as nx
        names = sorted(str(v) for v in graph.nodes)
        index = {v: i for i, v in enumerate(names)}
        weighted = nx.is_weighted(graph)                    # all edges carry `weight` (random_walk_generator.py:46)
        rowptr = np.zeros(len(names) + 1, dtype=np.int64)
        cols, ws = [], []
        w_is_int = True
        by_name = {str(v): v for v in graph.nodes}
        for i, name in enumerate(names):
            v = by_name[name]
            nbrs = list(graph.neighbors(v))
            rowptr[i + 1] = rowptr[i] + len(nbrs)
            cols.extend(index[str(x)] for x in nbrs)
            if weighted:
                for x in nbrs:
                    wt = graph[v][x]['weight']
                    w_is_int = w_is_int and isinstance(wt, (int, np.integer)) and not isinstance(wt, bool)
                    ws.append(float(wt))
        return CSRGraph.from_arrays(rowptr, np.array(cols, dtype=np.int32),
                                    np.array(ws, dtype=np.float64) if weighted else None, w_is_int,
                                    symmetric=not graph.is_directed(), names=names, device=device)

    @staticmethod
    def from_edges_device(src: torch.Tensor, dst: torch.Tensor, n_nodes: int, symmetrize: bool = True,
                          weights: Optional[torch.Tensor] = None, names: Optional[List[str]] = None) -> 'CSRGraph':
        """CSR from an edge list in HBM, built by the library's own kernels (`se_csr_build`, csrc/ingest.cu: degree count, scans,
        bucket fill, per-row bitonic sort, duplicate removal, weight prefix sums) with networkx's simple-graph semantics: self loops
        dropped, duplicate edges collapse to the LAST occurrence's weight, rows ascending (one array serves as CDF order and as
        membership order).  `weights`: float per edge (integral values keep the reference's int-weight arithmetic in exact mode)."""
        from shallow_encoders import _native as nat
        if not src.is_cuda:
            raise RuntimeError('from_edges_device needs CUDA tensors: the B200 path has no CPU fallback (use from_arrays / from_networkx on the host)')
        w64 = weights.to(torch.float64).contiguous() if weights is not None else None
        rowptr, col, w, wcdf, max_degree, skipped = nat.csr_build(src.to(torch.int32).contiguous(), dst.to(torch.int32).contiguous(), n_nodes,
                                                                  w64, symmetrize)
        if skipped:
            raise IndexError(f'{skipped} edges have an endpoint outside [0, {n_nodes})')
        w_is_int = bool(w is None or bool((w == w.round()).all().item()))
        return CSRGraph(rowptr, col, col, w, wcdf, w_is_int, symmetrize, names, max_degree)

    @staticmethod
    def from_edge_file(path: str, device='cuda', delimiter: Optional[str] = None, prefix: str = 'n', weighted: bool = False) -> 'CSRGraph':
        """An edge-list text file (`cora.cites` of the reference, graph/datasets.py:199-200: one `cited citing` pair per line) -> CSR on
        the device.  Node names are `prefix + token` as the reference's CoraDataset names them (`n<paperid>`), node id = lexicographic
        rank of the name (so embedding row = id + 1); parsing and the name -> id map are host work, everything after is `se_csr_build`."""
        import pandas as pd
        ncols = 3 if weighted else 2
        df = pd.read_csv(path, sep=delimiter if delimiter is not None else r'\s+', header=None, usecols=list(range(ncols)), dtype=str, engine='python')
        a, b = (prefix + df[0]).str.lower(), (prefix + df[1]).str.lower()
        names = sorted(set(a) | set(b))
        index = {n: i for i, n in enumerate(names)}
        src = torch.tensor([index[x] for x in a], dtype=torch.int32, device=device)
        dst = torch.tensor([index[x] for x in b], dtype=torch.int32, device=device)
        w = torch.tensor(df[2].astype(float).values, dtype=torch.float64, device=device) if weighted else None
        return CSRGraph.from_edges_device(src, dst, len(names), True, w, names)

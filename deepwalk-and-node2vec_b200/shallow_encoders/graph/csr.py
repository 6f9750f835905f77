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
    def from_edges_device(src: torch.Tensor, dst: torch.Tensor, n_nodes: int, symmetrize: bool = True) -> 'CSRGraph':
        """Unweighted CSR from a device edge list (torch sort/unique as plumbing; used for the synthetic benchmarks).
        Removes self loops and duplicate edges; adjacency rows come out ascending, so one array serves both orders."""
        dev = src.device
        src, dst = src.to(torch.int64), dst.to(torch.int64)
        keep = src != dst
        src, dst = src[keep], dst[keep]
        if symmetrize:
            src, dst = torch.cat([src, dst]), torch.cat([dst, src])
        key = torch.unique(src * n_nodes + dst)            # sorted
        del src, dst
        rows = torch.div(key, n_nodes, rounding_mode='floor')
        col = (key - rows * n_nodes).to(torch.int32)
        del key
        deg = torch.bincount(rows, minlength=n_nodes)
        del rows
        rowptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=dev)
        torch.cumsum(deg, 0, out=rowptr[1:])
        return CSRGraph(rowptr, col, col, None, None, True, symmetrize, None, int(deg.max().item()))

"""
Random-walk generators with the reference's interface
(shallow_encoders/graph/random_walk_generator.py:11-151: RandomWalk / DeepWalk / Node2Vec / random_walk_factory),
executed by the sm_100a walk kernels over a CSR copy of the graph in HBM.

  walk(node) -> 'n1 n2 ...'         the reference's one-walk API (one kernel launch per call; kept for drop-in use)
  walk_batch(nodes) -> int32[n, L]   the batched API everything else in this package uses
  walk_exact(nodes, uniforms)        reference-exact selection under a supplied uniform stream (parity mode)

Node2Vec uses the reference CODE rule by default (candidate == previous -> 1/p; previous in N(candidate) -> 1/q; else 1;
first step unbiased, :97-108).  `rule='paper'` selects the node2vec-paper rule instead.
"""
import itertools
from abc import ABC, abstractmethod
from typing import List, Optional, Sequence, Union

import networkx as nx
import torch

from shallow_encoders import _native as nat
from shallow_encoders.graph.csr import CSRGraph

_seed_counter = itertools.count(0x5EED)


class RandomWalk(ABC):
    """Walk-method interface: a graph, a walk length counted in NODES (length - 1 transitions)."""

    def __init__(self, graph: Union[nx.Graph, CSRGraph], length: int, device: Optional[str] = None,
                 seed: Optional[int] = None):
        assert length >= 1, 'Minimum walk length is 1!'
        self._graph = graph
        self._length = length
        self._device = torch.device(device if device is not None else 'cuda')
        self._csr = graph if isinstance(graph, CSRGraph) else None
        self._seed = seed if seed is not None else next(_seed_counter)
        self._calls = 0
        self._index = None

    # -- graph access ----------------------------------------------------------------------------------------
    @property
    def csr(self) -> CSRGraph:
        if self._csr is None:
            self._csr = CSRGraph.from_networkx(self._graph, device=self._device)
        if not getattr(self, '_degrees_checked', False):
            # the reference cannot walk out of a node without neighbours (`random.choices` on an empty population raises IndexError,
            # random_walk_generator.py:68,113); the kernels would park the walk there and feed self-pairs to SGNS.  Refuse such graphs once.
            self._degrees_checked = True
            c = self._csr
            if self._length > 1 and c.n_nodes > 0 and int((c.rowptr[1:] - c.rowptr[:-1]).min().item()) == 0:
                raise IndexError('the graph has nodes without neighbours: a random walk cannot leave them (the reference raises here too)')
        return self._csr

    @property
    def length(self) -> int:
        return self._length

    @property
    def node_names(self) -> List[str]:
        """Node names in id order (id = lexicographic rank); does not touch the device."""
        if self._csr is not None and self._csr.names is not None:
            return self._csr.names
        if isinstance(self._graph, CSRGraph):
            return [f'n{i:07d}' for i in range(self._graph.n_nodes)]
        return sorted(str(v) for v in self._graph.nodes)

    def _ids(self, nodes: Union[Sequence, torch.Tensor]) -> torch.Tensor:
        if isinstance(nodes, torch.Tensor):
            return nodes.to(device=self._device, dtype=torch.int32).contiguous()
        if self._index is None:
            self._index = {name: i for i, name in enumerate(self.node_names)}
        return torch.tensor([self._index[str(n)] for n in nodes], dtype=torch.int32, device=self._device)

    # The reference's per-node accessors (:41-53).  The kernels never call them (they read the CSR); they are kept so that code written
    # against the reference's RandomWalk keeps working, for networkx graphs and for CSR graphs alike.
    def _csr_row(self, node: str):
        c = self._graph
        if self._index is None:
            self._index = {name: i for i, name in enumerate(self.node_names)}
        v = self._index[str(node)]
        lo, hi = int(c.rowptr[v].item()), int(c.rowptr[v + 1].item())
        return lo, hi

    def get_node_neighbors(self, node: str) -> List[str]:
        """Neighbours in the order the CDF is built over (networkx adjacency order; `col` of a CSR graph)."""
        if isinstance(self._graph, CSRGraph):
            lo, hi = self._csr_row(node)
            names = self.node_names
            return [names[j] for j in self._graph.col[lo:hi].tolist()]
        return list(self._graph.neighbors(node))

    def get_node_unnormalized_edge_weights(self, node: str) -> list:
        """Edge weights aligned with `get_node_neighbors`; all 1 unless EVERY edge of the graph carries a `weight` (nx.is_weighted, :45-48).
        (The reference re-scans all edges on every call; the answer cannot change during training, so it is computed once.)"""
        if isinstance(self._graph, CSRGraph):
            lo, hi = self._csr_row(node)
            if self._graph.w is None:
                return [1] * (hi - lo)
            w = self._graph.w[lo:hi].tolist()
            return [int(x) for x in w] if self._graph.w_is_int else w
        if not hasattr(self, '_nx_weighted'):
            self._nx_weighted = nx.is_weighted(self._graph)
        neighbors = list(self._graph.neighbors(node))
        if not self._nx_weighted:
            return [1] * len(neighbors)
        return [self._graph[node][neighbor]['weight'] for neighbor in neighbors]

    def get_node_normalized_edge_weights(self, node: str) -> List[float]:
        weights = self.get_node_unnormalized_edge_weights(node)
        total = sum(weights)
        return [w / total for w in weights]

    # -- walking ---------------------------------------------------------------------------------------------
    @property
    @abstractmethod
    def _params(self) -> dict:
        """p, q, node2vec, rule for the kernels."""

    def walk_batch(self, nodes: Union[Sequence, torch.Tensor], seed: Optional[int] = None, walk_id_base: int = 0,
                   walk_id_stride: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Walks for a batch of start nodes (names or int32 ids) -> int32 ids [n, length] on the device."""
        if seed is None:
            seed = self._seed + 0x9E3779B97F4A7C15 * self._calls
            self._calls += 1
        k = self._params
        return nat.walk(self.csr, self._ids(nodes), self._length, k['p'], k['q'], k['node2vec'], k['rule'], seed,
                        walk_id_base, walk_id_stride, out=out)

    def walk_exact(self, nodes: Union[Sequence, torch.Tensor], uniforms: torch.Tensor) -> torch.Tensor:
        """Reference-exact walks: uniforms[n, length-1] float64, one per transition, as `random.random()` would return."""
        k = self._params
        u = uniforms.to(device=self._device, dtype=torch.float64).contiguous()
        return nat.walk_exact(self.csr, self._ids(nodes), self._length, k['p'], k['q'], k['node2vec'], k['rule'], u)

    def to_sentences(self, walks: torch.Tensor) -> List[str]:
        names = self.node_names
        return [' '.join(names[i] for i in row) for row in walks.cpu().tolist()]

    def walk(self, node: str) -> str:
        """One walk from `node` in the reference's sentence format (`n1 n2 n3`)."""
        return self.to_sentences(self.walk_batch([node]))[0]


class DeepWalk(RandomWalk):
    """Uniform (edge-weight proportional) random walk -- reference :56-72."""

    @property
    def _params(self) -> dict:
        return {'p': 1.0, 'q': 1.0, 'node2vec': False, 'rule': nat.RULE_REFERENCE}


class Node2Vec(RandomWalk):
    """Second-order p/q biased walk -- reference :75-119."""

    def __init__(self, graph: Union[nx.Graph, CSRGraph], length: int, p: float = 1.0, q: float = 1.0,
                 rule: str = 'reference', **kwargs):
        super().__init__(graph=graph, length=length, **kwargs)
        assert rule in ('reference', 'paper'), f'Unknown rule "{rule}"'
        self._p, self._q, self._rule = p, q, rule

    @property
    def _params(self) -> dict:
        return {'p': float(self._p), 'q': float(self._q), 'node2vec': True,
                'rule': nat.RULE_REFERENCE if self._rule == 'reference' else nat.RULE_PAPER}


def random_walk_factory(name: str, graph: Union[nx.Graph, CSRGraph], length: int,
                        additional_params: Optional[dict] = None) -> RandomWalk:
    """Method name -> generator ('deepwalk', 'dfs' -> DeepWalk; 'node2vec' -> Node2Vec), reference :122-151."""
    methods = {'deepwalk': DeepWalk, 'dfs': DeepWalk, 'node2vec': Node2Vec}
    name = name.lower()
    assert name in methods, f'Unknown method "{name}". Supported: {list(methods.keys())}'
    return methods[name](graph=graph, length=length, **(additional_params or {}))

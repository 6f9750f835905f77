"""
Graph random-walk datasets with the reference's protocol (shallow_encoders/graph/datasets.py:17-221):
`DATASET_REGISTRY[name](walks_per_node, walk_length, method, **kwargs)` objects that iterate over
`len(graph) * walks_per_node` walk sentences per epoch, node order reshuffled every epoch, and expose
`.graph / .labels / .features`.

Difference in mechanism, not behaviour: an epoch's walks are produced by ONE launch of the walk kernel
(`epoch_walks()` -> int32 [n_nodes * walks_per_node, walk_length] in HBM); the string iterator only formats them.
"""
import os
import random
from typing import Dict, Iterator, Optional

import networkx as nx
import numpy as np
import torch

from shallow_encoders.common.path import ASSETS_PATH
from shallow_encoders.graph.random_walk_generator import RandomWalk, random_walk_factory
from shallow_encoders.word2vec.dataloader.registry import register_dataset


class RandomWalkDataset:
    """A graph plus a walk generator; one epoch = `walks_per_node` walks from every node (reference :17-123)."""

    def __init__(self, graph: nx.Graph, walks_per_node: int, walk_length: int, method: str = 'deepwalk',
                 method_params: Optional[dict] = None, labels: Optional[Dict[str, str]] = None,
                 features: Optional[Dict[str, np.ndarray]] = None):
        self._graph = graph
        self._labels, self._features = labels, features
        self._walk_generator: RandomWalk = random_walk_factory(
            name=method, graph=graph, length=walk_length, additional_params=method_params or {})
        self._walks_per_node = walks_per_node
        self._order = self._shuffled(len(graph))    # node ids (lexicographic rank), shuffled per epoch (:45, :87)
        self._epoch = 0
        self._pending: Optional[Iterator[str]] = None

    @staticmethod
    def _shuffled(n: int) -> torch.Tensor:
        """A fresh permutation of the node ids driven by python's `random` state, like the reference's `random.shuffle(nodes)`
        (seed `random` to reproduce it); a tensor permutation so that 10 M-node graphs do not shuffle a python list."""
        gen = torch.Generator()
        gen.manual_seed(random.getrandbits(62))
        return torch.randperm(n, generator=gen).to(torch.int32)

    # -- reference protocol --------------------------------------------------------------------------------------
    @property
    def graph(self) -> nx.Graph:
        return self._graph

    @property
    def walk_generator(self) -> RandomWalk:
        return self._walk_generator

    def __len__(self) -> int:
        return len(self._graph) * self._walks_per_node

    def __iter__(self) -> 'RandomWalkDataset':
        self._pending = iter(self._walk_generator.to_sentences(self.epoch_walks()))
        return self

    def __next__(self) -> str:
        if self._pending is None:
            self.__iter__()
        try:
            return next(self._pending)
        except StopIteration:
            self._pending = None
            raise StopIteration('Finished.')

    @property
    def has_labels(self) -> bool:
        return self._labels is not None

    @property
    def labels(self) -> Dict[str, str]:
        assert self.has_labels, 'This dataset does not have any labels!'
        return self._labels

    @property
    def has_features(self) -> bool:
        return self._features is not None

    @property
    def features(self) -> Dict[str, np.ndarray]:
        assert self.has_features, 'This dataset does not have any features!'
        return self._features

    # -- batched device API ----------------------------------------------------------------------------------------
    def epoch_starts(self) -> torch.Tensor:
        """Start node of every walk of the current epoch: node k of the shuffled order, walks_per_node times in a
        row (`nodes[index // walks_per_node]`, reference :76)."""
        return self._order.repeat_interleave(self._walks_per_node)

    def epoch_walks(self, seed: Optional[int] = None, rank: int = 0, world: int = 1) -> torch.Tensor:
        """All walks of one epoch, int32 node ids [len(self), walk_length] on the device; reshuffles the node order
        for the next epoch (reference :86-88).  With world > 1 only walks rank, rank + world, ... of the epoch are generated
        (same node order on every rank -- seed python's `random` identically -- and the same `seed`): the union over
        ranks is exactly the single-GPU epoch, because Philox is keyed by the global walk id."""
        gen = self._walk_generator
        starts = self.epoch_starts()[rank::world].contiguous().to(gen.csr.device, non_blocking=True)
        walks = gen.walk_batch(starts, seed=seed, walk_id_base=self._epoch * len(self) + rank, walk_id_stride=world)
        self._epoch += 1
        self._order = self._shuffled(len(self._graph))
        return walks


@register_dataset('graph_triplets')
class GraphTriplets(RandomWalkDataset):
    """Sanity graph: NUM_CLUSTERS components x1 - x2 - x3 (the reference adds exactly these two edges per cluster,
    datasets.py:140-141), label = cluster index."""
    NUM_CLUSTERS = 3

    def __init__(self, walks_per_node: int, walk_length: int, method: str = 'deepwalk'):
        graph, labels = nx.Graph(), {}
        for c in range(self.NUM_CLUSTERS):
            letter = chr(ord('a') + c)
            nodes = [f'{letter}{k}' for k in (1, 2, 3)]
            nx.add_path(graph, nodes)
            labels.update({n: str(c) for n in nodes})
        super().__init__(graph=graph, walks_per_node=walks_per_node, walk_length=walk_length, method=method, labels=labels)


# Zachary karate club factions as listed by the reference (datasets.py:162-171): members of the second faction
_KARATE_FACTION_2 = {10, 15, 16, 19, 21, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34}


@register_dataset('graph_karate_club')
class KarateClubDataset(RandomWalkDataset):
    """networkx karate club (weighted edges), nodes renamed n01..n34 (reference :154-180)."""

    def __init__(self, walks_per_node: int, walk_length: int, method: str = 'deepwalk', **kwargs):
        graph = nx.karate_club_graph()
        graph = nx.relabel_nodes(graph, {v: f'n{v + 1:02d}' for v in graph.nodes})
        labels = {f'n{i:02d}': ('2' if i in _KARATE_FACTION_2 else '1') for i in range(1, 35)}
        super().__init__(graph=graph, walks_per_node=walks_per_node, walk_length=walk_length, method=method,
                         labels=labels, **kwargs)


@register_dataset('graph_cora')
class CoraDataset(RandomWalkDataset):
    """Cora citation graph from `assets/cora/{cora.cites,cora.content}` (reference :183-221).  The files are not
    shipped (and cannot be downloaded here); without them use `graph_cora_synthetic`."""

    def __init__(self, walks_per_node: int, walk_length: int, method: str = 'deepwalk', **kwargs):
        import pandas as pd
        root = os.path.join(ASSETS_PATH, 'cora')
        cites, content = os.path.join(root, 'cora.cites'), os.path.join(root, 'cora.content')
        if not (os.path.exists(cites) and os.path.exists(content)):
            raise FileNotFoundError(f'Cora assets not found under {root}; use dataset_name=graph_cora_synthetic')
        edges = pd.read_csv(cites, sep='\t', header=None, names=['target', 'source']).astype('str')
        graph = nx.Graph()
        graph.add_edges_from(('n' + t, 'n' + s, {'label': 'cites'}) for t, s in zip(edges.target, edges.source))
        n_feat = 1433
        nodes = pd.read_csv(content, sep='\t', header=None, names=[f'w_{i}' for i in range(n_feat)] + ['subject'])
        nodes.index = 'n' + nodes.index.astype(str)
        labels = nodes.subject.to_dict()
        feats = nodes.iloc[:, :n_feat].to_numpy()
        features = {name: feats[i] for i, name in enumerate(nodes.index)}
        super().__init__(graph=graph, walks_per_node=walks_per_node, walk_length=walk_length, method=method,
                         labels=labels, features=features, **kwargs)


@register_dataset('graph_cora_synthetic')
class SyntheticCoraDataset(RandomWalkDataset):
    """Cora-SHAPED labelled graph (2708 nodes, ~5429 edges, 7 classes): stochastic block model, seed 0, node names
    n0000000.. (letter prefix so the reference tokenizer keeps them, torch_dataset.py:38)."""

    def __init__(self, walks_per_node: int, walk_length: int, method: str = 'deepwalk', n_nodes: int = 2708,
                 n_edges: int = 5429, n_classes: int = 7, p_in: float = 0.85, seed: int = 0, **kwargs):
        rng = np.random.default_rng(seed)
        label = rng.integers(0, n_classes, n_nodes)
        members = [np.flatnonzero(label == c) for c in range(n_classes)]
        edges = set()
        ring = np.argsort(label, kind='stable')                     # ring through label order: no isolated node
        for a, b in zip(ring, np.roll(ring, 1)):
            edges.add((min(a, b), max(a, b)))
        while len(edges) < n_edges:
            a = int(rng.integers(0, n_nodes))
            b = int(rng.choice(members[label[a]])) if rng.random() < p_in else int(rng.integers(0, n_nodes))
            if a != b:
                edges.add((min(a, b), max(a, b)))
        names = [f'n{i:07d}' for i in range(n_nodes)]
        graph = nx.Graph()
        graph.add_nodes_from(names)
        graph.add_edges_from((names[a], names[b]) for a, b in sorted(edges))
        labels = {names[i]: str(int(label[i])) for i in range(n_nodes)}
        super().__init__(graph=graph, walks_per_node=walks_per_node, walk_length=walk_length, method=method,
                         labels=labels, **kwargs)


@register_dataset('graph_powerlaw_synthetic')
class SyntheticPowerLawDataset(RandomWalkDataset):
    """Large unlabelled power-law graph generated on the device (shallow_encoders/graph/synthetic.py), the shape of
    BASELINE.json's 10 M-node / 250 M-edge configuration; registered like the reference's graphs (datasets.py:126-221) so that
    `tools/train.py --config-name=sge_sg_powerlaw_synthetic` drives the walk + SGNS hot path at scale.  The graph lives only as
    CSR in HBM (`.graph` is the CSRGraph; there is no networkx object), node names are n0000000.. in id order."""

    def __init__(self, walks_per_node: int, walk_length: int, method: str = 'node2vec', n_nodes: int = 1_000_000,
                 n_edges: int = 25_000_000, seed: int = 0, **kwargs):
        from shallow_encoders.graph.synthetic import powerlaw_graph_device
        assert n_nodes <= 10_000_000, 'node names n%07d keep their lexicographic = numeric order up to 10^7 nodes'
        csr = powerlaw_graph_device(n_nodes, n_edges, seed, 'cuda')
        super().__init__(graph=csr, walks_per_node=walks_per_node, walk_length=walk_length, method=method, **kwargs)

"""
Dataset adapters and the skip-gram collate with the reference's interface
(shallow_encoders/word2vec/dataloader/torch_dataset.py:216-322: GraphDataset, W2VCollateFunctional).

GraphDataset
  * vocabulary = ['<unk>'] + lexicographically sorted (lower-cased) node names, assigned directly from the graph --
    the same order the reference obtains from a throw-away epoch of walks (:91-110), without generating it
  * iterating yields one LongTensor[walk_length] of vocabulary ids per walk (node id + 1), like the reference
  * `epoch_tokens()` is the batched device API: int32 [n_walks, walk_length] node ids in HBM (row = id + 1)

W2VCollateFunctional
  * `sg`: centres i in [r, L - r), inputs = text[i:i+1], targets = text[i-r:i] ++ text[i+1:i+1+r]   (:300-309)
  * same-length batches are windowed with one strided view instead of a python loop per centre
"""
from collections import Counter
from typing import Dict, Iterator, List, Optional, Tuple

import networkx as nx
import numpy as np
import torch
from torch.utils.data import IterableDataset

from shallow_encoders.graph import datasets as _graph_datasets  # noqa: F401  (registers the graph datasets)
from shallow_encoders.graph.datasets import RandomWalkDataset
from shallow_encoders.word2vec.dataloader.registry import DATASET_REGISTRY

UNK = '<unk>'


class Vocab:
    """The slice of torchtext's Vocab API the reference uses (len, in, [], (), get_stoi, get_itos, default index)."""

    def __init__(self, itos: List[str]):
        self._itos = list(itos)
        self._stoi = {t: i for i, t in enumerate(self._itos)}
        self._default: Optional[int] = None

    def __len__(self) -> int:
        return len(self._itos)

    def __contains__(self, token: str) -> bool:
        return token in self._stoi

    def __getitem__(self, token: str) -> int:
        idx = self._stoi.get(token, self._default)
        if idx is None:
            raise RuntimeError(f'Token {token} not found and default index is not set')
        return idx

    def __call__(self, tokens: List[str]) -> List[int]:
        return [self[t] for t in tokens]

    def set_default_index(self, index: int) -> None:
        self._default = index

    def get_stoi(self) -> Dict[str, int]:
        return dict(self._stoi)

    def get_itos(self) -> List[str]:
        return list(self._itos)


class GraphDataset(IterableDataset):
    """Graph walks as id sequences (reference :216-273)."""

    def __init__(self, dataset_name: str, context_radius: int = 5, additional_parameters: Optional[dict] = None):
        assert dataset_name in DATASET_REGISTRY, \
            f'Dataset "{dataset_name}" is not supported. Supported: {list(DATASET_REGISTRY.keys())}'
        self._context_radius = context_radius
        self._dataset = DATASET_REGISTRY[dataset_name](**(additional_parameters or {}))
        assert isinstance(self._dataset, RandomWalkDataset), \
            f'Expected RandomWalkDataset dataset but got {type(self._dataset)}!'
        names = self._dataset.walk_generator.node_names
        assert names == sorted(n.lower() for n in names), 'node names must be lower-case and unique after lower-casing'
        self._vocab = Vocab([UNK] + names)
        self._vocab.set_default_index(self._vocab[UNK])
        self._word_frequency: Optional[Dict[str, int]] = None
        self._pending: Optional[Iterator[torch.Tensor]] = None

    # -- reference surface ---------------------------------------------------------------------------------------
    @property
    def vocab(self) -> Vocab:
        return self._vocab

    @property
    def has_labels(self) -> bool:
        return self._dataset.has_labels

    @property
    def labels(self) -> Dict[str, str]:
        return self._dataset.labels

    @property
    def has_features(self) -> bool:
        return self._dataset.has_features

    @property
    def features(self) -> Dict[str, np.ndarray]:
        return self._dataset.features

    @property
    def graph(self) -> nx.Graph:
        return self._dataset.graph

    def __len__(self) -> int:
        return len(self._dataset)

    def __iter__(self) -> 'GraphDataset':
        tokens = self.epoch_tokens()
        if tokens.shape[1] < 2 * self._context_radius + 1:      # sentence filter (:154-155): every walk is too short
            self._pending = iter(())
        else:
            self._pending = iter((tokens.to(torch.int64) + 1).cpu().unbind(0))
        return self

    def __next__(self) -> torch.Tensor:
        if self._pending is None:
            self.__iter__()
        return next(self._pending)

    def get_n_most_frequent_words(self, n: int) -> Tuple[List[str], List[int]]:
        """Most visited nodes over one epoch of walks (the reference counts its vocabulary-building epoch, :113-119)."""
        if self._word_frequency is None:
            counts = torch.bincount(self.epoch_tokens().reshape(-1).to(torch.int64) + 1, minlength=len(self._vocab))
            itos = self._vocab.get_itos()
            self._word_frequency = {itos[i]: int(c) for i, c in enumerate(counts.cpu().tolist()) if c > 0}
        top = Counter(self._word_frequency).most_common(n)
        words = [w for w, _ in top]
        return words, [self._vocab[w] for w in words]

    # -- batched device API ----------------------------------------------------------------------------------------
    @property
    def row_offset(self) -> int:
        return 1                       # '<unk>' occupies table row 0

    def epoch_tokens(self, seed: Optional[int] = None, rank: int = 0, world: int = 1) -> torch.Tensor:
        """One epoch of walks as int32 node ids [n_walks, walk_length] on the device (world > 1: this rank's share)."""
        return self._dataset.epoch_walks(seed=seed, rank=rank, world=world)


class W2VCollateFunctional:
    """Batch collation for `sg` (and the mirrored `cbow`) windows (reference :276-322)."""

    def __init__(self, mode: str, context_radius: int, max_length: int):
        assert mode.lower() in ['sg', 'cbow'], 'Invalid collate mode! Choose "sg" or "cbow"!'
        self._mode = mode.lower()
        self._context_radius = context_radius
        self._min_text_length = 2 * context_radius + 1
        self._max_length = max_length

    def _windows(self, text: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """text [..., L] -> centres [..., L-2r, 1], contexts [..., L-2r, 2r] via one strided view."""
        r = self._context_radius
        text = text[..., :self._max_length]
        text_length = text.shape[-1]
        assert text_length >= self._min_text_length, \
            f'Text is too short! [{text_length=}] < [{self._min_text_length=}]'
        win = text.unfold(-1, 2 * r + 1, 1)                                    # [..., L-2r, 2r+1]
        return win[..., r:r + 1], torch.cat([win[..., :r], win[..., r + 1:]], dim=-1)

    def __call__(self, batch_text) -> Tuple[torch.Tensor, torch.Tensor]:
        if isinstance(batch_text, torch.Tensor):
            groups = [batch_text]
        elif len({int(t.shape[0]) for t in batch_text}) == 1:
            groups = [torch.stack(list(batch_text))]
        else:
            groups = [t.unsqueeze(0) for t in batch_text]                      # ragged batch: per-sentence windows
        centres, contexts = zip(*(self._windows(g) for g in groups))
        centres = torch.cat([c.reshape(-1, 1) for c in centres])
        contexts = torch.cat([c.reshape(-1, 2 * self._context_radius) for c in contexts])
        return (centres, contexts) if self._mode == 'sg' else (contexts, centres)

"""
Dataset adapters and the window collate with the reference's interface
(shallow_encoders/word2vec/dataloader/torch_dataset.py:23-322: tokenize, W2VDataset, GraphDataset, W2VCollateFunctional).

W2VDataset (text corpora)
  * tokenize / optional lemmatisation / min-frequency vocabulary in torchtext's order ('<unk>' first, then by (-frequency, token)) /
    word frequencies / id sequences, as the reference (:23-213); the corpus is tokenised ONCE (the reference re-tokenises every epoch)
  * `epoch_token_groups()` is the batched device API of the fused engine: the epoch's sentences grouped by (clipped) length as int32
    matrices in HBM

GraphDataset
  * vocabulary = ['<unk>'] + lexicographically sorted (lower-cased) node names, assigned directly from the graph --
    the same order the reference obtains from a throw-away epoch of walks (:91-110), without generating it
  * iterating yields one LongTensor[walk_length] of vocabulary ids per walk (node id + 1), like the reference
  * `epoch_tokens()` is the batched device API: int32 [n_walks, walk_length] node ids in HBM (row = id + 1)

W2VCollateFunctional
  * `sg`: centres i in [r, L - r), inputs = text[i:i+1], targets = text[i-r:i] ++ text[i+1:i+1+r]   (:300-309)
  * same-length batches are windowed with one strided view instead of a python loop per centre
"""
import re
from collections import Counter
from typing import Dict, Iterator, List, Optional, Tuple

import networkx as nx
import numpy as np
import torch
from torch.utils.data import IterableDataset

from shallow_encoders.graph import datasets as _graph_datasets  # noqa: F401  (registers the graph datasets)
from shallow_encoders.graph.datasets import RandomWalkDataset
from shallow_encoders.word2vec.dataloader import w2v_datasets as _text_datasets  # noqa: F401  (registers the text datasets)
from shallow_encoders.word2vec.dataloader.registry import DATASET_REGISTRY

UNK = '<unk>'
_TOKEN = re.compile(r"[A-Za-z]+[\w^\']*|[\w^\']*[A-Za-z]+[\w^\']*|<unk>")     # the reference's token pattern (:38), the parity contract


def tokenize(text: str) -> List[str]:
    """Lower-case, keep tokens that contain a letter (plus `<unk>`), drop punctuation and pure numbers (reference :23-39)."""
    return _TOKEN.findall(text.lower())


def lemmatize_sentence(text: str) -> str:
    """WordNet lemmatisation of every word as adjective, adverb, noun, verb in turn (reference :42-59); needs nltk."""
    try:
        from nltk.stem import WordNetLemmatizer
    except ImportError as e:            # nltk is not installable offline: fail loudly instead of silently skipping the step
        raise NotImplementedError('lemmatize=True needs nltk (WordNetLemmatizer), which is not installed') from e
    lemmatizer = WordNetLemmatizer()
    words = text.lower().split(' ')
    for tag in ('a', 'r', 'n', 'v'):
        words = [lemmatizer.lemmatize(w, tag) for w in words]
    return ' '.join(words)


def build_vocab(token_lists, min_freq: int, specials=(UNK,)) -> 'Vocab':
    """torchtext 0.15 `build_vocab_from_iterator(..., specials, min_freq)` ordering: specials first, then tokens with
    frequency >= min_freq by (-frequency, token)."""
    counter = Counter()
    for tokens in token_lists:
        counter.update(tokens)
    for sp in specials:
        counter.pop(sp, None)
    ordered = sorted(counter.items(), key=lambda kv: (-kv[1], kv[0]))
    return Vocab(list(specials) + [t for t, f in ordered if f >= min_freq])


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


class W2VDataset(IterableDataset):
    """Text corpus as id sequences (reference :61-213)."""

    def __init__(self, dataset_name: str, context_radius: int = 5, min_word_frequency: int = 20, lemmatize: bool = False,
                 sort_by_frequency: bool = True, additional_parameters: Optional[dict] = None):
        assert dataset_name in DATASET_REGISTRY, \
            f'Dataset "{dataset_name}" is not supported. Supported: {list(DATASET_REGISTRY.keys())}'
        self._context_radius = context_radius
        self._lemmatize = lemmatize
        self._dataset = DATASET_REGISTRY[dataset_name](**(additional_parameters or {}))
        # the whole corpus, tokenised once (the reference also loads it into memory here, :91)
        self._tokens: List[List[str]] = [self.sentence_pipeline(s, apply_filter=False) for s in self._dataset]
        vocab_source = self._tokens if sort_by_frequency else [[t] for t in {t for tl in self._tokens for t in tl}]
        self._vocab = build_vocab(vocab_source, min_word_frequency)
        self._vocab.set_default_index(self._vocab[UNK])
        freq = Counter(t for tl in self._tokens for t in tl if t in self._vocab)
        self._word_frequency: Dict[str, int] = dict(freq)
        self._pending: Optional[Iterator[torch.Tensor]] = None
        self._device_groups = None

    def sentence_pipeline(self, sentence: str, apply_filter: bool = True) -> Optional[List[str]]:
        sentence = lemmatize_sentence(sentence) if self._lemmatize else sentence
        tokens = tokenize(sentence)
        if apply_filter and len(tokens) < 2 * self._context_radius + 1:
            return None
        return tokens

    def get_iterator(self, apply_filter: bool = True) -> Iterator[List[str]]:
        keep = 2 * self._context_radius + 1 if apply_filter else 0
        return (tl for tl in self._tokens if len(tl) >= keep)

    def get_n_most_frequent_words(self, n: int) -> Tuple[List[str], List[int]]:
        top = sorted(self._word_frequency.items(), key=lambda kv: kv[1], reverse=True)[:n]      # stable, like the reference's sort (:168-169)
        words = [w for w, _ in top]
        return words, [self._vocab[w] for w in words]

    @property
    def vocab(self) -> Vocab:
        return self._vocab

    @property
    def has_labels(self) -> bool:
        return False

    @property
    def labels(self) -> Dict[str, str]:
        raise NotImplementedError('This function is not implemented!')

    @property
    def word_counts(self) -> np.ndarray:
        """Occurrences per vocabulary row (row 0 = '<unk>' counts the out-of-vocabulary tokens): the unigram table for alias negatives."""
        counts = np.zeros(len(self._vocab), dtype=np.float64)
        for tl in self._tokens:
            for i in self._vocab(tl):
                counts[i] += 1
        return counts

    def __iter__(self) -> 'W2VDataset':
        self._pending = (torch.tensor(self._vocab(tl), dtype=torch.long) for tl in self.get_iterator())
        return self

    def __next__(self) -> torch.Tensor:
        if self._pending is None:
            self.__iter__()
        return next(self._pending)

    # -- batched device API ----------------------------------------------------------------------------------------
    @property
    def row_offset(self) -> int:
        return 0                       # ids ARE table rows ('<unk>' = 0 is a vocabulary entry)

    def epoch_token_groups(self, max_length: int, device='cuda') -> Dict[int, torch.Tensor]:
        """The epoch's sentences (filtered, clipped to max_length) grouped by length: {L: int32 [n_L, L] on the device}.  The fused
        kernels take equal-length sequences; a launch per length class replaces the reference's per-sentence python collate."""
        if self._device_groups is None or self._device_groups[0] != (max_length, str(device)):
            groups: Dict[int, List[List[int]]] = {}
            for tl in self.get_iterator():
                ids = self._vocab(tl)[:max_length]
                groups.setdefault(len(ids), []).append(ids)
            self._device_groups = ((max_length, str(device)),
                                   {n: torch.tensor(rows, dtype=torch.int32, device=device) for n, rows in sorted(groups.items())})
        return self._device_groups[1]


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

"""Negative sampling (reference: shallow_encoders/word2vec/utils/sampling.py:7-21)."""
import itertools
from typing import Optional

import torch

from shallow_encoders import _native as nat

_draws = itertools.count()


def generate_noise_batch(batch_size: int, n_words: int, neg_samples: int, vocab_size: int, device=None,
                         seed: Optional[int] = None, alias=None) -> torch.Tensor:
    """LongTensor (batch_size, n_words, neg_samples) of noise ids drawn on the device.  Like the reference the ids are
    iid UNIFORM over [0, vocab_size) (index 0 = '<unk>' included); pass an alias table (`_native.alias_build`) for
    unigram^power sampling.  `seed=None` advances a process-wide Philox stream."""
    device = torch.device('cuda') if device is None else torch.device(device)
    n = batch_size * n_words * neg_samples
    if seed is None:
        seed, base = 0x6E6F697365, next(_draws) << 40
    else:
        base = 0
    out = nat.sample_negatives(n, vocab_size, seed, device, alias=alias, draw_id_base=base)
    return out.view(batch_size, n_words, neg_samples)

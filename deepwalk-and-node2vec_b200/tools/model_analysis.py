#!/usr/bin/env python
"""
Closest pairs of a trained model -- the analysis step of the reference's tools/model_analysis.py:33-83 that is arithmetic
(`show_closest_pairs_for_each_word`); its t-SNE plot and the Shakespeare analogy test are visualisation / need absent assets.

    python tools/model_analysis.py --config-name=w2v_sg_abcde

For the `max_words` most frequent words (or all): cosine similarity of the word's INPUT embedding against every OUTPUT embedding
(utils/func.py:7-20) and the `pairs_per_word` best matches, written to <experiment>/analysis/closest_pairs.txt in the reference's
format.  The similarity matrix is one tcgen05 tensor-core GEMM with the row norms in its epilogue (`se_cosine_similarity`), the
selection `se_topk_rows`; nothing leaves the device but the k indices per word.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from shallow_encoders import _native as nat  # noqa: E402
from shallow_encoders.config_parser import load_config  # noqa: E402


def closest_pairs(model, dataset, max_words: int = 100, pairs_per_word: int = 5):
    """[(word, [closest words])] in the reference's order (most frequent words first when the vocabulary exceeds max_words)."""
    w_in, w_out = model.tables
    itos = dataset.vocab.get_itos()
    vocab_size = len(itos)
    if vocab_size > max_words:
        _, sampled = dataset.get_n_most_frequent_words(max_words)
    else:
        sampled = list(range(vocab_size))
    rows = torch.tensor(sampled, dtype=torch.int64, device=w_in.device)
    sim = nat.cosine_similarity(nat.table_gather_rows(w_in, rows), w_out.contiguous())
    idx, _ = nat.topk_rows(sim, min(pairs_per_word, vocab_size))
    return [(itos[w], [itos[j] for j in row]) for w, row in zip(sampled, idx.cpu().tolist())]


def show_closest_pairs_for_each_word(model, dataset, output_path: str, max_words: int = 100, pairs_per_word: int = 5) -> str:
    text = [r'Closest pairs in format "{word}:{closest_word_pairs}"']
    text += [f'{word}: {", ".join(pairs)}' for word, pairs in closest_pairs(model, dataset, max_words, pairs_per_word)]
    text = '\n'.join(text)
    os.makedirs(output_path, exist_ok=True)
    result_path = os.path.join(output_path, 'closest_pairs.txt')
    with open(result_path, 'w', encoding='utf-8') as f:
        f.write(text)
    return result_path


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--config-name', required=True)
    ap.add_argument('overrides', nargs='*')
    a = ap.parse_args(argv)
    cfg = load_config(a.config_name, a.overrides)
    base = os.path.join(cfg.path.output_dir, cfg.datamodule.dataset_name, cfg.train.experiment)
    dataset = cfg.datamodule.instantiate_dataset()
    trainer = cfg.instantiate_trainer(dataset=dataset, checkpoint_path=os.path.join(base, 'checkpoints', cfg.analysis.get('checkpoint', 'last.ckpt')))
    cp = cfg.analysis.get('closest_pairs', {}) or {}
    if cp.get('enable', True):
        path = show_closest_pairs_for_each_word(trainer.model, dataset, os.path.join(base, 'analysis'), cp.get('max_words', 100),
                                                cp.get('pairs_per_word', 5))
        print(open(path).read())


if __name__ == '__main__':
    main()

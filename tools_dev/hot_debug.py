"""dev: first-order check of the window kernel variants against the dense-gradient oracle (tiny lr)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'deepwalk-and-node2vec_b200'), ROOT, os.path.join(ROOT, 'tests')]
import philox_ref
from oracle import sgns_oracle
from shallow_encoders import _native as nat
dev = torch.device('cuda:0')
rng = np.random.default_rng(3)
vocab, emb, radius, k, offset, n_seq, length = 5000, 128, 5, 5, 1, 1024, 32
p = 1.0 / np.arange(1, vocab)
tokens = rng.choice(vocab - 1, size=(n_seq, length), p=p / p.sum()).astype(np.int32)
counts = np.concatenate([[0.0], np.bincount(tokens.ravel(), minlength=vocab - 1).astype(np.float64)])
alias = nat.alias_build(counts, 0.75, dev)
w_in = (rng.standard_normal((vocab, emb)) * 0.2).astype(np.float32)
w_out = (rng.standard_normal((vocab, emb)) * 0.2).astype(np.float32)
lr = 1e-6
n_cen = length - 2 * radius
for use_alias in (False, True):
    for kk in (0, k):
        neg = philox_ref.negatives(11, np.arange(n_seq * n_cen), 2 * radius, kk, vocab, alias['prob'].cpu().numpy() if use_alias else None,
                                   alias['alias'].cpu().numpy() if use_alias else None) if kk else np.zeros((n_seq * n_cen, 2 * radius, 0), dtype=np.int64)
        inputs, targets = sgns_oracle.windows_from_walks(tokens.astype(np.int64), radius, offset)
        o = sgns_oracle.training_step(w_in.astype(np.float64), w_out.astype(np.float64), inputs, targets, neg)
        scale = lr * inputs.shape[0] * 2 * radius
        want_out = -scale * o['grad_out']; want_in = -scale * o['grad_in']
        for name, flags in (('plain', 0), ('hot48', nat.hot_rows_flag(48)), ('nowin', nat.NO_WINDOW)):
            t_in = torch.from_numpy(w_in).to(dev); t_out = torch.from_numpy(w_out).to(dev)
            st = nat.sgns_update_walks(t_in, t_out, torch.from_numpy(tokens).to(dev), radius, kk, offset, lr, 11, alias=alias if use_alias else None, flags=flags)
            d_out = t_out.cpu().numpy().astype(np.float64) - w_out; d_in = t_in.cpu().numpy().astype(np.float64) - w_in
            e_out = np.abs(d_out - want_out); e_in = np.abs(d_in - want_in)
            r = np.unravel_index(e_out.argmax(), e_out.shape)
            print(f'alias={use_alias} K={kk} {name:6s} moved_out {np.abs(want_out).max():.3e} err_out {e_out.max():.3e} at row {r[0]} (row err {e_out[r[0]].max():.2e}, want {np.abs(want_out[r[0]]).max():.2e}) '
                  f'err_in {e_in.max():.3e} moved_in {np.abs(want_in).max():.3e} loss {st["loss"]:.6f} vs {o["loss"]:.6f}')

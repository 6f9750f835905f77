"""
TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's skip-gram window rule, SkipGram
scoring, negative-sampling loss and its gradients.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module.

Parity status: PINNED.  `oracle/make_golden.py` checks every function below against the unmodified
reference modules (model.py / loss.py / torch_dataset.py, torch autograd in fp64 and fp32) before
writing `tests/golden/sgns_*.npz` and `tests/golden/collate_*.npz`.

Restated (reference file:line):
  * W2VCollateFunctional.__call__ (sg)   shallow_encoders/word2vec/dataloader/torch_dataset.py:293-322
        centres i in [r, L-r); inputs = text[i:i+1]; targets = text[i-r:i] ++ text[i+1:i+1+r]
        worked example in the source comment (:302-306): L=8, r=3 -> centres [3, 4]
  * SkipGram.forward                     shallow_encoders/word2vec/model.py:79-91
        scores[b, n] = <W_in[inputs[b, 0]], W_out[outputs[b, n]]>
  * NegativeSamplingLoss.forward         shallow_encoders/word2vec/loss.py:14-22
        L+ = -log(clamp(sigmoid(s+), 1e-6)); L- = -sum_k log(clamp(sigmoid(-s-), 1e-6)); means over (b, n)
  * Word2VecTrainer.training_step        shallow_encoders/word2vec/trainer.py:131-152
        noise (B, N, K) -> two forward passes -> loss dict; recall = mean(sigmoid(s+) >= .5),
        precision = 1 - mean(sigmoid(s-) >= .5)
  * generate_noise_batch                 shallow_encoders/word2vec/utils/sampling.py:7-21
        iid UNIFORM ids over [0, V)  (not unigram^0.75, despite the docstring)
"""
from typing import Dict, List, Tuple

import numpy as np

CLAMP_MIN = 1e-6


def collate_sg(texts: List[np.ndarray], context_radius: int, max_length: int) -> Tuple[np.ndarray, np.ndarray]:
    r = context_radius
    inputs, targets = [], []
    for text in texts:
        text = np.asarray(text)[:max_length]
        length = text.shape[0]
        assert length >= 2 * r + 1, f'Text is too short! {length} < {2 * r + 1}'
        for i in range(r, length - r):
            inputs.append(text[i:i + 1])
            targets.append(np.concatenate([text[i - r:i], text[i + 1:i + 1 + r]]))
    return np.stack(inputs).astype(np.int64), np.stack(targets).astype(np.int64)


def sigmoid(x: np.ndarray) -> np.ndarray:
    # numerically stable, dtype-preserving
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1 / (1 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1 + e)
    return out


def scores(w_in: np.ndarray, w_out: np.ndarray, inputs: np.ndarray, outputs: np.ndarray) -> np.ndarray:
    """SkipGram.forward(proba=False): inputs (B,1), outputs (B,M) -> (B,M)."""
    c = w_in[inputs[:, 0]]                 # (B, E)
    o = w_out[outputs]                     # (B, M, E)
    return np.einsum('bme,be->bm', o, c)


def loss_dict(pos_logits: np.ndarray, neg_logits: np.ndarray) -> Dict[str, float]:
    """NegativeSamplingLoss.forward: pos (B,N), neg (B,N,K)."""
    dt = pos_logits.dtype
    lo = dt.type(CLAMP_MIN)
    pl = -np.log(np.maximum(sigmoid(pos_logits), lo))
    nl = -np.log(np.maximum(sigmoid(-neg_logits), lo)).sum(-1)
    return {'loss': (pl + nl).mean(), 'positive-loss': pl.mean(), 'negative-loss': nl.mean()}


def training_step(w_in: np.ndarray, w_out: np.ndarray, inputs: np.ndarray, targets: np.ndarray,
                  noise: np.ndarray) -> Dict[str, np.ndarray]:
    """Loss triple, recall/precision and DENSE gradients of `loss` w.r.t. both tables for a fixed
    (inputs (B,1), targets (B,N), noise (B,N,K)) batch -- closed form of what autograd produces for
    trainer.py:133-139 (checked against torch autograd in make_golden.py)."""
    dt = w_in.dtype
    b, n = targets.shape
    k = noise.shape[2]
    sp = scores(w_in, w_out, inputs, targets)                              # (B,N)
    sn = scores(w_in, w_out, inputs, noise.reshape(b, n * k)).reshape(b, n, k)
    out = loss_dict(sp, sn)
    sig_p, sig_n, sig_mn = sigmoid(sp), sigmoid(sn), sigmoid(-sn)
    out['recall'] = (sig_p >= 0.5).astype(np.float64).mean()
    out['precision'] = 1.0 - ((sig_n >= 0.5).astype(np.float64).mean() if sig_n.size else 0.0)
    scale = dt.type(1.0) / dt.type(b * n)
    # clamp(x, min) passes gradient only where x > min
    gp = -(1 - sig_p) * (sig_p > dt.type(CLAMP_MIN)) * scale               # dL/ds+
    gn = sig_n * (sig_mn > dt.type(CLAMP_MIN)) * scale                     # dL/ds-
    c = w_in[inputs[:, 0]]
    g_in = np.zeros_like(w_in)
    g_out = np.zeros_like(w_out)
    gc = np.einsum('bn,bne->be', gp, w_out[targets]) + np.einsum('bnk,bnke->be', gn, w_out[noise])
    np.add.at(g_in, inputs[:, 0], gc)
    np.add.at(g_out, targets.reshape(-1), (gp[:, :, None] * c[:, None, :]).reshape(b * n, c.shape[1]))
    np.add.at(g_out, noise.reshape(-1), (gn[:, :, :, None] * c[:, None, None, :]).reshape(b * n * k, c.shape[1]))
    out['grad_in'], out['grad_out'] = g_in, g_out
    out['pos_logits'], out['neg_logits'] = sp, sn
    return out


def sgd_step(w_in: np.ndarray, w_out: np.ndarray, inputs: np.ndarray, targets: np.ndarray, noise: np.ndarray,
             lr: float) -> Tuple[np.ndarray, np.ndarray, Dict[str, np.ndarray]]:
    """One mini-batch SGD step on the reference's mean loss: W -= lr * dLoss/dW."""
    out = training_step(w_in, w_out, inputs, targets, noise)
    dt = w_in.dtype
    return w_in - dt.type(lr) * out['grad_in'], w_out - dt.type(lr) * out['grad_out'], out


def windows_from_walks(walks: np.ndarray, context_radius: int, row_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Batched collate over a dense (n_walks, L) token matrix (every row has the same length)."""
    return collate_sg([w + row_offset for w in np.asarray(walks)], context_radius, walks.shape[1])


def sequential_window_sgd(w_in: np.ndarray, w_out: np.ndarray, tokens: np.ndarray, context_radius: int, row_offset: int,
                          noise: np.ndarray, lr: float) -> Tuple[np.ndarray, np.ndarray, float]:
    """In-place SGD applied PAIR BY PAIR in the reference's batch order (sequence, centre, context;
    torch_dataset.py:300-309 windows, loss.py:15-16 per-pair loss, un-averaged: lr multiplies dL_pair): the limit of the
    reference's training loop for batches of one pair, and what one lane group of the window kernel computes for a
    sequence -- repeated tokens inside a window included.  Per centre the centre row is read once and written once (its
    gradient accumulates over the window); every context / negative row is read and updated immediately, so a row that
    occurs twice in a window sees its own earlier update.  noise: (n_centres, 2r, K) row ids.  Returns the updated
    copies and the summed loss."""
    w_in, w_out = w_in.astype(np.float64).copy(), w_out.astype(np.float64).copy()
    r = context_radius
    loss, c_idx = 0.0, 0
    for text in np.asarray(tokens):
        rows = text.astype(np.int64) + row_offset
        for i in range(r, len(rows) - r):
            c = w_in[rows[i]].copy()
            acc = np.zeros_like(c)
            ctx = np.concatenate([rows[i - r:i], rows[i + 1:i + 1 + r]])
            for n, o in enumerate(ctx):
                targets = np.concatenate([[o], noise[c_idx, n]]) if noise is not None and noise.shape[2] else np.array([o])
                vals = w_out[targets].copy()                  # all rows of one pair are read before any is updated
                s = vals @ c
                for t, (row, sc) in enumerate(zip(targets, s)):
                    x = sc if t == 0 else -sc
                    sig = 1.0 / (1.0 + np.exp(-x))
                    loss += -np.log(max(sig, CLAMP_MIN))
                    gmag = (1.0 - sig) if sig > CLAMP_MIN else 0.0
                    step = lr * gmag if t == 0 else -lr * gmag
                    acc += step * vals[t]
                    w_out[row] += step * c
            w_in[rows[i]] += acc
            c_idx += 1
    return w_in, w_out, loss


def lazy_adam_step(w_in, w_out, state, inputs, targets, noise, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam's update (lerp / addcmul / sqrt(v) / sqrt(1 - beta2^t) + eps, step lr / (1 - beta1^t)) applied ONLY to the
    rows that receive a gradient from this batch, with the row's own step count t -- on batches that always touch the same rows it
    is dense torch.optim.Adam (checked against torch in tests/test_gpu_adam.py).  `state` = dict of m_in, v_in, m_out, v_out
    (arrays like the tables) and t_in, t_out (int arrays [V]); updated in place.  fp64."""
    out = training_step(w_in, w_out, inputs, targets, noise)
    for w, g, m, v, t, rows in ((w_in, out['grad_in'], state['m_in'], state['v_in'], state['t_in'], np.unique(inputs)),
                                (w_out, out['grad_out'], state['m_out'], state['v_out'], state['t_out'],
                                 np.unique(np.concatenate([targets.ravel(), noise.ravel()])))):
        t[rows] += 1
        tt = t[rows][:, None].astype(np.float64)
        m[rows] = m[rows] + (g[rows] - m[rows]) * (1 - beta1)
        v[rows] = v[rows] * beta2 + (1 - beta2) * g[rows] * g[rows]
        denom = np.sqrt(v[rows]) / np.sqrt(1 - beta2 ** tt) + eps
        w[rows] = w[rows] - (lr / (1 - beta1 ** tt)) * (m[rows] / denom)
    return out

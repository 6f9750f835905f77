"""Negative-sampling loss with the reference's interface (shallow_encoders/word2vec/loss.py:10-22)."""
from typing import Dict

import torch
from torch import nn

from shallow_encoders import _native as nat


class _NsLoss(torch.autograd.Function):
    """(pos (B,N), neg (B,N,K)) -> (loss, positive-loss, negative-loss), one kernel forward, stored logit gradients."""

    @staticmethod
    def forward(ctx, pos, neg):
        stats, g_pos, g_neg = nat.ns_loss(pos.detach().contiguous(), neg.detach().contiguous(), want_grads=True)
        ctx.save_for_backward(g_pos, g_neg)
        pairs = stats[4].clamp(min=1.0)
        pl, nl = (stats[0] / pairs).to(pos.dtype), (stats[1] / pairs).to(pos.dtype)
        return pl + nl, pl, nl

    @staticmethod
    def backward(ctx, g_loss, g_pl, g_nl):
        g_pos, g_neg = ctx.saved_tensors
        return g_pos * (g_loss + g_pl), (g_neg * (g_loss + g_nl)) if g_neg is not None else None


class NegativeSamplingLoss(nn.Module):
    """mean_{b,n}[-log clamp(sigmoid(s+), 1e-6) - sum_k log clamp(sigmoid(-s-), 1e-6)] and its two parts."""

    def forward(self, positive_logits: torch.Tensor, negative_logits: torch.Tensor) -> Dict[str, torch.Tensor]:
        loss, pos, neg = _NsLoss.apply(positive_logits, negative_logits)
        return {'loss': loss, 'positive-loss': pos, 'negative-loss': neg}

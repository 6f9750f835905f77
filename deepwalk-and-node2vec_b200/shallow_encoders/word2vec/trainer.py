"""
Training wrapper with the reference's interface (shallow_encoders/word2vec/trainer.py:18-165), without Lightning.

Two ways to train:

  training_step(batch)   the reference semantics: (inputs (B,1), targets (B,N)) -> uniform noise (B,N,K) -> loss dict whose
                         'loss' carries autograd to both tables (dense gradients), so `loss.backward(); optimizer.step()` with
                         the YAML's torch optimizer behaves like the reference.  Forward + backward are ONE launch of
                         `se_sgns_grad`.
  fused_step(tokens,...) the production path: walks/tokens in HBM -> windows -> negatives -> in-place SGD in one launch of
                         `se_sgns_update_walks`; no index tensors, no gradients, no optimizer state.

`fit()` replays Lightning's automatic optimisation (tools/train.py:67-83 of the reference): per batch
training_step -> zero_grad -> backward -> step; scheduler.step() per epoch; per-epoch metric means.
"""
import itertools
from typing import Dict, List, Optional, Union

import torch
from torch import nn
from torch.optim import Optimizer

from shallow_encoders import _native as nat
from shallow_encoders.word2vec.loss import NegativeSamplingLoss
from shallow_encoders.word2vec.model import W2VBase
from shallow_encoders.word2vec.utils.meter import MetricMeter
from shallow_encoders.word2vec.utils.sampling import generate_noise_batch


class _SgnsStep(torch.autograd.Function):
    """Loss triple + metrics of a fixed (inputs, targets, noise) batch in one kernel; backward hands autograd the dense
    gradients of `loss` that the same launch produced."""

    @staticmethod
    def forward(ctx, w_in, w_out, inputs, targets, noise):
        stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=w_in.device)
        res = nat.sgns_grad(w_in.detach(), w_out.detach(), inputs.contiguous(), targets.contiguous(), noise.contiguous(),
                            want_grads=True, stats=stats)
        ctx.save_for_backward(res['grad_in'], res['grad_out'])
        pairs, negs = stats[4].clamp(min=1.0), stats[5].clamp(min=1.0)
        pl, nl = (stats[0] / pairs).float(), (stats[1] / pairs).float()
        recall, precision = (stats[2] / pairs).float(), (1.0 - stats[3] / negs).float()
        ctx.mark_non_differentiable(recall, precision)
        return pl + nl, pl, nl, recall, precision

    @staticmethod
    def backward(ctx, g_loss, g_pl, g_nl, _g_recall, _g_precision):
        g_in, g_out = ctx.saved_tensors
        # only `loss` is differentiated in training (Lightning backpropagates out['loss']); its parts share the batch gradient
        return g_in * g_loss, g_out * g_loss, None, None, None


class _CbowStep(torch.autograd.Function):
    """The same for a CBOW model: inputs (B, N) context ids, targets (B, 1), noise (B, 1, K); one launch of `se_cbow_grad`."""

    @staticmethod
    def forward(ctx, w_in, w_out, inputs, targets, noise):
        stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=w_in.device)
        res = nat.cbow_grad(w_in.detach(), w_out.detach(), inputs.contiguous(), targets.contiguous(), noise.contiguous(), want_grads=True, stats=stats)
        ctx.save_for_backward(res['grad_in'], res['grad_out'])
        pairs, negs = stats[4].clamp(min=1.0), stats[5].clamp(min=1.0)
        pl, nl = (stats[0] / pairs).float(), (stats[1] / pairs).float()
        recall, precision = (stats[2] / pairs).float(), (1.0 - stats[3] / negs).float()
        ctx.mark_non_differentiable(recall, precision)
        return pl + nl, pl, nl, recall, precision

    @staticmethod
    def backward(ctx, g_loss, g_pl, g_nl, _g_recall, _g_precision):
        g_in, g_out = ctx.saved_tensors
        return g_in * g_loss, g_out * g_loss, None, None, None


class Word2VecTrainer(nn.Module):
    """Trains a W2V model."""

    def __init__(self, model: W2VBase, optimizer: Optional[Optimizer], scheduler, neg_samples: int, vocab_size: int):
        super().__init__()
        self._optimizer = optimizer
        self._scheduler = scheduler
        self._loss_func = NegativeSamplingLoss()
        self._neg_samples = neg_samples
        self._vocab_size = vocab_size
        self._model = model
        self._meter = MetricMeter()
        self.logged: Dict[str, List[float]] = {}
        self.current_epoch = 0
        self.global_step = 0
        self._fused_launch = itertools.count()

    # -- reference surface ---------------------------------------------------------------------------------------
    @property
    def model(self) -> W2VBase:
        return self._model

    @property
    def optimizer(self) -> Optimizer:
        return self._optimizer

    @optimizer.setter
    def optimizer(self, optimizer: Optimizer) -> None:
        self._optimizer = optimizer

    @property
    def scheduler(self):
        return self._scheduler

    @scheduler.setter
    def scheduler(self, scheduler) -> None:
        self._scheduler = scheduler

    def log(self, name: str, value, **_kwargs) -> None:
        """Scalar log (Lightning's `self.log`); kept in memory, `tools/train.py` writes it out per epoch."""
        self.logged.setdefault(name, []).append(float(value))

    def forward(self, inputs: torch.Tensor, outputs: torch.Tensor, proba: bool = True) -> torch.Tensor:
        return self._model(inputs, outputs, proba=proba)

    def training_step(self, batch: List[torch.Tensor], *args, **kwargs) -> Dict[str, torch.Tensor]:
        inputs, outputs = batch
        w_in, w_out = self._model._input_embedding.weight, self._model._output_embedding.weight
        inputs, outputs = inputs.to(w_in.device), outputs.to(w_in.device)
        noise = generate_noise_batch(outputs.shape[0], outputs.shape[1], self._neg_samples, self._vocab_size, device=w_in.device)
        from shallow_encoders.word2vec.model import CBOW
        # the reference's two forward passes look the rows up through nn.Embedding(max_norm): same in-place renormalisation first
        self._model.renorm_(inputs, torch.cat([outputs.reshape(-1), noise.reshape(-1)]))
        step = _CbowStep if isinstance(self._model, CBOW) else _SgnsStep          # cbow collate: inputs (B, N) contexts, outputs (B, 1) centre
        loss, pos, neg, recall, precision = step.apply(w_in, w_out, inputs, outputs, noise)
        out = {'loss': loss, 'positive-loss': pos, 'negative-loss': neg}
        for name, value in out.items():
            value = value.detach()
            assert not torch.isnan(value).any(), f'Got nan value for key "{name}"!'
            self._meter.push(f'train-epoch/{name}', value)
            self.log(f'train/{name}', value)
        if self._optimizer is not None:
            self.log('epoch/lr', self._optimizer.param_groups[0]['lr'])
        self._meter.push('train-metrics/recall', recall)
        self._meter.push('train-metrics/precision', precision)
        return out

    def on_train_epoch_end(self) -> Dict[str, float]:
        means = dict(self._meter.get_all())
        for name, value in means.items():
            self.log(name, value)
        return means

    def configure_optimizers(self):
        return [self._optimizer], [self._scheduler]

    # -- loops ---------------------------------------------------------------------------------------------------
    def fit(self, dataloader, max_epochs: int, on_epoch_end=None) -> None:
        """Lightning-style automatic optimisation over `dataloader` batches of (inputs, targets)."""
        scheduler = self._scheduler['scheduler'] if isinstance(self._scheduler, dict) else self._scheduler
        for epoch in range(max_epochs):
            self.current_epoch = epoch
            for batch in dataloader:
                out = self.training_step(batch)
                self._optimizer.zero_grad()
                out['loss'].backward()
                self._optimizer.step()
                self.global_step += 1
            if scheduler is not None:
                scheduler.step()
            means = self.on_train_epoch_end()
            if on_epoch_end is not None:
                on_epoch_end(self, epoch, means)

    def fused_step(self, tokens: torch.Tensor, context_radius: int, lr: float, row_offset: int = 1, seed: int = 0,
                   alias=None, flags: int = nat.SCATTER_RED, stats: Optional[torch.Tensor] = None,
                   local_negatives: bool = False, check_tokens: bool = False) -> Optional[Dict[str, float]]:
        """In-place SGNS update from int32 token sequences [n_seq, L] in HBM; `lr` multiplies the un-averaged per-pair
        gradient (for the reference's mean loss over a launch of P pairs pass lr_batch / P)."""
        if getattr(self._model, 'max_norm', None) is not None:
            raise NotImplementedError('max_norm needs the rows of a step before it runs; the in-place SGD kernel draws its negatives '
                                      'in-kernel.  Use the Adam engine (train.engine=fused with torch.optim.Adam) or train.engine=reference.')
        w_in, w_out = self._model.tables
        launch = next(self._fused_launch)
        n_cen = tokens.shape[1] - 2 * context_radius
        return nat.sgns_update_walks(w_in, w_out, tokens, context_radius, self._neg_samples, row_offset, lr, seed,
                                     centre_id_base=launch * tokens.shape[0] * max(n_cen, 1), alias=alias, flags=flags, stats=stats,
                                     local_negatives=local_negatives, check_tokens=check_tokens)

    def fused_adam_step(self, tokens: torch.Tensor, context_radius: int, row_offset: int = 1, seed: int = 0,
                        stats: Optional[torch.Tensor] = None) -> Optional[Dict[str, float]]:
        """One mini-batch of the reference's loop with the YAML's Adam, on the device end to end: windows of the int32 token
        sequences [n_seq, L] (torch_dataset.py:300-309) -> uniform noise (sampling.py:21) -> `se_sgns_adam_step` (mean-loss
        gradient + row-sparse Adam).  The optimizer must be a RowSparseAdam (its lr follows the YAML's scheduler)."""
        from shallow_encoders.word2vec.optim import RowSparseAdam
        assert isinstance(self._optimizer, RowSparseAdam), 'fused_adam_step needs the RowSparseAdam optimizer'
        w_in, w_out = self._model._input_embedding.weight.data, self._model._output_embedding.weight.data
        r = context_radius
        rows = tokens.to(torch.int64) + row_offset
        win = rows.unfold(1, 2 * r + 1, 1)                                        # (n_seq, L - 2r, 2r + 1): one window per centre
        inputs = win[:, :, r].reshape(-1, 1).contiguous()
        targets = torch.cat([win[:, :, :r], win[:, :, r + 1:]], dim=2).reshape(-1, 2 * r).contiguous()
        launch = next(self._fused_launch)
        noise = generate_noise_batch(targets.shape[0], 2 * r, self._neg_samples, self._vocab_size, device=w_in.device,
                                     seed=seed * 0x9E3779B1 + launch)
        self._model.renorm_(inputs, torch.cat([targets.reshape(-1), noise.reshape(-1)]))
        return self._optimizer.step_batch(w_in, w_out, inputs, targets, noise, stats=stats)

    # -- checkpoints (state-dict keys `_model._input_embedding.weight`, `_model._output_embedding.weight`) ---------------
    def save_checkpoint(self, path: str) -> None:
        torch.save({'state_dict': {k: v.detach().cpu() for k, v in self.state_dict().items()},
                    'epoch': self.current_epoch, 'global_step': self.global_step}, path)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path: str, **init_kwargs) -> 'Word2VecTrainer':
        trainer = cls(**init_kwargs)
        ckpt = torch.load(checkpoint_path, map_location='cpu')
        trainer.load_state_dict(ckpt['state_dict'])
        trainer.current_epoch, trainer.global_step = ckpt.get('epoch', 0), ckpt.get('global_step', 0)
        return trainer

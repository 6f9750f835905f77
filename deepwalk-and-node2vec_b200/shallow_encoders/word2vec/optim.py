"""
Row-sparse Adam: what `train.optimizer._target_: torch.optim.Adam` (every shipped YAML, e.g. configs/sge_sg_karate_club.yaml:32-34;
instantiated at shallow_encoders/config_parser/core.py:43-94 of the reference) becomes in the fused engine.

It IS a `torch.optim.Optimizer` over the model's two tables, so the YAML's scheduler (`torch.optim.lr_scheduler.StepLR`)
drives `param_groups[0]['lr']` exactly as in the reference; but the update itself is `step_batch`, one call of the C ABI's
`se_sgns_adam_step`: gradient of the reference's mean loss for an explicit (inputs, targets, noise) batch, then Adam on the
rows that batch touched (moments, bias correction by the row's own step count).  No dense gradient, no pass over the other
V - touched rows.  `step()` (the dense autograd path) is refused: use torch.optim.Adam with `train.engine: reference` for that.
"""
from typing import Dict, Optional

import torch

from shallow_encoders import _native as nat


class RowSparseAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, **_unused):
        if weight_decay:
            raise NotImplementedError('row-sparse Adam has no weight decay (it would touch every row every step)')
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        self._state: Optional[nat.AdamState] = None

    @classmethod
    def from_torch_config(cls, params, cfg: Dict) -> 'RowSparseAdam':
        """Build from a YAML optimizer block whose `_target_` is torch.optim.Adam (same keyword names)."""
        kw = {k: v for k, v in cfg.items() if k != '_target_'}
        if kw.get('amsgrad'):
            raise NotImplementedError('amsgrad is not supported by the row-sparse Adam kernel')
        return cls(params, **kw)

    def step(self, closure=None):
        raise RuntimeError('RowSparseAdam updates through step_batch(model, inputs, targets, noise); the dense autograd path is '
                           'torch.optim.Adam with train.engine=reference')

    def step_batch(self, w_in: torch.Tensor, w_out: torch.Tensor, inputs: torch.Tensor, targets: torch.Tensor,
                   noise: Optional[torch.Tensor], stats: Optional[torch.Tensor] = None):
        group = self.param_groups[0]
        self._opt_called = True                     # lr schedulers check that the optimizer stepped before they do
        if self._state is None:
            self._state = nat.AdamState(w_in.shape[0], w_in.shape[1], w_in.device)
        return nat.sgns_adam_step(w_in, w_out, inputs, targets, noise, self._state, group['lr'], group['betas'][0], group['betas'][1],
                                  group['eps'], stats=stats)

"""
Word2vec models with the reference's interface (shallow_encoders/word2vec/model.py:10-110: W2VBase, SkipGram, CBOW).

The two tables are ordinary `nn.Embedding` parameters (state-dict keys `_input_embedding.weight`,
`_output_embedding.weight`, Xavier-uniform init) living in HBM; scoring and its backward run in the sm_100a kernels
(`se_skipgram_scores`, `se_skipgram_scores_backward`) through a `torch.autograd.Function`, so the reference's training
loop (loss.backward() + any torch optimizer) works unchanged.  The fast path does not go through here at all: see
`Word2VecTrainer.fused_step`, which updates the tables in place from walks.
"""
from typing import Optional

import torch
from torch import nn

from shallow_encoders import _native as nat


class _SkipGramScores(torch.autograd.Function):
    """scores[b, j] = <W_in[inputs[b]], W_out[outputs[b, j]]> with dense gradients, as autograd gives for model.py:85-88."""

    @staticmethod
    def forward(ctx, w_in, w_out, inputs, outputs):
        inputs, outputs = inputs.contiguous(), outputs.contiguous()
        ctx.save_for_backward(w_in, w_out, inputs, outputs)
        return nat.skipgram_scores(w_in.detach(), w_out.detach(), inputs, outputs, proba=False)

    @staticmethod
    def backward(ctx, grad_scores):
        w_in, w_out, inputs, outputs = ctx.saved_tensors
        g_in, g_out = torch.zeros_like(w_in), torch.zeros_like(w_out)
        nat.skipgram_scores_backward(w_in.detach(), w_out.detach(), inputs, outputs, grad_scores.contiguous(), g_in, g_out)
        return g_in, g_out, None, None


class W2VBase(nn.Module):
    """Input and context embedding tables."""

    def __init__(self, vocab_size: int, embedding_size: int, max_norm: Optional[float] = None, device=None,
                 shard: Optional[dict] = None):
        """`shard` (not in the reference, which is single-device): {'rank', 'world', 'exchange', 'seed'} -> both tables are
        ONE pair of `ShardedTable`s striped over the GPUs of the node instead of per-process `nn.Embedding`s; only the
        fused engine and `forward` without autograd work on them."""
        super().__init__()
        self.max_norm = None if max_norm is None else float(max_norm)
        if self.max_norm is not None and shard is not None and shard.get('world', 1) > 1:
            raise NotImplementedError('max_norm is not supported on striped multi-GPU tables')
        device = torch.device('cuda' if device is None else device)
        self._sharded = None
        if shard is not None and shard.get('world', 1) > 1:
            from shallow_encoders.word2vec.sharded import ReplicatedTable, ShardedTable
            bound = (6.0 / (vocab_size + embedding_size)) ** 0.5           # xavier_uniform_ (reference :26-27)
            # 'synced': a working copy per GPU + row-sharded masters, synchronised after every step (reference-exact global negatives at
            # NVLink bulk rate); otherwise ONE striped pair that the kernels read / update over NVLink per pair
            make = ReplicatedTable if shard.get('mode') == 'synced' else ShardedTable
            self._sharded = tuple(make(vocab_size, embedding_size, device, shard['rank'], shard['world'], shard['exchange'])
                                  for _ in range(2))
            for k, t in enumerate(self._sharded):
                t.fill_uniform(bound, int(shard.get('seed', 0)) * 2 + k)
            self._publish(device)
            return
        self._input_embedding = nn.Embedding(vocab_size, embedding_size, device=device)
        self._output_embedding = nn.Embedding(vocab_size, embedding_size, device=device)
        torch.nn.init.xavier_uniform_(self._input_embedding.weight)
        torch.nn.init.xavier_uniform_(self._output_embedding.weight)

    @staticmethod
    def _publish(device) -> None:
        """Every rank has written the stripes it owns: nobody may gather from, or red.add into, a peer's stripes before that peer's
        fill / load kernel has run (fresh cuMemCreate memory is not zeroed, and the owner's plain stores would overwrite early updates)."""
        import torch.distributed as dist
        torch.cuda.synchronize(device)
        if dist.is_available() and dist.is_initialized():
            dist.barrier()

    @property
    def input_embedding(self) -> torch.Tensor:
        """Input embedding weights as a CPU tensor (reference :29-37)."""
        if self._sharded is not None:
            return self._sharded[0].to_tensor().cpu()
        return self._input_embedding.weight.to('cpu').data

    @property
    def output_embedding(self) -> torch.Tensor:
        """Context embedding weights as a CPU tensor (reference :39-47)."""
        if self._sharded is not None:
            return self._sharded[1].to_tensor().cpu()
        return self._output_embedding.weight.to('cpu').data

    @property
    def tables(self):
        """(W_in, W_out) the fused kernels update in place: device tensors, or striped tables."""
        if self._sharded is not None:
            return self._sharded
        return self._input_embedding.weight.data, self._output_embedding.weight.data

    # striped tables are not nn.Parameters: give checkpoints the reference's keys anyway (trainer.py state-dict keys
    # `_model._input_embedding.weight`, `_model._output_embedding.weight`)
    def state_dict(self, *args, destination=None, prefix='', keep_vars=False):
        if self._sharded is None:
            return super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        destination = {} if destination is None else destination
        destination[prefix + '_input_embedding.weight'] = self._sharded[0].to_tensor()
        destination[prefix + '_output_embedding.weight'] = self._sharded[1].to_tensor()
        return destination

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        if self._sharded is None:
            return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        for table, key in zip(self._sharded, ('_input_embedding.weight', '_output_embedding.weight')):
            if hasattr(table, 'load_owned'):
                table.load_owned(state_dict[prefix + key])
            else:
                table.load(state_dict[prefix + key])
        self._publish(self._sharded[0].device)

    def renorm_(self, inputs: Optional[torch.Tensor] = None, outputs: Optional[torch.Tensor] = None) -> None:
        """`nn.Embedding(max_norm=...)` (reference :22-23; configs/w2v_sg_abcde.yaml:7) renormalises the rows it looks up, in
        place, before using them: the same side effect for the ids about to be scored.  No-op when max_norm is None."""
        if self.max_norm is None:
            return
        if inputs is not None:
            nat.table_renorm_rows(self._input_embedding.weight.data, inputs, self.max_norm)
        if outputs is not None:
            nat.table_renorm_rows(self._output_embedding.weight.data, outputs, self.max_norm)

    def embed_inputs(self, inputs: torch.Tensor) -> torch.Tensor:
        if self._sharded is not None:
            return self._sharded[0].gather(inputs.reshape(-1)).reshape(*inputs.shape, -1)
        self.renorm_(inputs=inputs)
        return self._input_embedding(inputs)

    def embed_outs(self, outputs: torch.Tensor) -> torch.Tensor:
        if self._sharded is not None:
            return self._sharded[1].gather(outputs.reshape(-1)).reshape(*outputs.shape, -1)
        self.renorm_(outputs=outputs)
        return self._output_embedding(outputs)


class SkipGram(W2VBase):
    """inputs (B, 1), outputs (B, N) -> scores (B, N); sigmoid applied when `proba` (reference :79-91)."""

    def forward(self, inputs: torch.Tensor, outputs: torch.Tensor, proba: bool = True) -> torch.Tensor:
        if self._sharded is not None:      # scoring only: striped tables train through the fused engine, not autograd
            dev = self._sharded[0].device
            return nat.skipgram_scores(self._sharded[0], self._sharded[1], inputs.to(dev).reshape(-1).contiguous(),
                                       outputs.to(dev).contiguous(), proba=proba)
        w_in, w_out = self._input_embedding.weight, self._output_embedding.weight
        inputs = inputs.to(w_in.device).reshape(-1)
        outputs = outputs.to(w_in.device)
        self.renorm_(inputs, outputs)
        if not (torch.is_grad_enabled() and (w_in.requires_grad or w_out.requires_grad)):
            return nat.skipgram_scores(w_in.detach(), w_out.detach(), inputs.contiguous(), outputs.contiguous(), proba=proba)
        scores = _SkipGramScores.apply(w_in, w_out, inputs, outputs)
        return torch.sigmoid(scores) if proba else scores


class _CbowScores(torch.autograd.Function):
    """scores[b, j] = <mean_n W_in[inputs[b, n]], W_out[outputs[b, j]]>; backward = what autograd gives for model.py:103-106,
    evaluated by the same kernel family (dense gradients through `se_cbow_grad`-style accumulation in torch index_add form)."""

    @staticmethod
    def forward(ctx, w_in, w_out, inputs, outputs):
        inputs, outputs = inputs.contiguous(), outputs.contiguous()
        ctx.save_for_backward(w_in, w_out, inputs, outputs)
        return nat.cbow_scores(w_in.detach(), w_out.detach(), inputs, outputs, proba=False)

    @staticmethod
    def backward(ctx, grad_scores):
        w_in, w_out, inputs, outputs = ctx.saved_tensors
        b, n = inputs.shape
        h = w_in.detach()[inputs].mean(dim=1)                                              # (B, E)
        g_out = torch.zeros_like(w_out).index_add_(0, outputs.reshape(-1), (grad_scores.unsqueeze(-1) * h.unsqueeze(1)).reshape(-1, h.shape[1]))
        gh = torch.einsum('bm,bme->be', grad_scores, w_out.detach()[outputs]) / n
        g_in = torch.zeros_like(w_in).index_add_(0, inputs.reshape(-1), gh.unsqueeze(1).expand(b, n, -1).reshape(-1, h.shape[1]))
        return g_in, g_out, None, None


class CBOW(W2VBase):
    """inputs (B, N) context ids, outputs (B, 1) centre ids (or (B, K) noise) -> scores (B, M) (reference :94-110)."""

    def forward(self, inputs: torch.Tensor, outputs: torch.Tensor, proba: bool = True) -> torch.Tensor:
        if self._sharded is not None:
            raise NotImplementedError('CBOW runs on per-GPU tables (the striped multi-GPU tables serve the SkipGram fused engine)')
        w_in, w_out = self._input_embedding.weight, self._output_embedding.weight
        inputs, outputs = inputs.to(w_in.device), outputs.to(w_in.device)
        self.renorm_(inputs, outputs)
        if not (torch.is_grad_enabled() and (w_in.requires_grad or w_out.requires_grad)):
            return nat.cbow_scores(w_in.detach(), w_out.detach(), inputs.contiguous(), outputs.contiguous(), proba=proba)
        scores = _CbowScores.apply(w_in, w_out, inputs, outputs)
        return torch.sigmoid(scores) if proba else scores

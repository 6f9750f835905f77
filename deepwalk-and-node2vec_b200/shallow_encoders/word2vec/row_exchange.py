"""
The NCCL BASELINE for multi-GPU SGNS: both tables row-sharded by `row % world` in ordinary per-GPU tensors, rows
fetched and gradients returned with `all_to_all_single` around a local gradient kernel -- the textbook row-sharded
word2vec.  It exists to be measured against the product path (`sharded.ShardedTable`: one fused kernel
over NVLink peer memory, no collective) and to cross-check it; it is synchronous mini-batch SGD per micro-batch
(every pair of a micro-batch sees the same row values), not Hogwild.

Per micro-batch of walks on every rank:
  windows (torch_dataset.py:300-309) + negatives (utils/sampling.py:21)  ->  unique row ids per table
  ids --all_to_all--> owners gather rows --all_to_all--> compact [n_unique x emb] tables
  se_sgns_grad on the compact tables (trainer.py:131-152 + backward, one launch)
  gradients --all_to_all--> owners apply  W[row] -= lr_pair * dL/dW   (duplicates from different ranks accumulate)

The exchange itself (`fetch` / `push`) is torch plumbing and device-agnostic, which is what the world-size-2 gloo
test covers; the compute is the CUDA kernel only (no CPU fallback).
"""
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from shallow_encoders import _native as nat


class RowShardedTables:
    def __init__(self, vocab: int, emb: int, rank: int, world: int, device, group=None):
        self.vocab, self.emb, self.rank, self.world, self.group = int(vocab), int(emb), int(rank), int(world), group
        self.device = torch.device(device)
        n_local = (self.vocab - self.rank + self.world - 1) // self.world          # rows rank, rank + world, ...
        self.local = {'in': torch.zeros((n_local, emb), dtype=torch.float32, device=self.device),
                      'out': torch.zeros((n_local, emb), dtype=torch.float32, device=self.device)}
        self.exchanged_bytes = 0

    # -- set-up ------------------------------------------------------------------------------------------------------
    def load_full(self, which: str, full: torch.Tensor) -> None:
        """Keep this rank's rows of a full [vocab x emb] table."""
        self.local[which].copy_(full[self.rank::self.world].to(self.device))

    def fill_uniform(self, bound: float, seed_in: int, seed_out: int) -> None:
        """Same content as `table_fill_uniform` gives a full table (keyed by the global element index)."""
        for which, seed in (('in', seed_in), ('out', seed_out)):
            full = torch.empty((self.vocab, self.emb), dtype=torch.float32, device=self.device)
            nat.table_fill_uniform(full, bound, seed)
            self.load_full(which, full)
            del full

    def gather_full(self, which: str) -> torch.Tensor:
        """Full table on every rank (tests / checkpoints)."""
        n_max = (self.vocab + self.world - 1) // self.world                         # all_gather wants equal shapes: pad
        mine = torch.zeros((n_max, self.emb), dtype=torch.float32, device=self.device)
        mine[:self.local[which].shape[0]] = self.local[which]
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(parts, mine, group=self.group)
        else:
            parts[0] = mine
        parts = [parts[r][:(self.vocab - r + self.world - 1) // self.world] for r in range(self.world)]
        full = torch.empty((self.vocab, self.emb), dtype=torch.float32, device=self.device)
        for r in range(self.world):
            full[r::self.world] = parts[r]
        return full

    # -- exchange ----------------------------------------------------------------------------------------------------
    def _a2a(self, send: torch.Tensor, send_counts, recv_counts) -> torch.Tensor:
        out = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        if self.world == 1:
            out.copy_(send)
        else:
            dist.all_to_all_single(out, send.contiguous(), output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts),
                                   group=self.group)
        self.exchanged_bytes += send.numel() * send.element_size()
        return out

    def plan(self, ids: torch.Tensor) -> Dict:
        """Route unique row ids to their owners; returns what fetch/push need (ids stay in the caller's order)."""
        owner = ids % self.world
        order = torch.argsort(owner, stable=True)
        send_counts = torch.bincount(owner, minlength=self.world)
        recv_counts = torch.empty_like(send_counts)
        if self.world > 1:
            dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        else:
            recv_counts.copy_(send_counts)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        wanted = self._a2a((ids[order] // self.world).contiguous(), sc, rc)          # local row indices the peers ask me for
        return {'order': order, 'send': sc, 'recv': rc, 'wanted': wanted}

    def fetch(self, which: str, plan: Dict) -> torch.Tensor:
        rows = self._a2a(self.local[which].index_select(0, plan['wanted']), plan['recv'], plan['send'])
        out = torch.empty_like(rows)
        out[plan['order']] = rows
        return out

    def push(self, which: str, plan: Dict, grads: torch.Tensor, lr: float) -> None:
        got = self._a2a(grads.index_select(0, plan['order']), plan['send'], plan['recv'])
        self.local[which].index_add_(0, plan['wanted'], got, alpha=-float(lr))

    # -- one training step -----------------------------------------------------------------------------------------
    @staticmethod
    def windows(walks: torch.Tensor, radius: int, row_offset: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(centres (B,1), contexts (B,2r)) int64 rows of the skip-gram windows (torch_dataset.py:300-309)."""
        win = walks.to(torch.int64).unfold(1, 2 * radius + 1, 1) + row_offset            # (n, L-2r, 2r+1)
        centres = win[..., radius].reshape(-1, 1)
        contexts = torch.cat([win[..., :radius], win[..., radius + 1:]], dim=-1).reshape(-1, 2 * radius)
        return centres, contexts

    def step(self, walks: torch.Tensor, radius: int, n_neg: int, row_offset: int, lr: float, seed: int, draw_id_base: int = 0,
             micro_walks: int = 8192, local_negatives: bool = False, stats: Optional[torch.Tensor] = None) -> None:
        """All micro-batches of `walks` (int32 [n, L]); `lr` multiplies the un-averaged per-pair gradient, as in the fused
        kernels.  Every rank must call it with the same number of micro-batches."""
        n_cen = walks.shape[1] - 2 * radius
        for lo in range(0, walks.shape[0], micro_walks):
            mb = walks[lo:lo + micro_walks]
            centres, contexts = self.windows(mb, radius, row_offset)
            b, n = contexts.shape
            if local_negatives and self.world > 1:
                n_local = self.local['out'].shape[0]
                noise = nat.sample_negatives(b * n * n_neg, n_local, seed, self.device, draw_id_base=draw_id_base + lo * n_cen * n * n_neg)
                noise = noise * self.world + self.rank
            else:
                noise = nat.sample_negatives(b * n * n_neg, self.vocab, seed, self.device, draw_id_base=draw_id_base + lo * n_cen * n * n_neg)
            in_ids, in_inv = torch.unique(centres.reshape(-1), return_inverse=True)
            out_ids, out_inv = torch.unique(torch.cat([contexts.reshape(-1), noise]), return_inverse=True)
            p_in, p_out = self.plan(in_ids), self.plan(out_ids)
            c_in, c_out = self.fetch('in', p_in), self.fetch('out', p_out)
            res = nat.sgns_grad(c_in, c_out, in_inv.reshape(b, 1), out_inv[:b * n].reshape(b, n).contiguous(),
                                out_inv[b * n:].reshape(b, n, n_neg).contiguous(), want_grads=True, stats=stats)
            scale = lr * b * n                        # se_sgns_grad returns gradients of the MEAN loss over the b*n pairs
            self.push('in', p_in, res['grad_in'], scale)
            self.push('out', p_out, res['grad_out'], scale)

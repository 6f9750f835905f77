#!/usr/bin/env python
"""
Trains a SkipGram model from a YAML config -- the reference's entry point (tools/train.py:46-83):

    python tools/train.py --config-name=sge_sg_karate_club [train.max_epochs=5 train.engine=fused ...]

Outputs follow the reference's conventions (tools/conventions.py:1-27):
    runs/<dataset>/<experiment>/checkpoints/{checkpoint_epoch=..._step=....ckpt, last.ckpt}
    runs/<dataset>/<experiment>/run_history/train_<timestamp>.yaml
    runs/tb_logs/<dataset>/<experiment>/scalars.jsonl            (scalar log; TensorBoard itself is out of scope)
Checkpoints hold {'state_dict': {'_model._input_embedding.weight', '_model._output_embedding.weight'}, ...}.

Engines (`train.engine`):
    reference   Lightning's loop restated: DataLoader batches of `batch_size` walks -> collate -> training_step ->
                backward -> YAML optimizer -> per-epoch scheduler.  Same arithmetic as the reference, kernels on the B200.
    fused       per epoch ONE walk-kernel launch, then per `batch_size` walks (the reference's mini-batch), all on the device:
                `_target_: torch.optim.Adam` (every shipped YAML) -> windows + noise + `se_sgns_adam_step` (mean-loss gradient and
                row-sparse Adam with the YAML's lr / betas / eps; lr stepped by the YAML's scheduler);
                `_target_: torch.optim.SGD` or an explicit train.fused_lr -> one fused window/negatives/in-place-SGD launch with
                lr / (pairs per launch) per pair, decayed by the YAML's StepLR schedule.
An existing experiment directory is replaced without the reference's interactive prompt (train.py:36-42) unless --keep.

Several GPUs (not in the reference, whose YAMLs all say `devices: '1'`): launch under torchrun, one process per GPU,
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/train.py --config-name=... train.engine=fused
Every rank generates walks rank, rank+G, ... of each epoch.  train.multi_gpu_negatives=synced (default): each GPU trains a working copy
with the reference's global negative draw and after every mini-batch ONE kernel per table sums the updates into row-sharded masters and
writes the rows back over NVLink (csrc/replica.cu).  global | local | owner: both tables exist ONCE, striped over the G HBMs, and the
fused kernel reads / updates peer rows per pair (Hogwild across GPUs, barrier per epoch).  Rank 0 writes the checkpoints.
"""
import argparse
import json
import os
import shutil
import sys
import time
from datetime import datetime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import yaml  # noqa: E402

from shallow_encoders.config_parser import load_config  # noqa: E402


def experiment_dirs(output_dir: str, dataset: str, experiment: str):
    base = os.path.join(output_dir, dataset, experiment)
    return {'checkpoints': os.path.join(base, 'checkpoints'), 'run_history': os.path.join(base, 'run_history'),
            'tb_logs': os.path.join(output_dir, 'tb_logs', dataset, experiment)}


def distributed_context():
    """(rank, world, shard spec or None); initialises NCCL when launched under torchrun with WORLD_SIZE > 1."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1:
        return 0, 1, None
    import random
    import torch.distributed as dist
    from shallow_encoders.word2vec.sharded import make_exchange
    rank, local = int(os.environ['RANK']), int(os.environ.get('LOCAL_RANK', os.environ['RANK']))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    random.seed(0x5EED)                   # every rank must shuffle the node order identically (graph/datasets.py)
    return rank, world, {'rank': rank, 'world': world, 'exchange': make_exchange(rank, world), 'seed': 0x5EED}


def train(cfg, keep: bool = False, quiet: bool = False):
    rank, world, shard = distributed_context()
    if world > 1 and cfg.train.engine != 'fused':
        raise ValueError('multi-GPU training runs the fused engine on striped tables: pass train.engine=fused')
    quiet = quiet or rank != 0
    dirs = experiment_dirs(cfg.path.output_dir, cfg.datamodule.dataset_name, cfg.train.experiment)
    if rank == 0:
        for key in ('checkpoints', 'tb_logs'):
            if os.path.exists(dirs[key]) and not keep:
                shutil.rmtree(dirs[key])
        for d in dirs.values():
            os.makedirs(d, exist_ok=True)
    if world > 1:
        torch.distributed.barrier()

    dataset = cfg.datamodule.instantiate_dataset()
    if shard is not None:
        shard['mode'] = 'synced' if (not cfg.train.local_negatives and cfg.train.multi_gpu_negatives == 'synced') else 'striped'
    trainer = cfg.instantiate_trainer(dataset=dataset, shard=shard)
    scalars = open(os.path.join(dirs['tb_logs'], 'scalars.jsonl'), 'a') if rank == 0 else None

    def end_of_epoch(tr, epoch, means):
        if world > 1:                     # all ranks' updates are in the striped tables before rank 0 reads them
            torch.cuda.synchronize()
            torch.distributed.barrier()
        if rank != 0:
            if world > 1:
                torch.distributed.barrier()
            return
        name = f'checkpoint_epoch={epoch:06d}_step={tr.global_step:09d}.ckpt'
        tr.save_checkpoint(os.path.join(dirs['checkpoints'], name))
        tr.save_checkpoint(os.path.join(dirs['checkpoints'], 'last.ckpt'))
        scalars.write(json.dumps({'epoch': epoch, 'step': tr.global_step, **means}) + '\n')
        scalars.flush()
        if not quiet:
            print(f'epoch {epoch:3d} step {tr.global_step:7d} ' + ' '.join(f'{k}={v:.4f}' for k, v in means.items()), flush=True)
        if world > 1:
            torch.distributed.barrier()

    t0 = time.time()
    if cfg.train.engine == 'reference':
        trainer.fit(cfg.datamodule.instantiate_dataloader(dataset=dataset), cfg.train.max_epochs, on_epoch_end=end_of_epoch)
    elif cfg.train.engine == 'fused':
        fit_fused(cfg, dataset, trainer, end_of_epoch, rank, world)
    else:
        raise ValueError(f'unknown train.engine "{cfg.train.engine}"')
    torch.cuda.synchronize()
    if scalars is not None:
        scalars.close()
    if not quiet:
        print(f'trained {cfg.train.max_epochs} epochs in {time.time() - t0:.1f}s; checkpoints in {dirs["checkpoints"]}')
    return trainer, dataset


def owner_centre_id_base(epoch: int, lo: int, share: int, world: int, n_walks_global: int, n_cen: int) -> int:
    """Philox id of centre 0 of RANK 0's walks in the owner-computes step that starts at walk `lo` of every rank's share of `epoch`: a function
    of (epoch, lo) only, so every rank derives the same base whatever its own iteration count is (ranks may hold one walk more or less when
    len(dataset) % world != 0), and consecutive steps / epochs get disjoint id ranges of world * share * n_cen centres each."""
    steps_per_epoch = -(-n_walks_global // (share * world))
    return (epoch * steps_per_epoch + lo // share) * world * share * n_cen


def fit_fused(cfg, dataset, trainer, end_of_epoch, rank: int = 0, world: int = 1):
    """Walk kernel + fused SGNS kernel; one launch per `batch_size` walks so that a launch is the reference's mini-batch.
    world > 1: this rank's share of every batch (walks rank, rank + world, ...) against the striped tables."""
    r = cfg.datamodule.context_radius
    sched = cfg.train.scheduler.get('scheduler', cfg.train.scheduler)
    step_size, gamma = int(sched.get('step_size', 10 ** 9)), float(sched.get('gamma', 1.0))
    from shallow_encoders import _native as nat
    stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device='cuda')
    kind = cfg.train.fused_optimizer_kind()
    if kind == 'adam':
        if world > 1:
            raise ValueError('the row-sparse Adam engine is single-GPU; multi-GPU runs use the in-place SGD kernels (set train.fused_lr)')
        return fit_fused_adam(cfg, dataset, trainer, end_of_epoch, stats)
    base_lr = cfg.train.fused_sgd_lr()
    if not cfg.datamodule.is_graph:
        if world > 1:
            raise ValueError('text corpora train on one GPU')
        return fit_fused_text_sgd(cfg, dataset, trainer, end_of_epoch, stats, base_lr, step_size, gamma)
    n_walks_global = len(dataset)
    for epoch in range(cfg.train.max_epochs):
        trainer.current_epoch = epoch
        lr_batch = base_lr * gamma ** (epoch // step_size)
        tokens = dataset.epoch_tokens(rank=rank, world=world)[:, :cfg.datamodule.max_length]
        stats.zero_()
        share = -(-cfg.datamodule.batch_size // world)             # this rank's walks of one global batch
        mode = 'local' if getattr(cfg.train, 'local_negatives', False) else getattr(cfg.train, 'multi_gpu_negatives', 'synced')
        mode = mode if world > 1 else 'global'
        n_cen = tokens.shape[1] - 2 * r
        n_min = len(dataset) // world                              # walks every rank is guaranteed to have this epoch
        for lo in range(0, tokens.shape[0], share):
            chunk = tokens[lo:lo + share].contiguous()
            pairs = chunk.shape[0] * n_cen * 2 * r * world
            seed = epoch * 1_000_003 + lo
            if mode == 'owner' and lo + share <= n_min:          # same decision on every rank (the step contains a collective);
                                                                 # a ragged last batch falls back to fetching the rows
                from shallow_encoders.word2vec.sharded import sgns_update_walks_owner_computes
                w_in, w_out = trainer.model.tables
                cid_base = owner_centre_id_base(epoch, lo, share, world, n_walks_global, n_cen)
                sgns_update_walks_owner_computes(w_in, w_out, chunk, r, cfg.train.loss.negative_samples, dataset.row_offset, lr_batch / pairs,
                                                 seed, cid_base, rank, world, stats=stats)
            else:
                trainer.fused_step(chunk, r, lr_batch / pairs, row_offset=dataset.row_offset, seed=seed * world + rank, stats=stats,
                                   local_negatives=mode == 'local')
                if mode == 'synced' and lo + share <= n_min:     # every rank reaches this together (the sync contains barriers)
                    from shallow_encoders.word2vec.sharded import sync_replicated
                    sync_replicated(trainer.model.tables, merge=cfg.train.multi_gpu_merge)
            trainer.global_step += 1
        if mode == 'synced':                                     # a ragged last batch is folded in here
            from shallow_encoders.word2vec.sharded import sync_replicated
            sync_replicated(trainer.model.tables, merge=cfg.train.multi_gpu_merge)
        if world > 1:
            torch.distributed.all_reduce(stats)
        s = stats.tolist()
        p = max(s[4], 1.0)
        means = {'train-epoch/loss': (s[0] + s[1]) / p, 'train-epoch/positive-loss': s[0] / p,
                 'train-epoch/negative-loss': s[1] / p, 'train-metrics/recall': s[2] / p,
                 'train-metrics/precision': 1 - s[3] / max(s[5], 1.0), 'epoch/lr': lr_batch}
        for name, value in means.items():
            trainer.log(name, value)
        end_of_epoch(trainer, epoch, means)


def _epoch_batches(cfg, dataset):
    """Equal-length int32 token matrices of at most `batch_size` sequences: one per mini-batch of a graph epoch, one per
    (length class, mini-batch) of a text epoch (sentences are grouped by clipped length; the kernels take equal-length rows)."""
    bs = cfg.datamodule.batch_size
    if cfg.datamodule.is_graph:
        tokens = dataset.epoch_tokens()[:, :cfg.datamodule.max_length]
        groups = [tokens]
    else:
        groups = list(dataset.epoch_token_groups(cfg.datamodule.max_length).values())
    for tokens in groups:
        for lo in range(0, tokens.shape[0], bs):
            yield tokens[lo:lo + bs]


def _epoch_means(stats, lr):
    s = stats.tolist()
    p = max(s[4], 1.0)
    return {'train-epoch/loss': (s[0] + s[1]) / p, 'train-epoch/positive-loss': s[0] / p, 'train-epoch/negative-loss': s[1] / p,
            'train-metrics/recall': s[2] / p, 'train-metrics/precision': 1 - s[3] / max(s[5], 1.0), 'epoch/lr': lr}


def fit_fused_text_sgd(cfg, dataset, trainer, end_of_epoch, stats, base_lr, step_size, gamma):
    """Text corpus through the fused window / negatives / in-place SGD kernel, one launch per mini-batch of equal-length sentences."""
    r = cfg.datamodule.context_radius
    from shallow_encoders import _native as nat
    for epoch in range(cfg.train.max_epochs):
        trainer.current_epoch = epoch
        lr_batch = base_lr * gamma ** (epoch // step_size)
        stats.zero_()
        for chunk in _epoch_batches(cfg, dataset):
            pairs = chunk.shape[0] * (chunk.shape[1] - 2 * r) * 2 * r
            trainer.fused_step(chunk.contiguous(), r, lr_batch / pairs, row_offset=dataset.row_offset, seed=epoch * 1_000_003 + trainer.global_step,
                               stats=stats, flags=nat.SCATTER_RED | nat.WINDOW_REFRESH)
            trainer.global_step += 1
        means = _epoch_means(stats, lr_batch)
        for name, value in means.items():
            trainer.log(name, value)
        end_of_epoch(trainer, epoch, means)


def fit_fused_adam(cfg, dataset, trainer, end_of_epoch, stats):
    """The reference's loop (tools/train.py:67-83: per batch training_step -> backward -> Adam.step; scheduler.step per epoch) with every
    stage on the device: walks of the epoch from one kernel launch, then per `batch_size` walks one `fused_adam_step`."""
    r = cfg.datamodule.context_radius
    if cfg.datamodule.mode.lower() != 'sg':
        raise ValueError('the fused engine trains SkipGram (mode: sg); CBOW runs with train.engine=reference')
    scheduler = trainer.scheduler['scheduler'] if isinstance(trainer.scheduler, dict) else trainer.scheduler
    for epoch in range(cfg.train.max_epochs):
        trainer.current_epoch = epoch
        stats.zero_()
        for chunk in _epoch_batches(cfg, dataset):
            trainer.fused_adam_step(chunk, r, row_offset=dataset.row_offset, seed=epoch, stats=stats)
            trainer.global_step += 1
        lr = trainer.optimizer.param_groups[0]['lr']
        if scheduler is not None:
            scheduler.step()
        means = _epoch_means(stats, lr)
        for name, value in means.items():
            trainer.log(name, value)
        end_of_epoch(trainer, epoch, means)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--config-name', required=True)
    ap.add_argument('--config-path', default=None)
    ap.add_argument('--keep', action='store_true', help='do not delete an existing experiment directory')
    ap.add_argument('overrides', nargs='*', help='hydra-style a.b.c=value overrides')
    a = ap.parse_args(argv)
    kwargs = {'config_path': a.config_path} if a.config_path else {}
    cfg = load_config(a.config_name, a.overrides, **kwargs)
    dirs = experiment_dirs(cfg.path.output_dir, cfg.datamodule.dataset_name, cfg.train.experiment)
    os.makedirs(dirs['run_history'], exist_ok=True)
    stamp = datetime.now().strftime('%Y-%m-%d_%H-%M-%S.%f')
    with open(os.path.join(dirs['run_history'], f'train_{stamp}.yaml'), 'w') as fh:
        yaml.safe_dump({'config_name': a.config_name, 'overrides': a.overrides}, fh)
    train(cfg, keep=a.keep)


if __name__ == '__main__':
    main()

"""
YAML experiment configs with the reference's schema (shallow_encoders/config_parser/core.py:28-327), without
hydra / omegaconf / pydantic (none of them is installable here): PyYAML + dataclasses + importlib.

Contract kept: top-level keys `model`, `datamodule`, `train`, `analysis`, `downstream`, `path`; a leading
`defaults: [w2v_config]` is tolerated; objects are built from `_target_` dotted paths with the remaining keys as kwargs
(`model` gets `vocab_size=len(dataset.vocab)`, the optimizer `params=`, the scheduler `optimizer=`); CLI overrides are
hydra-style `a.b.c=value`.  Extra key understood here: `train.engine: reference|fused` (+ `train.fused_lr`).
"""
import copy
import importlib
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import yaml

from shallow_encoders.common.path import CONFIG_PATH, RUNS_PATH


def locate(target: str):
    module, _, attr = target.rpartition('.')
    return getattr(importlib.import_module(module), attr)


def instantiate(cfg: Dict[str, Any], **extra):
    cfg = dict(cfg)
    target = cfg.pop('_target_')
    return locate(target)(**cfg, **extra)


@dataclass
class TrainLossConfig:
    negative_samples: int


@dataclass
class TrainConfig:
    experiment: str
    optimizer: dict
    scheduler: dict
    loss: TrainLossConfig
    max_epochs: int
    accelerator: str = 'gpu'
    devices: str = '1'
    engine: str = 'reference'          # 'reference': autograd + YAML optimizer (dense); 'fused': device-resident engine (below)
    # fused engine: the YAML optimizer decides the kernel -- torch.optim.Adam -> row-sparse Adam (se_sgns_adam_step), torch.optim.SGD ->
    # in-place SGD kernel with the YAML's lr on the mean loss.  An explicit fused_lr forces the in-place SGD kernel with that
    # mean-loss learning rate (a launch over P pairs applies fused_lr / P per pair) whatever the optimizer block says.
    fused_lr: Optional[float] = None
    local_negatives: bool = False      # multi-GPU fused engine: draw negatives among the rows the GPU owns
    # multi-GPU fused engine: synced (default: working copy per GPU + row-sharded masters, the reference's global draw, one fused
    # reduce-scatter/all-gather kernel per step) | global | local | owner (ONE striped table pair accessed per pair over NVLink)
    multi_gpu_negatives: str = 'synced'
    multi_gpu_merge: Any = 'sum'       # synced mode: sum (summed updates; measured best) | mean (model averaging) | weight in (0, 1]

    def fused_optimizer_kind(self) -> str:
        """'sgd' or 'adam': which kernel the fused engine runs for this config."""
        if self.fused_lr is not None:
            return 'sgd'
        target = str(self.optimizer.get('_target_', '')).rsplit('.', 1)[-1].lower()
        if target == 'adam':
            return 'adam'
        if target == 'sgd':
            return 'sgd'
        raise ValueError(f'train.engine=fused supports torch.optim.Adam (row-sparse Adam kernel) and torch.optim.SGD (in-place SGD kernel); '
                         f'got {self.optimizer.get("_target_")!r}.  Set train.fused_lr to force the SGD kernel, or use train.engine=reference.')

    def fused_sgd_lr(self) -> float:
        return float(self.fused_lr) if self.fused_lr is not None else float(self.optimizer.get('lr', 0.01))

    def instantiate_optimizer(self, params):
        if self.engine == 'fused' and self.fused_optimizer_kind() == 'adam':
            from shallow_encoders.word2vec.optim import RowSparseAdam
            return RowSparseAdam.from_torch_config(params, self.optimizer)
        return instantiate(self.optimizer, params=params)

    def instantiate_scheduler(self, optimizer):
        if '_target_' in self.scheduler:
            return instantiate(self.scheduler, optimizer=optimizer)
        assert 'scheduler' in self.scheduler, 'Missing scheduler object in scheduler configuration.'
        sched = copy.deepcopy(self.scheduler)
        sched['scheduler'] = instantiate(sched['scheduler'], optimizer=optimizer)
        return sched


@dataclass
class DatamoduleConfig:
    dataset_name: str
    mode: str
    context_radius: int
    max_length: int
    is_graph: bool
    batch_size: int
    num_workers: int = 0
    min_word_frequency: int = 0
    lemmatize: bool = False
    additional_parameters: dict = field(default_factory=dict)

    def instantiate_dataset(self):
        from shallow_encoders.word2vec.dataloader.torch_dataset import GraphDataset, W2VDataset
        if not self.is_graph:           # text corpora (reference core.py:115-134)
            return W2VDataset(dataset_name=self.dataset_name, context_radius=self.context_radius, min_word_frequency=self.min_word_frequency,
                              lemmatize=self.lemmatize, additional_parameters=self.additional_parameters)
        return GraphDataset(dataset_name=self.dataset_name, context_radius=self.context_radius,
                            additional_parameters=self.additional_parameters)

    def instantiate_collate_fn(self):
        from shallow_encoders.word2vec.dataloader.torch_dataset import W2VCollateFunctional
        return W2VCollateFunctional(mode=self.mode, context_radius=self.context_radius, max_length=self.max_length)

    def instantiate_dataloader(self, dataset=None):
        """Batches of `batch_size` walks -> (inputs, targets).  The walks come out of one kernel launch per epoch, so
        `num_workers` worker processes (which in the reference each generate a full extra copy of the epoch,
        core.py:173-178) are not spawned; set `additional_parameters.walks_per_node` x8 to train on as many walks."""
        from torch.utils.data import DataLoader
        dataset = self.instantiate_dataset() if dataset is None else dataset
        return DataLoader(dataset, batch_size=self.batch_size, num_workers=0, collate_fn=self.instantiate_collate_fn())


@dataclass
class PathConfig:
    output_dir: str = RUNS_PATH


@dataclass
class GlobalConfig:
    train: TrainConfig
    datamodule: DatamoduleConfig
    model: dict
    analysis: dict = field(default_factory=dict)
    path: PathConfig = field(default_factory=PathConfig)
    downstream: dict = field(default_factory=dict)

    def instantiate_model(self, dataset=None, shard: Optional[dict] = None):
        """`shard` = {'rank', 'world', 'exchange', 'seed'}: stripe the two tables over the GPUs of the node (multi-GPU runs)."""
        dataset = self.datamodule.instantiate_dataset() if dataset is None else dataset
        extra = {'shard': shard} if shard is not None else {}
        return instantiate(self.model, vocab_size=len(dataset.vocab), **extra)

    def instantiate_trainer(self, model=None, optimizer=None, scheduler=None, dataset=None, checkpoint_path: Optional[str] = None,
                            shard: Optional[dict] = None):
        from shallow_encoders.word2vec.trainer import Word2VecTrainer
        dataset = self.datamodule.instantiate_dataset() if dataset is None else dataset
        model = self.instantiate_model(dataset=dataset, shard=shard) if model is None else model
        params = list(model.parameters())
        if params:
            optimizer = self.train.instantiate_optimizer(params) if optimizer is None else optimizer
            scheduler = self.train.instantiate_scheduler(optimizer) if scheduler is None else scheduler
        # striped tables are not nn.Parameters: they train through the fused engine (in-place SGD), no torch optimizer
        kwargs = dict(model=model, optimizer=optimizer, scheduler=scheduler,
                      neg_samples=self.train.loss.negative_samples, vocab_size=len(dataset.vocab))
        if checkpoint_path is None:
            return Word2VecTrainer(**kwargs)
        return Word2VecTrainer.load_from_checkpoint(checkpoint_path=checkpoint_path, **kwargs)


def _set_path(tree: dict, dotted: str, value) -> None:
    keys = dotted.split('.')
    for k in keys[:-1]:
        tree = tree.setdefault(k, {})
    tree[keys[-1]] = value


def load_config(config_name: str, overrides: Optional[List[str]] = None, config_path: str = CONFIG_PATH) -> GlobalConfig:
    """`config_name` = YAML stem (or path); `overrides` = ['train.max_epochs=5', ...] parsed as YAML scalars."""
    path = config_name if os.path.exists(config_name) else os.path.join(
        config_path, config_name if config_name.endswith(('.yaml', '.yml')) else config_name + '.yaml')
    with open(path) as fh:
        raw = yaml.safe_load(fh)
    raw.pop('defaults', None)
    for item in overrides or []:
        key, _, value = item.partition('=')
        _set_path(raw, key.lstrip('+'), yaml.safe_load(value))
    train = dict(raw['train'])
    train['loss'] = TrainLossConfig(**train['loss'])
    train['devices'] = str(train.get('devices', '1'))
    return GlobalConfig(
        train=TrainConfig(**train), datamodule=DatamoduleConfig(**raw['datamodule']), model=raw['model'],
        analysis=raw.get('analysis', {}) or {}, path=PathConfig(**(raw.get('path') or {})),
        downstream=raw.get('downstream', {}) or {})

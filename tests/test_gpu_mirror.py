"""GPU: the reference-facing Python surface drives the kernels and reproduces the reference's results."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, SGNS_CASES, cuda_device
from shallow_encoders.graph.datasets import KarateClubDataset
from shallow_encoders.word2vec import trainer as trainer_mod
from shallow_encoders.word2vec.dataloader.torch_dataset import GraphDataset, W2VCollateFunctional
from shallow_encoders.word2vec.loss import NegativeSamplingLoss
from shallow_encoders.word2vec.model import SkipGram
from shallow_encoders.word2vec.trainer import Word2VecTrainer
from shallow_encoders.word2vec.utils.sampling import generate_noise_batch

pytestmark = pytest.mark.gpu


def test_walk_api_strings_and_exact_mode_match_golden():
    cuda_device()
    ds = KarateClubDataset(walks_per_node=8, walk_length=10, method='node2vec', method_params={'p': 1, 'q': 0.5})
    gen = ds.walk_generator
    sentence = gen.walk('n01')                                  # the reference's one-walk API
    toks = sentence.split(' ')
    assert len(toks) == 10 and toks[0] == 'n01' and all(ds.graph.has_edge(a, b) for a, b in zip(toks, toks[1:]))
    z = np.load(os.path.join(GOLDEN, 'walks_karate_yaml.npz'))
    names = [str(s) for s in z['names']]
    got = gen.walk_exact([names[i] for i in z['starts']], torch.from_numpy(z['uniforms']))
    assert np.array_equal(got.cpu().numpy(), z['walks'])        # same walks as the reference under the same draws
    assert gen.to_sentences(got[:1])[0] == ' '.join(names[i] for i in z['walks'][0])


def test_dataset_iteration_protocol():
    cuda_device()
    ds = KarateClubDataset(walks_per_node=4, walk_length=10, method='deepwalk')
    first = list(ds)
    assert len(first) == len(ds) == 136 and all(len(s.split(' ')) == 10 for s in first)
    starts = [s.split(' ')[0] for s in first]
    assert all(starts[i] == starts[i - i % 4] for i in range(136)) and len(set(starts)) == 34
    second = list(ds)
    assert [s.split(' ')[0] for s in second] != starts          # node order reshuffled per epoch (datasets.py:87)
    gd = GraphDataset('graph_karate_club', context_radius=2,
                      additional_parameters={'walks_per_node': 4, 'walk_length': 10, 'method': 'node2vec', 'method_params': {'p': 1, 'q': 0.5}})
    rows = list(gd)
    assert len(rows) == 136 and rows[0].dtype == torch.int64 and rows[0].shape == (10,) and not rows[0].is_cuda
    assert min(int(r.min()) for r in rows) >= 1 and max(int(r.max()) for r in rows) <= 34
    inputs, targets = W2VCollateFunctional('sg', 2, 256)(rows[:64])
    assert inputs.shape == (384, 1) and targets.shape == (384, 4)        # B' = 64 * (10 - 4), N = 2r
    words, idx = gd.get_n_most_frequent_words(3)
    assert len(words) == 3 and idx == [gd.vocab[w] for w in words]
    short = GraphDataset('graph_karate_club', context_radius=5, additional_parameters={'walks_per_node': 1, 'walk_length': 10})
    assert list(short) == []                                     # sentences shorter than 2r+1 are filtered (:154-155)


@pytest.mark.parametrize('tag', SGNS_CASES)
def test_trainer_training_step_matches_reference(tag, monkeypatch):
    dev = cuda_device()
    z = np.load(os.path.join(GOLDEN, f'sgns_{tag}.npz'))
    vocab, emb = z['w_in_f32'].shape
    model = SkipGram(vocab_size=vocab, embedding_size=emb)
    with torch.no_grad():
        model._input_embedding.weight.copy_(torch.from_numpy(z['w_in_f32']))
        model._output_embedding.weight.copy_(torch.from_numpy(z['w_out_f32']))
    noise = torch.from_numpy(z['noise']).to(dev)
    monkeypatch.setattr(trainer_mod, 'generate_noise_batch', lambda *a, **k: noise)      # pin the noise (trainer.py:133)
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    tr = Word2VecTrainer(model, opt, None, neg_samples=noise.shape[2], vocab_size=vocab)
    out = tr.training_step([torch.from_numpy(z['inputs']), torch.from_numpy(z['targets'])])
    got = np.array([float(out['loss']), float(out['positive-loss']), float(out['negative-loss'])])
    np.testing.assert_allclose(got, z['loss_f32'], rtol=1e-5)
    out['loss'].backward()
    den = max(np.abs(z['grad_in_f32']).max(), np.abs(z['grad_out_f32']).max())
    assert np.abs(model._input_embedding.weight.grad.cpu().numpy() - z['grad_in_f32']).max() / den <= 1e-5
    assert np.abs(model._output_embedding.weight.grad.cpu().numpy() - z['grad_out_f32']).max() / den <= 1e-5
    opt.step()                                                                           # any torch optimizer works on top
    np.testing.assert_allclose(model.input_embedding.numpy(), z['w_in_f32'] - 0.5 * z['grad_in_f32'], atol=1e-5 * max(den, 1.0))
    means = tr.on_train_epoch_end()
    np.testing.assert_allclose([means['train-metrics/recall'], means['train-metrics/precision']], z['metrics_f32'], atol=1e-6)
    assert sorted(tr.state_dict().keys()) == ['_model._input_embedding.weight', '_model._output_embedding.weight']


def test_skipgram_forward_and_loss_modules_match_oracle_with_autograd():
    dev = cuda_device()
    from oracle import sgns_oracle
    z = np.load(os.path.join(GOLDEN, 'sgns_e48.npz'))
    vocab, emb = z['w_in_f32'].shape
    model = SkipGram(vocab_size=vocab, embedding_size=emb)
    with torch.no_grad():
        model._input_embedding.weight.copy_(torch.from_numpy(z['w_in_f32']))
        model._output_embedding.weight.copy_(torch.from_numpy(z['w_out_f32']))
    inputs, targets, noise = (torch.from_numpy(z[k]).to(dev) for k in ('inputs', 'targets', 'noise'))
    b, n, k = noise.shape
    pos = model(inputs, targets, proba=False)                           # two forward passes + loss module, as trainer.py:135-139
    neg = model(inputs, noise.view(b, -1), proba=False).view(b, n, k)
    o = sgns_oracle.training_step(z['w_in_f64'], z['w_out_f64'], z['inputs'], z['targets'], z['noise'])
    np.testing.assert_allclose(pos.detach().cpu().numpy(), o['pos_logits'], rtol=1e-4, atol=1e-5)
    loss = NegativeSamplingLoss()(pos, neg)
    np.testing.assert_allclose([float(loss['loss']), float(loss['positive-loss']), float(loss['negative-loss'])], z['loss_f32'], rtol=1e-5)
    loss['loss'].backward()
    den = max(np.abs(z['grad_in_f32']).max(), np.abs(z['grad_out_f32']).max())
    assert np.abs(model._input_embedding.weight.grad.cpu().numpy() - z['grad_in_f32']).max() / den <= 1e-5
    assert np.abs(model._output_embedding.weight.grad.cpu().numpy() - z['grad_out_f32']).max() / den <= 1e-5
    with torch.no_grad():
        proba = model(inputs, targets)                                   # proba=True default
    np.testing.assert_allclose(proba.cpu().numpy(), sgns_oracle.sigmoid(o['pos_logits']), rtol=1e-4, atol=1e-6)
    assert not model.input_embedding.is_cuda and model.input_embedding.shape == (vocab, emb)
    assert SkipGram(vocab_size=10, embedding_size=4, max_norm=1.0).max_norm == 1.0     # renormalisation itself: tests/test_gpu_text.py
    with pytest.raises(NotImplementedError):
        SkipGram(vocab_size=10, embedding_size=4, max_norm=1.0, shard={'world': 2, 'rank': 0})


def test_generate_noise_batch_is_uniform_like_the_reference():
    from scipy.stats import chisquare
    cuda_device()
    noise = generate_noise_batch(384, 4, 5, 35)
    assert noise.shape == (384, 4, 5) and noise.dtype == torch.int64 and noise.is_cuda
    assert int(noise.min()) == 0 and int(noise.max()) == 34                 # index 0 = '<unk>' is drawn too
    big = generate_noise_batch(1000, 10, 50, 35).cpu().numpy().ravel()
    assert chisquare(np.bincount(big, minlength=35)).pvalue > 1e-6
    assert not torch.equal(generate_noise_batch(8, 2, 2, 1000), generate_noise_batch(8, 2, 2, 1000))
    assert torch.equal(generate_noise_batch(8, 2, 2, 1000, seed=5), generate_noise_batch(8, 2, 2, 1000, seed=5))


@pytest.mark.parametrize('engine', ['reference', 'fused'])
def test_train_tool_karate_end_to_end(engine, tmp_path):
    """tools/train.py on the shipped karate YAML (shortened): loss falls, last.ckpt has the reference's keys and the
    embeddings separate the two factions (README: 98 % node classification)."""
    cuda_device()
    import sys
    from conftest import PKG
    sys.path.insert(0, PKG)
    from tools.train import train
    from tools.downstream import node_classification
    from shallow_encoders.config_parser import load_config
    from shallow_encoders.config_parser.core import instantiate
    torch.manual_seed(0)
    over = [f'path.output_dir={tmp_path}', f'train.engine={engine}']
    over += ['train.max_epochs=12', 'train.scheduler.step_size=5'] if engine == 'reference' else \
            ['train.max_epochs=40', 'train.fused_lr=40.0', 'train.scheduler.step_size=15', 'model.embedding_size=8']
    cfg = load_config('sge_sg_karate_club', over)
    trainer, dataset = train(cfg, quiet=True)
    losses = trainer.logged['train-epoch/loss']
    assert losses[-1] < losses[0] - 0.05, losses
    ckpt = torch.load(os.path.join(tmp_path, 'graph_karate_club', 'SG_exp01_baseline', 'checkpoints', 'last.ckpt'))
    assert sorted(ckpt['state_dict']) == ['_model._input_embedding.weight', '_model._output_embedding.weight']
    nc = cfg.downstream['node_classification']
    mean_acc, best = node_classification(trainer.model.input_embedding.numpy(), dataset.vocab.get_itos(), dataset.labels,
                                         instantiate(nc['split_algorithm']), 20)
    assert mean_acc >= 0.85, (engine, mean_acc, best)


def test_powerlaw_synthetic_dataset_trains_through_the_cli_config(tmp_path):
    """`graph_powerlaw_synthetic` (registered like the reference's graphs) + configs/sge_sg_powerlaw_synthetic.yaml at a small size:
    the CSR-only dataset feeds the fused engine, the loss falls, the checkpoint has the reference's keys."""
    from shallow_encoders.config_parser import load_config
    from shallow_encoders.word2vec.dataloader.registry import DATASET_REGISTRY
    from tools.train import train
    cuda_device()
    assert 'graph_powerlaw_synthetic' in DATASET_REGISTRY
    over = ['datamodule.additional_parameters.n_nodes=20000', 'datamodule.additional_parameters.n_edges=200000', 'datamodule.batch_size=16384',
            'datamodule.additional_parameters.walks_per_node=4', 'datamodule.additional_parameters.walk_length=40', 'train.max_epochs=3',
            f'train.fused_lr={0.025 * 16384 * 30 * 10}', f'path.output_dir={tmp_path}']
    cfg = load_config('sge_sg_powerlaw_synthetic', over)
    trainer, ds = train(cfg, quiet=True)
    assert len(ds) == 80000 and len(ds.vocab) == 20001 and ds.vocab.get_itos()[1] == 'n0000000'
    losses = trainer.logged['train-epoch/loss']
    assert len(losses) == 3 and losses[-1] < losses[0] - 0.3, losses
    ckpt = torch.load(os.path.join(str(tmp_path), 'graph_powerlaw_synthetic', cfg.train.experiment, 'checkpoints', 'last.ckpt'), map_location='cpu')
    assert ckpt['state_dict']['_model._input_embedding.weight'].shape == (20001, 128)

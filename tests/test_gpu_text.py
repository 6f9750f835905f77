"""Text path on the device: CBOW scoring / training step against fixtures from the reference's CBOW
(shallow_encoders/word2vec/model.py:94-110 + loss.py + autograd), nn.Embedding(max_norm) renormalisation, the `abcde` corpus
through tools/train.py (fused Adam engine and CBOW reference engine) and the closest-pairs analysis on the tensor cores."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, cuda_device
from shallow_encoders import _native as nat

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(GOLDEN, 'cbow.npz'), allow_pickle=True)


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize('tag', ['abcde', 'e48', 'e100'])
def test_cbow_scores_loss_and_gradients_match_the_reference(tag):
    dev = cuda_device()
    w_in, w_out = _t(Z[f'{tag}_w_in_f32'], dev), _t(Z[f'{tag}_w_out_f32'], dev)
    inputs, targets, noise = _t(Z[f'{tag}_inputs'], dev), _t(Z[f'{tag}_targets'], dev), _t(Z[f'{tag}_noise'], dev)
    b, k = targets.shape[0], noise.shape[2]
    pos = nat.cbow_scores(w_in, w_out, inputs, targets, proba=False)
    neg = nat.cbow_scores(w_in, w_out, inputs, noise.view(b, -1).contiguous(), proba=False)
    scale = max(1.0, float(np.abs(Z[f'{tag}_neg_f64']).max()))
    np.testing.assert_allclose(pos.cpu().numpy(), Z[f'{tag}_pos_f64'], rtol=0, atol=2e-6 * scale)
    np.testing.assert_allclose(neg.cpu().numpy().reshape(b, 1, k), Z[f'{tag}_neg_f64'], rtol=0, atol=2e-6 * scale)
    np.testing.assert_allclose(nat.cbow_scores(w_in, w_out, inputs, targets, proba=True).cpu().numpy(), Z[f'{tag}_proba_f64'], rtol=0, atol=2e-6)
    res = nat.cbow_grad(w_in, w_out, inputs, targets, noise)
    want = Z[f'{tag}_loss_f64']
    got = np.array([res['loss'], res['positive-loss'], res['negative-loss']])
    np.testing.assert_allclose(got, want, rtol=1e-5)
    den = max(np.abs(Z[f'{tag}_grad_in_f64']).max(), np.abs(Z[f'{tag}_grad_out_f64']).max())
    assert np.abs(res['grad_in'].cpu().numpy() - Z[f'{tag}_grad_in_f64']).max() / den < 1e-5          # north-star tolerance: 1e-5 relative
    assert np.abs(res['grad_out'].cpu().numpy() - Z[f'{tag}_grad_out_f64']).max() / den < 1e-5
    assert res['pairs'] == b and res['negatives'] == b * k


def test_cbow_model_and_trainer_follow_the_reference_interface():
    """CBOW(vocab, E).forward + autograd, and Word2VecTrainer.training_step on a cbow batch, give the fixture's gradients."""
    dev = cuda_device()
    from shallow_encoders.word2vec.loss import NegativeSamplingLoss
    from shallow_encoders.word2vec.model import CBOW
    tag = 'e48'
    vocab, emb = Z[f'{tag}_w_in_f32'].shape
    model = CBOW(vocab_size=vocab, embedding_size=emb)
    with torch.no_grad():
        model._input_embedding.weight.copy_(_t(Z[f'{tag}_w_in_f32'], dev)); model._output_embedding.weight.copy_(_t(Z[f'{tag}_w_out_f32'], dev))
    inputs, targets, noise = _t(Z[f'{tag}_inputs'], dev), _t(Z[f'{tag}_targets'], dev), _t(Z[f'{tag}_noise'], dev)
    b = targets.shape[0]
    pos = model(inputs, targets, proba=False)
    neg = model(inputs, noise.view(b, -1), proba=False).view(b, 1, -1)
    loss = NegativeSamplingLoss()(pos, neg)
    loss['loss'].backward()
    den = max(np.abs(Z[f'{tag}_grad_in_f64']).max(), np.abs(Z[f'{tag}_grad_out_f64']).max())
    assert abs(float(loss['loss']) - Z[f'{tag}_loss_f64'][0]) < 1e-5 * Z[f'{tag}_loss_f64'][0]
    assert np.abs(model._input_embedding.weight.grad.cpu().numpy() - Z[f'{tag}_grad_in_f64']).max() / den < 1e-5
    assert np.abs(model._output_embedding.weight.grad.cpu().numpy() - Z[f'{tag}_grad_out_f64']).max() / den < 1e-5
    assert model.input_embedding.device.type == 'cpu' and model.input_embedding.shape == (vocab, emb)


def test_max_norm_renormalises_looked_up_rows_like_nn_embedding():
    dev = cuda_device()
    rng = np.random.default_rng(4)
    w = (rng.standard_normal((50, 6)) * rng.uniform(0.1, 2.0, (50, 1))).astype(np.float32)
    ids = torch.tensor([3, 7, 7, 11, 3, 49, 0])
    ref = torch.nn.Embedding(50, 6, max_norm=1.0)
    with torch.no_grad():
        ref.weight.copy_(torch.from_numpy(w))
    ref(ids)                                                                # look-up renormalises rows 0, 3, 7, 11, 49 in place
    t = _t(w, dev)
    nat.table_renorm_rows(t, ids.to(dev), 1.0)
    np.testing.assert_allclose(t.cpu().numpy(), ref.weight.detach().numpy(), rtol=1e-6, atol=1e-7)
    untouched = np.setdiff1d(np.arange(50), ids.numpy())
    assert np.array_equal(t.cpu().numpy()[untouched], w[untouched])
    assert (np.linalg.norm(t.cpu().numpy()[ids.numpy()], axis=1) <= 1.0 + 1e-6).all()


def test_abcde_trains_with_the_fused_adam_engine_and_closest_pairs_separate_the_groups(tmp_path):
    """configs/w2v_sg_abcde.yaml as shipped (SkipGram E=2, max_norm 1, Adam 0.1, 20 epochs): after training, the closest OUTPUT
    embedding to `a` is `b`'s and to `c` is `d`'s (what the reference's README shows for this toy corpus), computed by the
    tensor-core cosine kernel; every row respects max_norm."""
    cuda_device()
    from shallow_encoders.config_parser import load_config
    from tools.model_analysis import closest_pairs, show_closest_pairs_for_each_word
    from tools.train import train
    hits = 0
    for seed in range(3):
        torch.manual_seed(seed)
        cfg = load_config('w2v_sg_abcde', [f'path.output_dir={tmp_path}/s{seed}', 'train.max_epochs=40', 'train.scheduler.step_size=20'])
        trainer, ds = train(cfg, quiet=True)
        losses = trainer.logged['train-epoch/loss']
        assert losses[-1] < losses[0]
        # rows are renormalised when looked up, BEFORE the optimizer step (nn.Embedding(max_norm) semantics): after the last step a row
        # may exceed max_norm by at most that step
        assert float(trainer.model.input_embedding.norm(dim=1).max()) <= 1.0 + 0.1
        pairs = dict(closest_pairs(trainer.model, ds, max_words=100, pairs_per_word=3))
        assert set(pairs) == {'<unk>', 'a', 'b', 'c', 'd', 'e'} and all(len(v) == 3 for v in pairs.values())
        hits += (pairs['a'][0] == 'b') + (pairs['b'][0] == 'a') + (pairs['c'][0] == 'd') + (pairs['d'][0] == 'c')
        path = show_closest_pairs_for_each_word(trainer.model, ds, str(tmp_path / f'analysis{seed}'), 100, 3)
        text = open(path).read().splitlines()
        assert text[0] == 'Closest pairs in format "{word}:{closest_word_pairs}"' and len(text) == 7 and text[1].startswith('<unk>: ')
    assert hits >= 9, hits                                                  # 12 possible; E = 2 with one negative leaves some seeds ambiguous


def test_cbow_abcde_trains_with_the_reference_engine(tmp_path):
    cuda_device()
    from shallow_encoders.config_parser import load_config
    from shallow_encoders.word2vec.model import CBOW
    from tools.train import train
    cfg = load_config('w2v_cbow_abcde', [f'path.output_dir={tmp_path}', 'train.max_epochs=15'])
    trainer, _ = train(cfg, quiet=True)
    assert isinstance(trainer.model, CBOW)
    losses = trainer.logged['train-epoch/loss']
    assert len(losses) == 15 and losses[-1] < losses[0] - 0.05

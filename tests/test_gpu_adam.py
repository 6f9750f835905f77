"""Row-sparse Adam (csrc/adam.cu, se_sgns_adam_step): the optimizer every shipped YAML names, restricted to the rows a batch
touches.  Pinned three ways: equal to dense torch.optim.Adam on the reference's loss when every step touches the same rows;
equal to the fp64 lazy-Adam oracle on changing batches; and the fused engine of tools/train.py honours
`_target_: torch.optim.Adam` from the unmodified karate YAML."""
import os

import numpy as np
import pytest
import torch

from helpers import cuda_device
from oracle import sgns_oracle
from shallow_encoders import _native as nat

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _assert_adam_close(got, want, lr, steps):
    """Adam divides by sqrt(v) + eps: where a gradient element is ~eps (1e-8) the step is lr * g / (|g| + eps), so fp32 rounding of the
    accumulated gradient (atomics here, autograd's order there) moves that element by up to a few % of lr per step.  That concerns
    well under 1 % of the elements; all others must agree to fp32 accuracy."""
    err = np.abs(got - want)
    tight = err <= 2e-6 + 1e-4 * np.abs(want)
    assert tight.mean() > 0.99, tight.mean()
    assert err.max() < 0.02 * lr * steps, err.max()


def _reference_loss(w_in, w_out, inputs, targets, noise):
    """shallow_encoders/word2vec/model.py:83-91 + loss.py:15-19 in plain torch (the reference arithmetic)."""
    b, n = targets.shape
    c = w_in[inputs.view(-1)].view(b, -1, 1)
    pos = torch.bmm(w_out[targets], c).view(b, n)
    neg = torch.bmm(w_out[noise.view(b, -1)], c).view(b, n, -1)
    pl = -torch.log(torch.clamp(torch.sigmoid(pos), min=1e-6))
    nl = -torch.log(torch.clamp(torch.sigmoid(-neg), min=1e-6)).sum(-1)
    return torch.mean(pl + nl)


@pytest.mark.parametrize('vocab,emb,b,n,k,lr', [(35, 2, 384, 4, 1, 0.1), (301, 128, 96, 4, 5, 0.1), (500, 48, 40, 10, 3, 1e-3), (97, 100, 16, 6, 2, 0.01)])
def test_fixed_batch_equals_dense_torch_adam(vocab, emb, b, n, k, lr):
    """The same batch five times: every step touches the same rows, so the row-wise step counts equal torch's global step and the
    result must equal torch.optim.Adam on dense tables (autograd on the reference's loss), rows untouched by the batch included."""
    dev = cuda_device()
    rng = np.random.default_rng(vocab + emb)
    w_in = (rng.standard_normal((vocab, emb)) * 0.5).astype(np.float32)
    w_out = (rng.standard_normal((vocab, emb)) * 0.5).astype(np.float32)
    inputs = rng.integers(0, vocab, (b, 1)); targets = rng.integers(0, vocab, (b, n)); noise = rng.integers(0, vocab, (b, n, k))
    ti, to = torch.nn.Parameter(torch.from_numpy(w_in.copy())), torch.nn.Parameter(torch.from_numpy(w_out.copy()))
    opt = torch.optim.Adam([ti, to], lr=lr)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = _reference_loss(ti, to, torch.from_numpy(inputs), torch.from_numpy(targets), torch.from_numpy(noise))
        loss.backward()
        opt.step()
        losses.append(float(loss))
    g_in, g_out = _t(w_in, dev), _t(w_out, dev)
    state = nat.AdamState(vocab, emb, dev)
    got_losses = []
    for _ in range(5):
        st = nat.sgns_adam_step(g_in, g_out, _t(inputs, dev), _t(targets, dev), _t(noise, dev), state, lr)
        got_losses.append(st['loss'])
        # the library hands the state back clean: accumulators, flags and list lengths are zero again
        assert float(state.g_in.abs().max()) == 0.0 and float(state.g_out.abs().max()) == 0.0
        assert int(state.touched_in.sum()) == 0 and int(state.touched_out.sum()) == 0 and state.counts.tolist() == [0, 0, 0, 0]
    np.testing.assert_allclose(got_losses, losses, rtol=2e-5)
    _assert_adam_close(g_in.cpu().numpy(), ti.detach().numpy(), lr, 5)
    _assert_adam_close(g_out.cpu().numpy(), to.detach().numpy(), lr, 5)
    rows_in = np.unique(inputs); rows_out = np.unique(np.concatenate([targets.ravel(), noise.ravel()]))
    assert state.t_in.cpu().numpy()[rows_in].tolist() == [5] * len(rows_in) and int(state.t_in.sum()) == 5 * len(rows_in)
    assert int(state.t_out.sum()) == 5 * len(rows_out)
    untouched = np.setdiff1d(np.arange(vocab), rows_out)
    assert np.array_equal(g_out.cpu().numpy()[untouched], w_out[untouched])


def test_changing_batches_equal_the_lazy_adam_oracle():
    """Different rows every step: each row's moments and bias correction follow ITS OWN step count (oracle: fp64 lazy Adam)."""
    dev = cuda_device()
    rng = np.random.default_rng(12)
    vocab, emb, b, n, k, lr = 4000, 64, 64, 4, 3, 0.05
    w_in = (rng.standard_normal((vocab, emb)) * 0.4).astype(np.float32)
    w_out = (rng.standard_normal((vocab, emb)) * 0.4).astype(np.float32)
    o_in, o_out = w_in.astype(np.float64), w_out.astype(np.float64)
    ostate = {key: np.zeros((vocab, emb)) for key in ('m_in', 'v_in', 'm_out', 'v_out')}
    ostate.update(t_in=np.zeros(vocab, dtype=np.int64), t_out=np.zeros(vocab, dtype=np.int64))
    g_in, g_out = _t(w_in, dev), _t(w_out, dev)
    state = nat.AdamState(vocab, emb, dev)
    for step in range(6):
        hi = 300 if step % 2 else vocab                               # alternate a crowded and a sparse batch: rows recur at different rates
        inputs = rng.integers(0, hi, (b, 1)); targets = rng.integers(0, hi, (b, n)); noise = rng.integers(0, vocab, (b, n, k))
        o = sgns_oracle.lazy_adam_step(o_in, o_out, ostate, inputs, targets, noise, lr)
        st = nat.sgns_adam_step(g_in, g_out, _t(inputs, dev), _t(targets, dev), _t(noise, dev), state, lr)
        assert abs(st['loss'] - o['loss']) < 1e-4 * o['loss']
    _assert_adam_close(g_in.cpu().numpy(), o_in, lr, 6)
    _assert_adam_close(g_out.cpu().numpy(), o_out, lr, 6)
    assert np.array_equal(state.t_in.cpu().numpy(), ostate['t_in']) and np.array_equal(state.t_out.cpu().numpy(), ostate['t_out'])
    assert ostate['t_out'].max() > 1 and (ostate['t_out'] == 1).any()
    np.testing.assert_allclose(state.m_out.cpu().numpy(), ostate['m_out'], rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(state.v_out.cpu().numpy(), ostate['v_out'], rtol=4e-3, atol=1e-10)


def test_fused_engine_honours_the_yamls_adam(tmp_path):
    """`tools/train.py --config-name=sge_sg_karate_club train.engine=fused` (no fused_lr): the YAML's torch.optim.Adam (lr 0.1) and
    StepLR(10, 0.1) drive the row-sparse Adam kernel; loss falls like the reference engine's and the checkpoint keeps the
    reference's state-dict keys."""
    cuda_device()
    from shallow_encoders.config_parser import load_config
    from shallow_encoders.word2vec.optim import RowSparseAdam
    from tools.train import train
    over = ['train.engine=fused', 'train.max_epochs=12', f'path.output_dir={tmp_path}']
    cfg = load_config('sge_sg_karate_club', over)
    assert cfg.train.fused_optimizer_kind() == 'adam'
    trainer, dataset = train(cfg, quiet=True)
    assert isinstance(trainer.optimizer, RowSparseAdam)
    losses = trainer.logged['train-epoch/loss']
    assert len(losses) == 12 and losses[-1] < losses[0] - 0.05 and np.isfinite(losses).all()
    lrs = trainer.logged['epoch/lr']
    assert abs(lrs[0] - 0.1) < 1e-12 and abs(lrs[10] - 0.01) < 1e-12            # StepLR(step_size=10, gamma=0.1) from the YAML
    ckpt = torch.load(os.path.join(str(tmp_path), cfg.datamodule.dataset_name, cfg.train.experiment, 'checkpoints', 'last.ckpt'), map_location='cpu')
    assert set(ckpt['state_dict']) == {'_model._input_embedding.weight', '_model._output_embedding.weight'}
    # the reference engine on the same config reaches a comparable loss (same loss definition, dense Adam)
    cfg_ref = load_config('sge_sg_karate_club', ['train.engine=reference', 'train.max_epochs=12', f'path.output_dir={tmp_path}/ref'])
    ref_trainer, _ = train(cfg_ref, quiet=True)
    assert abs(ref_trainer.logged['train-epoch/loss'][-1] - losses[-1]) < 0.08

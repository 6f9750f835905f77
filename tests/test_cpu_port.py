"""CPU: the timed CPU arms are what they claim to be.

* oracle/cpu_port.py (kind "port", the fallback when the reference cannot be imported) is pinned to the reference's golden
  vectors: its SGNS step reproduces the golden loss triple and dense gradients, its collate the golden windows, its walk
  loop the golden walks under replayed uniforms.
* oracle/ref_pipeline.py (kind "reference") really drives the unmodified reference: every module on its path is imported
  from the reference root (the mount, or the byte-for-byte copy under baseline/_ref/), and the vendored copy is identical
  to the mount.
"""
import glob
import os
import random

import numpy as np
import pytest
import torch

from oracle import cpu_port, ref_import, ref_pipeline, vendor_ref, walk_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SGNS_CASES = sorted(os.path.basename(p)[len('sgns_'):-len('.npz')] for p in glob.glob(os.path.join(GOLDEN, 'sgns_*.npz')))


@pytest.mark.parametrize('tag', SGNS_CASES)
def test_port_sgns_step_matches_reference_golden(tag):
    z = np.load(os.path.join(GOLDEN, f'sgns_{tag}.npz'))
    vocab, emb = z['w_in_f32'].shape
    m = cpu_port.TorchCpuSgns(vocab, emb, z['noise'].shape[2])
    with torch.no_grad():
        m.w_in.weight.copy_(torch.from_numpy(z['w_in_f32']))
        m.w_out.weight.copy_(torch.from_numpy(z['w_out_f32']))
    out = m.loss(torch.from_numpy(z['inputs']), torch.from_numpy(z['targets']), torch.from_numpy(z['noise']))
    out['loss'].backward()
    got = np.array([float(out['loss']), float(out['positive-loss']), float(out['negative-loss'])])
    np.testing.assert_allclose(got, z['loss_f32'], rtol=1e-5)
    den = max(np.abs(z['grad_in_f32']).max(), np.abs(z['grad_out_f32']).max())
    assert np.abs(m.w_in.weight.grad.numpy() - z['grad_in_f32']).max() / den < 1e-5
    assert np.abs(m.w_out.weight.grad.numpy() - z['grad_out_f32']).max() / den < 1e-5


def test_port_collate_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN, 'collate.npz'))
    for name in ('karate', 'clip', 'tri'):
        inputs, targets = cpu_port.collate(z[f'{name}_texts'], int(z[f'{name}_r']), int(z[f'{name}_max_length']), 0)
        assert np.array_equal(inputs.numpy(), z[f'{name}_inputs']) and np.array_equal(targets.numpy(), z[f'{name}_targets'])


class _Replay(random.Random):
    def __init__(self, draws):
        super().__init__(0)
        self._it = iter(draws)

    def random(self):
        return next(self._it)


@pytest.mark.parametrize('tag', ['karate_yaml', 'gnm_unweighted', 'triplets_node2vec'])
def test_port_walk_loop_matches_reference_golden(tag, monkeypatch):
    z = np.load(os.path.join(GOLDEN, f'walks_{tag}.npz'))
    rowptr, col = z['rowptr'], z['col']
    adj = [list(map(int, col[rowptr[i]:rowptr[i + 1]])) for i in range(len(rowptr) - 1)]
    wts = None
    if bool(z['weighted']):
        conv = int if bool(z['w_is_int']) else float
        wts = [[conv(x) for x in z['w'][rowptr[i]:rowptr[i + 1]]] for i in range(len(rowptr) - 1)]
    g = walk_oracle.OracleGraph(adj, wts, [str(s) for s in z['names']])
    cpu_port._G['g'] = g
    monkeypatch.setattr(cpu_port.random, 'Random', lambda seed: _Replay(z['uniforms'].reshape(-1)))
    got = cpu_port._walk_chunk((list(map(int, z['starts'])), int(z['length']), float(z['p']), float(z['q']), bool(z['node2vec']), 0))
    assert np.array_equal(np.array(got), z['walks'])


needs_reference = pytest.mark.skipif(ref_import.reference_root() is None, reason='reference neither mounted nor vendored')


@needs_reference
def test_reference_pipeline_runs_the_unmodified_reference():
    r = ref_pipeline.call(graph='karate', method='node2vec', p=1.0, q=0.5, walk_len=10, radius=2, emb=2, neg=1, steps=2, warmup=1,
                          walks_per_step=64, workers=2, timeout=300)
    assert r.get('kind') == 'reference', r
    assert r['reference_root'] == ref_import.reference_root()
    assert r['pairs'] == 2 * 64 * 6 * 4 and r['walk_steps'] == 2 * 64 * 9          # (L - 2r) centres x 2r contexts; L - 1 transitions
    assert r['vocab'] == 35 and np.isfinite(r['loss'])


@needs_reference
def test_vendored_copy_is_byte_identical_to_the_mount():
    if not os.path.isdir('/root/reference/shallow_encoders'):
        pytest.skip('no mount to compare with (GPU box)')
    assert vendor_ref.vendor() > 0
    assert vendor_ref.verify()
    assert os.path.exists(os.path.join(vendor_ref.DST, 'shallow_encoders', 'graph', 'random_walk_generator.py'))

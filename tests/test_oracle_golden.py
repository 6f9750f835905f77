"""CPU: the oracle restatements (python + C) reproduce every golden vector generated from the
unmodified reference by oracle/make_golden.py."""
import glob
import os
import sys

import numpy as np
import pytest

from oracle import sgns_oracle, walk_oracle
from oracle.c_oracle import c_walks

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
WALK_CASES = sorted(os.path.basename(p)[len('walks_'):-len('.npz')] for p in glob.glob(os.path.join(GOLDEN, 'walks_*.npz')))
SGNS_CASES = sorted(os.path.basename(p)[len('sgns_'):-len('.npz')] for p in glob.glob(os.path.join(GOLDEN, 'sgns_*.npz')))


def load_walk_case(tag):
    z = np.load(os.path.join(GOLDEN, f'walks_{tag}.npz'))
    rowptr, col = z['rowptr'], z['col']
    w = z['w'] if bool(z['weighted']) else None
    adj = [list(map(int, col[rowptr[i]:rowptr[i + 1]])) for i in range(len(rowptr) - 1)]
    wts = None
    if w is not None:
        conv = (lambda x: int(x)) if bool(z['w_is_int']) else (lambda x: float(x))
        wts = [[conv(x) for x in w[rowptr[i]:rowptr[i + 1]]] for i in range(len(rowptr) - 1)]
    g = walk_oracle.OracleGraph(adj, wts, [str(s) for s in z['names']])
    return z, g, w


def test_fixture_inventory():
    assert len(WALK_CASES) >= 10 and len(SGNS_CASES) >= 5


@pytest.mark.parametrize('tag', WALK_CASES)
def test_python_walk_oracle_matches_reference(tag):
    z, g, _ = load_walk_case(tag)
    got = walk_oracle.walks(g, z['starts'], int(z['length']), z['uniforms'], float(z['p']), float(z['q']),
                            node2vec=bool(z['node2vec']))
    assert np.array_equal(got, z['walks'])


@pytest.mark.parametrize('tag', WALK_CASES)
def test_c_walk_oracle_matches_reference(tag):
    z, g, w = load_walk_case(tag)
    for sorted_membership in (False, True):
        col_sorted = None
        if sorted_membership:
            col_sorted = np.concatenate([np.sort(z['col'][z['rowptr'][i]:z['rowptr'][i + 1]]) for i in range(g.n_nodes)])
        got = c_walks(z['rowptr'], z['col'], w, bool(z['w_is_int']), z['starts'], int(z['length']), float(z['p']),
                      float(z['q']), bool(z['node2vec']), 0, z['uniforms'], col_sorted=col_sorted)
        assert np.array_equal(got, z['walks'])


@pytest.mark.skipif(sys.version_info < (3, 12), reason='builtin sum() is Neumaier-compensated from 3.12')
def test_py312_sum_restatement_equals_builtin():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        n = int(rng.integers(1, 40))
        vals = []
        for _ in range(n):
            if rng.random() < 0.5:
                vals.append(int(rng.integers(1, 9)))
            else:
                vals.append(float(rng.random() * 10.0 ** int(rng.integers(-8, 8))))
        a, b = walk_oracle.py312_sum(vals), sum(vals)
        assert type(a) is type(b) and a == b


def test_paper_rule_differs_from_code_rule():
    # SURVEY trap #1: 4-node graph, p=4, q=.25 -> code rule favours the common neighbour
    adj = [[1, 2], [0, 2, 3], [0, 1], [1]]           # t=0, v=1: x=0 (return), x=2 (common nbr), x=3 (distance 2)
    g = walk_oracle.OracleGraph(adj)
    ref = walk_oracle.transition_probabilities(g, 0, 1, 4.0, 0.25, True, walk_oracle.RULE_REFERENCE)
    pap = walk_oracle.transition_probabilities(g, 0, 1, 4.0, 0.25, True, walk_oracle.RULE_PAPER)
    np.testing.assert_allclose(ref, np.array([0.25, 4.0, 1.0]) / 5.25)
    np.testing.assert_allclose(pap, np.array([0.25, 1.0, 4.0]) / 5.25)


def test_collate_golden():
    z = np.load(os.path.join(GOLDEN, 'collate.npz'))
    i, t = sgns_oracle.collate_sg([np.arange(10, 18)], 3, 256)
    assert np.array_equal(i, z['worked_inputs']) and np.array_equal(t, z['worked_targets'])
    for tag in ('karate', 'clip', 'tri'):
        i, t = sgns_oracle.collate_sg(list(z[f'{tag}_texts']), int(z[f'{tag}_r']), int(z[f'{tag}_max_length']))
        assert np.array_equal(i, z[f'{tag}_inputs']) and np.array_equal(t, z[f'{tag}_targets'])
    with pytest.raises(AssertionError):
        sgns_oracle.collate_sg([np.arange(4)], 2, 256)   # shorter than 2r+1 (torch_dataset.py:298)


@pytest.mark.parametrize('tag', SGNS_CASES)
@pytest.mark.parametrize('prec', ['f32', 'f64'])
def test_sgns_oracle_matches_reference_autograd(tag, prec):
    z = np.load(os.path.join(GOLDEN, f'sgns_{tag}.npz'))
    o = sgns_oracle.training_step(z[f'w_in_{prec}'], z[f'w_out_{prec}'], z['inputs'], z['targets'], z['noise'])
    tol = 1e-12 if prec == 'f64' else 1e-5
    got = np.array([o['loss'], o['positive-loss'], o['negative-loss']], dtype=np.float64)
    np.testing.assert_allclose(got, z[f'loss_{prec}'], rtol=tol)
    den = max(np.abs(z[f'grad_in_{prec}']).max(), np.abs(z[f'grad_out_{prec}']).max())
    assert np.abs(o['grad_in'] - z[f'grad_in_{prec}']).max() / den <= tol
    assert np.abs(o['grad_out'] - z[f'grad_out_{prec}']).max() / den <= tol
    np.testing.assert_allclose([o['recall'], o['precision']], z[f'metrics_{prec}'], atol=1e-6)


def test_vocab_golden_order():
    z = np.load(os.path.join(GOLDEN, 'vocab.npz'))
    itos = [str(s) for s in z['karate_itos']]
    assert itos[0] == '<unk>' and itos[1:] == [f'n{i:02d}' for i in range(1, 35)]
    assert [str(s) for s in z['triplets_itos']] == ['<unk>', 'a1', 'a2', 'a3', 'b1', 'b2', 'b3', 'c1', 'c2', 'c3']


@pytest.mark.parametrize('seed,weights,method,p,q', [(11, None, 'node2vec', 0.5, 2.0), (12, 'int', 'node2vec', 4.0, 0.25), (13, 'float', 'node2vec', 1.0, 0.5),
                                                     (14, 'float', 'deepwalk', 1.0, 1.0), (15, 'int', 'dfs', 1.0, 1.0)])
def test_walk_oracles_match_the_live_reference_on_fresh_random_graphs(seed, weights, method, p, q):
    """Beyond the committed fixtures: wherever the reference can be loaded (/root/reference here, baseline/_ref on the GPU box) its
    DeepWalk.walk / Node2Vec.walk run under a replayed uniform stream on a NEW random graph (unsorted adjacency, optional int / float
    weights) and both oracles -- the python restatement and the C one the GPU tests compare against -- must reproduce every walk."""
    import importlib.util
    import random
    import types
    import networkx as nx
    from oracle import ref_import
    root = ref_import.reference_root()
    if not root:
        pytest.skip('reference not available')
    spec = importlib.util.spec_from_file_location('_ref_rwg_live', os.path.join(root, 'shallow_encoders', 'graph', 'random_walk_generator.py'))
    ref = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(ref)
    rnd = random.Random(seed)
    n, m, length = 60, 240, 12
    names = [f'n{i:03d}' for i in range(n)]
    order = list(range(n))
    rnd.shuffle(order)
    edges = {(min(a, b), max(a, b)) for a, b in zip(order, order[1:])}           # a spanning path: no isolated node
    while len(edges) < m:
        a, b = rnd.randrange(n), rnd.randrange(n)
        if a != b:
            edges.add((min(a, b), max(a, b)))
    edges = list(edges)
    rnd.shuffle(edges)                                                            # random insertion order -> unsorted adjacency lists
    g = nx.Graph()
    for a, b in edges:
        a, b = (b, a) if rnd.random() < 0.5 else (a, b)
        if weights == 'float':
            g.add_edge(names[a], names[b], weight=rnd.uniform(0.25, 4.0))
        elif weights == 'int':
            g.add_edge(names[a], names[b], weight=rnd.randint(1, 9))
        else:
            g.add_edge(names[a], names[b])
    og = walk_oracle.OracleGraph.from_networkx(g)
    rng = np.random.default_rng(seed)
    starts = rng.permutation(np.repeat(np.arange(n, dtype=np.int32), 2))
    uniforms = rng.random((len(starts), length - 1))

    class _Replay(random.Random):
        def __init__(self, draws):
            super().__init__(0)
            self.draws, self.i = [float(x) for x in draws], 0

        def random(self):
            self.i += 1
            return self.draws[self.i - 1]
    replay = _Replay(uniforms.reshape(-1))
    ref.random = types.SimpleNamespace(choices=replay.choices)                    # the module's `random.choices` now consumes the recorded stream
    gen = ref.random_walk_factory(method, g, length, {'p': p, 'q': q} if method == 'node2vec' else {})
    idx = {name: i for i, name in enumerate(og.names)}
    want = np.array([[idx[t] for t in gen.walk(og.names[s]).split(' ')] for s in starts], dtype=np.int32)
    assert replay.i == uniforms.size                                              # exactly one draw per transition
    node2vec = method == 'node2vec'
    assert np.array_equal(walk_oracle.walks(og, starts, length, uniforms, p, q, node2vec=node2vec), want)
    rowptr, col, w, w_is_int = og.to_csr()
    assert np.array_equal(c_walks(rowptr, col, w, w_is_int, starts, length, p, q, node2vec, 0, uniforms), want)


@pytest.mark.parametrize('vocab,emb,b,n,k,scale,seed', [(50, 8, 16, 4, 1, 1.0, 1), (300, 48, 32, 10, 3, 0.3, 2), (40, 128, 24, 4, 5, 6.0, 3)])
def test_sgns_oracle_matches_the_live_reference_modules(vocab, emb, b, n, k, scale, seed):
    """Beyond the committed fixtures: the reference's own SkipGram (word2vec/model.py:79-91) and NegativeSamplingLoss (loss.py:14-22), loaded
    from the reference tree where it is available, evaluated in fp64 with torch autograd on a fresh random batch (duplicate rows, and with
    scale 6 logits deep in the clamp region); the numpy oracle the GPU kernels are tested against must give the same loss triple and the same
    dense gradients to 1e-12."""
    import importlib.util
    import torch
    from oracle import ref_import
    root = ref_import.reference_root()
    if not root:
        pytest.skip('reference not available')
    mods = {}
    for name in ('model', 'loss'):
        spec = importlib.util.spec_from_file_location(f'_ref_w2v_{name}_live', os.path.join(root, 'shallow_encoders', 'word2vec', f'{name}.py'))
        mods[name] = importlib.util.module_from_spec(spec)
        sys.dont_write_bytecode = True
        spec.loader.exec_module(mods[name])
    rng = np.random.default_rng(seed)
    w_in = rng.standard_normal((vocab, emb)) * scale / np.sqrt(emb)
    w_out = rng.standard_normal((vocab, emb)) * scale / np.sqrt(emb) * 4
    inputs = rng.integers(0, vocab, (b, 1))
    targets = rng.integers(0, vocab, (b, n))
    noise = rng.integers(0, vocab, (b, n, k))
    model = mods['model'].SkipGram(vocab_size=vocab, embedding_size=emb).double()
    with torch.no_grad():
        model._input_embedding.weight.copy_(torch.from_numpy(w_in))
        model._output_embedding.weight.copy_(torch.from_numpy(w_out))
    t_in, t_tg, t_nz = torch.from_numpy(inputs), torch.from_numpy(targets), torch.from_numpy(noise)
    pos = model(t_in, t_tg, proba=False)                                         # trainer.py:135-139
    neg = model(t_in, t_nz.view(b, -1), proba=False).view(b, n, -1)
    out = mods['loss'].NegativeSamplingLoss()(pos, neg)
    out['loss'].backward()
    o = sgns_oracle.training_step(w_in, w_out, inputs, targets, noise)
    for key in ('loss', 'positive-loss', 'negative-loss'):
        ref_value = float(out[key].detach())
        assert abs(o[key] - ref_value) <= 1e-12 * max(1.0, abs(ref_value)), key
    g_in, g_out = model._input_embedding.weight.grad.numpy(), model._output_embedding.weight.grad.numpy()
    den = max(np.abs(g_in).max(), np.abs(g_out).max())
    assert np.abs(o['grad_in'] - g_in).max() <= 1e-12 * den and np.abs(o['grad_out'] - g_out).max() <= 1e-12 * den
    if scale > 5:
        assert (1.0 / (1.0 + np.exp(-o['neg_logits'])) > 1 - 1e-6).any() or (1.0 / (1.0 + np.exp(o['pos_logits'])) > 1 - 1e-6).any()   # clamp region reached


def test_lazy_adam_oracle_equals_dense_torch_adam_while_the_batch_touches_the_same_rows():
    """The row-sparse Adam kernel (csrc/adam.cu) is checked against `sgns_oracle.lazy_adam_step`; the oracle itself is pinned here, on the
    host: five steps on the SAME batch (every step touches the same rows, so per-row step counts equal torch's global step) must equal
    torch.optim.Adam on dense fp64 tables driven by autograd through the reference's loss arithmetic -- untouched rows included
    (zero gradient and zero moments: dense Adam does not move them either)."""
    import torch
    rng = np.random.default_rng(9)
    vocab, emb, b, n, k, lr = 60, 16, 20, 4, 3, 0.1
    w_in = rng.standard_normal((vocab, emb)) * 0.3
    w_out = rng.standard_normal((vocab, emb)) * 0.3
    inputs, targets, noise = rng.integers(0, vocab, (b, 1)), rng.integers(0, vocab, (b, n)), rng.integers(0, vocab, (b, n, k))
    ti, to = torch.nn.Parameter(torch.from_numpy(w_in.copy())), torch.nn.Parameter(torch.from_numpy(w_out.copy()))
    opt = torch.optim.Adam([ti, to], lr=lr)
    state = {key: np.zeros((vocab, emb)) for key in ('m_in', 'v_in', 'm_out', 'v_out')}
    state['t_in'], state['t_out'] = np.zeros(vocab, dtype=np.int64), np.zeros(vocab, dtype=np.int64)
    a, c = w_in.copy(), w_out.copy()
    t_inputs, t_targets, t_noise = torch.from_numpy(inputs), torch.from_numpy(targets), torch.from_numpy(noise)
    for _ in range(5):
        centre = ti[t_inputs[:, 0]]                                               # model.py:83-90 + loss.py:15-19 in plain torch
        pos = torch.einsum('be,bne->bn', centre, to[t_targets])
        neg = torch.einsum('be,bnke->bnk', centre, to[t_noise])
        loss = -(torch.log(torch.clamp(torch.sigmoid(pos), min=1e-6)).mean() + torch.log(torch.clamp(torch.sigmoid(-neg), min=1e-6)).sum(-1).mean())
        opt.zero_grad()
        loss.backward()
        opt.step()
        o = sgns_oracle.lazy_adam_step(a, c, state, inputs, targets, noise, lr)
        assert abs(o['loss'] - float(loss.detach())) < 1e-12
    np.testing.assert_allclose(a, ti.detach().numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(c, to.detach().numpy(), rtol=0, atol=1e-12)
    untouched = np.setdiff1d(np.arange(vocab), inputs)
    assert np.array_equal(a[untouched], w_in[untouched]) and (state['t_in'][np.unique(inputs)] == 5).all()


def test_philox_restatement_passes_the_published_known_answer_vectors():
    """tests/philox_ref.py predicts every in-kernel draw of the GPU tests (walk tries, negatives, negative edges).  It is the published
    Philox4x32-10 (Salmon et al., SC'11): the three known-answer vectors of the Random123 distribution (kat_vectors: counter / key all zero,
    all ones, digits of pi) come out bit for bit, so the kernels -- pinned to this restatement on the GPU -- run a standard generator."""
    import philox_ref
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox_ref.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want
    # the keyed form used by the kernels: (id low, id high, sub, stream) as the counter, the 64-bit seed as the key
    a = philox_ref.philox(0x299f31d0a4093822, np.array([0x85a308d3243f6a88], dtype=np.uint64), 0x13198a2e, 0x03707344)
    assert tuple(int(x[0]) for x in a) == kat[2][2]

"""
TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, build container only) and checks the oracle restatements against it on the way.

    python oracle/make_golden.py            # rewrites tests/golden/

Everything written here is produced by the reference's own code:
  walks_*    Node2Vec.walk / DeepWalk.walk          (graph/random_walk_generator.py:61-72,94-119)
             under a replayed uniform stream (one draw per transition)
  collate_*  W2VCollateFunctional.__call__           (word2vec/dataloader/torch_dataset.py:293-322)
  sgns_*     SkipGram.forward + NegativeSamplingLoss + torch autograd
             (word2vec/model.py:79-91, word2vec/loss.py:14-22) and Word2VecTrainer.training_step
             (word2vec/trainer.py:131-152) with `generate_noise_batch` pinned to a recorded tensor
  vocab_*    GraphDataset vocabulary order           (torch_dataset.py:99-110)
  downstream perform_node_classification / perform_edge_classification of tools/graph_model_downstream_classification.py
             (:94-148, :227-299) run UNMODIFIED (hydra / omegaconf / matplotlib / tools.utils stubbed: they only serve the CLI and
             the plots) on a fixed embedding of the karate-club graph: per-experiment accuracies
  edge_ops   the four edge operators                 (graph/edge_operators.py:10-64) applied the way
             create_edge_embeddings does (tools/graph_model_downstream_classification.py:203-224)

Versions used for the committed fixtures are stored inside each file (`meta`).
"""
import json
import os
import platform
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

ref_import.import_reference()

import networkx as nx  # noqa: E402
import torch  # noqa: E402

import shallow_encoders.graph.random_walk_generator as rwg  # noqa: E402  (the reference's)
from shallow_encoders.graph import datasets as ref_datasets  # noqa: E402
from shallow_encoders.word2vec.dataloader import torch_dataset as ref_td  # noqa: E402
from shallow_encoders.word2vec.loss import NegativeSamplingLoss  # noqa: E402
from shallow_encoders.word2vec.model import SkipGram  # noqa: E402
from shallow_encoders.word2vec import trainer as ref_trainer  # noqa: E402

from oracle import sgns_oracle, walk_oracle  # noqa: E402
from oracle.c_oracle import c_walks  # noqa: E402

assert rwg.__file__.startswith(ref_import.REFERENCE_ROOT), rwg.__file__

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
META = json.dumps({
    'python': platform.python_version(), 'networkx': nx.__version__, 'torch': torch.__version__,
    'numpy': np.__version__, 'reference': 'Robotmurlock/Deepwalk-and-Node2vec @ /root/reference',
})


class Replay(random.Random):
    """random.Random whose .random() replays a recorded float64 stream."""

    def __init__(self, draws):
        super().__init__(0)
        self._draws = [float(x) for x in draws]
        self._i = 0

    def random(self):
        x = self._draws[self._i]
        self._i += 1
        return x

    @property
    def consumed(self):
        return self._i


def reference_walks(graph, names, starts, length, method, params, uniforms):
    gen = rwg.random_walk_factory(method, graph, length, params)
    flat = np.asarray(uniforms, dtype=np.float64).reshape(-1)
    replay = Replay(flat)
    saved = rwg.random
    rwg.random = types.SimpleNamespace(choices=replay.choices)
    try:
        idx = {n: i for i, n in enumerate(names)}
        out = np.empty((len(starts), length), dtype=np.int32)
        for i, s in enumerate(starts):
            toks = gen.walk(names[s]).split(' ')
            out[i] = [idx[t] for t in toks]
    finally:
        rwg.random = saved
    assert replay.consumed == len(starts) * (length - 1), (replay.consumed, len(starts), length)
    return out


def emit_walk_case(tag, graph, length, method, p, q, walks_per_node, seed):
    og = walk_oracle.OracleGraph.from_networkx(graph)
    rowptr, col, w, w_is_int = og.to_csr()
    rng = np.random.default_rng(seed)
    starts = np.repeat(np.arange(og.n_nodes, dtype=np.int32), walks_per_node)
    rng.shuffle(starts)
    uniforms = rng.random((len(starts), length - 1))
    params = {'p': p, 'q': q} if method == 'node2vec' else {}
    ref = reference_walks(graph, og.names, starts, length, method, params, uniforms)
    mine = walk_oracle.walks(og, starts, length, uniforms, p, q, node2vec=(method == 'node2vec'))
    assert np.array_equal(ref, mine), f'{tag}: python oracle != reference'
    cw = c_walks(rowptr, col, w, w_is_int, starts, length, p, q, method == 'node2vec', 0, uniforms)
    assert np.array_equal(ref, cw), f'{tag}: C oracle != reference'
    np.savez_compressed(
        os.path.join(GOLDEN, f'walks_{tag}.npz'),
        rowptr=rowptr, col=col, w=(w if w is not None else np.zeros(0)), weighted=np.array(w is not None),
        w_is_int=np.array(w_is_int), starts=starts, uniforms=uniforms, walks=ref,
        length=np.array(length), p=np.array(float(p)), q=np.array(float(q)),
        node2vec=np.array(method == 'node2vec'), names=np.array(og.names), meta=np.array(META))
    print(f'walks_{tag}: {len(starts)} walks x {length}, nodes={og.n_nodes}, nnz={len(col)}, weighted={w is not None} OK')


def random_graph(n, m, seed, weights=None):
    """G(n, m) with edges inserted in random order -> UNSORTED adjacency lists."""
    rng = random.Random(seed)
    g = nx.Graph()
    names = [f'n{i:05d}' for i in range(n)]
    order = list(range(n))
    rng.shuffle(order)
    # spanning path first so there is no isolated node (reference crashes on degree 0)
    edges = {(min(order[i], order[i + 1]), max(order[i], order[i + 1])) for i in range(n - 1)}
    while len(edges) < m:
        a, b = rng.randrange(n), rng.randrange(n)
        if a != b:
            edges.add((min(a, b), max(a, b)))
    edges = list(edges)
    rng.shuffle(edges)
    for a, b in edges:
        if rng.random() < 0.5:
            a, b = b, a
        if weights == 'float':
            g.add_edge(names[a], names[b], weight=rng.uniform(0.25, 4.0))
        elif weights == 'int':
            g.add_edge(names[a], names[b], weight=rng.randint(1, 9))
        else:
            g.add_edge(names[a], names[b])
    return g


def emit_walks():
    triplets = ref_datasets.GraphTriplets(walks_per_node=1, walk_length=5).graph
    emit_walk_case('triplets_deepwalk', triplets, 5, 'deepwalk', 1, 1, 16, 1)
    emit_walk_case('triplets_node2vec', triplets, 7, 'node2vec', 0.25, 4.0, 16, 2)
    karate = ref_datasets.KarateClubDataset(walks_per_node=1, walk_length=10).graph
    assert nx.is_weighted(karate)
    emit_walk_case('karate_yaml', karate, 10, 'node2vec', 1, 0.5, 8, 3)     # configs/sge_sg_karate_club.yaml:17-22
    emit_walk_case('karate_p05_q2', karate, 10, 'node2vec', 0.5, 2.0, 8, 4)
    emit_walk_case('karate_deepwalk', karate, 10, 'deepwalk', 1, 1, 8, 5)
    emit_walk_case('gnm_unweighted', random_graph(300, 1200, 6), 20, 'node2vec', 0.5, 2.0, 3, 7)
    emit_walk_case('gnm_cora_yaml', random_graph(300, 600, 8), 10, 'node2vec', 1, 2.0, 3, 9)  # sge_sg_cora.yaml p=1 q=2
    emit_walk_case('gnm_floatw', random_graph(200, 900, 10, 'float'), 12, 'node2vec', 2.0, 0.5, 3, 11)
    emit_walk_case('gnm_intw_deepwalk', random_graph(200, 900, 12, 'int'), 12, 'deepwalk', 1, 1, 3, 13)
    # a hub: star + ring, exercises long CDFs (degree 400)
    star = nx.Graph()
    names = [f'n{i:05d}' for i in range(401)]
    perm = list(range(1, 401))
    random.Random(14).shuffle(perm)
    for i in perm:
        star.add_edge(names[0], names[i])
    for i in range(1, 400):
        star.add_edge(names[i], names[i + 1])
    emit_walk_case('star_hub', star, 16, 'node2vec', 4.0, 0.25, 2, 15)


def emit_collate():
    # worked example from the source comment, torch_dataset.py:302-306
    text = torch.arange(10, 18, dtype=torch.long)
    inp, tgt = ref_td.W2VCollateFunctional('sg', 3, 256)([text])
    assert inp.tolist() == [[13], [14]] and tgt.tolist() == [[10, 11, 12, 14, 15, 16], [11, 12, 13, 15, 16, 17]]
    rng = np.random.default_rng(21)
    cases = {}
    for tag, (n, length, r, max_len) in {'karate': (64, 10, 2, 256), 'clip': (5, 40, 5, 32), 'tri': (7, 5, 2, 256)}.items():
        texts = rng.integers(0, 1000, size=(n, length))
        inp, tgt = ref_td.W2VCollateFunctional('sg', r, max_len)([torch.tensor(t, dtype=torch.long) for t in texts])
        oi, ot = sgns_oracle.collate_sg(list(texts), r, max_len)
        assert np.array_equal(inp.numpy(), oi) and np.array_equal(tgt.numpy(), ot), tag
        cases[f'{tag}_texts'] = texts
        cases[f'{tag}_r'] = np.array(r)
        cases[f'{tag}_max_length'] = np.array(max_len)
        cases[f'{tag}_inputs'] = inp.numpy()
        cases[f'{tag}_targets'] = tgt.numpy()
    # ragged batch (different lengths) -- only the reference/oracle handle it; the dense GPU path is per-length
    ragged = [rng.integers(0, 50, size=(ln,)) for ln in (5, 9, 6, 12)]
    inp, tgt = ref_td.W2VCollateFunctional('sg', 2, 10)([torch.tensor(t, dtype=torch.long) for t in ragged])
    oi, ot = sgns_oracle.collate_sg(ragged, 2, 10)
    assert np.array_equal(inp.numpy(), oi) and np.array_equal(tgt.numpy(), ot)
    cases['worked_inputs'] = np.array([[13], [14]])
    cases['worked_targets'] = np.array([[10, 11, 12, 14, 15, 16], [11, 12, 13, 15, 16, 17]])
    np.savez_compressed(os.path.join(GOLDEN, 'collate.npz'), meta=np.array(META), **cases)
    print('collate OK')


def emit_sgns_case(tag, vocab, emb, b, n, k, seed, scale=1.0):
    torch.manual_seed(seed)
    out = {}
    model = SkipGram(vocab_size=vocab, embedding_size=emb)
    with torch.no_grad():
        model._input_embedding.weight.mul_(scale)
        model._output_embedding.weight.mul_(scale)
    g = torch.Generator().manual_seed(seed + 1)
    inputs = torch.randint(0, vocab, (b, 1), generator=g)
    targets = torch.randint(0, vocab, (b, n), generator=g)
    noise = torch.randint(0, vocab, (b, n, k), generator=g)

    for dt, name in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        m = SkipGram(vocab_size=vocab, embedding_size=emb).to(dt)
        m.load_state_dict({kk: v.to(dt) for kk, v in model.state_dict().items()})
        # the real training_step with the noise pinned (trainer.py:133)
        saved = ref_trainer.generate_noise_batch
        ref_trainer.generate_noise_batch = lambda *_a, **_k: noise.clone()
        try:
            tr = ref_trainer.Word2VecTrainer(m, optimizer=None, scheduler=None, neg_samples=k, vocab_size=vocab)
            tr._optimizer = types.SimpleNamespace(param_groups=[{'lr': 0.0}])
            loss = tr.training_step([inputs, targets])
        finally:
            ref_trainer.generate_noise_batch = saved
        loss['loss'].backward()
        pos = m(inputs, targets, proba=False)
        neg = m(inputs, noise.view(b, -1), proba=False).view(b, n, k)
        chk = NegativeSamplingLoss()(pos, neg)
        assert torch.equal(chk['loss'], loss['loss'])
        w_in = m._input_embedding.weight.detach().numpy()
        w_out = m._output_embedding.weight.detach().numpy()
        o = sgns_oracle.training_step(w_in, w_out, inputs.numpy(), targets.numpy(), noise.numpy())
        tol = 1e-12 if dt == torch.float64 else 2e-6
        for key in ('loss', 'positive-loss', 'negative-loss'):
            assert abs(float(o[key]) - float(loss[key])) <= tol * max(1.0, abs(float(loss[key]))), (tag, name, key)
        gi, go = m._input_embedding.weight.grad.numpy(), m._output_embedding.weight.grad.numpy()
        den = max(np.abs(gi).max(), np.abs(go).max(), 1e-30)
        err = max(np.abs(o['grad_in'] - gi).max(), np.abs(o['grad_out'] - go).max()) / den
        assert err <= (1e-12 if dt == torch.float64 else 1e-5), (tag, name, err)
        recall = float((torch.sigmoid(pos) >= 0.5).float().mean())
        precision = float(1 - (torch.sigmoid(neg) >= 0.5).float().mean())
        assert abs(recall - o['recall']) < 1e-6 and abs(precision - o['precision']) < 1e-6
        out.update({
            f'w_in_{name}': w_in, f'w_out_{name}': w_out,
            f'loss_{name}': np.array([float(loss['loss']), float(loss['positive-loss']), float(loss['negative-loss'])]),
            f'grad_in_{name}': gi, f'grad_out_{name}': go,
            f'metrics_{name}': np.array([recall, precision]),
        })
        clamp_hits = int((torch.sigmoid(pos) <= 1e-6).sum() + (torch.sigmoid(-neg) <= 1e-6).sum())
        print(f'sgns_{tag} {name}: loss={float(loss["loss"]):.6f} oracle-vs-autograd rel err={err:.2e} clamp_hits={clamp_hits}')
    np.savez_compressed(os.path.join(GOLDEN, f'sgns_{tag}.npz'), inputs=inputs.numpy(), targets=targets.numpy(),
                        noise=noise.numpy(), meta=np.array(META), **out)


def emit_sgns():
    emit_sgns_case('karate', 35, 2, 384, 4, 1, 31)               # sge_sg_karate_club.yaml shapes
    emit_sgns_case('cora128', 301, 128, 96, 4, 5, 32, scale=8.0)   # E=128, K=5, collisions guaranteed (V<B*N)
    emit_sgns_case('e48', 500, 48, 40, 10, 3, 33, scale=10.0)      # w2v_sg_wiki_text_103.yaml E=48, K=3, r=5
    emit_sgns_case('clamp', 64, 8, 32, 4, 5, 34, scale=60.0)       # |s| > 13.8 -> clamp region (loss.py:15-16)
    emit_sgns_case('e100', 97, 100, 16, 6, 2, 35, scale=6.0)       # E not a power of two


def emit_vocab():
    ds = ref_td.GraphDataset('graph_karate_club', context_radius=2,
                             additional_parameters={'walks_per_node': 2, 'walk_length': 10, 'method': 'deepwalk'})
    itos = ds.vocab.get_itos()
    assert itos[0] == '<unk>' and itos[1:] == sorted(itos[1:]) and len(itos) == 35
    first = next(iter(ds))
    assert first.dtype == torch.long and first.shape == (10,)
    tri = ref_td.GraphDataset('graph_triplets', context_radius=2,
                              additional_parameters={'walks_per_node': 4, 'walk_length': 5, 'method': 'deepwalk'})
    np.savez_compressed(os.path.join(GOLDEN, 'vocab.npz'), karate_itos=np.array(itos),
                        triplets_itos=np.array(tri.vocab.get_itos()), meta=np.array(META))
    print('vocab OK', itos[:3], tri.vocab.get_itos())


def emit_edge_ops():
    """Edge features of a random embedding table for a random edge list, computed by the reference's own operators."""
    from shallow_encoders.graph import edge_operators as ref_ops       # the reference's module
    from oracle import edge_oracle
    assert ref_ops.__file__.startswith(ref_import.REFERENCE_ROOT), ref_ops.__file__
    rng = np.random.default_rng(123)
    out = {}
    for emb in (2, 8, 100, 128):
        table = (rng.standard_normal((120, emb)) * 0.7).astype(np.float32)
        edges = rng.integers(0, 120, (150, 2))
        edges[:5, 1] = edges[:5, 0]                                     # self pairs
        out[f'table_{emb}'] = table
        out[f'edges_{emb}'] = edges.astype(np.int64)
        for name in ('average', 'hadamard', 'weighted_l1', 'weighted_l2'):
            op = ref_ops.edge_operator_factory(name)
            ref = np.stack([op(table[s, :], table[e, :]) for s, e in edges])     # create_edge_embeddings, :219-224
            assert ref.dtype == np.float32
            assert np.array_equal(ref, edge_oracle.edge_embeddings(table, edges, name)), name
            out[f'{name}_{emb}'] = ref
    try:
        ref_ops.edge_operator_factory('cosine')
        raise RuntimeError('reference accepted an unknown operator')
    except AssertionError as e:
        out['unknown_operator_message'] = np.array(str(e))
    np.savez_compressed(os.path.join(GOLDEN, 'edge_ops.npz'), meta=np.array(META), **out)
    print('edge_ops: 4 operators x 4 embedding sizes')


def emit_downstream():
    """The reference's own downstream evaluation on a FIXED embedding (karate club, E = 8): the accuracy yardstick of the north star,
    pinned so that tools/downstream.py (the restatement the B200 repo uses) can be checked against it."""
    import logging
    import types
    for name in ('hydra', 'omegaconf', 'matplotlib', 'matplotlib.pyplot'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['hydra'].main = lambda **_kw: (lambda fn: fn)
    sys.modules['omegaconf'].DictConfig = dict
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    fake_utils = types.ModuleType('tools.utils')
    fake_utils.setup_pipeline = lambda *a, **k: None
    fake_utils.MATPLOTLIB_COLORS = ['b', 'g', 'r']
    sys.modules['tools.utils'] = fake_utils
    import tools.graph_model_downstream_classification as ref_ds            # the reference's module
    from shallow_encoders.split import TrainTestRatioSplit as RefSplit
    assert ref_ds.__file__.startswith(ref_import.REFERENCE_ROOT), ref_ds.__file__
    logging.disable(logging.CRITICAL)

    karate = ref_datasets.KarateClubDataset(walks_per_node=1, walk_length=4)
    graph, labels = karate.graph, karate.labels
    itos = ['<unk>'] + sorted(graph.nodes)
    rng = np.random.default_rng(2024)
    emb = rng.standard_normal((len(itos), 8)).astype(np.float32)
    for i, name in enumerate(itos[1:], start=1):                            # weak class signal + weak neighbourhood signal
        emb[i, 0] += 0.9 if labels[name] == '1' else -0.9
    adj_mean = np.zeros_like(emb)
    for i, name in enumerate(itos[1:], start=1):
        adj_mean[i] = np.mean([emb[itos.index(x)] for x in graph.neighbors(name)], axis=0)
    emb = (0.6 * emb + 0.8 * adj_mean).astype(np.float32)

    class FakeModel:
        input_embedding = torch.from_numpy(emb)

    vocab = ref_import._StubVocab(itos)
    dataset = types.SimpleNamespace(vocab=vocab, labels=labels, has_features=False, graph=graph)

    accs = []
    real_fit = ref_ds.create_and_fit_classification_model

    def recording_fit(*a, **k):
        clf, acc = real_fit(*a, **k)
        accs.append(acc)
        return clf, acc

    ref_ds.create_and_fit_classification_model = recording_fit
    ref_ds.perform_node_classification(model=FakeModel, dataset=dataset, output_path='/tmp', split_algorithm=RefSplit(train_ratio=0.5, test_all=True),
                                       n_experiments=20, visualize=False, classifier_params=None)
    node_accs = np.array(accs, dtype=np.float64)
    accs.clear()
    random.seed(7)
    ref_ds.perform_edge_classification(model=FakeModel, dataset=dataset, train_ratio=0.5, n_experiments=300, edge_operator_name='hadamard',
                                       classifier_params=None)
    edge_accs = np.array(accs, dtype=np.float64)
    ref_ds.create_and_fit_classification_model = real_fit
    logging.disable(logging.NOTSET)
    edges = np.array([(itos.index(a), itos.index(b)) for a, b in graph.edges], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, 'downstream_karate.npz'), meta=np.array(META), embedding=emb, itos=np.array(itos),
                        labels=np.array([labels[v] for v in itos[1:]]), edges=edges, node_accuracies=node_accs,
                        edge_accuracies=edge_accs, sklearn=np.array(__import__('sklearn').__version__))
    print(f'downstream: node acc mean {node_accs.mean():.4f} best {node_accs.max():.4f}; edge acc mean {edge_accs.mean():.4f} (+- {edge_accs.std():.4f})')


if __name__ == '__main__':
    os.makedirs(GOLDEN, exist_ok=True)
    emit_walks()
    emit_collate()
    emit_sgns()
    emit_vocab()
    emit_edge_ops()
    emit_downstream()
    print('golden fixtures written to', GOLDEN)

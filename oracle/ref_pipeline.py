"""
TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product path.

Times the UNMODIFIED reference (imported through oracle/ref_import.py: /root/reference, or the byte-for-byte
copy under baseline/_ref/ on the GPU box) on the host cores, for bench.py's `--impl reference` arm and `cpu_baseline`
block (kind "reference").  It always runs in its OWN process (`python -m oracle.ref_pipeline ...`, one JSON line on
stdout): the reference's package is called `shallow_encoders` like this repository's mirror, so the two must never
share an interpreter, and no native library of this repository is loaded here.

What runs, all of it the reference's own code:
  walks     random_walk_factory(method, graph, L, {p, q}).walk(node)            graph/random_walk_generator.py:94-151
            in `workers` forked processes over disjoint start nodes (the reference's DataLoader workers are
            processes too, config_parser/core.py:173-178)
  tokens    tokenize(walk string) -> vocab(tokens) -> LongTensor                word2vec/dataloader/torch_dataset.py:23-39, 205-213
            (vocab = build_vocab_from_iterator over the node names, '<unk>' first: what GraphDataset builds, :99-110,
            without the throw-away epoch of walks the reference spends on it)
  collate   W2VCollateFunctional('sg', r, max_length)                           torch_dataset.py:276-322
  step      Word2VecTrainer.training_step -> loss['loss'].backward() -> optimizer.step()
                                                                                word2vec/trainer.py:131-152, tools/train.py:67-83
            with SkipGram(V, E) and the YAML's torch.optim.Adam (or SGD for the like-for-like row)

The only things that are not the reference's: the three import stubs of ref_import.py (nltk / torchtext / lightning,
none on this arithmetic path except the vocab ordering) and this driver loop, which replaces Lightning's.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_W = {}


def _worker_walks(args):
    starts, seed = args
    import random
    random.seed(seed)          # the reference uses the module-level `random` (random_walk_generator.py:68,113)
    gen = _W['gen']
    return [gen.walk(s) for s in starts]


def powerlaw_nx(n_nodes, n_edges, seed):
    """networkx graph with the adjacency of cpu_port.powerlaw_graph_host (same generator as the GPU workload), node names
    'n0000042' (a letter so the reference's tokenizer keeps them; zero-padded so lexicographic = numeric order)."""
    import networkx as nx
    from oracle import cpu_port
    g = cpu_port.powerlaw_graph_host(n_nodes, n_edges, seed)
    names = [f'n{i:07d}' for i in range(n_nodes)]
    graph = nx.Graph()
    graph.add_nodes_from(names)
    graph.add_edges_from((names[u], names[v]) for u in range(n_nodes) for v in g.adj[u] if u < v)
    return graph, names


def named_graph(kind, nodes, edges, seed):
    import networkx as nx
    if kind == 'karate':
        g = nx.karate_club_graph()
        return nx.relabel_nodes(g, {i: f'n{i + 1:02d}' for i in g.nodes}), None
    if kind == 'cora_shape':
        g = nx.gnm_random_graph(2708, 5429, seed=seed)
        for u in [n for n in g.nodes if g.degree(n) == 0]:
            g.add_edge(u, (u + 1) % 2708)
        return nx.relabel_nodes(g, {i: f'n{i:07d}' for i in g.nodes}), None
    return powerlaw_nx(nodes, edges, seed)


def run(a):
    from oracle import ref_import
    if ref_import.reference_root() is None:
        return {'unavailable': 'reference not mounted and baseline/_ref not vendored (run python -m oracle.vendor_ref in the build container)'}
    ref_import.import_reference()
    import torch
    from shallow_encoders.graph.random_walk_generator import random_walk_factory
    from shallow_encoders.word2vec.dataloader import torch_dataset as ref_td
    from shallow_encoders.word2vec.model import SkipGram
    from shallow_encoders.word2vec.trainer import Word2VecTrainer
    from torchtext.vocab import build_vocab_from_iterator

    root = ref_import.reference_root()
    for mod in (ref_td, sys.modules[SkipGram.__module__], sys.modules[Word2VecTrainer.__module__], sys.modules[random_walk_factory.__module__]):
        assert mod.__file__.startswith(root), f'{mod.__name__} was not imported from the reference ({mod.__file__})'

    workers = a.workers or (os.cpu_count() or 1)
    torch.set_num_threads(workers)
    torch.manual_seed(a.seed)
    graph, _ = named_graph(a.graph, a.nodes, a.edges, a.seed)
    nodes = sorted(graph.nodes)
    params = {'p': a.p, 'q': a.q} if a.method == 'node2vec' else {}
    gen = random_walk_factory(a.method, graph, a.walk_len, params)
    vocab = build_vocab_from_iterator([[t] for t in {n.lower() for n in nodes}], specials=['<unk>'], min_freq=0)
    vocab.set_default_index(vocab['<unk>'])
    collate = ref_td.W2VCollateFunctional('sg', a.radius, a.max_length)
    model = SkipGram(vocab_size=len(vocab), embedding_size=a.emb)
    opt = (torch.optim.Adam if a.optimizer == 'adam' else torch.optim.SGD)(model.parameters(), lr=a.lr)
    trainer = Word2VecTrainer(model=model, optimizer=opt, scheduler=None, neg_samples=a.neg, vocab_size=len(vocab))

    pool = None
    _W['gen'] = gen
    if workers > 1 and not a.no_pool:
        import multiprocessing as mp
        pool = mp.get_context('fork').Pool(workers)

    import random
    rnd = random.Random(a.seed)
    order = list(nodes)
    rnd.shuffle(order)                                          # graph/datasets.py:45

    def walks_for(it):
        # graph/datasets.py:76: walks_per_node consecutive walks per node of the shuffled list
        base = it * a.walks_per_step
        starts = [order[((base + j) // a.walks_per_node) % len(order)] for j in range(a.walks_per_step)]
        if pool is None:
            return _worker_walks((starts, a.seed * 7919 + it))
        n = min(workers, len(starts))
        chunks = [starts[i::n] for i in range(n)]
        parts = pool.map(_worker_walks, [(c, (a.seed * 7919 + it) * 1000 + i) for i, c in enumerate(chunks)])
        return [w for part in parts for w in part]

    t_walk = t_tok = t_sgns = 0.0
    pairs = steps_walked = 0
    loss_last = None
    try:
        for it in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            sentences = walks_for(it)
            t1 = time.perf_counter()
            texts = [torch.tensor(vocab(ref_td.tokenize(s)), dtype=torch.long) for s in sentences]     # torch_dataset.py:209-212
            batch = collate(texts)
            t2 = time.perf_counter()
            loss = trainer.training_step(list(batch))
            opt.zero_grad()
            loss['loss'].backward()
            opt.step()
            t3 = time.perf_counter()
            if it >= a.warmup:
                t_walk += t1 - t0
                t_tok += t2 - t1
                t_sgns += t3 - t2
                pairs += batch[1].shape[0] * batch[1].shape[1]
                steps_walked += len(sentences) * (a.walk_len - 1)
                loss_last = float(loss['loss'].detach())
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    total = t_walk + t_tok + t_sgns
    return {
        'kind': 'reference', 'reference_root': root, 'seconds': total, 'pairs': pairs, 'walk_steps': steps_walked,
        'pairs_per_s': pairs / total, 'walk_steps_per_s': steps_walked / max(t_walk, 1e-9), 'sgns_pairs_per_s': pairs / max(t_sgns, 1e-9),
        't_walk': t_walk, 't_collate': t_tok, 't_sgns': t_sgns, 'ms_per_step': 1e3 * total / max(a.steps, 1), 'cores': workers,
        'cpu_count': os.cpu_count(), 'torch_threads': torch.get_num_threads(), 'loss': loss_last, 'vocab': len(vocab),
        'graph': {'kind': a.graph, 'nodes': graph.number_of_nodes(), 'edges': graph.number_of_edges()},
        'versions': {'python': sys.version.split()[0], 'torch': torch.__version__, 'networkx': __import__('networkx').__version__},
    }


def sgns_only(a):
    """training_step + backward + optimizer.step on a fixed random index batch of the given shape."""
    from oracle import ref_import
    if ref_import.reference_root() is None:
        return {'unavailable': 'reference not available'}
    ref_import.import_reference()
    import torch
    from shallow_encoders.word2vec.model import SkipGram
    from shallow_encoders.word2vec.trainer import Word2VecTrainer
    workers = a.workers or (os.cpu_count() or 1)
    torch.set_num_threads(workers)
    torch.manual_seed(a.seed)
    model = SkipGram(vocab_size=a.vocab, embedding_size=a.emb)
    opt = (torch.optim.Adam if a.optimizer == 'adam' else torch.optim.SGD)(model.parameters(), lr=a.lr)
    trainer = Word2VecTrainer(model=model, optimizer=opt, scheduler=None, neg_samples=a.neg, vocab_size=a.vocab)
    inputs = torch.randint(0, a.vocab, (a.batch_rows, 1))
    targets = torch.randint(0, a.vocab, (a.batch_rows, 2 * a.radius))
    t = 0.0
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        loss = trainer.training_step([inputs, targets])
        opt.zero_grad()
        loss['loss'].backward()
        opt.step()
        if it >= a.warmup:
            t += time.perf_counter() - t0
    pairs = a.steps * a.batch_rows * 2 * a.radius
    return {'kind': 'reference', 'what': 'training_step + backward + optimizer.step', 'vocab': a.vocab, 'emb': a.emb, 'neg': a.neg,
            'batch_rows': a.batch_rows, 'n_ctx': 2 * a.radius, 'optimizer': a.optimizer, 'ms_per_step': 1e3 * t / a.steps,
            'pairs_per_s': pairs / t, 'cores': workers}


def walks_only(a):
    """Node2Vec.walk / DeepWalk.walk steps/s over disjoint start nodes, 1 process or `workers` processes."""
    from oracle import ref_import
    if ref_import.reference_root() is None:
        return {'unavailable': 'reference not available'}
    ref_import.import_reference()
    from shallow_encoders.graph.random_walk_generator import random_walk_factory
    graph, _ = named_graph(a.graph, a.nodes, a.edges, a.seed)
    nodes = sorted(graph.nodes)
    params = {'p': a.p, 'q': a.q} if a.method == 'node2vec' else {}
    _W['gen'] = random_walk_factory(a.method, graph, a.walk_len, params)
    workers = a.workers or (os.cpu_count() or 1)
    n_walks = a.walks_per_step * a.steps
    starts = [nodes[i % len(nodes)] for i in range(n_walks)]
    t0 = time.perf_counter()
    if workers == 1 or a.no_pool:
        _worker_walks((starts, a.seed))
        used = 1
    else:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(workers) as pool:
            t0 = time.perf_counter()
            pool.map(_worker_walks, [(starts[i::workers], a.seed + i) for i in range(workers)])
        used = workers
    dt = time.perf_counter() - t0
    return {'kind': 'reference', 'what': f'{a.method}.walk', 'graph': a.graph, 'nodes': graph.number_of_nodes(), 'walk_len': a.walk_len,
            'processes': used, 'walk_steps_per_s': n_walks * (a.walk_len - 1) / dt, 'seconds': dt}


def dataloader_epoch(a):
    """The reference's own data path end to end on karate: GraphDataset -> DataLoader(num_workers) -> W2VCollateFunctional ->
    training_step -> backward -> Adam (num_workers = 8 is the shipped setting and generates 8x the walks)."""
    from oracle import ref_import
    if ref_import.reference_root() is None:
        return {'unavailable': 'reference not available'}
    ref_import.import_reference()
    import torch
    from torch.utils.data import DataLoader
    from shallow_encoders.word2vec.dataloader.torch_dataset import GraphDataset, W2VCollateFunctional
    from shallow_encoders.word2vec.model import SkipGram
    from shallow_encoders.word2vec.trainer import Word2VecTrainer
    torch.manual_seed(a.seed)
    ds = GraphDataset('graph_karate_club', context_radius=2,
                      additional_parameters={'walks_per_node': 64, 'walk_length': 10, 'method': 'node2vec', 'method_params': {'p': 1, 'q': 0.5}})
    model = SkipGram(vocab_size=len(ds.vocab), embedding_size=2)
    opt = torch.optim.Adam(model.parameters(), lr=0.1)
    trainer = Word2VecTrainer(model=model, optimizer=opt, scheduler=None, neg_samples=1, vocab_size=len(ds.vocab))
    dl = DataLoader(ds, batch_size=64, num_workers=a.workers, collate_fn=W2VCollateFunctional('sg', 2, 256))
    pairs = rows = 0
    t0 = time.perf_counter()
    for _ in range(a.steps):                      # epochs
        for batch in dl:
            loss = trainer.training_step(list(batch))
            opt.zero_grad(); loss['loss'].backward(); opt.step()
            pairs += batch[1].shape[0] * batch[1].shape[1]; rows += batch[1].shape[0]
    dt = time.perf_counter() - t0
    return {'kind': 'reference', 'what': 'GraphDataset + DataLoader + training_step + Adam (karate YAML values)', 'num_workers': a.workers,
            'epochs': a.steps, 'rows_per_epoch': rows / a.steps, 'pairs_per_s': pairs / dt, 'walk_steps_per_s': rows / 6 * 9 / dt, 'seconds': dt}


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='pipeline', choices=['pipeline', 'sgns', 'walks', 'dataloader'])
    ap.add_argument('--graph', default='powerlaw', choices=['powerlaw', 'karate', 'cora_shape'])
    ap.add_argument('--nodes', type=int, default=100_000)
    ap.add_argument('--edges', type=int, default=2_500_000)
    ap.add_argument('--method', default='node2vec')
    ap.add_argument('--p', type=float, default=0.5)
    ap.add_argument('--q', type=float, default=2.0)
    ap.add_argument('--walk-len', type=int, default=80)
    ap.add_argument('--walks-per-node', type=int, default=10)
    ap.add_argument('--walks-per-step', type=int, default=64)
    ap.add_argument('--radius', type=int, default=5)
    ap.add_argument('--max-length', type=int, default=1 << 30)
    ap.add_argument('--emb', type=int, default=128)
    ap.add_argument('--neg', type=int, default=5)
    ap.add_argument('--optimizer', default='adam', choices=['adam', 'sgd'])
    ap.add_argument('--lr', type=float, default=0.1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--workers', type=int, default=0)
    ap.add_argument('--no-pool', action='store_true', help='generate the walks in the main process (num_workers = 0)')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--vocab', type=int, default=267_736)
    ap.add_argument('--batch-rows', type=int, default=1888)
    return ap.parse_args(argv)


def main():
    a = parse()
    real_stdout = os.dup(1)
    os.dup2(2, 1)                      # anything libraries print goes to stderr; stdout carries one JSON line
    out = {'pipeline': run, 'sgns': sgns_only, 'walks': walks_only, 'dataloader': dataloader_epoch}[a.mode](a)
    os.write(real_stdout, (json.dumps(out) + '\n').encode())


def call(timeout=900, **kw):
    """Run this module in a fresh interpreter and return its JSON (used by bench.py)."""
    import subprocess
    cmd = [sys.executable, '-m', 'oracle.ref_pipeline']
    for k, v in kw.items():
        flag = '--' + k.replace('_', '-')
        if isinstance(v, bool):
            if v:
                cmd.append(flag)
        else:
            cmd += [flag, str(v)]
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    res = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    if res.returncode != 0:
        return {'unavailable': f'reference pipeline exited {res.returncode}: {res.stderr.decode()[-400:]}'}
    return json.loads(res.stdout.decode().strip().splitlines()[-1])


if __name__ == '__main__':
    main()

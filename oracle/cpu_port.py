"""
TEST / BENCH INFRASTRUCTURE ONLY -- the reference's CPU path restated end to end so it can be TIMED on the GPU
box's host cores, where /root/reference does not exist (bench.py `cpu_baseline` and `--impl reference`, kind "port").
Never imported by the product path.

It keeps the reference's computational pattern, not just its results:
  * walks: pure-Python per-step list building, `t in N(x)` membership on python containers, normalise, one
    random() per transition through the inverse-CDF rule   (graph/random_walk_generator.py:94-119;
    restated in oracle/walk_oracle.py, which is pinned to the reference by tests/golden/walks_*.npz);
    one process per host core over disjoint start nodes (the reference's DataLoader workers are processes too,
    config_parser/core.py:173-178)
  * collate: python double loop over centres              (word2vec/dataloader/torch_dataset.py:293-322)
  * SGNS step: torch CPU nn.Embedding x2 -> bmm -> sigmoid/clamp/log -> mean -> backward -> dense torch.optim.Adam,
    noise from torch.randint on the CPU                  (word2vec/model.py:79-91, loss.py:14-22,
                                                          trainer.py:131-152, utils/sampling.py:21,
                                                          configs/*.yaml `_target_: torch.optim.Adam`)
    (`tests/test_cpu_port.py` checks this step against the golden loss/gradients.)
"""
import os
import random
import time
from typing import Dict, List

import numpy as np

from oracle import walk_oracle


def powerlaw_graph_host(n_nodes: int, n_edges: int, seed: int = 0, gamma: float = 2.0) -> walk_oracle.OracleGraph:
    """Same recipe as shallow_encoders.graph.synthetic.powerlaw_graph_device, in numpy."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n_nodes)
    a = perm[np.minimum((rng.random(n_edges) ** gamma * n_nodes).astype(np.int64), n_nodes - 1)]
    b = perm[np.minimum((rng.random(n_edges) ** gamma * n_nodes).astype(np.int64), n_nodes - 1)]
    ring = rng.permutation(n_nodes)
    src = np.concatenate([a, ring])
    dst = np.concatenate([b, np.roll(ring, 1)])
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = np.unique(np.concatenate([src * n_nodes + dst, dst * n_nodes + src]))
    rows, cols = key // n_nodes, key % n_nodes
    bounds = np.searchsorted(rows, np.arange(n_nodes + 1))
    adj = [cols[bounds[i]:bounds[i + 1]].tolist() for i in range(n_nodes)]
    return walk_oracle.OracleGraph(adj, None, None)


_G = {}


def _walk_chunk(args):
    starts, length, p, q, node2vec, seed = args
    g = _G['g']
    rnd = random.Random(seed)
    out = []
    for s in starts:
        prev, node = None, s
        walk = [s]
        while len(walk) < length:
            w = walk_oracle.transition_weights(g, prev, node, p, q, node2vec)
            child = g.adj[node][walk_oracle.choose(w, rnd.random())]
            walk.append(child)
            prev, node = node, child
        out.append(walk)
    return out


class WalkPool:
    """Process pool over the host cores; the graph is inherited by fork (no pickling per task)."""

    def __init__(self, graph: walk_oracle.OracleGraph, workers: int):
        import multiprocessing as mp
        _G['g'] = graph
        self.workers = max(1, workers)
        self._pool = mp.get_context('fork').Pool(self.workers) if self.workers > 1 else None

    def walks(self, starts: List[int], length: int, p: float, q: float, node2vec: bool, seed: int) -> np.ndarray:
        if self._pool is None:
            return np.array(_walk_chunk((list(starts), length, p, q, node2vec, seed)), dtype=np.int64)
        chunks = [list(c) for c in np.array_split(np.asarray(starts), self.workers) if len(c)]
        parts = self._pool.map(_walk_chunk, [(c, length, p, q, node2vec, seed * 1000 + i) for i, c in enumerate(chunks)])
        return np.array([w for part in parts for w in part], dtype=np.int64)

    def close(self):
        if self._pool is not None:
            self._pool.close()
            self._pool.join()


def collate(walks: np.ndarray, radius: int, max_length: int, row_offset: int):
    """Python double loop with per-centre torch.cat, as torch_dataset.py:293-322."""
    import torch
    batch_inputs, batch_targets = [], []
    for text in walks:
        text = torch.from_numpy(text + row_offset)[:max_length]
        n = text.shape[0]
        assert n >= 2 * radius + 1
        for i in range(radius, n - radius):
            batch_inputs.append(text[i:i + 1])
            batch_targets.append(torch.cat([text[i - radius:i], text[i + 1:i + 1 + radius]]))
    return torch.stack(batch_inputs), torch.stack(batch_targets)


class TorchCpuSgns:
    """The reference's model + loss + optimiser step in plain torch on the CPU."""

    def __init__(self, vocab: int, emb: int, n_neg: int, lr: float = 0.1, optimizer: str = 'adam', seed: int = 0):
        import torch
        torch.manual_seed(seed)
        self.torch = torch
        self.w_in = torch.nn.Embedding(vocab, emb)
        self.w_out = torch.nn.Embedding(vocab, emb)
        torch.nn.init.xavier_uniform_(self.w_in.weight)
        torch.nn.init.xavier_uniform_(self.w_out.weight)
        params = list(self.w_in.parameters()) + list(self.w_out.parameters())
        self.opt = torch.optim.Adam(params, lr=lr) if optimizer == 'adam' else torch.optim.SGD(params, lr=lr)
        self.vocab, self.n_neg = vocab, n_neg

    def scores(self, inputs, outputs):
        b = outputs.shape[0]
        return self.torch.bmm(self.w_out(outputs), self.w_in(inputs).view(b, -1, 1)).view(b, -1)

    def loss(self, inputs, targets, noise) -> Dict:
        torch = self.torch
        b, n = targets.shape
        pos = self.scores(inputs, targets)
        neg = self.scores(inputs, noise.view(b, -1)).view(b, n, -1)
        pl = -torch.log(torch.clamp(torch.sigmoid(pos), min=1e-6))
        nl = -torch.log(torch.clamp(torch.sigmoid(-neg), min=1e-6)).sum(-1)
        return {'loss': torch.mean(pl + nl), 'positive-loss': torch.mean(pl), 'negative-loss': torch.mean(nl)}

    def step(self, inputs, targets) -> float:
        torch = self.torch
        noise = torch.randint(0, self.vocab, (targets.shape[0], targets.shape[1], self.n_neg), dtype=torch.long)
        out = self.loss(inputs, targets, noise)
        self.opt.zero_grad()
        out['loss'].backward()
        self.opt.step()
        return float(out['loss'].detach())


def run_reference_pipeline(graph: walk_oracle.OracleGraph, walks_per_step: int, steps: int, warmup: int, length: int,
                           p: float, q: float, node2vec: bool, radius: int, emb: int, n_neg: int, workers: int,
                           optimizer: str = 'adam', seed: int = 0) -> Dict:
    """`steps` timed batches of `walks_per_step` walks through walk -> collate -> SGNS step; wall-clock seconds."""
    import torch
    torch.set_num_threads(max(1, workers))
    pool = WalkPool(graph, workers)
    model = TorchCpuSgns(graph.n_nodes + 1, emb, n_neg, optimizer=optimizer, seed=seed)
    rnd = random.Random(seed)
    nodes = list(range(graph.n_nodes))
    rnd.shuffle(nodes)
    t_walk = t_collate = t_sgns = 0.0
    pairs = n_steps = 0
    per_step = []
    try:
        for it in range(warmup + steps):
            starts = [nodes[(it * walks_per_step + j) % len(nodes)] for j in range(walks_per_step)]
            t0 = time.perf_counter()
            w = pool.walks(starts, length, p, q, node2vec, seed + it)
            t1 = time.perf_counter()
            inputs, targets = collate(w, radius, 1 << 30, 1)
            t2 = time.perf_counter()
            model.step(inputs, targets)
            t3 = time.perf_counter()
            if it >= warmup:
                t_walk += t1 - t0
                t_collate += t2 - t1
                t_sgns += t3 - t2
                per_step.append(t3 - t0)
                pairs += inputs.shape[0] * targets.shape[1]
                n_steps += walks_per_step * (length - 1)
    finally:
        pool.close()
    total = t_walk + t_collate + t_sgns
    return {'seconds': total, 'pairs': pairs, 'walk_steps': n_steps, 'pairs_per_s': pairs / total,
            'walk_steps_per_s': n_steps / max(t_walk, 1e-9), 'sgns_pairs_per_s': pairs / max(t_sgns, 1e-9),
            't_walk': t_walk, 't_collate': t_collate, 't_sgns': t_sgns, 'ms_per_step': 1e3 * total / max(steps, 1),
            'cores': workers, 'cpu_count': os.cpu_count()}


def train_reference_cpu(nx_graph, walks_per_node: int, walk_length: int, method: str, p: float, q: float, radius: int,
                        emb: int, n_neg: int, batch_size: int, lr: float, max_epochs: int, step_size: int, gamma: float,
                        workers: int = 1, seed: int = 0):
    """The reference's training run on the CPU (walk -> collate -> training_step -> backward -> Adam; StepLR per epoch,
    tools/train.py:67-83 + configs/*.yaml).  Returns the input embedding table [V, E] (row 0 = '<unk>')."""
    import torch
    torch.manual_seed(seed)
    g = walk_oracle.OracleGraph.from_networkx(nx_graph)
    pool = WalkPool(g, workers)
    model = TorchCpuSgns(g.n_nodes + 1, emb, n_neg, lr=lr, optimizer='adam', seed=seed)
    sched = torch.optim.lr_scheduler.StepLR(model.opt, step_size=step_size, gamma=gamma)
    rnd = random.Random(seed)
    nodes = list(range(g.n_nodes))
    losses = []
    try:
        for epoch in range(max_epochs):
            rnd.shuffle(nodes)
            starts = [n for n in nodes for _ in range(walks_per_node)]
            walks = pool.walks(starts, walk_length, p, q, method == 'node2vec', seed * 100003 + epoch)
            tot = 0.0
            for lo in range(0, len(walks), batch_size):
                inputs, targets = collate(walks[lo:lo + batch_size], radius, 1 << 30, 1)
                tot += model.step(inputs, targets)
            losses.append(tot / max(1, (len(walks) + batch_size - 1) // batch_size))
            sched.step()
    finally:
        pool.close()
    return model.w_in.weight.detach().numpy().copy(), g.names, losses

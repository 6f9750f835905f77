"""
Synthetic graphs of the benchmark shapes (BASELINE.json configs), generated on the device with torch as plumbing.
Not part of the reference; the reference's real datasets (graph/datasets.py:126-221) need files/network that are
not available here.
"""
import torch

from shallow_encoders.graph.csr import CSRGraph


def powerlaw_graph_device(n_nodes: int, n_edges: int, seed: int = 0, device='cuda', gamma: float = 2.0,
                          chunk: int = 1 << 26) -> CSRGraph:
    """Undirected power-law graph (Chung-Lu style): endpoints drawn with density ~ rank^(1/gamma - 1)
    (gamma = 2 -> degree exponent 3), node ids shuffled, plus one random-permutation edge per node so that no node
    is isolated (the reference crashes on degree-0 nodes).  ~n_edges undirected edges before de-duplication."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    perm = torch.randperm(n_nodes, generator=gen, device=dev)
    srcs, dsts = [], []
    left = n_edges
    while left > 0:
        m = min(left, chunk)
        a = (torch.rand(m, generator=gen, device=dev, dtype=torch.float64).pow_(gamma) * n_nodes).long().clamp_(max=n_nodes - 1)
        b = (torch.rand(m, generator=gen, device=dev, dtype=torch.float64).pow_(gamma) * n_nodes).long().clamp_(max=n_nodes - 1)
        srcs.append(perm[a].to(torch.int32))
        dsts.append(perm[b].to(torch.int32))
        left -= m
    ring = torch.randperm(n_nodes, generator=gen, device=dev)
    srcs.append(ring.to(torch.int32))
    dsts.append(torch.roll(ring, 1).to(torch.int32))
    src, dst = torch.cat(srcs), torch.cat(dsts)
    del srcs, dsts
    return CSRGraph.from_edges_device(src, dst, n_nodes, symmetrize=True)


def sbm_graph_device(n_nodes: int, n_edges: int, n_blocks: int, p_in: float = 0.85, seed: int = 0, device='cuda'):
    """Cora-shaped labelled graph: `n_blocks` communities, a fraction p_in of the edges inside a community.
    Returns (CSRGraph, labels int64[n_nodes])."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    labels = torch.randint(0, n_blocks, (n_nodes,), generator=gen, device=dev)
    order = torch.argsort(labels)
    counts = torch.bincount(labels, minlength=n_blocks)
    offs = torch.cumsum(counts, 0) - counts
    a = torch.randint(0, n_nodes, (n_edges,), generator=gen, device=dev)
    inside = torch.rand(n_edges, generator=gen, device=dev) < p_in
    la = labels[a]
    b_in = order[offs[la] + (torch.rand(n_edges, generator=gen, device=dev) * counts[la]).long().clamp_(max=n_nodes - 1)]
    b_out = torch.randint(0, n_nodes, (n_edges,), generator=gen, device=dev)
    b = torch.where(inside, b_in, b_out)
    ring = torch.randperm(n_nodes, generator=gen, device=dev)
    same = ring[torch.argsort(labels[ring], stable=True)]       # ring inside label order keeps most ring edges intra-block
    src = torch.cat([a, same])
    dst = torch.cat([b, torch.roll(same, 1)])
    return CSRGraph.from_edges_device(src, dst, n_nodes, symmetrize=True), labels

"""Developer probe (not the bench): quick device-timed throughput of the walk and SGNS kernels."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deepwalk-and-node2vec_b200'))
import torch

from shallow_encoders import _native as nat
from shallow_encoders.graph.synthetic import powerlaw_graph_device

ap = argparse.ArgumentParser()
ap.add_argument('--nodes', type=int, default=1_000_000)
ap.add_argument('--edges', type=int, default=25_000_000)
ap.add_argument('--walks', type=int, default=262144)
ap.add_argument('--len', type=int, default=80)
ap.add_argument('--emb', type=int, default=128)
ap.add_argument('--neg', type=int, default=5)
ap.add_argument('--radius', type=int, default=5)
ap.add_argument('--iters', type=int, default=5)
args = ap.parse_args()
dev = torch.device('cuda:0')
print(nat.version(), nat.device_info(), flush=True)
t0 = time.time()
csr = powerlaw_graph_device(args.nodes, args.edges, 0, dev)
torch.cuda.synchronize()
print(f'graph: n={csr.n_nodes} nnz={csr.nnz} max_deg={csr.max_degree} bytes={csr.nbytes() / 1e9:.2f} GB gen={time.time() - t0:.1f}s', flush=True)
starts = torch.randint(0, args.nodes, (args.walks,), device=dev, dtype=torch.int32)


def timed(fn, iters):
    fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i + 1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


walks = torch.empty((args.walks, args.len), dtype=torch.int32, device=dev)
steps = args.walks * (args.len - 1)
for name, kw in [('deepwalk', dict(p=1.0, q=1.0, node2vec=False)), ('node2vec p=.5 q=2', dict(p=0.5, q=2.0, node2vec=True)),
                 ('node2vec p=1 q=.5', dict(p=1.0, q=0.5, node2vec=True))]:
    for kname, kern in (('warp', nat.WALK_WARP), ('thread', nat.WALK_THREAD)):
        ms = timed(lambda i: nat.walk(csr, starts, args.len, kw['p'], kw['q'], kw['node2vec'], 0, seed=i, out=walks, kernel=kern), args.iters)
        print(f'walk {name:20s} {kname:6s}: {ms:8.3f} ms  {steps / ms / 1e6:8.3f} G steps/s', flush=True)

vocab = args.nodes + 1
w_in = (torch.rand(vocab, args.emb, device=dev) - 0.5) * 0.1
w_out = (torch.rand(vocab, args.emb, device=dev) - 0.5) * 0.1
pairs = args.walks * (args.len - 2 * args.radius) * 2 * args.radius
bpp = 2 * 4 * args.emb * (1 + args.neg + 1 / (2 * args.radius))
stats = torch.zeros(6, dtype=torch.float64, device=dev)
for name, flags in [('red', nat.SCATTER_RED), ('store', nat.SCATTER_STORE), ('generic red', 2), ('generic store', 3)]:
    ms = timed(lambda i: nat.sgns_update_walks(w_in, w_out, walks, args.radius, args.neg, 1, 0.025, seed=i, flags=flags, stats=stats), args.iters)
    print(f'sgns {name:14s}: {ms:8.3f} ms  {pairs / ms / 1e6:8.2f} M pairs/s  {pairs * bpp / ms / 1e6:8.1f} GB/s algorithmic '
          f'({pairs * bpp / ms / 1e6 / 6450.6 * 100:.1f}% of 6450.6)', flush=True)
print('tables GB', 2 * vocab * args.emb * 4 / 1e9)

# ---- link-prediction features: fused gather + operator, 3 * 4E bytes per edge (two random row reads, one streamed write) --------------
n_e = min(csr.nnz, 20_000_000)
deg = csr.rowptr[1:] - csr.rowptr[:-1]
e_src = torch.repeat_interleave(torch.arange(csr.n_nodes, device=dev), deg)[:n_e] + 1
e_dst = csr.col_sorted[:n_e].to(torch.int64) + 1
perm = torch.randperm(n_e, device=dev)          # random edge order: both row reads are random, as for sampled edge lists
e_src, e_dst = e_src[perm].contiguous(), e_dst[perm].contiguous()
for op in ('hadamard', 'weighted_l2'):
    ms = timed(lambda i: nat.edge_features(w_in, e_src, e_dst, op), args.iters)
    gbs = n_e * 3 * 4 * args.emb / ms / 1e6
    print(f'edge_features {op:12s}: {ms:8.3f} ms  {n_e / ms / 1e6:8.3f} G edges/s  {gbs:8.1f} GB/s algorithmic ({gbs / 6450.6 * 100:.1f}% of 6450.6)', flush=True)
ms = timed(lambda i: nat.sample_negative_edges(csr, n_e, seed=i), args.iters)
print(f'negative edges       : {ms:8.3f} ms  {n_e / ms / 1e6:8.3f} G samples/s', flush=True)

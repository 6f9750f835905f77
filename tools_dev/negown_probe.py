"""Developer probe (not the bench): the per-GPU cost of one owner-computes step of a WORLD-GPU job, measured on ONE GPU.

Rank 0 of a simulated `--world`-way striping (every stripe local, so no NVLink effects: the compute / HBM side only) processes the
walks of all ranks: positives of its own walks (window kernel, K = 0) + owned negatives of the gathered batch in walk order
(`sgns_negown_kernel`) and bucketed by centre row (`sgns_negown_grouped_kernel`).  S3 sizes by default."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deepwalk-and-node2vec_b200'))
import torch

from shallow_encoders import _native as nat
from shallow_encoders.graph.synthetic import powerlaw_graph_device
from shallow_encoders.word2vec.sharded import ShardedTable

ap = argparse.ArgumentParser()
ap.add_argument('--nodes', type=int, default=10_000_000)
ap.add_argument('--edges', type=int, default=250_000_000)
ap.add_argument('--walks', type=int, default=262144, help='walks per rank per step')
ap.add_argument('--world', type=int, default=8)
ap.add_argument('--len', type=int, default=80)
ap.add_argument('--emb', type=int, default=128)
ap.add_argument('--neg', type=int, default=5)
ap.add_argument('--radius', type=int, default=5)
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--modes', default='all-pairs,grouped,walk-order')
ap.add_argument('--out', default=None)
args = ap.parse_args()
dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
print(nat.version(), flush=True)
t0 = time.time()
csr = powerlaw_graph_device(args.nodes, args.edges, 0, dev)
torch.cuda.synchronize()
print(f'graph: n={csr.n_nodes} nnz={csr.nnz} max_deg={csr.max_degree} gen={time.time() - t0:.1f}s', flush=True)

vocab = args.nodes + 1
w_in = ShardedTable(vocab, args.emb, dev, rank=0, world=args.world, simulate=True)
w_out = ShardedTable(vocab, args.emb, dev, rank=0, world=args.world, simulate=True)
bound = (6.0 / (vocab + args.emb)) ** 0.5
for r in range(args.world):
    nat.table_fill_uniform(w_in.as_rank(r), bound, 101)
    nat.table_fill_uniform(w_out.as_rank(r), bound, 102)
all_walks = torch.empty((args.world * args.walks, args.len), dtype=torch.int32, device=dev)
g = torch.Generator(device='cpu'); g.manual_seed(0)
for r in range(args.world):
    starts = torch.randint(0, args.nodes, (args.walks,), generator=g, dtype=torch.int32).to(dev)
    nat.walk(csr, starts, args.len, 0.5, 2.0, True, nat.RULE_REFERENCE, 1, walk_id_base=r * args.walks, out=all_walks[r * args.walks:(r + 1) * args.walks])
torch.cuda.synchronize()
n_cen = args.len - 2 * args.radius
pairs_rank = args.walks * n_cen * 2 * args.radius
stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)


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


res = {'world': args.world, 'walks_per_rank': args.walks, 'pairs_per_rank_per_step': pairs_rank}
my = all_walks[:args.walks]
res['positives_pair_by_pair_ms'] = timed(lambda i: nat.sgns_update_walks(w_in, w_out, my, args.radius, 0, 1, 0.025, 7, centre_id_base=i * 10 ** 9, stats=stats),
                                         args.iters)
print(f"positives (window kernel, K=0) : {res['positives_pair_by_pair_ms']:8.2f} ms pair by pair", flush=True)
res['positives_ms'] = timed(lambda i: nat.sgns_update_walks(w_in, w_out, my, args.radius, 0, 1, 0.025, 7, centre_id_base=i * 10 ** 9, stats=stats,
                                                            flags=nat.BATCHED_POSITIVES), args.iters)
print(f"positives (window kernel, K=0) : {res['positives_ms']:8.2f} ms batched per centre (SE_SGNS_BATCHED_POSITIVES)", flush=True)
single = timed(lambda i: nat.sgns_update_walks(w_in, w_out, my, args.radius, args.neg, 1, 0.025, 7, centre_id_base=i * 10 ** 9, stats=stats), args.iters)
res['single_gpu_step_ms'] = single
print(f'single-GPU fused step          : {single:8.2f} ms  ({pairs_rank / single / 1e6:.3f} G pairs/s)', flush=True)
for mode in args.modes.split(','):
    stats.zero_()
    if mode == 'all-pairs':
        fn = lambda i: nat.sgns_update_pairs_owned(w_in, w_out, all_walks, args.radius, args.neg, 1, 0.025, 7, centre_id_base=i * 10 ** 9,   # noqa: E731
                                                   stats=stats, positives=True)
    else:
        fn = lambda i: nat.sgns_update_negatives_owned(w_in, w_out, all_walks, args.radius, args.neg, 1, 0.025, 7, centre_id_base=i * 10 ** 9,   # noqa: E731
                                                       stats=stats, grouped=mode == 'grouped')
    ms = timed(fn, args.iters)
    st = stats.tolist()
    res[mode + '_ms'] = ms
    res[mode + '_owned_negatives_per_launch'] = st[5] / (args.iters + 1)
    res[mode + '_owned_positives_per_launch'] = st[4] / (args.iters + 1)
    res[mode + '_mean_negative_loss'] = st[1] / max(st[5], 1)
    tot = ms + (0.0 if mode == 'all-pairs' else res['positives_ms'])
    print(f'{mode:12s}: {ms:8.2f} ms   SGNS stage {tot:8.2f} ms = {single / tot:.3f} of the single-GPU rate '
          f'(owned negatives per launch {st[5] / (args.iters + 1):.4g}, positives {st[4] / (args.iters + 1):.4g}, mean negative loss {st[1] / max(st[5], 1):.4f})',
          flush=True)
if args.out:
    with open(args.out, 'w') as fh:
        json.dump(res, fh, indent=1)
w_in.close(); w_out.close()

"""Does the multi-GPU design keep the model quality?  Trains the SAME labelled graph (stochastic block model) with the fused
engine in three ways and compares downstream node-classification accuracy of W_in:
    1 GPU, reference negatives (uniform over the whole table)
    G GPUs, one striped table pair, reference (global) negatives
    G GPUs, one striped table pair, LOCAL negatives (each GPU draws among the rows it owns; bench.py's default)
    G GPUs, one striped table pair, reference negatives, OWNER-COMPUTES (every GPU processes the negatives whose rows it owns)
    G GPUs, SYNCED working copies + row-sharded masters, reference negatives (bench.py's N > 1 headline: csrc/replica.cu)
Run under torchrun with G >= 2 ranks (rank 0 also does the 1-GPU run); writes gpurun_out/multi_gpu_accuracy.json.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools_dev/multi_gpu_accuracy.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deepwalk-and-node2vec_b200'))
import numpy as np
import torch
import torch.distributed as dist

from shallow_encoders import _native as nat
from shallow_encoders.graph.synthetic import sbm_graph_device
from shallow_encoders.word2vec.sharded import ReplicatedTable, ShardedTable, make_exchange, sgns_update_walks_owner_computes, sync_replicated

ap = argparse.ArgumentParser()
ap.add_argument('--nodes', type=int, default=200_000)
ap.add_argument('--edges', type=int, default=2_000_000)
ap.add_argument('--blocks', type=int, default=20)
ap.add_argument('--p-in', type=float, default=0.3)
ap.add_argument('--emb', type=int, default=128)
ap.add_argument('--walk-len', type=int, default=40)
ap.add_argument('--walks-per-node', type=int, default=10)
ap.add_argument('--radius', type=int, default=5)
ap.add_argument('--neg', type=int, default=5)
ap.add_argument('--epochs', type=int, default=2)
ap.add_argument('--lr', type=float, default=0.025)
ap.add_argument('--batch-walks', type=int, default=65536)
ap.add_argument('--arms', default='one,global,local,owner,synced,hybrid')
ap.add_argument('--merge', default='stable', help='synced arm: merge rules to try (mean | sum | weight), comma separated')
ap.add_argument('--owner-micro-walks', type=int, default=0, help='owner-computes arm: interleave positives / negatives in slices of this many walks')
a = ap.parse_args()

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
nat.load()
csr, labels = sbm_graph_device(a.nodes, a.edges, a.blocks, a.p_in, seed=0, device=dev)
vocab = a.nodes + 1
bound = (6.0 / (vocab + a.emb)) ** 0.5
n_cen = a.walk_len - 2 * a.radius


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def train(w_in, w_out, r, g, local_neg, owner=False, synced=False, merge='sum', micro=None):
    """Epochs of walks -> fused update; rank r of g takes walks r, r+g, ... of every batch."""
    gen = torch.Generator()
    gen.manual_seed(1)
    stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)
    losses = []
    for epoch in range(a.epochs):
        order = torch.randperm(a.nodes, generator=gen).to(torch.int32).repeat_interleave(a.walks_per_node)
        stats.zero_()
        for lo in range(0, order.numel(), a.batch_walks):
            starts = order[lo:lo + a.batch_walks][r::g].contiguous().to(dev)
            base = epoch * order.numel() + lo + r
            walks = nat.walk(csr, starts, a.walk_len, 1.0, 0.5, True, nat.RULE_REFERENCE, seed=7, walk_id_base=base, walk_id_stride=g)
            lr = a.lr * (1.0 - 0.5 * epoch / max(a.epochs, 1))
            if owner and walks.shape[0] * g == min(a.batch_walks, order.numel() - lo):      # same decision on every rank
                sgns_update_walks_owner_computes(w_in, w_out, walks, a.radius, a.neg, 1, lr, 11, (epoch * order.numel() + lo) * n_cen, r, g,
                                                 stats=stats, micro_walks=(a.owner_micro_walks or None) if micro is None else micro)
            else:
                nat.sgns_update_walks(w_in, w_out, walks, a.radius, a.neg, 1, lr, seed=11, centre_id_base=base * n_cen, stats=stats,
                                      local_negatives=local_neg)
            mg = merge
            if isinstance(merge, str) and merge.startswith('warmup:'):      # large-batch practice: ramp the merge weight from 2/G to 1 over K syncs
                k = float(merge.split(':')[1])
                n_sync = epoch * (-(-order.numel() // a.batch_walks)) + lo // a.batch_walks
                mg = min(1.0, 2.0 / g + (1.0 - 2.0 / g) * n_sync / k)
            if synced is True:
                sync_replicated([w_in, w_out], merge=mg)
            elif synced == 'hybrid':
                sync_replicated([w_in], merge=mg)
        s = stats.tolist()
        losses.append((s[0] + s[1]) / max(s[4], 1))
    return losses


def evaluate(w_in_dense):
    from sklearn.linear_model import LogisticRegression
    rng = np.random.default_rng(0)
    idx = rng.permutation(a.nodes)[:40000]
    x = w_in_dense[idx + 1].cpu().numpy()
    y = labels[torch.from_numpy(idx).to(dev)].cpu().numpy()
    clf = LogisticRegression(max_iter=300)
    clf.fit(x[:20000], y[:20000])
    return float((clf.predict(x[20000:]) == y[20000:]).mean())


out = {'config': vars(a), 'world': world}
# ---- 1 GPU, reference negatives (rank 0 only) ---------------------------------------------------------------------------
arms = set(a.arms.split(','))
if rank == 0 and 'one' in arms:
    w_in = torch.empty((vocab, a.emb), device=dev); w_out = torch.empty((vocab, a.emb), device=dev)
    nat.table_fill_uniform(w_in, bound, 101); nat.table_fill_uniform(w_out, bound, 102)
    t0 = time.time()
    losses = train(w_in, w_out, 0, 1, False)
    torch.cuda.synchronize()
    secs = time.time() - t0
    out['one_gpu_global_negatives'] = {'accuracy': evaluate(w_in), 'epoch_losses': losses, 'seconds': secs}
    print('1 GPU', out['one_gpu_global_negatives'], flush=True)
    del w_in, w_out
barrier()
# ---- G GPUs, striped tables ---------------------------------------------------------------------------------------------
if world > 1:
    ex = make_exchange(rank, world)
    plan = [('global', 'striped_global_negatives', False, False, False, None), ('local', 'striped_local_negatives', True, False, False, None),
            ('owner', 'striped_owner_computes_negatives', False, True, False, None)]
    plan += [('synced', f'synced_copies_global_negatives_merge_{mg}', False, False, True, mg) for mg in a.merge.split(',')]
    plan += [('hybrid', 'hybrid_w_in_synced_sum_w_out_striped_owner_computes', False, True, 'hybrid', 'sum')]
    plan += [('ownermicro', 'striped_owner_computes_micro4096', False, True, False, None)]
    for key, name, local_neg, owner, synced, merge in plan:
        if key not in arms:
            continue
        s_in = (ReplicatedTable if synced else ShardedTable)(vocab, a.emb, dev, rank, world, ex)
        s_out = (ReplicatedTable if synced is True else ShardedTable)(vocab, a.emb, dev, rank, world, ex)
        s_in.fill_uniform(bound, 101); s_out.fill_uniform(bound, 102)
        barrier()
        t0 = time.time()
        losses = train(s_in, s_out, rank, world, local_neg, owner, synced, merge, micro=(4096 // world) if key == 'ownermicro' else None)
        barrier()
        secs = time.time() - t0
        if rank == 0:
            out[name] = {'accuracy': evaluate(s_in.to_tensor()), 'epoch_losses_rank0': losses, 'seconds': secs, 'gpus': world}
            print(name, out[name], flush=True)
        barrier()
        s_in.close(); s_out.close()
    ex.close()
if rank == 0:
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'multi_gpu_accuracy.json'), 'w'), indent=1)
if world > 1:
    dist.destroy_process_group()

"""Worker of tests/test_gpu_sharded.py::test_two_gpus_train_one_pair_of_tables_over_nvlink (run under torchrun, one process
per GPU).  Every rank checks the SAME striped tables against the oracle after all ranks have updated them concurrently."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

import philox_ref
from oracle import sgns_oracle
from shallow_encoders import _native as nat
from shallow_encoders.word2vec.sharded import ShardedTable, local_rows, local_to_global, make_exchange


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    nat.load()
    ex = make_exchange(rank, world)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    emb, radius, k, offset, n_seq = 128, 2, 3, 1, 6
    vocab = 4096 * 7 + 13
    s_in = ShardedTable(vocab, emb, dev, rank, world, ex)
    s_out = ShardedTable(vocab, emb, dev, rank, world, ex)
    sr = s_in.stripe_rows
    assert sr == 4096 and s_in.n_stripes == 8

    # 1. fill: every rank writes its own stripes; every rank then reads the same full table over NVLink
    s_in.fill_uniform(0.3, 5); s_out.fill_uniform(0.3, 6)
    barrier()
    d_in = torch.empty((vocab, emb), device=dev); d_out = torch.empty((vocab, emb), device=dev)
    nat.table_fill_uniform(d_in, 0.3, 5); nat.table_fill_uniform(d_out, 0.3, 6)
    assert torch.equal(s_in.to_tensor(), d_in) and torch.equal(s_out.to_tensor(), d_out), 'peer stripes differ from the dense fill'
    w_in0, w_out0 = d_in.cpu().numpy(), d_out.cpu().numpy()
    barrier()

    length = 2 * radius + 1
    rng = np.random.default_rng(99)                       # same plan on every rank
    tokens_all = rng.permutation(vocab - offset)[:world * n_seq * length].reshape(world, n_seq, length).astype(np.int32)
    lr = 0.025

    def run_case(local_neg):
        for seed in range(100, 2000):
            negs = []
            for r in range(world):
                cid = np.arange(n_seq) + 1000 * (r + 1)
                if local_neg:
                    j = philox_ref.negatives(seed, cid, 2 * radius, k, local_rows(vocab, sr, world, r))
                    negs.append(local_to_global(j, sr, world, r))
                else:
                    negs.append(philox_ref.negatives(seed, cid, 2 * radius, k, vocab))
            neg = np.concatenate(negs)
            inputs, targets = sgns_oracle.windows_from_walks(tokens_all.reshape(-1, length).astype(np.int64), radius, offset)
            allrows = np.concatenate([targets.ravel(), neg.ravel()])
            if len(np.unique(allrows)) == allrows.size:
                break
        else:
            raise AssertionError('no collision-free seed')
        allr = torch.arange(vocab, device=dev)
        if rank == 0:
            s_in.scatter(allr, d_in); s_out.scatter(allr, d_out)
        barrier()
        st = nat.sgns_update_walks(s_in, s_out, torch.from_numpy(tokens_all[rank]).to(dev), radius, k, offset, lr, seed,
                                   centre_id_base=1000 * (rank + 1), local_negatives=local_neg)
        assert st['pairs'] == n_seq * 2 * radius
        barrier()
        rows = np.unique(np.concatenate([allrows, inputs.ravel()]))
        remap = {int(r): i for i, r in enumerate(rows)}
        rm = np.vectorize(remap.get)
        want_in, want_out, _ = sgns_oracle.sgd_step(w_in0[rows].astype(np.float64), w_out0[rows].astype(np.float64), rm(inputs),
                                                    rm(targets), rm(neg), lr * len(inputs) * 2 * radius)
        got_in, got_out = s_in.to_tensor().cpu().numpy(), s_out.to_tensor().cpu().numpy()
        np.testing.assert_allclose(got_in[rows], want_in, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(got_out[rows], want_out, rtol=1e-4, atol=1e-5)
        untouched = np.setdiff1d(np.arange(vocab), rows)
        assert np.array_equal(got_in[untouched], w_in0[untouched]) and np.array_equal(got_out[untouched], w_out0[untouched])
        assert np.abs(got_out[rows] - w_out0[rows]).max() > 1e-4
        barrier()

    # 2. / 3. every rank updates its own sequences; rows live on all GPUs (global negatives) or negatives on the own shard
    run_case(local_neg=False)
    run_case(local_neg=True)

    # 3b. owner-computes negatives: reference-exact GLOBAL draw, every rank processes the negatives it owns for the centres of
    #     all ranks (walks all-gathered); same pair updates as one GPU on the concatenated batch
    from shallow_encoders.word2vec.sharded import sgns_update_walks_owner_computes
    n_tot = world * n_seq
    for seed in range(100, 2000):
        neg = philox_ref.negatives(seed, np.arange(n_tot) + 7000, 2 * radius, k, vocab)
        inputs, targets = sgns_oracle.windows_from_walks(tokens_all.reshape(-1, length).astype(np.int64), radius, offset)
        allrows = np.concatenate([targets.ravel(), neg.ravel()])
        if len(np.unique(allrows)) == allrows.size:
            break
    else:
        raise AssertionError('no collision-free seed')
    allr = torch.arange(vocab, device=dev)
    if rank == 0:
        s_in.scatter(allr, d_in); s_out.scatter(allr, d_out)
    barrier()
    small = 1e-3
    stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)
    sgns_update_walks_owner_computes(s_in, s_out, torch.from_numpy(tokens_all[rank]).to(dev), radius, k, offset, small, seed, 7000,
                                     rank, world, stats=stats)
    barrier()
    first_pass = s_out.to_tensor().clone()
    barrier()
    # the same step interleaved in micro-batches of 4 walks (one all-gather, positives / negatives alternating) lands on the same tables
    if rank == 0:
        s_in.scatter(allr, d_in); s_out.scatter(allr, d_out)
    barrier()
    sgns_update_walks_owner_computes(s_in, s_out, torch.from_numpy(tokens_all[rank]).to(dev), radius, k, offset, small, seed, 7000,
                                     rank, world, micro_walks=4)
    barrier()
    assert float((s_out.to_tensor() - first_pass).abs().max()) < 2e-6
    barrier()
    dist.all_reduce(stats)
    assert stats[4].item() == n_tot * 2 * radius and stats[5].item() == n_tot * 2 * radius * k, stats.tolist()
    rows = np.unique(np.concatenate([allrows, inputs.ravel()]))
    remap = {int(r): i for i, r in enumerate(rows)}
    rm = np.vectorize(remap.get)
    want_in, want_out, _ = sgns_oracle.sgd_step(w_in0[rows].astype(np.float64), w_out0[rows].astype(np.float64), rm(inputs), rm(targets),
                                                rm(neg), small * len(inputs) * 2 * radius)
    got_in, got_out = s_in.to_tensor().cpu().numpy(), s_out.to_tensor().cpu().numpy()
    np.testing.assert_allclose(got_in[rows], want_in, rtol=0, atol=2e-6)
    np.testing.assert_allclose(got_out[rows], want_out, rtol=0, atol=2e-6)
    assert np.abs(got_out[rows] - w_out0[rows]).max() > 1e-4
    untouched = np.setdiff1d(np.arange(vocab), rows)
    assert np.array_equal(got_in[untouched], w_in0[untouched]) and np.array_equal(got_out[untouched], w_out0[untouched])
    barrier()

    # 4. contention: ALL ranks push the same update into the same rows at the same time; system-scope reductions must
    #    accumulate every contribution (to first order in lr: world x the single update)
    allr = torch.arange(vocab, device=dev)
    if rank == 0:
        s_in.scatter(allr, d_in); s_out.scatter(allr, d_out)
    barrier()
    tok = torch.from_numpy(tokens_all[0]).to(dev)
    small_lr = 5e-4
    for _ in range(8):
        nat.sgns_update_walks(s_in, s_out, tok, radius, k, offset, small_lr, 7, centre_id_base=0)
    barrier()
    t_in, t_out = d_in.clone(), d_out.clone()
    for _ in range(8 * world):
        nat.sgns_update_walks(t_in, t_out, tok, radius, k, offset, small_lr, 7, centre_id_base=0)
    torch.cuda.synchronize()
    got = s_out.to_tensor()
    delta_ref = (t_out - d_out)
    delta_got = (got - d_out)
    scale = float(delta_ref.abs().max())
    assert scale > 1e-4
    assert float((delta_got - delta_ref).abs().max()) < 0.05 * scale, (float((delta_got - delta_ref).abs().max()), scale)
    barrier()

    # 5. the host-buffer step on sharded tables with local negatives
    from helpers import random_csr
    from shallow_encoders.graph.csr import CSRGraph
    rowptr, col = random_csr(vocab - 1, 3 * vocab, 5, sort_rows=True)
    csr = CSRGraph.from_arrays(rowptr, col, device=dev)
    L = 12
    starts = torch.arange(rank, vocab - 1, world, dtype=torch.int32)[:2000].contiguous()
    scratch = {'starts': torch.empty(starts.numel(), dtype=torch.int32, device=dev),
               'walks': torch.empty((starts.numel(), L), dtype=torch.int32, device=dev),
               'stats': torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)}
    stats_host = torch.zeros(nat.STATS_LEN, dtype=torch.float64)
    nat.host_walk_sgns_step(csr, starts, L, 0.5, 2.0, True, nat.RULE_REFERENCE, 3, rank * 2000, s_in, s_out, radius, k, 1, 0.01,
                            scratch, stats_host, local_negatives=True)
    assert stats_host[4].item() == starts.numel() * (L - 2 * radius) * 2 * radius and np.isfinite(stats_host.numpy()).all()
    barrier()

    # 6. the NCCL all-to-all baseline (row % world sharding) is synchronous mini-batch SGD over the union of all ranks' pairs
    from shallow_encoders.word2vec.row_exchange import RowShardedTables
    rs = RowShardedTables(vocab, emb, rank, world, dev)
    rs.load_full('in', d_in); rs.load_full('out', d_out)
    tok_r = torch.from_numpy(tokens_all[rank]).to(dev)
    b_r = n_seq * 1
    rs.step(tok_r, radius, k, offset, lr, 23, draw_id_base=5000 * (rank + 1), micro_walks=n_seq)
    barrier()
    inputs, targets = sgns_oracle.windows_from_walks(tokens_all.reshape(-1, length).astype(np.int64), radius, offset)
    neg = np.concatenate([philox_ref.draws(23, b_r * 2 * radius * k, vocab, base=5000 * (r + 1)).reshape(b_r, 2 * radius, k) for r in range(world)])
    want_in, want_out, _ = sgns_oracle.sgd_step(w_in0.astype(np.float64), w_out0.astype(np.float64), inputs, targets, neg,
                                                lr * len(inputs) * 2 * radius)     # each rank applies lr * (sum of ITS pairs' gradients)
    np.testing.assert_allclose(rs.gather_full('in').cpu().numpy(), want_in, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rs.gather_full('out').cpu().numpy(), want_out, rtol=1e-4, atol=1e-5)
    assert rs.exchanged_bytes > 0
    barrier()
    # 7. ReplicatedTable (csrc/replica.cu): every rank trains its own working copy with the GLOBAL negative draw, then the fused
    #    reduce-scatter + all-gather over peer memory leaves start + sum_g (copy_g - start) in every copy
    from shallow_encoders.word2vec.sharded import ReplicatedTable, sync_replicated
    r_in = ReplicatedTable(vocab, emb, dev, rank, world, ex); r_out = ReplicatedTable(vocab, emb, dev, rank, world, ex)
    r_in.fill_uniform(0.3, 5); r_out.fill_uniform(0.3, 6)
    barrier()
    assert torch.equal(r_in.to_tensor(), d_in) and torch.equal(r_out.view(), d_out)
    walks_r = torch.from_numpy(np.random.default_rng(1000 + rank).integers(0, vocab - offset, (256, 9)).astype(np.int32)).to(dev)
    for it in range(2):
        before_in, before_out = r_in.to_tensor().double(), r_out.to_tensor().double()
        st = nat.sgns_update_walks(r_in, r_out, walks_r, radius, k, offset, 0.05, seed=50 + it, centre_id_base=(it * world + rank) * 256 * 5)
        assert st['pairs'] == 256 * 5 * 4
        mine_in, mine_out = r_in.to_tensor().double() - before_in, r_out.to_tensor().double() - before_out
        assert float(mine_in.abs().max()) > 1e-4
        dist.all_reduce(mine_in); dist.all_reduce(mine_out)                      # NCCL cross-check of the summed updates
        sync_replicated([r_in, r_out], merge='sum')
        torch.cuda.synchronize()
        assert float((r_in.to_tensor().double() - (before_in + mine_in)).abs().max()) < 2e-6
        assert float((r_out.to_tensor().double() - (before_out + mine_out)).abs().max()) < 2e-6
        flat = r_in.to_tensor().reshape(-1)
        lo, hi = nat.replica_chunk(vocab * emb, world, rank)
        hi = max(lo, min(hi, vocab * emb))
        assert torch.equal(r_in.master[:hi - lo], flat[lo:hi])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, flat), 'working copies differ between ranks after the sync'
    barrier()
    r_in.close(); r_out.close()
    s_in.close(); s_out.close(); ex.close()
    dist.destroy_process_group()
    print(f'MGPU_OK rank {rank}/{world}', flush=True)


if __name__ == '__main__':
    main()

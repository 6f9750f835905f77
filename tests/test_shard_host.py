"""Host-side logic of the multi-GPU sharding (no GPU): stripe ownership arithmetic against the C ABI, the unix-socket
file-descriptor exchange between two ranks, and the gloo world-size-2 rendezvous that names the sockets."""
import ctypes
import multiprocessing as mp
import os

import numpy as np
import pytest

from shallow_encoders import _native as nat
from shallow_encoders.word2vec import sharded


@pytest.mark.parametrize('vocab,stripe_rows,world', [(1, 4, 2), (4, 4, 2), (5, 4, 2), (8, 4, 2), (9, 4, 2), (100, 8, 4), (4096 * 5 + 100, 4096, 2),
                                                     (10_000_001, 4096, 8), (267_736, 4096, 4), (35, 16, 8), (1000, 7, 3)])
def test_local_rows_matches_c_abi_and_partitions_the_table(vocab, stripe_rows, world):
    lib = nat.load()
    total = 0
    seen = np.zeros(vocab, dtype=np.int32)
    for rank in range(world):
        spec = nat.ShardSpec(world, rank, stripe_rows, 0, 0)
        n = ctypes.c_int64()
        assert lib.se_shard_local_rows(vocab, ctypes.byref(spec), ctypes.byref(n)) == 0
        want = sharded.local_rows(vocab, stripe_rows, world, rank)
        assert n.value == want
        rows = sharded.local_to_global(np.arange(want, dtype=np.int64), stripe_rows, world, rank)
        assert (rows < vocab).all() and (np.diff(rows) > 0).all()
        assert ((rows // stripe_rows) % world == rank).all()       # every mapped row lives on this rank
        seen[rows] += 1
        total += want
    assert total == vocab and (seen == 1).all()                    # the shards partition [0, vocab)


def test_shard_spec_validation():
    lib = nat.load()
    n = ctypes.c_int64()
    for bad in (nat.ShardSpec(0, 0, 4, 0, 0), nat.ShardSpec(2, 2, 4, 0, 0), nat.ShardSpec(2, -1, 4, 0, 0), nat.ShardSpec(2, 0, 0, 0, 0)):
        assert lib.se_shard_local_rows(10, ctypes.byref(bad), ctypes.byref(n)) == -1
        assert b'shard spec' in lib.se_last_error()


def _fd_worker(rank, world, token, barrier, q):
    try:
        ex = sharded.FdExchange(rank, world, token, barrier.wait)
        # every rank owns pipes; it sends the WRITE ends to every peer in two batches and reads what the peers wrote
        pipes = [os.pipe() for _ in range(5)]
        for lo, hi in ((0, 3), (3, 5)):
            for peer in range(world):
                if peer != rank:
                    ex.send(peer, [w for _r, w in pipes[lo:hi]])
            for peer in range(world):
                if peer != rank:
                    for i, fd in enumerate(ex.recv(peer, hi - lo)):
                        os.write(fd, f'{rank}->{peer}:{lo + i};'.encode())
                        os.close(fd)
        barrier.wait()
        got = []
        for r, w in pipes:
            os.close(w)
            got.append(os.read(r, 4096).decode())
            os.close(r)
        ex.close()
        q.put((rank, got))
    except Exception as e:   # noqa: BLE001
        q.put((rank, repr(e)))


@pytest.mark.parametrize('world', [2, 3])
def test_fd_exchange_between_ranks(world):
    ctx = mp.get_context('fork')
    barrier, q = ctx.Barrier(world), ctx.Queue()
    token = f'test{os.getpid()}_{world}'
    procs = [ctx.Process(target=_fd_worker, args=(r, world, token, barrier, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    for rank in range(world):
        assert isinstance(res[rank], list), res[rank]
        for i, text in enumerate(res[rank]):
            parts = sorted(x for x in text.split(';') if x)
            assert parts == sorted(f'{peer}->{rank}:{i}' for peer in range(world) if peer != rank)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        dist.init_process_group('gloo', rank=rank, world_size=world)
        ex = sharded.make_exchange(rank, world)
        r, w = os.pipe()
        peer = 1 - rank
        ex.send(peer, [w])
        (fd,) = ex.recv(peer, 1)
        os.write(fd, b'hello from %d' % rank)
        os.close(fd)
        dist.barrier()
        os.close(w)
        msg = os.read(r, 100)
        ex.close()
        dist.destroy_process_group()
        q.put((rank, msg))
    except Exception as e:   # noqa: BLE001
        q.put((rank, repr(e)))


def test_make_exchange_over_gloo_world_size_2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res[0] == b'hello from 1' and res[1] == b'hello from 0', res


def _row_exchange_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from shallow_encoders.word2vec.row_exchange import RowShardedTables
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        dist.init_process_group('gloo', rank=rank, world_size=world)
        vocab, emb = 101, 4
        full = torch.arange(vocab * emb, dtype=torch.float32).reshape(vocab, emb)
        t = RowShardedTables(vocab, emb, rank, world, 'cpu')
        t.load_full('out', full)
        assert t.local['out'].shape[0] == len(range(rank, vocab, world))
        g = torch.Generator().manual_seed(100 + rank)
        ids = torch.unique(torch.randint(0, vocab, (40 + 7 * rank,), generator=g))
        plan = t.plan(ids)
        rows = t.fetch('out', plan)
        assert torch.equal(rows, full[ids]), 'fetched rows differ'
        grads = torch.full((ids.numel(), emb), float(rank + 1))
        t.push('out', plan, grads, lr=0.5)
        dist.barrier()
        got = t.gather_full('out')
        # expected: every rank r subtracted 0.5 * (r + 1) from the rows IT asked for
        want = full.clone()
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            ids_r = torch.unique(torch.randint(0, vocab, (40 + 7 * r,), generator=gr))
            want[ids_r] -= 0.5 * (r + 1)
        assert torch.equal(got, want), 'pushed gradients were not applied at the owners'
        c, x = RowShardedTables.windows(torch.arange(16, dtype=torch.int32).reshape(2, 8), 3, 1)
        assert c.reshape(-1).tolist() == [4, 5, 12, 13] and x[0].tolist() == [1, 2, 3, 5, 6, 7]     # torch_dataset.py:302-306
        dist.destroy_process_group()
        q.put((rank, 'ok'))
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


def test_row_exchange_fetch_and_push_over_gloo_world_size_2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29300 + os.getpid() % 300
    procs = [ctx.Process(target=_row_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res == {0: 'ok', 1: 'ok'}, res


def _owner_step_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from shallow_encoders import _native as nat
    from shallow_encoders.word2vec import sharded
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        dist.init_process_group('gloo', rank=rank, world_size=world)
        calls = []
        # the kernels are replaced by recorders: this test covers the HOST side of the multi-GPU headline step (all-gather layout, keys)
        nat.sgns_update_pairs_owned = lambda w_in, w_out, tokens, radius, n_neg, row_offset, lr, seed, centre_id_base=0, alias=None, stats=None, \
            positives=True, scratch=None: calls.append(('pairs', tokens.clone(), radius, n_neg, row_offset, lr, seed, centre_id_base, positives))
        n_walks, length, radius = 6, 9, 2
        mine = (torch.arange(n_walks * length, dtype=torch.int32).reshape(n_walks, length) + 1000 * rank)
        sharded.sgns_update_walks_owner_computes('W_IN', 'W_OUT', mine, radius, 3, 1, 0.01, seed=77, centre_id_base=5000, rank=rank, world=world)
        sharded.sgns_update_walks_owner_computes('W_IN', 'W_OUT', mine, radius, 3, 1, 0.01, seed=77, centre_id_base=5000, rank=rank, world=world,
                                                 micro_walks=4)
        dist.destroy_process_group()
        q.put((rank, [(c[0], c[1].tolist(), *c[2:]) for c in calls]))
    except Exception:   # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


def test_owner_computes_step_gathers_the_same_batch_on_every_rank_over_gloo_world_size_2():
    """The multi-GPU headline step (`sgns_update_walks_owner_computes`, grouped): after ONE all-gather every rank hands the SAME token matrix
    (rank r's walks in block r) and the SAME Philox keys to `se_sgns_update_pairs_owned`, so the ranks partition one draw; with micro-batches
    the slices of every rank's block follow with their own centre id bases.  Kernels are recorders here (world size 2, gloo, CPU)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29000 + os.getpid() % 300
    procs = [ctx.Process(target=_owner_step_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert isinstance(res[0], list) and isinstance(res[1], list), res
    assert res[0] == res[1], 'ranks disagree on the gathered batch or its keys'
    n_walks, length, radius, n_cen = 6, 9, 2, 5
    block = lambda r: [[1000 * r + w * length + j for j in range(length)] for w in range(n_walks)]      # noqa: E731
    first = res[0][0]
    assert first == ('pairs', block(0) + block(1), radius, 3, 1, 0.01, 77, 5000, True)
    micro = res[0][1:]
    want = []
    for lo, hi in ((0, 4), (4, 6)):
        for r in range(2):
            want.append(('pairs', block(r)[lo:hi], radius, 3, 1, 0.01, 77, 5000 + (r * n_walks + lo) * n_cen, True))
    assert micro == want


def test_replica_chunks_partition_the_table_in_float4_units():
    """se_replica_chunk (csrc/replica.cu): rank r owns a contiguous element range, the ranges tile [0, round_up4(n)) exactly."""
    from shallow_encoders import _native as nat
    nat.load()
    for n_elems in (0, 1, 70, 4096 * 128 + 12, 10_000_001 * 128):
        for world in (1, 2, 3, 8):
            edges = [nat.replica_chunk(n_elems, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == -(-n_elems // 4) * 4
            for (lo, hi), (lo2, _hi2) in zip(edges, edges[1:]):
                assert lo <= hi == lo2 and lo % 4 == 0
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(s for s in sizes if s or True) <= max(sizes)      # no chunk larger than ceil(n4 / world) float4s
            assert max(sizes) == -(-(-(-n_elems // 4)) // world) * 4 or n_elems == 0

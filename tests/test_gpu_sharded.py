"""Sharded tables (csrc/shard.cu, shallow_encoders/word2vec/sharded.py): the fused SGNS kernel on a table mapped through
CUDA virtual memory management gives the same results as on a torch tensor; local negatives land on rows the rank owns and
match the numpy Philox restatement; two real GPUs (when the box has them) train ONE pair of tables over NVLink."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import philox_ref
from helpers import cuda_device
from oracle import sgns_oracle
from shallow_encoders import _native as nat
from shallow_encoders.word2vec.sharded import ReplicatedTable, ShardedTable, local_rows, local_to_global

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_fill_is_independent_of_sharding_and_rows_round_trip():
    dev = cuda_device()
    vocab, emb = 4096 * 3 + 77, 128
    dense = torch.empty((vocab, emb), dtype=torch.float32, device=dev)
    nat.table_fill_uniform(dense, 0.25, seed=11)
    assert float(dense.abs().max()) <= 0.25 and float(dense.abs().max()) > 0.2 and abs(float(dense.mean())) < 1e-3
    # numpy restatement: element 4v+j = (2*u01(word j of Philox(seed; v, 0, 0x50000000)) - 1) * bound
    words = philox_ref.philox(11, np.arange(64, dtype=np.uint64), 0, 0x50000000)
    want = np.stack([(np.float32(2.0) * philox_ref.u01(w) - np.float32(1.0)) * np.float32(0.25) for w in words], axis=1).reshape(-1)
    assert np.array_equal(dense.reshape(-1)[:256].cpu().numpy(), want.astype(np.float32))
    one = ShardedTable(vocab, emb, dev)                                     # world 1: every stripe local
    one.fill_uniform(0.25, seed=11)
    assert torch.equal(one.to_tensor(), dense)
    # simulated 2-way sharding: each "rank" fills only the stripes it owns
    parts = []
    for rank in range(2):
        t = ShardedTable(vocab, emb, dev, rank=rank, world=2, simulate=True)
        assert t.stripe_rows == 4096 and t.n_stripes == 4
        t.scatter(torch.arange(vocab, device=dev), torch.zeros((vocab, emb), device=dev))
        t.fill_uniform(0.25, seed=11)
        full = t.to_tensor()
        own = t.owned_rows()
        assert own.numel() == local_rows(vocab, 4096, 2, rank)
        assert torch.equal(full[own], dense[own])
        mask = torch.ones(vocab, dtype=torch.bool, device=dev)
        mask[own] = False
        assert float(full[mask].abs().max()) == 0.0                        # rows of the other rank were not written
        parts.append(own)
        t.close()
    assert torch.equal(torch.sort(torch.cat(parts)).values, torch.arange(vocab, device=dev))
    rows = torch.tensor([0, 5, vocab - 1, 4096, 4095], device=dev)
    src = torch.randn((5, emb), device=dev)
    one.scatter(rows, src)
    assert torch.equal(one.gather(rows), src)
    one.close()


def test_empty_inputs_are_no_ops_on_the_new_entry_points():
    dev = cuda_device()
    t = ShardedTable(5000, 128, dev)
    t.fill_uniform(0.1, 3)
    before = t.to_tensor().clone()
    other = ShardedTable(5000, 128, dev)
    st = nat.sgns_update_walks(t, other, torch.empty((0, 12), dtype=torch.int32, device=dev), 2, 3, 1, 0.1, 1)
    assert st['pairs'] == 0 and torch.equal(t.to_tensor(), before)
    assert t.gather(torch.empty(0, dtype=torch.int64, device=dev)).shape == (0, 128)
    t.scatter(torch.empty(0, dtype=torch.int64, device=dev), torch.empty((0, 128), device=dev))
    dense = before.clone()
    assert nat.edge_features(dense, torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int64, device=dev), 'hadamard').shape == (0, 128)
    # sequences shorter than a window are refused exactly like the reference's collate assert (torch_dataset.py:298)
    with pytest.raises(AssertionError, match='Text is too short'):
        nat.sgns_update_walks(t, other, torch.zeros((4, 4), dtype=torch.int32, device=dev), 2, 3, 1, 0.1, 1)
    # mixing a striped and a torch table is refused
    with pytest.raises(ValueError):
        nat.sgns_update_walks(t, dense, torch.zeros((4, 12), dtype=torch.int32, device=dev), 2, 3, 1, 0.1, 1)
    t.close(); other.close()


def _collision_free_case(rng, emb, radius, k, n_seq, vocab, offset, neg_fn):
    length = 2 * radius + 1
    tokens = rng.permutation(vocab - offset)[:n_seq * length].reshape(n_seq, length).astype(np.int32)
    inputs, targets = sgns_oracle.windows_from_walks(tokens.astype(np.int64), radius, offset)
    for seed in range(77, 600):
        neg = neg_fn(seed)
        allrows = np.concatenate([targets.ravel(), neg.ravel()])
        if len(np.unique(allrows)) == allrows.size:
            return tokens, inputs, targets, neg, seed
    raise AssertionError('no collision-free seed')


@pytest.mark.parametrize('emb,radius,k', [(128, 5, 5), (128, 2, 3), (64, 2, 5), (256, 3, 2)])
def test_fused_update_on_vmm_table_equals_torch_table(emb, radius, k):
    dev = cuda_device()
    rng = np.random.default_rng(21)
    vocab, offset, n_seq = 50000, 1, 6
    tokens, inputs, targets, neg, seed = _collision_free_case(
        rng, emb, radius, k, n_seq, vocab, offset, lambda s: philox_ref.negatives(s, np.arange(n_seq) + 1000, 2 * radius, k, vocab))
    w_in = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    w_out = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    t_in, t_out = _t(w_in, dev), _t(w_out, dev)
    st_ref = nat.sgns_update_walks(t_in, t_out, _t(tokens, dev), radius, k, offset, 0.025, seed, centre_id_base=1000)
    s_in, s_out = ShardedTable(vocab, emb, dev), ShardedTable(vocab, emb, dev)
    allrows = torch.arange(vocab, device=dev)
    s_in.scatter(allrows, _t(w_in, dev)); s_out.scatter(allrows, _t(w_out, dev))
    st = nat.sgns_update_walks(s_in, s_out, _t(tokens, dev), radius, k, offset, 0.025, seed, centre_id_base=1000)
    assert torch.equal(s_in.to_tensor(), t_in) and torch.equal(s_out.to_tensor(), t_out)    # collision-free: deterministic
    assert st == st_ref and st['pairs'] == n_seq * 2 * radius
    s_in.close(); s_out.close()


@pytest.mark.parametrize('rank', [0, 1])
@pytest.mark.parametrize('use_alias', [False, True])
def test_local_negatives_stay_on_the_owning_shard_and_match_the_oracle(rank, use_alias):
    """world = 2 simulated on one GPU: negatives are local ids of `rank` mapped to its stripes (restated in numpy), and the
    update equals mini-batch SGD on exactly those rows."""
    dev = cuda_device()
    rng = np.random.default_rng(31 + rank)
    emb, radius, k, n_seq, offset, world = 128, 2, 4, 8, 1, 2
    vocab = 4096 * 5 + 100
    s_in = ShardedTable(vocab, emb, dev, rank=rank, world=world, simulate=True)
    s_out = ShardedTable(vocab, emb, dev, rank=rank, world=world, simulate=True)
    sr = s_in.stripe_rows
    n_local = local_rows(vocab, sr, world, rank)
    alias = None
    prob = ali = None
    if use_alias:
        counts = rng.integers(1, 50, n_local).astype(np.float64)
        alias = nat.alias_build(counts, 0.75, dev)
        prob, ali = alias['prob'].cpu().numpy(), alias['alias'].cpu().numpy()

    def neg_fn(seed):
        j = philox_ref.negatives(seed, np.arange(n_seq) + 500, 2 * radius, k, n_local, prob, ali)
        return local_to_global(j, sr, world, rank)

    tokens, inputs, targets, neg, seed = _collision_free_case(rng, emb, radius, k, n_seq, vocab, offset, neg_fn)
    assert ((neg // sr) % world == rank).all() and (neg < vocab).all()
    w_in = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    w_out = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    allrows = torch.arange(vocab, device=dev)
    s_in.scatter(allrows, _t(w_in, dev)); s_out.scatter(allrows, _t(w_out, dev))
    lr = 0.025
    st = nat.sgns_update_walks(s_in, s_out, _t(tokens, dev), radius, k, offset, lr, seed, centre_id_base=500, alias=alias,
                               local_negatives=True)
    rows = np.unique(np.concatenate([targets.ravel(), neg.ravel(), inputs.ravel()]))
    remap = {int(r): i for i, r in enumerate(rows)}
    rm = np.vectorize(remap.get)
    want_in, want_out, o = sgns_oracle.sgd_step(w_in[rows].astype(np.float64), w_out[rows].astype(np.float64), rm(inputs), rm(targets),
                                                rm(neg), lr * n_seq * 2 * radius)
    got_in, got_out = s_in.to_tensor().cpu().numpy(), s_out.to_tensor().cpu().numpy()
    np.testing.assert_allclose(got_in[rows], want_in, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(got_out[rows], want_out, rtol=1e-4, atol=1e-5)
    assert abs(st['loss'] - o['loss']) <= 1e-4 * abs(o['loss'])
    changed = np.nonzero(np.abs(got_out - w_out).max(axis=1) > 0)[0]
    assert set(changed) <= set(rows.tolist())
    neg_changed = np.setdiff1d(changed, targets.ravel())
    assert ((neg_changed // sr) % world == rank).all()                     # only this rank's stripes received negatives
    # plain stores are refused on sharded tables, and torch tables cannot ask for local negatives
    with pytest.raises(AssertionError, match='plain-store'):
        nat.sgns_update_walks(s_in, s_out, _t(tokens, dev), radius, k, offset, lr, seed, flags=nat.SCATTER_STORE)
    with pytest.raises(ValueError):
        nat.sgns_update_walks(_t(w_in, dev), _t(w_out, dev), _t(tokens, dev), radius, k, offset, lr, seed, local_negatives=True)
    s_in.close(); s_out.close()


@pytest.mark.parametrize('grouped', [True, False])
@pytest.mark.parametrize('use_alias,emb,world', [(False, 128, 2), (True, 128, 2), (False, 96, 2), (False, 128, 3), (True, 128, 4)])
def test_owner_computes_negatives_perform_the_same_pair_updates(use_alias, emb, world, grouped):
    """Positives on the home rank (window kernel, K = 0) + owner-computes negatives on every simulated rank == the pair
    updates of the ordinary fused kernel == the oracle's mini-batch SGD, with the GLOBAL negative distribution.
    world = 2 / 4 take the mask path of the ownership test (power-of-two GPU counts), world = 3 the modulo path."""
    dev = cuda_device()
    rng = np.random.default_rng(51)
    radius, k, n_seq, offset = 2, 4, 10, 1
    vocab = (4096 if emb == 128 else 16384) * 5 + 100          # 96 floats per row: 6 MiB stripes of 16384 rows
    alias = None
    prob = ali = None
    if use_alias:
        alias = nat.alias_build(rng.integers(1, 50, vocab).astype(np.float64), 0.75, dev)
        prob, ali = alias['prob'].cpu().numpy(), alias['alias'].cpu().numpy()
    tokens, inputs, targets, neg, seed = _collision_free_case(
        rng, emb, radius, k, n_seq, vocab, offset, lambda s_: philox_ref.negatives(s_, np.arange(n_seq) + 300, 2 * radius, k, vocab, prob, ali))
    w_in = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    w_out = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    lr = 1e-3          # the three launches see each other's updates of W_in: second order in lr, far below the tolerance
    s_in = ShardedTable(vocab, emb, dev, rank=0, world=world, simulate=True)
    s_out = ShardedTable(vocab, emb, dev, rank=0, world=world, simulate=True)
    allrows = torch.arange(vocab, device=dev)
    s_in.scatter(allrows, _t(w_in, dev)); s_out.scatter(allrows, _t(w_out, dev))
    tok = _t(tokens, dev)
    stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)
    if not grouped:          # round-1 form: positives on the home rank (window kernel, K = 0), owned negatives in walk order
        nat.sgns_update_walks(s_in, s_out, tok, radius, 0, offset, lr, seed, centre_id_base=300, stats=stats)
    owned_negs = 0
    for r in range(world):
        before, before_pos = stats[5].item(), stats[4].item()
        if grouped:          # every pair -- positive or negative -- on the rank that owns its W_out row, centres bucketed by row
            nat.sgns_update_pairs_owned(s_in.as_rank(r), s_out.as_rank(r), tok, radius, k, offset, lr, seed, centre_id_base=300,
                                        alias=alias, stats=stats, positives=True)
            assert stats[4].item() - before_pos == int((((targets // s_in.stripe_rows) % world) == r).sum())
        else:
            nat.sgns_update_negatives_owned(s_in.as_rank(r), s_out.as_rank(r), tok, radius, k, offset, lr, seed, centre_id_base=300,
                                            alias=alias, stats=stats, grouped=False)
        got_r = stats[5].item() - before
        assert got_r == int((((neg // s_in.stripe_rows) % world) == r).sum())                                 # exactly the rows rank r owns
        owned_negs += got_r
    assert owned_negs == neg.size and stats[4].item() == n_seq * 2 * radius
    rows = np.unique(np.concatenate([targets.ravel(), neg.ravel(), inputs.ravel()]))
    remap = {int(x): i for i, x in enumerate(rows)}
    rm = np.vectorize(remap.get)
    want_in, want_out, o = sgns_oracle.sgd_step(w_in[rows].astype(np.float64), w_out[rows].astype(np.float64), rm(inputs), rm(targets),
                                                rm(neg), lr * n_seq * 2 * radius)
    got_in, got_out = s_in.to_tensor().cpu().numpy(), s_out.to_tensor().cpu().numpy()
    np.testing.assert_allclose(got_out[rows], want_out, rtol=0, atol=2e-6)
    np.testing.assert_allclose(got_in[rows], want_in, rtol=0, atol=2e-6)
    assert np.abs(got_out[rows] - w_out[rows]).max() > 1e-4
    st = stats.tolist()
    assert abs((st[0] + st[1]) / st[4] - o['loss']) < 1e-3 * o['loss']
    untouched = np.setdiff1d(np.arange(vocab), rows)
    assert np.array_equal(got_in[untouched], w_in[untouched]) and np.array_equal(got_out[untouched], w_out[untouched])
    # the same launches as ONE ordinary fused call land on the same tables
    t_in, t_out = _t(w_in, dev), _t(w_out, dev)
    nat.sgns_update_walks(t_in, t_out, tok, radius, k, offset, lr, seed, centre_id_base=300, alias=alias)
    np.testing.assert_allclose(got_out, t_out.cpu().numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(got_in, t_in.cpu().numpy(), rtol=0, atol=2e-6)
    with pytest.raises(ValueError):
        nat.sgns_update_negatives_owned(t_in, t_out, tok, radius, k, offset, lr, seed)          # needs striped tables
    s_in.close(); s_out.close()


@pytest.mark.parametrize('grouped', [True, False])
def test_owner_computes_with_every_negative_owned_and_a_long_list(grouped):
    """world = 1 and N*K = 70 > 64: every drawn negative is owned, the per-warp list holds them all; result == the ordinary
    fused kernel on the same tokens (distinct rows, tiny lr)."""
    dev = cuda_device()
    rng = np.random.default_rng(61)
    emb, radius, k, n_seq, offset, vocab = 128, 5, 7, 4, 1, 2_000_000
    tokens = rng.permutation(vocab - offset)[:n_seq * 11].reshape(n_seq, 11).astype(np.int32)
    w = ShardedTable(vocab, emb, dev), ShardedTable(vocab, emb, dev)
    w[0].fill_uniform(0.3, 1); w[1].fill_uniform(0.3, 2)
    t_in, t_out = w[0].to_tensor().clone(), w[1].to_tensor().clone()
    tok = _t(tokens, dev)
    stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)
    if grouped:
        nat.sgns_update_pairs_owned(w[0], w[1], tok, radius, k, offset, 1e-3, 9, centre_id_base=40, stats=stats, positives=True)
    else:
        nat.sgns_update_walks(w[0], w[1], tok, radius, 0, offset, 1e-3, 9, centre_id_base=40, stats=stats)
        nat.sgns_update_negatives_owned(w[0], w[1], tok, radius, k, offset, 1e-3, 9, centre_id_base=40, stats=stats, grouped=False)
    assert stats[5].item() == n_seq * 2 * radius * k and stats[4].item() == n_seq * 2 * radius
    nat.sgns_update_walks(t_in, t_out, tok, radius, k, offset, 1e-3, 9, centre_id_base=40)
    assert float((w[1].to_tensor() - t_out).abs().max()) < 2e-6 and float((w[0].to_tensor() - t_in).abs().max()) < 2e-6
    w[0].close(); w[1].close()


@pytest.mark.parametrize('use_alias,emb,radius,k', [(False, 128, 5, 5), (True, 128, 2, 4), (False, 64, 3, 7)])
def test_grouped_owner_computes_on_repeated_centres_equals_walk_order(use_alias, emb, radius, k):
    """Walk-like sequences over a few dozen rows with one hub: a row is the centre hundreds of times, so the bucketed kernel sees
    runs that span several 32-entry chunks, keeps the centre row in registers across a run and draws two occurrences per Philox
    round (K = 5, r = 5) or one (K = 7, r = 3: 14 ... 21 drawing lanes).  The (centre, negative) pairs are those of the walk-order
    kernel -- the owned-negative COUNT is identical on every simulated rank -- and with a tiny learning rate (order of the
    updates matters at second order only) both kernels land on the same tables; rows nobody names stay bit-identical."""
    dev = cuda_device()
    rng = np.random.default_rng(71)
    n_seq, length, offset, world = 96, 2 * radius + 9, 1, 2
    vocab = (4096 if emb == 128 else 8192) * 5 + 100
    ids = rng.choice(vocab - offset, 40, replace=False)
    pick = rng.integers(0, 40, (n_seq, length))
    pick[rng.random((n_seq, length)) < 0.3] = 0                    # the hub: ~30 % of all centres
    tokens = ids[pick].astype(np.int32)
    alias = None
    if use_alias:
        alias = nat.alias_build(rng.integers(1, 50, vocab).astype(np.float64), 0.75, dev)
    w_in = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    w_out = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    lr, seed = 1e-6, 123
    tok = _t(tokens, dev)
    allrows = torch.arange(vocab, device=dev)
    res = {}
    for grouped in (False, True, 'walk-order with positives', 'all pairs'):
        s_in = ShardedTable(vocab, emb, dev, rank=0, world=world, simulate=True)
        s_out = ShardedTable(vocab, emb, dev, rank=0, world=world, simulate=True)
        s_in.scatter(allrows, _t(w_in, dev)); s_out.scatter(allrows, _t(w_out, dev))
        per_rank = []
        # (with the positives the hub's rows collect thousands of same-signed increments: at lr 1e-6 those are an ulp of the fp32 weights
        #  and the red.add roundings of the walk-order path pile up to 2.5e-5; ten times larger steps keep rounding out of the comparison)
        lr = 1e-5 if isinstance(grouped, str) else 1e-6
        if grouped == 'walk-order with positives':
            nat.sgns_update_walks(s_in, s_out, tok, radius, 0, offset, lr, seed, centre_id_base=7_000_000_000)
        for r in range(world):
            stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)
            if grouped == 'all pairs':
                nat.sgns_update_pairs_owned(s_in.as_rank(r), s_out.as_rank(r), tok, radius, k, offset, lr, seed, centre_id_base=7_000_000_000,
                                            alias=alias, stats=stats, positives=True)
            else:
                nat.sgns_update_negatives_owned(s_in.as_rank(r), s_out.as_rank(r), tok, radius, k, offset, lr, seed,
                                                centre_id_base=7_000_000_000, alias=alias, stats=stats, grouped=grouped is True)
            per_rank.append(stats.tolist())
        res[grouped] = (s_in.to_tensor().cpu().numpy().astype(np.float64), s_out.to_tensor().cpu().numpy().astype(np.float64), per_rank)
        s_in.close(); s_out.close()
    # every pair on the owner of its output row == positives in the window kernel + owned negatives in walk order
    a, b = res['walk-order with positives'], res['all pairs']
    assert sum(st[4] for st in b[2]) == n_seq * (length - 2 * radius) * 2 * radius and sum(st[5] for st in b[2]) == sum(st[5] for st in a[2])
    ctx_rows = sgns_oracle.windows_from_walks(tokens.astype(np.int64), radius, offset)[1]
    for r in range(world):
        assert b[2][r][4] == int((((ctx_rows // (4096 if emb == 128 else 8192)) % world) == r).sum())
    move = max(np.abs(a[0] - w_in).max(), np.abs(a[1] - w_out).max())
    assert move > 2e-4                                                               # positives move the few context rows a lot
    assert np.abs(b[0] - a[0]).max() <= 0.02 * move and np.abs(b[1] - a[1]).max() <= 0.02 * move
    lr = 1e-6
    n_cen = length - 2 * radius
    assert sum(st[5] for st in res[True][2]) == n_seq * n_cen * 2 * radius * k
    for r in range(world):
        a, b = res[False][2][r], res[True][2][r]
        assert a[5] == b[5] and abs(a[3] - b[3]) <= 2 and a[5] > 0             # same owned negatives, same false-positive count (ties aside)
        assert abs(a[1] - b[1]) <= 1e-5 * abs(a[1])
    d_in, d_out = res[False][0] - w_in, res[False][1] - w_out
    assert np.abs(d_in).max() > 20 * lr and np.abs(d_out).max() > 0.05 * lr        # the hub row moved by thousands of pair updates
    # fp32 accumulation order differs (one reduction per run against one per occurrence): a few ulps of the row values
    np.testing.assert_allclose(res[True][0], res[False][0], rtol=0, atol=5e-6)
    np.testing.assert_allclose(res[True][1], res[False][1], rtol=0, atol=5e-6)
    assert np.abs((res[True][0] - w_in) - d_in).max() <= 0.02 * np.abs(d_in).max()
    centres = np.unique(tokens[:, radius:length - radius]) + offset
    untouched = np.setdiff1d(np.arange(vocab), centres)
    assert np.array_equal(res[True][0][untouched], w_in[untouched].astype(np.float64))


def test_host_step_on_sharded_tables_matches_device_calls():
    dev = cuda_device()
    from helpers import random_csr
    from shallow_encoders.graph.csr import CSRGraph
    rowptr, col = random_csr(3000, 9000, 5, sort_rows=True)
    csr = CSRGraph.from_arrays(rowptr, col, device=dev)
    vocab, emb, radius, k, L = 3001, 128, 2, 3, 12
    starts = torch.arange(0, 3000, 3, dtype=torch.int32)
    res = []
    for table_kind in ('torch', 'vmm'):
        if table_kind == 'torch':
            w_in = torch.empty((vocab, emb), device=dev); w_out = torch.empty((vocab, emb), device=dev)
        else:
            w_in, w_out = ShardedTable(vocab, emb, dev), ShardedTable(vocab, emb, dev)
        nat.table_fill_uniform(w_in, 0.1, 1); nat.table_fill_uniform(w_out, 0.1, 2)
        scratch = {'starts': torch.empty(starts.numel(), dtype=torch.int32, device=dev),
                   'walks': torch.empty((starts.numel(), L), dtype=torch.int32, device=dev),
                   'stats': torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)}
        stats_host = torch.zeros(nat.STATS_LEN, dtype=torch.float64)
        walks_host = torch.empty((starts.numel(), L), dtype=torch.int32)
        nat.host_walk_sgns_step(csr, starts, L, 0.5, 2.0, True, nat.RULE_REFERENCE, 9, 0, w_in, w_out, radius, k, 1, 1e-4, scratch,
                                stats_host, walks_host=walks_host)
        dense = (w_in if table_kind == 'torch' else w_in.to_tensor()).cpu().numpy()
        res.append((stats_host.clone().numpy(), walks_host.clone().numpy(), dense))
    assert np.array_equal(res[0][1], res[1][1])
    assert res[0][0][4] == starts.numel() * (L - 2 * radius) * 2 * radius
    # Hogwild: rows are read while other warps update them, so two runs agree to the size of one update, not bit for bit
    assert np.array_equal(res[0][0][4:], res[1][0][4:])
    np.testing.assert_allclose(res[0][0][:4], res[1][0][:4], rtol=5e-3)
    np.testing.assert_allclose(res[0][2], res[1][2], rtol=0, atol=1e-4)


def test_nccl_baseline_step_equals_minibatch_sgd_on_one_gpu():
    """RowShardedTables (the all-to-all baseline) with world = 1: windows + device negatives + unique + se_sgns_grad + apply
    == the oracle's mini-batch SGD on the same pairs (negatives restated with the numpy Philox)."""
    dev = cuda_device()
    from shallow_encoders.word2vec.row_exchange import RowShardedTables
    rng = np.random.default_rng(41)
    vocab, emb, radius, k, n_seq, offset, lr, seed = 30000, 64, 2, 3, 40, 1, 0.025, 17
    length = 9
    tokens = rng.integers(0, vocab - offset, (n_seq, length)).astype(np.int32)          # repeats allowed: mini-batch semantics
    inputs, targets = sgns_oracle.windows_from_walks(tokens.astype(np.int64), radius, offset)
    b, n = targets.shape
    w_in = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    w_out = (rng.standard_normal((vocab, emb)) * 0.3).astype(np.float32)
    t = RowShardedTables(vocab, emb, 0, 1, dev)
    t.load_full('in', _t(w_in, dev)); t.load_full('out', _t(w_out, dev))
    stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)
    t.step(_t(tokens, dev), radius, k, offset, lr, seed, draw_id_base=1000, micro_walks=n_seq, stats=stats)
    neg = philox_ref.draws(seed, b * n * k, vocab, base=1000).reshape(b, n, k)
    want_in, want_out, o = sgns_oracle.sgd_step(w_in.astype(np.float64), w_out.astype(np.float64), inputs, targets, neg, lr * b * n)
    np.testing.assert_allclose(t.gather_full('in').cpu().numpy(), want_in, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(t.gather_full('out').cpu().numpy(), want_out, rtol=1e-4, atol=1e-5)
    st = stats.tolist()
    assert st[4] == b * n and abs((st[0] + st[1]) / st[4] - o['loss']) < 1e-4 * o['loss']


@pytest.mark.parametrize('world,vocab,emb', [(2, 5000, 128), (4, 4097, 48), (8, 35, 2), (3, 777, 20)])
def test_replicated_tables_sync_equals_the_sum_of_all_updates_simulated_ranks(world, vocab, emb):
    """ReplicatedTable (csrc/replica.cu) with every rank's working copy on ONE GPU: each simulated rank trains its copy on its
    own walks with the unchanged fused kernel (global negative draw), then every rank runs its share of the fused
    reduce-scatter + all-gather.  Afterwards every copy and every master chunk equals start + sum_g (copy_g - start) --
    synchronous data-parallel SGD with summed updates -- to fp32 rounding, and a second sync without updates changes nothing."""
    dev = cuda_device()
    radius, k, offset = 2, 3, 1
    t_in = ReplicatedTable(vocab, emb, dev, rank=0, world=world, simulate=True)
    t_out = ReplicatedTable(vocab, emb, dev, rank=0, world=world, simulate=True)
    start_in = torch.empty((vocab, emb), device=dev); start_out = torch.empty((vocab, emb), device=dev)
    nat.table_fill_uniform(start_in, 0.4, 21); nat.table_fill_uniform(start_out, 0.4, 22)
    rows = torch.arange(vocab, device=dev)
    for r in range(world):
        nat.table_scatter_rows(t_in.as_rank(r), rows, start_in); nat.table_scatter_rows(t_out.as_rank(r), rows, start_out)
        t_in.adopt(r); t_out.adopt(r)
    rng = np.random.default_rng(5)
    after_in, after_out = [], []
    for r in range(world):
        walks = _t(rng.integers(0, vocab - offset, (64, 9)).astype(np.int32), dev)
        st = nat.sgns_update_walks(t_in.as_rank(r), t_out.as_rank(r), walks, radius, k, offset, 0.05, seed=30 + r, centre_id_base=1000 * r)
        assert st['pairs'] == 64 * 5 * 4
        after_in.append(nat.table_gather_rows(t_in.as_rank(r), rows)); after_out.append(nat.table_gather_rows(t_out.as_rank(r), rows))
    assert float((after_in[0] - start_in).abs().max()) > 1e-4                  # the step moved something
    want_in = start_in.double() + sum((a.double() - start_in.double()) for a in after_in)
    want_out = start_out.double() + sum((a.double() - start_out.double()) for a in after_out)
    for r in range(world):
        t_in.sync_local(r); t_out.sync_local(r)
    for r in range(world):
        got_in = nat.table_gather_rows(t_in.as_rank(r), rows); got_out = nat.table_gather_rows(t_out.as_rank(r), rows)
        assert float((got_in.double() - want_in).abs().max()) < 2e-6 and float((got_out.double() - want_out).abs().max()) < 2e-6
        assert torch.equal(got_in, nat.table_gather_rows(t_in.as_rank(0), rows))          # all copies identical bit for bit
        lo, hi = nat.replica_chunk(vocab * emb, world, r)
        hi = max(lo, min(hi, vocab * emb))
        assert torch.equal(t_in._master_for(r)[:hi - lo], got_in.reshape(-1)[lo:hi])
    before = nat.table_gather_rows(t_in.as_rank(world - 1), rows)
    for r in range(world):
        t_in.sync_local(r)
    assert torch.equal(nat.table_gather_rows(t_in.as_rank(0), rows), before)         # idempotent without new updates
    # mode 2: masters pushed back into every copy (checkpoint restore path)
    nat.table_scatter_rows(t_in.as_rank(1 % world), rows, torch.zeros((vocab, emb), device=dev))
    for r in range(world):
        t_in.sync_local(r, mode=2)
    assert torch.equal(nat.table_gather_rows(t_in.as_rank(1 % world), rows), before)
    v = t_in.view()
    assert v.shape == (vocab, emb) and torch.equal(v, nat.table_gather_rows(t_in.as_rank(0), rows))
    t_in.close(); t_out.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs on one NVLink/NVSwitch node')
def test_two_gpus_train_one_pair_of_tables_over_nvlink():
    world = min(torch.cuda.device_count(), 4)
    world = 2 if world < 4 else 4
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'deepwalk-and-node2vec_b200'), ROOT, os.path.join(ROOT, 'tests')]))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
           '--master-port', str(29700 + os.getpid() % 200), os.path.join(ROOT, 'tests', 'mgpu_worker.py')]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count('MGPU_OK') == world, out.stdout[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs on one NVLink/NVSwitch node')
@pytest.mark.parametrize('mode', ['owner', 'synced'])
def test_train_cli_on_two_gpus_matches_single_gpu_accuracy(tmp_path, mode):
    """`torchrun tools/train.py ... train.engine=fused train.multi_gpu_negatives=owner` on 2 GPUs trains ONE striped model
    (positives on the home GPU, negatives on the GPU that owns their rows); rank 0's checkpoint has the
    reference's state-dict keys and its downstream node-classification accuracy (tools/graph_model_downstream_classification.py
    :94-148 restated in tools/downstream.py) matches what one GPU reaches on the same config (0.98 +- 0.01)."""
    from shallow_encoders.config_parser import load_config
    from shallow_encoders.config_parser.core import instantiate
    from tools.downstream import node_classification
    pkg = os.path.join(ROOT, 'deepwalk-and-node2vec_b200')
    over = ['train.engine=fused', 'train.fused_lr=60.0', 'train.max_epochs=8', 'train.scheduler.step_size=4', 'model.embedding_size=128',
            'datamodule.additional_parameters.method_params.q=0.5', f'train.multi_gpu_negatives={mode}', f'path.output_dir={tmp_path}']
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([pkg, ROOT]))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', str(29900 + os.getpid() % 90), os.path.join(pkg, 'tools', 'train.py'), '--config-name=sge_sg_cora', *over]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    cfg = load_config('sge_sg_cora', over)
    ckpt = torch.load(os.path.join(str(tmp_path), cfg.datamodule.dataset_name, cfg.train.experiment, 'checkpoints', 'last.ckpt'), map_location='cpu')
    sd = ckpt['state_dict']
    assert set(sd) == {'_model._input_embedding.weight', '_model._output_embedding.weight'}
    ds = cfg.datamodule.instantiate_dataset()
    w = sd['_model._input_embedding.weight'].numpy()
    assert w.shape == (len(ds.vocab), 128) and np.isfinite(w).all()
    nc = cfg.downstream['node_classification']
    acc = node_classification(w, ds.vocab.get_itos(), ds.labels, instantiate(nc['split_algorithm']), 10, nc.get('classifier_params'))
    assert acc[0] > 0.96, acc

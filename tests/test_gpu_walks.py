"""GPU parity tests for the walk kernels, through the C ABI.

  * exact mode: bit-identical to the reference (golden vectors) and to the oracle on larger random graphs
  * production mode (Philox + rejection): every transition is an edge, the empirical transition frequencies pass a
    chi-square test against the exact node2vec probabilities of the reference CODE rule, and sharded generation
    reproduces the single-launch walks bit for bit
"""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, WALK_CASES, cuda_device, oracle_graph_from_csr, random_csr
from oracle import walk_oracle
from oracle.c_oracle import c_walks
from shallow_encoders import _native as nat
from shallow_encoders.graph.csr import CSRGraph

pytestmark = pytest.mark.gpu


def _csr_from_golden(z, dev):
    w = z['w'] if bool(z['weighted']) else None
    return CSRGraph.from_arrays(z['rowptr'], z['col'], w, bool(z['w_is_int']), device=dev), w


@pytest.mark.parametrize('tag', WALK_CASES)
def test_exact_walks_match_reference_golden(tag):
    dev = cuda_device()
    z = np.load(os.path.join(GOLDEN, f'walks_{tag}.npz'))
    csr, _ = _csr_from_golden(z, dev)
    got = nat.walk_exact(csr, torch.from_numpy(z['starts']).to(dev), int(z['length']), float(z['p']), float(z['q']),
                         bool(z['node2vec']), nat.RULE_REFERENCE, torch.from_numpy(z['uniforms']).to(dev))
    assert np.array_equal(got.cpu().numpy(), z['walks'])


@pytest.mark.parametrize('n,m,length,p,q,weights', [
    (5000, 40000, 40, 0.5, 2.0, None),
    (5000, 40000, 40, 1.0, 1.0, None),
    (2708, 5429, 10, 1.0, 2.0, None),          # Cora-shaped, sge_sg_cora.yaml p/q
    (3000, 30000, 30, 4.0, 0.25, 'int'),
    (3000, 30000, 30, 0.3, 3.0, 'float'),
])
def test_exact_walks_match_c_oracle_on_random_graphs(n, m, length, p, q, weights):
    dev = cuda_device()
    rowptr, col = random_csr(n, m, seed=n + m)
    rng = np.random.default_rng(1)
    w, w_is_int = None, True
    if weights is not None:
        # symmetric weights: w(u,v) = f(min,max)
        src = np.repeat(np.arange(n), np.diff(rowptr))
        lo, hi = np.minimum(src, col).astype(np.int64), np.maximum(src, col).astype(np.int64)
        h = (lo * 1000003 + hi * 7919) % 1000
        w = (1 + h % 9).astype(np.float64) if weights == 'int' else 0.25 + h / 250.0
        w_is_int = weights == 'int'
    csr = CSRGraph.from_arrays(rowptr, col, w, w_is_int, device=dev)
    starts = rng.integers(0, n, 4096).astype(np.int32)
    uni = rng.random((len(starts), length - 1))
    for node2vec in (True, False):
        want = c_walks(rowptr, col, w, w_is_int, starts, length, p, q, node2vec, 0, uni)
        got = nat.walk_exact(csr, torch.from_numpy(starts).to(dev), length, p, q, node2vec, 0, torch.from_numpy(uni).to(dev))
        assert np.array_equal(got.cpu().numpy(), want)


def test_exact_paper_rule_flag():
    dev = cuda_device()
    rowptr, col = random_csr(500, 3000, seed=5)
    csr = CSRGraph.from_arrays(rowptr, col, device=dev)
    rng = np.random.default_rng(2)
    starts = rng.integers(0, 500, 512).astype(np.int32)
    uni = rng.random((512, 19))
    want = c_walks(rowptr, col, None, True, starts, 20, 0.5, 2.0, True, 1, uni)
    got = nat.walk_exact(csr, torch.from_numpy(starts).to(dev), 20, 0.5, 2.0, True, nat.RULE_PAPER, torch.from_numpy(uni).to(dev))
    assert np.array_equal(got.cpu().numpy(), want)
    ref_rule = c_walks(rowptr, col, None, True, starts, 20, 0.5, 2.0, True, 0, uni)
    assert not np.array_equal(want, ref_rule)


def test_exact_edge_cases():
    dev = cuda_device()
    rowptr, col = random_csr(50, 100, seed=9)
    csr = CSRGraph.from_arrays(rowptr, col, device=dev)
    starts = torch.arange(50, dtype=torch.int32, device=dev)
    # walk_len 1: just the start nodes, no uniforms consumed
    got = nat.walk_exact(csr, starts, 1, 1.0, 1.0, True, 0, torch.empty(0, dtype=torch.float64, device=dev))
    assert got[:, 0].tolist() == list(range(50))
    # empty batch
    got = nat.walk_exact(csr, starts[:0], 5, 1.0, 1.0, True, 0, torch.empty(0, dtype=torch.float64, device=dev))
    assert got.shape == (0, 5)
    # u -> 1 picks the last neighbour (bisect hi = n-1), u = 0 the first
    uni = torch.full((50, 1), np.nextafter(1.0, 0.0), dtype=torch.float64, device=dev)
    got = nat.walk_exact(csr, starts, 2, 1.0, 1.0, False, 0, uni).cpu().numpy()
    assert all(got[i, 1] == col[rowptr[i + 1] - 1] for i in range(50))
    got = nat.walk_exact(csr, starts, 2, 1.0, 1.0, False, 0, torch.zeros((50, 1), dtype=torch.float64, device=dev)).cpu().numpy()
    assert all(got[i, 1] == col[rowptr[i]] for i in range(50))
    with pytest.raises(AssertionError):
        nat.walk_exact(csr, starts, 0, 1.0, 1.0, True, 0, uni)     # random_walk_generator.py:21


# ------------------------------------------------------------------------------------------------------------------
def _edges_set(rowptr, col):
    src = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    return set(zip(src.tolist(), col.tolist()))


def _chi_square_transitions(walks, og, p, q, node2vec, min_expected=5.0):
    """Pooled chi-square over (prev, cur) buckets; returns (statistic, dof)."""
    from collections import defaultdict
    counts = defaultdict(lambda: defaultdict(int))
    w = walks
    for s in range(1, w.shape[1] - 1):
        for t, v, x in zip(w[:, s - 1], w[:, s], w[:, s + 1]):
            counts[(int(t), int(v))][int(x)] += 1
    stat, dof = 0.0, 0
    for (t, v), c in counts.items():
        total = sum(c.values())
        probs = walk_oracle.transition_probabilities(og, t, v, p, q, node2vec)
        exp = probs * total
        obs = np.array([c.get(x, 0) for x in og.adj[v]], dtype=np.float64)
        big = exp >= min_expected
        if big.sum() < 1 or (~big).sum() + big.sum() < 2:
            continue
        e = np.concatenate([exp[big], [exp[~big].sum()]]) if (~big).any() else exp[big]
        o = np.concatenate([obs[big], [obs[~big].sum()]]) if (~big).any() else obs[big]
        keep = e >= min_expected
        if keep.sum() < 2:
            continue
        e, o = e[keep], o[keep]
        # renormalise after dropping tiny cells
        e = e * (o.sum() / e.sum())
        stat += float(((o - e) ** 2 / e).sum())
        dof += len(e) - 1
    return stat, dof


@pytest.mark.parametrize('case', ['karate_weighted', 'triangle_rich', 'star_hub'])
@pytest.mark.parametrize('p,q', [(1.0, 0.5), (0.5, 2.0), (4.0, 0.25)])
def test_philox_walks_chi_square_vs_exact_probabilities(case, p, q):
    from scipy.stats import chi2
    dev = cuda_device()
    if case == 'karate_weighted':
        z = np.load(os.path.join(GOLDEN, 'walks_karate_yaml.npz'))
        rowptr, col, w, w_is_int = z['rowptr'], z['col'], z['w'], True
    elif case == 'star_hub':
        z = np.load(os.path.join(GOLDEN, 'walks_star_hub.npz'))     # degree-400 hub: unstaged lists (> 128 entries)
        rowptr, col, w, w_is_int = z['rowptr'], z['col'], None, True
    else:
        rowptr, col = random_csr(60, 400, seed=11)
        w, w_is_int = None, True
    og = oracle_graph_from_csr(rowptr, col, w, w_is_int)
    csr = CSRGraph.from_arrays(rowptr, col, w, w_is_int, device=dev)
    n = len(rowptr) - 1
    reps = 20000 // n + 1 if case != 'star_hub' else 40
    starts = torch.arange(n, dtype=torch.int32, device=dev).repeat(reps)
    walks = nat.walk(csr, starts, 12, p, q, True, nat.RULE_REFERENCE, seed=1234).cpu().numpy()
    edges = _edges_set(rowptr, col)
    assert all((int(a), int(b)) in edges for a, b in zip(walks[:, :-1].ravel(), walks[:, 1:].ravel()))
    assert np.array_equal(walks[:, 0], starts.cpu().numpy())
    stat, dof = _chi_square_transitions(walks, og, p, q, True)
    assert dof > 20
    # one pooled test at alpha = 1e-6 (a wrong multiplier rule fails by orders of magnitude)
    assert stat < chi2.ppf(1 - 1e-6, dof), (stat, dof)
    if case == 'star_hub':
        # transitions out of the degree-400 hub (unstaged list), classed as return / ring-neighbour of t / other
        obs, exp = np.zeros(3), np.zeros(3)
        for s_ in range(1, walks.shape[1] - 1):
            sel = (walks[:, s_] == 0)
            t_, x_ = walks[sel, s_ - 1], walks[sel, s_ + 1]
            obs += [(x_ == t_).sum(), (np.abs(x_ - t_) == 1).sum(), (np.abs(x_ - t_) > 1).sum()]
            ring = np.where((t_ == 1) | (t_ == 400), 1, 2)
            wsum = 1 / p + ring / q + (399 - ring)
            exp += [(1 / p / wsum).sum(), (ring / q / wsum).sum(), ((399 - ring) / wsum).sum()]
        assert obs.sum() > 20000
        assert float(((obs - exp) ** 2 / exp).sum()) < chi2.ppf(1 - 1e-6, 2), (obs, exp)
    # and the PAPER rule must be rejected by the same statistic when p != q-symmetric
    if case == 'triangle_rich' and q != 1.0:
        w2 = nat.walk(csr, starts, 12, p, q, True, nat.RULE_PAPER, seed=1234).cpu().numpy()
        stat2, dof2 = _chi_square_transitions(w2, og, p, q, True)
        assert stat2 > chi2.ppf(1 - 1e-6, dof2)


def test_philox_first_step_and_deepwalk_are_unbiased():
    from scipy.stats import chi2
    dev = cuda_device()
    z = np.load(os.path.join(GOLDEN, 'walks_karate_yaml.npz'))
    rowptr, col, w = z['rowptr'], z['col'], z['w']
    csr = CSRGraph.from_arrays(rowptr, col, w, True, device=dev)
    og = oracle_graph_from_csr(rowptr, col, w, True)
    starts = torch.arange(34, dtype=torch.int32, device=dev).repeat(3000)
    for node2vec in (True, False):
        walks = nat.walk(csr, starts, 3 if node2vec else 6, 0.25, 4.0, node2vec, 0, seed=7).cpu().numpy()
        stat, dof = 0.0, 0
        steps = [(0, 1)] if node2vec else [(s, s + 1) for s in range(5)]
        for v in range(34):
            obs = np.zeros(len(og.adj[v]))
            for a, b in steps:
                sel = walks[:, a] == v
                nxt = walks[sel, b]
                for i, x in enumerate(og.adj[v]):
                    obs[i] += (nxt == x).sum()
            probs = walk_oracle.transition_probabilities(og, None, v, 1, 1, False)
            exp = probs * obs.sum()
            stat += float(((obs - exp) ** 2 / exp).sum())
            dof += len(exp) - 1
        assert stat < chi2.ppf(1 - 1e-6, dof), (node2vec, stat, dof)


def test_philox_walks_are_shard_invariant_and_seeded():
    dev = cuda_device()
    rowptr, col = random_csr(2000, 16000, seed=21)
    csr = CSRGraph.from_arrays(rowptr, col, device=dev)
    starts = torch.from_numpy(np.random.default_rng(3).integers(0, 2000, 5000).astype(np.int32)).to(dev)
    full = nat.walk(csr, starts, 33, 0.5, 2.0, True, 0, seed=99)
    again = nat.walk(csr, starts, 33, 0.5, 2.0, True, 0, seed=99)
    assert torch.equal(full, again)
    assert not torch.equal(full, nat.walk(csr, starts, 33, 0.5, 2.0, True, 0, seed=100))
    # contiguous shards (walk_id_base) and strided shards (w mod G == g) reproduce the same rows
    for g_count in (2, 4, 8):
        for g in range(g_count):
            part = nat.walk(csr, starts[g::g_count].contiguous(), 33, 0.5, 2.0, True, 0, seed=99, walk_id_base=g, walk_id_stride=g_count)
            assert torch.equal(part, full[g::g_count])
    half = nat.walk(csr, starts[2500:].contiguous(), 33, 0.5, 2.0, True, 0, seed=99, walk_id_base=2500)
    assert torch.equal(half, full[2500:])


def test_philox_walk_lengths_and_dead_ends():
    dev = cuda_device()
    # path 0-1, node 2 isolated (the reference raises on an isolated node; we stay and count)
    rowptr = np.array([0, 1, 2, 2], dtype=np.int64)
    col = np.array([1, 0], dtype=np.int32)
    csr = CSRGraph.from_arrays(rowptr, col, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    starts = torch.tensor([0, 1, 2], dtype=torch.int32, device=dev)
    for length in (1, 2, 31, 32, 33, 64, 65, 80):
        err.zero_()
        w = nat.walk(csr, starts, length, 1.0, 1.0, True, 0, seed=5, err_count=err).cpu().numpy()
        assert w.shape == (3, length)
        assert w[0].tolist() == [i % 2 for i in range(length)]
        assert w[1].tolist() == [(i + 1) % 2 for i in range(length)]
        assert w[2].tolist() == [2] * length
        assert int(err.item()) == (1 if length > 1 else 0)


@pytest.mark.parametrize('weights', [None, 'int'])
def test_warp_and_thread_kernels_generate_identical_walks(weights):
    """One warp per walk (parallel tries, staged lists) and one thread per walk (sequential tries, interpolation search)
    consume the same Philox counters in the same order: bit-identical output, so kernel selection is only scheduling."""
    dev = cuda_device()
    for n, m, length in ((3000, 40000, 80), (500, 900, 13), (401, None, 16)):
        if m is None:
            z = np.load(os.path.join(GOLDEN, 'walks_star_hub.npz'))
            rowptr, col = z['rowptr'], z['col']
        else:
            rowptr, col = random_csr(n, m, seed=n)
        w = None
        if weights:
            src = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
            lo, hi = np.minimum(src, col).astype(np.int64), np.maximum(src, col).astype(np.int64)
            w = (1 + (lo * 1000003 + hi * 7919) % 9).astype(np.float64)
        csr = CSRGraph.from_arrays(rowptr, col, w, True, device=dev)
        starts = torch.from_numpy(np.random.default_rng(1).integers(0, len(rowptr) - 1, 6000).astype(np.int32)).to(dev)
        for p, q, node2vec, rule in ((0.5, 2.0, True, 0), (4.0, 0.25, True, 1), (1.0, 1.0, False, 0), (1.0, 1.0, True, 0)):
            a = nat.walk(csr, starts, length, p, q, node2vec, rule, seed=31, kernel=nat.WALK_WARP)
            b = nat.walk(csr, starts, length, p, q, node2vec, rule, seed=31, kernel=nat.WALK_THREAD)
            assert torch.equal(a, b), (n, p, q, node2vec, rule)
    # asymmetric (directed) CSR: membership is tested in N(candidate)
    rowptr = np.array([0, 2, 4, 6, 7], dtype=np.int64)
    col = np.array([1, 2, 2, 3, 0, 3, 0], dtype=np.int32)
    csr = CSRGraph.from_arrays(rowptr, col, symmetric=False, device=dev)
    starts = torch.arange(4, dtype=torch.int32, device=dev).repeat(500)
    a = nat.walk(csr, starts, 20, 0.5, 2.0, True, 0, seed=3, kernel=nat.WALK_WARP)
    b = nat.walk(csr, starts, 20, 0.5, 2.0, True, 0, seed=3, kernel=nat.WALK_THREAD)
    assert torch.equal(a, b)


def test_thread_kernel_chi_square_on_hub_graph():
    from scipy.stats import chi2
    dev = cuda_device()
    z = np.load(os.path.join(GOLDEN, 'walks_star_hub.npz'))
    rowptr, col = z['rowptr'], z['col']
    og = oracle_graph_from_csr(rowptr, col)
    csr = CSRGraph.from_arrays(rowptr, col, device=dev)
    starts = torch.arange(401, dtype=torch.int32, device=dev).repeat(40)
    walks = nat.walk(csr, starts, 12, 0.5, 2.0, True, 0, seed=77, kernel=nat.WALK_THREAD).cpu().numpy()
    stat, dof = _chi_square_transitions(walks, og, 0.5, 2.0, True)
    assert dof > 20 and stat < chi2.ppf(1 - 1e-6, dof), (stat, dof)


@pytest.mark.parametrize('p,q,node2vec,rule', [(0.5, 2.0, True, 0), (1.0, 0.5, True, 0), (4.0, 0.25, True, 1), (0.25, 4.0, True, 1), (1.0, 1.0, False, 0)])
def test_production_walks_equal_the_python_restatement_bit_for_bit(p, q, node2vec, rule):
    """Both production kernels == tests/walk_model.py (Philox keys, return-edge split, fp32 acceptance arithmetic) on a random graph
    and on the degree-400 hub graph: the sampler is pinned bit for bit, on top of the chi-square tests against the reference rule."""
    import walk_model
    dev = cuda_device()
    z = np.load(os.path.join(GOLDEN, 'walks_star_hub.npz'))
    for rowptr, col in (random_csr(300, 1500, seed=17, sort_rows=True), (z['rowptr'], np.concatenate([np.sort(z['col'][a:b]) for a, b in zip(z['rowptr'][:-1], z['rowptr'][1:])]).astype(np.int32))):
        csr = CSRGraph.from_arrays(rowptr, col, device=dev)
        n = len(rowptr) - 1
        starts = np.random.default_rng(4).integers(0, n, 300).astype(np.int32)
        want = walk_model.walks(rowptr, csr.col_sorted.cpu().numpy(), starts, 21, p, q, node2vec, rule == 0, seed=1234, walk_id_base=50, walk_id_stride=3)
        for kern in (nat.WALK_WARP, nat.WALK_THREAD):
            got = nat.walk(csr, torch.from_numpy(starts).to(dev), 21, p, q, node2vec, rule, seed=1234, walk_id_base=50, walk_id_stride=3, kernel=kern)
            assert np.array_equal(got.cpu().numpy(), want), (kern, np.argwhere(got.cpu().numpy() != want)[:5])


@pytest.mark.parametrize('p,q', [(0.5, 2.0), (1.0, 0.5), (2.0, 1.0)])
def test_weighted_production_walks_equal_the_python_restatement_bit_for_bit(p, q):
    """Weighted graphs (karate club with its integer weights, a random graph with symmetric weights 1..9): both production kernels ==
    tests/walk_model.py, including the fp64 weighted neighbour pick and the remembered weight of the return edge."""
    import walk_model
    dev = cuda_device()
    z = np.load(os.path.join(GOLDEN, 'walks_karate_yaml.npz'))
    rowptr2, col2 = random_csr(200, 900, seed=23, sort_rows=True)
    src = np.repeat(np.arange(200), np.diff(rowptr2))
    lo, hi = np.minimum(src, col2).astype(np.int64), np.maximum(src, col2).astype(np.int64)
    w2 = (1 + (lo * 1000003 + hi * 7919) % 9).astype(np.float64)
    for rowptr, col, w in ((z['rowptr'], z['col'], z['w']), (rowptr2, col2, w2)):
        csr = CSRGraph.from_arrays(rowptr, col, w, True, device=dev)
        n = len(rowptr) - 1
        starts = np.random.default_rng(6).integers(0, n, 200).astype(np.int32)
        want = walk_model.walks(csr.rowptr.cpu().numpy(), csr.col_sorted.cpu().numpy(), starts, 17, p, q, True, True, seed=77, walk_id_base=9,
                                wcdf=csr.wcdf.cpu().numpy())
        for kern in (nat.WALK_WARP, nat.WALK_THREAD):
            got = nat.walk(csr, torch.from_numpy(starts).to(dev), 17, p, q, True, 0, seed=77, walk_id_base=9, kernel=kern)
            assert np.array_equal(got.cpu().numpy(), want), (kern, np.argwhere(got.cpu().numpy() != want)[:5])

"""Python restatement of the PRODUCTION walk sampler (csrc/walk.cu: walk_kernel / walk_thread_kernel, undirected graphs,
unweighted or with fp32 per-row weight prefix sums `wcdf`), so that tests can demand bit-identical walks -- not only the right distribution -- from the CUDA kernels.

Per transition s of walk `walk_id` (try a = 0, 1, ...): r = Philox4x32-10(seed; walk_id, s, STREAM_WALK | a)
  first step / DeepWalk / p = q = 1:  x = N(v)[mulhi32(r.x, deg)]                                  (one try)
  otherwise, with m = max(1, 1/q), zt = 1/p, Z = fma(deg, m, zt)   (fp32):
      if  rz(r.w) * 2^-32 * Z < zt:          x = t                      (return edge split off, accepted outright)
      else cand = N(v)[mulhi32(r.x, deg)];   cand == t -> rejected
           mult = 1/q if (cand in N(t)) == (rule is REFERENCE) else 1;   accepted iff u01(r.z) * m < mult
Weighted graphs: the neighbour index is the first i with wcdf[i] > u64(r.x, r.y) * 2^-64 * wcdf[deg-1] (fp64, clamped to deg-1),
deg is replaced by W_v = wcdf[deg-1], and zt = w_prev / p where w_prev = wcdf[k] - wcdf[k-1] of the edge the walk arrived over.
All float arithmetic is fp32 exactly as the kernel performs it (the compiler contracts zt + W * m into one FMA).
"""
import numpy as np

import philox_ref

F = np.float32


def _rz(u32: int) -> np.float32:
    """__uint2float_rz: convert with truncation toward zero."""
    f = F(u32)
    if float(f) > u32:
        f = np.nextafter(f, F(0))
    return f


def _fma(a: np.float32, b: np.float32, c: np.float32) -> np.float32:
    return F(np.float64(a) * np.float64(b) + np.float64(c))      # exact product of two fp32 values, one rounding


def _pick(rx, ry, deg, wrow):
    if wrow is None:
        return (rx * deg) >> 32
    target = np.float64(np.uint64((rx << 32) | ry)) * np.float64(2.0 ** -64) * np.float64(wrow[deg - 1])
    lo, hi = 0, deg - 1
    while lo < hi:
        mid = (lo + hi) >> 1
        if target < np.float64(wrow[mid]):
            hi = mid
        else:
            lo = mid + 1
    return lo


def walks(rowptr, col_sorted, starts, walk_len, p, q, node2vec=True, rule_reference=True, seed=0, walk_id_base=0, walk_id_stride=1,
          wcdf=None):
    rowptr, col = np.asarray(rowptr, dtype=np.int64), np.asarray(col_sorted, dtype=np.int64)
    wcdf = None if wcdf is None else np.asarray(wcdf, dtype=np.float32)
    inv_p, inv_q = F(1.0 / p), F(1.0 / q)
    m_o = max(F(1.0), inv_q)
    any_bias = node2vec and not (inv_p == F(1.0) and inv_q == F(1.0))
    out = np.zeros((len(starts), walk_len), dtype=np.int32)
    nbr = [set(col[rowptr[i]:rowptr[i + 1]].tolist()) for i in range(len(rowptr) - 1)]
    for w, start in enumerate(starts):
        walk_id = walk_id_base + w * walk_id_stride
        v, t = int(start), -1
        wprev = F(1.0)
        out[w, 0] = v
        for s in range(1, walk_len):
            base, deg = int(rowptr[v]), int(rowptr[v + 1] - rowptr[v])
            if deg <= 0:
                out[w, s:] = v                    # the reference raises on an isolated node; the kernels stay
                break
            wrow = None if wcdf is None else wcdf[base:base + deg]
            edge_w = (lambda k_: F(wrow[k_] - (wrow[k_ - 1] if k_ > 0 else F(0.0)))) if wrow is not None else None
            attempt = 0
            while True:
                rx, ry, rz_, rw = (int(a[0]) for a in philox_ref.philox(seed, [walk_id], s, philox_ref.STREAM_WALK | attempt))
                if not (any_bias and t >= 0):
                    k = _pick(rx, ry, deg, wrow)
                    x = int(col[base + k])
                    if wrow is not None and any_bias:
                        wprev = edge_w(k)
                    break
                zt = inv_p if wrow is None else F(wprev * inv_p)
                ztot = _fma(F(deg) if wrow is None else wrow[deg - 1], m_o, zt)
                if F(F(_rz(rw) * F(2.3283064365386963e-10)) * ztot) < zt:
                    x = t                                     # the return edge keeps its weight
                    break
                k = _pick(rx, ry, deg, wrow)
                cand = int(col[base + k])
                if cand != t:
                    member = cand in nbr[t]
                    mult = inv_q if (member == rule_reference) else F(1.0)
                    if F(philox_ref.u01(np.uint32(rz_)) * m_o) < mult:
                        x = cand
                        if wrow is not None:
                            wprev = edge_w(k)
                        break
                attempt += 1
            out[w, s] = x
            t, v = v, x
    return out

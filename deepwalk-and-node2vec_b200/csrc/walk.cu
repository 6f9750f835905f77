// Random-walk kernels for sm_100a.
//
//   walk_exact_kernel : fp64 inverse-CDF walk consuming supplied uniforms; bit-identical to the reference's
//                       DeepWalk.walk / Node2Vec.walk (shallow_encoders/graph/random_walk_generator.py:61-72, 94-119)
//                       including CPython 3.12's sum()/accumulate()/bisect_right arithmetic.  Parity mode.
//   walk_kernel       : production mode.  One warp per walk, Philox4x32-10 keyed by (seed; walk id, step, try),
//                       32 rejection tries evaluated in parallel per round (first accepted lane wins, which is the
//                       same distribution as sequential tries), neighbour lists up to STAGE_CAP entries staged in
//                       shared memory so the membership test "x in N(t)" of the NEXT step is a shared-memory
//                       binary search instead of a chain of dependent global loads.
//
// Rejection scheme of the production kernels (undirected graphs).  Target: pi(x | t, v) ~ w_vx * mult(x) with
// mult(t) = 1/p and mult(x != t) in {1/q, 1} (random_walk_generator.py:101-108).  A naive envelope max(1/p, 1, 1/q)
// wastes tries whenever 1/p is the maximum (p < 1: every ordinary candidate is then accepted with probability <= p).
// Instead the RETURN EDGE IS SPLIT OFF: with m = max(1, 1/q) and Z = w_vt/p + W_v * m (W_v = total edge weight of v),
//   with probability (w_vt/p) / Z   take x = t, accepted outright (no memory access at all);
//   otherwise draw x ~ w_vx / W_v; x == t is rejected; x != t is accepted with probability mult(x) / m.
// P(accept x != t) = (W_v m / Z)(w_vx / W_v)(mult(x) / m) = w_vx mult(x) / Z and P(accept t) = (w_vt/p) / Z: exactly the
// target, and the acceptance rate is sum_x w_vx mult(x) / Z, i.e. ~1 try per step for q >= 1 whatever p is.  t is a
// neighbour of v because the walk arrived over that edge; its weight w_vt = w_tv is remembered from the previous step.
// Directed inputs (symmetric == 0) keep the plain envelope.
#include "common.cuh"

namespace se {
namespace {

constexpr int EXACT_WPB = 4;       // warps per block, exact mode
constexpr int EXACT_MAX_WARPS = 4096;

__device__ __forceinline__ bool member_sorted(const int32_t *__restrict__ col_sorted, int64_t lo, int64_t hi,
                                              int32_t key) {
    const int64_t end = hi;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        int32_t c = __ldg(col_sorted + mid);
        if (c < key) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(col_sorted + lo) == key;
}

__global__ void __launch_bounds__(EXACT_WPB * 32)
walk_exact_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                  const int32_t *__restrict__ col_sorted, const double *__restrict__ w, int w_is_int,
                  const int32_t *__restrict__ starts, int64_t n_walks, int walk_len, double inv_p, double inv_q,
                  int node2vec, int rule, const double *__restrict__ uniforms, double *wbuf, unsigned char *fbuf,
                  int64_t max_degree, int64_t fstride, int n_warps, int32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_warps) return;
    double *wb = wbuf + (int64_t)warp * max_degree;
    unsigned char *fb = fbuf + (int64_t)warp * fstride;

    for (int64_t wk = warp; wk < n_walks; wk += n_warps) {
        int32_t prev = -1, node = starts[wk];
        int32_t *o = out + wk * (int64_t)walk_len;
        if (lane == 0 && walk_len > 0) o[0] = node;
        for (int s = 1; s < walk_len; ++s) {
            const int64_t base = rowptr[node];
            const int64_t deg = rowptr[node + 1] - base;
            if (deg <= 0) {  // the reference raises (random.choices on an empty population); stay put
                if (lane == 0) for (int r = s; r < walk_len; ++r) o[r] = node;
                break;
            }
            // unnormalised weights, python object types tracked (int vs float) -- random_walk_generator.py:100-108
            for (int64_t i = lane; i < deg; i += 32) {
                const int32_t x = col[base + i];
                double wi = w ? w[base + i] : 1.0;
                unsigned char fl = (unsigned char)(w != nullptr && !w_is_int);
                if (node2vec && prev >= 0) {
                    if (x == prev) {
                        wi = __dmul_rn(wi, inv_p); fl = 1;
                    } else {
                        const bool m = member_sorted(col_sorted, rowptr[x], rowptr[x + 1], prev);
                        if ((rule == SE_RULE_REFERENCE) ? m : !m) { wi = __dmul_rn(wi, inv_q); fl = 1; }
                    }
                }
                wb[i] = wi; fb[i] = fl;
            }
            __syncwarp();
            int32_t child = 0;
            if (lane == 0) {
                // CPython 3.12 builtin sum(): leading ints exact, first float added plainly, later floats
                // Neumaier-compensated, later ints added plainly, compensation folded in at the end.
                int64_t i = 0;
                double acc = 0.0;
                while (i < deg && !fb[i]) { acc = __dadd_rn(acc, wb[i]); ++i; }
                double total_w = acc;
                if (i < deg) {
                    double f = __dadd_rn(acc, wb[i]); ++i;
                    double c = 0.0;
                    for (; i < deg; ++i) {
                        const double x = wb[i];
                        if (fb[i]) {
                            const double t = __dadd_rn(f, x);
                            if (fabs(f) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dadd_rn(f, -t), x));
                            else c = __dadd_rn(c, __dadd_rn(__dadd_rn(x, -t), f));
                            f = t;
                        } else {
                            f = __dadd_rn(f, x);
                        }
                    }
                    if (c != 0.0 && isfinite(c)) f = __dadd_rn(f, c);
                    total_w = f;
                }
                // cum = accumulate(w_i / s)
                double cum = 0.0;
                for (int64_t j = 0; j < deg; ++j) {
                    const double nw = __ddiv_rn(wb[j], total_w);
                    cum = (j == 0) ? nw : __dadd_rn(cum, nw);
                    wb[j] = cum;
                }
                // bisect_right(cum, u * (cum[-1] + 0.0), 0, deg - 1)
                const double target = __dmul_rn(uniforms[wk * (int64_t)(walk_len - 1) + (s - 1)], __dadd_rn(cum, 0.0));
                int64_t lo = 0, hi = deg - 1;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (target < wb[mid]) hi = mid; else lo = mid + 1;
                }
                child = col[base + lo];
                o[s] = child;
            }
            child = __shfl_sync(FULL, child, 0);
            __syncwarp();
            prev = node;
            node = child;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
constexpr int WALK_WPB = 8;      // warps per block
constexpr int STAGE_CAP = 128;   // neighbour-list entries staged per warp per buffer (2 buffers): 8 KB per block

template <bool WEIGHTED>
__device__ __forceinline__ int64_t pick_index(const float *__restrict__ wcdf, int64_t base, int64_t deg, uint32_t r0,
                                              uint32_t r1) {
    if (!WEIGHTED) {
        return (deg <= 0xffffffffll) ? (int64_t)mulhi32(r0, (uint32_t)deg)
                                     : (int64_t)__umul64hi(((uint64_t)r0 << 32) | r1, (uint64_t)deg);
    }
    const double total = (double)__ldg(wcdf + base + deg - 1);
    const double target = (double)(((uint64_t)r0 << 32) | r1) * (1.0 / 18446744073709551616.0) * total;
    int64_t lo = 0, hi = deg - 1;   // first i with wcdf[i] > target, clamped to deg-1
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (target < (double)__ldg(wcdf + base + mid)) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ bool member_smem(const int32_t *list, int n, int32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (list[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo < n && list[lo] == key;
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(WALK_WPB * 32)
walk_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const float *__restrict__ wcdf,
            int symmetric, const int32_t *__restrict__ starts, int64_t n_walks, int walk_len, float inv_p, float inv_q,
            int node2vec, int rule, uint64_t seed, int64_t walk_id_base, int64_t walk_id_stride,
            int32_t *__restrict__ out, int32_t *__restrict__ err_count) {
    __shared__ int32_t stage[WALK_WPB][2][STAGE_CAP];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * WALK_WPB;
    const float wmax = fmaxf(1.0f, fmaxf(inv_p, inv_q));      // plain envelope (directed inputs)
    const float m_o = fmaxf(1.0f, inv_q);                     // envelope of the candidates other than t (return edge split off)
    const bool any_bias = node2vec && !(inv_p == 1.0f && inv_q == 1.0f);

    for (int64_t wk = (int64_t)blockIdx.x * WALK_WPB + wib; wk < n_walks; wk += n_warps) {
        const uint64_t walk_id = (uint64_t)(walk_id_base + wk * walk_id_stride);
        int32_t *o = out + wk * (int64_t)walk_len;
        int32_t v = __ldg(starts + wk), t = -1;
        int64_t tbase = 0, tdeg = 0;
        float wprev = 1.0f;                                   // weight of the edge the walk arrived over
        bool prev_staged = false;
        int cb = 0;
        int32_t keep = v;  // lane (s & 31) keeps the node of step s; flushed as one coalesced store per 32 steps
        bool dead = false;
        for (int s = 1; s < walk_len; ++s) {
            const int64_t base = __ldg(rowptr + v);
            const int64_t deg = __ldg(rowptr + v + 1) - base;
            int32_t x = v;
            bool cur_staged = false;
            if (deg <= 0) {
                dead = true;   // the reference raises here; we stay on the node and count the walk
            } else {
                cur_staged = deg <= STAGE_CAP;
                int32_t *cur = stage[wib][cb];
                if (cur_staged) {
                    for (int i = lane; i < (int)deg; i += 32) cur[i] = __ldg(col + base + i);
                    __syncwarp();
                }
                if (!(any_bias && t >= 0)) {
                    // unbiased step (DeepWalk, or the first step of a node2vec walk: random_walk_generator.py:97)
                    const uint4 r = philox(seed, walk_id, (uint32_t)s, STREAM_WALK);
                    const int64_t k = pick_index<WEIGHTED>(wcdf, base, deg, r.x, r.y);
                    x = cur_staged ? cur[k] : __ldg(col + base + k);
                    if (WEIGHTED && any_bias) wprev = __ldg(wcdf + base + k) - (k > 0 ? __ldg(wcdf + base + k - 1) : 0.f);
                } else {
                    const int32_t *prv = stage[wib][cb ^ 1];
                    const float zt = symmetric ? (WEIGHTED ? wprev * inv_p : inv_p) : 0.f;           // return-edge mass
                    const float wtot = WEIGHTED ? __ldg(wcdf + base + deg - 1) : (float)deg;
                    const float env = symmetric ? m_o : wmax;
                    const float ztot = zt + wtot * env;
                    for (uint32_t round = 0;; ++round) {
                        const uint32_t attempt = round * 32u + (uint32_t)lane;
                        const uint4 r = philox(seed, walk_id, (uint32_t)s, STREAM_WALK | attempt);
                        int32_t cand = t;
                        int64_t k = -1;
                        bool acc;
                        if (symmetric && __uint2float_rz(r.w) * 2.3283064365386963e-10f * ztot < zt) {
                            acc = true;                                   // return to t, no memory touched
                        } else {
                            k = pick_index<WEIGHTED>(wcdf, base, deg, r.x, r.y);
                            cand = cur_staged ? cur[k] : __ldg(col + base + k);
                            float mult;
                            if (cand == t) {
                                mult = symmetric ? 0.f : inv_p;          // symmetric: t is only reachable through the split
                            } else {
                                bool m;
                                if (symmetric) {
                                    m = prev_staged ? member_smem(prv, (int)tdeg, cand)
                                                    : member_sorted(col, tbase, tbase + tdeg, cand);
                                } else {
                                    const int64_t xb = __ldg(rowptr + cand);
                                    m = member_sorted(col, xb, __ldg(rowptr + cand + 1), t);
                                }
                                mult = ((rule == SE_RULE_REFERENCE) ? m : !m) ? inv_q : 1.0f;
                            }
                            acc = u01(r.z) * env < mult;
                        }
                        acc = acc || attempt >= (1u << 25) - 1;
                        const unsigned ballot = __ballot_sync(FULL, acc);
                        if (ballot) {
                            const int win = __ffs(ballot) - 1;
                            x = __shfl_sync(FULL, cand, win);
                            if (WEIGHTED) {
                                float wsel = wprev;                       // the return edge keeps its weight
                                if (lane == win && k >= 0) wsel = __ldg(wcdf + base + k) - (k > 0 ? __ldg(wcdf + base + k - 1) : 0.f);
                                wprev = __shfl_sync(FULL, wsel, win);
                            }
                            break;
                        }
                    }
                }
            }
            if ((s & 31) == 0) {  // flush steps s-32 .. s-1
                o[s - 32 + lane] = keep;
            }
            if (lane == (s & 31)) keep = x;
            if (deg > 0) {
                t = v; tbase = base; tdeg = deg; prev_staged = cur_staged; cb ^= 1; v = x;
            }
            __syncwarp();
        }
        {   // tail flush: steps (walk_len-1) & ~31 .. walk_len-1
            const int first = (walk_len - 1) & ~31;
            if (first + lane < walk_len) o[first + lane] = keep;
        }
        if (dead && lane == 0 && err_count) atomicAdd(err_count, 1);
    }
}


// ------------------------------------------------------------------------------------------------------------
// walk_thread_kernel: one THREAD per walk, for large batches.  Pointer chasing is latency-bound, so the kernel buys
// parallelism instead: 32 independent dependency chains per warp (2048 walks per SM).  Rejection tries are sequential
// per walk and consume the SAME Philox counters in the same order as walk_kernel's lane-parallel rounds
// (try = round * 32 + lane, first accepted try wins), so both kernels produce bit-identical walks and the choice
// between them is a pure scheduling decision.  The loop is flattened into "one try per iteration" so a lane that
// accepts early moves on to its next step instead of idling while its neighbours retry.  Membership of the candidate in
// N(t) uses interpolation search (node ids in an adjacency row are close to uniformly spread) between VIRTUAL anchors
// (-1 before the row, n_nodes after it: no loads of the row's ends), falling back to binary search, which cuts the
// dependent loads on hub rows (degree 10^3..10^5) from 10-17 to ~4.  Outputs are written as
// 16-byte vectors (4 steps) when the walk rows are 16-byte aligned.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool member_interp(const int32_t *__restrict__ col, int64_t lo, int64_t n, int32_t key, int32_t n_nodes) {
    if (n <= 0) return false;
    // virtual anchors: id -1 just before the row and id n_nodes just after it, so the first probe needs no load of the row's ends
    int64_t l = lo - 1, r = lo + n;          // col[l] = vl < key < vr = col[r]  (virtually at the two ends)
    int32_t vl = -1, vr = n_nodes;
#pragma unroll 1
    for (int it = 0; it < 4 && r - l > 8; ++it) {
        const float frac = (float)((int64_t)key - vl) / (float)((int64_t)vr - vl);
        int64_t m = l + (int64_t)(frac * (float)(r - l));
        m = max(l + 1, min(r - 1, m));
        const int32_t vm = __ldg(col + m);
        if (vm == key) return true;
        if (vm < key) { l = m; vl = vm; } else { r = m; vr = vm; }
    }
    ++l;                                     // remaining candidates: (l, r) exclusive -> [l, r)
    while (l < r) {
        const int64_t m = l + ((r - l) >> 1);
        const int32_t vm = __ldg(col + m);
        if (vm == key) return true;
        if (vm < key) l = m + 1; else r = m;
    }
    return false;
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(256)
walk_thread_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const float *__restrict__ wcdf,
                   int symmetric, int32_t n_nodes, const int32_t *__restrict__ starts, int64_t n_walks, int walk_len, float inv_p,
                   float inv_q, int node2vec, int rule, uint64_t seed, int64_t walk_id_base, int64_t walk_id_stride,
                   int32_t *__restrict__ out, int32_t *__restrict__ err_count) {
    const float wmax = fmaxf(1.0f, fmaxf(inv_p, inv_q));
    const float m_o = fmaxf(1.0f, inv_q);
    const bool any_bias = node2vec && !(inv_p == 1.0f && inv_q == 1.0f);
    const bool vec_out = ((walk_len & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);

    for (int64_t wk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wk < n_walks; wk += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t walk_id = (uint64_t)(walk_id_base + wk * walk_id_stride);
        int32_t *o = out + wk * (int64_t)walk_len;
        int32_t v = __ldg(starts + wk), t = -1;
        int64_t base = 0, deg = 0, tbase = 0, tdeg = 0;
        float wprev = 1.0f, wtot = 0.f;
        int32_t pend[4];
        pend[0] = v;
        if (!vec_out && walk_len > 0) o[0] = v;
        bool dead = false, need_row = true;
        int s = 1;
        uint32_t attempt = 0;
        while (s < walk_len) {
            if (need_row) {
                base = __ldg(rowptr + v);
                deg = __ldg(rowptr + v + 1) - base;
                if (deg > 0 && any_bias) wtot = WEIGHTED ? __ldg(wcdf + base + deg - 1) : (float)deg;
                need_row = false;
                attempt = 0;
            }
            int32_t x = v;
            bool commit = true;
            if (deg <= 0) {
                dead = true;                    // the reference raises on an isolated node; stay and count
            } else {
                const uint4 r = philox(seed, walk_id, (uint32_t)s, STREAM_WALK | attempt);
                const bool biased = any_bias && t >= 0;
                const float zt = (biased && symmetric) ? (WEIGHTED ? wprev * inv_p : inv_p) : 0.f;
                const float env = symmetric ? m_o : wmax;
                if (biased && symmetric && __uint2float_rz(r.w) * 2.3283064365386963e-10f * (zt + wtot * env) < zt) {
                    x = t;                                  // return edge, accepted outright (see the scheme above)
                    ++attempt;
                } else {
                    const int64_t k = pick_index<WEIGHTED>(wcdf, base, deg, r.x, r.y);
                    const int32_t cand = __ldg(col + base + k);
                    x = cand;
                    if (biased) {
                        float mult;
                        if (cand == t) {
                            mult = symmetric ? 0.f : inv_p;
                        } else {
                            bool m;
                            if (symmetric) {
                                m = member_interp(col, tbase, tdeg, cand, n_nodes);
                            } else {
                                const int64_t xb = __ldg(rowptr + cand);
                                m = member_sorted(col, xb, __ldg(rowptr + cand + 1), t);
                            }
                            mult = ((rule == SE_RULE_REFERENCE) ? m : !m) ? inv_q : 1.0f;
                        }
                        commit = (u01(r.z) * env < mult) || attempt >= (1u << 25) - 1;
                        ++attempt;
                    }
                    if (WEIGHTED && any_bias && commit) wprev = __ldg(wcdf + base + k) - (k > 0 ? __ldg(wcdf + base + k - 1) : 0.f);
                }
            }
            if (commit) {
                if (vec_out) {
                    pend[s & 3] = x;
                    if ((s & 3) == 3) *reinterpret_cast<int4 *>(o + s - 3) = make_int4(pend[0], pend[1], pend[2], pend[3]);
                } else {
                    o[s] = x;
                }
                if (deg > 0) { t = v; tbase = base; tdeg = deg; v = x; need_row = true; }
                ++s;
            }
        }
        if (dead && err_count) atomicAdd(err_count, 1);
    }
}

}  // namespace
}  // namespace se

extern "C" int64_t se_walk_exact_scratch_bytes(int64_t max_degree, int64_t n_walks) {
    if (max_degree < 1) max_degree = 1;
    int64_t warps = n_walks < se::EXACT_MAX_WARPS ? n_walks : se::EXACT_MAX_WARPS;
    if (warps < 1) warps = 1;
    const int64_t fstride = (max_degree + 7) & ~7ll;
    return warps * (max_degree * 8 + fstride);
}

extern "C" int se_walk_exact(const int64_t *rowptr, const int32_t *col, const int32_t *col_sorted, const double *w,
                             int w_is_int, int64_t n_nodes, int64_t max_degree, const int32_t *starts,
                             int64_t n_walks, int walk_len, double p, double q, int node2vec, int rule,
                             const double *uniforms, void *scratch, int64_t scratch_bytes, int32_t *out,
                             void *stream) {
    SE_REQUIRE(n_nodes > 0 && n_walks >= 0 && walk_len >= 1, "se_walk_exact: bad sizes (walk length must be >= 1)");
    SE_REQUIRE(p > 0 && q > 0, "se_walk_exact: p and q must be positive");
    SE_REQUIRE(rule == SE_RULE_REFERENCE || rule == SE_RULE_PAPER, "se_walk_exact: unknown rule %d", rule);
    if (n_walks == 0) return SE_OK;
    SE_REQUIRE(rowptr && col && col_sorted && starts && out, "se_walk_exact: null graph/starts/out pointer");
    SE_REQUIRE(walk_len == 1 || uniforms, "se_walk_exact: uniforms required");
    if (max_degree < 1) max_degree = 1;
    const int64_t fstride = (max_degree + 7) & ~7ll;
    const int64_t per_warp = max_degree * 8 + fstride;
    SE_REQUIRE(scratch && scratch_bytes >= per_warp, "se_walk_exact: scratch too small (%lld < %lld)",
               (long long)scratch_bytes, (long long)per_warp);
    int64_t n_warps = scratch_bytes / per_warp;
    if (n_warps > se::EXACT_MAX_WARPS) n_warps = se::EXACT_MAX_WARPS;
    if (n_warps > n_walks) n_warps = n_walks;
    double *wbuf = (double *)scratch;
    unsigned char *fbuf = (unsigned char *)(wbuf + n_warps * max_degree);
    const int blocks = (int)((n_warps + se::EXACT_WPB - 1) / se::EXACT_WPB);
    se::walk_exact_kernel<<<blocks, se::EXACT_WPB * 32, 0, (cudaStream_t)stream>>>(
        rowptr, col, col_sorted, w, w_is_int, starts, n_walks, walk_len, 1.0 / p, 1.0 / q, node2vec, rule, uniforms,
        wbuf, fbuf, max_degree, fstride, (int)n_warps, out);
    return se::check_cuda(cudaGetLastError(), "walk_exact_kernel launch");
}

extern "C" int se_walk(const int64_t *rowptr, const int32_t *col, const float *wcdf, int64_t n_nodes, int symmetric,
                       const int32_t *starts, int64_t n_walks, int walk_len, double p, double q, int node2vec,
                       int rule, uint64_t seed, int64_t walk_id_base, int64_t walk_id_stride, int32_t *out,
                       int32_t *err_count, int flags, void *stream) {
    SE_REQUIRE(n_nodes > 0 && n_walks >= 0 && walk_len >= 1, "se_walk: bad sizes (walk length must be >= 1)");
    SE_REQUIRE(p > 0 && q > 0, "se_walk: p and q must be positive");
    SE_REQUIRE(rule == SE_RULE_REFERENCE || rule == SE_RULE_PAPER, "se_walk: unknown rule %d", rule);
    if (n_walks == 0) return SE_OK;
    SE_REQUIRE(rowptr && col && starts && out, "se_walk: null graph/starts/out pointer");
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    SE_REQUIRE(flags == SE_WALK_AUTO || flags == SE_WALK_WARP || flags == SE_WALK_THREAD, "se_walk: unknown flags %d", flags);
    const float inv_p = (float)(1.0 / p), inv_q = (float)(1.0 / q);
    cudaStream_t st = (cudaStream_t)stream;
    // both kernels generate identical walks; one thread per walk needs ~a full machine of walks to pay off
    const bool per_thread = flags == SE_WALK_THREAD || (flags == SE_WALK_AUTO && n_walks >= (int64_t)sms * 1024);
    if (per_thread) {
        int64_t blocks = (n_walks + 255) / 256;
        const int64_t cap = (int64_t)sms * 8;
        if (blocks > cap) blocks = cap;
        if (wcdf)
            se::walk_thread_kernel<true><<<(int)blocks, 256, 0, st>>>(rowptr, col, wcdf, symmetric, (int32_t)n_nodes, starts, n_walks, walk_len, inv_p,
                                                                      inv_q, node2vec, rule, seed, walk_id_base, walk_id_stride, out, err_count);
        else
            se::walk_thread_kernel<false><<<(int)blocks, 256, 0, st>>>(rowptr, col, wcdf, symmetric, (int32_t)n_nodes, starts, n_walks, walk_len, inv_p,
                                                                       inv_q, node2vec, rule, seed, walk_id_base, walk_id_stride, out, err_count);
        return se::check_cuda(cudaGetLastError(), "walk_thread_kernel launch");
    }
    // persistent grid: 8 resident blocks of 8 warps per SM (64 warps/SM), capped by the work available
    int64_t blocks = (n_walks + se::WALK_WPB - 1) / se::WALK_WPB;
    const int64_t cap = (int64_t)sms * 8;
    if (blocks > cap) blocks = cap;
    if (wcdf)
        se::walk_kernel<true><<<(int)blocks, se::WALK_WPB * 32, 0, st>>>(
            rowptr, col, wcdf, symmetric, starts, n_walks, walk_len, inv_p, inv_q, node2vec, rule, seed, walk_id_base,
            walk_id_stride, out, err_count);
    else
        se::walk_kernel<false><<<(int)blocks, se::WALK_WPB * 32, 0, st>>>(
            rowptr, col, wcdf, symmetric, starts, n_walks, walk_len, inv_p, inv_q, node2vec, rule, seed, walk_id_base,
            walk_id_stride, out, err_count);
    return se::check_cuda(cudaGetLastError(), "walk_kernel launch");
}

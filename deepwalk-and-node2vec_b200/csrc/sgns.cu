// Skip-gram negative-sampling kernels for sm_100a (HBM-bound gather / dot / sigmoid / scatter).
//
// One GROUP of G lanes (G = 32 for emb >= 128) owns one centre: it keeps the centre row of W_in and the centre's
// accumulated gradient in registers for the whole window (read once, written once), and for every context gathers
// the context row and its K negative rows of W_out with 128-bit loads issued back to back (all rows of a chunk
// are in flight before the first is used), reduces the dot products with warp shuffles, applies
// sigmoid / clamp / log exactly as shallow_encoders/word2vec/loss.py:15-16, and scatters the updates in place.
//
// Three entry kernels share that body:
//   MODE_GRAD  explicit (inputs, targets, noise) -> loss sums + DENSE gradients of the mean loss (parity kernel,
//              word2vec/trainer.py:131-152 + autograd)
//   MODE_STEP  explicit batch, in-place SGD (noise optional: in-kernel Philox negatives)
//   MODE_WALK  tokens[n_seq x L] -> windows (word2vec/dataloader/torch_dataset.py:300-309) + in-kernel negatives
//              (word2vec/utils/sampling.py:21 distribution, or alias table) + in-place SGD.  Nothing materialised.
#include "sgns_common.cuh"

namespace se {

int launch_win_wide(const SgnsArgs &a, cudaStream_t stream);        // sgns_win_wide.cu: window-resident kernel for 128 < emb <= 256 (two float4 per lane)

namespace {

template <int MODE, int VEC, int G, int R>
__global__ void __launch_bounds__(SGNS_THREADS)
sgns_kernel(const SgnsArgs a) {
    constexpr int GROUPS_PER_BLOCK = SGNS_THREADS / G;
    constexpr int CH = (G < 8) ? G : (8 / R);       // target rows in flight per chunk; every lane owns <= 1 of them
    constexpr bool FAST = MODE != MODE_GRAD;
    const int lg = threadIdx.x & (G - 1);
    const int64_t gid = (int64_t)blockIdx.x * GROUPS_PER_BLOCK + (threadIdx.x / G);
    const int64_t n_groups = (int64_t)gridDim.x * GROUPS_PER_BLOCK;
    const int E = a.emb, N = a.n_ctx, K = a.n_neg, T = 1 + a.n_neg;
    const unsigned gmask = group_mask<G>();

    int eoff[R]; bool ok[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { eoff[j] = (lg + j * G) * VEC; ok[j] = eoff[j] < E; }

    float loss_pos = 0.f, loss_neg = 0.f;
    unsigned cnt_recall = 0, cnt_fp = 0, cnt_pairs = 0;

    // MODE_WALK: contiguous span of centres per group (neighbouring centres of a walk share rows -> L2 locality and
    // no intra-walk races); explicit modes: grid-stride.
    int64_t u_begin, u_end, u_step;
    if constexpr (MODE == MODE_WALK) {
        const int64_t span = (a.n_units + n_groups - 1) / n_groups;
        u_begin = gid * span; u_end = min(a.n_units, u_begin + span); u_step = 1;
    } else {
        u_begin = gid; u_end = a.n_units; u_step = n_groups;
    }

    for (int64_t u = u_begin; u < u_end; u += u_step) {
        int64_t crow; const int32_t *seq = nullptr; int pos = 0;
        if constexpr (MODE == MODE_WALK) {
            const int64_t s = u / a.n_cen;
            pos = a.radius + (int)(u - s * a.n_cen);
            seq = a.tokens + s * a.seq_len;
            crow = (int64_t)__ldg(seq + pos) + a.row_offset;
        } else {
            crow = __ldg(a.inputs + u);
        }
        float cen[R][VEC], acc[R][VEC];
        const float *cptr = a.w_in + crow * E;
#pragma unroll
        for (int j = 0; j < R; ++j) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) { cen[j][e] = 0.f; acc[j][e] = 0.f; }
            if (ok[j]) load_vec<VEC>(cptr + eoff[j], cen[j]);
        }

        for (int n = 0; n < N; ++n) {
            for (int t0 = 0; t0 < T; t0 += CH) {
                const int cnt = min(CH, T - t0);
                // lane `lg` resolves the row id of target t0 + lg (0 = the context, 1.. = negatives)
                int my = 0;
                if (lg < cnt) {
                    const int tj = t0 + lg;
                    if (tj == 0) {
                        if constexpr (MODE == MODE_WALK) {
                            const int off = (n < a.radius) ? (pos - a.radius + n) : (pos + 1 + n - a.radius);
                            my = __ldg(seq + off) + a.row_offset;
                        } else {
                            my = (int)__ldg(a.targets + u * N + n);
                        }
                    } else if (MODE != MODE_WALK && a.noise != nullptr) {
                        my = (int)__ldg(a.noise + (u * N + n) * K + (tj - 1));
                    } else {
                        const uint64_t cid = (uint64_t)(a.id_base + u);
                        const uint32_t r0 = pick_word(neg_words(a.seed, cid, n, tj - 1, STREAM_NEG), n & 3);
                        const uint32_t r1 = a.alias_prob ? pick_word(neg_words(a.seed, cid, n, tj - 1, STREAM_NEG_COIN), n & 3) : 0u;
                        my = neg_row(a, r0, r1);
                    }
                }
                int tid[CH];
                float row[CH][R][VEC];
                float dot[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) tid[c] = __shfl_sync(gmask, my, c, G);
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const float *rp = a.w_out + (int64_t)tid[c] * E;
#pragma unroll
                    for (int j = 0; j < R; ++j) {
#pragma unroll
                        for (int e = 0; e < VEC; ++e) row[c][j][e] = 0.f;
                        if (c < cnt && ok[j]) load_vec<VEC>(rp + eoff[j], row[c][j]);
                    }
                }
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    float d = 0.f;
#pragma unroll
                    for (int j = 0; j < R; ++j)
#pragma unroll
                        for (int e = 0; e < VEC; ++e) d = fmaf(row[c][j][e], cen[j][e], d);
                    dot[c] = group_sum<G>(d, gmask);
                }
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (c < cnt) {
                        const bool positive = (t0 + c) == 0;
                        const float s = dot[c];
                        float g;   // dL_pair / ds
                        if (positive) {
                            const float sig = sigmoidf_<FAST>(s);
                            loss_pos -= logf_<FAST>(fmaxf(sig, CLAMP_MIN));
                            g = (sig > CLAMP_MIN) ? -sigmoidf_<FAST>(-s) : 0.f;
                            cnt_recall += sig >= 0.5f;
                            cnt_pairs += 1;
                        } else {
                            const float sig_m = sigmoidf_<FAST>(-s);
                            loss_neg -= logf_<FAST>(fmaxf(sig_m, CLAMP_MIN));
                            const float sig = sigmoidf_<FAST>(s);
                            g = (sig_m > CLAMP_MIN) ? sig : 0.f;
                            cnt_fp += sig >= 0.5f;
                        }
                        if constexpr (MODE == MODE_GRAD) {
                            const float gs = g * a.grad_scale;
#pragma unroll
                            for (int j = 0; j < R; ++j) {
                                float d[VEC];
#pragma unroll
                                for (int e = 0; e < VEC; ++e) { acc[j][e] = fmaf(gs, row[c][j][e], acc[j][e]); d[e] = gs * cen[j][e]; }
                                if (a.grad_out && ok[j]) red_vec<VEC>(a.grad_out + (int64_t)tid[c] * E + eoff[j], d);
                            }
                            if (a.touch_out && lg == 0) mark_row(a.touch_out, a.list_out, a.touch_counts + 1, tid[c]);
                        } else {
                            const float step = -a.lr * g;
                            float *rp = a.w_out + (int64_t)tid[c] * E;
#pragma unroll
                            for (int j = 0; j < R; ++j) {
                                float d[VEC];
#pragma unroll
                                for (int e = 0; e < VEC; ++e) {
                                    acc[j][e] = fmaf(step, row[c][j][e], acc[j][e]);
                                    d[e] = step * cen[j][e];
                                }
                                if (ok[j]) {
                                    if (a.scatter_store) {
#pragma unroll
                                        for (int e = 0; e < VEC; ++e) d[e] += row[c][j][e];
                                        store_vec<VEC>(rp + eoff[j], d);
                                    } else {
                                        red_vec<VEC>(rp + eoff[j], d, a.sys_scope);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        // centre row: one write per centre for the whole window
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (!ok[j]) continue;
            if constexpr (MODE == MODE_GRAD) {
                if (a.grad_in) red_vec<VEC>(a.grad_in + crow * E + eoff[j], acc[j]);
                if (j == 0 && a.touch_in && lg == 0) mark_row(a.touch_in, a.list_in, a.touch_counts, crow);
            } else {
                float *cp = a.w_in + crow * E + eoff[j];
                if (a.scatter_store) {
                    float d[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) d[e] = cen[j][e] + acc[j][e];
                    store_vec<VEC>(cp, d);
                } else {
                    red_vec<VEC>(cp, acc[j], a.sys_scope);
                }
            }
        }
    }

    flush_stats(a.stats, lg == 0 && cnt_pairs != 0, loss_pos, loss_neg, cnt_recall, cnt_fp, cnt_pairs, (double)cnt_pairs * (double)K);
}


// ------------------------------------------------------------------------------------------------------------------
// Fast path for the in-place kernels (MODE_STEP / MODE_WALK) for rows longer than 128 floats or more than 7 negatives:
// one WARP per centre, R float4 per lane per row, CH = 8/R target rows per chunk.  Versus sgns_kernel it
//   * resolves the row ids of the NEXT chunk before working on the current one,
//   * reduces the CH dot products with a transposed butterfly (CH-1 + log2(32/CH) shuffles instead of 5*CH) that leaves
//     dot c in lane group c, where sigmoid / clamp / log are evaluated ONCE per chunk instead of once per row,
//   * draws negatives from one Philox call per four contexts (see neg_words),
//   * knows the row length at compile time when E == 128*R (no predication, shift addressing).
// ------------------------------------------------------------------------------------------------------------------

template <int MODE, int R, bool EXACT>
__global__ void __launch_bounds__(SGNS_THREADS, 2)
sgns_fast_kernel(const SgnsArgs a) {
    constexpr int VEC = 4;
    constexpr int CH = 8 / R;
    constexpr int SHIFT = (CH == 8) ? 2 : (CH == 4) ? 3 : (CH == 2) ? 4 : 5;   // lanes per owner group = 1 << SHIFT
    const int lane = threadIdx.x & 31;
    const int64_t gid = (int64_t)blockIdx.x * (SGNS_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t n_groups = (int64_t)gridDim.x * (SGNS_THREADS / 32);
    const int E = EXACT ? 128 * R : a.emb;
    const int N = a.n_ctx, K = a.n_neg, T = 1 + a.n_neg;
    const int n_chunks = (T + CH - 1) / CH;
    const int Q = N * n_chunks;                     // chunks per centre
    const bool cache_words = n_chunks == 1;         // one Philox call then serves 4 consecutive contexts
    const bool explicit_noise = (MODE != MODE_WALK) && a.noise != nullptr;
    const int owner_c = lane >> SHIFT;              // the row whose dot this lane holds after the reduction
    const bool owner_rep = (lane & ((1 << SHIFT) - 1)) == 0;

    int eoff[R]; bool ok[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { eoff[j] = (lane + j * 32) * VEC; ok[j] = EXACT || eoff[j] < E; }

    float loss_pos = 0.f, loss_neg = 0.f;
    unsigned cnt_recall = 0, cnt_fp = 0, cnt_pairs = 0;

    int64_t u_begin, u_end, u_step;
    if constexpr (MODE == MODE_WALK) {
        const int64_t span = (a.n_units + n_groups - 1) / n_groups;
        u_begin = gid * span; u_end = min(a.n_units, u_begin + span); u_step = 1;
    } else {
        u_begin = gid; u_end = a.n_units; u_step = n_groups;
    }

    uint4 w_bucket = make_uint4(0, 0, 0, 0), w_coin = make_uint4(0, 0, 0, 0);

    // row id owned by this lane (lane < cnt) in chunk (n, t0) of centre u; refreshes the cached Philox words
    auto resolve = [&](int64_t u, int n, int t0) -> int {
        const int cnt = min(CH, T - t0);
        int my = -1;
        const int tj = t0 + lane;
        // all lanes refresh together (the words are cached across calls, so no lane may skip a refresh it will need later)
        if (K > 0 && !explicit_noise && (!cache_words || (n & 3) == 0)) {
            const uint64_t cid = (uint64_t)(a.id_base + u);
            w_bucket = neg_words(a.seed, cid, n, max(tj - 1, 0), STREAM_NEG);
            if (a.alias_prob) w_coin = neg_words(a.seed, cid, n, max(tj - 1, 0), STREAM_NEG_COIN);
        }
        if (lane < cnt) {
            if (tj == 0) {
                if constexpr (MODE == MODE_WALK) {
                    const int64_t s = u / a.n_cen;
                    const int pos = a.radius + (int)(u - s * a.n_cen);
                    const int off = (n < a.radius) ? (pos - a.radius + n) : (pos + 1 + n - a.radius);
                    my = __ldg(a.tokens + s * a.seq_len + off) + a.row_offset;
                } else {
                    my = (int)__ldg(a.targets + u * N + n);
                }
            } else if (explicit_noise) {
                my = (int)__ldg(a.noise + (u * N + n) * K + (tj - 1));
            } else {
                my = neg_row(a, pick_word(w_bucket, n & 3), pick_word(w_coin, n & 3));
            }
        }
        return my;
    };
    auto centre_row = [&](int64_t u) -> int64_t {
        if constexpr (MODE == MODE_WALK) {
            const int64_t s = u / a.n_cen;
            return (int64_t)__ldg(a.tokens + s * a.seq_len + a.radius + (int)(u - s * a.n_cen)) + a.row_offset;
        } else {
            return __ldg(a.inputs + u);
        }
    };

    int my_next = -1;
    if (u_begin < u_end) {
        my_next = resolve(u_begin, 0, 0);
    }

    for (int64_t u = u_begin; u < u_end; u += u_step) {
        const int64_t crow = centre_row(u);
        float cen[R][VEC], acc[R][VEC];
        const float *cptr = a.w_in + crow * E;
#pragma unroll
        for (int j = 0; j < R; ++j) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) { cen[j][e] = 0.f; acc[j][e] = 0.f; }
            if (ok[j]) load_vec<VEC>(cptr + eoff[j], cen[j]);
        }

        int n = 0, t0 = 0;
        for (int q = 0; q < Q; ++q) {
            const int my = my_next;
            const int cnt = min(CH, T - t0);
            const bool positive_chunk = t0 == 0;
            // ---- look ahead one chunk: resolve its row ids (Philox / index loads) off the critical path --------------
            int n2 = n, t2 = t0 + CH;
            if (t2 >= T) { t2 = 0; n2 = n + 1; }
            if (n2 < N) {
                my_next = resolve(u, n2, t2);
            } else if (u + u_step < u_end) {
                my_next = resolve(u + u_step, 0, 0);
            } else {
                my_next = -1;
            }
            // ---- gather the chunk -----------------------------------------------------------------------------
            int tid[CH];
            float row[CH][R][VEC];
            float dot[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) tid[c] = __shfl_sync(FULL, my, c);
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float *rp = a.w_out + (int64_t)tid[c] * E;
#pragma unroll
                for (int j = 0; j < R; ++j) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) row[c][j][e] = 0.f;
                    if (c < cnt && ok[j]) load_vec<VEC>(rp + eoff[j], row[c][j]);
                }
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                float d = 0.f;
#pragma unroll
                for (int j = 0; j < R; ++j)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) d = fmaf(row[c][j][e], cen[j][e], d);
                dot[c] = d;
            }
            const float s = transposed_reduce<CH>(dot, lane);
            // ---- sigmoid / clamp / log once per chunk, in the lane group that owns the row (loss.py:15-16) ---------
            float step_mine = 0.f;
            if (owner_c < cnt) {
                const bool positive = positive_chunk && owner_c == 0;
                const float x = positive ? s : -s;                       // loss = -log clamp(sigmoid(x), 1e-6)
                const float e = __expf(-x);
                const float sig = __fdividef(1.0f, 1.0f + e);
                const bool live = sig > CLAMP_MIN;
                const float g = live ? (positive ? -(e * sig) : (e * sig)) : 0.f;   // dL/ds: -sigmoid(-s) | sigmoid(s)
                step_mine = -a.lr * g;
                if (owner_rep) {
                    const float l = -__logf(fmaxf(sig, CLAMP_MIN));
                    if (positive) { loss_pos += l; cnt_recall += x >= 0.f; cnt_pairs += 1; }
                    else { loss_neg += l; cnt_fp += x <= 0.f; }
                }
            }
            // ---- scatter ----------------------------------------------------------------------------------------
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float step = __shfl_sync(FULL, step_mine, c << SHIFT);
                if (c < cnt) {
                    float *rp = a.w_out + (int64_t)tid[c] * E;
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        float d[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            acc[j][e] = fmaf(step, row[c][j][e], acc[j][e]);
                            d[e] = step * cen[j][e];
                        }
                        if (ok[j]) {
                            if (a.scatter_store) {
#pragma unroll
                                for (int e = 0; e < VEC; ++e) d[e] += row[c][j][e];
                                store_vec<VEC>(rp + eoff[j], d);
                            } else {
                                red_vec<VEC>(rp + eoff[j], d, a.sys_scope);
                            }
                        }
                    }
                }
            }
            t0 = t2; n = n2;
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (!ok[j]) continue;
            float *cp = a.w_in + crow * E + eoff[j];
            if (a.scatter_store) {
                float d[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) d[e] = cen[j][e] + acc[j][e];
                store_vec<VEC>(cp, d);
            } else {
                red_vec<VEC>(cp, acc[j], a.sys_scope);
            }
        }
    }

    flush_stats(a.stats, loss_pos != 0.f || loss_neg != 0.f || cnt_pairs != 0 || cnt_fp != 0 || cnt_recall != 0, loss_pos, loss_neg, cnt_recall,
                cnt_fp, cnt_pairs, (double)cnt_pairs * (double)K);
}


// ------------------------------------------------------------------------------------------------------------------
// Specialised hot kernel: rows of at most 128 floats (one float4 per lane), T = 1 + K <= 8 target rows per context
// known at compile time (lane t < T owns target t: 0 = context, 1.. = negatives), contexts processed in groups of
// four that share one Philox call.  All T row gathers of a context are issued back to back, the T dots are reduced
// with one transposed butterfly, sigmoid / clamp / log run once per context in the owner lanes.
// Measured on S3 (B200): explicit L2 prefetching of upcoming rows (prefetch.global.L2 / cp.async.bulk.prefetch.L2, one
// context or one group ahead) LOWERS throughput by 20-25 % -- the prefetched lines thrash the L2 that the scatter's
// dirty lines and the context-row reuse live in -- so the kernel relies on occupancy (24 warps/SM x T rows in flight).
// The prefetch path is kept behind a developer flag (flags bit 0x200) for re-measurement on other shapes.
// ------------------------------------------------------------------------------------------------------------------
template <int MODE, int T, bool EXACT>
__global__ void __launch_bounds__(SGNS_THREADS, 2)
sgns_ctx_kernel(const SgnsArgs a) {
    constexpr int K = T - 1;
    constexpr int P = (T <= 1) ? 1 : (T <= 2) ? 2 : (T <= 4) ? 4 : 8;                 // dots padded to a power of two
    constexpr int SHIFT = (P == 8) ? 2 : (P == 4) ? 3 : (P == 2) ? 4 : 5;             // lanes per owner group = 1 << SHIFT
    const int lane = threadIdx.x & 31;
    const int64_t gid = (int64_t)blockIdx.x * (SGNS_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t n_groups = (int64_t)gridDim.x * (SGNS_THREADS / 32);
    const int E = EXACT ? 128 : a.emb;
    const int eoff = lane * 4;
    const bool ok = EXACT || eoff < E;
    const int N = a.n_ctx, NG = (a.n_ctx + 3) >> 2;
    const bool explicit_noise = (MODE != MODE_WALK) && a.noise != nullptr;
    const int owner_t = lane >> SHIFT;
    const bool owner_rep = (lane & ((1 << SHIFT) - 1)) == 0;

    float loss_pos = 0.f, loss_neg = 0.f;
    unsigned cnt_recall = 0, cnt_fp = 0, cnt_pairs = 0;

    int64_t u_begin, u_end, u_step;
    if constexpr (MODE == MODE_WALK) {
        const int64_t span = (a.n_units + n_groups - 1) / n_groups;
        u_begin = gid * span; u_end = min(a.n_units, u_begin + span); u_step = 1;
    } else {
        u_begin = gid; u_end = a.n_units; u_step = n_groups;
    }

    auto centre_row = [&](int64_t u) -> int64_t {
        if constexpr (MODE == MODE_WALK) {
            const int64_t s = u / a.n_cen;
            return (int64_t)__ldg(a.tokens + s * a.seq_len + a.radius + (int)(u - s * a.n_cen)) + a.row_offset;
        } else {
            return __ldg(a.inputs + u);
        }
    };
    // ids of the rows this lane owns in contexts 4g .. 4g+3 of centre u
    auto resolve_group = [&](int64_t u, int g, int (&ids)[4]) {
        uint4 wb = make_uint4(0, 0, 0, 0), wc = make_uint4(0, 0, 0, 0);
        if (K > 0 && !explicit_noise) {
            const uint64_t cid = (uint64_t)(a.id_base + u);
            const int k = lane >= 1 ? lane - 1 : 0;
            wb = neg_words(a.seed, cid, g * 4, k, STREAM_NEG);
            if (a.alias_prob) wc = neg_words(a.seed, cid, g * 4, k, STREAM_NEG_COIN);
        }
        const int32_t *seq = nullptr; int pos = 0;
        if constexpr (MODE == MODE_WALK) {
            const int64_t s = u / a.n_cen;
            pos = a.radius + (int)(u - s * a.n_cen);
            seq = a.tokens + s * a.seq_len;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = g * 4 + j;
            int id = 0;
            if (n < N) {
                if (lane == 0) {
                    if constexpr (MODE == MODE_WALK) {
                        const int off = (n < a.radius) ? (pos - a.radius + n) : (pos + 1 + n - a.radius);
                        id = __ldg(seq + off) + a.row_offset;
                    } else {
                        id = (int)__ldg(a.targets + u * N + n);
                    }
                } else if (lane < T) {
                    if (explicit_noise) id = (int)__ldg(a.noise + (u * N + n) * K + (lane - 1));
                    else id = neg_row(a, pick_word(wb, j), pick_word(wc, j));
                }
            }
            ids[j] = id;
        }
    };

    int cur[4] = {0, 0, 0, 0};

    for (int64_t u = u_begin; u < u_end; u += u_step) {
        const int64_t crow = centre_row(u);
        float cen[4] = {0.f, 0.f, 0.f, 0.f}, acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok) load_vec<4>(a.w_in + crow * E + eoff, cen);

        for (int g = 0; g < NG; ++g) {
            resolve_group(u, g, cur);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (g * 4 + j < N) {
                    int tid[T];
                    float row[T][4];
                    float dot[P];
#pragma unroll
                    for (int t = 0; t < T; ++t) tid[t] = __shfl_sync(FULL, cur[j], t);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        row[t][0] = row[t][1] = row[t][2] = row[t][3] = 0.f;
                        if (ok) load_vec<4>(a.w_out + (int64_t)tid[t] * E + eoff, row[t]);
                    }
#pragma unroll
                    for (int t = 0; t < P; ++t) {
                        float d = 0.f;
                        if (t < T) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) d = fmaf(row[t][e], cen[e], d);
                        }
                        dot[t] = d;
                    }
                    const float s = transposed_reduce<P>(dot, lane);
                    float step_mine = 0.f;
                    if (owner_t < T) {
                        const bool positive = owner_t == 0;
                        const float x = positive ? s : -s;                        // loss = -log clamp(sigmoid(x), 1e-6)
                        const float ex = __expf(-x);
                        const float sig = __fdividef(1.0f, 1.0f + ex);
                        const bool live = sig > CLAMP_MIN;
                        const float gmag = live ? ex * sig : 0.f;                 // |dL/ds| = sigmoid(-x)
                        step_mine = positive ? a.lr * gmag : -a.lr * gmag;        // -lr * dL/ds
                        if (owner_rep) {
                            const float l = -__logf(fmaxf(sig, CLAMP_MIN));
                            if (positive) { loss_pos += l; cnt_recall += x >= 0.f; cnt_pairs += 1; }
                            else { loss_neg += l; cnt_fp += x <= 0.f; }
                        }
                    }
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const float step = __shfl_sync(FULL, step_mine, t << SHIFT);
                        float d[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[e] = fmaf(step, row[t][e], acc[e]);
                        if (ok) {
                            float *rp = a.w_out + (int64_t)tid[t] * E + eoff;
                            if (a.scatter_store) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) d[e] = fmaf(step, cen[e], row[t][e]);
                                store_vec<4>(rp, d);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) d[e] = step * cen[e];
                                red_vec<4>(rp, d, a.sys_scope);
                            }
                        }
                    }
                }
            }
        }
        if (ok) {
            float *cp = a.w_in + crow * E + eoff;
            if (a.scatter_store) {
                float d[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) d[e] = cen[e] + acc[e];
                store_vec<4>(cp, d);
            } else {
                red_vec<4>(cp, acc, a.sys_scope);
            }
        }
    }

    flush_stats(a.stats, loss_pos != 0.f || loss_neg != 0.f || cnt_pairs != 0 || cnt_fp != 0 || cnt_recall != 0, loss_pos, loss_neg, cnt_recall,
                cnt_fp, cnt_pairs, (double)cnt_pairs * (double)K);
}


// ------------------------------------------------------------------------------------------------------------------
// Window-resident kernel family: sgns_win.cuh, instantiated per lane-group width in sgns_win_g*.cu (16 <= emb <= 128), and
// sgns_win_wide.cu (128 < emb <= 256, two float4 per lane).
// Returns SE_ERR_UNSUPPORTED when the shape is not covered (caller tries the next kernel).
// ------------------------------------------------------------------------------------------------------------------
int launch_win(const SgnsArgs &a, cudaStream_t stream) {
    if (a.emb % 4 != 0 || a.emb < 16 || a.emb > 256 || a.n_neg > 7 || a.radius > 8 || a.scatter_store || a.no_window) return SE_ERR_UNSUPPORTED;
    if (((uintptr_t)a.w_in % 16) || ((uintptr_t)a.w_out % 16)) return SE_ERR_UNSUPPORTED;
    int rc = SE_ERR_UNSUPPORTED;
    if (a.emb > 128) {
        rc = launch_win_wide(a, stream);
    } else if (a.emb > 64) {
        rc = launch_win_g32(a, stream);
    } else if (a.emb > 32) {
        rc = launch_win_g16(a, stream);
    } else {
        rc = launch_win_g8(a, stream);
    }
    return rc;
}


// ------------------------------------------------------------------------------------------------------------------
// Owner-computes negatives (striped tables, the reference's GLOBAL negative distribution).  Fetching the K negative rows
// of a pair from their owners costs K rows each way over NVLink (measured: ~0.2 G pairs/s per GPU).  Here the centre row
// travels instead: every GPU walks over the centres of ALL GPUs' walks (tokens are all-gathered, 4 bytes per centre),
// re-draws each centre's N*K negatives from the same Philox keys as the single-GPU kernel, keeps the ones whose rows it
// OWNS, and for those computes dot / sigmoid / update against its local HBM; the centre row is read once per centre
// (peer load) and the centre's accumulated gradient goes back with one peer red.add -- 2 rows per centre per GPU over
// NVLink instead of 2*N*K.  Positive pairs are done by the walk's home GPU (sgns_win_kernel with K = 0).  Lane l < NG*K
// draws the ids of negative (l % K) for the four contexts of group (l / K) with ONE Philox call, exactly the keying of
// `neg_words`; owned ids are compacted into a per-warp list and processed eight rows at a time.
// ------------------------------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void __launch_bounds__(SGNS_THREADS, 2)
sgns_negown_kernel(const SgnsArgs a) {
    constexpr int P = 8, SHIFT = 2;
    __shared__ int own_list[SGNS_THREADS / 32][128];      // up to 32 drawing lanes x 4 contexts owned at once (world = 1)
    const int lane = threadIdx.x & 31;
    int *list = own_list[threadIdx.x >> 5];
    const int64_t gid = (int64_t)blockIdx.x * (SGNS_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t n_groups = (int64_t)gridDim.x * (SGNS_THREADS / 32);
    const int E = EXACT ? 128 : a.emb;
    const int eoff = lane * 4;
    const bool ok = EXACT || eoff < E;
    const int N = a.n_ctx, K = a.n_neg, NG = (a.n_ctx + 3) >> 2;
    const int l_g = lane / K, l_k = lane - l_g * K;
    const bool drawer = lane < NG * K;
    const int owner_t = lane >> SHIFT;
    const bool owner_rep = (lane & 3) == 0;
    const uint32_t world = (uint32_t)a.neg_world, me = (uint32_t)a.neg_rank;
    const int shift = a.own_shift;

    float loss_neg = 0.f;
    unsigned cnt_fp = 0, cnt_neg = 0;

    const int64_t span = (a.n_units + n_groups - 1) / n_groups;
    const int64_t u_end = min(a.n_units, gid * span + span);
    for (int64_t u = gid * span; u < u_end; ++u) {
        // ---- which of this centre's N*K negatives live in my HBM? --------------------------------------------------
        int ids[4] = {0, 0, 0, 0};
        bool own[4] = {false, false, false, false};
        if (drawer) {
            const uint64_t cid = (uint64_t)(a.id_base + u);
            const uint4 wb = neg_words(a.seed, cid, l_g * 4, l_k, STREAM_NEG);
            uint4 wc = make_uint4(0, 0, 0, 0);
            if (a.alias_prob) wc = neg_words(a.seed, cid, l_g * 4, l_k, STREAM_NEG_COIN);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (l_g * 4 + j < N) {
                    const uint32_t id = (uint32_t)draw_row(a.alias_prob, a.alias_idx, (uint32_t)a.vocab, pick_word(wb, j), pick_word(wc, j));
                    ids[j] = (int)id;
                    own[j] = ((id >> shift) % world) == me;
                }
            }
        }
        const int cnt = (int)own[0] + (int)own[1] + (int)own[2] + (int)own[3];
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total == 0) continue;
        __syncwarp();
        {
            int pos = incl - cnt;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (own[j]) list[pos++] = ids[j];
        }
        __syncwarp();

        const int64_t sq = u / a.n_cen;
        const int64_t crow = (int64_t)__ldg(a.tokens + sq * a.seq_len + a.radius + (int)(u - sq * a.n_cen)) + a.row_offset;
        float cen[4] = {0.f, 0.f, 0.f, 0.f}, acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok) load_vec<4>(a.w_in + crow * E + eoff, cen);

        for (int c0 = 0; c0 < total; c0 += P) {
            const int m = min(P, total - c0);
            int tid[P];
            float row[P][4];
            float dot[P];
#pragma unroll
            for (int t = 0; t < P; ++t) tid[t] = (t < m) ? list[c0 + t] : 0;
#pragma unroll
            for (int t = 0; t < P; ++t) {
                row[t][0] = row[t][1] = row[t][2] = row[t][3] = 0.f;
                if (t < m && ok) load_vec<4>(a.w_out + (int64_t)tid[t] * E + eoff, row[t]);
            }
#pragma unroll
            for (int t = 0; t < P; ++t) {
                float d = 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) d = fmaf(row[t][e], cen[e], d);
                dot[t] = d;
            }
            const float sc = transposed_reduce<P>(dot, lane);
            float step_mine = 0.f;
            if (owner_t < m) {
                const float x = -sc;                                          // loss = -log clamp(sigmoid(-s), 1e-6)   (loss.py:16)
                const float ex = __expf(-x);
                const float sig = __fdividef(1.0f, 1.0f + ex);
                const float gmag = (sig > CLAMP_MIN) ? ex * sig : 0.f;        // dL/ds = sigmoid(s)
                step_mine = -a.lr * gmag;
                if (owner_rep) { loss_neg -= __logf(fmaxf(sig, CLAMP_MIN)); cnt_fp += x <= 0.f; cnt_neg += 1; }
            }
#pragma unroll
            for (int t = 0; t < P; ++t) {
                const float step = __shfl_sync(FULL, step_mine, t << SHIFT);
                if (t < m) {
                    float d[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { acc[e] = fmaf(step, row[t][e], acc[e]); d[e] = step * cen[e]; }
                    if (ok) red_vec<4>(a.w_out + (int64_t)tid[t] * E + eoff, d, false);      // my own HBM: device scope is enough
                }
            }
        }
        if (ok) red_vec<4>(a.w_in + crow * E + eoff, acc, a.sys_scope);
    }

    flush_stats(a.stats, cnt_neg != 0, 0.f, loss_neg, 0u, cnt_fp, 0u, (double)cnt_neg);
}

template <bool EXACT>
int launch_negown(const SgnsArgs &a, cudaStream_t stream) {
    auto kern = sgns_negown_kernel<EXACT>;
    const int blocks = persistent_blocks(kern, 0, a.n_units, SGNS_THREADS / 32);
    if (blocks <= 0) return SE_ERR_CUDA;
    kern<<<blocks, SGNS_THREADS, 0, stream>>>(a);
    return check_cuda(cudaGetLastError(), "sgns_negown_kernel launch");
}

template <int MODE, int T, bool EXACT>
int launch_ctx_one(const SgnsArgs &a, cudaStream_t stream) {
    auto kern = sgns_ctx_kernel<MODE, T, EXACT>;
    const int blocks = persistent_blocks(kern, 0, a.n_units, SGNS_THREADS / 32);
    if (blocks <= 0) return SE_ERR_CUDA;
    kern<<<blocks, SGNS_THREADS, 0, stream>>>(a);
    return check_cuda(cudaGetLastError(), "sgns_ctx_kernel launch");
}

template <int MODE, bool EXACT>
int launch_ctx_t(const SgnsArgs &a, cudaStream_t stream) {
    switch (1 + a.n_neg) {
        case 1: return launch_ctx_one<MODE, 1, EXACT>(a, stream);
        case 2: return launch_ctx_one<MODE, 2, EXACT>(a, stream);
        case 3: return launch_ctx_one<MODE, 3, EXACT>(a, stream);
        case 4: return launch_ctx_one<MODE, 4, EXACT>(a, stream);
        case 5: return launch_ctx_one<MODE, 5, EXACT>(a, stream);
        case 6: return launch_ctx_one<MODE, 6, EXACT>(a, stream);
        case 7: return launch_ctx_one<MODE, 7, EXACT>(a, stream);
        case 8: return launch_ctx_one<MODE, 8, EXACT>(a, stream);
        default: return SE_ERR_UNSUPPORTED;
    }
}

// Returns SE_ERR_UNSUPPORTED when the shape is not covered (caller tries the next kernel).
template <int MODE>
int launch_ctx(const SgnsArgs &a, cudaStream_t stream) {
    if (a.emb % 4 != 0 || a.emb <= 32 || a.emb > 128 || a.n_neg > 7) return SE_ERR_UNSUPPORTED;
    if (((uintptr_t)a.w_in % 16) || ((uintptr_t)a.w_out % 16)) return SE_ERR_UNSUPPORTED;
    return a.emb == 128 ? launch_ctx_t<MODE, true>(a, stream) : launch_ctx_t<MODE, false>(a, stream);
}

template <int MODE, int R, bool EXACT>
int launch_fast_one(const SgnsArgs &a, cudaStream_t stream) {
    auto kern = sgns_fast_kernel<MODE, R, EXACT>;
    const int blocks = persistent_blocks(kern, 0, a.n_units, SGNS_THREADS / 32);
    if (blocks <= 0) return SE_ERR_CUDA;
    kern<<<blocks, SGNS_THREADS, 0, stream>>>(a);
    return check_cuda(cudaGetLastError(), "sgns_fast_kernel launch");
}

// Returns SE_ERR_UNSUPPORTED when the shape is not covered (caller falls back to sgns_kernel).
template <int MODE>
int launch_fast(const SgnsArgs &a, cudaStream_t stream) {
    if (a.emb % 4 != 0 || a.emb <= 64 || a.emb > 1024 || a.n_neg > 64) return SE_ERR_UNSUPPORTED;
    if (((uintptr_t)a.w_in % 16) || ((uintptr_t)a.w_out % 16)) return SE_ERR_UNSUPPORTED;
    const int nvec = a.emb / 4;
    if (a.emb == 128) return launch_fast_one<MODE, 1, true>(a, stream);
    if (a.emb == 256) return launch_fast_one<MODE, 2, true>(a, stream);
    if (a.emb == 512) return launch_fast_one<MODE, 4, true>(a, stream);
    if (a.emb == 1024) return launch_fast_one<MODE, 8, true>(a, stream);
    if (nvec <= 32) return launch_fast_one<MODE, 1, false>(a, stream);
    if (nvec <= 64) return launch_fast_one<MODE, 2, false>(a, stream);
    if (nvec <= 128) return launch_fast_one<MODE, 4, false>(a, stream);
    return launch_fast_one<MODE, 8, false>(a, stream);
}

template <int MODE, int VEC, int G, int R>
int launch_one(const SgnsArgs &a, cudaStream_t stream) {
    auto kern = sgns_kernel<MODE, VEC, G, R>;
    const int blocks = persistent_blocks(kern, 0, a.n_units, SGNS_THREADS / G);      // exactly one resident wave at most
    if (blocks <= 0) return SE_ERR_CUDA;
    kern<<<blocks, SGNS_THREADS, 0, stream>>>(a);
    return check_cuda(cudaGetLastError(), "sgns_kernel launch");
}

template <int MODE, int VEC>
int launch_vec(const SgnsArgs &a, int nvec, cudaStream_t stream) {
    if (nvec <= 1) return launch_one<MODE, VEC, 1, 1>(a, stream);
    if (nvec <= 2) return launch_one<MODE, VEC, 2, 1>(a, stream);
    if (nvec <= 4) return launch_one<MODE, VEC, 4, 1>(a, stream);
    if (nvec <= 8) return launch_one<MODE, VEC, 8, 1>(a, stream);
    if (nvec <= 16) return launch_one<MODE, VEC, 16, 1>(a, stream);
    if (nvec <= 32) return launch_one<MODE, VEC, 32, 1>(a, stream);
    if (nvec <= 64) return launch_one<MODE, VEC, 32, 2>(a, stream);
    if (nvec <= 128) return launch_one<MODE, VEC, 32, 4>(a, stream);
    if (nvec <= 256) return launch_one<MODE, VEC, 32, 8>(a, stream);
    set_error("embedding size %d not supported (max %d)", a.emb, 256 * VEC);
    return SE_ERR_UNSUPPORTED;
}

template <int MODE>
int launch(const SgnsArgs &a, cudaStream_t stream) {
    if (a.n_units <= 0) return SE_OK;
    if constexpr (MODE != MODE_GRAD) {
        if (!a.force_generic) {
            int rc = SE_ERR_UNSUPPORTED;
            if constexpr (MODE == MODE_WALK) rc = launch_win(a, stream);
            if (rc != SE_ERR_UNSUPPORTED) return rc;
            rc = launch_ctx<MODE>(a, stream);
            if (rc != SE_ERR_UNSUPPORTED) return rc;
            rc = launch_fast<MODE>(a, stream);
            if (rc != SE_ERR_UNSUPPORTED) return rc;
        }
    }
    const bool al16 = ((uintptr_t)a.w_in % 16 == 0) && ((uintptr_t)a.w_out % 16 == 0) &&
                      (MODE != MODE_GRAD || (((uintptr_t)a.grad_in % 16 == 0) && ((uintptr_t)a.grad_out % 16 == 0)));
    const bool al8 = ((uintptr_t)a.w_in % 8 == 0) && ((uintptr_t)a.w_out % 8 == 0) &&
                     (MODE != MODE_GRAD || (((uintptr_t)a.grad_in % 8 == 0) && ((uintptr_t)a.grad_out % 8 == 0)));
    if (a.emb % 4 == 0 && al16) return launch_vec<MODE, 4>(a, a.emb / 4, stream);
    if (a.emb % 2 == 0 && al8) return launch_vec<MODE, 2>(a, a.emb / 2, stream);
    return launch_vec<MODE, 1>(a, a.emb, stream);
}

// SkipGram.forward: one group per (b, j) score would starve small batches; one group per b, loop over m.
template <int G>
__global__ void __launch_bounds__(256)
scores_kernel(const float *__restrict__ w_in, const float *__restrict__ w_out, int emb, const int64_t *__restrict__ inputs,
              const int64_t *__restrict__ outputs, int64_t total, int m, int proba, float *__restrict__ out) {
    const int lg = threadIdx.x & (G - 1);
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / G;
    for (int64_t i = gid; i < total; i += n_groups) {
        const float *c = w_in + __ldg(inputs + i / m) * emb;
        const float *o = w_out + __ldg(outputs + i) * emb;
        float d = 0.f;
        for (int e = lg; e < emb; e += G) d = fmaf(__ldcg(o + e), __ldcg(c + e), d);
        d = group_sum<G>(d, group_mask<G>());
        if (lg == 0) out[i] = proba ? 1.0f / (1.0f + expf(-d)) : d;
    }
}

// Backward of scores_kernel (proba = 0), dense accumulate: one warp per batch row.
__global__ void __launch_bounds__(256)
scores_backward_kernel(const float *__restrict__ w_in, const float *__restrict__ w_out, int emb,
                       const int64_t *__restrict__ inputs, const int64_t *__restrict__ outputs, int64_t batch, int m,
                       const float *__restrict__ grad_scores, float *__restrict__ grad_in, float *__restrict__ grad_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < batch; b += n_warps) {
        const int64_t crow = __ldg(inputs + b);
        for (int e0 = 0; e0 < emb; e0 += 32) {
            const int e = e0 + lane;
            const float c = e < emb ? __ldcg(w_in + crow * emb + e) : 0.f;
            float acc = 0.f;
            for (int j = 0; j < m; ++j) {
                const int64_t orow = __ldg(outputs + b * m + j);
                const float g = __ldg(grad_scores + b * m + j);
                if (e < emb) {
                    acc = fmaf(g, __ldcg(w_out + orow * emb + e), acc);
                    atomicAdd(grad_out + orow * emb + e, g * c);
                }
            }
            if (e < emb) atomicAdd(grad_in + crow * emb + e, acc);
        }
    }
}

// NegativeSamplingLoss.forward on logits (loss.py:14-22) + d(mean loss)/d(logits); one thread per (b, n).
__global__ void __launch_bounds__(256)
ns_loss_kernel(const float *__restrict__ pos, const float *__restrict__ neg, int64_t pairs, int n_neg, float scale,
               double *__restrict__ stats, float *__restrict__ grad_pos, float *__restrict__ grad_neg) {
    float lp = 0.f, ln = 0.f;
    unsigned rec = 0, fp = 0, cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (int64_t)gridDim.x * blockDim.x) {
        const float s = pos[i];
        const float sig = 1.0f / (1.0f + expf(-s));
        lp -= logf(fmaxf(sig, CLAMP_MIN));
        if (grad_pos) grad_pos[i] = (sig > CLAMP_MIN) ? -scale / (1.0f + expf(s)) : 0.f;
        rec += sig >= 0.5f;
        cnt += 1;
        for (int k = 0; k < n_neg; ++k) {
            const float z = neg[i * n_neg + k];
            const float sm = 1.0f / (1.0f + expf(z)), sg = 1.0f / (1.0f + expf(-z));
            ln -= logf(fmaxf(sm, CLAMP_MIN));
            if (grad_neg) grad_neg[i * n_neg + k] = (sm > CLAMP_MIN) ? scale * sg : 0.f;
            fp += sg >= 0.5f;
        }
    }
    flush_stats(stats, cnt != 0, lp, ln, rec, fp, cnt, (double)cnt * n_neg);
}

int common_checks(const char *fn, const void *w_in, const void *w_out, int64_t vocab, int emb, int n_neg) {
    if (!w_in || !w_out) { set_error("%s: null embedding table", fn); return SE_ERR_INVALID_ARG; }
    if (vocab < 1 || vocab > 0x7fffffffll) { set_error("%s: vocab %lld out of range", fn, (long long)vocab); return SE_ERR_INVALID_ARG; }
    if (emb < 1) { set_error("%s: embedding size must be >= 1", fn); return SE_ERR_INVALID_ARG; }
    if (n_neg < 0) { set_error("%s: negative sample count must be >= 0", fn); return SE_ERR_INVALID_ARG; }
    return SE_OK;
}

}  // namespace
}  // namespace se

extern "C" int se_skipgram_scores(const float *w_in, const float *w_out, int64_t vocab, int emb, const int64_t *inputs,
                                  const int64_t *outputs, int64_t batch, int m, int proba, float *out, void *stream) {
    int rc = se::common_checks("se_skipgram_scores", w_in, w_out, vocab, emb, 0);
    if (rc != SE_OK) return rc;
    SE_REQUIRE(batch >= 0 && m >= 1, "se_skipgram_scores: bad shape");
    const int64_t total = batch * m;
    if (total == 0) return SE_OK;
    SE_REQUIRE(inputs && outputs && out, "se_skipgram_scores: null pointer");
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    if (emb >= 32) {
        int64_t blocks = (total + 7) / 8; if (blocks > sms * 8) blocks = sms * 8;
        se::scores_kernel<32><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w_in, w_out, emb, inputs, outputs, total, m, proba, out);
    } else {
        int64_t blocks = (total + 63) / 64; if (blocks > sms * 8) blocks = sms * 8;
        se::scores_kernel<4><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w_in, w_out, emb, inputs, outputs, total, m, proba, out);
    }
    return se::check_cuda(cudaGetLastError(), "scores_kernel launch");
}

extern "C" int se_sgns_grad(const float *w_in, const float *w_out, int64_t vocab, int emb, const int64_t *inputs,
                            const int64_t *targets, const int64_t *noise, int64_t batch, int n_ctx, int n_neg,
                            double *stats, float *grad_in, float *grad_out, void *stream) {
    int rc = se::common_checks("se_sgns_grad", w_in, w_out, vocab, emb, n_neg);
    if (rc != SE_OK) return rc;
    SE_REQUIRE(batch >= 0 && n_ctx >= 1, "se_sgns_grad: bad batch shape");
    if (batch == 0) return SE_OK;
    SE_REQUIRE(inputs && targets && (noise || n_neg == 0), "se_sgns_grad: null index tensor");
    SE_REQUIRE((grad_in == nullptr) == (grad_out == nullptr), "se_sgns_grad: pass both gradient buffers or neither");
    se::SgnsArgs a{};
    a.w_in = const_cast<float *>(w_in); a.w_out = const_cast<float *>(w_out);
    a.grad_in = grad_in; a.grad_out = grad_out;
    a.inputs = inputs; a.targets = targets; a.noise = noise;
    a.stats = stats; a.n_units = batch; a.vocab = vocab; a.emb = emb; a.n_ctx = n_ctx; a.n_neg = n_neg;
    a.grad_scale = batch > 0 ? 1.0f / (float)(batch * n_ctx) : 0.f;
    a.neg_vocab = (uint32_t)vocab; a.neg_shift = -1; a.neg_world = 1;
    return se::launch<se::MODE_GRAD>(a, (cudaStream_t)stream);
}

extern "C" int se_sgns_adam_step(float *w_in, float *w_out, int64_t vocab, int emb, const int64_t *inputs, const int64_t *targets,
                                 const int64_t *noise, int64_t batch, int n_ctx, int n_neg, const se_adam_state *st, float lr,
                                 float beta1, float beta2, float eps, double *stats, void *stream) {
    int rc = se::common_checks("se_sgns_adam_step", w_in, w_out, vocab, emb, n_neg);
    if (rc != SE_OK) return rc;
    SE_REQUIRE(batch >= 0 && n_ctx >= 1, "se_sgns_adam_step: bad batch shape");
    if (batch == 0) return SE_OK;
    SE_REQUIRE(inputs && targets && (noise || n_neg == 0), "se_sgns_adam_step: null index tensor");
    SE_REQUIRE(st && st->m_in && st->v_in && st->m_out && st->v_out && st->g_in && st->g_out && st->t_in && st->t_out && st->touched_in &&
               st->touched_out && st->list_in && st->list_out && st->counts, "se_sgns_adam_step: incomplete optimiser state");
    const int64_t need_in = batch < vocab ? batch : vocab;
    const int64_t tgt = batch * n_ctx * (int64_t)(1 + n_neg);
    const int64_t need_out = tgt < vocab ? tgt : vocab;
    SE_REQUIRE(st->list_in_capacity >= need_in && st->list_out_capacity >= need_out,
               "se_sgns_adam_step: row lists too small (need %lld / %lld entries)", (long long)need_in, (long long)need_out);
    SE_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "se_sgns_adam_step: bad hyper-parameters");
    se::SgnsArgs a{};
    a.w_in = w_in; a.w_out = w_out; a.grad_in = st->g_in; a.grad_out = st->g_out;
    a.inputs = inputs; a.targets = targets; a.noise = noise;
    a.stats = stats; a.n_units = batch; a.vocab = vocab; a.emb = emb; a.n_ctx = n_ctx; a.n_neg = n_neg;
    a.grad_scale = 1.0f / (float)(batch * n_ctx);                                   // the reference's MEAN loss (loss.py:19)
    a.neg_vocab = (uint32_t)vocab; a.neg_shift = -1; a.neg_world = 1;
    a.touch_in = st->touched_in; a.touch_out = st->touched_out; a.list_in = st->list_in; a.list_out = st->list_out; a.touch_counts = st->counts;
    rc = se::launch<se::MODE_GRAD>(a, (cudaStream_t)stream);
    if (rc != SE_OK) return rc;
    rc = se::adam_apply(w_in, st->m_in, st->v_in, st->g_in, st->t_in, st->touched_in, st->list_in, st->counts, st->counts + 2, need_in, emb, lr,
                        beta1, beta2, eps, (cudaStream_t)stream);
    if (rc != SE_OK) return rc;
    return se::adam_apply(w_out, st->m_out, st->v_out, st->g_out, st->t_out, st->touched_out, st->list_out, st->counts + 1, st->counts + 3, need_out,
                          emb, lr, beta1, beta2, eps, (cudaStream_t)stream);
}

extern "C" int se_sgns_step(float *w_in, float *w_out, int64_t vocab, int emb, const int64_t *inputs,
                            const int64_t *targets, const int64_t *noise, int64_t batch, int n_ctx, int n_neg,
                            const float *alias_prob, const int32_t *alias_idx, float lr, uint64_t seed,
                            int64_t pair_id_base, int flags, double *stats, void *stream) {
    int rc = se::common_checks("se_sgns_step", w_in, w_out, vocab, emb, n_neg);
    if (rc != SE_OK) return rc;
    SE_REQUIRE(batch >= 0 && n_ctx >= 1, "se_sgns_step: bad batch shape");
    if (batch == 0) return SE_OK;
    SE_REQUIRE(inputs && targets, "se_sgns_step: null index tensor");
    SE_REQUIRE((alias_prob == nullptr) == (alias_idx == nullptr), "se_sgns_step: pass both alias arrays or neither");
    SE_REQUIRE((flags & ~(SE_SGNS_SCATTER_STORE | SE_SGNS_GENERIC_KERNEL)) == 0, "se_sgns_step: unknown flags %d", flags);
    se::SgnsArgs a{};
    a.w_in = w_in; a.w_out = w_out; a.inputs = inputs; a.targets = targets; a.noise = noise;
    a.alias_prob = alias_prob; a.alias_idx = alias_idx;
    a.stats = stats; a.n_units = batch; a.vocab = vocab; a.emb = emb; a.n_ctx = n_ctx; a.n_neg = n_neg;
    a.lr = lr; a.seed = seed; a.id_base = pair_id_base; a.scatter_store = (flags & SE_SGNS_SCATTER_STORE) != 0;
    a.force_generic = (flags & SE_SGNS_GENERIC_KERNEL) != 0;
    a.neg_vocab = (uint32_t)vocab; a.neg_shift = -1; a.neg_world = 1;
    return se::launch<se::MODE_STEP>(a, (cudaStream_t)stream);
}

// rows of [0, vocab) owned by spec->rank: stripes s with s % world == rank
static int64_t shard_local_rows(int64_t vocab, const se_shard_spec *spec) {
    const int64_t sr = spec->stripe_rows;
    const int64_t n_stripes = (vocab + sr - 1) / sr;
    if (spec->rank >= n_stripes) return 0;
    const int64_t mine = (n_stripes - 1 - spec->rank) / spec->world + 1;      // stripes rank, rank + world, ...
    const int64_t last = spec->rank + (mine - 1) * spec->world;               // the last one may be partial
    return (mine - 1) * sr + ((last == n_stripes - 1) ? vocab - last * sr : sr);
}

extern "C" int se_shard_local_rows(int64_t vocab, const se_shard_spec *spec, int64_t *n_rows) {
    SE_REQUIRE(spec && n_rows && vocab >= 1, "se_shard_local_rows: bad arguments");
    SE_REQUIRE(spec->world >= 1 && spec->rank >= 0 && spec->rank < spec->world && spec->stripe_rows >= 1,
               "se_shard_local_rows: bad shard spec (world %d rank %d stripe_rows %lld)", spec->world, spec->rank,
               (long long)spec->stripe_rows);
    *n_rows = shard_local_rows(vocab, spec);
    return SE_OK;
}

// Fills the negative-sampler / scope fields of the kernel arguments from a shard spec (NULL = one unsharded table).
static int apply_shard_spec(const char *fn, se::SgnsArgs &a, const se_shard_spec *spec) {
    a.neg_vocab = (uint32_t)a.vocab; a.neg_shift = -1; a.neg_world = 1; a.neg_rank = 0; a.sys_scope = 0;
    if (!spec) return SE_OK;
    if (spec->world < 1 || spec->rank < 0 || spec->rank >= spec->world || spec->stripe_rows < 1) {
        se::set_error("%s: bad shard spec (world %d rank %d stripe_rows %lld)", fn, spec->world, spec->rank, (long long)spec->stripe_rows);
        return SE_ERR_INVALID_ARG;
    }
    a.sys_scope = spec->world > 1;
    if (spec->local_negatives && spec->world > 1) {
        const int64_t sr = spec->stripe_rows;
        if (sr & (sr - 1)) {
            se::set_error("%s: local negatives need a power-of-two stripe_rows (got %lld): choose emb and stripe_bytes so that "
                          "stripe_bytes / (4 * emb) is a power of two", fn, (long long)sr);
            return SE_ERR_UNSUPPORTED;
        }
        int shift = 0;
        while ((1ll << shift) < sr) ++shift;
        const int64_t local = shard_local_rows(a.vocab, spec);
        if (local < 1) { se::set_error("%s: rank %d owns no rows", fn, spec->rank); return SE_ERR_INVALID_ARG; }
        a.neg_vocab = (uint32_t)local; a.neg_shift = shift; a.neg_world = spec->world; a.neg_rank = spec->rank;
    }
    return SE_OK;
}

extern "C" int se_sgns_update_walks_sharded(float *w_in, float *w_out, int64_t vocab, int emb, const int32_t *tokens,
                                            int64_t n_seq, int seq_len, int radius, int n_neg, int row_offset,
                                            const float *alias_prob, const int32_t *alias_idx, float lr, uint64_t seed,
                                            int64_t centre_id_base, int flags, const se_shard_spec *spec, double *stats,
                                            void *stream) {
    int rc = se::common_checks("se_sgns_update_walks", w_in, w_out, vocab, emb, n_neg);
    if (rc != SE_OK) return rc;
    SE_REQUIRE(n_seq >= 0 && (tokens || n_seq == 0), "se_sgns_update_walks: null tokens");
    SE_REQUIRE(radius >= 1, "se_sgns_update_walks: context radius must be >= 1");
    // W2VCollateFunctional asserts text_length >= 2r+1 (torch_dataset.py:298)
    SE_REQUIRE(seq_len >= 2 * radius + 1, "Text is too short! [text_length=%d] < [min_text_length=%d]", seq_len, 2 * radius + 1);
    SE_REQUIRE((alias_prob == nullptr) == (alias_idx == nullptr), "se_sgns_update_walks: pass both alias arrays or neither");
    SE_REQUIRE((flags & ~(SE_SGNS_SCATTER_STORE | SE_SGNS_GENERIC_KERNEL | SE_SGNS_NO_WINDOW | SE_SGNS_WHOLE_SEQUENCES | SE_SGNS_WINDOW_REFRESH |
                          SE_SGNS_BATCHED_POSITIVES)) == 0, "se_sgns_update_walks: unknown flags %d", flags);
    se::SgnsArgs a{};
    a.w_in = w_in; a.w_out = w_out; a.tokens = tokens; a.alias_prob = alias_prob; a.alias_idx = alias_idx;
    a.stats = stats; a.vocab = vocab; a.emb = emb; a.n_ctx = 2 * radius; a.n_neg = n_neg;
    a.seq_len = seq_len; a.radius = radius; a.n_cen = seq_len - 2 * radius; a.row_offset = row_offset;
    a.n_units = n_seq * a.n_cen;
    a.lr = lr; a.seed = seed; a.id_base = centre_id_base; a.scatter_store = (flags & SE_SGNS_SCATTER_STORE) != 0;
    a.force_generic = (flags & SE_SGNS_GENERIC_KERNEL) != 0;
    a.no_window = (flags & SE_SGNS_NO_WINDOW) != 0;
    a.whole_seq = (flags & SE_SGNS_WHOLE_SEQUENCES) != 0;
    a.win_refresh = (flags & SE_SGNS_WINDOW_REFRESH) != 0;
    a.batch_pos = (flags & SE_SGNS_BATCHED_POSITIVES) != 0;
    a.n_seq = n_seq;
    rc = apply_shard_spec("se_sgns_update_walks_sharded", a, spec);
    if (rc != SE_OK) return rc;
    SE_REQUIRE(!(a.sys_scope && a.scatter_store), "se_sgns_update_walks_sharded: plain-store scatter is not supported on "
               "sharded tables (updates of other GPUs would be lost); use SE_SGNS_SCATTER_RED");
    return se::launch<se::MODE_WALK>(a, (cudaStream_t)stream);
}

extern "C" int se_sgns_update_negatives_owned(float *w_in, float *w_out, int64_t vocab, int emb, const int32_t *tokens,
                                              int64_t n_seq, int seq_len, int radius, int n_neg, int row_offset,
                                              const float *alias_prob, const int32_t *alias_idx, float lr, uint64_t seed,
                                              int64_t centre_id_base, const se_shard_spec *spec, double *stats, void *stream) {
    int rc = se::common_checks("se_sgns_update_negatives_owned", w_in, w_out, vocab, emb, n_neg);
    if (rc != SE_OK) return rc;
    SE_REQUIRE(spec, "se_sgns_update_negatives_owned: a shard spec is required");
    SE_REQUIRE(spec->world >= 1 && spec->rank >= 0 && spec->rank < spec->world && spec->stripe_rows >= 1,
               "se_sgns_update_negatives_owned: bad shard spec (world %d rank %d stripe_rows %lld)", spec->world, spec->rank,
               (long long)spec->stripe_rows);
    SE_REQUIRE(n_seq >= 0 && (tokens || n_seq == 0), "se_sgns_update_negatives_owned: null tokens");
    SE_REQUIRE(radius >= 1 && seq_len >= 2 * radius + 1, "Text is too short! [text_length=%d] < [min_text_length=%d]", seq_len, 2 * radius + 1);
    SE_REQUIRE((alias_prob == nullptr) == (alias_idx == nullptr), "se_sgns_update_negatives_owned: pass both alias arrays or neither");
    if (n_neg == 0 || n_seq == 0) return SE_OK;
    const int64_t sr = spec->stripe_rows;
    if ((sr & (sr - 1)) || emb % 4 != 0 || emb <= 32 || emb > 128 || n_neg > 7 || ((2 * radius + 3) / 4) * n_neg > 32 ||
        ((uintptr_t)w_in % 16) || ((uintptr_t)w_out % 16)) {
        se::set_error("se_sgns_update_negatives_owned: needs 32 < emb <= 128 (multiple of 4), n_neg <= 7, ceil(2r/4)*n_neg <= 32, "
                      "16-byte aligned tables and a power-of-two stripe_rows (emb %d, n_neg %d, radius %d, stripe_rows %lld)",
                      emb, n_neg, radius, (long long)sr);
        return SE_ERR_UNSUPPORTED;
    }
    se::SgnsArgs a{};
    a.w_in = w_in; a.w_out = w_out; a.tokens = tokens; a.alias_prob = alias_prob; a.alias_idx = alias_idx;
    a.stats = stats; a.vocab = vocab; a.emb = emb; a.n_ctx = 2 * radius; a.n_neg = n_neg;
    a.seq_len = seq_len; a.radius = radius; a.n_cen = seq_len - 2 * radius; a.row_offset = row_offset;
    a.n_units = n_seq * a.n_cen;
    a.lr = lr; a.seed = seed; a.id_base = centre_id_base;
    a.neg_vocab = (uint32_t)vocab; a.neg_shift = -1; a.neg_world = spec->world; a.neg_rank = spec->rank;
    a.sys_scope = spec->world > 1;
    a.own_shift = 0;
    while ((1ll << a.own_shift) < sr) ++a.own_shift;
    return emb == 128 ? se::launch_negown<true>(a, (cudaStream_t)stream) : se::launch_negown<false>(a, (cudaStream_t)stream);
}

extern "C" int se_sgns_update_walks(float *w_in, float *w_out, int64_t vocab, int emb, const int32_t *tokens,
                                    int64_t n_seq, int seq_len, int radius, int n_neg, int row_offset,
                                    const float *alias_prob, const int32_t *alias_idx, float lr, uint64_t seed,
                                    int64_t centre_id_base, int flags, double *stats, void *stream) {
    return se_sgns_update_walks_sharded(w_in, w_out, vocab, emb, tokens, n_seq, seq_len, radius, n_neg, row_offset,
                                        alias_prob, alias_idx, lr, seed, centre_id_base, flags, nullptr, stats, stream);
}

extern "C" int se_skipgram_scores_backward(const float *w_in, const float *w_out, int64_t vocab, int emb,
                                           const int64_t *inputs, const int64_t *outputs, int64_t batch, int m,
                                           const float *grad_scores, float *grad_in, float *grad_out, void *stream) {
    int rc = se::common_checks("se_skipgram_scores_backward", w_in, w_out, vocab, emb, 0);
    if (rc != SE_OK) return rc;
    SE_REQUIRE(batch >= 0 && m >= 1, "se_skipgram_scores_backward: bad shape");
    if (batch == 0) return SE_OK;
    SE_REQUIRE(inputs && outputs && grad_scores && grad_in && grad_out, "se_skipgram_scores_backward: null pointer");
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (batch + 7) / 8; if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::scores_backward_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w_in, w_out, emb, inputs, outputs, batch, m,
                                                                             grad_scores, grad_in, grad_out);
    return se::check_cuda(cudaGetLastError(), "scores_backward_kernel launch");
}

extern "C" int se_ns_loss(const float *pos_logits, const float *neg_logits, int64_t batch, int n_ctx, int n_neg,
                          double *stats, float *grad_pos, float *grad_neg, void *stream) {
    SE_REQUIRE(batch >= 0 && n_ctx >= 1 && n_neg >= 0, "se_ns_loss: bad shape");
    const int64_t pairs = batch * n_ctx;
    if (pairs == 0) return SE_OK;
    SE_REQUIRE(pos_logits && (neg_logits || n_neg == 0) && stats, "se_ns_loss: null pointer");
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (pairs + 255) / 256; if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::ns_loss_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pos_logits, neg_logits, pairs, n_neg,
                                                                     1.0f / (float)pairs, stats, grad_pos, grad_neg);
    return se::check_cuda(cudaGetLastError(), "ns_loss_kernel launch");
}

// Window-resident SGNS kernel (MODE_WALK: tokens[n_seq x L] -> windows, negatives, in-place update), sm_100a.
//
// A GROUP of G lanes (G = 32 / 16 / 8 for rows of up to 128 / 64 / 32 floats: 1 / 2 / 4 centres per warp, one float4 per lane)
// walks its span of centres with the W_out rows of the 2r+1 tokens around the current centre RESIDENT in shared memory (a ring
// of 2r+2 physical slots per group: current value + pending update).  A token's context row is fetched ONCE when it enters the
// window (cp.async, issued one centre ahead), serves as the positive row of up to 2r centres from shared memory, and its
// accumulated update leaves with ONE red.global.add.v4.f32 when the token slides out -- instead of 2r gathers and 2r scatters
// through L2 (window rule: word2vec/dataloader/torch_dataset.py:300-309).  Negatives, the centre row, the transposed reduction
// and the loss arithmetic (word2vec/loss.py:15-16) are those of sgns_ctx_kernel.
//
// Repeated tokens.  Random walks revisit nodes all the time (A-B-A at p = 0.5; a whole walk on a 3-node path).  Window POSITIONS
// (logical ring slots) are therefore mapped to PHYSICAL slots through a small table: a token that enters the window while its row
// is already resident ALIASES the resident slot, so inside one group every pair sees the row's latest value -- the pair-by-pair
// semantics of the reference's sequential updates and of the per-pair kernel -- and no duplicate is fetched.  When the older
// position slides out, the pending update is scattered (so updates never wait longer than one window length) and the slot lives on
// for its remaining aliases.
//
// Mid-life refresh (SE_SGNS_WINDOW_REFRESH).  On token streams whose frequent tokens sit in thousands of windows at once (a Zipf
// text corpus) a resident copy misses the other groups' updates for a whole window length; the flag scatters and re-fetches a
// token's row when the token is the centre (it is not a context then, so nothing in use is touched): S4 mean loss 2.78 -> 2.44
// (per-pair kernel: 2.45) for 6 % throughput.
//
// Tried and dropped (profiles/r02_s4_matrix.md): combining the most frequent rows per CTA in shared memory (fixed-point integer
// atomics; fp32 shared atomics compile to CAS loops).  It removes the hottest rows' red.global traffic (the busiest L2 slice's
// atomic unit is 68 % active against 23 % on average with unigram^0.75 negatives) but gains 3 % at best and delays the hot rows'
// updates, which costs optimisation progress -- the alias-negative slowdown is not an atomic-throughput limit.
#pragma once
#include "sgns_common.cuh"

namespace se {
namespace {

template <int G, int T, bool EXACT, int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
sgns_win_kernel(const SgnsArgs a) {
    constexpr int K = T - 1;
    constexpr int P = (T <= 1) ? 1 : (T <= 2) ? 2 : (T <= 4) ? 4 : 8;              // dots padded to a power of two
    constexpr int LOG_G = (G == 32) ? 5 : (G == 16) ? 4 : 3;
    constexpr int LOG_P = (P == 8) ? 3 : (P == 4) ? 2 : (P == 2) ? 1 : 0;
    constexpr int SHIFT = LOG_G - LOG_P;                                           // lanes per owner sub-group = 1 << SHIFT
    constexpr int GPB = THREADS / G;                                               // groups per block
    static_assert(P <= G && T <= G, "one lane per target row");
    extern __shared__ float4 win_smem[];
    const int lg = threadIdx.x & (G - 1);
    const int grp = threadIdx.x / G;
    const unsigned gmask = group_mask<G>();
    const int64_t gid = (int64_t)blockIdx.x * GPB + grp;
    const int64_t n_groups = (int64_t)gridDim.x * GPB;
    const int E = EXACT ? 4 * G : a.emb;
    const int eoff = lg * 4;
    const bool ok = EXACT || eoff < E;
    const int N = a.n_ctx, NG = (a.n_ctx + 3) >> 2, r = a.radius;
    const int RING = 2 * r + 2;
    float4 *cur = win_smem + (size_t)grp * 2 * RING * G + lg;                      // physical slot s: cur[s * G], del[s * G]
    float4 *del = cur + RING * G;
    int *ids_s = reinterpret_cast<int *>(win_smem + (size_t)GPB * 2 * RING * G) + grp * 2 * RING;   // row id of logical slot l
    int *phys_s = ids_s + RING;                                                                     // its physical slot
    const int owner_t = lg >> SHIFT;
    const bool owner_rep = (lg & ((1 << SHIFT) - 1)) == 0;

    float loss_pos = 0.f, loss_neg = 0.f;
    unsigned cnt_recall = 0, cnt_fp = 0, cnt_pairs = 0;
    unsigned free_mask = 0;

    auto load_row = [&](int rid, float (&v)[4]) {
        v[0] = v[1] = v[2] = v[3] = 0.f;
        if (ok) load_vec<4>(a.w_out + (int64_t)rid * E + eoff, v);
    };
    auto push_row = [&](int rid, const float (&d)[4]) {
        if (ok) red_vec<4>(a.w_out + (int64_t)rid * E + eoff, d, a.sys_scope);
    };
    auto scatter_slot = [&](int ph, int rid) {                                     // pending update of a resident slot -> global, then cleared
        if (!ok) return;
        const float4 d4 = del[ph * G];
        if (d4.x != 0.f || d4.y != 0.f || d4.z != 0.f || d4.w != 0.f) {
            const float d[4] = {d4.x, d4.y, d4.z, d4.w};
            red_vec<4>(a.w_out + (int64_t)rid * E + eoff, d, a.sys_scope);
            del[ph * G] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    // token with row `rid` enters logical slot l_new; the n_valid logical slots from first_l on are searched for the same row
    auto enter = [&](int l_new, int rid, int first_l, int n_valid) {
        int ph = -1;
        int l = first_l;
        for (int j = 0; j < n_valid; ++j) {
            if (ids_s[l] == rid) ph = phys_s[l];
            if (++l == RING) l = 0;
        }
        if (ph < 0) {
            ph = __ffs(free_mask) - 1;
            free_mask &= ~(1u << ph);
            if (ok) {
                cp_async16(cur + ph * G, a.w_out + (int64_t)rid * E + eoff);
                del[ph * G] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncwarp(gmask);
        if (lg == 0) { ids_s[l_new] = rid; phys_s[l_new] = ph; }
        __syncwarp(gmask);
    };

    // ids of the negatives this lane owns (lane t = negative t - 1) in contexts 4g .. 4g + 3 of centre uu
    auto draw_group = [&](int64_t uu, int g, int (&out)[4]) {
        out[0] = out[1] = out[2] = out[3] = 0;
        if (K > 0) {
            const uint64_t cid = (uint64_t)(a.id_base + uu);
            const int k = lg >= 1 ? lg - 1 : 0;
            const uint4 wb = neg_words(a.seed, cid, g * 4, k, STREAM_NEG);
            uint4 wc = make_uint4(0, 0, 0, 0);
            if (a.alias_prob) wc = neg_words(a.seed, cid, g * 4, k, STREAM_NEG_COIN);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (g * 4 + j < N && lg >= 1 && lg < T) out[j] = neg_row(a, pick_word(wb, j), pick_word(wc, j));
        }
    };

    int64_t span = (a.n_units + n_groups - 1) / n_groups;
    if (a.whole_seq) span = ((span + a.n_cen - 1) / a.n_cen) * a.n_cen;            // no sequence is split between two groups
    int64_t u = gid * span;
    const int64_t u_end = min(a.n_units, u + span);

    while (u < u_end) {
        // ---- one segment: consecutive centres of ONE sequence ----------------------------------------------------
        const int64_t s = u / a.n_cen;
        const int p0 = r + (int)(u - s * a.n_cen);
        const int m = (int)min(u_end - u, (int64_t)(a.n_cen - (p0 - r)));          // centres p0 .. p0 + m - 1
        const int32_t *seq = a.tokens + s * a.seq_len;
        free_mask = (RING >= 32) ? 0xffffffffu : ((1u << RING) - 1u);
        __syncwarp(gmask);
        // window of the first centre: positions p0 - r .. p0 + r -> logical slots 0 .. 2r
        for (int j = 0; j <= 2 * r; ++j) enter(j, __ldg(seq + p0 - r + j) + a.row_offset, 0, j);
        cp_async_wait_all();
        int head = 0;                                                              // logical slot of position p - r
        // T == 1 (positives only): a centre is over in a few hundred cycles, so its row is fetched while the previous centre is computed
        float cen_pre[4] = {0.f, 0.f, 0.f, 0.f};
        if (T == 1 && ok) load_vec<4>(a.w_in + ((int64_t)__ldg(seq + p0) + a.row_offset) * E + eoff, cen_pre);

        for (int p = p0; p < p0 + m; ++p, ++u) {
            // the row entering the window for the next centre goes to the free logical slot while this centre is processed
            int l_in = head - 1; if (l_in < 0) l_in += RING;
            const bool slide = p + 1 < p0 + m;
            if (slide) enter(l_in, __ldg(seq + p + r + 1) + a.row_offset, head, 2 * r + 1);
            const int64_t crow = (int64_t)__ldg(seq + p) + a.row_offset;
            float cen[4] = {0.f, 0.f, 0.f, 0.f}, acc[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (T == 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) cen[e] = cen_pre[e];
                // (issued after the reductions of all earlier centres, which a later load of the same thread observes; only THIS centre's
                //  update is missing from it and is added below when the next centre is the same row)
                if (slide && ok) load_vec<4>(a.w_in + ((int64_t)__ldg(seq + p + 1) + a.row_offset) * E + eoff, cen_pre);
            } else {
                if (ok) load_vec<4>(a.w_in + crow * E + eoff, cen);
            }
            // mid-life refresh: the token that is the centre right now is not a context of this centre, so its resident W_out row can
            // be scattered and re-fetched asynchronously without touching anything in use -- halves how stale a resident copy gets
            // relative to the other groups (matters for frequent tokens, which sit in thousands of windows at once)
            if (a.win_refresh && slide) {
                int lc = head + r; if (lc >= RING) lc -= RING;
                const int phc = phys_s[lc];
                bool aliased = false;
                int l = head;
                for (int j = 0; j <= 2 * r; ++j) {
                    aliased |= (l != lc) && phys_s[l] == phc;
                    if (++l == RING) l = 0;
                }
                if (!aliased) {
                    scatter_slot(phc, ids_s[lc]);
                    if (ok) cp_async16(cur + phc * G, a.w_out + (int64_t)ids_s[lc] * E + eoff);
                }
            }

            if (T == 1 && G == 32 && a.batch_pos) {
                // SE_SGNS_BATCHED_POSITIVES, positives only (the owner-computes mode runs the negatives elsewhere): all 2r context rows are resident, so the 2r dots
                // of a centre are reduced by ONE transposed butterfly (15 + 1 shuffles instead of 2r x 5 dependent ones) and the sigmoids
                // run side by side, two lanes per context.  The 2r pairs of a centre are scored against the same snapshot of the window;
                // their updates are then applied one after the other (a row that sits in the window twice receives both).
                float dot[16];
#pragma unroll
                for (int n = 0; n < 16; ++n) {
                    float d = 0.f;
                    if (n < N && ok) {
                        int l = head + ((n < r) ? n : n + 1);
                        if (l >= RING) l -= RING;
                        const float4 c4 = cur[phys_s[l] * G];
                        d = fmaf(c4.x, cen[0], fmaf(c4.y, cen[1], fmaf(c4.z, cen[2], c4.w * cen[3])));
                    }
                    dot[n] = d;
                }
                const float x = transposed_reduce<16, 32>(dot, lg, gmask);               // context n in lanes 2n, 2n + 1
                float step_mine = 0.f;
                if ((lg >> 1) < N) {
                    const float ex = __expf(-x);
                    const float sig = __fdividef(1.0f, 1.0f + ex);
                    step_mine = (sig > CLAMP_MIN) ? a.lr * ex * sig : 0.f;               // -lr * dL/ds, loss = -log clamp(sigmoid(x), 1e-6)
                    if ((lg & 1) == 0) { loss_pos -= __logf(fmaxf(sig, CLAMP_MIN)); cnt_recall += x >= 0.f; cnt_pairs += 1; }
                }
                for (int n = 0; n < N; ++n) {
                    const float step = __shfl_sync(gmask, step_mine, 2 * n, G);
                    int l = head + ((n < r) ? n : n + 1);
                    if (l >= RING) l -= RING;
                    const int ph = phys_s[l];
                    if (ok) {
                        const float4 c4 = cur[ph * G];
                        const float upd[4] = {step * cen[0], step * cen[1], step * cen[2], step * cen[3]};
                        acc[0] = fmaf(step, c4.x, acc[0]); acc[1] = fmaf(step, c4.y, acc[1]);
                        acc[2] = fmaf(step, c4.z, acc[2]); acc[3] = fmaf(step, c4.w, acc[3]);
                        cur[ph * G] = make_float4(c4.x + upd[0], c4.y + upd[1], c4.z + upd[2], c4.w + upd[3]);
                        float4 d4 = del[ph * G];
                        d4.x += upd[0]; d4.y += upd[1]; d4.z += upd[2]; d4.w += upd[3];
                        del[ph * G] = d4;
                    }
                }
            } else {
                for (int g = 0; g < NG; ++g) {
                    // ids of the negatives this lane owns (lane t = negative t - 1) in contexts 4g .. 4g + 3.  (Measured and dropped,
                    // profiles/r02_sgns_tuning.md: resolving them one group ahead -- no gain with alias tables, -2 % on S3 through register
                    // pressure; keeping the negative rows of TWO contexts in flight per warp at 12 warps / 165 registers -- 191 ms vs 182 ms.)
                    int ids[4];
                    draw_group(u, g, ids);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = g * 4 + j;
                        if (n < N) {
                            int l = head + ((n < r) ? n : n + 1);                      // window offset of context n (centre skipped)
                            if (l >= RING) l -= RING;
                            const int ph = phys_s[l];
                            int tid[T];
                            float row[T][4];
                            float dot[P];
#pragma unroll
                            for (int t = 1; t < T; ++t) tid[t] = __shfl_sync(gmask, ids[j], t, G);
                            row[0][0] = row[0][1] = row[0][2] = row[0][3] = 0.f;
                            if (ok) { const float4 c4 = cur[ph * G]; row[0][0] = c4.x; row[0][1] = c4.y; row[0][2] = c4.z; row[0][3] = c4.w; }
#pragma unroll
                            for (int t = 1; t < T; ++t) load_row(tid[t], row[t]);
#pragma unroll
                            for (int t = 0; t < P; ++t) {
                                float d = 0.f;
                                if (t < T) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) d = fmaf(row[t][e], cen[e], d);
                                }
                                dot[t] = d;
                            }
                            const float sc = transposed_reduce<P, G>(dot, lg, gmask);
                            float step_mine = 0.f;
                            if (owner_t < T) {
                                const bool positive = owner_t == 0;
                                const float x = positive ? sc : -sc;                      // loss = -log clamp(sigmoid(x), 1e-6)
                                const float ex = __expf(-x);
                                const float sig = __fdividef(1.0f, 1.0f + ex);
                                const bool live = sig > CLAMP_MIN;
                                const float gmag = live ? ex * sig : 0.f;                 // |dL/ds| = sigmoid(-x)
                                step_mine = positive ? a.lr * gmag : -a.lr * gmag;        // -lr * dL/ds
                                if (owner_rep) {
                                    const float lo = -__logf(fmaxf(sig, CLAMP_MIN));
                                    if (positive) { loss_pos += lo; cnt_recall += x >= 0.f; cnt_pairs += 1; }
                                    else { loss_neg += lo; cnt_fp += x <= 0.f; }
                                }
                            }
                            {   // positive row: update the resident copy and its pending update
                                const float step = __shfl_sync(gmask, step_mine, 0, G);
                                const float upd[4] = {step * cen[0], step * cen[1], step * cen[2], step * cen[3]};
#pragma unroll
                                for (int e = 0; e < 4; ++e) acc[e] = fmaf(step, row[0][e], acc[e]);
                                if (ok) {
                                    cur[ph * G] = make_float4(row[0][0] + upd[0], row[0][1] + upd[1], row[0][2] + upd[2], row[0][3] + upd[3]);
                                    float4 d4 = del[ph * G];
                                    d4.x += upd[0]; d4.y += upd[1]; d4.z += upd[2]; d4.w += upd[3];
                                    del[ph * G] = d4;
                                }
                            }
#pragma unroll
                            for (int t = 1; t < T; ++t) {
                                const float step = __shfl_sync(gmask, step_mine, t << SHIFT, G);
                                float d[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) { acc[e] = fmaf(step, row[t][e], acc[e]); d[e] = step * cen[e]; }
                                push_row(tid[t], d);
                            }
                        }
                    }
                }
            }
            if (ok) red_vec<4>(a.w_in + crow * E + eoff, acc, a.sys_scope);
            if constexpr (T == 1) {
                if (slide && (int64_t)__ldg(seq + p + 1) + a.row_offset == crow) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) cen_pre[e] += acc[e];
                }
            }
            // slide: the oldest position leaves the window.  Its pending update is scattered now (even if other positions alias
            // the slot: updates never wait longer than one window length); the slot is freed when no alias is left.
            if (slide) {
                const int ph = phys_s[head];
                scatter_slot(ph, ids_s[head]);
                bool aliased = false;
                int l = head;
                for (int j = 0; j <= 2 * r; ++j) {                                 // the window after the slide: head + 1 .. head + 2r + 1
                    if (++l == RING) l = 0;
                    aliased |= phys_s[l] == ph;
                }
                if (!aliased) free_mask |= 1u << ph;
                cp_async_wait_all();
                if (++head == RING) head = 0;
            }
        }
        // segment end: scatter what is still pending in the 2r + 1 resident positions (each physical slot once)
        {
            unsigned done = 0;
            int l = head;
            for (int j = 0; j <= 2 * r; ++j) {
                const int ph = phys_s[l];
                if (!((done >> ph) & 1u)) {
                    done |= 1u << ph;
                    scatter_slot(ph, ids_s[l]);
                }
                if (++l == RING) l = 0;
            }
        }
    }

    flush_stats(a.stats, loss_pos != 0.f || loss_neg != 0.f || cnt_pairs != 0 || cnt_fp != 0 || cnt_recall != 0, loss_pos, loss_neg, cnt_recall,
                cnt_fp, cnt_pairs, (double)cnt_pairs * (double)K);
}

inline size_t win_smem_bytes(int gpb, int g, int radius) {
    const int ring = 2 * radius + 2;
    return (size_t)gpb * 2 * ring * g * sizeof(float4) + (size_t)gpb * 2 * ring * sizeof(int);
}

template <int G, int T, bool EXACT>
int launch_win_one(const SgnsArgs &a_in, cudaStream_t stream) {
    constexpr int THREADS = SGNS_THREADS;
    constexpr int GPB = THREADS / G;
    auto kern = sgns_win_kernel<G, T, EXACT, THREADS>;
    SgnsArgs a = a_in;
    const size_t smem = win_smem_bytes(GPB, G, a.radius);
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute") != SE_OK) return SE_ERR_CUDA;
    const int blocks = persistent_blocks(kern, smem, a.n_units, GPB, true, THREADS);
    if (blocks < 0) return SE_ERR_UNSUPPORTED;               // the ring does not fit: the caller falls back to the per-context kernel
    if (blocks == 0) return SE_ERR_CUDA;
    if (a.n_seq >= (int64_t)blocks * GPB) a.whole_seq = 1;   // enough sequences for every group: never split one
    kern<<<blocks, THREADS, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), "sgns_win_kernel launch");
}

template <int G, bool EXACT>
int launch_win_t(const SgnsArgs &a, cudaStream_t stream) {
    switch (1 + a.n_neg) {
        case 1: return launch_win_one<G, 1, EXACT>(a, stream);
        case 2: return launch_win_one<G, 2, EXACT>(a, stream);
        case 3: return launch_win_one<G, 3, EXACT>(a, stream);
        case 4: return launch_win_one<G, 4, EXACT>(a, stream);
        case 5: return launch_win_one<G, 5, EXACT>(a, stream);
        case 6: return launch_win_one<G, 6, EXACT>(a, stream);
        case 7: return launch_win_one<G, 7, EXACT>(a, stream);
        case 8: return launch_win_one<G, 8, EXACT>(a, stream);
        default: return SE_ERR_UNSUPPORTED;
    }
}

}  // namespace
}  // namespace se

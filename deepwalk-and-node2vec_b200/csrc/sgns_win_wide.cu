// Window-resident SGNS kernel for WIDE rows (128 < emb <= 256), sm_100a: the scheme of sgns_win.cuh with R = 2 float4 per lane.
//
// One warp owns one centre; a row of up to 128 R floats is R coalesced 512-byte chunks (lane l holds floats (32 i + l) 4 .. + 3 of
// chunk i).  The W_out rows of the 2r+1 tokens around the centre stay RESIDENT in shared memory (ring of 2r+2 physical slots per
// warp, current value + pending update, 2 R KB per slot): a context row is fetched once when its token enters the window
// (cp.async.cg, issued one centre ahead), serves up to 2r centres from shared memory and leaves with one vector reduction per
// chunk when the token slides out (window rule: word2vec/dataloader/torch_dataset.py:300-309).  Repeated tokens alias one
// physical slot (pair-by-pair semantics inside a warp, as the reference's sequential updates), negatives are drawn in-kernel from
// the same Philox keys as every other SGNS kernel (uniform = word2vec/utils/sampling.py:21, or the alias table), loss
// arithmetic as word2vec/loss.py:15-16.  Before this kernel rows wider than 128 floats went to sgns_fast_kernel, which gathers
// and scatters every context row 2r times through L2.
//
// Geometry: blocks of 4 warps, 24 KB of ring per warp at r = 5: two blocks of 96 KB per SM (one block up to r = 8).  If the ring does not
//           fit the launcher reports SE_ERR_UNSUPPORTED and the caller falls back to sgns_fast_kernel.
// Measured (profiles/r02e_summary.md): S4 token stream at E = 256: 0.478 G pairs/s against 0.314 G for sgns_fast_kernel; on HBM-resident
// tables (S3 shape) 0.476 against 0.494 G -- 8 warps per SM are few for DRAM latency.  R = 4 (emb <= 512, 48 KB of ring per warp, 4 warps
// per SM) was built and measured too: 0.174 G against 0.224 G for sgns_fast_kernel on the S4 stream -- slower, so rows wider than 256
// floats stay on sgns_fast_kernel and only R = 2 is instantiated.
// The file is separate from sgns_win.cuh on purpose: that header is the profiled source set of the S3 bench kernel
// (profiles/sgns_traffic.json is stamped with its hash).
#include "sgns_common.cuh"

namespace se {
namespace {

template <int R, int T, int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
sgns_winw_kernel(const SgnsArgs a) {
    constexpr int G = 32;
    constexpr int K = T - 1;
    constexpr int P = (T <= 1) ? 1 : (T <= 2) ? 2 : (T <= 4) ? 4 : 8;              // dots padded to a power of two
    constexpr int LOG_P = (P == 8) ? 3 : (P == 4) ? 2 : (P == 2) ? 1 : 0;
    constexpr int SHIFT = 5 - LOG_P;                                               // lanes per owner sub-group = 1 << SHIFT
    constexpr int GPB = THREADS / G;                                               // warps per block
    constexpr int SLOT = R * G;                                                    // float4 per physical slot
    static_assert(T >= 2 && T <= 8, "positive + 1..7 negatives");
    extern __shared__ float4 winw_smem[];
    const int lg = threadIdx.x & 31;
    const int grp = threadIdx.x >> 5;
    const int64_t gid = (int64_t)blockIdx.x * GPB + grp;
    const int64_t n_groups = (int64_t)gridDim.x * GPB;
    const int E = a.emb;
    int eoff[R];
    bool ok[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { eoff[i] = (i * G + lg) * 4; ok[i] = eoff[i] < E; }
    const int N = a.n_ctx, NG = (a.n_ctx + 3) >> 2, r = a.radius;
    const int RING = 2 * r + 2;
    float4 *cur = winw_smem + (size_t)grp * 2 * RING * SLOT + lg;                  // physical slot s, chunk i: cur[s * SLOT + i * G]
    float4 *del = cur + RING * SLOT;
    int *ids_s = reinterpret_cast<int *>(winw_smem + (size_t)GPB * 2 * RING * SLOT) + grp * 2 * RING;   // row id of logical slot l
    int *phys_s = ids_s + RING;                                                                         // its physical slot
    const int owner_t = lg >> SHIFT;
    const bool owner_rep = (lg & ((1 << SHIFT) - 1)) == 0;

    float loss_pos = 0.f, loss_neg = 0.f;
    unsigned cnt_recall = 0, cnt_fp = 0, cnt_pairs = 0;
    unsigned free_mask = 0;

    auto fetch_slot = [&](int ph, int rid) {                                       // asynchronous: row `rid` of W_out -> slot ph
        const float *src = a.w_out + (int64_t)rid * E;
#pragma unroll
        for (int i = 0; i < R; ++i)
            if (ok[i]) cp_async16(cur + ph * SLOT + i * G, src + eoff[i]);
    };
    auto scatter_slot = [&](int ph, int rid) {                                     // pending update of a resident slot -> global, then cleared
        float *dst = a.w_out + (int64_t)rid * E;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (!ok[i]) continue;
            const float4 d4 = del[ph * SLOT + i * G];
            if (d4.x != 0.f || d4.y != 0.f || d4.z != 0.f || d4.w != 0.f) {
                const float d[4] = {d4.x, d4.y, d4.z, d4.w};
                red_vec<4>(dst + eoff[i], d, a.sys_scope);
                del[ph * SLOT + i * G] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
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
            fetch_slot(ph, rid);
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (ok[i]) del[ph * SLOT + i * G] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        if (lg == 0) { ids_s[l_new] = rid; phys_s[l_new] = ph; }
        __syncwarp();
    };
    // ids of the negatives this lane owns (lane t = negative t - 1) in contexts 4g .. 4g + 3 of centre uu
    auto draw_group = [&](int64_t uu, int g, int (&out)[4]) {
        out[0] = out[1] = out[2] = out[3] = 0;
        const uint64_t cid = (uint64_t)(a.id_base + uu);
        const int k = lg >= 1 ? lg - 1 : 0;
        const uint4 wb = neg_words(a.seed, cid, g * 4, k, STREAM_NEG);
        uint4 wc = make_uint4(0, 0, 0, 0);
        if (a.alias_prob) wc = neg_words(a.seed, cid, g * 4, k, STREAM_NEG_COIN);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (g * 4 + j < N && lg >= 1 && lg < T) out[j] = neg_row(a, pick_word(wb, j), pick_word(wc, j));
    };

    int64_t span = (a.n_units + n_groups - 1) / n_groups;
    if (a.whole_seq) span = ((span + a.n_cen - 1) / a.n_cen) * a.n_cen;            // no sequence is split between two warps
    int64_t u = gid * span;
    const int64_t u_end = min(a.n_units, u + span);

    while (u < u_end) {
        // ---- one segment: consecutive centres of ONE sequence ----------------------------------------------------
        const int64_t s = u / a.n_cen;
        const int p0 = r + (int)(u - s * a.n_cen);
        const int m = (int)min(u_end - u, (int64_t)(a.n_cen - (p0 - r)));          // centres p0 .. p0 + m - 1
        const int32_t *seq = a.tokens + s * a.seq_len;
        free_mask = (1u << RING) - 1u;                                             // RING <= 18
        __syncwarp();
        // window of the first centre: positions p0 - r .. p0 + r -> logical slots 0 .. 2r
        for (int j = 0; j <= 2 * r; ++j) enter(j, __ldg(seq + p0 - r + j) + a.row_offset, 0, j);
        cp_async_wait_all();
        int head = 0;                                                              // logical slot of position p - r

        for (int p = p0; p < p0 + m; ++p, ++u) {
            // the row entering the window for the next centre goes to the free logical slot while this centre is processed
            int l_in = head - 1; if (l_in < 0) l_in += RING;
            const bool slide = p + 1 < p0 + m;
            if (slide) enter(l_in, __ldg(seq + p + r + 1) + a.row_offset, head, 2 * r + 1);
            const int64_t crow = (int64_t)__ldg(seq + p) + a.row_offset;
            float cen[R][4], acc[R][4];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                cen[i][0] = cen[i][1] = cen[i][2] = cen[i][3] = 0.f;
                acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                if (ok[i]) load_vec<4>(a.w_in + crow * E + eoff[i], cen[i]);
            }
            // mid-life refresh (SE_SGNS_WINDOW_REFRESH): the centre's own token is not a context of this centre, so its resident W_out row
            // can be scattered and re-fetched asynchronously without touching anything in use
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
                    fetch_slot(phc, ids_s[lc]);
                }
            }

            for (int g = 0; g < NG; ++g) {
                int ids[4];
                draw_group(u, g, ids);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = g * 4 + j;
                    if (n < N) {
                        int l = head + ((n < r) ? n : n + 1);                          // window offset of context n (centre skipped)
                        if (l >= RING) l -= RING;
                        const int ph = phys_s[l];
                        int tid[T];
                        float row[T][R][4];
                        float dot[P];
#pragma unroll
                        for (int t = 1; t < T; ++t) tid[t] = __shfl_sync(FULL, ids[j], t);
                        // negative rows first (global, long latency), the resident positive row from shared memory behind them
#pragma unroll
                        for (int t = 1; t < T; ++t) {
                            const float *src = a.w_out + (int64_t)tid[t] * E;
#pragma unroll
                            for (int i = 0; i < R; ++i) {
                                row[t][i][0] = row[t][i][1] = row[t][i][2] = row[t][i][3] = 0.f;
                                if (ok[i]) load_vec<4>(src + eoff[i], row[t][i]);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < R; ++i) {
                            row[0][i][0] = row[0][i][1] = row[0][i][2] = row[0][i][3] = 0.f;
                            if (ok[i]) {
                                const float4 c4 = cur[ph * SLOT + i * G];
                                row[0][i][0] = c4.x; row[0][i][1] = c4.y; row[0][i][2] = c4.z; row[0][i][3] = c4.w;
                            }
                        }
#pragma unroll
                        for (int t = 0; t < P; ++t) {
                            float d = 0.f;
                            if (t < T) {
#pragma unroll
                                for (int i = 0; i < R; ++i)
#pragma unroll
                                    for (int e = 0; e < 4; ++e) d = fmaf(row[t][i][e], cen[i][e], d);
                            }
                            dot[t] = d;
                        }
                        const float sc = transposed_reduce<P, G>(dot, lg, FULL);
                        float step_mine = 0.f;
                        if (owner_t < T) {
                            const bool positive = owner_t == 0;
                            const float x = positive ? sc : -sc;                          // loss = -log clamp(sigmoid(x), 1e-6)
                            const float ex = __expf(-x);
                            const float sig = __fdividef(1.0f, 1.0f + ex);
                            const bool live = sig > CLAMP_MIN;
                            const float gmag = live ? ex * sig : 0.f;                     // |dL/ds| = sigmoid(-x)
                            step_mine = positive ? a.lr * gmag : -a.lr * gmag;            // -lr * dL/ds
                            if (owner_rep) {
                                const float lo = -__logf(fmaxf(sig, CLAMP_MIN));
                                if (positive) { loss_pos += lo; cnt_recall += x >= 0.f; cnt_pairs += 1; }
                                else { loss_neg += lo; cnt_fp += x <= 0.f; }
                            }
                        }
                        {   // positive row: update the resident copy and its pending update
                            const float step = __shfl_sync(FULL, step_mine, 0);
#pragma unroll
                            for (int i = 0; i < R; ++i) {
                                const float upd[4] = {step * cen[i][0], step * cen[i][1], step * cen[i][2], step * cen[i][3]};
#pragma unroll
                                for (int e = 0; e < 4; ++e) acc[i][e] = fmaf(step, row[0][i][e], acc[i][e]);
                                if (ok[i]) {
                                    cur[ph * SLOT + i * G] = make_float4(row[0][i][0] + upd[0], row[0][i][1] + upd[1], row[0][i][2] + upd[2], row[0][i][3] + upd[3]);
                                    float4 d4 = del[ph * SLOT + i * G];
                                    d4.x += upd[0]; d4.y += upd[1]; d4.z += upd[2]; d4.w += upd[3];
                                    del[ph * SLOT + i * G] = d4;
                                }
                            }
                        }
#pragma unroll
                        for (int t = 1; t < T; ++t) {
                            const float step = __shfl_sync(FULL, step_mine, t << SHIFT);
                            float *dst = a.w_out + (int64_t)tid[t] * E;
#pragma unroll
                            for (int i = 0; i < R; ++i) {
                                float d[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) { acc[i][e] = fmaf(step, row[t][i][e], acc[i][e]); d[e] = step * cen[i][e]; }
                                if (ok[i]) red_vec<4>(dst + eoff[i], d, a.sys_scope);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (ok[i]) red_vec<4>(a.w_in + crow * E + eoff[i], acc[i], a.sys_scope);
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

template <int R, int T>
int launch_winw_one(const SgnsArgs &a_in, cudaStream_t stream) {
    constexpr int THREADS = 128;
    constexpr int GPB = THREADS / 32;
    auto kern = sgns_winw_kernel<R, T, THREADS>;
    SgnsArgs a = a_in;
    const int ring = 2 * a.radius + 2;
    const size_t smem = (size_t)GPB * 2 * ring * R * 32 * sizeof(float4) + (size_t)GPB * 2 * ring * sizeof(int);
    if (smem + 1024 > 227u * 1024u) return SE_ERR_UNSUPPORTED;   // the ring (+ the block's static shared memory) does not fit: the caller falls back to sgns_fast_kernel
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute") != SE_OK) return SE_ERR_CUDA;
    const int blocks = persistent_blocks(kern, smem, a.n_units, GPB, true, THREADS);
    if (blocks < 0) return SE_ERR_UNSUPPORTED;
    if (blocks == 0) return SE_ERR_CUDA;
    if (a.n_seq >= (int64_t)blocks * GPB) a.whole_seq = 1;   // enough sequences for every warp: never split one
    kern<<<blocks, THREADS, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), "sgns_winw_kernel launch");
}

template <int R>
int launch_winw_t(const SgnsArgs &a, cudaStream_t stream) {
    switch (1 + a.n_neg) {
        case 2: return launch_winw_one<R, 2>(a, stream);
        case 3: return launch_winw_one<R, 3>(a, stream);
        case 4: return launch_winw_one<R, 4>(a, stream);
        case 5: return launch_winw_one<R, 5>(a, stream);
        case 6: return launch_winw_one<R, 6>(a, stream);
        case 7: return launch_winw_one<R, 7>(a, stream);
        case 8: return launch_winw_one<R, 8>(a, stream);
        default: return SE_ERR_UNSUPPORTED;                   // K = 0 (positives only) and K > 7 stay on the per-pair kernels
    }
}

}  // namespace

// 128 < emb <= 256, emb % 4 == 0, 1 <= n_neg <= 7, radius <= 8; SE_ERR_UNSUPPORTED otherwise (caller tries the next kernel)
int launch_win_wide(const SgnsArgs &a, cudaStream_t stream) {
    if (a.emb <= 128 || a.emb > 256 || a.emb % 4 != 0 || a.radius > 8) return SE_ERR_UNSUPPORTED;
    return launch_winw_t<2>(a, stream);
}

}  // namespace se

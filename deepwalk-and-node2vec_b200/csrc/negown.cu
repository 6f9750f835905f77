// Owner-computes pairs, grouped by centre row (striped tables, the reference's GLOBAL negative draw: word2vec/utils/sampling.py:21).
//
// `sgns_negown_kernel` (sgns.cu) visits the centres of all GPUs in walk order: every GPU reads and updates the centre row once per
// centre OCCURRENCE (147 M per step on S3 at 8 GPUs against 10 M distinct rows), fills on average 6 of the 8 row slots of a pass, and
// waits for the centre row before it can start on the negatives -- per-centre fixed work paid G times, which is what holds the mode at
// 0.55 of linear scaling on 8 GPUs (profiles/r02_multi_gpu.md, section 3).  Here the centre occurrences of the all-gathered batch are first
// bucketed by table row (counting sort: count / scan / two-pass fill, kernels of this file), then one warp owns a contiguous span of the sorted
// list, 32 entries at a time:
//   * a run of entries with the same row keeps the centre row in registers: it is read ONCE per run (prefetched while the previous run is
//     computed) and the accumulated update goes back with ONE vector reduction at the end of the run -- over NVLink that is two rows per
//     distinct (row, 32-entry chunk) instead of two per occurrence;
//   * the owned output rows of consecutive occurrences of a run share one list, so passes run with all eight row slots filled;
//   * when a centre needs at most 16 drawing lanes (K = 5, r = 5: 15), two occurrences are drawn per Philox round;
//   * with `positives`, the 2r context tokens of every occurrence (read from the gathered tokens, staged in shared memory one chunk ahead) join the list when
//     this GPU owns their W_out row: then EVERY pair of the batch, positive or negative, is computed by the owner of its output row,
//     W_out never crosses NVLink, and the separate positive pass (window kernel, K = 0: four peer rows per centre with nothing to
//     hide their latency behind) disappears.
// The draws are the single-GPU kernel's: negative k of context n of centre c comes from Philox(seed; centre_id_base + c, n, k), so the set
// of (centre, output row) pairs is unchanged -- only their order is.  Within a run the register copy of the centre row is advanced after
// every batch of passes (at most one draw round + 7 rows), a finer grain than the single-GPU kernel's 2r (1 + K) rows per centre; across
// GPUs and warps the usual Hogwild rules apply (a chunk boundary ends a run, so a hub row is refreshed from memory every 32 occurrences).
#include <type_traits>

#include "sgns_common.cuh"
#include "scan.cuh"

namespace se {
namespace {

constexpr int NGO_P = 8, NGO_SHIFT = 2;      // rows per pass; lanes per row after the transposed reduction (32 / 8 = 4)
constexpr int NGO_LIST = 176;                // a remainder of < 8 ids + one draw round of at most 128 negatives + 32 contexts (+ 8: vector reads of a ragged pass)

// 4-byte asynchronous copies for the context tokens staged a chunk ahead (kept here: sgns_common.cuh is part of the window kernel's
// profiled source set, profiles/sgns_traffic.json)
__device__ __forceinline__ void cp_async4(int *smem_dst, const int32_t *gsrc) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ int centre_row(const SgnsArgs &a, uint32_t u) {
    const uint32_t sq = u / (uint32_t)a.n_cen;
    return __ldg(a.tokens + (int64_t)sq * a.seq_len + a.radius + (int)(u - sq * (uint32_t)a.n_cen)) + a.row_offset;
}

// ---- bucketing: occurrences per row, then (after the scan) every centre into its row's range ------------------------------------------
__global__ void __launch_bounds__(256)
cen_count_kernel(const SgnsArgs a, unsigned long long *__restrict__ cnt) {
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < a.n_units; u += (int64_t)gridDim.x * blockDim.x) {
        const int row = centre_row(a, (uint32_t)u);
        if (row >= 0 && row < a.vocab) atomicAdd(cnt + row, 1ull);
    }
}

// Bucket fill in two passes.  Writing every centre straight to its row's range is 147 M random 8-byte stores into a 1.2 GB array
// (12 ms per step on S3 at 8 GPUs).  Pass 1 scatters the centres, in token order, into at most 2048 COARSE buckets of 2^cshift
// consecutive rows (a coarse bucket's range is known from the row-level scan): a block takes a tile of 16 K centres, histograms it in
// shared memory, reserves ONE range per non-empty bucket with a single global atomic, and writes its entries there -- few write
// heads, so sectors fill up in L2, and 8x fewer global atomics than entries.  Pass 2 reads that array in order -- the threads in
// flight cover a few coarse buckets at a time -- and places every entry in its row's range, inside a window of a few MB.
constexpr int COARSE_BUCKETS = 2048, COARSE_TILE = 16384;

__global__ void __launch_bounds__(256)
cen_coarse_kernel(const SgnsArgs a, const int64_t *__restrict__ start, int cshift, unsigned long long *__restrict__ ccur, int2 *__restrict__ tmp) {
    __shared__ int hist[COARSE_BUCKETS];
    __shared__ int base[COARSE_BUCKETS];
    for (int64_t t0 = (int64_t)blockIdx.x * COARSE_TILE; t0 < a.n_units; t0 += (int64_t)gridDim.x * COARSE_TILE) {
        const int64_t t1 = min(a.n_units, t0 + COARSE_TILE);
        for (int b = threadIdx.x; b < COARSE_BUCKETS; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        for (int64_t u = t0 + threadIdx.x; u < t1; u += blockDim.x) {
            const int row = centre_row(a, (uint32_t)u);
            if (row >= 0 && row < a.vocab) atomicAdd(&hist[row >> cshift], 1);
        }
        __syncthreads();
        for (int b = threadIdx.x; b < COARSE_BUCKETS; b += blockDim.x) {
            const int n = hist[b];
            if (n > 0) base[b] = (int)(start[(int64_t)b << cshift] + (int64_t)atomicAdd(ccur + b, (unsigned long long)n));
            hist[b] = 0;                                       // becomes the block's cursor inside its reserved range
        }
        __syncthreads();
        for (int64_t u = t0 + threadIdx.x; u < t1; u += blockDim.x) {
            const int row = centre_row(a, (uint32_t)u);
            if (row >= 0 && row < a.vocab) {
                const int b = row >> cshift;
                tmp[base[b] + atomicAdd(&hist[b], 1)] = make_int2((int)u, row);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
cen_fine_kernel(const int2 *__restrict__ tmp, const int64_t *__restrict__ n_entries_p, const int64_t *__restrict__ start,
                unsigned long long *__restrict__ cnt, int2 *__restrict__ entries) {
    const int64_t n = *n_entries_p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int2 t = tmp[i];
        const int64_t pos = start[t.y] + (int64_t)atomicAdd(cnt + t.y, ~0ull) - 1;            // the range fills from its end; order is immaterial
        entries[pos] = t;
    }
}

// ---- the update ---------------------------------------------------------------------------------------------------------------------
template <bool EXACT, bool POS>
__global__ void __launch_bounds__(SGNS_THREADS, 2)
sgns_owned_pairs_kernel(const SgnsArgs a, const int2 *__restrict__ entries, const int64_t *__restrict__ n_entries_p) {
    constexpr int P = NGO_P, SHIFT = NGO_SHIFT;
    __shared__ __align__(16) int own_list[SGNS_THREADS / 32][NGO_LIST];        // NGO_LIST * 4 bytes is a multiple of 16
    __shared__ int ctx_s[POS ? SGNS_THREADS / 32 : 1][2][POS ? 512 : 1];      // context tokens of this chunk / of the next one
    const int lane = threadIdx.x & 31;
    int *list = own_list[threadIdx.x >> 5];
    const int64_t gid = (int64_t)blockIdx.x * (SGNS_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (SGNS_THREADS / 32);
    const int64_t n_entries = *n_entries_p;
    const int64_t n_chunks = (n_entries + 31) >> 5;
    const int64_t span = (n_chunks + n_warps - 1) / n_warps;
    const int64_t c_begin = gid * span, c_end = min(n_chunks, c_begin + span);

    const int E = EXACT ? 128 : a.emb;
    const int eoff = lane * 4;
    const bool ok = EXACT || eoff < E;
    const int N = a.n_ctx, K = a.n_neg, NG = (a.n_ctx + 3) >> 2, D = NG * K;
    const bool two = D <= 16;                                  // two occurrences per draw round
    const int per = two ? 2 : 1;
    const int dl = two ? (lane & 15) : lane, half = two ? (lane >> 4) : 0;
    const int l_g = dl / max(K, 1), l_k = dl - l_g * K;
    const bool drawer = dl < D;
    const int owner_t = lane >> SHIFT;
    const bool owner_rep = (lane & 3) == 0;
    const uint32_t world = (uint32_t)a.neg_world, me = (uint32_t)a.neg_rank;
    const int shift = a.own_shift;

    // rows are striped round-robin: row id -> owner.  The usual power-of-two GPU count takes a mask instead of a division.
    const bool pow2 = (world & (world - 1u)) == 0u;
    auto owned = [&](uint32_t id) -> bool {
        const uint32_t stripe = id >> shift;
        return (pow2 ? (stripe & (world - 1u)) : (stripe % world)) == me;
    };
    // Positives: context nc of occurrence j + half is this lane's slot.  The context TOKENS of a whole chunk (32 entries x up to 16
    // contexts) are staged in shared memory with 4-byte cp.async one chunk ahead -- they are random reads of the gathered token
    // array, a DRAM latency each, and a round of two occurrences is far too short to hide one.
    const int nc = two ? (lane & 15) : lane;
    const int r = a.radius;
    int *ctx = ctx_s[threadIdx.x >> 5][0];
    auto stage_ctx = [&](int buf, int eu_, int er_) {
        if (!POS) return;
        int base = -1;                                         // index of token (sequence, centre position - r) of this lane's entry
        if (er_ >= 0) {
            const uint32_t sq = (uint32_t)eu_ / (uint32_t)a.n_cen;
            base = (int)(sq * (uint32_t)a.seq_len + ((uint32_t)eu_ - sq * (uint32_t)a.n_cen));
        }
        const int n = lane & 15, off = n < r ? n : n + 1;      // window offset of context n (the centre is skipped)
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int ent = i * 2 + (lane >> 4);
            const int b = __shfl_sync(FULL, base, ent);
            if (n < N && b >= 0) cp_async4(ctx + buf * 512 + ent * 16 + n, a.tokens + b + off);
        }
        cp_async_commit();
    };

    float loss_pos = 0.f, loss_neg = 0.f;
    unsigned cnt_fp = 0, cnt_neg = 0, cnt_recall = 0, cnt_pairs = 0;

    float cen[4] = {0.f, 0.f, 0.f, 0.f}, tot[4] = {0.f, 0.f, 0.f, 0.f}, nxt[4] = {0.f, 0.f, 0.f, 0.f};
    int cur_row = -1;
    int eu = 0, er = -1, nu = 0, nr = -1;                      // entries of this chunk / of the next one (one per lane)
    if (c_begin < c_end) {
        const int64_t i = c_begin * 32 + lane;
        if (i < n_entries) { const int2 t = __ldg(entries + i); eu = t.x; er = t.y; }
        const int first = __shfl_sync(FULL, er, 0);
        if (ok) load_vec<4>(a.w_in + (int64_t)first * E + eoff, nxt);
        stage_ctx(0, eu, er);
    }

    for (int64_t c = c_begin; c < c_end; ++c) {
        nu = 0; nr = -1;
        {
            const int64_t i = (c + 1) * 32 + lane;
            if (c + 1 < c_end && i < n_entries) { const int2 t = __ldg(entries + i); nu = t.x; nr = t.y; }
        }
        const int buf = (int)((c - c_begin) & 1);
        if (POS) {
            stage_ctx(buf ^ 1, nu, nr);                        // the next chunk's context tokens travel while this chunk is computed
            cp_async_wait_group1();                            // ... and this chunk's have arrived
            __syncwarp();
        }
        const int prev = __shfl_up_sync(FULL, er, 1);
        const int n_valid = __popc(__ballot_sync(FULL, er >= 0));                        // valid lanes form a prefix
        unsigned heads = __ballot_sync(FULL, er >= 0 && (lane == 0 || er != prev));
        while (heads) {
            const int b = __ffs(heads) - 1;
            heads &= heads - 1;
            const int e = heads ? (__ffs(heads) - 1) : n_valid;
            const int row = __shfl_sync(FULL, er, b);
            // the centre row was fetched while the previous run was computed, i.e. before that run's update went out: when the previous
            // run was the same row (a long run cut at a chunk boundary) my own update is added here
            const bool same = row == cur_row;
#pragma unroll
            for (int x = 0; x < 4; ++x) { cen[x] = nxt[x] + (same ? tot[x] : 0.f); tot[x] = 0.f; }
            cur_row = row;
            const int next_row = heads ? __shfl_sync(FULL, er, __ffs(heads) - 1) : __shfl_sync(FULL, nr, 0);
            if (next_row >= 0 && ok) load_vec<4>(a.w_in + (int64_t)next_row * E + eoff, nxt);

            int count = 0;                                     // ids waiting in the list (warp-uniform)
            for (int j = b; j < e; j += per) {
                // ---- which of the N*K negatives of occurrence j (and j + 1) live in my HBM? ---------------------------------
                const int jj = j + half;
                const uint32_t u = (uint32_t)__shfl_sync(FULL, eu, jj & 31);
                int ids[4] = {0, 0, 0, 0};
                bool own[4] = {false, false, false, false};
                if (drawer && jj < e) {
                    const uint64_t cid = (uint64_t)(a.id_base + (int64_t)u);
                    const uint4 wb = neg_words(a.seed, cid, l_g * 4, l_k, STREAM_NEG);
                    uint4 wc = make_uint4(0, 0, 0, 0);
                    if (a.alias_prob) wc = neg_words(a.seed, cid, l_g * 4, l_k, STREAM_NEG_COIN);
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        if (l_g * 4 + x < N) {
                            const uint32_t id = (uint32_t)draw_row(a.alias_prob, a.alias_idx, (uint32_t)a.vocab, pick_word(wb, x), pick_word(wc, x));
                            ids[x] = (int)id;
                            own[x] = owned(id);
                        }
                    }
                }
                // ---- and which of the 2r context rows?  (their tokens were staged while the previous chunk was computed) ----------
                bool own_p = false;
                int pid = 0;
                if (POS && nc < N && jj < e) {
                    pid = ctx[buf * 512 + jj * 16 + nc] + a.row_offset;
                    own_p = owned((uint32_t)pid);
                }
                const int cnt = (int)own[0] + (int)own[1] + (int)own[2] + (int)own[3] + (int)own_p;
                int incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, off);
                    if (lane >= off) incl += v;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                {
                    int pos = count + incl - cnt;
                    if (own_p) list[pos++] = pid | (int)0x80000000;              // sign bit: a positive pair
#pragma unroll
                    for (int x = 0; x < 4; ++x) if (own[x]) list[pos++] = ids[x];
                }
                count += total;
                const bool last = j + per >= e;
                const int todo = last ? count : (count & ~(P - 1));          // full passes; the ragged one only when the run ends
                if (todo == 0) continue;
                __syncwarp();

                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                // one pass = up to eight listed rows: gather, dots, one transposed reduction, sigmoids side by side, updates.  Nearly all
                // passes are full; their copy of the code carries no per-row predicates
                auto do_pass = [&](auto full_c, const int c0, const int m) {
                    constexpr bool FULLP = decltype(full_c)::value;
                    const int4 la = *reinterpret_cast<const int4 *>(list + c0), lb = *reinterpret_cast<const int4 *>(list + c0 + 4);
                    const int raw[P] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
                    int tid[P];
                    float rowv[P][4];
                    float dot[P];
#pragma unroll
                    for (int t = 0; t < P; ++t) tid[t] = (FULLP || t < m) ? (raw[t] & 0x7fffffff) : 0;
#pragma unroll
                    for (int t = 0; t < P; ++t) {
                        if (!(FULLP && EXACT)) rowv[t][0] = rowv[t][1] = rowv[t][2] = rowv[t][3] = 0.f;
                        if ((FULLP || t < m) && ok) load_vec<4>(a.w_out + (int64_t)tid[t] * E + eoff, rowv[t]);
                    }
#pragma unroll
                    for (int t = 0; t < P; ++t) {
                        float d = 0.f;
#pragma unroll
                        for (int x = 0; x < 4; ++x) d = fmaf(rowv[t][x], cen[x], d);
                        dot[t] = d;
                    }
                    const float sc = transposed_reduce<P>(dot, lane);
                    float step_mine = 0.f;
                    if (FULLP || owner_t < m) {
                        const bool positive = POS && list[c0 + owner_t] < 0;
                        const float xs = positive ? sc : -sc;                         // loss = -log clamp(sigmoid(+-s), 1e-6)   (loss.py:15-16)
                        const float ex = __expf(-xs);
                        const float sig = __fdividef(1.0f, 1.0f + ex);
                        const float gmag = (sig > CLAMP_MIN) ? ex * sig : 0.f;        // |dL/ds| = sigmoid(-xs)
                        step_mine = positive ? a.lr * gmag : -a.lr * gmag;
                        if (owner_rep) {
                            const float lo = -__logf(fmaxf(sig, CLAMP_MIN));
                            if (positive) { loss_pos += lo; cnt_recall += xs >= 0.f; cnt_pairs += 1; }
                            else { loss_neg += lo; cnt_fp += xs <= 0.f; cnt_neg += 1; }
                        }
                    }
#pragma unroll
                    for (int t = 0; t < P; ++t) {
                        const float step = __shfl_sync(FULL, step_mine, t << SHIFT);
                        if (FULLP || t < m) {
                            float d[4];
#pragma unroll
                            for (int x = 0; x < 4; ++x) { acc[x] = fmaf(step, rowv[t][x], acc[x]); d[x] = step * cen[x]; }
                            if (ok) red_vec<4>(a.w_out + (int64_t)tid[t] * E + eoff, d, false);      // my own HBM: device scope is enough
                        }
                    }
                };
                for (int c0 = 0; c0 < todo; c0 += P) {
                    if (todo - c0 >= P) do_pass(std::true_type{}, c0, P);
                    else do_pass(std::false_type{}, c0, todo - c0);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) { cen[x] += acc[x]; tot[x] += acc[x]; }
                // the ids of the unfinished pass move to the front of the list
                const int rem = count - todo;
                int keep = 0;
                if (lane < rem) keep = list[todo + lane];
                __syncwarp();
                if (lane < rem) list[lane] = keep;
                count = rem;
                __syncwarp();
            }
            if (ok) red_vec<4>(a.w_in + (int64_t)row * E + eoff, tot, a.sys_scope);
        }
        eu = nu; er = nr;
    }

    flush_stats(a.stats, cnt_neg != 0 || cnt_pairs != 0, loss_pos, loss_neg, cnt_recall, cnt_fp, cnt_pairs, (double)cnt_neg);
}

template <bool EXACT, bool POS>
int launch_owned_pairs(const SgnsArgs &a, const int2 *entries, const int64_t *n_entries, cudaStream_t stream) {
    auto kern = sgns_owned_pairs_kernel<EXACT, POS>;
    const int blocks = persistent_blocks(kern, 0, (a.n_units + 31) / 32, SGNS_THREADS / 32);
    if (blocks <= 0) return SE_ERR_CUDA;
    kern<<<blocks, SGNS_THREADS, 0, stream>>>(a, entries, n_entries);
    return check_cuda(cudaGetLastError(), "sgns_owned_pairs_kernel launch");
}

}  // namespace
}  // namespace se

using se::COARSE_BUCKETS;

// scratch layout (bytes): cnt[vocab + 1] u64 | start[vocab + 1] i64 | coarse cursors[2048] u64 | scan scratch |
//                         entries[n_centres] int2 (16-byte aligned) | coarse-ordered copy[n_centres] int2
extern "C" int64_t se_pairs_owned_scratch_bytes(int64_t vocab, int64_t n_seq, int seq_len, int radius) {
    if (vocab < 1 || n_seq < 0 || radius < 1 || seq_len < 2 * radius + 1) return -1;
    const int64_t n_units = n_seq * (seq_len - 2 * radius);
    return 8 * (2 * (vocab + 1) + COARSE_BUCKETS + se::scan_scratch_elems(vocab + 1)) + 16 * n_units + 64;
}

extern "C" int se_sgns_update_pairs_owned(float *w_in, float *w_out, int64_t vocab, int emb, const int32_t *tokens, int64_t n_seq,
                                          int seq_len, int radius, int n_neg, int positives, int row_offset,
                                          const float *alias_prob, const int32_t *alias_idx, float lr, uint64_t seed,
                                          int64_t centre_id_base, const se_shard_spec *spec, void *scratch,
                                          int64_t scratch_bytes, double *stats, void *stream) {
    const char *fn = "se_sgns_update_pairs_owned";
    SE_REQUIRE(w_in && w_out && vocab >= 1, "%s: null table or empty vocabulary", fn);
    SE_REQUIRE(spec, "%s: a shard spec is required", fn);
    SE_REQUIRE(spec->world >= 1 && spec->rank >= 0 && spec->rank < spec->world && spec->stripe_rows >= 1,
               "%s: bad shard spec (world %d rank %d stripe_rows %lld)", fn, spec->world, spec->rank, (long long)spec->stripe_rows);
    SE_REQUIRE(n_seq >= 0 && (tokens || n_seq == 0), "%s: null tokens", fn);
    SE_REQUIRE(radius >= 1 && seq_len >= 2 * radius + 1, "Text is too short! [text_length=%d] < [min_text_length=%d]", seq_len, 2 * radius + 1);
    SE_REQUIRE((alias_prob == nullptr) == (alias_idx == nullptr), "%s: pass both alias arrays or neither", fn);
    SE_REQUIRE(n_neg >= 0, "%s: negative n_neg", fn);
    if ((n_neg == 0 && !positives) || n_seq == 0) return SE_OK;
    const int64_t sr = spec->stripe_rows;
    const int64_t n_units = n_seq * (seq_len - 2 * radius);
    if ((sr & (sr - 1)) || emb % 4 != 0 || emb <= 32 || emb > 128 || n_neg > 7 || ((2 * radius + 3) / 4) * n_neg > 32 ||
        ((uintptr_t)w_in % 16) || ((uintptr_t)w_out % 16) || n_seq * seq_len >= 0x7fffffffll || vocab >= 0x7fffffffll) {
        se::set_error("%s: needs 32 < emb <= 128 (multiple of 4), n_neg <= 7, ceil(2r/4)*n_neg <= 32, 16-byte aligned tables, a power-of-two "
                      "stripe_rows and fewer than 2^31 tokens (emb %d, n_neg %d, radius %d, stripe_rows %lld, tokens %lld)",
                      fn, emb, n_neg, radius, (long long)sr, (long long)(n_seq * seq_len));
        return SE_ERR_UNSUPPORTED;
    }
    SE_REQUIRE(scratch && ((uintptr_t)scratch % 16) == 0 && scratch_bytes >= se_pairs_owned_scratch_bytes(vocab, n_seq, seq_len, radius),
               "%s: scratch needs %lld bytes, 16-byte aligned", fn, (long long)se_pairs_owned_scratch_bytes(vocab, n_seq, seq_len, radius));
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;

    se::SgnsArgs a{};
    a.w_in = w_in; a.w_out = w_out; a.tokens = tokens; a.alias_prob = alias_prob; a.alias_idx = alias_idx;
    a.stats = stats; a.vocab = vocab; a.emb = emb; a.n_ctx = 2 * radius; a.n_neg = n_neg;
    a.seq_len = seq_len; a.radius = radius; a.n_cen = seq_len - 2 * radius; a.row_offset = row_offset;
    a.n_units = n_units;
    a.lr = lr; a.seed = seed; a.id_base = centre_id_base;
    a.neg_vocab = (uint32_t)vocab; a.neg_shift = -1; a.neg_world = spec->world; a.neg_rank = spec->rank;
    a.sys_scope = spec->world > 1;
    a.own_shift = 0;
    while ((1ll << a.own_shift) < sr) ++a.own_shift;

    unsigned long long *cnt = reinterpret_cast<unsigned long long *>(scratch);
    int64_t *start = reinterpret_cast<int64_t *>(cnt + vocab + 1);
    unsigned long long *ccur = reinterpret_cast<unsigned long long *>(start + vocab + 1);
    int64_t *scan_scratch = reinterpret_cast<int64_t *>(ccur + COARSE_BUCKETS);
    int64_t *after = scan_scratch + se::scan_scratch_elems(vocab + 1);
    int2 *entries = reinterpret_cast<int2 *>((reinterpret_cast<uintptr_t>(after) + 15) & ~(uintptr_t)15);
    int2 *tmp = entries + n_units;
    int cshift = 0;                                            // coarse bucket = 2^cshift consecutive rows, at most COARSE_BUCKETS of them
    while (((vocab - 1) >> cshift) >= COARSE_BUCKETS) ++cshift;

    SE_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * (size_t)(vocab + 1), st));
    SE_CUDA(cudaMemsetAsync(ccur, 0, sizeof(unsigned long long) * COARSE_BUCKETS, st));
    int64_t cb = (n_units + 255) / 256; if (cb > (int64_t)sms * 16) cb = (int64_t)sms * 16; if (cb < 1) cb = 1;
    se::cen_count_kernel<<<(int)cb, 256, 0, st>>>(a, cnt);
    int rc = se::exclusive_scan(reinterpret_cast<const int64_t *>(cnt), vocab + 1, start, scan_scratch, st);      // start[vocab] = valid centres
    if (rc != SE_OK) return rc;
    int64_t tb = (n_units + se::COARSE_TILE - 1) / se::COARSE_TILE; if (tb > (int64_t)sms * 8) tb = (int64_t)sms * 8; if (tb < 1) tb = 1;
    se::cen_coarse_kernel<<<(int)tb, 256, 0, st>>>(a, start, cshift, ccur, tmp);
    se::cen_fine_kernel<<<(int)cb, 256, 0, st>>>(tmp, start + vocab, start, cnt, entries);
    rc = se::check_cuda(cudaGetLastError(), "centre bucketing");
    if (rc != SE_OK) return rc;
    if (positives)
        return emb == 128 ? se::launch_owned_pairs<true, true>(a, entries, start + vocab, st)
                          : se::launch_owned_pairs<false, true>(a, entries, start + vocab, st);
    return emb == 128 ? se::launch_owned_pairs<true, false>(a, entries, start + vocab, st)
                      : se::launch_owned_pairs<false, false>(a, entries, start + vocab, st);
}

// Shared pieces of the SGNS kernels (sgns.cu, sgns_win_*.cu): kernel arguments, 128-bit row loads / vector reductions, the in-kernel
// negative draw, the loss-statistics reduction and the persistent launch geometry.  Everything lives in an anonymous namespace:
// each translation unit gets its own copy (the argument struct and the cross-file launchers are ordinary declarations).
#pragma once
#include "common.cuh"

namespace se {

struct SgnsArgs {
    float *w_in, *w_out;                 // tables (read-only in MODE_GRAD)
    float *grad_in, *grad_out;           // MODE_GRAD only (may be null)
    const int64_t *inputs, *targets, *noise;   // explicit modes
    const int32_t *tokens;               // MODE_WALK
    const float *alias_prob; const int32_t *alias_idx;
    double *stats;
    int64_t n_units;                     // batch rows (explicit) or n_seq * (L - 2r) centres (walk)
    int64_t vocab;
    int emb, n_ctx, n_neg;
    int seq_len, radius, n_cen, row_offset;
    float lr, grad_scale;
    uint64_t seed; int64_t id_base;
    int scatter_store;
    int force_generic;                   // SE_SGNS_GENERIC_KERNEL: skip the fast path (testing / comparison)
    // sharded tables (se_shard_spec): negatives are drawn over neg_vocab ids; with neg_shift >= 0 those are LOCAL ids of
    // shard neg_rank (stripes of 1 << neg_shift rows, stripe s owned by rank s % neg_world) and are mapped to table rows
    uint32_t neg_vocab;
    int neg_shift, neg_world, neg_rank;
    int sys_scope;                       // rows may live in peer HBM: system-scope reductions
    int no_window;                       // SE_SGNS_NO_WINDOW: per-context kernel instead of the window-resident one
    int own_shift;                       // sgns_negown_kernel: log2(stripe_rows); row r is owned by (r >> own_shift) % neg_world
    int whole_seq;                       // window kernel: a group's span is rounded up to whole sequences (SE_SGNS_WHOLE_SEQUENCES)
    int64_t n_seq;                       // MODE_WALK: number of sequences
    int win_refresh;                     // window kernel: re-fetch a resident row when its token is the centre (SE_SGNS_WINDOW_REFRESH)
    int batch_pos;                       // window kernel, n_neg = 0: the 2r pairs of a centre share one butterfly (SE_SGNS_BATCHED_POSITIVES)
    // MODE_GRAD, row-sparse optimisers (se_sgns_adam_step): rows that receive a gradient are flagged and appended to a list
    int32_t *touch_in, *touch_out;       // [vocab] flags, zero between steps (may be null)
    int32_t *list_in, *list_out;         // touched rows
    int32_t *touch_counts;               // [0] = entries of list_in, [1] = entries of list_out
};

// row-sparse Adam apply (adam.cu): one warp per listed row; restores g = 0, flag = 0 and the list length
int adam_apply(float *w, float *m, float *v, float *g, int32_t *t, int32_t *flags, const int32_t *list, int32_t *count, int32_t *done,
               int64_t capacity, int emb, float lr, float beta1, float beta2, float eps, cudaStream_t stream);

// window-resident kernel family (sgns_win.cuh), one translation unit per lane-group width so they compile in parallel;
// each returns SE_ERR_UNSUPPORTED when the shape is not covered
int launch_win_g32(const SgnsArgs &a, cudaStream_t stream);        // 64 < emb <= 128: one centre per warp
int launch_win_g16(const SgnsArgs &a, cudaStream_t stream);        // 32 < emb <= 64: two centres per warp
int launch_win_g8(const SgnsArgs &a, cudaStream_t stream);         // 16 <= emb <= 32: four centres per warp

namespace {

constexpr int SGNS_THREADS = 256;
constexpr float CLAMP_MIN = 1e-6f;
enum { MODE_GRAD = 0, MODE_STEP = 1, MODE_WALK = 2 };

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

// L2-only loads: rows are updated by other SMs (and by our own red.global), so L1 must not serve them.
template <int VEC> __device__ __forceinline__ void load_vec(const float *p, float (&v)[VEC]) {
    if constexpr (VEC == 4) { float4 t = __ldcg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (VEC == 2) { float2 t = __ldcg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y; }
    else { v[0] = __ldcg(p); }
}
template <int VEC> __device__ __forceinline__ void store_vec(float *p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) __stcg(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    else if constexpr (VEC == 2) __stcg(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
    else __stcg(p, v[0]);
}
// no-return vector reduction at L2 (sm_90+): one instruction per 16 bytes.  `sys` selects system scope, required when
// the row may live in a peer GPU's HBM (sharded tables): the reduction then executes at the owner's L2 over NVLink.
template <int VEC> __device__ __forceinline__ void red_vec(float *p, const float (&v)[VEC], bool sys = false) {
    if (sys) {
        if constexpr (VEC == 4)
            asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
        else if constexpr (VEC == 2)
            asm volatile("red.relaxed.sys.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
        else
            asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory");
        return;
    }
    if constexpr (VEC == 4)
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    else if constexpr (VEC == 2)
        asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    else
        asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory");
}

// Shuffles name only the lanes of the calling group: groups of one warp may run different trip counts.
template <int G> __device__ __forceinline__ unsigned group_mask() {
    if constexpr (G == 32) return FULL;
    else return ((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1));
}
template <int G> __device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
    for (int off = G >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off, G);
    return v;
}


// in-kernel negative: uniform / alias draw, then (local negatives) local id -> row of the stripe this rank owns
__device__ __forceinline__ int neg_row(const SgnsArgs &a, uint32_t r0, uint32_t r1) {
    uint32_t j = (uint32_t)draw_row(a.alias_prob, a.alias_idx, a.neg_vocab, r0, r1);
    if (a.neg_shift >= 0) {
        const uint32_t s = j >> a.neg_shift;
        j = ((s * (uint32_t)a.neg_world + (uint32_t)a.neg_rank) << a.neg_shift) | (j & ((1u << a.neg_shift) - 1u));
    }
    return (int)j;
}

// first toucher of a row appends it to the list (flags are zero between steps)
__device__ __forceinline__ void mark_row(int32_t *flags, int32_t *list, int32_t *count, int64_t row) {
    if (flags[row] == 0 && atomicExch(flags + row, 1) == 0) list[atomicAdd(count, 1)] = (int32_t)row;
}

template <bool FAST> __device__ __forceinline__ float sigmoidf_(float x) {
    if constexpr (FAST) return __fdividef(1.0f, 1.0f + __expf(-x));
    else return 1.0f / (1.0f + expf(-x));
}
template <bool FAST> __device__ __forceinline__ float logf_(float x) {
    if constexpr (FAST) return __logf(x); else return logf(x);
}

// Block-level reduction of the loss statistics (layout: SE_STATS_LEN in se_b200.h): every contributing thread adds into shared
// memory, then one double atomic per statistic per block goes to `stats`.  Must be reached by all threads of the block.
__device__ __forceinline__ void flush_stats(double *stats, bool contribute, float loss_pos, float loss_neg, unsigned cnt_recall,
                                            unsigned cnt_fp, unsigned cnt_pairs, double cnt_neg) {
    __shared__ double sred[SE_STATS_LEN];
    if (threadIdx.x < SE_STATS_LEN) sred[threadIdx.x] = 0.0;
    __syncthreads();
    if (contribute) {
        atomicAdd(&sred[0], (double)loss_pos);
        atomicAdd(&sred[1], (double)loss_neg);
        atomicAdd(&sred[2], (double)cnt_recall);
        atomicAdd(&sred[3], (double)cnt_fp);
        atomicAdd(&sred[4], (double)cnt_pairs);
        atomicAdd(&sred[5], cnt_neg);
    }
    __syncthreads();
    if (threadIdx.x < SE_STATS_LEN && stats && sred[threadIdx.x] != 0.0) atomicAdd(stats + threadIdx.x, sred[threadIdx.x]);
}

// Persistent launch geometry shared by the SGNS kernels: enough blocks of SGNS_THREADS for `n_units` at `units_per_block`, capped at one
// resident wave (SMs x occupancy).  Returns 0 blocks on error (message set).
template <typename Kernel>
int persistent_blocks(Kernel kern, size_t smem, int64_t n_units, int units_per_block, bool need_resident = false, int threads = SGNS_THREADS) {
    int occ = 0;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem), "occupancy") != SE_OK) return 0;
    if (occ < 1) {
        if (need_resident) return -1;
        occ = 1;
    }
    const int sms = sm_count();
    if (sms <= 0) return 0;
    int64_t blocks = (n_units + units_per_block - 1) / units_per_block;
    const int64_t cap = (int64_t)sms * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// Transposed butterfly over a group of W lanes (W = 32: the whole warp): CH partial dots per lane in, ONE fully reduced dot per
// lane out -- lane l of the group ends up with the dot of row (l >> log2(W / CH)) -- in CH - 1 + log2(W / CH) shuffles instead of
// CH * log2(W).  `mask` names the lanes of the calling group, `lg` is the lane's index inside it.
template <int CH, int W = 32> __device__ __forceinline__ float transposed_reduce(float (&v)[CH], int lg, unsigned mask = FULL) {
    static_assert(CH <= W, "more dots than lanes");
    int off = W / 2;
#pragma unroll
    for (int n = CH; n > 1; n >>= 1, off >>= 1) {
        const bool hi = (lg & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = hi ? v[i] : v[i + n / 2];
            const float keep = hi ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(mask, send, off, W);
        }
    }
    float r = v[0];
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(mask, r, off, W);
    return r;
}

__device__ __forceinline__ void cp_async16(float4 *smem_dst, const float *gsrc) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

}  // namespace
}  // namespace se

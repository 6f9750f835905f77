// Dense A . B^T on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a: the only contraction-shaped work near the hot path.
//
//   c[i, j] = scale_a[i] * scale_b[j] * sum_k a[i, k] * b[j, k]          a [m x kd], b [n x kd], c [m x n], fp32 row-major
//
// Users: pairwise cosine similarity of embedding rows (tools/model_analysis.py:33-83 -> utils/func.py:7-20: x / |x| times (y / |y|)^T;
// the row norms enter as scale_a / scale_b in the epilogue) and the optional SHARED-NEGATIVES batch mode of SGNS, where the scores of
// B centres against S shared negatives are one dense B x S GEMM (word2vec/model.py:88 is a batched mat-vec per pair otherwise).
//
// One CTA (128 threads) per 128 x 128 output tile.  Per K block of 32 floats the four warps stage the A and B tiles in shared
// memory in the canonical K-major, no-swizzle UMMA layout (8-row x 16-byte core matrices; chunk c of 4 floats of row r at
// [c][r] float4), one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 with the fp32 accumulator tile (128 lanes x 128
// columns) in tensor memory, tcgen05.commit arrives on an mbarrier when the tile may be overwritten.  fp32 inputs are split as
// a = hi + lo (hi = the tf32 part) and three MMAs hi*hi + hi*lo + lo*hi recover fp32-level accuracy (3xTF32), so results match a
// plain fp32 matmul to ~1e-6 instead of tf32's 1e-3.  The epilogue reads the accumulator with tcgen05.ld (each warp its own 32
// TMEM lanes = 32 output rows), applies the two scale vectors and writes c.  Small problems (a few thousand rows): one stage, no TMA.
#include "common.cuh"

namespace se {
namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int GEMM_THREADS = 128;
constexpr int CHUNKS = BK / 4;                        // 16-byte chunks per row per K block
constexpr int TILE_FLOAT4 = CHUNKS * BM;              // one operand tile: [CHUNKS][128] float4 = 16 KB
constexpr uint32_t TMEM_COLS = 128;
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), SWIZZLE_NONE, K-major: start address, leading-dimension byte offset
// (distance between the two 16-byte K chunks of one MMA), stride byte offset (distance between 8-row core matrices), all >> 4
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46);                       // descriptor version 1 (Blackwell); base offset 0, layout type 0 = no swizzle
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(IDESC), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// bounded wait: a descriptor / barrier mistake must fault, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

#define TMEM_LD_32(r, taddr)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"               \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                   \
                 : "r"(taddr)                                                                                             \
                 : "memory")

// stage one 128-row operand tile of a K block: row `row0 + t` (zero beyond `rows` / `kd`), split into tf32 hi and the remainder
__device__ __forceinline__ void stage_tile(const float *__restrict__ src, int64_t rows, int kd, int64_t row0, int k0, float4 *hi, float4 *lo) {
    const int t = threadIdx.x;
    const int64_t row = row0 + t;
    const bool vec_ok = (kd % 4 == 0) && (((uintptr_t)src & 15u) == 0);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int k = k0 + 4 * c;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (row < rows) {
            const float *p = src + row * kd + k;
            if (vec_ok && k + 4 <= kd) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (k + e < kd) v[e] = __ldg(p + e);
            }
        }
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            h[e] = __uint_as_float(__float_as_uint(v[e]) & 0xFFFFE000u);      // what the tensor core keeps of an fp32 operand
            l[e] = v[e] - h[e];
        }
        hi[c * BM + t] = make_float4(h[0], h[1], h[2], h[3]);
        lo[c * BM + t] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_nt_tf32x3_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t m, int64_t n, int kd, const float *__restrict__ scale_a,
                      const float *__restrict__ scale_b, float *__restrict__ c) {
    extern __shared__ __align__(1024) float4 gemm_smem[];
    float4 *a_hi = gemm_smem, *a_lo = a_hi + TILE_FLOAT4, *b_hi = a_lo + TILE_FLOAT4, *b_lo = b_hi + TILE_FLOAT4;
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_done)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    const uint32_t bar = smem_u32(&mma_done);
    uint32_t parity = 0;
    const int n_kb = (kd + BK - 1) / BK;
    for (int kb = 0; kb < n_kb; ++kb) {
        stage_tile(a, m, kd, m0, kb * BK, a_hi, a_lo);
        stage_tile(b, n, kd, n0, kb * BK, b_hi, b_lo);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t lbo = BM * 16, sbo = 128;                          // next K chunk: one [128] float4 slab on; next 8 rows: 128 bytes on
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk) {                            // one MMA covers K = 8 tf32 values = two 16-byte chunks
                const uint32_t off = (uint32_t)kk * 2u * BM * 16u;
                const uint64_t dah = umma_desc(smem_u32(a_hi) + off, lbo, sbo), dal = umma_desc(smem_u32(a_lo) + off, lbo, sbo);
                const uint64_t dbh = umma_desc(smem_u32(b_hi) + off, lbo, sbo), dbl = umma_desc(smem_u32(b_lo) + off, lbo, sbo);
                umma_tf32(tmem_d, dal, dbh, (kb | kk) != 0 ? 1u : 0u);        // small terms first
                umma_tf32(tmem_d, dah, dbl, 1u);
                umma_tf32(tmem_d, dah, dbh, 1u);
            }
            // arrives on the mbarrier once every MMA issued so far has completed (implies tcgen05.fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        mbar_wait(bar, parity);                                              // the tiles may be overwritten / the accumulator is complete
        parity ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w owns TMEM lanes 32w .. 32w + 31 = output rows m0 + 32w + lane ------------------------------------
    const int64_t row = m0 + 32 * warp + lane;
    const float sa = (scale_a != nullptr && row < m) ? scale_a[row] : 1.0f;
#pragma unroll 1
    for (int j0 = 0; j0 < BN; j0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * warp) << 16) + (uint32_t)j0;
        TMEM_LD_32(r, taddr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < m) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int64_t col = n0 + j0 + j;
                if (col < n) c[row * n + col] = __uint_as_float(r[j]) * sa * (scale_b != nullptr ? __ldg(scale_b + col) : 1.0f);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
}

// out[i] = 1 / |x_i|  (utils/func.py:18-19 divides by torch.norm(x, dim=-1)); one warp per row
__global__ void __launch_bounds__(256)
row_inv_norm_kernel(const float *__restrict__ x, int64_t rows, int emb, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < rows; i += n_warps) {
        float s = 0.f;
        for (int e = lane; e < emb; e += 32) { const float v = x[i * emb + e]; s = fmaf(v, v, s); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
        if (lane == 0) out[i] = 1.0f / sqrtf(s);
    }
}

// k largest entries of every row, descending (torch.argsort(sim[i], descending=True)[:k], tools/model_analysis.py:71); one warp per
// row, k selection passes, ties broken by the smaller column
__global__ void __launch_bounds__(256)
topk_rows_kernel(const float *__restrict__ x, int64_t rows, int64_t cols, int k, int64_t *__restrict__ idx_out, float *__restrict__ val_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < rows; i += n_warps) {
        const float *xr = x + i * cols;
        float prev_v = INFINITY; int64_t prev_j = -1;
        for (int p = 0; p < k; ++p) {
            float best_v = -INFINITY; int64_t best_j = cols;
            for (int64_t j = lane; j < cols; j += 32) {
                const float v = xr[j];
                const bool after_prev = (v < prev_v) || (v == prev_v && j > prev_j);        // strictly later in (value desc, column asc) order
                if (after_prev && (v > best_v || (v == best_v && j < best_j))) { best_v = v; best_j = j; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(FULL, best_v, off);
                const int64_t oj = __shfl_xor_sync(FULL, best_j, off);
                if (ov > best_v || (ov == best_v && oj < best_j)) { best_v = ov; best_j = oj; }
            }
            if (lane == 0) { idx_out[i * k + p] = best_j < cols ? best_j : -1; if (val_out) val_out[i * k + p] = best_v; }
            prev_v = best_v; prev_j = best_j;
        }
    }
}

__global__ void __launch_bounds__(256)
transpose_kernel(const float *__restrict__ x, int64_t rows, int64_t cols, float *__restrict__ out) {
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
    for (int dy = threadIdx.y; dy < 32; dy += 8) {
        const int64_t r = r0 + dy, cc = c0 + threadIdx.x;
        tile[dy][threadIdx.x] = (r < rows && cc < cols) ? x[r * cols + cc] : 0.f;
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += 8) {
        const int64_t cc = c0 + dy, r = r0 + threadIdx.x;
        if (cc < cols && r < rows) out[cc * rows + r] = tile[threadIdx.x][dy];
    }
}

}  // namespace
}  // namespace se

extern "C" int se_gemm_nt(const float *a, const float *b, int64_t m, int64_t n, int kdim, const float *scale_a, const float *scale_b,
                          float *c, void *stream) {
    SE_REQUIRE(m >= 0 && n >= 0 && kdim >= 1, "se_gemm_nt: bad shape");
    if (m == 0 || n == 0) return SE_OK;
    SE_REQUIRE(a && b && c, "se_gemm_nt: null pointer");
    const int64_t gy = (m + se::BM - 1) / se::BM, gx = (n + se::BN - 1) / se::BN;
    SE_REQUIRE(gy <= 65535, "se_gemm_nt: too many rows for one launch (%lld; tile the call)", (long long)m);
    const size_t smem = 4 * (size_t)se::TILE_FLOAT4 * sizeof(float4);
    SE_CUDA(cudaFuncSetAttribute(se::gemm_nt_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    se::gemm_nt_tf32x3_kernel<<<dim3((unsigned)gx, (unsigned)gy), se::GEMM_THREADS, smem, (cudaStream_t)stream>>>(a, b, m, n, kdim, scale_a,
                                                                                                           scale_b, c);
    return se::check_cuda(cudaGetLastError(), "gemm_nt_tf32x3_kernel launch");
}

extern "C" int se_row_inv_norms(const float *x, int64_t rows, int emb, float *out, void *stream) {
    SE_REQUIRE(rows >= 0 && emb >= 1, "se_row_inv_norms: bad shape");
    if (rows == 0) return SE_OK;
    SE_REQUIRE(x && out, "se_row_inv_norms: null pointer");
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (rows + 7) / 8; if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::row_inv_norm_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, emb, out);
    return se::check_cuda(cudaGetLastError(), "row_inv_norm_kernel launch");
}

extern "C" int se_cosine_similarity(const float *x, const float *y, int64_t m, int64_t n, int emb, float *inv_norms, float *out, void *stream) {
    SE_REQUIRE(inv_norms || (m == 0 && n == 0), "se_cosine_similarity: scratch for m + n inverse norms required");
    int rc = se_row_inv_norms(x, m, emb, inv_norms, stream);
    if (rc != SE_OK) return rc;
    rc = se_row_inv_norms(y, n, emb, inv_norms + m, stream);
    if (rc != SE_OK) return rc;
    return se_gemm_nt(x, y, m, n, emb, inv_norms, inv_norms + m, out, stream);
}

extern "C" int se_topk_rows(const float *x, int64_t rows, int64_t cols, int k, int64_t *idx_out, float *val_out, void *stream) {
    SE_REQUIRE(rows >= 0 && cols >= 0 && k >= 1, "se_topk_rows: bad shape");
    if (rows == 0) return SE_OK;
    SE_REQUIRE(x && idx_out, "se_topk_rows: null pointer");
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (rows + 7) / 8; if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::topk_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, k, idx_out, val_out);
    return se::check_cuda(cudaGetLastError(), "topk_rows_kernel launch");
}

extern "C" int se_transpose(const float *x, int64_t rows, int64_t cols, float *out, void *stream) {
    SE_REQUIRE(rows >= 0 && cols >= 0, "se_transpose: bad shape");
    if (rows == 0 || cols == 0) return SE_OK;
    SE_REQUIRE(x && out, "se_transpose: null pointer");
    const int64_t gy = (rows + 31) / 32, gx = (cols + 31) / 32;
    SE_REQUIRE(gy <= 65535, "se_transpose: too many rows for one launch");
    se::transpose_kernel<<<dim3((unsigned)gx, (unsigned)gy), dim3(32, 8), 0, (cudaStream_t)stream>>>(x, rows, cols, out);
    return se::check_cuda(cudaGetLastError(), "transpose_kernel launch");
}

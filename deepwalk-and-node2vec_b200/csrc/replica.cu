// Multi-GPU SGNS with the reference's GLOBAL negative draw: per-GPU working copies + row-sharded masters, synchronised by ONE
// kernel over NVLink / NVSwitch peer memory.
//
// Why: with negatives uniform over the whole table (word2vec/utils/sampling.py:21) a step of the S3 workload touches every
// row of both tables many times (168 M tokens and 0.9 G negative draws per GPU over 10 M rows), so the deduplicated "rows
// out, gradients back" exchange of a row-sharded word2vec degenerates to "every row out, every row's summed update back".
// Doing that per pair (random 512-byte peer reads / reds) is NVLink-latency bound (~220 GB/s per direction measured);
// doing it ONCE per step as bulk coalesced traffic runs at the link rate.  So every GPU trains on a full working copy in
// its own HBM with the unchanged single-GPU kernels (Hogwild, device-scope reductions), and rank r OWNS the master of rows
// chunk r.  After a step, the owner of each element computes
//         master' = master + beta * sum_g (copy_g - master)     (beta = 1: the SUMMED updates of all GPUs since the last sync;
//                                                                beta = 1/G: their MEAN = local SGD with model averaging)
// reading the G copies over NVLink (7/8 of them peers), and stores master' back into all G copies -- reduce-scatter of the
// updates and all-gather of the rows fused in one pass, 2 (G-1)/G of the table per direction per GPU, no staging buffers,
// no NCCL on the data path (a barrier on either side is the only collective).  sum-of-differences keeps the arithmetic
// exact where copies did not move and equals synchronous data-parallel SGD with summed gradients.
//
// Layout: one virtual range of `world` segments of `stride_elems` floats, segment g physically in GPU g's HBM and mapped
// into every process (csrc/shard.cu VMM plumbing, shallow_encoders/word2vec/sharded.py::ReplicatedTable).
#include "common.cuh"

namespace se {
namespace {

constexpr int REPLICA_MAX_WORLD = 16;

__device__ __forceinline__ float4 ld_sys(const float4 *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(float4 *p, const float4 &v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// mode 0: sync (see above); mode 1: master <- own copy (initialisation); mode 2: every copy <- master (after a checkpoint load)
template <int W>
__global__ void __launch_bounds__(256)
replica_sync_kernel(float *__restrict__ base, int64_t stride_elems, int world, int rank, int64_t lo4, int64_t hi4,
                    float4 *__restrict__ master, int mode, float beta) {
    const int G = W > 0 ? W : world;
    for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += (int64_t)gridDim.x * blockDim.x) {
        if (mode == 1) {
            master[i - lo4] = ld_sys(reinterpret_cast<const float4 *>(base + (int64_t)rank * stride_elems) + i);
            continue;
        }
        const float4 m = master[i - lo4];
        float4 nv = m;
        if (mode == 0) {
            float4 s[W > 0 ? W : REPLICA_MAX_WORLD];
#pragma unroll
            for (int g = 0; g < (W > 0 ? W : REPLICA_MAX_WORLD); ++g)
                if (g < G) s[g] = ld_sys(reinterpret_cast<const float4 *>(base + (int64_t)g * stride_elems) + i);     // all loads in flight first
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int g = 0; g < (W > 0 ? W : REPLICA_MAX_WORLD); ++g)
                if (g < G) { acc.x += s[g].x - m.x; acc.y += s[g].y - m.y; acc.z += s[g].z - m.z; acc.w += s[g].w - m.w; }
            nv = make_float4(fmaf(beta, acc.x, m.x), fmaf(beta, acc.y, m.y), fmaf(beta, acc.z, m.z), fmaf(beta, acc.w, m.w));
            master[i - lo4] = nv;
        }
#pragma unroll
        for (int g = 0; g < (W > 0 ? W : REPLICA_MAX_WORLD); ++g)
            if (g < G) st_sys(reinterpret_cast<float4 *>(base + (int64_t)g * stride_elems) + i, nv);
    }
}

void chunk_bounds(int64_t n_elems, int world, int rank, int64_t &lo4, int64_t &hi4) {
    const int64_t n4 = (n_elems + 3) / 4;
    const int64_t chunk = (n4 + world - 1) / world;
    lo4 = (int64_t)rank * chunk; if (lo4 > n4) lo4 = n4;
    hi4 = lo4 + chunk; if (hi4 > n4) hi4 = n4;
}

}  // namespace
}  // namespace se

extern "C" int se_replica_chunk(int64_t n_elems, int world, int rank, int64_t *lo_elem, int64_t *hi_elem) {
    SE_REQUIRE(lo_elem && hi_elem && n_elems >= 0 && world >= 1 && rank >= 0 && rank < world, "se_replica_chunk: bad arguments");
    int64_t lo4, hi4;
    se::chunk_bounds(n_elems, world, rank, lo4, hi4);
    *lo_elem = lo4 * 4; *hi_elem = hi4 * 4;
    return SE_OK;
}

extern "C" int se_replica_sync(float *base, int64_t stride_elems, int world, int rank, int64_t n_elems, float *master, int mode,
                               float beta, void *stream) {
    SE_REQUIRE(base && master, "se_replica_sync: null pointer");
    SE_REQUIRE(world >= 1 && world <= se::REPLICA_MAX_WORLD && rank >= 0 && rank < world, "se_replica_sync: bad world %d / rank %d (max %d GPUs)",
               world, rank, se::REPLICA_MAX_WORLD);
    SE_REQUIRE(n_elems >= 0 && stride_elems >= ((n_elems + 3) / 4) * 4 && stride_elems % 4 == 0,
               "se_replica_sync: segment stride %lld too small for %lld elements (must hold them rounded up to 4)", (long long)stride_elems,
               (long long)n_elems);
    SE_REQUIRE(((uintptr_t)base % 16) == 0 && ((uintptr_t)master % 16) == 0, "se_replica_sync: buffers must be 16-byte aligned");
    SE_REQUIRE(mode >= 0 && mode <= 2, "se_replica_sync: unknown mode %d", mode);
    SE_REQUIRE(mode != 0 || (beta > 0.f && beta <= 1.f), "se_replica_sync: merge weight beta must be in (0, 1] (got %g)", (double)beta);
    int64_t lo4, hi4;
    se::chunk_bounds(n_elems, world, rank, lo4, hi4);
    if (hi4 <= lo4) return SE_OK;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (hi4 - lo4 + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;       // 8 x 256 threads per SM, each with `world` 16-byte peer loads in flight
    cudaStream_t st = (cudaStream_t)stream;
    float4 *m4 = reinterpret_cast<float4 *>(master);
    switch (world) {
        case 2: se::replica_sync_kernel<2><<<(int)blocks, 256, 0, st>>>(base, stride_elems, world, rank, lo4, hi4, m4, mode, beta); break;
        case 4: se::replica_sync_kernel<4><<<(int)blocks, 256, 0, st>>>(base, stride_elems, world, rank, lo4, hi4, m4, mode, beta); break;
        case 8: se::replica_sync_kernel<8><<<(int)blocks, 256, 0, st>>>(base, stride_elems, world, rank, lo4, hi4, m4, mode, beta); break;
        default: se::replica_sync_kernel<0><<<(int)blocks, 256, 0, st>>>(base, stride_elems, world, rank, lo4, hi4, m4, mode, beta); break;
    }
    return se::check_cuda(cudaGetLastError(), "replica_sync_kernel launch");
}

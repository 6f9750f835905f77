// Exclusive scan of int64 counts on the device (block sums, scan of the block sums, down-sweep): shared by the graph ingest
// (ingest.cu) and the centre bucketing of the grouped owner-computes negatives (negown.cu).  Anonymous namespace: one copy per
// translation unit.
#pragma once
#include "common.cuh"

namespace se {
namespace {

constexpr int SCAN_THREADS = 1024;

// ---- exclusive scan of int64 counts (three small kernels: block sums, scan of the block sums, down-sweep) ---------------------
__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_kernel(const int64_t *__restrict__ in, int64_t n, int64_t *__restrict__ out, int64_t *__restrict__ block_sums) {
    __shared__ int64_t warp_tot[32];
    const int64_t i = (int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t v = i < n ? in[i] : 0, incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int64_t t = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t w = warp_tot[lane], wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int64_t t = __shfl_up_sync(FULL, wi, off); if (lane >= off) wi += t; }
        warp_tot[lane] = wi - w;                         // exclusive prefix of the warp totals
        if (lane == 31 && block_sums) block_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    if (i < n) out[i] = warp_tot[warp] + incl - v;
}
__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int64_t *__restrict__ out, int64_t n, const int64_t *__restrict__ block_offsets) {
    const int64_t i = (int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x;
    if (i < n) out[i] += block_offsets[blockIdx.x];
}

// out[i] = sum of in[0 .. i-1] for i < n (in and out must not alias); two levels of block sums: n <= 2^30
int exclusive_scan(const int64_t *in, int64_t n, int64_t *out, int64_t *scratch, cudaStream_t st) {
    const int64_t nb0 = (n + SCAN_THREADS - 1) / SCAN_THREADS, nb1 = (nb0 + SCAN_THREADS - 1) / SCAN_THREADS;
    if (nb1 > SCAN_THREADS) { set_error("exclusive_scan: more than 2^30 elements are not supported"); return SE_ERR_UNSUPPORTED; }
    int64_t *sums0 = scratch, *sums0_scan = sums0 + nb0, *sums1 = sums0_scan + nb0, *sums1_scan = sums1 + nb1;
    scan_block_kernel<<<(unsigned)nb0, SCAN_THREADS, 0, st>>>(in, n, out, sums0);
    scan_block_kernel<<<(unsigned)nb1, SCAN_THREADS, 0, st>>>(sums0, nb0, sums0_scan, sums1);
    scan_block_kernel<<<1, SCAN_THREADS, 0, st>>>(sums1, nb1, sums1_scan, nullptr);
    scan_add_kernel<<<(unsigned)nb1, SCAN_THREADS, 0, st>>>(sums0_scan, nb0, sums1_scan);
    scan_add_kernel<<<(unsigned)nb0, SCAN_THREADS, 0, st>>>(out, n, sums0_scan);
    return check_cuda(cudaGetLastError(), "exclusive_scan");
}

// int64 elements of scratch exclusive_scan needs for n inputs
inline int64_t scan_scratch_elems(int64_t n) {
    const int64_t b0 = (n + 1 + SCAN_THREADS - 1) / SCAN_THREADS + 1;
    const int64_t b1 = (b0 + SCAN_THREADS - 1) / SCAN_THREADS + 1;
    return 2 * b0 + 2 * b1 + 8;
}

}  // namespace
}  // namespace se

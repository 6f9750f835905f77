// Micro-benchmark (developer tool): practical HBM ceiling for the SGNS access pattern -- random 512-byte rows,
// gathered with 128-bit loads and scattered back with red.add / plain stores, minimal arithmetic.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void red4(float *p, float4 v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void pf(const float *p) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(512) : "memory");
}

// mode 0: gather only; 1: gather + red; 2: gather + store; 3: gather + red with L2 prefetch one batch ahead
template <int ROWS, int MODE>
__global__ void __launch_bounds__(256) k(float *tab, uint32_t vocab, int64_t batches, float *sink) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int64_t b = warp; b < batches; b += nw) {
        float4 r[ROWS];
        uint32_t id[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; ++i) id[i] = __umulhi(hash32((uint32_t)(b * ROWS + i) * 2654435761u + 12345u), vocab);
        if (MODE == 3) {
            const int64_t b2 = b + nw;
            if (lane < ROWS && b2 < batches) pf(tab + (int64_t)__umulhi(hash32((uint32_t)(b2 * ROWS + lane) * 2654435761u + 12345u), vocab) * 128);
        }
#pragma unroll
        for (int i = 0; i < ROWS; ++i) r[i] = __ldcg(reinterpret_cast<const float4 *>(tab + (int64_t)id[i] * 128) + lane);
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            acc += r[i].x + r[i].w;
            if (MODE == 1 || MODE == 3) red4(tab + (int64_t)id[i] * 128 + lane * 4, make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f));
            if (MODE == 2) __stcg(reinterpret_cast<float4 *>(tab + (int64_t)id[i] * 128) + lane, make_float4(r[i].x + 1e-9f, r[i].y, r[i].z, r[i].w));
        }
    }
    if (acc == 123.456f) *sink = acc;
}

template <int ROWS, int MODE>
void run(float *tab, uint32_t vocab, float *sink, int blocks_per_sm, const char *name) {
    const int64_t batches = (int64_t)(1 << 24) / ROWS * 6;       // ~100 M rows
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * blocks_per_sm;
    k<ROWS, MODE><<<grid, 256>>>(tab, vocab, batches / 8, sink);
    cudaEventRecord(e0);
    k<ROWS, MODE><<<grid, 256>>>(tab, vocab, batches, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)batches * ROWS * 512.0 * (MODE == 0 ? 1 : 2);
    printf("%-28s rows/batch=%d blocks/SM=%d  %8.2f ms  %8.1f GB/s  (%s)\n", name, ROWS, blocks_per_sm, ms, bytes / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

// mode "mix": the row mix of the S3 window kernel without its arithmetic.  One warp per "centre": the centre row of W_in is read at the
// start and reduced at the end; N contexts follow, each gathering K random rows of W_out and reducing into the same rows with values that
// depend on the loaded data (so a red waits for its loads, as in the kernel); per centre one more W_out row is read (the context row that
// enters the window) and a different one is reduced (the row that leaves).  Bytes per centre = 512 * 2 * (1 + N K + 1).
template <int K, int N>
__global__ void __launch_bounds__(256) kmix(float *w_in, float *w_out, uint32_t vocab, int64_t centres, float *sink) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float keep = 0.f;
    for (int64_t c = warp; c < centres; c += nw) {
        const uint32_t base = (uint32_t)c * (uint32_t)(N * K + 3);
        const float4 cen = __ldcg(reinterpret_cast<const float4 *>(w_in + (int64_t)__umulhi(hash32(base * 2654435761u + 1u), vocab) * 128) + lane);
        const float4 ent = __ldcg(reinterpret_cast<const float4 *>(w_out + (int64_t)__umulhi(hash32((base + 1) * 2654435761u + 1u), vocab) * 128) + lane);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n = 0; n < N; ++n) {
            float4 r[K];
            uint32_t id[K];
#pragma unroll
            for (int i = 0; i < K; ++i) id[i] = __umulhi(hash32((base + 3 + n * K + i) * 2654435761u + 1u), vocab);
#pragma unroll
            for (int i = 0; i < K; ++i) r[i] = __ldcg(reinterpret_cast<const float4 *>(w_out + (int64_t)id[i] * 128) + lane);
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const float s = (r[i].x * cen.x + r[i].y * cen.y) * 1e-30f;
                acc.x += s; acc.y += r[i].z * 1e-30f;
                red4(w_out + (int64_t)id[i] * 128 + lane * 4, make_float4(s, s, s, s));
            }
        }
        red4(w_out + (int64_t)__umulhi(hash32((base + 2) * 2654435761u + 1u), vocab) * 128 + lane * 4, make_float4(ent.x * 1e-30f, acc.y, 0.f, 0.f));
        red4(w_in + (int64_t)__umulhi(hash32(base * 2654435761u + 1u), vocab) * 128 + lane * 4, acc);
        keep += acc.x;
    }
    if (keep == 123.456f) *sink = keep;
}

template <int K, int N>
void run_mix(float *w_in, float *w_out, uint32_t vocab, float *sink, int blocks_per_sm) {
    const int64_t centres = 18350080 / 4;                        // a quarter of one S3 launch (262,144 walks x 70 centres)
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * blocks_per_sm;
    kmix<K, N><<<grid, 256>>>(w_in, w_out, vocab, centres / 8, sink);
    cudaEventRecord(e0);
    kmix<K, N><<<grid, 256>>>(w_in, w_out, vocab, centres, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)centres * 512.0 * 2.0 * (1 + N * K + 1);
    printf("S3 row mix (K=%d, N=%d), no arithmetic   blocks/SM=%d (%d warps/SM)  %8.2f ms  %8.1f GB/s  (%s)\n", K, N, blocks_per_sm, blocks_per_sm * 8, ms,
           bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    if (argc > 1 && argv[1][0] == 'm') {                         // ./rowbw mix: only the S3 row-mix ceiling (two tables of 10 M x 128 floats, as S3)
        const uint32_t vocab = 10000001;
        float *w_in, *w_out, *sink;
        cudaMalloc(&w_in, (size_t)vocab * 512); cudaMemset(w_in, 0, (size_t)vocab * 512);
        cudaMalloc(&w_out, (size_t)vocab * 512); cudaMemset(w_out, 0, (size_t)vocab * 512); cudaMalloc(&sink, 4);
        for (int bps : {2, 3, 4, 8}) run_mix<5, 10>(w_in, w_out, vocab, sink, bps);
        return 0;
    }
    const uint32_t vocab = 10000001;
    float *tab, *sink;
    cudaMalloc(&tab, (size_t)vocab * 512); cudaMemset(tab, 0, (size_t)vocab * 512); cudaMalloc(&sink, 4);
    for (int bps : {2, 4, 8}) {
        run<6, 0>(tab, vocab, sink, bps, "gather only");
        run<6, 1>(tab, vocab, sink, bps, "gather + red.add.v4");
        run<6, 2>(tab, vocab, sink, bps, "gather + store");
        run<6, 3>(tab, vocab, sink, bps, "gather + red + L2 prefetch");
    }
    run<12, 1>(tab, vocab, sink, 4, "gather + red.add.v4");
    run<12, 3>(tab, vocab, sink, 4, "gather + red + L2 prefetch");
    // streaming copy for reference
    float *a, *b; size_t n = (size_t)1 << 30; cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemcpy(b, a, n * 4, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0); cudaMemcpy(b, a, n * 4, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D 4 GiB: %.2f ms  %.1f GB/s (read+write)\n", ms, 2.0 * n * 4 / ms / 1e6);
    return 0;
}

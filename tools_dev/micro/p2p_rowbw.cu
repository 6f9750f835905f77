// Micro-benchmark (developer tool): what NVLink sustains for the SGNS access pattern against PEER memory -- random
// 512-byte rows gathered with 128-bit loads and / or scattered with red.add.v4.f32 (system scope), one warp per row batch.
// One process, two GPUs, peer access enabled; "both" runs the same kernel on both GPUs against each other's table.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void red4(float *p, float4 v) {
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red1(float *p, float v) {
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// mode 0: gather; 1: red.v4; 2: gather + red.v4; 3: plain 128-bit stores; 4: scalar red.f32 (4 per lane)
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
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < ROWS; ++i) r[i] = __ldcg(reinterpret_cast<const float4 *>(tab + (int64_t)id[i] * 128) + lane);
#pragma unroll
            for (int i = 0; i < ROWS; ++i) acc += r[i].x + r[i].w;
        }
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            float *p = tab + (int64_t)id[i] * 128 + lane * 4;
            if (MODE == 1 || MODE == 2) red4(p, make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f));
            if (MODE == 3) __stcg(reinterpret_cast<float4 *>(p), make_float4(1e-9f, 0.f, 0.f, 0.f));
            if (MODE == 4) { red1(p, 1e-9f); red1(p + 1, 1e-9f); red1(p + 2, 1e-9f); red1(p + 3, 1e-9f); }
        }
    }
    if (acc == 123.456f) *sink = acc;
}

template <int ROWS, int MODE>
int run(float *tab[2], float *sink[2], uint32_t vocab, int bps, bool both, const char *name) {
    const int64_t batches = (int64_t)(1 << 22) / ROWS * 6;       // ~25 M rows
    cudaEvent_t e0[2], e1[2];
    const int n = both ? 2 : 1;
    for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d])); }
    for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); k<ROWS, MODE><<<148 * bps, 256>>>(tab[1 - d], vocab, batches / 8, sink[d]); }
    for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
    for (int d = 0; d < n; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaEventRecord(e0[d]));
        k<ROWS, MODE><<<148 * bps, 256>>>(tab[1 - d], vocab, batches, sink[d]);
        CK(cudaEventRecord(e1[d]));
    }
    float worst = 0.f;
    for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaEventSynchronize(e1[d])); float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); if (ms > worst) worst = ms; }
    const double bytes = (double)batches * ROWS * 512.0;      // payload per direction of use (gather: inbound, red: outbound)
    printf("%-22s %-5s rows/batch=%d blocks/SM=%d  %8.2f ms  %7.1f GB/s per GPU per direction used\n", name, both ? "both" : "one", ROWS, bps,
           worst, bytes / worst / 1e6);
    return 0;
}

int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    const uint32_t vocab = 5000001;
    float *tab[2], *sink[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&tab[d], (size_t)vocab * 512)); CK(cudaMemset(tab[d], 0, (size_t)vocab * 512)); CK(cudaMalloc(&sink[d], 4));
    }
    for (bool both : {false, true}) {
        for (int bps : {2, 8}) {
            if (run<6, 0>(tab, sink, vocab, bps, both, "peer gather")) return 1;
            if (run<6, 1>(tab, sink, vocab, bps, both, "peer red.v4")) return 1;
            if (run<6, 2>(tab, sink, vocab, bps, both, "peer gather+red.v4")) return 1;
            if (run<6, 3>(tab, sink, vocab, bps, both, "peer store.v4")) return 1;
        }
        if (run<6, 4>(tab, sink, vocab, 8, both, "peer red.f32 x4")) return 1;
        if (run<12, 0>(tab, sink, vocab, 8, both, "peer gather")) return 1;
    }
    // bulk copy for reference
    CK(cudaSetDevice(0));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    size_t nb = (size_t)1 << 31;
    CK(cudaMemcpyPeer(tab[0], 0, tab[1], 1, nb));
    CK(cudaEventRecord(a)); CK(cudaMemcpyPeer(tab[0], 0, tab[1], 1, nb)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    printf("cudaMemcpyPeer 2 GiB: %.2f ms  %.1f GB/s\n", ms, nb / ms / 1e6);
    return 0;
}

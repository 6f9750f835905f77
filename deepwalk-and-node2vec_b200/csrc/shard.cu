// Sharded embedding tables over NVLink / NVSwitch peer memory.
//
// A table is ONE flat fp32 [vocab x emb] array in every process's virtual address space, physically striped over the
// HBM of `world` GPUs: stripe s (stripe_bytes each, a multiple of the allocation granularity) lives on rank s % world.
// Each stripe is its own physical allocation (cuMemCreate) because cuMemMap cannot map at an offset into a handle;
// the owner exports it as a POSIX file descriptor, peers import it and map it at the same offset of their own
// reservation.  The SGNS kernels are unchanged: their 128-bit row loads and red.global.add.v4.f32 scatters travel to the
// owner's L2 over NVLink, so gather -> dot -> scatter and the row / gradient exchange are one kernel.
//
// The driver API is reached through cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda
// (it must load, and export every symbol, on a box without a GPU).
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace se {
namespace {

struct DriverApi {
    CUresult (*GetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);
    CUresult (*AddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*AddressFree)(CUdeviceptr, size_t);
    CUresult (*Create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
    CUresult (*Release)(CUmemGenericAllocationHandle);
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*Unmap)(CUdeviceptr, size_t);
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
    CUresult (*ExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
    CUresult (*ImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType);
    CUresult (*GetErrorString)(CUresult, const char **);
    bool ok;
};

int load_driver(DriverApi &d) {
    static thread_local DriverApi cached = {};
    if (cached.ok) { d = cached; return SE_OK; }
    SE_CUDA(cudaFree(0));     // make sure the primary context exists
    struct { const char *name; void **slot; } syms[] = {
        {"cuMemGetAllocationGranularity", (void **)&cached.GetAllocationGranularity},
        {"cuMemAddressReserve", (void **)&cached.AddressReserve},
        {"cuMemAddressFree", (void **)&cached.AddressFree},
        {"cuMemCreate", (void **)&cached.Create},
        {"cuMemRelease", (void **)&cached.Release},
        {"cuMemMap", (void **)&cached.Map},
        {"cuMemUnmap", (void **)&cached.Unmap},
        {"cuMemSetAccess", (void **)&cached.SetAccess},
        {"cuMemExportToShareableHandle", (void **)&cached.ExportToShareableHandle},
        {"cuMemImportFromShareableHandle", (void **)&cached.ImportFromShareableHandle},
        {"cuGetErrorString", (void **)&cached.GetErrorString},
    };
    for (auto &s : syms) {
        cudaDriverEntryPointQueryResult q;
        SE_CUDA(cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || *s.slot == nullptr) {
            set_error("driver entry point %s not available", s.name);
            return SE_ERR_UNSUPPORTED;
        }
    }
    cached.ok = true;
    d = cached;
    return SE_OK;
}

int check_cu(const DriverApi &d, CUresult r, const char *what) {
    if (r == CUDA_SUCCESS) return SE_OK;
    const char *msg = nullptr;
    if (d.GetErrorString) d.GetErrorString(r, &msg);
    set_error("CUDA driver error in %s: %s (%d)", what, msg ? msg : "?", (int)r);
    return SE_ERR_CUDA;
}

#define SE_CU(d, expr)                                     \
    do {                                                   \
        int _rc = check_cu((d), (d).expr, #expr);          \
        if (_rc != SE_OK) return _rc;                      \
    } while (0)

int stripe_prop(CUmemAllocationProp &prop) {
    int dev = 0;
    SE_CUDA(cudaGetDevice(&dev));
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return SE_OK;
}

constexpr uint32_t STREAM_INIT = 0x50000000u;

// Xavier-uniform fill keyed by the GLOBAL element index: the content of the table does not depend on how it is sharded.
// A rank writes only the stripes it owns (stripe_elems == 0: everything).
__global__ void __launch_bounds__(256)
fill_uniform_kernel(float *__restrict__ w, int64_t n_vec4, int64_t n_elems, float bound, uint64_t seed, int64_t stripe_elems,
                    int world, int rank) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec4; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e0 = v * 4;
        if (stripe_elems > 0 && (int)((e0 / stripe_elems) % world) != rank) continue;
        const uint4 r = philox(seed, (uint64_t)v, 0u, STREAM_INIT);
        const float x[4] = {(2.0f * u01(r.x) - 1.0f) * bound, (2.0f * u01(r.y) - 1.0f) * bound,
                            (2.0f * u01(r.z) - 1.0f) * bound, (2.0f * u01(r.w) - 1.0f) * bound};
        if (e0 + 4 <= n_elems) {
            *reinterpret_cast<float4 *>(w + e0) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
            for (int j = 0; j < 4 && e0 + j < n_elems; ++j) w[e0 + j] = x[j];
        }
    }
}

// rows of a (possibly sharded) table <-> a dense local buffer; one warp per row, L2-only accesses (rows are shared)
template <bool GATHER>
__global__ void __launch_bounds__(256)
rows_copy_kernel(float *__restrict__ w, int emb, const int64_t *__restrict__ rows, int64_t n, float *__restrict__ dense) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += n_warps) {
        float *trow = w + __ldg(rows + i) * emb;
        float *drow = dense + i * emb;
        for (int e = lane; e < emb; e += 32) {
            if (GATHER) drow[e] = __ldcg(trow + e);
            else __stcg(trow + e, drow[e]);
        }
    }
}

// nn.Embedding(max_norm=...) semantics (torch embedding_renorm_, p = 2): every listed row whose L2 norm exceeds max_norm is scaled IN
// PLACE by max_norm / (norm + 1e-7).  One warp per listed row; the list must not contain a row twice.
__global__ void __launch_bounds__(256)
renorm_rows_kernel(float *__restrict__ w, int emb, const int64_t *__restrict__ rows, int64_t n, float max_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += n_warps) {
        float *row = w + __ldg(rows + i) * emb;
        float s = 0.f;
        for (int e = lane; e < emb; e += 32) { const float v = __ldcg(row + e); s = fmaf(v, v, s); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
        const float norm = sqrtf(s);
        if (norm > max_norm) {
            const float scale = max_norm / (norm + 1e-7f);
            for (int e = lane; e < emb; e += 32) __stcg(row + e, __ldcg(row + e) * scale);
        }
    }
}

}  // namespace
}  // namespace se

extern "C" int se_table_renorm_rows(float *w, int emb, const int64_t *rows, int64_t n, float max_norm, void *stream) {
    SE_REQUIRE(w && emb >= 1 && n >= 0 && (n == 0 || rows) && max_norm > 0.f, "se_table_renorm_rows: bad arguments");
    if (n == 0) return SE_OK;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::renorm_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, emb, rows, n, max_norm);
    return se::check_cuda(cudaGetLastError(), "renorm_rows_kernel launch");
}

extern "C" int se_shard_granularity(int64_t *bytes) {
    SE_REQUIRE(bytes, "se_shard_granularity: null pointer");
    se::DriverApi d;
    int rc = se::load_driver(d);
    if (rc != SE_OK) return rc;
    CUmemAllocationProp prop;
    rc = se::stripe_prop(prop);
    if (rc != SE_OK) return rc;
    size_t g = 0;
    SE_CU(d, GetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    *bytes = (int64_t)g;
    return SE_OK;
}

extern "C" int se_shard_reserve(int64_t bytes, uint64_t *va) {
    SE_REQUIRE(va && bytes > 0, "se_shard_reserve: bad arguments");
    se::DriverApi d;
    int rc = se::load_driver(d);
    if (rc != SE_OK) return rc;
    CUdeviceptr p = 0;
    SE_CU(d, AddressReserve(&p, (size_t)bytes, 0, 0, 0));
    *va = (uint64_t)p;
    return SE_OK;
}

extern "C" int se_shard_unreserve(uint64_t va, int64_t bytes) {
    se::DriverApi d;
    int rc = se::load_driver(d);
    if (rc != SE_OK) return rc;
    SE_CU(d, AddressFree((CUdeviceptr)va, (size_t)bytes));
    return SE_OK;
}

extern "C" int se_shard_create(int64_t bytes, uint64_t *handle) {
    SE_REQUIRE(handle && bytes > 0, "se_shard_create: bad arguments");
    se::DriverApi d;
    int rc = se::load_driver(d);
    if (rc != SE_OK) return rc;
    CUmemAllocationProp prop;
    rc = se::stripe_prop(prop);
    if (rc != SE_OK) return rc;
    CUmemGenericAllocationHandle h = 0;
    SE_CU(d, Create(&h, (size_t)bytes, &prop, 0));
    *handle = (uint64_t)h;
    return SE_OK;
}

extern "C" int se_shard_release(uint64_t handle) {
    se::DriverApi d;
    int rc = se::load_driver(d);
    if (rc != SE_OK) return rc;
    SE_CU(d, Release((CUmemGenericAllocationHandle)handle));
    return SE_OK;
}

extern "C" int se_shard_export_fd(uint64_t handle, int *fd) {
    SE_REQUIRE(fd, "se_shard_export_fd: null pointer");
    se::DriverApi d;
    int rc = se::load_driver(d);
    if (rc != SE_OK) return rc;
    int out = -1;
    SE_CU(d, ExportToShareableHandle(&out, (CUmemGenericAllocationHandle)handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *fd = out;
    return SE_OK;
}

extern "C" int se_shard_import_fd(int fd, uint64_t *handle) {
    SE_REQUIRE(handle && fd >= 0, "se_shard_import_fd: bad arguments");
    se::DriverApi d;
    int rc = se::load_driver(d);
    if (rc != SE_OK) return rc;
    CUmemGenericAllocationHandle h = 0;
    SE_CU(d, ImportFromShareableHandle(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    *handle = (uint64_t)h;
    return SE_OK;
}

extern "C" int se_shard_map(uint64_t va, int64_t bytes, uint64_t handle) {
    SE_REQUIRE(va && bytes > 0, "se_shard_map: bad arguments");
    se::DriverApi d;
    int rc = se::load_driver(d);
    if (rc != SE_OK) return rc;
    int dev = 0;
    SE_CUDA(cudaGetDevice(&dev));
    SE_CU(d, Map((CUdeviceptr)va, (size_t)bytes, 0, (CUmemGenericAllocationHandle)handle, 0));
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    SE_CU(d, SetAccess((CUdeviceptr)va, (size_t)bytes, &acc, 1));
    return SE_OK;
}

extern "C" int se_shard_unmap(uint64_t va, int64_t bytes) {
    se::DriverApi d;
    int rc = se::load_driver(d);
    if (rc != SE_OK) return rc;
    SE_CU(d, Unmap((CUdeviceptr)va, (size_t)bytes));
    return SE_OK;
}

extern "C" int se_table_fill_uniform(float *w, int64_t n_elems, float bound, uint64_t seed, int64_t stripe_elems,
                                     int world, int rank, void *stream) {
    SE_REQUIRE(w && n_elems >= 0, "se_table_fill_uniform: bad arguments");
    SE_REQUIRE(((uintptr_t)w % 16) == 0, "se_table_fill_uniform: table must be 16-byte aligned");
    SE_REQUIRE(stripe_elems == 0 || (stripe_elems % 4 == 0 && world >= 1 && rank >= 0 && rank < world),
               "se_table_fill_uniform: bad shard spec");
    if (n_elems == 0) return SE_OK;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    const int64_t n_vec4 = (n_elems + 3) / 4;
    int64_t blocks = (n_vec4 + 255) / 256;
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    se::fill_uniform_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, n_vec4, n_elems, bound, seed, stripe_elems, world, rank);
    return se::check_cuda(cudaGetLastError(), "fill_uniform_kernel launch");
}

extern "C" int se_table_gather_rows(const float *w, int emb, const int64_t *rows, int64_t n, float *out, void *stream) {
    SE_REQUIRE(w && emb >= 1 && n >= 0 && (n == 0 || (rows && out)), "se_table_gather_rows: bad arguments");
    if (n == 0) return SE_OK;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::rows_copy_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(const_cast<float *>(w), emb, rows, n, out);
    return se::check_cuda(cudaGetLastError(), "rows_copy_kernel<gather> launch");
}

extern "C" int se_table_scatter_rows(float *w, int emb, const int64_t *rows, int64_t n, const float *src, void *stream) {
    SE_REQUIRE(w && emb >= 1 && n >= 0 && (n == 0 || (rows && src)), "se_table_scatter_rows: bad arguments");
    if (n == 0) return SE_OK;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::rows_copy_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, emb, rows, n, const_cast<float *>(src));
    return se::check_cuda(cudaGetLastError(), "rows_copy_kernel<scatter> launch");
}

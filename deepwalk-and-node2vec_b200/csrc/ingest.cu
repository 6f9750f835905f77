// Graph ingest on the device: edge list -> CSR (the step right before the hot path).
//
// The reference builds a networkx graph edge by edge (graph/datasets.py:126-221: `add_edge` per line of cora.cites, relabel_nodes)
// and then re-reads its adjacency dicts on every walk step.  Here an edge list that is already in HBM (synthetic generators, a parsed
// edge file) becomes the CSR the walk kernels read without leaving the device, with networkx's semantics for a simple undirected graph:
//   * `symmetrize`: every input edge (u, v) is stored in both rows (nx.Graph); self loops are dropped (the kernels never see them)
//   * duplicates collapse to ONE entry whose weight is that of the LAST occurrence in input order (a repeated `add_edge` overwrites
//     the edge attributes)
//   * rows come out ascending, so the same array serves as CDF order and as sorted membership order (`col` == `col_sorted`)
//   * optional edge weights -> fp64 weights aligned with the columns + fp32 per-row inclusive prefix sums (`wcdf`) for the sampler
//
// Pipeline (every stage a kernel of this file, no library sort / unique / scan):
//   count degrees (atomics) -> exclusive scan -> bucket fill (atomic cursor per row) -> per-row bitonic sort of (column, input index)
//   in shared memory (global memory for rows longer than 2048) + duplicate marking + unique count -> exclusive scan ->
//   per-row compaction with the weights of the surviving occurrences and their prefix sums.
#include "common.cuh"
#include "scan.cuh"

namespace se {
namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_SMEM_MAX = 2048;          // pairs sorted in shared memory (16 KB); longer rows sort in place in global memory

__global__ void __launch_bounds__(256)
set_total_kernel(const int64_t *__restrict__ counts, const int64_t *__restrict__ excl, int64_t n, int64_t *__restrict__ out_last) {
    if (blockIdx.x == 0 && threadIdx.x == 0) out_last[0] = n > 0 ? excl[n - 1] + counts[n - 1] : 0;
}

// ---- stage 1: degrees ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
count_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t n_edges, int64_t n_nodes, int symmetrize,
             int64_t *__restrict__ deg, int32_t *__restrict__ bad) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = src[e], v = dst[e];
        if (u < 0 || v < 0 || u >= n_nodes || v >= n_nodes) { atomicAdd(bad, 1); continue; }
        if (u == v) continue;
        atomicAdd(reinterpret_cast<unsigned long long *>(deg + u), 1ull);
        if (symmetrize) atomicAdd(reinterpret_cast<unsigned long long *>(deg + v), 1ull);
    }
}

// ---- stage 2: bucket fill -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fill_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t n_edges, int64_t n_nodes, int symmetrize,
            const int64_t *__restrict__ rowptr, int64_t *__restrict__ cursor, int2 *__restrict__ pairs) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = src[e], v = dst[e];
        if (u < 0 || v < 0 || u >= n_nodes || v >= n_nodes || u == v) continue;
        const int64_t pu = rowptr[u] + (int64_t)atomicAdd(reinterpret_cast<unsigned long long *>(cursor + u), 1ull);
        pairs[pu] = make_int2((int)v, (int)e);                        // (column, input index): the index orders duplicates
        if (symmetrize) {
            const int64_t pv = rowptr[v] + (int64_t)atomicAdd(reinterpret_cast<unsigned long long *>(cursor + v), 1ull);
            pairs[pv] = make_int2((int)u, (int)e);
        }
    }
}

__device__ __forceinline__ bool pair_less(const int2 &a, const int2 &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }

// ---- stage 3: per-row sort by (column, input index), duplicate marking, unique count -------------------------------------------
// One block per row (grid-stride).  Bitonic network in its direction-free form (first step of every merge compares i with its mirror
// i ^ (k - 1), the rest are half-cleaners i ^ j; every comparator leaves the smaller element at the lower index), which needs no
// padding to a power of two: a partner index >= n stands for +infinity and is simply skipped.  After the sort an entry survives when
// it is the LAST of its run of equal columns (networkx: a repeated add_edge overwrites the attributes); the survivor flag is the sign
// bit of the index field (fewer than 2^31 input edges).
__global__ void __launch_bounds__(SORT_THREADS)
sort_rows_kernel(const int64_t *__restrict__ rowptr, int64_t n_nodes, int2 *__restrict__ pairs, int64_t *__restrict__ uniq) {
    __shared__ int2 sh[SORT_SMEM_MAX];
    __shared__ int cnt_sh;
    for (int64_t row = blockIdx.x; row < n_nodes; row += gridDim.x) {
        const int64_t lo = rowptr[row];
        const int n = (int)(rowptr[row + 1] - lo);
        int2 *g = pairs + lo;
        const bool in_smem = n <= SORT_SMEM_MAX;
        int2 *a = in_smem ? sh : g;
        if (threadIdx.x == 0) cnt_sh = 0;
        if (in_smem) for (int i = threadIdx.x; i < n; i += SORT_THREADS) sh[i] = g[i];
        __syncthreads();
        int p2 = 1; while (p2 < n) p2 <<= 1;
        for (int k = 2; k <= p2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int mask = (j == (k >> 1)) ? (k - 1) : j;
                for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
                    const int l = i ^ mask;
                    if (l > i && l < n) {
                        const int2 x = a[i], y = a[l];
                        if (pair_less(y, x)) { a[i] = y; a[l] = x; }
                    }
                }
                __syncthreads();
            }
        }
        int mine = 0;
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
            const int2 x = a[i];
            const bool keep = (i == n - 1) || (a[i + 1].x != x.x);             // neighbours only read .x, which this sweep never changes
            mine += keep;
            g[i] = make_int2(x.x, keep ? (x.y | (int)0x80000000) : x.y);
        }
        if (!in_smem) __syncthreads();
        atomicAdd(&cnt_sh, mine);
        __syncthreads();
        if (threadIdx.x == 0) uniq[row] = cnt_sh;
        __syncthreads();
    }
}

// ---- stage 4: compaction + weights + per-row prefix sums -------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
compact_rows_kernel(const int64_t *__restrict__ rowptr_in, const int64_t *__restrict__ rowptr_out, int64_t n_nodes, const int2 *__restrict__ pairs,
                    const double *__restrict__ w_in, int32_t *__restrict__ col_out, double *__restrict__ w_out, float *__restrict__ wcdf_out) {
    __shared__ int warp_cnt[SORT_THREADS / 32];
    __shared__ double warp_w[SORT_THREADS / 32];
    __shared__ int base_cnt;
    __shared__ double base_w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t row = blockIdx.x; row < n_nodes; row += gridDim.x) {
        const int64_t lo = rowptr_in[row], out0 = rowptr_out[row];
        const int n = (int)(rowptr_in[row + 1] - lo);
        if (threadIdx.x == 0) { base_cnt = 0; base_w = 0.0; }
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += SORT_THREADS) {
            const int i = i0 + threadIdx.x;
            int2 a = make_int2(0, 0);
            bool keep = false;
            if (i < n) { a = pairs[lo + i]; keep = a.y < 0; }
            const int e = a.y & 0x7fffffff;
            const double wv = (keep && w_in) ? w_in[e] : 0.0;
            // block-wide inclusive scan of (keep, weight)
            int c = keep; double s = wv;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int tc = __shfl_up_sync(FULL, c, off); const double ts = __shfl_up_sync(FULL, s, off);
                if (lane >= off) { c += tc; s += ts; }
            }
            if (lane == 31) { warp_cnt[warp] = c; warp_w[warp] = s; }
            __syncthreads();
            int pc = base_cnt; double ps = base_w;
            for (int ww = 0; ww < warp; ++ww) { pc += warp_cnt[ww]; ps += warp_w[ww]; }
            if (keep) {
                const int64_t o = out0 + pc + c - 1;
                col_out[o] = a.x;
                if (w_out) w_out[o] = wv;
                if (wcdf_out) wcdf_out[o] = (float)(ps + s);
            }
            __syncthreads();
            if (threadIdx.x == SORT_THREADS - 1) { base_cnt = pc + c; base_w = ps + s; }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256)
max_kernel(const int64_t *__restrict__ x, int64_t n, int64_t *__restrict__ out) {
    int64_t m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, x[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(FULL, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(reinterpret_cast<long long *>(out), (long long)m);
}

}  // namespace
}  // namespace se

// scratch layout (bytes): deg[n] i64 | cursor[n] i64 | rowptr_raw[n + 1] i64 | uniq[n] i64 | scan scratch | pairs[2 n_edges] int2
using se::scan_scratch_elems;

extern "C" int64_t se_csr_build_scratch_bytes(int64_t n_nodes, int64_t n_edges, int symmetrize) {
    if (n_nodes < 0 || n_edges < 0) return -1;
    const int64_t slots = n_edges * (symmetrize ? 2 : 1);
    return 8 * (4 * n_nodes + 1 + scan_scratch_elems(n_nodes)) + 8 * slots + 64;
}

extern "C" int se_csr_build(const int32_t *src, const int32_t *dst, const double *w, int64_t n_edges, int64_t n_nodes, int symmetrize,
                            void *scratch, int64_t scratch_bytes, int64_t *rowptr_out, int32_t *col_out, double *w_out, float *wcdf_out,
                            int64_t *info, void *stream) {
    SE_REQUIRE(n_nodes >= 0 && n_edges >= 0 && n_edges < 0x7fffffffll, "se_csr_build: bad sizes (at most 2^31 - 1 input edges)");
    SE_REQUIRE(rowptr_out && info && (n_edges == 0 || (src && dst && col_out)), "se_csr_build: null pointer");
    SE_REQUIRE((w == nullptr) == (w_out == nullptr) || n_edges == 0, "se_csr_build: pass input and output weights together");
    SE_REQUIRE(wcdf_out == nullptr || w != nullptr, "se_csr_build: prefix sums need weights");
    SE_REQUIRE(scratch && scratch_bytes >= se_csr_build_scratch_bytes(n_nodes, n_edges, symmetrize) && ((uintptr_t)scratch % 16) == 0,
               "se_csr_build: scratch needs %lld bytes, 16-byte aligned", (long long)se_csr_build_scratch_bytes(n_nodes, n_edges, symmetrize));
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t *deg = reinterpret_cast<int64_t *>(scratch), *cursor = deg + n_nodes, *rowptr_raw = cursor + n_nodes, *uniq = rowptr_raw + n_nodes + 1;
    int64_t *scan_scratch = uniq + n_nodes;
    int64_t *after = scan_scratch + scan_scratch_elems(n_nodes);
    int2 *pairs = reinterpret_cast<int2 *>(reinterpret_cast<uintptr_t>(after + 1) & ~(uintptr_t)15);
    // info[0] = nnz, info[1] = max degree, info[2] = invalid endpoints (edges skipped)
    SE_CUDA(cudaMemsetAsync(deg, 0, sizeof(int64_t) * 2 * (size_t)n_nodes, st));
    SE_CUDA(cudaMemsetAsync(info, 0, sizeof(int64_t) * 3, st));
    SE_CUDA(cudaMemsetAsync(rowptr_out, 0, sizeof(int64_t) * (size_t)(n_nodes + 1), st));
    if (n_nodes == 0) return SE_OK;
    int32_t *bad = reinterpret_cast<int32_t *>(info + 2);
    int64_t eb = (n_edges + 255) / 256; if (eb > (int64_t)sms * 16) eb = (int64_t)sms * 16; if (eb < 1) eb = 1;
    se::count_kernel<<<(int)eb, 256, 0, st>>>(src, dst, n_edges, n_nodes, symmetrize, deg, bad);
    int rc = se::exclusive_scan(deg, n_nodes, rowptr_raw, scan_scratch, st);
    if (rc != SE_OK) return rc;
    se::set_total_kernel<<<1, 32, 0, st>>>(deg, rowptr_raw, n_nodes, rowptr_raw + n_nodes);
    se::fill_kernel<<<(int)eb, 256, 0, st>>>(src, dst, n_edges, n_nodes, symmetrize, rowptr_raw, cursor, pairs);
    int64_t rb = n_nodes; if (rb > (int64_t)sms * 8) rb = (int64_t)sms * 8;
    se::sort_rows_kernel<<<(int)rb, se::SORT_THREADS, 0, st>>>(rowptr_raw, n_nodes, pairs, uniq);
    rc = se::exclusive_scan(uniq, n_nodes, rowptr_out, scan_scratch, st);
    if (rc != SE_OK) return rc;
    se::set_total_kernel<<<1, 32, 0, st>>>(uniq, rowptr_out, n_nodes, rowptr_out + n_nodes);
    se::compact_rows_kernel<<<(int)rb, se::SORT_THREADS, 0, st>>>(rowptr_raw, rowptr_out, n_nodes, pairs, w, col_out, w_out, wcdf_out);
    se::set_total_kernel<<<1, 32, 0, st>>>(uniq, rowptr_out, n_nodes, info);
    int64_t mb = (n_nodes + 255) / 256; if (mb > (int64_t)sms * 8) mb = (int64_t)sms * 8;
    se::max_kernel<<<(int)mb, 256, 0, st>>>(uniq, n_nodes, info + 1);
    return se::check_cuda(cudaGetLastError(), "se_csr_build");
}

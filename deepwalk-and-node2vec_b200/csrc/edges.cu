// Link-prediction features on device-resident embeddings (the step right after the hot path).
//
//   edge_features_kernel    out[i, :] = op(table[src[i], :], table[dst[i], :])  -- create_edge_embeddings
//                           (tools/graph_model_downstream_classification.py:203-224) with the four operators of
//                           shallow_encoders/graph/edge_operators.py:10-64.  HBM-bound gather: 3 * 4E bytes per edge.
//   edge_op_kernel          the same operators on two dense [n x E] operands (edge_operator_factory(name)(lhs, rhs)).
//   negative_edges_kernel   sample_negative_edges (:170-200): node uniform over V, partner uniform over the nodes that
//                           are NOT neighbours of it (the node itself counts as a non-neighbour, as in the reference's
//                           set difference) -- rejection sampling over V with a sorted-row membership test, Philox keyed
//                           by (seed; sample id, try).
#include "common.cuh"

namespace se {
namespace {

constexpr uint32_t STREAM_EDGE = 0x60000000u;

__device__ __forceinline__ float apply_op(int op, float a, float b) {
    switch (op) {
        case SE_EDGE_AVERAGE: return (a + b) * 0.5f;          // (lhs + rhs) / 2
        case SE_EDGE_HADAMARD: return a * b;
        case SE_EDGE_WEIGHTED_L1: return fabsf(a - b);
        default: { const float d = a - b; return __fmul_rn(d, d); }   // weighted_l2: (lhs - rhs) ** 2
    }
}

// one warp per edge; float4 path when rows are 16-byte aligned
template <bool VEC4>
__global__ void __launch_bounds__(256)
edge_features_kernel(const float *__restrict__ table, int emb, const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                     int64_t n, int op, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += n_warps) {
        const float *a = table + __ldg(src + i) * emb;
        const float *b = table + __ldg(dst + i) * emb;
        float *o = out + i * emb;
        if (VEC4) {
            for (int e = lane * 4; e < emb; e += 128) {
                const float4 x = __ldg(reinterpret_cast<const float4 *>(a + e));
                const float4 y = __ldg(reinterpret_cast<const float4 *>(b + e));
                __stcs(reinterpret_cast<float4 *>(o + e),
                       make_float4(apply_op(op, x.x, y.x), apply_op(op, x.y, y.y), apply_op(op, x.z, y.z), apply_op(op, x.w, y.w)));
            }
        } else {
            for (int e = lane; e < emb; e += 32) o[e] = apply_op(op, __ldg(a + e), __ldg(b + e));
        }
    }
}

__global__ void __launch_bounds__(256)
edge_op_kernel(const float *__restrict__ lhs, const float *__restrict__ rhs, int64_t n, int op, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = apply_op(op, lhs[i], rhs[i]);
}

__global__ void __launch_bounds__(256)
negative_edges_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, uint32_t n_nodes, int64_t n, uint64_t seed,
                      int64_t sample_id_base, int32_t *__restrict__ out_src, int32_t *__restrict__ out_dst, int32_t *__restrict__ fail_count) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t id = (uint64_t)(sample_id_base + i);
        int32_t node = 0, other = 0;
        bool done = false;
        for (uint32_t attempt = 0; attempt < 4096 && !done; ++attempt) {
            // words x/y of every try: a fresh node (the reference re-draws the node only when it has no non-neighbour at all,
            // which for rejection sampling is the same as never accepting a partner for it) ...
            const uint4 r = philox(seed, id, attempt, STREAM_EDGE);
            if ((attempt & 1023) == 0) node = (int32_t)mulhi32(r.x, n_nodes);      // keep the node for 1024 partner tries
            const int32_t cand = (int32_t)mulhi32(r.y, n_nodes);
            const int64_t lo = __ldg(rowptr + node), hi = __ldg(rowptr + node + 1);
            int64_t l = lo, h = hi;
            while (l < h) {
                const int64_t m = l + ((h - l) >> 1);
                if (__ldg(col + m) < cand) l = m + 1; else h = m;
            }
            if (!(l < hi && __ldg(col + l) == cand)) { other = cand; done = true; }
        }
        if (!done && fail_count) atomicAdd(fail_count, 1);
        out_src[i] = node;
        out_dst[i] = other;
    }
}

int check_op(int op) {
    if (op < SE_EDGE_AVERAGE || op > SE_EDGE_WEIGHTED_L2) { set_error("unknown edge operator %d", op); return SE_ERR_INVALID_ARG; }
    return SE_OK;
}

}  // namespace
}  // namespace se

extern "C" int se_edge_features(const float *table, int64_t vocab, int emb, const int64_t *src_rows, const int64_t *dst_rows,
                                int64_t n_edges, int op, float *out, void *stream) {
    SE_REQUIRE(table && vocab >= 1 && emb >= 1 && n_edges >= 0, "se_edge_features: bad arguments");
    int rc = se::check_op(op);
    if (rc != SE_OK) return rc;
    if (n_edges == 0) return SE_OK;
    SE_REQUIRE(src_rows && dst_rows && out, "se_edge_features: null pointer");
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (n_edges + 7) / 8;
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    const bool vec = (emb % 4 == 0) && ((uintptr_t)table % 16 == 0) && ((uintptr_t)out % 16 == 0);
    if (vec) se::edge_features_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(table, emb, src_rows, dst_rows, n_edges, op, out);
    else se::edge_features_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(table, emb, src_rows, dst_rows, n_edges, op, out);
    return se::check_cuda(cudaGetLastError(), "edge_features_kernel launch");
}

extern "C" int se_edge_op(const float *lhs, const float *rhs, int64_t n_elems, int op, float *out, void *stream) {
    SE_REQUIRE(n_elems >= 0 && (n_elems == 0 || (lhs && rhs && out)), "se_edge_op: bad arguments");
    int rc = se::check_op(op);
    if (rc != SE_OK) return rc;
    if (n_elems == 0) return SE_OK;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (n_elems + 255) / 256;
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    se::edge_op_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(lhs, rhs, n_elems, op, out);
    return se::check_cuda(cudaGetLastError(), "edge_op_kernel launch");
}

extern "C" int se_sample_negative_edges(const int64_t *rowptr, const int32_t *col_sorted, int64_t n_nodes, int64_t n, uint64_t seed,
                                        int64_t sample_id_base, int32_t *out_src, int32_t *out_dst, int32_t *fail_count,
                                        void *stream) {
    SE_REQUIRE(rowptr && n_nodes >= 1 && n_nodes <= 0x7fffffffll && n >= 0, "se_sample_negative_edges: bad arguments");
    if (n == 0) return SE_OK;
    SE_REQUIRE(col_sorted && out_src && out_dst, "se_sample_negative_edges: null pointer");
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::negative_edges_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(rowptr, col_sorted, (uint32_t)n_nodes, n, seed, sample_id_base,
                                                                            out_src, out_dst, fail_count);
    return se::check_cuda(cudaGetLastError(), "negative_edges_kernel launch");
}

// C-ABI plumbing: error string, device info, alias table, negative sampler, host-buffer pipeline step.
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace se {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t err, const char *what) {
    if (err == cudaSuccess) return SE_OK;
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(err));
    return SE_ERR_CUDA;
}

int sm_count() {
    int dev = 0, sms = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice") != SE_OK) return -1;
    if (check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute") != SE_OK) return -1;
    return sms;
}

namespace {
__global__ void __launch_bounds__(256)
sample_negatives_kernel(const float *__restrict__ prob, const int32_t *__restrict__ alias, uint32_t vocab, uint64_t seed,
                        int64_t draw_id_base, int64_t n, int64_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = philox(seed, (uint64_t)(draw_id_base + i), 0u, STREAM_DRAW);
        out[i] = draw_row(prob, alias, vocab, r.x, r.y);
    }
}
__global__ void __launch_bounds__(256)
check_ids_kernel(const int32_t *__restrict__ ids, int64_t n, int64_t lo, int64_t hi, int32_t *__restrict__ bad_count) {
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = ids[i];
        bad += (v < lo) | (v >= hi);
    }
    bad = __reduce_add_sync(FULL, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(bad_count, bad);
}
}  // namespace
}  // namespace se

extern "C" const char *se_version(void) { return "se_b200 0.1.0 (sm_100a)"; }
extern "C" const char *se_last_error(void) { return se::g_err; }

extern "C" int se_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    SE_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SE_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return SE_OK;
}

// Vose's alias method over weights counts^power.
extern "C" int se_alias_build_host(const double *counts, int64_t vocab, double power, float *prob, int32_t *alias) {
    SE_REQUIRE(counts && prob && alias && vocab >= 1 && vocab <= 0x7fffffffll, "se_alias_build_host: bad arguments");
    std::vector<double> scaled((size_t)vocab);
    double total = 0.0;
    for (int64_t i = 0; i < vocab; ++i) {
        SE_REQUIRE(counts[i] >= 0.0, "se_alias_build_host: negative count at %lld", (long long)i);
        const double w = (power == 0.0) ? 1.0 : (counts[i] > 0.0 ? pow(counts[i], power) : 0.0);
        scaled[(size_t)i] = w;
        total += w;
    }
    SE_REQUIRE(total > 0.0, "se_alias_build_host: all weights are zero");
    std::vector<int32_t> small, large;
    small.reserve((size_t)vocab); large.reserve((size_t)vocab);
    for (int64_t i = 0; i < vocab; ++i) {
        scaled[(size_t)i] *= (double)vocab / total;
        (scaled[(size_t)i] < 1.0 ? small : large).push_back((int32_t)i);
    }
    while (!small.empty() && !large.empty()) {
        const int32_t s = small.back(); small.pop_back();
        const int32_t l = large.back(); large.pop_back();
        prob[s] = (float)scaled[(size_t)s];
        alias[s] = l;
        scaled[(size_t)l] = (scaled[(size_t)l] + scaled[(size_t)s]) - 1.0;
        (scaled[(size_t)l] < 1.0 ? small : large).push_back(l);
    }
    for (int32_t i : large) { prob[i] = 1.0f; alias[i] = i; }
    for (int32_t i : small) { prob[i] = 1.0f; alias[i] = i; }
    return SE_OK;
}

extern "C" int se_sample_negatives(const float *prob, const int32_t *alias, int64_t vocab, uint64_t seed,
                                   int64_t draw_id_base, int64_t n, int64_t *out, void *stream) {
    SE_REQUIRE(vocab >= 1 && vocab <= 0x7fffffffll && n >= 0 && (out || n == 0), "se_sample_negatives: bad arguments");
    SE_REQUIRE((prob == nullptr) == (alias == nullptr), "se_sample_negatives: pass both alias arrays or neither");
    if (n == 0) return SE_OK;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::sample_negatives_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(prob, alias, (uint32_t)vocab, seed,
                                                                              draw_id_base, n, out);
    return se::check_cuda(cudaGetLastError(), "sample_negatives_kernel launch");
}

extern "C" int se_host_walk_sgns_step_sharded(const int64_t *rowptr, const int32_t *col, const float *wcdf, int64_t n_nodes,
                                              int symmetric, const int32_t *starts_host, int64_t n_walks, int walk_len,
                                              double p, double q, int node2vec, int rule, uint64_t seed, int64_t walk_id_base,
                                              float *w_in, float *w_out, int64_t vocab, int emb, int radius, int n_neg,
                                              int row_offset, const float *alias_prob, const int32_t *alias_idx, float lr,
                                              int flags, const se_shard_spec *spec, int32_t *starts_dev, int32_t *walks_dev,
                                              double *stats_dev, int32_t *walks_host, double *stats_host, void *stream) {
    SE_REQUIRE(starts_host && starts_dev && walks_dev && stats_dev && stats_host, "se_host_walk_sgns_step: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    SE_CUDA(cudaMemcpyAsync(starts_dev, starts_host, sizeof(int32_t) * (size_t)n_walks, cudaMemcpyHostToDevice, st));
    SE_CUDA(cudaMemsetAsync(stats_dev, 0, sizeof(double) * SE_STATS_LEN, st));
    int rc = se_walk(rowptr, col, wcdf, n_nodes, symmetric, starts_dev, n_walks, walk_len, p, q, node2vec, rule, seed,
                     walk_id_base, 1, walks_dev, nullptr, SE_WALK_AUTO, stream);
    if (rc != SE_OK) return rc;
    rc = se_sgns_update_walks_sharded(w_in, w_out, vocab, emb, walks_dev, n_walks, walk_len, radius, n_neg, row_offset,
                                      alias_prob, alias_idx, lr, seed ^ 0x9E3779B97F4A7C15ull,
                                      walk_id_base * (int64_t)(walk_len - 2 * radius), flags, spec, stats_dev, stream);
    if (rc != SE_OK) return rc;
    if (walks_host)
        SE_CUDA(cudaMemcpyAsync(walks_host, walks_dev, sizeof(int32_t) * (size_t)n_walks * (size_t)walk_len,
                                cudaMemcpyDeviceToHost, st));
    SE_CUDA(cudaMemcpyAsync(stats_host, stats_dev, sizeof(double) * SE_STATS_LEN, cudaMemcpyDeviceToHost, st));
    SE_CUDA(cudaStreamSynchronize(st));
    return SE_OK;
}

extern "C" int se_host_walk_sgns_step(const int64_t *rowptr, const int32_t *col, const float *wcdf, int64_t n_nodes,
                                      int symmetric, const int32_t *starts_host, int64_t n_walks, int walk_len,
                                      double p, double q, int node2vec, int rule, uint64_t seed, int64_t walk_id_base,
                                      float *w_in, float *w_out, int64_t vocab, int emb, int radius, int n_neg,
                                      int row_offset, const float *alias_prob, const int32_t *alias_idx, float lr,
                                      int flags, int32_t *starts_dev, int32_t *walks_dev, double *stats_dev,
                                      int32_t *walks_host, double *stats_host, void *stream) {
    return se_host_walk_sgns_step_sharded(rowptr, col, wcdf, n_nodes, symmetric, starts_host, n_walks, walk_len, p, q,
                                          node2vec, rule, seed, walk_id_base, w_in, w_out, vocab, emb, radius, n_neg,
                                          row_offset, alias_prob, alias_idx, lr, flags, nullptr, starts_dev, walks_dev,
                                          stats_dev, walks_host, stats_host, stream);
}

extern "C" int se_host_sgns_update_tokens(const int32_t *tokens_host, int64_t n_seq, int seq_len, float *w_in, float *w_out,
                                          int64_t vocab, int emb, int radius, int n_neg, int row_offset,
                                          const float *alias_prob, const int32_t *alias_idx, float lr, uint64_t seed,
                                          int64_t centre_id_base, int flags, const se_shard_spec *spec, int32_t *tokens_dev,
                                          double *stats_dev, double *stats_host, void *stream) {
    SE_REQUIRE(tokens_host && tokens_dev && stats_dev && stats_host && n_seq >= 0 && seq_len >= 1, "se_host_sgns_update_tokens: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SE_CUDA(cudaMemcpyAsync(tokens_dev, tokens_host, sizeof(int32_t) * (size_t)n_seq * (size_t)seq_len, cudaMemcpyHostToDevice, st));
    SE_CUDA(cudaMemsetAsync(stats_dev, 0, sizeof(double) * SE_STATS_LEN, st));
    int rc = se_sgns_update_walks_sharded(w_in, w_out, vocab, emb, tokens_dev, n_seq, seq_len, radius, n_neg, row_offset, alias_prob,
                                          alias_idx, lr, seed, centre_id_base, flags, spec, stats_dev, stream);
    if (rc != SE_OK) return rc;
    SE_CUDA(cudaMemcpyAsync(stats_host, stats_dev, sizeof(double) * SE_STATS_LEN, cudaMemcpyDeviceToHost, st));
    SE_CUDA(cudaStreamSynchronize(st));
    return SE_OK;
}

extern "C" int se_check_ids(const int32_t *ids, int64_t n, int64_t lo, int64_t hi, int32_t *bad_count, void *stream) {
    SE_REQUIRE(n >= 0 && bad_count && (ids || n == 0), "se_check_ids: bad arguments");
    if (n == 0) return SE_OK;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::check_ids_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(ids, n, lo, hi, bad_count);
    return se::check_cuda(cudaGetLastError(), "check_ids_kernel launch");
}

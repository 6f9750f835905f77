// CBOW scoring, loss and gradients (the reference's other model: shallow_encoders/word2vec/model.py:94-110; collate mode `cbow`,
// word2vec/dataloader/torch_dataset.py:310-314: inputs = the 2r context ids, targets = the centre id).
//
//   h_b          = mean_n W_in[inputs[b, n]]                                     (model.py:103)
//   scores[b, j] = <h_b, W_out[outputs[b, j]]>                                   (model.py:104-106; sigmoid when proba)
//
// One warp per batch row: the mean vector stays in registers (emb <= 1024), every target row is read once with coalesced
// loads, dots are reduced with shuffles.  The gradient kernel evaluates the reference's training step for CBOW
// (trainer.py:131-139 + loss.py:14-22: positive logits (B, M), negative logits (B, M, K), MEAN over B * M) and accumulates the
// dense gradients autograd would produce: dW_out[o] += g * h_b, dW_in[inputs[b, n]] += (1 / N) * sum_j g_j * W_out[o_j].
#include "sgns_common.cuh"

namespace se {
namespace {

constexpr int CBOW_R = 32;      // floats per lane: emb <= 1024

template <bool GRAD>
__global__ void __launch_bounds__(256)
cbow_kernel(const float *__restrict__ w_in, const float *__restrict__ w_out, int emb, const int64_t *__restrict__ inputs,
            const int64_t *__restrict__ outputs, const int64_t *__restrict__ noise, int64_t batch, int n_in, int m, int k, int proba,
            float *__restrict__ scores, double *__restrict__ stats, float *__restrict__ grad_in, float *__restrict__ grad_out, float grad_scale) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nr = (emb + 31) / 32;
    const float inv_n = 1.0f / (float)n_in;
    float loss_pos = 0.f, loss_neg = 0.f;
    unsigned cnt_recall = 0, cnt_fp = 0, cnt_pairs = 0;
    for (int64_t b = warp; b < batch; b += n_warps) {
        float h[CBOW_R], gh[CBOW_R];
#pragma unroll
        for (int r = 0; r < CBOW_R; ++r) { h[r] = 0.f; gh[r] = 0.f; }
        for (int n = 0; n < n_in; ++n) {
            const float *row = w_in + __ldg(inputs + b * n_in + n) * emb;
#pragma unroll
            for (int r = 0; r < CBOW_R; ++r) { const int e = lane + 32 * r; if (r < nr && e < emb) h[r] += __ldcg(row + e); }
        }
#pragma unroll
        for (int r = 0; r < CBOW_R; ++r) h[r] *= inv_n;                                  // torch.mean(..., dim=1)
        const int per_pos = GRAD ? 1 + k : 1;
        for (int j = 0; j < m; ++j) {
            for (int t = 0; t < per_pos; ++t) {
                const int64_t orow = (t == 0) ? __ldg(outputs + b * m + j) : __ldg(noise + (b * m + j) * k + (t - 1));
                const float *row = w_out + orow * emb;
                float o[CBOW_R], d = 0.f;
#pragma unroll
                for (int r = 0; r < CBOW_R; ++r) {
                    const int e = lane + 32 * r;
                    o[r] = (r < nr && e < emb) ? __ldcg(row + e) : 0.f;
                    d = fmaf(o[r], h[r], d);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(FULL, d, off);
                if constexpr (!GRAD) {
                    if (lane == 0) scores[b * m + j] = proba ? 1.0f / (1.0f + expf(-d)) : d;
                } else {
                    float g;
                    if (t == 0) {
                        const float sig = 1.0f / (1.0f + expf(-d));
                        if (lane == 0) { loss_pos -= logf(fmaxf(sig, CLAMP_MIN)); cnt_recall += sig >= 0.5f; cnt_pairs += 1; }
                        g = (sig > CLAMP_MIN) ? -1.0f / (1.0f + expf(d)) : 0.f;
                    } else {
                        const float sig_m = 1.0f / (1.0f + expf(d)), sig = 1.0f / (1.0f + expf(-d));
                        if (lane == 0) { loss_neg -= logf(fmaxf(sig_m, CLAMP_MIN)); cnt_fp += sig >= 0.5f; }
                        g = (sig_m > CLAMP_MIN) ? sig : 0.f;
                    }
                    g *= grad_scale;
#pragma unroll
                    for (int r = 0; r < CBOW_R; ++r) {
                        const int e = lane + 32 * r;
                        if (r < nr && e < emb) {
                            gh[r] = fmaf(g, o[r], gh[r]);
                            if (grad_out) atomicAdd(grad_out + orow * emb + e, g * h[r]);
                        }
                    }
                }
            }
        }
        if constexpr (GRAD) {
            if (grad_in) {
                for (int n = 0; n < n_in; ++n) {
                    float *row = grad_in + __ldg(inputs + b * n_in + n) * emb;
#pragma unroll
                    for (int r = 0; r < CBOW_R; ++r) { const int e = lane + 32 * r; if (r < nr && e < emb) atomicAdd(row + e, gh[r] * inv_n); }
                }
            }
        }
    }
    if constexpr (GRAD) flush_stats(stats, lane == 0 && cnt_pairs != 0, loss_pos, loss_neg, cnt_recall, cnt_fp, cnt_pairs, (double)cnt_pairs * (double)k);
}

int cbow_blocks(int64_t batch) {
    const int sms = sm_count();
    if (sms <= 0) return 0;
    int64_t blocks = (batch + 7) / 8;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace se

extern "C" int se_cbow_scores(const float *w_in, const float *w_out, int64_t vocab, int emb, const int64_t *inputs, const int64_t *outputs,
                              int64_t batch, int n_in, int m, int proba, float *out, void *stream) {
    SE_REQUIRE(w_in && w_out && vocab >= 1 && emb >= 1 && emb <= 32 * se::CBOW_R, "se_cbow_scores: bad tables (1 <= emb <= %d)", 32 * se::CBOW_R);
    SE_REQUIRE(batch >= 0 && n_in >= 1 && m >= 1, "se_cbow_scores: bad shape");
    if (batch == 0) return SE_OK;
    SE_REQUIRE(inputs && outputs && out, "se_cbow_scores: null pointer");
    const int blocks = se::cbow_blocks(batch);
    if (blocks <= 0) return SE_ERR_CUDA;
    se::cbow_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(w_in, w_out, emb, inputs, outputs, nullptr, batch, n_in, m, 0, proba, out, nullptr,
                                                                     nullptr, nullptr, 0.f);
    return se::check_cuda(cudaGetLastError(), "cbow_kernel<scores> launch");
}

extern "C" int se_cbow_grad(const float *w_in, const float *w_out, int64_t vocab, int emb, const int64_t *inputs, const int64_t *targets,
                            const int64_t *noise, int64_t batch, int n_in, int m, int n_neg, double *stats, float *grad_in, float *grad_out,
                            void *stream) {
    SE_REQUIRE(w_in && w_out && vocab >= 1 && emb >= 1 && emb <= 32 * se::CBOW_R, "se_cbow_grad: bad tables (1 <= emb <= %d)", 32 * se::CBOW_R);
    SE_REQUIRE(batch >= 0 && n_in >= 1 && m >= 1 && n_neg >= 0, "se_cbow_grad: bad shape");
    if (batch == 0) return SE_OK;
    SE_REQUIRE(inputs && targets && (noise || n_neg == 0) && stats, "se_cbow_grad: null pointer");
    SE_REQUIRE((grad_in == nullptr) == (grad_out == nullptr), "se_cbow_grad: pass both gradient buffers or neither");
    const int blocks = se::cbow_blocks(batch);
    if (blocks <= 0) return SE_ERR_CUDA;
    se::cbow_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(w_in, w_out, emb, inputs, targets, noise, batch, n_in, m, n_neg, 0, nullptr, stats,
                                                                    grad_in, grad_out, 1.0f / (float)(batch * m));
    return se::check_cuda(cudaGetLastError(), "cbow_kernel<grad> launch");
}

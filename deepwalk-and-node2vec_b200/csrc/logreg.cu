// Softmax / logistic cross-entropy for the device-side downstream classifier (tools/graph_model_downstream_classification.py:85-91 fits
// sklearn's LogisticRegression on the embedding rows; at 10 M x 128 that host fit is the bottleneck SURVEY 8f rank 2 names).
// The two GEMMs of a gradient evaluation (logits = X W^T, dW = G^T X) run on the tensor cores (csrc/gemm.cu); this kernel is the
// elementwise part between them: per sample, numerically stable log-softmax (or the binary logistic form for ONE logit column),
// loss summed into *loss_sum, and G = d loss / d logits written in place of the logits.
#include "common.cuh"

namespace se {
namespace {

__global__ void __launch_bounds__(256)
xent_kernel(float *__restrict__ logits, const int32_t *__restrict__ labels, const float *__restrict__ bias, int64_t n, int n_cols, float scale,
            double *__restrict__ loss_sum, int32_t *__restrict__ n_correct) {
    float loss = 0.f; int correct = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float *z = logits + i * n_cols;
        const int y = labels[i];
        if (n_cols == 1) {                                   // binary: one logit, P(y = 1) = sigmoid(z)
            const float s = z[0] + (bias ? bias[0] : 0.f);
            const float p = 1.0f / (1.0f + expf(-s));
            loss += (y == 1) ? (s > 0.f ? log1pf(expf(-s)) : -s + log1pf(expf(s))) : (s > 0.f ? s + log1pf(expf(-s)) : log1pf(expf(s)));
            correct += ((s > 0.f) ? 1 : 0) == y;
            z[0] = (p - (float)y) * scale;
        } else {
            float m = -INFINITY; int arg = 0;
            for (int c = 0; c < n_cols; ++c) { const float v = z[c] + (bias ? bias[c] : 0.f); if (v > m) { m = v; arg = c; } }
            float den = 0.f;
            for (int c = 0; c < n_cols; ++c) den += expf(z[c] + (bias ? bias[c] : 0.f) - m);
            const float lse = m + logf(den);
            loss += lse - (z[y] + (bias ? bias[y] : 0.f));
            correct += arg == y;
            for (int c = 0; c < n_cols; ++c) z[c] = (expf(z[c] + (bias ? bias[c] : 0.f) - lse) - (c == y ? 1.f : 0.f)) * scale;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { loss += __shfl_xor_sync(FULL, loss, off); correct += __shfl_xor_sync(FULL, correct, off); }
    if ((threadIdx.x & 31) == 0) {
        if (loss_sum && loss != 0.f) atomicAdd(loss_sum, (double)loss);
        if (n_correct && correct) atomicAdd(n_correct, correct);
    }
}

}  // namespace
}  // namespace se

extern "C" int se_softmax_xent(float *logits, const int32_t *labels, const float *bias, int64_t n, int n_cols, float grad_scale,
                               double *loss_sum, int32_t *n_correct, void *stream) {
    SE_REQUIRE(n >= 0 && n_cols >= 1, "se_softmax_xent: bad shape");
    if (n == 0) return SE_OK;
    SE_REQUIRE(logits && labels, "se_softmax_xent: null pointer");
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (n + 255) / 256; if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    se::xent_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, labels, bias, n, n_cols, grad_scale, loss_sum, n_correct);
    return se::check_cuda(cudaGetLastError(), "xent_kernel launch");
}

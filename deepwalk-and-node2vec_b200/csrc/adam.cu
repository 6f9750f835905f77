// Row-sparse Adam for the two embedding tables (the optimiser every shipped YAML names: `_target_: torch.optim.Adam`,
// configs/sge_sg_karate_club.yaml:32-34, instantiated at config_parser/core.py:43-94).
//
// torch.optim.Adam on the reference's dense nn.Embedding tables reads and writes theta, m, v of ALL 2 V E parameters every
// step (1.9 GB per step at V = 267 k, E = 128, whatever the batch).  Here only the rows that received a gradient in the step
// are updated ("lazy" Adam): the gradient kernel (sgns.cu, MODE_GRAD) accumulates the mean-loss gradient of the batch into
// persistent [V x E] accumulators, flags every row it touches and appends it to a list; this kernel then runs one warp per listed
// row: m, v, theta update with torch's formulas (lerp for m, addcmul for v, sqrt(v) / sqrt(1 - beta2^t) + eps, step lr / (1 - beta1^t)),
// bias correction by the ROW's own step count t, and puts the accumulator row, the flag and the list length back to zero.
// On a batch that touches the same rows every step (or on the first step) this equals dense torch.optim.Adam exactly;
// rows that are not touched keep their value (dense Adam would keep moving them along their decaying momentum).
//
// Algorithmic bytes per touched row and step: gradient accumulate (read + write) + g, m, v, theta read + m, v, theta, g written
// = 10 row transfers of 4 E bytes, against 2 for in-place SGD.
#include <math.h>

#include "sgns_common.cuh"

namespace se {
namespace {

__global__ void __launch_bounds__(256)
adam_apply_kernel(float *__restrict__ w, float *__restrict__ m, float *__restrict__ v, float *__restrict__ g, int32_t *__restrict__ t,
                  int32_t *__restrict__ flags, const int32_t *__restrict__ list, int32_t *count, int32_t *done, int emb, float lr, float beta1,
                  float beta2, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n = *reinterpret_cast<volatile int32_t *>(count);
    for (int64_t i = warp; i < n; i += n_warps) {
        const int64_t row = list[i];
        int tt = 0;
        if (lane == 0) { tt = t[row] + 1; t[row] = tt; flags[row] = 0; }
        tt = __shfl_sync(FULL, tt, 0);
        const double bc1 = 1.0 - pow((double)beta1, (double)tt), bc2 = 1.0 - pow((double)beta2, (double)tt);
        const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
        for (int e = lane; e < emb; e += 32) {
            const int64_t k = row * emb + e;
            const float gg = g[k];
            const float mm = m[k] + (gg - m[k]) * (1.0f - beta1);                        // exp_avg.lerp_(grad, 1 - beta1)
            const float vv = fmaf(gg * gg, 1.0f - beta2, v[k] * beta2);                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const float denom = sqrtf(vv) / bc2_sqrt + eps;
            w[k] = w[k] - step_size * (mm / denom);                                     // param.addcdiv_(exp_avg, denom, value=-step_size)
            m[k] = mm; v[k] = vv; g[k] = 0.f;
        }
    }
    // the last block to finish puts the list length back to zero for the next step
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(done, 1) == (int)gridDim.x - 1) { *count = 0; *done = 0; }
    }
}

}  // namespace

int adam_apply(float *w, float *m, float *v, float *g, int32_t *t, int32_t *flags, const int32_t *list, int32_t *count, int32_t *done,
               int64_t capacity, int emb, float lr, float beta1, float beta2, float eps, cudaStream_t stream) {
    const int sms = sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int64_t blocks = (capacity + 7) / 8;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    if (blocks < 1) blocks = 1;
    adam_apply_kernel<<<(int)blocks, 256, 0, stream>>>(w, m, v, g, t, flags, list, count, done, emb, lr, beta1, beta2, eps);
    return check_cuda(cudaGetLastError(), "adam_apply_kernel launch");
}

}  // namespace se

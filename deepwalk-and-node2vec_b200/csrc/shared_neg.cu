// Optional SHARED-NEGATIVES batch mode of the SGNS step (north star, subsystem 3): one set of S negative rows serves every centre of
// the batch, so scoring is ONE dense B x S GEMM on the tensor cores (csrc/gemm.cu, tcgen05) instead of B * N * K warp-shuffle dots,
// and both updates are GEMMs too:
//     C  = W_in[inputs]  (B x E)          Nm = W_out[shared]  (S x E)
//     Sc = C . Nm^T                                                   scores              (word2vec/model.py:88 for all pairs at once)
//     G  = -lr * w * sigmoid(Sc) * [sigmoid(-Sc) > 1e-6]              -lr * dL/dscore     (word2vec/loss.py:16, clamp included)
//     W_in[inputs]  += G   . Nm                                       (B x E)
//     W_out[shared] += G^T . C                                        (S x E)
// with w = n_ctx * n_neg / S: every centre of the reference draws n_ctx * n_neg negatives of its own (utils/sampling.py:21); here each
// of the S shared rows stands for n_ctx * n_neg / S of them, so the negative term has the same expectation.  It is a DIFFERENT
// estimator (negatives are correlated across the batch) and therefore opt-in; the positive pairs of the batch go through
// se_sgns_step with n_neg = 0.  All of a step's scores use the tables as they were at its start (mini-batch semantics).
#include "common.cuh"

namespace se {
namespace {

// in place: score -> -lr * w * dL/dscore; loss / counters of the weighted negative term
__global__ void __launch_bounds__(256)
shared_neg_coeff_kernel(float *__restrict__ sc, int64_t n, float lr_w, float w, double *__restrict__ stats) {
    float loss = 0.f, fp = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float s = sc[i];
        const float sig = 1.0f / (1.0f + expf(-s)), sig_m = 1.0f / (1.0f + expf(s));
        loss -= logf(fmaxf(sig_m, 1e-6f));
        fp += sig >= 0.5f ? 1.f : 0.f;
        sc[i] = (sig_m > 1e-6f) ? -lr_w * sig : 0.f;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { loss += __shfl_xor_sync(FULL, loss, off); fp += __shfl_xor_sync(FULL, fp, off); }
    if ((threadIdx.x & 31) == 0 && stats && (loss != 0.f || fp != 0.f)) {
        atomicAdd(stats + 1, (double)loss * (double)w);
        atomicAdd(stats + 3, (double)fp * (double)w);
    }
}

// w[rows[i], :] += src[i, :]; duplicates in `rows` accumulate
__global__ void __launch_bounds__(256)
rows_add_kernel(float *__restrict__ w, int emb, const int64_t *__restrict__ rows, int64_t n, const float *__restrict__ src) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += n_warps) {
        float *row = w + __ldg(rows + i) * emb;
        for (int e = lane; e < emb; e += 32) atomicAdd(row + e, src[i * emb + e]);
    }
}

}  // namespace
}  // namespace se

extern "C" int64_t se_shared_negatives_scratch_floats(int64_t batch, int64_t n_shared, int emb) {
    if (batch < 0 || n_shared < 0 || emb < 1) return -1;
    // C, C^T, dC (3 B E) + Nm, Nm^T, dN (3 S E) + G, G^T (2 B S), each rounded up to 4 floats
    auto r4 = [](int64_t x) { return (x + 3) / 4 * 4; };
    return 3 * r4(batch * emb) + 3 * r4(n_shared * emb) + 2 * r4(batch * n_shared);
}

extern "C" int se_sgns_step_shared_negatives(float *w_in, float *w_out, int64_t vocab, int emb, const int64_t *inputs, int64_t batch,
                                             const int64_t *shared, int64_t n_shared, int n_ctx, int n_neg, float lr, float *scratch,
                                             int64_t scratch_floats, double *stats, void *stream) {
    SE_REQUIRE(w_in && w_out && vocab >= 1 && emb >= 1, "se_sgns_step_shared_negatives: bad tables");
    SE_REQUIRE(batch >= 0 && n_shared >= 0 && n_ctx >= 1 && n_neg >= 1, "se_sgns_step_shared_negatives: bad shape");
    if (batch == 0 || n_shared == 0) return SE_OK;
    SE_REQUIRE(inputs && shared && scratch, "se_sgns_step_shared_negatives: null pointer");
    SE_REQUIRE(scratch_floats >= se_shared_negatives_scratch_floats(batch, n_shared, emb) && ((uintptr_t)scratch % 16) == 0,
               "se_sgns_step_shared_negatives: scratch needs %lld floats, 16-byte aligned",
               (long long)se_shared_negatives_scratch_floats(batch, n_shared, emb));
    auto r4 = [](int64_t x) { return (x + 3) / 4 * 4; };
    float *c = scratch, *ct = c + r4(batch * emb), *dc = ct + r4(batch * emb);
    float *nm = dc + r4(batch * emb), *nmt = nm + r4(n_shared * emb), *dn = nmt + r4(n_shared * emb);
    float *g = dn + r4(n_shared * emb), *gt = g + r4(batch * n_shared);
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = se::sm_count();
    if (sms <= 0) return SE_ERR_CUDA;
    int rc;
    if ((rc = se_table_gather_rows(w_in, emb, inputs, batch, c, stream)) != SE_OK) return rc;
    if ((rc = se_table_gather_rows(w_out, emb, shared, n_shared, nm, stream)) != SE_OK) return rc;
    if ((rc = se_gemm_nt(c, nm, batch, n_shared, emb, nullptr, nullptr, g, stream)) != SE_OK) return rc;          // scores: tensor cores
    const float w = (float)n_ctx * (float)n_neg / (float)n_shared;
    {
        const int64_t n = batch * n_shared;
        int64_t blocks = (n + 255) / 256; if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
        se::shared_neg_coeff_kernel<<<(int)blocks, 256, 0, st>>>(g, n, lr * w, w, stats);
        SE_CUDA(cudaGetLastError());
    }
    if ((rc = se_transpose(nm, n_shared, emb, nmt, stream)) != SE_OK) return rc;                                   // Nm^T  (E x S)
    if ((rc = se_gemm_nt(g, nmt, batch, emb, (int)n_shared, nullptr, nullptr, dc, stream)) != SE_OK) return rc;    // dC = G . Nm
    if ((rc = se_transpose(g, batch, n_shared, gt, stream)) != SE_OK) return rc;                                   // G^T   (S x B)
    if ((rc = se_transpose(c, batch, emb, ct, stream)) != SE_OK) return rc;                                        // C^T   (E x B)
    SE_REQUIRE(batch <= 0x7fffffffll, "se_sgns_step_shared_negatives: batch too large for one call");
    if ((rc = se_gemm_nt(gt, ct, n_shared, emb, (int)batch, nullptr, nullptr, dn, stream)) != SE_OK) return rc;    // dN = G^T . C
    {
        int64_t blocks = (batch + 7) / 8; if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
        se::rows_add_kernel<<<(int)blocks, 256, 0, st>>>(w_in, emb, inputs, batch, dc);
        SE_CUDA(cudaGetLastError());
        blocks = (n_shared + 7) / 8; if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
        se::rows_add_kernel<<<(int)blocks, 256, 0, st>>>(w_out, emb, shared, n_shared, dn);
        SE_CUDA(cudaGetLastError());
    }
    return SE_OK;
}

/*
 * se_b200.h -- C ABI of the B200-native (sm_100a) random-walk + SGNS hot path.
 *
 * The reference (Robotmurlock/Deepwalk-and-Node2vec) is pure Python and has no FFI of its own; every
 * entry point below names the reference interface (file:line under the reference root) whose arithmetic
 * it replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary
 *   - every `*_dev` / unqualified buffer pointer is a DEVICE pointer owned by the caller; the library never
 *     allocates or frees caller-visible memory, scratch is passed in
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is enqueued on it and the
 *     call returns without synchronising, except the `se_host_*` entry points, which take HOST buffers, copy them
 *     and synchronise before returning
 *   - return value: SE_OK, or a negative SE_ERR_* code; se_last_error() gives a thread-local message.
 *     Nothing throws across the ABI.  There is no CPU fallback: without a usable sm_100 device every compute
 *     entry point returns SE_ERR_CUDA.
 *   - no global mutable state besides the thread-local error string: re-entrant across streams and devices
 *   - node ids are int32 in [0, n_nodes); embedding-table rows are int64 on the explicit-index API (the reference's
 *     LongTensors) and id + row_offset on the walk-driven API (row 0 = '<unk>',
 *     word2vec/dataloader/torch_dataset.py:99-110)
 */
#ifndef SE_B200_H
#define SE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE_OK 0
#define SE_ERR_INVALID_ARG (-1)
#define SE_ERR_CUDA (-2)
#define SE_ERR_UNSUPPORTED (-3)

/* node2vec multiplier rule */
#define SE_RULE_REFERENCE 0 /* what the reference CODE does: x==t -> 1/p ; t in N(x) -> 1/q ; else 1
                               (graph/random_walk_generator.py:101-108) */
#define SE_RULE_PAPER 1     /* node2vec paper / reference README: distance 1 -> 1 ; distance 2 -> 1/q */

/* flags for the SGNS update kernels */
#define SE_SGNS_SCATTER_RED 0   /* red.global.add.v4.f32 scatter: concurrent updates of a row all land */
#define SE_SGNS_SCATTER_STORE 1 /* plain read-modify-write stores: classic racy Hogwild */
#define SE_SGNS_GENERIC_KERNEL 2 /* bit flag: use the generic group-per-centre kernel even where the warp-per-centre
                                    fast kernel applies (same results; for testing and A/B timing) */
#define SE_SGNS_NO_WINDOW 4 /* bit flag (se_sgns_update_walks*): use the per-context kernel, which gathers and scatters the
                               positive row of every pair through L2, instead of the window-resident kernel, which keeps
                               the 2r+1 context rows around the centre in shared memory (for A/B timing) */

#define SE_SGNS_WHOLE_SEQUENCES 8 /* bit flag (se_sgns_update_walks*): the window kernel gives every lane group whole sequences
                                     (a group's span is rounded up to a multiple of L - 2r centres), so all pairs of a sequence are
                                     applied in order by one group.  Always on when there are at least as many sequences as groups */
#define SE_SGNS_WINDOW_REFRESH 16 /* bit flag (se_sgns_update_walks*): the window kernel scatters and re-fetches a token's resident
                                     W_out row when the token is the centre (it is not a context then), so a resident copy is at most r
                                     centres old instead of 2r: for token streams whose frequent tokens sit in many windows at once */
#define SE_SGNS_BATCHED_POSITIVES 32 /* bit flag (se_sgns_update_walks* with n_neg = 0 and 64 < emb <= 128): the 2r positive pairs of a
                                        centre are scored against ONE snapshot of the window (one transposed reduction for all 2r dots,
                                        sigmoids side by side) and then applied in order, instead of pair by pair -- the mini-batch-of-2r
                                        form; 2-3x faster when the negatives run elsewhere (the owner-computes multi-GPU mode sets it) */
/* stats layout written by the SGNS kernels (double[SE_STATS_LEN], ACCUMULATED into, caller zeroes):
 *   [0] sum over pairs of positive loss   -log clamp(sigmoid(s+), 1e-6)          (word2vec/loss.py:15)
 *   [1] sum over pairs of negative loss   -sum_k log clamp(sigmoid(-s-), 1e-6)   (word2vec/loss.py:16)
 *   [2] number of positive pairs with sigmoid(s+) >= 0.5   (recall numerator,     word2vec/trainer.py:145-146)
 *   [3] number of negatives with sigmoid(s-) >= 0.5        (1 - precision numer., word2vec/trainer.py:148-149)
 *   [4] number of positive pairs processed
 *   [5] number of negatives processed
 */
#define SE_STATS_LEN 6

const char *se_version(void);
const char *se_last_error(void);
/* sm count and compute capability of the current device. */
int se_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------------------------------
 * Random walks.  Replaces DeepWalk.walk / Node2Vec.walk (graph/random_walk_generator.py:61-72, 94-119) and the
 * per-node weight helpers (:41-53) over a CSR graph.  A walk has `walk_len` NODES (walk_len-1 transitions, :64,:98).
 * out[n_walks * walk_len] int32 node ids, row-major.
 * ---------------------------------------------------------------------------------------------------------- */

/* Exact mode: fp64 inverse-CDF selection consuming ONE supplied uniform per transition, bit-identical to the
 * reference under `random.choices` (CPython 3.12 sum()/accumulate/bisect semantics; see oracle/walk_oracle.py).
 *   rowptr[n_nodes+1], col[nnz] in the reference's CDF order (networkx adjacency order),
 *   col_sorted[nnz]  each row ascending (may alias col when rows are already sorted) -- membership tests,
 *   w[nnz] fp64 edge weights or NULL (unweighted); w_is_int != 0 when the weights are python ints,
 *   uniforms[n_walks * (walk_len-1)] fp64 in [0,1),
 *   scratch: >= se_walk_exact_scratch_bytes(max_degree, n_walks) bytes. */
int64_t se_walk_exact_scratch_bytes(int64_t max_degree, int64_t n_walks);
int se_walk_exact(const int64_t *rowptr, const int32_t *col, const int32_t *col_sorted, const double *w, int w_is_int,
                  int64_t n_nodes, int64_t max_degree, const int32_t *starts, int64_t n_walks, int walk_len,
                  double p, double q, int node2vec, int rule, const double *uniforms,
                  void *scratch, int64_t scratch_bytes, int32_t *out, void *stream);

/* Fast mode: counter-based Philox4x32-10 keyed by (seed; walk id, step, try), rejection sampling of the 1/p, 1, 1/q
 * multipliers with sorted-row membership tests.  Two kernels generate bit-identical walks: one warp per walk (32 tries
 * per round in parallel, neighbour lists up to 128 entries staged in shared memory) and one thread per walk
 * (interpolation search, vector stores) -- see `flags`.  Same transition distribution as the reference rule (validated by chi-square), not the same draws.
 *   col[nnz]: each row ASCENDING;  wcdf[nnz]: per-row inclusive prefix sums of the edge weights (fp32) or NULL,
 *   walk id of out row i = walk_id_base + i * walk_id_stride  (so any sharding reproduces the 1-GPU walks),
 *   symmetric != 0 promises x in N(t) <=> t in N(x) (undirected graph) and enables the staged-list membership test,
 *   err_count (int32[1], may be NULL) counts walks that hit a degree-0 node (the reference raises there; the walk
 *   stays on that node). */
#define SE_WALK_AUTO 0   /* pick the kernel from the batch size (results are identical either way) */
#define SE_WALK_WARP 1   /* one warp per walk: lowest latency for small batches */
#define SE_WALK_THREAD 2 /* one thread per walk: highest throughput for large batches */
int se_walk(const int64_t *rowptr, const int32_t *col, const float *wcdf, int64_t n_nodes, int symmetric,
            const int32_t *starts, int64_t n_walks, int walk_len, double p, double q, int node2vec, int rule,
            uint64_t seed, int64_t walk_id_base, int64_t walk_id_stride, int32_t *out, int32_t *err_count,
            int flags, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Negative sampling.  Replaces generate_noise_batch (word2vec/utils/sampling.py:7-21): the reference draws
 * UNIFORM ids over [0, V); the alias table adds unigram^power (power = 0.75 is word2vec's; power = 0 reproduces
 * the reference distribution).
 * ---------------------------------------------------------------------------------------------------------- */

/* Vose alias table on the HOST (one-off set-up, O(V)); counts_host[V] >= 0, prob_host[V] fp32, alias_host[V] int32. */
int se_alias_build_host(const double *counts_host, int64_t vocab, double power, float *prob_host, int32_t *alias_host);

/* out[n] int64 ids; id i is Philox(seed; draw_id_base + i).  prob/alias NULL -> uniform over [0, vocab). */
int se_sample_negatives(const float *prob, const int32_t *alias, int64_t vocab, uint64_t seed, int64_t draw_id_base,
                        int64_t n, int64_t *out, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * SkipGram / SGNS on explicit index tensors (the reference's (inputs (B,1), targets (B,N), noise (B,N,K)) int64).
 * Tables are fp32 row-major [vocab x emb], 16-byte aligned.
 * ---------------------------------------------------------------------------------------------------------- */

/* SkipGram.forward (word2vec/model.py:79-91): out[b*m + j] = <W_in[inputs[b]], W_out[outputs[b*m + j]]>,
 * sigmoid applied when proba != 0. */
int se_skipgram_scores(const float *w_in, const float *w_out, int64_t vocab, int emb, const int64_t *inputs,
                       const int64_t *outputs, int64_t batch, int m, int proba, float *out, void *stream);

/* Backward of se_skipgram_scores with proba = 0 (what autograd does for model.py:85-88): ACCUMULATES
 * grad_in[inputs[b]] += sum_j g[b,j] * W_out[outputs[b,j]] and grad_out[outputs[b,j]] += g[b,j] * W_in[inputs[b]]
 * into dense [vocab x emb] buffers. */
int se_skipgram_scores_backward(const float *w_in, const float *w_out, int64_t vocab, int emb, const int64_t *inputs,
                                const int64_t *outputs, int64_t batch, int m, const float *grad_scores, float *grad_in,
                                float *grad_out, void *stream);

/* CBOW.forward (word2vec/model.py:98-110): out[b*m + j] = <mean_n W_in[inputs[b*n_in + n]], W_out[outputs[b*m + j]]>, sigmoid when
 * proba != 0.  inputs (B, n_in) are the 2r context ids, outputs (B, m) the centre (m = 1) or noise ids (collate mode `cbow`,
 * word2vec/dataloader/torch_dataset.py:310-314).  emb <= 1024. */
int se_cbow_scores(const float *w_in, const float *w_out, int64_t vocab, int emb, const int64_t *inputs, const int64_t *outputs,
                   int64_t batch, int n_in, int m, int proba, float *out, void *stream);
/* Word2VecTrainer.training_step + backward for a CBOW model (trainer.py:131-139, loss.py:14-22): targets (B, m), noise (B, m, K);
 * stats as se_sgns_grad (pairs = B * m); dense gradients of the MEAN loss ACCUMULATED into grad_in / grad_out (both or neither). */
int se_cbow_grad(const float *w_in, const float *w_out, int64_t vocab, int emb, const int64_t *inputs, const int64_t *targets,
                 const int64_t *noise, int64_t batch, int n_in, int m, int n_neg, double *stats, float *grad_in, float *grad_out,
                 void *stream);

/* NegativeSamplingLoss.forward on logits (word2vec/loss.py:14-22): pos_logits (B,N), neg_logits (B,N,K) -> stats
 * (same layout as above) and, when non-NULL, grad_pos (B,N) = d mean(positive-loss)/d pos_logits and
 * grad_neg (B,N,K) = d mean(negative-loss)/d neg_logits. */
int se_ns_loss(const float *pos_logits, const float *neg_logits, int64_t batch, int n_ctx, int n_neg, double *stats,
               float *grad_pos, float *grad_neg, void *stream);

/* Word2VecTrainer.training_step + backward (word2vec/trainer.py:131-152, word2vec/loss.py:14-22): loss sums /
 * counters into stats (divide by stats[4] for the reference's means) and, when grad_in/grad_out are non-NULL,
 * ACCUMULATES the dense gradients of the MEAN loss into them (caller zeroes; same values as autograd). */
int se_sgns_grad(const float *w_in, const float *w_out, int64_t vocab, int emb, const int64_t *inputs,
                 const int64_t *targets, const int64_t *noise, int64_t batch, int n_ctx, int n_neg,
                 double *stats, float *grad_in, float *grad_out, void *stream);

/* Row-sparse Adam step on an explicit batch -- what `loss['loss'].backward(); optimizer.step()` does in the reference when the YAML
 * names `_target_: torch.optim.Adam` (configs/sge_sg_karate_club.yaml:32-34, config_parser/core.py:43-94), restricted to the rows
 * the batch touches: gradient of the MEAN loss (word2vec/loss.py:19) accumulated per row, then m, v, theta updated with torch's
 * formulas and bias correction by the row's own step count.  Equals dense torch.optim.Adam whenever every step touches the same
 * rows (and on the first step); untouched rows keep their value.  All state is caller-owned device memory:
 *   m_*, v_*, g_* fp32 [vocab x emb] (zero-initialised; g_* is zero again when the call returns), t_* int32 [vocab] step counts,
 *   touched_* int32 [vocab] flags (zero between calls), list_in / list_out int32 row lists with at least min(vocab, batch) /
 *   min(vocab, batch * n_ctx * (1 + n_neg)) entries, counts int32[4] (zero-initialised).  stats as se_sgns_grad. */
typedef struct se_adam_state {
    float *m_in, *v_in, *m_out, *v_out;
    float *g_in, *g_out;
    int32_t *t_in, *t_out;
    int32_t *touched_in, *touched_out;
    int32_t *list_in, *list_out;
    int64_t list_in_capacity, list_out_capacity;
    int32_t *counts;
} se_adam_state;
int se_sgns_adam_step(float *w_in, float *w_out, int64_t vocab, int emb, const int64_t *inputs, const int64_t *targets,
                      const int64_t *noise, int64_t batch, int n_ctx, int n_neg, const se_adam_state *state, float lr, float beta1,
                      float beta2, float eps, double *stats, void *stream);

/* Fused in-place SGD on an explicit batch: every row touched by pair (b, n) moves by -lr * dL_pair/drow where
 * L_pair is the un-averaged per-pair loss (pass lr = lr_ref / (batch * n_ctx) for the reference's mean loss).
 * noise NULL -> K negatives per pair drawn in-kernel (alias or uniform), keyed (seed; pair_id_base + b, n, k). */
int se_sgns_step(float *w_in, float *w_out, int64_t vocab, int emb, const int64_t *inputs, const int64_t *targets,
                 const int64_t *noise, int64_t batch, int n_ctx, int n_neg, const float *alias_prob,
                 const int32_t *alias_idx, float lr, uint64_t seed, int64_t pair_id_base, int flags, double *stats,
                 void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * The fused hot path: walks -> skip-gram windows -> negatives -> in-place SGNS update, nothing materialised.
 * Replaces W2VCollateFunctional sg branch (word2vec/dataloader/torch_dataset.py:293-322: centres i in
 * [radius, L - radius), 2*radius contexts each), generate_noise_batch, SkipGram.forward x2, NegativeSamplingLoss,
 * backward and the optimiser step (word2vec/trainer.py:131-152) for tokens[n_seq * seq_len] int32 (node ids or
 * token ids); table row = token + row_offset.  Negative k of (centre c, context n) is keyed
 * (seed; centre_id_base + c, n * n_neg + k) so results do not depend on the launch geometry.
 * ---------------------------------------------------------------------------------------------------------- */
int se_sgns_update_walks(float *w_in, float *w_out, int64_t vocab, int emb, const int32_t *tokens, int64_t n_seq,
                         int seq_len, int radius, int n_neg, int row_offset, const float *alias_prob,
                         const int32_t *alias_idx, float lr, uint64_t seed, int64_t centre_id_base, int flags,
                         double *stats, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Host-buffer pipeline step (what bench.py's e2e leg and RandomWalkDataset-driven training call): copies
 * starts_host[n_walks] to the device, runs se_walk + se_sgns_update_walks on `stream`, copies the SE_STATS_LEN
 * doubles back into stats_host and synchronises.  Device scratch: starts_dev[n_walks] int32,
 * walks_dev[n_walks*walk_len] int32, stats_dev[SE_STATS_LEN] double.  walks_host may be NULL; when non-NULL the
 * walks are copied back as well (the reference's RandomWalk.walk returns them to the caller).
 * ---------------------------------------------------------------------------------------------------------- */
int se_host_walk_sgns_step(const int64_t *rowptr, const int32_t *col, const float *wcdf, int64_t n_nodes,
                           int symmetric, const int32_t *starts_host, int64_t n_walks, int walk_len, double p,
                           double q, int node2vec, int rule, uint64_t seed, int64_t walk_id_base,
                           float *w_in, float *w_out, int64_t vocab, int emb, int radius, int n_neg, int row_offset,
                           const float *alias_prob, const int32_t *alias_idx, float lr, int flags,
                           int32_t *starts_dev, int32_t *walks_dev, double *stats_dev, int32_t *walks_host,
                           double *stats_host, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU: sharded embedding tables over NVLink / NVSwitch peer memory.  The reference is single-device
 * (`devices: '1'` in every YAML, e.g. configs/sge_sg_cora.yaml:30; tools/train.py:67-81 passes no strategy), so
 * this group has no reference counterpart: it lets the SAME fused SGNS kernel above run on G GPUs against ONE pair
 * of tables whose rows are striped over the G HBMs -- the row "all-to-all" and the gradient "all-to-all" of a
 * row-sharded word2vec become the kernel's own 128-bit loads and red.global.add.v4.f32 to peer memory.
 *
 * A sharded table is a flat fp32 [vocab x emb] array at one virtual address range per process; stripe s
 * (stripe_bytes each, a multiple of se_shard_granularity) is a separate physical allocation living on rank
 * s % world.  Set-up per process (one process per GPU; host plumbing in shallow_encoders/word2vec/sharded.py):
 *   se_shard_reserve the range; for every stripe: the owner se_shard_create + se_shard_export_fd and sends the fd
 *   over a unix socket (SCM_RIGHTS), peers se_shard_import_fd; everybody se_shard_map at offset s * stripe_bytes.
 * Handles and addresses are plain uint64 (CUmemGenericAllocationHandle / CUdeviceptr).
 * ---------------------------------------------------------------------------------------------------------- */
int se_shard_granularity(int64_t *bytes);                       /* minimum stripe size on the current device */
int se_shard_reserve(int64_t bytes, uint64_t *va);
int se_shard_unreserve(uint64_t va, int64_t bytes);
int se_shard_create(int64_t bytes, uint64_t *handle);           /* physical stripe in the current device's HBM */
int se_shard_release(uint64_t handle);
int se_shard_export_fd(uint64_t handle, int *fd);               /* caller closes fd after sending it */
int se_shard_import_fd(int fd, uint64_t *handle);               /* caller closes fd afterwards */
int se_shard_map(uint64_t va, int64_t bytes, uint64_t handle);  /* map + read/write access for the current device */
int se_shard_unmap(uint64_t va, int64_t bytes);

/* How the tables passed to the *_sharded entry points are striped. */
typedef struct se_shard_spec {
    int32_t world;           /* GPUs the tables are striped over (1 = one local table) */
    int32_t rank;            /* the shard whose stripes live in the calling GPU's HBM */
    int64_t stripe_rows;     /* rows per stripe: row r lives on rank (r / stripe_rows) % world */
    int32_t local_negatives; /* != 0: negatives are drawn only among the rows owned by `rank` (uniform, or from an
                                alias table built over those local rows: alias arrays then have se_shard_local_rows
                                entries).  Walks are dealt to GPUs by walk id, so over the job every centre still meets
                                negatives from every shard; it removes K/(K+1+1/N) of the NVLink traffic.  == 0: the
                                reference's distribution over all of [0, vocab) (word2vec/utils/sampling.py:21) */
    int32_t reserved;
} se_shard_spec;

/* number of rows of [0, vocab) owned by spec->rank */
int se_shard_local_rows(int64_t vocab, const se_shard_spec *spec, int64_t *n_rows);

/* se_sgns_update_walks / se_host_walk_sgns_step on sharded tables (spec NULL = the unsharded calls).  With
 * spec->world > 1 the scatter uses system-scope reductions (every GPU's concurrent updates of a row land at the
 * owner's L2); SE_SGNS_SCATTER_STORE is refused. */
int se_sgns_update_walks_sharded(float *w_in, float *w_out, int64_t vocab, int emb, const int32_t *tokens,
                                 int64_t n_seq, int seq_len, int radius, int n_neg, int row_offset,
                                 const float *alias_prob, const int32_t *alias_idx, float lr, uint64_t seed,
                                 int64_t centre_id_base, int flags, const se_shard_spec *spec, double *stats,
                                 void *stream);
/* Owner-computes negatives: the NEGATIVE half of se_sgns_update_walks for tokens that may belong to ANY GPU's walks
 * (all-gathered), restricted to the negatives whose rows spec->rank owns.  Negative k of (centre c, context n) is drawn
 * from Philox(seed; centre_id_base + c, n, k) exactly as in se_sgns_update_walks (word2vec/utils/sampling.py:21
 * distribution over the WHOLE table, or the alias table), so calling it on every rank with the same tokens and keys,
 * plus se_sgns_update_walks with n_neg = 0 on each rank's own walks for the positive pairs, performs the same pair
 * updates as one GPU would -- with 2 rows per centre per GPU over NVLink instead of 2*N*K.  stats: [1], [3], [5]. */
int se_sgns_update_negatives_owned(float *w_in, float *w_out, int64_t vocab, int emb, const int32_t *tokens,
                                   int64_t n_seq, int seq_len, int radius, int n_neg, int row_offset,
                                   const float *alias_prob, const int32_t *alias_idx, float lr, uint64_t seed,
                                   int64_t centre_id_base, const se_shard_spec *spec, double *stats, void *stream);
/* Owner-computes, grouped: the centres of the batch are first bucketed by table row on the device (counting sort in `scratch`,
 * se_pairs_owned_scratch_bytes; own kernels); a warp keeps the centre row of a run of equal rows in registers -- one read and one
 * reduction per run instead of per occurrence (10 M distinct rows against 147 M occurrences per step on S3 at 8 GPUs) -- and the
 * owned output rows of consecutive occurrences share full passes.  Negatives: same Philox keys, hence the same (centre, negative)
 * pairs as se_sgns_update_negatives_owned and se_sgns_update_walks (sampling.py:21 distribution over the WHOLE table).
 * positives != 0: the 2r context tokens of every centre (window rule torch_dataset.py:300-309) are processed as well when
 * spec->rank owns their W_out row, so that calling it on every rank with the same gathered tokens performs EVERY pair of the batch
 * exactly once, on the owner of its output row: W_out never crosses NVLink and no separate positive pass is needed
 * (stats [0], [2], [4] then count the owned positives).  The ORDER of the updates differs from walk order (Hogwild).
 * Token ids outside [0, vocab) are skipped as centres.  The default of the multi-GPU owner-computes mode. */
int64_t se_pairs_owned_scratch_bytes(int64_t vocab, int64_t n_seq, int seq_len, int radius);
int se_sgns_update_pairs_owned(float *w_in, float *w_out, int64_t vocab, int emb, const int32_t *tokens, int64_t n_seq,
                               int seq_len, int radius, int n_neg, int positives, int row_offset,
                               const float *alias_prob, const int32_t *alias_idx, float lr, uint64_t seed,
                               int64_t centre_id_base, const se_shard_spec *spec, void *scratch,
                               int64_t scratch_bytes, double *stats, void *stream);
int se_host_walk_sgns_step_sharded(const int64_t *rowptr, const int32_t *col, const float *wcdf, int64_t n_nodes,
                                   int symmetric, const int32_t *starts_host, int64_t n_walks, int walk_len, double p,
                                   double q, int node2vec, int rule, uint64_t seed, int64_t walk_id_base,
                                   float *w_in, float *w_out, int64_t vocab, int emb, int radius, int n_neg,
                                   int row_offset, const float *alias_prob, const int32_t *alias_idx, float lr,
                                   int flags, const se_shard_spec *spec, int32_t *starts_dev, int32_t *walks_dev,
                                   double *stats_dev, int32_t *walks_host, double *stats_host, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU with the reference's GLOBAL negative draw (word2vec/utils/sampling.py:21) at NVLink bulk rate: per-GPU
 * working copies + row-sharded masters (csrc/replica.cu).  No reference counterpart (`devices: '1'`).
 * `base` is one virtual range of `world` segments of stride_elems floats (segment g = GPU g's full [vocab x emb] working
 * copy, physically in GPU g's HBM, mapped into every process through the se_shard_* calls above); the single-GPU kernels
 * train on segment `rank`.  Rank r owns the master of the element chunk se_replica_chunk returns (contiguous rows).
 * se_replica_sync, enqueued by EVERY rank between two inter-GPU barriers, does for the chunk the caller owns
 *     mode 0:  master' = master + beta * sum_g (copy_g - master);  every copy_g <- master'   (fused reduce-scatter + all-gather)
 *     mode 1:  master <- copy_rank                                                     (after initialisation)
 *     mode 2:  every copy_g <- master                                                  (after loading a checkpoint)
 * beta = 1 is synchronous data-parallel SGD with SUMMED updates (right while a step touches a row about once), beta = 1 / world
 * is local SGD with model averaging (right when every GPU updates every row many times per step); 0 < beta <= 1.
 * `master` holds hi_elem - lo_elem floats.
 * ---------------------------------------------------------------------------------------------------------- */
int se_replica_chunk(int64_t n_elems, int world, int rank, int64_t *lo_elem, int64_t *hi_elem);
int se_replica_sync(float *base, int64_t stride_elems, int world, int rank, int64_t n_elems, float *master, int mode, float beta,
                    void *stream);

/* Input hygiene for token / node ids that come from outside the library (the reference's nn.Embedding raises IndexError on an
 * out-of-range id, word2vec/model.py:22-23; the fused kernels index the tables unchecked): ADDS to bad_count[0] the number of
 * ids[i] outside [lo, hi).  The Python wrappers run it before a fused update unless told the ids come from se_walk. */
int se_check_ids(const int32_t *ids, int64_t n, int64_t lo, int64_t hi, int32_t *bad_count, void *stream);

/* Host-buffer step for sequences that are already token ids (the text path after W2VDataset.sentence_pipeline,
 * word2vec/dataloader/torch_dataset.py:124-156, 205-213: one int32 id per token, sentences of equal length): copies
 * tokens_host[n_seq * seq_len] to tokens_dev, runs the fused window / negatives / SGNS update, copies the SE_STATS_LEN
 * doubles back and synchronises.  spec may be NULL. */
int se_host_sgns_update_tokens(const int32_t *tokens_host, int64_t n_seq, int seq_len, float *w_in, float *w_out,
                               int64_t vocab, int emb, int radius, int n_neg, int row_offset, const float *alias_prob,
                               const int32_t *alias_idx, float lr, uint64_t seed, int64_t centre_id_base, int flags,
                               const se_shard_spec *spec, int32_t *tokens_dev, double *stats_dev, double *stats_host,
                               void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Link-prediction features on device-resident embeddings (the evaluation step right after training).
 * Replaces shallow_encoders/graph/edge_operators.py:10-64 (average, hadamard, weighted_l1, weighted_l2),
 * create_edge_embeddings and sample_negative_edges (tools/graph_model_downstream_classification.py:170-224).
 * ---------------------------------------------------------------------------------------------------------- */
#define SE_EDGE_AVERAGE 0     /* (lhs + rhs) / 2 */
#define SE_EDGE_HADAMARD 1    /* lhs * rhs */
#define SE_EDGE_WEIGHTED_L1 2 /* |lhs - rhs| */
#define SE_EDGE_WEIGHTED_L2 3 /* (lhs - rhs) ** 2 */

/* out[i, :] = op(table[src_rows[i], :], table[dst_rows[i], :]);  table fp32 [vocab x emb], out fp32 [n_edges x emb]. */
int se_edge_features(const float *table, int64_t vocab, int emb, const int64_t *src_rows, const int64_t *dst_rows,
                     int64_t n_edges, int op, float *out, void *stream);
/* out[i] = op(lhs[i], rhs[i]) on dense operands (what edge_operator_factory(name)(lhs, rhs) computes). */
int se_edge_op(const float *lhs, const float *rhs, int64_t n_elems, int op, float *out, void *stream);
/* n non-edges: node uniform over [0, n_nodes), partner uniform over the nodes that are not its neighbours (the node
 * itself included, like the reference's set difference); col_sorted rows ascending; Philox(seed; sample_id_base + i, try).
 * fail_count (int32[1], may be NULL) counts samples that found no non-neighbour in 4096 tries (complete graphs). */
int se_sample_negative_edges(const int64_t *rowptr, const int32_t *col_sorted, int64_t n_nodes, int64_t n, uint64_t seed,
                             int64_t sample_id_base, int32_t *out_src, int32_t *out_dst, int32_t *fail_count, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Tensor-core contraction (csrc/gemm.cu: tcgen05.mma kind::tf32 with TMEM accumulators, 3xTF32 split = fp32-level accuracy).
 *   se_gemm_nt:            c[i, j] = scale_a[i] * scale_b[j] * <a_i, b_j>, a [m x kdim], b [n x kdim], c [m x n], fp32 row-major;
 *                          scale vectors may be NULL (= 1)
 *   se_row_inv_norms:      out[i] = 1 / |x_i|
 *   se_cosine_similarity:  pairwise_cosine_similarity(x, y) of shallow_encoders/word2vec/utils/func.py:7-20 (x / |x| times (y / |y|)^T),
 *                          what show_closest_pairs_for_each_word computes before its per-row argsort (tools/model_analysis.py:62-71);
 *                          inv_norms: scratch for m + n floats
 *   se_topk_rows:          the k largest entries per row, descending = torch.argsort(row, descending=True)[:k] (model_analysis.py:71);
 *                          idx_out int64 [rows x k], val_out fp32 [rows x k] or NULL
 *   se_transpose:          out [cols x rows] = x^T (operands of the shared-negatives update GEMMs)
 * ---------------------------------------------------------------------------------------------------------- */
int se_gemm_nt(const float *a, const float *b, int64_t m, int64_t n, int kdim, const float *scale_a, const float *scale_b, float *c,
               void *stream);
int se_row_inv_norms(const float *x, int64_t rows, int emb, float *out, void *stream);
int se_cosine_similarity(const float *x, const float *y, int64_t m, int64_t n, int emb, float *inv_norms, float *out, void *stream);
int se_topk_rows(const float *x, int64_t rows, int64_t cols, int k, int64_t *idx_out, float *val_out, void *stream);
int se_transpose(const float *x, int64_t rows, int64_t cols, float *out, void *stream);

/* Optional SHARED-NEGATIVES batch mode (csrc/shared_neg.cu): the negative half of an SGNS step for `batch` centres against ONE set
 * of n_shared negative rows, as three tensor-core GEMMs (scores B x S, centre updates B x E, negative-row updates S x E) instead of
 * per-pair dot products; each shared row is weighted n_ctx * n_neg / n_shared, the number of per-pair negatives of the reference
 * (word2vec/utils/sampling.py:21, trainer.py:133-138) it stands for.  inputs[batch] centre rows, shared[n_shared] negative rows
 * (duplicates accumulate); lr multiplies the un-averaged per-pair gradient like se_sgns_step; the positive pairs go through
 * se_sgns_step with n_neg = 0.  stats: [1] += weighted negative loss, [3] += weighted count of sigmoid(s-) >= 0.5.
 * scratch: se_shared_negatives_scratch_floats(batch, n_shared, emb) floats, 16-byte aligned. */
int64_t se_shared_negatives_scratch_floats(int64_t batch, int64_t n_shared, int emb);
int se_sgns_step_shared_negatives(float *w_in, float *w_out, int64_t vocab, int emb, const int64_t *inputs, int64_t batch,
                                  const int64_t *shared, int64_t n_shared, int n_ctx, int n_neg, float lr, float *scratch,
                                  int64_t scratch_floats, double *stats, void *stream);

/* Elementwise half of a logistic-regression gradient for the device-side downstream classifier (the reference fits sklearn's
 * LogisticRegression on the host, tools/graph_model_downstream_classification.py:85-91): logits fp32 [n x n_cols] (+ bias[n_cols] or
 * NULL), labels int32 [n].  n_cols == 1: binary logistic (label 1 = positive); else softmax.  ADDS the summed loss to *loss_sum and
 * the number of correctly classified rows to *n_correct (either may be NULL) and OVERWRITES logits with grad_scale * d loss / d logits.
 * The two GEMMs around it (X W^T and G^T X) are se_gemm_nt. */
int se_softmax_xent(float *logits, const int32_t *labels, const float *bias, int64_t n, int n_cols, float grad_scale, double *loss_sum,
                    int32_t *n_correct, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Graph ingest on the device (csrc/ingest.cu): edge list -> CSR with networkx's semantics for a simple graph -- what the reference
 * does edge by edge on the host (graph/datasets.py:126-221: nx.Graph.add_edge per line, e.g. cora.cites at :199-200) before
 * random_walk_generator.py:41-48 re-reads the adjacency on every step.
 *   src / dst int32 [n_edges] node ids, w fp64 [n_edges] or NULL; symmetrize != 0 stores every edge in both rows (undirected graph);
 *   self loops and edges with an endpoint outside [0, n_nodes) are dropped (the latter counted in info[2]); duplicate edges collapse
 *   to one entry with the weight of the LAST occurrence (a repeated add_edge overwrites the attributes).
 *   Outputs: rowptr_out int64 [n_nodes + 1]; col_out int32 (rows ascending: serves as CDF order and as membership order), w_out fp64 and
 *   wcdf_out fp32 (per-row inclusive prefix sums; NULL with unweighted input) sized for the worst case n_edges * (symmetrize ? 2 : 1);
 *   info int64[3] on the device: [0] nnz, [1] max degree, [2] skipped edges.  scratch: se_csr_build_scratch_bytes, 16-byte aligned.
 * ---------------------------------------------------------------------------------------------------------- */
int64_t se_csr_build_scratch_bytes(int64_t n_nodes, int64_t n_edges, int symmetrize);
int se_csr_build(const int32_t *src, const int32_t *dst, const double *w, int64_t n_edges, int64_t n_nodes, int symmetrize, void *scratch,
                 int64_t scratch_bytes, int64_t *rowptr_out, int32_t *col_out, double *w_out, float *wcdf_out, int64_t *info, void *stream);

/* Table utilities that work on local and sharded tables alike (W2VBase.__init__ xavier_uniform_, word2vec/model.py:22-27;
 * the input_embedding / output_embedding accessors, :29-47).
 *   fill: element i = (2u-1)*bound with u from Philox(seed; i/4) -- independent of the sharding; a rank writes only the
 *         stripes it owns (stripe_elems = stripe_rows * emb; 0 = write everything)
 *   gather / scatter: out[i,:] = w[rows[i],:]  /  w[rows[i],:] = src[i,:] */
int se_table_fill_uniform(float *w, int64_t n_elems, float bound, uint64_t seed, int64_t stripe_elems, int world,
                          int rank, void *stream);
int se_table_gather_rows(const float *w, int emb, const int64_t *rows, int64_t n, float *out, void *stream);
int se_table_scatter_rows(float *w, int emb, const int64_t *rows, int64_t n, const float *src, void *stream);
/* nn.Embedding(max_norm) (word2vec/model.py:22-23, set by configs/w2v_sg_abcde.yaml:7): rows[i] (DISTINCT ids) whose L2 norm exceeds
 * max_norm are rescaled in place by max_norm / (norm + 1e-7), what torch's embedding_renorm_ does at look-up time. */
int se_table_renorm_rows(float *w, int emb, const int64_t *rows, int64_t n, float max_norm, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SE_B200_H */

/*
 * demovlp_b200 -- C ABI of the B200-native region-word alignment path.
 *
 * Drop-in boundary for the hot path of showlab/DemoVLP (reference paths are
 * relative to the reference root).  Every entry point replaces a piece of the
 * reference's pure-PyTorch loss/metric code; the Python host mirror in
 * demovlp_b200/loss.py binds them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the
 *     name ends in _host; the caller owns every buffer;
 *   - nothing here allocates or synchronises: work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), workspace comes from the caller
 *     (size from dvlp_rw_workspace_bytes);
 *   - return value 0 = success, negative = DVLP_ERR_* (dvlp_error_string);
 *   - tensors are dense row-major ("contiguous"), d is the embedding width;
 *   - masks are the reference's ADDITIVE masks (0 = valid, -100 = padded,
 *     trainer/trainer_dist.py:156-159), fp32;
 *   - in_dtype: DVLP_BF16 (2-byte bfloat16) or DVLP_F32.
 */
#ifndef DEMOVLP_B200_H_
#define DEMOVLP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVLP_F32 0
#define DVLP_BF16 1

#define DVLP_OK 0
#define DVLP_ERR_BAD_SHAPE (-1)      /* non-positive or unsupported dimension            */
#define DVLP_ERR_BAD_DTYPE (-2)      /* in_dtype is neither DVLP_F32 nor DVLP_BF16       */
#define DVLP_ERR_NULL_POINTER (-3)   /* a required pointer is NULL                       */
#define DVLP_ERR_MISALIGNED (-4)     /* pointer not 16-byte aligned                      */
#define DVLP_ERR_WORKSPACE (-5)      /* workspace too small / missing                    */
#define DVLP_ERR_CUDA (-6)           /* a CUDA call failed (see dvlp_last_cuda_error)    */
#define DVLP_ERR_NO_DEVICE (-7)      /* no sm_100 device: there is NO CPU fallback       */
#define DVLP_ERR_BAD_ARG (-8)        /* other invalid argument                           */
#define DVLP_ERR_TIMEOUT (-9)        /* a peer-memory exchange gave up waiting for a rank */

/* implementation selector for the scoring kernels (debug / A-B testing) */
#define DVLP_IMPL_AUTO 0     /* tcgen05 path when the shape qualifies, else generic    */
#define DVLP_IMPL_GENERIC 1  /* generic-shape fp32 CUDA-core kernels                    */
#define DVLP_IMPL_TCGEN05 2  /* tcgen05/TMEM/TMA kernels; error if shape unsupported    */

int dvlp_version(void);
const char* dvlp_error_string(int code);
/* cudaError_t of the last failing CUDA call made by this library on this thread. */
int dvlp_last_cuda_error(void);
/* 1 when dvlp_rw_scores_* would take the tcgen05 path for this shape. */
int dvlp_rw_tcgen05_supported(int R, int W, int d);
/* which pair kernels DVLP_IMPL_AUTO selects for a shape: 0 = generic fp32 CUDA cores, 2 = tcgen05 small-shape kernels
 * (W = 32, R in {32,48,64,80}), 3 = tcgen05 large-shape streaming kernel (any R, W <= 128 it can tile: R = 288 x W = 64,
 * the reference's shipped R = 30 / 240 x W = 99, ...).  train = 0: forward-only calls (dvlp_rw_scores_infer). */
int dvlp_rw_path(int R, int W, int d, int train);
/* process-wide default for `impl` arguments given as DVLP_IMPL_AUTO. */
void dvlp_set_default_impl(int impl);
/* 1 when dvlp_rw_scores_bwd can write its gradients in `out_dtype` for this impl / shape / lambda
 * (DVLP_F32 always; DVLP_BF16 only where the tcgen05 finalisation kernel serves the shape). */
int dvlp_rw_bwd_out_dtype_supported(int impl, int R, int W, int d, float lambda, int out_dtype);
/* debugging aid: device buffer (5*32*16 int64) that the backward pair kernel of CTA 0 fills with per-phase clock
 * stamps (tools/trace_tiles.py); NULL (default) disables it.  Not part of the reference-facing surface. */
void dvlp_debug_set_trace(void* dev_buffer);
/* number of kernels this library has launched so far in this process. */
uint64_t dvlp_launch_count(void);

/*
 * Optional kernel timing for benchmarks: when enabled, CUDA events are recorded on the launching
 * stream around 0 = forward pair kernel, 1 = backward pair kernel, 2 = per-item prologue,
 * 3 = gradient finalisation, 4 = gradient GEMMs.  dvlp_profile_read synchronises those events, writes the
 * summed milliseconds and launch-group counts of the five slots (ms[5], counts[5], HOST pointers) and resets.
 */
void dvlp_profile_enable(int on);
int dvlp_profile_read(double* ms_host, int* counts_host);

/* Bytes of scratch dvlp_rw_scores_fwd / dvlp_rw_scores_bwd need for this shape.  The query does not take lambda: under
 * DVLP_IMPL_AUTO it covers both the tensor-core path and the fp32 path a lambda > 40 call is routed to. */
size_t dvlp_rw_workspace_bytes(int Bv, int Bt, int R, int W, int d, int in_dtype, int impl);

/*
 * Region-word pair scores, forward.
 * Replaces RWALoss.get_sim -> xattn_score_fast -> 2 x func_attention_fast ->
 * cosine_similarity (model/loss.py:64-71, 294-330, 209-271, 286-291).
 *   im      (Bv, R, d)  video region embeddings        s       (Bt, W, d)  word embeddings
 *   im_mask (Bv, R)     additive region mask           s_mask  (Bt, W)     additive word mask
 *   lambda  = lambda_softmax (20)      focal_equal = 1 for focal_type "equal", 0 otherwise
 *   scores  (Bv, Bt) fp32, overwritten.
 * The workspace keeps the per-item prologue (inverse row norms, Gram matrices)
 * that dvlp_rw_scores_bwd re-uses when called with the same workspace.
 */
int dvlp_rw_scores_fwd(const void* im, const void* s, const float* im_mask, const float* s_mask,
                       int Bv, int Bt, int R, int W, int d, float lambda, int focal_equal,
                       int in_dtype, int impl, float* scores,
                       void* workspace, size_t workspace_bytes, void* stream);

/*
 * Forward only (validation / retrieval: RWALoss.get_sim under no_grad, get_sim_by_segment,
 * model/loss.py:64-103, trainer/trainer_dist.py:369-385): same arguments and scores as
 * dvlp_rw_scores_fwd, but nothing is kept for a backward -- the workspace
 * (dvlp_rw_infer_workspace_bytes) holds the per-item prologue only, so a 1000 x 1000 evaluation
 * needs megabytes instead of gigabytes and the kernel writes nothing but the scores.
 */
size_t dvlp_rw_infer_workspace_bytes(int Bv, int Bt, int R, int W, int d, int in_dtype, int impl);
int dvlp_rw_scores_infer(const void* im, const void* s, const float* im_mask, const float* s_mask,
                         int Bv, int Bt, int R, int W, int d, float lambda, int focal_equal,
                         int in_dtype, int impl, float* scores,
                         void* workspace, size_t workspace_bytes, void* stream);

/*
 * Grouped forward (multiple-choice evaluation, trainer/trainer_dist.py:792-829): the reference expands ONE video
 * n_options times and calls get_sim on a 5 x 5 problem per question; here n_groups questions go through ONE launch:
 *   im (n_groups, R, d): the video of every question      s (n_groups * n_per_group, W, d): its options, group-major
 *   scores (n_groups, n_per_group) fp32: scores[g][o] = score(video g, option o of question g)
 * -- exactly row 0 of the reference's per-question local_sims_batch.  tcgen05 large-shape kernel only (d = 256,
 * W <= 128; dvlp_rw_grouped_workspace_bytes returns 0 for other shapes, which callers serve question by question
 * with dvlp_rw_scores_infer).
 */
size_t dvlp_rw_grouped_workspace_bytes(int n_groups, int n_per_group, int R, int W, int d, int in_dtype);
int dvlp_rw_scores_grouped_infer(const void* im, const void* s, const float* im_mask, const float* s_mask,
                                 int n_groups, int n_per_group, int R, int W, int d, float lambda, int focal_equal,
                                 int in_dtype, float* scores, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Backward of dvlp_rw_scores_fwd (what autograd does through model/loss.py:209-330):
 * given d_scores (Bv, Bt) fp32 produces d_im (Bv, R, d) and d_s (Bt, W, d) in `out_dtype`
 * (both overwritten; DVLP_F32 always, DVLP_BF16 on the tcgen05 path, else DVLP_ERR_BAD_DTYPE).
 * The attention tensors are recomputed, never stored.
 * `workspace` must be the one the forward call used (same shape, inputs unchanged).
 */
int dvlp_rw_scores_bwd(const void* im, const void* s, const float* im_mask, const float* s_mask,
                       int Bv, int Bt, int R, int W, int d, float lambda, int focal_equal,
                       int in_dtype, int impl, const float* d_scores, void* d_im, void* d_s, int out_dtype,
                       void* workspace, size_t workspace_bytes, void* stream);

/*
 * RWALoss tail (model/loss.py:108-114): KL(softmax(lambda*scores, dim=1) || eye + 1e-6).
 * scores is the (n_rows, n_cols) block holding global rows [row_offset, row_offset+n_rows);
 * the target 1 of local row i sits at column row_offset + i.
 *   loss_sum  1 float, overwritten with  sum_i KL_i / n_total
 *   d_scores  (n_rows, n_cols) or NULL:  d(loss_sum)/d(scores)
 * Single GPU: n_rows == n_cols == n_total, row_offset 0.
 */
int dvlp_rwa_loss_fwd_bwd(const float* scores, int n_rows, int n_cols, int row_offset, int n_total,
                          float lambda, float* loss_sum, float* d_scores, void* stream);

/*
 * sim_matrix (model/model.py:582-590): out[i][j] = a_i . b_j / (max(|a_i|,eps) max(|b_j|,eps)).
 * a (N, d), b (M, d) fp32 or bf16, out (N, M) fp32.  Backward: given d_out produces d_a, d_b
 * (either may be NULL) in `out_dtype` (DVLP_F32 always; DVLP_BF16 for bf16 inputs with d == 256, else
 * DVLP_ERR_BAD_DTYPE).  `scratch` holds N + M floats (inverse norms).
 */
int dvlp_sim_matrix_fwd(const void* a, const void* b, int N, int M, int d, float eps, int in_dtype,
                        float* out, float* scratch, void* stream);
int dvlp_sim_matrix_bwd(const void* a, const void* b, int N, int M, int d, float eps, int in_dtype,
                        const float* d_out, void* d_a, void* d_b, int out_dtype, float* scratch, void* stream);

/*
 * NormSoftmaxLoss (model/loss.py:126-138) on x (N, M) fp32:
 *   loss = -mean diag(log_softmax(x/T, 1)) - mean diag(log_softmax(x^T/T, 1)).
 * loss: 1 float, overwritten.  d_x (N, M) or NULL: d loss / d x.
 * `scratch` holds N + M floats (row / column log-sum-exp).
 */
int dvlp_normsoftmax_fwd_bwd(const float* x, int N, int M, float temperature,
                             float* loss, float* d_x, float* scratch, void* stream);

/*
 * Projection heads feeding the path: Y = act(X) W^T + b as one tcgen05 GEMM writing the bf16 rows the
 * scoring kernels read.  Replaces txt_proj = Sequential(ReLU(), Linear(768, 256)) (model/model.py:40-43, applied at :88;
 * relu_in = 1, bias given) and ObjectTransformer.proj = Linear(768, 256, bias=False) (model/object_transformer.py:381,
 * :451; relu_in = 0, bias NULL).
 *   x (M, K) fp32 or bf16 token rows (K a multiple of 64), w_bf16 (N = 256, K) bf16 row-major (nn.Linear.weight),
 *   bias (256) fp32 or NULL, out_bf16 (M, 256) bf16, overwritten.  Forward only: the host mirror's backward
 *   (demovlp_b200/proj.py) uses library GEMMs.
 */
size_t dvlp_proj_workspace_bytes(int M, int K, int in_dtype, int relu_in);
int dvlp_proj_fwd(const void* x, int in_dtype, int M, int K, int N, const void* w_bf16, const float* bias, int relu_in,
                  void* out_bf16, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Masked attention over the short token sequences of the ObjectTransformer:
 *   out = softmax(q k^T + mask, dim=-1) v       -- attn_mask of model/object_transformer.py:93-100, called by
 * VarAttention.forward (:125-196) on (batch x heads) sequences (space: 30 x 31 tokens, time: 8 x 9, CLS: 1 x (1 + F R)).
 *   q (B, I, d), k / v (B, J, d), mask (B, J) additive or NULL, out (B, I, d), lse (B, I) or NULL; all fp32, d = 64.
 *   dvlp_attn_small_supported: 1 = tensor-core kernels (I, J <= 64; fp32 operands split into bf16 hi + lo parts, so the
 *   result has fp32-class accuracy), 2 = single-query kernels (I == 1, J <= 1024), 0 = not served.
 *   One CTA per sequence; q, k, v, out are touched once.  The backward needs `out` and `lse` of the forward and
 *   overwrites d_q, d_k, d_v (no atomics: deterministic).
 */
int dvlp_attn_small_supported(int I, int J, int d);
int dvlp_attn_small_fwd(const float* q, const float* k, const float* v, const float* mask, int B, int I, int J, int d,
                        float* out, float* lse, void* stream);
int dvlp_attn_small_bwd(const float* q, const float* k, const float* v, const float* mask, const float* out,
                        const float* lse, const float* d_out, int B, int I, int J, int d, float* d_q, float* d_k,
                        float* d_v, void* stream);

/*
 * Retrieval ranks (model/metric.py).  sims is the (Nq, Nv) float64 matrix the
 * trainer builds (trainer/trainer_dist.py:369-385), rows = text queries.
 *   t2v (metric.py:22-66):  ranks[i] = #{ j : sims[i][j] > sims[i][gt(i)] },  gt(i) = i / (Nq / Nv)
 *   v2t (metric.py:142-189): for video v, over its captions c: #greater + (#equal-1)/2 in column v,
 *                            minimum over c; ranks (Nv) float64.
 * Integer / comparison work only: bit-exact with the reference.
 */
int dvlp_ranks_t2v(const double* sims, int Nq, int Nv, int32_t* ranks, void* stream);
int dvlp_ranks_v2t(const double* sims, int Nq, int Nv, double* ranks, void* stream);

/*
 * Peer-memory exchanges of the video-row sharded loss (one process per GPU of one NVLink / NVSwitch box).
 * They carry what the reference's AllGather_multi pattern (trainer/trainer_dist.py:13-31; all_gather forward,
 * own-slice backward) moves when the global batch is sharded by video rows: the all-gather of the text side,
 * the sum of the row-local RWALoss partials (model/loss.py:108-114) and the reduce-scatter of the text gradients.
 *
 * Every rank owns an ARENA of the same size and layout whose first dvlp_p2p_control_bytes() bytes are the
 * library's control block (zero it once, then barrier, before the first call).  `peer_arenas` is a DEVICE array
 * of `world` pointers: this process's mappings of the arenas of ranks 0..world-1 (peer_arenas[rank] is its own),
 * e.g. the buffer_ptrs of torch symmetric memory or cudaIpcOpenMemHandle results.  All ranks must issue the same
 * sequence of these calls.  The calls only enqueue kernels on `stream` (CUDA-graph capturable); flags carry an
 * epoch kept in the control block.
 *
 *   dvlp_p2p_allgather       for t < n_tensors: stores this rank's shard (shards_host[t], shard_bytes_host[t] bytes,
 *                            multiple of 16) at arena_offsets_host[t] + rank * shard_bytes in EVERY arena and returns
 *                            (in stream order) once all ranks' shards have landed in this rank's arena.
 *                            *_host arguments are HOST arrays (n_tensors <= 8).
 *   dvlp_p2p_sum_scalar      out = sum over ranks of *value, added in rank order (same bits on every rank).
 *   dvlp_p2p_reduce_scatter  every arena holds, at arena_offset, world * shard_floats fp32 partial sums (completed
 *                            by earlier work of `stream`); out (shard_floats elements of out_dtype) = sum over ranks
 *                            of rows [rank * shard_floats, (rank + 1) * shard_floats), added in rank order.
 * A data region may be overwritten again only after a later exchange of the same call sequence has completed
 * (the sharded loss alternates allgather -> sum_scalar -> reduce_scatter).
 * Flag waits are bounded by wall time: DVLP_P2P_TIMEOUT_S seconds (environment, default 600).  A wait that gives up
 * does NOT trap: it records a sticky error word in this rank's control block, the kernels finish with meaningless
 * data, and dvlp_p2p_check (which synchronises `stream`) returns DVLP_ERR_TIMEOUT from then on.
 */
size_t dvlp_p2p_control_bytes(void);
int dvlp_p2p_check(const void* own_arena, void* stream);
int dvlp_p2p_allgather(void* const* peer_arenas, int world, int rank, int n_tensors, const void* const* shards_host,
                       const size_t* shard_bytes_host, const size_t* arena_offsets_host, void* stream);
int dvlp_p2p_sum_scalar(void* const* peer_arenas, int world, int rank, const float* value, float* out, void* stream);
int dvlp_p2p_reduce_scatter(void* const* peer_arenas, int world, int rank, size_t arena_offset, size_t shard_floats,
                            void* out, int out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEMOVLP_B200_H_ */

/* tribe_b200.h — C ABI of libtribe_b200.so: the B200 (sm_100a) kernels behind TRIBE's FmriEncoder hot path.
 *
 * The reference (vovw/algonauts-2025) is pure Python/PyTorch and has no FFI; every entry point below replaces one
 * implicit ATen/cuBLAS call sequence of the reference (cited per function, paths relative to the reference root).
 * Conventions:
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless the name says host; no torch types;
 *   - buffers are owned by the caller (torch's caching allocator on the Python side); kernels never allocate;
 *   - every call is asynchronous on the `stream` argument (a cudaStream_t passed as void*);
 *   - return value: 0 on success, a negative TRIBE_E* code or a positive cudaError_t otherwise; no exceptions cross
 *     the ABI; tribe_last_error() returns a static description of the last failure on the calling thread;
 *   - bf16 buffers are `uint16_t`-sized elements (__nv_bfloat16), row-major unless stated.
 */
#ifndef TRIBE_B200_H_
#define TRIBE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRIBE_OK 0
#define TRIBE_EINVAL (-1)   /* bad argument (shape / alignment / unsupported combination) */
#define TRIBE_EDRIVER (-2)  /* CUDA driver entry point (cuTensorMapEncodeTiled) unavailable */
#define TRIBE_ETMAP (-3)    /* tensor-map encoding failed */

/* element types of `*_dtype` arguments */
#define TRIBE_DT_F32 0
#define TRIBE_DT_F64 1
#define TRIBE_DT_BF16 2
#define TRIBE_DT_F16 3

/* point-wise training losses of tribe_point_loss_fwd_bwd (MSE has its own entry point) */
#define TRIBE_LOSS_SMOOTH_L1 1 /* torch.nn.SmoothL1Loss(beta = param)  */
#define TRIBE_LOSS_HUBER 2     /* torch.nn.HuberLoss(delta = param)    */
#define TRIBE_LOSS_L1 3        /* torch.nn.L1Loss                      */

const char* tribe_last_error(void);
int tribe_abi_version(void);
/* Number of kernels launched through this library since load (bench.py's `gpu_launches`). */
int64_t tribe_launch_count(void);
/* Programmatic dependent launch for the hot kernels (GEMMs, attention, norm / sub-layer tails, column sums, Adam): a
 * kernel's prologue and launch latency overlap its predecessor's tail; stream order is preserved (every CTA waits for
 * the predecessor's completion before touching memory).  Off by default (measured neutral inside whole-step CUDA
 * graphs); TRIBE_PDL=1 in the environment or this call turns it on. */
int tribe_set_pdl(int32_t on);
/* cudaPeekAtLastError() of the calling thread's context (debugging aid: TRIBE_DEBUG_LASTERR=1 makes the Python layer
 * check it after every call, naming the first entry point after which the runtime's error state is set). */
int tribe_peek_last_error(void);
/* cudaGetLastError(): returns AND clears the runtime's non-sticky last-error slot.  Every entry point of this library
 * consumes the error of its own launches; a value found here was left behind by some other user of the runtime, and
 * torch would attribute it to whatever call it checks next (seen once in ~10 runs of the test suite as "invalid
 * argument" at an unrelated .item()).  The Python layer drains it before its own synchronisation points. */
int tribe_take_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * tcgen05 / TMEM / TMA GEMM:  D[z] = epilogue( alpha * A[z] (M x K) * B[z]^T (N x K) )
 * Replaces every dense contraction of the path: projector nn.Linear (algonauts2025/model.py:157), the
 * x_transformers Encoder's to_q/to_k/to_v/to_out/ff linears and attention einsums (model.py:173), the SubjectLayers
 * gather + einsum (modeling_utils/modeling_utils/models/common.py:61-66), InfoNCE logits (model.py:216), and their
 * autograd backward GEMMs.
 *
 * Operands are bf16 in global memory described as 3-D tensors (inner, row, batch):
 *   K-major  operand (mn_major=0): inner = K (contiguous), row = M or N index.
 *   MN-major operand (mn_major=1): inner = M or N (contiguous), row = K index.
 * `inner_off` / `zin_stride` shift the inner coordinate (e.g. head h of a packed [tokens, 3*H*dh] qkv buffer);
 * batch coordinate of z is  gather ? gather[z / zdiv] : z / zdiv.
 */
typedef struct TribeOperand {
  const void* ptr;       /* bf16 */
  int64_t inner;         /* extent of the contiguous dimension (elements) */
  int64_t rows;          /* extent of dim 1 */
  int64_t batch;         /* extent of dim 2 (>= 1) */
  int64_t row_stride;    /* elements; *2 bytes must be a multiple of 16 */
  int64_t batch_stride;  /* elements; *2 bytes must be a multiple of 16 (ignored when batch == 1) */
  int32_t mn_major;      /* 0: K-major, 1: MN-major */
  int32_t inner_off;     /* added to the inner coordinate */
  int32_t zin_stride;    /* inner coordinate += (z % z_inner) * zin_stride */
  int32_t zdiv;          /* batch coordinate = z / zdiv (>= 1) */
  const int64_t* gather; /* optional: batch coordinate = gather[z / zdiv] (subject ids, common.py:61) */
} TribeOperand;

enum TribeEpilogue {
  TRIBE_EPI_STORE = 0,     /* D = alpha*acc (+bias)                                                     */
  TRIBE_EPI_GELU = 1,      /* aux_out = bf16(acc+bias) (skipped when aux_out == NULL: inference);
                              D = gelu_erf(acc+bias)                                   (FF up-projection) */
  TRIBE_EPI_RESIDUAL = 2,  /* D = acc (+bias) + res[row % res_row_mod] * (rscale ? rscale[col] : 1)      */
  TRIBE_EPI_GELU_BWD = 3,  /* D = acc * gelu_erf'(aux_in)                                                */
  TRIBE_EPI_ROPE = 4       /* D = rotate(acc): interleaved-pair rotary on the first rope_dim dims of every
                              head_dim-wide head for columns < rope_cols; rope_sign=-1 applies the transpose */
};

typedef struct TribeGemm {
  TribeOperand a, b;
  int32_t m, n, k;        /* per-batch problem size */
  int32_t batch;          /* number of z */
  int32_t z_inner;        /* z = zo * z_inner + zi (>= 1) */
  /* K-gather (grouped wgrad of SubjectLayers): when kgroup != NULL the contraction runs over kgroup_len batches of
   * K each, taking batch j only if kgroup[j] == z; operands' batch coordinate is then j (zdiv/gather ignored). */
  const int64_t* kgroup;
  int32_t kgroup_len;
  /* output */
  void* d;
  int32_t d_f32;          /* 0: bf16, 1: fp32 */
  int32_t d_transposed;   /* 0: d[row*ldd + col], 1: d[col*ldd + row] */
  int64_t ldd, d_zo_stride, d_zi_stride; /* elements */
  /* epilogue */
  int32_t epilogue;       /* enum TribeEpilogue */
  float alpha;
  const float* bias;      /* optional fp32 [n] (+ bias_z_stride * batch coordinate of B when bias_gathered) */
  int32_t bias_gathered;
  int64_t bias_z_stride;
  const float* res;       /* fp32 residual, row-major ld_res */
  int64_t ld_res;
  int32_t res_row_mod;    /* residual row = row % res_row_mod (0: row) */
  int32_t res_batched;    /* 1: the residual shares the output's batch offsets (in-place accumulation D += acc) */
  const float* rscale;    /* optional fp32 [n] */
  const void* aux_in;     /* bf16 [m, ld_aux] (GELU_BWD) */
  void* aux_out;          /* bf16 [m, ld_aux] (GELU) */
  int64_t ld_aux;
  const float* rope;      /* fp32 (cos, sin) pairs [rope_t][rope_dim/2][2] */
  int32_t rope_t;         /* position = row % rope_t */
  int32_t rope_dim, head_dim, rope_cols;
  float rope_sign;
  int32_t block_n;        /* 0: auto; else one of 128, 160, 192, 256 */
  /* optional split-K workspace (caller-owned, its first 2560 bytes ZERO-initialised once; 20 MiB is plenty): lets the
   * ragged last wave of tiles be split along K over otherwise idle SMs (per-slice fp32 partial slots + arrival
   * counters, which the kernel resets).  Must not be shared by GEMMs running concurrently on different streams. */
  void* splitk_ws;
  int64_t splitk_ws_bytes;
  /* optional optimizer step fused into the epilogue (weight-gradient GEMMs; replaces the `optimizer.step()` pass of
   * Lightning's automatic optimisation for this weight, main.py:396-413 / defaults.py:126-141): with adam_p != NULL the
   * finished fp32 gradient element (after the epilogue, e.g. RESIDUAL with res = D for accumulation) at offset
   * o = batch offset + row*ldd + col never has to reach memory: the epilogue reads adam_p/m/v[o], applies one Adam step
   * with the device hyper-parameter block adam_hyper (tribe_adam_hyper layout) and writes adam_p/m/v[o] and the bf16
   * shadow adam_shadow[o].  D itself is written only when adam_keep_grad != 0.  Requires d_f32 = 1, d_transposed = 0. */
  float* adam_p;
  float* adam_m;
  float* adam_v;
  void* adam_shadow;        /* bf16, optional */
  const float* adam_hyper;
  int32_t adam_keep_grad;
} TribeGemm;

/* Kernel selection (all instances compute the same function): 2-CTA pairs (256 x 256 tiles) for large single problems;
 * the 1-CTA kernel (128 x BN tiles) for grouped / small problems; batched problems with 2-3 row tiles, <= 8 k-blocks and
 * n % 128 == 0 (the attention contractions P.V, dV, dQ, dK) run on the B-stationary multi-row-tile kernel
 * (csrc/gemm_mt_sm100.cuh; results bit-identical to the 1-CTA kernel; a non-zero block_n or TRIBE_GEMM_MT=0 selects the
 * 1-CTA kernel).  GELU / GELU_BWD evaluate Phi(x) from one exponential (|error| <= 5e-7 absolute). */
int tribe_gemm_bf16(const TribeGemm* g, void* stream);

/* Descriptor-probe variant used by tests only: overrides the UMMA shared-memory descriptor byte offsets
 * (lbo/sbo for K-major and MN-major operands); pass 0 to keep the defaults. */
/* Attention scores with the row softmax fused into the tcgen05 epilogue (whole score rows of <= 320 keys live in TMEM;
 * x_transformers attention behind model.py:173, restated in oracle/xt_encoder.py).  a / b: bf16 row-major (n_batch * T,
 * ld) matrices whose head h occupies columns [off + h*dh, off + (h+1)*dh); z = batch * heads + head.
 *   mode 0: out[z, i, :] = softmax_j(scale * <a_i, b_j>)  (bf16 (Z, T, Tp), columns >= T zeroed)      a = q, b = k
 *   mode 1: out[z, i, j] = P[z,i,j] * (<a_i, b_j> - sum_j' <a_i, b_j'> P[z,i,j']) * scale              a = dO, b = v, P = p_in
 * Requires dh % 64 == 0, T <= Tp <= 320, Tp % 8 == 0. */
int tribe_attn_scores(const void* a, int64_t a_ld, int64_t a_off, const void* b, int64_t b_ld, int64_t b_off, int64_t n_batch, int64_t T,
                      int64_t heads, int64_t dh, float scale, int32_t mode, const void* p_in, void* out, int64_t Tp, void* stream);
/* Flash-style attention forward, one launch per layer: O = softmax(scale * q k^T) v per (batch, head) with the scores and
 * the probabilities never leaving the SM (S in TMEM, P as the shared-memory A operand of P.V); replaces tribe_attn_scores
 * (mode 0) + the batched P.V GEMM.  q / k / v: bf16 row-major (n_batch * T, ld) with head h at columns [off + h*dh, ...);
 * p_out: optional bf16 (n_batch*heads, T, Tp) copy of P (the training backward reads it; NULL for inference);
 * o_out: bf16 (n_batch * T, o_ld), head h at columns [o_off + h*dh, ...).
 * Requires 160 < T <= Tp <= 320, Tp % 8 == 0, dh a multiple of 64 that is <= 256 or twice such a value (<= 448). */
int tribe_attn_fwd(const void* q, int64_t q_ld, int64_t q_off, const void* k, int64_t k_ld, int64_t k_off, const void* v, int64_t v_ld,
                   int64_t v_off, int64_t n_batch, int64_t T, int64_t heads, int64_t dh, float scale, void* p_out, int64_t Tp, void* o_out,
                   int64_t o_ld, int64_t o_off, void* stream);
/* Persistent GEMM grids use at most n_sms SMs from the next launch on (0 = all).  While a collective's CTAs occupy SMs
 * next to the backward pass, a one-CTA-per-SM grid would wait a whole kernel for the occupied SMs; a smaller grid runs
 * beside them (parallel.StepOverlap sets / clears this around the gradient all-reduces). */
int tribe_gemm_set_sm_limit(int32_t n_sms);
int tribe_gemm_bf16_probe(const TribeGemm* g, void* stream, uint32_t k_lbo, uint32_t k_sbo, uint32_t mn_lbo, uint32_t mn_sbo);

/* ------------------------------------------------------------------------------------------------------------------
 * Bandwidth kernels
 */

/* Feature ingest (model.py:147-155): x (B, L, D, T) fp32/fp64/bf16, T contiguous  ->  out bf16 (B*T, ld_out) with
 * out[b*T+t, col_off + l*D + d] = x[b,l,d,t]  (layer_mean=0, "cat")  or  mean_l x[b,l,d,t] at col_off + d (layer_mean=1).
 * src_dtype: 0 fp32, 1 fp64, 2 bf16, 3 fp16. */
int tribe_ingest_features(const void* x, int32_t src_dtype, int64_t B, int64_t L, int64_t D, int64_t T, int32_t layer_mean,
                          void* out_bf16, int64_t ld_out, int64_t col_off, void* stream);

/* ScaleNorm forward (x_transformers ScaleNorm, see oracle/xt_encoder.py): y = x / max(||x||, eps) * gain_mult * g[0].
 * x fp32 (rows, dim) -> y bf16 (rows, dim), rnorm fp32 (rows) = 1 / max(||x||, eps).
 * gain_mult = 0 -> sqrt(dim), eps = 0 -> 1e-12: the >= 2.x form F.normalize(x) * sqrt(dim) * g;
 * gain_mult = 1, eps = 1e-5: the 1.27.x form x / norm.clamp(min = eps) * g. */
int tribe_scalenorm_fwd(const float* x, const float* g, void* y_bf16, float* rnorm, int64_t rows, int64_t dim, float gain_mult, float eps,
                        void* stream);

/* Sub-layer backward tail: given dy_out (fp32 grad wrt the sub-layer output x_out = branch + x_in*rs) and d_xn (bf16
 * grad wrt the ScaleNorm output), produces dx_in (fp32 and bf16 copies), and accumulates d_rs (fp32 [dim]) and d_g
 * (fp32 [1]) atomically.  d_xn may be NULL (final norm handled by tribe_scalenorm_bwd). rs may be NULL (=1). */
int tribe_sublayer_bwd(const float* dy_out, const void* d_xn_bf16, const float* x_in, const float* rnorm, const float* g,
                       const float* rs, float* dx_in, void* dx_in_bf16, float* d_rs, float* d_g, int64_t rows, int64_t dim,
                       float gain_mult /* as in tribe_scalenorm_fwd */, void* stream);

/* Half-split rotary embedding of x_transformers 1.27.x, in place on a bf16 (rows, ld) buffer: for each of the n_heads
 * heads at columns col_off + h*head_dim, dims i and i + rot_dim/2 (i < rot_dim/2) are rotated by the angle of position
 * row % T and frequency i (table: fp32 (T, rot_dim/2, 2) = (cos, sin), the tribe_gemm rope table); sign = -1 applies
 * the transpose (backward).  The >= 2.x interleaved pairing is fused into the GEMM epilogue instead (TRIBE_EPI_ROPE). */
int tribe_rope_half(void* x_bf16, int64_t rows, int64_t ld, int64_t col_off, int64_t n_heads, int64_t head_dim, int64_t rot_dim,
                    const float* table, int64_t T, float sign, void* stream);

/* Row softmax over the first n_valid of ld columns: s fp32 (rows, ld) -> p bf16 (rows, ld), columns >= n_valid zeroed. */
int tribe_softmax_fwd(const float* s, void* p_bf16, int64_t rows, int64_t n_valid, int64_t ld, void* stream);
/* ds = p * (dp - sum_j p_j dp_j) * scale -> bf16 (rows, ld), padded columns zeroed. */
int tribe_softmax_bwd(const void* p_bf16, const float* dp, void* ds_bf16, float scale, int64_t rows, int64_t n_valid, int64_t ld,
                      void* stream);

/* Column sums (bias / residual-scale gradients): out[c] (+)= sum_r x[r, c] (* y[r, c] when y != NULL).
 * x_dtype/y_dtype: 0 fp32, 2 bf16.  accumulate: 0 overwrite (out must be zeroed by the call), 1 add. */
int tribe_colsum(const void* x, int32_t x_dtype, const void* y, int32_t y_dtype, float* out, int64_t rows, int64_t cols,
                 int64_t ld, int32_t accumulate, void* stream);

/* fp32 -> bf16 cast of a flat parameter buffer (bf16 shadow weights after the optimizer step). */
int tribe_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* dst_f32[i] = a * src[i] (+ dst when accumulate) — gradient bucket scaling / accumulation. */
int tribe_axpby_f32(const float* src, float* dst, float a, int32_t accumulate, int64_t n, void* stream);

/* AdaptiveAvgPool1d over the LAST dim (nn.AdaptiveAvgPool1d, model.py:60,119-122): x fp32 (rows, t_in) ->
 * y fp32 (rows, t_out), window i = [floor(i*t_in/t_out), ceil((i+1)*t_in/t_out)). */
int tribe_adaptive_avg_pool_fwd(const float* x, float* y, int64_t rows, int64_t t_in, int64_t t_out, void* stream);
int tribe_adaptive_avg_pool_bwd(const float* dy, float* dx, int64_t rows, int64_t t_in, int64_t t_out, void* stream);
/* Same windows along the TOKEN dim of a token-major activation: x bf16 (B, t_in, C) -> y bf16 (B, t_out, C)
 * (pool-before-readout: pool and SubjectLayers commute). */
int tribe_token_pool_fwd(const void* x_bf16, void* y_bf16, int64_t B, int64_t t_in, int64_t t_out, int64_t C, void* stream);
/* dx (B, t_in, C) = scatter of dy (B, t_out, C) / window length; dtypes 0 fp32 / 2 bf16 on either side. */
int tribe_token_pool_bwd(const void* dy, int32_t dy_dtype, void* dx, int32_t dx_dtype, int64_t B, int64_t t_in, int64_t t_out,
                         int64_t C, void* stream);

/* out[r, c] = (x ? x[r, c] : 0) + pos[r % row_mod, c], c < cols: the `x + time_pos_embed[:, :T]` of model.py:169-170 for
 * the stand-alone transformer_forward entry and for column blocks of dropped / absent modalities (model.py:143-144,158-159). */
int tribe_add_rows_periodic(const float* x, int64_t ld_x, const float* pos, int64_t ld_pos, float* out, int64_t ld_out, int64_t rows,
                            int64_t cols, int64_t row_mod, void* stream);

/* (B, O, T) fp32 -> (B, T, O) bf16 transpose-cast (gradient of the predictions into the readout's GEMM layout). */
int tribe_transpose_cast_bot(const float* x, void* y_bf16, int64_t B, int64_t O, int64_t T, void* stream);

/* SubjectLayers bias gradient: d_bias[subject[b], o] += sum_t dy[b, t, o]   (dy bf16 (B, T, O)). */
int tribe_subject_bias_grad(const void* dy_bf16, const int64_t* subjects, float* d_bias, int64_t B, int64_t T, int64_t O,
                            int64_t n_subjects, void* stream);

/* max(subjects) >= n_subjects check of SubjectLayers.forward (common.py:53-55) without a host sync on the hot path:
 * writes 1 into *flag_out (device int32) when violated (never clears it: the flag is sticky).  clamped_out (optional,
 * int64[n]) receives the ids clamped into [0, n_subjects): when the host examines the flag one step late (training
 * loops, CUDA-graph replays) the gathers of the already enqueued step stay inside the weight tensors. */
int tribe_check_subjects(const int64_t* subjects, int64_t n, int64_t n_subjects, int32_t* flag_out, int64_t* clamped_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Loss / evaluation reductions
 */

/* nn.MSELoss (pl_module.py:56 via losses/base.py:43-59), fused forward + gradient:
 * loss_out[0] = mean((pred-target)^2);  grad (optional, fp32, same shape) = grad_scale * 2 (pred-target) / n.
 * partial: workspace of >= 1024 doubles. */
int tribe_mse_fwd_bwd(const float* pred, const float* target, float* loss_out, float* grad, float grad_scale, int64_t n,
                      double* partial, void* stream);

/* Per-parcel Pearson sufficient statistics over rows of (n_rows, n_parcels)-shaped matrices (evaluation path:
 * main.py:459-477 scipy loop and pl_module.py:93-106 torchmetrics update).  The matrices are addressed as
 * element(r, p) = base[(r / t_len) * stride_b + p * stride_p + (r % t_len) * stride_t], which covers both the
 * flattened (b t) x d view of a (B, D, T) tensor (stride_b=D*T, stride_p=T, stride_t=1, t_len=T) and a plain
 * row-major (N, O) matrix (t_len=1, stride_b=O, stride_p=1).
 * stats: fp64 [6][n_parcels] = n, sum_x', sum_y', sum_x'x', sum_y'y', sum_x'y' with x' = x - shift[p], y' = y -
 * shift[n_parcels + p] (shift: fp32 [2][n_parcels] per-parcel pivots, NULL = 0).  scipy / torchmetrics centre before
 * they multiply; raw fp32 moments lose the variance once |mean| >> sigma, pivoted ones do not, and r is invariant to the
 * pivot.  All calls that accumulate into one stats block must pass the SAME pivots (tribe_pearson_pick_shift takes them
 * from the first row; tribe_pearson_recenter re-expresses a block about other pivots, e.g. before merging ranks).
 * Accumulated (+=) so that batches can be streamed; group: optional int64 per row/t_len block selecting
 * stats + group*6*n_parcels (GroupedMetric, metrics/base.py:52-78). */
int tribe_pearson_stats(const float* pred, const float* target, int64_t n_rows, int64_t n_parcels, int64_t t_len,
                        int64_t stride_b, int64_t stride_p, int64_t stride_t, const int64_t* group, int64_t n_groups,
                        const float* shift, double* stats, void* stream);
/* shift[p] = pred[p * stride_p], shift[n_parcels + p] = target[p * stride_p]: the first sample of every parcel (exactly
 * the column value for a constant column, so its shifted sums are exactly 0 and r = NaN like scipy.stats.pearsonr). */
int tribe_pearson_pick_shift(const float* pred, const float* target, int64_t n_parcels, int64_t stride_p, float* shift, void* stream);
/* In place: statistics taken about pivots shift_old -> the same statistics about shift_new (either may be NULL = 0), fp64. */
int tribe_pearson_recenter(double* stats, int64_t n_groups, int64_t n_parcels, const float* shift_old, const float* shift_new, void* stream);
/* r[p] = clamp(cov / sqrt(var_x var_y), -1, 1) from the statistics (NaN for a constant column or NaN input, like
 * scipy / torchmetrics; n = 1 gives NaN where scipy raises); r fp32 [n_parcels]; mean_out optional fp32 [1]. */
int tribe_pearson_finalize(const double* stats, int64_t n_parcels, float* r, float* mean_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Contrastive branch: symmetric InfoNCE over flattened (B*T) rows (algonauts2025/model.py:208-221).  The logits
 * (n, ld) fp32 = q^ k^T / tau come from tribe_gemm_bf16 on row-normalised operands; |logit| <= shift = 1/tau.
 */
/* row_sum[i] = sum_j exp(l_ij - shift); col_sum[j] = sum_i exp(l_ij - shift) (col_sum is zeroed by the call). n <= 8192. */
int tribe_nce_expsums(const float* logits, int64_t n, int64_t ld, float shift, float* row_sum, float* col_sum, void* stream);
/* loss = 0.5 * (CE(logits, arange) + CE(logits^T, arange)) from the sums above. */
int tribe_nce_loss(const float* logits, int64_t n, int64_t ld, float shift, const float* row_sum, const float* col_sum, float* loss_out,
                   void* stream);
/* g[i, j] = scale * upstream[0] * (softmax_row(i)_j + softmax_col(j)_i - 2 delta_ij) as bf16 (n, ldg), pad columns zeroed. */
int tribe_nce_grad(const float* logits, int64_t n, int64_t ld, float shift, const float* row_sum, const float* col_sum,
                   const float* upstream, float scale, void* g_bf16, int64_t ldg, void* stream);
/* bf16 -> fp32 cast (latents handed back to torch as fp32, model.py:178-183). */
int tribe_cast_bf16_f32(const void* src_bf16, float* dst, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Optimizer: one fused torch.optim.Adam step (amsgrad=False, maximize=False; recipe algonauts2025/grids/defaults.py:
 * 126-141) over a flat fp32 range, writing the bf16 shadow weights in the same pass (p_bf16 may be NULL).
 * `step` is the 1-based step count used for the bias corrections.  max_blocks > 0 bounds the grid (a step that runs
 * beside the backward GEMMs on a side stream should leave SM slots to them); 0 = fill the device.
 */
int tribe_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, double lr, double beta1, double beta2, double eps,
                    double weight_decay, int64_t step, int32_t max_blocks, void* stream);
/* The same step with its scalars read from DEVICE memory — hyper[6] = {beta1, beta2, lr / (1 - beta1^step),
 * 1 / sqrt(1 - beta2^step), eps, weight_decay} — so that a captured CUDA graph of the whole train step replays with the
 * scheduler's current lr / momentum (OneCycleLR cycles both per batch).  tribe_adam_hyper fills such a block from the
 * host-side values (one 1-block launch; the bias corrections are computed in fp64 on the host like torch does). */
int tribe_adam_step_dev(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, const float* hyper, int32_t max_blocks,
                        void* stream);
int tribe_adam_hyper(float* hyper_dev, double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step, void* stream);
/* The same for n blocks in one launch: block i lives at hyper_base + 8 * slots_host[i]; all arrays are HOST arrays of n
 * entries (read before the call returns). */
int tribe_adam_hyper_batch(float* hyper_base, const int32_t* slots_host, const int64_t* steps_host, const double* lr_host,
                           const double* beta1_host, const double* beta2_host, const double* eps_host, const double* wd_host, int32_t n,
                           void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Data-parallel step tail over NVLink 5 / NVSwitch.  Replaces Lightning DDP's gradient all-reduce + the replicated
 * optimizer step (algonauts2025/main.py:388-394, strategy "ddp_find_unused_parameters_true"; Adam recipe
 * algonauts2025/grids/defaults.py:126-141).  The flat fp32 gradient buffer, the bf16 shadow and the fp32 masters of every
 * rank live in SYMMETRIC memory (same size on every rank, peer-mapped, with an NVLS multicast mapping when the fabric
 * offers one; allocated and exchanged by the host side, parallel.ShardedStep).
 *
 * tribe_sharded_adam_step: ONE kernel for the [lo, lo + n) range THIS rank owns —
 *   g = (sum over ranks of grad[i]) / world   multimem.ld_reduce.add.f32 on grad_mc (in-switch reduction), or, when
 *                                             grad_mc == NULL, loads from grad_peer[0..world) summed in rank order
 *                                             (peer mappings, or local staging buffers filled by tribe_memcpy_async);
 *   Adam on the local param / m / v (same arithmetic as tribe_adam_step_dev, scalars from the device block `hyper`);
 *   bf16(param) -> every rank's shadow         multimem.st on shadow_mc, or (shadow_mc == NULL) stores to shadow_peer[r];
 *   bcast_master != 0: the fp32 param too      (parameters the kernels read as fp32: biases, gains, residual scales,
 *                                              positional embedding) — param_mc / param_peer.
 * All pointers already include the offset `lo`; n is a multiple of 8 and every pointer 16-byte aligned.  m / v (and the
 * fp32 master of ranges without bcast_master) are only current on the owner (gathered on demand for checkpoints).
 * max_blocks bounds the grid (default 888 = 6 CTAs of 128 threads per SM: the kernel runs after the backward pass with
 * the device to itself — a resident foreign CTA keeps the persistent 2-CTA GEMM off its SM, profiles/r02_coresidency.log).
 *
 * tribe_xgpu_barrier: all `world` ranks meet at `slot` (< TRIBE_XGPU_SLOTS).  flags->ptr[r] = rank r's flag block
 * (TRIBE_XGPU_SLOTS * TRIBE_XGPU_MAX_WORLD zero-initialised uint32 words of symmetric memory) as mapped into THIS
 * process.  Release/acquire at system scope: writes made before the barrier by any rank's earlier kernels on the
 * calling stream are visible to every rank's later kernels.  A rank that waits longer than timeout_s (<= 0: 30 s) stores
 * 1 + slot into err_flag (device uint32, sticky) and proceeds instead of hanging the GPU.  Capturable in CUDA graphs.
 */
#define TRIBE_XGPU_MAX_WORLD 16
#define TRIBE_XGPU_SLOTS 64
typedef struct TribeXgpuPeers {
  void* ptr[TRIBE_XGPU_MAX_WORLD];
} TribeXgpuPeers;
typedef struct TribeShardedAdam {
  float* param;            /* local fp32 master, m, v of the owned range */
  float* m;
  float* v;
  const float* hyper;      /* device: {beta1, beta2, lr / (1 - beta1^step), 1 / sqrt(1 - beta2^step), eps, weight_decay} */
  const float* grad_mc;    /* multicast addresses of the owned range (NULL: use the *_peer tables) */
  void* shadow_mc;         /* bf16 */
  float* param_mc;
  TribeXgpuPeers grad_peer;   /* per-rank addresses of the owned range as mapped into this process */
  TribeXgpuPeers shadow_peer;
  TribeXgpuPeers param_peer;
  int64_t n;
  int32_t world, rank, bcast_master, max_blocks;
} TribeShardedAdam;
int tribe_sharded_adam_step(const TribeShardedAdam* a, void* stream);
/* Copy-engine transfer dst <- src (pointers of this process's unified address space: local device memory, peer-mapped
 * symmetric memory, or PINNED host memory — cudaMemcpyDefault).  The data-parallel step PUSHES each finished gradient
 * bucket's slices into their owners' staging buffers with it while the backward pass keeps every SM (grad_peer[] of
 * tribe_sharded_adam_step then points at the LOCAL staged copies); the subject-range flag of SubjectLayers
 * (common.py:53-55) travels to its pinned host word with it, also as a memcpy node of a captured step. */
int tribe_memcpy_async(void* dst, const void* src, int64_t n_bytes, void* stream);
/* Micro-benchmark of the pieces of the kernel above (tools/xgpu_probe.py; not used by the product path): mode 0 =
 * multimem.ld_reduce only, 1 = peer loads only, 2 = local param/m/v stream only, 3 = multimem.st only, 4 = ld_reduce x4. */
/* Debug: `blocks` CTAs of `threads` threads spin for `seconds` (tools/coresidency_probe.py). */
int tribe_debug_spin(int32_t blocks, int32_t threads, double seconds, int32_t carveout_pct, uint32_t* sink, void* stream);
int tribe_xgpu_probe(const TribeShardedAdam* a, int32_t mode, int32_t blocks, float* sink, void* stream);
int tribe_xgpu_barrier(const TribeXgpuPeers* flags, int32_t rank, int32_t world, int32_t slot, uint32_t* err_flag, double timeout_s,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Alternative training losses (grid: algonauts2025/grids/run_ensemble.py:29; built by modeling_utils/losses/base.py:
 * 43-59 from torch.nn, PearsonLoss from modeling_utils/losses/losses.py:11-42).  Same contract as tribe_mse_fwd_bwd:
 * reduction="mean", loss_out[0] fp32, grad (optional) = grad_scale * dloss/dpred, partial >= 1024 doubles.
 */
int tribe_point_loss_fwd_bwd(const float* pred, const float* target, float* loss_out, float* grad, int32_t kind, float param,
                             float grad_scale, int64_t n, double* partial, void* stream);
/* PearsonLoss(dim=1) = reduce_p (1 - pcc_p), pcc_p = cov / (std_x std_y + 1e-8) over all rows of parcel p.
 * Forward: tribe_pearson_stats into a ZEROED stats block (with the pivots `shift`, or NULL), then this finalize, which writes loss_out[0] and (optional)
 * coef fp32 [4][n_parcels] = mean_x, mean_y, 1/D, cov*std_y/(D^2 std_x) for the backward pass.
 * Backward: grad[i] = -(upstream[0] * (reduction_mean ? 1/n_parcels : 1)) * (coef2[p] (y_i - coef1[p]) - coef3[p] (x_i - coef0[p]))
 * with p = (i / t_len) % n_parcels over contiguous (N, O) (t_len = 1) or (B, O, T) (t_len = T) tensors; upstream is a
 * device scalar (autograd's grad_output) or NULL for 1. */
int tribe_pearson_loss_finalize(const double* stats, int64_t n_parcels, int32_t reduction_mean, const float* shift, float* coef,
                                float* loss_out, void* stream);
int tribe_pearson_loss_bwd(const float* pred, const float* target, const float* coef, const float* upstream, int32_t reduction_mean,
                           float* grad, int64_t n, int64_t n_parcels, int64_t t_len, void* stream);

/* dst[i] = src[i] * scalar_dev[0]: backward of the fused losses (autograd hands the upstream gradient as a device
 * scalar); tribe_memset_zero: stream-ordered zero fill of gradient slices / statistics blocks (no kernel). */
int tribe_scale_dev(const float* src, const float* scalar_dev, float* dst, int64_t n, void* stream);
int tribe_memset_zero(void* ptr, int64_t n_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * The steps either side of the path.
 */
/* Window assembly (data_utils/base.py:167-198 TimedArray._overlap_slice / __iadd__, data_utils/segments.py:144-180):
 * out fp32 (n_windows, rows, t_out); window b copies length[b] samples starting at src_start[b] of every row of its
 * timeline array src_ptrs[b] (row-major (rows, t_total[b]), f32 or f64, device pointers in a device array) to columns
 * [dst_start[b], dst_start[b] + length[b]) and zero-fills the rest.  The index triples are computed on the host with
 * the reference's own rounding rules (windows.py) and must satisfy 0 <= src_start, src_start + length <= t_total,
 * 0 <= dst_start, dst_start + length <= t_out. */
int tribe_gather_windows(const void* const* src_ptrs, int32_t src_dtype, const int64_t* t_total, const int32_t* dst_start,
                         const int32_t* src_start, const int32_t* length, float* out, int64_t n_windows, int64_t rows, int64_t t_out,
                         void* stream);
/* Ensemble averaging (algonauts2025/grids/average_submissions.py:107-125): weights from the members' per-parcel
 * validation Pearson r (n_members, n_parcels): axis = 1 -> w[m, :] = softmax over PARCELS of r[m, :] / temperature, which
 * is what the reference computes (`pearsons.softmax(dim=1)`, :108-109); axis = 0 -> softmax over MEMBERS per parcel
 * (weights of a parcel sum to one).  out[n, p] = sum_m w[m, p] * preds[m, n, p] over stacked (n_members <= 48, n_rows,
 * n_parcels) fp32 predictions (w == NULL: plain mean, :123). */
int tribe_ensemble_weights(const float* r, int64_t n_members, int64_t n_parcels, float temperature, int32_t axis, float* w, void* stream);
int tribe_ensemble_average(const float* preds, const float* w, int64_t n_members, int64_t n_rows, int64_t n_parcels, float* out,
                           void* stream);
/* Retrieval metric (modeling_utils/metrics/metrics.py:66-121, TopkAcc :194-218; pl_module.py:100-101 feeds it the
 * time-averaged (B, O) predictions / targets): y[row] = mean(x[row, :t]); ranks[b] of the true candidate among
 * scores[b, o] = <x_b, y_o> / (1e-15 + |y_o|), ties averaged, NaN comparisons false, negative -> n / 2.
 * scores_out (optional) fp32 (n, n). */
int tribe_mean_lastdim(const float* x, float* y, int64_t rows, int64_t t, void* stream);
int tribe_retrieval_ranks(const float* x, const float* y, int64_t n, int64_t c, float* ranks, float* scores_out, void* stream);
/* y (batch, cols, rows) = transpose of x (batch, rows, cols), fp32, bit-exact: the materialised "b d t -> (b t) d" of
 * pl_module.py:54-55 for losses / metrics without a fused kernel, and `pred.T` of the submission assembly
 * (algonauts2025/callbacks.py:66). */
int tribe_transpose_last2(const float* x, float* y, int64_t batch, int64_t rows, int64_t cols, void* stream);
/* Stochastic weight averaging over the flat parameter buffer (algonauts2025/main.py:365-373; torch AveragedModel):
 * avg += (params - avg) / (n_averaged + 1). */
int tribe_swa_update(float* avg, const float* params, int64_t n, int64_t n_averaged, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRIBE_B200_H_ */

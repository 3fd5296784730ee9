/* libcmpc_b200 -- C ABI of the B200-native CMPC head (cross-modal progressive comprehension).
 *
 * Drop-in boundary for the hot path of zigonk/CMPC-Refseg: LSTM_model.build_graph,
 * CMPC_model.py:89-142, i.e. everything between the DeepLab taps (:74-76) / the word LSTM outputs
 * (:153-156) and the mask logits pred/up/sigm (:140-142).  The reference has no FFI of its own (it is a
 * pure TF-1 graph executed by sess.run, trainval_model.py:232, test.py:286); these are the entry points a
 * binding for that path would need.  Each function cites the reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*.
 *   - the caller owns all memory (inputs, outputs, workspace); nothing is allocated or freed here.
 *   - every call is asynchronous and stream-ordered on `stream` (a cudaStream_t passed as void*);
 *     no internal synchronisation; re-entrant across streams.
 *   - return value: 0 = ok, negative = cmpc_status; cmpc_last_error() gives the message (thread-local).
 *   - sm_100 only; any other device returns CMPC_ERR_ARCH.  There is no CPU or alternate path.
 *   - activations are row-major [rows, channels] with channels fastest ("NHWC flattened"), fp16 where
 *     they feed tensor-core GEMMs (fp32 accumulate, fp32 epilogue math), fp32 elsewhere.
 */
#ifndef CMPC_B200_H_
#define CMPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CMPC_OK = 0,
  CMPC_ERR_ARG = -1,     /* bad shape / null pointer / unsupported size */
  CMPC_ERR_ALIGN = -2,   /* pointer or leading dimension not 16-byte aligned */
  CMPC_ERR_ARCH = -3,    /* current device is not sm_100 */
  CMPC_ERR_LAUNCH = -4,  /* CUDA launch / driver error (message holds the cudaError string) */
  CMPC_ERR_WORKSPACE = -5
} cmpc_status;

const char* cmpc_last_error(void);
int cmpc_version(void);
/* Programmatic dependent launch for the kernels that support it (off by default; CMPC_PDL=1 in the Python host): a kernel is then
 * launched with cudaLaunchAttributeProgrammaticStreamSerialization and waits (griddepcontrol.wait) for the kernel in front of it only
 * after its own prologue, so launch latency and prologue overlap the predecessor's tail.  Results are unchanged. */
void cmpc_set_pdl(int32_t on);

/* ------------------------------------------------------------------------------------------------
 * Dense 1x1-convolution GEMM with fused epilogue (tcgen05 + TMEM + TMA, persistent, warp-specialised).
 * Replaces LSTM_model._conv for filter_size 1 (CMPC_model.py:412-417) and the elementwise TF nodes that
 * follow each call site (bias add, relu, gates, norm statistics).
 *
 *   acc[m, n] = sum_k A1[m, k] * W[n, k]  (+ sum_k A2[m, k] * W[n, K1pad + k])      fp16 x fp16 -> fp32
 *   v = acc * rscale[m] + bias[n] + sbias[b(m), n] + peephole            b(m) = m / rows_per_sample, rscale = 1 without a_row_sumsq
 *   v = act(v) * gate[b(m), n]
 *   out[m, n] = v        (fp16 or fp32)
 *   row_sumsq[m] += sum_n v^2          stats[b(m), group(n)] += (sum v, sum v^2)   (fp64 atomics)
 *
 * W is [N, Kw] row-major (K contiguous; one row per output channel).  K extents are rounded up to
 * multiples of 64 inside W (K1pad = 64*ceil(K1/64)); columns of A beyond K1/K2 are never read (TMA
 * zero fill).  Column n is "valid" iff (n % group_width) < group_valid; invalid columns are written as 0
 * and excluded from statistics.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  /* operands */
  const void* a1;  int64_t lda1;  int32_t k1;   /* fp16 [M, k1], leading dim lda1 (elements) */
  const void* a2;  int64_t lda2;  int32_t k2;   /* optional second K-segment (NULL / 0 if unused) */
  const void* w;   int64_t ldw;                 /* fp16 [N, >= K1pad + K2pad] */
  int64_t w_batch_stride;                       /* 0: W shared; else per-sample W_b = w + b*stride (elements),
                                                   tiles never cross a sample (m % rows_per_sample == 0) */
  int32_t w_rows;                               /* physical rows of W (0 = n); rows beyond read as zero */
  int32_t m, n;
  int32_t rows_per_sample;                      /* b(m) = m / rows_per_sample (>= 1) */
  /* epilogue */
  const float* row_scale;                       /* reserved, must be NULL */
  const float* bias;                            /* [N] or NULL */
  const float* sbias;   int64_t ld_sbias;       /* [B, ld] per-sample bias or NULL */
  const float* gate;    int64_t ld_gate;        /* [B, ld] per-sample multiplicative gate or NULL */
  int32_t act;                                  /* 0 none, 1 relu, 2 tanh, 3 sigmoid */
  int32_t group_width, group_valid;             /* column validity / statistics groups (0 = one group of N) */
  /* ConvLSTM peepholes (util/cell.py:48-50): v += peep_g[pixel(m), c] * cprev[m, c] for group 1 and 2 */
  const float* peep_i;  const float* peep_f;  int64_t ld_peep;
  const float* cprev;   int64_t ld_cprev;
  /* outputs */
  void* out;  int64_t ldo;  int32_t out_fp32;   /* fp16 (0) or fp32 (1) */
  float* row_sumsq;                             /* [M] accumulated (caller zeroes) or NULL */
  double* stats;                                /* [B, n_groups, 2] accumulated (caller zeroes) or NULL */
  int32_t peep_f16;                             /* 1: peep_i / peep_f / cprev point to fp16 data (ld in elements, 32-byte aligned rows) */
  const float* a_row_sumsq;                     /* optional [M]: the rows of A (BOTH K segments) carry a deferred l2_normalize -- the
                                                   accumulators are scaled by rsqrt(max(a_row_sumsq[m], 1e-12)) before bias / activation */
} cmpc_gemm_args;

int cmpc_gemm_f16(const cmpc_gemm_args* args, void* stream);
/* Tuning knob for measurements: 0 (default) = 2-SM MMA (cta_group::2) for clustered GEMMs, 1 = weight-multicast scheme. */
void cmpc_gemm_set_mode(int mode);

/* Entity perception (CMPC_model.py:295-328): five MUTAN heads in one GEMM + fused epilogue.
 *   out[m, c] = tanh( sum_{k<5} tanh(A[m,:] . Wk[c,:] + bias[k, c]) * lang[b(m), k, c] )     (fp32)
 *   row_sumsq[m] += sum_c out[m, c]^2         (for the l2_normalize over channels that follows, :324)
 * A = [visual | spatial] fp16 [M, k] (k = C + 8).  W is the packed weight produced by
 * pack_mutan_weights() (cmpc_refseg_b200/weights.py): rows ordered (chunk j, head k, cc) with channel c = 48*j + cc. */
typedef struct {
  const void* a;  int64_t lda;  int32_t k;
  const float* a_row_sumsq;                     /* optional [M]: the visual columns of A are NOT yet l2-normalised (:109-113);
                                                   accumulators are scaled by rsqrt(max(a_row_sumsq[m], 1e-12)) and the 8 spatial
                                                   columns of A must hold spatial * sqrt(max(a_row_sumsq[m], 1e-12))
                                                   (cmpc_spatial_fixup_f16) */
  const void* w;  int64_t ldw;                  /* fp16 [21*240, Kpad] */
  int32_t m, c;                                 /* c = channels (<= 1008) */
  int32_t rows_per_sample;
  const float* bias;                            /* [5, ld_bias] */
  int64_t ld_bias;
  const float* lang;                            /* [B][5][ld_lang] tanh(lang_trans) */
  int64_t ld_lang;
  int64_t lang_batch_stride;                    /* elements between samples (0 = 5 * ld_lang) */
  float* out;  int64_t ldo;                     /* fp32 [m, ldo], or fp16 [m, ldo] when out_f16 != 0 (then a void* in disguise) */
  float* row_sumsq;
  int32_t out_f16;                              /* forward only: store the un-normalised map as fp16 (row_sumsq still from the fp32 values) */
} cmpc_mutan_args;

int cmpc_mutan_f16(const cmpc_mutan_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Relation-aware reasoning (CMPC_model.py:376-410, :359-374)
 * ------------------------------------------------------------------------------------------------ */
/* Both softmaxes of the node-word affinity (:388-399).  affi fp32 [B*N, 32] (columns >= T zero; already
 * multiplied by R_t / sqrt(C)), seq_mask fp32 [B, T].  Writes W = softmax over words (masked with
 * tf.float32.min) and V = mask * softmax over nodes as fp16 [B*N, 32]; V is multiplied by v_scale.
 * gw_w / gw_v (fp32 [B*N, T], unscaled) are the reference's public attributes :395,:399 (both or neither). */
size_t cmpc_affinity_workspace_bytes(int32_t batch);
int cmpc_affinity_softmax(const float* affi, const float* seq_mask, int32_t batch, int32_t rows_per_sample, int32_t t,
                          float v_scale, void* w_f16, void* v_f16, float* gw_w, float* gw_v, void* workspace,
                          size_t workspace_bytes, void* stream);
/* Same, for node features whose l2_normalize (:324) is deferred: V additionally carries 1 / |x_j| (x_row_sumsq[j] = sum_c x_raw[j, c]^2),
 * so that the graph kernel may aggregate the un-normalised map.  gw_v (the reference's attribute) stays the plain softmax. */
int cmpc_affinity_softmax_scaled(const float* affi, const float* seq_mask, int32_t batch, int32_t rows_per_sample, int32_t t,
                                 float v_scale, void* w_f16, void* v_f16, float* gw_w, float* gw_v, const float* x_row_sumsq,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* Dense graph aggregation Y = (W V^T) X / v_scale without materialising the N x N adjacency (:400 + :362):
 * flash-style tcgen05/TMEM kernel fed by TMA.  w/v fp16 [B*N, 32], x fp16 [B*N, ldx] (c channels),
 * y fp16 [B*N, ldy]; stats[b] += (sum y, sum y^2) over the sample (fp64, caller zeroes) for the
 * layer norm at :364.  dbg_p (optional, fp32 [B, N, N]) receives the adjacency tiles for tests. */
/* ------------------------------------------------------------------------------------------------
 * Backward building blocks (TF autodiff of the head, CMPC_model.py:461)
 * ------------------------------------------------------------------------------------------------ */
/* Weight gradient of a 1x1 conv in the TF variable layout: out[i, j] += sum_m a[m, i] * b[m, j]  (dW[cin, cout] = X^T dY,
 * :412-417).  a fp16 [m, lda], b fp16 [m, ldb], out fp32 [a_cols, ldo] ACCUMULATED (caller zeroes; shared weights such as
 * the ConvLSTM kernel accumulate over their uses).  tcgen05 with both operands MN-major, split over the m rows
 * (splits <= 0: chosen to fill the GPU).  a_cols, b_cols, lda, ldb multiples of 8. */
int cmpc_gemm_atb_f16(const void* a_f16, int64_t lda, int32_t a_cols, const void* b_f16, int64_t ldb, int32_t b_cols,
                      int32_t m, float* out, int64_t ldo, int32_t splits, void* stream);
/* Same contraction per sample: out[b][i, j] += sum_{n < rows_per_sample} a[b, n, i] * b[b, n, j]; a / b hold batch * rows_per_sample
 * rows, out fp32 [batch][a_cols, ldo] at stride out_bstride. */
int cmpc_gemm_atb_batched_f16(const void* a_f16, int64_t lda, int32_t a_cols, const void* b_f16, int64_t ldb, int32_t b_cols,
                              int32_t rows_per_sample, int32_t batch, float* out, int64_t ldo, int64_t out_bstride, void* stream);

/* Seed of the backward pass (loss :439-445, util/loss.py:6-16; upsample :141 / :129-133; score conv :138):
 * dpred = resize_bilinear^T( scale * (sigmoid(up) - target) ), scale = loss weight / batch; dbias[0] += sum dpred.
 * cmpc_score_bwd_taps lays the nine shifted copies of dpred out as fp16 [rows, ld] (column 3*dy+dx, rest zero): the operand
 * of dF = D9 . w9 (cmpc_gemm_f16) and dw9 = D9^T F (cmpc_gemm_atb_f16). */
int cmpc_score_bwd_dpred(const float* up, const float* target, float scale, int32_t batch, int32_t h, int32_t w, int32_t out_h,
                         int32_t out_w, float* dpred, float* dbias, void* stream);
int cmpc_score_bwd_taps(const float* dpred, int32_t batch, int32_t h, int32_t w, void* d9_f16, int32_t ld, void* stream);
/* Backward of one ConvLSTM step around its five whole-sample layer norms (util/cell.py:46-75), three phases (see
 * csrc/convlstm_bwd.cu): 1 = sums for LN3/LN4, 2 = d o' / d c' + sums for LN0-2 + dW_co, 3 = d j / d i' / d f' (dy16, the
 * operand of the dgrad / wgrad GEMMs of the step's conv), d c_prev, dW_ci, dW_cf.  Phases 1 and 2 also fold their
 * partials: sums[b, 0..9] = sample means needed by the next phase, dgamma / dbeta [5, gw] accumulated.
 * Saved forward tensors: y16 fp16 [rows, 4 gw] (j, i', f', o), opre / cnew / cn / cprev fp32 [rows, gw], mr_g [B, 4, 2] and
 * mr_o [B, 2, 2] (mean, rstd) from cmpc_ln_finalize.  Incoming gradients: dh fp32 (row stride ld_dh), dcn_in (or NULL).
 * cprev == NULL is the first step (no state, no W_ci / W_cf). */
typedef struct cmpc_convlstm_bwd_args {
  const void* y16; const float* opre; const float* cnew; const float* cn; const float* cprev;
  const float* mr_g; const float* mr_o; const float* ln_gamma; const float* ln_beta;
  const float* w_ci; const float* w_cf; const float* w_co;
  const float* dh; int64_t ld_dh; const float* dcn_in;
  float* sums;                 /* [B, 10] */
  float* dcnew;                /* [rows, gw] written by phase 2, read by phase 3 */
  void* dy16;                  /* fp16 [rows, 4 gw] */
  float* dcprev_out;           /* [rows, gw] or NULL */
  float* dw_ci; float* dw_cf; float* dw_co;     /* [rows_per_sample, gw], accumulated */
  float* dgamma; float* dbeta; /* [5, gw], accumulated */
  float* ws_sample; float* ws_chan;             /* cmpc_convlstm_bwd_workspace_floats: [blocks, B, 6] then [N, 6, gw] */
  int32_t gw, m, rows_per_sample;
} cmpc_convlstm_bwd_args;
size_t cmpc_convlstm_bwd_workspace_floats(int32_t batch, int32_t rows_per_sample, int32_t gw);
int cmpc_convlstm_bwd(int32_t phase, const cmpc_convlstm_bwd_args* args, int32_t batch, void* stream);

/* Backward of one gated_exchange_module (:194-259, :272-284); see csrc/exchange_bwd.cu for the math.
 * cmpc_exg_bwd_rows: ds = l2_normalize^T(dout) (fp32 [rows, ld]), dp_i = ds * gate_i * [se_i > 0] (fp16, operands of the
 *   trans_feat dgrad / wgrad GEMMs), colsum[b] (4 rows of ld floats at colsum + b * colsum_bstride, ACCUMULATED): sum_n ds*relu_1,
 *   sum_n ds*relu_2 (-> d gate), sum_n dp_1, sum_n dp_2 (-> d bias).
 * cmpc_gv_gates_bwd: colsum [B, nmod, 4, ld] -> dpre1, dpre2 (d of the two lang_feat convs' outputs), dz (d of the gv_lang conv
 *   output), dpool, all [B, nmod, ld]; weights [nmod, mdim, mdim] with row = input channel.
 * cmpc_pool_bwd_rows: dfeat = ds + dgemm [+ extra] + attention-pooling terms (:226-236), du[b] += sum_n dlogit_n * scale * feat_n.
 * cmpc_small_atb_f32: out[z][i, j] += sum_b a[z][b, i] * c[z][b, j] (parameter gradients of the per-sample linear maps). */
int cmpc_exg_bwd_rows(const float* dout, int64_t ld_dout, const void* out_f16, const float* row_sumsq, const void* se1_f16,
                      const void* se2_f16, const float* gate1, const float* gate2, int64_t gate_bstride, int64_t ld, float* ds,
                      void* dp1_f16, void* dp2_f16, float* colsum, int64_t colsum_bstride, int32_t batch, int32_t rows_per_sample,
                      int32_t width, void* stream);
int cmpc_pool_bwd_rows(const void* feat_f16, int64_t ld, const float* u, int64_t u_bstride, const float* pool, const float* dpool, int64_t vec_bstride,
                       const float* pstats, int64_t pstats_bstride, float scale, const float* ds, const float* dgemm,
                       int64_t ld_dgemm, const float* extra, int64_t ld_extra, float* dfeat, float* du, int64_t du_bstride,
                       int32_t batch, int32_t rows_per_sample, int32_t width, void* stream);
int cmpc_gv_gates_bwd(const float* colsum, const float* gate1, const float* gate2, const float* gv, const float* pool,
                      const float* gvl, int64_t gvl_bstride, int64_t gvl_mstride, const float* wg, const float* wf1,
                      const float* wf2, int64_t w_mstride, int32_t batch, int32_t nmod, int32_t mdim, int64_t ld, float* dpre1,
                      float* dpre2, float* dz, float* dpool, void* stream);
/* Backward of the batch-coupled variant (forward: cmpc_gv_gates_batch): phase 1 emits dpre1 / dpre2, parks d gv in dz and ADDS
 * sum_b gv_b . dgv_b into batch_dot[nmod] (zeroed by the caller); phase 2 finishes dz / dpool with |z|^2 = batch_ss[mod] kept from the
 * forward.  A data-parallel caller all-reduces batch_dot between the phases only if it also all-reduced batch_ss in the forward. */
int cmpc_gv_gates_bwd_batch(const float* colsum, const float* gate1, const float* gate2, const float* gv, const float* pool,
                            const float* gvl, int64_t gvl_bstride, int64_t gvl_mstride, const float* wg, const float* wf1,
                            const float* wf2, int64_t w_mstride, int32_t batch, int32_t nmod, int32_t mdim, int64_t ld,
                            float* dpre1, float* dpre2, float* dz, float* dpool, int32_t phase, const float* batch_ss,
                            float* batch_dot, void* stream);
int cmpc_small_atb_f32(const float* a, int64_t lda, int64_t a_zstride, const float* c, int64_t ldc, int64_t c_zstride, float* out,
                       int64_t ldo, int64_t o_zstride, int32_t nz, int32_t nb, int32_t ni, int32_t nj, void* stream);

/* Row kernels of the per-level backward (fusion conv :338-344, graph_conv :359-374, affinity softmaxes :388-399); see
 * csrc/level_bwd.cu.  All buffers [batch * rows_per_sample, ld]; colsum / dgamma / dbeta / sums / drgate are ACCUMULATED. */
int cmpc_relu_mask_f16(const float* dout, int64_t ld_d, const void* act_f16, int64_t ld, void* dpre_f16, float* colsum /*[B, ld]*/,
                       int32_t batch, int32_t rows_per_sample, int32_t width, void* stream);
int cmpc_ln_bwd_sums(const float* dout, int64_t ld_d, const void* act_f16, const float* row_sumsq /* NULL: no l2_normalize in front */,
                     const void* pre_f16, int64_t ld, const float* mean_rstd, const float* gamma, float* dln, int64_t ld_ln,
                     double* sums /*[B, 2]*/, float* dgamma, float* dbeta, int32_t batch, int32_t rows_per_sample, int32_t width,
                     void* stream);
int cmpc_ln_bwd_apply(const float* dln, int64_t ld_ln, const void* pre_f16, int64_t ld, const float* mean_rstd, const float* gamma,
                      const double* sums, void* out_f16, float* colsum /*[ld] or NULL*/, int32_t batch, int32_t rows_per_sample,
                      int32_t width, void* stream);
int cmpc_affinity_bwd(const void* w_f16, const void* v_f16, const float* dw, const float* dv, const float* affi, const float* rgate,
                      float v_scale, int32_t batch, int32_t rows_per_sample, float* colsum_ws /*[B, 32]*/, void* draw_f16,
                      float* drgate /*[B, 32]*/, void* stream);
int cmpc_transpose_gt_f16(const void* gt_f16, int64_t ld, int32_t batch, int32_t t, int32_t rows_out, void* gtT_f16, void* stream);

/* Backward of MUTAN (:295-328) and of the lateral l2_normalize folded into it (:109-113):
 * cmpc_mutan_out_bwd: ds = d loss / d (sum of the five gated heads) from the (up to four) fp32 pieces of d loss / d vis_la_sp;
 * cmpc_mutan_bwd_f16: the MUTAN GEMM again with a backward epilogue: out = fp16 d(pre-activation) * row scale [m, chunks * 240]
 *   (packed column order of the MUTAN weight), dlang[b][5][ld_lang] += sum_n ds * tanh(pre), dbias[5][ld_bias] += sum_m d pre;
 * cmpc_lateral_bwd: d(lateral conv output) = G - xlat (xlat . G) / |xlat|^2 as fp16 (+ its column sums = bias gradient). */
int cmpc_mutan_out_bwd(const float* p0, const float* p1, const float* p2, const float* p3, int64_t ldp, const void* x_f16, int64_t ld,
                       const float* row_sumsq, float* ds, int64_t ld_ds, int64_t rows, int32_t width, void* stream);
int cmpc_mutan_bwd_f16(const cmpc_mutan_args* args, const float* ds, int64_t ld_ds, float* dlang, int64_t dlang_batch_stride,
                       float* dbias, void* stream);
/* out = dy * (1 - y^2) (kind 2, tanh) or dy * y (1 - y) (kind 3, sigmoid) on small fp32 vectors of the language side. */
int cmpc_act_bwd_f32(const float* dy, const float* y, float* out, int64_t n, int32_t kind, void* stream);
int cmpc_lateral_bwd(const float* g, int64_t ldg, const void* xlat_f16, int64_t ld, const float* row_sumsq, void* out_f16, float* colsum,
                     int32_t batch, int32_t rows_per_sample, int32_t width, void* stream);

/* Backward of the language side (:159-192, :347-357).  cmpc_lang_bwd (block per sentence): from d valid_lang, d nec_lang [B, r] and
 * the relation-gate gradient drgate [B, 32] (w.r.t. parse[..., 2] / sqrt(c)) -> dwords [B*T, r] (WRITTEN: the two weighted-sum
 * paths) and dlogit [B, T, 4] (gradient of the word-type logits before the masked softmax).  cmpc_l2norm_bwd_f32: the
 * l2_normalize of the LSTM outputs (:159); cmpc_relu_bwd_f32: out = dy * [y > 0]. */
int cmpc_lang_bwd(const float* words_f32, const float* parse, const float* seq_mask, const float* valid_f32, const float* nec_f32,
                  const float* d_valid, const float* d_nec, const float* drgate, int32_t batch, int32_t t, int32_t r, int32_t c,
                  float* dwords, float* dlogit, void* stream);
int cmpc_l2norm_bwd_f32(const float* dy, const float* y, const float* x, int32_t rows, int32_t r, float* dx, void* stream);
int cmpc_relu_bwd_f32(const float* dy, const float* y, float* out, int32_t rows, int32_t cols, int64_t ld, void* stream);

/* Optimizer step of train_op (:446-478): Adam on g = grad * grad_scale + weight_decay * w (grad_scale = 2 for `biases`, :464-475;
 * weight_decay only for `DW` variables, util/loss.py:28-32), in place on a flat fp32 group; lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t).
 * lr_t_dev (optional, device scalar) overrides lr_t, so that a captured CUDA graph of the step can follow the decaying rate. */
int cmpc_adam_f32(float* w, const float* grad, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2, float eps,
                  float grad_scale, float weight_decay, const float* lr_t_dev, void* stream);

/* Word encoder in front of the head (:144-157): embedding lookup as the fp16 A operand of the input-half GEMM,
 * and one step of tf LSTMCell (forget_bias 1, no peepholes) under dynamic_rnn(sequence_length): xg fp32 [B*T, 4r] = x K_x + b for all
 * steps (cmpc_gemm_f16), hg fp32 [B, 4r] = h_{t-1} K_h (cmpc_gemm_f16 per step; NULL at t = 0); c_state fp32 [B, r] and h fp16 [B, ldh]
 * are updated in place, out fp32 [B, T, r] receives h_t (zero past seq_len). */
int cmpc_embed_gather_f16(const int32_t* ids, const float* emb, int32_t vocab, int32_t e, int32_t rows, void* out_f16, int64_t ld, void* stream);
int cmpc_lstm_step(const float* xg, const float* hg, const int32_t* seq_len, int32_t t, int32_t steps, int32_t r, int32_t batch,
                   float* c_state, void* h_f16, int64_t ldh, float* out, void* stream);
/* Training form of the same step and its backward (the reference trains `Variable`, `rnn/lstm_cell/kernel`, `rnn/lstm_cell/bias`
 * with everything else under text_objseg/, :426-431).  Rows of xg / dz are time-major (t * B + b).  State slots: c fp32 [B, r] and
 * h fp16 [B, ldh] per step, the caller passes slot t as *_prev and slot t + 1 as *_new (slot 0 zeros); gates fp32 [B, 4r] of step t
 * receives (sigmoid i, tanh j, sigmoid(f + 1), sigmoid o).  cmpc_lstm_step_bwd(t): d_out fp32 [B, T, r] = d loss / d outputs,
 * g_rec fp32 [B, ldg] = scale * dz_{t+1} K_h^T (cmpc_gemm_f16; NULL at the last step), dc_state fp32 [B, r] carried between steps,
 * dz fp16 [T * B, ldz] row t * B + b receives scale * (dz_i, dz_j, dz_f, dz_o) (operand of d kernel = [X, H_prev]^T dz, of
 * d x = dz K_x^T and of the next g_rec), dbias fp32 [4r] += sum_b dz (unscaled).
 * cmpc_embed_scatter_add: demb[ids[row], :e] += scale * dx[row, :e] (the IndexedSlices gradient of embedding_lookup, summed densely). */
int cmpc_lstm_step_train(const float* xg, const float* hg, const int32_t* seq_len, int32_t t, int32_t steps, int32_t r, int32_t batch,
                         const float* c_prev, float* c_new, const void* h_prev_f16, void* h_new_f16, int64_t ldh, float* gates,
                         float* out, void* stream);
int cmpc_lstm_step_bwd(const float* d_out, const float* g_rec, int64_t ldg, const int32_t* seq_len, int32_t t, int32_t steps, int32_t r,
                       int32_t batch, const float* gates, const float* c_prev, const float* c_cur, float* dc_state, float scale,
                       void* dz_f16, int64_t ldz, float* dbias, void* stream);
int cmpc_embed_scatter_add(const int32_t* ids, const float* dx, int64_t ld, float scale, int32_t vocab, int32_t e, int32_t rows, float* demb,
                           void* stream);

/* Measurement knob: 0 (default) = persistent cta_group::1 kernel with TMA multicast, 2 = 2-SM MMA (tcgen05 cta_group::2)
 * variant (measured slower, kept for A/B runs; see graph_tc.cu). */
void cmpc_graph_set_mode(int mode);
int cmpc_graph_reason_f16(const void* w_f16, const void* v_f16, const void* x_f16, int64_t ldx, int32_t batch,
                          int32_t n_nodes, int32_t c, float v_scale, void* y_f16, int64_t ldy, double* stats,
                          float* dbg_p, void* stream);

/* Whole-sample layer-norm statistics (tf.contrib.layers.layer_norm, begin_norm_axis=1): turns n fp64 (sum, sumsq)
 * pairs accumulated by a producer epilogue into fp32 (mean, rsqrt(biased var + 1e-12)) pairs. */
int cmpc_ln_finalize(const double* stats, int32_t n, double count, float* mean_rstd, void* stream);
/* relu(X + LN(Y))  (:364-367).  mean_rstd = [B, 2] from cmpc_ln_finalize. */
int cmpc_ln_residual_relu_f16(const void* y, int64_t ldy, const void* x, int64_t ldx, const float* mean_rstd,
                              const float* gamma, const float* beta, void* out, int64_t ldo, int64_t rows, int32_t c,
                              int32_t rows_per_sample, void* stream);
/* l2_normalize_C(relu(LN(U)))  (:370-372, :408); optionally appends the 8 spatial channels at [c, c+8).
 * normalize == 0 stops at relu(LN(U)), the value graph_conv itself returns (:372). */
int cmpc_ln_relu_l2norm_f16(const void* u, int64_t ldu, const float* mean_rstd, const float* gamma, const float* beta,
                            void* out, int64_t ldo, int64_t rows, int32_t c, int32_t spatial_h, int32_t spatial_w,
                            int32_t rows_per_sample, int32_t normalize, float* row_sumsq /* optional [rows], for the backward */,
                            void* stream);
/* Deferred l2_normalize of the MUTAN map (CMPC_model.py:324): the map X stays un-normalised in memory next to its row sums of squares
 * ss, and its consumers apply 1 / |x| themselves -- cmpc_affinity_softmax_scaled (V), cmpc_gemm_args.a_row_sumsq (affinity and fusion
 * GEMMs), cmpc_ln_residual_relu_scaled_f16 (the residual X + LN(Y), :366) -- while cmpc_ln_relu_l2norm_scaled_f16 multiplies ITS output
 * rows by |x| = sqrt(max(out_row_sumsq, 1e-12)) because the fusion GEMM that reads them scales its whole accumulator by 1 / |x|. */
int cmpc_ln_residual_relu_scaled_f16(const void* y, int64_t ldy, const void* x, int64_t ldx, const float* x_row_sumsq,
                                     const float* mean_rstd, const float* gamma, const float* beta, void* out, int64_t ldo,
                                     int64_t rows, int32_t c, int32_t rows_per_sample, void* stream);
int cmpc_ln_relu_l2norm_scaled_f16(const void* u, int64_t ldu, const float* mean_rstd, const float* gamma, const float* beta,
                                   void* out, int64_t ldo, int64_t rows, int32_t c, int32_t spatial_h, int32_t spatial_w,
                                   int32_t rows_per_sample, int32_t normalize, float* row_sumsq, const float* out_row_sumsq,
                                   void* stream);
/* A/B knob: 0 (default) = wide rows of large maps go through the bulk-copy staged kernel, 1 = register kernels only. */
void cmpc_ln_relu_l2norm_set_mode(int32_t mode);

/* ------------------------------------------------------------------------------------------------
 * Element-wise / row-wise helpers
 * ------------------------------------------------------------------------------------------------ */
/* fp32 -> fp16 (backbone taps c3/c4/c5, CMPC_model.py:74-76, become GEMM operands). */
int cmpc_cast_f32_f16(const float* in, int64_t ldi, void* out, int64_t ldo, int64_t rows, int32_t cols, void* stream);
/* Same with a scale: fp16(in * scale) -- loads a caller-supplied gw_v (:391) as the graph kernel's scaled V operand. */
/* dst[c, r] = fp16(src[r, c]): a TF kernel [Cin, Cout] (fp32) -> the K-major fp16 GEMM operand [Cout, Cin] (operand refresh after an
 * optimizer step).  group > 0: destination row of source column c = (c / group) * group_stride + (c % group) * ld_dst elements from
 * dst (the interleaved (chunk, head, channel) rows of the MUTAN weight). */
int cmpc_transpose_cast_f32_f16(const float* src, int64_t ld_src, int32_t rows, int32_t cols, void* dst_f16, int64_t ld_dst, int32_t group,
                                int64_t group_stride, void* stream);
int cmpc_scale_cast_f32_f16(const float* in, int64_t ldi, float scale, void* out, int64_t ldo, int64_t rows, int32_t cols,
                            void* stream);
/* tf.nn.l2_normalize(x, 3) given per-row sum of squares (:109-113, :324): out = in * rsqrt(max(ss, 1e-12)) as fp16;
 * spatial_h > 0 appends generate_spatial_batch's 8 channels (util/processing_tools.py:5-17) at [c, c+8);
 * spatial_h == -1 appends a single 1.0 at column c (homogeneous coordinate used by the affinity GEMM). */
int cmpc_rownorm_f16(const float* in, int64_t ldi, const float* row_sumsq, void* out, int64_t ldo, int64_t rows,
                     int32_t c, int32_t spatial_h, int32_t spatial_w, int32_t rows_per_sample, void* stream);
/* Same with an fp16 input map (what cmpc_mutan_f16 writes with out_f16); in == out normalises in place. */
int cmpc_rownorm_h16(const void* in_f16, int64_t ldi, const float* row_sumsq, void* out, int64_t ldo, int64_t rows,
                     int32_t c, int32_t spatial_h, int32_t spatial_w, int32_t rows_per_sample, void* stream);
/* Companion of cmpc_mutan_f16(a_row_sumsq): writes columns [c, c+8) of the fp16 map x as
 * generate_spatial_batch(pixel) * sqrt(max(row_sumsq[m], 1e-12)) and zeroes [c+8, ldx). */
int cmpc_spatial_fixup_f16(void* x, int64_t ldx, const float* row_sumsq, int64_t rows, int32_t c, int32_t spatial_h,
                           int32_t spatial_w, void* stream);
/* l2_normalize_C(a + b + c)  (gated_exchange_module :258 + :272-284); pads are zero in all inputs.
 * normalize == 0 returns the plain sum, the value gated_exchange_module itself returns (:258). */
int cmpc_add3_l2norm_f16(const void* a, const void* b, const void* c, int64_t ld, void* out, int64_t ldo,
                         int64_t rows, int32_t width, int32_t normalize, float* row_sumsq /* optional [rows], for the backward */,
                         void* stream);
/* same with one leading dimension per input (b / c may be column slices of wider maps, e.g. the merged lang_se outputs) */
int cmpc_add3_l2norm_ld_f16(const void* a, int64_t lda, const void* b, int64_t ldb, const void* c, int64_t ldc, void* out, int64_t ldo,
                            int64_t rows, int32_t width, int32_t normalize, float* row_sumsq, void* stream);
/* global_vec attention pooling (:226-236) with the key conv folded into u = W_key q (softmax is shift
 * invariant): out[b, mod, :] = softmax_n(feat_mod[b, n, :] . u[b, mod, :] * scale)^T feat_mod[b].   Up to 3
 * modules per launch (feat0..2 fp16 [B*N, ld]); u fp32, sample b module m at u + b*u_bstride + m*ldu;
 * out fp32 [B, nmod, ldo]. */
size_t cmpc_global_pool_workspace_bytes(int32_t batch, int32_t nmod, int32_t width);
int cmpc_global_pool_f16(const void* feat0, const void* feat1, const void* feat2, int64_t ld, const float* u,
                         int64_t ldu, int64_t u_bstride, int32_t nmod, int32_t batch, int32_t rows_per_sample, int32_t width, float scale,
                         float* out, int64_t ldo, float* stats_out /* optional [B, nmod, 2] (max logit, sum exp), for the backward */,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Language side (CMPC_model.py:159-192, :347-357, :202-204, :223, :238-241)
 * ------------------------------------------------------------------------------------------------ */
/* words_feat = l2_normalize(lstm_outputs, -1) as fp32 [rows, r] and fp16 [rows, ld16]; seq_mask[rows]. */
int cmpc_words_prepare(const float* lstm_outputs, int32_t rows, int32_t r, float* words_f32, void* words_f16,
                       int64_t ld16, float* seq_mask, void* stream);
/* word-type attention: parse = softmax4(hidden W2 + b2) * mask (hidden = relu(words_parse_1), from the GEMM);
 * rgate[b, 32] = parse[..., 2] / sqrt(c) (relation weight, zero padded);  valid = l2norm(sum_t (E+A) w_t),
 * nec = l2norm(sum_t (E+A+R) w_t) as fp32 [B, r] and fp16 [B, ld16].
 * hidden == NULL: `parse` is an INPUT (caller-supplied word weights, valid_lang / nec_lang :166-192 on their own);
 * w2, b2 and seq_mask are then ignored. */
int cmpc_lang_parse(const float* hidden, int64_t ldh, int32_t hid, const float* w2, const float* b2,
                    const float* words_f32, const float* seq_mask, int32_t batch, int32_t t, int32_t r, int32_t c,
                    float* parse, float* rgate, float* valid_f32, float* nec_f32, void* valid_f16, void* nec_f16,
                    int64_t ld16, void* stream);
/* fp32 batched skinny matmul out[z] = act(x[z] W[z] + bias[z]), W row-major [k, n]; rows <= 64.
 * act: 0 none, 1 relu, 2 tanh, 3 sigmoid, 4 = accumulate into out (out += x W + bias; distinct out per z). */
int cmpc_small_linear_f32(const float* x, int64_t ldx, int64_t x_zstride, const float* w, int64_t ldw,
                          int64_t w_zstride, const float* bias, int64_t b_zstride, float* out, int64_t ldo,
                          int64_t o_zstride, int32_t nbatch, int32_t rows, int32_t k, int32_t n, int32_t act,
                          void* stream);
/* global_vec tail + lang_se gates: gv = l2norm(g Wg + gvl) per sample (:238-241), gate_f = sigmoid(gv Wf + bf) (:202-204). */
int cmpc_gv_gates(const float* g, int64_t ldg, const float* gvl, int64_t ldgvl, int64_t gvl_bstride, const float* wg, const float* wf1,
                  const float* bf1, const float* wf2, const float* bf2, int64_t w_mstride, int64_t b_mstride,
                  int32_t batch, int32_t nmod, int32_t mdim, float* gv, float* gate1, float* gate2, int64_t ldgate,
                  void* stream);
/* The literal graph at batch > 1: tf.nn.l2_normalize(gv_lang) has NO axis (CMPC_model.py:241), i.e. the sum of squares runs over
 * the batch as well.  Two launches around one [nmod] float buffer: phase 1 writes the un-normalised z = g Wg + gvl to gv and ADDS
 * sum |z|^2 into batch_ss[mod] (zeroed by the caller); phase 2 normalises by batch_ss[mod] and emits the gates.  A caller that shards
 * the batch over ranks all-reduces batch_ss between the phases to reproduce the reference at the global batch. */
/* Most general form: the two gates are written at gate_f + b * gf_bstride + mod * gf_mstride (strides in floats, may be negative;
 * both 0 = the default [B, nmod, ldgate] layout), phase 0 = per-sample norm in one launch, 1 / 2 = the batch-coupled pair above. */
int cmpc_gv_gates_ex(const float* g, int64_t ldg, const float* gvl, int64_t ldgvl, int64_t gvl_bstride, const float* wg,
                     const float* wf1, const float* bf1, const float* wf2, const float* bf2, int64_t w_mstride, int64_t b_mstride,
                     int32_t batch, int32_t nmod, int32_t mdim, float* gv, float* gate1, int64_t g1_bstride, int64_t g1_mstride,
                     float* gate2, int64_t g2_bstride, int64_t g2_mstride, int64_t ldgate, int32_t phase, float* batch_ss, void* stream);
int cmpc_gv_gates_batch(const float* g, int64_t ldg, const float* gvl, int64_t ldgvl, int64_t gvl_bstride, const float* wg,
                        const float* wf1, const float* bf1, const float* wf2, const float* bf2, int64_t w_mstride, int64_t b_mstride,
                        int32_t batch, int32_t nmod, int32_t mdim, float* gv, float* gate1, float* gate2, int64_t ldgate,
                        int32_t phase, float* batch_ss, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ConvLSTM fusion gates (util/cell.py:46-75) -- the 1x1 conv itself is cmpc_gemm_f16 with peepholes/stats.
 * y fp32 or fp16 [rows, 4*gw] (j,i,f,o), state fp32 [rows, gw], ln_gamma/beta fp32 [5, gw] (j,i,f,o,c).
 * ------------------------------------------------------------------------------------------------ */
int cmpc_convlstm_gates1(const void* y, int32_t y_fp16, int64_t ldy, int32_t gw, int32_t m, const float* mean_rstd_in /* [B,4,2] */,
                         const float* ln_gamma, const float* ln_beta, const float* cprev, const float* w_co,
                         float* cnew, float* opre, double* stats_out, int64_t rows, int32_t rows_per_sample,
                         void* stream);
int cmpc_convlstm_gates2(const float* opre, const float* cnew, int32_t gw, int32_t m, const float* mean_rstd /* [B,2,2] */,
                         const float* ln_gamma, const float* ln_beta, float* c_out, void* h_f16, float* h_f32,
                         int64_t rows, int32_t rows_per_sample, void* stream);
/* Same second half of the cell (util/cell.py:66-75) without the fp32 o' map: o' = o + W_co * c' (:66-67) is recomputed from the
 * GEMM's fp16 gate map (y_o_f16 = column block 3 of y, leading dimension ldy) -- pass opre = NULL to cmpc_convlstm_gates1 then.
 * c_out may be NULL here and in cmpc_convlstm_gates2 (last step of the sequence: only h is consumed, CMPC_model.py:290). */
int cmpc_convlstm_gates2_y16(const void* y_o_f16, int64_t ldy, const float* w_co, const void* cnew, int32_t gw, int32_t m,
                             const float* mean_rstd /* [B,2,2] */, const float* ln_gamma, const float* ln_beta, void* c_out,
                             void* h_f16, int32_t state_f16 /* cnew / c_out are fp16 */, int64_t rows, int32_t rows_per_sample,
                             void* stream);
/* First half of the cell for the inference path with the cell state kept in fp16 between the passes (cprev, cnew: fp16 [rows, gw];
 * the statistics of o' and c' are still taken from the fp32 values); no o' map (see cmpc_convlstm_gates2_y16). */
int cmpc_convlstm_gates1_h16(const void* y_f16, int64_t ldy, int32_t gw, int32_t m, const float* mean_rstd_in /* [B,4,2] */,
                             const float* ln_gamma, const float* ln_beta, const void* cprev_f16, const float* w_co,
                             void* cnew_f16, double* stats_out, int64_t rows, int32_t rows_per_sample, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Score head (CMPC_model.py:128-133, :138-142) and evaluation counts (:486-489, util/eval_tools.py:31-35)
 * ------------------------------------------------------------------------------------------------ */
size_t cmpc_score_workspace_bytes(int64_t rows);
/* pred = conv3x3(feat; M->1) + bias (SAME);  up = legacy resize_bilinear(pred, [out_h, out_w]);  sigm = sigmoid(up).
 * w9 fp32 [9, ld] (tap-major), up/sigm may be NULL. */
int cmpc_score_upsample(const void* feat_f16, int64_t ld, const float* w9, float bias, int32_t batch, int32_t h,
                        int32_t w, int32_t width, int32_t out_h, int32_t out_w, float* pred, float* up, float* sigm,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Same, with the nine tap dot products taps[pixel, 3*dy+dx] = feat[pixel, :] . w[dy, dx, :] already computed by
 * cmpc_gemm_f16 (N = 9 padded to 32): only the 3x3 gather + bias, the upsampling and the sigmoid run here. */
int cmpc_score_from_taps(const float* taps, int64_t ld_taps, float bias, const float* bias_dev /* optional: device scalar, overrides bias */,
                         int32_t batch, int32_t h, int32_t w, int32_t out_h, int32_t out_w, float* pred, float* up, float* sigm, void* stream);
/* sums[b] += sum over the sample's pixels of tf.nn.sigmoid_cross_entropy_with_logits(logits, target)
 * (util/loss.py:6-16 with pos/neg multipliers 1; CMPC_model.py:440-443 takes the mean of these over the batch). fp64, caller zeroes. */
int cmpc_sigmoid_ce_sums(const float* logits, const float* target, int32_t batch, int64_t per_sample, double* sums,
                         void* stream);
/* iu[b] += (|pred & gt|, |pred | gt|), pred = up > thresh (inclusive: >=), gt = target != 0.  uint64 [B, 2]. */
int cmpc_iou_counts(const float* up, const float* target, int32_t batch, int64_t per_sample, float thresh,
                    int32_t inclusive, uint64_t* iu, void* stream);
/* Post-processing of the test loop (trainval_model.py:243-245, 266; util/im_processing.py:25-41; util/eval_tools.py:31-35):
 * pred_raw = up >= score_thresh [batch, h, w]; per sample b the mask is resized (skimage.transform.resize semantics: order 1,
 * half-pixel centres, no anti-aliasing; reflect = 0: mode 'constant' cval 0 (skimage <= 0.14), 1: 'reflect') to res_h x res_w,
 * centre-cropped to gh x gw and compared with the ground truth; any non-zero resized value is foreground, as compute_mask_IU's
 * logical_and / logical_or treat it.  Ragged layout: sample b's gh * gw bytes of gt (and of pred_out, optional) start at
 * gt_offset[b]; meta[b] = {gh, gw, res_h, res_w, crop_h, crop_w} with the host-side integer arithmetic of resize_and_crop.
 * iu[b] += (|pred & gt|, |pred | gt|), uint64 [batch, 2], caller zeroes. */
int cmpc_postprocess_iou(const float* up, int32_t batch, int32_t h, int32_t w, float score_thresh, const uint8_t* gt,
                         const int64_t* gt_offset, const int32_t* meta, int32_t reflect, uint8_t* pred_out, uint64_t* iu,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMPC_B200_H_ */

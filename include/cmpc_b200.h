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

/* ------------------------------------------------------------------------------------------------
 * Dense 1x1-convolution GEMM with fused epilogue (tcgen05 + TMEM + TMA, persistent, warp-specialised).
 * Replaces LSTM_model._conv for filter_size 1 (CMPC_model.py:412-417) and the elementwise TF nodes that
 * follow each call site (bias add, relu, gates, norm statistics).
 *
 *   acc[m, n] = sum_k A1[m, k] * W[n, k]  (+ sum_k A2[m, k] * W[n, K1pad + k])      fp16 x fp16 -> fp32
 *   v = acc * row_scale[m] + bias[n] + sbias[b(m), n] + peephole          b(m) = m / rows_per_sample
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
  int32_t m, n;
  int32_t rows_per_sample;                      /* b(m) = m / rows_per_sample (>= 1) */
  /* epilogue */
  const float* row_scale;                       /* [M] or NULL */
  const float* bias;                            /* [N] or NULL */
  const float* sbias;   int64_t ld_sbias;       /* [B, ld] per-sample bias or NULL */
  const float* gate;    int64_t ld_gate;        /* [B, ld] per-sample multiplicative gate or NULL */
  int32_t act;                                  /* 0 none, 1 relu */
  int32_t group_width, group_valid;             /* column validity / statistics groups (0 = one group of N) */
  /* ConvLSTM peepholes (util/cell.py:48-50): v += peep_g[pixel(m), c] * cprev[m, c] for group 1 and 2 */
  const float* peep_i;  const float* peep_f;  int64_t ld_peep;
  const float* cprev;   int64_t ld_cprev;
  /* outputs */
  void* out;  int64_t ldo;  int32_t out_fp32;   /* fp16 (0) or fp32 (1) */
  float* row_sumsq;                             /* [M] accumulated (caller zeroes) or NULL */
  double* stats;                                /* [B, n_groups, 2] accumulated (caller zeroes) or NULL */
} cmpc_gemm_args;

int cmpc_gemm_f16(const cmpc_gemm_args* args, void* stream);

/* Entity perception (CMPC_model.py:295-328): five MUTAN heads in one GEMM + fused epilogue.
 *   out[m, c] = tanh( sum_{k<5} tanh(A[m,:] . Wk[c,:] + bias[k, c]) * lang[b(m), k, c] )     (fp32)
 *   row_sumsq[m] += sum_c out[m, c]^2         (for the l2_normalize over channels that follows, :324)
 * A = [visual | spatial] fp16 [M, k] (k = C + 8).  W is the packed weight produced by
 * cmpc_mutan_pack_layout(): rows ordered (chunk j, head k, cc) with channel c = 48*j + cc. */
typedef struct {
  const void* a;  int64_t lda;  int32_t k;
  const void* w;  int64_t ldw;                  /* fp16 [21*240, Kpad] */
  int32_t m, c;                                 /* c = channels (<= 1008) */
  int32_t rows_per_sample;
  const float* bias;                            /* [5, ld_bias] */
  int64_t ld_bias;
  const float* lang;                            /* [B, 5, ld_lang] tanh(lang_trans) */
  int64_t ld_lang;
  float* out;  int64_t ldo;
  float* row_sumsq;
} cmpc_mutan_args;

int cmpc_mutan_f16(const cmpc_mutan_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMPC_B200_H_ */

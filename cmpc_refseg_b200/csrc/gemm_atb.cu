// Weight-gradient GEMM  G[i, j] += sum_m A[m, i] * B[m, j]   ("A^T B": the contraction runs over the ROWS of two row-major
// fp16 matrices, e.g. dW[cin, cout] = X^T dY of a 1x1 conv -- the backward of CMPC_model.py:412-417 w.r.t. DW, in the TF
// variable layout [Cin, Cout], TF autodiff at :461).
//
// tcgen05 with BOTH operands MN-major: a stage holds 64 rows (the K dimension) of a 128-column slice of A and of a
// 256-column slice of B, as [64 rows x 64 cols] TMA boxes with the 128-byte swizzle (the same layout the graph kernel uses
// for its X tiles).  One CTA = one 128 x 256 output tile x one split of the rows; the fp32 accumulator (256 TMEM columns)
// is added to G with atomics (split-K over tens of thousands of rows is what fills 148 SMs: a 1000 x 1000 gradient has only
// 32 tiles).  Warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue (TMEM lane quadrant = warp % 4).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

constexpr int W_BI = 128;            // output rows per tile (columns of A)
constexpr int W_BJ = 256;            // output columns per tile (columns of B)
constexpr int W_BK = 64;             // contraction rows per stage
constexpr int W_STAGES = 4;
constexpr int W_THREADS = 192;
constexpr int W_BOX_BYTES = W_BK * 128;                   // one [64 rows x 64 cols] box, 8 KB
constexpr int W_A_BYTES = (W_BI / 64) * W_BOX_BYTES;      // 16 KB
constexpr int W_B_BYTES = (W_BJ / 64) * W_BOX_BYTES;      // 32 KB
constexpr int W_STAGE_BYTES = W_A_BYTES + W_B_BYTES;
constexpr int W_BAR_OFF = W_STAGES * W_STAGE_BYTES;
constexpr int W_SMEM = W_BAR_OFF + 128 + 1024;

struct AtbParams {
  int m, a_cols, b_cols, kblocks_per_split, splits, vec4;      // m = rows per batch entry
  float* out;
  long long ldo, out_bstride;
};

__global__ void __launch_bounds__(W_THREADS, 1)
gemm_atb_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const AtbParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + W_BAR_OFF);
  uint64_t* empty = full + W_STAGES;
  uint64_t* acc_full = empty + W_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * W_BI, j0 = blockIdx.y * W_BJ;
  const int kb_total = (p.m + W_BK - 1) / W_BK;
  const int bz = blockIdx.z / p.splits;                     // batch entry (per-sample contraction), 0 when not batched
  const int kb0 = (blockIdx.z - bz * p.splits) * p.kblocks_per_split;
  const int kb1 = min(kb_total, kb0 + p.kblocks_per_split);
  const int nkb = kb1 - kb0;                                // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int k = 0; k < nkb; ++k) {
        const int s = k % W_STAGES;
        mbar_wait(&empty[s], ((k / W_STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * W_STAGE_BYTES;
        uint8_t* sb = sa + W_A_BYTES;
        mbar_expect_tx(&full[s], W_STAGE_BYTES);
        const int row = (kb0 + k) * W_BK;                  // rows >= m are zero-filled by the tensor maps
#pragma unroll
        for (int b = 0; b < W_BI / 64; ++b) tma_load_3d(sa + b * W_BOX_BYTES, &tmA, &full[s], i0 + b * 64, row, bz);
#pragma unroll
        for (int b = 0; b < W_BJ / 64; ++b) tma_load_3d(sb + b * W_BOX_BYTES, &tmB, &full[s], j0 + b * 64, row, bz);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(W_BI, W_BJ, 0, /*A MN-major*/ 1, /*B MN-major*/ 1);
      for (int k = 0; k < nkb; ++k) {
        const int s = k % W_STAGES;
        mbar_wait(&full[s], (k / W_STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * W_STAGE_BYTES), sb = sa + W_A_BYTES;
        // MN-major, 128B swizzle: LBO = distance between 64-column boxes, SBO = 8 contraction rows (8 x 128 B)
        const uint64_t da = make_smem_desc(sa, W_BOX_BYTES, 1024, 2);
        const uint64_t db = make_smem_desc(sb, W_BOX_BYTES, 1024, 2);
#pragma unroll
        for (int kk = 0; kk < W_BK / 16; ++kk)
          umma_f16_ss(tmem_base, da + uint64_t((kk * 16 * 128) >> 4), db + uint64_t((kk * 16 * 128) >> 4), idesc, (k | kk) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // epilogue: thread = output row i0 + 32 * (warp % 4) + lane; 256 columns in chunks of 32
    const int q = warp & 3;
    const int i = i0 + q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* orow = p.out + (long long)bz * p.out_bstride + (long long)i * p.ldo + j0;
#pragma unroll 1
    for (int c = 0; c < W_BJ / 32; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem_base + (uint32_t(q * 32) << 16) + c * 32, r);
      tmem_wait_ld();
      if (i < p.a_cols) {
        if (p.vec4) {                        // 16-byte aligned rows: four columns per atomic (red.global.add.v4.f32)
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            if (j0 + c * 32 + e < p.b_cols)  // b_cols % 8 == 0: a group of four is entirely inside or outside
              atomicAdd(reinterpret_cast<float4*>(orow + c * 32 + e),
                        make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (j0 + c * 32 + e < p.b_cols) atomicAdd(orow + c * 32 + e, __uint_as_float(r[e]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace cmpc

using namespace cmpc;

static int launch_atb(const void* a_f16, int64_t lda, int32_t a_cols, const void* b_f16, int64_t ldb, int32_t b_cols, int32_t m, int32_t batch,
                      float* out, int64_t ldo, int64_t out_bstride, int32_t splits, cudaStream_t stream, const char* who) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(a_f16 && b_f16 && out && a_cols > 0 && b_cols > 0 && m > 0 && batch > 0, CMPC_ERR_ARG, "%s: bad args", who);
  CMPC_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && a_cols % 8 == 0 && b_cols % 8 == 0 && lda >= a_cols && ldb >= b_cols, CMPC_ERR_ALIGN,
               "%s: lda, ldb, a_cols, b_cols must be multiples of 8", who);
  CMPC_REQUIRE(ldo >= b_cols, CMPC_ERR_ARG, "%s: ldo < b_cols", who);
  CUtensorMap tA, tB;
  rc = make_tmap_3d(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a_f16, a_cols, m, batch, lda * 2, (uint64_t)m * lda * 2, 64, W_BK);
  if (rc) return rc;
  rc = make_tmap_3d(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b_f16, b_cols, m, batch, ldb * 2, (uint64_t)m * ldb * 2, 64, W_BK);
  if (rc) return rc;
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_atb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM);
    CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "cudaFuncSetAttribute(atb, smem=%d): %s", W_SMEM, cudaGetErrorString(e));
  }
  const int ti = (a_cols + W_BI - 1) / W_BI, tj = (b_cols + W_BJ - 1) / W_BJ;
  const int kb_total = (m + W_BK - 1) / W_BK;
  int sp = splits;
  if (sp <= 0) {
    // Split count by a wave model: one CTA per SM (193 KB of shared memory), all CTAs of a launch cost the same, so the kernel takes
    // ceil(CTAs / SMs) waves of (K-blocks per split + a fixed prologue / accumulator-drain cost of ~9 K-blocks); every extra split
    // adds one more pass of fp32 atomics over the whole output.  "Fill the machine about twice over" (the first heuristic) gave 320
    // CTAs = 2.16 waves -- three waves' time -- for the three commonest shapes of the head (1000 x 1000, 2048 x 1000, 1008 x 5000).
    const int sms = num_sms();
    const long long tiles = (long long)ti * tj * batch;
    const int max_sp = (kb_total + 7) / 8 > 0 ? (kb_total + 7) / 8 : 1;       // keep >= 8 K-blocks per CTA
    double best = 1e30;
    sp = 1;
    for (int c = 1; c <= max_sp && c <= 64; ++c) {
      const int kps = (kb_total + c - 1) / c;
      const int c_eff = (kb_total + kps - 1) / kps;                           // no empty split
      if (c_eff != c) continue;
      const long long waves = (tiles * c + sms - 1) / sms;
      const double cost = (double)waves * (kps + 9.0) + 1.5 * c;              // 1.5 K-blocks per additional pass of atomics (measured order)
      if (cost < best - 1e-9) { best = cost; sp = c; }
    }
  }
  if (sp > kb_total) sp = kb_total;
  AtbParams p{};
  p.m = m; p.a_cols = a_cols; p.b_cols = b_cols;
  p.kblocks_per_split = (kb_total + sp - 1) / sp;
  sp = (kb_total + p.kblocks_per_split - 1) / p.kblocks_per_split;       // no empty split
  p.splits = sp;
  p.out = out; p.ldo = ldo; p.out_bstride = out_bstride;
  p.vec4 = ((reinterpret_cast<uintptr_t>(out) & 15) == 0 && ldo % 4 == 0 && out_bstride % 4 == 0) ? 1 : 0;
  CMPC_REQUIRE((long long)sp * batch <= 65535, CMPC_ERR_ARG, "%s: batch * splits exceeds the grid limit", who);
  gemm_atb_kernel<<<dim3(ti, tj, sp * batch), W_THREADS, W_SMEM, stream>>>(tA, tB, p);
  return check_launch("gemm_atb_kernel");
}

extern "C" int cmpc_gemm_atb_f16(const void* a_f16, int64_t lda, int32_t a_cols, const void* b_f16, int64_t ldb, int32_t b_cols,
                                 int32_t m, float* out, int64_t ldo, int32_t splits, void* stream_) {
  return launch_atb(a_f16, lda, a_cols, b_f16, ldb, b_cols, m, 1, out, ldo, 0, splits, (cudaStream_t)stream_, "cmpc_gemm_atb_f16");
}

// per-sample contraction: out[b][i, j] += sum_{n < rows_per_sample} a[b, n, i] * c[b, n, j]   (the skinny [N x T] / [N x C] products of
// the graph-aggregation and affinity backward, CMPC_model.py:362, :384-400)
extern "C" int cmpc_gemm_atb_batched_f16(const void* a_f16, int64_t lda, int32_t a_cols, const void* b_f16, int64_t ldb, int32_t b_cols,
                                         int32_t rows_per_sample, int32_t batch, float* out, int64_t ldo, int64_t out_bstride,
                                         void* stream_) {
  return launch_atb(a_f16, lda, a_cols, b_f16, ldb, b_cols, rows_per_sample, batch, out, ldo, out_bstride, 0, (cudaStream_t)stream_,
                    "cmpc_gemm_atb_batched_f16");
}

// Backward of one text-guided exchange module (CMPC_model.py:194-259 + the l2_normalize of :272-284):
//   out = l2n(s),  s = feat + se1 + se2,  se_i = relu(f_i W_i + b_i) * gate_i[b],  gate_i = sigmoid(gv Wf_i + bf_i),
//   gv = l2n_sample(pool Wg + gvl),  pool = sum_n a_n feat_n,  a = softmax_n(feat_n . u * scale)
// Row kernels (warp per row, a warp walks a contiguous chunk of ONE sample so per-sample column sums stay in registers):
//   exg_bwd_rows  : ds = l2n^T(dout);  dP_i = ds * gate_i * [se_i > 0] (fp16, operand of the dgrad / wgrad GEMMs of the
//                   trans_feat convs);  per-sample column sums  S0 = sum ds*se1/gate1, S1 = sum ds*se2/gate2 (-> d gate),
//                   S2 = sum dP1, S3 = sum dP2 (-> d bias)
//   pool_bwd_rows : dfeat = ds + (lang_se dgrad GEMM output) [+ extra] + a_n dpool + dl_n scale u,
//                   dl_n = a_n (feat_n.dpool - pool.dpool);  per-sample column sums  du = sum_n dl_n scale feat_n
// Small per-sample kernels: gv_gates_bwd (the MLP between the column sums and dpool) and small_atb (parameter gradients
// that are sums over the batch of outer products).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

__device__ __forceinline__ void unpack8h(const uint4 u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8h(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

constexpr int XB_THREADS = 256;

// grid = (chunks, batch); each warp takes rows [r0 + warp, r1) step 8 of the block's chunk of sample blockIdx.y.
// The four per-sample column sums live in a warp-private shared-memory slice [4][MAXG * 256] (a lane only ever touches its own
// columns, so plain read-modify-write): as 64 more registers per thread they left one block per SM and a latency-bound kernel.
template <int MAXG>
__global__ void __launch_bounds__(XB_THREADS, 2)
exg_bwd_rows_kernel(const float* __restrict__ dout, long long ld_dout, const __half* __restrict__ out16, const float* __restrict__ row_ss,
                    const __half* __restrict__ se1, const __half* __restrict__ se2, const float* __restrict__ gate1, const float* __restrict__ gate2,
                    long long gate_bstride, long long ld, float* __restrict__ ds, __half* __restrict__ dp1, __half* __restrict__ dp2,
                    float* __restrict__ colsum /*b: [4, ld] at colsum + b * cs_bstride*/, long long cs_bstride, int rows_per_sample, int rows_per_chunk,
                    int width) {
  extern __shared__ float s_all[];                       // [warps][4][MAXG * 256]
  constexpr int NW = XB_THREADS / 32, W = MAXG * 256;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int groups = width / 8;
  const int p0 = blockIdx.x * rows_per_chunk, p1 = min(rows_per_sample, p0 + rows_per_chunk);
  float* sw = s_all + warp * 4 * W;
  for (int i = lane; i < 4 * W; i += 32) sw[i] = 0.f;
  __syncwarp();
  for (int pix = p0 + warp; pix < p1; pix += NW) {
    const long long r = (long long)b * rows_per_sample + pix;
    float o[MAXG][8], d[MAXG][8];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < groups) {
        unpack8h(__ldg(reinterpret_cast<const uint4*>(out16 + r * ld + g * 8)), o[k]);
        const float4 a = __ldg(reinterpret_cast<const float4*>(dout + r * ld_dout + g * 8)), c = __ldg(reinterpret_cast<const float4*>(dout + r * ld_dout + g * 8 + 4));
        d[k][0] = a.x; d[k][1] = a.y; d[k][2] = a.z; d[k][3] = a.w; d[k][4] = c.x; d[k][5] = c.y; d[k][6] = c.z; d[k][7] = c.w;
#pragma unroll
        for (int e = 0; e < 8; ++e) dot += o[k][e] * d[k][e];
      }
    }
    dot = warp_sum(dot);
    const float inv = rsqrtf(fmaxf(__ldg(row_ss + r), 1e-12f));
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < groups) {
        float s1[8], s2[8], v[8], q1[8], q2[8], g1[8], g2[8];
        unpack8h(__ldg(reinterpret_cast<const uint4*>(se1 + r * ld + g * 8)), s1);
        unpack8h(__ldg(reinterpret_cast<const uint4*>(se2 + r * ld + g * 8)), s2);
        {
          const float4 a = __ldg(reinterpret_cast<const float4*>(gate1 + b * gate_bstride + g * 8)), c = __ldg(reinterpret_cast<const float4*>(gate1 + b * gate_bstride + g * 8 + 4));
          g1[0] = a.x; g1[1] = a.y; g1[2] = a.z; g1[3] = a.w; g1[4] = c.x; g1[5] = c.y; g1[6] = c.z; g1[7] = c.w;
          const float4 a2 = __ldg(reinterpret_cast<const float4*>(gate2 + b * gate_bstride + g * 8)), c2 = __ldg(reinterpret_cast<const float4*>(gate2 + b * gate_bstride + g * 8 + 4));
          g2[0] = a2.x; g2[1] = a2.y; g2[2] = a2.z; g2[3] = a2.w; g2[4] = c2.x; g2[5] = c2.y; g2[6] = c2.z; g2[7] = c2.w;
        }
        float* sc = sw + g * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[e] = (d[k][e] - o[k][e] * dot) * inv;                       // d s
          q1[e] = s1[e] > 0.f ? v[e] * g1[e] : 0.f;                     // d (f1 W1 + b1)
          q2[e] = s2[e] > 0.f ? v[e] * g2[e] : 0.f;
          sc[e] += g1[e] > 0.f ? v[e] * s1[e] / g1[e] : 0.f;            // d gate1 (se1 / gate1 = relu(...))
          sc[W + e] += g2[e] > 0.f ? v[e] * s2[e] / g2[e] : 0.f;
          sc[2 * W + e] += q1[e];
          sc[3 * W + e] += q2[e];
        }
        *reinterpret_cast<float4*>(ds + r * ld + g * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(ds + r * ld + g * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        *reinterpret_cast<uint4*>(dp1 + r * ld + g * 8) = pack8h(q1);
        *reinterpret_cast<uint4*>(dp2 + r * ld + g * 8) = pack8h(q2);
      }
    }
  }
  __syncthreads();
  // fold the 8 warp slices, then one atomic per (quantity, column)
  for (int i = threadIdx.x; i < 4 * W; i += XB_THREADS) {
    const int q = i / W, c = i - q * W;
    if (c < width) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += s_all[w * 4 * W + i];
      atomicAdd(colsum + (long long)b * cs_bstride + (long long)q * ld + c, t);
    }
  }
}

template <int MAXG>
__global__ void __launch_bounds__(XB_THREADS)
pool_bwd_rows_kernel(const __half* __restrict__ feat, long long ld, const float* __restrict__ u, long long u_bstride, const float* __restrict__ pool,
                     const float* __restrict__ dpool, long long vec_bstride, const float* __restrict__ pstats, long long pstats_bstride,
                     float scale, const float* __restrict__ ds, const float* __restrict__ dgemm, long long ld_dgemm,
                     const float* __restrict__ extra, long long ld_extra, float* __restrict__ dfeat, float* __restrict__ du /*[B, ld]*/,
                     long long du_bstride, int rows_per_sample, int rows_per_chunk, int width) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int groups = width / 8;
  const int p0 = blockIdx.x * rows_per_chunk, p1 = min(rows_per_sample, p0 + rows_per_chunk);
  float uu[MAXG][8], dp[MAXG][8], acc[MAXG][8];
  float pd = 0.f;
#pragma unroll
  for (int k = 0; k < MAXG; ++k) {
    const int g = lane + 32 * k;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      uu[k][e] = g < groups ? __ldg(u + b * u_bstride + g * 8 + e) * scale : 0.f;
      dp[k][e] = g < groups ? __ldg(dpool + b * vec_bstride + g * 8 + e) : 0.f;
      pd += g < groups ? __ldg(pool + b * vec_bstride + g * 8 + e) * dp[k][e] : 0.f;
      acc[k][e] = 0.f;
    }
  }
  pd = warp_sum(pd);
  const float mx = __ldg(pstats + b * pstats_bstride), inv_l = 1.0f / __ldg(pstats + b * pstats_bstride + 1);
  for (int pix = p0 + warp; pix < p1; pix += XB_THREADS / 32) {
    const long long r = (long long)b * rows_per_sample + pix;
    float f[MAXG][8];
    float d1 = 0.f, lg = 0.f;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < groups) {
        unpack8h(__ldg(reinterpret_cast<const uint4*>(feat + r * ld + g * 8)), f[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) { d1 += f[k][e] * dp[k][e]; lg += f[k][e] * uu[k][e]; }
      }
    }
    d1 = warp_sum(d1);
    lg = warp_sum(lg);
    const float a = __expf(lg - mx) * inv_l;
    const float dl = a * (d1 - pd);
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < groups) {
        float v[8];
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(ds + r * ld + g * 8)), s1 = __ldg(reinterpret_cast<const float4*>(ds + r * ld + g * 8 + 4));
        const float4 m0 = __ldg(reinterpret_cast<const float4*>(dgemm + r * ld_dgemm + g * 8)), m1 = __ldg(reinterpret_cast<const float4*>(dgemm + r * ld_dgemm + g * 8 + 4));
        v[0] = s0.x + m0.x; v[1] = s0.y + m0.y; v[2] = s0.z + m0.z; v[3] = s0.w + m0.w;
        v[4] = s1.x + m1.x; v[5] = s1.y + m1.y; v[6] = s1.z + m1.z; v[7] = s1.w + m1.w;
        if (extra != nullptr) {
          const float4 x0 = __ldg(reinterpret_cast<const float4*>(extra + r * ld_extra + g * 8)), x1 = __ldg(reinterpret_cast<const float4*>(extra + r * ld_extra + g * 8 + 4));
          v[0] += x0.x; v[1] += x0.y; v[2] += x0.z; v[3] += x0.w; v[4] += x1.x; v[5] += x1.y; v[6] += x1.z; v[7] += x1.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[e] += a * dp[k][e] + dl * uu[k][e];
          acc[k][e] += dl * scale * f[k][e];
        }
        *reinterpret_cast<float4*>(dfeat + r * ld + g * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dfeat + r * ld + g * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  __shared__ float s_acc[XB_THREADS / 32][MAXG * 256];
#pragma unroll
  for (int k = 0; k < MAXG; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) s_acc[warp][(lane + 32 * k) * 8 + e] = acc[k][e];
  __syncthreads();
  for (int c = threadIdx.x; c < width; c += XB_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < XB_THREADS / 32; ++w) t += s_acc[w][c];
    atomicAdd(du + b * du_bstride + c, t);
  }
}

// one block per (sample, module): the small MLP between the lang_se column sums and the pooled vector
//   dpre_i = dgate_i * gate_i (1 - gate_i);  dgv = dpre1 Wf1^T + dpre2 Wf2^T;  z = pool Wg + gvl, gv = z / |z|;
//   dz = (dgv - gv (gv . dgv)) / |z|;  dpool = dz Wg^T.      W* are [k, n] row-major (k = input), mdim x mdim.
constexpr int GB_THREADS = 512;
__global__ void __launch_bounds__(GB_THREADS)
gv_gates_bwd_kernel(const float* __restrict__ colsum /*[B, nmod, 4, ld]*/, const float* __restrict__ gate1, const float* __restrict__ gate2,
                    const float* __restrict__ gv, const float* __restrict__ pool, const float* __restrict__ gvl, long long gvl_bstride,
                    long long gvl_mstride, const float* __restrict__ wg, const float* __restrict__ wf1, const float* __restrict__ wf2,
                    long long w_mstride, int nmod, int Mdim, long long ld, float* __restrict__ dpre1, float* __restrict__ dpre2,
                    float* __restrict__ dz_out, float* __restrict__ dpool, int mode, const float* __restrict__ batch_ss,
                    float* __restrict__ batch_dot) {
  // mode 0: per-sample l2_normalize.  Batch-coupled (forward cmpc_gv_gates_batch): mode 1 emits dpre1/2, parks dgv in dz_out and
  // adds gv . dgv to batch_dot[mod]; mode 2 finishes with |z|^2 = batch_ss[mod] and the batch-wide dot product.
  const int b = blockIdx.x, mod = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = GB_THREADS / 32;
  const long long bm = (long long)b * nmod + mod;
  __shared__ float s_p1[512], s_p2[512], s_dgv[512], s_dz[512], s_pool[512], s_red[NW];
  const float* W1 = wf1 + mod * w_mstride;
  const float* W2 = wf2 + mod * w_mstride;
  const float* WG = wg + mod * w_mstride;
  if (mode == 2) {
    if (tid < Mdim) s_dgv[tid] = dz_out[bm * ld + tid];
  } else {
    if (tid < Mdim) {
      const float a1 = gate1[bm * ld + tid], a2 = gate2[bm * ld + tid];
      const float p1 = colsum[(bm * 4 + 0) * ld + tid] * a1 * (1.f - a1), p2 = colsum[(bm * 4 + 1) * ld + tid] * a2 * (1.f - a2);
      s_p1[tid] = p1; s_p2[tid] = p2;
      dpre1[bm * ld + tid] = p1; dpre2[bm * ld + tid] = p2;
      s_pool[tid] = pool[bm * ld + tid];
    }
    __syncthreads();
    // dgv[k] = sum_n dpre1[n] Wf1[k, n] + dpre2[n] Wf2[k, n]     (warp per row k)
    for (int k = warp; k < Mdim; k += NW) {
      float t = 0.f;
      for (int n = lane; n < Mdim; n += 32) t += s_p1[n] * __ldg(W1 + (long long)k * Mdim + n) + s_p2[n] * __ldg(W2 + (long long)k * Mdim + n);
      t = warp_sum(t);
      if (lane == 0) s_dgv[k] = t;
    }
  }
  float tot = 0.f;
  if (mode == 0) {
    // z[n] = sum_k pool[k] Wg[k, n] + gvl[n]      (thread per column)
    float z = 0.f;
    if (tid < Mdim) {
      for (int k = 0; k < Mdim; ++k) z = fmaf(s_pool[k], __ldg(WG + (long long)k * Mdim + tid), z);
      z += __ldg(gvl + b * gvl_bstride + mod * gvl_mstride + tid);
    }
    float ss = warp_sum(tid < Mdim ? z * z : 0.f);
    if (lane == 0) s_red[warp] = ss;
    __syncthreads();
    for (int w = 0; w < NW; ++w) tot += s_red[w];
  } else if (mode == 2) {
    tot = batch_ss[mod];
  }
  __syncthreads();
  const float invn = rsqrtf(fmaxf(tot, 1e-12f));
  const float gvv = tid < Mdim ? gv[bm * ld + tid] : 0.f;
  float gd = 0.f;
  if (mode == 2) {
    gd = batch_dot[mod];
  } else {
    float dot = warp_sum(tid < Mdim ? gvv * s_dgv[tid] : 0.f);
    if (lane == 0) s_red[warp] = dot;
    __syncthreads();
    for (int w = 0; w < NW; ++w) gd += s_red[w];
    if (mode == 1) {
      if (tid < Mdim) dz_out[bm * ld + tid] = s_dgv[tid];
      if (tid == 0) atomicAdd(batch_dot + mod, gd);
      return;
    }
  }
  if (tid < Mdim) {
    const float dz = (s_dgv[tid] - gvv * gd) * invn;
    s_dz[tid] = dz;
    dz_out[bm * ld + tid] = dz;
  }
  __syncthreads();
  for (int k = warp; k < Mdim; k += NW) {
    float t = 0.f;
    for (int n = lane; n < Mdim; n += 32) t += s_dz[n] * __ldg(WG + (long long)k * Mdim + n);
    t = warp_sum(t);
    if (lane == 0) dpool[bm * ld + k] = t;
  }
}

// out[z][i, j] += sum_b a[z][b, i] * c[z][b, j]      (gradients of the small per-sample linear maps: sums over the batch of
// outer products).  grid = (ceil(ni / 8), ceil(nj / 128), nz), block = 128 threads over j.
__global__ void small_atb_kernel(const float* __restrict__ a, long long lda, long long a_zstride, const float* __restrict__ c, long long ldc,
                                 long long c_zstride, float* __restrict__ out, long long ldo, long long o_zstride, int nb, int ni, int nj) {
  const int z = blockIdx.z;
  const int i0 = blockIdx.x * 8;
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  a += z * a_zstride; c += z * c_zstride; out += z * o_zstride;
  __shared__ float s_a[64][8];                      // a[b0 .. b0+63, i0 .. i0+7]
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int b0 = 0; b0 < nb; b0 += 64) {
    for (int t = threadIdx.x; t < 64 * 8; t += blockDim.x) {
      const int bb = t >> 3, e = t & 7;
      s_a[bb][e] = (b0 + bb < nb && i0 + e < ni) ? __ldg(a + (long long)(b0 + bb) * lda + i0 + e) : 0.f;
    }
    __syncthreads();
    if (j < nj) {
      const int bn = min(64, nb - b0);
      for (int bb = 0; bb < bn; ++bb) {
        const float cv = __ldg(c + (long long)(b0 + bb) * ldc + j);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(s_a[bb][e], cv, acc[e]);
      }
    }
    __syncthreads();
  }
  if (j < nj) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (i0 + e < ni) out[(long long)(i0 + e) * ldo + j] += acc[e];
  }
}

static inline int chunks_for(int batch, int rows_per_sample, int* rows_per_chunk) {
  int chunks = (num_sms() * 4 + batch - 1) / batch;
  if (chunks > rows_per_sample) chunks = rows_per_sample;
  if (chunks < 1) chunks = 1;
  *rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  return (rows_per_sample + *rows_per_chunk - 1) / *rows_per_chunk;
}

}  // namespace cmpc

using namespace cmpc;

#define AL16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int cmpc_exg_bwd_rows(const float* dout, int64_t ld_dout, const void* out_f16, const float* row_sumsq, const void* se1_f16,
                                 const void* se2_f16, const float* gate1, const float* gate2, int64_t gate_bstride, int64_t ld, float* ds,
                                 void* dp1_f16, void* dp2_f16, float* colsum, int64_t colsum_bstride, int32_t batch, int32_t rows_per_sample,
                                 int32_t width, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(dout && out_f16 && row_sumsq && se1_f16 && se2_f16 && gate1 && gate2 && ds && dp1_f16 && dp2_f16 && colsum, CMPC_ERR_ARG,
               "cmpc_exg_bwd_rows: null pointer");
  CMPC_REQUIRE(batch > 0 && rows_per_sample > 0 && width > 0 && width % 8 == 0 && width <= 512 && ld >= width && ld % 8 == 0 && ld_dout % 4 == 0,
               CMPC_ERR_ARG, "cmpc_exg_bwd_rows: width must be a multiple of 8, <= 512");
  CMPC_REQUIRE(AL16(dout) && AL16(out_f16) && AL16(se1_f16) && AL16(se2_f16) && AL16(ds) && AL16(dp1_f16) && AL16(dp2_f16), CMPC_ERR_ALIGN,
               "cmpc_exg_bwd_rows: alignment");
  int rpc;
  const int chunks = chunks_for(batch, rows_per_sample, &rpc);
  dim3 grid(chunks, batch);
  static unsigned long long cfgd = 0;
  if (first_use_on_device(&cfgd)) {
    cudaFuncSetAttribute(exg_bwd_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (XB_THREADS / 32) * 4 * 512 * 4);
  }
  if (width <= 256)
    exg_bwd_rows_kernel<1><<<grid, XB_THREADS, (XB_THREADS / 32) * 4 * 256 * 4, (cudaStream_t)stream>>>(dout, ld_dout, (const __half*)out_f16, row_sumsq, (const __half*)se1_f16,
        (const __half*)se2_f16, gate1, gate2, gate_bstride, ld, ds, (__half*)dp1_f16, (__half*)dp2_f16, colsum, colsum_bstride, rows_per_sample, rpc, width);
  else
    exg_bwd_rows_kernel<2><<<grid, XB_THREADS, (XB_THREADS / 32) * 4 * 512 * 4, (cudaStream_t)stream>>>(dout, ld_dout, (const __half*)out_f16, row_sumsq, (const __half*)se1_f16,
        (const __half*)se2_f16, gate1, gate2, gate_bstride, ld, ds, (__half*)dp1_f16, (__half*)dp2_f16, colsum, colsum_bstride, rows_per_sample, rpc, width);
  return check_launch("exg_bwd_rows_kernel");
}

extern "C" int cmpc_pool_bwd_rows(const void* feat_f16, int64_t ld, const float* u, int64_t u_bstride, const float* pool, const float* dpool, int64_t vec_bstride,
                                  const float* pstats, int64_t pstats_bstride, float scale, const float* ds, const float* dgemm,
                                  int64_t ld_dgemm, const float* extra, int64_t ld_extra, float* dfeat, float* du, int64_t du_bstride,
                                  int32_t batch, int32_t rows_per_sample, int32_t width, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(feat_f16 && u && pool && dpool && pstats && ds && dgemm && dfeat && du, CMPC_ERR_ARG, "cmpc_pool_bwd_rows: null pointer");
  CMPC_REQUIRE(batch > 0 && rows_per_sample > 0 && width > 0 && width % 8 == 0 && width <= 512 && ld >= width && ld % 8 == 0 &&
                   ld_dgemm % 4 == 0 && ld_extra % 4 == 0, CMPC_ERR_ARG, "cmpc_pool_bwd_rows: width must be a multiple of 8, <= 512");
  CMPC_REQUIRE(AL16(feat_f16) && AL16(ds) && AL16(dgemm) && AL16(dfeat) && (!extra || AL16(extra)), CMPC_ERR_ALIGN, "cmpc_pool_bwd_rows: alignment");
  int rpc;
  const int chunks = chunks_for(batch, rows_per_sample, &rpc);
  dim3 grid(chunks, batch);
  if (width <= 256)
    pool_bwd_rows_kernel<1><<<grid, XB_THREADS, 0, (cudaStream_t)stream>>>((const __half*)feat_f16, ld, u, u_bstride, pool, dpool, vec_bstride, pstats,
        pstats_bstride, scale, ds, dgemm, ld_dgemm, extra, ld_extra, dfeat, du, du_bstride, rows_per_sample, rpc, width);
  else
    pool_bwd_rows_kernel<2><<<grid, XB_THREADS, 0, (cudaStream_t)stream>>>((const __half*)feat_f16, ld, u, u_bstride, pool, dpool, vec_bstride, pstats,
        pstats_bstride, scale, ds, dgemm, ld_dgemm, extra, ld_extra, dfeat, du, du_bstride, rows_per_sample, rpc, width);
  return check_launch("pool_bwd_rows_kernel");
}

static int launch_gv_gates_bwd(const float* colsum, const float* gate1, const float* gate2, const float* gv, const float* pool,
                               const float* gvl, int64_t gvl_bstride, int64_t gvl_mstride, const float* wg, const float* wf1,
                               const float* wf2, int64_t w_mstride, int32_t batch, int32_t nmod, int32_t mdim, int64_t ld, float* dpre1,
                               float* dpre2, float* dz, float* dpool, int mode, const float* batch_ss, float* batch_dot, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(colsum && gate1 && gate2 && gv && pool && gvl && wg && wf1 && wf2 && dpre1 && dpre2 && dz && dpool, CMPC_ERR_ARG,
               "cmpc_gv_gates_bwd: null pointer");
  CMPC_REQUIRE(batch > 0 && nmod > 0 && mdim > 0 && mdim <= 512 && ld >= mdim, CMPC_ERR_ARG, "cmpc_gv_gates_bwd: mlp_dim must be <= 512");
  CMPC_REQUIRE(mode == 0 || (batch_ss && batch_dot), CMPC_ERR_ARG, "cmpc_gv_gates_bwd_batch: batch_ss / batch_dot is null");
  gv_gates_bwd_kernel<<<dim3(batch, nmod), GB_THREADS, 0, (cudaStream_t)stream>>>(colsum, gate1, gate2, gv, pool, gvl, gvl_bstride, gvl_mstride,
                                                                                  wg, wf1, wf2, w_mstride, nmod, mdim, ld, dpre1, dpre2, dz, dpool,
                                                                                  mode, batch_ss, batch_dot);
  return check_launch("gv_gates_bwd_kernel");
}

extern "C" int cmpc_gv_gates_bwd(const float* colsum, const float* gate1, const float* gate2, const float* gv, const float* pool,
                                 const float* gvl, int64_t gvl_bstride, int64_t gvl_mstride, const float* wg, const float* wf1,
                                 const float* wf2, int64_t w_mstride, int32_t batch, int32_t nmod, int32_t mdim, int64_t ld, float* dpre1,
                                 float* dpre2, float* dz, float* dpool, void* stream) {
  return launch_gv_gates_bwd(colsum, gate1, gate2, gv, pool, gvl, gvl_bstride, gvl_mstride, wg, wf1, wf2, w_mstride, batch, nmod, mdim, ld,
                             dpre1, dpre2, dz, dpool, 0, nullptr, nullptr, stream);
}

extern "C" int cmpc_gv_gates_bwd_batch(const float* colsum, const float* gate1, const float* gate2, const float* gv, const float* pool,
                                       const float* gvl, int64_t gvl_bstride, int64_t gvl_mstride, const float* wg, const float* wf1,
                                       const float* wf2, int64_t w_mstride, int32_t batch, int32_t nmod, int32_t mdim, int64_t ld,
                                       float* dpre1, float* dpre2, float* dz, float* dpool, int32_t phase, const float* batch_ss,
                                       float* batch_dot, void* stream) {
  CMPC_REQUIRE(phase == 1 || phase == 2, CMPC_ERR_ARG, "cmpc_gv_gates_bwd_batch: phase must be 1 or 2");
  return launch_gv_gates_bwd(colsum, gate1, gate2, gv, pool, gvl, gvl_bstride, gvl_mstride, wg, wf1, wf2, w_mstride, batch, nmod, mdim, ld,
                             dpre1, dpre2, dz, dpool, phase, batch_ss, batch_dot, stream);
}

extern "C" int cmpc_small_atb_f32(const float* a, int64_t lda, int64_t a_zstride, const float* c, int64_t ldc, int64_t c_zstride, float* out,
                                  int64_t ldo, int64_t o_zstride, int32_t nz, int32_t nb, int32_t ni, int32_t nj, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(a && c && out && nz > 0 && nb > 0 && ni > 0 && nj > 0, CMPC_ERR_ARG, "cmpc_small_atb_f32: bad args");
  small_atb_kernel<<<dim3((ni + 7) / 8, (nj + 127) / 128, nz), 128, 0, (cudaStream_t)stream>>>(a, lda, a_zstride, c, ldc, c_zstride, out, ldo, o_zstride, nb, ni,
                                                                                         nj);
  return check_launch("small_atb_kernel");
}

// Row kernels of the per-level backward (CMPC_model.py:330-410: fusion conv, graph_conv with its two whole-sample layer
// norms, the two affinity softmaxes).  All are HBM-bound passes; a warp owns a row, a block walks a contiguous chunk of ONE
// sample, so per-sample and per-channel sums are reduced in registers / shared memory before they touch global memory.
//   relu_mask      : dpre = dout * [act > 0] (fp16 GEMM operand) + per-sample column sums (tiled-language rows, bias)
//   ln_bwd_sums    : [optional l2_normalize^T] -> relu mask -> d(LN output) (fp32, kept) ; per-sample sums of g and g*xhat
//                    (g = dLN*gamma), per-channel dgamma / dbeta
//   ln_bwd_apply   : d(LN input) = rstd (g - mean g - xhat mean(g xhat)) as fp16 (GEMM operand) + per-channel column sums
//   affinity_bwd   : backward of softmax_T (gw_w) and mask * softmax_N (gw_v) and of the relation gate (:388-399)
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

__device__ __forceinline__ void up8(const uint4 u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pk8(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ void ld8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

constexpr int LV_THREADS = 256;
constexpr int LV_WARPS = LV_THREADS / 32;

// block-level column reduction of per-lane accumulators (lane owns column groups lane + 32k) followed by atomics
template <int MAXG>
__device__ __forceinline__ void block_colsum(const float (&acc)[MAXG][8], float* s_acc /*[LV_WARPS][MAXG*256]*/, float* dst, int width) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < MAXG; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) s_acc[warp * (MAXG * 256) + (lane + 32 * k) * 8 + e] = acc[k][e];
  __syncthreads();
  for (int c = threadIdx.x; c < width; c += LV_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < LV_WARPS; ++w) t += s_acc[w * (MAXG * 256) + c];
    atomicAdd(dst + c, t);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
template <int MAXG>
__global__ void __launch_bounds__(LV_THREADS)
relu_mask_kernel(const float* __restrict__ dout, long long ld_d, const __half* __restrict__ act, long long ld, __half* __restrict__ dpre,
                 float* __restrict__ colsum /*[B, ld]*/, int rows_per_sample, int rows_per_chunk, int width) {
  extern __shared__ float s_acc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
  const int groups = width / 8, ogroups = (int)(ld / 8);
  const int p0 = blockIdx.x * rows_per_chunk, p1 = min(rows_per_sample, p0 + rows_per_chunk);
  float acc[MAXG][8];
#pragma unroll
  for (int k = 0; k < MAXG; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  for (int pix = p0 + warp; pix < p1; pix += LV_WARPS) {
    const long long r = (long long)b * rows_per_sample + pix;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < ogroups) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (g < groups) {
          float a[8], d[8];
          up8(__ldg(reinterpret_cast<const uint4*>(act + r * ld + g * 8)), a);
          ld8(dout + r * ld_d + g * 8, d);
#pragma unroll
          for (int e = 0; e < 8; ++e) { v[e] = a[e] > 0.f ? d[e] : 0.f; acc[k][e] += v[e]; }
        }
        *reinterpret_cast<uint4*>(dpre + r * ld + g * 8) = pk8(v);
      }
    }
  }
  block_colsum<MAXG>(acc, s_acc, colsum + (long long)b * ld, width);
}

// row_ss != NULL: dout is the gradient of l2_normalize(act0) given act = the NORMALISED map and row_ss = |act0|^2.
// dgamma / dbeta partial sums live in warp-private shared-memory slices, gamma in a block-shared one (as registers they cost
// 96 per thread at C = 1000 and left one block per SM).
template <int MAXG>
__global__ void __launch_bounds__(LV_THREADS, 2)
ln_bwd_sums_kernel(const float* __restrict__ dout, long long ld_d, const __half* __restrict__ act, const float* __restrict__ row_ss,
                   const __half* __restrict__ pre, long long ld, const float* __restrict__ mr /*[B,2]*/, const float* __restrict__ gamma,
                   float* __restrict__ dln, long long ld_ln, double* __restrict__ sums /*[B,2]*/, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, int rows_per_sample, int rows_per_chunk, int width) {
  extern __shared__ float s_acc[];                       // [warps][2][W] then gamma [W]
  constexpr int W = MAXG * 256;
  __shared__ float s_s[LV_WARPS][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
  const int groups = width / 8;
  const int p0 = blockIdx.x * rows_per_chunk, p1 = min(rows_per_sample, p0 + rows_per_chunk);
  const float2 ms = __ldg(reinterpret_cast<const float2*>(mr) + b);
  float* sw = s_acc + warp * 2 * W;
  float* s_gm = s_acc + LV_WARPS * 2 * W;
  for (int i = lane; i < 2 * W; i += 32) sw[i] = 0.f;
  for (int i = threadIdx.x; i < W; i += LV_THREADS) s_gm[i] = i < width ? __ldg(gamma + i) : 0.f;
  __syncthreads();
  float s0 = 0.f, s1 = 0.f;
  for (int pix = p0 + warp; pix < p1; pix += LV_WARPS) {
    const long long r = (long long)b * rows_per_sample + pix;
    float a[MAXG][8], d[MAXG][8];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < groups) {
        up8(__ldg(reinterpret_cast<const uint4*>(act + r * ld + g * 8)), a[k]);
        ld8(dout + r * ld_d + g * 8, d[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) dot += a[k][e] * d[k][e];
      }
    }
    float inv = 1.f;
    if (row_ss != nullptr) {
      dot = warp_sum(dot);
      inv = rsqrtf(fmaxf(__ldg(row_ss + r), 1e-12f));
    }
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < groups) {
        float x[8], v[8];
        up8(__ldg(reinterpret_cast<const uint4*>(pre + r * ld + g * 8)), x);
        float* sc = sw + g * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float t = row_ss != nullptr ? (d[k][e] - a[k][e] * dot) * inv : d[k][e];
          t = a[k][e] > 0.f ? t : 0.f;                               // relu
          v[e] = t;
          const float xh = (x[e] - ms.x) * ms.y;
          const float gg = t * s_gm[g * 8 + e];
          s0 += gg; s1 += gg * xh;
          sc[e] += t * xh; sc[W + e] += t;
        }
        st8(dln + r * ld_ln + g * 8, v);
      }
    }
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if (lane == 0) { s_s[warp][0] = s0; s_s[warp][1] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < LV_WARPS; ++w) t += (double)s_s[w][threadIdx.x];
    atomicAdd(sums + b * 2 + threadIdx.x, t);
  }
  for (int i = threadIdx.x; i < 2 * W; i += LV_THREADS) {
    const int q = i / W, c = i - q * W;
    if (c < width) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LV_WARPS; ++w) t += s_acc[w * 2 * W + i];
      atomicAdd((q ? dbeta : dgamma) + c, t);
    }
  }
}

template <int MAXG>
__global__ void __launch_bounds__(LV_THREADS)
ln_bwd_apply_kernel(const float* __restrict__ dln, long long ld_ln, const __half* __restrict__ pre, long long ld, const float* __restrict__ mr,
                    const float* __restrict__ gamma, const double* __restrict__ sums, float inv_count, __half* __restrict__ out,
                    float* __restrict__ colsum /*[ld] or null*/, int rows_per_sample, int rows_per_chunk, int width) {
  extern __shared__ float s_acc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
  const int groups = width / 8, ogroups = (int)(ld / 8);
  const int p0 = blockIdx.x * rows_per_chunk, p1 = min(rows_per_sample, p0 + rows_per_chunk);
  const float2 ms = __ldg(reinterpret_cast<const float2*>(mr) + b);
  const float m1 = (float)(sums[b * 2] * (double)inv_count), m2 = (float)(sums[b * 2 + 1] * (double)inv_count);
  float gm[MAXG][8], acc[MAXG][8];
#pragma unroll
  for (int k = 0; k < MAXG; ++k) {
    const int g = lane + 32 * k;
#pragma unroll
    for (int e = 0; e < 8; ++e) { gm[k][e] = g < groups ? __ldg(gamma + g * 8 + e) : 0.f; acc[k][e] = 0.f; }
  }
  for (int pix = p0 + warp; pix < p1; pix += LV_WARPS) {
    const long long r = (long long)b * rows_per_sample + pix;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < ogroups) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (g < groups) {
          float x[8], d[8];
          up8(__ldg(reinterpret_cast<const uint4*>(pre + r * ld + g * 8)), x);
          ld8(dln + r * ld_ln + g * 8, d);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xh = (x[e] - ms.x) * ms.y;
            v[e] = ms.y * (d[e] * gm[k][e] - m1 - xh * m2);
            acc[k][e] += v[e];
          }
        }
        *reinterpret_cast<uint4*>(out + r * ld + g * 8) = pk8(v);
      }
    }
  }
  if (colsum != nullptr) block_colsum<MAXG>(acc, s_acc, colsum, width);
}

// ---------------------------------------------------------------------------------------------------------------------
// affinity softmaxes (:388-399).  W = softmax_T(masked), V = mask * softmax_N: with dW, dV (fp32 [rows, 32]):
//   cs[b, t] = sum_n V dV (colsum kernel);  dA = W (dW - sum_t W dW) + V (dV - cs_t);  d raw = dA * r_t;  dr_t += sum_n dA * raw
__global__ void affinity_bwd_colsum_kernel(const __half* __restrict__ v16, const float* __restrict__ dv, float inv_vscale, int rows_per_sample,
                                           float* __restrict__ cs /*[B,32]*/) {
  const int b = blockIdx.y, t = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const int per = (rows_per_sample + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(rows_per_sample, p0 + per);
  float acc = 0.f;
  for (int n = p0 + sub; n < p1; n += blockDim.x / 32) {
    const long long r = (long long)b * rows_per_sample + n;
    acc += __half2float(v16[r * 32 + t]) * inv_vscale * dv[r * 32 + t];
  }
  __shared__ float s[8][32];
  s[sub][t] = acc;
  __syncthreads();
  if (sub == 0) {
    float tt = 0.f;
    for (int w = 0; w < (int)(blockDim.x / 32); ++w) tt += s[w][t];
    atomicAdd(cs + b * 32 + t, tt);
  }
}

__global__ void affinity_bwd_kernel(const __half* __restrict__ w16, const __half* __restrict__ v16, const float* __restrict__ dw,
                                    const float* __restrict__ dv, const float* __restrict__ affi, const float* __restrict__ rgate,
                                    const float* __restrict__ cs, float inv_vscale, int rows_per_sample, __half* __restrict__ draw16,
                                    float* __restrict__ drgate /*[B,32]*/) {
  const int b = blockIdx.y, t = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const int per = (rows_per_sample + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(rows_per_sample, p0 + per);
  const float rg = rgate[b * 32 + t], c = cs[b * 32 + t];
  float acc = 0.f;
  for (int n = p0 + sub; n < p1; n += blockDim.x / 32) {
    const long long r = (long long)b * rows_per_sample + n;
    const float W = __half2float(w16[r * 32 + t]), V = __half2float(v16[r * 32 + t]) * inv_vscale;
    const float dW = dw[r * 32 + t], dV = dv[r * 32 + t];
    const float rowdot = warp_sum(W * dW);
    const float dA = W * (dW - rowdot) + V * (dV - c);
    draw16[r * 32 + t] = __float2half_rn(dA * rg);
    acc += rg != 0.f ? dA * affi[r * 32 + t] / rg : 0.f;
  }
  __shared__ float s[8][32];
  s[sub][t] = acc;
  __syncthreads();
  if (sub == 0) {
    float tt = 0.f;
    for (int w = 0; w < (int)(blockDim.x / 32); ++w) tt += s[w][t];
    atomicAdd(drgate + b * 32 + t, tt);
  }
}

// gtT[b, c, t] = gt[b, t, c]  (fp16; [B, T rows, ld] -> [B, rows_out, 64], columns >= T zero): the B operand of d x = d raw . Gt
__global__ void transpose_gt_kernel(const __half* __restrict__ gt, long long ld, int T, int rows_out, __half* __restrict__ gtT) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= rows_out) return;
  __half* o = gtT + ((long long)b * rows_out + c) * 64;
  for (int t = 0; t < 64; ++t) o[t] = (t < T && c < ld) ? gt[((long long)b * T + t) * ld + c] : __float2half_rn(0.f);
}

// MUTAN output side (:319-324): x = m / |m|, m = tanh(s).  dx = sum of up to four fp32 pieces; ds = (dx - x (x.dx)) / |m| * (1 - m^2)
template <int MAXG>
__global__ void __launch_bounds__(LV_THREADS)
mutan_out_bwd_kernel(const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2, const float* __restrict__ p3,
                     long long ldp, const __half* __restrict__ x16, long long ld, const float* __restrict__ row_ss, float* __restrict__ ds,
                     long long ld_ds, long long rows, int width) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int groups = width / 8;
  for (long long r = warp0; r < rows; r += nwarps) {
    float x[MAXG][8], d[MAXG][8];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < groups) {
        up8(__ldg(reinterpret_cast<const uint4*>(x16 + r * ld + g * 8)), x[k]);
        ld8(p0 + r * ldp + g * 8, d[k]);
        float t[8];
        if (p1) { ld8(p1 + r * ldp + g * 8, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) d[k][e] += t[e]; }
        if (p2) { ld8(p2 + r * ldp + g * 8, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) d[k][e] += t[e]; }
        if (p3) { ld8(p3 + r * ldp + g * 8, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) d[k][e] += t[e]; }
#pragma unroll
        for (int e = 0; e < 8; ++e) dot += x[k][e] * d[k][e];
      }
    }
    dot = warp_sum(dot);
    const float ss = fmaxf(__ldg(row_ss + r), 1e-12f);
    const float inv = rsqrtf(ss), nrm = ss * inv;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < groups) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float m = x[k][e] * nrm;
          v[e] = (d[k][e] - x[k][e] * dot) * inv * (1.f - m * m);
        }
        st8(ds + r * ld_ds + g * 8, v);
      }
    }
  }
}

// lateral l2_normalize folded into the MUTAN GEMM (:109-113): with G = (d pre * rsc) . Wv^T, d xlat = G - xlat * (xlat . G) / |xlat|^2
template <int MAXG>
__global__ void __launch_bounds__(LV_THREADS)
lateral_bwd_kernel(const float* __restrict__ G, long long ldg, const __half* __restrict__ xlat, long long ld, const float* __restrict__ row_ss,
                   __half* __restrict__ out, float* __restrict__ colsum, int rows_per_sample, int rows_per_chunk, int width) {
  extern __shared__ float s_acc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
  const int groups = width / 8, ogroups = (int)(ld / 8);
  const int p0 = blockIdx.x * rows_per_chunk, p1 = min(rows_per_sample, p0 + rows_per_chunk);
  float acc[MAXG][8];
#pragma unroll
  for (int k = 0; k < MAXG; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  for (int pix = p0 + warp; pix < p1; pix += LV_WARPS) {
    const long long r = (long long)b * rows_per_sample + pix;
    float x[MAXG][8], g8[MAXG][8];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < groups) {
        up8(__ldg(reinterpret_cast<const uint4*>(xlat + r * ld + g * 8)), x[k]);
        ld8(G + r * ldg + g * 8, g8[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) dot += x[k][e] * g8[k][e];
      }
    }
    dot = warp_sum(dot) / fmaxf(__ldg(row_ss + r), 1e-12f);
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
      const int g = lane + 32 * k;
      if (g < ogroups) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (g < groups) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { v[e] = g8[k][e] - x[k][e] * dot; acc[k][e] += v[e]; }
        }
        *reinterpret_cast<uint4*>(out + r * ld + g * 8) = pk8(v);
      }
    }
  }
  block_colsum<MAXG>(acc, s_acc, colsum, width);
}

// out = dy * (1 - y^2)  (tanh) or dy * y * (1 - y) (sigmoid), small fp32 vectors of the language side
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ out, long long n, int kind) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = y[i];
    out[i] = dy[i] * (kind == 2 ? (1.f - v * v) : v * (1.f - v));
  }
}

static inline int lv_chunks(int batch, int rows_per_sample, int* rows_per_chunk) {
  int chunks = (num_sms() * 4 + batch - 1) / batch;
  if (chunks > rows_per_sample) chunks = rows_per_sample;
  if (chunks < 1) chunks = 1;
  *rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  return (rows_per_sample + *rows_per_chunk - 1) / *rows_per_chunk;
}

}  // namespace cmpc

using namespace cmpc;

#define LV_DISPATCH(kernel, ld_, ...)                                                                               \
  do {                                                                                                              \
    const size_t sm1 = (size_t)LV_WARPS * 256 * sizeof(float);                                                      \
    if ((ld_) <= 256) kernel<1><<<grid, LV_THREADS, sm1, (cudaStream_t)stream>>>(__VA_ARGS__);                      \
    else if ((ld_) <= 512) kernel<2><<<grid, LV_THREADS, 2 * sm1, (cudaStream_t)stream>>>(__VA_ARGS__);             \
    else {                                                                                                          \
      static unsigned long long cfgd = 0;                                                                                     \
      if (first_use_on_device(&cfgd)) { cudaFuncSetAttribute(kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sm1)); } \
      kernel<4><<<grid, LV_THREADS, 4 * sm1, (cudaStream_t)stream>>>(__VA_ARGS__);                                  \
    }                                                                                                               \
  } while (0)

extern "C" int cmpc_relu_mask_f16(const float* dout, int64_t ld_d, const void* act_f16, int64_t ld, void* dpre_f16, float* colsum,
                                  int32_t batch, int32_t rows_per_sample, int32_t width, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(dout && act_f16 && dpre_f16 && colsum && batch > 0 && rows_per_sample > 0 && width > 0 && width % 8 == 0 && ld % 8 == 0 &&
                   ld >= width && ld <= 1024 && ld_d % 4 == 0, CMPC_ERR_ARG, "cmpc_relu_mask_f16: bad args (width %% 8, ld <= 1024)");
  int rpc;
  dim3 grid(lv_chunks(batch, rows_per_sample, &rpc), batch);
  LV_DISPATCH(relu_mask_kernel, ld, dout, ld_d, (const __half*)act_f16, ld, (__half*)dpre_f16, colsum, rows_per_sample, rpc, width);
  return check_launch("relu_mask_kernel");
}

extern "C" int cmpc_ln_bwd_sums(const float* dout, int64_t ld_d, const void* act_f16, const float* row_sumsq, const void* pre_f16, int64_t ld,
                                const float* mean_rstd, const float* gamma, float* dln, int64_t ld_ln, double* sums, float* dgamma,
                                float* dbeta, int32_t batch, int32_t rows_per_sample, int32_t width, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(dout && act_f16 && pre_f16 && mean_rstd && gamma && dln && sums && dgamma && dbeta, CMPC_ERR_ARG, "cmpc_ln_bwd_sums: null pointer");
  CMPC_REQUIRE(batch > 0 && rows_per_sample > 0 && width > 0 && width % 8 == 0 && ld % 8 == 0 && ld >= width && ld <= 1024 && ld_d % 4 == 0 &&
                   ld_ln % 4 == 0, CMPC_ERR_ARG, "cmpc_ln_bwd_sums: bad shape");
  int rpc;
  dim3 grid(lv_chunks(batch, rows_per_sample, &rpc), batch);
  {
    const size_t per = (size_t)(LV_WARPS * 2 + 1) * 256 * sizeof(float);       // x MAXG
    static unsigned long long cfgd = 0;
    if (first_use_on_device(&cfgd)) {
      cudaFuncSetAttribute(ln_bwd_sums_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * per));
      cudaFuncSetAttribute(ln_bwd_sums_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * per));
    }
    if (ld <= 256)
      ln_bwd_sums_kernel<1><<<grid, LV_THREADS, per, (cudaStream_t)stream>>>(dout, ld_d, (const __half*)act_f16, row_sumsq, (const __half*)pre_f16, ld,
                                                                             mean_rstd, gamma, dln, ld_ln, sums, dgamma, dbeta, rows_per_sample, rpc, width);
    else if (ld <= 512)
      ln_bwd_sums_kernel<2><<<grid, LV_THREADS, 2 * per, (cudaStream_t)stream>>>(dout, ld_d, (const __half*)act_f16, row_sumsq, (const __half*)pre_f16, ld,
                                                                                 mean_rstd, gamma, dln, ld_ln, sums, dgamma, dbeta, rows_per_sample, rpc, width);
    else
      ln_bwd_sums_kernel<4><<<grid, LV_THREADS, 4 * per, (cudaStream_t)stream>>>(dout, ld_d, (const __half*)act_f16, row_sumsq, (const __half*)pre_f16, ld,
                                                                                 mean_rstd, gamma, dln, ld_ln, sums, dgamma, dbeta, rows_per_sample, rpc, width);
  }
  return check_launch("ln_bwd_sums_kernel");
}

extern "C" int cmpc_ln_bwd_apply(const float* dln, int64_t ld_ln, const void* pre_f16, int64_t ld, const float* mean_rstd, const float* gamma,
                                 const double* sums, void* out_f16, float* colsum, int32_t batch, int32_t rows_per_sample, int32_t width,
                                 void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(dln && pre_f16 && mean_rstd && gamma && sums && out_f16, CMPC_ERR_ARG, "cmpc_ln_bwd_apply: null pointer");
  CMPC_REQUIRE(batch > 0 && rows_per_sample > 0 && width > 0 && width % 8 == 0 && ld % 8 == 0 && ld >= width && ld <= 1024 && ld_ln % 4 == 0,
               CMPC_ERR_ARG, "cmpc_ln_bwd_apply: bad shape");
  int rpc;
  dim3 grid(lv_chunks(batch, rows_per_sample, &rpc), batch);
  const float inv_count = 1.0f / ((float)rows_per_sample * (float)width);
  LV_DISPATCH(ln_bwd_apply_kernel, ld, dln, ld_ln, (const __half*)pre_f16, ld, mean_rstd, gamma, sums, inv_count, (__half*)out_f16, colsum,
              rows_per_sample, rpc, width);
  return check_launch("ln_bwd_apply_kernel");
}

extern "C" int cmpc_affinity_bwd(const void* w_f16, const void* v_f16, const float* dw, const float* dv, const float* affi, const float* rgate,
                                 float v_scale, int32_t batch, int32_t rows_per_sample, float* colsum_ws /*[B,32], zeroed here*/,
                                 void* draw_f16, float* drgate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(w_f16 && v_f16 && dw && dv && affi && rgate && colsum_ws && draw_f16 && drgate && batch > 0 && rows_per_sample > 0 && v_scale > 0.f,
               CMPC_ERR_ARG, "cmpc_affinity_bwd: bad args");
  cudaError_t e = cudaMemsetAsync(colsum_ws, 0, (size_t)batch * 32 * sizeof(float), stream);
  CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "cmpc_affinity_bwd: memset: %s", cudaGetErrorString(e));
  int chunks = (num_sms() * 2 + batch - 1) / batch;
  if (chunks > (rows_per_sample + 7) / 8) chunks = (rows_per_sample + 7) / 8;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, batch);
  affinity_bwd_colsum_kernel<<<grid, 256, 0, stream>>>((const __half*)v_f16, dv, 1.0f / v_scale, rows_per_sample, colsum_ws);
  rc = check_launch("affinity_bwd_colsum_kernel");
  if (rc) return rc;
  affinity_bwd_kernel<<<grid, 256, 0, stream>>>((const __half*)w_f16, (const __half*)v_f16, dw, dv, affi, rgate, colsum_ws, 1.0f / v_scale,
                                                rows_per_sample, (__half*)draw_f16, drgate);
  return check_launch("affinity_bwd_kernel");
}

extern "C" int cmpc_transpose_gt_f16(const void* gt_f16, int64_t ld, int32_t batch, int32_t t, int32_t rows_out, void* gtT_f16, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(gt_f16 && gtT_f16 && batch > 0 && t > 0 && t <= 32 && rows_out > 0, CMPC_ERR_ARG, "cmpc_transpose_gt_f16: bad args");
  transpose_gt_kernel<<<dim3((rows_out + 127) / 128, batch), 128, 0, (cudaStream_t)stream>>>((const __half*)gt_f16, ld, t, rows_out, (__half*)gtT_f16);
  return check_launch("transpose_gt_kernel");
}

extern "C" int cmpc_mutan_out_bwd(const float* p0, const float* p1, const float* p2, const float* p3, int64_t ldp, const void* x_f16, int64_t ld,
                                  const float* row_sumsq, float* ds, int64_t ld_ds, int64_t rows, int32_t width, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(p0 && x_f16 && row_sumsq && ds && rows > 0 && width > 0 && width % 8 == 0 && ld % 8 == 0 && ld <= 1024 && ldp % 4 == 0 && ld_ds % 4 == 0,
               CMPC_ERR_ARG, "cmpc_mutan_out_bwd: bad args");
  long long blocks = (rows * 32 + LV_THREADS - 1) / LV_THREADS;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const dim3 grid((unsigned)blocks);
  if (ld <= 256) mutan_out_bwd_kernel<1><<<grid, LV_THREADS, 0, (cudaStream_t)stream>>>(p0, p1, p2, p3, ldp, (const __half*)x_f16, ld, row_sumsq, ds, ld_ds, rows, width);
  else if (ld <= 512) mutan_out_bwd_kernel<2><<<grid, LV_THREADS, 0, (cudaStream_t)stream>>>(p0, p1, p2, p3, ldp, (const __half*)x_f16, ld, row_sumsq, ds, ld_ds, rows, width);
  else mutan_out_bwd_kernel<4><<<grid, LV_THREADS, 0, (cudaStream_t)stream>>>(p0, p1, p2, p3, ldp, (const __half*)x_f16, ld, row_sumsq, ds, ld_ds, rows, width);
  return check_launch("mutan_out_bwd_kernel");
}

extern "C" int cmpc_lateral_bwd(const float* g, int64_t ldg, const void* xlat_f16, int64_t ld, const float* row_sumsq, void* out_f16, float* colsum,
                                int32_t batch, int32_t rows_per_sample, int32_t width, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(g && xlat_f16 && row_sumsq && out_f16 && colsum && batch > 0 && rows_per_sample > 0 && width > 0 && width % 8 == 0 && ld % 8 == 0 &&
                   ld <= 1024 && ldg % 4 == 0, CMPC_ERR_ARG, "cmpc_lateral_bwd: bad args");
  int rpc;
  dim3 grid(lv_chunks(batch, rows_per_sample, &rpc), batch);
  LV_DISPATCH(lateral_bwd_kernel, ld, g, ldg, (const __half*)xlat_f16, ld, row_sumsq, (__half*)out_f16, colsum, rows_per_sample, rpc, width);
  return check_launch("lateral_bwd_kernel");
}

extern "C" int cmpc_act_bwd_f32(const float* dy, const float* y, float* out, int64_t n, int32_t kind, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(dy && y && out && n > 0 && (kind == 2 || kind == 3), CMPC_ERR_ARG, "cmpc_act_bwd_f32: bad args (kind 2 = tanh, 3 = sigmoid)");
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  act_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dy, y, out, n, kind);
  return check_launch("act_bwd_kernel");
}

// Score head: 3x3 conv M -> 1 (SAME, zero pad), TF-1 legacy bilinear upsampling (align_corners=False, no
// half-pixel centres), sigmoid (CMPC_model.py:138-142; aux heads :128-133), and the integer mask I/U counts
// of the evaluation path (CMPC_model.py:486-489, util/eval_tools.py:31-35).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

// warp per pixel: taps[pix, k] = sum_c feat[pix, c] * w[k, c],  k = 3*dy + dx
template <int MAXG>
__global__ void score_taps_kernel(const __half* __restrict__ feat, long long ld, const float* __restrict__ w /*[9, ld]*/,
                                  int width, long long rows, float* __restrict__ taps /*[rows, 9]*/) {
  extern __shared__ float s_w[];   // [9][width]
  for (int i = threadIdx.x; i < 9 * width; i += blockDim.x) s_w[i] = __ldg(w + (long long)(i / width) * ld + (i % width));
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int groups = width / 8;
  for (long long r = warp0; r < rows; r += nwarps) {
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
#pragma unroll
    for (int gi = 0; gi < MAXG; ++gi) {
      const int g = lane + 32 * gi;
      if (g < groups) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(feat + r * ld + g * 8));
        const __half2* h = reinterpret_cast<const __half2*>(&u);
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float* ww = s_w + k * width + g * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[k] += f[e] * ww[e];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = warp_sum(acc[k]);
    if (lane < 9) {
      float v = acc[0];
#pragma unroll
      for (int k = 1; k < 9; ++k) v = (lane == k) ? acc[k] : v;
      taps[r * 9 + lane] = v;
    }
  }
}

__global__ void score_pred_kernel(const float* __restrict__ taps, int ldt, float bias, const float* __restrict__ bias_dev, int B, int h,
                                  int w, float* __restrict__ pred) {
  if (bias_dev != nullptr) bias = __ldg(bias_dev);       // training: the bias lives on the device (no host round trip per step)
  const long long total = (long long)B * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const long long b = i / ((long long)w * h);
    float v = bias;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        if (xx < 0 || xx >= w) continue;
        v += __ldg(taps + ((b * h + yy) * w + xx) * ldt + dy * 3 + dx);
      }
    }
    pred[i] = v;
  }
}

// thread per 4 consecutive output pixels of one row
__global__ void upsample_sigmoid_kernel(const float* __restrict__ pred, int B, int h, int w, int H, int W,
                                        float* __restrict__ up, float* __restrict__ sigm) {
  const int wq = W / 4;
  const long long total = (long long)B * H * wq;
  const float ys = (float)h / (float)H, xs = (float)w / (float)W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(i % wq);
    const int Y = (int)((i / wq) % H);
    const long long b = i / ((long long)wq * H);
    const float in_y = (float)Y * ys;
    const int y0 = (int)floorf(in_y);
    const int y1 = min(y0 + 1, h - 1);
    const float ly = in_y - (float)y0;
    const float* p0 = pred + (b * h + y0) * w;
    const float* p1 = pred + (b * h + y1) * w;
    float o[4], s[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int X = xq * 4 + e;
      const float in_x = (float)X * xs;
      const int x0 = (int)floorf(in_x);
      const int x1 = min(x0 + 1, w - 1);
      const float lx = in_x - (float)x0;
      const float tl = __ldg(p0 + x0), tr = __ldg(p0 + x1), bl = __ldg(p1 + x0), br = __ldg(p1 + x1);
      const float top = tl + (tr - tl) * lx;
      const float bot = bl + (br - bl) * lx;
      o[e] = top + (bot - top) * ly;
      s[e] = 1.0f / (1.0f + expf(-o[e]));
    }
    const long long off = (b * H + Y) * W + xq * 4;
    *reinterpret_cast<float4*>(up + off) = make_float4(o[0], o[1], o[2], o[3]);
    if (sigm) *reinterpret_cast<float4*>(sigm + off) = make_float4(s[0], s[1], s[2], s[3]);
  }
}

// per-sample integer intersection / union of (up > thresh) vs (target != 0)
__global__ void iou_counts_kernel(const float* __restrict__ up, const float* __restrict__ target, long long per_sample,
                                  float thresh, int inclusive, unsigned long long* __restrict__ iu /*[B,2]*/) {
  const int b = blockIdx.y;
  const float* u = up + (long long)b * per_sample;
  const float* t = target + (long long)b * per_sample;
  unsigned int ci = 0, cu = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(u + i);
    const bool p = inclusive ? (v >= thresh) : (v > thresh);
    const bool l = __ldg(t + i) != 0.f;
    ci += (p && l) ? 1u : 0u;
    cu += (p || l) ? 1u : 0u;
  }
  ci = __reduce_add_sync(0xffffffffu, ci);
  cu = __reduce_add_sync(0xffffffffu, cu);
  __shared__ unsigned int s_i[32], s_u[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_i[warp] = ci; s_u[warp] = cu; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    unsigned int a = lane < nw ? s_i[lane] : 0u, c = lane < nw ? s_u[lane] : 0u;
    a = __reduce_add_sync(0xffffffffu, a);
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) {
      atomicAdd(iu + 2 * b, (unsigned long long)a);
      atomicAdd(iu + 2 * b + 1, (unsigned long long)c);
    }
  }
}

// per-sample sum over pixels of sigmoid_cross_entropy_with_logits(x, z) = max(x,0) - x z + log1p(exp(-|x|))
// (util/loss.py:12-14 with pos/neg multipliers 1); fp32 per thread, fp64 across the block and across blocks
__global__ void sigmoid_ce_sums_kernel(const float* __restrict__ logits, const float* __restrict__ target, long long per_sample,
                                       double* __restrict__ sums /*[B]*/) {
  const int b = blockIdx.y;
  const float* x = logits + (long long)b * per_sample;
  const float* z = target + (long long)b * per_sample;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x) {
    const float xv = __ldg(x + i), zv = __ldg(z + i);
    acc += fmaxf(xv, 0.f) - xv * zv + log1pf(expf(-fabsf(xv)));
  }
  double d = (double)warp_sum(acc);
  __shared__ double s_d[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_d[warp] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_d[w];
    atomicAdd(sums + b, t);
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_sigmoid_ce_sums(const float* logits, const float* target, int32_t batch, int64_t per_sample, double* sums,
                                    void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(logits && target && sums && batch > 0 && per_sample > 0, CMPC_ERR_ARG, "cmpc_sigmoid_ce_sums: bad args");
  int bx = (int)((per_sample + 256 * 8 - 1) / (256 * 8));
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  sigmoid_ce_sums_kernel<<<dim3(bx, batch), 256, 0, (cudaStream_t)stream>>>(logits, target, per_sample, sums);
  return check_launch("sigmoid_ce_sums_kernel");
}

extern "C" size_t cmpc_score_workspace_bytes(int64_t rows) { return (size_t)rows * 9 * sizeof(float); }

extern "C" int cmpc_score_upsample(const void* feat_f16, int64_t ld, const float* w9, float bias, int32_t batch, int32_t h,
                                   int32_t w, int32_t width, int32_t out_h, int32_t out_w, float* pred, float* up,
                                   float* sigm, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(feat_f16 && w9 && pred && workspace && batch > 0 && h > 0 && w > 0, CMPC_ERR_ARG, "cmpc_score_upsample: bad args");
  CMPC_REQUIRE(width > 0 && width % 8 == 0 && width <= 1024 && ld >= width && ld % 8 == 0, CMPC_ERR_ARG,
               "cmpc_score_upsample: width must be a multiple of 8, <= 1024");
  CMPC_REQUIRE(up == nullptr || (out_w % 4 == 0 && out_h > 0), CMPC_ERR_ARG, "cmpc_score_upsample: W must be a multiple of 4");
  const long long rows = (long long)batch * h * w;
  CMPC_REQUIRE(workspace_bytes >= cmpc_score_workspace_bytes(rows), CMPC_ERR_WORKSPACE, "cmpc_score_upsample: workspace too small");
  float* taps = (float*)workspace;
  const int threads = 256;
  long long blocks = (rows * 32 + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)9 * width * sizeof(float);
  if (width <= 256) score_taps_kernel<1><<<(int)blocks, threads, smem, stream>>>((const __half*)feat_f16, ld, w9, width, rows, taps);
  else if (width <= 512) score_taps_kernel<2><<<(int)blocks, threads, smem, stream>>>((const __half*)feat_f16, ld, w9, width, rows, taps);
  else score_taps_kernel<4><<<(int)blocks, threads, smem, stream>>>((const __half*)feat_f16, ld, w9, width, rows, taps);
  rc = check_launch("score_taps_kernel");
  if (rc) return rc;
  score_pred_kernel<<<(int)((rows + 255) / 256), 256, 0, stream>>>(taps, 9, bias, nullptr, batch, h, w, pred);
  rc = check_launch("score_pred_kernel");
  if (rc || !up) return rc;
  const long long tot = (long long)batch * out_h * (out_w / 4);
  long long b2 = (tot + 255) / 256;
  if (b2 > (long long)num_sms() * 16) b2 = (long long)num_sms() * 16;
  upsample_sigmoid_kernel<<<(int)b2, 256, 0, stream>>>(pred, batch, h, w, out_h, out_w, up, sigm);
  return check_launch("upsample_sigmoid_kernel");
}

// Same head, but the nine per-pixel tap dot products come from a tensor-core GEMM (cmpc_gemm_f16 with N = 9 padded to 32):
// taps fp32 [B*h*w, ld_taps], tap k = 3*dy + dx in column k.
extern "C" int cmpc_score_from_taps(const float* taps, int64_t ld_taps, float bias, const float* bias_dev, int32_t batch, int32_t h, int32_t w,
                                    int32_t out_h, int32_t out_w, float* pred, float* up, float* sigm, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(taps && pred && batch > 0 && h > 0 && w > 0 && ld_taps >= 9, CMPC_ERR_ARG, "cmpc_score_from_taps: bad args");
  CMPC_REQUIRE(up == nullptr || (out_w % 4 == 0 && out_h > 0), CMPC_ERR_ARG, "cmpc_score_from_taps: W must be a multiple of 4");
  const long long rows = (long long)batch * h * w;
  score_pred_kernel<<<(int)((rows + 255) / 256), 256, 0, stream>>>(taps, (int)ld_taps, bias, bias_dev, batch, h, w, pred);
  rc = check_launch("score_pred_kernel");
  if (rc || !up) return rc;
  const long long tot = (long long)batch * out_h * (out_w / 4);
  long long b2 = (tot + 255) / 256;
  if (b2 > (long long)num_sms() * 16) b2 = (long long)num_sms() * 16;
  upsample_sigmoid_kernel<<<(int)b2, 256, 0, stream>>>(pred, batch, h, w, out_h, out_w, up, sigm);
  return check_launch("upsample_sigmoid_kernel");
}

extern "C" int cmpc_iou_counts(const float* up, const float* target, int32_t batch, int64_t per_sample, float thresh,
                               int32_t inclusive, uint64_t* iu, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(up && target && iu && batch > 0 && per_sample > 0, CMPC_ERR_ARG, "cmpc_iou_counts: bad args");
  int bx = (int)((per_sample + 256 * 8 - 1) / (256 * 8));
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  iou_counts_kernel<<<dim3(bx, batch), 256, 0, (cudaStream_t)stream>>>(up, target, per_sample, thresh, inclusive,
                                                                        (unsigned long long*)iu);
  return check_launch("iou_counts_kernel");
}

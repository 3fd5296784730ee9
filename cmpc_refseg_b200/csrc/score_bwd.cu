// Backward of the loss -> upsample -> 3x3 score conv tail (CMPC_model.py:138-142 and :128-133 for the aux heads, loss at
// :439-445 / util/loss.py:6-16): the seed of the whole backward pass.
//   loss_k = coef_k * mean_b sum_pixels sigmoid_ce(up, target)   =>  d up = coef_k / B * (sigmoid(up) - target)
//   up = legacy bilinear resize of pred (no half-pixel centres)    =>  d pred = resize^T d up   (gathered per pred pixel)
//   pred = conv3x3(F) + b                                          =>  the nine shifted copies of d pred ("D9", fp16 [rows, ld])
// feed two GEMMs: dF = D9 . w9 (cmpc_gemm_f16, K = 16) and dw9 = D9^T F (cmpc_gemm_atb_f16); db = sum d pred.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

__global__ void score_bwd_dpred_kernel(const float* __restrict__ up, const float* __restrict__ target, float scale, int B, int h, int w,
                                       int H, int W, float* __restrict__ dpred, float* __restrict__ dbias) {
  const long long total = (long long)B * h * w;
  const float ys = (float)h / (float)H, xs = (float)w / (float)W;
  float local = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const long long b = i / ((long long)w * h);
    // output rows / columns whose two source taps can include y / x (same arithmetic as upsample_sigmoid_kernel)
    const int Y0 = max(0, (int)floorf((float)(y - 1) / ys) - 1), Y1 = min(H - 1, (int)ceilf((float)(y + 1) / ys) + 1);
    const int X0 = max(0, (int)floorf((float)(x - 1) / xs) - 1), X1 = min(W - 1, (int)ceilf((float)(x + 1) / xs) + 1);
    float acc = 0.f;
    for (int Y = Y0; Y <= Y1; ++Y) {
      const float in_y = (float)Y * ys;
      const int y0 = (int)floorf(in_y), y1 = min(y0 + 1, h - 1);
      const float ly = in_y - (float)y0;
      const float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
      if (wy == 0.f) continue;
      const float* ur = up + (b * H + Y) * W;
      const float* tr = target + (b * H + Y) * W;
      float racc = 0.f;
      for (int X = X0; X <= X1; ++X) {
        const float in_x = (float)X * xs;
        const int x0 = (int)floorf(in_x), x1 = min(x0 + 1, w - 1);
        const float lx = in_x - (float)x0;
        const float wx = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
        if (wx == 0.f) continue;
        const float u = __ldg(ur + X);
        racc += wx * (1.0f / (1.0f + expf(-u)) - __ldg(tr + X));
      }
      acc += wy * racc;
    }
    acc *= scale;
    dpred[i] = acc;
    local += acc;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0 && dbias) atomicAdd(dbias, local);
}

// D9[m, 3*dy + dx] = dpred at the output pixel that read F[m] through tap (dy, dx); columns 9..ld-1 zero
__global__ void score_bwd_taps_kernel(const float* __restrict__ dpred, int B, int h, int w, __half* __restrict__ d9, int ld) {
  const long long total = (long long)B * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const long long b = i / ((long long)w * h);
    __half* o = d9 + i * ld;
    for (int k = 0; k < ld; ++k) {
      float v = 0.f;
      if (k < 9) {
        const int dy = k / 3, dx = k - dy * 3;
        const int yy = y - dy + 1, xx = x - dx + 1;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldg(dpred + (b * h + yy) * w + xx);
      }
      o[k] = __float2half_rn(v);
    }
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_score_bwd_dpred(const float* up, const float* target, float scale, int32_t batch, int32_t h, int32_t w, int32_t out_h,
                                    int32_t out_w, float* dpred, float* dbias, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(up && target && dpred && batch > 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0, CMPC_ERR_ARG, "cmpc_score_bwd_dpred: bad args");
  const long long total = (long long)batch * h * w;
  long long blocks = (total + 127) / 128;
  score_bwd_dpred_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(up, target, scale, batch, h, w, out_h, out_w, dpred, dbias);
  return check_launch("score_bwd_dpred_kernel");
}

extern "C" int cmpc_score_bwd_taps(const float* dpred, int32_t batch, int32_t h, int32_t w, void* d9_f16, int32_t ld, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(dpred && d9_f16 && batch > 0 && h > 0 && w > 0 && ld >= 16 && ld % 8 == 0, CMPC_ERR_ARG, "cmpc_score_bwd_taps: bad args (ld >= 16, %% 8)");
  const long long total = (long long)batch * h * w;
  score_bwd_taps_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dpred, batch, h, w, (__half*)d9_f16, ld);
  return check_launch("score_bwd_taps_kernel");
}

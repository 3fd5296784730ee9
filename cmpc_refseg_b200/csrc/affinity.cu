// Relation-aware reasoning, phase A: the two softmaxes of the node-word affinity (CMPC_model.py:388-399).
//   affi[b, n, t]  (already scaled by R_t / sqrt(C) in the affinity GEMM epilogue; columns t >= T are zero)
//   W[b, n, t] = softmax_t(mask * affi + (1 - mask) * FLT_MIN)        -> fp16 [B*N, 32]   (gw_w, :392-395)
//   V[b, n, t] = mask * softmax_n(affi)                               -> fp16 [B*N, 32]   (gw_v, :397-399)
// V is stored multiplied by v_scale (a power of two ~ N) so that P = W V^T is O(1) in fp16; the graph kernel
// divides it out again.  Padded word columns (t >= T) carry W = V = 0, so they add nothing to W V^T.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

constexpr int AFF_T = 32;           // padded word count (K of the first MMA of the graph kernel)
constexpr int CS_THREADS = 1024;    // 32 row-groups x 32 words

// per (sample, split): running column max and sum(exp) over its rows
__global__ void __launch_bounds__(CS_THREADS)
affinity_colstats_kernel(const float* __restrict__ affi, int rows_per_sample, int nsplit, float* __restrict__ part /*[B,nsplit,2,32]*/) {
  const int b = blockIdx.x, split = blockIdx.y;
  const int t = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int per = (rows_per_sample + nsplit - 1) / nsplit;
  const int r0 = split * per, r1 = min(rows_per_sample, r0 + per);
  const float* a = affi + (long long)b * rows_per_sample * AFF_T;
  float mx = -INFINITY, l = 0.f;
  for (int r = r0 + rg; r < r1; r += 32) {
    const float v = __ldg(a + (long long)r * AFF_T + t);
    const float mn = fmaxf(mx, v);
    l = l * __expf(mx - mn) + __expf(v - mn);
    mx = mn;
  }
  __shared__ float s_m[32][33], s_l[32][33];
  s_m[rg][t] = mx;
  s_l[rg][t] = l;
  __syncthreads();
  if (rg == 0) {
    float gm = -INFINITY;
    for (int g = 0; g < 32; ++g) gm = fmaxf(gm, s_m[g][t]);
    float gl = 0.f;
    for (int g = 0; g < 32; ++g) gl += (s_m[g][t] == -INFINITY) ? 0.f : s_l[g][t] * __expf(s_m[g][t] - gm);
    float* o = part + ((long long)(b * nsplit + split) * 2) * 32;
    o[t] = gm;
    o[32 + t] = gl;
  }
}

// warp per node row
__global__ void affinity_wv_kernel(const float* __restrict__ affi, const float* __restrict__ mask /*[B,T]*/,
                                   const float* __restrict__ part, int nsplit, int T, int rows_per_sample, long long rows,
                                   float v_scale, __half* __restrict__ w16, __half* __restrict__ v16,
                                   float* __restrict__ gw_w /*[rows,T] or null*/, float* __restrict__ gw_v,
                                   const float* __restrict__ x_row_ss /*[rows] or null*/) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    const int b = (int)(r / rows_per_sample);
    // merge the column statistics of the splits (tiny, L1-resident)
    float cm = -INFINITY;
    for (int s = 0; s < nsplit; ++s) cm = fmaxf(cm, __ldg(part + ((long long)(b * nsplit + s) * 2) * 32 + lane));
    float cl = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float m = __ldg(part + ((long long)(b * nsplit + s) * 2) * 32 + lane);
      cl += (m == -INFINITY) ? 0.f : __ldg(part + ((long long)(b * nsplit + s) * 2 + 1) * 32 + lane) * __expf(m - cm);
    }
    const float a = __ldg(affi + r * AFF_T + lane);
    const float mk = (lane < T) ? __ldg(mask + b * T + lane) : 0.f;
    // softmax over words: masked logits get tf.float32.min, padded columns are excluded outright
    const float logit = (lane < T) ? (mk * a + (1.0f - mk) * -3.4028234663852886e38f) : -INFINITY;
    const float rm = warp_max(logit);
    const float e = (lane < T) ? __expf(logit - rm) : 0.f;
    const float rsum = warp_sum(e);
    const float wv = e / rsum;
    const float vv = (lane < T) ? mk * __expf(a - cm) / cl : 0.f;
    // x_row_ss: the node features X the graph kernel will read are NOT l2-normalised yet (deferred l2_normalize of the MUTAN map,
    // CMPC_model.py:324): adj X = (W V^T) diag(1 / |x_j|) X_raw, so the per-node factor rides on V (clamped to the fp16 range)
    const float xs = x_row_ss ? rsqrtf(fmaxf(__ldg(x_row_ss + r), 1e-12f)) : 1.0f;
    w16[r * AFF_T + lane] = __float2half_rn(wv);
    v16[r * AFF_T + lane] = __float2half_rn(fminf(vv * v_scale * xs, 65504.f));
    if (gw_w && lane < T) {
      gw_w[r * T + lane] = wv;
      gw_v[r * T + lane] = vv;
    }
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" size_t cmpc_affinity_workspace_bytes(int32_t batch) { return (size_t)batch * 8 * 2 * 32 * sizeof(float); }

extern "C" int cmpc_affinity_softmax_scaled(const float* affi, const float* seq_mask, int32_t batch, int32_t rows_per_sample,
                                            int32_t t, float v_scale, void* w_f16, void* v_f16, float* gw_w, float* gw_v,
                                            const float* x_row_sumsq, void* workspace, size_t workspace_bytes, void* stream);

extern "C" int cmpc_affinity_softmax(const float* affi, const float* seq_mask, int32_t batch, int32_t rows_per_sample,
                                     int32_t t, float v_scale, void* w_f16, void* v_f16, float* gw_w, float* gw_v,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  return cmpc_affinity_softmax_scaled(affi, seq_mask, batch, rows_per_sample, t, v_scale, w_f16, v_f16, gw_w, gw_v, nullptr, workspace,
                                      workspace_bytes, stream);
}

extern "C" int cmpc_affinity_softmax_scaled(const float* affi, const float* seq_mask, int32_t batch, int32_t rows_per_sample,
                                            int32_t t, float v_scale, void* w_f16, void* v_f16, float* gw_w, float* gw_v,
                                            const float* x_row_sumsq, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(affi && seq_mask && w_f16 && v_f16 && workspace && batch > 0 && rows_per_sample > 0, CMPC_ERR_ARG,
               "cmpc_affinity_softmax: bad args");
  CMPC_REQUIRE(t >= 1 && t <= AFF_T, CMPC_ERR_ARG, "cmpc_affinity_softmax: T must be in [1, 32]");
  CMPC_REQUIRE((gw_w == nullptr) == (gw_v == nullptr), CMPC_ERR_ARG, "cmpc_affinity_softmax: gw_w / gw_v must both be set or null");
  CMPC_REQUIRE(workspace_bytes >= cmpc_affinity_workspace_bytes(batch), CMPC_ERR_WORKSPACE, "cmpc_affinity_softmax: workspace too small");
  const int nsplit = 8;
  affinity_colstats_kernel<<<dim3(batch, nsplit), CS_THREADS, 0, (cudaStream_t)stream>>>(affi, rows_per_sample, nsplit, (float*)workspace);
  rc = check_launch("affinity_colstats_kernel");
  if (rc) return rc;
  const long long rows = (long long)batch * rows_per_sample;
  long long blocks = (rows * 32 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  affinity_wv_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(affi, seq_mask, (const float*)workspace, nsplit, t, rows_per_sample,
                                                                     rows, v_scale, (__half*)w_f16, (__half*)v_f16, gw_w, gw_v, x_row_sumsq);
  return check_launch("affinity_wv_kernel");
}

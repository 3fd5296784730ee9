// Backward of the language side (CMPC_model.py:159-192, :347-357): word-type attention (softmax over the four word types,
// masked), the two weighted sentence vectors valid_lang / nec_lang with their l2_normalize, the relation weights R_t that
// gate the affinity, and the l2_normalize of the LSTM outputs.  A handful of rows: one block per sentence.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

constexpr int LB2_THREADS = 256;
constexpr int LB2_MAXC = 8;      // columns per thread: R <= 8 * 256

// dwords[b, t, :] = wv_t dvr + wn_t dnr  (WRITTEN);  dlogit[b, t, 0..3] (written)
__global__ void __launch_bounds__(LB2_THREADS)
lang_bwd_kernel(const float* __restrict__ words, const float* __restrict__ parse /*[B,T,4]*/, const float* __restrict__ mask,
                const float* __restrict__ valid, const float* __restrict__ nec, const float* __restrict__ d_valid,
                const float* __restrict__ d_nec, const float* __restrict__ drgate /*[B,32]*/, float inv_sqrt_c, int T, int R,
                float* __restrict__ dwords, float* __restrict__ dlogit /*[B,T,4]*/) {
  extern __shared__ float sm[];
  float* s_dv = sm;            // [R]  d(valid_raw)
  float* s_dn = s_dv + R;      // [R]  d(nec_raw)
  float* s_wv = s_dn + R;      // [T]
  float* s_wn = s_wv + T;      // [T]
  float* s_dwv = s_wn + T;     // [T]
  float* s_dwn = s_dwv + T;    // [T]
  float* s_red = s_dwn + T;    // [4 * warps]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = LB2_THREADS / 32;
  if (tid < T) {
    const float* p = parse + ((long long)b * T + tid) * 4;
    s_wv[tid] = p[0] + p[1];
    s_wn[tid] = (p[0] + p[1] + p[2] + p[3]) - p[3];
  }
  __syncthreads();
  // raw sentence vectors (for their norms) and the dots valid . d_valid, nec . d_nec
  float vraw[LB2_MAXC], nraw[LB2_MAXC];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int nc = 0;
  for (int c = tid; c < R; c += LB2_THREADS, ++nc) {
    float av = 0.f, an = 0.f;
    for (int t = 0; t < T; ++t) {
      const float w = __ldg(words + ((long long)b * T + t) * R + c);
      av += s_wv[t] * w;
      an += s_wn[t] * w;
    }
    vraw[nc] = av; nraw[nc] = an;
    a0 += av * av; a1 += an * an;
    a2 += __ldg(valid + (long long)b * R + c) * __ldg(d_valid + (long long)b * R + c);
    a3 += __ldg(nec + (long long)b * R + c) * __ldg(d_nec + (long long)b * R + c);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  if (lane == 0) { s_red[warp] = a0; s_red[NW + warp] = a1; s_red[2 * NW + warp] = a2; s_red[3 * NW + warp] = a3; }
  __syncthreads();
  float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
  for (int w = 0; w < NW; ++w) { t0 += s_red[w]; t1 += s_red[NW + w]; t2 += s_red[2 * NW + w]; t3 += s_red[3 * NW + w]; }
  const float iv = rsqrtf(fmaxf(t0, 1e-12f)), in = rsqrtf(fmaxf(t1, 1e-12f));
  nc = 0;
  for (int c = tid; c < R; c += LB2_THREADS, ++nc) {
    s_dv[c] = (__ldg(d_valid + (long long)b * R + c) - __ldg(valid + (long long)b * R + c) * t2) * iv;
    s_dn[c] = (__ldg(d_nec + (long long)b * R + c) - __ldg(nec + (long long)b * R + c) * t3) * in;
  }
  __syncthreads();
  // d wv_t = words_t . dv,  d wn_t = words_t . dn  (warp per word);  dwords_t = wv_t dv + wn_t dn
  for (int t = warp; t < T; t += NW) {
    float dv = 0.f, dn = 0.f;
    const float wv = s_wv[t], wn = s_wn[t];
    for (int c = lane; c < R; c += 32) {
      const float w = __ldg(words + ((long long)b * T + t) * R + c);
      dv += w * s_dv[c];
      dn += w * s_dn[c];
      dwords[((long long)b * T + t) * R + c] = wv * s_dv[c] + wn * s_dn[c];
    }
    dv = warp_sum(dv); dn = warp_sum(dn);
    if (lane == 0) { s_dwv[t] = dv; s_dwn[t] = dn; }
  }
  __syncthreads();
  if (tid < T) {
    const float* p = parse + ((long long)b * T + tid) * 4;
    const float mk = __ldg(mask + b * T + tid);
    // p = softmax * mask;  d p_j: E, A feed both sentence vectors, R feeds nec_lang and the relation gate, U nothing
    const float dp0 = s_dwv[tid] + s_dwn[tid], dp1 = dp0, dp2 = s_dwn[tid] + drgate[b * 32 + tid] * inv_sqrt_c, dp3 = 0.f;
    const float dot = p[0] * dp0 + p[1] * dp1 + p[2] * dp2 + p[3] * dp3;      // = mask * sum_k s_k dp_k
    float* o = dlogit + ((long long)b * T + tid) * 4;
    o[0] = p[0] * (dp0 * mk - dot);      // mask * s_j * (dp_j - sum_k s_k dp_k) with p_j = mask * s_j, mask in {0, 1}
    o[1] = p[1] * (dp1 * mk - dot);
    o[2] = p[2] * (dp2 * mk - dot);
    o[3] = p[3] * (dp3 * mk - dot);
  }
}

// d x = (d y - y (y . d y)) / |x|   for y = l2_normalize(x) row-wise (fp32), warp per row
__global__ void l2norm_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x, int rows, int R,
                                      float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float dot = 0.f, ss = 0.f;
  for (int c = lane; c < R; c += 32) {
    dot += dy[(long long)row * R + c] * y[(long long)row * R + c];
    const float v = x[(long long)row * R + c];
    ss += v * v;
  }
  dot = warp_sum(dot); ss = warp_sum(ss);
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
  for (int c = lane; c < R; c += 32) dx[(long long)row * R + c] = (dy[(long long)row * R + c] - y[(long long)row * R + c] * dot) * inv;
}

// out = dy * [y > 0]   (relu), small fp32 matrices with a row stride
__global__ void relu_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ out, int rows, int cols,
                                    long long ld) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    out[r * ld + c] = y[r * ld + c] > 0.f ? dy[r * ld + c] : 0.f;
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_lang_bwd(const float* words_f32, const float* parse, const float* seq_mask, const float* valid_f32, const float* nec_f32,
                             const float* d_valid, const float* d_nec, const float* drgate, int32_t batch, int32_t t, int32_t r, int32_t c,
                             float* dwords, float* dlogit, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(words_f32 && parse && seq_mask && valid_f32 && nec_f32 && d_valid && d_nec && drgate && dwords && dlogit, CMPC_ERR_ARG,
               "cmpc_lang_bwd: null pointer");
  CMPC_REQUIRE(batch > 0 && t > 0 && t <= 32 && r > 0 && r <= LB2_MAXC * LB2_THREADS && c > 0, CMPC_ERR_ARG, "cmpc_lang_bwd: need T <= 32, R <= 2048");
  const size_t smem = (size_t)(2 * r + 4 * t + 4 * (LB2_THREADS / 32)) * sizeof(float);
  lang_bwd_kernel<<<batch, LB2_THREADS, smem, (cudaStream_t)stream>>>(words_f32, parse, seq_mask, valid_f32, nec_f32, d_valid, d_nec, drgate,
                                                                      1.0f / sqrtf((float)c), t, r, dwords, dlogit);
  return check_launch("lang_bwd_kernel");
}

extern "C" int cmpc_l2norm_bwd_f32(const float* dy, const float* y, const float* x, int32_t rows, int32_t r, float* dx, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(dy && y && x && dx && rows > 0 && r > 0, CMPC_ERR_ARG, "cmpc_l2norm_bwd_f32: bad args");
  l2norm_bwd_f32_kernel<<<(rows * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dy, y, x, rows, r, dx);
  return check_launch("l2norm_bwd_f32_kernel");
}

extern "C" int cmpc_relu_bwd_f32(const float* dy, const float* y, float* out, int32_t rows, int32_t cols, int64_t ld, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(dy && y && out && rows > 0 && cols > 0 && ld >= cols, CMPC_ERR_ARG, "cmpc_relu_bwd_f32: bad args");
  long long blocks = ((long long)rows * cols + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  relu_bwd_f32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dy, y, out, rows, cols, ld);
  return check_launch("relu_bwd_f32_kernel");
}

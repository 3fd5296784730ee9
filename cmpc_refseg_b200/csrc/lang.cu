// Language-side kernels of the CMPC head (tiny: B*T = 640 word rows, B sentence rows).
//   words_prepare   l2-normalised word features + seq_mask           CMPC_model.py:159-163
//   lang_parse      word-type attention (E,A,R,U) + valid_lang / nec_lang sentence vectors  :347-357, :166-192
//   small_linear    fp32 batched skinny matmul (rows <= 64) for per-sentence projections (:223 collapsed key/query)
//   gv_gates        global_vec tail + lang_se gates: gv = l2norm(conv([g | lang])), sigmoid(conv(gv))  :238-241, :202-204
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

// warp per word row
__global__ void words_prepare_kernel(const float* __restrict__ lstm, int rows, int R, float* __restrict__ wf32,
                                     __half* __restrict__ wf16, long long ld16, float* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* x = lstm + (long long)row * R;
  float ss = 0.f;
  for (int c = lane; c < R; c += 32) { const float v = __ldg(x + c); ss += v * v; }
  ss = warp_sum(ss);
  const float sc = rsqrtf(fmaxf(ss, 1e-12f));
  float sa = 0.f;
  for (int c = lane; c < R; c += 32) {
    const float v = __ldg(x + c) * sc;
    sa += fabsf(v);
    wf32[(long long)row * R + c] = v;
    wf16[(long long)row * ld16 + c] = __float2half_rn(v);
  }
  for (int c = R + lane; c < ld16; c += 32) wf16[(long long)row * ld16 + c] = __float2half_rn(0.f);
  sa = warp_sum(sa);
  if (lane == 0) mask[row] = (sa != 0.f) ? 1.f : 0.f;
}

// one CTA per sentence
constexpr int PARSE_THREADS = 256;
__global__ void __launch_bounds__(PARSE_THREADS)
lang_parse_kernel(const float* __restrict__ hidden, long long ldh, int HID, const float* __restrict__ w2 /*[HID,4]*/,
                  const float* __restrict__ b2, const float* __restrict__ wf32, const float* __restrict__ mask, int T, int R,
                  float inv_sqrt_c, float* __restrict__ parse /*[B,T,4]*/, float* __restrict__ rgate /*[B,32]*/,
                  float* __restrict__ valid32, float* __restrict__ nec32 /*[B,R]*/, __half* __restrict__ valid16,
                  __half* __restrict__ nec16, long long ld16) {
  extern __shared__ float sm[];
  float* s_logit = sm;            // [T*4]
  float* s_wv = s_logit + T * 4;  // [T] E+A
  float* s_wn = s_wv + T;         // [T] E+A+R
  float* s_red = s_wn + T;        // [2 * warps]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = PARSE_THREADS / 32;
  // logits: one warp per word, all four word types at once (the hidden row is read once, the [HID, 4] weight as one float4 per
  // k; the k-loop is unrolled so every load of a word is in flight together); skipped when the caller supplies `parse`
  for (int t = warp; hidden != nullptr && t < T; t += NW) {
    const float* h = hidden + (long long)(b * T + t) * ldh;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int k = lane; k < HID; k += 32) {
      const float x = __ldg(h + k);
      const float4 w = __ldg(reinterpret_cast<const float4*>(w2) + k);
      a0 = fmaf(x, w.x, a0); a1 = fmaf(x, w.y, a1); a2 = fmaf(x, w.z, a2); a3 = fmaf(x, w.w, a3);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    if (lane == 0) {
      s_logit[t * 4 + 0] = a0 + __ldg(b2 + 0); s_logit[t * 4 + 1] = a1 + __ldg(b2 + 1);
      s_logit[t * 4 + 2] = a2 + __ldg(b2 + 2); s_logit[t * 4 + 3] = a3 + __ldg(b2 + 3);
    }
  }
  __syncthreads();
  if (tid < 32) {
    float r = 0.f;
    if (tid < T) {
      float* o = parse + ((long long)b * T + tid) * 4;
      float p0, p1, p2, p3;
      if (hidden != nullptr) {
        const float l0 = s_logit[tid * 4], l1 = s_logit[tid * 4 + 1], l2 = s_logit[tid * 4 + 2], l3 = s_logit[tid * 4 + 3];
        const float mx = fmaxf(fmaxf(l0, l1), fmaxf(l2, l3));
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx), e3 = expf(l3 - mx);
        const float inv = 1.0f / (e0 + e1 + e2 + e3);
        const float mk = __ldg(mask + b * T + tid);
        p0 = e0 * inv * mk; p1 = e1 * inv * mk; p2 = e2 * inv * mk; p3 = e3 * inv * mk;
        o[0] = p0; o[1] = p1; o[2] = p2; o[3] = p3;
      } else {
        p0 = o[0]; p1 = o[1]; p2 = o[2]; p3 = o[3];
      }
      s_wv[tid] = p0 + p1;
      s_wn[tid] = (p0 + p1 + p2 + p3) - p3;   // words_parse_sum - U   (CMPC_model.py:182-183)
      r = p2 * inv_sqrt_c;
    }
    rgate[b * 32 + tid] = r;
  }
  __syncthreads();
  // weighted sentence vectors; each thread owns columns c = tid, tid + 256, ...
  float sv = 0.f, sn = 0.f;
  float accv[8], accn[8];   // R <= 8 * 256
  int nc = 0;
  for (int c = tid; c < R; c += PARSE_THREADS, ++nc) {
    float av = 0.f, an = 0.f;
#pragma unroll 10
    for (int t = 0; t < T; ++t) {
      const float w = __ldg(wf32 + ((long long)(b * T + t)) * R + c);
      av += s_wv[t] * w;
      an += s_wn[t] * w;
    }
    accv[nc] = av; accn[nc] = an;
    sv += av * av; sn += an * an;
  }
  sv = warp_sum(sv); sn = warp_sum(sn);
  if (lane == 0) { s_red[warp] = sv; s_red[NW + warp] = sn; }
  __syncthreads();
  float tv = 0.f, tn = 0.f;
  for (int w = 0; w < NW; ++w) { tv += s_red[w]; tn += s_red[NW + w]; }
  const float scv = rsqrtf(fmaxf(tv, 1e-12f)), scn = rsqrtf(fmaxf(tn, 1e-12f));
  nc = 0;
  for (int c = tid; c < R; c += PARSE_THREADS, ++nc) {
    const float v = accv[nc] * scv, n = accn[nc] * scn;
    valid32[(long long)b * R + c] = v;
    nec32[(long long)b * R + c] = n;
    valid16[(long long)b * ld16 + c] = __float2half_rn(v);
    nec16[(long long)b * ld16 + c] = __float2half_rn(n);
  }
  for (int c = R + tid; c < ld16; c += PARSE_THREADS) {
    valid16[(long long)b * ld16 + c] = __float2half_rn(0.f);
    nec16[(long long)b * ld16 + c] = __float2half_rn(0.f);
  }
}

// out[z, r, n] = act(sum_k x[z, r, k] * W[z, k, n] + bias[z, n]);  thread per n, RB rows in registers.
constexpr int SL_THREADS = 128;
constexpr int SL_RB = 8;
constexpr int SL_KC = 64;
__global__ void __launch_bounds__(SL_THREADS)
small_linear_kernel(const float* __restrict__ x, long long ldx, long long x_zstride, const float* __restrict__ w,
                    long long ldw, long long w_zstride, const float* __restrict__ bias, long long b_zstride,
                    float* __restrict__ out, long long ldo, long long o_zstride, int rows, int K, int N, int act, int ksplit) {
  // accumulate mode (act == 4) may split K over ksplit blocks (folded into blockIdx.y) and add its partial with atomics: the
  // backward's [B, 5000] x [5000, 1000] products would otherwise run on 16 blocks
  const int z = blockIdx.z;
  const int n = blockIdx.x * SL_THREADS + threadIdx.x;
  const int ks = blockIdx.y % ksplit;
  const int r0 = (blockIdx.y / ksplit) * SL_RB;
  const int kper = ((K + ksplit - 1) / ksplit + SL_KC - 1) / SL_KC * SL_KC;
  const int kbeg = ks * kper, kend_all = min(K, kbeg + kper);
  x += z * x_zstride; w += z * w_zstride; out += z * o_zstride;
  __shared__ float sx[SL_RB][SL_KC];
  float acc[SL_RB];
#pragma unroll
  for (int r = 0; r < SL_RB; ++r) acc[r] = 0.f;
  for (int k0 = kbeg; k0 < kend_all; k0 += SL_KC) {
    for (int i = threadIdx.x; i < SL_RB * SL_KC; i += SL_THREADS) {
      const int r = i / SL_KC, k = i - r * SL_KC;
      sx[r][k] = (r0 + r < rows && k0 + k < kend_all) ? __ldg(x + (long long)(r0 + r) * ldx + k0 + k) : 0.f;
    }
    __syncthreads();
    if (n < N) {
      const int kend = min(SL_KC, kend_all - k0);
#pragma unroll 8
      for (int k = 0; k < kend; ++k) {
        const float wv = __ldg(w + (long long)(k0 + k) * ldw + n);
#pragma unroll
        for (int r = 0; r < SL_RB; ++r) acc[r] += sx[r][k] * wv;
      }
    }
    __syncthreads();
  }
  if (n < N) {
    const float bv = (bias && ks == 0) ? __ldg(bias + z * b_zstride + n) : 0.f;
#pragma unroll
    for (int r = 0; r < SL_RB; ++r) {
      if (r0 + r < rows) {
        float v = acc[r] + bv;
        if (ksplit > 1) { atomicAdd(out + (long long)(r0 + r) * ldo + n, v); continue; }
        if (act == 1) v = fmaxf(v, 0.f);
        else if (act == 2) v = tanh_acc(v);
        else if (act == 3) v = sigmoid_acc(v);
        else if (act == 4) v += out[(long long)(r0 + r) * ldo + n];      // accumulate (backward: several paths into one gradient)
        out[(long long)(r0 + r) * ldo + n] = v;
      }
    }
  }
}

// one CTA per (sample, module): gv = l2norm(g @ Wg + gvl); gate_f = sigmoid(gv @ Wf + bf), f = 1, 2.
// Thread (kh, n): output column n, half kh of the reduction; the k-loop is unrolled 10x so ten independent weight loads
// are in flight per thread (the serial version was pure L2 latency: 500 dependent-issue loads per stage).
constexpr int GV_NT = 512;                 // output columns per block (mlp_dim <= 512)
constexpr int GV_THREADS = 2 * GV_NT;
__global__ void __launch_bounds__(GV_THREADS)
gv_gates_kernel(const float* __restrict__ g, long long ldg, const float* __restrict__ gvl, long long ldgvl, long long gvl_bstride,
                const float* __restrict__ wg, const float* __restrict__ wf1, const float* __restrict__ bf1,
                const float* __restrict__ wf2, const float* __restrict__ bf2, long long w_mstride, long long b_mstride,
                int nmod, int Mdim, float* __restrict__ gv_out, float* __restrict__ gate1, float* __restrict__ gate2,
                long long ldgate, int mode, float* __restrict__ batch_ss, long long g1_bstride, long long g1_mstride,
                long long g2_bstride, long long g2_mstride) {
  // gate_f of (sample b, module mod) is written at gate_f + b * gf_bstride + mod * gf_mstride (signed: the head lays the gates of an
  // exchange round out per SOURCE map, see head._st_exchange_round); gv_out stays [B, nmod, ldgate].
  // mode 0: l2_normalize per sample.  Batch-coupled l2_normalize (tf.nn.l2_normalize without an axis, CMPC_model.py:241) in two
  // launches: mode 1 writes the un-normalised z to gv_out and adds |z|^2 to batch_ss[mod]; mode 2 reads both back and finishes.
  const int b = blockIdx.x, mod = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = tid % GV_NT, kh = tid / GV_NT;
  constexpr int NW = GV_THREADS / 32;
  __shared__ float s_in[GV_NT], s_gv[GV_NT], s_part[3][GV_NT], s_red[NW];
  const long long bm = (long long)b * nmod + mod;
  for (int k = tid; k < Mdim; k += GV_THREADS) s_in[k] = __ldg(g + bm * ldg + k);
  __syncthreads();
  const int k0 = kh * ((Mdim + 1) / 2), k1 = min(Mdim, k0 + (Mdim + 1) / 2);
  float v = 0.f;
  if (mode != 2 && n < Mdim) {
    const float* W = wg + mod * w_mstride + n;
#pragma unroll 10
    for (int k = k0; k < k1; ++k) v = fmaf(s_in[k], __ldg(W + (long long)k * Mdim), v);
  }
  if (kh == 1) s_part[0][n] = v;
  __syncthreads();
  float ss = 0.f;
  if (kh == 0 && n < Mdim) {
    if (mode == 2) v = gv_out[bm * ldgate + n];
    else v += s_part[0][n] + __ldg(gvl + (long long)b * gvl_bstride + (long long)mod * ldgvl + n);
    ss = v * v;
  } else {
    v = 0.f;
  }
  ss = warp_sum(ss);
  if (lane == 0) s_red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < NW; ++w) tot += s_red[w];
  if (mode == 1) {
    if (kh == 0 && n < ldgate) gv_out[bm * ldgate + n] = n < Mdim ? v : 0.f;
    if (tid == 0) atomicAdd(batch_ss + mod, tot);
    return;
  }
  if (mode == 2) tot = batch_ss[mod];
  const float gvn = v * rsqrtf(fmaxf(tot, 1e-12f));
  if (kh == 0) {
    if (n < Mdim) { s_gv[n] = gvn; gv_out[bm * ldgate + n] = gvn; }
    else if (n < ldgate) gv_out[bm * ldgate + n] = 0.f;
  }
  __syncthreads();
  float a1 = 0.f, a2 = 0.f;
  if (n < Mdim) {
    const float* W1 = wf1 + mod * w_mstride + n;
    const float* W2 = wf2 + mod * w_mstride + n;
#pragma unroll 5
    for (int k = k0; k < k1; ++k) {
      const float x = s_gv[k];
      a1 = fmaf(x, __ldg(W1 + (long long)k * Mdim), a1);
      a2 = fmaf(x, __ldg(W2 + (long long)k * Mdim), a2);
    }
  }
  if (kh == 1) { s_part[1][n] = a1; s_part[2][n] = a2; }
  __syncthreads();
  if (kh == 0) {
    float* o1 = gate1 + b * g1_bstride + mod * g1_mstride;
    float* o2 = gate2 + b * g2_bstride + mod * g2_mstride;
    if (n < Mdim) {
      o1[n] = sigmoid_acc(a1 + s_part[1][n] + __ldg(bf1 + mod * b_mstride + n));
      o2[n] = sigmoid_acc(a2 + s_part[2][n] + __ldg(bf2 + mod * b_mstride + n));
    } else if (n < ldgate) {
      o1[n] = 0.f;
      o2[n] = 0.f;
    }
  }
}

// Same computation with 128-bit weight loads: thread (ks, c4) owns four adjacent output columns and one eighth of the reduction, so
// a stage is 63 dependent-issue steps of LDG.128 instead of 250 of LDG.32 (the kernel is a chain of L2 latencies: more samples per
// CTA sharing the loads made it slower, shorter chains make it faster).  Needs Mdim % 4 == 0 (16-byte aligned weight rows).
constexpr int GV4_KS = 8;
__global__ void __launch_bounds__(GV_THREADS)
gv_gates_v4_kernel(const float* __restrict__ g, long long ldg, const float* __restrict__ gvl, long long ldgvl, long long gvl_bstride,
                   const float* __restrict__ wg, const float* __restrict__ wf1, const float* __restrict__ bf1,
                   const float* __restrict__ wf2, const float* __restrict__ bf2, long long w_mstride, long long b_mstride,
                   int nmod, int Mdim, float* __restrict__ gv_out, float* __restrict__ gate1, float* __restrict__ gate2,
                   long long ldgate, int mode, float* __restrict__ batch_ss, long long g1_bstride, long long g1_mstride,
                   long long g2_bstride, long long g2_mstride) {
  const int b = blockIdx.x, mod = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c4 = tid % (GV_NT / 4), ks = tid / (GV_NT / 4), col = c4 * 4;
  constexpr int NW = GV_THREADS / 32;
  __shared__ float s_in[GV_NT], s_gv[GV_NT], s_red[NW];
  __shared__ __align__(16) float s_part[2][GV4_KS][GV_NT];
  const long long bm = (long long)b * nmod + mod;
  for (int k = tid; k < Mdim; k += GV_THREADS) s_in[k] = __ldg(g + bm * ldg + k);
  __syncthreads();
  const int kper = (Mdim + GV4_KS - 1) / GV4_KS;
  const int k0 = ks * kper, k1 = min(Mdim, k0 + kper);
  const bool colok = col < Mdim;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (colok && mode != 2) {
    const float* W = wg + mod * w_mstride + col;
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(W + (long long)k * Mdim));
      const float x = s_in[k];
      v.x = fmaf(x, w.x, v.x); v.y = fmaf(x, w.y, v.y); v.z = fmaf(x, w.z, v.z); v.w = fmaf(x, w.w, v.w);
    }
  }
  *reinterpret_cast<float4*>(&s_part[0][ks][col]) = v;
  __syncthreads();
  const int n = tid;                       // second role of the first GV_NT threads: one output column each
  float val = 0.f, ss = 0.f;
  if (n < Mdim) {
    if (mode == 2) {
      val = gv_out[bm * ldgate + n];
    } else {
#pragma unroll
      for (int j = 0; j < GV4_KS; ++j) val += s_part[0][j][n];
      val += __ldg(gvl + (long long)b * gvl_bstride + (long long)mod * ldgvl + n);
    }
    ss = val * val;
  }
  ss = warp_sum(ss);
  if (lane == 0) s_red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < NW; ++w) tot += s_red[w];
  if (mode == 1) {
    if (n < ldgate) gv_out[bm * ldgate + n] = n < Mdim ? val : 0.f;
    if (tid == 0) atomicAdd(batch_ss + mod, tot);
    return;
  }
  if (mode == 2) tot = batch_ss[mod];
  if (n < GV_NT) {
    const float gvn = n < Mdim ? val * rsqrtf(fmaxf(tot, 1e-12f)) : 0.f;
    s_gv[n] = gvn;
    if (n < ldgate) gv_out[bm * ldgate + n] = gvn;
  }
  __syncthreads();
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  if (colok) {
    const float* W1 = wf1 + mod * w_mstride + col;
    const float* W2 = wf2 + mod * w_mstride + col;
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(W1 + (long long)k * Mdim));
      const float4 w2 = __ldg(reinterpret_cast<const float4*>(W2 + (long long)k * Mdim));
      const float x = s_gv[k];
      a1.x = fmaf(x, w1.x, a1.x); a1.y = fmaf(x, w1.y, a1.y); a1.z = fmaf(x, w1.z, a1.z); a1.w = fmaf(x, w1.w, a1.w);
      a2.x = fmaf(x, w2.x, a2.x); a2.y = fmaf(x, w2.y, a2.y); a2.z = fmaf(x, w2.z, a2.z); a2.w = fmaf(x, w2.w, a2.w);
    }
  }
  *reinterpret_cast<float4*>(&s_part[0][ks][col]) = a1;      // every thread is past its reads of the stage-1 partials
  *reinterpret_cast<float4*>(&s_part[1][ks][col]) = a2;
  __syncthreads();
  float* o1 = gate1 + b * g1_bstride + mod * g1_mstride;
  float* o2 = gate2 + b * g2_bstride + mod * g2_mstride;
  if (n < Mdim) {
    float t1 = __ldg(bf1 + mod * b_mstride + n), t2 = __ldg(bf2 + mod * b_mstride + n);
#pragma unroll
    for (int j = 0; j < GV4_KS; ++j) { t1 += s_part[0][j][n]; t2 += s_part[1][j][n]; }
    o1[n] = sigmoid_acc(t1);
    o2[n] = sigmoid_acc(t2);
  } else if (n < ldgate) {
    o1[n] = 0.f;
    o2[n] = 0.f;
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_words_prepare(const float* lstm_outputs, int32_t rows, int32_t r, float* words_f32, void* words_f16,
                                  int64_t ld16, float* seq_mask, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(lstm_outputs && words_f32 && words_f16 && seq_mask && rows > 0 && r > 0 && ld16 >= r, CMPC_ERR_ARG,
               "cmpc_words_prepare: bad args");
  const int threads = 128;
  words_prepare_kernel<<<(rows * 32 + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
      lstm_outputs, rows, r, words_f32, (__half*)words_f16, ld16, seq_mask);
  return check_launch("words_prepare_kernel");
}

extern "C" int cmpc_lang_parse(const float* hidden, int64_t ldh, int32_t hid, const float* w2, const float* b2,
                               const float* words_f32, const float* seq_mask, int32_t batch, int32_t t, int32_t r,
                               int32_t c, float* parse, float* rgate, float* valid_f32, float* nec_f32, void* valid_f16,
                               void* nec_f16, int64_t ld16, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE((hidden == nullptr || (w2 && b2 && seq_mask)) && words_f32 && parse && rgate && valid_f32 && nec_f32 && valid_f16 && nec_f16,
               CMPC_ERR_ARG, "cmpc_lang_parse: null pointer");
  CMPC_REQUIRE(batch > 0 && t > 0 && t <= 32 && r > 0 && r <= 8 * PARSE_THREADS && (hidden == nullptr || hid > 0) && ld16 >= r, CMPC_ERR_ARG,
               "cmpc_lang_parse: need T <= 32 and R <= 2048");
  CMPC_REQUIRE(hidden == nullptr || (reinterpret_cast<uintptr_t>(w2) & 15) == 0, CMPC_ERR_ALIGN, "cmpc_lang_parse: w2 must be 16-byte aligned");
  const size_t smem = (size_t)(t * 4 + 2 * t + 2 * (PARSE_THREADS / 32)) * sizeof(float);
  lang_parse_kernel<<<batch, PARSE_THREADS, smem, (cudaStream_t)stream>>>(
      hidden, ldh, hid, w2, b2, words_f32, seq_mask, t, r, 1.0f / sqrtf((float)c), parse, rgate, valid_f32, nec_f32,
      (__half*)valid_f16, (__half*)nec_f16, ld16);
  return check_launch("lang_parse_kernel");
}

extern "C" int cmpc_small_linear_f32(const float* x, int64_t ldx, int64_t x_zstride, const float* w, int64_t ldw,
                                     int64_t w_zstride, const float* bias, int64_t b_zstride, float* out, int64_t ldo,
                                     int64_t o_zstride, int32_t nbatch, int32_t rows, int32_t k, int32_t n, int32_t act,
                                     void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(x && w && out && nbatch > 0 && rows > 0 && k > 0 && n > 0, CMPC_ERR_ARG, "cmpc_small_linear_f32: bad args");
  const int nblk = ((n + SL_THREADS - 1) / SL_THREADS) * ((rows + SL_RB - 1) / SL_RB) * nbatch;
  int ksplit = 1;
  if (act == 4 && k >= 4 * SL_KC) {                 // accumulate mode: split K until ~2 blocks per SM
    ksplit = (2 * num_sms() + nblk - 1) / nblk;
    if (ksplit > k / (2 * SL_KC)) ksplit = k / (2 * SL_KC);
    if (ksplit < 1) ksplit = 1;
  }
  dim3 grid((n + SL_THREADS - 1) / SL_THREADS, ((rows + SL_RB - 1) / SL_RB) * ksplit, nbatch);
  small_linear_kernel<<<grid, SL_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, x_zstride, w, ldw, w_zstride, bias, b_zstride,
                                                                      out, ldo, o_zstride, rows, k, n, act, ksplit);
  return check_launch("small_linear_kernel");
}

static int launch_gv_gates(const float* g, int64_t ldg, const float* gvl, int64_t ldgvl, int64_t gvl_bstride, const float* wg,
                           const float* wf1, const float* bf1, const float* wf2, const float* bf2, int64_t w_mstride,
                           int64_t b_mstride, int32_t batch, int32_t nmod, int32_t mdim, float* gv, float* gate1,
                           float* gate2, int64_t ldgate, int mode, float* batch_ss, void* stream, int64_t g1_bstride = 0,
                           int64_t g1_mstride = 0, int64_t g2_bstride = 0, int64_t g2_mstride = 0) {
  if (g1_bstride == 0 && g1_mstride == 0) { g1_bstride = (int64_t)nmod * ldgate; g1_mstride = ldgate; }     // default [B, nmod, ldgate]
  if (g2_bstride == 0 && g2_mstride == 0) { g2_bstride = (int64_t)nmod * ldgate; g2_mstride = ldgate; }
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(g && gvl && wg && wf1 && bf1 && wf2 && bf2 && gv && gate1 && gate2, CMPC_ERR_ARG, "cmpc_gv_gates: null pointer");
  CMPC_REQUIRE(batch > 0 && nmod > 0 && mdim > 0 && mdim <= 512 && ldgate >= mdim && ldgate <= 512, CMPC_ERR_ARG,
               "cmpc_gv_gates: mlp_dim must be <= 512");
  CMPC_REQUIRE(mode == 0 || batch_ss, CMPC_ERR_ARG, "cmpc_gv_gates_batch: batch_ss is null");
  const bool v4 = mdim % 4 == 0 && w_mstride % 4 == 0 && ((reinterpret_cast<uintptr_t>(wg) | reinterpret_cast<uintptr_t>(wf1) |
                                                          reinterpret_cast<uintptr_t>(wf2)) & 15) == 0;
  if (v4)
    gv_gates_v4_kernel<<<dim3(batch, nmod), GV_THREADS, 0, (cudaStream_t)stream>>>(g, ldg, gvl, ldgvl, gvl_bstride, wg, wf1, bf1, wf2, bf2,
                                                                                    w_mstride, b_mstride, nmod, mdim, gv, gate1, gate2, ldgate,
                                                                                    mode, batch_ss, g1_bstride, g1_mstride, g2_bstride, g2_mstride);
  else
    gv_gates_kernel<<<dim3(batch, nmod), GV_THREADS, 0, (cudaStream_t)stream>>>(g, ldg, gvl, ldgvl, gvl_bstride, wg, wf1, bf1, wf2, bf2,
                                                                                 w_mstride, b_mstride, nmod, mdim, gv, gate1, gate2, ldgate,
                                                                                 mode, batch_ss, g1_bstride, g1_mstride, g2_bstride, g2_mstride);
  return check_launch("gv_gates_kernel");
}

extern "C" int cmpc_gv_gates(const float* g, int64_t ldg, const float* gvl, int64_t ldgvl, int64_t gvl_bstride, const float* wg,
                             const float* wf1, const float* bf1, const float* wf2, const float* bf2, int64_t w_mstride,
                             int64_t b_mstride, int32_t batch, int32_t nmod, int32_t mdim, float* gv, float* gate1,
                             float* gate2, int64_t ldgate, void* stream) {
  return launch_gv_gates(g, ldg, gvl, ldgvl, gvl_bstride, wg, wf1, bf1, wf2, bf2, w_mstride, b_mstride, batch, nmod, mdim, gv, gate1, gate2,
                         ldgate, 0, nullptr, stream);
}

extern "C" int cmpc_gv_gates_batch(const float* g, int64_t ldg, const float* gvl, int64_t ldgvl, int64_t gvl_bstride, const float* wg,
                                   const float* wf1, const float* bf1, const float* wf2, const float* bf2, int64_t w_mstride,
                                   int64_t b_mstride, int32_t batch, int32_t nmod, int32_t mdim, float* gv, float* gate1,
                                   float* gate2, int64_t ldgate, int32_t phase, float* batch_ss, void* stream) {
  CMPC_REQUIRE(phase == 1 || phase == 2, CMPC_ERR_ARG, "cmpc_gv_gates_batch: phase must be 1 or 2");
  return launch_gv_gates(g, ldg, gvl, ldgvl, gvl_bstride, wg, wf1, bf1, wf2, bf2, w_mstride, b_mstride, batch, nmod, mdim, gv, gate1, gate2,
                         ldgate, phase, batch_ss, stream);
}

extern "C" int cmpc_gv_gates_ex(const float* g, int64_t ldg, const float* gvl, int64_t ldgvl, int64_t gvl_bstride, const float* wg,
                                const float* wf1, const float* bf1, const float* wf2, const float* bf2, int64_t w_mstride, int64_t b_mstride,
                                int32_t batch, int32_t nmod, int32_t mdim, float* gv, float* gate1, int64_t g1_bstride, int64_t g1_mstride,
                                float* gate2, int64_t g2_bstride, int64_t g2_mstride, int64_t ldgate, int32_t phase, float* batch_ss,
                                void* stream) {
  CMPC_REQUIRE(phase >= 0 && phase <= 2, CMPC_ERR_ARG, "cmpc_gv_gates_ex: phase must be 0 (per-sample norm), 1 or 2 (batch-coupled)");
  return launch_gv_gates(g, ldg, gvl, ldgvl, gvl_bstride, wg, wf1, bf1, wf2, bf2, w_mstride, b_mstride, batch, nmod, mdim, gv, gate1, gate2,
                         ldgate, phase, batch_ss, stream, g1_bstride, g1_mstride, g2_bstride, g2_mstride);
}

// ConvLSTM fusion of the three exchanged levels (util/cell.py:36-79, driven by dynamic_rnn at
// CMPC_model.py:287-290).  The 1x1 "convolution" y = [x | h] K is the tcgen05 GEMM (its epilogue adds the
// W_ci / W_cf peepholes and accumulates the whole-map layer-norm statistics of j, i, f, o); the two kernels here
// are the HBM-bound gate math around the two rounds of whole-sample layer norms:
//   gates1:  j,i,f = LN(.) ; c' = c*sigmoid(f+1) + sigmoid(i)*tanh(j) ; o' = o + W_co*c' ; stats(o'), stats(c')
//   gates2:  o = LN(o') ; c = LN(c') ; h = sigmoid(o) * tanh(c)
// Layout: y fp32 [rows, 4*GW] (gate g at columns [g*GW, g*GW + M)), state fp32 [rows, GW], GW = padded M.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

// (mean, rstd) pairs come from cmpc_ln_finalize (fp64 sums -> fp32 once per sample/gate, not once per thread)
__device__ __forceinline__ void ln_ms(const float* mr, long long idx, float& mean, float& rstd) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(mr) + idx);
  mean = t.x;
  rstd = t.y;
}

constexpr int G_THREADS = 256;


// Thread layout: a thread owns one fixed float4 column group (so the layer-norm gamma / beta of its channels are loaded
// once, outside the loop) and walks down a contiguous range of rows of ONE sample: grid = (row chunks, samples).  The
// statistics of o' and c' are therefore reduced inside the block and cost 4 fp64 atomics per block; a strided row
// assignment made every warp flush on every iteration and the ~800k contended fp64 atomics cost 4x the kernel's HBM time.
__device__ __forceinline__ float4 ld_gate4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_gate4(const __half* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// cell-state element type CT = float (training, reference layout) or __half (inference: the state is kept in fp16 between passes)
__device__ __forceinline__ void st_state4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_state4(__half* p, const float4& v) {
  __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(p) = u;
}

// YT = float or __half: the gate pre-activations written by the GEMM (fp16 halves the largest tensor of the head; the
// layer-norm statistics were taken from the fp32 accumulators)
template <typename YT, typename CT>
__global__ void __launch_bounds__(G_THREADS, 4)
convlstm_gates1_kernel(const YT* __restrict__ y, long long ldy, int GW, int M, const float* __restrict__ stats_in /*[B,4] (mean,rstd)*/,
                       const float* __restrict__ ln_gamma /*[5,GW]*/, const float* __restrict__ ln_beta,
                       const CT* __restrict__ cprev /*or null*/, const float* __restrict__ w_co /*[pix,GW]*/,
                       CT* __restrict__ cnew, float* __restrict__ opre, double* __restrict__ stats_out /*[B,2,2]*/,
                       int rows_per_sample, int rows_per_chunk) {
  const int gpr = GW / 4;                       // float4 groups per row
  const int rows_per_iter = G_THREADS / gpr;
  const int g = threadIdx.x % gpr;
  const int sub = threadIdx.x / gpr;
  const int c = g * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const bool col_ok = c < M;                    // M % 4 == 0: a group is entirely valid or entirely padding
  float4 gj, bj, gi, bi, gf, bf;
  if (col_ok) {
    gj = __ldg(reinterpret_cast<const float4*>(ln_gamma + c));           bj = __ldg(reinterpret_cast<const float4*>(ln_beta + c));
    gi = __ldg(reinterpret_cast<const float4*>(ln_gamma + GW + c));      bi = __ldg(reinterpret_cast<const float4*>(ln_beta + GW + c));
    gf = __ldg(reinterpret_cast<const float4*>(ln_gamma + 2 * GW + c));  bf = __ldg(reinterpret_cast<const float4*>(ln_beta + 2 * GW + c));
  }
  float mj, rj, mi, ri, mf, rf;
  ln_ms(stats_in, (long long)b * 4 + 0, mj, rj);
  ln_ms(stats_in, (long long)b * 4 + 1, mi, ri);
  ln_ms(stats_in, (long long)b * 4 + 2, mf, rf);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int p0 = blockIdx.x * rows_per_chunk;
  const int p1 = min(rows_per_sample, p0 + rows_per_chunk);
  for (int pix = p0 + sub; pix < p1; pix += rows_per_iter) {
    const long long row = (long long)b * rows_per_sample + pix;
    float4 cn = make_float4(0.f, 0.f, 0.f, 0.f), op = cn;
    if (col_ok) {
      const YT* yr = y + row * ldy + c;
      const float4 vj = ld_gate4(yr);
      const float4 vi = ld_gate4(yr + GW);
      const float4 vf = ld_gate4(yr + 2 * GW);
      const float4 vo = ld_gate4(yr + 3 * GW);
      const float4 cp = cprev ? ld_gate4(cprev + row * GW + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 wc = __ldg(reinterpret_cast<const float4*>(w_co + (long long)pix * GW + c));
      const float aj[4] = {vj.x, vj.y, vj.z, vj.w}, ai[4] = {vi.x, vi.y, vi.z, vi.w}, af[4] = {vf.x, vf.y, vf.z, vf.w}, ao[4] = {vo.x, vo.y, vo.z, vo.w};
      const float ggj[4] = {gj.x, gj.y, gj.z, gj.w}, bbj[4] = {bj.x, bj.y, bj.z, bj.w};
      const float ggi[4] = {gi.x, gi.y, gi.z, gi.w}, bbi[4] = {bi.x, bi.y, bi.z, bi.w};
      const float ggf[4] = {gf.x, gf.y, gf.z, gf.w}, bbf[4] = {bf.x, bf.y, bf.z, bf.w};
      const float acp[4] = {cp.x, cp.y, cp.z, cp.w}, awc[4] = {wc.x, wc.y, wc.z, wc.w};
      float rc[4], ro[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float jn = (aj[e] - mj) * rj * ggj[e] + bbj[e];
        const float in = (ai[e] - mi) * ri * ggi[e] + bbi[e];
        const float fn = (af[e] - mf) * rf * ggf[e] + bbf[e];
        rc[e] = acp[e] * sigmoid_acc(fn + 1.0f) + sigmoid_acc(in) * tanh_acc(jn);
        ro[e] = ao[e] + awc[e] * rc[e];
        acc[0] += ro[e]; acc[1] += ro[e] * ro[e];
        acc[2] += rc[e]; acc[3] += rc[e] * rc[e];
      }
      cn = make_float4(rc[0], rc[1], rc[2], rc[3]);
      op = make_float4(ro[0], ro[1], ro[2], ro[3]);
    }
    st_state4(cnew + row * GW + c, cn);
    if (opre) *reinterpret_cast<float4*>(opre + row * GW + c) = op;
  }
  // block reduction -> 4 fp64 atomics per block
  __shared__ float s_red[G_THREADS / 32][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) { s_red[warp][0] = acc[0]; s_red[warp][1] = acc[1]; s_red[warp][2] = acc[2]; s_red[warp][3] = acc[3]; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < G_THREADS / 32; ++w) t += (double)s_red[w][threadIdx.x];
    atomicAdd(stats_out + (long long)b * 4 + threadIdx.x, t);
  }
}

// FROM_Y: o' is not read from an fp32 map but recomputed as o + W_co * c' from the GEMM's fp16 gate map (y16o = column block 3 of y,
// leading dimension ldy) and the per-pixel peephole weights -- the same fp32 expression gates1 took its statistics from; saves
// writing and re-reading an fp32 [rows, GW] map per step.  c_out may be null (last step: only h is consumed).
// Same thread layout as gates1: a thread owns one float4 column group (gamma / beta of o and c live in registers) and walks down
// a contiguous chunk of one sample's rows, G2_ROWS rows in flight; the flat grid-stride form re-read gamma / beta and the
// statistics for every element and ran at 3.4 TB/s.
constexpr int G2_ROWS = 4;
template <bool FROM_Y, typename CT>
__global__ void __launch_bounds__(G_THREADS, 4)
convlstm_gates2_kernel(const float* __restrict__ opre, const __half* __restrict__ y16o, long long ldy, const float* __restrict__ w_co,
                       const CT* __restrict__ cnew, int GW, int M,
                       const float* __restrict__ stats /*[B,2] (mean,rstd): o', c'*/, const float* __restrict__ ln_gamma /*[5,GW]*/,
                       const float* __restrict__ ln_beta, CT* __restrict__ c_out, __half* __restrict__ h16,
                       float* __restrict__ h32 /*or null*/, int rows_per_sample, int rows_per_chunk) {
  const int gpr = GW / 4;                       // float4 groups per row
  const int rows_per_iter = G_THREADS / gpr;
  const int g = threadIdx.x % gpr, sub = threadIdx.x / gpr;
  const int c = g * 4;
  const int b = blockIdx.y;
  const bool col_ok = c < M;                    // M % 4 == 0: a group is entirely valid or entirely padding
  float4 go = make_float4(0.f, 0.f, 0.f, 0.f), bo = go, gc = go, bc = go;
  if (col_ok) {
    go = __ldg(reinterpret_cast<const float4*>(ln_gamma + 3 * GW + c)); bo = __ldg(reinterpret_cast<const float4*>(ln_beta + 3 * GW + c));
    gc = __ldg(reinterpret_cast<const float4*>(ln_gamma + 4 * GW + c)); bc = __ldg(reinterpret_cast<const float4*>(ln_beta + 4 * GW + c));
  }
  float mo, ro, mc, rcs;
  ln_ms(stats, (long long)b * 2 + 0, mo, ro);
  ln_ms(stats, (long long)b * 2 + 1, mc, rcs);
  // o_n = o' * Ao + Bo,  c_n = c' * Ac + Bc
  const float Ao[4] = {ro * go.x, ro * go.y, ro * go.z, ro * go.w}, Ac[4] = {rcs * gc.x, rcs * gc.y, rcs * gc.z, rcs * gc.w};
  const float Bo[4] = {bo.x - mo * Ao[0], bo.y - mo * Ao[1], bo.z - mo * Ao[2], bo.w - mo * Ao[3]};
  const float Bc[4] = {bc.x - mc * Ac[0], bc.y - mc * Ac[1], bc.z - mc * Ac[2], bc.w - mc * Ac[3]};
  const int p0 = blockIdx.x * rows_per_chunk;
  const int p1 = min(rows_per_sample, p0 + rows_per_chunk);
  for (int pix0 = p0 + sub; pix0 < p1; pix0 += rows_per_iter * G2_ROWS) {
    float4 vo[G2_ROWS], vc[G2_ROWS];
#pragma unroll
    for (int i = 0; i < G2_ROWS; ++i) {           // all loads of the pass first
      const int pix = pix0 + i * rows_per_iter;
      vo[i] = make_float4(0.f, 0.f, 0.f, 0.f); vc[i] = vo[i];
      if (col_ok && pix < p1) {
        const long long r = (long long)b * rows_per_sample + pix;
        vc[i] = ld_gate4(cnew + r * GW + c);
        if (FROM_Y) {
          const float4 yo = ld_gate4(y16o + r * ldy + c);
          const float4 wc = __ldg(reinterpret_cast<const float4*>(w_co + (long long)pix * GW + c));
          vo[i] = make_float4(yo.x + wc.x * vc[i].x, yo.y + wc.y * vc[i].y, yo.z + wc.z * vc[i].z, yo.w + wc.w * vc[i].w);
        } else {
          vo[i] = __ldg(reinterpret_cast<const float4*>(opre + r * GW + c));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < G2_ROWS; ++i) {
      const int pix = pix0 + i * rows_per_iter;
      if (pix >= p1) break;
      const long long r = (long long)b * rows_per_sample + pix;
      float rc[4] = {0.f, 0.f, 0.f, 0.f}, rh[4] = {0.f, 0.f, 0.f, 0.f};
      if (col_ok) {
        const float ao[4] = {vo[i].x, vo[i].y, vo[i].z, vo[i].w}, ac[4] = {vc[i].x, vc[i].y, vc[i].z, vc[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float on = fmaf(ao[e], Ao[e], Bo[e]);
          const float cn = fmaf(ac[e], Ac[e], Bc[e]);
          rc[e] = cn;
          rh[e] = sigmoid_acc(on) * tanh_acc(cn);
        }
      }
      if (c_out) st_state4(c_out + r * GW + c, make_float4(rc[0], rc[1], rc[2], rc[3]));
      if (h32) *reinterpret_cast<float4*>(h32 + r * GW + c) = make_float4(rh[0], rh[1], rh[2], rh[3]);
      __half2 h0 = __floats2half2_rn(rh[0], rh[1]), h1 = __floats2half2_rn(rh[2], rh[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(h16 + r * GW + c) = u;
    }
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_convlstm_gates1(const void* y, int32_t y_fp16, int64_t ldy, int32_t gw, int32_t m, const float* stats_in,
                                    const float* ln_gamma, const float* ln_beta, const float* cprev, const float* w_co,
                                    float* cnew, float* opre, double* stats_out, int64_t rows, int32_t rows_per_sample,
                                    void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(y && stats_in && ln_gamma && ln_beta && w_co && cnew && stats_out, CMPC_ERR_ARG, "cmpc_convlstm_gates1: null pointer");   // opre may be null (cmpc_convlstm_gates2_y16 recomputes it)
  CMPC_REQUIRE(rows > 0 && rows < (1ll << 31) && rows_per_sample > 0 && m > 0 && m % 4 == 0 && gw >= m && ldy >= 4 * (int64_t)gw && ldy % 4 == 0,
               CMPC_ERR_ARG, "cmpc_convlstm_gates1: bad shape");
  CMPC_REQUIRE(gw % 128 == 0 && (G_THREADS * 4) % gw == 0, CMPC_ERR_ARG, "cmpc_convlstm_gates1: gw must be 128, 256, 512 or 1024");
  CMPC_REQUIRE(rows % rows_per_sample == 0, CMPC_ERR_ARG, "cmpc_convlstm_gates1: rows must be a multiple of rows_per_sample");
  const int batch = (int)(rows / rows_per_sample);
  // ~16 blocks per SM in total, each on a contiguous chunk of one sample's rows
  int chunks = (num_sms() * 16 + batch - 1) / batch;
  if (chunks > rows_per_sample) chunks = rows_per_sample;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  chunks = (rows_per_sample + rows_per_chunk - 1) / rows_per_chunk;
  if (y_fp16)
    convlstm_gates1_kernel<__half, float><<<dim3(chunks, batch), G_THREADS, 0, (cudaStream_t)stream>>>((const __half*)y, ldy, gw, m, stats_in, ln_gamma, ln_beta,
                                                                                                 cprev, w_co, cnew, opre, stats_out, rows_per_sample, rows_per_chunk);
  else
    convlstm_gates1_kernel<float, float><<<dim3(chunks, batch), G_THREADS, 0, (cudaStream_t)stream>>>((const float*)y, ldy, gw, m, stats_in, ln_gamma, ln_beta,
                                                                                                cprev, w_co, cnew, opre, stats_out, rows_per_sample, rows_per_chunk);
  return check_launch("convlstm_gates1_kernel");
}

// grid of both gates2 forms: ~16 blocks per SM in total, each on a contiguous chunk of one sample's rows
static void gates2_grid(int64_t rows, int32_t rows_per_sample, int& chunks, int& rows_per_chunk, int& batch) {
  batch = (int)(rows / rows_per_sample);
  chunks = (num_sms() * 16 + batch - 1) / batch;
  if (chunks > rows_per_sample) chunks = rows_per_sample;
  if (chunks < 1) chunks = 1;
  rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  chunks = (rows_per_sample + rows_per_chunk - 1) / rows_per_chunk;
}

extern "C" int cmpc_convlstm_gates2(const float* opre, const float* cnew, int32_t gw, int32_t m, const float* stats,
                                    const float* ln_gamma, const float* ln_beta, float* c_out, void* h_f16, float* h_f32,
                                    int64_t rows, int32_t rows_per_sample, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(opre && cnew && stats && ln_gamma && ln_beta && h_f16, CMPC_ERR_ARG, "cmpc_convlstm_gates2: null pointer");   // c_out may be null
  CMPC_REQUIRE(rows > 0 && rows < (1ll << 31) && rows_per_sample > 0 && rows % rows_per_sample == 0 && m > 0 && m % 4 == 0 && gw >= m,
               CMPC_ERR_ARG, "cmpc_convlstm_gates2: bad shape");
  CMPC_REQUIRE(gw % 128 == 0 && (G_THREADS * 4) % gw == 0, CMPC_ERR_ARG, "cmpc_convlstm_gates2: gw must be 128, 256, 512 or 1024");
  int chunks, rows_per_chunk, batch;
  gates2_grid(rows, rows_per_sample, chunks, rows_per_chunk, batch);
  convlstm_gates2_kernel<false, float><<<dim3(chunks, batch), G_THREADS, 0, (cudaStream_t)stream>>>(opre, nullptr, 0, nullptr, cnew, gw, m, stats, ln_gamma,
                                                                                              ln_beta, c_out, (__half*)h_f16, h_f32, rows_per_sample,
                                                                                              rows_per_chunk);
  return check_launch("convlstm_gates2_kernel");
}

extern "C" int cmpc_convlstm_gates2_y16(const void* y_o_f16, int64_t ldy, const float* w_co, const void* cnew, int32_t gw, int32_t m,
                                        const float* stats, const float* ln_gamma, const float* ln_beta, void* c_out, void* h_f16,
                                        int32_t state_f16, int64_t rows, int32_t rows_per_sample, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(y_o_f16 && w_co && cnew && stats && ln_gamma && ln_beta && h_f16, CMPC_ERR_ARG, "cmpc_convlstm_gates2_y16: null pointer");
  CMPC_REQUIRE(rows > 0 && rows < (1ll << 31) && rows_per_sample > 0 && rows % rows_per_sample == 0 && m > 0 && m % 4 == 0 && gw >= m && ldy >= gw &&
               ldy % 4 == 0, CMPC_ERR_ARG, "cmpc_convlstm_gates2_y16: bad shape");
  CMPC_REQUIRE(gw % 128 == 0 && (G_THREADS * 4) % gw == 0, CMPC_ERR_ARG, "cmpc_convlstm_gates2_y16: gw must be 128, 256, 512 or 1024");
  CMPC_REQUIRE((reinterpret_cast<uintptr_t>(y_o_f16) & 7) == 0, CMPC_ERR_ALIGN, "cmpc_convlstm_gates2_y16: y must be 8-byte aligned");
  int chunks, rows_per_chunk, batch;
  gates2_grid(rows, rows_per_sample, chunks, rows_per_chunk, batch);
  if (state_f16)
    convlstm_gates2_kernel<true, __half><<<dim3(chunks, batch), G_THREADS, 0, (cudaStream_t)stream>>>(
        nullptr, (const __half*)y_o_f16, ldy, w_co, (const __half*)cnew, gw, m, stats, ln_gamma, ln_beta, (__half*)c_out, (__half*)h_f16, nullptr,
        rows_per_sample, rows_per_chunk);
  else
    convlstm_gates2_kernel<true, float><<<dim3(chunks, batch), G_THREADS, 0, (cudaStream_t)stream>>>(
        nullptr, (const __half*)y_o_f16, ldy, w_co, (const float*)cnew, gw, m, stats, ln_gamma, ln_beta, (float*)c_out, (__half*)h_f16, nullptr,
        rows_per_sample, rows_per_chunk);
  return check_launch("convlstm_gates2_kernel");
}

extern "C" int cmpc_convlstm_gates1_h16(const void* y_f16, int64_t ldy, int32_t gw, int32_t m, const float* stats_in, const float* ln_gamma,
                                        const float* ln_beta, const void* cprev_f16, const float* w_co, void* cnew_f16,
                                        double* stats_out, int64_t rows, int32_t rows_per_sample, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(y_f16 && stats_in && ln_gamma && ln_beta && w_co && cnew_f16 && stats_out, CMPC_ERR_ARG, "cmpc_convlstm_gates1_h16: null pointer");
  CMPC_REQUIRE(rows > 0 && rows < (1ll << 31) && rows_per_sample > 0 && m > 0 && m % 4 == 0 && gw >= m && ldy >= 4 * (int64_t)gw && ldy % 4 == 0,
               CMPC_ERR_ARG, "cmpc_convlstm_gates1_h16: bad shape");
  CMPC_REQUIRE(gw % 128 == 0 && (G_THREADS * 4) % gw == 0, CMPC_ERR_ARG, "cmpc_convlstm_gates1_h16: gw must be 128, 256, 512 or 1024");
  CMPC_REQUIRE(rows % rows_per_sample == 0, CMPC_ERR_ARG, "cmpc_convlstm_gates1_h16: rows must be a multiple of rows_per_sample");
  const int batch = (int)(rows / rows_per_sample);
  int chunks = (num_sms() * 16 + batch - 1) / batch;
  if (chunks > rows_per_sample) chunks = rows_per_sample;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  chunks = (rows_per_sample + rows_per_chunk - 1) / rows_per_chunk;
  convlstm_gates1_kernel<__half, __half><<<dim3(chunks, batch), G_THREADS, 0, (cudaStream_t)stream>>>(
      (const __half*)y_f16, ldy, gw, m, stats_in, ln_gamma, ln_beta, (const __half*)cprev_f16, w_co, (__half*)cnew_f16, nullptr, stats_out,
      rows_per_sample, rows_per_chunk);
  return check_launch("convlstm_gates1_kernel");
}

// Backward of one ConvLSTM step (util/cell.py:36-79; TF autodiff at CMPC_model.py:461), the HBM-bound part around the
// five whole-sample layer norms.  Forward of the step (see convlstm.cu):
//   y = [x | h_prev] K  (gates j, i', f', o; the GEMM epilogue already added W_ci*c_prev, W_cf*c_prev to i', f')
//   jn, in, fn = LN0(j), LN1(i'), LN2(f');  c' = c_prev*sigmoid(fn+1) + sigmoid(in)*tanh(jn)
//   o' = o + W_co*c';  on = LN3(o');  cn = LN4(c');  h = sigmoid(on)*tanh(cn);  state = cn
// A whole-sample layer norm y = xh*gamma + beta, xh = (x - mean)*rstd, back-propagates as
//   dx = rstd * (g - mean_s(g) - xh * mean_s(g*xh)),  g = dy*gamma        (mean_s over the N*M elements of the sample)
// so the step needs two rounds of per-sample sums before it can emit gradients -> three phases:
//   phase 1: sums for LN3 / LN4 (+ their dgamma / dbeta)
//   phase 2: d o', d c' (stored), dW_co, sums for LN0-2 (+ dgamma / dbeta)
//   phase 3: d j, d i', d f' (-> dy for the dgrad / wgrad GEMMs), d c_prev, dW_ci, dW_cf
// No atomics: a block owns `rows_per_iter` pixels and loops over the batch, so peephole and per-channel sums (both are
// sums over the batch per pixel-channel) are plain stores into per-pixel partials [N, Q, GW], and per-sample sums go to
// per-block partials [blocks, B, K]; cmpc_convlstm_bwd_reduce folds the partials.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

constexpr int LB_THREADS = 256;

__device__ __forceinline__ float4 ldf4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldh4(const __half* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void sth4(__half* p, const float (&v)[4]) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(p) = u;
}
#define F4A(name, v) const float name[4] = {(v).x, (v).y, (v).z, (v).w}

template <int PHASE>
__global__ void __launch_bounds__(LB_THREADS)
convlstm_bwd_kernel(const cmpc_convlstm_bwd_args a, int batch, int rows_per_iter) {
  const int GW = a.gw, M = a.m, N = a.rows_per_sample;
  const int gpr = GW / 4;
  const int g = threadIdx.x % gpr, sub = threadIdx.x / gpr;
  const int c = g * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pix = blockIdx.x * rows_per_iter + sub;
  const bool live = c < M && pix < N && sub < rows_per_iter;
  constexpr int K = PHASE == 1 ? 4 : 6;          // per-sample sums produced by this phase
  constexpr int Q = PHASE == 1 ? 4 : (PHASE == 2 ? 6 : 0);
  const float inv_n = 1.0f / ((float)N * (float)M);
  __shared__ float s_red[LB_THREADS / 32][6];

  float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 g0 = z4, b0 = z4, g1 = z4, b1 = z4, g2 = z4, b2 = z4, g3 = z4, b3 = z4, g4 = z4, b4 = z4, wci = z4, wcf = z4, wco = z4;
  if (live) {
    g0 = ldf4(a.ln_gamma + c);           b0 = ldf4(a.ln_beta + c);
    g1 = ldf4(a.ln_gamma + GW + c);      b1 = ldf4(a.ln_beta + GW + c);
    g2 = ldf4(a.ln_gamma + 2 * GW + c);  b2 = ldf4(a.ln_beta + 2 * GW + c);
    g3 = ldf4(a.ln_gamma + 3 * GW + c);  b3 = ldf4(a.ln_beta + 3 * GW + c);
    g4 = ldf4(a.ln_gamma + 4 * GW + c);  b4 = ldf4(a.ln_beta + 4 * GW + c);
    wco = ldf4(a.w_co + (long long)pix * GW + c);
    if (a.cprev) { wci = ldf4(a.w_ci + (long long)pix * GW + c); wcf = ldf4(a.w_cf + (long long)pix * GW + c); }
  }
  F4A(G0, g0); F4A(B0, b0); F4A(G1, g1); F4A(B1, b1); F4A(G2, g2); F4A(B2, b2); F4A(G3, g3); F4A(B3, b3); F4A(G4, g4); F4A(B4, b4);
  F4A(WCI, wci); F4A(WCF, wcf); F4A(WCO, wco);
  float chan[Q > 0 ? Q : 1][4];                  // per-channel sums over the batch (this pixel)
#pragma unroll
  for (int q = 0; q < (Q > 0 ? Q : 1); ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) chan[q][e] = 0.f;
  float pco[4] = {0.f, 0.f, 0.f, 0.f}, pci[4] = {0.f, 0.f, 0.f, 0.f}, pcf[4] = {0.f, 0.f, 0.f, 0.f};   // peephole gradients

  for (int b = 0; b < batch; ++b) {
    const long long row = (long long)b * N + pix;
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    if (live) {
      // ---- recompute the output side of the step ----
      const float2 so = __ldg(reinterpret_cast<const float2*>(a.mr_o) + b * 2 + 0), sc = __ldg(reinterpret_cast<const float2*>(a.mr_o) + b * 2 + 1);
      const float4 vop = ldf4(a.opre + row * GW + c), vcw = ldf4(a.cnew + row * GW + c), vcn = ldf4(a.cn + row * GW + c);
      const float4 vdh = ldf4(a.dh + row * a.ld_dh + c);
      const float4 vdc = a.dcn_in ? ldf4(a.dcn_in + row * GW + c) : z4;
      F4A(OP, vop); F4A(CW, vcw); F4A(CN, vcn); F4A(DH, vdh); F4A(DC, vdc);
      float d_on[4], d_cn[4], xh3[4], xh4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xh3[e] = (OP[e] - so.x) * so.y;
        xh4[e] = (CW[e] - sc.x) * sc.y;
        const float og = sigmoid_acc(xh3[e] * G3[e] + B3[e]);
        const float tc = tanh_acc(CN[e]);
        d_on[e] = DH[e] * tc * og * (1.f - og);
        d_cn[e] = DH[e] * og * (1.f - tc * tc) + DC[e];
      }
      if (PHASE == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[0] += d_on[e] * G3[e]; acc[1] += d_on[e] * G3[e] * xh3[e];
          acc[2] += d_cn[e] * G4[e]; acc[3] += d_cn[e] * G4[e] * xh4[e];
          chan[0][e] += d_on[e] * xh3[e]; chan[1][e] += d_on[e];
          chan[2][e] += d_cn[e] * xh4[e]; chan[3][e] += d_cn[e];
        }
      } else {
        // ---- LN3 / LN4 backward with the sample means of phase 1, then the gate side ----
        const float* sm = a.sums + (long long)b * 10;
        float d_cw[4];
        if (PHASE == 2) {
          float d_op[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            d_op[e] = so.y * (d_on[e] * G3[e] - sm[0] - xh3[e] * sm[1]);
            d_cw[e] = sc.y * (d_cn[e] * G4[e] - sm[2] - xh4[e] * sm[3]) + d_op[e] * WCO[e];
            pco[e] += d_op[e] * CW[e];
          }
          sth4(reinterpret_cast<__half*>(a.dy16) + row * (4LL * GW) + 3 * GW + c, d_op);
          *reinterpret_cast<float4*>(a.dcnew + row * GW + c) = make_float4(d_cw[0], d_cw[1], d_cw[2], d_cw[3]);
        } else {
          const float4 v = ldf4(a.dcnew + row * GW + c);
          d_cw[0] = v.x; d_cw[1] = v.y; d_cw[2] = v.z; d_cw[3] = v.w;
        }
        const __half* yr = reinterpret_cast<const __half*>(a.y16) + row * (4LL * GW) + c;
        const float4 vj = ldh4(yr), vi = ldh4(yr + GW), vf = ldh4(yr + 2 * GW);
        const float4 vcp = a.cprev ? ldf4(a.cprev + row * GW + c) : z4;
        F4A(YJ, vj); F4A(YI, vi); F4A(YF, vf); F4A(CP, vcp);
        const float2 sj = __ldg(reinterpret_cast<const float2*>(a.mr_g) + b * 4 + 0), si = __ldg(reinterpret_cast<const float2*>(a.mr_g) + b * 4 + 1),
                     sf = __ldg(reinterpret_cast<const float2*>(a.mr_g) + b * 4 + 2);
        float d_jn[4], d_in[4], d_fn[4], xh0[4], xh1[4], xh2[4], fg[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xh0[e] = (YJ[e] - sj.x) * sj.y; xh1[e] = (YI[e] - si.x) * si.y; xh2[e] = (YF[e] - sf.x) * sf.y;
          const float jt = tanh_acc(xh0[e] * G0[e] + B0[e]);
          const float ig = sigmoid_acc(xh1[e] * G1[e] + B1[e]);
          fg[e] = sigmoid_acc(xh2[e] * G2[e] + B2[e] + 1.0f);
          d_jn[e] = d_cw[e] * ig * (1.f - jt * jt);
          d_in[e] = d_cw[e] * jt * ig * (1.f - ig);
          d_fn[e] = d_cw[e] * CP[e] * fg[e] * (1.f - fg[e]);
        }
        if (PHASE == 2) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[0] += d_jn[e] * G0[e]; acc[1] += d_jn[e] * G0[e] * xh0[e];
            acc[2] += d_in[e] * G1[e]; acc[3] += d_in[e] * G1[e] * xh1[e];
            acc[4] += d_fn[e] * G2[e]; acc[5] += d_fn[e] * G2[e] * xh2[e];
            chan[0][e] += d_jn[e] * xh0[e]; chan[1][e] += d_jn[e];
            chan[2][e] += d_in[e] * xh1[e]; chan[3][e] += d_in[e];
            chan[4][e] += d_fn[e] * xh2[e]; chan[5][e] += d_fn[e];
          }
        } else {
          float d_j[4], d_i[4], d_f[4], d_cp[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            d_j[e] = sj.y * (d_jn[e] * G0[e] - sm[4] - xh0[e] * sm[5]);
            d_i[e] = si.y * (d_in[e] * G1[e] - sm[6] - xh1[e] * sm[7]);
            d_f[e] = sf.y * (d_fn[e] * G2[e] - sm[8] - xh2[e] * sm[9]);
            d_cp[e] = d_cw[e] * fg[e] + d_i[e] * WCI[e] + d_f[e] * WCF[e];
            pci[e] += d_i[e] * CP[e];
            pcf[e] += d_f[e] * CP[e];
          }
          __half* dyr = reinterpret_cast<__half*>(a.dy16) + row * (4LL * GW) + c;
          sth4(dyr, d_j); sth4(dyr + GW, d_i); sth4(dyr + 2 * GW, d_f);
          if (a.dcprev_out) *reinterpret_cast<float4*>(a.dcprev_out + row * GW + c) = make_float4(d_cp[0], d_cp[1], d_cp[2], d_cp[3]);
        }
      }
    } else if (PHASE == 3 && pix < N && sub < rows_per_iter) {
      // padding channels of a valid row: dy and the state gradient must be exact zeros (they are GEMM operands)
      const float zz[4] = {0.f, 0.f, 0.f, 0.f};
      __half* dyr = reinterpret_cast<__half*>(a.dy16) + row * (4LL * GW) + c;
      sth4(dyr, zz); sth4(dyr + GW, zz); sth4(dyr + 2 * GW, zz); sth4(dyr + 3 * GW, zz);
      if (a.dcprev_out) *reinterpret_cast<float4*>(a.dcprev_out + row * GW + c) = z4;
    }
    if (PHASE != 3) {
      // per-sample partial of this block
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = warp_sum(acc[k]);
      if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) s_red[warp][k] = acc[k];
      __syncthreads();
      if (threadIdx.x < K) {
        float t = 0.f;
        for (int w = 0; w < LB_THREADS / 32; ++w) t += s_red[w][threadIdx.x];
        a.ws_sample[((long long)blockIdx.x * batch + b) * 6 + threadIdx.x] = t;
      }
      __syncthreads();
    }
  }
  if (pix < N && sub < rows_per_iter && c < GW) {
    if (PHASE != 3) {
#pragma unroll
      for (int q = 0; q < Q; ++q)
        *reinterpret_cast<float4*>(a.ws_chan + ((long long)pix * 6 + q) * GW + c) = make_float4(chan[q][0], chan[q][1], chan[q][2], chan[q][3]);
    }
    if (PHASE == 2 && live) {
      float4* d = reinterpret_cast<float4*>(a.dw_co + (long long)pix * GW + c);
      float4 o = *d;
      o.x += pco[0]; o.y += pco[1]; o.z += pco[2]; o.w += pco[3];
      *d = o;
    }
    if (PHASE == 3 && live && a.cprev) {
      float4* d = reinterpret_cast<float4*>(a.dw_ci + (long long)pix * GW + c);
      float4 o = *d;
      o.x += pci[0]; o.y += pci[1]; o.z += pci[2]; o.w += pci[3];
      *d = o;
      d = reinterpret_cast<float4*>(a.dw_cf + (long long)pix * GW + c);
      o = *d;
      o.x += pcf[0]; o.y += pcf[1]; o.z += pcf[2]; o.w += pcf[3];
      *d = o;
    }
  }
  (void)inv_n;
}

// folds the partials of one phase: per-sample sums -> means (sums_out[b, off + k] = sum / count);
// per-channel sums over the pixels -> accumulated into dgamma / dbeta ([5, GW] each; LN index = ln0 + q / 2)
__global__ void convlstm_bwd_reduce_kernel(const float* __restrict__ ws_sample, int nblk, int batch, int K, float inv_count,
                                           float* __restrict__ sums_out, int off, const float* __restrict__ ws_chan, int n_pix, int Q,
                                           int GW, int ln0, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.y == 0) {
    // sample sums: one warp per (b, k)
    const int w = tid >> 5, lane = tid & 31;
    if (w < batch * K) {
      const int b = w / K, k = w - b * K;
      double t = 0.0;
      for (int i = lane; i < nblk; i += 32) t += (double)ws_sample[((long long)i * batch + b) * 6 + k];
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) sums_out[b * 10 + off + k] = (float)(t * (double)inv_count);
    }
  } else {
    // channel sums: thread per (q, c), pixel range split over blockIdx.y - 1
    const int parts = gridDim.y - 1, part = blockIdx.y - 1;
    if (tid < Q * GW) {
      const int q = tid / GW, c = tid - q * GW;
      const int per = (n_pix + parts - 1) / parts;
      const int p0 = part * per, p1 = min(n_pix, p0 + per);
      float t = 0.f;
      for (int p = p0; p < p1; ++p) t += ws_chan[((long long)p * 6 + q) * GW + c];
      float* dst = (q & 1) ? dbeta : dgamma;
      atomicAdd(dst + (ln0 + q / 2) * GW + c, t);
    }
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" size_t cmpc_convlstm_bwd_workspace_floats(int32_t batch, int32_t rows_per_sample, int32_t gw) {
  const int rows_per_iter = LB_THREADS / (gw / 4) > 0 ? LB_THREADS / (gw / 4) : 1;
  const size_t nblk = (rows_per_sample + rows_per_iter - 1) / rows_per_iter;
  return (nblk * batch * 6 + 63) / 64 * 64 + (size_t)rows_per_sample * 6 * gw;     // ws_chan starts 256-byte aligned
}

extern "C" int cmpc_convlstm_bwd(int32_t phase, const cmpc_convlstm_bwd_args* a, int32_t batch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(a && phase >= 1 && phase <= 3 && batch > 0, CMPC_ERR_ARG, "cmpc_convlstm_bwd: bad args");
  CMPC_REQUIRE(a->y16 && a->opre && a->cnew && a->cn && a->mr_g && a->mr_o && a->ln_gamma && a->ln_beta && a->w_co && a->dh && a->dcnew &&
                   a->dy16 && a->ws_sample && a->ws_chan && a->dw_co && a->sums && a->dgamma && a->dbeta,
               CMPC_ERR_ARG, "cmpc_convlstm_bwd: null pointer");
  CMPC_REQUIRE(!a->cprev || (a->w_ci && a->w_cf && a->dw_ci && a->dw_cf), CMPC_ERR_ARG, "cmpc_convlstm_bwd: cprev needs the peepholes");
  CMPC_REQUIRE(a->gw % 128 == 0 && a->gw <= 1024 && a->m > 0 && a->m % 4 == 0 && a->m <= a->gw && a->rows_per_sample > 0 && a->ld_dh % 4 == 0,
               CMPC_ERR_ARG, "cmpc_convlstm_bwd: bad shape (gw 128..1024 multiple of 128, m %% 4 == 0)");
  const int gpr = a->gw / 4;
  const int rows_per_iter = LB_THREADS / gpr > 0 ? LB_THREADS / gpr : 1;
  CMPC_REQUIRE(gpr <= LB_THREADS, CMPC_ERR_ARG, "cmpc_convlstm_bwd: gw too large");
  const int nblk = (a->rows_per_sample + rows_per_iter - 1) / rows_per_iter;
  if (phase == 1) convlstm_bwd_kernel<1><<<nblk, LB_THREADS, 0, stream>>>(*a, batch, rows_per_iter);
  else if (phase == 2) convlstm_bwd_kernel<2><<<nblk, LB_THREADS, 0, stream>>>(*a, batch, rows_per_iter);
  else convlstm_bwd_kernel<3><<<nblk, LB_THREADS, 0, stream>>>(*a, batch, rows_per_iter);
  rc = check_launch("convlstm_bwd_kernel");
  if (rc || phase == 3) return rc;
  const int K = phase == 1 ? 4 : 6, Q = K;
  const int threads = 256;
  const int need = max(batch * K * 32, Q * a->gw);
  const float inv_count = 1.0f / ((float)a->rows_per_sample * (float)a->m);
  convlstm_bwd_reduce_kernel<<<dim3((need + threads - 1) / threads, 1 + 16), threads, 0, stream>>>(
      a->ws_sample, nblk, batch, K, inv_count, a->sums, phase == 1 ? 0 : 4, a->ws_chan, a->rows_per_sample, Q, a->gw, phase == 1 ? 3 : 0,
      a->dgamma, a->dbeta);
  return check_launch("convlstm_bwd_reduce_kernel");
}

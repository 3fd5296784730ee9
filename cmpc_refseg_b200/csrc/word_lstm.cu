// Word encoder in front of the head (CMPC_model.py:144-157): GloVe embedding lookup + one tf LSTMCell(rnn_size) unrolled by
// dynamic_rnn with sequence_length (outputs zero and state frozen past the end of a sentence).
//   z_t = [x_t, h_{t-1}] K + b;  i, j, f, o = split(z_t, 4);  c_t = sigmoid(f + 1) c_{t-1} + sigmoid(i) tanh(j);  h_t = sigmoid(o) tanh(c_t)
// The input half x_t K_x (+ b) of all T steps is ONE tensor-core GEMM; the recurrent half h_{t-1} K_h is a skinny GEMM per step
// (cmpc_gemm_f16, M = batch) followed by lstm_step_kernel.  (the recurrence is latency-bound: 20 dependent steps.)
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

// out16[r, :e] = emb[ids[r], :]  (fp16, pads to ld zeroed); one warp per row
__global__ void embed_gather_kernel(const int* __restrict__ ids, const float* __restrict__ emb, int vocab, int e, int rows, __half* __restrict__ out,
                                    long long ld) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  for (int c = lane; c < ld; c += 32) out[(long long)row * ld + c] = __float2half_rn(c < e ? __ldg(emb + (long long)id * e + c) : 0.f);
}

// one step: xg fp32 [B*T, 4R] (row b*T + t, input half + bias), hg fp32 [B, 4R] (recurrent half) or null at t = 0
__global__ void lstm_step_kernel(const float* __restrict__ xg, const float* __restrict__ hg, const int* __restrict__ seq_len, int t, int T, int R,
                                 int batch, float* __restrict__ c_state, __half* __restrict__ h16, long long ldh, float* __restrict__ out /*[B,T,R]*/) {
  const long long total = (long long)batch * R;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / R), k = (int)(i - (long long)b * R);
    const bool active = t < seq_len[b];
    float hv = 0.f;
    if (active) {
      const float* x = xg + ((long long)b * T + t) * 4 * R;
      const float* h = hg ? hg + (long long)b * 4 * R : nullptr;
      const float zi = x[k] + (h ? h[k] : 0.f), zj = x[R + k] + (h ? h[R + k] : 0.f);
      const float zf = x[2 * R + k] + (h ? h[2 * R + k] : 0.f), zo = x[3 * R + k] + (h ? h[3 * R + k] : 0.f);
      const float c = sigmoid_acc(zf + 1.0f) * c_state[i] + sigmoid_acc(zi) * tanh_acc(zj);
      c_state[i] = c;
      hv = sigmoid_acc(zo) * tanh_acc(c);
      h16[(long long)b * ldh + k] = __float2half_rn(hv);
    }
    out[((long long)b * T + t) * R + k] = hv;           // zero past the end of the sentence (dynamic_rnn)
  }
}

// ---- training: the same step with explicit state slots and saved gate activations, and its backward ----------------------
// Time-major rows (t * B + b) for xg / dz so that every step's slice is contiguous.  states: c fp32 [T+1, B, R], h fp16
// [T+1, B, ldh] (slot t = state entering step t, slot 0 = zeros); gates fp32 [T, B, 4R] = (sigmoid i, tanh j, sigmoid(f+1), sigmoid o).
__global__ void lstm_step_train_kernel(const float* __restrict__ xg, const float* __restrict__ hg, const int* __restrict__ seq_len, int t, int T, int R,
                                       int batch, const float* __restrict__ c_prev, float* __restrict__ c_new, const __half* __restrict__ h_prev,
                                       __half* __restrict__ h_new, long long ldh, float* __restrict__ gates, float* __restrict__ out /*[B,T,R]*/) {
  const long long total = (long long)batch * R;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / R), k = (int)(i - (long long)b * R);
    const bool active = t < seq_len[b];
    float hv = 0.f, gi = 0.f, gj = 0.f, gf = 0.f, go = 0.f;
    float c = c_prev[i];
    __half hn = h_prev[(long long)b * ldh + k];
    if (active) {
      const float* x = xg + ((long long)t * batch + b) * 4 * R;
      const float* h = hg ? hg + (long long)b * 4 * R : nullptr;
      gi = sigmoid_acc(x[k] + (h ? h[k] : 0.f));
      gj = tanh_acc(x[R + k] + (h ? h[R + k] : 0.f));
      gf = sigmoid_acc(x[2 * R + k] + (h ? h[2 * R + k] : 0.f) + 1.0f);
      go = sigmoid_acc(x[3 * R + k] + (h ? h[3 * R + k] : 0.f));
      c = gf * c + gi * gj;
      hv = go * tanh_acc(c);
      hn = __float2half_rn(hv);
    }
    c_new[i] = c;
    h_new[(long long)b * ldh + k] = hn;
    float* g = gates + (long long)b * 4 * R;
    g[k] = gi; g[R + k] = gj; g[2 * R + k] = gf; g[3 * R + k] = go;
    out[((long long)b * T + t) * R + k] = hv;
  }
}

// Backward of step t.  The final state feeds nothing and sentences are prefixes, so gradient only flows between active steps:
//   dh = d_out[b, t] + [t + 1 < len] G / S      (G = S * dz_{t+1} K_h^T from cmpc_gemm_f16),   dc = [t + 1 < len] dC + dh o (1 - tanh^2 c_t)
//   dz = (dc j i (1 - i), dc i (1 - j^2), dc c_{t-1} f (1 - f), dh tanh(c_t) o (1 - o)),   dC = dc f
// dz is written as fp16 * S (operand of the three gradient GEMMs); dbias += sum_b dz in fp32.  One thread per column k, looping
// over the batch, so the bias sum needs no atomics.
__global__ void lstm_step_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ G, long long ldg, const int* __restrict__ seq_len,
                                     int t, int T, int R, int batch, const float* __restrict__ gates, const float* __restrict__ c_prev,
                                     const float* __restrict__ c_cur, float* __restrict__ dC, float scale, __half* __restrict__ dz16, long long ldz,
                                     float* __restrict__ dbias) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= R) return;
  const float inv = 1.0f / scale;
  float sb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = 0; b < batch; ++b) {
    const int len = seq_len[b];
    float dz[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < len) {
      const bool next = t + 1 < len && t + 1 < T;
      const float* g = gates + (long long)b * 4 * R;
      const float gi = g[k], gj = g[R + k], gf = g[2 * R + k], go = g[3 * R + k];
      const long long i = (long long)b * R + k;
      const float tc = tanh_acc(c_cur[i]);
      const float dh = d_out[((long long)b * T + t) * R + k] + (next ? G[(long long)b * ldg + k] * inv : 0.f);
      const float dc = (next ? dC[i] : 0.f) + dh * go * (1.f - tc * tc);
      dz[0] = dc * gj * gi * (1.f - gi);
      dz[1] = dc * gi * (1.f - gj * gj);
      dz[2] = dc * c_prev[i] * gf * (1.f - gf);
      dz[3] = dh * tc * go * (1.f - go);
      dC[i] = dc * gf;
    }
    __half* z = dz16 + ((long long)t * batch + b) * ldz;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      z[q * R + k] = __float2half_rn(dz[q] * scale);
      sb[q] += dz[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) dbias[q * R + k] += sb[q];
}

// demb[ids[row], :] += dx[row, :e] * scale   (rows may repeat a word: atomics; tf sums duplicate IndexedSlices the same way)
__global__ void embed_scatter_add_kernel(const int* __restrict__ ids, const float* __restrict__ dx, long long ld, float scale, int vocab, int e,
                                         int rows, float* __restrict__ demb) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  for (int c = lane; c < e; c += 32) atomicAdd(demb + (long long)id * e + c, dx[(long long)row * ld + c] * scale);
}

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_lstm_step_train(const float* xg, const float* hg, const int32_t* seq_len, int32_t t, int32_t steps, int32_t r, int32_t batch,
                                    const float* c_prev, float* c_new, const void* h_prev_f16, void* h_new_f16, int64_t ldh, float* gates,
                                    float* out, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(xg && seq_len && c_prev && c_new && h_prev_f16 && h_new_f16 && gates && out && t >= 0 && t < steps && r > 0 && batch > 0 && ldh >= r,
               CMPC_ERR_ARG, "cmpc_lstm_step_train: bad args");
  const long long total = (long long)batch * r;
  lstm_step_train_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xg, hg, seq_len, t, steps, r, batch, c_prev, c_new,
                                                                                         (const __half*)h_prev_f16, (__half*)h_new_f16, ldh, gates, out);
  return check_launch("lstm_step_train_kernel");
}

extern "C" int cmpc_lstm_step_bwd(const float* d_out, const float* g_rec, int64_t ldg, const int32_t* seq_len, int32_t t, int32_t steps, int32_t r,
                                  int32_t batch, const float* gates, const float* c_prev, const float* c_cur, float* dc_state, float scale,
                                  void* dz_f16, int64_t ldz, float* dbias, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(d_out && seq_len && gates && c_prev && c_cur && dc_state && dz_f16 && dbias && t >= 0 && t < steps && r > 0 && batch > 0 &&
                   ldz >= 4 * (int64_t)r && scale > 0.f && (g_rec || t == steps - 1),
               CMPC_ERR_ARG, "cmpc_lstm_step_bwd: bad args");
  lstm_step_bwd_kernel<<<(r + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_out, g_rec, ldg, seq_len, t, steps, r, batch, gates, c_prev, c_cur, dc_state,
                                                                           scale, (__half*)dz_f16, ldz, dbias);
  return check_launch("lstm_step_bwd_kernel");
}

extern "C" int cmpc_embed_scatter_add(const int32_t* ids, const float* dx, int64_t ld, float scale, int32_t vocab, int32_t e, int32_t rows, float* demb,
                                      void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(ids && dx && demb && vocab > 0 && e > 0 && rows > 0 && ld >= e, CMPC_ERR_ARG, "cmpc_embed_scatter_add: bad args");
  embed_scatter_add_kernel<<<(rows * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ids, dx, ld, scale, vocab, e, rows, demb);
  return check_launch("embed_scatter_add_kernel");
}

extern "C" int cmpc_embed_gather_f16(const int32_t* ids, const float* emb, int32_t vocab, int32_t e, int32_t rows, void* out_f16, int64_t ld,
                                     void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(ids && emb && out_f16 && vocab > 0 && e > 0 && rows > 0 && ld >= e, CMPC_ERR_ARG, "cmpc_embed_gather_f16: bad args");
  embed_gather_kernel<<<(rows * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ids, emb, vocab, e, rows, (__half*)out_f16, ld);
  return check_launch("embed_gather_kernel");
}

extern "C" int cmpc_lstm_step(const float* xg, const float* hg, const int32_t* seq_len, int32_t t, int32_t steps, int32_t r, int32_t batch,
                              float* c_state, void* h_f16, int64_t ldh, float* out, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(xg && seq_len && c_state && h_f16 && out && t >= 0 && t < steps && r > 0 && batch > 0 && ldh >= r, CMPC_ERR_ARG,
               "cmpc_lstm_step: bad args");
  const long long total = (long long)batch * r;
  lstm_step_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xg, hg, seq_len, t, steps, r, batch, c_state, (__half*)h_f16, ldh, out);
  return check_launch("lstm_step_kernel");
}

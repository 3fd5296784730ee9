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

}  // namespace cmpc

using namespace cmpc;

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

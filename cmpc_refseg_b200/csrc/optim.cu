// Optimizer step of the reference's train_op (CMPC_model.py:446-478): tf.train.AdamOptimizer (beta1 0.9, beta2 0.999,
// eps 1e-8) on cost = cls_loss_all + weight_decay * sum_{DW} |w|^2 / 2, with the gradients of `biases` doubled (:464-475).
// One fused pass over a flat fp32 parameter group:  g = grad * grad_scale + weight_decay * w;  m, v updated in place;
// w -= lr_t * m / (sqrt(v) + eps),  lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) computed by the caller.
#include "common.cuh"

namespace cmpc {

__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v, long long n,
                            float lr_t, float beta1, float beta2, float eps, float grad_scale, float weight_decay,
                            const float* __restrict__ lr_t_dev) {
  if (lr_t_dev) lr_t = __ldg(lr_t_dev);          // step size from device memory: the launch can then live in a CUDA graph
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 ww = reinterpret_cast<float4*>(w)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(grad)[i];
    float wa[4] = {ww.x, ww.y, ww.z, ww.w}, ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
    const float ga[4] = {gg.x, gg.y, gg.z, gg.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float g = ga[e] * grad_scale + weight_decay * wa[e];
      ma[e] = beta1 * ma[e] + (1.f - beta1) * g;
      va[e] = beta2 * va[e] + (1.f - beta2) * g * g;
      wa[e] -= lr_t * ma[e] / (sqrtf(va[e]) + eps);
    }
    reinterpret_cast<float4*>(w)[i] = make_float4(wa[0], wa[1], wa[2], wa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {       // tail
    const long long i = n4 * 4 + threadIdx.x;
    const float g = grad[i] * grad_scale + weight_decay * w[i];
    m[i] = beta1 * m[i] + (1.f - beta1) * g;
    v[i] = beta2 * v[i] + (1.f - beta2) * g * g;
    w[i] -= lr_t * m[i] / (sqrtf(v[i]) + eps);
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_adam_f32(float* w, const float* grad, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2, float eps,
                             float grad_scale, float weight_decay, const float* lr_t_dev, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(w && grad && m && v && n > 0, CMPC_ERR_ARG, "cmpc_adam_f32: bad args");
  CMPC_REQUIRE(((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
               CMPC_ERR_ALIGN, "cmpc_adam_f32: buffers must be 16-byte aligned");
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, grad, m, v, n, lr_t, beta1, beta2, eps, grad_scale, weight_decay, lr_t_dev);
  return check_launch("adam_kernel");
}

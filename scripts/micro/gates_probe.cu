// Probe: why is convlstm_gates1 4x slower than its HBM traffic suggests?  Variants of the same access pattern.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x){float y; asm("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ float rcpf(float x){float y; asm("rcp.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ float sig(float x){return rcpf(1.f+ex2f(-1.4426950408889634f*x));}
__device__ __forceinline__ float tnh(float x){float e=ex2f(x*2.8853900817779268f); return 1.f-2.f*rcpf(e+1.f);}
template <int MODE>
__global__ void __launch_bounds__(256, 4) probe(const float* __restrict__ y, const float* __restrict__ cprev, const float* __restrict__ wco,
                                               float* __restrict__ cnew, float* __restrict__ opre, int rows, int rps) {
  const int GW = 512, gpr = 128;
  const int g = threadIdx.x % gpr, sub = threadIdx.x / gpr, c = g * 4;
  for (int row = blockIdx.x * 2 + sub; row < rows; row += gridDim.x * 2) {
    const int pix = row % rps;
    const float* yr = y + (long long)row * 2048 + c;
    float4 vj = __ldg((const float4*)yr), vi = __ldg((const float4*)(yr + GW)), vf = __ldg((const float4*)(yr + 2 * GW)), vo = __ldg((const float4*)(yr + 3 * GW));
    float4 cp = __ldg((const float4*)(cprev + (long long)row * GW + c));
    float4 wc = __ldg((const float4*)(wco + (long long)pix * GW + c));
    float4 cn, op;
    if (MODE == 0) {   // trivial math
      cn = make_float4(vj.x + vi.x + cp.x, vj.y + vi.y + cp.y, vj.z + vi.z + cp.z, vj.w + vi.w + cp.w);
      op = make_float4(vf.x + vo.x + wc.x, vf.y + vo.y + wc.y, vf.z + vo.z + wc.z, vf.w + vo.w + wc.w);
    } else {           // the real math
      float aj[4]={vj.x,vj.y,vj.z,vj.w}, ai[4]={vi.x,vi.y,vi.z,vi.w}, af[4]={vf.x,vf.y,vf.z,vf.w}, ao[4]={vo.x,vo.y,vo.z,vo.w}, ac[4]={cp.x,cp.y,cp.z,cp.w}, aw[4]={wc.x,wc.y,wc.z,wc.w};
      float rc[4], ro[4];
      for (int e = 0; e < 4; ++e) { rc[e] = ac[e] * sig(af[e] + 1.f) + sig(ai[e]) * tnh(aj[e]); ro[e] = ao[e] + aw[e] * rc[e]; }
      cn = make_float4(rc[0], rc[1], rc[2], rc[3]); op = make_float4(ro[0], ro[1], ro[2], ro[3]);
    }
    if (MODE == 2) { if (cn.x == 12345.f) *(float4*)(cnew + (long long)row * GW + c) = cn; }   // no stores
    else { *(float4*)(cnew + (long long)row * GW + c) = cn; *(float4*)(opre + (long long)row * GW + c) = op; }
  }
}
int main() {
  const int rows = 51200, rps = 1600;
  float *y, *cp, *wc, *cn, *op;
  cudaMalloc(&y, (size_t)rows * 2048 * 4); cudaMalloc(&cp, (size_t)rows * 512 * 4); cudaMalloc(&wc, (size_t)rps * 512 * 4);
  cudaMalloc(&cn, (size_t)rows * 512 * 4); cudaMalloc(&op, (size_t)rows * 512 * 4);
  cudaMemset(y, 0, (size_t)rows * 2048 * 4); cudaMemset(cp, 0, (size_t)rows * 512 * 4); cudaMemset(wc, 0, (size_t)rps * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {148 * 4, 148 * 16, 25600}) for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) probe<0><<<grid, 256>>>(y, cp, wc, cn, op, rows, rps);
      if (mode == 1) probe<1><<<grid, 256>>>(y, cp, wc, cn, op, rows, rps);
      if (mode == 2) probe<2><<<grid, 256>>>(y, cp, wc, cn, op, rows, rps);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("grid %5d mode %d (%s): %.1f us  (%.0f GB/s of 735 MB)\n", grid, mode, mode == 0 ? "trivial math" : mode == 1 ? "real math" : "real math, no stores", ms * 1e3, 0.735 / (ms * 1e-3));
  }
  return 0;
}

// Micro-benchmark: tcgen05.mma dispatch rate on B200 for the operand modes used by the CMPC kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu -I../../cmpc_refseg_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace cmpc;

// mode 0: SS, A K-major SW128, B K-major SW128        (GEMM kernel)
// mode 1: TS, A in TMEM, B MN-major SW128             (graph MMA2)
// mode 2: SS, A K-major, B MN-major
// mode 3: TS, A in TMEM, B K-major
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 32768;
    const uint32_t b_major = (mode == 1 || mode == 2) ? 1 : 0;
    const uint32_t idesc = make_idesc_f16(128, N, 0, 0, b_major);
    const uint64_t da = make_smem_desc(sa, 16, 1024, 2);
    const uint64_t db = b_major ? make_smem_desc(sb, 16384, 1024, 2) : make_smem_desc(sb, 16, 1024, 2);
    uint32_t ph = 0;
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int k = i & 3;
        if (mode == 0 || mode == 2) umma_f16_ss(tm, da + uint64_t(k * 2), b_major ? db + uint64_t(k * 128) : db + uint64_t(k * 2), idesc, 1);
        else umma_f16_ts(tm, tm + 256 + k * 8, b_major ? db + uint64_t(k * 128) : db + uint64_t(k * 2), idesc, 1);
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph); ph ^= 1;
      long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    out[blockIdx.x] = best;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N> void run(int grid) {
  long long* d; cudaMalloc(&d, grid * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 256;
  for (int mode = 0; mode < 4; ++mode) {
    rate_kernel<N><<<grid, 128, 100 * 1024>>>(mode, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
    printf("N=%3d grid=%3d mode %d: %6.1f cycles/dispatch (%s)  -> %.0f FLOP/clk/SM\n", N, grid, mode, (double)mx / iters, cudaGetErrorString(e),
           2.0 * 128 * N * 16 / ((double)mx / iters));
  }
  cudaFree(d);
}
int main() {
  run<256>(1); run<256>(148); run<128>(148); run<64>(148); run<240>(148);
  return 0;
}

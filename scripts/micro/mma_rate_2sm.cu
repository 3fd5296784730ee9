// Micro-benchmark: tcgen05.mma.cta_group::2 dispatch rate on B200 (M = 256 over a CTA pair), A in shared memory vs in TMEM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_2sm mma_rate_2sm.cu -I../../cmpc_refseg_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace cmpc;

// mode 0: SS, A K-major SW128, B K-major SW128;  mode 1: TS, A in TMEM, B MN-major SW128;  mode 2: SS, B MN-major;  mode 3: TS, B K-major
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc_2sm(&tptr, 512); tmem_relinquish_2sm(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 32768;
    const uint32_t b_major = (mode == 1 || mode == 2) ? 1 : 0;
    const uint32_t idesc = make_idesc_f16(256, N, 0, 0, b_major);
    const uint64_t da = make_smem_desc(sa, 16, 1024, 2);
    const uint64_t db = b_major ? make_smem_desc(sb, 16384, 1024, 2) : make_smem_desc(sb, 16, 1024, 2);
    uint32_t ph = 0;
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int k = i & 3;
        if (mode == 0 || mode == 2) umma_f16_ss_2sm(tm, da + uint64_t(k * 2), b_major ? db + uint64_t(k * 128) : db + uint64_t(k * 2), idesc, 1);
        else umma_f16_ts_2sm(tm, tm + 256 + k * 8, b_major ? db + uint64_t(k * 128) : db + uint64_t(k * 2), idesc, 1);
      }
      umma_commit_2sm_mc(&bar, 1);
      mbar_wait(&bar, ph); ph ^= 1;
      long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    out[blockIdx.x / 2] = best;
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc_2sm(tm, 512); }
}

template <int N> void run(int clusters) {
  long long* d; cudaMalloc(&d, clusters * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 256;
  for (int mode = 0; mode < 4; ++mode) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * 2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 100 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, rate_kernel<N>, mode, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, d, clusters * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < clusters; ++i) if (h[i] > mx) mx = h[i];
    printf("2SM N=%3d clusters=%3d mode %d: %6.1f cycles/dispatch (%s)  -> %.0f FLOP/clk/SM\n", N, clusters, mode, (double)mx / iters,
           cudaGetErrorString(e), 2.0 * 128 * N * 16 / ((double)mx / iters));
  }
  cudaFree(d);
}
int main() {
  run<256>(1); run<256>(74); run<128>(74);
  return 0;
}

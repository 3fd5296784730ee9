// How fast can one SM drain 64 KB of shared memory to global?  (plain 16-byte coalesced stores; rows of 128 B or 512 B,
// row pitch 2 KB like the [B*N, 1024] fp16 maps, or fully contiguous)   nvcc -arch=sm_100a -O3 store_rate.cu -o store_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint4* out, long long pitch16, int rowlen16, int iters, int nwarps_active, long long* clk) {
  const int sh = 31 - __clz(rowlen16);
  extern __shared__ uint4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_uint4(i, i, i, i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int it = 0; it < iters; ++it) {
      // 64 KB = 4096 x 16 B; rows of rowlen16 chunks
      uint4* base = out + ((long long)blockIdx.x * iters + it) * (4096 / rowlen16) * pitch16;
      for (int c = warp * 32 + lane; c < 4096; c += nwarps_active * 32) {
        const int row = c >> sh, col = c & (rowlen16 - 1);
        base[(long long)row * pitch16 + col] = sm[c];
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
  const int iters = 8;
  uint4* out; long long* clk;
  size_t bytes = (size_t)148 * iters * 4096 * 2048;   // worst case pitch
  cudaMalloc(&out, bytes); cudaMalloc(&clk, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  long long h[148];
  for (int grid : {1, 148}) for (int nw : {1, 2, 4, 8}) for (int mode = 0; mode < 3; ++mode) {
    int rowlen16 = mode == 0 ? 8 : (mode == 1 ? 32 : 4096);
    long long pitch16 = mode == 2 ? 4096 : 128;   // 2 KB pitch
    for (int rep = 0; rep < 2; ++rep) k<<<grid, 256, 65536>>>(out, pitch16, rowlen16, iters, nw, clk);
    cudaDeviceSynchronize();
    cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
    printf("grid %3d warps %d row %5d B pitch %s : %.0f clk per 64 KB  (%.1f B/clk/SM)  %s\n", grid, nw, rowlen16 * 16,
           mode == 2 ? "dense" : "2KB", s / grid / iters, 65536.0 * iters * grid / s, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

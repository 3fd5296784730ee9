// Relation-aware reasoning, phase B (CMPC_model.py:400 + :362): the dense graph aggregation
//     Y[b] = adj[b] @ X[b],   adj[b] = W[b] @ V[b]^T      (N x N, N = 1600 at 320^2, 4096 at 512^2)
// as a flash-attention-style tcgen05 kernel: the N x N adjacency exists only tile by tile in TMEM.
//
// One CTA = (sample b, 128 query nodes i, 256 output channels).  Per 128-node key tile j:
//   MMA1  S[128x128] = W_i[128x32] . V_j[128x32]^T        tcgen05.mma SS, K = 32 (T = 20 words zero-padded), fp32 in TMEM
//   CVT   P = fp16(S)                                      8 warps: tcgen05.ld.x64 -> cvt.rn.f16x2 -> tcgen05.st.x32 (P aliases S)
//   MMA2  O[128x256] += P[128x128] . X_j[128x256]          tcgen05.mma TS (A from TMEM), B = X tile MN-major in smem
// TMEM: O = columns [0,256); S/P double-buffered at [256,384) and [384,512) so CVT(j+1) overlaps MMA2(j).
//
// The key/value stream (X_j, V_j) is identical for all query tiles of a sample, and a CTA consumes it at ~57 B/clk,
// more than one SM's share of L2 bandwidth (measured: the MMA warp stalled on TMA, not on the convert warps).  So two
// CTAs with adjacent query tiles form a cluster: each loads HALF of every X_j / V_j tile and TMA-multicasts it into
// both CTAs' shared memory, halving L2 reads per SM.  A stage is recycled only when BOTH CTAs' MMAs have drained it
// (tcgen05.commit multicast onto both "empty" barriers).
//
// Out-of-range nodes (ragged last tile, N = 12.5 x 128; the padding CTA of an odd tile count) are zero-filled by the
// 3-D tensor maps on load and clipped on store.
// Epilogue: Y = O / v_scale -> fp16 staged in shared memory (128B swizzle) and written by TMA tile stores, plus the
// whole-sample layer-norm statistics (sum, sum^2) that tf.contrib.layers.layer_norm at :364 needs (fp64 atomics).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

// Optional in-kernel timeline (build with -DCMPC_GRAPH_TIMING; dbg_p is then reinterpreted as long long[grid][64]).
#ifdef CMPC_GRAPH_TIMING
#define G_TICK(slot) do { if (tl) tl[(slot)] = clock64(); } while (0)
#else
#define G_TICK(slot) do { } while (0)
#endif

constexpr int G_BM = 128;        // query nodes per CTA
constexpr int G_BJ = 128;        // key nodes per tile
constexpr int G_BC = 256;        // channels per CTA
constexpr int G_T = 32;          // padded words
constexpr int G_STAGES = 3;
constexpr int G_THREADS = 384;   // 4 control warps + 2 x 4 convert/epilogue warps
constexpr int G_CLUSTER = 2;     // CTAs sharing the key/value stream
constexpr int G_BOX_BYTES = G_BJ * 128;       // one [128 nodes x 64 ch] box, 16 KB
constexpr int G_X_BYTES = G_BJ * G_BC * 2;    // 65536
constexpr int G_V_BYTES = G_BJ * G_T * 2;     // 8192
constexpr int G_STAGE_BYTES = G_X_BYTES + G_V_BYTES;
constexpr int G_W_OFF = G_STAGES * G_STAGE_BYTES;
constexpr int G_BAR_OFF = G_W_OFF + G_BM * G_T * 2;
constexpr int G_SMEM = G_BAR_OFF + 256 + 1024;
constexpr uint32_t G_COL_O = 0, G_COL_S = 256;   // S buffer k at G_COL_S + 128 * k

struct GraphParams {
  int n_nodes, C, j_tiles;
  float inv_vscale;
  long long ldy;
  double* stats;          // [B, 2]
  float* dbg_p;           // optional [B, N, N] fp32 dump of P / v_scale (c-chunk 0 only)
};

__global__ void __launch_bounds__(G_THREADS, 1)
graph_reason_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const GraphParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + G_BAR_OFF);
  uint64_t* x_empty = x_full + G_STAGES;
  uint64_t* v_full = x_empty + G_STAGES;     // V tiles have their own ring: they are released right after MMA1, so the
  uint64_t* v_empty = v_full + G_STAGES;     // tiny S = W V^T MMA never waits behind a 64 KB X tile that is still in flight
  uint64_t* w_full = v_empty + G_STAGES;
  uint64_t* s_full = w_full + 1;     // [2]
  uint64_t* p_full = s_full + 2;     // [2]
  uint64_t* o_full = p_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cchunk = blockIdx.x, itile = blockIdx.y, b = blockIdx.z;
  const int i0 = itile * G_BM, c0 = cchunk * G_BC;
  const int J = p.j_tiles;
  const uint32_t rank = cluster_ctarank();            // cluster = (1, 2, 1): the two CTAs differ in itile only
  constexpr uint16_t kAll = (1u << G_CLUSTER) - 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], G_CLUSTER);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], G_CLUSTER);
    }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 8); }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // the peer's barriers must be initialised before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
#ifdef CMPC_GRAPH_TIMING
  long long* tl = (p.dbg_p && lane == 0) ? reinterpret_cast<long long*>(p.dbg_p) + (((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 64 : nullptr;
  if (warp == 1) G_TICK(0);
#endif

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, G_BM * G_T * 2);
      tma_load_3d(smem + G_W_OFF, &tmW, w_full, 0, i0, b);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < J; ++j) {
        uint8_t* sx = smem + s * G_STAGE_BYTES;
        // V: rows [64*rank, 64*rank + 64) of the key tile;  X: channel boxes 2*rank, 2*rank + 1
        mbar_wait(&v_empty[s], ph ^ 1);                       // both CTAs' MMA1 have read this V slot
        mbar_expect_tx(&v_full[s], G_V_BYTES);                // my half + the peer's half
        tma_load_3d_mc(sx + G_X_BYTES + rank * (G_V_BYTES / 2), &tmV, &v_full[s], 0, j * G_BJ + rank * (G_BJ / 2), b, kAll);
        mbar_wait(&x_empty[s], ph ^ 1);                       // both CTAs' MMA2 have drained this X stage
        mbar_expect_tx(&x_full[s], G_X_BYTES);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const int box = rank * 2 + m;
          tma_load_3d_mc(sx + box * G_BOX_BYTES, &tmX, &x_full[s], c0 + box * 64, j * G_BJ, b, kAll);
        }
        if (++s == G_STAGES) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Measured on B200 (scripts/micro/mma_rate.cu): a cta_group::1 dispatch costs >= ~105 clk whatever its N, 138.6 clk at
    // N = 256 with A in TMEM, 172 clk at N = 256 with A in smem.  So S is computed per full 128-key tile (2 dispatches)
    // rather than in finer pieces, and MMA2 keeps N = 256.
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_f16(G_BM, G_BJ, 0, 0, 0);   // S = W V^T, both K-major
      constexpr uint32_t idesc2 = make_idesc_f16(G_BM, G_BC, 0, 0, 1);   // O += P X, B (X) MN-major
      const uint32_t w_addr = smem_u32(smem + G_W_OFF);
      auto issue_mma1 = [&](int j) {
        const int st = j % G_STAGES;
        mbar_wait(&v_full[st], (uint32_t)((j / G_STAGES) & 1));
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + st * G_STAGE_BYTES + G_X_BYTES);
        const uint64_t dw = make_smem_desc(w_addr, 16, 512, 4);   // 64-byte swizzle, 8 rows x 64 B atoms
        const uint64_t dv = make_smem_desc(v_addr, 16, 512, 4);
        const uint32_t d = tmem_base + G_COL_S + 128 * (j & 1);
#pragma unroll
        for (int k = 0; k < G_T / 16; ++k) umma_f16_ss(d, dw + uint64_t(k * 2), dv + uint64_t(k * 2), idesc1, k != 0 ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
        if (j + G_STAGES < J) umma_commit_mc(&v_empty[st], kAll);   // V slot free in BOTH CTAs once their MMA1s have read it
      };
      mbar_wait(w_full, 0);
      G_TICK(1);
      tc_fence_after();
      issue_mma1(0);
      if (J > 1) issue_mma1(1);
      for (int j = 0; j < J; ++j) {
        const int st = j % G_STAGES;
        G_TICK(2 + 2 * j);
        mbar_wait(&x_full[st], (uint32_t)((j / G_STAGES) & 1));
        mbar_wait(&p_full[j & 1], (uint32_t)((j >> 1) & 1));
        G_TICK(3 + 2 * j);
        tc_fence_after();
        const uint32_t x_addr = smem_u32(smem + st * G_STAGE_BYTES);
        // MN-major, 128B swizzle: LBO = distance between 64-channel boxes, SBO = 8 key rows
        const uint64_t dx = make_smem_desc(x_addr, G_BOX_BYTES, 1024, 2);
        const uint32_t a_tmem = tmem_base + G_COL_S + 128 * (j & 1);
#pragma unroll
        for (int k = 0; k < G_BJ / 16; ++k)
          umma_f16_ts(tmem_base + G_COL_O, a_tmem + (k >> 2) * 64 + (k & 3) * 8, dx + uint64_t((k * 16 * 128) >> 4), idesc2,
                      (j | k) != 0 ? 1u : 0u);   // P half h (keys 64h..64h+63) lives at S-buffer columns [64h, 64h+32)
        // frees the stage in BOTH CTAs once these MMAs have read it; the last G_STAGES tiles are never refilled, and not
        // signalling them means no CTA touches its peer's barriers after the peer's own last wait -> either may exit first
        if (j + G_STAGES < J) umma_commit_mc(&x_empty[st], kAll);
        if (j + 2 < J) issue_mma1(j + 2);        // overwrites S/P buffer (j & 1): ordered after MMA2(j) by in-order MMA issue
      }
      umma_commit(o_full);
      G_TICK(40);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== convert (S -> P) and epilogue: two warpgroups, each owns half of the columns ==========
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;             // 0: keys / channels [0,64) / [0,128);  1: the upper half
    const int row = q * 32 + lane;
    const int i = i0 + row;                       // node inside the sample
    const bool row_ok = i < p.n_nodes;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    for (int j = 0; j < J; ++j) {
      mbar_wait(&s_full[j & 1], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint32_t sbuf = tmem_base + lane_off + G_COL_S + 128 * (j & 1) + half * 64;
      uint32_t r[64];
      tmem_ld_x64(sbuf, r);
      tmem_wait_ld();
#ifndef CMPC_GRAPH_TIMING
      if (p.dbg_p != nullptr && cchunk == 0 && row_ok) {
        float* d = p.dbg_p + ((long long)b * p.n_nodes + i) * p.n_nodes + j * G_BJ + half * 64;
        for (int e = 0; e < 64; ++e)
          if (j * G_BJ + half * 64 + e < p.n_nodes) d[e] = __uint_as_float(r[e]) * p.inv_vscale;
      }
#endif
      uint32_t pk[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const __half2 hh = __floats2half2_rn(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
        pk[e] = *reinterpret_cast<const uint32_t*>(&hh);
      }
      tmem_st_x32(sbuf, pk);      // overwrites only columns this thread has just read
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }
    // ---- epilogue: this warp owns channels [c0 + 128*half, +128) of its 32 rows = output boxes 2*half, 2*half+1 ----
    if (warp == 4) G_TICK(41);
    mbar_wait(o_full, 0);      // all MMAs done => every load has landed and every stage buffer is free for staging
    if (warp == 4) G_TICK(42);
    tc_fence_after();
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int bxl = 0; bxl < 2; ++bxl) {               // this warpgroup's two output boxes of 64 channels
      const int bx = half * 2 + bxl;
      const int col = bx * 64;                        // column inside the CTA's 256
      const int cb = c0 + col;
      uint32_t r[64];
      tmem_ld_x64(tmem_base + lane_off + G_COL_O + col, r);
      tmem_wait_ld();
      // stage as fp16 in the 128B-swizzled layout the TMA store expects: 16-byte chunk k of a 128-byte row at k ^ (row & 7)
      uint8_t* box = smem + bx * G_BOX_BYTES + row * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[e] = (cb + g * 8 + e < p.C) ? __uint_as_float(r[g * 8 + e]) * p.inv_vscale : 0.f;
          s1 += v[e];
          s2 += v[e] * v[e];
        }
        uint4 u;
        __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
        __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(box + ((g ^ (row & 7)) << 4)) = u;
      }
      fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the TMA (async proxy)
      named_bar_sync(1 + half, 128);       // the 4 warps of this warpgroup
      if (q == 0 && lane == 0 && cb < p.ldy) {
        tma_store_3d(&tmY, smem + bx * G_BOX_BYTES, cb, i0, b);   // rows >= N are clipped; overlaps the next box's drain
        tma_store_commit();
      }
    }
    if (warp == 4) G_TICK(49);
    if (q == 0 && lane == 0) tma_store_wait_read();   // smem may go once the TMA has read it; global writes finish on their own
    if (warp == 4) G_TICK(51);
    if (p.stats) {
      if (!row_ok) { s1 = 0.f; s2 = 0.f; }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        atomicAdd(p.stats + 2 * b, (double)s1);
        atomicAdd(p.stats + 2 * b + 1, (double)s2);
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) G_TICK(52);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
    G_TICK(43);
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_graph_reason_f16(const void* w_f16, const void* v_f16, const void* x_f16, int64_t ldx, int32_t batch,
                                     int32_t n_nodes, int32_t c, float v_scale, void* y_f16, int64_t ldy, double* stats,
                                     float* dbg_p, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(w_f16 && v_f16 && x_f16 && y_f16 && batch > 0 && n_nodes > 0 && c > 0, CMPC_ERR_ARG, "cmpc_graph_reason_f16: bad args");
  CMPC_REQUIRE(c % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= c && ldy >= c, CMPC_ERR_ARG,
               "cmpc_graph_reason_f16: c, ldx, ldy must be multiples of 8 with ld >= c");
  CMPC_REQUIRE(v_scale > 0.f, CMPC_ERR_ARG, "cmpc_graph_reason_f16: v_scale must be positive");
  CUtensorMap tW, tV, tX, tY;
  rc = make_tmap_3d_sw(&tW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w_f16, G_T, n_nodes, batch, G_T * 2, (uint64_t)n_nodes * G_T * 2, G_T,
                       G_BM, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  rc = make_tmap_3d_sw(&tV, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, v_f16, G_T, n_nodes, batch, G_T * 2, (uint64_t)n_nodes * G_T * 2, G_T,
                       G_BJ / G_CLUSTER, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  rc = make_tmap_3d_sw(&tX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x_f16, c, n_nodes, batch, ldx * 2, (uint64_t)n_nodes * ldx * 2, 64, G_BJ,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_3d_sw(&tY, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y_f16, ldy, n_nodes, batch, ldy * 2, (uint64_t)n_nodes * ldy * 2, 64, G_BM,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(graph_reason_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "cudaFuncSetAttribute(graph, smem=%d): %s", G_SMEM, cudaGetErrorString(e));
    configured = true;
  }
  GraphParams p{};
  p.n_nodes = n_nodes; p.C = c; p.j_tiles = (n_nodes + G_BJ - 1) / G_BJ;
  p.inv_vscale = 1.0f / v_scale;
  p.ldy = ldy; p.stats = stats; p.dbg_p = dbg_p;
  const int itiles = (n_nodes + G_BM - 1) / G_BM;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((c + G_BC - 1) / G_BC, (itiles + G_CLUSTER - 1) / G_CLUSTER * G_CLUSTER, batch);   // padded to whole clusters
  cfg.blockDim = dim3(G_THREADS);
  cfg.dynamicSmemBytes = G_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = G_CLUSTER; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, graph_reason_kernel, tW, tV, tX, tY, p);
  CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "graph_reason_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("graph_reason_kernel");
}

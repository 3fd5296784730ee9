// Relation-aware reasoning, phase B (CMPC_model.py:400 + :362): the dense graph aggregation
//     Y[b] = adj[b] @ X[b],   adj[b] = W[b] @ V[b]^T      (N x N, N = 1600 at 320^2, 4096 at 512^2)
// as a flash-attention-style tcgen05 kernel: the N x N adjacency exists only tile by tile in TMEM.
//
// One work unit of a CTA = (sample b, 128 query nodes i, 256 output channels).  Per 128-node key tile j:
//   MMA1  S[128x128] = W_i[128x32] . V_j[128x32]^T        tcgen05.mma SS, K = 32 (T = 20 words zero-padded), fp32 in TMEM
//   CVT   P = fp16(S)                                      8 warps: tcgen05.ld.x64 -> cvt.rn.f16x2 -> tcgen05.st.x32 (P aliases S)
//   MMA2  O[128x256] += P[128x128] . X_j[128x256]          tcgen05.mma TS (A from TMEM), B = X tile MN-major in smem
// TMEM: O = columns [0,256); S/P double-buffered at [256,384) and [384,512) so CVT(j+1) overlaps MMA2(j).
//
// The key/value stream (X_j, V_j) is identical for all query tiles of a sample, and a CTA consumes it at ~57 B/clk,
// more than one SM's share of L2 bandwidth (measured: the MMA warp stalled on TMA, not on the convert warps).  So two
// CTAs form a cluster that shares the stream: each loads HALF of every X_j / V_j tile and TMA-multicasts it into both
// CTAs' shared memory.  A stage is recycled only when BOTH CTAs' MMAs have drained it (tcgen05.commit multicast onto
// both "empty" barriers).
//
// PERSISTENT: one cluster per SM pair walks a static list of units, and every pipeline (TMA rings, S/P buffers) runs
// straight through the unit boundaries, so the loads, MMA1s and the first converts of unit u+1 overlap the epilogue of
// unit u.  (The non-persistent version spent 2.8 k clk in its prologue and 4.3 k clk in its epilogue per 18 k clk of
// MMAs, with one CTA per SM -- scripts/graph_timeline.py.)  Units of a sample:
//   * "pair" units: query tiles (2p, 2p+1) x one 256-channel chunk -- X boxes and V halves multicast as above;
//   * if the number of query tiles is odd (N = 1600: 12.5 -> 13), the last tile is paired across CHANNEL chunks
//     (2k, 2k+1) instead of being padded with an empty query tile: V is still shared, each CTA loads its own X columns.
// The last key tile issues only as many K = 16 MMA2 steps as it has valid keys (N = 1600: 4 of 8).
//
// Out-of-range nodes / channels are zero-filled by the 3-D tensor maps on load and clipped on store.
// Epilogue: Y = O / v_scale -> fp16 staged in shared memory (128B swizzle) in the ring stage that is refilled last, and
// written from there by TMA tile stores that two otherwise idle warps issue, plus the whole-sample layer-norm statistics
// (sum, sum^2) that tf.contrib.layers.layer_norm at :364 needs (fp64 atomics).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

// Optional in-kernel timeline of each CTA's FIRST unit (build with -DCMPC_GRAPH_TIMING; dbg_p is then reinterpreted as
// long long[grid][64]).
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
constexpr int G_THREADS = 384;   // TMA, MMA, 2 store warps | 2 x 4 convert/epilogue warps
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
  int n_nodes, C, j_tiles, i_tiles, c_chunks, batch;
  int units_per_sample, pair_units, total_units;
  int last_ksteps;        // K = 16 steps of MMA2 that hold valid keys in the last key tile
  float inv_vscale;
  long long ldy;
  double* stats;          // [B, 2]
  float* dbg_p;           // optional [B, N, N] fp32 dump of P / v_scale (c-chunk 0 only)
  int p16;                // MMA1 accumulates S in fp16: the S -> P step is a packed TMEM load + store without a conversion
  int ring_skip;          // the stage that stages a unit's output sits out one turn of the ring (see GRing)
  int epi_order;          // 1: first output box -> convert the next unit's tile 0 -> second box; 0: convert first
};

// Order in which the key tiles of consecutive units take the three ring stages.  Plain round robin makes key tile 2 of unit
// u + 1 wait for the stage that is still draining unit u's output (64 KB through the TMA while the SM's L2 port feeds the X
// stream: 4-5 k clk, released ~6.7 k clk after the unit boundary) although the stage of ITS key tile 0 is free two k clk after
// the boundary: a 4.4 k clk hole in every 26 k clk unit at N = 1600 (profiles/r02_graph_timeline.txt).  With ring_skip the staging
// stage z of a unit is left out ONCE: tiles 0, 1, 2, 3 of the next unit go to z+1, z+2, z+1, z+2 and z rejoins at tile 4, by which
// time its output has left.  Every role (producer, MMA1 / MMA2 issue, epilogue, store warps) steps its own copy of this
// counter in tile order, so all agree on (stage, phase) without communicating; both CTAs of a cluster run the same sequence.
struct GRing {
  uint32_t u0 = 0, u1 = 0, u2 = 0;       // uses of each stage so far (registers: no dynamically indexed array in the MMA thread)
  int cur = 0, skip = -1;
  __device__ __forceinline__ void next(int& st, uint32_t& ph) {
    static_assert(G_STAGES == 3, "GRing is written for three stages");
    int t = cur;
    if (t == skip) { t = (t + 1 == G_STAGES) ? 0 : t + 1; skip = -1; }
    cur = (t + 1 == G_STAGES) ? 0 : t + 1;
    st = t;
    ph = (t == 0 ? u0 : (t == 1 ? u1 : u2)) & 1u;
    u0 += (t == 0); u1 += (t == 1); u2 += (t == 2);
  }
  __device__ __forceinline__ void end_unit(int z, bool enable) { if (enable) skip = z; }
  // all J tiles of one unit at once: returns the stage of its last tile (= where the unit's output is staged)
  __device__ __forceinline__ int advance_unit(int J, bool enable) {
    int st = 0; uint32_t ph;
    for (int j = 0; j < J; ++j) next(st, ph);
    end_unit(st, enable);
    return st;
  }
};

struct GraphUnit { int b, i0, c0, chunk; bool own_x; };

__device__ __forceinline__ GraphUnit graph_unit(const GraphParams& p, int u, uint32_t rank) {
  GraphUnit g;
  g.b = u / p.units_per_sample;
  const int r = u - g.b * p.units_per_sample;
  if (r < p.pair_units) {                 // query tiles (2p, 2p + 1) of one channel chunk
    const int pair = r / p.c_chunks;
    g.chunk = r - pair * p.c_chunks;
    g.i0 = (2 * pair + (int)rank) * G_BM;
    g.own_x = false;
  } else {                                // odd last query tile: channel chunks (2k, 2k + 1)
    g.chunk = 2 * (r - p.pair_units) + (int)rank;
    g.i0 = (p.i_tiles - 1) * G_BM;
    g.own_x = true;
  }
  g.c0 = g.chunk * G_BC;
  return g;
}

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
  uint64_t* w_empty = w_full + 1;
  uint64_t* s_full = w_empty + 1;    // [2]
  uint64_t* p_full = s_full + 2;     // [2]
  uint64_t* o_full = p_full + 2;
  uint64_t* o_empty = o_full + 1;
  uint64_t* stage_free = o_empty + 1;   // the stage the epilogue staged its output in has been read by the TMA stores (both CTAs)
  uint64_t* staged = stage_free + 1;    // [4] output box bx is complete in shared memory (4 warps each)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(staged + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = p.j_tiles;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / G_CLUSTER, n_clusters = gridDim.x / G_CLUSTER;
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
    mbar_init(w_full, 1); mbar_init(w_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 8); }
    mbar_init(o_full, 1); mbar_init(o_empty, 8);
    mbar_init(stage_free, 2 * G_CLUSTER);
    for (int s = 0; s < 4; ++s) mbar_init(&staged[s], 4);
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
  long long* tl0 = (p.dbg_p && lane == 0) ? reinterpret_cast<long long*>(p.dbg_p) + (long long)blockIdx.x * 64 : nullptr;
  long long* tl = tl0;
#endif

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      GRing ring;                          // (stage, phase) of every key tile, in issue order
      const bool skip_on = p.ring_skip != 0;
      int need[G_STAGES] = {-1, -1, -1};   // last unit whose output is staged in this stage (its drain gates the refill)
      int ui = 0, staged = 0;              // units started; units whose output staging has been waited for
      // De-synchronise the clusters.  All units cost the same, so without this every SM reaches its epilogue in the same
      // few hundred cycles and 148 x 64 KB of output hit L2/HBM at once: the TMA stores then take 5.4 k clk instead of
      // ~1.4 k and hold the staging stage that long (measured, scripts/graph_timeline.py).  The clusters that get one
      // unit fewer than the others have a whole unit of slack: spread their start times over it.
      {
        const int rem = p.total_units % n_clusters;             // clusters [0, rem) run one unit more than the rest
        if (rem != 0 && cluster_id >= rem) {
          const long long unit_clk = (long long)J * 1400 + 3000;
          const long long delay = unit_clk * 9 / 10 * (cluster_id - rem + 1) / (n_clusters - rem + 1);
          const long long t0 = clock64();
          while (clock64() - t0 < delay) { }
        }
      }
      // Issue order inside a unit: V(0), V(1), X(0), W, V(2), X(1), V(3), X(2), ...  The 8 KB V tiles run two key tiles ahead of
      // the 64 KB X tiles (MMA1 is issued two tiles ahead of MMA2) and never queue behind one: with V(1) issued after X(0) the MMA
      // thread sat ~2 k clk at every unit boundary waiting for it before it even looked at tile 0.  W is requested after X(0): it
      // has to wait for the previous unit's last MMA1, the stage of X(0) is free 1.3 k clk earlier than that.
      GRing ringv;                         // stage sequence of the V stream (same sequence, two tiles ahead)
      for (int u = cluster_id; u < p.total_units; u += n_clusters, ++ui) {
        const GraphUnit un = graph_unit(p, u, rank);
        auto load_v = [&](int j) {
          int s; uint32_t ph;
          ringv.next(s, ph);
          if (j == J - 1) ringv.end_unit(s, skip_on);
          // V: rows [64*rank, 64*rank + 64) of the key tile
          mbar_wait(&v_empty[s], ph ^ 1);                       // both CTAs' MMA1 have read this V slot
          mbar_expect_tx(&v_full[s], G_V_BYTES);                // my half + the peer's half
          tma_load_3d_mc(smem + s * G_STAGE_BYTES + G_X_BYTES + rank * (G_V_BYTES / 2), &tmV, &v_full[s], 0, j * G_BJ + rank * (G_BJ / 2), un.b, kAll);
        };
        load_v(0);
        if (J > 1) load_v(1);
        for (int j = 0; j < J; ++j) {
          int s; uint32_t ph;
          ring.next(s, ph);
          uint8_t* sx = smem + s * G_STAGE_BYTES;
          mbar_wait(&x_empty[s], ph ^ 1);                       // both CTAs' MMA2 have drained this X stage
          // the epilogue of a unit stages its output in the stage of the unit's last key tile: its drain gates the refill
          while (staged <= need[s]) {
            mbar_wait(stage_free, (uint32_t)(staged & 1));
            ++staged;
            if (ui == 1) G_TICK(55);
          }
          if (ui == 1 && j == 0) G_TICK(56);
          if (ui == 1 && j == 2) G_TICK(57);
          if (j == J - 1) { need[s] = ui; ring.end_unit(s, skip_on); }
          mbar_expect_tx(&x_full[s], G_X_BYTES);
          if (!un.own_x) {                                      // channel boxes 2*rank, 2*rank + 1, multicast to both CTAs
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              const int box = rank * 2 + m;
              tma_load_3d_mc(sx + box * G_BOX_BYTES, &tmX, &x_full[s], un.c0 + box * 64, j * G_BJ, un.b, kAll);
            }
          } else {                                              // the peer works on other channels: all four boxes are mine
#pragma unroll
            for (int box = 0; box < 4; ++box)
              tma_load_3d(sx + box * G_BOX_BYTES, &tmX, &x_full[s], un.c0 + box * 64, j * G_BJ, un.b);
          }
          if (j == 0) {                                         // (before V(2): with the skip order V(2) reuses the slot of V(0),
            if (ui > 0) mbar_wait(w_empty, (uint32_t)((ui - 1) & 1));      //  whose MMA1 needs this W)
            mbar_expect_tx(w_full, G_BM * G_T * 2);             // every MMA1 of the previous unit has read W
            tma_load_3d(smem + G_W_OFF, &tmW, w_full, 0, un.i0, un.b);
          }
          if (j + 2 < J) load_v(j + 2);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Measured on B200 (scripts/micro/mma_rate.cu): a cta_group::1 dispatch costs >= ~105 clk whatever its N, 138.6 clk at
    // N = 256 with A in TMEM, 172 clk at N = 256 with A in smem.  So S is computed per full 128-key tile (2 dispatches)
    // rather than in finer pieces, and MMA2 keeps N = 256.
    if (lane == 0) {
      // S = W V^T, both K-major; with p16 the accumulator is fp16 (K = 32 products of O(1) fp16 operands: one rounding per K = 16 step)
      const uint32_t idesc1 = p.p16 ? make_idesc_f16(G_BM, G_BJ, 0, 0, 0, 0) : make_idesc_f16(G_BM, G_BJ, 0, 0, 0);
      constexpr uint32_t idesc2 = make_idesc_f16(G_BM, G_BC, 0, 0, 1);   // O += P X, B (X) MN-major
      const uint32_t w_addr = smem_u32(smem + G_W_OFF);
      uint32_t g0 = 0;
      int ui = 0;
      GRing ring1, ring2;                  // MMA1 runs two key tiles ahead of MMA2: one copy of the stage sequence each
      const bool skip_on = p.ring_skip != 0;
      for (int u = cluster_id; u < p.total_units; u += n_clusters, ++ui, g0 += (uint32_t)J) {
        auto issue_mma1 = [&](int j) {
          const uint32_t g = g0 + (uint32_t)j;
          int st; uint32_t sph;
          ring1.next(st, sph);
          if (j == J - 1) ring1.end_unit(st, skip_on);
          mbar_wait(&v_full[st], sph);
          if (ui == 1 && j < 2) G_TICK(59 + 2 * j);
          tc_fence_after();
          const uint32_t v_addr = smem_u32(smem + st * G_STAGE_BYTES + G_X_BYTES);
          const uint64_t dw = make_smem_desc(w_addr, 16, 512, 4);   // 64-byte swizzle, 8 rows x 64 B atoms
          const uint64_t dv = make_smem_desc(v_addr, 16, 512, 4);
          const uint32_t d = tmem_base + G_COL_S + 128 * (g & 1);
#pragma unroll
          for (int k = 0; k < G_T / 16; ++k) umma_f16_ss(d, dw + uint64_t(k * 2), dv + uint64_t(k * 2), idesc1, k != 0 ? 1u : 0u);
          umma_commit(&s_full[g & 1]);
          umma_commit_mc(&v_empty[st], kAll);        // V slot free in BOTH CTAs once their MMA1s have read it
          if (j == J - 1) umma_commit(w_empty);      // last MMA1 of the unit: W may be overwritten
          if (ui == 1 && j < 2) G_TICK(60 + 2 * j);
        };
        mbar_wait(w_full, (uint32_t)(ui & 1));
        if (ui == 1) G_TICK(1);
        tc_fence_after();
        issue_mma1(0);
        if (J > 1) issue_mma1(1);
        for (int j = 0; j < J; ++j) {
          const uint32_t g = g0 + (uint32_t)j;
          int st; uint32_t xph;
          ring2.next(st, xph);
          if (j == J - 1) ring2.end_unit(st, skip_on);
          if (ui == 1) G_TICK(2 + 2 * j);
          mbar_wait(&x_full[st], xph);
          if (ui == 1 && j == 6) G_TICK(58);
          mbar_wait(&p_full[g & 1], (g >> 1) & 1);
          if (j == 0 && ui > 0) mbar_wait(o_empty, (uint32_t)((ui - 1) & 1));    // the previous unit's O has left TMEM
          if (ui == 1) G_TICK(3 + 2 * j);
          tc_fence_after();
          const uint32_t x_addr = smem_u32(smem + st * G_STAGE_BYTES);
          // MN-major, 128B swizzle: LBO = distance between 64-channel boxes, SBO = 8 key rows
          const uint64_t dx = make_smem_desc(x_addr, G_BOX_BYTES, 1024, 2);
          const uint32_t a_tmem = tmem_base + G_COL_S + 128 * (g & 1);
          const int ksteps = (j == J - 1) ? p.last_ksteps : G_BJ / 16;     // keys past N are zero: skip their K steps
#pragma unroll
          for (int k = 0; k < G_BJ / 16; ++k)
            if (k < ksteps)
              umma_f16_ts(tmem_base + G_COL_O, a_tmem + (k >> 2) * 64 + (k & 3) * 8, dx + uint64_t((k * 16 * 128) >> 4), idesc2,
                          (j | k) != 0 ? 1u : 0u);   // P half h (keys 64h..64h+63) lives at S-buffer columns [64h, 64h+32)
          umma_commit_mc(&x_empty[st], kAll);        // frees the stage in BOTH CTAs once these MMAs have read it
          if (j + 2 < J) issue_mma1(j + 2);          // overwrites S/P buffer (g & 1): ordered after MMA2(j) by in-order MMA issue
        }
        umma_commit(o_full);
        if (ui == 0) G_TICK(0);
        if (ui == 1) G_TICK(40);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== convert (S -> P) and epilogue: two warpgroups, each owns half of the columns ==========
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;             // 0: keys / channels [0,64) / [0,128);  1: the upper half
    const int row = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    uint32_t g0 = 0;
    int ui = 0;
    GRing ring;
    for (int u = cluster_id; u < p.total_units; u += n_clusters, ++ui, g0 += (uint32_t)J) {
      const GraphUnit un = graph_unit(p, u, rank);
      const int zstage = ring.advance_unit(J, p.ring_skip != 0);     // stage of this unit's last key tile
      const int i = un.i0 + row;                  // node inside the sample
      const bool row_ok = i < p.n_nodes;
      // S -> P for key tile j of unit cu (global tile counter g)
      auto convert = [&](uint32_t g, const GraphUnit& cu, int j) {
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t sbuf = tmem_base + lane_off + G_COL_S + 128 * (g & 1) + half * 64;
        if (p.p16) {
          // fp16 accumulators sit one per 32-bit column: the packed load IS the A-operand layout of MMA2 (two keys per column)
          uint32_t pk16[32];
          tmem_ld_x32_pack16(sbuf, pk16);
          tmem_wait_ld();
#ifndef CMPC_GRAPH_TIMING
          if (p.dbg_p != nullptr && cu.chunk == 0 && cu.i0 + row < p.n_nodes) {
            float* d = p.dbg_p + ((long long)cu.b * p.n_nodes + cu.i0 + row) * p.n_nodes + j * G_BJ + half * 64;
            for (int e = 0; e < 32; ++e) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pk16[e]));
              if (j * G_BJ + half * 64 + 2 * e < p.n_nodes) d[2 * e] = f.x * p.inv_vscale;
              if (j * G_BJ + half * 64 + 2 * e + 1 < p.n_nodes) d[2 * e + 1] = f.y * p.inv_vscale;
            }
          }
#endif
          tmem_st_x32(sbuf, pk16);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[g & 1]);
          return;
        }
        uint32_t r[64];
        tmem_ld_x64(sbuf, r);
        tmem_wait_ld();
#ifndef CMPC_GRAPH_TIMING
        if (p.dbg_p != nullptr && cu.chunk == 0 && cu.i0 + row < p.n_nodes) {
          float* d = p.dbg_p + ((long long)cu.b * p.n_nodes + cu.i0 + row) * p.n_nodes + j * G_BJ + half * 64;
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
        if (lane == 0) mbar_arrive(&p_full[g & 1]);
      };
      // tile 0 of every unit but the first was converted ahead of the previous unit's epilogue (below)
      for (int j = (ui > 0) ? 1 : 0; j < J; ++j) {
        convert(g0 + (uint32_t)j, un, j);
        if (warp == 4 && ui == 1 && j < 2) G_TICK(53 + j);
      }
      // ---- epilogue: this warp owns channels [c0 + 128*half, +128) of its 32 rows = output boxes 2*half, 2*half+1 ----
      if (warp == 4 && ui == 0) G_TICK(41);
      mbar_wait(o_full, (uint32_t)(ui & 1));      // all MMAs of the unit done
      if (warp == 4 && ui == 0) G_TICK(42);
      tc_fence_after();
      // The next unit's first MMA2 needs P(0) AND the accumulator drained.  S(0) of the next unit lands ~600 clk after o_full (its
      // MMA1 waits for the next W tile), so the first output box is drained into shared memory in that gap, THEN tile 0 is
      // converted, then the second box leaves TMEM (timeline: MMA2(0) of the next unit started 2.9 k clk after o_full when the
      // convert came first and both boxes after it).  epi_order 0 keeps the old order for A/B runs.
      auto convert_next0 = [&]() {
        if (u + n_clusters < p.total_units) {
          const GraphUnit nx = graph_unit(p, u + n_clusters, rank);
          convert(g0 + (uint32_t)J, nx, 0);
          if (warp == 4 && ui == 0) G_TICK(53);
        }
      };
      if (p.epi_order == 0) convert_next0();
      // output staging: the X area of the ring stage of this unit's last key tile (consumed: o_full); the producer waits for
      // stage_free before it refills that stage (with ring_skip two key tiles later than round robin would)
      uint8_t* stg = smem + zstage * G_STAGE_BYTES;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int bxl = 0; bxl < 2; ++bxl) {               // this warpgroup's two output boxes of 64 channels
        const int bx = half * 2 + bxl;
        const int col = bx * 64;                        // column inside the CTA's 256
        const int cb = un.c0 + col;
        if (bxl == 1 && p.epi_order != 0) convert_next0();
        uint32_t r[64];
        tmem_ld_x64(tmem_base + lane_off + G_COL_O + col, r);
        tmem_wait_ld();
        if (bxl == 1) {                                 // O has left TMEM: the next unit's MMA2 may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty);
          if (warp == 4 && ui == 0) G_TICK(46);
        }
        // stage as fp16 in the 128B-swizzled layout the TMA store expects: 16-byte chunk k of a 128-byte row at k ^ (row & 7)
        uint8_t* box = stg + bx * G_BOX_BYTES + row * 128;
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v[e] = __uint_as_float(r[gq * 8 + e]) * p.inv_vscale;      // channels >= C are exact zeros (X is zero-filled there)
            s1 += v[e];
            s2 += v[e] * v[e];
          }
          uint4 uu;
          __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
          __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
          uu.x = *reinterpret_cast<uint32_t*>(&h0);
          uu.y = *reinterpret_cast<uint32_t*>(&h1);
          uu.z = *reinterpret_cast<uint32_t*>(&h2);
          uu.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(box + ((gq ^ (row & 7)) << 4)) = uu;
        }
        fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&staged[bx]);    // a store warp issues the TMA stores: the drain never blocks a convert warp
      }
      if (warp == 4 && ui == 0) G_TICK(49);
      if (p.stats) {
        if (!row_ok) { s1 = 0.f; s2 = 0.f; }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0 && un.c0 < p.C) {
          atomicAdd(p.stats + 2 * un.b, (double)s1);
          atomicAdd(p.stats + 2 * un.b + 1, (double)s2);
        }
      }
      if (warp == 4 && ui == 0) G_TICK(52);
    }
    tc_fence_before();
  } else {
    // ===================== store warps (2, 3): TMA stores of the staged output boxes of column half (warp - 2) ==========
    // Draining the 64 KB of a unit takes 4-5 k clk whatever issues it -- the L2 is busy feeding the X stream (measured here:
    // TMA tile stores 4.1 k clk, plain 16-byte stores from two / four warps 6.9 k / 5.1 k; an otherwise idle SM drains 64 KB
    // in 2.1 k clk, scripts/micro/store_rate.cu) -- so it must not sit on a convert warp's critical path: these two warps
    // issue the stores (32-row pieces from 8 lanes) and release the staging stage when the TMA has read it.
    {
      const int half = warp - 2;
      const int bxl = lane >> 2, piece = lane & 3;      // lanes 0..7 issue
      const uint32_t peer_stage_free = mapa_u32(stage_free, rank ^ 1u);
      int ui = 0;
      GRing ring;
      for (int u = cluster_id; u < p.total_units; u += n_clusters, ++ui) {
        const GraphUnit un = graph_unit(p, u, rank);
        uint8_t* stg = smem + ring.advance_unit(J, p.ring_skip != 0) * G_STAGE_BYTES;
        if (lane < 8) {
          const int bx = half * 2 + bxl;
          const int cb = un.c0 + bx * 64;
          mbar_wait(&staged[bx], (uint32_t)(ui & 1));
          if (cb < p.ldy) {
            tma_store_3d(&tmY, stg + bx * G_BOX_BYTES + piece * (G_BOX_BYTES / 4), cb, un.i0 + piece * 32, un.b);   // rows >= N are clipped
            tma_store_commit();
          }
          tma_store_wait_read();             // the staging stage may be refilled once the TMA has read it ...
        }
        __syncwarp();
        if (lane == 0) {
          if (warp == 2 && ui == 0) G_TICK(50);
          // (with fewer than 3 key tiles per unit nothing else keeps a CTA from running a whole unit ahead of its peer:
          //  the previous phase -- which includes the peer's arrivals on both barriers -- must be over before this one is fed)
          if (ui > 0) mbar_wait(stage_free, (uint32_t)((ui - 1) & 1));
          mbar_arrive(stage_free);           // ... in BOTH CTAs (the refill is a multicast)
          mbar_arrive_cluster(peer_stage_free);
          if (warp == 2 && ui == 0) G_TICK(51);
        }
      }
      if (lane < 8) tma_store_wait_all();
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // no CTA exits while its peer may still multicast into its smem / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
    G_TICK(43);
  }
}


// =====================================================================================================================
// 2-SM variant (NOT the default; cmpc_graph_set_mode(2) -- kept as a measured experiment): the CTA pair issues ONE
// tcgen05.mma.cta_group::2 stream with M = 256 (query rows 256p .. 256p+255, 128 in each CTA's TMEM).  The B operands are split between the two CTAs' shared memory instead of being multicast:
//   MMA1  S[256x128] = W[256x32] . V_j[128x32]^T   each CTA holds the 64 key rows 64r .. 64r+63 of V_j (4 KB)
//   MMA2  O[256x256] += P[256x128] . X_j[128x256]  each CTA holds channels 128r .. 128r+127 of X_j (32 KB), A = P from TMEM
// so a ring stage is 36 KB instead of 72 KB: four stages AND a dedicated 64 KB output staging buffer fit, the drain of a
// unit's output (4-5 k clk, see above) overlaps the whole next unit instead of holding a ring stage.
// Measured (B = 32, N = 1600): 228 us against 178 us for graph_reason_kernel.  Why it loses: (1) a cta_group::2 dispatch
// takes exactly as long as a cta_group::1 one -- 139 clk at M = 256, N = 256 with A in TMEM, 105.6 clk at N = 128
// (scripts/micro/mma_rate_2sm.cu) -- so there is no tensor-rate gain, only the smaller B footprint; (2) every S -> P
// hand-off now crosses the cluster twice (commit multicast to the peer, remote mbarrier arrive back), and with only two
// S/P buffers in TMEM that latency is on the critical path of every key tile; (3) the odd query tile costs a whole unit.
// The leader CTA (cluster rank 0) issues every MMA; both CTAs' TMA bytes complete on the leader's "full" barriers, its
// commits are multicast to both CTAs' "empty" / s_full / o_full barriers, and both CTAs' convert / epilogue warps arrive
// on the leader's p_full / o_empty barriers.  An odd last query tile is padded (its pair CTA has no valid rows).
// =====================================================================================================================
constexpr int G2_STAGES = 4;
constexpr int G2_X_BYTES = G_BJ * (G_BC / 2) * 2;          // 32768: two [128 keys x 64 ch] boxes
constexpr int G2_V_BYTES = (G_BJ / 2) * G_T * 2;           // 4096
constexpr int G2_STAGE_BYTES = G2_X_BYTES + G2_V_BYTES;    // 36864
constexpr int G2_W_OFF = G2_STAGES * G2_STAGE_BYTES;
constexpr int G2_STG_OFF = G2_W_OFF + G_BM * G_T * 2;      // 64 KB output staging, 1024-aligned
constexpr int G2_BAR_OFF = G2_STG_OFF + G_BM * G_BC * 2;
constexpr int G2_SMEM = G2_BAR_OFF + 384 + 1024;
static_assert(G2_STG_OFF % 1024 == 0 && G2_STAGE_BYTES % 1024 == 0, "swizzled tiles need 1024-byte alignment");

__global__ void __launch_bounds__(G_THREADS, 1)
graph_reason_2sm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmV,
                        const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const GraphParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + G2_BAR_OFF);   // leader's: both CTAs' X halves have landed
  uint64_t* x_empty = x_full + G2_STAGES;
  uint64_t* v_full = x_empty + G2_STAGES;
  uint64_t* v_empty = v_full + G2_STAGES;
  uint64_t* w_full = v_empty + G2_STAGES;
  uint64_t* w_empty = w_full + 1;
  uint64_t* s_full = w_empty + 1;    // [2]
  uint64_t* p_full = s_full + 2;     // [2] leader's: 16 convert warps
  uint64_t* o_full = p_full + 2;
  uint64_t* o_empty = o_full + 1;    // leader's: 16 epilogue warps
  uint64_t* stg_free = o_empty + 1;  // local: the store warps have drained the staging buffer
  uint64_t* staged = stg_free + 1;   // [4] local: output box bx is complete in the staging buffer
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(staged + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = p.j_tiles;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / G_CLUSTER, n_clusters = gridDim.x / G_CLUSTER;
  constexpr uint16_t kAll = (1u << G_CLUSTER) - 1;
  const int pairs = (p.i_tiles + 1) / 2;
  const int units_per_sample = pairs * p.c_chunks;
  const int total_units = units_per_sample * p.batch;
  auto unit_of = [&](int u) {
    GraphUnit g;
    g.b = u / units_per_sample;
    const int r = u - g.b * units_per_sample;
    const int pair = r / p.c_chunks;
    g.chunk = r - pair * p.c_chunks;
    g.i0 = (2 * pair + (int)rank) * G_BM;
    g.c0 = g.chunk * G_BC;
    g.own_x = false;
    return g;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    mbar_init(w_full, 1); mbar_init(w_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 8 * G_CLUSTER); }
    mbar_init(o_full, 1); mbar_init(o_empty, 8 * G_CLUSTER);
    mbar_init(stg_free, 2);
    for (int s = 0; s < 4; ++s) mbar_init(&staged[s], 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own halves, completion on the leader's barriers) =====================
    if (lane == 0) {
      {
        const int rem = total_units % n_clusters;               // see graph_reason_kernel: de-synchronise the output bursts
        if (rem != 0 && cluster_id >= rem) {
          const long long unit_clk = (long long)J * 1300 + 2500;
          const long long delay = unit_clk * 9 / 10 * (cluster_id - rem + 1) / (n_clusters - rem + 1);
          const long long t0 = clock64();
          while (clock64() - t0 < delay) { }
        }
      }
      const uint32_t lead_w_full = mapa_u32(w_full, 0);
      uint32_t g = 0;
      int ui = 0;
      for (int u = cluster_id; u < total_units; u += n_clusters, ++ui) {
        const GraphUnit un = unit_of(u);
        if (ui > 0) mbar_wait(w_empty, (uint32_t)((ui - 1) & 1));      // every MMA1 of the previous unit has read W (both CTAs)
        if (rank == 0) mbar_expect_tx(w_full, 2 * G_BM * G_T * 2);
        tma_load_3d_2sm(smem + G2_W_OFF, &tmW, lead_w_full, 0, un.i0, un.b);
        for (int j = 0; j < J; ++j, ++g) {
          const int s = (int)(g % G2_STAGES);
          const uint32_t ph = (g / G2_STAGES) & 1;
          uint8_t* sx = smem + s * G2_STAGE_BYTES;
          mbar_wait(&v_empty[s], ph ^ 1);
          if (rank == 0) mbar_expect_tx(&v_full[s], 2 * G2_V_BYTES);
          tma_load_3d_2sm(sx + G2_X_BYTES, &tmV, mapa_u32(&v_full[s], 0), 0, j * G_BJ + (int)rank * (G_BJ / 2), un.b);
          mbar_wait(&x_empty[s], ph ^ 1);
          if (rank == 0) mbar_expect_tx(&x_full[s], 2 * G2_X_BYTES);
          const uint32_t lead_x_full = mapa_u32(&x_full[s], 0);
#pragma unroll
          for (int m = 0; m < 2; ++m)
            tma_load_3d_2sm(sx + m * G_BOX_BYTES, &tmX, lead_x_full, un.c0 + ((int)rank * 2 + m) * 64, j * G_BJ, un.b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc1 = make_idesc_f16(2 * G_BM, G_BJ, 0, 0, 0);   // S = W V^T, both K-major
      constexpr uint32_t idesc2 = make_idesc_f16(2 * G_BM, G_BC, 0, 0, 1);   // O += P X, B (X) MN-major
      const uint32_t w_addr = smem_u32(smem + G2_W_OFF);
      uint32_t g0 = 0;
      int ui = 0;
      for (int u = cluster_id; u < total_units; u += n_clusters, ++ui, g0 += (uint32_t)J) {
        auto issue_mma1 = [&](int j) {
          const uint32_t g = g0 + (uint32_t)j;
          const int st = (int)(g % G2_STAGES);
          mbar_wait(&v_full[st], (g / G2_STAGES) & 1);
          tc_fence_after();
          const uint32_t v_addr = smem_u32(smem + st * G2_STAGE_BYTES + G2_X_BYTES);
          const uint64_t dw = make_smem_desc(w_addr, 16, 512, 4);   // 64-byte swizzle, 8 rows x 64 B atoms
          const uint64_t dv = make_smem_desc(v_addr, 16, 512, 4);
          const uint32_t d = tmem_base + G_COL_S + 128 * (g & 1);
#pragma unroll
          for (int k = 0; k < G_T / 16; ++k) umma_f16_ss_2sm(d, dw + uint64_t(k * 2), dv + uint64_t(k * 2), idesc1, k != 0 ? 1u : 0u);
          umma_commit_2sm_mc(&s_full[g & 1], kAll);
          umma_commit_2sm_mc(&v_empty[st], kAll);
          if (j == J - 1) umma_commit_2sm_mc(w_empty, kAll);
        };
        mbar_wait(w_full, (uint32_t)(ui & 1));
        tc_fence_after();
        issue_mma1(0);
        if (J > 1) issue_mma1(1);
        for (int j = 0; j < J; ++j) {
          const uint32_t g = g0 + (uint32_t)j;
          const int st = (int)(g % G2_STAGES);
          mbar_wait(&x_full[st], (g / G2_STAGES) & 1);
          mbar_wait(&p_full[g & 1], (g >> 1) & 1);
          if (j == 0 && ui > 0) mbar_wait(o_empty, (uint32_t)((ui - 1) & 1));    // the previous unit's O has left TMEM (both CTAs)
          tc_fence_after();
          const uint32_t x_addr = smem_u32(smem + st * G2_STAGE_BYTES);
          const uint64_t dx = make_smem_desc(x_addr, G_BOX_BYTES, 1024, 2);      // MN-major, 128B swizzle (see graph_reason_kernel)
          const uint32_t a_tmem = tmem_base + G_COL_S + 128 * (g & 1);
          const int ksteps = (j == J - 1) ? p.last_ksteps : G_BJ / 16;
#pragma unroll
          for (int k = 0; k < G_BJ / 16; ++k)
            if (k < ksteps)
              umma_f16_ts_2sm(tmem_base + G_COL_O, a_tmem + (k >> 2) * 64 + (k & 3) * 8, dx + uint64_t((k * 16 * 128) >> 4), idesc2,
                              (j | k) != 0 ? 1u : 0u);
          umma_commit_2sm_mc(&x_empty[st], kAll);
          if (j + 2 < J) issue_mma1(j + 2);
        }
        umma_commit_2sm_mc(o_full, kAll);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== convert (S -> P) and epilogue: identical per CTA to graph_reason_kernel =====================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t lead_p_full[2] = {mapa_u32(&p_full[0], 0), mapa_u32(&p_full[1], 0)};
    const uint32_t lead_o_empty = mapa_u32(o_empty, 0);
    uint8_t* stg = smem + G2_STG_OFF;
    uint32_t g0 = 0;
    int ui = 0;
    for (int u = cluster_id; u < total_units; u += n_clusters, ++ui, g0 += (uint32_t)J) {
      const GraphUnit un = unit_of(u);
      const bool row_ok = un.i0 + row < p.n_nodes;
      auto convert = [&](uint32_t g, const GraphUnit& cu, int j) {
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t sbuf = tmem_base + lane_off + G_COL_S + 128 * (g & 1) + half * 64;
        uint32_t r[64];
        tmem_ld_x64(sbuf, r);
        tmem_wait_ld();
        if (p.dbg_p != nullptr && cu.chunk == 0 && cu.i0 + row < p.n_nodes) {
          float* d = p.dbg_p + ((long long)cu.b * p.n_nodes + cu.i0 + row) * p.n_nodes + j * G_BJ + half * 64;
          for (int e = 0; e < 64; ++e)
            if (j * G_BJ + half * 64 + e < p.n_nodes) d[e] = __uint_as_float(r[e]) * p.inv_vscale;
        }
        uint32_t pk[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const __half2 hh = __floats2half2_rn(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
          pk[e] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        tmem_st_x32(sbuf, pk);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_p_full[g & 1]);
      };
      for (int j = (ui > 0) ? 1 : 0; j < J; ++j) convert(g0 + (uint32_t)j, un, j);
      mbar_wait(o_full, (uint32_t)(ui & 1));
      tc_fence_after();
      if (u + n_clusters < total_units) convert(g0 + (uint32_t)J, unit_of(u + n_clusters), 0);   // see graph_reason_kernel
      if (ui > 0) mbar_wait(stg_free, (uint32_t)((ui - 1) & 1));      // the previous unit's output has left the staging buffer
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int bxl = 0; bxl < 2; ++bxl) {
        const int bx = half * 2 + bxl;
        const int col = bx * 64;
        uint32_t r[64];
        tmem_ld_x64(tmem_base + lane_off + G_COL_O + col, r);
        tmem_wait_ld();
        if (bxl == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead_o_empty);
        }
        uint8_t* box = stg + bx * G_BOX_BYTES + row * 128;
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v[e] = __uint_as_float(r[gq * 8 + e]) * p.inv_vscale;      // channels >= C are exact zeros (X is zero-filled there)
            s1 += v[e];
            s2 += v[e] * v[e];
          }
          uint4 uu;
          __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
          __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
          uu.x = *reinterpret_cast<uint32_t*>(&h0);
          uu.y = *reinterpret_cast<uint32_t*>(&h1);
          uu.z = *reinterpret_cast<uint32_t*>(&h2);
          uu.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(box + ((gq ^ (row & 7)) << 4)) = uu;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&staged[bx]);
      }
      if (p.stats) {
        if (!row_ok) { s1 = 0.f; s2 = 0.f; }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0 && un.c0 < p.C) {
          atomicAdd(p.stats + 2 * un.b, (double)s1);
          atomicAdd(p.stats + 2 * un.b + 1, (double)s2);
        }
      }
    }
    tc_fence_before();
  } else {
    // ===================== store warps (2, 3): TMA stores of the staged boxes of column half (warp - 2) =====================
    const int half = warp - 2;
    const int bxl = lane >> 2, piece = lane & 3;      // lanes 0..7 issue one 32-row piece each
    uint8_t* stg = smem + G2_STG_OFF;
    int ui = 0;
    for (int u = cluster_id; u < total_units; u += n_clusters, ++ui) {
      const GraphUnit un = unit_of(u);
      if (lane < 8) {
        const int bx = half * 2 + bxl;
        const int cb = un.c0 + bx * 64;
        mbar_wait(&staged[bx], (uint32_t)(ui & 1));
        if (cb < p.ldy) {
          tma_store_3d(&tmY, stg + bx * G_BOX_BYTES + piece * (G_BOX_BYTES / 4), cb, un.i0 + piece * 32, un.b);   // rows >= N are clipped
          tma_store_commit();
        }
        tma_store_wait_read();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(stg_free);
    }
    if (lane < 8) tma_store_wait_all();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // no CTA exits while its peer may still signal its barriers / read its shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

static bool g_graph_two_sm = false;
static bool g_graph_p16 = false;
static bool g_graph_round_robin = false;     // mode 4: plain round-robin ring (the kernel before the skip-once order; A/B knob)
static bool g_graph_convert_first = false;   // modes 4, 5: the next unit's tile 0 is converted before the accumulator is drained

}  // namespace cmpc

using namespace cmpc;

extern "C" void cmpc_graph_set_mode(int mode) {
  cmpc::g_graph_two_sm = (mode == 2);
  cmpc::g_graph_p16 = (mode == 3);
  cmpc::g_graph_round_robin = (mode == 4);
  cmpc::g_graph_convert_first = (mode == 4 || mode == 5);
}

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
  rc = make_tmap_3d_sw(&tY, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y_f16, ldy, n_nodes, batch, ldy * 2, (uint64_t)n_nodes * ldy * 2, 64, 32,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(graph_reason_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "cudaFuncSetAttribute(graph, smem=%d): %s", G_SMEM, cudaGetErrorString(e));
    e = cudaFuncSetAttribute(graph_reason_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM);
    CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "cudaFuncSetAttribute(graph 2sm, smem=%d): %s", G2_SMEM, cudaGetErrorString(e));
  }
  GraphParams p{};
  p.n_nodes = n_nodes; p.C = c; p.j_tiles = (n_nodes + G_BJ - 1) / G_BJ;
  p.i_tiles = (n_nodes + G_BM - 1) / G_BM;
  p.c_chunks = (c + G_BC - 1) / G_BC;
  p.batch = batch;
  p.pair_units = (p.i_tiles / 2) * p.c_chunks;
  p.units_per_sample = p.pair_units + ((p.i_tiles & 1) ? (p.c_chunks + 1) / 2 : 0);
  p.total_units = p.units_per_sample * batch;
  p.last_ksteps = (n_nodes - (p.j_tiles - 1) * G_BJ + 15) / 16;
  p.inv_vscale = 1.0f / v_scale;
  p.ldy = ldy; p.stats = stats; p.dbg_p = dbg_p;
  p.p16 = g_graph_p16 ? 1 : 0;
  p.epi_order = g_graph_convert_first ? 0 : 1;
  p.ring_skip = (!g_graph_round_robin && p.j_tiles >= 6) ? 1 : 0;     // needs tiles 2..4 of a unit to exist
  const bool two = g_graph_two_sm;
  const int total_units = two ? ((p.i_tiles + 1) / 2) * p.c_chunks * batch : p.total_units;
  const int max_clusters = num_sms() / G_CLUSTER;
  const int clusters = total_units < max_clusters ? total_units : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * G_CLUSTER);
  cfg.blockDim = dim3(G_THREADS);
  cfg.dynamicSmemBytes = two ? G2_SMEM : G_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = two ? cudaLaunchKernelEx(&cfg, graph_reason_2sm_kernel, tW, tV, tX, tY, p)
                      : cudaLaunchKernelEx(&cfg, graph_reason_kernel, tW, tV, tX, tY, p);
  CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "graph_reason_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("graph_reason_kernel");
}

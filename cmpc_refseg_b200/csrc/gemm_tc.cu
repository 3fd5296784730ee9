// Persistent warp-specialised tcgen05 GEMM for the CMPC head's 1x1 convolutions (89 % of its FLOPs).
//
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, 4-stage mbarrier ring)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16; fp32 accumulators in TMEM)
//   warp 2      TMEM allocator (512 columns = 2 accumulator stages, so the epilogue of tile i overlaps tile i+1)
//   warps 4-7   epilogue       (tcgen05.ld 32x32b -> registers -> fused bias/act/gate/statistics -> global)
//
// Replaces LSTM_model._conv (CMPC_model.py:412-417) for filter_size 1 plus the elementwise nodes after each
// call site; the MUTAN variant replaces mutan_head/mutan_fusion (:295-323) for all five heads at once.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmpc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 fp16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;  // TMEM columns per accumulator stage

enum { EPI_GENERIC = 0, EPI_MUTAN = 1 };

struct GemmKernelParams {
  int M, N;              // N = number of weight rows covered by tiles (all of them)
  int kt1, kt2;          // k-tiles from A1 / A2
  int m_tiles, n_tiles;
  int rows_per_sample;
  int batched;           // 1: tiles are per sample, W is [B][w_rows][K] (3-D tensor map), rows masked at rows_per_sample
  int batch;
  // generic epilogue
  const float* row_scale;
  const float* bias;
  const float* sbias;  long long ld_sbias;
  const float* gate;   long long ld_gate;
  int act;
  int group_width, group_valid, n_groups;
  const float* peep_i; const float* peep_f; long long ld_peep;
  const float* cprev;  long long ld_cprev;
  void* out; long long ldo; int out_fp32;
  float* row_sumsq;
  double* stats;
  // mutan epilogue
  int C;                   // channels
  const float* mbias; long long ld_mbias;   // [5, ld]
  const float* lang;  long long ld_lang;  long long lang_bstride;   // [B][5][ld], sample stride lang_bstride
};

template <int BN>
struct SmemCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFF = BAR_OFF + 256;                 // per-warp staged epilogue vectors
  static constexpr int EPI_WARP_FLOATS = 2 * 256;               // [add | mul], 256 columns each
  static constexpr int TOTAL = EPI_OFF + 4 * EPI_WARP_FLOATS * 4 + 1024;  // + alignment slack
  static_assert(B_BYTES % 1024 == 0, "B tile must keep 1024-byte swizzle-atom alignment");
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------------------------------------
// epilogues (executed by the 128 epilogue threads; thread <-> accumulator row)
//
// Per-column epilogue operands (bias + per-sample bias -> "add", per-sample gate -> "mul") are fetched ONCE per
// tile by each warp into its private shared-memory slice *before* it waits for the accumulator, so their global
// latency hides behind the tile's MMAs; the per-chunk math is then branch-free (invalid columns carry add = mul = 0).
// ---------------------------------------------------------------------------------------------
struct EpiCtx {
  int m, mm, b, grp, cbase;
  bool row_ok, uniform, peep;
  float rs;
  const float* sb; const float* gt; const float* pe; const float* cp;
};

template <int BN>
__device__ __forceinline__ void epi_generic_prefetch(const GemmKernelParams& p, int m0, int n0, int tile_b, int q, int lane,
                                                     float* s_add, float* s_mul, EpiCtx& c) {
  // flattened: m0 is the global row of the tile; batched: m0 is the row inside sample tile_b
  const int lr = m0 + q * 32 + lane;
  c.row_ok = p.batched ? (lr < p.rows_per_sample) : (lr < p.M);
  c.m = p.batched ? tile_b * p.rows_per_sample + lr : lr;
  c.mm = c.row_ok ? c.m : 0;
  int b = p.batched ? tile_b : c.mm / p.rows_per_sample;
  const int b0 = __shfl_sync(0xffffffffu, b, 0);       // rows grow with the lane: lane 0 valid unless the whole warp is not
  if (!c.row_ok) b = b0;
  c.b = b;
  c.uniform = __all_sync(0xffffffffu, b == b0);
  const int pix = c.mm - b * p.rows_per_sample;
  const int gw = p.group_width > 0 ? p.group_width : (1 << 30);
  c.grp = n0 / gw;                 // tiles never straddle groups (checked on the host)
  c.cbase = n0 - c.grp * gw;       // column inside the group
  c.rs = (p.row_scale && c.row_ok) ? __ldg(p.row_scale + c.mm) : 1.0f;
  c.sb = p.sbias ? p.sbias + (long long)b * p.ld_sbias : nullptr;
  c.gt = p.gate ? p.gate + (long long)b * p.ld_gate : nullptr;
  c.peep = p.cprev != nullptr && (c.grp == 1 || c.grp == 2);
  c.pe = c.peep ? (c.grp == 1 ? p.peep_i : p.peep_f) + (long long)pix * p.ld_peep : nullptr;
  c.cp = c.peep ? p.cprev + (long long)c.mm * p.ld_cprev : nullptr;
  if (c.uniform) {
    constexpr int PER = BN / 32;   // columns staged by each lane
#pragma unroll
    for (int e4 = 0; e4 < (PER + 3) / 4; ++e4) {
      const int col = lane * PER + e4 * 4;
      const int cc = c.cbase + col, n = n0 + col;
      if (PER >= 4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), g = a;
        if (cc + 3 < p.group_valid) {    // group_valid % 4 == 0 (host check)
          g = make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.bias) a = ldg4(p.bias + n);
          if (c.sb) { const float4 t = ldg4(c.sb + n); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
          if (c.gt) g = ldg4(c.gt + n);
        }
        *reinterpret_cast<float4*>(s_add + col) = a;
        *reinterpret_cast<float4*>(s_mul + col) = g;
      } else {
        float a = 0.f, g = 0.f;
        if (cc < p.group_valid) {
          g = 1.f;
          if (p.bias) a = __ldg(p.bias + n);
          if (c.sb) a += __ldg(c.sb + n);
          if (c.gt) g = __ldg(c.gt + n);
        }
        s_add[col] = a;
        s_mul[col] = g;
      }
    }
    __syncwarp();
  }
}

template <int BN>
__device__ __forceinline__ void epi_generic_compute(const GemmKernelParams& p, uint32_t tmem_acc, int n0, int q, int lane,
                                                    const float* s_add, const float* s_mul, const EpiCtx& c) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
  for (int ch = 0; ch < BN / 32; ++ch) {
    const int nb = n0 + ch * 32;          // global column of r[0]
    const int cb = c.cbase + ch * 32;     // column within group
    if (nb >= p.ldo) break;               // warp-uniform; later chunks are further right
    float4 pe4[8], cp4[8];
    if (c.peep) {                         // ConvLSTM peepholes: issue all loads of the chunk before touching TMEM
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        pe4[j4] = ldg4(c.pe + cb + j4 * 4);
        cp4[j4] = ldg4(c.cp + cb + j4 * 4);
      }
    }
    uint32_t r[32];
    tmem_ld_x32(tmem_acc + (uint32_t(q * 32) << 16) + ch * 32, r);
    tmem_wait_ld();
    float v[32];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      float4 a, g;
      if (c.uniform) {
        a = *reinterpret_cast<const float4*>(s_add + ch * 32 + j4 * 4);
        g = *reinterpret_cast<const float4*>(s_mul + ch * 32 + j4 * 4);
      } else {   // rows of different samples in one warp (odd shapes only): per-thread loads
        a = make_float4(0.f, 0.f, 0.f, 0.f); g = a;
        if (cb + j4 * 4 + 3 < p.group_valid) {
          g = make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.bias) a = ldg4(p.bias + nb + j4 * 4);
          if (c.sb) { const float4 t = ldg4(c.sb + nb + j4 * 4); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
          if (c.gt) g = ldg4(c.gt + nb + j4 * 4);
        }
      }
      float x0 = fmaf(__uint_as_float(r[j4 * 4 + 0]), c.rs, a.x), x1 = fmaf(__uint_as_float(r[j4 * 4 + 1]), c.rs, a.y);
      float x2 = fmaf(__uint_as_float(r[j4 * 4 + 2]), c.rs, a.z), x3 = fmaf(__uint_as_float(r[j4 * 4 + 3]), c.rs, a.w);
      if (c.peep) {
        x0 = fmaf(pe4[j4].x, cp4[j4].x, x0); x1 = fmaf(pe4[j4].y, cp4[j4].y, x1);
        x2 = fmaf(pe4[j4].z, cp4[j4].z, x2); x3 = fmaf(pe4[j4].w, cp4[j4].w, x3);
      }
      if (p.act == 1) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
      else if (p.act == 2) { x0 = tanh_acc(x0); x1 = tanh_acc(x1); x2 = tanh_acc(x2); x3 = tanh_acc(x3); }
      else if (p.act == 3) { x0 = sigmoid_acc(x0); x1 = sigmoid_acc(x1); x2 = sigmoid_acc(x2); x3 = sigmoid_acc(x3); }
      x0 *= g.x; x1 *= g.y; x2 *= g.z; x3 *= g.w;      // invalid columns: add = mul = 0 -> exactly 0
      v[j4 * 4 + 0] = x0; v[j4 * 4 + 1] = x1; v[j4 * 4 + 2] = x2; v[j4 * 4 + 3] = x3;
      s1 += (x0 + x1) + (x2 + x3);
      s2 += (x0 * x0 + x1 * x1) + (x2 * x2 + x3 * x3);
    }
    if (c.row_ok) {
      if (p.out_fp32) {
        float* o = reinterpret_cast<float*>(p.out) + (long long)c.m * p.ldo + nb;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          if (nb + j4 * 4 + 3 < p.ldo)
            *reinterpret_cast<float4*>(o + j4 * 4) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
      } else {
        __half* o = reinterpret_cast<__half*>(p.out) + (long long)c.m * p.ldo + nb;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          if (nb + j8 * 8 + 7 < p.ldo) {
            uint4 u;
            __half2 h0 = __floats2half2_rn(v[j8 * 8 + 0], v[j8 * 8 + 1]);
            __half2 h1 = __floats2half2_rn(v[j8 * 8 + 2], v[j8 * 8 + 3]);
            __half2 h2 = __floats2half2_rn(v[j8 * 8 + 4], v[j8 * 8 + 5]);
            __half2 h3 = __floats2half2_rn(v[j8 * 8 + 6], v[j8 * 8 + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&h0);
            u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2);
            u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(o + j8 * 8) = u;
          }
        }
      }
    }
  }
  if (p.row_sumsq && c.row_ok) atomicAdd(p.row_sumsq + c.m, s2);
  if (p.stats) {
    if (!c.row_ok) { s1 = 0.f; s2 = 0.f; }
    if (c.uniform) {
      const float t1 = warp_sum(s1), t2 = warp_sum(s2);
      if (lane == 0) {
        double* st = p.stats + ((long long)c.b * p.n_groups + c.grp) * 2;
        atomicAdd(st, (double)t1);
        atomicAdd(st + 1, (double)t2);
      }
    } else if (c.row_ok) {
      double* st = p.stats + ((long long)c.b * p.n_groups + c.grp) * 2;
      atomicAdd(st, (double)s1);
      atomicAdd(st + 1, (double)s2);
    }
  }
}

// MUTAN: BN = 240 = 5 heads x 48 channels.  out = tanh(sum_k tanh(acc_k + bias_k) * lang_k)
__device__ __forceinline__ void epi_mutan_prefetch(const GemmKernelParams& p, int m0, int jchunk, int q, int lane,
                                                   float* s_bias, float* s_lang, EpiCtx& c) {
  c.m = m0 + q * 32 + lane;
  c.row_ok = c.m < p.M;
  c.mm = c.row_ok ? c.m : 0;
  int b = c.mm / p.rows_per_sample;
  const int b0 = __shfl_sync(0xffffffffu, b, 0);
  if (!c.row_ok) b = b0;
  c.b = b;
  c.uniform = __all_sync(0xffffffffu, b == b0);
  if (c.uniform) {
    const float* lang = p.lang + (long long)b * p.lang_bstride;
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int f = lane + rep * 32;          // float4 index inside [5 heads][12 float4]
      if (f < 60) {
        const int k = f / 12, i4 = f - k * 12;
        const int ch = jchunk * 48 + i4 * 4;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), ll = bb;
        if (ch + 3 < p.C) {
          bb = ldg4(p.mbias + (long long)k * p.ld_mbias + ch);
          ll = ldg4(lang + (long long)k * p.ld_lang + ch);
        }
        *reinterpret_cast<float4*>(s_bias + k * 48 + i4 * 4) = bb;
        *reinterpret_cast<float4*>(s_lang + k * 48 + i4 * 4) = ll;
      }
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void epi_mutan_compute(const GemmKernelParams& p, uint32_t tmem_acc, int jchunk, int q, int lane,
                                                  const float* s_bias, const float* s_lang, const EpiCtx& c) {
  const float* lang = p.lang + (long long)c.b * p.lang_bstride;
  float ss = 0.f;
#pragma unroll 1
  for (int s = 0; s < 3; ++s) {
    const int c0 = jchunk * 48 + s * 16;
    if (c0 >= p.ldo) break;  // warp-uniform
    uint32_t r[5][16];
#pragma unroll
    for (int k = 0; k < 5; ++k) tmem_ld_x16(tmem_acc + (uint32_t(q * 32) << 16) + k * 48 + s * 16, r[k]);
    tmem_wait_ld();
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        float4 bb, ll;
        if (c.uniform) {
          bb = *reinterpret_cast<const float4*>(s_bias + k * 48 + s * 16 + i4 * 4);
          ll = *reinterpret_cast<const float4*>(s_lang + k * 48 + s * 16 + i4 * 4);
        } else {
          bb = make_float4(0.f, 0.f, 0.f, 0.f); ll = bb;
          if (c0 + i4 * 4 + 3 < p.C) {
            bb = ldg4(p.mbias + (long long)k * p.ld_mbias + c0 + i4 * 4);
            ll = ldg4(lang + (long long)k * p.ld_lang + c0 + i4 * 4);
          }
        }
        acc[i4 * 4 + 0] = fmaf(tanh_acc(__uint_as_float(r[k][i4 * 4 + 0]) + bb.x), ll.x, acc[i4 * 4 + 0]);
        acc[i4 * 4 + 1] = fmaf(tanh_acc(__uint_as_float(r[k][i4 * 4 + 1]) + bb.y), ll.y, acc[i4 * 4 + 1]);
        acc[i4 * 4 + 2] = fmaf(tanh_acc(__uint_as_float(r[k][i4 * 4 + 2]) + bb.z), ll.z, acc[i4 * 4 + 2]);
        acc[i4 * 4 + 3] = fmaf(tanh_acc(__uint_as_float(r[k][i4 * 4 + 3]) + bb.w), ll.w, acc[i4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      acc[i] = tanh_acc(acc[i]);     // columns >= C have bias = lang = 0 and zero weights -> tanh(0) = 0
      ss += acc[i] * acc[i];
    }
    if (c.row_ok) {
      float* o = reinterpret_cast<float*>(p.out) + (long long)c.m * p.ldo + c0;
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4)
        if (c0 + i4 * 4 + 3 < p.ldo)
          *reinterpret_cast<float4*>(o + i4 * 4) = make_float4(acc[i4 * 4], acc[i4 * 4 + 1], acc[i4 * 4 + 2], acc[i4 * 4 + 3]);
    }
  }
  if (p.row_sumsq && c.row_ok) atomicAdd(p.row_sumsq + c.m, ss);
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const GemmKernelParams p) {
  using Cfg = SmemCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles * (p.batched ? p.batch : 1);   // batched: m_tiles is per sample
  const int kt_total = p.kt1 + p.kt2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    if (p.kt2 > 0) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles;
        const int n0 = (tile % p.n_tiles) * BN;
        // batched: mt = b * m_tiles_per_sample + local tile; A rows start at b * rows_per_sample + local * 128
        const int tb = p.batched ? mt / p.m_tiles : 0;
        const int m0 = p.batched ? tb * p.rows_per_sample + (mt - tb * p.m_tiles) * BLOCK_M : mt * BLOCK_M;
        for (int kt = 0; kt < kt_total; ++kt) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (kt < p.kt1) tma_load_2d(sa, &tmA1, &full_bar[s], kt * BLOCK_K, m0);
          else            tma_load_2d(sa, &tmA2, &full_bar[s], (kt - p.kt1) * BLOCK_K, m0);
          if (p.batched) tma_load_3d(sb, &tmW, &full_bar[s], kt * BLOCK_K, n0, tb);
          else           tma_load_2d(sb, &tmW, &full_bar[s], kt * BLOCK_K, n0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BLOCK_M, BN, /*fp16*/ 0, /*A K-major*/ 0, /*B K-major*/ 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
        for (int kt = 0; kt < kt_total; ++kt) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = make_smem_desc(sa, 16, 1024, 2);
          const uint64_t db = make_smem_desc(sb, 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in the address field
            umma_f16_ss(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kt | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full[as]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp - 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int mt = tile / p.n_tiles;
      const int nt = tile % p.n_tiles;
      const int tb = p.batched ? mt / p.m_tiles : 0;
      const int m0 = (p.batched ? (mt - tb * p.m_tiles) : mt) * BLOCK_M;   // batched: row inside the sample
      float* s_add = reinterpret_cast<float*>(smem + Cfg::EPI_OFF) + q * Cfg::EPI_WARP_FLOATS;
      float* s_mul = s_add + 256;
      EpiCtx ctx;
      if (EPI == EPI_GENERIC) epi_generic_prefetch<BN>(p, m0, nt * BN, tb, q, lane, s_add, s_mul, ctx);
      else                    epi_mutan_prefetch(p, m0, nt, q, lane, s_add, s_mul, ctx);
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const uint32_t acc = tmem_base + as * ACC_STRIDE;
      if (EPI == EPI_GENERIC) epi_generic_compute<BN>(p, acc, nt * BN, q, lane, s_add, s_mul, ctx);
      else                    epi_mutan_compute(p, acc, nt, q, lane, s_add, s_mul, ctx);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w, const GemmKernelParams& p,
                       cudaStream_t stream) {
  using Cfg = SmemCfg<BN>;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, EPI>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::TOTAL);
    CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "cudaFuncSetAttribute(smem=%d): %s", Cfg::TOTAL, cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * (p.batched ? p.batch : 1);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, GEMM_THREADS, Cfg::TOTAL, stream>>>(a1, a2, w, p);
  return check_launch("gemm_tc_kernel");
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_gemm_f16(const cmpc_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CMPC_REQUIRE(a != nullptr, CMPC_ERR_ARG, "cmpc_gemm_f16: null args");
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(a->a1 && a->w && a->out, CMPC_ERR_ARG, "cmpc_gemm_f16: null operand");
  CMPC_REQUIRE(a->m > 0 && a->n > 0 && a->k1 > 0 && a->k2 >= 0, CMPC_ERR_ARG, "cmpc_gemm_f16: bad shape m=%d n=%d k1=%d k2=%d",
               a->m, a->n, a->k1, a->k2);
  CMPC_REQUIRE(a->rows_per_sample >= 1, CMPC_ERR_ARG, "cmpc_gemm_f16: rows_per_sample must be >= 1");
  CMPC_REQUIRE(a->lda1 % 8 == 0 && a->lda2 % 8 == 0 && a->ldw % 8 == 0, CMPC_ERR_ALIGN, "cmpc_gemm_f16: lda / ldw must be multiples of 8");
  const int gw = a->group_width;
  const int gv = gw > 0 ? a->group_valid : a->n;
  CMPC_REQUIRE(gv % 4 == 0, CMPC_ERR_ARG, "cmpc_gemm_f16: valid columns (%d) must be a multiple of 4", gv);
  CMPC_REQUIRE(gw == 0 || (gw % 256 == 0 && a->n % gw == 0 && gv <= gw), CMPC_ERR_ARG,
               "cmpc_gemm_f16: group_width must be a multiple of 256 dividing n");
  CMPC_REQUIRE(a->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, CMPC_ERR_ALIGN,
               "cmpc_gemm_f16: out must be 16-byte aligned with ldo %% 8 == 0");
  CMPC_REQUIRE(a->ldo >= (gw > 0 ? (int64_t)(a->n - gw + gv) : (int64_t)a->n), CMPC_ERR_ARG, "cmpc_gemm_f16: ldo %lld smaller than n",
               (long long)a->ldo);
  const int kt1 = ceil_div(a->k1, BLOCK_K), kt2 = a->k2 > 0 ? ceil_div(a->k2, BLOCK_K) : 0;
  CMPC_REQUIRE(a->ldw >= (int64_t)(kt1 + kt2) * BLOCK_K || (kt2 == 0 && a->ldw >= a->k1), CMPC_ERR_ARG,
               "cmpc_gemm_f16: ldw %lld < padded K %d", (long long)a->ldw, (kt1 + kt2) * BLOCK_K);
  if (a->peep_i || a->peep_f || a->cprev)
    CMPC_REQUIRE(a->peep_i && a->peep_f && a->cprev && gw > 0, CMPC_ERR_ARG, "cmpc_gemm_f16: peepholes need peep_i, peep_f, cprev and groups");
  const bool batched = a->w_batch_stride != 0;
  const bool narrow = a->n <= 32;   // skinny outputs (affinity: N = T <= 32) use the 32-column tile
  CMPC_REQUIRE(!batched || (a->m % a->rows_per_sample == 0 && kt2 == 0), CMPC_ERR_ARG,
               "cmpc_gemm_f16: batched W needs m %% rows_per_sample == 0 and a single K segment");
  const int w_rows = a->w_rows > 0 ? a->w_rows : a->n;
  const int bn = narrow ? 32 : 256;
  // W inner extent: the padded K when W physically holds it, else the exact K (TMA zero-fills the tail)
  const uint64_t w_inner = (a->ldw >= (int64_t)(kt1 + kt2) * BLOCK_K) ? (uint64_t)(kt1 + kt2) * BLOCK_K : (uint64_t)a->k1;

  CUtensorMap tA1, tA2, tW;
  rc = make_tmap_2d(&tA1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->a1, a->k1, a->m, a->lda1 * 2, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  if (kt2 > 0) {
    CMPC_REQUIRE(a->a2 != nullptr, CMPC_ERR_ARG, "cmpc_gemm_f16: k2 > 0 but a2 is null");
    rc = make_tmap_2d(&tA2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->a2, a->k2, a->m, a->lda2 * 2, BLOCK_K, BLOCK_M);
    if (rc) return rc;
  } else {
    tA2 = tA1;
  }
  const int batch = batched ? a->m / a->rows_per_sample : 1;
  if (batched)
    rc = make_tmap_3d(&tW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->w, w_inner, w_rows, batch, a->ldw * 2, a->w_batch_stride * 2, BLOCK_K, bn);
  else
    rc = make_tmap_2d(&tW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->w, w_inner, w_rows, a->ldw * 2, BLOCK_K, bn);
  if (rc) return rc;

  GemmKernelParams p{};
  p.M = a->m; p.N = a->n; p.kt1 = kt1; p.kt2 = kt2;
  p.batched = batched ? 1 : 0; p.batch = batch;
  p.m_tiles = batched ? ceil_div(a->rows_per_sample, BLOCK_M) : ceil_div(a->m, BLOCK_M);
  p.n_tiles = ceil_div(a->n, bn);
  p.rows_per_sample = a->rows_per_sample;
  p.row_scale = a->row_scale; p.bias = a->bias;
  p.sbias = a->sbias; p.ld_sbias = a->ld_sbias;
  p.gate = a->gate; p.ld_gate = a->ld_gate;
  p.act = a->act;
  p.group_width = gw; p.group_valid = gv; p.n_groups = gw > 0 ? a->n / gw : 1;
  p.peep_i = a->peep_i; p.peep_f = a->peep_f; p.ld_peep = a->ld_peep;
  p.cprev = a->cprev; p.ld_cprev = a->ld_cprev;
  p.out = a->out; p.ldo = a->ldo; p.out_fp32 = a->out_fp32;
  p.row_sumsq = a->row_sumsq; p.stats = a->stats;
  if (narrow) return launch_gemm<32, EPI_GENERIC>(tA1, tA2, tW, p, stream);
  return launch_gemm<256, EPI_GENERIC>(tA1, tA2, tW, p, stream);
}

extern "C" int cmpc_mutan_f16(const cmpc_mutan_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CMPC_REQUIRE(a != nullptr, CMPC_ERR_ARG, "cmpc_mutan_f16: null args");
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(a->a && a->w && a->out && a->bias && a->lang, CMPC_ERR_ARG, "cmpc_mutan_f16: null operand");
  CMPC_REQUIRE(a->m > 0 && a->c > 0 && a->c % 8 == 0 && a->k > 0 && a->lda % 8 == 0 && a->ldw % 8 == 0, CMPC_ERR_ARG,
               "cmpc_mutan_f16: bad shape m=%d c=%d k=%d", a->m, a->c, a->k);
  CMPC_REQUIRE(a->ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, CMPC_ERR_ALIGN,
               "cmpc_mutan_f16: out must be 16-byte aligned with ldo %% 4 == 0");
  constexpr int BN = 240;
  const int kt = ceil_div(a->k, BLOCK_K);
  const int chunks = ceil_div(a->c, 48);
  CMPC_REQUIRE(a->ldw >= (int64_t)kt * BLOCK_K, CMPC_ERR_ARG, "cmpc_mutan_f16: ldw too small");
  CUtensorMap tA, tW;
  rc = make_tmap_2d(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->a, a->k, a->m, a->lda * 2, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap_2d(&tW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->w, (uint64_t)kt * BLOCK_K, (uint64_t)chunks * BN, a->ldw * 2, BLOCK_K, BN);
  if (rc) return rc;
  GemmKernelParams p{};
  p.M = a->m; p.N = chunks * BN; p.kt1 = kt; p.kt2 = 0;
  p.m_tiles = ceil_div(a->m, BLOCK_M); p.n_tiles = chunks;
  p.rows_per_sample = a->rows_per_sample;
  p.C = a->c; p.mbias = a->bias; p.ld_mbias = a->ld_bias; p.lang = a->lang; p.ld_lang = a->ld_lang; p.lang_bstride = a->lang_batch_stride > 0 ? a->lang_batch_stride : 5 * a->ld_lang;
  p.out = a->out; p.ldo = a->ldo; p.out_fp32 = 1; p.row_sumsq = a->row_sumsq;
  return launch_gemm<BN, EPI_MUTAN>(tA, tA, tW, p, stream);
}

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
#ifndef CMPC_TWO_STAGES
#define CMPC_TWO_STAGES 6
#endif
constexpr int GEMM_THREADS = 384;      // 4 control warps + 8 epilogue warps (two per TMEM lane quadrant)
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;  // TMEM columns per accumulator stage

enum { EPI_GENERIC = 0, EPI_MUTAN = 1, EPI_MUTAN_BWD = 2 };

#ifdef CMPC_GEMM_TIMING
static long long* g_gemm_dbg = nullptr;     // [grid][8]: total, wait_full, wait_tmem_empty, issue, tiles, epi_wait_full, epi_compute
#define GT_NOW() clock64()
#endif

struct GemmKernelParams {
  long long* dbg;
  int M, N;              // N = number of weight rows covered by tiles (all of them)
  int kt1, kt2;          // k-tiles from A1 / A2
  int m_tiles, n_tiles;
  int rows_per_sample;
  int batched;           // 1: tiles are per sample, W is [B][w_rows][K] (3-D tensor map), rows masked at rows_per_sample
  int batch;
  // generic epilogue
  const float* bias;
  const float* sbias;  long long ld_sbias;
  const float* gate;   long long ld_gate;
  int act;
  int group_width, group_valid, n_groups;
  int pdl;               // launched with programmatic stream serialization: griddepcontrol.wait before the first global access
  int group_shift, ntile_shift;   // log2 of group_width / n_tiles when they are powers of two (else -1): the per-tile index math runs in every epilogue thread
  const float* peep_i; const float* peep_f; long long ld_peep;
  const float* cprev;  long long ld_cprev;
  int peep16;            // peep_i / peep_f / cprev hold fp16 (the pointers are byte-addressed through peep_esz)
  void* out; long long ldo; int out_fp32;
  float* row_sumsq;
  double* stats;
  // mutan epilogue
  int C;                   // channels
  const float* mbias; long long ld_mbias;   // [5, ld]
  const float* a_row_ss;                    // optional [M]: A rows are un-normalised, scale accumulators by rsqrt(max(ss, 1e-12))
  const float* lang;  long long ld_lang;  long long lang_bstride;   // [B][5][ld], sample stride lang_bstride
  // mutan backward epilogue (EPI_MUTAN_BWD): out = fp16 d(pre-activation) * row scale, columns in the packed (chunk, head, channel) order
  const float* ds; long long ld_ds;         // [M, ld]: gradient w.r.t. the sum of the five heads (before the outer tanh)
  float* dlang; long long dlang_bstride;    // [B][5][ld_lang]: += sum_n ds * tanh(pre_k)
  float* dmbias;                            // [5][ld_mbias]:   += sum_m d pre_k
};

// TWO = 2-SM MMA (cta_group::2): each CTA of the pair keeps only half of the weight tile, so stages are 32 KB and six fit
template <int BN, bool TWO = false>
struct SmemCfg {
  static constexpr int NSTAGES = TWO ? CMPC_TWO_STAGES : STAGES;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (TWO ? BN / 2 : BN) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = NSTAGES * STAGE_BYTES;
  static constexpr int DESC_OFF = BAR_OFF + 256;                // tile descriptor table (generic epilogue): 48 x 16 B behind the barriers
  static constexpr int EPI_OFF = BAR_OFF + 1024;                // per-warp staged epilogue vectors
  static constexpr int EPI_WARP_FLOATS = 2 * 128;               // [add | mul], up to 128 columns per epilogue warp
  static constexpr int STG_OFF = EPI_OFF + 8 * EPI_WARP_FLOATS * 4;   // per-warp output staging for the TMA stores
  static constexpr int STG_WARP_BYTES = 3072;                         // 32 rows x 64 B (generic) or x 96 B (MUTAN)
  static constexpr int TOTAL = STG_OFF + 8 * STG_WARP_BYTES + 1024;   // + alignment slack
  static_assert(EPI_OFF % 16 == 0 && STG_OFF % 128 == 0, "staging alignment");
  static_assert((2 * NSTAGES + 4) * 8 + 8 <= 256, "barriers must fit in front of the descriptor table");
  static_assert(B_BYTES % 1024 == 0, "B tile must keep 1024-byte swizzle-atom alignment");
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// explicit shared-space accesses: the epilogue's pointers into the dynamic shared memory reach it through function arguments and
// the compiler otherwise falls back to generic LD.E / ST.E (longer latency, and it could not hoist them out of branch regions)
__device__ __forceinline__ float4 lds4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts4(uint32_t saddr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ void sts4f(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 256-bit read-only load (sm_100: LDG.E.256).  A thread that owns a whole row reads its 128-byte line in 4 instead of 8
// requests; with 32 different lines per warp instruction the L1 tag stage, not bandwidth, is what these loads cost.
__device__ __forceinline__ void ldg8(const float* p, float4& lo, float4& hi) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "l"(p));
}

// ---------------------------------------------------------------------------------------------
// epilogues (executed by the 128 epilogue threads; thread <-> accumulator row)
//
// Per-column epilogue operands (bias + per-sample bias -> "add", per-sample gate -> "mul") are fetched ONCE per
// tile by each warp into its private shared-memory slice *before* it waits for the accumulator, so their global
// latency hides behind the tile's MMAs; the per-chunk math is then branch-free (invalid columns carry add = mul = 0).
// ---------------------------------------------------------------------------------------------
struct EpiCtx {
  int m, mm, b, grp, cbase;
  float rs;              // row scale of the accumulators: rsqrt(max(a_row_ss[m], 1e-12)) when the A rows are un-normalised, else 1
  bool row_ok, uniform, peep;
  const float* sb; const float* gt; const float* pe; const float* cp;
#ifdef CMPC_GEMM_TIMING
  mutable long long t_ld = 0, t_wr = 0, t_st = 0;      // TMEM load + wait, wait for the previous store, staging + store issue
#endif
};

// EH = number of epilogue warps per TMEM lane quadrant (2 for the wide tiles: warp (q, h) owns columns [h*BN/2, +BN/2)).
// Two warps per scheduler is what hides the TMEM / shared-memory / ALU latencies of this per-element code; with one warp
// per scheduler the epilogue (12-24k clk per 128x256 tile, measured) was slower than the tile's MMAs.
// Tile context of one epilogue thread (ALU only): which row / sample / column group it works on.
__device__ __forceinline__ void epi_generic_ctx(const GemmKernelParams& p, int m0, int n0, int tile_b, int q, int lane, EpiCtx& c) {
  // flattened: m0 is the global row of the tile; batched: m0 is the row inside sample tile_b
  const int lr = m0 + q * 32 + lane;
  c.row_ok = p.batched ? (lr < p.rows_per_sample) : (lr < p.M);
  c.m = p.batched ? tile_b * p.rows_per_sample + lr : lr;
  c.mm = c.row_ok ? c.m : 0;
  c.rs = (p.a_row_ss != nullptr && c.row_ok) ? rsqrtf(fmaxf(__ldg(p.a_row_ss + c.mm), 1e-12f)) : 1.0f;
  int b = p.batched ? tile_b : c.mm / p.rows_per_sample;
  const int b0 = __shfl_sync(0xffffffffu, b, 0);       // rows grow with the lane: lane 0 valid unless the whole warp is not
  if (!c.row_ok) b = b0;
  c.b = b;
  c.uniform = __all_sync(0xffffffffu, b == b0);
  const int pix = c.mm - b * p.rows_per_sample;
  if (p.group_width <= 0) { c.grp = 0; c.cbase = n0; }
  else {
    c.grp = p.group_shift >= 0 ? (n0 >> p.group_shift) : n0 / p.group_width;     // tiles never straddle groups (checked on the host)
    c.cbase = n0 - c.grp * p.group_width;                                        // column inside the group
  }
  c.sb = p.sbias ? p.sbias + (long long)b * p.ld_sbias : nullptr;
  c.gt = p.gate ? p.gate + (long long)b * p.ld_gate : nullptr;
  c.peep = p.cprev != nullptr && (c.grp == 1 || c.grp == 2);
  const int esz = p.peep16 ? 2 : 4;                   // byte-addressed: the peephole operands are fp32 or fp16
  c.pe = c.peep ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(c.grp == 1 ? p.peep_i : p.peep_f) + (long long)pix * p.ld_peep * esz) : nullptr;
  c.cp = c.peep ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.cprev) + (long long)c.mm * p.ld_cprev * esz) : nullptr;
}

// The same context from a TILE DESCRIPTOR (m0, nt | tb << 16, sample of the tile's first row, first row of the next sample) that
// the CTA computed once for each of its tiles before the main loop.  Deriving it per tile in every epilogue thread -- two integer
// divisions, a shuffle and a vote behind a chain of parameter loads -- sat on the critical path of the epilogue-bound GEMMs:
// ~3 k clk of an 8 k clk tile at K = 500 (scripts/gemm_timeline.py: "wait acc" of the lang_se shape while the MMA thread idles).
constexpr int DESC_MAX = 48;
__device__ __forceinline__ void epi_generic_ctx_desc(const GemmKernelParams& p, const int4 d, int BN, int q, int lane, EpiCtx& c,
                                                     int& m0, int& nt, int& tb) {
  m0 = d.x; nt = d.y & 0xffff; tb = d.y >> 16;
  const int w0 = m0 + q * 32, lr = w0 + lane;
  c.row_ok = p.batched ? (lr < p.rows_per_sample) : (lr < p.M);
  c.m = p.batched ? tb * p.rows_per_sample + lr : lr;
  c.mm = c.row_ok ? c.m : 0;
  c.rs = (p.a_row_ss != nullptr && c.row_ok) ? rsqrtf(fmaxf(__ldg(p.a_row_ss + c.mm), 1e-12f)) : 1.0f;
  const int b = p.batched ? tb : d.z + (lr >= d.w ? 1 : 0);     // d.w = INT_MAX in the last sample: rows past M stay in it
  c.b = b;
  c.uniform = p.batched || !(d.w > w0 && d.w <= w0 + 31);
  const int pix = c.row_ok ? c.mm - b * p.rows_per_sample : 0;
  const int n0 = nt * BN;
  if (p.group_width <= 0) { c.grp = 0; c.cbase = n0; }
  else {
    c.grp = p.group_shift >= 0 ? (n0 >> p.group_shift) : n0 / p.group_width;
    c.cbase = n0 - c.grp * p.group_width;
  }
  c.sb = p.sbias ? p.sbias + (long long)b * p.ld_sbias : nullptr;
  c.gt = p.gate ? p.gate + (long long)b * p.ld_gate : nullptr;
  c.peep = p.cprev != nullptr && (c.grp == 1 || c.grp == 2);
  const int esz = p.peep16 ? 2 : 4;                   // byte-addressed: the peephole operands are fp32 or fp16
  c.pe = c.peep ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(c.grp == 1 ? p.peep_i : p.peep_f) + (long long)pix * p.ld_peep * esz) : nullptr;
  c.cp = c.peep ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.cprev) + (long long)c.mm * p.ld_cprev * esz) : nullptr;
}

// Per-column operands of a tile ("add" = bias + per-sample bias, "mul" = per-sample gate / validity mask), PER columns per lane,
// requested into registers: the kernel issues this for tile i+1 before it works on tile i, so the global latency of these
// loads is never in front of an accumulator that is already waiting (the epilogue-bound case), and writes them to the warp's
// shared-memory slice (epi_generic_commit) when tile i+1 starts.
template <int BN, int EH>
__device__ __forceinline__ void epi_generic_fetch(const GemmKernelParams& p, const EpiCtx& c, int n0, int h, int lane, float4& a, float4& g) {
  constexpr int W = BN / EH;         // columns owned by this warp
  constexpr int PER = W / 32;        // columns staged by each lane (4 or 1)
  a = make_float4(0.f, 0.f, 0.f, 0.f);
  g = a;
  if (!c.uniform) return;
  const int col = h * W + lane * PER;                // column inside the tile
  const int cc = c.cbase + col, n = n0 + col;
  if (PER >= 4) {
    if (cc + 3 < p.group_valid) {    // group_valid % 4 == 0 (host check)
      g = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p.bias) a = ldg4(p.bias + n);
      if (c.sb) { const float4 t = ldg4(c.sb + n); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
      if (c.gt) g = ldg4(c.gt + n);
    }
  } else {
    if (cc < p.group_valid) {
      g.x = 1.f;
      if (p.bias) a.x = __ldg(p.bias + n);
      if (c.sb) a.x += __ldg(c.sb + n);
      if (c.gt) g.x = __ldg(c.gt + n);
    }
  }
}

template <int BN, int EH>
__device__ __forceinline__ void epi_generic_commit(const EpiCtx& c, int lane, float* s_add, float* s_mul, const float4& a, const float4& g) {
  constexpr int PER = BN / EH / 32;
  if (!c.uniform) return;
  if (PER >= 4) {
    *reinterpret_cast<float4*>(s_add + lane * PER) = a;
    *reinterpret_cast<float4*>(s_mul + lane * PER) = g;
  } else {
    s_add[lane] = a.x;
    s_mul[lane] = g.x;
  }
  __syncwarp();
}

// ConvLSTM peepholes of one 32-column chunk (group column cb): 2 x 128 bytes per thread, straight from global / L2.
// P16: the operands are fp16 (inference: cell state and W_ci / W_cf copies in fp16) -- 2 x 64 bytes per thread in four 256-bit
// loads instead of eight; the raw halves stay packed in pe4[0..1] / cp4[0..1] until the math unpacks them.  With 32 different
// 128-byte lines per warp instruction these loads cost L1 tag lookups, not bandwidth: the count of instructions is what matters.
template <bool P16>
__device__ __forceinline__ void epi_peep_load(const EpiCtx& c, int cb, float4 (&pe4)[8], float4 (&cp4)[8]) {
  if (P16) {
    const float* pe = reinterpret_cast<const float*>(reinterpret_cast<const char*>(c.pe) + cb * 2);
    const float* cp = reinterpret_cast<const float*>(reinterpret_cast<const char*>(c.cp) + cb * 2);
#pragma unroll
    for (int j = 0; j < 2; ++j) {                      // 16 halves = 32 bytes per load
      ldg8(pe + j * 8, pe4[2 * j], pe4[2 * j + 1]);
      ldg8(cp + j * 8, cp4[2 * j], cp4[2 * j + 1]);
    }
  } else {
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      ldg8(c.pe + cb + j8 * 8, pe4[2 * j8], pe4[2 * j8 + 1]);
      ldg8(c.cp + cb + j8 * 8, cp4[2 * j8], cp4[2 * j8 + 1]);
    }
  }
}
// element pair (2w, 2w + 1) of a chunk's packed fp16 peephole operands (word w of the raw registers)
__device__ __forceinline__ float2 peep_h2(const float4 (&raw)[8], int w) {
  const float* f = reinterpret_cast<const float*>(raw);
  const uint32_t bits = __float_as_uint(f[w]);
  return __half22float2(*reinterpret_cast<const __half2*>(&bits));
}

// The chunk loop is software-pipelined by one chunk: r (and pe4 / cp4) arrive already REQUESTED -- by the prologue in
// epi_generic_compute or by the previous chunk -- and as soon as this chunk's math has consumed them the next chunk's TMEM load
// (and peephole loads) are issued, so that their latency runs under this chunk's store phase (wait for the staging buffer,
// st.shared, proxy fence, TMA store issue) instead of in front of the next chunk's math.
template <bool MUL, bool SUMS, bool PEEP, bool UNI, bool P16>
__device__ __forceinline__ void epi_chunk(const GemmKernelParams& p, const EpiCtx& c, uint32_t (&r)[32], float4 (&pe4)[8], float4 (&cp4)[8],
                                          bool has_next, uint32_t taddr_next, int nb, int cb,
                                          const float* s_add, const float* s_mul, float& s1, float& s2,
                                          const CUtensorMap* tmOut, uint8_t* stg, int row0, int tb, int lane) {
  // per-column operands of the chunk, fetched before the accumulator so that their latency overlaps the TMEM load; one
  // warp-uniform branch here keeps the math below free of divergence regions (which had pinned every load to its use)
  const bool need_g = MUL || (!PEEP && p.act >= 2);      // the peephole (ConvLSTM gate) GEMM has neither gate nor activation
  const uint32_t sa_u = smem_u32(s_add);
  // UNI (all 32 rows of the warp in one sample -- always, unless rows_per_sample % 32 != 0) is a template parameter: as a run-time
  // test it put a divergence region around each of the 16 operand loads of a chunk
  auto load_add = [&](int j4) -> float4 {
    if (UNI) return lds4(sa_u + j4 * 16);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);             // rows of different samples in one warp (odd shapes only): per-thread loads
    if (cb + j4 * 4 + 3 < p.group_valid) {
      if (p.bias) a = ldg4(p.bias + nb + j4 * 4);
      if (c.sb) { const float4 t = ldg4(c.sb + nb + j4 * 4); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
    }
    return a;
  };
  float4 a4[8];
  if (!PEEP) {           // (the peephole variant already holds 64 operand registers: it fetches these inside the math loop)
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) a4[j4] = load_add(j4);
  }
#ifdef CMPC_GEMM_TIMING
  const long long tl0 = GT_NOW();
#endif
  tmem_wait_ld();                          // the load of this chunk was issued one chunk ago
#ifdef CMPC_GEMM_TIMING
  c.t_ld += GT_NOW() - tl0;
#endif
  float v[32];
  const float lo = (p.act == 1) ? 0.f : -INFINITY;      // relu as a branch-free max
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 a = PEEP ? load_add(j4) : a4[j4];
    // acc * rs + add: rs = 1 unless the A rows carry a deferred l2_normalize (exact then: fma(x, 1, a) = x + a)
    float x0 = fmaf(__uint_as_float(r[j4 * 4 + 0]), c.rs, a.x), x1 = fmaf(__uint_as_float(r[j4 * 4 + 1]), c.rs, a.y);
    float x2 = fmaf(__uint_as_float(r[j4 * 4 + 2]), c.rs, a.z), x3 = fmaf(__uint_as_float(r[j4 * 4 + 3]), c.rs, a.w);
    if (PEEP && P16) {
      const float2 pa = peep_h2(pe4, 2 * j4), pb = peep_h2(pe4, 2 * j4 + 1), ca = peep_h2(cp4, 2 * j4), cb2 = peep_h2(cp4, 2 * j4 + 1);
      x0 = fmaf(pa.x, ca.x, x0); x1 = fmaf(pa.y, ca.y, x1);
      x2 = fmaf(pb.x, cb2.x, x2); x3 = fmaf(pb.y, cb2.y, x3);
    } else if (PEEP) {
      x0 = fmaf(pe4[j4].x, cp4[j4].x, x0); x1 = fmaf(pe4[j4].y, cp4[j4].y, x1);
      x2 = fmaf(pe4[j4].z, cp4[j4].z, x2); x3 = fmaf(pe4[j4].w, cp4[j4].w, x3);
    }
    v[j4 * 4 + 0] = fmaxf(x0, lo); v[j4 * 4 + 1] = fmaxf(x1, lo); v[j4 * 4 + 2] = fmaxf(x2, lo); v[j4 * 4 + 3] = fmaxf(x3, lo);
  }
  if (!PEEP && p.act >= 2) {             // tanh / sigmoid: only the tiny per-sentence GEMMs use these
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (p.act == 2) ? tanh_acc(v[j]) : sigmoid_acc(v[j]);
  }
  if (need_g) {                          // per-sample gate and / or validity mask (after the activation); a4 is dead by now
    float4 g4[8];
    if (UNI) {
      const uint32_t sm = smem_u32(s_mul);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) g4[j4] = lds4(sm + j4 * 16);
    } else {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        g4[j4] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cb + j4 * 4 + 3 < p.group_valid) g4[j4] = c.gt ? ldg4(c.gt + nb + j4 * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
    }
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      v[j4 * 4 + 0] *= g4[j4].x; v[j4 * 4 + 1] *= g4[j4].y; v[j4 * 4 + 2] *= g4[j4].z; v[j4 * 4 + 3] *= g4[j4].w;
    }
  }
  if (has_next) {                          // r / pe4 / cp4 are dead: request the next chunk (warp-uniform branch)
    tmem_ld_x32(taddr_next, r);
    if (PEEP && !p.out_fp32) epi_peep_load<P16>(c, cb + 32, pe4, cp4);     // (fp32 output keeps v[] live longer: it loads after its stores)
  }
  if (SUMS) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float x0 = v[j4 * 4], x1 = v[j4 * 4 + 1], x2 = v[j4 * 4 + 2], x3 = v[j4 * 4 + 3];
      s1 += (x0 + x1) + (x2 + x3);
      s2 += (x0 * x0 + x1 * x1) + (x2 * x2 + x3 * x3);
    }
  }
  // Store: stage the warp's 32 rows x 64 bytes in shared memory (64B swizzle) and let the TMA write them: a per-thread-row
  // st.global touches 32 different 128-byte lines per instruction and was measured to cost 7-17k clk per 128x256 tile,
  // more than the tile's MMAs.  Rows beyond the tensor (or, batched, beyond the sample) and columns >= ldo are clipped.
  const int sw = (lane >> 1) & 3;                     // 64B swizzle: 16-byte chunk k of row r sits at k ^ ((r >> 1) & 3)
  const uint32_t srow_s = smem_u32(stg) + lane * 64;
  if (p.out_fp32) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {                  // 16 fp32 columns = 64 bytes per pass
      if (lane == 0) tma_store_wait_read();           // the previous store has finished reading the staging buffer
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        sts4f(srow_s + ((k ^ sw) << 4), v[hh * 16 + k * 4], v[hh * 16 + k * 4 + 1], v[hh * 16 + k * 4 + 2], v[hh * 16 + k * 4 + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && nb + hh * 16 < p.ldo) {
        tma_store_3d(tmOut, stg, nb + hh * 16, row0, tb);
        tma_store_commit();
      }
    }
    if (PEEP && has_next) epi_peep_load<P16>(c, cb + 32, pe4, cp4);
  } else {
#ifdef CMPC_GEMM_TIMING
    const long long tw0 = GT_NOW();
#endif
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
#ifdef CMPC_GEMM_TIMING
    const long long tw1 = GT_NOW();
    c.t_wr += tw1 - tw0;
#endif
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4 u;
      __half2 h0 = __floats2half2_rn(v[k * 8 + 0], v[k * 8 + 1]);
      __half2 h1 = __floats2half2_rn(v[k * 8 + 2], v[k * 8 + 3]);
      __half2 h2 = __floats2half2_rn(v[k * 8 + 4], v[k * 8 + 5]);
      __half2 h3 = __floats2half2_rn(v[k * 8 + 6], v[k * 8 + 7]);
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2);
      u.w = *reinterpret_cast<uint32_t*>(&h3);
      sts4(srow_s + ((k ^ sw) << 4), u);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(tmOut, stg, nb, row0, tb);
      tma_store_commit();
    }
#ifdef CMPC_GEMM_TIMING
    c.t_st += GT_NOW() - tw1;
#endif
  }
}

// the chunk loop of one warp for one compile-time feature set (dispatched once per tile, so that the registers carried
// around the loop -- r, and pe4 / cp4 only in the peephole variant -- are those of this variant alone)
template <int BN, int EH, bool MUL, bool SUMS, bool PEEP, bool UNI, bool P16>
__device__ __forceinline__ void epi_generic_loop(const GemmKernelParams& p, uint32_t tmem_acc, int n0, int q, int h, int lane,
                                                 const float* s_add, const float* s_mul, const EpiCtx& c, float& s1, float& s2,
                                                 const CUtensorMap* tmOut, uint8_t* stg, int row0, int tb) {
  constexpr int W = BN / EH;
  int nch = 0;                              // chunks left of ldo (warp-uniform; later chunks are further right)
  while (nch < W / 32 && n0 + h * W + nch * 32 < p.ldo) ++nch;
  if (nch == 0) return;
  uint32_t r[32];
  float4 pe4[8], cp4[8];
  const uint32_t tbase = tmem_acc + (uint32_t(q * 32) << 16) + h * W;
  tmem_ld_x32(tbase, r);                    // prologue of the pipeline: request chunk 0
  if (PEEP) epi_peep_load<P16>(c, c.cbase + h * W, pe4, cp4);
#pragma unroll 1
  for (int ch = 0; ch < nch; ++ch) {
    const int col = h * W + ch * 32;      // column inside the tile
    epi_chunk<MUL, SUMS, PEEP, UNI, P16>(p, c, r, pe4, cp4, ch + 1 < nch, tbase + (ch + 1) * 32, n0 + col, c.cbase + col, s_add + ch * 32,
                               s_mul + ch * 32, s1, s2, tmOut, stg, row0, tb, lane);
  }
}

template <int BN, int EH>
__device__ __forceinline__ void epi_generic_compute(const GemmKernelParams& p, uint32_t tmem_acc, int n0, int q, int h, int lane,
                                                    const float* s_add, const float* s_mul, const EpiCtx& c,
                                                    const CUtensorMap* tmOut, uint8_t* stg, int row0, int tb) {
  float s1 = 0.f, s2 = 0.f;
  const bool mul = p.gate != nullptr || p.act >= 2 || !c.uniform;     // act >= 2 applies the validity mask through mul
  const bool sums = p.stats != nullptr || p.row_sumsq != nullptr;
  // with a gate-less relu/identity epilogue the validity mask is implicit: add = 0 and the accumulator is 0
#define CMPC_EPI_LOOP(M_, S_, P_, U_, H_) epi_generic_loop<BN, EH, M_, S_, P_, U_, H_>(p, tmem_acc, n0, q, h, lane, s_add, s_mul, c, s1, s2, tmOut, stg, row0, tb)
  if (!c.uniform) {     // a warp's rows straddle two samples (odd shapes only): per-thread operand loads, one generic variant
    if (c.peep) { if (p.peep16) CMPC_EPI_LOOP(false, true, true, false, true); else CMPC_EPI_LOOP(false, true, true, false, false); }
    else        CMPC_EPI_LOOP(true, true, false, false, false);
  }
  else if (c.peep) { if (p.peep16) CMPC_EPI_LOOP(false, true, true, true, true); else CMPC_EPI_LOOP(false, true, true, true, false); }
  else if (mul)    { if (sums) CMPC_EPI_LOOP(true, true, false, true, false); else CMPC_EPI_LOOP(true, false, false, true, false); }
  else             { if (sums) CMPC_EPI_LOOP(false, true, false, true, false); else CMPC_EPI_LOOP(false, false, false, true, false); }
#undef CMPC_EPI_LOOP
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

// MUTAN: BN = 240 = 5 heads x 48 channels.  out = tanh(sum_k tanh(acc_k + bias_k) * lang_k).  The five inner tanh use
// tanh.approx (one MUFU op, 2^-11 relative = one fp16 rounding; the epilogue was MUFU-bound with the two-op version), the
// outer one the accurate form.  Oracle emulation: logits move by < 2e-4, mask agreement unchanged.
// Warp (q, h) owns channels [24 h, 24 h + 24) of the tile's 48 = three groups of 8.
__device__ __forceinline__ void epi_mutan_prefetch(const GemmKernelParams& p, int m0, int jchunk, int q, int h, int lane,
                                                   float* s_bias, float* s_lang, EpiCtx& c) {
  c.m = m0 + q * 32 + lane;
  c.row_ok = c.m < p.M;
  c.mm = c.row_ok ? c.m : 0;
  int b = c.mm / p.rows_per_sample;
  const int b0 = __shfl_sync(0xffffffffu, b, 0);
  if (!c.row_ok) b = b0;
  c.b = b;
  c.uniform = __all_sync(0xffffffffu, b == b0);
  if (c.uniform && lane < 30) {
    const float* lang = p.lang + (long long)b * p.lang_bstride;
    const int k = lane / 6, i4 = lane - k * 6;           // [5 heads][6 float4] = this warp's 24 channels
    const int ch = jchunk * 48 + h * 24 + i4 * 4;
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), ll = bb;
    if (ch + 3 < p.C) {
      bb = ldg4(p.mbias + (long long)k * p.ld_mbias + ch);
      ll = ldg4(lang + (long long)k * p.ld_lang + ch);
    }
    *reinterpret_cast<float4*>(s_bias + k * 24 + i4 * 4) = bb;
    *reinterpret_cast<float4*>(s_lang + k * 24 + i4 * 4) = ll;
  }
  __syncwarp();
}

__device__ __forceinline__ void epi_mutan_compute(const GemmKernelParams& p, uint32_t tmem_acc, int jchunk, int q, int h, int lane,
                                                  const float* s_bias, const float* s_lang, const EpiCtx& c,
                                                  const CUtensorMap* tmOut, uint8_t* stg, int row0) {
  const float* lang = p.lang + (long long)c.b * p.lang_bstride;
  float ss = 0.f;
  // l2_normalize of the lateral features (CMPC_model.py:109-113) folded in: x_hat . W = (x . W) * rsqrt(max(|x|^2, 1e-12))
  const float rsc = (p.a_row_ss && c.row_ok) ? rsqrtf(fmaxf(__ldg(p.a_row_ss + c.mm), 1e-12f)) : 1.0f;
  const int cw = jchunk * 48 + h * 24;                  // first channel of this warp
  if (cw < p.ldo) {
    uint32_t r[5][24];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      tmem_ld_x16(tmem_acc + (uint32_t(q * 32) << 16) + k * 48 + h * 24, *reinterpret_cast<uint32_t(*)[16]>(&r[k][0]));
      tmem_ld_x8(tmem_acc + (uint32_t(q * 32) << 16) + k * 48 + h * 24 + 16, *reinterpret_cast<uint32_t(*)[8]>(&r[k][16]));
    }
    tmem_wait_ld();
    float acc[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) acc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
      for (int i4 = 0; i4 < 6; ++i4) {
        float4 bb, ll;
        if (c.uniform) {
          bb = *reinterpret_cast<const float4*>(s_bias + k * 24 + i4 * 4);
          ll = *reinterpret_cast<const float4*>(s_lang + k * 24 + i4 * 4);
        } else {
          bb = make_float4(0.f, 0.f, 0.f, 0.f); ll = bb;
          if (cw + i4 * 4 + 3 < p.C) {
            bb = ldg4(p.mbias + (long long)k * p.ld_mbias + cw + i4 * 4);
            ll = ldg4(lang + (long long)k * p.ld_lang + cw + i4 * 4);
          }
        }
        acc[i4 * 4 + 0] = fmaf(tanh_fast(fmaf(__uint_as_float(r[k][i4 * 4 + 0]), rsc, bb.x)), ll.x, acc[i4 * 4 + 0]);
        acc[i4 * 4 + 1] = fmaf(tanh_fast(fmaf(__uint_as_float(r[k][i4 * 4 + 1]), rsc, bb.y)), ll.y, acc[i4 * 4 + 1]);
        acc[i4 * 4 + 2] = fmaf(tanh_fast(fmaf(__uint_as_float(r[k][i4 * 4 + 2]), rsc, bb.z)), ll.z, acc[i4 * 4 + 2]);
        acc[i4 * 4 + 3] = fmaf(tanh_fast(fmaf(__uint_as_float(r[k][i4 * 4 + 3]), rsc, bb.w)), ll.w, acc[i4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      acc[i] = tanh_acc(acc[i]);     // columns >= C have bias = lang = 0 and zero weights -> tanh(0) = 0
      ss += acc[i] * acc[i];
    }
    // stage 32 rows x 24 fp32 (96-byte rows, no swizzle; 48-byte rows for an fp16 map) and TMA-store them; out-of-range rows /
    // columns are clipped
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
    if (p.out_fp32) {
#pragma unroll
      for (int i4 = 0; i4 < 6; ++i4)
        *reinterpret_cast<float4*>(stg + lane * 96 + i4 * 16) = make_float4(acc[i4 * 4], acc[i4 * 4 + 1], acc[i4 * 4 + 2], acc[i4 * 4 + 3]);
    } else {
#pragma unroll
      for (int i8 = 0; i8 < 3; ++i8) {
        uint4 u;
        __half2 h0 = __floats2half2_rn(acc[i8 * 8 + 0], acc[i8 * 8 + 1]), h1 = __floats2half2_rn(acc[i8 * 8 + 2], acc[i8 * 8 + 3]);
        __half2 h2 = __floats2half2_rn(acc[i8 * 8 + 4], acc[i8 * 8 + 5]), h3 = __floats2half2_rn(acc[i8 * 8 + 6], acc[i8 * 8 + 7]);
        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(stg + lane * 48 + i8 * 16) = u;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(tmOut, stg, cw, row0, 0);
      tma_store_commit();
    }
  }
  if (p.row_sumsq && c.row_ok) atomicAdd(p.row_sumsq + c.m, ss);
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// CL = CTAs per cluster.  With CL = 2 the two CTAs work on vertically adjacent row tiles of the SAME column tile, so the
// weight tile W[n0:n0+BN, k] is common: each CTA fetches half of its rows and TMA-multicasts them into both CTAs' shared
// memory.  This cuts L2 -> SM traffic per tile from (128 + BN) to (128 + BN/2) rows per k-step; measured on B200 the
// K >= 1000 GEMMs of the head were pinned at one SM's share of L2 bandwidth (~42 B/clk) before this change.
// ---------------------------------------------------------------------------------------------
// MUTAN backward epilogue (same tile / warp layout as the forward one): the GEMM recomputes acc_k = x . Wv_k for the five
// heads of 24 channels, and the epilogue turns the gradient ds of their gated sum into
//   t_k = tanh(acc_k * rsc + b_k);   d pre_k = ds * lang_k * (1 - t_k^2)
//   out[m, packed column] = fp16(d pre_k * rsc)      (operand of the dgrad / wgrad GEMMs; rsc = the folded lateral l2norm)
//   dlang_k[b, c] += sum_rows ds * t_k;              dbias_k[c] += sum_rows d pre_k
// Column sums over the 32 rows of a warp use a 31-shuffle butterfly per 32 columns (each lane ends with one column).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < n / 2; ++j) {
      const float send = up ? v[j] : v[j + n / 2];
      const float keep = up ? v[j + n / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];      // column `lane`
}

__device__ __forceinline__ void epi_mutan_bwd_compute(const GemmKernelParams& p, uint32_t tmem_acc, int jchunk, int q, int h, int lane,
                                                      const float* s_bias, const float* s_lang, const EpiCtx& c,
                                                      const CUtensorMap* tmOut, uint8_t* stg, int row0) {
  const float* lang = p.lang + (long long)c.b * p.lang_bstride;
  const float rsc = (p.a_row_ss && c.row_ok) ? rsqrtf(fmaxf(__ldg(p.a_row_ss + c.mm), 1e-12f)) : 1.0f;
  const int cw = jchunk * 48 + h * 24;                  // first channel of this warp
  if (cw >= p.C) return;
  float dsv[24];
#pragma unroll
  for (int i4 = 0; i4 < 6; ++i4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c.row_ok && cw + i4 * 4 + 3 < p.C) t = ldg4(p.ds + (long long)c.mm * p.ld_ds + cw + i4 * 4);
    dsv[i4 * 4] = t.x; dsv[i4 * 4 + 1] = t.y; dsv[i4 * 4 + 2] = t.z; dsv[i4 * 4 + 3] = t.w;
  }
#pragma unroll 1
  for (int k = 0; k < 5; ++k) {
    uint32_t r[24];
    tmem_ld_x16(tmem_acc + (uint32_t(q * 32) << 16) + k * 48 + h * 24, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
    tmem_ld_x8(tmem_acc + (uint32_t(q * 32) << 16) + k * 48 + h * 24 + 16, *reinterpret_cast<uint32_t(*)[8]>(&r[16]));
    tmem_wait_ld();
    float o[24], sl[32], sb[32];     // sl / sb: this lane's contributions to the column sums of head k (24 columns, padded)
#pragma unroll
    for (int i = 24; i < 32; ++i) { sl[i] = 0.f; sb[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      float bb, ll;
      if (c.uniform) { bb = s_bias[k * 24 + i]; ll = s_lang[k * 24 + i]; }
      else {
        const bool ok = cw + (i & ~3) + 3 < p.C;
        bb = ok ? __ldg(p.mbias + (long long)k * p.ld_mbias + cw + i) : 0.f;
        ll = ok ? __ldg(lang + (long long)k * p.ld_lang + cw + i) : 0.f;
      }
      const float t = tanh_fast(fmaf(__uint_as_float(r[i]), rsc, bb));
      const float dp = dsv[i] * ll * (1.f - t * t);
      o[i] = dp * rsc;
      sl[i] = dsv[i] * t;
      sb[i] = dp;
    }
    // stage 32 rows x 24 fp16 (48-byte rows, no swizzle) in one of two 1.5 KB slots and TMA-store them
    uint8_t* slot = stg + (k & 1) * 1536;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int i8 = 0; i8 < 3; ++i8) {
      uint4 u;
      __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) hh[e] = __floats2half2_rn(o[i8 * 8 + 2 * e], o[i8 * 8 + 2 * e + 1]);
      *reinterpret_cast<uint4*>(slot + lane * 48 + i8 * 16) = u;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(tmOut, slot, jchunk * 240 + k * 48 + h * 24, row0, 0);
      tma_store_commit();
    }
    // column sums over the warp's 32 rows
    if (c.uniform) {
      const float cl = colsum32(sl, lane), cb = colsum32(sb, lane);
      if (lane < 24 && cw + lane < p.C) {
        if (p.dlang) atomicAdd(p.dlang + (long long)c.b * p.dlang_bstride + (long long)k * p.ld_lang + cw + lane, cl);
        if (p.dmbias) atomicAdd(p.dmbias + (long long)k * p.ld_mbias + cw + lane, cb);
      }
    } else if (c.row_ok) {
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        if (cw + i < p.C) {
          if (p.dlang) atomicAdd(p.dlang + (long long)c.b * p.dlang_bstride + (long long)k * p.ld_lang + cw + i, sl[i]);
          if (p.dmbias) atomicAdd(p.dmbias + (long long)k * p.ld_mbias + cw + i, sb[i]);
        }
      }
    }
  }
}

// TWO (needs CL = 2) replaces the multicast scheme by the 2-SM MMA: the leader CTA issues tcgen05.mma.cta_group::2 with
// M = 256 over both CTAs' A tiles and weight halves (128-clk dispatches instead of 172, half the B-operand smem traffic,
// no duplicated weight tile).  Both producers complete their bytes on the LEADER's full barrier; the leader's commits are
// multicast to both CTAs' empty / tmem_full barriers; both CTAs' epilogue warps arrive on the leader's tmem_empty barrier.
template <int BN, int EPI, int CL, bool TWO>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut, const GemmKernelParams p) {
  static_assert(!TWO || CL == 2, "the 2-SM MMA needs a CTA pair");
  using Cfg = SmemCfg<BN, TWO>;
  constexpr int STAGES = Cfg::NSTAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
  // work unit = CL vertically adjacent row tiles x one column tile; p.m_tiles is already padded to a multiple of CL
  const int num_units = (p.m_tiles / CL) * p.n_tiles * (p.batched ? p.batch : 1);   // batched: m_tiles is per sample
  const int kt_total = p.kt1 + p.kt2;
  constexpr uint16_t kAll = (1u << CL) - 1;
  constexpr int EH = (BN >= 64) ? 2 : 1;     // epilogue warps per lane quadrant

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    if (p.kt2 > 0) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], TWO ? 1 : CL);   // multicast scheme: every CTA of the cluster must have drained the stage
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], TWO ? 8 * EH : 4 * EH);   // 2-SM: the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (TWO) { tmem_alloc_2sm(tmem_ptr, TMEM_COLS); tmem_relinquish_2sm(); }
    else     { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();        // peers' barriers are initialised before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // unit -> (row tile of this CTA, column tile, sample)
  auto decode = [&](int unit, int& mt_local, int& nt, int& tb) {
    const int mp = p.ntile_shift >= 0 ? (unit >> p.ntile_shift) : unit / p.n_tiles;   // index of the row-tile group
    nt = unit - mp * p.n_tiles;
    const int groups_per_sample = p.m_tiles / CL;
    tb = p.batched ? mp / groups_per_sample : 0;
    mt_local = (p.batched ? (mp - tb * groups_per_sample) : mp) * CL + rank;   // batched: tile inside the sample
  };

  if (p.pdl) pdl_launch_dependents();      // the next kernel's CTAs may take this SM as soon as this CTA has left it
  // generic epilogue: descriptors of this CTA's tiles (see epi_generic_ctx_desc), one thread per tile
  int4* desc_tab = reinterpret_cast<int4*>(smem + Cfg::DESC_OFF);
  const int n_my = cluster_id < num_units ? (num_units - 1 - cluster_id) / num_clusters + 1 : 0;
  const bool use_desc = EPI == EPI_GENERIC && n_my <= DESC_MAX && p.n_tiles < 65536 && (p.batched ? p.batch < 32768 : p.rows_per_sample >= BLOCK_M);
  if (use_desc) {
    if ((int)threadIdx.x < n_my) {
      int mtl, nt, tb;
      decode(cluster_id + (int)threadIdx.x * num_clusters, mtl, nt, tb);
      const int m0 = mtl * BLOCK_M;
      int bf = 0, bnd = 0x7fffffff;
      if (!p.batched) {
        const int last = (p.M - 1) / p.rows_per_sample;
        bf = min(m0 / p.rows_per_sample, last);
        if (bf < last) bnd = (bf + 1) * p.rows_per_sample;     // rows_per_sample >= BLOCK_M: at most one boundary inside a tile
      }
      desc_tab[threadIdx.x] = make_int4(m0, nt | (tb << 16), bf, bnd);
    }
    __syncthreads();
  }

  if (p.pdl) pdl_wait();                    // everything above touched no global memory: the kernel in front may still be running

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        int mtl, nt, tb;
        decode(unit, mtl, nt, tb);
        const int n0 = nt * BN;
        const int m0 = (p.batched ? tb * p.rows_per_sample : 0) + mtl * BLOCK_M;   // global row of the A tile
        for (int kt = 0; kt < kt_total; ++kt) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (TWO) {
            // both CTAs' bytes land on the leader's barrier; the leader arms it for the pair
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
            const uint32_t lead_full = mapa_u32(&full_bar[s], 0);
            if (kt < p.kt1) tma_load_2d_2sm(sa, &tmA1, lead_full, kt * BLOCK_K, m0);
            else            tma_load_2d_2sm(sa, &tmA2, lead_full, (kt - p.kt1) * BLOCK_K, m0);
            tma_load_2d_2sm(sb, &tmW, lead_full, kt * BLOCK_K, n0 + rank * (BN / 2));     // my half of the weight rows
            if (++s == STAGES) { s = 0; ph ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          if (kt < p.kt1) tma_load_2d(sa, &tmA1, &full_bar[s], kt * BLOCK_K, m0);
          else            tma_load_2d(sa, &tmA2, &full_bar[s], (kt - p.kt1) * BLOCK_K, m0);
          if (CL > 1) {
            // my share of the weight tile: rows [n0 + rank*BN/CL, +BN/CL), delivered to every CTA of the cluster
            tma_load_2d_mc(sb + rank * (Cfg::B_BYTES / CL), &tmW, &full_bar[s], kt * BLOCK_K, n0 + rank * (BN / CL), kAll);
          } else if (p.batched) {
            tma_load_3d(sb, &tmW, &full_bar[s], kt * BLOCK_K, n0, tb);
          } else {
            tma_load_2d(sb, &tmW, &full_bar[s], kt * BLOCK_K, n0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && (!TWO || rank == 0)) {
      constexpr uint32_t idesc = make_idesc_f16(TWO ? 2 * BLOCK_M : BLOCK_M, BN, /*fp16*/ 0, /*A K-major*/ 0, /*B K-major*/ 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
#ifdef CMPC_GEMM_TIMING
      long long t_begin = GT_NOW(), w_full = 0, w_te = 0;
#endif
      for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
#ifdef CMPC_GEMM_TIMING
        long long t0 = GT_NOW();
#endif
        mbar_wait(&tmem_empty[as], aph ^ 1);
#ifdef CMPC_GEMM_TIMING
        w_te += GT_NOW() - t0;
#endif
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
        for (int kt = 0; kt < kt_total; ++kt) {
#ifdef CMPC_GEMM_TIMING
          long long t1 = GT_NOW();
#endif
          mbar_wait(&full_bar[s], ph);
#ifdef CMPC_GEMM_TIMING
          w_full += GT_NOW() - t1;
#endif
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = make_smem_desc(sa, 16, 1024, 2);
          const uint64_t db = make_smem_desc(sb, 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in the address field
            if (TWO) umma_f16_ss_2sm(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kt | k) != 0 ? 1u : 0u);
            else     umma_f16_ss(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kt | k) != 0 ? 1u : 0u);
          }
          if (TWO)         umma_commit_2sm_mc(&empty_bar[s], kAll);
          else if (CL > 1) umma_commit_mc(&empty_bar[s], kAll);
          else             umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (TWO) umma_commit_2sm_mc(&tmem_full[as], kAll);
        else     umma_commit(&tmem_full[as]);
      }
#ifdef CMPC_GEMM_TIMING
      if (p.dbg) { long long* d = p.dbg + blockIdx.x * 8; d[0] = GT_NOW() - t_begin; d[2] = w_te; d[4] = it; }
#endif
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: warp (q, h) = rows of TMEM lane quadrant q, column half h =====================
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    if (h < EH) {
      int it = 0;
#ifdef CMPC_GEMM_TIMING
      long long e_wait = 0, e_comp = 0, e_ld = 0, e_wr = 0, e_st = 0;
#endif
      float* s_add = reinterpret_cast<float*>(smem + Cfg::EPI_OFF) + (warp - 4) * Cfg::EPI_WARP_FLOATS;
      float* s_mul = s_add + 128;
      uint8_t* stg = smem + Cfg::STG_OFF + (warp - 4) * Cfg::STG_WARP_BYTES;
      float4 pf_a = make_float4(0.f, 0.f, 0.f, 0.f), pf_g = pf_a;     // this lane's per-column operands of the NEXT tile
      if (EPI == EPI_GENERIC && cluster_id < num_units) {
        int mtl, nt, tb;
        EpiCtx nctx;
        if (use_desc) { int m0n; epi_generic_ctx_desc(p, desc_tab[0], BN, q, lane, nctx, m0n, nt, tb); }
        else { decode(cluster_id, mtl, nt, tb); epi_generic_ctx(p, mtl * BLOCK_M, nt * BN, tb, q, lane, nctx); }
        epi_generic_fetch<BN, EH>(p, nctx, nt * BN, h, lane, pf_a, pf_g);
      }
      for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        int mtl, nt, tb, m0;
#ifdef CMPC_GEMM_TIMING
        long long e0 = GT_NOW();
#endif
        EpiCtx ctx;
        if (EPI == EPI_GENERIC) {
          if (use_desc) epi_generic_ctx_desc(p, desc_tab[it], BN, q, lane, ctx, m0, nt, tb);
          else { decode(unit, mtl, nt, tb); m0 = mtl * BLOCK_M; epi_generic_ctx(p, m0, nt * BN, tb, q, lane, ctx); }
          epi_generic_commit<BN, EH>(ctx, lane, s_add, s_mul, pf_a, pf_g);      // requested one tile ago
          if (unit + num_clusters < num_units) {                                // request the next tile's now
            int mtl2, nt2, tb2, m02;
            EpiCtx nctx;
            if (use_desc) epi_generic_ctx_desc(p, desc_tab[it + 1], BN, q, lane, nctx, m02, nt2, tb2);
            else { decode(unit + num_clusters, mtl2, nt2, tb2); epi_generic_ctx(p, mtl2 * BLOCK_M, nt2 * BN, tb2, q, lane, nctx); }
            epi_generic_fetch<BN, EH>(p, nctx, nt2 * BN, h, lane, pf_a, pf_g);
          }
        } else {
          decode(unit, mtl, nt, tb);
          m0 = mtl * BLOCK_M;              // flattened: global row; batched: row inside sample tb
          epi_mutan_prefetch(p, m0, nt, q, h, lane, s_add, s_mul, ctx);
        }
        mbar_wait(&tmem_full[as], aph);
#ifdef CMPC_GEMM_TIMING
        long long e1 = GT_NOW(); e_wait += e1 - e0;
#endif
        tc_fence_after();
        const uint32_t acc = tmem_base + as * ACC_STRIDE;
        const int row0 = m0 + q * 32;     // first row of this warp (global, or inside sample tb when batched)
        if (EPI == EPI_GENERIC)    epi_generic_compute<BN, EH>(p, acc, nt * BN, q, h, lane, s_add, s_mul, ctx, &tmOut, stg, row0, tb);
        else if (EPI == EPI_MUTAN) epi_mutan_compute(p, acc, nt, q, h, lane, s_add, s_mul, ctx, &tmOut, stg, row0);
        else                       epi_mutan_bwd_compute(p, acc, nt, q, h, lane, s_add, s_mul, ctx, &tmOut, stg, row0);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (TWO) mbar_arrive_remote_slot(mapa_u32(&tmem_empty[as], 0));   // the leader's MMA warp waits for both CTAs
          else     mbar_arrive(&tmem_empty[as]);
        }
#ifdef CMPC_GEMM_TIMING
        e_comp += GT_NOW() - e1;
        e_ld += ctx.t_ld; e_wr += ctx.t_wr; e_st += ctx.t_st;
#endif
      }
      if (lane == 0) tma_store_wait_all();     // staging smem must outlive the last TMA store
#ifdef CMPC_GEMM_TIMING
      if (p.dbg && warp == 4 && lane == 0) { long long* d = p.dbg + blockIdx.x * 8; d[5] = e_wait; d[6] = e_comp; d[3] = e_ld; d[7] = e_wr; d[1] = e_st; }
#endif
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();        // no CTA exits while a peer may still signal its barriers or fill its smem
  if (warp == 2) {
    tc_fence_after();
    if (TWO) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else     tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
template <int BN, int EPI, int CL, bool TWO = false>
static int launch_gemm(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w, const CUtensorMap& o, GemmKernelParams p,
                       cudaStream_t stream) {
  using Cfg = SmemCfg<BN, TWO>;
  static unsigned long long configured = 0;
  auto kern = gemm_tc_kernel<BN, EPI, CL, TWO>;
  if (first_use_on_device(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::TOTAL);
    CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "cudaFuncSetAttribute(smem=%d): %s", Cfg::TOTAL, cudaGetErrorString(e));
  }
  p.m_tiles = (p.m_tiles + CL - 1) / CL * CL;     // whole clusters; the padding tile's rows are out of range everywhere
#ifdef CMPC_GEMM_TIMING
  p.dbg = g_gemm_dbg;
#endif
  const int units = (p.m_tiles / CL) * p.n_tiles * (p.batched ? p.batch : 1);
  const int max_clusters = num_sms() / CL;
  const int clusters = units < max_clusters ? units : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  p.pdl = pdl_enabled() ? 1 : 0;
  if (p.pdl) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a1, a2, w, o, p);
  CMPC_REQUIRE(e == cudaSuccess, CMPC_ERR_LAUNCH, "gemm_tc_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("gemm_tc_kernel");
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// clustered GEMMs use the 2-SM MMA; cmpc_gemm_set_mode(1) selects the older multicast scheme (A/B measurements only)
static bool g_two_sm = true;

}  // namespace cmpc

using namespace cmpc;

#ifdef CMPC_GEMM_TIMING
extern "C" void cmpc_gemm_set_debug(long long* buf) { cmpc::g_gemm_dbg = buf; }
#endif

extern "C" void cmpc_gemm_set_mode(int mode) { cmpc::g_two_sm = (mode != 1); }

extern "C" int cmpc_gemm_f16(const cmpc_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CMPC_REQUIRE(a != nullptr, CMPC_ERR_ARG, "cmpc_gemm_f16: null args");
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(a->a1 && a->w && a->out, CMPC_ERR_ARG, "cmpc_gemm_f16: null operand");
  CMPC_REQUIRE(a->m > 0 && a->n > 0 && a->k1 > 0 && a->k2 >= 0, CMPC_ERR_ARG, "cmpc_gemm_f16: bad shape m=%d n=%d k1=%d k2=%d",
               a->m, a->n, a->k1, a->k2);
  CMPC_REQUIRE(a->rows_per_sample >= 1, CMPC_ERR_ARG, "cmpc_gemm_f16: rows_per_sample must be >= 1");
  CMPC_REQUIRE(a->row_scale == nullptr, CMPC_ERR_ARG, "cmpc_gemm_f16: row_scale is reserved and must be NULL");
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
  {
    CMPC_REQUIRE(a->peep_i && a->peep_f && a->cprev && gw > 0, CMPC_ERR_ARG, "cmpc_gemm_f16: peepholes need peep_i, peep_f, cprev and groups");
    CMPC_REQUIRE(a->ld_peep % (a->peep_f16 ? 16 : 8) == 0 && a->ld_cprev % (a->peep_f16 ? 16 : 8) == 0 && gw % 32 == 0 &&
                 ((reinterpret_cast<uintptr_t>(a->peep_i) | reinterpret_cast<uintptr_t>(a->peep_f) | reinterpret_cast<uintptr_t>(a->cprev)) & 31) == 0,
                 CMPC_ERR_ALIGN, "cmpc_gemm_f16: peep_i / peep_f / cprev must be 32-byte aligned with ld %% 8 == 0 (256-bit loads)");
  }
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
  // wide shared-weight GEMMs run as 2-CTA clusters (multicast W halves); per-sample / skinny ones stay single-CTA
  const bool clustered = !batched && !narrow && ceil_div(a->m, BLOCK_M) >= 2;
  if (batched)
    rc = make_tmap_3d(&tW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->w, w_inner, w_rows, batch, a->ldw * 2, a->w_batch_stride * 2, BLOCK_K, bn);
  else
    rc = make_tmap_2d(&tW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->w, w_inner, w_rows, a->ldw * 2, BLOCK_K, clustered ? bn / 2 : bn);
  if (rc) return rc;

  GemmKernelParams p{};
  p.M = a->m; p.N = a->n; p.kt1 = kt1; p.kt2 = kt2;
  p.batched = batched ? 1 : 0; p.batch = batch;
  p.m_tiles = batched ? ceil_div(a->rows_per_sample, BLOCK_M) : ceil_div(a->m, BLOCK_M);
  p.n_tiles = ceil_div(a->n, bn);
  p.ntile_shift = ((p.n_tiles & (p.n_tiles - 1)) == 0) ? __builtin_ctz((unsigned)p.n_tiles) : -1;
  p.rows_per_sample = a->rows_per_sample;
  p.bias = a->bias;
  p.sbias = a->sbias; p.ld_sbias = a->ld_sbias;
  p.gate = a->gate; p.ld_gate = a->ld_gate;
  p.act = a->act;
  p.group_width = gw; p.group_valid = gv; p.n_groups = gw > 0 ? a->n / gw : 1;
  p.group_shift = (gw > 0 && (gw & (gw - 1)) == 0) ? __builtin_ctz((unsigned)gw) : -1;
  p.peep_i = a->peep_i; p.peep_f = a->peep_f; p.ld_peep = a->ld_peep; p.peep16 = a->peep_f16 ? 1 : 0;
  p.a_row_ss = a->a_row_sumsq;
  p.cprev = a->cprev; p.ld_cprev = a->ld_cprev;
  p.out = a->out; p.ldo = a->ldo; p.out_fp32 = a->out_fp32;
  p.row_sumsq = a->row_sumsq; p.stats = a->stats;
  // output tensor map: [batch][rows][ldo], box = 32 rows x 64 bytes, 64B swizzle (what the epilogue warps stage)
  CUtensorMap tO;
  const int esz = a->out_fp32 ? 4 : 2;
  const uint64_t orows = batched ? (uint64_t)a->rows_per_sample : (uint64_t)a->m;
  rc = make_tmap_3d_sw(&tO, a->out_fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, esz, a->out, (uint64_t)a->ldo,
                       orows, batched ? (uint64_t)batch : 1, (uint64_t)a->ldo * esz, orows * a->ldo * esz, 64 / esz, 32,
                       CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  if (narrow) return launch_gemm<32, EPI_GENERIC, 1>(tA1, tA2, tW, tO, p, stream);
  if (clustered) return g_two_sm ? launch_gemm<256, EPI_GENERIC, 2, true>(tA1, tA2, tW, tO, p, stream)
                                 : launch_gemm<256, EPI_GENERIC, 2, false>(tA1, tA2, tW, tO, p, stream);
  return launch_gemm<256, EPI_GENERIC, 1>(tA1, tA2, tW, tO, p, stream);
}

extern "C" int cmpc_mutan_f16(const cmpc_mutan_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CMPC_REQUIRE(a != nullptr, CMPC_ERR_ARG, "cmpc_mutan_f16: null args");
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(a->a && a->w && a->out && a->bias && a->lang, CMPC_ERR_ARG, "cmpc_mutan_f16: null operand");
  CMPC_REQUIRE(a->m > 0 && a->c > 0 && a->c % 8 == 0 && a->k > 0 && a->lda % 8 == 0 && a->ldw % 8 == 0, CMPC_ERR_ARG,
               "cmpc_mutan_f16: bad shape m=%d c=%d k=%d", a->m, a->c, a->k);
  CMPC_REQUIRE(a->ldo % (a->out_f16 ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, CMPC_ERR_ALIGN,
               "cmpc_mutan_f16: out must be 16-byte aligned with ldo %% 4 == 0 (fp32) / %% 8 == 0 (fp16)");
  constexpr int BN = 240;
  const int kt = ceil_div(a->k, BLOCK_K);
  const int chunks = ceil_div(a->c, 48);
  CMPC_REQUIRE(a->ldw >= (int64_t)kt * BLOCK_K, CMPC_ERR_ARG, "cmpc_mutan_f16: ldw too small");
  CUtensorMap tA, tW;
  rc = make_tmap_2d(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->a, a->k, a->m, a->lda * 2, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  const bool clustered = ceil_div(a->m, BLOCK_M) >= 2;
  rc = make_tmap_2d(&tW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->w, (uint64_t)kt * BLOCK_K, (uint64_t)chunks * BN, a->ldw * 2, BLOCK_K,
                    clustered ? BN / 2 : BN);
  if (rc) return rc;
  GemmKernelParams p{};
  p.M = a->m; p.N = chunks * BN; p.kt1 = kt; p.kt2 = 0;
  p.m_tiles = ceil_div(a->m, BLOCK_M); p.n_tiles = chunks; p.ntile_shift = -1; p.group_shift = -1;
  p.rows_per_sample = a->rows_per_sample;
  p.a_row_ss = a->a_row_sumsq;
  p.C = a->c; p.mbias = a->bias; p.ld_mbias = a->ld_bias; p.lang = a->lang; p.ld_lang = a->ld_lang; p.lang_bstride = a->lang_batch_stride > 0 ? a->lang_batch_stride : 5 * a->ld_lang;
  p.out = a->out; p.ldo = a->ldo; p.out_fp32 = a->out_f16 ? 0 : 1; p.row_sumsq = a->row_sumsq;
  CUtensorMap tO;   // fp32 (or fp16) [1][M][ldo], box = 32 rows x 24 columns (what one epilogue warp produces), no swizzle
  if (a->out_f16)
    rc = make_tmap_3d_sw(&tO, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->out, (uint64_t)a->ldo, (uint64_t)a->m, 1, (uint64_t)a->ldo * 2,
                         (uint64_t)a->m * a->ldo * 2, 24, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
  else
    rc = make_tmap_3d_sw(&tO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->out, (uint64_t)a->ldo, (uint64_t)a->m, 1, (uint64_t)a->ldo * 4,
                         (uint64_t)a->m * a->ldo * 4, 24, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  if (clustered) return g_two_sm ? launch_gemm<BN, EPI_MUTAN, 2, true>(tA, tA, tW, tO, p, stream)
                                 : launch_gemm<BN, EPI_MUTAN, 2, false>(tA, tA, tW, tO, p, stream);
  return launch_gemm<BN, EPI_MUTAN, 1>(tA, tA, tW, tO, p, stream);
}

// Backward of mutan_head x5 + mutan_fusion up to the pre-activations (CMPC_model.py:295-323); see epi_mutan_bwd_compute.
extern "C" int cmpc_mutan_bwd_f16(const cmpc_mutan_args* a, const float* ds, int64_t ld_ds, float* dlang, int64_t dlang_batch_stride,
                                  float* dbias, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CMPC_REQUIRE(a != nullptr && ds != nullptr, CMPC_ERR_ARG, "cmpc_mutan_bwd_f16: null args");
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(a->a && a->w && a->out && a->bias && a->lang, CMPC_ERR_ARG, "cmpc_mutan_bwd_f16: null operand");
  CMPC_REQUIRE(a->m > 0 && a->c > 0 && a->c % 8 == 0 && a->k > 0 && a->lda % 8 == 0 && a->ldw % 8 == 0 && ld_ds % 4 == 0, CMPC_ERR_ARG,
               "cmpc_mutan_bwd_f16: bad shape m=%d c=%d k=%d", a->m, a->c, a->k);
  constexpr int BN = 240;
  const int kt = ceil_div(a->k, BLOCK_K);
  const int chunks = ceil_div(a->c, 48);
  CMPC_REQUIRE(a->ldo % 8 == 0 && a->ldo >= (int64_t)chunks * BN && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, CMPC_ERR_ALIGN,
               "cmpc_mutan_bwd_f16: out (fp16 [m, ldo]) needs ldo %% 8 == 0, ldo >= chunks * 240");
  CMPC_REQUIRE(a->ldw >= (int64_t)kt * BLOCK_K, CMPC_ERR_ARG, "cmpc_mutan_bwd_f16: ldw too small");
  CUtensorMap tA, tW;
  rc = make_tmap_2d(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->a, a->k, a->m, a->lda * 2, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  const bool clustered = ceil_div(a->m, BLOCK_M) >= 2;
  rc = make_tmap_2d(&tW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->w, (uint64_t)kt * BLOCK_K, (uint64_t)chunks * BN, a->ldw * 2, BLOCK_K,
                    clustered ? BN / 2 : BN);
  if (rc) return rc;
  GemmKernelParams p{};
  p.M = a->m; p.N = chunks * BN; p.kt1 = kt; p.kt2 = 0;
  p.m_tiles = ceil_div(a->m, BLOCK_M); p.n_tiles = chunks; p.ntile_shift = -1; p.group_shift = -1;
  p.rows_per_sample = a->rows_per_sample;
  p.a_row_ss = a->a_row_sumsq;
  p.C = a->c; p.mbias = a->bias; p.ld_mbias = a->ld_bias; p.lang = a->lang; p.ld_lang = a->ld_lang;
  p.lang_bstride = a->lang_batch_stride > 0 ? a->lang_batch_stride : 5 * a->ld_lang;
  p.out = a->out; p.ldo = a->ldo; p.out_fp32 = 0;
  p.ds = ds; p.ld_ds = ld_ds; p.dlang = dlang; p.dlang_bstride = dlang_batch_stride; p.dmbias = dbias;
  CUtensorMap tO;   // fp16 [1][M][ldo], box = 32 rows x 24 columns (one head of one epilogue warp), no swizzle
  rc = make_tmap_3d_sw(&tO, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->out, (uint64_t)a->ldo, (uint64_t)a->m, 1, (uint64_t)a->ldo * 2,
                       (uint64_t)a->m * a->ldo * 2, 24, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  if (clustered) return g_two_sm ? launch_gemm<BN, EPI_MUTAN_BWD, 2, true>(tA, tA, tW, tO, p, stream)
                                 : launch_gemm<BN, EPI_MUTAN_BWD, 2, false>(tA, tA, tW, tO, p, stream);
  return launch_gemm<BN, EPI_MUTAN_BWD, 1>(tA, tA, tW, tO, p, stream);
}

#include "common.cuh"

#include <stdarg.h>
#include <string.h>

namespace cmpc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return CMPC_ERR_LAUNCH;
  }
  return CMPC_OK;
}

static int g_cc_major[64];
static int g_sms[64];
static bool g_dev_init[64];

static int query_device(int* dev_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("cudaGetDevice failed: %s (no CUDA device; libcmpc_b200 has no CPU path)", cudaGetErrorString(e));
    return CMPC_ERR_ARCH;
  }
  if (!g_dev_init[dev]) {
    int major = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_cc_major[dev] = major;
    g_sms[dev] = sms;
    g_dev_init[dev] = true;
  }
  *dev_out = dev;
  return CMPC_OK;
}

int require_sm100() {
  int dev;
  int rc = query_device(&dev);
  if (rc) return rc;
  if (g_cc_major[dev] != 10) {
    set_error("device %d has compute capability %d.x; libcmpc_b200 is sm_100a only", dev, g_cc_major[dev]);
    return CMPC_ERR_ARCH;
  }
  return CMPC_OK;
}

bool first_use_on_device(unsigned long long* flags) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  const unsigned long long bit = 1ull << dev;
  if (__atomic_fetch_or(flags, bit, __ATOMIC_ACQ_REL) & bit) return false;
  return true;
}

static int g_pdl = 0;
bool pdl_enabled() { return g_pdl != 0; }
void set_pdl(int on) { g_pdl = on; }

int num_sms() {
  int dev;
  if (query_device(&dev)) return 148;
  return g_sms[dev] > 0 ? g_sms[dev] : 148;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_3d(CUtensorMap* map, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t batch, uint64_t row_stride_bytes, uint64_t batch_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_3d_sw(map, dt, elt_bytes, base, inner, outer, batch, row_stride_bytes, batch_stride_bytes, box_inner,
                         box_outer, CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_3d_sw(CUtensorMap* map, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t inner,
                    uint64_t outer, uint64_t batch, uint64_t row_stride_bytes, uint64_t batch_stride_bytes,
                    uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle) {
  PFN_encodeTiled enc = get_encode();
  CMPC_REQUIRE(enc != nullptr, CMPC_ERR_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
  CMPC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, CMPC_ERR_ALIGN, "TMA base %p not 16-byte aligned", base);
  CMPC_REQUIRE((row_stride_bytes & 15) == 0, CMPC_ERR_ALIGN, "TMA row stride %llu bytes not a multiple of 16",
               (unsigned long long)row_stride_bytes);
  CMPC_REQUIRE(box_inner * (uint32_t)elt_bytes <= 128 && box_outer <= 256, CMPC_ERR_ARG, "TMA box too large");
  const bool three = batch > 0;
  cuuint64_t dims[3] = {inner, outer, three ? batch : 1};
  cuuint64_t strides[2] = {row_stride_bytes, three ? batch_stride_bytes : row_stride_bytes * outer};
  cuuint32_t box[3] = {box_inner, box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, dt, three ? 3 : 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CMPC_REQUIRE(r == CUDA_SUCCESS, CMPC_ERR_LAUNCH,
               "cuTensorMapEncodeTiled failed (%d): inner %llu outer %llu stride %llu box %ux%u", (int)r,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
               box_outer);
  return CMPC_OK;
}

int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_3d(map, dt, elt_bytes, base, inner, outer, 0, row_stride_bytes, 0, box_inner, box_outer);
}

}  // namespace cmpc

extern "C" const char* cmpc_last_error(void) { return cmpc::g_err; }
extern "C" int cmpc_version(void) { return 100; }

extern "C" void cmpc_set_pdl(int32_t on) { cmpc::set_pdl(on); }

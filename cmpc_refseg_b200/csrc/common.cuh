// Host-side helpers shared by every translation unit of libcmpc_b200: status codes, last-error
// string, launch checking, TMA tensor-map encoding through the driver entry point (no -lcuda).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cmpc_b200.h"

namespace cmpc {

void set_error(const char* fmt, ...);          // records cmpc_last_error()
int check_launch(const char* what);            // cudaGetLastError -> status
int require_sm100();                           // CMPC_ERR_ARCH unless the current device is sm_100

// 2-D row-major tensor map, 128-byte swizzle, zero fill out of bounds.
//   inner = contiguous extent (elements), outer = rows, row_stride_bytes multiple of 16.
int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);
// 3-D variant: [batch][outer][inner]
int make_tmap_3d(CUtensorMap* map, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t batch, uint64_t row_stride_bytes, uint64_t batch_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer);
// same, explicit swizzle (CU_TENSOR_MAP_SWIZZLE_64B for 64-byte rows)
int make_tmap_3d_sw(CUtensorMap* map, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t inner,
                    uint64_t outer, uint64_t batch, uint64_t row_stride_bytes, uint64_t batch_stride_bytes,
                    uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle);

int num_sms();
// Programmatic dependent launch (cmpc_set_pdl): kernels that support it are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and execute griddepcontrol.wait before their first global access, so their
// launch + prologue (barrier init, TMEM allocation, descriptor prefetch) overlap the tail of the kernel in front of them.
bool pdl_enabled();
void set_pdl(int on);
// true exactly once per (call site's flag word, current device): function attributes such as the dynamic shared-memory
// limit are per device, so a process that drives several GPUs must set them on each (flags = one bit per device ordinal)
bool first_use_on_device(unsigned long long* flags);

#define CMPC_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      cmpc::set_error(__VA_ARGS__);    \
      return (code);                   \
    }                                  \
  } while (0)

}  // namespace cmpc

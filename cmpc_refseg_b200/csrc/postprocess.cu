// Post-processing after the head (trainval_model.py:243-245, 266): threshold the upsampled logits, resize
// and centre-crop the mask to each sample's ground-truth size (util/im_processing.py:25-41 -> skimage.transform.resize,
// order 1, no anti-aliasing), and count |pred & gt|, |pred | gt| (util/eval_tools.py:31-35).  One pass, ragged outputs.
#include "common.cuh"

namespace cmpc {

struct PostMeta { int gh, gw, res_h, res_w, crop_h, crop_w; };

// position of resized-image index i on the input axis: (i + 0.5) * n_in / res - 0.5 in double, the operation order of the oracle
__device__ __forceinline__ void axis_taps(int i, int n_in, int res, int reflect, int& lo, int& hi, bool& use_hi, bool& lo_ok, bool& hi_ok) {
  const double pos = ((double)i + 0.5) * ((double)n_in / (double)res) - 0.5;
  const double fl = floor(pos);
  lo = (int)fl;
  hi = (int)ceil(pos);
  use_hi = (pos - fl) > 0.0;         // weight of the ceil neighbour; the floor neighbour's weight 1 - frac is always > 0
  if (reflect) {                     // skimage 'reflect' (numpy 'symmetric'): -1 -> 0, n -> n - 1
    lo = lo < 0 ? -lo - 1 : (lo >= n_in ? 2 * n_in - 1 - lo : lo);
    hi = hi < 0 ? -hi - 1 : (hi >= n_in ? 2 * n_in - 1 - hi : hi);
    lo = min(max(lo, 0), n_in - 1);
    hi = min(max(hi, 0), n_in - 1);
    lo_ok = hi_ok = true;
  } else {                           // 'constant', cval = 0: outside pixels contribute nothing
    lo_ok = lo >= 0 && lo < n_in;
    hi_ok = hi >= 0 && hi < n_in;
    lo = min(max(lo, 0), n_in - 1);
    hi = min(max(hi, 0), n_in - 1);
  }
}

// A bilinear blend of {0,1} pixels with non-negative weights is non-zero iff a neighbour with positive weight is 1, and
// compute_mask_IU only asks whether the resized value is non-zero: the mask is exact without evaluating the blend.
__global__ void postprocess_iou_kernel(const float* __restrict__ up, int H, int W, float thresh, const unsigned char* __restrict__ gt,
                                       const long long* __restrict__ gt_offset, const PostMeta* __restrict__ meta, int reflect,
                                       unsigned char* __restrict__ pred_out, unsigned long long* __restrict__ iu /*[B,2]*/) {
  const int b = blockIdx.y;
  const PostMeta m = meta[b];
  const float* u = up + (long long)b * H * W;
  const long long base = gt_offset[b];
  const long long total = (long long)m.gh * m.gw;
  unsigned int ci = 0, cu = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / m.gw), x = (int)(i - (long long)y * m.gw);
    int rlo, rhi, clo, chi;
    bool ur, uc, rlo_ok, rhi_ok, clo_ok, chi_ok;
    axis_taps(y + m.crop_h, H, m.res_h, reflect, rlo, rhi, ur, rlo_ok, rhi_ok);
    axis_taps(x + m.crop_w, W, m.res_w, reflect, clo, chi, uc, clo_ok, chi_ok);
    bool p = rlo_ok && clo_ok && __ldg(u + (long long)rlo * W + clo) >= thresh;
    if (uc) p = p || (rlo_ok && chi_ok && __ldg(u + (long long)rlo * W + chi) >= thresh);
    if (ur) {
      p = p || (rhi_ok && clo_ok && __ldg(u + (long long)rhi * W + clo) >= thresh);
      if (uc) p = p || (rhi_ok && chi_ok && __ldg(u + (long long)rhi * W + chi) >= thresh);
    }
    const bool l = gt[base + i] != 0;
    if (pred_out) pred_out[base + i] = p ? 1 : 0;
    ci += (p && l) ? 1u : 0u;
    cu += (p || l) ? 1u : 0u;
  }
  ci = __reduce_add_sync(0xffffffffu, ci);
  cu = __reduce_add_sync(0xffffffffu, cu);
  __shared__ unsigned int s_i[32], s_u[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_i[warp] = ci; s_u[warp] = cu; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    unsigned int a = lane < nw ? s_i[lane] : 0u, c = lane < nw ? s_u[lane] : 0u;
    a = __reduce_add_sync(0xffffffffu, a);
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && (a | c)) {
      atomicAdd(iu + 2 * b, (unsigned long long)a);
      atomicAdd(iu + 2 * b + 1, (unsigned long long)c);
    }
  }
}

}  // namespace cmpc

using namespace cmpc;

extern "C" int cmpc_postprocess_iou(const float* up, int32_t batch, int32_t h, int32_t w, float score_thresh, const uint8_t* gt,
                                    const int64_t* gt_offset, const int32_t* meta, int32_t reflect, uint8_t* pred_out, uint64_t* iu,
                                    void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CMPC_REQUIRE(up && gt && gt_offset && meta && iu && batch > 0 && h > 0 && w > 0, CMPC_ERR_ARG, "cmpc_postprocess_iou: bad args");
  static_assert(sizeof(PostMeta) == 6 * sizeof(int32_t), "meta layout");
  postprocess_iou_kernel<<<dim3(64, batch), 256, 0, (cudaStream_t)stream>>>(up, h, w, score_thresh, gt, (const long long*)gt_offset,
                                                                            (const PostMeta*)meta, reflect, pred_out,
                                                                            (unsigned long long*)iu);
  return check_launch("postprocess_iou_kernel");
}

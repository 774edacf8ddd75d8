// gs_depthkey.cuh — the 16-bit depth key of the reference's counting sort (index.js:557-561), shared by the full
// depth sort (gs_sort.cu) and the front-to-back slab path (gs_slab.cu).
#pragma once
#include "gs_common.cuh"

namespace gs {

// ECMAScript ToInt32 (index.js:561 `| 0`)
__device__ __forceinline__ int32_t js_to_int32(double d) {
  if (!isfinite(d)) return 0;
  double t = trunc(d);
  if (t >= -2147483648.0 && t <= 2147483647.0) return (int32_t)t;
  double m = fmod(t, 4294967296.0);
  if (m < 0) m += 4294967296.0;
  return (int32_t)(uint32_t)m;
}

// index.js:561: sizeList[i] = ((depthList[i] - minDepth) * depthInv) | 0
__device__ __forceinline__ int32_t depth_key(float depth_f32, double min_depth, double depth_inv) {
  return js_to_int32(__dmul_rn(__dsub_rn((double)depth_f32, min_depth), depth_inv));
}

struct DepthRange {
  double min_depth, depth_inv;
};
__device__ __forceinline__ DepthRange load_depth_range(const FrameCounters *ctr) {
  // min is stored bit-inverted so that a zero-initialised word means "no value yet"
  const double mn = dec_f64(~ctr->sort.min_enc);
  const double mx = dec_f64(ctr->sort.max_enc);
  DepthRange r;
  r.min_depth = mn;
  r.depth_inv = __ddiv_rn(65535.0, __dsub_rn(mx, mn));  // index.js:558
  return r;
}

}  // namespace gs

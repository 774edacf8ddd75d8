// gs_common.cuh — shared declarations of the B200 splat path (context, device counters, launch API).
//
// The whole library is compiled with --fmad=false: no implicit FMA contraction, so fp32/fp64
// expressions execute in the order written (the reference's JS fp64 and GLSL fp32 semantics are
// restated op by op; see DESIGN.md "numeric model").  Fused multiply-adds are written explicitly
// (__fmaf_rn) where the parity definition calls for them or where they are rounding-neutral
// accumulations inside the stated tolerance.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "../../include/gsplat_b200.h"

namespace gs {

constexpr int kTile = 16;                 // 16x16 screen tiles (north_star)
constexpr int kRadixThreads = 256;
constexpr int kRadixItems = 16;
constexpr int kRadixTile = kRadixThreads * kRadixItems;  // 4096 elements per look-back tile
constexpr int kEmitThreads = 256;
constexpr int kEmitItems = 4;
constexpr int kEmitTile = kEmitThreads * kEmitItems;      // 1024 sorted splats per emission tile
constexpr uint32_t kInvalidDigit = 0xFFFFFFFFu;
constexpr uint32_t kNoRect = 0xFFFFFFFFu;
constexpr uint16_t kNoTile = 0xFFFFu;
// depth sentinel: a splat rejected by the worker filter (index.js:548 keeps only depth < 0)
#define GS_DEPTH_REJECT 1.0f

// look-back status word: 2 flag bits + 30 value bits
constexpr uint32_t kFlagAgg = 1u << 30;
constexpr uint32_t kFlagIncl = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValMask = ~kFlagMask;
constexpr unsigned long long kFlagAgg64 = 1ull << 62;
constexpr unsigned long long kFlagIncl64 = 2ull << 62;
constexpr unsigned long long kFlagMask64 = 3ull << 62;

// Device-resident per-frame counters: zeroed by one memset at the start of every sort/render.
struct FrameCounters {
  unsigned long long min_enc;  // bit-inverted order-preserving encoding of the fp64 min depth (atomicMax)
  unsigned long long max_enc;  // order-preserving encoding of the fp64 max depth (atomicMax)
  unsigned long long n_inst;   // D: emitted tile instances (bounding-rectangle candidates)
  uint32_t n_valid;            // V: splats passing the worker filter
  uint32_t n_inrange;          // V - dropped: entries with a key in [0,65535]
  uint32_t n_dropped;          // quirk Q5
  uint32_t n_visible;          // V2
  uint32_t n_inst_kept;        // instances surviving the exact footprint test and the tile-ownership filter
  uint32_t overflow;           // instance buffer too small: frame must be re-run
};

struct RenderConsts {
  float proj[16];
  float mv[16];
  float vw, vh, focal;
  uint32_t width, height;
  uint32_t tiles_x, tiles_y, n_tiles;
  float bg[4];
  uint32_t shard_rank, shard_world;
  int32_t out_format;
  uint32_t out_tiled;
};

struct SortConsts {
  double view[4];
  double cutout[16];
  int has_cutout;
};

}  // namespace gs

struct gs_context {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string err;

  // ---- resident splat table (HBM layout a8: 16 B + 16 B + 4 B per splat) ----
  uint32_t n = 0, cap = 0;
  float4 *center_scale = nullptr;
  uint4 *cov_color = nullptr;
  float *size_alpha = nullptr;

  // ---- per-splat scratch (sized to cap) ----
  uint32_t scratch_cap = 0;
  float *depth = nullptr;        // f32 depth or GS_DEPTH_REJECT
  uint32_t *idx_a = nullptr;     // after depth pass 1
  uint8_t *dig_a = nullptr;
  uint32_t *order = nullptr;     // draw order (== reference sortedIndexes)
  float4 *proj_rec = nullptr;    // 2 x float4 per splat
  uint32_t *rect = nullptr;      // packed tile rect per splat
  uint32_t *table_n = nullptr;   // radix chunk histograms of the depth passes [256][table_n_stride]
  uint32_t table_n_stride = 0;
  uint32_t *totals = nullptr;    // [512]: digit totals of the depth / tile passes
  uint32_t *tile_total = nullptr;  // instances per 1024-entry emission slice

  // ---- per-instance scratch (sized to cap_inst) ----
  uint64_t cap_inst = 0;
  uint16_t *inst_tile = nullptr;
  uint32_t *inst_idx = nullptr;
  uint8_t *inst_dig_b = nullptr;
  uint32_t *inst_idx_b = nullptr;
  float4 *inst_rec = nullptr;    // 2 x float4 per instance, sorted by (tile, draw order)
  uint32_t *table_d = nullptr;   // radix chunk histograms of the tile passes [256][table_d_stride]
  uint32_t table_d_stride = 0;

  // ---- per-frame tables ----
  uint32_t tiles_cap = 0;
  uint32_t *tile_count = nullptr;  // [T]
  uint32_t *tile_start = nullptr;  // [T+1]
  gs::FrameCounters *counters = nullptr;     // device
  gs::FrameCounters *counters_host = nullptr;  // pinned
  double *quirk_table = nullptr;   // parseInt quirk thresholds (device)
  int quirk_n = 0;

  // ---- frame buffer owned by the context (used when the caller passes host memory) ----
  void *frame_dev = nullptr;
  size_t frame_bytes = 0;
  void *frame_pinned = nullptr;
  size_t frame_pinned_bytes = 0;

  uint32_t shard_rank = 0, shard_world = 1;
  bool have_order = false;
  uint32_t order_count = 0;
  gs_stats stats{};
  cudaEvent_t ev[8]{};
};

namespace gs {

// -- launchers (each .cu file owns its kernels) --
// sort
void launch_depth_cull(gs_context *c, const SortConsts &sc);
void launch_depth_radix(gs_context *c);  // two passes -> c->order
// pack
void launch_pack(gs_context *c, const uint8_t *rows_dev, uint32_t first, uint32_t n);
// project + bin
void launch_project(gs_context *c, const RenderConsts &rc);
void launch_emit(gs_context *c, const RenderConsts &rc);
void launch_tile_radix(gs_context *c);
void launch_tile_scan(gs_context *c, const RenderConsts &rc);
// raster
void launch_raster(gs_context *c, const RenderConsts &rc, void *out_dev);
void launch_assemble(gs_context *c, const void *gathered, uint32_t tiles_per_rank, uint32_t world, uint32_t width,
                     uint32_t height, int32_t format, void *out_frame);

// order-preserving u64 encoding of an fp64 value (for atomicMin / atomicMax)
__host__ __device__ inline unsigned long long enc_f64(double d) {
#ifdef __CUDA_ARCH__
  unsigned long long u = (unsigned long long)__double_as_longlong(d);
#else
  unsigned long long u;
  memcpy(&u, &d, 8);
#endif
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double dec_f64(unsigned long long e) {
  unsigned long long u = (e & 0x8000000000000000ull) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d;
  memcpy(&d, &u, 8);
  return d;
#endif
}

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace gs

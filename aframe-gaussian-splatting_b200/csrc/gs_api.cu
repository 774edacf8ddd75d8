// gs_api.cu — the C ABI of include/gsplat_b200.h: context lifetime, the worker protocol
// (clear / push / sort, reference index.js:572-598) and the draw (index.js:184-207 + shaders).
// Host code only orchestrates: every per-splat / per-pixel operation runs in the CUDA kernels of
// gs_sort.cu, gs_pack.cu, gs_project.cu and gs_raster.cu.  There is no CPU fallback.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <vector>

#include "gs_common.cuh"

namespace gs {
uint32_t owned_tiles_host(uint32_t width, uint32_t height, uint32_t rank, uint32_t world);
}
using namespace gs;

static thread_local std::string g_create_error;

#define GS_CUDA(ctx, expr)                                                                             \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      char _b[512];                                                                                    \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      (ctx)->err = _b;                                                                                 \
      return (_e == cudaErrorMemoryAllocation) ? GS_ERR_OOM : GS_ERR_CUDA;                             \
    }                                                                                                  \
  } while (0)

static int fail(gs_context *c, int code, const char *msg) {
  if (c) c->err = msg;
  return code;
}

template <class T>
static cudaError_t dev_alloc(T **p, size_t count) {
  return cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T));
}
template <class T>
static void dev_free(T *&p) {
  if (p) cudaFree((void *)p);
  p = nullptr;
}

// ---------------------------------------------------------------------------------------------
// capacity management
// ---------------------------------------------------------------------------------------------
static int ensure_table(gs_context *c, uint64_t need) {
  if (need <= c->cap) return GS_OK;
  if (need > 0x7FFFFFFFull) return fail(c, GS_ERR_CAPACITY, "more than 2^31-1 splats");
  uint64_t ncap = std::max<uint64_t>(need, (uint64_t)c->cap * 2);
  ncap = std::min<uint64_t>(std::max<uint64_t>(ncap, 1024), 0x7FFFFFFFull);
  float4 *cs = nullptr;
  uint4 *cc = nullptr;
  float *sa = nullptr;
  GS_CUDA(c, dev_alloc(&cs, ncap));
  GS_CUDA(c, dev_alloc(&cc, ncap));
  GS_CUDA(c, dev_alloc(&sa, ncap));
  if (c->n) {
    GS_CUDA(c, cudaMemcpyAsync(cs, c->center_scale, sizeof(float4) * c->n, cudaMemcpyDeviceToDevice, c->stream));
    GS_CUDA(c, cudaMemcpyAsync(cc, c->cov_color, sizeof(uint4) * c->n, cudaMemcpyDeviceToDevice, c->stream));
    GS_CUDA(c, cudaMemcpyAsync(sa, c->size_alpha, sizeof(float) * c->n, cudaMemcpyDeviceToDevice, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  dev_free(c->center_scale);
  dev_free(c->cov_color);
  dev_free(c->size_alpha);
  c->center_scale = cs;
  c->cov_color = cc;
  c->size_alpha = sa;
  c->cap = (uint32_t)ncap;
  return GS_OK;
}

static int ensure_scratch(gs_context *c) {
  if (c->scratch_cap >= c->cap && c->depth) return GS_OK;
  dev_free(c->depth); dev_free(c->idx_a); dev_free(c->dig_a); dev_free(c->order);
  dev_free(c->proj_rec); dev_free(c->rect); dev_free(c->table_n); dev_free(c->tile_total);
  const size_t n = c->cap;
  GS_CUDA(c, dev_alloc(&c->depth, n));
  GS_CUDA(c, dev_alloc(&c->idx_a, n));
  GS_CUDA(c, dev_alloc(&c->dig_a, n));
  GS_CUDA(c, dev_alloc(&c->order, n));
  GS_CUDA(c, dev_alloc(&c->proj_rec, 2 * n));
  GS_CUDA(c, dev_alloc(&c->rect, n));
  c->table_n_stride = (uint32_t)((n + kRadixTile - 1) / kRadixTile + 1);
  GS_CUDA(c, dev_alloc(&c->table_n, (size_t)256 * c->table_n_stride));
  GS_CUDA(c, dev_alloc(&c->tile_total, (n + kEmitTile - 1) / kEmitTile + 1));
  c->scratch_cap = c->cap;
  c->have_order = false;
  return GS_OK;
}

static int ensure_instances(gs_context *c, uint64_t need) {
  if (need <= c->cap_inst && c->inst_rec) return GS_OK;
  if (need >= (1ull << 30)) return fail(c, GS_ERR_CAPACITY, "more than 2^30 tile instances in one frame");
  dev_free(c->inst_tile); dev_free(c->inst_idx); dev_free(c->inst_dig_b); dev_free(c->inst_idx_b); dev_free(c->inst_rec);
  dev_free(c->table_d);
  c->table_d_stride = (uint32_t)((need + kRadixTile - 1) / kRadixTile + 1);
  GS_CUDA(c, dev_alloc(&c->table_d, (size_t)256 * c->table_d_stride));
  GS_CUDA(c, dev_alloc(&c->inst_tile, need));
  GS_CUDA(c, dev_alloc(&c->inst_idx, need));
  GS_CUDA(c, dev_alloc(&c->inst_dig_b, need));
  GS_CUDA(c, dev_alloc(&c->inst_idx_b, need));
  GS_CUDA(c, dev_alloc(&c->inst_rec, 2 * need));
  c->cap_inst = need;
  return GS_OK;
}

static int ensure_tiles(gs_context *c, uint32_t n_tiles) {
  if (n_tiles <= c->tiles_cap && c->tile_count) return GS_OK;
  dev_free(c->tile_count); dev_free(c->tile_start);
  GS_CUDA(c, dev_alloc(&c->tile_count, (size_t)n_tiles + 1));
  GS_CUDA(c, dev_alloc(&c->tile_start, (size_t)n_tiles + 2));
  c->tiles_cap = n_tiles;
  return GS_OK;
}

static int ensure_frame(gs_context *c, size_t bytes) {
  if (bytes <= c->frame_bytes && c->frame_dev) return GS_OK;
  if (c->frame_dev) cudaFree(c->frame_dev);
  c->frame_dev = nullptr;
  GS_CUDA(c, cudaMalloc(&c->frame_dev, bytes));
  c->frame_bytes = bytes;
  return GS_OK;
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
extern "C" const char *gs_version(void) { return "gsplat_b200 0.1 (sm_100a; restates aframe-gaussian-splatting index.js @ b50238f)"; }

extern "C" const char *gs_last_error(const gs_context *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int gs_create(int device_ordinal, gs_context **out_ctx) {
  if (!out_ctx) return GS_ERR_INVALID;
  *out_ctx = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
    return GS_ERR_CUDA;
  }
  if (device_ordinal < 0 || device_ordinal >= count) {
    g_create_error = "device ordinal out of range";
    return GS_ERR_INVALID;
  }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device_ordinal)) != cudaSuccess) {
    g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
    return GS_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only";
    return GS_ERR_CUDA;
  }
  gs_context *c = new (std::nothrow) gs_context();
  if (!c) return GS_ERR_OOM;
  c->device = device_ordinal;
  c->sm_count = prop.multiProcessorCount;
  auto bail = [&](const char *what, cudaError_t err) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    gs_destroy(c);
    return GS_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  for (auto &ev : c->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaHostAlloc((void **)&c->counters_host, sizeof(FrameCounters), cudaHostAllocDefault)) != cudaSuccess)
    return bail("cudaHostAlloc", e);
  // parseInt quirk table (gs_pack.cu): strtod("<d>e-<k>") for k = 323..7, d = 1..9, ascending
  std::vector<double> tab;
  for (int k = 323; k >= 7; --k)
    for (int d = 1; d <= 9; ++d) {
      char buf[32];
      snprintf(buf, sizeof(buf), "%de-%d", d, k);
      tab.push_back(strtod(buf, nullptr));
    }
  c->quirk_n = (int)tab.size();
  if ((e = cudaMalloc((void **)&c->counters, sizeof(FrameCounters))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc((void **)&c->totals, 512 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc((void **)&c->quirk_table, tab.size() * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemcpy(c->quirk_table, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
    return bail("cudaMemcpy", e);
  *out_ctx = c;
  return GS_OK;
}

extern "C" int gs_destroy(gs_context *c) {
  if (!c) return GS_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  dev_free(c->center_scale); dev_free(c->cov_color); dev_free(c->size_alpha);
  dev_free(c->depth); dev_free(c->idx_a); dev_free(c->dig_a); dev_free(c->order); dev_free(c->proj_rec); dev_free(c->rect);
  dev_free(c->inst_tile); dev_free(c->inst_idx); dev_free(c->inst_dig_b); dev_free(c->inst_idx_b); dev_free(c->inst_rec);
  dev_free(c->tile_count); dev_free(c->tile_start); dev_free(c->quirk_table);
  dev_free(c->table_n); dev_free(c->table_d); dev_free(c->tile_total); dev_free(c->totals); dev_free(c->counters);
  if (c->frame_dev) cudaFree(c->frame_dev);
  if (c->counters_host) cudaFreeHost(c->counters_host);
  for (auto &ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return GS_OK;
}

// ---------------------------------------------------------------------------------------------
// seam 1: clear / push / sort
// ---------------------------------------------------------------------------------------------
extern "C" int gs_clear(gs_context *c) {
  if (!c) return GS_ERR_INVALID;
  c->n = 0;
  c->have_order = false;
  c->order_count = 0;
  return GS_OK;
}

extern "C" int gs_num_splats(const gs_context *c, uint32_t *out_n) {
  if (!c || !out_n) return GS_ERR_INVALID;
  *out_n = c->n;
  return GS_OK;
}

extern "C" int gs_push_splats(gs_context *c, const void *rows32, uint32_t n) {
  if (!c || (!rows32 && n)) return GS_ERR_INVALID;
  if (!n) return GS_OK;
  GS_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_table(c, (uint64_t)c->n + n);
  if (rc) return rc;
  uint8_t *rows_dev = nullptr;
  GS_CUDA(c, cudaMalloc((void **)&rows_dev, (size_t)n * 32));
  cudaError_t e = cudaMemcpyAsync(rows_dev, rows32, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    launch_pack(c, rows_dev, c->n, n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(rows_dev);
  GS_CUDA(c, e);
  c->n += n;
  c->have_order = false;
  return GS_OK;
}

extern "C" int gs_push_packed(gs_context *c, const float *center_scale4, const uint32_t *cov_color4,
                              const float *size_alpha, uint32_t n) {
  if (!c || ((!center_scale4 || !cov_color4 || !size_alpha) && n)) return GS_ERR_INVALID;
  if (!n) return GS_OK;
  GS_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_table(c, (uint64_t)c->n + n);
  if (rc) return rc;
  GS_CUDA(c, cudaMemcpyAsync(c->center_scale + c->n, center_scale4, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  GS_CUDA(c, cudaMemcpyAsync(c->cov_color + c->n, cov_color4, sizeof(uint4) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  GS_CUDA(c, cudaMemcpyAsync(c->size_alpha + c->n, size_alpha, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  GS_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n += n;
  c->have_order = false;
  return GS_OK;
}

extern "C" int gs_read_packed(gs_context *c, uint32_t first, uint32_t n, float *center_scale4, uint32_t *cov_color4,
                              float *size_alpha) {
  if (!c || (uint64_t)first + n > c->n) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  if (center_scale4) GS_CUDA(c, cudaMemcpy(center_scale4, c->center_scale + first, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost));
  if (cov_color4) GS_CUDA(c, cudaMemcpy(cov_color4, c->cov_color + first, sizeof(uint4) * (size_t)n, cudaMemcpyDeviceToHost));
  if (size_alpha) GS_CUDA(c, cudaMemcpy(size_alpha, c->size_alpha + first, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
  return GS_OK;
}

static void fill_sort_consts(SortConsts &sc, const float view[4], const float *cutout) {
  memset(&sc, 0, sizeof(sc));
  for (int i = 0; i < 4; ++i) sc.view[i] = (double)view[i];
  sc.has_cutout = cutout ? 1 : 0;
  if (cutout)
    for (int i = 0; i < 16; ++i) sc.cutout[i] = (double)cutout[i];
}

// enqueue the sort kernels (counters must have been zeroed)
static uint32_t enqueue_sort(gs_context *c, const SortConsts &sc) {
  launch_depth_cull(c, sc);
  launch_depth_radix(c);
  return 7;
}

static void stats_from_counters(gs_context *c) {
  const FrameCounters &h = *c->counters_host;
  gs_stats &s = c->stats;
  s.n_splats = c->n;
  s.n_sorted = h.n_valid;
  s.n_dropped = h.n_dropped;
  s.n_visible = h.n_visible;
  s.n_instances = h.n_inst;
  s.n_instances_kept = h.n_inst_kept;
  s.min_depth = h.n_valid ? dec_f64(~h.min_enc) : INFINITY;
  s.max_depth = h.n_valid ? dec_f64(h.max_enc) : -INFINITY;
}

extern "C" int gs_sort(gs_context *c, const float view[4], const float *cutout16_or_null, uint32_t *out_idx,
                       uint32_t *out_count) {
  if (!c || !view) return GS_ERR_INVALID;
  if (c->n == 0) return fail(c, GS_ERR_EMPTY, "gs_sort before any push");
  GS_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_scratch(c);
  if (rc) return rc;
  SortConsts sc;
  fill_sort_consts(sc, view, cutout16_or_null);
  GS_CUDA(c, cudaMemsetAsync(c->counters, 0, sizeof(FrameCounters), c->stream));
  GS_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  const uint32_t launches = enqueue_sort(c, sc);
  GS_CUDA(c, cudaGetLastError());
  GS_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  GS_CUDA(c, cudaMemcpyAsync(c->counters_host, c->counters, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->stream));
  GS_CUDA(c, cudaStreamSynchronize(c->stream));
  memset(&c->stats, 0, sizeof(c->stats));
  stats_from_counters(c);
  c->stats.kernel_launches = launches;
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->stats.ms_sort = ms;
  c->stats.ms_total = ms;
  c->have_order = true;
  c->order_count = c->counters_host->n_valid;
  if (out_count) *out_count = c->order_count;
  if (out_idx && c->order_count)
    GS_CUDA(c, cudaMemcpy(out_idx, c->order, sizeof(uint32_t) * (size_t)c->order_count, cudaMemcpyDeviceToHost));
  return GS_OK;
}

// ---------------------------------------------------------------------------------------------
// seam 2: the draw
// ---------------------------------------------------------------------------------------------
extern "C" int gs_set_shard(gs_context *c, uint32_t rank, uint32_t world) {
  if (!c || world == 0 || rank >= world) return GS_ERR_INVALID;
  c->shard_rank = rank;
  c->shard_world = world;
  return GS_OK;
}

extern "C" uint32_t gs_owned_tiles(uint32_t width, uint32_t height, uint32_t rank, uint32_t world) {
  if (world == 0 || rank >= world) return 0;
  return owned_tiles_host(width, height, rank, world);
}

extern "C" int gs_render(gs_context *c, const gs_render_params *p, void *out_rgba, gs_stats *stats) {
  if (!c || !p || !out_rgba) return GS_ERR_INVALID;
  if (c->n == 0) return fail(c, GS_ERR_EMPTY, "gs_render before any push");
  if (p->width == 0 || p->height == 0 || p->width > 4096 || p->height > 4096)
    return fail(c, GS_ERR_INVALID, "frame size must be within 1..4096 per side");
  if (p->out_format != GS_FORMAT_RGBA8 && p->out_format != GS_FORMAT_RGBA32F) return fail(c, GS_ERR_INVALID, "bad out_format");
  GS_CUDA(c, cudaSetDevice(c->device));

  RenderConsts rc;
  memset(&rc, 0, sizeof(rc));
  memcpy(rc.proj, p->proj, sizeof(rc.proj));
  memcpy(rc.mv, p->modelview, sizeof(rc.mv));
  rc.width = p->width;
  rc.height = p->height;
  rc.vw = (float)p->width;
  rc.vh = (float)p->height;
  // index.js:191: focal = (viewport.w / 2.0) * Math.abs(projectionMatrix.elements[5]), fp64 then f32 uniform
  rc.focal = p->focal > 0.0f ? p->focal : (float)(((double)p->height / 2.0) * fabs((double)p->proj[5]));
  rc.tiles_x = (p->width + kTile - 1) / kTile;
  rc.tiles_y = (p->height + kTile - 1) / kTile;
  rc.n_tiles = rc.tiles_x * rc.tiles_y;
  if (rc.n_tiles >= 0xFFFFu) return fail(c, GS_ERR_INVALID, "more than 65534 tiles");
  memcpy(rc.bg, p->bg_rgba, sizeof(rc.bg));
  rc.shard_rank = c->shard_rank;
  rc.shard_world = c->shard_world;
  rc.out_format = p->out_format;
  rc.out_tiled = (p->flags & GS_RENDER_OUT_TILED) ? 1u : 0u;

  const bool reuse = (p->flags & GS_RENDER_REUSE_SORT) && c->have_order;
  int rcode = ensure_scratch(c);
  if (rcode) return rcode;
  if ((rcode = ensure_tiles(c, rc.n_tiles))) return rcode;
  if (c->cap_inst == 0) {
    if ((rcode = ensure_instances(c, std::max<uint64_t>(1u << 20, (uint64_t)c->n * 4)))) return rcode;
  }
  const size_t px_bytes = p->out_format == GS_FORMAT_RGBA8 ? 4 : 16;
  size_t out_pixels = (size_t)p->width * p->height;
  if (rc.out_tiled) out_pixels = (size_t)gs_owned_tiles(p->width, p->height, c->shard_rank, c->shard_world) * 256;
  const size_t out_bytes = out_pixels * px_bytes;
  void *out_dev = out_rgba;
  if (!(p->flags & GS_RENDER_OUT_DEVICE)) {
    if ((rcode = ensure_frame(c, out_bytes))) return rcode;
    out_dev = c->frame_dev;
  }

  SortConsts sc;
  const float view[4] = {p->modelview[2], p->modelview[6], p->modelview[10], p->modelview[14]};  // index.js:442
  fill_sort_consts(sc, view, p->has_cutout ? p->cutout16 : nullptr);

  // header of FrameCounters preserved across frames when the previous order is reused
  struct SortHeader { unsigned long long min_enc, max_enc; uint32_t n_valid, n_inrange, n_dropped; };
  SortHeader keep{};
  if (reuse) {
    keep.min_enc = c->counters_host->min_enc; keep.max_enc = c->counters_host->max_enc;
    keep.n_valid = c->counters_host->n_valid; keep.n_inrange = c->counters_host->n_inrange;
    keep.n_dropped = c->counters_host->n_dropped;
  }

  for (int attempt = 0; attempt < 8; ++attempt) {
    uint32_t launches = 0;
    GS_CUDA(c, cudaMemsetAsync(c->counters, 0, sizeof(FrameCounters), c->stream));
    GS_CUDA(c, cudaMemsetAsync(c->tile_count, 0, sizeof(uint32_t) * ((size_t)rc.n_tiles + 1), c->stream));
    GS_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    if (reuse) {
      GS_CUDA(c, cudaMemcpyAsync(&c->counters->min_enc, &keep.min_enc, 16, cudaMemcpyHostToDevice, c->stream));
      GS_CUDA(c, cudaMemcpyAsync(&c->counters->n_valid, &keep.n_valid, 12, cudaMemcpyHostToDevice, c->stream));
    } else {
      launches += enqueue_sort(c, sc);
    }
    GS_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    launch_project(c, rc);
    launches += 1;
    GS_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    launch_emit(c, rc);
    launch_tile_radix(c);
    launch_tile_scan(c, rc);
    launches += 9;
    GS_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    launch_raster(c, rc, out_dev);
    launches += 1;
    GS_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    GS_CUDA(c, cudaGetLastError());
    GS_CUDA(c, cudaMemcpyAsync(c->counters_host, c->counters, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->stream));
    if (!(p->flags & GS_RENDER_OUT_DEVICE))
      GS_CUDA(c, cudaMemcpyAsync(out_rgba, out_dev, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stats.kernel_launches = launches;
    if (!c->counters_host->overflow) break;
    // instance buffer too small: grow to the measured demand and run the frame again
    const uint64_t need = std::max<uint64_t>(c->counters_host->n_inst + c->counters_host->n_inst / 8, c->cap_inst * 2);
    if ((rcode = ensure_instances(c, need))) return rcode;
    if (attempt == 7) return fail(c, GS_ERR_CAPACITY, "instance buffer kept overflowing");
  }

  const uint32_t launches = c->stats.kernel_launches;
  memset(&c->stats, 0, sizeof(c->stats));
  stats_from_counters(c);
  c->stats.kernel_launches = launches;
  c->stats.n_tiles = rc.n_tiles;
  c->stats.width = p->width;
  c->stats.height = p->height;
  cudaEventElapsedTime(&c->stats.ms_sort, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&c->stats.ms_project, c->ev[1], c->ev[2]);
  cudaEventElapsedTime(&c->stats.ms_bin, c->ev[2], c->ev[3]);
  cudaEventElapsedTime(&c->stats.ms_raster, c->ev[3], c->ev[4]);
  cudaEventElapsedTime(&c->stats.ms_total, c->ev[0], c->ev[4]);
  c->have_order = true;
  c->order_count = c->counters_host->n_valid;
  if (stats) *stats = c->stats;
  return GS_OK;
}

extern "C" int gs_get_stats(const gs_context *c, gs_stats *out) {
  if (!c || !out) return GS_ERR_INVALID;
  *out = c->stats;
  return GS_OK;
}

extern "C" int gs_read_projected(gs_context *c, uint32_t first, uint32_t n, float *out8) {
  if (!c || !out8 || (uint64_t)first + n > c->n || !c->proj_rec) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  std::vector<float> rec((size_t)n * 8);
  std::vector<uint32_t> rect(n);
  GS_CUDA(c, cudaMemcpy(rec.data(), c->proj_rec + 2 * (size_t)first, sizeof(float) * 8 * (size_t)n, cudaMemcpyDeviceToHost));
  GS_CUDA(c, cudaMemcpy(rect.data(), c->rect + first, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < n; ++i) {
    memcpy(out8 + 8 * (size_t)i, rec.data() + 8 * (size_t)i, 32);
    memcpy(out8 + 8 * (size_t)i + 7, &rect[i], 4);
    if (rect[i] == kNoRect) {  // record slot holds stale data when the splat was not projected
      for (int k = 0; k < 7; ++k) out8[8 * (size_t)i + k] = 0.0f;
    }
  }
  return GS_OK;
}

extern "C" int gs_assemble_tiles(gs_context *c, const void *gathered, uint32_t tiles_per_rank, uint32_t world,
                                 uint32_t width, uint32_t height, int32_t format, void *out_frame) {
  if (!c || !gathered || !out_frame || world == 0) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  launch_assemble(c, gathered, tiles_per_rank, world, width, height, format, out_frame);
  GS_CUDA(c, cudaGetLastError());
  GS_CUDA(c, cudaStreamSynchronize(c->stream));
  return GS_OK;
}

// ---------------------------------------------------------------------------------------------
// device-memory helpers
// ---------------------------------------------------------------------------------------------
extern "C" int gs_device_alloc(gs_context *c, size_t bytes, void **out) {
  if (!c || !out) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  GS_CUDA(c, cudaMalloc(out, std::max<size_t>(bytes, 1)));
  return GS_OK;
}
extern "C" int gs_device_free(gs_context *c, void *p) {
  if (!c) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  if (p) GS_CUDA(c, cudaFree(p));
  return GS_OK;
}
extern "C" int gs_host_alloc(gs_context *c, size_t bytes, void **out) {
  if (!c || !out) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  GS_CUDA(c, cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocDefault));
  return GS_OK;
}
extern "C" int gs_host_free(gs_context *c, void *p) {
  if (!c) return GS_ERR_INVALID;
  if (p) GS_CUDA(c, cudaFreeHost(p));
  return GS_OK;
}
extern "C" int gs_memcpy_d2h(gs_context *c, void *dst, const void *src, size_t bytes) {
  if (!c || !dst || !src) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  GS_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  GS_CUDA(c, cudaStreamSynchronize(c->stream));
  return GS_OK;
}
extern "C" void *gs_stream(gs_context *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int gs_synchronize(gs_context *c) {
  if (!c) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  GS_CUDA(c, cudaStreamSynchronize(c->stream));
  return GS_OK;
}

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

static int drain(gs_context *c);

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
// Grow the resident table to hold `need` splats.  Growth is geometric (or exact, through gs_reserve) and is the one
// moment a push has to wait for the frames in flight: they read the buffers that are about to be replaced.
static int ensure_table(gs_context *c, uint64_t need, bool exact = false) {
  if (need <= c->cap) return GS_OK;
  if (need > 0x7FFFFFFFull) return fail(c, GS_ERR_CAPACITY, "more than 2^31-1 splats");
  int rc0 = drain(c);
  if (rc0) return rc0;
  uint64_t ncap = exact ? need : std::max<uint64_t>(need, (uint64_t)c->cap * 2);
  ncap = std::min<uint64_t>(std::max<uint64_t>(ncap, 1024), 0x7FFFFFFFull);
  float4 *cs = nullptr;
  uint4 *cc = nullptr;
  float *sa = nullptr;
  GS_CUDA(c, dev_alloc(&cs, ncap));
  GS_CUDA(c, dev_alloc(&cc, ncap));
  GS_CUDA(c, dev_alloc(&sa, ncap));
  if (c->n) {
    GS_CUDA(c, cudaMemcpyAsync(cs, c->center_scale, sizeof(float4) * c->n, cudaMemcpyDeviceToDevice, c->push_stream));
    GS_CUDA(c, cudaMemcpyAsync(cc, c->cov_color, sizeof(uint4) * c->n, cudaMemcpyDeviceToDevice, c->push_stream));
    GS_CUDA(c, cudaMemcpyAsync(sa, c->size_alpha, sizeof(float) * c->n, cudaMemcpyDeviceToDevice, c->push_stream));
  }
  GS_CUDA(c, cudaStreamSynchronize(c->push_stream));
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
  dev_free(c->depth); dev_free(c->idx_a); dev_free(c->dig_a); dev_free(c->table_n); dev_free(c->slice_total);
  for (int i = 0; i < 2; ++i) { dev_free(c->order[i]); dev_free(c->proj_rec[i]); dev_free(c->rect[i]); }
  dev_free(c->slice_prefix); dev_free(c->ent); dev_free(c->ent_off);
  const size_t n = c->cap;
  GS_CUDA(c, dev_alloc(&c->depth, n));
  GS_CUDA(c, dev_alloc(&c->idx_a, n));
  GS_CUDA(c, dev_alloc(&c->dig_a, n));
  for (int i = 0; i < 2; ++i) {
    GS_CUDA(c, dev_alloc(&c->order[i], n));
    GS_CUDA(c, dev_alloc(&c->proj_rec[i], 2 * n));
    GS_CUDA(c, dev_alloc(&c->rect[i], n));
  }
  c->table_n_stride = (uint32_t)((n + kRadixTile - 1) / kRadixTile + 1);
  GS_CUDA(c, dev_alloc(&c->table_n, (size_t)256 * c->table_n_stride));
  GS_CUDA(c, dev_alloc(&c->slice_total, (n + kEmitTile - 1) / kEmitTile + 1));
  GS_CUDA(c, dev_alloc(&c->slice_prefix, (n + kEmitTile - 1) / kEmitTile + 2));
  GS_CUDA(c, dev_alloc(&c->ent, n));
  GS_CUDA(c, dev_alloc(&c->ent_off, n));
  c->scratch_cap = c->cap;
  c->have_order = false;
  return GS_OK;
}

static int ensure_instances(gs_context *c, uint64_t need) {
  if (need <= c->cap_inst && c->inst_rec[0]) return GS_OK;
  if (need >= (1ull << 30)) return fail(c, GS_ERR_CAPACITY, "more than 2^30 tile instances in one frame");
  dev_free(c->inst_tile); dev_free(c->inst_idx); dev_free(c->inst_tile_b); dev_free(c->inst_tile_f); dev_free(c->inst_idx_b);
  dev_free(c->inst_rec[0]); dev_free(c->inst_rec[1]);
  dev_free(c->table_d);
  c->table_d_stride = (uint32_t)((need + kRadixTile / 2 - 1) / (kRadixTile / 2) + 2);  // one column per 2048-instance window
  GS_CUDA(c, dev_alloc(&c->table_d, (size_t)256 * c->table_d_stride));
  GS_CUDA(c, dev_alloc(&c->inst_tile, need));
  GS_CUDA(c, dev_alloc(&c->inst_idx, need));
  GS_CUDA(c, dev_alloc(&c->inst_tile_b, need));
  GS_CUDA(c, dev_alloc(&c->inst_tile_f, need));
  GS_CUDA(c, dev_alloc(&c->inst_idx_b, need));
  GS_CUDA(c, dev_alloc(&c->inst_rec[0], 2 * need));
  GS_CUDA(c, dev_alloc(&c->inst_rec[1], 2 * need));
  c->cap_inst = need;
  return GS_OK;
}

static int ensure_bins(gs_context *c, uint32_t n_bins) {
  if (n_bins <= c->bins_cap && c->bin_range[0]) return GS_OK;
  dev_free(c->bin_range[0]); dev_free(c->bin_range[1]);
  GS_CUDA(c, dev_alloc(&c->bin_range[0], (size_t)n_bins + 1));
  GS_CUDA(c, dev_alloc(&c->bin_range[1], (size_t)n_bins + 1));
  c->bins_cap = n_bins;
  return GS_OK;
}

static int ensure_tile_stats(gs_context *c, uint32_t n_tiles) {
  if (n_tiles <= c->tile_stats_cap && c->tile_stats) return GS_OK;
  dev_free(c->tile_stats);
  if (c->tile_stats_host) cudaFreeHost(c->tile_stats_host);
  c->tile_stats_host = nullptr;
  GS_CUDA(c, dev_alloc(&c->tile_stats, (size_t)n_tiles));
  GS_CUDA(c, cudaHostAlloc((void **)&c->tile_stats_host, sizeof(uint4) * (size_t)n_tiles, cudaHostAllocDefault));
  c->tile_stats_cap = n_tiles;
  return GS_OK;
}

// buffers of the front-to-back slab path (gs_slab.cu); the pipeline is idle when this runs
static int ensure_slab(gs_context *c, uint32_t n_tiles, uint32_t n_bins) {
  if (c->slab_cap < c->cap || !c->key32[0]) {
    dev_free(c->key32[0]); dev_free(c->key32[1]); dev_free(c->cidx); dev_free(c->ckey); dev_free(c->chunk_cnt[0]); dev_free(c->chunk_cnt[1]);
    GS_CUDA(c, dev_alloc(&c->key32[0], (size_t)c->cap + 8));
    GS_CUDA(c, dev_alloc(&c->key32[1], (size_t)c->cap + 8));
    GS_CUDA(c, dev_alloc(&c->cidx, (size_t)c->cap));
    GS_CUDA(c, dev_alloc(&c->ckey, (size_t)c->cap));
    c->chunk_row = (uint32_t)(c->cap / 2048 + 4);
    GS_CUDA(c, dev_alloc(&c->chunk_cnt[0], (size_t)c->chunk_row * kMaxSlabs));
    GS_CUDA(c, dev_alloc(&c->chunk_cnt[1], (size_t)c->chunk_row * kMaxSlabs));
    c->slab_cap = c->cap;
  }
  for (int i = 0; i < 2; ++i)
    if (!c->slab_tab[i]) GS_CUDA(c, dev_alloc(&c->slab_tab[i], 1));
  if (c->slab_tiles_cap < n_tiles || !c->pix_state) {
    dev_free(c->pix_state); dev_free(c->tile_closed); dev_free(c->bin_open);
    GS_CUDA(c, dev_alloc(&c->pix_state, (size_t)n_tiles * 256));
    GS_CUDA(c, dev_alloc(&c->tile_closed, (size_t)n_tiles));
    GS_CUDA(c, dev_alloc(&c->bin_open, (size_t)n_bins));  // bins <= tiles
    c->slab_tiles_cap = n_tiles;
  }
  return GS_OK;
}

static int ensure_frame(gs_context *c, gs_context::Slot &sl, size_t bytes) {
  if (bytes <= sl.frame_bytes && sl.frame_dev) return GS_OK;
  if (sl.frame_dev) cudaFree(sl.frame_dev);
  sl.frame_dev = nullptr;
  GS_CUDA(c, cudaMalloc(&sl.frame_dev, bytes));
  sl.frame_bytes = bytes;
  return GS_OK;
}

static void drop_graphs(gs_context *c) {
  auto kill = [](cudaGraphExec_t &g) { if (g) { cudaGraphExecDestroy(g); g = nullptr; } };
  for (auto &sl : c->slot)
    for (int i = 0; i < 2; ++i) {
      kill(sl.graph_a[i][0]); kill(sl.graph_a[i][1]); kill(sl.graph_b[i]); kill(sl.graph_r[i]); kill(sl.graph_rp[i]);
      kill(sl.graph_sa[i]); kill(sl.graph_sl[i][0]); kill(sl.graph_sl[i][1]); kill(sl.graph_sl[i][2]);
    }
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
extern "C" uint32_t gs_bin_size(void) { return (uint32_t)kBin; }

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
  // sort + binning kernels are short and latency-bound, the raster is one long issue-bound kernel: giving the
  // main / aux streams priority lets their CTAs slot in as raster CTAs retire, so frame k+1 is binned UNDER frame k's raster
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  if (const char *e = getenv("GS_PRIO")) {  // experiment knob: flat = every stage at one priority, inverse = raster first
    if (strcmp(e, "flat") == 0) prio_greatest = prio_least;
    if (strcmp(e, "inverse") == 0) std::swap(prio_least, prio_greatest);
  }
  if ((e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithPriority(&c->bstream, cudaStreamNonBlocking, prio_greatest)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithPriority(&c->rstream, cudaStreamNonBlocking, prio_least)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&c->push_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  for (int i = 0; i < 2; ++i)
    if ((e = cudaEventCreateWithFlags(&c->push_ev[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&c->push_done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  for (int i = 0; i < gs_context::kSlots; ++i) c->slot[i].index = i;
  if ((e = cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, prio_greatest)) != cudaSuccess) return bail("cudaStreamCreate", e);
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaEventCreateWithFlags(&c->ev_fork[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  }
  for (auto &ev : c->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
  for (auto &sl : c->slot) {
    for (auto &ev : sl.ev)
      if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (auto &ev : sl.evp)
      if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&sl.ev_binned, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&sl.ev_sorted, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&sl.ev_r0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&sl.ev_copied, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc((void **)&sl.ctr, sizeof(FrameCounters))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&sl.fp, sizeof(FrameParams))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaHostAlloc((void **)&sl.ctr_host, sizeof(FrameCounters), cudaHostAllocDefault)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if ((e = cudaHostAlloc((void **)&sl.fp_host, sizeof(FrameParams), cudaHostAllocDefault)) != cudaSuccess) return bail("cudaHostAlloc", e);
  }
  if ((e = cudaMalloc((void **)&c->sort_hdr, sizeof(SortHeader))) != cudaSuccess) return bail("cudaMalloc", e);
  c->use_graphs = getenv("GS_NO_GRAPH") == nullptr;
  // frames expected to sort at least GS_SLAB_MIN splats (default 16 M) are rendered front to back in depth slabs (gs_slab.cu);
  // GS_SLAB_FIRST = target entry count of the nearest slab (default 1 M, the following ones double)
  if (const char *e = getenv("GS_PDL")) c->use_pdl = strcmp(e, "1") == 0;
  if (const char *e = getenv("GS_EMIT")) c->emit_by_entry = strcmp(e, "windows") != 0;
  if (const char *e = getenv("GS_SLAB_MIN")) c->slab_min = (uint32_t)strtoull(e, nullptr, 10);
  if (const char *e = getenv("GS_SLAB_FIRST")) c->slab_first = std::max<uint32_t>(1024u, (uint32_t)strtoull(e, nullptr, 10));
  {  // pixel loop of the raster: packed fp32x2 (default) or scalar (GS_RASTER=scalar); both give identical frames
    const char *rk = getenv("GS_RASTER");
    c->raster_base_flags = (rk && strcmp(rk, "scalar") == 0) ? 0u : 1u;
  }
  // parseInt quirk table (gs_pack.cu): strtod("<d>e-<k>") for k = 323..7, d = 1..9, ascending
  std::vector<double> tab;
  for (int k = 323; k >= 7; --k)
    for (int d = 1; d <= 9; ++d) {
      char buf[32];
      snprintf(buf, sizeof(buf), "%de-%d", d, k);
      tab.push_back(strtod(buf, nullptr));
    }
  c->quirk_n = (int)tab.size();
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
  if (c->bstream) cudaStreamSynchronize(c->bstream);
  if (c->rstream) cudaStreamSynchronize(c->rstream);
  dev_free(c->center_scale); dev_free(c->cov_color); dev_free(c->size_alpha);
  dev_free(c->depth); dev_free(c->idx_a); dev_free(c->dig_a);
  for (int i = 0; i < 2; ++i) { dev_free(c->order[i]); dev_free(c->proj_rec[i]); dev_free(c->rect[i]); }
  dev_free(c->inst_tile); dev_free(c->inst_idx); dev_free(c->inst_tile_b); dev_free(c->inst_tile_f); dev_free(c->inst_idx_b);
  dev_free(c->inst_rec[0]); dev_free(c->inst_rec[1]);
  dev_free(c->bin_range[0]); dev_free(c->bin_range[1]); dev_free(c->quirk_table); dev_free(c->tile_stats);
  if (c->tile_stats_host) cudaFreeHost(c->tile_stats_host);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->push_stream) cudaStreamSynchronize(c->push_stream);
  for (int i = 0; i < 2; ++i) {
    if (c->push_pinned[i]) cudaFreeHost(c->push_pinned[i]);
    dev_free(c->push_dev[i]);
    if (c->push_ev[i]) cudaEventDestroy(c->push_ev[i]);
  }
  if (c->push_done) cudaEventDestroy(c->push_done);
  if (c->push_stream) cudaStreamDestroy(c->push_stream);
  drop_graphs(c);
  for (uint32_t r = 0; r < c->peer_world; ++r)
    if (r != c->peer_rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  if (c->peer_local) cudaFree(c->peer_local);
  dev_free(c->table_n); dev_free(c->table_d); dev_free(c->slice_total); dev_free(c->totals); dev_free(c->sort_hdr);
  dev_free(c->slice_prefix); dev_free(c->ent); dev_free(c->ent_off);
  dev_free(c->key32[0]); dev_free(c->key32[1]); dev_free(c->cidx); dev_free(c->ckey); dev_free(c->chunk_cnt[0]); dev_free(c->chunk_cnt[1]);
  dev_free(c->slab_tab[0]); dev_free(c->slab_tab[1]);
  dev_free(c->pix_state); dev_free(c->tile_closed); dev_free(c->bin_open);
  for (auto &sl : c->slot) {
    dev_free(sl.ctr); dev_free(sl.fp);
    if (sl.frame_dev) cudaFree(sl.frame_dev);
    if (sl.depth_dev) cudaFree(sl.depth_dev);
    if (sl.ctr_host) cudaFreeHost(sl.ctr_host);
    if (sl.fp_host) cudaFreeHost(sl.fp_host);
    for (auto &ev : sl.ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : sl.evp) if (ev) cudaEventDestroy(ev);
    if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    if (sl.ev_binned) cudaEventDestroy(sl.ev_binned);
    if (sl.ev_sorted) cudaEventDestroy(sl.ev_sorted);
    if (sl.ev_r0) cudaEventDestroy(sl.ev_r0);
    if (sl.ev_copied) cudaEventDestroy(sl.ev_copied);
    for (auto &pr : sl.slab_ev)
      for (auto &ev : pr)
        if (ev) cudaEventDestroy(ev);
  }
  for (auto &ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i) {
    if (c->ev_fork[i]) cudaEventDestroy(c->ev_fork[i]);
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
  }
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->rstream) cudaStreamDestroy(c->rstream);
  if (c->bstream) cudaStreamDestroy(c->bstream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return GS_OK;
}

// ---------------------------------------------------------------------------------------------
// seam 1: clear / push / sort
// ---------------------------------------------------------------------------------------------
extern "C" int gs_clear(gs_context *c) {
  if (!c) return GS_ERR_INVALID;
  int rc0 = drain(c);
  if (rc0) return rc0;
  GS_CUDA(c, cudaStreamSynchronize(c->push_stream));
  c->n = 0;
  c->have_order = false;
  c->have_last_sorted = false;
  c->order_count = 0;
  return GS_OK;
}

extern "C" int gs_num_splats(const gs_context *c, uint32_t *out_n) {
  if (!c || !out_n) return GS_ERR_INVALID;
  *out_n = c->n;
  return GS_OK;
}

// pinned + device staging for the push path, created on first use
static int ensure_push_staging(gs_context *c) {
  if (c->push_dev[1]) return GS_OK;
  const size_t bytes = (size_t)gs_context::kPushRows * 32;
  for (int i = 0; i < 2; ++i) {
    if (!c->push_pinned[i]) GS_CUDA(c, cudaHostAlloc(&c->push_pinned[i], bytes, cudaHostAllocDefault));
    if (!c->push_dev[i]) GS_CUDA(c, cudaMalloc((void **)&c->push_dev[i], bytes));
  }
  return GS_OK;
}

extern "C" int gs_reserve(gs_context *c, uint32_t n_total) {
  if (!c) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  return ensure_table(c, n_total, /*exact=*/true);
}

extern "C" int gs_push_splats(gs_context *c, const void *rows32, uint32_t n) {
  if (!c || (!rows32 && n)) return GS_ERR_INVALID;
  if (!n) return GS_OK;
  GS_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = ensure_table(c, (uint64_t)c->n + n))) return rc;
  if ((rc = ensure_push_staging(c))) return rc;
  // Frames in flight keep rendering: they read rows [0, n_at_submit) only, the pack below writes rows >= c->n, on
  // its own stream.  Chunks alternate between two pinned staging buffers, so the host copy of chunk k+1 overlaps the
  // DMA + pack of chunk k.  The caller's buffer is fully consumed when this returns.
  const uint8_t *src = (const uint8_t *)rows32;
  for (uint32_t off = 0; off < n; off += gs_context::kPushRows) {
    const uint32_t m = std::min<uint32_t>(gs_context::kPushRows, n - off);
    const int b = c->push_buf;
    c->push_buf ^= 1;
    GS_CUDA(c, cudaEventSynchronize(c->push_ev[b]));  // this staging pair's previous chunk has been packed
    memcpy(c->push_pinned[b], src + (size_t)off * 32, (size_t)m * 32);
    GS_CUDA(c, cudaMemcpyAsync(c->push_dev[b], c->push_pinned[b], (size_t)m * 32, cudaMemcpyHostToDevice, c->push_stream));
    launch_pack(c, c->push_dev[b], c->n + off, m, c->push_stream);
    GS_CUDA(c, cudaGetLastError());
    GS_CUDA(c, cudaEventRecord(c->push_ev[b], c->push_stream));
  }
  GS_CUDA(c, cudaEventRecord(c->push_done, c->push_stream));
  c->n += n;
  c->pushed = true;
  c->have_order = false;
  return GS_OK;
}

extern "C" int gs_push_packed(gs_context *c, const float *center_scale4, const uint32_t *cov_color4,
                              const float *size_alpha, uint32_t n) {
  if (!c || ((!center_scale4 || !cov_color4 || !size_alpha) && n)) return GS_ERR_INVALID;
  if (!n) return GS_OK;
  GS_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = ensure_table(c, (uint64_t)c->n + n))) return rc;
  GS_CUDA(c, cudaMemcpyAsync(c->center_scale + c->n, center_scale4, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->push_stream));
  GS_CUDA(c, cudaMemcpyAsync(c->cov_color + c->n, cov_color4, sizeof(uint4) * (size_t)n, cudaMemcpyHostToDevice, c->push_stream));
  GS_CUDA(c, cudaMemcpyAsync(c->size_alpha + c->n, size_alpha, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->push_stream));
  GS_CUDA(c, cudaStreamSynchronize(c->push_stream));  // pageable sources: the caller may reuse them on return
  GS_CUDA(c, cudaEventRecord(c->push_done, c->push_stream));
  c->n += n;
  c->pushed = true;
  c->have_order = false;
  return GS_OK;
}

extern "C" int gs_read_packed(gs_context *c, uint32_t first, uint32_t n, float *center_scale4, uint32_t *cov_color4,
                              float *size_alpha) {
  if (!c || (uint64_t)first + n > c->n) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  GS_CUDA(c, cudaStreamSynchronize(c->push_stream));  // pushes are asynchronous
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

static int wait_slot(gs_context *c, gs_context::Slot &sl, gs_stats *stats);

// finish whatever is in flight (before buffers are reallocated or the splat table changes)
static int drain(gs_context *c) {
  for (auto &sl : c->slot)
    if (sl.pending) {
      int rc = wait_slot(c, sl, nullptr);
      if (rc) return rc;
    }
  return GS_OK;
}

static void stats_from_counters(gs_context *c, const FrameCounters &h, uint32_t n_splats) {
  gs_stats &s = c->stats;
  s.n_splats = n_splats;
  s.n_sorted = h.sort.n_valid;
  s.n_dropped = h.sort.n_dropped;
  s.n_visible = h.n_visible;
  s.n_instances = h.n_inst;
  s.n_instances_kept = h.n_inst_kept;
  s.min_depth = h.sort.n_valid ? dec_f64(~h.sort.min_enc) : INFINITY;
  s.max_depth = h.sort.n_valid ? dec_f64(h.sort.max_enc) : -INFINITY;
}

extern "C" int gs_sort(gs_context *c, const float view[4], const float *cutout16_or_null, uint32_t *out_idx,
                       uint32_t *out_count) {
  if (!c || !view) return GS_ERR_INVALID;
  if (c->n == 0) return fail(c, GS_ERR_EMPTY, "gs_sort before any push");
  GS_CUDA(c, cudaSetDevice(c->device));
  int rc = drain(c);
  if (rc) return rc;
  if ((rc = ensure_scratch(c))) return rc;
  GS_CUDA(c, cudaStreamSynchronize(c->bstream));
  GS_CUDA(c, cudaStreamSynchronize(c->rstream));
  gs_context::Slot &sl = c->slot[0];
  c->last_set = 0;
  const FrameBufs bufs{c->order[0], c->proj_rec[0], c->rect[0], c->inst_rec[0], c->bin_range[0]};
  memset(sl.fp_host, 0, sizeof(FrameParams));
  fill_sort_consts(sl.fp_host->sc, view, cutout16_or_null);
  sl.fp_host->n_splats = c->n;
  if (c->pushed) GS_CUDA(c, cudaStreamWaitEvent(c->stream, c->push_done, 0));
  GS_CUDA(c, cudaMemcpyAsync(sl.fp, sl.fp_host, sizeof(FrameParams), cudaMemcpyHostToDevice, c->stream));
  GS_CUDA(c, cudaMemsetAsync(sl.ctr, 0, sizeof(FrameCounters), c->stream));
  GS_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  launch_depth_cull(c, sl.fp, sl.ctr, c->stream);
  launch_depth_radix(c, sl.fp, sl.ctr, bufs, c->stream);
  GS_CUDA(c, cudaGetLastError());
  GS_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  GS_CUDA(c, cudaMemcpyAsync(c->sort_hdr, sl.ctr, sizeof(SortHeader), cudaMemcpyDeviceToDevice, c->stream));
  GS_CUDA(c, cudaMemcpyAsync(sl.ctr_host, sl.ctr, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->stream));
  GS_CUDA(c, cudaStreamSynchronize(c->stream));
  memset(&c->stats, 0, sizeof(c->stats));
  stats_from_counters(c, *sl.ctr_host, sl.fp_host->n_splats);
  c->stats.kernel_launches = 7;
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->stats.ms_sort = ms;
  c->stats.ms_total = ms;
  c->have_order = true;
  c->order_count = sl.ctr_host->sort.n_valid;
  if (out_count) *out_count = c->order_count;
  if (out_idx && c->order_count)
    GS_CUDA(c, cudaMemcpy(out_idx, c->order[c->last_set], sizeof(uint32_t) * (size_t)c->order_count, cudaMemcpyDeviceToHost));
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

static FrameBufs slot_bufs(gs_context *c, const gs_context::Slot &sl) {
  return FrameBufs{c->order[sl.set], c->proj_rec[sl.set], c->rect[sl.set], c->inst_rec[sl.set], c->bin_range[sl.set]};
}

// Stage A of a frame (sort stream): per-frame inputs to the device, depth sort, vertex shader.  All per-frame
// inputs come from sl.fp (device memory), so each stage is captured once into a CUDA graph and replayed.
static cudaError_t enqueue_sort_stage(gs_context *c, gs_context::Slot &sl, bool reuse, bool external_events) {
  auto rec = [&](cudaEvent_t ev, cudaStream_t st) {
    return external_events ? cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal) : cudaEventRecord(ev, st);
  };
  cudaStream_t m = c->stream, x = c->aux_stream;
  const FrameBufs b = slot_bufs(c, sl);
  cudaError_t e;
  if ((e = cudaMemcpyAsync(sl.fp, sl.fp_host, sizeof(FrameParams), cudaMemcpyHostToDevice, m))) return e;
  if ((e = cudaMemsetAsync(sl.ctr, 0, sizeof(FrameCounters), m))) return e;
  if ((e = rec(sl.ev[0], m))) return e;
  if (reuse) {
    if ((e = cudaMemcpyAsync(sl.ctr, c->sort_hdr, sizeof(SortHeader), cudaMemcpyDeviceToDevice, m))) return e;
  } else {
    launch_depth_cull(c, sl.fp, sl.ctr, m);
  }
  // fork: the vertex-shader kernel only needs the cull result, so it runs beside the depth radix passes
  if ((e = cudaEventRecord(c->ev_fork[0], m))) return e;
  if ((e = cudaStreamWaitEvent(x, c->ev_fork[0], 0))) return e;
  if ((e = rec(sl.evp[0], x))) return e;
  launch_project(c, sl.fp, sl.ctr, b, x);
  if ((e = rec(sl.evp[1], x))) return e;
  if ((e = cudaEventRecord(c->ev_join[0], x))) return e;
  if (!reuse) {
    launch_depth_radix(c, sl.fp, sl.ctr, b, m);
    if ((e = cudaMemcpyAsync(c->sort_hdr, sl.ctr, sizeof(SortHeader), cudaMemcpyDeviceToDevice, m))) return e;
  }
  if ((e = rec(sl.ev[1], m))) return e;
  if ((e = cudaStreamWaitEvent(m, c->ev_join[0], 0))) return e;
  return cudaGetLastError();
}

// Stage B (bin stream): tile instances in draw order, stable sort by tile, per-tile record lists and ranges.
static cudaError_t enqueue_bin_stage(gs_context *c, gs_context::Slot &sl, uint32_t n_bins, bool external_events) {
  auto rec = [&](cudaEvent_t ev, cudaStream_t st) {
    return external_events ? cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal) : cudaEventRecord(ev, st);
  };
  cudaStream_t m = c->bstream;
  const FrameBufs b = slot_bufs(c, sl);
  cudaError_t e;
  if ((e = cudaMemsetAsync(b.bin_range, 0, sizeof(uint2) * (size_t)n_bins, m))) return e;
  if ((e = rec(sl.ev[2], m))) return e;
  launch_emit(c, sl.fp, sl.ctr, b, m);   // 2 launches (k_emit also histograms pass T1)
  launch_tile_radix(c, sl.ctr, b, n_bins, c->emit_by_entry, m);    // (hist +) scan + scatter; above 256 bins also pass T2 + k_tile_ranges
  if ((e = rec(sl.ev[3], m))) return e;
  return cudaGetLastError();
}

// Stage C (raster stream, low priority): reads only this frame's inst_rec / bin_range copy.
static cudaError_t enqueue_raster_stage(gs_context *c, gs_context::Slot &sl, uint32_t n_tiles, bool external_events) {
  auto rec = [&](cudaEvent_t ev, cudaStream_t st) {
    return external_events ? cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal) : cudaEventRecord(ev, st);
  };
  cudaError_t e;
  if (sl.peer) launch_peer_acquire(c, sl.fp, sl.ctr, c->rstream);
  if ((e = rec(sl.ev_r0, c->rstream))) return e;
  launch_raster(c, sl.fp, n_tiles, slot_bufs(c, sl), sl.raster_flags, c->rstream);
  if ((e = rec(sl.ev[4], c->rstream))) return e;
  if (sl.peer) launch_peer_signal_wait(c, sl.fp, sl.ctr, c->rstream);
  return cudaGetLastError();
}

template <class F>
static int run_graph(gs_context *c, cudaGraphExec_t &ge, cudaStream_t stream, F enqueue) {
  if (c->use_graphs) {
    if (!ge) {
      cudaGraph_t g = nullptr;
      GS_CUDA(c, cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      cudaError_t e = enqueue(true);
      cudaError_t e2 = cudaStreamEndCapture(stream, &g);
      if (e != cudaSuccess || e2 != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        c->use_graphs = false;  // fall back to plain launches for the rest of this context's life
      } else {
        e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { ge = nullptr; cudaGetLastError(); c->use_graphs = false; }
      }
    }
    if (ge) {
      GS_CUDA(c, cudaGraphLaunch(ge, stream));
      return GS_OK;
    }
  }
  GS_CUDA(c, enqueue(false));
  return GS_OK;
}

// Three frames overlap: while frame k is rasterised (stream C), frame k+1 is binned (stream B) and frame k+2 is
// sorted / projected (stream A).  A and B are high priority: their short latency-bound kernels slot in as the
// long issue-bound raster's CTAs retire.  Stage hand-offs are events; buffers between stages are double-buffered.
static int launch_frame(gs_context *c, gs_context::Slot &sl, bool reuse, uint32_t n_tiles, uint32_t n_bins) {
  // (re)capture when anything baked into the launches changed
  gs_context::GraphKey k;
  k.cap = c->cap; k.n_tiles = n_tiles; k.n_bins = n_bins; k.cap_inst = c->cap_inst; k.p0 = c->depth; k.p1 = c->inst_rec[0]; k.p2 = c->center_scale;
  if (memcmp(&k, &c->gkey, sizeof(k)) != 0) {
    drop_graphs(c);
    c->gkey = k;
  }
  const int set = sl.set;
  // A: order/proj_rec/rect[set] must no longer be read by the binning stage that used them last
  if (c->sort_set_free[set]) GS_CUDA(c, cudaStreamWaitEvent(c->stream, c->sort_set_free[set], 0));
  int rc = run_graph(c, sl.graph_a[set][reuse ? 1 : 0], c->stream, [&](bool ext) { return enqueue_sort_stage(c, sl, reuse, ext); });
  if (rc) return rc;
  GS_CUDA(c, cudaEventRecord(sl.ev_sorted, c->stream));
  // B: needs A of this frame; inst_rec/bin_range[set] must no longer be read by the raster that used them last
  GS_CUDA(c, cudaStreamWaitEvent(c->bstream, sl.ev_sorted, 0));
  if (c->bin_set_free[set]) GS_CUDA(c, cudaStreamWaitEvent(c->bstream, c->bin_set_free[set], 0));
  if ((rc = run_graph(c, sl.graph_b[set], c->bstream, [&](bool ext) { return enqueue_bin_stage(c, sl, n_bins, ext); }))) return rc;
  GS_CUDA(c, cudaEventRecord(sl.ev_binned, c->bstream));
  c->sort_set_free[set] = sl.ev_binned;
  // C
  GS_CUDA(c, cudaStreamWaitEvent(c->rstream, sl.ev_binned, 0));
  if (sl.raster_flags == c->raster_base_flags) {
    if ((rc = run_graph(c, sl.peer ? sl.graph_rp[set] : sl.graph_r[set], c->rstream,
                        [&](bool ext) { return enqueue_raster_stage(c, sl, n_tiles, ext); }))) return rc;
  } else {
    // depth-tested / statistics frames use other instantiations of the raster: plain launches, no cached graph
    GS_CUDA(c, enqueue_raster_stage(c, sl, n_tiles, false));
  }
  sl.launches = (reuse ? 0u : 7u) + 1u + (n_bins <= 256u ? 4u : 8u) + (c->emit_by_entry ? 1u : 0u) + 1u;
  return GS_OK;
}

// entries covered by the first k slabs: slab_first * (1 + 2 + 4 + ...) - each slab is twice the previous one (4x growth
// was measured slower at 80 M splats: the few bins that stay open then force much larger slabs through sort + projection)
static uint64_t slab_cumulative(uint32_t first, int k) { return (uint64_t)first * ((1ull << k) - 1ull); }

// Stage A of a slab frame (sort stream): depth + cull, keys + bucket histogram, slab plan, pixel-state reset.
static cudaError_t enqueue_slab_keys_stage(gs_context *c, gs_context::Slot &sl, bool external_events) {
  auto rec = [&](cudaEvent_t ev, cudaStream_t st) {
    return external_events ? cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal) : cudaEventRecord(ev, st);
  };
  cudaStream_t st = c->stream;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(sl.fp, sl.fp_host, sizeof(FrameParams), cudaMemcpyHostToDevice, st))) return e;
  if ((e = cudaMemsetAsync(sl.ctr, 0, sizeof(FrameCounters), st))) return e;
  if ((e = rec(sl.ev[0], st))) return e;
  launch_depth_cull(c, sl.fp, sl.ctr, st);
  launch_keys(c, sl.fp, sl.ctr, sl.set, st);
  launch_slab_plan(c, sl.fp, sl.ctr, sl.set, c->slab_first, sl.n_slabs, st);
  launch_compact_offsets(c, sl.fp, sl.set, sl.n_slabs, st);  // one pass over the keys for every slab's compaction offsets
  if ((e = rec(sl.ev[1], st))) return e;
  return cudaGetLastError();
}

// Stage B+C of a slab frame (raster stream): the slab loop and the resolve.  One chain: every slab depends on the tiles
// the previous one closed.
static cudaError_t enqueue_slab_loop_stage(gs_context *c, gs_context::Slot &sl, uint32_t n_tiles, uint32_t n_bins,
                                           bool external_events) {
  auto rec = [&](cudaEvent_t ev, cudaStream_t st) {
    return external_events ? cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal) : cudaEventRecord(ev, st);
  };
  cudaStream_t st = c->rstream;
  const FrameBufs b = slot_bufs(c, sl);
  cudaError_t e;
  launch_slab_init(c, sl.fp, sl.ctr, st);
  if ((e = rec(sl.ev[2], st))) return e;
  for (int s = 0; s < sl.n_slabs; ++s) {
    launch_slab_begin(c, sl.fp, sl.ctr, sl.set, s, st);   // entry count (0 once every bin is closed) + compaction
    launch_slab_sort(c, sl.fp, sl.ctr, b, st);            // draw order of the slab
    launch_project_entries(c, sl.fp, sl.ctr, b, st);      // vertex shader for the slab's entries
    if ((e = cudaMemsetAsync(b.bin_range, 0, sizeof(uint2) * (size_t)n_bins, st))) return e;
    launch_emit_slab(c, sl.fp, sl.ctr, b, st);
    launch_tile_radix(c, sl.ctr, b, n_bins, true, st);  // k_emit_entries leaves T1's histogram to its own kernel
    if ((e = rec(sl.slab_ev[s][0], st))) return e;
    launch_raster_slab(c, sl.fp, sl.ctr, n_tiles, b, (sl.raster_flags & 2u) != 0, st);
    if ((e = rec(sl.slab_ev[s][1], st))) return e;
  }
  launch_slab_end(c, sl.ctr, st);
  if ((e = rec(sl.ev[3], st))) return e;
  if (sl.peer) launch_peer_acquire(c, sl.fp, sl.ctr, st);
  if ((e = rec(sl.ev_r0, st))) return e;
  launch_resolve(c, sl.fp, n_tiles, st);
  if ((e = rec(sl.ev[4], st))) return e;
  if (sl.peer) launch_peer_signal_wait(c, sl.fp, sl.ctr, st);
  return cudaGetLastError();
}

// Front-to-back slab path (gs_slab.cu).  Two stages: A (keys of every splat, O(N), sort stream) and the slab loop
// (raster stream); stage A of frame k+1 runs under the loop of frame k (keys / slab table are double-buffered by set).
static int launch_frame_slabs(gs_context *c, gs_context::Slot &sl, uint32_t n_tiles, uint32_t n_bins) {
  gs_context::GraphKey k;
  k.cap = c->cap; k.n_tiles = n_tiles; k.n_bins = n_bins; k.cap_inst = c->cap_inst; k.p0 = c->depth; k.p1 = c->inst_rec[0]; k.p2 = c->center_scale;
  if (memcmp(&k, &c->gkey, sizeof(k)) != 0) {
    drop_graphs(c);
    c->gkey = k;
  }
  const int set = sl.set;
  // slabs of slab_first, 2x, 4x ... entries: enough of them to cover every resident splat
  int n_slabs = 1;
  while (n_slabs < kMaxSlabs && slab_cumulative(c->slab_first, n_slabs) < sl.n_splats) ++n_slabs;
  if (n_slabs != sl.graph_slabs[set]) {  // the captured loop bakes the slab count
    auto kill = [](cudaGraphExec_t &g) { if (g) { cudaGraphExecDestroy(g); g = nullptr; } };
    kill(sl.graph_sa[set]); kill(sl.graph_sl[set][0]); kill(sl.graph_sl[set][1]); kill(sl.graph_sl[set][2]);
    sl.graph_slabs[set] = n_slabs;
  }
  sl.n_slabs = n_slabs;
  for (int s = 0; s < n_slabs; ++s)
    for (int q = 0; q < 2; ++q)
      if (!sl.slab_ev[s][q]) GS_CUDA(c, cudaEventCreate(&sl.slab_ev[s][q]));
  // A: the keys / slab table of this set must no longer be read by the loop that used them last
  if (c->sort_set_free[set]) GS_CUDA(c, cudaStreamWaitEvent(c->stream, c->sort_set_free[set], 0));
  int rc = run_graph(c, sl.graph_sa[set], c->stream, [&](bool ext) { return enqueue_slab_keys_stage(c, sl, ext); });
  if (rc) return rc;
  GS_CUDA(c, cudaEventRecord(sl.ev_sorted, c->stream));
  // loop: needs A of this frame; consecutive loops are ordered by the stream itself
  GS_CUDA(c, cudaStreamWaitEvent(c->rstream, sl.ev_sorted, 0));
  // graph variants of the loop: [plain, depth-tested, fused peer exchange]
  const int variant = sl.peer ? 2 : ((sl.raster_flags & 2u) ? 1 : 0);
  if (sl.peer && (sl.raster_flags & 2u)) {  // depth-tested peer frames: rare, plain launches
    GS_CUDA(c, enqueue_slab_loop_stage(c, sl, n_tiles, n_bins, false));
  } else if ((rc = run_graph(c, sl.graph_sl[set][variant], c->rstream,
                             [&](bool ext) { return enqueue_slab_loop_stage(c, sl, n_tiles, n_bins, ext); }))) {
    return rc;
  }
  GS_CUDA(c, cudaEventRecord(sl.ev_binned, c->rstream));
  c->sort_set_free[set] = sl.ev_binned;
  sl.launches = 6u + 1u + (uint32_t)n_slabs * (n_bins <= 256u ? 16u : 20u) + 2u;
  return GS_OK;
}

// after the frame's kernels: counters (and the frame, when the caller's buffer is host memory) go to the host on
// the copy stream, so the next frame's kernels overlap the PCIe transfer
static int enqueue_readback(gs_context *c, gs_context::Slot &sl) {
  GS_CUDA(c, cudaEventRecord(sl.ev_done, c->rstream));
  c->bin_set_free[sl.set] = sl.ev_done;
  GS_CUDA(c, cudaStreamWaitEvent(c->copy_stream, sl.ev_done, 0));
  GS_CUDA(c, cudaMemcpyAsync(sl.ctr_host, sl.ctr, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->copy_stream));
  if (sl.host_out)
    GS_CUDA(c, cudaMemcpyAsync(sl.out_user, sl.frame_src, sl.out_bytes, cudaMemcpyDeviceToHost, c->copy_stream));
  if (sl.raster_flags & 4u)
    GS_CUDA(c, cudaMemcpyAsync(c->tile_stats_host, c->tile_stats, sizeof(uint4) * (size_t)sl.fp_host->rc.n_tiles,
                               cudaMemcpyDeviceToHost, c->copy_stream));
  GS_CUDA(c, cudaEventRecord(sl.ev_copied, c->copy_stream));
  return GS_OK;
}

static int submit(gs_context *c, gs_context::Slot &sl) {
  const gs_render_params *p = &sl.params;
  FrameParams &fp = *sl.fp_host;
  RenderConsts &rc = fp.rc;
  memset(&fp, 0, sizeof(fp));
  fp.n_splats = sl.n_splats;  // what was resident when the frame was submitted; later pushes append behind it
  if (c->pushed) GS_CUDA(c, cudaStreamWaitEvent(c->stream, c->push_done, 0));
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
  rc.bins_x = (p->width + kBin - 1) / kBin;
  rc.bins_y = (p->height + kBin - 1) / kBin;
  rc.n_bins = rc.bins_x * rc.bins_y;
  memcpy(rc.bg, p->bg_rgba, sizeof(rc.bg));
  rc.shard_rank = c->shard_rank;
  rc.shard_world = c->shard_world;
  rc.out_format = p->out_format;
  rc.out_tiled = (p->flags & GS_RENDER_OUT_TILED) ? 1u : 0u;
  const float view[4] = {p->modelview[2], p->modelview[6], p->modelview[10], p->modelview[14]};  // index.js:442
  fill_sort_consts(fp.sc, view, p->has_cutout ? p->cutout16 : nullptr);

  int rcode;
  const size_t px_bytes = p->out_format == GS_FORMAT_RGBA8 ? 4 : 16;
  size_t out_pixels = (size_t)p->width * p->height;
  if (rc.out_tiled) out_pixels = (size_t)gs_owned_tiles(p->width, p->height, c->shard_rank, c->shard_world) * 256;
  sl.out_bytes = out_pixels * px_bytes;
  sl.host_out = !(p->flags & GS_RENDER_OUT_DEVICE);
  sl.peer = (p->flags & GS_RENDER_OUT_PEER) != 0;
  if (sl.peer) {
    if (!c->peer_world || c->peer_world != c->shard_world || c->peer_rank != c->shard_rank)
      return fail(c, GS_ERR_INVALID, "GS_RENDER_OUT_PEER needs gs_peer_import with the rank/world of gs_set_shard");
    if (rc.out_tiled) return fail(c, GS_ERR_INVALID, "GS_RENDER_OUT_PEER writes row-major frames: do not combine with GS_RENDER_OUT_TILED");
    if (sl.out_bytes > c->peer_frame_bytes) return fail(c, GS_ERR_INVALID, "frame larger than the exported peer frame size");
    // ring slot and sequence number come from the count of PEER frames, which every rank submits in the same order
    // (tickets also count each rank's private frames, e.g. warm-up)
    sl.ring = (int)(c->peer_count % 3);
    sl.peer_seq = c->peer_count + 1;
    fp.n_peer = c->peer_world;
    fp.peer_rank = c->peer_rank;
    for (uint32_t r = 0; r < c->peer_world; ++r) {
      fp.peer_out[r] = peer_frame(c->peer_base[r], c->peer_frame_bytes, sl.ring);
      fp.peer_done[r] = peer_done_row(c->peer_base[r], sl.ring);
      fp.peer_released[r] = peer_released_row(c->peer_base[r], sl.ring);
    }
    fp.local_done = peer_done_row(c->peer_local, sl.ring);
    fp.local_released = peer_released_row(c->peer_local, sl.ring);
    fp.peer_seq = sl.peer_seq;
    fp.peer_need = c->peer_count >= 3 ? c->peer_count - 2 : 0;  // sequence number of this ring slot's previous frame
    fp.out = peer_frame(c->peer_local, c->peer_frame_bytes, sl.ring);
    sl.frame_src = fp.out;
    c->peer_count += 1;
  } else if (sl.host_out) {
    if ((rcode = ensure_frame(c, sl, sl.out_bytes))) return rcode;
    fp.out = sl.frame_dev;
    sl.frame_src = sl.frame_dev;
  } else {
    fp.out = sl.out_user;
    sl.frame_src = nullptr;
  }
  // raster instantiation: pixel loop (packed fp32x2 by default), depth test, statistics
  sl.raster_flags = c->raster_base_flags | (p->depth_in ? 2u : 0u) | ((p->flags & GS_RENDER_STATS) ? 4u : 0u);
  if (p->depth_in) {
    if (p->flags & GS_RENDER_DEPTH_DEVICE) {
      fp.depth_in = p->depth_in;
    } else {  // host depth buffer: staged per slot, copied on the sort stream (the raster stage is ordered after it)
      const size_t bytes = sizeof(float) * (size_t)p->width * p->height;
      if (bytes > sl.depth_bytes || !sl.depth_dev) {
        if (sl.depth_dev) cudaFree(sl.depth_dev);
        sl.depth_dev = nullptr;
        GS_CUDA(c, cudaMalloc(&sl.depth_dev, bytes));
        sl.depth_bytes = bytes;
      }
      GS_CUDA(c, cudaMemcpyAsync(sl.depth_dev, p->depth_in, bytes, cudaMemcpyHostToDevice, c->stream));
      fp.depth_in = sl.depth_dev;
    }
  }
  const bool reuse = (p->flags & GS_RENDER_REUSE_SORT) && c->have_order;
  // a frame normally takes the buffer set the previous frame did not; a frame that reuses the last sort must read
  // that sort's set, so it runs in it
  sl.set = reuse ? c->last_set : (c->last_set ^ 1);
  if (sl.slab) {
    if ((rcode = launch_frame_slabs(c, sl, rc.n_tiles, rc.n_bins))) return rcode;
  } else {
    if ((rcode = launch_frame(c, sl, reuse, rc.n_tiles, rc.n_bins))) return rcode;
  }
  if ((rcode = enqueue_readback(c, sl))) return rcode;
  c->last_set = sl.set;
  sl.pending = true;
  c->have_order = !sl.slab;  // a slab frame leaves no complete draw order behind (GS_RENDER_REUSE_SORT then sorts again)
  return GS_OK;
}

static int wait_slot(gs_context *c, gs_context::Slot &sl, gs_stats *stats) {
  if (!sl.pending) return fail(c, GS_ERR_INVALID, "gs_wait: no frame in flight for this ticket");
  for (int attempt = 0;; ++attempt) {
    GS_CUDA(c, cudaEventSynchronize(sl.ev_copied));
    sl.pending = false;
    if (sl.peer) {
      // the frame has been published to (and consumed by) the other ranks: it cannot be silently re-run
      const bool bad = sl.ctr_host->overflow || sl.ctr_host->peer_timeout;
      PeerRows rows{};
      for (uint32_t r = 0; r < c->peer_world; ++r) rows.p[r] = peer_released_row(c->peer_base[r], sl.ring);
      launch_peer_release(c, rows, c->peer_world, c->peer_rank, sl.peer_seq, c->copy_stream);
      GS_CUDA(c, cudaGetLastError());
      if (sl.ctr_host->peer_timeout) return fail(c, GS_ERR_CUDA, "fused exchange: a peer did not signal in time");
      if (sl.ctr_host->overflow)
        return fail(c, GS_ERR_CAPACITY, "instance buffer overflow in a GS_RENDER_OUT_PEER frame: size the buffers with one plain frame first");
      (void)bad;
      break;
    }
    if (!sl.ctr_host->overflow) break;
    if (attempt == 7) return fail(c, GS_ERR_CAPACITY, "instance buffer kept overflowing");
    // Instance buffer too small.  Frames submitted BEFORE this one that are still pending saw the same small
    // buffer: finish (and, if needed, re-run) them first so frames are always re-run in submission order and
    // last_set / have_order end up describing the most recently submitted frame.
    for (;;) {
      gs_context::Slot *older = nullptr;
      for (auto &o : c->slot)
        if (&o != &sl && o.pending && o.ticket < sl.ticket && (!older || o.ticket < older->ticket)) older = &o;
      if (!older) break;
      int rc_old = wait_slot(c, *older, nullptr);
      if (rc_old) return rc_old;
    }
    GS_CUDA(c, cudaStreamSynchronize(c->stream));  // the other slots' frames may still be using the buffers
    GS_CUDA(c, cudaStreamSynchronize(c->bstream));
    GS_CUDA(c, cudaStreamSynchronize(c->rstream));
    // grow once to the measured demand (+12.5 %); a frame whose overflow flag is stale (an earlier frame's regrow
    // already made room) is simply run again
    const uint64_t demand = sl.slab ? sl.ctr_host->n_inst_slab_max : sl.ctr_host->n_inst;
    if (demand > c->cap_inst) {
      const uint64_t need = std::max<uint64_t>(demand + demand / 8, c->cap_inst + c->cap_inst / 2);
      int rcode = ensure_instances(c, need);
      if (rcode) return rcode;
    }
    int rcode;
    if ((rcode = submit(c, sl))) return rcode;
  }
  memset(&c->stats, 0, sizeof(c->stats));
  stats_from_counters(c, *sl.ctr_host, sl.fp_host->n_splats);
  c->stats.kernel_launches = sl.launches;
  c->stats.n_tiles = sl.fp_host->rc.n_tiles;
  if (sl.raster_flags & 4u) {  // GS_RENDER_STATS: per-tile {records streamed, records kept, pair tests, pair hits}
    const RenderConsts &rc = sl.fp_host->rc;
    for (uint32_t t = 0; t < rc.n_tiles; ++t) {
      if (rc.shard_world > 1 && ((t % rc.tiles_x) / kTilesPerBin) % rc.shard_world != rc.shard_rank) continue;
      const uint4 v = c->tile_stats_host[t];
      c->stats.n_records_streamed += v.x;
      c->stats.n_tile_instances += v.y;
      c->stats.n_pair_tests += v.z;
      c->stats.n_pair_hits += v.w;
    }
  }
  c->stats.width = sl.params.width;
  c->stats.height = sl.params.height;
  cudaEventElapsedTime(&c->stats.ms_sort, sl.ev[0], sl.ev[1]);
  if (sl.slab) {
    // slab path: ms_sort = depth/cull + keys + plan; ms_raster = the slabs' rasters + the resolve; ms_bin = the rest of
    // the slab loop (compaction, per-slab sort, projection, binning)
    float loop = 0.f, res = 0.f, rs = 0.f;
    cudaEventElapsedTime(&loop, sl.ev[2], sl.ev[3]);
    cudaEventElapsedTime(&res, sl.ev_r0, sl.ev[4]);
    for (int s = 0; s < sl.n_slabs; ++s) {
      float t = 0.f;
      cudaEventElapsedTime(&t, sl.slab_ev[s][0], sl.slab_ev[s][1]);
      rs += t;
    }
    c->stats.ms_project = 0.f;
    c->stats.n_slabs = (uint32_t)sl.n_slabs;
    c->stats.n_slabs_run = sl.ctr_host->slabs_run;
    c->stats.n_slab_entries = sl.ctr_host->slab_entries;
    c->stats.ms_raster = rs + res;
    c->stats.ms_bin = loop - rs;
  } else {
    cudaEventElapsedTime(&c->stats.ms_project, sl.evp[0], sl.evp[1]);  // on the aux stream, overlapping the sort
    cudaEventElapsedTime(&c->stats.ms_bin, sl.ev[2], sl.ev[3]);  // on the bin stream
    cudaEventElapsedTime(&c->stats.ms_raster, sl.ev_r0, sl.ev[4]);
  }
  cudaEventElapsedTime(&c->stats.ms_total, sl.ev[0], sl.ev[4]);
  c->order_count = sl.ctr_host->sort.n_valid;
  c->last_sorted = sl.ctr_host->sort.n_valid;
  c->have_last_sorted = true;
  if (stats) *stats = c->stats;
  return GS_OK;
}

extern "C" int gs_render_async(gs_context *c, const gs_render_params *p, void *out_rgba, uint64_t *out_ticket) {
  if (!c || !p || !out_rgba) return GS_ERR_INVALID;
  if (c->n == 0) return fail(c, GS_ERR_EMPTY, "gs_render before any push");
  if (p->width == 0 || p->height == 0 || p->width > 4096 || p->height > 4096)
    return fail(c, GS_ERR_INVALID, "frame size must be within 1..4096 per side");
  if (p->out_format != GS_FORMAT_RGBA8 && p->out_format != GS_FORMAT_RGBA32F) return fail(c, GS_ERR_INVALID, "bad out_format");
  const uint32_t n_tiles = ((p->width + kTile - 1) / kTile) * ((p->height + kTile - 1) / kTile);
  const uint32_t n_bins = ((p->width + kBin - 1) / kBin) * ((p->height + kBin - 1) / kBin);  // <= 64*64: fits the 16-bit bin id
  GS_CUDA(c, cudaSetDevice(c->device));
  const uint64_t ticket = c->next_ticket;
  gs_context::Slot &sl = c->slot[ticket % gs_context::kSlots];
  int rcode;
  if (sl.pending && (rcode = wait_slot(c, sl, nullptr))) return rcode;  // slot reuse: its previous frame must be done
  if ((p->flags & GS_RENDER_OUT_PEER) && ticket >= 3) {
    // the shared frame ring of the fused exchange has three entries, released by gs_wait: at most three such frames
    gs_context::Slot &o = c->slot[(ticket - 3) % gs_context::kSlots];
    if (o.pending && o.ticket == ticket - 3 && (rcode = wait_slot(c, o, nullptr))) return rcode;
  }
  if ((p->flags & GS_RENDER_REUSE_SORT) && c->have_order && (rcode = drain(c))) return rcode;  // runs in the last sort's buffers
  if ((p->flags & GS_RENDER_STATS) && (rcode = drain(c))) return rcode;  // the per-tile statistics buffer is not double-buffered
  // large scenes render front to back in depth slabs; the two paths share scratch buffers, so a change drains
  // (the criterion is the number of SORTED splats: the last frame's count when there is one, else the resident count)
  const uint32_t expect_sorted = c->have_last_sorted ? c->last_sorted : c->n;
  const bool slab = expect_sorted >= c->slab_min && !(p->flags & (GS_RENDER_REUSE_SORT | GS_RENDER_STATS));
  if ((int)slab != c->last_mode) {
    if ((rcode = drain(c))) return rcode;
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->bstream));
    GS_CUDA(c, cudaStreamSynchronize(c->rstream));
    c->last_mode = (int)slab;
  }
  // growing any shared buffer needs an idle pipeline
  const bool grow = (slab && (c->slab_cap < c->cap || !c->key32[0] || c->slab_tiles_cap < n_tiles || !c->slab_tab[1])) || !(c->scratch_cap >= c->cap && c->depth) || !(n_bins <= c->bins_cap && c->bin_range[0]) || !(n_tiles <= c->tile_stats_cap && c->tile_stats) || c->cap_inst == 0;
  if (grow) {
    if ((rcode = drain(c))) return rcode;
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->bstream));
    GS_CUDA(c, cudaStreamSynchronize(c->rstream));
    if ((rcode = ensure_scratch(c))) return rcode;
    if ((rcode = ensure_bins(c, n_bins))) return rcode;
    if ((rcode = ensure_tile_stats(c, n_tiles))) return rcode;
    if (slab && (rcode = ensure_slab(c, n_tiles, n_bins))) return rcode;
    if (c->cap_inst == 0) {
      // first frame: room for two bin instances per resident splat (a typical scene needs ~1); GS_INST_CAP overrides
      // the initial size (tests of the overflow / regrow path)
      uint64_t first = std::max<uint64_t>(1u << 20, slab ? (uint64_t)c->n : (uint64_t)c->n * 2);
      if (const char *e = getenv("GS_INST_CAP")) first = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));
      if ((rcode = ensure_instances(c, first))) return rcode;
    }
  }
  sl.params = *p;
  sl.out_user = out_rgba;
  sl.ticket = ticket;
  sl.n_splats = c->n;
  sl.slab = slab;
  if ((rcode = submit(c, sl))) return rcode;
  c->next_ticket = ticket + 1;
  if (out_ticket) *out_ticket = ticket;
  return GS_OK;
}

extern "C" int gs_wait(gs_context *c, uint64_t ticket, gs_stats *stats) {
  if (!c) return GS_ERR_INVALID;
  if (ticket >= c->next_ticket) return fail(c, GS_ERR_INVALID, "gs_wait: unknown ticket");
  GS_CUDA(c, cudaSetDevice(c->device));
  gs_context::Slot &sl = c->slot[ticket % gs_context::kSlots];
  if (ticket + gs_context::kSlots < c->next_ticket || !sl.pending || sl.ticket != ticket) {  // already completed (e.g. by a slot-reuse wait): stats of that frame are gone, frame is in place
    if (stats) *stats = c->stats;
    return GS_OK;
  }
  return wait_slot(c, sl, stats);
}

extern "C" int gs_render(gs_context *c, const gs_render_params *p, void *out_rgba, gs_stats *stats) {
  uint64_t t = 0;
  int rc = gs_render_async(c, p, out_rgba, &t);
  if (rc) return rc;
  return gs_wait(c, t, stats);
}

// XR: one sort request per frame from the head camera (tick(), index.js:438-455), one draw per eye with that eye's
// matrices and viewport (onBeforeRender per eye camera, index.js:184-195)
extern "C" int gs_render_stereo(gs_context *c, const float view[4], const float *cutout16_or_null,
                                const gs_render_params eyes[2], void *const out_rgba[2], gs_stats *stats2_or_null) {
  if (!c || !view || !eyes || !out_rgba || !out_rgba[0] || !out_rgba[1]) return GS_ERR_INVALID;
  int rc = gs_sort(c, view, cutout16_or_null, nullptr, nullptr);
  if (rc) return rc;
  const float ms_sort = c->stats.ms_sort;
  for (int e = 0; e < 2; ++e) {
    gs_render_params p = eyes[e];
    p.flags |= GS_RENDER_REUSE_SORT;  // both eyes draw with the head camera's order
    gs_stats st;
    if ((rc = gs_render(c, &p, out_rgba[e], &st))) return rc;
    if (stats2_or_null) {
      stats2_or_null[e] = st;
      stats2_or_null[e].ms_sort = e == 0 ? ms_sort : 0.0f;  // the one sort is accounted to the first eye
    }
  }
  return GS_OK;
}

extern "C" int gs_peer_export(gs_context *c, size_t frame_bytes, void *handle_out) {
  if (!c || !handle_out || !frame_bytes) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  int rc = drain(c);
  if (rc) return rc;
  if (c->peer_local) return fail(c, GS_ERR_INVALID, "gs_peer_export: already exported");
  frame_bytes = (frame_bytes + 4095) / 4096 * 4096;
  const size_t total = kPeerFlagBytes + 3 * frame_bytes;
  GS_CUDA(c, cudaMalloc(&c->peer_local, total));
  GS_CUDA(c, cudaMemset(c->peer_local, 0, total));
  c->peer_frame_bytes = frame_bytes;
  cudaIpcMemHandle_t h;
  GS_CUDA(c, cudaIpcGetMemHandle(&h, c->peer_local));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, 64);
  return GS_OK;
}

extern "C" int gs_peer_import(gs_context *c, uint32_t rank, uint32_t world, const void *handles) {
  if (!c || !handles || world == 0 || world > (uint32_t)kMaxPeers || rank >= world) return GS_ERR_INVALID;
  if (!c->peer_local) return fail(c, GS_ERR_INVALID, "gs_peer_import before gs_peer_export");
  GS_CUDA(c, cudaSetDevice(c->device));
  for (uint32_t r = 0; r < world; ++r) {
    if (r == rank) { c->peer_base[r] = c->peer_local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + 64 * (size_t)r, 64);
    GS_CUDA(c, cudaIpcOpenMemHandle(&c->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  c->peer_rank = rank;
  c->peer_world = world;
  return GS_OK;
}

extern "C" int gs_peer_frame(gs_context *c, uint64_t ticket, void **out) {
  if (!c || !out || !c->peer_local || ticket >= c->next_ticket || ticket + 3 < c->next_ticket) return GS_ERR_INVALID;
  const gs_context::Slot &sl = c->slot[ticket % gs_context::kSlots];
  if (!sl.peer || sl.ticket != ticket) return GS_ERR_INVALID;
  *out = peer_frame(c->peer_local, c->peer_frame_bytes, sl.ring);
  return GS_OK;
}

extern "C" int gs_get_stats(const gs_context *c, gs_stats *out) {
  if (!c || !out) return GS_ERR_INVALID;
  *out = c->stats;
  return GS_OK;
}

extern "C" int gs_read_projected(gs_context *c, uint32_t first, uint32_t n, float *out8) {
  if (!c || !out8 || (uint64_t)first + n > c->n || !c->proj_rec[0]) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  std::vector<float> rec((size_t)n * 8);
  std::vector<uint32_t> rect(n);
  GS_CUDA(c, cudaMemcpy(rec.data(), c->proj_rec[c->last_set] + 2 * (size_t)first, sizeof(float) * 8 * (size_t)n, cudaMemcpyDeviceToHost));
  GS_CUDA(c, cudaMemcpy(rect.data(), c->rect[c->last_set] + first, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
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
  GS_CUDA(c, cudaGetLastError());  // stream-ordered: complete after gs_synchronize / any later stream work
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
  GS_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->rstream));
  GS_CUDA(c, cudaStreamSynchronize(c->rstream));
  return GS_OK;
}
// frames complete on the raster stream: external work ordered after a frame (collectives, copies) goes there
extern "C" void *gs_stream(gs_context *c) { return c ? (void *)c->rstream : nullptr; }
extern "C" int gs_synchronize(gs_context *c) {
  if (!c) return GS_ERR_INVALID;
  GS_CUDA(c, cudaSetDevice(c->device));
  GS_CUDA(c, cudaStreamSynchronize(c->stream));
  GS_CUDA(c, cudaStreamSynchronize(c->bstream));
  GS_CUDA(c, cudaStreamSynchronize(c->rstream));
  return GS_OK;
}

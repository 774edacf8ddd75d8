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
#include <stddef.h>
#include <string.h>

#include <string>

#include "../../include/gsplat_b200.h"

namespace gs {

constexpr int kTile = 16;                 // 16x16 screen tiles (north_star): one raster CTA per tile
// Binning granularity: splats are binned to square BINS of GS_BIN_TILES x GS_BIN_TILES tiles (96x96 pixels by default),
// not to tiles.  A splat meets ~5x fewer bins than tiles, so the instance emission and the stable sort by bin id handle
// ~5x fewer elements; each tile's raster CTA streams its bin's list and culls it against its own 16x16 pixels on the fly
// (exact footprint test, one record per thread).  Measured at config 2 (1 M splats, 1080p): 64 px 3323, 96 px 3601,
// 128 px 3626 frames/s, raster time unchanged; a 1920x1080 frame has 240 bins of 96 px, so the bin id is one byte and one
// radix pass sorts the instances.  96 px keeps more bin columns for multi-GPU ownership and closes bins earlier (slab path).
#ifndef GS_BIN_TILES
#define GS_BIN_TILES 6
#endif
constexpr int kTilesPerBin = GS_BIN_TILES;  // tile columns / rows per bin
constexpr int kBin = kTile * kTilesPerBin;  // bin edge in pixels (96 by default; gs_bin_size() reports it)
constexpr int kRadixThreads = 256;
#ifndef GS_RADIX_ITEMS
#define GS_RADIX_ITEMS 16
#endif
constexpr int kRadixItems = GS_RADIX_ITEMS;  // elements per thread of a radix chunk (chunk = 256 x this)
constexpr int kRadixTile = kRadixThreads * kRadixItems;  // 4096 elements per radix chunk
constexpr int kEmitThreads = 256;
constexpr int kEmitItems = 1;
constexpr int kEmitTile = kEmitThreads * kEmitItems;      // 256 draw-order entries per slice of the instance-offset scan
constexpr uint32_t kInvalidDigit = 0xFFFFFFFFu;  // element dropped by a radix pass
constexpr uint32_t kNoRect = 0xFFFFFFFFu;
constexpr uint16_t kNoTile = 0xFFFFu;
// depth sentinel: a splat rejected by the worker filter (index.js:548 keeps only depth < 0)
#define GS_DEPTH_REJECT 1.0f

// preserved part of FrameCounters when a frame reuses the previous draw order (GS_RENDER_REUSE_SORT): the sort's own
// results.  FrameCounters starts with exactly this header, so a sizeof(SortHeader) device copy saves / restores it.
struct SortHeader {
  unsigned long long min_enc;  // bit-inverted order-preserving encoding of the fp64 min depth (atomicMax)
  unsigned long long max_enc;  // order-preserving encoding of the fp64 max depth (atomicMax)
  uint32_t n_valid;            // V: splats passing the worker filter
  uint32_t n_inrange;          // V - dropped: entries with a key in [0,65535]
  uint32_t n_dropped;          // quirk Q5
  uint32_t pad;
};

// Device-resident per-frame counters: zeroed by one memset at the start of every sort/render.
struct FrameCounters {
  SortHeader sort;             // min_enc, max_enc, n_valid, n_inrange, n_dropped
  unsigned long long n_inst;   // D: emitted tile instances (bounding-rectangle candidates)
  uint32_t n_visible;          // V2
  uint32_t n_inst_kept;        // instances surviving the exact footprint test and the tile-ownership filter
  uint32_t overflow;           // instance buffer too small: frame must be re-run
  uint32_t peer_timeout;       // a peer flag was not seen in time (fused exchange)
  uint32_t count_done;         // k_count CTAs finished (the last one scans the slice totals)
  // ---- front-to-back slab path (large scenes): see gs_slab.cu ----
  uint32_t open_bins;          // bins of this rank that still have a live pixel
  uint32_t slab_real;          // real entries of the current slab (it may also hold quirk-Q5 repeats of splat 0)
  uint32_t total_valid;        // V and V - dropped of the whole frame (sort.n_valid / n_inrange hold the CURRENT slab's
  uint32_t total_inrange;      //   entry count while the slab loop runs, so the sort / emit kernels work unchanged)
  uint32_t n_kept_total;       // bin instances kept, summed over the slabs
  uint32_t slabs_run;          // slabs that found open bins and entries
  unsigned long long n_inst_total;  // bin-instance candidates, summed over the slabs
  unsigned long long n_inst_slab_max;  // ... of the largest slab (what the instance buffers must hold)
  unsigned long long slab_entries;     // entries (real + Q5 repeats) of the slabs that ran: what was compacted, sorted, projected
};
static_assert(offsetof(FrameCounters, sort) == 0 && sizeof(SortHeader) == 32, "SortHeader is the prefix of FrameCounters");

struct RenderConsts {
  float proj[16];
  float mv[16];
  float vw, vh, focal;
  uint32_t width, height;
  uint32_t tiles_x, tiles_y, n_tiles;
  uint32_t bins_x, bins_y, n_bins;
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

// Per-frame inputs, resident in device memory (one copy per pipeline slot) so that the whole frame is a static
// CUDA graph: a 400-byte host->device copy of this struct is the only per-frame input traffic.
constexpr int kMaxPeers = 16;

struct FrameParams {
  SortConsts sc;
  RenderConsts rc;
  uint32_t n_splats;  // resident splats of THIS frame (the table may be growing behind it: progressive push)
  void *out;  // frame (or packed owned tiles) destination of the raster
  const void *depth_in;  // optional window-space depth of foreign geometry (f32, width*height, row 0 = bottom)
  // ---- fused raster + exchange over NVLink peer memory (GS_RENDER_OUT_PEER) ----
  uint32_t n_peer;                       // 0: plain output; else every finished tile is stored into all ranks' frames
  uint32_t peer_rank;
  void *peer_out[kMaxPeers];             // this frame's slot in every rank's shared frame ring (own rank included)
  unsigned long long *peer_done[kMaxPeers];      // rank r's done[slot][*] flag row (we write [.][peer_rank])
  unsigned long long *peer_released[kMaxPeers];  // rank r's released[slot][*] flag row
  unsigned long long *local_done;        // our done[slot][*]
  unsigned long long *local_released;    // our released[slot][*]
  unsigned long long peer_seq;           // ticket + 1 of this frame
  unsigned long long peer_need;          // slot may be overwritten once every rank released seq >= peer_need
};


// ---- front-to-back slab path ----
constexpr int kMaxSlabs = 12;          // geometric slab sizes: 1 M, 2 M, 4 M ... entries (nearest first)
constexpr int kSlabBuckets = 4096;     // slab boundaries are chosen on a 4096-bucket histogram of the 16-bit keys
constexpr uint32_t kNoKey = 0xFFFFFFFFu;
struct SlabTable {
  uint32_t hist[kSlabBuckets];  // entries per 16-key bucket
  uint32_t klo[kMaxSlabs];      // slab s holds the keys [klo[s], khi[s])
  uint32_t khi[kMaxSlabs];
  uint32_t count[kMaxSlabs];    // entries of slab s
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
  // outputs of the sort/project stage, consumed by the binning stage of the same frame: double-buffered so that
  // frame k+1 is sorted while frame k is binned
  uint32_t *order[2] = {nullptr, nullptr};    // draw order (== reference sortedIndexes)
  float4 *proj_rec[2] = {nullptr, nullptr};   // 2 x float4 per splat
  uint32_t *rect[2] = {nullptr, nullptr};     // packed tile rect per splat
  uint32_t *table_n = nullptr;   // radix chunk histograms of the depth passes [256][table_n_stride]
  uint32_t table_n_stride = 0;
  uint32_t *totals = nullptr;    // [512]: digit totals of the depth / tile passes
  uint32_t *slice_total = nullptr;   // instances per 256-entry slice of the draw order
  uint32_t *slice_prefix = nullptr;  // exclusive scan of slice_total (+ total at the end)
  uint2 *ent = nullptr;              // per draw-order entry: {splat index, packed rect or kNoRect}
  uint32_t *ent_off = nullptr;       // per entry: exclusive instance offset inside its slice

  // ---- per-instance scratch (sized to cap_inst) ----
  uint64_t cap_inst = 0;
  uint16_t *inst_tile = nullptr;
  uint32_t *inst_idx = nullptr;
  uint16_t *inst_tile_b = nullptr;  // tile id carried through pass T1
  uint16_t *inst_tile_f = nullptr;  // tile id of every instance in final (tile, draw order) order
  uint32_t *inst_idx_b = nullptr;
  float4 *inst_rec[2] = {nullptr, nullptr};  // 2 x float4 per instance, sorted by (tile, draw order); one per slot:
                                             // the raster of frame k reads [k&1] while frame k+1 is binned into the other
  uint32_t *table_d = nullptr;   // radix chunk histograms of the tile passes [256][table_d_stride]
  uint32_t table_d_stride = 0;

  // ---- per-frame tables ----
  uint32_t bins_cap = 0;
  uint2 *bin_range[2] = {nullptr, nullptr};  // [bins] {start, end} into inst_rec, one per slot (read by the raster)
  // ---- front-to-back slab path (gs_slab.cu): allocated when a scene first crosses the slab threshold ----
  uint32_t slab_cap = 0;           // splats the per-splat slab buffers are sized for
  uint32_t *key32[2] = {nullptr, nullptr};  // [cap] 16-bit depth key of every splat, kNoKey if not in the sort (one per set)
  uint32_t *cidx = nullptr;        // [cap] splat indices of the current slab, in index order
  uint16_t *ckey = nullptr;        // [cap] their keys
  uint32_t *chunk_cnt[2] = {nullptr, nullptr};  // [kMaxSlabs][chunk_row] per-slab compaction offsets of every 2048-splat chunk (one per set)
  uint32_t chunk_row = 0;          // row stride of chunk_cnt: cap / 2048 + 4
  gs::SlabTable *slab_tab[2] = {nullptr, nullptr};
  float4 *pix_state = nullptr;     // [tiles * 256] {R, G, B, T} carried from slab to slab
  uint8_t *tile_closed = nullptr;  // [tiles]
  uint32_t *bin_open = nullptr;    // [bins] live tiles per bin (0 for bins of other ranks)
  uint32_t slab_tiles_cap = 0;
  uint32_t slab_min = 16u << 20;   // frames expected to SORT at least this many splats render front to back in slabs
  uint32_t last_sorted = 0;        // V of the most recently completed frame (predicts the next frame's)
  bool have_last_sorted = false;
  uint32_t slab_first = 1u << 20;  // target entry count of the nearest slab (the following ones double)
  int last_mode = 0;               // 0 = one pass (three-stage pipeline), 1 = slab path
  bool emit_by_entry = true;       // k_emit_entries + k_radix_hist<T1> (default; measured faster on every config) or, with
                                   // GS_EMIT=windows, the window-balanced k_emit of round 1 (one-pass path only)
  uint4 *tile_stats = nullptr;     // [tiles] per-tile counts of a GS_RENDER_STATS frame
  uint4 *tile_stats_host = nullptr;  // pinned copy
  uint32_t tile_stats_cap = 0;
  double *quirk_table = nullptr;   // parseInt quirk thresholds (device)
  int quirk_n = 0;
  gs::SortHeader *sort_hdr = nullptr;  // device: counters header of the last sort (for GS_RENDER_REUSE_SORT)

  // ---- pipeline slots (ticket % kSlots): frame k is rasterised while k+1 is binned, k+2 is sorted and k-1 is copied to
  // the host.  Three stages + the copy = four frames a caller can have outstanding; the GPU-side order of the stages is
  // kept by the buffer-set events below, a slot only holds a frame's parameters, counters, events and output staging.
  // (With three slots a caller that receives frames in host memory had to collect frame k-1's copy before it could
  // submit frame k+2, and the sort stage idled for the length of the copy: 3 200 instead of 3 800 frames/s.) ----
  static constexpr int kSlots = 4;
  struct Slot {
    gs::FrameCounters *ctr = nullptr;        // device
    gs::FrameCounters *ctr_host = nullptr;   // pinned
    gs::FrameParams *fp = nullptr;           // device
    gs::FrameParams *fp_host = nullptr;      // pinned staging
    void *frame_dev = nullptr;               // used when the caller's buffer is host memory
    size_t frame_bytes = 0;
    void *depth_dev = nullptr;               // staging of a host depth_in
    size_t depth_bytes = 0;
    uint32_t raster_flags = 0;               // k_raster instantiation of this frame (packed | depth | stats)
    uint32_t n_splats = 0;                   // resident splats when the frame was submitted
    bool slab = false;                       // rendered by the front-to-back slab path
    int n_slabs = 0;
    cudaEvent_t slab_ev[gs::kMaxSlabs][2] = {};  // raster of each slab (timing)
    cudaEvent_t ev[5]{};                     // stage boundaries (timing)
    cudaEvent_t evp[2]{};                    // k_project on the aux stream (timing)
    cudaEvent_t ev_done = nullptr, ev_copied = nullptr;
    // CUDA graphs of the three stages, one per buffer set this slot can be paired with
    cudaGraphExec_t graph_a[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // sort + project, [set][reuse_sort]
    cudaGraphExec_t graph_b[2] = {nullptr, nullptr};                           // binning, [set]
    cudaGraphExec_t graph_r[2] = {nullptr, nullptr};                           // raster, [set]
    cudaGraphExec_t graph_rp[2] = {nullptr, nullptr};                          // acquire + raster + signal/wait (fused exchange)
    cudaGraphExec_t graph_sa[2] = {nullptr, nullptr};                          // slab path: keys stage, [set]
    cudaGraphExec_t graph_sl[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // slab path: slab loop + resolve, [set][plain | depth | peer]
    int graph_slabs[2] = {0, 0};                                               // slab count baked into graph_sl
    bool peer = false;
    uint64_t ticket = 0;
    int ring = 0;                            // slot of the shared frame ring (fused exchange)
    unsigned long long peer_seq = 0;
    void *frame_src = nullptr;               // device buffer the host copy reads
    cudaEvent_t ev_sorted = nullptr;                // sort/project stage of this slot's frame finished
    cudaEvent_t ev_binned = nullptr;                // binning stage finished
    cudaEvent_t ev_r0 = nullptr;                    // raster start (timing)
    int index = 0;
    bool pending = false;
    bool host_out = false;
    void *out_user = nullptr;
    size_t out_bytes = 0;
    gs_render_params params{};
    uint32_t launches = 0;
    int set = 0;                                    // which order/proj_rec/rect and inst_rec/bin_range copy it uses
  } slot[kSlots];
  uint64_t next_ticket = 0;
  cudaStream_t bstream = nullptr;                   // binning stage (high priority, like the sort stage's `stream`)
  cudaEvent_t sort_set_free[2] = {nullptr, nullptr};  // last binning stage that read order/proj_rec/rect[i]
  cudaEvent_t bin_set_free[2] = {nullptr, nullptr};   // last raster that read inst_rec/bin_range[i]
  int last_set = 0;                                 // set holding the most recent sort (GS_RENDER_REUSE_SORT, read-backs)
  cudaStream_t rstream = nullptr;   // raster stream: frame k is rasterised here while frame k+1 is sorted / binned
  cudaStream_t copy_stream = nullptr;
  // ---- progressive push (index.js:259-298, 576-586): rows go host -> pinned staging -> device staging -> k_pack on
  // their own stream while frames keep rendering the prefix that was resident when they were submitted ----
  cudaStream_t push_stream = nullptr;
  static constexpr uint32_t kPushRows = 1u << 18;  // rows per staging buffer (8 MiB)
  void *push_pinned[2] = {nullptr, nullptr};
  uint8_t *push_dev[2] = {nullptr, nullptr};
  cudaEvent_t push_ev[2] = {nullptr, nullptr};      // staging buffer i is free again
  cudaEvent_t push_done = nullptr;                  // everything pushed so far is packed
  int push_buf = 0;
  bool pushed = false;
  cudaStream_t aux_stream = nullptr;             // runs k_project beside the depth radix passes
  cudaEvent_t ev_fork[2]{}, ev_join[2]{};
  bool use_graphs = true;
  bool use_pdl = false;                          // programmatic dependent launch inside the stage chains (GS_PDL=1 turns it on)
  uint32_t raster_base_flags = 1;                // default pixel loop: 1 = packed fp32x2, 0 = scalar
  // graph cache key: anything baked into the captured launches
  struct GraphKey { uint32_t cap = 0, n_tiles = 0, n_bins = 0, pad = 0; uint64_t cap_inst = 0; const void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr; } gkey;

  // ---- fused exchange: one shared allocation per rank = flag rows + a ring of 3 frames, opened by every peer ----
  void *peer_local = nullptr;            // our shared block
  size_t peer_frame_bytes = 0;
  void *peer_base[gs::kMaxPeers] = {};   // every rank's shared block as mapped here (own rank: peer_local)
  uint32_t peer_world = 0, peer_rank = 0;
  unsigned long long peer_count = 0;     // GS_RENDER_OUT_PEER frames submitted so far (identical on every rank)

  uint32_t shard_rank = 0, shard_world = 1;
  bool have_order = false;
  uint32_t order_count = 0;
  gs_stats stats{};
  cudaEvent_t ev[2]{};  // gs_sort timing
};

namespace gs {

// shared block layout: [done: 3 slots x kMaxPeers u64][released: 3 slots x kMaxPeers u64][pad to 4 KiB][frame 0][frame 1][frame 2]
constexpr size_t kPeerFlagBytes = 4096;
inline unsigned long long *peer_done_row(void *base, int slot) { return (unsigned long long *)base + (size_t)slot * kMaxPeers; }
inline unsigned long long *peer_released_row(void *base, int slot) { return (unsigned long long *)base + (size_t)(3 + slot) * kMaxPeers; }
inline void *peer_frame(void *base, size_t frame_bytes, int slot) { return (char *)base + kPeerFlagBytes + (size_t)slot * frame_bytes; }

// buffers one frame's stages hand to each other (a pair of double-buffered sets)
struct FrameBufs {
  uint32_t *order;
  float4 *proj_rec;
  uint32_t *rect;
  float4 *inst_rec;
  uint2 *bin_range;  // [n_bins] {start, end} of each bin's run in inst_rec
};

// -- launchers (each .cu file owns its kernels); every per-frame input comes from device memory (fp, ctr) --
void launch_depth_cull(gs_context *c, const FrameParams *fp, FrameCounters *ctr, cudaStream_t st);
void launch_depth_radix(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st);  // 6 launches -> b.order
void launch_pack(gs_context *c, const uint8_t *rows_dev, uint32_t first, uint32_t n, cudaStream_t st);
void launch_project(gs_context *c, const FrameParams *fp, const FrameCounters *ctr, const FrameBufs &b, cudaStream_t st);
void launch_emit(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st);  // 2 launches
void launch_tile_radix(gs_context *c, FrameCounters *ctr, const FrameBufs &b, uint32_t n_bins, bool hist_t1, cudaStream_t st);  // 2 .. 7 launches
void launch_tile_ranges(gs_context *c, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st);
void launch_raster(gs_context *c, const FrameParams *fp, uint32_t n_tiles, const FrameBufs &b, uint32_t flags, cudaStream_t st);
void launch_peer_acquire(gs_context *c, const FrameParams *fp, FrameCounters *ctr, cudaStream_t st);
void launch_peer_signal_wait(gs_context *c, const FrameParams *fp, FrameCounters *ctr, cudaStream_t st);
struct PeerRows { unsigned long long *p[kMaxPeers]; };
void launch_peer_release(gs_context *c, const PeerRows &rows, uint32_t world, uint32_t rank, unsigned long long seq,
                         cudaStream_t st);
// ---- slab path launchers (gs_slab.cu / gs_raster.cu) ----
void launch_keys(gs_context *c, const FrameParams *fp, FrameCounters *ctr, int set, cudaStream_t st);  // keys + bucket histogram
void launch_slab_plan(gs_context *c, const FrameParams *fp, FrameCounters *ctr, int set, uint32_t first_target, int n_slabs, cudaStream_t st);
void launch_slab_init(gs_context *c, const FrameParams *fp, FrameCounters *ctr, cudaStream_t st);
void launch_compact_offsets(gs_context *c, const FrameParams *fp, int set, int n_slabs, cudaStream_t st);  // every slab's chunk offsets: 2 launches
void launch_slab_begin(gs_context *c, const FrameParams *fp, FrameCounters *ctr, int set, int slab, cudaStream_t st);  // + compaction: 2 launches
void launch_slab_sort(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st);  // 6 launches
void launch_project_entries(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st);
void launch_emit_slab(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st);
void launch_slab_end(gs_context *c, FrameCounters *ctr, cudaStream_t st);
void launch_raster_slab(gs_context *c, const FrameParams *fp, FrameCounters *ctr, uint32_t n_tiles, const FrameBufs &b, bool depth,
                        cudaStream_t st);
void launch_resolve(gs_context *c, const FrameParams *fp, uint32_t n_tiles, cudaStream_t st);
void launch_assemble(gs_context *c, const void *gathered, uint32_t tiles_per_rank, uint32_t world, uint32_t width,
                     uint32_t height, int32_t format, void *out_frame);

// ---- programmatic dependent launch (PDL): the kernels of a stage form a chain of short dependent launches.  Launched
// with the programmatic-stream-serialization attribute, kernel k+1 is set up (CTAs scheduled, arguments loaded) while
// kernel k drains; it blocks in pdl_wait() until k has completed and its writes are visible.  Every kernel launched
// this way calls pdl_trigger() + pdl_wait() before its first global access.  OFF by default (GS_PDL=1 enables): measured at
// config 2 the isolated sort stage gains 8 % (0.094 -> 0.086 ms) but the pipelined frame rate drops 12 % (3784 -> 3335):
// early-launched CTAs sit on SM resources while they wait, and those are the resources the co-running raster needs. ----
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define GS_PDL_ENTRY() do { gs::pdl_trigger(); gs::pdl_wait(); } while (0)

template <class... KArgs, class... Args>
inline cudaError_t launch_chain(gs_context *c, void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = c->use_pdl ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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

// Exact footprint-vs-box test shared by the bin emission (kBin x kBin box) and the raster's cull (16x16 box).
// The box holds pixel CENTRES [x0, x0 + extent] x [y0, y0 + extent]; the footprint is the set r^2 = px^2 + py^2 <= 4 with
// (px, py) = (d.a2, d.a1), d = sample - centre (index.js:158-172).  r^2 is a convex quadratic of d, so when the centre
// lies outside the box its minimum over the box is attained on the edge(s) facing the centre; the test evaluates those
// minima in plain fp32 and keeps a 0.5 % slack (4.02) so that it can only err on the side of keeping the splat.
__device__ __forceinline__ bool footprint_meets_box(float cx, float cy, float a1x, float a1y, float a2x, float a2y, float x0,
                                                    float y0, float extent) {
  const float xa = x0 - cx, xb = xa + extent;
  const float ya = y0 - cy, yb = ya + extent;
  const bool in_x = (xa <= 0.0f) && (xb >= 0.0f), in_y = (ya <= 0.0f) && (yb >= 0.0f);
  if (in_x && in_y) return true;
  // q(d) = |(a2.d, a1.d)|^2 = M00 dx^2 + 2 M01 dx dy + M11 dy^2: edge minimisers need M01/M11, M01/M00
  const float cross = a2x * a2y + a1x * a1y;
  float qmin = 3.0e38f;
  if (!in_x) {  // nearest vertical edge, minimise over y on it; (px,py) evaluated at the found point
    const float dx = (xa > 0.0f) ? xa : xb;
    const float inv_yy = __fdividef(1.0f, a2y * a2y + a1y * a1y);
    const float t = fminf(fmaxf(-dx * cross * inv_yy, ya), yb);
    const float px = dx * a2x + t * a2y, py = dx * a1x + t * a1y;
    qmin = px * px + py * py;
  }
  if (!in_y) {  // nearest horizontal edge, minimise over x on it
    const float dy = (ya > 0.0f) ? ya : yb;
    const float inv_xx = __fdividef(1.0f, a2x * a2x + a1x * a1x);
    const float t = fminf(fmaxf(-dy * cross * inv_xx, xa), xb);
    const float px = t * a2x + dy * a2y, py = t * a1x + dy * a1y;
    qmin = fminf(qmin, px * px + py * py);
  }
  return !(qmin > 4.02f);  // NaN keeps
}

// ---- multi-GPU ownership: rank r owns the BIN COLUMNS bx with bx % world == r (kBin-pixel wide vertical stripes,
// interleaved), so the owned bins of any bin rectangle have a closed form; a tile belongs to the owner of its bin ----
__host__ __device__ inline uint32_t owned_cols(uint32_t bins_x, uint32_t rank, uint32_t world) {
  return rank < bins_x ? (bins_x - 1 - rank) / world + 1 : 0u;
}
// first owned bin column >= bx0 and number of owned columns in [bx0, bx1]
__host__ __device__ inline void owned_span(uint32_t bx0, uint32_t bx1, uint32_t rank, uint32_t world, uint32_t &first,
                                           uint32_t &ncols) {
  first = bx0 + (rank + world - bx0 % world) % world;
  ncols = first <= bx1 ? (bx1 - first) / world + 1 : 0u;
}
// owned TILE columns of a rank: 4 per owned bin column, except that the last bin column may hold fewer tiles
__host__ __device__ inline uint32_t owned_tile_cols(uint32_t tiles_x, uint32_t rank, uint32_t world) {
  const uint32_t full = tiles_x / kTilesPerBin, rem = tiles_x % kTilesPerBin;
  uint32_t n = kTilesPerBin * owned_cols(full, rank, world);
  if (rem && full % world == rank) n += rem;
  return n;
}
// packed index of owned tile (tx, ty) in rank's tile buffer (row-major over its own tile columns)
__host__ __device__ inline uint32_t owned_slot(uint32_t tx, uint32_t ty, uint32_t tiles_x, uint32_t rank, uint32_t world) {
  const uint32_t bx = tx / kTilesPerBin;
  return ty * owned_tile_cols(tiles_x, rank, world) + (bx - rank) / world * kTilesPerBin + (tx % kTilesPerBin);
}

}  // namespace gs

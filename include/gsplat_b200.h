/*
 * gsplat_b200.h — C ABI of the B200-native Gaussian-splat sort + raster path.
 *
 * Drop-in boundary for the two hot loops of quadjr/aframe-gaussian-splatting `index.js`
 * (v0.0.22 @ b50238f).  Every entry point names the reference interface it replaces
 * (file:line into the reference).  Plain pointers and sizes only: no torch / C++ types.
 *
 * Conventions
 *   - all matrices are 16 x f32, COLUMN-MAJOR (THREE.Matrix4.elements order), already in the
 *     "gs" convention the reference hands its worker / shader:
 *       proj      = getProjectionMatrix()  (index.js:456-466)  -> uniform gsProjectionMatrix
 *       modelview = getModelViewMatrix()   (index.js:467-487)  -> uniform gsModelViewMatrix
 *       view[4]   = row 2 of modelview     (index.js:442)
 *       cutout16  = inverse(cutout.matrixWorld) * object.matrixWorld (index.js:443-448), or NULL
 *   - frames are written in GL window orientation: row 0 is the BOTTOM row (what
 *     gl.readPixels returns for the reference's render target).
 *   - every function returns 0 on success or a negative gs_status; no exception crosses the ABI.
 *   - a context is single-owner (not thread-safe), one context per GPU (index.js runs one
 *     worker + one GL context per component).
 *   - there is NO CPU fallback: gs_create fails with GS_ERR_CUDA when no sm_100 device exists.
 */
#ifndef GSPLAT_B200_H
#define GSPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GS_API __attribute__((visibility("default")))
#else
#define GS_API
#endif

typedef struct gs_context gs_context;

typedef enum gs_status {
  GS_OK = 0,
  GS_ERR_INVALID = -1,  /* bad argument                                                     */
  GS_ERR_CUDA = -2,     /* CUDA runtime error / no usable device (see gs_last_error)         */
  GS_ERR_OOM = -3,      /* device allocation failed                                          */
  GS_ERR_CAPACITY = -4, /* more splats than 2^31-1 (the reference silently truncates at
                           MAX_TEXTURE_SIZE^2, index.js:31-36,329-335; we report instead)    */
  GS_ERR_EMPTY = -5     /* sort/render before any push (the reference replies [0],
                           index.js:588-590, quirk Q7 - not reproduced)                      */
} gs_status;

enum { GS_FORMAT_RGBA8 = 0, GS_FORMAT_RGBA32F = 1 };

/* gs_render flags */
enum {
  GS_RENDER_OUT_DEVICE = 1u << 0, /* out_rgba is a device pointer on the context's GPU        */
  GS_RENDER_REUSE_SORT = 1u << 1, /* reuse the draw order of the previous gs_sort/gs_render
                                     (reference behaviour when sortReady is false,
                                     index.js:206,439-440: the draw uses a stale order).  Frames
                                     that sort >= 16 M splats are rendered front to back in depth
                                     slabs and leave no complete order behind: after such a frame
                                     the flag is ignored (the frame sorts) unless gs_sort ran since */
  GS_RENDER_OUT_TILED = 1u << 2,  /* multi-GPU: write only the tiles this rank owns, packed as
                                     16x16 RGBA blocks in owned-tile order (see gs_set_shard) */
  GS_RENDER_OUT_PEER = 1u << 3,   /* multi-GPU, fused raster + exchange: every finished tile is stored
                                     straight into ALL ranks' frames over NVLink peer memory (see
                                     gs_peer_export / gs_peer_import); no collective, no un-tiling */
  GS_RENDER_STATS = 1u << 4,      /* also fill the gs_stats fields marked (STATS): exact count of 16x16 tile
                                     instances and pixel-splat pair counters (a diagnostic frame: the raster
                                     keeps culling closed tiles' lists, so it is slower than a plain frame)  */
  GS_RENDER_DEPTH_DEVICE = 1u << 5 /* gs_render_params.depth_in is a device pointer (default: host memory)  */
};

/* Per-frame counters (SURVEY.md 8d symbols) and device timings of the last gs_sort/gs_render */
typedef struct gs_stats {
  uint32_t n_splats;       /* N   resident splats                                            */
  uint32_t n_sorted;       /* V   splats passing the worker filter (index.js:548)            */
  uint32_t n_dropped;      /*     sorted entries whose 16-bit key fell outside [0,65535] (Q5)*/
  uint32_t n_visible;      /* V2  entries also passing the shader clip-cull (index.js:110-115)*/
  uint64_t n_instances;    /*     16x16 tile candidates (bounding rectangles of the footprints)*/
  uint32_t n_tiles;        /* T   tiles in the frame                                         */
  uint32_t width, height;  /*     frame size (P = width*height)                              */
  double min_depth, max_depth; /* fp64 depth range of the sorted set (index.js:552-553)      */
  float ms_sort;           /* depth/cull + histogram + two radix passes                      */
  float ms_project;        /* per-splat projection + tile rect (runs beside the sort's radix
                              passes on a second stream: overlaps ms_sort)                    */
  float ms_bin;            /* instance emission + two tile-radix passes                      */
  float ms_raster;         /* tile raster + composite                                        */
  float ms_total;          /* first kernel to last kernel of this frame on the device; with several
                              frames in flight it includes waiting behind the previous raster  */
  uint32_t kernel_launches;/* kernels launched by the call                                   */
  uint32_t n_instances_kept;/*    BIN instances kept: screen bins (gs_bin_size(), 96 px) really meeting the r<=2 footprint.
                                  (n_instances counts the bounding-rectangle candidates.)  Splats are binned
                                  to bins; each 16x16 tile culls its bin's list in the raster.               */
  uint64_t n_tile_instances;/* D  (STATS) 16x16 tiles meeting the footprint, summed over the drawn splats     */
  uint64_t n_records_streamed;/*  (STATS) bin records the raster CTAs pulled through shared memory            */
  uint64_t n_pair_tests;    /*    (STATS) pixel-splat pairs evaluated by live pixels                          */
  uint64_t n_pair_hits;     /*    (STATS) pairs that passed r^2 <= 4 (and the depth test) and were blended   */
  uint32_t n_slabs;         /*    front-to-back slab path (large scenes): depth slabs scheduled; 0 = one-pass frame */
  uint32_t n_slabs_run;     /*    ... slabs that still found an open bin (the others launch and find nothing to do) */
  uint64_t n_slab_entries;  /*    ... draw-order entries of the slabs that ran: compacted, sorted and projected      */
} gs_stats;

/* ---- lifetime ---------------------------------------------------------------------------- */

/* Replaces: `new Worker(...)` + `initGL` texture allocation (index.js:25-46, 229-236). */
GS_API int gs_create(int device_ordinal, gs_context **out_ctx);
GS_API int gs_destroy(gs_context *ctx);
/* Last error text of this context (or of the failed gs_create when ctx == NULL). */
GS_API const char *gs_last_error(const gs_context *ctx);
GS_API const char *gs_version(void);
/* Edge, in pixels, of the square screen bins splats are binned to (a multiple of the 16-pixel raster tile; 64 unless the
 * library was built with another GS_BIN_TILES).  Multi-GPU tile ownership is by bin column (gs_set_shard). */
GS_API uint32_t gs_bin_size(void);

/* ---- seam 1: the worker message protocol (index.js:572-598) ------------------------------ */

/* {method:"clear"} (index.js:236,573-575): drop all resident splats. */
GS_API int gs_clear(gs_context *ctx);

/*
 * pushDataBuffer + {method:"push"} (index.js:328-437, 576-586): append n raw 32-byte .splat
 * rows (f32 pos[3], f32 scale[3], u8 rgba[4], u8 rot[4] stored w,x,y,z).  The load-time pack
 * (index.js:343-402, fp64, incl. the parseInt quirk) runs on the device.  rows32 is host memory.
 */
GS_API int gs_push_splats(gs_context *ctx, const void *rows32, uint32_t n);
/*
 * Progressive loading (index.js:259-298: rows are pushed as they arrive while the scene is already being drawn):
 * gs_push_splats / gs_push_packed do NOT wait for frames in flight.  A frame draws the splats that were resident when
 * it was submitted; the pushed rows are staged through page-locked buffers and packed on a separate stream behind it,
 * and the next submitted frame sees them.  rows32 is fully consumed when the call returns.  The only push that waits
 * for the pipeline is one that outgrows the table's capacity (geometric growth) - never after gs_reserve.
 *
 * gs_reserve: size the resident table for n_total splats up front, what initGL(numVertexes) does with the
 * Content-Length (index.js:248-251, 26-46).
 */
GS_API int gs_reserve(gs_context *ctx, uint32_t n_total);

/*
 * Append n already-packed splats: the two data-texture records the reference uploads
 * (centerAndScaleData float4, covAndColorData uint4, index.js:40-46,378-394) and the worker's
 * matrices[15] (max(scale)*alpha/255, index.js:397).  Host pointers.
 */
GS_API int gs_push_packed(gs_context *ctx, const float *center_scale4, const uint32_t *cov_color4,
                          const float *size_alpha, uint32_t n);

GS_API int gs_num_splats(const gs_context *ctx, uint32_t *out_n);

/* Read back the packed records of splats [first, first+n) (testing the device-side pack). */
GS_API int gs_read_packed(gs_context *ctx, uint32_t first, uint32_t n, float *center_scale4, uint32_t *cov_color4,
                          float *size_alpha);

/*
 * {method:"sort", view, cutout} -> {sortedIndexes} (index.js:449-453, 507-570, 587-596).
 * out_idx (host, capacity gs_num_splats) receives the surviving splat indices back-to-front,
 * bit-identical to the reference's Uint32Array (16-bit bucket order, ties by index; tail zeros
 * of quirk Q5 included); *out_count = its length.  out_idx may be NULL to sort on the device
 * only (the order stays resident for GS_RENDER_REUSE_SORT).
 */
GS_API int gs_sort(gs_context *ctx, const float view[4], const float *cutout16_or_null, uint32_t *out_idx,
                   uint32_t *out_count);

/* ---- seam 2: the draw (index.js:68-195 uniforms + shaders + blend state) ------------------ */

typedef struct gs_render_params {
  float proj[16];      /* gsProjectionMatrix (index.js:74,186)                               */
  float modelview[16]; /* gsModelViewMatrix  (index.js:75,187)                               */
  uint32_t width;      /* viewport.z (index.js:192)                                          */
  uint32_t height;     /* viewport.w (index.js:193)                                          */
  float focal;         /* (viewport.w/2)*abs(proj[5]) (index.js:191); <=0 -> computed so     */
  float bg_rgba[4];    /* clear colour the blend starts from (A-Frame default 0,0,0,0)       */
  int32_t has_cutout;  /* non-zero: cutout16 is valid (cutoutEntity, index.js:4,19-21)       */
  float cutout16[16];
  int32_t out_format;  /* GS_FORMAT_RGBA8 | GS_FORMAT_RGBA32F                                */
  uint32_t flags;      /* GS_RENDER_*                                                        */
  const float *depth_in; /* optional depth buffer of the geometry already drawn (index.js:179-180:
                          depthTest true, depthWrite false, three.js default LessEqualDepth): width*height
                          f32 WINDOW-space depths in [0,1], row 0 = bottom.  A fragment is kept iff
                          z/w*0.5+0.5 <= depth_in[pixel]; nothing is written back.  NULL = no depth test.
                          Host memory unless GS_RENDER_DEPTH_DEVICE; must stay valid until gs_wait.     */
} gs_render_params;

/*
 * One frame: tick() sort request + the instanced draw (index.js:438-455 + 184-207 + shaders
 * 77-176 + blend 177-181), synchronously: sort and draw use the same camera.
 * out_rgba: width*height*4 elements (u8 or f32), row 0 = bottom; host memory unless
 * GS_RENDER_OUT_DEVICE.  stats may be NULL.
 */
GS_API int gs_render(gs_context *ctx, const gs_render_params *params, void *out_rgba, gs_stats *stats);

/*
 * Pipelined form of gs_render (= gs_render_async + gs_wait).  A frame is three stages on three internal streams,
 * each one CUDA graph: A depth sort + projection, B tile binning, C raster; its counters and (for a host out_rgba)
 * its RGBA frame are then copied to the host on a fourth stream.  gs_render_async enqueues all of that and returns
 * a ticket at once; gs_wait blocks until that frame is in out_rgba.  FOUR frames may be outstanding (slot =
 * ticket % 4): frame k is rasterised while frame k+1 is binned and frame k+2 sorted, and frame k-1 crosses PCIe
 * (the reference likewise overlaps its worker sort with drawing, index.js:206,439-440) - a caller that receives
 * frames in host memory should keep four tickets open, so that collecting frame k-1's copy never delays the
 * submission of frame k+2.  A FIFTH gs_render_async first waits for the oldest frame (GS_RENDER_OUT_PEER frames:
 * a fourth, the shared frame ring has three entries).
 * Buffer lifetime: out_rgba (and any device buffer passed with GS_RENDER_OUT_DEVICE / _TILED) must stay valid and
 * untouched until gs_wait of that ticket returns, i.e. across up to four outstanding frames; use page-locked
 * memory (gs_host_alloc) for a truly asynchronous copy.
 * gs_wait(ticket) on a ticket that was already retired (by an earlier gs_wait, or implicitly when its slot was
 * reused or the pipeline was drained by gs_clear / gs_sort / a growing push) returns GS_OK with the stats of the
 * MOST RECENTLY completed frame, not necessarily that ticket's: the frame itself is already in out_rgba.
 */
GS_API int gs_render_async(gs_context *ctx, const gs_render_params *params, void *out_rgba, uint64_t *out_ticket);
GS_API int gs_wait(gs_context *ctx, uint64_t ticket, gs_stats *stats);

/*
 * WebXR / stereo (index.js:13-15 xrPixelRatio, 184-195): the scene's one sort request per frame comes from the HEAD
 * camera (tick(), index.js:438-455: `view` = row 2 of its gsModelViewMatrix, plus the cutout), while the mesh is drawn
 * once per EYE with that eye's matrices and viewport (material.onBeforeRender runs per eye camera).  gs_render_stereo =
 * one gs_sort + two draws with that order; eyes[e].has_cutout / cutout16 are ignored (the cutout acts in the sort).
 * stats2_or_null, when given, receives the two eyes' stats.
 */
GS_API int gs_render_stereo(gs_context *ctx, const float view[4], const float *cutout16_or_null,
                            const gs_render_params eyes[2], void *const out_rgba[2], gs_stats *stats2_or_null);

/* Per-splat projected record of the last gs_render (testing the vertex-shader restatement):
 * 8 floats per resident splat {cx, cy, a1x, a1y, a2x, a2y, rgba8-as-bits, tile-rect-as-bits};
 * rect == 0xFFFFFFFF marks a splat that was not projected/visible. */
GS_API int gs_read_projected(gs_context *ctx, uint32_t first, uint32_t n, float *out8);

GS_API int gs_get_stats(const gs_context *ctx, gs_stats *out);

/* ---- multi-GPU (new capability): screen-tile ownership ---------------------- */

/*
 * Shard the FRAME, not the splat table: rank r of `world` rasters the 16x16 tiles t with
 * tx % world == r (interleaved 16-pixel tile columns).  Every rank holds the full splat table (32 B/splat) and computes the
 * same global draw order, so each pixel is composited on exactly one GPU in exactly the
 * reference order.  The only exchange is an all-gather of finished RGBA tiles.
 */
GS_API int gs_set_shard(gs_context *ctx, uint32_t rank, uint32_t world);
/* Number of tiles rank `rank` owns for a width x height frame (all ranks pad to the max). */
GS_API uint32_t gs_owned_tiles(uint32_t width, uint32_t height, uint32_t rank, uint32_t world);
/* Scatter `world` gathered tiled buffers (each tiles_per_rank*256 pixels) into a row-major frame.
 * gathered / out_frame are device pointers.  Stream-ordered on gs_stream(ctx); call gs_synchronize to wait. */
GS_API int gs_assemble_tiles(gs_context *ctx, const void *gathered, uint32_t tiles_per_rank, uint32_t world,
                             uint32_t width, uint32_t height, int32_t format, void *out_frame);

/*
 * Fused raster + exchange (one process per GPU, same node).  Each rank calls gs_peer_export, the 64-byte handles
 * are exchanged by the host program (any transport; bench.py uses torch.distributed), then every rank calls
 * gs_peer_import with all `world` handles in rank order.  A gs_render_async with GS_RENDER_OUT_PEER then leaves the
 * complete frame in this rank's shared ring (gs_peer_frame), and also copies it to out_rgba when that is host
 * memory.  Flow control: a frame slot is rewritten only after every rank's gs_wait released its previous frame.
 */
GS_API int gs_peer_export(gs_context *ctx, size_t frame_bytes, void *ipc_handle_out64);
GS_API int gs_peer_import(gs_context *ctx, uint32_t rank, uint32_t world, const void *ipc_handles_world_x64);
/* device pointer of the assembled frame of `ticket` inside this rank's shared ring (valid until the third
 * following gs_render_async) */
GS_API int gs_peer_frame(gs_context *ctx, uint64_t ticket, void **out_dev_ptr);

/* Device-memory helpers so a host language without a CUDA binding can keep frames on the GPU. */
GS_API int gs_device_alloc(gs_context *ctx, size_t bytes, void **out_dev_ptr);
GS_API int gs_device_free(gs_context *ctx, void *dev_ptr);
/* Page-locked host memory (so frame read-back runs at PCIe speed and can be asynchronous). */
GS_API int gs_host_alloc(gs_context *ctx, size_t bytes, void **out_host_ptr);
GS_API int gs_host_free(gs_context *ctx, void *host_ptr);
GS_API int gs_memcpy_d2h(gs_context *ctx, void *dst_host, const void *src_dev, size_t bytes);
/* The CUDA stream (cudaStream_t) on which frames COMPLETE (the raster stream): work enqueued there after
 * gs_render_async (collectives, copies, gs_assemble_tiles) is ordered after that frame.  Sort + binning of the
 * next frame run on an internal second stream underneath the raster. */
GS_API void *gs_stream(gs_context *ctx);
GS_API int gs_synchronize(gs_context *ctx);

#ifdef __cplusplus
}
#endif
#endif /* GSPLAT_B200_H */

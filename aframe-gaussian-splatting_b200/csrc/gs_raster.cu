// gs_raster.cu — one CTA per 16x16 tile: the reference's fragment shader (index.js:170-175) and its
// blend state (index.js:177-181), composited front-to-back with transmittance.
//
// The tile's instance range [tile_start[t], tile_start[t+1]) holds 32 B projected records already laid
// out contiguously in back-to-front draw order, so the kernel pulls them with 1-D TMA bulk copies
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) through a 4-stage shared-memory ring and
// walks every chunk from its end (nearest splat) to its start.
//
//   back-to-front (reference):  C <- c*a + C*(1-a),  A <- a + A*(1-a)      (index.js:177-178)
//   front-to-back (here):       C  = sum_i c_i a_i T_i + bg*T_end,  A = 1 - T_end + bg.a*T_end,
//                               T_i = prod_{j nearer than i} (1 - a_j)      
// The two are algebraically identical; a tile stops early once every pixel has T < 3e-4, which bounds the
// dropped contribution by 3e-4 per channel (the parity tolerance is 1e-3).
#include "gs_common.cuh"

namespace gs {

constexpr int kChunk = 128;   // records per TMA bulk copy (4 KB)
constexpr int kStages = 4;    // ring depth
constexpr float kTStop = 3e-4f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ uint32_t to_u8(float v) {
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  return (uint32_t)(v * 255.0f + 0.5f);
}

__global__ void __launch_bounds__(256) k_raster(const float4 *__restrict__ inst_rec,
                                                const uint2 *__restrict__ tile_range,
                                                const FrameParams *__restrict__ fp) {
  const RenderConsts &rc = fp->rc;
  void *out = fp->out;
  __shared__ __align__(128) float4 s_rec[kStages][kChunk * 2];
  __shared__ __align__(16) float4 s_col[kChunk];
  __shared__ __align__(8) uint64_t s_full[kStages];

  const uint32_t tile = blockIdx.x;
  const uint32_t tx = tile % rc.tiles_x, ty = tile / rc.tiles_x;
  if (rc.shard_world > 1 && (tx % rc.shard_world) != rc.shard_rank) return;

  const uint32_t tid = threadIdx.x;
  // a warp owns a compact 8x4 pixel block (fewer splats straddle it than a 16x2 strip): tid = [ty2 tx1 | y2 x3]
  const uint32_t lx = ((tid >> 5) & 1u) * 8u + (tid & 7u), ly = (tid >> 6) * 4u + ((tid >> 3) & 3u);
  const uint32_t x = tx * kTile + lx, y = ty * kTile + ly;
  const bool inside = (x < rc.width) && (y < rc.height);
  const float fx = (float)x + 0.5f, fy = (float)y + 0.5f;  // pixel centre, GL window coordinates

  const uint2 range = tile_range[tile];
  const uint32_t start = range.x, end = range.y;
  const uint32_t count = end - start;
  const uint32_t n_chunks = (count + kChunk - 1) / kChunk;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // chunk k covers records [lo_k, hi_k) with hi_k = end - k*kChunk (nearest first)
  auto issue = [&](uint32_t k) {
    const uint32_t hi = end - k * kChunk;
    const uint32_t lo = (hi - start > (uint32_t)kChunk) ? hi - kChunk : start;
    const uint32_t bytes = (hi - lo) * 32u;
    uint64_t *bar = &s_full[k % kStages];
    mbar_expect_tx(bar, bytes);
    bulk_g2s(&s_rec[k % kStages][0], inst_rec + 2 * (size_t)lo, bytes, bar);
  };
  if (tid == 0) {
    for (uint32_t k = 0; k < (uint32_t)kStages && k < n_chunks; ++k) issue(k);
  }

  float T = 1.0f, Cr = 0.0f, Cg = 0.0f, Cb = 0.0f;
  uint32_t k = 0;
  for (; k < n_chunks; ++k) {
    const uint32_t stage = k % kStages;
    mbar_wait(&s_full[stage], (k / kStages) & 1u);
    const uint32_t hi = end - k * kChunk;
    const uint32_t lo = (hi - start > (uint32_t)kChunk) ? hi - kChunk : start;
    const uint32_t m = hi - lo;
    const float4 *rec = &s_rec[stage][0];
    // colour bytes -> float once per record (index.js:152-157), not once per pixel
    if (tid < m) {
      const float4 r1 = rec[2 * tid + 1];
      const uint32_t bits = __float_as_uint(r1.z);
      s_col[tid] = make_float4(__fdiv_rn((float)(bits & 255u), 255.0f), __fdiv_rn((float)((bits >> 8) & 255u), 255.0f),
                               __fdiv_rn((float)((bits >> 16) & 255u), 255.0f), r1.w);
    }
    __syncthreads();
    if (T >= kTStop) {
      for (int j = (int)m - 1; j >= 0; --j) {
        const float4 r0 = rec[2 * j];      // cx, cy, a1x, a1y
        const float4 r1 = rec[2 * j + 1];  // a2x, a2y, rgba bits, alpha
        const float dx = __fsub_rn(fx, r0.x), dy = __fsub_rn(fy, r0.y);
        // vPosition = (px, py): same op order as the oracle (orc band_worker)
        const float px = __fmaf_rn(dy, r1.y, __fmul_rn(dx, r1.x));
        const float py = __fmaf_rn(dy, r0.w, __fmul_rn(dx, r0.z));
        const float r2 = __fmaf_rn(py, py, __fmul_rn(px, px));
        if (r2 <= 4.0f) {  // index.js:171-172: A = -r2; discard if A < -4
          const float4 col = s_col[j];
          const float alpha = __fmul_rn(__expf(-r2), col.w);  // index.js:173
          const float w = __fmul_rn(alpha, T);
          Cr = __fmaf_rn(col.x, w, Cr);
          Cg = __fmaf_rn(col.y, w, Cg);
          Cb = __fmaf_rn(col.z, w, Cb);
          T = __fsub_rn(T, w);
        }
      }
    }
    const int alive = __syncthreads_or(inside && (T >= kTStop));
    if (!alive) break;
    if (tid == 0 && k + kStages < n_chunks) issue(k + kStages);
  }
  // early exit: bulk copies already in flight must land before the CTA (and its shared memory) retires
  if (tid == 0 && k < n_chunks) {
    for (uint32_t kk = k + 1; kk < n_chunks && kk < k + kStages; ++kk) mbar_wait(&s_full[kk % kStages], (kk / kStages) & 1u);
  }

  // composite over the clear colour
  const float oR = __fmaf_rn(rc.bg[0], T, Cr), oG = __fmaf_rn(rc.bg[1], T, Cg), oB = __fmaf_rn(rc.bg[2], T, Cb);
  const float oA = __fmaf_rn(rc.bg[3], T, 1.0f - T);
  size_t pix;
  bool write;
  if (rc.out_tiled) {
    const uint32_t slot =
        (rc.shard_world > 1) ? owned_slot(tx, ty, rc.tiles_x, rc.shard_rank, rc.shard_world) : tile;
    pix = (size_t)slot * 256 + ly * 16 + lx;
    write = true;
  } else {
    pix = (size_t)y * rc.width + x;
    write = inside;
  }
  if (write) {
    if (fp->n_peer) {
      // fused exchange: the finished pixel goes straight into every rank's frame over NVLink peer stores, so the
      // transfer overlaps the raster tile by tile and no collective / un-tiling pass follows
      const uint32_t np = fp->n_peer;
      if (rc.out_format == GS_FORMAT_RGBA8) {
        const uint32_t v = to_u8(oR) | (to_u8(oG) << 8) | (to_u8(oB) << 16) | (to_u8(oA) << 24);
        for (uint32_t r = 0; r < np; ++r) ((uint32_t *)fp->peer_out[r])[pix] = v;
      } else {
        const float4 v = make_float4(oR, oG, oB, oA);
        for (uint32_t r = 0; r < np; ++r) ((float4 *)fp->peer_out[r])[pix] = v;
      }
    } else if (rc.out_format == GS_FORMAT_RGBA8) {
      const uint32_t v = inside ? (to_u8(oR) | (to_u8(oG) << 8) | (to_u8(oB) << 16) | (to_u8(oA) << 24)) : 0u;
      ((uint32_t *)out)[pix] = v;
    } else {
      ((float4 *)out)[pix] = inside ? make_float4(oR, oG, oB, oA) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ---- fused-exchange flow control: two monotonic flag rows per frame slot, written over peer memory ----
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *p >= need; gives up after ~2 s (a lost peer must not hang the GPU) and reports through ctr
__device__ __forceinline__ void wait_flag(const unsigned long long *p, unsigned long long need, FrameCounters *ctr) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < need) {
    if (clock64() - t0 > 4000000000ll) { atomicExch(&ctr->peer_timeout, 1u); break; }
    __nanosleep(200);
  }
}

// before the raster may overwrite this slot of every rank's frame ring: all ranks have released its previous frame
__global__ void k_peer_acquire(const FrameParams *__restrict__ fp, FrameCounters *ctr) {
  if (threadIdx.x < fp->n_peer && fp->peer_need) wait_flag(fp->local_released + threadIdx.x, fp->peer_need, ctr);
}

// after the raster: tell every rank our tiles of this frame have landed, then wait for everybody else's
__global__ void k_peer_signal_wait(const FrameParams *__restrict__ fp, FrameCounters *ctr) {
  const uint32_t r = threadIdx.x;
  if (r >= fp->n_peer) return;
  __threadfence_system();
  st_release_sys(fp->peer_done[r] + fp->peer_rank, fp->peer_seq);
  wait_flag(fp->local_done + r, fp->peer_seq, ctr);
}

// when the host has consumed a frame (gs_wait): every rank may overwrite our copy of that slot
__global__ void k_peer_release(PeerRows rows, uint32_t world, uint32_t rank, unsigned long long seq) {
  if (threadIdx.x < world) st_release_sys(rows.p[threadIdx.x] + rank, seq);
}

// scatter `world` gathered tiled buffers back into a row-major frame (one thread per pixel)
__global__ void __launch_bounds__(256) k_assemble(const void *__restrict__ gathered, uint32_t tiles_per_rank,
                                                  uint32_t world, uint32_t width, uint32_t height, int32_t format,
                                                  void *__restrict__ out) {
  const uint32_t tiles_x = (width + kTile - 1) / kTile;
  const uint32_t tile = blockIdx.x;
  const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
  const uint32_t tid = threadIdx.x;
  const uint32_t x = tx * kTile + (tid & 15u), y = ty * kTile + (tid >> 4);
  if (x >= width || y >= height) return;
  const uint32_t rank = tx % world;
  const uint32_t slot = (world > 1) ? owned_slot(tx, ty, tiles_x, rank, world) : tile;
  const size_t src = ((size_t)rank * tiles_per_rank + slot) * 256 + tid;
  const size_t dst = (size_t)y * width + x;
  if (format == GS_FORMAT_RGBA8) ((uint32_t *)out)[dst] = ((const uint32_t *)gathered)[src];
  else ((float4 *)out)[dst] = ((const float4 *)gathered)[src];
}

void launch_raster(gs_context *c, const FrameParams *fp, uint32_t n_tiles, const FrameBufs &b, cudaStream_t st) {
  k_raster<<<n_tiles, 256, 0, st>>>(b.inst_rec, b.tile_range, fp);
}

void launch_peer_acquire(gs_context *c, const FrameParams *fp, FrameCounters *ctr, cudaStream_t st) {
  k_peer_acquire<<<1, 32, 0, st>>>(fp, ctr);
}
void launch_peer_signal_wait(gs_context *c, const FrameParams *fp, FrameCounters *ctr, cudaStream_t st) {
  k_peer_signal_wait<<<1, 32, 0, st>>>(fp, ctr);
}
void launch_peer_release(gs_context *c, const PeerRows &rows, uint32_t world, uint32_t rank, unsigned long long seq,
                         cudaStream_t st) {
  k_peer_release<<<1, 32, 0, st>>>(rows, world, rank, seq);
}

void launch_assemble(gs_context *c, const void *gathered, uint32_t tiles_per_rank, uint32_t world, uint32_t width,
                     uint32_t height, int32_t format, void *out_frame) {
  const uint32_t tiles_x = (width + kTile - 1) / kTile, tiles_y = (height + kTile - 1) / kTile;
  k_assemble<<<tiles_x * tiles_y, 256, 0, c->rstream>>>(gathered, tiles_per_rank, world, width, height, format, out_frame);
}

uint32_t owned_tiles_host(uint32_t width, uint32_t height, uint32_t rank, uint32_t world) {
  const uint32_t tiles_x = (width + kTile - 1) / kTile, tiles_y = (height + kTile - 1) / kTile;
  if (world <= 1) return tiles_x * tiles_y;
  uint32_t n = 0;
  n = tiles_y * owned_cols(tiles_x, rank, world);
  return n;
}

}  // namespace gs

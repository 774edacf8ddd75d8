// gs_project.cu — per-splat projection (the reference's vertex shader, index.js:101-164) and the
// ordered emission of 16x16 tile instances.
//
//   k_project  : fp32 restatement of the vertex shader, op for op (no FMA contraction), producing a
//                32 B projected record per splat + its packed tile rectangle.
//   k_count    : instances per 1024-entry slice of the draw order (== reference sortedIndexes) + frame total D.
//   k_emit     : prefix of those totals + in-slice scan -> writes (tile, splat) instances in draw order, so that a
//                STABLE sort by tile id alone reproduces the reference's back-to-front order inside every tile.
//   k_tile_scan: exclusive scan of the per-tile instance counts -> tile ranges for the raster.
#include "gs_common.cuh"

namespace gs {

// unpackInt16 (index.js:92-99)
__device__ __forceinline__ void unpack_int16(uint32_t value, float &lo, float &hi) {
  const int32_t v = (int32_t)value;
  const int32_t v0 = v >> 16;
  int32_t v1 = v & 0xFFFF;
  if (v & 0x8000) v1 |= (int32_t)0xFFFF0000;
  lo = (float)v1;
  hi = (float)v0;
}

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define DIV(a, b) __fdiv_rn((a), (b))
// a0*b0 + a1*b1 + a2*b2, left to right
#define DOT3(a0, b0, a1, b1, a2, b2) ADD(ADD(MUL(a0, b0), MUL(a1, b1)), MUL(a2, b2))

// ---------------------------------------------------------------------------------------------
// K2: vertex shader restatement.  One thread per resident splat, index order (coalesced 16 B + 16 B
// loads, 32 B + 4 B stores).  Splats rejected by the worker filter are skipped, except splat 0 which
// the reference may draw through the zero tail of quirk Q5.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_project(const float4 *__restrict__ cs, const uint4 *__restrict__ cc,
                                                 const float *__restrict__ depth, uint32_t n,
                                                 const FrameParams *__restrict__ fp, float4 *__restrict__ rec_out,
                                                 uint32_t *__restrict__ rect_out) {
  const RenderConsts &rc = fp->rc;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t rect = kNoRect;
    const bool sorted = (__ldg(depth + i) != GS_DEPTH_REJECT) || (i == 0);
    if (sorted) {
      const float4 c = __ldg(cs + i);
      const float *mv = rc.mv, *P = rc.proj;
      // index.js:106-108: camspace = MV * (center,1); pos2d = P * camspace  (sum x,y,z,w left to right)
      float cam[4], p[4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        cam[r] = ADD(ADD(ADD(MUL(mv[r], c.x), MUL(mv[4 + r], c.y)), MUL(mv[8 + r], c.z)), MUL(mv[12 + r], 1.0f));
#pragma unroll
      for (int r = 0; r < 4; ++r)
        p[r] = ADD(ADD(ADD(MUL(P[r], cam[0]), MUL(P[4 + r], cam[1])), MUL(P[8 + r], cam[2])), MUL(P[12 + r], cam[3]));
      // index.js:110-115
      const float bounds = MUL(1.2f, p[3]);
      const bool culled = (p[2] < -p[3]) || (p[0] < -bounds) || (p[0] > bounds) || (p[1] < -bounds) || (p[1] > bounds);
      if (!culled) {
        const uint4 q = __ldg(cc + i);
        // index.js:117-125
        float c00, c01, c02, c11, c12, c22;
        unpack_int16(q.x, c00, c01);
        unpack_int16(q.y, c02, c11);
        unpack_int16(q.z, c12, c22);
        const float s = c.w;
        c00 = MUL(c00, s); c01 = MUL(c01, s); c02 = MUL(c02, s);
        c11 = MUL(c11, s); c12 = MUL(c12, s); c22 = MUL(c22, s);
        const float V[3][3] = {{c00, c01, c02}, {c01, c11, c12}, {c02, c12, c22}};
        // index.js:127-131 (GLSL mat3 constructor is column-major)
        const float zz = MUL(cam[2], cam[2]);
        float J[3][3];
        J[0][0] = DIV(rc.focal, cam[2]); J[1][0] = 0.0f; J[2][0] = DIV(-MUL(rc.focal, cam[0]), zz);
        J[0][1] = 0.0f; J[1][1] = DIV(-rc.focal, cam[2]); J[2][1] = DIV(MUL(rc.focal, cam[1]), zz);
        J[0][2] = 0.0f; J[1][2] = 0.0f; J[2][2] = 0.0f;
        // index.js:133-135
        float T[3][3], U[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int k = 0; k < 3; ++k)
            T[r][k] = DOT3(mv[r * 4 + 0], J[0][k], mv[r * 4 + 1], J[1][k], mv[r * 4 + 2], J[2][k]);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int k = 0; k < 3; ++k) U[r][k] = DOT3(T[0][r], V[0][k], T[1][r], V[1][k], T[2][r], V[2][k]);
        const float cov00 = DOT3(U[0][0], T[0][0], U[0][1], T[1][0], U[0][2], T[2][0]);
        const float cov10 = DOT3(U[1][0], T[0][0], U[1][1], T[1][0], U[1][2], T[2][0]);
        const float cov11 = DOT3(U[1][0], T[0][1], U[1][1], T[1][1], U[1][2], T[2][1]);
        // index.js:137-149
        const float vcx = DIV(p[0], p[3]), vcy = DIV(p[1], p[3]);
        const float diagonal1 = ADD(cov00, 0.3f);
        const float offDiagonal = cov10;
        const float diagonal2 = ADD(cov11, 0.3f);
        const float mid = MUL(0.5f, ADD(diagonal1, diagonal2));
        const float hd = DIV(SUB(diagonal1, diagonal2), 2.0f);
        const float radius = __fsqrt_rn(ADD(MUL(hd, hd), MUL(offDiagonal, offDiagonal)));
        const float lambda1 = ADD(mid, radius);
        const float l2raw = SUB(mid, radius);
        const float lambda2 = (l2raw < 0.1f) ? 0.1f : l2raw;
        const float dvx0 = offDiagonal, dvy0 = SUB(lambda1, diagonal1);
        const float dlen = __fsqrt_rn(ADD(MUL(dvx0, dvx0), MUL(dvy0, dvy0)));
        const float dvx = DIV(dvx0, dlen), dvy = DIV(dvy0, dlen);
        const float s1 = __fsqrt_rn(MUL(2.0f, lambda1)), s2 = __fsqrt_rn(MUL(2.0f, lambda2));
        const float l1 = (1024.0f < s1) ? 1024.0f : s1;
        const float l2 = (1024.0f < s2) ? 1024.0f : s2;
        const float v1x = MUL(l1, dvx), v1y = MUL(l1, dvy);
        const float v2x = MUL(l2, dvy), v2y = MUL(l2, -dvx);
        // index.js:160-163: the quad point q lands on window pixel c_px + q.x*v2 + q.y*v1
        const float zndc = DIV(p[2], p[3]);
        const float cx = MUL(ADD(MUL(vcx, 0.5f), 0.5f), rc.vw);
        const float cy = MUL(ADD(MUL(vcy, 0.5f), 0.5f), rc.vh);
        const float n1 = ADD(MUL(v1x, v1x), MUL(v1y, v1y));
        const float n2 = ADD(MUL(v2x, v2x), MUL(v2y, v2y));
        const float a1x = DIV(v1x, n1), a1y = DIV(v1y, n1);
        const float a2x = DIV(v2x, n2), a2y = DIV(v2y, n2);
        bool ok = (zndc <= 1.0f);  // GL clips the whole quad beyond the far plane (z/w > 1, w = 1)
        ok = ok && (a1x == a1x) && (a1y == a1y) && (a2x == a2x) && (a2y == a2y) && (cx == cx) && (cy == cy);
        if (ok) {
          // conservative pixel bounding box of the r<=2 disc image
          const float ex = 2.0f * sqrtf(v1x * v1x + v2x * v2x) + 0.01f;
          const float ey = 2.0f * sqrtf(v1y * v1y + v2y * v2y) + 0.01f;
          float fx0 = ceilf(cx - ex - 0.5f), fx1 = floorf(cx + ex - 0.5f);
          float fy0 = ceilf(cy - ey - 0.5f), fy1 = floorf(cy + ey - 0.5f);
          fx0 = fmaxf(fx0, 0.0f);
          fy0 = fmaxf(fy0, 0.0f);
          fx1 = fminf(fx1, (float)rc.width - 1.0f);
          fy1 = fminf(fy1, (float)rc.height - 1.0f);
          if (fx0 <= fx1 && fy0 <= fy1) {
            const uint32_t tx0 = (uint32_t)fx0 >> 4, tx1 = (uint32_t)fx1 >> 4;
            const uint32_t ty0 = (uint32_t)fy0 >> 4, ty1 = (uint32_t)fy1 >> 4;
            rect = tx0 | (tx1 << 8) | (ty0 << 16) | (ty1 << 24);
            // alpha as f32 = float(byte)/255.0 (index.js:156); rgb stay packed, converted in the raster
            const float alpha = DIV((float)(q.w >> 24), 255.0f);
            rec_out[2 * (size_t)i] = make_float4(cx, cy, a1x, a1y);
            rec_out[2 * (size_t)i + 1] = make_float4(a2x, a2y, __uint_as_float(q.w), alpha);
          }
        }
      }
    }
    rect_out[i] = rect;
  }
}

// candidate tiles of a packed rectangle that this rank owns (all of them on one GPU)
__device__ __forceinline__ uint32_t rect_count(uint32_t r, uint32_t rank, uint32_t world) {
  if (r == kNoRect) return 0u;
  const uint32_t h = (r >> 24) - ((r >> 16) & 255u) + 1u;
  uint32_t w = ((r >> 8) & 255u) - (r & 255u) + 1u;
  if (world > 1) {
    uint32_t first;
    owned_span(r & 255u, (r >> 8) & 255u, rank, world, first, w);
  }
  return w * h;
}

// ---------------------------------------------------------------------------------------------
// K3a: instances per emission tile (1024 consecutive draw-order entries) and the frame total D.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEmitThreads) k_count(const uint32_t *__restrict__ order,
                                                        const uint32_t *__restrict__ rect,
                                                        uint32_t *__restrict__ tile_total, FrameCounters *ctr,
                                                        const FrameParams *__restrict__ fp) {
  const uint32_t shard_rank = fp->rc.shard_rank, shard_world = fp->rc.shard_world;
  __shared__ uint32_t s_sum[kEmitThreads / 32], s_vis[kEmitThreads / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nv = ctr->n_valid;
  const uint32_t num_tiles = (nv + kEmitTile - 1) / kEmitTile;
  for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const uint32_t j0 = tile * kEmitTile + tid * kEmitItems;
    uint32_t sum = 0, vis = 0;
#pragma unroll
    for (int k = 0; k < kEmitItems; ++k) {
      const uint32_t j = j0 + k;
      if (j < nv) {
        const uint32_t r = __ldg(rect + __ldg(order + j));
        sum += rect_count(r, shard_rank, shard_world);
        vis += (r != kNoRect);
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      vis += __shfl_xor_sync(0xffffffffu, vis, o);
    }
    if (lane == 0) { s_sum[warp] = sum; s_vis[warp] = vis; }
    __syncthreads();
    if (tid == 0) {
      uint32_t t = 0, v = 0;
      for (int k = 0; k < kEmitThreads / 32; ++k) { t += s_sum[k]; v += s_vis[k]; }
      tile_total[tile] = t;
      if (t) atomicAdd(&ctr->n_inst, (unsigned long long)t);
      if (v) atomicAdd(&ctr->n_visible, v);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// K3b: ordered instance emission.  Instance (entry j, k-th tile of its rectangle) lands at
// position prefix(j) + k, so the instance array is in draw order whatever the execution order.
// One thread per instance (balanced expansion); every candidate tile of the bounding rectangle is
// tested exactly against the r<=2 footprint (closest point of the tile's pixel-centre box in the splat's
// (px,py) frame) and rejected tiles are written as kNoTile, which the T1 pass drops.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEmitThreads, 6) k_emit(const uint32_t *__restrict__ order,
                                                       const uint32_t *__restrict__ rect,
                                                       const float4 *__restrict__ proj_rec,
                                                       const FrameParams *__restrict__ fp, uint64_t cap_inst, const uint32_t *__restrict__ tile_total,
                                                       uint16_t *__restrict__ inst_tile, uint32_t *__restrict__ inst_idx,
                                                       uint32_t *__restrict__ tile_count, FrameCounters *ctr) {
  const RenderConsts &rc = fp->rc;
  __shared__ uint32_t s_off[kEmitTile];
  __shared__ uint32_t s_rect[kEmitTile];
  __shared__ uint32_t s_idx[kEmitTile];
  __shared__ uint32_t s_warp[kEmitThreads / 32];
  __shared__ unsigned long long s_red[kEmitThreads / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nv = ctr->n_valid;
  const uint32_t num_tiles = (nv + kEmitTile - 1) / kEmitTile;
  if (ctr->n_inst > cap_inst) {  // instance buffer too small: the host regrows it and re-runs the frame
    if (blockIdx.x == 0 && tid == 0) ctr->overflow = 1u;
    return;
  }
  unsigned long long base = 0;
  uint32_t summed_upto = 0;  // tile_total[0, summed_upto) is already in `base`
  for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    // ---- exclusive prefix of the tile totals (block reduction over the not-yet-summed range) ----
    unsigned long long part = 0;
    for (uint32_t t = summed_upto + tid; t < tile; t += kEmitThreads) part += tile_total[t];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[warp] = part;
    // ---- load the tile's entries ----
    const uint32_t j0 = tile * kEmitTile + tid * kEmitItems;
    uint32_t cnt[kEmitItems], sum = 0;
#pragma unroll
    for (int k = 0; k < kEmitItems; ++k) {
      const uint32_t j = j0 + k;
      uint32_t idx = 0, r = kNoRect;
      if (j < nv) {
        idx = __ldg(order + j);
        r = __ldg(rect + idx);
      }
      s_idx[tid * kEmitItems + k] = idx;
      s_rect[tid * kEmitItems + k] = r;
      cnt[k] = rect_count(r, rc.shard_rank, rc.shard_world);
      sum += cnt[k];
    }
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t k = 0; k < warp; ++k) wbase += s_warp[k];
    for (int k = 0; k < kEmitThreads / 32; ++k) base += s_red[k];
    summed_upto = tile;
    uint32_t run = wbase + incl - sum;
#pragma unroll
    for (int k = 0; k < kEmitItems; ++k) {
      s_off[tid * kEmitItems + k] = run;
      run += cnt[k];
    }
    __syncthreads();
    // ---- expansion: one thread per instance, strided over the slice's instance space (balanced whatever
    //      the rectangle sizes); the entry of instance e is found by binary search in the scanned offsets ----
    const uint32_t total = s_off[kEmitTile - 1] + rect_count(s_rect[kEmitTile - 1], rc.shard_rank, rc.shard_world);
    for (uint32_t e = tid; e < total; e += kEmitThreads) {
      uint32_t lo = 0, hi = kEmitTile;
#pragma unroll
      for (int it = 0; it < 10; ++it) {  // kEmitTile == 1024
        const uint32_t mid = (lo + hi) >> 1;
        if (s_off[mid] <= e) lo = mid; else hi = mid;
      }
      const uint32_t r = s_rect[lo];
      const uint32_t idx = s_idx[lo];
      const uint32_t k = e - s_off[lo];
      // owned columns of the rectangle: first, first + world, ... (all columns on one GPU)
      uint32_t tx0 = r & 255u, w = ((r >> 8) & 255u) - tx0 + 1u;
      const uint32_t ty0 = (r >> 16) & 255u;
      const uint32_t n_all = w * ((r >> 24) - ty0 + 1u);
      uint32_t step = 1u;
      if (rc.shard_world > 1) {
        owned_span(tx0, (r >> 8) & 255u, rc.shard_rank, rc.shard_world, tx0, w);
        step = rc.shard_world;
      }
      // k / w for k < 65536, w <= 256: float quotient of (k + 0.5) is never within rounding of an integer
      const uint32_t dy_t = (uint32_t)__fdividef((float)k + 0.5f, (float)w);
      const uint32_t tx = tx0 + (k - dy_t * w) * step, ty = ty0 + dy_t;
      bool keep = true;
      if (n_all > 1) {
        // footprint geometry: neighbouring instances share the splat, so these gathers mostly hit L1
        const float4 r0 = __ldg(proj_rec + 2 * (size_t)idx);                       // cx, cy, a1x, a1y
        const float2 r1 = __ldg((const float2 *)(proj_rec + 2 * (size_t)idx + 1));  // a2x, a2y
        // pixel-centre box of the tile, relative to the splat centre
        const float xa = (float)(tx * kTile) + 0.5f - r0.x, xb = xa + 15.0f;
        const float ya = (float)(ty * kTile) + 0.5f - r0.y, yb = ya + 15.0f;
        const bool in_x = (xa <= 0.0f) && (xb >= 0.0f), in_y = (ya <= 0.0f) && (yb >= 0.0f);
        if (!(in_x && in_y)) {
          float qmin = 3.0e38f;
          if (!in_x) {  // nearest vertical edge, minimise over y on it
            const float dx = (xa > 0.0f) ? xa : xb;
            const float px0 = dx * r1.x, py0 = dx * r0.z;
            float t = -__fdividef(px0 * r1.y + py0 * r0.w, r1.y * r1.y + r0.w * r0.w);
            t = fminf(fmaxf(t, ya), yb);
            const float px = px0 + t * r1.y, py = py0 + t * r0.w;
            qmin = px * px + py * py;
          }
          if (!in_y) {  // nearest horizontal edge, minimise over x on it
            const float dy = (ya > 0.0f) ? ya : yb;
            const float px0 = dy * r1.y, py0 = dy * r0.w;
            float t = -__fdividef(px0 * r1.x + py0 * r0.z, r1.x * r1.x + r0.z * r0.z);
            t = fminf(fmaxf(t, xa), xb);
            const float px = px0 + t * r1.x, py = py0 + t * r0.z;
            qmin = fminf(qmin, px * px + py * py);
          }
          keep = !(qmin > 4.02f);  // r^2 <= 4 with slack for fp32 rounding of the closest-point search
        }
      }
      uint32_t t = kNoTile;
      if (keep) {
        t = ty * rc.tiles_x + tx;
        atomicAdd(tile_count + t, 1u);
      }
      inst_tile[base + e] = (uint16_t)t;
      inst_idx[base + e] = idx;
    }
    __syncthreads();
  }
}

// exclusive scan of tile_count[T] -> tile_start[T+1]; one CTA
__global__ void __launch_bounds__(1024) k_tile_scan(const uint32_t *__restrict__ tile_count, uint32_t n_tiles,
                                                    uint32_t *__restrict__ tile_start) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t b = 0; b < n_tiles; b += 1024) {
    const uint32_t i = b + tid;
    const uint32_t v = (i < n_tiles) ? tile_count[i] : 0u;
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t k = 0; k < warp; ++k) wbase += s_warp[k];
    const uint32_t carry = s_carry;
    if (i < n_tiles) tile_start[i] = carry + wbase + incl - v;
    __syncthreads();
    if (tid == 1023) s_carry = carry + wbase + incl;
    __syncthreads();
  }
  if (tid == 0) tile_start[n_tiles] = s_carry;
}

void launch_project(gs_context *c, const FrameParams *fp) {
  uint64_t blocks = ((uint64_t)c->n + 255) / 256;
  const uint64_t cap = (uint64_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_project<<<(int)blocks, 256, 0, c->stream>>>(c->center_scale, c->cov_color, c->depth, c->n, fp, c->proj_rec, c->rect);
}

void launch_emit(gs_context *c, const FrameParams *fp, FrameCounters *ctr) {
  uint64_t tiles = ((uint64_t)c->n + kEmitTile - 1) / kEmitTile;
  const uint64_t cap = (uint64_t)c->sm_count * 8;
  if (tiles > cap) tiles = cap;
  if (tiles < 1) tiles = 1;
  k_count<<<(int)tiles, kEmitThreads, 0, c->stream>>>(c->order, c->rect, c->tile_total, ctr, fp);
  k_emit<<<(int)tiles, kEmitThreads, 0, c->stream>>>(c->order, c->rect, c->proj_rec, fp, c->cap_inst, c->tile_total,
                                                     c->inst_tile, c->inst_idx, c->tile_count, ctr);
}

void launch_tile_scan(gs_context *c, uint32_t n_tiles) {
  k_tile_scan<<<1, 1024, 0, c->stream>>>(c->tile_count, n_tiles, c->tile_start);
}

}  // namespace gs

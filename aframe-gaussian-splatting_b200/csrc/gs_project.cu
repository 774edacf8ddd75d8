// gs_project.cu — per-splat projection (the reference's vertex shader, index.js:101-164) and the
// ordered emission of bin instances (kBin x kBin-pixel bins = 6x6 raster tiles by default; "tile" below means bin).
//
//   k_project  : fp32 restatement of the vertex shader, op for op (no FMA contraction), producing a
//                32 B projected record per splat + its packed tile rectangle.
//   k_count    : per entry of the draw order (== reference sortedIndexes): instance offset inside its 256-entry slice;
//                per slice: total; last CTA: prefix over the slices + frame total D.
//                Sparse frames (fewer than half of the splats sorted): each chunk's survivors are compacted first.
//   k_emit_entries (default): one thread per draw-order entry writes its (bin, splat) instances at the entry's offset,
//                in draw order, so that a STABLE sort by bin id alone reproduces the reference's back-to-front order
//                inside every bin; rectangles of more than 8 bins are finished by the whole warp.
//   k_emit     (GS_EMIT=windows, round 1): one CTA per window of 2048 instance positions; also produces pass T1's
//                per-window digit histograms (table[digit][window]).
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
// One splat through the vertex shader: returns its packed bin rectangle (kNoRect when nothing is drawn) and stores the
// 32 B record at slot j.
__device__ __forceinline__ uint32_t project_one(const RenderConsts &rc, const float4 *__restrict__ cs,
                                                const uint4 *__restrict__ cc, uint32_t i, uint32_t j,
                                                float4 *__restrict__ rec_out) {
  uint32_t rect = kNoRect;
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
        // rectangle of BINS (at most 43 per axis for frames up to 4096 px at 96 px: never equals kNoRect)
        const uint32_t tx0 = (uint32_t)fx0 / (uint32_t)kBin, tx1 = (uint32_t)fx1 / (uint32_t)kBin;
        const uint32_t ty0 = (uint32_t)fy0 / (uint32_t)kBin, ty1 = (uint32_t)fy1 / (uint32_t)kBin;
        rect = tx0 | (tx1 << 8) | (ty0 << 16) | (ty1 << 24);
        // rgba stay packed (converted to float(byte)/255.0, index.js:152-157, once per record in the raster);
        // the last slot carries gl_Position.z/w (index.js:163) for the depth test against foreign geometry
        rec_out[2 * (size_t)j] = make_float4(cx, cy, a1x, a1y);
        rec_out[2 * (size_t)j + 1] = make_float4(a2x, a2y, __uint_as_float(q.w), zndc);
      }
    }
  }
  return rect;
}

// BY_ENTRY (slab path): one thread per ENTRY j of the current slab's draw order; record and rectangle are stored at j
// (the slab's instances then carry j, not the splat index), and only the slab's splats are projected.
// Index order, sparse frames (fewer than half of the resident splats passed the worker filter, e.g. a cutout box): the
// survivors of every 1024-splat chunk are first compacted into shared memory, so the shader runs in full warps
// instead of warps with a few live lanes each (20 M splats, 22 % kept: 0.41 ms -> see DESIGN.md 4).
template <bool BY_ENTRY>
__global__ void __launch_bounds__(256) k_project(const float4 *__restrict__ cs, const uint4 *__restrict__ cc,
                                                 const float *__restrict__ depth,
                                                 const FrameParams *__restrict__ fp, float4 *__restrict__ rec_out,
                                                 uint32_t *__restrict__ rect_out, const uint32_t *__restrict__ order,
                                                 const FrameCounters *__restrict__ ctr) {
  GS_PDL_ENTRY();
  const RenderConsts &rc = fp->rc;
  const uint32_t n = BY_ENTRY ? ctr->sort.n_valid : fp->n_splats;
  const uint32_t stride = gridDim.x * blockDim.x;
  if (!BY_ENTRY && (unsigned long long)ctr->sort.n_valid * 2ull < n) {
    constexpr uint32_t kChunk = 1024;  // 4 splats per thread
    __shared__ uint32_t s_list[kChunk];
    __shared__ uint32_t s_warp[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nchunks = (n + kChunk - 1) / kChunk;
    for (uint32_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint32_t base = ch * kChunk + tid * 4u;
      float d[4];
      if (base + 4u <= n) {
        const float4 v = __ldg((const float4 *)(depth + base));
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        // no rectangle unless a survivor stores one after the barrier below
        *(uint4 *)(rect_out + base) = make_uint4(kNoRect, kNoRect, kNoRect, kNoRect);
      } else {
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
          d[k] = (base + k < n) ? __ldg(depth + base + k) : GS_DEPTH_REJECT;
          if (base + k < n) rect_out[base + k] = kNoRect;
        }
      }
      bool f[4];
      uint32_t m = 0;
#pragma unroll
      for (uint32_t k = 0; k < 4; ++k) {
        f[k] = (base + k < n) && ((d[k] != GS_DEPTH_REJECT) || (base + k == 0u));  // splat 0: quirk Q5's zero tail
        m += f[k] ? 1u : 0u;
      }
      uint32_t incl = m;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      uint32_t pos = incl - m, total = 0;
#pragma unroll
      for (uint32_t w = 0; w < 8; ++w) {
        const uint32_t t = s_warp[w];
        pos += (w < warp) ? t : 0u;
        total += t;
      }
#pragma unroll
      for (uint32_t k = 0; k < 4; ++k)
        if (f[k]) s_list[pos++] = base + k;
      __syncthreads();
      for (uint32_t q = tid; q < total; q += blockDim.x) {
        const uint32_t i = s_list[q];
        const uint32_t rect = project_one(rc, cs, cc, i, i, rec_out);
        if (rect != kNoRect) rect_out[i] = rect;
      }
      __syncthreads();
    }
    return;
  }
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const uint32_t i = BY_ENTRY ? __ldg(order + j) : j;
    uint32_t rect = kNoRect;
    const bool sorted = BY_ENTRY || (__ldg(depth + i) != GS_DEPTH_REJECT) || (i == 0);
    if (sorted) rect = project_one(rc, cs, cc, i, j, rec_out);
    rect_out[j] = rect;
  }
}


constexpr int kEmitPerThread = kRadixTile / 2 / kEmitThreads;  // 8: one emission window = half a radix chunk
constexpr int kEmitWindow = kEmitThreads * kEmitPerThread;  // 2048 instances per CTA iteration

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
// K3a: per draw-order entry: its splat, rectangle and exclusive instance offset inside its slice of 256
// entries; per slice: its instance total; the LAST CTA to finish scans the slice totals (-> slice_prefix, D).
// ---------------------------------------------------------------------------------------------
// SLAB: rect is indexed by entry (k_project<true>), the entry's payload is j itself, and entries whose (small)
// rectangle holds only closed bins own nothing any more.
template <bool SLAB>
__global__ void __launch_bounds__(kEmitThreads) k_count(const uint32_t *__restrict__ order,
                                                        const uint32_t *__restrict__ rect,
                                                        uint2 *__restrict__ ent, uint32_t *__restrict__ ent_off,
                                                        uint32_t *__restrict__ slice_total,
                                                        uint32_t *__restrict__ slice_prefix, FrameCounters *ctr,
                                                        const FrameParams *__restrict__ fp,
                                                        const uint32_t *__restrict__ bin_open) {
  GS_PDL_ENTRY();
  const uint32_t shard_rank = fp->rc.shard_rank, shard_world = fp->rc.shard_world;
  __shared__ uint32_t s_warp[kEmitThreads / 32], s_vis[kEmitThreads / 32];
  __shared__ uint32_t s_last;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nv = ctr->sort.n_valid;
  const uint32_t num_slices = (nv + kEmitTile - 1) / kEmitTile;
  for (uint32_t sl = blockIdx.x; sl < num_slices; sl += gridDim.x) {
    const uint32_t j = sl * kEmitTile + tid;
    uint32_t idx = 0, r = kNoRect;
    if (j < nv) {
      idx = SLAB ? j : __ldg(order + j);
      r = __ldg(rect + idx);
    }
    uint32_t cnt = rect_count(r, shard_rank, shard_world);
    if (SLAB && cnt) {
      const uint32_t bx0 = r & 255u, bx1 = (r >> 8) & 255u, by0 = (r >> 16) & 255u, by1 = r >> 24;
      if ((bx1 - bx0 + 1u) * (by1 - by0 + 1u) <= 4u) {  // bins of other ranks count as closed (k_slab_init)
        bool any = false;
        for (uint32_t by = by0; by <= by1; ++by)
          for (uint32_t bx = bx0; bx <= bx1; ++bx) any = any || (__ldg(bin_open + by * fp->rc.bins_x + bx) != 0u);
        if (!any) cnt = 0;
      }
    }
    uint32_t incl = cnt, vis = (r != kNoRect);
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    for (int o = 16; o > 0; o >>= 1) vis += __shfl_xor_sync(0xffffffffu, vis, o);
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_vis[warp] = vis;
    __syncthreads();
    uint32_t wbase = 0, total = 0, v = 0;
    for (uint32_t k = 0; k < kEmitThreads / 32; ++k) {
      if (k < warp) wbase += s_warp[k];
      total += s_warp[k];
      v += s_vis[k];
    }
    if (j < nv) {
      ent[j] = make_uint2(idx, cnt ? r : kNoRect);  // entries owning no tile are skipped by the emit walk
      ent_off[j] = wbase + incl - cnt;
    }
    if (tid == 0) {
      slice_total[sl] = total;
      if (v) atomicAdd(&ctr->n_visible, v);
    }
    __syncthreads();
  }
  // ---- last CTA: exclusive scan of the slice totals ----
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&ctr->count_done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  unsigned long long carry = 0;
  for (uint32_t b = 0; b < num_slices; b += kEmitThreads) {
    const uint32_t i = b + tid;
    const uint32_t v = (i < num_slices) ? __ldcg(slice_total + i) : 0u;
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    for (uint32_t k = 0; k < kEmitThreads / 32; ++k) {
      if (k < warp) wbase += s_warp[k];
      total += s_warp[k];
    }
    // positions are 32-bit: a frame with >= 2^32 candidates overflows the instance buffer long before
    if (i < num_slices) slice_prefix[i] = (uint32_t)(carry + wbase + incl - v);
    carry += total;
  }
  if (tid == 0) {
    slice_prefix[num_slices] = (uint32_t)(carry > 0xFFFFFFFFull ? 0xFFFFFFFFull : carry);
    ctr->n_inst = carry;
  }
}

// ---------------------------------------------------------------------------------------------
// K3b: ordered instance emission, balanced by INSTANCES: CTA iteration = one window of 2048 consecutive
// positions of the instance array (near splats own thousands of tiles, far ones a few: balancing by entries
// would leave a few CTAs with most of the work).
//   1. two warps locate the window's first / last draw-order entry with 32-ary searches (5 dependent loads);
//   2. the entries of the window (and the footprint geometry of their splats) are staged in shared memory,
//      coalesced, at most kEmitEnt at a time;
//   3. thread t generates positions [8t, 8t+8) of the window: one binary search in shared memory, then an
//      incremental walk over tiles / entries; every candidate tile of the bounding rectangle is tested exactly
//      against the r<=2 footprint (closest point of the tile's pixel-centre box in the splat's (px,py) frame),
//      rejected tiles become kNoTile and are dropped by the T1 pass;
//   4. the window is written out coalesced.
// Position = prefix(entry) + k, so the array is in draw order whatever the execution order.
// ---------------------------------------------------------------------------------------------
constexpr int kEmitEnt = 896;  // staged entries per pass (static shared memory stays under 48 KB)

// largest i in [lo, hi) with key(i) <= p, keys non-decreasing, key(lo) <= p; all 32 lanes of a warp cooperate
template <class F>
__device__ __forceinline__ uint32_t warp_search_le(uint32_t lo, uint32_t hi, uint32_t p, F key) {
  const uint32_t lane = threadIdx.x & 31;
  while (hi - lo > 1) {
    const uint32_t step = (hi - lo + 31) / 32;
    const uint32_t i = lo + lane * step;
    const bool ok = (i < hi) && (key(i) <= p);
    const uint32_t c = __popc(__ballot_sync(0xffffffffu, ok));  // lanes 0..c-1 are ok (c >= 1)
    const uint32_t nlo = lo + (c - 1) * step;
    hi = min(hi, nlo + step);
    lo = nlo;
  }
  return lo;
}

__global__ void __launch_bounds__(kEmitThreads, 4) k_emit(const uint2 *__restrict__ ent,
                                                          const uint32_t *__restrict__ ent_off,
                                                          const uint32_t *__restrict__ slice_prefix,
                                                          const float4 *__restrict__ proj_rec,
                                                          const FrameParams *__restrict__ fp, uint64_t cap_inst,
                                                          uint16_t *__restrict__ inst_tile, uint32_t *__restrict__ inst_idx,
                                                          uint32_t *__restrict__ table_t1, uint32_t table_stride,
                                                          FrameCounters *ctr, const uint32_t *__restrict__ bin_open) {
  GS_PDL_ENTRY();
  const RenderConsts &rc = fp->rc;
  __shared__ uint32_t s_wi[kEmitThreads * (kEmitPerThread + 1)];  // stride 9: conflict-free staging
  __shared__ uint16_t s_wt[kEmitThreads * (kEmitPerThread + 1)];
  __shared__ uint2 s_ent[kEmitEnt];         // {splat index, rect}
  __shared__ uint32_t s_goff[kEmitEnt + 1];  // global instance offset of each staged entry
  __shared__ float4 s_g0[kEmitEnt];         // cx, cy, a1x, a1y
  __shared__ float2 s_g1[kEmitEnt];         // a2x, a2y
  __shared__ uint32_t s_jfl[2];
  __shared__ uint32_t s_cnt[kEmitThreads / 32];
  __shared__ uint32_t s_hist[256];  // low tile-id byte of the kept instances of this window: pass T1's histogram column
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  s_hist[tid] = 0;
  const uint32_t nv = ctr->sort.n_valid;
  const uint32_t num_slices = (nv + kEmitTile - 1) / kEmitTile;
  const unsigned long long d_all = ctr->n_inst;
  if (d_all > cap_inst) {  // instance buffer too small: the host regrows it and re-runs the frame
    if (blockIdx.x == 0 && tid == 0) ctr->overflow = 1u;
    return;
  }
  const uint32_t total = (uint32_t)d_all;
  const uint32_t num_windows = (total + kEmitWindow - 1) / kEmitWindow;
  auto goff = [&](uint32_t j) -> uint32_t {  // global instance offset of entry j (j == nv: the total)
    return j < nv ? __ldg(slice_prefix + (j >> 8)) + __ldg(ent_off + j) : total;
  };
  static_assert(kEmitTile == 256, "entry -> slice is j >> 8");
  static_assert(kEmitThreads == 256 && kEmitWindow * 2 == kRadixTile, "one histogram column per window, two per radix chunk");

  for (uint32_t win = blockIdx.x; win < num_windows; win += gridDim.x) {
    const uint32_t wb = win * kEmitWindow, we = min(wb + (uint32_t)kEmitWindow, total);
    // ---- 1. first / last entry of the window ----
    if (warp < 2) {
      const uint32_t p = warp == 0 ? wb : we - 1;
      const uint32_t sl = warp_search_le(0u, num_slices, p, [&](uint32_t i) { return __ldg(slice_prefix + i); });
      const uint32_t rel = p - __ldg(slice_prefix + sl);
      const uint32_t a = sl * kEmitTile, b = min(a + (uint32_t)kEmitTile, nv);
      const uint32_t j = warp_search_le(a, b, rel, [&](uint32_t i) { return __ldg(ent_off + i); });
      if (lane == 0) s_jfl[warp] = j;
    }
    __syncthreads();
    const uint32_t j_last = s_jfl[1];
    uint32_t j0 = s_jfl[0], pos = wb;
    while (pos < we) {
      // ---- 2. stage the entries that own tiles, compacted in order, from source entries [j0, ...) ----
      // (with the frame sharded over GPUs most entries own no tile of this rank: skipping them here keeps the
      //  staged window dense); batches of 256 source entries until the staging arrays are full
      uint32_t nE = 0, jn = j0;
      while (jn <= j_last && nE + kEmitThreads <= (uint32_t)kEmitEnt) {
        if ((jn & (uint32_t)(kEmitTile - 1)) == 0u) {
          // at a slice boundary: skip runs of slices that own no instance at all (slab path after most bins have
          // closed, or a rank that owns few of a region's bins) - 256 slices = 65536 entries per look
          const uint32_t sl = (jn >> 8) + tid, sl_last = j_last >> 8;
          const bool stop = (sl > sl_last) || (__ldg(slice_prefix + sl + 1) != __ldg(slice_prefix + sl));
          const uint32_t bal0 = __ballot_sync(0xffffffffu, stop);
          if (lane == 0) s_cnt[warp] = bal0 ? warp * 32u + (uint32_t)__ffs(bal0) - 1u : (uint32_t)kEmitThreads;
          __syncthreads();
          uint32_t skip = kEmitThreads;
          for (uint32_t k2 = 0; k2 < kEmitThreads / 32; ++k2) skip = min(skip, s_cnt[k2]);
          __syncthreads();  // s_cnt is reused by the batch below
          jn += skip * (uint32_t)kEmitTile;
          if (skip == (uint32_t)kEmitThreads) continue;
          if (jn > j_last) break;
        }
        const uint32_t jend = (jn | (uint32_t)(kEmitTile - 1)) + 1u;  // batches end at slice boundaries
        const uint32_t j = jn + tid;
        uint2 en = make_uint2(0u, kNoRect);
        if (j < jend && j <= j_last) en = __ldg(ent + j);
        const bool nz = en.y != kNoRect;
        const uint32_t bal = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t base_w = nE, tot = 0;
        for (uint32_t k2 = 0; k2 < kEmitThreads / 32; ++k2) {
          if (k2 < warp) base_w += s_cnt[k2];
          tot += s_cnt[k2];
        }
        if (nz) {
          const uint32_t q = base_w + __popc(bal & ((1u << lane) - 1u));
          s_ent[q] = en;
          s_goff[q] = goff(j);
          const uint32_t r = en.y;
          const uint32_t wfull = ((r >> 8) & 255u) - (r & 255u) + 1u, hfull = (r >> 24) - ((r >> 16) & 255u) + 1u;
          if (wfull * hfull > 1u) {  // footprint geometry for the exact tile test
            s_g0[q] = __ldg(proj_rec + 2 * (size_t)en.x);
            s_g1[q] = __ldg((const float2 *)(proj_rec + 2 * (size_t)en.x + 1));
          }
        }
        nE += tot;
        jn = jend;
        __syncthreads();
      }
      const uint32_t jE = min(jn, j_last + 1);
      if (tid == 0) s_goff[nE] = goff(jE);  // first position NOT covered by the staged entries
      __syncthreads();
      const uint32_t pe = min(we, s_goff[nE]);  // positions [pos, pe) belong to the staged entries
      // ---- 3. generate ----
      const uint32_t t0 = max(pos, wb + tid * kEmitPerThread), t1 = min(pe, wb + (tid + 1) * kEmitPerThread);
      if (t0 < t1) {
        uint32_t lo = 0, hi = nE;
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (s_goff[mid] <= t0) lo = mid; else hi = mid;
        }
        uint32_t e = lo;                  // staged entry
        uint32_t k = t0 - s_goff[lo];     // position inside it
        uint32_t idx = 0, txf = 0, w = 1, step = 1, tx = 0, ty = 0, n_all = 0, n_own = 0;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 r1 = make_float2(0.f, 0.f);
        float cross = 0.f, inv_yy = 0.f, inv_xx = 0.f;
        auto load_entry = [&](uint32_t ee, uint32_t kk) {
          const uint2 en = s_ent[ee];
          const uint32_t r = en.y;
          n_own = 0;
          if (r == kNoRect) return;
          idx = en.x;
          uint32_t tx0 = r & 255u;
          const uint32_t ty0 = (r >> 16) & 255u, h = (r >> 24) - ty0 + 1u;
          w = ((r >> 8) & 255u) - tx0 + 1u;
          n_all = w * h;
          step = 1u;
          if (rc.shard_world > 1) {  // owned columns of the rectangle: first, first + world, ...
            owned_span(tx0, (r >> 8) & 255u, rc.shard_rank, rc.shard_world, tx0, w);
            step = rc.shard_world;
          }
          n_own = w * h;
          if (n_own == 0) return;
          txf = tx0;
          // kk / w for kk < 65536, w <= 256: the float quotient of (kk + 0.5) is never within rounding of an integer
          const uint32_t row = (uint32_t)__fdividef((float)kk + 0.5f, (float)w);
          tx = tx0 + (kk - row * w) * step;
          ty = ty0 + row;
          if (n_all > 1) {
            r0 = s_g0[ee];
            r1 = s_g1[ee];
            // q(d) = |(a2.d, a1.d)|^2 = M00 dx^2 + 2 M01 dx dy + M11 dy^2: edge minimisers need M01/M11, M01/M00
            cross = r1.x * r1.y + r0.z * r0.w;
            inv_yy = __fdividef(1.0f, r1.y * r1.y + r0.w * r0.w);
            inv_xx = __fdividef(1.0f, r1.x * r1.x + r0.z * r0.z);
          }
        };
        load_entry(e, k);
#pragma unroll 1
        for (uint32_t p = t0; p < t1; ++p) {
          if (k >= n_own) {  // next staged entry (every staged entry owns tiles)
            ++e;
            load_entry(e, 0u);
            k = 0;
          }
          bool keep = true;
          if (n_all > 1) {
            // pixel-centre box of the bin, relative to the splat centre
            const float xa = (float)(tx * kBin) + 0.5f - r0.x, xb = xa + (float)(kBin - 1);
            const float ya = (float)(ty * kBin) + 0.5f - r0.y, yb = ya + (float)(kBin - 1);
            const bool in_x = (xa <= 0.0f) && (xb >= 0.0f), in_y = (ya <= 0.0f) && (yb >= 0.0f);
            if (!(in_x && in_y)) {
              float qmin = 3.0e38f;
              if (!in_x) {  // nearest vertical edge, minimise over y on it; (px,py) evaluated at the found point
                const float dx = (xa > 0.0f) ? xa : xb;
                const float t = fminf(fmaxf(-dx * cross * inv_yy, ya), yb);
                const float px = dx * r1.x + t * r1.y, py = dx * r0.z + t * r0.w;
                qmin = px * px + py * py;
              }
              if (!in_y) {  // nearest horizontal edge, minimise over x on it
                const float dy = (ya > 0.0f) ? ya : yb;
                const float t = fminf(fmaxf(-dy * cross * inv_xx, xa), xb);
                const float px = t * r1.x + dy * r1.y, py = t * r0.z + dy * r0.w;
                qmin = fminf(qmin, px * px + py * py);
              }
              keep = !(qmin > 4.02f);  // r^2 <= 4 with slack for fp32 rounding of the closest-point search
            }
          }
          uint32_t t = kNoTile;
          if (keep && bin_open) keep = __ldg(bin_open + ty * rc.bins_x + tx) != 0u;  // slab path: closed bins take nothing
          if (keep) {
            t = ty * rc.bins_x + tx;
            atomicAdd(&s_hist[t & 255u], 1u);
          }
          const uint32_t q = p - wb;  // window-relative position
          const uint32_t si = (q / kEmitPerThread) * (kEmitPerThread + 1) + (q % kEmitPerThread);
          s_wt[si] = (uint16_t)t;
          s_wi[si] = idx;
          // advance inside the rectangle (row-major over the owned columns)
          ++k;
          tx += step;
          if (tx >= txf + w * step) { tx = txf; ++ty; }
        }
      }
      __syncthreads();
      pos = pe;
      j0 = jE;
    }
    // ---- 4. write the window out ----
    const uint32_t wn = we - wb;
    for (uint32_t i = tid; i < wn; i += kEmitThreads) {
      const uint32_t si = (i / kEmitPerThread) * (kEmitPerThread + 1) + (i % kEmitPerThread);
      inst_tile[(size_t)wb + i] = s_wt[si];
      inst_idx[(size_t)wb + i] = s_wi[si];
    }
    table_t1[(size_t)tid * table_stride + win] = s_hist[tid];  // kEmitThreads == 256 digits
    s_hist[tid] = 0;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// K3b', slab path: instance emission by ENTRY.  In a slab most entries own nothing any more (closed bins, other
// ranks' bins), so the instance-window walk of k_emit would spend its time stepping over dead entries; here one thread
// owns one live entry and writes its instances at the entry's own offset (position = prefix(entry) + k, so the array is
// still in draw order).  Entries with large rectangles are finished by their whole warp, 32 candidates per step.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_candidate(const RenderConsts &rc, uint32_t bx, uint32_t by, bool multi, const float4 &r0,
                                               const float2 &r1, const uint32_t *__restrict__ bin_open, uint32_t payload,
                                               size_t pos, uint16_t *__restrict__ inst_tile, uint32_t *__restrict__ inst_idx) {
  bool keep = true;
  if (multi)
    keep = footprint_meets_box(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, (float)(bx * kBin) + 0.5f, (float)(by * kBin) + 0.5f,
                               (float)(kBin - 1));
  const uint32_t t = by * rc.bins_x + bx;
  if (keep && bin_open) keep = __ldg(bin_open + t) != 0u;
  inst_tile[pos] = keep ? (uint16_t)t : kNoTile;
  inst_idx[pos] = payload;
}

__global__ void __launch_bounds__(256) k_emit_entries(const uint2 *__restrict__ ent, const uint32_t *__restrict__ ent_off,
                                                      const uint32_t *__restrict__ slice_prefix,
                                                      const float4 *__restrict__ proj_rec, const FrameParams *__restrict__ fp,
                                                      uint64_t cap_inst, uint16_t *__restrict__ inst_tile,
                                                      uint32_t *__restrict__ inst_idx, FrameCounters *ctr,
                                                      const uint32_t *__restrict__ bin_open) {
  GS_PDL_ENTRY();
  const RenderConsts &rc = fp->rc;
  const uint32_t nv = ctr->sort.n_valid;
  if (ctr->n_inst > cap_inst) {  // instance buffer too small: the host regrows it and re-runs the frame
    if (blockIdx.x == 0 && threadIdx.x == 0) ctr->overflow = 1u;
    return;
  }
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t nv_pad = (nv + 31u) & ~31u;  // whole warps stay in the loop (cooperative part below)
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nv_pad; j += stride) {
    uint2 en = make_uint2(0u, kNoRect);
    if (j < nv) en = __ldg(ent + j);
    const uint32_t r = en.y;
    uint32_t bx0 = 0, by0 = 0, w = 0, h = 0, step = 1, n_own = 0;
    size_t base = 0;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 r1 = make_float2(0.f, 0.f);
    bool multi = false;
    if (r != kNoRect) {
      bx0 = r & 255u;
      by0 = (r >> 16) & 255u;
      h = (r >> 24) - by0 + 1u;
      w = ((r >> 8) & 255u) - bx0 + 1u;
      multi = w * h > 1u;
      if (rc.shard_world > 1) {  // owned columns of the rectangle: first, first + world, ...
        owned_span(bx0, (r >> 8) & 255u, rc.shard_rank, rc.shard_world, bx0, w);
        step = rc.shard_world;
      }
      n_own = w * h;
      base = (size_t)__ldg(slice_prefix + (j >> 8)) + __ldg(ent_off + j);
      if (multi) {
        r0 = __ldg(proj_rec + 2 * (size_t)en.x);
        r1 = __ldg((const float2 *)(proj_rec + 2 * (size_t)en.x + 1));
      }
    }
    const bool big = n_own > 8u;
    if (!big) {
      for (uint32_t k = 0; k < n_own; ++k) {
        const uint32_t row = k / w;
        emit_candidate(rc, bx0 + (k - row * w) * step, by0 + row, multi, r0, r1, bin_open, en.x, base + k, inst_tile, inst_idx);
      }
    }
    // large rectangles: the warp finishes them together
    uint32_t todo = __ballot_sync(0xffffffffu, big);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t s_bx0 = __shfl_sync(0xffffffffu, bx0, src), s_by0 = __shfl_sync(0xffffffffu, by0, src);
      const uint32_t s_w = __shfl_sync(0xffffffffu, w, src), s_step = __shfl_sync(0xffffffffu, step, src);
      const uint32_t s_n = __shfl_sync(0xffffffffu, n_own, src), s_pay = __shfl_sync(0xffffffffu, en.x, src);
      const unsigned long long s_base = __shfl_sync(0xffffffffu, (unsigned long long)base, src);
      float4 g0;
      float2 g1;
      g0.x = __shfl_sync(0xffffffffu, r0.x, src); g0.y = __shfl_sync(0xffffffffu, r0.y, src);
      g0.z = __shfl_sync(0xffffffffu, r0.z, src); g0.w = __shfl_sync(0xffffffffu, r0.w, src);
      g1.x = __shfl_sync(0xffffffffu, r1.x, src); g1.y = __shfl_sync(0xffffffffu, r1.y, src);
      for (uint32_t k = lane; k < s_n; k += 32u) {
        const uint32_t row = k / s_w;
        emit_candidate(rc, s_bx0 + (k - row * s_w) * s_step, s_by0 + row, true, g0, g1, bin_open, s_pay, (size_t)s_base + k, inst_tile,
                       inst_idx);
      }
    }
  }
}

void launch_project(gs_context *c, const FrameParams *fp, const FrameCounters *ctr, const FrameBufs &b, cudaStream_t stream) {
  uint64_t blocks = ((uint64_t)c->cap + 255) / 256;
  const uint64_t cap = (uint64_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  launch_chain(c, k_project<false>, (int)blocks, 256, stream, (const float4 *)c->center_scale, (const uint4 *)c->cov_color, (const float *)c->depth, fp, b.proj_rec, b.rect,
               (const uint32_t *)nullptr, ctr);  // ctr: the sorted count picks the sparse-frame path
}

void launch_project_entries(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t stream) {
  uint64_t blocks = ((uint64_t)c->cap + 255) / 256;
  const uint64_t cap = (uint64_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  launch_chain(c, k_project<true>, (int)blocks, 256, stream, (const float4 *)c->center_scale, (const uint4 *)c->cov_color, (const float *)c->depth, fp, b.proj_rec, b.rect,
               (const uint32_t *)b.order, (const FrameCounters *)ctr);
}

static void launch_emit_impl(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st, bool slab);
void launch_emit(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st) {
  launch_emit_impl(c, fp, ctr, b, st, false);
}
void launch_emit_slab(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st) {
  launch_emit_impl(c, fp, ctr, b, st, true);
}
static void launch_emit_impl(gs_context *c, const FrameParams *fp, FrameCounters *ctr, const FrameBufs &b, cudaStream_t st, bool slab) {
  uint64_t tiles = ((uint64_t)c->cap + kEmitTile - 1) / kEmitTile;
  const uint64_t cap = (uint64_t)c->sm_count * 8;
  if (tiles > cap) tiles = cap;
  if (tiles < 1) tiles = 1;
  if (slab)
    launch_chain(c, k_count<true>, (int)tiles, kEmitThreads, st, (const uint32_t *)b.order, (const uint32_t *)b.rect, c->ent, c->ent_off,
                 c->slice_total, c->slice_prefix, ctr, fp, (const uint32_t *)c->bin_open);
  else
    launch_chain(c, k_count<false>, (int)tiles, kEmitThreads, st, (const uint32_t *)b.order, (const uint32_t *)b.rect, c->ent, c->ent_off,
                 c->slice_total, c->slice_prefix, ctr, fp, (const uint32_t *)nullptr);
  if (slab || c->emit_by_entry) {
    launch_chain(c, k_emit_entries, (int)tiles, 256, st, (const uint2 *)c->ent, (const uint32_t *)c->ent_off, (const uint32_t *)c->slice_prefix,
                 (const float4 *)b.proj_rec, fp, (uint64_t)c->cap_inst, c->inst_tile, c->inst_idx, ctr,
                 (const uint32_t *)(slab ? c->bin_open : nullptr));
    return;
  }
  uint64_t wins = (c->cap_inst + kEmitWindow - 1) / kEmitWindow;
  if (wins > (uint64_t)c->sm_count * 4) wins = (uint64_t)c->sm_count * 4;
  if (wins < 1) wins = 1;
  k_emit<<<(int)wins, kEmitThreads, 0, st>>>(c->ent, c->ent_off, c->slice_prefix, b.proj_rec, fp, c->cap_inst,
                                                    c->inst_tile, c->inst_idx, c->table_d, c->table_d_stride, ctr,
                                                    slab ? c->bin_open : nullptr);
}

}  // namespace gs

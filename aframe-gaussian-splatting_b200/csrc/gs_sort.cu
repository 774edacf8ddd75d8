// gs_sort.cu — device restatement of the Web-Worker `sortSplats` (reference index.js:507-570)
// and the stable LSD radix passes shared by the depth sort and the tile binning.
//
//   k_depth_cull  : index.js:517-555  depth (fp64, left to right), cutout box, filter, min/max
//   k_key_hist    : index.js:557-563  16-bit key = ToInt32((f32 depth - min) * depthInv), digit histograms
//   k_radix<D1/D2>: index.js:564-567  stable counting sort, as two 8-bit passes with decoupled look-back
//   k_radix<T1/T2>: stable sort of tile instances by 16-bit tile id (T2 also gathers the 32 B records)
//
// Bit-exactness: JS evaluates in fp64 with IEEE rounding after every operation; the kernels use
// __dmul_rn/__dadd_rn so nothing is contracted, and ToInt32 is restated exactly (js_to_int32).
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
  const double mn = dec_f64(~ctr->min_enc);
  const double mx = dec_f64(ctr->max_enc);
  DepthRange r;
  r.min_depth = mn;
  r.depth_inv = __ddiv_rn(65535.0, __dsub_rn(mx, mn));  // index.js:558
  return r;
}

// ---------------------------------------------------------------------------------------------
// K1: depth + cull + min/max (index.js:517-555).  Reads 16 B + 4 B per splat, writes 4 B.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_depth_cull(const float4 *__restrict__ cs, const float *__restrict__ sa,
                                                    uint32_t n, SortConsts sc, float *__restrict__ depth_out,
                                                    FrameCounters *ctr) {
  double dmin = INFINITY, dmax = -INFINITY;
  uint32_t cnt = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 c = __ldg(cs + i);
    const float s = __ldg(sa + i);
    const double x = c.x, y = c.y, z = c.z;
    // index.js:519-523
    const double depth =
        __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sc.view[0], x), __dmul_rn(sc.view[1], y)), __dmul_rn(sc.view[2], z)),
                  sc.view[3]);
    bool in_box = true;
    if (sc.has_cutout) {
      // index.js:533 -> mul(cutout, x, -y, z) of index.js:492-500 (Q12: centre only, y negated)
      const double *e = sc.cutout;
      const double ny = -y;
      const double w = __ddiv_rn(
          1.0, __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(e[3], x), __dmul_rn(e[7], ny)), __dmul_rn(e[11], z)), e[15]));
      const double c0 = __dmul_rn(
          __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(e[0], x), __dmul_rn(e[4], ny)), __dmul_rn(e[8], z)), e[12]), w);
      const double c1 = __dmul_rn(
          __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(e[1], x), __dmul_rn(e[5], ny)), __dmul_rn(e[9], z)), e[13]), w);
      const double c2 = __dmul_rn(
          __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(e[2], x), __dmul_rn(e[6], ny)), __dmul_rn(e[10], z)), e[14]), w);
      if (c0 < -0.5 || c0 > 0.5 || c1 < -0.5 || c1 > 0.5 || c2 < -0.5 || c2 > 0.5) in_box = false;
    }
    // index.js:548
    const bool keep = (depth < 0.0) && ((double)s > __dmul_rn(-0.0001, depth)) && in_box;
    float out = GS_DEPTH_REJECT;
    if (keep) {
      out = (float)depth;  // Float32Array store (index.js:549)
      ++cnt;
      if (depth > dmax) dmax = depth;
      if (depth < dmin) dmin = depth;
    }
    depth_out[i] = out;
  }
  // block reduction
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, dmin, o);
    const double b = __shfl_xor_sync(0xffffffffu, dmax, o);
    dmin = a < dmin ? a : dmin;
    dmax = b > dmax ? b : dmax;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  __shared__ double s_min[8], s_max[8];
  __shared__ uint32_t s_cnt[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_min[w] = dmin; s_max[w] = dmax; s_cnt[w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      if (s_min[k] < dmin) dmin = s_min[k];
      if (s_max[k] > dmax) dmax = s_max[k];
      cnt += s_cnt[k];
    }
    if (cnt) {
      atomicMax(&ctr->min_enc, ~enc_f64(dmin));
      atomicMax(&ctr->max_enc, enc_f64(dmax));
      atomicAdd(&ctr->n_valid, cnt);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1b: key + digit histograms (index.js:560-563).  Reads 4 B per splat.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_key_hist(const float *__restrict__ depth, uint32_t n, FrameCounters *ctr) {
  __shared__ uint32_t h_lo[256], h_hi[256];
  __shared__ uint32_t s_in, s_drop;
  h_lo[threadIdx.x] = 0;
  h_hi[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_in = 0; s_drop = 0; }
  __syncthreads();
  if (ctr->n_valid != 0) {
    const DepthRange dr = load_depth_range(ctr);
    uint32_t in = 0, drop = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float d = __ldg(depth + i);
      if (d == GS_DEPTH_REJECT) continue;
      const int32_t key = depth_key(d, dr.min_depth, dr.depth_inv);
      if (key < 0 || key > 65535) { ++drop; continue; }  // typed-array write out of range: dropped (Q5)
      ++in;
      atomicAdd(&h_lo[key & 255], 1u);
      atomicAdd(&h_hi[key >> 8], 1u);
    }
    for (int o = 16; o > 0; o >>= 1) {
      in += __shfl_xor_sync(0xffffffffu, in, o);
      drop += __shfl_xor_sync(0xffffffffu, drop, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (in) atomicAdd(&s_in, in);
      if (drop) atomicAdd(&s_drop, drop);
    }
  }
  __syncthreads();
  if (h_lo[threadIdx.x]) atomicAdd(&ctr->hist_lo[threadIdx.x], h_lo[threadIdx.x]);
  if (h_hi[threadIdx.x]) atomicAdd(&ctr->hist_hi[threadIdx.x], h_hi[threadIdx.x]);
  if (threadIdx.x == 0) {
    if (s_in) atomicAdd(&ctr->n_inrange, s_in);
    if (s_drop) atomicAdd(&ctr->n_dropped, s_drop);
  }
}

// ---------------------------------------------------------------------------------------------
// Stable 8-bit radix pass with decoupled look-back (single read, single write per element).
// ---------------------------------------------------------------------------------------------
enum { PASS_D1 = 0, PASS_D2 = 1, PASS_T1 = 2, PASS_T2 = 3 };

struct RadixArgs {
  FrameCounters *ctr;
  uint32_t *status;
  uint32_t n_host;  // D1: number of resident splats
  uint64_t cap_inst;
  // depth passes
  const float *depth;
  uint32_t *idx_a;
  uint8_t *dig_a;
  uint32_t *order;
  // tile passes
  const uint16_t *inst_tile;
  const uint32_t *inst_idx;
  uint8_t *inst_dig_b;
  uint32_t *inst_idx_b;
  const float4 *proj_rec;
  float4 *inst_rec;
};

template <int PASS>
__global__ void __launch_bounds__(kRadixThreads) k_radix(RadixArgs a) {
  __shared__ uint32_t wcnt[kRadixThreads / 32][256];
  __shared__ uint32_t tile_off[256];
  __shared__ uint32_t s_warp_tot[8];
  __shared__ uint32_t s_tile;
  FrameCounters *ctr = a.ctr;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  uint32_t n;
  const uint32_t *hist;
  if (PASS == PASS_D1) { n = a.n_host; hist = ctr->hist_lo; }
  else if (PASS == PASS_D2) { n = ctr->n_inrange; hist = ctr->hist_hi; }
  else if (PASS == PASS_T1) {
    const unsigned long long d = ctr->n_inst;
    n = ctr->overflow ? 0u : (uint32_t)d;
    hist = ctr->thist_lo;
  } else { n = ctr->overflow ? 0u : ctr->n_inst_kept; hist = ctr->thist_hi; }
  const uint32_t num_tiles = (n + kRadixTile - 1) / kRadixTile;

  // exclusive scan of the global digit histogram -> first output slot of each digit
  uint32_t dbase;
  {
    const uint32_t h = hist[tid];
    uint32_t incl = h;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t k = 0; k < warp; ++k) wbase += s_warp_tot[k];
    dbase = wbase + incl - h;
  }

  DepthRange dr{0.0, 0.0};
  if (PASS == PASS_D1) {
    if (ctr->n_valid == 0) return;
    dr = load_depth_range(ctr);
  }
  if (PASS == PASS_D2) {
    // quirk Q5: the reference's output keeps length validCount; slots never written stay 0
    const uint32_t nv = ctr->n_valid;
    for (uint32_t j = n + blockIdx.x * blockDim.x + tid; j < nv; j += gridDim.x * blockDim.x) a.order[j] = 0u;
  }

  while (true) {
    if (tid == 0) s_tile = atomicAdd(&ctr->ticket[PASS], 1u);
    for (uint32_t k = tid; k < (kRadixThreads / 32) * 256; k += kRadixThreads) (&wcnt[0][0])[k] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= num_tiles) break;

    // ---- load (warp-striped: consecutive lanes read consecutive elements) ----
    const uint32_t base = tile * kRadixTile + warp * (32 * kRadixItems) + lane;
    uint32_t digit[kRadixItems], pay[kRadixItems], rank[kRadixItems];
    uint8_t hi[kRadixItems];
#pragma unroll
    for (int s = 0; s < kRadixItems; ++s) {
      const uint32_t i = base + s * 32;
      digit[s] = kInvalidDigit;
      pay[s] = 0;
      hi[s] = 0;
      if (i < n) {
        if (PASS == PASS_D1) {
          const float d = __ldg(a.depth + i);
          if (d != GS_DEPTH_REJECT) {
            const int32_t key = depth_key(d, dr.min_depth, dr.depth_inv);
            if (key >= 0 && key <= 65535) { digit[s] = key & 255; hi[s] = (uint8_t)(key >> 8); pay[s] = i; }
          }
        } else if (PASS == PASS_D2) {
          digit[s] = a.dig_a[i];
          pay[s] = a.idx_a[i];
        } else if (PASS == PASS_T1) {
          const uint16_t t = a.inst_tile[i];
          if (t != kNoTile) { digit[s] = t & 255; hi[s] = (uint8_t)(t >> 8); pay[s] = a.inst_idx[i]; }
        } else {
          digit[s] = a.inst_dig_b[i];
          pay[s] = a.inst_idx_b[i];
        }
      }
    }
    // ---- stable rank inside the warp (input order = lane order within a step, steps in order) ----
#pragma unroll
    for (int s = 0; s < kRadixItems; ++s) {
      const uint32_t d = digit[s];
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const uint32_t lt = __popc(peers & ((1u << lane) - 1u));
      uint32_t prior = 0;
      if (d != kInvalidDigit) prior = wcnt[warp][d];
      __syncwarp();
      if (d != kInvalidDigit && lt == 0) wcnt[warp][d] = prior + __popc(peers);
      __syncwarp();
      rank[s] = prior + lt;
    }
    __syncthreads();
    // ---- thread `tid` owns digit `tid`: scan over warps, publish, look back ----
    {
      uint32_t total = 0;
#pragma unroll
      for (int w = 0; w < kRadixThreads / 32; ++w) {
        const uint32_t c = wcnt[w][tid];
        wcnt[w][tid] = total;
        total += c;
      }
      uint32_t *st = a.status + (size_t)tile * 256 + tid;
      uint32_t excl = 0;
      if (tile == 0) {
        st_relaxed(st, kFlagIncl | total);
      } else {
        st_relaxed(st, kFlagAgg | total);
        uint32_t p = tile - 1;
        while (true) {
          const uint32_t v = ld_relaxed(a.status + (size_t)p * 256 + tid);
          if ((v & kFlagMask) == 0) continue;
          excl += v & kValMask;
          if (v & kFlagIncl) break;
          --p;
        }
        st_relaxed(st, kFlagIncl | (excl + total));
      }
      tile_off[tid] = dbase + excl;
    }
    __syncthreads();
    // ---- scatter ----
#pragma unroll
    for (int s = 0; s < kRadixItems; ++s) {
      const uint32_t d = digit[s];
      if (d == kInvalidDigit) continue;
      const uint32_t pos = tile_off[d] + wcnt[warp][d] + rank[s];
      if (PASS == PASS_D1) {
        a.idx_a[pos] = pay[s];
        a.dig_a[pos] = hi[s];
      } else if (PASS == PASS_D2) {
        a.order[pos] = pay[s];
      } else if (PASS == PASS_T1) {
        a.inst_idx_b[pos] = pay[s];
        a.inst_dig_b[pos] = hi[s];
      } else {
        const float4 r0 = __ldg(a.proj_rec + 2 * (size_t)pay[s]);
        const float4 r1 = __ldg(a.proj_rec + 2 * (size_t)pay[s] + 1);
        a.inst_rec[2 * (size_t)pos] = r0;
        a.inst_rec[2 * (size_t)pos + 1] = r1;
      }
    }
    __syncthreads();
  }
}

static int persistent_grid(gs_context *c, uint64_t n_elems, int per_cta, int ctas_per_sm) {
  uint64_t tiles = (n_elems + per_cta - 1) / per_cta;
  uint64_t cap = (uint64_t)c->sm_count * ctas_per_sm;
  if (tiles < 1) tiles = 1;
  return (int)(tiles < cap ? tiles : cap);
}

void launch_depth_cull(gs_context *c, const SortConsts &sc) {
  const int grid = persistent_grid(c, c->n, 256 * 4, 8);
  k_depth_cull<<<grid, 256, 0, c->stream>>>(c->center_scale, c->size_alpha, c->n, sc, c->depth, c->counters);
}

void launch_key_hist(gs_context *c) {
  const int grid = persistent_grid(c, c->n, 256 * 8, 8);
  k_key_hist<<<grid, 256, 0, c->stream>>>(c->depth, c->n, c->counters);
}

static RadixArgs make_args(gs_context *c) {
  RadixArgs a{};
  a.ctr = c->counters;
  a.n_host = c->n;
  a.cap_inst = c->cap_inst;
  a.depth = c->depth;
  a.idx_a = c->idx_a;
  a.dig_a = c->dig_a;
  a.order = c->order;
  a.inst_tile = c->inst_tile;
  a.inst_idx = c->inst_idx;
  a.inst_dig_b = c->inst_dig_b;
  a.inst_idx_b = c->inst_idx_b;
  a.proj_rec = c->proj_rec;
  a.inst_rec = c->inst_rec;
  return a;
}

void launch_depth_radix(gs_context *c) {
  RadixArgs a = make_args(c);
  const int grid = persistent_grid(c, c->n, kRadixTile, 4);
  a.status = c->status_d1;
  k_radix<PASS_D1><<<grid, kRadixThreads, 0, c->stream>>>(a);
  a.status = c->status_d2;
  k_radix<PASS_D2><<<grid, kRadixThreads, 0, c->stream>>>(a);
}

void launch_tile_radix(gs_context *c) {
  RadixArgs a = make_args(c);
  const int grid = persistent_grid(c, c->cap_inst, kRadixTile, 4);
  a.status = c->status_t1;
  k_radix<PASS_T1><<<grid, kRadixThreads, 0, c->stream>>>(a);
  a.status = c->status_t2;
  k_radix<PASS_T2><<<grid, kRadixThreads, 0, c->stream>>>(a);
}

}  // namespace gs

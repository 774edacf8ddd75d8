// gs_pack.cu — load-time pack on the device: the reference's `pushDataBuffer` loop (index.js:343-402).
//
// One thread per .splat row, all arithmetic in fp64 exactly as JavaScript evaluates it (Three.js r147
// Matrix4.compose / transpose / scale / premultiply restated entry by entry, sums left to right, no FMA):
//   q = ((b29-128)/128, (b30-128)/128, -(b31-128)/128, (b28-128)/128)   not normalised (quirk Q1)
//   Sigma = (R^T diag(s)) (R^T diag(s))^T ;  maxAbs over the 6 unique entries
//   float4 {cx, cy, -cz_file, (f32)(maxAbs/32767)} ; six int16 = parseInt(Sigma_ij*32767/maxAbs) ; rgba8
//   sizeAlpha = (f32)(max(scale) * alpha / 255)
// parseInt(Number) stringifies first (quirk Q2): for 0 < |x| < 1e-6 the string is in exponent form and
// parseInt returns the sign and FIRST digit of the shortest round-trip decimal.  That digit is d iff
// strtod("d e-k") <= |x| < strtod("(d+1) e-k"); the thresholds are tabulated on the host at gs_create
// (strtod is correctly rounded) and binary-searched here.
#include "gs_common.cuh"

namespace gs {

__device__ __forceinline__ int32_t to_int32_wrap(double d) {
  if (!isfinite(d)) return 0;
  double t = trunc(d);
  if (t >= -2147483648.0 && t <= 2147483647.0) return (int32_t)t;
  double m = fmod(t, 4294967296.0);
  if (m < 0) m += 4294967296.0;
  return (int32_t)(uint32_t)m;
}

// parseInt(Number) -> Int16Array store (NaN -> 0, ToInt16 wrap)
__device__ __forceinline__ int16_t parse_int_to_i16(double x, const double *__restrict__ tab, int nt) {
  if (!isfinite(x)) return 0;
  if (x == 0.0) return 0;
  const double ax = fabs(x);
  double r;
  if (ax >= 1e-6) {
    r = trunc(x);  // plain decimal notation: parseInt reads the integer part (|x| < 1e21 always holds here)
  } else {
    // largest table entry <= ax; entry e encodes digit (e % 9) + 1
    int lo = 0, hi = nt;  // invariant: tab[lo] <= ax (if any), answer in [lo, hi)
    if (ax < tab[0]) {
      r = 1.0;  // below 1e-323: not representable as a distinct one-digit decimal; unreachable in practice
    } else {
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tab[mid] <= ax) lo = mid; else hi = mid;
      }
      r = (double)((lo % 9) + 1);
    }
    if (x < 0) r = -r;
  }
  return (int16_t)(uint16_t)((uint32_t)to_int32_wrap(r) & 0xFFFFu);
}

__global__ void __launch_bounds__(256) k_pack(const uint4 *__restrict__ rows, uint32_t first, uint32_t n,
                                              float4 *__restrict__ cs, uint4 *__restrict__ cc,
                                              float *__restrict__ sa, const double *__restrict__ tab, int nt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 a = __ldg(rows + 2 * (size_t)i);      // pos.xyz, scale.x
  const uint4 b = __ldg(rows + 2 * (size_t)i + 1);  // scale.yz, rgba, rot
  const double px = __uint_as_float(a.x), py = __uint_as_float(a.y), pz = __uint_as_float(a.z);
  const double sx = __uint_as_float(a.w), sy = __uint_as_float(b.x), sz = __uint_as_float(b.y);
  const uint32_t rgba = b.z, rot = b.w;
  // index.js:344-349
  const double qx = __ddiv_rn(__dsub_rn((double)((rot >> 8) & 255u), 128.0), 128.0);
  const double qy = __ddiv_rn(__dsub_rn((double)((rot >> 16) & 255u), 128.0), 128.0);
  const double qz = -__ddiv_rn(__dsub_rn((double)(rot >> 24), 128.0), 128.0);
  const double qw = __ddiv_rn(__dsub_rn((double)(rot & 255u), 128.0), 128.0);
  // Matrix4.compose (index.js:362)
  const double x2 = __dadd_rn(qx, qx), y2 = __dadd_rn(qy, qy), z2 = __dadd_rn(qz, qz);
  const double xx = __dmul_rn(qx, x2), xy = __dmul_rn(qx, y2), xz = __dmul_rn(qx, z2);
  const double yy = __dmul_rn(qy, y2), yz = __dmul_rn(qy, z2), zz = __dmul_rn(qz, z2);
  const double wx = __dmul_rn(qw, x2), wy = __dmul_rn(qw, y2), wz = __dmul_rn(qw, z2);
  // R(row, col)
  const double R00 = __dsub_rn(1.0, __dadd_rn(yy, zz)), R10 = __dadd_rn(xy, wz), R20 = __dsub_rn(xz, wy);
  const double R01 = __dsub_rn(xy, wz), R11 = __dsub_rn(1.0, __dadd_rn(xx, zz)), R21 = __dadd_rn(yz, wx);
  const double R02 = __dadd_rn(xz, wy), R12 = __dsub_rn(yz, wx), R22 = __dsub_rn(1.0, __dadd_rn(xx, yy));
  // index.js:363-364: A = R^T with column k scaled by s_k: A(r,k) = R(k,r) * s_k
  const double A00 = __dmul_rn(R00, sx), A01 = __dmul_rn(R10, sy), A02 = __dmul_rn(R20, sz);
  const double A10 = __dmul_rn(R01, sx), A11 = __dmul_rn(R11, sy), A12 = __dmul_rn(R21, sz);
  const double A20 = __dmul_rn(R02, sx), A21 = __dmul_rn(R12, sy), A22 = __dmul_rn(R22, sz);
  // index.js:365-367: Sigma = A * A^T (multiplyMatrices sums left to right; the 4th terms are exact zeros)
#define SIG(r0, r1, r2, c0, c1, c2) \
  __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(r0, c0), __dmul_rn(r1, c1)), __dmul_rn(r2, c2)), 0.0)
  const double e0 = SIG(A00, A01, A02, A00, A01, A02);   // Sigma(0,0)
  const double e1 = SIG(A10, A11, A12, A00, A01, A02);   // Sigma(1,0)
  const double e2 = SIG(A20, A21, A22, A00, A01, A02);   // Sigma(2,0)
  const double e5 = SIG(A10, A11, A12, A10, A11, A12);   // Sigma(1,1)
  const double e6 = SIG(A20, A21, A22, A10, A11, A12);   // Sigma(2,1)
  const double e10 = SIG(A20, A21, A22, A20, A21, A22);  // Sigma(2,2)
#undef SIG
  // index.js:370-376
  double mx = 0.0;
  if (fabs(e0) > mx) mx = fabs(e0);
  if (fabs(e1) > mx) mx = fabs(e1);
  if (fabs(e2) > mx) mx = fabs(e2);
  if (fabs(e5) > mx) mx = fabs(e5);
  if (fabs(e6) > mx) mx = fabs(e6);
  if (fabs(e10) > mx) mx = fabs(e10);
  // index.js:378-382
  const size_t o = (size_t)first + i;
  cs[o] = make_float4((float)px, (float)py, (float)(-pz), (float)__ddiv_rn(mx, 32767.0));
  // index.js:384-394
  const uint32_t c0 = (uint16_t)parse_int_to_i16(__ddiv_rn(__dmul_rn(e0, 32767.0), mx), tab, nt);
  const uint32_t c1 = (uint16_t)parse_int_to_i16(__ddiv_rn(__dmul_rn(e1, 32767.0), mx), tab, nt);
  const uint32_t c2 = (uint16_t)parse_int_to_i16(__ddiv_rn(__dmul_rn(e2, 32767.0), mx), tab, nt);
  const uint32_t c3 = (uint16_t)parse_int_to_i16(__ddiv_rn(__dmul_rn(e5, 32767.0), mx), tab, nt);
  const uint32_t c4 = (uint16_t)parse_int_to_i16(__ddiv_rn(__dmul_rn(e6, 32767.0), mx), tab, nt);
  const uint32_t c5 = (uint16_t)parse_int_to_i16(__ddiv_rn(__dmul_rn(e10, 32767.0), mx), tab, nt);
  cc[o] = make_uint4(c0 | (c1 << 16), c2 | (c3 << 16), c4 | (c5 << 16), rgba);
  // index.js:396-397: Math.max(scale.x, scale.y, scale.z) * alpha / 255.0  (Math.max returns NaN if any is NaN)
  double ms = sx;
  if (sy > ms) ms = sy;
  if (sz > ms) ms = sz;
  if (isnan(sx) || isnan(sy) || isnan(sz)) ms = nan("");
  sa[o] = (float)__ddiv_rn(__dmul_rn(ms, (double)(rgba >> 24)), 255.0);
}

void launch_pack(gs_context *c, const uint8_t *rows_dev, uint32_t first, uint32_t n, cudaStream_t st) {
  if (!n) return;
  const uint32_t grid = (n + 255) / 256;
  k_pack<<<grid, 256, 0, st>>>((const uint4 *)rows_dev, first, n, c->center_scale, c->cov_color, c->size_alpha,
                                      c->quirk_table, c->quirk_n);
}

}  // namespace gs

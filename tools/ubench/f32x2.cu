// Microbenchmark: issue throughput of scalar FFMA (3-register form) vs packed FFMA2 / FMUL2 / FADD2 on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a --fmad=false -o f32x2 f32x2.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *in, int iters) {
  // 8 independent accumulator chains per thread (enough ILP to hide the 4-cycle latency at 8 warps/SMSP)
  float a0 = in[threadIdx.x], a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = in[1], c = in[2];
  if (MODE == 0) {
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a0 = __fmaf_rn(a0, b, c); a1 = __fmaf_rn(a1, b, c); a2 = __fmaf_rn(a2, b, c); a3 = __fmaf_rn(a3, b, c);
        a4 = __fmaf_rn(a4, b, c); a5 = __fmaf_rn(a5, b, c); a6 = __fmaf_rn(a6, b, c); a7 = __fmaf_rn(a7, b, c);
      }
    }
  } else if (MODE == 1) {
    float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
    float2 q0 = p0, q1 = p1, q2 = p2, q3 = p3;
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        p0 = __ffma2_rn(p0, b2, c2); p1 = __ffma2_rn(p1, b2, c2); p2 = __ffma2_rn(p2, b2, c2); p3 = __ffma2_rn(p3, b2, c2);
        q0 = __ffma2_rn(q0, b2, c2); q1 = __ffma2_rn(q1, b2, c2); q2 = __ffma2_rn(q2, b2, c2); q3 = __ffma2_rn(q3, b2, c2);
      }
    }
    a0 = p0.x + q0.x; a1 = p0.y + q0.y; a2 = p1.x + q1.x; a3 = p1.y + q1.y; a4 = p2.x + q2.x; a5 = p2.y + q2.y; a6 = p3.x + q3.x; a7 = p3.y + q3.y;
  } else if (MODE == 2) {  // mix the raster loop uses: FADD2, FMUL2, FFMA2
    float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        p0 = __fadd2_rn(p0, b2); p1 = __fmul2_rn(p1, c2); p2 = __ffma2_rn(p2, b2, c2); p3 = __fadd2_rn(p3, c2);
        p0 = __fmul2_rn(p0, c2); p1 = __ffma2_rn(p1, b2, p0); p2 = __fadd2_rn(p2, p3); p3 = __fmul2_rn(p3, b2);
      }
    }
    a0 = p0.x; a1 = p0.y; a2 = p1.x; a3 = p1.y; a4 = p2.x; a5 = p2.y; a6 = p3.x; a7 = p3.y;
  } else {  // scalar mix, same op count per element as MODE 2 but one lane-element per instruction
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a0 = __fadd_rn(a0, b); a1 = __fmul_rn(a1, c); a2 = __fmaf_rn(a2, b, c); a3 = __fadd_rn(a3, c);
        a0 = __fmul_rn(a0, c); a1 = __fmaf_rn(a1, b, a0); a2 = __fadd_rn(a2, a3); a3 = __fmul_rn(a3, b);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int MODE>
static void run(const char *name, float *out, float *in, double flop_per_inner) {
  const int iters = 4096, grid = 148 * 8;
  k<MODE><<<grid, 256>>>(out, in, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, in, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double elem_ops = (double)grid * 256 * iters * 8 * flop_per_inner;  // lane-element operations
  printf("%-28s %8.3f ms  %8.2f Tera lane-ops/s   (%s)\n", name, ms, elem_ops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float *out, *in;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0, 4096);
  run<0>("scalar FFMA x8 chains", out, in, 8);
  run<1>("packed FFMA2 x8 pairs", out, in, 16);
  run<3>("scalar FADD/FMUL/FFMA mix", out, in, 8);
  run<2>("packed FADD2/FMUL2/FFMA2 mix", out, in, 16);
  return 0;
}

// FP32 pipe probe for B200 (lab tool, not part of the library): how many warp instructions per
// clock per SM do scalar FFMA/FADD and the two-lane FFMA2/FADD2/FMUL2 sustain, alone and mixed?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float r;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float add1(float a, float b) {
  float r;
  asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float rcp1(float a) {
  float r;
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b) {
  unsigned r;
  asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// MODE: 0 FFMA, 1 FFMA2, 2 FADD, 3 FADD2, 4 FMUL2, 5 FFMA2+FFMA 1:1, 6 FFMA2+2 FFMA, 7 FFMA2+IADD 1:1,
//       8 FFMA+IADD 1:1, 9 the K4 candidate mix (8 packed + 5 scalar + 1 MUFU), 10 2 FFMA2 + 1 FFMA
template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, float seed, float* out) {
  float s[8];
  f32x2 p[8];
  unsigned u[8];
  for (int k = 0; k < 8; ++k) {
    s[k] = seed + k + threadIdx.x;
    p[k] = ((f32x2)__float_as_uint(seed + k) << 32) | __float_as_uint(seed * 0.5f + threadIdx.x);
    u[k] = threadIdx.x + k;
  }
  const float a = 1.0000001f, b = 1e-9f;
  const f32x2 a2 = ((f32x2)__float_as_uint(a) << 32) | __float_as_uint(a);
  const f32x2 b2 = ((f32x2)__float_as_uint(b) << 32) | __float_as_uint(b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) s[k] = fma1(s[k], a, b);
      if (MODE == 1) p[k] = fma2(p[k], a2, b2);
      if (MODE == 2) s[k] = add1(s[k], b);
      if (MODE == 3) p[k] = add2(p[k], b2);
      if (MODE == 4) p[k] = mul2(p[k], a2);
      if (MODE == 5) { p[k] = fma2(p[k], a2, b2); s[k] = fma1(s[k], a, b); }
      if (MODE == 6) { p[k] = fma2(p[k], a2, b2); s[k] = fma1(s[k], a, b); s[(k + 1) & 7] = fma1(s[(k + 1) & 7], a, b); }
      if (MODE == 7) { p[k] = fma2(p[k], a2, b2); u[k] = iadd(u[k], 3u); }
      if (MODE == 8) { s[k] = fma1(s[k], a, b); u[k] = iadd(u[k], 3u); }
      if (MODE == 10) { p[k] = fma2(p[k], a2, b2); p[(k + 4) & 7] = fma2(p[(k + 4) & 7], a2, b2); s[k] = fma1(s[k], a, b); }
      if (MODE == 9) {
        // sub2, mul2, fadd, fmul, fadd, rcp, ffma, ffma, mul2, fma2, fma2, add2 x3
        f32x2 d = add2(p[k], b2);
        f32x2 dd = mul2(d, d);
        float sq = add1(__uint_as_float((unsigned)dd), __uint_as_float((unsigned)(dd >> 32)));
        float den = add1(fma1(sq, sq, 0.f), 1.0f);
        float r = rcp1(den);
        float e = fma1(-den, r, 1.0f);
        r = fma1(r, e, r);
        f32x2 r2 = ((f32x2)__float_as_uint(r) << 32) | __float_as_uint(r);
        f32x2 t = mul2(d, r2);
        f32x2 m = fma2(b2, t, d);
        f32x2 q = fma2(r2, m, t);
        p[k] = add2(p[k], q);
        p[(k + 1) & 7] = add2(p[(k + 1) & 7], d);
        p[(k + 2) & 7] = add2(p[(k + 2) & 7], b2);
      }
    }
  }
  float acc = 0;
  for (int k = 0; k < 8; ++k) acc += s[k] + __uint_as_float((unsigned)p[k]) + __uint_as_float((unsigned)(p[k] >> 32)) + u[k];
  if (acc == 12345.678f) out[0] = acc;
}
template <int MODE>
void run(const char* name, int instr_per_k, double laneops_per_k) {
  const int iters = 4096, blocks = 148 * 8;
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<MODE><<<blocks, 256>>>(64, 1.0f, out);
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 256>>>(iters, 1.0f, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double cycles = ms * 1e-3 * 1.965e9;  // nominal boost; nvidia-smi reports the real one
  double warps = (double)blocks * 8;
  double winstr = warps * iters * 8.0 * instr_per_k;
  printf("{\"mode\": \"%s\", \"ms\": %.4f, \"warp_instr_per_clk_per_sm\": %.3f, \"lane_ops_per_clk_per_sm\": %.1f}\n",
         name, ms, winstr / cycles / 148.0, warps * 32 * iters * 8.0 * laneops_per_k / cycles / 148.0);
  cudaFree(out);
}
int main() {
  run<0>("FFMA", 1, 1);
  run<1>("FFMA2", 1, 2);
  run<2>("FADD", 1, 1);
  run<3>("FADD2", 1, 2);
  run<4>("FMUL2", 1, 2);
  run<5>("FFMA2+FFMA", 2, 3);
  run<6>("FFMA2+2FFMA", 3, 4);
  run<10>("2FFMA2+FFMA", 3, 5);
  run<7>("FFMA2+IADD", 2, 2);
  run<8>("FFMA+IADD", 2, 1);
  run<9>("K4 candidate mix (8 packed, 5 scalar, 1 MUFU)", 14, 21);
  return 0;
}

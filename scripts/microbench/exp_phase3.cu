// Softmax exp phase with PACKED fp32 instructions (Blackwell FFMA2 / FADD2 / FMUL2 on 64-bit register pairs) and the 3-input
// FMNMX3: same work per SM sub-partition as exp_phase2 (2 warps x 128 scores per thread, one query row per thread), variants of the
// instruction mix.  Reports cycles per 256 score-columns per SMSP (the shipped round-1 mix: ~2530; XU floor at x/8 on the FMA pipe:
// 2048 * (1 - x/8)).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_phase3 exp_phase3.cu && ./exp_phase3
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 256

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { __half2 p = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&p); }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = fmaf(0.05520550534129143f, f, 0.24261397123336792f);
  p = fmaf(p, f, 0.6932547688484192f);
  p = fmaf(p, f, 0.9999276995658875f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// two 2^x on the FMA pipe with packed instructions: 2 FMNMX + 6 packed + 2 shift-adds
__device__ __forceinline__ void exp2_poly2(float& x0, float& x1) {
  const u64 x = pk2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const u64 magic = pk2(12582912.0f, 12582912.0f);
  const u64 t = add2(x, magic);
  const u64 n = sub2(t, magic);
  const u64 f = sub2(x, n);
  u64 p = fma2(pk2(0.05520550534129143f, 0.05520550534129143f), f, pk2(0.24261397123336792f, 0.24261397123336792f));
  p = fma2(p, f, pk2(0.6932547688484192f, 0.6932547688484192f));
  p = fma2(p, f, pk2(0.9999276995658875f, 0.9999276995658875f));
  float p0, p1, t0, t1;
  up2(p, p0, p1);
  up2(t, t0, t1);
  x0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  x1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// MODE bit 0: FMNMX3 row max; bit 1: packed scale-sub (FFMA2); bit 2: packed row sum (FADD2); bit 3: packed polynomial
// POLY: of every 8 scores, this many on the FMA pipe.  NOMAX: skip the max pass (sensitivity)
template <int POLY, int MODE, int NOMAX>
__global__ void k_old(float* out, long long* cyc, float seed) {
  constexpr int NS = 128;
  float s[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = seed * (threadIdx.x + i);
  uint32_t pkd[NS / 2];
  float l = 0.f, m = 0.f;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (!NOMAX) {
      float mx0 = -1e30f, mx1 = -1e30f;
      if (MODE & 1) {
#pragma unroll
        for (int i = 0; i < NS; i += 4) { mx0 = max3(mx0, s[i], s[i + 1]); mx1 = max3(mx1, s[i + 2], s[i + 3]); }
      } else {
#pragma unroll
        for (int i = 0; i < NS; i += 2) { mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]); }
      }
      const float mx = fmaxf(mx0, mx1) * 0.18f;
      if (mx > m + 8.0f) m = mx;
    }
    float sum0 = 0.f, sum1 = 0.f;
    u64 sum2 = pk2(0.f, 0.f);
    const u64 sc2 = pk2(0.18f, 0.18f), nm2 = pk2(-m, -m);
#pragma unroll
    for (int q = 0; q < NS / 8; ++q) {
      float pv[8];
      if (MODE & 2) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) up2(fma2(pk2(s[8 * q + i], s[8 * q + i + 1]), sc2, nm2), pv[i], pv[i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) pv[i] = fmaf(s[8 * q + i], 0.18f, -m);
      }
      if (MODE & 8) {
#pragma unroll
        for (int i = 0; i < POLY; i += 2) exp2_poly2(pv[i], pv[i + 1]);
#pragma unroll
        for (int i = POLY; i < 8; ++i) pv[i] = ex2a(pv[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) pv[i] = (i < POLY) ? exp2_poly(pv[i]) : ex2a(pv[i]);
      }
      if (MODE & 4) {
        sum2 = add2(sum2, add2(add2(pk2(pv[0], pv[1]), pk2(pv[2], pv[3])), add2(pk2(pv[4], pv[5]), pk2(pv[6], pv[7]))));
      } else {
        sum0 += (pv[0] + pv[1]) + (pv[2] + pv[3]);
        sum1 += (pv[4] + pv[5]) + (pv[6] + pv[7]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) pkd[q * 4 + i] = pack(pv[2 * i], pv[2 * i + 1]);
    }
    if (MODE & 4) up2(sum2, sum0, sum1);
    l += sum0 + sum1;
#pragma unroll
    for (int i = 0; i < NS / 2; ++i) acc ^= pkd[i];
#pragma unroll
    for (int i = 0; i < NS; i += 16) s[i] += __uint_as_float(acc & 0x3fffff);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}


// ---- second generation: POLY16 of every 16 scores on the FMA pipe; MAXMODE 0 = two scalar FMNMX chains, 1 = FMNMX3, 2 = four scalar chains;
// PP = packed polynomial (pairs); scale-sub and row sum always packed (FFMA2 / FADD2)
template <int POLY16, int MAXMODE, int PP>
__global__ void k2(float* out, long long* cyc, float seed) {
  constexpr int NS = 128;
  float s[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = seed * (threadIdx.x + i);
  uint32_t pkd[NS / 2];
  float l = 0.f, m = 0.f;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    float mx;
    if (MAXMODE == 1) {
      float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
      for (int i = 0; i < NS; i += 4) { mx0 = max3(mx0, s[i], s[i + 1]); mx1 = max3(mx1, s[i + 2], s[i + 3]); }
      mx = fmaxf(mx0, mx1);
    } else if (MAXMODE == 2) {
      float a = -1e30f, b = -1e30f, c = -1e30f, d = -1e30f;
#pragma unroll
      for (int i = 0; i < NS; i += 4) { a = fmaxf(a, s[i]); b = fmaxf(b, s[i + 1]); c = fmaxf(c, s[i + 2]); d = fmaxf(d, s[i + 3]); }
      mx = fmaxf(fmaxf(a, b), fmaxf(c, d));
    } else {
      float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
      for (int i = 0; i < NS; i += 2) { mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]); }
      mx = fmaxf(mx0, mx1);
    }
    mx *= 0.18f;
    if (mx > m + 8.0f) m = mx;
    u64 sum2 = pk2(0.f, 0.f);
    const u64 sc2 = pk2(0.18f, 0.18f), nm2 = pk2(-m, -m);
#pragma unroll
    for (int q = 0; q < NS / 16; ++q) {
      float pv[16];
#pragma unroll
      for (int i = 0; i < 16; i += 2) up2(fma2(pk2(s[16 * q + i], s[16 * q + i + 1]), sc2, nm2), pv[i], pv[i + 1]);
      if (PP) {
#pragma unroll
        for (int i = 0; i + 1 < POLY16; i += 2) exp2_poly2(pv[i], pv[i + 1]);
        if (POLY16 & 1) pv[POLY16 - 1] = exp2_poly(pv[POLY16 - 1]);
      } else {
#pragma unroll
        for (int i = 0; i < POLY16; ++i) pv[i] = exp2_poly(pv[i]);
      }
#pragma unroll
      for (int i = POLY16; i < 16; ++i) pv[i] = ex2a(pv[i]);
      u64 a = add2(add2(pk2(pv[0], pv[1]), pk2(pv[2], pv[3])), add2(pk2(pv[4], pv[5]), pk2(pv[6], pv[7])));
      u64 b = add2(add2(pk2(pv[8], pv[9]), pk2(pv[10], pv[11])), add2(pk2(pv[12], pv[13]), pk2(pv[14], pv[15])));
      sum2 = add2(sum2, add2(a, b));
#pragma unroll
      for (int i = 0; i < 8; ++i) pkd[q * 8 + i] = pack(pv[2 * i], pv[2 * i + 1]);
    }
    float sum0, sum1;
    up2(sum2, sum0, sum1);
    l += sum0 + sum1;
#pragma unroll
    for (int i = 0; i < NS / 2; ++i) acc ^= pkd[i];
#pragma unroll
    for (int i = 0; i < NS; i += 16) s[i] += __uint_as_float(acc & 0x3fffff);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int POLY16, int MAXMODE, int PP>
void run2() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k2<POLY16, MAXMODE, PP><<<148, 256>>>(out, cyc, 1e-3f);
  cudaDeviceSynchronize();
  k2<POLY16, MAXMODE, PP><<<148, 256>>>(out, cyc, 1e-3f);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("packed scale+sum, poly=%2d/16 %s, max=%s : cycles per 256 score-columns per SMSP = %8.1f (XU floor %4d)\n", POLY16, PP ? "packed" : "scalar",
         MAXMODE == 1 ? "FMNMX3    " : MAXMODE == 2 ? "4 chains  " : "2 chains  ", avg / ITERS, 2048 * (16 - POLY16) / 16);
  cudaFree(out); cudaFree(cyc);
}

template <int POLY, int MODE, int NOMAX>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k_old<POLY, MODE, NOMAX><<<148, 256>>>(out, cyc, 1e-3f);
  cudaDeviceSynchronize();
  k_old<POLY, MODE, NOMAX><<<148, 256>>>(out, cyc, 1e-3f);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-58s poly=%d/8 mode=%2d  cycles per 256 score-columns per SMSP = %8.1f\n", name, POLY, MODE, avg / ITERS);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<2, 0, 0>("scalar (round-1 shipped mix)");
  run<2, 7, 0>("FMNMX3 + FFMA2 + FADD2, scalar poly");
  run2<0, 0, 0>(); run2<2, 0, 0>(); run2<3, 0, 0>(); run2<4, 0, 0>(); run2<5, 0, 0>(); run2<6, 0, 0>(); run2<8, 0, 0>();
  run2<4, 0, 1>(); run2<5, 0, 1>(); run2<6, 0, 1>(); run2<8, 0, 1>();
  run2<4, 2, 0>(); run2<5, 2, 0>(); run2<6, 2, 0>(); run2<6, 2, 1>();
  run2<4, 1, 0>(); run2<5, 1, 0>();
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

// Softmax exp phase, same work per SM sub-partition organised two ways: 2 warps x 128 scores per thread (one query row per thread,
// the shipped flash-attention layout) versus 4 warps x 64 scores per thread (each row split between two warps).  A warp's
// instruction stream is issued in order, so MUFU (8 clk), FFMA/FADD/FMNMX of ONE warp add up; only other warps fill the gaps.
// Reports cycles per iteration (= 256 score-columns per SMSP lane) for several MUFU/polynomial mixes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_phase2 exp_phase2.cu && ./exp_phase2
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 256

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { __half2 p = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&p); }
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

// MAXN: how many values the row-max pass covers (the split layout can re-read the whole 128-wide row so both halves agree on m)
template <int NS, int POLY, int MAXN>
__global__ void k(float* out, long long* cyc, float seed, int skew) {
  float s[NS];
  float extra[MAXN > NS ? MAXN - NS : 1];
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = seed * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < (MAXN > NS ? MAXN - NS : 1); ++i) extra[i] = seed * (threadIdx.x + 3 * i);
  uint32_t pk[NS / 2];
  float l = 0.f, m = 0.f;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (skew > 0) {  // de-synchronise the warps of a sub-partition (in the real kernel they never run in lockstep)
    const long long until = t0 + (long long)((threadIdx.x >> 7) * skew);
    while (clock64() < until) {}
  }
  for (int it = 0; it < ITERS; ++it) {
    float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
    for (int i = 0; i < NS; i += 2) { mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]); }
    if (MAXN > NS) {
#pragma unroll
      for (int i = 0; i < MAXN - NS; i += 2) { mx0 = fmaxf(mx0, extra[i]); mx1 = fmaxf(mx1, extra[i + 1]); }
    }
    const float mx = fmaxf(mx0, mx1) * 0.18f;
    if (mx > m + 8.0f) m = mx;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int q = 0; q < NS / 8; ++q) {
      float pv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pv[i] = fmaf(s[8 * q + i], 0.18f, -m);
#pragma unroll
      for (int i = 0; i < 8; ++i) pv[i] = (i < POLY) ? exp2_poly(pv[i]) : ex2a(pv[i]);
      sum0 += (pv[0] + pv[1]) + (pv[2] + pv[3]);
      sum1 += (pv[4] + pv[5]) + (pv[6] + pv[7]);
#pragma unroll
      for (int i = 0; i < 4; ++i) pk[q * 4 + i] = pack(pv[2 * i], pv[2 * i + 1]);
    }
    l += sum0 + sum1;
#pragma unroll
    for (int i = 0; i < NS / 2; ++i) acc ^= pk[i];
#pragma unroll
    for (int i = 0; i < NS; i += 16) s[i] += __uint_as_float(acc & 0x3fffff);
    if (MAXN > NS) extra[0] += __uint_as_float(acc & 0x3fffff);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NS, int POLY, int MAXN>
void run(const char* name, int wps, int skew = 0) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int threads = wps * 128;
  k<NS, POLY, MAXN><<<148, threads>>>(out, cyc, 1e-3f, skew);
  cudaDeviceSynchronize();
  k<NS, POLY, MAXN><<<148, threads>>>(out, cyc, 1e-3f, skew);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-44s warps/SMSP=%d scores/thread=%3d poly=%d/8 skew=%4d  cycles per %d score-columns per SMSP = %8.1f\n", name, wps, NS, POLY, skew, wps * NS, avg / ITERS);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<128, 0, 128>("row per thread", 2);
  run<128, 1, 128>("row per thread", 2);
  run<128, 2, 128>("row per thread (shipped)", 2);
  run<128, 3, 128>("row per thread", 2);
  run<64, 0, 64>("half row per thread, half-row max", 4);
  run<64, 1, 64>("half row per thread, half-row max", 4);
  run<64, 2, 64>("half row per thread, half-row max", 4);
  run<64, 3, 64>("half row per thread, half-row max", 4);
  run<64, 4, 64>("half row per thread, half-row max", 4);
  run<64, 1, 128>("half row per thread, full-row max", 4);
  run<64, 2, 128>("half row per thread, full-row max", 4);
  run<64, 3, 128>("half row per thread, full-row max", 4);
  for (int skew : {200, 500, 900, 1300}) {
    run<128, 2, 128>("row per thread (shipped), skewed start", 2, skew);
    run<128, 0, 128>("row per thread, skewed start", 2, skew);
    run<64, 2, 64>("half row per thread, half-row max, skewed", 4, skew);
    run<64, 2, 128>("half row per thread, full-row max, skewed", 4, skew);
    run<64, 0, 64>("half row per thread, half-row max, skewed", 4, skew);
  }
  run<32, 2, 32>("quarter row per thread", 8);
  run<32, 3, 32>("quarter row per thread", 8);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

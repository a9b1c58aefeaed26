// Pipe-throughput microbenchmark for sm_100a: cycles per warp-instruction per SMSP for the ops the attention softmax uses.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048
#define UNROLL 8

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[UNROLL];
  uint32_t h[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; h[i] = __float_as_uint(a[i]) & 0x3bff3bffu; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) % UNROLL]), "f"(seed));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(a[i]) : "f"(seed));
      if (OP == 4) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) % UNROLL]));
      if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) % UNROLL]));
      if (OP == 6) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 7) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 8) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(h[(i + 1) % UNROLL]));
      if (OP == 9) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) % UNROLL]), "f"(a[(i + 2) % UNROLL]));
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps_per_smsp : {1, 2, 4}) {
    int threads = warps_per_smsp * 4 * 32;
    k<OP><<<148, threads>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    k<OP><<<148, threads>>>(out, cyc, 0.5f);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-28s warps/SMSP=%d  cycles per warp-instr per SMSP = %.3f\n", name, warps_per_smsp, avg / (double(ITERS) * UNROLL * warps_per_smsp));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("MUFU.EX2 f32");
  run<1>("ex2 f16x2");
  run<7>("ex2 bf16x2");
  run<2>("FFMA 3-reg");
  run<3>("FFMA imm");
  run<4>("F2FP pack f16x2");
  run<5>("FMNMX");
  run<9>("FMNMX3");
  run<6>("FADD");
  run<8>("HFMA2");
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}

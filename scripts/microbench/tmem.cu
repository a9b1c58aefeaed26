// TMEM load/store throughput microbenchmark (sm_100a): cycles per tcgen05.ld/st 32x32b.x32 (4 KB per warp) vs warps per SMSP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../video_depth_normal_v2_b200/csrc -o tmem tmem.cu && ./tmem
#include <cstdio>
#include "vdn_common.cuh"
using namespace vdn;

#define ITERS 512

template <int OP>
__global__ void k(float* out, long long* cyc) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t r[32], r2[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { r[i] = threadIdx.x + i; r2[i] = i; }
  tmem_st32(base, r); tmem_st32(base + 32, r); tmem_st_wait();
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < ITERS; ++it) {
    if (OP == 0) { tmem_ld32(base, r); tmem_ld32(base + 32, r2); tmem_ld_wait(); acc ^= r[0] ^ r[17] ^ r2[5] ^ r2[31]; }
    if (OP == 1) { tmem_st32(base, r); tmem_st32(base + 32, r2); tmem_st_wait(); }
    if (OP == 2) { tmem_ld32(base, r); tmem_ld_wait(); acc ^= r[0] ^ r[17]; tmem_ld32(base + 32, r2); tmem_ld_wait(); acc ^= r2[5] ^ r2[31]; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + r[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tptr, 512); }
}

template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps : {1, 4, 8, 16}) {
    k<OP><<<148, warps * 32>>>(out, cyc);
    cudaDeviceSynchronize();
    k<OP><<<148, warps * 32>>>(out, cyc);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double per_op = avg / (ITERS * 2.0);  // one x32 op of one warp = 4 KB
    printf("%-22s warps/CTA=%2d  cycles per x32 op per warp = %7.2f   SM bytes/cycle = %7.1f\n", name, warps, per_op, warps * 4096.0 / per_op);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("LDTM x32 (2 in flight)");
  run<2>("LDTM x32 (serial)");
  run<1>("STTM x32 (2 in flight)");
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}

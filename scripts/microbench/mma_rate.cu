// tcgen05.mma issue/execution rate, cta_group::1, kind::f16, M=128: cycles per MMA (K=16) for N = 64/128/256, A from smem (SS) or TMEM (TS).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../video_depth_normal_v2_b200/csrc -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include "vdn_common.cuh"
using namespace vdn;

#define ITERS 1024

template <int N, int TS>
__global__ void k(long long* cyc, float* sink, int mufu_iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (warp == 0) tmem_alloc(&tptr, 512);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, tptr, 0);
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    constexpr uint32_t idesc = make_idesc(0, 128, N);
    t0 = clock64();
    if (elect_one()) {
      const uint64_t da = make_sdesc_sw128(a_addr), db = make_sdesc_sw128(b_addr);
#pragma unroll 1
      for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (TS) umma_f16_ts(tb, tb + 256 + kk * 8, db + 2 * kk, idesc, 1u);
          else umma_f16(tb, da + 2 * kk, db + 2 * kk, idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  else if (mufu_iters > 0) {
    // background: the other warps keep the XU pipe of their sub-partition saturated (like softmax warps)
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < mufu_iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, int TS>
void run(const char* name, int threads = 128, int mufu_iters = 0) {
  long long* cyc; cudaMalloc(&cyc, 148 * 16);
  float* sink; cudaMalloc(&sink, 148 * 1024 * 4);
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) { k<N, TS><<<148, threads, 64 * 1024>>>(cyc, sink, mufu_iters); cudaDeviceSynchronize(); }
  long long h[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double a = 0, b = 0; for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
  a /= 148; b /= 148;
  printf("%-28s issue %.1f clk/MMA, complete %.1f clk/MMA  -> %.0f FLOP/clk/SM\n", name, a / ITERS, b / ITERS, 2.0 * 128 * N * 16 / (b / ITERS));
  cudaFree(cyc);
}

int main() {
  run<64, 0>("M128 N64  K16 SS");
  run<128, 0>("M128 N128 K16 SS");
  run<256, 0>("M128 N256 K16 SS");
  run<64, 1>("M128 N64  K16 TS");
  run<128, 1>("M128 N128 K16 TS");
  run<256, 1>("M128 N256 K16 TS");
  // same with MUFU-saturating warps: 1 (warps 1-3 on other SMSPs only), 4 and 8 extra warps -> warps 4, 8, .. share SMSP 0 with the issuer
  run<64, 1>("N64 TS + 3 mufu warps", 128, 2000);
  run<64, 1>("N64 TS + 7 mufu warps", 256, 2000);
  run<64, 1>("N64 TS + 15 mufu warps", 512, 2000);
  run<128, 0>("N128 SS + 15 mufu warps", 512, 3000);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}

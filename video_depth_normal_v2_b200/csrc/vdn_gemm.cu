// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   warp 0      : TMA producer (one elected lane) — A and W tiles, 128B-swizzled, multi-stage mbarrier ring
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane), 128 x BLOCK_N x 16 UMMAs, fp32 accumulators in TMEM,
//                 two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2..9  : epilogue — tcgen05.ld TMEM -> registers, fused bias / GELU / ReLU / GEGLU / LayerScale / residual(s) /
//                 row remaps (pixel-shuffle, temporal transpose, patch tokens, QKV split with transposed V) / 1x1 "head" dot,
//                 direct vectorised global stores
//
// Convolution mode: A is an NHWC activation tensor described by a 4-D tensor map (C, W, H, B); a 128-pixel output tile is a
// TH x TW spatial box and each of the 9 taps is the same box shifted by (r-1, s-1) — TMA's out-of-bounds zero fill implements
// the padding, so no im2col buffer ever exists.
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/vdn_b200.h"
#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = 128 B = one swizzle row
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 64 + kNumEpiWarps * 32;
constexpr int kSmemBudget = 227 * 1024 - 2048;

struct GemmKParams {
  int M, N;
  int num_k_blocks;
  int num_m_tiles, num_n_blocks;
  // conv
  int conv, H, W, TH, TW, tiles_h, tiles_w, cin_blocks;
  // epilogue
  const float* bias;
  const float* gamma;
  const void* res;
  int res_f32;
  long long ld_res;
  const void* res2;
  long long ld_res2;
  void* out;
  int out_f32;
  long long ldc;
  void* out2;
  int out2_relu;
  long long ld_out2;
  int act, geglu, row_map, rm0, rm1, rm2, rm3;
  const float* head_w;
  float head_b;
  int fmt;
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStageBytesA = BLOCK_M * BLOCK_K * 2;
  static constexpr int kStageBytesB = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Per-thread output-row context, computed once per tile.
struct RowCtx {
  bool valid;
  long long out_row;  // row index for out / out2 / res2 (and res unless PATCH_TOKENS)
  long long res_row;
  int b, y, x;        // PIXEL_SHUFFLE: image, input row, input col.  QKV_SPLIT: b = frame, x = token
};

__device__ __forceinline__ void store8_f32(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8_16(void* base, long long idx, const float (&v)[8], int fmt) {
  uint4 u;
  u.x = pack16(v[0], v[1], fmt);
  u.y = pack16(v[2], v[3], fmt);
  u.z = pack16(v[4], v[5], fmt);
  u.w = pack16(v[6], v[7], fmt);
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + idx) = u;
}

// Epilogue for one group of 8 consecutive accumulator columns [n, n+8) of one row.
__device__ __forceinline__ void epilogue_group8(const GemmKParams& p, const RowCtx& rc, int n, float (&v)[8]) {
  if (p.bias != nullptr) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (p.geglu) {
    // interleaved (value, gate) pairs -> 4 outputs at column n/2
    uint2 u;
    u.x = pack16(v[0] * gelu_erf(v[1]), v[2] * gelu_erf(v[3]), p.fmt);
    u.y = pack16(v[4] * gelu_erf(v[5]), v[6] * gelu_erf(v[7]), p.fmt);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + rc.out_row * p.ldc + (n >> 1)) = u;
    return;
  }
  if (p.act == VDN_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
  } else if (p.act == VDN_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
  if (p.gamma != nullptr) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + n));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + n + 4));
    v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w;
    v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
  }
  if (p.row_map == VDN_ROWMAP_QKV_SPLIT) {
    const int twoC = 2 * p.rm2;
    if (n < twoC) {
      store8_16(p.out, rc.out_row * p.ldc + n, v, p.fmt);
    } else {
      const int c = n - twoC;  // h*64 + d
      const int heads = p.rm2 >> 6;
      uint16_t* vt = reinterpret_cast<uint16_t*>(p.out2) + ((long long)(rc.b * heads + (c >> 6)) * 64 + (c & 63)) * p.rm1 + rc.x;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t pk = pack16(v[i], 0.0f, p.fmt);
        vt[(long long)i * p.rm1] = static_cast<uint16_t>(pk & 0xFFFFu);
      }
    }
    return;
  }
  long long o_idx;
  if (p.row_map == VDN_ROWMAP_PIXEL_SHUFFLE) {
    const int s = p.rm2, Co = p.rm3;
    const int ij = n / Co, co = n - ij * Co;
    const int i = ij / s, j = ij - i * s;
    o_idx = (((long long)rc.b * (p.rm0 * s) + rc.y * s + i) * (p.rm1 * s) + rc.x * s + j) * p.ldc + co;
  } else {
    o_idx = rc.out_row * p.ldc + n;
  }
  if (p.res != nullptr) {
    if (p.res_f32) {
      const float* r = reinterpret_cast<const float*>(p.res) + rc.res_row * p.ld_res + n;
      const float4 r0 = *reinterpret_cast<const float4*>(r);
      const float4 r1 = *reinterpret_cast<const float4*>(r + 4);
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
      v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
    } else {
      const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.res) + rc.res_row * p.ld_res + n);
      float2 f;
      f = unpack16(u.x, p.fmt); v[0] += f.x; v[1] += f.y;
      f = unpack16(u.y, p.fmt); v[2] += f.x; v[3] += f.y;
      f = unpack16(u.z, p.fmt); v[4] += f.x; v[5] += f.y;
      f = unpack16(u.w, p.fmt); v[6] += f.x; v[7] += f.y;
    }
  }
  if (p.res2 != nullptr) {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.res2) + rc.out_row * p.ld_res2 + n);
    float2 f;
    f = unpack16(u.x, p.fmt); v[0] += f.x; v[1] += f.y;
    f = unpack16(u.y, p.fmt); v[2] += f.x; v[3] += f.y;
    f = unpack16(u.z, p.fmt); v[4] += f.x; v[5] += f.y;
    f = unpack16(u.w, p.fmt); v[6] += f.x; v[7] += f.y;
  }
  if (p.out_f32) {
    store8_f32(reinterpret_cast<float*>(p.out) + o_idx, v);
  } else {
    store8_16(p.out, o_idx, v, p.fmt);
  }
  if (p.out2 != nullptr) {
    if (p.out2_relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
    store8_16(p.out2, rc.out_row * p.ld_out2 + n, v, p.fmt);
  }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kStageBytesA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_blocks;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kNumEpiWarps * 32);
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.num_n_blocks;
        const int m_tile = tile / p.num_n_blocks;
        int img = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          const int tw_i = m_tile % p.tiles_w;
          const int t2 = m_tile / p.tiles_w;
          const int th_i = t2 % p.tiles_h;
          img = t2 / p.tiles_h;
          h0 = th_i * p.TH;
          w0 = tw_i * p.TW;
        }
        for (int k = 0; k < p.num_k_blocks; ++k) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (p.conv) {
            const int tap = k / p.cin_blocks;
            const int kc = k - tap * p.cin_blocks;
            const int r = tap / 3, s = tap - r * 3;
            tma_load_4d(smem_a + stage * Cfg::kStageBytesA, &tmA, &full_bar[stage], kc * BLOCK_K, w0 + s - 1, h0 + r - 1, img);
          } else {
            tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tmA, &full_bar[stage], k * BLOCK_K, m_tile * BLOCK_M);
          }
          tma_load_2d(smem_b + stage * Cfg::kStageBytesB, &tmB, &full_bar[stage], k * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.fmt ? 1u : 0u, BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int k = 0; k < p.num_k_blocks; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sdesc_sw128(smem_u32(smem_a + stage * Cfg::kStageBytesA));
          const uint64_t db = make_sdesc_sw128(smem_u32(smem_b + stage * Cfg::kStageBytesB));
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / 16; ++kk) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle row: +2 in the (addr >> 4) field
            umma_f16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int e = warp_idx - 2;
    const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
    const int half = e >> 2;           // column half handled by this warp
    constexpr int kChunks = BLOCK_N / 32;
    constexpr int kChunksPerHalf = (kChunks + 1) / 2;
    const int row_in_tile = quarter * 32 + lane;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int n_blk = tile % p.num_n_blocks;
      const int m_tile = tile / p.num_n_blocks;
      // ---- row context ----
      RowCtx rc;
      long long row;
      if (p.conv) {
        const int tw_i = m_tile % p.tiles_w;
        const int t2 = m_tile / p.tiles_w;
        const int th_i = t2 % p.tiles_h;
        const int img = t2 / p.tiles_h;
        const int hh = th_i * p.TH + row_in_tile / p.TW;
        const int ww = tw_i * p.TW + row_in_tile % p.TW;
        rc.valid = (hh < p.H) && (ww < p.W);
        row = ((long long)img * p.H + hh) * p.W + ww;
      } else {
        row = (long long)m_tile * BLOCK_M + row_in_tile;
        rc.valid = row < p.M;
      }
      rc.out_row = row;
      rc.res_row = row;
      rc.b = rc.y = rc.x = 0;
      if (p.row_map == VDN_ROWMAP_PIXEL_SHUFFLE) {
        const int hw = p.rm0 * p.rm1;
        rc.b = int(row / hw);
        const int rem = int(row - (long long)rc.b * hw);
        rc.y = rem / p.rm1;
        rc.x = rem - rc.y * p.rm1;
      } else if (p.row_map == VDN_ROWMAP_TEMPORAL) {
        const int T = p.rm0, D = p.rm1;
        const long long bd = row / T;
        const int f = int(row - bd * T);
        const long long bb = bd / D;
        const int d = int(bd - bb * D);
        rc.out_row = (bb * T + f) * D + d;
        rc.res_row = rc.out_row;
      } else if (p.row_map == VDN_ROWMAP_PATCH_TOKENS) {
        const int P = p.rm0;
        const long long bb = row / P;
        const int pp = int(row - bb * P);
        rc.out_row = bb * (P + 1) + 1 + pp;
        rc.res_row = 1 + pp;
      } else if (p.row_map == VDN_ROWMAP_QKV_SPLIT) {
        rc.b = int(row / p.rm0);
        rc.x = int(row - (long long)rc.b * p.rm0);
      }

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c = half * kChunksPerHalf; c < kChunks && c < (half + 1) * kChunksPerHalf; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        const int n0 = n_blk * BLOCK_N + c * 32;
        if (rc.valid) {
          if (p.head_w != nullptr) {
            // fused output_conv2: ReLU(conv3x3) -> 1x1 conv -> ReLU  (dpt.py:118-124), N <= 32 so one chunk per row
            float s = p.head_b;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < p.N) {
                const float a = fmaxf(__uint_as_float(r[j]) + __ldg(p.bias + j), 0.0f);
                s = fmaf(a, __ldg(p.head_w + j), s);
              }
            }
            reinterpret_cast<float*>(p.out)[rc.out_row] = fmaxf(s, 0.0f);
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int n = n0 + g * 8;
              if (n < p.N) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
                epilogue_group8(p, rc, n, v);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(gemm): ") + cudaGetErrorString(e));
    configured = true;
  }
  const int num_tiles = p.num_m_tiles * p.num_n_blocks;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  gemm_tc_kernel<BLOCK_N><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  count_launch();
  return check_launch("gemm_tc_kernel");
}

static void pick_spatial_tile(int H, int W, int* th, int* tw) {
  const int cand[][2] = {{8, 16}, {16, 8}, {4, 32}, {32, 4}, {2, 64}, {64, 2}, {1, 128}, {128, 1}};
  long long best = -1;
  for (auto& c : cand) {
    const long long tiles = (long long)((H + c[0] - 1) / c[0]) * ((W + c[1] - 1) / c[1]);
    if (best < 0 || tiles < best) {
      best = tiles;
      *th = c[0];
      *tw = c[1];
    }
  }
}

}  // namespace vdn

using namespace vdn;

extern "C" int vdn_gemm(const vdn_gemm_desc* d, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (d == nullptr || d->a == nullptr || d->w == nullptr || d->out == nullptr) return set_error("vdn_gemm: null pointer");
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return set_error("vdn_gemm: non-positive dimension");
  if (d->N % 8 != 0) return set_error("vdn_gemm: N must be a multiple of 8");
  if ((d->ldw * 2) % 16 != 0) return set_error("vdn_gemm: ldw*2 must be a multiple of 16 bytes");
  if (reinterpret_cast<uintptr_t>(d->a) % 16 || reinterpret_cast<uintptr_t>(d->w) % 16 || reinterpret_cast<uintptr_t>(d->out) % 16)
    return set_error("vdn_gemm: a / w / out must be 16-byte aligned");
  if (d->head_w != nullptr && (d->N > 32 || d->bias == nullptr || !d->out_f32)) return set_error("vdn_gemm: head mode needs N<=32, bias and fp32 out");
  if (d->geglu && (d->out_f32 || d->res || d->row_map != VDN_ROWMAP_IDENTITY)) return set_error("vdn_gemm: geglu supports plain 16-bit output only");
  const int fmt = get_operand_format();

  GemmKParams p{};
  p.M = (int)d->M;
  p.N = (int)d->N;
  p.bias = d->bias; p.gamma = d->gamma;
  p.res = d->res; p.res_f32 = d->res_f32; p.ld_res = d->ld_res;
  p.res2 = d->res2; p.ld_res2 = d->ld_res2;
  p.out = d->out; p.out_f32 = d->out_f32; p.ldc = d->ldc;
  p.out2 = d->out2; p.out2_relu = d->out2_relu; p.ld_out2 = d->ld_out2;
  p.act = d->act; p.geglu = d->geglu;
  p.row_map = d->row_map; p.rm0 = d->rm0; p.rm1 = d->rm1; p.rm2 = d->rm2; p.rm3 = d->rm3;
  p.head_w = d->head_w; p.head_b = d->head_b;
  p.fmt = fmt;
  p.conv = d->conv;

  // BLOCK_N: largest tile that does not waste more than necessary
  int block_n;
  if (d->N > 128) block_n = 256;
  else if (d->N > 64) block_n = 128;
  else if (d->N > 32) block_n = 64;
  else block_n = 32;
  // prefer 128-wide tiles when 256 would leave most of the last tile empty (e.g. N = 384)
  if (block_n == 256 && (d->N % 256) != 0 && (d->N % 256) <= 128) block_n = 128;
  p.num_n_blocks = (int)((d->N + block_n - 1) / block_n);

  CUtensorMap tmA, tmB;
  if (d->conv) {
    if (d->M != (int64_t)d->B * d->H * d->W) return set_error("vdn_gemm(conv): M != B*H*W");
    if ((d->K * 2) % 16 != 0) return set_error("vdn_gemm(conv): C_in*2 must be a multiple of 16 bytes");
    pick_spatial_tile(d->H, d->W, &p.TH, &p.TW);
    p.H = d->H; p.W = d->W;
    p.tiles_h = (d->H + p.TH - 1) / p.TH;
    p.tiles_w = (d->W + p.TW - 1) / p.TW;
    p.num_m_tiles = p.tiles_h * p.tiles_w * d->B;
    p.cin_blocks = (int)((d->K + BLOCK_K - 1) / BLOCK_K);
    p.num_k_blocks = 9 * p.cin_blocks;
    if (d->ldw < (int64_t)p.num_k_blocks * BLOCK_K) return set_error("vdn_gemm(conv): ldw < 9*roundup(C_in,64)");
    const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->K * 2, (uint64_t)d->K * 2 * d->W, (uint64_t)d->K * 2 * d->W * d->H};
    const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)p.TW, (uint32_t)p.TH, 1u};
    if (make_tensor_map(&tmA, d->a, fmt, 4, dims, strides, box)) return 1;
  } else {
    if ((d->lda * 2) % 16 != 0) return set_error("vdn_gemm: lda*2 must be a multiple of 16 bytes");
    p.num_m_tiles = (int)((d->M + BLOCK_M - 1) / BLOCK_M);
    p.num_k_blocks = (int)((d->K + BLOCK_K - 1) / BLOCK_K);
    const uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->M};
    const uint64_t strides[1] = {(uint64_t)d->lda * 2};
    const uint32_t box[2] = {(uint32_t)BLOCK_K, (uint32_t)BLOCK_M};
    if (make_tensor_map(&tmA, d->a, fmt, 2, dims, strides, box)) return 1;
  }
  {
    const uint64_t kw = d->conv ? (uint64_t)p.num_k_blocks * BLOCK_K : (uint64_t)d->K;
    const uint64_t dims[2] = {kw, (uint64_t)d->N};
    const uint64_t strides[1] = {(uint64_t)d->ldw * 2};
    const uint32_t box[2] = {(uint32_t)BLOCK_K, (uint32_t)block_n};
    if (make_tensor_map(&tmB, d->w, fmt, 2, dims, strides, box)) return 1;
  }
  switch (block_n) {
    case 256: return launch_gemm<256>(tmA, tmB, p, stream);
    case 128: return launch_gemm<128>(tmA, tmB, p, stream);
    case 64: return launch_gemm<64>(tmA, tmB, p, stream);
    default: return launch_gemm<32>(tmA, tmB, p, stream);
  }
}

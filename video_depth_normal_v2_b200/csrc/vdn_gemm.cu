// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   warp 0      : TMA producer (one elected lane) — A and W tiles, 128B-swizzled, multi-stage mbarrier ring
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane), 128 x BLOCK_N x 16 UMMAs, fp32 accumulators in TMEM,
//                 two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2..9  : epilogue — tcgen05.ld TMEM -> registers, fused bias / GELU / ReLU / GEGLU / LayerScale / residual(s) /
//                 row remaps (pixel-shuffle, temporal transpose, patch tokens, QKV split with transposed V) / 1x1 "head" dot,
//                 direct vectorised global stores.  The kernel is templated on <BLOCK_N, epilogue mode, operand format> so each
//                 instantiation carries one epilogue body (the all-modes-in-one version was instruction-fetch bound: ncu showed
//                 28 % stall_no_inst); the chunk loop is rolled and software-pipelined: the next accumulator chunk's tcgen05.ld
//                 and the fp32 residual two chunks ahead are in flight while the current chunk's math and stores run.
//
// Convolution mode: A is an NHWC activation tensor described by a 4-D tensor map (C, W, H, B); a 128-pixel output tile is a
// TH x TW spatial box and each of the 9 taps is the same box shifted by (r-1, s-1) — TMA's out-of-bounds zero fill implements
// the padding, so no im2col buffer ever exists.
#include <stdlib.h>

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
// Three warpgroups: control (warp 0 TMA producer, warp 1 MMA issuer, warps 2-3 idle) and two of epilogue warps.  Registers are granted
// per warpgroup (a 320-thread block is budgeted as 384: 168 registers each, and __maxnreg__(192) does not launch), so the control
// warpgroup hands most of its share to the epilogue with setmaxnreg: 128 x 56 + 256 x 224 = 64512 registers (no spills left, room for
// two chunks of look-ahead on both residual streams).
constexpr int kNumThreads = 128 + kNumEpiWarps * 32;
constexpr int kSmemBudget = 227 * 1024 - 256 /*barriers*/ - 2 * 256 * 4 /*bias+gamma*/;

// EPI_PLAIN: bias / act / out (+out2), no residual operands (keeps registers free so the GELU chains interleave);
// EPI_RES  : additionally LayerScale gamma + residual(s), with register prefetch buffers.
// EPI_TMA  : bias / act / gamma, identity row map; the tile leaves through swizzled shared-memory staging and TMA: a plain tensor
//            store, or an fp32 reduce-add when the output is accumulated in place (x += gamma * (A W^T + b)) — the old value is
//            never loaded into the SM.  Row-per-thread global accesses (32 rows x 16 B per warp instruction) cost 32 LSU
//            wavefronts each and made the epilogue, not the MMA, the bound of the K=1024 residual GEMMs (tensor pipe 29 % busy).
enum { EPI_PLAIN = 0, EPI_QKV = 1, EPI_GEGLU = 2, EPI_PIXSHUF = 3, EPI_HEAD = 4, EPI_RES = 5, EPI_TMA = 6 };
constexpr int kStagingBytesPerWarp = 4096;  // one 32 x 32 fp32 chunk, or two 32 x 32 16-bit chunks

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
  int act, row_map, rm0, rm1, rm2, rm3;
  const float* head_w;
  float head_b;
  int qkv_split, qkv_tpo, qkv_toff;
  int qkv_tma;  // the q|k columns leave through smem staging + TMA store (identity row placement only)
};

// HALO convolution (3x3, spatial tile 16 rows x 8 pixels): per 64-channel chunk the A operand is three column-shifted haloed
// tiles [18 rows x 8 px x 64 ch] (one per horizontal tap dx; each image row is one 1024-byte swizzle atom), and the nine taps are
// MMA descriptors into them — tap (dy, dx) starts dy atoms into tile dx.  54 KB of smem fill per chunk instead of 9 x 16 KB:
// the small-N convolutions (output_conv1/2) were bound by bytes in flight, not by the tensor pipe.
constexpr int kHaloTH = 16, kHaloTW = 8;
constexpr int kHaloBoxBytes = (kHaloTH + 2) * kHaloTW * BLOCK_K * 2;  // 18432
constexpr int kHaloABytes = 3 * kHaloBoxBytes;                        // 55296 per A stage
constexpr int kHaloAStages = 2;

template <int BLOCK_N, int EPI = EPI_PLAIN, int HALO = 0>
struct GemmCfg {
  static constexpr int kStageBytesA = BLOCK_M * BLOCK_K * 2;
  static constexpr int kStageBytesB = (HALO == 3 ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;  // pair form: each CTA stages half of the B tile
  static constexpr int kStageBytes = HALO == 1 ? kStageBytesB : kStageBytesA + kStageBytesB;  // HALO: the ring holds B (per-tap) tiles only
  static constexpr int kStagingBytes = (EPI == EPI_TMA || EPI == EPI_QKV) ? kNumEpiWarps * kStagingBytesPerWarp : 0;
  static constexpr int kARingBytes = HALO == 1 ? kHaloAStages * kHaloABytes : 0;
  static constexpr int kStagesRaw = (kSmemBudget - kStagingBytes - kARingBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kParamBytes = 2 * BLOCK_N * 4;  // this tile's bias and gamma slices
  static constexpr int kSmemBytes = kARingBytes + kStages * kStageBytes + kStagingBytes + 256 /*barriers*/ + kParamBytes;
};

// Per-thread output-row context, computed once per tile.
struct RowCtx {
  bool valid;
  long long out_row;  // row index for out / out2 / res2 (and res unless PATCH_TOKENS)
  long long res_row;
  int b, y, x;        // PIXEL_SHUFFLE: image, input row, input col.  QKV_SPLIT: b = frame, x = token
};

template <int FMT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (FMT == 1) return T16<__nv_bfloat16>::pack(a, b);
  else return T16<__half>::pack(a, b);
}
template <int FMT>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  if constexpr (FMT == 1) return T16<__nv_bfloat16>::unpack(u);
  else return T16<__half>::unpack(u);
}
__device__ __forceinline__ void store8_f32(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <int FMT>
__device__ __forceinline__ void store8_16(void* base, long long idx, const float* v) {
  uint4 u;
  u.x = pack2<FMT>(v[0], v[1]);
  u.y = pack2<FMT>(v[2], v[3]);
  u.z = pack2<FMT>(v[4], v[5]);
  u.w = pack2<FMT>(v[6], v[7]);
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + idx) = u;
}
template <int FMT>
__device__ __forceinline__ void add16x8(float* v, const uint4& u) {
  float2 f;
  f = unpack2<FMT>(u.x); v[0] += f.x; v[1] += f.y;
  f = unpack2<FMT>(u.y); v[2] += f.x; v[3] += f.y;
  f = unpack2<FMT>(u.z); v[4] += f.x; v[5] += f.y;
  f = unpack2<FMT>(u.w); v[6] += f.x; v[7] += f.y;
}

// Raw residual vectors of one 32-column chunk of one row (fp32: 8 x 16 B; 16-bit: r[0..3]).
struct ResBuf {
  uint4 r[8];
};

__device__ __forceinline__ void prefetch_res(const GemmKParams& p, const RowCtx& rc, int n0, ResBuf& rb) {
  if (p.res == nullptr || !rc.valid || n0 >= p.N) return;
  if (p.res_f32) {
    const float* r = reinterpret_cast<const float*>(p.res) + rc.res_row * p.ld_res + n0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (n0 + i * 4 < p.N) rb.r[i] = *reinterpret_cast<const uint4*>(r + i * 4);
  } else {
    const uint16_t* r = reinterpret_cast<const uint16_t*>(p.res) + rc.res_row * p.ld_res + n0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n0 + i * 8 < p.N) rb.r[i] = *reinterpret_cast<const uint4*>(r + i * 8);
  }
}
__device__ __forceinline__ void prefetch_res2(const GemmKParams& p, const RowCtx& rc, int n0, uint4 (&r2)[4]) {
  if (p.res2 == nullptr || !rc.valid || n0 >= p.N) return;
  const uint16_t* r = reinterpret_cast<const uint16_t*>(p.res2) + rc.out_row * p.ld_res2 + n0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (n0 + i * 8 < p.N) r2[i] = *reinterpret_cast<const uint4*>(r + i * 8);
}

// ---- epilogue bodies: v[32] already holds acc (+bias) for columns [n0, n0+32) of this thread's row ----------------------------
template <int FMT>
__device__ __forceinline__ void store_plain(const GemmKParams& p, const RowCtx& rc, int n0, float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = n0 + g * 8;
    if (n < p.N) {
      const long long o_idx = rc.out_row * p.ldc + n;
      if (p.out_f32) store8_f32(reinterpret_cast<float*>(p.out) + o_idx, v + g * 8);
      else store8_16<FMT>(p.out, o_idx, v + g * 8);
    }
  }
  if (p.out2 != nullptr) {
    if (p.out2_relu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (n0 + g * 8 < p.N) store8_16<FMT>(p.out2, rc.out_row * p.ld_out2 + n0 + g * 8, v + g * 8);
  }
}

template <int FMT>
__device__ __forceinline__ void epi_noresidual(const GemmKParams& p, const RowCtx& rc, int n0, float (&v)[32]) {
  if (p.act == VDN_ACT_GELU) {
    gelu_fast_batch<8>(v);
    gelu_fast_batch<8>(v + 8);
    gelu_fast_batch<8>(v + 16);
    gelu_fast_batch<8>(v + 24);
  } else if (p.act == VDN_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
  store_plain<FMT>(p, rc, n0, v);
}

template <int FMT>
__device__ __forceinline__ void epi_plain(const GemmKParams& p, const RowCtx& rc, int n0, int col_in_tile, float (&v)[32], ResBuf& rb,
                                          uint4 (&r2)[4], int n_res_next, int n_res2_next, const float* sgamma) {
  if (p.act == VDN_ACT_GELU) {
    gelu_fast_batch<8>(v);
    gelu_fast_batch<8>(v + 8);
    gelu_fast_batch<8>(v + 16);
    gelu_fast_batch<8>(v + 24);
  } else if (p.act == VDN_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
  if (p.gamma != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= sgamma[col_in_tile + i];
  }
  if (p.res != nullptr) {
    if (p.res_f32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (n0 + i * 4 < p.N) {
          v[4 * i] += __uint_as_float(rb.r[i].x); v[4 * i + 1] += __uint_as_float(rb.r[i].y);
          v[4 * i + 2] += __uint_as_float(rb.r[i].z); v[4 * i + 3] += __uint_as_float(rb.r[i].w);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n0 + i * 8 < p.N) add16x8<FMT>(v + 8 * i, rb.r[i]);
    }
    if (n_res_next >= 0) prefetch_res(p, rc, n_res_next, rb);  // two chunks ahead, different columns than the stores below
  }
  if (p.res2 != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n0 + i * 8 < p.N) add16x8<FMT>(v + 8 * i, r2[i]);
    if (n_res2_next >= 0) prefetch_res2(p, rc, n_res2_next, r2);
  }
  store_plain<FMT>(p, rc, n0, v);
}

template <int FMT>
__device__ __forceinline__ void epi_qkv(const GemmKParams& p, const RowCtx& rc, int n0, float (&v)[32]) {
  const int twoC = p.qkv_split;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = n0 + g * 8;
    if (n >= p.N) break;
    if (n < twoC) {
      store8_16<FMT>(p.out, rc.out_row * p.ldc + n, v + g * 8);
    } else {
      const int c = n - twoC;  // h*64 + d
      const int heads = (p.N - twoC) >> 6;
      uint16_t* vt = reinterpret_cast<uint16_t*>(p.out2) + ((long long)(rc.b * heads + (c >> 6)) * 64 + (c & 63)) * p.rm1 + rc.x;
#pragma unroll
      for (int i = 0; i < 8; ++i) vt[(long long)i * p.rm1] = static_cast<uint16_t>(pack2<FMT>(v[g * 8 + i], 0.0f) & 0xFFFFu);
    }
  }
}

template <int FMT>
__device__ __forceinline__ void epi_geglu(const GemmKParams& p, const RowCtx& rc, int n0, float (&v)[32]) {
  // interleaved (value, gate) pairs -> 16 outputs at column n0/2; gate activation GELU (GEGLU, motion_module/attention.py:382-384)
  // or SiLU (SwiGLU, dinov2_layers/swiglu_ffn.py:29-33)
  uint32_t o[8];
  float gt[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) gt[i] = v[2 * i + 1];
  if (p.act == VDN_ACT_SILU) silu_fast_batch<16>(gt);
  else gelu_fast_batch<16>(gt);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = pack2<FMT>(v[4 * i] * gt[2 * i], v[4 * i + 2] * gt[2 * i + 1]);
  uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + rc.out_row * p.ldc + (n0 >> 1);
  if (n0 + 16 <= p.N) *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
  else if (n0 + 8 <= p.N) *reinterpret_cast<uint2*>(dst) = make_uint2(o[0], o[1]);
  if (n0 + 32 <= p.N) *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
  else if (n0 + 24 <= p.N) *reinterpret_cast<uint2*>(dst + 8) = make_uint2(o[4], o[5]);
}

template <int FMT>
__device__ __forceinline__ void epi_pixshuf(const GemmKParams& p, const RowCtx& rc, int n0, float (&v)[32]) {
  const int s = p.rm2, Co = p.rm3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = n0 + g * 8;
    if (n >= p.N) break;
    const int ij = n / Co, co = n - ij * Co;
    const int i = ij / s, j = ij - i * s;
    const long long o_idx = (((long long)rc.b * (p.rm0 * s) + rc.y * s + i) * (p.rm1 * s) + rc.x * s + j) * p.ldc + co;
    store8_16<FMT>(p.out, o_idx, v + g * 8);
  }
}

__device__ __forceinline__ void epi_head(const GemmKParams& p, const RowCtx& rc, float (&v)[32]) {
  // fused output_conv2: ReLU(conv3x3) -> 1x1 conv -> ReLU  (dpt.py:118-124); N <= 32 so one chunk per row
  float s = p.head_b;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < p.N) s = fmaf(fmaxf(v[j], 0.0f), __ldg(p.head_w + j), s);
  reinterpret_cast<float*>(p.out)[rc.out_row] = fmaxf(s, 0.0f);
}

template <int BLOCK_N, int EPI, int FMT, int HALO = 0>
__global__ void __launch_bounds__(kNumThreads, 1)  // registers are granted per warpgroup: 320 threads count as 384, i.e. 168 registers per thread (__maxnreg__(192) does not launch)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmBh, const GemmKParams p) {
  // HALO == 2: plain GEMM on clusters of two CTAs that work on two M tiles of the same N block in lockstep; each CTA loads half of
  // the B (weight) tile and multicasts it to both (tmBh: box of BLOCK_N / 2 rows), halving the L2 -> SM traffic of B.
  // HALO == 3: the pair form - ONE tcgen05.mma.cta_group::2 of M = 256 per K step, issued by the leader CTA: each CTA stages its
  // own A tile and HALF of the B tile (no duplication in shared memory), the peer's loads signal the leader's full barrier, the
  // leader's commits release the stages and publish the accumulators in both CTAs, both epilogues report to the leader.
  constexpr bool CL = HALO == 2 || HALO == 3;
  constexpr bool PAIR = HALO == 3;
  using Cfg = GemmCfg<BLOCK_N, EPI, HALO>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (HALO == 1 ? Cfg::kARingBytes : kStages * Cfg::kStageBytesA);
  constexpr int kRingBytes = Cfg::kARingBytes + kStages * Cfg::kStageBytes;
  uint8_t* smem_stg = smem + kRingBytes;  // EPI_TMA staging (1024-byte aligned: stage sizes are multiples of 1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRingBytes + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* a_full = bars + 2 * kStages + 5;   // HALO: [2]
  uint64_t* a_empty = bars + 2 * kStages + 7;  // HALO: [2]
  float* sbias = reinterpret_cast<float*>(smem + kRingBytes + Cfg::kStagingBytes + 256);
  float* sgamma = sbias + BLOCK_N;

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_blocks;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (EPI == EPI_TMA || EPI == EPI_QKV) tma_prefetch_desc(&tmC);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], (CL && !PAIR) ? 2 : 1);  // cluster: a stage is free once the MMAs of both CTAs have read it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], (PAIR ? 2 : 1) * kNumEpiWarps * 32);  // pair: the leader hears from both epilogues
      if constexpr (HALO == 1) {
        mbar_init(&a_full[i], 1);
        mbar_init(&a_empty[i], 1);
      }
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    if constexpr (PAIR) tmem_alloc_2sm(tmem_ptr_smem, Cfg::kTmemCols);
    else tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
  }
  tc_fence_before();
  if constexpr (CL) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast into it
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // work iteration: whole tiles, or (cluster) pairs of M tiles of one N block, this CTA taking the tile of its rank
  const int cta_rank = CL ? (int)cluster_ctarank() : 0;
  const int it0 = CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int it_step = CL ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int it_end = CL ? ((p.num_m_tiles + 1) / 2) * p.num_n_blocks : num_tiles;

  if (warp_idx < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");  // inside the branch it governs, or ptxas budgets every role for the minimum
  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    // Producer and issuer warps run their (warp-uniform) loops with all lanes and elect one lane per issue: under a plain
    // `if (lane == 0)` the compiler cannot prove descriptors / coordinates uniform and wraps every tcgen05.mma and TMA in an
    // R2UR + ELECT + BRA.U.ANY waterfall (~80 cycles per MMA, longer than a 128x64x16 MMA itself).
    if constexpr (HALO == 1) {
      int stage = 0, astage = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.num_n_blocks;
        const int m_tile = tile / p.num_n_blocks;
        const int tw_i = m_tile % p.tiles_w;
        const int t2 = m_tile / p.tiles_w;
        const int th_i = t2 % p.tiles_h;
        const int img = t2 / p.tiles_h;
        const int h0 = th_i * kHaloTH, w0 = tw_i * kHaloTW;
        for (int kc = 0; kc < p.cin_blocks; ++kc) {
          mbar_wait(&a_empty[astage], aphase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[astage], kHaloABytes);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              tma_load_4d(smem_a + astage * kHaloABytes + dx * kHaloBoxBytes, &tmA, &a_full[astage], kc * BLOCK_K, w0 + dx - 1, h0 - 1, img);
          }
          __syncwarp();
          if (++astage == kHaloAStages) {
            astage = 0;
            aphase ^= 1;
          }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytesB);
              tma_load_2d(smem_b + stage * Cfg::kStageBytesB, &tmB, &full_bar[stage], (tap * p.cin_blocks + kc) * BLOCK_K, n_blk * BLOCK_N);
            }
            __syncwarp();
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else {
    int stage = 0;
    uint32_t phase = 0;
    for (int it = it0; it < it_end; it += it_step) {
      const int n_blk = it % p.num_n_blocks;
      const int m_tile = CL ? 2 * (it / p.num_n_blocks) + cta_rank : it / p.num_n_blocks;
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
        if constexpr (PAIR) {
          if (elect_one()) {
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);  // both CTAs' A tile + B half
            const uint32_t lead_bar = mapa_rank(smem_u32(&full_bar[stage]), 0);
            if (p.conv) {  // this CTA's spatial tile, shifted by the tap (OOB rows / columns / images are zero-filled)
              const int tap = k / p.cin_blocks;
              const int kc = k - tap * p.cin_blocks;
              const int r = tap / 3, s_ = tap - r * 3;
              tma_load_4d_2sm(smem_a + stage * Cfg::kStageBytesA, &tmA, lead_bar, kc * BLOCK_K, w0 + s_ - 1, h0 + r - 1, img);
            } else {
              tma_load_2d_2sm(smem_a + stage * Cfg::kStageBytesA, &tmA, lead_bar, k * BLOCK_K, m_tile * BLOCK_M);
            }
            tma_load_2d_2sm(smem_b + stage * Cfg::kStageBytesB, &tmBh, lead_bar, k * BLOCK_K, n_blk * BLOCK_N + cta_rank * (BLOCK_N / 2));
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (p.conv) {
            const int tap = k / p.cin_blocks;
            const int kc = k - tap * p.cin_blocks;
            const int r = tap / 3, s = tap - r * 3;
            tma_load_4d(smem_a + stage * Cfg::kStageBytesA, &tmA, &full_bar[stage], kc * BLOCK_K, w0 + s - 1, h0 + r - 1, img);
          } else {
            tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tmA, &full_bar[stage], k * BLOCK_K, m_tile * BLOCK_M);
          }
          if constexpr (CL)
            tma_load_2d_mc(smem_b + stage * Cfg::kStageBytesB + cta_rank * (Cfg::kStageBytesB / 2), &tmBh, &full_bar[stage], k * BLOCK_K,
                           n_blk * BLOCK_N + cta_rank * (BLOCK_N / 2), (uint16_t)0b11);
          else
            tma_load_2d(smem_b + stage * Cfg::kStageBytesB, &tmB, &full_bar[stage], k * BLOCK_K, n_blk * BLOCK_N);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(FMT ? 1u : 0u, BLOCK_M, BLOCK_N);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a_addr = smem_u32(smem_a), b_addr = smem_u32(smem_b);
    if constexpr (HALO == 1) {
      int stage = 0, astage = 0;
      uint32_t phase = 0, aphase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tb + acc * BLOCK_N;
        for (int kc = 0; kc < p.cin_blocks; ++kc) {
          mbar_wait(&a_full[astage], aphase);
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const int dy = tap / 3, dx = tap - dy * 3;
              // output pixel (y, x) of the 16 x 8 tile reads haloed row y + dy of the tile loaded at column offset dx
              const uint64_t da = make_sdesc_sw128(a_addr + astage * kHaloABytes + dx * kHaloBoxBytes + dy * 1024);
              const uint64_t db = make_sdesc_sw128(b_addr + stage * Cfg::kStageBytesB);
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / 16; ++kk) umma_f16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (kc | tap | kk) != 0 ? 1u : 0u);
              umma_commit(&empty_bar[stage]);
              if (tap == 8) {
                umma_commit(&a_empty[astage]);
                if (kc == p.cin_blocks - 1) umma_commit(&tmem_full_bar[acc]);
              }
            }
            __syncwarp();
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (++astage == kHaloAStages) {
            astage = 0;
            aphase ^= 1;
          }
        }
      }
    } else {
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int it = it0; it < it_end; it += it_step, ++local) {
      if (PAIR && cta_rank != 0) break;  // the leader's MMAs cover both CTAs
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tb + acc * BLOCK_N;
      for (int k = 0; k < p.num_k_blocks; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if constexpr (PAIR) {
          if (elect_one()) {
            constexpr uint32_t idesc2 = make_idesc(FMT ? 1u : 0u, 2 * BLOCK_M, BLOCK_N);
            const uint64_t da = make_sdesc_sw128(a_addr + stage * Cfg::kStageBytesA);
            const uint64_t db = make_sdesc_sw128(b_addr + stage * Cfg::kStageBytesB);
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / 16; ++kk) umma_f16_2sm(d_tmem, da + 2 * kk, db + 2 * kk, idesc2, (k | kk) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage], (uint16_t)0b11);
            if (k == p.num_k_blocks - 1) umma_commit_2sm(&tmem_full_bar[acc], (uint16_t)0b11);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        if (elect_one()) {
          const uint64_t da = make_sdesc_sw128(a_addr + stage * Cfg::kStageBytesA);
          const uint64_t db = make_sdesc_sw128(b_addr + stage * Cfg::kStageBytesB);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / 16; ++kk) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle row: +2 in the (addr >> 4) field
            umma_f16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
          }
          if constexpr (CL) umma_commit_mc(&empty_bar[stage], (uint16_t)0b11);  // ... in both CTAs: the peer's multicast writes here too
          else umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
          if (k == p.num_k_blocks - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    }
  }
  } else {
    // ===================== epilogue warps =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
    const int e = warp_idx - 4;
    const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
    const int half = e >> 2;           // column half handled by this warp
    constexpr int kChunks = BLOCK_N / 32;
    constexpr int kChunksPerHalf = (kChunks + 1) / 2;
    const int c_begin = half * kChunksPerHalf;
    const int c_end = (c_begin + kChunksPerHalf < kChunks) ? c_begin + kChunksPerHalf : kChunks;
    const int row_in_tile = quarter * 32 + lane;
    int local = 0;
    for (int it = it0; it < it_end; it += it_step, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int n_blk = it % p.num_n_blocks;
      const int m_tile = CL ? 2 * (it / p.num_n_blocks) + cta_rank : it / p.num_n_blocks;  // cluster: may be one past the last M tile (all rows invalid)
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
        rc.valid = (hh < p.H) && (ww < p.W) && (m_tile < p.num_m_tiles);  // cluster forms: the odd tile out of the last pair is empty
        row = ((long long)img * p.H + hh) * p.W + ww;
      } else {
        row = (long long)m_tile * BLOCK_M + row_in_tile;
        rc.valid = row < p.M;
      }
      rc.out_row = row;
      rc.res_row = row;
      rc.b = rc.y = rc.x = 0;
      if constexpr (EPI == EPI_PIXSHUF) {
        const int hw = p.rm0 * p.rm1;
        rc.b = int(row / hw);
        const int rem = int(row - (long long)rc.b * hw);
        rc.y = rem / p.rm1;
        rc.x = rem - rc.y * p.rm1;
      } else if constexpr (EPI == EPI_QKV) {
        rc.b = int(row / p.rm0);
        rc.x = int(row - (long long)rc.b * p.rm0) + p.qkv_toff;      // token index in the output (KV-cache slot offset)
        rc.out_row = (long long)rc.b * p.qkv_tpo + rc.x;
      } else if constexpr (EPI == EPI_PLAIN || EPI == EPI_RES) {
        if (p.row_map == VDN_ROWMAP_TEMPORAL) {
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
        }
      }

      // stage this tile's bias / gamma slices in shared memory (all epilogue warps; named barrier 1)
      asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiWarps * 32) : "memory");
      {
        const int t = threadIdx.x - 128;
        if (t < BLOCK_N) {
          const int n = n_blk * BLOCK_N + t;
          if (p.bias != nullptr) sbias[t] = n < p.N ? __ldg(p.bias + n) : 0.0f;
          if (p.gamma != nullptr) sgamma[t] = n < p.N ? __ldg(p.gamma + n) : 0.0f;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiWarps * 32) : "memory");

      const int nb = n_blk * BLOCK_N;
      const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BLOCK_N;
      if constexpr (EPI == EPI_TMA) {
        uint8_t* stg = smem_stg + e * kStagingBytesPerWarp;
        const int row0 = m_tile * BLOCK_M + quarter * 32;
        uint32_t accr[32];
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        if (c_begin < c_end) tmem_ld32(t_row + c_begin * 32, accr);
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          tmem_ld_wait();
          const int n0 = nb + c * 32;
          float v[32];
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(accr[i]) + sbias[c * 32 + i];
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(accr[i]);
          }
          if (c + 1 < c_end) tmem_ld32(t_row + (c + 1) * 32, accr);
          if (n0 < p.N) {  // warp-uniform
            if (p.act == VDN_ACT_GELU) {
              gelu_fast_batch<8>(v);
              gelu_fast_batch<8>(v + 8);
              gelu_fast_batch<8>(v + 16);
              gelu_fast_batch<8>(v + 24);
            } else if (p.act == VDN_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
            }
            if (p.gamma != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= sgamma[c * 32 + i];
            }
            if (p.out_f32) {
              // one 4 KB buffer: [32 rows][128 B], 128B swizzle (16-byte piece ^= row & 7) -> conflict-free 16-byte stores
              if (elect_one()) tma_store_wait_read<0>();
              __syncwarp();
              uint8_t* rowp = stg + lane * 128;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                *reinterpret_cast<float4*>(rowp + ((k ^ (lane & 7)) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              if (elect_one()) {
                if (p.res != nullptr) tma_reduce_add_2d(&tmC, stg, n0, row0);
                else tma_store_2d(&tmC, stg, n0, row0);
                tma_store_commit();
              }
            } else {
              // two 2 KB buffers: [32 rows][64 B], 64B swizzle (16-byte piece ^= (row >> 1) & 3)
              uint8_t* buf = stg + (c & 1) * 2048;
              if (elect_one()) tma_store_wait_read<1>();
              __syncwarp();
              uint8_t* rowp = buf + lane * 64;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint4*>(rowp + ((k ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(pack2<FMT>(v[8 * k], v[8 * k + 1]), pack2<FMT>(v[8 * k + 2], v[8 * k + 3]), pack2<FMT>(v[8 * k + 4], v[8 * k + 5]),
                               pack2<FMT>(v[8 * k + 6], v[8 * k + 7]));
              fence_proxy_async_smem();
              __syncwarp();
              if (elect_one()) {
                tma_store_2d(&tmC, buf, n0, row0);
                tma_store_commit();
              }
            }
          }
        }
        tc_fence_before();
        if constexpr (PAIR) mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty_bar[acc]), 0));
        else mbar_arrive(&tmem_empty_bar[acc]);
        continue;
      }
      uint32_t accr[32];
      ResBuf rb0, rb1;  // only live in the EPI_RES instantiation
      uint4 r2[4], r2b[4];  // second residual: two buffers, two chunks of look-ahead like the first (one chunk was not enough for DRAM latency)
      if constexpr (EPI == EPI_RES) {
        // residual operands of the first two chunks are requested before the accumulator is even ready
        if (c_begin < c_end) {
          prefetch_res(p, rc, nb + c_begin * 32, rb0);
          prefetch_res2(p, rc, nb + c_begin * 32, r2);
        }
        if (c_begin + 1 < c_end) {
          prefetch_res(p, rc, nb + (c_begin + 1) * 32, rb1);
          prefetch_res2(p, rc, nb + (c_begin + 1) * 32, r2b);
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if (c_begin < c_end) tmem_ld32(t_row + c_begin * 32, accr);

      auto chunk = [&](int c, ResBuf& rb, uint4 (&r2x)[4]) {
        tmem_ld_wait();
        const int n0 = nb + c * 32;
        float v[32];
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(accr[i]) + sbias[c * 32 + i];
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(accr[i]);
        }
        // the accumulator registers are free again: next chunk's TMEM load runs under this chunk's math and stores
        if (c + 1 < c_end) tmem_ld32(t_row + (c + 1) * 32, accr);
        if constexpr (EPI == EPI_QKV) {
          if (p.qkv_tma && n0 < p.qkv_split) {  // warp-uniform: q|k chunk -> 64B-swizzled staging tile -> TMA store (rows >= M are clipped)
            uint8_t* buf = smem_stg + e * kStagingBytesPerWarp + (c & 1) * 2048;
            if (elect_one()) tma_store_wait_read<1>();
            __syncwarp();
            uint8_t* rowp = buf + lane * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<uint4*>(rowp + ((k ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(pack2<FMT>(v[8 * k], v[8 * k + 1]), pack2<FMT>(v[8 * k + 2], v[8 * k + 3]), pack2<FMT>(v[8 * k + 4], v[8 * k + 5]),
                             pack2<FMT>(v[8 * k + 6], v[8 * k + 7]));
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
              tma_store_2d(&tmC, buf, n0, m_tile * BLOCK_M + quarter * 32);
              tma_store_commit();
            }
            return;
          }
        }
        if (rc.valid && n0 < p.N) {
          if constexpr (EPI == EPI_RES) {
            epi_plain<FMT>(p, rc, n0, c * 32, v, rb, r2x, (c + 2 < c_end) ? nb + (c + 2) * 32 : -1, (c + 2 < c_end) ? nb + (c + 2) * 32 : -1, sgamma);
          } else if constexpr (EPI == EPI_PLAIN) {
            epi_noresidual<FMT>(p, rc, n0, v);
          } else if constexpr (EPI == EPI_QKV) {
            epi_qkv<FMT>(p, rc, n0, v);
          } else if constexpr (EPI == EPI_GEGLU) {
            epi_geglu<FMT>(p, rc, n0, v);
          } else if constexpr (EPI == EPI_PIXSHUF) {
            epi_pixshuf<FMT>(p, rc, n0, v);
          } else {
            epi_head(p, rc, v);
          }
        }
      };
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 2) {
        chunk(c, rb0, r2);
        if (c + 1 < c_end) chunk(c + 1, rb1, r2b);
      }
      tc_fence_before();
      if constexpr (PAIR) mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty_bar[acc]), 0));
      else mbar_arrive(&tmem_empty_bar[acc]);
    }
    if constexpr (EPI == EPI_TMA || EPI == EPI_QKV) {
      if (elect_one()) tma_store_wait_all();  // staging memory and the global writes must outlive the bulk copies
      __syncwarp();
    }
  }

  tc_fence_before();
  if constexpr (CL) cluster_sync_all();  // the peer may still signal this CTA's barriers / fill its stages until it is done, too
  else __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool no_tma_epi_env() {
  static const bool v = getenv("VDN_NO_TMA_EPI") != nullptr;  // read once per process, like the other switches
  return v;
}

template <int BLOCK_N, int EPI, int FMT, int HALO = 0>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmKParams& p, cudaStream_t stream,
                       const CUtensorMap* tmBh = nullptr) {
  using Cfg = GemmCfg<BLOCK_N, EPI, HALO>;
  static bool configured_dev[kMaxDevices] = {};  // function attributes are per device
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, EPI, FMT, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(gemm): ") + cudaGetErrorString(e));
    configured = true;
  }
  if constexpr (HALO >= 2) {
    // clusters of two CTAs, each pair working through (M-tile pair, N block) items
    const int pairs = ((p.num_m_tiles + 1) / 2) * p.num_n_blocks;
    int clusters = num_sms() / 2;
    if (pairs < clusters) clusters = pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int max_clusters = -1;
    if (max_clusters < 0) {
      if (cudaOccupancyMaxActiveClusters(&max_clusters, gemm_tc_kernel<BLOCK_N, EPI, FMT, HALO>, &cfg) != cudaSuccess || max_clusters <= 0) {
        cudaGetLastError();
        max_clusters = 0;
      }
    }
    if (max_clusters > 0 && clusters > max_clusters) cfg.gridDim = dim3(2 * max_clusters);  // every cluster must be co-resident with its peer only, but keep the grid persistent
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BLOCK_N, EPI, FMT, HALO>, tmA, tmB, tmC, *tmBh, p);
    if (e != cudaSuccess) return set_error(std::string("cudaLaunchKernelEx(gemm cluster): ") + cudaGetErrorString(e));
    count_launch();
    return check_launch("gemm_tc_kernel(cluster)");
  } else {
    const int num_tiles = p.num_m_tiles * p.num_n_blocks;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    gemm_tc_kernel<BLOCK_N, EPI, FMT, HALO><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmC, tmB, p);
    count_launch();
    return check_launch("gemm_tc_kernel");
  }
}

template <int EPI, int FMT>
static int launch_gemm_bn(int block_n, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmKParams& p,
                          cudaStream_t stream) {
  if constexpr (EPI == EPI_HEAD) {
    return launch_gemm<32, EPI, FMT>(tmA, tmB, tmC, p, stream);
  } else if constexpr (EPI == EPI_PLAIN || EPI == EPI_RES || EPI == EPI_TMA) {
    switch (block_n) {
      case 256: return launch_gemm<256, EPI, FMT>(tmA, tmB, tmC, p, stream);
      case 128: return launch_gemm<128, EPI, FMT>(tmA, tmB, tmC, p, stream);
      case 64: return launch_gemm<64, EPI, FMT>(tmA, tmB, tmC, p, stream);
      default: return launch_gemm<32, EPI, FMT>(tmA, tmB, tmC, p, stream);
    }
  } else {
    // QKV / GEGLU / pixel-shuffle outputs are always wide (N >= 256): two tile widths suffice
    if (block_n == 256) return launch_gemm<256, EPI, FMT>(tmA, tmB, tmC, p, stream);
    return launch_gemm<128, EPI, FMT>(tmA, tmB, tmC, p, stream);
  }
}

// HALO convolutions: conv-capable epilogues only, tile widths <= 128 (the wide 256-channel convolutions already run tensor-bound)
template <int EPI, int FMT>
static int launch_halo_bn(int block_n, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmKParams& p, cudaStream_t stream) {
  if constexpr (EPI == EPI_HEAD) {
    return launch_gemm<32, EPI, FMT, 1>(tmA, tmB, tmC, p, stream);
  } else {
    switch (block_n) {
      case 128: return launch_gemm<128, EPI, FMT, 1>(tmA, tmB, tmC, p, stream);
      case 64: return launch_gemm<64, EPI, FMT, 1>(tmA, tmB, tmC, p, stream);
      default: return launch_gemm<32, EPI, FMT, 1>(tmA, tmB, tmC, p, stream);
    }
  }
}

template <int FMT>
static int launch_halo_epi(int epi, int block_n, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmKParams& p,
                           cudaStream_t stream) {
  switch (epi) {
    case EPI_HEAD: return launch_halo_bn<EPI_HEAD, FMT>(block_n, tmA, tmB, tmC, p, stream);
    case EPI_RES: return launch_halo_bn<EPI_RES, FMT>(block_n, tmA, tmB, tmC, p, stream);
    default: return launch_halo_bn<EPI_PLAIN, FMT>(block_n, tmA, tmB, tmC, p, stream);
  }
}

template <int FMT>
static int launch_gemm_epi(int epi, int block_n, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmKParams& p,
                           cudaStream_t stream) {
  switch (epi) {
    case EPI_QKV: return launch_gemm_bn<EPI_QKV, FMT>(block_n, tmA, tmB, tmC, p, stream);
    case EPI_GEGLU: return launch_gemm_bn<EPI_GEGLU, FMT>(block_n, tmA, tmB, tmC, p, stream);
    case EPI_PIXSHUF: return launch_gemm_bn<EPI_PIXSHUF, FMT>(block_n, tmA, tmB, tmC, p, stream);
    case EPI_HEAD: return launch_gemm_bn<EPI_HEAD, FMT>(block_n, tmA, tmB, tmC, p, stream);
    case EPI_RES: return launch_gemm_bn<EPI_RES, FMT>(block_n, tmA, tmB, tmC, p, stream);
    case EPI_TMA: return launch_gemm_bn<EPI_TMA, FMT>(block_n, tmA, tmB, tmC, p, stream);
    default: return launch_gemm_bn<EPI_PLAIN, FMT>(block_n, tmA, tmB, tmC, p, stream);
  }
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
  p.act = d->act;
  p.row_map = d->row_map; p.rm0 = d->rm0; p.rm1 = d->rm1; p.rm2 = d->rm2; p.rm3 = d->rm3;
  p.head_w = d->head_w; p.head_b = d->head_b;
  p.qkv_split = d->qkv_split > 0 ? d->qkv_split : 2 * d->rm2;
  p.qkv_tpo = d->qkv_tokens_out > 0 ? d->qkv_tokens_out : d->rm0;
  p.qkv_toff = d->qkv_token_offset;
  p.qkv_tma = 0;
  p.conv = d->conv;
  int epi = EPI_PLAIN;
  if (d->head_w != nullptr) epi = EPI_HEAD;
  else if (d->geglu) epi = EPI_GEGLU;
  else if (d->row_map == VDN_ROWMAP_QKV_SPLIT) epi = EPI_QKV;
  else if (d->row_map == VDN_ROWMAP_PIXEL_SHUFFLE) epi = EPI_PIXSHUF;
  else if (d->res != nullptr || d->res2 != nullptr || d->gamma != nullptr) epi = EPI_RES;
  // TMA-store epilogue: plain tiles, or in-place fp32 accumulation (out += ...) as an L2 reduce-add
  {
    static const char* env = getenv("VDN_NO_TMA_EPI");
    const int es = d->out_f32 ? 4 : 2;
    const bool inplace_f32 = d->res != nullptr && d->res == d->out && d->res_f32 && d->out_f32 && d->ld_res == d->ldc;
    if ((epi == EPI_PLAIN || epi == EPI_RES) && env == nullptr && !d->conv && d->row_map == VDN_ROWMAP_IDENTITY && d->out2 == nullptr &&
        d->res2 == nullptr && (d->res == nullptr || inplace_f32) && (d->ldc * es) % 16 == 0)
      epi = EPI_TMA;
  }
  if (epi == EPI_QKV && (d->out_f32 || d->out2 == nullptr || d->res || d->gamma || d->act)) return set_error("vdn_gemm: QKV split takes bias only, 16-bit out and out2 = V^T");
  if (epi == EPI_QKV && (p.qkv_split % 8 != 0 || p.qkv_split >= d->N || (d->N - p.qkv_split) % 64 != 0 || d->rm0 <= 0 || d->rm1 <= 0 ||
                         p.qkv_toff < 0 || p.qkv_toff + d->rm0 > p.qkv_tpo || p.qkv_tpo > d->rm1))
    return set_error("vdn_gemm: bad QKV split geometry");
  if (epi == EPI_PIXSHUF && (d->out_f32 || d->res || d->res2 || d->gamma || d->act || d->out2)) return set_error("vdn_gemm: pixel-shuffle takes bias only and a 16-bit output");
  if (epi == EPI_GEGLU && (d->gamma || (d->act != VDN_ACT_NONE && d->act != VDN_ACT_SILU) || d->out2 || d->res2)) return set_error("vdn_gemm: geglu takes bias only (act = VDN_ACT_SILU selects the SwiGLU gate)");
  if (epi != EPI_PLAIN && epi != EPI_RES && epi != EPI_HEAD && epi != EPI_TMA && d->N < 256) return set_error("vdn_gemm: QKV / GEGLU / pixel-shuffle epilogues need N >= 256");

  // BLOCK_N: largest tile that does not waste more than necessary
  int block_n;
  if (d->N > 128) block_n = 256;
  else if (d->N > 64) block_n = 128;
  else if (d->N > 32) block_n = 64;
  else block_n = 32;
  // prefer 128-wide tiles when 256 would leave most of the last tile empty (e.g. N = 384)
  if (block_n == 256 && (d->N % 256) != 0 && (d->N % 256) <= 128) block_n = 128;
  // fp32 read-modify-write outputs with a short K loop are HBM/latency bound in the epilogue: with 128-wide tiles each epilogue
  // warp owns two 32-column chunks and both residual chunks are requested a whole tile ahead of their use
  {
    static const char* env = getenv("VDN_RMW_BN");
    const int rmw_bn = env ? atoi(env) : 256;
    const long long k_total = (long long)d->K * (d->conv ? 9 : 1);
    if (block_n == 256 && epi == EPI_RES && d->out_f32 && d->res != nullptr && k_total <= 1024 && rmw_bn == 128) block_n = 128;
  }
  p.num_n_blocks = (int)((d->N + block_n - 1) / block_n);

  CUtensorMap tmA, tmB;
  bool halo = false;
  if (d->conv) {
    if (d->M != (int64_t)d->B * d->H * d->W) return set_error("vdn_gemm(conv): M != B*H*W");
    if ((d->K * 2) % 16 != 0) return set_error("vdn_gemm(conv): C_in*2 must be a multiple of 16 bytes");
    pick_spatial_tile(d->H, d->W, &p.TH, &p.TW);
    {
      // HALO path (16 x 8 tiles, three column-shifted haloed loads per channel chunk): narrow-N convolutions on maps large enough
      // that the fixed tile shape wastes < 12 % more pixels than the best free choice
      static const char* env = getenv("VDN_CONV_HALO");
      const long long best = (long long)((d->H + p.TH - 1) / p.TH) * ((d->W + p.TW - 1) / p.TW);
      const long long fixed = (long long)((d->H + kHaloTH - 1) / kHaloTH) * ((d->W + kHaloTW - 1) / kHaloTW);
      const bool want = env ? atoi(env) != 0 : true;
      halo = want && block_n <= 128 && (epi == EPI_PLAIN || epi == EPI_RES || epi == EPI_HEAD) && fixed * 100 <= best * 112;
      if (halo) { p.TH = kHaloTH; p.TW = kHaloTW; }
    }
    p.H = d->H; p.W = d->W;
    p.tiles_h = (d->H + p.TH - 1) / p.TH;
    p.tiles_w = (d->W + p.TW - 1) / p.TW;
    p.num_m_tiles = p.tiles_h * p.tiles_w * d->B;
    p.cin_blocks = (int)((d->K + BLOCK_K - 1) / BLOCK_K);
    p.num_k_blocks = 9 * p.cin_blocks;
    if (d->ldw < (int64_t)p.num_k_blocks * BLOCK_K) return set_error("vdn_gemm(conv): ldw < 9*roundup(C_in,64)");
    const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->K * 2, (uint64_t)d->K * 2 * d->W, (uint64_t)d->K * 2 * d->W * d->H};
    const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)p.TW, (uint32_t)(halo ? p.TH + 2 : p.TH), 1u};
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
  CUtensorMap tmC = tmB;  // placeholder unless the TMA epilogue is used
  if (epi == EPI_QKV && p.qkv_tpo == d->rm0 && p.qkv_toff == 0 && (d->ldc * 2) % 16 == 0 && p.qkv_split % 32 == 0 && !no_tma_epi_env()) {
    // q|k rows land at their own row index: one 32 x 32 box per warp chunk
    const uint64_t dims[2] = {(uint64_t)p.qkv_split, (uint64_t)d->M};
    const uint64_t strides[1] = {(uint64_t)d->ldc * 2};
    const uint32_t box[2] = {32u, 32u};
    if (make_tensor_map_ex(&tmC, d->out, fmt, 64, 2, dims, strides, box)) return 1;
    p.qkv_tma = 1;
  }
  if (epi == EPI_TMA) {
    const int es = d->out_f32 ? 4 : 2;
    const uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
    const uint64_t strides[1] = {(uint64_t)d->ldc * es};
    const uint32_t box[2] = {32u, 32u};
    if (make_tensor_map_ex(&tmC, d->out, d->out_f32 ? 2 : fmt, d->out_f32 ? 128 : 64, 2, dims, strides, box)) return 1;
  }
  {
    // 2-CTA clusters (VDN_GEMM_CLUSTER=0 disables): wide GEMMs / 256-wide convolutions with many M tiles.
    //  - multicast form (HALO == 2): two M = 128 MMAs, each CTA loads half of the weight tile and multicasts it to both.
    //    qkv 234 -> 227 us, fc1 316 -> 308 us.
    //  - pair form (HALO == 3): one 2-SM MMA of M = 256, each CTA stages only half of the weight tile.  proj 108.6 -> 102.6 us,
    //    fc2 302.8 -> 291.5 us, qkv 225 -> 211 us, but fc1 (GELU epilogue) 302 -> 306 us: the two epilogues of a pair gate one
    //    accumulator stage.  Default: QKV split, in-place fp32 accumulate, 256-wide convolutions.  VDN_GEMM_PAIR=0 never,
    //    =2 also every other TMA-epilogue GEMM.
    static const char* env = getenv("VDN_GEMM_CLUSTER");
    static const char* env_pair = getenv("VDN_GEMM_PAIR");
    const int pair_mode = env_pair ? atoi(env_pair) : 1;
    static const char* env_cp = getenv("VDN_CONV_PAIR");
    const bool conv_pair = env_cp == nullptr || atoi(env_cp) != 0;
    const bool wide = (env == nullptr || atoi(env) != 0) && block_n == 256 && p.num_m_tiles >= 32 && d->N % 256 == 0 && !halo;
    if (wide && ((!d->conv && (epi == EPI_TMA || epi == EPI_QKV)) || (d->conv && pair_mode != 0 && conv_pair && (epi == EPI_PLAIN || epi == EPI_RES)))) {
      CUtensorMap tmBh;
      const uint64_t kw = d->conv ? (uint64_t)p.num_k_blocks * BLOCK_K : (uint64_t)d->K;
      const uint64_t dims[2] = {kw, (uint64_t)d->N};
      const uint64_t strides[1] = {(uint64_t)d->ldw * 2};
      const uint32_t box[2] = {(uint32_t)BLOCK_K, 128u};
      if (make_tensor_map(&tmBh, d->w, fmt, 2, dims, strides, box)) return 1;
#define VDN_CL(E, H) (fmt ? launch_gemm<256, E, 1, H>(tmA, tmB, tmC, p, stream, &tmBh) : launch_gemm<256, E, 0, H>(tmA, tmB, tmC, p, stream, &tmBh))
      if (d->conv) return epi == EPI_RES ? VDN_CL(EPI_RES, 3) : VDN_CL(EPI_PLAIN, 3);
      const bool accumulate = d->out_f32 && d->res != nullptr;
      if (epi == EPI_TMA) return (pair_mode == 2 || (pair_mode == 1 && accumulate)) ? VDN_CL(EPI_TMA, 3) : VDN_CL(EPI_TMA, 2);
      return pair_mode != 0 ? VDN_CL(EPI_QKV, 3) : VDN_CL(EPI_QKV, 2);
#undef VDN_CL
    }
  }
  if (halo) return fmt ? launch_halo_epi<1>(epi, block_n, tmA, tmB, tmC, p, stream) : launch_halo_epi<0>(epi, block_n, tmA, tmB, tmC, p, stream);
  return fmt ? launch_gemm_epi<1>(epi, block_n, tmA, tmB, tmC, p, stream) : launch_gemm_epi<0>(epi, block_n, tmA, tmB, tmC, p, stream);
}

// Attention kernels.
//
// 1. vdn_flash_attn — spatial ViT attention (head_dim 64, non-causal) on tcgen05:
//      S  = Q K^T   : UMMA 128 x 128 x 64, Q/K tiles K-major in 128B-swizzled smem (TMA, 5-D map straight out of the QKV GEMM output)
//      P  = softmax : 4 softmax warps, one query row per thread, S read from TMEM twice (max pass, exp pass) to keep registers low,
//                     online rescaling in fp32, P written to smem as a 128B-swizzled K-major A operand
//      O += P V     : UMMA 128 x 64 x 128 with V^T (d-major rows, keys contiguous; produced transposed by the QKV GEMM epilogue)
//                     fresh TMEM accumulator per KV tile, running O kept in registers and rescaled there
//    One CTA per (128-query tile, head, frame); K/V^T double-buffered through TMA; 2 CTAs per SM overlap softmax with MMA.
//
// 2. vdn_temporal_attn — motion-module attention over T <= 32 frames per (pixel, head): tiny 32x32 problems, CUDA cores,
//    one warp per (pixel, head), lane = query frame, K/V broadcast from padded shared memory.
#include <stdlib.h>

#include "../../include/vdn_b200.h"
#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

// ------------------------------------------------------------------------------------------------
// flash attention
// ------------------------------------------------------------------------------------------------
constexpr int FA_BM = 128;     // queries per softmax group (one TMEM lane per query row)
constexpr int FA_BN = 128;     // keys per tile
constexpr int FA_D = 64;       // head dim
constexpr int FA_GROUPS = 2;   // query tiles per CTA, processed ping-pong by two softmax warpgroups
constexpr int FA_THREADS = FA_GROUPS * 128 + 128;  // + one control warpgroup: MMA issuer warp, TMA producer warp, 2 idle warps
constexpr int FA_TILE = FA_BM * FA_D * 2;         // 16 KB : one Q / K / V^T tile
constexpr int FA_VSTAGE_MAX = 2 * 80 * 128;        // a V^T stage of the ones-column form: 2 chunks of [80 rows x 64 keys]
constexpr int FA_SMEM = 2 * FA_TILE /*Q*/ + 2 * FA_TILE /*K x2*/ + 2 * FA_VSTAGE_MAX /*V^T x2*/ + 256;
#ifndef VDN_FA_POLY
#define VDN_FA_POLY 2
#endif
constexpr int FA_POLY = VDN_FA_POLY;  // of every 8 scores, this many take 2^x on the FMA pipe (exp2_poly), the rest on the XU pipe
constexpr int FA_TMEM_COLS = 512;  // 128-key tiles: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512); 112-key (ones-column) tiles: S 2 x 112, O 2 x 80, P 2 x 56

// ---- packed fp32 arithmetic (Blackwell FFMA2 / FADD2 on 64-bit register pairs) and the 3-input FMNMX3 for the softmax warps ----
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float a, float b) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float max3f(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
// two 2^x on the FMA pipe (same polynomial as exp2_poly): 2 FMNMX + 6 packed instructions + 2 shift-adds
__device__ __forceinline__ void exp2_poly_pair(float& x0, float& x1) {
  const f32x2_t x = pk2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const f32x2_t magic = pk2(12582912.0f, 12582912.0f);
  const f32x2_t t = add2(x, magic);
  const f32x2_t n = sub2(t, magic);
  const f32x2_t f = sub2(x, n);
  f32x2_t pq = fma2(pk2(0.05520550534129143f, 0.05520550534129143f), f, pk2(0.24261397123336792f, 0.24261397123336792f));
  pq = fma2(pq, f, pk2(0.6932547688484192f, 0.6932547688484192f));
  pq = fma2(pq, f, pk2(0.9999276995658875f, 0.9999276995658875f));
  float p0, p1, t0, t1;
  up2(pq, p0, p1);
  up2(t, t0, t1);
  x0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  x1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// Softmax instruction-mix variants (VDN_FA_VARIANT; measured with scripts/microbench/exp_phase3.cu and scripts/run_flash.py):
//   packed  : scale-sub as FFMA2 and the row sum as FADD2 (two scores per instruction)
//   max3    : row max with the 3-input FMNMX3
//   poly16  : of every 16 scores this many take 2^x on the FMA pipe, the rest on the XU pipe (MUFU.EX2)
//   polypk  : the polynomial runs on register pairs (exp2_poly_pair)
//   token   : the exp phases of the two softmax groups strictly alternate (g0 tile n, g1 tile n, g0 tile n+1, ...): a group enters
//             its exp phase only after the other group has left its own, so one group's MUFU stream runs against the other group's
//             MUFU-free work (wait for S, TMEM loads, row max, P store) instead of both streams fighting for the XU pipe and then
//             both leaving it idle.  This is also the order in which the in-order MMA issuer waits for the P tiles.
//   latewait: the wait for PV(j-1) (P buffer free, O safe to rescale) moves from before the exp phase to just before the first
//             tcgen05.st of P(j): half of the exponentials are computed while the previous P V MMA may still be queued
//   mma2    : one MMA issuer warp per softmax group (warps 8 and 10) instead of a single in-order issuer
// Round-2 experiments on the hand-off structure (profiles/r02_flash_experiments.txt has the numbers and the VDN_FA_TIMELINE
// traces): late wait for PV(j-1), two issuers, and a software-pipelined softmax (loads + row max of tile j+1 inside the exp phase of
// tile j, S issued as two 64-key halves) all measured slower than variant 6 — two softmax warps per sub-partition get the same
// exp throughput whether they alternate or overlap (~1150 cycles per 128-score warp tile), so only fewer issue cycles per score help.
//   ones    : KV tiles of 112 keys and a ones row appended to every V^T tile: O gets a 65th column that accumulates the row sum of the
//             (16-bit rounded) probabilities on the tensor core, so the softmax warps drop one packed add per two scores.  TMEM:
//             S 2 x 112, O 2 x 80, P 2 x 56 = 496 columns (128-key tiles would need 544)
struct FaVariant { int packed, max3, poly16, polypk, latewait, token, dbg, mma2, ones; };  // dbg (timing experiments only, wrong results): 1 = no exp, 2 = no row max, 3 = neither
__host__ __device__ constexpr FaVariant fa_variant(int v) {
  return v == 0 ? FaVariant{0, 0, 2 * FA_POLY, 0, 0, 0}   // round-1 kernel
       : v == 1 ? FaVariant{1, 1, 4, 1, 0, 1}
       : v == 2 ? FaVariant{1, 0, 4, 1, 0, 1}
       : v == 3 ? FaVariant{1, 1, 5, 1, 0, 1}
       : v == 4 ? FaVariant{1, 1, 2, 1, 0, 1}
       : v == 5 ? FaVariant{1, 1, 6, 1, 0, 1}
       : v == 6 ? FaVariant{1, 1, 4, 1, 0, 0}
       : v == 7 ? FaVariant{1, 1, 4, 1, 1, 1}
       : v == 8 ? FaVariant{1, 1, 4, 1, 0, 0, 1}
       : v == 9 ? FaVariant{1, 1, 4, 1, 0, 0, 2}
       : v == 10 ? FaVariant{1, 1, 4, 1, 0, 0, 3}
       : v == 11 ? FaVariant{1, 1, 4, 1, 1, 0}            // half of tile j's exponentials before the wait for PV(j-1), no token
       : v == 12 ? FaVariant{1, 1, 4, 1, 0, 0, 0, 1}      // one MMA issuer warp per softmax group
       : v == 13 ? FaVariant{1, 1, 4, 1, 0, 0, 1, 1}      // timing only: 12 without exp
       : v == 14 ? FaVariant{1, 1, 4, 1, 0, 0, 0, 0, 1}   // row sum on the tensor core (ones column), 112-key tiles
       : v == 15 ? FaVariant{1, 1, 3, 1, 0, 0, 0, 0, 1}
       : v == 16 ? FaVariant{1, 1, 5, 1, 0, 0, 0, 0, 1}
       : v == 17 ? FaVariant{1, 1, 4, 1, 0, 0, 1, 0, 1}   // timing only: 14 without exp
       : v == 18 ? FaVariant{1, 1, 6, 1, 0, 0, 0, 0, 1}
       : v == 19 ? FaVariant{1, 1, 7, 1, 0, 0, 0, 0, 1}
                 : FaVariant{1, 1, 8, 1, 0, 0, 0, 0, 1};
}
template <int VAR> constexpr FaVariant kFaVar = fa_variant(VAR);

// Debug builds only (VDN_EXTRA_NVCC_FLAGS=-DVDN_FA_TIMELINE): CTA 0 records (event, SM clock) pairs per warp role; read back with
// vdn_debug_fa_timeline (scripts/fa_timeline.py draws the hand-offs between the softmax groups and the MMA issuer).
#ifdef VDN_FA_TIMELINE
constexpr int FA_TL_CAP = 4096;
__device__ unsigned long long g_fa_tl[5][FA_TL_CAP];
__device__ int g_fa_tl_n[5];
__device__ __forceinline__ void fa_tl(int slot, int ev, int lane, int& n) {
  if (blockIdx.x == 0 && lane == 0 && n < FA_TL_CAP) {
    g_fa_tl[slot][n] = ((unsigned long long)ev << 48) | ((unsigned long long)clock64() & 0xffffffffffffull);
    g_fa_tl_n[slot] = ++n;
  }
}
#define FA_TL(slot, ev) fa_tl(slot, ev, lane + (warp_idx < 8 ? (warp_idx & 3) * 32 : 0), tl_n)
#else
#define FA_TL(slot, ev)
#endif
constexpr int FA_NUM_VARIANTS = 21;
#ifndef VDN_FA_DEFAULT_VARIANT
#define VDN_FA_DEFAULT_VARIANT 16
#endif


// The softmax of a head_dim-64 attention is bound by the XU pipe: one ex2 per score at 8 cycles per warp instruction per SM
// sub-partition (measured, scripts/microbench/pipes.cu), i.e. 1024 cycles per 128x128 score tile per SM, twice the tensor-pipe
// time of the two MMAs.  The kernel is therefore organised around keeping the XU pipe busy:
//   * one CTA per SM owns TWO 128-query tiles of the same (frame, head), one softmax warpgroup each (one query row per thread,
//     the whole 128-key score row in registers), so that one group's MUFU stream overlaps the other's non-MUFU work (wait for
//     S, TMEM loads, masking, row max, O rescale).  A warp's own instruction costs add up (MUFU 8 + FFMA 1 + FADD 1 + F2FP 0.5
//     cycles per score, measured in scripts/microbench/exp_phase.cu): one warp alone reaches 10.5 cycles per score, two
//     overlapping warps 8.5 — which is why the groups are NOT serialised by a token (tried: slower);
//   * P never touches shared memory: the probabilities are packed to 16 bits and written to TMEM (tcgen05.st), and O += P V
//     reads its A operand from TMEM.  With P in shared memory the operand fetches of the two MMAs plus the P stores came to
//     ~112 KB per tile per group against a 128 B/clk shared-memory pipe — as long as the exp phase itself;
//   * S tiles are issued two KV tiles ahead into per-group TMEM buffers, K and V^T have independent 2-stage TMA rings,
//     O accumulates in TMEM across KV tiles (lazy rescale: only when the running max grows by more than 2^8).
// Two independent CTAs per SM (the previous design) drifted into phase and left the XU pipe 47 % busy (ncu, profiles/).
template <int FMT, int VAR>
__global__ void __launch_bounds__(FA_THREADS, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmVT,
                  void* __restrict__ out, int tokens, int tokens_kv, int heads, int C, int num_units) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                      // [group]
  constexpr bool ONES = kFaVar<VAR>.ones != 0;
  constexpr int BN = ONES ? 112 : FA_BN;       // keys per tile
  constexpr int OC = ONES ? 80 : FA_D;         // O columns / V^T rows per MMA: head dim (+ the ones row, padded to 16)
  constexpr int VCH = OC * 128;                // bytes of one V^T chunk [OC rows x 64 keys]
  constexpr int VST = 2 * VCH;                 // bytes of one V^T stage
  constexpr uint32_t T_O = 2 * BN, T_P = 2 * BN + 2 * OC;  // TMEM columns: S_g at g BN, O_g at T_O + g OC, P_g at T_P + g BN / 2
  uint8_t* sK = smem + 2 * FA_TILE;        // [stage]
  uint8_t* sV = smem + 4 * FA_TILE;        // [stage], each = 2 chunks [OC d rows x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * FA_TILE + 2 * FA_VSTAGE_MAX);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [group]
  uint64_t* s_free = bars + 11;   // [group]
  uint64_t* p_full = bars + 13;   // [group]
  uint64_t* pv_done = bars + 15;  // [group]
  uint64_t* o_free = bars + 17;   // [group]
  uint64_t* q_empty = bars + 19;
  uint64_t* tok = bars + 20;      // [group]: exp-phase hand-over between the softmax groups
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef VDN_FA_TIMELINE
  int tl_n = 0;
#endif
  const int nt = (tokens_kv + BN - 1) / BN;  // KV tiles (cross-attention: tokens_kv != tokens)
  const int nq_tiles = (tokens + FA_BM - 1) / FA_BM;
  const int nqp = (nq_tiles + FA_GROUPS - 1) / FA_GROUPS;  // query-tile pairs per (frame, head)

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need 1024-byte alignment

  if (warp_idx == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmVT);
      constexpr int consumers = kFaVar<VAR>.mma2 ? 2 : 1;  // MMA issuer warps that release the Q / K / V^T stages
      mbar_init(q_full, 1);
      mbar_init(q_empty, consumers);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], consumers);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], consumers);
        mbar_init(&s_full[i], 1);
        mbar_init(&s_free[i], 128);
        mbar_init(&p_full[i], 128);
        mbar_init(&pv_done[i], 1);
        mbar_init(&o_free[i], 128);
        mbar_init(&tok[i], 128);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, FA_TMEM_COLS);
  }
  if constexpr (ONES) {
    // rows 64..79 of every V^T chunk are constant: row 64 = 1.0 for every key (O column 64 accumulates the row sum of P), rows 65..79
    // = 0.  TMA only ever writes rows 0..63.
    const uint32_t one2 = FMT ? 0x3F803F80u : 0x3C003C00u;
    for (int i = threadIdx.x; i < 4 * 128; i += FA_THREADS) {  // 4 regions (2 stages x 2 chunks) of 16 rows x 128 B, 16 bytes per item
      const int region = i >> 7, piece = i & 127;
      const uint32_t v = piece < 8 ? one2 : 0u;
      *reinterpret_cast<uint4*>(sV + (region >> 1) * VST + (region & 1) * VCH + 64 * 128 + piece * 16) = make_uint4(v, v, v, v);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // Persistent CTA: work unit u = (query-tile pair, head, frame); TMEM, barriers and the K / V rings live across units, so
  // the producer prefetches the next unit's Q / K / V under the current unit's last tiles and output stores (one CTA per SM:
  // a per-unit launch would expose ~3 us of allocation + first-load latency 21 times per SM).  All barrier parities come from
  // running counters.  Registers are allocated per warpgroup: the control warpgroup gives most of its share to the two softmax
  // warpgroups, whose threads each hold a whole 128-key score row (12 warps x 168 = 8 x 232 + 4 x 40 registers per lane).
  if (warp_idx >= 8) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
  if (warp_idx == 9) {
    // ---------------- TMA producer (warp-uniform control flow, one elected lane issues) ----------------
    int kvc = 0, ui = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++ui) {
      const int qp = u % nqp, h = (u / nqp) % heads, b = u / (nqp * heads);
      const int q_tile0 = qp * FA_GROUPS;
      const int ngroups = (q_tile0 + 1) * FA_BM < tokens ? 2 : 1;
      if (ui > 0) mbar_wait(q_empty, (ui - 1) & 1);  // every S MMA of the previous unit has read its Q tiles
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, ngroups * FA_TILE);
        for (int g = 0; g < ngroups; ++g) tma_load_4d(sQ + g * FA_TILE, &tmQ, q_full, 0, h, (q_tile0 + g) * FA_BM, b);
      }
      __syncwarp();
      for (int j = 0; j < nt; ++j, ++kvc) {
        const int st = kvc & 1;
        const uint32_t ph = (kvc >> 1) & 1;
        FA_TL(3, 30);
        mbar_wait(&k_empty[st], ph ^ 1);
        FA_TL(3, 31);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], BN * 128);
          tma_load_4d(sK + st * FA_TILE, &tmK, &k_full[st], 0, h, j * BN, b);
        }
        __syncwarp();
        mbar_wait(&v_empty[st], ph ^ 1);
        FA_TL(3, 32);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[st], FA_TILE);
          tma_load_3d(sV + st * VST, &tmVT, &v_full[st], j * BN, 0, b * heads + h);
          tma_load_3d(sV + st * VST + VCH, &tmVT, &v_full[st], j * BN + 64, 0, b * heads + h);
        }
        __syncwarp();
      }
    }
  } else if (warp_idx == 8 || (kFaVar<VAR>.mma2 && warp_idx == 10)) {
    // ---------------- MMA issuer(s) ----------------
    // The whole warp runs the (warp-uniform) control flow and one elected lane issues: with the loops under `if (lane == 0)`
    // the compiler cannot prove the descriptors uniform and wraps every tcgen05.mma in an R2UR / ELECT / BRA.U.ANY waterfall.
    // mma2: one issuer warp per softmax group (warps 8 and 10, on different sub-partitions).  A single in-order issuer makes each
    // group's chain  P(j) -> PV(j) -> pv_done -> exp(j+1)  wait for the other group's PV and S batches and for the issuer's own
    // barrier round trips (measured with VDN_FA_TIMELINE: the issuer loop, not the XU pipe, set the pace, and the two groups'
    // exp phases strictly alternated); with two issuers a group's P V is queued the moment its P tile lands.
    constexpr bool MMA2 = kFaVar<VAR>.mma2 != 0;
    const int gb = MMA2 ? (warp_idx - 8) >> 1 : 0;       // first group served by this warp
    const int ge = MMA2 ? gb + 1 : FA_GROUPS;            // one past the last
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc_s = make_idesc(FMT ? 1u : 0u, 128, BN);
    constexpr uint32_t idesc_pv = make_idesc(FMT ? 1u : 0u, 128, OC);
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    int kv0 = 0, ui = 0;
    int si[2] = {0, 0};  // S tiles issued per group (global): S tile n may overwrite the buffer once tile n-1 was pulled into registers
    int pi[2] = {0, 0};  // PV tiles issued per group (global)
    int ug[2] = {0, 0};  // units processed per group
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++ui) {
      const int qp = u % nqp;
      const int ngroups = (qp * FA_GROUPS + 1) * FA_BM < tokens ? 2 : 1;
      const int gl = ge < ngroups ? ge : ngroups;  // groups [gb, gl) of this unit are this warp's
      auto issue_s = [&](int g, int j) {  // S_g(j) = Q_g K(j)^T
        const int kst = (kv0 + j) & 1;
        FA_TL(2 + 2 * gb, 16 + g);
        if (si[g] > 0) mbar_wait(&s_free[g], (si[g] - 1) & 1);
        FA_TL(2 + 2 * gb, 18 + g);
        if (g == gb) mbar_wait(&k_full[kst], ((kv0 + j) >> 1) & 1);
        FA_TL(2 + 2 * gb, 20 + g);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dq = make_sdesc_sw128(q_addr + g * FA_TILE);
          const uint64_t dk = make_sdesc_sw128(k_addr + kst * FA_TILE);
#pragma unroll
          for (int kk = 0; kk < FA_D / 16; ++kk) umma_f16(tb + g * BN, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
          if (g == gl - 1) {
            umma_commit(&k_empty[kst]);
            if (j == nt - 1) umma_commit(q_empty);  // last S of this unit: Q may be replaced
          }
        }
        __syncwarp();
        FA_TL(2 + 2 * gb, 24 + g);
        ++si[g];
      };
      mbar_wait(q_full, ui & 1);
      if (MMA2 && gb >= ngroups) {
        // single-tile unit: this warp's group has no work, but the stages are released by two arrivals
        for (int j = 0; j < nt; ++j) {
          const int st = (kv0 + j) & 1;
          const uint32_t ph = ((kv0 + j) >> 1) & 1;
          mbar_wait(&k_full[st], ph);
          if (elect_one()) mbar_arrive(&k_empty[st]);
          __syncwarp();
          mbar_wait(&v_full[st], ph);
          if (elect_one()) mbar_arrive(&v_empty[st]);
          __syncwarp();
        }
        if (elect_one()) mbar_arrive(q_empty);
        __syncwarp();
        kv0 += nt;
        continue;
      }
      // prologue: S(0) and S(1) of both groups (each group's S buffer is refilled as soon as the group has pulled it into registers)
      for (int j = 0; j < 2 && j < nt; ++j)
        for (int g = gb; g < gl; ++g) issue_s(g, j);
      for (int j = 0; j < nt; ++j) {
        const int vst = (kv0 + j) & 1;
        for (int g = gb; g < gl; ++g) {
          FA_TL(2 + 2 * gb, 10 + g);
          mbar_wait(&p_full[g], pi[g] & 1);
          FA_TL(2 + 2 * gb, 12 + g);
          if (g == gb) mbar_wait(&v_full[vst], ((kv0 + j) >> 1) & 1);
          FA_TL(2 + 2 * gb, 14 + g);
          if (j == 0 && ug[g] > 0) mbar_wait(&o_free[g], (ug[g] - 1) & 1);  // the group has read the previous unit's O
          tc_fence_after();
          if (elect_one()) {
            const uint64_t dv0 = make_sdesc_sw128(v_addr + vst * VST);
#pragma unroll
            for (int kk = 0; kk < BN / 16; ++kk) {
              // B = V^T chunk (kk >> 2), 16 keys further per MMA; A = P_g from TMEM (16 keys = 8 packed columns per MMA);
              // O accumulates in TMEM across KV tiles
              const uint64_t dv = dv0 + uint64_t(((kk >> 2) * VCH) >> 4) + 2 * (kk & 3);
              umma_f16_ts(tb + T_O + g * OC, tb + T_P + g * (BN / 2) + kk * 8, dv, idesc_pv, (j | kk) != 0 ? 1u : 0u);
            }
            umma_commit(&pv_done[g]);
            if (g == gl - 1) umma_commit(&v_empty[vst]);
          }
          __syncwarp();
          FA_TL(2 + 2 * gb, 22 + g);
          ++pi[g];
          if (j + 2 < nt) issue_s(g, j + 2);
        }
      }
      kv0 += nt;
      for (int g = gb; g < gl; ++g) ++ug[g];
    }
  }
  } else {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
  {
    // ---------------- softmax / output warpgroups: one query row per thread ----------------
    const int g = warp_idx >> 2;
    const int r = threadIdx.x & 127;  // row in the query tile == TMEM lane
    const uint32_t lane_off = uint32_t((warp_idx & 3) * 32) << 16;
    const uint32_t tmem_S = tmem_base + g * BN + lane_off;
    const uint32_t tmem_O = tmem_base + T_O + g * OC + lane_off;
    const uint32_t tmem_P = tmem_base + T_P + g * (BN / 2) + lane_off;
    constexpr int N3 = ONES ? 16 : 32;  // scores of the fourth register block
    const float sc = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
    int t = 0;   // tiles processed by this group (global)
    int tn = 0;  // exp-phase turns taken by this group (global; includes the empty turns of group 1 in single-tile units)
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int qp = u % nqp, h = (u / nqp) % heads, b = u / (nqp * heads);
      const int q_tile0 = qp * FA_GROUPS;
      if (g == 1 && !((q_tile0 + 1) * FA_BM < tokens)) {  // this unit has a single query tile
        if constexpr (kFaVar<VAR>.token) {            // keep the hand-over going: group 0 waits for this group's turn after every tile
          for (int j = 0; j < nt; ++j, ++tn) {
            mbar_wait(&tok[0], tn & 1);
            mbar_arrive(&tok[1]);
          }
        }
        continue;
      }
      float m = -INFINITY, l = 0.0f;
      for (int j = 0; j < nt; ++j, ++t) {
        const int nvalid = min(BN, tokens_kv - j * BN);
        FA_TL(g, 1);
        mbar_wait(&s_full[g], t & 1);
        FA_TL(g, 2);
        tc_fence_after();
        uint32_t s0[32], s1[32], s2[32], s3[32];
        tmem_ld32(tmem_S + 0, s0);
        tmem_ld32(tmem_S + 32, s1);
        tmem_ld32(tmem_S + 64, s2);
        if constexpr (ONES) tmem_ld16(tmem_S + 96, reinterpret_cast<uint32_t (&)[16]>(s3));
        else tmem_ld32(tmem_S + 96, s3);
        tmem_ld_wait();
        FA_TL(g, 3);
        tc_fence_before();
        mbar_arrive(&s_free[g]);  // S(j) is out of TMEM: the issuer may overwrite it with S(j+2)
        if (nvalid < BN) {     // last KV tile: keys beyond the sequence are zero-filled by TMA -> mask them out
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= nvalid) s0[i] = 0xff800000u;
            if (32 + i >= nvalid) s1[i] = 0xff800000u;
            if (64 + i >= nvalid) s2[i] = 0xff800000u;
            if (i < N3 && 96 + i >= nvalid) s3[i] = 0xff800000u;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        if constexpr (kFaVar<VAR>.dbg & 2) {
          mx0 = __uint_as_float(s0[0]);
        } else if constexpr (kFaVar<VAR>.max3) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx0 = max3f(mx0, __uint_as_float(s0[i]), __uint_as_float(s0[i + 1]));
            mx1 = max3f(mx1, __uint_as_float(s1[i]), __uint_as_float(s1[i + 1]));
            mx2 = max3f(mx2, __uint_as_float(s2[i]), __uint_as_float(s2[i + 1]));
            if (i < N3) mx3 = max3f(mx3, __uint_as_float(s3[i]), __uint_as_float(s3[i + 1]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            mx0 = fmaxf(mx0, __uint_as_float(s0[i]));
            mx1 = fmaxf(mx1, __uint_as_float(s1[i]));
            mx2 = fmaxf(mx2, __uint_as_float(s2[i]));
            if (i < N3) mx3 = fmaxf(mx3, __uint_as_float(s3[i]));
          }
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
        // lazy rescaling: keep the stale running max unless it grows by more than 2^8 (p stays <= 256, exact after the final 1/l)
        float m_new = m;
        const bool grow = mx > m + 8.0f;
        if (grow) m_new = mx;
        const bool warp_rescale = __any_sync(0xffffffffu, grow);
        // P(j) may only overwrite P(j-1) once PV(j-1) has consumed it; the same wait makes O safe to rescale
        auto wait_pv_and_rescale = [&]() {
          if (j > 0) {
            FA_TL(g, 4);
            mbar_wait(&pv_done[g], (t - 1) & 1);
            FA_TL(g, 5);
            tc_fence_after();
            if (warp_rescale) {
              const float alpha = ex2_approx(m - m_new);  // 1 for lanes whose max did not move
              l *= alpha;
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                uint32_t o[32];
                tmem_ld32(tmem_O + c * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st32(tmem_O + c * 32, o);
              }
              if constexpr (ONES) {  // the row-sum column (and its padding) is rescaled with O
                uint32_t o[16];
                tmem_ld16(tmem_O + 64, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st16(tmem_O + 64, o);
              }
              tmem_st_wait();
            }
          }
        };
        if constexpr (!kFaVar<VAR>.latewait) wait_pv_and_rescale();
        if constexpr (kFaVar<VAR>.token) {  // my turn on the XU pipe: after the other group's previous (g = 0) / same (g = 1) turn
          if (g == 1) mbar_wait(&tok[0], tn & 1);
          else if (tn > 0) mbar_wait(&tok[1], (tn - 1) & 1);
        }
        float sum0 = 0.0f, sum1 = 0.0f;
        f32x2_t sum2 = pk2(0.0f, 0.0f);
        const f32x2_t sc2 = pk2(sc, sc), nm2 = pk2(-m_new, -m_new);
        uint32_t pk[32];  // 64 probabilities packed to 16 bits = 32 TMEM columns
        auto emit = [&](const uint32_t (&sv)[32], int c) {
          if constexpr (kFaVar<VAR>.packed) {
#pragma unroll
            for (int q = 0; q < ((ONES && c == 3) ? 1 : 2); ++q) {  // 16 scores at a time (112-key tiles: the fourth block has 16)
              float pv[16];
#pragma unroll
              for (int i = 0; i < 16; i += 2)
                up2(fma2(pk2(__uint_as_float(sv[16 * q + i]), __uint_as_float(sv[16 * q + i + 1])), sc2, nm2), pv[i], pv[i + 1]);
              if constexpr (kFaVar<VAR>.dbg & 1) {
              } else if constexpr (kFaVar<VAR>.polypk) {
#pragma unroll
                for (int i = 0; i + 1 < kFaVar<VAR>.poly16; i += 2) exp2_poly_pair(pv[i], pv[i + 1]);
                if constexpr (kFaVar<VAR>.poly16 & 1) pv[kFaVar<VAR>.poly16 - 1] = exp2_poly(pv[kFaVar<VAR>.poly16 - 1]);
              } else {
#pragma unroll
                for (int i = 0; i < kFaVar<VAR>.poly16; ++i) pv[i] = exp2_poly(pv[i]);
              }
              if constexpr (!(kFaVar<VAR>.dbg & 1)) {
#pragma unroll
                for (int i = kFaVar<VAR>.poly16; i < 16; ++i) pv[i] = ex2_approx(pv[i]);
              }
              const f32x2_t a = add2(add2(pk2(pv[0], pv[1]), pk2(pv[2], pv[3])), add2(pk2(pv[4], pv[5]), pk2(pv[6], pv[7])));
              const f32x2_t b2 = add2(add2(pk2(pv[8], pv[9]), pk2(pv[10], pv[11])), add2(pk2(pv[12], pv[13]), pk2(pv[14], pv[15])));
              if constexpr (!ONES) sum2 = add2(sum2, add2(a, b2));  // ONES: the tensor core sums the rounded probabilities (O column 64)
#pragma unroll
              for (int i = 0; i < 8; ++i) pk[(c & 1) * 16 + q * 8 + i] = T16f<FMT>::pack(pv[2 * i], pv[2 * i + 1]);
            }
          } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pv[i] = fmaf(__uint_as_float(sv[8 * q + i]), sc, -m_new);
#pragma unroll
            for (int i = 0; i < 8; ++i) pv[i] = (i < FA_POLY) ? exp2_poly(pv[i]) : ex2_approx(pv[i]);
            sum0 += (pv[0] + pv[1]) + (pv[2] + pv[3]);
            sum1 += (pv[4] + pv[5]) + (pv[6] + pv[7]);
#pragma unroll
            for (int i = 0; i < 4; ++i) pk[(c & 1) * 16 + q * 4 + i] = T16f<FMT>::pack(pv[2 * i], pv[2 * i + 1]);
          }
          }
        };
        emit(s0, 0);
        emit(s1, 1);
        if constexpr (kFaVar<VAR>.latewait) wait_pv_and_rescale();
        tmem_st32(tmem_P, pk);
        emit(s2, 2);
        emit(s3, 3);
        if constexpr (kFaVar<VAR>.token) {
          mbar_arrive(&tok[g]);
          ++tn;
        }
        if constexpr (ONES) {  // 48 probabilities = 24 words
          tmem_st16(tmem_P + 32, reinterpret_cast<const uint32_t (&)[16]>(pk));
          tmem_st8(tmem_P + 48, reinterpret_cast<const uint32_t (&)[8]>(pk[16]));
        } else {
          tmem_st32(tmem_P + 32, pk);
        }
        m = m_new;
        if constexpr (kFaVar<VAR>.packed) up2(sum2, sum0, sum1);
        l += sum0 + sum1;
        FA_TL(g, 6);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[g]);
        FA_TL(g, 7);
      }
      mbar_wait(&pv_done[g], (t - 1) & 1);
      tc_fence_after();
      const int q_row = (q_tile0 + g) * FA_BM + r;
      const bool ok = q_row < tokens;
      if constexpr (ONES) {
        uint32_t v16[16];
        tmem_ld16(tmem_O + 64, v16);
        tmem_ld_wait();
        l = __uint_as_float(v16[0]);
      }
      const float inv = 1.0f / l;
      uint16_t* o = reinterpret_cast<uint16_t*>(out) + ((long long)b * tokens + q_row) * C + h * FA_D;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_O + c * 32, v);
        tmem_ld_wait();
        if (c == 1) {  // O is in registers: the issuer may start the next unit's accumulation
          tc_fence_before();
          mbar_arrive(&o_free[g]);
        }
        if (ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u4;
            u4.x = T16f<FMT>::pack(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
            u4.y = T16f<FMT>::pack(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
            u4.z = T16f<FMT>::pack(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
            u4.w = T16f<FMT>::pack(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
            *reinterpret_cast<uint4*>(o + c * 32 + i) = u4;
          }
        }
      }
    }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FA_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// temporal attention (T <= 32)
// ------------------------------------------------------------------------------------------------
constexpr int TA_WARPS = 4;

template <typename T>
__global__ void __launch_bounds__(TA_WARPS * 32)
temporal_attn_kernel(const T* __restrict__ qkv, T* __restrict__ out, int D, int Tn, int C, int heads) {
  extern __shared__ __align__(16) uint8_t ta_smem[];
  const int dh = C / heads;
  const int ldp = dh + 2;  // padded row (elements): odd number of 32-bit words -> conflict-free per-lane rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* sq = reinterpret_cast<T*>(ta_smem) + (size_t)warp * 3 * 32 * ldp;
  T* sk = sq + 32 * ldp;
  T* sv = sk + 32 * ldp;
  const float scale = rsqrtf((float)dh);
  const long long npairs = (long long)D * heads;
  const int vec_per_row = dh / 8;  // 16-byte vectors per head row

  for (long long pair = (long long)blockIdx.x * TA_WARPS + warp; pair < npairs; pair += (long long)gridDim.x * TA_WARPS) {
    const long long d = pair / heads;
    const int hd = int(pair - d * heads);
    const T* base = qkv + (d * Tn) * (3LL * C) + hd * dh;
    // cooperative load of the q, k, v tiles [Tn x dh]
    for (int idx = lane; idx < 3 * Tn * vec_per_row; idx += 32) {
      const int which = idx / (Tn * vec_per_row);
      const int rem = idx - which * Tn * vec_per_row;
      const int f = rem / vec_per_row, v8 = rem - f * vec_per_row;
      const uint4 u = *reinterpret_cast<const uint4*>(base + (long long)f * 3 * C + which * C + v8 * 8);
      uint32_t* dst = reinterpret_cast<uint32_t*>(sq + (which * 32 + f) * ldp + v8 * 8);
      dst[0] = u.x; dst[1] = u.y; dst[2] = u.z; dst[3] = u.w;
    }
    __syncwarp();
    float s[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) s[j] = 0.0f;
    const int f = lane < Tn ? lane : 0;
    const uint32_t* qrow = reinterpret_cast<const uint32_t*>(sq + f * ldp);
    for (int c2 = 0; c2 < dh / 2; ++c2) {
      const float2 q2 = T16<T>::unpack(qrow[c2]);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < Tn) {
          const float2 k2 = T16<T>::unpack(reinterpret_cast<const uint32_t*>(sk + j * ldp)[c2]);
          s[j] = fmaf(q2.x, k2.x, fmaf(q2.y, k2.y, s[j]));
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < Tn) mx = fmaxf(mx, s[j] * scale);
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s[j] = (j < Tn) ? __expf(s[j] * scale - mx) : 0.0f;
      sum += s[j];
    }
    const float inv = 1.0f / sum;
    T* orow = out + (d * Tn + f) * (long long)C + hd * dh;
    for (int c0 = 0; c0 < dh; c0 += 8) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < Tn) {
          const uint32_t* vr = reinterpret_cast<const uint32_t*>(sv + j * ldp + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 v2 = T16<T>::unpack(vr[i]);
            acc[2 * i] = fmaf(s[j], v2.x, acc[2 * i]);
            acc[2 * i + 1] = fmaf(s[j], v2.y, acc[2 * i + 1]);
          }
        }
      }
      if (lane < Tn) {
        uint4 u;
        u.x = T16<T>::pack(acc[0] * inv, acc[1] * inv);
        u.y = T16<T>::pack(acc[2] * inv, acc[3] * inv);
        u.z = T16<T>::pack(acc[4] * inv, acc[5] * inv);
        u.w = T16<T>::pack(acc[6] * inv, acc[7] * inv);
        *reinterpret_cast<uint4*>(orow + c0) = u;
      }
    }
    __syncwarp();
  }
}


// ------------------------------------------------------------------------------------------------
// streaming temporal attention: one new frame against <= 32 cached frames (video_depth_stream.py:76-160, motion_module.py:252-269)
// ------------------------------------------------------------------------------------------------
// The reference re-projects the concatenated [cached, current] hidden states (+ positional encoding) on every frame.  to_q / to_k /
// to_v have no bias, so W (n_j + pe_j) = W n_j + W pe_j: each frame's projection W n is computed once when the frame arrives and
// cached, and the position-dependent part is a weights-only table pos[L, 3C] = pe W^T.  This kernel then reads L cached k / v rows
// per pixel (HBM-bound): warp per (pixel, head), lanes over channels, entries in a loop.
struct StreamAttnParams {
  const uint16_t* qkv[32];  // per entry: [D, ld] 16-bit rows, q | k | v at column offsets 0, C, 2C
  const float* pos;         // [32, 3C] fp32: row j = (Wq pe_j | Wk pe_j | Wv pe_j)
  void* out;                // [D, C] 16-bit
  int D, C, heads, L, ld, fmt;
  // ring mode (table != nullptr): entry j lives in slot table[j] of `pool` ([slots, D, ld]) or, for table[j] < 0, in `staging`
  // (this frame's projection).  The slot table is DEVICE memory, so the launch parameters do not change from frame to frame and
  // the whole streaming step can be replayed as one CUDA graph.
  const uint16_t* pool;
  const uint16_t* staging;
  const int* table;
  long long slot_stride;
};

__device__ __forceinline__ const uint16_t* stream_entry(const StreamAttnParams& p, int j) {
  if (p.table == nullptr) return p.qkv[j];
  const int s = __ldg(p.table + j);
  return s < 0 ? p.staging : p.pool + (long long)s * p.slot_stride;
}

// ring mode: staging -> pool[table[slot_index]] (16-byte vectors), the cache insertion of the frame inside the replayed graph
__global__ void __launch_bounds__(256) ring_store_kernel(const uint4* __restrict__ staging, uint4* __restrict__ pool, const int* __restrict__ table,
                                                         int slot_index, long long nvec) {
  const int s = __ldg(table + slot_index);
  if (s < 0) return;
  uint4* dst = pool + (long long)s * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) dst[i] = staging[i];
}

__global__ void __launch_bounds__(256) stream_temporal_attn_kernel(const StreamAttnParams p) {
  const int lane = threadIdx.x & 31;
  const int dh = p.C / p.heads;
  const float scale = rsqrtf((float)dh);
  const long long npairs = (long long)p.D * p.heads;
  const int L = p.L;
  for (long long pair = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pair < npairs; pair += ((long long)gridDim.x * blockDim.x) >> 5) {
    const long long d = pair / p.heads;
    const int h = int(pair - d * p.heads);
    const int c0 = h * dh;
    // query = current frame (entry L-1) at position L-1
    float q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      q[i] = c < dh ? load16(stream_entry(p, L - 1), d * p.ld + c0 + c, p.fmt) + __ldg(p.pos + (long long)(L - 1) * 3 * p.C + c0 + c) : 0.0f;
    }
    float sj = -INFINITY;  // lane j keeps score j
    for (int j = 0; j < L; ++j) {
      float acc = 0.0f;
      const uint16_t* ej = stream_entry(p, j);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        if (c < dh) acc = fmaf(q[i], load16(ej, d * p.ld + p.C + c0 + c, p.fmt) + __ldg(p.pos + (long long)j * 3 * p.C + p.C + c0 + c), acc);
      }
      acc = warp_sum(acc) * scale;
      if (lane == j) sj = acc;
    }
    const float mx = warp_max(sj);
    const float e = lane < L ? __expf(sj - mx) : 0.0f;
    const float pj = e / warp_sum(e);
    float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int j = 0; j < L; ++j) {
      const float w = __shfl_sync(0xffffffffu, pj, j);
      const uint16_t* ej = stream_entry(p, j);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        if (c < dh) o[i] = fmaf(w, load16(ej, d * p.ld + 2 * p.C + c0 + c, p.fmt) + __ldg(p.pos + (long long)j * 3 * p.C + 2 * p.C + c0 + c), o[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      if (c < dh) store16(p.out, d * p.C + c0 + c, o[i], p.fmt);
    }
  }
}

// Vectorised form for head_dim 32 / 64 / 128 (ViT-L motion modules).  The scalar kernel above reads 2 bytes per lane per load and
// re-reads the fp32 positional table from global memory for every (pixel, head): 0.7 TB/s.  Here a block serves one head, keeps that
// head's positional k / v slices in shared memory, and each warp walks its (pixel, head) with 16-byte loads: LPE = DH/8 lanes share
// one cached entry, 32/LPE entries are in flight per warp; scores meet in shared memory, the output is reduced across the entry
// groups with shuffles.
template <int DH, int FMT>
__global__ void __launch_bounds__(256) stream_temporal_attn_vec_kernel(const StreamAttnParams p) {
  constexpr int LPE = DH / 8;   // lanes per entry (one 16-byte vector of 8 channels each)
  constexpr int G = 32 / LPE;   // entries in flight per warp
  extern __shared__ float sta_smem[];
  float* pk = sta_smem;                 // [32][DH]
  float* pv = sta_smem + 32 * DH;       // [32][DH]
  float* sc = sta_smem + 64 * DH;       // [8 warps][32]
  const int h = blockIdx.y;
  const int L = p.L, C = p.C;
  for (int i = threadIdx.x; i < L * DH; i += blockDim.x) {
    const int j = i / DH, c = i - j * DH;
    pk[i] = __ldg(p.pos + (long long)j * 3 * C + C + h * DH + c);
    pv[i] = __ldg(p.pos + (long long)j * 3 * C + 2 * C + h * DH + c);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / LPE, v8 = lane % LPE;
  const float scale = rsqrtf((float)DH);
  float* myscores = sc + warp * 32;
  const int c_off = h * DH + v8 * 8;
  for (int d = blockIdx.x * 8 + warp; d < p.D; d += gridDim.x * 8) {
    const long long row = (long long)d * p.ld;
    float q[8];
    {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(stream_entry(p, L - 1) + row + c_off));
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(L - 1) * 3 * C + c_off));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(L - 1) * 3 * C + c_off) + 1);
      float2 t;
      t = T16f<FMT>::unpack(u.x); q[0] = t.x + a.x; q[1] = t.y + a.y;
      t = T16f<FMT>::unpack(u.y); q[2] = t.x + a.z; q[3] = t.y + a.w;
      t = T16f<FMT>::unpack(u.z); q[4] = t.x + b.x; q[5] = t.y + b.y;
      t = T16f<FMT>::unpack(u.w); q[6] = t.x + b.z; q[7] = t.y + b.w;
    }
    // batches of 4 entry groups: the four 16-byte loads of a batch are issued before the first dot product (the loop is otherwise
    // one exposed global-load latency per group)
    for (int jb = 0; jb < L; jb += 4 * G) {
      uint4 u[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = jb + t * G + grp;
        u[t] = j < L ? __ldg(reinterpret_cast<const uint4*>(stream_entry(p, j) + row + C + c_off)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = jb + t * G + grp;
        const bool valid = j < L;
        float acc = 0.0f;
        if (valid) {
          const float4 a = *reinterpret_cast<const float4*>(pk + j * DH + v8 * 8);
          const float4 b = *reinterpret_cast<const float4*>(pk + j * DH + v8 * 8 + 4);
          float2 f;
          f = T16f<FMT>::unpack(u[t].x); acc = fmaf(q[0], f.x + a.x, acc); acc = fmaf(q[1], f.y + a.y, acc);
          f = T16f<FMT>::unpack(u[t].y); acc = fmaf(q[2], f.x + a.z, acc); acc = fmaf(q[3], f.y + a.w, acc);
          f = T16f<FMT>::unpack(u[t].z); acc = fmaf(q[4], f.x + b.x, acc); acc = fmaf(q[5], f.y + b.y, acc);
          f = T16f<FMT>::unpack(u[t].w); acc = fmaf(q[6], f.x + b.z, acc); acc = fmaf(q[7], f.y + b.w, acc);
        }
#pragma unroll
        for (int o = LPE / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (valid && v8 == 0) myscores[j] = acc * scale;
      }
    }
    __syncwarp();
    const float sj = lane < L ? myscores[lane] : -INFINITY;
    const float mx = warp_max(sj);
    const float e = lane < L ? __expf(sj - mx) : 0.0f;
    const float pj = e / warp_sum(e);
    float o8[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (int jb = 0; jb < L; jb += 4 * G) {
      uint4 u[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = jb + t * G + grp;
        u[t] = j < L ? __ldg(reinterpret_cast<const uint4*>(stream_entry(p, j) + row + 2 * C + c_off)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = jb + t * G + grp;
        const bool valid = j < L;
        const float w = __shfl_sync(0xffffffffu, pj, valid ? j : 0);
        if (valid) {
          const float4 a = *reinterpret_cast<const float4*>(pv + j * DH + v8 * 8);
          const float4 b = *reinterpret_cast<const float4*>(pv + j * DH + v8 * 8 + 4);
          float2 f;
          f = T16f<FMT>::unpack(u[t].x); o8[0] = fmaf(w, f.x + a.x, o8[0]); o8[1] = fmaf(w, f.y + a.y, o8[1]);
          f = T16f<FMT>::unpack(u[t].y); o8[2] = fmaf(w, f.x + a.z, o8[2]); o8[3] = fmaf(w, f.y + a.w, o8[3]);
          f = T16f<FMT>::unpack(u[t].z); o8[4] = fmaf(w, f.x + b.x, o8[4]); o8[5] = fmaf(w, f.y + b.y, o8[5]);
          f = T16f<FMT>::unpack(u[t].w); o8[6] = fmaf(w, f.x + b.z, o8[6]); o8[7] = fmaf(w, f.y + b.w, o8[7]);
        }
      }
    }
#pragma unroll
    for (int o = LPE; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] += __shfl_xor_sync(0xffffffffu, o8[i], o);
    }
    if (grp == 0) {
      uint4 r;
      r.x = T16f<FMT>::pack(o8[0], o8[1]); r.y = T16f<FMT>::pack(o8[2], o8[3]);
      r.z = T16f<FMT>::pack(o8[4], o8[5]); r.w = T16f<FMT>::pack(o8[6], o8[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + (long long)d * C + c_off) = r;
    }
    __syncwarp();  // myscores is rewritten by the next pixel
  }
}

template <int DH>
static int launch_stream_vec(const StreamAttnParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)(64 * DH + 8 * 32) * sizeof(float);
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured && smem > 48 * 1024) {
    cudaFuncSetAttribute(stream_temporal_attn_vec_kernel<DH, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(stream_temporal_attn_vec_kernel<DH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  int bx = (p.D + 7) / 8;
  const int cap = (num_sms() * 8 + p.heads - 1) / p.heads;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)bx, (unsigned)p.heads);
  if (p.fmt) stream_temporal_attn_vec_kernel<DH, 1><<<grid, 256, smem, stream>>>(p);
  else stream_temporal_attn_vec_kernel<DH, 0><<<grid, 256, smem, stream>>>(p);
  count_launch();
  return check_launch("stream_temporal_attn_vec_kernel");
}

// all entries 16-byte addressable (rows ld * 2 bytes, channel offsets multiples of 8) and a head_dim the vector kernel is built for
static bool stream_vec_ok(const StreamAttnParams& p) {
  const int dh = p.C / p.heads;
  static const char* env = getenv("VDN_STREAM_ATTN_V1");
  return env == nullptr && (dh == 32 || dh == 64 || dh == 128) && p.ld % 8 == 0 && p.C % 8 == 0;
}

static int launch_stream_attn(const StreamAttnParams& p, cudaStream_t stream) {
  if (stream_vec_ok(p)) {
    const int dh = p.C / p.heads;
    if (dh == 32) return launch_stream_vec<32>(p, stream);
    if (dh == 64) return launch_stream_vec<64>(p, stream);
    return launch_stream_vec<128>(p, stream);
  }
  long long blocks = ((long long)p.D * p.heads + 7) / 8;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  stream_temporal_attn_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
  count_launch();
  return check_launch("stream_temporal_attn_kernel");
}

// ------------------------------------------------------------------------------------------------
// temporal attention on tcgen05 (T == 32 frames, head_dim 32 / 64 / 128)
// ------------------------------------------------------------------------------------------------
// A tile is 128 consecutive rows of the pixel-major sequence = 4 pixels x 32 frames.  S = Q K^T is computed for the whole
// 128 x 128 tile (4x the needed FLOPs — irrelevant, the op is HBM-bound at 8 bytes per token-channel) and only its four
// 32 x 32 diagonal blocks are kept: warp w owns rows 32w..32w+31 and reads exactly columns 32w..32w+31 of S from TMEM,
// normalises them and writes them to its 16 packed columns of the block-diagonal P operand (all other columns stay zero);
// O = P V then runs as a TS-mode MMA against V^T (produced transposed per tile by the QKV GEMM epilogue).
// Persistent CTAs, 2-stage TMA ring: the loads of unit i+1 run under the MMAs / softmax / stores of unit i.
constexpr int TT_THREADS = 192;  // 4 softmax warps, MMA issuer, TMA producer

template <int DH, int FMT>
__global__ void __launch_bounds__(TT_THREADS, 1)
temporal_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVT, void* __restrict__ out, long long rows,
                        int C, int num_tiles) {
  constexpr int CU = DH < 64 ? 64 : DH;        // channels per work unit (one 64-column TMA box minimum)
  constexpr int HPU = CU / DH;                 // heads per unit (2 for head_dim 32)
  constexpr int NB = CU / 64;                  // 64-column boxes per unit
  constexpr int QBYTES = NB * 128 * 128;       // Q (and K) bytes per stage
  constexpr int VBYTES = 2 * CU * 128;         // V^T: two 64-key chunks of [CU rows x 128 B]
  constexpr int STAGE = 2 * QBYTES + VBYTES;
  constexpr uint32_t S_COL = 0, P_COL = HPU * 128, O_COL = HPU * 128 + HPU * 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  uint64_t* full = bars;        // [2]
  uint64_t* empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* o_free = bars + 7;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units_per_tile = C / CU;
  const long long num_units = (long long)num_tiles * units_per_tile;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp_idx == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQK);
      tma_prefetch_desc(&tmVT);
      mbar_init(&full[0], 1); mbar_init(&full[1], 1);
      mbar_init(&empty[0], 1); mbar_init(&empty[1], 1);
      mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(o_free, 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 5) {
    // ---------------- TMA producer ----------------
    int it = 0;
    for (long long u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int tile = int(u / units_per_tile), cu = int(u % units_per_tile);
      mbar_wait(&empty[st], ph ^ 1);
      if (elect_one()) {
        uint8_t* sQ = smem + st * STAGE;
        uint8_t* sK = sQ + QBYTES;
        uint8_t* sV = sK + QBYTES;
        mbar_arrive_expect_tx(&full[st], STAGE);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          tma_load_2d(sQ + b * 16384, &tmQK, &full[st], cu * CU + b * 64, tile * 128);
          tma_load_2d(sK + b * 16384, &tmQK, &full[st], C + cu * CU + b * 64, tile * 128);
        }
        tma_load_3d(sV, &tmVT, &full[st], 0, cu * CU, tile);
        tma_load_3d(sV + CU * 128, &tmVT, &full[st], 64, cu * CU, tile);
      }
      __syncwarp();
    }
  } else if (warp_idx == 4) {
    // ---------------- MMA issuer ----------------
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc_s = make_idesc(FMT ? 1u : 0u, 128, 128);
    constexpr uint32_t idesc_pv = make_idesc(FMT ? 1u : 0u, 128, DH);
    const uint32_t smem_addr = smem_u32(smem);
    auto issue_s = [&](int st) {
      if (elect_one()) {
        const uint32_t q_addr = smem_addr + st * STAGE, k_addr = q_addr + QBYTES;
#pragma unroll
        for (int hh = 0; hh < HPU; ++hh) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) {
            const int col = hh * DH + kk * 16;  // column inside the unit
            const uint64_t dq = make_sdesc_sw128(q_addr + (col >> 6) * 16384) + 2 * ((col & 63) >> 4);
            const uint64_t dk = make_sdesc_sw128(k_addr + (col >> 6) * 16384) + 2 * ((col & 63) >> 4);
            umma_f16(tb + S_COL + hh * 128, dq, dk, idesc_s, kk != 0 ? 1u : 0u);
          }
        }
        umma_commit(s_full);
      }
      __syncwarp();
    };
    int it = 0;
    const long long first = blockIdx.x;
    if (first < num_units) {
      mbar_wait(&full[0], 0);
      tc_fence_after();
      issue_s(0);
    }
    for (long long u = first; u < num_units; u += gridDim.x, ++it) {
      const int st = it & 1;
      mbar_wait(p_full, it & 1);                     // P(it) written, S(it) consumed
      if (it > 0) mbar_wait(o_free, (it - 1) & 1);   // O(it-1) drained by the epilogue
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v_addr = smem_addr + st * STAGE + 2 * QBYTES;
#pragma unroll
        for (int hh = 0; hh < HPU; ++hh) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // 128 keys, 16 per MMA; B = V^T rows [hh*DH, +DH) of chunk kk >> 2
            const uint64_t dv = make_sdesc_sw128(v_addr + (kk >> 2) * (CU * 128) + hh * DH * 128) + 2 * (kk & 3);
            umma_f16_ts(tb + O_COL + hh * DH, tb + P_COL + hh * 64 + kk * 8, dv, idesc_pv, kk != 0 ? 1u : 0u);
          }
        }
        umma_commit(o_full);
        umma_commit(&empty[st]);
      }
      __syncwarp();
      if (u + gridDim.x < num_units) {  // S of the next unit runs under this unit's epilogue
        const int it2 = it + 1;
        mbar_wait(&full[it2 & 1], (it2 >> 1) & 1);
        tc_fence_after();
        issue_s(it2 & 1);
      }
    }
  } else {
    // ---------------- softmax + output warps ----------------
    const uint32_t lane_off = uint32_t(warp_idx * 32) << 16;
    const int r = warp_idx * 32 + lane;
    const float sc = rsqrtf((float)DH) * 1.4426950408889634f;
    // zero the P operand once: every warp only ever rewrites its own diagonal block
    {
      uint32_t z[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) z[i] = 0u;
#pragma unroll
      for (int c = 0; c < HPU * 2; ++c) tmem_st32(tmem_base + P_COL + lane_off + c * 32, z);
      tmem_st_wait();
    }
    int it = 0;
    for (long long u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
      const int tile = int(u / units_per_tile), cu = int(u % units_per_tile);
      mbar_wait(s_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < HPU; ++hh) {
        uint32_t sv[32];
        tmem_ld32(tmem_base + S_COL + hh * 128 + lane_off + warp_idx * 32, sv);  // this row's 32 keys = its own pixel's frames
        tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
        mx *= sc;
        float p[32], sum = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          p[i] = ex2_approx(fmaf(__uint_as_float(sv[i]), sc, -mx));
          sum += p[i];
        }
        const float inv = 1.0f / sum;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = T16f<FMT>::pack(p[2 * i] * inv, p[2 * i + 1] * inv);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                tmem_base + P_COL + hh * 64 + lane_off + warp_idx * 16),
            "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]), "r"(pk[8]), "r"(pk[9]), "r"(pk[10]),
            "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
            : "memory");
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      // output: O [128 x CU] fp32 in TMEM -> 16-bit rows of `out`
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      const long long row = (long long)tile * 128 + r;
      uint16_t* o = reinterpret_cast<uint16_t*>(out) + row * C + cu * CU;
#pragma unroll
      for (int c = 0; c < CU / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + O_COL + lane_off + c * 32, v);
        tmem_ld_wait();
        if (row < rows) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 q4;
            q4.x = T16f<FMT>::pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            q4.y = T16f<FMT>::pack(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            q4.z = T16f<FMT>::pack(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
            q4.w = T16f<FMT>::pack(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
            *reinterpret_cast<uint4*>(o + c * 32 + i) = q4;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(o_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH, int FMT>
static int launch_temporal_tc(const CUtensorMap& tmQK, const CUtensorMap& tmVT, void* out, long long rows, int C, int num_tiles, cudaStream_t stream) {
  constexpr int CU = DH < 64 ? 64 : DH;
  constexpr int STAGE = 2 * (CU / 64) * 16384 + 2 * CU * 128;
  constexpr int SMEM = 2 * STAGE + 256;
  static bool configured_dev[kMaxDevices] = {};  // function attributes are per device
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(temporal_attn_tc_kernel<DH, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(temporal_attn_tc): ") + cudaGetErrorString(e));
    configured = true;
  }
  const long long units = (long long)num_tiles * (C / CU);
  const int grid = units < num_sms() ? (int)units : num_sms();
  temporal_attn_tc_kernel<DH, FMT><<<grid, TT_THREADS, SMEM, stream>>>(tmQK, tmVT, out, rows, C, num_tiles);
  count_launch();
  return check_launch("temporal_attn_tc_kernel");
}

}  // namespace vdn

using namespace vdn;

namespace vdn {
template <int FMT, int VAR>
static int launch_flash_one(int grid, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmVT, void* out, int tokens_q,
                            int tokens_kv, int heads, int C, int units) {
  static bool configured_dev[kMaxDevices] = {};  // function attributes are per device
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_kernel<FMT, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(flash_attn): ") + cudaGetErrorString(e));
    configured = true;
  }
  flash_attn_kernel<FMT, VAR><<<grid, FA_THREADS, FA_SMEM, stream>>>(tmQ, tmK, tmVT, out, tokens_q, tokens_kv, heads, C, units);
  count_launch();
  return check_launch("flash_attn_kernel");
}
template <int VAR>
static int launch_flash_fmt(int fmt, int grid, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmVT, void* out,
                            int tokens_q, int tokens_kv, int heads, int C, int units) {
  return fmt ? launch_flash_one<1, VAR>(grid, stream, tmQ, tmK, tmVT, out, tokens_q, tokens_kv, heads, C, units)
             : launch_flash_one<0, VAR>(grid, stream, tmQ, tmK, tmVT, out, tokens_q, tokens_kv, heads, C, units);
}
static int launch_flash_variant(int variant, int fmt, int grid, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmVT,
                                void* out, int tokens_q, int tokens_kv, int heads, int C, int units) {
#define VDN_FA_CASE(V) case V: return launch_flash_fmt<V>(fmt, grid, stream, tmQ, tmK, tmVT, out, tokens_q, tokens_kv, heads, C, units)
  switch (variant) {
    VDN_FA_CASE(0); VDN_FA_CASE(1); VDN_FA_CASE(2); VDN_FA_CASE(3); VDN_FA_CASE(4); VDN_FA_CASE(5); VDN_FA_CASE(6); VDN_FA_CASE(7); VDN_FA_CASE(8);
    VDN_FA_CASE(9); VDN_FA_CASE(10); VDN_FA_CASE(11); VDN_FA_CASE(12); VDN_FA_CASE(13); VDN_FA_CASE(14); VDN_FA_CASE(15); VDN_FA_CASE(16);
    VDN_FA_CASE(17); VDN_FA_CASE(18); VDN_FA_CASE(19);
    default: return launch_flash_fmt<20>(fmt, grid, stream, tmQ, tmK, tmVT, out, tokens_q, tokens_kv, heads, C, units);
  }
#undef VDN_FA_CASE
}
}  // namespace vdn

#ifdef VDN_FA_TIMELINE
extern "C" int vdn_debug_fa_timeline(unsigned long long* host, int* counts) {
  int zero[5] = {0, 0, 0, 0, 0};
  cudaDeviceSynchronize();
  if (host == nullptr) return cudaMemcpyToSymbol(vdn::g_fa_tl_n, zero, sizeof(zero)) != cudaSuccess;  // reset
  if (cudaMemcpyFromSymbol(counts, vdn::g_fa_tl_n, sizeof(zero)) != cudaSuccess) return 1;
  return cudaMemcpyFromSymbol(host, vdn::g_fa_tl, sizeof(unsigned long long) * 5 * vdn::FA_TL_CAP) != cudaSuccess;
}
#endif

extern "C" int vdn_flash_attn_ex(const void* q, int64_t ld_q, int64_t q_batch_stride, const void* k, int64_t ld_k, int64_t k_batch_stride,
                                 const void* vT, int64_t ld_vT, void* out, int32_t B, int32_t tokens_q, int32_t tokens_kv, int32_t heads,
                                 void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!q || !k || !vT || !out) return set_error("vdn_flash_attn: null pointer");
  if (B <= 0 || tokens_q <= 0 || tokens_kv <= 0 || heads <= 0) return set_error("vdn_flash_attn: bad shape");
  const int C = heads * FA_D;
  if (ld_q < C || (ld_q * 2) % 16 != 0 || ld_k < C || (ld_k * 2) % 16 != 0) return set_error("vdn_flash_attn: ld_q / ld_k must be >= C and 16-byte aligned");
  if ((q_batch_stride * 2) % 16 != 0 || (k_batch_stride * 2) % 16 != 0) return set_error("vdn_flash_attn: batch strides must be 16-byte aligned");
  if (ld_vT < tokens_kv || (ld_vT * 2) % 16 != 0) return set_error("vdn_flash_attn: ld_vT must be >= tokens_kv and a multiple of 8");
  const int fmt = get_operand_format();
  static const int variant = [] {
    const char* env = getenv("VDN_FA_VARIANT");
    const int v = env ? atoi(env) : VDN_FA_DEFAULT_VARIANT;
    return v < 0 || v >= FA_NUM_VARIANTS ? VDN_FA_DEFAULT_VARIANT : v;
  }();
  const int bn = fa_variant(variant).ones ? 112 : FA_BN;  // keys per tile of the chosen kernel
  CUtensorMap tmQ, tmK, tmVT;
  {
    // element (b, t, h, d); innermost first: d, h, t, b
    const uint64_t dims[4] = {(uint64_t)FA_D, (uint64_t)heads, (uint64_t)tokens_q, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)FA_D * 2, (uint64_t)ld_q * 2, (uint64_t)q_batch_stride * 2};
    const uint32_t box[4] = {(uint32_t)FA_D, 1, (uint32_t)FA_BM, 1};
    if (make_tensor_map(&tmQ, q, fmt, 4, dims, strides, box)) return 1;
  }
  {
    const uint64_t dims[4] = {(uint64_t)FA_D, (uint64_t)heads, (uint64_t)tokens_kv, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)FA_D * 2, (uint64_t)ld_k * 2, (uint64_t)k_batch_stride * 2};
    const uint32_t box[4] = {(uint32_t)FA_D, 1, (uint32_t)bn, 1};
    if (make_tensor_map(&tmK, k, fmt, 4, dims, strides, box)) return 1;
  }
  {
    // V^T: (bh, d, t) with t contiguous; innermost first: t, d, bh.  Columns >= tokens_kv are out of bounds -> zero filled.
    const uint64_t dims[3] = {(uint64_t)tokens_kv, (uint64_t)FA_D, (uint64_t)B * heads};
    const uint64_t strides[2] = {(uint64_t)ld_vT * 2, (uint64_t)ld_vT * 2 * FA_D};
    const uint32_t box[3] = {64, (uint32_t)FA_D, 1};
    if (make_tensor_map(&tmVT, vT, fmt, 3, dims, strides, box)) return 1;
  }
  const long long units = (long long)((tokens_q + FA_GROUPS * FA_BM - 1) / (FA_GROUPS * FA_BM)) * heads * B;
  if (units > 0x7fffffffLL) return set_error("vdn_flash_attn: too many work units");
  const int grid = units < num_sms() ? (int)units : num_sms();
  return launch_flash_variant(variant, fmt, grid, stream, tmQ, tmK, tmVT, out, tokens_q, tokens_kv, heads, C, (int)units);
}

extern "C" int vdn_flash_attn(const void* qk, int64_t ld_qk, const void* vT, int64_t ld_vT, void* out, int32_t B, int32_t tokens, int32_t heads,
                              void* stream_v) {
  if (!qk) return set_error("vdn_flash_attn: null pointer");
  const int C = heads * FA_D;
  if (ld_qk < 2 * C) return set_error("vdn_flash_attn: ld_qk must be >= 2*C and 16-byte aligned");
  const uint16_t* base = reinterpret_cast<const uint16_t*>(qk);
  return vdn_flash_attn_ex(base, ld_qk, (int64_t)tokens * ld_qk, base + C, ld_qk, (int64_t)tokens * ld_qk, vT, ld_vT, out, B, tokens, tokens, heads, stream_v);
}

extern "C" int vdn_stream_temporal_attn(const void* const* qkv_entries, int32_t L, int64_t ld, const float* pos, void* out, int32_t D, int32_t C,
                                        int32_t heads, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!qkv_entries || !pos || !out) return set_error("vdn_stream_temporal_attn: null pointer");
  if (L < 1 || L > 32) return set_error("vdn_stream_temporal_attn: L must be in [1, 32]");
  if (heads <= 0 || C % heads != 0 || C / heads > 128 || ld < 3 * C) return set_error("vdn_stream_temporal_attn: head_dim must be <= 128 and ld >= 3*C");
  StreamAttnParams p{};
  for (int j = 0; j < L; ++j) {
    if (!qkv_entries[j]) return set_error("vdn_stream_temporal_attn: null entry");
    p.qkv[j] = reinterpret_cast<const uint16_t*>(qkv_entries[j]);
  }
  p.pos = pos; p.out = out; p.D = D; p.C = C; p.heads = heads; p.L = L; p.ld = (int)ld; p.fmt = get_operand_format();
  for (int j = 0; j < L; ++j)
    if ((reinterpret_cast<uintptr_t>(p.qkv[j]) & 15) != 0) return set_error("vdn_stream_temporal_attn: entries must be 16-byte aligned");
  return launch_stream_attn(p, stream);
}

extern "C" int vdn_stream_temporal_attn_ring(const void* pool, const void* staging, const int32_t* slot_table, int32_t L, int64_t ld, const float* pos,
                                             void* out, int32_t D, int32_t C, int32_t heads, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!pool || !staging || !slot_table || !pos || !out) return set_error("vdn_stream_temporal_attn_ring: null pointer");
  if (L < 1 || L > 32) return set_error("vdn_stream_temporal_attn_ring: L must be in [1, 32]");
  if (heads <= 0 || C % heads != 0 || C / heads > 128 || ld < 3 * C) return set_error("vdn_stream_temporal_attn_ring: head_dim must be <= 128 and ld >= 3*C");
  StreamAttnParams p{};
  p.pos = pos; p.out = out; p.D = D; p.C = C; p.heads = heads; p.L = L; p.ld = (int)ld; p.fmt = get_operand_format();
  p.pool = reinterpret_cast<const uint16_t*>(pool);
  p.staging = reinterpret_cast<const uint16_t*>(staging);
  p.table = slot_table;
  p.slot_stride = (long long)D * ld;
  if (((reinterpret_cast<uintptr_t>(pool) | reinterpret_cast<uintptr_t>(staging)) & 15) != 0) return set_error("vdn_stream_temporal_attn_ring: buffers must be 16-byte aligned");
  return launch_stream_attn(p, stream);
}

extern "C" int vdn_ring_store(const void* staging, void* pool, const int32_t* slot_table, int32_t slot_index, int64_t slot_elems, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!staging || !pool || !slot_table) return set_error("vdn_ring_store: null pointer");
  if (slot_elems <= 0 || slot_elems % 8 != 0 || slot_index < 0) return set_error("vdn_ring_store: slot_elems must be a positive multiple of 8");
  const long long nvec = slot_elems / 8;
  long long blocks = (nvec + 255) / 256;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  ring_store_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(staging), reinterpret_cast<uint4*>(pool), slot_table, slot_index, nvec);
  count_launch();
  return check_launch("ring_store_kernel");
}

extern "C" int vdn_temporal_attn_tc(const void* qk, int64_t ld_qk, const void* vT, void* out, int64_t rows, int32_t C, int32_t heads, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!qk || !vT || !out) return set_error("vdn_temporal_attn_tc: null pointer");
  if (rows <= 0 || rows % 32 != 0 || heads <= 0 || C % heads != 0) return set_error("vdn_temporal_attn_tc: rows must be a multiple of T = 32");
  const int dh = C / heads;
  if (dh != 32 && dh != 64 && dh != 128) return set_error("vdn_temporal_attn_tc: head_dim must be 32, 64 or 128");
  if (C % 64 != 0 || ld_qk < 2 * C || (ld_qk * 2) % 16 != 0) return set_error("vdn_temporal_attn_tc: bad C / ld_qk");
  const int fmt = get_operand_format();
  const int num_tiles = (int)((rows + 127) / 128);
  CUtensorMap tmQK, tmVT;
  {
    const uint64_t dims[2] = {(uint64_t)(2 * C), (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)ld_qk * 2};
    const uint32_t box[2] = {64u, 128u};
    if (make_tensor_map(&tmQK, qk, fmt, 2, dims, strides, box)) return 1;
  }
  {
    // V^T per tile: [tile][channel][128 keys]; innermost first: key, channel, tile
    const int cu = dh < 64 ? 64 : dh;
    const uint64_t dims[3] = {128u, (uint64_t)C, (uint64_t)num_tiles};
    const uint64_t strides[2] = {256u, (uint64_t)C * 256u};
    const uint32_t box[3] = {64u, (uint32_t)cu, 1u};
    if (make_tensor_map(&tmVT, vT, fmt, 3, dims, strides, box)) return 1;
  }
#define VDN_TT(DH) (fmt ? launch_temporal_tc<DH, 1>(tmQK, tmVT, out, rows, C, num_tiles, stream) : launch_temporal_tc<DH, 0>(tmQK, tmVT, out, rows, C, num_tiles, stream))
  if (dh == 32) return VDN_TT(32);
  if (dh == 64) return VDN_TT(64);
  return VDN_TT(128);
#undef VDN_TT
}

extern "C" int vdn_temporal_attn(const void* qkv, void* out, int32_t D, int32_t T, int32_t C, int32_t heads, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!qkv || !out) return set_error("vdn_temporal_attn: null pointer");
  if (T <= 0 || T > 32) return set_error("vdn_temporal_attn: T must be in [1, 32]");
  if (heads <= 0 || C % heads != 0 || (C / heads) % 8 != 0) return set_error("vdn_temporal_attn: head_dim must be a multiple of 8");
  const int dh = C / heads;
  const size_t smem = (size_t)TA_WARPS * 3 * 32 * (dh + 2) * 2;
  const long long npairs = (long long)D * heads;
  long long blocks = (npairs + TA_WARPS - 1) / TA_WARPS;
  const long long max_blocks = (long long)num_sms() * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  const int fmt = get_operand_format();
  cudaError_t e;
  if (fmt) {
    e = cudaFuncSetAttribute(temporal_attn_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(std::string("temporal_attn smem: ") + cudaGetErrorString(e));
    temporal_attn_kernel<__nv_bfloat16><<<(unsigned)blocks, TA_WARPS * 32, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                                                          reinterpret_cast<__nv_bfloat16*>(out), D, T, C, heads);
  } else {
    e = cudaFuncSetAttribute(temporal_attn_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(std::string("temporal_attn smem: ") + cudaGetErrorString(e));
    temporal_attn_kernel<__half><<<(unsigned)blocks, TA_WARPS * 32, smem, stream>>>(reinterpret_cast<const __half*>(qkv), reinterpret_cast<__half*>(out),
                                                                                 D, T, C, heads);
  }
  count_launch();
  return check_launch("temporal_attn_kernel");
}

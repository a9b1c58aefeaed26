// Head tail of the DPT decoder in one kernel (dpt_temporal.py:103-111 / dpt.py:179-187):
//
//     bilinear (align_corners) resize of output_conv1's map to (14 ph, 14 pw)  ->  output_conv2 = 3x3 conv 128 -> 32, ReLU, 1x1 conv 32 -> 1, ReLU
//
// The 128-channel full-resolution map (2.2 GB for a 32-frame 518x518 window) never exists: the upsampled pixels are produced straight
// into the shared-memory A operand of the convolution.  The convolution itself is organised around its narrow output (32 channels):
//
//   * an A tile is ONE image row segment: 128 MMA rows = 4 warps x 32 consecutive pixels, neighbouring warps overlapping by two
//     pixels (lane 0 and lane 31 of a warp are halo pixels), i.e. 120 output pixels per tile;
//   * the three HORIZONTAL taps go into the N dimension: E[x'][(dx, c)] = sum_k A[x'][k] W[c][k][dy][dx], N = 3 x 32 = 96, and the
//     epilogue combines them across lanes, out[x][c] = E[x-1][(0,c)] + E[x][(1,c)] + E[x+1][(2,c)] (two shuffles per channel);
//   * the three VERTICAL taps are three accumulating groups of MMAs over the row tiles y-1, y, y+1, which sit in a ring of four:
//     every row tile is built once and read by three output rows.
//   An MMA of 128 x 96 x 16 reads 4 KB of A and 3 KB of B from shared memory in 48 tensor cycles: 24 MMAs and 168 KB per output row
//   tile, against 72 MMAs of 128 x 32 x 16 and 360 KB for the 9-tap implicit GEMM it replaces (which ran at 0.27 of the tensor peak,
//   bound by shared-memory reads of the A operand).
//   * bias + ReLU + the 1x1 convolution + ReLU are the epilogue: one fp32 depth value per pixel leaves the SM.
//
// Warp roles (512 threads, registers re-divided per warpgroup with setmaxnreg): warp 0 TMA producer (unfused form only), warp 1 MMA
// issuer, warps 2-3 idle (56 registers); warps 4-7 epilogue (TMEM lane quarter = warp & 3; 112); warps 8-15 upsample producers
// (fused form; 168: two source rows of the next row tile in flight + the blended columns of the current one).
// Persistent over (frame, 120-pixel strip, chunk of rows) units.
#include <stdlib.h>

#include "../../include/vdn_b200.h"
#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int CT_C = 128;                      // input channels
constexpr int CT_NO = 32;                      // channels of the 3x3 convolution
constexpr int CT_N = 3 * CT_NO;                // MMA N: (dx, c)
constexpr int CT_VALID = 30;                   // output pixels per warp
constexpr int CT_STRIP = 4 * CT_VALID;         // output pixels per row tile
constexpr int CT_ATILE = 128 * CT_C * 2;       // 32 KB: [2 channel chunks][128 rows][128 B], 128B-swizzled
constexpr int CT_SLOTS = 4;                    // ring of row tiles
constexpr int CT_WCHUNK = CT_N * 128;          // 12 KB: one (dy, channel chunk) B tile [96 rows][128 B]
constexpr int CT_WBYTES = 3 * 2 * CT_WCHUNK;   // 72 KB
constexpr int CT_PRODUCERS = 8;                // upsample producer warps
constexpr int CT_COLS = 6;                     // source columns under a run of 8 output pixels (scale <= 4/7)
constexpr int CT_THREADS = (8 + CT_PRODUCERS) * 32;
constexpr int CT_TMEM_COLS = 256;              // two accumulators of 96 columns, 128 apart
constexpr int CT_SMEM = CT_SLOTS * CT_ATILE + CT_WBYTES + 1024;

struct TailParams {
  const uint4* src;      // fused form: output_conv1's map [B, Hs, Ws, 128], 16-bit
  int Hs, Ws;
  const uint4* wpacked;  // CT_WBYTES: pre-swizzled B tiles (packing.py::pack_conv_tail)
  const float* bias;     // [32]
  const float* head_w;   // [32]
  float head_b;
  float* out;            // [B, H, W] fp32
  int B, H, W;
  int strips, chunks, rows_per_chunk, num_units;
};

__device__ __forceinline__ void ct_coords(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {  // = ac_coords of vdn_elem.cu
  const float src = scale * (float)dst;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}
__device__ __forceinline__ float2 ct_fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  float2 c;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(rc));
  return c;
}
__device__ __forceinline__ float2 ct_ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

template <int FMT, bool FUSED>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tail_kernel(const __grid_constant__ CUtensorMap tmA, const TailParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                              // [slot]
  uint8_t* sW = smem + CT_SLOTS * CT_ATILE;        // [dy][chunk]
  uint8_t* tail = sW + CT_WBYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);  // [4]
  uint64_t* a_empty = a_full + 4;                  // [4]
  uint64_t* acc_full = a_full + 8;                 // [2]
  uint64_t* acc_empty = a_full + 10;               // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(a_full + 12);
  float* s_bias = reinterpret_cast<float*>(tail + 128);    // [32]
  float* s_hw = s_bias + 32;                               // [32]

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp_idx == 0) {
    if (lane == 0) {
      if (!FUSED) tma_prefetch_desc(&tmA);
      for (int i = 0; i < CT_SLOTS; ++i) {
        mbar_init(&a_full[i], FUSED ? CT_PRODUCERS * 32 : 1);
        mbar_init(&a_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], 128);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, CT_TMEM_COLS);
  }
  // the weights stay resident: every thread copies its share (generic proxy), made visible to the tensor core below
  for (int i = threadIdx.x; i < CT_WBYTES / 16; i += CT_THREADS) reinterpret_cast<uint4*>(sW)[i] = __ldg(p.wpacked + i);
  if (threadIdx.x < 32) {
    s_bias[threadIdx.x] = p.bias[threadIdx.x];
    s_hw[threadIdx.x] = p.head_w[threadIdx.x];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto unit_of = [&](int u, int& b, int& xs, int& y0, int& R) {
    const int chunk = u % p.chunks, strip = (u / p.chunks) % p.strips;
    b = u / (p.chunks * p.strips);
    xs = strip * CT_STRIP;
    y0 = chunk * p.rows_per_chunk;
    R = min(p.rows_per_chunk, p.H - y0);
  };

  // (each setmaxnreg sits inside the branch it governs: ptxas allocates a branch under its own limit only then)
  if (warp_idx < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
  if (warp_idx == 0) {
    if constexpr (!FUSED) {
      // ---------------- TMA producer: row tile = image row y0 - 1 + i, 4 x 32 pixels starting at xs + 30 w - 1, two channel chunks ----
      int cnt = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        int b, xs, y0, R;
        unit_of(u, b, xs, y0, R);
        for (int i = 0; i < R + 2; ++i, ++cnt) {
          const int slot = cnt % CT_SLOTS;
          mbar_wait(&a_empty[slot], ((cnt / CT_SLOTS) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[slot], CT_ATILE);
            for (int kc = 0; kc < 2; ++kc)
              for (int w = 0; w < 4; ++w)  // rows / columns outside the image are zero-filled: the convolution's padding
                tma_load_4d(sA + slot * CT_ATILE + kc * (CT_ATILE / 2) + w * 4096, &tmA, &a_full[slot], kc * 64, xs + CT_VALID * w - 1, y0 - 1 + i, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp_idx == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_idesc(FMT ? 1u : 0u, 128, CT_N);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
    int base = 0;   // row tiles of earlier units
    int local = 0;  // output rows so far (accumulator ring)
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      int b, xs, y0, R;
      unit_of(u, b, xs, y0, R);
      for (int j = 0; j < R; ++j, ++local) {
        const int acc = local & 1;
        mbar_wait(&acc_empty[acc], ((local >> 1) & 1) ^ 1);
        for (int dy = 0; dy < 3; ++dy) {
          const int gidx = base + j + dy;
          const int slot = gidx % CT_SLOTS;
          mbar_wait(&a_full[slot], (gidx / CT_SLOTS) & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
              const uint64_t da = make_sdesc_sw128(a_addr + slot * CT_ATILE + kc * (CT_ATILE / 2));
              const uint64_t db = make_sdesc_sw128(w_addr + (dy * 2 + kc) * CT_WCHUNK);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(tb + acc * 128, da + 2 * kk, db + 2 * kk, idesc, (dy | kc | kk) != 0 ? 1u : 0u);
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(&acc_full[acc]);
          umma_commit(&a_empty[(base + j) % CT_SLOTS]);  // row tile j was last needed by output row j
          if (j == R - 1) {
            umma_commit(&a_empty[(base + R) % CT_SLOTS]);
            umma_commit(&a_empty[(base + R + 1) % CT_SLOTS]);
          }
        }
        __syncwarp();
      }
      base += R + 2;
    }
  }
  } else if (warp_idx < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;" ::: "memory");
    // ---------------- epilogue: horizontal tap sum across lanes, bias, ReLU, 1x1 conv, ReLU ----------------
    const int q = warp_idx & 3;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    int local = 0;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      int b, xs, y0, R;
      unit_of(u, b, xs, y0, R);
      const int x = xs + CT_VALID * q + lane - 1;
      const bool store = lane >= 1 && lane <= CT_VALID && x < p.W;
      float* orow = p.out + ((long long)b * p.H + y0) * p.W + x;
      for (int j = 0; j < R; ++j, ++local) {
        const int acc = local & 1;
        mbar_wait_relaxed(&acc_full[acc], (local >> 1) & 1, 400);
        tc_fence_after();
        const uint32_t t = tmem_base + acc * 128 + lane_off;
        uint32_t e[32];
        float v[32];
        tmem_ld32(t + 32, e);  // dx = 1: this pixel
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(e[c]) + s_bias[c];
        tmem_ld32(t + 0, e);   // dx = 0: the contribution computed at pixel x - 1
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] += __shfl_up_sync(0xffffffffu, __uint_as_float(e[c]), 1);
        tmem_ld32(t + 64, e);  // dx = 2: at pixel x + 1
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[acc]);
        float d = p.head_b;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float s = v[c] + __shfl_down_sync(0xffffffffu, __uint_as_float(e[c]), 1);
          d = fmaf(fmaxf(s, 0.0f), s_hw[c], d);
        }
        if (store) orow[(long long)j * p.W] = fmaxf(d, 0.0f);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;" ::: "memory");
    if constexpr (FUSED) {
    // ---------------- upsample producers: 256 threads = 16 channel vectors x 16 runs of 8 consecutive pixels ----------------
    // Same arithmetic as bilinear_slide_kernel (vertical lerp of the source columns first, packed fp32), so the A operand is
    // bit-identical to what the unfused path reads back from HBM.  A run of 8 output pixels lies between at most CT_COLS source
    // columns (upsampling by >= 1.75, checked on the host); the 2 x CT_COLS 16-byte loads of the NEXT row tile are issued before
    // the current one is blended and stored, so no load latency is exposed inside a unit (a thread that loaded each column when
    // it slid onto it spent ~700 cycles per column: the fused kernel was 4x slower than resize + convolution).
    const int pt = threadIdx.x - 8 * 32;
    const int v = pt & 15, run = pt >> 4;
    const int r0 = run * 8;                       // first MMA row of the run
    const int cv = CT_C / 8;                      // 16-byte vectors per pixel
    const float sh = p.H > 1 ? (float)(p.Hs - 1) / (float)(p.H - 1) : 0.0f;
    const float sw = p.W > 1 ? (float)(p.Ws - 1) / (float)(p.W - 1) : 0.0f;
    const uint32_t a_addr = smem_u32(sA);
    // row r of a channel chunk: 128 B, 16-byte chunk j stored at j ^ (r & 7)
    const uint32_t st_off = (v >> 3) * (CT_ATILE / 2);
    const int jv = v & 7;
    int cnt = 0;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      int b, xs, y0, R;
      unit_of(u, b, xs, y0, R);
      const int xa = xs + CT_VALID * (r0 >> 5) + (r0 & 31) - 1;  // image column of the run's first pixel
      const int k0 = xa < 0 ? -xa : 0;                              // pixels k0 .. k1-1 of the run are inside the image
      const int k1 = min(8, p.W - xa);
      // per unit: the first source column under the run and, for each pair of neighbouring columns (wlo + c, wlo + c + 1), the first
      // pixel of the run that lies beyond it (the pixels of a run map to non-decreasing source columns): k0 + #{k : w0(k) - wlo <= c}
      int wlo = 0;
      int kend[CT_COLS - 1];
#pragma unroll
      for (int c = 0; c + 1 < CT_COLS; ++c) kend[c] = k0;
      if (k0 < k1) {
        int w1_;
        float lw_;
        ct_coords(xa + k0, sw, p.Ws, wlo, w1_, lw_);
        for (int k = k0; k < k1; ++k) {
          int w0;
          ct_coords(xa + k, sw, p.Ws, w0, w1_, lw_);
          if (w0 - wlo >= CT_COLS - 1) __trap();  // more source columns under the run than CT_COLS: the host must not choose the fused form
#pragma unroll
          for (int c = 0; c + 1 < CT_COLS; ++c)
            if (w0 - wlo <= c) ++kend[c];
        }
      }
      const uint4* img = p.src + (long long)b * p.Hs * p.Ws * cv + v;
      uint4 raw0[CT_COLS], raw1[CT_COLS];
      auto issue_loads = [&](int yy) {  // the source columns wlo .. wlo + CT_COLS - 1 (clamped) of the two source rows of image row yy
        if (yy < 0 || yy >= p.H || k0 >= k1) return;
        int h0, h1;
        float lh;
        ct_coords(yy, sh, p.Hs, h0, h1, lh);
        const uint4* row0 = img + (long long)h0 * p.Ws * cv;
        const uint4* row1 = img + (long long)h1 * p.Ws * cv;
#pragma unroll
        for (int c = 0; c < CT_COLS; ++c) {
          const int w = min(wlo + c, p.Ws - 1);
          raw0[c] = __ldg(row0 + (long long)w * cv);
          raw1[c] = __ldg(row1 + (long long)w * cv);
        }
      };
      issue_loads(y0 - 1);
      for (int i = 0; i < R + 2; ++i, ++cnt) {
        const int slot = cnt % CT_SLOTS;
        const int yy = y0 - 1 + i;
        const bool row_ok = yy >= 0 && yy < p.H && k0 < k1;
        float2 col[CT_COLS][4];
        if (row_ok) {
          int h0, h1;
          float lh;
          ct_coords(yy, sh, p.Hs, h0, h1, lh);
          const float2 wa = make_float2(1.0f - lh, 1.0f - lh), wb = make_float2(lh, lh);
#pragma unroll
          for (int c = 0; c < CT_COLS; ++c) {
            col[c][0] = ct_ffma2(T16f<FMT>::unpack(raw1[c].x), wb, ct_fmul2(T16f<FMT>::unpack(raw0[c].x), wa));
            col[c][1] = ct_ffma2(T16f<FMT>::unpack(raw1[c].y), wb, ct_fmul2(T16f<FMT>::unpack(raw0[c].y), wa));
            col[c][2] = ct_ffma2(T16f<FMT>::unpack(raw1[c].z), wb, ct_fmul2(T16f<FMT>::unpack(raw0[c].z), wa));
            col[c][3] = ct_ffma2(T16f<FMT>::unpack(raw1[c].w), wb, ct_fmul2(T16f<FMT>::unpack(raw0[c].w), wa));
          }
        }
        if (i + 1 < R + 2) issue_loads(yy + 1);  // in flight while this row tile is blended and stored
        mbar_wait_relaxed(&a_empty[slot], ((cnt / CT_SLOTS) & 1) ^ 1, 200);
        const uint32_t dst = a_addr + slot * CT_ATILE + st_off;
        auto store = [&](int k, const uint4& o) {
          const int r = r0 + k;
          const uint32_t addr = dst + r * 128 + ((jv ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        };
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);  // outside the image: the convolution's zero padding
        if (!row_ok) {
#pragma unroll
          for (int k = 0; k < 8; ++k) store(k, zero);
        } else {
          for (int k = 0; k < k0; ++k) store(k, zero);
          for (int k = k1; k < 8; ++k) store(k, zero);
          int k = k0;
          float xf = (float)(xa + k0);  // (float) of the pixel column, advanced by 1.0f per pixel (exact)
          auto between = [&](float2 l0, float2 l1, float2 l2, float2 l3, float2 q0, float2 q1, float2 q2, float2 q3, int c, int ke) {
            const float w0f = (float)(wlo + c);
            for (; k < ke; ++k, xf += 1.0f) {
              const float lw = __fmul_rn(sw, xf) - w0f;  // = ct_coords: src = scale * (float)dst, l1 = src - (float)i0 (two roundings)
              const float2 ua = make_float2(1.0f - lw, 1.0f - lw), ub = make_float2(lw, lw);  // at the right border w1 == w0 and lw == 0
              uint4 o;
              float2 y;
              y = ct_ffma2(q0, ub, ct_fmul2(l0, ua)); o.x = T16f<FMT>::pack(y.x, y.y);
              y = ct_ffma2(q1, ub, ct_fmul2(l1, ua)); o.y = T16f<FMT>::pack(y.x, y.y);
              y = ct_ffma2(q2, ub, ct_fmul2(l2, ua)); o.z = T16f<FMT>::pack(y.x, y.y);
              y = ct_ffma2(q3, ub, ct_fmul2(l3, ua)); o.w = T16f<FMT>::pack(y.x, y.y);
              store(k, o);
            }
          };
          static_assert(CT_COLS == 6, "one call per pair of neighbouring columns");
          between(col[0][0], col[0][1], col[0][2], col[0][3], col[1][0], col[1][1], col[1][2], col[1][3], 0, kend[0]);
          between(col[1][0], col[1][1], col[1][2], col[1][3], col[2][0], col[2][1], col[2][2], col[2][3], 1, kend[1]);
          between(col[2][0], col[2][1], col[2][2], col[2][3], col[3][0], col[3][1], col[3][2], col[3][3], 2, kend[2]);
          between(col[3][0], col[3][1], col[3][2], col[3][3], col[4][0], col[4][1], col[4][2], col[4][3], 3, kend[3]);
          between(col[4][0], col[4][1], col[4][2], col[4][3], col[5][0], col[5][1], col[5][2], col[5][3], 4, kend[4]);
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
        mbar_arrive(&a_full[slot]);
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, CT_TMEM_COLS);
  }
}

template <int FMT, bool FUSED>
static int launch_conv_tail(const CUtensorMap& tmA, const TailParams& p, cudaStream_t stream) {
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tail_kernel<FMT, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(conv_tail): ") + cudaGetErrorString(e));
    configured = true;
  }
  const int grid = p.num_units < num_sms() ? p.num_units : num_sms();
  conv_tail_kernel<FMT, FUSED><<<grid, CT_THREADS, CT_SMEM, stream>>>(tmA, p);
  count_launch();
  return check_launch("conv_tail_kernel");
}

// rows per unit: the split that minimises (waves of units) x (row tiles built per unit, two of them halo rows)
static void tail_partition(TailParams& p) {
  p.strips = (p.W + CT_STRIP - 1) / CT_STRIP;
  const int sms = num_sms();
  long long best = -1;
  for (int chunks = 1; chunks <= p.H; ++chunks) {
    const int rows = (p.H + chunks - 1) / chunks;
    if (rows < 8 && chunks > 1) break;
    const int real_chunks = (p.H + rows - 1) / rows;
    const long long units = (long long)p.B * p.strips * real_chunks;
    const long long cost = ((units + sms - 1) / sms) * (rows + 2);
    if (best < 0 || cost < best) {
      best = cost;
      p.chunks = real_chunks;
      p.rows_per_chunk = rows;
    }
  }
  p.num_units = p.B * p.strips * p.chunks;
}

static int tail_common(TailParams& p, const void* wpacked, const float* bias, const float* head_w, float head_b, float* out, int B, int H, int W) {
  if (!wpacked || !bias || !head_w || !out) return set_error("vdn_conv_tail: null pointer");
  if (B <= 0 || H <= 0 || W <= 0) return set_error("vdn_conv_tail: bad shape");
  if ((long long)B * ((W + CT_STRIP - 1) / CT_STRIP) * H > 0x7fffffffLL) return set_error("vdn_conv_tail: too many work units");
  p.wpacked = reinterpret_cast<const uint4*>(wpacked);
  p.bias = bias;
  p.head_w = head_w;
  p.head_b = head_b;
  p.out = out;
  p.B = B;
  p.H = H;
  p.W = W;
  tail_partition(p);
  return 0;
}

}  // namespace vdn

using namespace vdn;

extern "C" int vdn_conv_tail(const void* x, const void* wpacked, const float* bias, const float* head_w, float head_b, float* out, int32_t B,
                             int32_t H, int32_t W, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!x) return set_error("vdn_conv_tail: null pointer");
  TailParams p{};
  if (tail_common(p, wpacked, bias, head_w, head_b, out, B, H, W)) return 1;
  const int fmt = get_operand_format();
  CUtensorMap tmA;
  const uint64_t dims[4] = {(uint64_t)CT_C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)CT_C * 2, (uint64_t)W * CT_C * 2, (uint64_t)H * W * CT_C * 2};
  const uint32_t box[4] = {64, 32, 1, 1};
  if (make_tensor_map(&tmA, x, fmt, 4, dims, strides, box)) return 1;
  return fmt ? launch_conv_tail<1, false>(tmA, p, stream) : launch_conv_tail<0, false>(tmA, p, stream);
}

extern "C" int vdn_conv_tail_up(const void* src, int32_t Hs, int32_t Ws, const void* wpacked, const float* bias, const float* head_w, float head_b,
                                float* out, int32_t B, int32_t H, int32_t W, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!src) return set_error("vdn_conv_tail_up: null pointer");
  if (Hs <= 0 || Ws <= 0) return set_error("vdn_conv_tail_up: bad source shape");
  // a run of 8 output pixels must lie between at most CT_COLS source columns: 7 * (Ws - 1) / (W - 1) < 4 (the model resizes 8 p -> 14 p)
  if (W > 1 && 7LL * (Ws - 1) >= 4LL * (W - 1)) return set_error("vdn_conv_tail_up: horizontal scale above 4/7 (use vdn_bilinear_nhwc + vdn_conv_tail)");
  TailParams p{};
  if (tail_common(p, wpacked, bias, head_w, head_b, out, B, H, W)) return 1;
  p.src = reinterpret_cast<const uint4*>(src);
  p.Hs = Hs;
  p.Ws = Ws;
  CUtensorMap tmA{};  // unused by the fused form
  return get_operand_format() ? launch_conv_tail<1, true>(tmA, p, stream) : launch_conv_tail<0, true>(tmA, p, stream);
}

// Bandwidth-bound kernels: LayerNorm, GroupNorm (+ frame->pixel-major transpose), patch im2col, stride-2 im2col,
// bilinear resize (align_corners=True), ReLU / casts, window alignment reductions, Sobel normals.
// All are coalesced, 16-byte vectorised where the layout allows, warp-shuffle reductions; shared memory only for block reductions,
// the GroupNorm cluster merge (distributed shared memory) and the patch im2col's transposing copy of its output rows.
#include <stdlib.h>

#include "../../include/vdn_b200.h"
#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

static inline unsigned grid_for(long long work_items, int per_block, int waves = 16) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 1536: DINOv2 ViT-g), fp32 statistics, 16-bit output
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 12;  // float4 per lane -> C <= 32 * 12 * 4 = 1536

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, void* __restrict__ out, long long rows,
                 int C, float eps, int drop_first, int rows_per_batch, const float* __restrict__ pe, int pe_len, int fmt) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nvec = C >> 2;
  for (long long row = warp0; row < rows; row += nwarps) {
    long long orow = row;
    if (drop_first) {
      const long long bb = row / rows_per_batch;
      const int j = int(row - bb * rows_per_batch);
      if (j == 0) continue;
      orow = bb * (rows_per_batch - 1) + j - 1;
    }
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    float4 v[LN_MAXV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        v[i] = xr[idx];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        const float a = v[i].x - mean, bq = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + bq * bq) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    const float4* pr = pe ? reinterpret_cast<const float4*>(pe + (row % pe_len) * C) : nullptr;
    uint2* orow_p = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out) + orow * C);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(w) + idx);
        const float4 be = __ldg(reinterpret_cast<const float4*>(b) + idx);
        float4 y;
        y.x = (v[i].x - mean) * rstd * g.x + be.x;
        y.y = (v[i].y - mean) * rstd * g.y + be.y;
        y.z = (v[i].z - mean) * rstd * g.z + be.z;
        y.w = (v[i].w - mean) * rstd * g.w + be.w;
        if (pr) {
          const float4 p4 = __ldg(pr + idx);
          y.x += p4.x; y.y += p4.y; y.z += p4.z; y.w += p4.w;
        }
        uint2 u;
        u.x = pack16(y.x, y.y, fmt);
        u.y = pack16(y.z, y.w, fmt);
        orow_p[idx] = u;
      }
    }
  }
}

// C = 128 * NV: two rows per warp in flight (all 2 x NV 16-byte loads are issued before the first reduction), 32-bit index
// arithmetic, streaming loads — the ViT blocks' 48 LayerNorms per window.  PE adds the positional table row (row % pe_len) to the
// output (motion-module attention inputs), DROP iterates over OUTPUT rows and skips the first row of every batch of rpb source rows
// (the tapped features without their cls token): both used to take the generic kernel at less than half this kernel's rate.
template <int NV, int FMT, bool PE, bool DROP>
__global__ void __launch_bounds__(256)
layernorm_rows2_kernel(const float4* __restrict__ x, const float4* __restrict__ w, const float4* __restrict__ b, uint2* __restrict__ out, int rows,
                       float eps, const float4* __restrict__ pe, int pe_len, int rpb) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  constexpr float invC = 1.0f / (float)(NV * 128);
  for (int row = warp0; row < rows; row += 2 * nwarps) {   // `rows` / `row` count output rows
    const int row2 = row + nwarps;
    const bool has2 = row2 < rows;
    const int rb = has2 ? row2 : row;
    const int sa_row = DROP ? row + row / (rpb - 1) + 1 : row;   // source rows
    const int sb_row = DROP ? rb + rb / (rpb - 1) + 1 : rb;
    const float4* xa = x + (long long)sa_row * (NV * 32);
    const float4* xb = x + (long long)sb_row * (NV * 32);
    float4 va[NV], vb[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) va[i] = __ldcs(xa + lane + 32 * i);
#pragma unroll
    for (int i = 0; i < NV; ++i) vb[i] = __ldcs(xb + lane + 32 * i);
    float sa = 0.0f, sb = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sa += (va[i].x + va[i].y) + (va[i].z + va[i].w);
      sb += (vb[i].x + vb[i].y) + (vb[i].z + vb[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    const float ma = sa * invC, mb = sb * invC;
    float qa = 0.0f, qb = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a0 = va[i].x - ma, a1 = va[i].y - ma, a2 = va[i].z - ma, a3 = va[i].w - ma;
      const float b0 = vb[i].x - mb, b1 = vb[i].y - mb, b2 = vb[i].z - mb, b3 = vb[i].w - mb;
      qa += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      qb += (b0 * b0 + b1 * b1) + (b2 * b2 + b3 * b3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      qa += __shfl_xor_sync(0xffffffffu, qa, o);
      qb += __shfl_xor_sync(0xffffffffu, qb, o);
    }
    const float ra = rsqrtf(qa * invC + eps), rb_ = rsqrtf(qb * invC + eps);
    uint2* oa = out + (long long)row * (NV * 32);
    uint2* ob = out + (long long)row2 * (NV * 32);
    const float4* pa = PE ? pe + (long long)(sa_row % pe_len) * (NV * 32) : nullptr;
    const float4* pb = PE ? pe + (long long)(sb_row % pe_len) * (NV * 32) : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = __ldg(w + lane + 32 * i);
      const float4 be = __ldg(b + lane + 32 * i);
      float4 ya, yb;
      ya.x = (va[i].x - ma) * ra * g.x + be.x; ya.y = (va[i].y - ma) * ra * g.y + be.y;
      ya.z = (va[i].z - ma) * ra * g.z + be.z; ya.w = (va[i].w - ma) * ra * g.w + be.w;
      yb.x = (vb[i].x - mb) * rb_ * g.x + be.x; yb.y = (vb[i].y - mb) * rb_ * g.y + be.y;
      yb.z = (vb[i].z - mb) * rb_ * g.z + be.z; yb.w = (vb[i].w - mb) * rb_ * g.w + be.w;
      if (PE) {
        const float4 p0 = __ldg(pa + lane + 32 * i), p1 = __ldg(pb + lane + 32 * i);
        ya.x += p0.x; ya.y += p0.y; ya.z += p0.z; ya.w += p0.w;
        yb.x += p1.x; yb.y += p1.y; yb.z += p1.z; yb.w += p1.w;
      }
      uint2 u;
      u.x = T16f<FMT>::pack(ya.x, ya.y);
      u.y = T16f<FMT>::pack(ya.z, ya.w);
      oa[lane + 32 * i] = u;
      if (has2) {
        u.x = T16f<FMT>::pack(yb.x, yb.y);
        u.y = T16f<FMT>::pack(yb.z, yb.w);
        ob[lane + 32 * i] = u;
      }
    }
  }
}

// rows = output rows; pe / drop select the variant
template <int NV>
static void launch_layernorm_rows2(const float* x, const float* w, const float* b, void* out, long long rows, float eps, const float* pe, int pe_len,
                                   int drop_rpb, cudaStream_t stream) {
  const unsigned grid = grid_for((rows + 1) / 2, 8);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  const float4* p4 = reinterpret_cast<const float4*>(pe);
  uint2* o2 = reinterpret_cast<uint2*>(out);
#define VDN_LN2(F, P, D) layernorm_rows2_kernel<NV, F, P, D><<<grid, 256, 0, stream>>>(x4, w4, b4, o2, (int)rows, eps, p4, pe_len, drop_rpb)
  const int fmt = get_operand_format();
  if (pe != nullptr) { if (fmt) VDN_LN2(1, true, false); else VDN_LN2(0, true, false); }
  else if (drop_rpb > 0) { if (fmt) VDN_LN2(1, false, true); else VDN_LN2(0, false, true); }
  else { if (fmt) VDN_LN2(1, false, false); else VDN_LN2(0, false, false); }
#undef VDN_LN2
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: one block per (frame, group); two passes over an L2-resident slice
// ------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const void* __restrict__ x, float* __restrict__ stats, int D, int C, int groups, float eps, int fmt) {
  const int f = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cg = C / groups;
  const uint16_t* base = reinterpret_cast<const uint16_t*>(x) + (long long)f * D * C + g * cg;
  __shared__ double red[8];
  __shared__ float s_mean;
  const int n = D * cg;
  const int vpr = cg >> 3;  // 16-byte vectors per pixel row of this group (VEC path: cg % 8 == 0)
  const int nvec = D * vpr;
  // pass 1: mean
  float s = 0.0f;
  if (VEC) {
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const int d = i / vpr, v = i - d * vpr;
      const uint4 u = *reinterpret_cast<const uint4*>(base + (long long)d * C + v * 8);
      const float2 a = unpack16(u.x, fmt), b2 = unpack16(u.y, fmt), c2 = unpack16(u.z, fmt), d2 = unpack16(u.w, fmt);
      s += ((a.x + a.y) + (b2.x + b2.y)) + ((c2.x + c2.y) + (d2.x + d2.y));
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int d = i / cg, c = i - d * cg;
      s += load16(base, (long long)d * C + c, fmt);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (double)s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    s_mean = (float)(t / (double)n);
  }
  __syncthreads();
  const float mean = s_mean;
  // pass 2 (the slice is L2-resident): centred sum of squares
  float q = 0.0f;
  if (VEC) {
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const int d = i / vpr, v = i - d * vpr;
      const uint4 u = *reinterpret_cast<const uint4*>(base + (long long)d * C + v * 8);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = unpack16(w4[k], fmt);
        q = fmaf(a.x - mean, a.x - mean, q);
        q = fmaf(a.y - mean, a.y - mean, q);
      }
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int d = i / cg, c = i - d * cg;
      const float v = load16(base, (long long)d * C + c, fmt) - mean;
      q += v * v;
    }
  }
  q = warp_sum(q);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (double)q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    stats[2 * blockIdx.x] = mean;
    stats[2 * blockIdx.x + 1] = (float)(1.0 / sqrt(t / (double)n + (double)eps));
  }
}

// apply + transpose: one warp per (frame, pixel) row; out row = (b*D + d)*T + f
template <bool VEC>
__global__ void __launch_bounds__(256)
groupnorm_apply_tc_kernel(const void* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ w, const float* __restrict__ b,
                          void* __restrict__ out, int Bv, int T, int D, int C, int groups, int fmt) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long rows = (long long)Bv * T * D;
  const int cg = C / groups;
  const int nvec = C >> 3;  // 8 x 16-bit per 16-byte vector
  for (long long row = warp0; row < rows; row += nwarps) {
    const long long frame = row / D;  // b*T + f
    const int d = int(row - frame * D);
    const long long bb = frame / T;
    const int f = int(frame - bb * T);
    const long long orow = (bb * D + d) * T + f;
    const uint4* xr = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(x) + row * C);
    uint4* orp = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + orow * C);
    for (int idx = lane; idx < nvec; idx += 32) {
      const uint4 u = xr[idx];
      const int c0 = idx * 8;
      float v[8];
      float2 t;
      t = unpack16(u.x, fmt); v[0] = t.x; v[1] = t.y;
      t = unpack16(u.y, fmt); v[2] = t.x; v[3] = t.y;
      t = unpack16(u.z, fmt); v[4] = t.x; v[5] = t.y;
      t = unpack16(u.w, fmt); v[6] = t.x; v[7] = t.y;
      if (VEC) {  // cg % 8 == 0: the 8 channels of a vector share one group
        const float2 ms = __ldg(reinterpret_cast<const float2*>(stats) + frame * groups + c0 / cg);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(w + c0)), g1 = __ldg(reinterpret_cast<const float4*>(w + c0) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(b + c0) + 1);
        const float gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float gb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (v[i] - ms.x) * ms.y * gw[i] + gb[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = c0 + i;
          const int g = c / cg;
          const float mean = __ldg(stats + 2 * (frame * groups + g));
          const float rstd = __ldg(stats + 2 * (frame * groups + g) + 1);
          v[i] = (v[i] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
        }
      }
      uint4 o;
      o.x = pack16(v[0], v[1], fmt);
      o.y = pack16(v[2], v[3], fmt);
      o.z = pack16(v[4], v[5], fmt);
      o.w = pack16(v[6], v[7], fmt);
      orp[idx] = o;
    }
  }
}

// Statistics + apply + transpose in ONE launch (motion_module.py:103-115: GroupNorm(32, eps 1e-6) then the (b f) d c -> (b d) f c
// rearrange).  A cluster of GN_CLUSTER CTAs owns one frame, each CTA a contiguous slice of its pixel rows; a thread owns one
// 16-byte channel vector (8 channels of one group) and a row lane, so every warp access is one contiguous 512-byte run.
//   pass 1  statistics: each thread sums (x - K) and (x - K)^2 around its own pivot K and turns them into (n, mean, M2); the
//           threads of a group are Chan-merged through shared memory, the CTAs of the frame through distributed shared memory, in
//           a fixed order (every CTA does the same merge: bit-identical statistics in all of them, run to run).
//   pass 2  the CTA re-reads its own slice back to front (the rows read last are the ones most likely still in L2) and writes the
//           normalised rows pixel-major.
// Measured and dropped: 512 threads with the first 200 KB of the slice kept in shared memory (one CTA per SM, 3.5 waves of
// clusters: every wave idles through the two cluster barriers) was slower than the two-kernel form (0.33 against 0.27 ms per step).
// Against the two-kernel form (one block per (frame, group) reading 64-byte pieces at a 2 KB stride, twice; then one warp per row)
// the tensor comes from HBM once instead of twice and no access is narrower than a full warp line.
constexpr int GN_CLUSTER = 8;
constexpr int GN_MAXG = 64;
constexpr int GN_THREADS = 256;

__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  if (nb == 0.0f) return;
  const float nn = n + nb;
  const float delta = mb - mean;
  const float fb = nb / nn;
  mean = fmaf(delta, fb, mean);
  m2 = m2 + m2b + delta * delta * n * fb;
  n = nn;
}

__global__ void __cluster_dims__(GN_CLUSTER, 1, 1) __launch_bounds__(GN_THREADS)
groupnorm_fused_tc_kernel(const uint4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, uint4* __restrict__ out,
                          float* __restrict__ stats_out, int T, int D, int C, int groups, float eps, int fmt, int reverse) {
  __shared__ float s_n[GN_THREADS], s_mean[GN_THREADS], s_m2[GN_THREADS];
  __shared__ float part[3][GN_MAXG];   // this CTA's (n, mean, M2) per group: read by the whole cluster
  __shared__ float fin[2][GN_MAXG];    // (mean, rstd) per group
  const unsigned rank = blockIdx.x % GN_CLUSTER;
  const int frame = blockIdx.x / GN_CLUSTER;
  const int nv = C >> 3;                 // vectors per pixel row
  const int gv = (C / groups) >> 3;      // vectors per group per pixel row
  const int lanes = GN_THREADS / nv;     // row lanes of the block (host guarantees GN_THREADS % nv == 0)
  const int v = threadIdx.x % nv, rl = threadIdx.x / nv;
  const int rows_per = (D + GN_CLUSTER - 1) / GN_CLUSTER;
  const int r0 = min(D, (int)rank * rows_per), r1 = min(D, r0 + rows_per);
  const uint4* xf = x + (long long)frame * D * nv + v;

  // ---- pass 1: sums of (x - K) and (x - K)^2 around a per-thread pivot K (the mean of the thread's first vector, within sigma / sqrt(8)
  // of the mean: M2 = s2 - s1^2 / n loses nothing), turned into (n, mean, M2) once per thread.  (A Chan merge per vector cost a
  // division and ~50 instructions per 16 bytes: the kernel was issue-bound at 51 % issue-active, ncu.)
  float n = 0.0f, mean = 0.0f, m2 = 0.0f;
  constexpr int U = 8;
  {
    float K = 0.0f, s1 = 0.0f, s2 = 0.0f;
    for (int r = r0 + rl; r < r1; r += lanes * U) {
      const uint4* px = xf + (long long)r * nv;
      uint4 u[U];
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (r + k * lanes < r1) u[k] = px[k * GN_THREADS];  // lanes * nv = GN_THREADS vectors between a thread's consecutive rows
      if (r == r0 + rl) {
        const float2 a = unpack16(u[0].x, fmt), c2 = unpack16(u[0].y, fmt), d2 = unpack16(u[0].z, fmt), e2 = unpack16(u[0].w, fmt);
        K = (((a.x + a.y) + (c2.x + c2.y)) + ((d2.x + d2.y) + (e2.x + e2.y))) * 0.125f;
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (r + k * lanes < r1) {
          const float2 a = unpack16(u[k].x, fmt), c2 = unpack16(u[k].y, fmt), d2 = unpack16(u[k].z, fmt), e2 = unpack16(u[k].w, fmt);
          const float t0 = a.x - K, t1 = a.y - K, t2 = c2.x - K, t3 = c2.y - K, t4 = d2.x - K, t5 = d2.y - K, t6 = e2.x - K, t7 = e2.y - K;
          s1 += ((t0 + t1) + (t2 + t3)) + ((t4 + t5) + (t6 + t7));
          float qa = t0 * t0, qb = t1 * t1;
          qa = fmaf(t2, t2, qa); qb = fmaf(t3, t3, qb);
          qa = fmaf(t4, t4, qa); qb = fmaf(t5, t5, qb);
          qa = fmaf(t6, t6, qa); qb = fmaf(t7, t7, qb);
          s2 += qa + qb;
          n += 8.0f;
        }
      }
    }
    if (n > 0.0f) {
      const float d = s1 / n;
      mean = K + d;
      m2 = fmaxf(s2 - s1 * d, 0.0f);
    }
  }
  s_n[threadIdx.x] = n; s_mean[threadIdx.x] = mean; s_m2[threadIdx.x] = m2;
  __syncthreads();
  if ((int)threadIdx.x < groups) {  // the GN_THREADS / groups threads of group g: row lane l, vector g*gv + j
    const int g = threadIdx.x;
    float gn = 0.0f, gm = 0.0f, gq = 0.0f;
    for (int l = 0; l < lanes; ++l)
      for (int j = 0; j < gv; ++j) {
        const int t = l * nv + g * gv + j;
        chan_merge(gn, gm, gq, s_n[t], s_mean[t], s_m2[t]);
      }
    part[0][g] = gn; part[1][g] = gm; part[2][g] = gq;
  }
  cluster_sync_all();  // release / acquire: every CTA's part[] is visible to the others
  if ((int)threadIdx.x < groups) {
    const int g = threadIdx.x;
    float gn = 0.0f, gm = 0.0f, gq = 0.0f;
    for (unsigned cr = 0; cr < (unsigned)GN_CLUSTER; ++cr) {
      float pn, pm, pq;
      const uint32_t ra = mapa_rank(smem_u32(&part[0][g]), cr), rb = mapa_rank(smem_u32(&part[1][g]), cr), rcc = mapa_rank(smem_u32(&part[2][g]), cr);
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pn) : "r"(ra) : "memory");
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pm) : "r"(rb) : "memory");
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pq) : "r"(rcc) : "memory");
      chan_merge(gn, gm, gq, pn, pm, pq);
    }
    const float rstd = rsqrtf(gq / gn + eps);
    fin[0][g] = gm; fin[1][g] = rstd;
    if (rank == 0 && stats_out != nullptr) {
      stats_out[2 * (frame * groups + g)] = gm;
      stats_out[2 * (frame * groups + g) + 1] = rstd;
    }
  }
  cluster_sync_all();  // no CTA leaves (or reuses part[]) while a peer still reads it; also orders fin[] inside the CTA

  // ---- pass 2: y = x * a + b with a = rstd * w, b = bias - mean * a (one FFMA per value)
  const int g = v / gv;
  const float mu = fin[0][g], rstd = fin[1][g];
  float sc[8], sh[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v), g1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + 2 * v), b1 = __ldg(reinterpret_cast<const float4*>(b) + 2 * v + 1);
    sc[0] = g0.x; sc[1] = g0.y; sc[2] = g0.z; sc[3] = g0.w; sc[4] = g1.x; sc[5] = g1.y; sc[6] = g1.z; sc[7] = g1.w;
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sc[i] *= rstd;
      sh[i] = fmaf(-mu, sc[i], sh[i]);
    }
  }
  const int bb = frame / T, f = frame - bb * T;
  uint4* of = out + ((long long)bb * D * T + f) * nv + v;   // row (bb*D + d)*T + f  ->  + d * T * nv
  const int ostep = T * GN_THREADS;                         // output vectors between a thread's consecutive rows
  const int span = r1 - r0 - rl;  // rows r0 + rl, r0 + rl + lanes, ... < r1 belong to this thread
  const int nch = span > 0 ? (span - 1) / (lanes * U) + 1 : 0;
  for (int it = 0; it < nch; ++it) {
    const int ch = reverse ? nch - 1 - it : it;
    const int r = r0 + rl + ch * lanes * U;
    const uint4* px = xf + (long long)r * nv;
    uint4* po = of + (long long)r * T * nv;
    uint4 u[U];
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (r + k * lanes < r1) u[k] = px[k * GN_THREADS];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (r + k * lanes < r1) {
        const uint32_t wds[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = unpack16(wds[i], fmt);
          o[i] = pack16(fmaf(a.x, sc[2 * i], sh[2 * i]), fmaf(a.y, sc[2 * i + 1], sh[2 * i + 1]), fmt);
        }
        po[(long long)k * ostep] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layout kernels
// ------------------------------------------------------------------------------------------------
// One thread per (patch, channel c, patch row i): 14 contiguous fp32 pixels -> 14 contiguous 16-bit columns at c*196 + i*14
// (7 x 8-byte loads, 7 x 4-byte stores; the 42 threads of a patch write one contiguous 1176-byte row, the thread of (c=2, i=13) also
// zeroes the padding columns).  The element-per-thread version decoded (c, i, j) with three integer divisions per 2-byte store
// and ran at 0.15 of HBM bandwidth.
__global__ void __launch_bounds__(256)
patch_im2col_kernel(const float* __restrict__ img, void* __restrict__ out, int B, int H, int W, int Kp, int fmt) {
  const int ph = H / 14, pw = W / 14;
  const long long total = (long long)B * ph * pw * 42;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / 42;
    const int ci = int(idx - row * 42);
    const int c = ci / 14, i = ci - c * 14;
    const long long bimg = row / (ph * pw);
    const int p = int(row - bimg * ph * pw);
    const int py = p / pw, px = p - py * pw;
    const float2* src = reinterpret_cast<const float2*>(img + ((bimg * 3 + c) * H + py * 14 + i) * (long long)W + px * 14);  // 56-byte multiples: 8-byte aligned
    uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(out) + row * Kp + ci * 14);                      // 28-byte multiples: 4-byte aligned
    float2 v[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) v[j] = __ldg(src + j);
#pragma unroll
    for (int j = 0; j < 7; ++j) dst[j] = pack16(v[j].x, v[j].y, fmt);
    if (ci == 41) {
      uint16_t* pad = reinterpret_cast<uint16_t*>(out) + row * Kp + 588;
      for (int k = 0; k < Kp - 588; ++k) pad[k] = 0;
    }
  }
}

// Row form of the same layout change: a block owns the (up to 40) patches of one patch row of one image.  Work item = one
// (image-row piece (c, i), patch) pair = 14 contiguous floats; a warp takes a 4 x 8 tile of them (4 image rows x 8 neighbouring
// patches: four contiguous 448-byte runs), rounds them into a shared-memory copy of the output rows (row pitch Kp/2 + 4 words, so
// that the 32 lanes of a store hit 32 banks) and the block writes that copy out as contiguous 16-byte vectors.  The
// thread-per-(patch, c, i) kernel above reads 56-byte pieces at a 2 KB stride (0.32 of the copy peak).
__global__ void __launch_bounds__(256)
patch_im2col_rows_kernel(const float* __restrict__ img, uint4* __restrict__ out, int H, int W, int Kp, int pitch, int seg, int nseg, int fmt) {
  extern __shared__ uint4 pim_smem[];
  uint32_t* sm = reinterpret_cast<uint32_t*>(pim_smem);
  const int ph = H / 14, pw = W / 14;
  int bid = blockIdx.x;
  const int sg = bid % nseg;
  bid /= nseg;
  const int py = bid % ph;
  const int bimg = bid / ph;
  const int p0 = sg * seg;
  const int np = min(seg, pw - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int padw = (Kp - 588) >> 1;  // padding columns as 32-bit words
  for (int i = threadIdx.x; i < np * padw; i += 256) {
    const int p = i / padw, k = i - p * padw;
    sm[p * pitch + 294 + k] = 0u;
  }
  const float* base = img + ((long long)bimg * 3 * H + py * 14) * W + p0 * 14;
  const int lci = lane >> 3, lpp = lane & 7;
  const int tiles_pp = (np + 7) >> 3;
  const int ntiles = 11 * tiles_pp;  // 42 image-row pieces in groups of 4
  for (int t = warp; t < ntiles; t += 8) {
    const int tci = t / tiles_pp, tpp = t - tci * tiles_pp;
    const int ci = tci * 4 + lci, pp = tpp * 8 + lpp;
    if (ci < 42 && pp < np) {
      const int c = ci / 14, i = ci - c * 14;
      const float2* src = reinterpret_cast<const float2*>(base + ((long long)c * H + i) * W + pp * 14);
      float2 v[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) v[j] = __ldg(src + j);
      uint32_t* d = sm + pp * pitch + ci * 7;
#pragma unroll
      for (int j = 0; j < 7; ++j) d[j] = pack16(v[j].x, v[j].y, fmt);
    }
  }
  __syncthreads();
  const int kv = Kp >> 3, pv = pitch >> 2;
  uint4* dst = out + ((long long)(bimg * ph + py) * pw + p0) * kv;
  for (int r = warp; r < np; r += 8)
    for (int q = lane; q < kv; q += 32) dst[r * kv + q] = pim_smem[r * pv + q];
}

__global__ void write_cls_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int B, int tokens, int C) {
  const int total = B * C;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int bb = idx / C, c = idx - bb * C;
    x[(long long)bb * tokens * C + c] = cls[c] + pos[c];
  }
}

// use_clstoken readout input (dpt.py:129-132 / dpt_temporal.py:56-59): out[f*P + p] = [ tok[f*tok_pitch + p] | cls[f*cls_pitch] ], 16-byte vectors
__global__ void __launch_bounds__(256)
readout_concat_kernel(const uint4* __restrict__ tok, long long tok_pitch, const uint4* __restrict__ cls, long long cls_pitch, uint4* __restrict__ out,
                      long long frames, int P, int cv) {
  const long long total = frames * P * 2 * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = int(idx % (2 * cv));
    const long long row = idx / (2 * cv);
    const long long f = row / P;
    const int p = int(row - f * P);
    out[idx] = c < cv ? __ldg(tok + (f * tok_pitch + p) * cv + c) : __ldg(cls + f * cls_pitch * cv + (c - cv));
  }
}

// NHWC [B,H,W,C] -> [B*Ho*Wo, 9*C] for the 3x3 stride-2 pad-1 conv; column = (r*3+s)*C + c; 8 channels per thread
__global__ void __launch_bounds__(256)
im2col_3x3_s2_kernel(const void* __restrict__ x, void* __restrict__ out, int B, int H, int W, int C) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int cv = C >> 3;
  const long long total = (long long)B * Ho * Wo * 9 * cv;
  const uint4* xin = reinterpret_cast<const uint4*>(x);
  uint4* o = reinterpret_cast<uint4*>(out);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = int(idx % cv);
    long long t = idx / cv;
    const int tap = int(t % 9);
    t /= 9;
    const int wo = int(t % Wo);
    t /= Wo;
    const int ho = int(t % Ho);
    const long long bimg = t / Ho;
    const int hi = ho * 2 - 1 + tap / 3, wi = wo * 2 - 1 + tap % 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = xin[((bimg * H + hi) * W + wi) * cv + c8];
    o[idx] = v;
  }
}

// Pixel form: a block owns one (or, for narrow C, several) output pixels, a thread one 16-byte channel vector of all nine taps: the
// (b, ho, wo) decode happens once per thread in 32-bit arithmetic and the nine loads are independent (the element-per-thread kernel
// above does four 64-bit divisions per 16 bytes moved).
template <int NT>
__global__ void __launch_bounds__(NT)
im2col_3x3_s2_pix_kernel(const uint4* __restrict__ xin, uint4* __restrict__ o, int H, int W, int Ho, int Wo, int cv, int total_pix) {
  const int ppb = cv >= NT ? 1 : NT / cv;
  const int pl = cv >= NT ? 0 : (int)threadIdx.x / cv;
  const int c0 = cv >= NT ? (int)threadIdx.x : (int)threadIdx.x - pl * cv;
  const int pix = blockIdx.x * ppb + pl;
  if (pl >= ppb || pix >= total_pix) return;
  const int wo = pix % Wo;
  const int t = pix / Wo;
  const int ho = t % Ho;
  const int bimg = t / Ho;
  uint4* orow = o + (long long)pix * 9 * cv;
  for (int c8 = c0; c8 < cv; c8 += NT) {
    uint4 v[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hi = ho * 2 - 1 + tap / 3, wi = wo * 2 - 1 + tap % 3;
      v[tap] = (hi >= 0 && hi < H && wi >= 0 && wi < W) ? xin[((long long)(bimg * H + hi) * W + wi) * cv + c8] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) orow[tap * cv + c8] = v[tap];
  }
}

// PyTorch's align_corners=True source index: scale = (in-1)/(out-1) (0 if out==1), src = scale*dst
__device__ __forceinline__ void ac_coords(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  const float src = scale * (float)dst;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2): one issue slot for two lanes of a lerp.
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  float2 c;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(rc));
  return c;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// A thread owns one 16-byte channel vector over a run of L consecutive output pixels of one output row and keeps the two
// vertically-lerped source columns it is between in registers (fp32, packed FFMA2); when the run crosses a source pixel the right
// column becomes the left one and one new column is loaded (two 16-byte loads, L1/L2 hits).  No shared memory, no barrier;
// consecutive threads are consecutive channel vectors, so every access of a warp is one contiguous run of up to 512 bytes.
// Same coordinate arithmetic as ATen (ac_coords); the evaluation order (vertical first) differs from ATen's by fp32 rounding only.
// History: a gather kernel (four 16-byte loads and a full 4-tap blend per output vector) was issue-bound at 2.8 TB/s; a row-staged
// variant through shared memory reached 4.1-4.9 TB/s alone but stalled on its own shared-memory traffic; this one 4.2-5.7 TB/s.
// A straight-line form for ~2x up-sampling (8 pixels x 4 rows per thread, the pixel-to-column-pair assignment worked out once per
// thread as in the fused head tail's producers: 62 % -> 51 % issue-active, bit-identical outputs) took the same 384-388 us on the
// 148^2 -> 296^2 resize (profiles/r02_bandwidth_pass.txt): the kernel is not issue-bound, 80 % of its traffic is writes.  Removed.
template <int FMT, int RELU2>
__global__ void __launch_bounds__(128)
bilinear_slide_kernel(const uint4* __restrict__ xin, uint4* __restrict__ o, uint4* __restrict__ o_relu, int H, int W, int Ho, int Wo, int cv, int L,
                      int nseg, int relu_out) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= nseg * cv) return;
  const int seg = t / cv, v = t - seg * cv;
  const int ho = blockIdx.y;
  const long long bimg = blockIdx.z;
  const float sh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.0f;
  const float sw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.0f;
  int h0, h1;
  float lh;
  ac_coords(ho, sh, H, h0, h1, lh);
  const uint4* row0 = xin + (bimg * H + h0) * (long long)W * cv + v;
  const uint4* row1 = xin + (bimg * H + h1) * (long long)W * cv + v;
  const float2 wa = make_float2(1.0f - lh, 1.0f - lh), wb = make_float2(lh, lh);
  auto loadcol = [&](int w, float2 (&c)[4]) {
    const uint4 a = __ldg(row0 + (long long)w * cv);
    const uint4 b = __ldg(row1 + (long long)w * cv);
    c[0] = ffma2(T16f<FMT>::unpack(b.x), wb, fmul2(T16f<FMT>::unpack(a.x), wa));
    c[1] = ffma2(T16f<FMT>::unpack(b.y), wb, fmul2(T16f<FMT>::unpack(a.y), wa));
    c[2] = ffma2(T16f<FMT>::unpack(b.z), wb, fmul2(T16f<FMT>::unpack(a.z), wa));
    c[3] = ffma2(T16f<FMT>::unpack(b.w), wb, fmul2(T16f<FMT>::unpack(a.w), wa));
  };
  const int wo_a = seg * L;
  const int wo_b = min(wo_a + L, Wo);
  uint4* orow = o + (bimg * Ho + ho) * (long long)Wo * cv + v;
  uint4* orow_relu = RELU2 ? o_relu + (bimg * Ho + ho) * (long long)Wo * cv + v : nullptr;
  float2 A[4], B[4];
  int cur0 = -1, cur1 = -1;
  for (int wo = wo_a; wo < wo_b; ++wo) {
    int w0, w1;
    float lw;
    ac_coords(wo, sw, W, w0, w1, lw);
    if (w0 != cur0) {
      if (w0 == cur1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) A[k] = B[k];
      } else {
        loadcol(w0, A);
      }
      cur0 = w0;
    }
    if (w1 != cur1) {
      if (w1 == cur0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) B[k] = A[k];
      } else {
        loadcol(w1, B);
      }
      cur1 = w1;
    }
    const float2 ua = make_float2(1.0f - lw, 1.0f - lw), ub = make_float2(lw, lw);
    float2 y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      y[k] = ffma2(B[k], ub, fmul2(A[k], ua));
      if (relu_out) { y[k].x = fmaxf(y[k].x, 0.0f); y[k].y = fmaxf(y[k].y, 0.0f); }
    }
    uint4 r;
    r.x = T16f<FMT>::pack(y[0].x, y[0].y); r.y = T16f<FMT>::pack(y[1].x, y[1].y);
    r.z = T16f<FMT>::pack(y[2].x, y[2].y); r.w = T16f<FMT>::pack(y[3].x, y[3].y);
    orow[(long long)wo * cv] = r;
    if (RELU2) {
      uint4 rr;
      rr.x = T16f<FMT>::pack(fmaxf(y[0].x, 0.0f), fmaxf(y[0].y, 0.0f)); rr.y = T16f<FMT>::pack(fmaxf(y[1].x, 0.0f), fmaxf(y[1].y, 0.0f));
      rr.z = T16f<FMT>::pack(fmaxf(y[2].x, 0.0f), fmaxf(y[2].y, 0.0f)); rr.w = T16f<FMT>::pack(fmaxf(y[3].x, 0.0f), fmaxf(y[3].y, 0.0f));
      orow_relu[(long long)wo * cv] = rr;
    }
  }
}

__global__ void __launch_bounds__(256)
bilinear_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W, int Ho, int Wo, int relu) {
  const long long total = (long long)N * Ho * Wo;
  const float sh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.0f;
  const float sw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.0f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int wo = int(idx % Wo);
    long long t = idx / Wo;
    const int ho = int(t % Ho);
    const long long n = t / Ho;
    int h0, h1, w0, w1;
    float lh, lw;
    ac_coords(ho, sh, H, h0, h1, lh);
    ac_coords(wo, sw, W, w0, w1, lw);
    const float* p = x + n * (long long)H * W;
    // same association as ATen's upsample_bilinear2d: h0lambda*(w0lambda*a + w1lambda*b) + h1lambda*(w0lambda*c + w1lambda*d)
    float y = (1.0f - lh) * ((1.0f - lw) * p[h0 * (long long)W + w0] + lw * p[h0 * (long long)W + w1]) +
              lh * ((1.0f - lw) * p[h1 * (long long)W + w0] + lw * p[h1 * (long long)W + w1]);
    if (relu) y = fmaxf(y, 0.0f);
    out[idx] = y;
  }
}

__global__ void __launch_bounds__(256) relu16_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, long long nvec, int fmt) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nvec; idx += (long long)gridDim.x * blockDim.x) {
    const uint4 u = x[idx];
    const uint32_t* pu = &u.x;
    uint4 r;
    uint32_t* pr = &r.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack16(pu[i], fmt);
      pr[i] = pack16(fmaxf(f.x, 0.0f), fmaxf(f.y, 0.0f), fmt);
    }
    out[idx] = r;
  }
}

__global__ void __launch_bounds__(256) cast_f32_to_16_kernel(const float4* __restrict__ x, uint2* __restrict__ out, long long nvec, int fmt) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nvec; idx += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[idx];
    uint2 u;
    u.x = pack16(v.x, v.y, fmt);
    u.y = pack16(v.z, v.w, fmt);
    out[idx] = u;
  }
}

// ------------------------------------------------------------------------------------------------
// window alignment (least-squares scale/shift sums, affine + clamp, cross-fade)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lsq_sums_kernel(const float* __restrict__ p, const float* __restrict__ t, long long n, double* __restrict__ sums) {
  double a00 = 0, a01 = 0, b0 = 0, b1 = 0;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const double pv = p[idx], tv = t[idx];
    a00 += pv * pv;
    a01 += pv;
    b0 += pv * tv;
    b1 += tv;
  }
  __shared__ double red[4][8];
  double vals[4] = {a00, a01, b0, b1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double v = vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0;
    for (int i = 0; i < 8; ++i) v += red[threadIdx.x][i];
    const int slot = threadIdx.x == 0 ? 0 : threadIdx.x == 1 ? 1 : threadIdx.x == 2 ? 3 : 4;
    atomicAdd(&sums[slot], v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 4) atomicAdd(&sums[2], (double)n);
}

// 2x2 normal equations of the scale/shift fit (utils/util.py:40-62) solved on the device: no host round trip per window
__global__ void lsq_solve_kernel(const double* __restrict__ sums, float* __restrict__ ss) {
  const double a00 = sums[0], a01 = sums[1], a11 = sums[2], b0 = sums[3], b1 = sums[4];
  const double det = a00 * a11 - a01 * a01;
  double x0 = 1.0, x1 = 0.0;
  if (det != 0.0) {
    x0 = (a11 * b0 - a01 * b1) / det;
    x1 = (-a01 * b0 + a00 * b1) / det;
  }
  ss[0] = (float)x0;
  ss[1] = (float)x1;
}

__global__ void __launch_bounds__(256) affine_clamp_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, const float* __restrict__ ss) {
  const float sc = ss[0], sh = ss[1];
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
    out[idx] = fmaxf(x[idx] * sc + sh, 0.0f);
}

__global__ void __launch_bounds__(256)
crossfade_kernel(const float* __restrict__ pre, const float* __restrict__ post, float* __restrict__ out, long long n, const float* __restrict__ ss, float w) {
  const float sc = ss[0], sh = ss[1];
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const float pv = fmaxf(post[idx] * sc + sh, 0.0f);
    out[idx] = pre[idx] * (1.0f - w) + pv * w;
  }
}

// One launch finalises the output frames a window owns (video_depth.py:131-152, utils/util.py:65-74): slots 2..9 of window
// k >= 1 are cross-faded with the previous window's aligned slots 24..31, every later slot is scale * d + shift clamped at 0;
// window 0 is copied as it is.  (scale, shift) pairs are read from device memory: no host round trip.  VEC = floats per access.
template <int VEC>
struct FVec { float v[VEC]; };
template <> struct __align__(16) FVec<4> { float v[4]; };

template <int VEC>
__global__ void __launch_bounds__(256)
window_finalize_kernel(const float* __restrict__ cur, const float* __restrict__ prev_tail, const float* __restrict__ ss_cur,
                       const float* __restrict__ ss_prev, float* __restrict__ out, long long nv, int first_slot, int count, int is_first) {
  typedef FVec<VEC> V;
  float sc = 1.0f, sh = 0.0f, psc = 1.0f, psh = 0.0f;
  if (!is_first) { sc = ss_cur[0]; sh = ss_cur[1]; }
  const bool prev_affine = ss_prev != nullptr;
  if (prev_affine) { psc = ss_prev[0]; psh = ss_prev[1]; }
  const long long total = nv * count;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int fr = int(idx / nv);
    const long long px = idx - (long long)fr * nv;
    const int slot = first_slot + fr;
    V o = reinterpret_cast<const V*>(cur)[(long long)slot * nv + px];
    if (!is_first) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) o.v[i] = fmaxf(o.v[i] * sc + sh, 0.0f);
      if (slot < 10) {  // OVERLAP: slots ALIGN_LEN .. OVERLAP-1 blend with the predecessor's last INTERP_LEN frames
        const int j = slot - 2;
        const float w = j == 0 ? 0.0f : (j == 7 ? 1.0f : (float)(j * (1.0 / 7.0)));
        V p = reinterpret_cast<const V*>(prev_tail)[(long long)j * nv + px];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float pv = prev_affine ? fmaxf(p.v[i] * psc + psh, 0.0f) : p.v[i];
          o.v[i] = pv * (1.0f - w) + o.v[i] * w;
        }
      }
    }
    reinterpret_cast<V*>(out)[idx] = o;
  }
}

// the three key-frame depth maps the scale/shift chain needs from every window (slots 0, 1, 12): [3, n]
__global__ void __launch_bounds__(256) window_keys_kernel(const float* __restrict__ cur, float* __restrict__ keys, long long n) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < 3 * n; idx += (long long)gridDim.x * blockDim.x) {
    const int j = int(idx / n);
    const long long px = idx - (long long)j * n;
    keys[idx] = cur[(long long)(j == 2 ? 12 : j) * n + px];
  }
}

// ------------------------------------------------------------------------------------------------
// Sobel normals (reflect padding, kernel / 8): n = normalize(-Ix, -Iy, 1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(256)
sobel_normals_kernel(const float* __restrict__ depth, float* __restrict__ normals, int N, int H, int W, int ch) {
  const long long total = (long long)N * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = int(idx % W);
    long long t = idx / W;
    const int y = int(t % H);
    const long long n = t / H;
    const float* p = depth + n * (long long)H * W;
    const int ym = reflect(y - 1, H), yp = reflect(y + 1, H), xm = reflect(x - 1, W), xp = reflect(x + 1, W);
    const float a = p[ym * (long long)W + xm], b = p[ym * (long long)W + x], c = p[ym * (long long)W + xp];
    const float d = p[y * (long long)W + xm], f = p[y * (long long)W + xp];
    const float g = p[yp * (long long)W + xm], h = p[yp * (long long)W + x], i = p[yp * (long long)W + xp];
    const float ix = ((a - c) + 2.0f * (d - f) + (g - i)) * 0.125f;
    const float iy = ((a + 2.0f * b + c) - (g + 2.0f * h + i)) * 0.125f;
    const float nx = -ix, ny = -iy;
    const float inv = 1.0f / sqrtf(nx * nx + ny * ny + 1.0f + 1e-8f);
    float* o = normals + n * (long long)ch * H * W + (long long)y * W + x;
    o[0] = nx * inv;
    if (ch > 1) o[(long long)H * W] = ny * inv;
    if (ch > 2) o[2LL * H * W] = inv;
  }
}

// ------------------------------------------------------------------------------------------------
// frame pre-processing on the device: uint8 RGB HWC -> /255 -> cv2.INTER_CUBIC resize -> ImageNet normalise -> fp32 CHW
// (util/transform.py Resize + NormalizeImage + PrepareForNet as used at video_depth.py:74-99).  OpenCV's float cubic: A = -0.75,
// source coordinate (d + 0.5) * scale - 0.5, 4 taps at floor-1 .. floor+2 with replicated borders, horizontal pass then vertical.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float x, float (&c)[4]) {
  const float A = -0.75f;
  c[0] = ((A * (x + 1.0f) - 5.0f * A) * (x + 1.0f) + 8.0f * A) * (x + 1.0f) - 4.0f * A;
  c[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  c[2] = ((A + 2.0f) * (1.0f - x) - (A + 3.0f)) * (1.0f - x) * (1.0f - x) + 1.0f;
  c[3] = 1.0f - c[0] - c[1] - c[2];
}

__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ frames, float* __restrict__ out, int N, int H, int W, int h, int w, double scale_y, double scale_x,
                     float m0, float m1, float m2, float is0, float is1, float is2) {
  const long long total = (long long)N * h * w;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = int(idx % w);
    const int y = int((idx / w) % h);
    const long long n = idx / ((long long)w * h);
    float fx = (float)((x + 0.5) * scale_x - 0.5);
    float fy = (float)((y + 0.5) * scale_y - 0.5);
    const int sx = (int)floorf(fx), sy = (int)floorf(fy);
    fx -= (float)sx;
    fy -= (float)sy;
    float cx[4], cy[4];
    cubic_coeffs(fx, cx);
    cubic_coeffs(fy, cy);
    float acc[3] = {0.0f, 0.0f, 0.0f};
    const uint8_t* img = frames + n * (long long)H * W * 3;
    const bool identity = (H == h && W == w);  // cv2.resize returns the source unchanged for equal sizes
    if (identity) {
      const uint8_t* px = img + ((long long)y * W + x) * 3;
      acc[0] = (float)px[0] / 255.0f; acc[1] = (float)px[1] / 255.0f; acc[2] = (float)px[2] / 255.0f;
    } else {
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int yy = min(max(sy - 1 + ky, 0), H - 1);
        float row[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const int xx = min(max(sx - 1 + kx, 0), W - 1);
          const uint8_t* px = img + ((long long)yy * W + xx) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) row[c] = fmaf((float)px[c] / 255.0f, cx[kx], row[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = fmaf(row[c], cy[ky], acc[c]);
      }
    }
    float* o = out + n * 3LL * h * w + (long long)y * w + x;
    o[0] = (acc[0] - m0) * is0;
    o[(long long)h * w] = (acc[1] - m1) * is1;
    o[2LL * h * w] = (acc[2] - m2) * is2;
  }
}

// Equal source and network size (cv2.resize returns the source unchanged): four pixels per thread, 12 bytes in as three 32-bit
// words, one float4 per channel plane out; same arithmetic as the identity branch above.
__global__ void __launch_bounds__(256)
preprocess_u8_identity_kernel(const uint32_t* __restrict__ frames, float4* __restrict__ out, unsigned nquads, unsigned hw4, float m0, float m1, float m2,
                              float is0, float is1, float is2) {
  for (unsigned q = blockIdx.x * 256u + threadIdx.x; q < nquads; q += gridDim.x * 256u) {
    const unsigned n = q / hw4, p4 = q - n * hw4;
    const uint32_t a = __ldg(frames + 3ull * q), b = __ldg(frames + 3ull * q + 1), c = __ldg(frames + 3ull * q + 2);
    // little endian: a = R0 G0 B0 R1, b = G1 B1 R2 G2, c = B2 R3 G3 B3
    const float r[4] = {(float)(a & 255u), (float)(a >> 24), (float)((b >> 16) & 255u), (float)((c >> 8) & 255u)};
    const float g[4] = {(float)((a >> 8) & 255u), (float)(b & 255u), (float)(b >> 24), (float)((c >> 16) & 255u)};
    const float bl[4] = {(float)((a >> 16) & 255u), (float)((b >> 8) & 255u), (float)(c & 255u), (float)(c >> 24)};
    float4* o = out + (size_t)n * 3u * hw4 + p4;
    o[0] = make_float4((r[0] / 255.0f - m0) * is0, (r[1] / 255.0f - m0) * is0, (r[2] / 255.0f - m0) * is0, (r[3] / 255.0f - m0) * is0);
    o[hw4] = make_float4((g[0] / 255.0f - m1) * is1, (g[1] / 255.0f - m1) * is1, (g[2] / 255.0f - m1) * is1, (g[3] / 255.0f - m1) * is1);
    o[2 * (size_t)hw4] = make_float4((bl[0] / 255.0f - m2) * is2, (bl[1] / 255.0f - m2) * is2, (bl[2] / 255.0f - m2) * is2, (bl[3] / 255.0f - m2) * is2);
  }
}

}  // namespace vdn

using namespace vdn;
#define VDN_STREAM cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v)

extern "C" int vdn_layernorm(const float* x, const float* w, const float* b, void* out, int64_t rows, int32_t C, float eps, int32_t drop_first,
                             int32_t rows_per_batch, const float* pe, int32_t pe_len, void* stream_v) {
  VDN_STREAM;
  if (!x || !w || !b || !out) return set_error("vdn_layernorm: null pointer");
  if (C % 4 != 0 || C > 32 * LN_MAXV * 4) return set_error("vdn_layernorm: C must be a multiple of 4 and <= 1536");
  if (drop_first && (rows_per_batch < 2 || rows % rows_per_batch != 0)) return set_error("vdn_layernorm: bad rows_per_batch");
  if (pe && pe_len <= 0) return set_error("vdn_layernorm: bad pe_len");
  static const char* env = getenv("VDN_LN_V1");
  // the two-rows-per-warp kernel: plain, + positional table, or drop-first (not both at once); output rows fit 32 bits
  const bool fast = !(drop_first && pe != nullptr) && rows < 0x3fffffffLL && env == nullptr;
  const long long out_rows = drop_first ? rows / rows_per_batch * (rows_per_batch - 1) : rows;
  const int rpb = drop_first ? rows_per_batch : 0;
  if (fast && C == 1536) launch_layernorm_rows2<12>(x, w, b, out, out_rows, eps, pe, pe_len, rpb, stream);
  else if (fast && C == 1024) launch_layernorm_rows2<8>(x, w, b, out, out_rows, eps, pe, pe_len, rpb, stream);
  else if (fast && C == 768) launch_layernorm_rows2<6>(x, w, b, out, out_rows, eps, pe, pe_len, rpb, stream);
  else if (fast && C == 384) launch_layernorm_rows2<3>(x, w, b, out, out_rows, eps, pe, pe_len, rpb, stream);
  else if (fast && C == 256) launch_layernorm_rows2<2>(x, w, b, out, out_rows, eps, pe, pe_len, rpb, stream);
  else layernorm_kernel<<<grid_for(rows, 8), 256, 0, stream>>>(x, w, b, out, rows, C, eps, drop_first, rows_per_batch, pe, pe_len, get_operand_format());
  count_launch();
  return check_launch("layernorm_kernel");
}

extern "C" int vdn_groupnorm_stats(const void* x, float* stats, int32_t frames, int32_t D, int32_t C, int32_t groups, float eps, void* stream_v) {
  VDN_STREAM;
  if (!x || !stats) return set_error("vdn_groupnorm_stats: null pointer");
  if (groups <= 0 || C % groups != 0) return set_error("vdn_groupnorm_stats: C must be divisible by groups");
  if ((long long)D * C > 0x7fffffffLL) return set_error("vdn_groupnorm_stats: frame too large");
  if ((C / groups) % 8 == 0 && C % 8 == 0) groupnorm_stats_kernel<true><<<frames * groups, 256, 0, stream>>>(x, stats, D, C, groups, eps, get_operand_format());
  else groupnorm_stats_kernel<false><<<frames * groups, 256, 0, stream>>>(x, stats, D, C, groups, eps, get_operand_format());
  count_launch();
  return check_launch("groupnorm_stats_kernel");
}

extern "C" int vdn_groupnorm_apply_tc(const void* x, const float* stats, const float* w, const float* b, void* out, int32_t Bv, int32_t T, int32_t D,
                                      int32_t C, int32_t groups, void* stream_v) {
  VDN_STREAM;
  if (!x || !stats || !w || !b || !out) return set_error("vdn_groupnorm_apply_tc: null pointer");
  if (C % 8 != 0 || C % groups != 0) return set_error("vdn_groupnorm_apply_tc: C must be a multiple of 8 and of groups");
  const unsigned grid = grid_for((long long)Bv * T * D, 8);
  if ((C / groups) % 8 == 0) groupnorm_apply_tc_kernel<true><<<grid, 256, 0, stream>>>(x, stats, w, b, out, Bv, T, D, C, groups, get_operand_format());
  else groupnorm_apply_tc_kernel<false><<<grid, 256, 0, stream>>>(x, stats, w, b, out, Bv, T, D, C, groups, get_operand_format());
  count_launch();
  return check_launch("groupnorm_apply_tc_kernel");
}

extern "C" int vdn_groupnorm_to_tc(const void* x, const float* w, const float* b, void* out, float* stats, int32_t Bv, int32_t T, int32_t D,
                                   int32_t C, int32_t groups, float eps, void* stream_v) {
  VDN_STREAM;
  if (!x || !w || !b || !out || !stats) return set_error("vdn_groupnorm_to_tc: null pointer");
  if (Bv <= 0 || T <= 0 || D <= 0 || groups <= 0 || C % 8 != 0 || C % groups != 0) return set_error("vdn_groupnorm_to_tc: C must be a multiple of 8 and of groups");
  const int frames = Bv * T;
  const int nv = C / 8;
  static const char* env = getenv("VDN_GN_V1");
  static const char* fwd = getenv("VDN_GN_FWD");  // evaluation switch: pass 2 front to back
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  // one launch for the shapes of the ViT-L / ViT-g heads (C = 256, 512, 1024, 2048); a handful of frames (the streaming path) does not
  // fill the machine with 8 CTAs per frame and keeps the two-kernel form
  const bool fused = env == nullptr && aligned && (C / groups) % 8 == 0 && nv <= 256 && GN_THREADS % nv == 0 && groups <= GN_MAXG && frames >= 8 &&
                     (long long)frames * GN_CLUSTER < 0x7fffffffLL && (long long)D * nv < 0x7fffffffLL;
  if (!fused) {
    if (int rc = vdn_groupnorm_stats(x, stats, frames, D, C, groups, eps, stream_v)) return rc;
    return vdn_groupnorm_apply_tc(x, stats, w, b, out, Bv, T, D, C, groups, stream_v);
  }
  groupnorm_fused_tc_kernel<<<frames * GN_CLUSTER, GN_THREADS, 0, stream>>>(reinterpret_cast<const uint4*>(x), w, b, reinterpret_cast<uint4*>(out), stats, T, D, C,
                                                                    groups, eps, get_operand_format(), fwd == nullptr ? 1 : 0);
  count_launch();
  return check_launch("groupnorm_fused_tc_kernel");
}

extern "C" int vdn_patch_im2col(const float* img, void* out, int32_t B, int32_t H, int32_t W, int32_t Kp, void* stream_v) {
  VDN_STREAM;
  if (!img || !out) return set_error("vdn_patch_im2col: null pointer");
  if (H % 14 != 0 || W % 14 != 0) return set_error("vdn_patch_im2col: H and W must be multiples of the patch size 14");  // patch_embed.py:73-74
  if (Kp < 588 || Kp % 8 != 0) return set_error("vdn_patch_im2col: Kp must be >= 588 and a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(img) & 7) != 0 || (reinterpret_cast<uintptr_t>(out) & 3) != 0) return set_error("vdn_patch_im2col: img must be 8-byte and out 4-byte aligned");
  static const char* env = getenv("VDN_PATCH_V1");
  const int ph = H / 14, pw = W / 14;
  const int pitch = Kp / 2 + ((Kp / 2) % 8 == 0 ? 4 : 0);  // 32-bit words per shared-memory row, = 4 (mod 8): conflict-free 4 x 8 warp tiles
  const int maxseg = (48 * 1024) / (pitch * 4);            // patches whose output rows fit 48 KB of shared memory (40 at Kp = 592)
  if (env == nullptr && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && maxseg >= 1 && B > 0 && ph > 0 && pw > 0) {
    const int nseg = (pw + maxseg - 1) / maxseg;
    const int seg = (pw + nseg - 1) / nseg;
    const long long blocks = (long long)B * ph * nseg;
    if (blocks < 0x7fffffffLL) {
      patch_im2col_rows_kernel<<<(unsigned)blocks, 256, (size_t)seg * pitch * 4, stream>>>(img, reinterpret_cast<uint4*>(out), H, W, Kp, pitch, seg, nseg,
                                                                                           get_operand_format());
      count_launch();
      return check_launch("patch_im2col_rows_kernel");
    }
  }
  patch_im2col_kernel<<<grid_for((long long)B * (H / 14) * (W / 14) * 42, 256), 256, 0, stream>>>(img, out, B, H, W, Kp, get_operand_format());
  count_launch();
  return check_launch("patch_im2col_kernel");
}

extern "C" int vdn_write_cls(float* x, const float* cls, const float* pos, int32_t B, int32_t tokens, int32_t C, void* stream_v) {
  VDN_STREAM;
  if (!x || !cls || !pos) return set_error("vdn_write_cls: null pointer");
  write_cls_kernel<<<grid_for((long long)B * C, 256), 256, 0, stream>>>(x, cls, pos, B, tokens, C);
  count_launch();
  return check_launch("write_cls_kernel");
}

extern "C" int vdn_readout_concat(const void* tok, int64_t tok_frame_pitch, const void* cls, int64_t cls_frame_pitch, void* out, int64_t frames,
                                  int32_t P, int32_t C, void* stream_v) {
  VDN_STREAM;
  if (!tok || !cls || !out) return set_error("vdn_readout_concat: null pointer");
  if (C % 8 != 0 || P < 1 || frames <= 0 || tok_frame_pitch < P || cls_frame_pitch < 1) return set_error("vdn_readout_concat: C must be a multiple of 8, pitches >= extents");
  if (((reinterpret_cast<uintptr_t>(tok) | reinterpret_cast<uintptr_t>(cls) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return set_error("vdn_readout_concat: 16-byte alignment");
  readout_concat_kernel<<<grid_for(frames * P * 2 * (C / 8), 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(tok), tok_frame_pitch,
                                                                                    reinterpret_cast<const uint4*>(cls), cls_frame_pitch,
                                                                                    reinterpret_cast<uint4*>(out), frames, P, C / 8);
  count_launch();
  return check_launch("readout_concat_kernel");
}

extern "C" int vdn_im2col_3x3_s2(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C, void* stream_v) {
  VDN_STREAM;
  if (!x || !out) return set_error("vdn_im2col_3x3_s2: null pointer");
  if (C % 8 != 0) return set_error("vdn_im2col_3x3_s2: C must be a multiple of 8");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  static const char* env = getenv("VDN_IM2COL_V1");
  const long long total_pix = (long long)B * Ho * Wo;
  const int cv = C / 8;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (env == nullptr && aligned && total_pix > 0 && total_pix < 0x7fffffffLL) {
    constexpr int NT = 128;
    const int ppb = cv >= NT ? 1 : NT / cv;
    im2col_3x3_s2_pix_kernel<NT><<<(unsigned)((total_pix + ppb - 1) / ppb), NT, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), H,
                                                                                          W, Ho, Wo, cv, (int)total_pix);
    count_launch();
    return check_launch("im2col_3x3_s2_pix_kernel");
  }
  im2col_3x3_s2_kernel<<<grid_for((long long)B * Ho * Wo * 9 * (C / 8), 256), 256, 0, stream>>>(x, out, B, H, W, C);
  count_launch();
  return check_launch("im2col_3x3_s2_kernel");
}

static int launch_bilinear_nhwc(const void* x, void* out, void* out_relu, int B, int H, int W, int Ho, int Wo, int C, int relu_out, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return set_error("vdn_bilinear_nhwc: bad shape");
  if (Ho > 65535 || B > 65535) return set_error("vdn_bilinear_nhwc: Ho and B must fit a grid dimension");
  const int cv = C / 8;
  const uint4* xi = reinterpret_cast<const uint4*>(x);
  uint4* o = reinterpret_cast<uint4*>(out);
  uint4* orl = reinterpret_cast<uint4*>(out_relu);
  const int fmt = get_operand_format();
  static const char* env_run = getenv("VDN_BILINEAR_RUN");  // run length override (output pixels per thread)
  const int want = env_run && atoi(env_run) > 0 ? atoi(env_run) : (cv % 32 == 0 ? 16 : 64);  // measured: short runs win when a warp is one run
  // the number of runs is rounded so that a row's threads fill whole 128-thread blocks
  int mult = 1;
  while ((mult * cv) % 128 != 0 && mult < 8) mult <<= 1;
  int nseg = (Wo + want - 1) / want;
  nseg = (nseg + mult - 1) / mult * mult;
  if (nseg > Wo) nseg = Wo;
  const int L = (Wo + nseg - 1) / nseg;
  nseg = (Wo + L - 1) / L;
  dim3 g((unsigned)((nseg * cv + 127) / 128), Ho, B);
#define VDN_BILS(F, R) bilinear_slide_kernel<F, R><<<g, 128, 0, stream>>>(xi, o, orl, H, W, Ho, Wo, cv, L, nseg, relu_out)
  if (orl != nullptr) { if (fmt) VDN_BILS(1, 1); else VDN_BILS(0, 1); }
  else { if (fmt) VDN_BILS(1, 0); else VDN_BILS(0, 0); }
#undef VDN_BILS
  count_launch();
  return check_launch("bilinear_slide_kernel");
}

extern "C" int vdn_bilinear_nhwc(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t Ho, int32_t Wo, int32_t C, int32_t relu_out,
                                 void* stream_v) {
  VDN_STREAM;
  if (!x || !out) return set_error("vdn_bilinear_nhwc: null pointer");
  if (C % 8 != 0) return set_error("vdn_bilinear_nhwc: C must be a multiple of 8");
  return launch_bilinear_nhwc(x, out, nullptr, B, H, W, Ho, Wo, C, relu_out, stream);
}

extern "C" int vdn_bilinear_nhwc2(const void* x, void* out, void* out_relu, int32_t B, int32_t H, int32_t W, int32_t Ho, int32_t Wo, int32_t C,
                                  void* stream_v) {
  VDN_STREAM;
  if (!x || !out || !out_relu) return set_error("vdn_bilinear_nhwc2: null pointer");
  if (C % 8 != 0) return set_error("vdn_bilinear_nhwc2: C must be a multiple of 8");
  return launch_bilinear_nhwc(x, out, out_relu, B, H, W, Ho, Wo, C, 0, stream);
}

extern "C" int vdn_bilinear_f32(const float* x, float* out, int32_t N, int32_t H, int32_t W, int32_t Ho, int32_t Wo, int32_t relu, void* stream_v) {
  VDN_STREAM;
  if (!x || !out) return set_error("vdn_bilinear_f32: null pointer");
  bilinear_f32_kernel<<<grid_for((long long)N * Ho * Wo, 256, 32), 256, 0, stream>>>(x, out, N, H, W, Ho, Wo, relu);
  count_launch();
  return check_launch("bilinear_f32_kernel");
}

extern "C" int vdn_relu16(const void* x, void* out, int64_t n, void* stream_v) {
  VDN_STREAM;
  if (!x || !out) return set_error("vdn_relu16: null pointer");
  if (n % 8 != 0) return set_error("vdn_relu16: n must be a multiple of 8");
  relu16_kernel<<<grid_for(n / 8, 256, 32), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), n / 8, get_operand_format());
  count_launch();
  return check_launch("relu16_kernel");
}

extern "C" int vdn_cast_f32_to_16(const float* x, void* out, int64_t n, void* stream_v) {
  VDN_STREAM;
  if (!x || !out) return set_error("vdn_cast_f32_to_16: null pointer");
  if (n % 4 != 0) return set_error("vdn_cast_f32_to_16: n must be a multiple of 4");
  cast_f32_to_16_kernel<<<grid_for(n / 4, 256, 32), 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(out), n / 4,
                                                                    get_operand_format());
  count_launch();
  return check_launch("cast_f32_to_16_kernel");
}

extern "C" int vdn_lsq_sums(const float* pred, const float* target, int64_t n, double* sums5, void* stream_v) {
  VDN_STREAM;
  if (!pred || !target || !sums5) return set_error("vdn_lsq_sums: null pointer");
  cudaError_t e = cudaMemsetAsync(sums5, 0, 5 * sizeof(double), stream);
  if (e != cudaSuccess) return set_error(std::string("vdn_lsq_sums memset: ") + cudaGetErrorString(e));
  lsq_sums_kernel<<<grid_for(n, 256 * 8, 4), 256, 0, stream>>>(pred, target, n, sums5);
  count_launch();
  return check_launch("lsq_sums_kernel");
}

extern "C" int vdn_lsq_solve(const double* sums5, float* scale_shift, void* stream_v) {
  VDN_STREAM;
  if (!sums5 || !scale_shift) return set_error("vdn_lsq_solve: null pointer");
  lsq_solve_kernel<<<1, 1, 0, stream>>>(sums5, scale_shift);
  count_launch();
  return check_launch("lsq_solve_kernel");
}

extern "C" int vdn_affine_clamp(const float* x, float* out, int64_t n, const float* scale_shift, void* stream_v) {
  VDN_STREAM;
  if (!x || !out || !scale_shift) return set_error("vdn_affine_clamp: null pointer");
  affine_clamp_kernel<<<grid_for(n, 256 * 4), 256, 0, stream>>>(x, out, n, scale_shift);
  count_launch();
  return check_launch("affine_clamp_kernel");
}

extern "C" int vdn_window_finalize(const float* cur, const float* prev_tail, const float* ss_cur, const float* ss_prev, float* out, int64_t n,
                                   int first_slot, int count, int is_first, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!cur || !out) return set_error("vdn_window_finalize: null pointer");
  if (n <= 0) return set_error("vdn_window_finalize: empty frame");
  if (first_slot < 0 || count < 0 || first_slot + count > 32) return set_error("vdn_window_finalize: slots out of range");
  if (!is_first && (!ss_cur || (first_slot < 10 && (!prev_tail || first_slot < 2)))) return set_error("vdn_window_finalize: missing operand for the cross-fade");
  if (count == 0) return 0;
  const bool vec = n % 4 == 0 && (reinterpret_cast<uintptr_t>(cur) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(prev_tail)) % 16 == 0;
  if (vec)
    window_finalize_kernel<4><<<grid_for(n / 4 * count, 256 * 2), 256, 0, stream>>>(cur, prev_tail, ss_cur, ss_prev, out, n / 4, first_slot, count, is_first);
  else
    window_finalize_kernel<1><<<grid_for(n * count, 256 * 4), 256, 0, stream>>>(cur, prev_tail, ss_cur, ss_prev, out, n, first_slot, count, is_first);
  count_launch();
  return check_launch("window_finalize_kernel");
}

extern "C" int vdn_window_keys(const float* cur, float* keys, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!cur || !keys) return set_error("vdn_window_keys: null pointer");
  if (n <= 0) return set_error("vdn_window_keys: empty frame");
  window_keys_kernel<<<grid_for(3 * n, 256 * 4), 256, 0, stream>>>(cur, keys, n);
  count_launch();
  return check_launch("window_keys_kernel");
}

extern "C" int vdn_crossfade(const float* pre, const float* post, float* out, int64_t n, const float* scale_shift, float w, void* stream_v) {
  VDN_STREAM;
  if (!pre || !post || !out || !scale_shift) return set_error("vdn_crossfade: null pointer");
  crossfade_kernel<<<grid_for(n, 256 * 4), 256, 0, stream>>>(pre, post, out, n, scale_shift, w);
  count_launch();
  return check_launch("crossfade_kernel");
}

extern "C" int vdn_sobel_normals(const float* depth, float* normals, int32_t N, int32_t H, int32_t W, int32_t channels_out, void* stream_v) {
  VDN_STREAM;
  if (!depth || !normals) return set_error("vdn_sobel_normals: null pointer");
  if (H < 2 || W < 2 || channels_out < 1 || channels_out > 3) return set_error("vdn_sobel_normals: bad shape");
  sobel_normals_kernel<<<grid_for((long long)N * H * W, 256, 32), 256, 0, stream>>>(depth, normals, N, H, W, channels_out);
  count_launch();
  return check_launch("sobel_normals_kernel");
}

extern "C" int vdn_preprocess_u8(const void* frames, float* out, int32_t N, int32_t H, int32_t W, int32_t h, int32_t w, const float* mean3,
                                 const float* std3, void* stream_v) {
  VDN_STREAM;
  if (!frames || !out || !mean3 || !std3) return set_error("vdn_preprocess_u8: null pointer");
  if (N <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0) return set_error("vdn_preprocess_u8: bad shape");
  const long long hw = (long long)h * w;
  if (H == h && W == w && hw % 4 == 0 && (long long)N * hw / 4 < 0x7fffffffLL && (reinterpret_cast<uintptr_t>(frames) & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const unsigned nquads = (unsigned)((long long)N * hw / 4);
    preprocess_u8_identity_kernel<<<grid_for(nquads, 256, 32), 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(frames), reinterpret_cast<float4*>(out),
                                                                                  nquads, (unsigned)(hw / 4), mean3[0], mean3[1], mean3[2],
                                                                                  1.0f / std3[0], 1.0f / std3[1], 1.0f / std3[2]);
    count_launch();
    return check_launch("preprocess_u8_identity_kernel");
  }
  const double sy = 1.0 / ((double)h / (double)H), sx = 1.0 / ((double)w / (double)W);  // OpenCV: scale = 1 / inv_scale
  preprocess_u8_kernel<<<grid_for((long long)N * h * w, 256, 32), 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(frames), out, N, H, W, h, w, sy, sx,
                                                                                      mean3[0], mean3[1], mean3[2], 1.0f / std3[0], 1.0f / std3[1],
                                                                                      1.0f / std3[2]);
  count_launch();
  return check_launch("preprocess_u8_kernel");
}

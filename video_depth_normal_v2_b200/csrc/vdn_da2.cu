// Bandwidth-bound pieces of the DepthAnythingV2 memory block (depth_anything_v2/memory_block.py, sam2/modeling/*):
//   vdn_rope2d         : axial rotary encoding of q / k heads in place (position_encoding.py:186-239)
//   vdn_add_rowvec     : out_f32 = x + alpha * vec   (pos_enc_at_input, memory_attention.py:137-138; constant empty-bank cross-attention term)
//   vdn_add_rowscalar  : x[r, :] += m[r]             (pix_feat + downsampled mask, memory_encoder.py:173)
//   vdn_dwconv7_ln     : depthwise 7x7 conv + LayerNorm2d over channels, NHWC fp32 -> 16-bit rows (memory_encoder.py:96-99)
//   vdn_mask_down1 / 2 : the two MaskDownSamplers on the 1-channel sigmoid(depth) map (memory_encoder.py:17-60, memory_block.py:68-71)
#include <stdlib.h>

#include "../../include/vdn_b200.h"
#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

// IDX: index type of the (row, head, piece) decode — unsigned 32-bit whenever the element count allows (the four 64-bit divisions per
// 16 bytes of the long long form made the kernel issue-bound: 36 us per 45 MB call)
template <int FMT, typename IDX>
__global__ void __launch_bounds__(256)
rope2d_kernel(uint4* __restrict__ x, long long rows, int ld8 /*row pitch in 16-byte units*/, int col8 /*first column / 8*/, int heads,
              const float* __restrict__ cs, int P, long long rows_per_batch, long long batch_pitch, int per_head /*table row = (pos, head)*/) {
  const IDX total = (IDX)(rows * heads * 8);
  const IDX rpb = (IDX)rows_per_batch;
  for (IDX idx = (IDX)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (IDX)gridDim.x * blockDim.x) {
    const int v = int(idx & 7);  // 16-byte piece of the 64-wide head: pairs 4v .. 4v+3
    const IDX t = idx >> 3;
    const int h = int(t % (IDX)heads);
    const IDX r = t / (IDX)heads;
    const int pos = int(r % (IDX)P);
    const IDX bb = r / rpb;
    const long long phys = (long long)bb * batch_pitch + (long long)(r - bb * rpb);  // rows of one batch are contiguous, batches batch_pitch rows apart
    uint4* p = x + phys * ld8 + col8 + h * 8 + v;
    uint4 u = *p;
    uint32_t* w = &u.x;
    const long long trow = per_head ? (long long)pos * heads + h : (long long)pos;
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs + trow * 64 + 4 * v));
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(cs + trow * 64 + 32 + 4 * v));
    const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 ab = T16f<FMT>::unpack(w[i]);
      w[i] = T16f<FMT>::pack(ab.x * cc[i] - ab.y * ss[i], ab.x * ss[i] + ab.y * cc[i]);
    }
    *p = u;
  }
}

__global__ void __launch_bounds__(256)
add_rowvec_kernel(const void* __restrict__ x, int x_f32, const float* __restrict__ vec, float alpha, float* __restrict__ out, long long rows, int C, int fmt) {
  const long long total = rows * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = int(idx % C);
    const float v = x_f32 ? reinterpret_cast<const float*>(x)[idx] : load16(x, idx, fmt);
    out[idx] = v + alpha * __ldg(vec + c);
  }
}

__global__ void __launch_bounds__(256) add_rowscalar_kernel(float* __restrict__ x, const float* __restrict__ m, long long rows, int C) {
  const long long total = rows * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) x[idx] += __ldg(m + idx / C);
}

constexpr int DW_MAXC = 6;  // channels per thread (C <= 1536 with 256 threads: ViT-g)

__global__ void __launch_bounds__(256)
dwconv7_ln_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ ln_w,
                  const float* __restrict__ ln_b, void* __restrict__ out, int H, int W, int C, float eps, int fmt) {
  const long long pix = blockIdx.x;
  const int xx = int(pix % W);
  const int yy = int((pix / W) % H);
  const long long img = pix / ((long long)W * H);
  const float* base = x + img * (long long)H * W * C;
  float acc[DW_MAXC];
#pragma unroll
  for (int i = 0; i < DW_MAXC; ++i) {
    const int c = threadIdx.x + i * 256;
    acc[i] = c < C ? __ldg(bias + c) : 0.0f;
  }
  for (int dy = -3; dy <= 3; ++dy) {
    const int y2 = yy + dy;
    if (y2 < 0 || y2 >= H) continue;
    for (int dx = -3; dx <= 3; ++dx) {
      const int x2 = xx + dx;
      if (x2 < 0 || x2 >= W) continue;
      const float* src = base + ((long long)y2 * W + x2) * C;
      const float* wt = w + ((dy + 3) * 7 + (dx + 3)) * C;
#pragma unroll
      for (int i = 0; i < DW_MAXC; ++i) {
        const int c = threadIdx.x + i * 256;
        if (c < C) acc[i] = fmaf(src[c], __ldg(wt + c), acc[i]);
      }
    }
  }
  __shared__ float red[8];
  __shared__ float stat;
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < DW_MAXC; ++i)
    if (threadIdx.x + i * 256 < C) s += acc[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    stat = t / (float)C;
  }
  __syncthreads();
  const float mean = stat;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < DW_MAXC; ++i)
    if (threadIdx.x + i * 256 < C) q = fmaf(acc[i] - mean, acc[i] - mean, q);
  q = warp_sum(q);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    stat = rsqrtf(t / (float)C + eps);
  }
  __syncthreads();
  const float rstd = stat;
#pragma unroll
  for (int i = 0; i < DW_MAXC; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < C) store16(out, pix * C + c, (acc[i] - mean) * rstd * __ldg(ln_w + c) + __ldg(ln_b + c), fmt);
  }
}

// Row-segment form of the same op: a block owns DW_P consecutive pixels of one image row, a thread four channels.  Every input
// vector of the 7 x (DW_P + 6) neighbourhood is loaded once and feeds up to 7 output pixels (the block-per-pixel kernel above read
// 49 neighbours per output from L2: 8 TB/s of L2 traffic, 247 GB/s of algorithmic bytes).  LayerNorm2d over the channels of each pixel:
// two block reductions for all DW_P pixels at once.  Needs C % 128 == 0 (blockDim = C / 4).
constexpr int DW_P = 13;
template <int MAXT>  // 256 threads (C <= 1024): two blocks per SM; 384 (ViT-g, C = 1536): one
__global__ void __launch_bounds__(MAXT, MAXT <= 256 ? 2 : 1)
dwconv7_ln_row_kernel(const float4* __restrict__ x, const float4* __restrict__ w, const float4* __restrict__ bias, const float4* __restrict__ ln_w,
                      const float4* __restrict__ ln_b, void* __restrict__ out, int H, int W, int C4, float eps, int fmt) {
  const int segs = (W + DW_P - 1) / DW_P;
  const int seg = blockIdx.x % segs;
  const int yy = (blockIdx.x / segs) % H;
  const long long img = blockIdx.x / ((long long)segs * H);
  const int x0 = seg * DW_P;
  const int c = threadIdx.x;  // channel vector
  const float4* base = x + img * (long long)H * W * C4 + c;
  float4 acc[DW_P];
  const float4 b4 = __ldg(bias + c);
#pragma unroll
  for (int p = 0; p < DW_P; ++p) acc[p] = b4;
  for (int dy = 0; dy < 7; ++dy) {
    const int y2 = yy + dy - 3;
    if (y2 < 0 || y2 >= H) continue;
    float4 wt[7];
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) wt[dx] = __ldg(w + (dy * 7 + dx) * C4 + c);
    const float4* row = base + (long long)y2 * W * C4;
    // input column x0 - 3 + xi feeds output pixel p = xi - dx with tap dx.  The columns are fetched DW_B at a time before any of them
    // is used (columns outside the image read as zero: an exact no-op in the FMAs): with one conditional load per column right in
    // front of its FMAs every one of the 7 x 19 loads of a block exposed its full L2 latency (ncu: 65 % of the stall samples on
    // long-scoreboard, 236 us per launch)
    constexpr int DW_B = 7;
#pragma unroll
    for (int b0 = 0; b0 < DW_P + 6; b0 += DW_B) {
      float4 v[DW_B];
#pragma unroll
      for (int i = 0; i < DW_B; ++i) {
        const int x2 = x0 - 3 + b0 + i;
        v[i] = (b0 + i < DW_P + 6 && x2 >= 0 && x2 < W) ? __ldg(row + (long long)x2 * C4) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
#pragma unroll
      for (int i = 0; i < DW_B; ++i) {
        const int xi = b0 + i;
        if (xi < DW_P + 6) {
#pragma unroll
          for (int dx = 0; dx < 7; ++dx) {
            const int p = xi - dx;
            if (p >= 0 && p < DW_P) {
              acc[p].x = fmaf(v[i].x, wt[dx].x, acc[p].x);
              acc[p].y = fmaf(v[i].y, wt[dx].y, acc[p].y);
              acc[p].z = fmaf(v[i].z, wt[dx].z, acc[p].z);
              acc[p].w = fmaf(v[i].w, wt[dx].w, acc[p].w);
            }
          }
        }
      }
    }
  }
  __shared__ float red[12][DW_P];
  __shared__ float stat[DW_P];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float invC = 1.0f / (float)(4 * C4);
#pragma unroll
  for (int p = 0; p < DW_P; ++p) {
    const float s = warp_sum((acc[p].x + acc[p].y) + (acc[p].z + acc[p].w));
    if (lane == 0) red[warp][p] = s;
  }
  __syncthreads();
  if (threadIdx.x < DW_P) {
    float t = 0.0f;
    for (int i = 0; i < nwarps; ++i) t += red[i][threadIdx.x];
    stat[threadIdx.x] = t * invC;
  }
  __syncthreads();
  float mean[DW_P];
#pragma unroll
  for (int p = 0; p < DW_P; ++p) {
    mean[p] = stat[p];
    const float dx_ = acc[p].x - mean[p], dy_ = acc[p].y - mean[p], dz_ = acc[p].z - mean[p], dw_ = acc[p].w - mean[p];
    const float q = warp_sum((dx_ * dx_ + dy_ * dy_) + (dz_ * dz_ + dw_ * dw_));
    if (lane == 0) red[warp][p] = q;  // every warp has read stat[] into registers and red[] was consumed before the barrier above
  }
  __syncthreads();
  if (threadIdx.x < DW_P) {
    float t = 0.0f;
    for (int i = 0; i < nwarps; ++i) t += red[i][threadIdx.x];
    stat[threadIdx.x] = rsqrtf(t * invC + eps);
  }
  __syncthreads();
  const float4 g = __ldg(ln_w + c), be = __ldg(ln_b + c);
#pragma unroll
  for (int p = 0; p < DW_P; ++p) {
    if (x0 + p >= W) break;
    const float rstd = stat[p];
    const long long pix = (img * H + yy) * (long long)W + x0 + p;
    const uint32_t lo = pack16((acc[p].x - mean[p]) * rstd * g.x + be.x, (acc[p].y - mean[p]) * rstd * g.y + be.y, fmt);
    const uint32_t hi = pack16((acc[p].z - mean[p]) * rstd * g.z + be.z, (acc[p].w - mean[p]) * rstd * g.w + be.w, fmt);
    reinterpret_cast<uint2*>(out)[pix * C4 + c] = make_uint2(lo, hi);
  }
}

__device__ __forceinline__ float gelu_exact(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

// sigmoid -> conv 3x3 / stride 2 / pad 1 (1 -> 4) -> LayerNorm2d(4) -> GELU -> conv 1x1 (4 -> 1)
__global__ void __launch_bounds__(256)
mask_down1_kernel(const float* __restrict__ depth, const float* __restrict__ prm, float* __restrict__ out, int B, int H, int W, int Ho, int Wo) {
  // prm: w0[4][9], b0[4], lnw[4], lnb[4], w1[4], b1
  const long long total = (long long)B * Ho * Wo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int xo = int(idx % Wo);
    const int yo = int((idx / Wo) % Ho);
    const long long b = idx / ((long long)Wo * Ho);
    const float* p = depth + b * (long long)H * W;
    float t[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int y = 2 * yo - 1 + r, x = 2 * xo - 1 + s;
        t[r * 3 + s] = (y >= 0 && y < H && x >= 0 && x < W) ? 1.0f / (1.0f + expf(-p[(long long)y * W + x])) : 0.0f;
      }
    float c[4], mean = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = prm[36 + k];
#pragma unroll
      for (int j = 0; j < 9; ++j) a = fmaf(prm[k * 9 + j], t[j], a);
      c[k] = a;
      mean += a;
    }
    mean *= 0.25f;
    float var = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) var += (c[k] - mean) * (c[k] - mean);
    const float rstd = 1.0f / sqrtf(var * 0.25f + 1e-6f);
    float o = prm[52];
#pragma unroll
    for (int k = 0; k < 4; ++k) o = fmaf(prm[48 + k], gelu_exact((c[k] - mean) * rstd * prm[40 + k] + prm[44 + k]), o);
    out[idx] = o;
  }
}

// conv 7x7 / stride 7 (1 -> 49) -> LayerNorm2d(49) -> GELU -> conv 1x1 (49 -> 1); one warp per output pixel
__global__ void __launch_bounds__(256)
mask_down2_kernel(const float* __restrict__ in, const float* __restrict__ prm, float* __restrict__ out, int B, int Hi, int Wi, int Ho, int Wo) {
  // prm: w0[49][49], b0[49], lnw[49], lnb[49], w1[49], b1
  __shared__ float sp[49 * 49 + 4 * 49 + 1];
  for (int i = threadIdx.x; i < 49 * 49 + 4 * 49 + 1; i += blockDim.x) sp[i] = prm[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * Ho * Wo;
  for (long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pix < total; pix += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int xo = int(pix % Wo);
    const int yo = int((pix / Wo) % Ho);
    const long long b = pix / ((long long)Wo * Ho);
    const float* p = in + b * (long long)Hi * Wi + (long long)(7 * yo) * Wi + 7 * xo;
    // lane handles channels lane and lane + 32 (< 49)
    float c0 = sp[2401 + lane], c1 = lane + 32 < 49 ? sp[2401 + lane + 32] : 0.0f;
    for (int j = 0; j < 49; ++j) {
      const float t = p[(long long)(j / 7) * Wi + (j % 7)];
      c0 = fmaf(sp[lane * 49 + j], t, c0);
      if (lane + 32 < 49) c1 = fmaf(sp[(lane + 32) * 49 + j], t, c1);
    }
    const bool has1 = lane + 32 < 49;
    const float mean = warp_sum(c0 + (has1 ? c1 : 0.0f)) * (1.0f / 49.0f);
    const float d0 = c0 - mean, d1 = has1 ? c1 - mean : 0.0f;
    const float rstd = 1.0f / sqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.0f / 49.0f) + 1e-6f);
    float o = sp[2401 + 3 * 49 + lane] * gelu_exact(d0 * rstd * sp[2401 + 49 + lane] + sp[2401 + 2 * 49 + lane]);
    if (has1) o += sp[2401 + 3 * 49 + lane + 32] * gelu_exact(d1 * rstd * sp[2401 + 49 + lane + 32] + sp[2401 + 2 * 49 + lane + 32]);
    o = warp_sum(o);
    if (lane == 0) out[pix] = o + sp[2401 + 4 * 49];
  }
}

static inline unsigned blocks_for(long long work, int per_block) {
  long long b = (work + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace vdn

using namespace vdn;
#define VDN_STREAM cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v)

static void launch_rope2d(unsigned grid, cudaStream_t stream, uint4* x, long long rows, int ld8, int col8, int heads, const float* cs, int P, long long rpb,
                          long long pitch, int per_head) {
  // 32-bit decode when every index (and idx + one grid stride) fits: total + grid * 256 < 2^32
  const bool small = rows * heads * 8 + (long long)grid * 256 < 0xffffffffLL && rpb < 0xffffffffLL;
  const int fmt = get_operand_format();
  if (small) {
    if (fmt) rope2d_kernel<1, unsigned><<<grid, 256, 0, stream>>>(x, rows, ld8, col8, heads, cs, P, rpb, pitch, per_head);
    else rope2d_kernel<0, unsigned><<<grid, 256, 0, stream>>>(x, rows, ld8, col8, heads, cs, P, rpb, pitch, per_head);
  } else {
    if (fmt) rope2d_kernel<1, long long><<<grid, 256, 0, stream>>>(x, rows, ld8, col8, heads, cs, P, rpb, pitch, per_head);
    else rope2d_kernel<0, long long><<<grid, 256, 0, stream>>>(x, rows, ld8, col8, heads, cs, P, rpb, pitch, per_head);
  }
}

extern "C" int vdn_rope2d(void* x, int64_t rows, int64_t ld, int32_t col0, int32_t heads, const float* cos_sin, int32_t P, int64_t rows_per_batch,
                          int64_t batch_pitch, void* stream_v) {
  VDN_STREAM;
  if (!x || !cos_sin) return set_error("vdn_rope2d: null pointer");
  if (ld % 8 != 0 || col0 % 8 != 0 || heads <= 0 || P <= 0 || rows <= 0) return set_error("vdn_rope2d: bad geometry");
  if (col0 + (int64_t)heads * 64 > ld) return set_error("vdn_rope2d: heads exceed the row");
  if (rows_per_batch <= 0) { rows_per_batch = rows; batch_pitch = rows; }
  if (rows_per_batch % P != 0 || batch_pitch < rows_per_batch) return set_error("vdn_rope2d: rows_per_batch must be a multiple of P and <= batch_pitch");
  const unsigned grid = blocks_for(rows * heads * 8, 256);
  launch_rope2d(grid, stream, reinterpret_cast<uint4*>(x), rows, (int)(ld / 8), col0 / 8, heads, cos_sin, P, rows_per_batch, batch_pitch, 0);
  count_launch();
  return check_launch("rope2d_kernel");
}

extern "C" int vdn_rope_chunks(void* x, int64_t rows, int64_t ld, int32_t col0, int32_t chunks, const float* cos_sin, int32_t P, void* stream_v) {
  VDN_STREAM;
  if (!x || !cos_sin) return set_error("vdn_rope_chunks: null pointer");
  if (ld % 8 != 0 || col0 % 8 != 0 || chunks <= 0 || P <= 0 || rows <= 0) return set_error("vdn_rope_chunks: bad geometry");
  if (col0 + (int64_t)chunks * 64 > ld) return set_error("vdn_rope_chunks: chunks exceed the row");
  const unsigned grid = blocks_for(rows * chunks * 8, 256);
  launch_rope2d(grid, stream, reinterpret_cast<uint4*>(x), rows, (int)(ld / 8), col0 / 8, chunks, cos_sin, P, rows, rows, 1);
  count_launch();
  return check_launch("rope2d_kernel");
}

extern "C" int vdn_add_rowvec(const void* x, int32_t x_f32, const float* vec, float alpha, float* out, int64_t rows, int32_t C, void* stream_v) {
  VDN_STREAM;
  if (!x || !vec || !out) return set_error("vdn_add_rowvec: null pointer");
  add_rowvec_kernel<<<blocks_for(rows * C, 256), 256, 0, stream>>>(x, x_f32, vec, alpha, out, rows, C, get_operand_format());
  count_launch();
  return check_launch("add_rowvec_kernel");
}

extern "C" int vdn_add_rowscalar(float* x, const float* m, int64_t rows, int32_t C, void* stream_v) {
  VDN_STREAM;
  if (!x || !m) return set_error("vdn_add_rowscalar: null pointer");
  add_rowscalar_kernel<<<blocks_for(rows * C, 256), 256, 0, stream>>>(x, m, rows, C);
  count_launch();
  return check_launch("add_rowscalar_kernel");
}

extern "C" int vdn_dwconv7_ln(const float* x, const float* w, const float* bias, const float* ln_w, const float* ln_b, void* out, int32_t B, int32_t H,
                              int32_t W, int32_t C, float eps, void* stream_v) {
  VDN_STREAM;
  if (!x || !w || !bias || !ln_w || !ln_b || !out) return set_error("vdn_dwconv7_ln: null pointer");
  if (C <= 0 || C > 256 * DW_MAXC) return set_error("vdn_dwconv7_ln: C must be <= 1536");
  static const bool v1 = [] { const char* e = getenv("VDN_DWCONV_V1"); return e && atoi(e) != 0; }();
  if (C % 128 == 0 && !v1) {
    const long long blocks = (long long)B * H * ((W + DW_P - 1) / DW_P);
    if (C <= 1024)
      dwconv7_ln_row_kernel<256><<<(unsigned)blocks, C / 4, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(w),
                                                                       reinterpret_cast<const float4*>(bias), reinterpret_cast<const float4*>(ln_w),
                                                                       reinterpret_cast<const float4*>(ln_b), out, H, W, C / 4, eps, get_operand_format());
    else
      dwconv7_ln_row_kernel<384><<<(unsigned)blocks, C / 4, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(w),
                                                                       reinterpret_cast<const float4*>(bias), reinterpret_cast<const float4*>(ln_w),
                                                                       reinterpret_cast<const float4*>(ln_b), out, H, W, C / 4, eps, get_operand_format());
    count_launch();
    return check_launch("dwconv7_ln_row_kernel");
  }
  dwconv7_ln_kernel<<<(unsigned)((long long)B * H * W), 256, 0, stream>>>(x, w, bias, ln_w, ln_b, out, H, W, C, eps, get_operand_format());
  count_launch();
  return check_launch("dwconv7_ln_kernel");
}

extern "C" int vdn_mask_down1(const float* depth, const float* params, float* out, int32_t B, int32_t H, int32_t W, void* stream_v) {
  VDN_STREAM;
  if (!depth || !params || !out) return set_error("vdn_mask_down1: null pointer");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  mask_down1_kernel<<<blocks_for((long long)B * Ho * Wo, 256), 256, 0, stream>>>(depth, params, out, B, H, W, Ho, Wo);
  count_launch();
  return check_launch("mask_down1_kernel");
}

extern "C" int vdn_mask_down2(const float* in, const float* params, float* out, int32_t B, int32_t Hi, int32_t Wi, void* stream_v) {
  VDN_STREAM;
  if (!in || !params || !out) return set_error("vdn_mask_down2: null pointer");
  if (Hi < 7 || Wi < 7) return set_error("vdn_mask_down2: input smaller than the 7x7 kernel");
  const int Ho = (Hi - 7) / 7 + 1, Wo = (Wi - 7) / 7 + 1;
  mask_down2_kernel<<<blocks_for((long long)B * Ho * Wo * 32, 256), 256, 0, stream>>>(in, params, out, B, Hi, Wi, Ho, Wo);
  count_launch();
  return check_launch("mask_down2_kernel");
}

// Full-resolution, bandwidth-bound pieces of the v5 depth-refinement model (models/video_depth_model_v5.py:160-192):
//   vdn_frame_median_scale : GlobalQuantilePool2d(0.5) -> ZeroConv 1x1 -> exp(tanh(.))      (:63-87, :165-167)
//   vdn_v5_net_input       : scale, Sobel normals of the 224x224 resized depth, 3-channel network input (:169-178, normal_utils.py:4-52)
//   vdn_v5_residual        : bilinear to the input size, ReLU, shift_head affine, + scaled input, * max_depth (:183-192)
// The median is an exact radix select on the order-preserving integer image of the floats (4 passes of 8 bits per rank,
// block-wide shared-memory histograms), with torch.quantile's linear interpolation between the two middle ranks.
#include "../../include/vdn_b200.h"
#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// k-th smallest (0-based) of x[0..n) for the whole block; result broadcast to every thread.
__device__ float block_select(const float* __restrict__ x, long long n, long long k, uint32_t* hist, uint32_t* sh_prefix, long long* sh_k) {
  uint32_t prefix = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = float_key(x[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long kk = k;
      int b = 0;
      for (; b < 255; ++b) {
        if (kk < (long long)hist[b]) break;
        kk -= hist[b];
      }
      *sh_prefix = prefix | (uint32_t(b) << shift);
      *sh_k = kk;
    }
    __syncthreads();
    prefix = *sh_prefix;
    k = *sh_k;
    __syncthreads();
  }
  return key_float(prefix);
}

__global__ void __launch_bounds__(1024)
frame_median_scale_kernel(const float* __restrict__ x, float* __restrict__ med_out, float* __restrict__ scale_out, long long n, float inv_max, float w,
                          float b) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sh_prefix;
  __shared__ long long sh_k;
  const float* p = x + (long long)blockIdx.x * n;
  const long long lo = (n - 1) / 2, hi = n / 2;  // ranks floor / ceil of 0.5 * (n - 1)
  const float v_lo = block_select(p, n, lo, hist, &sh_prefix, &sh_k);
  float med = v_lo;
  if (hi != lo) {
    const float v_hi = block_select(p, n, hi, hist, &sh_prefix, &sh_k);
    med = v_lo + 0.5f * (v_hi - v_lo);  // torch.quantile: lerp(v_lo, v_hi, 0.5)
  }
  if (threadIdx.x == 0) {
    if (med_out != nullptr) med_out[blockIdx.x] = med;
    scale_out[blockIdx.x] = expf(tanhf(med * inv_max * w + b));
  }
}

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// r: [N, h, w] resized raw depth; x: [N, 3, h, w] = (r * scale / max, nx, ny) with normals of the scaled map
__global__ void __launch_bounds__(256)
v5_net_input_kernel(const float* __restrict__ r, const float* __restrict__ scale, float* __restrict__ x, int N, int h, int w, float inv_max) {
  const long long total = (long long)N * h * w;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int xx = int(idx % w);
    long long t = idx / w;
    const int yy = int(t % h);
    const int n = int(t / h);
    const float s = scale[n] * inv_max;
    const float* p = r + (long long)n * h * w;
    const int ym = reflect1(yy - 1, h), yp = reflect1(yy + 1, h), xm = reflect1(xx - 1, w), xp = reflect1(xx + 1, w);
    const float a = p[ym * w + xm] * s, b = p[ym * w + xx] * s, c = p[ym * w + xp] * s;
    const float d = p[yy * w + xm] * s, e = p[yy * w + xx] * s, f = p[yy * w + xp] * s;
    const float g = p[yp * w + xm] * s, hh = p[yp * w + xx] * s, i = p[yp * w + xp] * s;
    const float ix = ((a - c) + 2.0f * (d - f) + (g - i)) * 0.125f;
    const float iy = ((a + 2.0f * b + c) - (g + 2.0f * hh + i)) * 0.125f;
    const float inv = 1.0f / sqrtf(ix * ix + iy * iy + 1.0f + 1e-8f);
    float* o = x + (long long)n * 3 * h * w + (long long)yy * w + xx;
    o[0] = e;
    o[(long long)h * w] = -ix * inv;
    o[2LL * h * w] = -iy * inv;
  }
}

// out = (din / max * scale + relu(bilinear_ac(o -> H x W)) * ws + bs) * max
__global__ void __launch_bounds__(256)
v5_residual_kernel(const float* __restrict__ din, const float* __restrict__ o, const float* __restrict__ scale, float* __restrict__ out, int N, int H,
                   int W, int h, int w, float ws, float bs, float max_depth) {
  const long long total = (long long)N * H * W;
  const float sh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f;
  const float sw = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.0f;
  const float inv_max = 1.0f / max_depth;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int X = int(idx % W);
    long long t = idx / W;
    const int Y = int(t % H);
    const int n = int(t / H);
    const float sy = sh * (float)Y, sx = sw * (float)X;
    int y0 = (int)sy, x0 = (int)sx;
    if (y0 > h - 1) y0 = h - 1;
    if (x0 > w - 1) x0 = w - 1;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float* p = o + (long long)n * h * w;
    float v = (1.0f - ly) * ((1.0f - lx) * p[y0 * w + x0] + lx * p[y0 * w + x1]) + ly * ((1.0f - lx) * p[y1 * w + x0] + lx * p[y1 * w + x1]);
    v = fmaxf(v, 0.0f);
    const float d = din[idx] * inv_max * scale[n];
    out[idx] = (d + (v * ws + bs)) * max_depth;
  }
}

}  // namespace vdn

using namespace vdn;

extern "C" int vdn_frame_median_scale(const float* x, float* median, float* scale, int32_t N, int64_t n_per_frame, float inv_max, float w, float b,
                                      void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!x || !scale) return set_error("vdn_frame_median_scale: null pointer");
  if (N <= 0 || n_per_frame <= 0) return set_error("vdn_frame_median_scale: bad shape");
  frame_median_scale_kernel<<<N, 1024, 0, stream>>>(x, median, scale, n_per_frame, inv_max, w, b);
  count_launch();
  return check_launch("frame_median_scale_kernel");
}

extern "C" int vdn_v5_net_input(const float* r, const float* scale, float* x, int32_t N, int32_t h, int32_t w, float inv_max, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!r || !scale || !x) return set_error("vdn_v5_net_input: null pointer");
  if (h < 2 || w < 2) return set_error("vdn_v5_net_input: bad shape");
  const long long total = (long long)N * h * w;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
  v5_net_input_kernel<<<(unsigned)blocks, 256, 0, stream>>>(r, scale, x, N, h, w, inv_max);
  count_launch();
  return check_launch("v5_net_input_kernel");
}

extern "C" int vdn_v5_residual(const float* din, const float* o, const float* scale, float* out, int32_t N, int32_t H, int32_t W, int32_t h, int32_t w,
                               float ws, float bs, float max_depth, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!din || !o || !scale || !out) return set_error("vdn_v5_residual: null pointer");
  const long long total = (long long)N * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
  v5_residual_kernel<<<(unsigned)blocks, 256, 0, stream>>>(din, o, scale, out, N, H, W, h, w, ws, bs, max_depth);
  count_launch();
  return check_launch("v5_residual_kernel");
}

/* vdn_b200 — C ABI of the B200-native (sm_100a) depth/normal inference kernels.
 *
 * The reference (injun-baek/Video-Depth-Normal-v2) has no FFI layer: its "operator API" is the
 * nn.Module surface.  Each entry point below replaces the PyTorch/library op(s) the
 * reference calls at the cited lines; the Python host (video_depth_normal_v2_b200/) binds them with
 * ctypes and mirrors the reference's module API on top.  Plain pointers and sizes only: every pointer is
 * a CUDA device pointer unless noted, `stream` is a cudaStream_t passed as void*.
 *
 * All functions return 0 on success, non-zero on error (message via vdn_last_error()).  Kernels never
 * allocate; outputs/workspaces are caller-owned.  "16-bit" tensors are fp16 or bf16 according to
 * vdn_set_operand_format (library-wide; fp16 default).
 */
#ifndef VDN_B200_H_
#define VDN_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library state ------------------------------------------------------------------------- */
const char* vdn_last_error(void);
int vdn_version(void);
/* 0 = fp16 operands, 1 = bf16 operands (tcgen05 kind::f16 runs both at the same rate). */
int vdn_set_operand_format(int fmt);
int vdn_get_operand_format(void);
/* number of kernel launches issued by this library since the last reset (bench.py "gpu_launches"). */
int64_t vdn_launch_count(void);
void vdn_reset_launch_count(void);
/* account for kernels launched by replaying a CUDA graph that captured n of this library's launches */
void vdn_add_launch_count(int64_t n);

/* ---- tcgen05 GEMM / implicit-GEMM convolution with fused epilogue ---------------------------
 * out = epilogue( A[M,K] * W[N,K]^T ), fp32 accumulation in TMEM.
 * Replaces: nn.Linear qkv/proj/fc1/fc2 (dinov2_layers/attention.py:52,60; mlp.py:36-40) incl. LayerScale
 * (layer_scale.py:27-28) and the residual adds (block.py:105-106); PatchEmbed conv (patch_embed.py:76);
 * DPT 1x1 / 3x3 / ConvTranspose convs (dpt.py:60-124, util/blocks.py:20-32,78-91,159); motion-module linears,
 * GEGLU (motion_module/attention.py:382-384) and proj_out + residual (motion_module.py:128-133).            */
enum { VDN_ACT_NONE = 0, VDN_ACT_GELU = 1, VDN_ACT_RELU = 2, VDN_ACT_SILU = 3 /* gate activation of the GLU epilogue only (SwiGLU) */ };
enum {
  VDN_ROWMAP_IDENTITY = 0,
  VDN_ROWMAP_PIXEL_SHUFFLE = 1, /* ConvTranspose k==stride: rm0=H_in rm1=W_in rm2=stride rm3=C_out; N = stride^2*C_out, col=(i*s+j)*C_out+co */
  VDN_ROWMAP_TEMPORAL = 2,      /* rows (b*D+d)*T+f -> (b*T+f)*D+d : rm0=T rm1=D (applies to out, res, res2, out2) */
  VDN_ROWMAP_PATCH_TOKENS = 3,  /* rows b*P+p -> b*(P+1)+1+p, residual row 1+p (pos_embed) : rm0=P */
  VDN_ROWMAP_QKV_SPLIT = 4      /* cols [0,split) -> out[row', col]; cols [split,N) -> out2 = V^T[(b*heads+h)*64+d][t'] : rm0=tokens per batch of the
                                   rows, rm1=ld of V^T, rm2=C; split = qkv_split (0 -> 2C).  row' = b*qkv_tokens_out + t', t' = qkv_token_offset + t
                                   (qkv_tokens_out 0 -> rm0): lets a [k|v] projection of new memory tokens land in slot s of a per-layer KV cache */
};

typedef struct vdn_gemm_desc {
  const void* a;   /* plain: [M, K] 16-bit row-major, row stride lda elements (lda*2 % 16 == 0).
                      conv : NHWC [B, H, W, K] 16-bit, dense */
  const void* w;   /* [N, Kw] 16-bit, K contiguous, row stride ldw.  conv: Kw = 9 * roundup(K,64), tap-major (r*3+s) */
  int64_t M, N, K; /* conv: M = B*H*W, K = C_in */
  int64_t lda, ldw;
  int32_t conv;    /* 0 = plain GEMM, 1 = 3x3 stride-1 pad-1 convolution (implicit GEMM, 9 shifted TMA boxes) */
  int32_t B, H, W; /* conv geometry */
  const float* bias;  /* [N] or NULL */
  const float* gamma; /* [N] or NULL: out = res + gamma * (acc + bias)  (LayerScale) */
  const void* res;    /* residual or NULL */
  int32_t res_f32;    /* residual is fp32 (else 16-bit) */
  int64_t ld_res;
  const void* res2;   /* second residual, 16-bit, or NULL */
  int64_t ld_res2;
  void* out;
  int32_t out_f32;    /* output is fp32 (else 16-bit) */
  int64_t ldc;
  void* out2;         /* optional second 16-bit output: copy (or ReLU copy) of out; V^T for QKV_SPLIT */
  int32_t out2_relu;
  int64_t ld_out2;
  int32_t act;        /* VDN_ACT_* applied to (acc + bias) */
  int32_t geglu;      /* weights interleaved (value_j, gate_j): out[:, j] = value_j * gelu(gate_j); out has N/2 columns */
  int32_t row_map;
  int32_t rm0, rm1, rm2, rm3;
  const float* head_w; /* if non-NULL (N <= 32): out_f32[row] = relu(sum_j relu(acc_j+bias_j) * head_w[j] + head_b) */
  float head_b;
  int32_t qkv_split, qkv_tokens_out, qkv_token_offset; /* VDN_ROWMAP_QKV_SPLIT extensions (0 = defaults) */
} vdn_gemm_desc;

int vdn_gemm(const vdn_gemm_desc* d, void* stream);

/* ---- spatial (ViT) flash attention, tcgen05 ---------------------------------------------------
 * softmax(q k^T / sqrt(64)) v per (frame, head), head_dim 64.  Replaces attention.py:53-58 / the xformers
 * memory_efficient_attention call at attention.py:76.
 * qk : [B*tokens, ld_qk] 16-bit, q at cols [0,C), k at cols [C,2C) (head h = cols h*64..h*64+63)
 * vT : [B*heads, 64, ld_vT] 16-bit (written by vdn_gemm with VDN_ROWMAP_QKV_SPLIT)
 * out: [B*tokens, C] 16-bit                                                                              */
int vdn_flash_attn(const void* qk, int64_t ld_qk, const void* vT, int64_t ld_vT, void* out, int32_t B, int32_t tokens,
                   int32_t heads, void* stream);
/* general form: separate q [B, tokens_q, ld_q] and k [B, tokens_kv, ld_k] buffers (element strides), V^T [B*heads, 64, ld_vT];
   cross-attention to a memory bank (sam2/modeling/sam/transformer.py:275-311 after the projections and RoPE) */
int vdn_flash_attn_ex(const void* q, int64_t ld_q, int64_t q_batch_stride, const void* k, int64_t ld_k, int64_t k_batch_stride, const void* vT,
                      int64_t ld_vT, void* out, int32_t B, int32_t tokens_q, int32_t tokens_kv, int32_t heads, void* stream);

/* ---- temporal attention over T frames per (pixel, head) ----------------------------------------
 * Replaces CrossAttention._attention (motion_module/attention.py:182-211) as called from
 * TemporalAttention.forward (motion_module.py:296-313).  qkv: [D*T, 3C] rows (d, f), q|k|v column blocks.
 * out: [D*T, C].  head_dim = C/heads (any multiple of 8).                                              */
int vdn_temporal_attn(const void* qkv, void* out, int32_t D, int32_t T, int32_t C, int32_t heads, void* stream);
/* tcgen05 form for T == 32 frames, head_dim 32 / 64 / 128: qk [rows, ld_qk] = q | k (rows = (pixel, frame), 32 consecutive rows per
   pixel), vT [ceil(rows/128), C, 128] = V transposed per 128-row tile (vdn_gemm with VDN_ROWMAP_QKV_SPLIT, rm0 = rm1 = 128; must be
   zero beyond the last valid row), out [rows, C] */
int vdn_temporal_attn_tc(const void* qk, int64_t ld_qk, const void* vT, void* out, int64_t rows, int32_t C, int32_t heads, void* stream);
/* streaming form (video_depth_stream.py:76-160, motion_module.py:252-269): one new frame (entry L-1) attends to L <= 32 cached frames.
   qkv_entries[j] = the cached bias-free projection [D, ld] = (Wq n_j | Wk n_j | Wv n_j) of frame j's normed hidden state; pos [32, 3C] fp32 =
   the positional part (Wq pe_j | Wk pe_j | Wv pe_j); out [D, C] = softmax((q+pq)(k+pk)^T / sqrt(dh)) (v+pv) per (pixel, head) */
int vdn_stream_temporal_attn(const void* const* qkv_entries, int32_t L, int64_t ld, const float* pos, void* out, int32_t D, int32_t C, int32_t heads,
                             void* stream);
/* the same attention with the cached projections in a slot pool [slots, D, ld] and a DEVICE-side slot table (int32, entry j = slot of
   cached frame j, < 0 = `staging`, this frame's projection): launch parameters are frame-independent, so the streaming step replays
   as one CUDA graph.  vdn_ring_store copies `staging` into slot table[slot_index] of the pool (the cache insertion, also in the graph). */
int vdn_stream_temporal_attn_ring(const void* pool, const void* staging, const int32_t* slot_table, int32_t L, int64_t ld, const float* pos,
                                  void* out, int32_t D, int32_t C, int32_t heads, void* stream);
int vdn_ring_store(const void* staging, void* pool, const int32_t* slot_table, int32_t slot_index, int64_t slot_elems, void* stream);

/* ---- normalisation ----------------------------------------------------------------------------- */
/* LayerNorm over C of fp32 rows -> 16-bit.  out row = map(row):
 *   drop_first == 0: same row.  drop_first == 1 (ViT taps, dinov2.py:309-316): rows_per_batch rows per batch,
 *   row 0 (cls) is dropped, row j>0 of batch b goes to b*(rows_per_batch-1)+j-1.
 * pe (optional, fp32 [pe_len, C]): added after the affine, row r uses pe[r % pe_len] (motion_module.py:211). */
int vdn_layernorm(const float* x, const float* w, const float* b, void* out, int64_t rows, int32_t C, float eps, int32_t drop_first,
                  int32_t rows_per_batch, const float* pe, int32_t pe_len, void* stream);
/* GroupNorm(32) statistics of NHWC 16-bit frames: stats[f*32+g] = (mean, rstd).  motion_module.py:84,111 */
int vdn_groupnorm_stats(const void* x, float* stats, int32_t frames, int32_t D, int32_t C, int32_t groups, float eps, void* stream);
/* apply GroupNorm + affine and transpose frame-major [T, D, C] -> pixel-major rows (b*D+d)*T+f, 16-bit.  motion_module.py:111-114,253 */
int vdn_groupnorm_apply_tc(const void* x, const float* stats, const float* w, const float* b, void* out, int32_t Bv, int32_t T, int32_t D,
                           int32_t C, int32_t groups, void* stream);

/* both of the above in one call: statistics (written to stats[frames*groups*2] as well) + affine + transpose.  One cluster launch
   (8 CTAs per frame, distributed-shared-memory merge of the statistics) for C in {256, 512, 1024, 2048} with 8 or more frames;
   every other shape runs the two kernels above.  motion_module.py:103-115 */
int vdn_groupnorm_to_tc(const void* x, const float* w, const float* b, void* out, float* stats, int32_t Bv, int32_t T, int32_t D, int32_t C,
                        int32_t groups, float eps, void* stream);

/* ---- layout / elementwise ------------------------------------------------------------------------ */
/* fp32 NCHW image [B,3,H,W] -> 16-bit patch rows [B*ph*pw, Kp] (Kp >= 588, zero padded), col = c*196 + i*14 + j. patch_embed.py:76 */
int vdn_patch_im2col(const float* img, void* out, int32_t B, int32_t H, int32_t W, int32_t Kp, void* stream);
/* x[b, 0, :] = cls + pos[0]  for every frame (dinov2.py:219-220) */
int vdn_write_cls(float* x, const float* cls, const float* pos, int32_t B, int32_t tokens, int32_t C, void* stream);
/* use_clstoken readout input (dpt.py:129-132, dpt_temporal.py:56-59): out [frames*P, 2C] = [patch token | that frame's cls token], the
   operand of readout_projects[i] (Linear 2C->C + GELU).  Token p of frame f is row f*tok_frame_pitch + p of `tok`, the cls token of
   frame f row f*cls_frame_pitch of `cls` (rows of C 16-bit elements).  Video models: both point into the normed token matrix
   [frames*(P+1), C] (tok = row 1, pitches P+1); DepthAnythingV2's last tap takes its tokens from the memory block's output. */
int vdn_readout_concat(const void* tok, int64_t tok_frame_pitch, const void* cls, int64_t cls_frame_pitch, void* out, int64_t frames, int32_t P,
                       int32_t C, void* stream);
/* im2col for the 3x3 stride-2 pad-1 conv (dpt.py:84-89): NHWC [B,H,W,C] -> [B*Ho*Wo, 9*C] */
int vdn_im2col_3x3_s2(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* Head tail in one kernel (dpt_temporal.py:103-111, dpt.py:179-187): [bilinear resize to (H, W), align_corners=True ->]
 * 3x3 conv 128 -> 32 (pad 1) + bias -> ReLU -> 1x1 conv 32 -> 1 + bias -> ReLU, one fp32 value per pixel.
 * wpacked: the 3x3 filter as pre-swizzled tensor-core B tiles (packing.py::pack_conv_tail), 73728 bytes.
 * vdn_conv_tail reads an already resized 16-bit NHWC map x [B, H, W, 128]; vdn_conv_tail_up resizes src [B, Hs, Ws, 128] on the fly
 * (the full-resolution 128-channel map is never written). */
int vdn_conv_tail(const void* x, const void* wpacked, const float* bias, const float* head_w, float head_b, float* out, int32_t B, int32_t H,
                  int32_t W, void* stream);
int vdn_conv_tail_up(const void* src, int32_t Hs, int32_t Ws, const void* wpacked, const float* bias, const float* head_w, float head_b, float* out,
                     int32_t B, int32_t H, int32_t W, void* stream);
/* bilinear resize, align_corners=True, NHWC 16-bit (F.interpolate at util/blocks.py:155-157, dpt_temporal.py:104-106).
 * optional second input added before interpolation is NOT supported; relu_out writes max(x,0). */
int vdn_bilinear_nhwc(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t Ho, int32_t Wo, int32_t C, int32_t relu_out,
                      void* stream);
/* same, additionally writing relu(out) to out_relu (the ReLU'd copy the next ResidualConvUnit consumes, util/blocks.py:77) */
int vdn_bilinear_nhwc2(const void* x, void* out, void* out_relu, int32_t B, int32_t H, int32_t W, int32_t Ho, int32_t Wo, int32_t C, void* stream);
/* bilinear resize of fp32 planes [N,H,W] -> [N,Ho,Wo], align_corners=True, optional ReLU (video_depth.py:63-64,110) */
int vdn_bilinear_f32(const float* x, float* out, int32_t N, int32_t H, int32_t W, int32_t Ho, int32_t Wo, int32_t relu, void* stream);
/* out = relu(x), 16-bit, n elements (util/blocks.py:78) */
int vdn_relu16(const void* x, void* out, int64_t n, void* stream);
/* out16 = x (fp32 -> 16-bit), n elements */
int vdn_cast_f32_to_16(const float* x, void* out, int64_t n, void* stream);

/* ---- frame pre-processing (util/transform.py Resize(INTER_CUBIC) + NormalizeImage + PrepareForNet, video_depth.py:74-99) ----
 * frames uint8 [N, H, W, 3] RGB (device) -> out fp32 [N, 3, h, w] = ((cubic_resize(frames / 255) - mean) / std); mean3 / std3 are HOST pointers */
int vdn_preprocess_u8(const void* frames, float* out, int32_t N, int32_t H, int32_t W, int32_t h, int32_t w, const float* mean3, const float* std3,
                      void* stream);

/* ---- window alignment (video_depth.py:118-154, utils/util.py:40-74) ------------------------------- */
/* sums[0..4] = (sum p*p, sum p, n, sum p*t, sum t) over n elements, accumulated in fp64 on device */
int vdn_lsq_sums(const float* pred, const float* target, int64_t n, double* sums5, void* stream);
/* scale_shift[0..1] = solution of the 2x2 normal equations built from sums5 (identity (1, 0) when singular; utils/util.py:54-60) */
int vdn_lsq_solve(const double* sums5, float* scale_shift, void* stream);
/* out = max(x*scale + shift, 0) */
int vdn_affine_clamp(const float* x, float* out, int64_t n, const float* scale_shift, void* stream);
/* out = pre*(1-w) + max(post*scale+shift,0)*w */
int vdn_crossfade(const float* pre, const float* post, float* out, int64_t n, const float* scale_shift, float w, void* stream);

/* One launch for every output frame a window owns (video_depth.py:131-152 + utils/util.py:65-74): cur = raw depth [32, n] of window k,
 * prev_tail = raw depth [8, n] of window k-1's slots 24..31, ss_cur / ss_prev = device (scale, shift) of windows k / k-1 (ss_prev NULL:
 * the predecessor is window 0, which is never re-scaled).  out [count, n] = slots first_slot .. first_slot+count-1:
 * slot < 10 -> pre*(1-w_j) + max(cur*s+t, 0)*w_j with pre = max(prev_tail[j]*s'+t', 0), j = slot-2, w = (0, 1/7, .., 6/7, 1);
 * slot >= 10 -> max(cur*s+t, 0).  is_first (window 0): plain copy. */
int vdn_window_finalize(const float* cur, const float* prev_tail, const float* ss_cur, const float* ss_prev, float* out, int64_t n,
                        int32_t first_slot, int32_t count, int32_t is_first, void* stream);
/* keys [3, n] = slots (0, 1, 12) of cur [32, n]: what the sequential scale/shift chain needs from a window (video_depth.py:121-129,148-152) */
int vdn_window_keys(const float* cur, float* keys, int64_t n, void* stream);

/* ---- depth -> normals (utils/normal_utils.py:4-52): reflect-pad Sobel/8, n = normalize(-Ix,-Iy,1) ------- */
int vdn_sobel_normals(const float* depth, float* normals, int32_t N, int32_t H, int32_t W, int32_t channels_out, void* stream);

/* ---- v5 depth-refinement model, full-resolution pieces (models/video_depth_model_v5.py:160-192) -------------- */
/* per frame: median (torch.quantile 0.5 semantics; optional output) and scale = exp(tanh(median * inv_max * w + b))
   = GlobalScaleHead (:63-87) of the frame divided by max_depth */
int vdn_frame_median_scale(const float* x, float* median, float* scale, int32_t N, int64_t n_per_frame, float inv_max, float w, float b, void* stream);
/* r [N,h,w] (raw depth resized) -> x [N,3,h,w] = (r*scale/max, normal_x, normal_y) (:169-178, utils/normal_utils.py:4-52) */
int vdn_v5_net_input(const float* r, const float* scale, float* x, int32_t N, int32_t h, int32_t w, float inv_max, void* stream);
/* out = (din/max*scale + relu(bilinear(o [N,h,w] -> [N,H,W])) * ws + bs) * max   (:183-192) */
int vdn_v5_residual(const float* din, const float* o, const float* scale, float* out, int32_t N, int32_t H, int32_t W, int32_t h, int32_t w, float ws,
                    float bs, float max_depth, void* stream);

/* ---- DepthAnythingV2 memory block, bandwidth-bound pieces (depth_anything_v2/memory_block.py, sam2/modeling/*) ---- */
/* axial RoPE in place on `heads` 64-wide heads starting at column col0 of 16-bit rows [rows, ld]; row r sits at grid position r % P;
   cos_sin [P, 64] fp32 = cos[32] | sin[32] per position (sam2/modeling/position_encoding.py:186-239) */
int vdn_rope2d(void* x, int64_t rows, int64_t ld, int32_t col0, int32_t heads, const float* cos_sin, int32_t P, int64_t rows_per_batch,
               int64_t batch_pitch, void* stream); /* rows_per_batch consecutive rows per batch, batches batch_pitch rows apart (0 -> dense) */
/* temporal RoPE (pe='rope', motion_module.py:236-240,279-282; attention.py:403-429): rotate `chunks` 64-wide column chunks starting at
   col0 of 16-bit rows [rows, ld] in place; row r is frame r % P; cos_sin [P, chunks, 64] fp32 = cos[32] | sin[32] per (frame, chunk) —
   the reference's frequencies run over the whole query_dim, so every chunk has its own table */
int vdn_rope_chunks(void* x, int64_t rows, int64_t ld, int32_t col0, int32_t chunks, const float* cos_sin, int32_t P, void* stream);
/* out_f32[r, c] = x[r, c] + alpha * vec[c]   (x fp32 or 16-bit; in place allowed for fp32) */
int vdn_add_rowvec(const void* x, int32_t x_f32, const float* vec, float alpha, float* out, int64_t rows, int32_t C, void* stream);
/* x[r, :] += m[r] */
int vdn_add_rowscalar(float* x, const float* m, int64_t rows, int32_t C, void* stream);
/* ConvNeXt front half: depthwise 7x7 (pad 3, w [49, C] tap-major) + LayerNorm2d over channels; NHWC fp32 -> 16-bit [B*H*W, C]
   (sam2/modeling/memory_encoder.py:96-99) */
int vdn_dwconv7_ln(const float* x, const float* w, const float* bias, const float* ln_w, const float* ln_b, void* out, int32_t B, int32_t H, int32_t W,
                   int32_t C, float eps, void* stream);
/* MaskDownSampler stages on the 1-channel map (memory_encoder.py:17-60 as configured at memory_block.py:68-71):
   1: sigmoid -> conv3x3/s2/p1 (1->4) -> LN2d -> GELU -> 1x1; params = w0[4][9] b0[4] lnw[4] lnb[4] w1[4] b1      out [B, ceil(H/2), ceil(W/2)]
   2: conv7x7/s7 (1->49) -> LN2d -> GELU -> 1x1;             params = w0[49][49] b0[49] lnw[49] lnb[49] w1[49] b1 out [B, Hi/7, Wi/7] */
int vdn_mask_down1(const float* depth, const float* params, float* out, int32_t B, int32_t H, int32_t W, void* stream);
int vdn_mask_down2(const float* in, const float* params, float* out, int32_t B, int32_t Hi, int32_t Wi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VDN_B200_H_ */

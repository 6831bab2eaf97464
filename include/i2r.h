/*
 * i2r.h -- C ABI of libi2r_sm100.so: the B200 (sm_100a) kernels behind the I2R-Net forward path.
 *
 * Boundary: the reference has no FFI of its own -- its forward is a chain of torch
 * ops inside `model(x, pos_mask, length)` (lib/core/function.py:135).  Each entry point below
 * replaces the torch/cuDNN/cuBLAS call sequence of one group of reference lines; the Python module
 * that mirrors the reference's `lib/models/*.py` surface (get_pose_net / forward) loads this
 * library with ctypes and is the only caller.
 *
 * Conventions
 *   - every function returns 0 on success, a negative I2R_E_* code for argument errors, or the
 *     positive cudaError_t of a failed launch; i2r_last_error() gives a thread-local message;
 *   - nothing here allocates device memory or synchronises; all pointers are device pointers owned
 *     by the caller (torch), `stream` is a cudaStream_t passed as void*;
 *   - activations are NHWC fp16 ("pixel-major": one row of C channels per pixel); accumulation is
 *     fp32 in TMEM; weights are fp16 in the packed K-major core-matrix layout documented at
 *     i2r_conv_problem::w.
 */
#ifndef I2R_H_
#define I2R_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2R_ABI_VERSION 4

#define I2R_E_BADARG (-1)
#define I2R_E_UNSUPPORTED (-2)
#define I2R_E_DEVICE (-3)

#define I2R_MAX_TAPS 9
#define I2R_MAX_GROUP 6
#define I2R_MAX_CHAIN_PROBLEMS 40 /* problems of one chained halo launch (all layers together) */
#define I2R_MAX_CHAIN_LAYERS 16

/* i2r_conv_problem::flags */
#define I2R_F_RELU 1u        /* clamp at 0 after scale/bias/addends                         */
#define I2R_F_OUT_NCHW_F32 2u /* write fp32 NCHW (heatmap head) instead of fp16 NHWC         */
#define I2R_F_OUT_F32 4u      /* write fp32 NHWC (row-major [pixels, Cout])                  */
#define I2R_F_OUT_T16 16u     /* i2r_conv_halo only: write fp16 TRANSPOSED, y[c * out_pix_stride + p] (channel-major rows of
                              * out_pix_stride pixels; with I2R_F_SPLIT the lo rows follow the Cout hi rows): the V^T operand of
                              * i2r_attention_tc */
#define I2R_F_GELU 32u        /* erf-GELU instead of ReLU (MlpDWBN of HRFormer-B, lib/models/hrformer.py:1094-1119)            */
#define I2R_F_ACT_FIRST 64u   /* y = act(scale * acc + bias) + addends  (instead of act(... + addends)): x + mlp(x), :1236      */
/* Split-operand mode (the 1e-3 heatmap bar of the TransPose-H families needs ~22-bit operands): activations are
 * fp16 PAIRS -- a pixel holds 2*C channels, [0,C) = hi, [C,2C) = lo, value = hi + lo -- and a product is
 * x_hi*W_hi + x_lo*W_hi + x_hi*W_lo with fp32 accumulation, evaluated as ONE GEMM over K = [x_hi | x_lo | x_hi]
 * against weights packed as [W_hi | W_hi | W_lo] (3*ceil(Cin/64) K-chunks per tap).  x, add0/add1 and an fp16
 * NHWC output are all pair tensors (addend pointers address the hi half; lo is Cout channels further); Cin / Cout
 * stay the logical channel counts; the pixel strides are the physical ones (>= 2*C). */
#define I2R_F_SPLIT 8u

/*
 * One implicit-GEMM problem:  Y[p, n] = act( scale[n] * sum_{t,c} X[src(p,t), c] * W[t, c, n]
 *                                           + bias[n] + add0[p >> s0, n] + add1[p >> s1, n] )
 * with p an output pixel (b, oy, ox), taps t with integer offsets, and c the input channel.
 * Covers (reference lines): nn.Conv2d 3x3/1x1 stride 1/2 + eval BatchNorm2d + ReLU + residual
 * (lib/models/interformer_pureMulti.py:37-107, :543-582), HRNet fuse layers incl. nearest upsample
 * (:353-387, :392-410), nn.Linear (attention/FFN projections, :174-213), ConvTranspose2d 4x4 s2
 * as four 2x2-tap phases (:648-672), and the final 1x1 head (:489-495).
 */
typedef struct i2r_conv_problem {
  const void* x;     /* fp16 NHWC source [NB, IH>>in_shift, IW>>in_shift, *] ; pixel stride in_pix_stride  */
  const void* w;     /* fp16 packed [ntaps][Cin/KC][KC/8][Npad][8]  (B operand, K-major core matrices)      */
  const float* scale; /* [Npad] per-output-channel scale (folded BatchNorm gamma/sqrt(var+eps)), or 1       */
  const float* bias;  /* [Npad] per-output-channel bias (folded BN beta - mean*scale, or conv/linear bias)  */
  const void* add0;  /* optional fp16 NHWC addend at output resolution >> add0_shift, Cout channels        */
  const void* add1;  /* optional second addend                                                              */
  void* y;           /* output, see flags                                                                   */
  int32_t NB, IH, IW;      /* logical input extent (after the nearest-upsample by 2^in_shift)              */
  int32_t Cin, KC;         /* input channels, K-chunk per pipeline stage (48 or 64; divides Cin)            */
  int32_t in_pix_stride;   /* elements between consecutive source pixels (>= Cin)                          */
  int32_t in_shift;        /* source pixel = (iy >> in_shift, ix >> in_shift)                              */
  int32_t OH, OW;          /* GEMM-M index space: M = NB*OH*OW, iy = oy*stride + dy[t]                      */
  int32_t stride;
  int32_t Cout, Npad;      /* real / padded (multiple of 16, <= 256) output channels                        */
  int32_t out_pix_stride;  /* elements between consecutive output pixels (NHWC modes)                       */
  int32_t OHf, OWf;        /* full output extent; output pixel = (oy*out_mul+out_offy, ox*out_mul+out_offx) */
  int32_t out_mul, out_offy, out_offx;
  int32_t add0_shift, add1_shift;
  int32_t add_pix_stride;  /* elements between consecutive addend pixels (>= Cout; = Cout when dense)      */
  int32_t ntaps;
  int8_t dy[I2R_MAX_TAPS + 3];
  int8_t dx[I2R_MAX_TAPS + 3];
  uint32_t flags;
  const void* w_folded;    /* i2r_conv_halo only: fp16 [1 + ntaps*ceil(Cin/64)][Npad][64] -- block 0 is the bias  */
                           /* block (K slot 0 = fp16(bias), slot 1 = fp16(bias - slot 0)), blocks 1.. are `w`     */
                           /* with scale[n] folded into every row n before the fp16 rounding                     */
  int32_t w_folded_copies; /* >= 1: that image repeated back to back; CTA i streams copy i % copies, which spreads */
                           /* the L2 lines every CTA reads at the same time over more L2 slices                    */
  int32_t pair_lo_offset;  /* I2R_F_SPLIT: channel offset from the hi half to the lo half inside y / add0 / add1   */
                           /* rows; 0 = Cout.  Lets a problem produce a SLICE of the output channels of a wider   */
                           /* pair tensor (layers with more than 256 output channels run as several problems)     */
} i2r_conv_problem;

int i2r_version(void);
const char* i2r_last_error(void);
/* 0 iff device `dev` is compute capability 10.x (sm_100a code is loadable). */
int i2r_device_check(int dev);
int i2r_sm_count(int dev);

/* Launch `nprob` (1..I2R_MAX_GROUP) independent problems as ONE grid of 128-pixel tiles on the
 * tcgen05 path.  impl: 0 = tcgen05/TMEM kernel (product path), 1 = scalar SIMT check kernel
 * (tests only; same problem struct, same packed weights). */
int i2r_conv_igemm(const i2r_conv_problem* probs, int nprob, int impl, void* stream);

/* Persistent halo-tile variant for the problems that dominate the FLOPs: 3x3 convolutions of stride 1 (BasicBlock /
 * Bottleneck convs, interformer_pureMulti.py:37-107) and stride 2 (second stem conv :680-682, transitions :543-582,
 * fuse chains :353-387) and 1x1 / nn.Linear problems with no input resampling (in_shift 0, out_mul 1).  add0 is read at
 * output resolution; add1 may be a tensor of (OH >> add1_shift) x (OW >> add1_shift) pixels that is up-sampled (nearest)
 * in the epilogue (HRNet fuse, :392-410).  Stride 1: each 8x16-pixel tile's activation halo is staged once in shared
 * memory and the nine taps are shifted UMMA descriptor windows of it.  Stride 2: the (2*8+1) x (2*16+1) input pixels of
 * a tile are staged as four parity planes through TMA boxes with element strides 2, and every tap is a shifted window of
 * one plane.  Weights arrive as TMA bulk copies and stay resident in shared memory when they fit; two TMEM accumulators
 * overlap the epilogue with the next tile; fp16 NHWC outputs leave through shared-memory staging and TMA stores.
 * Split-operand problems (I2R_F_SPLIT) stage x_hi / x_lo and W_hi / W_lo once per real 64-channel chunk where that is the
 * faster scheme.  i2r_conv_halo_supported() returns 1 when a problem qualifies; others go through i2r_conv_igemm. */
int i2r_conv_halo_supported(const i2r_conv_problem* prob);
int i2r_conv_halo(const i2r_conv_problem* probs, int nprob, void* stream);

/* CHAINED launch: `nlayers` groups of problems (layer l = probs[sum(layer_count[:l]) ..][layer_count[l]], each group what
 * one i2r_conv_halo call would take) that would otherwise be `nlayers` back-to-back launches -- the conv1 / conv2 (+
 * residual) sequence of the BasicBlocks of an HRNet module (interformer_pureMulti.py:37-66, :284-330), the 1x1 / 3x3 / 1x1
 * sequence of the layer1 Bottlenecks (:69-107) -- run as ONE persistent grid.  Every CTA walks the layers in order; a tile
 * of a later layer starts as soon as the tiles it reads (the same IMAGE of the producing problems, found by address-range
 * overlap of inputs / addends with earlier outputs) have been stored, which the producers publish through per-image
 * pixel counters in `workspace` (device memory, >= i2r_conv_halo_chain_workspace(...) bytes, zeroed by this call on
 * `stream`).  No grid-wide barrier, no launch gap, one prologue per chain.
 * Requirements beyond i2r_conv_halo's: no CTA-pair problems, fp16 NHWC outputs, a
 * layer's outputs must not alias anything an earlier layer of the chain reads or writes, and at most ONE chained launch
 * may be in flight on a device at a time (CTAs of the grid wait for each other: issue chains on one stream).
 * Returns I2R_E_UNSUPPORTED (nothing launched) when the chain does not qualify -- launch the layers one by one then. */
int i2r_conv_halo_chain(const i2r_conv_problem* probs, const int* layer_count, int nlayers, void* workspace,
                        size_t workspace_bytes, void* stream);
size_t i2r_conv_halo_chain_workspace(const i2r_conv_problem* probs, int nprob);

/* Stem / mask convolution on fp32 NCHW input with tiny Cin (3 or 1): 3x3 stride 2 pad 1 + folded
 * BN + ReLU -> fp16 NHWC [NB, H/2, W/2, Cout] (split != 0: pair tensor [.., 2*Cout], see I2R_F_SPLIT).  Replaces conv1/bn1/relu
 * (interformer_pureMulti.py:677-679) and position_embedding.conv1/bn1/relu
 * (position_embedding.py:108-110).  w: fp32 [Cin*9][Cout] (k = (c*3+ky)*3+kx). */
int i2r_stem_conv3x3s2(const float* x, const float* w, const float* scale, const float* bias, void* y,
                       int NB, int Cin, int H, int W, int Cout, int split, void* stream);

/* The same layer on the tensor cores (the product path; the SIMT entry point above is the check implementation):
 * a 128-pixel tile is one implicit GEMM [128 x 32] x [32 x 64]; input and weights travel as fp16 (hi, lo) pairs so
 * the result keeps fp32-level accuracy.  wimg: i2r_stem_tc_weight_bytes() bytes from i2r_b200/packing.py
 * pack_stem_tc (BatchNorm scale folded into the weights); bias: fp32 [64]. */
int64_t i2r_stem_tc_weight_bytes(void);
int i2r_stem_conv3x3s2_tc(const float* x, const void* wimg, const float* bias, void* y, int NB, int Cin, int H, int W,
                          int Cout, int split, void* stream);

/* MaxPool2d(kernel 3, stride 2, padding 1) on fp16 NHWC (position_embedding.py:9,:113-114;
 * interformer.py:260-264). */
int i2r_maxpool3x3s2(const void* x, void* y, int NB, int H, int W, int C, int split, void* stream);

/* Single-head scaled-dot-product attention over ragged sequences (one per image):
 *   out[t, :] = softmax_j( scale * q[t,:] . k[j,:] ) v[j,:]   for j in the same sequence.
 * q,k,v,out: fp16 row-major with row strides ldq/ldk/ldv/ldo elements, head dim D (80 or 96);
 * cu_seqlens: int32 [nseq+1] token offsets on the device; total_tokens = cu_seqlens[nseq].
 * When few (sequence, query-tile) pairs exist the keys are split across CTAs and merged by a second
 * kernel; `workspace` (i2r_attention_workspace_bytes(), may be NULL/0 = no splitting) holds the fp32 partials.
 * Equivalent to nn.MultiheadAttention(nhead=1) with key_padding_mask on padded persons
 * (interformer_pureMulti.py:199-204; torch F.multi_head_attention_forward); with one crop per sequence it is the
 * intra-human attention of TransPose-H (transpose_h.py:165-240).  split != 0: split-operand mode (I2R_F_SPLIT) --
 * the lo half of each q/k/v/out row starts q_lo/k_lo/v_lo/o_lo elements after its hi half. */
int64_t i2r_attention_workspace_bytes(int total_tokens, int D, int nseq, int max_seqlen);
int i2r_attention_varlen(const void* q, const void* k, const void* v, void* out, int ldq, int ldk, int ldv,
                         int ldo, int D, const int32_t* cu_seqlens, int nseq, int max_seqlen, int total_tokens,
                         float scale, void* workspace, int64_t workspace_bytes, int split, int q_lo, int k_lo,
                         int v_lo, int o_lo, void* stream);

/* The same attention on the tcgen05 tensor cores (TMA-staged SWIZZLE_128B tiles, S/P/O in TMEM, one thread per
 * query row for the streaming softmax); the product path of both encoders.  Differences in the operand contract:
 *   - q, k: rows of D fp16 (split: 2*D -- the lo half DIRECTLY after the hi half), 16-byte aligned, strides % 8;
 *   - vt:   V TRANSPOSED, fp16 [D (split: 2*D, lo rows after the hi rows)][ldvt tokens] -- what i2r_conv_halo writes
 *           with I2R_F_OUT_T16; ldvt % 8 == 0;
 *   - out:  [T, D] (split: hi at column 0, lo at column o_lo).
 * Replaces F.multi_head_attention_forward's baddbmm / softmax / bmm (torch/nn/functional.py:6638-6650) as called
 * from transpose_h.py:165-240 and attention.py:68-73.  D = 96 or 80. */
int64_t i2r_attention_tc_workspace_bytes(int total_tokens, int D, int nseq, int max_seqlen);
int i2r_attention_tc(const void* q, const void* k, const void* vt, void* out, int ldq, int ldk, int ldvt, int ldo,
                     int D, const int32_t* cu_seqlens, int nseq, int max_seqlen, int total_tokens, float scale,
                     void* workspace, int64_t workspace_bytes, int split, int o_lo, void* stream);

/* Fused tail of one post-norm encoder layer (everything after the attention) for d_model 96 / dim_feedforward 192:
 *   x1 = attn W_o^T + b_o + src;  s1 = LN1(x1);  x2 = relu(s1 W_1^T + b_1) W_2^T + b_2 + s1;  out = LN2(x2);
 *   out_pos = out + pos (optional: the q/k input of the next layer).
 * attn / src / pos / out / out_pos: fp16 [T, 96] (split: [T, 192] pairs, lo directly after hi) with row strides
 * ld_attn / ld_src / ld (pos, out and out_pos share ld).  wimg / params: i2r_b200/packing.py pack_encoder_tail
 * (i2r_encoder_tail_weight_bytes(split) bytes of pre-swizzled fp16 weights; 768 fp32 biases and LayerNorm
 * parameters).  Replaces TransformerEncoderLayer.forward_post after self_attn (lib/models/transpose_h.py:205-222,
 * lib/models/attention.py:74-82, lib/models/interformer_pureMulti.py:205-213). */
int64_t i2r_encoder_tail_weight_bytes(int split);
int i2r_encoder_tail(const void* attn, int ld_attn, const void* src, int ld_src, const void* pos, void* out,
                     void* out_pos, int ld, const void* wimg, const float* params, int T, int d_model, int dim_ff,
                     float eps, int split, void* stream);

/* ---- HBM-bound building blocks of the HRFormer-B first stage (lib/models/hrformer.py) ----
 * Depthwise 3x3 convolution (padding 1, stride 1 or 2) + per-channel scale / bias (folded BatchNorm and conv bias) +
 * activation (0 none, 1 ReLU, 2 erf-GELU) on fp16 NHWC (split: pair tensors).  w: fp32 [9][C] tap-major.  Replaces
 * MlpDWBN.dw3x3/norm2/act2 (:1094-1119) and the depthwise stride-2 convs of the fuse layers (:1652-1703). */
int i2r_dwconv3x3(const void* x, const float* w, const float* scale, const float* bias, void* y, int NB, int H, int W,
                  int C, int stride, int act, int split, void* stream);
/* y = [relu](x0 + sum_k bilinear_up(t_k, 2^shift_k)), align_corners = False (F.interpolate arithmetic): the
 * higher-resolution outputs of HighResolutionTransformerModule's fuse layers (:1626-1644, :1714-1731) with the 1x1
 * conv + BN terms t_k computed at their own resolution.  t2 / t3 optional (NULL). */
int i2r_upsum_bilinear(const void* x0, const void* t1, int shift1, const void* t2, int shift2, const void* t3, int shift3,
                       void* y, int NB, int H, int W, int C, int relu, int split, void* stream);
/* LayerNorm over the first C_real channels of rows padded to C_pad (<= 768, multiple of 8) channels; pad channels are
 * written as zero.  eps is a parameter (1e-6 in GeneralTransformerBlock, :1198). */
int i2r_layernorm_padded(const void* x, const float* gamma, const float* beta, void* y, int rows, int C_real, int C_pad,
                         float eps, int split, void* stream);

/* Window-major token layout of InterlacedPoolAttention (:949-1000): the H x W map is centre-padded with zeros to
 * multiples of ws and cut into ws x ws windows; row = ((n*QH + qh)*QW + qw)*ws*ws + ph*ws + pw.
 * i2r_window_rows() = NB * Hp * Wp rows.  i2r_ln_window_gather: y[row] = LayerNorm(x[pixel]) (first C_real of C_pad
 * channels), zero rows at padded positions (the reference pads after norm1).  i2r_window_scatter_add:
 * y[pixel] = x[pixel] + a[row(pixel)] -- reverse permutation, de-pad and the residual of GeneralTransformerBlock (:1234). */
int64_t i2r_window_rows(int NB, int H, int W, int ws);
int i2r_ln_window_gather(const void* x, const float* gamma, const float* beta, void* y, int NB, int H, int W, int C_real,
                         int C_pad, int ws, float eps, int split, void* stream);
int i2r_window_scatter_add(const void* x, const void* a, void* y, int NB, int H, int W, int C, int ws, int split,
                           void* stream);
/* softmax(scale * q k^T) v for nwin windows of win_len consecutive token rows and `heads` heads of head_pad (= 48:
 * head_dim 39 zero-padded by the weight packing) channels each; no relative position bias (its addition is commented
 * out in the reference, :866-888) and no mask (padded tokens take part).  MHA_.forward (:627-935). */
int i2r_window_attention(const void* q, const void* k, const void* v, void* out, int ldq, int ldk, int ldv, int ldo,
                         int nwin, int win_len, int heads, int head_pad, float scale, int split, int q_lo, int k_lo,
                         int v_lo, int o_lo, void* stream);
/* The same operation on tcgen05 / TMEM / TMA for the shape HRFormer-B uses (win_len 49, head_pad 48): two windows per
 * 128-row tile, S = Q K^T and O = P V on the tensor cores (P from TMEM, V as an MN-major operand straight from the
 * token-major tile), block-diagonal softmax in the epilogue warps.  The product path; i2r_window_attention (mma.sync)
 * stays as the check implementation and for other window sizes.  Other shapes return I2R_E_UNSUPPORTED. */
int i2r_window_attention_tc(const void* q, const void* k, const void* v, void* out, int ldq, int ldk, int ldv, int ldo,
                         int nwin, int win_len, int heads, int head_pad, float scale, int split, int q_lo, int k_lo,
                         int v_lo, int o_lo, void* stream);

/* y = LayerNorm(x) * gamma + beta over the last dim C (<= 256, multiple of 8), fp16 in/out, fp32
 * math; optional y2 = y + pos (the next layer's q/k input).  (interformer_pureMulti.py:206,:209) */
int i2r_layernorm(const void* x, const float* gamma, const float* beta, const void* pos, void* y, void* y2,
                  int rows, int C, float eps, int split, void* stream);

/* y = a + b elementwise on fp16, n multiple of 8 (with_pos_embed, interformer_pureMulti.py:189; the residual
 * `single_res + x`, interformer.py:315).  split_c > 0: the operands are pair tensors with C = split_c. */
int i2r_add_f16(const void* a, const void* b, void* y, int64_t n, int split_c, void* stream);

/* y[n,h,w,:] = act(x0[n,h,w,:] + t1[n,h>>shift1,w>>shift1,:] + t2[n,h>>shift2,w>>shift2,:]) on fp16 NHWC (t2 may be
 * NULL; split != 0: pair tensors with C logical channels): the highest-resolution output branch of an HRNet fuse
 * layer -- identity + nearest-upsampled 1x1-conv terms evaluated at their own resolution, sum, ReLU
 * (interformer_pureMulti.py:392-410). */
int i2r_upsum(const void* x0, const void* t1, int shift1, const void* t2, int shift2, void* y, int NB, int H, int W,
              int C, int relu, int split, void* stream);

/* Profiling aid: while dev_buffer != NULL, CTA `cta` of every i2r_conv_halo launch writes (tag<<32 | tile, clock64)
 * pairs into four role regions (producer, MMA, epilogue, kernel start/end) of `capacity_events` pairs each (zero-filled by the caller).  Tags: 1/2/3
 * producer slot free / loads issued / stage published, 10/11/12 MMA accumulator free / operands landed / tile
 * committed, 20/21 epilogue accumulator ready / tile stored.  Pass NULL to switch tracing off. */
int i2r_debug_trace(void* dev_buffer, int capacity_events, int cta);
/* Profiling aid: ablation bits for i2r_conv_halo (results become wrong): 1 = epilogue hand-shake only,
 * 2 = epilogue without global stores / residual loads, 4 = activation TMA only for the first ring pass,
 * 8 = no MMAs (commits only).  0 restores the product behaviour. */
int i2r_debug_flags(int flags);
/* Profiling aid for i2r_conv_halo_chain (results become wrong with any bit set): 1 = tiles do not wait for their
 * producers, 2 = they poll but skip the acquire / proxy fences, 4 = finished tiles are published after the shared-memory
 * read of their TMA store instead of its completion.  0 restores the product behaviour. */
int i2r_debug_chain_flags(int flags);
/* Debug aid: every mbarrier wait of the tcgen05 kernels is time-bounded (2^31 SM cycles); a wait that times out traps
 * (the launch fails with cudaErrorLaunchFailure instead of hanging the GPU).  With a hang buffer installed --
 * HOST-MAPPED pinned memory of 4096 x 4 uint64, zero-filled by the caller, still readable after the context died --
 * lane 0 of every timed-out warp first records {blockIdx.x << 32 | threadIdx.x, source line << 32 | barrier shared
 * address, parity << 32 | dynamic shared base, clock64}.  NULL removes it.  Synchronises the device. */
int i2r_debug_hang_buffer(void* host_mapped);

/* `res` multi-person position embedding, stem (lib/models/position_embedding.py:14-18, :90-93): box masks fp32
 * [NB,1,H,W] -> conv_pre 3x3 (1 -> 3, w_pre fp32 [3][9]) -> resnet18 conv1 7x7 s2 p3 (3 -> 64, w1 fp32 [147][64] with
 * k = (c*7 + ky)*7 + kx) -> folded bn1 (scale / bias fp32 [64]) -> ReLU, fp16 NHWC [NB, H/2, W/2, 64] (split != 0: pair
 * tensor with 128 values per pixel).  The layers after it (max-pool, layer1, conv_end) use the generic entry points. */
int i2r_mask_res_stem(const float* mask, const float* w_pre, const float* w1, const float* scale, const float* bias,
                      void* y, int NB, int H, int W, int split, void* stream);

/* ---- post-processing on the device (N1 flip-test fusion, N2 heatmap decode) ------------------------ */

/* dst[r, w] = src[r, W-1-w] on fp32 rows of W elements (out of place): np.flip(input, 3) of the flip test
 * (lib/core/function.py:145-149) for x [S,3,H,W] and pos_mask [S,1,H,W] viewed as rows. */
int i2r_hflip_f32(const float* src, float* dst, int64_t rows, int W, void* stream);

/* y[s,k,h,w] = 0.5 * (out[s,k,h,w] + out_flipped[s, perm[k], h, W-1-w]): `flip_back` (lib/utils/transforms.py:16-30:
 * reverse the width axis, swap the matched left/right joints -- `perm` is that permutation as K int32 on the device)
 * fused with `(output + output_flipped) * 0.5` (lib/core/function.py:158-162).  fp32 NCHW heatmaps. */
int i2r_flip_merge(const float* out, const float* out_flipped, float* y, int S, int K, int H, int W, const int32_t* perm,
                   void* stream);

/* get_final_preds (lib/core/inference.py:90-112) for fp32 heatmaps hm [S,K,H,W] on the device: get_max_preds (:20-48:
 * arg-max, x = idx % W, y = floor(idx / W), zeroed when the maximum is <= 0), gaussian_blur (:73-87: zero-padded
 * cv2.GaussianBlur(ksize = blur_kernel, sigma from ksize) in double, renormalised to the original maximum), log of
 * max(., 1e-10), taylor (:51-70: second-order step, interior maxima only), and, when transform_back != 0,
 * transform_preds (lib/utils/transforms.py:50-56) with center / scale [S,2] fp32 on the device.
 * preds [S,K,2] fp32 (x, y), maxvals [S,K] fp32.  One CTA per heatmap; H*W*16 bytes of shared memory (<= 200 KB). */
int i2r_decode_heatmaps(const float* hm, int S, int K, int H, int W, const float* center, const float* scale,
                        int blur_kernel, int transform_back, float* preds, float* maxvals, void* stream);

/* ---- input pipeline on the device (N3) -------------------------------------------------------------- */

/* Person crops of ONE image: x[n] = Normalize(ToTensor(cv2.warpAffine(image, trans_n, (OW, OH), INTER_LINEAR)))
 * (lib/dataset/JointsDataset.py:296-303, :329-330; tools/test.py:126-134).  image: uint8 RGB [IH, IW, 3] on the device;
 * inv_affine: [N, 6] doubles on the device = the dst -> src matrix cv2.warpAffine derives from the forward matrix of
 * get_affine_transform (lib/utils/transforms.py:58-92); mean3 / std3: host pointers to 3 floats; x: fp32 [N,3,OH,OW].
 * 8-bit fixed-point arithmetic of cv2 (bit-exact against cv2 4.13). */
int i2r_crop_persons(const uint8_t* image, int IH, int IW, const double* inv_affine, int N, int OH, int OW,
                     const float* mean3, const float* std3, float* x, void* stream);

/* Per-person box masks of one image: ToTensor(cv2.resize(rotate_bound(get_position(box), 0), (OW, OH)))
 * (lib/dataset/JointsDataset.py:166-201, :323-331).  rect: int32 [N, 4] on the device, the inclusive corners
 * (x0, y0, x1, y1) cv2.rectangle fills = (int(x), int(y), int(x+w), int(y+h)) ordered; pos_mask: fp32 [N,1,OH,OW]. */
int i2r_box_masks(const int32_t* rect, int N, int IH, int IW, int OH, int OW, float* pos_mask, void* stream);

/* sizeof(i2r_conv_problem) as compiled -- lets the ctypes binding verify its struct layout. */
int i2r_sizeof_conv_problem(void);

#ifdef __cplusplus
}
#endif
#endif /* I2R_H_ */

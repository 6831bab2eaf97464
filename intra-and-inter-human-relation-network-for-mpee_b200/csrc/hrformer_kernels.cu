// HBM-bound building blocks of the HRFormer-B first stage (reference lib/models/hrformer.py):
//   * depthwise 3x3 convolution + folded BatchNorm + {none, ReLU, erf-GELU}, stride 1 (MlpDWBN.dw3x3, :1094-1119) or
//     stride 2 (the down-sampling fuse / transition chains, :1652-1703);
//   * bilinear (align_corners = False) x2^k up-sampling of up to three lower-resolution terms accumulated onto the
//     identity branch + ReLU (the fuse layers of HighResolutionTransformerModule, :1626-1644, :1714-1731);
//   * LayerNorm over the first C_real channels of rows padded to C_pad channels (78 -> 80, 156 -> 160, ...: the K = 16
//     granularity of the tensor-core GEMMs), eps parameter (1e-6 in GeneralTransformerBlock, :1198), pad channels
//     written as zero so they stay inert in every following GEMM.
// fp16 NHWC tensors; PAIR = split-operand pair tensors [hi(C) | lo(C)] (include/i2r.h, I2R_F_SPLIT); fp32 math.
#include "i2r_common.cuh"

namespace i2r {

__device__ __forceinline__ void hk_load8(const __half* p, float (&v)[8]) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = unpack_h2(w4[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <bool PAIR>
__device__ __forceinline__ void hk_load(const __half* row, int C, int c, float (&v)[8]) {
  hk_load8(row + c, v);
  if (PAIR) {
    float lo[8];
    hk_load8(row + C + c, lo);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += lo[i];
  }
}
template <bool PAIR>
__device__ __forceinline__ void hk_store(__half* row, int C, int c, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_h2(v[2 * i], v[2 * i + 1]);
    const float2 f = unpack_h2(h[i]);
    l[i] = pack_h2(v[2 * i] - f.x, v[2 * i + 1] - f.y);
  }
  *reinterpret_cast<uint4*>(row + c) = make_uint4(h[0], h[1], h[2], h[3]);
  if (PAIR) *reinterpret_cast<uint4*>(row + C + c) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ float hk_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return gelu_erf(v);   // nn.GELU (erf form)
  return v;
}
static int hk_grid(int64_t items, int block) {
  int64_t g = (items + block - 1) / block;
  const int64_t cap = 148 * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3
// thread = (output pixel, 8 channels); w fp32 [9][C] (tap-major), scale / bias fp32 [C]
template <bool PAIR>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const __half* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ bias, __half* __restrict__ y, int NB,
                                                        int H, int W, int C, int stride, int act) {
  pdl_launch_dependents();
  pdl_wait();
  const int OH = (H + stride - 1) / stride, OW = (W + stride - 1) / stride;
  const int cv = C >> 3;
  const int ld = PAIR ? 2 * C : C;
  const int64_t total = static_cast<int64_t>(NB) * OH * OW * cv;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % cv) * 8;
    const int64_t p = idx / cv;
    const int ox = static_cast<int>(p % OW);
    const int oy = static_cast<int>((p / OW) % OH);
    const int n = static_cast<int>(p / (static_cast<int64_t>(OW) * OH));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * stride - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * stride - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        hk_load<PAIR>(x + ((static_cast<int64_t>(n) * H + iy) * W + ix) * ld, C, c, v);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + c));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + c + 4));
        acc[0] = fmaf(v[0], w0.x, acc[0]);
        acc[1] = fmaf(v[1], w0.y, acc[1]);
        acc[2] = fmaf(v[2], w0.z, acc[2]);
        acc[3] = fmaf(v[3], w0.w, acc[3]);
        acc[4] = fmaf(v[4], w1.x, acc[4]);
        acc[5] = fmaf(v[5], w1.y, acc[5]);
        acc[6] = fmaf(v[6], w1.z, acc[6]);
        acc[7] = fmaf(v[7], w1.w, acc[7]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = hk_act(fmaf(acc[i], __ldg(scale + c + i), __ldg(bias + c + i)), act);
    hk_store<PAIR>(y + p * ld, C, c, acc);
  }
}

// Stride-1 variant with a shared-memory halo patch: CTA = 8 x 8 output pixels x 64 channels.  The (10 x 10 pixel) patch is
// loaded once (1.56x the tile instead of the 9x re-reads of the kernel above, which made the 320-channel MlpDWBN layer of
// HRFormer's highest-resolution branch run at 1.2 TB/s of L2 traffic), weights of the 64 channels sit in shared memory.
constexpr int DW_T = 8, DW_P = DW_T + 2, DW_CB = 64;
template <bool PAIR>
__global__ void __launch_bounds__(256) dwconv3x3_tile_kernel(const __half* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ bias, __half* __restrict__ y, int NB,
                                                             int H, int W, int C, int act) {
  __shared__ __align__(16) __half patch[(PAIR ? 2 : 1) * DW_P * DW_P * DW_CB];
  __shared__ __align__(16) float sw[9 * DW_CB];
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  const int cblocks = C / DW_CB;
  const int tiles_x = (W + DW_T - 1) / DW_T, tiles_y = (H + DW_T - 1) / DW_T;
  int b = blockIdx.x;
  const int cb = b % cblocks;
  b /= cblocks;
  const int tx = b % tiles_x;
  b /= tiles_x;
  const int ty = b % tiles_y;
  const int n = b / tiles_y;
  const int c0 = cb * DW_CB;
  const int ld = PAIR ? 2 * C : C;
  for (int i = tid; i < 9 * DW_CB; i += 256) sw[i] = __ldg(w + (i / DW_CB) * C + c0 + (i % DW_CB));
  pdl_wait();
  // patch[half][py][px][64 ch]: 8 chunks of 16 bytes per pixel and half
  constexpr int CHUNKS = DW_P * DW_P * (DW_CB / 8);
  for (int i = tid; i < CHUNKS * (PAIR ? 2 : 1); i += 256) {
    const int half = i / CHUNKS, j = i - half * CHUNKS;
    const int g = j % (DW_CB / 8), pp = j / (DW_CB / 8);
    const int iy = ty * DW_T - 1 + pp / DW_P, ix = tx * DW_T - 1 + pp % DW_P;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<int64_t>(n) * H + iy) * W + ix) * ld + half * C + c0 + g * 8));
    *reinterpret_cast<uint4*>(patch + (half * DW_P * DW_P + pp) * DW_CB + g * 8) = v;
  }
  __syncthreads();
  for (int it = tid; it < DW_T * DW_T * (DW_CB / 8); it += 256) {
    const int g = it % (DW_CB / 8), pp = it / (DW_CB / 8);
    const int oy = pp / DW_T, ox = pp % DW_T;
    const int gy = ty * DW_T + oy, gx = tx * DW_T + ox;
    if (gy >= H || gx >= W) continue;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int pi = (oy + t / 3) * DW_P + ox + t % 3;
      const uint4 hq = *reinterpret_cast<const uint4*>(patch + pi * DW_CB + g * 8);
      const uint32_t hw4[4] = {hq.x, hq.y, hq.z, hq.w};
      float v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_h2(hw4[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
      }
      if (PAIR) {
        const uint4 lq = *reinterpret_cast<const uint4*>(patch + (DW_P * DW_P + pi) * DW_CB + g * 8);
        const uint32_t lw4[4] = {lq.x, lq.y, lq.z, lq.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack_h2(lw4[i]);
          v[2 * i] += f.x;
          v[2 * i + 1] += f.y;
        }
      }
      const float4 w0 = *reinterpret_cast<const float4*>(sw + t * DW_CB + g * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(sw + t * DW_CB + g * 8 + 4);
      acc[0] = fmaf(v[0], w0.x, acc[0]);
      acc[1] = fmaf(v[1], w0.y, acc[1]);
      acc[2] = fmaf(v[2], w0.z, acc[2]);
      acc[3] = fmaf(v[3], w0.w, acc[3]);
      acc[4] = fmaf(v[4], w1.x, acc[4]);
      acc[5] = fmaf(v[5], w1.y, acc[5]);
      acc[6] = fmaf(v[6], w1.z, acc[6]);
      acc[7] = fmaf(v[7], w1.w, acc[7]);
    }
    const int c = c0 + g * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = hk_act(fmaf(acc[i], __ldg(scale + c + i), __ldg(bias + c + i)), act);
    hk_store<PAIR>(y + ((static_cast<int64_t>(n) * H + gy) * W + gx) * ld, C, c, acc);
  }
}

// ------------------------------------------------------------------------------------------------ bilinear fuse
// y = act(x0 + sum_k bilinear_up(t_k, 2^s_k)), torch F.interpolate(mode='bilinear', align_corners=False) arithmetic:
// src = max(0, (dst + 0.5) / 2^s - 0.5), i0 = floor(src), i1 = min(i0 + 1, in - 1), l1 = src - i0
struct BilinearTerm {
  const __half* t;
  int shift;
};
template <bool PAIR>
__device__ __forceinline__ void bilinear_add(const BilinearTerm T, int n, int hy, int wx, int H, int W, int C, int c,
                                             float (&v)[8]) {
  const int h = H >> T.shift, w = W >> T.shift;
  const int ld = PAIR ? 2 * C : C;
  const float inv = 1.f / static_cast<float>(1 << T.shift);
  const float sy = fmaxf((hy + 0.5f) * inv - 0.5f, 0.f), sx = fmaxf((wx + 0.5f) * inv - 0.5f, 0.f);
  const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly = sy - y0, lx = sx - x0;
  const __half* base = T.t + static_cast<int64_t>(n) * h * w * ld;
  float a[8], b[8], cc[8], d[8];
  hk_load<PAIR>(base + (static_cast<int64_t>(y0) * w + x0) * ld, C, c, a);
  hk_load<PAIR>(base + (static_cast<int64_t>(y0) * w + x1) * ld, C, c, b);
  hk_load<PAIR>(base + (static_cast<int64_t>(y1) * w + x0) * ld, C, c, cc);
  hk_load<PAIR>(base + (static_cast<int64_t>(y1) * w + x1) * ld, C, c, d);
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] += w00 * a[i] + w01 * b[i] + w10 * cc[i] + w11 * d[i];
}

template <bool PAIR>
__global__ void __launch_bounds__(256) upsum_bilinear_kernel(const __half* __restrict__ x0, BilinearTerm t1,
                                                             BilinearTerm t2, BilinearTerm t3,
                                                             __half* __restrict__ y, int NB, int H, int W, int C,
                                                             int relu) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = C >> 3;
  const int ld = PAIR ? 2 * C : C;
  const int64_t total = static_cast<int64_t>(NB) * H * W * cv;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % cv) * 8;
    const int64_t p = idx / cv;
    const int wx = static_cast<int>(p % W);
    const int hy = static_cast<int>((p / W) % H);
    const int n = static_cast<int>(p / (static_cast<int64_t>(W) * H));
    float v[8];
    hk_load<PAIR>(x0 + p * ld, C, c, v);
    bilinear_add<PAIR>(t1, n, hy, wx, H, W, C, c, v);
    if (t2.t != nullptr) bilinear_add<PAIR>(t2, n, hy, wx, H, W, C, c, v);
    if (t3.t != nullptr) bilinear_add<PAIR>(t3, n, hy, wx, H, W, C, c, v);
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    hk_store<PAIR>(y + p * ld, C, c, v);
  }
}

// ------------------------------------------------------------------------------------------------ padded LayerNorm
// one warp per row; lane handles the 8-channel groups lane, lane + 32, lane + 64 (C_pad <= 768)
template <bool PAIR, int LPR>
__global__ void __launch_bounds__(256) layernorm_padded_kernel(const __half* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, __half* __restrict__ y,
                                                               int rows, int Cr, int Cp, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int row_raw = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int lane = threadIdx.x % LPR;      // LPR lanes share a row (16 for rows of <= 128 channels: two rows per warp)
  const bool live = row_raw < rows;        // no early return: the full-mask shuffles below need every lane of the warp
  const int row = live ? row_raw : rows - 1;
  const int64_t ld = PAIR ? 2 * Cp : Cp;
  const __half* xr = x + row * ld;
  float v[3][8];
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int c = (lane + LPR * g) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[g][i] = 0.f;
    if (c < Cp) {
      hk_load<PAIR>(xr, Cp, c, v[g]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (c + i >= Cr) v[g][i] = 0.f;
        s += v[g][i];
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / Cr;
  float sq = 0.f;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int c = (lane + LPR * g) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (c + i < Cr) {
        const float d = v[g][i] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / Cr + eps);
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int c = (lane + LPR * g) * 8;
    if (c < Cp) {
      float o8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o8[i] = (c + i < Cr) ? (v[g][i] - mean) * rstd * __ldg(gamma + c + i) + __ldg(beta + c + i) : 0.f;
      if (live) hk_store<PAIR>(y + row * ld, Cp, c, o8);
    }
  }
}

// ------------------------------------------------------------------------------------------------ window layout
// Window-major token layout of InterlacedPoolAttention (lib/models/hrformer.py:949-986): the H x W map is centre-padded
// with zeros to multiples of ws, cut into ws x ws windows; output row = ((n * QH + qh) * QW + qw) * ws*ws + ph * ws + pw.
// ln_window_gather = LayerNorm (first C_real of C_pad channels) of every pixel written at its window row; rows of padded
// positions are zero (the reference pads AFTER norm1, so their q/k/v are the projection biases).
template <bool PAIR, int LPR>
__global__ void __launch_bounds__(256) ln_window_gather_kernel(const __half* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, __half* __restrict__ y,
                                                               int NB, int H, int W, int Cr, int Cp, int ws, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int Hp = (H + ws - 1) / ws * ws, Wp = (W + ws - 1) / ws * ws;
  const int pt = (Hp - H) / 2, pl = (Wp - W) / 2;
  const int QH = Hp / ws, QW = Wp / ws;
  const int64_t rows = static_cast<int64_t>(NB) * Hp * Wp;
  const int64_t row_raw = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / LPR;
  const int lane = threadIdx.x % LPR;      // LPR lanes share a row (16 for rows of <= 128 channels: two rows per warp)
  const bool live = row_raw < rows;        // no early return: the full-mask shuffles below need every lane of the warp
  const int64_t row = live ? row_raw : rows - 1;
  const int pos = static_cast<int>(row % (ws * ws));
  const int64_t win = row / (ws * ws);
  const int qw = static_cast<int>(win % QW), qh = static_cast<int>((win / QW) % QH);
  const int n = static_cast<int>(win / (static_cast<int64_t>(QW) * QH));
  const int py = qh * ws + pos / ws - pt, px = qw * ws + pos % ws - pl;
  const int64_t ld = PAIR ? 2 * Cp : Cp;
  __half* yr = y + row * ld;
  const bool inside = py >= 0 && py < H && px >= 0 && px < W;
  const __half* xr = x + ((static_cast<int64_t>(n) * H + (inside ? py : 0)) * W + (inside ? px : 0)) * ld;
  float v[3][8];
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int c = (lane + LPR * g) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[g][i] = 0.f;
    if (c < Cp && inside) {
      hk_load<PAIR>(xr, Cp, c, v[g]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (c + i >= Cr) v[g][i] = 0.f;
        s += v[g][i];
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / Cr;
  float sq = 0.f;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int c = (lane + LPR * g) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (c + i < Cr) {
        const float d = v[g][i] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / Cr + eps);
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int c = (lane + LPR * g) * 8;
    if (c < Cp) {
      float o8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o8[i] = (inside && c + i < Cr) ? (v[g][i] - mean) * rstd * __ldg(gamma + c + i) + __ldg(beta + c + i) : 0.f;
      if (live) hk_store<PAIR>(yr, Cp, c, o8);
    }
  }
}

// y[pixel] = x[pixel] + a[window row of that pixel]  (reverse permutation + de-pad + residual, :987-1000, :1234)
template <bool PAIR>
__global__ void __launch_bounds__(256) window_scatter_add_kernel(const __half* __restrict__ x,
                                                                 const __half* __restrict__ a, __half* __restrict__ y,
                                                                 int NB, int H, int W, int C, int ws) {
  pdl_launch_dependents();
  pdl_wait();
  const int Hp = (H + ws - 1) / ws * ws, Wp = (W + ws - 1) / ws * ws;
  const int pt = (Hp - H) / 2, pl = (Wp - W) / 2;
  const int QH = Hp / ws, QW = Wp / ws;
  const int cv = C >> 3;
  const int ld = PAIR ? 2 * C : C;
  const int64_t total = static_cast<int64_t>(NB) * H * W * cv;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % cv) * 8;
    const int64_t p = idx / cv;
    const int wx = static_cast<int>(p % W);
    const int hy = static_cast<int>((p / W) % H);
    const int n = static_cast<int>(p / (static_cast<int64_t>(W) * H));
    const int yy = hy + pt, xx = wx + pl;
    const int64_t row = ((static_cast<int64_t>(n) * QH + yy / ws) * QW + xx / ws) * (ws * ws) + (yy % ws) * ws + xx % ws;
    float v[8], t[8];
    hk_load<PAIR>(x + p * ld, C, c, v);
    hk_load<PAIR>(a + row * ld, C, c, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += t[i];
    hk_store<PAIR>(y + p * ld, C, c, v);
  }
}

}  // namespace i2r

using namespace i2r;

extern "C" int i2r_dwconv3x3(const void* x, const float* w, const float* scale, const float* bias, void* y, int NB, int H,
                             int W, int C, int stride, int act, int split, void* stream) {
  if (!x || !w || !scale || !bias || !y || NB <= 0 || H <= 0 || W <= 0 || C % 8 != 0 || (stride != 1 && stride != 2) ||
      act < 0 || act > 2) {
    set_error("i2r_dwconv3x3: bad arguments (C=%d stride=%d act=%d)", C, stride, act);
    return I2R_E_BADARG;
  }
  const int OH = (H + stride - 1) / stride, OW = (W + stride - 1) / stride;
  const int64_t items = static_cast<int64_t>(NB) * OH * OW * (C / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (stride == 1 && C % DW_CB == 0) {      // shared-memory tiled variant
    const int64_t ctas = static_cast<int64_t>(NB) * ((H + DW_T - 1) / DW_T) * ((W + DW_T - 1) / DW_T) * (C / DW_CB);
    if (ctas < (1ll << 31)) {
      if (split)
        launch_pdl(dwconv3x3_tile_kernel<true>, dim3(static_cast<unsigned>(ctas)), dim3(256), 0, st,
                   static_cast<const __half*>(x), w, scale, bias, static_cast<__half*>(y), NB, H, W, C, act);
      else
        launch_pdl(dwconv3x3_tile_kernel<false>, dim3(static_cast<unsigned>(ctas)), dim3(256), 0, st,
                   static_cast<const __half*>(x), w, scale, bias, static_cast<__half*>(y), NB, H, W, C, act);
      return check_launch("dwconv3x3_tile_kernel");
    }
  }
  if (split) {
    launch_pdl(dwconv3x3_kernel<true>, dim3(hk_grid(items, 256)), dim3(256), 0, st, static_cast<const __half*>(x), w,
               scale, bias, static_cast<__half*>(y), NB, H, W, C, stride, act);
  } else {
    launch_pdl(dwconv3x3_kernel<false>, dim3(hk_grid(items, 256)), dim3(256), 0, st, static_cast<const __half*>(x), w,
               scale, bias, static_cast<__half*>(y), NB, H, W, C, stride, act);
  }
  return check_launch("dwconv3x3_kernel");
}

extern "C" int i2r_upsum_bilinear(const void* x0, const void* t1, int shift1, const void* t2, int shift2, const void* t3,
                                  int shift3, void* y, int NB, int H, int W, int C, int relu, int split, void* stream) {
  auto bad = [&](const void* t, int s) { return t != nullptr && (s < 1 || s > 4 || ((H >> s) << s) != H || ((W >> s) << s) != W); };
  if (!x0 || !t1 || !y || NB <= 0 || H <= 0 || W <= 0 || C % 8 != 0 || bad(t1, shift1) || bad(t2, shift2) ||
      bad(t3, shift3) || (t3 && !t2)) {
    set_error("i2r_upsum_bilinear: bad arguments (H=%d W=%d C=%d shifts %d %d %d)", H, W, C, shift1, shift2, shift3);
    return I2R_E_BADARG;
  }
  const int64_t items = static_cast<int64_t>(NB) * H * W * (C / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const BilinearTerm a{static_cast<const __half*>(t1), shift1}, b{static_cast<const __half*>(t2), shift2},
      c{static_cast<const __half*>(t3), shift3};
  if (split) {
    launch_pdl(upsum_bilinear_kernel<true>, dim3(hk_grid(items, 256)), dim3(256), 0, st, static_cast<const __half*>(x0),
               a, b, c, static_cast<__half*>(y), NB, H, W, C, relu);
  } else {
    launch_pdl(upsum_bilinear_kernel<false>, dim3(hk_grid(items, 256)), dim3(256), 0, st,
               static_cast<const __half*>(x0), a, b, c, static_cast<__half*>(y), NB, H, W, C, relu);
  }
  return check_launch("upsum_bilinear_kernel");
}

extern "C" int i2r_layernorm_padded(const void* x, const float* gamma, const float* beta, void* y, int rows, int C_real,
                                    int C_pad, float eps, int split, void* stream) {
  if (!x || !gamma || !beta || !y || rows <= 0 || C_pad % 8 != 0 || C_pad > 768 || C_real < 1 || C_real > C_pad) {
    set_error("i2r_layernorm_padded: bad arguments (rows=%d C_real=%d C_pad=%d)", rows, C_real, C_pad);
    return I2R_E_BADARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool narrow = C_pad <= 128;      // 16 lanes per row: two rows per warp
  const int rpb = narrow ? 16 : 8;       // rows per 256-thread block
  const dim3 grid((rows + rpb - 1) / rpb);
  const __half* xs = static_cast<const __half*>(x);
  __half* ys = static_cast<__half*>(y);
  if (split) {
    if (narrow) launch_pdl(layernorm_padded_kernel<true, 16>, grid, dim3(256), 0, st, xs, gamma, beta, ys, rows, C_real, C_pad, eps);
    else launch_pdl(layernorm_padded_kernel<true, 32>, grid, dim3(256), 0, st, xs, gamma, beta, ys, rows, C_real, C_pad, eps);
  } else {
    if (narrow) launch_pdl(layernorm_padded_kernel<false, 16>, grid, dim3(256), 0, st, xs, gamma, beta, ys, rows, C_real, C_pad, eps);
    else launch_pdl(layernorm_padded_kernel<false, 32>, grid, dim3(256), 0, st, xs, gamma, beta, ys, rows, C_real, C_pad, eps);
  }
  return check_launch("layernorm_padded_kernel");
}

extern "C" int64_t i2r_window_rows(int NB, int H, int W, int ws) {
  if (NB <= 0 || H <= 0 || W <= 0 || ws <= 0) return 0;
  return static_cast<int64_t>(NB) * ((H + ws - 1) / ws * ws) * ((W + ws - 1) / ws * ws);
}

extern "C" int i2r_ln_window_gather(const void* x, const float* gamma, const float* beta, void* y, int NB, int H, int W,
                                    int C_real, int C_pad, int ws, float eps, int split, void* stream) {
  if (!x || !gamma || !beta || !y || NB <= 0 || H <= 0 || W <= 0 || ws <= 0 || C_pad % 8 != 0 || C_pad > 768 ||
      C_real < 1 || C_real > C_pad) {
    set_error("i2r_ln_window_gather: bad arguments (C_real=%d C_pad=%d ws=%d)", C_real, C_pad, ws);
    return I2R_E_BADARG;
  }
  const int64_t rows = i2r_window_rows(NB, H, W, ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool narrow = C_pad <= 128;      // 16 lanes per row: two rows per warp
  const int rpb = narrow ? 16 : 8;
  const dim3 grid(static_cast<unsigned>((rows + rpb - 1) / rpb));
  const __half* xs = static_cast<const __half*>(x);
  __half* ys = static_cast<__half*>(y);
  if (split) {
    if (narrow) launch_pdl(ln_window_gather_kernel<true, 16>, grid, dim3(256), 0, st, xs, gamma, beta, ys, NB, H, W, C_real, C_pad, ws, eps);
    else launch_pdl(ln_window_gather_kernel<true, 32>, grid, dim3(256), 0, st, xs, gamma, beta, ys, NB, H, W, C_real, C_pad, ws, eps);
  } else {
    if (narrow) launch_pdl(ln_window_gather_kernel<false, 16>, grid, dim3(256), 0, st, xs, gamma, beta, ys, NB, H, W, C_real, C_pad, ws, eps);
    else launch_pdl(ln_window_gather_kernel<false, 32>, grid, dim3(256), 0, st, xs, gamma, beta, ys, NB, H, W, C_real, C_pad, ws, eps);
  }
  return check_launch("ln_window_gather_kernel");
}

extern "C" int i2r_window_scatter_add(const void* x, const void* a, void* y, int NB, int H, int W, int C, int ws,
                                      int split, void* stream) {
  if (!x || !a || !y || NB <= 0 || H <= 0 || W <= 0 || ws <= 0 || C % 8 != 0) {
    set_error("i2r_window_scatter_add: bad arguments");
    return I2R_E_BADARG;
  }
  const int64_t items = static_cast<int64_t>(NB) * H * W * (C / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (split) {
    launch_pdl(window_scatter_add_kernel<true>, dim3(hk_grid(items, 256)), dim3(256), 0, st,
               static_cast<const __half*>(x), static_cast<const __half*>(a), static_cast<__half*>(y), NB, H, W, C, ws);
  } else {
    launch_pdl(window_scatter_add_kernel<false>, dim3(hk_grid(items, 256)), dim3(256), 0, st,
               static_cast<const __half*>(x), static_cast<const __half*>(a), static_cast<__half*>(y), NB, H, W, C, ws);
  }
  return check_launch("window_scatter_add_kernel");
}

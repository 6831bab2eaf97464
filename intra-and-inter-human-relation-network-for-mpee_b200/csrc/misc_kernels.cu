// HBM-bound helper kernels: tiny-Cin stem convolution (fp32 NCHW -> fp16 NHWC), max-pool,
// LayerNorm (+ fused positional add), elementwise add.  Warp-shuffle / 16-byte vectorised.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "i2r_common.cuh"

namespace i2r {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("I2R_PDL");
    v = (e != nullptr && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return v != 0;
}

void hang_sink_attention_tc(void*);
void hang_sink_conv_halo(void*);
void hang_sink_encoder_tail(void*);
void hang_sink_igemm_tc(void*);
void hang_sink_stem_tc(void*);
void hang_sink_window_attention_tc(void*);

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

// ---- split-operand ("pair") tensors: a row of 2*C fp16 = [hi(C) | lo(C)], value = hi + lo (include/i2r.h, I2R_F_SPLIT)
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = unpack_h2(w4[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8_pair(const __half* row, int C, int c, float (&v)[8]) {
  float lo[8];
  load8(row + c, v);
  load8(row + C + c, lo);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] += lo[i];
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
}
__device__ __forceinline__ void store8_pair(__half* row, int C, int c, const float (&v)[8]) {
  float lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) lo[i] = v[i] - __half2float(__float2half_rn(v[i]));
  store8(row + c, v);
  store8(row + C + c, lo);
}

// ------------------------------------------------------------------------------------------------
// 3x3 stride-2 pad-1 convolution with Cin in {1,3} and 64 output channels (stem / mask embedding).
// CTA = 8 x 32 output pixels: the (17 x 65 x CIN) fp32 input patch is staged in shared memory with coalesced
// loads; each thread owns one output pixel and all 64 channels (weights are warp-broadcast float4 reads),
// then writes its 128-byte NHWC row.
// CTA = 8 x 32 output pixels: the (17 x 65 x CIN) fp32 input patch is staged in shared memory with coalesced loads.
// A thread owns FOUR horizontally adjacent output pixels x 16 channels (a quarter of the 64): per tap it reads four
// patch values and four float4 weight vectors for 64 FMAs, i.e. one shared-memory load per 8 FMAs (the first version
// -- one pixel x 64 channels per thread -- issued one LDS per 4 FMAs and ran LDS-bound at 14x the HBM time).
constexpr int STEM_TH = 8, STEM_TW = 32, STEM_PH = 2 * STEM_TH + 1, STEM_PW = 2 * STEM_TW + 1, STEM_PWP = STEM_PW + 4;

template <int CIN>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ bias, __half* __restrict__ y,
                                                        int NB, int H, int W, int split) {
  constexpr int K = CIN * 9;
  pdl_launch_dependents();
  __shared__ __align__(16) float sw[K * 64];
  __shared__ __align__(16) float ssc[64];
  __shared__ __align__(16) float sbi[64];
  __shared__ float patch[CIN][STEM_PH][STEM_PWP];
  const int tid = threadIdx.x;
  // weights re-ordered so that the four channel quarters read by the lanes of one instruction are 64 contiguous
  // bytes: float4 slot (k*4 + q)*4 + cq holds channels [cq*16 + q*4, +4) of tap k (bank-conflict free)
  for (int i = tid; i < K * 64; i += 256) {
    const int k = i >> 6, ch = i & 63;
    const int cq_ = ch >> 4, q_ = (ch >> 2) & 3, e_ = ch & 3;
    sw[((k * 4 + q_) * 4 + cq_) * 4 + e_] = w[i];
  }
  if (tid < 64) {
    ssc[tid] = scale[tid];
    sbi[tid] = bias[tid];
  }
  pdl_wait();   // weights / scale / bias above are constants; the input and output below are not
  const int OH = H >> 1, OW = W >> 1;
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * STEM_TH, ox0 = blockIdx.x * STEM_TW;
  const int iy0 = oy0 * 2 - 1, ix0 = ox0 * 2 - 1;
  for (int i = tid; i < CIN * STEM_PH * STEM_PW; i += 256) {
    const int c = i / (STEM_PH * STEM_PW);
    const int r = i - c * (STEM_PH * STEM_PW);
    const int py = r / STEM_PW, px = r - py * STEM_PW;
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((static_cast<int64_t>(n) * CIN + c) * H + iy) * W + ix);
    patch[c][py][px + (px >> 5)] = v;   // skew one word per 32 columns: the 8 pixel groups of a warp hit 8 banks
  }
  __syncthreads();
  // thread -> (row ty, pixel group pg of 4 pixels, channel quarter cq): consecutive lanes take consecutive channel
  // quarters of the same pixels, so a warp's stores cover 8 pixel groups x 128 B contiguous rows
  const int cq = tid & 3, pg = (tid >> 2) & 7, ty = tid >> 5;
  float acc[4][16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[p][i] = 0.f;
  const float4* sw4 = reinterpret_cast<const float4*>(sw);
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      // the 4 pixels x 3 taps of this row touch 9 consecutive patch columns
      float pv[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) pv[j] = patch[c][2 * ty + ky][8 * pg + j + ((8 * pg + j) >> 5)];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int k = (c * 3 + ky) * 3 + kx;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wq = sw4[(k * 4 + q) * 4 + cq];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float v = pv[2 * p + kx];
            acc[p][4 * q + 0] = fmaf(v, wq.x, acc[p][4 * q + 0]);
            acc[p][4 * q + 1] = fmaf(v, wq.y, acc[p][4 * q + 1]);
            acc[p][4 * q + 2] = fmaf(v, wq.z, acc[p][4 * q + 2]);
            acc[p][4 * q + 3] = fmaf(v, wq.w, acc[p][4 * q + 3]);
          }
        }
      }
    }
  }
  const int oy = oy0 + ty;
  if (oy < OH) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int ox = ox0 + 4 * pg + p;
      if (ox >= OW) continue;
      __half* row = y + ((static_cast<int64_t>(n) * OH + oy) * OW + ox) * (split ? 128 : 64);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ch = cq * 16 + h * 8 + i;
          o[i] = fmaxf(acc[p][h * 8 + i] * ssc[ch] + sbi[ch], 0.f);
        }
        if (split) {
          store8_pair(row, 64, cq * 16 + h * 8, o);
        } else {
          store8(row + cq * 16 + h * 8, o);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                           int NB, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int OH = (H + 1) >> 1, OW = (W + 1) >> 1;
  const int cv = C >> 3;
  const int64_t total = static_cast<int64_t>(NB) * OH * OW * cv;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % cv);
    const int64_t p = idx / cv;
    const int ox = static_cast<int>(p % OW);
    const int oy = static_cast<int>((p / OW) % OH);
    const int n = static_cast<int>(p / (static_cast<int64_t>(OW) * OH));
    __half2 m[4];
    const __half2 ninf = __float2half2_rn(-65504.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = ninf;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        const uint4 q = *reinterpret_cast<const uint4*>(x + ((static_cast<int64_t>(n) * H + iy) * W + ix) * C + c8 * 8);
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], h[i]);
      }
    }
    *reinterpret_cast<uint4*>(y + p * C + c8 * 8) = *reinterpret_cast<uint4*>(m);
  }
}

// pair tensors: max over the VALUES hi + lo, re-split on store
__global__ void __launch_bounds__(256) maxpool3x3s2_pair_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                                int NB, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int OH = (H + 1) >> 1, OW = (W + 1) >> 1;
  const int cv = C >> 3;
  const int64_t total = static_cast<int64_t>(NB) * OH * OW * cv;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % cv);
    const int64_t p = idx / cv;
    const int ox = static_cast<int>(p % OW);
    const int oy = static_cast<int>((p / OW) % OH);
    const int n = static_cast<int>(p / (static_cast<int64_t>(OW) * OH));
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -3.0e38f;
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        load8_pair(x + ((static_cast<int64_t>(n) * H + iy) * W + ix) * 2 * C, C, c8 * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
      }
    }
    store8_pair(y + p * 2 * C, C, c8 * 8, m);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, 8 channels (16 B) per lane per pass, fp32 statistics (two-pass in
// registers), optional y2 = y + pos.
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        const __half* __restrict__ pos, __half* __restrict__ y,
                                                        __half* __restrict__ y2, int rows, int C, float eps, int split) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int c = lane * 8;
  const bool act = c < C;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  const int64_t ld = split ? 2 * C : C;   // pair tensors: rows of [hi(C) | lo(C)]
  if (act) {
    if (split) {
      load8_pair(x + warp * ld, C, c, v);
    } else {
      load8(x + warp * ld + c, v);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float sq = 0.f;
  if (act) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = v[i] - mean;
      sq += d * d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  if (act) {
    float o8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o8[i] = (v[i] - mean) * rstd * gamma[c + i] + beta[c + i];
    if (split) {
      store8_pair(y + warp * ld, C, c, o8);
      if (y2 != nullptr) {
        float pv[8];
        load8_pair(pos + warp * ld, C, c, pv);
#pragma unroll
        for (int i = 0; i < 8; ++i) pv[i] += o8[i];
        store8_pair(y2 + warp * ld, C, c, pv);
      }
      return;
    }
    uint4 q;
    q.x = pack_h2(o8[0], o8[1]);
    q.y = pack_h2(o8[2], o8[3]);
    q.z = pack_h2(o8[4], o8[5]);
    q.w = pack_h2(o8[6], o8[7]);
    *reinterpret_cast<uint4*>(y + static_cast<int64_t>(warp) * C + c) = q;
    if (y2 != nullptr) {
      const uint4 pq = *reinterpret_cast<const uint4*>(pos + static_cast<int64_t>(warp) * C + c);
      const uint32_t p4[4] = {pq.x, pq.y, pq.z, pq.w};
      uint32_t r4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_h2(p4[i]);
        r4[i] = pack_h2(o8[2 * i] + f.x, o8[2 * i + 1] + f.y);
      }
      *reinterpret_cast<uint4*>(y2 + static_cast<int64_t>(warp) * C + c) = make_uint4(r4[0], r4[1], r4[2], r4[3]);
    }
  }
}

__global__ void __launch_bounds__(256) add_f16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                      uint4* __restrict__ y, int64_t n8) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 qa = a[i], qb = b[i];
    const __half2* ha = reinterpret_cast<const __half2*>(&qa);
    const __half2* hb = reinterpret_cast<const __half2*>(&qb);
    __half2 r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = __hadd2(ha[k], hb[k]);
    y[i] = *reinterpret_cast<uint4*>(r);
  }
}

// pair tensors: rows of [hi(C) | lo(C)]; y = (a_hi + a_lo) + (b_hi + b_lo), re-split
__global__ void __launch_bounds__(256) add_pair_kernel(const __half* __restrict__ a, const __half* __restrict__ b,
                                                       __half* __restrict__ y, int64_t rows, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = C >> 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < rows * cv;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cv;
    const int c = static_cast<int>(i - r * cv) * 8;
    float va[8], vb[8];
    load8_pair(a + r * 2 * C, C, c, va);
    load8_pair(b + r * 2 * C, C, c, vb);
#pragma unroll
    for (int k = 0; k < 8; ++k) va[k] += vb[k];
    store8_pair(y + r * 2 * C, C, c, va);
  }
}

// y[n,h,w,:] = act(x0[n,h,w,:] + t1[n,h>>s1,w>>s1,:] + t2[n,h>>s2,w>>s2,:]) -- the highest-resolution output of an
// HRNet fuse layer: identity branch + nearest-upsampled 1x1 terms computed at their own (lower) resolution, because
// a 1x1 convolution commutes with nearest upsampling (interformer_pureMulti.py:392-410).  PAIR: pair tensors.
template <bool PAIR>
__global__ void __launch_bounds__(256) upsum_kernel(const __half* __restrict__ x0, const __half* __restrict__ t1,
                                                    const __half* __restrict__ t2, __half* __restrict__ y, int NB,
                                                    int H, int W, int C, int s1, int s2, int relu) {
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
    float v[8], a[8];
    if (PAIR) load8_pair(x0 + p * ld, C, c, v); else load8(x0 + p * ld + c, v);
    const int64_t p1 = (static_cast<int64_t>(n) * (H >> s1) + (hy >> s1)) * (W >> s1) + (wx >> s1);
    if (PAIR) load8_pair(t1 + p1 * ld, C, c, a); else load8(t1 + p1 * ld + c, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += a[i];
    if (t2 != nullptr) {
      const int64_t p2 = (static_cast<int64_t>(n) * (H >> s2) + (hy >> s2)) * (W >> s2) + (wx >> s2);
      if (PAIR) load8_pair(t2 + p2 * ld, C, c, a); else load8(t2 + p2 * ld + c, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += a[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (PAIR) store8_pair(y + p * ld, C, c, v); else store8(y + p * ld + c, v);
  }
}

static int grid_for(int64_t work_items, int block) {
  int64_t g = (work_items + block - 1) / block;
  const int64_t cap = 148 * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace i2r

using namespace i2r;

extern "C" int i2r_version(void) { return I2R_ABI_VERSION; }
extern "C" const char* i2r_last_error(void) { return g_err; }
extern "C" int i2r_sizeof_conv_problem(void) { return static_cast<int>(sizeof(i2r_conv_problem)); }

extern "C" int i2r_device_check(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    return I2R_E_DEVICE;
  }
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libi2r_sm100 contains sm_100a code only", dev, prop.major, prop.minor);
    return I2R_E_DEVICE;
  }
  return 0;
}

extern "C" int i2r_sm_count(int dev) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return I2R_E_DEVICE;
  return n;
}

extern "C" int i2r_stem_conv3x3s2(const float* x, const float* w, const float* scale, const float* bias, void* y,
                                  int NB, int Cin, int H, int W, int Cout, int split, void* stream) {
  if (!x || !w || !scale || !bias || !y || NB <= 0 || (H & 1) || (W & 1) || Cout != 64) {
    set_error("i2r_stem_conv3x3s2: bad arguments (Cin=%d H=%d W=%d Cout=%d; Cout must be 64)", Cin, H, W, Cout);
    return I2R_E_BADARG;
  }
  const dim3 grid((W / 2 + STEM_TW - 1) / STEM_TW, (H / 2 + STEM_TH - 1) / STEM_TH, NB);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Cin == 3) {
    launch_pdl(stem_conv_kernel<3>, dim3(grid), dim3(256), 0, st, x, w, scale, bias, static_cast<__half*>(y), NB, H, W, split);
  } else if (Cin == 1) {
    launch_pdl(stem_conv_kernel<1>, dim3(grid), dim3(256), 0, st, x, w, scale, bias, static_cast<__half*>(y), NB, H, W, split);
  } else {
    set_error("i2r_stem_conv3x3s2: Cin=%d unsupported (1 or 3)", Cin);
    return I2R_E_UNSUPPORTED;
  }
  return check_launch("stem_conv_kernel");
}

extern "C" int i2r_maxpool3x3s2(const void* x, void* y, int NB, int H, int W, int C, int split, void* stream) {
  if (!x || !y || NB <= 0 || H <= 0 || W <= 0 || C % 8 != 0) {
    set_error("i2r_maxpool3x3s2: bad arguments");
    return I2R_E_BADARG;
  }
  const int64_t items = static_cast<int64_t>(NB) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  if (split) {
    launch_pdl(maxpool3x3s2_pair_kernel, dim3(grid_for(items, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
        static_cast<const __half*>(x), static_cast<__half*>(y), NB, H, W, C);
  } else {
    launch_pdl(maxpool3x3s2_kernel, dim3(grid_for(items, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
        static_cast<const __half*>(x), static_cast<__half*>(y), NB, H, W, C);
  }
  return check_launch("maxpool3x3s2_kernel");
}

extern "C" int i2r_layernorm(const void* x, const float* gamma, const float* beta, const void* pos, void* y,
                             void* y2, int rows, int C, float eps, int split, void* stream) {
  if (!x || !gamma || !beta || !y || rows <= 0 || C % 8 != 0 || C > 256 || (y2 && !pos)) {
    set_error("i2r_layernorm: bad arguments (rows=%d C=%d)", rows, C);
    return I2R_E_BADARG;
  }
  const int wpb = 8;
  launch_pdl(layernorm_kernel, dim3((rows + wpb - 1) / wpb), dim3(wpb * 32), 0, static_cast<cudaStream_t>(stream),
      static_cast<const __half*>(x), gamma, beta, static_cast<const __half*>(pos), static_cast<__half*>(y),
      static_cast<__half*>(y2), rows, C, eps, split);
  return check_launch("layernorm_kernel");
}

extern "C" int i2r_add_f16(const void* a, const void* b, void* y, int64_t n, int split_c, void* stream) {
  if (!a || !b || !y || n <= 0 || n % 8 != 0) {
    set_error("i2r_add_f16: bad arguments");
    return I2R_E_BADARG;
  }
  if (split_c > 0) {
    if (split_c % 8 != 0 || n % (2 * split_c) != 0) {
      set_error("i2r_add_f16: pair tensors need C %% 8 == 0 and n a multiple of 2*C");
      return I2R_E_BADARG;
    }
    launch_pdl(add_pair_kernel, dim3(grid_for(n / 16, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
        static_cast<const __half*>(a), static_cast<const __half*>(b), static_cast<__half*>(y), n / (2 * split_c), split_c);
    return check_launch("add_pair_kernel");
  }
  launch_pdl(add_f16_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(y), n / 8);
  return check_launch("add_f16_kernel");
}

extern "C" int i2r_upsum(const void* x0, const void* t1, int shift1, const void* t2, int shift2, void* y, int NB, int H,
                         int W, int C, int relu, int split, void* stream) {
  if (!x0 || !t1 || !y || NB <= 0 || H <= 0 || W <= 0 || C % 8 != 0 || shift1 < 0 || shift2 < 0 ||
      (H >> shift1) << shift1 != H || (W >> shift1) << shift1 != W ||
      (t2 && ((H >> shift2) << shift2 != H || (W >> shift2) << shift2 != W))) {
    set_error("i2r_upsum: bad arguments (H=%d W=%d C=%d shifts %d %d)", H, W, C, shift1, shift2);
    return I2R_E_BADARG;
  }
  const int64_t items = static_cast<int64_t>(NB) * H * W * (C / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (split) {
    launch_pdl(upsum_kernel<true>, dim3(grid_for(items, 256)), dim3(256), 0, st, static_cast<const __half*>(x0), static_cast<const __half*>(t1),
                                                             static_cast<const __half*>(t2), static_cast<__half*>(y), NB,
                                                             H, W, C, shift1, shift2, relu);
  } else {
    launch_pdl(upsum_kernel<false>, dim3(grid_for(items, 256)), dim3(256), 0, st, static_cast<const __half*>(x0), static_cast<const __half*>(t1),
                                                              static_cast<const __half*>(t2), static_cast<__half*>(y), NB,
                                                              H, W, C, shift1, shift2, relu);
  }
  return check_launch("upsum_kernel");
}

// Debug aid: install (or remove, with nullptr) the host-mapped record buffer the bounded mbarrier waits report to before
// they trap (i2r_common.cuh: mbar_timeout).  4096 records of four 64-bit words; synchronises the device.
extern "C" int i2r_debug_hang_buffer(void* host_mapped) {
  using namespace i2r;
  hang_sink_attention_tc(host_mapped);
  hang_sink_conv_halo(host_mapped);
  hang_sink_encoder_tail(host_mapped);
  hang_sink_igemm_tc(host_mapped);
  hang_sink_stem_tc(host_mapped);
  hang_sink_window_attention_tc(host_mapped);
  return check_launch("i2r_debug_hang_buffer");
}

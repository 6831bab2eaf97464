// Stem / mask convolution on the tensor cores: 3x3 stride-2 pad-1, Cin in {1, 3}, 64 output channels, fp32 NCHW in,
// folded BatchNorm + ReLU, fp16 NHWC out (pair tensor in split-operand mode).  Replaces conv1/bn1/relu
// (lib/models/interformer_pureMulti.py:677-679) and position_embedding conv1/bn1/relu (position_embedding.py:108-110).
//
// The SIMT version of this layer (misc_kernels.cu, kept as the check implementation) spends 27 FMAs per output value
// and ran at ~9x the HBM time of the layer.  Here a tile of 128 output pixels (4 rows x 32 columns) is one implicit
// GEMM [128 x K] x [K x 64] with K = 9*Cin <= 27 padded to 32: every thread gathers the 27 inputs of its pixel from a
// shared-memory patch and writes them as ONE 128-byte SWIZZLE_128B operand row [x_hi (32 slots) | x_lo (32 slots)]
// (the fp32 input as an fp16 pair, value = hi + lo), and six tcgen05.mma (M128 x N64 x K16) evaluate
// x_hi W_hi + x_lo W_hi + x_hi W_lo into a TMEM accumulator -- the input and the weights keep ~22 bits in both
// precision modes, so the layer is as exact as the fp32 SIMT kernel it replaces.  The epilogue (thread = pixel) adds
// the bias, applies ReLU and writes the pixel's 128-byte (or 256-byte pair) NHWC row.  HBM-bound by design: one read
// of the fp32 image, one write of the fp16 feature map; several CTAs per SM overlap the phases of different tiles.
#include "i2r_tma.cuh"

namespace i2r {

constexpr int ST_TH = 4, ST_TW = 32;                       // output tile
constexpr int ST_PH = 2 * ST_TH + 1, ST_PW = 2 * ST_TW + 1, ST_PWP = ST_PW + 3;   // input patch (+ pad)
constexpr int ST_WIMG_BYTES = 2 * 64 * 128;                // [W_hi | W_hi] rows, then [W_lo | 0] rows

template <int CIN>
__global__ void __launch_bounds__(128) stem_tc_kernel(const float* __restrict__ x, const __half* __restrict__ wimg,
                                                      const float* __restrict__ bias, __half* __restrict__ y, int NB,
                                                      int H, int W, int split, int tiles_x, int tiles_y) {
  constexpr int K = CIN * 9;
  __shared__ __align__(1024) uint8_t sA_raw[128 * 128];
  __shared__ __align__(1024) uint8_t sB_raw[ST_WIMG_BYTES];
  __shared__ float patch[CIN][ST_PH][ST_PWP];
  __shared__ __align__(16) float sbias[64];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sA = smem_u32(sA_raw), sB = smem_u32(sB_raw);
  const uint32_t bW = smem_u32(&bars[0]), bM = smem_u32(&bars[1]);
  if (tid == 0) {
    mbar_init(bW, 1);
    mbar_init(bM, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(bW, ST_WIMG_BYTES);                 // weights: constants, loaded before the dependency wait
    bulk_g2s(sB, wimg, ST_WIMG_BYTES, bW);
  }
  if (tid < 64) sbias[tid] = bias[tid];
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();
  const int OH = H >> 1, OW = W >> 1;
  const int ntiles = NB * tiles_y * tiles_x;
  const uint32_t idesc = make_idesc_f16(128, 64);
  const uint32_t dhi = sw128_desc_hi(1024, 0);
  const int ty = tid >> 5, tx = tid & 31;
  uint32_t phase = 0;
  bool first = true;
  constexpr int TOTAL = CIN * ST_PH * ST_PW;
  constexpr int NLD = (TOTAL + 127) / 128;
  // input patch of one tile, coalesced along the image rows, as NLD register values per thread.  The loads of tile i+1
  // are issued right after the patch of tile i has been stored to shared memory, so their (cold, first-touch HBM)
  // latency hides behind the MMAs and the epilogue of tile i.
  auto load_patch = [&](int tile, float (&v)[NLD]) {
    const int n = tile / (tiles_y * tiles_x);
    const int rem = tile - n * (tiles_y * tiles_x);
    const int iy0 = (rem / tiles_x) * ST_TH * 2 - 1, ix0 = (rem % tiles_x) * ST_TW * 2 - 1;
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + u * 128;
      const int c = i / (ST_PH * ST_PW);
      const int r = i - c * (ST_PH * ST_PW);
      const int py = r / ST_PW, px = r - py * ST_PW;
      const int iy = iy0 + py, ix = ix0 + px;
      v[u] = 0.f;
      if (i < TOTAL && tile < ntiles && iy >= 0 && iy < H && ix >= 0 && ix < W)
        v[u] = __ldg(x + ((static_cast<int64_t>(n) * CIN + c) * H + iy) * W + ix);
    }
  };
  float pv[NLD];
  load_patch(blockIdx.x, pv);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_y * tiles_x);
    const int rem = tile - n * (tiles_y * tiles_x);
    const int oy0 = (rem / tiles_x) * ST_TH, ox0 = (rem % tiles_x) * ST_TW;
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + u * 128;
      const int c = i / (ST_PH * ST_PW);
      const int r = i - c * (ST_PH * ST_PW);
      const int py = r / ST_PW, px = r - py * ST_PW;
      if (i < TOTAL) patch[c][py][px] = pv[u];
    }
    load_patch(tile + gridDim.x, pv);      // prefetch (no-op past the last tile)
    __syncthreads();
    // ---- this pixel's operand row: slots [0, K) = hi, [32, 32 + K) = lo, the rest zero
    {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float v0 = 0.f, v1 = 0.f;
        if (2 * j < K) {
          const int k = 2 * j, c = k / 9, ky = (k % 9) / 3, kx = k % 3;
          v0 = patch[c][2 * ty + ky][2 * tx + kx];
        }
        if (2 * j + 1 < K) {
          const int k = 2 * j + 1, c = k / 9, ky = (k % 9) / 3, kx = k % 3;
          v1 = patch[c][2 * ty + ky][2 * tx + kx];
        }
        hi[j] = pack_h2(v0, v1);
        const float2 f = unpack_h2(hi[j]);
        lo[j] = pack_h2(v0 - f.x, v1 - f.y);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sA + sw128_off(tid, g)), "r"(hi[4 * g]),
                     "r"(hi[4 * g + 1]), "r"(hi[4 * g + 2]), "r"(hi[4 * g + 3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sA + sw128_off(tid, 4 + g)), "r"(lo[4 * g]),
                     "r"(lo[4 * g + 1]), "r"(lo[4 * g + 2]), "r"(lo[4 * g + 3])
                     : "memory");
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      if (first) mbar_wait(bW, 0);
      tc_fence_after();
      const uint32_t a0 = sw128_desc_lo(sA), b1 = sw128_desc_lo(sB), b2 = sw128_desc_lo(sB + 64 * 128);
#pragma unroll
      for (int s = 0; s < 4; ++s)      // [x_hi | x_lo] . [W_hi | W_hi]
        umma_f16(tmem, desc64(a0 + 2 * s, dhi), desc64(b1 + 2 * s, dhi), idesc, s ? 1u : 0u);
#pragma unroll
      for (int s = 0; s < 2; ++s)      // x_hi . W_lo
        umma_f16(tmem, desc64(a0 + 2 * s, dhi), desc64(b2 + 2 * s, dhi), idesc, 1u);
      umma_commit(bM);
    }
    first = false;
    mbar_wait(bM, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: bias, ReLU, NHWC row of this pixel
    const int oy = oy0 + ty, ox = ox0 + tx;
    const bool ok = oy < OH && ox < OW;
    __half* row = y + ((static_cast<int64_t>(n) * OH + oy) * OW + ox) * (split ? 128 : 64);
    const uint32_t tacc = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[16];
      tmem_ld16(tacc + c * 16, r);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t p[4], q[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = fmaxf(__uint_as_float(r[h * 8 + 2 * i]) + sbias[c * 16 + h * 8 + 2 * i], 0.f);
            const float b = fmaxf(__uint_as_float(r[h * 8 + 2 * i + 1]) + sbias[c * 16 + h * 8 + 2 * i + 1], 0.f);
            p[i] = pack_h2(a, b);
            const float2 f = unpack_h2(p[i]);
            q[i] = pack_h2(a - f.x, b - f.y);
          }
          *reinterpret_cast<uint4*>(row + c * 16 + h * 8) = make_uint4(p[0], p[1], p[2], p[3]);
          if (split) *reinterpret_cast<uint4*>(row + 64 + c * 16 + h * 8) = make_uint4(q[0], q[1], q[2], q[3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();     // accumulator, operand rows and patch are free for the next tile
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

I2R_HANG_SINK_SETTER(stem_tc)
}  // namespace i2r

extern "C" int64_t i2r_stem_tc_weight_bytes(void) { return i2r::ST_WIMG_BYTES; }

extern "C" int i2r_stem_conv3x3s2_tc(const float* x, const void* wimg, const float* bias, void* y, int NB, int Cin,
                                     int H, int W, int Cout, int split, void* stream) {
  using namespace i2r;
  if (!x || !wimg || !bias || !y || NB <= 0 || (H & 1) || (W & 1) || Cout != 64 ||
      (reinterpret_cast<uintptr_t>(wimg) & 15) != 0) {
    set_error("i2r_stem_conv3x3s2_tc: bad arguments (Cin=%d H=%d W=%d Cout=%d; Cout must be 64)", Cin, H, W, Cout);
    return I2R_E_BADARG;
  }
  const int tiles_x = (W / 2 + ST_TW - 1) / ST_TW, tiles_y = (H / 2 + ST_TH - 1) / ST_TH;
  const int64_t ntiles = static_cast<int64_t>(NB) * tiles_x * tiles_y;
  const int grid = static_cast<int>(ntiles < 4 * device_sms() ? ntiles : 4 * device_sms());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Cin == 3) {
    launch_pdl(stem_tc_kernel<3>, dim3(grid), dim3(128), 0, st, x, static_cast<const __half*>(wimg), bias,
               static_cast<__half*>(y), NB, H, W, split, tiles_x, tiles_y);
  } else if (Cin == 1) {
    launch_pdl(stem_tc_kernel<1>, dim3(grid), dim3(128), 0, st, x, static_cast<const __half*>(wimg), bias,
               static_cast<__half*>(y), NB, H, W, split, tiles_x, tiles_y);
  } else {
    set_error("i2r_stem_conv3x3s2_tc: Cin=%d unsupported (1 or 3)", Cin);
    return I2R_E_UNSUPPORTED;
  }
  return check_launch("stem_tc_kernel");
}

// `res` multi-person position embedding, stem part (lib/models/position_embedding.py:14-18, :90-93; the shipped
// experiments/OCHuman/interformer_ochuman_tph_192_p3_b8.yaml uses it): per-person box mask [S,1,H,W] fp32 ->
//   conv_pre  3x3 s1 p1, 1 -> 3, no bias            (position_embedding.py:15)
//   resnet18 conv1 7x7 s2 p3, 3 -> 64, no bias + bn1 + ReLU   (torchvision resnet18 children[0:3])
// as ONE kernel writing fp16 NHWC [S, H/2, W/2, 64] (pair tensor in split-operand mode).  The two convolutions are NOT
// composed into one 9x9 filter: conv1's zero padding applies to conv_pre's OUTPUT, whose border values are non-zero.
// SIMT fp32 (K = 147 per output with 3 input channels is no tensor-core shape); the max-pool, the two BasicBlocks of
// resnet18.layer1 and conv_end that follow run on the generic kernels.
#include "i2r_common.cuh"

namespace i2r {

constexpr int MR_TH = 8, MR_TW = 16;                         // output tile
constexpr int MR_PH = 2 * MR_TH + 5, MR_PW = 2 * MR_TW + 5;  // conv_pre patch 21 x 37
constexpr int MR_MH = MR_PH + 2, MR_MW = MR_PW + 2;          // mask patch 23 x 39
constexpr int MR_K = 147;

template <bool PAIR>
__global__ void __launch_bounds__(256) mask_res_stem_kernel(const float* __restrict__ mask, const float* __restrict__ w_pre,
                                                            const float* __restrict__ w1, const float* __restrict__ scale,
                                                            const float* __restrict__ bias, __half* __restrict__ y, int NB,
                                                            int H, int W) {
  extern __shared__ __align__(16) float mr_smem[];
  float* sw = mr_smem;                          // [147][64]
  float* spre = sw + MR_K * 64;                 // [21][37][3]
  float* smask = spre + MR_PH * MR_PW * 3;      // [23][39]
  pdl_launch_dependents();
  const int OH = H >> 1, OW = W >> 1;
  const int tiles_x = (OW + MR_TW - 1) / MR_TW, tiles_y = (OH + MR_TH - 1) / MR_TH;
  const int tid = threadIdx.x;
  for (int i = tid; i < MR_K * 64; i += 256) sw[i] = __ldg(w1 + i);
  float wp[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) wp[i] = __ldg(w_pre + i);
  pdl_wait();
  const int ntiles = NB * tiles_x * tiles_y;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int r = tile - n * tiles_x * tiles_y;
    const int oy0 = (r / tiles_x) * MR_TH, ox0 = (r % tiles_x) * MR_TW;
    const int py0 = 2 * oy0 - 3, px0 = 2 * ox0 - 3;          // conv_pre patch origin (input resolution)
    __syncthreads();
    for (int i = tid; i < MR_MH * MR_MW; i += 256) {
      const int my = py0 - 1 + i / MR_MW, mx = px0 - 1 + i % MR_MW;
      smask[i] = (my >= 0 && my < H && mx >= 0 && mx < W) ? __ldg(mask + (static_cast<int64_t>(n) * H + my) * W + mx) : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < MR_PH * MR_PW; i += 256) {
      const int ly = i / MR_PW, lx = i % MR_PW;
      const int gy = py0 + ly, gx = px0 + lx;
      float a[3] = {0.f, 0.f, 0.f};
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {      // outside the image conv1 sees its own zero padding
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float m = smask[(ly + ky) * MR_MW + lx + kx];
#pragma unroll
            for (int c = 0; c < 3; ++c) a[c] = fmaf(m, wp[c * 9 + ky * 3 + kx], a[c]);
          }
      }
      spre[i * 3] = a[0];
      spre[i * 3 + 1] = a[1];
      spre[i * 3 + 2] = a[2];
    }
    __syncthreads();
    // thread = (pixel of the 8 x 16 tile, 32-channel half)
    const int p = tid & 127, ch0 = (tid >> 7) * 32;
    const int ty = p / MR_TW, tx = p % MR_TW;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int c = 0; c < 3; ++c)
      for (int ky = 0; ky < 7; ++ky)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float v = spre[((2 * ty + ky) * MR_PW + 2 * tx + kx) * 3 + c];
          const float4* wr = reinterpret_cast<const float4*>(sw + ((c * 7 + ky) * 7 + kx) * 64 + ch0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 w4 = wr[q];
            acc[4 * q] = fmaf(v, w4.x, acc[4 * q]);
            acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
          }
        }
    const int oy = oy0 + ty, ox = ox0 + tx;
    if (oy < OH && ox < OW) {
      __half* row = y + ((static_cast<int64_t>(n) * OH + oy) * OW + ox) * (PAIR ? 128 : 64);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = ch0 + 8 * g + 2 * i;
          const float a = fmaxf(fmaf(acc[8 * g + 2 * i], __ldg(scale + c), __ldg(bias + c)), 0.f);
          const float b = fmaxf(fmaf(acc[8 * g + 2 * i + 1], __ldg(scale + c + 1), __ldg(bias + c + 1)), 0.f);
          h[i] = pack_h2(a, b);
          const float2 f = unpack_h2(h[i]);
          l[i] = pack_h2(a - f.x, b - f.y);
        }
        *reinterpret_cast<uint4*>(row + ch0 + 8 * g) = make_uint4(h[0], h[1], h[2], h[3]);
        if (PAIR) *reinterpret_cast<uint4*>(row + 64 + ch0 + 8 * g) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

}  // namespace i2r

using namespace i2r;

extern "C" int i2r_mask_res_stem(const float* mask, const float* w_pre, const float* w1, const float* scale,
                                 const float* bias, void* y, int NB, int H, int W, int split, void* stream) {
  if (!mask || !w_pre || !w1 || !scale || !bias || !y || NB <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1)) {
    set_error("i2r_mask_res_stem: bad arguments (even H, W)");
    return I2R_E_BADARG;
  }
  const size_t smem = sizeof(float) * (MR_K * 64 + MR_PH * MR_PW * 3 + MR_MH * MR_MW);
  static bool attr_done_dev[MAX_DEVICES][2] = {};
  bool& attr_done = attr_done_dev[current_device()][split ? 1 : 0];
  if (!attr_done) {
    cudaError_t e = split ? cudaFuncSetAttribute(mask_res_stem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem))
                          : cudaFuncSetAttribute(mask_res_stem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(mask_res_stem): %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_done = true;
  }
  const int OH = H / 2, OW = W / 2;
  const int ntiles = NB * ((OW + MR_TW - 1) / MR_TW) * ((OH + MR_TH - 1) / MR_TH);
  const int grid = ntiles < 148 * 4 ? ntiles : 148 * 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (split)
    launch_pdl(mask_res_stem_kernel<true>, dim3(grid), dim3(256), smem, st, mask, w_pre, w1, scale, bias,
               static_cast<__half*>(y), NB, H, W);
  else
    launch_pdl(mask_res_stem_kernel<false>, dim3(grid), dim3(256), smem, st, mask, w_pre, w1, scale, bias,
               static_cast<__half*>(y), NB, H, W);
  return check_launch("mask_res_stem_kernel");
}

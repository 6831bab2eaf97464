// Test-time input pipeline on the device: person crops by affine warp + normalisation and the
// per-person box masks, i.e. lib/dataset/JointsDataset.py:296-331 (`__getitem__`, is_train False) with its OpenCV calls
// restated in their own 8-bit fixed-point arithmetic, so that the tensors handed to the forward are the ones the
// reference's DataLoader would have produced (x bit-exact; pos_mask up to one grey level, see oracle/preproc_oracle.py).
// HBM-bound SIMT kernels: one thread per output pixel.
#include "i2r_common.cuh"

namespace i2r {

struct CropArgs {
  const uint8_t* img;      // [IH, IW, 3] RGB
  const double* inv;       // [N, 6] inverse affine (dst -> src), as cv2.warpAffine derives it from the forward matrix
  float* x;                // [N, 3, OH, OW]
  int IH, IW, N, OH, OW;
  float mean[3], stdv[3];
};

// cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) + ToTensor + Normalize (tools/test.py:126-134)
__global__ void __launch_bounds__(256) crop_warp_norm_kernel(const __grid_constant__ CropArgs A) {
  const int64_t total = static_cast<int64_t>(A.N) * A.OH * A.OW;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / (A.OH * A.OW));
    const int r = static_cast<int>(i - static_cast<int64_t>(n) * A.OH * A.OW);
    const int y = r / A.OW, x = r - y * A.OW;
    const double* m = A.inv + 6 * n;
    // 1/1024 fixed point, rounded to 1/32 pixel; the products are kept un-fused (cv2 evaluates them in plain double)
    const long long adelta = __double2ll_rn(__dmul_rn(__dmul_rn(m[0], static_cast<double>(x)), 1024.0));
    const long long bdelta = __double2ll_rn(__dmul_rn(__dmul_rn(m[3], static_cast<double>(x)), 1024.0));
    const long long x0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], static_cast<double>(y)), m[2]), 1024.0)) + 16;
    const long long y0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], static_cast<double>(y)), m[5]), 1024.0)) + 16;
    const long long xq = (x0 + adelta) >> 5, yq = (y0 + bdelta) >> 5;
    long long sx = xq >> 5, sy = yq >> 5;
    sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);
    sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
    const int fx = static_cast<int>(xq & 31), fy = static_cast<int>(yq & 31);
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    int acc[3] = {0, 0, 0};
    auto tap = [&](long long yy, long long xx, int w) {
      if (w != 0 && yy >= 0 && yy < A.IH && xx >= 0 && xx < A.IW) {
        const uint8_t* p = A.img + (yy * A.IW + xx) * 3;
        acc[0] += w * p[0];
        acc[1] += w * p[1];
        acc[2] += w * p[2];
      }
    };
    tap(sy, sx, w00);
    tap(sy, sx + 1, w01);
    tap(sy + 1, sx, w10);
    tap(sy + 1, sx + 1, w11);
    float* out = A.x + (static_cast<int64_t>(n) * 3) * A.OH * A.OW + r;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int v = (acc[c] + (1 << 14)) >> 15;
      // ToTensor: v / 255 ; Normalize: (t - mean) / std -- three IEEE float32 operations, as torch evaluates them
      out[static_cast<int64_t>(c) * A.OH * A.OW] =
          __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), A.mean[c]), A.stdv[c]);
    }
  }
}

struct MaskArgs {
  const int32_t* rect;     // [N, 4] inclusive x0, y0, x1, y1 (cv2.rectangle corners, already ordered)
  float* pm;               // [N, 1, OH, OW]
  int IH, IW, N, OH, OW;
};

// one axis of cv2.resize(INTER_LINEAR) on uint8: source offset + the two 11-bit coefficients of destination index d
__device__ __forceinline__ void resize_coeff(int d, int dn, int sn, int& ofs, int& c0, int& c1) {
  const double scale = 1.0 / (static_cast<double>(dn) / static_cast<double>(sn));
  float f = static_cast<float>(__dsub_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), 0.5));
  int s = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(s));
  if (s < 0) {
    s = 0;
    f = 0.f;
  }
  if (s >= sn - 1) {
    s = sn - 1;
    f = 0.f;
  }
  ofs = s;
  c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
  c1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

// get_position (box rectangle, JointsDataset.py:166-177) -> rotate_bound(angle 0) (:179-201: a half-pixel bilinear shift
// along every odd image dimension) -> cv2.resize to the network input size -> ToTensor
__global__ void __launch_bounds__(256) box_mask_kernel(const __grid_constant__ MaskArgs A) {
  const int64_t total = static_cast<int64_t>(A.N) * A.OH * A.OW;
  const int hx = A.IW & 1, hy = A.IH & 1;      // half-pixel shifts of rotate_bound
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / (A.OH * A.OW));
    const int r = static_cast<int>(i - static_cast<int64_t>(n) * A.OH * A.OW);
    const int y = r / A.OW, x = r - y * A.OW;
    const int rx0 = A.rect[4 * n], ry0 = A.rect[4 * n + 1], rx1 = A.rect[4 * n + 2], ry1 = A.rect[4 * n + 3];
    auto rect = [&](int yy, int xx) -> int {
      return (yy >= 0 && yy < A.IH && xx >= 0 && xx < A.IW && yy >= ry0 && yy <= ry1 && xx >= rx0 && xx <= rx1) ? 255 : 0;
    };
    // rotate_bound(0): warpAffine with translation (hx/2, hy/2): source = (x - 1 | x) with weights 16/32 when shifted
    auto rot = [&](int yy, int xx) -> int {
      const int fx = hx ? 16 : 0, fy = hy ? 16 : 0;
      const int sx = xx - hx, sy = yy - hy;
      const int acc = rect(sy, sx) * ((32 - fx) * (32 - fy) * 32) + rect(sy, sx + 1) * (fx * (32 - fy) * 32) +
                      rect(sy + 1, sx) * ((32 - fx) * fy * 32) + rect(sy + 1, sx + 1) * (fx * fy * 32);
      return (acc + (1 << 14)) >> 15;
    };
    int xo, xa0, xa1, yo, ya0, ya1;
    resize_coeff(x, A.OW, A.IW, xo, xa0, xa1);
    resize_coeff(y, A.OH, A.IH, yo, ya0, ya1);
    const int x1 = min(xo + 1, A.IW - 1), y1 = min(yo + 1, A.IH - 1);
    const int r0 = rot(yo, xo) * xa0 + rot(yo, x1) * xa1;
    const int r1 = rot(y1, xo) * xa0 + rot(y1, x1) * xa1;
    const int v = (((ya0 * (r0 >> 4)) >> 16) + ((ya1 * (r1 >> 4)) >> 16) + 2) >> 2;
    A.pm[i] = __fdiv_rn(static_cast<float>(v & 255), 255.0f);
  }
}

static int pre_grid(int64_t items) {
  int64_t g = (items + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace i2r

using namespace i2r;

extern "C" int i2r_crop_persons(const uint8_t* image, int IH, int IW, const double* inv_affine, int N, int OH, int OW,
                                const float* mean3, const float* std3, float* x, void* stream) {
  if (!image || !inv_affine || !x || !mean3 || !std3 || IH <= 0 || IW <= 0 || N <= 0 || OH <= 0 || OW <= 0) {
    set_error("i2r_crop_persons: bad arguments");
    return I2R_E_BADARG;
  }
  CropArgs A;
  A.img = image; A.inv = inv_affine; A.x = x;
  A.IH = IH; A.IW = IW; A.N = N; A.OH = OH; A.OW = OW;
  for (int c = 0; c < 3; ++c) {
    A.mean[c] = mean3[c];
    A.stdv[c] = std3[c];
  }
  crop_warp_norm_kernel<<<pre_grid(static_cast<int64_t>(N) * OH * OW), 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  return check_launch("crop_warp_norm_kernel");
}

extern "C" int i2r_box_masks(const int32_t* rect, int N, int IH, int IW, int OH, int OW, float* pos_mask, void* stream) {
  if (!rect || !pos_mask || IH <= 1 || IW <= 1 || N <= 0 || OH <= 0 || OW <= 0) {
    set_error("i2r_box_masks: bad arguments");
    return I2R_E_BADARG;
  }
  MaskArgs A;
  A.rect = rect; A.pm = pos_mask; A.IH = IH; A.IW = IW; A.N = N; A.OH = OH; A.OW = OW;
  box_mask_kernel<<<pre_grid(static_cast<int64_t>(N) * OH * OW), 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  return check_launch("box_mask_kernel");
}

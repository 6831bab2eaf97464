// Post-processing of the heatmaps on the device -- HBM-bound SIMT kernels:
//   * flip test (lib/core/function.py:142-162, lib/utils/transforms.py:16-30): horizontal flip of the inputs and
//     `(out + flip_back(out_flipped)) * 0.5` with the left/right joint swap, without the numpy round trip;
//   * heatmap decode (lib/core/inference.py:20-112): arg-max, DARK refinement (zero-padded Gaussian blur,
//     renormalisation to the original maximum, log, second-order Taylor step) and the inverse crop transform
//     (lib/utils/transforms.py:50-92) -- the reference runs per-joint python loops on the host over [S, K, 64, 48].
#include <cmath>

#include "i2r_common.cuh"

namespace i2r {

__global__ void hflip_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int W) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / W;
    const int w = static_cast<int>(i - r * W);
    dst[i] = __ldg(src + r * W + (W - 1 - w));
  }
}

// y[s,k,h,w] = 0.5 * (a[s,k,h,w] + b[s,perm[k],h,W-1-w])
__global__ void flip_merge_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                  int64_t n, int K, int HW, int W, const int32_t* __restrict__ perm) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t plane = i / HW;
    const int p = static_cast<int>(i - plane * HW);
    const int64_t s = plane / K;
    const int k = static_cast<int>(plane - s * K);
    const int h = p / W, w = p - h * W;
    const float f = __ldg(b + (s * K + __ldg(perm + k)) * HW + h * W + (W - 1 - w));
    y[i] = (__ldg(a + i) + f) * 0.5f;
  }
}

// ------------------------------------------------------------------------------------------------ decode
constexpr int DEC_THREADS = 256;
constexpr int DEC_MAX_KSIZE = 31;

struct DecodeArgs {
  const float* hm;       // [S, K, H, W]
  const float* center;   // [S, 2]
  const float* scale;    // [S, 2]
  float* preds;          // [S, K, 2]
  float* maxvals;        // [S, K]
  int K, H, W, ksize, transform_back;
  double kern[DEC_MAX_KSIZE];
};

// One CTA per heatmap.  Shared memory: the map (float), the row-filtered map (double, cv2 filters CV_64F data in
// double), the blurred map (float: the reference stores the blurred values back into its float32 array).
__global__ void __launch_bounds__(DEC_THREADS) decode_kernel(const __grid_constant__ DecodeArgs A) {
  extern __shared__ __align__(16) uint8_t dec_smem[];
  const int HW = A.H * A.W, W = A.W, H = A.H;
  double* rowf = reinterpret_cast<double*>(dec_smem);
  float* hm = reinterpret_cast<float*>(dec_smem + sizeof(double) * HW);
  float* bl = hm + HW;
  __shared__ float red_v[DEC_THREADS / 32];
  __shared__ int red_i[DEC_THREADS / 32];
  __shared__ float s_max, s_bmax;
  __shared__ int s_idx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t map = blockIdx.x;
  const float* src = A.hm + map * HW;
  // ---- load + arg-max (first index of the maximum, as np.argmax)
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = tid; i < HW; i += DEC_THREADS) {
    const float v = __ldg(src + i);
    hm[i] = v;
    if (v > bv) {           // i grows within a thread, so '>' keeps the first index
      bv = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    red_v[warp] = bv;
    red_i[warp] = bi;
  }
  __syncthreads();
  if (tid == 0) {
    float v = red_v[0];
    int idx = red_i[0];
    for (int w = 1; w < DEC_THREADS / 32; ++w)
      if (red_v[w] > v || (red_v[w] == v && red_i[w] < idx)) {
        v = red_v[w];
        idx = red_i[w];
      }
    s_max = v;
    s_idx = idx;
  }
  __syncthreads();
  const float maxval = s_max;
  // ---- zero-padded separable Gaussian blur in double (cv2.GaussianBlur on the float64 padded copy, inference.py:73-87)
  const int r = (A.ksize - 1) / 2;
  for (int i = tid; i < HW; i += DEC_THREADS) {
    const int y = i / W, x = i - y * W;
    double s = 0.0;
    for (int j = 0; j < A.ksize; ++j) {
      const int xx = x + j - r;
      if (xx >= 0 && xx < W) s += A.kern[j] * static_cast<double>(hm[y * W + xx]);
    }
    rowf[i] = s;
  }
  __syncthreads();
  float bmax = -INFINITY;
  for (int i = tid; i < HW; i += DEC_THREADS) {
    const int y = i / W, x = i - y * W;
    double s = 0.0;
    for (int j = 0; j < A.ksize; ++j) {
      const int yy = y + j - r;
      if (yy >= 0 && yy < H) s += A.kern[j] * rowf[yy * W + x];
    }
    const float f = static_cast<float>(s);
    bl[i] = f;
    bmax = fmaxf(bmax, f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
  if (lane == 0) red_v[warp] = bmax;
  __syncthreads();
  if (tid == 0) {
    float v = red_v[0];
    for (int w = 1; w < DEC_THREADS / 32; ++w) v = fmaxf(v, red_v[w]);
    s_bmax = v;
  }
  __syncthreads();
  if (tid != 0) return;
  // ---- coordinates, Taylor refinement on log(max(blurred * origin_max / blurred_max, 1e-10)), inverse transform
  const int idx = s_idx;
  float cx = static_cast<float>(idx % W), cy = floorf(static_cast<float>(idx) / static_cast<float>(W));
  if (!(maxval > 0.0f)) {
    cx = 0.f;
    cy = 0.f;
  }
  const float ratio = maxval / s_bmax;      // float32 scalar, as numpy computes it
  auto L = [&](int yy, int xx) -> float { return logf(fmaxf(bl[yy * W + xx] * ratio, 1e-10f)); };
  const int px = static_cast<int>(cx), py = static_cast<int>(cy);
  double ox = cx, oy = cy;
  if (1 < px && px < W - 2 && 1 < py && py < H - 2) {
    const double dx = 0.5 * static_cast<double>(L(py, px + 1) - L(py, px - 1));
    const double dy = 0.5 * static_cast<double>(L(py + 1, px) - L(py - 1, px));
    const float c2 = 2.0f * L(py, px);
    const double dxx = 0.25 * static_cast<double>(L(py, px + 2) - c2 + L(py, px - 2));
    const double dxy = 0.25 * static_cast<double>(L(py + 1, px + 1) - L(py - 1, px + 1) - L(py + 1, px - 1) + L(py - 1, px - 1));
    const double dyy = 0.25 * static_cast<double>(L(py + 2, px) - c2 + L(py - 2, px));
    const double det = dxx * dyy - dxy * dxy;
    if (det != 0.0) {
      // offset = -H^-1 g ; the reference adds it to the float32 coordinate array
      const double offx = -(dyy * dx - dxy * dy) / det;
      const double offy = -(-dxy * dx + dxx * dy) / det;
      ox = static_cast<double>(static_cast<float>(static_cast<double>(cx) + offx));
      oy = static_cast<double>(static_cast<float>(static_cast<double>(cy) + offy));
    }
  }
  if (A.transform_back) {
    // get_affine_transform(center, scale, 0, [W, H], inv=1) is a similarity: factor (scale_x * 200 - 1) / (W - 1)
    // about the centres (transforms.py:58-92 with rot = 0: only scale[0] enters)
    const int64_t s = map / A.K;
    const double c0 = A.center[2 * s], c1 = A.center[2 * s + 1];
    const double src_w = static_cast<double>(A.scale[2 * s]) * 200.0;
    const double k = (src_w - 1.0) / (static_cast<double>(W) - 1.0);
    ox = c0 + k * (ox - (static_cast<double>(W) - 1.0) * 0.5);
    oy = c1 + k * (oy - (static_cast<double>(H) - 1.0) * 0.5);
  }
  A.preds[2 * map] = static_cast<float>(ox);
  A.preds[2 * map + 1] = static_cast<float>(oy);
  A.maxvals[map] = maxval;
}

static int pp_grid(int64_t items, int block) {
  int64_t g = (items + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace i2r

using namespace i2r;

extern "C" int i2r_hflip_f32(const float* src, float* dst, int64_t rows, int W, void* stream) {
  if (!src || !dst || rows <= 0 || W <= 0 || src == dst) {
    set_error("i2r_hflip_f32: bad arguments (out-of-place, rows > 0, W > 0)");
    return I2R_E_BADARG;
  }
  const int64_t n = rows * W;
  hflip_f32_kernel<<<pp_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, W);
  return check_launch("hflip_f32_kernel");
}

extern "C" int i2r_flip_merge(const float* out, const float* out_flipped, float* y, int S, int K, int H, int W,
                              const int32_t* perm, void* stream) {
  if (!out || !out_flipped || !y || !perm || S <= 0 || K <= 0 || H <= 0 || W <= 0) {
    set_error("i2r_flip_merge: bad arguments");
    return I2R_E_BADARG;
  }
  const int64_t n = static_cast<int64_t>(S) * K * H * W;
  flip_merge_kernel<<<pp_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, out_flipped, y, n, K, H * W, W,
                                                                                     perm);
  return check_launch("flip_merge_kernel");
}

extern "C" int i2r_decode_heatmaps(const float* hm, int S, int K, int H, int W, const float* center, const float* scale,
                                   int blur_kernel, int transform_back, float* preds, float* maxvals, void* stream) {
  if (!hm || !preds || !maxvals || S <= 0 || K <= 0 || H < 5 || W < 5 || blur_kernel < 1 || blur_kernel > DEC_MAX_KSIZE ||
      blur_kernel % 2 == 0 || (transform_back && (!center || !scale))) {
    set_error("i2r_decode_heatmaps: bad arguments (odd blur kernel <= %d, maps >= 5x5)", DEC_MAX_KSIZE);
    return I2R_E_BADARG;
  }
  DecodeArgs A;
  A.hm = hm; A.center = center; A.scale = scale; A.preds = preds; A.maxvals = maxvals;
  A.K = K; A.H = H; A.W = W; A.ksize = blur_kernel; A.transform_back = transform_back;
  // cv2.getGaussianKernel(ksize, sigma <= 0): fixed tables up to 7 taps, else sigma = 0.3*((ksize-1)*0.5 - 1) + 0.8
  static const double small[4][7] = {{1.0}, {0.25, 0.5, 0.25}, {0.0625, 0.25, 0.375, 0.25, 0.0625},
                                     {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125}};
  if (blur_kernel <= 7) {
    for (int i = 0; i < blur_kernel; ++i) A.kern[i] = small[blur_kernel >> 1][i];
  } else {
    const double sigma = 0.3 * ((blur_kernel - 1) * 0.5 - 1.0) + 0.8;
    const double scale2x = -0.5 / (sigma * sigma);
    double sum = 0.0;
    for (int i = 0; i < blur_kernel; ++i) {
      const double x = i - (blur_kernel - 1) * 0.5;
      A.kern[i] = std::exp(scale2x * x * x);
      sum += A.kern[i];
    }
    for (int i = 0; i < blur_kernel; ++i) A.kern[i] *= 1.0 / sum;
  }
  const size_t smem = static_cast<size_t>(H) * W * (sizeof(double) + 2 * sizeof(float));
  if (smem > 200 * 1024) {
    set_error("i2r_decode_heatmaps: %dx%d maps need %zu B of shared memory", H, W, smem);
    return I2R_E_UNSUPPORTED;
  }
  static bool attr_done_dev[MAX_DEVICES] = {};
  bool& attr_done = attr_done_dev[current_device()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(decode_kernel): %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_done = true;
  }
  decode_kernel<<<S * K, DEC_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(A);
  return check_launch("decode_kernel");
}

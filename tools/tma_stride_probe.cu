// Micro-probe: TMA tensor loads with elementStrides = 2 along W and H (every other pixel of an NHWC fp16 tensor), the
// building block of a stride-2 3x3 convolution in the halo formulation (four parity planes of the input tile).
// Questions: (1) how many shared-memory rows does a box of boxDim (64, BW, BH, 1) with elementStrides (1, 2, 2, 1)
// write -- ceil(BW/2) * ceil(BH/2)? -- and in which order; (2) which source pixel lands in which row, including
// negative start coordinates (zero fill); (3) how many bytes complete the mbarrier transaction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/tma_stride_probe tools/tma_stride_probe.cu
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../intra-and-inter-human-relation-network-for-mpee_b200/csrc/i2r_common.cuh"

namespace i2r {
void set_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
}  // namespace i2r
using namespace i2r;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

constexpr int ROWS = 512;   // 128-byte rows of shared memory under observation

// mode 0: huge expect_tx, wait a fixed time, dump (which rows were written); mode 1: expect exactly `tx` bytes and wait
__global__ void probe(const __grid_constant__ CUtensorMap map, int x0, int y0, int mode, uint32_t tx, uint32_t* out,
                      int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + ROWS * 128;
  for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xdeadbeefu;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, mode == 0 ? (1u << 19) : tx);
    tma_load_4d(sbase, &map, 0, x0, y0, 0, bar);
    if (mode == 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < 400000) {
      }
      status[0] = 0;
    } else {
      const long long t0 = clock64();
      int ok = 0;
      while (clock64() - t0 < 4000000) {
        if (mbar_try(bar, 0)) {
          ok = 1;
          break;
        }
      }
      status[0] = ok;
    }
  }
  __syncthreads();
  // un-swizzle: 16-byte chunk j of row r sits at chunk j ^ (r & 7); report element 0 (channel 0) and element 8 of each row
  for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(smem + r * 128);
    out[2 * r] = row[((0 ^ (r & 7)) * 4)];
    out[2 * r + 1] = row[((1 ^ (r & 7)) * 4)];
  }
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  const int H = 40, W = 24, C = 64;
  std::vector<__half> h(static_cast<size_t>(H) * W * C);
  // channel 0 = y + 1 (so that zero fill is distinguishable), channel 1 = x + 1, channel 8 = 1000 + y, channel 9 = x
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < C; ++c) {
        float v = 0.f;
        if (c == 0) v = y + 1;
        if (c == 1) v = x + 1;
        if (c == 8) v = 1000 + y;
        if (c == 9) v = x;
        h[(static_cast<size_t>(y) * W + x) * C + c] = __float2half(v);
      }
  __half* d;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  uint32_t* dout;
  int* dstat;
  cudaMalloc(&dout, ROWS * 8);
  cudaMalloc(&dstat, 4);
  const int smem = ROWS * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Case { int bw, bh, x0, y0; };
  const Case cases[] = {{17, 33, -1, -1}, {16, 33, 0, -1}, {17, 32, -1, 0}, {16, 32, 0, 0}, {17, 33, 15, 15}, {15, 31, 0, 0}};
  for (const Case& cs : cases) {
    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
    const cuuint32_t box[4] = {64, (cuuint32_t)cs.bw, (cuuint32_t)cs.bh, 1};
    const cuuint32_t es[4] = {1, 2, 2, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box (64,%d,%d,1) strides (1,2,2,1) start (%d,%d): encode rc=%d\n", cs.bw, cs.bh, cs.x0, cs.y0, (int)r);
    if (r != CUDA_SUCCESS) continue;
    probe<<<1, 128, smem>>>(map, cs.x0, cs.y0, 0, 0, dout, dstat);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("  launch failed: %s\n", cudaGetErrorString(e));
      return 1;
    }
    std::vector<uint32_t> o(ROWS * 2);
    cudaMemcpy(o.data(), dout, ROWS * 8, cudaMemcpyDeviceToHost);
    int written = 0, last = -1;
    for (int rr = 0; rr < ROWS; ++rr)
      if (o[2 * rr] != 0xdeadbeefu) {
        ++written;
        last = rr;
      }
    printf("  rows written: %d (last row %d); ceil(bw/2)*ceil(bh/2) = %d\n", written, last, ((cs.bw + 1) / 2) * ((cs.bh + 1) / 2));
    const int pw = (cs.bw + 1) / 2;
    printf("  first rows (y+1, x+1 | chunk1: 1000+y, x):");
    for (int rr = 0; rr < 2 * pw + 2 && rr < ROWS; ++rr) {
      const __half2 a = *reinterpret_cast<const __half2*>(&o[2 * rr]);
      const __half2 b = *reinterpret_cast<const __half2*>(&o[2 * rr + 1]);
      printf(" [%d: %g,%g | %g,%g]", rr, __half2float(a.x), __half2float(a.y), __half2float(b.x), __half2float(b.y));
    }
    printf("\n");
    // check the hypothesis row = ly * pw + lx  <->  pixel (y0 + 2 ly, x0 + 2 lx)
    int bad = 0;
    for (int rr = 0; rr < written; ++rr) {
      const int ly = rr / pw, lx = rr % pw;
      const int y = cs.y0 + 2 * ly, x = cs.x0 + 2 * lx;
      const bool in = y >= 0 && y < H && x >= 0 && x < W;
      const __half2 a = *reinterpret_cast<const __half2*>(&o[2 * rr]);
      const float ey = in ? y + 1 : 0.f, ex = in ? x + 1 : 0.f;
      if (__half2float(a.x) != ey || __half2float(a.y) != ex) ++bad;
    }
    printf("  mapping row = ly*%d + lx <-> pixel (y0+2ly, x0+2lx), zero fill outside: %s (%d mismatches)\n", pw,
           bad ? "NO" : "yes", bad);
    const uint32_t tx = static_cast<uint32_t>(written) * 128u;
    probe<<<1, 128, smem>>>(map, cs.x0, cs.y0, 1, tx, dout, dstat);
    e = cudaDeviceSynchronize();
    int st = -1;
    cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost);
    printf("  expect_tx = rows written x 128 = %u bytes: barrier completed = %d (%s)\n", tx, st, cudaGetErrorString(e));
  }
  return 0;
}

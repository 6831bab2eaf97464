// Micro-probe: cycles per tcgen05.mma (kind::f16, M=128/64, K=16) as a function of N, of where A comes from
// (shared memory SW128 rows vs TMEM) and of the A window alignment (shifted conv windows).  One CTA per SM,
// a burst of back-to-back MMAs into one accumulator, clock64 around issue .. commit arrival.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../intra-and-inter-human-relation-network-for-mpee_b200/csrc/i2r_common.cuh"

namespace i2r {
void set_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
}  // namespace i2r
using namespace i2r;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Args {
  int M, N, mode;      // mode 0: SS aligned A; 1: SS shifted A windows (conv taps, SBO 1280); 2: TS (A in TMEM)
  int nmma;            // MMAs per burst
  int bursts;
  int rotate;          // distinct A/B K-slices to rotate through (1 = same operands every time)
  long long* out;      // per CTA: cycles
  int spin;            // 0: other warps idle at the final barrier; 1: they spin on an mbarrier (all lanes); 2: one lane per warp spins
};

template <int MODE, int ROT>
__global__ void __launch_bounds__(384, 1) probe(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase;             // one mbarrier
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  const uint32_t a_base = sbase + 1024;                 // A region: 64 KB
  const uint32_t b_base = sbase + 1024 + 64 * 1024;     // B region: 64 KB
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0x3c003c00u;  // fp16 1.0
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    mbar_init(bar + 16, 1);
    mbar_init(bar + 24, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 0) {
    tmem_alloc(smem_u32(slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(a.M, a.N);
    const uint32_t b_hi = sw128_desc_hi(1024, 0);
    const uint32_t a_hi = sw128_desc_hi(MODE == 1 ? 1280 : 1024, 0);
    const bool leader = elect_one();
    const uint32_t a_lo0 = sw128_desc_lo(a_base), b_lo0 = sw128_desc_lo(b_base);
    uint32_t ph[2] = {0, 0};
    long long total = 0;
    // pipelined bursts (one "tile" each): issue burst b into accumulator b&1, commit to bar2[b&1], then wait
    // for burst b-1 -- the MMA pipe always has the next tile queued, as in the persistent conv kernel.
    const uint32_t bar2 = bar + 16;
    __syncwarp();
    const long long t0 = clock64();
    for (int b = 0; b < a.bursts; ++b) {
      if (leader) {
        const uint32_t d = tmem + (b & 1) * 256;
        for (int i0 = 0; i0 < a.nmma; i0 += 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int r = ROT ? j : 0;
            const uint32_t koff = (r & 3) * 2;                        // 32-byte K slice inside the 128-byte row
            const uint32_t blk = (r >> 2);
            uint32_t a_addr = blk * 16384;
            if (MODE == 1) a_addr += ((j % 9) / 3 * 10 + (j % 9) % 3) * 128;   // shifted tap windows
            const uint32_t b_addr = blk * 8192;
            const uint64_t bd = desc64(b_lo0 + (b_addr >> 4) + koff, b_hi);
            if (MODE == 2) {
              umma_f16_ts(d, tmem + 480 + (r % 4) * 8, bd, idesc, 1u);
            } else {
              umma_f16(d, desc64(a_lo0 + (a_addr >> 4) + koff, a_hi), bd, idesc, (j | i0) ? 1u : 0u);
            }
          }
        }
        umma_commit(bar2 + 8 * (b & 1));
      }
      __syncwarp();
      if (b > 0) {
        mbar_wait(bar2 + 8 * ((b - 1) & 1), ph[(b - 1) & 1]);
        ph[(b - 1) & 1] ^= 1;
        tc_fence_after();
      }
    }
    mbar_wait(bar2 + 8 * ((a.bursts - 1) & 1), ph[(a.bursts - 1) & 1]);
    tc_fence_after();
    total = clock64() - t0;
    if (tid == 0) a.out[blockIdx.x] = total;
    if (tid == 0) mbar_arrive(bar + 8);
  } else if (a.spin == 1 || (a.spin == 2 && (tid & 31) == 0)) {
    mbar_wait(bar + 8, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = 1024 + 1024 + 128 * 1024;
  typedef void (*KFn)(Args);
  KFn fns[3][2] = {{probe<0, 0>, probe<0, 1>}, {probe<1, 0>, probe<1, 1>}, {probe<2, 0>, probe<2, 1>}};
  for (auto& row : fns)
    for (auto f : row) cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * sms);
  const int Ns[] = {16, 32, 48, 64, 80, 96, 112, 128, 144, 160, 192, 224, 256};
  const char* names[] = {"SS aligned", "SS shifted-taps", "TS (A in TMEM)"};
  printf("cycles per tcgen05.mma kind::f16 K=16 (burst of 256, 4 bursts averaged; per-CTA max over grid)\n");
  for (int nm : {16, 32, 64, 256}) {
    const int spin = 0;
    for (int M : {128}) {
      for (int mode = 0; mode < 3; ++mode) {
        for (int rotate : {16}) {
          const int grid = sms;
          printf("burst %3d M %3d %-16s rotate %2d :", nm, M, names[mode], rotate);
          for (int N : Ns) {
            if (M == 128 && N % 16) { printf("     -"); continue; }
            Args a{M, N, mode, nm, 4096 / nm, rotate, d_out, spin};
            fns[mode][rotate > 1]<<<grid, 384, smem>>>(a);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf(" ERR(%s)", cudaGetErrorString(e));
              return 1;
            }
            std::vector<long long> h(grid);
            cudaMemcpy(h.data(), d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (auto v : h) mx = v > mx ? v : mx;
            printf(" %5.1f", (double)mx / 4096.0);
          }
          printf("\n");
        }
      }
    }
  }
  printf("N columns:");
  for (int N : Ns) printf(" %5d", N);
  printf("\n");
  return 0;
}

// Micro-probe: how many tcgen05.mma can be in flight before the issuing thread blocks?  Issue k MMAs
// (M128 x N x K16, SS), read the clock right after the k-th issue (no wait), then commit + wait.
#include <cstdio>
#include <vector>

#include "../intra-and-inter-human-relation-network-for-mpee_b200/csrc/i2r_common.cuh"
namespace i2r {
void set_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
}  // namespace i2r
using namespace i2r;

template <int K>
__global__ void __launch_bounds__(128, 1) probe(int N, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  const uint32_t a_base = sbase + 1024, b_base = sbase + 1024 + 64 * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0x3c003c00u;
  if (tid == 0) {
    mbar_init(bar, 1);
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
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t b_hi = sw128_desc_hi(1024, 0), a_hi = sw128_desc_hi(1024, 0);
    const uint32_t a_lo0 = sw128_desc_lo(a_base), b_lo0 = sw128_desc_lo(b_base);
    const bool leader = elect_one();
    uint32_t ph = 0;
    long long t_issue = 0, t_done = 0;
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
      long long t1 = t0;
      if (leader) {
#pragma unroll
        for (int j = 0; j < K; ++j)
          umma_f16(tmem, desc64(a_lo0 + (j & 3) * 2, a_hi), desc64(b_lo0 + (j & 3) * 2, b_hi), idesc, 1u);
        t1 = clock64();
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, ph);
      ph ^= 1;
      tc_fence_after();
      const long long t2 = clock64();
      t_issue = t1 - t0;
      t_done = t2 - t0;
    }
    if (leader) {
      out[0] = t_issue;
      out[1] = t_done;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int K>
void run(int N, long long* d) {
  const int smem = 2048 + 128 * 1024;
  cudaFuncSetAttribute(probe<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<K><<<1, 128, smem>>>(N, d);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d k=%2d: issue %5lld clk, done %5lld clk\n", N, K, h[0], h[1]);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int N : {48, 96, 256}) {
    run<1>(N, d); run<2>(N, d); run<3>(N, d); run<4>(N, d); run<6>(N, d); run<8>(N, d); run<12>(N, d);
    run<16>(N, d); run<24>(N, d); run<32>(N, d); run<48>(N, d); run<64>(N, d);
  }
  return 0;
}

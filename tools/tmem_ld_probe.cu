// Micro-probe: tcgen05.ld cost for the epilogue pattern (8 warps, each 32 lanes x 24 fp32 columns per "tile")
// alone and while another warp keeps the tensor pipe busy with M128 x N48 x K16 MMAs.
#include <cstdio>
#include <vector>

#include "../intra-and-inter-human-relation-network-for-mpee_b200/csrc/i2r_common.cuh"
namespace i2r {
void set_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
}  // namespace i2r
using namespace i2r;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// mode: 0 = 3 x ld.x8 + wait ; 1 = ld.x8, wait each ; 2 = ld.x16 + ld.x8 + wait ; 3 = ld.x32 + wait (32 cols)
__global__ void __launch_bounds__(384, 1) probe(int mode, int with_mma, int nwarps, int iters, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  volatile int* stop = reinterpret_cast<volatile int*>(smem + 128);
  const uint32_t a_base = sbase + 1024, b_base = sbase + 1024 + 64 * 1024;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0x3c003c00u;
  if (tid == 0) {
    mbar_init(bar, 1);
    *stop = 0;
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 2) {
    tmem_alloc(smem_u32(slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 2) {
    if (with_mma) {
      const uint32_t idesc = make_idesc_f16(128, 48);
      const uint32_t b_hi = sw128_desc_hi(1024, 0), a_hi = sw128_desc_hi(1024, 0);
      const uint32_t a_lo0 = sw128_desc_lo(a_base), b_lo0 = sw128_desc_lo(b_base);
      const bool leader = elect_one();
      uint32_t ph = 0;
      while (!*stop) {
        if (leader) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            umma_f16(tmem + 256, desc64(a_lo0 + (j & 3) * 2, a_hi), desc64(b_lo0 + (j & 3) * 2, b_hi), idesc, 1u);
          umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, ph);
        ph ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 4 + nwarps) {
    const int quad = warp & 3;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(quad * 32) << 16) + ((warp - 4) >> 2) * 24;
    float acc = 0.f;
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
        uint32_t v[3][8];
        tmem_ld8(taddr, v[0]);
        tmem_ld8(taddr + 8, v[1]);
        tmem_ld8(taddr + 16, v[2]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += __uint_as_float(v[j][i]);
      } else if (mode == 1) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t v[8];
          tmem_ld8(taddr + 8 * j, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += __uint_as_float(v[i]);
        }
      } else if (mode == 2) {
        uint32_t v[16], w[8];
        tmem_ld16(taddr, v);
        tmem_ld8(taddr + 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += __uint_as_float(v[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += __uint_as_float(w[i]);
      } else {
        uint32_t v[32];
        tmem_ld32(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[i]);
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp - 4] = t1 - t0;
    if (acc == 123.f) sink[0] = acc;
    asm volatile("bar.sync 1, %0;" ::"r"(nwarps * 32));
    if (warp == 4 && lane == 0) *stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 64);
  cudaMalloc(&sink, 4);
  const int smem = 2048 + 128 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"3x ld.x8 + wait", "ld.x8+wait x3", "ld.x16+ld.x8+wait", "ld.x32+wait(32c)"};
  for (int with_mma = 0; with_mma < 2; ++with_mma)
    for (int nw : {1, 4, 8})
      for (int mode = 0; mode < 4; ++mode) {
        probe<<<1, 384, smem>>>(mode, with_mma, nw, 200, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 1; }
        long long h[8];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("mma %d warps %d %-18s : %7.1f clk per tile-iteration\n", with_mma, nw, names[mode], mx / 200.0);
      }
  return 0;
}

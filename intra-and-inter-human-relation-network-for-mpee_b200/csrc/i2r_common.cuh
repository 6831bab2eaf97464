// Shared device helpers for the sm_100a kernels: mbarrier, cp.async, bulk copy, tcgen05/TMEM PTX.
// Everything here is inline PTX for compute_100a; there is no fallback for other architectures.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/i2r.h"

namespace i2r {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// Per-device caches (SM count, shared-memory opt-in done) are indexed by the current device ordinal: the reference's
// tools/test.py wraps the module in nn.DataParallel(device_ids=cfg.GPUS), i.e. several devices in one process.
constexpr int MAX_DEVICES = 64;
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d < 0 || d >= MAX_DEVICES) ? 0 : d;
}

// ---------------------------------------------------------------- programmatic dependent launch
// The forward is a chain of ~130 short dependent kernels.  Every kernel is launched with the programmatic stream
// serialization attribute, calls pdl_launch_dependents() first (so the NEXT grid's CTAs are scheduled as soon as SMs
// free up in this grid's tail and run their prologue: barrier init, TMEM allocation, weight loads) and pdl_wait()
// before its first access to memory a predecessor may still be producing or reading.  Because every kernel of the
// chain waits, completion is transitive: when a kernel passes its wait all earlier kernels have finished.
// I2R_PDL=0 in the environment launches without the attribute (the device instructions are then no-ops).
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint32_t bar) {   // required before an mbarrier is initialised a second time
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded waits: a protocol bug reports itself and traps (launch error on the host) instead of hanging the GPU.
// The bound is TIME based (2^31 SM cycles, ~1.1 s: no legitimate wait of these kernels is longer than a few ms) so
// every stuck role of every CTA times out within the same window.  When a hang sink is installed
// (i2r_debug_hang_buffer: host-mapped memory, readable after the context died) lane 0 of each timed-out warp records
// {blockIdx.x << 32 | threadIdx.x, source line << 32 | barrier shared address, parity << 32 | dynamic-smem base,
// clock64} and lingers 2^28 cycles before the trap so that the other stuck roles get to report as well.
static __device__ unsigned long long* g_hang_sink = nullptr;   // one copy per translation unit (no -rdc)
static __device__ unsigned int g_hang_count = 0;
constexpr unsigned int HANG_SINK_RECORDS = 4096;
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int line) {
  extern __shared__ uint8_t hang_dyn_smem[];
  unsigned long long* sink = g_hang_sink;
  if (sink != nullptr) {
    const unsigned int idx = atomicAdd(&g_hang_count, 1u);
    if (idx < HANG_SINK_RECORDS) {
      volatile unsigned long long* e = sink + 4ull * idx;
      e[0] = (static_cast<unsigned long long>(blockIdx.x) << 32) | threadIdx.x;
      e[1] = (static_cast<unsigned long long>(line) << 32) | bar;
      e[2] = (static_cast<unsigned long long>(parity) << 32) |
             static_cast<uint32_t>(__cvta_generic_to_shared(hang_dyn_smem));
      e[3] = static_cast<unsigned long long>(clock64());
    }
    __threadfence_system();
    const long long t1 = clock64();
    while (clock64() - t1 < (1ll << 28)) {
    }
  }
  __trap();
}
#define I2R_HANG_SINK_SETTER(name)                                              \
  void hang_sink_##name(void* host_mapped) {                                    \
    unsigned long long* p = static_cast<unsigned long long*>(host_mapped);      \
    unsigned int zero = 0;                                                      \
    cudaMemcpyToSymbol(g_hang_sink, &p, sizeof(p));                             \
    cudaMemcpyToSymbol(g_hang_count, &zero, sizeof(zero));                      \
  }

__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_at(uint32_t bar, uint32_t parity, int line) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > (1ll << 31)) mbar_timeout(bar, parity, line);
  }
}
// Wait used by roles that run far ahead of their consumers (TMA producers): back off between polls so the
// spinning warp does not compete for issue slots with the epilogue warps of its scheduler partition.
__device__ __forceinline__ void mbar_wait_relaxed_at(uint32_t bar, uint32_t parity, int line) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    __nanosleep(128);
    if ((++spins & 255u) == 0 && clock64() - t0 > (1ll << 31)) mbar_timeout(bar, parity, line);
  }
}
// Wait for a WHOLE WARP whose lanes do not all take part in the work that follows (MMA issuer warps: the loop is
// warp-uniform, one elected lane issues).  ONE lane polls and the warp re-converges at __syncwarp.  Letting all 32
// lanes poll independently is NOT equivalent: the polling loop is a divergent exit, nothing forces the lanes back
// together before the next wait, and the elected lane -- the only one with work to do -- can run a whole ring
// revolution ahead of lanes still sitting in an earlier wait; the barrier they watch then completes a second phase,
// the parity they wait for becomes the parity of the CURRENT incomplete phase again, and they spin forever while
// every other role of the CTA finishes (the round-1 "unspecified launch failure": 31 lanes of the issuer warp stuck on
// wfull[s] of a finished CTA, profiles/r02_hang_hunt.txt).  Whether it happens depends on where the compiler places the
// re-convergence point of the polling loop and on the scheduler, hence the sensitivity to unrelated code changes.
__device__ __forceinline__ void mbar_wait_warp_at(uint32_t bar, uint32_t parity, int line) {
  if ((threadIdx.x & 31u) == 0) mbar_wait_at(bar, parity, line);
  __syncwarp();
}
#define mbar_wait(bar, parity) ::i2r::mbar_wait_at((bar), (parity), __LINE__)
#define mbar_wait_warp(bar, parity) ::i2r::mbar_wait_warp_at((bar), (parity), __LINE__)
#define mbar_wait_relaxed(bar, parity) ::i2r::mbar_wait_relaxed_at((bar), (parity), __LINE__)

// ---------------------------------------------------------------- async copies
// 16-byte global->shared copy, zero-filled when src_bytes == 0 (padding taps / rows past M).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy writes (cp.async / st.shared) -> visible to the async proxy (tcgen05.mma, TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global->shared (TMA unit, SASS UBLKCP), completion counted on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: lane (= accumulator row) of this thread, 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): core matrix = 8 rows x 16 B
// stored contiguously (128 B); SBO = byte distance between 8-row groups along M/N, LBO = byte
// distance between the two 8-element K halves of one K=16 step.  Bits: [0,14) addr>>4,
// [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1 (sm_100), [61,64) layout type 0.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// Instruction descriptor for kind::f16: fp16 A/B (format 0), fp32 accumulator (c_format 1 at bit 4),
// both operands K-major, N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// ---- SWIZZLE_128B K-major operands (the full-rate shared-memory path of the tensor core) ----------------
// A row (one pixel / one output channel) is 128 B = 64 fp16 of K; rows are 128 B apart, 8-row groups SBO bytes
// apart; inside a row the eight 16-byte chunks are XOR-ed with (row address >> 7) & 7.  One K=16 MMA step
// consumes 32 B of every row: advance the start address by 32 B.  base_offset = (start >> 7) & 7 when the
// start address is not 1024-byte aligned (shifted conv windows).
__device__ __forceinline__ uint32_t sw128_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t sw128_desc_hi(uint32_t sbo_bytes, uint32_t base_offset) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((base_offset & 7u) << 17) | (2u << 29);
}
// byte offset of 16-byte chunk `c` of row `row` inside a 1024-byte-aligned SW128 tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t c) { return row * 128u + ((c ^ (row & 7u)) << 4); }

// Split form for hot issue loops: `hi` is loop-invariant, `lo` = (addr >> 4) | (LBO >> 4) << 16 advances by
// plain 32-bit adds of (byte_offset >> 4).
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Straight-line MMA bursts (executed by the elected lane only): all descriptor offsets are immediates.
template <int KS>
__device__ __forceinline__ void issue_ksteps(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int k2 = 0; k2 < KS; ++k2)
    umma_f16(d_tmem, desc64(a_lo + 2 * k2, a_hi), desc64(b_lo + 2 * k2, b_hi), idesc, k2 ? 1u : acc_first);
}
// erf-GELU, 0.5 x (1 + erf(x / sqrt 2)), with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. at fp32
// rounding level): one reciprocal, one exp2 and seven FMAs instead of the ~45 instructions of erff().  The GELU layers
// of HRFormer-B (MlpDWBN, lib/models/hrformer.py:1094-1119) evaluate it 25 M times per crop in epilogue warps that are
// instruction-issue bound.
__device__ __forceinline__ float gelu_erf(float v) {
  const float z = fabsf(v) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float e = exp2f(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p * t, e, 1.0f);          // erf(|v| / sqrt 2)
  return 0.5f * v + 0.5f * fabsf(v) * erf_abs;          // 0.5 v (1 + sign(v) erf_abs)
}
// epilogue activation: ReLU (I2R_F_RELU), erf-GELU (I2R_F_GELU) or identity
// (deliberately NOT inlined: the GELU layers are cold paths of kernels whose hot loops must stay small)
static __device__ __noinline__ float epi_act(float v, uint32_t flags) {
  if (flags & I2R_F_GELU) return gelu_erf(v);
  if (flags & I2R_F_RELU) return fmaxf(v, 0.f);
  return v;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}

}  // namespace i2r

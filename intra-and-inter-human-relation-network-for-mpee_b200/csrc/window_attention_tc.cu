// Window attention of HRFormer-B's InterlacedPoolAttention / MHA_ (lib/models/hrformer.py:627-935, :949-1000) on the
// 5th-generation tensor cores: softmax(scale * q k^T) v over 7x7 = 49-token windows, `heads` heads of 39 channels padded
// to 48, no relative-position bias, no mask (padded tokens take part), split-operand (hi | lo) or plain fp16 rows.
//
// One CTA = one head of TWO windows: the 98 token rows [98 p, 98 p + 98) of the window-major q / k / v tensors form one
// 128-row tcgen05 tile (rows 98..127 belong to the next windows and are computed but never stored):
//   warp 0      TMA: Q, K, V tiles as [128 rows x 64 channels] SWIZZLE_128B boxes at channel h*48 (hi) and lo_off + h*48
//               (lo); one elected lane then issues S = Q K^T (M128 x N128 x K48, three MMAs per K step in split mode:
//               q_hi k_hi + q_lo k_hi + q_hi k_lo) into TMEM, waits for the softmax and issues O = P V with P read from
//               TMEM (A-from-TMEM form) and V consumed as an MN-MAJOR B operand straight from the token-major tile -- no
//               transposed copy of V exists anywhere (b_major bit of the instruction descriptor; 8-key groups 1024 B apart);
//   warps 1-4   softmax, one thread per query row (tcgen05.ld 32x32b): the block-diagonal structure is a column range
//               -- row r of window w = r / 49 attends to columns [49 w, 49 w + 49) -- everything else gets P = 0; P is
//               written back over S as packed fp16 (tcgen05.st); then O / l from TMEM to global memory (hi | lo).
// The work per CTA is tiny (9 + 16 MMAs): the kernel is latency bound like its mma.sync predecessor (attention.cu,
// kept as the check implementation and for other window sizes); what this version changes is that the whole HRFormer
// path now runs on tcgen05 / TMEM / TMA.
#include <stdlib.h>

#include "i2r_tma.cuh"

namespace i2r {

constexpr int WT_THREADS = 160;
constexpr int WT_WIN = 49, WT_HD = 48, WT_ROWS = 2 * WT_WIN;

struct WtArgs {
  __half* out;
  int ldo, o_lo, q_lo, k_lo, v_lo;
  int total_rows, heads;
  float scale_log2e;
  uint32_t v_lbo, v_sbo, v_major;   // descriptor fields of the MN-major V operand.  Probed on B200 (profiles/
                                    // r02_window_attention_tc.txt): b_major = 1, SBO = 1024 (8-key atoms), LBO unused for N <= 64;
                                    // LBO / SBO swapped or b_major = 0 give wrong results
};

template <bool SPLIT>
__global__ void __launch_bounds__(WT_THREADS, 2)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                           const __grid_constant__ CUtensorMap mapV, const WtArgs A) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y;
  const int row0 = blockIdx.x * WT_ROWS;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t CH = TC_CH_BYTES;                       // 128 rows x 128 B
  const uint32_t sQ = sbase, sK = sQ + (SPLIT ? 2 : 1) * CH, sV = sK + (SPLIT ? 2 : 1) * CH;
  const uint32_t sBar = sV + (SPLIT ? 2 : 1) * CH;
  const uint32_t bLoad = sBar, bS = sBar + 8, bP = sBar + 16, bO = sBar + 24, sSlot = sBar + 32;
  if (tid == 0) {
    mbar_init(bLoad, 1);
    mbar_init(bS, 1);
    mbar_init(bP, 128);
    mbar_init(bO, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(sSlot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sSlot));
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      pdl_wait();   // q / k / v come from the projection GEMMs
      mbar_arrive_expect_tx(bLoad, 3 * (SPLIT ? 2 : 1) * CH);
      tma_load_2d(sQ, &mapQ, h * WT_HD, row0, bLoad);
      tma_load_2d(sK, &mapK, h * WT_HD, row0, bLoad);
      tma_load_2d(sV, &mapV, h * WT_HD, row0, bLoad);
      if (SPLIT) {
        tma_load_2d(sQ + CH, &mapQ, A.q_lo + h * WT_HD, row0, bLoad);
        tma_load_2d(sK + CH, &mapK, A.k_lo + h * WT_HD, row0, bLoad);
        tma_load_2d(sV + CH, &mapV, A.v_lo + h * WT_HD, row0, bLoad);
      }
      mbar_wait(bLoad, 0);
      tc_fence_after();
      const uint32_t idS = make_idesc_f16(128, 128);
      const uint32_t hiK = sw128_desc_hi(1024, 0);
      const uint32_t q0 = sw128_desc_lo(sQ), k0 = sw128_desc_lo(sK);
#pragma unroll
      for (int s = 0; s < WT_HD / 16; ++s) {
        umma_f16(tS, desc64(q0 + 2 * s, hiK), desc64(k0 + 2 * s, hiK), idS, s ? 1u : 0u);
        if (SPLIT) {
          umma_f16(tS, desc64(q0 + (CH >> 4) + 2 * s, hiK), desc64(k0 + 2 * s, hiK), idS, 1u);
          umma_f16(tS, desc64(q0 + 2 * s, hiK), desc64(k0 + (CH >> 4) + 2 * s, hiK), idS, 1u);
        }
      }
      umma_commit(bS);
      mbar_wait(bP, 0);
      tc_fence_after();
      // O = P V: A = P (128 x 128 keys, packed fp16 in TMEM columns [0, 64)), B = V tile, MN-major: a 128-byte row holds
      // the 64 channels of ONE key, 8 keys = one 1024-byte swizzle atom, a K = 16 step = two atoms
      const uint32_t idO = make_idesc_f16(128, WT_HD) | (A.v_major << 16);
      const uint32_t hiV = ((A.v_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t v0 = ((sV >> 4) & 0x3FFFu) | (((A.v_lbo >> 4) & 0x3FFFu) << 16);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        umma_f16_ts(tO, tS + kk * 8, desc64(v0 + kk * (2048 >> 4), hiV), idO, kk ? 1u : 0u);
        if (SPLIT) umma_f16_ts(tO, tS + kk * 8, desc64(v0 + (CH >> 4) + kk * (2048 >> 4), hiV), idO, 1u);
      }
      umma_commit(bO);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const int w = row >= WT_ROWS ? 2 : (row >= WT_WIN ? 1 : 0);
    const int c_lo = WT_WIN * w, c_hi = c_lo + WT_WIN;       // this row's key columns (rows >= 98: none)
    mbar_wait(bS, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(tS + lane_base + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = c * 32 + i;
        mx = fmaxf(mx, (w < 2 && col >= c_lo && col < c_hi) ? __uint_as_float(r[i]) : -INFINITY);
      }
    }
    const float sc = A.scale_log2e;
    const float m = (w < 2) ? mx * sc : 0.f;
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(tS + lane_base + c * 32, r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = c * 32 + 2 * i;
        float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), sc, -m));
        float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), sc, -m));
        if (!(w < 2 && col >= c_lo && col < c_hi)) p0 = 0.f;
        if (!(w < 2 && col + 1 >= c_lo && col + 1 < c_hi)) p1 = 0.f;
        l += p0 + p1;
        pk[i] = pack_h2(p0, p1);
      }
      // (chunk c of S is fully in registers before its first half is overwritten by the packed P columns [16c, 16c+16))
      tmem_st16(tS + lane_base + c * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bP);
    mbar_wait(bO, 0);
    tc_fence_after();
    const int grow = row0 + row;
    const bool store = w < 2 && grow < A.total_rows;       // (the TMEM loads are warp-collective: every lane runs them)
    const float inv = store ? 1.f / l : 0.f;
    __half* orow = A.out + static_cast<int64_t>(store ? grow : 0) * A.ldo + h * WT_HD;
#pragma unroll
    for (int c = 0; c < WT_HD / 16; ++c) {
      uint32_t o[16];
      tmem_ld16(tO + lane_base + c * 16, o);
      tmem_ld_wait();
      uint32_t hv[8], lv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = __uint_as_float(o[2 * i]) * inv, b = __uint_as_float(o[2 * i + 1]) * inv;
        hv[i] = pack_h2(a, b);
        lv[i] = 0;
        if (SPLIT) {
          const float2 f = unpack_h2(hv[i]);
          lv[i] = pack_h2(a - f.x, b - f.y);
        }
      }
      if (store) {
        *reinterpret_cast<uint4*>(orow + c * 16) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
        *reinterpret_cast<uint4*>(orow + c * 16 + 8) = make_uint4(hv[4], hv[5], hv[6], hv[7]);
        if (SPLIT) {
          *reinterpret_cast<uint4*>(orow + A.o_lo + c * 16) = make_uint4(lv[0], lv[1], lv[2], lv[3]);
          *reinterpret_cast<uint4*>(orow + A.o_lo + c * 16 + 8) = make_uint4(lv[4], lv[5], lv[6], lv[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

template <bool SPLIT>
static int launch_window_attention_tc(const void* q, const void* k, const void* v, void* out, int ldq, int ldk, int ldv,
                                      int ldo, int nwin, int heads, float scale, int q_lo, int k_lo, int v_lo, int o_lo,
                                      cudaStream_t st) {
  constexpr int SMEM = 3 * (SPLIT ? 2 : 1) * TC_CH_BYTES + 64 + 1024;
  static bool attr_done_dev[MAX_DEVICES] = {};
  bool& attr_done = attr_done_dev[current_device()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_tc_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(window_attention_tc): %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_done = true;
  }
  const int rows = nwin * WT_WIN;
  const int hq = heads * WT_HD;
  CUtensorMap mq, mk, mv;
  // the maps span the whole row (hi | lo); boxes are 64 channels wide, 48 of them used (the rest belongs to the next
  // head or is zero fill past the row)
  int rc = encode_2d(&mq, q, SPLIT ? q_lo + hq : hq, rows, ldq, 64, 128);
  if (!rc) rc = encode_2d(&mk, k, SPLIT ? k_lo + hq : hq, rows, ldk, 64, 128);
  if (!rc) rc = encode_2d(&mv, v, SPLIT ? v_lo + hq : hq, rows, ldv, 64, 128);
  if (rc) return rc;
  WtArgs A;
  A.out = static_cast<__half*>(out);
  A.ldo = ldo;
  A.o_lo = o_lo;
  A.q_lo = q_lo;
  A.k_lo = k_lo;
  A.v_lo = v_lo;
  A.total_rows = rows;
  A.heads = heads;
  A.scale_log2e = scale * 1.4426950408889634f;
  A.v_lbo = 16;
  A.v_sbo = 1024;
  A.v_major = 1;
  dim3 grid((nwin + 1) / 2, heads);
  launch_pdl(window_attention_tc_kernel<SPLIT>, grid, dim3(WT_THREADS), static_cast<size_t>(SMEM), st, mq, mk, mv, A);
  return check_launch("window_attention_tc_kernel");
}

I2R_HANG_SINK_SETTER(window_attention_tc)
}  // namespace i2r

extern "C" int i2r_window_attention_tc(const void* q, const void* k, const void* v, void* out, int ldq, int ldk, int ldv,
                                       int ldo, int nwin, int win_len, int heads, int head_pad, float scale, int split,
                                       int q_lo, int k_lo, int v_lo, int o_lo, void* stream) {
  using namespace i2r;
  if (!q || !k || !v || !out || nwin <= 0 || heads <= 0 || heads > 65535 || (ldq | ldk | ldv | ldo) % 8 != 0 ||
      (split && (q_lo | k_lo | v_lo | o_lo) % 8 != 0) || scale <= 0.f ||
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(out)) & 15) != 0) {
    set_error("i2r_window_attention_tc: bad arguments (16-byte aligned pointers, strides multiples of 8)");
    return I2R_E_BADARG;
  }
  if (win_len != WT_WIN || head_pad != WT_HD) {
    set_error("i2r_window_attention_tc: 49-token windows of 48-channel heads only (got %d, %d)", win_len, head_pad);
    return I2R_E_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return split ? launch_window_attention_tc<true>(q, k, v, out, ldq, ldk, ldv, ldo, nwin, heads, scale, q_lo, k_lo, v_lo,
                                                  o_lo, st)
               : launch_window_attention_tc<false>(q, k, v, out, ldq, ldk, ldv, ldo, nwin, heads, scale, 0, 0, 0, 0, st);
}

// Ragged single-head attention on the 5th-generation tensor cores (tcgen05 + TMEM), the kernel behind the
// intra-human encoder of TransPose-H (3072 tokens per crop, lib/models/transpose_h.py:165-240) and the inter-human
// encoders (lib/models/attention.py:61-82, interformer_pureMulti.py:182-213).
//
// One CTA = up to two 128-query tiles of one sequence (they share every K/V block that is staged) and one range of
// 128-key blocks (split-KV; partials merged by attention_merge_kernel).  Ten warps:
//   warp 0      TMA producer: Q tiles once, then one K block [128 keys x (hi|lo) channels] and one V^T block
//               [(hi|lo) channels x 128 keys] per iteration, SWIZZLE_128B boxes, mbarrier complete_tx;
//   warp 1      MMA issuer (one elected lane): S_t = Q_t K^T (M128 x N128, operands in shared memory) into TMEM,
//               O_t += P_t V (M128 x N=HD, A = P_t read from TMEM, B = V^T rows in shared memory);
//   warps 2-5   softmax of tile 0, warps 6-9 softmax of tile 1: one thread per query row (tcgen05.ld 32x32b gives
//               a thread its own accumulator row): running max with LAZY rescaling (O and l are only rescaled when
//               the row maximum grew by more than 2^8, so P stays below 256 in fp16), P written back over S as
//               packed fp16 (tcgen05.st), final O / l written straight from TMEM to global memory.
// Issue order in steady state: PV_0(j), QK_0(j+1), PV_1(j), QK_1(j+1) -- the softmax of one tile runs while the
// tensor pipe works for the other.  Because tcgen05.mma of one thread complete in order, "S_t(j+1) is full" implies
// PV_t(j) has finished, which is what makes the in-place P and the O rescale race-free without extra barriers.
//
// Split-operand mode (I2R_F_SPLIT): rows of q and k are fp16 pairs [hi | lo] (lo directly after hi), V^T has the lo
// channel rows after the hi rows; S = q_hi k_hi + q_lo k_hi + q_hi k_lo (three MMAs per K step), O += P v_hi + P v_lo.
#include <stdlib.h>

#include "i2r_tma.cuh"

namespace i2r {

int attention_merge_launch(int HD, const float* opart, const float* mlpart, __half* out, int ldo, int rows, int nsplit,
                           int o_lo, cudaStream_t st);   // attention.cu

constexpr int TC_THREADS = 320;

struct TcArgs {
  __half* out;
  const int32_t* cu_seqlens;
  float* opart;
  float* mlpart;
  float scale_log2e;
  int ldo, o_lo, nsplit;
};

template <int HD, bool SPLIT>
struct TcCfg {
  static constexpr int KS = HD / 16;                                   // K steps over the head dim
  static constexpr int QCH = ((SPLIT ? 2 * HD : HD) + 63) / 64;        // 64-channel chunks of a q / k row
  static constexpr int Q_TILE = QCH * TC_CH_BYTES;
  static constexpr int K_BYTES = QCH * TC_CH_BYTES;
  static constexpr int VROWS = SPLIT ? 2 * HD : HD;                    // channel rows of a V^T block
  static constexpr int V_CH = VROWS * 128;                             // one 64-key chunk of V^T
  static constexpr int V_BYTES = 2 * V_CH;
  static constexpr int SMEM = 2 * Q_TILE + K_BYTES + V_BYTES + 256 + 1024;   // + barriers + alignment slack
};

template <int HD, bool SPLIT>
__device__ __forceinline__ void issue_qk(uint32_t d_tmem, uint32_t q_addr, uint32_t k_addr, uint32_t idesc) {
  constexpr int KS = HD / 16;
  const uint32_t hi = sw128_desc_hi(1024, 0);
  const uint32_t a0 = sw128_desc_lo(q_addr), b0 = sw128_desc_lo(k_addr);
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    umma_f16(d_tmem, desc64(a0 + (kstep_off(s) >> 4), hi), desc64(b0 + (kstep_off(s) >> 4), hi), idesc, s ? 1u : 0u);
    if (SPLIT) {
      umma_f16(d_tmem, desc64(a0 + (kstep_off(KS + s) >> 4), hi), desc64(b0 + (kstep_off(s) >> 4), hi), idesc, 1u);
      umma_f16(d_tmem, desc64(a0 + (kstep_off(s) >> 4), hi), desc64(b0 + (kstep_off(KS + s) >> 4), hi), idesc, 1u);
    }
  }
}

template <int HD, bool SPLIT>
__device__ __forceinline__ void issue_pv(uint32_t d_tmem, uint32_t p_tmem, uint32_t v_addr, uint32_t idesc,
                                         uint32_t acc_first) {
  constexpr int V_CH = (SPLIT ? 2 * HD : HD) * 128;
  const uint32_t hi = sw128_desc_hi(1024, 0);
  const uint32_t b0 = sw128_desc_lo(v_addr);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {   // 16 keys per step = 8 packed TMEM columns of P
    const uint32_t off = ((kk >> 2) * V_CH + (kk & 3) * 32) >> 4;
    umma_f16_ts(d_tmem, p_tmem + kk * 8, desc64(b0 + off, hi), idesc, kk ? 1u : acc_first);
    if (SPLIT) umma_f16_ts(d_tmem, p_tmem + kk * 8, desc64(b0 + off + ((HD * 128) >> 4), hi), idesc, 1u);
  }
}

template <int HD, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                    const __grid_constant__ CUtensorMap mapV, const TcArgs A) {
  using C = TcCfg<HD, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seq = blockIdx.y;
  pdl_wait();   // cu_seqlens / q / k / v are produced by predecessors; the output may still be read by one
  const int t0 = A.cu_seqlens[seq];
  const int L = A.cu_seqlens[seq + 1] - t0;
  const int q0 = blockIdx.x * 256;
  if (q0 >= L) return;
  const int nt = (L - q0 > 128) ? 2 : 1;
  const int nblk_all = (L + 127) >> 7;
  const int per_split = (nblk_all + A.nsplit - 1) / A.nsplit;
  const int jb = blockIdx.z * per_split;
  const int je = min(nblk_all, jb + per_split);
  const int n = je - jb;
  if (n <= 0) {
    // this key range is empty for this (short) sequence: neutral partials for the merge
    for (int r = tid; r < 256; r += TC_THREADS) {
      const int row = q0 + r;
      if (row < L) {
        float* op = A.opart + (static_cast<int64_t>(t0 + row) * A.nsplit + blockIdx.z) * HD;
        for (int c = 0; c < HD; ++c) op[c] = 0.f;
        *reinterpret_cast<float2*>(A.mlpart + (static_cast<int64_t>(t0 + row) * A.nsplit + blockIdx.z) * 2) =
            make_float2(-INFINITY, 0.f);
      }
    }
    return;
  }

  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sbase;
  const uint32_t sK = sQ + 2 * C::Q_TILE;
  const uint32_t sV = sK + C::K_BYTES;
  const uint32_t sBar = sV + C::V_BYTES;
  // barriers (8 bytes each)
  const uint32_t bQ = sBar, bKf = sBar + 8, bKe = sBar + 16, bVf = sBar + 24, bVe = sBar + 32;
  const uint32_t bS = sBar + 40;    // [2]
  const uint32_t bP = sBar + 56;    // [2]
  const uint32_t bO = sBar + 72;    // [2]
  const uint32_t sSlot = sBar + 96;
  if (tid == 0) {
    mbar_init(bQ, 1);
    mbar_init(bKf, 1);
    mbar_init(bKe, 1);
    mbar_init(bVf, 1);
    mbar_init(bVe, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(bS + 8 * t, 1);
      mbar_init(bP + 8 * t, 128);
      mbar_init(bO + 8 * t, 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(sSlot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sSlot));

  if (warp == 0) {
    // ------------------------------------------------------------------------------------- TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(bQ, nt * C::Q_TILE);
      for (int t = 0; t < nt; ++t)
#pragma unroll
        for (int ch = 0; ch < C::QCH; ++ch)
          tma_load_2d(sQ + t * C::Q_TILE + ch * TC_CH_BYTES, &mapQ, ch * 64, t0 + q0 + t * 128, bQ);
      for (int jj = 0; jj < n; ++jj) {
        const int krow = t0 + (jb + jj) * 128;
        if (jj > 0) mbar_wait_relaxed(bKe, (jj - 1) & 1);
        mbar_arrive_expect_tx(bKf, C::K_BYTES);
#pragma unroll
        for (int ch = 0; ch < C::QCH; ++ch) tma_load_2d(sK + ch * TC_CH_BYTES, &mapK, ch * 64, krow, bKf);
        if (jj > 0) mbar_wait_relaxed(bVe, (jj - 1) & 1);
        mbar_arrive_expect_tx(bVf, C::V_BYTES);
        tma_load_2d(sV, &mapV, krow, 0, bVf);
        tma_load_2d(sV + C::V_CH, &mapV, krow + 64, 0, bVf);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      const uint32_t idS = make_idesc_f16(128, 128), idO = make_idesc_f16(128, HD);
      mbar_wait(bQ, 0);
      mbar_wait(bKf, 0);
      tc_fence_after();
      for (int t = 0; t < nt; ++t) {
        issue_qk<HD, SPLIT>(tmem_base + t * 128, sQ + t * C::Q_TILE, sK, idS);
        umma_commit(bS + 8 * t);
      }
      umma_commit(bKe);
      for (int jj = 0; jj < n; ++jj) {
        const bool more = jj + 1 < n;
        mbar_wait(bVf, jj & 1);
        for (int t = 0; t < nt; ++t) {
          mbar_wait(bP + 8 * t, jj & 1);
          tc_fence_after();
          issue_pv<HD, SPLIT>(tmem_base + 256 + t * 128, tmem_base + t * 128, sV, idO, jj ? 1u : 0u);
          if (!more) umma_commit(bO + 8 * t);
          if (t == nt - 1) umma_commit(bVe);
          if (more) {
            if (t == 0) {
              mbar_wait(bKf, (jj + 1) & 1);
              tc_fence_after();
            }
            issue_qk<HD, SPLIT>(tmem_base + t * 128, sQ + t * C::Q_TILE, sK, idS);
            umma_commit(bS + 8 * t);
            if (t == nt - 1) umma_commit(bKe);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------- softmax + epilogue
    const int t = (warp - 2) >> 2;
    if (t < nt) {
      const int quad = warp & 3;              // the TMEM lane quadrant a warp may access is warp_id % 4
      const int row = quad * 32 + lane;       // accumulator row = query row inside the tile
      const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const uint32_t tS = lane_base + t * 128, tO = lane_base + 256 + t * 128;
      const float sc = A.scale_log2e;
      float m_used = -INFINITY, l = 0.f;
      for (int jj = 0; jj < n; ++jj) {
        mbar_wait(bS + 8 * t, jj & 1);
        tc_fence_after();
        const int valid = L - (jb + jj) * 128;     // keys of this block that belong to the sequence (>= 1)
        const bool partial = valid < 128;
        // pass 1: block maximum of the raw scores
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(tS + c * 32, r);
          tmem_ld_wait();
          if (partial) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c * 32 + i < valid) ? __uint_as_float(r[i]) : -INFINITY);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          }
        }
        const float m_new = fmaxf(m_used, mx * sc);
        if (__any_sync(0xffffffffu, m_new > m_used + 8.f)) {
          // lazy rescale (whole warp, so the TMEM accesses stay warp-collective); PV_t(jj-1) has completed (see top)
          const float alpha = ex2(m_used - m_new);
          l *= alpha;
          if (jj > 0) {
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
              uint32_t o[16];
              tmem_ld16(tO + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tO + c * 16, o);
            }
          }
          m_used = m_new;
        }
        // pass 2: P = 2^(s*scale - m_used) as packed fp16 over the first 64 columns of S
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(tS + c * 32, r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), sc, -m_used));
            float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), sc, -m_used));
            if (partial) {
              if (c * 32 + 2 * i >= valid) p0 = 0.f;
              if (c * 32 + 2 * i + 1 >= valid) p1 = 0.f;
            }
            l0 += p0;
            l1 += p1;
            pk[i] = pack_h2(p0, p1);
          }
          tmem_st16(tS + c * 16, pk);
        }
        l += l0 + l1;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bP + 8 * t);
      }
      // ---- epilogue: O / l straight from TMEM
      mbar_wait(bO + 8 * t, 0);
      tc_fence_after();
      const int qrow = q0 + t * 128 + row;
      const bool ok = qrow < L;
      const int64_t tok = t0 + qrow;
      if (A.nsplit == 1) {
        const float inv = 1.f / l;
        __half* orow = A.out + tok * A.ldo;
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) {
          uint32_t o[16];
          tmem_ld16(tO + c * 16, o);
          tmem_ld_wait();
          if (ok) {
            uint32_t h[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float a = __uint_as_float(o[2 * i]) * inv, b = __uint_as_float(o[2 * i + 1]) * inv;
              h[i] = pack_h2(a, b);
              if (SPLIT) {
                const float2 f = unpack_h2(h[i]);
                lo[i] = pack_h2(a - f.x, b - f.y);
              }
            }
            *reinterpret_cast<uint4*>(orow + c * 16) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(orow + c * 16 + 8) = make_uint4(h[4], h[5], h[6], h[7]);
            if (SPLIT) {
              *reinterpret_cast<uint4*>(orow + A.o_lo + c * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              *reinterpret_cast<uint4*>(orow + A.o_lo + c * 16 + 8) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
          }
        }
      } else {
        float* op = A.opart + (tok * A.nsplit + blockIdx.z) * HD;
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) {
          uint32_t o[16];
          tmem_ld16(tO + c * 16, o);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(op + c * 16 + 4 * i) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        }
        if (ok) *reinterpret_cast<float2*>(A.mlpart + (tok * A.nsplit + blockIdx.z) * 2) = make_float2(m_used, l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// number of key splits: minimise rounds(ctas / SMs) x blocks per split (+1 block per extra pass for the merge)
static int tc_choose_nsplit(int nseq, int max_seqlen, int sms) {
  const int qblocks = ((max_seqlen + 255) / 256) * nseq;
  const int nblk = (max_seqlen + 127) / 128;
  static const int forced = []() {
    const char* e = getenv("I2R_ATT_NSPLIT");     // tuning / test override
    return e ? atoi(e) : 0;
  }();
  if (forced > 0) return forced < nblk ? forced : nblk;
  int best = 1;
  double best_cost = 1e30;
  for (int ns = 1; ns <= 8 && ns <= nblk; ++ns) {
    const int per = (nblk + ns - 1) / ns;
    const int rounds = (qblocks * ns + sms - 1) / sms;
    const double cost = static_cast<double>(rounds) * (per + 0.75) + (ns > 1 ? 0.5 * ns : 0.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = ns;
    }
  }
  return best;
}

int device_sms() {
  static int sms_dev[MAX_DEVICES] = {};
  int& sms = sms_dev[current_device()];
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                     uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return I2R_E_DEVICE;
  }
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {row_stride_elems * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t ones[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for 2-D [%llu x %llu] stride %llu", (int)r,
              (unsigned long long)outer, (unsigned long long)inner, (unsigned long long)row_stride_elems);
    return I2R_E_DEVICE;
  }
  return 0;
}

template <int HD, bool SPLIT>
static int launch_attention_tc(const void* q, const void* k, const void* vt, void* out, int ldq, int ldk, int ldvt,
                               int ldo, const int32_t* cu, int nseq, int max_seqlen, int total_tokens, float scale,
                               void* ws, int64_t ws_bytes, int o_lo, cudaStream_t st) {
  using C = TcCfg<HD, SPLIT>;
  static bool attr_done_dev[MAX_DEVICES] = {};   // the opt-in is a per-device property
  bool& attr_done = attr_done_dev[current_device()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<HD, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attention_tc): %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_done = true;
  }
  int nsplit = tc_choose_nsplit(nseq, max_seqlen, device_sms());
  const int64_t need = static_cast<int64_t>(total_tokens) * nsplit * (HD + 2) * 4;
  if (nsplit > 1 && (ws == nullptr || ws_bytes < need)) nsplit = 1;
  CUtensorMap mq, mk, mv;
  const uint64_t width = SPLIT ? 2 * HD : HD;
  int rc = encode_2d(&mq, q, width, total_tokens, ldq, 64, 128);
  if (!rc) rc = encode_2d(&mk, k, width, total_tokens, ldk, 64, 128);
  if (!rc) rc = encode_2d(&mv, vt, total_tokens, C::VROWS, ldvt, 64, C::VROWS);
  if (rc) return rc;
  TcArgs A;
  A.out = static_cast<__half*>(out);
  A.cu_seqlens = cu;
  A.opart = static_cast<float*>(ws);
  A.mlpart = A.opart ? A.opart + static_cast<int64_t>(total_tokens) * nsplit * HD : nullptr;
  A.scale_log2e = scale * 1.4426950408889634f;
  A.ldo = ldo;
  A.o_lo = o_lo;
  A.nsplit = nsplit;
  dim3 grid((max_seqlen + 255) / 256, nseq, nsplit);
  launch_pdl(attention_tc_kernel<HD, SPLIT>, grid, dim3(TC_THREADS), static_cast<size_t>(C::SMEM), st, mq, mk, mv, A);
  rc = check_launch("attention_tc_kernel");
  if (rc || nsplit == 1) return rc;
  return attention_merge_launch(HD, A.opart, A.mlpart, A.out, ldo, total_tokens, nsplit, SPLIT ? o_lo : 0, st);
}

I2R_HANG_SINK_SETTER(attention_tc)
}  // namespace i2r

extern "C" int64_t i2r_attention_tc_workspace_bytes(int total_tokens, int D, int nseq, int max_seqlen) {
  const int ns = i2r::tc_choose_nsplit(nseq, max_seqlen, i2r::device_sms());
  return ns > 1 ? static_cast<int64_t>(total_tokens) * ns * (D + 2) * 4 : 0;
}

extern "C" int i2r_attention_tc(const void* q, const void* k, const void* vt, void* out, int ldq, int ldk, int ldvt,
                                int ldo, int D, const int32_t* cu_seqlens, int nseq, int max_seqlen, int total_tokens,
                                float scale, void* workspace, int64_t workspace_bytes, int split, int o_lo,
                                void* stream) {
  using namespace i2r;
  if (!q || !k || !vt || !out || !cu_seqlens || nseq <= 0 || max_seqlen <= 0 || total_tokens <= 0 ||
      (ldq | ldk | ldvt | ldo) % 8 != 0 || scale <= 0.f ||
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(vt) |
        reinterpret_cast<uintptr_t>(out)) & 15) != 0) {
    set_error("i2r_attention_tc: bad arguments (16-byte aligned pointers, strides multiple of 8, scale > 0)");
    return I2R_E_BADARG;
  }
  if (split && (o_lo % 8 != 0 || o_lo < D)) {
    set_error("i2r_attention_tc: o_lo must be a multiple of 8 and >= D");
    return I2R_E_BADARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (D) {
    case 96:
      return split ? launch_attention_tc<96, true>(q, k, vt, out, ldq, ldk, ldvt, ldo, cu_seqlens, nseq, max_seqlen,
                                                   total_tokens, scale, workspace, workspace_bytes, o_lo, st)
                   : launch_attention_tc<96, false>(q, k, vt, out, ldq, ldk, ldvt, ldo, cu_seqlens, nseq, max_seqlen,
                                                    total_tokens, scale, workspace, workspace_bytes, 0, st);
    case 80:      // d_model 78 padded to 80 (HRFormer-B inter-human stage)
      return split ? launch_attention_tc<80, true>(q, k, vt, out, ldq, ldk, ldvt, ldo, cu_seqlens, nseq, max_seqlen,
                                                   total_tokens, scale, workspace, workspace_bytes, o_lo, st)
                   : launch_attention_tc<80, false>(q, k, vt, out, ldq, ldk, ldvt, ldo, cu_seqlens, nseq, max_seqlen,
                                                    total_tokens, scale, workspace, workspace_bytes, 0, st);
    default:
      set_error("i2r_attention_tc: head dim %d unsupported (80, 96)", D);
      return I2R_E_UNSUPPORTED;
  }
}

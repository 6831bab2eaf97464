// Ragged single-head attention (one sequence per image): flash-style streaming softmax, scores never
// leave the SM.  v1 uses warp-level mma.sync (fp16 in, fp32 accumulate) -- this stage is <2 % of the
// vanilla model's FLOPs; the tcgen05 version is tracked in DESIGN.md ("next").
//
// CTA = 4 warps = 64 query rows of one sequence; K/V stream through a 2-stage cp.async ring in
// 64-key tiles.  Row pitch in shared memory is HD+8 halves so that ldmatrix rows fall in distinct banks.
#include "i2r_common.cuh"

namespace i2r {

constexpr int ATT_BQ = 64;
constexpr int ATT_BK = 64;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD>
__device__ __forceinline__ void load_tile(uint32_t sdst, const __half* __restrict__ g, int ld, int row0, int nrows_valid,
                                          int tid) {
  constexpr int PITCH = (HD + 8) * 2;  // bytes
  constexpr int CPR = HD / 8;          // 16-byte chunks per row
  for (int j = tid; j < ATT_BK * CPR; j += 128) {
    const int r = j / CPR, c = j - r * CPR;
    const bool ok = r < nrows_valid;
    const __half* src = ok ? g + static_cast<int64_t>(row0 + r) * ld + c * 8 : g;
    cp_async16(sdst + r * PITCH + c * 16, src, ok ? 16u : 0u);
  }
}

// SPLIT: split-operand mode (include/i2r.h, I2R_F_SPLIT) -- q/k/v rows carry a lo half `*_lo` elements after the hi
// half; scores = q_hi k_hi + q_hi k_lo + q_lo k_hi, O += P (v_hi + v_lo) with fp16 probabilities, and the output row
// is written as a pair (lo half o_lo elements after the hi half).
template <int HD, bool SPLIT>
__global__ void __launch_bounds__(128) attention_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                        const __half* __restrict__ v, __half* __restrict__ out,
                                                        int ldq, int ldk, int ldv, int ldo,
                                                        const int32_t* __restrict__ cu_seqlens, float scale_log2e,
                                                        int nsplit, float* __restrict__ opart,
                                                        float* __restrict__ mlpart, int q_lo, int k_lo, int v_lo,
                                                        int o_lo, int uniform_len, int head_stride) {
  constexpr int PITCH = (HD + 8) * 2;
  constexpr int TILE = ATT_BK * PITCH;
  constexpr int KS = HD / 16;  // k-steps over the head dim
  constexpr int DT = HD / 8;   // 8-wide output tiles over the head dim
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_launch_dependents();
  pdl_wait();
  const int seq = blockIdx.y;
  // window attention (HRFormer): equal-length sequences without an offset table, gridDim.z = heads (no key split)
  const int t0 = cu_seqlens != nullptr ? cu_seqlens[seq] : seq * uniform_len;
  const int L = cu_seqlens != nullptr ? cu_seqlens[seq + 1] - t0 : uniform_len;
  if (head_stride > 0) {
    const int hoff = blockIdx.z * head_stride;
    q += hoff;
    k += hoff;
    v += hoff;
    out += hoff;
  }
  const int q0 = blockIdx.x * ATT_BQ;
  if (q0 >= L) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + TILE;       // 2 stages
  const uint32_t sV = sK + 2 * TILE;   // 2 stages
  const uint32_t sQl = sV + 2 * TILE;  // SPLIT: lo halves, same arrangement
  const uint32_t sKl = sQl + TILE;
  const uint32_t sVl = sKl + 2 * TILE;

  // split-KV: this CTA covers key tiles [kt_begin, kt_end) of the sequence (all of them when nsplit == 1)
  const int ntiles_all = (L + ATT_BK - 1) / ATT_BK;
  const int per_split = (ntiles_all + nsplit - 1) / nsplit;
  const int kt_begin = (head_stride > 0 ? 0 : blockIdx.z) * per_split;
  const int kt_end = min(ntiles_all, kt_begin + per_split);
  if (kt_begin < kt_end) {
    load_tile<HD>(sQ, q, ldq, t0 + q0, min(ATT_BQ, L - q0), tid);
    load_tile<HD>(sK, k, ldk, t0 + kt_begin * ATT_BK, min(ATT_BK, L - kt_begin * ATT_BK), tid);
    load_tile<HD>(sV, v, ldv, t0 + kt_begin * ATT_BK, min(ATT_BK, L - kt_begin * ATT_BK), tid);
    if (SPLIT) {
      load_tile<HD>(sQl, q + q_lo, ldq, t0 + q0, min(ATT_BQ, L - q0), tid);
      load_tile<HD>(sKl, k + k_lo, ldk, t0 + kt_begin * ATT_BK, min(ATT_BK, L - kt_begin * ATT_BK), tid);
      load_tile<HD>(sVl, v + v_lo, ldv, t0 + kt_begin * ATT_BK, min(ATT_BK, L - kt_begin * ATT_BK), tid);
    }
    cp_async_commit();
  }

  uint32_t qf[KS][4];
  uint32_t qfl[SPLIT ? KS : 1][4];
  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kt = kt_begin; kt < kt_end; ++kt) {
    const int st = (kt - kt_begin) & 1;
    if (kt + 1 < kt_end) {
      const int kr = (kt + 1) * ATT_BK;
      load_tile<HD>(sK + (st ^ 1) * TILE, k, ldk, t0 + kr, min(ATT_BK, L - kr), tid);
      load_tile<HD>(sV + (st ^ 1) * TILE, v, ldv, t0 + kr, min(ATT_BK, L - kr), tid);
      if (SPLIT) {
        load_tile<HD>(sKl + (st ^ 1) * TILE, k + k_lo, ldk, t0 + kr, min(ATT_BK, L - kr), tid);
        load_tile<HD>(sVl + (st ^ 1) * TILE, v + v_lo, ldv, t0 + kr, min(ATT_BK, L - kr), tid);
      }
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == kt_begin) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = ks * 16 + (lane >> 4) * 8;
        ldsm_x4(sQ + row * PITCH + col * 2, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        if (SPLIT) ldsm_x4(sQl + row * PITCH + col * 2, qfl[ks][0], qfl[ks][1], qfl[ks][2], qfl[ks][3]);
      }
    }
    // ---- S = Q K^T for this warp's 16 rows x 64 keys
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    const uint32_t kb = sK + st * TILE;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nt = 0; nt < 8; nt += 2) {
        const int key = nt * 8 + (lane >> 4) * 8 + (lane & 7);
        const int col = ks * 16 + ((lane >> 3) & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kb + key * PITCH + col * 2, b0, b1, b2, b3);
        mma16816(s[nt], qf[ks], b0, b1);
        mma16816(s[nt + 1], qf[ks], b2, b3);
        if (SPLIT) {
          mma16816(s[nt], qfl[ks], b0, b1);          // q_lo k_hi
          mma16816(s[nt + 1], qfl[ks], b2, b3);
          ldsm_x4(sKl + st * TILE + key * PITCH + col * 2, b0, b1, b2, b3);
          mma16816(s[nt], qf[ks], b0, b1);           // q_hi k_lo
          mma16816(s[nt + 1], qf[ks], b2, b3);
        }
      }
    }
    // ---- online softmax (rows r0 = lane/4, r1 = r0+8 of the warp's 16)
    const int kbase = kt * ATT_BK;
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = kbase + nt * 8 + (lane & 3) * 2 + (j & 1);
        float val = s[nt][j] * scale_log2e;
        if (key >= L) val = -INFINITY;
        s[nt][j] = val;
        if (j < 2) mx0 = fmaxf(mx0, val); else mx1 = fmaxf(mx1, val);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);
    m0 = mx0;
    m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mx0);
      s[nt][1] = exp2f(s[nt][1] - mx0);
      s[nt][2] = exp2f(s[nt][2] - mx1);
      s[nt][3] = exp2f(s[nt][3] - mx1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      o[i][0] *= c0;
      o[i][1] *= c0;
      o[i][2] *= c1;
      o[i][3] *= c1;
    }
    // ---- O += P V
    const uint32_t vb = sV + st * TILE;
#pragma unroll
    for (int j = 0; j < ATT_BK / 16; ++j) {
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * j][0], s[2 * j][1]);
      pa[1] = pack_h2(s[2 * j][2], s[2 * j][3]);
      pa[2] = pack_h2(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[3] = pack_h2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int dt = 0; dt < DT; dt += 2) {
        const int key = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = dt * 8 + (lane >> 4) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vb + key * PITCH + col * 2, b0, b1, b2, b3);
        mma16816(o[dt], pa, b0, b1);
        mma16816(o[dt + 1], pa, b2, b3);
        if (SPLIT) {
          ldsm_x4_t(sVl + st * TILE + key * PITCH + col * 2, b0, b1, b2, b3);
          mma16816(o[dt], pa, b0, b1);
          mma16816(o[dt + 1], pa, b2, b3);
        }
      }
    }
    __syncthreads();  // all warps done with stage st before it is refilled
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int r0 = q0 + warp * 16 + (lane >> 2);
  const int r1 = r0 + 8;
  if (nsplit == 1) {
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const int col = dt * 8 + (lane & 3) * 2;
      if (r0 < L) {
        const uint32_t h = pack_h2(o[dt][0] * i0, o[dt][1] * i0);
        *reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(t0 + r0) * ldo + col) = h;
        if (SPLIT) {
          const float2 f = unpack_h2(h);
          *reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(t0 + r0) * ldo + o_lo + col) =
              pack_h2(o[dt][0] * i0 - f.x, o[dt][1] * i0 - f.y);
        }
      }
      if (r1 < L) {
        const uint32_t h = pack_h2(o[dt][2] * i1, o[dt][3] * i1);
        *reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(t0 + r1) * ldo + col) = h;
        if (SPLIT) {
          const float2 f = unpack_h2(h);
          *reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(t0 + r1) * ldo + o_lo + col) =
              pack_h2(o[dt][2] * i1 - f.x, o[dt][3] * i1 - f.y);
        }
      }
    }
  } else {
    // un-normalised partial result + (running max, running sum) for the merge kernel
    const int sp = blockIdx.z;
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const int col = dt * 8 + (lane & 3) * 2;
      if (r0 < L)
        *reinterpret_cast<float2*>(opart + (static_cast<int64_t>(t0 + r0) * nsplit + sp) * HD + col) = make_float2(o[dt][0], o[dt][1]);
      if (r1 < L)
        *reinterpret_cast<float2*>(opart + (static_cast<int64_t>(t0 + r1) * nsplit + sp) * HD + col) = make_float2(o[dt][2], o[dt][3]);
    }
    if ((lane & 3) == 0) {
      if (r0 < L) *reinterpret_cast<float2*>(mlpart + (static_cast<int64_t>(t0 + r0) * nsplit + sp) * 2) = make_float2(m0, l0);
      if (r1 < L) *reinterpret_cast<float2*>(mlpart + (static_cast<int64_t>(t0 + r1) * nsplit + sp) * 2) = make_float2(m1, l1);
    }
  }
}

// Merge the split-KV partials: out = sum_s O_s 2^(m_s - M) / sum_s l_s 2^(m_s - M).  One warp per token row.
template <int HD>
__global__ void __launch_bounds__(256) attention_merge_kernel(const float* __restrict__ opart,
                                                              const float* __restrict__ mlpart,
                                                              __half* __restrict__ out, int ldo, int rows, int nsplit,
                                                              int o_lo) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, mlpart[(static_cast<int64_t>(row) * nsplit + s) * 2]);
  float den = 0.f;
  float acc[(HD + 31) / 32];
#pragma unroll
  for (int i = 0; i < (HD + 31) / 32; ++i) acc[i] = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float2 ml = *reinterpret_cast<const float2*>(mlpart + (static_cast<int64_t>(row) * nsplit + s) * 2);
    const float wgt = (ml.x == -INFINITY) ? 0.f : exp2f(ml.x - M);
    den += ml.y * wgt;
#pragma unroll
    for (int i = 0; i < (HD + 31) / 32; ++i) {
      const int c = lane + 32 * i;
      if (c < HD) acc[i] += wgt * opart[(static_cast<int64_t>(row) * nsplit + s) * HD + c];
    }
  }
  const float inv = 1.f / den;
#pragma unroll
  for (int i = 0; i < (HD + 31) / 32; ++i) {
    const int c = lane + 32 * i;
    if (c < HD) {
      const __half h = __float2half_rn(acc[i] * inv);
      out[static_cast<int64_t>(row) * ldo + c] = h;
      if (o_lo > 0) out[static_cast<int64_t>(row) * ldo + o_lo + c] = __float2half_rn(acc[i] * inv - __half2float(h));
    }
  }
}

// shared with attention_tc.cu (same partial format: un-normalised fp32 O, (max, sum) in the log2 domain)
int attention_merge_launch(int HD, const float* opart, const float* mlpart, __half* out, int ldo, int rows, int nsplit,
                           int o_lo, cudaStream_t st) {
  const dim3 grid((rows + 7) / 8), block(256);
  switch (HD) {
    case 96:
      launch_pdl(attention_merge_kernel<96>, grid, block, 0, st, opart, mlpart, out, ldo, rows, nsplit, o_lo);
      break;
    case 80:
      launch_pdl(attention_merge_kernel<80>, grid, block, 0, st, opart, mlpart, out, ldo, rows, nsplit, o_lo);
      break;
    default:
      set_error("attention_merge: head dim %d unsupported", HD);
      return I2R_E_UNSUPPORTED;
  }
  return check_launch("attention_merge_kernel");
}

static int choose_nsplit(int nseq, int max_seqlen) {
  // enough CTAs for ~2 waves of 148 SMs, never more splits than key tiles, at most 8
  const int qtiles = (max_seqlen + ATT_BQ - 1) / ATT_BQ;
  const int ktiles = (max_seqlen + ATT_BK - 1) / ATT_BK;
  int ns = (2 * 148 + qtiles * nseq - 1) / (qtiles * nseq);
  if (ns > ktiles) ns = ktiles;
  if (ns > 8) ns = 8;
  if (ns < 1) ns = 1;
  return ns;
}

template <int HD, bool SPLIT>
static int launch_attention(const void* q, const void* k, const void* v, void* out, int ldq, int ldk, int ldv, int ldo,
                            const int32_t* cu, int nseq, int max_seqlen, int total_tokens, float scale, void* ws,
                            int64_t ws_bytes, int q_lo, int k_lo, int v_lo, int o_lo, cudaStream_t st) {
  constexpr int smem = (SPLIT ? 10 : 5) * ATT_BK * (HD + 8) * 2;
  static bool attr_done_dev[MAX_DEVICES] = {};   // the opt-in is a per-device property
  bool& attr_done = attr_done_dev[current_device()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<HD, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_done = true;
  }
  int nsplit = choose_nsplit(nseq, max_seqlen);
  const int64_t need = static_cast<int64_t>(total_tokens) * nsplit * (HD + 2) * 4;
  if (nsplit > 1 && (ws == nullptr || ws_bytes < need)) nsplit = 1;   // no workspace: single pass
  float* opart = static_cast<float*>(ws);
  float* mlpart = opart ? opart + static_cast<int64_t>(total_tokens) * nsplit * HD : nullptr;
  dim3 grid((max_seqlen + ATT_BQ - 1) / ATT_BQ, nseq, nsplit);
  launch_pdl(attention_kernel<HD, SPLIT>, grid, dim3(128), smem, st,
      static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(v),
      static_cast<__half*>(out), ldq, ldk, ldv, ldo, cu, scale * 1.4426950408889634f, nsplit, opart, mlpart, q_lo, k_lo,
      v_lo, o_lo, 0, 0);
  int rc = check_launch("attention_kernel");
  if (rc || nsplit == 1) return rc;
  launch_pdl(attention_merge_kernel<HD>, dim3((total_tokens + 7) / 8), dim3(256), 0, st, opart, mlpart, static_cast<__half*>(out), ldo,
                                                                     total_tokens, nsplit, SPLIT ? o_lo : 0);
  return check_launch("attention_merge_kernel");
}

}  // namespace i2r

extern "C" int64_t i2r_attention_workspace_bytes(int total_tokens, int D, int nseq, int max_seqlen) {
  const int ns = i2r::choose_nsplit(nseq, max_seqlen);
  return ns > 1 ? static_cast<int64_t>(total_tokens) * ns * (D + 2) * 4 : 0;
}

extern "C" int i2r_attention_varlen(const void* q, const void* k, const void* v, void* out, int ldq, int ldk, int ldv,
                                    int ldo, int D, const int32_t* cu_seqlens, int nseq, int max_seqlen,
                                    int total_tokens, float scale, void* workspace, int64_t workspace_bytes,
                                    int split, int q_lo, int k_lo, int v_lo, int o_lo, void* stream) {
  using namespace i2r;
  if (!q || !k || !v || !out || !cu_seqlens || nseq <= 0 || max_seqlen <= 0 || total_tokens <= 0 ||
      (ldq | ldk | ldv | ldo) % 8 != 0) {
    set_error("i2r_attention_varlen: bad arguments");
    return I2R_E_BADARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (split && ((q_lo | k_lo | v_lo | o_lo) % 8 != 0 || o_lo < D)) {
    set_error("i2r_attention_varlen: split-operand offsets must be multiples of 8 (o_lo >= D)");
    return I2R_E_BADARG;
  }
#define I2R_ATT_CASE(HD_)                                                                                              \
  case HD_:                                                                                                            \
    return split ? launch_attention<HD_, true>(q, k, v, out, ldq, ldk, ldv, ldo, cu_seqlens, nseq, max_seqlen,           \
                                               total_tokens, scale, workspace, workspace_bytes, q_lo, k_lo, v_lo, o_lo, \
                                               st)                                                                      \
                 : launch_attention<HD_, false>(q, k, v, out, ldq, ldk, ldv, ldo, cu_seqlens, nseq, max_seqlen,          \
                                                total_tokens, scale, workspace, workspace_bytes, 0, 0, 0, 0, st);
  switch (D) {
    I2R_ATT_CASE(96)
    I2R_ATT_CASE(80)
    default:
      set_error("i2r_attention_varlen: head dim %d unsupported (80 or 96)", D);
      return I2R_E_UNSUPPORTED;
  }
}


// Window attention of HRFormer-B (InterlacedPoolAttention / MHA_, lib/models/hrformer.py:1164-1180, :627-935): nwin
// windows of `win_len` (49) consecutive token rows each (window-major layout written by i2r_ln_window_gather, padded
// tokens included), `heads` heads of `head_pad` (48) channels each (head_dim 39 zero-padded by the packing of the
// q/k/v projections); softmax(scale q k^T) v per (window, head), no relative position bias (:866-888), no mask.
namespace i2r {
template <bool SPLIT>
static int launch_window_attention(const void* q, const void* k, const void* v, void* out, int ldq, int ldk, int ldv,
                                   int ldo, int nwin, int win_len, int heads, float scale, int q_lo, int k_lo, int v_lo,
                                   int o_lo, cudaStream_t st) {
  constexpr int HD = 48;
  constexpr int smem = (SPLIT ? 10 : 5) * ATT_BK * (HD + 8) * 2;
  static bool attr_done_dev[MAX_DEVICES] = {};   // the opt-in is a per-device property
  bool& attr_done = attr_done_dev[current_device()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<HD, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(window attention): %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_done = true;
  }
  dim3 grid((win_len + ATT_BQ - 1) / ATT_BQ, nwin, heads);
  launch_pdl(attention_kernel<HD, SPLIT>, grid, dim3(128), smem, st, static_cast<const __half*>(q),
             static_cast<const __half*>(k), static_cast<const __half*>(v), static_cast<__half*>(out), ldq, ldk, ldv, ldo,
             static_cast<const int32_t*>(nullptr), scale * 1.4426950408889634f, 1, static_cast<float*>(nullptr),
             static_cast<float*>(nullptr), q_lo, k_lo, v_lo, o_lo, win_len, HD);
  return check_launch("attention_kernel(window)");
}
}  // namespace i2r

extern "C" int i2r_window_attention(const void* q, const void* k, const void* v, void* out, int ldq, int ldk, int ldv,
                                    int ldo, int nwin, int win_len, int heads, int head_pad, float scale, int split,
                                    int q_lo, int k_lo, int v_lo, int o_lo, void* stream) {
  using namespace i2r;
  if (!q || !k || !v || !out || nwin <= 0 || win_len <= 0 || win_len > 64 * 1024 || heads <= 0 || head_pad != 48 ||
      heads > 65535 || (ldq | ldk | ldv | ldo) % 8 != 0 ||
      (split && (q_lo | k_lo | v_lo | o_lo) % 8 != 0)) {
    set_error("i2r_window_attention: bad arguments (head_pad must be 48, strides multiples of 8)");
    return I2R_E_BADARG;
  }
  if (nwin > 65535) {
    set_error("i2r_window_attention: more than 65535 windows per launch (grid.y); split the batch");
    return I2R_E_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return split ? launch_window_attention<true>(q, k, v, out, ldq, ldk, ldv, ldo, nwin, win_len, heads, scale, q_lo, k_lo,
                                               v_lo, o_lo, st)
               : launch_window_attention<false>(q, k, v, out, ldq, ldk, ldv, ldo, nwin, win_len, heads, scale, 0, 0, 0, 0,
                                                st);
}

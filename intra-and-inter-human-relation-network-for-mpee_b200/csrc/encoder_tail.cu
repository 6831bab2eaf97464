// Fused tail of one post-norm Transformer encoder layer (everything after the attention), one 128-token tile per CTA:
//     x1 = attn W_o^T + b_o + src          s1 = LayerNorm1(x1)
//     h  = relu(s1 W_1^T + b_1)            x2 = h W_2^T + b_2 + s1
//     src' = LayerNorm2(x2)                sp' = src' + pos   (q/k input of the next layer; optional)
// Reference: TransformerEncoderLayer.forward_post (lib/models/transpose_h.py:205-222; lib/models/attention.py:61-82;
// lib/models/interformer_pureMulti.py:182-213) with d_model 96, dim_feedforward 192, eval-mode dropout = identity.
// Replaces five launches (out-proj GEMM, LayerNorm, two FFN GEMMs, LayerNorm) and their HBM round trips.
//
// All five GEMMs are [128 x 96] x [96 -> 96] on tcgen05 (the FFN is evaluated as two 96-channel halves of the hidden
// layer: acc2 = h_a W_2a^T + h_b W_2b^T), accumulators in TMEM, A operands in shared memory as SWIZZLE_128B rows that
// the epilogue threads write themselves (one thread = one token row, so both LayerNorms are register-local), weights
// streamed by 1-D bulk copies through a ring of three slots from a pre-swizzled image (i2r_b200/packing.py
// pack_encoder_tail).  Warps: 0 = TMA producer, 1 = MMA issuer, 2-5 = epilogue (TMEM lane quadrant = warp % 4).
// Split-operand mode: activations and weights are (hi | lo) pairs along K and every K step issues the three MMAs
// x_hi W_hi + x_lo W_hi + x_hi W_lo, as everywhere else (include/i2r.h, I2R_F_SPLIT).
#include "i2r_tma.cuh"

namespace i2r {

constexpr int ET_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps
constexpr int ET_HC = 48;          // channels per epilogue thread (two threads share a token row)
constexpr int ET_D = 96;         // d_model
constexpr int ET_F = 192;        // dim_feedforward
constexpr int ET_NPARAM = 768;   // fp32: b_o[96] b_1[192] b_2[96] g1[96] be1[96] g2[96] be2[96]

template <bool SPLIT>
struct EtCfg {
  static constexpr int NCH = SPLIT ? 3 : 2;                 // 64-channel chunks of a (hi | lo) row of 96 channels
  static constexpr int A_BYTES = NCH * TC_CH_BYTES;         // activation operand tile: 128 rows
  static constexpr int W_CH = ET_D * 128;                   // one weight chunk: 96 rows x 128 B
  static constexpr int W_BYTES = NCH * W_CH;                // one [96 x 96] matrix
  static constexpr int SMEM = 2 * A_BYTES + 3 * W_BYTES + ET_NPARAM * 4 + 2048 + 256 + 1024;
};

struct EtArgs {
  const __half* wimg;    // [5][W_BYTES] fp16 pre-swizzled: W_o, W_1[0:96], W_1[96:192], W_2[:, 0:96], W_2[:, 96:192]
  const float* params;   // [ET_NPARAM]
  const __half* pos;     // optional [T, ld]
  __half* out;           // [T, ld] src'
  __half* out_pos;       // [T, ld] src' + pos (only with pos)
  int T, ld;             // tokens, row stride of pos / out / out_pos in elements
  float eps;
};

template <bool SPLIT>
__device__ __forceinline__ void issue_gemm96(uint32_t d_tmem, uint32_t a_addr, uint32_t w_addr, uint32_t idesc,
                                             uint32_t acc_first) {
  constexpr int KS = ET_D / 16;
  constexpr int WCH = ET_D * 128;
  const uint32_t hi = sw128_desc_hi(1024, 0);
  const uint32_t a0 = sw128_desc_lo(a_addr), b0 = sw128_desc_lo(w_addr);
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    umma_f16(d_tmem, desc64(a0 + (kstep_off(s) >> 4), hi), desc64(b0 + (kstep_off(s, WCH) >> 4), hi), idesc,
             s ? 1u : acc_first);
    if (SPLIT) {
      umma_f16(d_tmem, desc64(a0 + (kstep_off(KS + s) >> 4), hi), desc64(b0 + (kstep_off(s, WCH) >> 4), hi), idesc, 1u);
      umma_f16(d_tmem, desc64(a0 + (kstep_off(s) >> 4), hi), desc64(b0 + (kstep_off(KS + s, WCH) >> 4), hi), idesc, 1u);
    }
  }
}

// shared-memory byte offset of the 16-byte group holding channels [8g, 8g+8) of the (hi | lo) row `row`
// (g counts over the pair row: hi groups 0..11, lo groups 12..23)
__device__ __forceinline__ uint32_t a_group_off(int row, int g) {
  return (g >> 3) * TC_CH_BYTES + sw128_off(row, g & 7);
}

// write 16 consecutive channels (values v[0..15] of channel c0..c0+15) of this thread's row into an A operand tile
template <bool SPLIT>
__device__ __forceinline__ void store_a16(uint32_t tile, int row, int c0, const float (&v)[16]) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t h[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = v[half * 8 + 2 * i], b = v[half * 8 + 2 * i + 1];
      h[i] = pack_h2(a, b);
      if (SPLIT) {
        const float2 f = unpack_h2(h[i]);
        lo[i] = pack_h2(a - f.x, b - f.y);
      }
    }
    const int g = (c0 >> 3) + half;
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tile + a_group_off(row, g)), "r"(h[0]), "r"(h[1]),
                 "r"(h[2]), "r"(h[3])
                 : "memory");
    if (SPLIT)
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tile + a_group_off(row, g + ET_D / 8)), "r"(lo[0]),
                   "r"(lo[1]), "r"(lo[2]), "r"(lo[3])
                   : "memory");
  }
}

// read 16 consecutive channels of this thread's row from an A operand tile (hi + lo in split mode)
template <bool SPLIT>
__device__ __forceinline__ void load_a16(uint32_t tile, int row, int c0, float (&v)[16]) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int g = (c0 >> 3) + half;
    uint32_t h[4];
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3])
                 : "r"(tile + a_group_off(row, g)));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_h2(h[i]);
      v[half * 8 + 2 * i] = f.x;
      v[half * 8 + 2 * i + 1] = f.y;
    }
    if (SPLIT) {
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3])
                   : "r"(tile + a_group_off(row, g + ET_D / 8)));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_h2(h[i]);
        v[half * 8 + 2 * i] += f.x;
        v[half * 8 + 2 * i + 1] += f.y;
      }
    }
  }
}

__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// x[i] += p[i] for this thread's 48 channels (p: fp32 parameter vector in shared memory, 16-byte aligned)
__device__ __forceinline__ void add_param48(float (&x)[ET_HC], uint32_t p) {
#pragma unroll
  for (int i = 0; i < ET_HC / 4; ++i) {
    const float4 b = lds_f32x4(p + 16 * i);
    x[4 * i] += b.x;
    x[4 * i + 1] += b.y;
    x[4 * i + 2] += b.z;
    x[4 * i + 3] += b.w;
  }
}

// LayerNorm over the 96 channels of a row held by two threads (48 each): two-pass fp32 statistics, the halves
// exchanged through `scratch` (float [2][128]) around a named barrier of the 256 epilogue threads.
__device__ __forceinline__ void layernorm_pair(float (&x)[ET_HC], uint32_t gamma, uint32_t beta, float eps,
                                               uint32_t scratch, int row, int hsel) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < ET_HC; i += 4) {
    s0 += x[i];
    s1 += x[i + 1];
    s2 += x[i + 2];
    s3 += x[i + 3];
  }
  const float part = (s0 + s1) + (s2 + s3);
  sts_f32(scratch + 4 * (hsel * 128 + row), part);
  epi_bar();
  const float mean = (part + lds_f32(scratch + 4 * ((hsel ^ 1) * 128 + row))) * (1.f / ET_D);
  s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
  for (int i = 0; i < ET_HC; i += 4) {
    const float d0 = x[i] - mean, d1 = x[i + 1] - mean, d2 = x[i + 2] - mean, d3 = x[i + 3] - mean;
    s0 += d0 * d0;
    s1 += d1 * d1;
    s2 += d2 * d2;
    s3 += d3 * d3;
  }
  const float partq = (s0 + s1) + (s2 + s3);
  sts_f32(scratch + 4 * (256 + hsel * 128 + row), partq);
  epi_bar();
  const float rstd = rsqrtf((partq + lds_f32(scratch + 4 * (256 + (hsel ^ 1) * 128 + row))) * (1.f / ET_D) + eps);
#pragma unroll
  for (int i = 0; i < ET_HC / 4; ++i) {
    const float4 g = lds_f32x4(gamma + 16 * i), b = lds_f32x4(beta + 16 * i);
    x[4 * i] = (x[4 * i] - mean) * rstd * g.x + b.x;
    x[4 * i + 1] = (x[4 * i + 1] - mean) * rstd * g.y + b.y;
    x[4 * i + 2] = (x[4 * i + 2] - mean) * rstd * g.z + b.z;
    x[4 * i + 3] = (x[4 * i + 3] - mean) * rstd * g.w + b.w;
  }
}

template <bool SPLIT>
__global__ void __launch_bounds__(ET_THREADS, 1)
encoder_tail_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapS, const EtArgs E) {
  using C = EtCfg<SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tok0 = blockIdx.x * 128;

  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = sbase;                    // attention output, then s1
  const uint32_t sH = sA + C::A_BYTES;          // src (residual of x1), then the hidden halves
  const uint32_t sW = sH + C::A_BYTES;          // 3 weight slots
  const uint32_t sPar = sW + 3 * C::W_BYTES;
  const uint32_t sScr = sPar + ET_NPARAM * 4;   // LayerNorm exchange: float [2 stats][2 halves][128 rows]
  const uint32_t sBar = sScr + 2048;
  const uint32_t bA = sBar, bSrc = sBar + 8;
  const uint32_t bW = sBar + 16;      // [3] weight slot full
  const uint32_t bWe = sBar + 40;     // [3] weight slot free
  const uint32_t bAcc1 = sBar + 64, bS1 = sBar + 72, bH0 = sBar + 80, bH1 = sBar + 88, bHfull = sBar + 96,
                 bHfree = sBar + 104, bAcc2 = sBar + 112;
  const uint32_t sSlot = sBar + 128;
  if (tid == 0) {
    mbar_init(bA, 1);
    mbar_init(bSrc, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(bW + 8 * i, 1);
      mbar_init(bWe + 8 * i, 1);
    }
    mbar_init(bAcc1, 1);
    mbar_init(bS1, 256);
    mbar_init(bH0, 1);
    mbar_init(bH1, 1);
    mbar_init(bHfull, 256);
    mbar_init(bHfree, 1);
    mbar_init(bAcc2, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(sSlot, 512);
    tmem_relinquish();
  }
  // layer constants: parameters never change between launches, so they may be read before the dependency wait
  for (int i = tid; i < ET_NPARAM; i += ET_THREADS)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sPar + 4 * i), "f"(E.params[i]) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sSlot));
  const uint32_t tAcc1 = tmem_base, tH0 = tmem_base + 128, tH1 = tmem_base + 256, tAcc2 = tmem_base + 384;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------- producer
    if (elect_one()) {
      const uint8_t* wimg = reinterpret_cast<const uint8_t*>(E.wimg);
      // weights first (constants), activations after the dependency wait
      for (int i = 0; i < 3; ++i) {
        mbar_arrive_expect_tx(bW + 8 * i, C::W_BYTES);
        bulk_g2s(sW + i * C::W_BYTES, wimg + static_cast<size_t>(i) * C::W_BYTES, C::W_BYTES, bW + 8 * i);
      }
      pdl_wait();
      mbar_arrive_expect_tx(bA, C::A_BYTES);
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch) tma_load_2d(sA + ch * TC_CH_BYTES, &mapA, ch * 64, tok0, bA);
      mbar_arrive_expect_tx(bSrc, C::A_BYTES);
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch) tma_load_2d(sH + ch * TC_CH_BYTES, &mapS, ch * 64, tok0, bSrc);
      for (int i = 3; i < 5; ++i) {
        const int slot = i - 3;
        mbar_wait_relaxed(bWe + 8 * slot, 0);
        mbar_arrive_expect_tx(bW + 8 * slot, C::W_BYTES);
        bulk_g2s(sW + slot * C::W_BYTES, wimg + static_cast<size_t>(i) * C::W_BYTES, C::W_BYTES, bW + 8 * slot);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      const uint32_t id = make_idesc_f16(128, ET_D);
      mbar_wait(bA, 0);
      mbar_wait(bW, 0);
      tc_fence_after();
      issue_gemm96<SPLIT>(tAcc1, sA, sW, id, 0u);                       // attn W_o^T
      umma_commit(bAcc1);
      umma_commit(bWe);
      mbar_wait(bS1, 0);
      mbar_wait(bW + 8, 0);
      tc_fence_after();
      issue_gemm96<SPLIT>(tH0, sA, sW + C::W_BYTES, id, 0u);            // s1 W_1a^T
      umma_commit(bH0);
      umma_commit(bWe + 8);
      mbar_wait(bW + 16, 0);
      tc_fence_after();
      issue_gemm96<SPLIT>(tH1, sA, sW + 2 * C::W_BYTES, id, 0u);        // s1 W_1b^T
      umma_commit(bH1);
      mbar_wait(bHfull, 0);
      mbar_wait(bW, 1);
      tc_fence_after();
      issue_gemm96<SPLIT>(tAcc2, sH, sW, id, 0u);                       // h_a W_2a^T
      umma_commit(bHfree);
      mbar_wait(bHfull, 1);
      mbar_wait(bW + 8, 1);
      tc_fence_after();
      issue_gemm96<SPLIT>(tAcc2, sH, sW + C::W_BYTES, id, 1u);          // + h_b W_2b^T
      umma_commit(bAcc2);
    }
  } else {
    // ------------------------------------------------------------------------------------- epilogues
    // two threads per token row: warps 2-5 take channels [0, 48), warps 6-9 channels [48, 96) of every vector
    const int quad = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int ch0 = hsel * ET_HC;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (static_cast<uint32_t>(quad * 32) << 16) + ch0;
    const int tok = tok0 + row;
    const bool ok = tok < E.T;
    const uint32_t pBo = sPar + 4 * ch0, pB1 = sPar + 4 * (96 + ch0), pB2 = sPar + 4 * (288 + ch0),
                   pG1 = sPar + 4 * (384 + ch0), pBe1 = sPar + 4 * (480 + ch0), pG2 = sPar + 4 * (576 + ch0),
                   pBe2 = sPar + 4 * (672 + ch0);
    float x[ET_HC];
    // ---- x1 = acc1 + b_o + src ; s1 = LN1(x1) -> sA
    mbar_wait(bSrc, 0);
    mbar_wait(bAcc1, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < ET_HC / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(tAcc1 + lane_off + c * 16, r);
      float v[16];
      load_a16<SPLIT>(sH, row, ch0 + c * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) x[c * 16 + i] = __uint_as_float(r[i]) + v[i];
    }
    add_param48(x, pBo);
    layernorm_pair(x, pG1, pBe1, E.eps, sScr, row, hsel);
#pragma unroll
    for (int c = 0; c < ET_HC / 16; ++c) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = x[c * 16 + i];
      store_a16<SPLIT>(sA, row, ch0 + c * 16, v);     // GEMM 1 has completed (bAcc1), so the attention tile is dead
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(bS1);
    // ---- hidden halves: h = relu(acc + b_1) -> sH.  The src tile in sH is dead once BOTH threads of every row have
    //      read it, which the LayerNorm barriers above already guarantee.
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      mbar_wait(hf ? bH1 : bH0, 0);
      tc_fence_after();
      if (hf) {
        mbar_wait(bHfree, 0);     // h_a W_2a^T has consumed the buffer
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < ET_HC / 16; ++c) {
        uint32_t r[16];
        tmem_ld16((hf ? tH1 : tH0) + lane_off + c * 16, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b = lds_f32x4(pB1 + 4 * (hf * ET_D + c * 16 + 4 * i));
          v[4 * i] = fmaxf(__uint_as_float(r[4 * i]) + b.x, 0.f);
          v[4 * i + 1] = fmaxf(__uint_as_float(r[4 * i + 1]) + b.y, 0.f);
          v[4 * i + 2] = fmaxf(__uint_as_float(r[4 * i + 2]) + b.z, 0.f);
          v[4 * i + 3] = fmaxf(__uint_as_float(r[4 * i + 3]) + b.w, 0.f);
        }
        store_a16<SPLIT>(sH, row, ch0 + c * 16, v);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bHfull);
    }
    // ---- x2 = acc2 + b_2 + s1 ; src' = LN2(x2).  The pos row (a constant of the model) is fetched first so that its
    //      global-memory latency hides behind the last GEMM.
    uint4 ph[ET_HC / 8], pl[SPLIT ? ET_HC / 8 : 1];
    if (E.pos != nullptr && ok) {
      const __half* prow = E.pos + static_cast<int64_t>(tok) * E.ld + ch0;
#pragma unroll
      for (int g = 0; g < ET_HC / 8; ++g) {
        ph[g] = __ldg(reinterpret_cast<const uint4*>(prow + g * 8));
        if (SPLIT) pl[g] = __ldg(reinterpret_cast<const uint4*>(prow + ET_D + g * 8));
      }
    }
    mbar_wait(bAcc2, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < ET_HC / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(tAcc2 + lane_off + c * 16, r);
      float v[16];
      load_a16<SPLIT>(sA, row, ch0 + c * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) x[c * 16 + i] = __uint_as_float(r[i]) + v[i];
    }
    add_param48(x, pB2);
    layernorm_pair(x, pG2, pBe2, E.eps, sScr, row, hsel);
    if (ok) {
      __half* orow = E.out + static_cast<int64_t>(tok) * E.ld + ch0;
      const bool prow = E.pos != nullptr;
      __half* qrow = E.pos ? E.out_pos + static_cast<int64_t>(tok) * E.ld + ch0 : nullptr;
#pragma unroll
      for (int g = 0; g < ET_HC / 8; ++g) {
        uint32_t h[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = x[g * 8 + 2 * i], b = x[g * 8 + 2 * i + 1];
          h[i] = pack_h2(a, b);
          if (SPLIT) {
            const float2 f = unpack_h2(h[i]);
            lo[i] = pack_h2(a - f.x, b - f.y);
          }
        }
        *reinterpret_cast<uint4*>(orow + g * 8) = make_uint4(h[0], h[1], h[2], h[3]);
        if (SPLIT) *reinterpret_cast<uint4*>(orow + ET_D + g * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        if (prow) {
          const uint32_t pw[4] = {ph[g].x, ph[g].y, ph[g].z, ph[g].w};
          float s[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = unpack_h2(pw[i]);
            s[2 * i] = x[g * 8 + 2 * i] + f.x;
            s[2 * i + 1] = x[g * 8 + 2 * i + 1] + f.y;
          }
          if (SPLIT) {
            const uint32_t pv[4] = {pl[g].x, pl[g].y, pl[g].z, pl[g].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = unpack_h2(pv[i]);
              s[2 * i] += f.x;
              s[2 * i + 1] += f.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            h[i] = pack_h2(s[2 * i], s[2 * i + 1]);
            if (SPLIT) {
              const float2 f = unpack_h2(h[i]);
              lo[i] = pack_h2(s[2 * i] - f.x, s[2 * i + 1] - f.y);
            }
          }
          *reinterpret_cast<uint4*>(qrow + g * 8) = make_uint4(h[0], h[1], h[2], h[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(qrow + ET_D + g * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <bool SPLIT>
static int launch_encoder_tail(const void* attn, int ld_attn, const void* src, int ld_src, const void* pos, void* out,
                               void* out_pos, int ld, const void* wimg, const float* params, int T, float eps,
                               cudaStream_t st) {
  using C = EtCfg<SPLIT>;
  static bool attr_done_dev[MAX_DEVICES] = {};   // the opt-in is a per-device property
  bool& attr_done = attr_done_dev[current_device()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(encoder_tail_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(encoder_tail): %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_done = true;
  }
  CUtensorMap ma, ms;
  const uint64_t width = SPLIT ? 2 * ET_D : ET_D;
  int rc = encode_2d(&ma, attn, width, T, ld_attn, 64, 128);
  if (!rc) rc = encode_2d(&ms, src, width, T, ld_src, 64, 128);
  if (rc) return rc;
  EtArgs E;
  E.wimg = static_cast<const __half*>(wimg);
  E.params = params;
  E.pos = static_cast<const __half*>(pos);
  E.out = static_cast<__half*>(out);
  E.out_pos = static_cast<__half*>(out_pos);
  E.T = T;
  E.ld = ld;
  E.eps = eps;
  launch_pdl(encoder_tail_kernel<SPLIT>, dim3((T + 127) / 128), dim3(ET_THREADS), static_cast<size_t>(C::SMEM), st, ma,
             ms, E);
  return check_launch("encoder_tail_kernel");
}

I2R_HANG_SINK_SETTER(encoder_tail)
}  // namespace i2r

extern "C" int64_t i2r_encoder_tail_weight_bytes(int split) {
  return 5ll * (split ? i2r::EtCfg<true>::W_BYTES : i2r::EtCfg<false>::W_BYTES);
}

extern "C" int i2r_encoder_tail(const void* attn, int ld_attn, const void* src, int ld_src, const void* pos, void* out,
                                void* out_pos, int ld, const void* wimg, const float* params, int T, int d_model,
                                int dim_ff, float eps, int split, void* stream) {
  using namespace i2r;
  if (!attn || !src || !out || !wimg || !params || T <= 0 || (pos && !out_pos) || (ld_attn | ld_src | ld) % 8 != 0 ||
      ((reinterpret_cast<uintptr_t>(attn) | reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out) |
        reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(out_pos) | reinterpret_cast<uintptr_t>(wimg)) &
       15) != 0) {
    set_error("i2r_encoder_tail: bad arguments (16-byte aligned pointers, strides multiple of 8)");
    return I2R_E_BADARG;
  }
  if (d_model != ET_D || dim_ff != ET_F) {
    set_error("i2r_encoder_tail: d_model %d / dim_feedforward %d unsupported (96 / 192)", d_model, dim_ff);
    return I2R_E_UNSUPPORTED;
  }
  const int width = split ? 2 * ET_D : ET_D;
  if (ld < width || ld_attn < width || ld_src < width) {
    set_error("i2r_encoder_tail: row strides must cover %d channels", width);
    return I2R_E_BADARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return split ? launch_encoder_tail<true>(attn, ld_attn, src, ld_src, pos, out, out_pos, ld, wimg, params, T, eps, st)
               : launch_encoder_tail<false>(attn, ld_attn, src, ld_src, pos, out, out_pos, ld, wimg, params, T, eps, st);
}

// Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (sm_100a).
//
// One CTA computes a tile of 128 output pixels x Npad output channels:
//   warps 0-3 (128 threads): A-operand producers (im2col gather with cp.async, zero-fill for padding
//                            taps) and, afterwards, the epilogue (TMEM -> registers -> global);
//   warp 4  : TMEM allocation + the single MMA-issuing thread (tcgen05.mma, accumulator in TMEM).
// The B operand (weights, pre-packed on the host in the exact shared-memory core-matrix layout)
// arrives with one 1-D bulk copy (TMA unit) per pipeline stage.  Stages form an mbarrier ring
// (full[]: 128 producer arrivals + 1 expect_tx arrival; empty[]: tcgen05.commit).
//
// Shared-memory operand layout: SWIZZLE_128B K-major -- one 128-byte row per output pixel (A) or per
// output channel (B) holding the K-chunk's channels (48 or 64 real, zero-padded to 64 slots for B; A's pad
// slots are never read because only KC/16 K-steps are issued); the 16-byte chunks of a row are XOR-ed
// with (row & 7), which also makes the producers' cp.async writes bank-conflict free.
#include <cstdio>

#include "i2r_common.cuh"

namespace i2r {

struct ConvGroup {
  int nprob;
  int tile_end[I2R_MAX_GROUP];  // exclusive prefix of 128-pixel tiles per problem
  i2r_conv_problem p[I2R_MAX_GROUP];
};

constexpr int BM = 128;
constexpr int NPROD = 128;
constexpr int NTHREADS = 160;
constexpr int A_STAGE = BM * 128;           // one 128-byte swizzled row per output pixel
constexpr int SMEM_HDR = 3072;             // barriers (<=128 B) | tmem ptr | scale[256] | bias[256]; stages 1024-aligned
constexpr int SMEM_SCALE_OFF = 256;
constexpr int SMEM_BIAS_OFF = 256 + 1024;


template <int STAGES>
__device__ __forceinline__ void run_tile(const i2r_conv_problem& P, const int tile, uint8_t* smem) {
  constexpr int LOOK = STAGES > 2 ? STAGES - 2 : 1;   // cp.async groups in flight; the other stage(s) cover the MMA hand-over round trip
  constexpr int KG = 8;   // 16-byte chunk slots per 128-byte row; the last K-chunk may use fewer (Cin % 64 != 0)
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase;                  // STAGES x 8 B
  const uint32_t bar_empty = sbase + 8 * STAGES;    // STAGES x 8 B
  const uint32_t bar_accum = sbase + 16 * STAGES;   // 8 B
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16 * STAGES + 8);
  float* s_scale = reinterpret_cast<float*>(smem + SMEM_SCALE_OFF);
  float* s_bias = reinterpret_cast<float*>(smem + SMEM_BIAS_OFF);

  const int Npad = P.Npad;
  const int a_bytes = A_STAGE;
  const int b_bytes = Npad * 128;
  const int st_bytes = a_bytes + b_bytes;
  const uint32_t stages0 = sbase + SMEM_HDR;

  // split-operand mode: K runs over [x_hi | x_lo | x_hi] (three passes over the nkr real K-chunks; the input tensor
  // stores hi at channel 0 and lo at channel Cin) against weights packed as [W_hi | W_hi | W_lo]
  const bool split = (P.flags & I2R_F_SPLIT) != 0;
  const int nkr = (P.Cin + 63) >> 6;
  const int nchunks = split ? 3 * nkr : nkr;
  const int niter = P.ntaps * nchunks;
  const int M = P.NB * P.OH * P.OW;

  uint32_t ncols = 32;
  while (ncols < static_cast<uint32_t>(Npad)) ncols <<= 1;

  // ---------------- setup
  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, NPROD + 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_accum, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), ncols);
    tmem_relinquish();
  }
  if (tid < NPROD) {
    for (int i = tid; i < Npad; i += NPROD) {
      s_scale[i] = P.scale[i];
      s_bias[i] = P.bias[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above is private set-up; activations, addends and the output are touched below

  if (warp < 4) {
    // =============================================================== A/B producers
    const __half* __restrict__ X = reinterpret_cast<const __half*>(P.x);
    const uint8_t* __restrict__ Wp = reinterpret_cast<const uint8_t*>(P.w);
    const int sh = P.in_shift;
    const int IHs = P.IH >> sh, IWs = P.IW >> sh;
    const int ohow = P.OH * P.OW;

    // Per-thread gather slots: chunk j = tid + i*128 -> (row, k-group).
    int r_oy[KG], r_ox[KG], r_nb[KG], r_base[KG];
    uint32_t r_dst[KG];
    int r_g[KG];
#pragma unroll
    for (int i = 0; i < KG; ++i) {
      const int j = tid + i * NPROD;
      const int row = j / KG;
      const int g = j - row * KG;
      r_g[i] = g;
      r_dst[i] = sw128_off(row, g);
      const int p = tile * BM + row;
      if (p < M) {
        const int n = p / ohow;
        const int rem = p - n * ohow;
        const int oy = rem / P.OW;
        const int ox = rem - oy * P.OW;
        r_oy[i] = oy * P.stride;
        r_ox[i] = ox * P.stride;
        r_nb[i] = n * IHs;
        r_base[i] = ((n * IHs + oy * P.stride) * IWs + ox * P.stride) * P.in_pix_stride + g * 8;   // in_shift == 0 only
      } else {
        r_oy[i] = -100000;  // always out of range -> zero fill
        r_ox[i] = 0;
        r_nb[i] = 0;
        r_base[i] = 0;
      }
    }

    int it = 0;
    for (int t = 0; t < P.ntaps; ++t) {
      const int dy = P.dy[t], dx = P.dx[t];
      const int tapoff = (dy * IWs + dx) * P.in_pix_stride;   // element offset of this tap (in_shift == 0)
      for (int c = 0; c < nchunks; ++c, ++it) {
        const int third = c / nkr, kr = c - third * nkr;
        const int coff = (third == 1 ? P.Cin : 0) + kr * 64;   // channel offset of this K-chunk in the source pixel
        const int kgc = min(8, (P.Cin - kr * 64) >> 3);   // real 16-byte chunks per row in this K-chunk
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t a_s = stages0 + s * st_bytes;
        if (tid == 0) {
          mbar_arrive_expect_tx(bar_full + 8 * s, b_bytes);
          bulk_g2s(a_s + a_bytes, Wp + static_cast<size_t>(it) * b_bytes, b_bytes, bar_full + 8 * s);
        }
        if (sh == 0) {
          // fast path: source address = per-row base + per-tap offset (+ K-chunk), bounds from two compares
          const __half* xc = X + tapoff + coff;
#pragma unroll
          for (int i = 0; i < KG; ++i) {
            if (r_g[i] < kgc) {
              const bool ok = (static_cast<unsigned>(r_oy[i] + dy) < static_cast<unsigned>(P.IH)) &&
                              (static_cast<unsigned>(r_ox[i] + dx) < static_cast<unsigned>(P.IW));
              cp_async16(a_s + r_dst[i], ok ? static_cast<const void*>(xc + r_base[i]) : static_cast<const void*>(X),
                         ok ? 16u : 0u);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < KG; ++i) {
            const int iy = r_oy[i] + dy;
            const int ix = r_ox[i] + dx;
            const bool ok = (static_cast<unsigned>(iy) < static_cast<unsigned>(P.IH)) &&
                            (static_cast<unsigned>(ix) < static_cast<unsigned>(P.IW));
            if (r_g[i] < kgc) {
              const __half* src = X;
              if (ok) {
                const int64_t pix = static_cast<int64_t>(r_nb[i] + (iy >> sh)) * IWs + (ix >> sh);
                src = X + pix * P.in_pix_stride + coff + r_g[i] * 8;
              }
              cp_async16(a_s + r_dst[i], src, ok ? 16u : 0u);
            }
          }
        }
        cp_async_commit();
        if (it >= LOOK) {
          cp_async_wait<LOOK>();
          fence_proxy_async();
          mbar_arrive(bar_full + 8 * ((it - LOOK) % STAGES));
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int j = (niter > LOOK ? niter - LOOK : 0); j < niter; ++j) mbar_arrive(bar_full + 8 * (j % STAGES));

    // =============================================================== epilogue
    mbar_wait(bar_accum, 0);
    tc_fence_after();

    const int row = warp * 32 + lane;
    const int p = tile * BM + row;
    const bool valid = p < M;
    int n = 0, oyf = 0, oxf = 0;
    if (valid) {
      n = p / ohow;
      const int rem = p - n * ohow;
      const int oy = rem / P.OW;
      const int ox = rem - oy * P.OW;
      oyf = oy * P.out_mul + P.out_offy;
      oxf = ox * P.out_mul + P.out_offx;
    }
    const int Cout = P.Cout;
    const int lo_off = P.pair_lo_offset > 0 ? P.pair_lo_offset : Cout;   // split mode: hi -> lo channel offset
    const __half* a0 = nullptr;
    const __half* a1 = nullptr;
    if (P.add0 != nullptr) {
      const int s0 = P.add0_shift;
      const int64_t ap = (static_cast<int64_t>(n) * (P.OHf >> s0) + (oyf >> s0)) * (P.OWf >> s0) + (oxf >> s0);
      a0 = reinterpret_cast<const __half*>(P.add0) + ap * P.add_pix_stride;
    }
    if (P.add1 != nullptr) {
      const int s1 = P.add1_shift;
      const int64_t ap = (static_cast<int64_t>(n) * (P.OHf >> s1) + (oyf >> s1)) * (P.OWf >> s1) + (oxf >> s1);
      a1 = reinterpret_cast<const __half*>(P.add1) + ap * P.add_pix_stride;
    }
    const int64_t opix = (static_cast<int64_t>(n) * P.OHf + oyf) * P.OWf + oxf;
    const bool relu = (P.flags & I2R_F_RELU) != 0;
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

    for (int c0 = 0; c0 < Npad; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(taddr_row + c0, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * s_scale[c0 + i] + s_bias[c0 + i];
      if (valid) {
        const bool act_first = (P.flags & I2R_F_ACT_FIRST) != 0;
        if (act_first) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = epi_act(v[i], P.flags);
        }
        for (int part = 0; part < (split ? 2 : 1); ++part) {   // split addends: hi half, then lo half at +Cout
        const int po = part * lo_off;
        if (a0 != nullptr) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (c0 + h * 8 < Cout) {
              const uint4 q = *reinterpret_cast<const uint4*>(a0 + po + c0 + h * 8);
              const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_h2(w4[i]);
                v[h * 8 + 2 * i] += f.x;
                v[h * 8 + 2 * i + 1] += f.y;
              }
            }
          }
        }
        if (a1 != nullptr) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (c0 + h * 8 < Cout) {
              const uint4 q = *reinterpret_cast<const uint4*>(a1 + po + c0 + h * 8);
              const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_h2(w4[i]);
                v[h * 8 + 2 * i] += f.x;
                v[h * 8 + 2 * i + 1] += f.y;
              }
            }
          }
        }
        }
        if (!act_first) {
          if (P.flags & I2R_F_GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = epi_act(v[i], P.flags);
          } else if (relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
          }
        }
        if (P.flags & I2R_F_OUT_NCHW_F32) {
          float* Y = reinterpret_cast<float*>(P.y);
          const int64_t plane = static_cast<int64_t>(P.OHf) * P.OWf;
          const int64_t base = static_cast<int64_t>(n) * Cout * plane + static_cast<int64_t>(oyf) * P.OWf + oxf;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < Cout) Y[base + (c0 + i) * plane] = v[i];
        } else if (P.flags & I2R_F_OUT_F32) {
          float* Y = reinterpret_cast<float*>(P.y) + opix * P.out_pix_stride + c0;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < Cout) Y[i] = v[i];
        } else {
          __half* Y = reinterpret_cast<__half*>(P.y) + opix * P.out_pix_stride + c0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (c0 + h * 8 < Cout) {
              uint4 q;
              q.x = pack_h2(v[h * 8 + 0], v[h * 8 + 1]);
              q.y = pack_h2(v[h * 8 + 2], v[h * 8 + 3]);
              q.z = pack_h2(v[h * 8 + 4], v[h * 8 + 5]);
              q.w = pack_h2(v[h * 8 + 6], v[h * 8 + 7]);
              *reinterpret_cast<uint4*>(Y + h * 8) = q;
              if (split) {   // lo half = fp16(v - fp16(v)) at channel offset Cout
                const uint32_t hq[4] = {q.x, q.y, q.z, q.w};
                uint32_t lq[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = unpack_h2(hq[i]);
                  lq[i] = pack_h2(v[h * 8 + 2 * i] - f.x, v[h * 8 + 2 * i + 1] - f.y);
                }
                *reinterpret_cast<uint4*>(Y + lo_off + h * 8) = make_uint4(lq[0], lq[1], lq[2], lq[3]);
              }
            }
          }
        }
      }
    }
  } else {
    // =============================================================== MMA issuer (warp-uniform loop, one lane issues)
    const uint32_t idesc = make_idesc_f16(BM, Npad);
    const uint32_t a_hi = sw128_desc_hi(1024, 0), b_hi = a_hi;
    const uint32_t a_lo0 = sw128_desc_lo(stages0);
    const uint32_t b_lo0 = sw128_desc_lo(stages0 + a_bytes);
    const uint32_t st16 = static_cast<uint32_t>(st_bytes) >> 4;
    const bool leader = elect_one();
    uint32_t accum = 0;
    for (int it = 0; it < niter; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait_warp(bar_full + 8 * s, ph);   // one lane polls (see i2r_common.cuh)
      tc_fence_after();
      const uint32_t a_lo = a_lo0 + s * st16, b_lo = b_lo0 + s * st16;
      const int ksteps = min(4, (P.Cin - ((it % nchunks) % nkr) * 64) >> 4);
      if (leader) {
        switch (ksteps) {
          case 4: issue_ksteps<4>(tmem_base, a_lo, a_hi, b_lo, b_hi, idesc, accum); break;
          case 3: issue_ksteps<3>(tmem_base, a_lo, a_hi, b_lo, b_hi, idesc, accum); break;
          case 2: issue_ksteps<2>(tmem_base, a_lo, a_hi, b_lo, b_hi, idesc, accum); break;
          default: issue_ksteps<1>(tmem_base, a_lo, a_hi, b_lo, b_hi, idesc, accum); break;
        }
      }
      accum = 1;
      if (leader) umma_commit(bar_empty + 8 * s);
    }
    if (leader) umma_commit(bar_accum);
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, ncols);
}

template <int STAGES>
__global__ void __launch_bounds__(NTHREADS) igemm_tc_kernel(const __grid_constant__ ConvGroup G) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SW128 tiles need 1024-B alignment
  int tile = blockIdx.x;
  int pi = 0;
  while (pi < G.nprob - 1 && tile >= G.tile_end[pi]) ++pi;
  if (pi > 0) tile -= G.tile_end[pi - 1];
  const i2r_conv_problem& P = G.p[pi];
  run_tile<STAGES>(P, tile, smem);
}

// ------------------------------------------------------------------------------------------------
// Scalar check kernel (tests only): same problem struct and packed weights, one thread per output
// pixel, fp32 accumulation of fp16 products.
__global__ void __launch_bounds__(128) igemm_check_kernel(const __grid_constant__ ConvGroup G) {
  pdl_launch_dependents();
  pdl_wait();
  int tile = blockIdx.x;
  int pi = 0;
  while (pi < G.nprob - 1 && tile >= G.tile_end[pi]) ++pi;
  if (pi > 0) tile -= G.tile_end[pi - 1];
  const i2r_conv_problem& P = G.p[pi];
  const int M = P.NB * P.OH * P.OW;
  const int p = tile * BM + threadIdx.x;
  if (p >= M) return;
  const int ohow = P.OH * P.OW;
  const int n = p / ohow;
  const int rem = p - n * ohow;
  const int oy = rem / P.OW, ox = rem - (rem / P.OW) * P.OW;
  const int oyf = oy * P.out_mul + P.out_offy, oxf = ox * P.out_mul + P.out_offx;
  const int sh = P.in_shift;
  const int IHs = P.IH >> sh, IWs = P.IW >> sh;
  const int lo_off_c = P.pair_lo_offset > 0 ? P.pair_lo_offset : P.Cout;
  const bool split = (P.flags & I2R_F_SPLIT) != 0;
  const int nkr = (P.Cin + 63) >> 6;
  const int nchunks = split ? 3 * nkr : nkr;
  const __half* X = reinterpret_cast<const __half*>(P.x);
  const __half* Wp = reinterpret_cast<const __half*>(P.w);
  const int64_t opix = (static_cast<int64_t>(n) * P.OHf + oyf) * P.OWf + oxf;
  for (int co = 0; co < P.Cout; ++co) {
    float acc = 0.f;
    for (int t = 0; t < P.ntaps; ++t) {
      const int iy = oy * P.stride + P.dy[t], ix = ox * P.stride + P.dx[t];
      if (static_cast<unsigned>(iy) >= static_cast<unsigned>(P.IH) ||
          static_cast<unsigned>(ix) >= static_cast<unsigned>(P.IW))
        continue;
      const __half* xp = X + (static_cast<int64_t>(n * IHs + (iy >> sh)) * IWs + (ix >> sh)) * P.in_pix_stride;
      for (int c = 0; c < P.Cin; ++c) {
        const int ch = c >> 6, g = (c & 63) >> 3, e = c & 7;
        const int64_t wi = (static_cast<int64_t>(t * nchunks + ch) * P.Npad + co) * 64 + ((g ^ (co & 7)) << 3) + e;
        acc += __half2float(xp[c]) * __half2float(Wp[wi]);
        if (split) {   // + x_lo * W_hi (second third of K) + x_hi * W_lo (last third)
          const int64_t step = static_cast<int64_t>(nkr) * P.Npad * 64;
          acc += __half2float(xp[P.Cin + c]) * __half2float(Wp[wi + step]);
          acc += __half2float(xp[c]) * __half2float(Wp[wi + 2 * step]);
        }
      }
    }
    float v = acc * P.scale[co] + P.bias[co];
    if (P.flags & I2R_F_ACT_FIRST) v = epi_act(v, P.flags);
    if (P.add0) {
      const int s0 = P.add0_shift;
      const int64_t ap = (static_cast<int64_t>(n) * (P.OHf >> s0) + (oyf >> s0)) * (P.OWf >> s0) + (oxf >> s0);
      v += __half2float(reinterpret_cast<const __half*>(P.add0)[ap * P.add_pix_stride + co]);
      if (split) v += __half2float(reinterpret_cast<const __half*>(P.add0)[ap * P.add_pix_stride + lo_off_c + co]);
    }
    if (P.add1) {
      const int s1 = P.add1_shift;
      const int64_t ap = (static_cast<int64_t>(n) * (P.OHf >> s1) + (oyf >> s1)) * (P.OWf >> s1) + (oxf >> s1);
      v += __half2float(reinterpret_cast<const __half*>(P.add1)[ap * P.add_pix_stride + co]);
      if (split) v += __half2float(reinterpret_cast<const __half*>(P.add1)[ap * P.add_pix_stride + lo_off_c + co]);
    }
    if (!(P.flags & I2R_F_ACT_FIRST)) v = epi_act(v, P.flags);
    if (P.flags & I2R_F_OUT_NCHW_F32) {
      const int64_t plane = static_cast<int64_t>(P.OHf) * P.OWf;
      reinterpret_cast<float*>(P.y)[(static_cast<int64_t>(n) * P.Cout + co) * plane +
                                    static_cast<int64_t>(oyf) * P.OWf + oxf] = v;
    } else if (P.flags & I2R_F_OUT_F32) {
      reinterpret_cast<float*>(P.y)[opix * P.out_pix_stride + co] = v;
    } else {
      const __half hv = __float2half_rn(v);
      reinterpret_cast<__half*>(P.y)[opix * P.out_pix_stride + co] = hv;
      if (split) reinterpret_cast<__half*>(P.y)[opix * P.out_pix_stride + lo_off_c + co] = __float2half_rn(v - __half2float(hv));
    }
  }
}

static int validate(const i2r_conv_problem& P, int idx) {
  if (!P.x || !P.w || !P.scale || !P.bias || !P.y) {
    set_error("conv problem %d: null pointer", idx);
    return I2R_E_BADARG;
  }
  if (P.KC != 64) {
    set_error("conv problem %d: KC=%d unsupported (weights are packed in 64-slot K-chunks)", idx, P.KC);
    return I2R_E_UNSUPPORTED;
  }
  if (P.Cin <= 0 || P.Cin % 16 != 0) {
    set_error("conv problem %d: Cin=%d not a multiple of 16", idx, P.Cin);
    return I2R_E_BADARG;
  }
  if (P.Npad < 16 || P.Npad > 256 || P.Npad % 16 != 0 || P.Cout > P.Npad || P.Cout <= 0) {
    set_error("conv problem %d: Cout=%d Npad=%d invalid", idx, P.Cout, P.Npad);
    return I2R_E_BADARG;
  }
  if (P.ntaps < 1 || P.ntaps > I2R_MAX_TAPS) {
    set_error("conv problem %d: ntaps=%d", idx, P.ntaps);
    return I2R_E_BADARG;
  }
  const bool nhwc16 = !(P.flags & (I2R_F_OUT_NCHW_F32 | I2R_F_OUT_F32));
  if (nhwc16 && (P.Cout % 8 != 0 || P.out_pix_stride % 8 != 0)) {
    set_error("conv problem %d: fp16 NHWC output needs Cout, out_pix_stride multiples of 8", idx);
    return I2R_E_BADARG;
  }
  if ((P.add0 || P.add1) && (P.Cout % 8 != 0 || P.add_pix_stride % 8 != 0 || P.add_pix_stride < P.Cout)) {
    set_error("conv problem %d: addends need Cout and add_pix_stride (>= Cout) multiples of 8", idx);
    return I2R_E_BADARG;
  }
  if (P.in_pix_stride % 8 != 0 || P.in_pix_stride < P.Cin) {
    set_error("conv problem %d: in_pix_stride=%d", idx, P.in_pix_stride);
    return I2R_E_BADARG;
  }
  if (static_cast<int64_t>(P.NB) * P.IH * P.IW * P.in_pix_stride >= (1ll << 31)) {
    set_error("conv problem %d: input larger than 2^31 elements", idx);
    return I2R_E_UNSUPPORTED;
  }
  if (P.NB <= 0 || P.OH <= 0 || P.OW <= 0 || P.IH <= 0 || P.IW <= 0 || P.stride <= 0 || P.out_mul <= 0) {
    set_error("conv problem %d: bad extents", idx);
    return I2R_E_BADARG;
  }
  return 0;
}

template <int STAGES>
static int launch_tc(const ConvGroup& G, int tiles, size_t smem, cudaStream_t st) {
  static bool attr_done_dev[MAX_DEVICES] = {};   // the opt-in is a per-device property
  bool& attr_done = attr_done_dev[current_device()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(igemm_tc): %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_done = true;
  }
  launch_pdl(igemm_tc_kernel<STAGES>, dim3(tiles), dim3(NTHREADS), smem, st, G);
  return check_launch("igemm_tc_kernel");
}

I2R_HANG_SINK_SETTER(igemm_tc)
}  // namespace i2r

extern "C" int i2r_conv_igemm(const i2r_conv_problem* probs, int nprob, int impl, void* stream) {
  using namespace i2r;
  if (!probs || nprob < 1 || nprob > I2R_MAX_GROUP) {
    set_error("i2r_conv_igemm: nprob=%d out of range", nprob);
    return I2R_E_BADARG;
  }
  ConvGroup G;
  G.nprob = nprob;
  int tiles = 0;
  int max_stage = 0;
  for (int i = 0; i < nprob; ++i) {
    int rc = validate(probs[i], i);
    if (rc) return rc;
    if (probs[i].flags & I2R_F_OUT_T16) {
      set_error("i2r_conv_igemm: problem %d asks for transposed output (I2R_F_OUT_T16), which only i2r_conv_halo writes", i);
      return I2R_E_UNSUPPORTED;
    }
    G.p[i] = probs[i];
    const int64_t M = static_cast<int64_t>(probs[i].NB) * probs[i].OH * probs[i].OW;
    tiles += static_cast<int>((M + BM - 1) / BM);
    G.tile_end[i] = tiles;
    const int sb = A_STAGE + probs[i].Npad * 128;
    if (sb > max_stage) max_stage = sb;
  }
  for (int i = nprob; i < I2R_MAX_GROUP; ++i) G.tile_end[i] = tiles;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (impl == 1) {
    launch_pdl(igemm_check_kernel, dim3(tiles), dim3(128), 0, st, G);
    return check_launch("igemm_check_kernel");
  }
  if (impl != 0) {
    set_error("i2r_conv_igemm: impl=%d", impl);
    return I2R_E_BADARG;
  }
  int stages = (96 * 1024) / max_stage;
  if (stages > 4) stages = 4;
  if (stages < 2) stages = 2;
  const size_t smem = 1024 + SMEM_HDR + static_cast<size_t>(stages) * max_stage;
  switch (stages) {
    case 2: return launch_tc<2>(G, tiles, smem, st);
    case 3: return launch_tc<3>(G, tiles, smem, st);
    default: return launch_tc<4>(G, tiles, smem, st);
  }
}

// Persistent, warp-specialised convolution kernel for stride-1 3x3 and 1x1 problems (sm_100a):
// halo-tile activation staging -> tcgen05.mma with TMEM accumulators -> fused epilogue.
//
//   tile        : 8 (x) by 16 (y) output pixels = 128 accumulator rows, all Npad output channels
//   A operand   : per (tile, 64-channel K-chunk) ONE 4-D TMA box (64 ch, 10 px, 18 lines, 1 image) with
//                 CU_TENSOR_MAP_SWIZZLE_128B deposits the (8+2) x (16+2) halo tile as 128-byte rows, one per
//                 halo pixel; pixels outside the image and channels past C are zero-filled by the TMA unit
//                 (= conv padding / K padding).  The nine taps of a 3x3 filter are nine shifted windows of
//                 that single tile: the UMMA descriptor's start address moves by (dy*10+dx)*128 B with
//                 SBO = one halo line (1280 B) -- the swizzle XOR is a function of the absolute shared-memory
//                 address on sm_100 (probed on B200, see DESIGN.md: descriptor base_offset must stay 0), so shifted windows stay consistent
//                 with what TMA wrote.  Every activation byte is fetched once instead of nine times.
//   B operand   : weights pre-packed in core-matrix order, moved by the TMA unit as 1-D bulk copies;
//                 resident in shared memory for the whole kernel when they fit (<= 120 KB), else streamed
//                 per (K-chunk, tap) through an mbarrier ring.
//   D           : two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//   warps       : 0 = activation TMA producer, 1 = weight producer, 2 = TMEM owner + MMA issuer,
//                 4..11 = epilogue (TMEM -> registers -> scale/bias/residual/ReLU -> global), two warps per
//                 TMEM lane quadrant, each owning half of the output channels.
//   grid        : persistent; each problem of a grouped launch owns a contiguous CTA range sized
//                 by its share of the work, CTAs stride over that problem's tiles.
#include <cuda.h>

#include <cstdio>
#include <cstring>

#include "i2r_common.cuh"

namespace i2r {

struct HaloProblem {
  const __half* x;
  const uint8_t* w;
  const float* scale;
  const float* bias;
  const __half* add0;
  const __half* add1;
  void* y;
  int NB, H, W, C, Cout, Npad;
  int ntaps, halo;
  int KCH, nkc;        // channels per A stage, A stages per tile
  int kgp, nchp;       // packed-weight geometry: k-groups per packed chunk, packed chunks per tap
  int tiles_x, tiles_per_img, ntiles;
  int in_pix_stride, out_pix_stride, add_pix_stride;
  int plane;           // real OH*OW of the output (NCHW addressing)
  uint32_t flags;
  int cta_begin, cta_count;
  int w_resident;
  uint32_t w_total_bytes, w_stage_bytes;
  uint32_t a_stage_bytes, a_tx_bytes;
  int a_stages, w_stages;
  uint32_t w_off;      // byte offset of the weight region in dynamic smem
};

struct HaloGroup {
  CUtensorMap amap[I2R_MAX_GROUP];
  unsigned long long* trace;   // optional event trace (tools/trace_halo.py): three role regions of trace_cap (tag<<32|tile, clock64) pairs
  int trace_cta, trace_cap;
  HaloProblem p[I2R_MAX_GROUP];
  int nprob;
};

constexpr int T_THREADS = 384;
constexpr int T_TW = 8, T_TH = 16;
constexpr uint32_t T_A_OFF = 3072;          // dynamic smem: [0,256) barriers | [256,2304) scale,bias | A ring
constexpr uint32_t T_MAX_SMEM = 226 * 1024;   // + 1 KB alignment slack = 227 KB opt-in limit
constexpr uint32_t T_W_RES_MAX = 120 * 1024;

// Fire-and-forget trace record (no atomics: each role owns region `role` of the buffer and a private counter).
__device__ __forceinline__ void trace_ev(unsigned long long* tr, int cap, int role, int& idx, int tag, int tile) {
  if (tr != nullptr && idx < cap) {
    unsigned long long* e = tr + (static_cast<size_t>(role) * cap + idx) * 2;
    e[0] = (static_cast<unsigned long long>(tag) << 32) | static_cast<unsigned>(tile);
    e[1] = clock64();
    ++idx;
  }
}

template <int NTAPS, int KS>
__device__ __forceinline__ void issue_taps(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_tap,
                                           uint32_t b_hi, uint32_t idesc, uint32_t acc_first) {
  constexpr int PW = NTAPS == 9 ? T_TW + 2 : T_TW;
#pragma unroll
  for (int tap = 0; tap < NTAPS; ++tap) {
    const uint32_t a_t = a_lo + (NTAPS == 9 ? static_cast<uint32_t>(((tap / 3) * PW + (tap % 3)) * 8) : 0u);
    issue_ksteps<KS>(d_tmem, a_t, a_hi, b_lo + tap * b_tap, b_hi, idesc, tap ? 1u : acc_first);
  }
}

// MMA issue loop.  Everything here is warp-uniform (kernel parameters, loop counters, shared-memory
// addresses), so descriptor arithmetic stays on the uniform datapath and costs two 32-bit adds per
// tcgen05.mma: the loop must sustain one MMA every ~25 cycles for N = 48.
template <int NTAPS>
__device__ __forceinline__ void mma_role(const HaloProblem& P, const int cta, const uint32_t sbase,
                                         const uint32_t tmem_base, const uint32_t ncols, unsigned long long* tr,
                                         const int trcap) {
  constexpr int HALO = NTAPS == 9 ? 1 : 0;
  constexpr int PW = T_TW + 2 * HALO;   // halo line = PW pixels = PW 128-byte rows (dense TMA box)
  constexpr uint32_t A_SBO = PW * 128;
  const uint32_t bar_afull = sbase, bar_aempty = sbase + 32, bar_wfull = sbase + 64, bar_wempty = sbase + 96;
  const uint32_t bar_accfull = sbase + 128, bar_accempty = sbase + 144, bar_wres = sbase + 160;
  const uint32_t a_base = sbase + T_A_OFF;
  const uint32_t w_base = sbase + P.w_off;
  const int Npad = P.Npad;
  const uint32_t idesc = make_idesc_f16(128, Npad);
  const uint32_t b_hi = sw128_desc_hi(1024, 0);
  const uint32_t a_hi = sw128_desc_hi(A_SBO, 0);   // base_offset 0: the swizzle XOR follows absolute smem address bits
  const uint32_t w_stage16 = P.w_stage_bytes >> 4;                         // one (tap, K-chunk) block = Npad rows x 128 B
  const uint32_t b_tap = static_cast<uint32_t>(P.nkc) * w_stage16;         // resident: next tap
  const uint32_t w_lo0 = sw128_desc_lo(w_base);
  const uint32_t a_stage16 = P.a_stage_bytes >> 4;
  const uint32_t a_lo0 = sw128_desc_lo(a_base);
  const bool resident = P.w_resident != 0;
  const bool leader = elect_one();
  int tri = 0;
  int as = 0, ws = 0, acc = 0;
  uint32_t aph = 0, wph = 0, accph = 0;
  if (resident) mbar_wait(bar_wres, 0);
  for (int t = cta; t < P.ntiles; t += P.cta_count) {
    mbar_wait(bar_accempty + 8 * acc, accph ^ 1);
    tc_fence_after();
    if (leader) trace_ev(tr, trcap, 1, tri, 10, t);
    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * (ncols >> 1);
    for (int kc = 0; kc < P.nkc; ++kc) {
      mbar_wait(bar_afull + 8 * as, aph);
      tc_fence_after();
      if (leader) trace_ev(tr, trcap, 1, tri, 11, t);
      const uint32_t a_lo = a_lo0 + as * a_stage16;
      const uint32_t b_lo_kc = w_lo0 + kc * w_stage16;
      const int ksteps = min(4, (P.C - kc * 64) >> 4);   // K=16 steps holding real channels in this chunk
      const uint32_t acc_kc = (kc != 0) ? 1u : 0u;
      if (resident) {
        // one straight-line burst of NTAPS x ksteps MMAs issued by the elected lane
        if (leader) {
          switch (ksteps) {
            case 4: issue_taps<NTAPS, 4>(d_tmem, a_lo, a_hi, b_lo_kc, b_tap, b_hi, idesc, acc_kc); break;
            case 3: issue_taps<NTAPS, 3>(d_tmem, a_lo, a_hi, b_lo_kc, b_tap, b_hi, idesc, acc_kc); break;
            case 2: issue_taps<NTAPS, 2>(d_tmem, a_lo, a_hi, b_lo_kc, b_tap, b_hi, idesc, acc_kc); break;
            default: issue_taps<NTAPS, 1>(d_tmem, a_lo, a_hi, b_lo_kc, b_tap, b_hi, idesc, acc_kc); break;
          }
        }
      } else {
#pragma unroll
        for (int tap = 0; tap < NTAPS; ++tap) {
          mbar_wait(bar_wfull + 8 * ws, wph);
          tc_fence_after();
          if (leader) {
            const uint32_t b_lo = w_lo0 + ws * w_stage16;
            const uint32_t a_t = a_lo + (HALO ? static_cast<uint32_t>(((tap / 3) * PW + (tap % 3)) * 8) : 0u);
            const uint32_t accf = tap ? 1u : acc_kc;
            switch (ksteps) {
              case 4: issue_ksteps<4>(d_tmem, a_t, a_hi, b_lo, b_hi, idesc, accf); break;
              case 3: issue_ksteps<3>(d_tmem, a_t, a_hi, b_lo, b_hi, idesc, accf); break;
              case 2: issue_ksteps<2>(d_tmem, a_t, a_hi, b_lo, b_hi, idesc, accf); break;
              default: issue_ksteps<1>(d_tmem, a_t, a_hi, b_lo, b_hi, idesc, accf); break;
            }
            umma_commit(bar_wempty + 8 * ws);
          }
          if (++ws == P.w_stages) {
            ws = 0;
            wph ^= 1;
          }
        }
      }
      if (leader) umma_commit(bar_aempty + 8 * as);
      if (++as == P.a_stages) {
        as = 0;
        aph ^= 1;
      }
    }
    if (leader) umma_commit(bar_accfull + 8 * acc);
    if (leader) trace_ev(tr, trcap, 1, tri, 12, t);
    acc ^= 1;
    if (acc == 0) accph ^= 1;
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__global__ void __launch_bounds__(T_THREADS, 1) conv_halo_kernel(const __grid_constant__ HaloGroup G) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SW128 tiles need 1024-B alignment
  int pi = 0;
  while (pi < G.nprob - 1 && static_cast<int>(blockIdx.x) >= G.p[pi].cta_begin + G.p[pi].cta_count) ++pi;
  const HaloProblem& P = G.p[pi];
  const int cta = blockIdx.x - P.cta_begin;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_afull = sbase, bar_aempty = sbase + 32, bar_wfull = sbase + 64, bar_wempty = sbase + 96;
  const uint32_t bar_accfull = sbase + 128, bar_accempty = sbase + 144, bar_wres = sbase + 160;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 176);
  float* s_scale = reinterpret_cast<float*>(smem + 256);
  float* s_bias = reinterpret_cast<float*>(smem + 256 + 1024);
  const uint32_t a_base = sbase + T_A_OFF;
  const uint32_t w_base = sbase + P.w_off;

  const int Npad = P.Npad;
  uint32_t ncols = 32;
  while (ncols < static_cast<uint32_t>(2 * Npad)) ncols <<= 1;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_afull + 8 * i, 1);
      mbar_init(bar_aempty + 8 * i, 1);
      mbar_init(bar_wfull + 8 * i, 1);
      mbar_init(bar_wempty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_accfull + 8 * i, 1);
      mbar_init(bar_accempty + 8 * i, 8);
    }
    mbar_init(bar_wres, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), ncols);
    tmem_relinquish();
  }
  for (int i = tid; i < Npad; i += T_THREADS) {
    s_scale[i] = P.scale[i];
    s_bias[i] = P.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  unsigned long long* tr = (G.trace != nullptr && static_cast<int>(blockIdx.x) == G.trace_cta) ? G.trace : nullptr;

  if (warp == 0) {
    if (lane == 0) {
      // ================================================= activation producer: one TMA box per (tile, K-chunk)
      const CUtensorMap* amap = &G.amap[pi];
      prefetch_tmap(amap);
      int s = 0, tri = 0;
      uint32_t ph = 0;
      for (int t = cta; t < P.ntiles; t += P.cta_count) {
        const int n = t / P.tiles_per_img;
        const int r = t - n * P.tiles_per_img;
        const int ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
        const int x0 = tx * T_TW - P.halo, y0 = ty * T_TH - P.halo;
        for (int kc = 0; kc < P.nkc; ++kc) {
          mbar_wait(bar_aempty + 8 * s, ph ^ 1);
          trace_ev(tr, G.trace_cap, 0, tri, 1, t);
          mbar_arrive_expect_tx(bar_afull + 8 * s, P.a_tx_bytes);
          tma_load_4d(a_base + s * P.a_stage_bytes, amap, kc * 64, x0, y0, n, bar_afull + 8 * s);
          trace_ev(tr, G.trace_cap, 0, tri, 2, t);
          if (++s == P.a_stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
    // ================================================= weight producer (1-D bulk copies on the TMA unit)
    if (P.w_resident) {
      mbar_arrive_expect_tx(bar_wres, P.w_total_bytes);
      for (uint32_t off = 0; off < P.w_total_bytes; off += 16384) {
        const uint32_t sz = min(16384u, P.w_total_bytes - off);
        bulk_g2s(w_base + off, P.w + off, sz, bar_wres);
      }
    } else {
      int s = 0;
      uint32_t ph = 0;
      for (int t = cta; t < P.ntiles; t += P.cta_count) {
        for (int kc = 0; kc < P.nkc; ++kc) {
          for (int tap = 0; tap < P.ntaps; ++tap) {
            mbar_wait(bar_wempty + 8 * s, ph ^ 1);
            mbar_arrive_expect_tx(bar_wfull + 8 * s, P.w_stage_bytes);
            bulk_g2s(w_base + s * P.w_stage_bytes,
                     P.w + static_cast<size_t>(tap * P.nchp + kc) * P.w_stage_bytes, P.w_stage_bytes,
                     bar_wfull + 8 * s);
            if (++s == P.w_stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
    }
  } else if (warp == 2) {
    // ================================================= MMA issuer (whole warp runs the loop, one lane issues)
    if (P.ntaps == 9) {
      mma_role<9>(P, cta, sbase, tmem_base, ncols, tr, G.trace_cap);
    } else {
      mma_role<1>(P, cta, sbase, tmem_base, ncols, tr, G.trace_cap);
    }
  } else if (warp >= 4) {
    // ================================================= epilogue: 8 warps, two per TMEM lane quadrant, each
    // owning half of the output channels of its 32 rows; residual loads are issued before the accumulator
    // wait so their latency hides behind the MMAs of the tile.
    const int ew = warp - 4;
    const int quad = warp & 3;             // warp % 4: the TMEM lanes this warp may read
    const int row = quad * 32 + lane;
    const int ty_in = row >> 3, tx_in = row & 7;
    const int Cout = P.Cout;
    const bool relu = (P.flags & I2R_F_RELU) != 0;
    const int n8 = Npad >> 3, half8 = (n8 + 1) >> 1;
    const int cb = (ew >> 2) ? half8 : 0, ce = (ew >> 2) ? n8 : half8;   // 8-column chunks [cb, ce)
    const float4* s_scale4 = reinterpret_cast<const float4*>(s_scale);
    const float4* s_bias4 = reinterpret_cast<const float4*>(s_bias);
    int acc = 0;
    int tri = 0;
    uint32_t accph = 0;
    for (int t = cta; t < P.ntiles; t += P.cta_count) {
      const int n = t / P.tiles_per_img;
      const int r = t - n * P.tiles_per_img;
      const int ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
      const int x = tx * T_TW + tx_in, y = ty * T_TH + ty_in;
      const bool valid = (x < P.W) && (y < P.H);
      const int64_t p = (static_cast<int64_t>(n) * P.H + y) * P.W + x;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc) * (ncols >> 1);
      const __half* a0 = (P.add0 && valid) ? P.add0 + p * P.add_pix_stride : nullptr;
      const __half* a1 = (P.add1 && valid) ? P.add1 + p * P.add_pix_stride : nullptr;
      bool waited = false;
      for (int c = cb; c < ce; c += 4) {
        const int nc = min(4, ce - c);
        uint4 r0[4], r1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          r0[j] = make_uint4(0, 0, 0, 0);
          r1[j] = make_uint4(0, 0, 0, 0);
          if (j < nc && (c + j) * 8 < Cout) {
            if (a0 != nullptr) r0[j] = __ldg(reinterpret_cast<const uint4*>(a0 + (c + j) * 8));
            if (a1 != nullptr) r1[j] = __ldg(reinterpret_cast<const uint4*>(a1 + (c + j) * 8));
          }
        }
        if (!waited) {
          mbar_wait(bar_accfull + 8 * acc, accph);
          tc_fence_after();
          waited = true;
          if (ew == 0 && lane == 0) trace_ev(tr, G.trace_cap, 2, tri, 20, t);
        }
        uint32_t av[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nc) tmem_ld8(taddr + (c + j) * 8, av[j]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nc && valid && (c + j) * 8 < Cout) {
            const int c0 = (c + j) * 8;
            const float4 sa = s_scale4[2 * (c + j)], sb = s_scale4[2 * (c + j) + 1];
            const float4 ba = s_bias4[2 * (c + j)], bb = s_bias4[2 * (c + j) + 1];
            float v[8];
            v[0] = __uint_as_float(av[j][0]) * sa.x + ba.x;
            v[1] = __uint_as_float(av[j][1]) * sa.y + ba.y;
            v[2] = __uint_as_float(av[j][2]) * sa.z + ba.z;
            v[3] = __uint_as_float(av[j][3]) * sa.w + ba.w;
            v[4] = __uint_as_float(av[j][4]) * sb.x + bb.x;
            v[5] = __uint_as_float(av[j][5]) * sb.y + bb.y;
            v[6] = __uint_as_float(av[j][6]) * sb.z + bb.z;
            v[7] = __uint_as_float(av[j][7]) * sb.w + bb.w;
            const uint32_t q0[4] = {r0[j].x, r0[j].y, r0[j].z, r0[j].w};
            const uint32_t q1[4] = {r1[j].x, r1[j].y, r1[j].z, r1[j].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f0 = unpack_h2(q0[i]), f1 = unpack_h2(q1[i]);
              v[2 * i] += f0.x + f1.x;
              v[2 * i + 1] += f0.y + f1.y;
            }
            if (relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.0f);
            }
            if (P.flags & I2R_F_OUT_NCHW_F32) {
              float* Y = reinterpret_cast<float*>(P.y);
              const int64_t nr = p / P.plane, rem = p - nr * P.plane;
              const int64_t base = nr * Cout * P.plane + rem;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (c0 + i < Cout) Y[base + static_cast<int64_t>(c0 + i) * P.plane] = v[i];
            } else if (P.flags & I2R_F_OUT_F32) {
              float* Y = reinterpret_cast<float*>(P.y) + p * P.out_pix_stride + c0;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (c0 + i < Cout) Y[i] = v[i];
            } else {
              uint4 q;
              q.x = pack_h2(v[0], v[1]);
              q.y = pack_h2(v[2], v[3]);
              q.z = pack_h2(v[4], v[5]);
              q.w = pack_h2(v[6], v[7]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(P.y) + p * P.out_pix_stride + c0) = q;
            }
          }
        }
      }
      if (!waited) {  // no columns assigned to this warp (tiny N): still take part in the hand-shake
        mbar_wait(bar_accfull + 8 * acc, accph);
        tc_fence_after();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accempty + 8 * acc);
      if (ew == 0 && lane == 0) trace_ev(tr, G.trace_cap, 2, tri, 21, t);
      acc ^= 1;
      if (acc == 0) accph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// NHWC activation tensor as a 4-D TMA tensor (C, W, H, N); box = (64 channels, halo width, halo height, 1) with
// 128-byte swizzle: one 128-byte shared-memory row per pixel.  Out-of-range pixels and channels read as zero.
static int encode_amap(CUtensorMap* map, const void* x, int NB, int H, int W, int C, int pix_stride, int hw, int hh) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return I2R_E_DEVICE;
  }
  const cuuint64_t pb = static_cast<cuuint64_t>(pix_stride) * 2;
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
  const cuuint64_t strides[3] = {pb, pb * W, pb * W * H};
  const cuuint32_t box[4] = {64, (cuuint32_t)hw, (cuuint32_t)hh, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] pix_stride %d", (int)r, NB, H, W, C, pix_stride);
    return I2R_E_DEVICE;
  }
  return 0;
}

static unsigned long long* g_trace = nullptr;
static int g_trace_cta = 0, g_trace_cap = 0;
static bool is_std3x3(const i2r_conv_problem& P) {
  if (P.ntaps != 9) return false;
  for (int t = 0; t < 9; ++t)
    if (P.dy[t] != t / 3 - 1 || P.dx[t] != t % 3 - 1) return false;
  return true;
}

}  // namespace i2r

extern "C" int i2r_conv_halo_supported(const i2r_conv_problem* P) {
  using namespace i2r;
  if (!P) return 0;
  const bool k1 = P->ntaps == 1 && P->dy[0] == 0 && P->dx[0] == 0;
  if (!(k1 || is_std3x3(*P))) return 0;
  if (P->stride != 1 || P->in_shift != 0 || P->out_mul != 1 || P->out_offy != 0 || P->out_offx != 0) return 0;
  if (P->OH != P->IH || P->OW != P->IW || P->OHf != P->OH || P->OWf != P->OW) return 0;
  if ((P->add0 && P->add0_shift != 0) || (P->add1 && P->add1_shift != 0)) return 0;
  if (P->Cin % 16 != 0 || P->Npad > 256 || P->Npad % 16 != 0) return 0;
  if (P->KC != 64) return 0;
  if (P->in_pix_stride % 8 != 0) return 0;
  if ((P->add0 || P->add1) && (P->add_pix_stride % 8 != 0 || P->add_pix_stride < P->Cout)) return 0;
  return 1;
}

extern "C" int i2r_conv_halo(const i2r_conv_problem* probs, int nprob, void* stream) {
  using namespace i2r;
  if (!probs || nprob < 1 || nprob > I2R_MAX_GROUP) {
    set_error("i2r_conv_halo: nprob=%d out of range", nprob);
    return I2R_E_BADARG;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  HaloGroup G;
  memset(&G, 0, sizeof(G));
  G.nprob = nprob;
  G.trace = g_trace;
  G.trace_cta = g_trace_cta;
  G.trace_cap = g_trace_cap;
  double cost[I2R_MAX_GROUP];
  int total_tiles = 0;
  uint32_t smem_need = 0;
  for (int i = 0; i < nprob; ++i) {
    const i2r_conv_problem& S = probs[i];
    if (!i2r_conv_halo_supported(&S)) {
      set_error("i2r_conv_halo: problem %d is not a stride-1 3x3 / 1x1 problem this kernel supports", i);
      return I2R_E_UNSUPPORTED;
    }
    if (!S.x || !S.w || !S.scale || !S.bias || !S.y) {
      set_error("i2r_conv_halo: problem %d: null pointer", i);
      return I2R_E_BADARG;
    }
    HaloProblem& P = G.p[i];
    P.x = static_cast<const __half*>(S.x);
    P.w = static_cast<const uint8_t*>(S.w);
    P.scale = S.scale;
    P.bias = S.bias;
    P.add0 = static_cast<const __half*>(S.add0);
    P.add1 = static_cast<const __half*>(S.add1);
    P.y = S.y;
    P.ntaps = S.ntaps;
    P.halo = S.ntaps == 9 ? 1 : 0;
    P.NB = S.NB;
    P.H = S.IH;
    P.W = S.IW;
    const int64_t mtot = static_cast<int64_t>(S.NB) * S.IH * S.IW;
    P.plane = S.OHf * S.OWf;
    if (S.ntaps == 1 && mtot % 8 == 0) {  // pixels are independent: re-tile as an 8-wide strip
      P.NB = 1;
      P.W = 8;
      P.H = static_cast<int>(mtot / 8);
    }
    P.C = S.Cin;
    P.Cout = S.Cout;
    P.Npad = S.Npad;
    P.KCH = 64;
    P.nkc = (S.Cin + 63) / 64;
    P.kgp = 8;
    P.nchp = P.nkc;
    P.tiles_x = (P.W + T_TW - 1) / T_TW;
    P.tiles_per_img = P.tiles_x * ((P.H + T_TH - 1) / T_TH);
    P.ntiles = P.tiles_per_img * P.NB;
    P.in_pix_stride = S.in_pix_stride;
    P.out_pix_stride = S.out_pix_stride;
    P.add_pix_stride = S.add_pix_stride;
    P.flags = S.flags;
    const int hw = T_TW + 2 * P.halo, hh = T_TH + 2 * P.halo;
    P.a_tx_bytes = static_cast<uint32_t>(hh * hw * 128);          // full box, zero-filled parts included
    P.a_stage_bytes = (P.a_tx_bytes + 1023u) & ~1023u;            // stages stay 1024-byte aligned (SW128)
    P.w_total_bytes = static_cast<uint32_t>(S.ntaps) * P.nkc * S.Npad * 128;
    P.w_resident = P.w_total_bytes <= T_W_RES_MAX ? 1 : 0;
    P.w_stage_bytes = static_cast<uint32_t>(S.Npad) * 128;
    uint32_t wregion;
    if (P.w_resident) {
      P.w_stages = 1;
      wregion = P.w_total_bytes;
    } else {
      P.w_stages = 4;
      while (P.w_stages > 2 && P.w_stages * P.w_stage_bytes > 110 * 1024) --P.w_stages;
      wregion = P.w_stages * P.w_stage_bytes;
    }
    int astg = static_cast<int>((T_MAX_SMEM - T_A_OFF - wregion) / P.a_stage_bytes);
    if (astg > 4) astg = 4;
    if (astg < 2) {
      set_error("i2r_conv_halo: problem %d does not fit shared memory (A stage %u B, W region %u B)", i,
                P.a_stage_bytes, wregion);
      return I2R_E_UNSUPPORTED;
    }
    P.a_stages = astg;
    {
      int rc = encode_amap(&G.amap[i], S.x, P.NB, P.H, P.W, P.C, S.in_pix_stride, hw, hh);
      if (rc) return rc;
    }
    P.w_off = T_A_OFF + static_cast<uint32_t>(astg) * P.a_stage_bytes;
    const uint32_t need = P.w_off + wregion;
    if (need > smem_need) smem_need = need;
    cost[i] = static_cast<double>(S.ntaps) * (S.Cin / 16) * (4096.0 + 32.0 * S.Npad) * (P.w_resident ? 1.0 : 1.3);  // per tile
    total_tiles += P.ntiles;
  }
  // CTA ranges: one CTA per tile while they fit; otherwise hand the SMs out greedily to whichever problem
  // currently has the longest makespan ceil(tiles / CTAs) * tile_cost (tile_cost from the measured
  // operand-fetch-bound MMA rate: ~ (4096 + 32 * Npad) bytes of shared-memory operands per K=16 step).
  int begin = 0;
  if (total_tiles <= num_sms) {
    for (int i = 0; i < nprob; ++i) {
      G.p[i].cta_begin = begin;
      G.p[i].cta_count = G.p[i].ntiles;
      begin += G.p[i].ntiles;
    }
  } else {
    int cnt[I2R_MAX_GROUP];
    int left = num_sms;
    for (int i = 0; i < nprob; ++i) {
      cnt[i] = 1;
      --left;
    }
    while (left > 0) {
      int best = -1;
      double bestv = -1;
      for (int i = 0; i < nprob; ++i) {
        if (cnt[i] >= G.p[i].ntiles) continue;
        const double v = static_cast<double>((G.p[i].ntiles + cnt[i] - 1) / cnt[i]) * cost[i];
        if (v > bestv) {
          bestv = v;
          best = i;
        }
      }
      if (best < 0) break;
      ++cnt[best];
      --left;
    }
    for (int i = 0; i < nprob; ++i) {
      G.p[i].cta_begin = begin;
      G.p[i].cta_count = cnt[i];
      begin += cnt[i];
    }
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_MAX_SMEM + 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(conv_halo): %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_done = true;
  }
  conv_halo_kernel<<<begin, T_THREADS, smem_need + 1024, static_cast<cudaStream_t>(stream)>>>(G);
  return check_launch("conv_halo_kernel");
}

extern "C" int i2r_debug_trace(void* dev_buffer, int capacity_events, int cta) {
  i2r::g_trace = static_cast<unsigned long long*>(dev_buffer);
  i2r::g_trace_cap = capacity_events;
  i2r::g_trace_cta = cta;
  return 0;
}

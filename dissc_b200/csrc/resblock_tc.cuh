// Fused ResBlock1 pair for the narrow vocoder stages (C = 16 / 32):
//
//     x' = x + conv2( lrelu( conv1( lrelu(x), dilation d ) ) )                    (sr/models.py:36-40)
//     [+ MRF accumulate / divide / next-stage leaky-relu in the same epilogue]   (:104-110)
//
// in ONE kernel.  The unfused path moves 24 B per element and pair through HBM (planes in, planes out, planes in,
// residual in, fp32 out, planes out); here x is read once as fp32 (with the dilation halo) and x' written once:
// ~8-10 B.  The intermediate xt never leaves the SM, and the fp16 hi/lo operand split of lrelu(x) is done on chip.
//
// Per tile of M_out = 128-(k-1) output rows:
//   producer warp : bulk-TMA the fp32 tile x[t0-p2-p1 .. +R1) (R1 = 128+(k-1)d rows, C/8 slabs) into a staging buffer
//   worker group  : (4 warps) convert staging -> lrelu -> fp16 hi/lo -> UMMA operand tile (zero outside [0,T))
//   MMA thread    : conv1 as k shifted MMAs into TMEM acc1   (128 rows of xt: t0-p2 .. t0-p2+127)
//   worker group  : epilogue 1: TMEM acc1 -> +bias -> lrelu -> zero outside [0,T) -> fp16 hi/lo -> xt operand tile (smem)
//   MMA thread    : conv2 (dilation 1) from the xt tile into TMEM acc2
//   worker group  : epilogue 2: acc2 + bias + x (residual, L2-hot re-read) [+ xs] [/ n] -> fp32 / planes / plain output
// Two worker groups alternate tiles (every buffer and accumulator is per group), so the conversions and epilogues of
// one tile overlap the MMAs of the other; the single MMA thread issues conv1(s), conv2(s-1), conv1(s+1), ...
// Both convs' weights (fp16 hi/lo, 2-MMA split of conv_tc.cuh) stay resident in shared memory.
//
// HBM layout: f32h = fp32 [B][C/8][Tpf][8] with `f_halo` rows of slack in front and >= 160 behind, so a tile's halo reads
// never leave its slab; the CONTENT of the slack rows is irrelevant (rows outside [0,T) are zeroed by index).
#pragma once
#include "conv_tc.cuh"

namespace dissc {

// warp 0 producer, warp 1 MMA issuer + TMEM owner, then two worker groups of WPG warps each.  C=32: 8 warps per group
// (two warps share a TMEM lane quarter and split the channel groups) -- one CTA per SM, so the extra warps are what
// hides the epilogue latencies; C=16: 4 warps per group, two CTAs per SM.
template <int NC>
struct PairCfg {
  static constexpr int WPG = (NC >= 32) ? 8 : 4;
  static constexpr int THREADS = 64 + 2 * WPG * 32;
  static constexpr int C8 = NC / 8;
  static constexpr int CH = C8 * 4 / WPG;  // 8-channel groups per worker warp
};

struct PairParams {
  const float* x;       // f32h [B][C/8][Tpf][8]
  const __half* w1;     // packed [tap][c8][hi|lo][C][8]
  const __half* w2;
  const float* b1;      // [C]
  const float* b2;
  float inv1, inv2;     // 2^-s of the two weight scalings, with 1 / in_scale resp. 1 / xt_scale folded in by the host
  float in_scale, xt_scale, plane_scale;  // power-of-two activation scales (TcParams): the on-chip operand tile of
                                          // lrelu(x), the xt tile and the output planes hold value * scale
  const float* acc_in;  // f32h or null (MRF accumulator xs)
  float* out_f;         // f32h or null
  __half* out_hi;       // planes [B][C/8][Tp][8] or null (leaky-relu(plane_slope) iff plane_act)
  __half* out_lo;
  float* out_plain;     // (B, C, T) fp32 or null (leaky-relu(plain_slope) iff plain_act)
  const int* lengths;
  int len_mul;
  int B, T, Tpf, f_halo, Tp, p_halo;
  int k, dil;
  int tiles_per_b, n_tiles;
  int tmem_cols;
  float div;
  int plane_act, plain_act;
  float plane_slope, plain_slope;
};

template <int NC>
__global__ void __launch_bounds__(PairCfg<NC>::THREADS, (NC == 16) ? 2 : 1) resblock_pair_tc_kernel(const PairParams p) {
  constexpr int C8 = NC / 8;
  constexpr int WPG = PairCfg<NC>::WPG, CH = PairCfg<NC>::CH, kPairThreads = PairCfg<NC>::THREADS;
  constexpr int KS = NC / 16;
  constexpr uint32_t lbo_b = 2u * NC * 16;
  constexpr uint32_t w_tap_bytes = (uint32_t)C8 * lbo_b;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int k = p.k, d = p.dil;
  const int p2 = (k - 1) / 2, p1 = d * (k - 1) / 2;
  const int R1 = 128 + (k - 1) * d;   // x rows per tile
  const int R2 = 128 + (k - 1);       // xt rows addressable by conv2 (rows >= 128 stay zero)
  const int M_out = 128 - (k - 1);
  const uint32_t stg_bytes = (uint32_t)C8 * R1 * 32;
  const uint32_t xop_plane = (uint32_t)C8 * R1 * 16, xop_bytes = 2 * xop_plane;
  const uint32_t xt_plane = (uint32_t)C8 * R2 * 16, xt_bytes = 2 * xt_plane;
  const uint32_t w_bytes = (uint32_t)k * w_tap_bytes;
  unsigned char* sStg = smem_raw;                   // [2][stg_bytes]
  unsigned char* sXop = sStg + 2 * stg_bytes;       // [2][xop_bytes]
  unsigned char* sXt = sXop + 2 * xop_bytes;        // [2][xt_bytes]
  unsigned char* sW1 = sXt + 2 * xt_bytes;
  unsigned char* sW2 = sW1 + w_bytes;
  float* s_b1 = reinterpret_cast<float*>(sW2 + w_bytes);
  float* s_b2 = s_b1 + NC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b2 + NC);
  uint64_t* stg_full = bars;         // [2]
  uint64_t* stg_empty = bars + 2;    // [2]
  uint64_t* xop_full = bars + 4;     // [2]
  uint64_t* xop_empty = bars + 6;    // [2]
  uint64_t* acc1_full = bars + 8;    // [2]
  uint64_t* xt_full = bars + 10;     // [2]
  uint64_t* acc2_full = bars + 12;   // [2]
  uint64_t* acc2_empty = bars + 14;  // [2]
  uint64_t* w_full = bars + 16;      // [1]
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&stg_full[i], 1);
      mbar_init(&stg_empty[i], WPG);
      mbar_init(&xop_full[i], WPG);
      mbar_init(&xop_empty[i], 1);
      mbar_init(&acc1_full[i], 1);
      mbar_init(&xt_full[i], WPG);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], WPG);
    }
    mbar_init(&w_full[0], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < NC; i += kPairThreads) {
    s_b1[i] = p.b1 ? p.b1[i] : 0.f;
    s_b2[i] = p.b2 ? p.b2[i] : 0.f;
  }
  // rows >= 128 of both xt tiles are read by conv2's discarded output rows only: keep them finite (zero)
  for (int i = tid; i < (int)(2 * xt_bytes / 16); i += kPairThreads) reinterpret_cast<uint4*>(sXt)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  pdl_wait();                 // the prologue above touched only weights / shared memory; activations from here on
  pdl_launch_dependents();
  // TMEM columns: group g: acc1 at g*4NC (main | cross), acc2 at g*4NC + 2NC

  if (warp == 0) {
    // ===================== producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&w_full[0], 2 * w_bytes);
      tma_load_1d(sW1, p.w1, w_bytes, &w_full[0]);
      tma_load_1d(sW2, p.w2, w_bytes, &w_full[0]);
      uint32_t s = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++s) {
        const uint32_t g = s & 1, ph = (s >> 1) & 1;
        const int b = tile / p.tiles_per_b;
        const int t0 = (tile - b * p.tiles_per_b) * M_out;
        mbar_wait(&stg_empty[g], ph ^ 1);
        mbar_arrive_expect_tx(&stg_full[g], stg_bytes);
        const float* src = p.x + (((size_t)b * C8) * p.Tpf + p.f_halo + t0 - p2 - p1) * 8;
        for (int c8 = 0; c8 < C8; ++c8)
          tma_load_1d(sStg + g * stg_bytes + (size_t)c8 * R1 * 32, src + (size_t)c8 * p.Tpf * 8, (uint32_t)R1 * 32,
                      &stg_full[g]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_n = (1u << 4) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * NC) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lbo_x = (uint32_t)R1 * 16, lbo_t = (uint32_t)R2 * 16;
      const uint32_t w1d = umma_desc_lo(smem_u32(sW1), lbo_b), w2d = umma_desc_lo(smem_u32(sW2), lbo_b);
      mbar_wait(&w_full[0], 0);
      auto conv = [&](uint32_t a_addr, uint32_t a_plane, uint32_t lbo_a, uint32_t wdesc, int tap_rows, uint32_t d_main) {
        const uint32_t a0 = umma_desc_lo(a_addr, lbo_a);
        const uint32_t a_kstep = (2 * lbo_a) >> 4, b_kstep = (2 * lbo_b) >> 4, lo_off = a_plane >> 4;
        uint32_t accum = 0, ad_t = a0, wd_t = wdesc;
        for (int j = 0; j < k; ++j, ad_t += (uint32_t)tap_rows, wd_t += (w_tap_bytes >> 4)) {
          uint32_t ad = ad_t, wd = wd_t;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks, ad += a_kstep, wd += b_kstep) {
            umma_f16(d_main, umma_desc(ad), umma_desc(wd), idesc_2n, accum);            // [main | cross]
            umma_f16(d_main + NC, umma_desc(ad + lo_off), umma_desc(wd), idesc_n, 1);   // cross += lo * hi
            accum = 1;
          }
        }
      };
      int n_mine = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) ++n_mine;
      for (int s = 0; s <= n_mine; ++s) {
        if (s < n_mine) {
          const uint32_t g = s & 1, ph = (s >> 1) & 1;
          mbar_wait(&xop_full[g], ph);
          tc_fence_after();
          conv(smem_u32(sXop + g * xop_bytes), xop_plane, lbo_x, w1d, d, tmem_base + g * 4u * NC);
          umma_commit(&xop_empty[g]);
          umma_commit(&acc1_full[g]);
        }
        if (s >= 1) {
          const uint32_t sp = (uint32_t)(s - 1), g = sp & 1, ph = (sp >> 1) & 1;
          mbar_wait(&xt_full[g], ph);
          mbar_wait(&acc2_empty[g], ph ^ 1);
          tc_fence_after();
          conv(smem_u32(sXt + g * xt_bytes), xt_plane, lbo_t, w2d, 1, tmem_base + g * 4u * NC + 2u * NC);
          umma_commit(&acc2_full[g]);
        }
      }
    }
  } else {
    // ===================== worker groups: convert -> epilogue 1 -> epilogue 2 =====================
    const int g = (warp - 2) / WPG;         // worker group
    const int wi = (warp - 2) - g * WPG;    // warp inside the group
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int c8_0 = (wi >> 2) * CH;        // first of the CH channel groups this warp handles in the epilogues
    const int wt = wi * 32 + lane;          // thread index inside the group
    const int row = quarter * 32 + lane;    // TMEM lane = tile row
    unsigned char* stg = sStg + g * stg_bytes;
    unsigned char* xop = sXop + g * xop_bytes;
    unsigned char* xt = sXt + g * xt_bytes;
    const uint32_t t_acc1 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 4u * NC;
    const uint32_t t_acc2 = t_acc1 + 2u * NC;
    const unsigned r1_magic = 0xFFFFFFFFu / (unsigned)R1 + 1u;   // ceil(2^32 / R1): umulhi(item, magic) == item / R1 for item < 2^16
    // convert: fp32 staging tile -> lrelu -> fp16 hi/lo operand tile (rows outside [0, Tvalid) are zeros)
    auto convert = [&](int tile, uint32_t ph) {
      const int b = tile / p.tiles_per_b;
      const int t0 = (tile - b * p.tiles_per_b) * M_out;
      const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
      mbar_wait(&stg_full[g], ph);
      mbar_wait(&xop_empty[g], ph ^ 1);
      const int tx0 = t0 - p2 - p1;
      for (int item = wt; item < C8 * R1; item += WPG * 32) {
        const int c8 = (int)__umulhi((unsigned)item, r1_magic), i = item - c8 * R1;   // item / R1 (exact: item < 2^16)
        const int t = tx0 + i;
        if (t >= 0 && t < Tvalid) {
          const float4 a = *reinterpret_cast<const float4*>(stg + (size_t)item * 32);
          const float4 c = *reinterpret_cast<const float4*>(stg + (size_t)item * 32 + 16);
          float v[8];
          const float si = p.in_scale;   // leaky(x) * s == leaky(x * s) for s > 0
          v[0] = leaky(a.x * si, 0.1f); v[1] = leaky(a.y * si, 0.1f); v[2] = leaky(a.z * si, 0.1f); v[3] = leaky(a.w * si, 0.1f);
          v[4] = leaky(c.x * si, 0.1f); v[5] = leaky(c.y * si, 0.1f); v[6] = leaky(c.z * si, 0.1f); v[7] = leaky(c.w * si, 0.1f);
          split_store8(reinterpret_cast<__half*>(xop + (size_t)item * 16),
                       reinterpret_cast<__half*>(xop + xop_plane + (size_t)item * 16), v);
        } else {
          *reinterpret_cast<uint4*>(xop + (size_t)item * 16) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(xop + xop_plane + (size_t)item * 16) = make_uint4(0, 0, 0, 0);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (elect_one()) {
        mbar_arrive(&xop_full[g]);
        mbar_arrive(&stg_empty[g]);
      }
    };
    uint32_t it = 0;
    const int first = blockIdx.x + g * gridDim.x, step = 2 * gridDim.x;
    if (first < p.n_tiles) convert(first, 0);
    for (int tile = first; tile < p.n_tiles; tile += step, ++it) {
      const uint32_t ph = it & 1;
      const int b = tile / p.tiles_per_b;
      const int t0 = (tile - b * p.tiles_per_b) * M_out;
      const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
      // ---- residual (and MRF accumulator) prefetch for epilogue 2: row t0+row of the same fp32 tensor (L2-hot)
      const int t_out = t0 + row;
      const bool out_valid = row < M_out && t_out < Tvalid;
      float4 rq[CH * 2], aq[CH * 2];
      if (out_valid) {
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          const size_t fi = (((size_t)b * C8 + c8_0 + q) * p.Tpf + p.f_halo + t_out) * 8;
          ldg8(p.x + fi, rq[2 * q], rq[2 * q + 1]);
          if (p.acc_in) ldg8(p.acc_in + fi, aq[2 * q], aq[2 * q + 1]);
        }
      }
      // ---- epilogue 1: acc1 -> xt tile
      mbar_wait(&acc1_full[g], ph);
      tc_fence_after();
      {
        const int t = t0 - p2 + row;
        const bool v_ok = t >= 0 && t < Tvalid;
        float m[CH][8], x8[CH][8];
#pragma unroll
        for (int q = 0; q < CH; ++q) {   // all TMEM loads in flight, one wait
          tmem_ld8(t_acc1 + (c8_0 + q) * 8, m[q]);
          tmem_ld8(t_acc1 + NC + (c8_0 + q) * 8, x8[q]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          const int c8 = c8_0 + q;
          const size_t o = ((size_t)c8 * R2 + row) * 16;
          if (v_ok) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = leaky((m[q][e] + x8[q][e]) * p.inv1 + s_b1[c8 * 8 + e], 0.1f) * p.xt_scale;
            split_store8(reinterpret_cast<__half*>(xt + o), reinterpret_cast<__half*>(xt + xt_plane + o), v);
          } else {
            *reinterpret_cast<uint4*>(xt + o) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(xt + xt_plane + o) = make_uint4(0, 0, 0, 0);
          }
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&xt_full[g]);
      // ---- operand tile of this group's NEXT tile, so the MMA thread never waits for it
      if (tile + step < p.n_tiles) convert(tile + step, ph ^ 1);
      // ---- epilogue 2: acc2 + bias + residual [+ xs] [/ n] -> outputs
      mbar_wait(&acc2_full[g], ph);
      tc_fence_after();
      float m2[CH][8], y2[CH][8];
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        tmem_ld8(t_acc2 + (c8_0 + q) * 8, m2[q]);
        tmem_ld8(t_acc2 + NC + (c8_0 + q) * 8, y2[q]);
      }
      tmem_ld_wait();
      // the accumulator is in registers: release it before the (long) store phase
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_empty[g]);
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int c8 = c8_0 + q;
        if (!out_valid && !(p.out_hi && row < M_out && t_out < p.T)) continue;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (m2[q][e] + y2[q][e]) * p.inv2 + s_b2[c8 * 8 + e];
        if (out_valid) {
          v[0] += rq[2 * q].x; v[1] += rq[2 * q].y; v[2] += rq[2 * q].z; v[3] += rq[2 * q].w;
          v[4] += rq[2 * q + 1].x; v[5] += rq[2 * q + 1].y; v[6] += rq[2 * q + 1].z; v[7] += rq[2 * q + 1].w;
          if (p.acc_in) {
            v[0] = aq[2 * q].x + v[0]; v[1] = aq[2 * q].y + v[1]; v[2] = aq[2 * q].z + v[2]; v[3] = aq[2 * q].w + v[3];
            v[4] = aq[2 * q + 1].x + v[4]; v[5] = aq[2 * q + 1].y + v[5]; v[6] = aq[2 * q + 1].z + v[6];
            v[7] = aq[2 * q + 1].w + v[7];
          }
          if (p.div != 0.f) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = v[e] / p.div;
          }
          if (p.out_f) {
            stg8(p.out_f + (((size_t)b * C8 + c8) * p.Tpf + p.f_halo + t_out) * 8, v);
          }
          if (p.out_plain) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              p.out_plain[((size_t)b * NC + c8 * 8 + e) * p.T + t_out] = p.plain_act ? leaky(v[e], p.plain_slope) : v[e];
          }
        }
        if (p.out_hi) {
          const size_t o = (((size_t)b * C8 + c8) * p.Tp + p.p_halo + t_out) * 8;
          if (out_valid) {
            float a[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = (p.plane_act ? leaky(v[e], p.plane_slope) : v[e]) * p.plane_scale;
            split_store8(p.out_hi + o, p.out_lo + o, a);
          } else {
            *reinterpret_cast<uint4*>(p.out_hi + o) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(p.out_lo + o) = make_uint4(0, 0, 0, 0);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

}  // namespace dissc

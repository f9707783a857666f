// Fused ResBlock1 pair for the C = 64 vocoder stage:
//
//     x' = x + conv2( lrelu( conv1( lrelu(x), dilation d ) ) )                    (sr/models.py:36-40)
//     [+ MRF accumulate / divide / next-stage leaky-relu in the same epilogue]   (:104-110)
//
// The unfused path moves 6 tensor passes per pair through HBM (c1: planes in, planes out; c2: planes in, fp32 residual
// in, fp32 out, planes out) and its c2 launches run at the HBM roof.  Here a pair reads the split planes of lrelu(x)
// once and writes the planes of lrelu(x') once; xt never leaves the SM and there is NO fp32 residual stream: the
// residual x is rebuilt in the epilogue from the same planes, x = min(y, y / slope) with y = hi + lo (22 significant
// bits, well inside the 1e-4 waveform tolerance; the planes of the tile were just fetched by TMA, so the re-read hits L2).
//
// Shared memory cannot hold both convs' weights at C = 64 (2 x k x 16 KB), so the taps stream through a ring of NS
// 16 KB slots in MMA issue order (W1 taps of tile s, W2 taps of tile s-1, ...), all CTAs hitting the same L2 lines.
//
// Per tile of M_out = 128-(k-1) output rows:
//   producer warp : bulk-TMA the planes x[t0-p2-p1 .. +R1) (R1 = 128+(k-1)d rows, 8 slabs x hi/lo) into the operand
//                   tile of the tile's worker group; weight taps into the ring
//   MMA thread    : conv1 as k shifted tap steps (2-MMA split of conv_tc.cuh) into TMEM acc1 (128 xt rows: t0-p2 ..)
//   worker group  : epilogue 1: acc1 -> +bias -> lrelu -> zero outside [0,T) -> fp16 hi/lo -> xt operand tile (smem)
//   MMA thread    : conv2 (dilation 1) from the xt tile into TMEM acc2
//   worker group  : epilogue 2: acc2 + bias + x [+ xs] [/ n] -> f32b and / or leaky-relu planes
// Two worker groups (8 warps each) alternate tiles; the MMA thread issues conv1(s), conv2(s-1), conv1(s+1), ... so one
// tile's epilogues overlap the other tile's MMAs.  TMEM: 2 groups x (acc1, acc2) x (main | cross) x 64 = 512 columns.
#pragma once
#include "conv_tc.cuh"

namespace dissc {

constexpr int kPair64Threads = 64 + 2 * 8 * 32;
constexpr uint32_t kPair64TapBytes = 8u * 2u * 64u * 16u;  // one tap: [c8 = 8][hi|lo][64][8] fp16 = 16 KB

struct Pair64Params {
  const __half* x_hi;   // planes [B][8][Tp][8] of lrelu(x, in_slope); zero rows outside [0, valid length)
  const __half* x_lo;
  const __half* w1;     // packed by pack_weights_tc with KB = 32: [cb = 2][tap][c8 = 4][hi|lo][64][8]
  const __half* w2;
  const float* b1;      // [64]
  const float* b2;
  float inv1, inv2;     // 2^-s of the two weight scalings, with 1 / in_scale resp. 1 / xt_scale folded in by the host
  float in_scale, xt_scale, plane_scale;  // power-of-two activation scales (TcParams): the input planes, the on-chip xt
                                          // tile and the output planes hold value * scale
  float in_inv_slope;   // 1 / in_slope: x = min(y, y * in_inv_slope)
  const float* acc_in;  // f32b [B][8][Tr][8] or null (MRF accumulator xs)
  float* out_f;         // f32b or null
  __half* out_hi;       // planes [B][8][Tp][8] or null: lrelu(result, plane_slope), zeros for rows in [valid, T)
  __half* out_lo;
  const int* lengths;
  int len_mul;
  int B, T, Tp, Tr, halo;
  int k, dil, NS;
  int tiles_per_b, n_tiles;
  float div, plane_slope;
};

__global__ void __launch_bounds__(kPair64Threads, 1) resblock_pair64_tc_kernel(const Pair64Params p) {
  constexpr int NC = 64, C8 = 8, KS = 4, WPG = 8, CH = 4;   // CH: 8-channel groups per worker warp
  constexpr uint32_t lbo_b = 2u * NC * 16;                  // [c8][hi|lo][NC][8]
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int k = p.k, d = p.dil;
  const int p2 = (k - 1) / 2, p1 = d * (k - 1) / 2;
  const int R1 = 128 + (k - 1) * d;   // x rows per tile
  const int R2 = 128 + (k - 1);       // xt rows addressable by conv2 (rows >= 128 stay zero)
  const int M_out = 128 - (k - 1);
  const uint32_t xop_plane = (uint32_t)C8 * R1 * 16, xop_bytes = 2 * xop_plane;
  const uint32_t xt_plane = (uint32_t)C8 * R2 * 16, xt_bytes = 2 * xt_plane;
  unsigned char* sXop = smem_raw;                   // [2][xop_bytes]
  unsigned char* sXt = sXop + 2 * xop_bytes;        // [2][xt_bytes]
  unsigned char* sW = sXt + 2 * xt_bytes;           // [NS][16 KB]
  float* s_b1 = reinterpret_cast<float*>(sW + (size_t)p.NS * kPair64TapBytes);
  float* s_b2 = s_b1 + NC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b2 + NC);
  uint64_t* xop_full = bars;         // [2]
  uint64_t* xop_empty = bars + 2;    // [2]
  uint64_t* acc1_full = bars + 4;    // [2]
  uint64_t* xt_full = bars + 6;      // [2]
  uint64_t* acc2_full = bars + 8;    // [2]
  uint64_t* acc2_empty = bars + 10;  // [2]
  uint64_t* w_full = bars + 12;      // [NS]
  uint64_t* w_empty = w_full + p.NS; // [NS]
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&xop_full[i], 1);
      mbar_init(&xop_empty[i], 1);
      mbar_init(&acc1_full[i], 1);
      mbar_init(&xt_full[i], WPG);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], WPG);
    }
    for (int i = 0; i < p.NS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < NC; i += kPair64Threads) {
    s_b1[i] = p.b1 ? p.b1[i] : 0.f;
    s_b2[i] = p.b2 ? p.b2[i] : 0.f;
  }
  // rows >= 128 of both xt tiles are read by conv2's discarded output rows only: keep them finite (zero)
  for (int i = tid; i < (int)(2 * xt_bytes / 16); i += kPair64Threads) reinterpret_cast<uint4*>(sXt)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  pdl_wait();                 // the prologue above touched only weights / shared memory; activations from here on
  pdl_launch_dependents();
  // TMEM columns: group g: acc1 at g*256 (main | cross), acc2 at g*256 + 128

  int n_mine = 0;
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) ++n_mine;

  if (warp == 0) {
    // ===================== producer =====================
    if (elect_one()) {
      uint32_t ws = 0, wph = 0;
      auto load_w = [&](const __half* w) {
        const unsigned char* wb = reinterpret_cast<const unsigned char*>(w);
        for (int j = 0; j < k; ++j) {
          mbar_wait(&w_empty[ws], wph ^ 1);
          mbar_arrive_expect_tx(&w_full[ws], kPair64TapBytes);
          unsigned char* dst = sW + (size_t)ws * kPair64TapBytes;
          tma_load_1d(dst, wb + (size_t)j * (kPair64TapBytes / 2), kPair64TapBytes / 2, &w_full[ws]);              // cb 0
          tma_load_1d(dst + kPair64TapBytes / 2, wb + (size_t)(k + j) * (kPair64TapBytes / 2), kPair64TapBytes / 2,
                      &w_full[ws]);                                                                                  // cb 1
          if (++ws == (uint32_t)p.NS) { ws = 0; wph ^= 1; }
        }
      };
      int tile = blockIdx.x;
      for (int s = 0; s <= n_mine; ++s, tile += gridDim.x) {
        if (s < n_mine) {
          const uint32_t g = s & 1, ph = (s >> 1) & 1;
          const int b = tile / p.tiles_per_b;
          const int t0 = (tile - b * p.tiles_per_b) * M_out;
          mbar_wait(&xop_empty[g], ph ^ 1);
          mbar_arrive_expect_tx(&xop_full[g], xop_bytes);
          const size_t off = (((size_t)b * C8) * p.Tp + p.halo + t0 - p2 - p1) * 8;
          unsigned char* dst = sXop + g * xop_bytes;
          for (int c8 = 0; c8 < C8; ++c8) {
            tma_load_1d(dst + (size_t)c8 * R1 * 16, p.x_hi + off + (size_t)c8 * p.Tp * 8, (uint32_t)R1 * 16, &xop_full[g]);
            tma_load_1d(dst + xop_plane + (size_t)c8 * R1 * 16, p.x_lo + off + (size_t)c8 * p.Tp * 8, (uint32_t)R1 * 16,
                        &xop_full[g]);
          }
          load_w(p.w1);
        }
        if (s >= 1) load_w(p.w2);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_n = (1u << 4) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * NC) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lbo_x = (uint32_t)R1 * 16, lbo_t = (uint32_t)R2 * 16;
      const uint32_t sW0 = smem_u32(sW);
      uint32_t ws = 0, wph = 0;
      auto conv = [&](uint32_t a_addr, uint32_t a_plane, uint32_t lbo_a, int tap_rows, uint32_t d_main) {
        const uint32_t a_kstep = (2 * lbo_a) >> 4, b_kstep = (2 * lbo_b) >> 4, lo_off = a_plane >> 4;
        uint32_t accum = 0, ad_t = umma_desc_lo(a_addr, lbo_a);
        for (int j = 0; j < k; ++j, ad_t += (uint32_t)tap_rows) {
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          uint32_t ad = ad_t, wd = umma_desc_lo(sW0 + ws * kPair64TapBytes, lbo_b);
#pragma unroll
          for (int ks = 0; ks < KS; ++ks, ad += a_kstep, wd += b_kstep) {
            umma_f16(d_main, umma_desc(ad), umma_desc(wd), idesc_2n, accum);            // [main | cross]
            umma_f16(d_main + NC, umma_desc(ad + lo_off), umma_desc(wd), idesc_n, 1);   // cross += lo * hi
            accum = 1;
          }
          umma_commit(&w_empty[ws]);
          if (++ws == (uint32_t)p.NS) { ws = 0; wph ^= 1; }
        }
      };
      for (int s = 0; s <= n_mine; ++s) {
        if (s < n_mine) {
          const uint32_t g = s & 1, ph = (s >> 1) & 1;
          mbar_wait(&xop_full[g], ph);
          tc_fence_after();
          conv(smem_u32(sXop + g * xop_bytes), xop_plane, lbo_x, d, tmem_base + g * 256u);
          umma_commit(&xop_empty[g]);
          umma_commit(&acc1_full[g]);
        }
        if (s >= 1) {
          const uint32_t sp = (uint32_t)(s - 1), g = sp & 1, ph = (sp >> 1) & 1;
          mbar_wait(&xt_full[g], ph);
          mbar_wait(&acc2_empty[g], ph ^ 1);
          tc_fence_after();
          conv(smem_u32(sXt + g * xt_bytes), xt_plane, lbo_t, 1, tmem_base + g * 256u + 128u);
          umma_commit(&acc2_full[g]);
        }
      }
    }
  } else {
    // ===================== worker groups: epilogue 1 -> epilogue 2 =====================
    const int g = (warp - 2) / WPG;         // worker group
    const int wi = (warp - 2) - g * WPG;    // warp inside the group
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int c8_0 = (wi >> 2) * CH;        // first of the CH channel groups this warp handles
    const int row = quarter * 32 + lane;    // TMEM lane = tile row
    unsigned char* xt = sXt + g * xt_bytes;
    const uint32_t t_acc1 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 256u;
    const uint32_t t_acc2 = t_acc1 + 128u;
    const float slope = p.plane_slope;
    const float un_in = 1.f / p.in_scale, inv_in = p.in_inv_slope * un_in;   // residual: x = min(y, y / slope) / in_scale
    uint32_t it = 0;
    const int first = blockIdx.x + g * gridDim.x, step = 2 * gridDim.x;
    for (int tile = first; tile < p.n_tiles; tile += step, ++it) {
      const uint32_t ph = it & 1;
      const int b = tile / p.tiles_per_b;
      const int t0 = (tile - b * p.tiles_per_b) * M_out;
      const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
      // ---- epilogue 1: acc1 -> xt tile
      mbar_wait(&acc1_full[g], ph);
      tc_fence_after();
      {
        const int t = t0 - p2 + row;
        const bool v_ok = t >= 0 && t < Tvalid;
#pragma unroll
        for (int qb = 0; qb < CH; qb += 2) {
          float m[2][8], x8[2][8];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            tmem_ld8(t_acc1 + (c8_0 + qb + q) * 8, m[q]);
            tmem_ld8(t_acc1 + NC + (c8_0 + qb + q) * 8, x8[q]);
          }
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c8 = c8_0 + qb + q;
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
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&xt_full[g]);
      // ---- residual prefetch for epilogue 2: the planes of lrelu(x) at row t0+row (fetched by this tile's TMA a moment
      //      ago: L2-hot); issued now so the latency hides behind conv2
      const int t_out = t0 + row;
      const bool out_valid = row < M_out && t_out < Tvalid;
      const bool out_inb = row < M_out && t_out < p.T;
      uint4 rh[CH], rl[CH];
      if (out_valid) {
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          const size_t pi = (((size_t)b * C8 + c8_0 + q) * p.Tp + p.halo + t_out) * 8;
          rh[q] = *reinterpret_cast<const uint4*>(p.x_hi + pi);
          rl[q] = *reinterpret_cast<const uint4*>(p.x_lo + pi);
        }
      }
      // ---- epilogue 2: acc2 + bias + residual [+ xs] [/ n] -> outputs
      mbar_wait(&acc2_full[g], ph);
      tc_fence_after();
#pragma unroll
      for (int qb = 0; qb < CH; qb += 2) {
        float m2[2][8], y2[2][8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          tmem_ld8(t_acc2 + (c8_0 + qb + q) * 8, m2[q]);
          tmem_ld8(t_acc2 + NC + (c8_0 + qb + q) * 8, y2[q]);
        }
        tmem_ld_wait();
        if (qb + 2 >= CH) {
          // the accumulator is in registers: release it before the (long) store phase
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc2_empty[g]);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c8 = c8_0 + qb + q;
          if (!out_inb) continue;
          const size_t po = (((size_t)b * C8 + c8) * p.Tp + p.halo + t_out) * 8;
          if (!out_valid) {
            if (p.out_hi) {
              *reinterpret_cast<uint4*>(p.out_hi + po) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(p.out_lo + po) = make_uint4(0, 0, 0, 0);
            }
            continue;
          }
          float v[8];
          {
            const uint32_t hh[4] = {rh[qb + q].x, rh[qb + q].y, rh[qb + q].z, rh[qb + q].w};
            const uint32_t ll[4] = {rl[qb + q].x, rl[qb + q].y, rl[qb + q].z, rl[qb + q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hh[e]));
              const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&ll[e]));
              const float y0 = fh.x + fl.x, y1 = fh.y + fl.y;
              v[2 * e] = fminf(y0 * un_in, y0 * inv_in);          // inverse leaky-relu: x = y (y >= 0), y / slope (y < 0)
              v[2 * e + 1] = fminf(y1 * un_in, y1 * inv_in);
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += (m2[q][e] + y2[q][e]) * p.inv2 + s_b2[c8 * 8 + e];
          const size_t fo = (((size_t)b * C8 + c8) * p.Tr + t_out) * 8;
          if (p.acc_in) {
            float4 a0, a1;
            ldg8(p.acc_in + fo, a0, a1);
            v[0] = a0.x + v[0]; v[1] = a0.y + v[1]; v[2] = a0.z + v[2]; v[3] = a0.w + v[3];
            v[4] = a1.x + v[4]; v[5] = a1.y + v[5]; v[6] = a1.z + v[6]; v[7] = a1.w + v[7];
          }
          if (p.div != 0.f) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = v[e] / p.div;
          }
          if (p.out_f) stg8(p.out_f + fo, v);
          if (p.out_hi) {
            float a[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = leaky(v[e], slope) * p.plane_scale;
            split_store8(p.out_hi + po, p.out_lo + po, a);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace dissc

// Fused ResBlock1 pair for the C = 16 vocoder stage, conv2 in TWO-PHASE form.
//
//     x' = x + conv2( lrelu( conv1( lrelu(x), dilation d ) ) )                    (sr/models.py:36-40)
//     [+ MRF accumulate / divide / final leaky-relu in the same epilogue]        (:104-110)
//
// Why another kernel: at C = 16 every tcgen05.mma of resblock_tc.cuh has N = 32 / 16 columns, and such an MMA costs
// its 4 KB A-operand shared-memory read (32 cycles), not its math (8-17 cycles): ncu shows the tensor pipe 90 % busy
// in the k = 7 / 11 pairs doing 1/3 useful work.  The undilated conv2 can be re-blocked so that ONE GEMM row carries
// TWO time steps:
//
//     X2[rho] = [ xt[2 rho], xt[2 rho + 1] ]   (K = 2C),      Y2[rho] = [ y[2 rho], y[2 rho + 1] ]   (N = 2C)
//     Y2[rho] = sum_s X2[rho + s] * V_s ,   V_s[(q', ci), (q, co)] = W[co, ci, j = 2 s + q' - q]  (0 outside 0 <= j < k)
//
// (k + 1) / 2 row shifts instead of k taps, every V_s dense: half as many MMAs per output sample, each with N = 64 / 32
// columns -- 1.7x less tensor-pipe time for conv2.  conv1 keeps its dilation (the trick needs d = 1) and runs as two
// ordinary M = 128 sub-tiles whose epilogue writes xt straight into the two-phase operand layout.
//
// Per tile of M_out = 256 - (k - 1) output samples:
//   producer warp : bulk-TMA the fp32 tile x[t0-p2-p1 .. +R1) (R1 = 256+(k-1)d rows, 2 slabs) into a staging buffer
//   worker group  : (4 warps) convert staging -> lrelu -> fp16 hi/lo -> operand tile (zero outside [0,T))
//   MMA thread    : conv1 on sub-tiles a, b (k shifted MMAs each) into TMEM acc1a / acc1b
//   worker group  : epilogue 1: acc1 -> +bias -> lrelu -> zero outside [0,T) -> hi/lo -> two-phase xt tile
//   MMA thread    : conv2: (k + 1) / 2 shifts x 2 K-steps from the xt tile into TMEM acc2 (128 rows x [phase 0 | phase 1])
//   worker group  : epilogue 2: acc2 + bias + x (residual, L2-hot re-read) [+ xs] [/ n] -> fp32 / planes / plain output
// THREE worker groups (4 warps each) rotate over tiles: with its 256-sample tiles the kernel needs 205 KB of shared
// memory, i.e. one CTA per SM, and two tiles in flight left the tensor pipe idle 26 % of the time (first version of this
// kernel); the MMA thread issues conv1(s), conv2(s-2), conv1(s+1), ...  All weights stay resident in shared memory.
#pragma once
#include "conv_tc.cuh"

namespace dissc {

constexpr int kP16Groups = 3;                       // tiles in flight per CTA
constexpr int kP16Wpg = 4;                          // worker warps per group
constexpr int kP16Threads = 64 + kP16Groups * kP16Wpg * 32;

struct Pair16Params {
  const float* x;       // f32h [B][2][Tpf][8]
  const __half* w1;     // conv1, packed [tap][c8 = 2][hi|lo][16][8]  (pack_weights_tc, KB = 16)
  const __half* v2;     // conv2 in two-phase form, packed [shift][chunk = (q', c8)][hi|lo][32][8]
  const float* b1;      // [16]
  const float* b2;
  float inv1, inv2;     // 2^-s of the two weight scalings
  const float* acc_in;  // f32h or null (MRF accumulator xs)
  float* out_f;         // f32h or null
  __half* out_hi;       // planes [B][2][Tp][8] or null (leaky-relu(plane_slope))
  __half* out_lo;
  float* out_plain;     // (B, 16, T) fp32 or null (leaky-relu(plain_slope) iff plain_act)
  const int* lengths;
  int len_mul;
  int B, T, Tpf, f_halo, Tp, p_halo;
  int k, dil;
  int tiles_per_b, n_tiles;
  float div;
  int plain_act;
  float plane_slope, plain_slope;
};

__global__ void __launch_bounds__(kP16Threads, 1) resblock_pair16_tc_kernel(const Pair16Params p) {
  constexpr int C = 16, C8 = 2, WPG = kP16Wpg, G = kP16Groups;
  constexpr uint32_t lbo_w1 = 2u * C * 16;        // conv1 weights: [c8][hi|lo][16][8]
  constexpr uint32_t w1_tap_bytes = C8 * lbo_w1;  // 1 KB
  constexpr uint32_t lbo_v = 2u * 2 * C * 16;     // conv2 V: [chunk][hi|lo][32][8] = 1 KB per chunk
  constexpr uint32_t v_shift_bytes = 4 * lbo_v;   // 4 chunks (q', c8) per shift
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int k = p.k, d = p.dil;
  const int p2 = (k - 1) / 2, p1 = d * (k - 1) / 2;
  const int NSH = (k + 1) / 2;         // row shifts of the two-phase conv2
  const int R1 = 256 + (k - 1) * d;    // x rows per tile
  const int R2 = 128 + NSH;            // rows of the two-phase xt tile (rows >= 128 stay zero)
  const int M_out = 256 - (k - 1);
  const uint32_t stg_bytes = (uint32_t)C8 * R1 * 32;
  const uint32_t xop_plane = (uint32_t)C8 * R1 * 16, xop_bytes = 2 * xop_plane;
  const uint32_t xt_plane = (uint32_t)4 * R2 * 16, xt_bytes = 2 * xt_plane;   // 4 chunks: (phase, c8)
  const uint32_t w1_bytes = (uint32_t)k * w1_tap_bytes, v_bytes = (uint32_t)NSH * v_shift_bytes;
  unsigned char* sStg = smem_raw;                   // [G][stg_bytes]
  unsigned char* sXop = sStg + G * stg_bytes;       // [G][xop_bytes]
  unsigned char* sXt = sXop + G * xop_bytes;        // [G][xt_bytes]
  unsigned char* sW1 = sXt + G * xt_bytes;
  unsigned char* sV = sW1 + w1_bytes;
  float* s_b1 = reinterpret_cast<float*>(sV + v_bytes);
  float* s_b2 = s_b1 + C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b2 + C);
  uint64_t* stg_full = bars;             // [G]
  uint64_t* stg_empty = bars + G;        // [G]
  uint64_t* xop_full = bars + 2 * G;     // [G]
  uint64_t* xop_empty = bars + 3 * G;    // [G]
  uint64_t* acc1_full = bars + 4 * G;    // [G]
  uint64_t* xt_full = bars + 5 * G;      // [G]
  uint64_t* acc2_full = bars + 6 * G;    // [G]
  uint64_t* acc2_empty = bars + 7 * G;   // [G]
  uint64_t* w_full = bars + 8 * G;       // [1]
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < G; ++i) {
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
  for (int i = tid; i < C; i += kP16Threads) {
    s_b1[i] = p.b1 ? p.b1[i] : 0.f;
    s_b2[i] = p.b2 ? p.b2[i] : 0.f;
  }
  // rows >= 128 of both xt tiles feed conv2's discarded output rows only: keep them finite (zero)
  for (int i = tid; i < (int)(G * xt_bytes / 16); i += kP16Threads) reinterpret_cast<uint4*>(sXt)[i] = make_uint4(0, 0, 0, 0);
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
  // TMEM columns, group g at g*128: acc1a [0,32) (main | cross), acc1b [32,64), acc2 [64,128) (main 32 | cross 32)

  if (warp == 0) {
    // ===================== producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&w_full[0], w1_bytes + v_bytes);
      tma_load_1d(sW1, p.w1, w1_bytes, &w_full[0]);
      tma_load_1d(sV, p.v2, v_bytes, &w_full[0]);
      uint32_t s = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++s) {
        const uint32_t g = s % G, ph = (s / G) & 1;
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
      constexpr uint32_t idesc_16 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lbo_x = (uint32_t)R1 * 16, lbo_t = (uint32_t)R2 * 16;
      const uint32_t w1d = umma_desc_lo(smem_u32(sW1), lbo_w1), vd0 = umma_desc_lo(smem_u32(sV), lbo_v);
      mbar_wait(&w_full[0], 0);
      int n_mine = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) ++n_mine;
      // issue order conv1(s), conv2(s - (G-1)): a tile's epilogue 1 has G-1 conv slots to finish before its conv2
      for (int s = 0; s < n_mine + G - 1; ++s) {
        if (s < n_mine) {
          const uint32_t g = (uint32_t)s % G, ph = ((uint32_t)s / G) & 1;
          mbar_wait(&xop_full[g], ph);
          tc_fence_after();
          const uint32_t xa = smem_u32(sXop + g * xop_bytes);
          // conv1: xt row tau (0..255) = sum_j x_row[tau + j d]; sub-tile u covers tau = 128 u .. 128 u + 127
#pragma unroll 1
          for (int u = 0; u < 2; ++u) {
            const uint32_t d_acc = tmem_base + g * 128u + (uint32_t)u * 32u;
            uint32_t ad = umma_desc_lo(xa + (uint32_t)u * 128u * 16u, lbo_x), wd = w1d, accum = 0;
            for (int j = 0; j < k; ++j, ad += (uint32_t)d, wd += (w1_tap_bytes >> 4)) {
              umma_f16(d_acc, umma_desc(ad), umma_desc(wd), idesc_32, accum);                        // [main | cross]
              umma_f16(d_acc + C, umma_desc(ad + (xop_plane >> 4)), umma_desc(wd), idesc_16, 1);     // cross += lo * hi
              accum = 1;
            }
          }
          umma_commit(&xop_empty[g]);
          umma_commit(&acc1_full[g]);
        }
        if (s >= G - 1) {
          const uint32_t sp = (uint32_t)(s - (G - 1)), g = sp % G, ph = (sp / G) & 1;
          mbar_wait(&xt_full[g], ph);
          mbar_wait(&acc2_empty[g], ph ^ 1);
          tc_fence_after();
          const uint32_t d_acc = tmem_base + g * 128u + 64u;
          uint32_t ad_s = umma_desc_lo(smem_u32(sXt + g * xt_bytes), lbo_t), vd_s = vd0, accum = 0;
          for (int sh = 0; sh < NSH; ++sh, ad_s += 1u, vd_s += (v_shift_bytes >> 4)) {
            uint32_t ad = ad_s, vd = vd_s;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks, ad += (2 * lbo_t) >> 4, vd += (2 * lbo_v) >> 4) {
              umma_f16(d_acc, umma_desc(ad), umma_desc(vd), idesc_64, accum);                          // [main 32 | cross 32]
              umma_f16(d_acc + 32, umma_desc(ad + (xt_plane >> 4)), umma_desc(vd), idesc_32, 1);       // cross += lo * hi
              accum = 1;
            }
          }
          umma_commit(&acc2_full[g]);
        }
      }
    }
  } else {
    // ===================== worker groups: convert -> epilogue 1 -> epilogue 2 =====================
    const int g = (warp - 2) / WPG;         // worker group
    const int wi = (warp - 2) - g * WPG;    // warp inside the group
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int wt = wi * 32 + lane;          // thread index inside the group
    const int row = quarter * 32 + lane;    // TMEM lane = tile row
    unsigned char* stg = sStg + g * stg_bytes;
    unsigned char* xop = sXop + g * xop_bytes;
    unsigned char* xt = sXt + g * xt_bytes;
    const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 128u;
    const unsigned r1_magic = 0xFFFFFFFFu / (unsigned)R1 + 1u;   // umulhi(item, magic) == item / R1 for item < 2^16
    // convert: fp32 staging tile -> lrelu -> fp16 hi/lo operand tile (rows outside [0, Tvalid) are zeros)
    auto convert = [&](int tile, uint32_t ph) {
      const int b = tile / p.tiles_per_b;
      const int t0 = (tile - b * p.tiles_per_b) * M_out;
      const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
      mbar_wait(&stg_full[g], ph);
      mbar_wait(&xop_empty[g], ph ^ 1);
      const int tx0 = t0 - p2 - p1;
      for (int item = wt; item < C8 * R1; item += WPG * 32) {
        const int c8 = (int)__umulhi((unsigned)item, r1_magic), i = item - c8 * R1;
        const int t = tx0 + i;
        if (t >= 0 && t < Tvalid) {
          const float4 a = *reinterpret_cast<const float4*>(stg + (size_t)item * 32);
          const float4 c = *reinterpret_cast<const float4*>(stg + (size_t)item * 32 + 16);
          float v[8];
          v[0] = leaky(a.x, 0.1f); v[1] = leaky(a.y, 0.1f); v[2] = leaky(a.z, 0.1f); v[3] = leaky(a.w, 0.1f);
          v[4] = leaky(c.x, 0.1f); v[5] = leaky(c.y, 0.1f); v[6] = leaky(c.z, 0.1f); v[7] = leaky(c.w, 0.1f);
          split_store8(reinterpret_cast<__half*>(xop + (size_t)item * 16),
                       reinterpret_cast<__half*>(xop + xop_plane + (size_t)item * 16), v);
        } else {
          *reinterpret_cast<uint4*>(xop + (size_t)item * 16) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(xop + xop_plane + (size_t)item * 16) = make_uint4(0, 0, 0, 0);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&xop_full[g]);
        mbar_arrive(&stg_empty[g]);
      }
    };
    uint32_t it = 0;
    const int first = blockIdx.x + g * gridDim.x, step = G * gridDim.x;
    if (first < p.n_tiles) convert(first, 0);
    for (int tile = first; tile < p.n_tiles; tile += step, ++it) {
      const uint32_t ph = it & 1;
      const int b = tile / p.tiles_per_b;
      const int t0 = (tile - b * p.tiles_per_b) * M_out;
      const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
      // ---- residual prefetch for epilogue 2 (GEMM row `row` carries the samples t0 + 2 row + q, q = 0, 1)
      float4 rq[2][C8 * 2], aq[2][C8 * 2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int o_idx = 2 * row + q, t_out = t0 + o_idx;
        if (o_idx < M_out && t_out < Tvalid) {
#pragma unroll
          for (int c8 = 0; c8 < C8; ++c8) {
            const size_t fi = (((size_t)b * C8 + c8) * p.Tpf + p.f_halo + t_out) * 8;
            ldg8(p.x + fi, rq[q][2 * c8], rq[q][2 * c8 + 1]);
            if (p.acc_in) ldg8(p.acc_in + fi, aq[q][2 * c8], aq[q][2 * c8 + 1]);
          }
        }
      }
      // ---- epilogue 1: acc1 (sub-tile u, row `row`) -> xt at tile time tau = 128 u + row, two-phase layout:
      //      chunk (tau & 1, c8), row tau >> 1
      mbar_wait(&acc1_full[g], ph);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int tau = 128 * u + row;
        const int t = t0 - p2 + tau;
        const bool v_ok = t >= 0 && t < Tvalid;
        float m[C8][8], x8[C8][8];
#pragma unroll
        for (int c8 = 0; c8 < C8; ++c8) {
          tmem_ld8(t_base + (uint32_t)u * 32u + c8 * 8, m[c8]);
          tmem_ld8(t_base + (uint32_t)u * 32u + C + c8 * 8, x8[c8]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < C8; ++c8) {
          const size_t o = ((size_t)((tau & 1) * C8 + c8) * R2 + (tau >> 1)) * 16;
          if (v_ok) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = leaky((m[c8][e] + x8[c8][e]) * p.inv1 + s_b1[c8 * 8 + e], 0.1f);
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
      // ---- epilogue 2: acc2 columns [q*16 + co] (main) and [32 + q*16 + co] (cross) + bias + residual [+ xs] [/ n]
      mbar_wait(&acc2_full[g], ph);
      tc_fence_after();
      float m2[2][C8][8], y2[2][C8][8];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int c8 = 0; c8 < C8; ++c8) {
          tmem_ld8(t_base + 64u + (uint32_t)q * 16u + c8 * 8, m2[q][c8]);
          tmem_ld8(t_base + 96u + (uint32_t)q * 16u + c8 * 8, y2[q][c8]);
        }
      tmem_ld_wait();
      // the accumulator is in registers: release it before the (long) store phase
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_empty[g]);
      const int t_e = t0 + 2 * row;                 // the row's two samples: t_e (q = 0) and t_e + 1 (q = 1); t_e is even
      bool inb[2], valid[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        inb[q] = 2 * row + q < M_out && t_e + q < p.T;
        valid[q] = 2 * row + q < M_out && t_e + q < Tvalid;
      }
      if (inb[0]) {
#pragma unroll
        for (int c8 = 0; c8 < C8; ++c8) {
          float v[2][8];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[q][e] = (m2[q][c8][e] + y2[q][c8][e]) * p.inv2 + s_b2[c8 * 8 + e];
            if (!valid[q]) continue;
            v[q][0] += rq[q][2 * c8].x; v[q][1] += rq[q][2 * c8].y; v[q][2] += rq[q][2 * c8].z; v[q][3] += rq[q][2 * c8].w;
            v[q][4] += rq[q][2 * c8 + 1].x; v[q][5] += rq[q][2 * c8 + 1].y; v[q][6] += rq[q][2 * c8 + 1].z;
            v[q][7] += rq[q][2 * c8 + 1].w;
            if (p.acc_in) {
              v[q][0] = aq[q][2 * c8].x + v[q][0]; v[q][1] = aq[q][2 * c8].y + v[q][1]; v[q][2] = aq[q][2 * c8].z + v[q][2];
              v[q][3] = aq[q][2 * c8].w + v[q][3]; v[q][4] = aq[q][2 * c8 + 1].x + v[q][4];
              v[q][5] = aq[q][2 * c8 + 1].y + v[q][5]; v[q][6] = aq[q][2 * c8 + 1].z + v[q][6];
              v[q][7] = aq[q][2 * c8 + 1].w + v[q][7];
            }
            if (p.div != 0.f) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[q][e] = v[q][e] / p.div;
            }
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (!inb[q]) continue;
            const size_t po = (((size_t)b * C8 + c8) * p.Tp + p.p_halo + t_e + q) * 8;
            if (!valid[q]) {
              if (p.out_hi) {
                *reinterpret_cast<uint4*>(p.out_hi + po) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(p.out_lo + po) = make_uint4(0, 0, 0, 0);
              }
              continue;
            }
            if (p.out_f) stg8(p.out_f + (((size_t)b * C8 + c8) * p.Tpf + p.f_halo + t_e + q) * 8, v[q]);
            if (p.out_hi) {
              float a[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) a[e] = leaky(v[q][e], p.plane_slope);
              split_store8(p.out_hi + po, p.out_lo + po, a);
            }
          }
          if (p.out_plain) {
            // (B, 16, T) fp32: the row's two samples are adjacent in memory -> one 8-byte store per channel (t_e and T even)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float* o = p.out_plain + ((size_t)b * C + c8 * 8 + e) * p.T + t_e;
              const float y0 = p.plain_act ? leaky(v[0][e], p.plain_slope) : v[0][e];
              const float y1 = p.plain_act ? leaky(v[1][e], p.plain_slope) : v[1][e];
              if (valid[1] && !(p.T & 1))
                *reinterpret_cast<float2*>(o) = make_float2(y0, y1);
              else {
                if (valid[0]) o[0] = y0;
                if (valid[1]) o[1] = y1;
              }
            }
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

// Fused ResBlock1 pair for the C = 16 vocoder stage with TWO SAMPLES PER GEMM ROW:
//
//     x' = x + conv2( lrelu( conv1( lrelu(x), dilation d ) ) )                    (sr/models.py:36-40)
//     [+ MRF accumulate / divide / next-stage leaky-relu in the same epilogue]   (:104-110)
//
// Same dataflow, HBM layouts and parameters as resblock_tc.cuh (fp32 "f32h" tile in by TMA, on-chip leaky-relu + fp16
// hi/lo split, conv1 -> TMEM -> epilogue 1 -> xt tile in shared memory -> conv2 -> epilogue 2, two worker groups
// alternating tiles).  What changes is the GEMM shape.  With one sample per row a C = 16 conv is k taps of 128 x 16 x 16
// MMAs (N = 32 / 16 with the hi|lo split): each costs the 4 KB read of its A operand from shared memory, not its math,
// and ncu shows the tensor pipe 90 % "busy" at 28 % of the tensor roof (profiles/README.md, r01_d s4k11).  Here a GEMM
// row is a PAIR of consecutive samples: K = N = 2C = 32, and a k-tap conv becomes S = (k+1)/2 row shifts of a
// block-Toeplitz weight
//
//     out[2r + qo] = sum_j W_j x[2r + qo + j]   ->   D[r, (qo, co)] = sum_{s < S} A[r + s, (qi, ci)] B_s[(qi, ci), (qo, co)],
//     B_s[(qi, ci), (qo, co)] = W_{2s + qi - qo}[ci -> co]   (zero outside 0 <= j < k: 2 of the 4S blocks)
//
// so the same output costs S x 2 k-steps x (N = 64, N = 32) MMAs instead of k x (N = 32, N = 16): 0.55x the A-operand
// bytes, MMAs that are wide enough to be limited by their math.  A DILATED conv1 (d = 3, 5) fits the same form after
// decimation: the samples of a tile are stored phase by phase (phase = index mod d, then pairs of consecutive
// decimated samples per row), which turns the dilation-d conv into d independent dilation-1 convs over the decimated
// index -- the operand tile is written on chip by the convert step, so its order is free.  Epilogue 1 undoes the
// permutation while it writes the xt tile (natural order, pairs per row) for conv2.
//
// Tile: U xt samples (even, chosen by the host so that all d phases fit the 128 TMEM lanes), M_out = U - (k-1) outputs.
#pragma once
#include "conv_tc.cuh"

namespace dissc {

constexpr int kPack2MaxDil = 8;
// producer warp, MMA-issuer warp, G worker groups of 8 warps (one tile in flight per group)
__host__ __device__ constexpr int pack2_threads(int G) { return 64 + G * 8 * 32; }

struct Pack2Params {
  const float* x;       // f32h [B][2][Tpf][8]
  const __half* w1;     // block-Toeplitz weights, packed [shift][k8 = (qi, c8)][hi|lo][(qo, co) = 32][8]
  const __half* w2;
  const float* b1;      // [16]
  const float* b2;
  float inv1, inv2;     // 2^-s of the two weight scalings, with 1 / in_scale resp. 1 / xt_scale folded in by the host
  float in_scale, xt_scale, plane_scale;  // power-of-two activation scales (TcParams)
  const float* acc_in;  // f32h or null (MRF accumulator xs)
  float* out_f;         // f32h or null
  __half* out_hi;       // planes [B][2][Tp][8] or null (leaky-relu(plane_slope) iff plane_act)
  __half* out_lo;
  float* out_plain;     // (B, 16, T) fp32 or null (leaky-relu(plain_slope) iff plain_act)
  const int* lengths;
  int len_mul;
  int B, T, Tpf, f_halo, Tp, p_halo;
  int k, dil, S;        // taps, dilation of conv1, row shifts per conv = (k+1)/2
  int U, M_out;         // xt samples / output samples per tile
  int base[kPack2MaxDil];   // first operand row of phase f
  int n_out[kPack2MaxDil];  // xt samples of phase f in a tile: ceil((U - f) / d)
  unsigned d_magic;     // ceil(2^32 / d): umulhi(v, d_magic) == v / d for v < 2^16
  int tiles_per_b, n_tiles;
  float div;
  int plane_act, plain_act;
  float plane_slope, plain_slope;
};

template <int G>
__global__ void __launch_bounds__(pack2_threads(G), 1) resblock_pack2_tc_kernel(const Pack2Params p) {
  constexpr int kPack2Threads = pack2_threads(G);
  constexpr int C = 16, C8 = 2;            // real channels
  constexpr int NG = 32, G8 = 4, KS = 2;   // GEMM width (two samples x C), its 8-wide K groups, its 16-wide K steps
  constexpr int WPG = 8, CH = 2;
  constexpr uint32_t lbo_b = 2u * NG * 16;
  constexpr uint32_t w_tap_bytes = (uint32_t)G8 * lbo_b;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int k = p.k, d = p.dil, S = p.S;
  const int p2 = (k - 1) / 2, p1 = d * (k - 1) / 2;
  const int Lx = p.U + (k - 1) * d;   // x samples per tile
  const int RX = 128 + S - 1;         // operand rows addressable by conv1 / conv2 (rows past the data stay zero)
  const int M_out = p.M_out;
  const uint32_t stg_bytes = (uint32_t)C8 * Lx * 32;
  const uint32_t xop_plane = (uint32_t)G8 * RX * 16, xop_bytes = 2 * xop_plane;
  const uint32_t xt_plane = xop_plane, xt_bytes = xop_bytes;
  const uint32_t w_bytes = (uint32_t)S * w_tap_bytes;
  unsigned char* sStg = smem_raw;                   // [G][stg_bytes]
  unsigned char* sXop = sStg + G * stg_bytes;       // [G][xop_bytes]
  unsigned char* sXt = sXop + G * xop_bytes;        // [G][xt_bytes]
  unsigned char* sW1 = sXt + G * xt_bytes;
  unsigned char* sW2 = sW1 + w_bytes;
  float* s_b1 = reinterpret_cast<float*>(sW2 + w_bytes);
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
  for (int i = tid; i < C; i += kPack2Threads) {
    s_b1[i] = p.b1 ? p.b1[i] : 0.f;
    s_b2[i] = p.b2 ? p.b2[i] : 0.f;
  }
  // every operand slot the convert step / epilogue 1 never writes (odd phase tails, rows past the data that only
  // discarded output rows read) must hold finite values: zero both operand tile pairs once
  for (int i = tid; i < (int)((G * xop_bytes + G * xt_bytes) / 16); i += kPack2Threads)
    reinterpret_cast<uint4*>(sXop)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(G <= 2 ? 256 : 512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  pdl_wait();                 // the prologue above touched only weights / shared memory; activations from here on
  pdl_launch_dependents();
  // TMEM columns: group g: acc1 at g*4NG (main | cross), acc2 at g*4NG + 2NG;  column inside an accumulator = qo*C + co

  if (warp == 0) {
    // ===================== producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&w_full[0], 2 * w_bytes);
      tma_load_1d(sW1, p.w1, w_bytes, &w_full[0]);
      tma_load_1d(sW2, p.w2, w_bytes, &w_full[0]);
      uint32_t g = 0, ph = 0;   // tile s of this CTA belongs to group s % G, its phase is (s / G) & 1
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_b;
        const int t0 = (tile - b * p.tiles_per_b) * M_out;
        mbar_wait(&stg_empty[g], ph ^ 1);
        mbar_arrive_expect_tx(&stg_full[g], stg_bytes);
        const float* src = p.x + (((size_t)b * C8) * p.Tpf + p.f_halo + t0 - p2 - p1) * 8;
        for (int c8 = 0; c8 < C8; ++c8)
          tma_load_1d(sStg + g * stg_bytes + (size_t)c8 * Lx * 32, src + (size_t)c8 * p.Tpf * 8, (uint32_t)Lx * 32,
                      &stg_full[g]);
        if (++g == (uint32_t)G) { g = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_n = (1u << 4) | ((uint32_t)(NG >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * NG) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lbo_a = (uint32_t)RX * 16;
      const uint32_t w1d = umma_desc_lo(smem_u32(sW1), lbo_b), w2d = umma_desc_lo(smem_u32(sW2), lbo_b);
      mbar_wait(&w_full[0], 0);
      // both convs are dilation-free in row space: shift s reads operand rows [s, s + 128)
      auto conv = [&](uint32_t a_addr, uint32_t wdesc, uint32_t d_main) {
        const uint32_t a_kstep = (2 * lbo_a) >> 4, b_kstep = (2 * lbo_b) >> 4, lo_off = xop_plane >> 4;
        uint32_t accum = 0, ad_t = umma_desc_lo(a_addr, lbo_a), wd_t = wdesc;
        for (int s = 0; s < S; ++s, ad_t += 1u, wd_t += (w_tap_bytes >> 4)) {
          uint32_t ad = ad_t, wd = wd_t;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks, ad += a_kstep, wd += b_kstep) {
            umma_f16(d_main, umma_desc(ad), umma_desc(wd), idesc_2n, accum);            // [main | cross], N = 64
            umma_f16(d_main + NG, umma_desc(ad + lo_off), umma_desc(wd), idesc_n, 1);   // cross += lo * hi, N = 32
            accum = 1;
          }
        }
      };
      int n_mine = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) ++n_mine;
      // conv1 runs kLead tiles ahead of conv2: conv1(s), conv2(s - kLead), conv1(s + 1), ... so a tile's epilogue 1 has
      // kLead conv1's worth of tensor time before the MMA thread needs its xt
      constexpr int kLead = G - 1;
      uint32_t g1 = 0, ph1 = 0, g2 = 0, ph2 = 0;
      for (int s = 0; s < n_mine + kLead; ++s) {
        if (s < n_mine) {
          mbar_wait(&xop_full[g1], ph1);
          tc_fence_after();
          conv(smem_u32(sXop + g1 * xop_bytes), w1d, tmem_base + g1 * 4u * NG);
          umma_commit(&xop_empty[g1]);
          umma_commit(&acc1_full[g1]);
          if (++g1 == (uint32_t)G) { g1 = 0; ph1 ^= 1; }
        }
        if (s >= kLead) {
          mbar_wait(&xt_full[g2], ph2);
          mbar_wait(&acc2_empty[g2], ph2 ^ 1);
          tc_fence_after();
          conv(smem_u32(sXt + g2 * xt_bytes), w2d, tmem_base + g2 * 4u * NG + 2u * NG);
          umma_commit(&acc2_full[g2]);
          if (++g2 == (uint32_t)G) { g2 = 0; ph2 ^= 1; }
        }
      }
    }
  } else {
    // ===================== worker groups: convert -> epilogue 1 -> epilogue 2 =====================
    const int g = (warp - 2) / WPG;         // worker group
    const int wi = (warp - 2) - g * WPG;    // warp inside the group
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int qo = wi >> 2;                 // which sample of the row pair this warp handles in the epilogues
    const int wt = wi * 32 + lane;          // thread index inside the group
    const int row = quarter * 32 + lane;    // TMEM lane = GEMM row
    unsigned char* stg = sStg + g * stg_bytes;
    unsigned char* xop = sXop + g * xop_bytes;
    unsigned char* xt = sXt + g * xt_bytes;
    const uint32_t t_acc1 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 4u * NG + (uint32_t)(qo * C);
    const uint32_t t_acc2 = t_acc1 + 2u * NG;
    // phase of this thread's conv1 output row, decimated index and xt sample of its qo half (fixed per thread)
    int f = 0;
    for (int i = 1; i < d; ++i)
      if (row >= p.base[i]) f = i;
    const int mo = 2 * (row - p.base[f]) + qo;      // decimated output index inside phase f
    const bool has_xt = mo < p.n_out[f];            // this (row, qo) is an xt sample of the tile (else: halo / pad row)
    const int u_xt = d * mo + f;                    // its position in the xt tile [0, U)
    const unsigned lx_magic = 0xFFFFFFFFu / (unsigned)Lx + 1u;   // ceil(2^32 / Lx)
    // convert: fp32 staging tile -> lrelu -> fp16 hi/lo operand tile in (phase, decimated pair) order
    auto convert = [&](int tile, uint32_t ph) {
      const int b = tile / p.tiles_per_b;
      const int t0 = (tile - b * p.tiles_per_b) * M_out;
      const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
      mbar_wait(&stg_full[g], ph);
      mbar_wait(&xop_empty[g], ph ^ 1);
      const int tx0 = t0 - p2 - p1;
      for (int item = wt; item < C8 * Lx; item += WPG * 32) {
        const int c8 = (int)__umulhi((unsigned)item, lx_magic), v = item - c8 * Lx;   // item / Lx (exact: item < 2^16)
        const int m = d == 1 ? v : (int)__umulhi((unsigned)v, p.d_magic), fv = v - m * d;   // decimated index, phase
        const int orow = p.base[fv] + (m >> 1);
        const size_t o = ((size_t)((m & 1) * C8 + c8) * RX + orow) * 16;
        const int t = tx0 + v;
        if (t >= 0 && t < Tvalid) {
          const float4 a = *reinterpret_cast<const float4*>(stg + (size_t)item * 32);
          const float4 c = *reinterpret_cast<const float4*>(stg + (size_t)item * 32 + 16);
          float w[8];
          const float si = p.in_scale;   // leaky(x) * s == leaky(x * s) for s > 0
          w[0] = leaky(a.x * si, 0.1f); w[1] = leaky(a.y * si, 0.1f); w[2] = leaky(a.z * si, 0.1f); w[3] = leaky(a.w * si, 0.1f);
          w[4] = leaky(c.x * si, 0.1f); w[5] = leaky(c.y * si, 0.1f); w[6] = leaky(c.z * si, 0.1f); w[7] = leaky(c.w * si, 0.1f);
          split_store8(reinterpret_cast<__half*>(xop + o), reinterpret_cast<__half*>(xop + xop_plane + o), w);
        } else {
          *reinterpret_cast<uint4*>(xop + o) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(xop + xop_plane + o) = make_uint4(0, 0, 0, 0);
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
    const int first = blockIdx.x + g * gridDim.x, step = G * gridDim.x;
    if (first < p.n_tiles) convert(first, 0);
    for (int tile = first; tile < p.n_tiles; tile += step, ++it) {
      const uint32_t ph = it & 1;
      const int b = tile / p.tiles_per_b;
      const int t0 = (tile - b * p.tiles_per_b) * M_out;
      const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
      // ---- residual (and MRF accumulator) prefetch for epilogue 2: sample t0 + 2*row + qo of the same fp32 tensor (L2-hot)
      const int w_out = 2 * row + qo;
      const int t_out = t0 + w_out;
      const bool out_valid = w_out < M_out && t_out < Tvalid;
      float4 rq[CH * 2], aq[CH * 2];
      if (out_valid) {
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          const size_t fi = (((size_t)b * C8 + q) * p.Tpf + p.f_halo + t_out) * 8;
          ldg8(p.x + fi, rq[2 * q], rq[2 * q + 1]);
          if (p.acc_in) ldg8(p.acc_in + fi, aq[2 * q], aq[2 * q + 1]);
        }
      }
      // ---- epilogue 1: acc1 -> xt tile (natural sample order, two samples per row)
      mbar_wait(&acc1_full[g], ph);
      tc_fence_after();
      {
        const int t = t0 - p2 + u_xt;
        const bool v_ok = has_xt && t >= 0 && t < Tvalid;
        float m[CH][8], x8[CH][8];
#pragma unroll
        for (int q = 0; q < CH; ++q) {   // all TMEM loads in flight, one wait
          tmem_ld8(t_acc1 + q * 8, m[q]);
          tmem_ld8(t_acc1 + NG + q * 8, x8[q]);
        }
        tmem_ld_wait();
        if (has_xt) {
#pragma unroll
          for (int q = 0; q < CH; ++q) {
            const size_t o = ((size_t)((u_xt & 1) * C8 + q) * RX + (u_xt >> 1)) * 16;
            if (v_ok) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = leaky((m[q][e] + x8[q][e]) * p.inv1 + s_b1[q * 8 + e], 0.1f) * p.xt_scale;
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
      // ---- operand tile of this group's NEXT tile, so the MMA thread never waits for it
      if (tile + step < p.n_tiles) convert(tile + step, ph ^ 1);
      // ---- epilogue 2: acc2 + bias + residual [+ xs] [/ n] -> outputs
      mbar_wait(&acc2_full[g], ph);
      tc_fence_after();
      float m2[CH][8], y2[CH][8];
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        tmem_ld8(t_acc2 + q * 8, m2[q]);
        tmem_ld8(t_acc2 + NG + q * 8, y2[q]);
      }
      tmem_ld_wait();
      // the accumulator is in registers: release it before the (long) store phase
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_empty[g]);
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        if (!out_valid && !(p.out_hi && w_out < M_out && t_out < p.T)) continue;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (m2[q][e] + y2[q][e]) * p.inv2 + s_b2[q * 8 + e];
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
            stg8(p.out_f + (((size_t)b * C8 + q) * p.Tpf + p.f_halo + t_out) * 8, v);
          }
          if (p.out_plain) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              p.out_plain[((size_t)b * C + q * 8 + e) * p.T + t_out] = p.plain_act ? leaky(v[e], p.plain_slope) : v[e];
          }
        }
        if (p.out_hi) {
          const size_t o = (((size_t)b * C8 + q) * p.Tp + p.p_halo + t_out) * 8;
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G <= 2 ? 256 : 512));
  }
}

}  // namespace dissc

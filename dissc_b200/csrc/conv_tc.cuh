// Tensor-core implicit-GEMM convolution for the vocoder: tcgen05.mma (kind::f16) on split-fp16
// operands so the result keeps fp32 accuracy.  One kernel serves
//   * the dilated Conv1d of the ResBlocks (sr/models.py:36-40) with the residual / MRF epilogues,
//   * conv_pre (:99) on the gathered embedding planes, and
//   * the polyphase ConvTranspose1d upsamplers (:102): every output phase is a small dense conv
//     over the INPUT frames, the phases are laid side by side along the GEMM N dimension.
//
//   D[r, n] = sum_j sum_ci A[r + j*dil - pad, ci] * W_j[ci, n]          r: 128 rows (time / frames) per tile
//
// Split precision: x = x_hi + x_lo (both fp16, weights pre-scaled by 2^s so w_lo stays normal) and
//   x*w ~= x_hi*w_hi + (x_lo*w_hi + x_hi*w_lo)                          (error ~2^-22 relative)
// N <= 128: TWO MMAs per 16-channel k-step:  A_hi x [W_hi | W_lo] -> [main | cross],  A_lo x W_hi -> cross
//           (the weight planes sit side by side in shared memory so one N=2*NC MMA reads A once);
// N == 256: three MMAs (hi*hi -> main, lo*hi -> cross, hi*lo -> cross), or all into one accumulator
//           (single_acc) which frees half of TMEM for double buffering.
// The cross terms get their own accumulator because the tensor core truncates the fp32 accumulator
// after every MMA; the epilogue adds main + cross with a proper round-to-nearest.
//
// HBM layouts (channel-blocked, time-major):
//   planes hi/lo : fp16 [B][C/8][Tp][8]   Tp = roundup(T,128) + 2*kTcHalo, kTcHalo zero rows either side
//   f32b         : fp32 [B][C/8][Tr][8]   Tr = roundup(T,128)            (residual / MRF accumulator)
// One (c8, row range) slab is contiguous, so an activation block is a handful of 1-D bulk TMA copies
// and lands in shared memory as [c8][row][8] = the canonical no-swizzle K-major UMMA operand layout
// (core matrix 8 rows x 16 B, SBO 128 B, LBO rows*16 B).  A tap shift of j*dil rows is +j*dil*16 B on
// the descriptor start address: one resident block serves all taps.  The conv zero padding is the halo.
// Weights are packed [chunk][cb][tap][c8][hi|lo][NC][8]: a pipeline stage is one contiguous bulk copy.
//
// CTA = 6 warps, persistent over work items (tile x N-chunk): warp 0 = producer (bulk TMA), warp 1 =
// MMA issuer (one thread) + TMEM owner, warps 2-5 = epilogue (TMEM -> registers -> bias / residual /
// MRF / leaky-relu -> fp16 split -> global).  TMEM accumulators ping-pong so the epilogue of item i
// overlaps the MMAs of item i+1.  All slot / phase bookkeeping is incremental (no div/mod in the loops).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace dissc {

// CTA = producer warp + MMA warp + EPW epilogue warps.  N >= 128 (and N = 64 when one CTA fills the SM) uses 8 epilogue
// warps, two per TMEM lane quarter, each taking half of the columns: a 128 x N tile is too much epilogue for 4 warps to
// hide behind the MMAs of the next tile.
constexpr int kTcHalo = 32;       // zero rows either side of every plane row-slab (>= max conv padding 25)
constexpr int kTcMaxStages = 64;  // barrier slots for the weight pipeline (resident mode: one per stage)

struct TcParams {
  const __half* a_hi;  // planes [B][Cin8][Tp_in][8]
  const __half* a_lo;
  const __half* w;     // packed [chunk][cb][tap][KB/8][2][NC][8]
  const float* bias;   // [Cout]
  float w_inv_scale;   // 2^-s
  const float* res;     // f32b [B][Cout/8][Tr][8] or null
  const float* acc_in;  // f32b or null
  float* out_f32b;      // f32b or null (raw value)
  __half* out_hi;       // planes [B][Cout/8][Tp][8] (leaky-relu(plane_slope) applied iff plane_act) or null
  __half* out_lo;
  float* out_plain;     // (B, Cout, T) fp32 (leaky-relu(plain_slope) iff plain_act) or null
  const int* lengths;
  int len_mul;          // valid OUTPUT rows of utterance b = lengths[b] * len_mul
  int B, Cin8, Cout, T, Tp, Tr, Tp_in;
  int k, dil, pad;      // taps, dilation, left padding in rows (transposed conv: k = taps per phase, pad = k-1)
  int KB;               // channels per activation block (16 or 32)
  int n_cb;             // Cin_padded / KB
  int JG;               // taps per weight stage
  int SPC;              // weight stages per channel block = ceil(k / JG)
  int NS;               // weight slots in shared memory
  int resident;         // 1: NS == n_cb*SPC and n_chunks == 1, weights loaded once per CTA
  int n_chunks;         // N chunks of NC columns
  int tiles_per_b, n_items;
  int tmem_cols;        // allocation (power of two)
  int acc_cols;         // TMEM columns per accumulator buffer (2*NC, or NC with single_acc)
  int NA;               // activation-block buffers in shared memory (2..4)
  int nbuf;             // TMEM accumulator buffers (1 or 2)
  int single_acc;       // NC == 256 only: all three MMAs into one accumulator
  int up;               // 0: conv.  u > 0: transposed conv with stride u; tile rows are input frames,
  int up_P;             //   a chunk holds up_P phases x Cout channels, output row = frame*u + phase - up_pad
  int up_pad;
  float div;
  int plane_act, plain_act;  // 0 none, 1 leaky-relu(slope), 2 GELU (erf)
  float plane_slope, plain_slope;
  int halo;              // zero rows either side of every plane slab (kTcHalo for the vocoder)
  int pre_act;           // activation applied right after the bias, BEFORE residual / accumulate (0 none, 2 GELU)
  int out_deint;         // planes output de-interleaved for a following stride-2 conv: row t -> slab phase t&1, row t>>1
  int f_halo;            // f32b tensors (res / acc_in / out_f32b): rows of slack in front of every slab (0 = plain f32b;
                         //   > 0 = the "f32h" layout of resblock_tc.cuh, with Tr = rows per slab incl. slack)
  int groups;            // grouped conv: chunk g reads channel blocks [g*n_cb, (g+1)*n_cb) and writes group_c8 8-channel
  int group_c8;          //   groups of output channels starting at g*group_c8 (NC >= 8*group_c8, padded columns dropped)
  int cout_log2;         // log2(Cout) when Cout is a power of two, else -1 (kTcUp needs it)
  int split_w;           // EPW == 8 kernels: 1 = weights have their own producer thread (warp 2), 0 = warp 0 loads both
  int cb_split, k_hi;    // channel blocks >= cb_split carry only their first k_hi taps (the others are structural zeros:
                         //   the odd phase of a stride-2 conv in frame form); cb_split == 0: every block has k taps
  // Activation scale (powers of two, chosen per tensor at load time so the fp16 hi / lo planes sit in the middle of the
  // fp16 range whatever the magnitude of the activations; exact in fp32, so results do not depend on it unless a plane
  // would otherwise underflow / saturate).  The INPUT planes hold value * in_scale: the host folds 1 / in_scale into
  // w_inv_scale.  The OUTPUT planes are written as act(value) * plane_scale.  0 means 1.
  float in_scale, plane_scale;
  // 2-CTA clusters sharing ONE weight stream (EPW == 8, streamed weights): the two CTAs of a cluster work on two
  // different tiles of the same N chunk in lockstep; each loads half of every weight stage and multicasts it into both
  // CTAs' shared memory, which halves the L2 -> SM weight traffic (43 B/clk/SM without it, the chip's L2 limit).
  int cluster2;
  int pair2;            // CTA pair issuing ONE 256-row tcgen05.mma.cta_group::2 per k-step: each CTA stages its own 128
                        // activation rows and HALF the weight columns (see the kernel)
  int n_tiles;           // B * tiles_per_b
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float tc_act(float v, int mode, float slope) {
  return mode == 1 ? leaky(v, slope) : (mode == 2 ? gelu_erf(v) : v);
}

// ---- tcgen05 wrappers ------------------------------------------------------------------------
// K-major, no swizzle: start address >> 4 | LBO >> 4 (K-chunk stride) << 16 ; high word: SBO (=128 B, stride between
// 8-row groups) >> 4 | descriptor version 1 << 14.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
constexpr uint32_t kUmmaDescHi = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint64_t umma_desc(uint32_t lo) { return ((uint64_t)kUmmaDescHi << 32) | lo; }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// commit that arrives on the mbarrier at the same offset in every CTA of `cta_mask` (a weight slot shared by a CTA pair
// is free once BOTH CTAs' MMAs have read it)
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- CTA-pair (cta_group::2) forms: issued by the leader CTA for both ----
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 pair -> packed fp16x2, round to nearest, saturating to +-65504 (one F2FP.SATFINITE instruction)
__device__ __forceinline__ uint32_t pack_half2_sat(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 8 values -> one 16-byte store per plane
__device__ __forceinline__ void split_store8(__half* hi_dst, __half* lo_dst, const float v[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_half2_sat(v[2 * i], v[2 * i + 1]);
    const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
    l[i] = pack_half2_sat(v[2 * i] - back.x, v[2 * i + 1] - back.y);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

// MODE selects how much of the epilogue is compiled in (the runtime switches of TcParams cost instructions in the
// per-element loops, and the narrow layers are bound by exactly that):
//   kTcGeneric : everything (HuBERT: GELU, grouped / de-interleaved outputs, plain outputs, ...)
//   kTcConv    : vocoder Conv1d -- bias [+ residual] [+ MRF accumulate] [/ n] -> f32b and/or leaky-relu planes
//   kTcUp      : vocoder polyphase ConvTranspose1d -- bias -> f32b and/or leaky-relu planes (Cout a power of two)
enum { kTcGeneric = 0, kTcConv = 1, kTcUp = 2 };

// 32-byte global accesses (one f32b row of 8 channels): LDG.E.256 / STG.E.256 on sm_100
__device__ __forceinline__ void ldg8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg8(float* p, const float v[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// one f32b row (8 floats): a single 32-byte access where the caller guarantees 32-byte alignment (WIDE), else two float4
template <bool WIDE>
__device__ __forceinline__ void ld_row8(const float* p, float4& a, float4& b) {
  if constexpr (WIDE) {
    ldg8(p, a, b);
  } else {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
}
template <bool WIDE>
__device__ __forceinline__ void st_row8(float* p, const float v[8]) {
  if constexpr (WIDE) {
    stg8(p, v);
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// Warp roles: warp 0 = activation producer, warp 1 = MMA issuer + TMEM owner, [warp 2 = weight producer], then EPW epilogue
// warps.  The one-CTA-per-SM variants (EPW == 8: the layers that STREAM their weights) give the weights their own producer
// thread: a single in-order producer can only request the next activation block after it has queued all weight stages of
// the current one, i.e. (ring depth) stages before the block is needed -- ncu showed the MMA thread re-trying 44 % of its
// activation waits in the stage-1 k = 11 layers.  With two threads the block is requested the moment its buffer frees.
__host__ __device__ constexpr int tc_pre_warps(int epw) { return epw == 8 ? 3 : 2; }
__host__ __device__ constexpr int tc_threads(int epw) { return (tc_pre_warps(epw) + epw) * 32; }

// PAIR: the CTA-pair form (p.pair2) is its own instantiation -- a kernel that merely CONTAINS cta_group::2 instructions can
// only be launched as a cluster of two.
template <int NC, int EPW, int MODE, bool PAIR = false>
__global__ void __launch_bounds__(tc_threads(EPW), (EPW == 8) ? 1 : ((NC <= 32) ? 3 : 2)) conv_tc_kernel(const TcParams p) {
  constexpr int PW = tc_pre_warps(EPW);   // warps in front of the epilogue warps
  constexpr int kTcThreads = tc_threads(EPW);
  constexpr bool kTwoMma = (NC <= 128);
  constexpr bool kFast = (MODE != kTcGeneric);
  // switches the specialised modes fold away at compile time
  const int up = (MODE == kTcConv) ? 0 : p.up;
  const int groups = kFast ? 0 : p.groups;
  const int pre_act = kFast ? 0 : p.pre_act;
  const int out_deint = kFast ? 0 : p.out_deint;
  float* const out_plain = kFast ? nullptr : p.out_plain;
  const float* const res_p = (MODE == kTcUp) ? nullptr : p.res;
  const float* const acc_p = (MODE == kTcUp) ? nullptr : p.acc_in;
  const float div = p.div;  // unused by kTcUp
  const float plane_scale = p.plane_scale;
  const int single_acc = (NC == 256) ? p.single_acc : 0;
  constexpr int G = NC / 8;            // 8-channel groups per chunk
  constexpr int EB = 2;                // groups per epilogue batch
  constexpr int NB = G / EB;
  constexpr bool kAccPrefetch = (NC >= 64);  // MRF accumulator prefetched like the residual (registers permitting)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int R = 128 + (p.k - 1) * p.dil;                          // activation rows a tile needs
  const uint32_t lbo_a = (uint32_t)R * 16;                        // bytes between 8-channel chunks of A
  const uint32_t a_plane_bytes = (uint32_t)(p.KB / 8) * lbo_a;    // one plane of one block
  const uint32_t a_bytes = 2 * a_plane_bytes;
  // CTA pair (tcgen05.mma.cta_group::2, M = 256): the pair walks (chunk, tile pair) items like the multicast clusters; each
  // CTA stages the activation rows of ITS tile and the weight columns [rank * NC/2, rank * NC/2 + NC/2) of every stage --
  // half the weight bytes per SM, which is what bounds the streamed-weight layers (shared-memory fill rate) -- and the
  // leader (rank 0) issues every MMA for both.  The other CTA's MMA warp relays its "operands landed" barriers to the
  // leader (a non-tensor cp.async.bulk cannot complete on another CTA's mbarrier: tried, the signal never arrives); slots
  // are released in both CTAs by multicast commits.
  constexpr bool pr2 = PAIR && (PW == 3);
  const uint32_t wcols = pr2 ? (uint32_t)NC / 2 : (uint32_t)NC;   // weight columns staged by this CTA
  const uint32_t lbo_b = 2u * wcols * 16;                         // [c8][hi|lo][wcols][8]
  const uint32_t w_tap_bytes = (uint32_t)(p.KB / 8) * lbo_b;      // per CTA; the global stage holds both halves
  const uint32_t w_tap_gbytes = pr2 ? 2 * w_tap_bytes : w_tap_bytes;
  const uint32_t w_slot_bytes = (uint32_t)p.JG * w_tap_bytes;
  unsigned char* sA = smem_raw;
  unsigned char* sW = sA + (size_t)p.NA * a_bytes;
  float* s_bias = reinterpret_cast<float*>(sW + (size_t)p.NS * w_slot_bytes);  // [n_chunks*NC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + ((p.n_chunks * NC + 1) & ~1));
  uint64_t* a_full = bars;            // [4]
  uint64_t* a_empty = bars + 4;       // [4]
  uint64_t* acc_full = bars + 8;      // [2]
  uint64_t* acc_empty = bars + 10;    // [2]
  uint64_t* w_full = bars + 12;       // [NS]
  uint64_t* w_empty = w_full + p.NS;  // [NS]
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // cluster mode: the pair (2i, 2i+1) walks "pair items" (chunk, tile pair) together; member r takes tile 2*pair + r.  An
  // odd tile count makes the last pair compute its last tile twice (identical stores) so both keep the same weight sequence.
  const bool cl2 = (PW == 3) && (p.cluster2 || pr2);   // item mapping / cluster syncs shared by both pair modes
  const bool mc2 = cl2 && !pr2;                            // multicast weight stream (two independent M = 128 MMAs)
  const uint32_t crank = cl2 ? cluster_ctarank() : 0;
  const bool leader = !pr2 || crank == 0;
  const int it_first = cl2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int it_step = cl2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int it_count = cl2 ? ((p.n_tiles + 1) >> 1) * p.n_chunks : p.n_items;
  auto item_of = [&](int it, int& chunk, int& tile) {
    const int q = it / p.n_chunks;
    chunk = it - q * p.n_chunks;
    tile = cl2 ? min(2 * q + (int)crank, p.n_tiles - 1) : q;
  };

  if (tid == 0) {
    // pair mode, leader: "full" = own producer (+ bytes) and the partner's relay; "accumulator drained" = both epilogues
    const uint32_t full_n = (pr2 && crank == 0) ? 2 : 1;
    for (int i = 0; i < 4; ++i) {
      mbar_init(&a_full[i], full_n);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], (pr2 && crank == 0) ? 2 * EPW : EPW);
    }
    for (int i = 0; i < p.NS; ++i) {
      mbar_init(&w_full[i], full_n);
      mbar_init(&w_empty[i], mc2 ? 2 : 1);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < p.n_chunks * NC; i += kTcThreads) {
    int co = up ? (i % p.Cout) : i;
    if (groups) {
      const int ch = i / NC, n = i - ch * NC;
      co = n < p.group_c8 * 8 ? ch * p.group_c8 * 8 + n : p.Cout;
    }
    s_bias[i] = (p.bias && co < p.Cout) ? p.bias[co] : 0.f;
  }
  if (warp == 1) {
    if (pr2) {   // both CTAs of the pair, same warp, same shared-memory word
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                   "r"(p.tmem_cols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                   "r"(p.tmem_cols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  if (cl2) cluster_sync_all();   // both CTAs' mbarriers exist before either multicasts into / arrives on the other's
  pdl_wait();                 // the prologue above touched only weights / shared memory; activations from here on
  pdl_launch_dependents();

  if (warp == 0 || (PW == 3 && warp == 2)) {
    // ===================== producers: bulk TMA =====================
    // PW == 2: warp 0 loads activation blocks AND weight stages (in MMA order).  PW == 3: warp 0 activations, warp 2 weights.
    const bool split = (PW == 3) && p.split_w;
    const bool do_a = (warp == 0), do_w = split ? (warp == 2) : (warp == 0);
    if (elect_one()) {
      uint32_t abuf = 0, aph = 0, ws = 0, wph = 0;
      bool first = true;
      const int kb8 = p.KB / 8;
      for (int item = it_first; item < it_count; item += it_step) {
        int chunk, tile;
        item_of(item, chunk, tile);
        const int b = tile / p.tiles_per_b;
        const int r0 = (tile - b * p.tiles_per_b) * 128;
        const size_t in0 = (((size_t)b * p.Cin8 + (groups ? (size_t)chunk * p.n_cb * kb8 : 0)) * p.Tp_in + p.halo + r0 -
                            p.pad) * 8;
        const __half* hi0 = p.a_hi + in0;
        const __half* lo0 = p.a_lo + in0;
        const unsigned char* wchunk = reinterpret_cast<const unsigned char*>(p.w) +
                                      (size_t)chunk * p.n_cb * p.k * w_tap_gbytes;
        for (int cb = 0; cb < p.n_cb; ++cb) {
          if (do_a) {
            mbar_wait(&a_empty[abuf], aph ^ 1);
            mbar_arrive_expect_tx(&a_full[abuf], a_bytes);
            unsigned char* dst = sA + abuf * a_bytes;
            for (int c8 = 0; c8 < kb8; ++c8) {
              const size_t off = (size_t)(cb * kb8 + c8) * p.Tp_in * 8;
              tma_load_1d(dst + (size_t)c8 * lbo_a, hi0 + off, lbo_a, &a_full[abuf]);
              tma_load_1d(dst + a_plane_bytes + (size_t)c8 * lbo_a, lo0 + off, lbo_a, &a_full[abuf]);
            }
            if (++abuf == (uint32_t)p.NA) { abuf = 0; aph ^= 1; }
          }
          if (do_w && (!p.resident || first)) {
            int j0 = 0;
            for (int g = 0; g < p.SPC; ++g, j0 += p.JG) {
              const int nt = min(p.JG, p.k - j0);
              // a stage made only of structurally zero taps (odd-kernel stride-2 frame form, second half of the channel
              // blocks) is neither loaded nor waited for: the MMA thread skips it the same way
              if (!p.resident && p.cb_split && cb >= p.cb_split && j0 >= p.k_hi) continue;
              const uint32_t slot = p.resident ? (uint32_t)(cb * p.SPC + g) : ws;
              if (!p.resident) mbar_wait(&w_empty[slot], wph ^ 1);
              const uint32_t wbytes = (uint32_t)nt * w_tap_bytes;
              mbar_arrive_expect_tx(&w_full[slot], wbytes);
              const unsigned char* wsrc = wchunk + ((size_t)cb * p.k + j0) * w_tap_gbytes;
              if (pr2) {
                // this CTA's column half of every tap of the stage ([tap][rank][...] in global memory)
                for (int jj = 0; jj < nt; ++jj)
                  tma_load_1d(sW + (size_t)slot * w_slot_bytes + (size_t)jj * w_tap_bytes,
                              wsrc + (size_t)jj * w_tap_gbytes + crank * w_tap_bytes, w_tap_bytes, &w_full[slot]);
              } else if (mc2) {
                // this CTA's half of the stage, into BOTH CTAs' slot; the partner delivers the other half
                const uint32_t half = wbytes >> 1;
                tma_load_1d_multicast(sW + (size_t)slot * w_slot_bytes + crank * half, wsrc + crank * half, half,
                                      &w_full[slot], (uint16_t)3);
              } else {
                tma_load_1d(sW + (size_t)slot * w_slot_bytes, wsrc, wbytes, &w_full[slot]);
              }
              if (++ws == (uint32_t)p.NS) { ws = 0; wph ^= 1; }
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1 && !leader) {
    // ===================== pair mode, second CTA: relay "operands landed" to the leader's barriers =====================
    // Same walk as the MMA thread below; an arrival on the leader's a_full / w_full says this CTA's share of the block /
    // stage is in ITS shared memory (the pair MMA reads both).
    if (elect_one()) {
      uint32_t abuf = 0, aph = 0, ws = 0, wph = 0;
      const uint32_t a_full0 = mapa_u32(smem_u32(a_full), 0), w_full0 = mapa_u32(smem_u32(w_full), 0);
      for (int item = it_first; item < it_count; item += it_step) {
        for (int cb = 0; cb < p.n_cb; ++cb) {
          mbar_wait(&a_full[abuf], aph);
          mbar_arrive_cluster(a_full0 + abuf * 8);
          if (++abuf == (uint32_t)p.NA) { abuf = 0; aph ^= 1; }
          int j0 = 0;
          const int kc = (p.cb_split && cb >= p.cb_split) ? p.k_hi : p.k;
          for (int g = 0; g < p.SPC; ++g, j0 += p.JG) {
            if (min(p.JG, kc - j0) <= 0) continue;
            mbar_wait(&w_full[ws], wph);
            mbar_arrive_cluster(w_full0 + ws * 8);
            if (++ws == (uint32_t)p.NS) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (elect_one()) {
      // instruction descriptor: D=f32 (1<<4), A=B=f16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
      constexpr uint32_t idesc_n = (1u << 4) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * NC) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_pair = (1u << 4) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // M = 256
      const uint32_t sA0 = smem_u32(sA), sW0 = smem_u32(sW);
      const int KS = p.KB / 16;
      const uint32_t a_kstep = (2 * lbo_a) >> 4, b_kstep = (2 * lbo_b) >> 4;  // descriptor units (16 B)
      const uint32_t a_lo_off = a_plane_bytes >> 4;
      uint32_t abuf = 0, aph = 0, ws = 0, wph = 0, ab = 0, accph = 0;
      bool first = true;
      for (int item = it_first; item < it_count; item += it_step) {
        if (pr2) mbar_wait_cluster(&acc_empty[ab], accph ^ 1); else mbar_wait(&acc_empty[ab], accph ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + ab * (uint32_t)p.acc_cols;
        const uint32_t d_cross = single_acc ? d_main : d_main + NC;
        uint32_t accum = 0;
        for (int cb = 0; cb < p.n_cb; ++cb) {
          if (pr2) mbar_wait_cluster(&a_full[abuf], aph); else mbar_wait(&a_full[abuf], aph);
          const uint32_t a_desc0 = umma_desc_lo(sA0 + abuf * a_bytes, lbo_a);
          int j0 = 0;
          const int kc = (p.cb_split && cb >= p.cb_split) ? p.k_hi : p.k;   // taps with non-zero weights in this block
          for (int g = 0; g < p.SPC; ++g, j0 += p.JG) {
            const int nt = min(p.JG, kc - j0);   // <= 0: resident weights: the stage is only waited for; streamed: skipped
            if (!p.resident && nt <= 0) continue;   // (the weight producer did not load it either)
            const uint32_t slot = p.resident ? (uint32_t)(cb * p.SPC + g) : ws;
            if (pr2)
              mbar_wait_cluster(&w_full[slot], wph);
            else if (!p.resident)
              mbar_wait(&w_full[slot], wph);
            else if (first)
              mbar_wait(&w_full[slot], 0);
            tc_fence_after();
            uint32_t w_desc = umma_desc_lo(sW0 + slot * w_slot_bytes, lbo_b);
            uint32_t a_desc = a_desc0 + (uint32_t)(j0 * p.dil);
            for (int jj = 0; jj < nt; ++jj, a_desc += (uint32_t)p.dil, w_desc += (w_tap_bytes >> 4)) {
              uint32_t ad = a_desc, wd = w_desc;
              for (int ks = 0; ks < KS; ++ks, ad += a_kstep, wd += b_kstep) {
                if (pr2) {
                  // three M = 256 MMAs; this CTA's weight columns sit at the same offsets in the partner's stage
                  umma2_f16(d_main, umma_desc(ad), umma_desc(wd), idesc_pair, accum);
                  umma2_f16(d_cross, umma_desc(ad + a_lo_off), umma_desc(wd), idesc_pair, single_acc ? 1u : accum);
                  umma2_f16(d_cross, umma_desc(ad), umma_desc(wd + NC / 2), idesc_pair, 1);
                } else if constexpr (kTwoMma) {
                  umma_f16(d_main, umma_desc(ad), umma_desc(wd), idesc_2n, accum);            // [main | cross]
                  umma_f16(d_cross, umma_desc(ad + a_lo_off), umma_desc(wd), idesc_n, 1);     // cross += lo*hi
                } else {
                  umma_f16(d_main, umma_desc(ad), umma_desc(wd), idesc_n, accum);
                  umma_f16(d_cross, umma_desc(ad + a_lo_off), umma_desc(wd), idesc_n, single_acc ? 1u : accum);
                  umma_f16(d_cross, umma_desc(ad), umma_desc(wd + NC), idesc_n, 1);
                }
                accum = 1;
              }
            }
            if (!p.resident) {
              if (pr2)
                umma2_commit_multicast(&w_empty[slot], (uint16_t)3);  // the pair MMA has read both CTAs' halves
              else if (mc2)
                umma_commit_multicast(&w_empty[slot], (uint16_t)3);   // frees the slot in both CTAs of the pair
              else
                umma_commit(&w_empty[slot]);
            }
            if (++ws == (uint32_t)p.NS) { ws = 0; wph ^= 1; }
          }
          if (pr2) umma2_commit_multicast(&a_empty[abuf], (uint16_t)3); else umma_commit(&a_empty[abuf]);
          if (++abuf == (uint32_t)p.NA) { abuf = 0; aph ^= 1; }
        }
        if (pr2) umma2_commit_multicast(&acc_full[ab], (uint16_t)3); else umma_commit(&acc_full[ab]);
        if (++ab == (uint32_t)p.nbuf) { ab = 0; accph ^= 1; }
        first = false;
      }
    }
  } else {
    // ===================== epilogue: 4 warps, TMEM lane quarter = warp % 4 =====================
    const int quarter = warp & 3;
    constexpr int NBH = NB / (EPW / 4);                 // epilogue batches per warp
    const int bi0 = ((warp - PW) >> 2) * NBH;           // this warp's first batch (column half)
    const int row = quarter * 32 + lane;
    const int cout8 = p.Cout / 8;
    uint32_t ab = 0, accph = 0;
    for (int item = it_first; item < it_count; item += it_step) {
      int chunk, tile;
      item_of(item, chunk, tile);
      const int b = tile / p.tiles_per_b;
      const int r = (tile - b * p.tiles_per_b) * 128 + row;
      const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
      // residual of the first batch: issued before the accumulator wait so its latency hides behind the MMAs
      float4 rq[EB * 2];
      const bool conv_valid = (!up) && r < Tvalid;
      const size_t fbase = (((size_t)b * cout8 + (size_t)chunk * (groups ? p.group_c8 : G)) * p.Tr + p.f_halo + r) * 8;
      const size_t fstride = (size_t)p.Tr * 8;
      float4 aq[EB * 2];
      if (res_p && conv_valid) {
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          ld_row8<kFast>(res_p + fbase + (bi0 * EB + e) * fstride, rq[2 * e], rq[2 * e + 1]);
        }
      }
      if (kAccPrefetch && acc_p && conv_valid) {
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          ld_row8<kFast>(acc_p + fbase + (bi0 * EB + e) * fstride, aq[2 * e], aq[2 * e + 1]);
        }
      }
      mbar_wait(&acc_full[ab], accph);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * (uint32_t)p.acc_cols;
#pragma unroll 1
      for (int bi = bi0; bi < bi0 + NBH; ++bi) {
        float m[EB][8], x[EB][8];
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          tmem_ld8(taddr0 + (bi * EB + e) * 8, m[e]);
          if (!single_acc) tmem_ld8(taddr0 + NC + (bi * EB + e) * 8, x[e]);
        }
        float4 rn[EB * 2];
        if (NBH > 1 && bi + 1 < bi0 + NBH && res_p && conv_valid) {
          // prefetch the next batch's residual while this one is processed
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            ld_row8<kFast>(res_p + fbase + ((bi + 1) * EB + e) * fstride, rn[2 * e], rn[2 * e + 1]);
          }
        }
        float4 an[EB * 2];
        if (kAccPrefetch && NBH > 1 && bi + 1 < bi0 + NBH && acc_p && conv_valid) {
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            ld_row8<kFast>(acc_p + fbase + ((bi + 1) * EB + e) * fstride, an[2 * e], an[2 * e + 1]);
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          const int g8 = bi * EB + e;      // 8-column group inside the chunk
          const int n0 = chunk * NC + g8 * 8;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float s = single_acc ? m[e][i] : (m[e][i] + x[e][i]);
            v[i] = s * p.w_inv_scale + s_bias[n0 + i];
          }
          int t, c8o;
          bool inb, valid;
          if (up) {
            const int phase = (MODE == kTcUp) ? (n0 >> p.cout_log2) : (n0 / p.Cout);
            c8o = (n0 - phase * p.Cout) >> 3;
            t = r * up + phase - p.up_pad;
            inb = t >= 0 && t < p.T;
            valid = inb && t < Tvalid;
          } else {
            c8o = groups ? chunk * p.group_c8 + g8 : (n0 >> 3);
            t = r;
            inb = true;
            valid = conv_valid;
            if (groups && g8 >= p.group_c8) continue;  // padded output columns of a group
          }
          if (!kFast && c8o >= cout8) continue;  // padded output columns
          if (pre_act) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = tc_act(v[i], pre_act, 0.f);
          }
          const size_t fidx = (((size_t)b * cout8 + c8o) * p.Tr + p.f_halo + t) * 8;
          if (valid) {
            if (res_p) {
              v[0] += rq[2 * e].x; v[1] += rq[2 * e].y; v[2] += rq[2 * e].z; v[3] += rq[2 * e].w;
              v[4] += rq[2 * e + 1].x; v[5] += rq[2 * e + 1].y; v[6] += rq[2 * e + 1].z; v[7] += rq[2 * e + 1].w;
            }
            if (acc_p) {
              float4 r0, r1;
              if constexpr (kAccPrefetch) {
                r0 = aq[2 * e]; r1 = aq[2 * e + 1];
              } else {
                ld_row8<kFast>(acc_p + fidx, r0, r1);
              }
              v[0] = r0.x + v[0]; v[1] = r0.y + v[1]; v[2] = r0.z + v[2]; v[3] = r0.w + v[3];
              v[4] = r1.x + v[4]; v[5] = r1.y + v[5]; v[6] = r1.z + v[6]; v[7] = r1.w + v[7];
            }
            if constexpr (MODE != kTcUp) {
              if (div != 0.f) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = v[i] / div;
              }
            }
            if (p.out_f32b) {
              st_row8<kFast>(p.out_f32b + fidx, v);
            }
            if (out_plain) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                out_plain[((size_t)b * p.Cout + c8o * 8 + i) * p.T + t] = tc_act(v[i], p.plain_act, p.plain_slope);
            }
          }
          if (p.out_hi && inb) {
            // rows >= valid length are written as zeros: they are the next conv's zero padding
            const size_t pidx = out_deint
                                    ? (((size_t)b * 2 * cout8 + (size_t)(t & 1) * cout8 + c8o) * p.Tp + p.halo + (t >> 1)) * 8
                                    : (((size_t)b * cout8 + c8o) * p.Tp + p.halo + t) * 8;
            if (valid) {
              float a[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                a[i] = (kFast ? leaky(v[i], p.plane_slope) : tc_act(v[i], p.plane_act, p.plane_slope)) * plane_scale;
              split_store8(p.out_hi + pidx, p.out_lo + pidx, a);
            } else {
              *reinterpret_cast<uint4*>(p.out_hi + pidx) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(p.out_lo + pidx) = make_uint4(0, 0, 0, 0);
            }
          }
        }
        if (NBH > 1) {
#pragma unroll
          for (int e = 0; e < EB * 2; ++e) {
            rq[e] = rn[e];
            if constexpr (kAccPrefetch) aq[e] = an[e];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (pr2 && crank != 0)
          mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[ab]), 0));   // the leader's MMA thread waits for both epilogues
        else
          mbar_arrive(&acc_empty[ab]);
      }
      if (++ab == (uint32_t)p.nbuf) { ab = 0; accph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cl2) cluster_sync_all();   // neither CTA leaves while the other may still multicast into it / arrive on its barriers
  if (warp == 1) {
    if (pr2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

// Zero rows [0,HP) and [HP+T, Tp) of every (b, c8) slab of a plane pair: the conv zero padding
// (left halo, the round-up rows [T,Tr) and the right halo).
static __global__ void tc_zero_halos_kernel(__half* hi, __half* lo, int slabs, int Tp, int T, int halo) {
  const int per = halo + (Tp - halo - T);  // rows per slab to clear
  const long long total = (long long)slabs * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i / per);
    int r = (int)(i - (long long)s * per);
    r = r < halo ? r : (T + r);  // second range starts at row halo+T
    const size_t off = ((size_t)s * Tp + r) * 8;
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(0, 0, 0, 0);
  }
}

// Fused gather/concat of CodeGenerator.forward (sr/models.py:189,:206-215) straight into the split planes that
// conv_pre consumes: channels [0,E) = dict[code[b,t]], [E] = f0[b,t], then spkr_emb[spkr[b]] repeated over time;
// channels >= Cin and every row outside [0, valid length) are written as zeros (halos included).
struct EmbedParams {
  const long long* code;  // (B, T)
  const float* f0;        // (B, T) or null
  const long long* spkr;  // (B) or null
  const float* dict_w;    // (num_embeddings, E)
  const float* spkr_w;    // (rows, E)
  const int* lengths;
  int E, f0_ch, spk_base, Cin;  // f0_ch / spk_base = -1 if absent
  const float* extra;     // (B, n_extra) or null: per-utterance conditioning features repeated over time as channels
  int extra_base, n_extra;  //   [extra_base, extra_base + n_extra) (sr/models.py:216-221, e.g. `f0_stats`)
  int B, C8, T, Tp;
  float scale;                   // planes hold value * scale (TcParams::in_scale of conv_pre)
  int n_code_rows, n_spkr_rows;  // table rows: ids outside [0, rows) set *err (common.cuh::checked_row)
  int* err;
  __half* hi;
  __half* lo;
};
static __global__ void tc_embed_planes_kernel(const EmbedParams p) {
  const long long total = (long long)p.B * p.C8 * p.Tp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i % p.Tp);
    const long long s = i / p.Tp;
    const int c8 = (int)(s % p.C8), b = (int)(s / p.C8);
    const int t = row - kTcHalo;
    const int Tvalid = p.lengths ? min(p.T, p.lengths[b]) : p.T;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ci = c8 * 8 + e;
      float x = 0.f;
      if (t >= 0 && t < Tvalid && ci < p.Cin) {
        if (ci < p.E) {
          x = __ldg(p.dict_w + (size_t)checked_row(p.code[(size_t)b * p.T + t], p.n_code_rows, p.err, kIdxUnit) * p.E + ci);
        } else if (ci == p.f0_ch) {
          x = __ldg(p.f0 + (size_t)b * p.T + t);
        } else if (p.n_extra > 0 && ci >= p.extra_base) {
          x = __ldg(p.extra + (size_t)b * p.n_extra + (ci - p.extra_base));
        } else if (p.spk_base >= 0 && ci >= p.spk_base) {
          x = __ldg(p.spkr_w + (size_t)checked_row(p.spkr[b], p.n_spkr_rows, p.err, kIdxSpeaker) * p.E + (ci - p.spk_base));
        }
      }
      v[e] = x * p.scale;
    }
    split_store8(p.hi + (size_t)i * 8, p.lo + (size_t)i * 8, v);
  }
}

// plain (B,C,T) fp32 -> split planes [B][C8][Tp][8] (+ optional leaky-relu); rows >= valid length and channels >= C
// are zero.  Layer-test helper (the model writes planes straight from the producing kernel's epilogue).
static __global__ void tc_pack_planes_kernel(const float* in, __half* hi, __half* lo, const int* lengths, int len_mul, int B, int C,
                                      int C8, int T, int Tp, int act, float slope, float scale = 1.f) {
  const long long total = (long long)B * C8 * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long s = i / T;  // slab = b*C8 + c8
    const int b = (int)(s / C8), c8 = (int)(s % C8);
    const int Tvalid = lengths ? min(T, lengths[b] * len_mul) : T;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c8 * 8 + e;
      float x = (t < Tvalid && c < C) ? in[((size_t)b * C + c) * T + t] : 0.f;
      v[e] = (act ? leaky(x, slope) : x) * scale;
    }
    const size_t off = ((size_t)s * Tp + kTcHalo + t) * 8;
    split_store8(hi + off, lo + off, v);
  }
}

// plain (B,C,T) fp32 <-> blocked f32b [B][C/8][Tr][8] (layer-test helpers)
static __global__ void tc_plain_to_f32b_kernel(const float* in, float* out, int B, int C, int T, int Tr) {
  const long long total = (long long)B * C * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long bc = i / T;
    const int c = (int)(bc % C), b = (int)(bc / C);
    out[(((size_t)b * (C / 8) + c / 8) * Tr + t) * 8 + (c & 7)] = in[i];
  }
}
static __global__ void tc_f32b_to_plain_kernel(const float* in, float* out, int B, int C, int T, int Tr) {
  const long long total = (long long)B * C * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long bc = i / T;
    const int c = (int)(bc % C), b = (int)(bc / C);
    out[i] = in[(((size_t)b * (C / 8) + c / 8) * Tr + t) * 8 + (c & 7)];
  }
}
// planes -> plain fp32 (hi + lo), for tests
static __global__ void tc_planes_to_plain_kernel(const __half* hi, const __half* lo, float* out, int B, int C, int T, int Tp,
                                                 float unscale = 1.f) {
  const long long total = (long long)B * C * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long bc = i / T;
    const int c = (int)(bc % C), b = (int)(bc / C);
    const size_t off = (((size_t)b * (C / 8) + c / 8) * Tp + kTcHalo + t) * 8 + (c & 7);
    out[i] = (__half2float(hi[off]) + __half2float(lo[off])) * unscale;
  }
}

}  // namespace dissc

// Tensor-core dilated Conv1d for the vocoder ResBlocks: tcgen05.mma (kind::f16) with
// split-precision operands so the result keeps fp32 accuracy.
//
//   out[t, co] = sum_j sum_ci A[t + j*d - pad, ci] * W_j[ci, co]
//
// is run as an implicit GEMM per 128-row time tile: D[128 x N] (fp32, in TMEM) accumulates,
// for every tap j and every 16-channel k-step, THREE MMAs
//     A_hi*W_hi + A_lo*W_hi + A_hi*W_lo          (x = x_hi + x_lo, both fp16)
// which reproduces the fp32 product to ~2^-22 (measured end to end: 1.6e-6 max-abs on the
// waveform vs fp64, the same as plain fp32; a single fp16/tf32 pass gives 7e-4 / 2e-3).
//
// Data layout (HBM): activations live channel-blocked and time-major,
//     planes  hi/lo : fp16 [B][C/8][Tp][8]   Tp = Tr + 2*HP, Tr = roundup(T,128), HP zero halo rows
//     f32b          : fp32 [B][C/8][Tr][8]   (residual / MRF accumulator)
// so that (a) one (c8, row-range) slab is CONTIGUOUS -> an activation block is fetched with
// plain 1-D bulk TMA copies and lands in shared memory as [c8][row][8], which is exactly the
// canonical no-swizzle K-major UMMA operand layout (core matrix = 8 rows x 16 B, SBO = 128 B,
// LBO = rows*16 B); (b) a tap shift of j*d rows is just +j*d*16 B on the descriptor start
// address, so one resident block serves all k taps; (c) the conv zero padding is the zero halo;
// (d) the epilogue thread that owns TMEM lane t holds 8 consecutive channels of row t =
// one 16 B (fp16) / 32 B (fp32) store, a warp writes 512 B / 1 KB contiguous.
// Weights are pre-split, pre-scaled by 2^s (so w_lo stays a normal fp16) and pre-packed
// [cb][tap][hi|lo][c8][co][8]: a pipeline stage is one contiguous bulk copy.
//
// CTA = 6 warps, persistent over tiles: warp 0 producer (bulk TMA), warp 1 MMA issuer (one
// thread) + TMEM owner, warps 2-5 epilogue (TMEM -> registers -> bias/residual/MRF/leaky-relu
// -> fp16 split -> global).  Two TMEM accumulators ping-pong so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Small layers keep the whole weight set resident in shared memory.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace dissc {

constexpr int kTcThreads = 192;
constexpr int kTcHalo = 32;      // HP: zero rows either side of every plane row-slab (>= max pad 25)
constexpr int kTcMaxStages = 64;  // barrier slots for the weight pipeline (resident mode uses one per stage)

struct TcConvParams {
  const __half* a_hi;  // planes [B][Cin/8][Tp][8]
  const __half* a_lo;
  const __half* w;     // packed [cb][k][2][KB/8][N][8]
  const float* bias;   // [N]
  float w_inv_scale;   // 2^-s
  const float* res;     // f32b or null
  const float* acc_in;  // f32b or null
  float* out_f32b;      // f32b or null (raw value)
  __half* out_hi;       // planes (leaky-relu(plane_slope) applied iff plane_act) or null
  __half* out_lo;
  float* out_plain;     // (B, N, T) fp32 (leaky-relu(plain_slope) iff plain_act) or null
  const int* lengths;
  int len_mul;
  int B, Cin, N, T, Tp, Tr;
  int k, dil, pad;
  int KB;         // channels per activation block (16 or 32)
  int n_cb;       // Cin / KB
  int JG;         // taps per weight stage
  int SPC;        // weight stages per channel block = ceil(k / JG)
  int NS;         // weight slots in shared memory
  int resident;   // 1: NS == n_cb*SPC, weights loaded once per CTA
  int tiles_per_b, n_tiles;
  int tmem_cols;  // allocation (power of two >= nbuf*2*N)
  int NA;         // activation-block buffers in shared memory (2..4)
  int nbuf;       // TMEM tile buffers (2 when 4*N <= 512, else 1); each holds a main and a cross accumulator
  float div;
  int plane_act, plain_act;
  float plane_slope, plain_slope;
};

__device__ __forceinline__ uint64_t umma_desc_kmajor_noswz(uint32_t saddr, uint32_t lbo_bytes) {
  // start address >> 4 | LBO >> 4 (K-chunk stride) << 16 | SBO (=128 B, 8-row group stride) >> 4 << 32 | version 1 << 46
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}

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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void split_store8(__half* hi_dst, __half* lo_dst, const float v[8]) {
  __align__(16) __half h[8];
  __align__(16) __half l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float c = fminf(fmaxf(v[i], -65504.f), 65504.f);
    h[i] = __float2half_rn(c);
    l[i] = __float2half_rn(c - __half2float(h[i]));
  }
  *reinterpret_cast<uint4*>(hi_dst) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(lo_dst) = *reinterpret_cast<const uint4*>(l);
}

__global__ void __launch_bounds__(kTcThreads, 1) conv1d_tc_kernel(const TcConvParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int R = 128 + (p.k - 1) * p.dil;                     // activation rows a tile needs
  const uint32_t a_plane_bytes = (uint32_t)(p.KB / 8) * R * 16;  // one plane of one block
  const uint32_t a_bytes = 2 * a_plane_bytes;
  const uint32_t w_plane_bytes = (uint32_t)(p.KB / 8) * p.N * 16;
  const uint32_t w_tap_bytes = 2 * w_plane_bytes;
  const uint32_t w_slot_bytes = (uint32_t)p.JG * w_tap_bytes;
  unsigned char* sA = smem_raw;
  unsigned char* sW = sA + (size_t)p.NA * a_bytes;
  float* s_bias = reinterpret_cast<float*>(sW + (size_t)p.NS * w_slot_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + ((p.N + 1) & ~1));
  uint64_t* a_full = bars;            // [4]
  uint64_t* a_empty = bars + 4;       // [4]
  uint64_t* acc_full = bars + 8;      // [2]
  uint64_t* acc_empty = bars + 10;    // [2]
  uint64_t* w_full = bars + 12;       // [NS]
  uint64_t* w_empty = w_full + p.NS;  // [NS]
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    for (int i = 0; i < p.NS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < p.N; i += kTcThreads) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int n8 = p.N / 8, cin8 = p.Cin / 8;

  if (warp == 0) {
    // ===================== producer: bulk TMA =====================
    if (lane == 0) {
      uint32_t qa = 0, qw = 0;
      bool first_tile = true;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_b;
        const int t0 = (tile - b * p.tiles_per_b) * 128;
        for (int cb = 0; cb < p.n_cb; ++cb) {
          const uint32_t buf = qa % (uint32_t)p.NA;
          mbar_wait(&a_empty[buf], ((qa / (uint32_t)p.NA) & 1) ^ 1);
          mbar_arrive_expect_tx(&a_full[buf], a_bytes);
          unsigned char* dst = sA + buf * a_bytes;
          for (int pl = 0; pl < 2; ++pl) {
            const __half* plane = pl ? p.a_lo : p.a_hi;
            for (int c8 = 0; c8 < p.KB / 8; ++c8) {
              const size_t row0 = ((size_t)b * cin8 + (size_t)cb * (p.KB / 8) + c8) * p.Tp + kTcHalo + t0 - p.pad;
              tma_load_1d(dst + pl * a_plane_bytes + (size_t)c8 * R * 16, plane + row0 * 8, (uint32_t)R * 16,
                          &a_full[buf]);
            }
          }
          ++qa;
          if (!p.resident || first_tile) {
            for (int g = 0; g < p.SPC; ++g) {
              const int j0 = g * p.JG;
              const int nt = min(p.JG, p.k - j0);
              const uint32_t slot = p.resident ? (uint32_t)(cb * p.SPC + g) : qw % (uint32_t)p.NS;
              if (!p.resident) mbar_wait(&w_empty[slot], ((qw / (uint32_t)p.NS) & 1) ^ 1);
              mbar_arrive_expect_tx(&w_full[slot], (uint32_t)nt * w_tap_bytes);
              tma_load_1d(sW + (size_t)slot * w_slot_bytes,
                          reinterpret_cast<const unsigned char*>(p.w) + ((size_t)cb * p.k + j0) * w_tap_bytes,
                          (uint32_t)nt * w_tap_bytes, &w_full[slot]);
              ++qw;
            }
          }
        }
        first_tile = false;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // instruction descriptor: D=f32 (1<<4), A=B=f16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lbo_a = (uint32_t)R * 16, lbo_b = (uint32_t)p.N * 16;
      uint32_t qa = 0, qw = 0, ti = 0;
      bool first_tile = true;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ti) {
        const uint32_t ab = ti % (uint32_t)p.nbuf;
        mbar_wait(&acc_empty[ab], ((ti / (uint32_t)p.nbuf) & 1) ^ 1);
        tc_fence_after();
        // Two accumulators per tile: the tensor core truncates the fp32 accumulator on every MMA, so the
        // tiny cross terms (2^-11 of the main term) get their own accumulator instead of re-rounding the
        // main one twice per k-step; the epilogue adds the two with a proper fp32 round-to-nearest.
        const uint32_t d_main = tmem_base + ab * 2u * (uint32_t)p.N;
        const uint32_t d_cross = d_main + (uint32_t)p.N;
        uint32_t accum = 0;
        for (int cb = 0; cb < p.n_cb; ++cb) {
          const uint32_t buf = qa % (uint32_t)p.NA;
          mbar_wait(&a_full[buf], (qa / (uint32_t)p.NA) & 1);
          const uint32_t a_hi = smem_u32(sA + buf * a_bytes);
          const uint32_t a_lo = a_hi + a_plane_bytes;
          for (int g = 0; g < p.SPC; ++g) {
            const int j0 = g * p.JG;
            const int nt = min(p.JG, p.k - j0);
            const uint32_t slot = p.resident ? (uint32_t)(cb * p.SPC + g) : qw % (uint32_t)p.NS;
            if (!p.resident)
              mbar_wait(&w_full[slot], (qw / (uint32_t)p.NS) & 1);
            else if (first_tile)
              mbar_wait(&w_full[slot], 0);
            tc_fence_after();
            const uint32_t wslot = smem_u32(sW + (size_t)slot * w_slot_bytes);
            for (int jj = 0; jj < nt; ++jj) {
              const uint32_t row_off = (uint32_t)((j0 + jj) * p.dil) * 16;
              const uint32_t w_hi = wslot + (uint32_t)jj * w_tap_bytes;
              const uint32_t w_lo = w_hi + w_plane_bytes;
              for (int ks = 0; ks < p.KB / 16; ++ks) {
                const uint64_t da_hi = umma_desc_kmajor_noswz(a_hi + row_off + ks * 2 * lbo_a, lbo_a);
                const uint64_t da_lo = umma_desc_kmajor_noswz(a_lo + row_off + ks * 2 * lbo_a, lbo_a);
                const uint64_t db_hi = umma_desc_kmajor_noswz(w_hi + ks * 2 * lbo_b, lbo_b);
                const uint64_t db_lo = umma_desc_kmajor_noswz(w_lo + ks * 2 * lbo_b, lbo_b);
                umma_f16(d_main, da_hi, db_hi, idesc, accum);
                umma_f16(d_cross, da_lo, db_hi, idesc, accum);
                umma_f16(d_cross, da_hi, db_lo, idesc, 1);
                accum = 1;
              }
            }
            if (!p.resident) umma_commit(&w_empty[slot]);
            ++qw;
          }
          umma_commit(&a_empty[buf]);
          ++qa;
        }
        umma_commit(&acc_full[ab]);
        first_tile = false;
      }
    }
  } else {
    // ===================== epilogue: 4 warps, TMEM lane quarter = warp % 4 =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ti) {
      const int b = tile / p.tiles_per_b;
      const int t = (tile - b * p.tiles_per_b) * 128 + row;
      const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
      const bool valid = t < Tvalid;
      const uint32_t ab = ti % (uint32_t)p.nbuf;
      mbar_wait(&acc_full[ab], (ti / (uint32_t)p.nbuf) & 1);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * 2u * (uint32_t)p.N;
      for (int c8 = 0; c8 < n8; ++c8) {
        uint32_t r[8], x[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr0 + c8 * 8));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                     : "r"(taddr0 + (uint32_t)p.N + c8 * 8));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[i] = (__uint_as_float(r[i]) + __uint_as_float(x[i])) * p.w_inv_scale + s_bias[c8 * 8 + i];
        const size_t fidx = (((size_t)b * n8 + c8) * p.Tr + t) * 8;
        if (valid) {
          if (p.res) {
            const float4 r0 = *reinterpret_cast<const float4*>(p.res + fidx);
            const float4 r1 = *reinterpret_cast<const float4*>(p.res + fidx + 4);
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
          }
          if (p.acc_in) {
            const float4 r0 = *reinterpret_cast<const float4*>(p.acc_in + fidx);
            const float4 r1 = *reinterpret_cast<const float4*>(p.acc_in + fidx + 4);
            v[0] = r0.x + v[0]; v[1] = r0.y + v[1]; v[2] = r0.z + v[2]; v[3] = r0.w + v[3];
            v[4] = r1.x + v[4]; v[5] = r1.y + v[5]; v[6] = r1.z + v[6]; v[7] = r1.w + v[7];
          }
          if (p.div != 0.f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v[i] / p.div;
          }
          if (p.out_f32b) {
            *reinterpret_cast<float4*>(p.out_f32b + fidx) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(p.out_f32b + fidx + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (p.out_plain) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              p.out_plain[((size_t)b * p.N + c8 * 8 + i) * p.T + t] = p.plain_act ? leaky(v[i], p.plain_slope) : v[i];
          }
        }
        if (p.out_hi) {
          // rows >= valid length are written as zeros: they are the next conv's zero padding
          float a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = valid ? (p.plane_act ? leaky(v[i], p.plane_slope) : v[i]) : 0.f;
          const size_t pidx = (((size_t)b * n8 + c8) * p.Tp + kTcHalo + t) * 8;
          split_store8(p.out_hi + pidx, p.out_lo + pidx, a);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

// Zero rows [0,HP) and [HP+T, Tp) of every (b, c8) slab of a plane pair: the conv zero padding
// (left halo, the round-up rows [T,Tr) and the right halo).
__global__ void tc_zero_halos_kernel(__half* hi, __half* lo, int slabs, int Tp, int T) {
  const int per = kTcHalo + (Tp - kTcHalo - T);  // rows per slab to clear
  const long long total = (long long)slabs * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i / per);
    int r = (int)(i - (long long)s * per);
    r = r < kTcHalo ? r : (T + r);  // second range starts at row HP+T
    const size_t off = ((size_t)s * Tp + r) * 8;
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(0, 0, 0, 0);
  }
}

// plain (B,C,T) fp32 -> split planes [B][C/8][Tp][8] (+ optional leaky-relu); rows >= valid length are zero.
// Layer-test helper (the model writes planes straight from the producing kernel's epilogue).
__global__ void tc_pack_planes_kernel(const float* in, __half* hi, __half* lo, const int* lengths, int len_mul, int B, int C,
                                      int T, int Tp, int act, float slope) {
  const long long total = (long long)B * (C / 8) * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long s = i / T;  // slab = b*C/8 + c8
    const int b = (int)(s / (C / 8)), c8 = (int)(s % (C / 8));
    const int Tvalid = lengths ? min(T, lengths[b] * len_mul) : T;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float x = (t < Tvalid) ? in[((size_t)b * C + c8 * 8 + e) * T + t] : 0.f;
      v[e] = act ? leaky(x, slope) : x;
    }
    const size_t off = ((size_t)s * Tp + kTcHalo + t) * 8;
    split_store8(hi + off, lo + off, v);
  }
}

// plain (B,C,T) fp32 <-> blocked f32b [B][C/8][Tr][8] (layer-test helpers)
__global__ void tc_plain_to_f32b_kernel(const float* in, float* out, int B, int C, int T, int Tr) {
  const long long total = (long long)B * C * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long bc = i / T;
    const int c = (int)(bc % C), b = (int)(bc / C);
    out[(((size_t)b * (C / 8) + c / 8) * Tr + t) * 8 + (c & 7)] = in[i];
  }
}
__global__ void tc_f32b_to_plain_kernel(const float* in, float* out, int B, int C, int T, int Tr) {
  const long long total = (long long)B * C * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long bc = i / T;
    const int c = (int)(bc % C), b = (int)(bc / C);
    out[i] = in[(((size_t)b * (C / 8) + c / 8) * Tr + t) * 8 + (c & 7)];
  }
}
// planes -> plain fp32 (hi + lo), for tests
__global__ void tc_planes_to_plain_kernel(const __half* hi, const __half* lo, float* out, int B, int C, int T, int Tp) {
  const long long total = (long long)B * C * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long bc = i / T;
    const int c = (int)(bc % C), b = (int)(bc / C);
    const size_t off = (((size_t)b * (C / 8) + c / 8) * Tp + kTcHalo + t) * 8 + (c & 7);
    out[i] = __half2float(hi[off]) + __half2float(lo[off]);
  }
}

}  // namespace dissc

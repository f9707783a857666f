// Polyphase ConvTranspose1d (fp32), the vocoder's upsamplers (sr/models.py:82-86,:102).
//
//   out[b,co,t] = bias[co] + sum_ci sum_{i : 0 <= t+p-i*u < k} in[b,ci,i] * W[ci,co,t+p-i*u]
//
// With s = t+p, q = s / u ("frame"), phi = s % u, the taps of output s are
// j = phi + m*u (m = 0..ceil(k/u)-1) reading input frame q-m: every phase is a
// tiny dense conv at the INPUT rate, no zero insertion.  The producer of `in`
// already applied the leaky-relu (conv_pre / the previous stage's last
// epilogue), so the input is consumed as is.
//
// A CTA of 256 threads owns CO_TILE channels x F_TILE frames (x u phases);
// each thread keeps RCO(4) x RF(4) x U accumulators, frames strided across
// lanes (conflict-free shared-memory reads).  Weight chunks arrive by 1-D bulk
// TMA, activation rows are staged through registers, double-buffered.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace dissc {

struct ConvTParams {
  const float* in;     // (B, Cin, Tin)   already activated
  const float* w;      // packed [co_tile][chunk][CI_CHUNK][KW][CO_TILE]
  const float* bias;   // (Cout)
  float* out;          // (B, Cout, Tout)
  const int* lengths;  // (B) or null
  int len_mul;         // valid input frames = lengths[b]*len_mul
  int B, Cin, Cout, Tin, Tout;
  int pad;
  // tensor-core stage hand-off (conv1d_tc.cuh layouts); when out_f32b != null `out` is unused:
  float* out_f32b;     // raw x_up, fp32 [B][Cout/8][Tr][8]
  __half* out_hi;      // leaky_relu(x_up, plane_slope) split into fp16 planes [B][Cout/8][Tp][8]
  __half* out_lo;
  int Tr, Tp, halo;
  float plane_slope;
};

template <int CO_TILE, int KW, int U, int CI_CHUNK>
struct ConvTCfg {
  static constexpr int RCO = 4, RF = 4;
  static constexpr int M = (KW + U - 1) / U;  // taps per phase (max)
  static constexpr int NCG = CO_TILE / RCO;
  static constexpr int TT = kThreads / NCG;
  static constexpr int F_TILE = TT * RF;
  static constexpr int XROW = F_TILE + M - 1;
  static constexpr int W_CHUNK = CI_CHUNK * KW * CO_TILE;
  static constexpr int X_CHUNK = CI_CHUNK * XROW;
  static constexpr int NX = (X_CHUNK + kThreads - 1) / kThreads;
  static constexpr size_t SMEM = sizeof(float) * 2 * (W_CHUNK + X_CHUNK) + 2 * sizeof(uint64_t);
  static_assert(TT >= 32 && TT % 32 == 0, "a warp must share one channel group");
  static_assert((W_CHUNK * 4) % 16 == 0, "bulk copy needs 16-byte multiples");
};

template <int CO_TILE, int KW, int U, int CI_CHUNK>
__global__ void __launch_bounds__(kThreads, 2) convt1d_kernel(const ConvTParams p) {
  using C = ConvTCfg<CO_TILE, KW, U, CI_CHUNK>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ws = reinterpret_cast<float*>(smem_raw);
  float* xs = ws + 2 * C::W_CHUNK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs + 2 * C::X_CHUNK);

  const int tid = threadIdx.x;
  const int cg = tid / C::TT;
  const int tl = tid % C::TT;
  const int b = blockIdx.z;
  const int co_tile = blockIdx.y;
  const int q0 = blockIdx.x * C::F_TILE;
  const int Tvalid = p.lengths ? min(p.Tin, p.lengths[b] * p.len_mul) : p.Tin;
  // reads only padding; outputs never consumed (the tensor-core hand-off needs its zeros written, so no exit there)
  if (q0 - (C::M - 1) >= Tvalid && q0 > 0 && !p.out_f32b) return;
  const int nchunks = (p.Cin + CI_CHUNK - 1) / CI_CHUNK;
  const float* wbase = p.w + (size_t)co_tile * nchunks * C::W_CHUNK;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  float xr[C::NX];
  auto load_x = [&](int chunk) {
#pragma unroll
    for (int i = 0; i < C::NX; ++i) {
      const int e = tid + i * kThreads;
      float v = 0.f;
      if (e < C::X_CHUNK) {
        const int cl = e / C::XROW;
        const int ff = e - cl * C::XROW;
        const int ci = chunk * CI_CHUNK + cl;
        const int f = q0 - (C::M - 1) + ff;
        if (ci < p.Cin && f >= 0 && f < Tvalid) v = __ldg(p.in + ((size_t)b * p.Cin + ci) * p.Tin + f);
      }
      xr[i] = v;
    }
  };
  auto store_x = [&](int buf) {
    float* dst = xs + buf * C::X_CHUNK;
#pragma unroll
    for (int i = 0; i < C::NX; ++i) {
      const int e = tid + i * kThreads;
      if (e < C::X_CHUNK) dst[e] = xr[i];
    }
  };

  float acc[C::RCO][C::RF][U];
#pragma unroll
  for (int c = 0; c < C::RCO; ++c)
#pragma unroll
    for (int f = 0; f < C::RF; ++f)
#pragma unroll
      for (int u = 0; u < U; ++u) acc[c][f][u] = 0.f;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], C::W_CHUNK * 4);
    tma_load_1d(ws, wbase, C::W_CHUNK * 4, &bars[0]);
  }
  load_x(0);
  store_x(0);
  __syncthreads();

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    const bool has_next = (c + 1 < nchunks);
    if (has_next) {
      if (tid == 0) {
        mbar_arrive_expect_tx(&bars[buf ^ 1], C::W_CHUNK * 4);
        tma_load_1d(ws + (buf ^ 1) * C::W_CHUNK, wbase + (size_t)(c + 1) * C::W_CHUNK, C::W_CHUNK * 4,
                    &bars[buf ^ 1]);
      }
      load_x(c + 1);
    }
    mbar_wait(&bars[buf], (c >> 1) & 1);

    const float* wsb = ws + buf * C::W_CHUNK + cg * C::RCO;
    const float* xsb = xs + buf * C::X_CHUNK + tl + (C::M - 1);
#pragma unroll 1
    for (int ci = 0; ci < CI_CHUNK; ++ci) {
      const float* wrow = wsb + ci * KW * CO_TILE;
      const float* xrow = xsb + ci * C::XROW;
      float xv[C::RF][C::M];
#pragma unroll
      for (int f = 0; f < C::RF; ++f)
#pragma unroll
        for (int m = 0; m < C::M; ++m) xv[f][m] = xrow[f * C::TT - m];
#pragma unroll
      for (int j = 0; j < KW; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + j * CO_TILE);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
        const int phi = j % U;
        const int m = j / U;
#pragma unroll
        for (int cc = 0; cc < C::RCO; ++cc)
#pragma unroll
          for (int f = 0; f < C::RF; ++f) acc[cc][f][phi] = fmaf(wv[cc], xv[f][m], acc[cc][f][phi]);
      }
    }
    if (has_next) store_x(buf ^ 1);
    __syncthreads();
  }

  const int co_base = co_tile * CO_TILE + cg * C::RCO;
  if (p.out_f32b) {
    // blocked hand-off: this thread's 4 consecutive channels are one 16 B (fp32) / 8 B (fp16) store per time step
    if (co_base + C::RCO > p.Cout) return;  // Cout % 16 == 0 in this mode
    float bv[C::RCO];
#pragma unroll
    for (int cc = 0; cc < C::RCO; ++cc) bv[cc] = p.bias ? __ldg(p.bias + co_base + cc) : 0.f;
    const int Tv_out = p.lengths ? min(p.Tout, Tvalid * U) : p.Tout;
    const size_t slab = (size_t)b * (p.Cout / 8) + co_base / 8;
    const int sub = co_base & 7;
#pragma unroll
    for (int f = 0; f < C::RF; ++f) {
      const int q = q0 + tl + f * C::TT;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = q * U + u - p.pad;
        if (t < 0 || t >= p.Tout) continue;
        float v[C::RCO];
#pragma unroll
        for (int cc = 0; cc < C::RCO; ++cc) v[cc] = acc[cc][f][u] + bv[cc];
        const bool valid = t < Tv_out;
        if (valid) *reinterpret_cast<float4*>(p.out_f32b + (slab * p.Tr + t) * 8 + sub) = make_float4(v[0], v[1], v[2], v[3]);
        __align__(8) __half h[4];
        __align__(8) __half l[4];
#pragma unroll
        for (int cc = 0; cc < C::RCO; ++cc) {
          float a = valid ? leaky(v[cc], p.plane_slope) : 0.f;
          a = fminf(fmaxf(a, -65504.f), 65504.f);
          h[cc] = __float2half_rn(a);
          l[cc] = __float2half_rn(a - __half2float(h[cc]));
        }
        const size_t po = (slab * p.Tp + p.halo + t) * 8 + sub;
        *reinterpret_cast<uint2*>(p.out_hi + po) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(p.out_lo + po) = *reinterpret_cast<const uint2*>(l);
      }
    }
    return;
  }
#pragma unroll
  for (int cc = 0; cc < C::RCO; ++cc) {
    const int co = co_base + cc;
    if (co >= p.Cout) continue;
    const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
    float* orow = p.out + ((size_t)b * p.Cout + co) * p.Tout;
#pragma unroll
    for (int f = 0; f < C::RF; ++f) {
      const int q = q0 + tl + f * C::TT;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = q * U + u - p.pad;
        if (t >= 0 && t < p.Tout) orow[t] = acc[cc][f][u] + bv;
      }
    }
  }
}

}  // namespace dissc

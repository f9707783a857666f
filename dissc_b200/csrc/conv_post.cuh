// conv_post + tanh (+ optional int16 conversion): sr/models.py:110-112, sr/inference.py:73-75.
//
//   y[b,t] = tanh(bias + sum_{ci,j} w[ci,j] * a[b,ci,t+j-pad])
//
// `a` already carries the final LeakyReLU (slope 0.01, :110) -- applied by the
// epilogue of the last MRF kernel.  Pure streaming op: Cin rows in, one row
// out; the tile (+halo) is staged in shared memory with coalesced loads and
// each thread produces RT outputs strided across the tile.
#pragma once
#include "common.cuh"

namespace dissc {

struct ConvPostParams {
  const float* in;    // (B, Cin, T)
  const float* w;     // (Cin, KW) fp32 (Cout == 1)
  const float* bias;  // (1)
  float* out_f32;     // (B, T) or null
  int16_t* out_i16;   // (B, T) or null
  const int* lengths;
  int len_mul;
  int B, Cin, T;
};

constexpr int kPostTile = 2048;
constexpr int kPostMaxCin = 32;

template <int KW>
__global__ void __launch_bounds__(kThreads) conv_post_kernel(const ConvPostParams p) {
  constexpr int PAD = (KW - 1) / 2;
  constexpr int XROW = kPostTile + KW - 1;
  constexpr int CI_STEP = 4;
  __shared__ float xs[CI_STEP][XROW];
  __shared__ float wsm[kPostMaxCin * KW];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kPostTile;
  const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
  for (int i = tid; i < p.Cin * KW; i += kThreads) wsm[i] = p.w[i];
  constexpr int RT = kPostTile / kThreads;
  float acc[RT];
#pragma unroll
  for (int m = 0; m < RT; ++m) acc[m] = 0.f;
  for (int c0 = 0; c0 < p.Cin; c0 += CI_STEP) {
    __syncthreads();
    for (int e = tid; e < CI_STEP * XROW; e += kThreads) {
      const int cl = e / XROW, tt = e - cl * XROW;
      const int ci = c0 + cl, t = t0 - PAD + tt;
      float v = 0.f;
      if (ci < p.Cin && t >= 0 && t < Tvalid) v = __ldg(p.in + ((size_t)b * p.Cin + ci) * p.T + t);
      xs[cl][tt] = v;
    }
    __syncthreads();
#pragma unroll
    for (int cl = 0; cl < CI_STEP; ++cl) {
      if (c0 + cl >= p.Cin) break;
#pragma unroll
      for (int j = 0; j < KW; ++j) {
        const float wv = wsm[(c0 + cl) * KW + j];
#pragma unroll
        for (int m = 0; m < RT; ++m) acc[m] = fmaf(wv, xs[cl][tid + m * kThreads + j], acc[m]);
      }
    }
  }
  const float bv = p.bias[0];
#pragma unroll
  for (int m = 0; m < RT; ++m) {
    const int t = t0 + tid + m * kThreads;
    if (t >= p.T) continue;
    float y = (t < Tvalid) ? tanhf(acc[m] + bv) : 0.f;
    if (p.out_f32) p.out_f32[(size_t)b * p.T + t] = y;
    if (p.out_i16) {
      // numpy: (y*32768).astype(int16) -- truncate toward zero, then wrap modulo 2^16
      const int iv = (int)(y * 32768.0f);
      p.out_i16[(size_t)b * p.T + t] = (int16_t)(unsigned short)(iv & 0xffff);
    }
  }
}

}  // namespace dissc

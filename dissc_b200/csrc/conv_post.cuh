// conv_post + tanh (+ optional int16 conversion): sr/models.py:110-112, sr/inference.py:73-75.
//
//   y[b,t] = tanh(bias + sum_{ci,j} w[ci,j] * a[b,ci,t+j-pad])
//
// `a` already carries the final LeakyReLU (slope 0.01, :110) -- applied by the
// epilogue of the last MRF kernel.  Pure streaming op: Cin rows in, one row out.
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

constexpr int kPostMaxCin = 32;
constexpr int kPostRT = 8;  // consecutive outputs per thread

// Register-tiled streaming version: a thread produces kPostRT consecutive samples of one utterance.  Per input channel it
// needs x[t-PAD .. t+RT-1+PAD]: four aligned 16-byte loads (t-4 .. t+11) straight from global memory when the rows are
// 16-byte aligned (T % 4 == 0, always true for hop 320); neighbouring threads' overlaps hit L1.  No shared-memory
// staging, no barriers: the kernel is bound by the single pass over the Cin x T activation (HBM).
template <int KW>
__global__ void __launch_bounds__(kThreads) conv_post_kernel(const ConvPostParams p) {
  constexpr int PAD = (KW - 1) / 2;
  static_assert(PAD <= 4 && kPostRT == 8, "window t-4 .. t+11 must cover the taps");
  __shared__ float wsm[kPostMaxCin * KW];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t = (blockIdx.x * kThreads + tid) * kPostRT;   // first output of this thread
  const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
  for (int i = tid; i < p.Cin * KW; i += kThreads) wsm[i] = p.w[i];
  __syncthreads();
  if (t >= p.T) return;
  float acc[kPostRT];
#pragma unroll
  for (int m = 0; m < kPostRT; ++m) acc[m] = 0.f;
  const bool in_al = (reinterpret_cast<uintptr_t>(p.in) & 15) == 0;
  const bool out_al = ((reinterpret_cast<uintptr_t>(p.out_f32) | reinterpret_cast<uintptr_t>(p.out_i16)) & 15) == 0;
  const bool vec = in_al && (p.T & 3) == 0 && t >= 4 && t + 12 <= Tvalid;   // whole window inside the valid signal, aligned
  if (t < Tvalid) {
    for (int ci = 0; ci < p.Cin; ++ci) {
      const float* row = p.in + ((size_t)b * p.Cin + ci) * p.T;
      float x[16];  // x[i] = a[t - 4 + i]
      if (vec) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(row + t - 4) + q);
          x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int tt = t - 4 + i;
          x[i] = (tt >= 0 && tt < Tvalid) ? __ldg(row + tt) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < KW; ++j) {
        const float wv = wsm[ci * KW + j];
#pragma unroll
        for (int m = 0; m < kPostRT; ++m) acc[m] = fmaf(wv, x[m + j + 4 - PAD], acc[m]);
      }
    }
  }
  const float bv = p.bias[0];
  float y[kPostRT];
#pragma unroll
  for (int m = 0; m < kPostRT; ++m) y[m] = (t + m < Tvalid) ? tanhf(acc[m] + bv) : 0.f;
  if (out_al && (p.T & 7) == 0) {   // t is a multiple of 8: the thread's 8 outputs are one aligned 32-byte / 16-byte store
    if (p.out_f32) {
      float4* o = reinterpret_cast<float4*>(p.out_f32 + (size_t)b * p.T + t);
      o[0] = make_float4(y[0], y[1], y[2], y[3]);
      o[1] = make_float4(y[4], y[5], y[6], y[7]);
    }
    if (p.out_i16) {
      // numpy: (y*32768).astype(int16) -- truncate toward zero, then wrap modulo 2^16
      uint32_t w[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int i0 = (int)(y[2 * m] * 32768.0f), i1 = (int)(y[2 * m + 1] * 32768.0f);
        w[m] = (uint32_t)(i0 & 0xffff) | ((uint32_t)(i1 & 0xffff) << 16);
      }
      *reinterpret_cast<uint4*>(p.out_i16 + (size_t)b * p.T + t) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
#pragma unroll
    for (int m = 0; m < kPostRT; ++m) {
      if (t + m >= p.T) break;
      if (p.out_f32) p.out_f32[(size_t)b * p.T + t + m] = y[m];
      if (p.out_i16) {
        const int iv = (int)(y[m] * 32768.0f);
        p.out_i16[(size_t)b * p.T + t + m] = (int16_t)(unsigned short)(iv & 0xffff);
      }
    }
  }
}

}  // namespace dissc

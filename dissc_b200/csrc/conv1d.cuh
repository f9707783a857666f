// Fused dilated Conv1d for the vocoder (fp32 CUDA-core implicit GEMM).
//
//   out = post( [acc_in +] [res +] bias + conv1d(pre(in), w, dilation) ) [/ div]
//
// One kernel covers: conv_pre with the embedding gather/concat fused into its
// operand load (sr/models.py:189,:206-215,:99), ResBlock1's first half
// lrelu -> dilated conv -> lrelu (:36-38), its second half conv -> +x (:39-40),
// and the MRF accumulate / divide (:104-109) folded into the last epilogue.
//
// Tiling: a CTA of 256 threads owns CO_TILE output channels x T_TILE time
// steps of one utterance.  Each thread keeps an 8(co) x 8(t) fp32 register
// tile; its 8 time steps are strided by the number of time-threads so that a
// warp's shared-memory reads of the input row are conflict-free and its
// global stores are 128-byte coalesced.  The reduction over input channels
// runs in chunks of CI_CHUNK: the weight tile of a chunk (pre-packed
// [ci][tap][co], contiguous) arrives by one 1-D bulk TMA copy signalled on an
// mbarrier, the activation rows (with the dilation halo, zero outside
// [0, valid length), leaky-relu applied once on the way in) are staged
// through registers; both are double-buffered against the FMA loop.
#pragma once
#include "common.cuh"

namespace dissc {

struct ConvParams {
  // input: plain activations ...
  const float* in;  // (B, Cin, T)
  // ... or the fused embedding concat (IN_EMBED): channels [0,E) = dict[code], [E] = f0, then spkr_emb
  const long long* code;  // (B, T)
  const float* f0;        // (B, T) or null
  const long long* spkr;  // (B) or null
  const float* dict_w;    // (num_embeddings, E)
  const float* spkr_w;    // (rows, E)
  int E, f0_ch, spk_base; // f0_ch = -1 if absent; spk_base = first speaker channel or -1
  const float* extra;     // (B, n_extra) or null: per-utterance conditioning features, channels [extra_base, +n_extra)
  int extra_base, n_extra;
  int n_code_rows, n_spkr_rows;  // table rows: ids outside [0, rows) set *err (common.cuh::checked_row)
  int* err;

  const float* w;       // packed [co_tile][chunk][CI_CHUNK][KW][CO_TILE], zero padded
  const float* bias;    // (Cout)
  const float* res;     // (B, Cout, T) or null
  const float* acc_in;  // (B, Cout, T) or null
  float* out;           // (B, Cout, T)
  const int* lengths;   // (B) or null
  int len_mul;          // valid time steps = lengths[b] * len_mul
  int B, Cin, Cout, T;
  int pad;
  int pre_act, post_act;
  float pre_slope, post_slope;
  float div;  // 0 = none
};

template <int CO_TILE, int KW, int DIL, int CI_CHUNK>
struct ConvCfg {
  static constexpr int RCO = 8, RT = 8;
  static constexpr int NCG = CO_TILE / RCO;          // channel groups of 8
  static constexpr int TT = kThreads / NCG;          // time threads
  static constexpr int T_TILE = TT * RT;
  static constexpr int HALO = (KW - 1) * DIL;
  static constexpr int XROW = T_TILE + HALO;
  static constexpr int W_CHUNK = CI_CHUNK * KW * CO_TILE;  // floats per weight chunk
  static constexpr int X_CHUNK = CI_CHUNK * XROW;          // floats per activation chunk
  static constexpr int NX = (X_CHUNK + kThreads - 1) / kThreads;
  static constexpr size_t SMEM = sizeof(float) * 2 * (W_CHUNK + X_CHUNK) + 2 * sizeof(uint64_t);
  static_assert(TT >= 32 && TT % 32 == 0, "a warp must share one channel group");
  static_assert((W_CHUNK * 4) % 16 == 0, "bulk copy needs 16-byte multiples");
};

template <int CO_TILE, int KW, int DIL, int CI_CHUNK, bool IN_EMBED>
__global__ void __launch_bounds__(kThreads, 2) conv1d_fused_kernel(const ConvParams p) {
  using C = ConvCfg<CO_TILE, KW, DIL, CI_CHUNK>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ws = reinterpret_cast<float*>(smem_raw);                 // [2][W_CHUNK]
  float* xs = ws + 2 * C::W_CHUNK;                                // [2][X_CHUNK]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs + 2 * C::X_CHUNK);

  const int tid = threadIdx.x;
  const int cg = tid / C::TT;   // warp-uniform
  const int tl = tid % C::TT;
  const int b = blockIdx.z;
  const int co_tile = blockIdx.y;
  const int t0 = blockIdx.x * C::T_TILE;
  const int Tvalid = p.lengths ? min(p.T, p.lengths[b] * p.len_mul) : p.T;
  if (t0 >= Tvalid + p.pad && t0 > 0) {
    // tile entirely in the padded region: every tap reads zeros; nothing downstream reads it.
    return;
  }
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
        const int tt = e - cl * C::XROW;
        const int ci = chunk * CI_CHUNK + cl;
        const int t = t0 - p.pad + tt;
        if (ci < p.Cin && t >= 0 && t < Tvalid) {
          if constexpr (IN_EMBED) {
            if (ci < p.E) {
              v = __ldg(p.dict_w + (size_t)checked_row(p.code[(size_t)b * p.T + t], p.n_code_rows, p.err, kIdxUnit) * p.E + ci);
            } else if (ci == p.f0_ch) {
              v = __ldg(p.f0 + (size_t)b * p.T + t);
            } else if (p.n_extra > 0 && ci >= p.extra_base) {
              v = __ldg(p.extra + (size_t)b * p.n_extra + (ci - p.extra_base));
            } else {
              v = __ldg(p.spkr_w + (size_t)checked_row(p.spkr[b], p.n_spkr_rows, p.err, kIdxSpeaker) * p.E + (ci - p.spk_base));
            }
          } else {
            v = __ldg(p.in + ((size_t)b * p.Cin + ci) * p.T + t);
          }
        }
      }
      xr[i] = v;
    }
  };
  auto store_x = [&](int buf) {
    float* dst = xs + buf * C::X_CHUNK;
#pragma unroll
    for (int i = 0; i < C::NX; ++i) {
      const int e = tid + i * kThreads;
      // leaky-relu applied here, after the FMA loop, so the global load's latency hides behind it
      if (e < C::X_CHUNK) dst[e] = p.pre_act ? leaky(xr[i], p.pre_slope) : xr[i];
    }
  };

  float acc[C::RCO][C::RT];
#pragma unroll
  for (int c = 0; c < C::RCO; ++c)
#pragma unroll
    for (int m = 0; m < C::RT; ++m) acc[c][m] = 0.f;

  // prologue: chunk 0
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
    const float* xsb = xs + buf * C::X_CHUNK + tl;
#pragma unroll 1
    for (int ci = 0; ci < CI_CHUNK; ++ci) {
      const float* wrow = wsb + ci * KW * CO_TILE;
      const float* xrow = xsb + ci * C::XROW;
#pragma unroll
      for (int j = 0; j < KW; ++j) {
        const float4 wa = *reinterpret_cast<const float4*>(wrow + j * CO_TILE);
        const float4 wb = *reinterpret_cast<const float4*>(wrow + j * CO_TILE + 4);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        float xv[C::RT];
#pragma unroll
        for (int m = 0; m < C::RT; ++m) xv[m] = xrow[m * C::TT + j * DIL];
#pragma unroll
        for (int cc = 0; cc < C::RCO; ++cc)
#pragma unroll
          for (int m = 0; m < C::RT; ++m) acc[cc][m] = fmaf(wv[cc], xv[m], acc[cc][m]);
      }
    }
    if (has_next) store_x(buf ^ 1);
    __syncthreads();
  }

  // epilogue
  const int co_base = co_tile * CO_TILE + cg * C::RCO;
#pragma unroll
  for (int cc = 0; cc < C::RCO; ++cc) {
    const int co = co_base + cc;
    if (co >= p.Cout) continue;
    const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
    const size_t row = ((size_t)b * p.Cout + co) * p.T;
#pragma unroll
    for (int m = 0; m < C::RT; ++m) {
      const int t = t0 + tl + m * C::TT;
      if (t >= p.T) continue;
      float v = acc[cc][m] + bv;
      if (p.res) v += p.res[row + t];
      if (p.acc_in) v = p.acc_in[row + t] + v;
      if (p.div != 0.f) v = v / p.div;
      if (p.post_act) v = leaky(v, p.post_slope);
      p.out[row + t] = v;
    }
  }
}

}  // namespace dissc

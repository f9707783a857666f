// Host-side planning / packing / launch of the tensor-core conv kernel (conv_tc.cuh), shared between translation
// units (definitions in generator.cu).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "conv_tc.cuh"

namespace dissc {

struct TcLayer {
  bool ok = false;
  int Cin = 0, Cin_pad = 0, Cout = 0, NC = 0, n_chunks = 1;
  int k = 0, dil = 1, pad = 0;       // taps / dilation / left padding of the implicit GEMM
  int up = 0, up_P = 0, up_pad = 0;  // transposed conv: stride, phases per chunk, padding
  int KB = 0, n_cb = 0, JG = 0, SPC = 0, NS = 0, resident = 0, tmem_cols = 0, acc_cols = 0, nbuf = 1, NA = 2;
  int single_acc = 0, ctas_per_sm = 1;
  int cin8_total = 0;                // 8-channel groups per batch row of the INPUT tensor (0: Cin_pad/8)
  int groups = 0, group_c8 = 0;      // grouped conv: n_chunks groups of group_c8*8 output channels (NC-padded)
  int split_w = 1;                   // separate weight-producer thread (N >= 128 kernels)
  int cluster2 = -1;                 // 2-CTA clusters with a multicast weight stream: 1 / 0, -1 = the global default
  int pair2 = 0;                     // CTA pairs issuing 256-row cta_group::2 MMAs (weights packed per column half)
  int cb_split = 0, k_hi = 0;        // channel blocks >= cb_split use only their first k_hi taps (rest: structural zeros)
  size_t smem = 0;
  __half* w = nullptr;  // packed [chunk][cb][tap][KB/8][hi|lo][NC][8], device
  float inv_scale = 1.f;
};

constexpr size_t kSmemPerSm = 227 * 1024;

// Cin input channels, ncols GEMM columns (Cout, or u*Cout for a transposed conv), `taps` shifted by `dil` rows, left
// padding `pad` rows; the plane buffers carry `halo` zero rows either side.
// force_nc: chunk width (0 = by the column count); single_acc: -1 = the process default (NC = 256 only)
// pair2: plan for CTA pairs (conv_tc.cuh): half of every weight stage per CTA, so twice the stages in the same shared memory
bool tc_plan(int Cin, int ncols, int taps, int dil, int pad, TcLayer* L, int halo, int force_nc = 0, int single_acc = -1,
             int pair2 = 0);
bool tc_plan_conv(int Cin, int Cout, int k, int dil, TcLayer* L);
// p carries the tensors, B, T (output rows), Tr, Tp, Tp_in, lengths and the epilogue switches; `rows` is the number of
// GEMM rows per utterance (output time steps for a conv, input frames for a transposed conv).
int launch_conv_tc(TcParams p, const TcLayer& L, int rows, cudaStream_t st);
int launch_zero_halos(__half* hi, __half* lo, int slabs, int Tp, int T, cudaStream_t st, int halo = kTcHalo);

// Generic packer: wval(n, ci, tap) is the GEMM weight of column n (0 <= n < n_chunks*NC).  Output fp16 hi/lo planes of
// w*2^s, layout [chunk][cb][tap][KB/8][hi|lo][NC][8].
template <typename F>
std::vector<__half> pack_weights_tc(const TcLayer& L, F wval, float* inv_scale) {
  const int ncols = L.n_chunks * L.NC;
  float mx = 0.f;
  for (int n = 0; n < ncols; ++n)
    for (int ci = 0; ci < L.Cin; ++ci)
      for (int j = 0; j < L.k; ++j) mx = std::max(mx, std::fabs(wval(n, ci, j)));
  int s = 0;
  if (mx > 0.f) {
    int e;
    std::frexp(mx, &e);  // mx = f * 2^e, f in [0.5,1)
    s = 4 - e;           // mx * 2^s in [8,16)
    s = std::max(-14, std::min(24, s));
  }
  const float scale = std::ldexp(1.f, s);
  *inv_scale = std::ldexp(1.f, -s);
  const int kb8 = L.KB / 8;
  std::vector<__half> out((size_t)L.n_chunks * L.n_cb * L.k * kb8 * 2 * L.NC * 8);
  size_t o = 0;
  for (int ch = 0; ch < L.n_chunks; ++ch)
    for (int cb = 0; cb < L.n_cb; ++cb)
      for (int j = 0; j < L.k; ++j)
        // pair2: [rank][KB/8][hi|lo][NC/2][8] -- each CTA of a pair streams the columns [rank*NC/2, rank*NC/2 + NC/2)
        for (int half = 0; half < (L.pair2 ? 2 : 1); ++half) {
          const int wc = L.pair2 ? L.NC / 2 : L.NC, n0 = half * wc;
          for (int c8 = 0; c8 < kb8; ++c8) {
            __half* hi = &out[o];
            __half* lo = hi + (size_t)wc * 8;
            o += (size_t)2 * wc * 8;
            for (int n = 0; n < wc; ++n)
              for (int e = 0; e < 8; ++e) {
                const int ci = cb * L.KB + c8 * 8 + e;
                const float v = (ci < L.Cin ? wval(ch * L.NC + n0 + n, ci, j) : 0.f) * scale;
                const __half h = __float2half_rn(v);
                hi[n * 8 + e] = h;
                lo[n * 8 + e] = __float2half_rn(v - __half2float(h));
              }
          }
        }
  return out;
}


}  // namespace dissc

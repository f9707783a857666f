// HuBERT-base unit encoder on the GPU (include/dissc_b200.h, "Unit encoder").
//
// Replaces textless `SpeechEncoder.forward` as called at data/encode.py:21-22,32: HuBERT-base (fairseq
// hubert_base_ls960) features at transformer layer 6 -> nearest of K k-means centroids.  The arithmetic is third-party
// (textlesslib / fairseq@dd106d95, absent from the reference tree -- SURVEY.md 8c), restated here from the published
// architecture (fairseq models/hubert/hubert.py + models/wav2vec/wav2vec2.py, extractor mode "default", post-LN
// encoder); oracle/hubert_oracle.py is the CPU restatement and is cross-checked against torchaudio.models.hubert_base.
//
//   wave (B,N) fp32
//   conv0   Conv1d(1->512,k10,s5,no bias) -> GroupNorm(512,512) over time -> GELU      CUDA cores, 2 passes (stats, apply)
//   conv1-6 Conv1d(512->512,k3|2,s2,no bias) -> GELU                                   tcgen05 (conv_tc.cuh): the producer
//           writes its output de-interleaved (even / odd time steps as two channel planes), so a stride-2 conv becomes
//           a stride-1 implicit GEMM over 1024-channel "frames" with ceil(k/2) taps
//   LayerNorm(512) -> Linear 512->768                                                  row kernel + tcgen05 (k=1)
//   x + GELU(pos_conv(x)): Conv1d(768,768,k128,pad64,groups16), last frame dropped     tcgen05 grouped implicit GEMM
//   LayerNorm(768); 6 x post-LN layer: fused QKV GEMM -> attention (fp32, online softmax) -> out proj + residual -> LN
//           -> fc1 + GELU -> fc2 + residual -> LN                                      tcgen05 (k=1) + CUDA cores
//   units = argmin_j ||x - c_j||^2  (first index on ties)                              warp per frame, shuffle reduce
//
// All GEMMs use the split-fp16 scheme of conv_tc.cuh (fp32-accurate), so the layer-6 features agree with an fp32
// evaluation to ~1e-5 and the argmin changes only on near-ties (tests report the margin).
#include <cmath>
#include <cstring>
#include <cfloat>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "tc_host.cuh"

namespace dissc {

constexpr int kHubHalo = 64;  // plane halo rows of the encoder (pos_conv pads 64 frames)
constexpr int kConv0K = 10, kConv0S = 5, kConv0Chunk = 128;  // frames per CTA in the conv0 kernels

// ---- valid lengths of every extractor layer -----------------------------------------------------------------
// lens[l*B + b] = frames after conv layer l (l = 0..6) for clip b with n_samples[b] samples
__global__ void hub_lengths_kernel(const int* n_samples, int N, int B, int* lens) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int n = n_samples ? min(n_samples[b], N) : N;
  const int ks[7] = {10, 3, 3, 3, 3, 2, 2}, ss[7] = {5, 2, 2, 2, 2, 2, 2};
  for (int l = 0; l < 7; ++l) {
    n = n >= ks[l] ? (n - ks[l]) / ss[l] + 1 : 0;
    lens[l * B + b] = n;
  }
}

// Debug aid (DISSC_HUB_RANGE_CHECK=1): counts the fp16 "hi" values of a plane tensor that sit at the format's limit
// (|v| >= 65504: cvt.rn.satfinite clamped them) -- a real checkpoint with outlier activations beyond the fp16 range would
// otherwise lose them silently.  Rows [halo, halo + rows) of every slab.
__global__ void hub_range_check_kernel(const __half* hi, long long slabs, int Tp, int halo, int rows,
                                       unsigned long long* count) {
  const long long n = slabs * rows * 8;
  unsigned long long local = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long slab = i / ((long long)rows * 8), rem = i - slab * rows * 8;
    const unsigned short bits = reinterpret_cast<const unsigned short*>(hi)[(slab * Tp + halo) * 8 + rem];
    local += (bits & 0x7fffu) >= 0x7bffu;
  }
  if (local) atomicAdd(count, local);
}

// row_off[b] = sum of the frame counts of clips 0 .. b-1 (row_off[B] = total): where clip b starts in the packed row
// space of the transformer stack.  B <= a few thousand: one thread.
__global__ void hub_row_offsets_kernel(const int* lens, int B, int T, int* row_off) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) {
      row_off[b] = acc;
      acc += min(T, max(lens[b], 0));
    }
    row_off[B] = acc;
  }
}

// ---- conv0 + GroupNorm statistics ----------------------------------------------------------------------------
// GroupNorm(C groups, C channels) needs, per clip and channel, the mean and variance over time of y_c[t] = sum_k w_c[k]
// x[5t + k].  Both are functions of 65 MOMENTS OF THE WAVE that do not depend on the channel:
//     S[k]    = sum_t x[5t + k]                    (10)         mean_c   = w_c . S / n
//     R[k][j] = sum_t x[5t + k] x[5t + j], k <= j  (55)         E[y_c^2] = w_c^T R w_c / n
// so the statistics cost 65 multiply-adds per frame instead of a full 512-channel conv0 pass (0.19 ms -> 0.02 ms at
// 32 x 96 000 samples).  A filter that rejects most of the signal makes w^T R w a sum of large cancelling terms, so the
// moments are accumulated in fp64 (products of two fp32 numbers are exact there) and so is the quadratic form.
constexpr int kWaveMoments = 65;
// moment m -> (k, j): m < 10: S[m] (k = m, j = -1); else the pairs k <= j in row-major order
__device__ __forceinline__ void wave_moment_kj(int m, int& k, int& j) {
  if (m < kConv0K) { k = m; j = -1; return; }
  int r = m - kConv0K;
  k = 0;
  while (r >= kConv0K - k) { r -= kConv0K - k; ++k; }
  j = k + r;
}

// partial[(b*nchunk + chunk)*65 + m]: moment m over the chunk's valid frames.  Four threads per moment (frames mod 4).
__global__ void __launch_bounds__(256) hub_wave_moments_kernel(const float* wave, const int* lens0, int N, int nchunk,
                                                               double* partial) {
  __shared__ float sw[kConv0Chunk * kConv0S + kConv0K];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int T0 = lens0[b];
  const int t0 = chunk * kConv0Chunk;
  const int nf = max(0, min(kConv0Chunk, T0 - t0));
  const int ns = nf > 0 ? (nf - 1) * kConv0S + kConv0K : 0;
  for (int i = threadIdx.x; i < ns; i += blockDim.x) sw[i] = wave[(size_t)b * N + (size_t)t0 * kConv0S + i];
  __syncthreads();
  const int sub = threadIdx.x & 3;
  for (int m = threadIdx.x >> 2; m < 128; m += 64) {   // every thread runs both rounds: the shuffles need whole warps
    int k = 0, j = -1;
    if (m < kWaveMoments) wave_moment_kj(m, k, j);
    double acc = 0.0;
    if (m < kWaveMoments) {
      for (int f = sub; f < nf; f += 4) {
        const double xk = (double)sw[f * kConv0S + k];
        acc = j < 0 ? acc + xk : fma(xk, (double)sw[f * kConv0S + j], acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (sub == 0 && m < kWaveMoments) partial[((size_t)b * nchunk + chunk) * kWaveMoments + m] = acc;
  }
}

// scale/shift of GroupNorm(C groups, C channels): y = (x - mean) * rstd * gamma + beta  (biased variance, eps 1e-5).
// One CTA per clip: the chunk partials are summed into shared memory, then every thread takes channels.
__global__ void __launch_bounds__(256) hub_gn_finalize_kernel(const double* partial, const float* w0, const float* gamma,
                                                              const float* beta, const int* lens0, int C, int nchunk,
                                                              float2* scale_shift) {
  __shared__ double s_mom[kWaveMoments];
  __shared__ double s_R[kConv0K][kConv0K];
  const int b = blockIdx.x;
  const int sub = threadIdx.x & 3;
  for (int m = threadIdx.x >> 2; m < 128; m += 64) {
    double acc = 0.0;
    if (m < kWaveMoments)
      for (int k = sub; k < nchunk; k += 4) acc += partial[((size_t)b * nchunk + k) * kWaveMoments + m];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (sub == 0 && m < kWaveMoments) s_mom[m] = acc;
  }
  __syncthreads();
  if (threadIdx.x >= kConv0K && threadIdx.x < kWaveMoments) {
    int k, j;
    wave_moment_kj(threadIdx.x, k, j);
    s_R[k][j] = s_R[j][k] = s_mom[threadIdx.x];
  }
  __syncthreads();
  const double n = (double)max(lens0[b], 1);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double w[kConv0K];
#pragma unroll
    for (int k = 0; k < kConv0K; ++k) w[k] = (double)w0[c * kConv0K + k];
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int k = 0; k < kConv0K; ++k) {
      s = fma(w[k], s_mom[k], s);
      double row = 0.0;
#pragma unroll
      for (int j = 0; j < kConv0K; ++j) row = fma(w[j], s_R[k][j], row);
      q = fma(w[k], row, q);
    }
    const double mean = s / n;
    const double var = fmax(q / n - mean * mean, 0.0);
    const double rstd = 1.0 / sqrt(var + 1e-5);
    const float sc = (float)(rstd * gamma[c]);
    scale_shift[(size_t)b * C + c] = make_float2(sc, (float)(beta[c] - mean * rstd * gamma[c]));
  }
}

// GELU for the 314 M conv0 outputs of a 32-clip batch, where the erf is what the kernel issues most: x * Phi(x) with
// erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2), t = 1 / (1 + p z)  (Abramowitz & Stegun 7.1.26, |error|
// <= 1.5e-7) -- 16 instructions instead of the 26 of erff's branch-free piecewise polynomial, and the same absolute accuracy
// once either form is evaluated in fp32 (max |error| vs fp64 over [-8, 8]: 6.1e-7 this form, 6.8e-7 0.5 x (1 + erff)).
__device__ __forceinline__ float gelu_as(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float h = p * t * e;                 // erfc(|x| / sqrt 2) / 2
  return x * (x >= 0.f ? 1.f - h : h);
}

// conv0 recomputed, normalised, GELU'd and written as split planes, de-interleaved for the stride-2 conv1:
// out planes [B][2*C/8][Tp][8]: time step t -> slab (t&1)*C/8 + c8, row halo + (t>>1).  Frames >= T0 are zeros.
__global__ void __launch_bounds__(256) hub_conv0_apply_kernel(const float* wave, const float* w0, const float2* scale_shift,
                                                              const int* lens0, int N, int C, int Tp, int halo, __half* hi,
                                                              __half* lo) {
  extern __shared__ float smem[];
  float* sw = smem;                                        // wave segment
  float* sW = sw + kConv0Chunk * kConv0S + kConv0K + 2;    // [C][10]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int T0 = lens0[b];
  const int t0 = chunk * kConv0Chunk;
  const int nf = max(0, min(kConv0Chunk, T0 - t0));
  const int ns = nf > 0 ? (nf - 1) * kConv0S + kConv0K : 0;
  for (int i = threadIdx.x; i < ns; i += blockDim.x) sw[i] = wave[(size_t)b * N + (size_t)t0 * kConv0S + i];
  for (int i = threadIdx.x; i < C * kConv0K; i += blockDim.x) sW[i] = w0[i];
  __syncthreads();
  const int c8n = C / 8;
  // thread = (q, phase) of the chunk -- lanes run over q, so a warp writes 32 consecutive rows of one slab -- and walks
  // the channel groups: its frame, the ten samples under it and all address arithmetic are loop-invariant
  constexpr int qn = kConv0Chunk / 2;
  static_assert(256 % (2 * qn) == 0, "thread -> (q, phase) mapping");
  const int q = threadIdx.x % qn, ph = (threadIdx.x / qn) & 1;
  const int f = 2 * q + ph;  // frame inside the chunk
  const bool have = f < nf;
  float x[kConv0K];
#pragma unroll
  for (int j = 0; j < kConv0K; ++j) x[j] = have ? sw[f * kConv0S + j] : 0.f;
  const float2* ss = scale_shift + (size_t)b * C;
  for (int c8 = threadIdx.x / (2 * qn); c8 < c8n; c8 += 256 / (2 * qn)) {
    float v[8];
    if (have) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = c8 * 8 + e;
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < kConv0K; ++j) a = fmaf(x[j], sW[c * kConv0K + j], a);
        const float2 sc = __ldg(ss + c);
        v[e] = gelu_as(fmaf(a, sc.x, sc.y));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    const size_t off = (((size_t)b * 2 * c8n + (size_t)ph * c8n + c8) * Tp + halo + (t0 / 2) + q) * 8;
    split_store8(hi + off, lo + off, v);
  }
}

// ---- LayerNorm over channels, f32b in -> f32b and/or planes out ---------------------------------------------------
// 8 lanes per row (frame); each lane keeps its 8-channel groups in registers (C <= 768 -> <= 12 groups per lane).
// One CTA = 32 consecutive frames of one utterance x all channels: warp w owns the 8-channel groups w, w + 16, ...
// and lane = frame, so every global access of a warp is 32 x 32 contiguous bytes; per-frame sums go through shared
// memory (16 partials per frame).  G = C8 / 16 groups per thread (4 for 512 channels, 6 for 768).
// `row_off` != null: the output goes to the PACKED layout of the transformer stack -- one row space for the whole batch,
// clip b's valid frames at rows [row_off[b], row_off[b] + len_b) of a single slab set [1][C8][oTr | oTp][8] -- and only
// valid frames are stored.
template <int G>
__global__ void __launch_bounds__(512) hub_layernorm_kernel(const float* in, const float* gamma, const float* beta,
                                                            const int* lengths, int B, int C8, int T, int Tr, int oTr,
                                                            int oTp, int halo, const int* row_off, float* out_f,
                                                            __half* out_hi, __half* out_lo) {
  __shared__ float s_sum[16][32], s_sq[16][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (T + 31) / 32;
  const int b = blockIdx.x / tiles, t = (blockIdx.x - b * tiles) * 32 + lane;
  const bool live = t < T;
  const bool valid = live && t < (lengths ? min(T, lengths[b]) : T);
  float x[G][8];
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c8 = warp + g * 16;
    if (valid) {
      float4 a, c;
      ldg8(in + (((size_t)b * C8 + c8) * Tr + t) * 8, a, c);
      x[g][0] = a.x; x[g][1] = a.y; x[g][2] = a.z; x[g][3] = a.w;
      x[g][4] = c.x; x[g][5] = c.y; x[g][6] = c.z; x[g][7] = c.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[g][e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s += x[g][e];
  }
  s_sum[warp][lane] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int w = 0; w < 16; ++w) s += s_sum[w][lane];
  const float mean = s / (float)(C8 * 8);
  float q = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = x[g][e] - mean;
      q = fmaf(d, d, q);
    }
  }
  s_sq[warp][lane] = q;
  __syncthreads();
  q = 0.f;
#pragma unroll
  for (int w = 0; w < 16; ++w) q += s_sq[w][lane];
  const float rstd = rsqrtf(q / (float)(C8 * 8) + 1e-5f);
  if (!live || (row_off && !valid)) return;
  const size_t ob = row_off ? 0 : (size_t)b;                  // output batch slab / row
  const int orow = row_off ? row_off[b] + t : t;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c8 = warp + g * 16;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c8 * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c8 * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c8 * 8 + 4));
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = valid ? fmaf((x[g][e] - mean) * rstd, gm[e], bt[e]) : 0.f;
    if (out_f) stg8(out_f + ((ob * C8 + c8) * oTr + orow) * 8, y);
    if (out_hi) {
      const size_t off = ((ob * C8 + c8) * oTp + halo + orow) * 8;
      split_store8(out_hi + off, out_lo + off, y);
    }
  }
}

// ---- self-attention (fp32, online softmax) ------------------------------------------------------------------------
// qkv f32b [B][3*D/8][Tr][8] (q pre-scaled); out: planes [B][D/8][Tp][8].  Head size 64.
// kAttnQ query rows per CTA, keys / values streamed through shared memory in tiles of 32.  A LANE PAIR owns two query rows
// (t and t + 16 inside the warp's 32 rows) and half of the head dimension each (the first / second float4 of every
// 8-channel group), so one 16-byte shared-memory load of a key / value row feeds 8 FMAs (two queries x four dims) instead
// of 4: the one-thread-per-row version was bound by exactly that load (0.55 ms per layer at 32 clips; this one 4x the
// FMA rate).  Partial dot products are completed with one __shfl_xor per (query, key); the online softmax rescales the
// accumulators once per chunk of 8 keys.
constexpr int kAttnKT = 32;
constexpr int kAttnKC = 8;
constexpr int kAttnQ = 64;   // query rows (= threads) per CTA: 299 frames waste 6 % of the last tile instead of 22 % at 128
// Packed layout: qkv f32b [1][3*D8][Tr][8] and the output planes [1][D8][Tp][8] hold clip b's frames at rows row_off[b] + t.
__global__ void __launch_bounds__(kAttnQ, 6) hub_attention_kernel(const float* qkv, const int* lengths, const int* row_off,
                                                            int D8, int T, int Tr, int Tp, int halo, __half* out_hi,
                                                            __half* out_lo) {
  __shared__ __align__(16) float sK[2][kAttnKT][64];   // double-buffered: tile i+1 lands (cp.async) while tile i is used
  __shared__ __align__(16) float sV[2][kAttnKT][64];
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hf = lane & 1;                                   // which float4 of every 8-channel group
  const int tq[2] = {(int)blockIdx.x * kAttnQ + warp * 32 + (lane >> 1), (int)blockIdx.x * kAttnQ + warp * 32 + 16 + (lane >> 1)};
  const int Tv = lengths ? min(T, lengths[b]) : T;
  const size_t bq = 0;
  const int r0 = row_off[b];
  // stage one K / V tile: 32 keys x 64 dims each; element (key, c, e) <- f32b[(D8 + h*8 + c)][k0+key][e].  Consecutive
  // threads take consecutive float4 of a key row (conflict-free shared stores); keys past the valid length are zero-filled.
  auto stage = [&](int buf, int k0) {
    for (int i = threadIdx.x; i < kAttnKT * 16; i += kAttnQ) {
      const int c4 = i & 15, key = i >> 4;  // c4: 16 float4 per key row
      const int c = c4 >> 1, half = c4 & 1;
      const bool ok = k0 + key < Tv;
      const int kk = ok ? k0 + key : 0;
      const float* gk = qkv + ((bq + D8 + h * 8 + c) * Tr + r0 + kk) * 8 + half * 4;
      const float* gv = qkv + ((bq + 2 * D8 + h * 8 + c) * Tr + r0 + kk) * 8 + half * 4;
      const uint32_t n = ok ? 16u : 0u;     // src-size 0: the 16 destination bytes are written as zeros
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(&sK[buf][key][c4 * 4])), "l"(gk), "r"(n) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(&sV[buf][key][c4 * 4])), "l"(gv), "r"(n) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0, 0);
  float q[2][32], o[2][32];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 a = make_float4(0, 0, 0, 0);
      if (tq[r] < Tv) a = *reinterpret_cast<const float4*>(qkv + ((bq + h * 8 + c) * Tr + r0 + tq[r]) * 8 + hf * 4);
      q[r][c * 4 + 0] = a.x; q[r][c * 4 + 1] = a.y; q[r][c * 4 + 2] = a.z; q[r][c * 4 + 3] = a.w;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) o[r][i] = 0.f;
  }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  int buf = 0;
  for (int k0 = 0; k0 < Tv; k0 += kAttnKT, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // this thread's copies of tile k0 have landed
    __syncthreads();                                        // ... everyone's have, and everyone is done with tile k0 - 32
    if (k0 + kAttnKT < Tv) stage(buf ^ 1, k0 + kAttnKT);    // prefetch the next tile into the buffer just released
    const int nk = min(kAttnKT, Tv - k0);
#pragma unroll 1
    for (int j0 = 0; j0 < nk; j0 += kAttnKC) {
      float sc[2][kAttnKC];
#pragma unroll
      for (int jj = 0; jj < kAttnKC; ++jj) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 kk = *reinterpret_cast<const float4*>(&sK[buf][j0 + jj][c * 8 + hf * 4]);
          a0 = fmaf(q[0][c * 4], kk.x, a0); a0 = fmaf(q[0][c * 4 + 1], kk.y, a0);
          a0 = fmaf(q[0][c * 4 + 2], kk.z, a0); a0 = fmaf(q[0][c * 4 + 3], kk.w, a0);
          a1 = fmaf(q[1][c * 4], kk.x, a1); a1 = fmaf(q[1][c * 4 + 1], kk.y, a1);
          a1 = fmaf(q[1][c * 4 + 2], kk.z, a1); a1 = fmaf(q[1][c * 4 + 3], kk.w, a1);
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        const bool ok = j0 + jj < nk;
        sc[0][jj] = ok ? a0 : -INFINITY;
        sc[1][jj] = ok ? a1 : -INFINITY;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float mx = m[r];
#pragma unroll
        for (int jj = 0; jj < kAttnKC; ++jj) mx = fmaxf(mx, sc[r][jj]);
        const float alpha = (m[r] == -INFINITY) ? 0.f : expf(m[r] - mx);
        l[r] *= alpha;
#pragma unroll
        for (int i = 0; i < 32; ++i) o[r][i] *= alpha;
#pragma unroll
        for (int jj = 0; jj < kAttnKC; ++jj) {
          sc[r][jj] = (j0 + jj < nk) ? expf(sc[r][jj] - mx) : 0.f;   // both lanes of the pair hold the same value
          l[r] += sc[r][jj];
        }
        m[r] = mx;
      }
#pragma unroll
      for (int jj = 0; jj < kAttnKC; ++jj) {
        const float p0 = sc[0][jj], p1 = sc[1][jj];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 vv = *reinterpret_cast<const float4*>(&sV[buf][j0 + jj][c * 8 + hf * 4]);
          o[0][c * 4] = fmaf(p0, vv.x, o[0][c * 4]); o[0][c * 4 + 1] = fmaf(p0, vv.y, o[0][c * 4 + 1]);
          o[0][c * 4 + 2] = fmaf(p0, vv.z, o[0][c * 4 + 2]); o[0][c * 4 + 3] = fmaf(p0, vv.w, o[0][c * 4 + 3]);
          o[1][c * 4] = fmaf(p1, vv.x, o[1][c * 4]); o[1][c * 4 + 1] = fmaf(p1, vv.y, o[1][c * 4 + 1]);
          o[1][c * 4 + 2] = fmaf(p1, vv.z, o[1][c * 4 + 2]); o[1][c * 4 + 3] = fmaf(p1, vv.w, o[1][c * 4 + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int t = tq[r];
    if (t >= Tv) continue;     // packed rows: frames past the clip's length belong to the next clip
    const float inv = l[r] > 0.f ? 1.f / l[r] : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      // this lane's four channels of the 8-channel group: one 8-byte store per plane
      uint32_t hh[2], ll[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float y0 = o[r][c * 4 + 2 * i] * inv, y1 = o[r][c * 4 + 2 * i + 1] * inv;
        hh[i] = pack_half2_sat(y0, y1);
        const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hh[i]));
        ll[i] = pack_half2_sat(y0 - back.x, y1 - back.y);
      }
      const size_t off = (((size_t)h * 8 + c) * Tp + halo + r0 + t) * 8 + hf * 4;
      *reinterpret_cast<uint2*>(out_hi + off) = make_uint2(hh[0], hh[1]);
      *reinterpret_cast<uint2*>(out_lo + off) = make_uint2(ll[0], ll[1]);
    }
  }
}

// ---- attention on the tensor cores ---------------------------------------------------------------------------------
// One CTA = one (clip, head, 128-query tile); T <= kAttnTcKeys keys (a 160 000-sample clip has 499 frames: those
// batches keep the CUDA-core kernel above).  Inputs are the split-fp16 planes of the fused QKV projection
// [B][3*D8][Tp][8] (q already scaled by head_dim^-0.5), i.e. exactly the K-major UMMA operand layout for Q (A: rows =
// queries) and K (B: rows = keys), so both tiles arrive by bulk TMA:
//   S = Q K^T        128 x 320, three split-precision MMAs per 16 dims into ONE TMEM accumulator (12 accumulations)
//   softmax          thread r owns query row r = TMEM lane r: two passes over its 320 scores (max; exp + sum), keys
//                    >= the clip's valid length masked; P written as fp16 hi/lo K-major tiles over the dead K tile
//   O = P V          V transposed on chip ([keys][dims] planes -> K-major [key group][hi|lo][dim][8 keys]) while the
//                    S MMAs run; per 160-key block  P_hi x [V_hi|V_lo] -> [main|cross],  P_lo x V_hi -> cross
//   out              (main + cross) / sum -> fp16 hi/lo planes [B][D8][Tp][8] (the out-projection's operand)
// fp32-accurate like every other GEMM here (the units must not move), at tensor-core instead of FMA-pipe speed.
int g_hub_attn_tc = -1;               // dissc_tc_set_tuning key 4 / env DISSC_HUB_ATTN_TC (default 1)
constexpr int kAttnTcKeys = 320;      // keys per CTA (multiple of 16, two N = 160 chunks for S)
constexpr int kAttnTcThreads = 512;   // four threads per query row (TMEM lane): the kernel is latency-bound, warps are what it needs
constexpr uint32_t kVtStride = 2048 + 16;                    // bytes between key groups of the V^T tile (bank padding)
constexpr size_t kAttnTcSmem = 32768 + 81920 + (kAttnTcKeys / 8) * kVtStride + 64;   // Q | K (later P) | V^T | barriers

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kAttnTcThreads, 1) hub_attention_tc_kernel(const __half* q_hi, const __half* q_lo,
                                                                            const int* lengths, const int* row_off, int D8,
                                                                            int T, int Tp_in, int Tp, int halo,
                                                                            __half* out_hi, __half* out_lo) {
  constexpr int NK = kAttnTcKeys, NKB = NK / 2;          // keys, keys per P block
  constexpr uint32_t kQPlane = 8 * 128 * 16;             // one plane of the Q tile: [c8][128 rows][16 B]
  constexpr uint32_t kKPlane = 8 * NK * 16;              // K tile: [c8][320 rows][16 B]
  constexpr uint32_t kPPlane = (NKB / 8) * 128 * 16;     // P block: [key group (20)][128 rows][16 B]
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sQ = smem_raw;                          // hi plane, lo plane
  unsigned char* sK = sQ + 2 * kQPlane;                  // hi plane, lo plane; reused for P (hi block, lo block)
  unsigned char* sV = sK + 2 * kKPlane;                  // V^T: [key group (40), stride kVtStride][hi|lo][64 dims][16 B = 8 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + (NK / 8) * kVtStride);
  uint64_t* qk_full = bars;      // TMA: Q and K tiles landed
  uint64_t* mma_done = bars + 1; // tcgen05.commit after S, after each PV block
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_red[4][128];   // per-row partial max / sum of the four threads that share a row
  // four threads per query row: threads r, r + 128, r + 256, r + 384 (warps w, w + 4, w + 8, w + 12 address the same TMEM
  // lane quarter) take every fourth 16-key chunk of every pass and a quarter of the output each
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, hsel = tid >> 7;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int Tv = lengths ? min(T, lengths[b]) : T;
  if (tid == 0) {
    mbar_init(qk_full, 1);
    mbar_init(mma_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const size_t slab = (size_t)Tp_in * 8;                                   // halves per (b, c8) slab of the QKV planes
  // packed layout: one slab set for the batch, clip b's frames at rows row_off[b] + t (a tile may read on into the next
  // clip's rows or the slack behind the last one: those keys are masked, those query rows are not stored)
  const int r0 = row_off[b];
  const size_t base_q = ((size_t)h * 8) * slab + (size_t)(halo + r0) * 8;
  const size_t base_k = base_q + (size_t)D8 * slab, base_v = base_q + (size_t)2 * D8 * slab;
  constexpr uint32_t idesc_s = (1u << 4) | ((uint32_t)((NK / 2) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // N = 160
  constexpr uint32_t idesc_o2 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);       // N = 128
  constexpr uint32_t idesc_o1 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);        // N = 64
  if (warp == 0 && elect_one()) {
    // ---- Q / K tiles by bulk TMA, then S = Q K^T (hi*hi + hi*lo + lo*hi into one accumulator, TMEM columns [0, 320))
    mbar_arrive_expect_tx(qk_full, 2 * kQPlane + 2 * kKPlane);
    for (int c8 = 0; c8 < 8; ++c8) {
      tma_load_1d(sQ + c8 * 2048, q_hi + base_q + c8 * slab + (size_t)q0 * 8, 2048, qk_full);
      tma_load_1d(sQ + kQPlane + c8 * 2048, q_lo + base_q + c8 * slab + (size_t)q0 * 8, 2048, qk_full);
      tma_load_1d(sK + c8 * (NK * 16), q_hi + base_k + c8 * slab, NK * 16, qk_full);
      tma_load_1d(sK + kKPlane + c8 * (NK * 16), q_lo + base_k + c8 * slab, NK * 16, qk_full);
    }
    mbar_wait(qk_full, 0);
    tc_fence_after();
    const uint32_t qd = umma_desc_lo(smem_u32(sQ), 2048), kd = umma_desc_lo(smem_u32(sK), NK * 16);
    const uint32_t q_lo_off = kQPlane >> 4, k_lo_off = kKPlane >> 4;
    for (int nch = 0; nch < 2; ++nch) {
      const uint32_t d = tmem_base + nch * (NK / 2);
      const uint32_t kdn = kd + ((uint32_t)(nch * (NK / 2) * 16) >> 4);
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t qa = qd + ks * ((2 * 2048) >> 4), ka = kdn + ks * ((2 * NK * 16) >> 4);
        umma_f16(d, umma_desc(qa), umma_desc(ka), idesc_s, ks > 0);
        umma_f16(d, umma_desc(qa), umma_desc(ka + k_lo_off), idesc_s, 1);
        umma_f16(d, umma_desc(qa + q_lo_off), umma_desc(ka), idesc_s, 1);
      }
    }
    umma_commit(mma_done);
  }
  __syncwarp();
  // ---- meanwhile: V^T.  A thread transposes one 8-key x 8-dim block per plane in registers: eight 16-byte global
  //      loads (8 consecutive keys of one 8-dim slab: 128 contiguous bytes, lanes = consecutive key groups), 32 byte
  //      permutes, eight 16-byte shared stores (one per dim: its 8 keys).  The key-group stride of the V^T tile is padded
  //      to 2048 + 16 bytes so the 8 lanes of a store phase hit 8 different 16-byte bank groups.
  //      Warp 0 issues the TMA loads and the MMAs, so the items go to the other warps (the row-max pass below starts with a
  //      CTA barrier: a late warp 0 would hold everybody up), and a thread has the loads of BOTH planes in flight before it
  //      permutes the first -- L2 latency is what this phase costs.
  for (int item = tid - 32; item >= 0 && item < (NK / 8) * 8; item += kAttnTcThreads - 32) {
    const int c8 = item / (NK / 8), kg = item - c8 * (NK / 8);
    uint4 a[2][8];
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      const __half* src = (pl ? q_lo : q_hi) + base_v + c8 * slab + (size_t)kg * 64;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        a[pl][kk] = (kg * 8 + kk < Tv) ? *reinterpret_cast<const uint4*>(src + kk * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      unsigned char* dst = sV + (size_t)kg * kVtStride + pl * 1024 + (size_t)(c8 * 8) * 16;
#pragma unroll
      for (int n = 0; n < 8; ++n) {       // dim n of the slab: its 8 keys, two per word
        const uint32_t sel = (n & 1) ? 0x7632u : 0x5410u;
        uint32_t o[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint4 &e = a[pl][2 * w], &d = a[pl][2 * w + 1];
          const uint32_t lo_w = (n >> 1) == 0 ? e.x : (n >> 1) == 1 ? e.y : (n >> 1) == 2 ? e.z : e.w;
          const uint32_t hi_w = (n >> 1) == 0 ? d.x : (n >> 1) == 1 ? d.y : (n >> 1) == 2 ? d.z : d.w;
          o[w] = __byte_perm(lo_w, hi_w, sel);
        }
        *reinterpret_cast<uint4*>(dst + n * 16) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  // ---- softmax: query row = TMEM lane; pass 1: row maximum over the valid keys
  mbar_wait(mma_done, 0);
  __syncwarp();               // tcgen05.ld is warp-collective
  tc_fence_after();
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  float mx = -INFINITY;
  for (int c0 = hsel * 16; c0 < NK; c0 += 64) {
    if (c0 >= Tv) break;
    float sc[16];
    tmem_ld16(lane_addr + c0, sc);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < Tv) mx = fmaxf(mx, sc[i]);
  }
  s_red[hsel][row] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0][row], s_red[1][row]), fmaxf(s_red[2][row], s_red[3][row]));
  constexpr float kLog2e = 1.4426950408889634f;
  const float mxl = mx * kLog2e;
  float l = 0.f;
  uint32_t done_phase = 1;
  const uint32_t pd = umma_desc_lo(smem_u32(sK), 128 * 16), vd = umma_desc_lo(smem_u32(sV), kVtStride);
  for (int blk = 0; blk < 2; ++blk) {
    // P block = keys [blk*160, blk*160 + 160): exp(s - max), fp16 hi / lo, K-major [key group][row][8 keys]
    for (int ci = hsel; ci < NKB / 16; ci += 4) {
      const int c0 = blk * NKB + ci * 16;
      float sc[16];
      if (c0 < Tv) {
        tmem_ld16(lane_addr + c0, sc);
        tmem_ld_wait();
      }
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        float pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int key = c0 + g8 * 8 + i;
          // exp(s - max) as one FFMA + ex2.approx (relative error 2^-22, far inside the fp16 hi / lo split of P)
          pv[i] = (key < Tv) ? exp2f(fmaf(sc[g8 * 8 + i], kLog2e, -mxl)) : 0.f;
          l += pv[i];
        }
        unsigned char* dst = sK + (size_t)(ci * 2 + g8) * 2048 + row * 16;
        split_store8(reinterpret_cast<__half*>(dst), reinterpret_cast<__half*>(dst + kPPlane), pv);
      }
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      const uint32_t d = tmem_base + NK;   // O accumulator: [main 64 | cross 64]
      for (int ks = 0; ks < NKB / 16; ++ks) {
        const uint32_t pa = pd + ks * ((2 * 2048) >> 4);
        const uint32_t va = vd + (uint32_t)(blk * (NKB / 8) + 2 * ks) * (kVtStride >> 4);
        umma_f16(d, umma_desc(pa), umma_desc(va), idesc_o2, (blk | ks) != 0);                 // P_hi x [V_hi|V_lo]
        umma_f16(d + 64, umma_desc(pa + (kPPlane >> 4)), umma_desc(va), idesc_o1, 1);          // P_lo x V_hi -> cross
      }
      umma_commit(mma_done);
    }
    __syncwarp();
    mbar_wait(mma_done, done_phase);   // the P buffer may be overwritten / O is complete
    done_phase ^= 1;
    __syncwarp();
    tc_fence_after();
  }
  s_red[hsel][row] = l;
  __syncthreads();
  l = (s_red[0][row] + s_red[1][row]) + (s_red[2][row] + s_red[3][row]);
  // ---- O / l -> planes (this thread's 16 of the row's 64 dims)
  {
    const int t = q0 + row;
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const uint32_t o_addr = lane_addr + NK + hsel * 16;
    float om[16], oc[16];
    tmem_ld16(o_addr, om);
    tmem_ld16(o_addr + 64, oc);
    tmem_ld_wait();
    if (t < Tv) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = (om[c * 8 + e] + oc[c * 8 + e]) * inv;
        const size_t off = (((size_t)h * 8 + hsel * 2 + c) * Tp + halo + r0 + t) * 8;
        split_store8(out_hi + off, out_lo + off, y);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- k-means assignment ----------------------------------------------------------------------------------------------
// one warp per frame; x f32b [B][D8][Tr][8]; centroids (K, D) row-major; units int64 (B, T) (-1 past the valid length)
// A warp assigns kKmFrames consecutive frames at once: every centroid value it loads from L2 is used for all of them
// (the one-frame-per-warp version pulled the whole 300 KB codebook through L2 for every frame: 0.32 ms per 32 clips).
// The per-frame arithmetic and its order are unchanged -- 8 channels per lane and group, groups in order, xor-shuffle
// tree, strict '<' so the first index wins ties -- so the units are bit-identical to the previous kernel's.
constexpr int kKmFrames = 4;
// x: packed f32b [1][D8][Tr][8], clip b's frame t at row row_off[b] + t
__global__ void __launch_bounds__(256) hub_kmeans_f32b_kernel(const float* x, const float* cent, const int* lengths,
                                                              const int* row_off, int B, int D8, int T, int Tr, int K,
                                                              long long* units, float* feat_out /* (B,T,D) or null */) {
  const long long w0 = ((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5) * kKmFrames;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * T;
  if (w0 >= total) return;
  float v[kKmFrames][4][8];  // D8 <= 128
  bool valid[kKmFrames];
#pragma unroll
  for (int f = 0; f < kKmFrames; ++f) {
    const long long wid = w0 + f;
    valid[f] = false;
    if (wid < total) {
      const int b = (int)(wid / T), t = (int)(wid - (long long)b * T);
      valid[f] = t < (lengths ? min(T, lengths[b]) : T);
      if (!valid[f]) {
        if (lane == 0) units[wid] = -1;
        if (feat_out)
          for (int i = lane; i < D8 * 8; i += 32) feat_out[wid * D8 * 8 + i] = 0.f;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c8 = lane + g * 32;
        float4 a = make_float4(0, 0, 0, 0), c = a;
        if (valid[f] && c8 < D8) {
          const float* p = x + ((size_t)c8 * Tr + row_off[b] + t) * 8;
          a = *reinterpret_cast<const float4*>(p);
          c = *reinterpret_cast<const float4*>(p + 4);
          if (feat_out) {
            float* o = feat_out + wid * D8 * 8 + c8 * 8;
            *reinterpret_cast<float4*>(o) = a;
            *reinterpret_cast<float4*>(o + 4) = c;
          }
        }
        v[f][g][0] = a.x; v[f][g][1] = a.y; v[f][g][2] = a.z; v[f][g][3] = a.w;
        v[f][g][4] = c.x; v[f][g][5] = c.y; v[f][g][6] = c.z; v[f][g][7] = c.w;
      }
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[f][g][e] = 0.f;
    }
  }
  float best[kKmFrames];
  int besti[kKmFrames];
#pragma unroll
  for (int f = 0; f < kKmFrames; ++f) { best[f] = INFINITY; besti[f] = 0; }
  for (int j = 0; j < K; ++j) {
    float d[kKmFrames];
#pragma unroll
    for (int f = 0; f < kKmFrames; ++f) d[f] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c8 = lane + g * 32;
      if (c8 < D8) {
        const float* c = cent + (size_t)j * D8 * 8 + c8 * 8;
        const float4 a = __ldg(reinterpret_cast<const float4*>(c)), e = __ldg(reinterpret_cast<const float4*>(c + 4));
#pragma unroll
        for (int f = 0; f < kKmFrames; ++f) {
          float u;
          u = v[f][g][0] - a.x; d[f] = fmaf(u, u, d[f]);
          u = v[f][g][1] - a.y; d[f] = fmaf(u, u, d[f]);
          u = v[f][g][2] - a.z; d[f] = fmaf(u, u, d[f]);
          u = v[f][g][3] - a.w; d[f] = fmaf(u, u, d[f]);
          u = v[f][g][4] - e.x; d[f] = fmaf(u, u, d[f]);
          u = v[f][g][5] - e.y; d[f] = fmaf(u, u, d[f]);
          u = v[f][g][6] - e.z; d[f] = fmaf(u, u, d[f]);
          u = v[f][g][7] - e.w; d[f] = fmaf(u, u, d[f]);
        }
      }
    }
#pragma unroll
    for (int f = 0; f < kKmFrames; ++f) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) d[f] += __shfl_xor_sync(0xffffffffu, d[f], o);
      if (d[f] < best[f]) {  // strict: first index wins ties (argmin semantics)
        best[f] = d[f];
        besti[f] = j;
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int f = 0; f < kKmFrames; ++f)
      if (valid[f]) units[w0 + f] = besti[f];
  }
}


// K <= 128 centroids: the assignment is a GEMM.  argmin_j |x - c_j|^2 = argmin_j (|c_j|^2 - 2 x.c_j) -- the form
// fairseq's ApplyKmeans and sklearn's predict evaluate -- so the centroids are one more streamed-weight layer (W = -2 C,
// bias = |c|^2 in fp64, +FLT_MAX for the padded columns) fed by the planes the layer-6 LayerNorm writes anyway: 75 row
// tiles x 3 split-fp16 MMAs instead of 1.9 G CUDA-core multiply-adds (0.18 -> 0.02 ms per 32 clips).  This kernel then
// takes the row minimum: dist f32b [1][16][Tr][8] (packed rows), lowest index on ties, -1 past the valid length.
constexpr int kKmMaxK = 128;
__global__ void __launch_bounds__(128) hub_kmeans_argmin_kernel(const float* dist, const int* lengths, const int* row_off,
                                                                int T, int Tr, long long* units) {
  const int tiles_per_b = (T + 127) / 128;
  const int b = blockIdx.x / tiles_per_b, t = (blockIdx.x - b * tiles_per_b) * 128 + threadIdx.x;
  if (t >= T) return;
  const int Tv = lengths ? min(T, lengths[b]) : T;
  if (t >= Tv) {
    units[(size_t)b * T + t] = -1;
    return;
  }
  const size_t row = (size_t)row_off[b] + t;
  float best = INFINITY;
  int besti = 0;
#pragma unroll 4
  for (int g8 = 0; g8 < kKmMaxK / 8; ++g8) {
    float4 a, c;
    ldg8(dist + ((size_t)g8 * Tr + row) * 8, a, c);
    const float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (v[i] < best) {   // ascending index, strict '<': the lowest index wins ties
        best = v[i];
        besti = g8 * 8 + i;
      }
  }
  units[(size_t)b * T + t] = besti;
}

// dense features out: packed f32b [1][D8][Tr][8] -> (B, T, D) row-major, zeros past the valid length
__global__ void __launch_bounds__(256) hub_features_out_kernel(const float* x, const int* lengths, const int* row_off, int D8,
                                                               int T, int Tr, float* feat_out) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int Tv = lengths ? min(T, lengths[b]) : T;
  const size_t row = (size_t)row_off[b] + t;
  for (int i = threadIdx.x; i < D8 * 2; i += blockDim.x) {
    const int c8 = i >> 1, half = i & 1;
    float4 v = make_float4(0, 0, 0, 0);
    if (t < Tv) v = *reinterpret_cast<const float4*>(x + ((size_t)c8 * Tr + row) * 8 + half * 4);
    *reinterpret_cast<float4*>(feat_out + ((size_t)b * T + t) * (D8 * 8) + c8 * 8 + half * 4) = v;
  }
}

// standalone: x (M, D) row-major
__global__ void __launch_bounds__(256) kmeans_rowmajor_kernel(const float* x, const float* cent, int M, int D, int K,
                                                              long long* out) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= M) return;
  float best = INFINITY;
  int besti = 0;
  for (int j = 0; j < K; ++j) {
    float d = 0.f;
    for (int i = lane; i < D; i += 32) {
      const float u = x[wid * D + i] - __ldg(cent + (size_t)j * D + i);
      d = fmaf(u, u, d);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (d < best) {
      best = d;
      besti = j;
    }
  }
  if (lane == 0) out[wid] = besti;
}

}  // namespace dissc

using namespace dissc;

struct HubLayer {
  TcLayer qkv, out, fc1, fc2;
  float *qkv_b = nullptr, *out_b = nullptr, *fc1_b = nullptr, *fc2_b = nullptr;
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
};

struct dissc_hubert {
  dissc_hubert_cfg cfg;
  int device = 0;
  std::vector<void*> allocs;
  float *w0 = nullptr, *gn_w = nullptr, *gn_b = nullptr;
  TcLayer conv[6];
  float *ln0_w = nullptr, *ln0_b = nullptr;
  TcLayer proj, pos;
  float *proj_b = nullptr, *pos_b = nullptr, *eln_w = nullptr, *eln_b = nullptr;
  std::vector<HubLayer> layers;
  float* cent = nullptr;
  TcLayer km;               // the k-means codebook as a GEMM layer (K <= 128): W = -2 C, bias = |c|^2
  float* km_b = nullptr;
  bool km_gemm = false;
};

namespace dissc {

struct HubWeights {
  std::map<std::string, const dissc_tensor*> m;
  const dissc_tensor* get(const std::string& k) const {
    auto it = m.find(k);
    return it == m.end() ? nullptr : it->second;
  }
};

static int hub_upload_f(dissc_hubert* g, const float* host, size_t n, float** out) {
  float* d = nullptr;
  DISSC_CUDA(cudaMalloc(&d, std::max<size_t>(n, 4) * sizeof(float)));
  g->allocs.push_back(d);
  DISSC_CUDA(cudaMemcpy(d, host, n * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return DISSC_OK;
}
static int hub_upload_h(dissc_hubert* g, const std::vector<__half>& packed, __half** out) {
  __half* d = nullptr;
  DISSC_CUDA(cudaMalloc(&d, packed.size() * sizeof(__half)));
  g->allocs.push_back(d);
  DISSC_CUDA(cudaMemcpy(d, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
  *out = d;
  return DISSC_OK;
}
static int hub_vec(dissc_hubert* g, const HubWeights& wm, const std::string& name, int n, float** out) {
  const dissc_tensor* t = wm.get(name);
  DISSC_CHECK(t && t->numel == n, DISSC_EMISSING, "missing tensor %s (%d elements)", name.c_str(), n);
  return hub_upload_f(g, t->data, n, out);
}
// Linear (Cout, Cin) [+ bias] as a k=1 tensor-core conv
// 256-column chunks with ONE accumulator for the encoder's GEMMs (DISSC_HUB_NC256=0: 128-column chunks, main + cross).
// These layers have one or two taps, so an activation tile is used for few MMAs and the operand stream from L2 -- not the
// tensor pipe -- is what bounds them: 148 SMs sustain ~35 B/clk each, a 128 x 128 split-fp16 tile needs 64 at full MMA rate
// (with the 2-CTA multicast weight stream), a 128 x 256 tile 43.
static int hub_nc256(int ncols) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DISSC_HUB_NC256");
    on = e ? (atoi(e) != 0) : 1;
  }
  return (on && ncols % 256 == 0) ? 256 : 0;
}
static bool hub_plan(int Cin, int ncols, int taps, TcLayer* L) {
  const int nc = hub_nc256(ncols);
  static int pair = -1;   // DISSC_HUB_PAIR2=1: CTA pairs with 256-row cta_group::2 MMAs (conv_tc.cuh)
  if (pair < 0) {
    const char* e = getenv("DISSC_HUB_PAIR2");
    pair = e ? (atoi(e) != 0) : 0;
  }
  if (nc && pair && tc_plan(Cin, ncols, taps, 1, 0, L, kHubHalo, nc, 1, 1)) return true;
  return tc_plan(Cin, ncols, taps, 1, 0, L, kHubHalo, nc, nc ? 1 : -1);
}

static int hub_linear(dissc_hubert* g, const HubWeights& wm, const std::string& name, int Cin, int Cout, TcLayer* L,
                      float** bias) {
  const dissc_tensor* w = wm.get(name + ".weight");
  DISSC_CHECK(w && w->numel == (int64_t)Cin * Cout, DISSC_EMISSING, "missing tensor %s.weight (%d,%d)", name.c_str(), Cout,
              Cin);
  DISSC_CHECK(hub_plan(Cin, Cout, 1, L), DISSC_EUNSUPPORTED, "%s: no tcgen05 plan for %d->%d", name.c_str(),
              Cin, Cout);
  L->Cout = Cout;
  const float* wd = w->data;
  auto packed = pack_weights_tc(*L, [=](int n, int ci, int) { return wd[(size_t)n * Cin + ci]; }, &L->inv_scale);
  int rc = hub_upload_h(g, packed, &L->w);
  if (rc) return rc;
  return hub_vec(g, wm, name + ".bias", Cout, bias);
}

struct Bump {
  char* base;
  size_t off = 0;
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  }
};

static int conv_out_len(int n, int k, int s) { return n >= k ? (n - k) / s + 1 : 0; }

struct HubShapes {
  int T[7];   // frames after conv layer l for the longest clip
  int Tq[6];  // rows of the de-interleaved planes feeding conv l+1: ceil(T[l]/2)
};
static HubShapes hub_shapes(int N) {
  HubShapes s;
  const int ks[7] = {10, 3, 3, 3, 3, 2, 2}, ss[7] = {5, 2, 2, 2, 2, 2, 2};
  int n = N;
  for (int l = 0; l < 7; ++l) {
    n = conv_out_len(n, ks[l], ss[l]);
    s.T[l] = n;
    if (l < 6) s.Tq[l] = (n + 1) / 2;
  }
  return s;
}
static size_t ru(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Packed row space of the transformer stack: after pos_conv no op looks across frames except attention, which is per
// clip anyway, so the B clips' VALID frames are laid end to end in one [1][C/8][rows][8] slab set and every GEMM sees
// M = sum of the frame counts instead of B x roundup(T, 128): 75 row tiles instead of 96 for 32 clips of 299 frames.
// R: upper bound of the used rows; Rr: rows of the f32b slabs; Rp: rows of the plane slabs -- halo, then Rr, then slack
// for the attention tiles that start at a clip's first row and run 320 keys / 384 query rows on whatever follows.
struct HubPacked {
  int R, Rr, Rp;
};
static HubPacked hub_packed_rows(int B, int T) {
  HubPacked k;
  k.R = B * T;
  k.Rr = (int)ru(k.R, 128);
  k.Rp = k.Rr + 2 * kHubHalo + 384;
  return k;
}

struct HubBuffers {
  int* lens;
  int* row_off;   // B + 1: first packed row of every clip, then the total
  unsigned long long* sat;   // DISSC_HUB_RANGE_CHECK: saturated fp16 activations seen in this forward
  double* gn_partial;   // [B][nchunk][65] wave moments
  float2* gn_ss;
  __half* dA[2];  // de-interleaved plane pairs (hi at [0], lo at hi + plane_elems)
  __half* dB[2];
  float *X6, *X7, *X8, *H, *Y, *QKV;
  __half *P6[2], *P7[2], *PH[2], *PA[2], *PF[2], *PQ[2];   // PQ: planes of the fused QKV projection (tensor-core attention)
  size_t total;
};

static HubBuffers hub_layout(const dissc_hubert* g, int B, int N, void* ws) {
  const dissc_hubert_cfg& c = g->cfg;
  const HubShapes s = hub_shapes(N);
  Bump bp{static_cast<char*>(ws)};
  HubBuffers b{};
  const int C = c.conv_dim, D = c.embed_dim;
  const int nchunk = (std::max(s.T[0], 1) + kConv0Chunk - 1) / kConv0Chunk;
  b.lens = (int*)bp.take((size_t)7 * B * sizeof(int));
  b.row_off = (int*)bp.take((size_t)(B + 1) * sizeof(int));
  b.sat = (unsigned long long*)bp.take(sizeof(unsigned long long));
  b.gn_partial = (double*)bp.take((size_t)B * nchunk * kWaveMoments * sizeof(double));
  b.gn_ss = (float2*)bp.take((size_t)B * C * sizeof(float2));
  auto plane_bytes = [&](int ch, int rows) { return (size_t)B * (ch / 8) * (ru(std::max(rows, 1), 128) + 2 * kHubHalo) * 16 + 4096; };
  // ping-pong: dA holds D0, D2, D4; dB holds D1, D3, D5
  const size_t a_bytes = plane_bytes(2 * C, s.Tq[0]), b_bytes = plane_bytes(2 * C, s.Tq[1]);
  b.dA[0] = (__half*)bp.take(a_bytes); b.dA[1] = (__half*)bp.take(a_bytes);
  b.dB[0] = (__half*)bp.take(b_bytes); b.dB[1] = (__half*)bp.take(b_bytes);
  const int T = std::max(s.T[6], 1);
  const size_t Tr = ru(T, 128);
  auto f32b_bytes = [&](int ch) { return (size_t)B * (ch / 8) * Tr * 32 + 4096; };
  // the transformer stack runs on ONE packed row space (hub_packed_rows): its plane buffers hold whichever is larger
  const HubPacked pk = hub_packed_rows(B, T);
  auto tplane_bytes = [&](int ch) { return std::max(plane_bytes(ch, T), (size_t)(ch / 8) * pk.Rp * 16 + 4096); };
  b.X6 = (float*)bp.take(f32b_bytes(C));
  b.X7 = (float*)bp.take(f32b_bytes(D));
  b.X8 = (float*)bp.take(f32b_bytes(D));
  b.H = (float*)bp.take(f32b_bytes(D));
  b.Y = (float*)bp.take(f32b_bytes(D));
  b.QKV = (float*)bp.take(f32b_bytes(3 * D));
  for (int i = 0; i < 2; ++i) {
    b.P6[i] = (__half*)bp.take(plane_bytes(C, T));
    b.P7[i] = (__half*)bp.take(plane_bytes(D, T));
    b.PH[i] = (__half*)bp.take(tplane_bytes(D));
    b.PA[i] = (__half*)bp.take(tplane_bytes(D));
    b.PF[i] = (__half*)bp.take(tplane_bytes(c.ffn_dim));
    b.PQ[i] = (__half*)bp.take(tplane_bytes(3 * D));
  }
  b.total = bp.off;
  return b;
}

// row_off == null: per-clip layout in and out (oTr = Tr, oTp = Tp); else packed output (see the kernel)
static int hub_layernorm(const float* in, const float* gw, const float* gb, const int* lengths, int B, int C, int T, int Tr,
                         int Tp, float* out_f, __half* out_hi, __half* out_lo, cudaStream_t st, const int* row_off = nullptr,
                         int oTr = 0, int oTp = 0) {
  const int blocks = B * ((T + 31) / 32);
  if (!row_off) { oTr = Tr; oTp = Tp; }
#define HUB_LN(G) hub_layernorm_kernel<G><<<blocks, 512, 0, st>>>(in, gw, gb, lengths, B, C / 8, T, Tr, oTr, oTp, kHubHalo, row_off, out_f, out_hi, out_lo)
  switch (C) {
    case 256: HUB_LN(2); break;
    case 512: HUB_LN(4); break;
    case 768: HUB_LN(6); break;
    case 1024: HUB_LN(8); break;
    default: return set_err(DISSC_EUNSUPPORTED, "layer norm over %d channels (256 / 512 / 768 / 1024 supported)", C);
  }
#undef HUB_LN
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

}  // namespace dissc

#define HUB_TRY(expr)                \
  do {                               \
    int _rc = (expr);                \
    if (_rc != DISSC_OK) return _rc; \
  } while (0)

extern "C" {

int dissc_hubert_num_frames(int n_samples) { return hub_shapes(n_samples).T[6]; }

int dissc_hubert_create(dissc_hubert_t** out, const dissc_hubert_cfg* cfg, const dissc_tensor* weights, int n_weights,
                        int device) {
  DISSC_CHECK(out && cfg && weights, DISSC_EINVAL, "null argument");
  *out = nullptr;
  const dissc_hubert_cfg& c = *cfg;
  DISSC_CHECK(c.conv_dim == 512 && c.embed_dim % 256 == 0 && c.embed_dim <= 1024 && c.ffn_dim % 256 == 0 &&
                  c.n_heads * 64 == c.embed_dim && c.n_layers >= 0 && c.pos_groups > 0 &&
                  c.embed_dim % c.pos_groups == 0 && (c.embed_dim / c.pos_groups) % 16 == 0 &&
                  c.embed_dim / c.pos_groups <= 64 && c.pos_kernel % 2 == 0 && c.pos_kernel / 2 <= kHubHalo &&
                  c.n_clusters > 0,
              DISSC_EUNSUPPORTED,
              "unsupported HuBERT geometry (need conv_dim 512, head size 64, embed/ffn multiples of 256, pos_conv group "
              "width a multiple of 16 and <= 64, even pos kernel <= 128)");
  DISSC_CUDA(cudaSetDevice(device));
  HubWeights wm;
  for (int i = 0; i < n_weights; ++i) wm.m[weights[i].name] = &weights[i];
  dissc_hubert* g = new dissc_hubert();
  g->cfg = c;
  g->device = device;
  auto fail = [&](int rc) {
    dissc_hubert_destroy(g);
    return rc;
  };
  int rc;
  const int C = c.conv_dim, D = c.embed_dim;
  const std::string fe = "feature_extractor.conv_layers.";
  if ((rc = hub_vec(g, wm, fe + "0.0.weight", C * kConv0K, &g->w0))) return fail(rc);
  if ((rc = hub_vec(g, wm, fe + "0.2.weight", C, &g->gn_w))) return fail(rc);
  if ((rc = hub_vec(g, wm, fe + "0.2.bias", C, &g->gn_b))) return fail(rc);
  for (int l = 1; l <= 6; ++l) {
    const int k = l <= 4 ? 3 : 2;
    const dissc_tensor* w = wm.get(fe + std::to_string(l) + ".0.weight");
    if (!w || w->numel != (int64_t)C * C * k) return fail(set_err(DISSC_EMISSING, "missing %s%d.0.weight (%d,%d,%d)", fe.c_str(), l, C, C, k));
    TcLayer* L = &g->conv[l - 1];
    const int taps = (k + 1) / 2;
    if (!hub_plan(2 * C, C, taps, L)) return fail(set_err(DISSC_EUNSUPPORTED, "no tcgen05 plan for extractor conv %d", l));
    L->Cout = C;
    if (k & 1) {
      // odd kernel: the last tap of the odd phase (jj = k) does not exist -> the second half of the channel blocks has
      // one tap less; the MMA thread skips those all-zero steps (1/4 of the layer's MMAs for k = 3)
      L->cb_split = L->n_cb / 2;
      L->k_hi = taps - 1;
    }
    const float* wd = w->data;
    // frame form of a stride-2 conv: channel ci' = phase*C + ci of frame q is x[ci, 2q + phase]; tap j' reads frame q + j'
    auto packed = pack_weights_tc(*L, [=](int n, int cip, int jp) {
      const int phase = cip / C, ci = cip % C;
      const int jj = 2 * jp + phase;
      return jj < k ? wd[((size_t)n * C + ci) * k + jj] : 0.f;
    }, &L->inv_scale);
    if ((rc = hub_upload_h(g, packed, &L->w))) return fail(rc);
  }
  if ((rc = hub_vec(g, wm, "layer_norm.weight", C, &g->ln0_w))) return fail(rc);
  if ((rc = hub_vec(g, wm, "layer_norm.bias", C, &g->ln0_b))) return fail(rc);
  if ((rc = hub_linear(g, wm, "post_extract_proj", C, D, &g->proj, &g->proj_b))) return fail(rc);
  {
    // pos_conv: (D, D/groups, k) folded weight; grouped implicit GEMM, one N-chunk per group
    const int gw = D / c.pos_groups, k = c.pos_kernel;
    const dissc_tensor* w = wm.get("encoder.pos_conv.0.weight");
    if (!w || w->numel != (int64_t)D * gw * k) return fail(set_err(DISSC_EMISSING, "missing encoder.pos_conv.0.weight (%d,%d,%d) [weight-norm folded]", D, gw, k));
    TcLayer* L = &g->pos;
    const int NCg = gw <= 16 ? 16 : (gw <= 32 ? 32 : 64);
    if (!tc_plan(gw, c.pos_groups * NCg, k, 1, k / 2, L, kHubHalo, NCg))
      return fail(set_err(DISSC_EUNSUPPORTED, "no tcgen05 plan for pos_conv"));
    L->Cout = D; L->groups = 1; L->group_c8 = gw / 8; L->cin8_total = D / 8;
    // grouped layers stream per-group weights: never resident across chunks
    const float* wd = w->data;
    auto packed = pack_weights_tc(*L, [=](int n, int ci, int j) {
      const int grp = n / NCg, nn = n % NCg;
      return nn < gw ? wd[((size_t)(grp * gw + nn) * gw + ci) * k + j] : 0.f;
    }, &L->inv_scale);
    if ((rc = hub_upload_h(g, packed, &L->w))) return fail(rc);
    if ((rc = hub_vec(g, wm, "encoder.pos_conv.0.bias", D, &g->pos_b))) return fail(rc);
  }
  if ((rc = hub_vec(g, wm, "encoder.layer_norm.weight", D, &g->eln_w))) return fail(rc);
  if ((rc = hub_vec(g, wm, "encoder.layer_norm.bias", D, &g->eln_b))) return fail(rc);
  g->layers.resize(c.n_layers);
  for (int l = 0; l < c.n_layers; ++l) {
    HubLayer& Ly = g->layers[l];
    const std::string p = "encoder.layers." + std::to_string(l) + ".";
    // fused QKV: rows [q * head_dim^-0.5 ; k ; v]
    const dissc_tensor *qw = wm.get(p + "self_attn.q_proj.weight"), *kw = wm.get(p + "self_attn.k_proj.weight"),
                       *vw = wm.get(p + "self_attn.v_proj.weight"), *qb = wm.get(p + "self_attn.q_proj.bias"),
                       *kb = wm.get(p + "self_attn.k_proj.bias"), *vb = wm.get(p + "self_attn.v_proj.bias");
    if (!qw || !kw || !vw || !qb || !kb || !vb || qw->numel != (int64_t)D * D || kw->numel != qw->numel ||
        vw->numel != qw->numel || qb->numel != D || kb->numel != D || vb->numel != D)
      return fail(set_err(DISSC_EMISSING, "missing %sself_attn.{q,k,v}_proj.{weight,bias}", p.c_str()));
    if (!hub_plan(D, 3 * D, 1, &Ly.qkv)) return fail(set_err(DISSC_EUNSUPPORTED, "no tcgen05 plan for QKV"));
    Ly.qkv.Cout = 3 * D;
    const float scaling = 0.125f;  // head_dim 64 ** -0.5 (fairseq MultiheadAttention.scaling)
    const float *qd = qw->data, *kd = kw->data, *vd = vw->data;
    auto packed = pack_weights_tc(Ly.qkv, [=](int n, int ci, int) {
      return n < D ? qd[(size_t)n * D + ci] * scaling : (n < 2 * D ? kd[(size_t)(n - D) * D + ci] : vd[(size_t)(n - 2 * D) * D + ci]);
    }, &Ly.qkv.inv_scale);
    if ((rc = hub_upload_h(g, packed, &Ly.qkv.w))) return fail(rc);
    std::vector<float> bias(3 * D);
    for (int i = 0; i < D; ++i) {
      bias[i] = qb->data[i] * scaling;
      bias[D + i] = kb->data[i];
      bias[2 * D + i] = vb->data[i];
    }
    if ((rc = hub_upload_f(g, bias.data(), bias.size(), &Ly.qkv_b))) return fail(rc);
    if ((rc = hub_linear(g, wm, p + "self_attn.out_proj", D, D, &Ly.out, &Ly.out_b))) return fail(rc);
    if ((rc = hub_linear(g, wm, p + "fc1", D, c.ffn_dim, &Ly.fc1, &Ly.fc1_b))) return fail(rc);
    if ((rc = hub_linear(g, wm, p + "fc2", c.ffn_dim, D, &Ly.fc2, &Ly.fc2_b))) return fail(rc);
    if ((rc = hub_vec(g, wm, p + "self_attn_layer_norm.weight", D, &Ly.ln1_w))) return fail(rc);
    if ((rc = hub_vec(g, wm, p + "self_attn_layer_norm.bias", D, &Ly.ln1_b))) return fail(rc);
    if ((rc = hub_vec(g, wm, p + "final_layer_norm.weight", D, &Ly.ln2_w))) return fail(rc);
    if ((rc = hub_vec(g, wm, p + "final_layer_norm.bias", D, &Ly.ln2_b))) return fail(rc);
  }
  if ((rc = hub_vec(g, wm, "kmeans.cluster_centers", c.n_clusters * D, &g->cent))) return fail(rc);
  if (c.n_clusters <= kKmMaxK && tc_plan(D, kKmMaxK, 1, 1, 0, &g->km, kHubHalo)) {
    const float* cd = wm.get("kmeans.cluster_centers")->data;
    const int K = c.n_clusters;
    g->km.Cout = kKmMaxK;
    auto packed = pack_weights_tc(g->km, [=](int n, int ci, int) { return n < K ? -2.f * cd[(size_t)n * D + ci] : 0.f; },
                                  &g->km.inv_scale);
    if ((rc = hub_upload_h(g, packed, &g->km.w))) return fail(rc);
    std::vector<float> nb(kKmMaxK, FLT_MAX);   // padded columns never win
    for (int n = 0; n < K; ++n) {
      double acc = 0.0;
      for (int i = 0; i < D; ++i) acc += (double)cd[(size_t)n * D + i] * (double)cd[(size_t)n * D + i];
      nb[n] = (float)acc;
    }
    if ((rc = hub_upload_f(g, nb.data(), nb.size(), &g->km_b))) return fail(rc);
    g->km_gemm = true;
  }
  // the encoder's streamed-weight GEMMs stream 57 B/clk/SM of weights (the stride-2 frame form also loads the taps it
  // skips): 2-CTA clusters with one multicast weight stream are worth 1.6 % here (neutral in the vocoder, where they stay off)
  {
    const char* e = getenv("DISSC_HUB_CLUSTER2");   // A/B switch
    const int cl = e ? (atoi(e) != 0) : 1;
    for (int l = 0; l < 6; ++l) g->conv[l].cluster2 = cl;
    g->proj.cluster2 = cl;
    for (HubLayer& Ly : g->layers) Ly.qkv.cluster2 = Ly.out.cluster2 = Ly.fc1.cluster2 = Ly.fc2.cluster2 = cl;
    if (g->km_gemm) g->km.cluster2 = cl;
  }
  *out = g;
  return DISSC_OK;
}

void dissc_hubert_destroy(dissc_hubert_t* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  for (void* p : g->allocs) cudaFree(p);
  delete g;
}

int dissc_hubert_workspace_bytes(const dissc_hubert_t* g, int B, int N, size_t* bytes) {
  DISSC_CHECK(g && bytes && B > 0 && N > 0, DISSC_EINVAL, "bad argument");
  *bytes = hub_layout(g, B, N, nullptr).total;
  return DISSC_OK;
}

int dissc_hubert_forward(dissc_hubert_t* g, const float* wave, const int32_t* n_samples, int B, int N, int64_t* units,
                         int32_t* n_frames, float* features, void* workspace, size_t workspace_bytes, void* stream) {
  DISSC_CHECK(g && wave && units && B > 0 && N > 0, DISSC_EINVAL, "bad argument");
  DISSC_CHECK(B <= 65535, DISSC_EINVAL, "B=%d exceeds the grid limit 65535", B);
  const HubShapes s = hub_shapes(N);
  DISSC_CHECK(s.T[6] > 0, DISSC_EINVAL, "clips of %d samples are shorter than the 400-sample receptive field", N);
  DISSC_CHECK((long long)B * s.T[6] < (1ll << 30), DISSC_EINVAL, "%d clips x %d frames exceed the packed row space", B, s.T[6]);
  size_t need = 0;
  dissc_hubert_workspace_bytes(g, B, N, &need);
  DISSC_CHECK(workspace && workspace_bytes >= need, DISSC_EINVAL, "workspace %zu bytes < required %zu", workspace_bytes, need);
  int dev = -1;
  DISSC_CUDA(cudaGetDevice(&dev));
  DISSC_CHECK(dev == g->device, DISSC_EINVAL, "current device %d != handle device %d", dev, g->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dissc_hubert_cfg& c = g->cfg;
  const int C = c.conv_dim, D = c.embed_dim;
  HubBuffers bf = hub_layout(g, B, N, workspace);
  // DISSC_HUB_RANGE_CHECK=1 (debug; synchronises): fail when an activation was clamped at the fp16 limit on its way into a
  // plane tensor -- for first contact with a real checkpoint, whose outlier channels random look-alike weights do not have
  static int range_check = -1;
  if (range_check < 0) {
    const char* e = getenv("DISSC_HUB_RANGE_CHECK");
    range_check = e ? (atoi(e) != 0) : 0;
  }
  if (range_check) DISSC_CUDA(cudaMemsetAsync(bf.sat, 0, sizeof(unsigned long long), st));
  auto check = [&](const __half* hi, long long slabs, int Tp_, int rows) {
    if (range_check && rows > 0) hub_range_check_kernel<<<296, 256, 0, st>>>(hi, slabs, Tp_, kHubHalo, rows, bf.sat);
  };

  hub_lengths_kernel<<<(B + 127) / 128, 128, 0, st>>>(n_samples, N, B, bf.lens);
  DISSC_CUDA(cudaGetLastError());
  // conv0 + GroupNorm + GELU -> D0 (de-interleaved planes)
  {
    const int T0 = s.T[0];
    const int nchunk = (T0 + kConv0Chunk - 1) / kConv0Chunk;
    hub_wave_moments_kernel<<<dim3(nchunk, B), 256, 0, st>>>(wave, bf.lens, N, nchunk, bf.gn_partial);
    DISSC_CUDA(cudaGetLastError());
    hub_gn_finalize_kernel<<<B, 256, 0, st>>>(bf.gn_partial, g->w0, g->gn_w, g->gn_b, bf.lens, C, nchunk, bf.gn_ss);
    DISSC_CUDA(cudaGetLastError());
    const int Tq = s.Tq[0];
    const int Tp = (int)ru(Tq, 128) + 2 * kHubHalo;
    HUB_TRY(launch_zero_halos(bf.dA[0], bf.dA[1], B * 2 * C / 8, Tp, Tq, st, kHubHalo));
    const size_t smem = (size_t)(kConv0Chunk * kConv0S + kConv0K + 2 + C * kConv0K) * sizeof(float);
    // per-device setting, a few microseconds: set on every call rather than caching a per-process flag
    DISSC_CUDA(cudaFuncSetAttribute(hub_conv0_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    // chunks cover every frame up to the rounded-up row count so that rows >= T0 inside [0, Tq) are written as zeros
    const int nchunk_apply = (2 * Tq + kConv0Chunk - 1) / kConv0Chunk;
    hub_conv0_apply_kernel<<<dim3(nchunk_apply, B), 256, smem, st>>>(wave, g->w0, bf.gn_ss, bf.lens, N, C, Tp, kHubHalo,
                                                                     bf.dA[0], bf.dA[1]);
    DISSC_CUDA(cudaGetLastError());
    check(bf.dA[0], (long long)B * 2 * C / 8, Tp, Tq);
  }
  // conv1..conv6 on the tensor cores
  __half* cur[2] = {bf.dA[0], bf.dA[1]};
  __half* nxt[2] = {bf.dB[0], bf.dB[1]};
  const int T = s.T[6];
  const int Tr = (int)ru(T, 128), Tp = Tr + 2 * kHubHalo;
  for (int l = 1; l <= 6; ++l) {
    const int Tin_q = s.Tq[l - 1], Tout = s.T[l];
    TcParams p{};
    p.a_hi = cur[0]; p.a_lo = cur[1];
    p.lengths = bf.lens + l * B; p.len_mul = 1;
    p.B = B; p.T = Tout; p.halo = kHubHalo;
    p.Tp_in = (int)ru(Tin_q, 128) + 2 * kHubHalo;
    if (l < 6) {
      const int Tq = s.Tq[l];
      p.Tp = (int)ru(Tq, 128) + 2 * kHubHalo; p.Tr = (int)ru(Tout, 128);
      p.out_hi = nxt[0]; p.out_lo = nxt[1]; p.plane_act = 2; p.out_deint = 1;
      HUB_TRY(launch_zero_halos(nxt[0], nxt[1], B * 2 * C / 8, p.Tp, Tq, st, kHubHalo));
    } else {
      p.Tp = Tp; p.Tr = Tr;
      p.out_f32b = bf.X6; p.pre_act = 2;  // GELU, then LayerNorm below
    }
    HUB_TRY(launch_conv_tc(p, g->conv[l - 1], Tout, st));
    if (l < 6) check(nxt[0], (long long)B * 2 * C / 8, p.Tp, s.Tq[l]);
    std::swap(cur[0], nxt[0]);
    std::swap(cur[1], nxt[1]);
  }
  const int* lenT = bf.lens + 6 * B;
  // LayerNorm(512) -> planes; post_extract_proj -> x (f32b + planes)
  HUB_TRY(launch_zero_halos(bf.P6[0], bf.P6[1], B * C / 8, Tp, T, st, kHubHalo));
  HUB_TRY(hub_layernorm(bf.X6, g->ln0_w, g->ln0_b, lenT, B, C, T, Tr, Tp, nullptr, bf.P6[0], bf.P6[1], st));
  check(bf.P6[0], (long long)B * C / 8, Tp, T);
  auto base = [&]() {
    TcParams p{};
    p.lengths = lenT; p.len_mul = 1; p.B = B; p.T = T; p.Tr = Tr; p.Tp = Tp; p.Tp_in = Tp; p.halo = kHubHalo;
    return p;
  };
  {
    TcParams p = base();
    p.a_hi = bf.P6[0]; p.a_lo = bf.P6[1]; p.bias = g->proj_b;
    p.out_f32b = bf.X7; p.out_hi = bf.P7[0]; p.out_lo = bf.P7[1];
    HUB_TRY(launch_zero_halos(bf.P7[0], bf.P7[1], B * D / 8, Tp, T, st, kHubHalo));
    HUB_TRY(launch_conv_tc(p, g->proj, T, st));
    check(bf.P7[0], (long long)B * D / 8, Tp, T);
  }
  {
    // x + GELU(pos_conv(x) + bias)   (fairseq TransformerEncoder.extract_features: x = x + x_conv)
    TcParams p = base();
    p.a_hi = bf.P7[0]; p.a_lo = bf.P7[1]; p.bias = g->pos_b; p.pre_act = 2; p.res = bf.X7; p.out_f32b = bf.X8;
    HUB_TRY(launch_conv_tc(p, g->pos, T, st));
  }
  // ---- transformer stack on the packed row space (HubPacked) ----
  const HubPacked pk = hub_packed_rows(B, T);
  const int* row_off = bf.row_off;
  const int* len_all = bf.row_off + B;   // "length" of the one packed pseudo-clip
  hub_row_offsets_kernel<<<1, 32, 0, st>>>(lenT, B, T, bf.row_off);
  DISSC_CUDA(cudaGetLastError());
  // PH / PA are written row by row (LayerNorm, attention: valid frames only) and read tile by tile: the rows between the
  // total and the end of the last tile must hold finite numbers.  PF / PQ are written by GEMM epilogues, whole tiles.
  for (int i = 0; i < 2; ++i) {
    DISSC_CUDA(cudaMemsetAsync(bf.PH[i], 0, (size_t)(D / 8) * pk.Rp * 16, st));
    DISSC_CUDA(cudaMemsetAsync(bf.PA[i], 0, (size_t)(D / 8) * pk.Rp * 16, st));
  }
  HUB_TRY(hub_layernorm(bf.X8, g->eln_w, g->eln_b, lenT, B, D, T, Tr, Tp, bf.H, bf.PH[0], bf.PH[1], st, row_off, pk.Rr, pk.Rp));
  auto pbase = [&]() {
    TcParams p{};
    p.lengths = len_all; p.len_mul = 1; p.B = 1; p.T = pk.R; p.Tr = pk.Rr; p.Tp = pk.Rp; p.Tp_in = pk.Rp; p.halo = kHubHalo;
    return p;
  };
  auto packed_ln = [&](const float* in, const float* w, const float* bsv) {
    return hub_layernorm(in, w, bsv, len_all, 1, D, pk.R, pk.Rr, pk.Rp, bf.H, bf.PH[0], bf.PH[1], st);
  };
  // attention on the tensor cores when all keys of a clip fit one CTA (<= 320 frames = 102 000 samples; BASELINE
  // configs[3] clips have 299); longer batches use the CUDA-core kernel.  DISSC_HUB_ATTN_TC=0 forces the latter.
  if (g_hub_attn_tc < 0) {
    const char* e = getenv("DISSC_HUB_ATTN_TC");
    g_hub_attn_tc = e ? (atoi(e) != 0) : 1;
  }
  const bool attn_tc = g_hub_attn_tc && T <= kAttnTcKeys;
  if (attn_tc) DISSC_CUDA(cudaFuncSetAttribute(hub_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnTcSmem));
  for (int l = 0; l < c.n_layers; ++l) {
    const HubLayer& Ly = g->layers[l];
    {
      TcParams p = pbase();
      p.a_hi = bf.PH[0]; p.a_lo = bf.PH[1]; p.bias = Ly.qkv_b;
      if (attn_tc) {
        // planes for the tensor-core attention; leaky-relu with slope 1 is the identity (max(v, v * 1)) and keeps the
        // specialised epilogue
        p.out_hi = bf.PQ[0]; p.out_lo = bf.PQ[1]; p.plane_act = 1; p.plane_slope = 1.0f;
      } else {
        p.out_f32b = bf.QKV;
      }
      HUB_TRY(launch_conv_tc(p, Ly.qkv, pk.R, st));
      check(bf.PH[0], D / 8, pk.Rp, pk.R);
      if (attn_tc) check(bf.PQ[0], 3 * D / 8, pk.Rp, pk.R);
    }
    if (attn_tc) {
      hub_attention_tc_kernel<<<dim3((T + 127) / 128, c.n_heads, B), kAttnTcThreads, kAttnTcSmem, st>>>(
          bf.PQ[0], bf.PQ[1], lenT, row_off, D / 8, T, pk.Rp, pk.Rp, kHubHalo, bf.PA[0], bf.PA[1]);
    } else {
      hub_attention_kernel<<<dim3((T + kAttnQ - 1) / kAttnQ, c.n_heads, B), kAttnQ, 0, st>>>(
          bf.QKV, lenT, row_off, D / 8, T, pk.Rr, pk.Rp, kHubHalo, bf.PA[0], bf.PA[1]);
    }
    DISSC_CUDA(cudaGetLastError());
    {
      TcParams p = pbase();
      p.a_hi = bf.PA[0]; p.a_lo = bf.PA[1]; p.bias = Ly.out_b; p.res = bf.H; p.out_f32b = bf.Y;
      HUB_TRY(launch_conv_tc(p, Ly.out, pk.R, st));
    }
    HUB_TRY(packed_ln(bf.Y, Ly.ln1_w, Ly.ln1_b));
    {
      TcParams p = pbase();
      p.a_hi = bf.PH[0]; p.a_lo = bf.PH[1]; p.bias = Ly.fc1_b; p.out_hi = bf.PF[0]; p.out_lo = bf.PF[1]; p.plane_act = 2;
      HUB_TRY(launch_conv_tc(p, Ly.fc1, pk.R, st));
      check(bf.PH[0], D / 8, pk.Rp, pk.R);
      check(bf.PF[0], c.ffn_dim / 8, pk.Rp, pk.R);
    }
    {
      TcParams p = pbase();
      p.a_hi = bf.PF[0]; p.a_lo = bf.PF[1]; p.bias = Ly.fc2_b; p.res = bf.H; p.out_f32b = bf.Y;
      HUB_TRY(launch_conv_tc(p, Ly.fc2, pk.R, st));
    }
    HUB_TRY(packed_ln(bf.Y, Ly.ln2_w, Ly.ln2_b));
  }
  {
    if (g->km_gemm) {
      // distances (up to the per-row constant |x|^2) by one more GEMM on the layer-6 planes, into the idle QKV buffer
      TcParams p = pbase();
      p.a_hi = bf.PH[0]; p.a_lo = bf.PH[1]; p.bias = g->km_b; p.out_f32b = bf.QKV;
      HUB_TRY(launch_conv_tc(p, g->km, pk.R, st));
      hub_kmeans_argmin_kernel<<<B * ((T + 127) / 128), 128, 0, st>>>(bf.QKV, lenT, row_off, T, pk.Rr,
                                                                       reinterpret_cast<long long*>(units));
      if (features) {
        DISSC_CUDA(cudaGetLastError());
        hub_features_out_kernel<<<dim3(T, B), 192, 0, st>>>(bf.H, lenT, row_off, D / 8, T, pk.Rr, features);
      }
    } else {
      const long long threads = (((long long)B * T + kKmFrames - 1) / kKmFrames) * 32;
      hub_kmeans_f32b_kernel<<<(int)((threads + 255) / 256), 256, 0, st>>>(
          bf.H, g->cent, lenT, row_off, B, D / 8, T, pk.Rr, c.n_clusters, reinterpret_cast<long long*>(units), features);
    }
    DISSC_CUDA(cudaGetLastError());
  }
  if (n_frames) DISSC_CUDA(cudaMemcpyAsync(n_frames, lenT, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
  if (range_check) {
    unsigned long long sat = 0;
    DISSC_CUDA(cudaMemcpyAsync(&sat, bf.sat, sizeof(sat), cudaMemcpyDeviceToHost, st));
    DISSC_CUDA(cudaStreamSynchronize(st));
    DISSC_CHECK(sat == 0, DISSC_EUNSUPPORTED,
                "%llu activations reached the fp16 limit (65504) in the split planes: this checkpoint needs activation scaling",
                sat);
  }
  return DISSC_OK;
}

int dissc_kmeans_assign(const float* x, const float* centroids, int M, int D, int K, int64_t* out, void* stream) {
  DISSC_CHECK(x && centroids && out && M > 0 && D > 0 && K > 0, DISSC_EINVAL, "bad argument");
  const long long threads = (long long)M * 32;
  kmeans_rowmajor_kernel<<<(int)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, centroids, M, D, K, reinterpret_cast<long long*>(out));
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

}  // extern "C"

// Prosody predictors on the GPU (include/dissc_b200.h, "Prosody predictors"):
//   LenPredictor        model/len_predictor.py:35-52
//   PitchPredictor      model/pitch_predictor.py:72-104   ("new": speaker embedding + linear-ramp positional table)
//   PitchPredictorBase  model/pitch_predictor.py:145-176  (BatchNorm after every conv)
// plus the sequence glue of infer.py:24-45,158-172 (dedup, carry-over rounding, repeat_interleave).
//
// Every model is: embedding gather + concat -> a stack of Conv1d(k=3, 128 ch) [+ BatchNorm(eval)] + LeakyReLU(0.01)
// -> 1-channel head(s).  Eval-mode BatchNorm is an affine map per channel, folded into the preceding conv's weights
// and bias at load time; the LenPredictor's "* norm_std + norm_mean" is applied in the head kernel's epilogue
// as the reference does (multiply then add).  Convs run on the fused fp32 conv kernel (conv1d.cuh) with per-utterance
// `lengths`, so a padded batch reproduces the reference's B=1 zero padding at every utterance's true ends.
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "launch.cuh"

namespace dissc {

struct PredConv {
  int Cin = 0, Cout = 0, k = 0, co_tile = 0;
  float* w = nullptr;
  float* bias = nullptr;
};

// token_emb[seq] ++ (spk_emb[spk] [+ pe[t]])  ->  (B, 2E, L) fp32, zero at t >= lengths[b]
__global__ void pred_embed_kernel(const long long* seq, const long long* spk, const int* lengths, const float* tok_w,
                                  const float* spk_w, const float* pe, int E, int B, int L, float* out, int tok_rows,
                                  int spk_rows, int* err) {
  const long long total = (long long)B * 2 * E * L;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % L);
    const long long bc = i / L;
    const int c = (int)(bc % (2 * E)), b = (int)(bc / (2 * E));
    const int n = lengths ? min(L, lengths[b]) : L;
    float v = 0.f;
    if (t < n) {
      if (c < E) {
        v = __ldg(tok_w + (size_t)checked_row(seq[(size_t)b * L + t], tok_rows, err, kIdxUnit) * E + c);
      } else {
        v = __ldg(spk_w + (size_t)checked_row(spk[b], spk_rows, err, kIdxSpeaker) * E + (c - E));
        if (pe) v = v + __ldg(pe + (size_t)t * E + (c - E));  // PositionalEncoding.forward, :31-38
      }
    }
    out[i] = v;
  }
}

// out = x * scale + shift  (LenPredictor: cnn2(...) * norm_std + norm_mean, model/len_predictor.py:52)
__global__ void pred_affine_kernel(const float* x, float scale, float shift, const int* lengths, int B, int L, float* out) {
  const int total = B * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / L, t = i - b * L;
    const int n = lengths ? min(L, lengths[b]) : L;
    out[i] = t < n ? __fadd_rn(__fmul_rn(x[i], scale), shift) : 0.f;
  }
}

// calc_freq (model/pitch_predictor.py:100-104): (class > 0) * (norm ? reg : mean[spk] + reg * std[spk])
__global__ void pitch_calc_freq_kernel(const float* cls, const float* reg, const long long* spk, const float* mean,
                                       const float* stdv, int n_rows, const int* lengths, int B, int L, float* out) {
  const int total = B * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / L, t = i - b * L;
    const int n = lengths ? min(L, lengths[b]) : L;
    float r = reg[i];
    bool bad = false;
    if (mean) {
      const long long s = spk[b];
      bad = (unsigned long long)s >= (unsigned long long)n_rows;  // the reference raises IndexError: NaN marks the row
      if (!bad) r = __fadd_rn(mean[s], __fmul_rn(r, stdv[s]));
    }
    out[i] = bad ? __int_as_float(0x7fc00000) : ((t < n && cls[i] > 0.f) ? r : 0.f);  // mask * value: masked positions are +0 (or -0*... the reference
                                                 // yields 0*r = +-0; both compare equal and print as 0.0/-0.0)
  }
}

// len_carryover_correction (infer.py:158-172): r = round_half_even(max(lens,1)); a = lens - r; running fp32 sum of a,
// emit +1 / -1 whenever it reaches +-1.  One thread per utterance (a serial scan over <= a few hundred tokens).
__global__ void len_carryover_kernel(const float* lens, const int* lengths, int B, int L, int* out, int* totals) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = lengths ? min(L, lengths[b]) : L;
  float total = 0.f;
  int sum = 0;
  for (int t = 0; t < L; ++t) {
    int v = 0;
    if (t < n) {
      const float x = lens[(size_t)b * L + t];
      const float r = rintf(fmaxf(x, 1.f));
      total = __fadd_rn(total, __fsub_rn(x, r));
      int adj = 0;
      if (total >= 1.f) {
        adj = 1;
        total = __fsub_rn(total, 1.f);
      } else if (total <= -1.f) {
        adj = -1;
        total = __fadd_rn(total, 1.f);
      }
      v = (int)r + adj;
    }
    out[(size_t)b * L + t] = v;
    sum += v;
  }
  if (totals) totals[b] = sum;
}

// dedup_seq (dataset/utils.py:14-16): run-length encode; pad tokens (== pad_token, infer.py:25) are dropped first.
__global__ void dedup_units_kernel(const long long* seq, const int* lengths, long long pad_token, int B, int L,
                                   long long* dd, int* counts, int* dd_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = lengths ? min(L, lengths[b]) : L;
  int m = -1;
  long long prev = 0;
  for (int t = 0; t < n; ++t) {
    const long long u = seq[(size_t)b * L + t];
    if (u == pad_token) continue;
    if (m < 0 || u != prev) {
      ++m;
      dd[(size_t)b * L + m] = u;
      counts[(size_t)b * L + m] = 0;
      prev = u;
    }
    counts[(size_t)b * L + m] += 1;
  }
  ++m;
  for (int t = m; t < L; ++t) {
    dd[(size_t)b * L + t] = pad_token;
    counts[(size_t)b * L + t] = 0;
  }
  dd_len[b] = m;
}

// torch.repeat_interleave(dd_seq, lens) per utterance (infer.py:32); one CTA per utterance.
__global__ void repeat_interleave_kernel(const long long* dd, const int* counts, const int* dd_len, long long pad_token,
                                         int B, int L, int Lout, long long* out, int* out_len) {
  const int b = blockIdx.x;
  __shared__ int s_off[1025];
  const int n = min(dd_len[b], L);
  // serial prefix (n <= 1024 per chunk) by thread 0; sequences are a few hundred tokens
  int base = 0;
  for (int c0 = 0; c0 < n; c0 += 1024) {
    const int m = min(1024, n - c0);
    if (threadIdx.x == 0) {
      int acc = base;
      for (int i = 0; i < m; ++i) {
        s_off[i] = acc;
        acc += max(counts[(size_t)b * L + c0 + i], 0);
      }
      s_off[m] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const long long u = dd[(size_t)b * L + c0 + i];
      for (int o = s_off[i]; o < s_off[i + 1] && o < Lout; ++o) out[(size_t)b * Lout + o] = u;
    }
    base = s_off[m];
    __syncthreads();
  }
  for (int o = base + threadIdx.x; o < Lout; o += blockDim.x) out[(size_t)b * Lout + o] = pad_token;
  if (threadIdx.x == 0) out_len[b] = min(base, Lout);
}

}  // namespace dissc

using namespace dissc;

struct dissc_pred {
  int kind = 0;  // DISSC_PRED_LEN / _PITCH_NEW / _PITCH_BASE
  int device = 0, E = 32, n_tokens = 0, n_speakers = 0, pe_len = 0, spk_rows = 0;
  dissc::ErrFlag err;  // out-of-range token / speaker ids (dissc_pred_status)
  std::vector<void*> allocs;
  float* tok_w = nullptr;
  float* spk_w = nullptr;
  float* pe = nullptr;
  std::vector<PredConv> trunk;  // conv + folded BN + leaky-relu each
  PredConv head_a1, head_a2, head_b1, head_b2;  // len: head_a2 only; pitch: class = a, reg = b
};

namespace dissc {

struct PredWeights {
  std::map<std::string, const dissc_tensor*> m;
  const dissc_tensor* get(const std::string& k) const {
    auto it = m.find(k);
    return it == m.end() ? nullptr : it->second;
  }
};

static int pred_upload(dissc_pred* g, const float* host, size_t n, float** out) {
  float* d = nullptr;
  DISSC_CUDA(cudaMalloc(&d, std::max<size_t>(n, 4) * sizeof(float)));
  g->allocs.push_back(d);
  DISSC_CUDA(cudaMemcpy(d, host, n * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return DISSC_OK;
}

// conv `name` (Cout,Cin,k) with optional eval-mode BatchNorm `bn` folded in:
//   y = (conv(x) - mean) * gamma / sqrt(var + eps) + beta
static int pred_make_conv(dissc_pred* g, const PredWeights& wm, const std::string& name, const std::string& bn, int Cin,
                          int Cout, int k, PredConv* L) {
  const dissc_tensor* w = wm.get(name + ".weight");
  const dissc_tensor* b = wm.get(name + ".bias");
  DISSC_CHECK(w && b, DISSC_EMISSING, "missing tensor %s.{weight,bias}", name.c_str());
  DISSC_CHECK(w->numel == (int64_t)Cin * Cout * k && b->numel == Cout, DISSC_EINVAL, "%s: expected (%d,%d,%d)",
              name.c_str(), Cout, Cin, k);
  std::vector<float> wf(w->data, w->data + w->numel), bf(b->data, b->data + Cout);
  if (!bn.empty()) {
    const dissc_tensor* ga = wm.get(bn + ".weight");
    const dissc_tensor* be = wm.get(bn + ".bias");
    const dissc_tensor* mu = wm.get(bn + ".running_mean");
    const dissc_tensor* va = wm.get(bn + ".running_var");
    DISSC_CHECK(ga && be && mu && va, DISSC_EMISSING, "missing BatchNorm tensors %s.*", bn.c_str());
    DISSC_CHECK(ga->numel == Cout && be->numel == Cout && mu->numel == Cout && va->numel == Cout, DISSC_EINVAL,
                "%s: BatchNorm tensors must have %d elements", bn.c_str(), Cout);
    for (int co = 0; co < Cout; ++co) {
      const double s = (double)ga->data[co] / std::sqrt((double)va->data[co] + 1e-5);  // nn.BatchNorm1d eps
      for (int i = 0; i < Cin * k; ++i) wf[(size_t)co * Cin * k + i] = (float)(wf[(size_t)co * Cin * k + i] * s);
      bf[co] = (float)(((double)bf[co] - mu->data[co]) * s + be->data[co]);
    }
  }
  L->Cin = Cin; L->Cout = Cout; L->k = k; L->co_tile = conv_co_tile(Cout);
  auto packed = pack_weights(wf.data(), Cin, Cout, k, L->co_tile, conv_ci_chunk(L->co_tile), false);
  int rc = pred_upload(g, packed.data(), packed.size(), &L->w);
  if (rc) return rc;
  return pred_upload(g, bf.data(), Cout, &L->bias);
}

static int pred_conv(const PredConv& L, const float* in, float* out, const int* lengths, int B, int T, bool act,
                     cudaStream_t st) {
  ConvParams p{};
  p.in = in; p.w = L.w; p.bias = L.bias; p.out = out;
  p.lengths = lengths; p.len_mul = 1;
  p.B = B; p.Cin = L.Cin; p.Cout = L.Cout; p.T = T; p.pad = (L.k - 1) / 2;
  p.post_act = act ? 1 : 0; p.post_slope = 0.01f;  // nn.LeakyReLU() default slope
  return launch_conv(p, L.k, 1, L.co_tile, false, st);
}

constexpr int kPredC = 128;

}  // namespace dissc

extern "C" {

int dissc_pred_create(dissc_pred_t** out, int kind, int n_tokens, int n_speakers, const dissc_tensor* weights,
                      int n_weights, int device) {
  DISSC_CHECK(out && weights, DISSC_EINVAL, "null argument");
  *out = nullptr;
  DISSC_CHECK(kind == DISSC_PRED_LEN || kind == DISSC_PRED_PITCH_NEW || kind == DISSC_PRED_PITCH_BASE, DISSC_EINVAL,
              "unknown predictor kind %d", kind);
  struct DeviceGuard {   // the caller's current device is restored on every exit path
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  } guard;
  cudaGetDevice(&guard.prev);
  DISSC_CUDA(cudaSetDevice(device));
  PredWeights wm;
  for (int i = 0; i < n_weights; ++i) wm.m[weights[i].name] = &weights[i];
  dissc_pred* g = new dissc_pred();
  g->kind = kind; g->device = device; g->n_tokens = n_tokens; g->n_speakers = n_speakers;
  auto fail = [&](int rc) {
    dissc_pred_destroy(g);
    return rc;
  };
  int rc;
  const dissc_tensor* tok = wm.get("token_emb.weight");
  const dissc_tensor* spk = wm.get("spk_emb.weight");
  if (!tok || !spk) return fail(set_err(DISSC_EMISSING, "missing token_emb.weight / spk_emb.weight"));
  const int spk_rows = n_speakers + (kind == DISSC_PRED_LEN ? 0 : 1);  // pitch models add a padding row (:52, :117)
  if (tok->numel % (n_tokens + 1) || spk->numel % spk_rows || tok->numel / (n_tokens + 1) != spk->numel / spk_rows)
    return fail(set_err(DISSC_EINVAL, "embedding tables do not match n_tokens=%d n_speakers=%d", n_tokens, n_speakers));
  g->E = (int)(tok->numel / (n_tokens + 1));
  g->spk_rows = spk_rows;
  if ((rc = err_flag_create(&g->err))) return fail(rc);
  if ((rc = pred_upload(g, tok->data, tok->numel, &g->tok_w))) return fail(rc);
  if ((rc = pred_upload(g, spk->data, spk->numel, &g->spk_w))) return fail(rc);
  if (kind == DISSC_PRED_PITCH_NEW) {
    const dissc_tensor* pe = wm.get("pe.pe");
    if (!pe || pe->numel % g->E) return fail(set_err(DISSC_EMISSING, "missing pe.pe (PositionalEncoding buffer)"));
    g->pe_len = (int)(pe->numel / g->E);
    if ((rc = pred_upload(g, pe->data, pe->numel, &g->pe))) return fail(rc);
  }
  const bool bn_trunk = (kind != DISSC_PRED_PITCH_NEW);
  const int n_trunk = (kind == DISSC_PRED_LEN) ? 7 : 8;
  g->trunk.resize(n_trunk + (kind == DISSC_PRED_LEN ? 0 : 1));
  for (int i = 0; i < n_trunk; ++i) {
    const std::string sfx = i == 0 ? "1" : "1" + std::to_string(i);
    if ((rc = pred_make_conv(g, wm, "cnn" + sfx, bn_trunk ? "bn" + sfx : "", i == 0 ? 2 * g->E : kPredC, kPredC, 3,
                             &g->trunk[i])))
      return fail(rc);
  }
  if (kind == DISSC_PRED_LEN) {
    if ((rc = pred_make_conv(g, wm, "cnn2", "", kPredC, 1, 3, &g->head_a2))) return fail(rc);
  } else {
    const bool nw = (kind == DISSC_PRED_PITCH_NEW);
    if ((rc = pred_make_conv(g, wm, "cnn2", nw ? "bn2" : "", kPredC, kPredC, 3, &g->trunk[n_trunk]))) return fail(rc);
    if ((rc = pred_make_conv(g, wm, "cnn_class1", nw ? "" : "bn_c1", kPredC, kPredC, 3, &g->head_a1))) return fail(rc);
    if ((rc = pred_make_conv(g, wm, "cnn_class2", "", kPredC, 1, 1, &g->head_a2))) return fail(rc);
    if ((rc = pred_make_conv(g, wm, "cnn_reg1", nw ? "" : "bn_r1", kPredC, kPredC, 3, &g->head_b1))) return fail(rc);
    if ((rc = pred_make_conv(g, wm, "cnn_reg2", "", kPredC, 1, 1, &g->head_b2))) return fail(rc);
  }
  *out = g;
  return DISSC_OK;
}

void dissc_pred_destroy(dissc_pred_t* g) {
  if (!g) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(g->device);
  for (void* p : g->allocs) cudaFree(p);
  err_flag_destroy(&g->err);
  delete g;
  if (prev >= 0) cudaSetDevice(prev);
}

int dissc_pred_status(dissc_pred_t* g) {
  DISSC_CHECK(g, DISSC_EINVAL, "null handle");
  return err_flag_take(&g->err, "predictor forward", g->n_tokens + 1, g->spk_rows);
}

int dissc_pred_workspace_bytes(const dissc_pred_t* g, int B, int L, size_t* bytes) {
  DISSC_CHECK(g && bytes && B > 0 && L > 0, DISSC_EINVAL, "bad argument");
  *bytes = (size_t)3 * B * kPredC * L * sizeof(float) + 1024;
  return DISSC_OK;
}

// trunk -> returns the buffer index holding the trunk output
static int pred_trunk(dissc_pred* g, const int64_t* seq, const int64_t* spk, const int32_t* lengths, int B, int L,
                      float* bufs[3], cudaStream_t st, int* out_idx) {
  DISSC_CHECK(g->kind != DISSC_PRED_PITCH_NEW || L <= g->pe_len, DISSC_EINVAL,
              "sequence length %d exceeds the positional table (%d rows, model/pitch_predictor.py:7)", L, g->pe_len);
  const long long tot = (long long)B * 2 * g->E * L;
  pred_embed_kernel<<<(int)std::min<long long>((tot + 255) / 256, 148 * 8), 256, 0, st>>>(
      reinterpret_cast<const long long*>(seq), reinterpret_cast<const long long*>(spk), lengths, g->tok_w, g->spk_w,
      g->pe, g->E, B, L, bufs[0], g->n_tokens + 1, g->spk_rows, g->err.dev);
  DISSC_CUDA(cudaGetLastError());
  int cur = 0;
  for (const PredConv& c : g->trunk) {
    int rc = pred_conv(c, bufs[cur], bufs[cur ^ 1], lengths, B, L, true, st);
    if (rc) return rc;
    cur ^= 1;
  }
  *out_idx = cur;
  return DISSC_OK;
}

static int pred_check(dissc_pred* g, const void* seq, const void* spk, int B, int L, void* ws, size_t ws_bytes) {
  DISSC_CHECK(g && seq && spk && B > 0 && L > 0, DISSC_EINVAL, "bad argument");
  size_t need = 0;
  dissc_pred_workspace_bytes(g, B, L, &need);
  DISSC_CHECK(ws && ws_bytes >= need, DISSC_EINVAL, "workspace %zu bytes < required %zu", ws_bytes, need);
  int dev = -1;
  DISSC_CUDA(cudaGetDevice(&dev));
  DISSC_CHECK(dev == g->device, DISSC_EINVAL, "current device %d != handle device %d", dev, g->device);
  return err_flag_take(&g->err, "an earlier predictor forward", g->n_tokens + 1, g->spk_rows);
}

int dissc_len_forward(dissc_pred_t* g, const int64_t* seq, const int64_t* spk, const int32_t* lengths, int B, int L,
                      float norm_mean, float norm_std, float* out, void* workspace, size_t workspace_bytes,
                      void* stream) {
  int rc = pred_check(g, seq, spk, B, L, workspace, workspace_bytes);
  if (rc) return rc;
  DISSC_CHECK(g->kind == DISSC_PRED_LEN && out, DISSC_EINVAL, "handle is not a LenPredictor / null out");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* base = static_cast<float*>(workspace);
  float* bufs[3] = {base, base + (size_t)B * kPredC * L, base + (size_t)2 * B * kPredC * L};
  int cur = 0;
  if ((rc = pred_trunk(g, seq, spk, lengths, B, L, bufs, st, &cur))) return rc;
  if ((rc = pred_conv(g->head_a2, bufs[cur], bufs[2], lengths, B, L, false, st))) return rc;
  pred_affine_kernel<<<std::min((B * L + 255) / 256, 148 * 8), 256, 0, st>>>(bufs[2], norm_std, norm_mean, lengths, B, L,
                                                                             out);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

int dissc_pitch_forward(dissc_pred_t* g, const int64_t* seq, const int64_t* spk, const int32_t* lengths, int B, int L,
                        float* cls, float* reg, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = pred_check(g, seq, spk, B, L, workspace, workspace_bytes);
  if (rc) return rc;
  DISSC_CHECK(g->kind != DISSC_PRED_LEN && cls && reg, DISSC_EINVAL, "handle is not a PitchPredictor / null out");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* base = static_cast<float*>(workspace);
  float* bufs[3] = {base, base + (size_t)B * kPredC * L, base + (size_t)2 * B * kPredC * L};
  int cur = 0;
  if ((rc = pred_trunk(g, seq, spk, lengths, B, L, bufs, st, &cur))) return rc;
  float* tmp = bufs[cur ^ 1];
  if ((rc = pred_conv(g->head_a1, bufs[cur], tmp, lengths, B, L, true, st))) return rc;
  if ((rc = pred_conv(g->head_a2, tmp, cls, lengths, B, L, false, st))) return rc;
  if ((rc = pred_conv(g->head_b1, bufs[cur], tmp, lengths, B, L, true, st))) return rc;
  if ((rc = pred_conv(g->head_b2, tmp, reg, lengths, B, L, false, st))) return rc;
  return DISSC_OK;
}

int dissc_pitch_calc_freq(const float* cls, const float* reg, const int64_t* spk, const float* mean, const float* std,
                          int n_stats_rows, const int32_t* lengths, int B, int L, float* out, void* stream) {
  DISSC_CHECK(cls && reg && out && B > 0 && L > 0, DISSC_EINVAL, "bad argument");
  DISSC_CHECK((mean == nullptr) == (std == nullptr) && (mean == nullptr || spk != nullptr), DISSC_EINVAL,
              "mean/std must both be given (with spk) or both NULL (normalised output)");
  DISSC_CHECK(mean == nullptr || n_stats_rows > 0, DISSC_EINVAL, "n_stats_rows=%d must be the length of mean/std", n_stats_rows);
  pitch_calc_freq_kernel<<<std::min((B * L + 255) / 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      cls, reg, reinterpret_cast<const long long*>(spk), mean, std, n_stats_rows, lengths, B, L, out);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

int dissc_len_carryover(const float* lens, const int32_t* lengths, int B, int L, int32_t* out, int32_t* totals,
                        void* stream) {
  DISSC_CHECK(lens && out && B > 0 && L > 0, DISSC_EINVAL, "bad argument");
  len_carryover_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(lens, lengths, B, L, out, totals);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

int dissc_dedup_units(const int64_t* seq, const int32_t* lengths, int64_t pad_token, int B, int L, int64_t* dd,
                      int32_t* counts, int32_t* dd_len, void* stream) {
  DISSC_CHECK(seq && dd && counts && dd_len && B > 0 && L > 0, DISSC_EINVAL, "bad argument");
  dedup_units_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(seq), lengths, pad_token, B, L, reinterpret_cast<long long*>(dd), counts, dd_len);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

int dissc_repeat_interleave(const int64_t* dd, const int32_t* counts, const int32_t* dd_len, int64_t pad_token, int B,
                            int L, int L_out, int64_t* out, int32_t* out_len, void* stream) {
  DISSC_CHECK(dd && counts && dd_len && out && out_len && B > 0 && L > 0 && L_out > 0, DISSC_EINVAL, "bad argument");
  repeat_interleave_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(dd), counts, dd_len, pad_token, B, L, L_out, reinterpret_cast<long long*>(out),
      out_len);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

}  // extern "C"

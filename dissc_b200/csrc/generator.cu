// C-ABI + host-side plan of the vocoder forward (include/dissc_b200.h).
//
// Launch plan of one forward (resblock "1", the shipped configs: 97 launches):
//   conv_pre  [embedding gather fused]                      -> act   (lrelu 0.1 applied in epilogue)
//   per stage i:
//     convT(act)                                            -> x_up
//     per resblock j, per dilation m:
//       K3: lrelu -> dilated conv -> lrelu                  -> xt    (in = x_up | r)
//       K4: conv(xt) + residual                             -> r     (m < last)
//           ... last m: MRF accumulate folded in: j==0 -> xs, j>0 -> xs += , last j -> lrelu((xs+v)/n_rk) -> act
//   conv_post + tanh                                        -> out
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv1d.cuh"
#include "conv_tc.cuh"
#include "conv_post.cuh"
#include "convt1d.cuh"
#include "launch.cuh"
#include "tc_host.cuh"
#include "resblock_tc.cuh"
#include "resblock64_tc.cuh"
#include "resblock_pack2_tc.cuh"

namespace dissc {

thread_local char g_err[512] = "";

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int err_flag_create(ErrFlag* f) {
  DISSC_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&f->host), 64, cudaHostAllocMapped));
  f->host[0] = f->host[1] = 0;
  DISSC_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&f->dev), f->host, 0));
  return DISSC_OK;
}
void err_flag_destroy(ErrFlag* f) {
  if (f->host) cudaFreeHost(f->host);
  f->host = f->dev = nullptr;
}
int err_flag_take(ErrFlag* f, const char* what, int unit_rows, int spkr_rows) {
  if (!f->host) return DISSC_OK;
  const int v = (__atomic_exchange_n(&f->host[0], 0, __ATOMIC_ACQ_REL) ? kIdxUnit : 0) |
                (__atomic_exchange_n(&f->host[1], 0, __ATOMIC_ACQ_REL) ? kIdxSpeaker : 0);
  if (!v) return DISSC_OK;
  return set_err(DISSC_EINDEX, "%s: index out of range in self:%s%s%s (unit table has %d rows, speaker table %d)", what,
                 (v & kIdxUnit) ? " unit id" : "", (v & kIdxUnit) && (v & kIdxSpeaker) ? " and" : "",
                 (v & kIdxSpeaker) ? " speaker id" : "", unit_rows, spkr_rows);
}

// One-time per-(kernel, device) setup (cudaFuncSetAttribute is a per-device setting): slot of the current device in the
// static flag arrays below.  Ordinals >= kMaxDevices share the last slot and simply repeat the (idempotent) call.
constexpr int kMaxDevices = 64;
static int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < kMaxDevices - 1 ? dev : kMaxDevices - 1;
}

// ------------------------------------------------------------------------
// weight packing (host)
// ------------------------------------------------------------------------
// -> [co_tile][chunk][CI_CHUNK][k][CO_TILE], zero padded.  transposed: source is (Cin,Cout,k).
std::vector<float> pack_weights(const float* w, int Cin, int Cout, int k, int co_tile, int ci_chunk, bool transposed) {
  const int n_cot = (Cout + co_tile - 1) / co_tile;
  const int n_chunk = (Cin + ci_chunk - 1) / ci_chunk;
  std::vector<float> out((size_t)n_cot * n_chunk * ci_chunk * k * co_tile, 0.f);
  for (int ct = 0; ct < n_cot; ++ct)
    for (int ch = 0; ch < n_chunk; ++ch)
      for (int cl = 0; cl < ci_chunk; ++cl) {
        const int ci = ch * ci_chunk + cl;
        if (ci >= Cin) continue;
        for (int j = 0; j < k; ++j)
          for (int c = 0; c < co_tile; ++c) {
            const int co = ct * co_tile + c;
            if (co >= Cout) continue;
            const float v = transposed ? w[((size_t)ci * Cout + co) * k + j] : w[((size_t)co * Cin + ci) * k + j];
            out[((((size_t)ct * n_chunk + ch) * ci_chunk + cl) * k + j) * co_tile + c] = v;
          }
      }
  return out;
}

int conv_co_tile(int Cout) { return Cout >= 64 ? 64 : (Cout >= 32 ? 32 : 16); }
int conv_ci_chunk(int co_tile) { return co_tile == 16 ? 4 : 8; }
static int convt_co_tile(int Cout) { return Cout >= 32 ? 32 : 16; }

// ------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------
template <int CO_TILE, int KW, int DIL, int CI_CHUNK, bool EMB>
static int launch_conv_inst(const ConvParams& p, cudaStream_t st) {
  using C = ConvCfg<CO_TILE, KW, DIL, CI_CHUNK>;
  auto kern = conv1d_fused_kernel<CO_TILE, KW, DIL, CI_CHUNK, EMB>;
  static bool attr_set[kMaxDevices] = {};   // the attribute is per device: one flag per device ordinal
  const int dev_ = current_device_slot();
  if (!attr_set[dev_]) {
    DISSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set[dev_] = true;
  }
  dim3 grid((p.T + C::T_TILE - 1) / C::T_TILE, (p.Cout + CO_TILE - 1) / CO_TILE, p.B);
  kern<<<grid, kThreads, C::SMEM, st>>>(p);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

template <int CO_TILE, int CI_CHUNK>
static int launch_conv_cot(const ConvParams& p, int k, int dil, bool emb, cudaStream_t st) {
#define DISSC_CONV_CASE(KW, DIL)                                                     \
  if (k == KW && dil == DIL) {                                                       \
    if (emb) {                                                                       \
      if constexpr (KW == 7 && DIL == 1)                                             \
        return launch_conv_inst<CO_TILE, KW, DIL, CI_CHUNK, true>(p, st);            \
      else                                                                           \
        return set_err(DISSC_EUNSUPPORTED, "embedding-fused conv only for k=7");     \
    }                                                                                \
    return launch_conv_inst<CO_TILE, KW, DIL, CI_CHUNK, false>(p, st);               \
  }
  DISSC_CONV_CASE(1, 1)
  DISSC_CONV_CASE(3, 1) DISSC_CONV_CASE(3, 3) DISSC_CONV_CASE(3, 5)
  DISSC_CONV_CASE(5, 1) DISSC_CONV_CASE(5, 3) DISSC_CONV_CASE(5, 5)
  DISSC_CONV_CASE(7, 1) DISSC_CONV_CASE(7, 3) DISSC_CONV_CASE(7, 5)
  DISSC_CONV_CASE(11, 1) DISSC_CONV_CASE(11, 3) DISSC_CONV_CASE(11, 5)
#undef DISSC_CONV_CASE
  return set_err(DISSC_EUNSUPPORTED, "conv1d kernel_size=%d dilation=%d has no sm_100a instantiation", k, dil);
}

bool conv_supported(int k, int dil) {
  if (k == 1) return dil == 1;
  return (k == 3 || k == 5 || k == 7 || k == 11) && (dil == 1 || dil == 3 || dil == 5);
}

int launch_conv(const ConvParams& p, int k, int dil, int co_tile, bool emb, cudaStream_t st) {
  switch (co_tile) {
    case 64: return launch_conv_cot<64, 8>(p, k, dil, emb, st);
    case 32: return launch_conv_cot<32, 8>(p, k, dil, emb, st);
    case 16: return launch_conv_cot<16, 4>(p, k, dil, emb, st);
  }
  return set_err(DISSC_EINVAL, "bad co_tile %d", co_tile);
}

template <int CO_TILE, int KW, int U>
static int launch_convt_inst(const ConvTParams& p, cudaStream_t st) {
  using C = ConvTCfg<CO_TILE, KW, U, 8>;
  auto kern = convt1d_kernel<CO_TILE, KW, U, 8>;
  static bool attr_set[kMaxDevices] = {};   // the attribute is per device: one flag per device ordinal
  const int dev_ = current_device_slot();
  if (!attr_set[dev_]) {
    DISSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set[dev_] = true;
  }
  const int n_frames = (p.Tout + p.pad - 1) / U + 1;
  dim3 grid((n_frames + C::F_TILE - 1) / C::F_TILE, (p.Cout + CO_TILE - 1) / CO_TILE, p.B);
  kern<<<grid, kThreads, C::SMEM, st>>>(p);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

static bool convt_supported(int k, int u) {
  return (k == 11 && u == 5) || (k == 8 && u == 4) || (k == 4 && u == 2) || (k == 16 && u == 8);
}

static int launch_convt(const ConvTParams& p, int k, int u, int co_tile, cudaStream_t st) {
#define DISSC_CONVT_CASE(KW, U)                                             \
  if (k == KW && u == U) {                                                  \
    if (co_tile == 32) return launch_convt_inst<32, KW, U>(p, st);          \
    return launch_convt_inst<16, KW, U>(p, st);                             \
  }
  DISSC_CONVT_CASE(11, 5) DISSC_CONVT_CASE(8, 4) DISSC_CONVT_CASE(4, 2) DISSC_CONVT_CASE(16, 8)
#undef DISSC_CONVT_CASE
  return set_err(DISSC_EUNSUPPORTED, "conv_transpose1d kernel_size=%d stride=%d has no sm_100a instantiation", k, u);
}

static int launch_conv_post(const ConvPostParams& p, int k, cudaStream_t st) {
  if (k != 7) return set_err(DISSC_EUNSUPPORTED, "conv_post kernel_size=%d (only 7)", k);
  if (p.Cin > kPostMaxCin) return set_err(DISSC_EUNSUPPORTED, "conv_post Cin=%d > %d", p.Cin, kPostMaxCin);
  const int per_cta = kThreads * kPostRT;
  dim3 grid((p.T + per_cta - 1) / per_cta, p.B);
  conv_post_kernel<7><<<grid, kThreads, 0, st>>>(p);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

// ------------------------------------------------------------------------
// tensor-core path: plan + weight packing + launch
// ------------------------------------------------------------------------
// NC == 256: one accumulator (so TMEM double-buffers) instead of main+cross.  Measured end to end (profiles/README.md):
// 1.8e-5 max-abs vs fp64 instead of 7e-6, stage 0 23 % faster.  DISSC_TC_SINGLE_ACC=0 restores the dual accumulator.
static int g_single_acc256 = -1;   // -1: env DISSC_TC_SINGLE_ACC or the default (0)
// plan-time tuning knobs (dissc_tc_set_tuning): activation buffers preferred by the streamed-weight layers, and whether
// the N >= 128 kernels use their separate weight-producer thread
static int g_tc_na_pref = -1, g_tc_split_w = 1;
// 256-column GEMMs as two 128-column chunks (each with its own main + cross accumulators, double-buffered in TMEM)
// instead of one 256-column chunk: -1 = env DISSC_TC_SPLIT256 or the default
static int g_tc_kb64 = -1;       // 64-channel activation blocks for the NC = 128 kernels (env DISSC_TC_KB64)
static int g_tc_split256 = -1;
extern int g_hub_attn_tc;        // hubert.cu: tensor-core attention (key 4)
static int g_tc_cluster2 = -1;   // 2-CTA clusters with multicast weights (-1: env DISSC_TC_CLUSTER2 or the default 0)

// Cin input channels, ncols GEMM columns (Cout, or u*Cout for a transposed conv), `taps` shifted by `dil` rows.
bool tc_plan(int Cin, int ncols, int taps, int dil, int pad, TcLayer* L, int halo, int force_nc, int single_acc, int pair2) {
  L->ok = false;
  L->pair2 = 0;
  if (Cin < 1 || taps < 1 || pad > halo || pad < 0) return false;
  if ((taps - 1) * dil - pad > halo) return false;  // right halo
  if (g_tc_split256 < 0) {
    const char* e = getenv("DISSC_TC_SPLIT256");
    g_tc_split256 = e ? (atoi(e) != 0) : 1;
  }
  int NC;
  if (force_nc == 0 && g_tc_split256 && ncols >= 256 && ncols % 256 == 0) force_nc = 128;
  if (force_nc) {
    if ((force_nc != 16 && force_nc != 32 && force_nc != 64 && force_nc != 128 && force_nc != 256) || ncols % force_nc)
      return false;
    NC = force_nc;
  } else if (ncols <= 256) {
    if (ncols != 16 && ncols != 32 && ncols != 64 && ncols != 128 && ncols != 256) return false;
    NC = ncols;
  } else {
    if (ncols % 256) return false;
    NC = 256;
  }
  if (g_single_acc256 < 0) {
    const char* e = getenv("DISSC_TC_SINGLE_ACC");
    g_single_acc256 = e ? (atoi(e) != 0) : 0;
  }
  L->Cin = Cin; L->Cin_pad = (Cin + 15) / 16 * 16; L->NC = NC; L->n_chunks = ncols / NC;
  L->k = taps; L->dil = dil; L->pad = pad;
  L->KB = (L->Cin_pad % 32 == 0) ? 32 : 16;
  // streamed-weight 128-column kernels: 64-channel activation blocks halve the number of barrier round trips per MMA
  // cycle (a block = k taps x 4 k-steps, a weight stage = one 32 KB tap), which is what the one-chunk 256-column
  // configuration has and why it was faster (profiles/README.md r02)
  if (g_tc_kb64 < 0) {
    const char* e = getenv("DISSC_TC_KB64");
    g_tc_kb64 = e ? (atoi(e) != 0) : 1;
  }
  if (g_tc_kb64 && NC == 128 && L->Cin_pad % 64 == 0 && L->Cin_pad >= 128) L->KB = 64;
  {
    // 256-column single-accumulator chunks with one or two taps (the HuBERT GEMMs, which ask for them explicitly): a block of
    // 32 channels is only 2 k-steps x 3 MMAs = 768 tensor cycles, and the per-block barrier round trips then cost ~10 % of
    // the encoder (7.23 -> 6.93 ms per 32 clips with 64-channel blocks; DISSC_TC_KB64_256=0 restores 32)
    static int kb64_256 = -1;
    if (kb64_256 < 0) {
      const char* e = getenv("DISSC_TC_KB64_256");
      kb64_256 = e ? atoi(e) : 1;
    }
    if (kb64_256 && NC == 256 && single_acc == 1 && taps <= 2 && L->Cin_pad % 64 == 0 && L->Cin_pad >= 128) L->KB = 64;
  }
  L->n_cb = L->Cin_pad / L->KB;
  L->single_acc = (NC == 256) ? (single_acc >= 0 ? single_acc : g_single_acc256) : 0;
  L->acc_cols = L->single_acc ? NC : 2 * NC;
  L->nbuf = (2 * L->acc_cols <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < L->nbuf * L->acc_cols) cols *= 2;
  L->tmem_cols = cols;
  const int R = 128 + (taps - 1) * dil;
  const size_t a_bytes = (size_t)4 * L->KB * R;
  const size_t w_tap = (size_t)4 * L->KB * NC / (pair2 ? 2 : 1);   // bytes of one tap of a stage in ONE CTA's shared memory
  L->SPC = (int)std::max<size_t>(1, ((size_t)taps * w_tap + 32767) / 32768);
  L->JG = (taps + L->SPC - 1) / L->SPC;
  L->SPC = (taps + L->JG - 1) / L->JG;
  const size_t slot = (size_t)L->JG * w_tap;
  const int total_slots = L->n_cb * L->SPC;
  int cap = NC <= 32 ? 3 : (NC <= 64 ? 2 : 1);  // matches __launch_bounds__ of conv_tc_kernel<NC>
  cap = std::min(cap, 512 / cols);
  for (int ctas = cap; ctas >= 1; --ctas) {
    const size_t budget = kSmemPerSm / ctas - 1536;  // static shared + alignment + per-CTA reservation
    auto misc = [&](int ns) { return (size_t)((L->n_chunks * NC + 1) & ~1) * 4 + (size_t)(12 + 2 * ns) * 8 + 128; };
    if (!pair2 && L->n_chunks == 1 && total_slots <= kTcMaxStages) {
      for (int na = 4; na >= 2; --na) {
        const size_t need = (size_t)na * a_bytes + (size_t)total_slots * slot + misc(total_slots);
        if (need <= budget) {
          L->resident = 1; L->NS = total_slots; L->NA = na; L->smem = need; L->ctas_per_sm = ctas; L->ok = true;
          return true;
        }
      }
    }
    // streamed weights.  The N >= 128 kernels have separate activation / weight producer threads, so a third activation
    // buffer really doubles the prefetch window of a block (one loaded-HBM latency at N = 128, k = 11 with two); it is
    // taken when >= 4 weight slots remain (the ring only has to cover the L2 latency).  DISSC_TC_NA=2 restores two.
    if (g_tc_na_pref < 0) {
      const char* e = getenv("DISSC_TC_NA");
      g_tc_na_pref = e ? atoi(e) : 0;   // 0: the heuristic below
    }
    L->split_w = g_tc_split_w;
    // measured (profiles/README.md, r02): with 8 channel blocks per item (Cin = 256) four activation buffers pay, with
    // 4 (Cin = 128) they do not
    const int na_pref = g_tc_na_pref > 0 ? g_tc_na_pref : (L->n_cb >= 8 ? 4 : 2);
    for (int na = (NC >= 128) ? std::max(2, std::min(4, na_pref)) : 2; na >= 2; --na) {
      const size_t fixed = (size_t)na * a_bytes + misc(8);
      if (budget <= fixed + 2 * slot) continue;
      const int ns = (int)std::min<size_t>(8, (budget - fixed) / slot);
      if (na > 2 && ns < (pair2 ? 3 : 4)) continue;
      L->resident = 0; L->NA = na;
      L->NS = ns;
      L->smem = (size_t)na * a_bytes + (size_t)L->NS * slot + misc(L->NS);
      L->ctas_per_sm = ctas; L->ok = true;
      L->pair2 = (pair2 && NC >= 128 && ctas == 1 && L->split_w) ? 1 : 0;
      if (pair2 && !L->pair2) { L->ok = false; return false; }   // the caller asked for the pair layout: no silent fallback
      return true;
    }
  }
  return false;
}

bool tc_plan_conv(int Cin, int Cout, int k, int dil, TcLayer* L) {
  if (Cout % 8) return false;
  if (!tc_plan(Cin, Cout, k, dil, (k * dil - dil) / 2, L, kTcHalo)) return false;
  L->Cout = Cout;
  return true;
}
// Polyphase transposed conv (Cin, Cout, k) stride u, padding (k-u)/2: M = ceil(k/u) taps per phase; GEMM column
// n = phase*Cout + co; tap j' reads input frame q - (M-1) + j' and carries W[:, co, phase + (M-1-j')*u].
static bool tc_plan_convt(int Cin, int Cout, int k, int u, TcLayer* L) {
  if (Cout % 8 || Cout > 256 || u < 1) return false;
  const int M = (k + u - 1) / u;
  if (!tc_plan(Cin, u * Cout, M, 1, M - 1, L, kTcHalo)) return false;
  // a chunk holds whole output phases, or a phase spans whole chunks (the epilogue derives phase and channel from the
  // global column index either way)
  if (L->NC % Cout && Cout % L->NC) { L->ok = false; return false; }
  L->Cout = Cout; L->up = u; L->up_P = L->NC / Cout; L->up_pad = (k - u) / 2;
  return true;
}

// SM count of the CURRENT device (one cached value per device ordinal; relaxed atomics: racing threads compute the
// same number).
static int num_sms() {
  static int cache[kMaxDevices] = {};
  const int slot = current_device_slot();
  int n = __atomic_load_n(&cache[slot], __ATOMIC_RELAXED);
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    __atomic_store_n(&cache[slot], n, __ATOMIC_RELAXED);
  }
  return n;
}

// Launch with the programmatic-stream-serialization attribute (PDL, common.cuh): only for kernels that call pdl_wait()
// before their first dependent global access.  DISSC_PDL=0 falls back to ordinary launches.
static int g_pdl = -1;
template <typename P>
static cudaError_t launch_pdl(void (*kern)(P), int grid, int block, size_t smem, cudaStream_t st, const P& p,
                              int cluster = 1) {
  if (g_pdl < 0) {
    const char* e = getenv("DISSC_PDL");
    g_pdl = e ? (atoi(e) != 0) : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int NC, int EPW, int MODE, bool PAIR = false>
static int launch_conv_tc_inst(const TcParams& p, const TcLayer& L, int grid, cudaStream_t st) {
  static bool attr_set[kMaxDevices] = {};   // the attribute is per device: one flag per device ordinal
  const int dev_ = current_device_slot();
  if (!attr_set[dev_]) {
    DISSC_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NC, EPW, MODE, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kSmemPerSm - 1024)));
    attr_set[dev_] = true;
  }
  const cudaError_t le = launch_pdl(conv_tc_kernel<NC, EPW, MODE, PAIR>, grid, tc_threads(EPW), L.smem, st, p,
                                    (p.cluster2 || PAIR) ? 2 : 1);
  if (le != cudaSuccess) {
    cudaGetLastError();
    return set_err(DISSC_ECUDA, "conv_tc_kernel<%d,%d,%d> grid %d cluster %d pair %d smem %zu tiles %d chunks %d: %s", NC, EPW, MODE,
                   grid, p.cluster2, p.pair2, L.smem, p.n_tiles, p.n_chunks, cudaGetErrorString(le));
  }
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

template <int NC, int EPW>
static int launch_conv_tc_mode(const TcParams& p, const TcLayer& L, int mode, int grid, cudaStream_t st) {
  if constexpr (NC >= 128 && EPW == 8) {
    if (p.pair2) {   // CTA-pair instantiations: plain convs and the generic epilogue
      if (mode == kTcConv) return launch_conv_tc_inst<NC, EPW, kTcConv, true>(p, L, grid, st);
      if (mode == kTcGeneric) return launch_conv_tc_inst<NC, EPW, kTcGeneric, true>(p, L, grid, st);
      return set_err(DISSC_EUNSUPPORTED, "no CTA-pair kernel for transposed convs");
    }
  } else {
    if (p.pair2) return set_err(DISSC_EINVAL, "CTA-pair plan on a %d-column chunk", NC);
  }
  switch (mode) {
    case kTcConv: return launch_conv_tc_inst<NC, EPW, kTcConv>(p, L, grid, st);
    case kTcUp: return launch_conv_tc_inst<NC, EPW, kTcUp>(p, L, grid, st);
  }
  if constexpr (NC == 64 && EPW == 8) {
    return set_err(DISSC_EINVAL, "no generic 8-warp N=64 kernel");
  } else {
    return launch_conv_tc_inst<NC, EPW, kTcGeneric>(p, L, grid, st);
  }
}

// DISSC_TC_FAST=0: always the generic epilogue.  DISSC_TC_EPW64=4: four (not eight) epilogue warps for the one-CTA-per-SM
// N=64 layers.  Both exist for A/B measurements (scripts/profile_layers.py).
static int g_tc_fast = -1, g_tc_epw64 = -1;

static bool aligned32(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; }

// Which specialised epilogue (conv_tc.cuh MODE) serves this launch; kTcGeneric when any rarely used switch is on.
static int tc_mode(const TcParams& p, const TcLayer& L) {
  if (g_tc_fast < 0) {
    const char* e = getenv("DISSC_TC_FAST");
    g_tc_fast = e ? (atoi(e) != 0) : 1;
  }
  if (!g_tc_fast || p.groups || p.pre_act || p.out_deint || p.out_plain || (p.out_hi && p.plane_act != 1)) return kTcGeneric;
  if (!aligned32(p.res) || !aligned32(p.acc_in) || !aligned32(p.out_f32b)) return kTcGeneric;
  if (!L.up) return (L.n_chunks * L.NC == L.Cout) ? kTcConv : kTcGeneric;
  if (p.res || p.acc_in || p.div != 0.f || p.cout_log2 < 0) return kTcGeneric;
  return kTcUp;
}

// p carries the tensors, B, T (output rows), Tr, Tp, Tp_in, lengths and the epilogue switches; `rows` is the number of
// GEMM rows per utterance (output time steps for a conv, input frames for a transposed conv).
int launch_conv_tc(TcParams p, const TcLayer& L, int rows, cudaStream_t st) {
  if (p.halo == 0) p.halo = kTcHalo;
  if (L.groups) { p.groups = 1; p.group_c8 = L.group_c8; }
  p.k = L.k; p.dil = L.dil; p.pad = L.pad; p.KB = L.KB; p.n_cb = L.n_cb; p.JG = L.JG; p.SPC = L.SPC; p.NS = L.NS;
  p.resident = L.resident; p.tmem_cols = L.tmem_cols; p.acc_cols = L.acc_cols; p.nbuf = L.nbuf; p.NA = L.NA;
  if (p.in_scale == 0.f) p.in_scale = 1.f;
  if (p.plane_scale == 0.f) p.plane_scale = 1.f;
  p.single_acc = L.single_acc; p.w = L.w; p.w_inv_scale = L.inv_scale / p.in_scale;   // powers of two: exact
  p.Cin8 = L.cin8_total ? L.cin8_total : L.Cin_pad / 8; p.Cout = L.Cout; p.n_chunks = L.n_chunks;
  p.up = L.up; p.up_P = L.up_P; p.up_pad = L.up_pad;
  p.cb_split = L.cb_split; p.k_hi = L.k_hi; p.split_w = L.split_w;
  p.cout_log2 = -1;
  for (int s = 3; s < 12; ++s)
    if ((1 << s) == L.Cout) p.cout_log2 = s;
  p.tiles_per_b = (rows + 127) / 128;
  p.n_tiles = p.B * p.tiles_per_b;
  p.n_items = p.n_tiles * L.n_chunks;
  int grid = std::min(p.n_items, num_sms() * L.ctas_per_sm);
  const int mode = tc_mode(p, L);
  if (g_tc_epw64 < 0) {
    const char* e = getenv("DISSC_TC_EPW64");
    g_tc_epw64 = e ? atoi(e) : 8;
  }
  // 2-CTA clusters sharing one multicast weight stream: the 8-epilogue-warp kernels (one CTA per SM) with streamed weights
  // and a dedicated weight-producer thread, when there are at least two tiles to pair up
  if (g_tc_cluster2 < 0) {
    const char* e = getenv("DISSC_TC_CLUSTER2");
    g_tc_cluster2 = e ? (atoi(e) != 0) : 0;   // measured: no change (profiles/README.md r02) -> off by default
  }
  const bool epw8 = L.NC >= 128 || (L.NC == 64 && L.ctas_per_sm == 1 && g_tc_epw64 == 8 && mode != kTcGeneric);
  p.cluster2 = 0;
  p.pair2 = 0;
  if (L.pair2) {
    // weights are packed per column half: always the pair form (a lone tile is computed by both CTAs)
    const int pair_items = ((p.n_tiles + 1) / 2) * L.n_chunks;
    grid = std::max(2, std::min(2 * pair_items, num_sms()) & ~1);
    p.pair2 = 1;
  } else if ((L.cluster2 < 0 ? g_tc_cluster2 : L.cluster2) && epw8 && !L.resident && L.split_w && L.ctas_per_sm == 1 && p.n_tiles >= 2) {
    const int pair_items = ((p.n_tiles + 1) / 2) * L.n_chunks;
    grid = std::min(2 * pair_items, num_sms()) & ~1;
    p.cluster2 = grid >= 2 ? 1 : 0;
    if (!p.cluster2) grid = std::min(p.n_items, num_sms() * L.ctas_per_sm);
  }
  switch (L.NC) {
    case 16: return launch_conv_tc_mode<16, 4>(p, L, mode, grid, st);
    case 32: return launch_conv_tc_mode<32, 4>(p, L, mode, grid, st);
    case 64:
      // a lone CTA per SM (weights of the k = 7 / 11 layers fill shared memory): four epilogue warps cannot keep up
      // with the tensor pipe, eight can
      if (L.ctas_per_sm == 1 && g_tc_epw64 == 8 && mode != kTcGeneric) return launch_conv_tc_mode<64, 8>(p, L, mode, grid, st);
      return launch_conv_tc_mode<64, 4>(p, L, mode, grid, st);
    case 128: return launch_conv_tc_mode<128, 8>(p, L, mode, grid, st);
    case 256: return launch_conv_tc_mode<256, 8>(p, L, mode, grid, st);
  }
  return set_err(DISSC_EINVAL, "bad tensor-core chunk width %d", L.NC);
}

int launch_zero_halos(__half* hi, __half* lo, int slabs, int Tp, int T, cudaStream_t st, int halo) {
  const long long total = (long long)slabs * (Tp - T);
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  tc_zero_halos_kernel<<<std::max(blocks, 1), 256, 0, st>>>(hi, lo, slabs, Tp, T, halo);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

// ------------------------------------------------------------------------
// fused ResBlock pair (resblock_tc.cuh): plan + launch
// ------------------------------------------------------------------------
constexpr int kPairHalo = 32;    // f32h: slack rows in front of every slab (>= p1 + p2 = 30 for k=11, d=5)
constexpr int kPairSlack = 352;  // f32h: Tpf = roundup(T,128) + kPairSlack (a 256-sample tile of the packed kernel reads
                                 // up to ~256 + 30 rows past the last sample)
static int g_use_pair = -1;      // DISSC_TC_PAIR=0 disables the fused pair kernel

struct PairLayer {
  bool ok = false;
  int C = 0, k = 0, dil = 1, tmem_cols = 0, ctas_per_sm = 1;
  size_t smem = 0;
};

static bool pair_plan(int C, int k, int dil, const TcLayer& c1, const TcLayer& c2, PairLayer* L) {
  L->ok = false;
  if (g_use_pair < 0) {
    const char* e = getenv("DISSC_TC_PAIR");
    g_use_pair = e ? (atoi(e) != 0) : 1;
  }
  if (!g_use_pair || (C != 16 && C != 32) || !(k & 1) || k < 1 || k > 33) return false;
  if (!c1.ok || !c2.ok || c1.n_cb != 1 || c2.n_cb != 1 || c1.NC != C || c2.NC != C || c1.n_chunks != 1) return false;
  const int p1 = dil * (k - 1) / 2, p2 = (k - 1) / 2;
  if (p1 + p2 > kPairHalo || 128 - (k - 1) < 64) return false;
  const size_t R1 = 128 + (size_t)(k - 1) * dil, R2 = 128 + (k - 1), C8 = C / 8;
  const size_t stg = C8 * R1 * 32, xop = 2 * C8 * R1 * 16, xt = 2 * C8 * R2 * 16, w = (size_t)k * C8 * 2 * C * 16;
  L->smem = 2 * (stg + xop + xt + w) + 2 * C * 4 + 17 * 8 + 128;
  if (L->smem > kSmemPerSm - 1536) return false;
  L->C = C; L->k = k; L->dil = dil;
  L->tmem_cols = 8 * C;  // two groups x (acc1, acc2) x (main | cross)
  L->ctas_per_sm = (C == 16) ? (int)std::max<size_t>(1, std::min<size_t>(2, kSmemPerSm / (L->smem + 1536))) : 1;
  L->ok = true;
  return true;
}

template <int NC>
static int launch_pair_nc(const PairParams& p, const PairLayer& L, int grid, cudaStream_t st) {
  static bool attr_set[kMaxDevices] = {};   // the attribute is per device: one flag per device ordinal
  const int dev_ = current_device_slot();
  if (!attr_set[dev_]) {
    DISSC_CUDA(cudaFuncSetAttribute(resblock_pair_tc_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kSmemPerSm - 1024)));
    attr_set[dev_] = true;
  }
  DISSC_CUDA(launch_pdl(resblock_pair_tc_kernel<NC>, grid, PairCfg<NC>::THREADS, L.smem, st, p));
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

static int launch_pair(PairParams p, const PairLayer& L, const TcLayer& c1, const TcLayer& c2, cudaStream_t st) {
  p.k = L.k; p.dil = L.dil; p.tmem_cols = L.tmem_cols;
  if (p.in_scale == 0.f) p.in_scale = 1.f;
  if (p.xt_scale == 0.f) p.xt_scale = 1.f;
  if (p.plane_scale == 0.f) p.plane_scale = 1.f;
  p.w1 = c1.w; p.w2 = c2.w; p.inv1 = c1.inv_scale / p.in_scale; p.inv2 = c2.inv_scale / p.xt_scale;
  const int m_out = 128 - (L.k - 1);
  p.tiles_per_b = (p.T + m_out - 1) / m_out;
  p.n_tiles = p.B * p.tiles_per_b;
  const int grid = std::min(p.n_tiles, num_sms() * L.ctas_per_sm);
  if (L.C == 16) return launch_pair_nc<16>(p, L, grid, st);
  return launch_pair_nc<32>(p, L, grid, st);
}

// ------------------------------------------------------------------------
// fused ResBlock pair, C = 16, two samples per GEMM row (resblock_pack2_tc.cuh): plan + weights + launch
// ------------------------------------------------------------------------
static int g_use_pack2 = -1;  // DISSC_TC_PACK2=0 falls back to the one-sample-per-row pair kernel
static int g_pack2_groups = -1;  // DISSC_TC_PACK2_GROUPS=2|3: tiles in flight per CTA (default 2)

struct Pack2Layer {
  bool ok = false;
  int G = 2;             // worker groups = tiles in flight per CTA
  int k = 0, dil = 1, S = 0, U = 0, M_out = 0;
  int base[kPack2MaxDil] = {}, n_out[kPack2MaxDil] = {};
  unsigned d_magic = 0;
  size_t smem = 0;
  __half* w1 = nullptr;  // block-Toeplitz packed weights (device)
  __half* w2 = nullptr;
  float inv1 = 1.f, inv2 = 1.f;
};

// Largest even U (xt samples per tile) whose d phases -- each ceil(n/2) rows of decimated sample pairs, inputs of a
// phase directly followed by the next phase -- keep every OUTPUT row inside the 128 TMEM lanes and every input row
// inside the 128 + S - 1 operand rows.
static bool pack2_plan(int C, int k, int dil, Pack2Layer* L) {
  L->ok = false;
  if (g_use_pack2 < 0) {
    const char* e = getenv("DISSC_TC_PACK2");
    g_use_pack2 = e ? (atoi(e) != 0) : 1;
  }
  if (!g_use_pack2 || C != 16 || !(k & 1) || k < 1 || k > 33 || dil < 1 || dil > kPack2MaxDil) return false;
  const int p1 = dil * (k - 1) / 2, p2 = (k - 1) / 2, S = (k + 1) / 2, RX = 128 + S - 1;
  if (p1 + p2 > kPairHalo) return false;
  for (int U = 256; U >= 64 + (k - 1); U -= 2) {
    const int Lx = U + (k - 1) * dil;
    int base = 0;
    bool ok = true;
    for (int f = 0; f < dil && ok; ++f) {
      const int n_out = U > f ? (U - f + dil - 1) / dil : 0;
      const int n_in = (Lx - f + dil - 1) / dil;
      L->base[f] = base;
      L->n_out[f] = n_out;
      ok = base + (n_out + 1) / 2 <= 128 && base + (n_in + 1) / 2 <= RX;
      base += (n_in + 1) / 2;
    }
    if (!ok) continue;
    L->k = k; L->dil = dil; L->S = S; L->U = U; L->M_out = U - (k - 1);
    L->d_magic = (unsigned)(0xFFFFFFFFu / (unsigned)dil + 1u);   // ceil(2^32 / d); d = 1: wraps to 0, handled below
    const size_t stg = (size_t)2 * Lx * 32, tile = (size_t)2 * 4 * RX * 16, w = (size_t)S * 4096;
    if (g_pack2_groups < 0) {
      const char* e = getenv("DISSC_TC_PACK2_GROUPS");
      g_pack2_groups = e ? atoi(e) : 2;
    }
    // two tiles in flight (96 registers per thread).  Three fit in shared memory, but 26 warps leave 72 registers per
    // thread: the kernel spills and the MRF-accumulate pairs get 25 % slower (profiles/README.md r02) -- opt-in only
    for (int G = std::max(2, std::min(3, g_pack2_groups)); G >= 2; --G) {
      L->smem = G * (stg + 2 * tile) + 2 * w + 2 * 16 * 4 + (8 * G + 1) * 8 + 128;
      if (L->smem <= kSmemPerSm - 1536) {
        L->G = G;
        L->ok = true;
        return true;
      }
    }
    return false;
  }
  return false;
}

// B_s[(qi, ci), (qo, co)] = W[co][ci][2s + qi - qo] (zero outside the kernel), packed like every other tensor-core weight
// ([shift][k8][hi|lo][32][8], power-of-two pre-scale)
static std::vector<__half> pack2_weights(const float* w, int k, int S, float* inv_scale) {
  TcLayer V;
  V.Cin = 32; V.Cin_pad = 32; V.NC = 32; V.n_chunks = 1; V.k = S; V.KB = 32; V.n_cb = 1;
  return pack_weights_tc(V, [=](int n, int cip, int sh) {
    const int qo = n >> 4, co = n & 15, qi = cip >> 4, ci = cip & 15, j = 2 * sh + qi - qo;
    return (j >= 0 && j < k) ? w[((size_t)co * 16 + ci) * k + j] : 0.f;
  }, inv_scale);
}

static int launch_pack2(Pack2Params p, const Pack2Layer& L, cudaStream_t st) {
  static bool attr_set[kMaxDevices] = {};   // the attribute is per device: one flag per device ordinal
  const int dev_ = current_device_slot();
  if (!attr_set[dev_]) {
    DISSC_CUDA(cudaFuncSetAttribute(resblock_pack2_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kSmemPerSm - 1024)));
    DISSC_CUDA(cudaFuncSetAttribute(resblock_pack2_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kSmemPerSm - 1024)));
    attr_set[dev_] = true;
  }
  p.k = L.k; p.dil = L.dil; p.S = L.S; p.U = L.U; p.M_out = L.M_out;
  for (int f = 0; f < kPack2MaxDil; ++f) { p.base[f] = L.base[f]; p.n_out[f] = L.n_out[f]; }
  p.d_magic = L.d_magic;
  if (p.in_scale == 0.f) p.in_scale = 1.f;
  if (p.xt_scale == 0.f) p.xt_scale = 1.f;
  if (p.plane_scale == 0.f) p.plane_scale = 1.f;
  p.w1 = L.w1; p.w2 = L.w2; p.inv1 = L.inv1 / p.in_scale; p.inv2 = L.inv2 / p.xt_scale;
  p.tiles_per_b = (p.T + L.M_out - 1) / L.M_out;
  p.n_tiles = p.B * p.tiles_per_b;
  const int grid = std::min(p.n_tiles, num_sms());
  if (L.G == 3)
    DISSC_CUDA(launch_pdl(resblock_pack2_tc_kernel<3>, grid, pack2_threads(3), L.smem, st, p));
  else
    DISSC_CUDA(launch_pdl(resblock_pack2_tc_kernel<2>, grid, pack2_threads(2), L.smem, st, p));
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

// ------------------------------------------------------------------------
// fused ResBlock pair, C = 64 (resblock64_tc.cuh): planes in / planes out, streamed weights
// ------------------------------------------------------------------------
static int g_use_pair64 = -1;  // DISSC_TC_PAIR64=0 disables it (the stage then runs the unfused c1 / c2 launches)

struct Pair64Layer {
  bool ok = false;
  int k = 0, dil = 1, NS = 0;
  size_t smem = 0;
};

static bool pair64_plan(int C, int k, int dil, const TcLayer& c1, const TcLayer& c2, Pair64Layer* L) {
  L->ok = false;
  if (g_use_pair64 < 0) {
    const char* e = getenv("DISSC_TC_PAIR64");
    g_use_pair64 = e ? (atoi(e) != 0) : 1;
  }
  if (!g_use_pair64 || C != 64 || !(k & 1) || k < 1 || k > 33) return false;
  // weights as packed for conv_tc with KB = 32: [cb = 2][tap][c8 = 4][hi|lo][64][8]
  if (!c1.ok || !c2.ok || c1.NC != 64 || c2.NC != 64 || c1.n_chunks != 1 || c2.n_chunks != 1 || c1.KB != 32 || c2.KB != 32 ||
      c1.n_cb != 2 || c2.n_cb != 2)
    return false;
  const int p1 = dil * (k - 1) / 2, p2 = (k - 1) / 2;
  if (p1 + p2 > kTcHalo || 128 - (k - 1) < 64) return false;
  const size_t R1 = 128 + (size_t)(k - 1) * dil, R2 = 128 + (k - 1);
  const size_t tiles = 2 * (2 * 8 * R1 * 16 + 2 * 8 * R2 * 16);
  const size_t fixed = tiles + 2 * 64 * 4 + 128;
  for (int ns = std::min(8, 2 * k); ns >= 2; --ns) {
    const size_t need = fixed + (size_t)ns * kPair64TapBytes + (size_t)(12 + 2 * ns) * 8;
    if (need <= kSmemPerSm - 1536) {
      L->k = k; L->dil = dil; L->NS = ns; L->smem = need; L->ok = true;
      return true;
    }
  }
  return false;
}

static int launch_pair64(Pair64Params p, const Pair64Layer& L, const TcLayer& c1, const TcLayer& c2, cudaStream_t st) {
  static bool attr_set[kMaxDevices] = {};   // the attribute is per device: one flag per device ordinal
  const int dev_ = current_device_slot();
  if (!attr_set[dev_]) {
    DISSC_CUDA(cudaFuncSetAttribute(resblock_pair64_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kSmemPerSm - 1024)));
    attr_set[dev_] = true;
  }
  p.k = L.k; p.dil = L.dil; p.NS = L.NS;
  if (p.in_scale == 0.f) p.in_scale = 1.f;
  if (p.xt_scale == 0.f) p.xt_scale = 1.f;
  if (p.plane_scale == 0.f) p.plane_scale = 1.f;
  p.w1 = c1.w; p.w2 = c2.w; p.inv1 = c1.inv_scale / p.in_scale; p.inv2 = c2.inv_scale / p.xt_scale;
  if (p.halo == 0) p.halo = kTcHalo;
  const int m_out = 128 - (L.k - 1);
  p.tiles_per_b = (p.T + m_out - 1) / m_out;
  p.n_tiles = p.B * p.tiles_per_b;
  const int grid = std::min(p.n_tiles, num_sms());
  DISSC_CUDA(launch_pdl(resblock_pair64_tc_kernel, grid, kPair64Threads, L.smem, st, p));
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

// ------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------
struct ConvLayer {
  int Cin = 0, Cout = 0, k = 0, dil = 1, pad = 0, co_tile = 0;
  float* w = nullptr;  // packed, device
  float* bias = nullptr;
};
struct ConvTLayer {
  int Cin = 0, Cout = 0, k = 0, u = 0, pad = 0, co_tile = 0;
  float* w = nullptr;
  float* bias = nullptr;
};

struct Profiler {
  char (*names)[64];
  float* ms;
  double* flops;
  double* bytes;  // algorithmic bytes of the launch (layer-fused traffic model)
  int cap;
  int n = 0;
  std::vector<cudaEvent_t> ev;
};

}  // namespace dissc

using namespace dissc;

struct dissc_gen {
  dissc_gen_cfg cfg;
  int device = 0;
  int hop = 1;
  std::vector<void*> allocs;
  ConvLayer pre, post;
  float* post_w_plain = nullptr;  // (Cin,k) for conv_post
  ConvTLayer ups[DISSC_MAX_STAGES];
  // rb[stage][j][m][0|1]  (ResBlock2 uses only [0])
  ConvLayer rb[DISSC_MAX_STAGES][DISSC_MAX_KERNELS][DISSC_MAX_DILATIONS][2];
  // tensor-core twins of rb[][][][] and per-stage eligibility
  TcLayer rb_tc[DISSC_MAX_STAGES][DISSC_MAX_KERNELS][DISSC_MAX_DILATIONS][2];
  bool stage_tc[DISSC_MAX_STAGES] = {};
  PairLayer rb_pair[DISSC_MAX_STAGES][DISSC_MAX_KERNELS][DISSC_MAX_DILATIONS];  // fused (c1,c2) pairs, narrow stages
  bool stage_pair[DISSC_MAX_STAGES] = {};
  dissc::Pack2Layer rb_pack2[DISSC_MAX_STAGES][DISSC_MAX_KERNELS][DISSC_MAX_DILATIONS];  // C = 16 pairs, two samples per row
  dissc::Pair64Layer rb_pair64[DISSC_MAX_STAGES][DISSC_MAX_KERNELS][DISSC_MAX_DILATIONS];  // fused pairs of the C = 64 stage
  bool stage_pair64[DISSC_MAX_STAGES] = {};
  TcLayer pre_tc, ups_tc[DISSC_MAX_STAGES];  // conv_pre / upsamplers on the tensor cores
  bool tc_all = false;                        // every layer but conv_post has a tcgen05 plan: planes flow end to end
  int use_tc = 1;
  float* dict_w = nullptr;
  float* spkr_w = nullptr;
  int n_launches = 0;
  int n_extra = 0;     // per-utterance conditioning channels after the speaker embedding (model_in_dim - the three standard parts)
  dissc::ErrFlag err;  // out-of-range unit / speaker ids (dissc_gen_status)
  // power-of-two activation scales of the split-fp16 planes (estimate_act_scales): embedding planes, conv_pre output,
  // per stage the residual stream x (x_up and every r) and the stage output, per conv pair the intermediate xt
  float s_emb = 1.f, s_pre = 1.f, s_x[DISSC_MAX_STAGES], s_out[DISSC_MAX_STAGES];
  float s_xt[DISSC_MAX_STAGES][DISSC_MAX_KERNELS][DISSC_MAX_DILATIONS];
  // host-entry staging arena
  void* arena = nullptr;
  size_t arena_bytes = 0;
  cudaStream_t hstream = nullptr;                  // compute stream of the host entry points
  cudaStream_t cstream = nullptr;                  // copy-back stream
  cudaEvent_t fwd_done[2] = {nullptr, nullptr};    // per slot: forward finished (compute stream)
  cudaEvent_t d2h_done[2] = {nullptr, nullptr};    // per slot: output is in host memory (copy stream)
};
constexpr int kHostSlots = 2;

namespace dissc {

static int dev_upload(dissc_gen* g, const float* host, size_t n, float** out) {
  float* d = nullptr;
  DISSC_CUDA(cudaMalloc(&d, std::max<size_t>(n, 4) * sizeof(float)));
  g->allocs.push_back(d);
  DISSC_CUDA(cudaMemcpy(d, host, n * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return DISSC_OK;
}

struct WeightMap {
  std::map<std::string, const dissc_tensor*> m;
  const dissc_tensor* get(const std::string& k) const {
    auto it = m.find(k);
    return it == m.end() ? nullptr : it->second;
  }
};

static int make_conv(dissc_gen* g, const WeightMap& wm, const std::string& prefix, int Cin, int Cout, int k, int dil,
                     ConvLayer* L) {
  const dissc_tensor* w = wm.get(prefix + ".weight");
  const dissc_tensor* b = wm.get(prefix + ".bias");
  DISSC_CHECK(w && b, DISSC_EMISSING, "missing tensor %s.{weight,bias}", prefix.c_str());
  DISSC_CHECK(w->numel == (int64_t)Cin * Cout * k, DISSC_EINVAL, "%s.weight has %lld elements, expected %d*%d*%d",
              prefix.c_str(), (long long)w->numel, Cout, Cin, k);
  DISSC_CHECK(b->numel == Cout, DISSC_EINVAL, "%s.bias has %lld elements, expected %d", prefix.c_str(),
              (long long)b->numel, Cout);
  DISSC_CHECK(conv_supported(k, dil), DISSC_EUNSUPPORTED, "%s: conv1d kernel_size=%d dilation=%d unsupported",
              prefix.c_str(), k, dil);
  L->Cin = Cin; L->Cout = Cout; L->k = k; L->dil = dil; L->pad = (k * dil - dil) / 2;
  L->co_tile = conv_co_tile(Cout);
  auto packed = pack_weights(w->data, Cin, Cout, k, L->co_tile, conv_ci_chunk(L->co_tile), false);
  int rc = dev_upload(g, packed.data(), packed.size(), &L->w);
  if (rc) return rc;
  return dev_upload(g, b->data, Cout, &L->bias);
}

static int tc_upload(dissc_gen* g, const std::vector<__half>& packed, __half** out) {
  __half* d = nullptr;
  DISSC_CUDA(cudaMalloc(&d, packed.size() * sizeof(__half)));
  g->allocs.push_back(d);
  DISSC_CUDA(cudaMemcpy(d, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
  *out = d;
  return DISSC_OK;
}

static int make_conv_tc(dissc_gen* g, const WeightMap& wm, const std::string& prefix, int Cin, int Cout, int k, int dil,
                        TcLayer* L) {
  if (!tc_plan_conv(Cin, Cout, k, dil, L)) return DISSC_OK;  // not eligible: the fp32 CUDA-core kernel handles it
  const float* w = wm.get(prefix + ".weight")->data;         // (Cout, Cin, k)
  auto packed = pack_weights_tc(*L, [=](int n, int ci, int j) { return w[((size_t)n * Cin + ci) * k + j]; },
                                &L->inv_scale);
  return tc_upload(g, packed, &L->w);
}

static int make_convt_tc(dissc_gen* g, const WeightMap& wm, const std::string& prefix, int Cin, int Cout, int k, int u,
                         TcLayer* L) {
  if (!tc_plan_convt(Cin, Cout, k, u, L)) return DISSC_OK;
  const float* w = wm.get(prefix + ".weight")->data;  // (Cin, Cout, k)
  const int M = L->k;
  auto packed = pack_weights_tc(*L, [=](int n, int ci, int jp) {
    const int phase = n / Cout, co = n % Cout;
    const int jj = phase + (M - 1 - jp) * u;
    return jj < k ? w[((size_t)ci * Cout + co) * k + jj] : 0.f;
  }, &L->inv_scale);
  return tc_upload(g, packed, &L->w);
}

static int make_convt(dissc_gen* g, const WeightMap& wm, const std::string& prefix, int Cin, int Cout, int k, int u,
                      ConvTLayer* L) {
  const dissc_tensor* w = wm.get(prefix + ".weight");
  const dissc_tensor* b = wm.get(prefix + ".bias");
  DISSC_CHECK(w && b, DISSC_EMISSING, "missing tensor %s.{weight,bias}", prefix.c_str());
  DISSC_CHECK(w->numel == (int64_t)Cin * Cout * k && b->numel == Cout, DISSC_EINVAL, "%s: bad tensor sizes",
              prefix.c_str());
  DISSC_CHECK(convt_supported(k, u), DISSC_EUNSUPPORTED, "%s: conv_transpose1d kernel_size=%d stride=%d unsupported",
              prefix.c_str(), k, u);
  L->Cin = Cin; L->Cout = Cout; L->k = k; L->u = u; L->pad = (k - u) / 2;
  L->co_tile = convt_co_tile(Cout);
  auto packed = pack_weights(w->data, Cin, Cout, k, L->co_tile, 8, true);
  int rc = dev_upload(g, packed.data(), packed.size(), &L->w);
  if (rc) return rc;
  return dev_upload(g, b->data, Cout, &L->bias);
}

// ------------------------------------------------------------------------
// Activation scale of the split-fp16 planes
// ------------------------------------------------------------------------
// A plane pair represents x as fp16 hi + fp16 lo.  That is fp32-accurate while |x| sits well inside the fp16 range:
// `lo` keeps all 11 bits for |x| >= 2^-3 and `hi` saturates at 65504.  The weights get their own power-of-two scale
// (pack_weights_tc); the activations get one per tensor, estimated HERE from the weights alone: a second-moment
// propagation through the graph (mean square of every tensor under independent zero-mean inputs: a conv multiplies it
// by sum(w^2) / Cout and adds mean(b^2); leaky-relu(0.1) keeps 0.505 of it; a residual add sums; the MRF mean keeps it),
// and each plane is scaled so its estimated RMS lands in [16, 32): values from RMS / 128 to RMS * 2047 are exact to 22
// bits.  The estimate only has to be right within a few octaves; being powers of two the scales change no result bit
// unless a plane would otherwise leave that window (tests/test_layers_gpu.py::*_activation_scale_sweep,
// tests/test_generator_gpu.py::test_config2_rows_vs_oracle on the N(0, 0.01) `init_weights` recipe, whose late
// activations are ~1e-5).  DISSC_ACT_SCALE=0 forces every scale to 1.
static float pow2_scale_for_rms(double rms) {
  if (!(rms > 0.0) || !std::isfinite(rms)) return 1.f;
  int e = 0;
  std::frexp(rms, &e);   // rms = f * 2^e, f in [0.5, 1)
  const int sh = std::max(-40, std::min(40, 5 - e));
  return std::ldexp(1.f, sh);
}
static double mean_sq(const dissc_tensor* t) {
  if (!t || t->numel <= 0) return 0.0;
  double a = 0.0;
  for (int64_t i = 0; i < t->numel; ++i) a += (double)t->data[i] * t->data[i];
  return a / (double)t->numel;
}
// second-moment gain of a conv: sum(w^2) / n_out_positions_per_weight_set, plus mean(b^2)
static void conv_moments(const WeightMap& wm, const std::string& prefix, int Cout, int stride, double* gain, double* b2) {
  const dissc_tensor* w = wm.get(prefix + ".weight");
  const dissc_tensor* b = wm.get(prefix + ".bias");
  *gain = w ? mean_sq(w) * (double)w->numel / ((double)Cout * stride) : 0.0;
  *b2 = mean_sq(b);
}
static void estimate_act_scales(dissc_gen* g, const WeightMap& wm) {
  const dissc_gen_cfg& c = g->cfg;
  g->s_emb = g->s_pre = 1.f;
  for (int i = 0; i < DISSC_MAX_STAGES; ++i) {
    g->s_x[i] = g->s_out[i] = 1.f;
    for (int j = 0; j < DISSC_MAX_KERNELS; ++j)
      for (int m = 0; m < DISSC_MAX_DILATIONS; ++m) g->s_xt[i][j][m] = 1.f;
  }
  if (const char* e = getenv("DISSC_ACT_SCALE"))
    if (atoi(e) == 0) return;
  if (!g->tc_all) return;   // mixed pipelines (a CUDA-core upsampler writing planes) keep unscaled planes
  constexpr double kLrelu = 0.505;   // E[lrelu(x, 0.1)^2] / E[x^2] for a symmetric x
  const double ms_dict = mean_sq(wm.get("dict.weight")), ms_spk = c.has_spkr ? mean_sq(wm.get("spkr.weight")) : 0.0;
  const double ms_f0 = c.has_f0 ? 1.0 : 0.0;   // speaker-normalised F0, sr/dataset.py:297-312
  // extra conditioning channels (`f0_stats`: a speaker's F0 mean / std in Hz, ~1e2) share the embedding planes: leave
  // them 2^7 of headroom on top of the table values
  const double ms_extra = g->n_extra > 0 ? 128.0 * 128.0 : 0.0;
  g->s_emb = pow2_scale_for_rms(std::sqrt(std::max(std::max(ms_dict, ms_extra), std::max(ms_spk, ms_f0))));
  double ms = (c.embedding_dim * ms_dict + ms_f0 + (c.has_spkr ? c.embedding_dim * ms_spk : 0.0)) / std::max(1, c.model_in_dim);
  double gain, b2;
  conv_moments(wm, "conv_pre", c.c0, 1, &gain, &b2);
  ms = ms * gain + b2;
  g->s_pre = pow2_scale_for_rms(std::sqrt(kLrelu * ms));
  for (int i = 0; i < c.n_up; ++i) {
    const int ch = c.c0 >> (i + 1);
    conv_moments(wm, "ups." + std::to_string(i), ch, c.up_rates[i], &gain, &b2);
    const double ms_up = kLrelu * ms * gain + b2;
    g->s_x[i] = pow2_scale_for_rms(std::sqrt(kLrelu * ms_up));
    double ms_sum = 0.0;
    for (int j = 0; j < c.n_rk; ++j) {
      const std::string p = "resblocks." + std::to_string(i * c.n_rk + j);
      double ms_r = ms_up;
      for (int m = 0; m < c.n_dil; ++m) {
        if (c.resblock == 1) {
          conv_moments(wm, p + ".convs1." + std::to_string(m), ch, 1, &gain, &b2);
          const double ms_xt = kLrelu * ms_r * gain + b2;
          g->s_xt[i][j][m] = pow2_scale_for_rms(std::sqrt(kLrelu * ms_xt));
          conv_moments(wm, p + ".convs2." + std::to_string(m), ch, 1, &gain, &b2);
          ms_r += kLrelu * ms_xt * gain + b2;
        } else {
          conv_moments(wm, p + ".convs." + std::to_string(m), ch, 1, &gain, &b2);
          ms_r += kLrelu * ms_r * gain + b2;
        }
      }
      ms_sum += ms_r;
    }
    ms = ms_sum / std::max(1, c.n_rk);
    g->s_out[i] = pow2_scale_for_rms(std::sqrt(kLrelu * ms));
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// bytes of one workspace region: large enough for any stage's tensor in any layout
// (plain (B,C,T) fp32, blocked f32b [B][C/8][Tr][8] fp32, or an fp16 hi+lo plane pair [B][C/8][Tp][8] x2),
// plus slack: the last frame tile of a transposed conv may read up to 128 rows past a slab (discarded rows).
static size_t region_bytes(const dissc_gen* g, int B, int T) {
  const size_t tp0 = round_up(T, 128) + 2 * kTcHalo;
  const size_t cin_pad = round_up(g->cfg.model_in_dim, 16);
  size_t mx = (size_t)B * std::max<size_t>(g->cfg.c0, cin_pad) * tp0 * 4;
  size_t t = T;
  for (int i = 0; i < g->cfg.n_up; ++i) {
    const ConvTLayer& U = g->ups[i];
    t = (t - 1) * U.u - 2 * U.pad + U.k;
    const size_t tp = round_up(t, 128) + std::max(2 * kTcHalo, kPairSlack);
    mx = std::max(mx, (size_t)B * round_up(U.Cout, 8) * tp * 4);
  }
  return round_up(mx, 1024) + 8192;
}
constexpr int kNumRegions = 8;

static int out_len(const dissc_gen* g, int T) {
  long t = T;
  for (int i = 0; i < g->cfg.n_up; ++i) t = (t - 1) * g->ups[i].u - 2 * g->ups[i].pad + g->ups[i].k;
  return (int)t;
}

struct Launcher {
  cudaStream_t st;
  Profiler* prof;
  int count = 0;
  int begin(const char* name, double flops, double bytes = 0.0) {
    ++count;
    if (!prof) return DISSC_OK;
    if (prof->n >= prof->cap) return set_err(DISSC_EINVAL, "profile capacity %d too small", prof->cap);
    snprintf(prof->names[prof->n], 64, "%s", name);
    prof->flops[prof->n] = flops;
    if (prof->bytes) prof->bytes[prof->n] = bytes;
    cudaEvent_t e0, e1;
    DISSC_CUDA(cudaEventCreate(&e0));
    DISSC_CUDA(cudaEventCreate(&e1));
    prof->ev.push_back(e0);
    prof->ev.push_back(e1);
    DISSC_CUDA(cudaEventRecord(e0, st));
    return DISSC_OK;
  }
  int end() {
    if (!prof) return DISSC_OK;
    DISSC_CUDA(cudaEventRecord(prof->ev.back(), st));
    prof->n++;
    return DISSC_OK;
  }
};

#define DISSC_TRY(expr)        \
  do {                         \
    int _rc = (expr);          \
    if (_rc != DISSC_OK) return _rc; \
  } while (0)

static int forward_impl(dissc_gen* g, const int64_t* code, const float* f0, const int64_t* spkr,
                        const int32_t* lengths, int B, int T, float* out_f32, int16_t* out_i16, void* workspace,
                        size_t workspace_bytes, cudaStream_t st, Profiler* prof, const float* extra = nullptr) {
  DISSC_CHECK(g && code && (out_f32 || out_i16), DISSC_EINVAL, "null handle / code / out");
  DISSC_CHECK(g->n_extra == 0 || extra, DISSC_EINVAL,
              "the config has %d extra conditioning channels (model_in_dim %d): use dissc_gen_forward_ex with `extra`",
              g->n_extra, g->cfg.model_in_dim);
  DISSC_CHECK(B > 0 && T > 0, DISSC_EINVAL, "B=%d T=%d must be positive", B, T);
  DISSC_CHECK(!g->cfg.has_f0 || f0, DISSC_EINVAL, "config has f0 but f0 == NULL");
  DISSC_CHECK(!g->cfg.has_spkr || spkr, DISSC_EINVAL, "config is multi-speaker but spkr == NULL");
  DISSC_CHECK(B <= 65535, DISSC_EINVAL, "B=%d exceeds gridDim.z limit 65535", B);
  size_t need = 0;
  dissc_gen_workspace_bytes(g, B, T, &need);
  DISSC_CHECK(workspace && workspace_bytes >= need, DISSC_EINVAL, "workspace %zu bytes < required %zu", workspace_bytes,
              need);
  int dev = -1;
  DISSC_CUDA(cudaGetDevice(&dev));
  DISSC_CHECK(dev == g->device, DISSC_EINVAL, "current device %d != handle device %d", dev, g->device);
  DISSC_TRY(err_flag_take(&g->err, "an earlier forward", g->cfg.num_embeddings, g->cfg.n_spkr_rows));

  const size_t RS = region_bytes(g, B, T);
  char* wsb = static_cast<char*>(workspace);
  auto region = [&](int i) { return wsb + (size_t)i * RS; };
  float* act[2] = {reinterpret_cast<float*>(region(0)), reinterpret_cast<float*>(region(1))};
  // CUDA-core stages: x_up / xt / r are plain (B,C,T) fp32 in regions 2,3,4.
  // Tensor-core stages: F_up, F_r, F_xs (blocked fp32) in 2,4,5; plane pairs P_xt, P_up, P_r in 3,6,7.
  float* x_up = reinterpret_cast<float*>(region(2));
  float* xt = reinterpret_cast<float*>(region(3));
  float* r = reinterpret_cast<float*>(region(4));
  float* F_up = x_up;
  float* F_r = r;
  float* F_xs = reinterpret_cast<float*>(region(5));
  Launcher L{st, prof};
  const dissc_gen_cfg& c = g->cfg;
  char name[64];
  const bool tc_all = g->use_tc && g->tc_all;
  struct Planes { __half* hi; __half* lo; };
  auto planes = [&](int i) { return Planes{reinterpret_cast<__half*>(region(i)), reinterpret_cast<__half*>(region(i) + RS / 2)}; };
  // Tensor-core stages: F_up, F_r, F_xs (blocked fp32) in regions 2,4,5; plane pairs P_xt, P_up, P_r in 3,6,7;
  // with tc_all the stage inputs (leaky-relu'd planes) ping-pong in regions 0,1 and the embedding planes use region 2.
  const Planes P_xt = planes(3), P_up = planes(6), P_r = planes(7);
  const Planes P_act[2] = {planes(0), planes(1)};
  const int Tr0 = (int)round_up(T, 128), Tp0 = Tr0 + 2 * kTcHalo;

  if (tc_all) {
    // gather/concat -> split planes (zero halos written by the same kernel)   (sr/models.py:189,:206-215)
    const Planes P_emb = planes(2);
    EmbedParams e{};
    e.code = reinterpret_cast<const long long*>(code); e.f0 = f0; e.spkr = reinterpret_cast<const long long*>(spkr);
    e.dict_w = g->dict_w; e.spkr_w = g->spkr_w; e.lengths = lengths;
    e.E = c.embedding_dim; e.f0_ch = c.has_f0 ? c.embedding_dim : -1;
    e.spk_base = c.has_spkr ? c.embedding_dim + (c.has_f0 ? 1 : 0) : -1;
    e.Cin = c.model_in_dim; e.B = B; e.C8 = g->pre_tc.Cin_pad / 8; e.T = T; e.Tp = Tp0;
    e.extra = extra; e.n_extra = g->n_extra; e.extra_base = c.model_in_dim - g->n_extra;
    e.hi = P_emb.hi; e.lo = P_emb.lo; e.scale = g->s_emb;
    e.n_code_rows = c.num_embeddings; e.n_spkr_rows = c.n_spkr_rows; e.err = g->err.dev;
    DISSC_TRY(L.begin("embed", 0));
    const long long tot = (long long)B * e.C8 * Tp0;
    tc_embed_planes_kernel<<<(int)std::min<long long>((tot + 255) / 256, 148 * 8), 256, 0, st>>>(e);
    DISSC_CUDA(cudaGetLastError());
    DISSC_TRY(launch_zero_halos(P_act[0].hi, P_act[0].lo, B * c.c0 / 8, Tp0, T, st));
    DISSC_TRY(L.end());
    L.count += 1;
    // conv_pre; epilogue applies the first stage's leaky-relu (sr/models.py:99,:101) and writes planes
    TcParams p{};
    p.a_hi = P_emb.hi; p.a_lo = P_emb.lo; p.bias = g->pre.bias;
    p.out_hi = P_act[0].hi; p.out_lo = P_act[0].lo; p.plane_act = 1; p.plane_slope = 0.1f;
    p.in_scale = g->s_emb; p.plane_scale = g->s_pre;
    p.lengths = lengths; p.len_mul = 1;
    p.B = B; p.T = T; p.Tr = Tr0; p.Tp = Tp0; p.Tp_in = Tp0;
    DISSC_TRY(L.begin("conv_pre.tc", 2.0 * g->pre.Cin * g->pre.Cout * g->pre.k * (double)T * B,
                      (8.0 + (c.has_f0 ? 4.0 : 0.0)) * T * B + (c.has_spkr ? 8.0 * B : 0.0) + 4.0 * g->pre.Cout * (double)T * B + 4.0 * g->pre.Cin * g->pre.Cout * g->pre.k));
    DISSC_TRY(launch_conv_tc(p, g->pre_tc, T, st));
    DISSC_TRY(L.end());
  } else {
    // conv_pre with fused gather/concat; epilogue applies the first stage's leaky-relu (sr/models.py:99,:101)
    ConvParams p{};
    p.code = reinterpret_cast<const long long*>(code);
    p.f0 = f0;
    p.spkr = reinterpret_cast<const long long*>(spkr);
    p.dict_w = g->dict_w;
    p.spkr_w = g->spkr_w;
    p.E = c.embedding_dim;
    p.f0_ch = c.has_f0 ? c.embedding_dim : -1;
    p.spk_base = c.has_spkr ? c.embedding_dim + (c.has_f0 ? 1 : 0) : -1;
    p.n_code_rows = c.num_embeddings; p.n_spkr_rows = c.n_spkr_rows; p.err = g->err.dev;
    p.extra = extra; p.n_extra = g->n_extra; p.extra_base = c.model_in_dim - g->n_extra;
    p.w = g->pre.w; p.bias = g->pre.bias; p.out = act[0];
    p.lengths = lengths; p.len_mul = 1;
    p.B = B; p.Cin = g->pre.Cin; p.Cout = g->pre.Cout; p.T = T; p.pad = g->pre.pad;
    p.post_act = 1; p.post_slope = 0.1f;
    if (c.n_up == 0) { p.post_slope = 0.01f; }
    DISSC_TRY(L.begin("conv_pre", 2.0 * p.Cin * p.Cout * g->pre.k * (double)T * B,
                      (8.0 + (c.has_f0 ? 4.0 : 0.0)) * T * B + (c.has_spkr ? 8.0 * B : 0.0) + 4.0 * g->pre.Cout * (double)T * B + 4.0 * g->pre.Cin * g->pre.Cout * g->pre.k));
    DISSC_TRY(launch_conv(p, g->pre.k, 1, g->pre.co_tile, true, st));
    DISSC_TRY(L.end());
  }

  int cur = 0;     // act[cur] / P_act[cur] holds the (already activated) stage input
  int Tcur = T;    // time steps at the current rate
  int mul = 1;     // valid length multiplier (product of rates so far)
  for (int i = 0; i < c.n_up; ++i) {
    const ConvTLayer& U = g->ups[i];
    const int Tin = Tcur;
    const int Tout = (Tcur - 1) * U.u - 2 * U.pad + U.k;
    const bool tc = g->use_tc && g->stage_tc[i];
    const int Tr_in = (int)round_up(Tin, 128), Tp_in = Tr_in + 2 * kTcHalo;
    const int Tr = (int)round_up(Tout, 128), Tp = Tr + 2 * kTcHalo;
    const int ch = U.Cout;
    const bool last_stage = (i == c.n_up - 1);
    const bool pair = tc_all && g->stage_pair[i];            // fused (c1,c2) pairs on fp32 "f32h" tensors
    const bool pair64 = tc_all && g->stage_pair64[i];        // fused (c1,c2) pairs on planes (C = 64)
    const int Tpf = Tr + kPairSlack;
    if (tc) {
      // zero padding of the plane pairs this stage writes (halo rows + round-up rows)
      DISSC_TRY(L.begin("zero_halos", 0));
      if (!pair) {
        DISSC_TRY(launch_zero_halos(P_up.hi, P_up.lo, B * ch / 8, Tp, Tout, st));
        DISSC_TRY(launch_zero_halos(P_xt.hi, P_xt.lo, B * ch / 8, Tp, Tout, st));
        DISSC_TRY(launch_zero_halos(P_r.hi, P_r.lo, B * ch / 8, Tp, Tout, st));
        L.count += 2;
      } else if (last_stage) {
        --L.count;  // nothing to clear: begin() counted a launch that does not happen
      }
      if (tc_all && !last_stage) {
        DISSC_TRY(launch_zero_halos(P_act[cur ^ 1].hi, P_act[cur ^ 1].lo, B * ch / 8, Tp, Tout, st));
        L.count += 1;
      }
      DISSC_TRY(L.end());
    }
    snprintf(name, sizeof(name), tc_all ? "ups.%d.tc" : "ups.%d", i);
    DISSC_TRY(L.begin(name, 2.0 * U.Cin * U.Cout * U.k * (double)Tcur * B,
                      4.0 * B * ((double)U.Cin * Tin + (double)U.Cout * Tout) + 4.0 * U.Cin * U.Cout * U.k));
    if (tc_all) {
      // polyphase transposed conv on the tensor cores: frames x (phase, channel)
      TcParams p{};
      p.a_hi = P_act[cur].hi; p.a_lo = P_act[cur].lo; p.bias = U.bias;
      p.out_f32b = F_up; p.out_hi = P_up.hi; p.out_lo = P_up.lo; p.plane_act = 1; p.plane_slope = 0.1f;
      p.in_scale = (i == 0) ? g->s_pre : g->s_out[i - 1]; p.plane_scale = g->s_x[i];
      p.lengths = lengths; p.len_mul = mul * U.u;
      p.B = B; p.T = Tout; p.Tr = Tr; p.Tp = Tp; p.Tp_in = Tp_in;
      if (pair) {  // the fused pairs read x as fp32 only
        p.out_hi = nullptr; p.out_lo = nullptr; p.Tr = Tpf; p.f_halo = kPairHalo;
      }
      if (pair64) p.out_f32b = nullptr;  // the C = 64 fused pairs read (and rebuild the residual from) the planes only
      const int n_frames = (Tout + U.pad - 1) / U.u + 1;
      DISSC_TRY(launch_conv_tc(p, g->ups_tc[i], n_frames, st));
    } else {
      ConvTParams p{};
      p.in = act[cur]; p.w = U.w; p.bias = U.bias; p.out = x_up;
      p.lengths = lengths; p.len_mul = mul;
      p.B = B; p.Cin = U.Cin; p.Cout = U.Cout; p.Tin = Tcur; p.Tout = Tout; p.pad = U.pad;
      if (tc) {
        p.out = nullptr; p.out_f32b = F_up; p.out_hi = P_up.hi; p.out_lo = P_up.lo;
        p.Tr = Tr; p.Tp = Tp; p.halo = kTcHalo; p.plane_slope = 0.1f;
      }
      DISSC_TRY(launch_convt(p, U.k, U.u, U.co_tile, st));
    }
    DISSC_TRY(L.end());
    Tcur = Tout;
    mul *= U.u;
    float* xs = act[cur ^ 1];  // MRF accumulator (CUDA-core path) / next stage's activated input
    const float next_slope = last_stage ? 0.01f : 0.1f;  // sr/models.py:110 vs :101
    for (int j = 0; j < c.n_rk; ++j) {
      for (int m = 0; m < c.n_dil; ++m) {
        const bool last_m = (m == c.n_dil - 1);
        const bool last_j = (j == c.n_rk - 1);
        const ConvLayer& c1 = g->rb[i][j][m][0];
        const double fl = 2.0 * ch * ch * (double)c1.k * Tcur * B;
        const double S = 4.0 * ch * (double)Tcur * B, wb = 4.0 * ch * ch * (double)c1.k;
        const double by1 = 2 * S + wb;                                                       // K3: 1R + 1W
        const double by2 = 3 * S + wb + ((m == c.n_dil - 1 && j > 0) ? S : 0.0);             // K4: 2R + 1W (+ xs read)
        if (pair) {
          // K3+K4 fused: x' = x + conv2(lrelu(conv1(lrelu(x))))  [-> MRF accumulate]   (sr/models.py:36-40, :104-109)
          float* F_rr[2] = {F_r, xt};  // ping-pong (a tile reads halo rows its neighbours write)
          PairParams q{};
          q.x = (m == 0) ? F_up : F_rr[(m - 1) & 1];
          q.b1 = c1.bias; q.b2 = g->rb[i][j][m][1].bias;
          q.lengths = lengths; q.len_mul = mul;
          q.B = B; q.T = Tcur; q.Tpf = Tpf; q.f_halo = kPairHalo; q.Tp = Tp; q.p_halo = kTcHalo;
          q.in_scale = g->s_x[i]; q.xt_scale = g->s_xt[i][j][m]; q.plane_scale = g->s_out[i];
          if (!last_m) {
            q.out_f = F_rr[m & 1];
          } else {
            if (j > 0) q.acc_in = F_xs;
            if (!last_j) {
              q.out_f = F_xs;
            } else {
              q.div = (float)c.n_rk;
              if (!last_stage) {
                q.out_hi = P_act[cur ^ 1].hi; q.out_lo = P_act[cur ^ 1].lo; q.plane_act = 1; q.plane_slope = next_slope;
              } else {
                q.out_plain = xs; q.plain_act = 1; q.plain_slope = next_slope;
              }
            }
          }
          if (g->rb_pack2[i][j][m].ok) {   // C = 16: two samples per GEMM row (resblock_pack2_tc.cuh), same tensors
            Pack2Params r{};
            r.x = q.x; r.b1 = q.b1; r.b2 = q.b2; r.acc_in = q.acc_in; r.out_f = q.out_f; r.out_hi = q.out_hi;
            r.out_lo = q.out_lo; r.out_plain = q.out_plain; r.lengths = q.lengths; r.len_mul = q.len_mul;
            r.B = q.B; r.T = q.T; r.Tpf = q.Tpf; r.f_halo = q.f_halo; r.Tp = q.Tp; r.p_halo = q.p_halo;
            r.in_scale = q.in_scale; r.xt_scale = q.xt_scale; r.plane_scale = q.plane_scale;
            r.div = q.div; r.plane_act = q.plane_act; r.plain_act = q.plain_act; r.plane_slope = q.plane_slope;
            r.plain_slope = q.plain_slope;
            snprintf(name, sizeof(name), "s%d.rb%d.pair.%d.pk2", i, j, m);
            DISSC_TRY(L.begin(name, 2 * fl, by1 + by2));
            DISSC_TRY(launch_pack2(r, g->rb_pack2[i][j][m], st));
            DISSC_TRY(L.end());
            continue;
          }
          snprintf(name, sizeof(name), "s%d.rb%d.pair.%d.ptc", i, j, m);
          DISSC_TRY(L.begin(name, 2 * fl, by1 + by2));
          DISSC_TRY(launch_pair(q, g->rb_pair[i][j][m], g->rb_tc[i][j][m][0], g->rb_tc[i][j][m][1], st));
          DISSC_TRY(L.end());
          continue;
        }
        if (pair64) {
          // K3+K4 fused, planes in / planes out; the residual is rebuilt from the input planes (resblock64_tc.cuh)
          const Planes PP[2] = {P_r, P_xt};  // ping-pong (a tile reads halo rows its neighbours write)
          const Planes pin = (m == 0) ? P_up : PP[(m - 1) & 1];
          Pair64Params q{};
          q.x_hi = pin.hi; q.x_lo = pin.lo; q.in_inv_slope = 10.0f;  // planes hold lrelu(x, 0.1)
          q.b1 = c1.bias; q.b2 = g->rb[i][j][m][1].bias;
          q.lengths = lengths; q.len_mul = mul;
          q.B = B; q.T = Tcur; q.Tp = Tp; q.Tr = Tr; q.halo = kTcHalo;
          q.plane_slope = 0.1f;
          q.in_scale = g->s_x[i]; q.xt_scale = g->s_xt[i][j][m]; q.plane_scale = g->s_x[i];
          if (!last_m) {
            q.out_hi = PP[m & 1].hi; q.out_lo = PP[m & 1].lo;
          } else {
            if (j > 0) q.acc_in = F_xs;
            if (!last_j) {
              q.out_f = F_xs;
            } else {
              q.div = (float)c.n_rk;
              q.out_hi = P_act[cur ^ 1].hi; q.out_lo = P_act[cur ^ 1].lo; q.plane_slope = next_slope;
              q.plane_scale = g->s_out[i];
            }
          }
          snprintf(name, sizeof(name), "s%d.rb%d.pair.%d.p64", i, j, m);
          DISSC_TRY(L.begin(name, 2 * fl, by1 + by2));
          DISSC_TRY(launch_pair64(q, g->rb_pair64[i][j][m], g->rb_tc[i][j][m][0], g->rb_tc[i][j][m][1], st));
          DISSC_TRY(L.end());
          continue;
        }
        if (tc) {
          const Planes rin_p = (m == 0) ? P_up : P_r;
          const float* rin_f = (m == 0) ? F_up : F_r;
          TcParams base{};
          base.lengths = lengths; base.len_mul = mul;
          base.B = B; base.T = Tcur; base.Tr = Tr; base.Tp = Tp; base.Tp_in = Tp;
          if (c.resblock == 1) {
            // K3: (leaky-relu'd planes) -> dilated conv -> leaky-relu -> planes   (sr/models.py:36-38)
            TcParams p = base;
            p.a_hi = rin_p.hi; p.a_lo = rin_p.lo; p.bias = c1.bias;
            p.out_hi = P_xt.hi; p.out_lo = P_xt.lo; p.plane_act = 1; p.plane_slope = 0.1f;
            p.in_scale = g->s_x[i]; p.plane_scale = g->s_xt[i][j][m];
            snprintf(name, sizeof(name), "s%d.rb%d.c1.%d.tc", i, j, m);
            DISSC_TRY(L.begin(name, fl, by1));
            DISSC_TRY(launch_conv_tc(p, g->rb_tc[i][j][m][0], Tcur, st));
            DISSC_TRY(L.end());
          }
          // K4: conv -> + residual [-> MRF accumulate]   (:39-40, :104-109)
          const int which = (c.resblock == 1) ? 1 : 0;
          TcParams q = base;
          const Planes qin = (c.resblock == 1) ? P_xt : rin_p;
          q.a_hi = qin.hi; q.a_lo = qin.lo; q.bias = g->rb[i][j][m][which].bias; q.res = rin_f;
          q.in_scale = (c.resblock == 1) ? g->s_xt[i][j][m] : g->s_x[i];
          if (!last_m) {
            q.out_f32b = F_r; q.out_hi = P_r.hi; q.out_lo = P_r.lo; q.plane_act = 1; q.plane_slope = 0.1f;
            q.plane_scale = g->s_x[i];
          } else {
            if (j > 0) q.acc_in = F_xs;
            if (!last_j) {
              q.out_f32b = F_xs;
            } else {
              q.div = (float)c.n_rk;
              if (tc_all && !last_stage) {
                q.out_hi = P_act[cur ^ 1].hi; q.out_lo = P_act[cur ^ 1].lo; q.plane_act = 1; q.plane_slope = next_slope;
                q.plane_scale = g->s_out[i];
              } else {
                q.out_plain = xs; q.plain_act = 1; q.plain_slope = next_slope;
              }
            }
          }
          snprintf(name, sizeof(name), "s%d.rb%d.c2.%d.tc", i, j, m);
          DISSC_TRY(L.begin(name, fl, by2));
          DISSC_TRY(launch_conv_tc(q, g->rb_tc[i][j][m][which], Tcur, st));
          DISSC_TRY(L.end());
          continue;
        }
        const float* rin = (m == 0) ? x_up : r;
        ConvParams p{};
        p.lengths = lengths; p.len_mul = mul;
        p.B = B; p.Cin = ch; p.Cout = ch; p.T = Tcur;
        if (c.resblock == 1) {
          // K3: lrelu -> dilated conv -> lrelu     (sr/models.py:36-38)
          p.in = rin; p.w = c1.w; p.bias = c1.bias; p.out = xt; p.pad = c1.pad;
          p.pre_act = 1; p.pre_slope = 0.1f; p.post_act = 1; p.post_slope = 0.1f;
          snprintf(name, sizeof(name), "s%d.rb%d.c1.%d", i, j, m);
          DISSC_TRY(L.begin(name, fl, by1));
          DISSC_TRY(launch_conv(p, c1.k, c1.dil, c1.co_tile, false, st));
          DISSC_TRY(L.end());
        }
        // K4: conv -> + residual [-> MRF accumulate]  (:39-40, :104-109)
        const ConvLayer& c2 = (c.resblock == 1) ? g->rb[i][j][m][1] : c1;
        ConvParams q{};
        q.lengths = lengths; q.len_mul = mul;
        q.B = B; q.Cin = ch; q.Cout = ch; q.T = Tcur;
        q.w = c2.w; q.bias = c2.bias; q.pad = c2.pad; q.res = rin;
        if (c.resblock == 1) {
          q.in = xt;
        } else {
          q.in = rin; q.pre_act = 1; q.pre_slope = 0.1f;  // ResBlock2: lrelu -> conv -> +x (:63-66)
        }
        if (!last_m) {
          q.out = r;
        } else {
          q.out = xs;
          if (j > 0) q.acc_in = xs;
          if (last_j) {
            q.div = (float)c.n_rk;
            q.post_act = 1; q.post_slope = next_slope;
          }
        }
        snprintf(name, sizeof(name), "s%d.rb%d.c2.%d", i, j, m);
        DISSC_TRY(L.begin(name, fl, by2));
        DISSC_TRY(launch_conv(q, c2.k, c2.dil, c2.co_tile, false, st));
        DISSC_TRY(L.end());
      }
    }
    cur ^= 1;
  }

  {
    ConvPostParams p{};
    p.in = act[cur]; p.w = g->post_w_plain; p.bias = g->post.bias;
    p.out_f32 = out_f32; p.out_i16 = out_i16;
    p.lengths = lengths; p.len_mul = mul;
    p.B = B; p.Cin = g->post.Cin; p.T = Tcur;
    DISSC_TRY(L.begin("conv_post", 2.0 * p.Cin * g->post.k * (double)Tcur * B,
                      4.0 * B * ((double)p.Cin * Tcur + Tcur) + 4.0 * p.Cin * g->post.k));
    DISSC_TRY(launch_conv_post(p, g->post.k, st));
    DISSC_TRY(L.end());
  }
  g->n_launches = L.count;
  return DISSC_OK;
}

// Layer-test helpers: the test entry points pick the activation scales from the tensors they are handed (absmax -> a
// power of two), which a model does at load time from its weights (estimate_act_scales).  Synchronous; test entries only.
static double device_absmax(const float* d, size_t n, cudaStream_t st) {
  if (!d || !n) return 0.0;
  std::vector<float> h(n);
  cudaStreamSynchronize(st);
  if (cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 0.0;
  double m = 0.0;
  for (float v : h)
    if (std::isfinite(v)) m = std::max(m, (double)std::fabs(v));
  return m;
}
static double host_absmax(const float* h, size_t n) {
  double m = 0.0;
  for (size_t i = 0; h && i < n; ++i) m = std::max(m, (double)std::fabs(h[i]));
  return m;
}
// largest l2 norm of one output channel's weights: the typical |output| of a conv is this times the input's RMS
static double host_row_norm(const float* w, int rows, size_t per_row) {
  double m = 0.0;
  for (int r = 0; w && r < rows; ++r) {
    double a = 0.0;
    for (size_t i = 0; i < per_row; ++i) a += (double)w[r * per_row + i] * w[r * per_row + i];
    m = std::max(m, std::sqrt(a));
  }
  return m;
}
// scale that puts `absmax` at ~2^9 (the RMS of a Gaussian-ish tensor then sits near 2^7; 2^7 headroom to 65504)
static float pow2_scale_for_absmax(double absmax) { return pow2_scale_for_rms(absmax / 32.0); }

// Layer-level test of the C = 64 fused pair: plain (B,64,T) fp32 in, raw x' [+acc][/div] and leaky-relu planes out.
static int pair64_layer_test(const float* in, const float* w1_host, const float* b1_host, const float* w2_host,
                             const float* b2_host, const float* acc_in, float* out_raw, float* out_planes,
                             const int32_t* lengths, int len_mul, int B, int T, int k, int dilation, float div,
                             float plane_slope, cudaStream_t st) {
  const int C = 64;
  TcLayer c1, c2;
  Pair64Layer L;
  DISSC_CHECK(tc_plan_conv(C, C, k, dilation, &c1) && tc_plan_conv(C, C, k, 1, &c2) && pair64_plan(C, k, dilation, c1, c2, &L),
              DISSC_EUNSUPPORTED, "no fused-pair plan for C=64 kernel_size=%d dilation=%d (odd k, halo <= %d)", k, dilation,
              kTcHalo);
  const int Tr = (int)round_up(T, 128), Tp = Tr + 2 * kTcHalo;
  const size_t f_elems = (size_t)B * (C / 8) * Tr * 8, plane_elems = (size_t)B * (C / 8) * Tp * 8 + 4096;
  auto pk1 = pack_weights_tc(c1, [=](int n, int ci, int j) { return w1_host[((size_t)n * C + ci) * k + j]; }, &c1.inv_scale);
  auto pk2 = pack_weights_tc(c2, [=](int n, int ci, int j) { return w2_host[((size_t)n * C + ci) * k + j]; }, &c2.inv_scale);
  std::vector<void*> tmp;
  auto dalloc = [&](size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    tmp.push_back(d);
    return d;
  };
  auto cleanup = [&]() { for (void* d : tmp) cudaFree(d); };
  __half* dw1 = (__half*)dalloc(pk1.size() * 2);
  __half* dw2 = (__half*)dalloc(pk2.size() * 2);
  float* db = (float*)dalloc((size_t)2 * C * 4);
  __half* i_hi = (__half*)dalloc(plane_elems * 2);
  __half* i_lo = (__half*)dalloc(plane_elems * 2);
  float* f_acc = (float*)dalloc(f_elems * 4);
  float* f_out = (float*)dalloc(f_elems * 4);
  __half* o_hi = (__half*)dalloc(plane_elems * 2);
  __half* o_lo = (__half*)dalloc(plane_elems * 2);
  if (!dw1 || !dw2 || !db || !i_hi || !i_lo || !f_acc || !f_out || !o_hi || !o_lo) {
    cleanup();
    return set_err(DISSC_ENOMEM, "cudaMalloc failed in dissc_resblock_pair_tc");
  }
  cudaMemcpyAsync(dw1, pk1.data(), pk1.size() * 2, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dw2, pk2.data(), pk2.size() * 2, cudaMemcpyHostToDevice, st);
  std::vector<float> bias(2 * C, 0.f);
  if (b1_host) memcpy(bias.data(), b1_host, C * 4);
  if (b2_host) memcpy(bias.data() + C, b2_host, C * 4);
  cudaMemcpyAsync(db, bias.data(), (size_t)2 * C * 4, cudaMemcpyHostToDevice, st);
  // garbage everywhere first: the kernel must not depend on anything but the zeroed halos / rows past the valid length
  cudaMemsetAsync(i_hi, 0x7b, plane_elems * 2, st);
  cudaMemsetAsync(i_lo, 0x7b, plane_elems * 2, st);
  cudaMemsetAsync(f_acc, 0xff, f_elems * 4, st);
  cudaMemsetAsync(f_out, 0xff, f_elems * 4, st);
  cudaMemsetAsync(o_hi, 0x7b, plane_elems * 2, st);
  cudaMemsetAsync(o_lo, 0x7b, plane_elems * 2, st);
  const int nb = 148 * 4;
  const double ax = device_absmax(in, (size_t)B * C * T, st), aacc = device_absmax(acc_in, acc_in ? (size_t)B * C * T : 0, st);
  const double axt = ax * host_row_norm(w1_host, C, (size_t)C * k) + host_absmax(b1_host, C);
  const double aout = ax + axt * host_row_norm(w2_host, C, (size_t)C * k) + host_absmax(b2_host, C) + aacc;
  const float s_in = pow2_scale_for_absmax(ax), s_xt = pow2_scale_for_absmax(axt), s_out = pow2_scale_for_absmax(aout);
  tc_pack_planes_kernel<<<nb, 256, 0, st>>>(in, i_hi, i_lo, lengths, len_mul, B, C, C / 8, T, Tp, 1, 0.1f, s_in);
  int rc = launch_zero_halos(i_hi, i_lo, B * C / 8, Tp, T, st);
  if (!rc) rc = launch_zero_halos(o_hi, o_lo, B * C / 8, Tp, T, st);
  if (acc_in) tc_plain_to_f32b_kernel<<<nb, 256, 0, st>>>(acc_in, f_acc, B, C, T, Tr);
  c1.w = dw1; c2.w = dw2;
  Pair64Params p{};
  p.x_hi = i_hi; p.x_lo = i_lo; p.in_inv_slope = 10.0f;
  p.in_scale = s_in; p.xt_scale = s_xt; p.plane_scale = s_out;
  p.b1 = db; p.b2 = db + C; p.acc_in = acc_in ? f_acc : nullptr;
  p.out_f = f_out; p.out_hi = out_planes ? o_hi : nullptr; p.out_lo = out_planes ? o_lo : nullptr;
  p.lengths = lengths; p.len_mul = len_mul;
  p.B = B; p.T = T; p.Tp = Tp; p.Tr = Tr; p.halo = kTcHalo;
  p.div = div; p.plane_slope = plane_slope;
  if (!rc) rc = launch_pair64(p, L, c1, c2, st);
  if (!rc && out_raw) tc_f32b_to_plain_kernel<<<nb, 256, 0, st>>>(f_out, out_raw, B, C, T, Tr);
  if (!rc && out_planes) tc_planes_to_plain_kernel<<<nb, 256, 0, st>>>(o_hi, o_lo, out_planes, B, C, T, Tp, 1.f / s_out);
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (rc) return rc;
  DISSC_CUDA(e);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

}  // namespace dissc

// ------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------
extern "C" {

const char* dissc_last_error(void) { return dissc::g_err; }
const char* dissc_version(void) { return "dissc_b200 0.1 (sm_100a)"; }

int dissc_gen_create(dissc_gen_t** out, const dissc_gen_cfg* cfg, const dissc_tensor* weights, int n_weights,
                     int device) {
  DISSC_CHECK(out && cfg && weights, DISSC_EINVAL, "null argument");
  *out = nullptr;
  const dissc_gen_cfg& c = *cfg;
  DISSC_CHECK(c.n_up >= 0 && c.n_up <= DISSC_MAX_STAGES && c.n_rk > 0 && c.n_rk <= DISSC_MAX_KERNELS && c.n_dil > 0 &&
                  c.n_dil <= DISSC_MAX_DILATIONS,
              DISSC_EINVAL, "bad stage/kernel counts");
  DISSC_CHECK(c.resblock == 1 || c.resblock == 2, DISSC_EUNSUPPORTED, "resblock must be \"1\" or \"2\"");
  const int expect_in = c.embedding_dim + (c.has_f0 ? 1 : 0) + (c.has_spkr ? c.embedding_dim : 0);
  DISSC_CHECK(c.model_in_dim >= expect_in, DISSC_EUNSUPPORTED,
              "model_in_dim=%d but embedding_dim/f0/multispkr alone give %d channels", c.model_in_dim, expect_in);
  struct DeviceGuard {   // the caller's current device is restored on every exit path
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  } guard;
  cudaGetDevice(&guard.prev);
  DISSC_CUDA(cudaSetDevice(device));
  WeightMap wm;
  for (int i = 0; i < n_weights; ++i) wm.m[weights[i].name] = &weights[i];

  dissc_gen* g = new dissc_gen();
  g->cfg = c;
  g->device = device;
  g->n_extra = c.model_in_dim - expect_in;   // e.g. 2 for `f0_feats` (the f0_stats mean / std channels, sr/models.py:216-221)
  auto fail = [&](int rc) {
    dissc_gen_destroy(g);
    return rc;
  };
  int rc;
  if ((rc = err_flag_create(&g->err))) return fail(rc);
  if ((rc = make_conv(g, wm, "conv_pre", c.model_in_dim, c.c0, 7, 1, &g->pre))) return fail(rc);
  int ch = c.c0;
  g->hop = 1;
  for (int i = 0; i < c.n_up; ++i) {
    const int ci = c.c0 >> i, co = c.c0 >> (i + 1);
    if (co < 1) return fail(set_err(DISSC_EINVAL, "upsample_initial_channel too small for %d stages", c.n_up));
    if ((rc = make_convt(g, wm, "ups." + std::to_string(i), ci, co, c.up_kernels[i], c.up_rates[i], &g->ups[i])))
      return fail(rc);
    ch = co;
    g->hop *= c.up_rates[i];
    for (int j = 0; j < c.n_rk; ++j)
      for (int m = 0; m < c.n_dil; ++m) {
        const std::string p = "resblocks." + std::to_string(i * c.n_rk + j);
        if (c.resblock == 1) {
          if ((rc = make_conv(g, wm, p + ".convs1." + std::to_string(m), ch, ch, c.rk[j], c.dil[j][m],
                              &g->rb[i][j][m][0])))
            return fail(rc);
          if ((rc = make_conv(g, wm, p + ".convs2." + std::to_string(m), ch, ch, c.rk[j], 1, &g->rb[i][j][m][1])))
            return fail(rc);
          if ((rc = make_conv_tc(g, wm, p + ".convs1." + std::to_string(m), ch, ch, c.rk[j], c.dil[j][m],
                                 &g->rb_tc[i][j][m][0])))
            return fail(rc);
          if ((rc = make_conv_tc(g, wm, p + ".convs2." + std::to_string(m), ch, ch, c.rk[j], 1, &g->rb_tc[i][j][m][1])))
            return fail(rc);
        } else {
          if ((rc = make_conv(g, wm, p + ".convs." + std::to_string(m), ch, ch, c.rk[j], c.dil[j][m],
                              &g->rb[i][j][m][0])))
            return fail(rc);
          if ((rc = make_conv_tc(g, wm, p + ".convs." + std::to_string(m), ch, ch, c.rk[j], c.dil[j][m],
                                 &g->rb_tc[i][j][m][0])))
            return fail(rc);
        }
      }
    // a stage runs on the tensor cores iff every one of its convs has a tcgen05 plan
    g->stage_tc[i] = true;
    for (int j = 0; j < c.n_rk; ++j)
      for (int m = 0; m < c.n_dil; ++m) {
        g->stage_tc[i] = g->stage_tc[i] && g->rb_tc[i][j][m][0].ok && (c.resblock != 1 || g->rb_tc[i][j][m][1].ok);
      }
  }
  // conv_pre and the upsamplers on the tensor cores: only when every stage runs there (planes flow end to end)
  g->tc_all = c.n_up > 0;
  for (int i = 0; i < c.n_up; ++i) g->tc_all = g->tc_all && g->stage_tc[i];
  if (g->tc_all) {
    if ((rc = make_conv_tc(g, wm, "conv_pre", c.model_in_dim, c.c0, 7, 1, &g->pre_tc))) return fail(rc);
    g->tc_all = g->pre_tc.ok;
    for (int i = 0; i < c.n_up && g->tc_all; ++i) {
      if ((rc = make_convt_tc(g, wm, "ups." + std::to_string(i), c.c0 >> i, c.c0 >> (i + 1), c.up_kernels[i],
                              c.up_rates[i], &g->ups_tc[i])))
        return fail(rc);
      g->tc_all = g->ups_tc[i].ok;
    }
  }
  // fused ResBlock pairs for the narrow stages (only inside the all-tensor-core pipeline, ResBlock1 only)
  for (int i = 0; i < c.n_up; ++i) {
    g->stage_pair[i] = g->tc_all && c.resblock == 1;
    for (int j = 0; j < c.n_rk && g->stage_pair[i]; ++j)
      for (int m = 0; m < c.n_dil && g->stage_pair[i]; ++m)
        g->stage_pair[i] = pair_plan(c.c0 >> (i + 1), c.rk[j], c.dil[j][m], g->rb_tc[i][j][m][0], g->rb_tc[i][j][m][1],
                                     &g->rb_pair[i][j][m]);
    // C = 16: the same pairs with two samples per GEMM row (wider MMAs); per pair, falls back to the kernel above
    if (g->stage_pair[i] && (c.c0 >> (i + 1)) == 16) {
      for (int j = 0; j < c.n_rk; ++j)
        for (int m = 0; m < c.n_dil; ++m) {
          Pack2Layer& P = g->rb_pack2[i][j][m];
          if (!pack2_plan(16, c.rk[j], c.dil[j][m], &P)) continue;
          const std::string pfx = "resblocks." + std::to_string(i * c.n_rk + j);
          auto k1 = pack2_weights(wm.get(pfx + ".convs1." + std::to_string(m) + ".weight")->data, P.k, P.S, &P.inv1);
          auto k2 = pack2_weights(wm.get(pfx + ".convs2." + std::to_string(m) + ".weight")->data, P.k, P.S, &P.inv2);
          if ((rc = tc_upload(g, k1, &P.w1)) || (rc = tc_upload(g, k2, &P.w2))) return fail(rc);
        }
    }
    // C = 64: planes-in / planes-out fused pairs with streamed weights (the stage after must not be the last one: its
    // output goes on as planes)
    g->stage_pair64[i] = g->tc_all && c.resblock == 1 && !g->stage_pair[i] && (c.c0 >> (i + 1)) == 64 && i + 1 < c.n_up;
    for (int j = 0; j < c.n_rk && g->stage_pair64[i]; ++j)
      for (int m = 0; m < c.n_dil && g->stage_pair64[i]; ++m)
        g->stage_pair64[i] = pair64_plan(64, c.rk[j], c.dil[j][m], g->rb_tc[i][j][m][0], g->rb_tc[i][j][m][1],
                                         &g->rb_pair64[i][j][m]);
  }
  // conv_post: (1, ch, 7) -> plain (ch, 7)
  {
    const dissc_tensor* w = wm.get("conv_post.weight");
    const dissc_tensor* b = wm.get("conv_post.bias");
    if (!w || !b) return fail(set_err(DISSC_EMISSING, "missing tensor conv_post.{weight,bias}"));
    if (w->numel != (int64_t)ch * 7 || b->numel != 1) return fail(set_err(DISSC_EINVAL, "conv_post: bad tensor sizes"));
    g->post.Cin = ch; g->post.Cout = 1; g->post.k = 7; g->post.pad = 3;
    if ((rc = dev_upload(g, w->data, (size_t)ch * 7, &g->post_w_plain))) return fail(rc);
    if ((rc = dev_upload(g, b->data, 1, &g->post.bias))) return fail(rc);
  }
  {
    const dissc_tensor* d = wm.get("dict.weight");
    if (!d || d->numel != (int64_t)c.num_embeddings * c.embedding_dim)
      return fail(set_err(DISSC_EMISSING, "dict.weight missing or not (%d,%d)", c.num_embeddings, c.embedding_dim));
    if ((rc = dev_upload(g, d->data, d->numel, &g->dict_w))) return fail(rc);
    if (c.has_spkr) {
      const dissc_tensor* s = wm.get("spkr.weight");
      if (!s || s->numel != (int64_t)c.n_spkr_rows * c.embedding_dim)
        return fail(set_err(DISSC_EMISSING, "spkr.weight missing or not (%d,%d)", c.n_spkr_rows, c.embedding_dim));
      if ((rc = dev_upload(g, s->data, s->numel, &g->spkr_w))) return fail(rc);
    }
  }
  g->n_launches = 1 + c.n_up * (1 + c.n_rk * c.n_dil * (c.resblock == 1 ? 2 : 1)) + 1;
  if (const char* e = getenv("DISSC_TC")) g->use_tc = atoi(e) != 0;
  estimate_act_scales(g, wm);
  *out = g;
  return DISSC_OK;
}

void dissc_gen_destroy(dissc_gen_t* g) {
  if (!g) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(g->device);
  for (void* p : g->allocs) cudaFree(p);
  if (g->arena) cudaFree(g->arena);
  if (g->hstream) cudaStreamDestroy(g->hstream);
  if (g->cstream) cudaStreamDestroy(g->cstream);
  for (int i = 0; i < kHostSlots; ++i) {
    if (g->fwd_done[i]) cudaEventDestroy(g->fwd_done[i]);
    if (g->d2h_done[i]) cudaEventDestroy(g->d2h_done[i]);
  }
  err_flag_destroy(&g->err);
  delete g;
  if (prev >= 0) cudaSetDevice(prev);
}

int dissc_gen_hop(const dissc_gen_t* g) { return g ? g->hop : 0; }

int dissc_gen_status(dissc_gen_t* g) {
  DISSC_CHECK(g, DISSC_EINVAL, "null handle");
  return err_flag_take(&g->err, "CodeGenerator forward", g->cfg.num_embeddings, g->cfg.n_spkr_rows);
}

int dissc_gen_set_tensor_cores(dissc_gen_t* g, int enable) {
  DISSC_CHECK(g, DISSC_EINVAL, "null handle");
  g->use_tc = enable != 0;
  return DISSC_OK;
}

int dissc_tc_set_tuning(int key, int value) {
  switch (key) {
    case 0: dissc::g_tc_na_pref = value; return DISSC_OK;
    case 1: dissc::g_tc_split_w = value ? 1 : 0; return DISSC_OK;
    case 2: dissc::g_tc_split256 = value ? 1 : 0; return DISSC_OK;
    case 3: dissc::g_tc_cluster2 = value ? 1 : 0; return DISSC_OK;
    case 4: dissc::g_hub_attn_tc = value ? 1 : 0; return DISSC_OK;
  }
  return dissc::set_err(DISSC_EINVAL, "unknown tuning key %d", key);
}

int dissc_tc_set_single_accumulator(int enable) {
  const int prev = g_single_acc256;
  g_single_acc256 = enable ? 1 : 0;
  return prev;
}

int dissc_gen_tensor_core_stages(const dissc_gen_t* g) {
  if (!g || !g->use_tc) return 0;
  int n = 0;
  for (int i = 0; i < g->cfg.n_up; ++i) n += g->stage_tc[i] ? 1 : 0;
  return n;
}
int dissc_gen_launches_per_forward(const dissc_gen_t* g) { return g ? g->n_launches : 0; }

int dissc_gen_workspace_bytes(const dissc_gen_t* g, int B, int T, size_t* bytes) {
  DISSC_CHECK(g && bytes && B > 0 && T > 0, DISSC_EINVAL, "bad argument");
  *bytes = kNumRegions * region_bytes(g, B, T);
  return DISSC_OK;
}

int dissc_gen_forward(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr,
                      const int32_t* lengths, int B, int T, float* out, void* workspace, size_t workspace_bytes,
                      void* stream) {
  return forward_impl(g, code, f0, spkr, lengths, B, T, out, nullptr, workspace, workspace_bytes,
                      static_cast<cudaStream_t>(stream), nullptr);
}

int dissc_gen_forward_i16(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr,
                          const int32_t* lengths, int B, int T, int16_t* out_i16, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return forward_impl(g, code, f0, spkr, lengths, B, T, nullptr, out_i16, workspace, workspace_bytes,
                      static_cast<cudaStream_t>(stream), nullptr);
}

// ---- host entry points -----------------------------------------------------------------------------------------
// Two input / output slots in device memory, one compute stream, one copy-back stream:
//   compute stream : H2D(slot s) -> forward(slot s)                 (one shared workspace: forwards are stream-ordered)
//   copy stream    : wait(forward s done) -> D2H(slot s) -> d2h_done[s]
// so with two batches in flight the 12-25 MB D2H of batch i rides under the forward of batch i+1.
static int host_arena_reserve(dissc_gen* g, int B, int T) {
  size_t ws = 0;
  dissc_gen_workspace_bytes(g, B, T, &ws);
  const size_t n_out = (size_t)B * out_len(g, T);
  const size_t b_io = align_up((size_t)B * T * 8, 256) + align_up((size_t)B * T * 4, 256) + align_up((size_t)B * 8, 256) +
                      align_up((size_t)B * 4, 256) + align_up(n_out * 4, 256);
  const size_t total = kHostSlots * b_io + ws;
  if (total <= g->arena_bytes) return DISSC_OK;
  // growing the arena: nothing may still be using the old one
  if (g->hstream) DISSC_CUDA(cudaStreamSynchronize(g->hstream));
  if (g->cstream) DISSC_CUDA(cudaStreamSynchronize(g->cstream));
  if (g->arena) DISSC_CUDA(cudaFree(g->arena));
  g->arena = nullptr;
  g->arena_bytes = 0;
  DISSC_CUDA(cudaMalloc(&g->arena, total));
  g->arena_bytes = total;
  return DISSC_OK;
}

static int host_streams(dissc_gen* g) {
  if (!g->hstream) DISSC_CUDA(cudaStreamCreateWithFlags(&g->hstream, cudaStreamNonBlocking));
  if (!g->cstream) DISSC_CUDA(cudaStreamCreateWithFlags(&g->cstream, cudaStreamNonBlocking));
  for (int i = 0; i < kHostSlots; ++i) {
    if (!g->fwd_done[i]) DISSC_CUDA(cudaEventCreateWithFlags(&g->fwd_done[i], cudaEventDisableTiming));
    if (!g->d2h_done[i]) DISSC_CUDA(cudaEventCreateWithFlags(&g->d2h_done[i], cudaEventDisableTiming));
  }
  return DISSC_OK;
}

int dissc_gen_host_reserve(dissc_gen_t* g, int B, int T) {
  DISSC_CHECK(g && B > 0 && T > 0, DISSC_EINVAL, "bad argument");
  struct DeviceGuard {
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  } guard;
  cudaGetDevice(&guard.prev);
  DISSC_CUDA(cudaSetDevice(g->device));
  DISSC_TRY(host_streams(g));
  return host_arena_reserve(g, B, T);
}

int dissc_gen_forward_ex(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr, const float* extra,
                         const int32_t* lengths, int B, int T, float* out_f32, int16_t* out_i16, void* workspace,
                         size_t workspace_bytes, void* stream) {
  DISSC_CHECK((out_f32 != nullptr) != (out_i16 != nullptr), DISSC_EINVAL, "exactly one of out_f32 / out_i16");
  return forward_impl(g, code, f0, spkr, lengths, B, T, out_f32, out_i16, workspace, workspace_bytes,
                      static_cast<cudaStream_t>(stream), nullptr, extra);
}

int dissc_gen_n_extra(const dissc_gen_t* g) { return g ? g->n_extra : 0; }

int dissc_gen_forward_host_submit(dissc_gen_t* g, int slot, const int64_t* code, const float* f0, const int64_t* spkr,
                                  const int32_t* lengths, int B, int T, float* out_f32, int16_t* out_i16) {
  DISSC_CHECK(g && code && ((out_f32 != nullptr) != (out_i16 != nullptr)), DISSC_EINVAL,
              "need a handle, code and exactly one output buffer");
  DISSC_CHECK(slot >= 0 && slot < kHostSlots, DISSC_EINVAL, "slot %d outside [0, %d)", slot, kHostSlots);
  DISSC_CHECK(B > 0 && T > 0, DISSC_EINVAL, "B=%d T=%d must be positive", B, T);
  struct DeviceGuard {   // the caller's current device is restored on every exit path
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  } guard;
  cudaGetDevice(&guard.prev);
  DISSC_CUDA(cudaSetDevice(g->device));
  DISSC_TRY(host_streams(g));
  DISSC_TRY(host_arena_reserve(g, B, T));
  size_t ws = 0;
  dissc_gen_workspace_bytes(g, B, T, &ws);
  const size_t n_out = (size_t)B * out_len(g, T);
  const size_t b_code = align_up((size_t)B * T * 8, 256), b_f0 = align_up((size_t)B * T * 4, 256),
               b_spk = align_up((size_t)B * 8, 256), b_len = align_up((size_t)B * 4, 256), b_out = align_up(n_out * 4, 256);
  const size_t b_io = b_code + b_f0 + b_spk + b_len + b_out;
  char* a = static_cast<char*>(g->arena) + (size_t)slot * b_io;
  int64_t* d_code = reinterpret_cast<int64_t*>(a);
  float* d_f0 = reinterpret_cast<float*>(a + b_code);
  int64_t* d_spk = reinterpret_cast<int64_t*>(a + b_code + b_f0);
  int32_t* d_len = reinterpret_cast<int32_t*>(a + b_code + b_f0 + b_spk);
  char* d_out = a + b_code + b_f0 + b_spk + b_len;
  void* d_ws = static_cast<char*>(g->arena) + (size_t)kHostSlots * b_io;
  cudaStream_t st = g->hstream;
  // the slot's previous copy-back must have left its output buffer before this forward overwrites it
  DISSC_CUDA(cudaStreamWaitEvent(st, g->d2h_done[slot], 0));
  DISSC_CUDA(cudaMemcpyAsync(d_code, code, (size_t)B * T * 8, cudaMemcpyHostToDevice, st));
  if (f0) DISSC_CUDA(cudaMemcpyAsync(d_f0, f0, (size_t)B * T * 4, cudaMemcpyHostToDevice, st));
  if (spkr) DISSC_CUDA(cudaMemcpyAsync(d_spk, spkr, (size_t)B * 8, cudaMemcpyHostToDevice, st));
  if (lengths) DISSC_CUDA(cudaMemcpyAsync(d_len, lengths, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  int rc = forward_impl(g, d_code, f0 ? d_f0 : nullptr, spkr ? d_spk : nullptr, lengths ? d_len : nullptr, B, T,
                        out_f32 ? reinterpret_cast<float*>(d_out) : nullptr,
                        out_i16 ? reinterpret_cast<int16_t*>(d_out) : nullptr, d_ws, ws, st, nullptr);
  if (rc) return rc;
  DISSC_CUDA(cudaEventRecord(g->fwd_done[slot], st));
  DISSC_CUDA(cudaStreamWaitEvent(g->cstream, g->fwd_done[slot], 0));
  if (out_f32)
    DISSC_CUDA(cudaMemcpyAsync(out_f32, d_out, n_out * 4, cudaMemcpyDeviceToHost, g->cstream));
  else
    DISSC_CUDA(cudaMemcpyAsync(out_i16, d_out, n_out * 2, cudaMemcpyDeviceToHost, g->cstream));
  DISSC_CUDA(cudaEventRecord(g->d2h_done[slot], g->cstream));
  return DISSC_OK;
}

int dissc_gen_forward_host_wait(dissc_gen_t* g, int slot) {
  DISSC_CHECK(g, DISSC_EINVAL, "null handle");
  DISSC_CHECK(slot >= 0 && slot < kHostSlots, DISSC_EINVAL, "slot %d outside [0, %d)", slot, kHostSlots);
  if (!g->d2h_done[slot]) return DISSC_OK;   // nothing was ever submitted
  DISSC_CUDA(cudaEventSynchronize(g->d2h_done[slot]));
  return err_flag_take(&g->err, "CodeGenerator forward", g->cfg.num_embeddings, g->cfg.n_spkr_rows);
}

int dissc_gen_forward_host(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr,
                           const int32_t* lengths, int B, int T, float* out_f32, int16_t* out_i16) {
  int rc = dissc_gen_forward_host_submit(g, 0, code, f0, spkr, lengths, B, T, out_f32, out_i16);
  if (rc) return rc;
  return dissc_gen_forward_host_wait(g, 0);
}

int dissc_gen_cost(const dissc_gen_t* g, int B, int T, double* flops, double* bytes) {
  DISSC_CHECK(g && B > 0 && T > 0, DISSC_EINVAL, "bad argument");
  const dissc_gen_cfg& c = g->cfg;
  double fl = 2.0 * g->pre.Cin * g->pre.Cout * g->pre.k * T;
  // layer-fused traffic model: every fused kernel reads each input once, writes its output once.
  double by = (8.0 + (c.has_f0 ? 4.0 : 0.0)) * T + (c.has_spkr ? 8.0 : 0.0) + 4.0 * g->pre.Cout * T;
  double wbytes = 4.0 * g->pre.Cin * g->pre.Cout * g->pre.k;
  double t = T;
  double s_prev = 4.0 * g->pre.Cout * T;
  for (int i = 0; i < c.n_up; ++i) {
    const ConvTLayer& U = g->ups[i];
    fl += 2.0 * U.Cin * U.Cout * U.k * t;
    wbytes += 4.0 * U.Cin * U.Cout * U.k;
    t = (t - 1) * U.u - 2 * U.pad + U.k;
    const double s = 4.0 * U.Cout * t;
    by += s_prev + s;
    const int per_pair = (c.resblock == 1) ? 5 : 3;  // K3: 1R+1W, K4: 2R+1W
    by += s * (c.n_rk * c.n_dil * per_pair + (c.n_rk - 1));  // + the xs reads of the MRF accumulate
    for (int j = 0; j < c.n_rk; ++j)
      for (int m = 0; m < c.n_dil; ++m) {
        const int nconv = (c.resblock == 1) ? 2 : 1;
        fl += nconv * 2.0 * U.Cout * U.Cout * c.rk[j] * t;
        wbytes += nconv * 4.0 * U.Cout * U.Cout * c.rk[j];
      }
    s_prev = s;
  }
  fl += 2.0 * g->post.Cin * g->post.k * t;
  by += s_prev + 4.0 * t;
  if (flops) *flops = fl * B;
  if (bytes) *bytes = by * B + wbytes;
  return DISSC_OK;
}

int dissc_gen_profile(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr,
                      const int32_t* lengths, int B, int T, float* out, void* workspace, size_t workspace_bytes,
                      char (*names)[64], float* ms, double* flops, double* bytes, int cap, int* n) {
  DISSC_CHECK(names && ms && flops && n, DISSC_EINVAL, "null profile arrays");
  Profiler prof{names, ms, flops, bytes, cap};
  int rc = forward_impl(g, code, f0, spkr, lengths, B, T, out, nullptr, workspace, workspace_bytes, nullptr, &prof);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc == DISSC_OK && e == cudaSuccess)
    for (int i = 0; i < prof.n; ++i) cudaEventElapsedTime(&ms[i], prof.ev[2 * i], prof.ev[2 * i + 1]);
  for (cudaEvent_t ev : prof.ev) cudaEventDestroy(ev);
  *n = prof.n;
  if (rc) return rc;
  DISSC_CUDA(e);
  return DISSC_OK;
}

// ---- layer-level test entry points -------------------------------------
int dissc_conv1d_fused(const float* in, const float* w_host, const float* bias_host, const float* res,
                       const float* acc_in, float* out, const int32_t* lengths, int len_mul, int B, int Cin, int Cout,
                       int T, int k, int dilation, int pre_act, float pre_slope, int post_act, float post_slope,
                       float div, void* stream) {
  DISSC_CHECK(in && w_host && out && B > 0 && Cin > 0 && Cout > 0 && T > 0, DISSC_EINVAL, "bad argument");
  DISSC_CHECK(conv_supported(k, dilation), DISSC_EUNSUPPORTED, "conv1d kernel_size=%d dilation=%d unsupported", k,
              dilation);
  const int cot = conv_co_tile(Cout);
  auto packed = pack_weights(w_host, Cin, Cout, k, cot, conv_ci_chunk(cot), false);
  float *dw = nullptr, *db = nullptr;
  DISSC_CUDA(cudaMalloc(&dw, packed.size() * 4));
  cudaMemcpy(dw, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice);
  if (bias_host) {
    cudaMalloc(&db, Cout * 4);
    cudaMemcpy(db, bias_host, Cout * 4, cudaMemcpyHostToDevice);
  }
  ConvParams p{};
  p.in = in; p.w = dw; p.bias = db; p.res = res; p.acc_in = acc_in; p.out = out;
  p.lengths = lengths; p.len_mul = len_mul;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.T = T; p.pad = (k * dilation - dilation) / 2;
  p.pre_act = pre_act; p.pre_slope = pre_slope; p.post_act = post_act; p.post_slope = post_slope; p.div = div;
  int rc = launch_conv(p, k, dilation, cot, false, static_cast<cudaStream_t>(stream));
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  cudaFree(dw);
  if (db) cudaFree(db);
  if (rc) return rc;
  DISSC_CUDA(e);
  return DISSC_OK;
}

int dissc_conv_transpose1d(const float* in, const float* w_host, const float* bias_host, float* out,
                           const int32_t* lengths, int len_mul, int B, int Cin, int Cout, int T_in, int k, int u,
                           void* stream) {
  DISSC_CHECK(in && w_host && out && B > 0 && Cin > 0 && Cout > 0 && T_in > 0, DISSC_EINVAL, "bad argument");
  DISSC_CHECK(convt_supported(k, u), DISSC_EUNSUPPORTED, "conv_transpose1d kernel_size=%d stride=%d unsupported", k, u);
  const int cot = convt_co_tile(Cout);
  auto packed = pack_weights(w_host, Cin, Cout, k, cot, 8, true);
  float *dw = nullptr, *db = nullptr;
  DISSC_CUDA(cudaMalloc(&dw, packed.size() * 4));
  cudaMemcpy(dw, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice);
  if (bias_host) {
    cudaMalloc(&db, Cout * 4);
    cudaMemcpy(db, bias_host, Cout * 4, cudaMemcpyHostToDevice);
  }
  ConvTParams p{};
  p.in = in; p.w = dw; p.bias = db; p.out = out; p.lengths = lengths; p.len_mul = len_mul;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.Tin = T_in; p.pad = (k - u) / 2;
  p.Tout = (T_in - 1) * u - 2 * p.pad + k;
  int rc = launch_convt(p, k, u, cot, static_cast<cudaStream_t>(stream));
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  cudaFree(dw);
  if (db) cudaFree(db);
  if (rc) return rc;
  DISSC_CUDA(e);
  return DISSC_OK;
}

int dissc_conv1d_tc(const float* in, const float* w_host, const float* bias_host, const float* res,
                    const float* acc_in, float* out_plain, float* out_raw, float* out_planes, const int32_t* lengths,
                    int len_mul, int B, int Cin, int Cout, int T, int k, int dilation, int pre_act, float pre_slope,
                    int post_act, float post_slope, float div, void* stream) {
  DISSC_CHECK(in && w_host && B > 0 && Cin > 0 && Cout > 0 && T > 0, DISSC_EINVAL, "bad argument");
  TcLayer L;
  DISSC_CHECK(tc_plan_conv(Cin, Cout, k, dilation, &L), DISSC_EUNSUPPORTED,
              "no tcgen05 plan for Cin=%d Cout=%d kernel_size=%d dilation=%d (Cout in {16,32,64,128,256} or a "
              "multiple of 256, padding <= %d)", Cin, Cout, k, dilation, kTcHalo);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Tr = (int)round_up(T, 128), Tp = Tr + 2 * kTcHalo;
  const int cin8 = L.Cin_pad / 8;
  const size_t in_elems = (size_t)B * cin8 * Tp * 8 + 4096, plane_elems = (size_t)B * (Cout / 8) * Tp * 8,
               f_elems = (size_t)B * (Cout / 8) * Tr * 8;
  auto packed = pack_weights_tc(L, [=](int n, int ci, int j) { return w_host[((size_t)n * Cin + ci) * k + j]; },
                                &L.inv_scale);
  std::vector<void*> tmp;
  auto dalloc = [&](size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    tmp.push_back(d);
    return d;
  };
  auto cleanup = [&]() { for (void* d : tmp) cudaFree(d); };
  __half* dw = (__half*)dalloc(packed.size() * 2);
  float* db = (float*)dalloc((size_t)Cout * 4);
  __half* a_hi = (__half*)dalloc(in_elems * 2);
  __half* a_lo = (__half*)dalloc(in_elems * 2);
  __half* o_hi = (__half*)dalloc(plane_elems * 2);
  __half* o_lo = (__half*)dalloc(plane_elems * 2);
  float* f_res = (float*)dalloc(f_elems * 4);
  float* f_acc = (float*)dalloc(f_elems * 4);
  float* f_out = (float*)dalloc(f_elems * 4);
  if (!dw || !db || !a_hi || !a_lo || !o_hi || !o_lo || !f_res || !f_acc || !f_out) {
    cleanup();
    return set_err(DISSC_ENOMEM, "cudaMalloc failed in dissc_conv1d_tc");
  }
  cudaMemcpyAsync(dw, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, st);
  std::vector<float> zero_bias(Cout, 0.f);
  cudaMemcpyAsync(db, bias_host ? bias_host : zero_bias.data(), (size_t)Cout * 4, cudaMemcpyHostToDevice, st);
  // garbage everywhere first: the kernel must not depend on anything but the zeroed halos
  cudaMemsetAsync(a_hi, 0x7b, in_elems * 2, st);
  cudaMemsetAsync(a_lo, 0x7b, in_elems * 2, st);
  cudaMemsetAsync(o_hi, 0x7b, plane_elems * 2, st);
  cudaMemsetAsync(o_lo, 0x7b, plane_elems * 2, st);
  const int nb = 148 * 4;
  const double ax = device_absmax(in, (size_t)B * Cin * T, st);
  const double aout = ax * host_row_norm(w_host, Cout, (size_t)Cin * k) + host_absmax(bias_host, Cout) +
                      device_absmax(res, res ? (size_t)B * Cout * T : 0, st) +
                      device_absmax(acc_in, acc_in ? (size_t)B * Cout * T : 0, st);
  const float s_in = pow2_scale_for_absmax(ax), s_out = pow2_scale_for_absmax(aout);
  tc_pack_planes_kernel<<<nb, 256, 0, st>>>(in, a_hi, a_lo, lengths, len_mul, B, Cin, cin8, T, Tp, pre_act, pre_slope, s_in);
  int rc = launch_zero_halos(a_hi, a_lo, B * cin8, Tp, T, st);
  if (!rc) rc = launch_zero_halos(o_hi, o_lo, B * Cout / 8, Tp, T, st);
  if (res) tc_plain_to_f32b_kernel<<<nb, 256, 0, st>>>(res, f_res, B, Cout, T, Tr);
  if (acc_in) tc_plain_to_f32b_kernel<<<nb, 256, 0, st>>>(acc_in, f_acc, B, Cout, T, Tr);
  L.w = dw;
  TcParams p{};
  p.a_hi = a_hi; p.a_lo = a_lo; p.bias = db;
  p.res = res ? f_res : nullptr; p.acc_in = acc_in ? f_acc : nullptr;
  p.out_f32b = out_raw ? f_out : nullptr;
  p.out_hi = out_planes ? o_hi : nullptr; p.out_lo = out_planes ? o_lo : nullptr;
  p.out_plain = out_plain;
  p.lengths = lengths; p.len_mul = len_mul;
  p.B = B; p.T = T; p.Tr = Tr; p.Tp = Tp; p.Tp_in = Tp;
  p.div = div; p.plane_act = post_act; p.plane_slope = post_slope; p.plain_act = post_act; p.plain_slope = post_slope;
  p.in_scale = s_in; p.plane_scale = s_out;
  if (!rc) rc = launch_conv_tc(p, L, T, st);
  if (!rc && out_raw) tc_f32b_to_plain_kernel<<<nb, 256, 0, st>>>(f_out, out_raw, B, Cout, T, Tr);
  if (!rc && out_planes) tc_planes_to_plain_kernel<<<nb, 256, 0, st>>>(o_hi, o_lo, out_planes, B, Cout, T, Tp, 1.f / s_out);
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (rc) return rc;
  DISSC_CUDA(e);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

int dissc_conv_transpose1d_tc(const float* in, const float* w_host, const float* bias_host, float* out_raw,
                              float* out_planes, const int32_t* lengths, int len_mul, int B, int Cin, int Cout,
                              int T_in, int k, int u, float plane_slope, void* stream) {
  DISSC_CHECK(in && w_host && B > 0 && Cin > 0 && Cout > 0 && T_in > 0 && u > 0, DISSC_EINVAL, "bad argument");
  DISSC_CHECK(k >= u && (k - u) % 2 == 0, DISSC_EUNSUPPORTED, "conv_transpose1d needs k >= stride and k - stride even");
  TcLayer L;
  DISSC_CHECK(tc_plan_convt(Cin, Cout, k, u, &L), DISSC_EUNSUPPORTED,
              "no tcgen05 plan for conv_transpose1d Cin=%d Cout=%d kernel_size=%d stride=%d", Cin, Cout, k, u);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int pad = (k - u) / 2;
  const int Tout = (T_in - 1) * u - 2 * pad + k;
  const int Tr_in = (int)round_up(T_in, 128), Tp_in = Tr_in + 2 * kTcHalo;
  const int Tr = (int)round_up(Tout, 128), Tp = Tr + 2 * kTcHalo;
  const int cin8 = L.Cin_pad / 8, M = L.k;
  const size_t in_elems = (size_t)B * cin8 * Tp_in * 8 + 4096, plane_elems = (size_t)B * (Cout / 8) * Tp * 8,
               f_elems = (size_t)B * (Cout / 8) * Tr * 8;
  auto packed = pack_weights_tc(L, [=](int n, int ci, int jp) {
    const int phase = n / Cout, co = n % Cout;
    const int jj = phase + (M - 1 - jp) * u;
    return jj < k ? w_host[((size_t)ci * Cout + co) * k + jj] : 0.f;
  }, &L.inv_scale);
  std::vector<void*> tmp;
  auto dalloc = [&](size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    tmp.push_back(d);
    return d;
  };
  auto cleanup = [&]() { for (void* d : tmp) cudaFree(d); };
  __half* dw = (__half*)dalloc(packed.size() * 2);
  float* db = (float*)dalloc((size_t)Cout * 4);
  __half* a_hi = (__half*)dalloc(in_elems * 2);
  __half* a_lo = (__half*)dalloc(in_elems * 2);
  __half* o_hi = (__half*)dalloc(plane_elems * 2);
  __half* o_lo = (__half*)dalloc(plane_elems * 2);
  float* f_out = (float*)dalloc(f_elems * 4);
  if (!dw || !db || !a_hi || !a_lo || !o_hi || !o_lo || !f_out) {
    cleanup();
    return set_err(DISSC_ENOMEM, "cudaMalloc failed in dissc_conv_transpose1d_tc");
  }
  cudaMemcpyAsync(dw, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, st);
  std::vector<float> zero_bias(Cout, 0.f);
  cudaMemcpyAsync(db, bias_host ? bias_host : zero_bias.data(), (size_t)Cout * 4, cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(a_hi, 0x7b, in_elems * 2, st);
  cudaMemsetAsync(a_lo, 0x7b, in_elems * 2, st);
  cudaMemsetAsync(o_hi, 0x7b, plane_elems * 2, st);
  cudaMemsetAsync(o_lo, 0x7b, plane_elems * 2, st);
  cudaMemsetAsync(f_out, 0x7b, f_elems * 4, st);
  const int nb = 148 * 4;
  const double ax = device_absmax(in, (size_t)B * Cin * T_in, st);
  // (Cin, Cout, k): bound the output by the input's absmax times the largest l2 norm over a whole input-channel row set
  const double aout = ax * std::sqrt((double)host_row_norm(w_host, 1, (size_t)Cin * Cout * k) *
                                     host_row_norm(w_host, 1, (size_t)Cin * Cout * k) / std::max(1, Cout * u)) +
                      host_absmax(bias_host, Cout);
  const float s_in = pow2_scale_for_absmax(ax), s_out = pow2_scale_for_absmax(aout);
  tc_pack_planes_kernel<<<nb, 256, 0, st>>>(in, a_hi, a_lo, lengths, len_mul, B, Cin, cin8, T_in, Tp_in, 0, 0.f, s_in);
  int rc = launch_zero_halos(a_hi, a_lo, B * cin8, Tp_in, T_in, st);
  if (!rc) rc = launch_zero_halos(o_hi, o_lo, B * Cout / 8, Tp, Tout, st);
  L.w = dw;
  TcParams p{};
  p.a_hi = a_hi; p.a_lo = a_lo; p.bias = db;
  p.out_f32b = f_out; p.out_hi = o_hi; p.out_lo = o_lo; p.plane_act = 1; p.plane_slope = plane_slope;
  p.lengths = lengths; p.len_mul = len_mul * u;
  p.B = B; p.T = Tout; p.Tr = Tr; p.Tp = Tp; p.Tp_in = Tp_in;
  p.in_scale = s_in; p.plane_scale = s_out;
  const int n_frames = (Tout + pad - 1) / u + 1;
  if (!rc) rc = launch_conv_tc(p, L, n_frames, st);
  if (!rc && out_raw) tc_f32b_to_plain_kernel<<<nb, 256, 0, st>>>(f_out, out_raw, B, Cout, Tout, Tr);
  if (!rc && out_planes) tc_planes_to_plain_kernel<<<nb, 256, 0, st>>>(o_hi, o_lo, out_planes, B, Cout, Tout, Tp, 1.f / s_out);
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (rc) return rc;
  DISSC_CUDA(e);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

int dissc_resblock_pair_tc(const float* in, const float* w1_host, const float* b1_host, const float* w2_host,
                           const float* b2_host, const float* acc_in, float* out_raw, float* out_planes,
                           const int32_t* lengths, int len_mul, int B, int C, int T, int k, int dilation, float div,
                           float plane_slope, void* stream) {
  DISSC_CHECK(in && w1_host && w2_host && B > 0 && C > 0 && T > 0, DISSC_EINVAL, "bad argument");
  if (C == 64)
    return pair64_layer_test(in, w1_host, b1_host, w2_host, b2_host, acc_in, out_raw, out_planes, lengths, len_mul, B, T, k,
                             dilation, div, plane_slope, static_cast<cudaStream_t>(stream));
  TcLayer c1, c2;
  PairLayer L;
  DISSC_CHECK(tc_plan_conv(C, C, k, dilation, &c1) && tc_plan_conv(C, C, k, 1, &c2) && pair_plan(C, k, dilation, c1, c2, &L),
              DISSC_EUNSUPPORTED, "no fused-pair plan for C=%d kernel_size=%d dilation=%d (C in {16,32}, odd k, halo <= %d)",
              C, k, dilation, kPairHalo);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Tr = (int)round_up(T, 128), Tp = Tr + 2 * kTcHalo, Tpf = Tr + kPairSlack;
  const size_t f_elems = (size_t)B * (C / 8) * Tpf * 8, plane_elems = (size_t)B * (C / 8) * Tp * 8;
  auto pk1 = pack_weights_tc(c1, [=](int n, int ci, int j) { return w1_host[((size_t)n * C + ci) * k + j]; }, &c1.inv_scale);
  auto pk2 = pack_weights_tc(c2, [=](int n, int ci, int j) { return w2_host[((size_t)n * C + ci) * k + j]; }, &c2.inv_scale);
  std::vector<void*> tmp;
  auto dalloc = [&](size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    tmp.push_back(d);
    return d;
  };
  auto cleanup = [&]() { for (void* d : tmp) cudaFree(d); };
  __half* dw1 = (__half*)dalloc(pk1.size() * 2);
  __half* dw2 = (__half*)dalloc(pk2.size() * 2);
  float* db = (float*)dalloc((size_t)2 * C * 4);
  float* f_in = (float*)dalloc(f_elems * 4);
  float* f_acc = (float*)dalloc(f_elems * 4);
  float* f_out = (float*)dalloc(f_elems * 4);
  __half* o_hi = (__half*)dalloc(plane_elems * 2);
  __half* o_lo = (__half*)dalloc(plane_elems * 2);
  if (!dw1 || !dw2 || !db || !f_in || !f_acc || !f_out || !o_hi || !o_lo) {
    cleanup();
    return set_err(DISSC_ENOMEM, "cudaMalloc failed in dissc_resblock_pair_tc");
  }
  cudaMemcpyAsync(dw1, pk1.data(), pk1.size() * 2, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dw2, pk2.data(), pk2.size() * 2, cudaMemcpyHostToDevice, st);
  std::vector<float> bias(2 * C, 0.f);
  if (b1_host) memcpy(bias.data(), b1_host, C * 4);
  if (b2_host) memcpy(bias.data() + C, b2_host, C * 4);
  cudaMemcpyAsync(db, bias.data(), (size_t)2 * C * 4, cudaMemcpyHostToDevice, st);
  // NaN bit patterns everywhere first: the kernel must not depend on the slack rows or on rows past the valid length
  cudaMemsetAsync(f_in, 0xff, f_elems * 4, st);
  cudaMemsetAsync(f_acc, 0xff, f_elems * 4, st);
  cudaMemsetAsync(f_out, 0xff, f_elems * 4, st);
  cudaMemsetAsync(o_hi, 0x7b, plane_elems * 2, st);
  cudaMemsetAsync(o_lo, 0x7b, plane_elems * 2, st);
  const int nb = 148 * 4;
  const double ax = device_absmax(in, (size_t)B * C * T, st), aacc = device_absmax(acc_in, acc_in ? (size_t)B * C * T : 0, st);
  const double axt = ax * host_row_norm(w1_host, C, (size_t)C * k) + host_absmax(b1_host, C);
  const double aout = ax + axt * host_row_norm(w2_host, C, (size_t)C * k) + host_absmax(b2_host, C) + aacc;
  const float s_in = pow2_scale_for_absmax(ax), s_xt = pow2_scale_for_absmax(axt), s_out = pow2_scale_for_absmax(aout);
  tc_plain_to_f32b_kernel<<<nb, 256, 0, st>>>(in, f_in + (size_t)kPairHalo * 8, B, C, T, Tpf);
  if (acc_in) tc_plain_to_f32b_kernel<<<nb, 256, 0, st>>>(acc_in, f_acc + (size_t)kPairHalo * 8, B, C, T, Tpf);
  int rc = launch_zero_halos(o_hi, o_lo, B * C / 8, Tp, T, st);
  c1.w = dw1; c2.w = dw2;
  PairParams p{};
  p.x = f_in; p.b1 = db; p.b2 = db + C; p.acc_in = acc_in ? f_acc : nullptr;
  p.out_f = f_out; p.out_hi = out_planes ? o_hi : nullptr; p.out_lo = out_planes ? o_lo : nullptr;
  p.lengths = lengths; p.len_mul = len_mul;
  p.B = B; p.T = T; p.Tpf = Tpf; p.f_halo = kPairHalo; p.Tp = Tp; p.p_halo = kTcHalo;
  p.div = div; p.plane_act = 1; p.plane_slope = plane_slope;
  p.in_scale = s_in; p.xt_scale = s_xt; p.plane_scale = s_out;
  Pack2Layer P2;
  if (!rc && pack2_plan(C, k, dilation, &P2)) {   // C = 16: the two-samples-per-row kernel is the one the model runs
    auto q1 = pack2_weights(w1_host, k, P2.S, &P2.inv1), q2 = pack2_weights(w2_host, k, P2.S, &P2.inv2);
    P2.w1 = (__half*)dalloc(q1.size() * 2);
    P2.w2 = (__half*)dalloc(q2.size() * 2);
    if (!P2.w1 || !P2.w2) {
      cleanup();
      return set_err(DISSC_ENOMEM, "cudaMalloc failed in dissc_resblock_pair_tc");
    }
    cudaMemcpyAsync(P2.w1, q1.data(), q1.size() * 2, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(P2.w2, q2.data(), q2.size() * 2, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);   // q1 / q2 are about to go out of scope
    Pack2Params r{};
    r.x = p.x; r.b1 = p.b1; r.b2 = p.b2; r.acc_in = p.acc_in; r.out_f = p.out_f; r.out_hi = p.out_hi; r.out_lo = p.out_lo;
    r.lengths = p.lengths; r.len_mul = p.len_mul; r.B = p.B; r.T = p.T; r.Tpf = p.Tpf; r.f_halo = p.f_halo; r.Tp = p.Tp;
    r.p_halo = p.p_halo; r.in_scale = s_in; r.xt_scale = s_xt; r.plane_scale = s_out; r.div = div; r.plane_act = 1;
    r.plane_slope = plane_slope;
    rc = launch_pack2(r, P2, st);
  } else if (!rc) {
    rc = launch_pair(p, L, c1, c2, st);
  }
  if (!rc && out_raw) tc_f32b_to_plain_kernel<<<nb, 256, 0, st>>>(f_out + (size_t)kPairHalo * 8, out_raw, B, C, T, Tpf);
  if (!rc && out_planes) tc_planes_to_plain_kernel<<<nb, 256, 0, st>>>(o_hi, o_lo, out_planes, B, C, T, Tp, 1.f / s_out);
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (rc) return rc;
  DISSC_CUDA(e);
  DISSC_CUDA(cudaGetLastError());
  return DISSC_OK;
}

}  // extern "C"

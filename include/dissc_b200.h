/*
 * dissc_b200 -- C ABI of the B200-native DISSC inference hot path.
 *
 * The reference (gallilmaimon/DISSC) has no FFI of its own: its seam is the
 * Python nn.Module surface used by its CLIs.  Each entry point below states
 * the reference interface it replaces (file:line under the reference tree);
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 (DISSC_OK) or a negative DISSC_E* code; the
 *     message for the calling thread's last failure is dissc_last_error().
 *     Nothing throws, nothing calls exit().
 *   - a handle is bound to one CUDA device and owns only its re-packed
 *     weights (plus, for the *_host entry points, a cached staging arena).
 *     Calls are asynchronous and ordered on the stream passed in (a
 *     cudaStream_t passed as void*; NULL = the legacy default stream); no
 *     internal host synchronisation unless the name ends in _host.
 *   - a handle is not re-entrant across threads; distinct handles are
 *     independent.  One process per GPU.
 *   - activation layout is the reference's: (B, C, T) fp32, T contiguous.
 */
#ifndef DISSC_B200_H
#define DISSC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DISSC_OK 0
#define DISSC_EINVAL (-1)       /* bad argument / shape */
#define DISSC_EUNSUPPORTED (-2) /* geometry the kernels do not implement */
#define DISSC_ECUDA (-3)        /* CUDA runtime error (message has the cudaError string) */
#define DISSC_ENOMEM (-4)
#define DISSC_EMISSING (-5)     /* a required tensor is absent from the weight list */
#define DISSC_EINDEX (-6)       /* a unit / speaker id outside its embedding table (nn.Embedding raises IndexError) */

#define DISSC_MAX_STAGES 8
#define DISSC_MAX_KERNELS 8
#define DISSC_MAX_DILATIONS 8

typedef struct {
  const char* name; /* reference state-dict key, e.g. "resblocks.3.convs1.0.weight" */
  const float* data; /* HOST pointer, fp32, contiguous, PyTorch layout */
  int64_t numel;
} dissc_tensor;

/* ------------------------------------------------------------------ *
 * Vocoder: CodeGenerator  (sr/models.py:72-225)
 * ------------------------------------------------------------------ */

/* Geometry = the keys of <ckpt_dir>/config.json that sr/models.py reads
 * (Generator.__init__ :73-96, CodeGenerator.__init__ :126-156). */
typedef struct {
  int n_up;                                                 /* len(upsample_rates) */
  int up_rates[DISSC_MAX_STAGES];                           /* upsample_rates */
  int up_kernels[DISSC_MAX_STAGES];                         /* upsample_kernel_sizes */
  int n_rk;                                                 /* len(resblock_kernel_sizes) */
  int rk[DISSC_MAX_KERNELS];                                /* resblock_kernel_sizes */
  int n_dil;                                                /* len(resblock_dilation_sizes[j]) */
  int dil[DISSC_MAX_KERNELS][DISSC_MAX_DILATIONS];          /* resblock_dilation_sizes */
  int c0;                                                   /* upsample_initial_channel */
  int embedding_dim;                                        /* embedding_dim */
  int num_embeddings;                                       /* num_embeddings (rows of dict.weight) */
  int n_spkr_rows;                                          /* rows of spkr.weight (200, sr/models.py:133) */
  int model_in_dim;                                         /* model_in_dim */
  int resblock;                                             /* 1 = ResBlock1, 2 = ResBlock2 */
  int has_f0;                                               /* h.f0 */
  int has_spkr;                                             /* h.multispkr */
} dissc_gen_cfg;

typedef struct dissc_gen dissc_gen_t;

/* Replaces: CodeGenerator(h).to(dev); load_state_dict(...); eval(); remove_weight_norm()
 * (sr/inference.py:114-120,162-163).  `weights` are the FOLDED tensors
 * (weight = g*v/||v||, sr/models.py:116-122) keyed by the reference names with
 * suffix ".weight"/".bias", plus "dict.weight" and "spkr.weight".  They are
 * copied and re-packed on the device; the caller may free them on return. */
int dissc_gen_create(dissc_gen_t** out, const dissc_gen_cfg* cfg, const dissc_tensor* weights, int n_weights,
                     int device);
void dissc_gen_destroy(dissc_gen_t* g);

/* Total upsampling factor (prod(upsample_rates), 320 for the shipped configs). */
int dissc_gen_hop(const dissc_gen_t* g);

/* Bytes of device scratch dissc_gen_forward needs for a (B,T) batch. */
int dissc_gen_workspace_bytes(const dissc_gen_t* g, int B, int T, size_t* bytes);

/* Replaces: generator(code=..., f0=..., spkr=...)  (sr/inference.py:69 ->
 * CodeGenerator.forward sr/models.py:179-225 -> Generator.forward :98-114).
 * All pointers are DEVICE pointers on the handle's device.
 *   code    int64 (B,T)      unit ids in [0,num_embeddings)
 *   f0      fp32  (B,1,T)    may be NULL iff !has_f0
 *   spkr    int64 (B,1)      may be NULL iff !has_spkr
 *   lengths int32 (B)        valid frames per utterance, NULL = all T.  Frames
 *                            >= lengths[b] behave exactly like the zero padding
 *                            a B=1 reference call sees at the utterance's true
 *                            end; out[b, t >= hop*lengths[b]] is written as 0.
 *   out     fp32  (B,1,hop*T)
 * Asynchronous on `stream`. */
int dissc_gen_forward(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr,
                      const int32_t* lengths, int B, int T, float* out, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Same, with the int16 conversion of generate() fused into the last kernel:
 * audio = (y*32768).astype(int16)  (sr/inference.py:73-75; C-style truncation,
 * +1.0 wraps to -32768 exactly like numpy).  out_i16 is (B, hop*T). */
int dissc_gen_forward_i16(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr,
                          const int32_t* lengths, int B, int T, int16_t* out_i16, void* workspace,
                          size_t workspace_bytes, void* stream);

/* The same forward for configs with EXTRA conditioning features (`f0_feats`: CodeGenerator.forward appends every
 * other keyword argument as channels repeated over time, sr/models.py:216-221; sr/inference.py:237-245 passes
 * `f0_stats` = the target speaker's [mean, std]).  `extra` fp32 (B, n_extra) on the device, n_extra = model_in_dim -
 * embedding_dim - has_f0 - has_spkr * embedding_dim = dissc_gen_n_extra(g); channel order = order of the columns.
 * Exactly one of out_f32 / out_i16 is non-NULL.  dissc_gen_forward / _i16 / _host reject handles with n_extra > 0. */
int dissc_gen_forward_ex(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr, const float* extra,
                         const int32_t* lengths, int B, int T, float* out_f32, int16_t* out_i16, void* workspace,
                         size_t workspace_bytes, void* stream);
int dissc_gen_n_extra(const dissc_gen_t* g);

/* End-to-end call with HOST buffers (what sr/inference.py:178 + :69 + :75 do per
 * utterance, batched): H2D of the inputs, forward, D2H of the waveform, stream
 * synchronise.  Pinned host memory makes the copies asynchronous; pageable
 * memory works but serialises.  Exactly one of out_f32 / out_i16 is non-NULL. */
int dissc_gen_forward_host(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr,
                           const int32_t* lengths, int B, int T, float* out_f32, int16_t* out_i16);

/* Pipelined form of dissc_gen_forward_host: DISSC_HOST_SLOTS batches may be in flight.  _submit enqueues H2D ->
 * forward -> D2H for `slot` and returns without waiting; _wait(slot) blocks until that slot's output is in the host
 * buffer given to _submit (and reports out-of-range ids like dissc_gen_forward_host).  Submitting slot s+1 before
 * waiting for slot s hides the copy-back of one batch under the forward of the next; host buffers of a slot must stay
 * valid (and pinned, for the copies to be asynchronous) until its _wait returns.  Re-submitting a slot without
 * waiting for it is allowed only if its output buffer may be overwritten.  dissc_gen_forward_host == submit(0) + wait(0).
 * dissc_gen_host_reserve pre-sizes the device staging arena for (B,T) batches so that no call allocates (the arena
 * otherwise grows on the first call that needs more). */
#define DISSC_HOST_SLOTS 2
int dissc_gen_forward_host_submit(dissc_gen_t* g, int slot, const int64_t* code, const float* f0, const int64_t* spkr,
                                  const int32_t* lengths, int B, int T, float* out_f32, int16_t* out_i16);
int dissc_gen_forward_host_wait(dissc_gen_t* g, int slot);
int dissc_gen_host_reserve(dissc_gen_t* g, int B, int T);

/* Unit and speaker ids are range-checked on the device (nn.Embedding raises IndexError / a device assert on an id
 * outside its table, sr/models.py:128,133,189,213): a bad id never reads outside the table (row 0 is used), and it sets
 * a per-handle flag.  dissc_gen_status returns DISSC_EINDEX (and clears the flag) if any forward enqueued through this
 * handle so far saw one -- it does not synchronise, so synchronise the stream first for a definitive answer.
 * dissc_gen_forward_host checks after its own synchronisation; dissc_gen_forward{,_i16} check on entry (i.e. report a
 * bad id of the PREVIOUS batch at the latest). */
int dissc_gen_status(dissc_gen_t* g);

/* The ResBlock stages whose geometry allows it (C in {16,32,64,128,256}, padding <= 32) run on the tcgen05
 * tensor cores with split-fp16 operands (fp32-accurate); when every stage does, conv_pre and the
 * transposed convs run there too.  The others, and everything when disabled, use the fp32 CUDA-core kernels.
 * Default: enabled (env DISSC_TC=0 disables). */
int dissc_gen_set_tensor_cores(dissc_gen_t* g, int enable);
int dissc_gen_tensor_core_stages(const dissc_gen_t* g); /* how many stages currently take the tensor-core path */

/* 256-column GEMMs (the C = 256 stage, conv_pre, the first upsampler).  Default: two 128-column chunks, each with
 * separate main and cross-term accumulators double-buffered in TMEM (waveform error 7e-6 max-abs vs fp64 on the benchmark
 * weights, 4e-5 on the `hot` recipe).  Opt-in fast mode: dissc_tc_set_tuning(2, 0) + dissc_tc_set_single_accumulator(1) =
 * one 256-column chunk with all three split-precision MMAs into ONE accumulator (stage 0 10 % faster; the tensor core
 * truncates the accumulator after every MMA, so the error is 2.5x larger: 1.8e-5 / 1.0e-4).  Both apply to handles /
 * layer calls created afterwards.  dissc_tc_set_single_accumulator returns the previous setting (-1 = never set: env
 * DISSC_TC_SINGLE_ACC or the default 0 decides). */
int dissc_tc_set_single_accumulator(int enable);
/* Plan-time tuning of the tensor-core conv kernel, read when a handle is created (A/B measurements, scripts/ab_tuning.py):
 * key 0 = number of activation buffers of the streamed-weight layers (2..4; 0 = heuristic), key 1 = separate
 * weight-producer thread in the N >= 128 kernels (0 / 1), key 2 = 256-column GEMMs as two 128-column chunks (1, default)
 * or one 256-column chunk (0), key 3 = 2-CTA clusters that share one multicast weight stream in the streamed-weight
 * kernels (0 = off, default: measured neutral; read at launch time), key 4 = HuBERT attention on the tensor cores for
 * clips of <= 320 frames (1, default) or always the fp32 CUDA-core kernel (0).  No reference counterpart. */
int dissc_tc_set_tuning(int key, int value);

/* Number of kernel launches one forward issues (for bench.py's gpu_launches). */
int dissc_gen_launches_per_forward(const dissc_gen_t* g);

/* Algorithmic FLOPs and layer-fused-model bytes of one (B,T) forward. */
int dissc_gen_cost(const dissc_gen_t* g, int B, int T, double* flops, double* bytes);

/* Per-layer device timing of one forward (cudaEvent around every launch; debug /
 * profiling aid).  names/ms/flops/bytes are caller arrays of capacity `cap`
 * (bytes may be NULL; it receives each launch's algorithmic bytes under the
 * layer-fused traffic model); *n receives the number of launches.  Synchronises. */
int dissc_gen_profile(dissc_gen_t* g, const int64_t* code, const float* f0, const int64_t* spkr,
                      const int32_t* lengths, int B, int T, float* out, void* workspace, size_t workspace_bytes,
                      char (*names)[64], float* ms, double* flops, double* bytes, int cap, int* n);

/* ------------------------------------------------------------------ *
 * Prosody predictors  (model/len_predictor.py, model/pitch_predictor.py, infer.py)
 * ------------------------------------------------------------------ */
#define DISSC_PRED_LEN 0        /* LenPredictor         model/len_predictor.py:5-52 */
#define DISSC_PRED_PITCH_NEW 1  /* PitchPredictor       model/pitch_predictor.py:41-104 */
#define DISSC_PRED_PITCH_BASE 2 /* PitchPredictorBase   model/pitch_predictor.py:106-176 */

typedef struct dissc_pred dissc_pred_t;

/* Replaces: LenPredictor(n_tokens, n_speakers) / PitchPredictor[Base](n_tokens, n_speakers, ...) + .to(dev) +
 * load_state_dict(torch.load(dir + 'best_model.pth')) + eval()  (infer.py:68-84).  `weights` are the float tensors of
 * the reference state dict under their reference names ("cnn1.weight", "bn1.running_var", "token_emb.weight",
 * "pe.pe", ...); eval-mode BatchNorm is folded into the preceding conv here. */
int dissc_pred_create(dissc_pred_t** out, int kind, int n_tokens, int n_speakers, const dissc_tensor* weights,
                      int n_weights, int device);
void dissc_pred_destroy(dissc_pred_t* g);
int dissc_pred_workspace_bytes(const dissc_pred_t* g, int B, int L, size_t* bytes);
/* Same contract as dissc_gen_status for token ids (rows n_tokens + 1) and speaker ids of the predictors
 * (nn.Embedding at model/len_predictor.py:15-16, model/pitch_predictor.py:51-52). */
int dissc_pred_status(dissc_pred_t* g);

/* Replaces: len_model(dd_seq, spk_id)  (infer.py:30 -> LenPredictor.forward model/len_predictor.py:35-52), batched:
 * seq int64 (B,L) deduplicated units, spk int64 (B), lengths int32 (B) valid tokens per row (NULL = L; rows behave like
 * B=1 calls on the unpadded sequence), norm_mean/norm_std = len_norm_stats.pth (infer.py:72); out fp32 (B,L). */
int dissc_len_forward(dissc_pred_t* g, const int64_t* seq, const int64_t* spk, const int32_t* lengths, int B, int L,
                      float norm_mean, float norm_std, float* out, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Replaces: PitchPredictor[Base].forward  (model/pitch_predictor.py:72-94 / :145-166): cls, reg fp32 (B,L). */
int dissc_pitch_forward(dissc_pred_t* g, const int64_t* seq, const int64_t* spk, const int32_t* lengths, int B, int L,
                        float* cls, float* reg, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces: calc_freq  (model/pitch_predictor.py:100-104): out = (cls > 0) * (mean[spk] + reg * std[spk]), or
 * (cls > 0) * reg when mean == std == NULL (norm=True).  mean/std are DEVICE arrays of n_stats_rows entries indexed by
 * speaker id; a speaker id outside [0, n_stats_rows) (the reference raises IndexError) yields NaN for that row. */
int dissc_pitch_calc_freq(const float* cls, const float* reg, const int64_t* spk, const float* mean, const float* std,
                          int n_stats_rows, const int32_t* lengths, int B, int L, float* out, void* stream);

/* Replaces: len_carryover_correction  (infer.py:158-172), batched.  lens fp32 (B,L) -> out int32 (B,L) (0 past the
 * valid length), totals int32 (B) = sum of the corrected lengths (may be NULL).  Bit-exact: fp32 running sum,
 * round-half-to-even of clamp(lens, 1). */
int dissc_len_carryover(const float* lens, const int32_t* lengths, int B, int L, int32_t* out, int32_t* totals,
                        void* stream);

/* Replaces: seqs[seqs != n_tokens] + dedup_seq  (infer.py:25-27, dataset/utils.py:14-16), batched: dd int64 (B,L)
 * (padded with pad_token), counts int32 (B,L) run lengths, dd_len int32 (B). */
int dissc_dedup_units(const int64_t* seq, const int32_t* lengths, int64_t pad_token, int B, int L, int64_t* dd,
                      int32_t* counts, int32_t* dd_len, void* stream);

/* Replaces: torch.repeat_interleave(dd_seq, lens)  (infer.py:32), batched: out int64 (B,L_out) padded with pad_token,
 * out_len int32 (B) (clipped to L_out). */
int dissc_repeat_interleave(const int64_t* dd, const int32_t* counts, const int32_t* dd_len, int64_t pad_token, int B,
                            int L, int L_out, int64_t* out, int32_t* out_len, void* stream);

/* ------------------------------------------------------------------ *
 * Unit encoder: HuBERT-base layer-6 features -> k-means units  (data/encode.py:21-22,32)
 * ------------------------------------------------------------------ */

/* Geometry of the fairseq HuBERT-base graph (hubert_base_ls960: 7-layer conv extractor (512; k 10,3,3,3,3,2,2;
 * s 5,2,2,2,2,2,2; mode "default"), embed 768, ffn 3072, 12 heads, pos_conv k128 g16) and of the quantiser. */
typedef struct {
  int n_layers;   /* transformer layers to run (output_layer = 6 for the shipped pipeline) */
  int embed_dim;  /* 768 */
  int ffn_dim;    /* 3072 */
  int n_heads;    /* 12 (head size must be 64) */
  int conv_dim;   /* 512 */
  int pos_kernel; /* 128 */
  int pos_groups; /* 16 */
  int n_clusters; /* k-means vocabulary (100) */
} dissc_hubert_cfg;

typedef struct dissc_hubert dissc_hubert_t;

/* Replaces: SpeechEncoder.by_name(dense_model_name='hubert-base-ls960', quantizer_model_name='kmeans', vocab_size=100,
 * deduplicate=False).to(device)  (data/encode.py:21-22).  `weights`: fp32 tensors under the fairseq HuBERT state-dict
 * names ("feature_extractor.conv_layers.0.0.weight", "feature_extractor.conv_layers.0.2.{weight,bias}" (GroupNorm),
 * "layer_norm.*", "post_extract_proj.*", "encoder.pos_conv.0.{weight,bias}" with the weight-norm ALREADY FOLDED
 * (w = g * v / ||v|| over dims 0,1 per tap), "encoder.layer_norm.*", "encoder.layers.{l}.self_attn.{q,k,v,out}_proj.*",
 * "encoder.layers.{l}.self_attn_layer_norm.*", "encoder.layers.{l}.fc{1,2}.*", "encoder.layers.{l}.final_layer_norm.*")
 * plus "kmeans.cluster_centers" (n_clusters, embed_dim) = sklearn cluster_centers_ of km.bin. */
int dissc_hubert_create(dissc_hubert_t** out, const dissc_hubert_cfg* cfg, const dissc_tensor* weights, int n_weights,
                        int device);
void dissc_hubert_destroy(dissc_hubert_t* g);

/* Frames produced for a clip of n_samples samples: the (n-k)/s+1 chain of the 7 convs (= floor((n-400)/320)+1). */
int dissc_hubert_num_frames(int n_samples);
int dissc_hubert_workspace_bytes(const dissc_hubert_t* g, int B, int N, size_t* bytes);

/* Replaces: encoder(waveform)  (data/encode.py:32), batched.  DEVICE pointers:
 *   wave      fp32 (B,N)   16 kHz samples, rows zero-padded to N
 *   n_samples int32 (B)    valid samples per clip (NULL = N); each row is processed exactly like a B=1 call on the
 *                          unpadded clip (keys past its last frame are masked in attention, convs see zero padding)
 *   units     int64 (B,T)  T = dissc_hubert_num_frames(N); -1 past a clip's last frame
 *   n_frames  int32 (B)    frames per clip (may be NULL)
 *   features  fp32 (B,T,embed_dim) layer-`n_layers` features ('dense'), may be NULL
 * Asynchronous on `stream`. */
int dissc_hubert_forward(dissc_hubert_t* g, const float* wave, const int32_t* n_samples, int B, int N, int64_t* units,
                         int32_t* n_frames, float* features, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces: KMeansQuantizer.forward  (textless): out[i] = argmin_j ||x[i] - centroids[j]||^2, lowest index on ties.
 * x fp32 (M,D) row-major, centroids fp32 (K,D), out int64 (M); DEVICE pointers. */
int dissc_kmeans_assign(const float* x, const float* centroids, int M, int D, int K, int64_t* out, void* stream);

/* ------------------------------------------------------------------ *
 * Generic fused layers (exposed for layer-level parity tests)
 * ------------------------------------------------------------------ */

/* out = post( [acc_in +] [res +] bias + conv1d(pre(in), w) ) [/ div]
 *   pre(x)  = leaky_relu(x, pre_slope)  if pre_act  else x
 *   post(x) = leaky_relu(x, post_slope) if post_act else x
 * w is (Cout,Cin,k) PyTorch layout on the HOST (packed internally per call --
 * test entry point, not a fast path).  in/res/acc_in/out are DEVICE pointers. */
int dissc_conv1d_fused(const float* in, const float* w_host, const float* bias_host, const float* res,
                       const float* acc_in, float* out, const int32_t* lengths, int len_mul, int B, int Cin, int Cout,
                       int T, int k, int dilation, int pre_act, float pre_slope, int post_act, float post_slope,
                       float div, void* stream);

/* out = bias + conv_transpose1d(in, w) with w (Cin,Cout,k) on the HOST, stride u, padding (k-u)/2. */
int dissc_conv_transpose1d(const float* in, const float* w_host, const float* bias_host, float* out,
                           const int32_t* lengths, int len_mul, int B, int Cin, int Cout, int T_in, int k, int u,
                           void* stream);

/* Tensor-core twin of dissc_conv1d_fused.  Plain (B,C,T) fp32 device tensors in and out; the entry point converts
 * to/from the blocked tensor-core layouts itself (test entry, not a fast path).  Any of out_plain (post-activated),
 * out_raw (value before post), out_planes (fp16 hi+lo of the post-activated value, summed back to fp32) may be NULL.
 * Cout must be 16/32/64/128/256 or a multiple of 256; Cin is zero-padded to a multiple of 16 internally. */
int dissc_conv1d_tc(const float* in, const float* w_host, const float* bias_host, const float* res,
                    const float* acc_in, float* out_plain, float* out_raw, float* out_planes, const int32_t* lengths,
                    int len_mul, int B, int Cin, int Cout, int T, int k, int dilation, int pre_act, float pre_slope,
                    int post_act, float post_slope, float div, void* stream);

/* Tensor-core twin of dissc_conv_transpose1d (polyphase implicit GEMM; ConvTranspose1d of sr/models.py:82-86,:102).
 * out_raw = bias + conv_transpose1d(in, w); out_planes = leaky_relu(out_raw, plane_slope) through the fp16 split. */
int dissc_conv_transpose1d_tc(const float* in, const float* w_host, const float* bias_host, float* out_raw,
                              float* out_planes, const int32_t* lengths, int len_mul, int B, int Cin, int Cout,
                              int T_in, int k, int u, float plane_slope, void* stream);

/* Fused ResBlock1 pair on the tensor cores (C = 16 or 32; sr/models.py:36-40):
 *   out_raw = [acc_in +] in + conv1d(lrelu(conv1d(lrelu(in, 0.1), w1, dilation), 0.1), w2)   [/ div]
 *   out_planes = leaky_relu(out_raw, plane_slope) through the fp16 split.
 * Plain (B,C,T) fp32 DEVICE tensors; weights (C,C,k) / biases (C) on the HOST (test entry, not a fast path). */
int dissc_resblock_pair_tc(const float* in, const float* w1_host, const float* b1_host, const float* w2_host,
                           const float* b2_host, const float* acc_in, float* out_raw, float* out_planes,
                           const int32_t* lengths, int len_mul, int B, int C, int T, int k, int dilation, float div,
                           float plane_slope, void* stream);

const char* dissc_last_error(void);
const char* dissc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DISSC_B200_H */

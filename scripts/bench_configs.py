"""BASELINE configs[2..4] on synthetic data (importable; bench.py folds the records into its JSON line as `configs`,
scripts/bench_pipeline.py prints them one per line at full size).

  config 3  infer.py prosody path: LenPredictor + PitchPredictor -> CodeGenerator     (infer.py:24-45,101-122)
  config 4  data/encode.py: HuBERT-base layer-6 features + k-means-100 units          (data/encode.py:27-41)
  config 5  encode -> len / pitch predict -> vocode chained on the same clips

One process per GPU, every rank on its own shard (seed 1234 + rank), no data-path collective; times are CUDA events
on the launching stream, max over ranks via `mx`.  Weights: seeded synthetic checkpoints of the shipped geometries
(no pretrained files offline).
"""
from __future__ import annotations

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HUBERT_GFLOP_PER_CLIP = 59.6   # 96 000-sample clip, 6 layers


def timed(fn, iters, dev, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters, out


def build_predictors(dev):
    from dissc_b200 import synthetic as syn
    from dissc_b200.predictors import LenPredictor, PitchPredictor
    lm = LenPredictor(100, 108).to(dev)
    lm.load_state_dict(syn.synthetic_len_predictor_state_dict(100, 108, seed=21))
    # rhythm statistics: ~2.5 frames per deduplicated unit (VCTK-like).  The spread is kept small because the random-init
    # network's raw output is O(1..10): with the SURVEY's 1.5 most predictions clamp to 1 and the carry-over diffusion
    # then deletes units, which would shrink the vocoder's share of the step to ~50 frames per utterance
    lm.norm_mean, lm.norm_std = torch.tensor(2.5), torch.tensor(0.05)
    mean, std = syn.synthetic_pitch_stats(108, seed=22)
    pm = PitchPredictor(100, 108, id2pitch_mean=mean.to(dev), id2pitch_std=std.to(dev)).to(dev)
    pm.load_state_dict(syn.synthetic_pitch_predictor_state_dict("new", 100, 108, seed=23))
    return lm, pm


def build_generator(dev):
    from dissc_b200 import AttrDict, CodeGenerator
    from dissc_b200 import synthetic as syn
    gen = CodeGenerator(AttrDict(syn.VCTK_CONFIG)).to(dev)
    gen.load_state_dict(syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0))
    gen.eval()
    gen.remove_weight_norm()
    return gen


def build_encoder(dev):
    """HuBERT-base with torchaudio's random initialisation (seed 0) under fairseq's names + 100 random centroids."""
    import torchaudio
    from dissc_b200.hubert import SpeechEncoder
    from oracle import hubert_oracle as ho   # weight-NAME mapping only (synthetic weights = torchaudio's init)
    torch.manual_seed(0)
    hsd = ho.from_torchaudio(torchaudio.models.hubert_base().eval(), 6)
    cent = torch.randn(100, 768, generator=torch.Generator().manual_seed(5))
    return SpeechEncoder.from_state_dict(hsd, cent).to(dev), hsd


def synthetic_unit_sequences(B, g):
    """T ~ U{200..400} raw frames, run lengths ~ geometric (mean 2.5); padded with token 100."""
    seqs = torch.full((B, 400), 100, dtype=torch.int64)
    for b in range(B):
        T = int(torch.randint(200, 401, (1,), generator=g))
        toks, t = [], 0
        while t < T:
            run = int(torch.distributions.Geometric(probs=torch.tensor(0.4)).sample()) + 1
            tok = int(torch.randint(0, 100, (1,), generator=g))
            toks += [tok] * min(run, T - t)
            t += run
        seqs[b, :T] = torch.tensor(toks[:T])
    return seqs


def vocode_sorted(gen, out_seq, f0, spk, out_len, vocode_batch=64):
    """length-sorted sub-batches of `vocode_batch` utterances -> (total samples vocoded, last int16 batch)"""
    order = torch.argsort(out_len, descending=True)
    n, y = 0, None
    lens_h = out_len.cpu()
    order_h = order.cpu()
    for i0 in range(0, len(order), vocode_batch):
        idx = order[i0:i0 + vocode_batch]
        L = int(lens_h[order_h[i0]])
        code = out_seq[idx, :L].clone()
        code[code >= 100] = 0
        y = gen.generate_int16(code, f0[idx, :L].contiguous(), spk[idx], lengths=out_len[idx])
        n += int(lens_h[order_h[i0:i0 + vocode_batch]].sum()) * gen.hop
    return n, y


def run_config1(gen, dev, cpu_leg=False, n=30):
    """BASELINE configs[0]: ONE utterance of 50 units (16 000 samples), Generator forward only -- the reference's own use
    case (batch size 1, sr/inference.py:178,247): latency of an eager forward and of a CUDA-graph replay."""
    from dissc_b200 import synthetic as syn
    code, f0, spkr = (t.to(dev) for t in syn.synthetic_inputs(1, 50, seed=1234))
    for _ in range(5):
        y = gen(code=code, f0=f0, spkr=spkr)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(n):
        y = gen(code=code, f0=f0, spkr=spkr)
    torch.cuda.synchronize(dev)
    eager_ms = (time.perf_counter() - t0) / n * 1e3
    gf = gen.capture_graph(1, 50, dev)
    yg = gf(code=code, f0=f0, spkr=spkr)
    torch.cuda.synchronize(dev)
    same = bool(torch.equal(yg.view(-1), y.view(-1)))
    t0 = time.perf_counter()
    for _ in range(n):
        gf()
    torch.cuda.synchronize(dev)
    graph_ms = (time.perf_counter() - t0) / n * 1e3
    rec = {"workload": "sr/inference.py unit of work: 1 synthetic utterance, 50 units -> 16 000 samples, Generator forward only "
                       "(BASELINE configs[0])",
           "latency_ms_cuda_graph": graph_ms, "latency_ms_eager": eager_ms, "graph_bit_identical_to_eager": same,
           "samples_per_s": 16000 / graph_ms * 1e3, "launches": gen.launches_per_forward()}
    if cpu_leg:
        from oracle import generator_oracle as go
        torch.set_num_threads(len(os.sched_getaffinity(0)))
        sd = go.folded_state_dict(syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0))
        c, f, s_ = code.cpu(), f0.cpu(), spkr.cpu()
        go.code_generator_forward(sd, syn.VCTK_CONFIG, c, f, s_)
        t0 = time.perf_counter()
        for _ in range(5):
            go.code_generator_forward(sd, syn.VCTK_CONFIG, c, f, s_)
        cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
        rec["cpu_baseline"] = {"value": 16000 / cpu_ms * 1e3, "unit": "samples/s", "latency_ms": cpu_ms,
                               "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "5 forwards of the oracle port on the same utterance (ATen/oneDNN, all host threads)"}
    return rec


def run_config3(gen, lm, pm, dev, rank, world, mx, utts=256, iters=3, vocode_batch=64):
    from dissc_b200.infer import convert_batch
    g = torch.Generator().manual_seed(1234 + rank)
    torch.manual_seed(1234 + rank)   # torch.distributions draws from the global generator
    seqs = synthetic_unit_sequences(utts, g).to(dev)
    spk = torch.randint(0, 108, (utts, 1), generator=g).to(dev)
    lm.norm_mean, lm.norm_std = torch.tensor(2.5), torch.tensor(0.05)

    def full():
        out_seq, f0, out_len = convert_batch(seqs, spk, 100, lm, pm, norm_pitch=True)
        return vocode_sorted(gen, out_seq, f0, spk, out_len, vocode_batch)[0]

    ms, n = timed(full, iters, dev)
    ms = mx(ms)
    msp, _ = timed(lambda: convert_batch(seqs, spk, 100, lm, pm, norm_pitch=True), iters, dev)
    msp = mx(msp)
    return {"workload": f"infer.py prosody path (LenPredictor + PitchPredictor) -> CodeGenerator, {utts} synthetic "
                        f"utterances per GPU (BASELINE configs[2])",
            "n_gpus": world, "utterances_per_gpu": utts, "iters": iters, "ms_per_step": ms, "prosody_only_ms": msp,
            "utterances_per_s": world * utts / ms * 1e3, "samples_per_s": world * n / ms * 1e3,
            "mean_output_frames": n / gen.hop / utts}


def run_config4(enc, hsd, dev, rank, world, mx, clips=32, n_samples=96000, iters=3, tensor_peak_tflops=None,
                cpu_leg=False):
    g = torch.Generator().manual_seed(1234 + rank)
    wave = (0.1 * torch.randn(clips, n_samples, generator=g)).to(dev)
    ms, (units, n_frames, _) = timed(lambda: enc.encode_batch(wave, return_dense=False), iters, dev)
    ms = mx(ms)
    cps = world * clips / ms * 1e3
    rec = {"workload": f"data/encode.py HuBERT-base layer-6 + k-means-100 units, {clips} synthetic {n_samples}-sample "
                       f"clips per GPU and step (BASELINE configs[3]: 8 000 clips = {8000 // max(1, world)} per GPU)",
           "n_gpus": world, "clips_per_gpu_step": clips, "iters": iters, "ms_per_step": ms, "clips_per_s": cps,
           "audio_samples_per_s": cps * n_samples, "s_for_8000_clips": 8000 / cps, "frames_per_clip": int(n_frames[0])}
    gflop = HUBERT_GFLOP_PER_CLIP * n_samples / 96000
    tf32eq = cps / world * gflop / 1e3
    rec["tflops_fp32_equivalent_per_gpu"] = tf32eq
    if tensor_peak_tflops:
        # every GEMM-shaped op of the encoder runs as three fp16 MMAs per fp32-accurate product (split precision)
        rec["roofline"] = {"bound": "tensor", "achieved": 3 * tf32eq, "peak": tensor_peak_tflops, "unit": "TFLOP/s",
                           "frac": 3 * tf32eq / tensor_peak_tflops,
                           "note": f"{gflop:.1f} GFLOP per clip x 3 fp16 MMAs / measured sustained cuBLAS bf16 rate"}
    if cpu_leg:
        from oracle import hubert_oracle as ho
        torch.set_num_threads(len(os.sched_getaffinity(0)))
        w1 = wave[:1].cpu()
        ho.extract_features(hsd, w1, 6)
        t0 = time.perf_counter()
        n_rep = 3
        for _ in range(n_rep):
            ho.extract_features(hsd, w1, 6)
        dt = (time.perf_counter() - t0) / n_rep
        rec["cpu_baseline"] = {"value": 1.0 / dt, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{n_rep} x 1 clip of {n_samples} samples through the oracle port "
                                         f"(oracle/hubert_oracle.py, ATen/oneDNN, all host threads), {dt * 1e3:.0f} ms each"}
    return rec


def run_config5(gen, lm, pm, enc, dev, rank, world, mx, clips=32, n_samples=96000, iters=3, vocode_batch=64):
    from dissc_b200.infer import convert_batch
    g = torch.Generator().manual_seed(4321 + rank)
    wave = (0.1 * torch.randn(clips, n_samples, generator=g)).to(dev)
    spk = torch.randint(0, 108, (clips, 1), generator=g).to(dev)
    # random centroids give unit sequences without run structure (dedup keeps ~all 299 frames), so the synthetic rhythm
    # statistics are set to ~1 frame per unit here: output duration ~ input duration (6 s), inside PitchPredictor's
    # 850-frame positional table (model/pitch_predictor.py:7)
    lm.norm_mean, lm.norm_std = torch.tensor(1.0), torch.tensor(0.02)

    def full():
        u, nf, _ = enc.encode_batch(wave, return_dense=False)
        s = u.clone()
        s[s < 0] = 100
        out_seq, f0, out_len = convert_batch(s, spk, 100, lm, pm, norm_pitch=True)
        return vocode_sorted(gen, out_seq, f0, spk, out_len, vocode_batch)[0]

    ms, n = timed(full, iters, dev)
    ms = mx(ms)
    ups = world * clips / ms * 1e3
    return {"workload": f"encode -> len / pitch predict -> vocode, {clips} synthetic {n_samples}-sample utterances per GPU "
                        f"and step (BASELINE configs[4]: 32 000 utterances = {32000 // max(1, world)} per GPU)",
            "n_gpus": world, "utterances_per_gpu_step": clips, "iters": iters, "ms_per_step": ms, "utterances_per_s": ups,
            "samples_vocoded_per_s": world * n / ms * 1e3, "s_for_32000_utterances": 32000 / ups,
            "mean_output_frames": n / gen.hop / clips}

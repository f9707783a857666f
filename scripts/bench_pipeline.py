"""BASELINE configs[2..4] on synthetic data, one process per GPU (weak scaling, utterance-sharded, no data-path collective):

  config 3  infer.py prosody path: LenPredictor + PitchPredictor -> CodeGenerator, 256 utterances per GPU
  config 4  data/encode.py: HuBERT-base layer-6 features + k-means-100 units, 96 000-sample clips (1 000 per GPU at full size)
  config 5  encode -> predict -> vocode chained on the same clips

    python scripts/bench_pipeline.py [--utts 256] [--clips 64] [--iters 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_pipeline.py

Every rank works on its own shard (seed 1234 + rank); times are CUDA events, max over ranks; rank 0 prints one JSON line
per config.  Weights: seeded synthetic checkpoints of the shipped geometries (no pretrained files offline).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissc_b200 import AttrDict, CodeGenerator, dist as ddist  # noqa: E402
from dissc_b200 import synthetic as syn  # noqa: E402
from dissc_b200.infer import convert_batch  # noqa: E402
from dissc_b200.predictors import LenPredictor, PitchPredictor  # noqa: E402


def timed(fn, iters, dev):
    fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=256, help="utterances per GPU for config 3")
    ap.add_argument("--clips", type=int, default=64, help="96 000-sample clips per GPU and step for configs 4 / 5")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--vocode-batch", type=int, default=64)
    a = ap.parse_args()
    rank, world, local = ddist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    import torch.distributed as dist

    def mx(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = torch.Generator().manual_seed(1234 + rank)
    # models
    gen = CodeGenerator(AttrDict(syn.VCTK_CONFIG)).to(dev)
    gen.load_state_dict(syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0))
    gen.eval()
    gen.remove_weight_norm()
    lm = LenPredictor(100, 108).to(dev)
    lm.load_state_dict(syn.synthetic_len_predictor_state_dict(100, 108, seed=21))
    # rhythm statistics: ~2.5 frames per deduplicated unit (VCTK-like).  The spread is kept small because the random-init
    # network's raw output is O(1..10): with the SURVEY's 1.5 most predictions clamp to 1 and the carry-over diffusion
    # then deletes units, which would shrink the vocoder's share of the step to ~50 frames per utterance
    lm.norm_mean, lm.norm_std = torch.tensor(2.5), torch.tensor(0.05)
    mean, std = syn.synthetic_pitch_stats(108, seed=22)
    pm = PitchPredictor(100, 108, id2pitch_mean=mean.to(dev), id2pitch_std=std.to(dev)).to(dev)
    pm.load_state_dict(syn.synthetic_pitch_predictor_state_dict("new", 100, 108, seed=23))

    def vocode(out_seq, f0, spk, out_len):
        """length-sorted sub-batches of --vocode-batch utterances -> total samples vocoded"""
        order = torch.argsort(out_len, descending=True)
        n = 0
        for i0 in range(0, len(order), a.vocode_batch):
            idx = order[i0:i0 + a.vocode_batch]
            L = int(out_len[idx[0]])
            code = out_seq[idx, :L].clone()
            code[code >= 100] = 0
            y = gen.generate_int16(code, f0[idx, :L].contiguous(), spk[idx], lengths=out_len[idx])
            n += int(out_len[idx].sum()) * gen.hop
        return n, y

    # ---- config 3: units (T ~ U{200..400}, run lengths ~ geometric mean 2.5) -> prosody -> vocoder
    B = a.utts
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
    seqs = seqs.to(dev)
    spk = torch.randint(0, 108, (B, 1), generator=g).to(dev)

    def config3():
        out_seq, f0, out_len = convert_batch(seqs, spk, 100, lm, pm, norm_pitch=True)
        n, _ = vocode(out_seq, f0, spk, out_len)
        return n

    ms3, n3 = timed(config3, a.iters, dev)
    ms3 = mx(ms3)

    def prosody_only():
        return convert_batch(seqs, spk, 100, lm, pm, norm_pitch=True)

    ms3p, _ = timed(prosody_only, a.iters, dev)
    ms3p = mx(ms3p)
    if rank == 0:
        print(json.dumps({"config": "3: infer.py prosody (len + pitch predictors) -> CodeGenerator", "n_gpus": world,
                          "utterances_per_gpu": B, "ms_per_step": ms3, "prosody_only_ms": ms3p,
                          "utterances_per_s": world * B / ms3 * 1e3, "samples_per_s": world * n3 / ms3 * 1e3,
                          "samples_vocoded_per_gpu_step": n3, "mean_output_frames": n3 / gen.hop / B}))

    # ---- rank 0 owns inputs and outputs: NCCL scatter of (code, f0, spkr, lengths) -> local vocode -> NCCL gather of the
    #      int16 waveforms (replaces the reference's Pool(8) + files on disk, sr/inference.py:351-359)
    sb = ddist.ShardedBatch(rank, world, dev)
    Bl, Tg = 64, 300
    if rank == 0:
        cg, fg, sg = syn.synthetic_inputs(world * Bl, Tg, seed=99)
        cg, fg, sg = cg.to(dev), fg.reshape(world * Bl, Tg).to(dev), sg.reshape(world * Bl).to(dev)
        lg = torch.full((world * Bl,), Tg, dtype=torch.int32, device=dev)
    else:
        cg = fg = sg = lg = None

    def scatter_vocode_gather():
        c, f, s_, l_ = sb.scatter(cg, fg, sg, lg, Bl, Tg)
        y = gen.generate_int16(c, f, s_, lengths=l_)
        return sb.gather(y)

    msg, yg = timed(scatter_vocode_gather, a.iters, dev)
    msg = mx(msg)
    if rank == 0:
        print(json.dumps({"config": "2 with rank 0 owning inputs/outputs: NCCL scatter -> vocode -> NCCL gather (int16)",
                          "n_gpus": world, "utterances": world * Bl, "units": Tg, "ms_per_step": msg,
                          "samples_per_s": world * Bl * Tg * gen.hop / msg * 1e3,
                          "gathered_bytes": int(yg.numel() * 2), "gathered_shape": list(yg.shape)}))

    # ---- config 4 / 5: clips -> HuBERT units (-> prosody -> vocoder)
    try:
        import torchaudio
        from dissc_b200.hubert import SpeechEncoder
        from oracle import hubert_oracle as ho   # weight-name mapping only (synthetic weights = torchaudio's init)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"config": "4/5", "skipped": repr(e)}))
        return
    torch.manual_seed(0)
    hsd = ho.from_torchaudio(torchaudio.models.hubert_base().eval(), 6)
    cent = torch.randn(100, 768, generator=torch.Generator().manual_seed(5))
    enc = SpeechEncoder.from_state_dict(hsd, cent).to(dev)
    C = a.clips
    wave = (0.1 * torch.randn(C, 96000, generator=g)).to(dev)
    spk5 = torch.randint(0, 108, (C, 1), generator=g).to(dev)

    def config4():
        return enc.encode_batch(wave, return_dense=False)

    ms4, (units, n_frames, _) = timed(config4, a.iters, dev)
    ms4 = mx(ms4)
    if rank == 0:
        print(json.dumps({"config": "4: data/encode.py HuBERT-base layer-6 + k-means-100 units", "n_gpus": world,
                          "clips_per_gpu_step": C, "ms_per_step": ms4, "clips_per_s": world * C / ms4 * 1e3,
                          "audio_samples_per_s": world * C * 96000 / ms4 * 1e3,
                          "s_for_8000_clips": 8000 / (world * C / ms4 * 1e3)}))

    # random centroids give unit sequences without run structure (dedup keeps ~all 299 frames), so the synthetic rhythm
    # statistics are set to ~1 frame per unit here: output duration ~ input duration (6 s), inside PitchPredictor's
    # 850-frame positional table (model/pitch_predictor.py:7)
    lm.norm_mean, lm.norm_std = torch.tensor(1.0), torch.tensor(0.02)

    def config5():
        u, nf, _ = enc.encode_batch(wave, return_dense=False)
        s = u.clone()
        s[s < 0] = 100
        out_seq, f0, out_len = convert_batch(s, spk5, 100, lm, pm, norm_pitch=True)
        n, _ = vocode(out_seq, f0, spk5, out_len)
        return n

    ms5, n5 = timed(config5, a.iters, dev)
    ms5 = mx(ms5)
    if rank == 0:
        print(json.dumps({"config": "5: encode -> len/pitch predict -> vocode", "n_gpus": world, "utterances_per_gpu_step": C,
                          "ms_per_step": ms5, "utterances_per_s": world * C / ms5 * 1e3,
                          "samples_vocoded_per_s": world * n5 / ms5 * 1e3,
                          "s_for_32000_utterances": 32000 / (world * C / ms5 * 1e3),
                          "mean_output_frames": n5 / gen.hop / C}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

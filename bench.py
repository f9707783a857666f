#!/usr/bin/env python
"""Headline benchmark: 16 kHz audio samples/s vocoded (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- CodeGenerator forward,
batch 64 VCTK-shape utterances x 300 units (96 000 samples each), per GPU
(weak scaling: every rank vocodes its own 64-utterance shard, no data-path
collective).  A step = one forward over one batch.

  value    device-resident inputs, CUDA-event time, max over ranks
  e2e      the C-ABI host entry (dissc_gen_forward_host_submit / _wait): pinned host
           inputs -> H2D -> forward -> D2H waveform, every step inside the timed region
  gathered the N-GPU data path: rank 0 owns inputs and outputs; one packed NCCL
           scatter -> forward -> NCCL gather of the int16 waveforms (overlapped)
  configs  short runs of BASELINE configs[2..4] (prosody -> vocode, HuBERT units,
           encode -> predict -> vocode) on this rank's shard
  roofline / cpu_baseline: see DESIGN.md "Measurement"

`--impl reference` times the reference's CPU path (the oracle port of
sr/models.py::CodeGenerator.forward, same ATen calls, all host threads) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "16kHz audio samples/sec vocoded"
UNIT = "samples/s"
B_PER_GPU, T_UNITS, HOP = 64, 300, 320
CPU_SAMPLE_B = 4
TRAFFIC_FILE = "r02_traffic.json"   # ncu launch list of one forward of the CURRENT kernels (scripts/ncu_traffic.py)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (pynvml, 50 ms period)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.mask, self.max_mhz, self.ok = [], 0, None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.mask |= int(fn(self.h))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self.t.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if self.mask & b], "samples": len(self.samples)}


def cpu_port_time(B, T, steps, warmup, threads=None):
    """Times the oracle port (reference algorithm, ATen/oneDNN on the host) -> (samples/s, ms/step, cores)."""
    from dissc_b200 import synthetic as syn
    from oracle import generator_oracle as go
    # torchrun exports OMP_NUM_THREADS=1: ask for every core this process may run on explicitly
    torch.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = torch.get_num_threads()
    sd = go.folded_state_dict(syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0))
    code, f0, spkr = syn.synthetic_inputs(B, T, seed=1234)
    for _ in range(warmup):
        go.code_generator_forward(sd, syn.VCTK_CONFIG, code, f0, spkr)
    t0 = time.perf_counter()
    for _ in range(steps):
        y = go.code_generator_forward(sd, syn.VCTK_CONFIG, code, f0, spkr)
    dt = (time.perf_counter() - t0) / steps
    return B * y.shape[-1] / dt, dt * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, ms, cores = cpu_port_time(CPU_SAMPLE_B, T_UNITS, args.steps, args.warmup)
    sample = f"B={CPU_SAMPLE_B} of the {B_PER_GPU}-utterance batch x {T_UNITS} units per step (identical seeds/weights)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CodeGenerator forward, VCTK geometry, 300 units/utterance (BASELINE configs[1])",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dissc_b200", choices=["dissc_b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="utterances per GPU (default = BASELINE config 2)")
    ap.add_argument("--units", type=int, default=T_UNITS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs[2..4] sub-records")
    ap.add_argument("--settle", type=float, default=1.5,
                    help="seconds of untimed back-to-back forwards before the extra `sustained` leg: under its power cap "
                         "the chip's clock keeps sinking for the first seconds of continuous load, so the legs measured "
                         "later (e2e, gathered, sustained) run 2-6%% below the one measured first (value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    from dissc_b200 import AttrDict, CodeGenerator, dist as ddist
    from dissc_b200 import synthetic as syn
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dissc_b200 has no CPU path")
    # NCCL may print its version banner on stdout when the first communicator is created: keep stdout to the ONE JSON
    # line by pointing fd 1 at stderr until the line itself is printed
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = ddist.init_from_env()
    if world > 1:
        torch.cuda.set_device(local)
        dist.all_reduce(torch.zeros(1, device=torch.device("cuda", local)))
        torch.cuda.synchronize()
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B, T = args.batch, args.units

    gen = CodeGenerator(AttrDict(syn.VCTK_CONFIG)).to(dev)
    gen.load_state_dict(syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0))
    gen.eval()
    gen.remove_weight_norm()
    code_h, f0_h, spkr_h = syn.synthetic_inputs(B, T, seed=1234 + rank)
    code, f0, spkr = code_h.to(dev), f0_h.to(dev), spkr_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------
    for _ in range(args.warmup):
        y = gen(code=code, f0=f0, spkr=spkr)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0.record()
    for _ in range(args.steps):
        y = gen(code=code, f0=f0, spkr=spkr)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    n_samples = y.shape[-1] * B
    value = world * n_samples / (ms_step * 1e-3)
    gen.check_indices(synchronize=False)

    # ---- end to end through the C-ABI host entry (pinned host buffers) ----
    # dissc_gen_forward_host_submit / _wait with two batches in flight: every step copies its inputs host -> device,
    # runs the forward and copies the waveform device -> host; step i's copy-back rides under step i+1's forward.
    code_p, f0_p = code_h.pin_memory(), f0_h.reshape(B, T).contiguous().pin_memory()
    spkr_p = spkr_h.reshape(B).contiguous().pin_memory()
    out_p = [torch.empty((B, gen.hop * T), dtype=torch.float32).pin_memory() for _ in range(2)]
    gen.host_reserve(B, T, device=local)
    for i in range(max(2, args.warmup)):
        gen.forward_host(code_p, f0_p, spkr_p, out=out_p[i % 2], device=local)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        gen.forward_host_submit(i % 2, code_p, f0_p, spkr_p, out=out_p[i % 2], device=local)
        if i > 0:
            gen.forward_host_wait((i - 1) % 2)       # step i-1's waveform is in host memory
    gen.forward_host_wait((args.steps - 1) % 2)
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    barrier()
    h2d = code_p.numel() * 8 + f0_p.numel() * 4 + spkr_p.numel() * 8
    d2h = out_p[0].numel() * 4
    e2e_val = world * n_samples / e2e_s
    y_host = y.reshape(B, -1).cpu()
    parity_e2e = bool(torch.equal(out_p[0], y_host) and torch.equal(out_p[1], y_host))

    # ---- the N-GPU data path of the north-star: rank 0 owns inputs and outputs -----------------------------------
    # one packed NCCL scatter -> forward (int16 written by conv_post into the gather's send buffer) -> NCCL gather on a
    # side stream, double-buffered against the next step (dissc_b200/dist.py::ScatterGatherPipeline).  At N = 1 the same
    # pipeline runs without collectives, which is the denominator of its scaling efficiency.
    def fwd_i16(c_, f_, s_, l_, out):
        gen.generate_int16(c_, f_, s_, lengths=l_, out=out)

    pipe = ddist.ScatterGatherPipeline(rank, world, dev, B, T, gen.hop, fwd_i16)
    packed = None
    if rank == 0:
        cg, fg, sg = syn.synthetic_inputs(world * B, T, seed=99)
        lg = torch.full((world * B,), T, dtype=torch.int32)
        packed = ddist.pack_inputs(cg.to(dev), fg.reshape(world * B, T).to(dev), sg.reshape(world * B).to(dev), lg.to(dev),
                                   world)
    for _ in range(max(3, args.warmup)):
        pipe.step(packed)
    pipe.flush()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    g0.record()
    for _ in range(args.steps):
        last = pipe.step(packed)
    pipe.flush()                                    # the compute stream waits for the outstanding gathers
    g1.record()
    torch.cuda.synchronize()
    barrier()
    ms_gath = max_over_ranks(g0.elapsed_time(g1)) / args.steps
    gathered = {"value": world * n_samples / (ms_gath * 1e-3), "unit": UNIT, "ms_per_step": ms_gath,
                "scatter_bytes_per_step": int(world * pipe.row_bytes), "gather_bytes_per_step": int(world * n_samples * 2),
                "collectives": "1 packed scatter + 1 int16 gather per step (NCCL), gather on a side stream / own "
                               "communicator, double-buffered" if world > 1 else "none (N=1: same pipeline, no collective)",
                "output": "int16 (N*B, hop*T) on rank 0"}
    if rank == 0:
        # rank 0's own rows of the gathered result = its local forward of the same inputs, bit for bit
        want = gen.generate_int16(cg[:B].to(dev), fg[:B].reshape(B, T).to(dev), sg[:B].reshape(B).to(dev),
                                  lengths=lg[:B].to(dev))
        gathered["bit_identical_to_local_forward"] = bool(torch.equal(pipe.gathered[last][0], want))
        if world > 1:   # and the last rank's rows = the forward of ITS inputs (run here)
            want_l = gen.generate_int16(cg[-B:].to(dev), fg[-B:].reshape(B, T).to(dev), sg[-B:].reshape(B).to(dev),
                                        lengths=lg[-B:].to(dev))
            gathered["bit_identical_last_rank"] = bool(torch.equal(pipe.gathered[last][world - 1], want_l))

    # ---- the same device-resident leg again after --settle seconds of continuous load (steady thermal / power state) --
    sustained = None
    if args.settle > 0:
        t_settle = time.perf_counter()
        while time.perf_counter() - t_settle < args.settle:
            for _ in range(5):
                y = gen(code=code, f0=f0, spkr=spkr)
            torch.cuda.synchronize()
        sampler2 = ClockSampler(local)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        sampler2.start()
        s0.record()
        for _ in range(args.steps):
            y = gen(code=code, f0=f0, spkr=spkr)
        s1.record()
        torch.cuda.synchronize()
        clocks2 = sampler2.stop()
        barrier()
        ms_sus = max_over_ranks(s0.elapsed_time(s1)) / args.steps
        sustained = {"value": world * n_samples / (ms_sus * 1e-3), "unit": UNIT, "ms_per_step": ms_sus,
                     "sm_mhz": clocks2.get("sm_mhz"),
                     "note": f"the `value` leg repeated after {args.settle:g} s more of back-to-back forwards on top of the "
                             "e2e and gathered legs: the power-capped clock has settled"}

    # ---- BASELINE configs[2..4]: short, bounded runs of the other configs on this rank's shard ---------------------
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    configs = {}
    if not args.no_configs:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        try:
            import bench_configs as bc
            if rank == 0:   # a latency, not a throughput: one rank measures it
                configs["configs[0]"] = bc.run_config1(gen, dev, cpu_leg=(world == 1 and not args.no_cpu_baseline))
            lm, pm = bc.build_predictors(dev)
            configs["configs[2]"] = bc.run_config3(gen, lm, pm, dev, rank, world, max_over_ranks, utts=256, iters=2)
            enc, hsd = bc.build_encoder(dev)
            configs["configs[3]"] = bc.run_config4(enc, hsd, dev, rank, world, max_over_ranks, clips=32, iters=3,
                                                   tensor_peak_tflops=tensor_peak,
                                                   cpu_leg=(rank == 0 and world == 1 and not args.no_cpu_baseline))
            configs["configs[4]"] = bc.run_config5(gen, lm, pm, enc, dev, rank, world, max_over_ranks, clips=32, iters=2)
        except Exception as e:  # noqa: BLE001 -- the headline line must still be printed
            configs["error"] = repr(e)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_fd, 1)
    os.close(saved_fd)
    if rank != 0:
        return
    # ---- roofline -------------------------------------------------------------------------
    # Dominant kernel = the kernel family with the largest share of the step.  The path is a dense contraction
    # (85 FLOP/B): the tensor pipe binds, so `roofline` is quoted against the measured sustained tensor rate --
    # achieved = 3 x the algorithmic FLOPs of the family's launches (three fp16 MMAs per fp32-accurate product)
    # / their summed CUDA-event durations, measured on one extra forward with an event pair around every launch on the
    # launching stream.  `hbm` keeps the algorithmic-bytes fraction BASELINE.json's metric asks for (same launches,
    # layer-fused traffic model, SURVEY.md 8d / DESIGN.md), `forward` the whole step over the timed region.
    flops, abytes = gen.cost(B, T, dev)
    peak_gbs, sm_max_mhz, peak_src = measured_peaks()
    rows = gen.profile(code, f0, spkr)  # one extra, untimed forward with an event pair around every launch
    fam = {}
    for name, ms, fl, by in rows:
        key = "conv_tc_kernel" if name.endswith(".tc") else "resblock_pair_tc_kernel" if name.endswith(".ptc") \
            else "resblock_pair64_tc_kernel" if name.endswith(".p64") else "resblock_pack2_tc_kernel" \
            if name.endswith(".pk2") else "conv_post_kernel" if name == "conv_post" \
            else "convt1d_kernel" if name.startswith("ups") else "tc_embed_planes+tc_zero_halos" \
            if name in ("embed", "zero_halos") else "conv1d_fused_kernel"
        a = fam.setdefault(key, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        a["launches"] += 1
        a["ms"] += ms
        a["flops"] += fl
        a["bytes"] += by
    tot_ms = sum(a["ms"] for a in fam.values())
    for a in fam.values():
        a["share"] = round(a["ms"] / tot_ms, 4)
        a["tflops"] = round(a["flops"] / a["ms"] / 1e9, 2)
        a["algorithmic_gbs"] = round(a["bytes"] / a["ms"] / 1e6, 1)
        a["ms"] = round(a["ms"], 3)
        del a["flops"], a["bytes"]
    dom_name = max(fam, key=lambda k: fam[k]["ms"])
    dom = fam[dom_name]
    sfx = {"conv_tc_kernel": ".tc", "resblock_pair_tc_kernel": ".ptc", "resblock_pair64_tc_kernel": ".p64",
           "resblock_pack2_tc_kernel": ".pk2"}.get(dom_name, "")
    dom_rows = [r for r in rows if r[0].endswith(sfx)]
    dom_bytes = sum(r[3] for r in dom_rows)
    dom_ms = sum(r[1] for r in dom_rows)
    dom_flops = sum(r[2] for r in dom_rows)
    n_dom = max(1, len(dom_rows))
    ach_gbs = dom_bytes / (dom_ms * 1e-3) / 1e9
    fwd_gbs = abytes / (ms_step * 1e-3) / 1e9
    split_tflops = 3.0 * dom_flops / (dom_ms * 1e-3) / 1e12   # three fp16 MMAs per fp32-accurate product
    fwd_tflops = 3.0 * flops / (ms_step * 1e-3) / 1e12
    # measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant kernel, from the
    # committed ncu launch list of the same command on the current kernels (scripts/ncu_traffic.py); null otherwise
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", TRAFFIC_FILE)
    if (B, T) == (B_PER_GPU, T_UNITS) and os.path.isfile(tpath):
        tj = json.load(open(tpath))
        tk = tj["kernels"].get(dom_name)
        if tk and tk["launches"] == len(dom_rows):
            traffic = tk["dram_bytes_per_launch"]
            traffic_src = f"profiles/{TRAFFIC_FILE} (ncu launch list of one forward, {tj.get('commit', 'this round')})"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"CodeGenerator forward, batch={B} VCTK-shape utterances x {T} units "
                               f"(x{gen.hop} -> {gen.hop * T} samples each) per GPU (BASELINE configs[1])",
                   "weights": "seeded synthetic checkpoint, shipped VCTK geometry (13.7M params)",
                   "arithmetic": "fp32 results (<=1e-4 max-abs vs the oracle, tests/test_generator_gpu.py) from "
                                 "split-fp16 tcgen05 MMAs, fp32 accumulate",
                   "l2": "every layer's working set (>=0.8 GB at B=64) exceeds the 126 MB L2; no flush needed",
                   "parallelism": f"utterance-sharded x{world}"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3,
                "api": "dissc_gen_forward_host_submit / _wait (C ABI, pinned host buffers, two batches in flight: "
                       "every step does H2D + forward + D2H, a step's D2H overlaps the next step's forward)",
                "bit_identical_to_device_path": parity_e2e},
        "gathered": gathered,
        "sustained": sustained,
        "gpu_launches": gen.launches_per_forward() * args.steps,
        "roofline": {"bound": "tensor", "achieved": split_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                     "frac": split_tflops / tensor_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16, sustained: the kernel is "
                                    "timed inside a long step)",
                     "kernel": f"{dom_name} ({dom['launches']} of {len(rows)} launches per forward, "
                               f"{100 * dom['share']:.1f}% of the step)",
                     "algorithmic_flops_per_launch_avg": dom_flops / n_dom,
                     "mma_flops_per_launch_avg": 3.0 * dom_flops / n_dom,
                     "algorithmic_bytes_per_launch_avg": dom_bytes / n_dom,
                     "launch_ms_avg": dom_ms / n_dom,
                     "why_tensor": "dense contraction, 85 FLOP/B algorithmic; three fp16 MMAs per fp32-accurate product "
                                   "put the tensor floor (13.4 ms) above the HBM floor (11.1 ms); ncu: tensor pipe "
                                   "80-94% busy in the k>=7 layers, DRAM 7-40% (profiles/README.md)",
                     "hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": peak_gbs, "unit": "GB/s",
                             "frac": ach_gbs / peak_gbs, "peak_source": peak_src,
                             "note": "same launches, algorithmic bytes of the layer-fused traffic model"},
                     "forward": {"hbm": {"achieved": fwd_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": fwd_gbs / peak_gbs,
                                         "algorithmic_bytes_per_step": abytes},
                                 "tensor": {"achieved": fwd_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                                            "frac": fwd_tflops / tensor_peak, "algorithmic_flops_per_step": flops},
                                 "note": "whole step over the timed region (all kernels)"},
                     "kernels": fam},
    }
    if configs:
        line["configs"] = configs
    if world == 1 and not args.no_cpu_baseline:
        val, ms, cores = cpu_port_time(CPU_SAMPLE_B, T, steps=3, warmup=1)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"B={CPU_SAMPLE_B} slice of the batch x {T} units, 3 timed forwards "
                                          f"({ms:.0f} ms each) of the oracle port (ATen/oneDNN, all host threads)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()

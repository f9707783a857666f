"""In-process A/B of plan-time tuning knobs (dissc_tc_set_tuning): one generator per variant in the SAME process, forwards
timed alternately (A, B, A, B ...) so clock / thermal drift and box-to-box spread cancel.

    python scripts/ab_tuning.py "base:1=0" "split:1=1" "split_na3:1=1,0=3"        (label:key=value,key=value)
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissc_b200 import AttrDict, CodeGenerator, _lib  # noqa: E402
from dissc_b200 import synthetic as syn  # noqa: E402


def main():
    variants = []
    for v in sys.argv[1:]:
        label, _, kv = v.partition(":")
        variants.append((label, [tuple(int(x) for x in p.split("=")) for p in kv.split(",") if p]))
    dev = torch.device("cuda", 0)
    B, T = 64, 300
    sd = syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0)
    code, f0, spkr = (t.to(dev) for t in syn.synthetic_inputs(B, T))
    gens = []
    for label, kvs in variants:
        for k, val in kvs:
            _lib.check(_lib.lib().dissc_tc_set_tuning(k, val))
        g = CodeGenerator(AttrDict(syn.VCTK_CONFIG)).to(dev)
        g.load_state_dict(sd)
        g.eval()
        g.remove_weight_norm()
        y = g(code=code, f0=f0, spkr=spkr)   # creates the handle (plans read the knobs now)
        gens.append((label, g, y.clone()))
    torch.cuda.synchronize()
    for label, g, y in gens[1:]:
        print(f"{label}: bit-identical to {gens[0][0]}: {bool(torch.equal(y, gens[0][2]))}")
    times = {label: [] for label, _, _ in gens}
    import random
    random.seed(0)
    for r in range(int(os.environ.get("AB_ROUNDS", "8"))):
        order = list(gens)
        random.shuffle(order)
        for label, g, _ in order:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                g(code=code, f0=f0, spkr=spkr)
            e1.record()
            torch.cuda.synchronize()
            times[label].append(e0.elapsed_time(e1) / 4)
    for label, ts in times.items():
        print(f"{label:14s} median {statistics.median(ts):7.3f} ms  mean {statistics.mean(ts):7.3f}  min {min(ts):7.3f}   n={len(ts)}")
    # per-stage medians from the profiler, alternating as well
    rows = {label: [] for label, _, _ in gens}
    for r in range(3):
        for label, g, _ in gens:
            rows[label].append(g.profile(code, f0, spkr))
    stages = ["ups", "s0", "s1", "s2", "s3", "s4"]
    for label in rows:
        agg = {}
        for st in stages:
            agg[st] = statistics.median(sum(ms for name, ms, _, _ in run if name.split(".")[0] == st) for run in rows[label])
        print(f"{label:14s} " + "  ".join(f"{st} {agg[st]:.3f}" for st in stages))


if __name__ == "__main__":
    main()

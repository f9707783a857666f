"""A/B per-layer timings: runs scripts/profile_layers.py as a subprocess once per (variant, round), variants interleaved
round-robin so clock / thermal drift hits all of them alike, and prints the per-layer and per-stage medians.

    python scripts/ab_layers.py --rounds 3 base: fast:DISSC_TC_PAIR64=1 ...      (label:ENV=VAL,ENV=VAL)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="+")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("-B", type=int, default=64)
    ap.add_argument("-T", type=int, default=300)
    a = ap.parse_args()
    variants = []
    for v in a.variants:
        label, _, envs = v.partition(":")
        env = dict(kv.split("=", 1) for kv in envs.split(",") if kv)
        variants.append((label, env))
    times = {label: {} for label, _ in variants}
    order = []
    for r in range(a.rounds):
        for label, env in variants:
            e = dict(os.environ)
            e.update(env)
            out = subprocess.run([sys.executable, os.path.join(HERE, "profile_layers.py"), str(a.B), str(a.T)], env=e,
                                 capture_output=True, text=True, timeout=600)
            if out.returncode != 0:
                print(f"[{label}] FAILED rc={out.returncode}\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
                continue
            rows = json.load(open(f"gpurun_out/layers_B{a.B}_T{a.T}.json"))
            seen = {}
            for name, ms, fl, by in rows:
                n = seen.get(name, 0)
                seen[name] = n + 1
                key = name if n == 0 else f"{name}#{n}"
                if key not in order:
                    order.append(key)
                times[label].setdefault(key, []).append(ms)
    labels = [l for l, _ in variants]
    med = {l: {k: statistics.median(v) for k, v in times[l].items()} for l in labels}
    print(f"{'layer':24s}" + "".join(f"{l:>12s}" for l in labels))
    for k in order:
        print(f"{k:24s}" + "".join(f"{med[l].get(k, float('nan')):12.3f}" for l in labels))
    print("---- per stage")
    stages = []
    for k in order:
        st = k.split(".")[0].split("#")[0]
        if st not in stages:
            stages.append(st)
    for st in stages:
        print(f"{st:24s}" + "".join(
            f"{sum(v for k, v in med[l].items() if k.split('.')[0].split('#')[0] == st):12.3f}" for l in labels))
    print(f"{'TOTAL':24s}" + "".join(f"{sum(med[l].values()):12.3f}" for l in labels))


if __name__ == "__main__":
    main()

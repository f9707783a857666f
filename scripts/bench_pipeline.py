"""BASELINE configs[2..4] at full per-GPU size, one JSON line per config (workloads: scripts/bench_configs.py; bench.py
reports the same records, with bounded iteration counts, in its `configs` sub-record).

    python scripts/bench_pipeline.py [--utts 256] [--clips 64] [--iters 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_pipeline.py
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc  # noqa: E402
from dissc_b200 import dist as ddist  # noqa: E402


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

    gen = bc.build_generator(dev)
    lm, pm = bc.build_predictors(dev)
    recs = [bc.run_config3(gen, lm, pm, dev, rank, world, mx, a.utts, a.iters, a.vocode_batch)]
    enc, hsd = bc.build_encoder(dev)
    recs.append(bc.run_config4(enc, hsd, dev, rank, world, mx, a.clips, 96000, a.iters, cpu_leg=(rank == 0 and world == 1)))
    recs.append(bc.run_config5(gen, lm, pm, enc, dev, rank, world, mx, a.clips, 96000, a.iters, a.vocode_batch))
    if rank == 0:
        for r in recs:
            print(json.dumps(r))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

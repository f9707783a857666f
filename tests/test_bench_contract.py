"""CPU check of bench.py's reference arm (the line the driver parses): one JSON object on stdout with the contract's keys.
The GPU arm's line is produced on the GPU box (profiles/r01_*_bench.json hold examples)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "16kHz audio samples/sec vocoded" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_committed_gpu_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
              "gathered", "configs"):
        assert k in d, k
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["unit"] == {"hbm": "GB/s", "tensor": "TFLOP/s"}[r["bound"]]
    assert r["traffic"] is None or r["traffic"] > 0
    assert r["hbm"]["unit"] == "GB/s" and abs(r["hbm"]["frac"] - r["hbm"]["achieved"] / r["hbm"]["peak"]) < 1e-9
    g = d["gathered"]
    assert g["value"] > 0 and g["bit_identical_to_local_forward"] is True
    assert set(d["configs"]) >= {"configs[2]", "configs[3]", "configs[4]"}
    assert d["configs"]["configs[3]"]["roofline"]["bound"] == "tensor"
    t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))   # what bench.py reads for roofline.traffic
    assert t["kernels"]["conv_tc_kernel"]["launches"] == 42 and t["kernels"]["conv_tc_kernel"]["dram_bytes_per_launch"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1

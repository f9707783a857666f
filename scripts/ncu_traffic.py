"""Aggregate an ncu launch list (CSV of gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch)
into per-kernel totals -> profiles/<tag>_traffic.json, the file bench.py reads for `roofline.traffic`.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python scripts/one_forward.py 64 300 1
    python scripts/ncu_traffic.py gpurun_out/launches.csv profiles/r01_d_traffic.json
"""
import csv
import json
import re
import sys


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m.get(unit, 1)


def to_us(v, unit):
    m = {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}
    return float(v.replace(",", "")) * m.get(unit, 1)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    lines = open(src).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    launches = {}
    for r in rows:
        lid = int(r["ID"])
        d = launches.setdefault(lid, {"kernel": r["Kernel Name"]})
        name, unit, val = r["Metric Name"], r["Metric Unit"], r["Metric Value"]
        if name == "gpu__time_duration.sum":
            d["us"] = to_us(val, unit)
        elif name == "dram__bytes_read.sum":
            d["rd"] = to_bytes(val, unit)
        elif name == "dram__bytes_write.sum":
            d["wr"] = to_bytes(val, unit)
    fam = {}
    for lid in sorted(launches):
        d = launches[lid]
        k = re.sub(r"^(void )?(dissc::)?", "", d["kernel"])
        k = re.sub(r"[<(].*$", "", k)
        a = fam.setdefault(k, {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        a["launches"] += 1
        a["us"] += d.get("us", 0.0)
        a["dram_read_bytes"] += d.get("rd", 0.0)
        a["dram_write_bytes"] += d.get("wr", 0.0)
    tot = sum(a["us"] for a in fam.values())
    for a in fam.values():
        a["share"] = round(a["us"] / tot, 4)
        a["dram_bytes_per_launch"] = (a["dram_read_bytes"] + a["dram_write_bytes"]) / a["launches"]
    out = {"source": src, "command": "python scripts/one_forward.py 64 300 1 (one forward, BASELINE configs[1])",
           "total_us": tot, "total_dram_bytes": sum(a["dram_read_bytes"] + a["dram_write_bytes"] for a in fam.values()),
           "kernels": fam}
    json.dump(out, open(dst, "w"), indent=1)
    for k, a in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{k:34s} {a['launches']:3d} launches {a['us'] / 1e3:8.3f} ms {100 * a['share']:5.1f}%  "
              f"{(a['dram_read_bytes'] + a['dram_write_bytes']) / 1e9:7.2f} GB DRAM")
    print(f"total {tot / 1e3:.3f} ms, {out['total_dram_bytes'] / 1e9:.2f} GB")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per CUDA kernel name the launches, total / average duration, share of the serialised kernel time and (when the DRAM
metrics were collected) the average DRAM traffic per launch.

    tools/ncu_summarize.py launches.csv [--steps K] [--json profiles/ncu_traffic.json] > summary.txt

`--steps K`: the capture covers K identical steps (reported per step).  `--json`: also write {kernel name: average
DRAM bytes per launch}, the file bench.py reads for `roofline.traffic`.
"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)                     # drop the parameter list
    name = name.replace("rnvp::", "").replace("<unnamed>::", "")
    return re.sub(r"\s+", "", name)


def main():
    args = sys.argv[1:]
    path = args[0]
    steps = int(args[args.index("--steps") + 1]) if "--steps" in args else 1
    jpath = args[args.index("--json") + 1] if "--json" in args else None
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: defaultdict(float))
    ids = defaultdict(set)
    for r in rows[1:]:
        k = short(r[ix["Kernel Name"]])
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        m = r[ix["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        elif m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        per[k][m] += v
        ids[k].add(r[ix["ID"]])
    tot = sum(p["gpu__time_duration.sum"] for p in per.values())
    n_launch = sum(len(v) for v in ids.values())
    print(f"{path}: {n_launch} launches, {tot / 1e3:.2f} ms of serialised kernel time over {steps} step(s) "
          f"(cold-cache, serialised: compare SHARES, not absolutes)")
    has_dram = any("dram__bytes_read.sum" in p for p in per.values())
    traffic = {}
    print(f"{'kernel':64s} {'launches/step':>13s} {'ms/step':>9s} {'share':>7s} {'avg us':>8s}" + ("  avg DRAM MB/launch" if has_dram else ""))
    for k, p in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(ids[k])
        t = p["gpu__time_duration.sum"]
        line = f"{k[:64]:64s} {n / steps:13.1f} {t / 1e3 / steps:9.3f} {t / tot:7.1%} {t / n:8.1f}"
        if has_dram:
            b = (p["dram__bytes_read.sum"] + p["dram__bytes_write.sum"]) / n
            traffic[k] = b
            line += f"  {b / 1e6:10.2f}"
        print(line)
    if jpath and has_dram:
        out = {"_source": f"{path}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                          "--clock-control none; dram__bytes_read.sum + dram__bytes_write.sum averaged over the launches "
                          "of each kernel name in one training step (bench.py --steps 1)"}
        out.update({k: round(v) for k, v in traffic.items()})
        json.dump(out, open(jpath, "w"), indent=1)


if __name__ == "__main__":
    main()

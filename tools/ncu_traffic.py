#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list:
per kernel name -> launches, total time, DRAM bytes read + written.  Writes JSON (for bench.py's roofline.traffic) and
prints a table.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/launches_dram.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --profile-only
    python tools/ncu_traffic.py gpurun_out/launches_dram.csv profiles/r01_dram_traffic.json --steps 1
"""
import csv, json, re, sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 1
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).split("::")[-1].strip()
        name = re.sub(r"<.*", "", name)
        metric, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
        d = per.setdefault(name, {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0})
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1)
        if metric == "gpu__time_duration.sum":
            d["launches"] += 1; d["time_us"] += val * scale
        elif metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            d["dram_bytes"] += val * scale
    if "--steps" not in sys.argv:                       # one top-k launch per step
        steps = max(1, per.get("sigmoid_topk_kernel", {"launches": 1})["launches"])
    tot_t = sum(d["time_us"] for d in per.values())
    out = {"source": src, "steps": steps, "note": "ncu launch list: cold-cache, serialised launches; shares of the step are meaningful, absolute times are not",
           "kernels": {k: {"launches_per_step": v["launches"] / steps, "time_us_per_step": v["time_us"] / steps, "share": v["time_us"] / tot_t,
                           "dram_bytes_per_step": v["dram_bytes"] / steps} for k, v in sorted(per.items(), key=lambda t: -t[1]["time_us"])}}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out["kernels"].items():
        print("%-24s %5.0f launches %9.1f us %5.1f%% %9.1f MB DRAM" % (k, v["launches_per_step"], v["time_us_per_step"], 100 * v["share"], v["dram_bytes_per_step"] / 1e6))


if __name__ == "__main__":
    main()

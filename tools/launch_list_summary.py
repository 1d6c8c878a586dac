"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum[,lts__t_bytes.sum]` launch
list (one row per metric per launch) into a per-kernel table (markdown on stdout) and, with --json KEY OUT, add the average DRAM
bytes per conv launch to OUT[KEY] (read by bench.py as roofline.traffic).

    python tools/launch_list_summary.py gpurun_out/launches.csv [--json res101_b1 profiles/r02_conv_traffic.json] [--note "..."]
"""
import csv
import json
import os
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("frcnn::", "")
    m = re.match(r"([A-Za-z0-9_:]+(<[^>(]*>)?)", name)
    return m.group(1) if m else name[:60]


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iv, iu, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    launches = {}
    for r in rd:
        d = launches.setdefault(int(r[ii]), {"name": short(r[ik])})
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
                 "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r[im]] = v * scale
    agg = {}
    for d in launches.values():
        a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "l2": 0.0})
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["l2"] += d.get("lts__t_bytes.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    print("| kernel | launches | time us | share | DRAM read MB | DRAM write MB | L2 traffic MB | DRAM GB/s |")
    print("|---|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = (a["rd"] + a["wr"]) / a["us"] / 1e3 if a["us"] else 0.0
        print("| %s | %d | %.1f | %.1f%% | %.1f | %.1f | %.1f | %.0f |" % (k, a["n"], a["us"], 100 * a["us"] / tot, a["rd"] / 1e6, a["wr"] / 1e6,
                                                                        a["l2"] / 1e6, gbs))
    print("\nTotal %.0f us over %d launches." % (tot, len(launches)))
    if "--json" in sys.argv:
        i = sys.argv.index("--json")
        key, out = sys.argv[i + 1], sys.argv[i + 2]
        note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
        conv = [a for k, a in agg.items() if k.startswith("conv_gemm") or k.startswith("tail_reduce")]
        nconv = sum(a["n"] for k, a in agg.items() if k.startswith("conv_gemm"))
        byts = sum(a["rd"] + a["wr"] for a in conv)
        j = json.load(open(out)) if os.path.exists(out) else {}
        j[key] = {"dram_bytes_per_conv_launch_avg": byts / max(nconv, 1), "conv_launches": nconv, "dram_bytes_conv_total": byts,
                  "dram_bytes_all_kernels": sum(a["rd"] + a["wr"] for a in agg.values()),
                  "note": note or "dram__bytes_read+write of the conv_gemm launches (+ their tail_reduce passes) of one step / number of conv launches; source %s" % os.path.basename(path)}
        with open(out, "w") as f:
            json.dump(j, f, indent=1)


if __name__ == "__main__":
    main()

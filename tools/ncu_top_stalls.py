"""Read an `ncu --page source --csv --print-source sass` dump and list the SASS instructions with the most warp-stall samples.

    ncu -i rep.ncu-rep --page source --csv --print-source sass --kernel-id :::N > src.csv ; python tools/ncu_top_stalls.py src.csv [top]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ia = hdr.index("Warp Stall Sampling (All Samples)")
    ie = hdr.index("Instructions Executed")

    def n(v):
        try:
            return int(v)
        except ValueError:
            return 0
    tot = sum(n(r[ia]) for r in data)
    print("total samples", tot, "instructions", len(data))
    top = sorted(range(len(data)), key=lambda k: -n(data[k][ia]))[:top_n]
    for k in sorted(top):
        r = data[k]
        print("%5d %6d %5.1f%% exec=%-8s %s" % (k, n(r[ia]), 100.0 * n(r[ia]) / max(tot, 1), r[ie], r[1][:100]))


if __name__ == "__main__":
    main()

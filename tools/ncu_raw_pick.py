"""Print selected metrics of every kernel in an `ncu --page raw --csv` export (one column per kernel).

    python tools/ncu_raw_pick.py raw.csv [metric-prefix ...]
"""
import csv
import sys

DEFAULT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.max"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2:] or DEFAULT
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    names = [d[ik].replace("frcnn::", "").replace("void ", "")[:28] for d in data]
    print("%-70s %-8s " % ("metric", "unit") + " ".join("%14s" % n[:14] for n in names))
    for i, h in enumerate(hdr):
        if any(h == w or (w.endswith("*") and h.startswith(w[:-1])) for w in want):
            print("%-70s %-8s " % (h[:70], units[i][:8]) + " ".join("%14s" % d[i][:14] for d in data))


if __name__ == "__main__":
    main()

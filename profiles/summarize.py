"""Turn gpurun_out ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/rN_launches.txt
    python profiles/summarize.py raw      gpurun_out/raw_X.csv       > profiles/rN_kernels.txt
(raw csv = `ncu -i prof.ncu-rep --page raw --csv`)
"""
import collections
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    d = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        d[row["Kernel Name"][:90]].append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-92s %5s %10s %10s %6s" % ("kernel", "n", "avg_us", "total_us", "share"))
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("%-92s %5d %10.1f %10.1f %5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none, one launch per kernel")
    for d in data:
        print("== " + d[idx["Kernel Name"]][:110])
        for w in WANT:
            if w in idx:
                print("   %-66s %16s %s" % (w, d[idx[w]][:16], units[idx[w]]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])

"""Summarises ncu outputs (launch list CSV + raw page of a .ncu-rep) into the text files kept in profiles/."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, tot = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"us": 1.0, "ns": 1e-3, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
        k = d["Kernel Name"][:60]
        tot[k][0] += 1
        tot[k][1] += v
    s = sum(v[1] for v in tot.values())
    out = ["kernel | launches | total us | avg us | share of listed GPU time"]
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append("%-62s %5d %12.1f %9.1f %6.3f" % (k, n, t, t / n, t / s))
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "sm__cycles_elapsed.max ",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ", "smsp__inst_executed.sum ",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum ", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"]


def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for i, h in enumerate(hdr):
        if any((h + " ").startswith(w) or w.strip() == h for w in WANT) or "issue_stalled" in h and "pct" in h:
            out.append("%-90s %-8s %s" % (h, units[i], [r[i] for r in rows[2:5]]))
    return "\n".join(out)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2]))
    else:
        print(raw(sys.argv[2]))

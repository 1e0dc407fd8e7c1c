#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

  python profiles/summarize.py launches gpurun_out/launches_x.csv            > profiles/rN_launch_shares_x.txt
  python profiles/summarize.py metrics  gpurun_out/prof_x.ncu-rep [regex]    > profiles/rN_ncu_metrics_x.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'lts__t_bytes.sum']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        name = r[ki].split('(')[0][:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# per-kernel totals of gpu__time_duration.sum (ns) from %s; cold-cache, serialised: compare SHARES" % path)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s launches=%5d  total=%10.3f ms  share=%5.1f%%" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
    print("total %.3f ms" % (tot / 1e6))


def metrics(path, regex=None):
    cmd = ["ncu", "-i", path, "--page", "raw", "--csv"]
    if regex:
        cmd += ["--kernel-name", "regex:" + regex]
    rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    ki = hdr.index('Kernel Name')
    print("# selected metrics from %s (ncu --set full --clock-control none)" % path)
    for r in rows[2:]:
        print(r[ki][:90])
        for k, i in idx:
            print("    %-66s %s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "metrics": metrics}[sys.argv[1]](*sys.argv[2:])

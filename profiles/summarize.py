#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

  python profiles/summarize.py launches gpurun_out/launches_x.csv            > profiles/rN_launch_shares_x.txt
  python profiles/summarize.py metrics  gpurun_out/prof_x.ncu-rep [regex]    > profiles/rN_ncu_metrics_x.txt
  python profiles/summarize.py traffic  gpurun_out/prof_x.ncu-rep <candidates per launch> "note" > profiles/rN_traffic.json
      (per kernel: mean DRAM read+write bytes and duration per launch; bench.py copies it into roofline.traffic)
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
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


def raw_rows(path, regex=None):
    """rows of `ncu --page raw --csv`: from a report, or from a csv already exported on the GPU box (reports of
    more than ~25 kernels exceed what gpurun brings back)"""
    if path.endswith(".csv"):
        return [r for r in csv.reader(open(path)) if len(r) > 10]
    cmd = ["ncu", "-i", path, "--page", "raw", "--csv"]
    if regex:
        cmd += ["--kernel-name", "regex:" + regex]
    return list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))


def metrics(path, regex=None):
    rows = raw_rows(path, regex)
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    ki = hdr.index('Kernel Name')
    print("# selected metrics from %s (ncu --set full --clock-control none)" % path)
    for r in rows[2:]:
        print(r[ki][:90])
        for k, i in idx:
            print("    %-66s %s %s" % (k, r[i], units[i]))


UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0,
        'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0}


def traffic(path, candidates_per_launch, note=""):
    import json
    import re
    rows = raw_rows(path)
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum')}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r'\(.*$', '', r[col['Kernel Name']]).replace('void ', '').replace('cto::', '').replace('tc::', '').strip()
        rd = float(r[col['dram__bytes_read.sum']]) * UNIT[units[col['dram__bytes_read.sum']]]
        wr = float(r[col['dram__bytes_write.sum']]) * UNIT[units[col['dram__bytes_write.sum']]]
        t = float(r[col['gpu__time_duration.sum']]) * UNIT[units[col['gpu__time_duration.sum']]]
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += rd; a[2] += wr; a[3] += t
    out = {"source": "%s (ncu --set full --clock-control none; cold cache, per launch)" % path, "note": note,
           "candidates_per_launch": int(candidates_per_launch), "kernels": {}}
    for k, a in agg.items():
        out["kernels"][k] = dict(launches=a[0], dram_read_bytes=a[1] / a[0], dram_write_bytes=a[2] / a[0],
                                 dram_bytes=(a[1] + a[2]) / a[0], seconds=a[3] / a[0])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "metrics": metrics, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])

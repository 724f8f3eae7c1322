"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the small, tracked
summaries under profiles/: the launch list, key metrics per kernel, and traffic.json (DRAM
bytes per launch of each of our kernels, read by bench.py for roofline.traffic).

    python tools/summarize_ncu.py <launches.csv> <full.ncu-rep> <tag>
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}


def main(launch_csv, rep, tag):
    out = [f"# ncu summary {tag}\n"]
    rows = [r for r in csv.reader(l for l in open(launch_csv) if l.startswith('"'))]
    h = rows[0]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    d = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki].split('(')[0][:70]].append(float(r[vi].replace(',', '')))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    out += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, "
            "serialised: compare shares)\n", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v) / 1e6:.3f} | {sum(v) / tot:.3f} |")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rawpath = os.path.join(ROOT, "profiles", f"{tag}_ncu_raw.csv")
    with open(rawpath, "w") as f:
        f.write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index('Kernel Name')
    per = collections.OrderedDict()
    for vals in rows[2:]:
        per.setdefault(vals[ik].split('(')[0].replace('void ', ''), []).append(vals)
    traffic = {}
    out.append("\n## Full captures (`ncu --set full --clock-control none --import-source on`)\n")
    for name, insts in per.items():
        last = insts[-1]

        def f(k):
            return float(last[hdr.index(k)].replace(',', ''))
        tr = f('dram__bytes_read.sum') * SCALE[units[hdr.index('dram__bytes_read.sum')]] + \
            f('dram__bytes_write.sum') * SCALE[units[hdr.index('dram__bytes_write.sum')]]
        traffic[name.split('<')[0]] = {"dram_bytes_per_launch": tr, "kernel": name}
        out += [f"### `{name}` (last of {len(insts)} captured launches)\n", "| metric | value | unit |", "|---|---|---|"]
        for k in WANT:
            if k in hdr:
                out.append(f"| {k} | {last[hdr.index(k)]} | {units[hdr.index(k)]} |")
        out.append("")
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    merged = {}
    if os.path.exists(tpath):           # captures of different kernels come from different runs
        with open(tpath) as f:
            merged = json.load(f)
    for k, v in traffic.items():
        v["capture"] = f"profiles/{tag}_ncu_raw.csv"
        merged[k] = v
    with open(tpath, "w") as f:
        json.dump(merged, f, indent=1)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(*sys.argv[1:4])

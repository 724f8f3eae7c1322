"""Executed-instruction mix and hottest SASS lines of one kernel from an ncu report.

    python tools/ncu_source_mix.py <report.ncu-rep> [top_n]
"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    allrows = list(csv.reader(io.StringIO(raw)))
    # one section per kernel: a "Kernel Name" row, a header row, then the SASS lines
    starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
    for a, b in zip(starts[:-1], starts[1:]):
        print("\n=====", allrows[a][1][:100])
        section(allrows[a:b], top)


def section(rows, top):
    h = rows[1]
    isrc, iex, ism = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ithr = h.index("Avg. Threads Executed")
    tot, by, lines = 0, collections.Counter(), []
    for k, r in enumerate(rows[2:]):
        try:
            n = int(r[iex])
        except (ValueError, IndexError):
            continue
        toks = r[isrc].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        by[op.split(".")[0]] += n
        tot += n
        lines.append((k, n, int(r[ism] or 0), r[ithr], r[isrc]))
    print("warp instructions executed:", tot)
    for k, v in by.most_common(25):
        print(f"  {k:10s} {v:12d} {v / tot:.3f}")
    nsamp = sum(l[2] for l in lines)
    print("\nhottest lines by stall samples (line, executed, samples share, avg threads, sass)")
    for l in sorted(lines, key=lambda l: -l[2])[:top]:
        print(f"  {l[0]:5d} {l[1]:10d} {l[2] / max(nsamp, 1):.3f} {l[3]:>5s}  {l[4][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

"""Counts, per kernel of libbsgpu.so, the SASS instructions that prove the TMA / mbarrier /
256-bit-load claims of DESIGN.md (`cp.async.bulk` -> UBLKCP, mbarrier -> SYNCS, the filter gather ->
LDG.E...256) and writes profiles/<tag>_sass_grep.md.

    python tools/sass_grep.py r02
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "biostrings_b200", "lib", "libbsgpu.so")
PATTERNS = {"UBLKCP (cp.async.bulk global->shared)": r"\bUBLKCP", "SYNCS (mbarrier arrive / try_wait)": r"\bSYNCS\.",
            "LDG...256 (one 32-byte filter block per probe)": r"LDG\.E\.[A-Z0-9.]*256", "LDGSTS (cp.async)": r"\bLDGSTS",
            "ATOMG/RED (global atomics)": r"\b(ATOMG|RED)\.", "SHFL": r"\bSHFL\.", "LDS.128": r"\bLDS\.128"}


def main(tag):
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            fn = fn.replace("void ", "")
            per.setdefault(fn, collections.Counter())
            continue
        if fn is None:
            continue
        for name, pat in PATTERNS.items():
            if re.search(pat, line):
                per[fn][name] += 1
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            per[fn]["instructions"] += 1
    out = [f"# SASS instruction grep ({tag}): `cuobjdump -sass biostrings_b200/lib/libbsgpu.so`\n",
           "Built with `nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3` (biostrings_b200/build.py).\n",
           "| kernel | instructions | " + " | ".join(PATTERNS) + " |", "|---|---|" + "---|" * len(PATTERNS)]
    for fn, c in per.items():
        if not any(c[k] for k in PATTERNS) or fn.startswith("cub::"):
            continue
        out.append(f"| `{fn}` | {c['instructions']} | " + " | ".join(str(c[k]) for k in PATTERNS) + " |")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_grep.md")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")

#!/bin/bash
# rebuilds the library with different nvcc -D settings on the GPU box and times the exact scan
#   tools/exact_sweep.sh "<settings for bench_exact --only>" "<nvcc -D flags>" ...
set -u
only=$1; shift
for cfg in "$@"; do
  echo "=== $cfg"
  BSGPU_NVCC_EXTRA="$cfg" python -m biostrings_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  python tools/bench_exact.py --only "$only" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['setting'], round(d['ms_per_step'],4), {k:round(v[1],4) for k,v in d['kernel_ms'].items()}, d['matches'], d['same_counts_as_first'])"
done
python -m biostrings_b200.build --force > /dev/null 2>&1

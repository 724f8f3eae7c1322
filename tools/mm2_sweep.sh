#!/bin/bash
# rebuilds the library with different qscan2 tile settings on the GPU box and times max.mismatch 2
#   tools/mm2_sweep.sh "<nvcc -D flags>" "<nvcc -D flags>" ...
set -u
for cfg in "$@"; do
  echo "=== $cfg"
  BSGPU_NVCC_EXTRA="$cfg" python -m biostrings_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  python bench.py --mm 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/sweep_tmp.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/sweep_tmp.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['roofline'].get('kernel_ms'), d['matches_total'])"
done
python -m biostrings_b200.build --force > /dev/null 2>&1

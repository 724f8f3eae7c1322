#!/bin/bash
# One gpurun call that refreshes the round's measurements (1 GPU):
#   tools/final_measure.sh <tag>      e.g. r02b  ->  gpurun_out/<tag>_*
tag=${1:-r02b}
o=gpurun_out
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $o/${tag}_reference_arm.json 2> $o/${tag}_reference_arm.err
python tools/bench_configs.py > $o/${tag}_configs.jsonl 2> $o/${tag}_configs.err
python tools/bench_dense.py > $o/${tag}_dense.json 2> $o/${tag}_dense.err
# launch list (shares of the step), then one full capture of the dominant kernels; numbers printed under ncu are never bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $o/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'scan_xseed|qscan2|pack_subject' -c 9 -f -o $o/${tag}_full \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $o/${tag}_ncu_full.log 2>&1
# the exact scan once more with the L2 as the previous step left it (bench.py runs its steps back to back:
# the 32 MB filter is resident; ncu's default flushes the caches before every replay pass)
ncu --set full --clock-control none --cache-control none -k regex:'scan_xseed' -s 3 -c 1 -f -o $o/${tag}_xseed_warm \
    python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-e2e --no-mm2 > $o/${tag}_ncu_warm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'scan_tbdir' -c 2 -f -o $o/${tag}_tbdir \
    python tools/bench_configs.py --only C3 > $o/${tag}_ncu_tbdir.log 2>&1
tail -c 400 $o/${tag}_bench_n1.json

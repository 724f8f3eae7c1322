#!/bin/bash
# usage: tools/ncu_quick.sh <kernel regex> <tag> <bench.py args...>   (one --set full capture of the 2nd launch)
k=$1; tag=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/$tag python bench.py "$@" > gpurun_out/$tag.log 2>&1
tail -c 300 gpurun_out/$tag.log

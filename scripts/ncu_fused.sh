#!/bin/bash
# ncu capture of the fused mobilenetV3 kernels inside one eager cfg2 step (run under gpurun; plain run first)
set -x
CMD="python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-micro"
$CMD > gpurun_out/plain_fused.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mbv3 -s 180 -c 24 -o gpurun_out/prof_fused -f $CMD > gpurun_out/ncu_fused.log 2>&1
tail -3 gpurun_out/ncu_fused.log

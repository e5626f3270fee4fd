#!/bin/bash
mkdir -p gpurun_out
for v in 2 3; do
LVKB200_REMAP_KERNEL=$v timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_easu_remap -s 12 -c 1 -o gpurun_out/remap_k$v python tools/bench_remap.py --res 1080p --iters 5 > gpurun_out/ncu_k$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_remap_gpu.py tests/test_golden_gpu.py tests/test_scaling_gpu.py -q -m gpu -x 2>&1 | tail -6
for occ in 4 5 3; do echo "occ=$occ"; LVKB200_REMAP_OCC=$occ python tools/bench_remap.py --res 1080p; LVKB200_REMAP_OCC=$occ python tools/bench_remap.py --res 4k; done | tee gpurun_out/remap_flat_occ.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_easu_remap -s 12 -c 1 -o gpurun_out/remap_flat python tools/bench_remap.py --res 1080p --iters 5 > gpurun_out/ncu_flat.log 2>&1
tail -1 gpurun_out/ncu_flat.log

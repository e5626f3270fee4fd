#!/bin/bash
mkdir -p gpurun_out
python tools/dbg/remap_diff.py > gpurun_out/remap_diff.txt 2>&1
grep -c "differ" gpurun_out/remap_diff.txt; grep "differ" gpurun_out/remap_diff.txt | grep -v " 0 px"
for occ in 2 3 4; do
  echo "occ=$occ"
  LVKB200_REMAP_OCC=$occ python tools/bench_remap.py --res 1080p
  LVKB200_REMAP_OCC=$occ python tools/bench_remap.py --res 4k
done 2>&1 | tee gpurun_out/remap_occ.txt
timeout 600 python -m pytest tests/test_remap_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -5

#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/dbg/remap_diff.py > gpurun_out/remap_diff.txt 2>&1
grep -c "differ" gpurun_out/remap_diff.txt; grep "differ" gpurun_out/remap_diff.txt | grep -v " 0 px"; tail -3 gpurun_out/remap_diff.txt
timeout 300 python -m pytest tests/test_remap_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -5
for occ in 2 3 4; do
  echo "v3 occ=$occ"
  LVKB200_REMAP_OCC=$occ timeout 120 python tools/bench_remap.py --res 1080p
  LVKB200_REMAP_OCC=$occ timeout 120 python tools/bench_remap.py --res 4k
done 2>&1 | tee gpurun_out/remap_occ_v3.txt
for v in 2 3; do
  echo "bench kernel v$v"
  LVKB200_REMAP_KERNEL=$v timeout 300 python bench.py 2>/dev/null | tee gpurun_out/bench_v$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['avg_kernel_us'], d['stage_us'])"
done

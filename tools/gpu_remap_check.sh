#!/bin/bash
# GPU recipe: parity of both EASU builds, stand-alone kernel timing, one ncu --set full capture of the default build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_remap_gpu.py -q 2>&1 | tail -8
for occ in 5 6 7; do for r in 1080p 4k; do echo "occ=$occ"; LVKB200_REMAP_OCC=$occ python tools/bench_remap.py --res $r; done; done 2>&1 | tee gpurun_out/remap_bench.txt
ncu --set full --clock-control none --import-source on -k regex:k_easu_remap_fast -s 12 -c 1 -f -o gpurun_out/remap_fast_1080p \
    python tools/bench_remap.py --res 1080p --iters 4 > gpurun_out/ncu_remap.log 2>&1
ncu -i gpurun_out/remap_fast_1080p.ncu-rep --page raw --csv > gpurun_out/remap_fast_1080p_raw.csv 2>/dev/null
ncu -i gpurun_out/remap_fast_1080p.ncu-rep --page source --csv > gpurun_out/remap_fast_1080p_source.csv 2>/dev/null
tail -2 gpurun_out/ncu_remap.log

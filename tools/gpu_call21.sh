#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_deblock_gpu.py -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_rows.py --res 4k 2>&1 | head -1
timeout 300 python tools/bench_rows.py --res 1080p 2>&1 | head -1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_deblock -c 3 --csv --log-file gpurun_out/deblock_kernels2.csv python tools/profile_deblock.py --resolution 4k > gpurun_out/deblock_prof.log 2>&1
grep -E "gpu__time_duration|inst_executed" gpurun_out/deblock_kernels2.csv | cut -d, -f5,13-15 | head -9

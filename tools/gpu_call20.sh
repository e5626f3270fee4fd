#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_rows.py --res 4k 2>&1 | tail -6 | tee gpurun_out/rows_bench.txt
timeout 300 python tools/bench_rows.py --res 1080p 2>&1 | tail -6 | tee -a gpurun_out/rows_bench.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_deblock -c 9 --csv --log-file gpurun_out/deblock_kernels.csv python tools/profile_deblock.py --resolution 4k > gpurun_out/deblock_prof.log 2>&1
tail -2 gpurun_out/deblock_prof.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mesh_cgls -s 8 -c 1 -o gpurun_out/mesh_cgls python bench.py --preset F --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/ncu_mesh.log 2>&1
tail -1 gpurun_out/ncu_mesh.log | cut -c1-200

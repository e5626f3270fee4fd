#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lk_track -s 20 -c 1 -o gpurun_out/lk_r1 python tools/profile_stream.py --frames 40 > gpurun_out/ncu_lk.log 2>&1
tail -1 gpurun_out/ncu_lk.log | cut -c1-150
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ransac_refine -s 20 -c 1 -o gpurun_out/refine_r1 python tools/profile_stream.py --frames 40 > gpurun_out/ncu_refine.log 2>&1
tail -1 gpurun_out/ncu_refine.log | cut -c1-150

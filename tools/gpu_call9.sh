#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scaling_gpu.py tests/test_compat_gpu.py -q -m gpu 2>&1 | tail -15
timeout 300 python tools/bench_scaling.py > gpurun_out/scaling_bench2.txt 2>gpurun_out/scaling_bench.err; cat gpurun_out/scaling_bench2.txt; tail -3 gpurun_out/scaling_bench.err
timeout 300 python tools/bench_scaling.py --src 720p --dst 1080p >> gpurun_out/scaling_bench2.txt 2>>gpurun_out/scaling_bench.err; tail -3 gpurun_out/scaling_bench2.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_rcas -s 12 -c 1 -o gpurun_out/rcas_r2 python tools/bench_scaling.py --iters 5 > gpurun_out/ncu_rcas.log 2>&1
tail -2 gpurun_out/ncu_rcas.log

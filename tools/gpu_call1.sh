#!/bin/bash
# One gpurun call: FFMA2 probe, remap parity + micro-bench, full GPU test suite, bench line, ncu capture of the remap.
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe tools/dbg/ffma2_probe.cu && /tmp/probe > gpurun_out/ffma2_probe.txt 2>&1
timeout 300 python -m pytest tests/test_remap_gpu.py -x -q -m gpu > gpurun_out/t_remap.log 2>&1
tail -3 gpurun_out/t_remap.log
python tools/bench_remap.py --res 1080p > gpurun_out/remap_1080p.json 2>gpurun_out/remap_1080p.err
python tools/bench_remap.py --res 4k > gpurun_out/remap_4k.json 2>gpurun_out/remap_4k.err
cat gpurun_out/remap_1080p.json gpurun_out/remap_4k.json
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1
tail -5 gpurun_out/t_all.log
timeout 300 python bench.py > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
cat gpurun_out/bench_1080p.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_easu_remap -s 12 -c 1 -o gpurun_out/remap_v2 python tools/bench_remap.py --res 1080p --iters 5 > gpurun_out/ncu_remap.log 2>&1
ls -la gpurun_out

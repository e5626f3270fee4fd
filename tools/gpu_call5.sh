#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_formats_gpu.py -x -q -m gpu 2>&1 | tail -15
python tools/bench_remap.py --res 1080p; python tools/bench_remap.py --res 4k
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8

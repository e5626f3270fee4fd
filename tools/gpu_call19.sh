#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k "concurrent or lookahead" 2>&1 | tail -4
timeout 600 python tools/bench_multistream.py | tee gpurun_out/multistream_1080p.txt
timeout 600 python tools/bench_multistream.py --resolution 4k --streams 1 2 4 --frames 150 | tee gpurun_out/multistream_4k.txt

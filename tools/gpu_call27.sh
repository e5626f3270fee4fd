#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tracking_gpu.py tests/test_golden_gpu.py tests/test_pipeline_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_lk.err > gpurun_out/bench_lk.json; python -c "import json; d=json.load(open('gpurun_out/bench_lk.json')); print('lkhoist', d['value'], d['e2e']['value'], {k: round(v,1) for k,v in d['stage_us'].items()}, d['parity_failures'])"

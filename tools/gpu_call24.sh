#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/t_final2.log
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_final2.err > gpurun_out/bench_final2.json; python -c "import json; d=json.load(open('gpurun_out/bench_final2.json')); print('final2', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['parity_failures'])"

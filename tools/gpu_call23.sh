#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tracking_gpu.py tests/test_pipeline_gpu.py -q -m gpu -s -k "local_motions or F-1080p or F-False" 2>&1 | grep -E "^\[mesh|passed|failed|Error|assert" | head
timeout 300 python bench.py --preset F --steps 150 --warmup 30 --no-cpu-baseline 2>gpurun_out/bench_F3.err | tee gpurun_out/bench_F3.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('F', round(d['value'],1), round(d['e2e']['value'],1), {k: round(v,1) for k,v in d['stage_us'].items()}, d.get('parity_failures'))"
tail -2 gpurun_out/bench_F3.err

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tracking_gpu.py -x -q -m gpu -s -k "local_motions" 2>&1 | grep -E "^\[|passed|failed|Error|error|assert" | head -30
timeout 600 python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu -s -k "free_running" 2>&1 | grep -E "^\[|passed|failed|Error|error|assert" | head -20
timeout 300 python bench.py --preset F --steps 120 --warmup 20 --no-cpu-baseline 2>gpurun_out/bench_F2.err | tee gpurun_out/bench_F2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('F', d['value'], d['e2e']['value'], d['stage_us'], d.get('parity_failures'))"
tail -3 gpurun_out/bench_F2.err

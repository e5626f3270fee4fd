#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu -s -k "free_running" 2>&1 | grep -E "^\[|passed|failed|Error|assert" | head -20
timeout 300 python bench.py --preset F --steps 120 --warmup 20 --no-cpu-baseline 2>gpurun_out/bench_F.err | tee gpurun_out/bench_F.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('F', d['value'], d['e2e']['value'], d['stage_us'])"
tail -3 gpurun_out/bench_F.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_easu_remap -s 12 -c 1 -o gpurun_out/remap_final python tools/bench_remap.py --res 1080p --iters 5 > gpurun_out/ncu_final.log 2>&1

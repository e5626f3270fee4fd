#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_compat_gpu.py tests/test_deblock_gpu.py -q -m gpu -x 2>&1 | tail -12
run() { name=$1; shift; timeout 600 python bench.py "$@" 2>gpurun_out/bench_$name.err | tee gpurun_out/bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e'].get('apply_fps_rank0',0),1), round(d['roofline']['avg_kernel_us'],1), d.get('parity_failures'), d['gpu_launches'])"; tail -2 gpurun_out/bench_$name.err; }
run la --no-cpu-baseline
run nola --no-cpu-baseline --no-lookahead
run la4k --no-cpu-baseline --resolution 4k --steps 200
run laD --no-cpu-baseline --preset D
run lachain --no-cpu-baseline --deblock

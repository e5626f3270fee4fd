#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_remap_gpu.py tests/test_golden_gpu.py tests/test_scaling_gpu.py -q -m gpu -x 2>&1 | tail -4
for occ in 4 5; do echo "occ=$occ"; LVKB200_REMAP_OCC=$occ python tools/bench_remap.py --res 1080p; LVKB200_REMAP_OCC=$occ python tools/bench_remap.py --res 4k; done | tee gpurun_out/remap_flat2_occ.txt
run() { name=$1; shift; timeout 600 python bench.py "$@" 2>gpurun_out/bench_$name.err | tee gpurun_out/bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['avg_kernel_us'],1), d.get('parity_failures'))"; tail -2 gpurun_out/bench_$name.err; }
LVKB200_REMAP_OCC=4 run o4 --no-cpu-baseline
LVKB200_REMAP_OCC=5 run o5 --no-cpu-baseline
LVKB200_REMAP_OCC=4 run o4_4k --no-cpu-baseline --resolution 4k --steps 200
LVKB200_REMAP_OCC=5 run o5_4k --no-cpu-baseline --resolution 4k --steps 200

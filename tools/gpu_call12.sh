#!/bin/bash
# compute-sanitizer memcheck over a small-size subset of the GPU tests (out-of-bounds / misaligned accesses in any kernel)
mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5"
run() { name=$1; shift; timeout 900 $CS python -m pytest "$@" -q -m gpu -x -p no:cacheprovider > gpurun_out/memcheck_$name.log 2>&1; echo "memcheck $name rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/memcheck_$name.log | tr '\n' ' ')"; }
run scaling tests/test_scaling_gpu.py -k "every_ring or unaligned or same_size or (upscale_bit_exact and (333 or 37 or 8-8)) or (sharpen_bit_exact and 0.8 and (963 or 129 or 127 or 5-4 or 3-3 or 2-7 or 1-1))"
run remap tests/test_remap_gpu.py
run formats tests/test_formats_gpu.py -k "64-36 or 482"
run deblock tests/test_deblock_gpu.py -k "170 or 963 or settings"
run tracking tests/test_tracking_gpu.py
run pipeline tests/test_pipeline_gpu.py -k "D-720p or pipelined or pure_delay"
run pipelineF tests/test_pipeline_gpu.py -k "F-1080p"

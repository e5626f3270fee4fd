#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_compat_gpu.py tests/test_pipeline_gpu.py -q -m gpu -k "compat or pipelined" 2>&1 | tail -8
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5"
run() { name=$1; shift; timeout 900 $CS python -m pytest "$@" -q -m gpu -x -p no:cacheprovider > gpurun_out/memcheck_$name.log 2>&1; echo "memcheck $name rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/memcheck_$name.log | tr '\n' ' ')"; }
run formats tests/test_formats_gpu.py -k "size0 or size2 or unsupported"
run scaling2 tests/test_scaling_gpu.py -k "src_wh0 or src_wh3 or src_wh5 or src_wh6 or (0.8 and (wh1 or wh2 or wh3 or wh4 or wh5 or wh6 or wh7 or wh8)) or scaling_filter_vs_oracle and src_size1"
run golden tests/test_golden_gpu.py -k "not full_size"

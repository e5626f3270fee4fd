#!/bin/bash
# GPU recipe: what the driver runs at round end - the whole GPU test suite, smoke(), the default bench line, the
# 20-step form and the reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.err | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_k20.json 2>/dev/null
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_k20.json 2>/dev/null
timeout 100 python tools/bench_rows.py --res 4k 2>&1 | head -1 | cut -c1-160

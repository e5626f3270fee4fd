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
timeout 200 python bench.py --preset F --steps 100 --warmup 20 --no-extra-configs --no-cpu-baseline > gpurun_out/bench_F.json 2>/dev/null
timeout 200 python bench.py --preset D --steps 100 --warmup 20 --no-extra-configs --no-cpu-baseline > gpurun_out/bench_D.json 2>/dev/null
python - <<'PY'
import json
for n in ("bench_default", "bench_k20", "bench_reference_k20", "bench_F", "bench_D"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "launches", j.get("gpu_launches"))
    except Exception as e:
        print(n, "unreadable", e)
PY

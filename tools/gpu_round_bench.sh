#!/bin/bash
# GPU recipe: the round's bench records (default line with the configs block, the driver's 20-step form, the other
# presets / the chained configuration, one-GPU multi-stream capacity).  Everything lands under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --steps 20 --warmup 5 --no-extra-configs > gpurun_out/bench_k20.json 2>/dev/null
for p in D F; do python bench.py --preset $p --steps 100 --no-extra-configs --no-cpu-baseline > gpurun_out/bench_preset_$p.json 2>/dev/null; done
python bench.py --deblock --steps 100 --no-extra-configs --no-cpu-baseline > gpurun_out/bench_chain_1080p.json 2>/dev/null
python bench.py --deblock --resolution 4k --steps 60 --no-extra-configs --no-cpu-baseline > gpurun_out/bench_chain_4k.json 2>/dev/null
python tools/bench_multistream.py --streams 1 2 4 > gpurun_out/multistream_1080p.txt 2>&1
python tools/bench_multistream.py --streams 1 4 --resolution 4k --frames 150 > gpurun_out/multistream_4k.txt 2>&1
tail -3 gpurun_out/multistream_1080p.txt

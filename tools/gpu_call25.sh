#!/bin/bash
mkdir -p gpurun_out
for k in 1 2; do timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_fr.err > gpurun_out/bench_fr$k.json; python -c "import json; d=json.load(open('gpurun_out/bench_fr$k.json')); print('frameref', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['samples'], d['parity_failures'])"; done
timeout 600 python bench.py --no-cpu-baseline --resolution 4k --steps 200 2>gpurun_out/bench_fr.err > gpurun_out/bench_fr4k.json; python -c "import json; d=json.load(open('gpurun_out/bench_fr4k.json')); print('frameref4k', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_failures'])"

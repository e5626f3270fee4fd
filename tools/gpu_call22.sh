#!/bin/bash
# round-1 final evidence run: full GPU suite, smoke(), the default bench line, the reference arm, the ncu launch list of
# the same bench command and one full capture of the remap kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/t_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/bench_final.err > gpurun_out/bench_final.json; python -c "import json; d=json.load(open('gpurun_out/bench_final.json')); print('final', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_kernel_us'], d['cpu_baseline']['value'], d['clocks'], d['parity_failures'], d['gpu_launches'])"
tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_reference_final.json; cut -c1-220 gpurun_out/bench_reference_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 40 --warmup 12 --no-cpu-baseline > gpurun_out/launches_final.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_final.csv | tee gpurun_out/launches_final.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_easu_remap -s 12 -c 1 -o gpurun_out/remap_r1final python tools/bench_remap.py --res 1080p --iters 5 > gpurun_out/ncu_r1final.log 2>&1
tail -1 gpurun_out/ncu_r1final.log

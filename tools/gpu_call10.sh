#!/bin/bash
# full GPU suite + the default bench line + the ncu launch list of the same bench command (profiles/)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/t_all2.log
timeout 600 python bench.py 2>gpurun_out/bench_r1f.err | tee gpurun_out/bench_r1f.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1080p', d['value'], d['e2e']['value'], d['roofline'], d['stage_us'], d.get('parity_failures'), d['cpu_baseline'])"
tail -3 gpurun_out/bench_r1f.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 40 --warmup 12 --no-cpu-baseline > gpurun_out/launches_r1f.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r1f.csv | tee gpurun_out/launches_r1f.txt

#!/bin/bash
# the other BASELINE configs on one GPU: 4K (configs[2]), presets D and F, the Deblocking -> Stabilization chain (config 5)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" 2>gpurun_out/bench_$name.err | tee gpurun_out/bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e'].get('apply_fps_rank0',0),1), round(d['roofline']['avg_kernel_us'],1), round(d['roofline']['frac'],4), {k: round(v,1) for k,v in d['stage_us'].items()}, d.get('parity_failures'), d.get('clocks',{}).get('sm_mhz'), d.get('cpu_baseline',{}).get('value'))"; tail -2 gpurun_out/bench_$name.err; }
run 4k --resolution 4k --steps 200 --warmup 30
run D --preset D --no-cpu-baseline
run F --preset F --steps 150 --warmup 30 --no-cpu-baseline
run chain4k --resolution 4k --deblock --steps 150 --warmup 30 --no-cpu-baseline
run chain1080 --deblock --no-cpu-baseline
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_reference.json
python tools/bench_remap.py --res 1080p; python tools/bench_remap.py --res 4k

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 30 2>gpurun_out/bench_n2.err | tee gpurun_out/bench_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N2', round(d['value'],1), round(d['e2e']['value'],1), d['n_gpus'], d.get('parity_failures'), d['roofline']['avg_kernel_us'])"
tail -3 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 | tail -2 | cut -c1-300

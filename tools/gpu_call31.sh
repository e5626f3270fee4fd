#!/bin/bash
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
timeout 120 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -x -k "pipelined or lookahead_equals_plain_submit and H-False or pure_delay or configure" 2>&1 | tail -2
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('small', round(d['value']), round(d['e2e']['value']), d['outputs'], d['parity_failures'])"

#!/bin/bash
timeout 300 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('small', d['steps'], d['warmup'], round(d['value']), round(d['e2e']['value']), d['outputs'], d['parity_failures'], d['cpu_baseline']['value'])"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ref', d['steps'], d['warmup'], d['value'])"

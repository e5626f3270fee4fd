#!/bin/bash
# GPU recipe: where the held-back remap is launched relative to the tracking chain (LVKB200_REMAP_OVERLAP = 0 behind the
# chain, 1 beside the whole chain, 2 beside everything after LK), 1080p and 4K.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for ov in 2 0 1; do for res in 1080p 4k; do
  LVKB200_REMAP_OVERLAP=$ov timeout 200 python bench.py --resolution $res --steps 100 --windows 2 --no-extra-configs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('overlap=$ov $res value', round(d['value']), 'e2e', round(d['e2e']['value']), 'nv12', d['e2e_nv12'] and round(d['e2e_nv12']['value']), 'remap_us', round(d['roofline']['avg_kernel_us'],1), 'parity', d['parity_failures'])"
done; done 2>&1 | tee gpurun_out/overlap_ab.txt

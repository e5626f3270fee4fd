#!/bin/bash
# GPU recipe: per-launch kernel durations of the default bench command (ncu --metrics gpu__time_duration.sum, the
# B200_PROFILING.md launch-list pass: cold-cache, serialised - the kernels' SHARES are what matters).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 200 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 40 --warmup 12 --windows 1 --no-extra-configs --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches.csv') if not l.startswith('==')))
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
t = collections.defaultdict(list)
for r in rows[1:]:
    try: t[r[ki].split('(')[0].split('<')[0].replace('void lvkb200::','').replace('<unnamed>::','')].append(float(r[vi].replace(',','')))
    except Exception: pass
tot = sum(sum(v) for v in t.values())
out = ['kernel                              launches   avg_us   share']
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    out.append(f'{k[:34]:34s} {len(v):8d} {sum(v)/len(v)/1000:8.2f} {100*sum(v)/tot:6.1f} %')
open('gpurun_out/launches.txt','w').write('\n'.join(out)+'\n'); print('\n'.join(out))
PY

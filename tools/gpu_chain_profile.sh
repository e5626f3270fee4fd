#!/bin/bash
# GPU recipe: ncu --set full captures of the tracking chain's two longest kernels inside the default bench workload.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for k in k_ransac_refine k_lk_track; do
  ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:$k -s 40 -c 1 -f -o gpurun_out/$k \
      python bench.py --steps 60 --warmup 12 --windows 1 --no-extra-configs --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/$k.ncu-rep --page raw --csv > gpurun_out/${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$k.ncu-rep --page source --csv > gpurun_out/${k}_source.csv 2>/dev/null
  tail -1 gpurun_out/ncu_$k.log | cut -c1-200
done

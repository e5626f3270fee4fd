cd "${GRAFT_REPO_ROOT:-/root/repo}"
ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on --graph-profiling node -k regex:k_ransac_refine -s 40 -c 1 -f -o gpurun_out/k_ransac_refine2 \
      python bench.py --steps 60 --warmup 12 --windows 1 --no-extra-configs --no-cpu-baseline > gpurun_out/ncu_refine2.log 2>&1
ncu -i gpurun_out/k_ransac_refine2.ncu-rep --page source --csv > gpurun_out/k_ransac_refine2_source.csv 2>/dev/null
ncu -i gpurun_out/k_ransac_refine2.ncu-rep --page raw --csv > gpurun_out/k_ransac_refine2_raw.csv 2>/dev/null
tail -1 gpurun_out/ncu_refine2.log | cut -c1-100

# GPU recipe: the motion-mesh solver (preset F, 16x16 mesh): parity tests, the preset's bench line, and a source-level
# ncu capture of one k_mesh_cgls2 launch.  LVKB200_MESH_V1=1 selects the round-1 kernel for an A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tracking_gpu.py -m gpu -q -k "local_motions" -s 2>&1 | grep -E "mesh|passed|failed|Error|assert" | head -20
timeout 500 python -m pytest -m gpu -q "tests/test_pipeline_gpu.py::test_free_running_vs_oracle[F-1080p-30]" "tests/test_pipeline_gpu.py::test_lookahead_equals_plain_submit[F-False]" \
    "tests/test_parity_e2e_gpu.py::test_free_running_final_pixels[F-1080p]" 2>&1 | tail -5
timeout 200 python bench.py --preset F --steps 100 --warmup 20 --no-extra-configs --no-cpu-baseline > gpurun_out/bench_F.json 2>/dev/null
python - <<'PY'
import json
j = json.loads(open('gpurun_out/bench_F.json').read().strip().splitlines()[-1])
print('preset F value', round(j['value']), 'ms/step', j['ms_per_step'], 'e2e', round(j['e2e']['value']))
PY
timeout 300 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on --graph-profiling node -k regex:k_mesh_cgls -s 40 -c 1 -f -o gpurun_out/k_mesh_cgls \
      python bench.py --preset F --steps 60 --warmup 12 --windows 1 --no-extra-configs --no-cpu-baseline > gpurun_out/ncu_mesh.log 2>&1
ncu -i gpurun_out/k_mesh_cgls.ncu-rep --page source --csv > gpurun_out/k_mesh_cgls_source.csv 2>/dev/null
ncu -i gpurun_out/k_mesh_cgls.ncu-rep --page raw --csv > gpurun_out/k_mesh_cgls_raw.csv 2>/dev/null
tail -1 gpurun_out/ncu_mesh.log | cut -c1-100

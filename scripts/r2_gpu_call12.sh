set -x
cd "$GRAFT_REPO_ROOT"
SWEEP_SHORT=1 timeout 500 python scripts/build_reinsert_sweep.py kitchen sanmiguel demoscene > gpurun_out/r2_build_reinsert_other_scenes.log 2>&1
cat gpurun_out/r2_build_reinsert_other_scenes.log
timeout 300 python -m pytest tests/test_gpu_build.py tests/test_gpu_fullsize.py -x -q -k "build or device_built" 2>&1 | tail -5

set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_build.py -x -q 2>&1 | tail -15
timeout 900 python scripts/build_reinsert_sweep.py hairball kitchen sanmiguel > gpurun_out/r2_build_reinsert_sweep.log 2>&1
cat gpurun_out/r2_build_reinsert_sweep.log

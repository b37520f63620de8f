set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t37_pytest.log; cat gpurun_out/r2_t37_pytest.log
for scene in hairball kitchen; do
  for lib in libtray_cuda.so libtray_cuda_mhi.so libtray_cuda.so libtray_cuda_mhi.so; do
    TRAY_CUDA_LIB=$PWD/tray_racing_b200/$lib timeout 300 python scripts/r2_perf.py $scene --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight"
  done
done 2>&1 | tee gpurun_out/r2_pipe_balance_ab3.log
run() { env "$@" timeout 300 python scripts/r2_perf.py hairball --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight" | sed "s/^/$* /"; }
{
run TRAY_CUDA_TRI_WEIGHT=4
run TRAY_CUDA_TRI_WEIGHT=2
run TRAY_CUDA_REFILL_MIN=6
} 2>&1 | tee -a gpurun_out/r2_pipe_balance_ab3.log
TRAY_CUDA_LIB=$PWD/tray_racing_b200/libtray_cuda.so timeout 300 python scripts/r2_perf.py sanmiguel --w 3840 --h 2160 --frames 20 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight" | tee -a gpurun_out/r2_pipe_balance_ab3.log

set -x
cd "$GRAFT_REPO_ROOT"
for scene in hairball kitchen; do
  for lib in libtray_cuda.so libtray_cuda_imad.so libtray_cuda_i2fy1.so libtray_cuda_i2fy2.so libtray_cuda.so libtray_cuda_imad.so libtray_cuda_i2fy1.so; do
    TRAY_CUDA_LIB=$PWD/tray_racing_b200/$lib timeout 300 python scripts/r2_perf.py $scene --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight"
  done
done 2>&1 | tee gpurun_out/r2_pipe_balance_ab.log

set -x
cd "$GRAFT_REPO_ROOT"
for scene in hairball kitchen; do
  for lib in libtray_cuda.so libtray_cuda_pl.so libtray_cuda_ptxpl.so libtray_cuda.so libtray_cuda_pl.so libtray_cuda_ptxpl.so; do
    TRAY_CUDA_LIB=$PWD/tray_racing_b200/$lib timeout 300 python scripts/r2_perf.py $scene --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight"
  done
done 2>&1 | tee gpurun_out/r2_push_late_ab.log

set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t27_pytest.log; cat gpurun_out/r2_t27_pytest.log
for scene in hairball kitchen; do
  for lib in libtray_cuda_nobatch.so libtray_cuda.so libtray_cuda_nobatch.so libtray_cuda.so; do
    TRAY_CUDA_LIB=$PWD/tray_racing_b200/$lib timeout 300 python scripts/r2_perf.py $scene --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight"
  done
done 2>&1 | tee gpurun_out/r2_batch_refill_ab.log
for rm in 2 3; do
  TRAY_CUDA_REFILL_MIN=$rm TRAY_CUDA_LIB=$PWD/tray_racing_b200/libtray_cuda.so timeout 300 python scripts/r2_perf.py hairball --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight" | sed "s/^/refill_min=$rm /"
done 2>&1 | tee -a gpurun_out/r2_batch_refill_ab.log
for lib in libtray_cuda_nobatch.so libtray_cuda.so; do
  TRAY_CUDA_LIB=$PWD/tray_racing_b200/$lib timeout 300 python scripts/r2_perf.py sanmiguel --w 3840 --h 2160 --frames 20 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight"
done 2>&1 | tee -a gpurun_out/r2_batch_refill_ab.log

set -x
cd "$GRAFT_REPO_ROOT"
for lib in libtray_cuda.so libtray_cuda_tp.so libtray_cuda.so libtray_cuda_tp.so; do
  TRAY_CUDA_LIB=$PWD/tray_racing_b200/$lib timeout 120 python scripts/r2_perf.py hairball --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight"
done 2>&1 | tee gpurun_out/r2_tri_packed_ab.log
TRAY_CUDA_LIB=$PWD/tray_racing_b200/libtray_cuda_tp.so timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_anyhit_f16.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/r2_tri_packed_ab.log

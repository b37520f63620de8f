set -x
cd "$GRAFT_REPO_ROOT"
for scene in hairball; do
  for lib in libtray_cuda_r2base.so libtray_cuda.so libtray_cuda_ld128.so libtray_cuda_r2base.so libtray_cuda.so libtray_cuda_ld128.so; do
    TRAY_CUDA_LIB=$PWD/tray_racing_b200/$lib timeout 300 python scripts/r2_perf.py $scene --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight"
  done
done 2>&1 | tee gpurun_out/r2_tnode_ab2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 2 -o gpurun_out/r2_tnode_hairball python scripts/render_frames.py --scene hairball --frames 3 > gpurun_out/r2_ncu_tnode.log 2>&1
ls -la gpurun_out/

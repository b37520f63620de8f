set -x
cd "$GRAFT_REPO_ROOT"
for scene in hairball kitchen sanmiguel; do
  for tw in 3 4 3 4; do
    TRAY_CUDA_TRI_WEIGHT=$tw timeout 300 python scripts/r2_perf.py $scene --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x1 in flight|x2 in flight" | sed "s/^/tw=$tw /"
  done
done 2>&1 | tee gpurun_out/r2_triweight_ab.log

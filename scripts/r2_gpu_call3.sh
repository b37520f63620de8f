set -x
cd "$GRAFT_REPO_ROOT"
L=tray_racing_b200
python -m pytest tests/test_gpu_relaxed.py -x -q 2>&1 | tail -15 > gpurun_out/r2_t3_pytest.log
for lib in libtray_cuda.so libtray_cuda_tri2.so libtray_cuda_node2.so libtray_cuda_tri2node2.so; do
  TRAY_CUDA_LIB=$PWD/$L/$lib python scripts/r2_perf.py hairball
  TRAY_CUDA_LIB=$PWD/$L/$lib python scripts/r2_perf.py kitchen
  TRAY_CUDA_LIB=$PWD/$L/$lib python scripts/r2_perf.py sanmiguel --w 3840 --h 2160 --shards 8 --device-build
done > gpurun_out/r2_perf_ab2.log 2>&1
python scripts/relaxed_census.py --out gpurun_out/r2_relaxed_census.json > gpurun_out/r2_relaxed.log 2>&1
for pm in 6 8 16 24; do for tw in 2 4 8; do echo "post_max $pm tri_weight $tw"; TRAY_CUDA_POST_MAX=$pm TRAY_CUDA_RELAXED_TRI_WEIGHT=$tw python scripts/relaxed_census.py --configs c3 --out gpurun_out/tmp.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('c3 '):
        d = json.loads(l[3:]); print(d['exact']['mrays_s'], d['relaxed']['mrays_s'], d['speedup'], d['relaxed']['nodes_per_ray'], d['score_primary']['pass'], d['score_bounce']['pass'])
"; done; done > gpurun_out/r2_relaxed_sweep.log 2>&1
cat gpurun_out/r2_t3_pytest.log
grep -v "^+" gpurun_out/r2_perf_ab2.log
cat gpurun_out/r2_relaxed.log
cat gpurun_out/r2_relaxed_sweep.log

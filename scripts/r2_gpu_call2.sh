set -x
cd "$GRAFT_REPO_ROOT"
L=tray_racing_b200
for lib in libtray_cuda.so libtray_cuda_tri2.so libtray_cuda_w1.so libtray_cuda_w2.so; do
  TRAY_CUDA_LIB=$PWD/$L/$lib python scripts/r2_perf.py hairball
  TRAY_CUDA_LIB=$PWD/$L/$lib python scripts/r2_perf.py kitchen
  TRAY_CUDA_LIB=$PWD/$L/$lib python scripts/r2_perf.py sanmiguel --w 3840 --h 2160 --shards 8 --device-build
done > gpurun_out/r2_perf_ab1.log 2>&1
python scripts/r2_perf.py sanmiguel --w 3840 --h 2160 --device-build >> gpurun_out/r2_perf_ab1.log 2>&1
python scripts/variant_census.py --out gpurun_out/r2_variant_census.json > gpurun_out/r2_census.log 2>&1
python bench.py --steps 100 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -5 gpurun_out/r2_bench_a.err
grep -v "^+" gpurun_out/r2_perf_ab1.log
tail -8 gpurun_out/r2_census.log
head -c 3000 gpurun_out/r2_bench_a.json

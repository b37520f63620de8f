set -x
cd "$GRAFT_REPO_ROOT"
run() { env "$@" timeout 300 python scripts/r2_perf.py hairball --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight" | sed "s/^/$* /"; }
{
run TRAY_CUDA_GEN_MIN=4
run TRAY_CUDA_GEN_MIN=2
run TRAY_CUDA_GEN_MIN=8
run TRAY_CUDA_GEN_MIN=16
run TRAY_CUDA_BOUNCE_SORT=2
run TRAY_CUDA_BOUNCE_SORT=0
run TRAY_CUDA_REFILL_MIN=3
run TRAY_CUDA_REFILL_MIN=6
run TRAY_CUDA_GEN_MIN=4
} 2>&1 | tee gpurun_out/r2_knobs_tw3.log

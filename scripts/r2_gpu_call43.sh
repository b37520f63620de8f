set -x
cd "$GRAFT_REPO_ROOT"
{
timeout 60 python scripts/r2_perf.py kitchen --frames 30 2>&1 | grep -E "primary|x1 in flight|x2 in flight"
timeout 60 python scripts/r2_perf.py demoscene --frames 30 2>&1 | grep -E "primary|x1 in flight|x2 in flight"
timeout 90 python scripts/r2_perf.py sanmiguel --w 3840 --h 2160 --frames 12 2>&1 | grep -E "primary|x1 in flight|x2 in flight"
} | tee gpurun_out/r2_final_kernel_other_configs.log

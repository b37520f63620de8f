set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t39_pytest.log; cat gpurun_out/r2_t39_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
run() { scene=$1; shift; env "$@" timeout 300 python scripts/r2_perf.py $scene --frames 40 2>&1 | grep -v "^+" | grep -E "primary|x2 in flight" | sed "s/^/$* /"; }
{
run hairball TRAY_CUDA_REFILL_MIN=4
run hairball TRAY_CUDA_REFILL_MIN=6
run hairball TRAY_CUDA_REFILL_MIN=8
run kitchen TRAY_CUDA_REFILL_MIN=4
run kitchen TRAY_CUDA_REFILL_MIN=6
run kitchen TRAY_CUDA_REFILL_MIN=8
run sanmiguel TRAY_CUDA_REFILL_MIN=4
run sanmiguel TRAY_CUDA_REFILL_MIN=6
} 2>&1 | tee gpurun_out/r2_refill_min_ab.log

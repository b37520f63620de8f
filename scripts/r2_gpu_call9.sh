set -x
cd "$GRAFT_REPO_ROOT"
timeout 90 python - <<'PY'
import os, sys, time
sys.path.insert(0, '.')
from tray_racing_b200 import cuda, host
import numpy as np
m = host.Mesh.generate("hairball", 3, 0.05)
for passes in (0, 1, 2, 4):
    os.environ["TRAY_CUDA_BUILD_REINSERT"] = str(passes)
    g = cuda.TrayCudaScene.build(m.tris())
    print("small", passes, g.build_stats, flush=True)
    g.close()
PY
echo "small rc=$?"
timeout 240 python -m pytest tests/test_gpu_build.py -x -q 2>&1 | tail -15
echo "pytest rc=$?"
timeout 400 python scripts/build_reinsert_sweep.py hairball > gpurun_out/r2_build_reinsert_sweep.log 2>&1
echo "sweep rc=$?"
cat gpurun_out/r2_build_reinsert_sweep.log

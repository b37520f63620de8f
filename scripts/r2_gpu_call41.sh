set -x
cd "$GRAFT_REPO_ROOT"
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from tray_racing_b200 import cuda, host
g = np.load("tests/golden/cornell_box.npz")
cam = host.Camera(tuple(g["eye"]), tuple(g["look_at"]), float(g["fov"]))
mesh = host.Mesh.from_tris(g["tris"], g["offsets"], cam)
for tlas in (False, True):
    p = host.PackedScene(mesh, use_tlas=tlas)
    view = host.view_from_camera(cam, 160, 96, p.tlas_start)
    sc = cuda.TrayCudaScene.from_packed(p)
    sc.render(view, 160, 96, 0, cuda.RENDER_BOUNCE | cuda.RENDER_RGBA)
    sc.render(view, 160, 96, 0, cuda.RENDER_BOUNCE | cuda.RENDER_RGBA | cuda.RENDER_OVERLAP)
    sc.sync()
    sc.close()
print("sanitizer script done")
PY
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py 2>&1 | tail -8 | tee gpurun_out/r2_sanitizer_memcheck.log

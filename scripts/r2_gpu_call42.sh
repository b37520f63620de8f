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
scenes = [(mesh, False, 48), (mesh, True, 64), (mesh, False, 24)]
if len(sys.argv) > 1:
    m = host.Mesh.generate("hairball", 3, 0.02)
    scenes.append((m, False, 48))
for m, tlas, stride in scenes:
    p = host.PackedScene(m, use_tlas=tlas, tri_stride=stride)
    view = host.view_from_camera(m.camera, 160, 96, p.tlas_start)
    sc = cuda.TrayCudaScene.from_packed(p)
    sc.set_frames_in_flight(2)
    for _ in range(2):
        sc.render(view, 160, 96, 0, cuda.RENDER_BOUNCE | cuda.RENDER_RGBA)
        sc.render(view, 160, 96, 0, cuda.RENDER_BOUNCE | cuda.RENDER_RGBA | cuda.RENDER_OVERLAP)
    sc.sync()
    rays = np.zeros(5000, dtype=host.RAY_DTYPE); rng = np.random.default_rng(1)
    rays["o"] = rng.uniform(-1, 1, (5000, 3)); rays["d"] = rng.normal(size=(5000, 3)); rays["tmax"] = 1e30
    sc.traverse(rays); sc.traverse(rays, any_hit=True)
    sc.close()
b = cuda.TrayCudaScene.build(g["tris"]); b.render(host.view_from_camera(cam, 160, 96), 160, 96, 0, cuda.RENDER_BOUNCE); b.sync(); b.close()
print("sanitizer script done")
PY
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool"; timeout 60 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san.py big 2>&1 | grep -E "ERROR SUMMARY|done|Error|error" | head -5
done 2>&1 | tee gpurun_out/r2_sanitizer.log

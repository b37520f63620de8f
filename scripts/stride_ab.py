"""Dev: kernel times with 48-byte {v0,e1,e2} vs 64-byte {v0,e1,e2,ng} triangle records (ng recomputed vs loaded)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host
w, h = 1920, 1080
for scene, seed in (("hairball", 3), ("kitchen", 1), ("sanmiguel", 4)):
    m = host.Mesh.generate(scene, seed, 1.0)
    view = host.view_from_camera(m.camera, w, h)
    for rep in range(2):
        for stride in (48, 64):
            p = host.PackedScene(m, tri_stride=stride)
            sc = cuda.TrayCudaScene.from_packed(p)
            k = [sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_RGBA) for _ in range(7)][1:]
            sc.close()
            kp, kb = min(a for a, _ in k), min(b for _, b in k)
            print(f"{scene} stride {stride}: primary {kp:.3f} bounce {kb:.3f} sum {kp + kb:.3f}", flush=True)

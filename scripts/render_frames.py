"""Dev/profiling driver: render a few frames of a synthetic scene (used under ncu)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="hairball")
ap.add_argument("--seed", type=int, default=3)
ap.add_argument("--size", type=float, default=1.0)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--tlas", action="store_true")
ap.add_argument("--stride", type=int, default=48)
a = ap.parse_args()
m = host.Mesh.generate(a.scene, a.seed, a.size)
p = host.PackedScene(m, use_tlas=a.tlas, tri_stride=a.stride)
view = host.view_from_camera(m.camera, a.width, a.height, p.tlas_start)
sc = cuda.TrayCudaScene.from_packed(p)
best = None
for f in range(a.frames):
    ms = sc.render(view, a.width, a.height, 0, cuda.RENDER_BOUNCE | cuda.RENDER_RGBA)
    best = ms if best is None else (min(best[0], ms[0]), min(best[1], ms[1]))
print(a.scene, "tris", p.n_tris, "nodes", p.n_nodes, "ms primary/bounce", best)

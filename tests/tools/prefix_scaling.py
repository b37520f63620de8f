"""Dev: kernel time vs number of rays (prefixes of one frame's primary / bounce rays, tile order) — separates the
per-ray cost from the per-launch cost."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tray_racing_b200 import cuda, host  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "hairball"
w, h = 1920, 1080
m = host.Mesh.generate(scene, 3, 1.0)
p = host.PackedScene(m)
view = host.view_from_camera(m.camera, w, h)
sc = cuda.TrayCudaScene.from_packed(p)
sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_KEEP_RAYS)
out = sc.download(primary=True, bounce_rays=True)
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
from sched_sim import tile_order  # noqa: E402
import oracle_binding as ob  # noqa: E402
order = tile_order(w, h)
prim = np.ascontiguousarray(ob.primary_rays(view, w, h)[order])
hit = out["primary"]["prim"][order] != 0xFFFFFFFF
brays = np.ascontiguousarray(out["bounce_rays"][order][hit])
rng = np.random.default_rng(1)
for name, rays in (("primary", prim), ("bounce", brays), ("primary-shuffled-tiles", None)):
    if rays is None:
        t = prim.reshape(-1, 256)
        rays = np.ascontiguousarray(t[rng.permutation(t.shape[0])].reshape(-1))
    n = rays.shape[0]
    for frac in (1 / 1024, 1 / 256, 1 / 64, 1 / 16, 1 / 8, 1 / 4, 1 / 2, 1.0):
        k = max(32, int(n * frac))
        # a strided subset keeps the ray mix of the whole frame: every (1/frac)-th tile of 256 rays
        t = rays[: (n // 256) * 256].reshape(-1, 256)
        step = max(1, int(round(1 / frac)))
        sub = np.ascontiguousarray(t[::step].reshape(-1))
        best = 1e9
        for _ in range(5):
            tm = {}
            sc.traverse(sub, tm)
            best = min(best, tm["ms_kernel"])
        print(f"{scene} {name}: {sub.shape[0]:8d} rays  kernel {best:.4f} ms  {sub.shape[0] / best / 1e3:8.1f} Mrays/s", flush=True)

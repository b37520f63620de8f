"""Dev: per-kernel times of the §8 f4 variants (f16 triangle records, any-hit AO rays) beside the parity path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "hairball"
w, h = 1920, 1080
m = host.Mesh.generate(scene, 3, 1.0)
for stride in (48, 24):
    p = host.PackedScene(m, tri_stride=stride)
    view = host.view_from_camera(m.camera, w, h)
    sc = cuda.TrayCudaScene.from_packed(p)
    for name, extra in (("closest-hit AO", 0), ("any-hit AO", cuda.RENDER_ANYHIT_AO)):
        flags = cuda.RENDER_BOUNCE | cuda.RENDER_RGBA | extra
        k = [sc.render(view, w, h, 0, flags) for _ in range(8)][2:]
        sc.render(view, w, h, 0, flags | cuda.RENDER_COUNTERS)
        cp, cb = sc.counters()
        kp, kb = min(a for a, _ in k), min(b for _, b in k)
        print(f"{scene} stride {stride} {name}: primary {kp:.3f} ms ({cp['rays'] / kp / 1e3:.0f} Mrays/s, {cp['nodes'] / cp['rays']:.1f} nodes {cp['tris'] / cp['rays']:.1f} tris /ray) | "
              f"bounce {kb:.3f} ms ({cb['rays'] / kb / 1e3:.0f} Mrays/s, {cb['nodes'] / max(1, cb['rays']):.1f} nodes {cb['tris'] / max(1, cb['rays']):.1f} tris /ray)", flush=True)
    sc.close()

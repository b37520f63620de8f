"""Dev: device-side builder vs the host producer — build time, node count, and traversal cost of the two BVHs."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "hairball"
size = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
radii = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [14]
w, h = 1920, 1080
m = host.Mesh.generate(scene, 3, size)
tris = m.tris()
view = host.view_from_camera(m.camera, w, h)
cuda.TrayCudaScene.build(tris[:1000]).close()       # context + module warm-up
for radius in radii:
    best = None
    for rep in range(3):
        t0 = time.time(); g = cuda.TrayCudaScene.build(tris, search_radius=radius); t_gpu = time.time() - t0
        if best is None or t_gpu < best[0]:
            best = (t_gpu, dict(g.build_stats))
        if rep < 2:
            g.close()
    t_gpu, st = best
    k = [g.render(view, w, h, 0) for _ in range(6)][2:]
    g.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_COUNTERS); cp, cb = g.counters()
    print(f"{scene} x{size} ({tris.shape[0]} tris): GPU build r={radius}: {t_gpu * 1e3:.1f} ms wall ({st['ms_upload']:.1f} upload + {st['ms_sort']:.1f} sort + {st['ms_ploc']:.1f} ploc/{st['ploc_iterations']} it + "
          f"{st['ms_collapse']:.1f} collapse/{st['levels']} lv), {st['n_nodes']} nodes | primary {min(a for a, _ in k):.3f} ms bounce {min(b for _, b in k):.3f} ms, "
          f"{cp['nodes'] / cp['rays']:.1f} nodes {cp['tris'] / cp['rays']:.1f} tris /primary ray, {cb['nodes'] / max(1, cb['rays']):.1f} nodes /bounce ray", flush=True)
    g.close()
if os.environ.get("SKIP_HOST") != "1":
    t0 = time.time(); p = host.PackedScene(m); t_host = time.time() - t0
    t0 = time.time(); c = cuda.TrayCudaScene.from_packed(p); t_up = time.time() - t0
    k = [c.render(view, w, h, 0) for _ in range(6)][2:]
    c.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_COUNTERS); cp, cb = c.counters()
    print(f"{scene} x{size}: host build {t_host * 1e3:.0f} ms (+ upload {t_up * 1e3:.0f} ms), {p.n_nodes} nodes | primary {min(a for a, _ in k):.3f} ms bounce {min(b for _, b in k):.3f} ms, "
          f"{cp['nodes'] / cp['rays']:.1f} nodes {cp['tris'] / cp['rays']:.1f} tris /primary ray, {cb['nodes'] / max(1, cb['rays']):.1f} nodes /bounce ray")

"""Dev: two-launch vs one-launch frame on ONE tile shard of a 3840x2160 frame (what each rank renders at N GPUs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host
w, h = 3840, 2160
for scene, seed, tlas in (("sanmiguel", 4, False), ("caldera", 5, True)):
    m = host.Mesh.generate(scene, seed, 1.0)
    p = host.PackedScene(m, use_tlas=tlas)
    view = host.view_from_camera(m.camera, w, h, p.tlas_start)
    sc = cuda.TrayCudaScene.from_packed(p)
    for shards in (1, 2, 4, 8):
        res = {}
        for rep in range(2):
            for name, extra in (("two", 0), ("one", cuda.RENDER_OVERLAP)):
                k = [sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_RGBA | extra, shard=0, shards=shards) for _ in range(6)][1:]
                best = min(a + b for a, b in k)
                res[name] = min(res.get(name, 1e9), best)
        print(f"{scene}{' --tlas' if tlas else ''} 4K shard 0/{shards}: two launches {res['two']:.3f} ms, one launch {res['one']:.3f} ms ({(res['one'] / res['two'] - 1) * 100:+.1f} %)", flush=True)
    sc.close()

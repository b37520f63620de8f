"""Dev: device-builder timings, A/B of the single-block PLOC tail (TRAY_BUILD_PLOC_TAIL=0 keeps the multi-kernel loop)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host
for scene, seed, tlas in (("hairball", 3, False), ("kitchen", 1, False), ("caldera", 5, True)):
    m = host.Mesh.generate(scene, seed, 1.0)
    tris = m.tris()
    offs = m.object_offsets() if tlas else None
    cuda.TrayCudaScene.build(tris, object_offsets=offs).close()
    best = {}
    for rep in range(5):
        for tail in ("1", "0"):
            os.environ["TRAY_BUILD_PLOC_TAIL"] = tail
            t0 = time.time(); g = cuda.TrayCudaScene.build(tris, object_offsets=offs); t = (time.time() - t0) * 1e3
            st = dict(g.build_stats); g.close()
            if tail not in best or st["ms_total"] < best[tail]["ms_total"]:
                best[tail] = dict(st, wall=t)
    for tail in ("1", "0"):
        st = best[tail]
        print(f"{scene}{' --tlas' if tlas else ''} ({tris.shape[0]} tris) tail={tail}: best of 5: total {st['ms_total']:.1f} ms (upload {st['ms_upload']:.1f} ploc {st['ms_ploc']:.1f}/{st['ploc_iterations']} it collapse {st['ms_collapse']:.1f}), wall {st['wall']:.1f}", flush=True)

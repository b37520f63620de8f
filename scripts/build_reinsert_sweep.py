#!/usr/bin/env python
"""Dev: the device builder's reinsertion passes (TRAY_CUDA_BUILD_REINSERT / _VISITS) against BVH quality and build time.
usage: build_reinsert_sweep.py [scene ...]   (default: hairball kitchen sanmiguel)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host  # noqa: E402

SEEDS = {"kitchen": 1, "demoscene": 2, "hairball": 3, "sanmiguel": 4, "caldera": 5}
scenes = sys.argv[1:] or ["hairball", "kitchen", "sanmiguel"]
for scene in scenes:
    m = host.Mesh.generate(scene, SEEDS[scene], 1.0)
    tris = m.tris()
    w, h = (3840, 2160) if scene in ("sanmiguel", "caldera") else (1920, 1080)
    view = host.view_from_camera(m.camera, w, h)
    p = host.PackedScene(m)
    sc = cuda.TrayCudaScene.from_packed(p)
    k = [sc.render(view, w, h, 0, cuda.RENDER_BOUNCE) for _ in range(6)][2:]
    sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_COUNTERS)
    cp, cb = sc.counters()
    sc.close()
    print(f"{scene} host binned-SAH BVH: {p.n_nodes} nodes, primary {min(a for a, _ in k):.3f} ms bounce {min(b for _, b in k):.3f} ms, "
          f"nodes/ray {cp['nodes'] / cp['rays']:.2f} {cb['nodes'] / max(1, cb['rays']):.2f}", flush=True)
    configs = ((0, 192, 1), (1, 192, 1), (1, 192, 3), (2, 192, 3), (2, 192, 6), (3, 192, 4), (4, 192, 3), (4, 64, 3), (8, 192, 3))
    if os.environ.get("SWEEP_SHORT"):
        configs = ((0, 192, 1), (4, 192, 3))
    for passes, visits, rounds in configs:
        os.environ["TRAY_CUDA_BUILD_REINSERT"] = str(passes)
        os.environ["TRAY_CUDA_BUILD_REINSERT_VISITS"] = str(visits)
        os.environ["TRAY_CUDA_BUILD_REINSERT_ROUNDS"] = str(rounds)
        cuda.TrayCudaScene.build(tris).close()
        best = None
        for _ in range(3):
            g = cuda.TrayCudaScene.build(tris)
            st = g.build_stats
            if best is None or st["ms_total"] < best["ms_total"]:
                best = dict(st)
            if _ < 2:
                g.close()
        k = [g.render(view, w, h, 0, cuda.RENDER_BOUNCE) for _ in range(6)][2:]
        g.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_COUNTERS)
        cp, cb = g.counters()
        g.close()
        print(f"{scene} device PLOC + {passes} reinsertion pass(es) x {visits} visits x {rounds} rounds: {best['n_nodes']} nodes, build {best['ms_total']:.1f} ms (reinsert {best['ms_reinsert']:.1f}), "
              f"moves {best['reinsert_moves']}, SAH {best['sah_before']:.4g} -> {best['sah_after']:.4g}, primary {min(a for a, _ in k):.3f} ms bounce {min(b for _, b in k):.3f} ms, "
              f"nodes/ray {cp['nodes'] / cp['rays']:.2f} {cb['nodes'] / max(1, cb['rays']):.2f}", flush=True)

"""Dev probe (GPU box): parity of the CUDA path against the oracle on a few scenes + first timings.
Not part of the test suite; tests/test_gpu_*.py hold the real checks."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from tray_racing_b200 import cuda, host  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def compare(name, gpu, orc):
    same_prim = gpu["prim"] == orc["prim"]
    same_t = bits(gpu["t"]) == bits(orc["t"])
    n = len(gpu)
    print(f"  {name}: prim equal {same_prim.sum()}/{n}, t bit-equal {same_t.sum()}/{n}")
    return bool(same_prim.all() and same_t.all())


def check_scene(label, mesh, w, h, use_tlas=False, stride=48):
    p = host.PackedScene(mesh, use_tlas=use_tlas, tri_stride=stride)
    view = host.view_from_camera(mesh.camera, w, h, p.tlas_start)
    orc = ob.Oracle.from_packed(p)
    ref = orc.render(view, w, h, frame_count=0)
    sc = cuda.TrayCudaScene.from_packed(p)
    print(f"[{label}] tris {p.n_tris} nodes {p.n_nodes} tlas={use_tlas} stride={stride} info={sc.info()}")
    flags = cuda.RENDER_BOUNCE | cuda.RENDER_RGBA | cuda.RENDER_COUNTERS | cuda.RENDER_KEEP_RAYS
    ms = sc.render(view, w, h, 0, flags)
    out = sc.download(primary=True, bounce=True, bounce_rays=True, rgba=True)
    ok = compare("primary", out["primary"], ref["primary"])
    rays_equal = (out["bounce_rays"].view(np.uint32).reshape(-1, 8) == ref["bounce_rays"].view(np.uint32).reshape(-1, 8)).all(axis=1)
    print(f"  bounce rays bit-equal {rays_equal.sum()}/{w*h}")
    ok &= bool(rays_equal.all())
    ok &= compare("bounce", out["bounce"], ref["bounce"])
    cp, cb = sc.counters()
    print("  counters gpu", cp, cb)
    print("  counters orc", ref["primary_totals"], ref["bounce_totals"])
    ok &= cp["nodes"] == ref["primary_totals"]["nodes"] and cp["tris"] == ref["primary_totals"]["tris"]
    ok &= cb["nodes"] == ref["bounce_totals"]["nodes"] and cb["tris"] == ref["bounce_totals"]["tris"]
    # batch operator on the same primary rays
    rays = ob.primary_rays(view, w, h)
    t = {}
    hits = sc.traverse(rays, t)
    ok &= compare("traverse(primary rays)", hits, ref["primary"])
    print(f"  kernels ms (counting build) primary/bounce = {ms}, traverse {t}  -> {'OK' if ok else 'MISMATCH'}")
    sc.close()
    return ok


def timing(label, mesh, w, h, frames=5):
    t0 = time.time()
    p = host.PackedScene(mesh)
    print(f"[{label}] build {time.time()-t0:.1f}s tris {p.n_tris} nodes {p.n_nodes} WS {p.working_set_bytes()/1e6:.1f} MB")
    view = host.view_from_camera(mesh.camera, w, h)
    sc = cuda.TrayCudaScene.from_packed(p)
    flags = cuda.RENDER_BOUNCE | cuda.RENDER_RGBA
    sc.render(view, w, h, 0, flags | cuda.RENDER_COUNTERS)
    cp, cb = sc.counters()
    res = []
    for f in range(frames + 2):
        a, b = sc.render(view, w, h, 0, flags)
        if f >= 2:
            res.append((a, b))
    a = min(r[0] for r in res); b = min(r[1] for r in res)
    bytes_p = 80 * cp["nodes"] + 48 * cp["tris"] + 8 * cp["rays"]
    bytes_b = 80 * cb["nodes"] + 48 * cb["tris"] + 8 * cb["rays"]
    print(f"  primary {a:.3f} ms  {cp['rays']/a/1e3:.1f} Mrays/s  {bytes_p/a/1e6:.1f} GB/s algorithmic ({bytes_p/cp['rays']:.0f} B/ray, {cp['nodes']/cp['rays']:.1f} nodes {cp['tris']/cp['rays']:.1f} tris)")
    if cb["rays"]:
        print(f"  bounce  {b:.3f} ms  {cb['rays']/b/1e3:.1f} Mrays/s  {bytes_b/b/1e6:.1f} GB/s algorithmic ({bytes_b/cb['rays']:.0f} B/ray, {cb['nodes']/cb['rays']:.1f} nodes {cb['tris']/cb['rays']:.1f} tris)")
    sc.close()
    return dict(label=label, ms_primary=a, ms_bounce=b, primary=cp, bounce=cb)


def main():
    print("devices", cuda.device_count())
    g = np.load(os.path.join(ROOT, "tests/golden/cornell_box.npz"))
    cam = host.Camera(tuple(g["eye"]), tuple(g["look_at"]), float(g["fov"]))
    cornell = host.Mesh.from_tris(g["tris"], g["offsets"], cam)
    ok = True
    ok &= check_scene("cornell flat", cornell, 640, 360)
    ok &= check_scene("cornell flat s64", cornell, 640, 360, stride=64)
    ok &= check_scene("cornell tlas", cornell, 640, 360, use_tlas=True)
    ok &= check_scene("hairball 5%", host.Mesh.generate("hairball", 3, 0.05), 640, 360)
    ok &= check_scene("caldera 1% tlas", host.Mesh.generate("caldera", 5, 0.01), 640, 360, use_tlas=True)
    print("PARITY", "OK" if ok else "FAILED")
    results = []
    if "--timing" in sys.argv:
        results.append(timing("kitchen", host.Mesh.generate("kitchen", 1, 1.0), 1920, 1080))
        results.append(timing("hairball", host.Mesh.generate("hairball", 3, 1.0), 1920, 1080))
        results.append(timing("demoscene", host.Mesh.generate("demoscene", 2, 1.0), 1920, 1080))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out/probe.json"), "w") as f:
        json.dump(dict(parity=bool(ok), timing=results), f, indent=1)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

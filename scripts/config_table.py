"""The five BASELINE.json configs on one GPU: kernel times (min of 6 frames, CUDA events), Mrays/s, and the algorithmic
bytes per ray B = 80 Nn + 48 Nt + 4 Ni + 8 from the counting build.  Prints a markdown table.
  python scripts/config_table.py [out.md]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host  # noqa: E402

CONFIGS = [("C1 kitchen-sized interior", "kitchen", 1, 1920, 1080, False),
           ("C2 demoscene stand-in (height field)", "demoscene", 2, 1920, 1080, False),
           ("C3 hairball-like soup", "hairball", 3, 1920, 1080, False),
           ("C4 San-Miguel-sized", "sanmiguel", 4, 3840, 2160, False),
           ("C5 Caldera-sized, --tlas", "caldera", 5, 3840, 2160, True)]
rows = ["| config | tris | nodes | working set MB | frame | primary ms | Mrays/s | nodes / tris / inst per ray | B/ray | GB/s alg. | bounce rays | bounce ms | Mrays/s | nodes / tris / inst per ray | B/ray | GB/s alg. |",
        "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
only = os.environ.get("CONFIGS")
for label, name, seed, w, h, tlas in CONFIGS:
    if only and name not in only.split(","):
        continue
    t0 = time.time()
    m = host.Mesh.generate(name, seed, 1.0)
    p = host.PackedScene(m, use_tlas=tlas)
    view = host.view_from_camera(m.camera, w, h, p.tlas_start)
    sc = cuda.TrayCudaScene.from_packed(p)
    t_setup = time.time() - t0
    flags = cuda.RENDER_BOUNCE | cuda.RENDER_RGBA
    k = [sc.render(view, w, h, 0, flags) for _ in range(7)][1:]
    sc.render(view, w, h, 0, flags | cuda.RENDER_COUNTERS)
    cp, cb = sc.counters()
    info = sc.info()
    sc.close()
    kp, kb = min(a for a, _ in k), min(b for _, b in k)

    def cols(c, ms):
        n = max(1, c["rays"])
        bpr = (80 * c["nodes"] + p.tri_stride * c["tris"] + 4 * c["instances"]) / n + 8
        return (f"{ms:.3f} | {c['rays'] / ms / 1e3:.0f} | {c['nodes'] / n:.1f} / {c['tris'] / n:.1f} / {c['instances'] / n:.2f} | "
                f"{bpr:.0f} | {bpr * c['rays'] / ms / 1e6:.0f}")
    ws = (info["n_nodes"] * 80 + info["n_tris"] * p.tri_stride) / 1e6
    rows.append(f"| {label} | {info['n_tris']} | {info['n_nodes']} | {ws:.0f} | {w}x{h} | {cols(cp, kp)} | {cb['rays']} | {cols(cb, kb)} |")
    print(rows[-1], f"   (setup {t_setup:.1f} s)", flush=True)
text = "\n".join(rows) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text)

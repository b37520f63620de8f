import os, sys, time
sys.path.insert(0, "/root/repo")
from tray_racing_b200 import cuda, host
m = host.Mesh.generate("hairball", 3, 1.0)
tris = m.tris()
cuda.TrayCudaScene.build(tris[:1000]).close()
def run(tag):
    for rep in range(4):
        t0 = time.time(); g = cuda.TrayCudaScene.build(tris); t = (time.time() - t0) * 1e3
        st = g.build_stats
        t1 = time.time(); g.close(); tc = (time.time() - t1) * 1e3
        print(f"{tag} rep {rep}: wall {t:.1f} ms (upload {st['ms_upload']:.1f} ploc {st['ms_ploc']:.1f} collapse {st['ms_collapse']:.1f} total {st['ms_total']:.1f}), close {tc:.1f} ms", flush=True)
run("plain")
import torch
x = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); x.fill_(1); torch.cuda.synchronize()
run("with torch + 256MiB")
p = host.PackedScene(m)
sc = cuda.TrayCudaScene.from_packed(p)
view = host.view_from_camera(m.camera, 1920, 1080)
for _ in range(20): sc.render(view, 1920, 1080, 0)
run("after renders, scene alive")

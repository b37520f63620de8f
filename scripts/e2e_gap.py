"""Dev: where does the end-to-end frame loop lose time against the device-timed step? (wall clock per frame)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host
scene_name = sys.argv[1] if len(sys.argv) > 1 else "demoscene"
seed = {"demoscene": 2, "hairball": 3, "kitchen": 1}.get(scene_name, 3)
m = host.Mesh.generate(scene_name, seed, 1.0)
p = host.PackedScene(m)
w, h = 1920, 1080
view = host.view_from_camera(m.camera, w, h)
sc = cuda.TrayCudaScene.from_packed(p)
flags = cuda.RENDER_BOUNCE | cuda.RENDER_RGBA
frames = [torch.empty((h, w, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
N = 200
def run(label, body, tail=None):
    for i in range(5): body(i)
    sc.sync()
    t0 = time.perf_counter()
    for i in range(N): body(i)
    if tail: tail()
    sc.sync()
    print(f"{scene_name} {label}: {(time.perf_counter() - t0) / N * 1e3:.3f} ms per frame", flush=True)
run("render only (enqueue, one sync at the end)", lambda i: sc.render(view, w, h, 0, flags, timed=False))
run("render + sync every frame", lambda i: (sc.render(view, w, h, 0, flags, timed=False), sc.sync()))
def rb(i):
    sc.render(view, w, h, 0, flags, timed=False); sc.readback_begin(frames[i & 1], i & 1)
run("render + readback_begin (waits only when a slot is reused)", rb, lambda: (sc.readback_wait(0), sc.readback_wait(1)))
def full(i):
    sc.render(view, w, h, 0, flags, timed=False); sc.readback_begin(frames[i & 1], i & 1)
    if i > 0: sc.readback_wait((i & 1) ^ 1)
run("render + readback_begin + wait for the previous frame (bench e2e)", full, lambda: (sc.readback_wait(0), sc.readback_wait(1)))
sc.close()

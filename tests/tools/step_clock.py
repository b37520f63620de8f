"""Dev: cycle breakdown of the traversal loop, from the instrumented build (`make -C tray_racing_b200/csrc stepclock`).
Run on a GPU:  TRAY_CUDA_LIB=$PWD/tray_racing_b200/libtray_cuda_stepclock.so python tests/tools/step_clock.py [scene]
Cases: the 32 / 4736 longest primary rays one per warp (the drain phase in isolation) and the whole 1080p primary batch."""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("TRAY_EXIT_LOG_FILE", "/tmp/stepclock")
import oracle_binding as ob
from tray_racing_b200 import cuda, host

scene = sys.argv[1] if len(sys.argv) > 1 else "hairball"
m = host.Mesh.generate(scene, 3, 1.0)
p = host.PackedScene(m)
w, h = 1920, 1080
view = host.view_from_camera(m.camera, w, h)
orc = ob.Oracle.from_packed(p)
rays = ob.primary_rays(view, w, h)
_, cnt, tot = orc.trace(rays, counts=True)
steps = cnt["nodes"].astype(np.int64) + cnt["tris"]
order = np.argsort(-steps, kind="stable")
sc = cuda.TrayCudaScene.from_packed(p)
names = ["vote", "node pre", "node load wait", "node test", "node post", "tri pre", "tri load wait", "tri test+post"]


def report(label, batch):
    for f in glob.glob(os.environ["TRAY_EXIT_LOG_FILE"] + ".*.bin"):
        os.remove(f)
    for _ in range(3):
        t = {}
        sc.traverse(batch, t)
    f = sorted(glob.glob(os.environ["TRAY_EXIT_LOG_FILE"] + ".*.bin"), key=lambda x: int(x.split(".")[-2]))[-1]
    a = np.fromfile(f, dtype=np.int64).reshape(-1, 12)
    a = a[a[:, 11] == 1]
    a = a[np.argsort(-a[:, 10])][: max(1, len(a) // 8)]          # the longest-running eighth of the warps
    n_node, n_tri, total = a[:, 8].mean(), a[:, 9].mean(), a[:, 10].mean()
    print(f"== {label}: kernel {t['ms_kernel'] * 1e3:.0f} us; longest-running {len(a)} warps: {total:.0f} cycles, {n_node:.0f} node + {n_tri:.0f} tri iterations")
    acc = a[:, :8].mean(axis=0)
    for i, nm in enumerate(names):
        per = acc[i] / (n_node + n_tri if i == 0 else n_node if i < 5 else n_tri)
        print(f"   {nm:16s} {acc[i] / total * 100:5.1f} %   {per:7.1f} cycles per {'iteration' if i == 0 else 'node step' if i < 5 else 'tri step'}")
    print(f"   unaccounted      {(1 - acc.sum() / total) * 100:5.1f} %")


for L in (32, 4736):
    spread = np.zeros(L * 32, dtype=ob.RAY_DTYPE)
    spread["o"] = rays["o"][0]; spread["d"] = rays["d"][0]; spread["tmax"] = 0.0
    spread[::32] = rays[order[:L]]
    report(f"{L} longest rays, one per warp (longest {steps[order[0]]} steps)", spread)
tile = np.fromfile  # noqa
report("whole primary batch (row-major order)", rays)
sc.close()

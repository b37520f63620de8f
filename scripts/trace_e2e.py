"""Dev: batch-grain operator (tray_cuda_trace: host rays in, host hits out) — kernel ms vs call ms, Mrays/s end to end."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host
m = host.Mesh.generate(sys.argv[1] if len(sys.argv) > 1 else "hairball", 3, 1.0)
p = host.PackedScene(m)
sc = cuda.TrayCudaScene.from_packed(p)
rng = np.random.default_rng(1)
for n in (100_000, 2_073_600, 8_294_400, 33_177_600):
    rays = np.zeros(n, dtype=host.RAY_DTYPE)
    o = rng.normal(size=(n, 3)).astype(np.float32); o /= np.linalg.norm(o, axis=1, keepdims=True); o *= 7
    d = -o + rng.normal(size=(n, 3)).astype(np.float32) * 2; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["o"], rays["d"], rays["tmax"] = o, d, np.float32(3.4e38)
    best = None
    for _ in range(4):
        t = {}
        t0 = time.perf_counter(); sc.traverse(rays, t); wall = (time.perf_counter() - t0) * 1e3
        if best is None or wall < best[0]:
            best = (wall, t["ms_kernel"], t["ms_total"])
    print(f"n {n}: call {best[0]:.2f} ms (C ABI {best[2]:.2f} ms, kernel {best[1]:.2f} ms) -> {n / best[0] / 1e3:.0f} Mrays/s end to end, {n / best[1] / 1e3:.0f} kernel", flush=True)
sc.close()

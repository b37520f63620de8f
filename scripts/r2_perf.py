#!/usr/bin/env python
"""Dev A/B (round 2): per-kernel times and frame THROUGHPUT (one / two frames in flight x two-launch / one-launch frame path)
of one build of the CUDA library on one scene, optionally on one tile shard of a larger frame (what one GPU of N traces).

  TRAY_CUDA_LIB=$PWD/tray_racing_b200/libtray_cuda_x.so python scripts/r2_perf.py hairball [--w 1920 --h 1080] [--shards 8]
Prints one line per measurement; `sum=` is a checksum of the hits so that builds can be compared for bit-equality."""
import argparse
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host  # noqa: E402

SEEDS = {"kitchen": 1, "demoscene": 2, "hairball": 3, "sanmiguel": 4, "caldera": 5}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--shards", type=int, default=1)
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--tlas", action="store_true")
    ap.add_argument("--device-build", action="store_true")
    a = ap.parse_args()
    tag = os.path.basename(os.environ.get("TRAY_CUDA_LIB", "libtray_cuda.so"))
    m = host.Mesh.generate(a.scene, SEEDS[a.scene], 1.0)
    if a.device_build:
        sc = cuda.TrayCudaScene.build(m.tris())
        tlas_start = 0
    else:
        p = host.PackedScene(m, use_tlas=a.tlas)
        sc = cuda.TrayCudaScene.from_packed(p)
        tlas_start = p.tlas_start
    w, h, S = a.w, a.h, a.shards
    view = host.view_from_camera(m.camera, w, h, tlas_start)
    base = cuda.RENDER_BOUNCE | cuda.RENDER_RGBA
    k = [sc.render(view, w, h, 0, base, 0, S) for _ in range(8)][2:]
    out = sc.download(primary=True, bounce=True)
    crc = zlib.crc32(out["bounce"].tobytes(), zlib.crc32(out["primary"].tobytes()))
    sc.render(view, w, h, 0, base | cuda.RENDER_COUNTERS, 0, S)
    cp, cb = sc.counters()
    rays = cp["rays"] + cb["rays"]
    kp, kb = min(x for x, _ in k), min(y for _, y in k)
    print(f"{tag} {a.scene} {w}x{h}/{S}: primary {kp:.3f} ms ({cp['rays'] / kp / 1e3:.0f} Mrays/s) bounce {kb:.3f} ms ({cb['rays'] / max(kb, 1e-9) / 1e3:.0f} Mrays/s) "
          f"sum {kp + kb:.3f} nodes/ray {cp['nodes'] / cp['rays']:.2f} {cb['nodes'] / max(1, cb['rays']):.2f} crc={crc:08x}", flush=True)
    for ov in (0, 1):
        for nf in (1, 2, 3):
            sc.set_frames_in_flight(nf)
            fl = base | (cuda.RENDER_OVERLAP if ov else 0)
            best = None
            for rep in range(3):
                for _ in range(4):
                    sc.render(view, w, h, 0, fl, 0, S, timed=False)
                sc.sync()
                t0 = time.perf_counter()
                for _ in range(a.frames):
                    sc.render(view, w, h, 0, fl, 0, S, timed=False)
                sc.sync()
                ms = (time.perf_counter() - t0) * 1e3 / a.frames
                best = ms if best is None else min(best, ms)
            print(f"{tag} {a.scene} {w}x{h}/{S}: {'one' if ov else 'two'}-launch x{nf} in flight: {best:.3f} ms/frame ({rays / best / 1e3:.0f} Mrays/s)", flush=True)
    sc.set_frames_in_flight(1)
    sc.close()


if __name__ == "__main__":
    main()

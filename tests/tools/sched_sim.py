#!/usr/bin/env python
"""Warp-schedule study for the traversal kernel (analysis tool, test infrastructure: it uses the CPU oracle).

  python tests/tools/sched_sim.py [--scene hairball] [--w 960 --h 540]

Replays the oracle's per-ray step logs through the issue-slot model in sched_sim.c for several lane organisations
and prints slots/ray and lane occupancy, so that a kernel restructuring can be sized before it is written."""
import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from tray_racing_b200 import host  # noqa: E402


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("policy", "K", "n_warps", "refill_min", "tri_weight", "cn", "ct", "cr", "csel")]


class Out(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("slots", "makespan", "node_steps", "tri_steps", "node_lanes", "tri_lanes", "refills")]


def sim_lib():
    so = os.path.join(HERE, "libsched_sim.so")
    src = os.path.join(HERE, "sched_sim.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.sched_sim.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Cfg), C.POINTER(Out)]
    L.sched_sim4.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Cfg), C.POINTER(Out), C.POINTER(C.c_double)]
    return L


def tile_order(w, h):
    """row-major pixel index of the kernel's work item j (32x8 tiles, 8x4 sub-tiles; traverse.cuh item_to_pixel)"""
    tx, ty = (w + 31) // 32, (h + 7) // 8
    j = np.arange(tx * ty * 256, dtype=np.int64)
    k, wi = j >> 8, j & 255
    sub, l = wi >> 5, wi & 31
    px = (k % tx) * 32 + (sub & 3) * 8 + (l & 7)
    py = (k // tx) * 8 + (sub >> 2) * 4 + (l >> 3)
    ok = (px < w) & (py < h)
    return (py * w + px)[ok]


def oplog(orc, rays, detail=False):
    hits, cnt, tot = orc.trace(rays, counts=True)
    n_ops = cnt["nodes"].astype(np.uint64) + cnt["tris"] + cnt["insts"]
    offsets = np.zeros(rays.shape[0] + 1, dtype=np.uint64)
    np.cumsum(n_ops, out=offsets[1:])
    ops = np.zeros(int(offsets[-1]) + 1, dtype=np.uint8)
    L = ob.lib()
    fn = L.orc_trace_oplog_detail if detail else L.orc_trace_oplog
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(ob.OrcScene), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = fn(C.byref(orc.scene), rays.ctypes.data, rays.shape[0], cnt.ctypes.data, ops.ctypes.data, offsets.ctypes.data, 0)
    assert rc == 0, rc
    return ops, offsets, tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="hairball")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--w", type=int, default=960)
    ap.add_argument("--h", type=int, default=540)
    ap.add_argument("--warps", type=int, default=0, help="resident warps (default: scaled from 148 SMs x 32 by the pixel count)")
    ap.add_argument("--cn", type=int, default=300)
    ap.add_argument("--ct", type=int, default=165)
    ap.add_argument("--cr", type=int, default=80)
    ap.add_argument("--csel", type=int, default=25)
    ap.add_argument("--spec", action="store_true", help="round 2: only the tri2 / speculative-node-step study (policy 4, detailed log)")
    a = ap.parse_args()
    mesh = host.Mesh.generate(a.scene, a.seed, 1.0)
    packed = host.PackedScene(mesh, use_tlas=False, tri_stride=48)
    orc = ob.Oracle.from_packed(packed)
    view = host.view_from_camera(mesh.camera, a.w, a.h)
    order = tile_order(a.w, a.h)
    prim_rays = ob.primary_rays(view, a.w, a.h)[order]
    r = orc.render(view, a.w, a.h, 0)
    hitpix = r["primary"]["prim"][order] != ob.INVALID_PRIM
    bounce_rays = r["bounce_rays"][order][hitpix]
    n_warps = a.warps or max(64, int(148 * 32 * (a.w * a.h) / (1920 * 1080)))
    L = sim_lib()
    for name, rays in (("primary", prim_rays), ("bounce", bounce_rays)):
        rays = np.ascontiguousarray(rays)
        ops, offsets, tot = oplog(orc, rays)
        n = rays.shape[0]
        print(f"== {name}: {n} rays, nodes/ray {tot['nodes'] / n:.2f}, tris/ray {tot['tris'] / n:.2f}, warps {n_warps}")
        ideal = (tot["nodes"] * a.cn + tot["tris"] * a.ct) / 32.0 / n
        print(f"   ideal (32/32 lanes, no overhead): {ideal:.1f} warp-slots/ray")
        if a.spec:
            ops, offsets, tot = oplog(orc, rays, detail=True)
            base = None
            for label, K, tw, cn, ct, csel in (("one triangle per step (r1)", 0, 4, 300, 130, 0), ("tri2 (the r2 kernel)", 1, 4, 290, 170, 0), ("tri2 tw=3", 1, 3, 290, 170, 0),
                                               ("tri2 + speculative node step (+20)", 3, 4, 290, 175, 20), ("  .. tw=3", 3, 3, 290, 175, 20), ("  .. tw=2", 3, 2, 290, 175, 20), ("  .. tw=6", 3, 6, 290, 175, 20),
                                               ("  .. free (+0)", 3, 4, 290, 170, 0),
                                               ("tri2, refill cost 140 (rm 4)", 1, 4, 290, 170, -140), ("tri2, refill 60, rm 4", 1, 4, 290, 170, -60), ("tri2, refill 60, rm 2", 1, 4, 290, 170, -60 - 2000),
                                               ("tri2, refill 60, rm 1", 1, 4, 290, 170, -60 - 1000), ("tri2, refill 140, rm 2", 1, 4, 290, 170, -140 - 2000), ("tri2, refill 140, rm 1", 1, 4, 290, 170, -140 - 1000)):
                rm, cr = 4, a.cr
                if csel < 0:            # encoded refill study: -(cost) - 1000 * refill_min
                    v = -csel; rm = v // 1000 or 4; cr = v % 1000; csel = 0
                cfg = Cfg(4, K, n_warps, rm, tw, cn, ct, cr, csel)
                out = Out(); extra = (C.c_double * 3)()
                L.sched_sim4(ops.ctypes.data, offsets.ctypes.data, n, C.byref(cfg), C.byref(out), extra)
                spr = out.slots / n
                base = base or spr
                print(f"   {label:36s} slots/ray {spr:7.1f}  x{base / spr:4.2f}  node steps/ray {out.node_steps * 32 / n:6.2f} lanes {out.node_lanes / max(1, out.node_steps):5.2f}  "
                      f"tri steps/ray {out.tri_steps * 32 / n:5.2f} lanes {out.tri_lanes / max(1, out.tri_steps):5.2f}  spec/ray {extra[0] / n:4.2f} committed {extra[1] / max(1, extra[0]):4.2f}")
            continue
        base = None
        for label, policy, K, rmin, tw in (("K=1 (round-1 kernel)", 0, 1, 4, 4), ("K=1 tw=2", 0, 1, 4, 2), ("K=1 tw=8", 0, 1, 4, 8),
                                           ("drain tw=4 (cn 300, loop 45 + 110/round)", 2, 1, 4, 4), ("drain tw=2", 2, 1, 4, 2), ("drain tw=1", 2, 1, 4, 1), ("drain tw=8", 2, 1, 4, 8), ("K=2 per lane", 0, 2, 8, 4), ("K=2 per lane tw=2", 0, 2, 8, 2), ("K=3 per lane", 0, 3, 8, 4),
                                           ("K=4 per lane", 0, 4, 8, 4),
                                           ("narrow<=4 (pass 120)", 3, 4, 4, 4), ("narrow<=8 (pass 120)", 3, 8, 4, 4), ("narrow<=8 tw=8", 3, 8, 4, 8),
                                           ("pool 40 tw=1", 1, 40, 4, 1), ("pool 48 tw=1", 1, 48, 8, 1), ("pool 48 tw=2", 1, 48, 8, 2), ("pool 64 tw=4", 1, 64, 8, 4), ("pool 64 tw=2", 1, 64, 8, 2), ("pool 64 tw=1", 1, 64, 8, 1),
                                           ("pool 64 tw=1 rm=16", 1, 64, 16, 1), ("pool 64 tw=1 rm=4", 1, 64, 4, 1), ("pool 96 tw=1", 1, 96, 8, 1), ("pool 128 tw=1", 1, 128, 8, 1)):
            slots = K if policy == 1 else 32 * (K if policy == 0 else 1)
            cfg = Cfg(policy, K, max(1, n_warps * 32 // slots), rmin, tw, a.cn, a.ct if policy != 2 else 110, a.cr, 120 if policy == 3 else (a.csel if policy != 2 else 45))
            out = Out()
            L.sched_sim(ops.ctypes.data, offsets.ctypes.data, n, C.byref(cfg), C.byref(out))
            spr = out.slots / n
            base = base or spr
            print(f"   {label:24s} slots/ray {spr:7.1f}  x{base / spr:4.2f}  node lanes {out.node_lanes / max(1, out.node_steps):5.2f}/32 "
                  f"tri lanes {out.tri_lanes / max(1, out.tri_steps):5.2f}/32  tail {out.makespan * cfg.n_warps / out.slots:4.2f}")


if __name__ == "__main__":
    main()

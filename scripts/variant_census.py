#!/usr/bin/env python
"""Census of the unpinned semantics (VERDICT r1 "next" #1b): on FULL-SIZE frames of the five BASELINE.json configs, how many
rays change (prim, t) when the kernel runs the other reading of each semantic that is sourced from memory of the un-vendored
obvhs crate (tray_cuda_scene_set_variant).  GPU only — the product's MODE 1 kernels; the oracle checks them in
tests/test_gpu_variants.py.  Writes profiles/r2_variant_census.json and prints a markdown table.

  python scripts/variant_census.py [--configs c1,c2,c3,c4,c5] [--out profiles/r2_variant_census.json]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tray_racing_b200 import cuda, host  # noqa: E402

CONFIGS = {
    "c1": ("kitchen", 1, 1920, 1080, False), "c2": ("demoscene", 2, 1920, 1080, False), "c3": ("hairball", 3, 1920, 1080, False),
    "c4": ("sanmiguel", 4, 3840, 2160, False), "c5": ("caldera", 5, 3840, 2160, True),
}
SWITCHES = {"box_divide": cuda.VARIANT_BOX_DIVIDE, "tie_last": cuda.VARIANT_TIE_LAST, "box_tmin_ray": cuda.VARIANT_BOX_TMIN_RAY,
            "zerodir_box_only": cuda.VARIANT_ZERODIR_BOX_ONLY, "all_four": 0xF}
INVALID = 0xFFFFFFFF


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def compare(base, out):
    d = {}
    for k in ("primary", "bounce"):
        b, o = base[k], out[k]
        ch = (b["prim"] != o["prim"]) | (bits(b["t"]) != bits(o["t"]))
        both = ch & (b["prim"] != INVALID) & (o["prim"] != INVALID)
        rel = np.abs(o["t"][both].astype(np.float64) - b["t"][both]) / np.abs(b["t"][both].astype(np.float64)) if both.any() else np.zeros(0)
        d[k] = {"changed": int(ch.sum()), "prim_changed": int((b["prim"] != o["prim"]).sum()),
                "hit_miss_flips": int(((b["prim"] == INVALID) != (o["prim"] == INVALID)).sum()),
                "max_rel_dt": float(rel.max()) if rel.size else 0.0, "over_1e-5": int((rel > 1e-5).sum())}
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_variant_census.json"))
    a = ap.parse_args()
    res = {}
    for c in a.configs.split(","):
        name, seed, w, h, tlas = CONFIGS[c]
        m = host.Mesh.generate(name, seed, 1.0)
        p = host.PackedScene(m, use_tlas=tlas)
        view = host.view_from_camera(m.camera, w, h, p.tlas_start)
        sc = cuda.TrayCudaScene.from_packed(p)
        # the rays of the DEFAULT frame: its primary rays (a function of the pixel) and its bounce rays, kept, so that every
        # switch is asked about the very same rays (a bounce ray regenerated from a moved primary hit is a different ray)
        sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_KEEP_RAYS)
        fr = sc.download(primary=True, bounce=True, bounce_rays=True)
        hitpix = fr["primary"]["prim"] != INVALID
        brays = np.ascontiguousarray(fr["bounce_rays"][hitpix])
        base = {"primary": fr["primary"], "bounce": sc.traverse(brays)}
        assert (base["bounce"]["prim"] == fr["bounce"]["prim"][hitpix]).all()
        n_p, n_b = w * h, int(hitpix.sum())
        res[c] = {"scene": name, "n_tris": p.n_tris, "frame": [w, h], "tlas": tlas, "primary_rays": n_p, "bounce_rays": n_b,
                  "bounce_rays_with_a_zero_direction_component": int((brays["d"] == 0).any(axis=1).sum()), "switches": {}}
        for sw, v in SWITCHES.items():
            sc.set_variant(v)
            sc.render(view, w, h, 0, 0)
            out = {"primary": sc.download(primary=True)["primary"], "bounce": sc.traverse(brays)}
            res[c]["switches"][sw] = compare(base, out)
        sc.set_variant(0)
        sc.render(view, w, h, 0, 0)
        again = {"primary": sc.download(primary=True)["primary"], "bounce": sc.traverse(brays)}
        assert compare(base, again)["primary"]["changed"] == 0 and compare(base, again)["bounce"]["changed"] == 0
        sc.close()
        print(c, json.dumps(res[c]["switches"]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print("\n| config | rays (primary + bounce) | " + " | ".join(SWITCHES) + " |")
    print("|---|---|" + "---|" * len(SWITCHES))
    for c, r in res.items():
        cells = []
        for sw in SWITCHES:
            s = r["switches"][sw]
            cells.append(f"{s['primary']['changed'] + s['bounce']['changed']} ({s['primary']['prim_changed'] + s['bounce']['prim_changed']} prim, "
                         f"max rel dt {max(s['primary']['max_rel_dt'], s['bounce']['max_rel_dt']):.1e})")
        print(f"| {c} {r['scene']} | {r['primary_rays']} + {r['bounce_rays']} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()

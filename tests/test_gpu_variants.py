"""Parity hardening (VERDICT r1 "next" #1): the semantics of the reference's CPU path that are sourced from memory of the
un-vendored obvhs crate — tie rule, slab-test lower clamp, reach of the zero-direction patch — and the divide-form box test of
the HLSL twin are RUN-TIME switches of both the oracle (orc_set_variant) and the CUDA path (tray_cuda_scene_set_variant).
For every switch and combination the kernel is bit-identical to the oracle under the same switch; and a census reports how
many rays change (prim, t) against the default, so the exposure of the unpinned parity is a number (scripts/variant_census.py
produces the full-size table in profiles/)."""
import numpy as np
import pytest

import oracle_binding as ob
from conftest import random_rays
from tray_racing_b200 import cuda, host

pytestmark = pytest.mark.gpu

SWITCHES = [cuda.VARIANT_BOX_DIVIDE, cuda.VARIANT_TIE_LAST, cuda.VARIANT_BOX_TMIN_RAY, cuda.VARIANT_ZERODIR_BOX_ONLY]
COMBOS = SWITCHES + [cuda.VARIANT_TIE_LAST | cuda.VARIANT_ZERODIR_BOX_ONLY, 0xF]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def identical(a, b):
    return bool((a["prim"] == b["prim"]).all() and (bits(a["t"]) == bits(b["t"])).all())


@pytest.fixture(autouse=True)
def _reset_oracle_variant():
    yield
    ob.set_variant(0)


def test_flag_values_agree_between_oracle_and_product():
    assert (ob.VARIANT_BOX_DIVIDE, ob.VARIANT_TIE_LAST, ob.VARIANT_BOX_TMIN_RAY, ob.VARIANT_ZERODIR_BOX_ONLY) == tuple(SWITCHES)


@pytest.mark.parametrize("use_tlas,stride", [(False, 48), (True, 64), (False, 24)])
def test_every_switch_matches_the_oracle_on_random_rays(cornell, use_tlas, stride):
    """Rays with exact-zero direction components (10 %) and finite [tmin, tmax] windows (30 %): the inputs the switches bite on."""
    p = host.PackedScene(cornell, use_tlas=use_tlas, tri_stride=stride)
    orc = ob.Oracle.from_packed(p)
    sc = cuda.TrayCudaScene.from_packed(p)
    rays = random_rays(200003, 11, axis_fraction=0.1, bounded_fraction=0.3)
    try:
        base = sc.traverse(rays)
        assert identical(base, orc.trace(rays))
        changed = {}
        for v in COMBOS:
            ob.set_variant(v)
            sc.set_variant(v)
            got, want = sc.traverse(rays), orc.trace(rays)
            assert identical(got, want), f"variant 0x{v:x}: kernel and oracle disagree on {(got['prim'] != want['prim']).sum()} prims"
            changed[v] = int(((got["prim"] != base["prim"]) | (bits(got["t"]) != bits(base["t"]))).sum())
        sc.set_variant(0)
        ob.set_variant(0)
        assert identical(sc.traverse(rays), base)                       # and back
        # the zero-direction switch must bite on these rays (axis-parallel rays graze box faces exactly), the census is not vacuous
        assert changed[cuda.VARIANT_ZERODIR_BOX_ONLY] > 0 or changed[0xF] > 0
    finally:
        sc.close()


def test_switches_on_a_frame_with_coincident_triangles(box):
    """box.obj under --tlas holds coplanar, coincident faces (equal-t hits): the tie rule decides the primitive there."""
    p = host.PackedScene(box, use_tlas=True)
    w, h = 320, 200
    view = host.view_from_camera(box.camera, w, h, p.tlas_start)
    orc = ob.Oracle.from_packed(p)
    sc = cuda.TrayCudaScene.from_packed(p)
    try:
        for v in [0] + COMBOS:
            ob.set_variant(v)
            sc.set_variant(v)
            ref = orc.render(view, w, h, 0)
            sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_OVERLAP)         # OVERLAP is ignored under a variant (two launches)
            out = sc.download(primary=True, bounce=True)
            for k in ("primary", "bounce"):
                assert identical(out[k], ref[k]), f"variant 0x{v:x} {k}"
    finally:
        sc.close()


def test_variants_refuse_what_they_do_not_cover(cornell):
    p = host.PackedScene(cornell)
    sc = cuda.TrayCudaScene.from_packed(p)
    try:
        with pytest.raises(cuda.TrayCudaError, match="unknown variant"):
            sc.set_variant(0x10)
        sc.set_variant(cuda.VARIANT_TIE_LAST)
        rays = random_rays(100, 1)
        with pytest.raises(cuda.TrayCudaError, match="closest-hit kernels without counters"):
            sc.traverse(rays, any_hit=True)
        sc.set_counting(True)
        with pytest.raises(cuda.TrayCudaError, match="closest-hit kernels without counters"):
            sc.traverse(rays)
    finally:
        sc.close()


@pytest.mark.parametrize("name,seed,size,tlas,w,h", [("kitchen", 1, 1.0, False, 960, 540), ("hairball", 3, 0.25, False, 960, 540),
                                                     ("caldera", 5, 0.05, True, 960, 540)])
def test_census_of_affected_rays_on_frames(name, seed, size, tlas, w, h):
    """How many rays of a frame (primary + bounce) change (prim, t) under each switch — kernel == oracle for each, and the
    count is small: the exposure of the three from-memory semantics.  (Even frame widths put ndc.x == 0 on a pixel column only
    when the camera is axis-aligned; the census says what that costs.)"""
    m = host.Mesh.generate(name, seed, size)
    p = host.PackedScene(m, use_tlas=tlas)
    view = host.view_from_camera(m.camera, w, h, p.tlas_start)
    orc = ob.Oracle.from_packed(p)
    sc = cuda.TrayCudaScene.from_packed(p)
    try:
        sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_KEEP_RAYS)
        base = sc.download(primary=True, bounce=True, bounce_rays=True)
        n_rays = w * h + int((base["primary"]["prim"] != ob.INVALID_PRIM).sum())
        for v in SWITCHES:
            ob.set_variant(v)
            sc.set_variant(v)
            sc.render(view, w, h, 0, cuda.RENDER_BOUNCE)
            out = sc.download(primary=True, bounce=True)
            ref = orc.render(view, w, h, 0)
            for k in ("primary", "bounce"):
                assert identical(out[k], ref[k]), f"{name} variant 0x{v:x} {k}"
            # bounce rays descend from primary hits: count a pixel once if either of its rays changed
            ch = ((out["primary"]["prim"] != base["primary"]["prim"]) | (bits(out["primary"]["t"]) != bits(base["primary"]["t"])) |
                  (out["bounce"]["prim"] != base["bounce"]["prim"]) | (bits(out["bounce"]["t"]) != bits(base["bounce"]["t"])))
            assert ch.sum() <= 0.002 * n_rays, f"{name} variant 0x{v:x}: {ch.sum()} of {n_rays} rays changed"
    finally:
        sc.close()

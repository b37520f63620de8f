"""any-hit query and the f16-compressed triangle records of the reference's wgpu path
(src/rt_gpu/mod.rs:39-43; rt_gpu_software_query.hlsl:75-85,440-445; rt_cpu.rs:78-79).

CPU tests pin the oracle's half decode and the host packer's half encode against numpy's IEEE binary16, and the
any-hit predicate against the closest-hit traversal.  GPU tests (through the C ABI) hold the kernel to the oracle on
the SAME records, bit for bit.  The f16 records are NOT the parity path against rt_cpu: the last test measures how far
they move hit distances."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
from conftest import random_rays
from tray_racing_b200 import cuda, host

F32_MAX = np.float32(3.402823466e+38)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---------------------------------------------------------------- CPU: records and the oracle -----
def test_packer_f16_records_match_numpy_half(cornell):
    """tray_host_pack(stride 24) == numpy's round-to-nearest-even binary16 of the same edges, on real geometry and on
    edge magnitudes (half subnormals, ties, overflow to inf, signed zeros)."""
    p = host.PackedScene(cornell, tri_stride=24)
    tris = cornell.tris()[p.prim_to_mesh_tri]
    assert np.array_equal(p.tri_bytes, host.tri_records(tris, 24))
    rng = np.random.default_rng(7)
    n = 4096
    t = np.zeros((n, 3, 3), dtype=np.float32)
    mag = np.float32(2.0) ** rng.integers(-27, 18, size=(n, 3)).astype(np.float32)
    t[:, 1] = (rng.uniform(-2, 2, size=(n, 3)).astype(np.float32) * mag)
    t[:, 2] = (rng.integers(-4096, 4096, size=(n, 3)).astype(np.float32) * np.float32(2.0 ** -25))   # exact ties around subnormals
    t[0, 1] = [65519.9, -65520.0, 65504.0]; t[0, 2] = [0.0, -0.0, 6.1e-5]
    m = host.Mesh.from_tris(t.reshape(-1, 9))
    q = host.PackedScene(m, tri_stride=24)
    assert np.array_equal(q.tri_bytes, host.tri_records(t[q.prim_to_mesh_tri], 24))


def test_oracle_half_decode_all_values(cornell):
    """The oracle decodes every one of the 65536 binary16 patterns like numpy, and its triangle test on an f16 record
    equals its test on the f32 record made of the decoded edges."""
    L = ob.lib()
    pats = np.arange(65536, dtype=np.uint32)
    ref = pats.astype(np.uint16).view(np.float16).astype(np.float32)
    got = np.array([L.orc_half_to_float(int(h)) for h in pats], dtype=np.float32)
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(got), nan) and np.array_equal(bits(got[~nan]), bits(ref[~nan]))
    p = host.PackedScene(cornell, tri_stride=24)
    rec24 = p.tri_bytes.view(np.uint32).reshape(-1, 6)
    rec48 = np.zeros((rec24.shape[0], 12), dtype=np.float32)
    rec48[:, 0:3] = rec24[:, 0:3].view(np.float32)
    rec48[:, 8:11] = (rec24[:, 3:6] & 0xffff).astype(np.uint16).view(np.float16).astype(np.float32)          # e2
    rec48[:, 4:7] = -((rec24[:, 3:6] >> 16).astype(np.uint16).view(np.float16).astype(np.float32))           # e1 = -(v1 - v0)
    a = ob.Oracle(p.bvh_bytes, p.tri_bytes, 24)
    b = ob.Oracle(p.bvh_bytes, rec48.view(np.uint8).reshape(-1), 48)
    rays = random_rays(5000, seed=3)
    ha, hb = a.trace(rays), b.trace(rays)
    assert np.array_equal(ha["prim"], hb["prim"]) and np.array_equal(bits(ha["t"]), bits(hb["t"]))
    assert (ha["prim"] != ob.INVALID_PRIM).sum() > 1000


def test_anyhit_predicate_equals_closest_hit(cornell):
    """orc_trace_any finds a hit exactly when orc_trace does, tests no more nodes / triangles than it, and the hit it
    reports is a real intersection (same t as the triangle test alone)."""
    for use_tlas in (False, True):
        p = host.PackedScene(cornell, use_tlas=use_tlas)
        orc = ob.Oracle.from_packed(p)
        rays = random_rays(20000, seed=11)
        closest, cc, _ = orc.trace(rays, counts=True)
        anyh, ca, _ = orc.trace(rays, counts=True, any_hit=True)
        assert np.array_equal(anyh["prim"] != ob.INVALID_PRIM, closest["prim"] != ob.INVALID_PRIM)
        assert (ca["nodes"] <= cc["nodes"]).all() and (ca["tris"] <= cc["tris"]).all()
        assert ca["tris"].sum() < cc["tris"].sum()
        hit = np.nonzero(anyh["prim"] != ob.INVALID_PRIM)[0][:500]
        for i in hit:
            assert bits(np.float32(orc.intersect_tri(anyh["prim"][i], rays[i:i + 1])))[()] == bits(anyh["t"][i:i + 1])[0]
        assert (anyh["t"][hit] >= closest["t"][hit]).all()


# ---------------------------------------------------------------- GPU -------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("use_tlas", [False, True])
def test_gpu_f16_records_frame(cornell, use_tlas):
    """stride-24 records: primary hits, generated bounce rays and bounce hits bit-identical to the oracle on the same
    records, with identical node / triangle counts."""
    from test_gpu_parity import render_and_compare
    render_and_compare(cornell, 320, 184, use_tlas=use_tlas, stride=24)
    render_and_compare(host.Mesh.generate("hairball", 3, 0.05), 256, 144, use_tlas=False, stride=24)


@pytest.mark.gpu
@pytest.mark.parametrize("stride,use_tlas", [(48, False), (64, True), (24, False)])
def test_gpu_anyhit_matches_oracle(cornell, stride, use_tlas):
    p = host.PackedScene(cornell, use_tlas=use_tlas, tri_stride=stride)
    orc = ob.Oracle.from_packed(p)
    sc = cuda.TrayCudaScene.from_packed(p)
    try:
        for n in (1, 31, 4097, 50000):
            rays = random_rays(n, seed=n)
            want = orc.trace(rays, any_hit=True)
            got = sc.traverse(rays, any_hit=True)
            assert np.array_equal(got["prim"], want["prim"]) and np.array_equal(bits(got["t"]), bits(want["t"]))
            closest = sc.traverse(rays)
            assert np.array_equal(got["prim"] != ob.INVALID_PRIM, closest["prim"] != ob.INVALID_PRIM)
    finally:
        sc.close()


@pytest.mark.gpu
def test_gpu_anyhit_ao_frame(cornell):
    """TRAY_RENDER_ANYHIT_AO: the bounce buffer holds each AO ray's first hit (== oracle any-hit on the same rays), the
    image is 0 / 255 visibility on hit pixels."""
    w, h = 320, 184
    p = host.PackedScene(cornell)
    view = host.view_from_camera(cornell.camera, w, h)
    sc = cuda.TrayCudaScene.from_packed(p)
    try:
        sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_RGBA | cuda.RENDER_KEEP_RAYS | cuda.RENDER_ANYHIT_AO)
        out = sc.download(primary=True, bounce=True, bounce_rays=True, rgba=True)
    finally:
        sc.close()
    orc = ob.Oracle.from_packed(p)
    ref = orc.render(view, w, h, 0)
    assert np.array_equal(out["bounce_rays"].view(np.uint32), ref["bounce_rays"].view(np.uint32))
    shot = ref["primary"]["prim"] != ob.INVALID_PRIM
    want = orc.trace(ref["bounce_rays"][shot], any_hit=True)
    assert np.array_equal(out["bounce"]["prim"][shot], want["prim"]) and np.array_equal(bits(out["bounce"]["t"][shot]), bits(want["t"]))
    occluded = want["prim"] != ob.INVALID_PRIM
    grey = out["rgba"].reshape(-1, 4)[shot][:, 0]
    assert (grey[occluded] == 0).all() and (grey[~occluded] == 255).all()
    assert np.array_equal(occluded, ref["bounce"]["prim"][shot] != ob.INVALID_PRIM)


@pytest.mark.gpu
def test_gpu_f16_records_deviation_from_f32(cornell):
    """What the f16 edges cost: against the f32 parity path the hit distance moves by up to ~3e-4 relative and a few
    silhouette pixels change primitive — the reason the parity path keeps f32 records (SURVEY.md §8c delta i)."""
    w, h = 320, 184
    view = host.view_from_camera(cornell.camera, w, h)
    outs = {}
    for stride in (48, 24):
        sc = cuda.TrayCudaScene.from_packed(host.PackedScene(cornell, tri_stride=stride))
        try:
            sc.render(view, w, h, 0, 0)
            outs[stride] = sc.download(primary=True)["primary"]
        finally:
            sc.close()
    a, b = outs[48], outs[24]
    both = (a["prim"] != ob.INVALID_PRIM) & (b["prim"] != ob.INVALID_PRIM)
    assert both.mean() > 0.4
    same_prim = both & (a["prim"] == b["prim"])
    assert same_prim.sum() > 0.98 * both.sum()
    rel = np.abs(a["t"][same_prim] - b["t"][same_prim]) / a["t"][same_prim]
    assert rel.max() < 5e-3 and rel.max() > 1e-5     # measurably outside the 1e-5 bar, hence "not the parity path"

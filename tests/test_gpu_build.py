"""the device-side CWBVH builder (tray_cuda_scene_build), held to the same checks as the host producer:
structural validation in the spirit of obvhs `bvh.validate` (src/cwbvh.rs:102-104), traversal results bit-identical to the
CPU oracle run on the downloaded bytes, closest hits equal to an exhaustive search, and byte-for-byte determinism."""
import numpy as np
import pytest

import oracle_binding as ob
from conftest import random_rays
from tray_racing_b200 import cuda, host

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def build_and_check(tris, stride=48, max_leaf=3, radius=0):
    tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 3, 3)
    sc = cuda.TrayCudaScene.build(tris, tri_stride=stride, max_prims_per_leaf=max_leaf, search_radius=radius)
    nodes, tri_bytes, pi = sc.download_bvh()
    n = tris.shape[0]
    assert sc.build_stats["n_tris"] == n and sc.build_stats["n_nodes"] == nodes.size // 80
    if n:
        assert np.array_equal(np.sort(pi), np.arange(n, dtype=np.uint32)), "prim_indices is not a permutation"
        assert np.array_equal(tri_bytes, host.tri_records(tris[pi], stride)), "triangle records differ from RtTriangle::from"
        rc, rep = host.validate_cwbvh(nodes, pi, tris.min(axis=1), tris.max(axis=1), stack_limit=48)
        assert rc == 0, rep
        assert rep["prims_reached"] == n and rep["nodes_reached"] == nodes.size // 80
    return sc, nodes, tri_bytes, pi


def test_build_cornell_box_validates_and_traces_like_the_oracle(cornell):
    tris = cornell.tris()
    sc, nodes, tri_bytes, pi = build_and_check(tris)
    try:
        orc = ob.Oracle(nodes, tri_bytes, 48)
        rays = random_rays(60000, seed=5)
        got, want = sc.traverse(rays), orc.trace(rays)
        assert np.array_equal(got["prim"], want["prim"]) and np.array_equal(bits(got["t"]), bits(want["t"]))
        assert (want["prim"] != ob.INVALID_PRIM).sum() > 10000
        # against an exhaustive search: same t bits (primitive may differ only between equal-t triangles)
        sub = rays[:3000]
        brute, ties = orc.brute_force(sub)
        fast = got[:3000]
        hit = brute["prim"] != ob.INVALID_PRIM
        assert np.array_equal(hit, fast["prim"] != ob.INVALID_PRIM)
        close = np.abs(fast["t"][hit] - brute["t"][hit]) <= 1e-5 * np.abs(brute["t"][hit])
        assert close.all()
        exact = bits(fast["t"][hit]) == bits(brute["t"][hit])
        assert exact.mean() > 0.999
        # a whole frame through the frame operator on the device-built scene
        w, h = 320, 184
        view = host.view_from_camera(cornell.camera, w, h)
        sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_RGBA | cuda.RENDER_KEEP_RAYS)
        out = sc.download(primary=True, bounce=True)
        ref = orc.render(view, w, h, 0)
        for k in ("primary", "bounce"):
            assert np.array_equal(out[k]["prim"], ref[k]["prim"]) and np.array_equal(bits(out[k]["t"]), bits(ref[k]["t"]))
    finally:
        sc.close()


@pytest.mark.parametrize("stride,max_leaf,radius", [(48, 3, 0), (64, 1, 4), (24, 2, 32), (48, 3, 1)])
def test_build_parameters(cornell, stride, max_leaf, radius):
    sc, nodes, tri_bytes, pi = build_and_check(cornell.tris(), stride, max_leaf, radius)
    try:
        orc = ob.Oracle(nodes, tri_bytes, stride)
        rays = random_rays(20000, seed=stride + radius)
        got, want = sc.traverse(rays), orc.trace(rays)
        assert np.array_equal(got["prim"], want["prim"]) and np.array_equal(bits(got["t"]), bits(want["t"]))
    finally:
        sc.close()


def test_build_edge_cases():
    """empty soup, one triangle, two, coincident triangles (identical Morton codes), a degenerate (zero-area) triangle,
    and sizes around the block / radius boundaries"""
    rng = np.random.default_rng(3)
    one = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float32)
    cases = [np.zeros((0, 3, 3), np.float32), one, np.concatenate([one, one + 2]), np.repeat(one, 37, axis=0),
             np.concatenate([one, np.zeros((1, 3, 3), np.float32), one + 1])]
    for n in (3, 4, 8, 9, 24, 25, 255, 256, 257, 1000):
        c = rng.uniform(-1, 1, size=(n, 1, 3)).astype(np.float32)
        cases.append(c + rng.uniform(-0.05, 0.05, size=(n, 3, 3)).astype(np.float32))
    for tris in cases:
        sc, nodes, tri_bytes, pi = build_and_check(tris)
        try:
            rays = random_rays(2000, seed=tris.shape[0])
            got = sc.traverse(rays)
            if tris.shape[0]:
                want = ob.Oracle(nodes, tri_bytes, 48).trace(rays)
                assert np.array_equal(got["prim"], want["prim"]) and np.array_equal(bits(got["t"]), bits(want["t"]))
            else:
                assert (got["prim"] == ob.INVALID_PRIM).all()
        finally:
            sc.close()


def test_build_is_deterministic_and_comparable_to_the_host_producer():
    """Same triangles -> same bytes (ranks that each build their replica agree on every primitive id); and the PLOC tree is
    in the same quality class as the host producer's binned-SAH tree (nodes fetched per primary ray within 1.5x)."""
    m = host.Mesh.generate("hairball", 3, 0.1)
    tris = m.tris()
    a = cuda.TrayCudaScene.build(tris)
    b = cuda.TrayCudaScene.build(tris)
    try:
        na, ta, pa = a.download_bvh()
        nb, tb, pb = b.download_bvh()
        assert np.array_equal(na, nb) and np.array_equal(ta, tb) and np.array_equal(pa, pb)
        w, h = 480, 270
        view = host.view_from_camera(m.camera, w, h)
        a.render(view, w, h, 0, cuda.RENDER_COUNTERS)
        ca, _ = a.counters()
        p = host.PackedScene(m)
        c = cuda.TrayCudaScene.from_packed(p)
        try:
            c.render(view, w, h, 0, cuda.RENDER_COUNTERS)
            cc, _ = c.counters()
            # same geometry, two BVHs: every pixel must see the same surface
            ha = a.download(primary=True)["primary"]
            hc = c.download(primary=True)["primary"]
        finally:
            c.close()
        assert np.array_equal(ha["prim"] != ob.INVALID_PRIM, hc["prim"] != ob.INVALID_PRIM)
        hit = ha["prim"] != ob.INVALID_PRIM
        assert (np.abs(ha["t"][hit] - hc["t"][hit]) <= 1e-5 * hc["t"][hit]).mean() > 0.9999
        orig_a, orig_c = pa[ha["prim"][hit]], p.prim_to_mesh_tri[hc["prim"][hit]]
        assert (orig_a == orig_c).mean() > 0.999          # same input triangle, up to equal-t neighbours
        assert ca["nodes"] < 1.5 * cc["nodes"], (ca, cc)
    finally:
        a.close(); b.close()


@pytest.mark.parametrize("radius", [0, 3, 40])
def test_single_block_ploc_tail_builds_the_same_bvh(cornell, monkeypatch, radius):
    """The last PLOC iterations (<= 1024 clusters) run in one block; TRAY_BUILD_PLOC_TAIL=0 keeps the multi-kernel loop to the
    end.  Both must produce the same bytes: flat scenes, a two-level scene (segmented run + TLAS run), several radii."""
    meshes = [(cornell.tris(), None), (host.Mesh.generate("kitchen", 1, 1.0).tris(), None), (cornell.tris()[:700], None),
              (cornell.tris(), cornell.object_offsets())]
    cal = host.Mesh.generate("caldera", 5, 0.02)
    meshes.append((cal.tris(), cal.object_offsets()))
    for tris, offs in meshes:
        got = []
        for tail in ("1", "0"):
            monkeypatch.setenv("TRAY_BUILD_PLOC_TAIL", tail)
            sc = cuda.TrayCudaScene.build(tris, search_radius=radius, object_offsets=offs)
            try:
                got.append(sc.download_bvh() + ((sc.download_instances(),) if offs is not None else ()) + (sc.build_stats["ploc_iterations"],))
            finally:
                sc.close()
        for a, b in zip(*got):
            assert np.array_equal(a, b)


def two_level_check(mesh, w, h, stride=48, max_leaf=3):
    """device-built BLAS forest + TLAS: layout invariants, oracle parity in two-level mode, and the same picture as a flat
    device build of the same triangles"""
    tris, offs = mesh.tris(), mesh.object_offsets()
    n_obj = offs.size - 1
    sc = cuda.TrayCudaScene.build(tris, tri_stride=stride, max_prims_per_leaf=max_leaf, object_offsets=offs)
    try:
        info = sc.info()
        assert info["is_tlas"] and info["n_instances"] == n_obj and info["tlas_start"] > 0
        nodes, tri_bytes, pi = sc.download_bvh()
        blas = sc.download_instances()
        assert np.array_equal(np.sort(pi), np.arange(tris.shape[0], dtype=np.uint32))
        assert np.array_equal(tri_bytes, host.tri_records(tris[pi], stride))
        assert np.array_equal(np.sort(blas), np.arange(n_obj, dtype=np.uint32))      # BLAS k starts at node k; TLAS-leaf order permutes them
        assert info["tlas_start"] + 1 <= nodes.size // 80
        # every triangle slot belongs to the object whose BLAS holds it: slots of one object are exactly its triangles
        obj_of_tri = np.searchsorted(offs, pi, side="right") - 1
        assert (np.bincount(obj_of_tri, minlength=n_obj) == np.diff(offs).astype(np.int64)).all()
        orc = ob.Oracle(nodes, tri_bytes, stride, blas_offsets=blas, tlas_start=info["tlas_start"], use_tlas=True)
        rays = random_rays(30000, seed=n_obj)
        got, want = sc.traverse(rays), orc.trace(rays)
        assert np.array_equal(got["prim"], want["prim"]) and np.array_equal(bits(got["t"]), bits(want["t"]))
        view = host.view_from_camera(mesh.camera, w, h, info["tlas_start"])
        sc.render(view, w, h, 0, cuda.RENDER_BOUNCE | cuda.RENDER_COUNTERS)
        out = sc.download(primary=True, bounce=True)
        cp, _ = sc.counters()
        ref = orc.render(view, w, h, 0)
        for k in ("primary", "bounce"):
            assert np.array_equal(out[k]["prim"], ref[k]["prim"]) and np.array_equal(bits(out[k]["t"]), bits(ref[k]["t"]))
        assert cp["instances"] == ref["primary_totals"]["insts"] and cp["instances"] > 0
    finally:
        sc.close()
    flat = cuda.TrayCudaScene.build(tris, tri_stride=stride, max_prims_per_leaf=max_leaf)
    try:
        _, _, pf = flat.download_bvh()
        flat.render(host.view_from_camera(mesh.camera, w, h), w, h, 0, 0)
        f = flat.download(primary=True)["primary"]
    finally:
        flat.close()
    a = out["primary"]
    hit = f["prim"] != ob.INVALID_PRIM
    assert np.array_equal(a["prim"] != ob.INVALID_PRIM, hit)
    assert (np.abs(a["t"][hit] - f["t"][hit]) <= 1e-5 * f["t"][hit]).mean() > 0.9999
    assert (pi[a["prim"][hit]] == pf[f["prim"][hit]]).mean() > 0.999


def test_two_level_build_cornell_box(cornell):
    two_level_check(cornell, 320, 184)
    two_level_check(cornell, 160, 92, stride=64, max_leaf=1)


def test_two_level_build_many_objects_and_tiny_blas(box):
    two_level_check(box, 200, 120)                                          # BLASes of a handful of triangles
    two_level_check(host.Mesh.generate("caldera", 5, 0.02), 320, 180)       # ~80 objects of 1k-20k triangles


def test_two_level_build_argument_errors(cornell):
    tris, offs = cornell.tris(), cornell.object_offsets().copy()
    bad = offs.copy(); bad[1] = bad[0]                                      # an empty object
    with pytest.raises(cuda.TrayCudaError):
        cuda.TrayCudaScene.build(tris, object_offsets=bad)
    bad = offs.copy(); bad[-1] -= 1                                         # offsets do not cover the triangles
    with pytest.raises(cuda.TrayCudaError):
        cuda.TrayCudaScene.build(tris, object_offsets=bad)

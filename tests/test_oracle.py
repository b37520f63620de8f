"""CPU tests of the oracle itself: known-answer vectors derived by hand from the in-tree spec
(reference src/rt_gpu/rt_gpu_software_query.hlsl, sampling.hlsl), an independent numpy restatement of the
ray generator, and brute-force cross-checks.  The reference holds no golden vectors: these
are the pins there are."""
import numpy as np
import pytest

import oracle_binding as ob
from conftest import random_rays
from tray_racing_b200 import host

F32_MAX = np.float32(3.402823466e+38)


def make_node(p=(0, 0, 0), e=(127, 127, 127), imask=0, child_base=1, prim_base=0, meta=(0,) * 8,
              lo=((0,) * 8,) * 3, hi=((0,) * 8,) * 3):
    n = np.zeros(80, dtype=np.uint8)
    n[0:12] = np.array(p, dtype=np.float32).view(np.uint8)
    n[12:15] = e
    n[15] = imask
    n[16:20] = np.array([child_base], dtype=np.uint32).view(np.uint8)
    n[20:24] = np.array([prim_base], dtype=np.uint32).view(np.uint8)
    n[24:32] = meta
    for a in range(3):
        n[32 + 16 * a:40 + 16 * a] = lo[a]
        n[40 + 16 * a:48 + 16 * a] = hi[a]
    return n


def ray(o, d, tmin=0.0, tmax=F32_MAX):
    r = np.zeros(1, dtype=ob.RAY_DTYPE)
    r["o"], r["d"], r["tmin"], r["tmax"] = o, d, tmin, tmax
    return r


def assert_matches_brute_force(hits, bf, ties):
    """Closest-hit traversal vs exhaustive search.  t agrees to the last bits; where the box test of a flat,
    axis-aligned node culls a coplanar triangle whose t is an ulp closer (box and triangle arithmetic round
    differently — inherent to the algorithm, reference included), t may differ by a few ulps and the id with
    it.  Ids differ ONLY on exact ties or such near-ties."""
    ht, bt = hits["t"].astype(np.float64), bf["t"].astype(np.float64)
    both = np.isfinite(ht) & np.isfinite(bt)
    assert (np.isfinite(ht) == np.isfinite(bt)).all()
    assert (np.abs(ht[both] - bt[both]) <= 1e-6 * np.abs(bt[both])).all()
    exact = hits["t"].view(np.uint32) == bf["t"].view(np.uint32)
    assert exact.mean() > 0.999
    differ = hits["prim"] != bf["prim"]
    assert ((ties[differ] > 1) | ~exact[differ]).all()
    assert differ.mean() < 0.01


def test_record_sizes():
    assert ob.RAY_DTYPE.itemsize == 32 and ob.HIT_DTYPE.itemsize == 8


def test_uhash_known_answers():
    # independent evaluation of sampling.hlsl:5-15 with Python integers
    def uhash(a, b):
        m = 0xFFFFFFFF
        x = ((a * 1597334673) & m) ^ ((b * 3812015801) & m)
        x ^= x >> 16; x = (x * 0x7feb352d) & m
        x ^= x >> 15; x = (x * 0x846ca68b) & m
        x ^= x >> 16
        return x
    L = ob.lib()
    for a, b in [(0, 0), (1, 0), (0, 1), (1919, (1079 << 11) + 1024), (0xFFFFFFFF, 0xFFFFFFFF), (12345, 67890)]:
        assert L.orc_uhash(a, b) == uhash(a, b)
    # hash_noise = unormf(uhash(x, (y << 11) + frame)) (sampling.hlsl:17-27); 1/float(0xffffffff) == 2**-32 in f32
    for x, y, f in [(0, 0, 0), (7, 9, 3), (1919, 1079, 1024)]:
        want = np.float32(np.float32(uhash(x, ((y << 11) + f) & 0xFFFFFFFF)) * np.float32(2.0 ** -32))
        assert np.float32(L.orc_hash_noise(x, y, f)) == want
        assert 0.0 <= want <= 1.0


def test_sincos_tau_accuracy_and_quadrants():
    for u in np.linspace(0, 1, 4097, dtype=np.float32):
        s, c = ob.sincos_tau(float(u))
        assert abs(s - np.sin(2 * np.pi * float(u))) < 5e-7
        assert abs(c - np.cos(2 * np.pi * float(u))) < 5e-7
    assert ob.sincos_tau(0.0) == (0.0, 1.0)
    assert ob.sincos_tau(0.25) == (1.0, -0.0) or ob.sincos_tau(0.25) == (1.0, 0.0)


def test_node_decode_leaf_and_inner_bits():
    """One node, unit grid (e=127 -> scale 1), origin p=0.  Slot 0: inner child box [0,1]^3.  Slot 3: leaf with
    2 triangles at offset 5, box [2,3]x[0,1]x[0,1].  Slot 6: leaf, 3 triangles at offset 7, box far off the ray."""
    meta = [0] * 8
    meta[0] = 0x20 | 24           # inner: 0b001_00000 | (24 + slot)   (bvh_embree_to_cwbvh.rs:153-154)
    meta[3] = 0x60 | 5            # leaf, 2 triangles (unary 011), first at offset 5  (:157-165)
    meta[6] = 0xE0 | 7            # leaf, 3 triangles (unary 111), first at offset 7
    lo = [[0] * 8 for _ in range(3)]; hi = [[0] * 8 for _ in range(3)]
    for a in range(3):
        hi[a][0] = 1
    lo[0][3], hi[0][3], hi[1][3], hi[2][3] = 2, 3, 1, 1
    lo[1][6], hi[1][6], hi[0][6], hi[2][6] = 200, 201, 1, 1
    n = make_node(imask=1, meta=meta, lo=lo, hi=hi)
    # ray along +x through y=z=0.5: hits slot 0 and slot 3, misses slot 6.  All directions >= 0 -> oct_inv = 7,
    # inner slot 0 lands on bit 24 + (0 ^ 7) = 31 (query.hlsl:253,297); leaf bits are offset..offset+count-1
    m = ob.node_intersect(n, ray((-1, 0.5, 0.5), (1, 0, 0)), F32_MAX)
    assert m == (1 << 31) | (0b11 << 5)
    # same ray reversed (-x): oct_inv = 3 (x bit clear) -> inner slot 0 lands on bit 24 + (0 ^ 3) = 27
    m = ob.node_intersect(n, ray((5, 0.5, 0.5), (-1, 0, 0)), F32_MAX)
    assert m == (1 << 27) | (0b11 << 5)
    # max_distance culls: the leaf box starts at t = 3 from x = -1, the inner box at t = 1
    assert ob.node_intersect(n, ray((-1, 0.5, 0.5), (1, 0, 0)), 2.5) == (1 << 31)
    assert ob.node_intersect(n, ray((-1, 0.5, 0.5), (1, 0, 0)), 0.5) == 0
    # ray up the y axis at x=z=0.5 reaches slot 6's box (y in [200,201]) and slot 0
    m = ob.node_intersect(n, ray((0.5, -1, 0.5), (0, 1, 0)), F32_MAX)
    assert m == (1 << 31) | (0b111 << 7)


def test_node_scale_exponent_and_origin():
    """e = 125 -> scale 2^-2; p = (10, 20, 30).  Child box = p + [4,8] * 0.25 = [1,2] offset."""
    meta = [0] * 8; meta[2] = 0x20 | 1      # leaf, 1 triangle at offset 1
    lo = [[0] * 8 for _ in range(3)]; hi = [[0] * 8 for _ in range(3)]
    for a in range(3):
        lo[a][2], hi[a][2] = 4, 8
    n = make_node(p=(10, 20, 30), e=(125, 125, 125), meta=meta, lo=lo, hi=hi)
    assert ob.node_intersect(n, ray((11.5, 21.5, 0), (0, 0, 1)), F32_MAX) == 0b10
    assert ob.node_intersect(n, ray((12.5, 21.5, 0), (0, 0, 1)), F32_MAX) == 0
    assert ob.node_intersect(n, ray((11.5, 21.5, 40), (0, 0, 1)), F32_MAX) == 0      # box behind the origin
    assert ob.node_intersect(n, ray((11.5, 21.5, 40), (0, 0, -1)), F32_MAX) == 0b10


def one_tri_scene(v0, v1, v2, stride=48):
    tri = np.array([v0, v1, v2], dtype=np.float32).reshape(1, 9)
    nodes, pidx, depth = host.build_cwbvh(tri)
    return ob.Oracle(nodes, host.tri_records(tri, stride), stride), nodes


def test_triangle_known_answers():
    o, nodes = one_tri_scene((0, 0, 0), (1, 0, 0), (0, 1, 0))
    assert nodes.shape == (1, 80)
    r = ray((0.25, 0.25, 1), (0, 0, -1))
    assert o.intersect_tri(0, r) == 1.0
    assert o.trace(r)[0]["prim"] == 0 and o.trace(r)[0]["t"] == 1.0
    # back face hits too (no culling, query.hlsl:94,116)
    assert o.intersect_tri(0, ray((0.25, 0.25, -2), (0, 0, 1))) == 2.0
    # outside the triangle, parallel to it, behind the origin: miss = +inf
    assert np.isinf(o.intersect_tri(0, ray((0.75, 0.75, 1), (0, 0, -1))))
    assert np.isinf(o.intersect_tri(0, ray((0.25, 0.25, 1), (1, 0, 0))))
    assert np.isinf(o.intersect_tri(0, ray((0.25, 0.25, 1), (0, 0, 1))))
    # [tmin, tmax] is inclusive at both ends (query.hlsl:119)
    assert o.intersect_tri(0, ray((0.25, 0.25, 1), (0, 0, -1), 1.0, 1.0)) == 1.0
    assert np.isinf(o.intersect_tri(0, ray((0.25, 0.25, 1), (0, 0, -1), 0.0, 0.999)))
    assert np.isinf(o.intersect_tri(0, ray((0.25, 0.25, 1), (0, 0, -1), 1.001, 5.0)))
    miss = o.trace(ray((0.75, 0.75, 1), (0, 0, -1)))[0]
    assert np.isinf(miss["t"]) and miss["prim"] == ob.INVALID_PRIM


def test_triangle_stride64_identical():
    rng = np.random.default_rng(5)
    tris = rng.uniform(-1, 1, size=(64, 9)).astype(np.float32)
    nodes, pidx, _ = host.build_cwbvh(tris)
    rays = random_rays(2000, 11)
    a = ob.Oracle(nodes, host.tri_records(tris[pidx], 48), 48).trace(rays)
    b = ob.Oracle(nodes, host.tri_records(tris[pidx], 64), 64).trace(rays)
    assert (a["prim"] == b["prim"]).all() and (a["t"].view(np.uint32) == b["t"].view(np.uint32)).all()


def test_tie_rule_first_wins_vs_hlsl_last_wins():
    """Two coincident triangles: the CPU rule `t < tmax` keeps the first one tested, the HLSL `tt <= t`
    (query.hlsl:120) the last.  Triangles of a leaf are tested highest bit first (:398)."""
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]] * 2, dtype=np.float32)
    nodes, pidx, _ = host.build_cwbvh(tri)
    o = ob.Oracle(nodes, host.tri_records(tri[pidx]), 48)
    r = ray((0.25, 0.25, 1), (0, 0, -1))
    first = o.trace(r)[0]["prim"]
    ob.set_variant(ob.VARIANT_TIE_LAST)
    try:
        last = o.trace(r)[0]["prim"]
    finally:
        ob.set_variant(0)
    assert {int(first), int(last)} == {0, 1} and first == 1      # bit 1 is tested before bit 0


def test_primary_rays_match_numpy_restatement(cornell):
    """orc_primary_ray against an independent float32 numpy evaluation of rt_cpu.rs:38-55."""
    w, h = 64, 40
    v = host.view_from_camera(cornell.camera, w, h)
    rays = ob.primary_rays(v, w, h)
    f = np.float32
    pinv = np.array(v.proj_inv, dtype=f).reshape(4, 4).T      # column-major -> rows
    vinv = np.array(v.view_inv, dtype=f).reshape(4, 4).T
    eye = np.array(v.eye, dtype=f)
    i = np.arange(w * h)
    px, py = (i % w).astype(f), (i // w).astype(f)
    uvx, uvy = px / f(w), f(1) - py / f(h)
    clip = np.stack([uvx * f(2) - f(1), uvy * f(2) - f(1), np.ones_like(uvx), np.ones_like(uvx)], 1)

    def mat_vec(m, x):          # ((c0*x + c1*y) + c2*z) + c3*w, rounded after every op
        return ((m[:, 0] * x[:, 0:1] + m[:, 1] * x[:, 1:2]) + m[:, 2] * x[:, 2:3]) + m[:, 3] * x[:, 3:4]
    vs = mat_vec(pinv, clip); vs = vs / vs[:, 3:4]
    d = mat_vec(vinv, vs)[:, :3] - eye
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    d = d / ln[:, None]
    assert (rays["d"].view(np.uint32) == d.astype(f).view(np.uint32)).all()
    assert (rays["o"] == eye).all() and (rays["tmin"] == 0).all() and (rays["tmax"] == F32_MAX).all()
    # centre ray looks at look_at; rays are unit length
    c = rays[(h // 2) * w + w // 2]["d"]
    look = np.array(cornell.camera.look_at) - np.array(cornell.camera.eye)
    assert np.dot(c, look / np.linalg.norm(look)) > 0.999


@pytest.mark.parametrize("scene,use_tlas", [("cornell", False), ("cornell", True), ("box", False), ("box", True)])
def test_oracle_vs_brute_force_golden_scenes(scene, use_tlas, cornell, box):
    mesh = cornell if scene == "cornell" else box
    p = host.PackedScene(mesh, use_tlas=use_tlas)
    o = ob.Oracle.from_packed(p)
    w, h = (160, 90)
    rays = np.concatenate([ob.primary_rays(host.view_from_camera(mesh.camera, w, h, p.tlas_start), w, h), random_rays(3000, 3)])
    hits = o.trace(rays)
    bf, ties = o.brute_force(rays)
    assert_matches_brute_force(hits, bf, ties)
    assert (hits["prim"] != ob.INVALID_PRIM).sum() > 1000


def test_oracle_vs_brute_force_soup_with_bounded_rays():
    m = host.Mesh.generate("soup", 7, 0.004)
    p = host.PackedScene(m)
    o = ob.Oracle.from_packed(p)
    rays = random_rays(6000, 9, lo=-1, hi=1, axis_fraction=0.05, bounded_fraction=0.3)
    hits = o.trace(rays)
    bf, ties = o.brute_force(rays)
    assert_matches_brute_force(hits, bf, ties)
    hit = hits["prim"] != ob.INVALID_PRIM
    assert (hits["t"][hit] >= rays["tmin"][hit]).all() and (hits["t"][hit] <= rays["tmax"][hit]).all()


def test_box_test_variants_do_not_change_results(cornell):
    """The box test of the CPU path multiplies by a cached reciprocal, the HLSL twin divides
    (query.hlsl:237-242, SURVEY.md §8c vi): only the set of visited nodes may differ, never (prim, t)."""
    p = host.PackedScene(cornell)
    o = ob.Oracle.from_packed(p)
    rays = np.concatenate([ob.primary_rays(host.view_from_camera(cornell.camera, 320, 180), 320, 180), random_rays(20000, 4)])
    a, ca, _ = o.trace(rays, counts=True)
    ob.set_variant(ob.VARIANT_BOX_DIVIDE)
    try:
        b, cb, _ = o.trace(rays, counts=True)
    finally:
        ob.set_variant(0)
    assert (a["prim"] == b["prim"]).all() and (a["t"].view(np.uint32) == b["t"].view(np.uint32)).all()
    assert abs(int(ca["nodes"].sum()) - int(cb["nodes"].sum())) < 0.001 * ca["nodes"].sum()


def test_tlas_equals_flat(cornell):
    """--tlas and flat traversal of the same triangles return the same closest hit
    (reference src/cwbvh.rs:144-193; global prim -> (geometry_id, primitive_id))."""
    flat, tl = host.PackedScene(cornell, use_tlas=False), host.PackedScene(cornell, use_tlas=True)
    assert tl.n_instances == 5 and tl.tlas_start == tl.n_nodes - (tl.n_nodes - tl.tlas_start)
    rays = np.concatenate([ob.primary_rays(host.view_from_camera(cornell.camera, 200, 120), 200, 120), random_rays(5000, 8)])
    a = ob.Oracle.from_packed(flat).trace(rays)
    b, cnt, tot = ob.Oracle.from_packed(tl).trace(rays, counts=True)
    assert (a["t"].view(np.uint32) == b["t"].view(np.uint32)).mean() > 0.999      # near-ties aside (see above)
    assert np.allclose(a["t"], b["t"], rtol=1e-6)
    hit = a["prim"] != ob.INVALID_PRIM
    assert (flat.prim_to_mesh_tri[a["prim"][hit]] == tl.prim_to_mesh_tri[b["prim"][hit]]).mean() > 0.999   # ties aside
    geom, local = tl.geometry_of(b["prim"][hit])
    offs = cornell.object_offsets()
    assert (offs[geom] + 0 <= tl.prim_to_mesh_tri[b["prim"][hit]]).all() and (tl.prim_to_mesh_tri[b["prim"][hit]] < offs[geom + 1]).all()
    assert tot["insts"] > 0


def test_render_frame_semantics(cornell):
    """orc_render = rt_cpu.rs:35-91: bounce rays exist exactly for primary hits, start 0.01 before the hit
    point, are unit length and leave on the side the primary ray came from."""
    p = host.PackedScene(cornell)
    o = ob.Oracle.from_packed(p)
    w, h = 160, 96
    v = host.view_from_camera(cornell.camera, w, h)
    r = o.render(v, w, h, frame_count=0, rgba=True)
    prim, b, br = r["primary"], r["bounce"], r["bounce_rays"]
    hit = prim["prim"] != ob.INVALID_PRIM
    assert hit.sum() == r["primary_totals"]["hits"] == r["bounce_totals"]["rays"]
    assert (br["tmax"][~hit] == 0).all() and (br["tmax"][hit] == F32_MAX).all()
    assert (b["prim"][~hit] == ob.INVALID_PRIM).all()
    d = br["d"][hit]
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    rays = ob.primary_rays(v, w, h)
    want_o = rays["o"][hit] + rays["d"][hit] * prim["t"][hit, None] - rays["d"][hit] * np.float32(0.01)
    assert np.allclose(br["o"][hit], want_o, atol=1e-5)
    # bounce rays re-traced through the batch operator give the recorded bounce hits
    again = o.trace(br[hit])
    assert (again["prim"] == b["prim"][hit]).all()
    # different frame_count -> different directions (hash_noise(px, frame)), same origins
    r2 = o.render(v, w, h, frame_count=1)
    assert (r2["bounce_rays"]["o"][hit] == br["o"][hit]).all() and (r2["bounce_rays"]["d"][hit] != d).any()
    img = r["rgba"].reshape(h, w, 4)
    assert (img[..., 3] == 255).all() and img[..., 0].max() > 100 and (img[..., 0][~hit.reshape(h, w)] == 0).all()


def test_empty_inputs():
    o = ob.Oracle(np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    assert len(o.trace(np.zeros(0, dtype=ob.RAY_DTYPE))) == 0
    h = o.trace(random_rays(10, 1))
    assert (h["prim"] == ob.INVALID_PRIM).all() and np.isinf(h["t"]).all()


def test_step_log_matches_the_visit_counters(cornell):
    """orc_trace_oplog (input of tests/tools/sched_sim.py): one byte per step, 'N' per node fetched and 'T' per
    triangle tested, in the ray's own order; every ray starts with the root node."""
    import ctypes as C
    p = host.PackedScene(cornell)
    orc = ob.Oracle.from_packed(p)
    rays = random_rays(3000, 11)
    hits, cnt, tot = orc.trace(rays, counts=True)
    n_ops = cnt["nodes"].astype(np.uint64) + cnt["tris"] + cnt["insts"]
    offsets = np.zeros(rays.shape[0] + 1, dtype=np.uint64)
    np.cumsum(n_ops, out=offsets[1:])
    ops = np.zeros(int(offsets[-1]) + 1, dtype=np.uint8)
    L = ob.lib()
    L.orc_trace_oplog.restype = C.c_int
    L.orc_trace_oplog.argtypes = [C.POINTER(ob.OrcScene), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    assert L.orc_trace_oplog(C.byref(orc.scene), rays.ctypes.data, rays.shape[0], cnt.ctypes.data, ops.ctypes.data, offsets.ctypes.data, 0) == 0
    assert (ops[:-1] == ord("N")).sum() == tot["nodes"] and (ops[:-1] == ord("T")).sum() == tot["tris"]
    assert (ops[offsets[:-1].astype(np.int64)] == ord("N")).all()


def test_simd_node_test_is_bit_identical_to_scalar(cornell):
    """The oracle's AVX2 node test (8 children per vector) against its scalar statement of query.hlsl:213-303: same hit
    masks on random nodes x rays, same hits AND same node / triangle counts on whole traversals — including rays with
    denormal direction components, whose non-finite per-node constants take the scalar fallback."""
    if not ob.lib().orc_simd():
        ob.set_simd(True)
        if not ob.simd():
            pytest.skip("no AVX2 on this CPU: the oracle is scalar only")
    try:
        p = host.PackedScene(cornell)
        nodes = p.bvh_bytes.reshape(-1, 80)
        rays = random_rays(4000, seed=21)
        rays["d"][:40, 0] = np.float32(1e-42)                       # denormal: 1/d overflows to inf
        rays["d"][40:80, 1] = np.float32(-3e-45)
        rng = np.random.default_rng(2)
        pick = rng.integers(0, nodes.shape[0], size=4000)
        tmax = rng.uniform(0.0, 5.0, size=4000).astype(np.float32)
        # aim most probe rays at the middle of the node they are tested against, so that children are really hit
        aimed = rays.copy()
        org = nodes[pick, 0:12].copy().view(np.float32)
        ext = np.ldexp(np.float32(128.0), nodes[pick, 12:15].astype(np.int32) - 127).astype(np.float32)
        d = (org + ext) - aimed["o"]
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        aimed["d"][80:] = d[80:].astype(np.float32)
        masks = {}
        for simd in (False, True):
            ob.set_simd(simd)
            masks[simd] = np.array([ob.node_intersect(nodes[pick[i]], aimed[i], tmax[i]) for i in range(4000)], dtype=np.uint32)
        assert np.array_equal(masks[False], masks[True]) and (masks[True] != 0).sum() > 500
        out = {}
        for use_tlas in (False, True):
            q = host.PackedScene(cornell, use_tlas=use_tlas)
            orc = ob.Oracle.from_packed(q)
            for simd in (False, True):
                ob.set_simd(simd)
                out[simd] = orc.trace(rays, counts=True)
            assert np.array_equal(out[False][0].view(np.uint8), out[True][0].view(np.uint8))
            assert np.array_equal(out[False][1].view(np.uint8), out[True][1].view(np.uint8))
    finally:
        ob.set_simd(True)


def test_oracle_matches_a_line_by_line_transliteration_of_the_shader(cornell, box):
    """tests/hlsl_transliteration.py restates rt_gpu_software_query.hlsl:89-438 in Python, written separately from oracle/.
    With the two documented CPU-vs-HLSL switches set to the shader's side (box test divides, an equal t replaces), the C
    oracle must return its primitive, its t bit for bit, and its node / triangle visit counts on every ray."""
    import hlsl_transliteration as hl
    ob.set_variant(ob.VARIANT_BOX_DIVIDE | ob.VARIANT_TIE_LAST)
    try:
        for mesh, n, seed in ((cornell, 260, 41), (box, 120, 42)):
            p = host.PackedScene(mesh)
            orc = ob.Oracle.from_packed(p)
            rays = random_rays(n, seed, axis_fraction=0.15, bounded_fraction=0.0)
            hits, cnt, _ = orc.trace(rays, counts=True)
            sc = hl.Scene(p.bvh_bytes, p.tri_bytes, p.tri_stride)
            n_hit = 0
            for i in range(n):
                t, prim, nn, nt = hl.traverse_bvh(sc, rays["o"][i], rays["d"][i])
                if prim < 0:
                    assert hits["prim"][i] == ob.INVALID_PRIM, i
                else:
                    n_hit += 1
                    assert hits["prim"][i] == prim, (i, hits["prim"][i], prim)
                    assert np.float32(hits["t"][i]).view(np.uint32) == np.float32(t).view(np.uint32), (i, hits["t"][i], t)
                assert cnt["nodes"][i] == nn and cnt["tris"][i] == nt, (i, cnt[i], nn, nt)
            assert n_hit > n // 4
        # two-level traversal (rt_gpu_software_query_tlas.hlsl:333-500) on the buffers cwbvh_gpu_runner's layout prescribes
        for mesh, n, seed in ((cornell, 200, 43), (box, 100, 44)):
            p = host.PackedScene(mesh, use_tlas=True)
            orc = ob.Oracle.from_packed(p)
            rays = random_rays(n, seed, axis_fraction=0.15, bounded_fraction=0.0)
            hits, cnt, _ = orc.trace(rays, counts=True)
            sc = hl.Scene(p.bvh_bytes, p.tri_bytes, p.tri_stride, p.blas_offsets, p.tlas_start)
            for i in range(n):
                t, prim, nn, nt, ni = hl.traverse_bvh_tlas(sc, rays["o"][i], rays["d"][i])
                if prim < 0:
                    assert hits["prim"][i] == ob.INVALID_PRIM, i
                else:
                    assert hits["prim"][i] == prim, (i, hits["prim"][i], prim)
                    assert np.float32(hits["t"][i]).view(np.uint32) == np.float32(t).view(np.uint32), (i, hits["t"][i], t)
                assert cnt["nodes"][i] == nn and cnt["tris"][i] == nt and cnt["insts"][i] == ni, (i, cnt[i], nn, nt, ni)
    finally:
        ob.set_variant(0)


def test_new_variant_switches_touch_only_what_they_name(cornell):
    """ORC_VARIANT_BOX_TMIN_RAY: with the slab test clamped at ray.tmin = 0 instead of 1e-4 a superset of the boxes is entered;
    (prim, t) can only change for rays that start closer than 1e-4 to a surface — the default never enters a box that ends before
    t = 1e-4, so it cannot see a hit that close (the bounce rays of rt_cpu.rs:67 start 0.01 off the surface for that reason).
    ORC_VARIANT_ZERODIR_BOX_ONLY: only rays with an exact-zero direction component can change; those that do keep their
    primitive, with t moved by the 1.19e-7 tilt of the patched component — up to ~1e-5 relative on a surface the ray grazes,
    i.e. AT the north_star tolerance, which is why this switch is the one the census (scripts/variant_census.py) watches."""
    p = host.PackedScene(cornell)
    o = ob.Oracle.from_packed(p)
    rays = random_rays(60000, 77, axis_fraction=0.2, bounded_fraction=0.0)
    a, ca, _ = o.trace(rays, counts=True)
    try:
        ob.set_variant(ob.VARIANT_BOX_TMIN_RAY)
        b, cb, _ = o.trace(rays, counts=True)
        diff = (a["prim"] != b["prim"]) | (a["t"].view(np.uint32) != b["t"].view(np.uint32))
        assert diff.sum() <= 5 and (b["t"][diff] < 1.0001e-4).all()        # only hits the default cannot see: t < 1e-4
        assert int(cb["nodes"].sum()) >= int(ca["nodes"].sum())            # clamp at 0 <= 1e-4: a superset of the boxes
        ob.set_variant(ob.VARIANT_ZERODIR_BOX_ONLY)
        c = o.trace(rays)
    finally:
        ob.set_variant(0)
    zero = (rays["d"] == 0).any(axis=1)
    changed = (a["prim"] != c["prim"]) | (a["t"].view(np.uint32) != c["t"].view(np.uint32))
    assert zero.sum() >= 10000 and not changed[~zero].any()
    assert changed.any()                                                   # the switch is not a no-op
    both = changed & (a["prim"] != ob.INVALID_PRIM) & (c["prim"] != ob.INVALID_PRIM)
    rel = np.abs(a["t"][both] - c["t"][both]) / np.abs(a["t"][both])
    assert (rel <= 1e-4).all() and (rel <= 1e-5).mean() > 0.9
    assert (a["prim"][both] == c["prim"][both]).mean() > 0.99


def test_child_order_table_equals_the_octant_permutation():
    """The lane kernel keeps hit inner children in SLOT space and picks the next one through child_order[oct_inv][hit byte]
    (traverse.cuh, node test in slot space).  The reference permutes instead: child `slot` sets bit slot ^ oct_inv, firstbithigh
    picks the highest bit, `slot = bit ^ oct_inv` undoes it (query.hlsl:256-262, 358, 370).  For every octant and every hit
    byte both give the same child, and the remaining set keeps giving the same order."""
    for oi in range(8):
        for t in range(1, 256):
            order_ref, order_tab = [], []
            perm = 0
            for j in range(8):
                if (t >> j) & 1:
                    perm |= 1 << (j ^ oi)
            while perm:
                off = perm.bit_length() - 1                  # firstbithigh
                perm &= ~(1 << off)
                order_ref.append(off ^ oi)                   # slot_index
            rest = t
            while rest:
                best = max((j for j in range(8) if (rest >> j) & 1), key=lambda j: j ^ oi)    # child_order_fill
                rest &= ~(1 << best)
                order_tab.append(best)
            assert order_ref == order_tab, (oi, t)

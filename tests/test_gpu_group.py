"""GPU tests of the boundary rows VERDICT r1 marked partial / missing: CPU-style TLAS hit records (geometry_id, primitive_id),
two frames in flight, and the one-process multi-GPU entry points (tray_group, tray_cuda_start_multi).  The group tests run
with however many devices the box has (1 on the driver's test box, up to 8 under SCALE)."""
import numpy as np
import pytest

import oracle_binding as ob
from conftest import random_rays
from tray_racing_b200 import cuda, host

pytestmark = pytest.mark.gpu
FLAGS = cuda.RENDER_BOUNCE | cuda.RENDER_RGBA


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_hits_to_geometry_matches_cwbvh_tlas_scene_semantics(cornell):
    """`CwBvhTlasScene::traverse` (src/cwbvh.rs:144-166): geometry_id = BLAS index, primitive_id = index into that BLAS's own
    permuted triangle array; the GPU buffers carry one global index (src/rt_gpu/mod.rs:45-47).  5 objects in cornell_box.obj."""
    p = host.PackedScene(cornell, use_tlas=True)
    assert p.blas_tri_offsets.size == 6
    sc = cuda.TrayCudaScene.from_packed(p)
    try:
        rays = random_rays(50001, 21)
        hits = sc.traverse(rays)
        g0, p0 = sc.hits_to_geometry(hits)                               # no table yet: flat-scene convention
        assert (g0 == 0xFFFFFFFF).all() and (p0 == hits["prim"]).all()
        sc.set_geometry_offsets(p.blas_tri_offsets)
        g, pr = sc.hits_to_geometry(hits)
        hit = hits["prim"] != ob.INVALID_PRIM
        assert hit.sum() > 10000 and (~hit).sum() > 100
        wg, wp = p.geometry_of(hits["prim"][hit])
        assert (g[hit] == wg).all() and (pr[hit] == wp).all()
        assert (g[~hit] == 0xFFFFFFFF).all() and (pr[~hit] == 0xFFFFFFFF).all()   # RayHit::none()
        assert set(np.unique(g[hit])) <= set(range(5)) and len(np.unique(g[hit])) >= 4
        # the pair addresses the triangle the global index addresses: meshes[geometry_id][primitive_id]
        assert (p.blas_tri_offsets[g[hit]] + pr[hit] == hits["prim"][hit]).all()
        assert len(sc.hits_to_geometry(hits[:0])[0]) == 0
        with pytest.raises(cuda.TrayCudaError, match="run from 0 to n_tris"):
            sc.set_geometry_offsets(np.array([0, 5], dtype=np.uint32))
        with pytest.raises(cuda.TrayCudaError, match="ascend"):
            sc.set_geometry_offsets(np.array([0, 9, 5, p.n_tris], dtype=np.uint32))
    finally:
        sc.close()


@pytest.mark.parametrize("n_in_flight", [2, 3])
@pytest.mark.parametrize("overlap", [False, True])
def test_two_frames_in_flight_are_the_same_frames(cornell, overlap, n_in_flight):
    """tray_cuda_scene_set_frames_in_flight(2): consecutive frames alternate between two buffer sets / streams.  Every frame is
    bit-identical to the one-at-a-time frame; frame_count (--animate) changes the bounce rays, so a mix-up of slots would show."""
    p = host.PackedScene(cornell)
    w, h = 320, 184
    view = host.view_from_camera(cornell.camera, w, h)
    orc = ob.Oracle.from_packed(p)
    refs = [orc.render(view, w, h, fc, rgba=True) for fc in range(4)]
    sc = cuda.TrayCudaScene.from_packed(p)
    fl = FLAGS | (cuda.RENDER_OVERLAP if overlap else 0)
    try:
        sc.set_frames_in_flight(n_in_flight)
        for fc in range(4):                                              # download right after each render: "the last frame"
            sc.render(view, w, h, fc, fl, timed=False)
            out = sc.download(primary=True, bounce=True)
            for k in ("primary", "bounce"):
                assert (out[k]["prim"] == refs[fc][k]["prim"]).all() and (bits(out[k]["t"]) == bits(refs[fc][k]["t"])).all(), (fc, k)
        # four frames enqueued back to back, read back asynchronously two at a time
        frames = [np.zeros((h, w, 4), dtype=np.uint8) for _ in range(4)]
        for fc in range(4):
            sc.render(view, w, h, fc, fl, timed=False)
            sc.readback_begin(frames[fc], fc & 1)
            if fc >= 1:
                sc.readback_wait((fc - 1) & 1)
        sc.readback_wait(1)
        sc.sync()
        for fc in range(4):
            assert np.abs(frames[fc].reshape(-1, 4).astype(int) - refs[fc]["rgba"].astype(int)).max() <= 1, fc
        assert (frames[0] != frames[1]).any()                            # the frames do differ (hash_noise(px, frame))
        # timed renders and counters work per slot
        a, b = sc.render(view, w, h, 0, fl | cuda.RENDER_COUNTERS)
        cp, cb = sc.counters()
        assert cp["nodes"] == refs[0]["primary_totals"]["nodes"] and cb["tris"] == refs[0]["bounce_totals"]["tris"]
        assert sc.render_frame_ms(view, w, h, 1, fl) > 0
        sc.set_frames_in_flight(1)
        sc.render(view, w, h, 2, fl, timed=False)
        out = sc.download(primary=True, bounce=True)
        assert (out["bounce"]["prim"] == refs[2]["bounce"]["prim"]).all()
    finally:
        sc.close()


def test_fence_and_after_order_foreign_streams(cornell):
    """tray_cuda_scene_fence / _after: a caller-owned stream is ordered against frames that run on the scene's two streams."""
    import torch
    p = host.PackedScene(cornell)
    w, h = 320, 184
    view = host.view_from_camera(cornell.camera, w, h)
    sc = cuda.TrayCudaScene.from_packed(p)
    try:
        sc.set_frames_in_flight(2)
        st = torch.cuda.Stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            sc.after(st.cuda_stream)
            for fc in range(6):
                sc.render(view, w, h, fc, FLAGS, timed=False)
            sc.fence(st.cuda_stream)
            e1.record(st)
        e1.synchronize()
        assert e0.elapsed_time(e1) > 0.05                                # six frames of work lie between the two events
        out = sc.download(primary=True)
        assert (out["primary"]["prim"] == ob.Oracle.from_packed(p).render(view, w, h, 5)["primary"]["prim"]).all()
    finally:
        sc.close()


def _devices():
    return list(range(min(cuda.device_count(), 8)))


@pytest.mark.parametrize("push", [True, False])
@pytest.mark.parametrize("in_flight", [1, 2, 3])
def test_group_frame_equals_single_gpu_frame(in_flight, push):
    """One process, all the box's GPUs: tiles dealt round-robin, pixels stored into ONE frame on devices[0] over peer access,
    completion by events.  The frame equals the single-scene frame byte for byte; per-device shards hold the oracle's hits."""
    m = host.Mesh.generate("kitchen", 1, 1.0)
    p = host.PackedScene(m)
    w, h = 640, 368
    view = host.view_from_camera(m.camera, w, h)
    one = cuda.TrayCudaScene.from_packed(p)
    try:
        one.render(view, w, h, 0, FLAGS)
        want0 = one.download(rgba=True)["rgba"].copy()
        one.render(view, w, h, 7, FLAGS)
        want7 = one.download(rgba=True)["rgba"].copy()
    finally:
        one.close()
    devs = _devices()
    g = cuda.TrayCudaGroup.from_packed(p, devices=devs)
    try:
        g.set_frames_in_flight(in_flight)
        g.set_exchange(push)            # DMA push of the compact shards + untile on devices[0], or pixel stores into the frame
        for rep in range(3):
            g.render(view, w, h, 0, FLAGS)
            assert (g.frame() == want0).all(), rep
            g.render(view, w, h, 7, FLAGS)
            assert (g.frame() == want7).all(), rep
        # pipelined: frames back to back, read back asynchronously, alternating targets
        frames = [np.zeros((h, w, 4), dtype=np.uint8) for _ in range(6)]
        for i in range(6):
            g.render(view, w, h, 7 if i & 1 else 0, FLAGS)
            g.readback_begin(frames[i], i & 1)
            if i >= 1:
                g.readback_wait((i - 1) & 1)
        g.readback_wait(1)
        g.sync()
        for i in range(6):
            assert (frames[i] == (want7 if i & 1 else want0)).all(), i
        ms = g.render(view, w, h, 0, FLAGS, timed=True)
        assert ms > 0
        # hits of every device's shard against the oracle
        ref = ob.Oracle.from_packed(p).render(view, w, h, 0)
        acc = {}
        for i in range(len(devs)):
            sc = g.scene(i)
            sc.frame_size = (w, h)
            sc.download(primary=True, bounce=True, into=acc, merge=True)
        for k in ("primary", "bounce"):
            assert (acc[k]["prim"] == ref[k]["prim"]).all() and (bits(acc[k]["t"]) == bits(ref[k]["t"])).all(), k
    finally:
        g.close()


def test_start_multi_runs_the_reference_protocol(cornell):
    """tray_cuda_start_multi = rt_gpu_software::start on the box's GPUs: min <= mean, frames counted, --tlas buffers accepted."""
    p = host.PackedScene(cornell, use_tlas=True)
    w, h = 640, 360
    view = host.view_from_camera(cornell.camera, w, h, p.tlas_start)
    devs = _devices()
    mn, mean, frames = cuda.start_multi(devs, p.bvh_bytes, p.instance_bytes, p.tri_bytes, p.tlas_start, view, w, h,
                                        render_time=0.2, benchmark=True, animate=True, use_tlas=True, tri_stride=p.tri_stride)
    assert frames >= 2 and 0 < mn <= mean
    with pytest.raises(cuda.TrayCudaError, match="multiple of 80"):
        cuda.start_multi(devs, p.bvh_bytes[:-1], p.instance_bytes, p.tri_bytes, p.tlas_start, view, w, h, render_time=0.05)
    with pytest.raises(cuda.TrayCudaError, match="out of range"):
        cuda.start_multi(list(range(cuda.device_count() + 1)), p.bvh_bytes, p.instance_bytes, p.tri_bytes, p.tlas_start, view, w, h, render_time=0.05)


def test_completion_flags_are_stream_memory_operations(cornell):
    """tray_cuda_frame_signal / _wait_flag: a 32-bit flag written and awaited by the stream front-end (no kernel, no SM slot) —
    the completion primitive of the multi-process peer exchange.  Here on one device: order, values, wrap-around compare."""
    import torch
    p = host.PackedScene(cornell)
    w, h = 320, 184
    view = host.view_from_camera(cornell.camera, w, h)
    sc = cuda.TrayCudaScene.from_packed(p)
    flag = cuda.frame_alloc(64)
    try:
        words = torch.as_tensor(cuda.DeviceArray(flag, (16,), "<u4", sc), device="cuda")
        sc.render(view, w, h, 0, FLAGS, timed=False)
        sc.signal(flag, 7)                          # behind the frame's kernels
        sc.wait_flag(flag, 7)                       # passes: the write is ahead of it on the same stream
        sc.wait_flag(flag, 5)                       # GEQ
        sc.signal(flag + 4, 0xFFFFFFF0)
        sc.wait_flag(flag + 4, 0xFFFFFFE0)          # (int32)(flag - value) >= 0 across the wrap
        sc.sync()
        assert int(words[0].item()) == 7 and int(words[1].item()) == 0xFFFFFFF0
        # the next frame (other slot, other stream) only starts once the last frame has signalled
        sc.set_frames_in_flight(2)
        for k in range(8, 14):
            sc.render(view, w, h, k, FLAGS, timed=False)
            sc.signal(flag + 8, k)
            sc.wait_flag(flag + 8, k, before_next_frame=True)
        sc.sync()
        assert int(words[2].item()) == 13
        out = sc.download(primary=True)
        assert (out["primary"]["prim"] == ob.Oracle.from_packed(p).render(view, w, h, 13)["primary"]["prim"]).all()
    finally:
        sc.close()
        cuda.frame_free(flag)


@pytest.mark.parametrize("w,h,shards", [(640, 368, 3), (101, 37, 2), (1920, 1080, 8)])
def test_push_and_untile_shards_assemble_the_frame(cornell, w, h, shards):
    """tray_cuda_frame_push + tray_cuda_untile_shards: every shard's compact RGBA moved by one DMA copy into a staging array, then
    one launch writes the row-major frame — the same bytes as the single-shard frame (ragged sizes included)."""
    import torch
    p = host.PackedScene(cornell)
    view = host.view_from_camera(cornell.camera, w, h)
    sc = cuda.TrayCudaScene.from_packed(p)
    items = cuda.local_items(w, h, 0, shards)
    assert items == cuda.lib().tray_cuda_shard_items(w, h, 0, shards)
    staging = cuda.frame_alloc(shards * items * 4 + w * h * 4)
    try:
        sc.render(view, w, h, 3, FLAGS)
        want = sc.download(rgba=True)["rgba"].copy()
        for s in range(shards):
            sc.render(view, w, h, 3, FLAGS, shard=s, shards=shards, timed=False)
            sc.push(staging + s * items * 4)
        sc.untile_shards(staging, w, h, shards, staging + shards * items * 4)
        sc.sync()
        got = torch.as_tensor(cuda.DeviceArray(staging + shards * items * 4, (h, w, 4), "|u1", sc), device="cuda").cpu().numpy()
        assert (got == want).all()
    finally:
        sc.close()
        cuda.frame_free(staging)

"""binary scene / ray / hit dumps.  The oracle stands in for rt_cpu as the writer of hits.bin
(the real writer is the Rust snippet of INTEGRATION.md §5); the product-side reader / scorer never sees the oracle."""
import json
import subprocess
import sys

import numpy as np
import pytest

import oracle_binding as ob
from conftest import ROOT, random_rays
from tray_racing_b200 import dump, host


def make_dump(tmp_path, mesh, use_tlas=False, stride=48, n=20000, seed=7):
    p = host.PackedScene(mesh, use_tlas=use_tlas, tri_stride=stride)
    rays = random_rays(n, seed)
    hits = ob.Oracle.from_packed(p).trace(rays)
    d = str(tmp_path / f"dump_{int(use_tlas)}_{stride}")
    dump.write_dump(d, p.bvh_bytes, p.tri_bytes, stride, p.tlas_start, use_tlas, prim_indices=p.prim_to_mesh_tri,
                    blas_offsets=p.blas_offsets if use_tlas else None, rays=rays, hits=hits, note="test")
    return d, p, rays, hits


@pytest.mark.parametrize("use_tlas,stride", [(False, 48), (True, 64)])
def test_roundtrip_is_byte_exact(tmp_path, cornell, use_tlas, stride):
    d, p, rays, hits = make_dump(tmp_path, cornell, use_tlas, stride, n=2000)
    r = dump.read_dump(d)
    assert r.tri_stride == stride and r.use_tlas == use_tlas and r.tlas_start == p.tlas_start
    assert (r.bvh_bytes == p.bvh_bytes).all() and (r.tri_bytes == p.tri_bytes).all()
    assert r.n_nodes == p.n_nodes and r.n_tris == p.n_tris
    assert (r.prim_indices == p.prim_to_mesh_tri).all()
    assert (r.rays.view(np.uint8) == rays.view(np.uint8)).all() and (r.hits.view(np.uint8) == hits.view(np.uint8)).all()
    assert (r.blas_offsets == p.blas_offsets).all() if use_tlas else r.blas_offsets is None
    # the reloaded buffers drive the oracle to the same answers
    again = ob.Oracle(r.bvh_bytes, r.tri_bytes, r.tri_stride, r.blas_offsets, r.tlas_start, r.use_tlas).trace(r.rays)
    assert (again.view(np.uint8) == hits.view(np.uint8)).all()
    out = subprocess.run([sys.executable, "-m", "tray_racing_b200.dump", "info", d], capture_output=True, text=True, cwd=ROOT)
    assert json.loads(out.stdout)["n_nodes"] == p.n_nodes


def test_reader_rejects_bad_shapes(tmp_path, box):
    d, p, rays, hits = make_dump(tmp_path, box, n=64)
    with open(f"{d}/nodes.bin", "ab") as f:
        f.write(b"\0" * 7)
    with pytest.raises(ValueError, match="multiple of 80"):
        dump.read_dump(d)
    dump.write_dump(d, p.bvh_bytes, p.tri_bytes[:-3], 48, rays=rays, hits=hits)
    with pytest.raises(ValueError, match="tri_stride"):
        dump.read_dump(d)
    dump.write_dump(d, p.bvh_bytes, p.tri_bytes, 48, rays=rays, hits=hits[:-1])
    with pytest.raises(ValueError, match="hits"):
        dump.read_dump(d)
    with pytest.raises(FileNotFoundError):
        dump.read_dump(str(tmp_path / "nowhere"))


def test_score_implements_the_north_star_bar():
    want = np.zeros(6, dtype=dump.HIT_DTYPE)
    want["t"] = [1.0, 2.0, 3.0, np.inf, 3.402823466e+38, 5.0]
    want["prim"] = [10, 11, 12, dump.INVALID_PRIM, dump.INVALID_PRIM, 15]
    got = want.copy()
    got["t"][4] = np.inf                                  # both spellings of RayHit::none() are a miss
    assert dump.score(got, want)["pass"]
    got["t"][0] = np.float32(1.0 + 5e-6)                  # inside 1e-5 relative
    got["prim"][1] = 99                                   # same t, other triangle: a genuine tie
    rep = dump.score(got, want)
    assert rep["pass"] and rep["prim_mismatch"] == 1 and rep["prim_mismatch_genuine_ties"] == 1 and rep["t_bit_identical"] == 3
    bad = got.copy(); bad["t"][2] = np.float32(3.001)     # t off by 3e-4 relative
    assert not dump.score(bad, want)["pass"]
    bad = got.copy(); bad["prim"][3] = 5; bad["t"][3] = 1.0   # hit where the reference missed
    assert dump.score(bad, want)["hit_miss_disagree"] == 1
    bad = got.copy(); bad["prim"][5] = 16; bad["t"][5] = np.float32(5.01)   # other triangle AND other t: not a tie
    rep = dump.score(bad, want)
    assert not rep["pass"] and rep["prim_mismatch"] == 2 and rep["prim_mismatch_genuine_ties"] == 1


@pytest.mark.gpu
@pytest.mark.parametrize("use_tlas,stride", [(False, 48), (True, 48), (False, 64)])
def test_check_dump_on_gpu(tmp_path, cornell, use_tlas, stride):
    d, p, rays, hits = make_dump(tmp_path, cornell, use_tlas, stride, n=100000)
    rep = dump.check(d)
    assert rep["pass"] and rep["hits"] > 10000
    assert rep["t_bit_identical"] == rep["hits"] and rep["prim_mismatch"] == 0     # the stronger bar this kernel meets
    out = subprocess.run([sys.executable, "-m", "tray_racing_b200.dump", "check", d], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and json.loads(out.stdout)["pass"]
    # a corrupted expectation is caught
    hits2 = hits.copy(); k = np.nonzero(hits["prim"] != dump.INVALID_PRIM)[0][0]; hits2["t"][k] *= np.float32(1.01)
    dump.write_dump(d, p.bvh_bytes, p.tri_bytes, stride, p.tlas_start, use_tlas, blas_offsets=p.blas_offsets if use_tlas else None,
                    rays=rays, hits=hits2)
    assert not dump.check(d)["pass"]

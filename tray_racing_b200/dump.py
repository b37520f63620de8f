"""Binary scene / ray / hit dumps: the bridge that lets REAL obvhs output reach this backend
without a Rust toolchain on this side.

A dump directory holds little-endian raw arrays, exactly the bytes the reference already has in memory:

  nodes.bin         n_nodes x 80 B   `bvh.nodes` as `cwbvh_gpu_runner` casts them (reference src/rt_gpu/mod.rs:62-69,101):
                                     flat = one BLAS with the root at 0; --tlas = BLAS0 | BLAS1 | ... | TLAS
  tris.bin          n_tris x stride  BVH-ordered triangles (reference src/rt_cpu/mod.rs:38-43): {v0,e1,e2} 48 B or
                                     {v0,e1,e2,ng} 64 B per `RtTriangle`, each vec3 padded to 16 B
  prim_indices.bin  n_tris x u32     optional: `bvh.primitive_indices` (BVH slot -> original triangle, src/cwbvh.rs:185-187)
  blas_offsets.bin  n_inst x u32     --tlas only: node offset of the BLAS behind each TLAS leaf (src/rt_gpu/mod.rs:72-78)
  rays.bin          n x 32 B         optional: `Ray::new(origin, dir, tmin, tmax)` as [o.xyz, tmin, d.xyz, tmax]
  hits.bin          n x 8 B          optional: rt_cpu's answer for those rays as [t: f32, primitive_id: u32]
                                     (`RayHit::none()` = t +inf / f32::MAX and id u32::MAX)
  meta.json                          {"tri_stride", "tlas_start", "use_tlas", optional "width", "height", "note"}

INTEGRATION.md §5 has the ~20 lines of Rust that write these from `cwbvh_cpu_runner`.  `check()` runs the rays of a
dump through the CUDA path and scores them against hits.bin with BASELINE.json's bar: primitive ids identical except
genuine ties, t within 1e-5 relative.  Nothing here touches the CPU oracle.

  python -m tray_racing_b200.dump check DIR [--device 0]
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("prim", "<u4")])
INVALID_PRIM = 0xFFFFFFFF
F32_MAX = np.float32(3.402823466e+38)


@dataclass
class Dump:
    bvh_bytes: np.ndarray
    tri_bytes: np.ndarray
    tri_stride: int = 48
    tlas_start: int = 0
    use_tlas: bool = False
    prim_indices: np.ndarray | None = None
    blas_offsets: np.ndarray | None = None
    rays: np.ndarray | None = None
    hits: np.ndarray | None = None
    meta: dict | None = None

    @property
    def n_nodes(self) -> int:
        return self.bvh_bytes.size // 80

    @property
    def n_tris(self) -> int:
        return self.tri_bytes.size // self.tri_stride


def write_dump(path: str, bvh_bytes, tri_bytes, tri_stride: int = 48, tlas_start: int = 0, use_tlas: bool = False,
               prim_indices=None, blas_offsets=None, rays=None, hits=None, **meta) -> None:
    os.makedirs(path, exist_ok=True)

    def put(name, arr, dtype):
        if arr is not None:
            np.ascontiguousarray(arr, dtype=dtype).tofile(os.path.join(path, name))

    put("nodes.bin", bvh_bytes, np.uint8)
    put("tris.bin", tri_bytes, np.uint8)
    put("prim_indices.bin", prim_indices, "<u4")
    put("blas_offsets.bin", blas_offsets if use_tlas else None, "<u4")
    put("rays.bin", rays, RAY_DTYPE)
    put("hits.bin", hits, HIT_DTYPE)
    m = dict(meta, tri_stride=int(tri_stride), tlas_start=int(tlas_start), use_tlas=bool(use_tlas))
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump(m, f, indent=1, sort_keys=True)


def read_dump(path: str) -> Dump:
    def get(name, dtype):
        p = os.path.join(path, name)
        return np.fromfile(p, dtype=dtype) if os.path.exists(p) else None

    meta = {}
    mp = os.path.join(path, "meta.json")
    if os.path.exists(mp):
        with open(mp) as f:
            meta = json.load(f)
    nodes, tris = get("nodes.bin", np.uint8), get("tris.bin", np.uint8)
    if nodes is None or tris is None:
        raise FileNotFoundError(f"{path}: nodes.bin and tris.bin are required")
    stride = int(meta.get("tri_stride", 48))
    if nodes.size % 80:
        raise ValueError(f"nodes.bin: {nodes.size} bytes is not a multiple of 80")        # reference src/rt_gpu/mod.rs:70,105
    if stride not in (48, 64) or tris.size % stride:
        raise ValueError(f"tris.bin: {tris.size} bytes is not a multiple of tri_stride {stride}")
    blas = get("blas_offsets.bin", "<u4")
    use_tlas = bool(meta.get("use_tlas", blas is not None))
    if use_tlas and blas is None:
        raise FileNotFoundError(f"{path}: --tlas dump without blas_offsets.bin")
    rays, hits = get("rays.bin", RAY_DTYPE), get("hits.bin", HIT_DTYPE)
    if rays is not None and hits is not None and rays.shape[0] != hits.shape[0]:
        raise ValueError(f"rays.bin holds {rays.shape[0]} rays but hits.bin {hits.shape[0]} hits")
    pi = get("prim_indices.bin", "<u4")
    if pi is not None and pi.size != tris.size // stride:
        raise ValueError("prim_indices.bin does not match tris.bin")
    return Dump(nodes, tris, stride, int(meta.get("tlas_start", 0)), use_tlas, pi, blas, rays, hits, meta)


def score(got: np.ndarray, want: np.ndarray, tie_checker=None, rtol: float = 1e-5) -> dict:
    """BASELINE.json's bar.  `want` may encode a miss as t = +inf or t >= f32::MAX (rt_cpu.rs:61), id u32::MAX.
    A primitive-id mismatch counts as a genuine tie when both ids give the same t (|dt| <= rtol * |t|)."""
    miss_w = (want["prim"] == INVALID_PRIM) | ~(want["t"] < F32_MAX)
    miss_g = got["prim"] == INVALID_PRIM
    hit_both = ~miss_w & ~miss_g
    dt = np.zeros(got.shape[0], dtype=np.float64)
    dt[hit_both] = np.abs(got["t"][hit_both].astype(np.float64) - want["t"][hit_both]) / np.maximum(np.abs(want["t"][hit_both]), 1e-30)
    t_bad = hit_both & (dt > rtol)
    id_diff = hit_both & (got["prim"] != want["prim"])
    ties = id_diff & ~t_bad                     # same t, other triangle: a genuine tie
    if tie_checker is not None and ties.any():
        idx = np.nonzero(ties)[0]
        ties[idx] = tie_checker(idx)
    rep = {
        "rays": int(got.shape[0]),
        "hit_miss_disagree": int((miss_w != miss_g).sum()),
        "t_beyond_tolerance": int(t_bad.sum()),
        "t_bit_identical": int((hit_both & (got["t"].view(np.uint32) == want["t"].view(np.uint32))).sum()),
        "hits": int(hit_both.sum()),
        "prim_mismatch": int(id_diff.sum()),
        "prim_mismatch_genuine_ties": int(ties.sum()),
        "max_rel_dt": float(dt.max()) if dt.size else 0.0,
    }
    rep["pass"] = rep["hit_miss_disagree"] == 0 and rep["t_beyond_tolerance"] == 0 and rep["prim_mismatch"] == rep["prim_mismatch_genuine_ties"]
    return rep


def check(path: str, device: int = 0, batch: int = 1 << 22) -> dict:
    """Trace the dump's rays on the GPU through the C ABI and score them against its hits."""
    from . import cuda
    d = read_dump(path)
    if d.rays is None or d.hits is None:
        raise FileNotFoundError(f"{path}: rays.bin and hits.bin are needed for a check")
    sc = cuda.TrayCudaScene(d.bvh_bytes, d.tri_bytes, d.tri_stride, d.blas_offsets if d.use_tlas else None, d.tlas_start, device)
    try:
        got = np.empty(d.rays.shape[0], dtype=HIT_DTYPE)
        for a in range(0, d.rays.shape[0], batch):
            got[a:a + batch] = sc.traverse(d.rays[a:a + batch])
    finally:
        sc.close()
    rep = score(got, d.hits)
    rep["n_nodes"], rep["n_tris"] = d.n_nodes, d.n_tris
    return rep


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("check", help="trace rays.bin on the GPU and compare with hits.bin")
    c.add_argument("dir")
    c.add_argument("--device", type=int, default=0)
    i = sub.add_parser("info", help="print the shapes found in a dump directory")
    i.add_argument("dir")
    a = ap.parse_args(argv)
    if a.cmd == "info":
        d = read_dump(a.dir)
        print(json.dumps({"n_nodes": d.n_nodes, "n_tris": d.n_tris, "tri_stride": d.tri_stride, "use_tlas": d.use_tlas,
                          "tlas_start": d.tlas_start, "n_instances": 0 if d.blas_offsets is None else int(d.blas_offsets.size),
                          "rays": None if d.rays is None else int(d.rays.shape[0]),
                          "hits": None if d.hits is None else int(d.hits.shape[0])}))
        return 0
    rep = check(a.dir, a.device)
    print(json.dumps(rep))
    return 0 if rep["pass"] else 1


if __name__ == "__main__":
    raise SystemExit(main())

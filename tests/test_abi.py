"""CPU tests of the drop-in boundary: every function include/*.h declares is exported by the built library,
POD layouts, argument validation and the no-device error path (no compute call is made without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from tray_racing_b200 import cuda, host


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tray_(?:cuda|host)_[a-z0-9_]+)\s*\(", src)))


def test_cuda_library_exports_every_declared_symbol():
    names = declared_functions("tray_cuda.h")
    assert len(names) >= 16
    L = cuda.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tray_cuda.h but not exported by libtray_cuda.so"
    assert sorted(cuda.EXPORTS) == names
    assert L.tray_cuda_abi_version() == 3


def test_rust_mirror_declares_every_entry_point():
    """tray_cuda/src/lib.rs (the crate the Rust host links) cannot be compiled in this image, so its extern
    block is at least kept complete: every function of include/tray_cuda.h is declared there, with the same arity."""
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "tray_cuda.h")).read(), flags=re.S)
    rust = re.sub(r"//.*", "", open(os.path.join(ROOT, "tray_cuda", "src", "lib.rs")).read())
    ext = rust[rust.index('extern "C" {'):]
    ext = ext[:ext.index("\n}")]
    for n in declared_functions("tray_cuda.h"):
        m = re.search(r"pub fn " + n + r"\s*\((.*?)\)", ext, flags=re.S)
        assert m, f"{n} missing from tray_cuda/src/lib.rs"
        c = re.search(r"\b" + n + r"\s*\((.*?)\)\s*;", header, flags=re.S).group(1).strip()
        n_c = 0 if c in ("", "void") else c.count(",") + 1
        r = m.group(1).strip()
        n_r = 0 if r == "" else r.count(",") + 1
        assert n_c == n_r, f"{n}: {n_c} arguments in the header, {n_r} in the Rust declaration"
    build_rs = open(os.path.join(ROOT, "tray_cuda", "build.rs")).read()
    assert "tray_cuda.cu" in build_rs and "build_gpu.cu" in build_rs and "sm_100a" in build_rs


def test_render_flags_agree_between_header_python_and_rust():
    header = open(os.path.join(ROOT, "include", "tray_cuda.h")).read()
    flags = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define TRAY_RENDER_([A-Z_]+)\s+0x([0-9a-fA-F]+)u", header)}
    assert set(flags) == {"BOUNCE", "RGBA", "COUNTERS", "KEEP_RAYS", "ANYHIT_AO", "OVERLAP"}
    assert len(set(flags.values())) == len(flags)
    rust = open(os.path.join(ROOT, "tray_cuda", "src", "lib.rs")).read()
    for name, value in flags.items():
        assert getattr(cuda, "RENDER_" + name) == value, name
        m = re.search(r"pub const RENDER_" + name + r": u32 = 0x([0-9a-fA-F]+);", rust)
        assert m and int(m.group(1), 16) == value, name


def test_host_library_exports_every_declared_symbol():
    L = host.lib()
    for n in declared_functions("tray_host.h"):
        assert hasattr(L, n), n


def test_cuda_library_is_sm100a_and_has_no_cpu_path():
    """The product library carries sm_100a SASS for the traversal kernel and does not link the oracle."""
    out = subprocess.run(["cuobjdump", "-lelf", cuda.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    deps = subprocess.run(["ldd", cuda.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    for root, _, files in os.walk(os.path.join(ROOT, "tray_racing_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "libtray_oracle" not in text and "oracle_binding" not in text, f


def test_pod_layouts():
    assert host.RAY_DTYPE.itemsize == 32 and host.HIT_DTYPE.itemsize == 8 and host.NODE_BYTES == 80
    assert C.sizeof(host.TrayView) == 160 and C.sizeof(cuda.Counters) == 40
    assert host.RAY_DTYPE.fields["tmin"][1] == 12 and host.RAY_DTYPE.fields["d"][1] == 16 and host.RAY_DTYPE.fields["tmax"][1] == 28


def test_shard_pixels_partition_the_frame():
    for w, h in [(1920, 1080), (3840, 2160), (100, 37), (33, 9), (8, 8)]:
        for shards in (1, 2, 3, 4, 8):
            assert sum(cuda.shard_pixels(w, h, s, shards) for s in range(shards)) == w * h
    # interleaved tiles balance the shards to within one 32x8 tile row
    px = [cuda.shard_pixels(3840, 2160, s, 8) for s in range(8)]
    assert max(px) - min(px) <= 256


def test_argument_validation_without_touching_a_device():
    nodes, tris = np.zeros(80, np.uint8), np.zeros(48, np.uint8)
    with pytest.raises(ValueError):
        cuda.TrayCudaScene(np.zeros(81, np.uint8), tris)                  # reference asserts len % 80, mod.rs:70,105
    with pytest.raises(ValueError):
        cuda.TrayCudaScene(nodes, np.zeros(47, np.uint8))                 # reference asserts tri stride, mod.rs:86,107
    L = cuda.lib()
    h = C.c_void_p()
    assert L.tray_cuda_scene_create(nodes.ctypes.data, 1, tris.ctypes.data, 1, 32, None, 0, 0, 0, C.byref(h)) == -1   # strides: 48, 64 (f32) or 24 (f16)
    assert b"tri_stride" in L.tray_cuda_last_error()
    assert L.tray_cuda_scene_create(None, 1, tris.ctypes.data, 1, 48, None, 0, 0, 0, C.byref(h)) == -1
    assert L.tray_cuda_scene_create(nodes.ctypes.data, 1, tris.ctypes.data, 1, 48, None, 3, 0, 0, C.byref(h)) == -1
    assert L.tray_cuda_render(None, None, 8, 8, 0, 0, 0, 1, None, None) == -1
    assert L.tray_cuda_trace(None, None, 0, None, None, None) == -1


def test_no_device_means_error_not_fallback():
    if cuda.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cuda.TrayCudaError, match="no CUDA device"):
        cuda.TrayCudaScene(np.zeros(80, np.uint8), np.zeros(48, np.uint8))
    v = host.TrayView()
    with pytest.raises(cuda.TrayCudaError):
        cuda.start(np.zeros(80, np.uint8), np.zeros(16, np.uint8), np.zeros(48, np.uint8), 0, v, 64, 64, 0.01)


def test_probes_reject_bad_arguments_and_need_a_device():
    """The roofline probes (bandwidth, L1 gather) validate their arguments before touching a device and report a missing
    device as an error, like every compute entry point."""
    with pytest.raises(cuda.TrayCudaError):
        cuda.l1_gather_probe(nbytes=128 << 10)            # the table must fit the L1: <= 64 KiB
    with pytest.raises(cuda.TrayCudaError):
        cuda.l1_gather_probe(nbytes=32 << 10, iters=0)
    with pytest.raises(cuda.TrayCudaError):
        cuda.bandwidth_probe(16, 1)
    if cuda.device_count() == 0:
        with pytest.raises(cuda.TrayCudaError, match="no CUDA device"):
            cuda.l1_gather_probe()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(cuda, "_lib", None)
    monkeypatch.setattr(cuda, "LIB_PATH", str(tmp_path / "libtray_cuda.so"))
    with pytest.raises(cuda.TrayCudaError, match="no fallback"):
        cuda.lib()

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once (no-op when the .so files are fresh)."""
    import __graft_entry__ as g
    need = [os.path.join(ROOT, "tray_racing_b200", n) for n in ("libtray_cuda.so", "libtray_host.so")]
    need.append(os.path.join(ROOT, "oracle", "libtray_oracle.so"))
    if not all(os.path.exists(p) for p in need):
        g.build()


def load_golden_mesh(name):
    from tray_racing_b200 import host
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
    cam = host.Camera(tuple(float(x) for x in g["eye"]), tuple(float(x) for x in g["look_at"]), float(g["fov"]))
    return host.Mesh.from_tris(g["tris"], g["offsets"], cam)


@pytest.fixture(scope="session")
def cornell():
    return load_golden_mesh("cornell_box")


@pytest.fixture(scope="session")
def box():
    return load_golden_mesh("box")


def random_rays(n, seed, lo=-1.5, hi=1.5, axis_fraction=0.02, bounded_fraction=0.1):
    """Rays towards the scene from a surrounding shell; a few axis-parallel (exact-zero components, the
    zero-direction fix-up of query.hlsl:334) and a few with finite [tmin, tmax]."""
    import oracle_binding as ob
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=ob.RAY_DTYPE)
    o = rng.normal(size=(n, 3)); o /= np.linalg.norm(o, axis=1, keepdims=True)
    o *= rng.uniform(0.2, 3.0, size=(n, 1)) * (hi - lo) / 2
    tgt = rng.uniform(lo, hi, size=(n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = int(n * axis_fraction)
    if k:
        ax = rng.integers(0, 3, size=k)
        d[:k] = 0
        d[np.arange(k), ax] = rng.choice([-1.0, 1.0], size=k)
        o[:k] = tgt[:k] - d[:k] * 4
    rays["o"] = o.astype(np.float32)
    rays["d"] = d.astype(np.float32)
    rays["tmin"] = 0
    rays["tmax"] = np.float32(3.402823466e+38)
    b = int(n * bounded_fraction)
    if b:
        rays["tmin"][-b:] = rng.uniform(0, 1, size=b).astype(np.float32)
        rays["tmax"][-b:] = rays["tmin"][-b:] + rng.uniform(0.1, 3, size=b).astype(np.float32)
    return rays

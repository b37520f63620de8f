"""Host-side mirror of the reference's operator interface for the CWBVH path.

`cwbvh_cuda_runner` is the sibling of `cwbvh_gpu_runner` (reference src/rt_gpu/mod.rs:16-112) and
`cwbvh_cpu_runner` (reference src/rt_cpu/mod.rs:17-74): same inputs (objects, options, scene), same
result (frame time in ms), dispatched from the `--cuda` arm that INTEGRATION.md adds next to
reference src/main.rs:456-470."""
from __future__ import annotations

from dataclasses import dataclass, field

from . import cuda, host


@dataclass
class Options:
    """The fields of the reference's `Options` (src/main.rs:65-171) that reach this path."""
    width: int = 1920
    height: int = 1080
    render_time: float = 1.0
    benchmark: bool = False
    animate: bool = False
    tlas: bool = False
    flatten_blas: bool = False
    max_prims_per_leaf: int = 3
    verbose: bool = False
    build: str = "ploc_cwbvh"
    tri_stride: int = 48
    device: int = 0


@dataclass
class Scene:
    """reference `Scene` (src/main.rs:627-632)"""
    model_path: str = ""
    camera: host.Camera = field(default_factory=lambda: host.Camera((0, 0, 5), (0, 0, 0), 90.0))
    sun_direction: tuple = (0.22, -1.0, -0.2)


@dataclass
class Stats:
    """what a run reports (reference `Stats`, src/main.rs:634-640)"""
    traversal_ms: float
    mean_ms: float
    frames: int
    blas_build_time_s: float
    tlas_build_time_ms: float


def cwbvh_cuda_runner(objects: host.Mesh, options: Options, scene: Scene) -> Stats:
    if "cwbvh" not in options.build:
        raise ValueError("NO BVH BUILDER SPECIFIED")                       # reference src/cwbvh.rs:98-100
    if options.width % 8 or options.height % 8:
        pass  # the reference silently drops the tail (dispatch W/8 x H/8, rt_gpu_software.rs:298); we render it
    use_tlas = options.tlas and not options.flatten_blas                   # reference src/main.rs:300-308
    packed = host.PackedScene(objects, use_tlas=use_tlas, tri_stride=options.tri_stride,
                              max_prims_per_leaf=options.max_prims_per_leaf)
    view = host.view_from_camera(scene.camera, options.width, options.height, packed.tlas_start)
    if options.verbose:
        print(f"{objects.n_objects} objects, triangles {objects.n_tris}, nodes {packed.n_nodes}")
    mn, mean, frames = cuda.start(packed.bvh_bytes, packed.instance_bytes, packed.tri_bytes, packed.tlas_start, view,
                                  options.width, options.height, options.render_time, options.benchmark, options.animate,
                                  use_tlas, options.tri_stride, options.device)
    return Stats(mn, mean, frames, packed.build_seconds, packed.tlas_build_seconds * 1e3)

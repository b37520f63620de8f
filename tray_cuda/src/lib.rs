//! Thin Rust face of `include/tray_cuda.h`.  Mirrors `rt_gpu_software::start` (tray_racing
//! src/rt_gpu/rt_gpu_software.rs:24-32) so that `cwbvh_cuda_runner` reads like `cwbvh_gpu_runner`.
//! Authored against the header; not compiled in the build image (no Rust toolchain there).
use std::ffi::{c_char, c_int, c_void, CStr};

#[repr(C)]
#[derive(Clone, Copy, bytemuck::Pod, bytemuck::Zeroable)]
pub struct TrayRay { pub origin: [f32; 3], pub tmin: f32, pub dir: [f32; 3], pub tmax: f32 }

#[repr(C)]
#[derive(Clone, Copy, bytemuck::Pod, bytemuck::Zeroable)]
pub struct TrayHit { pub t: f32, pub prim: u32 }

/// `ViewUniform` of tray_racing (src/main.rs:589-597) padded to 160 bytes.
#[repr(C)]
#[derive(Clone, Copy, bytemuck::Pod, bytemuck::Zeroable)]
pub struct TrayView {
    pub view_inv: [f32; 16], pub proj_inv: [f32; 16], pub eye: [f32; 3], pub exposure: f32,
    pub tlas_start: u32, pub pad: [u32; 3],
}

/// f32 triangle record: obvhs `RtTriangle` without `ng` (48 B) — `(&tri).into()` then copy v0, e1, e2.
#[repr(C)]
#[derive(Clone, Copy, bytemuck::Pod, bytemuck::Zeroable)]
pub struct TrayTri48 { pub v0: [f32; 3], pub p0: f32, pub e1: [f32; 3], pub p1: f32, pub e2: [f32; 3], pub p2: f32 }

#[repr(C)] pub struct TrayScene { _private: [u8; 0] }

/// `tray_build_stats` of include/tray_cuda.h
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct TrayBuildStats {
    pub n_tris: u64, pub n_nodes: u64, pub ploc_iterations: u32, pub levels: u32,
    pub ms_upload: f32, pub ms_sort: f32, pub ms_ploc: f32, pub ms_collapse: f32, pub ms_total: f32,
}

extern "C" {
    pub fn tray_cuda_device_count() -> c_int;
    pub fn tray_cuda_last_error() -> *const c_char;
    pub fn tray_cuda_scene_create(nodes: *const c_void, n_nodes: u64, tris: *const c_void, n_tris: u64, tri_stride: u32,
        blas_offsets: *const u32, n_instances: u32, tlas_start: u32, device: c_int, out: *mut *mut TrayScene) -> c_int;
    pub fn tray_cuda_scene_destroy(scene: *mut TrayScene);
    pub fn tray_cuda_trace(scene: *mut TrayScene, rays: *const TrayRay, n: u64, hits: *mut TrayHit,
        ms_kernel: *mut f32, ms_total: *mut f32) -> c_int;
    pub fn tray_cuda_trace_any(scene: *mut TrayScene, rays: *const TrayRay, n: u64, hits: *mut TrayHit,
        ms_kernel: *mut f32, ms_total: *mut f32) -> c_int;
    pub fn tray_cuda_render(scene: *mut TrayScene, view: *const TrayView, width: u32, height: u32, frame_count: u32,
        flags: u32, shard_index: u32, shard_count: u32, ms_primary: *mut f32, ms_bounce: *mut f32) -> c_int;
    pub fn tray_cuda_render_timed(scene: *mut TrayScene, view: *const TrayView, width: u32, height: u32, frame_count: u32,
        flags: u32, shard_index: u32, shard_count: u32, ms_frame: *mut f32) -> c_int;
    pub fn tray_cuda_frame_readback_begin(scene: *mut TrayScene, rgba_host: *mut u8, slot: u32) -> c_int;
    pub fn tray_cuda_frame_readback_wait(scene: *mut TrayScene, slot: u32) -> c_int;
    pub fn tray_cuda_scene_build(tris9: *const f32, n_tris: u64, tri_stride: u32, max_prims_per_leaf: u32, search_radius: u32,
        device: c_int, out: *mut *mut TrayScene, stats: *mut TrayBuildStats) -> c_int;
    pub fn tray_cuda_scene_download(scene: *mut TrayScene, nodes: *mut c_void, tris: *mut c_void, prim_indices: *mut u32) -> c_int;
    pub fn tray_cuda_frame_download(scene: *mut TrayScene, primary: *mut TrayHit, bounce: *mut TrayHit,
        bounce_rays: *mut TrayRay, rgba: *mut u8) -> c_int;
    pub fn tray_cuda_start(bvh: *const c_void, bvh_len: u64, inst: *const c_void, inst_len: u64, tris: *const c_void,
        tri_len: u64, tri_stride: u32, tlas_start: u32, use_tlas: c_int, view: *const TrayView, width: u32, height: u32,
        render_time_s: f32, benchmark: c_int, animate: c_int, device: c_int,
        out_min_ms: *mut f32, out_mean_ms: *mut f32, out_frames: *mut u32) -> c_int;
}

fn check(rc: c_int) {
    if rc != 0 {
        // the reference panics on every error (src/main.rs:178-180, rt_gpu_software.rs:99,124,307)
        let msg = unsafe { CStr::from_ptr(tray_cuda_last_error()) }.to_string_lossy().into_owned();
        panic!("tray_cuda error {rc}: {msg}");
    }
}

pub struct StartArgs<'a> {
    pub bvh_bytes: &'a [u8], pub instance_bytes: &'a [u8], pub tri_bytes: &'a [u8], pub tlas_start: u32, pub use_tlas: bool,
    pub view: TrayView, pub width: u32, pub height: u32, pub render_time: f32, pub benchmark: bool, pub animate: bool,
}

/// Drop-in for `rt_gpu_software::start`: returns the min frame time in ms.
pub fn start(a: StartArgs) -> f32 {
    let (mut min_ms, mut mean_ms, mut frames) = (0f32, 0f32, 0u32);
    check(unsafe {
        tray_cuda_start(a.bvh_bytes.as_ptr().cast(), a.bvh_bytes.len() as u64, a.instance_bytes.as_ptr().cast(),
            a.instance_bytes.len() as u64, a.tri_bytes.as_ptr().cast(), a.tri_bytes.len() as u64, 48, a.tlas_start,
            a.use_tlas as c_int, &a.view, a.width, a.height, a.render_time, a.benchmark as c_int, a.animate as c_int, 0,
            &mut min_ms, &mut mean_ms, &mut frames)
    });
    min_ms
}

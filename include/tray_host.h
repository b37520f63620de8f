/*
 * tray_host.h — CPU-side helpers around the tray_cuda ABI (libtray_host.so, no CUDA dependency).
 *
 * These are the host-side mirror of what surrounds the hot path in the reference:
 *   - a CWBVH producer standing in for obvhs' external PLOC builder
 *       (`cwbvh_from_tris` src/cwbvh.rs:24-105, `tlas_from_blas` src/cwbvh.rs:108-137)
 *   - the marshalling of `cwbvh_gpu_runner` (src/rt_gpu/mod.rs:16-112): per-object BLAS, triangles
 *     permuted into BVH order, primitive_base_idx globalised, BLAS|TLAS concatenation, blas_offsets
 *   - scene input: a minimal OBJ reader with `load_meshs` semantics (src/main.rs:493-561) and
 *     deterministic synthetic scenes sized like BASELINE.json's configs (the real assets are absent)
 * None of this is on the timed path; all of it is plain C ABI so a Rust host can ignore it entirely
 * and hand real obvhs output to tray_cuda.h.
 */
#ifndef TRAY_HOST_H
#define TRAY_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "tray_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- CWBVH builder --------------------------------------------------------------------------- */
typedef struct tray_cwbvh tray_cwbvh;   /* = obvhs CwBvh {nodes, primitive_indices, total_aabb} (src/cwbvh.rs:70-86) */

int tray_host_build_cwbvh_from_tris(const float* tris9, uint64_t n_tris, uint32_t max_prims_per_leaf,
                                    int nthreads, tray_cwbvh** out);
int tray_host_build_cwbvh_from_aabbs(const float* bmin3, const float* bmax3, uint64_t n,
                                     uint32_t max_prims_per_leaf, int nthreads, tray_cwbvh** out);
uint64_t tray_host_cwbvh_node_count(const tray_cwbvh* h);
uint64_t tray_host_cwbvh_prim_count(const tray_cwbvh* h);
uint32_t tray_host_cwbvh_max_depth(const tray_cwbvh* h);
const tray_cwbvh_node* tray_host_cwbvh_nodes(const tray_cwbvh* h);
const uint32_t* tray_host_cwbvh_prim_indices(const tray_cwbvh* h);
void tray_host_cwbvh_aabb(const tray_cwbvh* h, float mn[3], float mx[3]);
void tray_host_cwbvh_free(tray_cwbvh* h);

typedef struct tray_validate_report {
    uint64_t nodes_reached, nodes_unreached, prims_reached, prims_missing;
    uint64_t prim_seen_twice, node_visited_twice, bad_child_index, bad_prim_index;
    uint64_t bad_meta, bad_quant, box_violations;
    uint64_t children_total, leaf_children;
    uint32_t max_depth, max_stack, stack_too_deep, pad;
} tray_validate_report;

/* structural validation (role of obvhs `bvh.validate`, src/cwbvh.rs:102-104); 0 = valid */
int tray_host_cwbvh_validate(const tray_cwbvh_node* nodes, uint64_t n_nodes, const uint32_t* prim_indices,
                             uint64_t n_prims, const float* prim_min3, const float* prim_max3,
                             uint32_t stack_limit, tray_validate_report* report);

/* ---- meshes: Vec<Vec<Triangle>> (src/main.rs:493-561) ---------------------------------------- */
typedef struct tray_mesh tray_mesh;

/* OBJ: `v`, `f` (triangle, quad -> (a,b,c),(a,c,d); longer polygons contribute their first triangle,
 * as main.rs:535-551 does), one object per `o`. */
int tray_host_mesh_load_obj(const char* path, tray_mesh** out);
int tray_host_mesh_from_tris(const float* tris9, uint64_t n_tris, const uint64_t* object_offsets,
                             uint32_t n_objects, tray_mesh** out);
/* synthetic scenes: "kitchen" | "demoscene" | "hairball" | "sanmiguel" | "caldera" | "soup";
 * `size` in (0,1] scales the triangle count (1 = the BASELINE.json size). */
int tray_host_mesh_generate(const char* name, uint64_t seed, double size, tray_mesh** out);
uint64_t tray_host_mesh_tri_count(const tray_mesh* m);
uint32_t tray_host_mesh_object_count(const tray_mesh* m);
const float* tray_host_mesh_tris(const tray_mesh* m);               /* n_tris x 9 floats (v0,v1,v2) */
const uint64_t* tray_host_mesh_object_offsets(const tray_mesh* m);  /* n_objects + 1 */
/* camera that goes with a synthetic scene (eye, look_at, fov degrees), as the .ron files give it */
void tray_host_mesh_camera(const tray_mesh* m, float eye[3], float look_at[3], float* fov_deg);
void tray_host_mesh_free(tray_mesh* m);

/* ---- cwbvh_gpu_runner marshalling (src/rt_gpu/mod.rs:16-112) --------------------------------- */
typedef struct tray_packed tray_packed;

/* use_tlas = 0: objects are flattened into one BLAS first (src/main.rs:300-308).
 * tri_stride: 48 or 64 (tray_cuda.h). */
int tray_host_pack(const tray_mesh* mesh, int use_tlas, uint32_t tri_stride, uint32_t max_prims_per_leaf,
                   int nthreads, tray_packed** out);
const void* tray_host_packed_bvh_bytes(const tray_packed* p, uint64_t* len);
const void* tray_host_packed_tri_bytes(const tray_packed* p, uint64_t* len);
const void* tray_host_packed_instance_bytes(const tray_packed* p, uint64_t* len);
/* global BVH-ordered triangle slot -> index into the mesh's triangle list */
const uint32_t* tray_host_packed_prim_to_mesh_tri(const tray_packed* p, uint64_t* count);
/* per-BLAS first global triangle slot (n_blas + 1 entries): maps a global prim to (geometry_id, local) */
const uint64_t* tray_host_packed_blas_tri_offsets(const tray_packed* p, uint32_t* n_blas);
uint32_t tray_host_packed_tlas_start(const tray_packed* p);
uint32_t tray_host_packed_max_depth(const tray_packed* p);
double tray_host_packed_build_seconds(const tray_packed* p, double* tlas_seconds);
void tray_host_packed_free(tray_packed* p);

/* ViewUniform::from_camera (src/main.rs:599-616), f32 arithmetic in glam's operation order */
void tray_host_view_from_camera(const float eye[3], const float look_at[3], float fov_deg,
                                float width, float height, float exposure, uint32_t tlas_start,
                                tray_view* out);

#ifdef __cplusplus
}
#endif
#endif

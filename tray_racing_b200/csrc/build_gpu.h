// build_gpu.h — internal interface between the C ABI (tray_cuda.cu) and the device-side builder (build_gpu.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "tray_cuda.h"

namespace tray_build {

struct Result {
    uint8_t* d_nodes;            // n_nodes x 80 bytes (allocation holds n_tris + 1 nodes); caller owns (cudaFree)
    uint64_t n_nodes;
    uint8_t* d_tris;             // n_tris x tri_stride, BVH order; caller owns
    uint32_t* d_prim_indices;    // BVH slot -> input triangle; caller owns
    uint32_t* d_blas_offsets;    // two-level builds: node index of the BLAS behind TLAS leaf k; caller owns
    uint32_t n_instances, tlas_start;
    bool force_exact;            // some node scale >= 2^40 (see tray_scene::force_exact)
    tray_build_stats stats;
};

// Builds on the current device, synchronously with respect to `st`.  Returns 0, or a negative code with a message in `err`.
int build(const float* tris9_host, uint64_t n_tris, uint32_t tri_stride, uint32_t max_prims_per_leaf, uint32_t search_radius,
          cudaStream_t st, Result* out, char* err, size_t errlen);

// Two-level build: one BLAS per object (object k = triangles [object_offsets[k], object_offsets[k + 1])) + a TLAS over them.
int build_tlas(const float* tris9_host, uint64_t n_tris, const uint64_t* object_offsets, uint32_t n_objects, uint32_t tri_stride,
               uint32_t max_prims_per_leaf, uint32_t search_radius, cudaStream_t st, Result* out, char* err, size_t errlen);

}  // namespace tray_build

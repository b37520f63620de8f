// traverse_pool.cuh — the pooled traversal kernel (kernel v6): same per-ray state machine, new lane assignment.
//
// trace_kernel (traverse.cuh) ties one ray to one lane: on every iteration the lanes whose ray wants the OTHER kind
// of step sit idle (measured 21/32 active lanes on primary rays, 15.8/32 on bounce rays; the issue-slot model in
// tests/tools/sched_sim.py reproduces both).  Here every warp owns a POOL of 64 rays whose whole state lives in
// shared memory.  Each iteration the warp takes a census of the pool, votes node step or triangle step, COMPACTS up
// to 32 rays that want that step onto its 32 lanes (prefix-popc rank -> tiny shared selection list), runs the step
// from / back to shared memory, and refills retired slots from the global cursor.  Any lane can run any ray, so the
// lanes stay full in both phases (model: 29/32 and 26/32).
//
// Per-ray semantics are untouched: each ray still performs exactly the reference's sequence of node tests and
// triangle tests with its own shrinking tmax (query.hlsl:328-438), so (prim, t) and the node/triangle counters stay
// bit-identical to the oracle whatever the pool does.
//
// Shared memory per warp (64 slots): 5 x 16 B state records + an 8-entry stack of 8 B = 144 B per ray = 9.2 KB,
// 4 warps per CTA, 6 CTAs per SM.  Stack entries past 8 spill to a per-slot global scratch area (rare).
#pragma once
#include "traverse.cuh"

namespace tray {

constexpr int POOL_WARPS = 4;
constexpr int POOL_SLOTS = 64;            // rays per warp
#ifndef TRAY_POOL_STACK
#define TRAY_POOL_STACK 8
#endif
#ifndef TRAY_POOL_MIN_BLOCKS
#define TRAY_POOL_MIN_BLOCKS 6
#endif
constexpr int POOL_STACK_SMEM = TRAY_POOL_STACK;   // stack entries per ray in shared memory
constexpr int POOL_STACK_SPILL = 48 - POOL_STACK_SMEM;      // further entries per ray in global scratch (total 48 > obvhs' 32)

struct PoolWarpSmem {
    float4 A[POOL_SLOTS];                 // origin.xyz, best_t (the ray's current tmax)
    float4 B[POOL_SLOTS];                 // 1/dir.xyz, misc: oct_inv4 (bits 0-2 of every byte) | wide << 31
    float4 C[POOL_SLOTS];                 // dir.xyz (after the zero fix-up), tmin
    uint4 D[POOL_SLOTS];                  // cur_x, tri_x, cur_y, tri_y | sp << 24     (idle: cur_y == 0 and tri bits == 0)
    uint4 E[POOL_SLOTS];                  // best_prim, ray index, bvh_off, tlas_sp
    uint2 stack[POOL_STACK_SMEM][POOL_SLOTS];
    uint8_t sel[32];                      // slot run by lane i in this iteration
};

template <bool TLAS, bool COUNT, int TRI_STRIDE>
__global__ void __launch_bounds__(POOL_WARPS * 32, TRAY_POOL_MIN_BLOCKS) trace_pool_kernel(const __grid_constant__ TraceParams P) {
    __shared__ PoolWarpSmem s_pool[POOL_WARPS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    PoolWarpSmem& W = s_pool[warp];
    uint2* const spill = P.spill + (size_t)(blockIdx.x * POOL_WARPS + warp) * POOL_STACK_SPILL * POOL_SLOTS;
    const uint32_t k4b = P.k4b;
    const uint32_t n_work = P.n_work_dev ? *P.n_work_dev : P.n_work;
    unsigned long long c_rays = 0, c_nodes = 0, c_tris = 0, c_insts = 0, c_hits = 0;

    W.D[lane] = make_uint4(0, 0, 0, 0);
    W.D[lane + 32] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    bool exhausted = false;        // warp-uniform
    unsigned flip = 0;             // alternates which half of the pool is ranked first

    auto push = [&](uint32_t s, uint32_t& sp, uint32_t x, uint32_t y) {
        if (sp < POOL_STACK_SMEM) W.stack[sp][s] = make_uint2(x, y);
        else if (sp < POOL_STACK_SMEM + POOL_STACK_SPILL) spill[(sp - POOL_STACK_SMEM) * POOL_SLOTS + s] = make_uint2(x, y);
        else { atomicOr(P.overflow, 1u); return; }
        sp++;
    };
    // the ray in slot s has neither triangles nor nodes in hand: pop its stack, or retire it
    // (query.hlsl:417-427; tlas:480-486; a popped triangle group is query.hlsl:389-393)
    auto pop_or_retire = [&](uint32_t s, uint32_t& sp, uint32_t& cur_x, uint32_t& cur_y, uint32_t& tri_x, uint32_t& tri_y,
                             float best_t, uint32_t best_prim, uint32_t ray_idx, uint32_t& bvh_off, uint32_t& tlas_sp) {
        if (sp == 0) {
            tray_hit h;
            h.t = best_prim != INVALID ? best_t : __int_as_float(0x7f800000);   // RayHit::none()
            h.prim = best_prim;
            const uint32_t item = P.ray_item ? __ldg(P.ray_item + ray_idx) : ray_idx;
            P.hits_out[item] = h;
            if (P.rgba_out) {
                float col;
                if (P.shade_mode == SHADE_PRIMARY) col = __fdiv_rn(1.0f, h.t);                       // rt_cpu.rs:59
                else col = h.t < F32_MAX_ ? __fdiv_rn(h.t, add(1.0f, h.t)) : 1.0f;                   // rt_cpu.rs:82-87
                const long long o = rgba_slot(item, P.frame_w, P.frame_h, P.frame_tiles_x, P.frame_shard, P.frame_shards);
                if (o >= 0) P.rgba_out[o] = shade(col);
            }
            if (COUNT && best_prim != INVALID) c_hits++;
            cur_x = 0; cur_y = 0; tri_x = 0; tri_y = 0;                                              // slot is idle
        } else {
            if (TLAS && sp == tlas_sp) { tlas_sp = INVALID; bvh_off = P.tlas_start; }
            sp--;
            const uint2 e = sp < POOL_STACK_SMEM ? W.stack[sp][s] : spill[(sp - POOL_STACK_SMEM) * POOL_SLOTS + s];
            if (e.y & 0xff000000u) { cur_x = e.x; cur_y = e.y; tri_y = 0; }
            else { tri_x = e.x; tri_y = e.y; cur_x = 0; cur_y = 0; }
        }
    };

    for (;;) {
        // ---- census: lane l reports on slots l and l + 32 ----
        const uint2 q0 = reinterpret_cast<const uint2*>(&W.D[lane])[1];          // (cur_y, tri_y | sp << 24)
        const uint2 q1 = reinterpret_cast<const uint2*>(&W.D[lane + 32])[1];
        const bool wt0 = (q0.y & 0x00ffffffu) != 0u, wn0 = !wt0 && q0.x >= 0x01000000u;
        const bool wt1 = (q1.y & 0x00ffffffu) != 0u, wn1 = !wt1 && q1.x >= 0x01000000u;
        const unsigned mt0 = __ballot_sync(FULL, wt0), mt1 = __ballot_sync(FULL, wt1);
        const unsigned mn0 = __ballot_sync(FULL, wn0), mn1 = __ballot_sync(FULL, wn1);
        const unsigned n_tri = (unsigned)(__popc(mt0) + __popc(mt1)), n_node = (unsigned)(__popc(mn0) + __popc(mn1));
        const unsigned busy = n_tri + n_node;

        // ---- refill retired slots from the global cursor ----
        if (busy != POOL_SLOTS) {
            if (!exhausted && (POOL_SLOTS - busy >= P.refill_min || busy == 0u)) {
                const unsigned i0 = ~(mt0 | mn0), i1 = ~(mt1 | mn1);
                const uint32_t n0 = (uint32_t)__popc(i0), n_idle = n0 + (uint32_t)__popc(i1);
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(P.cursor, n_idle);
                base = __shfl_sync(FULL, base, 0);
                if (base + n_idle >= n_work) exhausted = true;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const unsigned im = half ? i1 : i0;
                    if ((im >> lane) & 1u) {
                        const uint32_t ray_idx = base + (half ? n0 : 0u) + (uint32_t)__popc(im & lt);
                        if (base < n_work && ray_idx < n_work) {
                            const uint32_t s = lane + 32u * half;
                            const float4* rp = reinterpret_cast<const float4*>(P.rays + ray_idx);
                            const float4 a = __ldg(rp), b = __ldg(rp + 1);
                            RayConst r; prepare_ray(r, a.x, a.y, a.z, b.x, b.y, b.z, a.w);
                            W.A[s] = make_float4(r.ox, r.oy, r.oz, b.w);
                            W.B[s] = make_float4(r.ix, r.iy, r.iz, __uint_as_float(r.oct_inv4 | (r.wide ? 0x80000000u : 0u)));
                            W.C[s] = make_float4(r.dx, r.dy, r.dz, r.tmin);
                            W.D[s] = make_uint4(0u, 0u, 0x80000000u, 0u);                 // root group, query.hlsl:343
                            W.E[s] = make_uint4(INVALID, ray_idx, TLAS ? P.tlas_start : 0u, INVALID);
                            if (COUNT) c_rays++;
                        }
                    }
                }
                __syncwarp();
                continue;                                    // census again with the new rays
            }
            if (busy == 0u) break;                           // cursor exhausted and nothing in flight
        }

        // ---- vote, then compact up to 32 rays of the voted kind onto the lanes ----
        const unsigned l_tri = n_tri < 32u ? n_tri : 32u, l_node = n_node < 32u ? n_node : 32u;
        const bool tri_phase = n_node == 0u || l_tri * P.tri_weight >= l_node;
        const unsigned m0 = tri_phase ? mt0 : mn0, m1 = tri_phase ? mt1 : mn1;
        flip ^= 1u;
        const unsigned mf = flip ? m1 : m0, ms = flip ? m0 : m1;          // mf is ranked first
        const unsigned nf = (unsigned)__popc(mf);
        if ((mf >> lane) & 1u) W.sel[__popc(mf & lt)] = (uint8_t)(lane + (flip ? 32u : 0u));
        if ((ms >> lane) & 1u) {
            const unsigned rk = nf + (unsigned)__popc(ms & lt);
            if (rk < 32u) W.sel[rk] = (uint8_t)(lane + (flip ? 0u : 32u));
        }
        __syncwarp();
        const unsigned n_sel = tri_phase ? l_tri : l_node;

        if (lane < n_sel) {
            const uint32_t s = W.sel[lane];
            const float4 a = W.A[s];
            const uint4 d = W.D[s];
            uint32_t cur_x = d.x, tri_x = d.y, cur_y = d.z, tri_y = d.w & 0x00ffffffu, sp = d.w >> 24;
            uint32_t bvh_off = 0, tlas_sp = INVALID;
            if (TLAS) { const uint2 e = reinterpret_cast<const uint2*>(&W.E[s])[1]; bvh_off = e.x; tlas_sp = e.y; }
            const uint32_t bvh_off_in = bvh_off, tlas_sp_in = tlas_sp;
            float best_t = a.w;

            if (!tri_phase) {
                // ---- node step (query.hlsl:354-387) ----
                const float4 b = W.B[s];
                const uint32_t misc = __float_as_uint(b.w);
                RayConst r;
                r.ox = a.x; r.oy = a.y; r.oz = a.z; r.ix = b.x; r.iy = b.y; r.iz = b.z;
                r.dx = b.x; r.dy = b.y; r.dz = b.z;            // only the SIGN of dir is used by the node test; sign(1/d) == sign(d)
                r.tmin = 0.f; r.oct_inv4 = misc & 0x07070707u; r.wide = (misc >> 31) != 0u;
                const uint32_t hits_imask = cur_y;
                const uint32_t off = 31u - (uint32_t)__clz((int)hits_imask);                 // query.hlsl:358
                cur_y &= ~(1u << off);                                                         // :362
                if (cur_y & 0xff000000u) push(s, sp, cur_x, cur_y);                            // :365-368
                const uint32_t slot = (off - 24u) ^ (r.oct_inv4 & 0xffu);                      // :370
                const uint32_t rel = (uint32_t)__popc(hits_imask & ~(0xffffffffu << slot));    // :371
                const uint4* np = P.nodes + (size_t)(bvh_off + cur_x + rel) * 5u;              // :373, tlas:383
                const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
                if (COUNT) c_nodes++;
                const uint32_t hitmask = (r.wide || P.force_exact) ? node_test(r, best_t, n0, n1, n2, n3, n4, k4b)
                                                : node_test_fast(r, best_t, n0, n1, n2, n3, n4, k4b, P.zero);   // :380
                cur_x = n1.x; tri_x = n1.y;                                                    // :383-384
                cur_y = (hitmask & 0xff000000u) | (n0.w >> 24);                                // :386
                tri_y = hitmask & 0x00ffffffu;                                                 // :387
                if (tri_y == 0u && cur_y < 0x01000000u) {
                    const uint2 e = reinterpret_cast<const uint2*>(&W.E[s])[0];
                    pop_or_retire(s, sp, cur_x, cur_y, tri_x, tri_y, best_t, e.x, e.y, bvh_off, tlas_sp);
                }
            } else {
                // ---- triangle step (query.hlsl:396-413), or TLAS instance entry (query_tlas.hlsl:410-446) ----
                const uint32_t local = 31u - (uint32_t)__clz((int)tri_y);                      // :398
                tri_y &= ~(1u << local);                                                       // :401
                const uint32_t g = tri_x + local;                                              // :403
                if (TLAS && tlas_sp == INVALID) {
                    if (tri_y != 0u) push(s, sp, tri_x, tri_y);
                    if (cur_y & 0xff000000u) push(s, sp, cur_x, cur_y);
                    tlas_sp = sp;
                    bvh_off = __ldg(P.blas_offsets + g);
                    if (COUNT) c_insts++;
                    cur_x = 0; cur_y = 0x80000000u; tri_y = 0;
                } else {
                    const float4 c = W.C[s];
                    RayConst r;
                    r.ox = a.x; r.oy = a.y; r.oz = a.z; r.dx = c.x; r.dy = c.y; r.dz = c.z; r.tmin = c.w;
                    if (COUNT) c_tris++;
                    const float t = tri_test<TRI_STRIDE>(r, best_t, P.tris, g);
                    uint2 e = reinterpret_cast<const uint2*>(&W.E[s])[0];       // best_prim, ray index
                    if (t < best_t) {                                           // CPU tie rule: first of equal t wins (§8a a11)
                        best_t = t; e.x = g;
                        W.A[s].w = t; W.E[s].x = g;
                    }
                    if (tri_y == 0u && cur_y < 0x01000000u)
                        pop_or_retire(s, sp, cur_x, cur_y, tri_x, tri_y, best_t, e.x, e.y, bvh_off, tlas_sp);
                }
            }
            W.D[s] = make_uint4(cur_x, tri_x, cur_y, tri_y | (sp << 24));
            if (TLAS && (bvh_off != bvh_off_in || tlas_sp != tlas_sp_in))
                reinterpret_cast<uint2*>(&W.E[s])[1] = make_uint2(bvh_off, tlas_sp);
        }
        __syncwarp();
    }

    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_rays += __shfl_xor_sync(FULL, c_rays, o); c_nodes += __shfl_xor_sync(FULL, c_nodes, o);
            c_tris += __shfl_xor_sync(FULL, c_tris, o); c_insts += __shfl_xor_sync(FULL, c_insts, o);
            c_hits += __shfl_xor_sync(FULL, c_hits, o);
        }
        if (lane == 0) {
            atomicAdd(P.counters + 0, c_rays); atomicAdd(P.counters + 1, c_nodes); atomicAdd(P.counters + 2, c_tris);
            atomicAdd(P.counters + 3, c_insts); atomicAdd(P.counters + 4, c_hits);
        }
    }
}

}  // namespace tray

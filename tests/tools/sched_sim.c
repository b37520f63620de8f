/* sched_sim.c — issue-slot model of the traversal kernel's WARP schedule (test/analysis infrastructure).
 *
 * Input: the per-ray step logs of the CPU oracle ('N' node step, 'T' triangle step, 'I' instance entry), in the
 * order the kernel's cursor hands rays out.  The model replays them through persistent 32-lane warps with ray
 * replacement and a per-iteration phase vote, for three lane organisations:
 *   policy 0  K ray slots per LANE: a lane joins the voted phase if any of its K rays wants it (K = 1 is the kernel
 *             of round 1)
 *   policy 1  warp-level POOL of K rays: any lane can run any ray (upper bound for in-warp compaction)
 *   policy 2  one ray per lane, but a triangle phase DRAINS every participating lane's triangle group
 *             (cost = csel + ct per round, rounds = the longest group)
 *   policy 3  one ray per lane; a node step with at most K participating lanes runs NARROW: 8 lanes test the 8
 *             children of one ray, 4 rays per pass, cost = csel per pass (csel = slots of one narrow pass)
 * Each step costs issue slots (cn, ct, plus csel when K > 1); the result is total slots, per-phase lane occupancy
 * and the makespan over warps.  Nothing here is on the product path. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct sim_cfg {
    int policy, K, n_warps, refill_min, tri_weight;
    int cn, ct, cr, csel;          /* issue slots: node step, triangle step, refill, per-step selection overhead */
} sim_cfg;

typedef struct sim_out {
    double slots, makespan, node_steps, tri_steps, node_lanes, tri_lanes, refills;
} sim_out;

typedef struct warp {
    double t;
    uint64_t* pos;   /* 32*K cursors into ops (end == idle) */
    uint64_t* end;
    int exhausted;
} warp;

static void heap_sift(int* h, int n, int i, const warp* w) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && w[h[l]].t < w[h[m]].t) m = l;
        if (r < n && w[h[r]].t < w[h[m]].t) m = r;
        if (m == i) return;
        int tmp = h[i]; h[i] = h[m]; h[m] = tmp; i = m;
    }
}

int sched_sim(const uint8_t* ops, const uint64_t* offsets, uint64_t n_rays, const sim_cfg* c, sim_out* o) {
    const int K = c->policy == 1 ? 1 : (c->policy >= 2 ? 1 : c->K), S = c->policy == 1 ? c->K : 32 * K;   /* policy 1: K is the pool size */
    warp* w = (warp*)calloc((size_t)c->n_warps, sizeof(warp));
    int* heap = (int*)malloc(sizeof(int) * (size_t)c->n_warps);
    for (int i = 0; i < c->n_warps; i++) {
        w[i].pos = (uint64_t*)calloc((size_t)S, 8); w[i].end = (uint64_t*)calloc((size_t)S, 8); heap[i] = i;
    }
    memset(o, 0, sizeof *o);
    uint64_t cursor = 0;
    int live = c->n_warps;
    while (live > 0) {
        warp* W = &w[heap[0]];
        /* census */
        int idle_slots = 0, lanes_tri = 0, lanes_node = 0, pool_tri = 0, pool_node = 0;
        for (int l = 0; l < (c->policy == 1 ? S : 32); l++) {
            int lt = 0, ln = 0;
            for (int k = 0; k < K; k++) {
                const int s = l * K + k;
                if (W->pos[s] == W->end[s]) { idle_slots++; continue; }
                if (ops[W->pos[s]] == 'N') { ln = 1; pool_node++; } else { lt = 1; pool_tri++; }
            }
            lanes_tri += lt; lanes_node += ln;
        }
        int n_tri = c->policy == 1 ? (pool_tri < 32 ? pool_tri : 32) : lanes_tri;
        int n_node = c->policy == 1 ? (pool_node < 32 ? pool_node : 32) : lanes_node;
        const int busy = S - idle_slots;
        if (idle_slots > 0 && !W->exhausted && (idle_slots >= c->refill_min * (K > 1 ? 1 : 1) || busy == 0)) {
            for (int s = 0; s < S; s++)
                if (W->pos[s] == W->end[s]) {
                    if (cursor < n_rays) { W->pos[s] = offsets[cursor]; W->end[s] = offsets[cursor + 1]; cursor++; }
                }
            if (cursor >= n_rays) W->exhausted = 1;
            W->t += c->cr; o->slots += c->cr; o->refills += 1;
        } else if (busy == 0) {
            /* retire the warp */
            if (W->t > o->makespan) o->makespan = W->t;
            heap[0] = heap[--live];
            if (live > 0) heap_sift(heap, live, 0, w);
            continue;
        } else {
            const int tri_phase = n_node == 0 || n_tri * c->tri_weight >= n_node;
            const uint8_t want = tri_phase ? 'T' : 'N';
            int done = 0;
            if (c->policy == 2 && tri_phase) {
                /* drain: rounds until no participating lane has a 'T' at its cursor */
                int rounds = 0;
                for (;;) {
                    int any = 0;
                    for (int s = 0; s < S; s++)
                        if (W->pos[s] != W->end[s] && ops[W->pos[s]] != 'N') { W->pos[s]++; any++; }
                    if (!any) break;
                    rounds++; o->tri_steps += 1; o->tri_lanes += any;
                }
                const int cost = c->csel + rounds * c->ct;
                W->t += cost; o->slots += cost;
                heap_sift(heap, live, 0, w);
                continue;
            }
            if (c->policy == 1) {
                for (int s = 0; s < S && done < 32; s++)
                    if (W->pos[s] != W->end[s] && ((ops[W->pos[s]] == 'N') == (want == 'N'))) { W->pos[s]++; done++; }
            } else {
                for (int l = 0; l < 32; l++)
                    for (int k = 0; k < K; k++) {
                        const int s = l * K + k;
                        if (W->pos[s] != W->end[s] && ((ops[W->pos[s]] == 'N') == (want == 'N'))) { W->pos[s]++; done++; break; }
                    }
            }
            int cost = (tri_phase ? c->ct : c->cn) + ((c->K > 1 && c->policy == 0) || c->policy == 1 ? c->csel : 0);
            if (c->policy == 3 && !tri_phase && done <= c->K) cost = ((done + 3) / 4) * c->csel;
            W->t += cost; o->slots += cost;
            if (tri_phase) { o->tri_steps += 1; o->tri_lanes += done; } else { o->node_steps += 1; o->node_lanes += done; }
        }
        heap_sift(heap, live, 0, w);
    }
    for (int i = 0; i < c->n_warps; i++) { free(w[i].pos); free(w[i].end); }
    free(w); free(heap);
    return 0;
}

/* ---- policy 4 (round 2): one ray per lane, two triangles of a group per triangle step (tri2), and optionally a SPECULATIVE
 * node step: a lane that waits with a triangle group whose successor is a node step joins the warp's node phase early, testing
 * that node against its current tmax; when its triangle group ends without having moved tmax the result stands (the node step
 * is skipped), otherwise it is thrown away and the step is replayed.  Input: the detailed log (orc_trace_oplog_detail).
 * cfg.K: bit 0 = tri2, bit 1 = speculate; csel = extra slots of a node step that carries speculative lanes. */
static int is_later(uint8_t c) { return c == 'U' || c == 'u'; }
int sched_sim4(const uint8_t* ops, const uint64_t* offsets, uint64_t n_rays, const sim_cfg* c, sim_out* o, double* extra) {
    const int tri2 = c->K & 1, spec_on = (c->K >> 1) & 1;
    warp* w = (warp*)calloc((size_t)c->n_warps, sizeof(warp));
    int* heap = (int*)malloc(sizeof(int) * (size_t)c->n_warps);
    uint8_t* spec = (uint8_t*)calloc((size_t)c->n_warps * 32, 1);     /* 0 none, 1 speculated, 2 speculated and dirty */
    for (int i = 0; i < c->n_warps; i++) { w[i].pos = (uint64_t*)calloc(32, 8); w[i].end = (uint64_t*)calloc(32, 8); heap[i] = i; }
    memset(o, 0, sizeof *o);
    double n_spec = 0, n_commit = 0, n_replay = 0;
    uint64_t cursor = 0;
    int live = c->n_warps;
    while (live > 0) {
        const int wi = heap[0];
        warp* W = &w[wi];
        uint8_t* sp = spec + (size_t)wi * 32;
        int idle = 0, n_tri = 0, n_node = 0;
        for (int l = 0; l < 32; l++) {
            if (W->pos[l] == W->end[l]) { idle++; continue; }
            if (ops[W->pos[l]] == 'N') n_node++; else n_tri++;
        }
        const int busy = 32 - idle;
        if (idle > 0 && !W->exhausted && (idle >= c->refill_min || busy == 0)) {
            for (int l = 0; l < 32; l++)
                if (W->pos[l] == W->end[l] && cursor < n_rays) { W->pos[l] = offsets[cursor]; W->end[l] = offsets[cursor + 1]; cursor++; sp[l] = 0; }
            if (cursor >= n_rays) W->exhausted = 1;
            W->t += c->cr; o->slots += c->cr; o->refills += 1;
        } else if (busy == 0) {
            if (W->t > o->makespan) o->makespan = W->t;
            heap[0] = heap[--live];
            if (live > 0) heap_sift(heap, live, 0, w);
            continue;
        } else {
            const int tri_phase = n_node == 0 || n_tri * c->tri_weight >= n_node;
            if (!tri_phase) {
                int done = 0, specs = 0;
                for (int l = 0; l < 32; l++) {
                    if (W->pos[l] == W->end[l]) continue;
                    const uint8_t op = ops[W->pos[l]];
                    if (op == 'N') { W->pos[l]++; done++; }
                    else if (spec_on && sp[l] == 0 && op != 'I') {
                        uint64_t e = W->pos[l] + 1;
                        while (e < W->end[l] && is_later(ops[e])) e++;
                        if (e < W->end[l] && ops[e] == 'N') { sp[l] = 1; specs++; }
                    }
                }
                const int cost = c->cn + (specs ? c->csel : 0);
                W->t += cost; o->slots += cost; o->node_steps += 1; o->node_lanes += done + specs; n_spec += specs;
            } else {
                int done = 0;
                for (int l = 0; l < 32; l++) {
                    if (W->pos[l] == W->end[l] || ops[W->pos[l]] == 'N') continue;
                    done++;
                    int n = 1;
                    if (tri2 && ops[W->pos[l]] != 'I' && W->pos[l] + 1 < W->end[l] && is_later(ops[W->pos[l] + 1])) n = 2;
                    for (int k = 0; k < n; k++) { if ((ops[W->pos[l]] & 0x20) && sp[l] == 1) sp[l] = 2; W->pos[l]++; }
                    const int group_over = W->pos[l] == W->end[l] || !is_later(ops[W->pos[l]]);
                    if (group_over && sp[l]) {
                        if (sp[l] == 1 && W->pos[l] < W->end[l] && ops[W->pos[l]] == 'N') { W->pos[l]++; n_commit += 1; }
                        else n_replay += 1;
                        sp[l] = 0;
                    }
                }
                W->t += c->ct; o->slots += c->ct; o->tri_steps += 1; o->tri_lanes += done;
            }
        }
        heap_sift(heap, live, 0, w);
    }
    if (extra) { extra[0] = n_spec; extra[1] = n_commit; extra[2] = n_replay; }
    for (int i = 0; i < c->n_warps; i++) { free(w[i].pos); free(w[i].end); }
    free(w); free(heap); free(spec);
    return 0;
}

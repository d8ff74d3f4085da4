// world_step.cuh -- the rest of Physics.Engine.Main.updateWorld on the device (SURVEY.md section 8f,
// ranks 2 and 4): body state resident in HBM, applyExternal, applyCachedSlns' velocity updates,
// improveWorld and advance.  Included by shapes_b200.cu (one translation unit).
//
//   updateWorld (shapes/src/Physics/Engine/Main.hs:71-86):
//     keys      <- culledKeys                       } run_frame (K0..K3, the hot path)
//     applyExternal exts dt world                   } k_external          (World.hs:156-158)
//     kContacts <- prepareFrame keys world          } run_frame
//     applyCachedSlns ...                           } k_warm_join + pass 0 of k_solve
//     improveWorld solutionProcessor ... (x2)       } passes 1.. of k_solve (Solvers/Contact.hs:124-157)
//     advance dt world; moveShapes world            } k_advance (+ next frame's K0)
//
// The reference solver is a strictly sequential Gauss-Seidel walk over the contact list.  A contact
// only reads and writes the velocities of its own two bodies, so the walk is a DAG: contact k must
// wait for the previous contact (in list order, across sweeps) that touches either of its bodies, and
// for nothing else.  k_solve executes that DAG as a dataflow graph: one node per (sweep, pair with
// contacts), a dependency counter per node, a ready queue, persistent threads that follow a chain
// as long as the node they just released is ready.  Every body therefore sees its velocity updates
// in exactly the reference's order and the result is BIT-IDENTICAL to the sequential walk, whatever
// the schedule.  The sweeps pipeline into each other (sweep s+1 starts on a body as soon as sweep s
// has left it), so the cost is set by the longest dependency chain, not by sweeps x contacts.
//
// Known caveat: a fully static body (inv_lin == 0 && inv_rot == 0) is not a dependency carrier and
// its velocity is never written; the reference adds (j*l)*0 to it, which differs only in the sign
// of a zero velocity component (or when an impulse is non-finite).

#pragma once

#include "shapes_sincos.h"

#include <cub/device/device_radix_sort.cuh>

namespace {

constexpr unsigned SOLVE_NONE = 0xffffffffu;     // "no node" / empty queue slot
constexpr int SOLVE_PASS_SHIFT = 28;             // queue entry = node | pass << 28
constexpr unsigned SOLVE_NODE_MASK = (1u << SOLVE_PASS_SHIFT) - 1u;
constexpr int SOLVE_THREADS = 128;
constexpr unsigned long long SOLVE_IDLE_LIMIT_NS = 15000000000ull;  // a warp without work for 15 s of wall clock gives up

// Shared scheduler words, one 128 B line each: the idle pollers, the pushers and the chain ends
// must not queue up behind each other on one L2 sector.
struct __align__(128) SolveState {
    unsigned long long head;      // next queue slot to claim
    unsigned long long pad0[15];
    unsigned long long tail;      // next queue slot to fill
    unsigned long long pad1[15];
    unsigned done;                // body chains that finished their last sweep
    unsigned pad2[31];
    unsigned expected;            // body chains that exist (dynamic bodies with at least one contact)
    int overflow;                 // 1 = queue overflow, 2 = idle time-out (both are internal errors)
    unsigned long long nodes_run; // statistics
    unsigned long long queue_cap;
    unsigned long long pad3[13];
};

// Everything static a node needs, one 32 B sector: bodies, first contact row, successor links.
struct __align__(32) NodeRec {
    int i, j;                     // pair_i (larger key), pair_j
    unsigned r0;                  // first contact row of the pair
    unsigned flags;               // rows [0,2) | body i dynamic [2] | body j dynamic [3]
    int next_i, next_j;           // >= 0: next node on that body in this sweep; < 0: -1 - (first node of the body's chain)
    unsigned succ_r0_i, succ_r0_j; // first contact row of those two successor nodes (so their rows can be prefetched a node ahead)
};
// One contact row as the solver reads it (array of structures, 192 B = 6 sectors): the SoA result
// columns stay the product's output; this copy is packed by k_pack_rows for the gather-heavy solver.
constexpr int ROW_DOUBLES = 24;   // jn[6] jf[6] b_np inv_eff_np inv_eff_f ra[2] rb[2] rn[2] hit pad[2]
struct __align__(32) BodyRec { double inv_lin, inv_rot, mu, bounce; };

struct SolveParams {
    const FrameState *st;
    NodeRec *node;
    double *rows;                 // n_contacts x ROW_DOUBLES
    const BodyRec *body;
    const int32_t *pair_i, *pair_j;
    const uint32_t *ccnt, *coff;
    const double *j_np[6], *b_np, *ra_x, *ra_y, *rb_x, *rb_y, *rn_x, *rn_y, *j_f[6], *inv_eff_np, *inv_eff_f;
    const uint8_t *hit;
    double *lam_np, *lam_f;
    double2 *vel;                 // per body: [2b] = (vx, vy), [2b+1] = (w, unused)
    const double2 *mass;          // (inv_lin, inv_rot), packed by K0
    const double *mu, *bounce;
    int32_t *next_i, *next_j;     // per pair: >= 0 next node on that body in this sweep; < 0: -1 - (first node of the body) = wrap
    int32_t *first_i, *first_j;   // per body: first live pair where the body is the larger / the smaller key
    int32_t *cnt;                 // per pair: unmet dependencies of its next execution
    uint32_t *sort_key[2], *sort_val[2];
    uint32_t *queue;
    SolveState *ss;
    int n_slots;
    int p_begin, p_end;           // sweeps [p_begin, p_end): 0 = applyCachedSlns, >= 1 = improveWorld
    unsigned long long queue_cap;
    int lanes;                    // executor lanes per warp (the others exit at once)
    unsigned sleep_cap;           // longest idle back-off in ns (0 = spin)
    unsigned claim_after;         // iterations an idle lane of a busy warp waits for a sibling's hand-over before it claims a queue slot
    int prefer_i;                 // both successors ready: 1 = continue on body i's chain, 0 = on body j's
};

__device__ __forceinline__ bool body_dynamic(double2 m) { return !(m.x == 0.0 && m.y == 0.0); }   // not isStatic (Constraint.hs:123-125)

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) { return *reinterpret_cast<const volatile unsigned *>(p); }
__device__ __forceinline__ void st_volatile_u32(unsigned *p, unsigned v) { *reinterpret_cast<volatile unsigned *>(p) = v; }

// ---- body state ----------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_pack_vel(int n, const double *vx, const double *vy, const double *w, double2 *vel)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    vel[2 * s] = make_double2(vx[s], vy[s]);
    vel[2 * s + 1] = make_double2(w[s], 0.0);
}

__global__ void __launch_bounds__(256) k_unpack_vel(int n, const double2 *vel, double *vx, double *vy, double *w)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double2 a = vel[2 * s], b = vel[2 * s + 1];
    vx[s] = a.x; vy[s] = a.y; w[s] = b.x;
}

// shapes_sincos over a column (moveShapes' rotate22 with the shared host/device routine)
__global__ void __launch_bounds__(256) k_sincos(int n, const double *rot, double *c, double *s)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double cc, ss;
    shapes_sincos_inline(rot[k], &cc, &ss);
    c[k] = cc; s[k] = ss;
}

// applyExternal (World.hs:156-158) over the filled slots; constantAccel / constantForce
// (World/External.hs:16-28; the latter as the reference parses it: (v + f*dt) * inv_lin).
__global__ void __launch_bounds__(256) k_external(int n, const uint8_t *alive, int kind, double ex, double ey, double dt,
                                                  const double *inv_lin, double2 *vel)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || !alive[s]) return;
    const double il = inv_lin[s];
    double2 v = vel[2 * s];
    if (kind == SHAPES_EXT_ACCEL) {
        if (0.0 == il) return;                       // isStaticLin (Constraint.hs:128-130)
        v.x = fadd(v.x, fmul(ex, dt)); v.y = fadd(v.y, fmul(ey, dt));
    } else {
        v.x = fmul(fadd(v.x, fmul(ex, dt)), il); v.y = fmul(fadd(v.y, fmul(ey, dt)), il);
    }
    vel[2 * s] = v;
}

// advance (World.hs:167-169; advanceObj Constraint.hs:225-229): pos' = (vel*dt) + pos,
// rot' = (dt*rotVel) + rot, then the rotation moveShapes will use (World.hs:132-140).
__global__ void __launch_bounds__(256) k_advance(int n, const uint8_t *alive, double dt, const double2 *vel,
                                                 double *pos_x, double *pos_y, double *rot, double *c, double *sn)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || !alive[s]) return;
    const double2 v = vel[2 * s];
    const double w = vel[2 * s + 1].x;
    pos_x[s] = fadd(fmul(v.x, dt), pos_x[s]);
    pos_y[s] = fadd(fmul(v.y, dt), pos_y[s]);
    const double r = fadd(fmul(dt, w), rot[s]);
    rot[s] = r;
    double cc, ss;
    shapes_sincos_inline(r, &cc, &ss);
    c[s] = cc; sn[s] = ss;
}

__global__ void __launch_bounds__(256) k_pack_body(int n, const double *inv_lin, const double *inv_rot, const double *mu,
                                                   const double *bounce, BodyRec *body)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    body[s] = BodyRec{ inv_lin[s], inv_rot[s], mu[s], bounce[s] };
}

// Contact rows, SoA columns -> 192 B records.  A warp transposes 32 rows through shared memory:
// coalesced column reads, coalesced record writes.
constexpr int PACK_WARPS = 4;
__global__ void __launch_bounds__(PACK_WARPS * 32) k_pack_rows(SolveParams S)
{
    __shared__ double tile[PACK_WARPS][32][ROW_DOUBLES + 1];
    const long long n_rows = S.st->n_contacts;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long n_tiles = (n_rows + 31) / 32;
    for (long long t = (long long)blockIdx.x * PACK_WARPS + wid; t < n_tiles; t += (long long)gridDim.x * PACK_WARPS) {
        const long long r = t * 32 + lane;
        double (*T)[ROW_DOUBLES + 1] = tile[wid];
        if (r < n_rows) {
#pragma unroll
            for (int k = 0; k < 6; ++k) { T[lane][k] = S.j_np[k][r]; T[lane][6 + k] = S.j_f[k][r]; }
            T[lane][12] = S.b_np[r]; T[lane][13] = S.inv_eff_np[r]; T[lane][14] = S.inv_eff_f[r];
            T[lane][15] = S.ra_x[r]; T[lane][16] = S.ra_y[r]; T[lane][17] = S.rb_x[r]; T[lane][18] = S.rb_y[r];
            T[lane][19] = S.rn_x[r]; T[lane][20] = S.rn_y[r];
            T[lane][21] = S.hit[r] ? 1.0 : 0.0; T[lane][22] = 0.0; T[lane][23] = 0.0;
        }
        __syncwarp();
        const long long base = t * 32 * ROW_DOUBLES;                 // doubles
        const long long limit = n_rows * ROW_DOUBLES;
        for (int e = lane; e < 32 * ROW_DOUBLES; e += 32)
            if (base + e < limit) S.rows[base + e] = T[e / ROW_DOUBLES][e % ROW_DOUBLES];
        __syncwarp();
    }
}

// ---- dependency chains -----------------------------------------------------------------------------
// Body b is touched, in list order, first by the pairs where it is the SMALLER key (pair_j == b:
// they belong to larger first keys, which come earlier in the descending list) and then by its own
// contiguous block of pairs (pair_i == b).  The first group is found by a stable radix sort of the
// pair indices on pair_j; pairs without contacts take no part.

__global__ void __launch_bounds__(256) k_chain_keys(SolveParams S)
{
    const long long n_pairs = S.st->n_pairs;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += (long long)gridDim.x * blockDim.x) {
        S.sort_key[0][q] = S.ccnt[q] > 0 ? (uint32_t)S.pair_j[q] : (uint32_t)S.n_slots;
        S.sort_val[0][q] = (uint32_t)q;
    }
}

__global__ void __launch_bounds__(256) k_chain_links(SolveParams S, const uint32_t *skey, const uint32_t *sval)
{
    const long long n_pairs = S.st->n_pairs;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += (long long)gridDim.x * blockDim.x) {
        // (a) position t of the pair_j-sorted order
        const uint32_t key = skey[t];
        if (key != (uint32_t)S.n_slots) {
            const uint32_t q = sval[t];
            const bool has_next = t + 1 < n_pairs && skey[t + 1] == key;
            S.next_j[q] = has_next ? (int32_t)sval[t + 1] : -1;
            if (t == 0 || skey[t - 1] != key) S.first_j[key] = (int32_t)q;
        }
        // (b) pair t in list order: next live pair of the same pair_i block
        if (S.ccnt[t] > 0) {
            const int i = S.pair_i[t];
            long long u = t + 1;
            while (u < n_pairs && S.pair_i[u] == i && S.ccnt[u] == 0) ++u;
            S.next_i[t] = (u < n_pairs && S.pair_i[u] == i) ? (int32_t)u : -1;
            long long d = t - 1;
            while (d >= 0 && S.pair_i[d] == i && S.ccnt[d] == 0) --d;
            if (d < 0 || S.pair_i[d] != i) S.first_i[i] = (int32_t)t;
        }
    }
}

__global__ void __launch_bounds__(256) k_chain_finish(SolveParams S)
{
    const long long n_pairs = S.st->n_pairs;
    SolveState *ss = S.ss;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += (long long)gridDim.x * blockDim.x) {
        if (S.ccnt[q] == 0) continue;
        const int i = S.pair_i[q], j = S.pair_j[q];
        const bool dyn_i = body_dynamic(S.mass[i]), dyn_j = body_dynamic(S.mass[j]);
        const int head_i = S.first_j[i] >= 0 ? S.first_j[i] : S.first_i[i];   // first node of body i's chain
        const int head_j = S.first_j[j] >= 0 ? S.first_j[j] : S.first_i[j];
        unsigned chains = 0;
        if (S.next_i[q] < 0) { S.next_i[q] = -1 - head_i; chains += dyn_i ? 1u : 0u; }   // the own block ends the chain
        if (S.next_j[q] < 0) {
            if (S.first_i[j] >= 0) S.next_j[q] = S.first_i[j];                  // on to body j's own block
            else { S.next_j[q] = -1 - head_j; chains += dyn_j ? 1u : 0u; }
        }
        const int deps = ((dyn_i && head_i != (int)q) ? 1 : 0) + ((dyn_j && head_j != (int)q) ? 1 : 0);
        S.cnt[q] = deps;
        NodeRec rec;
        rec.i = i; rec.j = j; rec.r0 = S.coff[q];
        rec.flags = (S.ccnt[q] & 3u) | (dyn_i ? 4u : 0u) | (dyn_j ? 8u : 0u);
        rec.next_i = S.next_i[q]; rec.next_j = S.next_j[q];
        rec.succ_r0_i = S.coff[rec.next_i >= 0 ? rec.next_i : -1 - rec.next_i];
        rec.succ_r0_j = S.coff[rec.next_j >= 0 ? rec.next_j : -1 - rec.next_j];
        S.node[q] = rec;
        if (chains) atomicAdd(&ss->expected, chains);
        if (deps == 0) {
            const unsigned long long slot = atomicAdd(&ss->tail, 1ull);
            if (slot < ss->queue_cap) S.queue[slot] = (unsigned)q | ((unsigned)S.p_begin << SOLVE_PASS_SHIFT);
            else ss->overflow = 1;
        }
    }
}

// ---- the solver -----------------------------------------------------------------------------------

// dotV6 (Template.hs:108-110)
__device__ __forceinline__ double dot6(const double *a, const double *b)
{
    double s = fmul(a[0], b[0]);
#pragma unroll
    for (int k = 1; k < 6; ++k) s = fadd(s, fmul(a[k], b[k]));
    return s;
}
// applyLagrangian (Constraint.hs:216-222): v_k + ((j_k * l) * im_k)
__device__ __forceinline__ void apply_lagrangian(double l, const double *j, const double *im, double *v)
{
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = fadd(v[k], fmul(fmul(j[k], l), im[k]));
}
// Ord Double's default min / max (GHC.Classes): max x y = if x <= y then y else x
__device__ __forceinline__ double hs_max(double x, double y) { return (x <= y) ? y : x; }
__device__ __forceinline__ double hs_min(double x, double y) { return (x <= y) ? x : y; }

__device__ __forceinline__ NodeRec load_node(const NodeRec *p)
{
    const int4 a = __ldg(reinterpret_cast<const int4 *>(p)), b = __ldg(reinterpret_cast<const int4 *>(p) + 1);
    NodeRec r;
    r.i = a.x; r.j = a.y; r.r0 = (unsigned)a.z; r.flags = (unsigned)a.w;
    r.next_i = b.x; r.next_j = b.y; r.succ_r0_i = (unsigned)b.z; r.succ_r0_j = (unsigned)b.w;
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// release fence: the velocity / lambda / counter stores before it are visible device-wide before
// the counter decrements after it
__device__ __forceinline__ void fence_release() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// Consumer side of the hand-over: a thread that learned through a relaxed atomic's return value (or a polled queue
// slot) that a node's predecessors are done must not read their results before this fence (PTX memory model:
// release fence + relaxed write -> relaxed read + acquire fence is the synchronising pattern).
__device__ __forceinline__ void fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// The same guarantee without a second fence on the hop: the decrement itself is an ACQUIRE read-modify-write (it reads
// from the release sequence headed by the other predecessor's decrement, which its release fence precedes), and a
// polled queue slot is read with an acquire load.  (A separate acquire fence after a relaxed atomic cost 36 -> 44 ms
// on the 1M-box pile; these cost nothing measurable: every mutable datum is read at L2 anyway.)
__device__ __forceinline__ int atomic_dec_acquire(int *p)
{
    int old;
    asm volatile("atom.acquire.gpu.global.add.s32 %0, [%1], -1;" : "=r"(old) : "l"(p) : "memory");
    return old;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }


// One contact row against the pair's velocities v (improveContactSln, Solvers/Contact.hs:124-143).
// R = the packed row (ROW_DOUBLES doubles).  Returns the new cached Lagrangians.
__device__ __forceinline__ void improve_row(const double2 *R, const double *im, double bounciness, double pair_mu,
                                            double cached_np, double cached_f, double *v, double *out_np, double *out_f)
{
    const double jn[6] = { R[0].x, R[0].y, R[1].x, R[1].y, R[2].x, R[2].y };
    const double jf[6] = { R[3].x, R[3].y, R[4].x, R[4].y, R[5].x, R[5].y };
    const double b_np = R[6].x, en = R[6].y, ef = R[7].x;
    const double rax = R[7].y, ray = R[8].x, rbx = R[8].y, rby = R[9].x, rnx = R[9].y, rny = R[10].x;
    // bounceB (Restitution.hs:34-47)
    const double nwa = -v[2];
    const double nwa_x = -fmul(nwa, ray), nwa_y = fmul(nwa, rax);       // zcrossV2 (Linear.hs:127-130)
    const double wb_x = -fmul(v[5], rby), wb_y = fmul(v[5], rbx);
    const double cv_x = fadd(fadd(fadd(-v[0], nwa_x), v[3]), wb_x);
    const double cv_y = fadd(fadd(fadd(-v[1], nwa_y), v[4]), wb_y);
    const double bounce_b = hs_min(0.0, fmul(bounciness, fadd(fmul(cv_x, rnx), fmul(cv_y, rny))));
    // contactLagrangian (Constraints/Contact.hs:87-97) = lagrangian2 (Constraint.hs:164-169) twice, both
    // from the velocities as read; effMassM2 is the velocity-independent value K3 computed
    const double new_np = fdiv(-fadd(dot6(jn, v), fadd(b_np, bounce_b)), en);
    const double new_f = fdiv(-fadd(dot6(jf, v), 0.0), ef);
    // solutionProcessor (Constraints/Contact.hs:99-110): positive, then clampAbs
    const double apply_np = hs_max(new_np, -cached_np);
    const double cache_np = fadd(cached_np, apply_np);
    const double max_thresh = fmul(cache_np, pair_mu), min_thresh = -max_thresh;
    const double accum = fadd(cached_f, new_f);
    const double accum2 = (accum > max_thresh) ? max_thresh : ((accum < min_thresh) ? min_thresh : accum);
    const double apply_f = fsub(accum2, cached_f);
    // applySln (Solvers/Contact.hs:54-65): non-penetration first, then friction
    apply_lagrangian(apply_np, jn, im, v);
    apply_lagrangian(apply_f, jf, im, v);
    *out_np = cache_np; *out_f = accum2;
}

// useCache (Solvers/Contact.hs:99-112): applySln with the cached ContactLagrangian, rows the join hit
__device__ __forceinline__ void cached_row(const double2 *R, const double *im, double l_np, double l_f, double *v)
{
    if (R[10].y == 0.0) return;                                           // newCache: nothing to apply
    const double jn[6] = { R[0].x, R[0].y, R[1].x, R[1].y, R[2].x, R[2].y };
    const double jf[6] = { R[3].x, R[3].y, R[4].x, R[4].y, R[5].x, R[5].y };
    apply_lagrangian(l_np, jn, im, v);
    apply_lagrangian(l_f, jf, im, v);
}

// Persistent dataflow executor.  A thread runs a node, publishes its results (release fence),
// decrements the dependency counters of the node's two successors (the next contact pair on body
// i, on body j), continues with a successor it made ready and queues the other one.  Idle threads
// each wait on their own queue slot.  No thread ever waits for a particular other thread, so
// residency and scheduling order cannot deadlock it.
//
// The critical path of a step is the longest dependency chain (thousands of nodes in a pile), so
// the loop is organised around the latency of ONE hop: every static datum of a node sits in one
// NodeRec sector and 192 B row records; the NodeRecs of both successors are fetched while the
// current node computes and their rows are pulled into L2, so that after the counter decrement
// only L2 hits (velocities of the other body, prefetched rows) separate a thread from the math.
// Mutable data (velocities, Lagrangians, counters, queue) is only ever accessed at L2
// (ld.cg / st.cg / atomics / volatile), never through the non-coherent L1.
__global__ void __launch_bounds__(SOLVE_THREADS) k_solve(SolveParams S)
{
    SolveState *ss = S.ss;
    const unsigned expected = ss->expected;
    if (expected == 0 || ss->overflow) return;
    const int lane = (int)(threadIdx.x & 31);
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool executor = lane < S.lanes;
    const double2 *rows = reinterpret_cast<const double2 *>(S.rows);
    const double2 *body = reinterpret_cast<const double2 *>(S.body);
    unsigned node = SOLVE_NONE;
    NodeRec rec = {};
    int pass = 0;
    long long my_slot = -1;
    unsigned backoff = 32;
    unsigned idle_iters = 0;
    unsigned long long ran = 0, idle_ns = 0;
    // The warp stays converged: every lane takes part in the votes of each iteration, lanes with a
    // node run it in lock step (their loads overlap), a lane that made two successors ready hands
    // the second one to an idle sibling through shuffles (no memory round trip) and only queues it
    // globally when the warp has no free lane.  Only a warp with no work at all sleeps -- a
    // sleeping lane in a busy warp would hold its siblings back.
    for (;;) {
        __syncwarp();
        // idle lanes: a claimed queue slot is polled once per iteration; a slot is claimed when the
        // whole warp is idle (see below) or after a few iterations without a sibling's hand-over
        if (node == SOLVE_NONE && executor) {
            if (my_slot < 0 && idle_iters >= S.claim_after) my_slot = (long long)atomicAdd(&ss->head, 1ull);
            if (my_slot >= 0) {
                unsigned e = SOLVE_NONE;
                if ((unsigned long long)my_slot < S.queue_cap) e = ld_acquire_u32(&S.queue[my_slot]);   // the pusher's results become visible with the slot
                if (e != SOLVE_NONE) {
                    node = e & SOLVE_NODE_MASK; pass = (int)(e >> SOLVE_PASS_SHIFT);
                    rec = load_node(&S.node[node]);
                    my_slot = -1;
                }
            }
        }
        const unsigned busy = __ballot_sync(0xffffffffu, node != SOLVE_NONE);
        if (busy == 0) {
            int stop = 0;
            if (lane == 0) {
                if (*reinterpret_cast<volatile unsigned *>(&ss->done) >= expected) stop = 1;
                else if (*reinterpret_cast<volatile int *>(&ss->overflow)) stop = 1;
                else {
                    // never spin forever: wall clock (%globaltimer) since this warp last had work; the host reports it
                    const unsigned long long now = global_ns();
                    if (idle_ns == 0) idle_ns = now;
                    else if (now - idle_ns > SOLVE_IDLE_LIMIT_NS) { ss->overflow = 2; stop = 1; }
                }
            }
            stop = __shfl_sync(0xffffffffu, stop, 0);
            if (stop) break;
            idle_iters = S.claim_after;                 // an idle warp listens to the global queue
            if (S.sleep_cap) __nanosleep(backoff);
            if (backoff < S.sleep_cap) backoff <<= 1;
            continue;
        }
        backoff = 32; idle_ns = 0;
        // what this lane offers to / continues with after the iteration
        unsigned spare = SOLVE_NONE; int spare_pass = 0; NodeRec spare_rec = {};
        if (node == SOLVE_NONE) ++idle_iters;
        else {
            idle_iters = 0;
            const unsigned q = node;
            const int i = rec.i, j = rec.j;
            const unsigned m = rec.flags & 3u, r0 = rec.r0;
            const bool dyn_i = (rec.flags & 4u) != 0, dyn_j = (rec.flags & 8u) != 0;
            // ---- every load of this node, issued together
            double2 R0[11], R1[11];
#pragma unroll
            for (int t = 0; t < 11; ++t) R0[t] = __ldg(&rows[(size_t)r0 * (ROW_DOUBLES / 2) + t]);
            if (m > 1) {
#pragma unroll
                for (int t = 0; t < 11; ++t) R1[t] = __ldg(&rows[(size_t)(r0 + 1) * (ROW_DOUBLES / 2) + t]);
            }
            const double2 a0 = __ldcg(&S.vel[2 * i]), a1 = __ldcg(&S.vel[2 * i + 1]);
            const double2 b0 = __ldcg(&S.vel[2 * j]), b1 = __ldcg(&S.vel[2 * j + 1]);
            const double2 mi = __ldg(&body[2 * i]), ui = __ldg(&body[2 * i + 1]);
            const double2 mj = __ldg(&body[2 * j]), uj = __ldg(&body[2 * j + 1]);
            double l0n = __ldcg(&S.lam_np[r0]), l0f = __ldcg(&S.lam_f[r0]), l1n = 0.0, l1f = 0.0;
            if (m > 1) { l1n = __ldcg(&S.lam_np[r0 + 1]); l1f = __ldcg(&S.lam_f[r0 + 1]); }
            // both successors' records (static), needed only after the math below
            const int li = rec.next_i, lj = rec.next_j;
            const unsigned si = (unsigned)(li >= 0 ? li : -1 - li), sj = (unsigned)(lj >= 0 ? lj : -1 - lj);
            const NodeRec rsi = load_node(&S.node[si]), rsj = load_node(&S.node[sj]);
            // ... and their rows / Lagrangians start moving towards L2 now, a whole node ahead of their use
            // (nothing should be in flight to DRAM when the release fence below drains the memory pipeline)
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!(side == 0 ? dyn_i : dyn_j)) continue;
                const unsigned sr = side == 0 ? rec.succ_r0_i : rec.succ_r0_j;
                const double2 *p = &rows[(size_t)sr * (ROW_DOUBLES / 2)];
                prefetch_l2(p); prefetch_l2(p + 8); prefetch_l2(p + 16);
                prefetch_l2(&S.lam_np[sr]); prefetch_l2(&S.lam_f[sr]);
            }
            // ---- the node
            double v[6] = { a0.x, a0.y, a1.x, b0.x, b0.y, b1.x };
            const double im[6] = { mi.x, mi.x, mi.y, mj.x, mj.x, mj.y };     // invMassM2 (Constraint.hs:118-120)
            if (pass == 0) {
                cached_row(R0, im, l0n, l0f, v);
                if (m > 1) cached_row(R1, im, l1n, l1f, v);
            } else {
                const double bounciness = hs_min(ui.y, uj.y);                    // uncurry min (Restitution.hs:47)
                const double pair_mu = fdiv(fadd(ui.x, uj.x), 2.0);              // pairMu (Friction.hs:46-48)
                improve_row(R0, im, bounciness, pair_mu, l0n, l0f, v, &l0n, &l0f);
                __stcg(&S.lam_np[r0], l0n); __stcg(&S.lam_f[r0], l0f);
                if (m > 1) {
                    improve_row(R1, im, bounciness, pair_mu, l1n, l1f, v, &l1n, &l1f);
                    __stcg(&S.lam_np[r0 + 1], l1n); __stcg(&S.lam_f[r0 + 1], l1f);
                }
            }
            if (dyn_i) { __stcg(&S.vel[2 * i], make_double2(v[0], v[1])); __stcg(&S.vel[2 * i + 1], make_double2(v[2], 0.0)); }
            if (dyn_j) { __stcg(&S.vel[2 * j], make_double2(v[3], v[4])); __stcg(&S.vel[2 * j + 1], make_double2(v[5], 0.0)); }
            ++ran;
            if (pass + 1 < S.p_end) S.cnt[q] = (dyn_i ? 1 : 0) + (dyn_j ? 1 : 0);   // re-arm for the next sweep
            fence_release();
            // both counter decrements are issued before either result is looked at (one L2 round trip, not two)
            const bool wrap_i = li < 0, wrap_j = lj < 0;
            const int sp_i = pass + (wrap_i ? 1 : 0), sp_j = pass + (wrap_j ? 1 : 0);
            const bool end_i = dyn_i && sp_i >= S.p_end, end_j = dyn_j && sp_j >= S.p_end;   // the body's chain is finished
            const bool go_i = dyn_i && !end_i, go_j = dyn_j && !end_j;
            int old_i = 0, old_j = 0;
            if (go_i) old_i = atomic_dec_acquire(&S.cnt[si]);
            if (go_j) old_j = atomic_dec_acquire(&S.cnt[sj]);
            if (end_i || end_j) atomicAdd(&ss->done, (end_i ? 1u : 0u) + (end_j ? 1u : 0u));
            const bool rdy_i = go_i && old_i == 1, rdy_j = go_j && old_j == 1;
            // (last arriver: the other predecessor's stores were published before its decrement, which this acquire RMW read from)
            // continue along one ready successor; a second one is offered to the warp
            node = SOLVE_NONE;
            if (rdy_i || rdy_j) {
                const bool take_i = rdy_i && (!rdy_j || S.prefer_i);
                node = take_i ? si : sj; pass = take_i ? sp_i : sp_j; rec = take_i ? rsi : rsj;
                if (rdy_i && rdy_j) { spare = take_i ? sj : si; spare_pass = take_i ? sp_j : sp_i; spare_rec = take_i ? rsj : rsi; }
            }
        }
        // ---- hand spare nodes to free sibling lanes (k-th spare to the k-th free lane); the rest is queued
        const unsigned spare_mask = __ballot_sync(0xffffffffu, spare != SOLVE_NONE);
        if (spare_mask) {
            const unsigned free_mask = __ballot_sync(0xffffffffu, executor && node == SOLVE_NONE && my_slot < 0);
            const int n_free = __popc(free_mask);
            const bool i_am_free = ((free_mask >> lane) & 1u) != 0;
            const int my_rank = i_am_free ? __popc(free_mask & lt_mask) : __popc(spare_mask & lt_mask);
            // a free lane of rank r reads from the r-th spare holder (if there is one)
            const int src = (i_am_free && my_rank < __popc(spare_mask)) ? (int)__fns(spare_mask, 0, my_rank + 1) : lane;
            const unsigned g_node = __shfl_sync(0xffffffffu, spare, src);
            const int g_pass = __shfl_sync(0xffffffffu, spare_pass, src);
            NodeRec g;
            g.i = __shfl_sync(0xffffffffu, spare_rec.i, src); g.j = __shfl_sync(0xffffffffu, spare_rec.j, src);
            g.r0 = __shfl_sync(0xffffffffu, spare_rec.r0, src); g.flags = __shfl_sync(0xffffffffu, spare_rec.flags, src);
            g.next_i = __shfl_sync(0xffffffffu, spare_rec.next_i, src); g.next_j = __shfl_sync(0xffffffffu, spare_rec.next_j, src);
            g.succ_r0_i = __shfl_sync(0xffffffffu, spare_rec.succ_r0_i, src); g.succ_r0_j = __shfl_sync(0xffffffffu, spare_rec.succ_r0_j, src);
            if (i_am_free && src != lane) { node = g_node; pass = g_pass; rec = g; idle_iters = 0; }
            else if (spare != SOLVE_NONE && my_rank >= n_free) {
                const unsigned long long slot = atomicAdd(&ss->tail, 1ull);
                if (slot < S.queue_cap) st_volatile_u32(&S.queue[slot], spare | ((unsigned)spare_pass << SOLVE_PASS_SHIFT));
                else ss->overflow = 1;
            }
        }
    }
    if (ran) atomicAdd(&ss->nodes_run, ran);
}

} // namespace

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------

struct WorldStep {
    bool uploaded = false;
    int64_t n = 0;
    double *col[9] = {};           // pos_x, pos_y, rot, cos, sin, inv_lin, inv_rot, mu, bounce
    double2 *vel = nullptr;
    double *tmp[3] = {};           // unpack scratch
    int32_t *next_i = nullptr, *next_j = nullptr, *first_i = nullptr, *first_j = nullptr, *cnt = nullptr;
    NodeRec *node = nullptr;
    double *rows = nullptr;
    BodyRec *body = nullptr;
    uint32_t *sort_key[2] = {}, *sort_val[2] = {};
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    uint32_t *queue = nullptr;
    size_t queue_alloc = 0;        // slots allocated
    size_t queue_cap = 0;          // slots this step may use (pairs x sweeps + threads)
    SolveState *ss = nullptr;
    SolveState *h_ss = nullptr;    // pinned
    cudaEvent_t ev[5] = {};
    int solve_blocks = 0;
    int lanes = 32;
    unsigned sleep_cap = 1024;
    int prefer_i = 1;
    unsigned claim_after = 0;
    int64_t steps = 0;
};

namespace {

constexpr int WORLD_MAX_SWEEPS = 1 + 14;   // the pass number travels in 4 bits of a queue entry

int world_alloc(shapes_ctx *c)
{
    if (c->ws) return SHAPES_OK;
    if (c->world != 1) { c->err = "shapes_world_*: single-GPU ctx only (the device-resident world does not shard yet)"; return SHAPES_E_ARG; }
    if (c->max_pairs >= (int64_t)SOLVE_NODE_MASK) { c->err = "shapes_world_*: max_pairs must stay below 2^28"; return SHAPES_E_ARG; }
    CU_TRY(c, cudaSetDevice(c->device));
    WorldStep *w = new WorldStep();
    c->ws = w;
    const size_t N = (size_t)std::max<int64_t>(c->max_shapes, 1), P = (size_t)std::max<int64_t>(c->max_pairs, 1);
    int rc;
#define WS_ALLOC(ptr, count) do { rc = dev_alloc(c, ptr, count); if (rc != SHAPES_OK) return rc; } while (0)
    for (int k = 0; k < 9; ++k) WS_ALLOC(&w->col[k], N);
    for (int k = 0; k < 3; ++k) WS_ALLOC(&w->tmp[k], N);
    WS_ALLOC(&w->vel, 2 * N);
    WS_ALLOC(&w->next_i, P); WS_ALLOC(&w->next_j, P); WS_ALLOC(&w->cnt, P);
    WS_ALLOC(&w->node, P);
    WS_ALLOC(&w->rows, (size_t)std::max<int64_t>(c->max_contacts, 1) * ROW_DOUBLES);
    WS_ALLOC(&w->body, N);
    WS_ALLOC(&w->first_i, N); WS_ALLOC(&w->first_j, N);
    for (int k = 0; k < 2; ++k) { WS_ALLOC(&w->sort_key[k], P); WS_ALLOC(&w->sort_val[k], P); }
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(nullptr, w->sort_tmp_bytes, w->sort_key[0], w->sort_key[1], w->sort_val[0],
                                              w->sort_val[1], (int)P, 0, 32, c->stream));
    WS_ALLOC(reinterpret_cast<uint8_t **>(&w->sort_tmp), w->sort_tmp_bytes);
    int per_sm = 2;
    if (const char *e = std::getenv("SHAPES_B200_SOLVE_BLOCKS_PER_SM")) per_sm = std::max(1, std::atoi(e));
    if (const char *e = std::getenv("SHAPES_B200_SOLVE_LANES")) w->lanes = std::min(32, std::max(1, std::atoi(e)));
    if (const char *e = std::getenv("SHAPES_B200_SOLVE_CLAIM_AFTER")) w->claim_after = (unsigned)std::max(0, std::atoi(e));
    if (const char *e = std::getenv("SHAPES_B200_SOLVE_PREFER_I")) w->prefer_i = std::atoi(e) ? 1 : 0;
    if (const char *e = std::getenv("SHAPES_B200_SOLVE_SLEEP")) w->sleep_cap = (unsigned)std::max(0, std::atoi(e));
    int occ = 0;
    CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_solve, SOLVE_THREADS, 0));
    per_sm = std::min(per_sm, std::max(occ, 1));
    w->solve_blocks = c->sm_count * per_sm;
    // every node is queued at most once per sweep, every idle thread holds at most one claimed slot;
    // sized for updateWorld's 3 sweeps here, re-allocated by shapes_world_step when a step needs more
    w->queue_alloc = P * (size_t)3 + (size_t)w->solve_blocks * SOLVE_THREADS + 64;
    WS_ALLOC(&w->queue, w->queue_alloc);
    WS_ALLOC(&w->ss, 1);
#undef WS_ALLOC
    CU_TRY(c, cudaMallocHost(&w->h_ss, sizeof(SolveState)));
    for (int k = 0; k < 5; ++k) CU_TRY(c, cudaEventCreate(&w->ev[k]));
    return SHAPES_OK;
}

// shapes_grow: the pair / contact sized buffers of the device-resident world (all of them per-step scratch; the body
// columns, velocities and materials are sized by slots and stay where they are).
int world_grow(shapes_ctx *c, bool pairs, bool contacts)
{
    WorldStep *w = c->ws;
    if (!w) return SHAPES_OK;
    const size_t P = (size_t)std::max<int64_t>(c->max_pairs, 1);
    auto fresh = [&](auto **p, size_t count) -> int {
        using T = std::remove_pointer_t<std::remove_pointer_t<decltype(p)>>;
        T *q = nullptr;
        CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&q), count * sizeof(T)));
        for (void *&a : c->allocs) if (a == static_cast<void *>(*p)) { cudaFree(a); a = q; }
        *p = q;
        return SHAPES_OK;
    };
#define WS_FRESH(ptr, count) do { int rc__ = fresh(ptr, count); if (rc__ != SHAPES_OK) return rc__; } while (0)
    if (pairs) {
        WS_FRESH(&w->next_i, P); WS_FRESH(&w->next_j, P); WS_FRESH(&w->cnt, P); WS_FRESH(&w->node, P);
        for (int k = 0; k < 2; ++k) { WS_FRESH(&w->sort_key[k], P); WS_FRESH(&w->sort_val[k], P); }
        size_t tb = 0;
        CU_TRY(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, w->sort_key[0], w->sort_key[1], w->sort_val[0], w->sort_val[1],
                                                  (int)P, 0, 32, c->stream));
        if (tb > w->sort_tmp_bytes) {
            uint8_t *tmp = static_cast<uint8_t *>(w->sort_tmp);
            WS_FRESH(&tmp, tb);
            w->sort_tmp = tmp; w->sort_tmp_bytes = tb;
        }
    }
    if (contacts) WS_FRESH(&w->rows, (size_t)std::max<int64_t>(c->max_contacts, 1) * ROW_DOUBLES);
#undef WS_FRESH
    return SHAPES_OK;
}

void world_free(shapes_ctx *c)
{
    WorldStep *w = c->ws;
    if (!w) return;
    if (w->h_ss) cudaFreeHost(w->h_ss);
    for (int k = 0; k < 5; ++k) if (w->ev[k]) cudaEventDestroy(w->ev[k]);
    delete w;
    c->ws = nullptr;
}

} // namespace

extern "C" {

void shapes_sincos(int64_t n, const double *rot, double *cos_out, double *sin_out)
{
    for (int64_t k = 0; k < n; ++k) shapes_sincos_inline(rot[k], &cos_out[k], &sin_out[k]);
}

int shapes_world_upload(shapes_ctx *c, int64_t n_slots, const double *vel_x, const double *vel_y, const double *rot_vel,
                        const double *pos_x, const double *pos_y, const double *rot,
                        const double *cos_rot, const double *sin_rot,
                        const double *inv_lin, const double *inv_rot, const double *mu, const double *bounce)
{
    if (!c) return SHAPES_E_ARG;
    if (!c->hulls_set || n_slots != c->n_slots) { c->err = "shapes_world_upload: n_slots differs from shapes_set_hulls"; return SHAPES_E_ARG; }
    if (n_slots > 0 && (!vel_x || !vel_y || !rot_vel || !pos_x || !pos_y || !rot || !inv_lin || !inv_rot || !mu || !bounce ||
                        ((cos_rot == nullptr) != (sin_rot == nullptr)))) {
        c->err = "shapes_world_upload: missing column";
        return SHAPES_E_ARG;
    }
    int rc = world_alloc(c);
    if (rc != SHAPES_OK) return rc;
    WorldStep *w = c->ws;
    cudaStream_t s = c->stream;
    const size_t bytes = sizeof(double) * (size_t)n_slots;
    const int N = (int)n_slots;
    if (N > 0) {
        const double *src[9] = { pos_x, pos_y, rot, cos_rot, sin_rot, inv_lin, inv_rot, mu, bounce };
        for (int k = 0; k < 9; ++k)
            if (src[k]) CU_TRY(c, cudaMemcpyAsync(w->col[k], src[k], bytes, cudaMemcpyHostToDevice, s));
        const double *vsrc[3] = { vel_x, vel_y, rot_vel };
        for (int k = 0; k < 3; ++k) CU_TRY(c, cudaMemcpyAsync(w->tmp[k], vsrc[k], bytes, cudaMemcpyHostToDevice, s));
        k_pack_vel<<<grid_for(N, 256, 1 << 30), 256, 0, s>>>(N, w->tmp[0], w->tmp[1], w->tmp[2], w->vel); ++c->launches;
        k_pack_body<<<grid_for(N, 256, 1 << 30), 256, 0, s>>>(N, w->col[5], w->col[6], w->col[7], w->col[8], w->body); ++c->launches;
        if (!cos_rot) { k_sincos<<<grid_for(N, 256, 1 << 30), 256, 0, s>>>(N, w->col[2], w->col[3], w->col[4]); ++c->launches; }
        CU_TRY(c, cudaGetLastError());
    }
    CU_TRY(c, cudaStreamSynchronize(s));
    w->uploaded = true; w->n = n_slots; w->steps = 0;
    // a new world state starts with an empty EngineCache (initEngine, Engine/Main.hs:43-46)
    c->have_frame = false; c->cache_valid = false; c->plan_valid = false;
    return SHAPES_OK;
}

int shapes_world_download(shapes_ctx *c, int64_t n_slots, double *vel_x, double *vel_y, double *rot_vel,
                          double *pos_x, double *pos_y, double *rot, double *cos_rot, double *sin_rot)
{
    if (!c) return SHAPES_E_ARG;
    WorldStep *w = c->ws;
    if (!w || !w->uploaded || n_slots != w->n) { c->err = "shapes_world_download: no uploaded world of that size"; return SHAPES_E_ARG; }
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t bytes = sizeof(double) * (size_t)n_slots;
    const int N = (int)n_slots;
    if (N > 0) {
        if (vel_x || vel_y || rot_vel) {
            k_unpack_vel<<<grid_for(N, 256, 1 << 30), 256, 0, s>>>(N, w->vel, w->tmp[0], w->tmp[1], w->tmp[2]); ++c->launches;
            double *dst[3] = { vel_x, vel_y, rot_vel };
            for (int k = 0; k < 3; ++k) if (dst[k]) CU_TRY(c, cudaMemcpyAsync(dst[k], w->tmp[k], bytes, cudaMemcpyDeviceToHost, s));
        }
        double *dst[5] = { pos_x, pos_y, rot, cos_rot, sin_rot };
        for (int k = 0; k < 5; ++k) if (dst[k]) CU_TRY(c, cudaMemcpyAsync(dst[k], w->col[k], bytes, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(c, cudaStreamSynchronize(s));
    return SHAPES_OK;
}

int shapes_world_step(shapes_ctx *c, const shapes_step_config *cfg, shapes_step_stats *stats)
{
    if (!c || !cfg) return SHAPES_E_ARG;
    WorldStep *w = c->ws;
    if (!w || !w->uploaded) { c->err = "shapes_world_step: shapes_world_upload has not been called"; return SHAPES_E_ARG; }
    if (cfg->solver_iterations < 0 || cfg->solver_iterations > WORLD_MAX_SWEEPS - 1 ||
        cfg->external_kind < SHAPES_EXT_NONE || cfg->external_kind > SHAPES_EXT_FORCE) {
        c->err = "shapes_world_step: bad configuration";
        return SHAPES_E_ARG;
    }
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int N = (int)w->n;
    Params &P = c->P;
    if (!cfg->warm_start) c->cache_valid = false;
    CU_TRY(c, cudaEventRecord(w->ev[0], s));
    // culledKeys + prepareFrame + constraintGen (+ the cache join) on the shapes as last moved
    const double *in[7] = { w->col[0], w->col[1], w->col[2], w->col[3], w->col[4], w->col[5], w->col[6] };
    shapes_frame_out fo;
    std::memset(&fo, 0, sizeof(fo));
    int rc = run_frame(c, w->n, in, cfg->dt, cfg->baumgarte, cfg->slop, false, &fo);
    if (stats) { std::memset(stats, 0, sizeof(*stats)); stats->n_pairs = fo.n_pairs; stats->n_contacts = fo.n_contacts; stats->frame_ms = fo.device_ms; }
    if (rc != SHAPES_OK) return rc;            // capacity: nothing of the world has been touched
    const bool warm = c->warm_done;
    const int64_t n_pairs = c->last_pairs, n_contacts = c->last_contacts;
    CU_TRY(c, cudaEventRecord(w->ev[1], s));
    if (!warm && n_contacts > 0) {             // newCache everywhere: ContactLagrangian 0 0 (Solvers/Contact.hs:85-97)
        CU_TRY(c, cudaMemsetAsync(P.warm_np, 0, sizeof(double) * (size_t)n_contacts, s));
        CU_TRY(c, cudaMemsetAsync(P.warm_f, 0, sizeof(double) * (size_t)n_contacts, s));
        CU_TRY(c, cudaMemsetAsync(P.warm_hit, 0, (size_t)n_contacts, s));
    }
    c->warm_done = true;                       // warm_np / warm_f / warm_hit now describe this frame
    // applyExternal (before the solver touches the velocities, Engine/Main.hs:76)
    if (N > 0 && cfg->external_kind != SHAPES_EXT_NONE) {
        k_external<<<grid_for(N, 256, 1 << 30), 256, 0, s>>>(N, c->d_alive, cfg->external_kind, cfg->external_x, cfg->external_y,
                                                              cfg->dt, w->col[5], w->vel);
        ++c->launches;
    }
    SolveParams S;
    std::memset(&S, 0, sizeof(S));
    S.st = P.st; S.pair_i = P.pair_i; S.pair_j = P.pair_j; S.ccnt = P.ccnt; S.coff = P.coff;
    for (int q = 0; q < 6; ++q) { S.j_np[q] = P.j_np[q]; S.j_f[q] = P.j_f[q]; }
    S.b_np = P.b_np; S.ra_x = P.ra_x; S.ra_y = P.ra_y; S.rb_x = P.rb_x; S.rb_y = P.rb_y; S.rn_x = P.rn_x; S.rn_y = P.rn_y;
    S.inv_eff_np = P.inv_eff_np; S.inv_eff_f = P.inv_eff_f;
    S.hit = P.warm_hit; S.lam_np = P.warm_np; S.lam_f = P.warm_f;
    S.vel = w->vel; S.mass = P.mass; S.mu = w->col[7]; S.bounce = w->col[8];
    S.node = w->node; S.rows = w->rows; S.body = w->body;
    S.next_i = w->next_i; S.next_j = w->next_j; S.first_i = w->first_i; S.first_j = w->first_j; S.cnt = w->cnt;
    for (int k = 0; k < 2; ++k) { S.sort_key[k] = w->sort_key[k]; S.sort_val[k] = w->sort_val[k]; }
    S.queue = w->queue; S.ss = w->ss; S.n_slots = N;
    S.p_begin = warm ? 0 : 1; S.p_end = 1 + cfg->solver_iterations;
    S.lanes = w->lanes; S.sleep_cap = w->sleep_cap; S.queue_cap = w->queue_cap; S.prefer_i = w->prefer_i; S.claim_after = w->claim_after;
    const bool solve = n_contacts > 0 && n_pairs > 0 && S.p_begin < S.p_end;
    if (solve) {
        const int sms = c->sm_count;
        SolveState init;
        std::memset(&init, 0, sizeof(init));
        // queue slots this step can need: one per (sweep, pair) + one claimed slot per thread
        const size_t need = (size_t)n_pairs * (size_t)(S.p_end - S.p_begin) + (size_t)w->solve_blocks * SOLVE_THREADS + 64;
        if (need > w->queue_alloc) {
            uint32_t *bigger = nullptr;
            CU_TRY(c, cudaStreamSynchronize(s));
            CU_TRY(c, cudaMalloc(&bigger, sizeof(uint32_t) * need));
            for (void *&q : c->allocs) if (q == w->queue) { cudaFree(q); q = bigger; }
            w->queue = bigger; w->queue_alloc = need; S.queue = bigger;
        }
        w->queue_cap = need; S.queue_cap = need;
        init.queue_cap = w->queue_cap;
        *w->h_ss = init;
        CU_TRY(c, cudaMemcpyAsync(w->ss, w->h_ss, sizeof(SolveState), cudaMemcpyHostToDevice, s));
        CU_TRY(c, cudaMemsetAsync(w->queue, 0xff, sizeof(uint32_t) * w->queue_cap, s));
        CU_TRY(c, cudaMemsetAsync(w->first_i, 0xff, sizeof(int32_t) * (size_t)std::max(N, 1), s));
        CU_TRY(c, cudaMemsetAsync(w->first_j, 0xff, sizeof(int32_t) * (size_t)std::max(N, 1), s));
        k_chain_keys<<<sms * 8, 256, 0, s>>>(S); ++c->launches;
        int bits = 1;
        while ((1ll << bits) <= (long long)N) ++bits;          // keys are 0..N (N = pairs without contacts)
        size_t tb = w->sort_tmp_bytes;
        cub::DoubleBuffer<uint32_t> dk(w->sort_key[0], w->sort_key[1]), dv(w->sort_val[0], w->sort_val[1]);
        CU_TRY(c, cub::DeviceRadixSort::SortPairs(w->sort_tmp, tb, dk, dv, (int)n_pairs, 0, bits, s));
        k_chain_links<<<sms * 8, 256, 0, s>>>(S, dk.Current(), dv.Current()); ++c->launches;
        k_chain_finish<<<sms * 8, 256, 0, s>>>(S); ++c->launches;
        k_pack_rows<<<sms * 8, PACK_WARPS * 32, 0, s>>>(S); ++c->launches;
    }
    CU_TRY(c, cudaEventRecord(w->ev[2], s));
    if (solve) { k_solve<<<w->solve_blocks, SOLVE_THREADS, 0, s>>>(S); ++c->launches; }
    CU_TRY(c, cudaEventRecord(w->ev[3], s));
    // advance + the rotation moveShapes will use (Engine/Main.hs:84-85)
    if (N > 0) {
        k_advance<<<grid_for(N, 256, 1 << 30), 256, 0, s>>>(N, c->d_alive, cfg->dt, w->vel, w->col[0], w->col[1], w->col[2], w->col[3], w->col[4]);
        ++c->launches;
    }
    // this frame's Lagrangians are the next frame's EngineCache (Engine/Main.hs:32,60-68)
    if (n_contacts > 0) {
        CU_TRY(c, cudaMemcpyAsync(c->d_cache_np, P.warm_np, sizeof(double) * (size_t)n_contacts, cudaMemcpyDeviceToDevice, s));
        CU_TRY(c, cudaMemcpyAsync(c->d_cache_f, P.warm_f, sizeof(double) * (size_t)n_contacts, cudaMemcpyDeviceToDevice, s));
    }
    if (solve) CU_TRY(c, cudaMemcpyAsync(w->h_ss, w->ss, sizeof(SolveState), cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaEventRecord(w->ev[4], s));
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaStreamSynchronize(s));
    c->cache_valid = true;
    ++w->steps;
    if (solve && (w->h_ss->overflow || w->h_ss->done != w->h_ss->expected)) {
        c->err = "shapes_world_step: solver did not complete (internal error)";
        return SHAPES_E_CUDA;
    }
    if (stats) {
        float ms = 0.f;
        CU_TRY(c, cudaEventElapsedTime(&ms, w->ev[1], w->ev[2])); stats->chains_ms = ms;
        CU_TRY(c, cudaEventElapsedTime(&ms, w->ev[2], w->ev[3])); stats->solve_ms = ms;
        CU_TRY(c, cudaEventElapsedTime(&ms, w->ev[3], w->ev[4])); stats->integrate_ms = ms;
        CU_TRY(c, cudaEventElapsedTime(&ms, w->ev[0], w->ev[4])); stats->total_ms = ms;
        stats->solver_nodes = solve ? (int64_t)w->h_ss->nodes_run : 0;
        stats->queue_pushes = solve ? (int64_t)w->h_ss->tail : 0;
        stats->body_chains = solve ? (int64_t)w->h_ss->expected : 0;
        stats->warm = warm ? 1 : 0;
    }
    return SHAPES_OK;
}

} // extern "C"

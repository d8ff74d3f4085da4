// shapes_b200.cu -- B200 (sm_100a) collision pipeline behind include/shapes_b200.h.
//
// Replaces, for one frame, the reference's
//   Aabb.culledKeys / Grid.culledKeys   (shapes/src/Physics/Broadphase/Aabb.hs:168-183, Grid.hs:67-100)
//   prepareFrame                        (shapes/src/Physics/Solvers/Contact.hs:40-52)
//   constraintGen per contact           (shapes/src/Physics/Constraints/Contact.hs:60-72)
// All reference paths below are relative to /root/reference/.
//
// Pipeline of one frame on one GPU (one stream, one CUDA graph, no host round trip between kernels):
//   K0  k_begin_frame + k_transform_aabb   moveShapes + toAabb per slot (World.hs:132-140, Aabb.hs:81-110); the grid was
//       planned from the PREVIOUS frame's bounds, so K0 also computes each shape's cell key and bins it (stale plans
//       are detected on the device and the frame is re-seeded inside the call)
//   K1  counting sort on the cell table: exclusive scan of the histogram + k_scatter_sorted
//   K2  boxes-only worlds: k_sweep<count> -> exclusive scan over slots in DESCENDING key order -> k_sweep<emit>;
//       general polygon worlds ("sorted mode"): ONE sweep pass that also appends every pair to a work list in grid-CELL
//       order; (+ k_big<count/emit> for shapes spanning more than 2 cells or with non-finite bounds)
//       => pairs come out in the reference's descending (i, j) order by construction (a pair's place is
//       off[rank of i] + its rank among i's partners).
//   K3a k_manifolds<4> (boxes, one thread per pair) / k_manifolds_coop (general polygons, 16 lanes per pair, walks the
//       cell-ordered work list so that neighbouring tiles share their hulls): SAT both directions + incident-edge
//       clipping -> per pair a contact count and a 64 B manifold record
//   K3b exclusive scan of the counts + k_row_map, then k_rows: one lane per contact row -- NonPenetration / Friction /
//       Restitution generators + inverse effective masses, every column store a whole number of 32 B sectors
//   (+ k_warm_join: descZipVector of this frame's keys against the previous frame's Lagrangian cache)
// Several GPUs: "rows mode", the k_rw_* kernels further down (the comment there describes the exchange).
//
// Arithmetic: IEEE binary64, every operation a separately rounded __dmul_rn/__dadd_rn/... so no
// FMA contraction can occur whatever the compiler flags; expression trees follow the reference's
// TH-generated left folds (shapes-math/src/Shapes/Linear/Template.hs:108-110).
//
// There is no CPU fallback anywhere in this file.

#include "shapes_b200.h"


#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <nccl.h>   // types only: NCCL is bound at run time (see NcclApi), never at link time
#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// small device math layer (Physics.Linear, shapes/src/Physics/Linear.hs)
// ---------------------------------------------------------------------------------------------

struct V2 { double x, y; };

__device__ __forceinline__ double fmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fdiv(double a, double b) { return __ddiv_rn(a, b); }

// dotV2 (Template.hs:108-110): (a0*b0)+(a1*b1)
__device__ __forceinline__ double dot2(V2 a, V2 b) { return fadd(fmul(a.x, b.x), fmul(a.y, b.y)); }
// minusV2 (Linear.hs:114-116)
__device__ __forceinline__ V2 sub2(V2 a, V2 b) { return V2{ fsub(a.x, b.x), fsub(a.y, b.y) }; }
// crossV2 (Linear.hs:118-120)
__device__ __forceinline__ double cross2(V2 a, V2 b) { return fsub(fmul(a.x, b.y), fmul(a.y, b.x)); }
// clockwiseV2 (Linear.hs:161-163)
__device__ __forceinline__ V2 clockwise2(V2 a) { return V2{ a.y, -a.x }; }
__device__ __forceinline__ V2 neg2(V2 a) { return V2{ -a.x, -a.y }; }

// unitEdgeNormal (ConvexHull.hs:218-226) = normalizeV2 (clockwiseV2 (v' - v)) (Linear.hs:165-168)
__device__ __forceinline__ V2 unit_edge_normal(V2 v, V2 vnext)
{
    V2 e = clockwise2(sub2(vnext, v));
    double len = __dsqrt_rn(fadd(fmul(e.x, e.x), fmul(e.y, e.y)));
    return V2{ fdiv(e.x, len), fdiv(e.y, len) };
}

// Forward matrix of toTransform pos ori = translate(pos) . rotate(ori)
// (Transform.hs:34-38,73-77; Linear.hs:349-381), first two rows. Every entry is the generic
// mul3x3x3 dot product ((t0*r0)+(t1*r1))+(t2*r2) (MatrixTemplate.hs:47-67) so that signed zeros
// and non-finite inputs behave as in the reference.
struct Aff { double m00, m01, m02, m10, m11, m12; };

__device__ __forceinline__ Aff to_transform(double px, double py, double c, double s)
{
    const double ns = -s;
    Aff m;
    m.m00 = fadd(fadd(fmul(1.0, c), fmul(0.0, s)), fmul(px, 0.0));
    m.m01 = fadd(fadd(fmul(1.0, ns), fmul(0.0, c)), fmul(px, 0.0));
    m.m02 = fadd(fadd(fmul(1.0, 0.0), fmul(0.0, 0.0)), fmul(px, 1.0));
    m.m10 = fadd(fadd(fmul(0.0, c), fmul(1.0, s)), fmul(py, 0.0));
    m.m11 = fadd(fadd(fmul(0.0, ns), fmul(1.0, c)), fmul(py, 0.0));
    m.m12 = fadd(fadd(fmul(0.0, 0.0), fmul(1.0, 0.0)), fmul(py, 1.0));
    return m;
}

// afmul (Linear.hs:217-220): t `mul3x3c` (a, b, 1.0)
__device__ __forceinline__ V2 afmul(const Aff &m, V2 p)
{
    return V2{ fadd(fadd(fmul(m.m00, p.x), fmul(m.m01, p.y)), fmul(m.m02, 1.0)),
               fadd(fadd(fmul(m.m10, p.x), fmul(m.m11, p.y)), fmul(m.m12, 1.0)) };
}

// boundsOverlap (Aabb.hs:69-72): not (c > b || d < a); NaN bounds therefore "overlap".
__device__ __forceinline__ bool bounds_overlap(double a, double b, double c, double d)
{
    return !((c > b) || (d < a));
}

__device__ __forceinline__ void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct __align__(32) Box { double min_x, max_x, min_y, max_y; };
struct __align__(32) Xf { double px, py, c, s; };

// aabbCheck boxA boxB (Aabb.hs:75-78); A is the shape with the larger key.
__device__ __forceinline__ bool aabb_check(const Box &a, const Box &b)
{
    return bounds_overlap(a.min_x, a.max_x, b.min_x, b.max_x) &&
           bounds_overlap(a.min_y, a.max_y, b.min_y, b.max_y);
}

__device__ __forceinline__ bool finite4(const Box &b)
{
    return isfinite(b.min_x) && isfinite(b.max_x) && isfinite(b.min_y) && isfinite(b.max_y);
}

// order-preserving map double -> uint64 for atomicMin/atomicMax
__device__ __forceinline__ unsigned long long enc_ordered(double d)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_ordered(unsigned long long u)
{
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}

constexpr int SHAPES_MAX_RANKS = 16;
constexpr int ROW_BINS = 4096;       // rows mode: resolution of the per-row work histogram that balances the row cuts
enum { RW_PHASE_SEED = 0, RW_PHASE_KEYS = 1, RW_PHASE_CNT = 2, RW_PHASE_OFF = 3, RW_PHASE_RESULTS = 4, RW_PHASE_COUNTS = 5, RW_PHASES = 6 };
constexpr int ERR_PAIR_CAP = 1;
constexpr int ERR_PEER_TIMEOUT = 4;
constexpr int ERR_CONTACT_CAP = 2;
constexpr int ERR_REPLAN = 8;       // plan-ahead: too many shapes fell outside the planned grid -- re-seed and redo the frame
constexpr int MAX_STAGED_VERTS = 8; // hulls up to this many vertices are staged in shared memory

// Device-resident per-frame state: lets every kernel be launched without a host round trip.
struct FrameState {
    unsigned long long bmin_x, bmin_y, bmax_x, bmax_y; // ordered encodings of the finite world bounds
    double ox, oy, h;   // grid origin and (possibly coarsened) cell edge
    int W, H;           // grid columns / rows
    unsigned n_cells;
    unsigned n_big;     // shapes on the big-shape path
    unsigned n_small;   // shapes in the grid
    int error;
    long long n_pairs;
    long long n_contacts;
    unsigned long long work_cursor; // sorted mode: next free entry of the cell-ordered SAT work list
    unsigned long long n_pairs_hit; // pairs with at least one contact (statistics: the roofline accounting)
    // rows mode (multi-rank): the frame counter lives here so that frames replay as CUDA graphs
    unsigned long long frame_no;
    int row_lo, row_hi;         // grid rows this rank sweeps, [lo, hi)
    unsigned cell_lo, cell_end; // cells it keeps: rows [row_lo - 1, row_hi + 1); every other mode: [0, n_cells)
    int cut[SHAPES_MAX_RANKS + 1]; // row cuts of all ranks: rank g sweeps rows [cut[g], cut[g + 1])
    unsigned n_kept;            // shapes in this rank's grid
    unsigned n_list;            // rows mode: entries of the kept-slot list (k_rw_bin appends, k_rw_hulls walks it)
    unsigned push_cursor[SHAPES_MAX_RANKS];   // rows mode, home side: records appended to every rank's inbox this frame
    int peer_error;             // OR of every rank's error word (exchanged at the results barrier)
    unsigned long long loc_fold, loc_contig;   // rows mode: work entries whose home would be this rank under the folded /
                                               // the contiguous home layout (decides which one the next frames use)
};

struct Params;
__device__ __forceinline__ bool slot_static(const Params &P, int s);
__device__ __forceinline__ Xf slot_xf(const Params &P, int s);
__device__ __forceinline__ double2 slot_mass(const Params &P, int s);

// Clipped contact manifold of one pair, handed from k_manifolds to k_rows.
struct __align__(64) ManRec {
    double nx, ny;        // unit normal of the penetrated edge
    double ref_d;         // n . (first vertex of the penetrated edge)
    double c0x, c0y, c1x, c1y; // manifold points, descending feature index
    unsigned long long bits;   // edge [0,20) | pen0 [20,40) | pen1 [40,60) | flip [60]
};

// Rows mode: what a sweeping rank delivers for one pair, into the arena of the pair's home: a 16 B header in a DENSE
// array (the home unpacks the pair columns and the contact counts from it with fully coalesced reads) and, when the
// pair has contacts, a 96 B body written as one request -- instead of up to nine scattered stores.
struct __align__(16) PairHdr { int32_t i, j; uint32_t cnt, pad; };
struct __align__(128) PairRec {
    ManRec man;            // the manifold
    double4 pj;            // (pos_j, inverse masses of j)
    double pad[4];
};
static_assert(sizeof(PairRec) == 128 && sizeof(PairHdr) == 16, "pair record layout");

// Rows mode: what a home sends to a rank that sweeps one of its shapes -- one 96 B record, appended densely to the
// receiver's inbox (a warp's records leave as whole 128 B lines, not as three scattered stores per shape).
struct __align__(32) HomeRec {
    Xf xf;                 // packed transform
    Box box;               // the AABB its home folded: the sweep does not wait for the sweeper's own hull pass
    double2 mass;          // inverse masses
    uint32_t slot, key;    // key: RW_KEY_* encoding, bit 31 = isStatic
    uint32_t pad[2];
};
static_assert(sizeof(HomeRec) == 96, "HomeRec layout");
constexpr int HOME_REC_CHUNKS = (int)(sizeof(HomeRec) / 16);

// Everything the kernels need, passed by value.
struct Params {
    // static world
    int n_slots;
    int own_lo, own_hi;         // slots this rank owns as the larger key of a pair
    const uint8_t *alive;
    const int32_t *vert_offset;
    const double2 *local;       // interleaved local vertices
    const int32_t *ext_min, *ext_max;   // per edge (hull-relative)
    const unsigned long long *ext_packed; // per slot, 3+3 bits per edge, hulls with <= 8 vertices
    const double *radius;       // per slot: >= 0 = CircleShape of that radius (no vertices), < 0 = hull; NULL = no circles
    double2 *circ;              // world centre of the owned circle slots (setCircleTransform)
    // per-frame inputs
    const double *pos_x, *pos_y, *rot, *cos_rot, *sin_rot, *inv_lin, *inv_rot;
    double dt, baumgarte, slop;
    // per-frame derived
    Xf *xf;                     // (px, py, cos, sin)
    double2 *mass;              // (inv_lin, inv_rot)
    Box *box;                   // per slot
    double *world_x, *world_y;  // optional debug output
    double2 *wv, *wn;           // world vertices / unit edge normals of the OWNED slots (moveShapes result)
    const uint4 *hh;            // per slot, static: CSR offset, vertex count, packed extents (lo, hi) -- one 16 B gather per hull
    // Sorted mode (general polygon worlds): the SAT stage visits the pairs in grid-CELL order through a work
    // list written by the (single pass) sweep, so the hulls a tile of pairs touches are shared with the
    // neighbouring tiles whatever the host's slot numbering is; results go to the pair's place in the
    // reference order, off[r(i)] + a.
    int sorted_mode;
    uint32_t *w_i, *w_j, *w_a;  // SAT work list, cell order: both slots and the pair's rank among i's partners
    uint32_t *keys, *keys_sorted; // cell key per slot / per sorted position
    uint32_t *rank;             // per slot: arrival order within its cell (counting sort)
    Box *sbox;                  // AABB records in sorted order
    uint32_t *smeta;            // slot | static << 31, sorted order
    uint32_t *cell_count;       // per cell: shapes binned this frame
    uint32_t *cell_begin;       // exclusive scan of cell_count: cell c = sorted positions [begin[c], begin[c+1])
    uint8_t *cell_mark;         // multi-rank: cells inside the 3x3 neighbourhood of an owned shape
    int multi_rank;
    int plan_ahead;             // 1 = the grid was planned from the PREVIOUS frame's bounds (k_begin_frame): K0 keys and bins
    unsigned big_limit;         // plan-ahead: more big-list entries than this = the plan is stale (ERR_REPLAN)
    unsigned cell_cap;
    unsigned cell_limit;        // cells the planner may use this frame (<= cell_cap); what the scan covers
    uint32_t key_none;          // sort key of slots outside the grid (dead / big): first value past the cell table
    double cell_size;
    uint32_t *big_idx;
    unsigned long long *rank_bounds; // world x 4 ordered-uint bounds (multi-rank)
    // peer-to-peer exchange (multi-rank, CUDA IPC): every rank's buffers, indexed by rank (own included)
    Box *peer_box[SHAPES_MAX_RANKS];
    uint32_t *peer_keys[SHAPES_MAX_RANKS];
    const double *peer_in[7][SHAPES_MAX_RANKS]; // every rank's body columns (host API: each rank uploads only its own slots)
    int remote_inputs;          // 1 = body columns of foreign slots are read through peer_in
    uint32_t *gkeys;            // cell key of EVERY slot (own: computed here; others: pushed by their owner)
    int chunk;                  // slots per rank
    unsigned long long *peer_bounds[SHAPES_MAX_RANKS];
    unsigned long long *peer_flags[SHAPES_MAX_RANKS];
    unsigned long long *flags;  // this rank's flag words, written by the peers (one per rank)
    int n_peers;                // 0 = exchange through NCCL
    int my_rank;
    unsigned long long frame_no;
    unsigned long long *cnt, *off; // per query slot, indexed own_hi-1-i
    unsigned long long *hitmask;   // per sorted position: k_sweep<count>'s hits, replayed by k_sweep<emit>
    int64_t max_pairs, max_contacts;
    int32_t *pair_i, *pair_j;
    // contacts
    ManRec *man;                // per pair: clipped manifold (written only when it has contacts)
    uint32_t *ccnt, *coff;      // per pair: contact count, exclusive row offset
    uint32_t *row_map;          // per contact row: pair << 1 | manifold point
    int32_t *key_i, *key_j, *feat_a, *feat_b;
    uint8_t *flip;
    double *normal_x, *normal_y, *center_x, *center_y, *depth;
    double *j_np[6], *b_np, *ra_x, *ra_y, *rb_x, *rb_y, *rn_x, *rn_y, *j_f[6];
    double *inv_eff_np, *inv_eff_f;
    // warm start (descZipVector): previous frame's keys + the host solver's Lagrangian cache for them
    const int32_t *pk_i, *pk_j, *pk_fa, *pk_fb;
    const double *cache_np, *cache_f;
    const long long *n_prev;    // device word, set outside the captured graph (the count changes every frame)
    double *warm_np, *warm_f;
    uint8_t *warm_hit;
    FrameState *st;
    // ---- rows mode (multi-rank with mapped peers): the SWEEP / SAT work is partitioned by grid rows (cuts balanced by
    // the pair counts the rows produced in the previous frame), results are delivered to the slot-range HOME of the
    // pair's larger key, so the global order stays "rank G-1's rows, then G-2's, ..." with no merge.
    int work_mode;              // 0: SAT walks the pairs in reference order; 1: cell-ordered work list, results at
                                // off[r(i)] + a; 2: rows mode -- results stay in work order on the sweeping rank
    uint32_t *sat_ccnt;         // rows mode: per work entry, the SAT stage's contact count (local copy: marks the pairs
                                // left to the per-thread pass); otherwise = ccnt
    uint32_t *q_off;            // rows mode, per slot i this rank sweeps: first index of i's pairs in its home's arrays
    uint32_t *roww;             // rows mode: pairs produced per row bin this frame [ROW_BINS]
    uint32_t *mat_stamp;        // rows mode, per slot: frame in which k_rw_hulls materialised its world vertices / normals
    uint32_t *kept_list;        // rows mode: the slots this rank keeps, in (roughly) ascending slot order
    HomeRec *rw_inbox[SHAPES_MAX_RANKS];     // every rank's inbox [G sources][inbox_cap]; I append to section my_rank
    unsigned *rw_inbox_cnt[SHAPES_MAX_RANKS];// every rank's [G] record counts of its inbox sections (written at the KEYS barrier)
    const HomeRec *inbox;                    // mine
    const unsigned *inbox_cnt;
    int inbox_cap;                           // records per section
    int dbg_local_stores;       // experiment (SHAPES_B200_DBG_LOCAL_STORES): SAT results stay on the sweeping rank -- WRONG results, timing only
    // homes: the slot space is cut into 2G blocks of rw_blk slots, rank g is home to blocks g and 2G-1-g.  Whatever the
    // host's numbering, each home then holds the same number of slots AND (the larger key of a pair being uniform or
    // linear in the slot index) the same number of pairs; its slice of the result is two runs of the global order.
    // (rw_fold 0: one contiguous block per rank, the high block is empty -- chosen after the first frames when the slot
    // numbering turns out to follow the geometry, so that a slot's home is also the rank that sweeps it and its results
    // stay on the GPU; the decision is taken from counters every rank holds, so all ranks switch in the same frame.)
    int rw_blk, rw_fold;
    int rw_lo_lo, rw_lo_hi, rw_hi_lo, rw_hi_hi;   // my low block [lo_lo, lo_hi), my high block [hi_lo, hi_hi)
    uint32_t *rw_cq[SHAPES_MAX_RANKS];           // every rank's per-slot word, pushed by the sweeping rank: rank << 28 | partner count
    uint32_t *rw_qoff[SHAPES_MAX_RANKS];         // every rank's q_off (pushed by the homes after their scan)
    PairRec *rw_prec[SHAPES_MAX_RANKS];          // every rank's pair records (home side): the sweeping rank stores each pair
                                                 // straight into its final place
    PairRec *prec;                               // mine
    PairHdr *rw_phdr[SHAPES_MAX_RANKS];          // ... and their dense headers
    PairHdr *phdr;
    double2 *rw_mass[SHAPES_MAX_RANKS];          // every rank's inverse masses (its home slots are valid)
    Xf *rw_xf[SHAPES_MAX_RANKS];             // every rank's packed transforms (its own slot range is valid)
    uint32_t *rw_weights[SHAPES_MAX_RANKS];  // every rank's [G][ROW_BINS] inbox of row weights (this frame's parity)
    const uint32_t *rw_weights_prev;         // my inbox of the previous frame
    const unsigned long long *rw_bounds_prev;// bounds every rank pushed in the previous frame [4 G]
    long long *rw_counts[SHAPES_MAX_RANKS];  // every rank's [2 G] inbox of (pairs, contacts)
    int *rw_err[SHAPES_MAX_RANKS];           // every rank's [G] inbox of error words
};

// Body column k (0 pos_x, 1 pos_y, 2 rot, 3 cos, 4 sin, 5 inv_lin, 6 inv_rot) of ANY slot.  When every
// rank uploaded only its own slots (shapes_frame with the peer exchange) the value of a foreign slot
// is pulled from its owner's column over NVLink.
__device__ __forceinline__ double in_col(const Params &P, int k, const double *local, int s)
{
    if (P.remote_inputs) return P.peer_in[k][s / P.chunk][s];
    return local[s];
}
// isStatic (Constraint.hs:123-125), straight from the host's inverse-mass columns
__device__ __forceinline__ bool slot_static(const Params &P, int s)
{
    return in_col(P, 5, P.inv_lin, s) == 0.0 && in_col(P, 6, P.inv_rot, s) == 0.0;
}
// (px, py, cos, sin): K0 packs it for the rank's own slots; other slots are read from the raw columns
__device__ __forceinline__ int rw_home(const Params &P, int s);
__device__ __forceinline__ Xf slot_xf(const Params &P, int s)
{
    if (P.work_mode == 2) return P.rw_xf[rw_home(P, s)][s];      // rows mode: its home's record (mine: local)
    if (s >= P.own_lo && s < P.own_hi) return P.xf[s];
    double c, sn;
    if (P.cos_rot) { c = in_col(P, 3, P.cos_rot, s); sn = in_col(P, 4, P.sin_rot, s); }
    else sincos(in_col(P, 2, P.rot, s), &sn, &c);
    return Xf{ in_col(P, 0, P.pos_x, s), in_col(P, 1, P.pos_y, s), c, sn };
}
__device__ __forceinline__ double2 slot_mass(const Params &P, int s)
{
    if (P.work_mode == 2) return P.rw_mass[rw_home(P, s)][s];
    if (s >= P.own_lo && s < P.own_hi) return P.mass[s];
    return make_double2(in_col(P, 5, P.inv_lin, s), in_col(P, 6, P.inv_rot, s));
}

// AABB record of any slot: with the peer exchange a rank's box array holds only its own slots and
// the records of other ranks are PULLED through the peer pointers (NVLink loads) where needed.
__device__ __forceinline__ Box box_of(const Params &P, int s)
{
    if (P.work_mode == 2) return P.peer_box[rw_home(P, s)][s];
    if (P.n_peers > 0) return P.peer_box[s / P.chunk][s];
    return P.box[s];
}

constexpr uint32_t KEY_STATIC_BIT = 0x80000000u;   // rows mode: pushed cell keys carry isStatic in bit 31
// rows mode key encoding (so that a cleared key array reads "nothing here"): 0 = no cell, 1 = big-shape path, cell + 2
constexpr uint32_t RW_KEY_NONE = 0u, RW_KEY_BIG = 1u, RW_KEY_BASE = 2u;

__device__ __forceinline__ int rw_home(const Params &P, int s)
{
    const int b = s / P.rw_blk;
    return (!P.rw_fold || b < P.n_peers) ? b : 2 * P.n_peers - 1 - b;
}
__device__ __forceinline__ bool rw_mine(const Params &P, int s)
{
    return (s >= P.rw_lo_lo && s < P.rw_lo_hi) || (s >= P.rw_hi_lo && s < P.rw_hi_hi);
}
// my home slots in the order my slice lists them: descending through the high block, then through the low block
__device__ __forceinline__ int rw_qslot(const Params &P, int r)
{
    const int n_hi = P.rw_hi_hi - P.rw_hi_lo;
    return r < n_hi ? P.rw_hi_hi - 1 - r : P.rw_lo_hi - 1 - (r - n_hi);
}

// ---------------------------------------------------------------------------------------------
// K0: moveShapes + toAabb
// ---------------------------------------------------------------------------------------------

__global__ void k_set_i64(long long *p, long long v) { *p = v; }

__global__ void k_reset_state(FrameState *st)
{
    st->bmin_x = st->bmin_y = ~0ull;
    st->bmax_x = st->bmax_y = 0ull;
    st->ox = st->oy = 0.0;
    st->h = 1.0;
    st->W = st->H = 1;
    st->n_cells = 1;
    st->n_big = 0;
    st->n_small = 0;
    st->error = 0;
    st->n_pairs = 0;
    st->n_contacts = 0;
    st->work_cursor = 0ull;
    st->n_pairs_hit = 0ull;
    st->peer_error = 0;
    st->cell_lo = 0u; st->cell_end = 1u; st->row_lo = 0; st->row_hi = 1;
}

// Choose origin, cell edge and grid extent for finite world bounds [bmin, bmax] (ordered-uint encodings in st).
// The cell edge starts at the static estimate (largest hull diameter outside the big set) and doubles until the
// table fits.  `margin` cells are added on every side (plan-ahead: shapes may have moved since the bounds were taken;
// whatever still falls outside goes to the exact big-shape path, so results never depend on the plan).
__device__ void plan_grid_from_bounds(FrameState *st, double cell_size, unsigned cell_limit, double margin)
{
    if (st->bmax_x == 0ull) { // no finite shape
        st->ox = st->oy = 0.0; st->h = cell_size; st->W = st->H = 1; st->n_cells = 1;
        st->cell_lo = 0u; st->cell_end = 1u; st->row_lo = 0; st->row_hi = 1;
        return;
    }
    const double bx = dec_ordered(st->bmin_x), by = dec_ordered(st->bmin_y);
    const double ex = dec_ordered(st->bmax_x) - bx, ey = dec_ordered(st->bmax_y) - by;
    double h = cell_size;
    double wx = 0.0, wy = 0.0;
    bool ok = false;
    for (int it = 0; it < 2200; ++it) {
        wx = floor(ex / h) + 1.0 + 2.0 * margin;
        wy = floor(ey / h) + 1.0 + 2.0 * margin;
        if (isfinite(wx) && isfinite(wy) && wx < 1073741824.0 && wy < 1073741824.0 &&
            wx * wy <= (double)cell_limit) { ok = true; break; }
        h *= 2.0;
    }
    if (!ok) { wx = wy = 1.0; h = INFINITY; margin = 0.0; } // every finite shape lands in cell (0,0) or the big set
    const double ox = bx - margin * h, oy = by - margin * h;
    st->ox = isfinite(ox) ? ox : bx; st->oy = isfinite(oy) ? oy : by; st->h = h;
    st->W = (int)wx; st->H = (int)wy;
    st->n_cells = (unsigned)(st->W * st->H);
    st->cell_lo = 0u; st->cell_end = st->n_cells;
    st->row_lo = 0; st->row_hi = st->H;
}

// Plan-ahead frames (single rank): first kernel of the frame, one thread.  The bounds K0 reduced in the PREVIOUS
// frame (or the bounds-only pass after shapes_set_hulls) become this frame's grid, two cells of margin around
// them; then the per-frame counters are reset.  K0 can therefore key and bin every shape as it computes its AABB:
// no bounds -> plan -> keys dependency inside the frame.
__global__ void k_begin_frame(Params P)
{
    FrameState *st = P.st;
    plan_grid_from_bounds(st, P.cell_size, P.cell_limit, 2.0);
    st->bmin_x = st->bmin_y = ~0ull;
    st->bmax_x = st->bmax_y = 0ull;
    st->n_big = 0;
    st->n_small = 0;
    st->error = 0;
    st->n_pairs = 0;
    st->n_contacts = 0;
    st->work_cursor = 0ull;
    st->n_pairs_hit = 0ull;
    st->peer_error = 0;
}

__device__ __forceinline__ double warp_min(double v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One thread per OWNED slot: packs the body state later kernels gather (xf, mass), transforms the
// hull, stores world vertices + recomputed unit edge normals, folds the AABB and reduces the
// finite world bounds of the rank's shapes.
// moveShape (World.hs:132-134) -> setHullTransform (ConvexHull.hs:184-195): world vertex =
// afmul (toTransform pos rot) local; hullToAabb (Aabb.hs:81-84) = foldl1 mergeAabb with
// mergeRange's `if a < c then a else c` / `if b > d then b else d` (Aabb.hs:104-110).
__device__ __forceinline__ bool small_cell(const Box &b, const FrameState *st, int &cx, int &cy);

// BOUNDS_ONLY: nothing is stored -- the pass that seeds the plan-ahead grid after shapes_set_hulls.
template <bool BOUNDS_ONLY>
__global__ void __launch_bounds__(256) k_transform_aabb(Params P, int lo, int hi)
{
    double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int s = lo + blockIdx.x * blockDim.x + threadIdx.x; s < hi; s += gridDim.x * blockDim.x) {
        if (BOUNDS_ONLY) {
            if (!P.alive[s]) continue;
            const double px = P.pos_x[s], py = P.pos_y[s];
            double c, sn;
            if (P.cos_rot) { c = P.cos_rot[s]; sn = P.sin_rot[s]; }
            else sincos(P.rot[s], &sn, &c);
            const Aff m = to_transform(px, py, c, sn);
            const int o = P.vert_offset[s], n = P.vert_offset[s + 1] - o;
            const double rad = P.radius ? P.radius[s] : -1.0;
            Box b{ 0.0, 0.0, 0.0, 0.0 };
            if (rad >= 0.0) {
                const V2 ctr = afmul(m, V2{ 0.0, 0.0 });
                b.min_x = fsub(ctr.x, rad); b.max_x = fadd(ctr.x, rad);
                b.min_y = fsub(ctr.y, rad); b.max_y = fadd(ctr.y, rad);
            }
            for (int k = 0; k < n; ++k) {
                const double2 l = __ldg(&P.local[o + k]);
                const V2 w = afmul(m, V2{ l.x, l.y });
                if (k == 0) { b.min_x = b.max_x = w.x; b.min_y = b.max_y = w.y; }
                else {
                    b.min_x = (b.min_x < w.x) ? b.min_x : w.x; b.max_x = (b.max_x > w.x) ? b.max_x : w.x;
                    b.min_y = (b.min_y < w.y) ? b.min_y : w.y; b.max_y = (b.max_y > w.y) ? b.max_y : w.y;
                }
            }
            if (finite4(b)) {
                mnx = fmin(mnx, b.min_x); mxx = fmax(mxx, b.max_x);
                mny = fmin(mny, b.min_y); mxy = fmax(mxy, b.max_y);
            }
            continue;
        }
        // level 1: everything indexed by the slot, requested before anything is consumed
        const double px = P.pos_x[s], py = P.pos_y[s];
        const double il = P.inv_lin[s], ir = P.inv_rot[s];
        const bool live = P.alive[s] != 0;
        const int o = P.vert_offset[s];
        const int n = P.vert_offset[s + 1] - o;
        const double rad = P.radius ? P.radius[s] : -1.0;
        double c, sn;
        if (P.cos_rot) { c = P.cos_rot[s]; sn = P.sin_rot[s]; }
        else sincos(P.rot[s], &sn, &c); // not bit-exact against libm (documented at the ABI)
        P.xf[s] = Xf{ px, py, c, sn };
        P.mass[s] = make_double2(il, ir);
        if (!live) { if (P.plan_ahead) P.keys[s] = P.key_none; continue; }
        const Aff m = to_transform(px, py, c, sn);
        Box b;
        if (rad >= 0.0) {
            // setCircleTransform (Circle.hs:55-59): centre = transform applied to the local origin;
            // circleToAabb (Aabb.hs:86-88)
            const V2 ctr = afmul(m, V2{ 0.0, 0.0 });
            P.circ[s] = make_double2(ctr.x, ctr.y);
            b.min_x = fsub(ctr.x, rad); b.max_x = fadd(ctr.x, rad);
            b.min_y = fsub(ctr.y, rad); b.max_y = fadd(ctr.y, rad);
        }
        if (n <= MAX_STAGED_VERTS) {
            // level 2: the hull's local vertices as ONE batch of loads; world vertices stay in registers for the
            // normals (the r1 kernel re-read them from global memory after storing them, one dependent step per vertex)
            double2 l[MAX_STAGED_VERTS];
#pragma unroll
            for (int k = 0; k < MAX_STAGED_VERTS; ++k) if (k < n) l[k] = __ldg(&P.local[o + k]);
            V2 w[MAX_STAGED_VERTS];
#pragma unroll
            for (int k = 0; k < MAX_STAGED_VERTS; ++k) {
                if (k >= n) break;
                w[k] = afmul(m, V2{ l[k].x, l[k].y });
                P.wv[o + k] = make_double2(w[k].x, w[k].y);
                if (P.world_x) { P.world_x[o + k] = w[k].x; P.world_y[o + k] = w[k].y; }
                if (k == 0) { b.min_x = b.max_x = w[k].x; b.min_y = b.max_y = w[k].y; }
                else {
                    b.min_x = (b.min_x < w[k].x) ? b.min_x : w[k].x;
                    b.max_x = (b.max_x > w[k].x) ? b.max_x : w[k].x;
                    b.min_y = (b.min_y < w[k].y) ? b.min_y : w[k].y;
                    b.max_y = (b.max_y > w[k].y) ? b.max_y : w[k].y;
                }
            }
            // setHullTransform (ConvexHull.hs:193-194): unit edge normals recomputed from the NEW vertices
#pragma unroll
            for (int k = 0; k < MAX_STAGED_VERTS; ++k) {
                if (k >= n) break;
                const V2 nxt = (k + 1 < MAX_STAGED_VERTS && k + 1 < n) ? w[(k + 1) & (MAX_STAGED_VERTS - 1)] : w[0];
                const V2 nn = unit_edge_normal(w[k], nxt);
                P.wn[o + k] = make_double2(nn.x, nn.y);
            }
        } else {
            for (int k = 0; k < n; ++k) {
                double2 l = __ldg(&P.local[o + k]);
                V2 w = afmul(m, V2{ l.x, l.y });
                P.wv[o + k] = make_double2(w.x, w.y);
                if (P.world_x) { P.world_x[o + k] = w.x; P.world_y[o + k] = w.y; }
                if (k == 0) { b.min_x = b.max_x = w.x; b.min_y = b.max_y = w.y; }
                else {
                    b.min_x = (b.min_x < w.x) ? b.min_x : w.x;
                    b.max_x = (b.max_x > w.x) ? b.max_x : w.x;
                    b.min_y = (b.min_y < w.y) ? b.min_y : w.y;
                    b.max_y = (b.max_y > w.y) ? b.max_y : w.y;
                }
            }
            double2 v0 = P.wv[o], va = v0;
            for (int k = 0; k < n; ++k) {
                const double2 vb = (k + 1 < n) ? P.wv[o + k + 1] : v0;
                const V2 nn = unit_edge_normal(V2{ va.x, va.y }, V2{ vb.x, vb.y });
                P.wn[o + k] = make_double2(nn.x, nn.y);
                va = vb;
            }
        }
        P.box[s] = b;
        if (finite4(b)) {
            mnx = fmin(mnx, b.min_x); mxx = fmax(mxx, b.max_x);
            mny = fmin(mny, b.min_y); mxy = fmax(mxy, b.max_y);
        }
        if (P.plan_ahead) {
            // K0b/K1a fused (the grid of this frame was planned before the frame started): cell key of the min corner,
            // histogram of the cell table (the arrival order is the counting sort's scatter slot); shapes outside the
            // planned grid, spanning more than 2 cells or with non-finite bounds go to the big list
            int cx, cy;
            uint32_t key = P.key_none;
            if (small_cell(b, P.st, cx, cy)) {
                key = (uint32_t)cy * (uint32_t)P.st->W + (uint32_t)cx;
                P.rank[s] = atomicAdd(&P.cell_count[key], 1u);
            } else {
                const unsigned pos = atomicAdd(&P.st->n_big, 1u);
                P.big_idx[pos] = (uint32_t)s;
                if (pos >= P.big_limit) atomicOr(&P.st->error, ERR_REPLAN);   // a stale plan must not turn into an O(N^2) frame
            }
            P.keys[s] = key;
        }
    }
    // block-level reduction, then one set of atomics per block
    __shared__ double s_red[4][8];
    mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][warp] = mnx; s_red[1][warp] = mny; s_red[2][warp] = mxx; s_red[3][warp] = mxy; }
    __syncthreads();
    if (warp == 0) {
        mnx = lane < 8 ? s_red[0][lane] : INFINITY; mny = lane < 8 ? s_red[1][lane] : INFINITY;
        mxx = lane < 8 ? s_red[2][lane] : -INFINITY; mxy = lane < 8 ? s_red[3][lane] : -INFINITY;
        mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
        if (lane == 0 && mnx <= mxx) {
            atomicMin(&P.st->bmin_x, enc_ordered(mnx));
            atomicMin(&P.st->bmin_y, enc_ordered(mny));
            atomicMax(&P.st->bmax_x, enc_ordered(mxx));
            atomicMax(&P.st->bmax_y, enc_ordered(mxy));
        }
    }
}

// Peer exchange barrier, arrive side (phase 0: after K0, also carries this rank's bounds; phase 1:
// after the cell keys were pushed): raise this frame's flag in every peer.  The peer stores of
// earlier kernels of this stream are ordered before the flag by the system-scope fence.
__global__ void k_publish_peers(Params P, int phase)
{
    const FrameState *st = P.st;
    const int r = threadIdx.x;
    if (r >= P.n_peers) return;
    if (phase == 0) {
        unsigned long long *dst = P.peer_bounds[r] + 4 * P.my_rank;
        dst[0] = st->bmin_x; dst[1] = st->bmin_y; dst[2] = st->bmax_x; dst[3] = st->bmax_y;
    }
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(P.peer_flags[r] + phase * SHAPES_MAX_RANKS + P.my_rank) = P.frame_no;
}

// Peer exchange, step 3: wait until every rank has raised this frame's flag here (their records
// and bounds are then visible).  Bounded spin: a missing peer turns into an error, not a hang.
__global__ void k_wait_peers(Params P, int phase)
{
    const int r = threadIdx.x;
    if (r >= P.n_peers) return;
    const volatile unsigned long long *flag = P.flags + phase * SHAPES_MAX_RANKS + r;
    const long long t0 = clock64();
    while (*flag < P.frame_no) {
        if (clock64() - t0 > 8000000000ll) { atomicOr(&P.st->error, ERR_PEER_TIMEOUT); break; } // ~4 s
        __nanosleep(200);
    }
    __threadfence_system();
}

// Multi-rank: publish this rank's bounds (still in ordered-uint form) for the bounds all-gather.
__global__ void k_publish_bounds(Params P, int rank)
{
    const FrameState *st = P.st;
    unsigned long long *dst = P.rank_bounds + 4 * rank;
    dst[0] = st->bmin_x; dst[1] = st->bmin_y; dst[2] = st->bmax_x; dst[3] = st->bmax_y;
}

// ---------------------------------------------------------------------------------------------
// K0b: world bounds, grid plan, cell keys
// ---------------------------------------------------------------------------------------------

// Single thread: choose origin, cell edge and grid extent.  The cell edge starts at the static
// estimate (largest hull diameter outside the big set) and doubles until the table fits.
__global__ void k_plan_grid(Params P, int world)
{
    FrameState *st = P.st;
    for (int r = 0; r < world && world > 1; ++r) { // merge every rank's bounds (all-gathered)
        const unsigned long long *b = P.rank_bounds + 4 * r;
        if (b[0] < st->bmin_x) st->bmin_x = b[0];
        if (b[1] < st->bmin_y) st->bmin_y = b[1];
        if (b[2] > st->bmax_x) st->bmax_x = b[2];
        if (b[3] > st->bmax_y) st->bmax_y = b[3];
    }
    plan_grid_from_bounds(st, P.cell_size, P.cell_limit, 0.0);
}

// Cell of an AABB's min corner if the box spans at most 2 cells per axis ("small"), else -1.
// Monotonicity of x -> floor((x - ox) / h) alone guarantees that two overlapping small boxes have
// min-corner cells at most 1 apart per axis, so the 3x3 neighbourhood search is exhaustive.
__device__ __forceinline__ bool small_cell(const Box &b, const FrameState *st, int &cx, int &cy)
{
    if (!finite4(b)) return false;
    const double ox = st->ox, oy = st->oy, h = st->h;
    const double x0 = floor((b.min_x - ox) / h), x1 = floor((b.max_x - ox) / h);
    const double y0 = floor((b.min_y - oy) / h), y1 = floor((b.max_y - oy) / h);
    if (!(isfinite(x0) && isfinite(x1) && isfinite(y0) && isfinite(y1))) return false;
    if (!(x1 - x0 <= 1.0 && y1 - y0 <= 1.0 && x0 >= 0.0 && y0 >= 0.0 && x0 < (double)st->W && y0 < (double)st->H)) return false;
    cx = (int)x0; cy = (int)y0;
    return true;
}

__global__ void __launch_bounds__(256) k_clear_cells(Params P)
{
    const unsigned n = P.st->n_cells;
    for (unsigned c = blockIdx.x * blockDim.x + threadIdx.x; c <= n; c += gridDim.x * blockDim.x) {
        P.cell_count[c] = 0u;
        if (P.multi_rank && c < n) P.cell_mark[c] = 0;
    }
}

// K1a: cell key of the slots in [lo, hi) from their AABB records; with the peer exchange the 4 B key
// is the only thing PUSHED to the other ranks (key_none = not in any grid, key_none + 1 = big shape).
// Multi-rank: owned small shapes also mark the cells their partners can live in.
__global__ void __launch_bounds__(256) k_keys(Params P, int lo, int hi)
{
    const FrameState *st = P.st;
    const int W = st->W, H = st->H;
    for (int s = lo + blockIdx.x * blockDim.x + threadIdx.x; s < hi; s += gridDim.x * blockDim.x) {
        uint32_t key = P.key_none;
        if (P.alive[s]) {
            int cx, cy;
            if (small_cell(P.box[s], st, cx, cy)) {
                key = (uint32_t)cy * (uint32_t)W + (uint32_t)cx;
                if (P.multi_rank && s >= P.own_lo && s < P.own_hi)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int nx = cx + dx, ny = cy + dy;
                            if (nx >= 0 && nx < W && ny >= 0 && ny < H) P.cell_mark[(size_t)ny * W + nx] = 1;
                        }
            } else key = P.key_none + 1u;
        }
        if (P.n_peers > 0) { for (int r = 0; r < P.n_peers; ++r) P.peer_keys[r][s] = key; }
        else P.gkeys[s] = key;
    }
}

// K1a': histogram of the cell table over ALL slots' keys (the shape's arrival rank in its cell is
// the counting sort's scatter slot); big shapes go to the big list and are tested against
// everything; multi-rank: shapes of other ranks outside the marked cells are left out.
__global__ void __launch_bounds__(256) k_bin(Params P)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < P.n_slots; s += gridDim.x * blockDim.x) {
        uint32_t key = P.gkeys[s];
        if (key == P.key_none + 1u) {
            const unsigned pos = atomicAdd(&P.st->n_big, 1u);
            P.big_idx[pos] = (uint32_t)s;
            key = P.key_none;
        } else if (key < P.key_none) {
            const bool own = s >= P.own_lo && s < P.own_hi;
            if (!P.multi_rank || own || P.cell_mark[key]) P.rank[s] = atomicAdd(&P.cell_count[key], 1u);
            else key = P.key_none;
        }
        P.keys[s] = key;
    }
}

// K1b: scatter into cell order (after the exclusive scan of the histogram): AABB records, slot ids
// and keys in sorted order, contiguous per cell and per grid row.
__global__ void __launch_bounds__(256) k_scatter_sorted(Params P)
{
    if (P.work_mode == 2) {      // rows mode: the kept list of inbox records is what this rank walks (dense reads)
        const unsigned n_list = P.st->n_list;
        for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n_list; q += gridDim.x * blockDim.x) {
            const HomeRec &r = P.inbox[P.kept_list[q]];
            const uint32_t key = (r.key & ~KEY_STATIC_BIT) - RW_KEY_BASE;
            const uint32_t p = P.cell_begin[key] + P.rank[q];
            P.smeta[p] = r.slot | ((r.key & KEY_STATIC_BIT) ? 0x80000000u : 0u);
            P.keys_sorted[p] = key;
            P.sbox[p] = r.box;            // the AABB its home folded
        }
        return;
    }
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < P.n_slots; s += gridDim.x * blockDim.x) {
        const uint32_t key = P.keys[s];
        if (key >= P.key_none) continue;
        const uint32_t p = P.cell_begin[key] + P.rank[s];
        if (P.work_mode != 2) P.sbox[p] = box_of(P, s);     // rows mode: k_rw_hulls folds the record at the sorted position
        const bool st_flag = P.work_mode == 2 ? (P.gkeys[s] & KEY_STATIC_BIT) != 0u : slot_static(P, s);
        P.smeta[p] = (uint32_t)s | ((uint32_t)st_flag << 31);
        P.keys_sorted[p] = key;
    }
}

// ---------------------------------------------------------------------------------------------
// K2: grid sweep.  Query = every owned small shape i; candidates = shapes j < i in the 3x3 cell
// neighbourhood plus every big shape j < i.  Count pass, scan in descending i, emit pass that
// also orders each i's partners descending => Aabb.culledKeys order (Aabb.hs:155-183).
// ---------------------------------------------------------------------------------------------

constexpr int EMIT_LOCAL = 24;

// Exclusive prefix sum of one value per thread over a 128-thread block; `total` = the block's sum.
__device__ __forceinline__ unsigned long long block128_exclusive(unsigned long long v, unsigned long long &total)
{
    __shared__ unsigned long long s_warp[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();            // s_warp may still be read by the previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long before = 0;
    total = 0;
    for (int w = 0; w < 4; ++w) { if (w < warp) before += s_warp[w]; total += s_warp[w]; }
    return before + inc - v;
}

enum { SWEEP_COUNT = 0, SWEEP_EMIT = 1, SWEEP_FUSED = 2 };

// Candidates of query i (a small shape in cell (cx, cy), AABB bi, static flag si) in enumeration order: the three
// grid-row runs of its 3x3 neighbourhood, then the big list.  `hit(j)` is called for every partner j < i whose
// AABB passes aabbCheck; returns the hit mask (bit c = candidate c hit; bit 63 = more than 63 candidates, no mask).
template <typename F>
__device__ __forceinline__ unsigned long long sweep_test_all(const Params &P, const FrameState *st, int i, bool si, const Box &bi,
                                                             int cx, int cy, F &&hit)
{
    const int W = st->W, H = st->H;
    unsigned long long mask = 0;
    unsigned cand = 0;
    for (int dy = -1; dy <= 1; ++dy) {
        const int ny = cy + dy;
        if (ny < 0 || ny >= H) continue;
        // the three cells of a grid row are consecutive keys => one contiguous run of candidates
        const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, W - 1);
        const unsigned q_lo = __ldg(&P.cell_begin[(size_t)ny * W + x_lo]);
        const unsigned q_hi = __ldg(&P.cell_begin[(size_t)ny * W + x_hi + 1]);
        for (unsigned q = q_lo; q < q_hi; ++q, ++cand) {
            const uint32_t m = __ldg(&P.smeta[q]);
            const int j = (int)(m & 0x7fffffffu);
            if (j >= i) continue;
            if (si && (m >> 31)) continue; // never pair two static shapes (Aabb.hs:172-176)
            const Box bj = P.sbox[q];
            if (aabb_check(bi, bj)) { hit(j); if (cand < 63u) mask |= 1ull << cand; }
        }
    }
    const unsigned n_big = st->n_big;
    for (unsigned b = 0; b < n_big; ++b, ++cand) {
        const int j = (int)P.big_idx[b];
        if (j >= i) continue;
        if (si && (P.work_mode == 2 ? (P.gkeys[j] & KEY_STATIC_BIT) != 0u : slot_static(P, j))) continue;
        const Box bj = P.work_mode == 2 ? P.box[j] : box_of(P, j);      // rows mode: delivered with j's record (or my own fold)
        if (aabb_check(bi, bj)) { hit(j); if (cand < 63u) mask |= 1ull << cand; }
    }
    return (cand > 63u) ? (1ull << 63) : mask;
}

// The same hits again from the mask sweep_test_all returned (bit 63 clear): no AABB records, no overlap tests.
template <typename F>
__device__ __forceinline__ void sweep_replay(const Params &P, const FrameState *st, int cx, int cy, unsigned long long want, F &&hit)
{
    const int W = st->W, H = st->H;
    unsigned long long m = want;
    unsigned seen = 0;
    for (int dy = -1; dy <= 1 && m; ++dy) {
        const int ny = cy + dy;
        if (ny < 0 || ny >= H) continue;
        const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, W - 1);
        const unsigned q_lo = __ldg(&P.cell_begin[(size_t)ny * W + x_lo]);
        const unsigned q_hi = __ldg(&P.cell_begin[(size_t)ny * W + x_hi + 1]);
        const unsigned len = q_hi - q_lo;
        // bits [seen, seen + len) belong to this run
        unsigned long long run = (len >= 64u - seen) ? (m >> seen) : ((m >> seen) & ((1ull << len) - 1ull));
        while (run) {
            const int b = __ffsll((long long)run) - 1;
            run &= run - 1;
            hit((int)(__ldg(&P.smeta[q_lo + (unsigned)b]) & 0x7fffffffu));
        }
        seen += len;
        if (seen >= 63u) break;
    }
    if (seen < 63u) {
        unsigned long long run = (want & 0x7fffffffffffffffull) >> seen;
        while (run) {
            const int b = __ffsll((long long)run) - 1;
            run &= run - 1;
            hit((int)P.big_idx[b]);
        }
    }
}

// MODE SWEEP_COUNT / SWEEP_EMIT: the two-pass sweep -- counts per query, (scan in descending i), then the pairs
// straight into their place of the reference order; the count pass leaves a hit mask per query so that the emit
// pass revisits only the hits.
// MODE SWEEP_FUSED (sorted mode): ONE pass.  A query counts its partners (mask in a register), the 128 queries of
// a tile reserve a run of the cell-ordered SAT work list with one atomic (tiles are handed out in launch order, so
// the list follows the cell order closely), and each query replays its hits into (i, j, a) entries, a = the rank
// of j among i's partners in descending order.  The pair's index in the reference order is off[r(i)] + a, resolved
// by the SAT stage after the scan; pair_i / pair_j are written there too.
template <int MODE>
__global__ void __launch_bounds__(128) k_sweep(Params P)
{
    constexpr bool EMIT = MODE == SWEEP_EMIT, FUSED = MODE == SWEEP_FUSED;
    const FrameState *st = P.st;
    if ((EMIT && st->error) || (st->error & ERR_REPLAN)) return;
    const unsigned n_sorted = P.cell_begin[st->cell_end]; // shapes in this rank's grid
    const bool rows = P.work_mode == 2;
    __shared__ unsigned long long s_wbase;
    // blocks walk whole 128-position tiles, so that the work-list reservation below is block uniform
    for (unsigned tile = blockIdx.x * 128u; tile < n_sorted; tile += gridDim.x * 128u) {
        const unsigned p = tile + threadIdx.x;
        uint32_t meta = 0;
        int i = -1;
        bool query = false;
        int cx = 0, cy = 0;
        if (p < n_sorted) {
            meta = P.smeta[p];
            i = (int)(meta & 0x7fffffffu);
            if (rows) {   // rows mode: a query belongs to the rank that sweeps its grid row
                const uint32_t key = P.keys_sorted[p];
                cy = (int)(key / (uint32_t)st->W); cx = (int)(key % (uint32_t)st->W);
                query = cy >= st->row_lo && cy < st->row_hi;
            } else query = i >= P.own_lo && i < P.own_hi;
        }
        const int r = P.own_hi - 1 - i;
        const bool si = (meta >> 31) != 0;
        Box bi{ 0.0, 0.0, 0.0, 0.0 };
        if (query) {
            bi = P.sbox[p];
            if (!rows) {
                const uint32_t key = P.keys_sorted[p];
                cy = (int)(key / (uint32_t)st->W); cx = (int)(key % (uint32_t)st->W);
            }
        }
        unsigned long long count = 0, mask = 0;
        if (query && !EMIT) {
            mask = sweep_test_all(P, st, i, si, bi, cx, cy, [&](int) { ++count; });
            if (!rows) P.cnt[r] = count;
            if (!FUSED) P.hitmask[p] = mask;
        }
        if (MODE == SWEEP_COUNT) continue;

        unsigned long long base = 0;      // first output (EMIT) / work-list (FUSED) index of this query
        bool room = true;
        if (FUSED) {
            unsigned long long total;
            const unsigned long long before = block128_exclusive(count, total);
            if (threadIdx.x == 0) s_wbase = total ? atomicAdd(&P.st->work_cursor, total) : 0ull;
            __syncthreads();
            base = s_wbase + before;
            room = s_wbase + total <= (unsigned long long)P.max_pairs;   // else k_finish_pairs raises the capacity error
            if (rows) {
                if (!room && threadIdx.x == 0) atomicOr(&P.st->error, ERR_PAIR_CAP);   // this rank's work list is full
                if (query) {  // the HOME of i learns how many partners i has and who found them
                    P.rw_cq[rw_home(P, i)][i] = ((uint32_t)P.my_rank << 28) | (uint32_t)count;
                    if (count) {   // would i's results stay on this GPU? (under either home layout; sampled: one query in 16)
                        if ((p & 15u) == 0u) {
                            const int G = P.n_peers;
                            const int chunk_blk = P.rw_fold ? 2 * P.rw_blk : P.rw_blk, half_blk = P.rw_fold ? P.rw_blk : (P.rw_blk + 1) / 2;
                            const int bf = i / half_blk;
                            if ((bf < G ? bf : 2 * G - 1 - bf) == P.my_rank) atomicAdd(&P.st->loc_fold, count);
                            if (i / chunk_blk == P.my_rank) atomicAdd(&P.st->loc_contig, count);
                        }
                    }
                }
                // pairs per row bin, for the next frame's cuts: the tile's total goes to the bin of its first row
                // (a tile of 128 consecutive cell-sorted positions spans a row or two)
                if (threadIdx.x == 0 && total) {
                    const unsigned row0 = P.keys_sorted[tile] / (uint32_t)st->W;
                    atomicAdd(&P.roww[(unsigned)(((unsigned long long)row0 * ROW_BINS) / (unsigned)st->H)], (uint32_t)total);
                }
            }
        } else if (query) {
            base = P.off[r];
            mask = P.hitmask[p];
            count = 0;
        }
        if (!query || !room) continue;
        if (FUSED && count == 0) continue;

        int32_t *const out_j = FUSED ? reinterpret_cast<int32_t *>(P.w_j) : P.pair_j;
        int local[EMIT_LOCAL];
        unsigned long long n_hit = 0;
        auto collect = [&](int j) {
            if (n_hit < EMIT_LOCAL) local[n_hit] = j;
            else out_j[base + n_hit] = j;
            ++n_hit;
        };
        if (!(mask >> 63)) sweep_replay(P, st, cx, cy, mask, collect);
        else sweep_test_all(P, st, i, si, bi, cx, cy, collect);
        if (n_hit == 0) continue;
        // each i's partners in descending order => Aabb.culledKeys order (Aabb.hs:155-183)
        if (n_hit <= EMIT_LOCAL) {
            const int n = (int)n_hit;
            for (int a = 1; a < n; ++a) { // insertion sort, descending
                const int v = local[a];
                int b2 = a - 1;
                while (b2 >= 0 && local[b2] < v) { local[b2 + 1] = local[b2]; --b2; }
                local[b2 + 1] = v;
            }
            for (int a = 0; a < n; ++a) out_j[base + a] = local[a];
        } else {
            for (int a = 0; a < EMIT_LOCAL; ++a) out_j[base + a] = local[a];
            int32_t *seg = out_j + base;
            for (unsigned long long a = 1; a < n_hit; ++a) {
                const int v = seg[a];
                long long b2 = (long long)a - 1;
                while (b2 >= 0 && seg[b2] < v) { seg[b2 + 1] = seg[b2]; --b2; }
                seg[b2 + 1] = v;
            }
        }
        if (FUSED) for (unsigned long long a = 0; a < n_hit; ++a) { P.w_i[base + a] = (uint32_t)i; P.w_a[base + a] = (uint32_t)a; }
        else for (unsigned long long a = 0; a < n_hit; ++a) P.pair_i[base + a] = i;
    }
}

// Big shapes as queries: one block per big shape, all slots j < i in descending j.
template <bool EMIT>
__global__ void __launch_bounds__(256) k_big(Params P)
{
    const FrameState *st = P.st;
    if ((EMIT && st->error) || (st->error & ERR_REPLAN)) return;
    __shared__ unsigned s_warp[8];
    __shared__ unsigned long long s_run, s_wbase;
    const unsigned n_big = st->n_big;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned b = blockIdx.x; b < n_big; b += gridDim.x) {
        const int i = (int)P.big_idx[b];
        const bool rows = P.work_mode == 2;       // rows mode: big queries stay with their home
        if (rows ? !rw_mine(P, i) : (i < P.own_lo || i >= P.own_hi)) continue;
        const bool listed = P.sorted_mode || rows; // results go to the SAT work list
        const Box bi = P.box[i];
        const bool si = rows ? (P.gkeys[i] & KEY_STATIC_BIT) != 0u : slot_static(P, i);
        const int r = P.own_hi - 1 - i;
        const unsigned long long base = (EMIT && !rows) ? P.off[r] : 0ull;
        if (threadIdx.x == 0) {
            s_run = 0;
            // this query's run of the SAT work list (see k_sweep)
            if (EMIT && listed) {
                const unsigned long long n = rows ? (unsigned long long)(P.rw_cq[P.my_rank][i] & 0x0fffffffu) : P.cnt[r];
                s_wbase = n ? atomicAdd(&P.st->work_cursor, n) : 0ull;
                if (rows && s_wbase + n > (unsigned long long)P.max_pairs) atomicOr(&P.st->error, ERR_PAIR_CAP);
            }
        }
        __syncthreads();
        const unsigned long long wbase = (EMIT && listed) ? s_wbase : 0ull;
        const unsigned long long n_mine = !listed ? 0ull : rows ? (unsigned long long)(P.rw_cq[P.my_rank][i] & 0x0fffffffu) : P.cnt[r];
        const bool room = wbase + n_mine <= (unsigned long long)P.max_pairs;
        for (int top = i - 1; top >= 0; top -= (int)blockDim.x) {
            const int j = top - (int)threadIdx.x;
            bool pred = false;
            if (j >= 0 && P.alive[j] && !(si && (rows ? (P.peer_keys[rw_home(P, j)][j] & KEY_STATIC_BIT) != 0u : slot_static(P, j))))
                pred = aabb_check(bi, box_of(P, j));
            const unsigned bal = __ballot_sync(0xffffffffu, pred);
            if (lane == 0) s_warp[warp] = __popc(bal);
            __syncthreads();
            unsigned before = 0, total = 0;
            for (int w = 0; w < 8; ++w) { if (w < warp) before += s_warp[w]; total += s_warp[w]; }
            const unsigned long long run = s_run;
            if (EMIT && pred) {
                const unsigned long long pos = base + run + before + __popc(bal & ((1u << lane) - 1u));
                if (listed) {   // work-list entry; single rank: the SAT stage writes pair_i / pair_j at off[r] + a
                    const unsigned long long a = pos - base, w = wbase + a;
                    if (room) { P.w_i[w] = (uint32_t)i; P.w_j[w] = (uint32_t)j; P.w_a[w] = (uint32_t)a; }
                } else {
                    P.pair_i[pos] = i;
                    P.pair_j[pos] = j;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_run = run + total;
            __syncthreads();
        }
        if (!EMIT && threadIdx.x == 0) { if (rows) P.rw_cq[P.my_rank][i] = ((uint32_t)P.my_rank << 28) | (uint32_t)s_run; else P.cnt[r] = s_run; }
        __syncthreads();
    }
}

__global__ void k_finish_pairs(Params P, int n_query)
{
    long long total = 0;
    if (n_query > 0) total = (long long)(P.off[n_query - 1] + P.cnt[n_query - 1]);
    P.st->n_pairs = total;
    if (total > P.max_pairs) P.st->error |= ERR_PAIR_CAP;
}

// ---------------------------------------------------------------------------------------------
// K3: SAT + clipping + constraint generators, one thread per pair, ordered compaction
// ---------------------------------------------------------------------------------------------

constexpr int CT_THREADS = 128;
#ifndef CT_MIN_BLOCKS_OVERRIDE
#define CT_MIN_BLOCKS_OVERRIDE 6
#endif
constexpr int CT_MIN_BLOCKS = CT_MIN_BLOCKS_OVERRIDE;

struct HullAcc {
    int slot, off, n;
    int which;                 // 0 = a, 1 = b (selects the shared-memory plane)
    bool owned;                // world vertices / normals of this slot were materialised by K0 on this rank
    unsigned long long ext;    // packed extents (n <= 8)
    V2 n0;                     // unit normal of edge 0, fetched while staging
};

struct SatRes { bool sep; int edge; double depth; int pen; };

template <int MAXV>
struct ContactKernel {
    const Params &P;
    double2 (*sv)[MAXV][32]; // [2][MAXV][lane] world vertices staged by (and private to) each lane
    int tid;                 // lane

    __device__ __forceinline__ V2 slow_vertex(const HullAcc &h, int k) const
    {
        if (h.owned) { const double2 v = P.wv[h.off + k]; return V2{ v.x, v.y }; }
        const Xf x = slot_xf(P, h.slot);
        const Aff m = to_transform(x.px, x.py, x.c, x.s);
        const double2 l = __ldg(&P.local[h.off + k]);
        return afmul(m, V2{ l.x, l.y });
    }
    __device__ __forceinline__ V2 vtx(const HullAcc &h, int k) const
    {
        if (h.n <= MAXV) { const double2 v = sv[h.which][k][tid]; return V2{ v.x, v.y }; }
        return slow_vertex(h, k);
    }
    __device__ __forceinline__ void ext(const HullAcc &h, int e, int &imin, int &imax) const
    {
        if (h.n <= MAX_STAGED_VERTS) {
            const unsigned bits = (unsigned)(h.ext >> (6 * e));
            imin = bits & 7; imax = (bits >> 3) & 7;
        } else { imin = P.ext_min[h.off + e]; imax = P.ext_max[h.off + e]; }
    }
    __device__ __forceinline__ void stage(HullAcc &h) const
    {
        // rows mode: the hulls this rank keeps were materialised by k_rw_hulls (a big query's partners may not be)
        h.owned = P.work_mode == 2 ? P.mat_stamp[h.slot] == (uint32_t)P.st->frame_no : (h.slot >= P.own_lo && h.slot < P.own_hi);
        if (h.n <= MAXV) {
            if (h.owned) {
                if (MAXV > 4) {
                    // general worlds: partners are far apart in memory, so every vertex load is
                    // issued before the first is consumed (one DRAM latency instead of n)
                    double2 tmp[MAXV];
#pragma unroll
                    for (int k = 0; k < MAXV; ++k) if (k < h.n) tmp[k] = P.wv[h.off + k];
#pragma unroll
                    for (int k = 0; k < MAXV; ++k) if (k < h.n) sv[h.which][k][tid] = tmp[k];
                } else {
                    for (int k = 0; k < h.n; ++k) sv[h.which][k][tid] = P.wv[h.off + k];
                }
            } else {
                const Xf x = slot_xf(P, h.slot);
                const Aff m = to_transform(x.px, x.py, x.c, x.s);
                for (int k = 0; k < h.n; ++k) {
                    const double2 l = __ldg(&P.local[h.off + k]);
                    const V2 w = afmul(m, V2{ l.x, l.y });
                    sv[h.which][k][tid] = make_double2(w.x, w.y);
                }
            }
        }
        h.n0 = normal(h, 0);
    }
    // unit normal of edge e: K0's value for owned slots, else recomputed from the (identical) vertices
    __device__ __forceinline__ V2 normal(const HullAcc &h, int e) const
    {
        if (h.owned) { const double2 v = P.wn[h.off + e]; return V2{ v.x, v.y }; }
        const int e1 = (e < h.n - 1) ? e + 1 : 0; // nextIndex (ConvexHull.hs:228-230)
        return unit_edge_normal(vtx(h, e), vtx(h, e1));
    }

    // minOverlap' sEdge sPen (SAT.hs:121-143): the first separating edge wins (later edges cannot
    // change the fold), else strictly smaller depth.  NE / NP: compile-time vertex counts (0 = read
    // them from the hull).  Two schedules for the edge normals:
    //  * boxes-only kernel (MAXV <= 4): partners are neighbours in memory, normals come from L1/L2
    //    one edge ahead of their use (keeps registers and shared memory small => more L1);
    //  * general kernel: partners are scattered, so all normals of E are requested as one batch.
    //    Loops have compile-time bounds with `break` guards so lanes with different vertex counts
    //    stay converged (runtime trip counts under `#pragma unroll` made ptxas emit remainder
    //    jumps that serialised the lanes: 4.7 active threads per instruction).
    template <int NE, int NP>
    __device__ __forceinline__ SatRes min_overlap(const HullAcc &E, const HullAcc &Pn) const
    {
        const int ne = NE ? NE : E.n, np = NP ? NP : Pn.n;
        SatRes best{ false, 0, 0.0, 0 };
        if (MAXV <= 4 || NE) {
            V2 dir_next = E.n0;
#pragma unroll
            for (int e = 0; e < (NE ? NE : MAXV); ++e) {
                if (e >= ne) break;
                const V2 dir = dir_next;
                if (e + 1 < ne) dir_next = normal(E, e + 1); // issued one edge ahead of its use
                if (edge_test<NP>(E, Pn, e, dir, np, best)) return best;
            }
        } else if (ne <= MAXV) {
            V2 nrm[MAXV];
            nrm[0] = E.n0;
#pragma unroll
            for (int e = 1; e < MAXV; ++e) if (e < ne) nrm[e] = normal(E, e);
#pragma unroll
            for (int e = 0; e < MAXV; ++e) {
                if (e >= ne) break;
                if (edge_test<NP>(E, Pn, e, nrm[e], np, best)) return best;
            }
        } else {
#pragma unroll 1
            for (int e = 0; e < ne; ++e)
                if (edge_test<NP>(E, Pn, e, normal(E, e), np, best)) return best;
        }
        return best;
    }

    // overlap sEdge edge sPen (SAT.hs:103-117) folded into minOverlap's accumulator; true = separated
    template <int NP>
    __device__ __forceinline__ bool edge_test(const HullAcc &E, const HullAcc &Pn, int e, V2 dir, int np, SatRes &best) const
    {
        int imin, imax;
        ext(E, e, imin, imax);
        // extentAlongSelf (ConvexHull.hs:111-118): the cached extreme vertices only
        const double s_min = dot2(vtx(E, imin), dir);
        const double s_max = dot2(vtx(E, imax), dir);
        // extentAlong (ConvexHull.hs:81-100): first minimum / first maximum win
        double p_min = dot2(vtx(Pn, 0), dir), p_max = p_min;
        int p_idx = 0;
        if (NP) {
#pragma unroll
            for (int k = 1; k < NP; ++k) {
                const double d = dot2(vtx(Pn, k), dir);
                if (d < p_min) { p_min = d; p_idx = k; }
                if (d > p_max) p_max = d;
            }
        } else {
#pragma unroll 1
            for (int k = 1; k < np; ++k) {
                const double d = dot2(vtx(Pn, k), dir);
                if (d < p_min) { p_min = d; p_idx = k; }
                if (d > p_max) p_max = d;
            }
        }
        if ((p_min > s_max) || (p_max < s_min)) { best.sep = true; best.edge = e; return true; } // overlapTest (SAT.hs:74-83)
        const double depth = fsub(s_max, p_min); // overlapAmount (SAT.hs:86-96)
        if (e == 0 || depth < best.depth) { best.edge = e; best.depth = depth; best.pen = p_idx; }
        return false;
    }

    // ---- circles: Circle.contact, GJK closestSimplex, CircleVsHull ----

    // world centre of a circle slot: K0's value for owned slots, else setCircleTransform here
    __device__ __forceinline__ V2 circle_center(int slot) const
    {
        if (P.work_mode == 2 ? P.mat_stamp[slot] == (uint32_t)P.st->frame_no : (slot >= P.own_lo && slot < P.own_hi)) {
            const double2 c = P.circ[slot]; return V2{ c.x, c.y };
        }
        const Xf x = slot_xf(P, slot);
        return afmul(to_transform(x.px, x.py, x.c, x.s), V2{ 0.0, 0.0 });
    }

    // support (ConvexHull.hs:124-128): first maximum of dir . v
    __device__ __forceinline__ int support(const HullAcc &h, V2 dir) const
    {
        int best = 0;
        double bd = dot2(vtx(h, 0), dir);
#pragma unroll 1
        for (int k = 1; k < h.n; ++k) {
            const double d = dot2(vtx(h, k), dir);
            if (d > bd) { bd = d; best = k; }
        }
        return best;
    }

    // closestSimplex hull origin (GJK.hs:52-69): size of the final simplex (1, 2; 3 = encloses the
    // target; 0 = iteration cap) and its vertex indices, most recently added first.
    __device__ int closest_simplex(const HullAcc &h, V2 origin, int &i0, int &i1) const
    {
        int n = 1, i2 = 0;
        i0 = 0; i1 = 0;
        V2 d = sub2(origin, vtx(h, 0));
#pragma unroll 1
        for (int it = 0; it < 64; ++it) {
            const int aa = support(h, d);
            // extendSimplex (GJK.hs:71-90): a repeated vertex ends the search
            if (n == 1) { if (i0 == aa) return 1; i1 = i0; i0 = aa; n = 2; }
            else { if (i0 == aa || i1 == aa) return 2; i2 = i1; i1 = i0; i0 = aa; n = 3; }
            const V2 a = vtx(h, i0), b = vtx(h, i1);
            const V2 ab = sub2(b, a), ao = sub2(origin, a);
            if (n == 2) { // shiftSimplex2 (GJK.hs:98-112)
                if (dot2(ab, ao) > 0.0) d = cross_v2v2(ab, ao, ab);
                else { n = 1; d = ao; }
            } else {      // shiftSimplex3 (GJK.hs:114-136)
                const V2 c = vtx(h, i2);
                const V2 ac = sub2(c, a);
                const double abc = cross2(ab, ac);
                const V2 abcac{ -fmul(abc, ac.y), fmul(abc, ac.x) };  // abc `zcrossV2` ac (Linear.hs:126-129)
                const V2 ababc{ fmul(ab.y, abc), -fmul(ab.x, abc) };  // ab `crosszV2` abc (Linear.hs:121-124)
                bool star = false;
                if (dot2(abcac, ao) > 0.0) {
                    if (dot2(ac, ao) > 0.0) { i1 = i2; n = 2; d = cross_v2v2(ac, ao, ac); }
                    else star = true;
                } else if (dot2(ababc, ao) > 0.0) star = true;
                else return 3; // the simplex encloses the target
                if (star) {
                    if (dot2(ab, ao) > 0.0) { n = 2; d = cross_v2v2(ab, ao, ab); }
                    else { n = 1; d = ao; }
                }
            }
        }
        return 0;
    }
    // crossV2V2 (Linear.hs:135-139)
    __device__ __forceinline__ static V2 cross_v2v2(V2 a, V2 b, V2 c)
    {
        const double abz = fsub(fmul(a.x, b.y), fmul(a.y, b.x));
        return V2{ -fmul(abz, c.y), fmul(abz, c.x) };
    }

    // Circle.contact circleA circleB (Circle.hs:30-53): A is the penetratee, normal out of A.
    __device__ bool circle_circle(V2 a, double ra, V2 b, double rb, V2 &normal, V2 &center, double &depth) const
    {
        const V2 ab = sub2(b, a);
        const double rab = fadd(ra, rb);
        const double ab_sq = fadd(fmul(ab.x, ab.x), fmul(ab.y, ab.y));
        if (!(fmul(rab, rab) >= ab_sq)) return false;
        const double ab_len = __dsqrt_rn(ab_sq);
        normal = V2{ fdiv(ab.x, ab_len), fdiv(ab.y, ab_len) };
        const V2 a1{ fadd(fmul(normal.x, ra), a.x), fadd(fmul(normal.y, ra), a.y) };
        const double nrb = -rb;
        const V2 b1{ fadd(fmul(normal.x, nrb), b.x), fadd(fmul(normal.y, nrb), b.y) };
        center = V2{ fdiv(fadd(a1.x, b1.x), 2.0), fdiv(fadd(a1.y, b1.y), 2.0) }; // midpointP2
        depth = fsub(fadd(ra, rb), ab_len);
        return true;
    }

    // CircleVsHull.generateContacts (CircleVsHull.hs:18-69): the circle is always the penetrator
    __device__ bool circle_hull(V2 ctr, double r, const HullAcc &h, int &feature, V2 &normal, V2 &point, double &depth) const
    {
        int i0, i1;
        const int n = closest_simplex(h, ctr, i0, i1);
        if (n != 1 && n != 2) return false; // Simplex3' (deep overlap) => Nothing (CircleVsHull.hs:29)
        V2 a = vtx(h, i0);
        if (n == 2) { // closestAlong (CircleVsHull.hs:60-69)
            const V2 b = vtx(h, i1);
            const V2 ao = sub2(ctr, a), ab = sub2(b, a);
            const double len = __dsqrt_rn(fadd(fmul(ab.x, ab.x), fmul(ab.y, ab.y)));
            const V2 abn{ fdiv(ab.x, len), fdiv(ab.y, len) };      // normalizeV2
            const double sc = dot2(ao, abn);
            a = V2{ fadd(fmul(abn.x, sc), a.x), fadd(fmul(abn.y, sc), a.y) };
        }
        // processSimplex_ (CircleVsHull.hs:42-58)
        const V2 ab = sub2(ctr, a);
        const double ab_sq = fadd(fmul(ab.x, ab.x), fmul(ab.y, ab.y));
        if (fmul(r, r) < ab_sq) return false;
        const double ab_len = __dsqrt_rn(ab_sq);
        normal = V2{ -fdiv(ab.x, ab_len), -fdiv(ab.y, ab_len) };   // negateV2 normal
        point = a;
        depth = fsub(r, ab_len);
        feature = i0;
        return true;
    }
};

enum { CLIP_LEFT = 0, CLIP_RIGHT = 1, CLIP_BOTH = 2, CLIP_NONE = 3 };

// clipSegment (Linear.hs:327-343) with intersect2 (Linear.hs:244-251) and invM2x2 (Linear.hs:194-199).
// The incident line (ip, in) and its offset ib = ip . in are the same for all three clips.
__device__ __forceinline__ int clip_segment(V2 bp, V2 bn, V2 in, double ib, V2 a, V2 b, V2 &c)
{
    const double b0 = dot2(bp, bn);
    const double det = fsub(fmul(bn.x, in.y), fmul(bn.y, in.x));
    const double inv = fdiv(1.0, det);
    const double m00 = fmul(in.y, inv), m01 = fmul(-bn.y, inv);
    const double m10 = fmul(-in.x, inv), m11 = fmul(bn.x, inv);
    c.x = fadd(fmul(m00, b0), fmul(m01, ib));
    c.y = fadd(fmul(m10, b0), fmul(m11, ib));
    const double a1 = dot2(a, bn), b1 = dot2(b, bn), c1 = dot2(c, bn);
    if (a1 < c1) return (b1 < c1) ? CLIP_BOTH : CLIP_LEFT;
    if (b1 < c1) return CLIP_RIGHT;
    return CLIP_NONE;
}

// The clipped manifold of an overlapping hull pair (both SAT directions found no separating axis):
// penetratedEdge / penetratingEdge (SAT.hs:152-171), clipEdge (SAT.hs:190-218), flattenContactPoints
// (SAT.hs:181-187).  E = penetrated hull, Pn = penetrating hull; `acc` supplies their world vertices and
// the unit normal of E's edge.  Writes the ManRec when the pair has contacts; returns their number.
template <typename Acc>
__device__ __forceinline__ unsigned emit_manifold(ManRec *out_rec, const Acc &acc, int e_n, int pn_n,
                                                  int edge, int pen, bool same)
{
    unsigned cnt = 0;
    const V2 n = acc.normal_e(edge); // overlapNormal (SAT.hs:98-100)
    // penetratedEdge (SAT.hs:169-171)
    const int e1 = (edge < e_n - 1) ? edge + 1 : 0;
    const V2 ra = acc.ve(edge), rb = acc.ve(e1);
    // penetratingEdge (SAT.hs:152-166)
    const int ib = pen;
    const int ic = (ib < pn_n - 1) ? ib + 1 : 0;
    const int ia = (ib > 0) ? ib - 1 : pn_n - 1;
    const V2 va = acc.vp(ia), vb = acc.vp(ib), vc = acc.vp(ic);
    const double abn = fabs(dot2(sub2(vb, va), n));
    const double bcn = fabs(dot2(sub2(vc, vb), n));
    V2 q0, q1;
    int i0, i1;
    if (bcn < abn) { q0 = vb; i0 = ib; q1 = vc; i1 = ic; }
    else { q0 = va; i0 = ia; q1 = vb; i1 = ib; }
    // clipEdge (SAT.hs:190-218)
    const V2 inc_n = clockwise2(sub2(q1, q0)); // toLine2 c d, unclipped endpoints
    const double inc_b = dot2(q0, inc_n);
    V2 x;
    bool alive = true;
    int r = clip_segment(ra, sub2(rb, ra), inc_n, inc_b, q0, q1, x); // perpLine2 a b
    if (r == CLIP_BOTH) alive = false;
    else if (r == CLIP_LEFT) q0 = x;
    else if (r == CLIP_RIGHT) q1 = x;
    if (alive) {
        r = clip_segment(rb, sub2(ra, rb), inc_n, inc_b, q0, q1, x); // perpLine2 b a
        if (r == CLIP_BOTH) alive = false;
        else if (r == CLIP_LEFT) q0 = x;
        else if (r == CLIP_RIGHT) q1 = x;
    }
    if (alive) {
        r = clip_segment(ra, neg2(n), inc_n, inc_b, q0, q1, x); // Line2 a (negateV2 n)
        // applyClip'' (Linear.hs:285-292) removes the clipped endpoint;
        // flattenContactPoints (SAT.hs:181-187): descending feature index
        V2 c0 = q0, c1 = q1;
        int p0 = i0, p1 = i1;
        if (r == CLIP_LEFT) { cnt = 1; c0 = q1; p0 = i1; }
        else if (r == CLIP_RIGHT) { cnt = 1; }
        else if (r == CLIP_NONE) {
            cnt = 2;
            if (!(i0 > i1)) { c0 = q1; p0 = i1; c1 = q0; p1 = i0; }
        }
        if (cnt) {
            ManRec rec;
            rec.nx = n.x; rec.ny = n.y;
            rec.ref_d = dot2(ra, n); // contactDepth_ (HullVsHull.hs:30-37): f v, f = afdot' n
            rec.c0x = c0.x; rec.c0y = c0.y; rec.c1x = c1.x; rec.c1y = c1.y;
            rec.bits = (unsigned long long)(unsigned)edge | ((unsigned long long)(unsigned)p0 << 20) |
                       ((unsigned long long)(unsigned)p1 << 40) | ((unsigned long long)(same ? 0 : 1) << 60);
            if (out_rec) *out_rec = rec;      // null: the home's arrays are full (it raises the capacity error itself)
        }
    }
    return cnt;
}

// Accessor over the per-thread kernel's staged hulls
template <int MAXV>
struct StagedAcc {
    const ContactKernel<MAXV> &K;
    const HullAcc &E, &Pn;
    __device__ __forceinline__ V2 normal_e(int e) const { return K.normal(E, e); }
    __device__ __forceinline__ V2 ve(int k) const { return K.vtx(E, k); }
    __device__ __forceinline__ V2 vp(int k) const { return K.vtx(Pn, k); }
};

// One hull pair on ONE thread: stage both hulls, SAT both ways, clip.  Returns the contact count.
template <int MAXV>
__device__ __forceinline__ unsigned hull_pair_manifold(const ContactKernel<MAXV> &K, ManRec *out_rec, HullAcc &A, HullAcc &B)
{
    K.stage(A);
    K.stage(B);
    // contactDebug (SAT.hs:238-248): eitherBranchBoth (Utils.hs:230-235) -- a separating
    // axis on either side means no contact; else depth_ab < depth_ba ? Same : Flip.
    const bool boxes = (A.n == 4) & (B.n == 4);
    const SatRes ab = boxes ? K.template min_overlap<4, 4>(A, B) : K.template min_overlap<0, 0>(A, B);
    if (ab.sep) return 0;
    const SatRes ba = boxes ? K.template min_overlap<4, 4>(B, A) : K.template min_overlap<0, 0>(B, A);
    if (ba.sep) return 0;
    const bool same = ab.depth < ba.depth;
    const HullAcc &E = same ? A : B;
    const HullAcc &Pn = same ? B : A;
    const SatRes ov = same ? ab : ba;
    return emit_manifold(out_rec, StagedAcc<MAXV>{ K, E, Pn }, E.n, Pn.n, ov.edge, ov.pen, same);
}

// Where a pair's results go.  work_mode 0: SAT walks the pairs in reference order, results at the pair's own index.
// 1: cell-ordered work list, results at off[r(i)] + a of this rank's arrays, pair_i / pair_j written too.
// 2 (rows mode): results at q_off[i] + a of the arrays of i's HOME (stores over NVLink), plus the partner's position
// and inverse masses for k_rows there; `ok` false = beyond the home's capacity (it raises the error itself).
struct PairOut {
    long long idx;
    uint32_t *ccnt; ManRec *man; int32_t *pair_i, *pair_j;
    PairRec *rec;      // rows mode: the pair's body record at its home (null: the home's arrays are full)
    PairHdr *hdr;      //            and its header
    bool ok;
};
__device__ __forceinline__ PairOut pair_out(const Params &P, long long w, int i)
{
    PairOut o{ w, P.ccnt, P.man, nullptr, nullptr, nullptr, nullptr, true };
    if (P.work_mode == 1) { o.idx = (long long)(P.off[P.own_hi - 1 - i] + P.w_a[w]); o.pair_i = P.pair_i; o.pair_j = P.pair_j; }
    else if (P.work_mode == 2) {
        const int h = P.dbg_local_stores ? P.my_rank : rw_home(P, i);
        o.idx = (long long)P.q_off[i] + (long long)P.w_a[w];
        o.ok = o.idx < P.max_pairs;
        o.rec = o.ok ? &P.rw_prec[h][o.idx] : nullptr;
        o.hdr = o.ok ? &P.rw_phdr[h][o.idx] : nullptr;
        o.ccnt = nullptr; o.man = nullptr;
    }
    return o;
}
// where the pair's manifold record goes (null: nowhere)
__device__ __forceinline__ ManRec *pair_out_man(const Params &P, const PairOut &o)
{
    if (!o.ok) return nullptr;
    return P.work_mode == 2 ? &o.rec->man : &o.man[o.idx];
}
// rows mode: (pos_j, inverse masses of j) for the home's k_rows.  j's records were pushed here if this rank keeps j;
// a big query's far partner is read from its home
__device__ __forceinline__ double4 partner_record(const Params &P, int j)
{
    if (P.mat_stamp[j] == (uint32_t)P.st->frame_no) {      // kept here: its record is in my inbox (keys[] = rec_of[])
        const HomeRec &r = P.inbox[P.keys[j]];
        return make_double4(r.xf.px, r.xf.py, r.mass.x, r.mass.y);
    }
    const int hj = rw_home(P, j);
    const Xf x = P.rw_xf[hj][j];
    const double2 m = P.rw_mass[hj][j];
    return make_double4(x.px, x.py, m.x, m.y);
}
// the pair's index entries and, in rows mode when it has contacts, the partner's body record
__device__ __forceinline__ void pair_out_finish(const Params &P, const PairOut &o, int i, int j, unsigned cnt)
{
    if (!o.ok) return;
    if (P.work_mode == 2) {
        *reinterpret_cast<int4 *>(o.hdr) = make_int4(i, j, (int)cnt, 0);
        if (cnt != 0u && cnt != 0xffffffffu) o.rec->pj = partner_record(P, j);
        return;
    }
    o.ccnt[o.idx] = cnt;
    if (o.pair_i) { o.pair_i[o.idx] = i; o.pair_j[o.idx] = j; }
}

// K3a: one thread per pair.  SAT both ways + incident-edge clipping; writes the pair's contact
// count and, when it has contacts, its manifold record.  No ordering between pairs.
#ifndef CT_MIN_BLOCKS_GENERAL
#define CT_MIN_BLOCKS_GENERAL 4
#endif
constexpr unsigned CCNT_FALLBACK = 0xffffffffu;   // k_manifolds_coop: "left to the per-thread kernel"

template <int MAXV, bool CIRCLES, bool FLAGGED_ONLY = false>
__global__ void __launch_bounds__(CT_THREADS, MAXV <= 4 ? CT_MIN_BLOCKS : CT_MIN_BLOCKS_GENERAL) k_manifolds(Params P)
{
    __shared__ double2 s_verts[CT_THREADS / 32][2][MAXV][32];

    const FrameState *st = P.st;
    if (st->error) return;
    const bool rows = P.work_mode == 2;   // rows mode: p walks this rank's work list, results stay in work order
    const long long n_pairs = rows ? (long long)st->work_cursor : st->n_pairs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ContactKernel<MAXV> K{ P, s_verts[warp], lane };

    for (long long p = (long long)blockIdx.x * CT_THREADS + threadIdx.x; p < n_pairs;
         p += (long long)gridDim.x * CT_THREADS) {
        if (FLAGGED_ONLY && P.sat_ccnt[p] != CCNT_FALLBACK) continue;   // second pass after k_manifolds_coop
        const int i = rows ? (int)P.w_i[p] : P.pair_i[p], j = rows ? (int)P.w_j[p] : P.pair_j[p];
        // (work_mode 1 reaches this kernel only as the flagged pass, which walks the reference order: like mode 0)
        PairOut o{ p, P.ccnt, P.man, nullptr, nullptr, nullptr, nullptr, true };
        if (rows) o = pair_out(P, p, i);
        ManRec *const out_rec = pair_out_man(P, o);
        HullAcc A, B; // A = shape with the larger key (Aabb.hs:174-179, Solvers/Contact.hs:48-51)
        A.slot = i; A.off = P.vert_offset[i]; A.n = P.vert_offset[i + 1] - A.off; A.which = 0;
        B.slot = j; B.off = P.vert_offset[j]; B.n = P.vert_offset[j + 1] - B.off; B.which = 1;
        A.ext = P.ext_packed[i];
        B.ext = P.ext_packed[j];
        if (CIRCLES) {
            // generateContacts dispatch (shapes/src/Physics/Contact.hs:22-40)
            const double ra = P.radius[i], rb = P.radius[j];
            if (ra >= 0.0 || rb >= 0.0) {
                V2 normal{ 0.0, 0.0 }, center{ 0.0, 0.0 };
                double depth = 0.0;
                int feature = 0, flip = 0;
                bool hit;
                if (ra >= 0.0 && rb >= 0.0) {            // ((0, 0), Same contact)
                    hit = K.circle_circle(K.circle_center(i), ra, K.circle_center(j), rb, normal, center, depth);
                } else if (ra >= 0.0) {                   // ((0, hullFeature), Same contact)
                    K.stage(B);
                    hit = K.circle_hull(K.circle_center(i), ra, B, feature, normal, center, depth);
                } else {                                  // ((hullFeature, 0), Flip contact)
                    K.stage(A);
                    hit = K.circle_hull(K.circle_center(j), rb, A, feature, normal, center, depth);
                    flip = 1;
                }
                if (hit) {
                    ManRec rec;
                    rec.nx = normal.x; rec.ny = normal.y;
                    rec.ref_d = depth;                    // explicit depth (bit 61): not derived from a reference edge
                    rec.c0x = center.x; rec.c0y = center.y; rec.c1x = 0.0; rec.c1y = 0.0;
                    rec.bits = ((unsigned long long)(unsigned)feature << 20) | ((unsigned long long)flip << 60) | (1ull << 61);
                    if (out_rec) *out_rec = rec;
                }
                pair_out_finish(P, o, i, j, hit ? 1u : 0u);
                if (rows) P.sat_ccnt[p] = hit ? 1u : 0u;
                continue;
            }
        }
        const unsigned cnt = hull_pair_manifold<MAXV>(K, out_rec, A, B);
        pair_out_finish(P, o, i, j, cnt);
        if (rows) P.sat_ccnt[p] = cnt;
    }
}


// K3a for general convex polygons (hulls of 3..8 vertices): 16 lanes per pair.
// In `k_manifolds` a thread walks (edges of A) x (vertices of B) + (edges of B) x (vertices of A) on its
// own; with 3..8 vertices per hull the trip counts differ from lane to lane (14.9 of 32 threads active
// per instruction in ncu) and every lane stages two whole hulls.  Here lane (dir, e) of a half warp owns
// ONE candidate axis: the unit normal of edge e of the penetrated hull of direction dir (0: A <- B,
// 1: B <- A), against which it projects the other hull's <= 8 vertices.  A warp does its 32 pairs two at a
// time (phase 1, 16 steps), then every lane clips one pair (phase 2, `emit_manifold`).
//  * The 64 vertex / normal records of a step (2 pairs x 2 hulls x 16) are copied global -> shared with one
//    16 B cp.async per lane and hull, one step AHEAD of their use (double buffer): the SAT arithmetic never waits
//    for a global load, and its 11 operand loads per lane are LDS with immediate offsets.
//  * The minOverlap fold (SAT.hs:121-143) over the eight per-edge results is a shuffle minimum of the depths
//    followed by a ballot that picks the FIRST edge holding that minimum -- the sequential fold's strict `<`,
//    with its NaN behaviour kept (see fold_min_overlap).
//  * Each direction's winner goes to shared memory; the Same / Flip decision moves to phase 2.
// Pairs with a hull of more than 8 vertices, or a hull whose world vertices were not materialised on this
// rank, are flagged and finished by a second, per-thread pass.
// History (1M random polygons, SAT stage): one thread per pair 0.381 ms; r1 16 lanes per pair with global operand
// loads, a lexicographic shuffle fold and index shuffles 0.224 ms (285 warp instructions per step, 62 % issue
// utilisation); 8 lanes per pair / both directions per lane 0.225 ms (fewer instructions, but two dependent load
// batches per step and spills at 80 registers) -- profiles/r2_summary.md.
constexpr int CO_WARPS = 4;

struct GlobalAcc {      // emit_manifold over materialised world vertices / normals in global memory
    const double2 *wv, *wn;
    int e_off, pn_off;
    __device__ __forceinline__ V2 normal_e(int e) const { const double2 v = wn[e_off + e]; return V2{ v.x, v.y }; }
    __device__ __forceinline__ V2 ve(int k) const { const double2 v = wv[e_off + k]; return V2{ v.x, v.y }; }
    __device__ __forceinline__ V2 vp(int k) const { const double2 v = wv[pn_off + k]; return V2{ v.x, v.y }; }
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// minOverlap' (SAT.hs:121-143) over the edges held by the 8 lanes of my group: edge 0 first, a later edge replaces
// the best one only with a strictly smaller depth.  That fold equals "the first edge holding the minimum depth, a
// NaN depth counting as +inf" -- unless edge 0's depth is NaN, in which case nothing ever replaces edge 0.
// (A NaN lane can only be the first holder of the minimum when that minimum is +inf and lane 0 is not itself
// +inf / NaN -- impossible, lane 0 would hold it first.)  Returns the winning edge; depth / pen come from its lane.
__device__ __forceinline__ int fold_min_overlap(bool active, double depth, int group_base)
{
#ifdef COOP_REDUX_FOLD   // measured slower (sub-warp REDUX masks serialise): 1M polygons 0.190 vs 0.161 ms in the SAT stage
    // Depths that matter are >= 0 (no axis separates: pMin <= sMax) or NaN, so after folding -0 into +0 the bit
    // pattern orders like the value and the minimum is two 32-bit warp reductions (REDUX) over the group's 8 lanes.
    // (When some axis separates the depths may be negative and the winner garbage: it is never used.)
    const unsigned gmask = 0xffu << group_base;
    const bool is_nan = depth != depth;
    const double key = (active && !is_nan) ? fadd(depth, 0.0) : __longlong_as_double(0x7ff0000000000000ll);
    const unsigned hi = (unsigned)__double2hiint(key), lo = (unsigned)__double2loint(key);
    const unsigned hmin = __reduce_min_sync(gmask, hi);
    const unsigned lmin = __reduce_min_sync(gmask, hi == hmin ? lo : 0xffffffffu);
    const unsigned holders = (__ballot_sync(0xffffffffu, active && hi == hmin && lo == lmin) >> group_base) & 0xffu;
#else
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const bool is_nan = depth != depth;
    const double key = (active && !is_nan) ? depth : inf;
    double kmin = key;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const double k2 = __shfl_xor_sync(0xffffffffu, kmin, o);
        kmin = (k2 < kmin) ? k2 : kmin;
    }
    const unsigned holders = (__ballot_sync(0xffffffffu, active && key == kmin) >> group_base) & 0xffu;
#endif
    const unsigned nans = (__ballot_sync(0xffffffffu, active && is_nan) >> group_base) & 0xffu;
    if ((nans & 1u) || holders == 0u) return 0;
    return __ffs((int)holders) - 1;
}

struct __align__(16) CoopRes { double depth; int edge_pen; int sep; };   // per (pair, direction): minOverlap's winner

// SORTED: the tile's 32 pairs come from the cell-ordered work list (w_i / w_j / w_a), so neighbouring tiles touch
// the same hulls whatever the slot numbering; results go to the pair's place in the reference order, off[r(i)] + a,
// and pair_i / pair_j are written here.  Otherwise the tile is 32 consecutive pairs of the reference order.
#ifndef COOP_MIN_BLOCKS
#define COOP_MIN_BLOCKS 6
#endif
template <int WMODE>   // = Params::work_mode
__global__ void __launch_bounds__(CO_WARPS * 32, COOP_MIN_BLOCKS) k_manifolds_coop(Params P)
{
    __shared__ int4 s_meta[CO_WARPS][32];                 // per pair of the tile: offset / count of hull A, of hull B
    __shared__ ulonglong2 s_ext[CO_WARPS][32];            // packed extents of both hulls
    __shared__ double2 s_hull[CO_WARPS][2][2][32];        // [buffer][pair of the step][A verts, A normals, B verts, B normals]
    __shared__ CoopRes s_res[CO_WARPS][32][2];            // phase 1 -> phase 2
    constexpr bool SORTED = WMODE == 1, ROWS = WMODE == 2;

    const FrameState *st = P.st;
    if (st->error) return;
    const long long n_pairs = ROWS ? (long long)st->work_cursor : st->n_pairs;
    const double2 *const WV = P.wv;
    const double2 *const WN = P.wn;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, dir = (lane >> 3) & 1, e = lane & 7;
    const int group_base = lane & ~7;                 // first lane of my (pair, direction) group
    const int idx16 = lane & 15;                      // staging: element of the A / B half of my pair's record
    const double2 *const src_arr = (idx16 < 8) ? WV : WN;
    const long long n_tiles = (n_pairs + 31) / 32;

    // cp.async of step t's records into buffer t & 1: lane (q, idx16) copies element idx16 of hull A and of hull B
    auto stage = [&](int t) {
        const int src = 2 * t + half;
        const int4 m = s_meta[warp][src];
        const int k = idx16 & 7;         // (pairs left to the per-thread pass carry vertex counts 0: nothing is copied)
        if (k < m.y) cp_async16(&s_hull[warp][t & 1][half][idx16], src_arr + (unsigned)(m.x + k));
        if (k < m.w) cp_async16(&s_hull[warp][t & 1][half][16 + idx16], src_arr + (unsigned)(m.z + k));
        cp_async_commit();
    };

    for (long long tile = (long long)blockIdx.x * CO_WARPS + warp; tile < n_tiles; tile += (long long)gridDim.x * CO_WARPS) {
        const long long base = tile * 32;
        // ---- prologue: lane L reads the header of pair base + L into shared memory (vertex count 0 = no pair,
        // MAX_STAGED_VERTS + 1 = leave it to the per-thread pass)
        int my_i = 0, my_j = 0, my_oa = 0, my_na = 0, my_ob = 0, my_nb = 0;
        unsigned long long my_xa = 0, my_xb = 0;
        bool fallback = false;
        if (base + lane < n_pairs) {
            if (SORTED || ROWS) { my_i = (int)P.w_i[base + lane]; my_j = (int)P.w_j[base + lane]; }
            else { my_i = P.pair_i[base + lane]; my_j = P.pair_j[base + lane]; }
            const uint4 ha = __ldg(&P.hh[my_i]), hb = __ldg(&P.hh[my_j]);
            my_oa = (int)ha.x; my_na = (int)ha.y; my_ob = (int)hb.x; my_nb = (int)hb.y;
            my_xa = (unsigned long long)ha.z | ((unsigned long long)ha.w << 32);
            my_xb = (unsigned long long)hb.z | ((unsigned long long)hb.w << 32);
            const bool own = ROWS ? (P.mat_stamp[my_i] == (uint32_t)st->frame_no && P.mat_stamp[my_j] == (uint32_t)st->frame_no)
                                  : (my_i >= P.own_lo && my_i < P.own_hi && my_j >= P.own_lo && my_j < P.own_hi);
            fallback = !own || my_na < 1 || my_nb < 1 || my_na > MAX_STAGED_VERTS || my_nb > MAX_STAGED_VERTS;
        }
        // phase 1 sees vertex counts 0 for pairs it must not touch (no pair / per-thread pass)
        s_meta[warp][lane] = make_int4(my_oa, fallback ? 0 : my_na, my_ob, fallback ? 0 : my_nb);
        s_ext[warp][lane] = make_ulonglong2(my_xa, my_xb);
        __syncwarp();
        // ---- phase 1: SAT, two pairs per step, operands staged one step ahead
        stage(0);
#pragma unroll 1
        for (int t = 0; t < 16; ++t) {
            if (t + 1 < 16) { stage(t + 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncwarp();
            const int src = 2 * t + half;
            const int4 m = s_meta[warp][src];
            const int na = m.y, nb = m.w;
            // my direction: E = penetrated hull (its edge normals are the axes), Pn = the other hull
            const int e_n = dir ? nb : na, pn_n = dir ? na : nb;
            const double2 *const rec = s_hull[warp][t & 1][half];
            const double2 *const ev = rec + (dir ? 16 : 0), *const pvs = rec + (dir ? 0 : 16);
            const bool active = e < e_n;
            bool sep = false;
            double depth = 0.0;
            int pen = 0;
            if (active) {
                // overlap sEdge edge sPen (SAT.hs:103-117)
                const ulonglong2 x = s_ext[warp][src];
                const unsigned bits = (unsigned)((dir ? x.y : x.x) >> (6 * e));
                const double2 dn = ev[8 + e];
                const double2 vmin = ev[bits & 7], vmax = ev[(bits >> 3) & 7];
                const V2 d{ dn.x, dn.y };
                // extentAlongSelf (ConvexHull.hs:111-118): the cached extreme vertices only
                const double s_min = dot2(V2{ vmin.x, vmin.y }, d), s_max = dot2(V2{ vmax.x, vmax.y }, d);
                // extentAlong (ConvexHull.hs:81-100): first minimum wins.  The maximum is only ever compared with
                // s_min (overlapTest, SAT.hs:74-83: pMax < sMin), so instead of folding it, `below` keeps "every
                // projection so far is < s_min or unordered" -- with the fold's NaN rules that is pMax < sMin
                // exactly when the first projection and s_min are not NaN (a NaN start freezes the fold's maximum
                // at NaN, later NaNs are skipped by its `>`).
                const double2 p0 = pvs[0];
                const double q0 = dot2(V2{ p0.x, p0.y }, d);
                double p_min = q0;
                bool below = !(q0 >= s_min);
#pragma unroll
                for (int k = 1; k < MAX_STAGED_VERTS; ++k) {
                    if (k >= pn_n) break;
                    const double2 pk = pvs[k];
                    const double q = dot2(V2{ pk.x, pk.y }, d);
                    if (q < p_min) { p_min = q; pen = k; }
                    below = below && !(q >= s_min);
                }
                sep = (p_min > s_max) || (below && q0 == q0 && s_min == s_min);   // overlapTest (SAT.hs:74-83)
                depth = fsub(s_max, p_min);                                       // overlapAmount (SAT.hs:86-96)
            }
            // a separating axis on either side means no contact (contactDebug, SAT.hs:238-248)
            const unsigned sep_mask = __ballot_sync(0xffffffffu, sep);
            const int edge = fold_min_overlap(active, depth, group_base);
            const double best = __shfl_sync(0xffffffffu, depth, group_base + edge);
            const int bpen = __shfl_sync(0xffffffffu, pen, group_base + edge);
            if (e == 0) {
                CoopRes r;
                r.depth = best;
                r.edge_pen = edge | (bpen << 8);
                r.sep = (int)((sep_mask >> (16 * half)) & 0xffffu);
                s_res[warp][src][dir] = r;
            }
            __syncwarp();       // buffer t & 1 is free again before stage(t + 2) refills it
        }
        // ---- phase 2: one pair per lane
        unsigned long long dst = 0ull;      // rows mode: where my pair's body record goes
        unsigned long long dst_hdr = 0ull;  //            and its header
        unsigned n_chunks = 0;              //            and how many of its 16 B chunks are live
        int4 hdr = make_int4(0, 0, 0, 0);
        double4 pj = make_double4(0.0, 0.0, 0.0, 0.0);
        ManRec staged;                      //            the manifold on its way to shared memory
        if (base + lane < n_pairs) {
            PairOut o{ base + lane, P.ccnt, P.man, nullptr, nullptr, nullptr, nullptr, true };
            if (SORTED || ROWS) o = pair_out(P, base + lane, my_i);     // the pair's place in the reference order
            const CoopRes r0 = s_res[warp][lane][0], r1 = s_res[warp][lane][1];
            unsigned cnt = 0;
            if (fallback) cnt = CCNT_FALLBACK;                          // finished by k_manifolds<.., FLAGGED_ONLY>
            else if (!r0.sep) {
                const bool same = r0.depth < r1.depth;                  // depth_ab < depth_ba ? Same : Flip (ties: Flip)
                const int ep = same ? r0.edge_pen : r1.edge_pen;
                ManRec *const out_rec = ROWS ? &staged : pair_out_man(P, o);
                cnt = emit_manifold(out_rec, GlobalAcc{ WV, WN, same ? my_oa : my_ob, same ? my_ob : my_oa },
                                    same ? my_na : my_nb, same ? my_nb : my_na, ep & 0xff, (ep >> 8) & 0xff, same);
            }
            if (ROWS) {
                P.sat_ccnt[base + lane] = cnt;
                if (cnt != CCNT_FALLBACK && o.ok) {
                    hdr = make_int4(my_i, my_j, (int)cnt, 0);
                    if (cnt) pj = partner_record(P, my_j);
                    dst = (unsigned long long)o.rec;
                    dst_hdr = (unsigned long long)o.hdr;
                    n_chunks = cnt ? 7u : 1u;
                }
            } else pair_out_finish(P, o, my_i, my_j, cnt);
        }
        __syncwarp();
        if (ROWS) {
            // The tile's records leave as whole records: staged chunk-major in the (now idle) operand buffers of phase 1
            // -- 16 pairs at a time, no extra shared memory: the kernel lives on its L1 -- and written out 8 lanes per
            // pair, 4 pairs per step: lanes 0..5 the 96 B body (ONE write request), lane 6 the 16 B header.
            int4 *const stage4 = reinterpret_cast<int4 *>(&s_hull[warp][0][0][0]);       // 128 x 16 B = [8 chunks][16 pairs]
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if ((lane >> 4) == h && n_chunks) {
                    const int c16 = lane & 15;
                    stage4[6 * 16 + c16] = hdr;
                    if (n_chunks > 1u) {
                        const int4 *m4 = reinterpret_cast<const int4 *>(&staged);
#pragma unroll
                        for (int q = 0; q < 4; ++q) stage4[q * 16 + c16] = m4[q];
                        stage4[4 * 16 + c16] = make_int4(__double2loint(pj.x), __double2hiint(pj.x), __double2loint(pj.y), __double2hiint(pj.y));
                        stage4[5 * 16 + c16] = make_int4(__double2loint(pj.z), __double2hiint(pj.z), __double2loint(pj.w), __double2hiint(pj.w));
                    }
                }
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r16 = it * 4 + (lane >> 3), ch = lane & 7;
                    const unsigned long long d = __shfl_sync(0xffffffffu, dst, h * 16 + r16);
                    const unsigned long long dh = __shfl_sync(0xffffffffu, dst_hdr, h * 16 + r16);
                    const unsigned nc = __shfl_sync(0xffffffffu, n_chunks, h * 16 + r16);
                    if (nc != 0u) {
                        if (ch == 6) *reinterpret_cast<int4 *>(dh) = stage4[6 * 16 + r16];
                        else if (ch < 6 && nc > 1u) reinterpret_cast<int4 *>(d)[ch] = stage4[ch * 16 + r16];
                    }
                }
                __syncwarp();
            }
        }
    }
}

// Rows mode, home side, after the RESULTS barrier: the pair columns of my slice and the contact counts the row-offset
// scan runs over, from the headers of the records the sweeping ranks stored.
__global__ void __launch_bounds__(256) k_rw_unpack(Params P)
{
    const FrameState *st = P.st;
    if (st->error) return;
    const long long n_pairs = st->n_pairs < P.max_pairs ? st->n_pairs : P.max_pairs;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (long long)gridDim.x * blockDim.x) {
        const int4 h = __ldcs(reinterpret_cast<const int4 *>(&P.phdr[p]));
        P.pair_i[p] = h.x; P.pair_j[p] = h.y; P.ccnt[p] = (uint32_t)h.z;
    }
}

// Row map: row r of the output belongs to manifold point k of pair p, stored as p << 1 | k.
// Also finishes the contact count.  (Runs after the exclusive scan of the per-pair counts.)
__global__ void __launch_bounds__(256) k_row_map(Params P)
{
    FrameState *st = P.st;
    if (st->error) return;
    const long long n_pairs = st->n_pairs;
    unsigned hit = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (long long)gridDim.x * blockDim.x) {
        const unsigned cnt = P.ccnt[p], off = P.coff[p];
        hit += cnt ? 1u : 0u;
        for (unsigned k = 0; k < cnt; ++k)
            if ((long long)off + k < P.max_contacts) P.row_map[off + k] = (uint32_t)((p << 1) | k);
        if (p == n_pairs - 1) {
            const long long total = (long long)off + cnt;
            st->n_contacts = total;
            if (total > P.max_contacts) atomicOr(&st->error, ERR_CONTACT_CAP);
        }
    }
    for (int o = 16; o > 0; o >>= 1) hit += __shfl_xor_sync(0xffffffffu, hit, o);
    if ((threadIdx.x & 31) == 0 && hit) atomicAdd(&st->n_pairs_hit, (unsigned long long)hit);
}

// K3b: one lane per contact ROW, warps own 32 ALIGNED rows.  Every column store of a warp is a
// whole number of 32 B sectors (f64: 256 B, i32: 128 B, u8: 32 B): on B200 a warp store that
// straddles sector boundaries costs ~2.3x (profiles/micro/write_align.cu), so the kernel is
// organised around the output layout and gathers its inputs through the row map.
// flattenContactResult (HullVsHull.hs:54-76) + constraintGen (Constraints/Contact.hs:60-72).
__global__ void __launch_bounds__(256) k_rows(Params P)
{
    const FrameState *st = P.st;
    if (st->error) return;
    const long long n_rows = st->n_contacts < P.max_contacts ? st->n_contacts : P.max_contacts;
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
         row += (long long)gridDim.x * blockDim.x) {
        const uint32_t m = P.row_map[row];
        const long long q = m >> 1;
        const int k = m & 1;
        const int i = P.pair_i[q], j = P.pair_j[q];
        // rows mode: the rank that swept the pair stored the manifold and the partner's position / inverse masses in
        // the pair's record
        const bool rows = P.work_mode == 2;
        const ManRec rec = rows ? P.prec[q].man : P.man[q];
        // i is always an owned slot; j may belong to another rank (raw input columns; rows mode: its home's records)
        const double2 xi = *reinterpret_cast<const double2 *>(&P.xf[i]);
        const double4 pjr = rows ? P.prec[q].pj : make_double4(0.0, 0.0, 0.0, 0.0);
        const bool j_own = j >= P.own_lo && j < P.own_hi;
        const double2 xj = rows ? make_double2(pjr.x, pjr.y)
                         : j_own ? *reinterpret_cast<const double2 *>(&P.xf[j])
                                 : make_double2(in_col(P, 0, P.pos_x, j), in_col(P, 1, P.pos_y, j));
        const double2 mi = P.mass[i], mj = rows ? make_double2(pjr.z, pjr.w) : slot_mass(P, j);
        const int flip = (int)((rec.bits >> 60) & 1u);
        const int edge = (int)(rec.bits & 0xfffffu);
        const int pen = (int)((rec.bits >> (k ? 40 : 20)) & 0xfffffu);
        const V2 n{ rec.nx, rec.ny };
        const V2 c = k ? V2{ rec.c1x, rec.c1y } : V2{ rec.c0x, rec.c0y };
        const V2 pos_i{ xi.x, xi.y }, pos_j{ xj.x, xj.y };
        // contactDepth_ (HullVsHull.hs:30-37): f v - f p, f = afdot' n
        // circle contacts carry their depth (Circle.hs:53, CircleVsHull.hs:54) instead of a reference edge
        const double d = ((rec.bits >> 61) & 1u) ? rec.ref_d : fsub(rec.ref_d, dot2(c, n));
        P.key_i[row] = i; P.key_j[row] = j;
        // flipExtractPair fst (HullVsHull.hs:73-75, Utils.hs:184-186)
        P.feat_a[row] = flip ? pen : edge;
        P.feat_b[row] = flip ? edge : pen;
        P.flip[row] = (uint8_t)flip;
        P.normal_x[row] = n.x; P.normal_y[row] = n.y;
        P.center_x[row] = c.x; P.center_y[row] = c.y;
        P.depth[row] = d;
        // generators run on (penetrated, penetrator) = (a,b) for Same, (b,a) for Flip, and
        // flipExtract swaps the Jacobian halves back (Utils.hs:175-177,212-215; Constraint.hs:96-98)
        const V2 xa = flip ? pos_j : pos_i;
        const V2 xb = flip ? pos_i : pos_j;
        // NonPenetration.jacobian (NonPenetration.hs:34-43)
        const double np_a = cross2(sub2(xa, c), n), np_b = cross2(sub2(c, xb), n);
        // Friction.jacobian (Friction.hs:31-44)
        const V2 tb = clockwise2(n), ta = neg2(tb);
        const double f_a = cross2(sub2(c, xa), ta), f_b = cross2(sub2(c, xb), tb);
        double jn[6], jf[6];
        jn[0] = flip ? n.x : -n.x; jn[1] = flip ? n.y : -n.y; jn[2] = flip ? np_b : np_a;
        jn[3] = flip ? -n.x : n.x; jn[4] = flip ? -n.y : n.y; jn[5] = flip ? np_a : np_b;
        jf[0] = flip ? tb.x : ta.x; jf[1] = flip ? tb.y : ta.y; jf[2] = flip ? f_b : f_a;
        jf[3] = flip ? ta.x : tb.x; jf[4] = flip ? ta.y : tb.y; jf[5] = flip ? f_a : f_b;
#pragma unroll
        for (int t = 0; t < 6; ++t) { P.j_np[t][row] = jn[t]; P.j_f[t][row] = jf[t]; }
        // baumgarte (NonPenetration.hs:48-55)
        P.b_np[row] = (d > P.slop) ? fmul(fdiv(P.baumgarte, P.dt), fsub(P.slop, d)) : 0.0;
        // Restitution.constraintGen (Restitution.hs:21-31): radii from the unflipped pair
        P.ra_x[row] = fsub(c.x, pos_i.x); P.ra_y[row] = fsub(c.y, pos_i.y);
        P.rb_x[row] = fsub(c.x, pos_j.x); P.rb_y[row] = fsub(c.y, pos_j.y);
        P.rn_x[row] = flip ? -n.x : n.x; P.rn_y[row] = flip ? -n.y : n.y;
        // effMassM2 (Constraint.hs:173-179): left fold of (j_k * im_k) * j_k over the unflipped pair
        const double im[6] = { mi.x, mi.x, mi.y, mj.x, mj.x, mj.y };
        double en = fmul(fmul(jn[0], im[0]), jn[0]), ef = fmul(fmul(jf[0], im[0]), jf[0]);
#pragma unroll
        for (int t = 1; t < 6; ++t) {
            en = fadd(en, fmul(fmul(jn[t], im[t]), jn[t]));
            ef = fadd(ef, fmul(fmul(jf[t], im[t]), jf[t]));
        }
        P.inv_eff_np[row] = en; P.inv_eff_f[row] = ef;
    }
}

// ---------------------------------------------------------------------------------------------
// Warm start: the cache join of applyCachedSlns
// ---------------------------------------------------------------------------------------------

// descZipVector (Utils/Descending.hs:47-71) over this frame's contacts and the previous frame's
// (ObjectFeatureKey, ContactLagrangian) cache, both descending with unique keys: a contact whose key
// existed last frame gets the cached Lagrangians (useCache, Solvers/Contact.hs:99-112), any other
// contact gets ContactLagrangian 0 0 (newCache, :85-97).  With unique descending keys the sequential
// two-pointer walk equals an independent lookup per contact: a binary search on the (i, j) columns,
// then the (at most two) rows of that pair are compared on the feature keys.
// Rows are aligned to the output (32 consecutive rows per warp) like k_rows.
__device__ __forceinline__ unsigned long long prev_pair_key(const Params &P, long long t)
{
    return ((unsigned long long)(unsigned)__ldg(&P.pk_i[t]) << 32) | (unsigned)__ldg(&P.pk_j[t]);
}

// first previous row in [lo, hi) whose packed (i, j) is <= key (rows are descending)
__device__ __forceinline__ long long prev_lower_bound(const Params &P, unsigned long long key, long long lo, long long hi)
{
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (prev_pair_key(P, mid) > key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_warm_join(Params P)
{
    const FrameState *st = P.st;
    if (st->error) return;
    const long long n_rows = st->n_contacts < P.max_contacts ? st->n_contacts : P.max_contacts;
    const long long n_prev = *P.n_prev;
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
         row += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = ((unsigned long long)(unsigned)P.key_i[row] << 32) | (unsigned)P.key_j[row];
        const int fa = P.feat_a[row], fb = P.feat_b[row];
        // Contact sets change little from frame to frame, so the match sits near the same relative
        // position: gallop outwards from there (O(log distance) probes) instead of bisecting all
        // n_prev rows.  lo/hi bracket the first previous row whose (i, j) is <= key.
        long long lo = 0, hi = n_prev;
        if (n_prev > 0) {
            long long t = (long long)((double)row * (double)n_prev / (double)n_rows);
            if (t >= n_prev) t = n_prev - 1;
            if (prev_pair_key(P, t) > key) {            // answer is after t
                long long step = 1;
                lo = t + 1;
                while (lo < n_prev) {
                    const long long probe = (lo + step - 1 < n_prev) ? lo + step - 1 : n_prev - 1;
                    if (prev_pair_key(P, probe) > key) { lo = probe + 1; step <<= 1; }
                    else { hi = probe; break; }
                }
            } else {                                     // answer is at or before t
                long long step = 1;
                hi = t;
                while (hi > 0) {
                    const long long probe = (hi - step > 0) ? hi - step : 0;
                    if (prev_pair_key(P, probe) > key) { lo = probe + 1; break; }
                    hi = probe; step <<= 1;
                }
            }
        }
        lo = prev_lower_bound(P, key, lo, hi);
        double np = 0.0, f = 0.0;
        uint8_t hit = 0;
        for (long long t = lo; t < n_prev; ++t) {
            if (prev_pair_key(P, t) != key) break;
            if (P.pk_fa[t] == fa && P.pk_fb[t] == fb) { np = P.cache_np[t]; f = P.cache_f[t]; hit = 1; break; }
        }
        P.warm_np[row] = np; P.warm_f[row] = f; P.warm_hit[row] = hit;
    }
}

// ---------------------------------------------------------------------------------------------
// Rows mode: one frame on G ranks with mapped peer memory (CUDA IPC between processes,
// peer access inside one process).
//
//   home(s)  = rank whose slot block holds s (two folded blocks per rank, or one contiguous block when the slot
//              numbering follows the geometry): owns s's body columns, computes its AABB / cell key, and ends up with
//              the pairs whose LARGER key is s -- the global descending order is a fixed concatenation of the ranks'
//              slices, no merge;
//   sweeper  = rank whose grid-ROW range holds the cell of the pair's larger key: finds the pair and runs SAT on it.
//              Rows are cut so that every rank gets the same number of pairs, measured per row in the previous frame
//              (ROW_BINS bins, exchanged with the CNT barrier): a Gaussian blob is balanced like a uniform world.
//
//   K0 (home slots)  AABB + cell key; a 96 B record (transform, AABB, inverse masses, slot, key + static bit) is
//                    APPENDED over NVLink to the inbox of the rank(s) whose rows + halo hold the cell (big shapes:
//                    of every rank), a warp's records leaving as whole lines                        k_rw_transform
//   -- barrier KEYS (inbox counts; this frame's bounds ride along: they plan the NEXT frame's grid) -- k_rw_sync
//   my inbox: histogram, slot -> record index, big list, kept list; counting sort   k_rw_bin, k_scan_cells_*, k_scatter_sorted
//   world vertices / normals of the kept hulls from the records' transforms -- on a SIDE STREAM,
//   joined before the SAT stage (only SAT reads them)                                               k_rw_hulls
//   single-pass sweep of my rows -> local work list in cell order; every query's partner count
//   is pushed to home(i) (4 B)                                                                      k_sweep<FUSED>, k_big
//   -- barrier CNT (row weights ride along) --
//   home: scan of the counts in descending slot order; each slot's first pair index goes back to
//   its sweeper (4 B)                                                     k_rw_home_counts, scan, k_finish_pairs, k_rw_push_offsets
//   -- barrier OFF (error words ride along) --
//   SAT over the local work list; every pair is STORED into its final place at its home: a dense 16 B header
//   (i, j, contact count) and, when it has contacts, a 96 B body (manifold + the partner's body record)
//   written as one request                                                                          k_manifolds*
//   -- barrier RESULTS --
//   home: pair columns / counts out of the headers, row offsets, k_rows over local memory only,
//   cache join                                                                    k_rw_unpack, scan, k_row_map, k_rows
//   -- barrier COUNTS (every rank's pair / contact totals: global row offsets of the slices) --
// The grid and the row cuts of frame f come from what frame f-1 exchanged (bounds, row weights), so no barrier
// sits between K0 and the keys; the frame number lives in device memory, so frames replay as CUDA graphs.
// ---------------------------------------------------------------------------------------------

// Barrier, arrive side: payload stores, system fence, then this frame's number into every peer's flag word.
__global__ void k_rw_publish(Params P, int phase)
{
    FrameState *st = P.st;
    const int G = P.n_peers, me = P.my_rank;
    if (phase == RW_PHASE_SEED || phase == RW_PHASE_KEYS) {
        // my finite bounds of this frame: frame f + 1 plans its grid from them (SEED: this frame does)
        for (int r = threadIdx.x; r < G; r += blockDim.x) {
            unsigned long long *dst = P.peer_bounds[r] + 4 * me;
            dst[0] = st->bmin_x; dst[1] = st->bmin_y; dst[2] = st->bmax_x; dst[3] = st->bmax_y;
        }
    } else if (phase == RW_PHASE_CNT) {
        // pairs per row bin (balances the next frame's cuts)
        for (int k = threadIdx.x; k < G * ROW_BINS; k += blockDim.x) {
            const int r = k / ROW_BINS, b = k % ROW_BINS;
            P.rw_weights[r][(size_t)me * ROW_BINS + b] = P.roww[b];
        }
    } else if (phase == RW_PHASE_OFF || phase == RW_PHASE_RESULTS) {
        for (int r = threadIdx.x; r < G; r += blockDim.x) P.rw_err[r][(phase == RW_PHASE_OFF ? 0 : G) + me] = st->error;
    } else {
        // (pairs, contacts) of my high block's run and of my low block's run
        const int n_hi = P.rw_hi_hi - P.rw_hi_lo, n_q = n_hi + (P.rw_lo_hi - P.rw_lo_lo);
        long long p_hi = 0, c_hi = 0;
        if (!st->error && n_q > 0) {
            p_hi = n_hi < n_q ? (long long)P.off[n_hi] : st->n_pairs;
            if (n_hi == 0) p_hi = 0;
            c_hi = p_hi < st->n_pairs ? (long long)P.coff[p_hi] : st->n_contacts;
        }
        for (int r = threadIdx.x; r < G; r += blockDim.x) {
            long long *dst = P.rw_counts[r] + 4 * me;
            dst[0] = p_hi; dst[1] = st->n_pairs - p_hi; dst[2] = c_hi; dst[3] = st->n_contacts - c_hi;
            long long *loc = P.rw_counts[r] + 4 * G + 2 * me;
            loc[0] = (long long)st->loc_fold; loc[1] = (long long)st->loc_contig;
        }
    }
    __threadfence_system();
    __syncthreads();
    for (int r = threadIdx.x; r < G; r += blockDim.x)
        *reinterpret_cast<volatile unsigned long long *>(P.peer_flags[r] + phase * SHAPES_MAX_RANKS + me) = st->frame_no;
}

// Barrier, wait side.  Bounded spin: a missing peer turns into an error, not a hang.
__global__ void k_rw_wait(Params P, int phase)
{
    FrameState *st = P.st;
    const int r = threadIdx.x;
    if (r < P.n_peers) {
        const volatile unsigned long long *flag = P.flags + phase * SHAPES_MAX_RANKS + r;
        const long long t0 = clock64();
        while (*flag < st->frame_no) {
            if (clock64() - t0 > 8000000000ll) { atomicOr(&st->error, ERR_PEER_TIMEOUT); break; } // ~4 s
            __nanosleep(200);
        }
    }
    __threadfence_system();
    __syncthreads();
    if ((phase == RW_PHASE_OFF || phase == RW_PHASE_RESULTS) && threadIdx.x == 0) {
        int e = 0;
        for (int q = 0; q < P.n_peers; ++q) e |= reinterpret_cast<volatile int *>(P.rw_err[P.my_rank])[(phase == RW_PHASE_OFF ? 0 : P.n_peers) + q];
        st->peer_error |= e;
        if (e) st->error |= e & (ERR_PAIR_CAP | ERR_PEER_TIMEOUT);   // a full list anywhere voids the frame everywhere
    }
}

// Both sides of a barrier in one launch (one block of 1024 threads): payload + flags out, then the bounded spin on
// my own flag words.  Saves a kernel boundary per barrier on the frame's critical path.
__global__ void __launch_bounds__(1024) k_rw_sync(Params P, int phase)
{
    FrameState *st = P.st;
    const int G = P.n_peers, me = P.my_rank;
    if (phase == RW_PHASE_KEYS) {
        for (int r = threadIdx.x; r < G; r += blockDim.x) {
            unsigned long long *dst = P.peer_bounds[r] + 4 * me;
            dst[0] = st->bmin_x; dst[1] = st->bmin_y; dst[2] = st->bmax_x; dst[3] = st->bmax_y;
            P.rw_inbox_cnt[r][me] = st->push_cursor[r];     // how many records I appended to r's inbox
        }
    } else if (phase == RW_PHASE_CNT) {
        for (int k = threadIdx.x; k < G * ROW_BINS; k += blockDim.x) {
            const int r = k / ROW_BINS, b = k % ROW_BINS;
            P.rw_weights[r][(size_t)me * ROW_BINS + b] = P.roww[b];
        }
    } else if (phase == RW_PHASE_OFF || phase == RW_PHASE_RESULTS) {
        for (int r = threadIdx.x; r < G; r += blockDim.x) P.rw_err[r][(phase == RW_PHASE_OFF ? 0 : G) + me] = st->error;
    } else {
        const int n_hi = P.rw_hi_hi - P.rw_hi_lo, n_q = n_hi + (P.rw_lo_hi - P.rw_lo_lo);
        long long p_hi = 0, c_hi = 0;
        if (!st->error && n_q > 0) {
            p_hi = n_hi < n_q ? (long long)P.off[n_hi] : st->n_pairs;
            if (n_hi == 0) p_hi = 0;
            c_hi = p_hi < st->n_pairs ? (long long)P.coff[p_hi] : st->n_contacts;
        }
        for (int r = threadIdx.x; r < G; r += blockDim.x) {
            long long *dst = P.rw_counts[r] + 4 * me;
            dst[0] = p_hi; dst[1] = st->n_pairs - p_hi; dst[2] = c_hi; dst[3] = st->n_contacts - c_hi;
            long long *loc = P.rw_counts[r] + 4 * G + 2 * me;
            loc[0] = (long long)st->loc_fold; loc[1] = (long long)st->loc_contig;
        }
    }
    __threadfence_system();
    __syncthreads();
    const unsigned long long frame = st->frame_no;
    for (int r = threadIdx.x; r < G; r += blockDim.x)
        *reinterpret_cast<volatile unsigned long long *>(P.peer_flags[r] + phase * SHAPES_MAX_RANKS + me) = frame;
    // ---- wait side
    if ((int)threadIdx.x < G) {
        const volatile unsigned long long *flag = P.flags + phase * SHAPES_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while (*flag < frame) {
            if (clock64() - t0 > 8000000000ll) { atomicOr(&st->error, ERR_PEER_TIMEOUT); break; } // ~4 s
            __nanosleep(100);
        }
    }
    __threadfence_system();
    __syncthreads();
    if ((phase == RW_PHASE_OFF || phase == RW_PHASE_RESULTS) && threadIdx.x == 0) {
        int e = 0;
        for (int q = 0; q < G; ++q) e |= reinterpret_cast<volatile int *>(P.rw_err[me])[(phase == RW_PHASE_OFF ? 0 : G) + q];
        st->peer_error |= e;
        if (e) st->error |= e & (ERR_PAIR_CAP | ERR_PEER_TIMEOUT);   // a full list anywhere voids the frame everywhere
    }
}

// First kernel of a rows-mode frame (one block).  Frame counter, the grid from the bounds every rank pushed in the
// previous frame (two cells of margin: whatever moved further goes to the exact big-shape path), the row cuts from
// the row weights of the previous frame, and the per-frame counters.
__global__ void __launch_bounds__(1024) k_rw_begin(Params P, int advance)
{
    FrameState *st = P.st;
    __shared__ unsigned long long s_pre[1024];
    const int G = P.n_peers, t = threadIdx.x;
    if (t == 0) {
        if (advance) st->frame_no += 1;
        st->bmin_x = st->bmin_y = ~0ull;
        st->bmax_x = st->bmax_y = 0ull;
        for (int r = 0; r < G; ++r) {
            const unsigned long long *b = P.rw_bounds_prev + 4 * r;
            if (b[0] < st->bmin_x) st->bmin_x = b[0];
            if (b[1] < st->bmin_y) st->bmin_y = b[1];
            if (b[2] > st->bmax_x) st->bmax_x = b[2];
            if (b[3] > st->bmax_y) st->bmax_y = b[3];
        }
        plan_grid_from_bounds(st, P.cell_size, P.cell_limit, 2.0);
        st->bmin_x = st->bmin_y = ~0ull;       // K0 reduces this frame's bounds into them
        st->bmax_x = st->bmax_y = 0ull;
        st->n_big = 0; st->n_small = 0; st->error = 0; st->peer_error = 0;
        st->n_pairs = 0; st->n_contacts = 0; st->work_cursor = 0ull; st->n_pairs_hit = 0ull; st->n_kept = 0u; st->n_list = 0u;
        for (int r = 0; r < SHAPES_MAX_RANKS; ++r) st->push_cursor[r] = 0u;
        st->loc_fold = 0ull; st->loc_contig = 0ull;
    }
    // row weights: ROW_BINS bins over the rows, summed over the ranks that measured them; 4 bins per thread
    unsigned long long mine = 0;
    for (int b = 4 * t; b < 4 * t + 4 && b < ROW_BINS; ++b)
        for (int r = 0; r < G; ++r) mine += P.rw_weights_prev[(size_t)r * ROW_BINS + b];
    s_pre[t] = mine;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {       // inclusive scan (Hillis-Steele; one block, once per frame)
        const unsigned long long v = t >= o ? s_pre[t - o] : 0ull;
        __syncthreads();
        s_pre[t] += v;
        __syncthreads();
    }
    if (t == 0) {
        const int H = st->H;
        const unsigned long long total = s_pre[1023];
        st->cut[0] = 0;
        for (int g = 1; g < G; ++g) {
            int row;
            if (total < (unsigned long long)(64 * G)) row = (int)(((long long)H * g) / G);   // nothing measured yet: equal rows
            else {
                // first group of 4 bins whose running total reaches g / G of the work
                const unsigned long long want = (total * (unsigned long long)g) / (unsigned long long)G;
                int lo = 0, hi = 1023;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_pre[mid] >= want) hi = mid; else lo = mid + 1; }
                row = (int)(((long long)(4 * lo + 2) * H) / ROW_BINS);
            }
            if (row < st->cut[g - 1]) row = st->cut[g - 1];
            if (row > H) row = H;
            st->cut[g] = row;
        }
        st->cut[G] = H;
        st->row_lo = st->cut[P.my_rank]; st->row_hi = st->cut[P.my_rank + 1];
        const int k_lo = st->row_lo > 0 ? st->row_lo - 1 : 0, k_hi = st->row_hi < H ? st->row_hi + 1 : H;
        st->cell_lo = (unsigned)k_lo * (unsigned)st->W;
        st->cell_end = (st->row_hi > st->row_lo) ? (unsigned)k_hi * (unsigned)st->W : st->cell_lo;   // no rows: keep nothing
    }
    for (int b = t; b < ROW_BINS; b += blockDim.x) P.roww[b] = 0u;
}

// K0 of a rows-mode frame, one thread per HOME slot (my low block, then my high block): packed transform, inverse
// masses, AABB, this rank's finite bounds, and the cell key -- bit 31 = isStatic, RW_KEY_* encoding.  The records stay
// in my arena (k_rows and k_big read them) and are PUSHED, 4 + 32 + 16 bytes, to the ranks whose rows (plus halo)
// contain the shape's cell -- to every rank for big shapes: after the KEYS barrier a sweeping rank finds everything it
// keeps in its own memory.  World vertices are not materialised here: that is the sweeping rank's job (k_rw_hulls).
template <bool BOUNDS_ONLY>
__global__ void __launch_bounds__(256) k_rw_transform(Params P, int n_home)
{
    const FrameState *st = P.st;
    const int n_lo = P.rw_lo_hi - P.rw_lo_lo;
    const int G = P.n_peers;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    __shared__ unsigned s_cnt[8][SHAPES_MAX_RANKS];     // per warp and destination: records of this tile, then their offset
    __shared__ unsigned s_base[SHAPES_MAX_RANKS];       // per destination: the tile's first index in my section of its inbox
    __shared__ HomeRec s_rec[8][32];                    // per warp: the records going to one destination, compacted
    double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    // the block walks whole 256-slot tiles (block-uniform trip count: the reservation below synchronises the block)
    for (int tile = blockIdx.x * blockDim.x; tile < n_home; tile += gridDim.x * blockDim.x) {
        const int t = tile + (int)threadIdx.x;
        const bool valid = t < n_home;
        const int s = !valid ? 0 : (t < n_lo ? P.rw_lo_lo + t : P.rw_hi_lo + (t - n_lo));
        uint32_t key = RW_KEY_NONE;
        int cy = -1;
        bool live = false;
        Xf x{ 0.0, 0.0, 1.0, 0.0 };
        double2 mass = make_double2(0.0, 0.0);
        Box b{ 0.0, 0.0, 0.0, 0.0 };
        if (valid) {
            const double px = P.pos_x[s], py = P.pos_y[s];
            const double il = P.inv_lin[s], ir = P.inv_rot[s];
            live = P.alive[s] != 0;
            const int o = P.vert_offset[s];
            const int n = P.vert_offset[s + 1] - o;
            const double rad = P.radius ? P.radius[s] : -1.0;
            double c, sn;
            if (P.cos_rot) { c = P.cos_rot[s]; sn = P.sin_rot[s]; }
            else sincos(P.rot[s], &sn, &c);
            x = Xf{ px, py, c, sn };
            mass = make_double2(il, ir);
            if (!BOUNDS_ONLY) { P.xf[s] = x; P.mass[s] = mass; }
            if (live) {
                const Aff m = to_transform(px, py, c, sn);
                if (rad >= 0.0) {   // setCircleTransform (Circle.hs:55-59), circleToAabb (Aabb.hs:86-88)
                    const V2 ctr = afmul(m, V2{ 0.0, 0.0 });
                    b.min_x = fsub(ctr.x, rad); b.max_x = fadd(ctr.x, rad);
                    b.min_y = fsub(ctr.y, rad); b.max_y = fadd(ctr.y, rad);
                }
                // hullToAabb (Aabb.hs:81-84): foldl1 mergeAabb.  Hulls of up to 8 vertices: the local vertices as ONE batch
                // of loads (a load per loop iteration made this kernel latency-bound: 0.15 ms for 2M slots)
                if (n <= MAX_STAGED_VERTS) {
                    double2 l[MAX_STAGED_VERTS];
#pragma unroll
                    for (int k = 0; k < MAX_STAGED_VERTS; ++k) if (k < n) l[k] = __ldg(&P.local[o + k]);
#pragma unroll
                    for (int k = 0; k < MAX_STAGED_VERTS; ++k) {
                        if (k >= n) break;
                        const V2 w = afmul(m, V2{ l[k].x, l[k].y });
                        if (k == 0) { b.min_x = b.max_x = w.x; b.min_y = b.max_y = w.y; }
                        else {
                            b.min_x = (b.min_x < w.x) ? b.min_x : w.x; b.max_x = (b.max_x > w.x) ? b.max_x : w.x;
                            b.min_y = (b.min_y < w.y) ? b.min_y : w.y; b.max_y = (b.max_y > w.y) ? b.max_y : w.y;
                        }
                    }
                } else
                for (int k = 0; k < n; ++k) {
                    const double2 l = __ldg(&P.local[o + k]);
                    const V2 w = afmul(m, V2{ l.x, l.y });
                    if (k == 0) { b.min_x = b.max_x = w.x; b.min_y = b.max_y = w.y; }
                    else {
                        b.min_x = (b.min_x < w.x) ? b.min_x : w.x; b.max_x = (b.max_x > w.x) ? b.max_x : w.x;
                        b.min_y = (b.min_y < w.y) ? b.min_y : w.y; b.max_y = (b.max_y > w.y) ? b.max_y : w.y;
                    }
                }
                if (finite4(b)) {
                    mnx = fmin(mnx, b.min_x); mxx = fmax(mxx, b.max_x);
                    mny = fmin(mny, b.min_y); mxy = fmax(mxy, b.max_y);
                }
                if (!BOUNDS_ONLY) {
                    int cx;
                    key = small_cell(b, st, cx, cy) ? RW_KEY_BASE + (uint32_t)cy * (uint32_t)st->W + (uint32_t)cx : RW_KEY_BIG;
                    if (key == RW_KEY_BIG) cy = -1;
                    if (il == 0.0 && ir == 0.0) key |= KEY_STATIC_BIT;      // isStatic (Constraint.hs:123-125)
                }
            }
            if (!BOUNDS_ONLY) { P.box[s] = b; P.gkeys[s] = key; }
        }
        if (BOUNDS_ONLY) continue;
        // ---- who sweeps this shape: rank g keeps rows [cut[g] - 1, cut[g + 1]] when it sweeps any row at all; big
        // shapes go to every rank (mine included: the inbox is the only way into a rank's grid)
        unsigned wants = 0;
        if (live)
            for (int g = 0; g < G; ++g)
                if (cy < 0 || (st->cut[g + 1] > st->cut[g] && cy >= st->cut[g] - 1 && cy <= st->cut[g + 1])) wants |= 1u << g;
        if (P.dbg_local_stores) wants &= 1u << P.my_rank;       // experiment: nothing leaves this GPU (WRONG results, timing only)
        for (int g = 0; g < G; ++g) {
            const unsigned bal = __ballot_sync(0xffffffffu, (wants >> g) & 1u);
            if (lane == 0) s_cnt[warp][g] = (unsigned)__popc(bal);
        }
        __syncthreads();
        if ((int)threadIdx.x < G) {      // one reservation per destination and tile
            const int g = (int)threadIdx.x;
            unsigned tot = 0;
            for (int w = 0; w < 8; ++w) { const unsigned c = s_cnt[w][g]; s_cnt[w][g] = tot; tot += c; }
            s_base[g] = tot ? atomicAdd(&P.st->push_cursor[g], tot) : 0u;
        }
        __syncthreads();
        // ---- per destination: the warp's records compacted in shared memory, then copied out 16 B per lane, so that
        // they leave as contiguous 512 B runs (whole lines over NVLink) -- one 96 B record per shape instead of three packets
        for (int g = 0; g < G; ++g) {
            const unsigned bal = __ballot_sync(0xffffffffu, (wants >> g) & 1u);
            if (bal == 0u) continue;
            if ((wants >> g) & 1u) {
                HomeRec r;
                r.xf = x; r.box = b; r.mass = mass; r.slot = (uint32_t)s; r.key = key; r.pad[0] = 0u; r.pad[1] = 0u;
                s_rec[warp][__popc(bal & lt_mask)] = r;
            }
            __syncwarp();
            const unsigned first = s_base[g] + s_cnt[warp][g];
            const int n_chunks = HOME_REC_CHUNKS * __popc(bal);
            if (first + (unsigned)__popc(bal) <= (unsigned)P.inbox_cap) {       // (cannot overflow: a section holds a whole home)
                int4 *dst = reinterpret_cast<int4 *>(P.rw_inbox[g] + ((size_t)P.my_rank * (size_t)P.inbox_cap + first));
                const int4 *src = reinterpret_cast<const int4 *>(&s_rec[warp][0]);
                for (int c = lane; c < n_chunks; c += 32) dst[c] = src[c];
            }
            __syncwarp();
        }
        __syncthreads();     // s_cnt / s_base are rewritten by the next tile
    }
    __shared__ double s_red[4][8];
    mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
    if (lane == 0) { s_red[0][warp] = mnx; s_red[1][warp] = mny; s_red[2][warp] = mxx; s_red[3][warp] = mxy; }
    __syncthreads();
    if (warp == 0) {
        mnx = lane < 8 ? s_red[0][lane] : INFINITY; mny = lane < 8 ? s_red[1][lane] : INFINITY;
        mxx = lane < 8 ? s_red[2][lane] : -INFINITY; mxy = lane < 8 ? s_red[3][lane] : -INFINITY;
        mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
        if (lane == 0 && mnx <= mxx) {
            atomicMin(&P.st->bmin_x, enc_ordered(mnx)); atomicMin(&P.st->bmin_y, enc_ordered(mny));
            atomicMax(&P.st->bmax_x, enc_ordered(mxx)); atomicMax(&P.st->bmax_y, enc_ordered(mxy));
        }
    }
}

// The records my inbox received (one section per home rank, counts published with the KEYS barrier): every shape
// whose cell lies in my rows or the halo row on either side, and every big shape.  Histogram of the cell table (the
// arrival order is the counting sort's scatter slot), slot -> record index (rec_of, kept in keys[]), the big list,
// and the kept list of record indices (one reservation per block, the inbox order inside it: sections are in ascending
// slot order up to the interleaving of the sender's blocks).
__global__ void __launch_bounds__(256) k_rw_bin(Params P)
{
    const FrameState *st = P.st;
    const unsigned c_lo = st->cell_lo, c_end = st->cell_end;
    const int G = P.n_peers;
    __shared__ unsigned s_wsum[8];
    __shared__ unsigned s_base;
    __shared__ unsigned s_pre[SHAPES_MAX_RANKS + 1];
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int h = 0; h < G; ++h) { s_pre[h] = t; t += min(P.inbox_cnt[h], (unsigned)P.inbox_cap); }
        s_pre[G] = t;
    }
    __syncthreads();
    const unsigned total = s_pre[G];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        const unsigned e = base + threadIdx.x;
        bool keep = false;
        int s = 0;
        unsigned at = 0, cell_rank = 0;       // the record's index in my inbox; its arrival order within its cell
        if (e < total) {
            int h = 0;
            while (e >= s_pre[h + 1]) ++h;
            at = (unsigned)h * (unsigned)P.inbox_cap + (e - s_pre[h]);
            // slot and key sit in the record's last 16 B: one sector per record, nothing is copied for a small shape --
            // the later kernels read the record itself (through the kept list, or rec_of[slot] for a pair's partner)
            const uint4 tail = *reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(&P.inbox[at]) + 80);
            s = (int)tail.x;
            const uint32_t enc = tail.y & ~KEY_STATIC_BIT;
            P.keys[s] = at;                   // rows mode: keys[] is rec_of[] (slot -> inbox record)
            if (enc == RW_KEY_BIG) {
                // big shapes (few): slot-indexed copies, read by the sweep's big-candidate loop and the hull pass
                const HomeRec r = P.inbox[at];
                P.xf[s] = r.xf; P.mass[s] = r.mass; P.gkeys[s] = r.key;
                if (!rw_mine(P, s)) P.box[s] = r.box;       // (home slots: K0 stored it; peers read that copy)
                const unsigned pos = atomicAdd(&P.st->n_big, 1u);
                P.big_idx[pos] = (uint32_t)s;
                if (pos >= P.big_limit) atomicOr(&P.st->error, ERR_REPLAN);
            } else if (enc >= RW_KEY_BASE) {
                const uint32_t key = enc - RW_KEY_BASE;
                if (key >= c_lo && key < c_end) {
                    cell_rank = atomicAdd(&P.cell_count[key], 1u);
                    keep = true;
                }
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wsum[warp] = (unsigned)__popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int w = 0; w < 8; ++w) { const unsigned c = s_wsum[w]; s_wsum[w] = t; t += c; }
            s_base = t ? atomicAdd(&P.st->n_list, t) : 0u;
        }
        __syncthreads();
        if (keep) {
            const unsigned pos = s_base + s_wsum[warp] + (unsigned)__popc(bal & ((1u << lane) - 1u));
            P.kept_list[pos] = at;            // the kept list holds inbox indices ...
            P.rank[pos] = cell_rank;          // ... and, next to them, each shape's scatter slot within its cell
        }
        __syncthreads();
    }
}

// Exclusive scan of cell_count over the kept cell range [cell_lo, cell_end) (device side bounds) into cell_begin,
// cell_begin[cell_end] = total.  Two passes over a small table: chunk sums, then each block scans its chunk.
constexpr int SCAN_BLOCKS = 296, SCAN_THREADS = 256;
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_cells_sums(Params P, unsigned *chunk_sum)
{
    const FrameState *st = P.st;
    const unsigned lo = st->cell_lo, len = st->cell_end - st->cell_lo;
    const unsigned chunk = (len + SCAN_BLOCKS - 1) / SCAN_BLOCKS;
    const unsigned a = min(len, blockIdx.x * chunk), b = min(len, a + chunk);
    unsigned sum = 0;
    for (unsigned k = a + threadIdx.x; k < b; k += SCAN_THREADS) sum += P.cell_count[lo + k];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __shared__ unsigned s_w[SCAN_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = 0; for (int w = 0; w < SCAN_THREADS / 32; ++w) t += s_w[w]; chunk_sum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_cells_apply(Params P, const unsigned *chunk_sum)
{
    FrameState *st = P.st;
    const unsigned lo = st->cell_lo, len = st->cell_end - st->cell_lo;
    const unsigned chunk = (len + SCAN_BLOCKS - 1) / SCAN_BLOCKS;
    const unsigned a = min(len, blockIdx.x * chunk), b = min(len, a + chunk);
    __shared__ unsigned s_w[SCAN_THREADS / 32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (unsigned q = 0; q < blockIdx.x; ++q) t += chunk_sum[q];
        s_carry = t;
        if (blockIdx.x == SCAN_BLOCKS - 1) { P.cell_begin[lo + len] = t + chunk_sum[blockIdx.x]; st->n_kept = t + chunk_sum[blockIdx.x]; }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned base = a; base < b; base += SCAN_THREADS) {
        const unsigned k = base + threadIdx.x;
        const unsigned v = k < b ? P.cell_count[lo + k] : 0u;
        unsigned inc = v;
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) { if (w < warp) before += s_w[w]; total += s_w[w]; }
        const unsigned carry = s_carry;
        if (k < b) P.cell_begin[lo + k] = carry + before + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

// moveShapes (World.hs:132-140) for the shapes this rank keeps (its rows, the halo rows, the big list): one thread
// per shape, walking the kept-slot list (ascending slots inside every 256-entry run: the gathers of the static
// geometry and the stores of the world vertices / normals are as dense as K0's, whatever share of the world this
// rank keeps).  Same body as K0 (k_transform_aabb) without the AABB (that came with the shape's record): the hull's
// local vertices as one batch of loads, world vertices kept in registers for the unit normals (setHullTransform,
// ConvexHull.hs:184-195: normals recomputed from the NEW vertices), with the packed transform the home pushed.
// Only the SAT stage reads what this kernel writes, so it runs on a second stream next to the sweep, the CNT / OFF
// barriers and the homes' offset scan -- a chain of small, latency-bound kernels that leaves the GPU idle otherwise.
__global__ void __launch_bounds__(256) k_rw_hulls(Params P)
{
    const FrameState *st = P.st;
    if (st->error & ERR_REPLAN) return;
    const unsigned n_kept = st->n_list, n_all = n_kept + st->n_big;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n_all; q += gridDim.x * blockDim.x) {
        int s;
        Xf x;
        if (q < n_kept) { const HomeRec &r = P.inbox[P.kept_list[q]]; s = (int)r.slot; x = r.xf; }   // the record its home sent
        else { s = (int)P.big_idx[q - n_kept]; x = P.xf[s]; }
        const int o = P.vert_offset[s], n = P.vert_offset[s + 1] - o;
        const double rad = P.radius ? P.radius[s] : -1.0;
        P.mat_stamp[s] = (uint32_t)st->frame_no;
        const Aff m = to_transform(x.px, x.py, x.c, x.s);
        if (rad >= 0.0) {      // setCircleTransform (Circle.hs:55-59)
            const V2 ctr = afmul(m, V2{ 0.0, 0.0 });
            P.circ[s] = make_double2(ctr.x, ctr.y);
        }
        if (n <= MAX_STAGED_VERTS) {
            double2 l[MAX_STAGED_VERTS];
#pragma unroll
            for (int k = 0; k < MAX_STAGED_VERTS; ++k) if (k < n) l[k] = __ldg(&P.local[o + k]);
            V2 w[MAX_STAGED_VERTS];
#pragma unroll
            for (int k = 0; k < MAX_STAGED_VERTS; ++k) {
                if (k >= n) break;
                w[k] = afmul(m, V2{ l[k].x, l[k].y });
                P.wv[o + k] = make_double2(w[k].x, w[k].y);
            }
#pragma unroll
            for (int k = 0; k < MAX_STAGED_VERTS; ++k) {
                if (k >= n) break;
                const V2 nxt = (k + 1 < MAX_STAGED_VERTS && k + 1 < n) ? w[(k + 1) & (MAX_STAGED_VERTS - 1)] : w[0];
                const V2 nn = unit_edge_normal(w[k], nxt);
                P.wn[o + k] = make_double2(nn.x, nn.y);
            }
        } else {
            V2 w0{ 0.0, 0.0 }, prev{ 0.0, 0.0 };
            for (int v = 0; v < n; ++v) {
                const double2 l = __ldg(&P.local[o + v]);
                const V2 w = afmul(m, V2{ l.x, l.y });
                P.wv[o + v] = make_double2(w.x, w.y);
                if (v == 0) w0 = w;
                else { const V2 nn = unit_edge_normal(prev, w); P.wn[o + v - 1] = make_double2(nn.x, nn.y); }
                prev = w;
            }
            if (n > 0) { const V2 nn = unit_edge_normal(prev, w0); P.wn[o + n - 1] = make_double2(nn.x, nn.y); }
        }
    }
}

// Home side, after the CNT barrier: the counts the sweeping ranks pushed for my slots, in the descending order the
// scan runs in.
__global__ void __launch_bounds__(256) k_rw_home_counts(Params P, int n_query)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_query; r += gridDim.x * blockDim.x)
        P.cnt[r] = (unsigned long long)(P.rw_cq[P.my_rank][rw_qslot(P, r)] & 0x0fffffffu);
}

// Home side, after the scan: every slot's first pair index goes back to the rank that swept it (4 B over NVLink), so
// that its SAT stage stores each pair straight into its final place here.
__global__ void __launch_bounds__(256) k_rw_push_offsets(Params P, int n_query)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_query; r += gridDim.x * blockDim.x) {
        const int i = rw_qslot(P, r);
        const uint32_t w = P.rw_cq[P.my_rank][i];
        if ((w & 0x0fffffffu) == 0u) continue;
        const unsigned long long off = P.off[r];
        P.rw_qoff[w >> 28][i] = off < 0xffffffffull ? (uint32_t)off : 0xffffffffu;
    }
}

// ---------------------------------------------------------------------------------------------
// static geometry kernels
// ---------------------------------------------------------------------------------------------

// _hullExtents (ConvexHull.hs:151-167): per edge, (argmin, argmax) of the LOCAL vertices along
// the local unit edge normal, first minimum / first maximum on ties.  One thread per edge.
__global__ void k_hull_extents(int n_slots, const int32_t *vert_offset, const double2 *local,
                               int32_t *ext_min, int32_t *ext_max)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int o = vert_offset[s], n = vert_offset[s + 1] - o;
    for (int e = 0; e < n; ++e) {
        const int e1 = (e < n - 1) ? e + 1 : 0;
        const double2 a = local[o + e], b = local[o + e1];
        const V2 dir = unit_edge_normal(V2{ a.x, a.y }, V2{ b.x, b.y });
        double vmin = 0.0, vmax = 0.0;
        int imin = 0, imax = 0;
        for (int k = 0; k < n; ++k) {
            const double2 v = local[o + k];
            const double d = dot2(V2{ v.x, v.y }, dir);
            if (k == 0) { vmin = vmax = d; }
            else {
                if (d < vmin) { vmin = d; imin = k; }
                if (d > vmax) { vmax = d; imax = k; }
            }
        }
        ext_min[o + e] = imin;
        ext_max[o + e] = imax;
    }
}

__global__ void k_pack_extents(int n_slots, const int32_t *vert_offset, const int32_t *ext_min,
                               const int32_t *ext_max, unsigned long long *packed, uint4 *hh)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int o = vert_offset[s], n = vert_offset[s + 1] - o;
    unsigned long long w = 0;
    if (n <= MAX_STAGED_VERTS)
        for (int e = 0; e < n; ++e)
            w |= ((unsigned long long)(ext_min[o + e] & 7) | ((unsigned long long)(ext_max[o + e] & 7) << 3)) << (6 * e);
    packed[s] = w;
    hh[s] = make_uint4((unsigned)o, (unsigned)n, (unsigned)w, (unsigned)(w >> 32));
}

__global__ void k_split_boxes(int n, const Box *box, double *a, double *b, double *c, double *d)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const Box x = box[s];
    a[s] = x.min_x; b[s] = x.max_x; c[s] = x.min_y; d[s] = x.max_y;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// NCCL is resolved with dlopen/dlsym the first time a multi-rank ctx (or a unique id) is asked
// for.  A process that already carries an NCCL (e.g. PyTorch's bundled copy) keeps using that
// one; single-GPU use never loads NCCL at all.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool ok = false;
};

static NcclApi &nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) { api.error = std::string("dlopen libnccl.so.2: ") + dlerror(); return api; }
    bool all = true;
    auto sym = [&](const char *name) { void *p = dlsym(h, name); if (!p) { all = false; api.error = std::string("dlsym ") + name; } return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = all;
    return api;
}

struct FrameKey {   // everything a captured frame graph bakes in
    int64_t n; const double *in[7]; double dt, baumgarte, slop, cell; bool world, profiling, warm, p2p, remote, sorted; int64_t geometry, n_prev; unsigned cell_limit, big_limit;
};

struct WorldStep;   // device-resident world (world_step.cuh)

struct shapes_ctx {
    WorldStep *ws = nullptr;
    int device = 0;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;      // rows mode: the hull pass runs here, next to the sweep and the offset exchange
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool profiling = false;
    cudaEvent_t stage_ev[SHAPES_N_STAGES + 1] = {};
    float stage_ms[SHAPES_N_STAGES] = {};
    int sm_count = 148;
    int64_t max_shapes = 0, max_verts = 0, max_pairs = 0, max_contacts = 0;
    int64_t n_slots = 0, n_verts = 0;
    int64_t chunk = 0;          // slots per rank (all-gather granule)
    bool hulls_set = false;
    bool use_graph = true;
    cudaGraphExec_t graph_exec[4] = { nullptr, nullptr, nullptr, nullptr }; // per key-buffer parity (x frame parity in rows mode)
    FrameKey graph_key[4] = {};
    int64_t graph_launches = 0;
    int64_t geometry_version = 0;
    int max_hull_verts = 0;
    unsigned cell_limit = 0;     // sticky per-frame cell budget (0 = not chosen yet)
    int ct_blocks[3] = { 4, 4, 4 }; // resident k_manifolds blocks per SM (boxes / general / with circles)
    int coop_blocks = 4;          // resident k_manifolds_coop blocks per SM
    int coop_blocks_rows = 4;     // ... of the rows-mode instantiation (more shared memory: the staged pair records)
    bool use_coop = true;         // general polygons: 16 lanes per pair (SHAPES_B200_NO_COOP=1: one thread per pair)
    // rows mode (multi-rank with mapped peers): one exchange arena per rank, same layout everywhere
    bool use_rows = true;         // SHAPES_B200_NO_ROWS=1: slot-range ownership of the whole path (the r1 exchange)
    bool rows_ready = false;      // every peer's arena is mapped
    bool rows_fold = true;        // home layout: folded blocks (any slot numbering) / contiguous chunks (numbering follows the geometry)
    int rows_frames = 0;          // rows-mode frames since the geometry was set (the layout decision is taken after the second)
    char *rw_arena = nullptr;
    char *peer_arena[SHAPES_MAX_RANKS] = {};
    struct RowsLayout {
        size_t gkeys[2], box, xf, mass, in[7], cq, qoff, bounds[2], weights[2], counts, err, flags, prec, phdr, inbox, inbox_cnt, total;
    } rwl{};
    uint32_t *d_roww = nullptr, *d_ccnt_w = nullptr, *d_w_j = nullptr;
    int32_t *d_pair_i = nullptr, *d_pair_j = nullptr; uint32_t *d_ccnt = nullptr; ManRec *d_man = nullptr;   // single-rank homes of the result arrays
    Xf *d_xf = nullptr;
    double2 *d_mass = nullptr;
    unsigned *d_chunk_sum = nullptr;
    uint32_t *d_mat_stamp = nullptr, *d_kept_list = nullptr;
    bool pending_warm = false, pending_seed = false, pending_rows = false, pending_plan_ahead = false;
    bool use_plan_ahead = true;   // single rank: grid planned from the previous frame's bounds, K0 keys and bins (SHAPES_B200_NO_PLAN_AHEAD=1: plan inside the frame)
    bool plan_valid = false;      // the bounds in FrameState describe the last completed frame of the current geometry
    int64_t big_seen = 0;         // big-list length of the last frame that ran on an exactly seeded plan
    bool use_sorted = true;       // general polygons: single-pass sweep + SAT work list in cell order (SHAPES_B200_NO_SORTED=1: two-pass sweep, SAT in the reference's pair order)
    uint4 *d_hh = nullptr;        // static hull headers (Params::hh)
    bool has_circles = false;
    double *d_radius = nullptr;
    int rows_blocks = 4;         // resident k_rows blocks per SM
    double auto_cell = 1.0, user_cell = 0.0;
    int64_t launches = 0;
    std::string err;
    std::vector<void *> allocs;
    // device buffers
    uint8_t *d_alive = nullptr;
    int32_t *d_vert_offset = nullptr, *d_ext_min = nullptr, *d_ext_max = nullptr;
    double2 *d_local = nullptr;
    unsigned long long *d_ext_packed = nullptr;
    double *d_in[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    double *d_world_x = nullptr, *d_world_y = nullptr, *d_split = nullptr;
    void *d_scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    FrameState *h_state = nullptr; // pinned
    long long *d_n_prev = nullptr; // rows of the previous frame's key columns (read by k_warm_join)
    int64_t *d_counts = nullptr;   // world x 2 (pairs, contacts), all-gathered
    int64_t *h_counts = nullptr;   // pinned
    Params P{};
    // peer-to-peer exchange: boxes are double buffered by frame parity so that a rank running one
    // frame ahead never overwrites records a slower rank is still reading
    Box *d_box2[2] = { nullptr, nullptr };
    uint32_t *d_gkeys2[2] = { nullptr, nullptr };
    double *d_in2[2][7] = {};   // body columns, double buffered by frame parity (multi-rank)
    const double *peer_in2[2][7][SHAPES_MAX_RANKS] = {};
    uint32_t *peer_keys2[2][SHAPES_MAX_RANKS] = {};
    unsigned long long *d_bounds2[2] = { nullptr, nullptr };
    unsigned long long *d_flags = nullptr;
    bool peers_ready = false;
    bool use_p2p = true;
    Box *peer_box2[2][SHAPES_MAX_RANKS] = {};
    unsigned long long *peer_bounds2[2][SHAPES_MAX_RANKS] = {};
    unsigned long long *peer_flags[SHAPES_MAX_RANKS] = {};
    std::vector<void *> ipc_opened;
    unsigned long long frame_no = 0;
    // warm start: the other half of the double-buffered key columns, and the cache
    int32_t *alt_key[4] = { nullptr, nullptr, nullptr, nullptr };
    double *d_cache_np = nullptr, *d_cache_f = nullptr;
    int64_t n_prev_keys = 0;      // rows of the previous completed frame (keys live in alt_key after the swap)
    bool cache_valid = false;     // a Lagrangian cache for those rows has been supplied
    bool warm_done = false;       // the last frame ran the join
    int parity = 0;
    // last frame
    int64_t last_pairs = 0, last_contacts = 0;
    // SHAPES_B200_KERNEL_TIMES=1: per-kernel CUDA events inside profiled rows-mode frames, averages printed by shapes_destroy
    bool dbg_times = false;
    std::vector<cudaEvent_t> dbg_ev;
    std::vector<const char *> dbg_name;
    std::vector<double> dbg_sum;
    int dbg_marks = 0, dbg_marks_seen = 0, dbg_frames = 0;
    bool have_frame = false;      // key columns (and counts) of a completed frame exist: the join's "previous frame"
    bool results_valid = false;   // the result arrays hold that frame (false after a failed attempt until the next success)
};

namespace {

thread_local std::string g_create_error;

#define CU_TRY(ctx, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                     \
            return SHAPES_E_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

#define NCCL_TRY(ctx, expr)                                                                       \
    do {                                                                                          \
        ncclResult_t r__ = (expr);                                                                \
        if (r__ != ncclSuccess) {                                                                 \
            (ctx)->err = std::string(#expr) + ": " + nccl_api().GetErrorString(r__);                     \
            return SHAPES_E_NCCL;                                                                 \
        }                                                                                         \
    } while (0)

template <typename T>
int dev_alloc(shapes_ctx *c, T **p, size_t count)
{
    void *q = nullptr;
    CU_TRY(c, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    c->allocs.push_back(q);
    *p = static_cast<T *>(q);
    return SHAPES_OK;
}

inline int grid_for(int64_t n, int threads, int cap)
{
    int64_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

int create_impl(shapes_ctx **out, int device_id, int rank, int world, const void *nccl_id,
                int64_t max_shapes, int64_t max_verts, int64_t max_pairs, int64_t max_contacts, bool need_nccl = true)
{
    if (!out || max_shapes < 0 || max_verts < 0 || max_pairs < 0 || max_contacts < 0 ||
        max_shapes > 0x7ffffff0ll || max_verts > 0x7ffffff0ll || max_pairs > 0x7ffffff0ll ||
        max_contacts > 0xfffffff0ll || world < 1 || world > SHAPES_MAX_RANKS || rank < 0 || rank >= world ||
        (world > 1 && need_nccl && !nccl_id)) {
        g_create_error = "shapes_create: bad argument";
        return SHAPES_E_ARG;
    }
    shapes_ctx *c = new shapes_ctx();
    c->device = device_id; c->rank = rank; c->world = world;
    c->max_shapes = max_shapes; c->max_verts = max_verts;
    c->max_pairs = max_pairs; c->max_contacts = max_contacts;
    c->use_graph = std::getenv("SHAPES_B200_NO_GRAPH") == nullptr;
    c->dbg_times = std::getenv("SHAPES_B200_KERNEL_TIMES") != nullptr;
    auto fail = [&](int code) { g_create_error = c->err; shapes_destroy(c); return code; };
#define TRY_CREATE(expr) do { int rc__ = (expr); if (rc__ != SHAPES_OK) return fail(rc__); } while (0)
    auto cu = [&](cudaError_t e, const char *what) {
        if (e == cudaSuccess) return SHAPES_OK;
        c->err = std::string(what) + ": " + cudaGetErrorString(e);
        return SHAPES_E_CUDA;
    };
    TRY_CREATE(cu(cudaSetDevice(device_id), "cudaSetDevice"));
    cudaDeviceProp prop;
    TRY_CREATE(cu(cudaGetDeviceProperties(&prop, device_id), "cudaGetDeviceProperties"));
    c->sm_count = prop.multiProcessorCount;
    TRY_CREATE(cu(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate"));
    TRY_CREATE(cu(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking), "cudaStreamCreate"));
    TRY_CREATE(cu(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming), "cudaEventCreate"));
    TRY_CREATE(cu(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming), "cudaEventCreate"));
    TRY_CREATE(cu(cudaEventCreate(&c->ev0), "cudaEventCreate"));
    TRY_CREATE(cu(cudaEventCreate(&c->ev1), "cudaEventCreate"));
    for (int k = 0; k <= SHAPES_N_STAGES; ++k) TRY_CREATE(cu(cudaEventCreate(&c->stage_ev[k]), "cudaEventCreate"));
    if (world > 1 && nccl_id) {
        ncclUniqueId id;
        static_assert(sizeof(ncclUniqueId) <= SHAPES_NCCL_ID_BYTES, "nccl id size");
        std::memcpy(&id, nccl_id, sizeof(id));
        NcclApi &api = nccl_api();
        if (!api.ok) { c->err = "NCCL unavailable: " + api.error; return fail(SHAPES_E_NCCL); }
        ncclResult_t r = api.CommInitRank(&c->comm, world, id, rank);
        if (r != ncclSuccess) { c->err = std::string("ncclCommInitRank: ") + api.GetErrorString(r); return fail(SHAPES_E_NCCL); }
    }
    const int64_t N = max_shapes, V = max_verts;
    c->chunk = (N + world - 1) / world;
    const int64_t Npad = std::max<int64_t>(c->chunk * world, 1);
    Params &P = c->P;
    TRY_CREATE(dev_alloc(c, &c->d_alive, N));
    TRY_CREATE(dev_alloc(c, &c->d_vert_offset, N + 1));
    TRY_CREATE(dev_alloc(c, &c->d_ext_min, V));
    TRY_CREATE(dev_alloc(c, &c->d_ext_max, V));
    TRY_CREATE(dev_alloc(c, &c->d_local, V));
    TRY_CREATE(dev_alloc(c, &c->d_ext_packed, N));
    for (int k = 0; k < 7; ++k) {
        TRY_CREATE(dev_alloc(c, &c->d_in[k], N));
        c->d_in2[0][k] = c->d_in[k];
        TRY_CREATE(dev_alloc(c, &c->d_in2[1][k], world > 1 ? N : 1));
    }
    TRY_CREATE(dev_alloc(c, &c->d_xf, N));
    P.xf = c->d_xf;
    TRY_CREATE(dev_alloc(c, &c->d_mass, N));
    P.mass = c->d_mass;
    TRY_CREATE(dev_alloc(c, &c->d_box2[0], Npad));
    TRY_CREATE(dev_alloc(c, &c->d_box2[1], world > 1 ? Npad : 1));
    P.box = c->d_box2[0];
    TRY_CREATE(dev_alloc(c, &c->d_gkeys2[0], Npad));
    TRY_CREATE(dev_alloc(c, &c->d_gkeys2[1], world > 1 ? Npad : 1));
    P.gkeys = c->d_gkeys2[0];
    P.chunk = (int)std::max<int64_t>(c->chunk, 1);
    TRY_CREATE(dev_alloc(c, &P.wv, V));
    TRY_CREATE(dev_alloc(c, &P.wn, V));
    TRY_CREATE(dev_alloc(c, &P.circ, N));
    c->use_sorted = std::getenv("SHAPES_B200_NO_SORTED") == nullptr;
    c->use_plan_ahead = std::getenv("SHAPES_B200_NO_PLAN_AHEAD") == nullptr;
    TRY_CREATE(dev_alloc(c, &c->d_hh, N));
    P.hh = c->d_hh;
    TRY_CREATE(dev_alloc(c, &P.w_i, (c->use_sorted || world > 1) ? max_pairs : 1));
    TRY_CREATE(dev_alloc(c, &c->d_w_j, (c->use_sorted || world > 1) ? max_pairs : 1));
    TRY_CREATE(dev_alloc(c, &P.w_a, (c->use_sorted || world > 1) ? max_pairs : 1));
    P.w_j = c->d_w_j;
    TRY_CREATE(dev_alloc(c, &c->d_radius, N));
    TRY_CREATE(dev_alloc(c, &P.keys, N));
    TRY_CREATE(dev_alloc(c, &P.keys_sorted, N));
    TRY_CREATE(dev_alloc(c, &P.rank, N));
    TRY_CREATE(dev_alloc(c, &P.sbox, N));
    TRY_CREATE(dev_alloc(c, &P.smeta, N));
    P.cell_cap = (unsigned)std::min<int64_t>(std::max<int64_t>(4 * N, 1 << 16), 1ll << 28);
    TRY_CREATE(dev_alloc(c, &P.cell_count, (size_t)P.cell_cap + 2));
    TRY_CREATE(dev_alloc(c, &P.cell_begin, (size_t)P.cell_cap + 2));
    TRY_CREATE(dev_alloc(c, &P.cell_mark, (size_t)P.cell_cap + 2));
    TRY_CREATE(cu(cudaMemset(P.cell_count, 0, ((size_t)P.cell_cap + 2) * sizeof(uint32_t)), "cudaMemset"));
    P.key_none = P.cell_cap;
    P.multi_rank = world > 1 ? 1 : 0;
    TRY_CREATE(dev_alloc(c, &P.big_idx, N));
    TRY_CREATE(dev_alloc(c, &c->d_bounds2[0], (size_t)4 * world));
    TRY_CREATE(dev_alloc(c, &c->d_bounds2[1], (size_t)4 * world));
    TRY_CREATE(dev_alloc(c, &c->d_flags, (size_t)2 * SHAPES_MAX_RANKS));
    TRY_CREATE(cu(cudaMemset(c->d_flags, 0, sizeof(unsigned long long) * 2 * SHAPES_MAX_RANKS), "cudaMemset"));
    P.rank_bounds = c->d_bounds2[0];
    P.flags = c->d_flags; P.n_peers = 0; P.my_rank = rank;
    c->use_p2p = std::getenv("SHAPES_B200_NO_P2P") == nullptr;
    TRY_CREATE(dev_alloc(c, &P.cnt, N));
    TRY_CREATE(dev_alloc(c, &P.off, N));
    TRY_CREATE(dev_alloc(c, &P.hitmask, N));
    TRY_CREATE(dev_alloc(c, &P.pair_i, max_pairs));
    TRY_CREATE(dev_alloc(c, &P.pair_j, max_pairs));
    TRY_CREATE(dev_alloc(c, &P.man, max_pairs));
    TRY_CREATE(dev_alloc(c, &P.ccnt, max_pairs));
    c->d_pair_i = P.pair_i; c->d_pair_j = P.pair_j; c->d_man = P.man; c->d_ccnt = P.ccnt;
    TRY_CREATE(dev_alloc(c, &P.coff, max_pairs));
    TRY_CREATE(dev_alloc(c, &P.row_map, max_contacts));
    TRY_CREATE(cu(cudaMemset(P.ccnt, 0, std::max<int64_t>(max_pairs, 1) * sizeof(uint32_t)), "cudaMemset"));
    const int64_t C = max_contacts;
    TRY_CREATE(dev_alloc(c, &P.key_i, C)); TRY_CREATE(dev_alloc(c, &P.key_j, C));
    TRY_CREATE(dev_alloc(c, &P.feat_a, C)); TRY_CREATE(dev_alloc(c, &P.feat_b, C));
    TRY_CREATE(dev_alloc(c, &P.flip, C));
    for (int q = 0; q < 4; ++q) TRY_CREATE(dev_alloc(c, &c->alt_key[q], C));
    TRY_CREATE(dev_alloc(c, &c->d_cache_np, C)); TRY_CREATE(dev_alloc(c, &c->d_cache_f, C));
    TRY_CREATE(dev_alloc(c, &P.warm_np, C)); TRY_CREATE(dev_alloc(c, &P.warm_f, C)); TRY_CREATE(dev_alloc(c, &P.warm_hit, C));
    double **cols[] = { &P.normal_x, &P.normal_y, &P.center_x, &P.center_y, &P.depth, &P.b_np,
                        &P.ra_x, &P.ra_y, &P.rb_x, &P.rb_y, &P.rn_x, &P.rn_y, &P.inv_eff_np, &P.inv_eff_f };
    for (double **col : cols) TRY_CREATE(dev_alloc(c, col, C));
    for (int q = 0; q < 6; ++q) { TRY_CREATE(dev_alloc(c, &P.j_np[q], C)); TRY_CREATE(dev_alloc(c, &P.j_f[q], C)); }
    TRY_CREATE(dev_alloc(c, &P.st, 1));
    TRY_CREATE(cu(cudaMemset(P.st, 0, sizeof(FrameState)), "cudaMemset"));
    c->use_rows = std::getenv("SHAPES_B200_NO_ROWS") == nullptr;
    if (const char *lay = std::getenv("SHAPES_B200_ROWS_LAYOUT")) c->rows_fold = std::strcmp(lay, "contiguous") != 0;   // pin the home layout
    if (world > 1 && c->use_rows && max_pairs < (1ll << 28)) {
        // the exchange arena of rows mode: one allocation, identical layout on every rank (the capacities are the
        // same everywhere), so a peer's buffer is its arena base + the local offset
        shapes_ctx::RowsLayout &L = c->rwl;
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t at = off; off += (bytes + 255) & ~size_t(255); return at; };
        L.gkeys[0] = take(sizeof(uint32_t) * Npad); L.gkeys[1] = take(sizeof(uint32_t) * Npad);
        L.box = take(sizeof(Box) * Npad); L.xf = take(sizeof(Xf) * Npad); L.mass = take(sizeof(double2) * Npad);
        for (int k = 0; k < 7; ++k) L.in[k] = take(sizeof(double) * Npad);
        L.cq = take(sizeof(uint32_t) * Npad); L.qoff = take(sizeof(uint32_t) * Npad);
        for (int q = 0; q < 2; ++q) { L.bounds[q] = take(sizeof(unsigned long long) * 4 * world); L.weights[q] = take(sizeof(uint32_t) * ROW_BINS * world); }
        L.counts = take(sizeof(long long) * 6 * world); L.err = take(sizeof(int) * 2 * world);
        L.flags = take(sizeof(unsigned long long) * RW_PHASES * SHAPES_MAX_RANKS);
        const size_t MP = (size_t)std::max<int64_t>(max_pairs, 1);
        L.prec = take(sizeof(PairRec) * MP); L.phdr = take(sizeof(PairHdr) * MP);
        // inbox: one section per home rank, each large enough for that rank's whole home (2 folded blocks <= chunk + 1 slots)
        L.inbox = take(sizeof(HomeRec) * (size_t)(c->chunk + 2) * (size_t)world); L.inbox_cnt = take(sizeof(unsigned) * SHAPES_MAX_RANKS);
        L.total = off;
        TRY_CREATE(dev_alloc(c, &c->rw_arena, L.total));
        TRY_CREATE(cu(cudaMemset(c->rw_arena, 0, L.total), "cudaMemset"));
        TRY_CREATE(dev_alloc(c, &c->d_roww, ROW_BINS));
        TRY_CREATE(dev_alloc(c, &c->d_ccnt_w, max_pairs));
        TRY_CREATE(dev_alloc(c, &c->d_chunk_sum, SCAN_BLOCKS));
        TRY_CREATE(dev_alloc(c, &c->d_mat_stamp, N));
        TRY_CREATE(dev_alloc(c, &c->d_kept_list, N));
        TRY_CREATE(cu(cudaMemset(c->d_mat_stamp, 0, sizeof(uint32_t) * (size_t)std::max<int64_t>(N, 1)), "cudaMemset"));
        c->peer_arena[rank] = c->rw_arena;
    }
    TRY_CREATE(dev_alloc(c, &c->d_counts, 2 * world));
    TRY_CREATE(dev_alloc(c, &c->d_n_prev, 1));
    TRY_CREATE(cu(cudaMallocHost(&c->h_state, sizeof(FrameState)), "cudaMallocHost"));
    TRY_CREATE(cu(cudaMallocHost(&c->h_counts, sizeof(int64_t) * 6 * world), "cudaMallocHost"));
    // library scratch: radix sort of (cell key, slot) and the offset scan
    size_t cb = 0;
    TRY_CREATE(cu(cub::DeviceScan::ExclusiveSum(nullptr, cb, P.cnt, P.off, (int)std::max<int64_t>(N, 1), c->stream),
                  "cub scan size"));
    size_t cb3 = 0;
    TRY_CREATE(cu(cub::DeviceScan::ExclusiveSum(nullptr, cb3, P.cell_count, P.cell_begin, (int)P.cell_cap + 1, c->stream),
                  "cub scan size"));
    cb = std::max(cb, cb3);
    size_t cb2 = 0;
    TRY_CREATE(cu(cub::DeviceScan::ExclusiveSum(nullptr, cb2, P.ccnt, P.coff, (int)std::max<int64_t>(max_pairs, 1), c->stream),
                  "cub scan size"));
    cb = std::max(cb, cb2);
    c->scan_tmp_bytes = cb;
    {   // grid-stride k_manifolds grid: the number of co-resident blocks
        int b4 = 0, b8 = 0;
        int bc = 0;
        TRY_CREATE(cu(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b4, k_manifolds<4, false>, CT_THREADS, 0), "occupancy"));
        TRY_CREATE(cu(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b8, k_manifolds<MAX_STAGED_VERTS, false>, CT_THREADS, 0), "occupancy"));
        TRY_CREATE(cu(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bc, k_manifolds<MAX_STAGED_VERTS, true>, CT_THREADS, 0), "occupancy"));
        c->ct_blocks[0] = std::max(b4, 1); c->ct_blocks[1] = std::max(b8, 1); c->ct_blocks[2] = std::max(bc, 1);
        int bco = 0;
        TRY_CREATE(cu(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bco, k_manifolds_coop<1>, CO_WARPS * 32, 0), "occupancy"));
        c->coop_blocks = std::max(bco, 1);
        TRY_CREATE(cu(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bco, k_manifolds_coop<2>, CO_WARPS * 32, 0), "occupancy"));
        c->coop_blocks_rows = std::max(bco, 1);
        c->use_coop = std::getenv("SHAPES_B200_NO_COOP") == nullptr;
        int br = 0;
        TRY_CREATE(cu(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&br, k_rows, 256, 0), "occupancy"));
        c->rows_blocks = std::max(br, 1);
    }
    TRY_CREATE(dev_alloc(c, reinterpret_cast<uint8_t **>(&c->d_scan_tmp), cb));
    P.max_pairs = max_pairs; P.max_contacts = max_contacts;
    P.alive = c->d_alive; P.vert_offset = c->d_vert_offset; P.local = c->d_local;
    P.ext_min = c->d_ext_min; P.ext_max = c->d_ext_max; P.ext_packed = c->d_ext_packed;
#undef TRY_CREATE
    *out = c;
    return SHAPES_OK;
}

// Issue one frame on the ctx stream.  `in` = device pointers (pos_x, pos_y, rot, cos, sin, inv_lin, inv_rot).
// rows mode: the two home blocks of a rank, [lo[0], hi[0]) = low block, [lo[1], hi[1]) = high block
inline void rows_home_blocks(const shapes_ctx *c, int rank, int64_t n_slots, int64_t lo[2], int64_t hi[2])
{
    if (!c->rows_fold) {     // contiguous homes: one block per rank
        lo[0] = std::min<int64_t>(rank * c->chunk, n_slots); hi[0] = std::min<int64_t>((rank + 1) * c->chunk, n_slots);
        lo[1] = hi[1] = n_slots;
        return;
    }
    const int64_t blk = std::max<int64_t>((c->chunk + 1) / 2, 1);
    lo[0] = std::min<int64_t>(rank * blk, n_slots); hi[0] = std::min<int64_t>((rank + 1) * blk, n_slots);
    lo[1] = std::min<int64_t>((2 * c->world - 1 - rank) * blk, n_slots); hi[1] = std::min<int64_t>((2 * c->world - rank) * blk, n_slots);
}

constexpr int SHAPES_I_REPLAN = 1;   // internal: the frame ran on a stale grid plan and must be issued again

// Enqueue one frame on the ctx stream (no host synchronisation).  `in` = device pointers (pos_x, pos_y, rot, cos,
// sin, inv_lin, inv_rot).  frame_finish() waits for it and does the bookkeeping.
int frame_launch(shapes_ctx *c, int64_t n_slots, const double *const in[7], double dt, double baumgarte,
                 double slop, bool want_world, bool own_slots_only)
{
    if (!c->hulls_set) { c->err = "shapes_frame: shapes_set_hulls has not been called"; return SHAPES_E_ARG; }
    if (n_slots != c->n_slots) { c->err = "shapes_frame: n_slots differs from shapes_set_hulls"; return SHAPES_E_ARG; }
    if (!in[0] || !in[1] || !in[5] || !in[6] || (!in[2] && !(in[3] && in[4])) || ((in[3] == nullptr) != (in[4] == nullptr))) {
        c->err = "shapes_frame: missing input column";
        return SHAPES_E_ARG;
    }
    CU_TRY(c, cudaSetDevice(c->device));
    Params &P = c->P;
    const int N = (int)n_slots;
    P.n_slots = N;
    P.own_lo = (int)std::min<int64_t>(c->rank * c->chunk, N);
    P.own_hi = (int)std::min<int64_t>((c->rank + 1) * c->chunk, N);
    const int n_query = P.own_hi - P.own_lo;
    P.pos_x = in[0]; P.pos_y = in[1]; P.rot = in[2]; P.cos_rot = in[3]; P.sin_rot = in[4];
    P.inv_lin = in[5]; P.inv_rot = in[6];
    P.dt = dt; P.baumgarte = baumgarte; P.slop = slop;
    P.cell_size = c->user_cell > 0.0 ? c->user_cell : c->auto_cell;
    // general polygon worlds on one rank: hull records and the SAT work list in cell order
    P.sorted_mode = (c->use_sorted && c->use_coop && c->world == 1 && !c->has_circles && c->max_hull_verts > 4) ? 1 : 0;
    P.plan_ahead = (c->use_plan_ahead && c->world == 1) ? 1 : 0;
    const bool seed_plan = P.plan_ahead && !c->plan_valid;
    // a freshly seeded plan is exact: whatever is on the big list then belongs there
    P.big_limit = seed_plan ? 0xffffffffu : (unsigned)std::max<int64_t>(std::max<int64_t>(1024, n_slots / 256), 2 * c->big_seen + 64);
    // the previous frame's key columns become the join's "that" side; this frame writes the other set
    if (c->have_frame) {
        std::swap(P.key_i, c->alt_key[0]); std::swap(P.key_j, c->alt_key[1]);
        std::swap(P.feat_a, c->alt_key[2]); std::swap(P.feat_b, c->alt_key[3]);
        c->parity ^= 1;
        c->n_prev_keys = c->last_contacts;
    } else { c->n_prev_keys = 0; c->cache_valid = false; }
    const bool warm = c->cache_valid && c->n_prev_keys >= 0;
    P.pk_i = c->alt_key[0]; P.pk_j = c->alt_key[1]; P.pk_fa = c->alt_key[2]; P.pk_fb = c->alt_key[3];
    P.cache_np = c->d_cache_np; P.cache_f = c->d_cache_f;
    const long long n_prev_now = warm ? c->n_prev_keys : 0;
    P.n_prev = c->d_n_prev;
    // Cell budget of this frame: the scan and the clear cover exactly this many cells, so it follows
    // the grid the last frame needed (x2 head-room) instead of the table capacity; the planner coarsens
    // the cells if the world outgrows it within one frame (results do not depend on the cell size).
    if (c->cell_limit == 0) c->cell_limit = P.cell_cap;
    else if (c->have_frame) {
        const unsigned used = c->h_state->n_cells;
        if (used > c->cell_limit / 2 + c->cell_limit / 4 || (unsigned long long)used * 8 < c->cell_limit)
            c->cell_limit = (unsigned)std::min<unsigned long long>(P.cell_cap, std::max<unsigned long long>(65536ull, 2ull * used));
    }
    P.cell_limit = c->cell_limit;
    ++c->frame_no;
    P.frame_no = c->frame_no;
    const int fpar = (int)(c->frame_no & 1);
    const bool rows = c->world > 1 && c->rows_ready && c->use_rows;
    const bool p2p = rows || (c->world > 1 && c->peers_ready && c->use_p2p);
    if (c->world > 1) { P.box = c->d_box2[fpar]; P.rank_bounds = c->d_bounds2[fpar]; P.gkeys = c->d_gkeys2[fpar]; }
    P.n_peers = p2p ? c->world : 0;
    P.remote_inputs = (p2p && own_slots_only) ? 1 : 0;
    for (int r = 0; r < c->world && p2p && !rows; ++r) {
        P.peer_box[r] = c->peer_box2[fpar][r]; P.peer_bounds[r] = c->peer_bounds2[fpar][r]; P.peer_flags[r] = c->peer_flags[r];
        P.peer_keys[r] = c->peer_keys2[fpar][r];
        for (int k = 0; k < 7; ++k) P.peer_in[k][r] = c->peer_in2[fpar][k][r];
    }
    P.work_mode = rows ? 2 : (P.sorted_mode ? 1 : 0);
    P.pair_i = c->d_pair_i; P.pair_j = c->d_pair_j; P.man = c->d_man; P.ccnt = c->d_ccnt;
    P.sat_ccnt = P.ccnt; P.xf = c->d_xf; P.mass = c->d_mass;
    const bool seed_rows = rows && !c->plan_valid;
    int n_home = 0;
    if (rows) {
        // every exchanged buffer lives in the ranks' arenas: peer pointer = that rank's arena base + my offset
        const shapes_ctx::RowsLayout &L = c->rwl;
        char *mine = c->rw_arena;
        P.sorted_mode = 0; P.plan_ahead = 0;
        P.big_limit = seed_rows ? 0xffffffffu : (unsigned)std::max<int64_t>(std::max<int64_t>(1024, n_slots / 256), 2 * c->big_seen + 64);
        P.box = reinterpret_cast<Box *>(mine + L.box); P.gkeys = reinterpret_cast<uint32_t *>(mine + L.gkeys[fpar]);
        P.xf = reinterpret_cast<Xf *>(mine + L.xf); P.mass = reinterpret_cast<double2 *>(mine + L.mass);
        // homes: blocks g and 2G-1-g of the slot space
        int64_t hlo[2], hhi[2];
        rows_home_blocks(c, c->rank, N, hlo, hhi);
        P.rw_fold = c->rows_fold ? 1 : 0;
        P.rw_blk = (int)std::max<int64_t>(c->rows_fold ? (c->chunk + 1) / 2 : c->chunk, 1);
        P.rw_lo_lo = (int)hlo[0]; P.rw_lo_hi = (int)hhi[0]; P.rw_hi_lo = (int)hlo[1]; P.rw_hi_hi = (int)hhi[1];
        P.own_lo = P.rw_lo_lo; P.own_hi = P.rw_lo_hi;
        n_home = (P.rw_lo_hi - P.rw_lo_lo) + (P.rw_hi_hi - P.rw_hi_lo);
        P.flags = reinterpret_cast<unsigned long long *>(mine + L.flags);
        // the result arrays of a home live in its arena: the sweeping ranks store into them
        P.prec = reinterpret_cast<PairRec *>(mine + L.prec); P.phdr = reinterpret_cast<PairHdr *>(mine + L.phdr); P.q_off = reinterpret_cast<uint32_t *>(mine + L.qoff);
        P.sat_ccnt = c->d_ccnt_w;
        P.roww = c->d_roww; P.mat_stamp = c->d_mat_stamp; P.kept_list = c->d_kept_list;
        P.inbox = reinterpret_cast<const HomeRec *>(mine + L.inbox); P.inbox_cnt = reinterpret_cast<const unsigned *>(mine + L.inbox_cnt);
        P.inbox_cap = (int)(c->chunk + 2);
        P.dbg_local_stores = std::getenv("SHAPES_B200_DBG_LOCAL_STORES") ? 1 : 0;
        P.rw_weights_prev = reinterpret_cast<const uint32_t *>(mine + L.weights[fpar ^ 1]);
        P.rw_bounds_prev = reinterpret_cast<const unsigned long long *>(mine + L.bounds[fpar ^ 1]);
        for (int r = 0; r < c->world; ++r) {
            char *a = c->peer_arena[r];
            P.peer_box[r] = reinterpret_cast<Box *>(a + L.box); P.peer_keys[r] = reinterpret_cast<uint32_t *>(a + L.gkeys[fpar]);
            P.rw_xf[r] = reinterpret_cast<Xf *>(a + L.xf); P.rw_mass[r] = reinterpret_cast<double2 *>(a + L.mass);
            P.rw_cq[r] = reinterpret_cast<uint32_t *>(a + L.cq); P.rw_qoff[r] = reinterpret_cast<uint32_t *>(a + L.qoff);
            P.rw_prec[r] = reinterpret_cast<PairRec *>(a + L.prec); P.rw_phdr[r] = reinterpret_cast<PairHdr *>(a + L.phdr);
            P.rw_inbox[r] = reinterpret_cast<HomeRec *>(a + L.inbox); P.rw_inbox_cnt[r] = reinterpret_cast<unsigned *>(a + L.inbox_cnt);
            P.peer_bounds[r] = reinterpret_cast<unsigned long long *>(a + L.bounds[fpar]);
            P.rw_weights[r] = reinterpret_cast<uint32_t *>(a + L.weights[fpar]);
            P.rw_counts[r] = reinterpret_cast<long long *>(a + L.counts); P.rw_err[r] = reinterpret_cast<int *>(a + L.err);
            P.peer_flags[r] = reinterpret_cast<unsigned long long *>(a + L.flags);

        }
    } else P.flags = c->d_flags;
    P.world_x = want_world ? c->d_world_x : nullptr;
    P.world_y = want_world ? c->d_world_y : nullptr;
    cudaStream_t s = c->stream;
    const int sms = c->sm_count;

    // The launch sequence of a frame is fixed for given buffers and scalars (all counts live in
    // device memory), so it is captured once into a CUDA graph and replayed: one cudaGraphLaunch
    // instead of ~25 launches per frame.
    // rows mode (multi-rank, peers mapped): see the k_rw_* kernels
    auto issue_rows = [&](bool advance) -> int {
        int stage = 0;
        c->dbg_marks = 0;
        auto kt = [&](const char *name) {      // per-kernel timing marks (debug aid, profiled frames only)
            if (!c->dbg_times || !c->profiling) return;
            if ((int)c->dbg_ev.size() <= c->dbg_marks) { cudaEvent_t e; cudaEventCreate(&e); c->dbg_ev.push_back(e); c->dbg_name.push_back(name); c->dbg_sum.push_back(0.0); }
            c->dbg_name[c->dbg_marks] = name;
            cudaEventRecord(c->dbg_ev[c->dbg_marks++], s);
        };
        kt("start");
    #define STAGE_MARK() do { if (c->profiling) CU_TRY(c, cudaEventRecord(c->stage_ev[stage], s)); ++stage; } while (0)
        const int n_query = n_home;     // my slice: both home blocks
        const int gq = grid_for(n_query, 256, sms * 8);
        const int gk = grid_for(2 * c->chunk + 1024, 256, sms * 8);     // the shapes a rank keeps: about a home's worth plus halo and big list
        STAGE_MARK(); // 0: transform (home slots) + record push
        k_rw_begin<<<1, 1024, 0, s>>>(P, advance ? 1 : 0); ++c->launches;
        kt("k_rw_begin");
        CU_TRY(c, cudaMemsetAsync(P.cell_count, 0, sizeof(uint32_t) * ((size_t)P.cell_limit + 2), s));
        kt("memset");
        if (n_query > 0) { k_rw_transform<false><<<gq, 256, 0, s>>>(P, n_query); ++c->launches; }
        kt("k_rw_transform");
        STAGE_MARK(); // 1: barrier KEYS (this frame's bounds ride along, for the next frame's grid)
        k_rw_sync<<<1, 1024, 0, s>>>(P, RW_PHASE_KEYS); ++c->launches;
        kt("k_rw_sync:KEYS");
        STAGE_MARK(); // 2: keep my rows' keys
        if (N > 0) { k_rw_bin<<<gk, 256, 0, s>>>(P); ++c->launches; }
        kt("k_rw_bin");
        STAGE_MARK(); // 3: cell offsets
        k_scan_cells_sums<<<SCAN_BLOCKS, SCAN_THREADS, 0, s>>>(P, c->d_chunk_sum); ++c->launches;
        kt("k_scan_cells_sums");
        k_scan_cells_apply<<<SCAN_BLOCKS, SCAN_THREADS, 0, s>>>(P, c->d_chunk_sum); ++c->launches;
        kt("k_scan_cells_apply");
        STAGE_MARK(); // 4: scatter into cell order (AABBs come with the records); the hull pass of the kept shapes is forked off
        if (N > 0) {
            k_scatter_sorted<<<gk, 256, 0, s>>>(P); ++c->launches;
            kt("k_scatter_sorted");
            // the hull pass only feeds the SAT stage: fork it onto the side stream (inside a captured graph this is a
            // parallel branch), join before the manifolds
            CU_TRY(c, cudaEventRecord(c->ev_fork, s));
            CU_TRY(c, cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
            k_rw_hulls<<<sms * 8, 256, 0, c->side_stream>>>(P); ++c->launches;
            CU_TRY(c, cudaEventRecord(c->ev_join, c->side_stream));
        }
        STAGE_MARK(); // 5: single-pass sweep of my rows; every query's count is pushed to its home; barrier CNT
        if (N > 0) {
            k_sweep<SWEEP_FUSED><<<grid_for(n_query + n_query / 2 + 4096, 128, 1 << 30), 128, 0, s>>>(P); ++c->launches;
            kt("k_sweep");
            k_big<false><<<64, 256, 0, s>>>(P); ++c->launches;
            kt("k_big");
            k_big<true><<<64, 256, 0, s>>>(P); ++c->launches;
            kt("k_big");
        }
        k_rw_sync<<<1, 1024, 0, s>>>(P, RW_PHASE_CNT); ++c->launches;
        kt("k_rw_sync:CNT");
        STAGE_MARK(); // 6: home -- offsets of my slice, sent back to the sweeping ranks; barrier OFF
        if (n_query > 0) {
            k_rw_home_counts<<<gq, 256, 0, s>>>(P, n_query); ++c->launches;
            kt("k_rw_home_counts");
            size_t cb = c->scan_tmp_bytes;
            CU_TRY(c, cub::DeviceScan::ExclusiveSum(c->d_scan_tmp, cb, P.cnt, P.off, n_query, s));
            kt("cub_scan");
        }
        k_finish_pairs<<<1, 1, 0, s>>>(P, n_query); ++c->launches;
        kt("k_finish_pairs");
        if (n_query > 0) { k_rw_push_offsets<<<gq, 256, 0, s>>>(P, n_query); ++c->launches; }
        kt("k_rw_push_offsets");
        k_rw_sync<<<1, 1024, 0, s>>>(P, RW_PHASE_OFF); ++c->launches;
        kt("k_rw_sync:OFF");
        STAGE_MARK(); // 7
        STAGE_MARK(); // 8: manifolds over my work list, every pair stored into its final place at its home; barrier RESULTS
        if (N > 0) {
            CU_TRY(c, cudaStreamWaitEvent(s, c->ev_join, 0));      // world vertices / normals of the kept hulls are in place
            kt("join:k_rw_hulls");
            if (c->has_circles) k_manifolds<MAX_STAGED_VERTS, true><<<sms * c->ct_blocks[2], CT_THREADS, 0, s>>>(P);
            else if (c->max_hull_verts <= 4) k_manifolds<4, false><<<sms * c->ct_blocks[0], CT_THREADS, 0, s>>>(P);
            else if (c->use_coop) {
                k_manifolds_coop<2><<<sms * c->coop_blocks_rows, CO_WARPS * 32, 0, s>>>(P);
                kt("k_manifolds_coop");
                {   // hulls with more than 8 vertices, partners of big queries this rank does not keep
                    k_manifolds<MAX_STAGED_VERTS, false, true><<<sms * c->ct_blocks[1], CT_THREADS, 0, s>>>(P); ++c->launches;
                }
            }
            else k_manifolds<MAX_STAGED_VERTS, false><<<sms * c->ct_blocks[1], CT_THREADS, 0, s>>>(P);
            kt("k_manifolds");
            ++c->launches;
        }
        k_rw_sync<<<1, 1024, 0, s>>>(P, RW_PHASE_RESULTS); ++c->launches;
        kt("k_rw_sync:RESULTS");
        STAGE_MARK(); // 9: home -- pair columns and counts out of the delivered records, row offsets
        if (c->max_pairs > 0) {
            k_rw_unpack<<<sms * 8, 256, 0, s>>>(P); ++c->launches;
            kt("k_rw_unpack");
            size_t cb = c->scan_tmp_bytes;
            CU_TRY(c, cub::DeviceScan::ExclusiveSum(c->d_scan_tmp, cb, P.ccnt, P.coff, (int)c->max_pairs, s));
            kt("cub_scan");
            k_row_map<<<sms * 8, 256, 0, s>>>(P); ++c->launches;
            kt("k_row_map");
        }
        STAGE_MARK(); // 10: contact rows
        if (N > 0) { k_rows<<<sms * c->rows_blocks, 256, 0, s>>>(P); ++c->launches; }
        kt("k_rows");
        STAGE_MARK(); // 11: warm-start cache join
        if (N > 0 && warm) { k_warm_join<<<sms * 8, 256, 0, s>>>(P); ++c->launches; }
        kt("k_warm_join");
        STAGE_MARK(); // end
    #undef STAGE_MARK
        // barrier COUNTS: every rank learns every rank's pair / contact totals (global row offsets of the slices)
        k_rw_sync<<<1, 1024, 0, s>>>(P, RW_PHASE_COUNTS); ++c->launches;
        kt("k_rw_sync:COUNTS");
        CU_TRY(c, cudaGetLastError());
        CU_TRY(c, cudaMemcpyAsync(c->h_counts, c->rw_arena + c->rwl.counts, sizeof(int64_t) * 6 * c->world, cudaMemcpyDeviceToHost, s));
        CU_TRY(c, cudaMemcpyAsync(c->h_state, P.st, sizeof(FrameState), cudaMemcpyDeviceToHost, s));
        return SHAPES_OK;
    };
    auto issue = [&]() -> int {
        if (rows) return issue_rows(true);
        int stage = 0;
    #define STAGE_MARK() do { if (c->profiling) CU_TRY(c, cudaEventRecord(c->stage_ev[stage], s)); ++stage; } while (0)
        STAGE_MARK(); // 0: transform
        if (P.plan_ahead) {
            k_begin_frame<<<1, 1, 0, s>>>(P); ++c->launches;
            CU_TRY(c, cudaMemsetAsync(P.cell_count, 0, sizeof(uint32_t) * ((size_t)P.cell_limit + 2), s));
        } else { k_reset_state<<<1, 1, 0, s>>>(P.st); ++c->launches; }
        if (N > 0) {
            CU_TRY(c, cudaMemsetAsync(P.cnt, 0, sizeof(unsigned long long) * (size_t)std::max(n_query, 1), s));
            k_transform_aabb<false><<<grid_for(n_query, 256, sms * 8), 256, 0, s>>>(P, P.own_lo, P.own_hi); ++c->launches;
        }
        STAGE_MARK(); // 1: allgather
        if (N > 0 && c->world > 1 && p2p) {
            // exchange #1, peer-to-peer, barrier A: every rank's own AABB records and bounds are in place
            k_publish_peers<<<1, SHAPES_MAX_RANKS, 0, s>>>(P, 0); ++c->launches;
            k_wait_peers<<<1, SHAPES_MAX_RANKS, 0, s>>>(P, 0); ++c->launches;
        } else if (N > 0 && c->world > 1) {
            // exchange #1 through NCCL: AABB records of every rank's slot range (in place), plus each
            // rank's finite bounds (32 B per rank) so that nobody re-reduces all N boxes
            k_publish_bounds<<<1, 1, 0, s>>>(P, c->rank); ++c->launches;
            NCCL_TRY(c, nccl_api().GroupStart());
            NCCL_TRY(c, nccl_api().AllGather(reinterpret_cast<const char *>(P.box) + sizeof(Box) * c->chunk * c->rank, P.box,
                                             sizeof(Box) * c->chunk, ncclChar, c->comm, s));
            NCCL_TRY(c, nccl_api().AllGather(P.rank_bounds + 4 * c->rank, P.rank_bounds, 4, ncclUint64, c->comm, s));
            NCCL_TRY(c, nccl_api().GroupEnd());
        }
        STAGE_MARK(); // 2: grid keys (plan-ahead frames: done by K0)
        if (N > 0 && !P.plan_ahead) {
            k_plan_grid<<<1, 1, 0, s>>>(P, c->world); ++c->launches;
            k_clear_cells<<<sms * 4, 256, 0, s>>>(P); ++c->launches;
            if (p2p) {
                // keys of the OWN slots, pushed (4 B each) to every rank; barrier B; then every rank bins all keys
                k_keys<<<grid_for(n_query, 256, sms * 8), 256, 0, s>>>(P, P.own_lo, P.own_hi); ++c->launches;
                k_publish_peers<<<1, SHAPES_MAX_RANKS, 0, s>>>(P, 1); ++c->launches;
                k_wait_peers<<<1, SHAPES_MAX_RANKS, 0, s>>>(P, 1); ++c->launches;
            } else { k_keys<<<grid_for(N, 256, sms * 8), 256, 0, s>>>(P, 0, N); ++c->launches; }
            k_bin<<<grid_for(N, 256, sms * 8), 256, 0, s>>>(P); ++c->launches;
        }
        STAGE_MARK(); // 3: cell offsets (exclusive scan of the histogram)
        if (N > 0) {
            size_t cb = c->scan_tmp_bytes;
            CU_TRY(c, cub::DeviceScan::ExclusiveSum(c->d_scan_tmp, cb, P.cell_count, P.cell_begin, (int)P.cell_limit + 1, s));
        }
        STAGE_MARK(); // 4: scatter into cell order
        if (N > 0) { k_scatter_sorted<<<grid_for(N, 256, sms * 8), 256, 0, s>>>(P); ++c->launches; }
        STAGE_MARK(); // 5: sweep count (sorted mode: the whole single-pass sweep)
        if (N > 0) {
            const int g = grid_for(n_query + n_query / 2 + 4096, 128, 1 << 30);
            if (P.sorted_mode) k_sweep<SWEEP_FUSED><<<g, 128, 0, s>>>(P);
            else k_sweep<SWEEP_COUNT><<<g, 128, 0, s>>>(P);
            ++c->launches;
            k_big<false><<<64, 256, 0, s>>>(P); ++c->launches;
        }
        STAGE_MARK(); // 6: scan
        if (N > 0) {
            if (n_query > 0) {
                size_t cb = c->scan_tmp_bytes;
                CU_TRY(c, cub::DeviceScan::ExclusiveSum(c->d_scan_tmp, cb, P.cnt, P.off, n_query, s));
            }
            k_finish_pairs<<<1, 1, 0, s>>>(P, n_query); ++c->launches;
        }
        STAGE_MARK(); // 7: sweep emit
        if (N > 0) {
            if (!P.sorted_mode) { k_sweep<SWEEP_EMIT><<<grid_for(n_query + n_query / 2 + 4096, 128, 1 << 30), 128, 0, s>>>(P); ++c->launches; }
            k_big<true><<<64, 256, 0, s>>>(P); ++c->launches;
        }
        STAGE_MARK(); // 8: manifolds (SAT + clipping)
        if (N > 0) {
            if (c->has_circles) k_manifolds<MAX_STAGED_VERTS, true><<<sms * c->ct_blocks[2], CT_THREADS, 0, s>>>(P);
            else if (c->max_hull_verts <= 4) k_manifolds<4, false><<<sms * c->ct_blocks[0], CT_THREADS, 0, s>>>(P);
            else if (c->use_coop) {
                if (P.sorted_mode) k_manifolds_coop<1><<<sms * c->coop_blocks, CO_WARPS * 32, 0, s>>>(P);
                else k_manifolds_coop<0><<<sms * c->coop_blocks, CO_WARPS * 32, 0, s>>>(P);
                // hulls with more than 8 vertices / foreign hulls (multi-rank): per-thread pass over the flagged pairs
                if (c->max_hull_verts > MAX_STAGED_VERTS || c->world > 1) {
                    k_manifolds<MAX_STAGED_VERTS, false, true><<<sms * c->ct_blocks[1], CT_THREADS, 0, s>>>(P); ++c->launches;
                }
            }
            else k_manifolds<MAX_STAGED_VERTS, false><<<sms * c->ct_blocks[1], CT_THREADS, 0, s>>>(P);
            ++c->launches;
        }
        STAGE_MARK(); // 9: contact row offsets
        if (N > 0 && c->max_pairs > 0) {
            size_t cb = c->scan_tmp_bytes;
            CU_TRY(c, cub::DeviceScan::ExclusiveSum(c->d_scan_tmp, cb, P.ccnt, P.coff, (int)c->max_pairs, s));
            k_row_map<<<sms * 8, 256, 0, s>>>(P); ++c->launches;
        }
        STAGE_MARK(); // 10: contact rows (flatten + constraint generators)
        if (N > 0) { k_rows<<<sms * c->rows_blocks, 256, 0, s>>>(P); ++c->launches; }
        STAGE_MARK(); // 11: warm-start cache join
        if (N > 0 && warm) { k_warm_join<<<sms * 8, 256, 0, s>>>(P); ++c->launches; }
        STAGE_MARK(); // end
    #undef STAGE_MARK
        CU_TRY(c, cudaGetLastError());
        if (c->world > 1) {
            // exchange #2 (counts): every rank learns every rank's pair / contact counts, so the
            // global row offset of each rank's slice is known everywhere.
            NCCL_TRY(c, nccl_api().AllGather(&P.st->n_pairs, c->d_counts, 2, ncclInt64, c->comm, s));
            CU_TRY(c, cudaMemcpyAsync(c->h_counts, c->d_counts, sizeof(int64_t) * 2 * c->world, cudaMemcpyDeviceToHost, s));
        }
        CU_TRY(c, cudaMemcpyAsync(c->h_state, P.st, sizeof(FrameState), cudaMemcpyDeviceToHost, s));
        return SHAPES_OK;
    };
    FrameKey key;
    std::memset(&key, 0, sizeof(key));
    key.n = n_slots; for (int k = 0; k < 7; ++k) key.in[k] = in[k];
    key.dt = dt; key.baumgarte = baumgarte; key.slop = slop; key.cell = P.cell_size;
    key.world = want_world; key.profiling = c->profiling; key.geometry = c->geometry_version;
    key.warm = warm; key.n_prev = 0; key.p2p = p2p; key.cell_limit = P.cell_limit; key.remote = P.remote_inputs != 0; key.sorted = P.sorted_mode != 0;
    key.big_limit = P.big_limit;
    CU_TRY(c, cudaEventRecord(c->ev0, s));
    if (warm) { k_set_i64<<<1, 1, 0, s>>>(c->d_n_prev, n_prev_now); ++c->launches; }   // outside the graph: varies per frame
    if (seed_plan) {
        // first frame of a geometry (or after a failed frame): a bounds-only pass seeds the grid plan
        k_reset_state<<<1, 1, 0, s>>>(P.st); ++c->launches;
        if (N > 0) { k_transform_aabb<true><<<grid_for(N, 256, sms * 8), 256, 0, s>>>(P, 0, N); ++c->launches; }
    }
    c->pending_warm = warm; c->pending_seed = seed_plan || seed_rows; c->pending_rows = rows; c->pending_plan_ahead = P.plan_ahead != 0;
    const int64_t launches_before = c->launches;
    if (seed_rows) {
        // rows mode, first frame of a geometry: every rank reduces the bounds of its home slots and pushes them into
        // the inbox the frame's k_rw_begin reads (the "previous frame" one); barrier SEED; then the frame itself,
        // without advancing the frame counter again.  Not replayed from a graph.
        Params Ps = P;
        for (int r = 0; r < c->world; ++r)
            Ps.peer_bounds[r] = reinterpret_cast<unsigned long long *>(c->peer_arena[r] + c->rwl.bounds[fpar ^ 1]);
        k_rw_begin<<<1, 1024, 0, s>>>(P, 1); ++c->launches;      // advances the counter, resets the bounds accumulators
        if (n_home > 0) { k_rw_transform<true><<<grid_for(n_home, 256, sms * 8), 256, 0, s>>>(P, n_home); ++c->launches; }
        k_rw_publish<<<1, 1024, 0, s>>>(Ps, RW_PHASE_SEED); ++c->launches;
        k_rw_wait<<<1, 32, 0, s>>>(P, RW_PHASE_SEED); ++c->launches;
        const int rc = issue_rows(false);
        if (rc != SHAPES_OK) return rc;
    } else if (!c->use_graph || c->profiling || (c->world > 1 && !rows)) {
        // per-stage events cannot be timed from inside a graph; the r1 multi-rank exchange carries the frame number
        // in its kernel arguments (rows mode keeps it in device memory and replays)
        const int rc = issue();
        if (rc != SHAPES_OK) return rc;
    } else {
        const int gi = c->parity + (rows ? 2 * fpar : 0);
        cudaGraphExec_t &gexec = c->graph_exec[gi];
        if (!gexec || std::memcmp(&key, &c->graph_key[gi], sizeof(key)) != 0) {
            if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
            cudaGraph_t g = nullptr;
            CU_TRY(c, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            const int rc = issue();
            cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (rc != SHAPES_OK) { if (g) cudaGraphDestroy(g); return rc; }
            CU_TRY(c, ce);
            CU_TRY(c, cudaGraphInstantiate(&gexec, g, 0));
            cudaGraphDestroy(g);
            c->graph_key[gi] = key;
            c->graph_launches = c->launches - launches_before;
        } else c->launches += c->graph_launches;
        CU_TRY(c, cudaGraphLaunch(gexec, s));
    }
    CU_TRY(c, cudaEventRecord(c->ev1, s));
    return SHAPES_OK;
}

// Wait for the frame frame_launch() enqueued and do the bookkeeping.  SHAPES_I_REPLAN: the grid plan was stale,
// everything of the attempt has been undone and the caller must launch the frame again (every rank of a multi-rank
// job takes the same decision: the condition only depends on data all of them hold).
int frame_finish(shapes_ctx *c, shapes_frame_out *out)
{
    CU_TRY(c, cudaSetDevice(c->device));
    Params &P = c->P;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    const FrameState &st = *c->h_state;
    const bool warm = c->pending_warm;
    if (st.error & ERR_REPLAN) {
        // the plan was stale (the world moved further than the grid's margin since the last frame): seed it from
        // this frame's positions and run the frame again -- nothing of the aborted attempt is visible
        c->plan_valid = false;
        if (c->have_frame) {   // undo the key-column swap of this attempt
            std::swap(P.key_i, c->alt_key[0]); std::swap(P.key_j, c->alt_key[1]);
            std::swap(P.feat_a, c->alt_key[2]); std::swap(P.feat_b, c->alt_key[3]);
            c->parity ^= 1;
        }
        c->cache_valid = warm;
        return SHAPES_I_REPLAN;
    }
    c->plan_valid = (c->pending_plan_ahead || c->pending_rows) && st.error == 0;
    if (c->pending_rows && st.error == 0) {
        // Home layout for the next frames, from the locality counters every rank received with the counts: if most
        // pairs would stay on the GPU that sweeps them with contiguous homes (slot numbering follows the geometry) use
        // those; otherwise the folded blocks, which balance slots and pairs for any numbering.  Hysteresis 0.45 / 0.55.
        ++c->rows_frames;
        long long lf = 0, lc = 0, tot = 0;
        for (int r = 0; r < c->world; ++r) { lf += c->h_counts[4 * c->world + 2 * r]; lc += c->h_counts[4 * c->world + 2 * r + 1]; }
        for (int r = 0; r < c->world; ++r) tot += c->h_counts[4 * r] + c->h_counts[4 * r + 1];
        (void)lf;
        const double f_contig = tot > 0 ? 16.0 * (double)lc / (double)tot : 0.0;   // one query in 16 is sampled
        const bool want_fold = c->rows_fold ? !(f_contig > 0.55) : (f_contig < 0.45);
        if (want_fold != c->rows_fold && std::getenv("SHAPES_B200_ROWS_LAYOUT") == nullptr) {
            c->rows_fold = want_fold;
            for (int q = 0; q < 4; ++q) if (c->graph_exec[q]) { cudaGraphExecDestroy(c->graph_exec[q]); c->graph_exec[q] = nullptr; }
        }
    }
    if (c->pending_seed) c->big_seen = st.n_big;
    const int64_t now_pairs = st.n_pairs;
    const int64_t now_contacts = (st.error & ERR_PAIR_CAP) ? 2 * st.n_pairs : st.n_contacts;
    const bool capacity = st.error != 0 && !(st.error & ERR_PEER_TIMEOUT);
    if (capacity) {
        // SHAPES_E_CAPACITY leaves the ctx as it was before the call: the previous frame's key columns go back to the
        // "current" side, its Lagrangian cache stays valid, so after shapes_grow the retried frame joins against the
        // same EngineCache the failed attempt saw (Solvers/Contact.hs:84-121).  Only the result arrays are stale.
        if (c->have_frame) {
            std::swap(P.key_i, c->alt_key[0]); std::swap(P.key_j, c->alt_key[1]);
            std::swap(P.feat_a, c->alt_key[2]); std::swap(P.feat_b, c->alt_key[3]);
            c->parity ^= 1;
        }
        c->cache_valid = warm;
        c->results_valid = false;
    } else {
        c->last_pairs = now_pairs;
        c->last_contacts = now_contacts;
        c->have_frame = (st.error == 0);
        c->results_valid = c->have_frame;
        c->warm_done = warm && c->have_frame;
        c->cache_valid = false; // a cache describes exactly one previous frame
        if (c->world == 1) { c->h_counts[0] = c->last_pairs; c->h_counts[1] = c->last_contacts; }
    }
    if (out) {
        out->n_pairs = now_pairs;
        out->n_contacts = now_contacts;
        out->n_big = st.n_big;
        out->grid_w = st.W; out->grid_h = st.H; out->cell_size = st.h;
        float ms = 0.f;
        CU_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        out->device_ms = ms;
        out->total_ms = ms;
    }
    if (c->profiling)
        for (int k = 0; k < SHAPES_N_STAGES; ++k) CU_TRY(c, cudaEventElapsedTime(&c->stage_ms[k], c->stage_ev[k], c->stage_ev[k + 1]));
    if (c->profiling && c->dbg_times && c->pending_rows && c->dbg_marks > 1) {
        for (int k = 1; k < c->dbg_marks; ++k) { float ms = 0.f; cudaEventElapsedTime(&ms, c->dbg_ev[k - 1], c->dbg_ev[k]); c->dbg_sum[k] += ms; }
        ++c->dbg_frames; c->dbg_marks_seen = c->dbg_marks;
    }
    if (st.error & ERR_PEER_TIMEOUT) { c->err = "peer exchange timed out: a rank did not publish its records"; c->have_frame = false; c->results_valid = false; return SHAPES_E_NCCL; }
    if (st.error) {
        c->err = (st.error & ERR_PAIR_CAP) ? "capacity: max_pairs too small (required count in n_pairs; rows mode: on this or another rank)"
                                           : "capacity: max_contacts too small (required count in n_contacts)";
        return SHAPES_E_CAPACITY;
    }
    return SHAPES_OK;
}

// One frame: launch, wait, and once more if the grid plan turned out stale.
int run_frame(shapes_ctx *c, int64_t n_slots, const double *const in[7], double dt, double baumgarte,
              double slop, bool want_world, shapes_frame_out *out, bool own_slots_only = false)
{
    for (int attempt = 0;; ++attempt) {
        int rc = frame_launch(c, n_slots, in, dt, baumgarte, slop, want_world, own_slots_only);
        if (rc != SHAPES_OK) return rc;
        rc = frame_finish(c, out);
        if (rc != SHAPES_I_REPLAN) return rc;
        if (attempt >= 2) { c->err = "grid plan did not settle"; return SHAPES_E_CUDA; }
    }
}

template <typename T>
int fetch_col(shapes_ctx *c, T *dst, const T *src, int64_t n)
{
    if (!dst || n <= 0) return SHAPES_OK;
    CU_TRY(c, cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    return SHAPES_OK;
}

void world_free(shapes_ctx *c);
int world_grow(shapes_ctx *c, bool pairs, bool contacts);

} // namespace

#include "world_step.cuh"

extern "C" {

const char *shapes_version(void) { return "shapes_b200 0.1 (sm_100a)"; }

int shapes_create(shapes_ctx **out, int device_id, int64_t max_shapes, int64_t max_verts,
                  int64_t max_pairs, int64_t max_contacts)
{
    return create_impl(out, device_id, 0, 1, nullptr, max_shapes, max_verts, max_pairs, max_contacts);
}

int shapes_create_ranked(shapes_ctx **out, int device_id, int rank, int world_size, const void *nccl_id,
                         int64_t max_shapes, int64_t max_verts, int64_t max_pairs, int64_t max_contacts)
{
    return create_impl(out, device_id, rank, world_size, nccl_id, max_shapes, max_verts, max_pairs, max_contacts);
}

int shapes_nccl_unique_id(void *out_id)
{
    if (!out_id) return SHAPES_E_ARG;
    NcclApi &api = nccl_api();
    if (!api.ok) { g_create_error = "NCCL unavailable: " + api.error; return SHAPES_E_NCCL; }
    ncclUniqueId id;
    ncclResult_t r = api.GetUniqueId(&id);
    if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + api.GetErrorString(r); return SHAPES_E_NCCL; }
    std::memset(out_id, 0, SHAPES_NCCL_ID_BYTES);
    std::memcpy(out_id, &id, sizeof(id));
    return SHAPES_OK;
}

void shapes_destroy(shapes_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int q = 0; q < 4; ++q) if (c->graph_exec[q]) cudaGraphExecDestroy(c->graph_exec[q]);
    if (c->dbg_times && c->dbg_frames > 0) {
        std::string line = "[shapes_b200 rank " + std::to_string(c->rank) + "] rows-mode kernel ms (mean of " + std::to_string(c->dbg_frames) + " profiled frames):";
        for (int k = 1; k < c->dbg_marks_seen; ++k) { char buf[96]; std::snprintf(buf, sizeof(buf), " %s %.4f", c->dbg_name[k], c->dbg_sum[k] / c->dbg_frames); line += buf; }
        std::fprintf(stderr, "%s\n", line.c_str());
    }
    for (cudaEvent_t e : c->dbg_ev) cudaEventDestroy(e);
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->comm) nccl_api().CommDestroy(c->comm);
    world_free(c);
    for (void *p : c->allocs) cudaFree(p);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int k = 0; k <= SHAPES_N_STAGES; ++k) if (c->stage_ev[k]) cudaEventDestroy(c->stage_ev[k]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char *shapes_last_error(const shapes_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int shapes_set_hulls(shapes_ctx *c, int64_t n_slots, const uint8_t *alive, const int32_t *vert_offset,
                     const double *local_x, const double *local_y, const int32_t *ext_min, const int32_t *ext_max)
{
    return shapes_set_shapes(c, n_slots, alive, vert_offset, local_x, local_y, ext_min, ext_max, nullptr);
}

int shapes_set_shapes(shapes_ctx *c, int64_t n_slots, const uint8_t *alive, const int32_t *vert_offset,
                      const double *local_x, const double *local_y, const int32_t *ext_min, const int32_t *ext_max,
                      const double *radius)
{
    if (!c) return SHAPES_E_ARG;
    if (n_slots < 0 || n_slots > c->max_shapes || (n_slots > 0 && !vert_offset) ||
        (n_slots > 0 && vert_offset[n_slots] > 0 && (!local_x || !local_y)) ||
        ((ext_min == nullptr) != (ext_max == nullptr))) {
        c->err = "shapes_set_shapes: bad argument";
        return SHAPES_E_ARG;
    }
    CU_TRY(c, cudaSetDevice(c->device));
    const int64_t n_verts = n_slots > 0 ? vert_offset[n_slots] : 0;
    if (n_verts > c->max_verts || (n_slots > 0 && vert_offset[0] != 0)) {
        c->err = "shapes_set_hulls: vertex count exceeds max_verts (or vert_offset[0] != 0)";
        return SHAPES_E_ARG;
    }
    std::vector<uint8_t> live((size_t)n_slots, 1);
    if (alive) std::memcpy(live.data(), alive, (size_t)n_slots);
    // layout conversion (interleave x/y) and the static cell-size estimate: hull diameters
    std::vector<double2> inter((size_t)n_verts);
    std::vector<double> diam;
    diam.reserve((size_t)n_slots);
    int max_verts_seen = 0;
    bool any_circle = false;
    for (int64_t s = 0; s < n_slots; ++s) {
        const int32_t o = vert_offset[s], n = vert_offset[s + 1] - o;
        const bool circle = radius && radius[s] >= 0.0;
        if (n < 0 || (live[s] && !circle && n < 1) || (circle && n != 0)) {
            c->err = "shapes_set_shapes: a filled hull slot has no vertices (or a circle slot has some)";
            return SHAPES_E_ARG;
        }
        if (circle) { // Circle (Contact/Circle.hs:16-19): diameter 2r for the cell-size estimate
            if (live[s]) { any_circle = true; if (std::isfinite(radius[s])) diam.push_back(2.0 * radius[s]); }
            continue;
        }
        if (live[s] && n > max_verts_seen) max_verts_seen = n;
        double r2 = 0.0;
        for (int32_t k = 0; k < n; ++k) {
            inter[o + k] = make_double2(local_x[o + k], local_y[o + k]);
            const double d2 = local_x[o + k] * local_x[o + k] + local_y[o + k] * local_y[o + k];
            if (d2 > r2) r2 = d2;
        }
        if (live[s] && std::isfinite(r2)) diam.push_back(2.0 * std::sqrt(r2));
    }
    // Cell edge = largest diameter outside the "obviously big" set: among the 64 largest hulls,
    // everything above the last >= 2x gap (a floor among boxes) is left to the big-shape path.
    double cell = 1.0;
    if (!diam.empty()) {
        const size_t top = std::min<size_t>(diam.size(), 65);
        std::partial_sort(diam.begin(), diam.begin() + top, diam.end(), std::greater<double>());
        size_t cut = 0;
        for (size_t k = 1; k < top; ++k)
            if (diam[k - 1] > 2.0 * diam[k]) cut = k;
        cell = diam[cut] * (1.0 + 1e-6);
        if (!(cell > 0.0)) cell = 1.0;
    }
    c->auto_cell = cell;
    cudaStream_t s = c->stream;
    if (n_slots > 0) {
        CU_TRY(c, cudaMemcpyAsync(c->d_alive, live.data(), (size_t)n_slots, cudaMemcpyHostToDevice, s));
        CU_TRY(c, cudaMemcpyAsync(c->d_vert_offset, vert_offset, sizeof(int32_t) * (size_t)(n_slots + 1), cudaMemcpyHostToDevice, s));
        if (n_verts > 0) CU_TRY(c, cudaMemcpyAsync(c->d_local, inter.data(), sizeof(double2) * (size_t)n_verts, cudaMemcpyHostToDevice, s));
        if (any_circle) CU_TRY(c, cudaMemcpyAsync(c->d_radius, radius, sizeof(double) * (size_t)n_slots, cudaMemcpyHostToDevice, s));
        if (ext_min) {
            for (int64_t sl = 0; sl < n_slots; ++sl) {
                // _hullExtents entries index the hull's own vertices (ConvexHull.hs:63-77): anything else would be an
                // out-of-bounds read of the staged hull
                const int32_t o = vert_offset[sl], n = vert_offset[sl + 1] - o;
                for (int32_t k = 0; k < n; ++k)
                    if (ext_min[o + k] < 0 || ext_max[o + k] < 0 || ext_min[o + k] >= n || ext_max[o + k] >= n) {
                        c->err = "shapes_set_hulls: extent index outside its hull";
                        return SHAPES_E_ARG;
                    }
            }
            CU_TRY(c, cudaMemcpyAsync(c->d_ext_min, ext_min, sizeof(int32_t) * (size_t)n_verts, cudaMemcpyHostToDevice, s));
            CU_TRY(c, cudaMemcpyAsync(c->d_ext_max, ext_max, sizeof(int32_t) * (size_t)n_verts, cudaMemcpyHostToDevice, s));
        } else {
            k_hull_extents<<<grid_for(n_slots, 128, 1 << 30), 128, 0, s>>>((int)n_slots, c->d_vert_offset, c->d_local, c->d_ext_min, c->d_ext_max);
            ++c->launches;
        }
        k_pack_extents<<<grid_for(n_slots, 128, 1 << 30), 128, 0, s>>>((int)n_slots, c->d_vert_offset, c->d_ext_min, c->d_ext_max, c->d_ext_packed, c->d_hh);
        ++c->launches;
        CU_TRY(c, cudaGetLastError());
    }
    CU_TRY(c, cudaStreamSynchronize(s));
    c->n_slots = n_slots; c->n_verts = n_verts;
    c->max_hull_verts = max_verts_seen;
    c->has_circles = any_circle;
    c->P.radius = any_circle ? c->d_radius : nullptr;
    if (c->rw_arena)   // rows mode: slots nobody sweeps (dead ones) must read as "no partners"
        CU_TRY(c, cudaMemset(c->rw_arena + c->rwl.cq, 0, sizeof(uint32_t) * (size_t)std::max<int64_t>(c->chunk * c->world, 1)));
    c->hulls_set = true;
    c->have_frame = false;
    c->plan_valid = false;
    ++c->geometry_version;
    return SHAPES_OK;
}

int shapes_set_cell_size(shapes_ctx *c, double cell_size)
{
    if (!c) return SHAPES_E_ARG;
    c->user_cell = (cell_size > 0.0 && std::isfinite(cell_size)) ? cell_size : 0.0;
    c->plan_valid = false;
    return SHAPES_OK;
}

// Grow the pair / contact capacities in place.  Everything that describes the world between frames survives: the
// previous frame's key columns (both halves of the double buffer), the Lagrangian cache supplied for them, the
// uploaded world (world_step.cuh) and the grid plan -- i.e. the reference's EngineCache (Engine/Main.hs:32,60-68) is
// NOT dropped, which re-creating the ctx would do.  Per-frame result arrays are simply re-allocated.
int shapes_grow(shapes_ctx *c, int64_t max_pairs, int64_t max_contacts)
{
    if (!c) return SHAPES_E_ARG;
    if (c->world != 1) { c->err = "shapes_grow: single-GPU ctx only (a multi-rank job re-creates its ctxs collectively)"; return SHAPES_E_ARG; }
    max_pairs = std::max(max_pairs, c->max_pairs); max_contacts = std::max(max_contacts, c->max_contacts);
    if (max_pairs > 0x7ffffff0ll || max_contacts > 0xfffffff0ll || (c->ws && max_pairs >= (int64_t)SOLVE_NODE_MASK)) {
        c->err = "shapes_grow: capacity out of range";
        return SHAPES_E_ARG;
    }
    if (max_pairs == c->max_pairs && max_contacts == c->max_contacts) return SHAPES_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    Params &P = c->P;
    const size_t oldP = (size_t)std::max<int64_t>(c->max_pairs, 1), newP = (size_t)std::max<int64_t>(max_pairs, 1);
    const size_t oldC = (size_t)std::max<int64_t>(c->max_contacts, 1), newC = (size_t)std::max<int64_t>(max_contacts, 1);
    // one buffer: new allocation, the first `keep` elements copied over, old one released
    auto regrow = [&](auto **p, size_t count, size_t keep) -> int {
        using T = std::remove_pointer_t<std::remove_pointer_t<decltype(p)>>;
        T *q = nullptr;
        CU_TRY(c, cudaMalloc(reinterpret_cast<void **>(&q), count * sizeof(T)));
        if (keep > 0 && *p) CU_TRY(c, cudaMemcpy(q, *p, keep * sizeof(T), cudaMemcpyDeviceToDevice));
        bool tracked = false;
        for (void *&a : c->allocs) if (a == static_cast<void *>(*p)) { cudaFree(a); a = q; tracked = true; }
        if (!tracked) c->allocs.push_back(q);
        *p = q;
        return SHAPES_OK;
    };
#define GROW(ptr, count, keep) do { int rc__ = regrow(ptr, count, keep); if (rc__ != SHAPES_OK) return rc__; } while (0)
    if (newP != oldP) {
        const bool lists = c->use_sorted;
        if (lists) { GROW(&P.w_i, newP, 0); GROW(&c->d_w_j, newP, 0); GROW(&P.w_a, newP, 0); P.w_j = c->d_w_j; }
        GROW(&c->d_pair_i, newP, 0); GROW(&c->d_pair_j, newP, 0); GROW(&c->d_man, newP, 0); GROW(&c->d_ccnt, newP, 0);
        P.pair_i = c->d_pair_i; P.pair_j = c->d_pair_j; P.man = c->d_man; P.ccnt = c->d_ccnt; P.sat_ccnt = P.ccnt;
        GROW(&P.coff, newP, 0);
        CU_TRY(c, cudaMemset(P.ccnt, 0, newP * sizeof(uint32_t)));
        size_t cb = 0;
        CU_TRY(c, cub::DeviceScan::ExclusiveSum(nullptr, cb, P.ccnt, P.coff, (int)newP, c->stream));
        if (cb > c->scan_tmp_bytes) {
            uint8_t *tmp = static_cast<uint8_t *>(c->d_scan_tmp);
            GROW(&tmp, cb, 0);
            c->d_scan_tmp = tmp; c->scan_tmp_bytes = cb;
        }
    }
    if (newC != oldC) {
        GROW(&P.row_map, newC, 0);
        // the key columns of the last completed frame and the cache given for them must survive
        const size_t keep = c->have_frame ? (size_t)std::min<int64_t>(c->last_contacts, (int64_t)oldC) : 0;
        GROW(&P.key_i, newC, keep); GROW(&P.key_j, newC, keep); GROW(&P.feat_a, newC, keep); GROW(&P.feat_b, newC, keep);
        for (int q = 0; q < 4; ++q) GROW(&c->alt_key[q], newC, keep);
        GROW(&c->d_cache_np, newC, keep); GROW(&c->d_cache_f, newC, keep);
        GROW(&P.flip, newC, 0); GROW(&P.warm_np, newC, 0); GROW(&P.warm_f, newC, 0); GROW(&P.warm_hit, newC, 0);
        double **cols[] = { &P.normal_x, &P.normal_y, &P.center_x, &P.center_y, &P.depth, &P.b_np,
                            &P.ra_x, &P.ra_y, &P.rb_x, &P.rb_y, &P.rn_x, &P.rn_y, &P.inv_eff_np, &P.inv_eff_f };
        for (double **col : cols) GROW(col, newC, 0);
        for (int q = 0; q < 6; ++q) { GROW(&P.j_np[q], newC, 0); GROW(&P.j_f[q], newC, 0); }
    }
    c->max_pairs = max_pairs; c->max_contacts = max_contacts;
    P.max_pairs = max_pairs; P.max_contacts = max_contacts;
    const int rc = world_grow(c, newP != oldP, newC != oldC);
    if (rc != SHAPES_OK) return rc;
#undef GROW
    // captured frames bake the old pointers in
    for (int q = 0; q < 4; ++q) if (c->graph_exec[q]) { cudaGraphExecDestroy(c->graph_exec[q]); c->graph_exec[q] = nullptr; }
    c->results_valid = false;
    return SHAPES_OK;
}

int shapes_frame_device(shapes_ctx *c, int64_t n_slots, const double *pos_x, const double *pos_y,
                        const double *rot, const double *cos_rot, const double *sin_rot,
                        const double *inv_lin, const double *inv_rot, double dt, double baumgarte,
                        double slop, shapes_frame_out *out)
{
    if (!c) return SHAPES_E_ARG;
    const double *in[7] = { pos_x, pos_y, rot, cos_rot, sin_rot, inv_lin, inv_rot };
    return run_frame(c, n_slots, in, dt, baumgarte, slop, false, out);
}

static int fetch_enqueue(shapes_ctx *c, shapes_frame_out *out)
{
    if (!c || !out) return SHAPES_E_ARG;
    if (!c->have_frame || !c->results_valid) { c->err = "shapes_fetch: no completed frame"; return SHAPES_E_ARG; }
    CU_TRY(c, cudaSetDevice(c->device));
    const Params &P = c->P;
    const int64_t np = c->last_pairs, nc = c->last_contacts;
    int rc = SHAPES_OK;
#define FETCH(dst, src, n) do { rc = fetch_col(c, dst, src, n); if (rc != SHAPES_OK) return rc; } while (0)
    FETCH(out->pair_i, P.pair_i, np); FETCH(out->pair_j, P.pair_j, np);
    FETCH(out->key_i, P.key_i, nc); FETCH(out->key_j, P.key_j, nc);
    FETCH(out->feat_a, P.feat_a, nc); FETCH(out->feat_b, P.feat_b, nc);
    FETCH(out->flip, P.flip, nc);
    FETCH(out->normal_x, P.normal_x, nc); FETCH(out->normal_y, P.normal_y, nc);
    FETCH(out->center_x, P.center_x, nc); FETCH(out->center_y, P.center_y, nc);
    FETCH(out->depth, P.depth, nc);
    for (int q = 0; q < 6; ++q) { FETCH(out->j_np[q], P.j_np[q], nc); FETCH(out->j_f[q], P.j_f[q], nc); }
    FETCH(out->b_np, P.b_np, nc);
    FETCH(out->ra_x, P.ra_x, nc); FETCH(out->ra_y, P.ra_y, nc);
    FETCH(out->rb_x, P.rb_x, nc); FETCH(out->rb_y, P.rb_y, nc);
    FETCH(out->rn_x, P.rn_x, nc); FETCH(out->rn_y, P.rn_y, nc);
    FETCH(out->inv_eff_np, P.inv_eff_np, nc); FETCH(out->inv_eff_f, P.inv_eff_f, nc);
    if (c->warm_done) { FETCH(out->warm_np, P.warm_np, nc); FETCH(out->warm_f, P.warm_f, nc); FETCH(out->warm_hit, P.warm_hit, nc); }
    if (out->aabb_min_x || out->aabb_max_x || out->aabb_min_y || out->aabb_max_y) {
        const int64_t N = c->n_slots;
        if (!c->d_split) { rc = dev_alloc(c, &c->d_split, 4 * (size_t)std::max<int64_t>(c->max_shapes, 1)); if (rc) return rc; }
        double *a = c->d_split, *b = a + c->max_shapes, *cc = b + c->max_shapes, *d = cc + c->max_shapes;
        if (N > 0) { k_split_boxes<<<grid_for(N, 256, 1 << 30), 256, 0, c->stream>>>((int)N, P.box, a, b, cc, d); ++c->launches; }
        FETCH(out->aabb_min_x, a, N); FETCH(out->aabb_max_x, b, N);
        FETCH(out->aabb_min_y, cc, N); FETCH(out->aabb_max_y, d, N);
    }
    if (out->world_x && c->d_world_x) { FETCH(out->world_x, c->d_world_x, c->n_verts); FETCH(out->world_y, c->d_world_y, c->n_verts); }
#undef FETCH
    return SHAPES_OK;
}

static int fetch_complete(shapes_ctx *c, shapes_frame_out *out)
{
    const int64_t nc = c->last_contacts;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (out->b_f && nc > 0) std::memset(out->b_f, 0, sizeof(double) * (size_t)nc); // Friction.toConstraint: b = 0 (Friction.hs:26-29)
    if (!c->warm_done && nc > 0) { // no cache was supplied: every contact is a newCache (ContactLagrangian 0 0)
        if (out->warm_np) std::memset(out->warm_np, 0, sizeof(double) * (size_t)nc);
        if (out->warm_f) std::memset(out->warm_f, 0, sizeof(double) * (size_t)nc);
        if (out->warm_hit) std::memset(out->warm_hit, 0, (size_t)nc);
    }
    return SHAPES_OK;
}

int shapes_fetch(shapes_ctx *c, shapes_frame_out *out)
{
    const int rc = fetch_enqueue(c, out);
    return rc != SHAPES_OK ? rc : fetch_complete(c, out);
}

int shapes_frame(shapes_ctx *c, int64_t n_slots, const double *pos_x, const double *pos_y,
                 const double *rot, const double *cos_rot, const double *sin_rot,
                 const double *inv_lin, const double *inv_rot, double dt, double baumgarte,
                 double slop, shapes_frame_out *out)
{
    if (!c || !out) return SHAPES_E_ARG;
    if (n_slots != c->n_slots || !c->hulls_set) { c->err = "shapes_frame: n_slots differs from shapes_set_hulls"; return SHAPES_E_ARG; }
    CU_TRY(c, cudaSetDevice(c->device));
    cudaEvent_t t0, t1;
    CU_TRY(c, cudaEventCreate(&t0));
    CU_TRY(c, cudaEventCreate(&t1));
    CU_TRY(c, cudaEventRecord(t0, c->stream));
    const double *host[7] = { pos_x, pos_y, rot, cos_rot, sin_rot, inv_lin, inv_rot };
    const double *dev[7];
    // With the peer exchange every rank uploads only ITS slots (kernels pull foreign bodies from
    // their owners' columns); otherwise the whole world is uploaded on every rank.
    const bool rows = c->world > 1 && c->rows_ready && c->use_rows;
    const bool own_only = rows || (c->world > 1 && c->peers_ready && c->use_p2p);
    const int fpar = (int)((c->frame_no + 1) & 1);
    int64_t lo[2] = { own_only ? std::min<int64_t>(c->rank * c->chunk, n_slots) : 0, 0 };
    int64_t hi[2] = { own_only ? std::min<int64_t>((c->rank + 1) * c->chunk, n_slots) : n_slots, 0 };
    if (rows) rows_home_blocks(c, c->rank, n_slots, lo, hi);      // my two home blocks
    for (int k = 0; k < 7; ++k) {
        dev[k] = nullptr;
        double *dst = rows ? reinterpret_cast<double *>(c->rw_arena + c->rwl.in[k]) : c->world > 1 ? c->d_in2[fpar][k] : c->d_in[k];
        for (int q = 0; q < 2; ++q)
            if (host[k] && hi[q] > lo[q])
                CU_TRY(c, cudaMemcpyAsync(dst + lo[q], host[k] + lo[q], sizeof(double) * (size_t)(hi[q] - lo[q]), cudaMemcpyHostToDevice, c->stream));
        if (host[k]) dev[k] = dst;
    }
    const bool want_world = out->world_x && out->world_y;
    if (want_world && !c->d_world_x) {
        int rc = dev_alloc(c, &c->d_world_x, (size_t)c->max_verts); if (rc) return rc;
        rc = dev_alloc(c, &c->d_world_y, (size_t)c->max_verts); if (rc) return rc;
    }
    int rc = run_frame(c, n_slots, dev, dt, baumgarte, slop, want_world, out, own_only);
    if (rc == SHAPES_OK) rc = shapes_fetch(c, out);
    cudaEventRecord(t1, c->stream);
    cudaEventSynchronize(t1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    out->total_ms = ms;
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    return rc;
}

static int set_cache(shapes_ctx *c, int64_t n_prev, const double *np, const double *f, cudaMemcpyKind kind)
{
    if (!c) return SHAPES_E_ARG;
    if (!c->have_frame || n_prev != c->last_contacts || (n_prev > 0 && (!np || !f))) {
        c->err = "shapes_set_lagrangian_cache: n_prev must equal the last completed frame's n_contacts";
        return SHAPES_E_ARG;
    }
    CU_TRY(c, cudaSetDevice(c->device));
    if (n_prev > 0) {
        CU_TRY(c, cudaMemcpyAsync(c->d_cache_np, np, sizeof(double) * (size_t)n_prev, kind, c->stream));
        CU_TRY(c, cudaMemcpyAsync(c->d_cache_f, f, sizeof(double) * (size_t)n_prev, kind, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
    }
    c->cache_valid = true;
    return SHAPES_OK;
}

int shapes_set_lagrangian_cache(shapes_ctx *c, int64_t n_prev, const double *lambda_np, const double *lambda_f)
{
    return set_cache(c, n_prev, lambda_np, lambda_f, cudaMemcpyHostToDevice);
}

int shapes_set_lagrangian_cache_device(shapes_ctx *c, int64_t n_prev, const double *lambda_np, const double *lambda_f)
{
    return set_cache(c, n_prev, lambda_np, lambda_f, cudaMemcpyDeviceToDevice);
}

int shapes_device_view_get(shapes_ctx *c, shapes_device_view *v)
{
    if (!c || !v) return SHAPES_E_ARG;
    if (!c->have_frame || !c->results_valid) { c->err = "shapes_device_view_get: no completed frame"; return SHAPES_E_ARG; }
    const Params &P = c->P;
    v->n_pairs = c->last_pairs; v->n_contacts = c->last_contacts;
    v->pair_i = P.pair_i; v->pair_j = P.pair_j;
    v->key_i = P.key_i; v->key_j = P.key_j; v->feat_a = P.feat_a; v->feat_b = P.feat_b;
    v->flip = P.flip;
    v->normal_x = P.normal_x; v->normal_y = P.normal_y; v->center_x = P.center_x; v->center_y = P.center_y;
    v->depth = P.depth;
    for (int q = 0; q < 6; ++q) { v->j_np[q] = P.j_np[q]; v->j_f[q] = P.j_f[q]; }
    v->b_np = P.b_np;
    v->ra_x = P.ra_x; v->ra_y = P.ra_y; v->rb_x = P.rb_x; v->rb_y = P.rb_y; v->rn_x = P.rn_x; v->rn_y = P.rn_y;
    v->inv_eff_np = P.inv_eff_np; v->inv_eff_f = P.inv_eff_f;
    v->warm_np = c->warm_done ? P.warm_np : nullptr; v->warm_f = c->warm_done ? P.warm_f : nullptr;
    v->warm_hit = c->warm_done ? P.warm_hit : nullptr;
    v->aabb = reinterpret_cast<const double *>(P.box);
    return SHAPES_OK;
}

int shapes_ipc_export(shapes_ctx *c, void *out)
{
    if (!c || !out) return SHAPES_E_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h[22];
    std::memset(h, 0, sizeof(h));
    if (c->rw_arena) CU_TRY(c, cudaIpcGetMemHandle(&h[21], c->rw_arena));
    for (int par = 0; par < 2; ++par)
        for (int k = 0; k < 7; ++k) CU_TRY(c, cudaIpcGetMemHandle(&h[7 + par * 7 + k], c->d_in2[par][k]));
    CU_TRY(c, cudaIpcGetMemHandle(&h[0], c->d_box2[0]));
    CU_TRY(c, cudaIpcGetMemHandle(&h[1], c->d_box2[1]));
    CU_TRY(c, cudaIpcGetMemHandle(&h[2], c->d_bounds2[0]));
    CU_TRY(c, cudaIpcGetMemHandle(&h[3], c->d_bounds2[1]));
    CU_TRY(c, cudaIpcGetMemHandle(&h[4], c->d_flags));
    CU_TRY(c, cudaIpcGetMemHandle(&h[5], c->d_gkeys2[0]));
    CU_TRY(c, cudaIpcGetMemHandle(&h[6], c->d_gkeys2[1]));
    static_assert(sizeof(h) <= SHAPES_IPC_BYTES, "ipc blob size");
    std::memset(out, 0, SHAPES_IPC_BYTES);
    std::memcpy(out, h, sizeof(h));
    return SHAPES_OK;
}

int shapes_ipc_import(shapes_ctx *c, const void *all_blobs)
{
    if (!c || !all_blobs || c->world < 2) return SHAPES_E_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) {
            c->peer_box2[0][r] = c->d_box2[0]; c->peer_box2[1][r] = c->d_box2[1];
            c->peer_bounds2[0][r] = c->d_bounds2[0]; c->peer_bounds2[1][r] = c->d_bounds2[1];
            c->peer_flags[r] = c->d_flags;
            c->peer_keys2[0][r] = c->d_gkeys2[0]; c->peer_keys2[1][r] = c->d_gkeys2[1];
            for (int par = 0; par < 2; ++par)
                for (int k = 0; k < 7; ++k) c->peer_in2[par][k][r] = c->d_in2[par][k];
            continue;
        }
        cudaIpcMemHandle_t h[22];
        std::memcpy(h, static_cast<const char *>(all_blobs) + (size_t)r * SHAPES_IPC_BYTES, sizeof(h));
        void *p[22];
        for (int q = 0; q < 22; ++q) {
            if (q == 21 && !c->rw_arena) { p[q] = nullptr; continue; }
            CU_TRY(c, cudaIpcOpenMemHandle(&p[q], h[q], cudaIpcMemLazyEnablePeerAccess));
            c->ipc_opened.push_back(p[q]);
        }
        c->peer_arena[r] = static_cast<char *>(p[21]);
        c->peer_box2[0][r] = static_cast<Box *>(p[0]); c->peer_box2[1][r] = static_cast<Box *>(p[1]);
        c->peer_bounds2[0][r] = static_cast<unsigned long long *>(p[2]);
        c->peer_bounds2[1][r] = static_cast<unsigned long long *>(p[3]);
        c->peer_flags[r] = static_cast<unsigned long long *>(p[4]);
        c->peer_keys2[0][r] = static_cast<uint32_t *>(p[5]); c->peer_keys2[1][r] = static_cast<uint32_t *>(p[6]);
        for (int par = 0; par < 2; ++par)
            for (int k = 0; k < 7; ++k) c->peer_in2[par][k][r] = static_cast<const double *>(p[7 + par * 7 + k]);
    }
    c->peers_ready = true;
    c->rows_ready = c->rw_arena != nullptr;
    c->plan_valid = false;
    return SHAPES_OK;
}

int shapes_rank_info(shapes_ctx *c, int64_t *own_lo, int64_t *own_hi, int64_t *all_pairs, int64_t *all_contacts)
{
    if (!c) return SHAPES_E_ARG;
    if (own_lo) *own_lo = std::min<int64_t>(c->rank * c->chunk, c->n_slots);
    if (own_hi) *own_hi = std::min<int64_t>((c->rank + 1) * c->chunk, c->n_slots);
    const bool rows = c->world > 1 && c->rows_ready && c->use_rows;
    for (int r = 0; r < c->world; ++r) {
        if (all_pairs) all_pairs[r] = rows ? c->h_counts[4 * r] + c->h_counts[4 * r + 1] : c->h_counts[2 * r];
        if (all_contacts) all_contacts[r] = rows ? c->h_counts[4 * r + 2] + c->h_counts[4 * r + 3] : c->h_counts[2 * r + 1];
    }
    return SHAPES_OK;
}

int shapes_rank_segments(shapes_ctx *c, int rank, int64_t seg_lo[2], int64_t seg_hi[2], int64_t seg_pairs[2], int64_t seg_contacts[2])
{
    if (!c || rank < 0 || rank >= c->world) return SHAPES_E_ARG;
    const bool rows = c->world > 1 && c->rows_ready && c->use_rows;
    if (rows) {
        int64_t lo[2], hi[2];
        rows_home_blocks(c, rank, c->n_slots, lo, hi);
        // run 0 = the high block (first in the rank's arrays), run 1 = the low block
        if (seg_lo) { seg_lo[0] = lo[1]; seg_lo[1] = lo[0]; }
        if (seg_hi) { seg_hi[0] = hi[1]; seg_hi[1] = hi[0]; }
        if (seg_pairs) { seg_pairs[0] = c->h_counts[4 * rank]; seg_pairs[1] = c->h_counts[4 * rank + 1]; }
        if (seg_contacts) { seg_contacts[0] = c->h_counts[4 * rank + 2]; seg_contacts[1] = c->h_counts[4 * rank + 3]; }
    } else {
        // one slot range per rank: everything is "run 1" (the global order is then rank G-1's rows, G-2's, ...)
        if (seg_lo) { seg_lo[1] = std::min<int64_t>(rank * c->chunk, c->n_slots); seg_lo[0] = 0; }
        if (seg_hi) { seg_hi[1] = std::min<int64_t>((rank + 1) * c->chunk, c->n_slots); seg_hi[0] = 0; }
        const int64_t *h = c->h_counts;
        if (seg_pairs) { seg_pairs[1] = c->world == 1 ? h[0] : h[2 * rank]; seg_pairs[0] = 0; }
        if (seg_contacts) { seg_contacts[1] = c->world == 1 ? h[1] : h[2 * rank + 1]; seg_contacts[0] = 0; }
    }
    return SHAPES_OK;
}

/* ---- one process, several GPUs (the host is ONE single-threaded ST computation) ---- */

struct shapes_multi {
    int n = 0;
    std::vector<shapes_ctx *> rank;
    std::string err;
};

static thread_local std::string g_multi_error;

void shapes_multi_destroy(shapes_multi *m)
{
    if (!m) return;
    for (shapes_ctx *c : m->rank) shapes_destroy(c);
    delete m;
}

const char *shapes_multi_last_error(const shapes_multi *m) { return m ? m->err.c_str() : g_multi_error.c_str(); }

int shapes_create_multi(shapes_multi **out, int n_gpus, const int *device_ids, int64_t max_shapes, int64_t max_verts,
                        int64_t max_pairs_per_gpu, int64_t max_contacts_per_gpu)
{
    if (!out || n_gpus < 1 || n_gpus > SHAPES_MAX_RANKS) { g_multi_error = "shapes_create_multi: bad argument"; return SHAPES_E_ARG; }
    shapes_multi *m = new shapes_multi();
    m->n = n_gpus;
    auto fail = [&](int rc, const std::string &why) { g_multi_error = why; shapes_multi_destroy(m); return rc; };
    for (int r = 0; r < n_gpus; ++r) {
        shapes_ctx *c = nullptr;
        const int dev = device_ids ? device_ids[r] : r;
        const int rc = create_impl(&c, dev, r, n_gpus, nullptr, max_shapes, max_verts, max_pairs_per_gpu, max_contacts_per_gpu, false);
        if (rc != SHAPES_OK) return fail(rc, g_create_error);
        m->rank.push_back(c);
    }
    if (n_gpus > 1) {
        // peers are plain pointers inside one process: enable peer access both ways and hand every rank the others' arenas
        for (int a = 0; a < n_gpus; ++a) {
            if (cudaSetDevice(m->rank[a]->device) != cudaSuccess) return fail(SHAPES_E_CUDA, "cudaSetDevice");
            for (int b = 0; b < n_gpus; ++b) {
                if (a == b || m->rank[a]->device == m->rank[b]->device) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, m->rank[a]->device, m->rank[b]->device);
                if (!can) return fail(SHAPES_E_CUDA, "shapes_create_multi: the GPUs cannot access each other's memory");
                const cudaError_t e = cudaDeviceEnablePeerAccess(m->rank[b]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(SHAPES_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
        }
        for (int a = 0; a < n_gpus; ++a) {
            shapes_ctx *c = m->rank[a];
            if (!c->rw_arena) return fail(SHAPES_E_ARG, "shapes_create_multi: rows mode is disabled (SHAPES_B200_NO_ROWS) or max_pairs_per_gpu >= 2^28");
            for (int b = 0; b < n_gpus; ++b) c->peer_arena[b] = m->rank[b]->rw_arena;
            c->peers_ready = true; c->rows_ready = true; c->plan_valid = false;
        }
    }
    *out = m;
    return SHAPES_OK;
}

shapes_ctx *shapes_multi_rank(shapes_multi *m, int rank) { return (m && rank >= 0 && rank < m->n) ? m->rank[rank] : nullptr; }

int shapes_multi_set_shapes(shapes_multi *m, int64_t n_slots, const uint8_t *alive, const int32_t *vert_offset,
                            const double *local_x, const double *local_y, const int32_t *ext_min, const int32_t *ext_max,
                            const double *radius)
{
    if (!m) return SHAPES_E_ARG;
    for (shapes_ctx *c : m->rank) {
        const int rc = shapes_set_shapes(c, n_slots, alive, vert_offset, local_x, local_y, ext_min, ext_max, radius);
        if (rc != SHAPES_OK) { m->err = c->err; return rc; }
    }
    return SHAPES_OK;
}

// One frame on all GPUs from ONE host thread: every rank's slice of the body columns goes up over its own PCIe link,
// the frames are enqueued back to back (their device-side barriers meet on the GPUs), and every rank's slice of the
// result comes down, again over its own link, straight to its global row offset in the caller's buffers.
int shapes_multi_frame(shapes_multi *m, int64_t n_slots, const double *pos_x, const double *pos_y,
                       const double *rot, const double *cos_rot, const double *sin_rot,
                       const double *inv_lin, const double *inv_rot, double dt, double baumgarte,
                       double slop, shapes_frame_out *out)
{
    if (!m || !out) return SHAPES_E_ARG;
    const int G = m->n;
    if (G == 1) return shapes_frame(m->rank[0], n_slots, pos_x, pos_y, rot, cos_rot, sin_rot, inv_lin, inv_rot, dt, baumgarte, slop, out);
    const double *host[7] = { pos_x, pos_y, rot, cos_rot, sin_rot, inv_lin, inv_rot };
    std::vector<std::array<const double *, 7>> dev((size_t)G);
    auto fail = [&](shapes_ctx *c, int rc) { m->err = c->err; return rc; };
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaSetDevice(m->rank[0]->device);
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaEventRecord(t0, m->rank[0]->stream);
    for (int r = 0; r < G; ++r) {
        shapes_ctx *c = m->rank[r];
        if (n_slots != c->n_slots || !c->hulls_set) { c->err = "shapes_multi_frame: n_slots differs from shapes_multi_set_shapes"; return fail(c, SHAPES_E_ARG); }
        CU_TRY(c, cudaSetDevice(c->device));
        int64_t lo[2], hi[2];
        rows_home_blocks(c, c->rank, n_slots, lo, hi);
        for (int k = 0; k < 7; ++k) {
            dev[r][k] = nullptr;
            double *dst = reinterpret_cast<double *>(c->rw_arena + c->rwl.in[k]);
            for (int q = 0; q < 2; ++q)
                if (host[k] && hi[q] > lo[q])
                    CU_TRY(c, cudaMemcpyAsync(dst + lo[q], host[k] + lo[q], sizeof(double) * (size_t)(hi[q] - lo[q]), cudaMemcpyHostToDevice, c->stream));
            if (host[k]) dev[r][k] = dst;
        }
    }
    std::vector<shapes_frame_out> part((size_t)G, *out);
    for (int attempt = 0;; ++attempt) {
        for (int r = 0; r < G; ++r) {
            const int rc = frame_launch(m->rank[r], n_slots, dev[r].data(), dt, baumgarte, slop, false, true);
            if (rc != SHAPES_OK) return fail(m->rank[r], rc);
        }
        bool replan = false;
        int bad = SHAPES_OK;
        shapes_ctx *bad_ctx = nullptr;
        for (int r = 0; r < G; ++r) {
            const int rc = frame_finish(m->rank[r], &part[r]);
            if (rc == SHAPES_I_REPLAN) replan = true;
            else if (rc != SHAPES_OK && bad == SHAPES_OK) { bad = rc; bad_ctx = m->rank[r]; }
        }
        if (bad != SHAPES_OK) {
            out->n_pairs = 0; out->n_contacts = 0;
            for (int r = 0; r < G; ++r) { out->n_pairs = std::max(out->n_pairs, part[r].n_pairs); out->n_contacts = std::max(out->n_contacts, part[r].n_contacts); }
            return fail(bad_ctx, bad);   // capacity: the largest per-GPU requirement is reported
        }
        if (!replan) break;
        if (attempt >= 2) { m->err = "grid plan did not settle"; return SHAPES_E_CUDA; }
    }
    // The global descending order walks the 2G home blocks from the top: block 2G-1 (rank 0's high block), 2G-2 (rank
    // 1's), ..., G (rank G-1's), then the low blocks G-1 (rank G-1) ... 0 (rank 0).  Every rank holds its high block's
    // rows first, then its low block's: two copies per column and rank, each to its global offset.
    const int64_t *cnts = m->rank[0]->h_counts;      // per rank: pairs hi, pairs lo, contacts hi, contacts lo
    std::vector<int64_t> pair_at((size_t)2 * G), row_at((size_t)2 * G);   // [2 r + 0] = rank r's high run, [2 r + 1] = its low run
    int64_t pair_off = 0, row_off = 0;
    for (int r = 0; r < G; ++r) { pair_at[2 * r] = pair_off; row_at[2 * r] = row_off; pair_off += cnts[4 * r]; row_off += cnts[4 * r + 2]; }
    for (int r = G - 1; r >= 0; --r) { pair_at[2 * r + 1] = pair_off; row_at[2 * r + 1] = row_off; pair_off += cnts[4 * r + 1]; row_off += cnts[4 * r + 3]; }
    for (int r = 0; r < G; ++r) {
        shapes_ctx *c = m->rank[r];
        const Params &P = c->P;
        const int64_t np[2] = { cnts[4 * r], cnts[4 * r + 1] }, nc[2] = { cnts[4 * r + 2], cnts[4 * r + 3] };
        CU_TRY(c, cudaSetDevice(c->device));
        for (int q = 0; q < 2; ++q) {
            const int64_t ps = q ? np[0] : 0, cs = q ? nc[0] : 0;         // where the run starts in the rank's arrays
            const int64_t pd = pair_at[2 * r + q], cd = row_at[2 * r + q]; // ... and in the caller's
            int rc = SHAPES_OK;
#define RUNP(field, src) do { if (out->field) { rc = fetch_col(c, out->field + pd, src + ps, np[q]); if (rc) return fail(c, rc); } } while (0)
#define RUNC(field, src) do { if (out->field) { rc = fetch_col(c, out->field + cd, src + cs, nc[q]); if (rc) return fail(c, rc); } } while (0)
            RUNP(pair_i, P.pair_i); RUNP(pair_j, P.pair_j);
            RUNC(key_i, P.key_i); RUNC(key_j, P.key_j); RUNC(feat_a, P.feat_a); RUNC(feat_b, P.feat_b); RUNC(flip, P.flip);
            RUNC(normal_x, P.normal_x); RUNC(normal_y, P.normal_y); RUNC(center_x, P.center_x); RUNC(center_y, P.center_y); RUNC(depth, P.depth);
            for (int k = 0; k < 6; ++k) { RUNC(j_np[k], P.j_np[k]); RUNC(j_f[k], P.j_f[k]); }
            RUNC(b_np, P.b_np);
            RUNC(ra_x, P.ra_x); RUNC(ra_y, P.ra_y); RUNC(rb_x, P.rb_x); RUNC(rb_y, P.rb_y); RUNC(rn_x, P.rn_x); RUNC(rn_y, P.rn_y);
            RUNC(inv_eff_np, P.inv_eff_np); RUNC(inv_eff_f, P.inv_eff_f);
            if (c->warm_done) { RUNC(warm_np, P.warm_np); RUNC(warm_f, P.warm_f); RUNC(warm_hit, P.warm_hit); }
#undef RUNP
#undef RUNC
        }
    }
    for (int r = 0; r < G; ++r) { shapes_ctx *c = m->rank[r]; CU_TRY(c, cudaSetDevice(c->device)); CU_TRY(c, cudaStreamSynchronize(c->stream)); }
    if (out->b_f && row_off > 0) std::memset(out->b_f, 0, sizeof(double) * (size_t)row_off);   // Friction.toConstraint: b = 0
    if (!m->rank[0]->warm_done && row_off > 0) {
        if (out->warm_np) std::memset(out->warm_np, 0, sizeof(double) * (size_t)row_off);
        if (out->warm_f) std::memset(out->warm_f, 0, sizeof(double) * (size_t)row_off);
        if (out->warm_hit) std::memset(out->warm_hit, 0, (size_t)row_off);
    }
    out->n_pairs = pair_off; out->n_contacts = row_off;
    out->n_big = part[0].n_big; out->grid_w = part[0].grid_w; out->grid_h = part[0].grid_h; out->cell_size = part[0].cell_size;
    float dev_ms = 0.f;
    for (int r = 0; r < G; ++r) dev_ms = std::max(dev_ms, part[r].device_ms);
    out->device_ms = dev_ms;
    cudaSetDevice(m->rank[0]->device);
    cudaEventRecord(t1, m->rank[0]->stream);
    cudaEventSynchronize(t1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    out->total_ms = ms;
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    return SHAPES_OK;
}

void *shapes_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

void shapes_host_free(void *p) { if (p) cudaFreeHost(p); }

void *shapes_stream(shapes_ctx *c) { return c ? (void *)c->stream : nullptr; }

int64_t shapes_launch_count(const shapes_ctx *c) { return c ? c->launches : 0; }

int shapes_last_frame_info(shapes_ctx *c, shapes_frame_info *info)
{
    if (!c || !info) return SHAPES_E_ARG;
    if (!c->have_frame || !c->results_valid) { c->err = "shapes_last_frame_info: no completed frame"; return SHAPES_E_ARG; }
    info->pairs_with_contacts = (int64_t)c->h_state->n_pairs_hit;
    info->sorted_mode = c->P.sorted_mode;
    info->sat_kernel = c->has_circles ? SHAPES_SAT_PER_THREAD_CIRCLES : c->max_hull_verts <= 4 ? SHAPES_SAT_PER_THREAD_BOXES
                     : c->use_coop ? SHAPES_SAT_COOP : SHAPES_SAT_PER_THREAD;
    return SHAPES_OK;
}

int shapes_set_profiling(shapes_ctx *c, int enabled)
{
    if (!c) return SHAPES_E_ARG;
    c->profiling = enabled != 0;
    return SHAPES_OK;
}

int shapes_stage_ms(const shapes_ctx *c, float *out_ms)
{
    if (!c || !out_ms) return SHAPES_E_ARG;
    for (int k = 0; k < SHAPES_N_STAGES; ++k) out_ms[k] = c->stage_ms[k];
    return SHAPES_OK;
}

const char *shapes_stage_name(int stage)
{
    static const char *names[SHAPES_N_STAGES] = { "transform_aabb", "allgather_aabb", "grid_keys", "cell_scan",
                                                  "scatter_sorted", "sweep_count", "scan_offsets", "sweep_emit",
                                                  "manifolds", "scan_rows", "contact_rows", "warm_join" };
    return (stage >= 0 && stage < SHAPES_N_STAGES) ? names[stage] : "";
}

} // extern "C"

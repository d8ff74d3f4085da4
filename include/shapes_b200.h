/*
 * shapes_b200.h -- C ABI of the B200 collision pipeline for ublubu/shapes.
 *
 * The reference (Haskell, /root/reference) has no FFI boundary of its own
 * (no `foreign import` anywhere).  The boundary is therefore defined by the
 * three expressions this library replaces inside
 * Physics.Engine.Main.updateWorld (shapes/src/Physics/Engine/Main.hs:71-86):
 *
 *   keys      <- G.culledKeys <$> G.toGrid gridAxes world      (Main.hs:75)
 *                 == Aabb.culledKeys world                      (Broadphase/Aabb.hs:168-183)
 *   kContacts <- prepareFrame keys world                        (Main.hs:77, Solvers/Contact.hs:40-52)
 *   constraintGen beh dt fContact ab   -- per contact, inside applyCachedSlns
 *                                                               (Solvers/Contact.hs:93,107;
 *                                                                Constraints/Contact.hs:60-72)
 *
 * One shapes_frame() call returns all three results for one frame as
 * structure-of-arrays FP64/int32 buffers, in the reference's order
 * (descending ObjectFeatureKey).  The sequential solver, warm starting and
 * integration stay in the host engine.  INTEGRATION.md shows the
 * `foreign import ccall` binding a maintainer of the reference would add.
 *
 * Conventions: plain C, no CUDA or torch types.  Every function returns 0 on
 * success or a negative SHAPES_E_* code, never aborts, never throws.  A ctx is
 * single-caller and blocking; distinct ctxs may be used from distinct threads.
 * There is no CPU fallback: without a usable CUDA device every call fails with
 * SHAPES_E_CUDA.
 */
#ifndef SHAPES_B200_H
#define SHAPES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHAPES_OK          0
#define SHAPES_E_ARG      (-1) /* bad argument / call order */
#define SHAPES_E_CUDA     (-2) /* CUDA runtime error (text in shapes_last_error) */
#define SHAPES_E_NCCL     (-3) /* NCCL error */
#define SHAPES_E_CAPACITY (-4) /* max_pairs / max_contacts too small: required
                                  sizes are in the frame's n_pairs / n_contacts;
                                  nothing else was written; grow and retry */

#define SHAPES_NCCL_ID_BYTES 128

typedef struct shapes_ctx shapes_ctx;

/* Per-frame results.  Every array pointer is caller-owned and may be NULL
 * (= not wanted, nothing is copied).  Non-NULL pair arrays must hold
 * max_pairs elements and contact arrays max_contacts elements (the capacities
 * given to shapes_create).  Row k of every contact array describes the same
 * contact.
 *
 * For shapes_frame the pointers are HOST memory (pinned recommended:
 * shapes_host_alloc); for shapes_frame_device they are ignored and results
 * stay in HBM (shapes_device_view). */
typedef struct shapes_frame_out {
    /* Aabb.culledKeys / Grid.culledKeys: pairs (i, j), i > j, descending
     * lexicographic (Broadphase/Aabb.hs:155-183). */
    int64_t  n_pairs;
    int32_t *pair_i, *pair_j;

    /* prepareFrame: (ObjectFeatureKey (i,j) (featA,featB), Flipping Contact)
     * (Constraints/Contact.hs:36-57, Contact/Types.hs:28-35), descending. */
    int64_t  n_contacts;
    int32_t *key_i, *key_j, *feat_a, *feat_b;
    uint8_t *flip;                         /* 0 = Same, 1 = Flip (Utils/Utils.hs:147) */
    double  *normal_x, *normal_y;          /* _contactNormal */
    double  *center_x, *center_y;          /* _contactCenter */
    double  *depth;                        /* _contactDepth  */

    /* constraintGen: ContactConstraint (Constraints/Types.hs:41-51). */
    double  *j_np[6], *b_np;               /* _ccNonPen      (NonPenetration.hs:16-55) */
    double  *ra_x, *ra_y, *rb_x, *rb_y;    /* _ccRestitution radii (Restitution.hs:21-31) */
    double  *rn_x, *rn_y;                  /* _ccRestitution normal */
    double  *j_f[6], *b_f;                 /* _ccFriction    (Friction.hs:18-44); b_f is always 0 */
    /* effMassM2 of both constraints (Constraint.hs:173-179); velocity independent,
     * the reference recomputes it in every lagrangian2 call. Optional. */
    double  *inv_eff_np, *inv_eff_f;

    /* Warm start, the cache join of applyCachedSlns (Solvers/Contact.hs:84-121) = descZipVector
     * (Utils/Descending.hs:47-71): per contact, the ContactLagrangian (non-penetration, friction)
     * cached for the same ObjectFeatureKey in the previous frame, else 0 (newCache); warm_hit = 1
     * where useCache applies (the host then calls applySln for those rows).  Filled from the cache
     * given to shapes_set_lagrangian_cache; all zeros when none was given. Optional. */
    double  *warm_np, *warm_f;
    uint8_t *warm_hit;

    /* optional debug outputs, n_slots / n_verts elements */
    double  *aabb_min_x, *aabb_max_x, *aabb_min_y, *aabb_max_y; /* toAabb (Aabb.hs:81-110) */
    double  *world_x, *world_y;            /* _hullVertices after moveShapes (World.hs:136-140) */

    /* stats of this frame */
    int64_t  n_big;        /* shapes that took the big-shape path */
    int32_t  grid_w, grid_h;
    double   cell_size;
    float    device_ms;    /* kernels only, CUDA events on the ctx stream */
    float    total_ms;     /* device_ms + H2D + D2H (shapes_frame only) */
} shapes_frame_out;

/* Device-resident view of the last frame's results (pointers into ctx-owned
 * HBM, valid until the next frame / destroy).  Same meaning as above. */
typedef struct shapes_device_view {
    int64_t  n_pairs, n_contacts;
    const int32_t *pair_i, *pair_j;
    const int32_t *key_i, *key_j, *feat_a, *feat_b;
    const uint8_t *flip;
    const double  *normal_x, *normal_y, *center_x, *center_y, *depth;
    const double  *j_np[6], *b_np;
    const double  *ra_x, *ra_y, *rb_x, *rb_y, *rn_x, *rn_y;
    const double  *j_f[6];
    const double  *inv_eff_np, *inv_eff_f;
    const double  *warm_np, *warm_f;     /* NULL unless the frame ran the cache join */
    const uint8_t *warm_hit;
    const double  *aabb;   /* n_slots x (min_x, max_x, min_y, max_y) */
} shapes_device_view;

/* ---- lifetime ------------------------------------------------------- */

/* One ctx per GPU.  Owns all device memory, one stream and (world_size > 1)
 * one NCCL communicator.  max_* are capacities: slots, vertices, broadphase
 * pairs and contacts owned by THIS rank. */
int  shapes_create(shapes_ctx **out, int device_id,
                   int64_t max_shapes, int64_t max_verts,
                   int64_t max_pairs, int64_t max_contacts);

/* Rank `rank` of `world_size` cooperating ctxs (one process or thread per GPU
 * of one NVSwitch box).  nccl_id: SHAPES_NCCL_ID_BYTES from
 * shapes_nccl_unique_id() on rank 0, distributed by the host's own plumbing. */
int  shapes_create_ranked(shapes_ctx **out, int device_id, int rank, int world_size,
                          const void *nccl_id,
                          int64_t max_shapes, int64_t max_verts,
                          int64_t max_pairs, int64_t max_contacts);
int  shapes_nccl_unique_id(void *out_id /* SHAPES_NCCL_ID_BYTES */);
void shapes_destroy(shapes_ctx *);
const char *shapes_last_error(const shapes_ctx *);   /* NULL ctx: last create error */

/* ---- static geometry: once per world, and after append/delete -------- */

/* Mirrors World.append / listToHull (World.hs:77-84, ConvexHull.hs:151-167).
 * alive      n_slots  EmptiesVector filled flags (NULL = all filled)
 * vert_offset n_slots+1  CSR offsets into local_x / local_y
 * local_x/y  CCW local-space vertices (_hullLocalVertices)
 * ext_min/max per edge: _hullExtents (hull-relative vertex index); NULL => the
 *            library computes them with the listToHull rule.
 * Host pointers. Every rank registers the whole world. */
int  shapes_set_hulls(shapes_ctx *, int64_t n_slots, const uint8_t *alive,
                      const int32_t *vert_offset,
                      const double *local_x, const double *local_y,
                      const int32_t *ext_min, const int32_t *ext_max);

/* Same, for worlds that also contain CircleShapes (Contact.hs:19,
 * Contact/Circle.hs:16-22, Engine.makeCircle Engine.hs:53-54): radius[s] >= 0 makes slot s a
 * circle of that radius (its CSR vertex range must be empty), radius[s] < 0 a hull; radius == NULL
 * is shapes_set_hulls.  Frames then follow the full generateContacts dispatch (Contact.hs:22-40):
 * circle/circle (key (0,0), Same), circle/hull (key (0, hull feature), Same) and hull/circle
 * (key (hull feature, 0), Flip) through GJK closestSimplex, hull/hull through SAT. */
int  shapes_set_shapes(shapes_ctx *, int64_t n_slots, const uint8_t *alive,
                       const int32_t *vert_offset,
                       const double *local_x, const double *local_y,
                       const int32_t *ext_min, const int32_t *ext_max,
                       const double *radius);

/* Broadphase cell edge; <= 0 restores the automatic choice made by
 * shapes_set_hulls.  Performance only: results do not depend on it. */
int  shapes_set_cell_size(shapes_ctx *, double cell_size);

/* Grow max_pairs / max_contacts IN PLACE (values below the current capacity are ignored): the caller's answer to
 * SHAPES_E_CAPACITY.  A frame that returned SHAPES_E_CAPACITY left the ctx as it was before the call -- the previous
 * frame's ObjectFeatureKey columns, the Lagrangian cache supplied for them (the reference's EngineCache,
 * Engine/Main.hs:32,60-68) and an uploaded world all stay valid -- and shapes_grow keeps them, so the retried frame
 * (or world step) warm-starts exactly as the reference's applyCachedSlns would (Solvers/Contact.hs:84-121).
 * Only the result arrays of the last frame become unavailable (shapes_fetch / shapes_device_view_get fail until the
 * next completed frame).  Single-GPU ctxs only: ranks of a multi-GPU job re-create their ctxs collectively. */
int  shapes_grow(shapes_ctx *, int64_t max_pairs, int64_t max_contacts);

/* ---- per frame -------------------------------------------------------- */

/* Inputs are the reference's own SoA columns of _wPhysObjs (World.hs:47,
 * Constraint.hs:52-63): position, rotation, inverse masses.
 * cos_rot/sin_rot: cos/sin of rot as the host's libm computes them
 * (Linear.hs:353-357) for bit-exact parity; both NULL => the device computes
 * sincos(rot) (<= 2 ulp from libm, flagged not bit-exact).  rot may be NULL
 * when cos/sin are given.  static <=> inv_lin == 0 && inv_rot == 0
 * (Constraint.hs:123-125).  dt, baumgarte, slop: EngineConfig / ContactBehavior
 * (Engine/Main.hs:34-37, Contact/Types.hs:20-25).
 * Blocking: returns after the requested outputs are in `out`. */
int  shapes_frame(shapes_ctx *, int64_t n_slots,
                  const double *pos_x, const double *pos_y,
                  const double *rot, const double *cos_rot, const double *sin_rot,
                  const double *inv_lin, const double *inv_rot,
                  double dt, double baumgarte, double slop,
                  shapes_frame_out *out);

/* Same, with DEVICE input pointers (on the ctx's GPU); results stay in HBM.
 * out->n_* and stats are filled, array pointers in `out` are ignored. */
int  shapes_frame_device(shapes_ctx *, int64_t n_slots,
                         const double *pos_x, const double *pos_y,
                         const double *rot, const double *cos_rot, const double *sin_rot,
                         const double *inv_lin, const double *inv_rot,
                         double dt, double baumgarte, double slop,
                         shapes_frame_out *out);

int  shapes_device_view_get(shapes_ctx *, shapes_device_view *view);

/* Warm start.  The ctx keeps the ObjectFeatureKeys of the last
 * completed frame on the device.  Hand it the (key, ContactLagrangian) cache the host solver left
 * for that frame (EngineCache, Engine/Main.hs:32; row k <-> that frame's contact k; n_prev must
 * equal that frame's n_contacts) and the NEXT frame also produces warm_np / warm_f / warm_hit.
 * A cache is consumed by exactly one frame.  Host pointers / device pointers. */
int  shapes_set_lagrangian_cache(shapes_ctx *, int64_t n_prev, const double *lambda_np, const double *lambda_f);
int  shapes_set_lagrangian_cache_device(shapes_ctx *, int64_t n_prev, const double *lambda_np, const double *lambda_f);

/* Copy the last frame's results from HBM into the non-NULL arrays of `out`
 * (what shapes_frame does after the kernels). */
int  shapes_fetch(shapes_ctx *, shapes_frame_out *out);

/* ---- multi-GPU (world_size > 1) --------------------------------------- */

/* Rank r owns the pairs/contacts whose larger key i lies in
 * [own_lo, own_hi); the global descending order is rank world_size-1's rows,
 * then world_size-2's, ...  After a frame, all_pairs/all_contacts hold every
 * rank's counts (all-gathered), each world_size long. */
int  shapes_rank_info(shapes_ctx *, int64_t *own_lo, int64_t *own_hi,
                      int64_t *all_pairs, int64_t *all_contacts);

/* Rows mode (the default once the peers are mapped: shapes_ipc_import, or shapes_create_multi) splits the slot space
 * into 2 x world_size blocks and makes rank g the home of blocks g and 2 world_size - 1 - g: whatever the host's slot
 * numbering, every rank then holds the same number of slots and of pairs.  A rank's arrays hold the rows of its HIGH
 * block first (run 0), then those of its low block (run 1); the global descending order is run 0 of ranks 0, 1, ...,
 * world_size - 1 followed by run 1 of ranks world_size - 1, ..., 0.  For any rank of the job: the slot range and the
 * pair / contact counts of its two runs after the last frame (the other exchanges have one slot range per rank: run 0 is
 * empty, and the rule above reads "rank world_size - 1's rows, then world_size - 2's, ..."). */
int  shapes_rank_segments(shapes_ctx *, int rank, int64_t seg_lo[2], int64_t seg_hi[2],
                          int64_t seg_pairs[2], int64_t seg_contacts[2]);

/* Mapped peer memory (optional, replaces the NCCL all-gather of the AABB records): each rank exports
 * SHAPES_IPC_BYTES (CUDA IPC handles of its exchange arena); the host gathers the blobs of all ranks (rank order)
 * and hands them to every rank.  Ranks live in different processes (CUDA IPC does not map a process's own handles;
 * one process with several GPUs uses shapes_create_multi below).  With the peers mapped the frame runs in "rows
 * mode": homes push 4 B cell keys + 48 B body records over NVLink only to the ranks whose grid rows need them, the
 * sweep / SAT work is split by grid rows balanced on the previous frame's pair counts, every pair is stored straight
 * into its final place at the rank that owns its larger key, and per-frame flag words are the cross-GPU barriers
 * (five per frame, bounded spins).  shapes_frame then uploads only the rank's own slots of the body columns.
 * Without an import the exchange goes through NCCL.  SHAPES_B200_NO_P2P=1 forces the NCCL path,
 * SHAPES_B200_NO_ROWS=1 the round-1 exchange (keys pushed to every peer, AABB records pulled, slot-range ownership
 * of the whole path). */
#define SHAPES_IPC_BYTES 2048
int  shapes_ipc_export(shapes_ctx *, void *out_blob /* SHAPES_IPC_BYTES */);
int  shapes_ipc_import(shapes_ctx *, const void *all_blobs /* world_size x SHAPES_IPC_BYTES */);

/* ---- one process, several GPUs ----------------------------------------------------------------
 *
 * The reference's host is one single-threaded ST computation (Engine/Main.hs:38,71-86): it cannot run one process
 * per GPU.  shapes_create_multi builds one ctx per listed GPU of an NVSwitch box inside THIS process (peer access
 * instead of CUDA IPC, no NCCL), and shapes_multi_frame is shapes_frame over all of them from one host thread:
 * every GPU uploads only its slot range of the body columns over its own PCIe link, the sweep / SAT work is split by
 * grid rows balanced on the pair counts of the previous frame, every pair is delivered to the GPU that owns its
 * larger key, and each GPU's slice of the result is copied straight to its global row offset in `out` -- which
 * therefore holds the whole frame in the reference's descending order, exactly what a single-GPU shapes_frame
 * returns.  Array pointers in `out` must hold n_gpus x max_pairs_per_gpu pairs / n_gpus x max_contacts_per_gpu
 * rows.  The aabb / world debug outputs are not gathered (use shapes_multi_rank + shapes_fetch). */
typedef struct shapes_multi shapes_multi;
int  shapes_create_multi(shapes_multi **out, int n_gpus, const int *device_ids /* NULL: 0..n_gpus-1 */,
                         int64_t max_shapes, int64_t max_verts,
                         int64_t max_pairs_per_gpu, int64_t max_contacts_per_gpu);
void shapes_multi_destroy(shapes_multi *);
const char *shapes_multi_last_error(const shapes_multi *);   /* NULL: last create error */
int  shapes_multi_set_shapes(shapes_multi *, int64_t n_slots, const uint8_t *alive,
                             const int32_t *vert_offset,
                             const double *local_x, const double *local_y,
                             const int32_t *ext_min, const int32_t *ext_max,
                             const double *radius);
int  shapes_multi_frame(shapes_multi *, int64_t n_slots,
                        const double *pos_x, const double *pos_y,
                        const double *rot, const double *cos_rot, const double *sin_rot,
                        const double *inv_lin, const double *inv_rot,
                        double dt, double baumgarte, double slop,
                        shapes_frame_out *out);
/* The ctx of one GPU (stage timing, shapes_rank_info, shapes_fetch of its slice, ...). */
shapes_ctx *shapes_multi_rank(shapes_multi *, int rank);

/* ---- device-resident world ---------------------------- */

/* The rest of Physics.Engine.Main.updateWorld (Engine/Main.hs:71-86) on the GPU: the body state
 * (_wPhysObjs, _wMaterials; World.hs:46-52) lives in HBM, and one shapes_world_step call runs
 *   culledKeys -> applyExternal -> prepareFrame -> applyCachedSlns -> improveWorld x iterations
 *   -> advance -> moveShapes
 * without any per-frame host<->device traffic.  The solver reproduces the reference's SEQUENTIAL
 * Gauss-Seidel walk bit for bit: contacts are executed as the dependency graph that walk defines
 * (a contact waits only for the previous contact on either of its two bodies), so every body sees
 * its velocity updates in the reference's order.  Opt-in: a host that keeps its own solver simply
 * never calls these functions.  Single-GPU ctxs only.
 *
 * Rotation: moveShapes needs cos/sin of the advanced rotation; the device uses shapes_sincos
 * (include/shapes_sincos.h), a plain-IEEE routine the host can call too (same bits on both sides,
 * <= 1 ulp from libm).  A step sequence is therefore bit-identical to the reference's updateWorld
 * evaluated with shapes_sincos in place of libm's cos/sin in rotate22 (Linear.hs:353-357). */

#define SHAPES_EXT_NONE  0
#define SHAPES_EXT_ACCEL 1   /* constantAccel (World/External.hs:23-27) */
#define SHAPES_EXT_FORCE 2   /* constantForce (World/External.hs:16-20), as the reference parses it */

typedef struct shapes_step_config {
    double dt;                 /* _engineTimestep (Engine/Main.hs:34-37) */
    double baumgarte, slop;    /* ContactBehavior (Contact/Types.hs:20-25) */
    int32_t external_kind;     /* SHAPES_EXT_* : the External applied each frame (World.hs:32-33) */
    int32_t solver_iterations; /* improveWorld sweeps; updateWorld runs 2 (Engine/Main.hs:82-83) */
    double external_x, external_y;
    int32_t warm_start;        /* 1 = applyCachedSlns against the previous step's cache (updateWorld);
                                  0 = every contact starts from ContactLagrangian 0 0 */
    int32_t reserved;
} shapes_step_config;

typedef struct shapes_step_stats {
    int64_t n_pairs, n_contacts;
    int64_t solver_nodes;      /* (sweep, contact pair) nodes executed */
    int64_t queue_pushes;      /* nodes that went through the ready queue instead of being chained */
    int64_t body_chains;       /* dynamic bodies with at least one contact */
    int32_t warm;              /* 1 = the cache join ran (sweep 0 applied cached Lagrangians) */
    int32_t reserved;
    float   frame_ms;          /* the hot path (K0..K3 + join), device time */
    float   chains_ms;         /* external + dependency chains (sort + links) */
    float   solve_ms;          /* k_solve */
    float   integrate_ms;      /* advance + cache hand-over */
    float   total_ms;          /* whole step, device time */
} shapes_step_stats;

/* Host columns of the n_slots objects: velocity, position, rotation (PhysicalObj, Constraint.hs:52-63),
 * inverse masses, materials (mu = _mMu, bounce = _mBounce, World.hs:36-40).  cos_rot / sin_rot: the
 * rotation the shapes were last moved with (makeWorldObj's moveShape); both NULL = shapes_sincos(rot).
 * Requires shapes_set_hulls / shapes_set_shapes.  Resets the EngineCache (initEngine). */
int  shapes_world_upload(shapes_ctx *, int64_t n_slots,
                         const double *vel_x, const double *vel_y, const double *rot_vel,
                         const double *pos_x, const double *pos_y, const double *rot,
                         const double *cos_rot, const double *sin_rot,
                         const double *inv_lin, const double *inv_rot,
                         const double *mu, const double *bounce);
/* Copy the body state back; any pointer may be NULL. */
int  shapes_world_download(shapes_ctx *, int64_t n_slots,
                           double *vel_x, double *vel_y, double *rot_vel,
                           double *pos_x, double *pos_y, double *rot,
                           double *cos_rot, double *sin_rot);
/* One updateWorld.  SHAPES_E_CAPACITY (required counts in stats->n_pairs / n_contacts) leaves the
 * world untouched.  Afterwards shapes_fetch / shapes_device_view_get describe the frame the step
 * generated its contacts from; warm_hit is the cache join, warm_np / warm_f are the Lagrangians
 * the solver LEFT (= the EngineCache the next step joins against). */
int  shapes_world_step(shapes_ctx *, const shapes_step_config *cfg, shapes_step_stats *stats /* may be NULL */);

/* Host-side evaluation of the shared cos/sin (include/shapes_sincos.h). Needs no GPU. */
void shapes_sincos(int64_t n, const double *rot, double *cos_out, double *sin_out);

/* ---- helpers ----------------------------------------------------------- */

void *shapes_host_alloc(size_t bytes);     /* pinned host memory, NULL on failure */
void  shapes_host_free(void *);
/* CUDA stream handle (cudaStream_t) the ctx launches on, for event timing. */
void *shapes_stream(shapes_ctx *);
/* Number of kernels the library launched on the ctx stream so far. */
int64_t shapes_launch_count(const shapes_ctx *);
const char *shapes_version(void);

/* Statistics of the last completed frame (bench / profiling; nothing here changes results). */
#define SHAPES_SAT_PER_THREAD_BOXES   0  /* k_manifolds<4>: boxes-only worlds, one thread per pair            */
#define SHAPES_SAT_PER_THREAD         1  /* k_manifolds<8>: one thread per pair                               */
#define SHAPES_SAT_PER_THREAD_CIRCLES 2  /* k_manifolds<8, circles>: the full generateContacts dispatch       */
#define SHAPES_SAT_COOP               3  /* k_manifolds_coop: 16 lanes per pair (general polygon worlds)      */
typedef struct shapes_frame_info {
    int64_t pairs_with_contacts;   /* broadphase pairs whose narrow phase produced at least one contact */
    int32_t sorted_mode;           /* 1 = hull records and the SAT work list are kept in grid-cell order */
    int32_t sat_kernel;            /* SHAPES_SAT_* */
} shapes_frame_info;
int  shapes_last_frame_info(shapes_ctx *, shapes_frame_info *info);

/* Per-stage device timing (CUDA events on the ctx stream between the stages of a frame).
 * Off by default; when on, shapes_stage_ms fills SHAPES_N_STAGES milliseconds of the last
 * frame, in the order shapes_stage_name reports. */
#define SHAPES_N_STAGES 12
int  shapes_set_profiling(shapes_ctx *, int enabled);
int  shapes_stage_ms(const shapes_ctx *, float *out_ms /* SHAPES_N_STAGES */);
const char *shapes_stage_name(int stage);

#ifdef __cplusplus
}
#endif
#endif /* SHAPES_B200_H */

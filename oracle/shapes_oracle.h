/*
 * shapes_oracle.h -- CPU restatement of the ublubu/shapes collision hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or
 * the reported CPU baseline.  The product (shapes_b200/csrc) never links,
 * includes or calls it.
 *
 * PARITY STATUS: *parity unpinned by reference outputs*.  The reference is
 * Haskell; no GHC/stack/cabal exists in this environment, so the reference
 * could not be run and holds no golden outputs for SAT / clipping / the
 * constraint generators.  This restatement follows the
 * cited reference lines operation for operation and is pinned only by
 *   - the two properties the reference's own tests hold
 *     (shapes/test/Physics/Broadphase/AabbSpec.hs:8-11,
 *      shapes-math/test/Shapes/Linear/TemplateSpec.hs:25-35), and
 *   - hand-derived known-answer vectors for the reference's bench fixtures
 *     (tests/golden/, derivations in tests/golden/README.md).
 *
 * All reals are IEEE binary64, no FMA contraction (-ffp-contract=off), no
 * reassociation.  All paths below are relative to /root/reference/.
 */
#ifndef SHAPES_ORACLE_H
#define SHAPES_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* World geometry in the flattened (CSR) form the boundary uses.
 * Slot s owns vertices [vert_offset[s], vert_offset[s+1]); alive[s]==0 means
 * an empty EmptiesVector slot (shapes/src/Utils/EmptiesVector.hs:136-161). */

/* listToHull's _hullExtents (shapes/src/Physics/Contact/ConvexHull.hs:151-167):
 * per edge e of every hull, the (argmin, argmax) vertex index (hull-relative)
 * along the LOCAL unit edge normal, first-min/first-max on ties (:81-92). */
void orc_hull_extents(int64_t n_slots, const int32_t *vert_offset,
                      const double *local_x, const double *local_y,
                      int32_t *ext_min, int32_t *ext_max);

/* cos/sin exactly as rotate22 obtains them: libm through cosDouble#/sinDouble#
 * (shapes/src/Physics/Linear.hs:353-357). */
void orc_cos_sin(int64_t n, const double *rot, double *cos_out, double *sin_out);

/* moveShapes (shapes/src/Physics/World.hs:132-140) -> setHullTransform
 * (ConvexHull.hs:184-195): world vertices and recomputed unit edge normals. */
void orc_move_shapes(int64_t n_slots, const uint8_t *alive, const int32_t *vert_offset,
                     const double *local_x, const double *local_y,
                     const double *pos_x, const double *pos_y,
                     const double *cos_rot, const double *sin_rot,
                     double *world_x, double *world_y,
                     double *normal_x, double *normal_y);

/* hullToAabb (shapes/src/Physics/Broadphase/Aabb.hs:81-110). */
void orc_aabbs(int64_t n_slots, const uint8_t *alive, const int32_t *vert_offset,
               const double *world_x, const double *world_y,
               double *min_x, double *max_x, double *min_y, double *max_y);

/* isStatic (shapes/src/Physics/Constraint.hs:123-125). */
void orc_is_static(int64_t n_slots, const double *inv_lin, const double *inv_rot,
                   uint8_t *is_static);

/* Aabb.culledKeys (Aabb.hs:168-183): Theta(n^2) enumeration in unorderedPairs
 * order (:155-163).  Returns the number of pairs found; writes at most `cap`. */
int64_t orc_culled_keys_aabb(int64_t n_slots, const uint8_t *alive,
                             const double *min_x, const double *max_x,
                             const double *min_y, const double *max_y,
                             const uint8_t *is_static,
                             int64_t cap, int32_t *pair_i, int32_t *pair_j);

/* Grid.toGrid + Grid.culledKeys (shapes/src/Physics/Broadphase/Grid.hs:67-141)
 * -- the variant updateWorld calls (Engine/Main.hs:75). */
int64_t orc_culled_keys_grid(int64_t n_slots, const uint8_t *alive,
                             const double *min_x, const double *max_x,
                             const double *min_y, const double *max_y,
                             const uint8_t *is_static,
                             int32_t grid_len_x, double grid_unit_x, double grid_origin_x,
                             int32_t grid_len_y, double grid_unit_y, double grid_origin_y,
                             int64_t cap, int32_t *pair_i, int32_t *pair_j);

/* Same pair set and order as orc_culled_keys_aabb, found by an x-sorted sweep
 * with the identical predicate; for worlds where n^2 is infeasible.  Not a
 * restatement of reference code -- validated against orc_culled_keys_aabb. */
int64_t orc_culled_keys_sweep(int64_t n_slots, const uint8_t *alive,
                              const double *min_x, const double *max_x,
                              const double *min_y, const double *max_y,
                              const uint8_t *is_static,
                              int64_t cap, int32_t *pair_i, int32_t *pair_j);

/* unorderedPairs n (Aabb.hs:155-163); returns the count, writes at most cap. */
int64_t orc_unordered_pairs(int64_t n, int64_t cap, int32_t *xs, int32_t *ys);

/* Output rows of prepareFrame + constraintGen, one row per contact, in the
 * reference's descending ObjectFeatureKey order.  Any pointer may be NULL. */
typedef struct orc_contacts_out {
    int64_t cap;            /* capacity of every non-NULL array (rows) */
    int32_t *key_i, *key_j; /* _ofkObjKeys  (Constraints/Contact.hs:36-39) */
    int32_t *feat_a, *feat_b; /* _ofkFeatKeys */
    uint8_t *flip;          /* 0 = Same, 1 = Flip (Utils/Utils.hs:147) */
    double *normal_x, *normal_y, *center_x, *center_y, *depth; /* Contact/Types.hs:28-35 */
    double *j_np[6], *b_np; /* NonPenetration constraint (NonPenetration.hs:16-55) */
    double *ra_x, *ra_y, *rb_x, *rb_y, *rn_x, *rn_y; /* Restitution.hs:21-31 */
    double *j_f[6], *b_f;   /* Friction constraint (Friction.hs:18-44) */
    double *inv_eff_np, *inv_eff_f; /* effMassM2 (Constraint.hs:173-179) */
} orc_contacts_out;

/* prepareFrame (Solvers/Contact.hs:40-52) over the given pairs followed by
 * constraintGen (Constraints/Contact.hs:60-72) per contact.
 * Returns the number of contacts (which may exceed out->cap; rows beyond cap
 * are not written). */
int64_t orc_contacts(int64_t n_pairs, const int32_t *pair_i, const int32_t *pair_j,
                     const int32_t *vert_offset,
                     const double *world_x, const double *world_y,
                     const double *normal_x, const double *normal_y,
                     const int32_t *ext_min, const int32_t *ext_max,
                     const double *pos_x, const double *pos_y,
                     const double *inv_lin, const double *inv_rot,
                     double dt, double baumgarte, double slop,
                     orc_contacts_out *out);

/* ---- circles ------------------------------------------------
 * A slot with radius[s] >= 0 is a CircleShape of that radius (its CSR vertex range is empty);
 * radius == NULL or radius[s] < 0 means HullShape.
 *
 * orc_move_circles: setCircleTransform (shapes/src/Physics/Contact/Circle.hs:55-59): the world
 * centre is the transform applied to the local origin (afmul with the same toTransform matrix).
 * orc_aabbs_circles: circleToAabb (shapes/src/Physics/Broadphase/Aabb.hs:86-88), overwrites the
 * AABB entries of the circle slots.
 * orc_contacts_shapes: prepareFrame + constraintGen with the full dispatch of
 * Physics.Contact.generateContacts (shapes/src/Physics/Contact.hs:22-40): circle/circle
 * (Circle.hs:30-53), circle/hull and hull/circle through GJK closestSimplex
 * (Contact/GJK.hs:52-146, Contact/CircleVsHull.hs:18-69), hull/hull as orc_contacts. */
void orc_move_circles(int64_t n_slots, const uint8_t *alive, const double *radius,
                      const double *pos_x, const double *pos_y,
                      const double *cos_rot, const double *sin_rot,
                      double *center_x, double *center_y);
void orc_aabbs_circles(int64_t n_slots, const uint8_t *alive, const double *radius,
                       const double *center_x, const double *center_y,
                       double *min_x, double *max_x, double *min_y, double *max_y);
int64_t orc_contacts_shapes(int64_t n_pairs, const int32_t *pair_i, const int32_t *pair_j,
                            const int32_t *vert_offset,
                            const double *world_x, const double *world_y,
                            const double *normal_x, const double *normal_y,
                            const int32_t *ext_min, const int32_t *ext_max,
                            const double *radius, const double *circle_x, const double *circle_y,
                            const double *pos_x, const double *pos_y,
                            const double *inv_lin, const double *inv_rot,
                            double dt, double baumgarte, double slop,
                            orc_contacts_out *out);

/* The cache join of applyCachedSlns (shapes/src/Physics/Solvers/Contact.hs:84-121): descZipVector
 * (shapes/src/Utils/Descending.hs:47-71) walks this frame's contacts (descending ObjectFeatureKey)
 * against the previous frame's (key, ContactLagrangian) cache (descending): keys equal => useCache
 * (the cached Lagrangians are propagated, hit = 1), otherwise newCache (ContactLagrangian 0 0).
 * Keys are (i, j, featA, featB) compared lexicographically (derived Ord, Constraints/Contact.hs:36-39).
 * The velocity update of useCache (applySln) is host-side and not restated. */
void orc_warm_join(int64_t n_this, const int32_t *ti, const int32_t *tj, const int32_t *tfa, const int32_t *tfb,
                   int64_t n_that, const int32_t *pi, const int32_t *pj, const int32_t *pfa, const int32_t *pfb,
                   const double *that_np, const double *that_f,
                   double *out_np, double *out_f, uint8_t *out_hit);

/* dotV2 / mul2x2x2 as the TH templates generate them
 * (shapes-math/src/Shapes/Linear/Template.hs:108-110, MatrixTemplate.hs:47-67);
 * exported so the TemplateSpec property can be restated. */
double orc_dot_v2(double ax, double ay, double bx, double by);
void orc_mul2x2x2(const double *a, const double *b, double *out);

/* solveConstraint (Constraint.hs:164-204) for the one known-answer vector of
 * shapes/bench/Physics/Constraint/Benchmark.hs.  vel6 in/out = (va, wa, vb, wb). */
void orc_solve_constraint(const double *j6, double b, const double *inv_mass6,
                          double *vel6);

/* ---- the rest of updateWorld: shapes_oracle_step.c -------
 * applyExternal (World.hs:156-158) with constantAccel (kind 1) / constantForce (kind 2)
 * (World/External.hs:16-28); kind 0 = no external. */
void orc_apply_external(int64_t n_slots, const uint8_t *alive, int kind, double ex, double ey, double dt,
                        const double *inv_lin, double *vel_x, double *vel_y);
/* advance (World.hs:167-169, advanceObj Constraint.hs:225-229). */
void orc_advance(int64_t n_slots, const uint8_t *alive, double dt,
                 const double *vel_x, const double *vel_y, const double *rot_vel,
                 double *pos_x, double *pos_y, double *rot);
/* useCache's applySln over the rows the warm join hit (Solvers/Contact.hs:54-65,99-112). */
void orc_apply_cached(int64_t n_contacts, const int32_t *key_i, const int32_t *key_j, const uint8_t *hit,
                      const double *lam_np, const double *lam_f,
                      const double *const j_np[6], const double *const j_f[6],
                      const double *inv_lin, const double *inv_rot,
                      double *vel_x, double *vel_y, double *rot_vel);
/* One improveWorld solutionProcessor sweep (Solvers/Contact.hs:124-157, Constraints/Contact.hs:74-110,
 * SolutionProcessors.hs:12-53, Restitution.hs:34-47); lam_np / lam_f are the EngineCache values,
 * updated in place. */
void orc_improve_world(int64_t n_contacts, const int32_t *key_i, const int32_t *key_j,
                       const double *const j_np[6], const double *b_np,
                       const double *ra_x, const double *ra_y, const double *rb_x, const double *rb_y,
                       const double *rn_x, const double *rn_y,
                       const double *const j_f[6],
                       const double *mu, const double *bounce,
                       const double *inv_lin, const double *inv_rot,
                       double *vel_x, double *vel_y, double *rot_vel,
                       double *lam_np, double *lam_f);

#ifdef __cplusplus
}
#endif
#endif

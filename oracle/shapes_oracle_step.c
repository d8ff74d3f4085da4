/*
 * shapes_oracle_step.c -- CPU restatement of the rest of Physics.Engine.Main.updateWorld
 * (shapes/src/Physics/Engine/Main.hs:71-86) around the collision hot path: applyExternal, the
 * velocity part of applyCachedSlns, improveWorld with solutionProcessor, advance.
 * 
 *
 * TEST INFRASTRUCTURE ONLY (see shapes_oracle.h).  PARITY UNPINNED by reference outputs except
 * KAT-2 (solveConstraint on shapes/bench/Physics/Constraint/Benchmark.hs:11-35, checked through
 * orc_solve_constraint and orc_improve_world in tests/test_oracle_kat.py).
 *
 * The loops below are strictly sequential in the reference's order (descending ObjectFeatureKey =
 * row order); every floating-point expression is parenthesised as the reference evaluates it.
 * Paths are relative to /root/reference/.
 */
#include "shapes_oracle.h"

#include <math.h>

/* Ord Double has no min/max of its own in GHC 8.2 (ghc-prim GHC/Classes.hs): the class defaults
 * apply, `max x y = if x <= y then y else x`, `min x y = if x <= y then x else y`. */
static inline double hs_max(double x, double y) { return (x <= y) ? y : x; }
static inline double hs_min(double x, double y) { return (x <= y) ? x : y; }

/* dotV6 (shapes-math/src/Shapes/Linear/Template.hs:108-110): foldl1 (+) of the products */
static inline double dot6(const double *a, const double *b)
{
    return (((((a[0] * b[0]) + (a[1] * b[1])) + (a[2] * b[2])) + (a[3] * b[3])) + (a[4] * b[4])) + (a[5] * b[5]);
}

/* applyLagrangian (Constraint.hs:216-222) -> applyLagrangian2 (:197-203) -> updateVelocity2_
 * (:188-193) with constraintImpulse2 (:181-185): v + (im `vmulDiag6'` (l `smulV6` j)), i.e.
 * v_k + ((j_k * l) * im_k)  (smulV6 Linear.hs:84-86, vmulDiag6' :145-147, plusV6 :104-106). */
static inline void apply_lagrangian(double l, const double *j, const double *im, double *v)
{
    for (int k = 0; k < 6; ++k) v[k] = v[k] + ((j[k] * l) * im[k]);
}

/* effMassM2 (Constraint.hs:173-179): (j `vmulDiag6` im) `dotV6` j */
static inline double eff_mass(const double *j, const double *im)
{
    double t[6];
    for (int k = 0; k < 6; ++k) t[k] = j[k] * im[k];
    return dot6(t, j);
}

/* lagrangian2 (Constraint.hs:164-169): (-(j . v + b)) / mc */
static inline double lagrangian2(const double *j, double b, const double *im, const double *v)
{
    return (-(dot6(j, v) + b)) / eff_mass(j, im);
}

/* applyExternal (World.hs:156-158) over the filled slots with one of the two externals the
 * reference defines (World/External.hs:16-28):
 *   kind 1  constantAccel a: v + a*dt unless isStaticLin (inv_lin == 0)      (:23-27)
 *   kind 2  constantForce f: (v + f*dt) * inv_lin -- the reference's own parse of
 *           "v `plusV2` (f `smulV2'` dt) `smulV2'` im" (all backtick operators are infixl 9)  (:16-20)
 * smulV2' v s = each component * s (Linear.hs:72-78). */
void orc_apply_external(int64_t n_slots, const uint8_t *alive, int kind, double ex, double ey, double dt,
                        const double *inv_lin, double *vel_x, double *vel_y)
{
    for (int64_t s = 0; s < n_slots; ++s) {
        if (alive && !alive[s]) continue;
        if (kind == 1) {
            if (0.0 == inv_lin[s]) continue;
            vel_x[s] = vel_x[s] + (ex * dt);
            vel_y[s] = vel_y[s] + (ey * dt);
        } else if (kind == 2) {
            vel_x[s] = (vel_x[s] + (ex * dt)) * inv_lin[s];
            vel_y[s] = (vel_y[s] + (ey * dt)) * inv_lin[s];
        }
    }
}

/* advance (World.hs:167-169) -> advanceObj (Constraint.hs:225-229):
 * pos' = (dt `smulV2` vel) `plusV2` pos = (vel*dt) + pos;  rot' = (dt * rotVel) + rot. */
void orc_advance(int64_t n_slots, const uint8_t *alive, double dt,
                 const double *vel_x, const double *vel_y, const double *rot_vel,
                 double *pos_x, double *pos_y, double *rot)
{
    for (int64_t s = 0; s < n_slots; ++s) {
        if (alive && !alive[s]) continue;
        pos_x[s] = (vel_x[s] * dt) + pos_x[s];
        pos_y[s] = (vel_y[s] * dt) + pos_y[s];
        rot[s] = (dt * rot_vel[s]) + rot[s];
    }
}

static inline void load_pair(int32_t i, int32_t j, const double *vel_x, const double *vel_y, const double *rot_vel,
                             const double *inv_lin, const double *inv_rot, double *v, double *im)
{
    /* _constrainedVel6 (Constraint.hs:104-105), invMassM2 (:118-120) */
    v[0] = vel_x[i]; v[1] = vel_y[i]; v[2] = rot_vel[i];
    v[3] = vel_x[j]; v[4] = vel_y[j]; v[5] = rot_vel[j];
    im[0] = inv_lin[i]; im[1] = inv_lin[i]; im[2] = inv_rot[i];
    im[3] = inv_lin[j]; im[4] = inv_lin[j]; im[5] = inv_rot[j];
}

static inline void store_pair(int32_t i, int32_t j, const double *v, double *vel_x, double *vel_y, double *rot_vel)
{
    vel_x[i] = v[0]; vel_y[i] = v[1]; rot_vel[i] = v[2];
    vel_x[j] = v[3]; vel_y[j] = v[4]; rot_vel[j] = v[5];
}

/* The velocity side of applyCachedSlns (Solvers/Contact.hs:84-121): walking the contacts in order,
 * a contact whose key was cached (hit) applies the cached ContactLagrangian with applySln
 * (useCache :99-112; applySln :54-65 = applyFriction . applyNonPen).  The join itself is
 * orc_warm_join; constraintGen is position-only, so the rows computed up front are the rows
 * useCache would compute. */
void orc_apply_cached(int64_t n_contacts, const int32_t *key_i, const int32_t *key_j, const uint8_t *hit,
                      const double *lam_np, const double *lam_f,
                      const double *const j_np[6], const double *const j_f[6],
                      const double *inv_lin, const double *inv_rot,
                      double *vel_x, double *vel_y, double *rot_vel)
{
    for (int64_t r = 0; r < n_contacts; ++r) {
        if (!hit[r]) continue;
        double v[6], im[6], jn[6], jf[6];
        load_pair(key_i[r], key_j[r], vel_x, vel_y, rot_vel, inv_lin, inv_rot, v, im);
        for (int k = 0; k < 6; ++k) { jn[k] = j_np[k][r]; jf[k] = j_f[k][r]; }
        apply_lagrangian(lam_np[r], jn, im, v);
        apply_lagrangian(lam_f[r], jf, im, v);
        store_pair(key_i[r], key_j[r], v, vel_x, vel_y, rot_vel);
    }
}

/* One improveWorld sweep (Solvers/Contact.hs:146-157): improveContactSln (:124-143) per contact in
 * order.  Per contact, with (a, b) = (obj_i, obj_j) read ONCE before either constraint is solved:
 *   bounceB (Restitution.hs:34-47): min 0 (bounciness * (closingVelocity . rn)), bounciness =
 *     uncurry min (bounce_i, bounce_j), closingVelocity = ((-va + (-wa) x ra) + vb) + wb x rb,
 *     zcrossV2 z (x, y) = (-(z*y), z*x) (Linear.hs:127-130)
 *   nonPenWithRestitution (Constraints/Contact.hs:74-85): b = b_np + bounceB
 *   contactLagrangian (:87-97): lagrangian2 for non-penetration and friction (b_f = 0)
 *   solutionProcessor (:99-110): NonPenetration = positive (SolutionProcessors.hs:28-35):
 *     apply = max new (-cached), cache' = cached + apply; Friction = clampAbs (:37-53) with
 *     maxThresh = cache_np' * pairMu, pairMu (ua, ub) = (ua + ub) / 2 (Friction.hs:46-55)
 *   applySln toApply (non-penetration first, then friction), cache written back.
 * Materials are per object: mu, bounce (World.hs:36-40 Material). */
void orc_improve_world(int64_t n_contacts, const int32_t *key_i, const int32_t *key_j,
                       const double *const j_np[6], const double *b_np,
                       const double *ra_x, const double *ra_y, const double *rb_x, const double *rb_y,
                       const double *rn_x, const double *rn_y,
                       const double *const j_f[6],
                       const double *mu, const double *bounce,
                       const double *inv_lin, const double *inv_rot,
                       double *vel_x, double *vel_y, double *rot_vel,
                       double *lam_np, double *lam_f)
{
    for (int64_t r = 0; r < n_contacts; ++r) {
        const int32_t i = key_i[r], j = key_j[r];
        double v[6], im[6], jn[6], jf[6];
        load_pair(i, j, vel_x, vel_y, rot_vel, inv_lin, inv_rot, v, im);
        for (int k = 0; k < 6; ++k) { jn[k] = j_np[k][r]; jf[k] = j_f[k][r]; }
        /* bounceB */
        const double bounciness = hs_min(bounce[i], bounce[j]);
        const double nwa = -v[2];
        const double nwa_x = -(nwa * ra_y[r]), nwa_y = nwa * ra_x[r];
        const double wb_x = -(v[5] * rb_y[r]), wb_y = v[5] * rb_x[r];
        const double cv_x = (((-v[0]) + nwa_x) + v[3]) + wb_x;
        const double cv_y = (((-v[1]) + nwa_y) + v[4]) + wb_y;
        const double bounce_b = hs_min(0.0, bounciness * ((cv_x * rn_x[r]) + (cv_y * rn_y[r])));
        /* contactLagrangian */
        const double new_np = lagrangian2(jn, b_np[r] + bounce_b, im, v);
        const double new_f = lagrangian2(jf, 0.0, im, v);
        /* solutionProcessor */
        const double cached_np = lam_np[r], cached_f = lam_f[r];
        const double apply_np = hs_max(new_np, -cached_np);
        const double cache_np = cached_np + apply_np;
        const double max_thresh = cache_np * ((mu[i] + mu[j]) / 2.0);
        const double min_thresh = -max_thresh;
        const double accum = cached_f + new_f;
        const double accum2 = (accum > max_thresh) ? max_thresh : ((accum < min_thresh) ? min_thresh : accum);
        const double apply_f = accum2 - cached_f;
        /* applySln */
        apply_lagrangian(apply_np, jn, im, v);
        apply_lagrangian(apply_f, jf, im, v);
        store_pair(i, j, v, vel_x, vel_y, rot_vel);
        lam_np[r] = cache_np;
        lam_f[r] = accum2;
    }
}

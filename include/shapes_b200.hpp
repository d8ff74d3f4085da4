// shapes_b200.hpp -- C++17 host-side mirror of the reference's contact path over the C ABI.
//
// The reference engine is compiled Haskell; GHC is not available in this build environment, so
// the host layer above include/shapes_b200.h is provided in C++ (header only) with the reference's
// names and argument meaning:
//
//   shapes::World            Physics.World.World          (shapes/src/Physics/World.hs:46-84)
//   shapes::makeRectangleHull / makeHull / makeCircle      (shapes/src/Physics/Engine.hs:47-54)
//   shapes::makePhysicalObj  + toInvMass2                 (Engine.hs:32-39, Constraint.hs:79-83)
//   shapes::culledKeys       Aabb.culledKeys / Grid.culledKeys   (Broadphase/Aabb.hs:168-183)
//   shapes::updateWorld      Physics.Engine.Main.updateWorld     (Engine/Main.hs:71-86), on device-resident bodies
//   shapes::prepareFrame     Solvers.Contact.prepareFrame        (Solvers/Contact.hs:40-52)
//   shapes::constraintGen    Constraints.Contact.constraintGen   (Constraints/Contact.hs:60-72)
//
// Errors: the replaced functions are total; every non-zero C-ABI code is thrown as shapes::Error
// (SHAPES_E_CAPACITY is handled by shapes_grow and retrying, the failed frame being side-effect free).
// There is no CPU fallback: without the CUDA library/device the constructor throws.
#pragma once

#include "shapes_b200.h"

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace shapes {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &msg) : std::runtime_error("shapes_b200 error " + std::to_string(c) + ": " + msg), code(c) {}
};

struct V2 { double x, y; };

// ContactBehavior (shapes/src/Physics/Contact/Types.hs:20-25)
struct ContactBehavior { double contactBaumgarte = 0.0, contactPenetrationSlop = 0.0; };

// PhysicalObj, position part + inverse mass (Constraint.hs:52-63); velocities stay with the host solver
struct PhysicalObj { V2 pos; double rotPos; double invLin, invRot; V2 vel{ 0.0, 0.0 }; double rotVel = 0.0; };

// toInvMass2 (Constraint.hs:79-83): mass 0 means infinite mass
inline PhysicalObj makePhysicalObj(V2 pos, double rotPos, std::pair<double, double> mass)
{
    auto inv = [](double m) { return m == 0.0 ? 0.0 : 1.0 / m; };
    return PhysicalObj{ pos, rotPos, inv(mass.first), inv(mass.second) };
}
// makePhysicalObj vel rotvel pos rotpos mass (Engine.hs:32-39), the reference's argument order
inline PhysicalObj makePhysicalObj(V2 vel, double rotVel, V2 pos, double rotPos, std::pair<double, double> mass)
{
    PhysicalObj o = makePhysicalObj(pos, rotPos, mass);
    o.vel = vel; o.rotVel = rotVel;
    return o;
}

using Hull = std::vector<V2>;    // CCW local vertices (_hullLocalVertices)

// rectangleVertices (ConvexHull.hs:135-145)
inline Hull makeRectangleHull(double w, double h)
{
    const double w2 = w / 2.0, h2 = h / 2.0;
    return Hull{ { w2, h2 }, { -w2, h2 }, { -w2, -h2 }, { w2, -h2 } };
}
inline Hull makeHull(std::vector<V2> vertices) { return vertices; }
// makeCircle (Engine.hs:53-54): CircleShape (circleWithRadius r)
struct Circle { double radius; };
inline Circle makeCircle(double radius) { return Circle{ radius }; }

// World: SoA body columns + CSR hull geometry + EmptiesVector filled flags.
class World {
public:
    std::vector<uint8_t> alive;
    std::vector<int32_t> vert_offset{ 0 };
    std::vector<double> local_x, local_y;
    std::vector<double> pos_x, pos_y, rot, inv_lin, inv_rot;
    std::vector<double> vel_x, vel_y, rot_vel;   // _physObjVel, _physObjRotVel (used by the device-resident world step)
    std::vector<double> mu, bounce;               // Material (World.hs:36-40)
    std::vector<double> radius;   // >= 0: CircleShape of that radius (no vertices); < 0: HullShape
    bool has_circles = false;
    bool geometry_dirty = true;

    // World.append (World.hs:77-84): returns the new object's key
    int append(const PhysicalObj &obj, const Hull &hull)
    {
        for (const V2 &v : hull) { local_x.push_back(v.x); local_y.push_back(v.y); }
        vert_offset.push_back((int32_t)local_x.size());
        alive.push_back(1);
        radius.push_back(-1.0);
        vel_x.push_back(obj.vel.x); vel_y.push_back(obj.vel.y); rot_vel.push_back(obj.rotVel);
        mu.push_back(0.0); bounce.push_back(0.0);
        pos_x.push_back(obj.pos.x); pos_y.push_back(obj.pos.y); rot.push_back(obj.rotPos);
        inv_lin.push_back(obj.invLin); inv_rot.push_back(obj.invRot);
        geometry_dirty = true;
        return (int)alive.size() - 1;
    }
    int append(const PhysicalObj &obj, const Circle &circle)
    {
        const int key = append(obj, Hull{});
        radius.back() = circle.radius;
        has_circles = true;
        return key;
    }
    // World.delete (World.hs:86-87): the slot stays, keys remain sparse
    void remove(int key) { alive.at((size_t)key) = 0; geometry_dirty = true; }
    int64_t slots() const { return (int64_t)alive.size(); }
};

// ObjectFeatureKey (Constraints/Contact.hs:36-39) + Flipping Contact (Contact/Types.hs:28-35)
struct KeyedContact {
    int i, j, featA, featB;
    bool flip;            // false = Same, true = Flip (Utils/Utils.hs:147)
    V2 normal, center;
    double depth;
};
// ContactConstraint (Constraints/Types.hs:41-51)
struct Constraint { double j[6]; double b; };
struct ContactConstraint {
    Constraint nonPen;
    V2 radiusA, radiusB, normal;   // RestitutionConstraint
    Constraint friction;
    double invEffNonPen, invEffFriction;   // effMassM2 (Constraint.hs:173-179)
};

struct Frame {
    std::vector<std::pair<int, int>> keys;            // Descending (Int, Int)
    std::vector<KeyedContact> contacts;               // Descending (ObjectFeatureKey Int, Flipping Contact)
    std::vector<ContactConstraint> constraints;       // row k belongs to contacts[k]
    double device_ms = 0.0;
};

class Engine {
public:
    explicit Engine(int device = 0, int64_t max_pairs = 1 << 16, int64_t max_contacts = 1 << 17)
        : device_(device), max_pairs_(max_pairs), max_contacts_(max_contacts) {}
    ~Engine() { if (ctx_) shapes_destroy(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    // one frame of the hot path; cos/sin are taken from the host libm exactly as rotate22 does
    // (Linear.hs:353-357), which keeps the results bit-identical to the reference arithmetic
    Frame frame(World &w, const ContactBehavior &beh, double dt, bool want_contacts = true, bool want_constraints = true)
    {
        ensure(w);
        const size_t n = (size_t)w.slots();
        std::vector<double> c(n), s(n);
        for (size_t k = 0; k < n; ++k) { c[k] = std::cos(w.rot[k]); s[k] = std::sin(w.rot[k]); }
        for (int attempt = 0;; ++attempt) {
            Columns col(max_pairs_, want_contacts ? max_contacts_ : 0, want_constraints ? max_contacts_ : 0);
            shapes_frame_out out{};
            col.bind(out);
            const int rc = shapes_frame(ctx_, (int64_t)n, w.pos_x.data(), w.pos_y.data(), w.rot.data(), c.data(), s.data(),
                                        w.inv_lin.data(), w.inv_rot.data(), dt, beh.contactBaumgarte,
                                        beh.contactPenetrationSlop, &out);
            if (rc == SHAPES_E_CAPACITY && attempt < 4) { grow(w, out.n_pairs, out.n_contacts); continue; }
            if (rc != SHAPES_OK) throw Error(rc, shapes_last_error(ctx_));
            return col.unpack(out, want_contacts, want_constraints, w);
        }
    }

    // ---- device-resident world: the whole updateWorld on the GPU -------
    // shapes_world_upload: the bodies of `w` move to HBM (and the EngineCache starts empty)
    void worldUpload(World &w)
    {
        ensure(w);
        const int rc = shapes_world_upload(ctx_, w.slots(), w.vel_x.data(), w.vel_y.data(), w.rot_vel.data(), w.pos_x.data(),
                                           w.pos_y.data(), w.rot.data(), nullptr, nullptr, w.inv_lin.data(), w.inv_rot.data(),
                                           w.mu.data(), w.bounce.data());
        if (rc != SHAPES_OK) throw Error(rc, shapes_last_error(ctx_));
        uploaded_ = &w;
    }
    // shapes_world_step: one updateWorld; on SHAPES_E_CAPACITY (the step leaves the world and the EngineCache
    // untouched) the capacities grow in place and the step is issued again
    shapes_step_stats worldStep(World &w, const shapes_step_config &cfg)
    {
        if (uploaded_ != &w) worldUpload(w);
        for (int attempt = 0;; ++attempt) {
            shapes_step_stats st{};
            const int rc = shapes_world_step(ctx_, &cfg, &st);
            if (rc == SHAPES_E_CAPACITY && attempt < 4) { grow(w, st.n_pairs, st.n_contacts); continue; }
            if (rc != SHAPES_OK) throw Error(rc, shapes_last_error(ctx_));
            return st;
        }
    }
    // shapes_world_download: velocities, positions and rotations back into `w`
    void worldDownload(World &w)
    {
        const int rc = shapes_world_download(ctx_, w.slots(), w.vel_x.data(), w.vel_y.data(), w.rot_vel.data(), w.pos_x.data(),
                                             w.pos_y.data(), w.rot.data(), nullptr, nullptr);
        if (rc != SHAPES_OK) throw Error(rc, shapes_last_error(ctx_));
    }

private:
    struct Columns {
        std::vector<int32_t> pi, pj, ki, kj, fa, fb;
        std::vector<uint8_t> flip;
        std::vector<double> f64[5], jn[6], jf[6], bnp, r[6], ie[2];
        Columns(int64_t np, int64_t nc, int64_t ncon)
        {
            pi.resize((size_t)np + 1); pj.resize((size_t)np + 1);
            for (auto *v : { &ki, &kj, &fa, &fb }) v->resize((size_t)nc + 1);
            flip.resize((size_t)nc + 1);
            for (auto &v : f64) v.resize((size_t)nc + 1);
            for (auto &v : jn) v.resize((size_t)ncon + 1);
            for (auto &v : jf) v.resize((size_t)ncon + 1);
            bnp.resize((size_t)ncon + 1);
            for (auto &v : r) v.resize((size_t)ncon + 1);
            for (auto &v : ie) v.resize((size_t)ncon + 1);
            has_c = nc > 0; has_k = ncon > 0;
        }
        bool has_c, has_k;
        void bind(shapes_frame_out &o)
        {
            o.pair_i = pi.data(); o.pair_j = pj.data();
            if (has_c) {
                o.key_i = ki.data(); o.key_j = kj.data(); o.feat_a = fa.data(); o.feat_b = fb.data(); o.flip = flip.data();
                o.normal_x = f64[0].data(); o.normal_y = f64[1].data(); o.center_x = f64[2].data();
                o.center_y = f64[3].data(); o.depth = f64[4].data();
            }
            if (has_k) {
                // Compact wire format: of a row's 27 doubles only the 11 independent ones cross PCIe -- the contact
                // (normal, centre, depth: above), the four cross terms, the Baumgarte bias and the two inverse
                // effective masses.  A NULL pointer means "not wanted, nothing is copied"; unpack() rebuilds the other
                // sixteen (sign copies of the normal / tangent, centre - position) bit for bit.
                compact = has_c;
                for (int q = 0; q < 6; ++q) {
                    const bool shipped = !compact || q == 2 || q == 5;
                    o.j_np[q] = shipped ? jn[q].data() : nullptr; o.j_f[q] = shipped ? jf[q].data() : nullptr;
                }
                o.b_np = bnp.data();
                if (!compact) {
                    o.ra_x = r[0].data(); o.ra_y = r[1].data(); o.rb_x = r[2].data(); o.rb_y = r[3].data();
                    o.rn_x = r[4].data(); o.rn_y = r[5].data();
                }
                o.inv_eff_np = ie[0].data(); o.inv_eff_f = ie[1].data();
            }
        }
        bool compact = false;
        Frame unpack(const shapes_frame_out &o, bool want_c, bool want_k, const World &w)
        {
            if (want_k && compact)
                for (int64_t k = 0; k < o.n_contacts; ++k) {
                    // With n the contact normal and t = clockwiseV2 n = (n.y, -n.x): J_np = (-n, x, n, x), J_f = (-t, x, t, x)
                    // on (penetrated, penetrator), halves swapped back for Flip (NonPenetration.hs:34-43,
                    // Friction.hs:31-44, Constraint.hs:96-98); ra = c - pos_i, rb = c - pos_j, rn = n for Same / -n for
                    // Flip (Restitution.hs:21-31).  Negation and one IEEE subtraction reproduce the device's bits.
                    const size_t u = (size_t)k;
                    const double nx = f64[0][u], ny = f64[1][u];
                    const double sx = flip[u] ? nx : -nx, sy = flip[u] ? ny : -ny;
                    jn[0][u] = sx; jn[1][u] = sy; jn[3][u] = -sx; jn[4][u] = -sy;
                    jf[0][u] = sy; jf[1][u] = -sx; jf[3][u] = -sy; jf[4][u] = sx;
                    r[4][u] = -sx; r[5][u] = -sy;
                    r[0][u] = f64[2][u] - w.pos_x[(size_t)ki[u]]; r[1][u] = f64[3][u] - w.pos_y[(size_t)ki[u]];
                    r[2][u] = f64[2][u] - w.pos_x[(size_t)kj[u]]; r[3][u] = f64[3][u] - w.pos_y[(size_t)kj[u]];
                }
            Frame f;
            f.device_ms = o.device_ms;
            f.keys.reserve((size_t)o.n_pairs);
            for (int64_t k = 0; k < o.n_pairs; ++k) f.keys.emplace_back(pi[(size_t)k], pj[(size_t)k]);
            if (want_c)
                for (int64_t k = 0; k < o.n_contacts; ++k) {
                    const size_t u = (size_t)k;
                    f.contacts.push_back(KeyedContact{ ki[u], kj[u], fa[u], fb[u], flip[u] != 0, { f64[0][u], f64[1][u] },
                                                       { f64[2][u], f64[3][u] }, f64[4][u] });
                }
            if (want_k)
                for (int64_t k = 0; k < o.n_contacts; ++k) {
                    const size_t u = (size_t)k;
                    ContactConstraint cc{};
                    for (int q = 0; q < 6; ++q) { cc.nonPen.j[q] = jn[q][u]; cc.friction.j[q] = jf[q][u]; }
                    cc.nonPen.b = bnp[u]; cc.friction.b = 0.0;     // Friction.toConstraint (Friction.hs:26-29)
                    cc.radiusA = { r[0][u], r[1][u] }; cc.radiusB = { r[2][u], r[3][u] }; cc.normal = { r[4][u], r[5][u] };
                    cc.invEffNonPen = ie[0][u]; cc.invEffFriction = ie[1][u];
                    f.constraints.push_back(cc);
                }
            return f;
        }
    };

    void create(const World &w)
    {
        if (ctx_) { shapes_destroy(ctx_); ctx_ = nullptr; }
        uploaded_ = nullptr;
        cap_slots_ = std::max<int64_t>(w.slots(), 1);
        cap_verts_ = std::max<int64_t>((int64_t)w.local_x.size(), 1);
        const int rc = shapes_create(&ctx_, device_, cap_slots_, cap_verts_, max_pairs_, max_contacts_);
        if (rc != SHAPES_OK) throw Error(rc, shapes_last_error(nullptr));
    }
    void ensure(World &w)
    {
        if (!ctx_ || w.slots() > cap_slots_ || (int64_t)w.local_x.size() > cap_verts_) { create(w); w.geometry_dirty = true; }
        if (w.geometry_dirty) {
            const int rc = shapes_set_shapes(ctx_, w.slots(), w.alive.data(), w.vert_offset.data(), w.local_x.data(),
                                             w.local_y.data(), nullptr, nullptr, w.has_circles ? w.radius.data() : nullptr);
            if (rc != SHAPES_OK) throw Error(rc, shapes_last_error(ctx_));
            w.geometry_dirty = false;
        }
    }
    // shapes_grow: capacities grow in place; key columns, EngineCache and an uploaded world survive
    void grow(World &w, int64_t need_pairs, int64_t need_contacts)
    {
        (void)w;
        max_pairs_ = std::max(max_pairs_, need_pairs + need_pairs / 4 + 1024);
        max_contacts_ = std::max(max_contacts_, need_contacts + need_contacts / 4 + 1024);
        const int rc = shapes_grow(ctx_, max_pairs_, max_contacts_);
        if (rc != SHAPES_OK) throw Error(rc, shapes_last_error(ctx_));
    }

    shapes_ctx *ctx_ = nullptr;
    World *uploaded_ = nullptr;
    int device_;
    int64_t max_pairs_, max_contacts_, cap_slots_ = 0, cap_verts_ = 0;
};

// ---- the reference's entry points, by name ------------------------------------------------------

// Aabb.culledKeys world :: Descending (Int, Int)
inline std::vector<std::pair<int, int>> culledKeys(Engine &e, World &w)
{
    return e.frame(w, ContactBehavior{}, 1.0, false, false).keys;
}
// prepareFrame keys world :: Descending (ObjectFeatureKey Int, Flipping Contact)
inline std::vector<KeyedContact> prepareFrame(Engine &e, World &w)
{
    return e.frame(w, ContactBehavior{}, 1.0, true, false).contacts;
}
// constraintGen beh dt fContact ab, for every contact of the frame (row k <-> contact k)
inline Frame constraintGen(Engine &e, const ContactBehavior &beh, double dt, World &w)
{
    return e.frame(w, beh, dt, true, true);
}

// External (World.hs:32-33): the two the reference defines (World/External.hs:16-28)
struct External { int kind = SHAPES_EXT_NONE; V2 v{ 0.0, 0.0 }; };
inline External makeConstantAccel(V2 a) { return External{ SHAPES_EXT_ACCEL, a }; }     // Engine.hs:44-45
inline External makeConstantForce(V2 f) { return External{ SHAPES_EXT_FORCE, f }; }

// Physics.Engine.Main.updateWorld (Engine/Main.hs:71-86) on the device-resident copy of `w`:
// culledKeys, applyExternal, prepareFrame, applyCachedSlns, improveWorld x2, advance, moveShapes.
// The bodies stay in HBM between calls; Engine::worldDownload brings them back into `w`.
inline shapes_step_stats updateWorld(Engine &e, World &w, double dt, const ContactBehavior &beh, const External &ext)
{
    shapes_step_config cfg{};
    cfg.dt = dt; cfg.baumgarte = beh.contactBaumgarte; cfg.slop = beh.contactPenetrationSlop;
    cfg.external_kind = ext.kind; cfg.external_x = ext.v.x; cfg.external_y = ext.v.y;
    cfg.solver_iterations = 2; cfg.warm_start = 1;
    return e.worldStep(w, cfg);
}

} // namespace shapes

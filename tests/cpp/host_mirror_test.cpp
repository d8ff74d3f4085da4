// Exercises include/shapes_b200.hpp the way the reference's own bench fixtures would:
//   * testOptBoxes   (shapes/bench/Physics/Contact/Benchmark.hs:16-27)  -> KAT-1
//   * testWorld      (shapes/bench/Physics/Broadphase/Benchmark.hs:50-52) -> KAT-3
//   * Stacks.makeScene (30,30) 0 (shapes/src/Physics/Scenes/Stacks.hs:110-113) -> config 1 counts
//   * updateWorld on the Stacks scene's floor + one box (device-resident world step)
// Exit code 0 = all checks passed; 3 = no usable GPU (the library refuses, it never falls back).
#include "shapes_b200.hpp"

#include <cmath>
#include <cstdio>

using namespace shapes;

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++fails; } } while (0)

// stacks / boxStack with the reference's repeated additions (Stacks.hs:34-56)
static void stacks(World &w, double bw, double bh, double center, double bottom, double spacing, int n_w, int n_h)
{
    double left = center - (bw * double(n_w - 1) / 2.0);
    for (int c = 0; c < n_w; ++c) {
        double y = bottom;
        for (int r = 0; r < n_h; ++r) {
            w.append(makePhysicalObj({ left, y }, 0.0, { 2.0, 1.0 }), makeRectangleHull(bw, bh));
            y = y + (bh + spacing);
        }
        left = left + bw;
    }
}

int main()
{
    try {
        Engine eng(0);
        {   // KAT-1
            World w;
            w.append(makePhysicalObj({ 1.0, 3.0 }, 0.0, { 1.0, 1.0 }), makeRectangleHull(2.0, 2.0));
            w.append(makePhysicalObj({ 0.0, 0.0 }, 0.0, { 1.0, 1.0 }), makeRectangleHull(4.0, 4.0));
            auto keys = culledKeys(eng, w);
            CHECK(keys.size() == 1 && keys[0] == std::make_pair(1, 0));
            auto cs = prepareFrame(eng, w);
            CHECK(cs.size() == 2);
            if (cs.size() == 2) {
                CHECK(cs[0].featA == 1 && cs[0].featB == 2 && cs[0].flip);
                CHECK(cs[1].featA == 0 && cs[1].featB == 2 && cs[1].flip);
                CHECK(cs[0].normal.x == 0.0 && cs[0].normal.y == -1.0);
                CHECK(cs[0].center.x == 0.0 && cs[0].center.y == 2.0 && cs[0].depth == 0.0);
                CHECK(cs[1].center.x == 2.0 && cs[1].center.y == 2.0 && cs[1].depth == 0.0);
            }
            Frame f = constraintGen(eng, ContactBehavior{ 0.01, 0.02 }, 0.01, w);
            CHECK(f.constraints.size() == 2);
            if (f.constraints.size() == 2) {
                // Flip: penetrated body is b (key 0, at (1,3)); J halves swapped back (Constraint.hs:96-98)
                const ContactConstraint &c0 = f.constraints[0];
                CHECK(c0.nonPen.j[0] == 0.0 && c0.nonPen.j[1] == -1.0);   // +n for the penetrator (a), n = (0,-1)
                CHECK(c0.nonPen.j[3] == -0.0 && c0.nonPen.j[4] == 1.0);   // -n for the penetrated body (b)
                CHECK(c0.nonPen.b == 0.0 && c0.friction.b == 0.0);
                CHECK(c0.normal.x == -0.0 && c0.normal.y == 1.0);         // Restitution: -n for Flip
            }
        }
        {   // KAT-6 (tests/golden/README.md): right triangle into the top edge of a 32x16 box next to its corner --
            // Same, ClipLeft, two points, depth > slop.  The C++ host takes cos/sin from libm, so the triangle is given
            // unrotated here (local = world - position): the same world vertices as the fixture's quarter turn.
            World w;
            w.append(makePhysicalObj({ -14.0, 6.0 }, 0.0, { 0.5, 4.0 }), makeHull({ { -1.5, -0.5 }, { 0.5, -2.0 }, { 3.3125, 1.75 } }));
            w.append(makePhysicalObj({ 1.0, -2.0 }, 0.0, { 2.0, 8.0 }), makeRectangleHull(32.0, 16.0));
            Frame f = constraintGen(eng, ContactBehavior{ 0.5, 0.25 }, 0.25, w);
            CHECK(f.keys.size() == 1 && f.contacts.size() == 2 && f.constraints.size() == 2);
            if (f.contacts.size() == 2 && f.constraints.size() == 2) {
                CHECK(f.contacts[0].featA == 0 && f.contacts[0].featB == 1 && !f.contacts[0].flip);
                CHECK(f.contacts[1].featA == 0 && f.contacts[1].featB == 0 && !f.contacts[1].flip);
                CHECK(f.contacts[0].normal.x == 0.0 && f.contacts[0].normal.y == 1.0);
                CHECK(f.contacts[0].center.x == -13.5 && f.contacts[0].center.y == 4.0 && f.contacts[0].depth == 2.0);
                CHECK(f.contacts[1].center.x == -15.0 && f.contacts[1].center.y == 5.125 && f.contacts[1].depth == 0.875);
                const ContactConstraint &c0 = f.constraints[0], &c1 = f.constraints[1];
                const double jn0[6] = { -0.0, -1.0, 14.5, 0.0, 1.0, 0.5 }, jf0[6] = { -1.0, 0.0, 6.0, 1.0, -0.0, 2.0 };
                const double jn1[6] = { -0.0, -1.0, 16.0, 0.0, 1.0, -1.0 }, jf1[6] = { -1.0, 0.0, 7.125, 1.0, -0.0, 0.875 };
                for (int q = 0; q < 6; ++q) {
                    CHECK(c0.nonPen.j[q] == jn0[q] && std::signbit(c0.nonPen.j[q]) == std::signbit(jn0[q]));
                    CHECK(c0.friction.j[q] == jf0[q] && std::signbit(c0.friction.j[q]) == std::signbit(jf0[q]));
                    CHECK(c1.nonPen.j[q] == jn1[q] && c1.friction.j[q] == jf1[q]);
                }
                CHECK(c0.nonPen.b == -3.5 && c1.nonPen.b == -1.25 && c0.friction.b == 0.0);     // (0.5 / 0.25) * (0.25 - depth)
                CHECK(c0.radiusA.x == -14.5 && c0.radiusA.y == 6.0 && c0.radiusB.x == 0.5 && c0.radiusB.y == -2.0);
                CHECK(c1.radiusA.x == -16.0 && c1.radiusA.y == 7.125 && c1.radiusB.x == -1.0 && c1.radiusB.y == -0.875);
                CHECK(c0.normal.x == 0.0 && c0.normal.y == 1.0);
            }
        }
        {   // KAT-4 / KAT-5: circles (Circle.contact; CircleVsHull through GJK)
            World w;
            w.append(makePhysicalObj({ 3.0, 0.0 }, 0.0, { 1.0, 1.0 }), makeCircle(1.5));
            w.append(makePhysicalObj({ 0.0, 0.0 }, 0.0, { 1.0, 1.0 }), makeCircle(2.0));
            auto cs = prepareFrame(eng, w);
            CHECK(cs.size() == 1);
            if (cs.size() == 1) {
                CHECK(cs[0].featA == 0 && cs[0].featB == 0 && !cs[0].flip);
                CHECK(cs[0].normal.x == 1.0 && cs[0].normal.y == 0.0);
                CHECK(cs[0].center.x == 1.75 && cs[0].center.y == 0.0 && cs[0].depth == 0.5);
            }
            World w2;
            w2.append(makePhysicalObj({ 0.0, 1.4 }, 0.0, { 1.0, 1.0 }), makeCircle(0.5));
            w2.append(makePhysicalObj({ 0.0, 0.0 }, 0.0, { 1.0, 1.0 }), makeRectangleHull(2.0, 2.0));
            auto c2 = prepareFrame(eng, w2);      // hull = shape a (larger key) => ((hullFeature, 0), Flip)
            CHECK(c2.size() == 1);
            if (c2.size() == 1) {
                CHECK(c2[0].featA == 1 && c2[0].featB == 0 && c2[0].flip);
                CHECK(c2[0].normal.x == 0.0 && c2[0].normal.y == -1.0);
                CHECK(c2[0].center.x == 0.0 && c2[0].center.y == 1.0 && c2[0].depth == 0.10000000000000009);
            }
        }
        {   // KAT-3
            World w;
            stacks(w, 0.2, 0.2, 0.0, -4.5, 0.0, 30, 30);
            auto keys = culledKeys(eng, w);
            CHECK(keys.size() == 3076);
            CHECK(!keys.empty() && keys[0] == std::make_pair(899, 898));
            World w1;
            stacks(w1, 0.2, 0.2, 0.0, -4.5, 1.0, 30, 30);
            auto keys1 = culledKeys(eng, w1);
            CHECK(keys1.size() == 840);
            CHECK(!keys1.empty() && keys1[0] == std::make_pair(899, 869));
        }
        {   // config 1: floor + 900 boxes; keys strictly descending, delete keeps keys sparse
            World w;
            w.append(makePhysicalObj({ 0.0, -6.0 }, 0.0, { 0.0, 0.0 }), makeRectangleHull(18.0, 1.0));
            stacks(w, 0.2, 0.2, 0.0, -4.5, 0.0, 30, 30);
            Frame f = constraintGen(eng, ContactBehavior{ 0.01, 0.02 }, 0.01, w);
            CHECK(f.keys.size() == 3076);   // the floor is 0.9 below the bottom row at frame 0
            for (size_t k = 1; k < f.keys.size(); ++k) CHECK(f.keys[k - 1] > f.keys[k]);
            CHECK(f.contacts.size() == f.constraints.size() && !f.contacts.empty());
            w.remove(900);
            auto keys = culledKeys(eng, w);
            for (auto &p : keys) CHECK(p.first != 900 && p.second != 900);
            CHECK(keys.size() < 3076);
        }
        {   // updateWorld (Engine/Main.hs:71-86) on device-resident bodies: Stacks' floor and one box
            // (mu 0.2, bounce 0, gravity (0,-2), ContactBehavior 0.01 0.02, dt 0.01): the box lands and rests
            World w;
            w.append(makePhysicalObj({ 0.0, -6.0 }, 0.0, { 0.0, 0.0 }), makeRectangleHull(18.0, 1.0));
            w.append(makePhysicalObj({ 0.0, 0.0 }, 0.0, { 0.0, -4.5 }, 0.0, { 2.0, 1.0 }), makeRectangleHull(0.2, 0.2));
            w.mu = { 0.2, 0.2 }; w.bounce = { 0.0, 0.0 };
            const External gravity = makeConstantAccel({ 0.0, -2.0 });
            shapes_step_stats st{};
            for (int f = 0; f < 400; ++f) st = updateWorld(eng, w, 0.01, ContactBehavior{ 0.01, 0.02 }, gravity);
            eng.worldDownload(w);
            CHECK(st.n_pairs == 1 && st.n_contacts >= 1 && st.warm == 1);
            CHECK(w.pos_y[0] == -6.0 && w.vel_y[0] == 0.0);                    // the static floor never moves
            CHECK(std::fabs(w.pos_y[1] - (-6.0 + 0.5 + 0.1)) < 0.01);          // resting on the floor's top face
            CHECK(std::fabs(w.vel_y[1]) < 0.02 && std::fabs(w.rot[1]) < 0.05);
            // a capacity error inside a step grows the ctx and carries on from the same state
            Engine small(0, 1, 1);
            World w2;
            w2.append(makePhysicalObj({ 0.0, -6.0 }, 0.0, { 0.0, 0.0 }), makeRectangleHull(18.0, 1.0));
            stacks(w2, 0.2, 0.2, 0.0, -5.41, 0.0, 6, 4);
            w2.mu.assign(w2.mu.size(), 0.2);
            for (int f = 0; f < 5; ++f) st = updateWorld(small, w2, 0.01, ContactBehavior{ 0.01, 0.02 }, gravity);
            CHECK(st.n_pairs > 1 && st.n_contacts > 1);
        }
    } catch (const Error &e) {
        std::printf("shapes::Error: %s\n", e.what());
        return e.code == SHAPES_E_CUDA ? 3 : 2;
    }
    std::printf(fails ? "host mirror: %d check(s) failed\n" : "host mirror: all checks passed\n", fails);
    return fails ? 1 : 0;
}

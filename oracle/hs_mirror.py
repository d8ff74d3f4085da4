"""A SECOND, independently structured restatement of the reference path, in pure Python.

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  PARITY UNPINNED by reference outputs: the
Haskell program cannot run here.  Purpose: oracle/shapes_oracle.c flattens the reference into index
arithmetic over SoA columns; this file instead keeps the reference's own shapes -- `Neighborhood`
records with next/prev links, `Maybe` / `Either` / `Flipping` values, `foldl1` over lists, `ClipResult`
constructors -- function by function under the reference's names.  The two were written from the
Haskell source separately; tests/test_oracle_mirror.py requires them to agree bit for bit, which catches
transcription slips in either (it cannot catch a shared misreading of the Haskell).

Python floats are IEEE binary64 and CPython never fuses a*b+c, so every expression below rounds like
GHC's SSE2 code.  Pure-Python loops: small worlds only.  Paths are relative to /root/reference/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional


# ---- Physics.Linear (shapes/src/Physics/Linear.hs) ---------------------------------------------------

def dotV2(a, b):                      # shapes-math/src/Shapes/Linear/Template.hs:108-110: foldl1 (+) of products
    return (a[0] * b[0]) + (a[1] * b[1])


def plusV2(a, b): return (a[0] + b[0], a[1] + b[1])                  # :100-102
def minusV2(a, b): return (a[0] - b[0], a[1] - b[1])                 # :114-116
def negateV2(a): return (-a[0], -a[1])                               # :201-203
def clockwiseV2(a): return (a[1], -a[0])                             # :161-163
def crossV2(a, b): return (a[0] * b[1]) - (a[1] * b[0])              # :118-120
def smulV2(s, v): return (v[0] * s, v[1] * s)                        # :72-74  liftV2 (*## s)
def zcrossV2(z, b): return (-(z * b[1]), z * b[0])                   # :127-130


def normalizeV2(v):                                                  # :165-168
    n = math.sqrt((v[0] * v[0]) + (v[1] * v[1]))
    return (_div(v[0], n), _div(v[1], n))


def _div(a, b):
    """IEEE division (Python raises on a zero divisor; Double# does not)."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


def mul3x3x3(a, b):                   # MatrixTemplate.hs:47-67: rows of a, columns of b, dotE each
    return [[((a[r][0] * b[0][c]) + (a[r][1] * b[1][c])) + (a[r][2] * b[2][c]) for c in range(3)] for r in range(3)]


def toTransform(pos, ori):            # Transform.hs:34-38,73-77 (forward matrix): translate . rotate
    c, s = ori                        # rotate22 (Linear.hs:353-357): (cos, sin) come from the caller
    translate = [[1.0, 0.0, pos[0]], [0.0, 1.0, pos[1]], [0.0, 0.0, 1.0]]
    rotate = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]
    return mul3x3x3(translate, rotate)


def afmul(m, p):                      # Linear.hs:217-220: m `mul3x3c` (x, y, 1)
    return (((m[0][0] * p[0]) + (m[0][1] * p[1])) + (m[0][2] * 1.0),
            ((m[1][0] * p[0]) + (m[1][1] * p[1])) + (m[1][2] * 1.0))


# lines and clipping (Linear.hs:233-343)
def toLine2(a, b): return (a, clockwiseV2(minusV2(b, a)))            # :237-240
def perpLine2(a, b): return (a, minusV2(b, a))                       # :242-245


def intersect2(l0, l1):               # :248-256 with invM2x2 (:194-199) and mul2x2c
    (p, (n0, n1)), (p2, (n2, n3)) = l0, l1
    b0, b1 = dotV2(p, (n0, n1)), dotV2(p2, (n2, n3))
    inv_det = _div(1.0, (n0 * n3) - (n1 * n2))
    m = ((inv_det * n3, inv_det * (-n1)), (inv_det * (-n2), inv_det * n0))
    return ((m[0][0] * b0) + (m[0][1] * b1), (m[1][0] * b0) + (m[1][1] * b1))


def clipSegment(boundary, incident, seg):                            # :327-343
    a, b = seg
    c = intersect2(boundary, incident)
    n = boundary[1]
    a1, b1, c1 = dotV2(a, n), dotV2(b, n), dotV2(c, n)
    if a1 < c1:
        return ("ClipBoth", c) if b1 < c1 else ("ClipLeft", c)
    if b1 < c1:
        return ("ClipRight", c)
    return ("ClipNone", None)


# ---- Physics.Contact.ConvexHull ----------------------------------------------------------------------

@dataclass
class Neighborhood:                                                  # ConvexHull.hs:29-34
    hull: "ConvexHull"
    index: int
    center: tuple
    unit_normal: tuple

    def next(self):
        return self.hull.neighborhoods[nextIndex(self.hull.count - 1, self.index)]

    def prev(self):
        return self.hull.neighborhoods[prevIndex(self.hull.count - 1, self.index)]

    def with_center(self, c):                                        # `set neighborhoodCenter c`
        return Neighborhood(self.hull, self.index, c, self.unit_normal)


def nextIndex(max_i, i): return i + 1 if i < max_i else 0            # :228-230
def prevIndex(max_i, i): return i - 1 if i > 0 else max_i            # :232-234


class ConvexHull:                                                    # :61-68
    def __init__(self, local_vertices):
        self.count = len(local_vertices)
        self.local = list(local_vertices)
        self.vertices = list(local_vertices)
        self._refresh()
        # listToHull (:151-167): _hullExtents from the LOCAL vertices, frozen afterwards
        self.extents = [extentIndices(extentAlong(self, n)) for n in self.edge_normals]

    def _refresh(self):
        mx = self.count - 1
        self.edge_normals = [normalizeV2(clockwiseV2(minusV2(self.vertices[nextIndex(mx, i)], self.vertices[i])))
                             for i in range(self.count)]             # unitEdgeNormal (:218-226)
        self.neighborhoods = [Neighborhood(self, i, self.vertices[i], self.edge_normals[i]) for i in range(self.count)]

    def setHullTransform(self, f):                                   # :184-195: from the local vertices; extents untouched
        self.vertices = [f(v) for v in self.local]
        self._refresh()
        return self


def extentAlong_(hull, d):                                           # extentAlong' (:81-92)
    items = [((dotV2(n.center, d), n), (dotV2(n.center, d), n)) for n in hull.neighborhoods]
    acc = items[0]
    for (min_b, max_b) in items[1:]:                                 # foldl1 g
        min_a, max_a = acc
        acc = (min_b if min_b[0] < min_a[0] else min_a, max_b if max_b[0] > max_a[0] else max_a)
    return acc[0][1], acc[1][1]


def extentAlong(hull, d):                                            # :94-100
    mn, mx = extentAlong_(hull, d)
    return {"min": mn, "max": mx, "proj": (dotV2(mn.center, d), dotV2(mx.center, d))}


def extentIndices(ext): return (ext["min"].index, ext["max"].index)  # :102-105


def extentAlongSelf(hull, index, d):                                 # :111-118
    i_min, i_max = hull.extents[index]
    mn, mx = hull.neighborhoods[i_min], hull.neighborhoods[i_max]
    return {"min": mn, "max": mx, "proj": (dotV2(mn.center, d), dotV2(mx.center, d))}


# ---- Physics.Contact.SAT -----------------------------------------------------------------------------

def overlapTest(x, y):                                               # SAT.hs:74-83
    (a, b), (c, d) = x, y
    return not (c > b or d < a)


def overlapAmount(x, y):                                             # :86-96 -> Maybe
    return (x[1] - y[0]) if overlapTest(x, y) else None


def overlap(s_edge, edge, s_pen):                                    # :103-117 -> Maybe Overlap
    d = edge.unit_normal
    ext_s = extentAlongSelf(s_edge, edge.index, d)
    ext_p = extentAlong(s_pen, d)
    oval = overlapAmount(ext_s["proj"], ext_p["proj"])
    return None if oval is None else {"edge": edge, "depth": oval, "penetrator": ext_p["min"]}


def minOverlap_(a, b):                                               # minOverlap' / minOverlap (:121-148): foldl1 f
    results = []
    for edge in a.neighborhoods:
        o = overlap(a, edge, b)
        results.append(("Separated", edge) if o is None else ("MinOverlap", o))
    acc = results[0]
    for r in results[1:]:
        if acc[0] == "Separated":
            continue
        if r[0] == "Separated":
            acc = r
        elif r[1]["depth"] < acc[1]["depth"]:
            acc = r
    return acc


def penetratingEdge(ovl):                                            # :152-166
    b = ovl["penetrator"]
    c, a = b.next(), b.prev()
    n = ovl["edge"].unit_normal
    abn = abs(dotV2(minusV2(b.center, a.center), n))
    bcn = abs(dotV2(minusV2(c.center, b.center), n))
    return (b, c) if bcn < abn else (a, b)


def penetratedEdge(ovl):                                             # :169-171
    return ovl["edge"], ovl["edge"].next()


def lApplyClip_(res, seg):                                           # lApplyClip' (Linear.hs:316-320) with neighborhoodCenter
    a, b = seg
    kind, c = res
    if kind == "ClipBoth":
        return None
    if kind == "ClipLeft":
        return (a.with_center(c), b)
    if kind == "ClipRight":
        return (a, b.with_center(c))
    return (a, b)


def applyClip__(res, seg):                                           # applyClip'' (Linear.hs:285-292)
    a, b = seg
    kind, _ = res
    if kind == "ClipLeft":
        return ("Left", b)
    if kind == "ClipRight":
        return ("Left", a)
    if kind == "ClipBoth":
        return None
    return ("Right", (a, b))


def clipEdge(edge_pair, n, inc_):                                    # SAT.hs:190-218
    a, b = edge_pair[0].center, edge_pair[1].center
    c, d = inc_[0].center, inc_[1].center
    cd = toLine2(c, d)                                               # from the UNCLIPPED incident edge
    inc1 = lApplyClip_(clipSegment(perpLine2(a, b), cd, (c, d)), inc_)
    if inc1 is None:
        return None
    inc2 = lApplyClip_(clipSegment(perpLine2(b, a), cd, (inc1[0].center, inc1[1].center)), inc1)
    if inc2 is None:
        return None
    return applyClip__(clipSegment((a, negateV2(n)), cd, (inc2[0].center, inc2[1].center)), inc2)


def contact_(ovl):                                                   # :261-267 -> Maybe Contact
    pts = clipEdge(penetratedEdge(ovl), ovl["edge"].unit_normal, penetratingEdge(ovl))
    return None if pts is None else {"edge": ovl["edge"], "penetrator": pts}


def contact(a, b):                                                   # contactDebug / contact (:238-258)
    ab, ba = minOverlap_(a, b), minOverlap_(b, a)
    # eitherBranchBoth ((<) `on` depth) (Utils.hs:230-235): the first Left wins
    if ab[0] == "Separated":
        return ("Same", ("Left", ab[1]))
    if ba[0] == "Separated":
        return ("Flip", ("Left", ba[1]))
    flipping, ovl = ("Same", ab[1]) if ab[1]["depth"] < ba[1]["depth"] else ("Flip", ba[1])
    c = contact_(ovl)                                                # convertContactResult: Right Nothing -> Nothing
    return None if c is None else (flipping, ("Right", c))


def flattenContactPoints(pts):                                       # :181-187
    if pts[0] == "Left":
        return [pts[1]]
    p1, p2 = pts[1]
    return [p1, p2] if p1.index > p2.index else [p2, p1]


# ---- Physics.Contact.HullVsHull + Physics.Constraints.Contact ----------------------------------------

def contactDepth(edge, pen):                                         # HullVsHull.hs:25-37: f v - f p, f = afdot' n
    n = edge.unit_normal
    return dotV2(edge.center, n) - dotV2(pen.center, n)


def generateContacts(a, b):                                          # HullVsHull.hs:54-90 -> [((featA, featB), Flipping Contact)]
    res = contact(a, b)
    if res is None or res[1][0] == "Left":                           # unwrapContactResult: separating axis -> Nothing
        return []
    flipping, (_, c) = res
    out = []
    for pen in flattenContactPoints(c["penetrator"]):
        feat = (c["edge"].index, pen.index)
        if flipping == "Flip":                                       # flipExtractPair fst (Utils.hs:184-186)
            feat = (feat[1], feat[0])
        out.append((feat, (flipping, {"normal": c["edge"].unit_normal, "center": pen.center,
                                      "depth": contactDepth(c["edge"], pen)})))
    return out


def keyedContacts(ij, shapes):                                       # Constraints/Contact.hs:41-57
    return [((ij, feat), fc) for feat, fc in generateContacts(*shapes)]


# constraint generators ------------------------------------------------------------------------------

def _flip_map(f, fc, ab):                                            # flipMap + flipExtract for Constraint (Utils.hs:175-215)
    flipping, c = fc
    if flipping == "Same":
        return f(c, ab)
    j, b = f(c, (ab[1], ab[0]))
    return (j[3:] + j[:3], b)                                        # flip3v3 (Linear.hs:149-151)


def nonpen_constraintGen(beh, dt, fc, ab):                           # NonPenetration.hs:16-55
    def to_constraint(c, pair):
        a, b = pair
        n, p = c["normal"], c["center"]
        ja = negateV2(n) + (crossV2(minusV2(a["pos"], p), n),)
        jb = n + (crossV2(minusV2(p, b["pos"]), n),)
        d = c["depth"]
        bias = (_div(beh[0], dt)) * (beh[1] - d) if d > beh[1] else 0.0
        return ja + jb, bias
    return _flip_map(to_constraint, fc, ab)


def friction_constraintGen(fc, ab):                                  # Friction.hs:18-44
    def to_constraint(c, pair):
        a, b = pair
        p = c["center"]
        tb = clockwiseV2(c["normal"])
        ta = negateV2(tb)
        ja = ta + (crossV2(minusV2(p, a["pos"]), ta),)
        jb = tb + (crossV2(minusV2(p, b["pos"]), tb),)
        return ja + jb, 0.0
    return _flip_map(to_constraint, fc, ab)


def restitution_constraintGen(fc, ab):                               # Restitution.hs:21-31
    flipping, c = fc
    a, b = ab
    n = c["normal"] if flipping == "Same" else negateV2(c["normal"])
    return minusV2(c["center"], a["pos"]), minusV2(c["center"], b["pos"]), n


def constraintGen(beh, dt, fc, ab):                                  # Constraints/Contact.hs:60-72
    return {"nonpen": nonpen_constraintGen(beh, dt, fc, ab), "restitution": restitution_constraintGen(fc, ab),
            "friction": friction_constraintGen(fc, ab)}


# ---- Physics.Constraint: solving -----------------------------------------------------------------------

def dotV6(a, b):
    acc = a[0] * b[0]
    for k in range(1, 6):
        acc = acc + (a[k] * b[k])
    return acc


def invMassM2(a, b):                                                 # Constraint.hs:118-120
    return (a["inv"][0], a["inv"][0], a["inv"][1], b["inv"][0], b["inv"][0], b["inv"][1])


def velocity2(a, b):                                                 # :150-157
    return a["vel"] + (a["rotvel"],) + b["vel"] + (b["rotvel"],)


def effMassM2(j, a, b):                                              # :173-179
    im = invMassM2(a, b)
    return dotV6(tuple(j[k] * im[k] for k in range(6)), j)


def lagrangian2(ab, constraint):                                     # :164-169
    j, b = constraint
    return _div(-(dotV6(j, velocity2(*ab)) + b), effMassM2(j, *ab))


def applyLagrangian(lagr, constraint, ab):                           # :216-222 -> applyLagrangian2 (:197-203)
    j = constraint[0]
    im = invMassM2(*ab)
    v = velocity2(*ab)
    pc = tuple(jk * lagr for jk in j)                                # constraintImpulse2: l `smulV6` j
    v2 = tuple(v[k] + (pc[k] * im[k]) for k in range(6))             # updateVelocity2_: v + (im `vmulDiag6'` pc)
    a, b = ab
    return (dict(a, vel=(v2[0], v2[1]), rotvel=v2[2]), dict(b, vel=(v2[3], v2[4]), rotvel=v2[5]))


def hs_max(x, y): return y if x <= y else x                          # Ord's default max (GHC.Classes)
def hs_min(x, y): return x if x <= y else y


def bounceB(bounciness_ab, rc, ab):                                  # Restitution.hs:34-47
    ra, rb, n = rc
    a, b = ab
    nva = negateV2(a["vel"])
    nwa = zcrossV2(-a["rotvel"], ra)
    wb = zcrossV2(b["rotvel"], rb)
    closing = plusV2(plusV2(plusV2(nva, nwa), b["vel"]), wb)
    return hs_min(0.0, hs_min(*bounciness_ab) * dotV2(closing, n))


def improveContactSln(cc, cached, mu_ab, bounce_ab, ab):             # Solvers/Contact.hs:124-143 for one contact
    """Returns (new pair, new cached ContactLagrangian)."""
    jn, bn = cc["nonpen"]
    nonpen = (jn, bn + bounceB(bounce_ab, cc["restitution"], ab))    # nonPenWithRestitution (Constraints/Contact.hs:74-85)
    new_np, new_f = lagrangian2(ab, nonpen), lagrangian2(ab, cc["friction"])   # contactLagrangian (:87-97)
    # solutionProcessor (:99-110): positive (SolutionProcessors.hs:28-35), clampAbs (:37-53)
    apply_np = hs_max(new_np, -cached[0])
    cache_np = cached[0] + apply_np
    max_thresh = cache_np * _div(mu_ab[0] + mu_ab[1], 2)
    accum = cached[1] + new_f
    accum2 = max_thresh if accum > max_thresh else (-max_thresh if accum < -max_thresh else accum)
    apply_f = accum2 - cached[1]
    ab = applyLagrangian(apply_np, cc["nonpen"], ab)                 # applySln (:54-65): applyFriction . applyNonPen
    ab = applyLagrangian(apply_f, cc["friction"], ab)
    return ab, (cache_np, accum2)


# ---- driver: one frame's contacts of a flattened world (shapes_b200.world.World) ------------------------

def hulls_of(world, cos_rot, sin_rot):
    """listToHull once (local vertices), then moveShape (World.hs:132-134) = setHullTransform (transform ...)."""
    hulls = []
    for s in range(world.n_slots):
        o, e = int(world.vert_offset[s]), int(world.vert_offset[s + 1])
        if not world.alive[s] or e == o:
            hulls.append(None)
            continue
        h = ConvexHull([(float(world.local_x[k]), float(world.local_y[k])) for k in range(o, e)])
        m = toTransform((float(world.pos_x[s]), float(world.pos_y[s])), (float(cos_rot[s]), float(sin_rot[s])))
        hulls.append(h.setHullTransform(lambda p, m=m: afmul(m, p)))
    return hulls


def prepare_frame(world, hulls, pairs, beh, dt):
    """prepareFrame (Solvers/Contact.hs:40-52) + constraintGen per contact: list of rows."""
    rows = []
    for (i, j) in pairs:
        a = {"pos": (float(world.pos_x[i]), float(world.pos_y[i]))}
        b = {"pos": (float(world.pos_x[j]), float(world.pos_y[j]))}
        for (ij, feat), fc in keyedContacts((i, j), (hulls[i], hulls[j])):
            rows.append({"key": ij + feat, "flip": 0 if fc[0] == "Same" else 1, "contact": fc[1],
                         "constraint": constraintGen(beh, dt, fc, (a, b))})
    return rows


# ---- circles: Physics.Contact.Circle / GJK / CircleVsHull, Physics.Contact.generateContacts ----------

def sdivV2(s, v): return (_div(v[0], s), _div(v[1], s))              # Linear.hs:80-82  liftV2 (/## s)
def sqLengthV2(v): return (v[0] * v[0]) + (v[1] * v[1])              # :175-176
def crosszV2(a, bz): return (a[1] * bz, -(a[0] * bz))                # :121-124
def sameDirection(a, b): return dotV2(a, b) > 0.0                    # GJK.hs:138-139


def crossV2V2(a, b, c):                                              # Linear.hs:135-139
    abz = (a[0] * b[1]) - (a[1] * b[0])
    return (-(abz * c[1]), abz * c[0])


def circle_contact(ca, ra, cb, rb):                                  # Circle.contact (Circle.hs:30-53); A = penetratee
    ab = minusV2(cb, ca)
    rab = ra + rb
    ab_sq = sqLengthV2(ab)
    if not (rab * rab >= ab_sq):
        return None
    ab_len = math.sqrt(ab_sq)
    ab_n = sdivV2(ab_len, ab)
    a1 = plusV2(smulV2(ra, ab_n), ca)
    b1 = plusV2(smulV2(-rb, ab_n), cb)
    return {"normal": ab_n, "center": sdivV2(2, plusV2(a1, b1)), "depth": ra + rb - ab_len}


def support(hull, d):                                                # ConvexHull.hs:124-128: first maximum
    acc = (dotV2(hull.neighborhoods[0].center, d), hull.neighborhoods[0])
    for n in hull.neighborhoods[1:]:
        dist = dotV2(n.center, d)
        if dist > acc[0]:
            acc = (dist, n)
    return acc[1]


def closestSimplex(hull, origin):                                    # GJK.hs:52-69
    simplex = [hull.neighborhoods[0]]                                # Left (Simplex1 a); most recent vertex first
    d = minusV2(origin, simplex[0].center)
    for _ in range(64):                                              # the reference loops until a repeat; bounded here like the C oracle
        aa = support(hull, d)
        if any(aa.index == s.index for s in simplex):                # extendSimplex1/2: a repeat ends the search
            return ("Simplex12", simplex)
        simplex = [aa] + simplex
        a = simplex[0].center
        ao = minusV2(origin, a)
        ab = minusV2(simplex[1].center, a)
        if len(simplex) == 2:                                        # shiftSimplex2 (:98-112)
            if sameDirection(ab, ao):
                d = crossV2V2(ab, ao, ab)
            else:
                simplex, d = [simplex[0]], ao
        else:                                                        # shiftSimplex3 (:114-136)
            ac = minusV2(simplex[2].center, a)
            abc = crossV2(ab, ac)
            abcac = zcrossV2(abc, ac)
            ababc = crosszV2(ab, abc)
            star = None
            if sameDirection(abcac, ao):
                if sameDirection(ac, ao):
                    simplex, d = [simplex[0], simplex[2]], crossV2V2(ac, ao, ac)
                else:
                    star = True
            elif sameDirection(ababc, ao):
                star = True
            else:
                return ("Simplex3", simplex)                         # encloses the target
            if star:
                if sameDirection(ab, ao):
                    simplex, d = [simplex[0], simplex[1]], crossV2V2(ab, ao, ab)
                else:
                    simplex, d = [simplex[0]], ao
    return ("Simplex12", None)


def circle_hull_contacts(center, radius, hull):                      # CircleVsHull.generateContacts (:18-69)
    kind, simplex = closestSimplex(hull, center)
    if kind != "Simplex12" or simplex is None:
        return None                                                  # Simplex3': deep overlap -> Nothing
    feature = simplex[0]
    a = feature.center
    if len(simplex) == 2:                                            # closestAlong (:60-69)
        b = simplex[1].center
        ao, ab = minusV2(center, a), minusV2(b, a)
        ab_norm = normalizeV2(ab)
        a = plusV2(smulV2(dotV2(ao, ab_norm), ab_norm), a)
    ab = minusV2(center, a)                                          # processSimplex_ (:42-58)
    ab_sq = sqLengthV2(ab)
    if radius * radius < ab_sq:
        return None
    ab_len = math.sqrt(ab_sq)
    return feature.index, {"normal": negateV2(sdivV2(ab_len, ab)), "center": a, "depth": radius - ab_len}


def generateContactsShapes(sa, sb):                                  # Physics.Contact.generateContacts (Contact.hs:22-40)
    """A shape is ("hull", ConvexHull) or ("circle", centre, radius)."""
    if sa[0] == "circle" and sb[0] == "circle":
        c = circle_contact(sa[1], sa[2], sb[1], sb[2])
        return [] if c is None else [((0, 0), ("Same", c))]
    if sa[0] == "circle":
        r = circle_hull_contacts(sa[1], sa[2], sb[1])
        return [] if r is None else [((0, r[0]), ("Same", r[1]))]
    if sb[0] == "circle":
        r = circle_hull_contacts(sb[1], sb[2], sa[1])
        return [] if r is None else [((r[0], 0), ("Flip", r[1]))]
    return generateContacts(sa[1], sb[1])


# ---- Physics.Broadphase.Aabb -------------------------------------------------------------------------

def mergeRange(x, y):                                                # Aabb.hs:104-110
    (a, b), (c, d) = x, y
    return (a if a < c else c, b if b > d else d)


def hullToAabb(hull):                                                # :81-84: foldl1 mergeAabb over the vertices
    acc = ((hull.vertices[0][0],) * 2, (hull.vertices[0][1],) * 2)
    for v in hull.vertices[1:]:
        acc = (mergeRange(acc[0], (v[0], v[0])), mergeRange(acc[1], (v[1], v[1])))
    return acc


def circleToAabb(center, r):                                         # :86-88
    return ((center[0] - r, center[0] + r), (center[1] - r, center[1] + r))


def boundsOverlap(x, y):                                             # :69-72
    (a, b), (c, d) = x, y
    return not (c > b or d < a)


def aabbCheck(p, q): return boundsOverlap(p[0], q[0]) and boundsOverlap(p[1], q[1])   # :75-78


def unorderedPairs(n):                                               # :155-163
    out = []
    if n < 2:
        return out
    x, y = n - 1, n - 2
    while True:
        out.append((x, y))
        if (x, y) == (1, 0):
            return out
        if y == 0:
            x, y = x - 1, x - 2
        else:
            y -= 1


def culledKeys(tagged):                                              # :168-183; tagged = [(key, aabb, isStatic)] in traversal order
    out = []
    for (i, j) in unorderedPairs(len(tagged)):
        ki, a, sa = tagged[i]
        kj, b, sb = tagged[j]
        if not (sa and sb) and aabbCheck(a, b):
            out.append((ki, kj))
    return out


# ---- Utils.Descending.descZipVector as applyCachedSlns uses it ----------------------------------------

def descZipVector(these, those):                                     # Descending.hs:47-71 -> per `this`: matching `that` index or None
    out, that_i = [], 0
    for this_key in these:
        match = None
        while that_i < len(those):
            that_key = those[that_i]
            if this_key < that_key:
                that_i += 1                                          # keep looking
                continue
            if this_key == that_key:
                match = that_i
                that_i += 1
            break
        out.append(match)
    return out


# ---- Physics.Broadphase.Grid (the variant updateWorld calls, Engine/Main.hs:75) ------------------------

def axialIndex(axis, val):                                           # Grid.hs:124-127; axis = (length, unit, origin)
    return math.floor(_div(val - axis[2], axis[1]))


def boxIndices(axes, box):                                           # :129-141
    x_axis, y_axis = axes
    xs = range(axialIndex(x_axis, box[0][0]), axialIndex(x_axis, box[0][1]) + 1)
    ys = range(axialIndex(y_axis, box[1][0]), axialIndex(y_axis, box[1][1]) + 1)
    return [x + (y * x_axis[0]) for x in xs for y in ys]              # flattenIndex' (:117-118)


def fromTaggedAabbs(axes, tagged):                                   # :96-105: IntMap (IntMap TaggedAabb)
    grid = {}
    for key, box, is_static in tagged:
        for index in boxIndices(axes, box):
            grid.setdefault(index, {})[key] = (is_static, box)
    return grid


def allPairs(xs):                                                    # :84-89 (the accumulation order does not survive the sort)
    out = []
    for a in range(len(xs)):
        for b in range(a + 1, len(xs)):
            out.append((xs[a], xs[b]))
    return out


def grid_culledKeys(axes, tagged):                                   # toGrid + culledKeys + culledKeys' + uniq (:67-95)
    pairs = []
    for square in fromTaggedAabbs(axes, tagged).values():
        desc = sorted(square.items(), key=lambda kv: -kv[0])         # IM.toDescList
        for (a, (sa, box_a)), (b, (sb, box_b)) in allPairs(desc):
            if sa and sb:
                continue                                             # two static shapes
            if aabbCheck(box_a, box_b):
                pairs.append((a, b))
    pairs.sort(reverse=True)                                         # sortBy (flip compare)
    out = []
    for p in pairs:                                                  # uniq
        if not out or out[-1] != p:
            out.append(p)
    return out


# ---- Physics.Engine.Main.updateWorld, assembled from the pieces above ---------------------------------

def constantAccel(a):                                                # World/External.hs:23-27
    def ext(dt, o):
        if 0.0 == o["inv"][0]:                                       # isStaticLin
            return o
        return dict(o, vel=plusV2(o["vel"], smulV2(dt, a)))
    return ext


def advanceObj(o, dt):                                               # Constraint.hs:225-229
    return dict(o, pos=plusV2(smulV2(dt, o["vel"]), o["pos"]), rot=(dt * o["rotvel"]) + o["rot"])


def applySln(lagr, cc, ab):                                          # Solvers/Contact.hs:54-65: applyFriction . applyNonPen
    ab = applyLagrangian(lagr[0], cc["nonpen"], ab)
    return applyLagrangian(lagr[1], cc["friction"], ab)


def updateWorld(objs, hulls_local, mats, cache, ext, dt, beh, sincos, grid_axes=None):
    """One Physics.Engine.Main.updateWorld (Engine/Main.hs:71-86).
    objs: list of dicts {vel, rotvel, pos, rot, inv, cs=(cos, sin) the shape was last moved with} (all slots filled);
    hulls_local: local CCW vertices per object; mats: [(mu, bounce)]; cache: [(key, (lam_np, lam_f))] descending.
    Returns the new cache; objs are updated in place."""
    n = len(objs)
    # the shapes as moveShapes left them (World.hs:132-140)
    hulls = [ConvexHull(hulls_local[k]).setHullTransform(lambda p, m=toTransform(objs[k]["pos"], objs[k]["cs"]): afmul(m, p))
             for k in range(n)]
    tagged = [(k, hullToAabb(hulls[k]), objs[k]["inv"] == (0.0, 0.0)) for k in range(n)]
    keys = grid_culledKeys(grid_axes, tagged) if grid_axes else culledKeys(tagged)      # :75
    for k in range(n):                                               # applyExternal (:76, World.hs:156-158)
        objs[k] = ext(dt, objs[k])
    k_contacts = []                                                  # prepareFrame (:77)
    for (i, j) in keys:
        k_contacts += keyedContacts((i, j), (hulls[i], hulls[j]))
    # applyCachedSlns (:78-80, Solvers/Contact.hs:74-121)
    this_keys = [ij + feat for (ij, feat), _ in k_contacts]
    match = descZipVector(this_keys, [k for k, _ in cache])
    lagrangians, constraints = [], []
    for (key, fc), m in zip(k_contacts, match):
        (i, j), _ = key
        cc = constraintGen(beh, dt, fc, (objs[i], objs[j]))
        if m is None:
            lagrangians.append((key[0] + key[1], (0.0, 0.0)))        # newCache
        else:
            objs[i], objs[j] = applySln(cache[m][1], cc, (objs[i], objs[j]))   # useCache
            lagrangians.append(cache[m])
        constraints.append(cc)
    for _ in range(2):                                               # improveWorld x2 (:81-82)
        for k, (key, _) in enumerate(k_contacts):
            (i, j), _ = key
            (objs[i], objs[j]), lam = improveContactSln(constraints[k], lagrangians[k][1], (mats[i][0], mats[j][0]),
                                                         (mats[i][1], mats[j][1]), (objs[i], objs[j]))
            lagrangians[k] = (lagrangians[k][0], lam)
    for k in range(n):                                               # advance (:83) + moveShapes (:84)
        objs[k] = advanceObj(objs[k], dt)
        objs[k]["cs"] = sincos(objs[k]["rot"])
    return lagrangians

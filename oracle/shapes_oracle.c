/*
 * shapes_oracle.c -- CPU restatement of the ublubu/shapes collision hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see shapes_oracle.h).  PARITY UNPINNED by reference
 * outputs: the Haskell reference cannot be built here (no GHC); this file
 * follows the cited lines operation for operation and is pinned by the
 * hand-derived known-answer vectors under tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile).  Every
 * floating-point expression below is parenthesised exactly as the reference
 * evaluates it: the TH-generated dot product is a left fold of separate
 * products (shapes-math/src/Shapes/Linear/Template.hs:108-110) and GHC's NCG
 * emits one SSE2 scalar instruction per primop (no FMA).
 *
 * Paths in comments are relative to /root/reference/.
 */
#include "shapes_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* Physics.Linear (shapes/src/Physics/Linear.hs)                       */
/* ------------------------------------------------------------------ */

typedef struct { double x, y; } v2;

/* dotV2: Template.hs:108-110 -> (a0*b0)+(a1*b1) */
static inline double dot2(v2 a, v2 b) { return (a.x * b.x) + (a.y * b.y); }
/* minusV2: Linear.hs:114-116 */
static inline v2 sub2(v2 a, v2 b) { v2 r = { a.x - b.x, a.y - b.y }; return r; }
/* negateV2: Linear.hs:201-203 */
static inline v2 neg2(v2 a) { v2 r = { -a.x, -a.y }; return r; }
/* clockwiseV2: Linear.hs:161-163 */
static inline v2 clockwise2(v2 a) { v2 r = { a.y, -a.x }; return r; }
/* crossV2: Linear.hs:118-120 */
static inline double cross2(v2 a, v2 b) { return (a.x * b.y) - (a.y * b.x); }
/* normalizeV2: Linear.hs:165-168 */
static inline v2 normalize2(v2 a)
{
    double n = sqrt((a.x * a.x) + (a.y * a.y));
    v2 r = { a.x / n, a.y / n };
    return r;
}

double orc_dot_v2(double ax, double ay, double bx, double by)
{
    v2 a = { ax, ay }, b = { bx, by };
    return dot2(a, b);
}

/* mul2x2x2: MatrixTemplate.hs:47-67 (rows of a, columns of b, dotE each) */
void orc_mul2x2x2(const double *a, const double *b, double *out)
{
    out[0] = (a[0] * b[0]) + (a[1] * b[2]);
    out[1] = (a[0] * b[1]) + (a[1] * b[3]);
    out[2] = (a[2] * b[0]) + (a[3] * b[2]);
    out[3] = (a[2] * b[1]) + (a[3] * b[3]);
}

/* mul3x3x3: MatrixTemplate.hs:47-67, instantiated Linear.hs:36.
 * Row-major a, b; each entry is dotE row col = ((r0*c0)+(r1*c1))+(r2*c2). */
static void mul3x3x3(const double a[9], const double b[9], double out[9])
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            out[3 * r + c] = ((a[3 * r + 0] * b[0 + c]) + (a[3 * r + 1] * b[3 + c])) +
                             (a[3 * r + 2] * b[6 + c]);
}

/* afmul: Linear.hs:217-220 -- t `mul3x3c` (a, b, 1.0), keep x and y. */
static inline v2 afmul(const double t[9], v2 p)
{
    v2 r;
    r.x = ((t[0] * p.x) + (t[1] * p.y)) + (t[2] * 1.0);
    r.y = ((t[3] * p.x) + (t[4] * p.y)) + (t[5] * 1.0);
    return r;
}

/* toTransform pos ori = translate(pos) . rotate(ori), forward matrix only
 * (Transform.hs:34-38, 73-77; Linear.hs:349-381). cos/sin are passed in. */
static void to_transform(double px, double py, double c, double s, double out[9])
{
    const double transl[9] = { 1.0, 0.0, px, 0.0, 1.0, py, 0.0, 0.0, 1.0 }; /* aftranslate33 */
    const double rot[9] = { c, -s, 0.0, s, c, 0.0, 0.0, 0.0, 1.0 }; /* afmat33 (rotate22_ c s) */
    mul3x3x3(transl, rot, out);
}

void orc_cos_sin(int64_t n, const double *rot, double *cos_out, double *sin_out)
{
    for (int64_t i = 0; i < n; ++i) {
        cos_out[i] = cos(rot[i]);
        sin_out[i] = sin(rot[i]);
    }
}

/* ------------------------------------------------------------------ */
/* Physics.Contact.ConvexHull (shapes/src/Physics/Contact/ConvexHull.hs) */
/* ------------------------------------------------------------------ */

typedef struct {
    int n;                 /* _hullVertexCount */
    const double *x, *y;   /* _hullVertices (world) */
    const double *nx, *ny; /* _hullEdgeNormals (world, unit) */
    const int32_t *emin, *emax; /* _hullExtents (frozen at construction) */
} hull_t;

static inline int next_index(int n, int i) { return i < n - 1 ? i + 1 : 0; } /* :228-230 */
static inline int prev_index(int n, int i) { return i > 0 ? i - 1 : n - 1; } /* :232-234 */

static inline v2 hull_vertex(const hull_t *h, int i) { v2 r = { h->x[i], h->y[i] }; return r; }
static inline v2 hull_normal(const hull_t *h, int i) { v2 r = { h->nx[i], h->ny[i] }; return r; }

/* unitEdgeNormal (:218-226): normalize (clockwise (v_next - v_i)) */
static inline v2 unit_edge_normal(const double *x, const double *y, int n, int i)
{
    int j = next_index(n, i);
    v2 v = { x[i], y[i] }, v1 = { x[j], y[j] };
    return normalize2(clockwise2(sub2(v1, v)));
}

typedef struct { int min_i, max_i; double min_v, max_v; } extent_t;

/* extentAlong' / extentAlong (:81-100): fold over vertices in index order,
 * strict < / > so the first minimum / first maximum wins. */
static extent_t extent_along(const double *x, const double *y, int n, v2 dir)
{
    extent_t e;
    v2 p0 = { x[0], y[0] };
    double d0 = dot2(p0, dir); /* distanceAlong: dir `afdot'` center = dotV2 center dir (:76-79) */
    e.min_i = e.max_i = 0;
    e.min_v = e.max_v = d0;
    for (int k = 1; k < n; ++k) {
        v2 p = { x[k], y[k] };
        double d = dot2(p, dir);
        if (d < e.min_v) { e.min_v = d; e.min_i = k; }
        if (d > e.max_v) { e.max_v = d; e.max_i = k; }
    }
    return e;
}

void orc_hull_extents(int64_t n_slots, const int32_t *vert_offset,
                      const double *local_x, const double *local_y,
                      int32_t *ext_min, int32_t *ext_max)
{
    /* listToHull (:151-167): edgeNormals from the local vertices, then
     * extents = fmap (extentIndices . extentAlong hull) edgeNormals. */
    for (int64_t s = 0; s < n_slots; ++s) {
        int32_t o = vert_offset[s];
        int n = vert_offset[s + 1] - o;
        for (int e = 0; e < n; ++e) {
            v2 dir = unit_edge_normal(local_x + o, local_y + o, n, e);
            extent_t ex = extent_along(local_x + o, local_y + o, n, dir);
            ext_min[o + e] = ex.min_i;
            ext_max[o + e] = ex.max_i;
        }
    }
}

void orc_move_shapes(int64_t n_slots, const uint8_t *alive, const int32_t *vert_offset,
                     const double *local_x, const double *local_y,
                     const double *pos_x, const double *pos_y,
                     const double *cos_rot, const double *sin_rot,
                     double *world_x, double *world_y,
                     double *normal_x, double *normal_y)
{
    /* moveShape (World.hs:132-134): setShapeTransform shape (transform (_physObjTransform obj))
     * setHullTransform (ConvexHull.hs:184-195): vertices = fmap fromLocalSpace local;
     * edgeNormals = unitEdgeNormal over the NEW vertices. */
    for (int64_t s = 0; s < n_slots; ++s) {
        if (alive && !alive[s]) continue;
        int32_t o = vert_offset[s];
        int n = vert_offset[s + 1] - o;
        double t[9];
        to_transform(pos_x[s], pos_y[s], cos_rot[s], sin_rot[s], t);
        for (int k = 0; k < n; ++k) {
            v2 l = { local_x[o + k], local_y[o + k] };
            v2 w = afmul(t, l);
            world_x[o + k] = w.x;
            world_y[o + k] = w.y;
        }
        for (int e = 0; e < n; ++e) {
            v2 nn = unit_edge_normal(world_x + o, world_y + o, n, e);
            normal_x[o + e] = nn.x;
            normal_y[o + e] = nn.y;
        }
    }
}

/* ------------------------------------------------------------------ */
/* Physics.Broadphase.Aabb (shapes/src/Physics/Broadphase/Aabb.hs)     */
/* ------------------------------------------------------------------ */

void orc_aabbs(int64_t n_slots, const uint8_t *alive, const int32_t *vert_offset,
               const double *world_x, const double *world_y,
               double *min_x, double *max_x, double *min_y, double *max_y)
{
    /* hullToAabb = foldl1 mergeAabb (toAabb_ <$> vertices) (:81-84);
     * mergeRange (Bounds a b) (Bounds c d): min = if a < c then a else c,
     * max = if b > d then b else d (:104-110); accumulator on the left. */
    for (int64_t s = 0; s < n_slots; ++s) {
        if (alive && !alive[s]) continue;
        int32_t o = vert_offset[s];
        int n = vert_offset[s + 1] - o;
        double ax0 = world_x[o], ax1 = world_x[o], ay0 = world_y[o], ay1 = world_y[o];
        for (int k = 1; k < n; ++k) {
            double cx = world_x[o + k], cy = world_y[o + k];
            ax0 = (ax0 < cx) ? ax0 : cx;
            ax1 = (ax1 > cx) ? ax1 : cx;
            ay0 = (ay0 < cy) ? ay0 : cy;
            ay1 = (ay1 > cy) ? ay1 : cy;
        }
        min_x[s] = ax0; max_x[s] = ax1; min_y[s] = ay0; max_y[s] = ay1;
    }
}

void orc_is_static(int64_t n_slots, const double *inv_lin, const double *inv_rot,
                   uint8_t *is_static)
{
    /* isStatic = (== InvMass2 0.0 0.0) (Constraint.hs:123-125) */
    for (int64_t s = 0; s < n_slots; ++s)
        is_static[s] = (inv_lin[s] == 0.0 && inv_rot[s] == 0.0) ? 1 : 0;
}

/* boundsOverlap (Aabb.hs:69-72) */
static inline int bounds_overlap(double a, double b, double c, double d)
{
    return !((c > b) || (d < a));
}

/* aabbCheck (Aabb.hs:75-78) with box A = slot i, box B = slot j */
static inline int aabb_check(const double *min_x, const double *max_x,
                             const double *min_y, const double *max_y, int64_t i, int64_t j)
{
    return bounds_overlap(min_x[i], max_x[i], min_x[j], max_x[j]) &&
           bounds_overlap(min_y[i], max_y[i], min_y[j], max_y[j]);
}

int64_t orc_unordered_pairs(int64_t n, int64_t cap, int32_t *xs, int32_t *ys)
{
    /* unorderedPairs (Aabb.hs:155-163): (n-1,n-2), (n-1,n-3), ..., (1,0) */
    int64_t k = 0;
    if (n < 2) return 0;
    for (int64_t x = n - 1; x >= 1; --x)
        for (int64_t y = x - 1; y >= 0; --y) {
            if (k < cap) { xs[k] = (int32_t)x; ys[k] = (int32_t)y; }
            ++k;
        }
    return k;
}

int64_t orc_culled_keys_aabb(int64_t n_slots, const uint8_t *alive,
                             const double *min_x, const double *max_x,
                             const double *min_y, const double *max_y,
                             const uint8_t *is_static,
                             int64_t cap, int32_t *pair_i, int32_t *pair_j)
{
    /* toTaggedAabbs (Aabb.hs:136-146): filled slots in ascending order. */
    int32_t *filled = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_slots > 0 ? n_slots : 1));
    int64_t m = 0;
    for (int64_t s = 0; s < n_slots; ++s)
        if (!alive || alive[s]) filled[m++] = (int32_t)s;
    /* culledKeys (:168-183): unorderedPairs (length taggedAabbs), keep
     * not (isStaticA && isStaticB) && aabbCheck a b, emit world keys (i', j'). */
    int64_t k = 0;
    for (int64_t x = m - 1; x >= 1; --x) {
        int32_t i = filled[x];
        for (int64_t y = x - 1; y >= 0; --y) {
            int32_t j = filled[y];
            if (!(is_static[i] && is_static[j]) && aabb_check(min_x, max_x, min_y, max_y, i, j)) {
                if (k < cap) { pair_i[k] = i; pair_j[k] = j; }
                ++k;
            }
        }
    }
    free(filled);
    return k;
}

typedef struct { int64_t cell; int32_t key; } cellref_t;

static int cmp_cellref(const void *pa, const void *pb)
{
    const cellref_t *a = (const cellref_t *)pa, *b = (const cellref_t *)pb;
    if (a->cell != b->cell) return a->cell < b->cell ? -1 : 1;
    if (a->key != b->key) return a->key < b->key ? -1 : 1;
    return 0;
}

static int cmp_pair_desc(const void *pa, const void *pb)
{
    uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
    return a > b ? -1 : (a < b ? 1 : 0);
}

/* sort packed (i<<32|j) descending, optionally drop duplicates, unpack */
static int64_t finish_pairs(uint64_t *packed, int64_t n, int uniq,
                            int64_t cap, int32_t *pair_i, int32_t *pair_j)
{
    qsort(packed, (size_t)n, sizeof(uint64_t), cmp_pair_desc);
    int64_t k = 0;
    for (int64_t t = 0; t < n; ++t) {
        if (uniq && t > 0 && packed[t] == packed[t - 1]) continue;
        if (k < cap) {
            pair_i[k] = (int32_t)(packed[t] >> 32);
            pair_j[k] = (int32_t)(packed[t] & 0xffffffffu);
        }
        ++k;
    }
    return k;
}

typedef struct { uint64_t *v; int64_t n, cap; } u64vec;
static void u64vec_push(u64vec *p, uint64_t x)
{
    if (p->n == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 1024;
        p->v = (uint64_t *)realloc(p->v, sizeof(uint64_t) * (size_t)p->cap);
    }
    p->v[p->n++] = x;
}

int64_t orc_culled_keys_grid(int64_t n_slots, const uint8_t *alive,
                             const double *min_x, const double *max_x,
                             const double *min_y, const double *max_y,
                             const uint8_t *is_static,
                             int32_t grid_len_x, double grid_unit_x, double grid_origin_x,
                             int32_t grid_len_y, double grid_unit_y, double grid_origin_y,
                             int64_t cap, int32_t *pair_i, int32_t *pair_j)
{
    (void)grid_len_y; /* flattenIndex' uses only the x axis length (Grid.hs:116-118) */
    /* fromTaggedAabbs (Grid.hs:102-110): insert every box into every cell of
     * boxIndices (:132-141): [axialIndex min .. axialIndex max] per axis,
     * axialIndex v = floor ((v - origin) / unit) (:127-129), flat = x + y*len. */
    cellref_t *refs = NULL;
    int64_t n_refs = 0, refs_cap = 0;
    for (int64_t s = 0; s < n_slots; ++s) {
        if (alive && !alive[s]) continue;
        int64_t x0 = (int64_t)floor((min_x[s] - grid_origin_x) / grid_unit_x);
        int64_t x1 = (int64_t)floor((max_x[s] - grid_origin_x) / grid_unit_x);
        int64_t y0 = (int64_t)floor((min_y[s] - grid_origin_y) / grid_unit_y);
        int64_t y1 = (int64_t)floor((max_y[s] - grid_origin_y) / grid_unit_y);
        for (int64_t x = x0; x <= x1; ++x)
            for (int64_t y = y0; y <= y1; ++y) {
                if (n_refs == refs_cap) {
                    refs_cap = refs_cap ? refs_cap * 2 : 4096;
                    refs = (cellref_t *)realloc(refs, sizeof(cellref_t) * (size_t)refs_cap);
                }
                refs[n_refs].cell = x + y * (int64_t)grid_len_x;
                refs[n_refs].key = (int32_t)s;
                ++n_refs;
            }
    }
    /* the inner IntMap keeps one entry per key per cell (flat-index aliasing
     * can present the same key twice) */
    qsort(refs, (size_t)n_refs, sizeof(cellref_t), cmp_cellref);
    int64_t n_uniq = 0;
    for (int64_t t = 0; t < n_refs; ++t)
        if (t == 0 || refs[t].cell != refs[t - 1].cell || refs[t].key != refs[t - 1].key)
            refs[n_uniq++] = refs[t];
    n_refs = n_uniq;
    u64vec found = { 0, 0, 0 };
    /* culledKeys' (:80-86): allPairs over IM.toDescList square, i.e. (a, b) with a > b;
     * skip static/static, keep aabbCheck boxA boxB. */
    int64_t lo = 0;
    while (lo < n_refs) {
        int64_t hi = lo;
        while (hi < n_refs && refs[hi].cell == refs[lo].cell) ++hi;
        for (int64_t p = hi - 1; p > lo; --p)
            for (int64_t q = p - 1; q >= lo; --q) {
                int32_t a = refs[p].key, b = refs[q].key;
                if (is_static[a] && is_static[b]) continue;
                if (aabb_check(min_x, max_x, min_y, max_y, a, b))
                    u64vec_push(&found, ((uint64_t)(uint32_t)a << 32) | (uint32_t)b);
            }
        lo = hi;
    }
    /* culledKeys (:74-78): concat, sortBy descending, uniq */
    int64_t k = finish_pairs(found.v, found.n, 1, cap, pair_i, pair_j);
    free(found.v);
    free(refs);
    return k;
}

static const double *g_sort_key;
static int cmp_by_key(const void *pa, const void *pb)
{
    double a = g_sort_key[*(const int32_t *)pa], b = g_sort_key[*(const int32_t *)pb];
    if (a < b) return -1;
    if (a > b) return 1;
    int32_t ia = *(const int32_t *)pa, ib = *(const int32_t *)pb;
    return ia < ib ? -1 : (ia > ib ? 1 : 0);
}

int64_t orc_culled_keys_sweep(int64_t n_slots, const uint8_t *alive,
                              const double *min_x, const double *max_x,
                              const double *min_y, const double *max_y,
                              const uint8_t *is_static,
                              int64_t cap, int32_t *pair_i, int32_t *pair_j)
{
    /* Same predicate as Aabb.culledKeys; shapes with any non-finite bound are
     * compared against everything (NaN bounds "overlap" under boundsOverlap). */
    int32_t *tame = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_slots > 0 ? n_slots : 1));
    int32_t *wild = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_slots > 0 ? n_slots : 1));
    int64_t n_tame = 0, n_wild = 0;
    for (int64_t s = 0; s < n_slots; ++s) {
        if (alive && !alive[s]) continue;
        if (isfinite(min_x[s]) && isfinite(max_x[s]) && isfinite(min_y[s]) && isfinite(max_y[s]))
            tame[n_tame++] = (int32_t)s;
        else
            wild[n_wild++] = (int32_t)s;
    }
    g_sort_key = min_x;
    qsort(tame, (size_t)n_tame, sizeof(int32_t), cmp_by_key);
    u64vec found = { 0, 0, 0 };
    for (int64_t p = 0; p < n_tame; ++p) {
        int32_t a = tame[p];
        for (int64_t q = p + 1; q < n_tame && !(min_x[tame[q]] > max_x[a]); ++q) {
            int32_t b = tame[q];
            int32_t i = a > b ? a : b, j = a > b ? b : a;
            if (is_static[i] && is_static[j]) continue;
            if (aabb_check(min_x, max_x, min_y, max_y, i, j))
                u64vec_push(&found, ((uint64_t)(uint32_t)i << 32) | (uint32_t)j);
        }
    }
    for (int64_t w = 0; w < n_wild; ++w) {
        int32_t a = wild[w];
        for (int64_t s = 0; s < n_slots; ++s) {
            if (alive && !alive[s]) continue;
            int32_t b = (int32_t)s;
            if (a == b) continue;
            int b_wild = !(isfinite(min_x[b]) && isfinite(max_x[b]) && isfinite(min_y[b]) && isfinite(max_y[b]));
            if (b_wild && b > a) continue; /* wild/wild pair handled once, from the larger key */
            int32_t i = a > b ? a : b, j = a > b ? b : a;
            if (is_static[i] && is_static[j]) continue;
            if (aabb_check(min_x, max_x, min_y, max_y, i, j))
                u64vec_push(&found, ((uint64_t)(uint32_t)i << 32) | (uint32_t)j);
        }
    }
    int64_t k = finish_pairs(found.v, found.n, 0, cap, pair_i, pair_j);
    free(found.v);
    free(tame);
    free(wild);
    return k;
}

/* ------------------------------------------------------------------ */
/* Physics.Contact.SAT (shapes/src/Physics/Contact/SAT.hs)             */
/* ------------------------------------------------------------------ */

typedef struct {
    int separated; /* 1 = Separated edge, 0 = MinOverlap */
    int edge;      /* _overlapEdge index (in the receiving hull) */
    double depth;  /* _overlapDepth */
    int pen;       /* _overlapPenetrator index (in the penetrating hull) */
} sat_t;

/* overlap (SAT.hs:103-117) */
static sat_t sat_overlap(const hull_t *s_edge, int e, const hull_t *s_pen)
{
    sat_t r;
    v2 dir = hull_normal(s_edge, e);
    /* extentAlongSelf (ConvexHull.hs:111-118): the two cached vertices only */
    double s_min = dot2(hull_vertex(s_edge, s_edge->emin[e]), dir);
    double s_max = dot2(hull_vertex(s_edge, s_edge->emax[e]), dir);
    extent_t p = extent_along(s_pen->x, s_pen->y, s_pen->n, dir);
    /* overlapTest (SAT.hs:74-83): not (c > b || d < a); overlapAmount (:86-96): edge - penetrator */
    int ov = !((p.min_v > s_max) || (p.max_v < s_min));
    r.separated = !ov;
    r.edge = e;
    r.depth = s_max - p.min_v;
    r.pen = p.min_i;
    return r;
}

/* minOverlap / minOverlap' (SAT.hs:121-143): foldl1 over edges in index order;
 * the first Separated wins, else strictly smaller depth replaces. */
static sat_t sat_min_overlap(const hull_t *s_edge, const hull_t *s_pen)
{
    sat_t acc = sat_overlap(s_edge, 0, s_pen);
    for (int e = 1; e < s_edge->n; ++e) {
        sat_t o = sat_overlap(s_edge, e, s_pen);
        if (acc.separated) continue;
        if (o.separated) { acc = o; continue; }
        if (o.depth < acc.depth) acc = o;
    }
    return acc;
}

typedef struct { v2 p; int idx; } nb_t;      /* the parts of a Neighborhood clipping touches */
typedef struct { v2 p; v2 n; } line2_t;      /* Line2 (Linear.hs:229-231) */

/* intersect2 (Linear.hs:244-251) with invM2x2 (:194-199) and mul2x2c */
static v2 intersect2(line2_t l0, line2_t l1)
{
    double n0 = l0.n.x, n1 = l0.n.y, n2 = l1.n.x, n3 = l1.n.y;
    double b0 = dot2(l0.p, l0.n);
    double b1 = dot2(l1.p, l1.n);
    double det = (n0 * n3) - (n1 * n2);
    double inv_det = 1.0 / det;
    /* invDet `smulM2x2` M2x2 d (-b) (-c) a, each entry multiplied */
    double m00 = n3 * inv_det, m01 = (-n1) * inv_det;
    double m10 = (-n2) * inv_det, m11 = n0 * inv_det;
    v2 r = { (m00 * b0) + (m01 * b1), (m10 * b0) + (m11 * b1) };
    return r;
}

enum { CLIP_LEFT, CLIP_RIGHT, CLIP_BOTH, CLIP_NONE };

/* clipSegment (Linear.hs:327-343) */
static int clip_segment(line2_t boundary, line2_t incident, v2 a, v2 b, v2 *c_out)
{
    v2 c = intersect2(boundary, incident);
    v2 n = boundary.n;
    double a1 = dot2(a, n), b1 = dot2(b, n), c1 = dot2(c, n);
    *c_out = c;
    if (a1 < c1) return (b1 < c1) ? CLIP_BOTH : CLIP_LEFT;
    if (b1 < c1) return CLIP_RIGHT;
    return CLIP_NONE;
}

/* lApplyClip' (Linear.hs:303-322): replace the clipped endpoint's point, keep its index */
static int l_apply_clip(int res, v2 c, nb_t seg[2])
{
    switch (res) {
    case CLIP_LEFT: seg[0].p = c; return 1;
    case CLIP_RIGHT: seg[1].p = c; return 1;
    case CLIP_BOTH: return 0;
    default: return 1;
    }
}

/* clipEdge (SAT.hs:190-218). Returns the number of manifold points (0 = Nothing). */
static int clip_edge(nb_t aa, nb_t bb, v2 n, nb_t inc0, nb_t inc1, nb_t out[2])
{
    v2 a = aa.p, b = bb.p, c = inc0.p, d = inc1.p;
    line2_t a_bound = { a, sub2(b, a) };  /* perpLine2 a b (Linear.hs:238-241) */
    line2_t b_bound = { b, sub2(a, b) };  /* perpLine2 b a */
    line2_t ab_bound = { a, neg2(n) };    /* Line2 a (negateV2 n) */
    line2_t cd = { c, clockwise2(sub2(d, c)) }; /* toLine2 c d (Linear.hs:233-236), unclipped endpoints */
    nb_t seg[2] = { inc0, inc1 };
    v2 x;
    int r = clip_segment(a_bound, cd, seg[0].p, seg[1].p, &x);
    if (!l_apply_clip(r, x, seg)) return 0;
    r = clip_segment(b_bound, cd, seg[0].p, seg[1].p, &x);
    if (!l_apply_clip(r, x, seg)) return 0;
    r = clip_segment(ab_bound, cd, seg[0].p, seg[1].p, &x);
    /* applyClip'' (Linear.hs:285-292): REMOVES the out-of-bounds endpoint */
    switch (r) {
    case CLIP_LEFT: out[0] = seg[1]; return 1;
    case CLIP_RIGHT: out[0] = seg[0]; return 1;
    case CLIP_BOTH: return 0;
    default: out[0] = seg[0]; out[1] = seg[1]; return 2;
    }
}

typedef struct {
    int n;         /* number of flattened contacts: 0, 1 or 2 */
    int flip;      /* 0 = Same (a penetrated by b), 1 = Flip */
    int edge;      /* penetrated edge index (in the penetrated hull) */
    v2 normal;     /* unit normal of the penetrated edge */
    int pen[2];    /* penetrating feature index, descending */
    v2 center[2];
    double depth[2];
} manifold_t;

/* contact / contactDebug (SAT.hs:238-258), contact_ (:261-267), then
 * flattenContactResult (HullVsHull.hs:54-76). a = shape with the larger key. */
static manifold_t hull_vs_hull(const hull_t *a, const hull_t *b)
{
    manifold_t m;
    memset(&m, 0, sizeof m);
    sat_t ab = sat_min_overlap(a, b);
    sat_t ba = sat_min_overlap(b, a);
    /* eitherBranchBoth (Utils.hs:230-235): first Left wins; else depth_ab < depth_ba ? Same : Flip */
    if (ab.separated) return m;
    if (ba.separated) return m;
    const hull_t *s_edge, *s_pen;
    sat_t ov;
    if (ab.depth < ba.depth) { m.flip = 0; ov = ab; s_edge = a; s_pen = b; }
    else { m.flip = 1; ov = ba; s_edge = b; s_pen = a; }

    v2 n = hull_normal(s_edge, ov.edge); /* overlapNormal (:98-100) */
    /* penetratedEdge (:169-171) */
    int e0 = ov.edge, e1 = next_index(s_edge->n, e0);
    nb_t aa = { hull_vertex(s_edge, e0), e0 }, bb = { hull_vertex(s_edge, e1), e1 };
    /* penetratingEdge (:152-166) */
    int ib = ov.pen, ic = next_index(s_pen->n, ib), ia = prev_index(s_pen->n, ib);
    v2 pa = hull_vertex(s_pen, ia), pb = hull_vertex(s_pen, ib), pc = hull_vertex(s_pen, ic);
    double abn = fabs(dot2(sub2(pb, pa), n));
    double bcn = fabs(dot2(sub2(pc, pb), n));
    nb_t inc0, inc1;
    if (bcn < abn) { inc0.p = pb; inc0.idx = ib; inc1.p = pc; inc1.idx = ic; }
    else { inc0.p = pa; inc0.idx = ia; inc1.p = pb; inc1.idx = ib; }

    nb_t pts[2];
    int np = clip_edge(aa, bb, n, inc0, inc1, pts);
    if (np == 0) return m; /* clipEdge = Nothing => contact = Nothing */
    /* flattenContactPoints (:181-187): descending feature index */
    if (np == 2 && !(pts[0].idx > pts[1].idx)) { nb_t t = pts[0]; pts[0] = pts[1]; pts[1] = t; }
    m.n = np;
    m.edge = ov.edge;
    m.normal = n;
    for (int k = 0; k < np; ++k) {
        m.pen[k] = pts[k].idx;
        m.center[k] = pts[k].p;
        /* contactDepth_ (HullVsHull.hs:30-37): f v - f p, f = afdot' n */
        m.depth[k] = dot2(aa.p, n) - dot2(pts[k].p, n);
    }
    return m;
}

/* ------------------------------------------------------------------ */
/* Circles: Physics.Contact.Circle, CircleVsHull, GJK                   */
/* ------------------------------------------------------------------ */

/* Circle.contact circleA circleB (shapes/src/Physics/Contact/Circle.hs:30-53): A is the penetratee,
 * the normal points out of A.  The flattened key is (0, 0), Same (Contact.hs:25-29). */
static manifold_t circle_vs_circle(v2 a, double ra, v2 b, double rb)
{
    manifold_t m;
    memset(&m, 0, sizeof m);
    v2 ab = sub2(b, a);                                /* diffP2 b a */
    double rab = ra + rb;
    double ab_sq = (ab.x * ab.x) + (ab.y * ab.y);       /* sqLengthV2 */
    if (!(rab * rab >= ab_sq)) return m;
    double ab_len = sqrt(ab_sq);
    v2 abn = { ab.x / ab_len, ab.y / ab_len };         /* abLength `sdivV2` ab */
    v2 a1 = { (abn.x * ra) + a.x, (abn.y * ra) + a.y };        /* (ra `smulV2` abN) `vplusP2` a */
    double nrb = -rb;
    v2 b1 = { (abn.x * nrb) + b.x, (abn.y * nrb) + b.y };      /* ((-rb) `smulV2` abN) `vplusP2` b */
    m.n = 1;
    m.flip = 0;
    m.edge = 0;
    m.pen[0] = 0;
    m.normal = abn;
    m.center[0].x = (a1.x + b1.x) / 2.0;               /* midpointP2: 2 `sdivV2` (v0 `plusV2` v1) */
    m.center[0].y = (a1.y + b1.y) / 2.0;
    m.depth[0] = (ra + rb) - ab_len;
    return m;
}

/* support (ConvexHull.hs:124-128): first maximum of dir . v */
static int hull_support(const hull_t *h, v2 dir)
{
    int best = 0;
    double bd = dot2(hull_vertex(h, 0), dir);
    for (int k = 1; k < h->n; ++k) {
        double d = dot2(hull_vertex(h, k), dir);
        if (d > bd) { bd = d; best = k; }
    }
    return best;
}

static inline int same_direction(v2 a, v2 b) { return dot2(a, b) > 0.0; } /* GJK.hs:138-139 */
/* crossV2V2 (Linear.hs:135-139) */
static inline v2 cross_v2v2(v2 a, v2 b, v2 c)
{
    double abz = (a.x * b.y) - (a.y * b.x);
    v2 r = { -(abz * c.y), abz * c.x };
    return r;
}

#define GJK_MAX_ITER 64 /* the reference loops until a support vertex repeats; bounded here */

/* closestSimplex hull origin (GJK.hs:52-69).  Returns the simplex size (1, 2; 3 = encloses the
 * target; 0 = iteration cap hit) and its vertex indices, most recently added first. */
static int gjk_closest_simplex(const hull_t *h, v2 origin, int idx[3])
{
    int n = 1;
    idx[0] = 0;
    v2 d = sub2(origin, hull_vertex(h, 0));
    for (int it = 0; it < GJK_MAX_ITER; ++it) {
        int aa = hull_support(h, d);
        /* extendSimplex (GJK.hs:71-90): a repeated vertex ends the search */
        if (n == 1) { if (idx[0] == aa) return 1; idx[1] = idx[0]; idx[0] = aa; n = 2; }
        else { if (idx[0] == aa || idx[1] == aa) return 2; idx[2] = idx[1]; idx[1] = idx[0]; idx[0] = aa; n = 3; }
        v2 a = hull_vertex(h, idx[0]), b = hull_vertex(h, idx[1]);
        v2 ab = sub2(b, a), ao = sub2(origin, a);
        if (n == 2) {
            /* shiftSimplex2 (GJK.hs:98-112) */
            if (same_direction(ab, ao)) d = cross_v2v2(ab, ao, ab);
            else { n = 1; d = ao; }
        } else {
            /* shiftSimplex3 (GJK.hs:114-136) */
            v2 c = hull_vertex(h, idx[2]);
            v2 ac = sub2(c, a);
            double abc = cross2(ab, ac);
            v2 abcac = { -(abc * ac.y), abc * ac.x };   /* abc `zcrossV2` ac (Linear.hs:126-129) */
            v2 ababc = { ab.y * abc, -(ab.x * abc) };   /* ab `crosszV2` abc (Linear.hs:121-124) */
            int star = 0;
            if (same_direction(abcac, ao)) {
                if (same_direction(ac, ao)) { idx[1] = idx[2]; n = 2; d = cross_v2v2(ac, ao, ac); }
                else star = 1;
            } else if (same_direction(ababc, ao)) star = 1;
            else return 3; /* simplex encloses the origin */
            if (star) {
                if (same_direction(ab, ao)) { n = 2; d = cross_v2v2(ab, ao, ab); }
                else { n = 1; d = ao; }
            }
        }
    }
    return 0;
}

/* CircleVsHull.generateContacts (CircleVsHull.hs:18-69): the circle is always the penetrator.
 * flip = 0: (CircleShape a, HullShape b) -> ((0, hullFeature), Same contact);
 * flip = 1: (HullShape a, CircleShape b) -> ((hullFeature, 0), Flip contact)  (Contact.hs:30-39). */
static manifold_t circle_vs_hull(v2 center, double r, const hull_t *h, int flip)
{
    manifold_t m;
    memset(&m, 0, sizeof m);
    int idx[3] = { 0, 0, 0 };
    int n = gjk_closest_simplex(h, center, idx);
    if (n != 1 && n != 2) return m;            /* Simplex3' (deep overlap) => Nothing (CircleVsHull.hs:29) */
    v2 a = hull_vertex(h, idx[0]);
    if (n == 2) {
        /* closestAlong (CircleVsHull.hs:60-69); the feature is the most recently added vertex */
        v2 b = hull_vertex(h, idx[1]);
        v2 ao = sub2(center, a), ab = sub2(b, a);
        v2 abn = normalize2(ab);
        double s = dot2(ao, abn);
        v2 along = { abn.x * s, abn.y * s };
        a.x = along.x + a.x; a.y = along.y + a.y;
    }
    /* processSimplex_ (CircleVsHull.hs:42-58) */
    double sq_radius = r * r;
    v2 ab = sub2(center, a);
    double ab_sq = (ab.x * ab.x) + (ab.y * ab.y);
    if (sq_radius < ab_sq) return m;
    double ab_len = sqrt(ab_sq);
    v2 normal = { ab.x / ab_len, ab.y / ab_len };
    m.n = 1;
    m.flip = flip;
    m.edge = 0;
    m.pen[0] = idx[0];
    m.normal = neg2(normal);
    m.center[0] = a;
    m.depth[0] = r - ab_len;
    return m;
}

void orc_move_circles(int64_t n_slots, const uint8_t *alive, const double *radius,
                      const double *pos_x, const double *pos_y,
                      const double *cos_rot, const double *sin_rot,
                      double *center_x, double *center_y)
{
    /* setCircleTransform (Circle.hs:55-59): Circle (fromLocalSpace zeroP2) radius */
    for (int64_t s = 0; s < n_slots; ++s) {
        if ((alive && !alive[s]) || !(radius[s] >= 0.0)) continue;
        double t[9];
        to_transform(pos_x[s], pos_y[s], cos_rot[s], sin_rot[s], t);
        v2 zero = { 0.0, 0.0 };
        v2 c = afmul(t, zero);
        center_x[s] = c.x; center_y[s] = c.y;
    }
}

void orc_aabbs_circles(int64_t n_slots, const uint8_t *alive, const double *radius,
                       const double *center_x, const double *center_y,
                       double *min_x, double *max_x, double *min_y, double *max_y)
{
    /* circleToAabb (Aabb.hs:86-88) */
    for (int64_t s = 0; s < n_slots; ++s) {
        if ((alive && !alive[s]) || !(radius[s] >= 0.0)) continue;
        double x = center_x[s], y = center_y[s], r = radius[s];
        min_x[s] = x - r; max_x[s] = x + r; min_y[s] = y - r; max_y[s] = y + r;
    }
}

/* ------------------------------------------------------------------ */
/* Constraint generators                                               */
/* ------------------------------------------------------------------ */

/* NonPenetration.jacobian (NonPenetration.hs:34-43): a is penetrated by b */
static void np_jacobian(v2 n, v2 p, v2 xa, v2 xb, double j[6])
{
    v2 nn = neg2(n);
    j[0] = nn.x; j[1] = nn.y; j[2] = cross2(sub2(xa, p), n);
    j[3] = n.x;  j[4] = n.y;  j[5] = cross2(sub2(p, xb), n);
}

/* Friction.jacobian (Friction.hs:31-44) */
static void f_jacobian(v2 n, v2 p, v2 xa, v2 xb, double j[6])
{
    v2 tb = clockwise2(n);
    v2 ta = neg2(tb);
    j[0] = ta.x; j[1] = ta.y; j[2] = cross2(sub2(p, xa), ta);
    j[3] = tb.x; j[4] = tb.y; j[5] = cross2(sub2(p, xb), tb);
}

/* flip3v3 (Linear.hs:149-151) */
static void flip3v3(double j[6])
{
    double t;
    t = j[0]; j[0] = j[3]; j[3] = t;
    t = j[1]; j[1] = j[4]; j[4] = t;
    t = j[2]; j[2] = j[5]; j[5] = t;
}

/* effMassM2 (Constraint.hs:173-179): (j `vmulDiag6` im) `dotV6` j, left fold */
static double eff_mass(const double j[6], const double im[6])
{
    double acc = (j[0] * im[0]) * j[0];
    for (int k = 1; k < 6; ++k) acc = acc + ((j[k] * im[k]) * j[k]);
    return acc;
}

#define PUT(arr, row, val) do { if (out->arr) out->arr[row] = (val); } while (0)

int64_t orc_contacts_shapes(int64_t n_pairs, const int32_t *pair_i, const int32_t *pair_j,
                     const int32_t *vert_offset,
                     const double *world_x, const double *world_y,
                     const double *normal_x, const double *normal_y,
                     const int32_t *ext_min, const int32_t *ext_max,
                     const double *radius, const double *circle_x, const double *circle_y,
                     const double *pos_x, const double *pos_y,
                     const double *inv_lin, const double *inv_rot,
                     double dt, double baumgarte, double slop,
                     orc_contacts_out *out)
{
    int64_t row = 0;
    for (int64_t t = 0; t < n_pairs; ++t) {
        int32_t i = pair_i[t], j = pair_j[t];
        hull_t ha, hb;
        int32_t oa = vert_offset[i], ob = vert_offset[j];
        ha.n = vert_offset[i + 1] - oa;
        ha.x = world_x + oa; ha.y = world_y + oa; ha.nx = normal_x + oa; ha.ny = normal_y + oa;
        ha.emin = ext_min + oa; ha.emax = ext_max + oa;
        hb.n = vert_offset[j + 1] - ob;
        hb.x = world_x + ob; hb.y = world_y + ob; hb.nx = normal_x + ob; hb.ny = normal_y + ob;
        hb.emin = ext_min + ob; hb.emax = ext_max + ob;

        /* keyedContacts (Constraints/Contact.hs:49-56) -> generateContacts
         * (Contact.hs:40, HullVsHull.hs:85-89) */
        /* generateContacts dispatch (shapes/src/Physics/Contact.hs:22-40) */
        const int ca = radius && radius[i] >= 0.0, cb = radius && radius[j] >= 0.0;
        manifold_t m;
        if (ca && cb) {
            v2 pa = { circle_x[i], circle_y[i] }, pb = { circle_x[j], circle_y[j] };
            m = circle_vs_circle(pa, radius[i], pb, radius[j]);          /* ((0,0), Same contact) */
        } else if (ca) {
            v2 pa = { circle_x[i], circle_y[i] };
            m = circle_vs_hull(pa, radius[i], &hb, 0);                   /* ((0, hullFeature), Same contact) */
        } else if (cb) {
            v2 pb = { circle_x[j], circle_y[j] };
            m = circle_vs_hull(pb, radius[j], &ha, 1);                   /* ((hullFeature, 0), Flip contact) */
        } else m = hull_vs_hull(&ha, &hb);
        v2 xi = { pos_x[i], pos_y[i] }, xj = { pos_x[j], pos_y[j] };
        for (int k = 0; k < m.n; ++k, ++row) {
            if (!out || row >= out->cap) continue;
            v2 n = m.normal, p = m.center[k];
            double d = m.depth[k];
            PUT(key_i, row, i);
            PUT(key_j, row, j);
            /* flipExtractPair fst (HullVsHull.hs:73-75, Utils.hs:184-186) */
            PUT(feat_a, row, m.flip ? m.pen[k] : m.edge);
            PUT(feat_b, row, m.flip ? m.edge : m.pen[k]);
            PUT(flip, row, (uint8_t)m.flip);
            PUT(normal_x, row, n.x); PUT(normal_y, row, n.y);
            PUT(center_x, row, p.x); PUT(center_y, row, p.y);
            PUT(depth, row, d);

            /* constraintGen (Constraints/Contact.hs:60-72); flipMap evaluates
             * the generator on (b, a) for Flip and flipExtract swaps the J
             * halves back (Utils.hs:175-177, 212-215; Constraint.hs:96-98). */
            double jn[6], jf[6];
            if (!m.flip) { np_jacobian(n, p, xi, xj, jn); f_jacobian(n, p, xi, xj, jf); }
            else {
                np_jacobian(n, p, xj, xi, jn); flip3v3(jn);
                f_jacobian(n, p, xj, xi, jf); flip3v3(jf);
            }
            /* baumgarte (NonPenetration.hs:48-55) */
            double bnp = (d > slop) ? (baumgarte / dt) * (slop - d) : 0.0;
            for (int q = 0; q < 6; ++q) { PUT(j_np[q], row, jn[q]); PUT(j_f[q], row, jf[q]); }
            PUT(b_np, row, bnp);
            PUT(b_f, row, 0.0);
            /* Restitution.constraintGen (Restitution.hs:21-31): radii from the UNflipped pair */
            v2 ra = sub2(p, xi), rb = sub2(p, xj);
            v2 rn = m.flip ? neg2(n) : n;
            PUT(ra_x, row, ra.x); PUT(ra_y, row, ra.y);
            PUT(rb_x, row, rb.x); PUT(rb_y, row, rb.y);
            PUT(rn_x, row, rn.x); PUT(rn_y, row, rn.y);
            /* effMassM2 with invMassM2 of the unflipped (i, j) (Constraint.hs:118-120, 138-140) */
            double im[6] = { inv_lin[i], inv_lin[i], inv_rot[i], inv_lin[j], inv_lin[j], inv_rot[j] };
            PUT(inv_eff_np, row, eff_mass(jn, im));
            PUT(inv_eff_f, row, eff_mass(jf, im));
        }
    }
    return row;
}


int64_t orc_contacts(int64_t n_pairs, const int32_t *pair_i, const int32_t *pair_j,
                     const int32_t *vert_offset,
                     const double *world_x, const double *world_y,
                     const double *normal_x, const double *normal_y,
                     const int32_t *ext_min, const int32_t *ext_max,
                     const double *pos_x, const double *pos_y,
                     const double *inv_lin, const double *inv_rot,
                     double dt, double baumgarte, double slop,
                     orc_contacts_out *out)
{
    return orc_contacts_shapes(n_pairs, pair_i, pair_j, vert_offset, world_x, world_y, normal_x, normal_y,
                               ext_min, ext_max, NULL, NULL, NULL, pos_x, pos_y, inv_lin, inv_rot,
                               dt, baumgarte, slop, out);
}

void orc_solve_constraint(const double *j6, double b, const double *inv_mass6, double *vel6)
{
    /* lagrangian2 (Constraint.hs:164-169): (-(j.v + b)) / effMass */
    double jv = j6[0] * vel6[0];
    for (int k = 1; k < 6; ++k) jv = jv + (j6[k] * vel6[k]);
    double mc = eff_mass(j6, inv_mass6);
    double lagr = (-(jv + b)) / mc;
    /* applyLagrangian2 (:197-204): v + im * (lagr * j) */
    for (int k = 0; k < 6; ++k) {
        double pc = j6[k] * lagr;           /* constraintImpulse2: l `smulV6` j */
        vel6[k] = vel6[k] + (pc * inv_mass6[k]); /* updateVelocity2_: v + (im vmulDiag6' pc) */
    }
}

/* ------------------------------------------------------------------ */
/* Warm start: descZipVector (Utils/Descending.hs:47-71)               */
/* ------------------------------------------------------------------ */

static int key_cmp(int32_t ai, int32_t aj, int32_t afa, int32_t afb,
                   int32_t bi, int32_t bj, int32_t bfa, int32_t bfb)
{
    /* derived Ord on ObjectFeatureKey: (_ofkObjKeys, _ofkFeatKeys) lexicographic */
    if (ai != bi) return ai < bi ? -1 : 1;
    if (aj != bj) return aj < bj ? -1 : 1;
    if (afa != bfa) return afa < bfa ? -1 : 1;
    if (afb != bfb) return afb < bfb ? -1 : 1;
    return 0;
}

void orc_warm_join(int64_t n_this, const int32_t *ti, const int32_t *tj, const int32_t *tfa, const int32_t *tfb,
                   int64_t n_that, const int32_t *pi, const int32_t *pj, const int32_t *pfa, const int32_t *pfb,
                   const double *that_np, const double *that_f,
                   double *out_np, double *out_f, uint8_t *out_hit)
{
    /* foldM f (0, accum0) these: that_i only ever moves forward */
    int64_t that_i = 0;
    for (int64_t k = 0; k < n_this; ++k) {
        int done = 0;
        while (!done) {
            if (that_i < n_that) {
                int c = key_cmp(ti[k], tj[k], tfa[k], tfb[k], pi[that_i], pj[that_i], pfa[that_i], pfb[that_i]);
                if (c < 0) { ++that_i; continue; }             /* thisKey < thatKey: keep looking */
                if (c == 0) {                                  /* accumBoth = useCache */
                    out_np[k] = that_np[that_i]; out_f[k] = that_f[that_i]; out_hit[k] = 1;
                    ++that_i;
                } else {                                       /* accumThis = newCache: ContactLagrangian 0 0 */
                    out_np[k] = 0.0; out_f[k] = 0.0; out_hit[k] = 0;
                }
            } else { out_np[k] = 0.0; out_f[k] = 0.0; out_hit[k] = 0; }
            done = 1;
        }
    }
}

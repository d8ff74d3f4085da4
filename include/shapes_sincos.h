/*
 * shapes_sincos.h -- one deterministic cos/sin for the host AND the device.
 *
 * Why it exists: the reference rotates hulls with libm's cos/sin (rotate22,
 * shapes/src/Physics/Linear.hs:353-357).  No device routine reproduces a given libm bit for bit,
 * so a world whose rotation lives in HBM (shapes_world_step) needs a
 * cos/sin the host can evaluate too.  This header is that function: plain IEEE binary64
 * multiplies/adds in a fixed order (no FMA, no tables, no libm calls), so gcc on the host
 * (-ffp-contract=off) and nvcc on the device produce the SAME bits.  A host engine that wants
 * its own moveShapes to agree bit-exactly with device-stepped worlds calls shapes_sincos()
 * (exported by the library, host code) instead of libm -- see INTEGRATION.md.
 *
 * Method (public-domain fdlibm scheme, restated): Cody-Waite reduction by pi/2 with a 33+33+33+53
 * bit split of pi/2 (the multiple fn is itself split so that every product is exact) carried as a
 * double-double remainder, then the degree-13 / degree-14 minimax kernels on [-pi/4, pi/4].
 * Error < 1 ulp for |x| <= 1e12 (tests/test_sincos.py measures it against libm).  Beyond 1e12 the multiple of
 * pi/2 no longer splits into two 20-bit halves and the "every product is exact" argument fails, so instead of a
 * silently inaccurate rotation |x| > 1e12 (or non-finite x) gives NaN -- visible in the very next frame (NaN bodies
 * take the exact big-shape path) instead of a slow drift.  A body reaches 1e12 rad only after ~1.6e9 revolutions.
 *
 * Host bit-parity needs un-fused arithmetic: the host functions carry GCC's optimize("fp-contract=off") attribute
 * (clang: FP_CONTRACT OFF pragma), so an includer built with contraction on -- GCC's default on targets with FMA,
 * e.g. aarch64 -- still gets separately rounded multiplies and adds.
 */
#ifndef SHAPES_SINCOS_H
#define SHAPES_SINCOS_H

#include <stdint.h>

#if defined(__CUDACC__)
#define SHAPES_SC_FN static __host__ __device__ __forceinline__
#define SHAPES_SC_NOFMA
#else
#include <string.h>
#if defined(__clang__)
#define SHAPES_SC_FN static inline
#define SHAPES_SC_NOFMA _Pragma("STDC FP_CONTRACT OFF")
#elif defined(__GNUC__)
#define SHAPES_SC_FN static inline __attribute__((optimize("fp-contract=off")))
#define SHAPES_SC_NOFMA
#else
#define SHAPES_SC_FN static inline
#define SHAPES_SC_NOFMA
#endif
#endif

#if defined(__CUDA_ARCH__)
#define SC_MUL(a, b) __dmul_rn((a), (b))
#define SC_ADD(a, b) __dadd_rn((a), (b))
#define SC_SUB(a, b) __dsub_rn((a), (b))
SHAPES_SC_FN uint64_t shapes_sc_bits(double x) { return (uint64_t)__double_as_longlong(x); }
SHAPES_SC_FN double shapes_sc_from_bits(uint64_t u) { return __longlong_as_double((long long)u); }
#else
#define SC_MUL(a, b) ((a) * (b))
#define SC_ADD(a, b) ((a) + (b))
#define SC_SUB(a, b) ((a) - (b))
SHAPES_SC_FN uint64_t shapes_sc_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
SHAPES_SC_FN double shapes_sc_from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#endif

/* sin on [-pi/4, pi/4] of the double-double x + y */
SHAPES_SC_FN double shapes_sc_ksin(double x, double y)
{
    SHAPES_SC_NOFMA
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = SC_MUL(x, x);
    const double v = SC_MUL(z, x);
    const double r = SC_ADD(S2, SC_MUL(z, SC_ADD(S3, SC_MUL(z, SC_ADD(S4, SC_MUL(z, SC_ADD(S5, SC_MUL(z, S6))))))));
    /* x - ((z*(y/2 - v*r) - y) - v*S1) */
    return SC_SUB(x, SC_SUB(SC_SUB(SC_MUL(z, SC_SUB(SC_MUL(0.5, y), SC_MUL(v, r))), y), SC_MUL(v, S1)));
}

/* cos on [-pi/4, pi/4] of the double-double x + y */
SHAPES_SC_FN double shapes_sc_kcos(double x, double y)
{
    SHAPES_SC_NOFMA
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = SC_MUL(x, x);
    const double r = SC_MUL(z, SC_ADD(C1, SC_MUL(z, SC_ADD(C2, SC_MUL(z, SC_ADD(C3, SC_MUL(z, SC_ADD(C4, SC_MUL(z, SC_ADD(C5, SC_MUL(z, C6)))))))))));
    const double zr_xy = SC_SUB(SC_MUL(z, r), SC_MUL(x, y));
    const uint64_t ax = shapes_sc_bits(x) & 0x7fffffffffffffffull;
    const uint32_t hx = (uint32_t)(ax >> 32);
    if (hx < 0x3fd33333u)                       /* |x| < 0.3 */
        return SC_SUB(1.0, SC_SUB(SC_MUL(0.5, z), zr_xy));
    /* 1 - qx is exact for qx ~ |x|/4 cut to its high word: keeps the subtraction from 1 error free */
    const double qx = (hx > 0x3fe90000u) ? 0.28125 : shapes_sc_from_bits((uint64_t)(hx - 0x00200000u) << 32);
    const double hz = SC_SUB(SC_MUL(0.5, z), qx);
    const double a = SC_SUB(1.0, qx);
    return SC_SUB(a, SC_SUB(hz, zr_xy));
}

/* s + e = a + b exactly (Knuth TwoSum, branch free) */
SHAPES_SC_FN double shapes_sc_two_sum(double a, double b, double *e)
{
    SHAPES_SC_NOFMA
    const double s = SC_ADD(a, b);
    const double bb = SC_SUB(s, a);
    *e = SC_ADD(SC_SUB(a, SC_SUB(s, bb)), SC_SUB(b, bb));
    return s;
}

/* cos(x), sin(x).  Same bits on host and device. */
SHAPES_SC_FN void shapes_sincos_inline(double x, double *cos_out, double *sin_out)
{
    SHAPES_SC_NOFMA
    const uint64_t ux = shapes_sc_bits(x);
    const uint64_t ax = ux & 0x7fffffffffffffffull;
    if (ax > 0x426d1a94a2000000ull) {           /* |x| > 1e12, inf, NaN: outside the proven range of the reduction */
        const double nan = shapes_sc_from_bits(0x7ff8000000000000ull);
        *cos_out = nan; *sin_out = nan;
        return;
    }
    double y0 = x, y1 = 0.0;
    int64_t n = 0;
    if (ax > 0x3fe921fb54442d18ull) {           /* |x| > pi/4 */
        const double INVPIO2 = 6.36619772367581382433e-01;
        const double PIO2_1 = 1.57079632673412561417e+00;   /* first 33 bits of pi/2 */
        const double PIO2_2 = 6.07710050630396597660e-11;   /* next 33 bits */
        const double PIO2_3 = 2.02226624871116645580e-21;   /* next 33 bits */
        const double PIO2_3T = 8.47842766036889956997e-32;  /* pi/2 - (PIO2_1 + PIO2_2 + PIO2_3) */
        const double MAGIC = 6755399441055744.0;             /* 1.5 * 2^52: adds and removes => round to nearest integer */
        const double t = SC_MUL(x, INVPIO2);
        const uint64_t at = shapes_sc_bits(t) & 0x7fffffffffffffffull;
        const double fn = (at < 0x4320000000000000ull) ? SC_SUB(SC_ADD(t, MAGIC), MAGIC) : t;   /* |t| < 2^51 */
        n = (int64_t)fn;
        /* fn = fh + fl, fh a multiple of 2^20: with |fn| < 2^40 both halves have <= 20 significant
         * bits, so every product with a 33-bit piece of pi/2 is exact */
        const double SPLIT = 7083549724304467820544.0;      /* 1.5 * 2^72 */
        const double fh = SC_SUB(SC_ADD(fn, SPLIT), SPLIT);
        const double fl = SC_SUB(fn, fh);
        double e1, e2, e3, e4;
        const double r0 = SC_SUB(SC_SUB(x, SC_MUL(fh, PIO2_1)), SC_MUL(fl, PIO2_1));   /* exact */
        const double r1 = shapes_sc_two_sum(r0, -SC_MUL(fh, PIO2_2), &e1);
        const double r2 = shapes_sc_two_sum(r1, -SC_MUL(fl, PIO2_2), &e2);
        const double r2b = shapes_sc_two_sum(r2, -SC_MUL(fh, PIO2_3), &e3);
        const double r3 = shapes_sc_two_sum(r2b, -SC_MUL(fl, PIO2_3), &e4);
        const double lo = SC_SUB(SC_ADD(SC_ADD(SC_ADD(e1, e2), e3), e4), SC_MUL(fn, PIO2_3T));
        y0 = SC_ADD(r3, lo);
        y1 = SC_ADD(SC_SUB(r3, y0), lo);
    }
    const double ks = shapes_sc_ksin(y0, y1), kc = shapes_sc_kcos(y0, y1);
    switch ((int)(n & 3)) {
    case 0: *sin_out = ks; *cos_out = kc; break;
    case 1: *sin_out = kc; *cos_out = -ks; break;
    case 2: *sin_out = -ks; *cos_out = -kc; break;
    default: *sin_out = -kc; *cos_out = ks; break;
    }
}

#endif /* SHAPES_SINCOS_H */

/*
 * nixis_oracle4.c -- CPU ORACLE, 4-D OpenSimplex.  TEST INFRASTRUCTURE ONLY (see nixis_oracle.c).
 *
 * IEEE-double restatement of opensimplex.py:133-141 (extrapolate4d) and :771-1958 (noise4d),
 * pinned bit-for-bit to tests/golden/noise.npz (v4_*: 7 500 points x 2 seeds computed by the
 * unmodified reference).  The reference never calls noise4d from its pipeline; the
 * fBm driver nxo_sample_octaves4 below is builder-defined by analogy with terrain.py:12-59.
 *
 * Restated around lattice OFFSETS: a contribution is base + o[4]; its displacement is
 * (d0 - o) - mult*SQUISH where mult is the multiplier the reference spells out (normally the
 * coordinate sum of the offset; opensimplex.py:1297-1321 uses 3 and 2 for the same point when
 * a|b covers all four axes, which is reproduced literally).  Where the reference applies an
 * integer step AFTER the squish term (`dy_ext1 += 1`, `dx_ext2 -= 2`, ...) ld[] holds that step
 * so the roundings happen in the same order.
 */
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NXO_API __attribute__((visibility("default")))

static const double ST4 = -0.138196601125011, SQ4 = 0.309016994374947;

typedef struct { int o[4]; int mult; int ld[4]; } ext4_t;

static inline int64_t ffloor4(double x)
{
    int64_t xi = (int64_t)x;
    return x < (double)xi ? xi - 1 : xi;
}

/* GRADIENTS_4D[idx..idx+3], idx = 4 * (16 sign patterns x 4 members): member a carries 3 on axis a,
 * sign bit c of the pattern negates axis c (opensimplex.py:64-83). */
static inline double grad4(const int32_t *perm, const int64_t b[4], const int o[4], const double d[4])
{
    int idx = perm[(perm[(perm[(perm[(b[0] + o[0]) & 0xFF] + (b[1] + o[1])) & 0xFF] + (b[2] + o[2])) & 0xFF]
                    + (b[3] + o[3])) & 0xFF] & 0xFC;
    int a = (idx >> 2) & 3, q = idx >> 4;
    double g[4];
    for (int c = 0; c < 4; ++c) g[c] = ((q >> c) & 1 ? -1.0 : 1.0) * (a == c ? 3.0 : 1.0);
    return g[0] * d[0] + g[1] * d[1] + g[2] * d[2] + g[3] * d[3];
}

static inline void add4(double *v, const int32_t *perm, const int64_t b[4], const double d0[4], const ext4_t *e)
{
    static const double SQM[5] = { 0.0, 0.309016994374947, 2 * 0.309016994374947, 3 * 0.309016994374947,
                                   4 * 0.309016994374947 };
    double d[4];
    for (int c = 0; c < 4; ++c) {
        double t = d0[c] - (double)(e->o[c] - e->ld[c]);
        t = t - SQM[e->mult];
        if (e->ld[c]) t = t - (double)e->ld[c];
        d[c] = t;
    }
    double at = 2 - d[0] * d[0] - d[1] * d[1] - d[2] * d[2] - d[3] * d[3];
    if (at > 0) { at *= at; *v += at * at * grad4(perm, b, e->o, d); }
}

static inline void corner4(double *v, const int32_t *perm, const int64_t b[4], const double d0[4], int code)
{
    ext4_t e = { { code & 1, (code >> 1) & 1, (code >> 2) & 1, (code >> 3) & 1 }, 0, { 0, 0, 0, 0 } };
    e.mult = e.o[0] + e.o[1] + e.o[2] + e.o[3];
    add4(v, perm, b, d0, &e);
}

static inline int first_set(int c) { return (c & 1) ? 0 : ((c & 2) ? 1 : ((c & 4) ? 2 : 3)); }
static inline int first_clear(int c) { return !(c & 1) ? 0 : (!(c & 2) ? 1 : (!(c & 4) ? 2 : 3)); }

/* opensimplex.py:1337-1381 / 1383-1431: two extras derived from a point code c on the SMALL side of
 * the first dispentachoron (bits -> +1, one missing axis each -> -1), multiplier 1. */
static void ext_pair_minus(int c, ext4_t *e0, ext4_t *e1)
{
    for (int a = 0; a < 4; ++a) { e0->o[a] = e1->o[a] = (c >> a) & 1; e0->ld[a] = e1->ld[a] = 0; }
    e0->mult = e1->mult = 1;
    if (!(c & 1)) e0->o[0] = -1;                                   /* dx0 + 1 - SQ: generic order */
    if (!(c & 2)) { if ((c & 1) == 1) { e0->o[1] = -1; e0->ld[1] = -1; } else { e1->o[1] = -1; e1->ld[1] = -1; } }
    if (!(c & 4)) { if ((c & 3) == 3) { e0->o[2] = -1; e0->ld[2] = -1; } else { e1->o[2] = -1; e1->ld[2] = -1; } }
    if (!(c & 8)) e1->o[3] = -1;                                   /* dw0 + 1 - SQ: generic order */
}

/* opensimplex.py:1719-1763 / 1765-1815: two extras from a code c (bits -> 1, one of them -> 2), mult 3 */
static void ext_pair_plus(int c, ext4_t *e0, ext4_t *e1)
{
    for (int a = 0; a < 4; ++a) { e0->o[a] = e1->o[a] = (c >> a) & 1; e0->ld[a] = e1->ld[a] = 0; }
    e0->mult = e1->mult = 3;
    if (c & 1) e0->o[0] = 2;
    if (c & 2) { if ((c & 1) == 0) { e0->o[1] = 2; e0->ld[1] = 1; } else { e1->o[1] = 2; e1->ld[1] = 1; } }
    if (c & 4) { if ((c & 3) == 0) { e0->o[2] = 2; e0->ld[2] = 1; } else { e1->o[2] = 2; e1->ld[2] = 1; } }
    if (c & 8) e1->o[3] = 2;
}

NXO_API double nxo_noise4(double x, double y, double z, double w, const int32_t *perm)
{
    double so = (x + y + z + w) * ST4;
    double s[4] = { x + so, y + so, z + so, w + so };
    int64_t b[4] = { ffloor4(s[0]), ffloor4(s[1]), ffloor4(s[2]), ffloor4(s[3]) };
    double qo = (double)(b[0] + b[1] + b[2] + b[3]) * SQ4;
    double f[4], d0[4];
    const double in[4] = { x, y, z, w };
    for (int c = 0; c < 4; ++c) { f[c] = s[c] - (double)b[c]; d0[c] = in[c] - ((double)b[c] + qo); }
    double fsum = f[0] + f[1] + f[2] + f[3];
    double v = 0;
    ext4_t e0 = { {0,0,0,0}, 0, {0,0,0,0} }, e1 = e0, e2 = e0;

    if (fsum <= 1) {                                   /* pentachoron at (0,0,0,0): :811-975 */
        int ap = 1, bp = 2; double as = f[0], bs = f[1];
        if (as >= bs && f[2] > bs) { bs = f[2]; bp = 4; } else if (as < bs && f[2] > as) { as = f[2]; ap = 4; }
        if (as >= bs && f[3] > bs) { bs = f[3]; bp = 8; } else if (as < bs && f[3] > as) { as = f[3]; ap = 8; }
        double u = 1 - fsum;
        if (u > as || u > bs) {
            int c = (bs > as) ? bp : ap;               /* single axis */
            for (int a = 0; a < 4; ++a) e0.o[a] = e1.o[a] = e2.o[a] = (c >> a) & 1;
            e0.mult = e1.mult = e2.mult = 0;
            if (!(c & 1)) e0.o[0] = -1;
            if (!(c & 2)) { if ((c & 1) == 1) e0.o[1] = -1; else e1.o[1] = -1; }
            if (!(c & 4)) { if ((c & 3) != 0) { if ((c & 3) == 3) e0.o[2] = -1; else e1.o[2] = -1; } else e2.o[2] = -1; }
            if (!(c & 8)) e2.o[3] = -1;
        } else {
            int c = ap | bp;                           /* two axes */
            for (int a = 0; a < 4; ++a) e0.o[a] = e1.o[a] = e2.o[a] = (c >> a) & 1;
            e0.mult = 2; e1.mult = e2.mult = 1;
            if (!(c & 1)) e1.o[0] = -1;                /* dx0 + 1 - SQ */
            if (!(c & 2)) { if ((c & 1) == 1) { e1.o[1] = -1; e1.ld[1] = -1; } else { e2.o[1] = -1; e2.ld[1] = -1; } }
            if (!(c & 4)) { if ((c & 3) == 3) { e1.o[2] = -1; e1.ld[2] = -1; } else { e2.o[2] = -1; e2.ld[2] = -1; } }
            if (!(c & 8)) e2.o[3] = -1;                /* dw0 + 1 - SQ */
        }
        corner4(&v, perm, b, d0, 0); corner4(&v, perm, b, d0, 1); corner4(&v, perm, b, d0, 2);
        corner4(&v, perm, b, d0, 4); corner4(&v, perm, b, d0, 8);
    } else if (fsum >= 3) {                            /* pentachoron at (1,1,1,1): :976-1166 */
        int ap = 0xE, bp = 0xD; double as = f[0], bs = f[1];
        if (as <= bs && f[2] < bs) { bs = f[2]; bp = 0xB; } else if (as > bs && f[2] < as) { as = f[2]; ap = 0xB; }
        if (as <= bs && f[3] < bs) { bs = f[3]; bp = 0x7; } else if (as > bs && f[3] < as) { as = f[3]; ap = 0x7; }
        double u = 4 - fsum;
        if (u < as || u < bs) {
            int c = (bs < as) ? bp : ap;               /* three axes */
            for (int a = 0; a < 4; ++a) e0.o[a] = e1.o[a] = e2.o[a] = (c >> a) & 1;
            e0.mult = e1.mult = e2.mult = 4;
            if (c & 1) e0.o[0] = 2;
            if (c & 2) { if (c & 1) { e1.o[1] = 2; e1.ld[1] = 1; } else { e0.o[1] = 2; e0.ld[1] = 1; } }
            if (c & 4) {
                if ((c & 3) != 3) { if ((c & 3) == 0) { e0.o[2] = 2; e0.ld[2] = 1; } else { e1.o[2] = 2; e1.ld[2] = 1; } }
                else { e2.o[2] = 2; e2.ld[2] = 1; }
            }
            if (c & 8) e2.o[3] = 2;
        } else {
            int c = ap & bp;                           /* two axes */
            for (int a = 0; a < 4; ++a) e0.o[a] = e1.o[a] = e2.o[a] = (c >> a) & 1;
            e0.mult = 2; e1.mult = e2.mult = 3;
            if (c & 1) e1.o[0] = 2;
            if (c & 2) { if (c & 1) { e2.o[1] = 2; e2.ld[1] = 1; } else { e1.o[1] = 2; e1.ld[1] = 1; } }
            if (c & 4) { if ((c & 3) != 0) { e2.o[2] = 2; e2.ld[2] = 1; } else { e1.o[2] = 2; e1.ld[2] = 1; } }
            if (c & 8) e2.o[3] = 2;
        }
        corner4(&v, perm, b, d0, 7); corner4(&v, perm, b, d0, 0xB); corner4(&v, perm, b, d0, 0xD);
        corner4(&v, perm, b, d0, 0xE); corner4(&v, perm, b, d0, 0xF);
    } else if (fsum <= 2) {                            /* first dispentachoron: :1167-1559 */
        int abig = 1, bbig = 1, ap, bp; double as, bs, sc;
        if (f[0] + f[1] > f[2] + f[3]) { as = f[0] + f[1]; ap = 0x3; } else { as = f[2] + f[3]; ap = 0xC; }
        if (f[0] + f[2] > f[1] + f[3]) { bs = f[0] + f[2]; bp = 0x5; } else { bs = f[1] + f[3]; bp = 0xA; }
        if (f[0] + f[3] > f[1] + f[2]) {
            sc = f[0] + f[3];
            if (as >= bs && sc > bs) { bs = sc; bp = 0x9; } else if (as < bs && sc > as) { as = sc; ap = 0x9; }
        } else {
            sc = f[1] + f[2];
            if (as >= bs && sc > bs) { bs = sc; bp = 0x6; } else if (as < bs && sc > as) { as = sc; ap = 0x6; }
        }
        for (int a = 0; a < 4; ++a) {                  /* (1,0,0,0) .. (0,0,0,1); the last does not update scores */
            double p = 2 - fsum + f[a];
            if (as >= bs && p > bs) { if (a < 3) bs = p; bp = 1 << a; bbig = 0; }
            else if (as < bs && p > as) { if (a < 3) as = p; ap = 1 << a; abig = 0; }
        }
        if (abig == bbig) {
            if (abig) {
                int c1 = ap | bp, c2 = ap & bp;
                for (int a = 0; a < 4; ++a) { int on = (c1 >> a) & 1; e0.o[a] = on; e1.o[a] = on ? 1 : -1; }
                e0.mult = 3; e1.mult = 2;
                e2.mult = 2;
                { int ax = first_set(c2 ? c2 : 8); e2.o[ax] = 2; e2.ld[ax] = 2; }
            } else {
                e2.mult = 0;                           /* (0,0,0,0) */
                ext_pair_minus(ap | bp, &e0, &e1);
            }
        } else {
            int c1 = abig ? ap : bp, c2 = abig ? bp : ap;
            ext_pair_minus(c1, &e0, &e1);
            e2.mult = 2;
            { int ax = first_set(c2); e2.o[ax] = 2; e2.ld[ax] = 2; }
        }
        static const int order[10] = { 1, 2, 4, 8, 3, 5, 9, 6, 0xA, 0xC };
        for (int i = 0; i < 10; ++i) corner4(&v, perm, b, d0, order[i]);
    } else {                                           /* second dispentachoron: :1560-1935 */
        int abig = 1, bbig = 1, ap, bp; double as, bs, sc;
        if (f[0] + f[1] < f[2] + f[3]) { as = f[0] + f[1]; ap = 0xC; } else { as = f[2] + f[3]; ap = 0x3; }
        if (f[0] + f[2] < f[1] + f[3]) { bs = f[0] + f[2]; bp = 0xA; } else { bs = f[1] + f[3]; bp = 0x5; }
        if (f[0] + f[3] < f[1] + f[2]) {
            sc = f[0] + f[3];
            if (as <= bs && sc < bs) { bs = sc; bp = 0x6; } else if (as > bs && sc < as) { as = sc; ap = 0x6; }
        } else {
            sc = f[1] + f[2];
            if (as <= bs && sc < bs) { bs = sc; bp = 0x9; } else if (as > bs && sc < as) { as = sc; ap = 0x9; }
        }
        for (int a = 0; a < 4; ++a) {                  /* (0,1,1,1) .. (1,1,1,0) */
            double p = 3 - fsum + f[a];
            int code = 0xF & ~(1 << a);
            if (as <= bs && p < bs) { if (a < 3) bs = p; bp = code; bbig = 0; }
            else if (as > bs && p < as) { if (a < 3) as = p; ap = code; abig = 0; }
        }
        if (abig == bbig) {
            if (abig) {
                int c1 = ap & bp, c2 = ap | bp;
                int ax = first_set(c1 ? c1 : 8);
                e0.mult = 1; e0.o[ax] = 1; e0.ld[ax] = 1;
                e1.mult = 2; e1.o[ax] = 2; e1.ld[ax] = 2;
                for (int a = 0; a < 4; ++a) e2.o[a] = 1;
                e2.mult = 2;
                { int az = first_clear(c2); e2.o[az] = -1; e2.ld[az] = -2; }
            } else {
                for (int a = 0; a < 4; ++a) e2.o[a] = 1;
                e2.mult = 4;
                ext_pair_plus(ap & bp, &e0, &e1);
            }
        } else {
            int c1 = abig ? ap : bp, c2 = abig ? bp : ap;
            ext_pair_plus(c1, &e0, &e1);
            for (int a = 0; a < 4; ++a) e2.o[a] = 1;
            e2.mult = 2;
            { int az = first_clear(c2); e2.o[az] = -1; e2.ld[az] = -2; }
        }
        static const int order[10] = { 7, 0xB, 0xD, 0xE, 3, 5, 9, 6, 0xA, 0xC };
        for (int i = 0; i < 10; ++i) corner4(&v, perm, b, d0, order[i]);
    }
    add4(&v, perm, b, d0, &e0);
    add4(&v, perm, b, d0, &e1);
    add4(&v, perm, b, d0, &e2);
    return v / 30;
}

/* opensimplex.py:762-768 */
NXO_API void nxo_noise4_array(int64_t n, const double *x, const double *y, const double *z, const double *w,
                              const int32_t *perm, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = nxo_noise4(x[i], y[i], z[i], w[i], perm);
}

/* 4-D fBm (builder-defined, by analogy with terrain.py:12-59): octave o samples
 * noise4d(v*nr, w_scale*freq_o) and adds ((e + 1) * 0.5 * ns) * R. */
NXO_API void nxo_sample_octaves4(int64_t n, const double *verts, double *elev, const int32_t *perm, int n_octaves,
                                 double f0, double a0, double roughness, double persistence,
                                 double world_radius, double w_scale, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    double fr = f0, am = a0;
    for (int o = 0; o < n_octaves; ++o) {
        double nr = fr / world_radius, ns = am / world_radius, ww = w_scale * fr;
#pragma omp parallel for schedule(static)
        for (int64_t v = 0; v < n; ++v) {
            double e = nxo_noise4(verts[3 * v] * nr, verts[3 * v + 1] * nr, verts[3 * v + 2] * nr, ww, perm);
            elev[v] += (e + 1) * 0.5 * ns * world_radius;
        }
        fr *= roughness; am *= persistence;
    }
}

/*
 * nixis_oracle.c -- CPU ORACLE for the nixis terrain hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a from-scratch IEEE-double restatement of the algorithms that
 * MightyBOBcnc/nixis runs on the CPU (numba) for the path BASELINE.json names.
 * It exists so that the CUDA product path (nixis_b200/csrc) can be checked
 * against the reference's arithmetic on a box where /root/reference is absent.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load it.  The product package never imports it.
 *
 * Parity status: PINNED.  tests/golden/*.npz were produced by running the
 * unmodified reference (numba) in the build container (tests/golden/gen_golden.py)
 * and this restatement reproduces them bit-for-bit (tests/test_oracle_golden.py).
 * Exception: the icosphere generator lives in oracle/icosphere.py and is
 * "parity unpinned" (meshzoo is absent from the reference tree).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC   (oracle/Makefile)
 * -ffp-contract=off matters: numba emits no FMA contraction, so every a*b+c
 * below must round twice exactly like the reference.
 *
 * Each function cites the reference file:line it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NXO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* constants: opensimplex.py:24-33                                            */
static const double ST2 = -0.211324865405187, SQ2 = 0.366025403784439;
static const double ST3 = -1.0 / 6, SQ3 = 1.0 / 3;
static const double ST4 = -0.138196601125011, SQ4 = 0.309016994374947;

/* gradient sets: opensimplex.py:40-83.  Generated instead of listed:
 * 3-D: 8 sign octants x {11 on x, 11 on y, 11 on z}, octant bits = (x-,y-,z-)
 * with the x bit inverted (index 0 is (-11,4,4)).
 * 4-D: 16 sign patterns x {3 on x,y,z,w}. 2-D: octagon (5,2)/(2,5). */
static double G2[16], G3[72], G4[256];
static int tables_ready = 0;

static void build_gradients(void)
{
    if (tables_ready) return;
    for (int q = 0; q < 8; ++q) {
        double sx = (q & 1) ? 1.0 : -1.0, sy = (q & 2) ? -1.0 : 1.0, sz = (q & 4) ? -1.0 : 1.0;
        for (int a = 0; a < 3; ++a) {
            double *g = &G3[(q * 3 + a) * 3];
            g[0] = sx * (a == 0 ? 11 : 4);
            g[1] = sy * (a == 1 ? 11 : 4);
            g[2] = sz * (a == 2 ? 11 : 4);
        }
    }
    for (int q = 0; q < 16; ++q) {
        double s[4] = { (q & 1) ? -1.0 : 1.0, (q & 2) ? -1.0 : 1.0, (q & 4) ? -1.0 : 1.0, (q & 8) ? -1.0 : 1.0 };
        for (int a = 0; a < 4; ++a)
            for (int c = 0; c < 4; ++c)
                G4[(q * 4 + a) * 4 + c] = s[c] * (a == c ? 3 : 1);
    }
    for (int q = 0; q < 4; ++q) {
        double sx = (q & 1) ? -1.0 : 1.0, sy = (q & 2) ? -1.0 : 1.0;
        G2[q * 4 + 0] = sx * 5; G2[q * 4 + 1] = sy * 2;
        G2[q * 4 + 2] = sx * 2; G2[q * 4 + 3] = sy * 5;
    }
    tables_ready = 1;
}

/* opensimplex.py:18-21 */
static inline int64_t ffloor(double x)
{
    int64_t xi = (int64_t)x;
    return x < (double)xi ? xi - 1 : xi;
}

/* ------------------------------------------------------------------------- */
/* opensimplex.py:90-112.  `over` is declared int32(int32): every LCG state is
 * the wrapped 64-bit product truncated to a signed 32-bit value.             */
static inline int64_t lcg32(int64_t s)
{
    uint64_t u = (uint64_t)s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (int64_t)(int32_t)(uint32_t)u;
}

NXO_API void nxo_init(int64_t seed, int32_t *perm, int32_t *pgi)
{
    int64_t source[256];
    for (int i = 0; i < 256; ++i) source[i] = i;
    seed = lcg32(seed); seed = lcg32(seed); seed = lcg32(seed);
    for (int i = 255; i >= 0; --i) {
        seed = lcg32(seed);
        int64_t r = (seed + 31) % (i + 1);
        if (r < 0) r += i + 1;             /* Python floor-mod */
        perm[i] = (int32_t)source[r];
        pgi[i] = (int32_t)(fmod((double)perm[i], 24.0) * 3);
        source[r] = source[i];
    }
}

/* ------------------------------------------------------------------------- */
/* opensimplex.py:115-120, 153-254                                            */
static inline double grad2(const int32_t *perm, int64_t xb, int64_t yb, double dx, double dy)
{
    int idx = perm[(perm[xb & 0xFF] + yb) & 0xFF] & 0x0E;
    return G2[idx] * dx + G2[idx + 1] * dy;
}

static inline void add2(double *v, const int32_t *perm, int64_t xb, int64_t yb, double dx, double dy)
{
    double at = 2 - dx * dx - dy * dy;
    if (at > 0) { at *= at; *v += at * at * grad2(perm, xb, yb, dx, dy); }
}

NXO_API double nxo_noise2(double x, double y, const int32_t *perm)
{
    build_gradients();
    double so = (x + y) * ST2;
    double xs = x + so, ys = y + so;
    int64_t xb = ffloor(xs), yb = ffloor(ys);
    double qo = (double)(xb + yb) * SQ2;
    double ox = (double)xb + qo, oy = (double)yb + qo;
    double fx = xs - (double)xb, fy = ys - (double)yb;
    double fsum = fx + fy;
    double dx0 = x - ox, dy0 = y - oy;
    double v = 0;
    add2(&v, perm, xb + 1, yb, dx0 - 1 - SQ2, dy0 - 0 - SQ2);
    add2(&v, perm, xb, yb + 1, dx0 - 0 - SQ2, dy0 - 1 - SQ2);
    int64_t ex, ey; double edx, edy;
    if (fsum <= 1) {
        double fz = 1 - fsum;
        if (fz > fx || fz > fy) {
            if (fx > fy) { ex = xb + 1; ey = yb - 1; edx = dx0 - 1; edy = dy0 + 1; }
            else         { ex = xb - 1; ey = yb + 1; edx = dx0 + 1; edy = dy0 - 1; }
        } else { ex = xb + 1; ey = yb + 1; edx = dx0 - 1 - 2 * SQ2; edy = dy0 - 1 - 2 * SQ2; }
    } else {
        double fz = 2 - fsum;
        if (fz < fx || fz < fy) {
            if (fx > fy) { ex = xb + 2; ey = yb;     edx = dx0 - 2 - 2 * SQ2; edy = dy0 + 0 - 2 * SQ2; }
            else         { ex = xb;     ey = yb + 2; edx = dx0 + 0 - 2 * SQ2; edy = dy0 - 2 - 2 * SQ2; }
        } else { ex = xb; ey = yb; edx = dx0; edy = dy0; }
        xb += 1; yb += 1;
        dx0 = dx0 - 1 - 2 * SQ2; dy0 = dy0 - 1 - 2 * SQ2;
    }
    add2(&v, perm, xb, yb, dx0, dy0);
    add2(&v, perm, ex, ey, edx, edy);
    return v / 47;
}

/* ------------------------------------------------------------------------- */
/* 3-D.  opensimplex.py:123-130 (gradient), 266-759 (noise3d).
 *
 * Restated around lattice OFFSETS: every contributing lattice point is
 * base + (i,j,k); its displacement is (d0 - i) - m*SQ3 with m = i+j+k (the
 * reference writes the same thing as `dx0 - 1 - 2*SQUISH`, i.e. two roundings).
 * Two reference spellings subtract the integer AFTER the squish term
 * (`dy_ext1 -= 1`, opensimplex.py:438-443; `dx_ext1 -= 2`, :683-691); the
 * `late` mask reproduces that rounding order.                                 */
typedef struct { int o[3]; int late; /* bit a set: axis a integer part applied last */ int lateamt; } ext3_t;

static inline double grad3(const int32_t *perm, const int32_t *pgi,
                           int64_t xb, int64_t yb, int64_t zb, double dx, double dy, double dz)
{
    int idx = pgi[(perm[(perm[xb & 0xFF] + yb) & 0xFF] + zb) & 0xFF];
    return G3[idx] * dx + G3[idx + 1] * dy + G3[idx + 2] * dz;
}

static inline void add3(double *v, const int32_t *perm, const int32_t *pgi,
                        int64_t xb, int64_t yb, int64_t zb, double dx, double dy, double dz)
{
    double at = 2 - dx * dx - dy * dy - dz * dz;
    if (at > 0) { at *= at; *v += at * at * grad3(perm, pgi, xb, yb, zb, dx, dy, dz); }
}

static const double SQ3M[4] = { 0.0, 1.0 / 3, 2 * (1.0 / 3), 3 * (1.0 / 3) };

/* lattice point at fixed offset (i,j,k) in {0,1}^3 */
static inline void corner3(double *v, const int32_t *perm, const int32_t *pgi,
                           const int64_t b[3], const double d0[3], int i, int j, int k)
{
    double c = SQ3M[i + j + k];
    add3(v, perm, pgi, b[0] + i, b[1] + j, b[2] + k, d0[0] - i - c, d0[1] - j - c, d0[2] - k - c);
}

static inline void extra3(double *v, const int32_t *perm, const int32_t *pgi,
                          const int64_t b[3], const double d0[3], const ext3_t *e)
{
    int m = e->o[0] + e->o[1] + e->o[2];
    double c = SQ3M[m], d[3];
    for (int a = 0; a < 3; ++a) {
        if (e->late & (1 << a)) d[a] = (d0[a] - (e->o[a] - e->lateamt)) - c - e->lateamt;
        else                    d[a] = d0[a] - e->o[a] - c;
    }
    add3(v, perm, pgi, b[0] + e->o[0], b[1] + e->o[1], b[2] + e->o[2], d[0], d[1], d[2]);
}

NXO_API double nxo_noise3(double x, double y, double z, const int32_t *perm, const int32_t *pgi)
{
    build_gradients();
    double so = (x + y + z) * ST3;
    double s[3] = { x + so, y + so, z + so };
    int64_t b[3] = { ffloor(s[0]), ffloor(s[1]), ffloor(s[2]) };
    double qo = (double)(b[0] + b[1] + b[2]) * SQ3;
    double f[3] = { s[0] - (double)b[0], s[1] - (double)b[1], s[2] - (double)b[2] };
    double fsum = f[0] + f[1] + f[2];
    double d0[3] = { x - ((double)b[0] + qo), y - ((double)b[1] + qo), z - ((double)b[2] + qo) };
    double v = 0;
    ext3_t e0 = { {0, 0, 0}, 0, 0 }, e1 = { {0, 0, 0}, 0, 0 };

    if (fsum <= 1) {                       /* tetrahedron at (0,0,0): :299-416 */
        int ap = 1, bp = 2; double as = f[0], bs = f[1];
        if (as >= bs && f[2] > bs) { bs = f[2]; bp = 4; }
        else if (as < bs && f[2] > as) { as = f[2]; ap = 4; }
        double w = 1 - fsum;
        if (w > as || w > bs) {
            int c = (bs > as) ? bp : ap;   /* single axis bit */
            for (int a = 0; a < 3; ++a) { e0.o[a] = e1.o[a] = (c >> a) & 1; }
            /* the two axes not in c: one gets -1 on ext0, the other -1 on ext1 */
            if (!(c & 1)) { e0.o[0] = -1; }
            if (!(c & 2)) { if (!(c & 1)) e1.o[1] = -1; else e0.o[1] = -1; }
            if (!(c & 4)) { e1.o[2] = -1; }
        } else {
            int c = ap | bp;               /* two axis bits */
            for (int a = 0; a < 3; ++a) {
                int on = (c >> a) & 1;
                e0.o[a] = on; e1.o[a] = on ? 1 : -1;
            }
        }
        corner3(&v, perm, pgi, b, d0, 0, 0, 0);
        corner3(&v, perm, pgi, b, d0, 1, 0, 0);
        corner3(&v, perm, pgi, b, d0, 0, 1, 0);
        corner3(&v, perm, pgi, b, d0, 0, 0, 1);
    } else if (fsum >= 2) {                /* tetrahedron at (1,1,1): :417-534 */
        int ap = 6, bp = 5; double as = f[0], bs = f[1];
        if (as <= bs && f[2] < bs) { bs = f[2]; bp = 3; }
        else if (as > bs && f[2] < as) { as = f[2]; ap = 3; }
        double w = 3 - fsum;
        if (w < as || w < bs) {
            int c = (bs < as) ? bp : ap;   /* two axis bits */
            for (int a = 0; a < 3; ++a) { e0.o[a] = e1.o[a] = (c >> a) & 1; }
            if (c & 1) { e0.o[0] = 2; }
            if (c & 2) {                   /* `+= 1` after the squish term */
                if (c & 1) { e1.o[1] = 2; e1.late = 2; e1.lateamt = 1; }
                else       { e0.o[1] = 2; e0.late = 2; e0.lateamt = 1; }
            }
            if (c & 4) { e1.o[2] = 2; }
        } else {
            int c = ap & bp;               /* single axis bit */
            for (int a = 0; a < 3; ++a) {
                int on = (c >> a) & 1;
                e0.o[a] = on; e1.o[a] = 2 * on;
            }
        }
        corner3(&v, perm, pgi, b, d0, 1, 1, 0);
        corner3(&v, perm, pgi, b, d0, 1, 0, 1);
        corner3(&v, perm, pgi, b, d0, 0, 1, 1);
        corner3(&v, perm, pgi, b, d0, 1, 1, 1);
    } else {                               /* octahedron: :535-745 */
        double as, bs, sc; int ap, bp, afar, bfar;
        double p1 = f[0] + f[1];
        if (p1 > 1) { as = p1 - 1; ap = 3; afar = 1; } else { as = 1 - p1; ap = 4; afar = 0; }
        double p2 = f[0] + f[2];
        if (p2 > 1) { bs = p2 - 1; bp = 5; bfar = 1; } else { bs = 1 - p2; bp = 2; bfar = 0; }
        double p3 = f[1] + f[2];
        if (p3 > 1) {
            sc = p3 - 1;
            if (as <= bs && as < sc) { ap = 6; afar = 1; }
            else if (as > bs && bs < sc) { bp = 6; bfar = 1; }
        } else {
            sc = 1 - p3;
            if (as <= bs && as < sc) { ap = 1; afar = 0; }
            else if (as > bs && bs < sc) { bp = 1; bfar = 0; }
        }
        if (afar == bfar) {
            if (afar) {                    /* both near (1,1,1): ext0=(1,1,1), ext1 = 2 on shared axis */
                e0.o[0] = e0.o[1] = e0.o[2] = 1;
                int c = ap & bp;
                int ax = (c & 1) ? 0 : ((c & 2) ? 1 : 2);
                e1.o[ax] = 2;
            } else {                       /* both near (0,0,0): ext0=(0,0,0), ext1 = -1 on omitted axis */
                int c = ap | bp;
                int ax = !(c & 1) ? 0 : (!(c & 2) ? 1 : 2);
                e1.o[0] = e1.o[1] = e1.o[2] = 1; e1.o[ax] = -1;
            }
        } else {
            int c1 = afar ? ap : bp, c2 = afar ? bp : ap;
            int ax = !(c1 & 1) ? 0 : (!(c1 & 2) ? 1 : 2);
            e0.o[0] = e0.o[1] = e0.o[2] = 1; e0.o[ax] = -1;
            int ay = (c2 & 1) ? 0 : ((c2 & 2) ? 1 : 2);
            e1.o[ay] = 2; e1.late = 1 << ay; e1.lateamt = 2;   /* `-= 2` after the squish term */
        }
        corner3(&v, perm, pgi, b, d0, 1, 0, 0);
        corner3(&v, perm, pgi, b, d0, 0, 1, 0);
        corner3(&v, perm, pgi, b, d0, 0, 0, 1);
        corner3(&v, perm, pgi, b, d0, 1, 1, 0);
        corner3(&v, perm, pgi, b, d0, 1, 0, 1);
        corner3(&v, perm, pgi, b, d0, 0, 1, 1);
    }
    extra3(&v, perm, pgi, b, d0, &e0);
    extra3(&v, perm, pgi, b, d0, &e1);
    return v / 103;
}

/* opensimplex.py:257-263 */
NXO_API void nxo_noise3_array(int64_t n, const double *x, const double *y, const double *z,
                              const int32_t *perm, const int32_t *pgi, double *out)
{
    build_gradients();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = nxo_noise3(x[i], y[i], z[i], perm, pgi);
}

NXO_API void nxo_noise2_array(int64_t n, const double *x, const double *y, const int32_t *perm, double *out)
{
    build_gradients();
    for (int64_t i = 0; i < n; ++i) out[i] = nxo_noise2(x[i], y[i], perm);
}

/* ------------------------------------------------------------------------- */
/* terrain.py:12-29 (sample_noise) + terrain.py:32-47 (octave loop).
 * verts are the caller's (radius-scaled) positions, f64[n][3].               */
NXO_API void nxo_sample_octaves(int64_t n, const double *verts, double *elev,
                                const int32_t *perm, const int32_t *pgi, int n_octaves,
                                double f0, double a0, double roughness, double persistence,
                                double world_radius, int nthreads)
{
    build_gradients();
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    double fr = f0, am = a0;
    for (int o = 0; o < n_octaves; ++o) {
        double nr = fr / world_radius, ns = am / world_radius;
#pragma omp parallel for schedule(static)
        for (int64_t v = 0; v < n; ++v) {
            double e = nxo_noise3(verts[3 * v] * nr, verts[3 * v + 1] * nr, verts[3 * v + 2] * nr, perm, pgi);
            elev[v] += (e + 1) * 0.5 * ns * world_radius;
        }
        fr *= roughness; am *= persistence;
    }
}

/* terrain.py:61-72 */
NXO_API void nxo_mask_le(int64_t n, const double *h, double level, uint8_t *mask)
{
    for (int64_t i = 0; i < n; ++i) mask[i] = h[i] <= level;
}

/* ------------------------------------------------------------------------- */
/* util.py:110-175.  mode: 0 None, 1 'lower', 2 'upper'.  has_* flags stand in
 * for Python None.  Returns 0, or 1 for the "mode without mid" error path
 * (reference prints and returns x itself; out is then a copy of x).          */
NXO_API int nxo_rescale(int64_t n, const double *x, double *out, double lower, double upper,
                        int has_mid, double mid, int mode,
                        int has_umin, double umin, int has_umax, double umax)
{
    double lo = x[0], hi = x[0];
    for (int64_t i = 1; i < n; ++i) { if (x[i] < lo) lo = x[i]; if (x[i] > hi) hi = x[i]; }
    if (has_umin && umin < lo) lo = umin;
    if (has_umax && umax > hi) hi = umax;
    if (out != x) memcpy(out, x, (size_t)n * sizeof(double));
    if (mode == 0) {
        if (!has_mid) {
            double xr = hi - lo, nr = upper - lower;
            for (int64_t i = 0; i < n; ++i) out[i] = ((x[i] - lo) / xr) * nr + lower;
        } else {
            double xlr = mid - lo, nlr = mid - lower, xur = hi - mid, nur = upper - mid;
            for (int64_t i = 0; i < n; ++i)
                out[i] = (x[i] <= mid) ? ((x[i] - lo) / xlr) * nlr + lower
                                       : ((x[i] - mid) / xur) * nur + mid;
        }
        return 0;
    }
    if (!has_mid) return 1;
    if (mode == 1) {
        double xr = mid - lo, nr = mid - lower;
        for (int64_t i = 0; i < n; ++i) if (x[i] <= mid) out[i] = ((x[i] - lo) / xr) * nr + lower;
    } else {
        double xr = hi - mid, nr = upper - mid;
        for (int64_t i = 0; i < n; ++i) if (x[i] >= mid) out[i] = ((x[i] - mid) / xr) * nr + mid;
    }
    return 0;
}

/* util.py:178-254.  mode: -1 None, 0 "not mask", 1 "mask".  The sequential
 * if/elif scan (:203-214) is kept literally: an element that lowers the
 * running minimum can never raise the running maximum.
 * stats[4] = x_min, x_max, mask_lower, mask_upper (the four printed values). */
NXO_API void nxo_power_rescale(int64_t n, const double *x, const uint8_t *mask, int mode,
                               double power, double *out, double *stats)
{
    double xmin = x[0], xmax = x[0];
    for (int64_t i = 1; i < n; ++i) { if (x[i] < xmin) xmin = x[i]; if (x[i] > xmax) xmax = x[i]; }
    double mlo = xmax, mhi = xmin;
    if (mode == 0 || mode == 1) {
        for (int64_t i = 0; i < n; ++i) {
            int sel = mode == 1 ? (mask[i] != 0) : (mask[i] == 0);
            if (sel && x[i] < mlo) mlo = x[i];
            else if (sel && x[i] > mhi) mhi = x[i];
        }
    }
    if (stats) { stats[0] = xmin; stats[1] = xmax; stats[2] = mlo; stats[3] = mhi; }
    double mr = mhi - mlo, tr = 1.0 - 0.0;
    for (int64_t i = 0; i < n; ++i) {
        int sel = mode == 1 ? (mask[i] != 0) : (mode == 0 ? (mask[i] == 0) : 0);
        if (!sel) { out[i] = x[i]; continue; }
        double t = ((x[i] - mlo) / mr) * tr + 0.0;
        t = pow(t, power);
        out[i] = ((t - 0.0) / tr) * mr + mlo;
    }
}

/* util.py:556-566 */
NXO_API double nxo_find_percent_val(double minval, double maxval, double percent)
{
    if (!(0.0 < percent && percent < 100.0)) percent = 50.0;
    return minval + ((maxval - minval) * percent / 100.0);
}

/* ------------------------------------------------------------------------- */
/* util.py:580-613: serial append of directed edges at the first free slot.
 * cells are int64[T][3] (meshzoo dtype=int on Linux).  Returns -1 if a row
 * overflows (the reference would silently write slot 5; we report instead).  */
NXO_API int nxo_build_adjacency(int64_t T, const int64_t *cells, int32_t *adj)
{
    int64_t V = (T + 4) / 2;
    for (int64_t i = 0; i < V * 6; ++i) adj[i] = -1;
    int rc = 0;
    for (int64_t t = 0; t < T; ++t) {
        for (int c = 0; c < 3; ++c) {
            int64_t v = cells[3 * t + c], nx = cells[3 * t + (c + 1) % 3];
            int32_t *row = adj + 6 * v;
            int slot = -1;
            for (int q = 0; q < 6; ++q) if (row[q] == -1) { slot = q; break; }
            if (slot < 0) { rc = -1; slot = 5; }
            row[slot] = (int32_t)nx;
        }
    }
    return rc;
}

/* util.py:623-662.  Race-free version of the in-place ring walk: rows are
 * read from the unsorted input and written to a separate output (membership
 * tests are order independent).                                   */
static inline int32_t next_in_ring(int32_t v, const int32_t *row_idx, const int32_t *row_nv)
{
    for (int a = 0; a < 6; ++a) {
        int32_t cand = row_idx[a];
        for (int q = 0; q < 6; ++q)
            if (row_nv[q] == cand) { if (cand != v) return cand; break; }
    }
    return -1;
}

NXO_API void nxo_sort_adjacency(int64_t V, const int32_t *adj_in, int32_t *adj_out)
{
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < V; ++idx) {
        int n = idx < 12 ? 5 : 6;
        int32_t ring[6] = { -1, -1, -1, -1, -1, -1 };
        const int32_t *row = adj_in + 6 * idx;
        int32_t pv = (int32_t)idx, nv = row[0];
        for (int s = 0; s < n - 1; ++s) {
            ring[s] = nv;
            /* python: adj[nv] with nv == -1 wraps to the last row */
            const int32_t *nrow = adj_in + 6 * (nv >= 0 ? (int64_t)nv : V + nv);
            nv = next_in_ring(pv, row, nrow);
            pv = ring[s];
        }
        ring[n - 1] = nv;
        memcpy(adj_out + 6 * idx, ring, sizeof ring);
    }
}

/* ------------------------------------------------------------------------- */
/* erosion.py:76-99.  One sweep: w = r + 0.0005*(#higher - #lower).           */
NXO_API void nxo_erosion_iteration1(int64_t V, const int32_t *adj, const double *r, double *w)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < V; ++i) {
        double me = r[i], amt = 0;
        for (int q = 0; q < 6; ++q) {
            int32_t n = adj[6 * i + q];
            if (n == -1) continue;
            if (r[n] > me) amt += 0.0005;
            else if (r[n] < me) amt -= 0.0005;
        }
        w[i] = me + amt;
    }
}

/* erosion.py:42-73: driver; heights updated in place every pass. */
NXO_API void nxo_erode_terrain1(int64_t V, const int32_t *adj, double *h, int num_iter)
{
    if (num_iter <= 0) num_iter = 1;
    double *w = (double *)malloc((size_t)V * sizeof(double));
    for (int it = 0; it < num_iter; ++it) {
        nxo_erosion_iteration1(V, adj, h, w);
        memcpy(h, w, (size_t)V * sizeof(double));
    }
    free(w);
}

/* erosion.py:34-40, 197-279.  One sweep of the live variant; reads the old
 * h/wat/sed, writes new values back in place (through private buffers).      */
NXO_API void nxo_erosion_iteration3(int64_t V, const double *verts, const int32_t *adj,
                                    double *h, double *wat, double *sed)
{
    const double evaporation = 0.1 / 320, solubility = 0.01 / 320, capacity = 0.2 / 320;
    double *hb = (double *)malloc((size_t)V * sizeof(double));
    double *wb = (double *)malloc((size_t)V * sizeof(double));
    double *sb = (double *)malloc((size_t)V * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < V; ++i) {
        double me = h[i], sed_amt = sed[i], wat_amt = wat[i];
        const double *pi = verts + 3 * i;
        for (int q = 0; q < 6; ++q) {
            int32_t n = adj[6 * i + q];
            if (n == -1) continue;
            const double *pn = verts + 3 * (int64_t)n;
            double ax = pi[0] - pn[0], ay = pi[1] - pn[1], az = pi[2] - pn[2];
            double d = sqrt(ax * ax + ay * ay + az * az);
            double slope = (h[n] - me) / (d + 0.00001);
            if (slope > 0)      { sed_amt += solubility * wat[n]; wat_amt += wat[n] * d; }
            else if (slope < 0) { sed_amt -= solubility * wat[n]; wat_amt -= wat[n] * d; }
        }
        double hn = me - sed_amt;
        double sn = sed[i] + sed_amt;
        double wn = wat[i] + (wat_amt - wat_amt * evaporation);
        if (sn > capacity * wn) {
            hn += sn - capacity * wn;
            sn -= sn - capacity * wn;
        }
        hb[i] = hn; sb[i] = sn; wb[i] = wn;
    }
    memcpy(h, hb, (size_t)V * sizeof(double));
    memcpy(wat, wb, (size_t)V * sizeof(double));
    memcpy(sed, sb, (size_t)V * sizeof(double));
    free(hb); free(wb); free(sb);
}

/* erosion.py:172-192: water/sediment start at zero, rain added before every
 * sweep.  wat/sed may be NULL (reference discards them); if given they are
 * zeroed first and hold the final state.                                      */
NXO_API void nxo_erode_terrain3(int64_t V, const double *verts, const int32_t *adj, double *h,
                                int num_iter, double *wat_out, double *sed_out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (num_iter <= 0) num_iter = 1;
    double *wat = wat_out ? wat_out : (double *)malloc((size_t)V * sizeof(double));
    double *sed = sed_out ? sed_out : (double *)malloc((size_t)V * sizeof(double));
    memset(wat, 0, (size_t)V * sizeof(double));
    memset(sed, 0, (size_t)V * sizeof(double));
    const double rain = 0.3 / 320;
    for (int it = 0; it < num_iter; ++it) {
        for (int64_t i = 0; i < V; ++i) wat[i] += rain;
        nxo_erosion_iteration3(V, verts, adj, h, wat, sed);
    }
    if (!wat_out) free(wat);
    if (!sed_out) free(sed);
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must use the host's cores */
NXO_API void nxo_set_num_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
}

NXO_API int nxo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

NXO_API int nxo_version(void) { return 1; }

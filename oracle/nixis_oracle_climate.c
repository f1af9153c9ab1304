/* CPU ORACLE of the per-vertex climate kernels -- TEST INFRASTRUCTURE, never shipped, never on the
 * product path (only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this).
 *
 * Restates climate.py:193-201, 345-597 and util.py:59-88 of /root/reference in IEEE double / float
 * exactly as numba types them (float32 accumulators, float64 trigonometry from libm, no FMA
 * contraction: build with -ffp-contract=off).  Pinned bit-for-bit to tests/golden/climate.npz, which
 * was produced by running the unmodified reference (tests/golden/gen_golden.py gen_climate).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define NXO_API __attribute__((visibility("default")))
static const double PI = 3.141592653589793;

/* util.py:59-75  lat = degrees(arcsin(clamp(z / r))), lon = degrees(arctan2(y, x)) */
static inline double degrees_(double x) { return x * (180.0 / PI); }
static inline void xyz2latlon(double x, double y, double z, double r, double *lat, double *lon)
{
    double q = z / r;
    if (!(q > -1.0)) q = -1.0;      /* max(q, -1) */
    if (!(q < 1.0)) q = 1.0;        /* min(.., 1) */
    *lat = degrees_(asin(q));
    *lon = degrees_(atan2(y, x));
}

/* climate.py:193-201 */
NXO_API double nxo_seasonal_tilt(double axial_tilt, double degrees)
{
    return sin(degrees * PI / 180) * axial_tilt;
}

/* climate.py:345-372: alt_intensity is the literal 0, so h2 = rescale(altitudes, 0, 0) is all
 * zeros (NaN when the altitudes are constant) and the `h2[v] > 0` branch is never taken; the
 * branch is kept for fidelity. */
NXO_API void nxo_assign_surface_temp(int64_t n, const double *verts, const double *altitudes, double radius,
                                     double tilt, float *out)
{
    double mn = altitudes[0], mx = altitudes[0];
    for (int64_t i = 1; i < n; ++i) { if (altitudes[i] < mn) mn = altitudes[i]; if (altitudes[i] > mx) mx = altitudes[i]; }
    const double range = mx - mn, alt_intensity = 0.0, new_range = alt_intensity - 0.0;
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < n; ++v) {
        const double h2 = ((altitudes[v] - mn) / range) * new_range + 0.0;
        double lat, lon;
        xyz2latlon(verts[3 * v], verts[3 * v + 1], verts[3 * v + 2], radius, &lat, &lon);
        const double d = lat - tilt;
        double c = cos(fabs(d) * PI / 180);
        if (!(c > 0.0)) c = 0.0;
        out[v] = (float)(h2 > 0 ? c - fabs(h2) + 0.1 : c);
    }
}

/* climate.py:415-448 sample_insolation: arr (float32) += one rotation's term, n_rot times with
 * rotation = rot0 + i * rot_step accumulated the way brute_daily_insolation / calc_insolation_slice
 * do it (`rotation += rot_amt`, climate.py:466-470, 520-523). */
NXO_API void nxo_sample_insolation(int64_t n, float *arr, const double *verts, double radius, double rot0,
                                   double rot_step, int n_rot, double tilt)
{
    double rotation = rot0;
    const double ct = cos(tilt * PI / 180), st = sin(tilt * PI / 180);
    for (int i = 0; i < n_rot; ++i) {
        const double cr = cos(rotation * PI / 180), sr = sin(rotation * PI / 180);
#pragma omp parallel for schedule(static)
        for (int64_t v = 0; v < n; ++v) {
            const double x = verts[3 * v], y = verts[3 * v + 1], z = verts[3 * v + 2];
            const double rx = x * cr - y * sr;
            const double ry = x * sr + y * cr;
            const double rz = z;
            const double tx = rx * ct + rz * st;
            const double ty = ry;
            const double tz = rz * ct - rx * st;
            double lat, lon;
            xyz2latlon(tx, ty, tz, radius, &lat, &lon);
            double a = cos(fabs(lat) * PI / 180), b = cos(lon * PI / 180);
            if (!(a > 0.0)) a = 0.0;
            if (!(b > 0.0)) b = 0.0;
            arr[v] = (float)((double)arr[v] + a * b);
        }
        rotation += rot_step;
    }
}

/* climate.py:503-536 calc_insolation_slice: 181 vertices at integer latitudes on lon 0; the loop
 * `for i in range(-90, 91): verts[i] = ...` stores negative latitudes through negative indices, so
 * result[0..90] = lat 0..90 and result[91..180] = lat -90..-1. */
NXO_API void nxo_insolation_slice(double radius, double tilt, float *result /*[181]*/)
{
    double verts[181 * 3];
    for (int i = -90; i <= 90; ++i) {
        const int row = i < 0 ? 181 + i : i;
        const double lat = (double)i, lon = 0.0;
        verts[3 * row + 0] = radius * cos(lat * (PI / 180)) * cos(lon * (PI / 180));   /* util.py:85-87 */
        verts[3 * row + 1] = radius * cos(lat * (PI / 180)) * sin(lon * (PI / 180));
        verts[3 * row + 2] = radius * sin(lat * (PI / 180));
    }
    for (int i = 0; i < 181; ++i) result[i] = 0.0f;
    nxo_sample_insolation(181, result, verts, radius, -180.0, 360.0 / 360, 360, tilt);
}

/* climate.py:551-577 interpolate_insolation (negative `lower` / `upper` index from the end) */
NXO_API void nxo_interpolate_insolation(int64_t n, const double *verts, const float *table /*[181]*/,
                                        float *insolation, double radius)
{
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < n; ++v) {
        double lat, lon;
        xyz2latlon(verts[3 * v], verts[3 * v + 1], verts[3 * v + 2], radius, &lat, &lon);
        const int64_t lower = (int64_t)floor(lat), upper = (int64_t)ceil(lat);
        const float tl = table[lower < 0 ? 181 + lower : lower], tu = table[upper < 0 ? 181 + upper : upper];
        if (lower == upper) insolation[v] = tl;
        else {
            const float diff = tu - tl;                              /* float32 - float32 */
            insolation[v] = (float)((double)tl + (lat - (double)lower) * ((double)diff / (double)(upper - lower)));
        }
    }
}

/* climate.py:539-548 */
NXO_API void nxo_daily_insolation(int64_t n, const double *verts, double radius, double tilt, float *out)
{
    float table[181];
    nxo_insolation_slice(radius, tilt, table);
    nxo_interpolate_insolation(n, verts, table, out, radius);
}

/* climate.py:579-597: 360 days, annual (float32) += daily (float32) */
NXO_API void nxo_yearly_insolation(int64_t n, const double *verts, double radius, double axial_tilt, float *annual)
{
    float *daily = (float *)malloc(sizeof(float) * (size_t)n);
    for (int64_t v = 0; v < n; ++v) annual[v] = 0.0f;
    for (int x = 0; x < 360; ++x) {
        const double tilt = nxo_seasonal_tilt(axial_tilt, (double)x);
        nxo_daily_insolation(n, verts, radius, tilt, daily);
        for (int64_t v = 0; v < n; ++v) annual[v] = annual[v] + daily[v];
    }
    free(daily);
}

// Reference-exact 3-D fBm: IEEE double, no FMA contraction (this file is compiled with -fmad=false),
// every operation in the order the reference performs it (opensimplex.py:266-759, terrain.py:12-47).
// Given the same float64 vertices the result is BIT-IDENTICAL to the reference's numba output
// (tests/test_gpu_parity.py::test_sample_octaves_exact_mode_bit_identical).  It is the parity mode of
// terrain.sample_octaves(exact=True); the FP32 kernel (nxb_noise.cu) is the throughput mode.  B200
// runs FP64 at half the FP32 rate, so this costs ~4x the fast kernel -- 1 % of a whole terrain step.
//
// Structure: lattice OFFSETS.  A contribution is base + (i,j,k); its displacement is
// (d0 - i) - m*(1/3) with m = i+j+k, two roundings, as the reference spells `dx0 - 1 - 2*SQUISH`.
// Two reference spellings subtract the integer AFTER the squish term (`dy_ext1 -= 1`,
// opensimplex.py:438-443; `dx_ext1 -= 2`, :683-691); `late` / `lateamt` reproduce that order.
#include "nxb_noise.cuh"

struct Ext3 { int ox, oy, oz; int late; int lateamt; };   // late: axis index + 1 (0 = none)

struct Ctx64 {
    const uint8_t *perm, *grad;
    long long bx, by, bz;
    double dx0, dy0, dz0;
    double v;
};

__device__ __forceinline__ double sqm(int m)
{
    // 0, 1/3, 2*(1/3), 3*(1/3) evaluated in double exactly like the Python constants
    const double SQ = 1.0 / 3;
    return m == 0 ? 0.0 : (m == 1 ? SQ : (m == 2 ? 2 * SQ : 3 * SQ));
}

__device__ __forceinline__ void add64(Ctx64 &c, int i, int j, int k, double dx, double dy, double dz)
{
    double at = 2 - dx * dx - dy * dy - dz * dz;
    if (at > 0) {
        unsigned h = c.perm[(c.bx + i) & 255];
        h = c.perm[(h + c.by + j) & 255];
        const unsigned g = c.grad[(h + c.bz + k) & 255];          // bit0 x neg, bit1 y neg, bit2 z neg, bits3-4 axis of the 11
        const unsigned ax = g >> 3;
        const double gx = ((g & 1) ? -1.0 : 1.0) * (ax == 0 ? 11.0 : 4.0);
        const double gy = ((g & 2) ? -1.0 : 1.0) * (ax == 1 ? 11.0 : 4.0);
        const double gz = ((g & 4) ? -1.0 : 1.0) * (ax == 2 ? 11.0 : 4.0);
        at *= at;
        c.v += at * at * (gx * dx + gy * dy + gz * dz);
    }
}

__device__ __forceinline__ void corner64(Ctx64 &c, int i, int j, int k)
{
    const double s = sqm(i + j + k);
    add64(c, i, j, k, c.dx0 - i - s, c.dy0 - j - s, c.dz0 - k - s);
}

__device__ __forceinline__ double disp64(double d0, int o, double s, bool late, int amt)
{
    if (late) return (d0 - (o - amt)) - s - amt;
    return d0 - o - s;
}

__device__ __forceinline__ void extra64(Ctx64 &c, const Ext3 &e)
{
    const double s = sqm(e.ox + e.oy + e.oz);
    add64(c, e.ox, e.oy, e.oz,
          disp64(c.dx0, e.ox, s, e.late == 1, e.lateamt),
          disp64(c.dy0, e.oy, s, e.late == 2, e.lateamt),
          disp64(c.dz0, e.oz, s, e.late == 3, e.lateamt));
}

__device__ __forceinline__ long long ffloor64(double x)
{
    long long xi = (long long)x;                     // opensimplex.py:18-21
    return x < (double)xi ? xi - 1 : xi;
}

__device__ double noise3_f64(double x, double y, double z, const uint8_t *perm, const uint8_t *grad)
{
    const double so = (x + y + z) * (-1.0 / 6);
    const double xs = x + so, ys = y + so, zs = z + so;
    Ctx64 c;
    c.perm = perm; c.grad = grad;
    c.bx = ffloor64(xs); c.by = ffloor64(ys); c.bz = ffloor64(zs);
    const double qo = (double)(c.bx + c.by + c.bz) * (1.0 / 3);
    const double fx = xs - (double)c.bx, fy = ys - (double)c.by, fz = zs - (double)c.bz;
    const double fsum = fx + fy + fz;
    c.dx0 = x - ((double)c.bx + qo); c.dy0 = y - ((double)c.by + qo); c.dz0 = z - ((double)c.bz + qo);
    c.v = 0;
    Ext3 e0 = {0, 0, 0, 0, 0}, e1 = {0, 0, 0, 0, 0};

    if (fsum <= 1) {
        int ap = 1, bp = 2; double as = fx, bs = fy;
        if (as >= bs && fz > bs) { bs = fz; bp = 4; } else if (as < bs && fz > as) { as = fz; ap = 4; }
        const double w = 1 - fsum;
        if (w > as || w > bs) {
            const int cc = (bs > as) ? bp : ap;
            e0.ox = e1.ox = cc & 1; e0.oy = e1.oy = (cc >> 1) & 1; e0.oz = e1.oz = (cc >> 2) & 1;
            if (!(cc & 1)) e0.ox = -1;
            if (!(cc & 2)) { if (!(cc & 1)) e1.oy = -1; else e0.oy = -1; }
            if (!(cc & 4)) e1.oz = -1;
        } else {
            const int cc = ap | bp;
            e0.ox = cc & 1; e0.oy = (cc >> 1) & 1; e0.oz = (cc >> 2) & 1;
            e1.ox = 2 * e0.ox - 1; e1.oy = 2 * e0.oy - 1; e1.oz = 2 * e0.oz - 1;
        }
        corner64(c, 0, 0, 0); corner64(c, 1, 0, 0); corner64(c, 0, 1, 0); corner64(c, 0, 0, 1);
    } else if (fsum >= 2) {
        int ap = 6, bp = 5; double as = fx, bs = fy;
        if (as <= bs && fz < bs) { bs = fz; bp = 3; } else if (as > bs && fz < as) { as = fz; ap = 3; }
        const double w = 3 - fsum;
        if (w < as || w < bs) {
            const int cc = (bs < as) ? bp : ap;
            e0.ox = e1.ox = cc & 1; e0.oy = e1.oy = (cc >> 1) & 1; e0.oz = e1.oz = (cc >> 2) & 1;
            if (cc & 1) e0.ox = 2;
            if (cc & 2) { if (cc & 1) { e1.oy = 2; e1.late = 2; e1.lateamt = 1; } else { e0.oy = 2; e0.late = 2; e0.lateamt = 1; } }
            if (cc & 4) e1.oz = 2;
        } else {
            const int cc = ap & bp;
            e0.ox = cc & 1; e0.oy = (cc >> 1) & 1; e0.oz = (cc >> 2) & 1;
            e1.ox = 2 * e0.ox; e1.oy = 2 * e0.oy; e1.oz = 2 * e0.oz;
        }
        corner64(c, 1, 1, 0); corner64(c, 1, 0, 1); corner64(c, 0, 1, 1); corner64(c, 1, 1, 1);
    } else {
        double as, bs, sc; int ap, bp; bool afar, bfar;
        const double p1 = fx + fy, p2 = fx + fz, p3 = fy + fz;
        if (p1 > 1) { as = p1 - 1; ap = 3; afar = true; } else { as = 1 - p1; ap = 4; afar = false; }
        if (p2 > 1) { bs = p2 - 1; bp = 5; bfar = true; } else { bs = 1 - p2; bp = 2; bfar = false; }
        if (p3 > 1) {
            sc = p3 - 1;
            if (as <= bs && as < sc) { ap = 6; afar = true; } else if (as > bs && bs < sc) { bp = 6; bfar = true; }
        } else {
            sc = 1 - p3;
            if (as <= bs && as < sc) { ap = 1; afar = false; } else if (as > bs && bs < sc) { bp = 1; bfar = false; }
        }
        if (afar == bfar) {
            if (afar) {
                e0.ox = e0.oy = e0.oz = 1;
                const int cc = ap & bp;
                if (cc & 1) e1.ox = 2; else if (cc & 2) e1.oy = 2; else e1.oz = 2;
            } else {
                const int cc = ap | bp;
                e1.ox = e1.oy = e1.oz = 1;
                if (!(cc & 1)) e1.ox = -1; else if (!(cc & 2)) e1.oy = -1; else e1.oz = -1;
            }
        } else {
            const int c1 = afar ? ap : bp, c2 = afar ? bp : ap;
            e0.ox = e0.oy = e0.oz = 1;
            if (!(c1 & 1)) e0.ox = -1; else if (!(c1 & 2)) e0.oy = -1; else e0.oz = -1;
            if (c2 & 1) { e1.ox = 2; e1.late = 1; } else if (c2 & 2) { e1.oy = 2; e1.late = 2; } else { e1.oz = 2; e1.late = 3; }
            e1.lateamt = 2;
        }
        corner64(c, 1, 0, 0); corner64(c, 0, 1, 0); corner64(c, 0, 0, 1);
        corner64(c, 1, 1, 0); corner64(c, 1, 0, 1); corner64(c, 0, 1, 1);
    }
    extra64(c, e0);
    extra64(c, e1);
    return c.v / 103;
}

#define NXB_MAX_OCT64 32
struct Fbm64Params {
    double nr[NXB_MAX_OCT64];       // n_freq / world_radius   (terrain.py:43)
    double ns[NXB_MAX_OCT64];       // n_amp / world_radius
    double radius, scale;
    int n_oct;
};

__global__ void __launch_bounds__(256)
fbm3_f64_kernel(const NxbTables *__restrict__ tab, const double *__restrict__ verts, int64_t n,
                const __grid_constant__ Fbm64Params prm, const double *__restrict__ init, double *__restrict__ out)
{
    __shared__ __align__(16) uint8_t s_tab[NXB_TABLE_BYTES];
    nxb_stage_tables(tab, s_tab);
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        // nixis.py:249 `points *= world_radius` (scale) happens before sample_octaves sees the vertices
        const double px = verts[3 * v] * prm.scale, py = verts[3 * v + 1] * prm.scale, pz = verts[3 * v + 2] * prm.scale;
        double acc = init ? init[v] : 0.0;
        for (int o = 0; o < prm.n_oct; ++o) {
            const double e = noise3_f64(px * prm.nr[o], py * prm.nr[o], pz * prm.nr[o], s_tab, s_tab + 256);
            acc = acc + (e + 1) * 0.5 * prm.ns[o] * prm.radius;        // terrain.py:28, :43
        }
        out[v] = acc;
    }
}

NXB_API int nxb_fbm3_f64(void *tables, const double *verts, int64_t n, int n_oct,
                         const double *nr_host, const double *ns_host, double radius, double scale,
                         const double *init, double *out, void *stream)
{
    NXB_ARG(tables && n >= 0 && n_oct >= 0 && n_oct <= NXB_MAX_OCT64);
    if (n == 0) return NXB_OK;
    NXB_ARG(verts && out && (n_oct == 0 || (nr_host && ns_host)));
    Fbm64Params prm;
    for (int o = 0; o < NXB_MAX_OCT64; ++o) { prm.nr[o] = o < n_oct ? nr_host[o] : 0.0; prm.ns[o] = o < n_oct ? ns_host[o] : 0.0; }
    prm.radius = radius; prm.scale = scale; prm.n_oct = n_oct;
    int grid = nxb_grid_resident(fbm3_f64_kernel, 256, 0, (n + 255) / 256);
    fbm3_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const NxbTables *)tables, verts, n, prm, init, out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// ---------------------------------------------------------------------------------------------
// Reference-exact erosion sweep (erosion.py:34-40, 197-279): float64 state and positions, no FMA
// contraction (this file is built with -fmad=false), neighbour loop in the table's slot order,
// sqrt / divide in IEEE double.  One thread per vertex with global gathers: the parity mode of
// erosion.erode_terrain3(exact=True); the FP32 tile-plan kernel (nxb_erosion.cu) is the throughput
// mode.  `rain` is added to every water value read, which is exactly the reference's
// `water += rain_amount` (erosion.py:182-183) followed by the sweep.
__global__ void __launch_bounds__(256)
erode3_f64_kernel(const double *__restrict__ nodes, const int32_t *__restrict__ adj,
                  const double *__restrict__ h_in, const double *__restrict__ w_in, const double *__restrict__ s_in,
                  double *__restrict__ h_out, double *__restrict__ w_out, double *__restrict__ s_out,
                  int64_t n, double rain)
{
    const double evaporation = 0.1 / 320, solubility = 0.01 / 320, capacity = 0.2 / 320;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double me = h_in[i], sed_i = s_in[i], wat_i = w_in[i] + rain;
        double sed_amt = sed_i, wat_amt = wat_i;
        const double px = nodes[3 * i], py = nodes[3 * i + 1], pz = nodes[3 * i + 2];
        for (int q = 0; q < 6; ++q) {
            const int32_t nb = adj[6 * i + q];
            if (nb == -1) continue;
            const double ax = px - nodes[3 * (int64_t)nb], ay = py - nodes[3 * (int64_t)nb + 1], az = pz - nodes[3 * (int64_t)nb + 2];
            const double d = sqrt(ax * ax + ay * ay + az * az);
            const double wn = w_in[nb] + rain;
            const double slope = (h_in[nb] - me) / (d + 0.00001);
            if (slope > 0)      { sed_amt += solubility * wn; wat_amt += wn * d; }
            else if (slope < 0) { sed_amt -= solubility * wn; wat_amt -= wn * d; }
        }
        double hh = me - sed_amt;
        double ss = sed_i + sed_amt;
        const double ww = wat_i + (wat_amt - wat_amt * evaporation);
        if (ss > capacity * ww) {
            hh += ss - capacity * ww;
            ss -= ss - capacity * ww;
        }
        h_out[i] = hh; w_out[i] = ww; s_out[i] = ss;
    }
}

NXB_API int nxb_erode3_step_f64(const double *nodes, const int32_t *adj,
                                const double *h_in, const double *w_in, const double *s_in,
                                double *h_out, double *w_out, double *s_out, int64_t n, double rain, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(nodes && adj && h_in && w_in && s_in && h_out && w_out && s_out && h_in != h_out && w_in != w_out && s_in != s_out);
    int grid = nxb_grid_resident(erode3_f64_kernel, 256, 0, (n + 255) / 256);
    erode3_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(nodes, adj, h_in, w_in, s_in, h_out, w_out, s_out, n, rain);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

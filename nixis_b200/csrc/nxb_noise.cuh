// OpenSimplex (legacy KdotJPG) noise in FP32 for sm_100a -- device functions.
//
// Follows the candidate-selection logic of the reference exactly
// (opensimplex.py:153-254 2-D, :266-759 3-D): the legacy algorithm is NOT a full
// lattice sum -- it evaluates 4/6 simplex corners plus 2 "extra" lattice points,
// and dropping or adding candidates changes the value by up to 5e-5 -- so the
// same comparisons pick the same candidates here; only the arithmetic is FP32.
//
// Every contributing lattice point is base + (i,j,k); with m = i+j+k its
// displacement is d0 - (i,j,k) - m/3 and its weight max(0, 2-|d|^2)^4.
//
// Tables (built by nxb_tables_create, staged into shared memory by each CTA):
//   perm8[256]  : the reference's perm (opensimplex.py:90-107), one byte each
//   grad8[256]  : for the LAST hash level h: a descriptor of GRADIENTS_3D[pgi[h]]:
//                 bit0 x negative, bit1 y negative, bit2 z negative, bits3-4 the
//                 axis that carries 11 (others carry 4)   (opensimplex.py:52-61,123-130)
// The gradient dot product is decoded arithmetically (sign flips by XOR, then
// 4*(sum) + 7*axis component) so no per-lane gradient loads are needed.
#pragma once
#include "nxb_common.cuh"

struct NxbTables {
    uint8_t perm8[256];
    uint8_t grad8[256];     // 3-D gradient descriptor of pgi[h]
    uint8_t grad4[256];     // 4-D: perm[h] & 0xFC  (opensimplex.py:133-141)
    uint8_t grad2[256];     // 2-D: perm[h] & 0x0E  (opensimplex.py:115-120)
};

#define NXB_TABLE_BYTES 1024

// Stage the tables into shared memory (all threads of the CTA participate).
__device__ __forceinline__ void nxb_stage_tables(const NxbTables *__restrict__ g, uint8_t *s)
{
    const uint32_t *src = reinterpret_cast<const uint32_t *>(g);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s);
    for (int i = threadIdx.x; i < NXB_TABLE_BYTES / 4; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

__device__ __forceinline__ int nxb_fastfloor(float x)
{
    // opensimplex.py:18-21 (trunc, then step down for negatives) == floor for finite x
    return __float2int_rd(x);
}

// ---------------------------------------------------------------------------------
// 3-D
struct Noise3Ctx {
    const uint8_t *perm;   // shared
    const uint8_t *grad;   // shared
    int xb, yb, zb;
    float dx0, dy0, dz0;
    float value;
};

__device__ __forceinline__ float nxb_grad3_dot(uint32_t g, float dx, float dy, float dz)
{
    // g: bit0 x neg, bit1 y neg, bit2 z neg, bits 3-4 axis of the 11
    float sx = __uint_as_float(__float_as_uint(dx) ^ (g << 31));
    float sy = __uint_as_float(__float_as_uint(dy) ^ ((g << 30) & 0x80000000u));
    float sz = __uint_as_float(__float_as_uint(dz) ^ ((g << 29) & 0x80000000u));
    uint32_t ax = g >> 3;
    float da = ax == 0 ? sx : (ax == 1 ? sy : sz);
    return fmaf(7.0f, da, 4.0f * (sx + sy + sz));
}

__device__ __forceinline__ void nxb_add3(Noise3Ctx &c, int i, int j, int k, float dx, float dy, float dz)
{
    float at = 2.0f - dx * dx - dy * dy - dz * dz;
    if (at > 0.0f) {
        uint32_t h = c.perm[(c.xb + i) & 255];
        h = c.perm[(h + c.yb + j) & 255];
        uint32_t g = c.grad[(h + c.zb + k) & 255];
        at *= at;
        c.value = fmaf(at * at, nxb_grad3_dot(g, dx, dy, dz), c.value);
    }
}

#define NXB_SQ3 0.33333333333333333f

template <int I, int J, int K>
__device__ __forceinline__ void nxb_corner3(Noise3Ctx &c)
{
    constexpr float m = (float)(I + J + K) * (1.0f / 3.0f);
    nxb_add3(c, I, J, K, c.dx0 - ((float)I + m), c.dy0 - ((float)J + m), c.dz0 - ((float)K + m));
}

__device__ __forceinline__ void nxb_extra3(Noise3Ctx &c, int i, int j, int k)
{
    float m = (float)(i + j + k) * (1.0f / 3.0f);
    nxb_add3(c, i, j, k, c.dx0 - (float)i - m, c.dy0 - (float)j - m, c.dz0 - (float)k - m);
}

// opensimplex.py:266-759.  perm/grad are the shared-memory tables.
__device__ __forceinline__ float nxb_noise3(float x, float y, float z, const uint8_t *perm, const uint8_t *grad)
{
    float so = (x + y + z) * (-1.0f / 6.0f);
    float xs = x + so, ys = y + so, zs = z + so;
    Noise3Ctx c;
    c.perm = perm; c.grad = grad;
    c.xb = nxb_fastfloor(xs); c.yb = nxb_fastfloor(ys); c.zb = nxb_fastfloor(zs);
    float fxb = (float)c.xb, fyb = (float)c.yb, fzb = (float)c.zb;
    float qo = (fxb + fyb + fzb) * (1.0f / 3.0f);
    float fx = xs - fxb, fy = ys - fyb, fz = zs - fzb;
    float fsum = fx + fy + fz;
    c.dx0 = x - (fxb + qo); c.dy0 = y - (fyb + qo); c.dz0 = z - (fzb + qo);
    c.value = 0.0f;
    // extras as lattice offsets
    int e0x = 0, e0y = 0, e0z = 0, e1x = 0, e1y = 0, e1z = 0;

    if (fsum <= 1.0f) {                         // tetrahedron at (0,0,0)
        int ap = 1, bp = 2; float as = fx, bs = fy;
        if (as >= bs && fz > bs) { bs = fz; bp = 4; }
        else if (as < bs && fz > as) { as = fz; ap = 4; }
        float w = 1.0f - fsum;
        if (w > as || w > bs) {
            int cc = (bs > as) ? bp : ap;
            e0x = e1x = cc & 1; e0y = e1y = (cc >> 1) & 1; e0z = e1z = (cc >> 2) & 1;
            if (!(cc & 1)) e0x = -1;
            if (!(cc & 2)) { if (!(cc & 1)) e1y = -1; else e0y = -1; }
            if (!(cc & 4)) e1z = -1;
        } else {
            int cc = ap | bp;
            e0x = cc & 1; e0y = (cc >> 1) & 1; e0z = (cc >> 2) & 1;
            e1x = 2 * e0x - 1; e1y = 2 * e0y - 1; e1z = 2 * e0z - 1;
        }
        nxb_corner3<0, 0, 0>(c);
        nxb_corner3<1, 0, 0>(c);
        nxb_corner3<0, 1, 0>(c);
        nxb_corner3<0, 0, 1>(c);
    } else if (fsum >= 2.0f) {                  // tetrahedron at (1,1,1)
        int ap = 6, bp = 5; float as = fx, bs = fy;
        if (as <= bs && fz < bs) { bs = fz; bp = 3; }
        else if (as > bs && fz < as) { as = fz; ap = 3; }
        float w = 3.0f - fsum;
        if (w < as || w < bs) {
            int cc = (bs < as) ? bp : ap;
            e0x = e1x = cc & 1; e0y = e1y = (cc >> 1) & 1; e0z = e1z = (cc >> 2) & 1;
            if (cc & 1) e0x = 2;
            if (cc & 2) { if (cc & 1) e1y = 2; else e0y = 2; }
            if (cc & 4) e1z = 2;
        } else {
            int cc = ap & bp;
            e0x = cc & 1; e0y = (cc >> 1) & 1; e0z = (cc >> 2) & 1;
            e1x = 2 * e0x; e1y = 2 * e0y; e1z = 2 * e0z;
        }
        nxb_corner3<1, 1, 0>(c);
        nxb_corner3<1, 0, 1>(c);
        nxb_corner3<0, 1, 1>(c);
        nxb_corner3<1, 1, 1>(c);
    } else {                                    // octahedron
        float as, bs, sc; int ap, bp; bool afar, bfar;
        float p1 = fx + fy;
        if (p1 > 1.0f) { as = p1 - 1.0f; ap = 3; afar = true; } else { as = 1.0f - p1; ap = 4; afar = false; }
        float p2 = fx + fz;
        if (p2 > 1.0f) { bs = p2 - 1.0f; bp = 5; bfar = true; } else { bs = 1.0f - p2; bp = 2; bfar = false; }
        float p3 = fy + fz;
        if (p3 > 1.0f) {
            sc = p3 - 1.0f;
            if (as <= bs && as < sc) { ap = 6; afar = true; }
            else if (as > bs && bs < sc) { bp = 6; bfar = true; }
        } else {
            sc = 1.0f - p3;
            if (as <= bs && as < sc) { ap = 1; afar = false; }
            else if (as > bs && bs < sc) { bp = 1; bfar = false; }
        }
        if (afar == bfar) {
            if (afar) {
                e0x = e0y = e0z = 1;
                int cc = ap & bp;
                if (cc & 1) e1x = 2; else if (cc & 2) e1y = 2; else e1z = 2;
            } else {
                int cc = ap | bp;
                e1x = e1y = e1z = 1;
                if (!(cc & 1)) e1x = -1; else if (!(cc & 2)) e1y = -1; else e1z = -1;
            }
        } else {
            int c1 = afar ? ap : bp, c2 = afar ? bp : ap;
            e0x = e0y = e0z = 1;
            if (!(c1 & 1)) e0x = -1; else if (!(c1 & 2)) e0y = -1; else e0z = -1;
            if (c2 & 1) e1x = 2; else if (c2 & 2) e1y = 2; else e1z = 2;
        }
        nxb_corner3<1, 0, 0>(c);
        nxb_corner3<0, 1, 0>(c);
        nxb_corner3<0, 0, 1>(c);
        nxb_corner3<1, 1, 0>(c);
        nxb_corner3<1, 0, 1>(c);
        nxb_corner3<0, 1, 1>(c);
    }
    nxb_extra3(c, e0x, e0y, e0z);
    nxb_extra3(c, e1x, e1y, e1z);
    return c.value * (1.0f / 103.0f);
}

// ---------------------------------------------------------------------------------
// 2-D (opensimplex.py:115-120, 153-254) -- API parity only
__device__ __forceinline__ void nxb_add2(float &v, const uint8_t *perm, const uint8_t *g2, int xb, int yb, float dx, float dy)
{
    float at = 2.0f - dx * dx - dy * dy;
    if (at > 0.0f) {
        uint32_t idx = g2[(perm[xb & 255] + yb) & 255];          // perm[..] & 0x0E
        // GRADIENTS_2D: idx/2 in 0..7: (5,2),(2,5),(-5,2),(-2,5),(5,-2),(2,-5),(-5,-2),(-2,-5)
        uint32_t q = idx >> 1;
        float gx = (q & 1) ? 2.0f : 5.0f, gy = (q & 1) ? 5.0f : 2.0f;
        if (q & 2) gx = -gx;
        if (q & 4) gy = -gy;
        at *= at;
        v = fmaf(at * at, gx * dx + gy * dy, v);
    }
}

__device__ __forceinline__ float nxb_noise2(float x, float y, const uint8_t *perm, const uint8_t *g2)
{
    const float ST = -0.211324865405187f, SQ = 0.366025403784439f;
    float so = (x + y) * ST;
    float xs = x + so, ys = y + so;
    int xb = nxb_fastfloor(xs), yb = nxb_fastfloor(ys);
    float qo = (float)(xb + yb) * SQ;
    float fx = xs - (float)xb, fy = ys - (float)yb, fsum = fx + fy;
    float dx0 = x - ((float)xb + qo), dy0 = y - ((float)yb + qo);
    float v = 0.0f;
    nxb_add2(v, perm, g2, xb + 1, yb, dx0 - 1.0f - SQ, dy0 - SQ);
    nxb_add2(v, perm, g2, xb, yb + 1, dx0 - SQ, dy0 - 1.0f - SQ);
    int ex, ey; float edx, edy;
    if (fsum <= 1.0f) {
        float fz = 1.0f - fsum;
        if (fz > fx || fz > fy) {
            if (fx > fy) { ex = xb + 1; ey = yb - 1; edx = dx0 - 1.0f; edy = dy0 + 1.0f; }
            else         { ex = xb - 1; ey = yb + 1; edx = dx0 + 1.0f; edy = dy0 - 1.0f; }
        } else { ex = xb + 1; ey = yb + 1; edx = dx0 - 1.0f - 2.0f * SQ; edy = dy0 - 1.0f - 2.0f * SQ; }
    } else {
        float fz = 2.0f - fsum;
        if (fz < fx || fz < fy) {
            if (fx > fy) { ex = xb + 2; ey = yb;     edx = dx0 - 2.0f - 2.0f * SQ; edy = dy0 - 2.0f * SQ; }
            else         { ex = xb;     ey = yb + 2; edx = dx0 - 2.0f * SQ;        edy = dy0 - 2.0f - 2.0f * SQ; }
        } else { ex = xb; ey = yb; edx = dx0; edy = dy0; }
        xb += 1; yb += 1;
        dx0 = dx0 - 1.0f - 2.0f * SQ; dy0 = dy0 - 1.0f - 2.0f * SQ;
    }
    nxb_add2(v, perm, g2, xb, yb, dx0, dy0);
    nxb_add2(v, perm, g2, ex, ey, edx, edy);
    return v * (1.0f / 47.0f);
}

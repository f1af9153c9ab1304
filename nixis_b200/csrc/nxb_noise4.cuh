// OpenSimplex 4-D (legacy KdotJPG) in FP32 for sm_100a -- opensimplex.py:133-141, 771-1958.
//
// Same candidate selection as the reference (5 or 10 simplex corners + 3 "extra" lattice points
// chosen by the reference's comparisons), arithmetic in FP32.  A contribution is base + o[4] with
// displacement d0 - o - mult*SQUISH_4D; `mult` is the multiplier the reference writes (normally the
// coordinate sum; opensimplex.py:1297-1321 is reproduced literally).  All small arrays are indexed
// with compile-time indices (unrolled) so they live in registers.
// Tables (shared memory): perm8[256] and grad4[256] = perm & 0xFC (index into GRADIENTS_4D).
// The reference never calls noise4d from its pipeline: this exists for BASELINE
// config 5 and API parity, and is not tuned like the 3-D kernel.
#pragma once
#include "nxb_common.cuh"

#define NXB_SQ4 0.309016994374947f
#define NXB_ST4 (-0.138196601125011f)

struct Ext4 { int o[4]; int mult; };

struct Noise4Ctx {
    const uint8_t *perm, *grad;
    int b[4];
    float d0[4];
    float v;
};

__device__ __forceinline__ void nxb_add4(Noise4Ctx &c, const Ext4 &e)
{
    const float sq = (float)e.mult * NXB_SQ4;
    float d[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) d[a] = c.d0[a] - (float)e.o[a] - sq;
    float at = 2.0f - d[0] * d[0] - d[1] * d[1] - d[2] * d[2] - d[3] * d[3];
    if (at > 0.0f) {
        uint32_t h = c.perm[(c.b[0] + e.o[0]) & 255];
        h = c.perm[(h + c.b[1] + e.o[1]) & 255];
        h = c.perm[(h + c.b[2] + e.o[2]) & 255];
        const uint32_t idx = c.grad[(h + c.b[3] + e.o[3]) & 255];     // perm & 0xFC
        const uint32_t ax = (idx >> 2) & 3u, q = idx >> 4;
        // GRADIENTS_4D (opensimplex.py:64-83): sign bit k of q negates axis k, member ax carries 3
        float dot = 0.0f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float s = ((q >> a) & 1u) ? -d[a] : d[a];
            dot += (ax == (uint32_t)a) ? 3.0f * s : s;
        }
        at *= at;
        c.v = fmaf(at * at, dot, c.v);
    }
}

__device__ __forceinline__ void nxb_corner4(Noise4Ctx &c, int code)
{
    Ext4 e;
#pragma unroll
    for (int a = 0; a < 4; ++a) e.o[a] = (code >> a) & 1;
    e.mult = e.o[0] + e.o[1] + e.o[2] + e.o[3];
    nxb_add4(c, e);
}

__device__ __forceinline__ int nxb_first_set(int c) { return (c & 1) ? 0 : ((c & 2) ? 1 : ((c & 4) ? 2 : 3)); }
__device__ __forceinline__ int nxb_first_clear(int c) { return !(c & 1) ? 0 : (!(c & 2) ? 1 : (!(c & 4) ? 2 : 3)); }

__device__ __forceinline__ void nxb_bits4(Ext4 &e, int c)
{
#pragma unroll
    for (int a = 0; a < 4; ++a) e.o[a] = (c >> a) & 1;
}
__device__ __forceinline__ void nxb_set_axis(Ext4 &e, int ax, int val)
{
#pragma unroll
    for (int a = 0; a < 4; ++a) if (a == ax) e.o[a] = val;
}

// opensimplex.py:1337-1381 / 1383-1431
__device__ __forceinline__ void nxb_pair_minus(int c, Ext4 &e0, Ext4 &e1)
{
    nxb_bits4(e0, c); nxb_bits4(e1, c);
    e0.mult = e1.mult = 1;
    if (!(c & 1)) e0.o[0] = -1;
    if (!(c & 2)) { if ((c & 1) == 1) e0.o[1] = -1; else e1.o[1] = -1; }
    if (!(c & 4)) { if ((c & 3) == 3) e0.o[2] = -1; else e1.o[2] = -1; }
    if (!(c & 8)) e1.o[3] = -1;
}
// opensimplex.py:1719-1763 / 1765-1815
__device__ __forceinline__ void nxb_pair_plus(int c, Ext4 &e0, Ext4 &e1)
{
    nxb_bits4(e0, c); nxb_bits4(e1, c);
    e0.mult = e1.mult = 3;
    if (c & 1) e0.o[0] = 2;
    if (c & 2) { if ((c & 1) == 0) e0.o[1] = 2; else e1.o[1] = 2; }
    if (c & 4) { if ((c & 3) == 0) e0.o[2] = 2; else e1.o[2] = 2; }
    if (c & 8) e1.o[3] = 2;
}

__device__ __forceinline__ float nxb_noise4(float x, float y, float z, float w, const uint8_t *perm, const uint8_t *grad)
{
    const float so = (x + y + z + w) * NXB_ST4;
    const float s[4] = {x + so, y + so, z + so, w + so};
    Noise4Ctx c;
    c.perm = perm; c.grad = grad; c.v = 0.0f;
    float f[4];
    float bsum = 0.0f;
#pragma unroll
    for (int a = 0; a < 4; ++a) { c.b[a] = __float2int_rd(s[a]); f[a] = s[a] - (float)c.b[a]; bsum += (float)c.b[a]; }
    const float qo = bsum * NXB_SQ4;
    const float in[4] = {x, y, z, w};
#pragma unroll
    for (int a = 0; a < 4; ++a) c.d0[a] = in[a] - ((float)c.b[a] + qo);
    const float fsum = f[0] + f[1] + f[2] + f[3];
    Ext4 e0, e1, e2;
    nxb_bits4(e0, 0); nxb_bits4(e1, 0); nxb_bits4(e2, 0);
    e0.mult = e1.mult = e2.mult = 0;

    if (fsum <= 1.0f) {                                 // pentachoron at (0,0,0,0)
        int ap = 1, bp = 2; float as = f[0], bs = f[1];
        if (as >= bs && f[2] > bs) { bs = f[2]; bp = 4; } else if (as < bs && f[2] > as) { as = f[2]; ap = 4; }
        if (as >= bs && f[3] > bs) { bs = f[3]; bp = 8; } else if (as < bs && f[3] > as) { as = f[3]; ap = 8; }
        const float u = 1.0f - fsum;
        if (u > as || u > bs) {
            const int cc = (bs > as) ? bp : ap;
            nxb_bits4(e0, cc); nxb_bits4(e1, cc); nxb_bits4(e2, cc);
            if (!(cc & 1)) e0.o[0] = -1;
            if (!(cc & 2)) { if ((cc & 1) == 1) e0.o[1] = -1; else e1.o[1] = -1; }
            if (!(cc & 4)) { if ((cc & 3) != 0) { if ((cc & 3) == 3) e0.o[2] = -1; else e1.o[2] = -1; } else e2.o[2] = -1; }
            if (!(cc & 8)) e2.o[3] = -1;
        } else {
            const int cc = ap | bp;
            nxb_bits4(e0, cc); nxb_bits4(e1, cc); nxb_bits4(e2, cc);
            e0.mult = 2; e1.mult = e2.mult = 1;
            if (!(cc & 1)) e1.o[0] = -1;
            if (!(cc & 2)) { if ((cc & 1) == 1) e1.o[1] = -1; else e2.o[1] = -1; }
            if (!(cc & 4)) { if ((cc & 3) == 3) e1.o[2] = -1; else e2.o[2] = -1; }
            if (!(cc & 8)) e2.o[3] = -1;
        }
        nxb_corner4(c, 0); nxb_corner4(c, 1); nxb_corner4(c, 2); nxb_corner4(c, 4); nxb_corner4(c, 8);
    } else if (fsum >= 3.0f) {                          // pentachoron at (1,1,1,1)
        int ap = 0xE, bp = 0xD; float as = f[0], bs = f[1];
        if (as <= bs && f[2] < bs) { bs = f[2]; bp = 0xB; } else if (as > bs && f[2] < as) { as = f[2]; ap = 0xB; }
        if (as <= bs && f[3] < bs) { bs = f[3]; bp = 0x7; } else if (as > bs && f[3] < as) { as = f[3]; ap = 0x7; }
        const float u = 4.0f - fsum;
        if (u < as || u < bs) {
            const int cc = (bs < as) ? bp : ap;
            nxb_bits4(e0, cc); nxb_bits4(e1, cc); nxb_bits4(e2, cc);
            e0.mult = e1.mult = e2.mult = 4;
            if (cc & 1) e0.o[0] = 2;
            if (cc & 2) { if (cc & 1) e1.o[1] = 2; else e0.o[1] = 2; }
            if (cc & 4) { if ((cc & 3) != 3) { if ((cc & 3) == 0) e0.o[2] = 2; else e1.o[2] = 2; } else e2.o[2] = 2; }
            if (cc & 8) e2.o[3] = 2;
        } else {
            const int cc = ap & bp;
            nxb_bits4(e0, cc); nxb_bits4(e1, cc); nxb_bits4(e2, cc);
            e0.mult = 2; e1.mult = e2.mult = 3;
            if (cc & 1) e1.o[0] = 2;
            if (cc & 2) { if (cc & 1) e2.o[1] = 2; else e1.o[1] = 2; }
            if (cc & 4) { if ((cc & 3) != 0) e2.o[2] = 2; else e1.o[2] = 2; }
            if (cc & 8) e2.o[3] = 2;
        }
        nxb_corner4(c, 7); nxb_corner4(c, 0xB); nxb_corner4(c, 0xD); nxb_corner4(c, 0xE); nxb_corner4(c, 0xF);
    } else if (fsum <= 2.0f) {                          // first dispentachoron
        bool abig = true, bbig = true; int ap, bp; float as, bs;
        if (f[0] + f[1] > f[2] + f[3]) { as = f[0] + f[1]; ap = 0x3; } else { as = f[2] + f[3]; ap = 0xC; }
        if (f[0] + f[2] > f[1] + f[3]) { bs = f[0] + f[2]; bp = 0x5; } else { bs = f[1] + f[3]; bp = 0xA; }
        if (f[0] + f[3] > f[1] + f[2]) {
            const float sc = f[0] + f[3];
            if (as >= bs && sc > bs) { bs = sc; bp = 0x9; } else if (as < bs && sc > as) { as = sc; ap = 0x9; }
        } else {
            const float sc = f[1] + f[2];
            if (as >= bs && sc > bs) { bs = sc; bp = 0x6; } else if (as < bs && sc > as) { as = sc; ap = 0x6; }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float p = 2.0f - fsum + f[a];
            if (as >= bs && p > bs) { if (a < 3) bs = p; bp = 1 << a; bbig = false; }
            else if (as < bs && p > as) { if (a < 3) as = p; ap = 1 << a; abig = false; }
        }
        if (abig == bbig) {
            if (abig) {
                const int c1 = ap | bp, c2 = ap & bp;
#pragma unroll
                for (int a = 0; a < 4; ++a) { const int on = (c1 >> a) & 1; e0.o[a] = on; e1.o[a] = on ? 1 : -1; }
                e0.mult = 3; e1.mult = 2; e2.mult = 2;
                nxb_set_axis(e2, nxb_first_set(c2), 2);
            } else {
                nxb_pair_minus(ap | bp, e0, e1);        // e2 = (0,0,0,0), mult 0
            }
        } else {
            const int c1 = abig ? ap : bp, c2 = abig ? bp : ap;
            nxb_pair_minus(c1, e0, e1);
            e2.mult = 2;
            nxb_set_axis(e2, nxb_first_set(c2), 2);
        }
        nxb_corner4(c, 1); nxb_corner4(c, 2); nxb_corner4(c, 4); nxb_corner4(c, 8); nxb_corner4(c, 3);
        nxb_corner4(c, 5); nxb_corner4(c, 9); nxb_corner4(c, 6); nxb_corner4(c, 0xA); nxb_corner4(c, 0xC);
    } else {                                            // second dispentachoron
        bool abig = true, bbig = true; int ap, bp; float as, bs;
        if (f[0] + f[1] < f[2] + f[3]) { as = f[0] + f[1]; ap = 0xC; } else { as = f[2] + f[3]; ap = 0x3; }
        if (f[0] + f[2] < f[1] + f[3]) { bs = f[0] + f[2]; bp = 0xA; } else { bs = f[1] + f[3]; bp = 0x5; }
        if (f[0] + f[3] < f[1] + f[2]) {
            const float sc = f[0] + f[3];
            if (as <= bs && sc < bs) { bs = sc; bp = 0x6; } else if (as > bs && sc < as) { as = sc; ap = 0x6; }
        } else {
            const float sc = f[1] + f[2];
            if (as <= bs && sc < bs) { bs = sc; bp = 0x9; } else if (as > bs && sc < as) { as = sc; ap = 0x9; }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float p = 3.0f - fsum + f[a];
            const int code = 0xF & ~(1 << a);
            if (as <= bs && p < bs) { if (a < 3) bs = p; bp = code; bbig = false; }
            else if (as > bs && p < as) { if (a < 3) as = p; ap = code; abig = false; }
        }
        if (abig == bbig) {
            if (abig) {
                const int c1 = ap & bp, c2 = ap | bp;
                const int ax = nxb_first_set(c1);
                e0.mult = 1; nxb_set_axis(e0, ax, 1);
                e1.mult = 2; nxb_set_axis(e1, ax, 2);
                nxb_bits4(e2, 0xF); e2.mult = 2;
                nxb_set_axis(e2, nxb_first_clear(c2), -1);
            } else {
                nxb_bits4(e2, 0xF); e2.mult = 4;
                nxb_pair_plus(ap & bp, e0, e1);
            }
        } else {
            const int c1 = abig ? ap : bp, c2 = abig ? bp : ap;
            nxb_pair_plus(c1, e0, e1);
            nxb_bits4(e2, 0xF); e2.mult = 2;
            nxb_set_axis(e2, nxb_first_clear(c2), -1);
        }
        nxb_corner4(c, 7); nxb_corner4(c, 0xB); nxb_corner4(c, 0xD); nxb_corner4(c, 0xE); nxb_corner4(c, 3);
        nxb_corner4(c, 5); nxb_corner4(c, 9); nxb_corner4(c, 6); nxb_corner4(c, 0xA); nxb_corner4(c, 0xC);
    }
    nxb_add4(c, e0);
    nxb_add4(c, e1);
    nxb_add4(c, e2);
    return c.v * (1.0f / 30.0f);
}

// Branch-free OpenSimplex 3-D evaluation for the fused fBm kernel (sm_100a).
//
// Same candidate set as the reference (opensimplex.py:266-759) -- the legacy algorithm is not a
// full lattice sum, so the SAME comparisons must pick the SAME lattice points -- but organised so
// that every lane of a warp executes the same instruction stream whatever region of the lattice
// cell it falls in (profiles/r01_ncu_summary.json: the branchy version ran at 16.4 of 32 active
// lanes and 706 warp-instructions per evaluation):
//
//   * the 8 corners of the lattice cube are ALWAYS evaluated; a per-lane constant T (2 or -16)
//     replaces the "2" in attn = 2 - |d|^2, so a corner the reference would not visit gets a
//     negative attn and contributes exactly 0.  Which corners are live: the region's simplex
//     corners plus, when an "extra" lattice point of the reference happens to be a cube corner,
//     that corner.
//   * at most two extras are NOT cube corners (offsets containing -1 or 2).  They come from a
//     19-entry table of displacement constants / hash offsets indexed by a small per-lane code.
//   * only the ~25-instruction selection logic branches on the region (three short warp-level
//     branches; coherent warps execute one of them).
//   * hash and gradient tables live in shared memory in a bank-conflict-free layout: entry e of
//     lane l sits at word e*32 + l, so every lookup is one wavefront whatever the indices are.
//     Values of the permutation table are stored pre-multiplied by the row pitch (128 B), the
//     last hash level indexes three float tables holding the gradient components directly
//     (GRADIENTS_3D[perm_grad_index_3D[h]], opensimplex.py:123-130): no decode arithmetic.
//   * round 2: lattice cell, in-cell coordinates and the whole candidate selection are float64
//     (nxf_noise3_x103_d): same comparisons on the reference's own quantities, so the candidate
//     set is the reference's bit for bit; they issue on the otherwise idle FP64 pipe.  floor()
//     uses the 1.5*2^52 magic-number trick (exact for |x| < 2^51), no F2I/I2F.
//
// This header is plain C++ apart from a few intrinsics so that tools/host_noise_check.cpp can
// compile it with g++ and compare against the float64 oracle without a GPU.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#include <math.h>
#include <string.h>
#define NXF_DEV inline
static inline float nxf_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t nxf_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#define nxf_fma(a, b, c) fmaf(a, b, c)
#else
#define NXF_DEV __device__ __forceinline__
#define nxf_as_float(u) __uint_as_float(u)
#define nxf_as_uint(f) __float_as_uint(f)
#define nxf_fma(a, b, c) fmaf(a, b, c)
#endif

// ---- packed pairs of floats: sm_100a has two-wide FP32 instructions (FFMA2 / FADD2 / FMUL2, one issue slot
// for two results).  The kernel is issue-bound, so the ten contributions of an evaluation are computed as
// five PAIRS.  Host build: plain scalar code with the same roundings.
#ifdef __CUDACC__
typedef float2 nxf_f2;
#define nxf_fma2(a, b, c) __ffma2_rn(a, b, c)
#define nxf_add2(a, b) __fadd2_rn(a, b)
#define nxf_mul2(a, b) __fmul2_rn(a, b)
#else
struct nxf_f2 { float x, y; };
static inline nxf_f2 nxf_fma2(nxf_f2 a, nxf_f2 b, nxf_f2 c) { nxf_f2 r = {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; return r; }
static inline nxf_f2 nxf_add2(nxf_f2 a, nxf_f2 b) { nxf_f2 r = {a.x + b.x, a.y + b.y}; return r; }
static inline nxf_f2 nxf_mul2(nxf_f2 a, nxf_f2 b) { nxf_f2 r = {a.x * b.x, a.y * b.y}; return r; }
#endif
NXF_DEV nxf_f2 nxf_mk2(float x, float y) { nxf_f2 r; r.x = x; r.y = y; return r; }

#define NXF_ROW 128u                    // bytes per table row (32 lanes x 4 B)
#define NXF_MASK 0x7F80u                // (e & 255) * 128
#define NXF_TAB_BYTES (256u * NXF_ROW)  // one replicated 256-entry table: 32 KB
#define NXF_OFF_P 0u
#define NXF_OFF_GX (1u * NXF_TAB_BYTES)
#define NXF_OFF_GY (2u * NXF_TAB_BYTES)
#define NXF_OFF_GZ (3u * NXF_TAB_BYTES)
#define NXF_OFF_EXT (4u * NXF_TAB_BYTES)        // 19 extra records x 32 B (not replicated)
#define NXF_SMEM_BYTES (NXF_OFF_EXT + 32u * 20u)
#define NXF_MAX_COORD 1.0e9              // |lattice coordinate| bound (the hash keeps the low 32 bits of the integer part)

// record of one non-cube extra lattice point: displacement constants (offset + m/3), the live
// constant T (2, or -16 for "no extra"), hash offsets pre-multiplied by the row pitch
struct NxfExtra { float ncx, ncy, ncz, nT; int32_t ox, oy, oz, pad; };   // -(offset + m/3), -T

// offsets of the 18 non-cube extras, in the order the selection logic indexes them
//  0..5   tet(0,0,0), (0,0,0) among the two closest: axis k -> 2k, 2k+1     (opensimplex.py:321-350)
//  6..11  tet(1,1,1), (1,1,1) among the two closest: pair p -> 6+2p, 7+2p   (opensimplex.py:436-466)
//  12..14 (1,1,1) - 2 e_k                                                   (m = 1)
//  15..17 2 e_k                                                             (m = 2)
//  18     none
#ifdef __CUDACC__
__device__
#endif
static const int8_t NXF_EXTRA_OFFSETS[19][3] = {
    {1, -1, 0}, {1, 0, -1}, {-1, 1, 0}, {0, 1, -1}, {-1, 0, 1}, {0, -1, 1},
    {2, 1, 0}, {1, 2, 0}, {2, 0, 1}, {1, 0, 2}, {0, 2, 1}, {0, 1, 2},
    {-1, 1, 1}, {1, -1, 1}, {1, 1, -1},
    {2, 0, 0}, {0, 2, 0}, {0, 0, 2},
    {0, 0, 0}};

struct NxfCtx {
    const char *sm;         // shared-memory base of the tables (host build: plain pointer)
    uint32_t lane4;         // lane * 4; on the device OR-ed with the 32 KB-aligned shared address of
                            // the tables, so that (hash & MASK) | lane4 IS the load address
    uint32_t xb7, yb7, zb7; // (lattice base & 255) * 128, unmasked
    float dx0, dy0, dz0;
    float v;
    nxf_f2 v2;              // packed accumulator of the five contribution pairs
};

#ifdef __CUDACC__
// idx carries the absolute shared address (tables are 32 KB aligned): one LOP3 + one LDS per lookup,
// table selected by the immediate offset.
template <uint32_t OFF>
__device__ __forceinline__ uint32_t nxf_ldu_t(uint32_t idx)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(idx), "n"(OFF));
    return v;
}
template <uint32_t OFF>
__device__ __forceinline__ float nxf_ldf_t(uint32_t idx)
{
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(idx), "n"(OFF));
    return v;
}
#define nxf_ldu(c, off, idx) nxf_ldu_t<off>(idx)
#define nxf_ldf(c, off, idx) nxf_ldf_t<off>(idx)
#else
NXF_DEV uint32_t nxf_ldu(const NxfCtx &c, uint32_t off, uint32_t idx)
{
    return *reinterpret_cast<const uint32_t *>(c.sm + off + idx);
}
NXF_DEV float nxf_ldf(const NxfCtx &c, uint32_t off, uint32_t idx)
{
    return *reinterpret_cast<const float *>(c.sm + off + idx);
}
#endif
// one hash level: row ((h + add) & 255) of the P table, this lane's copy
NXF_DEV uint32_t nxf_hash(const NxfCtx &c, uint32_t h_plus_add)
{
    return nxf_ldu(c, NXF_OFF_P, (h_plus_add & NXF_MASK) | c.lane4);
}

// contributions of TWO lattice points at once: last-level table indices idxA / idxB (bytes), displacement
// pairs d = {d_A, d_B} per axis, nT = {-T_A, -T_B} (T = 2 for a live point, -16 for a dead one).
//   acc = |d|^2 - T = -attn (same roundings as T - |d|^2, sign flipped); m = min(acc, 0) = -max(attn, 0);
//   contribution = m^4 * (g . d)
NXF_DEV void nxf_contrib2(NxfCtx &c, uint32_t idxA, uint32_t idxB, nxf_f2 dx, nxf_f2 dy, nxf_f2 dz, nxf_f2 nT)
{
    nxf_f2 acc = nxf_fma2(dz, dz, nxf_fma2(dy, dy, nxf_fma2(dx, dx, nT)));
    const nxf_f2 gx = nxf_mk2(nxf_ldf(c, NXF_OFF_GX, idxA), nxf_ldf(c, NXF_OFF_GX, idxB));
    const nxf_f2 gy = nxf_mk2(nxf_ldf(c, NXF_OFF_GY, idxA), nxf_ldf(c, NXF_OFF_GY, idxB));
    const nxf_f2 gz = nxf_mk2(nxf_ldf(c, NXF_OFF_GZ, idxA), nxf_ldf(c, NXF_OFF_GZ, idxB));
    const nxf_f2 dot = nxf_fma2(gz, dz, nxf_fma2(gy, dy, nxf_mul2(gx, dx)));
    nxf_f2 m = nxf_mk2(fminf(acc.x, 0.0f), fminf(acc.y, 0.0f));
    m = nxf_mul2(m, m);
    c.v2 = nxf_fma2(nxf_mul2(m, m), dot, c.v2);
}

template <int IA, int JA, int KA, int IB, int JB, int KB>
NXF_DEV void nxf_corner2(NxfCtx &c, nxf_f2 dx0, nxf_f2 dy0, nxf_f2 dz0, uint32_t hA, uint32_t hB, float TA, float TB)
{
    constexpr float mA = (float)(IA + JA + KA) * (1.0f / 3.0f), mB = (float)(IB + JB + KB) * (1.0f / 3.0f);
    const uint32_t idxA = ((hA + c.zb7 + (uint32_t)KA * NXF_ROW) & NXF_MASK) | c.lane4;
    const uint32_t idxB = ((hB + c.zb7 + (uint32_t)KB * NXF_ROW) & NXF_MASK) | c.lane4;
    nxf_contrib2(c, idxA, idxB,
                 nxf_add2(dx0, nxf_mk2(-((float)IA + mA), -((float)IB + mB))),
                 nxf_add2(dy0, nxf_mk2(-((float)JA + mA), -((float)JB + mB))),
                 nxf_add2(dz0, nxf_mk2(-((float)KA + mA), -((float)KB + mB))), nxf_mk2(-TA, -TB));
}

NXF_DEV void nxf_extra2(NxfCtx &c, nxf_f2 dx0, nxf_f2 dy0, nxf_f2 dz0, int e0, int e1)
{
    const NxfExtra &r0 = *reinterpret_cast<const NxfExtra *>(c.sm + NXF_OFF_EXT + (uint32_t)e0 * 32u);
    const NxfExtra &r1 = *reinterpret_cast<const NxfExtra *>(c.sm + NXF_OFF_EXT + (uint32_t)e1 * 32u);
    uint32_t h0 = nxf_hash(c, c.xb7 + (uint32_t)r0.ox), h1 = nxf_hash(c, c.xb7 + (uint32_t)r1.ox);
    h0 = nxf_hash(c, h0 + c.yb7 + (uint32_t)r0.oy); h1 = nxf_hash(c, h1 + c.yb7 + (uint32_t)r1.oy);
    const uint32_t idx0 = ((h0 + c.zb7 + (uint32_t)r0.oz) & NXF_MASK) | c.lane4;
    const uint32_t idx1 = ((h1 + c.zb7 + (uint32_t)r1.oz) & NXF_MASK) | c.lane4;
    nxf_contrib2(c, idx0, idx1, nxf_add2(dx0, nxf_mk2(r0.ncx, r1.ncx)), nxf_add2(dy0, nxf_mk2(r0.ncy, r1.ncy)),
                 nxf_add2(dz0, nxf_mk2(r0.ncz, r1.ncz)), nxf_mk2(r0.nT, r1.nT));
}

// opensimplex.py:306-312 / 421-427 / 584-599: keep the two best of three candidates.
// Starts with a = cand0, b = cand1; cand2 replaces the worse of them if it beats it.
// "better" = larger score.  Tie rules are the reference's (>= on a vs b, strict on the newcomer).
NXF_DEV void nxf_pick2(float s0, float s1, float s2, int c0, int c1, int c2, int &ap, float &as, int &bp, float &bs)
{
    ap = c0; as = s0; bp = c1; bs = s1;
    if (as >= bs && s2 > bs) { bs = s2; bp = c2; }
    else if (as < bs && s2 > as) { as = s2; ap = c2; }
}

// Candidate selection, literal form (one branch per lattice region, as the reference is written).
// Kept as the specification the branch-free nxf_select below is checked against
// (tools/host_noise_check.cpp compares them exactly, ties included).
NXF_DEV void nxf_select_branchy(float fx, float fy, float fz, float fsum,
                                float &T000, float &T100, float &T010, float &T001,
                                float &T110, float &T101, float &T011, float &T111, int &e0, int &e1)
{
    const float LIVE = 2.0f, DEAD = -16.0f;
    e1 = 18;
    if (fsum <= 1.0f) {                                    // tetrahedron at (0,0,0)
        int ap, bp; float as, bs;
        nxf_pick2(fx, fy, fz, 1, 2, 4, ap, as, bp, bs);
        const float w = 1.0f - fsum;
        T000 = T100 = T010 = T001 = LIVE; T111 = DEAD;
        if (w > as || w > bs) {
            const int cc = (bs > as) ? bp : ap;            // 1, 2 or 4
            e0 = (cc >> 1) * 2; e1 = e0 + 1;
            T110 = T101 = T011 = DEAD;
        } else {
            const int cc = ap | bp;                        // 3, 5 or 6: that cube corner is the first extra
            e0 = 12 + ((6 - cc + 1) >> 1);                 // (1,1,1) - 2 e_k, k = the axis not in cc... see table
            T110 = cc == 3 ? LIVE : DEAD; T101 = cc == 5 ? LIVE : DEAD; T011 = cc == 6 ? LIVE : DEAD;
        }
    } else if (fsum >= 2.0f) {                             // tetrahedron at (1,1,1)
        int ap, bp; float as, bs;
        nxf_pick2(-fx, -fy, -fz, 6, 5, 3, ap, as, bp, bs); // two SMALLEST of fx,fy,fz
        const float w = fsum - 3.0f;                       // -(3 - fsum), same sign convention as the scores
        T110 = T101 = T011 = T111 = LIVE; T000 = DEAD;
        if (w > as || w > bs) {
            const int cc = (bs > as) ? bp : ap;            // 3, 5 or 6
            e0 = 6 + (cc >> 1 == 1 ? 0 : (cc == 5 ? 2 : 4)); e1 = e0 + 1;
            T100 = T010 = T001 = DEAD;
        } else {
            const int cc = ap & bp;                        // 1, 2 or 4
            e0 = 15 + (cc >> 1);
            T100 = cc == 1 ? LIVE : DEAD; T010 = cc == 2 ? LIVE : DEAD; T001 = cc == 4 ? LIVE : DEAD;
        }
    } else {                                               // octahedron
        const float p1 = fx + fy, p2 = fx + fz, p3 = fy + fz;
        const bool f1 = p1 > 1.0f, f2 = p2 > 1.0f, f3 = p3 > 1.0f;
        float as = f1 ? p1 - 1.0f : 1.0f - p1;  int ap = f1 ? 3 : 4;  bool afar = f1;
        float bs = f2 ? p2 - 1.0f : 1.0f - p2;  int bp = f2 ? 5 : 2;  bool bfar = f2;
        const float sc = f3 ? p3 - 1.0f : 1.0f - p3;  const int cp = f3 ? 6 : 1;
        if (as <= bs && as < sc) { ap = cp; afar = f3; }
        else if (as > bs && bs < sc) { bp = cp; bfar = f3; }
        T100 = T010 = T001 = T110 = T101 = T011 = LIVE;
        T000 = T111 = DEAD;
        if (afar == bfar) {
            if (afar) {                                    // (1,1,1) + 2 e_k on the shared axis
                const int cc = ap & bp;
                T111 = LIVE;
                e0 = 15 + ((cc & 1) ? 0 : ((cc & 2) ? 1 : 2));
            } else {                                       // (0,0,0) + (1,1,1) - 2 e_k on the omitted axis
                const int cc = ap | bp;
                T000 = LIVE;
                e0 = 12 + (!(cc & 1) ? 0 : (!(cc & 2) ? 1 : 2));
            }
        } else {
            const int c1 = afar ? ap : bp, c2 = afar ? bp : ap;
            e0 = 12 + (!(c1 & 1) ? 0 : (!(c1 & 2) ? 1 : 2));
            e1 = 15 + ((c2 & 1) ? 0 : ((c2 & 2) ? 1 : 2));
        }
    }
}

NXF_DEV float nxf_abs(float v) { return fabsf(v); }
NXF_DEV double nxf_abs(double v) { return fabs(v); }

#ifdef __CUDACC__
#define NXF_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define NXF_ANY(p) (p)
#endif

// Candidate selection, branch-free form.  Same decisions as nxf_select_branchy:
//   * the two tetrahedra share one routine: the (1,1,1) tetrahedron is the (0,0,0) one with the
//     scores negated (exact in floating point), candidate i <-> axis i in both;
//   * in every region the work is "keep the two best of three candidates" (nxf_pick2 semantics)
//     followed by a two-way classification; candidates are tracked as axis numbers 0..2 and the
//     extra codes / live corners are computed arithmetically from them;
//   * the two blocks are skipped with a warp vote when no lane of the warp needs them, so a
//     coherent warp pays for one block, a mixed warp for both, and no lane ever idles.
//
// R = float: the FP32 selection of round 1.  R = double: the SAME comparisons on the reference's
// own float64 quantities -- the candidate set is then the reference's, bit for bit, ties included;
// only predicates leave this routine, so no 64-bit value is ever selected or stored.
template <typename R>
NXF_DEV void nxf_select_r(R fx, R fy, R fz, R fsum,
                          float &T000, float &T100, float &T010, float &T001,
                          float &T110, float &T101, float &T011, float &T111, int &e0, int &e1)
{
    const float LIVE = 2.0f, DEAD = -16.0f;
    const bool r0 = fsum <= (R)1, r1 = fsum >= (R)2, tet = r0 || r1;
    // outputs of the tetrahedron block
    int te0 = 18, te1 = 18; bool opt0 = false, opt1 = false, opt2 = false;
    if (NXF_ANY(tet)) {
        // tet1 = tet0 with negated scores (exact); 1 - fsum == -(fsum - 1) and 3 - fsum == -(fsum - 3) exactly
        const R s0 = r1 ? -fx : fx, s1 = r1 ? -fy : fy, s2 = r1 ? -fz : fz;
        const R wm = fsum - (r1 ? (R)3 : (R)1);
        const R w = r1 ? wm : -wm;
        const bool ge = s0 >= s1, g21 = s2 > s1, g20 = s2 > s0, g12 = s1 > s2;
        const bool w0 = w > s0, w1 = w > s1, w2 = w > s2;
        const bool c1 = ge && g21;                      // newcomer replaces b
        const bool c2 = !ge && g20;                     // newcomer replaces a
        const int ia = c2 ? 2 : 0, ib = c1 ? 2 : 1;     // as = s[ia], bs = s[ib]
        const bool caseA = (c2 ? w2 : w0) || (c1 ? w2 : w1);
        const bool bgta = c1 ? g20 : (c2 ? g12 : !ge);  // bs > as
        const int cs = bgta ? ib : ia;                  // the single closest candidate (case A)
        const int io = 3 - ia - ib;                     // the axis not among the two closest (case B)
        // tet0: A -> extras 2cs, 2cs+1;            B -> cube corner omitting io + (1,1,1)-2e_io (12+io)
        // tet1: A -> extras 10-2cs, 11-2cs;        B -> cube corner e_io         + 2e_io       (15+io)
        const int a0 = r1 ? 10 - 2 * cs : 2 * cs;
        const int b0 = (r1 ? 15 : 12) + io;
        te0 = caseA ? a0 : b0;
        te1 = caseA ? a0 + 1 : 18;
        opt0 = !caseA && io == 0; opt1 = !caseA && io == 1; opt2 = !caseA && io == 2;
    }
    // outputs of the octahedron block
    int oe0 = 18, oe1 = 18; bool near2 = false, far2 = false;
    if (NXF_ANY(!tet)) {
        const R p1 = fx + fy, p2 = fx + fz, p3 = fy + fz;
        const bool f1 = p1 > (R)1, f2 = p2 > (R)1, f3 = p3 > (R)1;
        // |p - 1| is the reference's `p - 1` (p > 1) or `1 - p` (else): the same rounded value
        const R d1 = p1 - (R)1, d2 = p2 - (R)1, d3 = p3 - (R)1;
        const R q1 = nxf_abs(d1), q2 = nxf_abs(d2), q3 = nxf_abs(d3);
        // reference: a <- p1, b <- p2; p3 replaces a if (as <= bs && as < sc), else b if (as > bs && bs < sc)
        const bool le = q1 <= q2;
        const bool ra = le && q1 < q3;
        const bool rb = !le && q2 < q3;
        const bool fa = ra ? f3 : f1, fb = rb ? f3 : f2;
        const int ka = ra ? 0 : 2, kb = rb ? 0 : 1;     // candidate p1 <-> axis z, p2 <-> y, p3 <-> x
        const int k3 = 3 - ka - kb;
        const bool same = fa == fb;
        const int kf = fa ? ka : kb, kn = fa ? kb : ka;
        oe0 = same ? (fa ? 15 : 12) + k3 : 12 + kf;
        oe1 = same ? 18 : 15 + kn;
        near2 = same && !fa; far2 = same && fa;
    }
    e0 = tet ? te0 : oe0;
    e1 = tet ? te1 : oe1;
    T000 = (r0 || (!tet && near2)) ? LIVE : DEAD;
    T111 = (r1 || (!tet && far2)) ? LIVE : DEAD;
    // m = 1 corners (axis k): live in tet0 and the octahedron, in tet1 only as the case-B extra
    T100 = (!r1 || opt0) ? LIVE : DEAD;
    T010 = (!r1 || opt1) ? LIVE : DEAD;
    T001 = (!r1 || opt2) ? LIVE : DEAD;
    // m = 2 corners (omitting axis k): live in tet1 and the octahedron, in tet0 only as the case-B extra
    T011 = (!r0 || opt0) ? LIVE : DEAD;
    T101 = (!r0 || opt1) ? LIVE : DEAD;
    T110 = (!r0 || opt2) ? LIVE : DEAD;
}

NXF_DEV void nxf_select(float fx, float fy, float fz, float fsum,
                        float &T000, float &T100, float &T010, float &T001,
                        float &T110, float &T101, float &T011, float &T111, int &e0, int &e1)
{
    nxf_select_r<float>(fx, fy, fz, fsum, T000, T100, T010, T001, T110, T101, T011, T111, e0, e1);
}

// the 8 cube corners: shared hash tree (2 + 4 lookups), one leaf each, evaluated as four PAIRS; the two
// non-cube extras are the fifth pair
NXF_DEV float nxf_contributions(NxfCtx &c, float T000, float T100, float T010, float T001,
                                float T110, float T101, float T011, float T111, int e0, int e1)
{
    const uint32_t hx0 = nxf_hash(c, c.xb7), hx1 = nxf_hash(c, c.xb7 + NXF_ROW);
    const uint32_t h00 = nxf_hash(c, hx0 + c.yb7), h01 = nxf_hash(c, hx0 + c.yb7 + NXF_ROW);
    const uint32_t h10 = nxf_hash(c, hx1 + c.yb7), h11 = nxf_hash(c, hx1 + c.yb7 + NXF_ROW);
    const nxf_f2 dx0 = nxf_mk2(c.dx0, c.dx0), dy0 = nxf_mk2(c.dy0, c.dy0), dz0 = nxf_mk2(c.dz0, c.dz0);
    c.v2 = nxf_mk2(0.0f, 0.0f);
    nxf_corner2<0, 0, 0, 1, 1, 1>(c, dx0, dy0, dz0, h00, h11, T000, T111);
    nxf_corner2<1, 0, 0, 0, 1, 1>(c, dx0, dy0, dz0, h10, h01, T100, T011);
    nxf_corner2<0, 1, 0, 1, 0, 1>(c, dx0, dy0, dz0, h01, h10, T010, T101);
    nxf_corner2<0, 0, 1, 1, 1, 0>(c, dx0, dy0, dz0, h00, h11, T001, T110);
    nxf_extra2(c, dx0, dy0, dz0, e0, e1);
    return c.v2.x + c.v2.y;
}

// ---- float64 prologue + selection (device: __dadd_rn / __dmul_rn are never contracted into FMAs;
// host build of this header: plain double arithmetic, compile with -ffp-contract=off) -----------
#ifdef __CUDACC__
#define nxf_dadd(a, b) __dadd_rn(a, b)
#define nxf_dmul(a, b) __dmul_rn(a, b)
#define nxf_dlo(d) ((uint32_t)__double2loint(d))
#else
#define nxf_dadd(a, b) ((a) + (b))
#define nxf_dmul(a, b) ((a) * (b))
static inline uint32_t nxf_dlo(double d) { uint64_t u; memcpy(&u, &d, 8); return (uint32_t)u; }
#endif

// x,y,z: the reference's float64 noise3d arguments (opensimplex.py:266).  Lattice cell, in-cell
// coordinates and every comparison of the candidate selection are evaluated in float64, operation
// for operation as the reference does (opensimplex.py:270-300 and the region tests) -- they run on
// the FP64 pipe, which this kernel leaves idle otherwise -- so the candidate set is the reference's
// own, exact ties included.  The contributions (displacements, attn^4 * gradient dot) are FP32.
// Returns noise3d * 103.
NXF_DEV float nxf_noise3_x103_d(double x, double y, double z, const char *sm, uint32_t lane4)
{
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: (x + MAGIC) - MAGIC rounds to nearest integer
    const double so = nxf_dmul(nxf_dadd(nxf_dadd(x, y), z), -1.0 / 6);          // opensimplex.py:270 stretch_offset
    const double xs = nxf_dadd(x, so), ys = nxf_dadd(y, so), zs = nxf_dadd(z, so);
    const double tx = nxf_dadd(xs, MAGIC), ty = nxf_dadd(ys, MAGIC), tz = nxf_dadd(zs, MAGIC);
    double bx = nxf_dadd(tx, -MAGIC), by = nxf_dadd(ty, -MAGIC), bz = nxf_dadd(tz, -MAGIC);
    // floor = round-to-nearest, minus 1 where that rounded up (fastfloor, opensimplex.py:18-21; exact for |x| < 2^51)
    const bool ux = bx > xs, uy = by > ys, uz = bz > zs;
    bx = nxf_dadd(bx, ux ? -1.0 : 0.0); by = nxf_dadd(by, uy ? -1.0 : 0.0); bz = nxf_dadd(bz, uz ? -1.0 : 0.0);
    NxfCtx c;
    c.sm = sm; c.lane4 = lane4;
    // low mantissa bits of (xs + MAGIC) are the integer; only bits 0..7 survive the masks
    c.xb7 = (nxf_dlo(tx) - (ux ? 1u : 0u)) << 7;
    c.yb7 = (nxf_dlo(ty) - (uy ? 1u : 0u)) << 7;
    c.zb7 = (nxf_dlo(tz) - (uz ? 1u : 0u)) << 7;
    const double fx = nxf_dadd(xs, -bx), fy = nxf_dadd(ys, -by), fz = nxf_dadd(zs, -bz);   // :285-287 xins..
    const double fsum = nxf_dadd(nxf_dadd(fx, fy), fz);                           // :290 in_sum
    // position relative to the cell origin, dx0 = x - xb (opensimplex.py:279-295): in exact arithmetic
    // x - (xsb + (xsb+ysb+zsb)/3) = xins + in_sum/3 (the un-skew of the in-cell coordinates), which
    // needs no large-number cancellation and is taken in FP32 -- it only feeds the FP32 contributions
    const float fsum_f = (float)fsum;
    c.dx0 = nxf_fma(fsum_f, 1.0f / 3.0f, (float)fx);
    c.dy0 = nxf_fma(fsum_f, 1.0f / 3.0f, (float)fy);
    c.dz0 = nxf_fma(fsum_f, 1.0f / 3.0f, (float)fz);
    c.v = 0.0f;
    float T000, T100, T010, T001, T110, T101, T011, T111;
    int e0, e1;
    nxf_select_r<double>(fx, fy, fz, fsum, T000, T100, T010, T001, T110, T101, T011, T111, e0, e1);
    return nxf_contributions(c, T000, T100, T010, T001, T110, T101, T011, T111, e0, e1);
}

// Fill the replicated tables.  perm8 / grad8 as in NxbTables (nxb_noise.cuh).  Called by every
// thread of the CTA (tid, nthreads); on the host with (0, 1).
NXF_DEV void nxf_build_tables(const uint8_t *perm8, const uint8_t *grad8, char *sm, int tid, int nthreads)
{
    uint32_t *P = reinterpret_cast<uint32_t *>(sm + NXF_OFF_P);
    float *GX = reinterpret_cast<float *>(sm + NXF_OFF_GX);
    float *GY = reinterpret_cast<float *>(sm + NXF_OFF_GY);
    float *GZ = reinterpret_cast<float *>(sm + NXF_OFF_GZ);
    for (int w = tid; w < 256 * 32; w += nthreads) {
        const int e = w >> 5;
        const uint32_t g = grad8[e], ax = g >> 3;
        P[w] = (uint32_t)perm8[e] << 7;
        GX[w] = ((g & 1) ? -1.0f : 1.0f) * (ax == 0 ? 11.0f : 4.0f);
        GY[w] = ((g & 2) ? -1.0f : 1.0f) * (ax == 1 ? 11.0f : 4.0f);
        GZ[w] = ((g & 4) ? -1.0f : 1.0f) * (ax == 2 ? 11.0f : 4.0f);
    }
    NxfExtra *rec = reinterpret_cast<NxfExtra *>(sm + NXF_OFF_EXT);
    for (int e = tid; e < 19; e += nthreads) {
        const int8_t *o = NXF_EXTRA_OFFSETS[e];
        const float cc = (float)(o[0] + o[1] + o[2]) * (1.0f / 3.0f);
        rec[e].ncx = -((float)o[0] + cc); rec[e].ncy = -((float)o[1] + cc); rec[e].ncz = -((float)o[2] + cc);
        rec[e].nT = e == 18 ? 16.0f : -2.0f;
        rec[e].ox = o[0] * (int)NXF_ROW; rec[e].oy = o[1] * (int)NXF_ROW; rec[e].oz = o[2] * (int)NXF_ROW;
        rec[e].pad = 0;
    }
}

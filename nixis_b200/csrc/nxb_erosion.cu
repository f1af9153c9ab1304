// Erosion sweeps -- erosion.py:34-40, 76-99, 197-279 -- as a shared-memory-staged stencil.
//
// No single unit is saturated (v11 at d=2500: HBM 78 % of the measured copy peak, issue slots 66 %, L2 40 %); the sweep
// is bound by latency times concurrency -- see DESIGN.md section 4.2 for the five variants that established it.
// Algorithmic traffic per vertex-iteration of erosion_iteration3:
//   own h,w,s read 12 B + write 12 B + adjacency row 24 B + own position 12 B = 60 B.
// What this implementation streams from HBM per vertex-iteration (nxb_erosion_plan.cuh):
//   kind 3 (affine tile, one length per edge)  h,w,s read 12 + write 12 + dist3 12       = 36 B
//   kind 4 (two affine pieces + 4 exceptions)  as kind 3 but for the exception vertices   = 36 B
//   kind 2 (affine tile, full length rows)     ... + 6 edge lengths 24                   = 48 B
//   kind 1 (explicit codes)                    ... + 24 + 16-bit tile-local adjacency 12 = 60 B
//
// History (profiles/r01_ncu_summary.json): v1, one thread per vertex with 18 global gathers
// (xyz, h, w of 6 neighbours), ran at 45 % of HBM peak with minimal DRAM traffic -- latency bound;
// v2 streamed the tile's own data with cp.async.bulk but kept the gathers and was then bound by
// gather latency + instruction issue (322 instr/vertex, six IEEE sqrt sequences).  v3 removed both:
//
//   * EDGE LENGTHS ARE PRECOMPUTED once, in FP64, and stored as FP32 (erosion.py:227-229 uses the
//     undisplaced sphere positions, so they never change).  No neighbour positions are read and no
//     sqrt is evaluated in the sweep; it also removes the cancellation error of differencing FP32
//     positions (relative 1e-4 at d=2500).
//   * TILE PLAN (nxb_erosion_plan.cuh): for each tile of 256 consecutive vertices the neighbours
//     are the tile itself plus <= 8 contiguous index runs.  A producer warp brings the tile's own
//     streams AND those runs of h / w into shared memory with cp.async.bulk (TMA bulk copy, SASS
//     UBLKCP) completing on an mbarrier, n_stages tiles ahead; eight consumer warps then read every
//     neighbour value from shared memory.  No global gathers, no L1 dependence.  The halo runs
//     were just streamed by a neighbouring tile, so they come from L2, not HBM.
//   * v5 IMPLICIT ADJACENCY (kind 2): no adjacency codes on tiles whose six neighbour distances
//     are the same for all 256 vertices: 0.59 -> 0.51-0.56 ms per sweep at d=2500.
//   * v6 (round 2) ONE LENGTH PER EDGE ON AFFINE TILES (kind 3): the round-1 attempt decoded the
//     owner row per vertex from spare code bits (~50 instructions per vertex, slower than the
//     bytes it saved); on affine tiles the owner row and entry are per-tile constants, so the
//     lookup costs nothing.  The explicit-code dist3 variant, the in-kernel flag wait and the
//     "boundary tiles first" order of round 1 (all measured slower) are gone
//     from the hot kernel.
//   * v7 HEIGHT AND WATER INTERLEAVED ({h, w} float2 per vertex): a neighbour's two values are always
//     read together -- one bulk copy per run instead of two, one 8-byte shared load per neighbour,
//     one 8-byte peer store per boundary vertex.
//   * v8 THE VERTEX'S OWN DATA BYPASSES THE BULK-COPY ENGINE.  ncu on v6 (profiles/r02_*): DRAM 59 %,
//     issue 58 %, and 55 % of all warp stall samples on the consumers' full-barrier wait -- although the
//     producer never waits for a free stage, pipeline depth 2 runs as fast as depth 3, and an L2
//     prefetch of the next tiles through cp.async.bulk.prefetch makes the sweep 28 % SLOWER: the time
//     follows the bytes queued on the per-SM bulk-copy (TMA) engine, ~21 B/clk/SM, whatever their
//     source.  So only what NEIGHBOURS share is staged (the {h, w} window and runs); the vertex's own
//     sediment, its six edge lengths (kind 3: dist3[3 v + D_q]; kinds 1 / 2: its row of the full table)
//     and, on kind-1 tiles, its 16-bit neighbour codes are plain coalesced loads: 6.3 KB per tile through
//     the engine instead of 13.7 KB, and a pipeline stage shrinks from 18 KB to 7.4 KB.
//   * v9 / v10 THOSE LOADS RUN ONE TILE AHEAD: the producer puts the next tile's descriptor words into the
//     stage header and every consumer issues the next tile's loads behind the math of the current one, into
//     the registers it reads an iteration later (one prefetch scoreboard: see the comments at the loads);
//     32-bit indices, water stored already rained, slopes taken inside the tile-kind branches:
//     0.545-0.576 -> 0.458-0.475 ms per sweep at d=2500.
//   * v11 TWO-PIECE TILES (kind 4): a tile with the end of a mesh row inside is swept as kind 3 with two sets
//     of constants and four exception vertices: 98 % of the tiles read no adjacency and one length per edge,
//     2.62 -> 2.35 GB per sweep, same time.
//   * ping-pong buffers replace the reference's three np.copy + copy-back pass (erosion.py:199-201,
//     274-277); `water += rain` (erosion.py:182-183) is fused into the reads.
//   * the sweep LOOP lives here (nxb_erode3_run_*): n sweeps are n launches issued from C with
//     programmatic dependent launch, so the next sweep's CTAs are resident with barriers
//     initialised and the first descriptor fetched by the time the previous sweep drains.
#include "nxb_common.cuh"
#include "nxb_erosion_plan.cuh"
#include <string.h>
#include <stdlib.h>

#define ERO_STAGES_MAX 8          // pipeline depth is a launch parameter (a stage is 7.3 KB: only the shared {h, w} values)
#define ERO_CONSUMER_WARPS (ERO_TILE / 32)
#define ERO_THREADS (ERO_TILE + 32)

#define ERO_MAX_PEERS 8
#define ERO_SEND_SCAN 8           // a tile with more send entries (or a vertex sent more than twice) takes the staged path

struct __align__(128) EroStage {
    float2 hw[ERO_STAGE_ELEMS];         // {height, water}; kind 1: [own tile | halo runs]; kinds 2/3: [window 264 | halo runs]
    // header, written by the producer warp before lane 0 arms the full barrier
    int32_t kind, irregular;
    int32_t nmode;                      // what the consumers prefetch for the CTA's NEXT tile (ERO_PRE_*)
    int32_t send_info;                  // multi-GPU: 0 = this tile sends nothing; sparse tile: n entries | consumer-warp mask << 16; dense tile: 0x100
    int32_t affk8[8];                   // [0..5] kinds 2/3: byte offset of slot q's neighbour relative to &hw[c]; [6], [7]: this tile's send range
    uint32_t nd[8];                     // NEXT tile, kinds 3 / 4: [0..3], [6], [7] = 3 * v0_next + D_q, q = 0..3, 4, 5 (index into dist3 of the tile's vertex 0)
    int32_t affkB8[6];                  // kind 4: the K_q of piece B (c >= split + ERO_EXC)
    int32_t split, nsplit;              // kind 4: first exception vertex of THIS tile / of the NEXT tile
    uint32_t ndB[6];                    // NEXT tile, kind 4: 3 * v0_next + D_q of piece B
    int32_t pad1[2];
    uint2 send[ERO_SEND_SCAN];          // multi-GPU, SPARSE send tile: its entries {dst, c | peer << 16}, staged by the producer warp
};

// what a consumer thread loads for itself, one tile ahead (nmode of the stage header)
#define ERO_PRE_NONE 0                  // no next tile
#define ERO_PRE_CODES 1                 // full row of six lengths + the 16-bit neighbour codes (kind 1, regular)
#define ERO_PRE_ROW 2                   // full row of six lengths (kind 2, irregular tiles)
#define ERO_PRE_D3 3                    // six entries of the one-length-per-edge table (kind 3)
#define ERO_PRE_TWO 4                   // kind 4: ERO_PRE_D3 with the constants of the vertex's piece; the exception vertices ERO_PRE_CODES

// A vertex's own per-sweep data, loaded straight from global memory ONE TILE AHEAD of its use (the
// loads are issued right after the barrier wait of the previous tile of this CTA and are consumed an
// iteration later, so neither the descriptor -> length dependency nor DRAM latency is ever waited for).
struct EroPre { float so; float d[6]; uint32_t c0, c1, c2; };


// One boundary value this rank owes a peer: vertex `c` of a tile goes to element `dst` (< 2^28) of peer
// slot `peer`'s output buffer.  Entries are grouped by tile; the tile's range [send0, send1) sits in its
// descriptor (send0 stored as -1 - send0 for a DENSE tile) and reaches the consumers through the stage header.
struct EroSendEntry { int32_t dst; uint16_t c; uint16_t peer; };

// Fused halo exchange (multi-GPU shards; all pointers null / counts zero on a single GPU):
//   * consumers store the freshly computed {h, w} pair of boundary vertices straight into the peers'
//     halo slots (NVLink-mapped peer memory) right after computing them;
//   * the last CTA to finish raises this rank's flag in every peer (after system-scope fences);
//   * the peers' flags of the previous sweep are awaited by a one-warp kernel in front of the
//     sweep (nxb_halo.cu), overlapped with this kernel's prologue by programmatic dependent launch.
struct EroComm {
    const EroSendEntry *send_entries;
    float2 *peer_hw[ERO_MAX_PEERS];     // peers' OUTPUT {height, water} buffer of this sweep
    uint32_t *peer_flag[ERO_MAX_PEERS]; // peers' flag slot for this rank
    int n_send_peers;
    uint32_t flag_value;
    unsigned int *ticket;
    // flags of the PREVIOUS sweep awaited inside this kernel (n_wait = 0: a separate wait kernel ran in front)
    const uint32_t *wait_flags;         // this rank's flag array (one slot per source rank)
    const int32_t *wait_ranks;
    int n_wait;
    uint32_t wait_target;
};

struct EroPlanArgs {
    const EroTileDesc *desc; const uint16_t *adj16;     // plan
    const int32_t *adj;                                 // int32 ELL (irregular tiles only)
    const float *dist;                                  // full table [.][6]
    const float *dist3;                                 // one entry per edge [.][3] (null: full table only)
    int use_affine;                                     // honour the plan's affine tiles (implicit adjacency)
    int use_two;                                        // honour the plan's two-piece tiles (kind 4; needs dist3)
    int wait_hint_ns;                                   // consumers' barrier wait: suspend-time hint (0: plain poll loop)
    unsigned long long *prof;                           // NXB_ERO_PROFILE builds: per-CTA cycle counters [grid][8] (else null)
    int n_stages;                                       // pipeline depth (<= ERO_STAGES_MAX)
    const float2 *hw_in;                                // {height, water} interleaved: ONE bulk copy per run
    const float *s_in;
    float2 *hw_out;
    float *s_out;
    int64_t n_own;                                      // < 2^31: vertices are indexed with 32 bits in the hot loop
    float rain;                                         // erosion.py:182-183 `water += rain`
    int rain_on_store;                                  // store water + rain (the NEXT sweep's rained water; not on the last sweep of a run)
    EroComm comm;
};

// erosion.py:210-267 for one vertex, in two steps.
//
// Step 1 (ero_slopes), applied to the neighbour values RIGHT WHERE THEY WERE LOADED (inside the branch of
// the tile kind): dh = hn - me and swq = the neighbour's rained water with the sign of dh.
//   slope = (hn - me) / (d + 1e-5): only its sign is used and d + 1e-5 > 0.
//     dh > 0: sed_amt += sol * wq, wat_amt += wq * d;   dh < 0: the same with a minus sign;
//     dh == 0 or NaN: nothing (erosion.py:232-247).
//   Branch-free: the sign bit of dh is copied onto wq, then two FMAs predicated on dh != 0 -- sol * (+-wq)
//   and (+-wq) * d are the very products the branchy form contracts to, so results are bit-identical.
// Why inside the branch: the three tile kinds fetch neighbours differently (shared memory on the hot
// paths, global gathers on the 0.1 % irregular tiles).  If the raw values merged after the branch, the
// first instruction behind the merge would wait on the scoreboard of the gathers -- the scoreboard the
// one-tile-ahead prefetch loads of EVERY tile are in flight on -- and every tile would sit out its own
// prefetch (63 % of all stall samples of v10's first cut, profiles/r02_erode3_v10_*).
// PRERAIN: the stored water already holds this sweep's rain (fl(w + rain), added by the previous sweep of
// the same run when it stored), else it is added here -- the same single rounding either way.
template <bool PRERAIN>
__device__ __forceinline__ void ero_slopes(float me, const float (&hn)[6], const float (&wn)[6], float rain,
                                           float (&dh)[6], float (&swq)[6])
{
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const float wq = PRERAIN ? wn[q] : wn[q] + rain;
        dh[q] = hn[q] - me;
        swq[q] = __uint_as_float(__float_as_uint(wq) ^ (__float_as_uint(dh[q]) & 0x80000000u));
    }
}

// Step 2: the accumulation and the deposit rule.
template <bool PRERAIN>
__device__ __forceinline__ void erode3_math(float me, float wat_own, float sed_i, const float (&dh)[6],
                                            const float (&swq)[6], const float (&d)[6], float rain,
                                            float &hh, float &ww, float &ss)
{
    const float evaporation = (float)(0.1 / 320), solubility = (float)(0.01 / 320), capacity = (float)(0.2 / 320);
    const float wat_i = PRERAIN ? wat_own : wat_own + rain;
    float sed_amt = sed_i, wat_amt = wat_i;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        asm("{\n\t.reg .pred p;\n\tsetp.ne.f32 p, %2, 0f00000000;\n\t"
            "@p fma.rn.f32 %0, %3, %4, %0;\n\t@p fma.rn.f32 %1, %4, %5, %1;\n\t}"
            : "+f"(sed_amt), "+f"(wat_amt) : "f"(dh[q]), "f"(solubility), "f"(swq[q]), "f"(d[q]));
    }
    hh = me - sed_amt;
    ss = sed_i + sed_amt;
    ww = wat_i + (wat_amt - wat_amt * evaporation);
    const float cw = capacity * ww;
    if (ss > cw) { hh += ss - cw; ss -= ss - cw; }
}

__device__ __forceinline__ float2 lds_f32x2_at(const char *base, int byte_off)
{
    return *reinterpret_cast<const float2 *>(base + byte_off);
}

// base + 4 * idx in ONE instruction (IMAD.WIDE.U32 with a register-pair addend)
__device__ __forceinline__ const float *ero_f32_at(const float *base, uint32_t idx)
{
    const float *r;
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(r) : "r"(idx), "l"(base));
    return r;
}

// kind 3: six entries of the one-length-per-edge table; i0..i5 = dist3 index of the TILE's vertex 0 for
// slot q (3 v0 + D_q), p3 = dist3 + 3 c
__device__ __forceinline__ void ero_prefetch_d3(const EroPlanArgs &a, const float *p3, uint32_t i0, uint32_t i1, uint32_t i2,
                                                uint32_t i3, uint32_t i4, uint32_t i5, uint32_t v, EroPre &p)
{
    p.so = __ldg(ero_f32_at(a.s_in, v));                        // buffers are allocated in whole tiles
    p.d[0] = __ldg(ero_f32_at(p3, i0)); p.d[1] = __ldg(ero_f32_at(p3, i1)); p.d[2] = __ldg(ero_f32_at(p3, i2));
    p.d[3] = __ldg(ero_f32_at(p3, i3)); p.d[4] = __ldg(ero_f32_at(p3, i4)); p.d[5] = __ldg(ero_f32_at(p3, i5));
}

// kinds 1 / 2 and irregular tiles: the vertex's row of the full table, kind 1 also its 16-bit neighbour codes
__device__ __forceinline__ void ero_prefetch_row(const EroPlanArgs &a, bool codes, uint32_t v, EroPre &p)
{
    p.so = __ldg(ero_f32_at(a.s_in, v));
    const uint32_t v6 = v * 6u;
    const float2 *dp = reinterpret_cast<const float2 *>(ero_f32_at(a.dist, v6));
    const float2 d0 = __ldg(dp), d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
    p.d[0] = d0.x; p.d[1] = d0.y; p.d[2] = d1.x; p.d[3] = d1.y; p.d[4] = d2.x; p.d[5] = d2.y;
    if (codes) {
        const uint32_t *ap = reinterpret_cast<const uint32_t *>(a.adj16 + v6);
        p.c0 = __ldg(ap); p.c1 = __ldg(ap + 1); p.c2 = __ldg(ap + 2);
    }
}

// COMM = false: single-GPU instantiation, every exchange-related test compiled out of the hot loop.
// PRERAIN: the input water already holds this sweep's rain (every sweep of a run but the first).
// 4 CTAs per SM (<= 56 registers; the 5-CTA / 40-register build of v8 spills once the prefetch registers exist).
template <bool COMM, bool PRERAIN>
__global__ void __launch_bounds__(ERO_THREADS, 4)
erode3_plan_kernel(const __grid_constant__ EroPlanArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    EroStage *stage = reinterpret_cast<EroStage *>(smem_raw);
    __shared__ __align__(8) uint64_t full[ERO_STAGES_MAX], empty[ERO_STAGES_MAX];
    const int n_stages = a.n_stages;
    __shared__ float2 send_hw[COMM ? ERO_TILE : 1];
    __shared__ uint32_t s_last, s_sent;
    bool cta_sent = false;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t n_own = (uint32_t)a.n_own;
    const uint32_t n_tiles = (n_own + ERO_TILE - 1) / ERO_TILE;
    const uint32_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool d3_on = a.dist3 != nullptr;

    // Programmatic dependent launch: let the next kernel of the stream (the flag wait / the next
    // sweep) become resident as this grid's CTAs retire; it blocks in its own griddepcontrol.wait
    // until this grid has completed and flushed.  No-ops when launched without the attribute.
    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < ERO_STAGES_MAX; ++s) { nxb_mbar_init(&full[s], 1); nxb_mbar_init(&empty[s], ERO_CONSUMER_WARPS); }
        nxb_fence_mbar_init();
        s_sent = 0u;
    }
    // the plan is constant: the first descriptors are fetched before the previous sweep has drained
    const int32_t *dw = reinterpret_cast<const int32_t *>(a.desc);
    int32_t word = 0, word_n = 0;               // producer warp: this lane's word of the descriptors of tiles it, it + 1
    int32_t whi = 0, whi_n = 0;                 // ... and of their second halves (two-piece constants)
    int psplit = 0;                             // consumers: first exception vertex of the CTA's first tile (kind 4)
    uint32_t idx3B[6] = {0, 0, 0, 0, 0, 0};
    const bool two_on = a.use_two && a.use_affine && d3_on;
    uint2 ent = make_uint2(0u, 0u);             // producer lane e < ERO_SEND_SCAN: entry e of the current tile's sparse send list
    int pmode = ERO_PRE_NONE;                   // consumers: what to load for the CTA's first tile
    uint32_t idx3[6] = {0, 0, 0, 0, 0, 0};
    if (my_tiles > 0) {
        const int32_t *tw = dw + (size_t)blockIdx.x * ERO_DESC_WORDS;
        if (warp == 0) {
            word = __ldg(tw + lane); whi = __ldg(tw + 32 + lane);
            if (my_tiles > 1) { word_n = __ldg(tw + (size_t)gridDim.x * ERO_DESC_WORDS + lane); whi_n = __ldg(tw + (size_t)gridDim.x * ERO_DESC_WORDS + 32 + lane); }
            if (COMM && a.comm.n_send_peers > 0) {
                const int s0 = __shfl_sync(0xffffffffu, word, ERO_DW_SEND), s1 = __shfl_sync(0xffffffffu, word, ERO_DW_SEND + 1);
                if (s0 >= 0 && lane < s1 - s0 && lane < ERO_SEND_SCAN)
                    ent = __ldg(reinterpret_cast<const uint2 *>(a.comm.send_entries) + s0 + lane);
            }
        } else {
            const int4 f = __ldg(reinterpret_cast<const int4 *>(tw + ERO_DW_NSEG));        // nseg, irregular, halo_used, d3
            const int aff = (a.use_affine && !f.y) ? __ldg(tw + ERO_DW_AFFINE) : 0;
            pmode = (aff && d3_on && (f.w & 1)) ? ERO_PRE_D3 : ((aff || f.y) ? ERO_PRE_ROW : ERO_PRE_CODES);
            if (pmode == ERO_PRE_CODES && two_on && __ldg(tw + ERO_DW_TWO)) {
                pmode = ERO_PRE_TWO;
                psplit = __ldg(tw + ERO_DW_SPLIT);
                const int2 b0 = __ldg(reinterpret_cast<const int2 *>(tw + ERO_DW_D3OFFB + 1)), b1 = __ldg(reinterpret_cast<const int2 *>(tw + ERO_DW_D3OFFB + 3));
                const uint32_t b3 = blockIdx.x * (3u * ERO_TILE);
                idx3B[0] = b3 + (uint32_t)__ldg(tw + ERO_DW_D3OFFB); idx3B[1] = b3 + (uint32_t)b0.x; idx3B[2] = b3 + (uint32_t)b0.y;
                idx3B[3] = b3 + (uint32_t)b1.x; idx3B[4] = b3 + (uint32_t)b1.y; idx3B[5] = b3 + (uint32_t)__ldg(tw + ERO_DW_D3OFFB + 5);
            }
            if (pmode == ERO_PRE_D3 || pmode == ERO_PRE_TWO) {
                const int4 o0 = __ldg(reinterpret_cast<const int4 *>(tw + ERO_DW_D3OFF));
                const int2 o1 = __ldg(reinterpret_cast<const int2 *>(tw + ERO_DW_D3OFF45));
                const uint32_t b3 = blockIdx.x * (3u * ERO_TILE);
                idx3[0] = b3 + (uint32_t)o0.x; idx3[1] = b3 + (uint32_t)o0.y; idx3[2] = b3 + (uint32_t)o0.z;
                idx3[3] = b3 + (uint32_t)o0.w; idx3[4] = b3 + (uint32_t)o1.x; idx3[5] = b3 + (uint32_t)o1.y;
            }
        }
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");      // the previous sweep's output is complete and visible
    if (COMM && a.comm.n_wait > 0) {
        // The peers' boundary values of the previous sweep have landed in this rank's halo slots once their flags
        // show that sweep's number.  Every CTA checks for itself (one lane per source rank; the flags are in
        // local memory): no separate wait kernel, one programmatic hand-over per sweep instead of two.
        if (warp == 0) {
            if (lane < a.comm.n_wait) {
                const uint32_t *f = a.comm.wait_flags + a.comm.wait_ranks[lane];
                uint32_t seen;
                for (;;) {                      // acquire at system scope: no full memory barrier per poll or per CTA
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
                    if ((int32_t)(seen - a.comm.wait_target) >= 0) break;   // flags only grow; wrap-safe
                    __nanosleep(32);
                }
            }
            __syncwarp();
            asm volatile("fence.proxy.async;" ::: "memory");    // the halo slots are read by this warp's bulk copies (async proxy)
        }
        __syncthreads();
    }

    if (warp == 0) {
        // ---------------- producer warp: lane 0 = own streams, lanes 1..ERO_NSEG = halo segments
        int s = 0;
        uint32_t ph_empty = 1;                  // parity the empty barrier of stage s must have passed
#ifdef NXB_ERO_PROFILE
        long long pt_decode = 0, pt_wait = 0, pt_hdr = 0, pt_issue = 0, pt0 = clock64(), pt_begin = pt0;
#endif
        for (uint32_t it = 0; it < my_tiles; ++it) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            const size_t v0 = (size_t)tile * ERO_TILE;
            const int32_t cur = word, nxt = word_n;     // descriptor words: see ERO_DW_*
            const int32_t cur_hi = whi, nxt_hi = whi_n;
            word = word_n; whi = whi_n;
            word_n = (it + 2 < my_tiles) ? __ldg(dw + ((size_t)tile + 2 * (size_t)gridDim.x) * ERO_DESC_WORDS + lane) : 0;
            whi_n = (it + 2 < my_tiles) ? __ldg(dw + ((size_t)tile + 2 * (size_t)gridDim.x) * ERO_DESC_WORDS + 32 + lane) : 0;
            const int nseg = __shfl_sync(0xffffffffu, cur, ERO_DW_NSEG);
            const int irregular = __shfl_sync(0xffffffffu, cur, ERO_DW_IRREGULAR);
            const int halo_used = __shfl_sync(0xffffffffu, cur, ERO_DW_HALO_USED);
            const int d3word = __shfl_sync(0xffffffffu, cur, ERO_DW_D3);
            const int affine = (a.use_affine && !irregular) ? __shfl_sync(0xffffffffu, cur, ERO_DW_AFFINE) : 0;
            const int two = (two_on && !irregular && !affine) ? __shfl_sync(0xffffffffu, cur_hi, ERO_DW_TWO - 32) : 0;
            const int kind = two ? ERO_KIND_TWO : (!affine ? ERO_KIND_CODES : ((d3_on && (d3word & 1)) ? ERO_KIND_AFFINE3 : ERO_KIND_AFFINE));
            // what the consumers load for themselves for the NEXT tile while they work on this one
            const int irr_n = __shfl_sync(0xffffffffu, nxt, ERO_DW_IRREGULAR);
            const int aff_n = (a.use_affine && !irr_n) ? __shfl_sync(0xffffffffu, nxt, ERO_DW_AFFINE) : 0;
            const int d3_n = __shfl_sync(0xffffffffu, nxt, ERO_DW_D3);
            const int two_n = (two_on && !irr_n && !aff_n) ? __shfl_sync(0xffffffffu, nxt_hi, ERO_DW_TWO - 32) : 0;
            const int nmode = it + 1 >= my_tiles ? ERO_PRE_NONE
                              : (two_n ? ERO_PRE_TWO : ((aff_n && d3_on && (d3_n & 1)) ? ERO_PRE_D3 : ((aff_n || irr_n) ? ERO_PRE_ROW : ERO_PRE_CODES)));
            const int split = __shfl_sync(0xffffffffu, cur_hi, ERO_DW_SPLIT - 32), split_n = __shfl_sync(0xffffffffu, nxt_hi, ERO_DW_SPLIT - 32);
            const int kb0 = __shfl_sync(0xffffffffu, cur_hi, ERO_DW_AFFKB - 32), kb1 = __shfl_sync(0xffffffffu, cur_hi, ERO_DW_AFFKB - 31),
                      kb2 = __shfl_sync(0xffffffffu, cur_hi, ERO_DW_AFFKB - 30);
            const int q = lane - 1;             // segment handled by this lane
            const int qq = q < 0 ? 0 : (q >= ERO_NSEG ? ERO_NSEG - 1 : q);
            const int32_t seg_start = __shfl_sync(0xffffffffu, cur, qq);
            const uint32_t lens = (uint32_t)__shfl_sync(0xffffffffu, cur, ERO_DW_LEN + (qq >> 1));
            const uint32_t offs = (uint32_t)__shfl_sync(0xffffffffu, cur, ERO_DW_OFF + (qq >> 1));
            const uint32_t seg_len = (qq & 1) ? (lens >> 16) : (lens & 0xffffu);
            const uint32_t seg_off = (qq & 1) ? (offs >> 16) : (offs & 0xffffu);
            // header words travel lane -> lane 0: K_q (3 words), this tile's send range
            const int kw0 = __shfl_sync(0xffffffffu, cur, ERO_DW_AFFK), kw1 = __shfl_sync(0xffffffffu, cur, ERO_DW_AFFK + 1),
                      kw2 = __shfl_sync(0xffffffffu, cur, ERO_DW_AFFK + 2);
            const int snd0 = __shfl_sync(0xffffffffu, cur, ERO_DW_SEND), snd1 = __shfl_sync(0xffffffffu, cur, ERO_DW_SEND + 1);
            // the NEXT tile's sparse send entries, one per lane, requested NOW and written into the next stage header
            // an iteration later (the consumers read them from shared memory: no global load on their side, and
            // none the producer has to sit out between two tiles either)
            uint2 ent_n = make_uint2(0u, 0u);
            if (COMM && a.comm.n_send_peers > 0) {
                const int s0 = __shfl_sync(0xffffffffu, nxt, ERO_DW_SEND), s1 = __shfl_sync(0xffffffffu, nxt, ERO_DW_SEND + 1);
                if (it + 1 < my_tiles && s0 >= 0 && lane < s1 - s0 && lane < ERO_SEND_SCAN)
                    ent_n = __ldg(reinterpret_cast<const uint2 *>(a.comm.send_entries) + s0 + lane);
            }
            int send_info = 0;
            if (COMM && a.comm.n_send_peers > 0 && snd1 > (snd0 < 0 ? -1 - snd0 : snd0)) {
                if (snd0 < 0) send_info = 0x100;                    // dense: staged path, entries read from global memory
                else {
                    // sparse: which consumer warps hold a vertex of the list (the others skip the scan altogether)
                    const int n = snd1 - snd0;
                    const uint32_t wbit = lane < n ? 1u << ((ent.y & 0xffffu) >> 5) : 0u;
                    send_info = n | (int)(__reduce_or_sync(0xffffffffu, wbit) << 16);
                }
            }
#ifdef NXB_ERO_PROFILE
            { const long long t = clock64(); pt_decode += t - pt0; pt0 = t; }
#endif
            nxb_mbar_wait(&empty[s], ph_empty);
#ifdef NXB_ERO_PROFILE
            { const long long t = clock64(); pt_wait += t - pt0; pt0 = t; }
#endif
            EroStage &st = stage[s];
            // lanes holding D_q of the next tile turn it into the dist3 index of that tile's vertex 0
            if (lane >= ERO_DW_D3OFF) st.nd[lane - ERO_DW_D3OFF] = (uint32_t)nxt + (tile + gridDim.x) * (3u * ERO_TILE);
            if (lane >= ERO_DW_D3OFFB - 32 && lane < ERO_DW_D3OFFB - 32 + 6)
                st.ndB[lane - (ERO_DW_D3OFFB - 32)] = (uint32_t)nxt_hi + (tile + gridDim.x) * (3u * ERO_TILE);
            if (COMM && lane < ERO_SEND_SCAN) st.send[lane] = ent;
            if (lane == 0) {
                st.kind = kind; st.irregular = irregular; st.nmode = nmode; st.send_info = send_info;
                st.affk8[0] = (int)(int16_t)(kw0 & 0xffff) * 8; st.affk8[1] = (kw0 >> 16) * 8;
                st.affk8[2] = (int)(int16_t)(kw1 & 0xffff) * 8; st.affk8[3] = (kw1 >> 16) * 8;
                st.affk8[4] = (int)(int16_t)(kw2 & 0xffff) * 8; st.affk8[5] = (kw2 >> 16) * 8;
                st.affk8[6] = snd0; st.affk8[7] = snd1;
                st.affkB8[0] = (int)(int16_t)(kb0 & 0xffff) * 8; st.affkB8[1] = (kb0 >> 16) * 8;
                st.affkB8[2] = (int)(int16_t)(kb1 & 0xffff) * 8; st.affkB8[3] = (kb1 >> 16) * 8;
                st.affkB8[4] = (int)(int16_t)(kb2 & 0xffff) * 8; st.affkB8[5] = (kb2 >> 16) * 8;
                st.split = split; st.nsplit = split_n;
            }
            __syncwarp();                       // the whole header is written before lane 0 arms the barrier
#ifdef NXB_ERO_PROFILE
            { const long long t = clock64(); pt_hdr += t - pt0; pt0 = t; }
#endif
            if (kind != ERO_KIND_CODES) {
                // window layout: [v0 - 4, v0 + 260) of {h, w}, halo runs behind it; no adjacency codes
                if (lane == 0) {
                    nxb_mbar_expect_tx(&full[s], (uint32_t)(ERO_WIN * 8) + (uint32_t)halo_used * 8u);
                    nxb_bulk_g2s(st.hw, a.hw_in + v0 - ERO_WIN_PAD, ERO_WIN * 8, &full[s]);
                } else if (q < nseg) {
                    nxb_bulk_g2s(st.hw + ERO_WIN + seg_off, a.hw_in + seg_start, seg_len * 8u, &full[s]);
                }
            } else if (lane == 0) {
                const uint32_t halo_bytes = irregular ? 0u : (uint32_t)halo_used * 8u;
                nxb_mbar_expect_tx(&full[s], (uint32_t)(ERO_TILE * 8) + halo_bytes);
                nxb_bulk_g2s(st.hw, a.hw_in + v0, ERO_TILE * 8, &full[s]);
            } else if (q < nseg && !irregular) {
                nxb_bulk_g2s(st.hw + ERO_TILE + seg_off, a.hw_in + seg_start, seg_len * 8u, &full[s]);
            }
            __syncwarp();
            ent = ent_n;
            if (++s == n_stages) { s = 0; ph_empty ^= 1u; }
#ifdef NXB_ERO_PROFILE
            { const long long t = clock64(); pt_issue += t - pt0; pt0 = t; }
#endif
        }
#ifdef NXB_ERO_PROFILE
        if (a.prof && lane == 0) {
            unsigned long long *o = a.prof + (size_t)blockIdx.x * 8;
            o[0] = (unsigned long long)pt_decode; o[1] = (unsigned long long)pt_wait; o[2] = (unsigned long long)pt_hdr;
            o[3] = (unsigned long long)pt_issue; o[7] = (unsigned long long)(clock64() - pt_begin);
        }
#endif
    } else {
        // ---------------- consumer warps: thread c owns vertex v0 + c of every tile of this CTA
        const int c = tid - 32;
        int s = 0;
        uint32_t ph_full = 0;
        // running shared-window addresses of full[s] / empty[s] and a running stage pointer: the loop
        // head otherwise re-derives them from s every tile
        const uint32_t full_a0 = nxb_smem_u32(&full[0]), empty_a0 = nxb_smem_u32(&empty[0]);
        uint32_t full_a = full_a0, empty_a = empty_a0;
        const EroStage *stp = stage;
        // ---- what this vertex alone needs comes straight from global memory, one tile AHEAD: its sediment
        // and its six edge lengths (kind 3: dist3[3 v + D_q]; kinds 1 / 2: its row of the full table; kind 1:
        // also its 16-bit neighbour codes).  These bytes -- 16 of the 36 B per vertex-sweep on kind-3 tiles --
        // bypass the bulk-copy engine.  v8 issued them at the top of the tile's own iteration, behind two
        // dependent descriptor loads: 58 % of all stall samples sat on their first use (profiles/r02_erode3_v8_*).
        // v9: the producer puts the NEXT tile's words into the stage header, the loads for tile it + 1 are
        // issued right after the barrier wait of tile it and are consumed an iteration later.
        // v10: fewer instructions per vertex (187 -> ~125 warp-instructions on a kind-3 tile): 32-bit vertex
        // and table indices, water stored already rained, slopes taken where the neighbours are loaded.
        const float *p3 = a.dist3 + 3 * c;
        uint32_t v = blockIdx.x * (uint32_t)ERO_TILE + (uint32_t)c;
        const uint32_t v_step = gridDim.x * (uint32_t)ERO_TILE;

#ifdef NXB_ERO_PROFILE
        long long ct_wait = 0, ct_work = 0, ct0 = clock64(), ct_begin = ct0;
#endif
        auto body = [&](EroPre &cur) {
#ifdef NXB_ERO_PROFILE
            { const long long t = clock64(); ct_work += t - ct0; ct0 = t; }
#endif
            if (a.wait_hint_ns) nxb_mbar_wait_hint(full_a, ph_full, (uint32_t)a.wait_hint_ns);
            else nxb_mbar_wait_a(full_a, ph_full);
#ifdef NXB_ERO_PROFILE
            { const long long t = clock64(); ct_wait += t - ct0; ct0 = t; }
#endif
            const EroStage &st = *stp;
            const int kind = st.kind;
            // ---- multi-GPU: what this tile owes the peers.  A SPARSE tile (<= ERO_SEND_SCAN entries, a vertex
            // at most twice: the row ends next to the mesh skeleton, 20 % of a shard's tiles with ~3 entries
            // each) is handled per thread: every thread scans the tile's entries -- staged in the stage header by
            // the producer warp, which fetched them a tile ahead -- and keeps the (peer, slot) pairs of ITS
            // vertex; after the math it stores its own {h, w} straight to the peer.  No block barrier, no
            // global load on the consumer side.  (Round 2, first cut: every send tile went through two named
            // barriers and a dependent load, 1.1 us per send tile, +23 us per sweep at 4 GPUs.)  DENSE tiles
            // (seam rows, the skeleton's own tiles) keep the staged path.
            uint32_t snd0 = 0xffffffffu, snd1 = 0xffffffffu;   // (peer << 28) | slot in the peer's buffer
            const int send_info = COMM ? st.send_info : 0;
            if (COMM && (send_info >> (15 + warp)) & 1) {       // sparse tile and this warp holds one of its vertices
                const int n = send_info & 0xff;
                for (int e = 0; e < n; ++e) {
                    const uint2 en = st.send[e];                // {dst, c | peer << 16}: shared-memory broadcast
                    if ((en.y & 0xffffu) == (uint32_t)c) {
                        const uint32_t packed = ((en.y >> 16) << 28) | en.x;
                        if (snd0 == 0xffffffffu) snd0 = packed; else snd1 = packed;
                    }
                }
            }
            int32_t e0 = 0, e1 = 0;
            if (COMM && send_info == 0x100) {                   // dense tile: its range of the global entry list
                const int2 sr = *reinterpret_cast<const int2 *>(st.affk8 + 6);
                e0 = -1 - sr.x; e1 = sr.y;
            }
            float dh[6], swq[6];
            float me, wo;
            if (kind == ERO_KIND_TWO) {
                // two-piece tile: window layout; K_q of the vertex's piece, the ERO_EXC vertices at the row end
                // through their explicit codes (which address the [own | halo] layout: + 4 / + 8 in the window layout)
                float hn[6], wn[6];
                const float2 own = st.hw[c + ERO_WIN_PAD];
                me = own.x; wo = own.y;
                const int split = st.split;
                if ((uint32_t)(c - split) < (uint32_t)ERO_EXC) {
                    const uint32_t code[6] = {cur.c0 & 0xffffu, cur.c0 >> 16, cur.c1 & 0xffffu, cur.c1 >> 16, cur.c2 & 0xffffu, cur.c2 >> 16};
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const uint32_t pos = code[q] & ERO_CODE_POS;
                        const float2 nq = st.hw[pos + (pos >= ERO_TILE ? 2 * ERO_WIN_PAD : ERO_WIN_PAD)];
                        hn[q] = nq.x; wn[q] = nq.y;
                    }
                } else {
                    const char *hwb = reinterpret_cast<const char *>(st.hw + c);
                    const int32_t *kk = c < split ? st.affk8 : st.affkB8;
                    const int2 k01 = *reinterpret_cast<const int2 *>(kk), k23 = *reinterpret_cast<const int2 *>(kk + 2),
                               k45 = *reinterpret_cast<const int2 *>(kk + 4);
                    const float2 n0 = lds_f32x2_at(hwb, k01.x), n1 = lds_f32x2_at(hwb, k01.y), n2 = lds_f32x2_at(hwb, k23.x),
                                 n3 = lds_f32x2_at(hwb, k23.y), n4 = lds_f32x2_at(hwb, k45.x), n5 = lds_f32x2_at(hwb, k45.y);
                    hn[0] = n0.x; wn[0] = n0.y; hn[1] = n1.x; wn[1] = n1.y; hn[2] = n2.x; wn[2] = n2.y;
                    hn[3] = n3.x; wn[3] = n3.y; hn[4] = n4.x; wn[4] = n4.y; hn[5] = n5.x; wn[5] = n5.y;
                }
                ero_slopes<PRERAIN>(me, hn, wn, a.rain, dh, swq);
            } else if (kind != ERO_KIND_CODES) {
                float hn[6], wn[6];
                // implicit adjacency: slot q's neighbour is at a per-tile constant distance from c
                const char *hwb = reinterpret_cast<const char *>(st.hw + c);
                const float2 own = st.hw[c + ERO_WIN_PAD];
                me = own.x; wo = own.y;
                const int4 ka = *reinterpret_cast<const int4 *>(st.affk8);
                const int2 kb = *reinterpret_cast<const int2 *>(st.affk8 + 4);
                const float2 n0 = lds_f32x2_at(hwb, ka.x), n1 = lds_f32x2_at(hwb, ka.y), n2 = lds_f32x2_at(hwb, ka.z),
                             n3 = lds_f32x2_at(hwb, ka.w), n4 = lds_f32x2_at(hwb, kb.x), n5 = lds_f32x2_at(hwb, kb.y);
                hn[0] = n0.x; wn[0] = n0.y; hn[1] = n1.x; wn[1] = n1.y; hn[2] = n2.x; wn[2] = n2.y;
                hn[3] = n3.x; wn[3] = n3.y; hn[4] = n4.x; wn[4] = n4.y; hn[5] = n5.x; wn[5] = n5.y;
                ero_slopes<PRERAIN>(me, hn, wn, a.rain, dh, swq);
            } else {
                float hn[6], wn[6];
                const float2 own = st.hw[c];
                me = own.x; wo = own.y;
                if (!st.irregular) {
                    const uint32_t code[6] = {cur.c0 & 0xffffu, cur.c0 >> 16, cur.c1 & 0xffffu, cur.c1 >> 16, cur.c2 & 0xffffu, cur.c2 >> 16};
#pragma unroll
                    for (int q = 0; q < 6; ++q) { const float2 nq = st.hw[code[q] & ERO_CODE_POS]; hn[q] = nq.x; wn[q] = nq.y; }
                    ero_slopes<PRERAIN>(me, hn, wn, a.rain, dh, swq);
                } else {
                    // neighbours of this tile are scattered (mesh skeleton, shard seams): global gathers
                    const uint32_t vv = v < n_own ? v : n_own - 1;
                    const int2 *rp = reinterpret_cast<const int2 *>(a.adj + (size_t)vv * 6);
                    const int2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
                    const int32_t row[6] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const uint32_t n = row[q] < 0 ? vv : (uint32_t)row[q];
                        const float2 nq = __ldg(a.hw_in + n);
                        hn[q] = nq.x; wn[q] = nq.y;
                    }
                    ero_slopes<PRERAIN>(me, hn, wn, a.rain, dh, swq);
                }
            }
            float hh, ww, ss;
            erode3_math<PRERAIN>(me, wo, cur.so, dh, swq, cur.d, a.rain, hh, ww, ss);
            // ---- the next tile's per-vertex loads, issued only now that every value prefetched for THIS tile has
            // been consumed.  All these loads share one hardware scoreboard (a counter: waiting for it waits for
            // every load in flight on it), so issuing them any earlier makes the first use of this tile's values
            // wait for the next tile's loads as well -- v10's first cut did exactly that and gained nothing from
            // prefetching.  From here the loads have the stores, the loop tail and the next tile's barrier wait,
            // shared-memory reads and slope step to land: about 0.7 of an iteration.
            // (The next tile's words are read from the stage header here, so the stage is released only now,
            // ~70 instructions later than strictly necessary: cheaper than carrying seven registers through the math.)
            {
                const int nmode = st.nmode;
                if (nmode == ERO_PRE_D3) {
                    const uint4 o0 = *reinterpret_cast<const uint4 *>(st.nd);
                    const uint2 o1 = *reinterpret_cast<const uint2 *>(st.nd + 6);
                    ero_prefetch_d3(a, p3, o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, v + v_step, cur);
                } else if (nmode == ERO_PRE_TWO) {
                    const int ns = st.nsplit;
                    if ((uint32_t)(c - ns) < (uint32_t)ERO_EXC) ero_prefetch_row(a, true, v + v_step, cur);
                    else if (c < ns) {
                        const uint4 o0 = *reinterpret_cast<const uint4 *>(st.nd);
                        const uint2 o1 = *reinterpret_cast<const uint2 *>(st.nd + 6);
                        ero_prefetch_d3(a, p3, o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, v + v_step, cur);
                    } else {
                        const uint2 o0 = *reinterpret_cast<const uint2 *>(st.ndB), o1 = *reinterpret_cast<const uint2 *>(st.ndB + 2),
                                    o2 = *reinterpret_cast<const uint2 *>(st.ndB + 4);
                        ero_prefetch_d3(a, p3, o0.x, o0.y, o1.x, o1.y, o2.x, o2.y, v + v_step, cur);
                    }
                } else if (nmode != ERO_PRE_NONE) {
                    ero_prefetch_row(a, nmode == ERO_PRE_CODES, v + v_step, cur);
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_a) : "memory");
            if (a.rain_on_store) ww += a.rain;              // the next sweep's `water += rain`, erosion.py:182-183
            if (v < n_own) { a.hw_out[v] = make_float2(hh, ww); a.s_out[v] = ss; }
            if (COMM && send_info != 0) {         // uniform over the 8 consumer warps
                if (send_info != 0x100) {
                    if (snd0 != 0xffffffffu) a.comm.peer_hw[snd0 >> 28][snd0 & 0x0fffffffu] = make_float2(hh, ww);   // one 8-byte store over NVLink
                    if (snd1 != 0xffffffffu) a.comm.peer_hw[snd1 >> 28][snd1 & 0x0fffffffu] = make_float2(hh, ww);
                } else {
                    send_hw[c] = make_float2(hh, ww);
                    asm volatile("bar.sync 1, %0;" ::"n"(ERO_TILE) : "memory");
                    for (int32_t e = e0 + c; e < e1; e += ERO_TILE) {
                        const EroSendEntry en = a.comm.send_entries[e];
                        a.comm.peer_hw[en.peer][en.dst] = send_hw[en.c];
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(ERO_TILE) : "memory");
                }
                cta_sent = true;
            }
            if (++s == n_stages) { s = 0; ph_full ^= 1u; full_a = full_a0; empty_a = empty_a0; stp = stage; }
            else { full_a += 8; empty_a += 8; ++stp; }
            v += v_step;
        };

        // ONE register set: by the time the next tile's loads are issued (behind the math) this tile's values
        // are dead, so the loads land in the very registers they are read from an iteration later
        EroPre pre;
        pre.c0 = pre.c1 = pre.c2 = 0u;
        if (pmode == ERO_PRE_TWO) {
            if ((uint32_t)(c - psplit) < (uint32_t)ERO_EXC) ero_prefetch_row(a, true, v, pre);
            else if (c < psplit) ero_prefetch_d3(a, p3, idx3[0], idx3[1], idx3[2], idx3[3], idx3[4], idx3[5], v, pre);
            else ero_prefetch_d3(a, p3, idx3B[0], idx3B[1], idx3B[2], idx3B[3], idx3B[4], idx3B[5], v, pre);
        } else if (pmode == ERO_PRE_D3) ero_prefetch_d3(a, p3, idx3[0], idx3[1], idx3[2], idx3[3], idx3[4], idx3[5], v, pre);
        else if (pmode != ERO_PRE_NONE) ero_prefetch_row(a, pmode == ERO_PRE_CODES, v, pre);
        // The first tile's values are LANDED here (one real instruction that reads them all): the loop head is a
        // merge of this prologue and the back edge, and with loads of the prologue still in flight the compiler
        // must guard the registers the loop head reuses with a wait on the prefetch scoreboard -- a wait EVERY
        // tile would then pay right behind its barrier wait.  (The bit pattern is a NaN payload no arithmetic
        // produces; the store never happens.)
        {
            const uint32_t x = __float_as_uint(pre.so) ^ __float_as_uint(pre.d[0]) ^ __float_as_uint(pre.d[1]) ^
                               __float_as_uint(pre.d[2]) ^ __float_as_uint(pre.d[3]) ^ __float_as_uint(pre.d[4]) ^
                               __float_as_uint(pre.d[5]) ^ pre.c0 ^ pre.c1 ^ pre.c2;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %0, 0x7fd5a3c1;\n\t@p st.shared.u32 [%1], %0;\n\t}"
                         :: "r"(x), "r"(nxb_smem_u32(&s_last)) : "memory");
        }
        for (uint32_t it = 0; it < my_tiles; ++it) body(pre);
#ifdef NXB_ERO_PROFILE
        if (a.prof && tid == 32) {
            unsigned long long *o = a.prof + (size_t)blockIdx.x * 8;
            o[4] = (unsigned long long)ct_wait; o[5] = (unsigned long long)ct_work; o[6] = my_tiles;
        }
#endif
    }
    if (COMM && a.comm.n_send_peers > 0) {
        // Every peer store of this CTA is visible system-wide before the CTA checks in: the consumers' stores
        // happen-before the CTA barrier, thread 0's system-scope fence behind the barrier is cumulative over
        // them (ONE fence per CTA; a fence by each of the 288 threads cost ~10 us per sweep).  The last CTA of
        // the grid then raises this rank's flag in every peer.
        if (cta_sent) s_sent = 1u;
        __syncthreads();
        if (tid == 0) {
            if (s_sent) __threadfence_system();
            s_last = (atomicAdd(a.comm.ticket, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            if (tid < a.comm.n_send_peers) {
                __threadfence_system();
                volatile uint32_t *f = a.comm.peer_flag[tid];
                *f = a.comm.flag_value;
            }
            if (tid == 0) *a.comm.ticket = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Plan builder: one CTA per tile.  Greedy covering of the tile's out-of-tile neighbour indices by
// 4-aligned windows of at most ERO_MAXSEG elements, smallest index first.
__device__ __forceinline__ int block_reduce_min(int v, int *scratch)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = scratch[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = min(r, scratch[w]);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(ERO_TILE)
ero_plan_kernel(const int32_t *__restrict__ adj, int64_t n_own, int64_t capacity,
                EroTileDesc *__restrict__ desc, uint16_t *__restrict__ adj16, int32_t *__restrict__ stats)
{
    __shared__ int scratch[ERO_TILE / 32];
    const int64_t tile = blockIdx.x, v0 = tile * ERO_TILE;
    const int c = threadIdx.x;
    const int64_t v = v0 + c;
    int32_t nb[6];
    uint32_t code[6];
    bool open[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        nb[q] = v < n_own ? adj[v * 6 + q] : -1;
        open[q] = false;
        if (nb[q] < 0) code[q] = (uint32_t)c;                                  // pad: the vertex itself
        else if (nb[q] >= v0 && nb[q] < v0 + ERO_TILE) code[q] = (uint32_t)(nb[q] - v0);
        else { code[q] = 0; open[q] = true; }
    }
    EroTileDesc d;
    memset(&d, 0, sizeof d);
    const int BIG = 0x7fffffff;
    for (int k = 0; k <= ERO_NSEG; ++k) {
        int m = BIG;
#pragma unroll
        for (int q = 0; q < 6; ++q) if (open[q]) m = min(m, nb[q]);
        m = block_reduce_min(m, scratch);
        if (m == BIG) break;                                                    // everything covered
        if (k == ERO_NSEG) { d.irregular = 1; break; }
        const int s = m & ~3;
        // window: at most ERO_MAXSEG elements, and never across the tile itself (the elements just
        // before and just after the tile must not be merged into one segment spanning it)
        int wend = s + ERO_MAXSEG;
        if (s < v0 && wend > v0) wend = (int)v0;
        // The run ends at the first gap of more than ERO_GAP unneeded elements: without this the
        // element just after the tile and the start of the next mesh row, ~60..320 indices apart on
        // rows of 260..580 vertices, were merged into one 320-slot window, the halo area overflowed
        // and a fifth of the tiles at d = 1000 fell back to global gathers.
        __shared__ unsigned occ[ERO_MAXSEG / 32];
        __shared__ int e_sh;
        if (c < ERO_MAXSEG / 32) occ[c] = 0u;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 6; ++q)
            if (open[q] && nb[q] >= s && nb[q] < wend) atomicOr(&occ[(nb[q] - s) >> 5], 1u << ((nb[q] - s) & 31));
        __syncthreads();
        if (c == 0) {
            int last = m - s, gap = 0;
            for (int i = m - s + 1; i < wend - s; ++i) {
                if (occ[i >> 5] >> (i & 31) & 1u) { last = i; gap = 0; }
                else if (++gap > ERO_GAP) break;
            }
            e_sh = s + last;
        }
        __syncthreads();
        const int e = e_sh;                                                     // last needed index of the run
        const int len = ((e + 1 - s) + 3) & ~3;
        if (d.halo_used + len > ERO_HALO_CAP || (int64_t)s + len > capacity) { d.irregular = 1; break; }
#pragma unroll
        for (int q = 0; q < 6; ++q)
            if (open[q] && nb[q] < s + len) { code[q] = (uint32_t)(ERO_TILE + d.halo_used + (nb[q] - s)); open[q] = false; }
        d.seg_start[k] = s; d.seg_len[k] = (uint16_t)len; d.seg_off[k] = (uint16_t)d.halo_used;
        d.halo_used += len;
        d.nseg = k + 1;
    }
    // ---- implicit adjacency: is the staging index of every slot's neighbour c + K_q for all c? ----
    // ---- one length per edge (kind 3): slot q's length of vertex v is dist3[3 v + D_q] ----
    //   forward slot (neighbour index > v): own row, entry = rank among v's forward slots        D_q = entry
    //   backward slot: the neighbour's row, entry = rank of v among ITS forward slots            D_q = 3 (n - v) + entry
    // A tile is kind 3 when K_q and D_q are the same for all 256 vertices and nobody has more than 3 forward
    // neighbours (rows with more -- mesh skeleton, shard seams -- keep only their first 3 lengths in dist3 and
    // must never be referenced); kind 4 (two-piece) when that holds separately for the vertices in front of and
    // behind a window of ERO_EXC exception vertices (the end of a mesh row inside the tile).
    __shared__ int k0[6], j0[6], k1[6], j1[6];
    int kq[6], jq[6];
    const bool tile_ok = !d.irregular && v0 + ERO_TILE <= n_own && v0 >= ERO_WIN_PAD && v0 + ERO_TILE + ERO_WIN_PAD <= capacity &&
                         ERO_WIN + d.halo_used <= ERO_STAGE_ELEMS;
    bool vk_ok = tile_ok, vd_ok = tile_ok;          // this vertex fits the implicit / one-length-per-edge scheme at all
    {
        int fcnt = 0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int64_t n = nb[q];
            kq[q] = 0; jq[q] = 0;
            if (!tile_ok) continue;
            if (n < 0) { vk_ok = false; vd_ok = false; continue; }
            int widx;
            if (n >= v0 - ERO_WIN_PAD && n < v0 + ERO_TILE + ERO_WIN_PAD) widx = (int)(n - v0) + ERO_WIN_PAD;
            else widx = ERO_WIN + ((int)(code[q] & ERO_CODE_POS) - ERO_TILE);
            kq[q] = widx - c;
            if (n > v) { jq[q] = fcnt; ++fcnt; }
            else {                              // (6 scattered reads per backward slot, setup only)
                int fn = 0, idx = -1;
                for (int t = 0; t < 6; ++t) {
                    const int32_t m = adj[n * 6 + t];
                    if ((int64_t)m > n) { if ((int64_t)m == v) idx = fn; ++fn; }
                }
                if (fn > 3 || idx < 0) vd_ok = false;
                jq[q] = (int)(n - v) * 3 + (idx < 0 ? 0 : idx);
            }
        }
        if (fcnt > 3) vd_ok = false;
    }
    if (c == 0) for (int q = 0; q < 6; ++q) { k0[q] = kq[q]; j0[q] = jq[q]; }
    if (c == ERO_TILE - 1) for (int q = 0; q < 6; ++q) { k1[q] = kq[q]; j1[q] = jq[q]; }
    __syncthreads();
    bool same_k0 = vk_ok, same_j0 = vk_ok && vd_ok, same_1 = vk_ok && vd_ok;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        same_k0 = same_k0 && kq[q] == k0[q];
        same_j0 = same_j0 && kq[q] == k0[q] && jq[q] == j0[q];
        same_1 = same_1 && kq[q] == k1[q] && jq[q] == j1[q];
    }
    const int all_ok = __syncthreads_and(same_k0 ? 1 : 0);
    const int all3 = __syncthreads_and(same_j0 ? 1 : 0);
    // two-piece: first vertex that does not follow vertex 0's constants, then ERO_EXC exceptions, then vertex 255's
    const int BIGC = 1 << 20;
    const int m0 = block_reduce_min(same_j0 ? BIGC : c, scratch);
    const int tail_ok = __syncthreads_and((c < m0 + ERO_EXC || same_1) ? 1 : 0);
    const int two = (tile_ok && !all3 && m0 >= 1 && m0 <= ERO_TILE - ERO_EXC - 1 && tail_ok) ? 1 : 0;
    d.affine = all_ok;
    d.d3 = all3 ? 1 : 0;
    d.two = two;
    d.split = two ? m0 : 0;
    for (int q = 0; q < 6; ++q) {
        d.aff_k[q] = (int16_t)((all_ok || two) ? k0[q] : 0);
        d.aff_kB[q] = (int16_t)(two ? k1[q] : 0);
        d.d3_offB[q] = two ? j1[q] : 0;
    }
    for (int q = 0; q < 4; ++q) d.d3_off[q] = (all3 || two) ? j0[q] : 0;
    d.d3_off45[0] = (all3 || two) ? j0[4] : 0;
    d.d3_off45[1] = (all3 || two) ? j0[5] : 0;
#pragma unroll
    for (int q = 0; q < 6; ++q) adj16[v * 6 + q] = (uint16_t)code[q];          // adj16 is allocated in whole tiles
    if (c == 0) {
        desc[tile] = d;
        if (d.irregular) atomicAdd(stats, 1);
        atomicMax(stats + 1, d.halo_used);
        if (d.affine) atomicAdd(stats + 2, 1);
        if (d.d3 & 1) atomicAdd(stats + 3, 1);
        if (d.two) atomicAdd(stats + 4, 1);
    }
}

// dist3[v][i] = length of the edge to v's i-th larger-numbered neighbour in slot order (rows with
// more than 3 such neighbours -- the mesh skeleton, shard seams -- are never referenced: their tiles
// are not kind 3)
__global__ void __launch_bounds__(256)
dist3_build_kernel(const int32_t *__restrict__ adj, const float *__restrict__ dist, int64_t n_own, int64_t n_rows,
                   float *__restrict__ dist3)
{
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_rows; v += (int64_t)gridDim.x * blockDim.x) {
        float out[3] = {0.0f, 0.0f, 0.0f};
        if (v < n_own) {
            int f = 0;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int32_t n = adj[v * 6 + q];
                if ((int64_t)n > v) { if (f < 3) out[f] = dist[v * 6 + q]; ++f; }
            }
        }
        dist3[v * 3] = out[0]; dist3[v * 3 + 1] = out[1]; dist3[v * 3 + 2] = out[2];
    }
}

// rows of the dist3 table: the padded own range plus one tile (the last affine tile's window reads 4 rows past it)
NXB_API int64_t nxb_erode_dist3_floats(int64_t n_own)
{
    const int64_t n_tiles = (n_own + ERO_TILE - 1) / ERO_TILE;
    return (n_tiles + 1) * ERO_TILE * 3;
}

NXB_API int nxb_erode_dist3_build(const int32_t *adj, const float *dist, int64_t n_own, float *dist3, void *stream)
{
    NXB_ARG(n_own >= 0);
    if (n_own == 0) return NXB_OK;
    NXB_ARG(adj && dist && dist3 && (((uintptr_t)dist3) & 15) == 0);
    const int64_t n_rows = nxb_erode_dist3_floats(n_own) / 3;
    dist3_build_kernel<<<nxb_grid_resident(dist3_build_kernel, 256, 0, (n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        adj, dist, n_own, n_rows, dist3);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

NXB_API int64_t nxb_erode_plan_bytes(int64_t n_own)
{
    const int64_t n_tiles = (n_own + ERO_TILE - 1) / ERO_TILE;
    return n_tiles * (int64_t)sizeof(EroTileDesc) + n_tiles * ERO_TILE * 6 * (int64_t)sizeof(uint16_t);
}

NXB_API int nxb_erode_plan_build(const int32_t *adj, int64_t n_own, int64_t capacity, void *plan_mem,
                                 int32_t *stats_host, void *stream)
{
    NXB_ARG(n_own >= 0 && capacity % ERO_TILE == 0 && capacity >= (n_own + ERO_TILE - 1) / ERO_TILE * ERO_TILE);
    if (n_own == 0) return NXB_OK;
    NXB_ARG(adj && plan_mem && (((uintptr_t)plan_mem) & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_tiles = (n_own + ERO_TILE - 1) / ERO_TILE;
    EroTileDesc *desc = (EroTileDesc *)plan_mem;
    uint16_t *adj16 = (uint16_t *)((char *)plan_mem + n_tiles * sizeof(EroTileDesc));
    int32_t *stats = nullptr;
    NXB_CUDA(cudaMalloc(&stats, 32));
    NXB_CUDA(cudaMemsetAsync(stats, 0, 32, st));
    ero_plan_kernel<<<(unsigned)n_tiles, ERO_TILE, 0, st>>>(adj, n_own, capacity, desc, adj16, stats);
    NXB_LAUNCH_CHECK();
    int32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    NXB_CUDA(cudaMemcpyAsync(h, stats, 32, cudaMemcpyDeviceToHost, st));
    NXB_CUDA(cudaStreamSynchronize(st));
    NXB_CUDA(cudaFree(stats));
    if (stats_host) { stats_host[0] = (int32_t)n_tiles; stats_host[1] = h[0]; stats_host[2] = h[1]; stats_host[3] = h[2]; stats_host[4] = h[3]; stats_host[5] = h[4]; }
    return NXB_OK;
}

// ---------------------------------------------------------------------------------------------
// Launch side.  The configuration (pipeline depth, grid, environment switches) is resolved once
// per call, the sweep loop runs here, not in Python.
struct EroLaunchCfg {
    int stages, use_affine, use_dist3, use_two, pdl, wait_in_sweep, wait_hint_ns;
    size_t smem;
    int grid[2];                // [COMM]
};

static bool g_ero_attr_set[64] = {false};
#ifdef NXB_ERO_PROFILE
static unsigned long long *g_ero_prof = nullptr;       // [4096][8] cycle counters of the last sweep launched
NXB_API int nxb_erode_profile_read(unsigned long long *host, int n_ctas)
{
    if (!g_ero_prof) return NXB_ERR_ARG;
    NXB_CUDA(cudaDeviceSynchronize());
    NXB_CUDA(cudaMemcpy(host, g_ero_prof, (size_t)n_ctas * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return NXB_OK;
}
#endif

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int ero_launch_cfg(int64_t n_own, EroLaunchCfg &cfg)
{
    cfg.stages = env_int("NXB_ERO_STAGES", 3);
    if (cfg.stages < 2 || cfg.stages > ERO_STAGES_MAX) cfg.stages = 3;
    cfg.use_affine = env_int("NXB_ERO_AFFINE", 1);          // read per call: tests toggle it
    cfg.use_dist3 = env_int("NXB_ERO_DIST3", 1);
    cfg.wait_hint_ns = env_int("NXB_ERO_WAIT_HINT", 0);
    cfg.use_two = env_int("NXB_ERO_TWO", 1);                // two-piece tiles (kind 4); 0: they run as kind 1
    cfg.pdl = env_int("NXB_ERO_PDL", 1);
    cfg.wait_in_sweep = env_int("NXB_ERO_WAIT_IN_SWEEP", 1);        // 0: separate one-warp wait kernel in front of every sweep
    cfg.smem = sizeof(EroStage) * cfg.stages;
    int dev = 0;
    NXB_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !g_ero_attr_set[dev]) {
        const int max_smem = (int)(sizeof(EroStage) * ERO_STAGES_MAX);
        NXB_CUDA(cudaFuncSetAttribute(erode3_plan_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        NXB_CUDA(cudaFuncSetAttribute(erode3_plan_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        NXB_CUDA(cudaFuncSetAttribute(erode3_plan_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        NXB_CUDA(cudaFuncSetAttribute(erode3_plan_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        g_ero_attr_set[dev] = true;
    }
    const int64_t n_tiles = (n_own + ERO_TILE - 1) / ERO_TILE;
    cfg.grid[0] = nxb_grid_resident(erode3_plan_kernel<false, true>, ERO_THREADS, cfg.smem, n_tiles);
    cfg.grid[1] = nxb_grid_resident(erode3_plan_kernel<true, true>, ERO_THREADS, cfg.smem, n_tiles);
    return NXB_OK;
}

// prerain: the input water already holds this sweep's rain (stored so by the previous sweep of the run)
template <bool COMM>
static int ero_launch(const EroLaunchCfg &cfg, const EroPlanArgs &a, bool prerain, cudaStream_t st)
{
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3((unsigned)cfg.grid[COMM ? 1 : 0]);
    lc.blockDim = dim3(ERO_THREADS);
    lc.dynamicSmemBytes = cfg.smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = cfg.pdl ? 1 : 0;
    if (prerain) NXB_CUDA(cudaLaunchKernelEx(&lc, erode3_plan_kernel<COMM, true>, a));
    else NXB_CUDA(cudaLaunchKernelEx(&lc, erode3_plan_kernel<COMM, false>, a));
    return NXB_OK;
}

static int ero_base_args(EroPlanArgs &a, const EroLaunchCfg &cfg, const void *plan_mem, const int32_t *adj,
                         const float *dist, const float *dist3, int64_t n_own, float rain)
{
    NXB_ARG(plan_mem && adj && dist);
    NXB_ARG((((uintptr_t)plan_mem | (uintptr_t)dist | (uintptr_t)dist3) & 15) == 0);
    NXB_ARG(n_own < ((int64_t)1 << 30));                    // 32-bit vertex / table indices in the sweep (3 v + D_q, 6 v)
    const int64_t n_tiles = (n_own + ERO_TILE - 1) / ERO_TILE;
    memset(&a, 0, sizeof a);
    a.desc = (const EroTileDesc *)plan_mem;
    a.adj16 = (const uint16_t *)((const char *)plan_mem + n_tiles * sizeof(EroTileDesc));
    a.adj = adj; a.dist = dist; a.dist3 = cfg.use_dist3 ? dist3 : nullptr;
    a.n_own = n_own; a.rain = rain;
#ifdef NXB_ERO_PROFILE
    if (!g_ero_prof) { NXB_CUDA(cudaMalloc(&g_ero_prof, 4096 * 8 * sizeof(unsigned long long))); NXB_CUDA(cudaMemset(g_ero_prof, 0, 4096 * 8 * sizeof(unsigned long long))); }
    a.prof = g_ero_prof;
#endif
    a.n_stages = cfg.stages; a.use_affine = cfg.use_affine; a.use_two = cfg.use_two; a.wait_hint_ns = cfg.wait_hint_ns;
    return NXB_OK;
}

static int ero_set_buffers(EroPlanArgs &a, const float *hw_in, const float *s_in, float *hw_out, float *s_out)
{
    NXB_ARG(hw_in && s_in && hw_out && s_out);
    NXB_ARG(hw_in != hw_out && s_in != s_out);
    NXB_ARG((((uintptr_t)hw_in | (uintptr_t)s_in | (uintptr_t)hw_out) & 15) == 0);
    a.hw_in = (const float2 *)hw_in; a.s_in = s_in;
    a.hw_out = (float2 *)hw_out; a.s_out = s_out;
    return NXB_OK;
}

// n_sweeps sweeps, ping-pong between buffer sets A and B (sweep 0 reads A): the result is in A when
// n_sweeps is even, in B when it is odd.  hw_*: {height, water} interleaved, float[capacity][2].
// Between the sweeps of one call the stored water is fl(water + rain) -- the next sweep's rained water,
// the same single rounding erosion.py:182-183 applies -- so only the first sweep adds the rain when it
// loads; the last sweep stores the plain water (what the caller sees is the reference's state).
NXB_API int nxb_erode3_run_f32(const void *plan_mem, const int32_t *adj, const float *dist, const float *dist3,
                               float *hw_a, float *s_a, float *hw_b, float *s_b,
                               int64_t n_own, float rain, int64_t n_sweeps, void *stream)
{
    NXB_ARG(n_own >= 0 && n_sweeps >= 0);
    if (n_own == 0 || n_sweeps == 0) return NXB_OK;
    EroLaunchCfg cfg;
    int rc = ero_launch_cfg(n_own, cfg);
    if (rc) return rc;
    EroPlanArgs a;
    if ((rc = ero_base_args(a, cfg, plan_mem, adj, dist, dist3, n_own, rain))) return rc;
    for (int64_t i = 0; i < n_sweeps; ++i) {
        rc = (i & 1) ? ero_set_buffers(a, hw_b, s_b, hw_a, s_a) : ero_set_buffers(a, hw_a, s_a, hw_b, s_b);
        if (rc) return rc;
        a.rain_on_store = i + 1 < n_sweeps;
        if ((rc = ero_launch<false>(cfg, a, i > 0, (cudaStream_t)stream))) return rc;
    }
    return NXB_OK;
}

NXB_API int nxb_erode3_plan_step_f32(const void *plan_mem, const int32_t *adj, const float *dist, const float *dist3,
                                     const float *hw_in, const float *s_in, float *hw_out, float *s_out,
                                     int64_t n_own, float rain, void *stream)
{
    return nxb_erode3_run_f32(plan_mem, adj, dist, dist3, (float *)hw_in, (float *)s_in, hw_out, s_out, n_own, rain, 1, stream);
}

int nxb_halo_wait_launch(const void *flags, const int32_t *src_ranks, int npeers, uint32_t target, int pdl, cudaStream_t st);

// The sharded sweep loop: per sweep a one-warp wait for the peers' flags of the previous sweep, then
// the sweep fused with the halo exchange (EroComm).  Sweep i (0-based) reads buffer set A when i is
// even; it stores boundary results into the peers' OTHER set (peer_h_b / peer_w_b when i is even),
// waits for flag value sweep_base + 1 + i and raises sweep_base + 2 + i.
NXB_API int nxb_erode3_run_comm_f32(const void *plan_mem, const int32_t *adj, const float *dist, const float *dist3,
                                    float *hw_a, float *s_a, float *hw_b, float *s_b,
                                    int64_t n_own, float rain, int64_t n_sweeps,
                                    const void *send_entries, int n_send_peers,
                                    void *const *peer_hw_a, void *const *peer_hw_b, void *const *peer_flag,
                                    const void *flags, const int32_t *wait_ranks_dev, int n_wait,
                                    uint32_t sweep_base, void *ticket, void *stream)
{
    NXB_ARG(n_own >= 0 && n_sweeps >= 0);
    NXB_ARG(n_send_peers >= 0 && n_send_peers <= ERO_MAX_PEERS && n_wait >= 0 && n_wait <= 32);
    NXB_ARG(n_send_peers == 0 || (send_entries && peer_hw_a && peer_hw_b && peer_flag && ticket));
    NXB_ARG(n_wait == 0 || (flags && wait_ranks_dev));
    if (n_sweeps == 0) return NXB_OK;
    EroLaunchCfg cfg;
    int rc = ero_launch_cfg(n_own, cfg);
    if (rc) return rc;
    EroPlanArgs a;
    if (n_own > 0 && (rc = ero_base_args(a, cfg, plan_mem, adj, dist, dist3, n_own, rain))) return rc;
    for (int64_t i = 0; i < n_sweeps; ++i) {
        const bool wait_here = n_wait > 0 && (!cfg.wait_in_sweep || n_own == 0);
        if (wait_here && (rc = nxb_halo_wait_launch(flags, wait_ranks_dev, n_wait, sweep_base + 1u + (uint32_t)i, cfg.pdl, (cudaStream_t)stream))) return rc;
        if (n_own == 0) continue;
        const bool odd = (i & 1) != 0;
        rc = odd ? ero_set_buffers(a, hw_b, s_b, hw_a, s_a) : ero_set_buffers(a, hw_a, s_a, hw_b, s_b);
        if (rc) return rc;
        memset(&a.comm, 0, sizeof a.comm);
        if (n_send_peers > 0) {
            a.comm.send_entries = (const EroSendEntry *)send_entries;
            for (int p = 0; p < n_send_peers; ++p) {
                a.comm.peer_hw[p] = (float2 *)(odd ? peer_hw_a[p] : peer_hw_b[p]);
                a.comm.peer_flag[p] = (uint32_t *)peer_flag[p];
            }
            a.comm.n_send_peers = n_send_peers;
            a.comm.ticket = (unsigned int *)ticket;
            a.comm.flag_value = sweep_base + 2u + (uint32_t)i;
        }
        if (n_wait > 0 && !wait_here) {
            a.comm.wait_flags = (const uint32_t *)flags; a.comm.wait_ranks = wait_ranks_dev;
            a.comm.n_wait = n_wait; a.comm.wait_target = sweep_base + 1u + (uint32_t)i;
        }
        a.rain_on_store = i + 1 < n_sweeps;         // what goes to the peers is what is stored: consistent on every rank
        if ((rc = ero_launch<true>(cfg, a, i > 0, (cudaStream_t)stream))) return rc;
    }
    return NXB_OK;
}

// ---------------------------------------------------------------------------------------------
// Edge lengths from caller-supplied float64 positions (erosion.py:34-40 calc_distance), FP64
// arithmetic, FP32 result.  adj rows [0, n_own) index into nodes[.][3].
__global__ void __launch_bounds__(256)
edge_lengths_kernel(const double *__restrict__ nodes, const int32_t *__restrict__ adj, int64_t n_own,
                    float *__restrict__ dist)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_own * 6; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = i / 6;
        const int32_t n = adj[i];
        float r = 0.0f;
        if (n >= 0) {
            const double ax = nodes[3 * v] - nodes[3 * (int64_t)n], ay = nodes[3 * v + 1] - nodes[3 * (int64_t)n + 1],
                         az = nodes[3 * v + 2] - nodes[3 * (int64_t)n + 2];
            r = (float)sqrt(ax * ax + ay * ay + az * az);
        }
        dist[i] = r;
    }
}

NXB_API int nxb_edge_lengths_f64(const double *nodes, const int32_t *adj, int64_t n_own, float *dist, void *stream)
{
    NXB_ARG(n_own >= 0);
    if (n_own == 0) return NXB_OK;
    NXB_ARG(nodes && adj && dist);
    int grid = nxb_grid_resident(edge_lengths_kernel, 256, 0, (n_own * 6 + 255) / 256);
    edge_lengths_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(nodes, adj, n_own, dist);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// ---------------------------------------------------------------------------------------------
// erosion.py:76-99
__global__ void __launch_bounds__(256)
erode1_kernel(const int32_t *__restrict__ adj, const float *__restrict__ h_in, float *__restrict__ h_out,
              int64_t v_begin, int64_t v_end)
{
    for (int64_t i = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v_end;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int2 *p = reinterpret_cast<const int2 *>(adj + i * 6);
        const int2 r0 = __ldg(p), r1 = __ldg(p + 1), r2 = __ldg(p + 2);
        const int32_t row[6] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
        const float me = h_in[i];
        float hn[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) hn[q] = __ldg(h_in + (row[q] < 0 ? i : (int64_t)row[q]));
        float amt = 0.0f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            if (hn[q] > me) amt += 0.0005f;
            else if (hn[q] < me) amt -= 0.0005f;
        }
        h_out[i] = me + amt;
    }
}

NXB_API int nxb_erode1_step_f32(const int32_t *adj, const float *h_in, float *h_out,
                                int64_t v_begin, int64_t v_end, void *stream)
{
    NXB_ARG(v_begin >= 0 && v_end >= v_begin);
    if (v_end == v_begin) return NXB_OK;
    NXB_ARG(adj && h_in && h_out && h_in != h_out);
    int grid = nxb_grid_resident(erode1_kernel, 256, 0, (v_end - v_begin + 255) / 256);
    erode1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(adj, h_in, h_out, v_begin, v_end);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

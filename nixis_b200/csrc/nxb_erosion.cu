// Erosion sweeps -- erosion.py:34-40, 76-99, 197-279 -- as a shared-memory-staged stencil.
//
// Bound by HBM and by consumer instruction issue at the same time (profiles/r01_ncu_*).
// Algorithmic traffic per vertex-iteration of erosion_iteration3:
//   own h,w,s read 12 B + write 12 B + adjacency row 24 B + own position 12 B = 60 B.
// What this implementation actually streams per vertex-iteration:
//   h,w,s read 12 B + write 12 B + 6 edge lengths 24 B = 48 B on affine tiles (implicit adjacency,
//   ~80 % of the tiles at d = 2500), + 12 B of 16-bit tile-local adjacency on the other tiles.
//
// History (profiles/r01_ncu_summary.json): v1, one thread per vertex with 18 global gathers
// (xyz, h, w of 6 neighbours), ran at 45 % of HBM peak with minimal DRAM traffic -- latency bound;
// v2 streamed the tile's own data with cp.async.bulk but kept the gathers and was then bound by
// gather latency + instruction issue (322 instr/vertex, six IEEE sqrt sequences, L1 squeezed out by
// the shared-memory carve-out).  v3 (this file) removes both:
//
//   * EDGE LENGTHS ARE PRECOMPUTED once, in FP64, and stored as FP32 (erosion.py:227-229 uses the
//     undisplaced sphere positions, so they never change).  No neighbour positions are read and no
//     sqrt is evaluated in the sweep; it also removes the cancellation error of differencing FP32
//     positions (relative 1e-4 at d=2500).
//   * TILE PLAN (nxb_erosion_plan.cuh): for each tile of 256 consecutive vertices the neighbours
//     are the tile itself plus <= 8 contiguous index runs (mesh rows above / below, the elements
//     next to the tile ends).  A producer warp brings the tile's own streams AND those runs of
//     h / w into shared memory with cp.async.bulk (TMA bulk copy, SASS UBLKCP) completing on an
//     mbarrier, n_stages tiles ahead; eight consumer warps then read every neighbour value from
//     shared memory through a 16-bit tile-local adjacency.  No global gathers, no L1 dependence.
//     The halo runs were just streamed by a neighbouring tile, so they come from L2, not HBM.
//   * IMPLICIT ADJACENCY (nxb_erosion_plan.cuh, kind 2): tiles whose six neighbour distances are the
//     same for all 256 vertices are staged as a 264-element window + halo runs and swept without any
//     adjacency codes: -12 B and -27 instructions per vertex, 0.59 -> 0.51-0.56 ms per sweep at d=2500.
//   * ONE LENGTH PER EDGE (dist3, nxb_erosion_plan.cuh) is implemented and bit-identical, but OFF by
//     default: it cuts the DRAM reads from 3.16 GB to 2.53 GB per sweep at d = 2500 (ncu) and still
//     runs 715-750 us against 590 us, because at 590 us the consumer warps are already issue-bound
//     (72 % issue-slot utilisation, ~230 instructions per vertex) and decoding which row holds a
//     slot's length adds ~60 instructions per vertex.  Enable with NXB_ERO_DIST3=1.
//   * ping-pong buffers replace the reference's three np.copy + copy-back pass (erosion.py:199-201,
//     274-277); `water += rain` (erosion.py:182-183) is fused into the reads.
#include "nxb_common.cuh"
#include "nxb_erosion_plan.cuh"
#include <string.h>
#include <stdlib.h>

#define ERO_STAGES_MAX 4          // pipeline depth is a launch parameter (3: 4 CTAs/SM, 4: 3 CTAs/SM)
#define ERO_CONSUMER_WARPS (ERO_TILE / 32)
#define ERO_THREADS (ERO_TILE + 32)

struct __align__(128) EroStage {
    float h[ERO_STAGE_ELEMS];           // [own tile | halo segments]
    float w[ERO_STAGE_ELEMS];
    float s[ERO_TILE];
    float dist[ERO_TILE * 3 + ERO_D3_CAP * 3];   // kind 1: [256][6] full rows; kind 0: dist3 [own 256 | staged halo slots]
    uint16_t adj[ERO_TILE * 6];
    float exc[ERO_EXC * 6];             // kind 0: full rows of the tile's heavy vertices
    int32_t irregular;
    int32_t tile;
    int32_t kind;                       // 0: dist holds dist3 rows, 1: full rows, 2: affine tile (window layout, no codes)
    int32_t pad0;
    int32_t affk4[6];                   // kind 2: byte offset of slot q's neighbour relative to &h[c] / &w[c]
    int32_t pad[30];
};

#define ERO_MAX_PEERS 8

// One boundary value this rank owes a peer: vertex `c` of a tile goes to element `dst` of peer slot
// `peer`'s output buffers.  Entries are grouped by tile (CSR: send_ptr[tile] .. send_ptr[tile+1]).
struct EroSendEntry { int32_t dst; uint16_t c; uint16_t peer; };

// Fused halo exchange (multi-GPU shards; all pointers null / counts zero on a single GPU):
//   * consumers store the freshly computed h / w of boundary vertices straight into the peers'
//     halo slots (NVLink-mapped peer memory) right after computing them;
//   * the producer warp spins on this rank's flag words only before the first tile that needs halo
//     data (a segment in the halo area, or an irregular tile);
//   * the last CTA to finish raises this rank's flag in every peer (after system-scope fences).
struct EroComm {
    const int32_t *send_ptr;            // [n_tiles + 1], null = no sends
    const EroSendEntry *send_entries;
    float *peer_h[ERO_MAX_PEERS], *peer_w[ERO_MAX_PEERS];   // peers' OUTPUT buffers of this sweep
    uint32_t *peer_flag[ERO_MAX_PEERS]; // peers' flag slot for this rank
    int n_send_peers;
    const uint32_t *flags;              // this rank's flag array (written by the peers)
    int32_t wait_rank[ERO_MAX_PEERS];
    int n_wait;
    uint32_t wait_target, flag_value;
    int64_t halo_begin;                 // first halo slot (n_own_pad)
    unsigned int *ticket;
    // processing order of the tiles (null = index order).  The sharded driver puts the tiles that
    // read halo slots LAST, so by the time a CTA reaches them the peers' flags are already up and
    // the wait hides behind interior work.
    const int32_t *tile_order;
    // > 0: the first n_early tiles of the processing order are the boundary set (every tile that
    // sends or reads halo slots).  The flags are raised as soon as those tiles are done -- their
    // boundary values are in the peers' memory and this rank no longer reads its halo slots of this
    // sweep -- not at the end of the grid, so the flag latency, the peers' launch gap and this rank's
    // tail hide behind the interior tiles.
    int n_early;
};

struct EroPlanArgs {
    const EroTileDesc *desc; const uint16_t *adj16;     // plan
    const int32_t *adj;                                 // int32 ELL (irregular tiles only)
    const float *dist;                                  // full table [.][6]
    const float *dist3;                                 // one entry per edge [.][3] (null: full table only)
    const float *exc;                                   // [n_tiles][ERO_EXC][6] rows of heavy vertices
    int use_affine;                                     // honour the plan's affine tiles (implicit adjacency)
    int n_stages;                                       // pipeline depth (<= ERO_STAGES_MAX)
    const float *h_in, *w_in, *s_in;
    float *h_out, *w_out, *s_out;
    int64_t n_own;
    float rain;
    EroComm comm;
};

// erosion.py:210-267 for one vertex, neighbour values already fetched.
__device__ __forceinline__ void erode3_math(float me, float wat_own, float sed_i, const float (&hn)[6],
                                            const float (&wn)[6], const float (&d)[6], float rain,
                                            float &hh, float &ww, float &ss)
{
    const float evaporation = (float)(0.1 / 320), solubility = (float)(0.01 / 320), capacity = (float)(0.2 / 320);
    const float wat_i = wat_own + rain;
    float sed_amt = sed_i, wat_amt = wat_i;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const float wq = wn[q] + rain;
        // slope = (hn - me) / (d + 1e-5): only its sign is used and d + 1e-5 > 0.
        //   dh > 0: sed_amt += sol * wq, wat_amt += wq * d;   dh < 0: the same with a minus sign;
        //   dh == 0 or NaN: nothing (erosion.py:232-247).
        // Branch-free: the sign bit of dh is copied onto sol and wq, then two predicated FMAs --
        // the very FMAs the branchy form contracts to, so results are bit-identical to it.
        const float dh = hn[q] - me;
        const uint32_t sgn = __float_as_uint(dh) & 0x80000000u;
        const float ssol = __uint_as_float(__float_as_uint(solubility) | sgn);
        const float swq = __uint_as_float(__float_as_uint(wq) ^ sgn);
        asm("{\n\t.reg .pred p;\n\tsetp.ne.f32 p, %2, 0f00000000;\n\t"
            "@p fma.rn.f32 %0, %3, %4, %0;\n\t@p fma.rn.f32 %1, %5, %6, %1;\n\t}"
            : "+f"(sed_amt), "+f"(wat_amt) : "f"(dh), "f"(ssol), "f"(wq), "f"(swq), "f"(d[q]));
    }
    hh = me - sed_amt;
    ss = sed_i + sed_amt;
    ww = wat_i + (wat_amt - wat_amt * evaporation);
    const float cw = capacity * ww;
    if (ss > cw) { hh += ss - cw; ss -= ss - cw; }
}

// COMM = false: single-GPU instantiation, every exchange-related test compiled out of the hot loop
// (the sweep is issue-co-limited: each instruction per vertex counts).
template <bool COMM>
__global__ void __launch_bounds__(ERO_THREADS)
erode3_plan_kernel(const __grid_constant__ EroPlanArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    EroStage *stage = reinterpret_cast<EroStage *>(smem_raw);
    __shared__ __align__(8) uint64_t full[ERO_STAGES_MAX], empty[ERO_STAGES_MAX];
    const int n_stages = a.n_stages;
    __shared__ float send_h[ERO_TILE], send_w[ERO_TILE];
    __shared__ bool s_last;
    bool cta_sent = false;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (a.n_own + ERO_TILE - 1) / ERO_TILE;
    const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < ERO_STAGES_MAX; ++s) { nxb_mbar_init(&full[s], 1); nxb_mbar_init(&empty[s], ERO_CONSUMER_WARPS); }
        nxb_fence_mbar_init();
    }
    __syncthreads();

#ifdef NXB_ERO_DEBUG_WAIT
    unsigned long long dbg_t0 = 0; long long dbg_c0 = 0;
    if (blockIdx.x == 0 && tid == 0 && a.comm.ticket) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
        dbg_c0 = clock64();
        a.comm.ticket[4 + 8 * (a.comm.flag_value % 32) + 0] = (unsigned)dbg_t0;
    }
#endif
    if (warp == 0) {
        // ---------------- producer warp: lane 0 = own streams, lanes 1..ERO_NSEG = halo segments
        const int32_t *dw = reinterpret_cast<const int32_t *>(a.desc);
        int32_t word = 0;                       // this lane's word of the 16-word descriptor
        bool halo_ready = !COMM || a.comm.n_wait == 0;
        const int32_t *order = COMM ? a.comm.tile_order : nullptr;
        // order lookups are batched: lane l holds the entry of iteration 32 b + l, the next batch is
        // already in flight.  (One dependent __ldg per tile in front of the descriptor load made the
        // producer latency-bound: +30% per sweep.)  tile_of is called with i = 0, 1, 2, ... in turn.
        int32_t ord_cur = 0, ord_nxt = 0;
        auto ord_load = [&](int64_t b) -> int32_t {
            const int64_t i = b * 32 + lane;
            return i < my_tiles ? __ldg(order + blockIdx.x + i * gridDim.x) : 0;
        };
        if (order) { ord_cur = ord_load(0); ord_nxt = ord_load(1); }
        auto tile_of = [&](int64_t i) -> int64_t {       // i-th tile of this CTA
            if (!order) return blockIdx.x + i * gridDim.x;
            if (i > 0 && (i & 31) == 0) { ord_cur = ord_nxt; ord_nxt = ord_load((i >> 5) + 1); }
            return (int64_t)__shfl_sync(0xffffffffu, ord_cur, (int)(i & 31));
        };
        int64_t tile = my_tiles > 0 ? tile_of(0) : 0;
        int64_t tile_next = my_tiles > 1 ? tile_of(1) : 0;
        if (my_tiles > 0) word = __ldg(dw + tile * ERO_DESC_WORDS + lane);
        int s = 0;
        uint32_t ph_empty = 1;                  // parity the empty barrier of stage s must have passed
        for (int it = 0; it < (int)my_tiles; ++it) {
            const int64_t v0 = tile * ERO_TILE;
            const int32_t tile_id = (int32_t)tile;
            // descriptor words: see ERO_DW_* (seg_start | seg_len pairs | seg_off pairs | nseg | irregular | halo_used | d3 | affine | K_q)
            const int32_t cur = word;
            if (it + 1 < my_tiles) word = __ldg(dw + tile_next * ERO_DESC_WORDS + lane);
            tile = tile_next;
            if (it + 2 < my_tiles) tile_next = tile_of(it + 2);
            const int nseg = __shfl_sync(0xffffffffu, cur, ERO_DW_NSEG);
            const int irregular = __shfl_sync(0xffffffffu, cur, ERO_DW_IRREGULAR);
            const int halo_used = __shfl_sync(0xffffffffu, cur, ERO_DW_HALO_USED);
            const int d3word = __shfl_sync(0xffffffffu, cur, ERO_DW_D3);
            const int affine = (a.use_affine && !irregular) ? __shfl_sync(0xffffffffu, cur, ERO_DW_AFFINE) : 0;
            const int kw0 = __shfl_sync(0xffffffffu, cur, ERO_DW_AFFK), kw1 = __shfl_sync(0xffffffffu, cur, ERO_DW_AFFK + 1), kw2 = __shfl_sync(0xffffffffu, cur, ERO_DW_AFFK + 2);
            const int kind = affine ? ERO_KIND_AFFINE : ((a.dist3 == nullptr || irregular) ? 1 : (d3word & 0xff));   // 0: dist3, 1: full rows
            const uint32_t d3_used = kind == 0 ? (uint32_t)(d3word >> 8) : 0u;
            const int q = lane - 1;             // segment handled by this lane
            const int qq = q < 0 ? 0 : (q >= ERO_NSEG ? ERO_NSEG - 1 : q);
            const int32_t seg_start = __shfl_sync(0xffffffffu, cur, qq);
            const uint32_t lens = (uint32_t)__shfl_sync(0xffffffffu, cur, ERO_DW_LEN + (qq >> 1));
            const uint32_t offs = (uint32_t)__shfl_sync(0xffffffffu, cur, ERO_DW_OFF + (qq >> 1));
            const uint32_t seg_len = (qq & 1) ? (lens >> 16) : (lens & 0xffffu);
            const uint32_t seg_off = (qq & 1) ? (offs >> 16) : (offs & 0xffffu);
            if (!halo_ready) {
                // does this tile read halo slots?  (a segment in the halo area, or global gathers)
                // (an affine tile reads the 4 elements after its end through its window, not a run)
                const bool mine = (q >= 0 && q < nseg && (int64_t)seg_start >= a.comm.halo_begin) || irregular ||
                                  (affine && v0 + ERO_TILE + ERO_WIN_PAD > a.comm.halo_begin);
                if (__any_sync(0xffffffffu, mine)) {
                    if (lane < a.comm.n_wait) {
                        const volatile uint32_t *f = a.comm.flags + a.comm.wait_rank[lane];
                        // gentle polling: hundreds of CTAs hammering one L2 line delay the very
                        // NVLink write they are waiting for
                        unsigned ns = 64;
#ifdef NXB_ERO_DEBUG_WAIT
                        unsigned long long t0, t1;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                        bool spun = false;
#endif
                        while ((int32_t)(*f - a.comm.wait_target) < 0) {
                            __nanosleep(ns); if (ns < 1024) ns *= 2;
#ifdef NXB_ERO_DEBUG_WAIT
                            spun = true;
#endif
                        }
#ifdef NXB_ERO_DEBUG_WAIT
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (spun) { atomicAdd(a.comm.ticket + 1, 1u); atomicMax(a.comm.ticket + 2, (unsigned)(t1 - t0)); }
                        atomicMax(a.comm.ticket + 3, (unsigned)it);
                        if (blockIdx.x == 0 && lane == 0) a.comm.ticket[4 + 8 * (a.comm.flag_value % 32) + 1] = (unsigned)t1;
#endif
                    }
                    __threadfence_system();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __syncwarp();
                    halo_ready = true;
                }
            }
            nxb_mbar_wait(&empty[s], ph_empty);
            EroStage &st = stage[s];
            if (kind == ERO_KIND_AFFINE) {
                // window layout: [v0 - 4, v0 + 260) of h and w, halo runs behind it; no adjacency codes
                if (lane == 0) {
                    st.irregular = 0; st.tile = tile_id; st.kind = kind;
                    st.affk4[0] = (int)(int16_t)(kw0 & 0xffff) * 4; st.affk4[1] = (kw0 >> 16) * 4;
                    st.affk4[2] = (int)(int16_t)(kw1 & 0xffff) * 4; st.affk4[3] = (kw1 >> 16) * 4;
                    st.affk4[4] = (int)(int16_t)(kw2 & 0xffff) * 4; st.affk4[5] = (kw2 >> 16) * 4;
                    nxb_mbar_expect_tx(&full[s], (uint32_t)(ERO_WIN * 8 + ERO_TILE * (4 + 24)) + (uint32_t)halo_used * 8u);
                    nxb_bulk_g2s(st.h, a.h_in + v0 - ERO_WIN_PAD, ERO_WIN * 4, &full[s]);
                    nxb_bulk_g2s(st.w, a.w_in + v0 - ERO_WIN_PAD, ERO_WIN * 4, &full[s]);
                    nxb_bulk_g2s(st.s, a.s_in + v0, ERO_TILE * 4, &full[s]);
                    nxb_bulk_g2s(st.dist, a.dist + v0 * 6, ERO_TILE * 24, &full[s]);
                } else if (q < nseg) {
                    nxb_bulk_g2s(st.h + ERO_WIN + seg_off, a.h_in + seg_start, seg_len * 4u, &full[s]);
                    nxb_bulk_g2s(st.w + ERO_WIN + seg_off, a.w_in + seg_start, seg_len * 4u, &full[s]);
                }
            } else if (lane == 0) {
                st.irregular = irregular;
                st.tile = tile_id;
                st.kind = kind;
                const uint32_t halo_bytes = irregular ? 0u : (uint32_t)halo_used * 8u;
                const uint32_t dist_bytes = kind == 0 ? (uint32_t)(ERO_TILE * 12 + ERO_EXC * 24) + d3_used * 12u : (uint32_t)(ERO_TILE * 24);
                nxb_mbar_expect_tx(&full[s], (uint32_t)(ERO_TILE * (4 * 3 + 12)) + dist_bytes + halo_bytes);
                nxb_bulk_g2s(st.h, a.h_in + v0, ERO_TILE * 4, &full[s]);
                nxb_bulk_g2s(st.w, a.w_in + v0, ERO_TILE * 4, &full[s]);
                nxb_bulk_g2s(st.s, a.s_in + v0, ERO_TILE * 4, &full[s]);
                if (kind == 0) {
                    nxb_bulk_g2s(st.dist, a.dist3 + v0 * 3, ERO_TILE * 12, &full[s]);
                    nxb_bulk_g2s(st.exc, a.exc + (int64_t)tile_id * (ERO_EXC * 6), ERO_EXC * 24, &full[s]);
                }
                else           nxb_bulk_g2s(st.dist, a.dist + v0 * 6, ERO_TILE * 24, &full[s]);
                nxb_bulk_g2s(st.adj, a.adj16 + v0 * 6, ERO_TILE * 12, &full[s]);
            } else if (q < nseg && !irregular) {
                nxb_bulk_g2s(st.h + ERO_TILE + seg_off, a.h_in + seg_start, seg_len * 4u, &full[s]);
                nxb_bulk_g2s(st.w + ERO_TILE + seg_off, a.w_in + seg_start, seg_len * 4u, &full[s]);
                if (seg_off < d3_used) {
                    // dist3 rows of the (smaller-numbered) vertices of this run: owners of backward edges
                    const uint32_t n3 = min(seg_len, d3_used - seg_off);
                    nxb_bulk_g2s(st.dist + (ERO_TILE + seg_off) * 3, a.dist3 + (int64_t)seg_start * 3, n3 * 12u, &full[s]);
                }
            }
            __syncwarp();
            if (++s == n_stages) { s = 0; ph_empty ^= 1u; }
        }
    } else {
        // ---------------- consumer warps: thread c owns vertex v0 + c of every tile of this CTA
        const int c = tid - 32;
        int s = 0;
        uint32_t ph_full = 0;
        const bool sending = COMM && a.comm.send_ptr != nullptr;
        const int n_early = COMM ? a.comm.n_early : 0;
        // running shared-window addresses of full[s] / empty[s] and a running stage pointer: the loop
        // head otherwise re-derives them from s every tile (the sweep is issue-co-limited)
        const uint32_t full_a0 = nxb_smem_u32(&full[0]), empty_a0 = nxb_smem_u32(&empty[0]);
        uint32_t full_a = full_a0, empty_a = empty_a0;
        const EroStage *stp = stage;
        for (int it = 0; it < (int)my_tiles; ++it) {
            nxb_mbar_wait_a(full_a, ph_full);   // (sleeping between polls lowers power, not time: measured, dropped)
            const EroStage &st = *stp;
            const int64_t tile = st.tile;
            const int64_t v = tile * (int64_t)ERO_TILE + c;
            float hn[6], wn[6], d[6];
            float me, wo, so;
            if (st.kind == ERO_KIND_AFFINE) {
                // implicit adjacency: slot q's neighbour is at a per-tile constant distance from c
                const char *hb = reinterpret_cast<const char *>(st.h + c), *wb = reinterpret_cast<const char *>(st.w + c);
                me = st.h[c + ERO_WIN_PAD]; wo = st.w[c + ERO_WIN_PAD]; so = st.s[c];
                const float2 *dp = reinterpret_cast<const float2 *>(st.dist + c * 6);
                const float2 d0 = dp[0], d1 = dp[1], d2 = dp[2];
                d[0] = d0.x; d[1] = d0.y; d[2] = d1.x; d[3] = d1.y; d[4] = d2.x; d[5] = d2.y;
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const int k4 = st.affk4[q];
                    hn[q] = *reinterpret_cast<const float *>(hb + k4);
                    wn[q] = *reinterpret_cast<const float *>(wb + k4);
                }
            } else {
            me = st.h[c]; wo = st.w[c]; so = st.s[c];
            const uint32_t *ap = reinterpret_cast<const uint32_t *>(st.adj + c * 6);
            const uint32_t a0 = ap[0], a1 = ap[1], a2 = ap[2];
            const uint32_t code[6] = {a0 & 0xffffu, a0 >> 16, a1 & 0xffffu, a1 >> 16, a2 & 0xffffu, a2 >> 16};
            if (st.kind == 0 && !(code[0] & ERO_CODE_HEAVY)) {
                // one stored length per edge: from this vertex's dist3 row, or from the row of the
                // (smaller-numbered, staged) neighbour that owns the edge
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const uint32_t row = (code[q] & ERO_CODE_BACK) ? (code[q] & ERO_CODE_POS) : (uint32_t)c;
                    d[q] = st.dist[row * 3 + ((code[q] >> ERO_CODE_I_SHIFT) & 3u)];
                }
            } else if (st.kind == 0) {
                // skeleton-adjacent vertex: its full row travels with the tile (exception rows)
                const float *ep = st.exc + ((code[0] >> ERO_CODE_EXC_SHIFT) & 3u) * 6;
#pragma unroll
                for (int q = 0; q < 6; ++q) d[q] = ep[q];
            } else {
                const float2 *dp = reinterpret_cast<const float2 *>(st.dist + c * 6);
                const float2 d0 = dp[0], d1 = dp[1], d2 = dp[2];
                d[0] = d0.x; d[1] = d0.y; d[2] = d1.x; d[3] = d1.y; d[4] = d2.x; d[5] = d2.y;
            }
            if (!st.irregular) {
#pragma unroll
                for (int q = 0; q < 6; ++q) { hn[q] = st.h[code[q] & ERO_CODE_POS]; wn[q] = st.w[code[q] & ERO_CODE_POS]; }
            } else {
                // neighbours of this tile are scattered (mesh skeleton, shard seams): global gathers
                const int64_t vv = v < a.n_own ? v : a.n_own - 1;
                const int2 *rp = reinterpret_cast<const int2 *>(a.adj + vv * 6);
                const int2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
                const int32_t row[6] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const int64_t n = row[q] < 0 ? vv : (int64_t)row[q];
                    hn[q] = __ldg(a.h_in + n); wn[q] = __ldg(a.w_in + n);
                }
            }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_a) : "memory");
            float hh, ww, ss;
            erode3_math(me, wo, so, hn, wn, d, a.rain, hh, ww, ss);
            if (v < a.n_own) { a.h_out[v] = hh; a.w_out[v] = ww; a.s_out[v] = ss; }
            if (sending) {
                const int32_t e0 = __ldg(a.comm.send_ptr + tile), e1 = __ldg(a.comm.send_ptr + tile + 1);
                if (e1 > e0) {                  // uniform over the 8 consumer warps
                    send_h[c] = hh; send_w[c] = ww;
                    asm volatile("bar.sync 1, %0;" ::"n"(ERO_TILE) : "memory");
                    for (int32_t e = e0 + c; e < e1; e += ERO_TILE) {
                        const EroSendEntry en = a.comm.send_entries[e];
                        a.comm.peer_h[en.peer][en.dst] = send_h[en.c];
                        a.comm.peer_w[en.peer][en.dst] = send_w[en.c];
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(ERO_TILE) : "memory");
                    cta_sent = true;
                }
            }
            const int slot = (int)blockIdx.x + it * (int)gridDim.x;
            if (COMM && slot < n_early && slot + (int)gridDim.x >= n_early) {
                // this CTA's LAST boundary tile is done: all 8 consumer warps have read their inputs
                // and stored to the peers.  One system fence per CTA (a fence per tile costs
                // microseconds each while NVLink stores are in flight), then check in.
                asm volatile("bar.sync 1, %0;" ::"n"(ERO_TILE) : "memory");
                if (c == 0) {
                    __threadfence_system();         // cumulative over the CTA's peer stores (barrier above)
                    const unsigned n_cta = a.comm.n_early < (int)gridDim.x ? (unsigned)a.comm.n_early : gridDim.x;
                    if (atomicAdd(a.comm.ticket, 1u) == n_cta - 1u) {
                        *a.comm.ticket = 0;
                        __threadfence_system();
                        for (int p = 0; p < a.comm.n_send_peers; ++p) {
                            volatile uint32_t *f = a.comm.peer_flag[p];
                            *f = a.comm.flag_value;
                        }
                        __threadfence_system();
#ifdef NXB_ERO_DEBUG_WAIT
                        { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                          a.comm.ticket[4 + 8 * (a.comm.flag_value % 32) + 2] = (unsigned)t; }
#endif
                    }
                }
            }
            if (++s == n_stages) { s = 0; ph_full ^= 1u; full_a = full_a0; empty_a = empty_a0; stp = stage; }
            else { full_a += 8; empty_a += 8; ++stp; }
        }
    }
#ifdef NXB_ERO_DEBUG_WAIT
    if (a.comm.ticket && (tid == 0 || tid == 32)) {           // when did this CTA's producer / consumers run dry
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(a.comm.ticket + 4 + 8 * (a.comm.flag_value % 32) + (tid == 0 ? 4 : 5), (unsigned)t);
    }
    if (blockIdx.x == 0 && tid == 0 && a.comm.ticket) {       // SM clock (MHz) seen by CTA 0 over its lifetime
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        long long c1 = clock64();
        a.comm.ticket[4 + 8 * (a.comm.flag_value % 32) + 3] = (unsigned)((c1 - dbg_c0) * 1000 / (long long)(t1 - dbg_t0 + 1));
    }
#endif
    if (COMM && a.comm.n_send_peers > 0 && a.comm.n_early == 0) {
        // every peer store of this CTA is visible system-wide before the CTA checks in; the last
        // CTA of the grid then raises this rank's flag in every peer
        if (cta_sent) __threadfence_system();
        __syncthreads();
#ifdef NXB_ERO_DEBUG_WAIT
        if (tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                        atomicMax(a.comm.ticket + 4 + 8 * (a.comm.flag_value % 32) + 6, (unsigned)t); }
#endif
        if (tid == 0) s_last = (atomicAdd(a.comm.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            if (tid < a.comm.n_send_peers) {
                __threadfence_system();
                volatile uint32_t *f = a.comm.peer_flag[tid];
                *f = a.comm.flag_value;
                __threadfence_system();
            }
            if (tid == 0) *a.comm.ticket = 0;
#ifdef NXB_ERO_DEBUG_WAIT
            if (tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                            a.comm.ticket[4 + 8 * (a.comm.flag_value % 32) + 2] = (unsigned)t; }
#endif
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
                EroTileDesc *__restrict__ desc, uint16_t *__restrict__ adj16, int32_t *__restrict__ stats, int want_d3)
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
    for (int k = 0; k < ERO_NSEG; ++k) { d.seg_start[k] = 0; d.seg_len[k] = 0; d.seg_off[k] = 0; }
    d.nseg = 0; d.irregular = 0; d.halo_used = 0; d.d3 = 0;
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
    // ---- where each slot's edge length lives (dist3, see nxb_erosion_plan.cuh) ----
    int d3_need = 0;                    // staged dist3 halo slots this vertex needs (0 = none)
    bool is_heavy = false;
    if (!d.irregular && want_d3) {      // 36 scattered reads per vertex: only when the dist3 sweep is asked for
        bool heavy = false;
        int fcnt = 0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            if (nb[q] < 0) continue;
            if ((int64_t)nb[q] > v) { code[q] |= (uint32_t)(fcnt & 3) << ERO_CODE_I_SHIFT; ++fcnt; }
            else {
                // the neighbour n owns the edge: entry = rank of v among n's forward neighbours
                const int64_t n = nb[q];
                int fn = 0, idx = -1;
                for (int t = 0; t < 6; ++t) {
                    const int32_t m = adj[n * 6 + t];
                    if ((int64_t)m > n) { if ((int64_t)m == v) idx = fn; ++fn; }
                }
                if (fn > 3 || idx < 0) heavy = true;
                else {
                    code[q] |= ERO_CODE_BACK | (uint32_t)idx << ERO_CODE_I_SHIFT;
                    const int pos = (int)(code[q] & ERO_CODE_POS);
                    if (pos >= ERO_TILE) d3_need = max(d3_need, pos - ERO_TILE + 1);
                }
            }
        }
        if (fcnt > 3) heavy = true;
        is_heavy = heavy;
        if (heavy) d3_need = 0;
    }
    // exception row of a heavy vertex = number of heavy vertices before it in the tile
    {
        const unsigned bal = __ballot_sync(0xffffffffu, is_heavy);
        __shared__ int wcount[ERO_TILE / 32];
        if ((c & 31) == 0) wcount[c >> 5] = __popc(bal);
        __syncthreads();
        int before = __popc(bal & ((1u << (c & 31)) - 1u)), total = 0;
        for (int w = 0; w < ERO_TILE / 32; ++w) { if (w < (c >> 5)) before += wcount[w]; total += wcount[w]; }
        __syncthreads();
        if (is_heavy) code[0] |= ERO_CODE_HEAVY | (uint32_t)(before & 3) << ERO_CODE_EXC_SHIFT;
        if (total > ERO_EXC) d3_need = ERO_D3_CAP + 4;      // too many: the tile streams the full table
    }
    d3_need = -block_reduce_min(-d3_need, scratch);
    d3_need = (d3_need + 3) & ~3;
    d.d3 = (!want_d3 || d3_need > ERO_D3_CAP) ? 1 : (d3_need << 8);
    // ---- implicit adjacency: is the staging index of every slot's neighbour c + K_q for all c? ----
    {
        __shared__ int k0[6];
        int kq[6];
        bool ok = !d.irregular && v < n_own && v0 >= ERO_WIN_PAD && v0 + ERO_TILE + ERO_WIN_PAD <= capacity &&
                  ERO_WIN + d.halo_used <= ERO_STAGE_ELEMS;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int64_t n = nb[q];
            if (n < 0) { ok = false; kq[q] = 0; continue; }
            int widx;
            if (n >= v0 - ERO_WIN_PAD && n < v0 + ERO_TILE + ERO_WIN_PAD) widx = (int)(n - v0) + ERO_WIN_PAD;
            else widx = ERO_WIN + ((int)(code[q] & ERO_CODE_POS) - ERO_TILE);
            kq[q] = widx - c;
        }
        if (c == 0) for (int q = 0; q < 6; ++q) k0[q] = kq[q];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 6; ++q) ok = ok && kq[q] == k0[q];
        const int all_ok = __syncthreads_and(ok ? 1 : 0);
        d.affine = all_ok;
        for (int q = 0; q < 6; ++q) d.aff_k[q] = (int16_t)(all_ok ? k0[q] : 0);
        for (int t = 0; t < 8; ++t) d.pad[t] = 0;
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) adj16[v * 6 + q] = (uint16_t)code[q];          // adj16 is allocated in whole tiles
    if (c == 0) {
        desc[tile] = d;
        if (d.affine) atomicAdd(stats + 2, 1);
        if (d.irregular) atomicAdd(stats, 1);
        atomicMax(stats + 1, d.halo_used);
    }
}

// dist3[v][i] = length of the edge to v's i-th larger-numbered neighbour in slot order (rows with
// more than 3 such neighbours -- the mesh skeleton -- are never referenced: see the plan's heavy bit)
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

// exception rows: the full 6 lengths of every heavy vertex, at [tile][row from the vertex's code]
__global__ void __launch_bounds__(256)
exc_build_kernel(const uint16_t *__restrict__ adj16, const float *__restrict__ dist, int64_t n_own, float *__restrict__ exc)
{
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_own; v += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t c0 = adj16[v * 6];
        if (c0 & ERO_CODE_HEAVY) {
            float *row = exc + (v / ERO_TILE) * (ERO_EXC * 6) + ((c0 >> ERO_CODE_EXC_SHIFT) & 3u) * 6;
#pragma unroll
            for (int q = 0; q < 6; ++q) row[q] = dist[v * 6 + q];
        }
    }
}

NXB_API int64_t nxb_erode_dist3_floats(int64_t n_own)
{
    const int64_t n_tiles = (n_own + ERO_TILE - 1) / ERO_TILE;
    return n_tiles * ERO_TILE * 3 + n_tiles * ERO_EXC * 6;
}

NXB_API int nxb_erode_dist3_build(const void *plan_mem, const int32_t *adj, const float *dist, int64_t n_own, float *dist3, void *stream)
{
    NXB_ARG(n_own >= 0);
    if (n_own == 0) return NXB_OK;
    NXB_ARG(plan_mem && adj && dist && dist3 && (((uintptr_t)dist3) & 15) == 0);
    const int64_t n_tiles = (n_own + ERO_TILE - 1) / ERO_TILE, n_rows = n_tiles * ERO_TILE;
    const uint16_t *adj16 = (const uint16_t *)((const char *)plan_mem + n_tiles * sizeof(EroTileDesc));
    float *exc = dist3 + n_rows * 3;
    NXB_CUDA(cudaMemsetAsync(exc, 0, sizeof(float) * n_tiles * ERO_EXC * 6, (cudaStream_t)stream));
    dist3_build_kernel<<<nxb_grid_resident(dist3_build_kernel, 256, 0, (n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        adj, dist, n_own, n_rows, dist3);
    NXB_LAUNCH_CHECK();
    exc_build_kernel<<<nxb_grid_resident(exc_build_kernel, 256, 0, (n_own + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        adj16, dist, n_own, exc);
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
    NXB_CUDA(cudaMalloc(&stats, 12));
    NXB_CUDA(cudaMemsetAsync(stats, 0, 12, st));
    const char *d3env = getenv("NXB_ERO_DIST3");
    ero_plan_kernel<<<(unsigned)n_tiles, ERO_TILE, 0, st>>>(adj, n_own, capacity, desc, adj16, stats, d3env && atoi(d3env) == 1);
    NXB_LAUNCH_CHECK();
    int32_t h[3] = {0, 0, 0};
    NXB_CUDA(cudaMemcpyAsync(h, stats, 12, cudaMemcpyDeviceToHost, st));
    NXB_CUDA(cudaStreamSynchronize(st));
    NXB_CUDA(cudaFree(stats));
    if (stats_host) { stats_host[0] = (int32_t)n_tiles; stats_host[1] = h[0]; stats_host[2] = h[1]; stats_host[3] = h[2]; }
    return NXB_OK;
}

static bool g_ero_attr_set[64] = {false};

static int erode3_plan_launch(const void *plan_mem, const int32_t *adj, const float *dist, const float *dist3,
                              const float *h_in, const float *w_in, const float *s_in,
                              float *h_out, float *w_out, float *s_out,
                              int64_t n_own, float rain, const EroComm &comm, void *stream)
{
    NXB_ARG(n_own >= 0);
    if (n_own == 0) return NXB_OK;
    NXB_ARG(plan_mem && adj && dist && h_in && w_in && s_in && h_out && w_out && s_out);
    NXB_ARG(h_in != h_out && w_in != w_out && s_in != s_out);
    NXB_ARG((((uintptr_t)plan_mem | (uintptr_t)dist | (uintptr_t)dist3 | (uintptr_t)h_in | (uintptr_t)w_in | (uintptr_t)s_in) & 15) == 0);
    const int64_t n_tiles = (n_own + ERO_TILE - 1) / ERO_TILE;
    EroPlanArgs a;
    a.desc = (const EroTileDesc *)plan_mem;
    a.adj16 = (const uint16_t *)((const char *)plan_mem + n_tiles * sizeof(EroTileDesc));
    a.adj = adj; a.dist = dist; a.dist3 = dist3;
    a.exc = dist3 ? dist3 + n_tiles * ERO_TILE * 3 : nullptr;
    a.h_in = h_in; a.w_in = w_in; a.s_in = s_in;
    a.h_out = h_out; a.w_out = w_out; a.s_out = s_out;
    a.n_own = n_own; a.rain = rain;
    a.comm = comm;
    int dev = 0;
    NXB_CUDA(cudaGetDevice(&dev));
    static int cfg_stages = 0;
    if (cfg_stages == 0) {
        const char *e = getenv("NXB_ERO_STAGES");
        cfg_stages = e ? atoi(e) : 3;
        if (cfg_stages < 2 || cfg_stages > ERO_STAGES_MAX) cfg_stages = 3;
    }
    a.n_stages = cfg_stages;
    { const char *e = getenv("NXB_ERO_AFFINE"); a.use_affine = e ? atoi(e) : 1; }      // read per launch: tests toggle it
    const size_t smem = sizeof(EroStage) * cfg_stages;
    if (dev < 64 && !g_ero_attr_set[dev]) {
        NXB_CUDA(cudaFuncSetAttribute(erode3_plan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NXB_CUDA(cudaFuncSetAttribute(erode3_plan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_ero_attr_set[dev] = true;
    }
    const bool use_comm = comm.n_send_peers > 0 || comm.n_wait > 0 || comm.tile_order != nullptr || comm.ticket != nullptr;
    if (use_comm) {
        int grid = nxb_grid_resident(erode3_plan_kernel<true>, ERO_THREADS, smem, n_tiles);
        erode3_plan_kernel<true><<<grid, ERO_THREADS, smem, (cudaStream_t)stream>>>(a);
    } else {
        int grid = nxb_grid_resident(erode3_plan_kernel<false>, ERO_THREADS, smem, n_tiles);
        erode3_plan_kernel<false><<<grid, ERO_THREADS, smem, (cudaStream_t)stream>>>(a);
    }
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

NXB_API int nxb_erode3_plan_step_f32(const void *plan_mem, const int32_t *adj, const float *dist, const float *dist3,
                                     const float *h_in, const float *w_in, const float *s_in,
                                     float *h_out, float *w_out, float *s_out,
                                     int64_t n_own, float rain, void *stream)
{
    EroComm comm;
    memset(&comm, 0, sizeof comm);
    return erode3_plan_launch(plan_mem, adj, dist, dist3, h_in, w_in, s_in, h_out, w_out, s_out, n_own, rain, comm, stream);
}

// Sweep + halo exchange in ONE kernel (see EroComm).  send_ptr / send_entries: device CSR of
// {int32 dst, uint16 vertex-in-tile, uint16 peer slot}; peer_h / peer_w / peer_flag: host arrays of
// n_send_peers NVLink-mapped pointers (the peers' OUTPUT buffers of this sweep and their flag slot
// for this rank); flags: this rank's flag array; wait_rank: host int32[n_wait] source ranks whose
// flag must reach wait_target before halo slots are read; flag_value: raised in the peers when the
// whole grid has finished; halo_begin: first halo slot; ticket: device uint32, zero.
NXB_API int nxb_erode3_plan_step_comm_f32(const void *plan_mem, const int32_t *adj, const float *dist, const float *dist3,
                                          const float *h_in, const float *w_in, const float *s_in,
                                          float *h_out, float *w_out, float *s_out,
                                          int64_t n_own, float rain,
                                          const int32_t *send_ptr, const void *send_entries, int n_send_peers,
                                          void *const *peer_h, void *const *peer_w, void *const *peer_flag,
                                          const void *flags, const int32_t *wait_rank, int n_wait,
                                          uint32_t wait_target, uint32_t flag_value, int64_t halo_begin,
                                          void *ticket, const int32_t *tile_order, int64_t n_early, void *stream)
{
    NXB_ARG(n_early >= 0 && n_early < (1ll << 31) && (n_early == 0 || (tile_order && n_send_peers > 0)));
    NXB_ARG(n_send_peers >= 0 && n_send_peers <= ERO_MAX_PEERS && n_wait >= 0 && n_wait <= ERO_MAX_PEERS);
    NXB_ARG(n_send_peers == 0 || (send_ptr && send_entries && peer_h && peer_w && peer_flag && ticket));
    NXB_ARG(n_wait == 0 || (flags && wait_rank));
    EroComm comm;
    memset(&comm, 0, sizeof comm);
    if (n_send_peers > 0) {
        comm.send_ptr = send_ptr; comm.send_entries = (const EroSendEntry *)send_entries;
        for (int p = 0; p < n_send_peers; ++p) {
            comm.peer_h[p] = (float *)peer_h[p]; comm.peer_w[p] = (float *)peer_w[p]; comm.peer_flag[p] = (uint32_t *)peer_flag[p];
        }
        comm.n_send_peers = n_send_peers;
        comm.ticket = (unsigned int *)ticket;
    }
    comm.flags = (const uint32_t *)flags;
    for (int p = 0; p < n_wait; ++p) comm.wait_rank[p] = wait_rank[p];
    comm.n_wait = n_wait;
    comm.wait_target = wait_target; comm.flag_value = flag_value; comm.halo_begin = halo_begin;
    comm.tile_order = tile_order;
    comm.n_early = (int)n_early;
    return erode3_plan_launch(plan_mem, adj, dist, dist3, h_in, w_in, s_in, h_out, w_out, s_out, n_own, rain, comm, stream);
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

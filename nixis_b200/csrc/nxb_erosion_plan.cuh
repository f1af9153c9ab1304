// Tile plan of the erosion stencil (shared between the plan builder and the sweep kernels).
//
// The vertex range [0, n_own) is cut into tiles of ERO_TILE consecutive vertices.  In meshzoo
// order the neighbours of a tile are the tile itself plus a few CONTIGUOUS index runs (the mesh
// rows above and below, the element before / after the tile; for a multi-GPU shard also runs of
// halo slots).  The plan stores, per tile,
//   * up to ERO_NSEG halo segments (start, length), 4-element aligned so that a segment of a
//     float array is a legal cp.async.bulk source (16-byte aligned, multiple of 16 bytes);
//   * the adjacency re-encoded as uint16 codes into the tile's staging buffer:
//       code <  ERO_TILE          : own tile, vertex v0 + code
//       code >= ERO_TILE          : halo slot code - ERO_TILE of the staging buffer
//     (-1 pads of valence-5 vertices point at the vertex itself, which contributes nothing).
// A tile whose neighbours do not fit (more than ERO_NSEG segments or more than ERO_HALO_CAP halo
// slots) is flagged irregular and processed with global gathers through the int32 table.
//
// Three tile kinds, chosen per tile by the plan builder from the actual neighbour table (any vertex
// order qualifies or not on its own merits):
//   ERO_KIND_CODES    explicit 16-bit codes + the full float[6] edge-length rows (60 B / vertex-sweep)
//   ERO_KIND_AFFINE   IMPLICIT ADJACENCY.  Away from the mesh skeleton and from row ends the
//       neighbours of vertex v0 + c sit at FIXED distances: c - 1, c + 1, and four positions in the
//       rows above / below.  The staging index of slot q's neighbour is then c + K_q with six
//       per-tile constants K_q, provided the tile's own values are staged as the 264-element window
//       [v0 - 4, v0 + 260) (vertex c at window index c + 4, so the elements just before / after the
//       tile are ordinary window entries) followed by the halo runs at index ERO_WIN.  No per-vertex
//       adjacency is read (48 B / vertex-sweep) and the per-vertex code unpacking disappears.
//   ERO_KIND_AFFINE3  affine AND ONE STORED LENGTH PER EDGE.  The length of edge {a, b}, a < b,
//       lives in the row of a: dist3[a][i], i = rank of b among a's larger-numbered ("forward")
//       neighbours in slot order.  In the mesh interior every vertex has exactly 3 forward
//       neighbours (next in its row, two in the next row), so dist3 is float[.][3] -- half of the
//       6-per-vertex table.  On an affine tile the owner row of slot q is at a per-tile constant
//       INDEX distance G_q from the vertex (0 for a forward slot, the neighbour's index minus the
//       vertex's for a backward slot) and the entry is a per-tile constant too, so slot q's length
//       of vertex v is dist3[3 v + D_q] with D_q = 3 G_q + entry: six coalesced loads per thread,
//       issued BEFORE the thread waits for the tile's bulk copies -- they, and the vertex's own
//       sediment, bypass the bulk-copy engine (see nxb_erosion.cu: that engine, not HBM, is what
//       the staged bytes queue on).  36 B / vertex-sweep from HBM; the rows of the previous mesh
//       row were just streamed by another tile and come from L2.
//   ERO_KIND_TWO      TWO-PIECE kind 3.  A tile that contains the end of a mesh row is not affine: the
//       vertices of the next row see their neighbours at distances that differ by one, and the two
//       row-end vertices (plus the vertex behind them, whose backward edge is stored in a shorter row)
//       neighbour the mesh skeleton.  Such a tile -- 20 % of all tiles at d = 2500, all of the former
//       kind-1 tiles -- is kind 3 with TWO sets of constants (K_q, D_q for c < split and for
//       c >= split + ERO_EXC) and ERO_EXC exception vertices [split, split + ERO_EXC) that go through
//       their explicit 16-bit codes and full length rows like a kind-1 vertex: 36 B / vertex-sweep
//       instead of 60 for all but four vertices of the tile.
#pragma once
#include <stdint.h>

#define ERO_TILE 256
#define ERO_NSEG 8
#define ERO_MAXSEG 320          // longest single segment
#define ERO_GAP 16              // a run of needed indices ends at a longer gap
#define ERO_HALO_CAP 640        // halo slots per tile
#define ERO_CODE_POS 0x03ffu
#define ERO_STAGE_ELEMS (ERO_TILE + ERO_HALO_CAP)

#define ERO_WIN_PAD 4
#define ERO_WIN (ERO_TILE + 2 * ERO_WIN_PAD)     // 264
#define ERO_KIND_CODES 1
#define ERO_KIND_AFFINE 2
#define ERO_KIND_AFFINE3 3
#define ERO_KIND_TWO 4
#define ERO_EXC 4               // kind 4: vertices [split, split + ERO_EXC) of the tile are handled through their explicit codes

struct EroTileDesc {            // 256 bytes
    int32_t seg_start[ERO_NSEG];
    uint16_t seg_len[ERO_NSEG];
    uint16_t seg_off[ERO_NSEG]; // offset of the segment inside the halo area
    int32_t nseg;
    int32_t irregular;
    int32_t halo_used;
    int32_t d3;                 // bit 0: tile qualifies for kind 3 (d3_off valid)
    int32_t affine;             // 1: implicit adjacency, aff_k valid
    int16_t aff_k[6];           // K_q: staging index of slot q's neighbour minus c (window layout)
    int32_t d3_off[4];          // kind 3: D_q = float index of slot q's length in dist3 minus 3 v, q = 0..3
    int32_t send0, send1;       // multi-GPU shard: this tile's range of the send-entry list (filled by the driver)
    int32_t d3_off45[2];        // D_4, D_5
    // ---- second half: kind 4 (two-piece) only
    int32_t two;                // 1: tile qualifies for kind 4; aff_k / d3_off describe piece A (c < split)
    int32_t split;              // first exception vertex
    int16_t aff_kB[6];          // piece B (c >= split + ERO_EXC)
    int32_t d3_offB[6];
    int32_t pad[21];
};
static_assert(sizeof(EroTileDesc) == 256, "EroTileDesc layout");
#define ERO_DESC_WORDS 64
// word indices of the descriptor fields (the producer warp holds one word per lane)
#define ERO_DW_LEN (ERO_NSEG)
#define ERO_DW_OFF (ERO_NSEG + ERO_NSEG / 2)
#define ERO_DW_NSEG (2 * ERO_NSEG)
#define ERO_DW_IRREGULAR (2 * ERO_NSEG + 1)
#define ERO_DW_HALO_USED (2 * ERO_NSEG + 2)
#define ERO_DW_D3 (2 * ERO_NSEG + 3)
#define ERO_DW_AFFINE (2 * ERO_NSEG + 4)
#define ERO_DW_AFFK (2 * ERO_NSEG + 5)      // 3 words
#define ERO_DW_D3OFF (2 * ERO_NSEG + 8)     // 4 words (D_0..D_3); D_4, D_5 at ERO_DW_D3OFF45
#define ERO_DW_D3OFF45 (2 * ERO_NSEG + 14)
#define ERO_DW_SEND (2 * ERO_NSEG + 12)     // 2 words
#define ERO_DW_TWO 32
#define ERO_DW_SPLIT 33
#define ERO_DW_AFFKB 34                     // 3 words
#define ERO_DW_D3OFFB 37                    // 6 words

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
// EDGE LENGTHS ARE STORED ONCE PER EDGE (dist3).  The length of edge {a, b}, a < b, lives in the
// row of a: dist3[a][i], i = rank of b among a's larger-numbered ("forward") neighbours in slot
// order.  In the mesh interior every vertex has exactly 3 forward neighbours (next in its row, two
// in the next row), so dist3 is float[V][3] -- half of the 6-per-vertex table.  The upper bits of
// each 16-bit code say where the slot's length is:
//     bits 10-11  i        entry of the owner's dist3 row
//     bit  12     backward the owner is the NEIGHBOUR (staged: own tile or a halo run with smaller
//                          indices, whose dist3 rows the producer stages as well)
//     bit  15     (slot 0 only) heavy vertex: some edge has an owner with more than 3 forward
//                          neighbours (mesh skeleton).  Its 6 lengths are copied into the tile's
//                          exception rows (ERO_EXC per tile, staged with the tile); bits 13-14 of
//                          slot 0 hold the row.  A tile with more heavy vertices is kind 1.
// A regular tile whose backward halo positions exceed ERO_D3_CAP keeps streaming the full table
// (kind 1).
#pragma once
#include <stdint.h>

#define ERO_TILE 256
#define ERO_NSEG 8
#define ERO_MAXSEG 320          // longest single segment
#define ERO_GAP 16              // a run of needed indices ends at a longer gap
#define ERO_HALO_CAP 640        // halo slots per tile
#define ERO_D3_CAP 296          // leading halo slots whose dist3 rows can be staged
#define ERO_CODE_POS 0x03ffu
#define ERO_CODE_I_SHIFT 10
#define ERO_CODE_BACK 0x1000u
#define ERO_CODE_HEAVY 0x8000u
#define ERO_CODE_EXC_SHIFT 13
#define ERO_EXC 4               // exception rows (6 lengths each) per tile
#define ERO_STAGE_ELEMS (ERO_TILE + ERO_HALO_CAP)

// IMPLICIT ADJACENCY (kind 2).  Away from the mesh skeleton and from row ends the neighbours of
// vertex v0 + c sit at FIXED distances: c - 1, c + 1, and four positions in the rows above / below.
// For such an "affine" tile the staging index of slot q's neighbour is c + K_q with six per-tile
// constants K_q, provided the tile's own values are staged as the 264-element window
// [v0 - 4, v0 + 260) (vertex c at window index c + 4, so the elements just before / after the tile
// are ordinary window entries) followed by the halo runs at index ERO_WIN.  The sweep then needs no
// per-vertex adjacency at all: the 12 B/vertex code stream is not read (48 B per vertex-sweep) and
// the per-vertex code unpacking disappears.  The plan marks a tile affine when all 256 vertices
// are valid, have six neighbours, and agree on every K_q.
#define ERO_WIN_PAD 4
#define ERO_WIN (ERO_TILE + 2 * ERO_WIN_PAD)     // 264
#define ERO_KIND_AFFINE 2

struct EroTileDesc {            // 128 bytes
    int32_t seg_start[ERO_NSEG];
    uint16_t seg_len[ERO_NSEG];
    uint16_t seg_off[ERO_NSEG]; // offset of the segment inside the halo area
    int32_t nseg;
    int32_t irregular;
    int32_t halo_used;
    int32_t d3;                 // bits 0-7: 0 = edge lengths from dist3, 1 = from the full table; bits 8..: staged dist3 halo slots
    int32_t affine;             // 1: implicit adjacency, aff_k valid
    int16_t aff_k[6];           // K_q: staging index of slot q's neighbour minus c (window layout)
    int32_t pad[8];
};
static_assert(sizeof(EroTileDesc) == 128, "EroTileDesc layout");
#define ERO_DESC_WORDS 32
// word indices of the descriptor fields (the producer warp holds one word per lane)
#define ERO_DW_LEN (ERO_NSEG)
#define ERO_DW_OFF (ERO_NSEG + ERO_NSEG / 2)
#define ERO_DW_NSEG (2 * ERO_NSEG)
#define ERO_DW_IRREGULAR (2 * ERO_NSEG + 1)
#define ERO_DW_HALO_USED (2 * ERO_NSEG + 2)
#define ERO_DW_D3 (2 * ERO_NSEG + 3)
#define ERO_DW_AFFINE (2 * ERO_NSEG + 4)
#define ERO_DW_AFFK (2 * ERO_NSEG + 5)

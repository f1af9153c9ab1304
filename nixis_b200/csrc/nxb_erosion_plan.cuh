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
#pragma once
#include <stdint.h>

#define ERO_TILE 256
#define ERO_NSEG 6
#define ERO_MAXSEG 320          // longest single segment
#define ERO_HALO_CAP 768        // halo slots per tile
#define ERO_STAGE_ELEMS (ERO_TILE + ERO_HALO_CAP)

struct EroTileDesc {            // 64 bytes
    int32_t seg_start[ERO_NSEG];
    uint16_t seg_len[ERO_NSEG];
    uint16_t seg_off[ERO_NSEG]; // offset of the segment inside the halo area
    int32_t nseg;
    int32_t irregular;
    int32_t halo_used;
    int32_t pad;
};
static_assert(sizeof(EroTileDesc) == 64, "EroTileDesc layout");

// stream.cuh -- the packed per-chunk byte stream the SpMV kernel consumes (DESIGN.md "data layout").
//
// The reference keeps one global array per format plus ~14 B of per-tile headers spread over five
// arrays (tilespmv_cuda.h:505-512).  The planner re-packs the (bit-exact) Tile_matrix into ONE
// contiguous stream cut into scheduler chunks of <= chunk_bytes, each 16-byte aligned so that a
// warp fetches its whole chunk with a single TMA bulk copy (cp.async.bulk) into shared memory:
//
//   ChunkHeader                         32 B
//   RowRec   [nrows]                     8 B each   block rows (or pieces of long block rows)
//   TileDesc [ntiles]                    8 B each   tile column, format, width, aux
//   SideCnt  [#rows with side][16] u16  32 B each   per-row counts of extracted (COO-tile) nonzeros
//   SideCol  [nside] u32                 pad 8      GLOBAL columns of the extracted nonzeros
//   SideVal  [nside] T                   pad 8
//   payload of every tile, in order, each padded to 8 B:
//     CSR      : rowstart[16] u8 | val[nnz] T (pad 8) | nibbles ceil(nnz/2) (pad 8)
//     ELL/HYB  : val[w][16] T (slot-major, rows padded to 16) | nibbles[w][16] -> w*8 B
//     Dense    : val[16][16] T column-major (padded)
//     DenseRow : val[ndr][16] T row-major (padded); row ids as a 16-bit mask in the descriptor
//     DenseCol : val[ndc][16] T slot-major | 16 column nibbles in one u64
//   pad to 16
// Nibble parity is TILE-LOCAL here (element e of a tile sits in byte e/2, high nibble when e is
// even); the reference's global-position parity (csr2tile.h:973, :982) only exists in Tile_matrix.
// COO tiles are not in the stream as tiles: their nonzeros live in the side part exactly once
// (SURVEY.md A.6 design (i), fused into the block row's epilogue).
#pragma once
#include <cstdint>

namespace tsp
{

struct ChunkHeader // 32 B
{
    uint16_t nrows;
    uint16_t ntiles;
    uint32_t nside;
    uint32_t off_tiledesc;
    uint32_t off_sidecnt;
    uint32_t off_sidecol;
    uint32_t off_sideval;
    uint32_t off_payload;
    uint32_t total_bytes;
};
static_assert(sizeof(ChunkHeader) == 32, "ChunkHeader must be 32 bytes");

// RowRec as uint2: x = dest (block row index, or 0x80000000 | partial-sum slot for a piece of a
// split block row); y = ntiles | rowlen << 16 | flags << 24
constexpr uint32_t ROW_PARTIAL = 0x80000000u;
constexpr uint32_t ROWF_HAS_SIDE = 1u;
// TileDesc as uint2: x = tile column; y = format | width << 8 | aux << 16
//   aux: CSR nnz; DenseRow 16-bit row mask; otherwise 0

__host__ __device__ inline uint32_t pad8(uint32_t b) { return (b + 7u) & ~7u; }
__host__ __device__ inline uint32_t pad16(uint32_t b) { return (b + 15u) & ~15u; }

// payload bytes of one tile (format f, stored per the table above); vs = sizeof(value)
__host__ __device__ inline uint32_t tile_payload_bytes(int f, int nnz, int width, int nd, uint32_t vs)
{
    switch (f)
    {
    case 0: // CSR
        return 16u + pad8((uint32_t)nnz * vs) + pad8(((uint32_t)nnz + 1u) / 2u);
    case 2: // ELL
    case 3: // HYB (ELL part)
        return (uint32_t)width * 16u * vs + (uint32_t)width * 8u;
    case 4: // Dense
        return 256u * vs;
    case 5: // DenseRow
        return (uint32_t)nd * 16u * vs;
    case 6: // DenseCol
        return (uint32_t)nd * 16u * vs + 8u;
    default: // COO: lives in the side part
        return 0u;
    }
}

// total size of a chunk from its counters (must match pack_kernel's layout exactly)
__host__ __device__ inline uint32_t chunk_layout_bytes(uint32_t nrows, uint32_t ntiles, uint32_t nsiderows,
                                                       uint32_t nside, uint32_t payload, uint32_t vs)
{
    return pad16(32u + 8u * nrows + 8u * ntiles + 32u * nsiderows + pad8(4u * nside) + pad8(vs * nside) + payload);
}

// one schedulable unit: a whole block row, or a piece of a long one
struct PlanItem
{
    int br;        // block row
    int t0, t1;    // tile range [t0, t1) in Tile_matrix order (COO tiles inside are skipped)
    int s0, s1;    // side-CSR entry range [s0, s1) (global positions in deferredcoo_*)
    uint32_t dest; // block row, or ROW_PARTIAL | slot
    int rowlen;
};

} // namespace tsp

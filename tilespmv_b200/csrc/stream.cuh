// stream.cuh -- the packed per-chunk byte stream the SpMV kernel consumes (DESIGN.md "data layout").
//
// The reference keeps one global array per format plus ~14 B of per-tile headers spread over five
// arrays (tilespmv_cuda.h:505-512).  The planner re-packs the (bit-exact) Tile_matrix into ONE
// contiguous stream cut into scheduler chunks of <= chunk_bytes, each 16-byte aligned so that a
// warp fetches its whole chunk with a single TMA bulk copy (cp.async.bulk) into shared memory:
//
//   ChunkHeader                          32 B
//   RowRec   [nrows]                     16 B each  block rows (or pieces of long block rows)
//   ODesc    [nother]                     8 B each  descriptors of the non-ELL tiles, pad 16
//   SideHdr  [#rows with side][20] u16   40 B each  17 exclusive row starts of the extracted nnz + mask of long rows, pad 16
//   SideVal  [nside] T                   pad 16
//   payload, row after row:
//     ELL group of the row -- ALL slot-rows (16 values, one per local row) of its ELL/HYB tiles,
//     flattened so the kernel runs one branch-free loop over them:
//         val [nsr][16] T | nibbles [nsr][8] B | xsel [nsr] u8 (x segment of the slot-row), pad 16
//     then the other tiles of the row in order, each padded to 16 B:
//         CSR      : rowstart[16] u8 | val[nnz] T (pad 8) | nibbles ceil(nnz/2) (pad 8)
//         CSR group: ALL CSR tiles of a block row with <= 16 stream tiles merged into one jagged list (stream-only
//                    pseudo format TSP_FMT_CSRGROUP): slot-row s holds the s-th entry of every local row that has
//                    one (counted across the row's CSR tiles in tile order), rows ascending, no padding:
//                    hdr[nsrg] u32 {row mask | offset << 16} (pad 16) | val[n] T (pad 16) | idx[n+1] u8 (pad 16),
//                    idx = (ordinal of the tile among the row's stream tiles) << 4 | local column = offset into the
//                    row's window of staged x.  One branch-free loop per block row instead of one per tile.
//         Dense    : val[16][16] T column-major (rows / columns padded with zeros)
//         DenseRow : val[ndr][16] T row-major (padded); row ids as a 16-bit mask in the descriptor
//         DenseCol : val[ndc][16] T slot-major | 16 column nibbles in one u64
//   x-staging lists of the chunk the SAME WARP processes next (chunk index + nw, nw = warps of the
//   persistent grid): TileCol [next ntiles] u32 pad 16 | SideCol [next nside] u32 pad 16
//     = tile column of every stream tile / GLOBAL column of every extracted nonzero of that chunk.
//     Carrying them one chunk early lets the warp stage the x operand of chunk k+1 while it works
//     on chunk k without chunk k+1 having arrived, so two TMA stages per warp suffice.  The lists
//     of each warp's first chunk live in a small separate array (plan.head).
// FLAT chunks.  A chunk whose block rows hold NO stream tile -- only extracted (side) entries: every chunk of a uniform
// random matrix, of an x-panel sub-plan, most chunks of a power-law graph -- uses a second layout (CHF_FLAT set in
// ChunkHeader::nrows) made for rows with a handful of entries each, where the per-block-row bookkeeping of the general
// path (~450 warp instructions per block row) dwarfs the arithmetic:
//   ChunkHeader | RowRec[nrows <= 16] | len u8 [16 nrows] | FlatLong[nside / FLAT_LONG_ROW] pad 16 | val T[nside] pad 16 | lists
// The 16 nrows local rows are processed 32 at a time ("rounds": lane = row).  Inside a round the entries are stored
// slot-major WITHOUT padding (jagged diagonals): first entry 0 of every row that has one, rows ascending, then entry 1
// of every row that has two, ...; the kernel finds its position with one ballot + popc per slot, all loads of a slot are
// contiguous (conflict-free), every lane sums its own row in input order and the 32 y values leave as two 128-byte
// stores -- no shuffles, no atomics, no per-row headers.  Rows with >= FLAT_LONG_ROW entries (pieces of hub rows) have
// len = 0 there; their entries follow the rounds in row order and are summed by the whole warp (FlatLong records).
// Header fields in a flat chunk: off_sidehdr = offset of len[], off_payload = offset of FlatLong[], off_sideval =
// offset of val[], off_odesc = number of FlatLong records.  The x-staging list of the chunk is in the same order as val.
// Nibble parity is TILE-LOCAL here (element e sits in byte e/2, high nibble when e is even); the
// reference's global-position parity (csr2tile.h:973, :982) only exists in Tile_matrix.
// COO tiles are not in the stream as tiles: their nonzeros live in the side part exactly once
// (SURVEY.md A.6 design (i), fused into the block row's epilogue).
#pragma once
#include <cstdint>

namespace tsp
{

struct ChunkHeader // 32 B, read by the kernel as two 128-bit shared-memory loads
{
    uint16_t nrows;
    uint16_t ntiles;       // stream tiles (ELL + other) of THIS chunk = staged x segments (<= 256)
    uint16_t next_ntiles;  // counts of the lists at off_nextlist
    uint16_t next_nside;
    uint16_t off_nextlist; // byte offsets from the start of the chunk (chunks are < 64 KB)
    uint16_t off_odesc;
    uint16_t off_sidehdr;
    uint16_t off_sideval;
    uint32_t off_payload;
    uint16_t next_flags;   // CHF_* of the next chunk's x staging
    uint16_t nside;        // extracted nonzeros of THIS chunk
    uint32_t issue_off16;  // TMA descriptor of the chunk this warp fetches into the stage this chunk
    uint32_t issue_bytes;  //   frees: chunk index + stages * nw, {byte offset / 16, bytes}; 0 bytes = none
};
static_assert(sizeof(ChunkHeader) == 32, "ChunkHeader must be 32 bytes");
constexpr uint32_t CHF_PARTIAL_X = 1u; // some x segment sticks out past colA (zero-filled staging path)
constexpr uint32_t CHF_FLAT = 0x8000u; // bit of ChunkHeader::nrows: side-only chunk in the flat layout
constexpr int FLAT_MAX_ROWS = 16;      // block rows per flat chunk (8 rounds of 32 local rows)
constexpr int FLAT_LONG_ROW = 32;      // local rows with at least this many entries are summed by the whole warp
struct FlatLong                        // 8 B
{
    uint16_t row;   // local row of the chunk (block-row ordinal * 16 + row)
    uint16_t start; // first entry (index into val[] / the staged x)
    uint16_t count;
    uint16_t pad;
};
constexpr uint32_t CHUNK_OFF_ROWS = 32;

struct RowRec // 16 B, one 128-bit shared-memory load
{
    uint32_t dest;       // block row index, or ROW_PARTIAL | partial-sum slot for a piece of a split row
    uint16_t nsr;        // slot-rows in the row's ELL group
    uint16_t nother;     // non-ELL tiles
    uint8_t rowlen;      // rows that exist (16 except in the last block row)
    uint8_t flags;       // ROWF_*
    uint16_t side_nit;   // trip count of the side loop: max over local rows of ceil(#entries / 4)
    uint16_t ell_bytes16; // bytes of the ELL group / 16
    uint16_t pad0;
};
static_assert(sizeof(RowRec) == 16, "RowRec must be 16 bytes");
constexpr uint32_t ROW_PARTIAL = 0x80000000u;
constexpr uint32_t ROWF_HAS_SIDE = 1u;
constexpr uint32_t SIDEHDR_BYTES = 40; // 17 x u16 row starts + u16 mask of the long rows
constexpr int SIDE_LONG_ROW = 64;      // local rows with at least this many side entries are summed by the whole warp (R-MAT 2^21: 32 -> 198.7 us, 64 -> 194.2, 128 -> 197.8)

// ODesc as uint2: x = format | xsel << 8 | width << 16 ; y = aux (CSR nnz, DenseRow row mask)
constexpr int TSP_FMT_CSRGROUP = 8;      // ODesc: xsel = first x segment of the row, width = slot-rows, aux = entries
constexpr int CSRGROUP_MAX_TILES = 16;   // stream tiles of a block row that may use the group (4-bit tile ordinal)
constexpr int CSRGROUP_MAX_SLOTROWS = 256;

__host__ __device__ inline uint32_t pad8(uint32_t b) { return (b + 7u) & ~7u; }
__host__ __device__ inline uint32_t pad16(uint32_t b) { return (b + 15u) & ~15u; }

__host__ __device__ inline bool fmt_is_ell(int f) { return f == 2 || f == 3; }
__host__ __device__ inline bool fmt_is_other(int f) { return f == 0 || f == 4 || f == 5 || f == 6; }

// payload bytes of one non-ELL stream tile; vs = sizeof(value)
__host__ __device__ inline uint32_t other_payload_bytes(int f, int nnz, int nd, uint32_t vs)
{
    switch (f) // every payload is padded to 16 B so that 128-bit shared-memory loads stay aligned
    {
    case 0: // CSR
        return pad16(16u + pad8((uint32_t)nnz * vs) + pad8(((uint32_t)nnz + 1u) / 2u));
    case 4: // Dense
        return 256u * vs;
    case 5: // DenseRow
        return (uint32_t)nd * 16u * vs;
    case 6: // DenseCol
        return pad16((uint32_t)nd * 16u * vs + 8u);
    default:
        return 0u;
    }
}
// bytes of a CSR group with nsrg slot-rows and n entries
__host__ __device__ inline uint32_t csr_group_bytes(uint32_t nsrg, uint32_t n, uint32_t vs)
{
    return pad16(4u * nsrg) + pad16(n * vs) + pad16(n + 1u);
}
// bytes of a row's ELL group with nsr slot-rows
__host__ __device__ inline uint32_t ell_group_bytes(uint32_t nsr, uint32_t vs)
{
    return nsr * 16u * vs + pad16(nsr * 9u); // values | 8 B of nibbles + 1 B xsel per slot-row
}

// bytes of the x-staging lists of a chunk with ntiles stream tiles and nside extracted nonzeros
__host__ __device__ inline uint32_t list_bytes(uint32_t ntiles, uint32_t nside)
{
    return pad16(4u * ntiles) + pad16(4u * nside);
}
// size of a chunk WITHOUT the trailing lists, from its counters (must match pack_kernel's layout
// exactly); `payload` = sum over rows of ell_group_bytes + other payloads
__host__ __device__ inline uint32_t chunk_main_bytes(uint32_t nrows, uint32_t nother, uint32_t nsiderows,
                                                     uint32_t nside, uint32_t payload, uint32_t vs)
{
    return CHUNK_OFF_ROWS + 16u * nrows + pad16(8u * nother) + pad16(SIDEHDR_BYTES * nsiderows) + pad16(vs * nside) +
           payload; // payload parts are multiples of 16
}
// the same for a flat chunk (nside extracted nonzeros in nrows block rows, no stream tile)
__host__ __device__ inline uint32_t flat_chunk_main_bytes(uint32_t nrows, uint32_t nside, uint32_t vs)
{
    return CHUNK_OFF_ROWS + 16u * nrows + 16u * nrows + pad16(8u * (nside / (uint32_t)FLAT_LONG_ROW)) + pad16(vs * nside);
}
// head record of a warp's first chunk (plan.head): 16-byte header {ntiles | nside << 16, flags} + lists
constexpr uint32_t HEAD_HDR_BYTES = 16;

// one schedulable unit: a whole block row, or a piece of a long one
struct PlanItem
{
    int br;        // block row
    int t0, t1;    // tile range [t0, t1) in Tile_matrix order (COO tiles inside are skipped)
    int s0, s1;    // side-CSR entry range [s0, s1) (global positions in deferredcoo_*)
    uint32_t dest; // block row, or ROW_PARTIAL | slot
    int rowlen;
    int g_nsrg = 0, g_n = 0; // CSR group of the item: slot-rows / entries (g_n = 0: CSR tiles stay individual tiles)
};

} // namespace tsp

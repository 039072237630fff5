// plan.cu -- builds the SpMV plan from a device-resident Tile_matrix:
//   1. per-tile stream cost + prefix sums (device)
//   2. byte-balanced chunking of block rows, long block rows cut into pieces (host, O(tilem))
//   3. packing of every chunk into the 16-byte-aligned stream (device, one CTA per chunk)
// This replaces what the reference does with a throw-away SpMV launch (stir_spmv_cuda_kernel_v5
// recording per-warp COO replay lists, tilespmv_cuda.h:5-392, :1042-1056) and with the <=4-tile
// warp chunks of tilespmv_cpu.h:68-118: chunks here are bounded by BYTES, not tile counts.
#include <algorithm>

#include "plan.cuh"
#include "primitives.cuh"

namespace tsp
{

constexpr int PL_THREADS = 256;
constexpr double SPMV_SMEM_FIXED = 1024.0; // barriers + zero block of the SpMV kernel (spmv.cu), rounded up
constexpr int PACK_THREADS = 128;

// ---------------------------------------------------------------------------------------------
// 1. per-tile costs (prefix sums over tiles in Tile_matrix order)
// ---------------------------------------------------------------------------------------------
enum TileCostKind
{
    TC_STREAM_TILES, // 1 for every tile that lives in the stream (everything but COO)
    TC_OTHER_TILES,  // 1 for CSR / Dense / DenseRow / DenseCol
    TC_SLOTROWS,     // ELL / HYB width
    TC_OTHER_BYTES,  // payload bytes of the non-ELL tiles
    TC_NONCSR_TILES, // 1 for Dense / DenseRow / DenseCol
    TC_NONCSR_BYTES, // payload bytes of Dense / DenseRow / DenseCol tiles
    TC_HYB_IDXBYTES  // bytes of a HYB tile in hybIdx: ceil(w * rowlen / 2) nibble bytes + one byte per spilled entry
};
struct TileCostIn
{
    const char *fmt;
    const int *tile_nnz;
    const char *width;
    const int *dnsrowptr, *dnscolptr;
    int T;
    uint32_t vs;
    int kind;
    const int *blknnz = nullptr; // TC_HYB_IDXBYTES only
    int last_row_tile0 = 0, last_rowlen = TS;
    __device__ __forceinline__ int operator()(size_t i) const
    {
        if (i >= (size_t)T)
            return 0;
        const int f = fmt[i];
        switch (kind)
        {
        case TC_STREAM_TILES:
            return f != TILESPMV_FMT_COO ? 1 : 0;
        case TC_OTHER_TILES:
            return fmt_is_other(f) ? 1 : 0;
        case TC_NONCSR_TILES:
            return fmt_is_other(f) && f != TILESPMV_FMT_CSR ? 1 : 0;
        case TC_SLOTROWS:
            return fmt_is_ell(f) ? (int)(unsigned char)width[i] : 0;
        case TC_HYB_IDXBYTES:
        {
            if (f != TILESPMV_FMT_HYB)
                return 0;
            const int rowlen = (int)i >= last_row_tile0 ? last_rowlen : TS;
            const int ell = (int)(unsigned char)width[i] * rowlen;
            return (ell + 1) / 2 + (blknnz[i + 1] - blknnz[i] - ell);
        }
        default:
        {
            if (!fmt_is_other(f) || (kind == TC_NONCSR_BYTES && f == TILESPMV_FMT_CSR))
                return 0;
            const int nd = f == TILESPMV_FMT_DENSEROW ? dnsrowptr[i + 1] - dnsrowptr[i]
                                                      : (f == TILESPMV_FMT_DENSECOL ? dnscolptr[i + 1] - dnscolptr[i] : 0);
            return (int)other_payload_bytes(f, tile_nnz[i + 1] - tile_nnz[i], nd, vs);
        }
        }
    }
};

// per-format cost profiling (cf. DEBUG_FORMATCOST / formatprofile, tilespmv_cuda.h:102-111, main.cu:12): a plan restricted
// to some formats sees every other tile as a COO tile, i.e. as not part of the stream
__global__ void __launch_bounds__(PL_THREADS) filter_format_kernel(int T, const char *__restrict__ fmt, unsigned mask, char *__restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T)
        out[t] = ((mask >> (unsigned)fmt[t]) & 1u) ? fmt[t] : (char)TILESPMV_FMT_COO;
}

// the per-tile prefix sums of every block row that has to be cut into pieces, gathered into compact arrays so that the
// host-side piece cutter needs ONE device-to-host copy (one small copy per hub row cost 1.5 s on R-MAT 2^24)
__global__ void __launch_bounds__(PL_THREADS)
    gather_long_rows_kernel(int nlong, const int *__restrict__ row_ta, const long long *__restrict__ row_off, const long long *__restrict__ ob,
                            const int *__restrict__ nc, const int *__restrict__ oc, const int *__restrict__ ws, long long *__restrict__ c_ob,
                            int *__restrict__ c_nc, int *__restrict__ c_oc, int *__restrict__ c_ws)
{
    const int i = blockIdx.x;
    if (i >= nlong)
        return;
    const int ta = row_ta[i];
    const long long o = row_off[i], cnt = row_off[i + 1] - o;
    for (long long k = threadIdx.x; k < cnt; k += blockDim.x)
    {
        c_ob[o + k] = ob[ta + k];
        c_nc[o + k] = nc[ta + k];
        c_oc[o + k] = oc[ta + k];
        c_ws[o + k] = ws[ta + k];
    }
}

struct TileScans // all T+1 entries
{
    const int *nc;        // stream tiles
    const int *oc;        // other tiles
    const int *ws;        // slot-rows
    const long long *ob;  // other payload bytes
    const int *oc2;       // other tiles that are not CSR
    const long long *ob2; // their payload bytes
};

__global__ void __launch_bounds__(PL_THREADS)
    row_summary_kernel(int tilem, int rowA, const int *__restrict__ tile_ptr, TileScans sc,
                       const int *__restrict__ side_ptr, int *__restrict__ row_nt, int *__restrict__ row_no,
                       int *__restrict__ row_nsr, long long *__restrict__ row_ob, int *__restrict__ row_s0,
                       int *__restrict__ row_no2, long long *__restrict__ row_ob2)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > tilem)
        return;
    int r0 = b * TS < rowA ? b * TS : rowA;
    row_s0[b] = side_ptr[r0]; // side entries of block row b are [row_s0[b], row_s0[b+1])
    if (b < tilem)
    {
        int t0 = tile_ptr[b], t1 = tile_ptr[b + 1];
        row_nt[b] = sc.nc[t1] - sc.nc[t0];
        row_no[b] = sc.oc[t1] - sc.oc[t0];
        row_nsr[b] = sc.ws[t1] - sc.ws[t0];
        row_ob[b] = sc.ob[t1] - sc.ob[t0];
        row_no2[b] = sc.oc2[t1] - sc.oc2[t0];
        row_ob2[b] = sc.ob2[t1] - sc.ob2[t0];
    }
}

// CSR tiles of every block row that is short enough for a CSR group (<= CSRGROUP_MAX_TILES stream tiles): number
// of CSR tiles, their nonzeros, and the number of slot-rows = the longest local row summed over those tiles.
__global__ void __launch_bounds__(PL_THREADS)
    row_csr_summary_kernel(int tilem, int rowA, const int *__restrict__ tile_ptr, const char *__restrict__ Format,
                           const int *__restrict__ tile_nnz, const int *__restrict__ csrptr_offset,
                           const unsigned char *__restrict__ Blockcsr_Ptr, const int *__restrict__ row_nt,
                           int *__restrict__ row_csr_cnt, int *__restrict__ row_csr_nnz, int *__restrict__ row_csr_nsrg)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= tilem)
        return;
    int cnt = 0, nnz_sum = 0, nsrg = 0;
    if (row_nt[b] <= CSRGROUP_MAX_TILES)
    {
        const int rowlen = b == tilem - 1 ? rowA - (tilem - 1) * TS : TS;
        int L[TS];
#pragma unroll
        for (int r = 0; r < TS; r++)
            L[r] = 0;
        for (int t = tile_ptr[b]; t < tile_ptr[b + 1]; t++)
        {
            if (Format[t] != TILESPMV_FMT_CSR)
                continue;
            const int nnz = tile_nnz[t + 1] - tile_nnz[t], po = csrptr_offset[t];
            cnt++;
            nnz_sum += nnz;
#pragma unroll
            for (int r = 0; r < TS; r++)
                if (r < rowlen)
                    L[r] += (r + 1 < rowlen ? (int)Blockcsr_Ptr[po + r + 1] : nnz) - (int)Blockcsr_Ptr[po + r];
        }
#pragma unroll
        for (int r = 0; r < TS; r++)
            nsrg = max(nsrg, L[r]);
    }
    row_csr_cnt[b] = cnt;
    row_csr_nnz[b] = nnz_sum;
    row_csr_nsrg[b] = nsrg;
}

// ---------------------------------------------------------------------------------------------
// 3. packing
// ---------------------------------------------------------------------------------------------
template <class T>
struct PackArgs
{
    // plan tables
    const PlanItem *items;
    const long long *chunk_item0; // [nchunks+1]
    const unsigned long long *chunk_off;
    TileScans sc;
    unsigned char *stream;
    unsigned char *head; // lists of every warp's first chunk, head_stride bytes apart
    long long nw;        // lookahead distance = warps of the persistent grid
    int stages;          // TMA pipeline depth the stream is laid out for
    int head_stride;
    int side_long_row;   // local rows with at least this many side entries are summed by the whole warp
    int allow_flat;      // side-only chunks are packed in the flat layout (stream.cuh)
    int *error_flag;
    // Tile_matrix (device)
    int rowA, colA, tilem, tilen;
    const int *tile_columnidx, *tile_nnz;
    const char *Format, *tilewidth;
    const int *csr_offset, *csrptr_offset, *ell_offset, *hyb_offset, *dns_offset, *dnsrow_offset, *dnscol_offset,
        *dnsrowptr, *dnscolptr;
    const int *hyb_idxoff; // byte offset of every tile in hybIdx (0 for non-HYB tiles' own bytes)
    const T *Blockhyb_Val;
    const unsigned char *hybIdx;
    const T *Blockcsr_Val, *Blockell_Val, *Blockdense_Val, *Blockdenserow_Val, *Blockdensecol_Val;
    const unsigned char *Blockcsr_Ptr, *csr_compressedIdx, *ell_compressedIdx;
    const char *denserowid, *densecolid;
    const int *side_ptr, *side_col;
    const T *side_val;
};

__device__ __forceinline__ unsigned nib_global(const unsigned char *packed, int pos)
{
    unsigned b = packed[pos >> 1];
    return (pos & 1) ? (b & 15u) : (b >> 4);
}

// ELL / HYB tile -> w slot-rows of the row's ELL group (runs in one thread).  A HYB tile (format 3,
// csr2tile.h:505-548) contributes its ELL part (w x rowlen slot-major values in Blockhyb_Val, nibbles packed
// per tile in hybIdx, csr2tile.h:984-1008); its spilled entries are already in the side CSR (new_coocount).
template <class T>
__device__ void pack_ell_tile(const PackArgs<T> &a, int t, int br, unsigned xsel, T *vals, unsigned char *idx,
                              unsigned char *xs)
{
    const int rowlen = br == a.tilem - 1 ? a.rowA - (a.tilem - 1) * TS : TS;
    const bool hyb = a.Format[t] == TILESPMV_FMT_HYB;
    const int o = hyb ? a.hyb_offset[t] : a.ell_offset[t];
    const T *src = hyb ? a.Blockhyb_Val : a.Blockell_Val;
    const unsigned char *nib = hyb ? a.hybIdx + a.hyb_idxoff[t] : a.ell_compressedIdx;
    const int nib0 = hyb ? 0 : o; // hybIdx nibbles are tile-local, ell_compressedIdx positions global
    const int w = (int)(unsigned char)a.tilewidth[t];
    for (int s = 0; s < w; s++)
    {
        xs[s] = (unsigned char)xsel;
        for (int r = 0; r < TS; r += 2)
        {
            unsigned n0 = 0, n1 = 0;
            T v0 = 0, v1 = 0;
            if (r < rowlen)
            {
                v0 = src[o + s * rowlen + r];
                n0 = nib_global(nib, nib0 + s * rowlen + r);
            }
            if (r + 1 < rowlen)
            {
                v1 = src[o + s * rowlen + r + 1];
                n1 = nib_global(nib, nib0 + s * rowlen + r + 1);
            }
            vals[s * 16 + r] = v0;
            vals[s * 16 + r + 1] = v1;
            idx[(s * 16 + r) >> 1] = (unsigned char)((n0 << 4) | n1);
        }
    }
}

// non-ELL tile: descriptor + payload (runs in one thread; payloads are <= 2 KB)
template <class T>
__device__ void pack_other_tile(const PackArgs<T> &a, int t, int br, unsigned xsel, unsigned char *desc_out,
                                unsigned char *pay)
{
    const int f = a.Format[t];
    const int tc = a.tile_columnidx[t];
    const int nnz = a.tile_nnz[t + 1] - a.tile_nnz[t];
    const int rowlen = br == a.tilem - 1 ? a.rowA - (a.tilem - 1) * TS : TS;
    const int collen = tc == a.tilen - 1 ? a.colA - (a.tilen - 1) * TS : TS;
    uint32_t w = 0, aux = 0;
    T *pv = reinterpret_cast<T *>(pay);
    switch (f)
    {
    case TILESPMV_FMT_CSR:
    {
        const int o = a.csr_offset[t], po = a.csrptr_offset[t];
        int maxlen = 0;
        for (int r = 0; r < TS; r++)
        {
            pay[r] = r < rowlen ? a.Blockcsr_Ptr[po + r] : (unsigned char)nnz;
            if (r < rowlen)
            {
                const int end = r + 1 < rowlen ? (int)a.Blockcsr_Ptr[po + r + 1] : nnz;
                maxlen = max(maxlen, end - (int)a.Blockcsr_Ptr[po + r]);
            }
        }
        w = (uint32_t)((maxlen + 3) >> 2); // trip count of the kernel's 4-lanes-per-row loop
        T *v = reinterpret_cast<T *>(pay + 16);
        unsigned char *ix = pay + 16 + pad8((uint32_t)nnz * (uint32_t)sizeof(T));
        for (int k = 0; k < nnz; k++)
            v[k] = a.Blockcsr_Val[o + k];
        for (int k = 0; k < nnz; k += 2)
        {
            unsigned hi = nib_global(a.csr_compressedIdx, o + k);
            unsigned lo = k + 1 < nnz ? nib_global(a.csr_compressedIdx, o + k + 1) : 0u;
            ix[k >> 1] = (unsigned char)((hi << 4) | lo);
        }
        aux = (uint32_t)nnz;
        break;
    }
    case TILESPMV_FMT_DENSE:
    {
        const int o = a.dns_offset[t];
        for (int c = 0; c < TS; c++)
            for (int r = 0; r < TS; r++)
                pv[c * 16 + r] = (c < collen && r < rowlen) ? a.Blockdense_Val[o + c * rowlen + r] : (T)0;
        break;
    }
    case TILESPMV_FMT_DENSEROW:
    {
        const int o = a.dnsrow_offset[t], ro = a.dnsrowptr[t];
        const int ndr = a.dnsrowptr[t + 1] - ro;
        w = (uint32_t)ndr;
        for (int i = 0; i < ndr; i++)
        {
            aux |= 1u << ((unsigned)a.denserowid[ro + i] & 15u);
            for (int c = 0; c < TS; c++)
                pv[i * 16 + c] = c < collen ? a.Blockdenserow_Val[o + i * collen + c] : (T)0;
        }
        break;
    }
    case TILESPMV_FMT_DENSECOL:
    {
        const int o = a.dnscol_offset[t], co = a.dnscolptr[t];
        const int ndc = a.dnscolptr[t + 1] - co;
        w = (uint32_t)ndc;
        unsigned long long ids = 0;
        for (int k = 0; k < ndc; k++)
        {
            ids |= (unsigned long long)((unsigned)a.densecolid[co + k] & 15u) << (4 * k);
            for (int r = 0; r < TS; r++)
                pv[k * 16 + r] = r < rowlen ? a.Blockdensecol_Val[o + k * rowlen + r] : (T)0;
        }
        *reinterpret_cast<unsigned long long *>(pay + (size_t)ndc * 16 * sizeof(T)) = ids;
        break;
    }
    default:
        break;
    }
    uint2 d;
    d.x = (uint32_t)f | (xsel << 8) | (w << 16);
    d.y = aux;
    *reinterpret_cast<uint2 *>(desc_out) = d;
}

// CSR group of one item (runs on the whole CTA): slot-row s holds the s-th entry of every local row that has one,
// entries of a row counted across the item's CSR tiles in tile order (layout: stream.cuh)
template <class T>
__device__ void pack_csr_group(const PackArgs<T> &a, const PlanItem &it, unsigned xsel_base, unsigned char *desc_out,
                               unsigned char *pay, int (*g_start)[CSRGROUP_MAX_TILES], int *g_len, unsigned *g_hdr)
{
    const int rowlen = it.rowlen;
    const int nsrg = it.g_nsrg, n = it.g_n;
    __syncthreads();
    if (threadIdx.x < TS) // thread r: where local row r of every CSR tile starts in the row's merged list
    {
        const int r = threadIdx.x;
        int L = 0, ord = 0;
        for (int t = it.t0; t < it.t1; t++)
        {
            if (a.Format[t] != TILESPMV_FMT_CSR)
                continue;
            const int nnz = a.tile_nnz[t + 1] - a.tile_nnz[t], po = a.csrptr_offset[t];
            g_start[r][ord++] = L;
            if (r < rowlen)
                L += (r + 1 < rowlen ? (int)a.Blockcsr_Ptr[po + r + 1] : nnz) - (int)a.Blockcsr_Ptr[po + r];
        }
        g_len[r] = L;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned off = 0;
        for (int s = 0; s < nsrg; s++)
        {
            unsigned mask = 0;
            for (int r = 0; r < TS; r++)
                mask |= g_len[r] > s ? 1u << r : 0u;
            g_hdr[s] = mask | (off << 16);
            off += (unsigned)__popc(mask);
        }
        if (off != (unsigned)n || nsrg > CSRGROUP_MAX_SLOTROWS || n > 0xffff)
            atomicExch(a.error_flag, 1);
        // leading slot-rows in which all 16 rows have an entry: the kernel runs them ELL-style (128-bit value
        // loads, no mask arithmetic); entry (s, r) of that part sits at position 16 s + r
        unsigned nfull = 0;
        while (nfull < (unsigned)nsrg && (g_hdr[nfull] & 0xffffu) == 0xffffu)
            nfull++;
        uint2 d;
        d.x = (uint32_t)TSP_FMT_CSRGROUP | (xsel_base << 8) | ((uint32_t)nsrg << 16);
        d.y = (uint32_t)n | (nfull << 16);
        *reinterpret_cast<uint2 *>(desc_out) = d;
    }
    __syncthreads();
    uint32_t *hdr_out = reinterpret_cast<uint32_t *>(pay);
    for (int s = threadIdx.x; s < nsrg; s += PACK_THREADS)
        hdr_out[s] = g_hdr[s];
    T *val_out = reinterpret_cast<T *>(pay + pad16(4u * (uint32_t)nsrg));
    unsigned char *idx_out = pay + pad16(4u * (uint32_t)nsrg) + pad16((uint32_t)n * (uint32_t)sizeof(T));
    int ord = 0;
    for (int t = it.t0; t < it.t1; t++)
    {
        if (a.Format[t] != TILESPMV_FMT_CSR)
            continue;
        const int nnz = a.tile_nnz[t + 1] - a.tile_nnz[t], po = a.csrptr_offset[t], o = a.csr_offset[t];
        const unsigned tsel = (unsigned)(a.sc.nc[t] - a.sc.nc[it.t0]); // ordinal among the item's stream tiles (< 16)
        for (int k = threadIdx.x; k < nnz; k += PACK_THREADS)
        {
            int r = 0; // last row whose start is <= k
            for (int q = 1; q < rowlen; q++)
                if ((int)a.Blockcsr_Ptr[po + q] <= k)
                    r = q;
            const int s = g_start[r][ord] + (k - (int)a.Blockcsr_Ptr[po + r]);
            const unsigned h = g_hdr[s];
            const unsigned pos = (h >> 16) + (unsigned)__popc(h & 0xffffu & ((1u << r) - 1u));
            val_out[pos] = a.Blockcsr_Val[o + k];
            idx_out[pos] = (unsigned char)((tsel << 4) | nib_global(a.csr_compressedIdx, o + k));
        }
        ord++;
    }
}

// Order of the extracted nonzeros inside a FLAT chunk (stream.cuh): 32 local rows per round, inside a round slot-major
// without padding, rows with >= FLAT_LONG_ROW entries after all rounds in row order.  Runs on ONE warp; calls
// f(position in the chunk, index into the side arrays) for every entry, fills len[] / FlatLong[] when given.  The SpMV
// kernel (process_chunk_flat) walks the rounds with the same ballot / popc arithmetic.
template <class T, class F>
__device__ void flat_walk(const PackArgs<T> &a, long long c, int lane, F f, unsigned char *lens, FlatLong *longtab, int *nlong_out)
{
    const long long j0 = a.chunk_item0[c];
    const int nrows16 = (int)(a.chunk_item0[c + 1] - j0) * TS;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t pos = 0;
    int nlong = 0;
    for (int pass = 0; pass < 2; pass++)
        for (int rd = 0; rd * 32 < nrows16; rd++)
        {
            const int rho = rd * 32 + lane;
            int len = 0, src = 0;
            if (rho < nrows16)
            {
                const PlanItem it = a.items[j0 + (rho >> 4)];
                const int r = rho & 15;
                if (r < it.rowlen)
                {
                    int lo = a.side_ptr[it.br * TS + r], hi = a.side_ptr[it.br * TS + r + 1];
                    lo = min(max(lo, it.s0), it.s1);
                    hi = min(max(hi, it.s0), it.s1);
                    len = hi - lo;
                    src = lo;
                }
            }
            const bool is_long = len >= FLAT_LONG_ROW;
            if (pass == 0)
            {
                if (lens && rho < nrows16)
                    lens[rho] = (unsigned char)(is_long ? 0 : len);
                const int jl = is_long ? 0 : len;
                for (int j = 0;; j++)
                {
                    const unsigned mask = __ballot_sync(0xffffffffu, jl > j);
                    if (!mask)
                        break;
                    if (jl > j)
                        f(pos + (uint32_t)__popc(mask & lt), src + j);
                    pos += (uint32_t)__popc(mask);
                }
            }
            else
            {
                unsigned lm = __ballot_sync(0xffffffffu, is_long);
                while (lm)
                {
                    const int l = __ffs((int)lm) - 1;
                    lm &= lm - 1u;
                    const int L = __shfl_sync(0xffffffffu, len, l), S = __shfl_sync(0xffffffffu, src, l);
                    for (int q = lane; q < L; q += 32)
                        f(pos + (uint32_t)q, S + q);
                    if (lane == 0 && longtab)
                    {
                        FlatLong fl;
                        fl.row = (uint16_t)(rd * 32 + l);
                        fl.start = (uint16_t)pos;
                        fl.count = (uint16_t)L;
                        fl.pad = 0;
                        longtab[nlong] = fl;
                    }
                    nlong++;
                    pos += (uint32_t)L;
                }
            }
        }
    if (nlong_out)
        *nlong_out = nlong;
}

constexpr int PACK_ITEM_INTS = 6; // per item: tile base, other base, payload base, side base, sidehdr idx, nsr

// x-staging lists of chunk cn (tile column of every stream tile, global column of every extracted
// nonzero, in the order the chunk's xsel / side indices use); runs on the whole CTA.  Returns this
// thread's share of the CHF_* flags.
template <class T>
__device__ unsigned write_lists(const PackArgs<T> &a, long long cn, uint32_t *tilecol, uint32_t *sidecol)
{
    const long long j0 = a.chunk_item0[cn], j1 = a.chunk_item0[cn + 1];
    const bool ragged_cols = (a.colA % TS) != 0;
    unsigned flags = 0;
    int tbase = 0, sbase = 0;
    if (a.allow_flat)
    {
        // a chunk without stream tiles is flat: its side columns go in the flat order (warp 0 walks it)
        bool any_tile = false;
        for (long long j = j0; j < j1 && !any_tile; j++)
            any_tile = a.sc.nc[a.items[j].t1] != a.sc.nc[a.items[j].t0];
        if (!any_tile)
        {
            if (threadIdx.x < 32)
                flat_walk<T>(a, cn, (int)threadIdx.x, [&](uint32_t pos, int src) { sidecol[pos] = (uint32_t)a.side_col[src]; }, nullptr, nullptr, nullptr);
            return 0u;
        }
    }
    for (long long j = j0; j < j1; j++)
    {
        const PlanItem it = a.items[j];
        const int nc0 = a.sc.nc[it.t0];
        for (int t = it.t0 + (int)threadIdx.x; t < it.t1; t += (int)blockDim.x)
        {
            if (a.Format[t] == TILESPMV_FMT_COO)
                continue;
            const int tc = a.tile_columnidx[t];
            tilecol[tbase + (a.sc.nc[t] - nc0)] = (uint32_t)tc;
            if (ragged_cols && tc == a.tilen - 1)
                flags |= CHF_PARTIAL_X;
        }
        tbase += a.sc.nc[it.t1] - nc0;
        const int ns = it.s1 - it.s0;
        for (int e = (int)threadIdx.x; e < ns; e += (int)blockDim.x)
            sidecol[sbase + e] = (uint32_t)a.side_col[it.s0 + e];
        sbase += ns;
    }
    return flags;
}

// counts of a chunk's lists (serial, one thread)
template <class T>
__device__ void list_counts(const PackArgs<T> &a, long long cn, uint32_t &nt, uint32_t &ns)
{
    nt = ns = 0;
    for (long long j = a.chunk_item0[cn]; j < a.chunk_item0[cn + 1]; j++)
    {
        const PlanItem it = a.items[j];
        nt += (uint32_t)(a.sc.nc[it.t1] - a.sc.nc[it.t0]);
        ns += (uint32_t)(it.s1 - it.s0);
    }
}

template <class T>
__global__ void __launch_bounds__(PACK_THREADS) pack_kernel(PackArgs<T> a, long long nchunks)
{
    extern __shared__ int s_base[];
    __shared__ ChunkHeader hdr;
    __shared__ int s_start[TS + 1];
    __shared__ unsigned s_flags, s_head_flags;
    __shared__ uint32_t s_own_nt, s_own_ns;
    __shared__ int s_flat;
    __shared__ int g_start[TS][CSRGROUP_MAX_TILES], g_len[TS];
    __shared__ unsigned g_hdr[CSRGROUP_MAX_SLOTROWS];
    const long long c = blockIdx.x;
    if (c >= nchunks)
        return;
    const long long i0 = a.chunk_item0[c], i1 = a.chunk_item0[c + 1];
    const int nitems = (int)(i1 - i0);
    const long long cn = c + a.nw; // the chunk whose lists this one carries
    unsigned char *out = a.stream + a.chunk_off[c];
    constexpr uint32_t vs = (uint32_t)sizeof(T);

    if (threadIdx.x == 0)
    {
        uint32_t ntiles = 0, nother = 0, pay = 0, nside = 0, nsiderows = 0;
        for (int k = 0; k < nitems; k++)
        {
            const PlanItem it = a.items[i0 + k];
            const int nt = a.sc.nc[it.t1] - a.sc.nc[it.t0];
            const int nsr = a.sc.ws[it.t1] - a.sc.ws[it.t0];
            // a grouped item replaces its CSR tiles by ONE descriptor + payload placed before the other tiles
            const int no = it.g_n ? a.sc.oc2[it.t1] - a.sc.oc2[it.t0] + 1 : a.sc.oc[it.t1] - a.sc.oc[it.t0];
            const uint32_t ob = it.g_n ? (uint32_t)(a.sc.ob2[it.t1] - a.sc.ob2[it.t0]) + csr_group_bytes((uint32_t)it.g_nsrg, (uint32_t)it.g_n, vs)
                                       : (uint32_t)(a.sc.ob[it.t1] - a.sc.ob[it.t0]);
            const int ns = it.s1 - it.s0;
            int *sb = s_base + PACK_ITEM_INTS * k;
            sb[0] = (int)ntiles;
            sb[1] = (int)nother;
            sb[2] = (int)pay;
            sb[3] = (int)nside;
            sb[4] = (int)nsiderows;
            sb[5] = nsr;
            RowRec rec;
            rec.dest = it.dest;
            rec.nsr = (uint16_t)nsr;
            rec.nother = (uint16_t)no;
            rec.rowlen = (uint8_t)it.rowlen;
            rec.flags = (uint8_t)(ns > 0 ? ROWF_HAS_SIDE : 0u);
            rec.side_nit = 0; // filled in below once the row starts are known
            rec.ell_bytes16 = (uint16_t)(ell_group_bytes((uint32_t)nsr, vs) / 16u);
            rec.pad0 = 0;
            *reinterpret_cast<RowRec *>(out + CHUNK_OFF_ROWS + 16 * k) = rec;
            ntiles += (uint32_t)nt;
            nother += (uint32_t)no;
            pay += ell_group_bytes((uint32_t)nsr, vs) + ob;
            nside += (uint32_t)ns;
            nsiderows += ns > 0 ? 1u : 0u;
        }
        uint32_t next_nt = 0, next_ns = 0;
        if (cn < nchunks)
            list_counts<T>(a, cn, next_nt, next_ns);
        hdr.nrows = (uint16_t)nitems;
        hdr.ntiles = (uint16_t)ntiles;
        hdr.next_ntiles = (uint16_t)next_nt;
        hdr.next_nside = (uint16_t)next_ns;
        const bool flat = a.allow_flat && ntiles == 0; // no stream tile in the chunk: flat layout (stream.cuh)
        uint32_t off_odesc = CHUNK_OFF_ROWS + 16u * (uint32_t)nitems;
        uint32_t off_sidehdr = off_odesc + pad16(8u * nother);
        uint32_t off_sideval = off_sidehdr + pad16(SIDEHDR_BYTES * nsiderows);
        uint32_t off_payload = off_sideval + pad16(vs * nside);
        uint32_t off_nextlist = off_payload + pay;
        if (flat)
        {
            off_sidehdr = CHUNK_OFF_ROWS + 16u * (uint32_t)nitems;                       // len[16 nrows]
            off_payload = off_sidehdr + 16u * (uint32_t)nitems;                          // FlatLong[]
            off_sideval = off_payload + pad16(8u * (nside / (uint32_t)FLAT_LONG_ROW));   // val[nside]
            off_nextlist = off_sideval + pad16(vs * nside);
            off_odesc = 0;                                                               // number of FlatLong records, set by the walk
            hdr.nrows = (uint16_t)((uint32_t)nitems | CHF_FLAT);
            if (nitems > FLAT_MAX_ROWS)
                atomicExch(a.error_flag, 1);
        }
        s_flat = flat ? 1 : 0;
        hdr.off_nextlist = (uint16_t)off_nextlist;
        hdr.off_odesc = (uint16_t)off_odesc;
        hdr.off_sidehdr = (uint16_t)off_sidehdr;
        hdr.off_sideval = (uint16_t)off_sideval;
        hdr.off_payload = off_payload;
        hdr.next_flags = 0;
        hdr.nside = (uint16_t)nside;
        {
            const long long ci = c + (long long)a.stages * a.nw; // fetched into this chunk's stage once it is consumed
            hdr.issue_off16 = ci < nchunks ? (uint32_t)(a.chunk_off[ci] / 16) : 0u;
            hdr.issue_bytes = ci < nchunks ? (uint32_t)(a.chunk_off[ci + 1] - a.chunk_off[ci]) : 0u;
        }
        *reinterpret_cast<ChunkHeader *>(out) = hdr;
        if ((unsigned long long)(off_nextlist + list_bytes(next_nt, next_ns)) != a.chunk_off[c + 1] - a.chunk_off[c] ||
            ntiles > 256u || nside > 0xffffu || off_nextlist > 0xffffu)
            atomicExch(a.error_flag, 1);
        s_flags = 0;
        s_head_flags = 0;
        s_own_nt = ntiles;
        s_own_ns = nside;
    }
    __syncthreads();

    if (s_flat) // uniform across the CTA: values in the flat order, len[] and the long-row table (one warp walks the chunk)
    {
        if (threadIdx.x < 32)
        {
            T *ov = reinterpret_cast<T *>(out + hdr.off_sideval);
            int nlong = 0;
            flat_walk<T>(a, c, (int)threadIdx.x, [&](uint32_t pos, int src) { ov[pos] = a.side_val[src]; }, out + hdr.off_sidehdr,
                         reinterpret_cast<FlatLong *>(out + hdr.off_payload), &nlong);
            if (threadIdx.x == 0)
            {
                reinterpret_cast<ChunkHeader *>(out)->off_odesc = (uint16_t)nlong;
                if ((uint32_t)nlong > hdr.nside / (uint32_t)FLAT_LONG_ROW)
                    atomicExch(a.error_flag, 1);
            }
        }
    }
    for (int k = 0; k < (s_flat ? 0 : nitems); k++)
    {
        const PlanItem it = a.items[i0 + k];
        const int *sb = s_base + PACK_ITEM_INTS * k;
        const uint32_t nsr = (uint32_t)sb[5];
        unsigned char *rowpay = out + hdr.off_payload + (uint32_t)sb[2];
        T *ell_vals = reinterpret_cast<T *>(rowpay);
        unsigned char *ell_idx = rowpay + nsr * 16u * vs;
        unsigned char *ell_xsel = ell_idx + nsr * 8u;
        unsigned char *other_pay = rowpay + ell_group_bytes(nsr, vs);
        // tiles of this item, one thread per tile
        for (int t = it.t0 + (int)threadIdx.x; t < it.t1; t += PACK_THREADS)
        {
            const int f = a.Format[t];
            if (f == TILESPMV_FMT_COO)
                continue;
            const unsigned xsel = (unsigned)(sb[0] + (a.sc.nc[t] - a.sc.nc[it.t0])); // x segment in the chunk
            if (fmt_is_ell(f))
            {
                const uint32_t so = (uint32_t)(a.sc.ws[t] - a.sc.ws[it.t0]);
                pack_ell_tile<T>(a, t, it.br, xsel, ell_vals + so * 16u, ell_idx + so * 8u, ell_xsel + so);
            }
            else if (it.g_n == 0)
            {
                const uint32_t oi = (uint32_t)(sb[1] + (a.sc.oc[t] - a.sc.oc[it.t0]));
                const uint32_t po = (uint32_t)(a.sc.ob[t] - a.sc.ob[it.t0]);
                pack_other_tile<T>(a, t, it.br, xsel, out + hdr.off_odesc + 8u * oi, other_pay + po);
            }
            else if (f != TILESPMV_FMT_CSR) // grouped item: the group comes first, then the non-CSR tiles
            {
                const uint32_t oi = (uint32_t)(sb[1] + 1 + (a.sc.oc2[t] - a.sc.oc2[it.t0]));
                const uint32_t po = csr_group_bytes((uint32_t)it.g_nsrg, (uint32_t)it.g_n, vs) + (uint32_t)(a.sc.ob2[t] - a.sc.ob2[it.t0]);
                pack_other_tile<T>(a, t, it.br, xsel, out + hdr.off_odesc + 8u * oi, other_pay + po);
            }
        }
        if (it.g_n) // uniform across the CTA
            pack_csr_group<T>(a, it, (unsigned)sb[0], out + hdr.off_odesc + 8u * (uint32_t)sb[1], other_pay, g_start, g_len, g_hdr);
        const int ns = it.s1 - it.s0; // uniform across the CTA
        if (ns > 0)
        {
            if (threadIdx.x <= TS)
            {
                // exclusive starts of every local row inside [s0,s1); entry 16 = ns
                const int r = threadIdx.x;
                int st = ns;
                if (r < it.rowlen)
                {
                    int lo = a.side_ptr[it.br * TS + r];
                    lo = lo > it.s0 ? lo : it.s0;
                    lo = lo < it.s1 ? lo : it.s1;
                    st = lo - it.s0;
                }
                s_start[r] = st;
                reinterpret_cast<uint16_t *>(out + hdr.off_sidehdr + SIDEHDR_BYTES * (uint32_t)sb[4])[r] = (uint16_t)st;
            }
            T *ov = reinterpret_cast<T *>(out + hdr.off_sideval) + sb[3];
            for (int e = threadIdx.x; e < ns; e += PACK_THREADS)
                ov[e] = a.side_val[it.s0 + e];
            __syncthreads();
            if (threadIdx.x == 0)
            {
                // rows with >= SIDE_LONG_ROW entries (pieces of hub rows) are summed by the whole warp, the others by
                // 4 lanes per row pair in nit trips
                int nit = 0;
                unsigned longmask = 0;
                for (int r = 0; r < TS; r++)
                {
                    const int len = s_start[r + 1] - s_start[r];
                    if (len >= a.side_long_row)
                        longmask |= 1u << r;
                    else
                        nit = max(nit, (len + 3) >> 2);
                }
                reinterpret_cast<RowRec *>(out + CHUNK_OFF_ROWS + 16 * k)->side_nit = (uint16_t)nit;
                reinterpret_cast<uint16_t *>(out + hdr.off_sidehdr + SIDEHDR_BYTES * (uint32_t)sb[4])[TS + 1] = (uint16_t)longmask;
            }
            __syncthreads();
        }
    }

    // lists of the chunk the same warp processes next
    if (cn < nchunks)
    {
        uint32_t *tl = reinterpret_cast<uint32_t *>(out + hdr.off_nextlist);
        const unsigned f = write_lists<T>(a, cn, tl, tl + pad16(4u * hdr.next_ntiles) / 4u);
        if (f)
            atomicOr(&s_flags, f);
    }
    // every warp's first chunk: its own lists go to the head array
    if (c < a.nw)
    {
        unsigned char *h = a.head + (size_t)c * a.head_stride;
        uint32_t *tl = reinterpret_cast<uint32_t *>(h + HEAD_HDR_BYTES);
        const unsigned f = write_lists<T>(a, c, tl, tl + pad16(4u * s_own_nt) / 4u);
        if (f)
            atomicOr(&s_head_flags, f);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (s_flags)
            reinterpret_cast<ChunkHeader *>(out)->next_flags = (uint16_t)s_flags;
        if (c < a.nw)
        {
            uint32_t *h = reinterpret_cast<uint32_t *>(a.head + (size_t)c * a.head_stride);
            h[0] = s_own_nt | (s_own_ns << 16);
            h[1] = s_head_flags;
            h[2] = h[3] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host: chunking
// ---------------------------------------------------------------------------------------------
namespace
{
struct ChunkAcc
{
    uint32_t nrows = 0, ntiles = 0, nother = 0, nsiderows = 0, nside = 0, payload = 0;
    bool allow_flat = false; // side-only chunks use the flat layout (stream.cuh)
    bool empty() const { return nrows == 0; }
    bool is_flat() const { return allow_flat && ntiles == 0 && nother == 0 && payload == 0; }
    uint32_t main_bytes(uint32_t vs) const
    {
        return is_flat() ? flat_chunk_main_bytes(nrows, nside, vs) : chunk_main_bytes(nrows, nother, nsiderows, nside, payload, vs);
    }
    // budget check: the chunk with lists as long as its own (the lists it really carries are those
    // of the chunk the same warp processes next; the stage stride is widened afterwards if needed)
    uint32_t bytes(uint32_t vs) const { return main_bytes(vs) + list_bytes(ntiles, nside); }
    uint32_t xbytes(uint32_t vs) const { return ntiles * 16u * vs + nside * vs; }
    void add(const ChunkAcc &o)
    {
        allow_flat = allow_flat || o.allow_flat;
        nrows += o.nrows;
        ntiles += o.ntiles;
        nother += o.nother;
        nsiderows += o.nsiderows;
        nside += o.nside;
        payload += o.payload;
    }
};
} // namespace

// what one (sub-)plan is built from: the tiles of dm (or none) + a side matrix (dm's own, or one column panel of it)
struct PlanSource
{
    bool tiles;
    const int *tile_ptr; // dm->tile_ptr, or tilem+1 zeros when tiles is false
    const int *side_ptr, *side_col;
    const void *side_val;
};

template <class T>
static int plan_build_t(const tilespmv_dmat *dm, const PlanSource &src, tilespmv_plan *P, cudaStream_t s)
{
    const uint32_t vs = (uint32_t)sizeof(T);
    const int T_ = src.tiles ? dm->tilenum : 0, tilem = dm->tilem, rowA = dm->rowA;
    uint32_t C = (uint32_t)P->chunk_bytes, X = (uint32_t)P->xstage_bytes; // 0 = chosen below from the row sizes
    ScanWorkspace ws;

    // formats this plan covers (0 = all): excluded tiles are treated like COO tiles (skipped by the stream)
    DevBuf d_fmt_eff;
    const char *fmt_eff = dm->Format.as<char>();
    const unsigned fmask = (unsigned)P->format_mask & 0x7fu;
    if (P->format_mask != 0 && fmask != 0x7fu && T_ > 0) // mask 0x80: restricted to NO format (only the fixed per-row cost is left)
    {
        TSP_TRY(d_fmt_eff.alloc((size_t)T_, false));
        TSP_LAUNCH(filter_format_kernel, grid_for((size_t)T_, PL_THREADS), PL_THREADS, 0, s, T_, dm->Format.as<char>(), fmask, d_fmt_eff.as<char>());
        fmt_eff = d_fmt_eff.as<char>();
    }

    // ---- 1. per-tile prefix sums ----
    DevBuf d_nc, d_oc, d_ws, d_ob, d_hi, d_oc2, d_ob2;
    TSP_TRY(d_oc2.alloc((size_t)(T_ + 1) * sizeof(int), true, s));
    TSP_TRY(d_ob2.alloc((size_t)(T_ + 1) * sizeof(long long), true, s));
    TSP_TRY(d_nc.alloc((size_t)(T_ + 1) * sizeof(int), true, s));
    TSP_TRY(d_oc.alloc((size_t)(T_ + 1) * sizeof(int), true, s));
    TSP_TRY(d_ws.alloc((size_t)(T_ + 1) * sizeof(int), true, s));
    TSP_TRY(d_ob.alloc((size_t)(T_ + 1) * sizeof(long long), true, s));
    TSP_TRY(d_hi.alloc((size_t)(T_ + 1) * sizeof(int), true, s));
    if (T_)
    {
        TileCostIn tc{fmt_eff, dm->tile_nnz.as<int>(), dm->tilewidth.as<char>(),
                      dm->dnsrowptr.as<int>(), dm->dnscolptr.as<int>(), T_, vs, TC_STREAM_TILES};
        TSP_TRY(exclusive_scan(tc, (size_t)T_ + 1, d_nc.as<int>(), ws, s, nullptr));
        tc.kind = TC_OTHER_TILES;
        TSP_TRY(exclusive_scan(tc, (size_t)T_ + 1, d_oc.as<int>(), ws, s, nullptr));
        tc.kind = TC_SLOTROWS;
        TSP_TRY(exclusive_scan(tc, (size_t)T_ + 1, d_ws.as<int>(), ws, s, nullptr));
        tc.kind = TC_OTHER_BYTES;
        TSP_TRY(exclusive_scan(tc, (size_t)T_ + 1, d_ob.as<long long>(), ws, s, nullptr));
        tc.kind = TC_NONCSR_TILES;
        TSP_TRY(exclusive_scan(tc, (size_t)T_ + 1, d_oc2.as<int>(), ws, s, nullptr));
        tc.kind = TC_NONCSR_BYTES;
        TSP_TRY(exclusive_scan(tc, (size_t)T_ + 1, d_ob2.as<long long>(), ws, s, nullptr));
        if (dm->fmt_hist[TILESPMV_FMT_HYB] > 0)
        {
            // byte offset of every HYB tile inside hybIdx (what the reference calls ptroffset2, tilespmv_cpu.h:196)
            int last_tile0 = 0;
            TSP_CUDA(cudaMemcpyAsync(&last_tile0, dm->tile_ptr.as<int>() + (tilem - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
            TSP_CUDA(cudaStreamSynchronize(s));
            tc.kind = TC_HYB_IDXBYTES;
            tc.blknnz = dm->blknnz.as<int>();
            tc.last_row_tile0 = last_tile0;
            tc.last_rowlen = rowA - (tilem - 1) * TS;
            TSP_TRY(exclusive_scan(tc, (size_t)T_ + 1, d_hi.as<int>(), ws, s, nullptr));
        }
    }
    TileScans sc{d_nc.as<int>(), d_oc.as<int>(), d_ws.as<int>(), d_ob.as<long long>(), d_oc2.as<int>(), d_ob2.as<long long>()};
    DevBuf d_row_nt, d_row_no, d_row_nsr, d_row_ob, d_row_s0, d_row_no2, d_row_ob2, d_row_cc, d_row_cn, d_row_cs;
    const bool use_groups = !(P->flags & TILESPMV_PLAN_NO_CSR_GROUPS) && src.tiles && dm->fmt_hist[TILESPMV_FMT_CSR] > 0;
    const size_t nb1 = (size_t)tilem + 1;
    TSP_TRY(d_row_nt.alloc(nb1 * sizeof(int), true, s));
    TSP_TRY(d_row_no.alloc(nb1 * sizeof(int), true, s));
    TSP_TRY(d_row_nsr.alloc(nb1 * sizeof(int), true, s));
    TSP_TRY(d_row_ob.alloc(nb1 * sizeof(long long), true, s));
    TSP_TRY(d_row_s0.alloc(nb1 * sizeof(int), true, s));
    TSP_TRY(d_row_no2.alloc(nb1 * sizeof(int), true, s));
    TSP_TRY(d_row_ob2.alloc(nb1 * sizeof(long long), true, s));
    TSP_LAUNCH(row_summary_kernel, grid_for(nb1, PL_THREADS), PL_THREADS, 0, s, tilem, rowA, src.tile_ptr, sc,
               src.side_ptr, d_row_nt.as<int>(), d_row_no.as<int>(), d_row_nsr.as<int>(),
               d_row_ob.as<long long>(), d_row_s0.as<int>(), d_row_no2.as<int>(), d_row_ob2.as<long long>());
    std::vector<long long> row_ob(nb1), row_ob2(nb1);
    std::vector<int> row_nt(nb1), row_no(nb1), row_nsr(nb1), row_s0(nb1), tile_ptr(nb1), row_no2(nb1);
    std::vector<int> row_cc, row_cn, row_cs; // CSR tiles / their nonzeros / slot-rows of a would-be CSR group
    if (use_groups && tilem > 0)
    {
        TSP_TRY(d_row_cc.alloc((size_t)tilem * sizeof(int), true, s));
        TSP_TRY(d_row_cn.alloc((size_t)tilem * sizeof(int), true, s));
        TSP_TRY(d_row_cs.alloc((size_t)tilem * sizeof(int), true, s));
        TSP_LAUNCH(row_csr_summary_kernel, grid_for((size_t)tilem, PL_THREADS), PL_THREADS, 0, s, tilem, rowA,
                   dm->tile_ptr.as<int>(), fmt_eff, dm->tile_nnz.as<int>(), dm->csrptr_offset.as<int>(),
                   dm->Blockcsr_Ptr.as<unsigned char>(), d_row_nt.as<int>(), d_row_cc.as<int>(), d_row_cn.as<int>(),
                   d_row_cs.as<int>());
        row_cc.resize(tilem);
        row_cn.resize(tilem);
        row_cs.resize(tilem);
        TSP_CUDA(cudaMemcpyAsync(row_cc.data(), d_row_cc.p, (size_t)tilem * sizeof(int), cudaMemcpyDeviceToHost, s));
        TSP_CUDA(cudaMemcpyAsync(row_cn.data(), d_row_cn.p, (size_t)tilem * sizeof(int), cudaMemcpyDeviceToHost, s));
        TSP_CUDA(cudaMemcpyAsync(row_cs.data(), d_row_cs.p, (size_t)tilem * sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    TSP_CUDA(cudaMemcpyAsync(row_ob2.data(), d_row_ob2.p, nb1 * sizeof(long long), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaMemcpyAsync(row_no2.data(), d_row_no2.p, nb1 * sizeof(int), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaMemcpyAsync(row_ob.data(), d_row_ob.p, nb1 * sizeof(long long), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaMemcpyAsync(row_nt.data(), d_row_nt.p, nb1 * sizeof(int), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaMemcpyAsync(row_no.data(), d_row_no.p, nb1 * sizeof(int), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaMemcpyAsync(row_nsr.data(), d_row_nsr.p, nb1 * sizeof(int), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaMemcpyAsync(row_s0.data(), d_row_s0.p, nb1 * sizeof(int), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaMemcpyAsync(tile_ptr.data(), src.tile_ptr, nb1 * sizeof(int), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaStreamSynchronize(s));

    // ---- 1b. chunk size.  A block row that does not fit a chunk is cut into pieces whose partial sums
    //          take a second pass, so the stage should hold a typical block row (a 27-point stencil
    //          row is 4 KB, a 37-per-row band 5.5 KB); larger stages mean fewer resident warps.
    //          Pick the smallest candidate that leaves at most ~5 % of the stream bytes (or what the
    //          largest candidate leaves, + 5 %) in rows that have to be cut.
    // a block row whose CSR tiles can be merged into a CSR group (stream.cuh): payload / descriptor count with the group
    auto groupable = [&](int b) {
        return use_groups && row_cc[b] > 0 && row_nt[b] <= CSRGROUP_MAX_TILES && row_cs[b] <= CSRGROUP_MAX_SLOTROWS;
    };
    auto row_other_bytes = [&](int b) {
        return groupable(b) ? row_ob2[b] + (long long)csr_group_bytes((uint32_t)row_cs[b], (uint32_t)row_cn[b], vs) : row_ob[b];
    };
    auto row_other_count = [&](int b) { return groupable(b) ? row_no2[b] + 1 : row_no[b]; };
    // mostly extracted (side) entries among the nonzeros THIS (sub-)plan handles?  (also drives the shared-memory cap)
    const int64_t my_side = (int64_t)row_s0[tilem] - row_s0[0], my_tiled = src.tiles ? dm->nnz - dm->coototal : 0;
    const bool gather_bound = my_side > 0 && my_side >= my_tiled;
    P->gather_bound = gather_bound;
    if (C == 0)
    {
        // Matrices made mostly of extracted entries: score of a candidate = resident warps per SM it leaves (2 stages +
        // 2 x buffers per warp in 227 KB, at most 20) x (1 - half the share of stream bytes in rows that would have to
        // be cut); 4 KB first: it wins ties.  Everything else: the smallest stage >= 4 KB that fits the rows.
        const uint32_t cand[7] = {4096u, 5120u, 6144u, 7168u, 8192u, 3072u, 2560u};
        double unfit[7] = {0, 0, 0, 0, 0, 0, 0}, total = 0, total_x = 0;
        for (int b = 0; b < tilem; b++)
        {
            const int ns = row_s0[b + 1] - row_s0[b];
            const double bytes = (double)CHUNK_OFF_ROWS + 16.0 + (double)pad16(8u * (uint32_t)row_other_count(b)) + (ns > 0 ? 48.0 : 0.0) +
                                 (double)vs * ns + (double)ell_group_bytes((uint32_t)row_nsr[b], vs) + (double)row_other_bytes(b) +
                                 (double)list_bytes((uint32_t)row_nt[b], (uint32_t)ns);
            const double xb = (double)row_nt[b] * 16.0 * vs + (double)ns * vs;
            total += bytes;
            total_x += xb;
            for (int k = 0; k < 7; k++)
                if (bytes > cand[k] || xb > (X ? X : cand[k] * vs / 8u))
                    unfit[k] += bytes;
        }
        int pick = 0;
        if (total > 0 && gather_bound)
        {
            // scattered x gathers are latency-bound: resident warps matter more than rows cut into pieces
            double best = -1.0;
            for (int k = 0; k < 7; k++)
            {
                const double xk = X ? (double)X : std::min((double)(cand[k] * vs / 8u), total_x / total * cand[k] * 1.3 + 256.0);
                const double warps = std::min(20.0, std::floor(((double)P->gather_smem_cap - SPMV_SMEM_FIXED) / (2.0 * cand[k] + 2.0 * xk)));
                const double score = warps * (1.0 - 0.5 * unfit[k] / total);
                if (score > best + 1e-9)
                {
                    best = score;
                    pick = k;
                }
            }
        }
        else if (total > 0)
        {
            // streaming tiles: the smallest stage >= 4 KB that leaves at most ~5 % of the stream bytes (or what the
            // largest candidate leaves, + 5 %) in rows that have to be cut
            const double limit = std::max(0.05, unfit[4] / total + 0.05);
            while (pick < 4 && unfit[pick] / total > limit)
                pick++;
        }
        C = cand[pick];
        P->chunk_bytes = (int)C;
    }
    if (X == 0)
    {
        X = C * vs / 8u; // as many staged x values as a chunk of pure side entries can hold (C / 8 per fp64 ...)
        P->xstage_bytes = (int)X;
    }

    // ---- 2. greedy byte-bounded chunking over block rows ----
    std::vector<PlanItem> items;
    std::vector<long long> chunk_item0;
    std::vector<unsigned long long> chunk_off;
    std::vector<int> split_tab; // 4 ints per split row
    items.reserve((size_t)tilem + 16);
    int64_t nslots = 0;
    const bool allow_flat = !(P->flags & TILESPMV_PLAN_NO_FLAT_SIDE);
    ChunkAcc acc;
    acc.allow_flat = allow_flat;
    std::vector<uint32_t> ch_main, ch_nt, ch_ns; // per chunk: bytes without lists, stream tiles, side entries
    auto close_chunk = [&]() {
        if (acc.empty())
            return;
        ch_main.push_back(acc.main_bytes(vs));
        ch_nt.push_back(acc.ntiles);
        ch_ns.push_back(acc.nside);
        chunk_item0.push_back((long long)items.size());
        acc = ChunkAcc();
        acc.allow_flat = allow_flat;
    };
    chunk_item0.push_back(0);
    // does block row b fit a chunk on its own?  (rows that do not are cut into pieces below)
    auto row_fits_alone = [&](int b, ChunkAcc &one) {
        const int ns = row_s0[b + 1] - row_s0[b];
        const long long pay_ll = (long long)ell_group_bytes((uint32_t)row_nsr[b], vs) + row_other_bytes(b);
        one = ChunkAcc();
        one.allow_flat = allow_flat;
        one.nrows = 1;
        one.ntiles = (uint32_t)row_nt[b];
        one.nother = (uint32_t)row_other_count(b);
        one.nside = (uint32_t)ns;
        one.nsiderows = ns > 0 ? 1 : 0;
        bool fits = pay_ll < (long long)C && row_nt[b] <= 256 && row_nsr[b] < 60000 && ns < (int)C;
        if (fits)
        {
            one.payload = (uint32_t)pay_ll;
            fits = one.bytes(vs) <= C && one.xbytes(vs) <= X;
        }
        return fits;
    };
    // per-tile prefix sums of all long rows that hold stream tiles, in one gather + one copy
    std::vector<long long> h_ob;
    std::vector<int> h_nc, h_oc, h_ws;
    std::vector<long long> long_off; // per long row (in block-row order): offset into the compact arrays
    {
        std::vector<int> long_ta;
        long_off.push_back(0);
        for (int b = 0; b < tilem; b++)
        {
            const int ns = row_s0[b + 1] - row_s0[b];
            if (!P->keep_all_rows && ns == 0 && row_nt[b] == 0)
                continue;
            ChunkAcc one;
            if (row_nt[b] > 0 && !row_fits_alone(b, one))
            {
                long_ta.push_back(tile_ptr[b]);
                long_off.push_back(long_off.back() + (tile_ptr[b + 1] - tile_ptr[b]) + 1);
            }
        }
        const size_t nlong = long_ta.size(), total = (size_t)long_off.back();
        if (nlong > 0)
        {
            DevBuf d_ta, d_off, c_ob, c_nc, c_oc, c_ws;
            TSP_TRY(d_ta.alloc(nlong * sizeof(int), false));
            TSP_TRY(d_off.alloc((nlong + 1) * sizeof(long long), false));
            TSP_TRY(c_ob.alloc(total * sizeof(long long), false));
            TSP_TRY(c_nc.alloc(total * sizeof(int), false));
            TSP_TRY(c_oc.alloc(total * sizeof(int), false));
            TSP_TRY(c_ws.alloc(total * sizeof(int), false));
            TSP_CUDA(cudaMemcpyAsync(d_ta.p, long_ta.data(), nlong * sizeof(int), cudaMemcpyHostToDevice, s));
            TSP_CUDA(cudaMemcpyAsync(d_off.p, long_off.data(), (nlong + 1) * sizeof(long long), cudaMemcpyHostToDevice, s));
            TSP_LAUNCH(gather_long_rows_kernel, (unsigned)nlong, PL_THREADS, 0, s, (int)nlong, d_ta.as<int>(), d_off.as<long long>(),
                       d_ob.as<long long>(), d_nc.as<int>(), d_oc.as<int>(), d_ws.as<int>(), c_ob.as<long long>(), c_nc.as<int>(),
                       c_oc.as<int>(), c_ws.as<int>());
            h_ob.resize(total);
            h_nc.resize(total);
            h_oc.resize(total);
            h_ws.resize(total);
            TSP_CUDA(cudaMemcpyAsync(h_ob.data(), c_ob.p, total * sizeof(long long), cudaMemcpyDeviceToHost, s));
            TSP_CUDA(cudaMemcpyAsync(h_nc.data(), c_nc.p, total * sizeof(int), cudaMemcpyDeviceToHost, s));
            TSP_CUDA(cudaMemcpyAsync(h_oc.data(), c_oc.p, total * sizeof(int), cudaMemcpyDeviceToHost, s));
            TSP_CUDA(cudaMemcpyAsync(h_ws.data(), c_ws.p, total * sizeof(int), cudaMemcpyDeviceToHost, s));
            TSP_CUDA(cudaStreamSynchronize(s));
        }
    }
    size_t long_idx = 0;
    for (int b = 0; b < tilem; b++)
    {
        const int rowlen = b == tilem - 1 ? rowA - (tilem - 1) * TS : TS;
        const int ns = row_s0[b + 1] - row_s0[b];
        if (!P->keep_all_rows && ns == 0 && row_nt[b] == 0)
            continue; // an accumulating panel plan has nothing to add to this block row
        ChunkAcc one;
        const bool fits_alone = row_fits_alone(b, one);
        if (fits_alone)
        {
            ChunkAcc trial = acc;
            trial.add(one);
            if (trial.bytes(vs) > C || trial.xbytes(vs) > X || trial.nrows > (trial.is_flat() ? (uint32_t)FLAT_MAX_ROWS : 1000u) ||
                trial.ntiles > 256u)
            {
                close_chunk();
                trial = one;
            }
            acc = trial;
            PlanItem item{b, tile_ptr[b], tile_ptr[b + 1], row_s0[b], row_s0[b + 1], (uint32_t)b, rowlen};
            if (groupable(b))
            {
                item.g_nsrg = row_cs[b];
                item.g_n = row_cn[b];
                P->csr_groups++;
            }
            items.push_back(item);
            continue;
        }
        // ---- long block row: cut into pieces, each piece is its own chunk ----
        close_chunk();
        const int tb = tile_ptr[b + 1];
        // the compact arrays hold this row's prefix sums at long_off[long_idx]; index them like the full arrays
        const int ta = tile_ptr[b] - (row_nt[b] > 0 ? (int)long_off[long_idx] : 0);
        const int ta_real = tile_ptr[b];
        const int64_t slot0 = nslots;
        if (row_nt[b] > 0)
        {
            long_idx++;
            int t = ta_real;
            while (t < tb)
            {
                // grow the piece [t, te) tile by tile while it still fits
                int te = t;
                uint32_t p_nt = 0, p_no = 0, p_nsr = 0, p_ob = 0;
                while (te < tb)
                {
                    const uint32_t is_tile = (uint32_t)(h_nc[te + 1 - ta] - h_nc[te - ta]);
                    const uint32_t t_no = (uint32_t)(h_oc[te + 1 - ta] - h_oc[te - ta]);
                    const uint32_t t_nsr = (uint32_t)(h_ws[te + 1 - ta] - h_ws[te - ta]);
                    const uint32_t t_ob = (uint32_t)(h_ob[te + 1 - ta] - h_ob[te - ta]);
                    ChunkAcc trial; // pieces with stream tiles are never flat
                    trial.nrows = 1;
                    trial.ntiles = p_nt + is_tile;
                    trial.nother = p_no + t_no;
                    trial.payload = ell_group_bytes(p_nsr + t_nsr, vs) + p_ob + t_ob;
                    if (is_tile && p_nt > 0 && (trial.bytes(vs) > C || trial.xbytes(vs) > X || trial.ntiles > 256u))
                        break;
                    p_nt += is_tile;
                    p_no += t_no;
                    p_nsr += t_nsr;
                    p_ob += t_ob;
                    te++;
                }
                if (p_nt > 0)
                {
                    acc.nrows = 1;
                    acc.ntiles = p_nt;
                    acc.nother = p_no;
                    acc.payload = ell_group_bytes(p_nsr, vs) + p_ob;
                    items.push_back(PlanItem{b, t, te, row_s0[b], row_s0[b], ROW_PARTIAL | (uint32_t)nslots, rowlen});
                    nslots++;
                    close_chunk();
                }
                t = te;
            }
        }
        if (ns > 0)
        {
            // header, rec, side header (or the flat layout's lens + long-row table padding), value / list paddings
            const uint32_t fixed = CHUNK_OFF_ROWS + 16u + pad16(SIDEHDR_BYTES) + 16u + 16u + (allow_flat ? 32u : 0u);
            uint32_t max_side = (C - fixed) / (4u + vs + (allow_flat ? 1u : 0u));
            if (max_side > X / vs)
                max_side = X / vs;
            for (int s0 = row_s0[b]; s0 < row_s0[b + 1]; s0 += (int)max_side)
            {
                const int s1 = std::min(s0 + (int)max_side, row_s0[b + 1]);
                acc.nrows = 1;
                acc.nside = (uint32_t)(s1 - s0);
                acc.nsiderows = 1;
                items.push_back(PlanItem{b, tb, tb, s0, s1, ROW_PARTIAL | (uint32_t)nslots, rowlen});
                nslots++;
                close_chunk();
            }
        }
        if (nslots == slot0)
        {
            // cannot happen (a long row has tiles or side entries); keep y defined anyway
            acc.nrows = 1;
            items.push_back(PlanItem{b, ta_real, ta_real, row_s0[b], row_s0[b], (uint32_t)b, rowlen});
            close_chunk();
        }
        else
        {
            split_tab.push_back(b);
            split_tab.push_back((int)slot0);
            split_tab.push_back((int)(nslots - slot0));
            split_tab.push_back(rowlen);
        }
        if (nslots > 0x7ffffff0ll)
        {
            set_error("plan: too many partial-sum slots");
            return TILESPMV_ERR_UNSUPPORTED;
        }
    }
    close_chunk();
    const long long nchunks = (long long)ch_main.size();
    if (nchunks > 0x7fffffffll)
    {
        set_error("plan: too many chunks");
        return TILESPMV_ERR_UNSUPPORTED;
    }
    P->nchunks = nchunks;
    P->nsplit = (int64_t)split_tab.size() / 4;
    {
        // rows cut into many pieces (hub rows of power-law matrices) are combined by a CTA each, the
        // others by one thread per row: small rows first in the table
        std::vector<int> small, big;
        for (size_t i = 0; i + 3 < split_tab.size(); i += 4)
        {
            std::vector<int> &dst = split_tab[i + 2] > SPLIT_BIG_SLOTS ? big : small;
            dst.insert(dst.end(), split_tab.begin() + i, split_tab.begin() + i + 4);
        }
        P->nsplit_small = (int64_t)small.size() / 4;
        split_tab = small;
        split_tab.insert(split_tab.end(), big.begin(), big.end());
    }
    P->nslots = nslots;

    // ---- 2b. launch shape and the lookahead distance nw.  Chunk c carries the x-staging lists of
    //          chunk c + nw (what the same warp processes next), so its size -- and with it the
    //          stage stride, the warps that fit and nw itself -- depend on each other: iterate to
    //          the fixed point (uniform matrices converge at once), else fall back to the
    //          nw-independent bound max(main) + max(list).
    uint32_t max_main = 0, max_list = 0, max_x = 0;
    for (long long c = 0; c < nchunks; c++)
    {
        max_x = std::max(max_x, ch_nt[(size_t)c] * 16u * vs + ch_ns[(size_t)c] * vs);
        max_main = std::max(max_main, ch_main[(size_t)c]);
        max_list = std::max(max_list, list_bytes(ch_nt[(size_t)c], ch_ns[(size_t)c]));
    }
    auto stride_for_nw = [&](long long nw) {
        uint32_t need = 0;
        for (long long c = 0; c < nchunks; c++)
        {
            const long long n = c + nw;
            const uint32_t lb = n < nchunks ? list_bytes(ch_nt[(size_t)n], ch_ns[(size_t)n]) : 0u;
            need = std::max(need, ch_main[(size_t)c] + lb);
        }
        return (int)((need + 127u) & ~127u);
    };
    P->xstage_bytes = std::max(128, (int)((max_x + 127u) & ~127u)); // what the chunks really need (<= the X budget)
    P->stage_stride = std::max(128, (int)((std::max(max_main, std::min((uint32_t)C, max_main + max_list)) + 127u) & ~127u));
    bool converged = false;
    for (int it = 0; it < 6 && !converged; it++)
    {
        TSP_TRY(spmv_configure(P));
        const int need = nchunks ? stride_for_nw(P->nw) : 128;
        if (need <= P->stage_stride)
            converged = true;
        else
            P->stage_stride = need;
    }
    if (!converged)
    {
        P->stage_stride = (int)((max_main + max_list + 127u) & ~127u);
        TSP_TRY(spmv_configure(P));
    }
    const long long nw = P->nw;
    chunk_off.resize((size_t)nchunks + 1);
    chunk_off[0] = 0;
    for (long long c = 0; c < nchunks; c++)
    {
        const long long n = c + nw;
        const uint32_t lb = n < nchunks ? list_bytes(ch_nt[(size_t)n], ch_ns[(size_t)n]) : 0u;
        chunk_off[(size_t)c + 1] = chunk_off[(size_t)c] + ch_main[(size_t)c] + lb;
    }
    const unsigned long long off = chunk_off[(size_t)nchunks];
    P->stream_bytes = (int64_t)off;
    // lists of every warp's first chunk
    const long long nhead = std::min(nw, nchunks);
    uint32_t head_list = 0;
    for (long long c = 0; c < nhead; c++)
        head_list = std::max(head_list, list_bytes(ch_nt[(size_t)c], ch_ns[(size_t)c]));
    P->head_stride = (int)(HEAD_HDR_BYTES + head_list);
    TSP_TRY(P->head.alloc((size_t)std::max<long long>(nhead, 1) * (size_t)P->head_stride, true, s));

    // ---- 3. upload tables, pack ----
    DevBuf d_items, d_chunk_item0, d_err;
    TSP_TRY(d_items.alloc(items.size() * sizeof(PlanItem), false));
    TSP_TRY(d_chunk_item0.alloc(chunk_item0.size() * sizeof(long long), false));
    TSP_TRY(P->chunk_off.alloc(chunk_off.size() * sizeof(unsigned long long), false));
    // what the SpMV kernel reads: {offset / 16, bytes} per chunk, one 8-byte load
    if (off / 16 > 0xffffffffull)
    {
        set_error("plan: packed stream larger than 64 GB");
        return TILESPMV_ERR_UNSUPPORTED;
    }
    std::vector<uint2> chunk_desc((size_t)nchunks + 1);
    for (long long c = 0; c < nchunks; c++)
        chunk_desc[(size_t)c] = make_uint2((unsigned)(chunk_off[(size_t)c] / 16), (unsigned)(chunk_off[(size_t)c + 1] - chunk_off[(size_t)c]));
    chunk_desc[(size_t)nchunks] = make_uint2(0u, 0u);
    TSP_TRY(P->chunk_desc.alloc(chunk_desc.size() * sizeof(uint2), false));
    TSP_CUDA(cudaMemcpyAsync(P->chunk_desc.p, chunk_desc.data(), chunk_desc.size() * sizeof(uint2), cudaMemcpyHostToDevice, s));
    TSP_TRY(P->stream.alloc((size_t)off + 16, true, s));
    TSP_TRY(d_err.alloc(sizeof(int), true, s));
    TSP_TRY(P->scratch.alloc((size_t)nslots * TS * vs, true, s));
    TSP_TRY(P->split_tab.alloc(split_tab.size() * sizeof(int), false));
    if (!items.empty())
        TSP_CUDA(cudaMemcpyAsync(d_items.p, items.data(), items.size() * sizeof(PlanItem), cudaMemcpyHostToDevice, s));
    TSP_CUDA(cudaMemcpyAsync(d_chunk_item0.p, chunk_item0.data(), chunk_item0.size() * sizeof(long long), cudaMemcpyHostToDevice, s));
    TSP_CUDA(cudaMemcpyAsync(P->chunk_off.p, chunk_off.data(), chunk_off.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    if (!split_tab.empty())
        TSP_CUDA(cudaMemcpyAsync(P->split_tab.p, split_tab.data(), split_tab.size() * sizeof(int), cudaMemcpyHostToDevice, s));

    if (nchunks > 0)
    {
        long long max_items = 0;
        for (long long c = 0; c < nchunks; c++)
            max_items = std::max(max_items, chunk_item0[c + 1] - chunk_item0[c]);
        PackArgs<T> a;
        a.items = d_items.as<PlanItem>();
        a.chunk_item0 = d_chunk_item0.as<long long>();
        a.chunk_off = P->chunk_off.as<unsigned long long>();
        a.sc = sc;
        a.stream = P->stream.as<unsigned char>();
        a.head = P->head.as<unsigned char>();
        a.nw = P->nw;
        a.stages = P->stages;
        a.head_stride = P->head_stride;
        a.allow_flat = allow_flat ? 1 : 0;
        a.side_long_row = SIDE_LONG_ROW;
        if (const char *e = getenv("TILESPMV_SIDE_LONG_ROW")) // experiments
            a.side_long_row = atoi(e);
        a.error_flag = d_err.as<int>();
        a.rowA = dm->rowA;
        a.colA = dm->colA;
        a.tilem = dm->tilem;
        a.tilen = dm->tilen;
        a.tile_columnidx = dm->tile_columnidx.as<int>();
        a.tile_nnz = dm->tile_nnz.as<int>();
        a.Format = fmt_eff;
        a.tilewidth = dm->tilewidth.as<char>();
        a.csr_offset = dm->csr_offset.as<int>();
        a.csrptr_offset = dm->csrptr_offset.as<int>();
        a.ell_offset = dm->ell_offset.as<int>();
        a.hyb_offset = dm->hyb_offset.as<int>();
        a.hyb_idxoff = d_hi.as<int>();
        a.Blockhyb_Val = dm->Blockhyb_Val.as<T>();
        a.hybIdx = dm->hybIdx.as<unsigned char>();
        a.dns_offset = dm->dns_offset.as<int>();
        a.dnsrow_offset = dm->dnsrow_offset.as<int>();
        a.dnscol_offset = dm->dnscol_offset.as<int>();
        a.dnsrowptr = dm->dnsrowptr.as<int>();
        a.dnscolptr = dm->dnscolptr.as<int>();
        a.Blockcsr_Val = dm->Blockcsr_Val.as<T>();
        a.Blockell_Val = dm->Blockell_Val.as<T>();
        a.Blockdense_Val = dm->Blockdense_Val.as<T>();
        a.Blockdenserow_Val = dm->Blockdenserow_Val.as<T>();
        a.Blockdensecol_Val = dm->Blockdensecol_Val.as<T>();
        a.Blockcsr_Ptr = dm->Blockcsr_Ptr.as<unsigned char>();
        a.csr_compressedIdx = dm->csr_compressedIdx.as<unsigned char>();
        a.ell_compressedIdx = dm->ell_compressedIdx.as<unsigned char>();
        a.denserowid = dm->denserowid.as<char>();
        a.densecolid = dm->densecolid.as<char>();
        a.side_ptr = src.side_ptr;
        a.side_col = src.side_col;
        a.side_val = static_cast<const T *>(src.side_val);
        const size_t shm = (size_t)max_items * PACK_ITEM_INTS * sizeof(int);
        if (shm > 200 * 1024)
        {
            set_error("plan: too many block rows in one chunk");
            return TILESPMV_ERR_UNSUPPORTED;
        }
        if (shm > 48 * 1024)
            TSP_CUDA(cudaFuncSetAttribute(pack_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
        if (nchunks > 0x7fffffffll)
        {
            set_error("plan: too many chunks");
            return TILESPMV_ERR_UNSUPPORTED;
        }
        TSP_LAUNCH((pack_kernel<T>), (unsigned)nchunks, PACK_THREADS, shm, s, a, nchunks);
    }
    int h_err = 0;
    TSP_CUDA(cudaMemcpyAsync(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaStreamSynchronize(s));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
    {
        set_error("plan: pack kernel failed: %s", cudaGetErrorString(e));
        return TILESPMV_ERR_CUDA;
    }
    if (h_err)
    {
        set_error("plan: internal error, chunk layout size mismatch between host and device");
        return TILESPMV_ERR_CUDA;
    }
    return TILESPMV_OK;
}

// ---------------------------------------------------------------------------------------------
// x panels: the side matrix re-ordered panel-major (panel = column / panel_cols).  Rows are ascending by
// column (csr2tile.h:952-960), so the entries of (row i, panel p) are a contiguous piece of row i.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_col(const int *col, int lo, int hi, long long key)
{
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if ((long long)col[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}
struct PanelCountIn // flattened [P][rowA + 1]: entries of (panel, row); the last slot of every panel is 0
{
    const int *side_ptr, *side_col;
    int rowA, npanels;
    const long long *cuts; // [npanels + 1] ascending column cuts; panel p = columns [cuts[p], cuts[p+1])
    __device__ __forceinline__ int operator()(size_t k) const
    {
        const size_t stride = (size_t)rowA + 1;
        if (k >= stride * (size_t)npanels)
            return 0;
        const int p = (int)(k / stride), i = (int)(k % stride);
        if (i >= rowA)
            return 0;
        const int lo = side_ptr[i], hi = side_ptr[i + 1];
        return lower_bound_col(side_col, lo, hi, cuts[p + 1]) - lower_bound_col(side_col, lo, hi, cuts[p]);
    }
};
template <class T>
__global__ void __launch_bounds__(PL_THREADS)
    panel_scatter_kernel(int rowA, int npanels, const long long *__restrict__ cuts, const int *__restrict__ side_ptr,
                         const int *__restrict__ side_col, const T *__restrict__ side_val, const int *__restrict__ ptr2,
                         int *__restrict__ col2, T *__restrict__ val2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rowA)
        return;
    const int lo = side_ptr[i], hi = side_ptr[i + 1];
    int q = lower_bound_col(side_col, lo, hi, cuts[0]);
    for (int p = 0; p < npanels && q < hi; p++)
    {
        const int e = lower_bound_col(side_col, q, hi, cuts[p + 1]);
        int dst = ptr2[(size_t)p * ((size_t)rowA + 1) + i];
        for (; q < e; q++, dst++)
        {
            col2[dst] = side_col[q];
            val2[dst] = side_val[q];
        }
    }
}

// smallest / largest x column a plan source reads: tile columns of the stream tiles (COO tiles live in the side
// matrix) and the global columns of the side entries.  out = {min, max} (INT_MAX / -1 when nothing is read).
__global__ void __launch_bounds__(PL_THREADS)
    col_range_kernel(long long ntiles, const int *__restrict__ tile_columnidx, const char *__restrict__ Format,
                     long long nside, const int *__restrict__ side_col, int *__restrict__ out)
{
    int lo = 0x7fffffff, hi = -1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += stride)
        if (Format[t] != TILESPMV_FMT_COO)
        {
            const int c = tile_columnidx[t] * TS;
            lo = min(lo, c);
            hi = max(hi, c + TS - 1);
        }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nside; e += stride)
    {
        const int c = side_col[e];
        lo = min(lo, c);
        hi = max(hi, c);
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (lo != 0x7fffffff)
            atomicMin(out, lo);
        if (hi >= 0)
            atomicMax(out + 1, hi);
    }
}

// records [xcol_lo, xcol_hi) of a (sub-)plan: which part of x its launch reads (drives the per-launch dependencies
// of the pipelined multi-GPU exchange, comm.cu)
static int plan_col_range(const tilespmv_dmat *dm, bool tiles, const int *side_col, long long nside, tilespmv_plan *P, cudaStream_t s)
{
    DevBuf d;
    TSP_TRY(d.alloc(2 * sizeof(int), false));
    const int init[2] = {0x7fffffff, -1};
    TSP_CUDA(cudaMemcpyAsync(d.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
    const long long nt = tiles ? dm->tilenum : 0;
    if (nt + nside > 0)
        TSP_LAUNCH(col_range_kernel, std::min<unsigned>(grid_for((size_t)std::max(nt, nside), PL_THREADS), 2048u), PL_THREADS, 0, s, nt,
                   dm->tile_columnidx.as<int>(), dm->Format.as<char>(), nside, side_col, d.as<int>());
    int h[2];
    TSP_CUDA(cudaMemcpyAsync(h, d.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaStreamSynchronize(s));
    P->xcol_lo = h[1] >= 0 ? h[0] : 0;
    P->xcol_hi = h[1] >= 0 ? std::min<long long>((long long)h[1] + 1, dm->colA) : 0;
    return TILESPMV_OK;
}

// cuts[npanels + 1]: ascending column cuts (cuts[0] = 0, cuts[npanels] >= colA); order[npanels]: the launch order of the
// panels.  The plan itself takes the tiles + panel order[0] and WRITES y; sub[i-1] takes panel order[i] and accumulates.
template <class T>
static int plan_build_panels(const tilespmv_dmat *dm, tilespmv_plan *P, const std::vector<long long> &cuts,
                             const std::vector<int> &order, cudaStream_t s)
{
    const int rowA = dm->rowA, npanels = (int)order.size();
    const size_t stride = (size_t)rowA + 1, flat = stride * (size_t)npanels;
    ScanWorkspace ws;
    DevBuf ptr2, col2, val2, zeros, d_cuts;
    TSP_TRY(ptr2.alloc((flat + 1) * sizeof(int), false));
    TSP_TRY(col2.alloc((size_t)dm->coototal * sizeof(int), false));
    TSP_TRY(val2.alloc((size_t)dm->coototal * sizeof(T), false));
    TSP_TRY(zeros.alloc(((size_t)dm->tilem + 1) * sizeof(int), true, s));
    TSP_TRY(d_cuts.alloc(cuts.size() * sizeof(long long), false));
    TSP_CUDA(cudaMemcpyAsync(d_cuts.p, cuts.data(), cuts.size() * sizeof(long long), cudaMemcpyHostToDevice, s));
    long long total = 0;
    PanelCountIn in{dm->deferredcoo_ptr.as<int>(), dm->deferredcoo_colidx.as<int>(), rowA, npanels, d_cuts.as<long long>()};
    TSP_TRY(exclusive_scan(in, flat + 1, ptr2.as<int>(), ws, s, &total));
    if (total != (long long)dm->coototal)
    {
        set_error("plan: x-panel split lost entries (%lld of %d)", total, dm->coototal);
        return TILESPMV_ERR_CUDA;
    }
    TSP_LAUNCH((panel_scatter_kernel<T>), grid_for((size_t)rowA, PL_THREADS), PL_THREADS, 0, s, rowA, npanels, d_cuts.as<long long>(),
               dm->deferredcoo_ptr.as<int>(), dm->deferredcoo_colidx.as<int>(), dm->deferredcoo_val.as<T>(), ptr2.as<int>(),
               col2.as<int>(), val2.as<T>());
    const int user_chunk = P->chunk_bytes, user_xstage = P->xstage_bytes;
    for (int i = 0; i < npanels; i++)
    {
        const int p = order[(size_t)i];
        tilespmv_plan *Q = P;
        if (i > 0)
        {
            Q = new (std::nothrow) tilespmv_plan();
            if (!Q)
                return TILESPMV_ERR_ALLOC;
            P->sub.push_back(Q);
            Q->precision = P->precision;
            Q->rowA = P->rowA;
            Q->colA = P->colA;
            Q->tilem = P->tilem;
            Q->ctas_per_sm = P->ctas_per_sm;
            Q->stages = P->stages;
            Q->max_warps = P->max_warps;
            Q->flags = P->flags;
            Q->gather_smem_cap = P->gather_smem_cap;
            Q->chunk_bytes = user_chunk;
            Q->xstage_bytes = user_xstage;
            Q->accumulate = true;
            Q->keep_all_rows = i == npanels - 1; // the last launch visits every row: it carries the fused peer stores
        }
        PlanSource src{i == 0, i == 0 ? dm->tile_ptr.as<int>() : zeros.as<int>(), ptr2.as<int>() + (size_t)p * stride, col2.as<int>(),
                       val2.p};
        TSP_TRY(plan_build_t<T>(dm, src, Q, s));
        // the part of x this launch reads: its panel (clipped to colA), widened by the tile columns for the first one
        Q->xcol_lo = std::min<long long>(cuts[(size_t)p], dm->colA);
        Q->xcol_hi = std::min<long long>(cuts[(size_t)p + 1], dm->colA);
        if (i == 0 && dm->tilenum > dm->fmt_hist[TILESPMV_FMT_COO])
        {
            tilespmv_plan tmp;
            TSP_TRY(plan_col_range(dm, true, nullptr, 0, &tmp, s));
            if (tmp.xcol_hi > tmp.xcol_lo)
            {
                Q->xcol_lo = std::min(Q->xcol_lo, tmp.xcol_lo);
                Q->xcol_hi = std::max(Q->xcol_hi, tmp.xcol_hi);
            }
        }
    }
    TSP_CUDA(cudaStreamSynchronize(s));
    return TILESPMV_OK;
}

// x panels (plan.cuh): automatic when x is larger than L2 (panels of half of L2) and at least a quarter of the nonzeros are
// side entries (random gathers); xpanel_bytes > 0 forces that panel width, < 0 switches panels off.  *out = 0: no panels.
int plan_panel_bytes(const tilespmv_dmat *dm, int xpanel_bytes, long long *out)
{
    *out = 0;
    if (xpanel_bytes > 0)
        *out = xpanel_bytes;
    else if (xpanel_bytes == 0 && dm->coototal > 0 && (int64_t)dm->coototal * 4 >= dm->nnz)
    {
        int dev = 0, l2 = 0;
        TSP_CUDA(cudaGetDevice(&dev));
        TSP_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
        const long long budget = (long long)l2 / 2; // config-5 shard: 33 MB panels 1.32 ms, 50 MB 1.08, 67 MB 1.06, 100 MB 1.25
        if ((long long)dm->colA * dm->precision > 2 * budget)
            *out = budget;
    }
    return TILESPMV_OK;
}

int plan_build(const tilespmv_dmat *dm, const tilespmv_plan_options *opts, tilespmv_plan *P, cudaStream_t s, const PanelSpec *spec)
{
    P->precision = dm->precision;
    P->rowA = dm->rowA;
    P->colA = dm->colA;
    P->tilem = dm->tilem;
    P->nnz = dm->nnz;
    const int vs = dm->precision;
    P->chunk_bytes = opts ? opts->chunk_bytes : 0;   // 0 = chosen from the row sizes (4..8 KB)
    P->xstage_bytes = opts ? opts->xstage_bytes : 0; // 0 = follows chunk_bytes
    P->ctas_per_sm = opts ? opts->ctas_per_sm : 0;
    P->stages = opts ? opts->stages : 0;
    P->max_warps = opts ? opts->max_warps : 0;
    P->flags = opts ? opts->flags : 0;
    P->xpanel_bytes = opts ? opts->xpanel_bytes : 0;
    P->format_mask = opts ? opts->format_mask : 0;
    const bool with_side = P->format_mask == 0 || ((P->format_mask >> TILESPMV_FMT_COO) & 1);
    if (const char *e = getenv("TILESPMV_GATHER_SMEM_KB")) // experiments
        P->gather_smem_cap = atoi(e) * 1024;
    if ((P->chunk_bytes != 0 && (P->chunk_bytes < 2560 || P->chunk_bytes > 32768 || (P->chunk_bytes & 127))) ||
        (P->xstage_bytes != 0 && (P->xstage_bytes < 16 * vs || P->xstage_bytes > 32768 || (P->xstage_bytes & 127))))
    {
        set_error("plan: chunk_bytes must be a multiple of 128 in [2560, 32768], xstage_bytes a multiple of 128 in [%d, 32768]",
                  16 * vs);
        return TILESPMV_ERR_INVALID;
    }
    long long panel_bytes = 0;
    TSP_TRY(plan_panel_bytes(dm, P->xpanel_bytes, &panel_bytes));
    std::vector<long long> cuts;
    std::vector<int> order;
    if (spec && spec->cuts.size() >= 2 && dm->coototal > 0)
    {
        // explicit cuts (comm.cu: panels aligned with the row blocks of the ranks, launch order starting at the panel this
        // rank owns): every range wider than panel_bytes (when > 0) is cut further into equal pieces
        const long long maxc = panel_bytes > 0 ? std::max<long long>(TS, panel_bytes / vs / TS * TS) : 0;
        int first = 0;
        for (size_t k = 0; k + 1 < spec->cuts.size(); k++)
        {
            const long long a = spec->cuts[k], b = spec->cuts[k + 1];
            if (b <= a)
                continue;
            const long long pieces = maxc > 0 ? (b - a + maxc - 1) / maxc : 1;
            if ((int)k == spec->first_range)
                first = (int)cuts.size();
            for (long long q = 0; q < pieces; q++)
                cuts.push_back(a + (b - a) * q / pieces);
        }
        cuts.push_back(std::max<long long>(spec->cuts.back(), dm->colA));
        cuts[0] = 0;
        const int np = (int)cuts.size() - 1;
        if (np > 256)
        {
            set_error("plan: %d x panels (limit 256)", np);
            return TILESPMV_ERR_INVALID;
        }
        for (int i = 0; i < np; i++)
            order.push_back((first + i) % np);
    }
    else if (panel_bytes > 0 && dm->coototal > 0)
    {
        long long panel_cols = std::max<long long>(TS, panel_bytes / vs / TS * TS);
        long long np = ((long long)dm->colA + panel_cols - 1) / panel_cols;
        if (np > 64) // bound the number of launches / y passes
        {
            np = 64;
            panel_cols = (((long long)dm->colA + np - 1) / np + TS - 1) / TS * TS;
            np = ((long long)dm->colA + panel_cols - 1) / panel_cols;
        }
        for (long long p = 0; p <= np; p++)
            cuts.push_back(p * panel_cols);
        for (int p = 0; p < (int)np; p++)
            order.push_back(p);
    }
    const bool restricted = P->format_mask != 0 && (P->format_mask & 0x7f) != 0x7f;
    if (restricted) // per-format profiling plans are single launches
        order.clear();
    if (order.size() > 1)
    {
        if (vs == 8)
            TSP_TRY(plan_build_panels<double>(dm, P, cuts, order, s));
        else
            TSP_TRY(plan_build_panels<float>(dm, P, cuts, order, s));
    }
    else
    {
        DevBuf no_side; // a plan without the COO bit leaves the extracted entries out
        if (!with_side)
            TSP_TRY(no_side.alloc(((size_t)dm->rowA + 1) * sizeof(int), true, s));
        PlanSource src{true, dm->tile_ptr.as<int>(), with_side ? dm->deferredcoo_ptr.as<int>() : no_side.as<int>(),
                       dm->deferredcoo_colidx.as<int>(), dm->deferredcoo_val.p};
        if (vs == 8)
            TSP_TRY(plan_build_t<double>(dm, src, P, s));
        else
            TSP_TRY(plan_build_t<float>(dm, src, P, s));
        TSP_CUDA(cudaStreamSynchronize(s));
        TSP_TRY(plan_col_range(dm, true, dm->deferredcoo_colidx.as<int>(), dm->coototal, P, s));
    }

    // roofline accounting, SURVEY.md 8(d): every quantity from the (bit-exact) Tile_matrix
    const int64_t T_coo = dm->fmt_hist[TILESPMV_FMT_COO], T_csr = dm->fmt_hist[TILESPMV_FMT_CSR];
    const int64_t nnz_ext = dm->coototal, nnz_tiled = dm->nnz - nnz_ext, T_tiled = (int64_t)dm->tilenum - T_coo;
    const int64_t m = dm->rowA, n = dm->colA;
    P->b_alg = (nnz_tiled * (2 * vs + 1)) / 2 + 5 * T_tiled + 16 * T_csr + 4 * ((int64_t)dm->tilem + 1) +
               nnz_ext * (vs + 4) + (nnz_ext > 0 ? 4 * (m + 1) : 0) + (int64_t)vs * (n + m);
    P->b_csr = dm->nnz * (vs + 4) + 4 * (m + 1) + (int64_t)vs * (n + m);
    return TILESPMV_OK;
}

// ---------------------------------------------------------------------------------------------
// binary cache of a packed plan (the reference re-converts and re-uploads on every run,
// tilespmv_cuda.h:794-1056).  File = PlanFileHeader, then per (sub-)plan a PlanRecord followed by its four device
// buffers (packed stream, chunk descriptors, head lists, split table) as raw bytes; FNV-1a checksum of everything
// after the header.  A plan is tied to the launch shape it was packed for (the lookahead lists depend on grid x warps),
// so loading on a GPU with another SM count / shared-memory size is refused and the caller re-plans.
// ---------------------------------------------------------------------------------------------
namespace
{
struct PlanFileHeader
{
    char magic[8]; // "TSPPLAN1"
    uint32_t version, nplans;
    int32_t sm_count, smem_optin;
    uint64_t payload_bytes, checksum;
};
struct PlanRecord
{
    int32_t precision, rowA, colA, tilem, chunk_bytes, xstage_bytes, head_stride, stage_stride;
    int32_t grid, block, smem, ctas_per_sm, sm_count, stages, max_warps, flags, format_mask, xpanel_bytes;
    int32_t accumulate, keep_all_rows, gather_bound, pad0;
    int64_t nnz, nchunks, stream_bytes, nw, nsplit, nsplit_small, nslots, csr_groups, b_alg, b_csr, xcol_lo, xcol_hi;
    uint64_t bytes_stream, bytes_desc, bytes_head, bytes_split;
};
constexpr uint32_t PLAN_FILE_VERSION = 1;

uint64_t fnv1a(const unsigned char *p, size_t n, uint64_t h)
{
    for (size_t i = 0; i < n; i++)
    {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

void fill_record(const tilespmv_plan *P, PlanRecord &r)
{
    memset(&r, 0, sizeof(r));
    r.precision = P->precision;
    r.rowA = P->rowA;
    r.colA = P->colA;
    r.tilem = P->tilem;
    r.chunk_bytes = P->chunk_bytes;
    r.xstage_bytes = P->xstage_bytes;
    r.head_stride = P->head_stride;
    r.stage_stride = P->stage_stride;
    r.grid = P->grid;
    r.block = P->block;
    r.smem = P->smem;
    r.ctas_per_sm = P->ctas_per_sm;
    r.sm_count = P->sm_count;
    r.stages = P->stages;
    r.max_warps = P->max_warps;
    r.flags = P->flags;
    r.format_mask = P->format_mask;
    r.xpanel_bytes = P->xpanel_bytes;
    r.accumulate = P->accumulate;
    r.keep_all_rows = P->keep_all_rows;
    r.gather_bound = P->gather_bound;
    r.nnz = P->nnz;
    r.nchunks = P->nchunks;
    r.stream_bytes = P->stream_bytes;
    r.nw = P->nw;
    r.nsplit = P->nsplit;
    r.nsplit_small = P->nsplit_small;
    r.nslots = P->nslots;
    r.csr_groups = P->csr_groups;
    r.b_alg = P->b_alg;
    r.b_csr = P->b_csr;
    r.xcol_lo = P->xcol_lo;
    r.xcol_hi = P->xcol_hi;
    r.bytes_stream = P->stream.bytes;
    r.bytes_desc = P->chunk_desc.bytes;
    r.bytes_head = P->head.bytes;
    r.bytes_split = (uint64_t)P->nsplit * 4 * sizeof(int);
}
} // namespace

int plan_save(const tilespmv_plan *P, const char *path)
{
    std::vector<const tilespmv_plan *> all{P};
    for (const tilespmv_plan *q : P->sub)
        all.push_back(q);
    int dev = 0, smem_optin = 0;
    TSP_CUDA(cudaGetDevice(&dev));
    TSP_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f)
    {
        set_error("plan_save: cannot open %s", tmp.c_str());
        return TILESPMV_ERR_IO;
    }
    PlanFileHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "TSPPLAN1", 8);
    h.version = PLAN_FILE_VERSION;
    h.nplans = (uint32_t)all.size();
    h.sm_count = P->sm_count;
    h.smem_optin = smem_optin;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1; // rewritten at the end with the sizes
    uint64_t sum = 1469598103934665603ull, payload = 0;
    std::vector<unsigned char> host;
    auto put = [&](const void *p, size_t n) {
        ok = ok && (n == 0 || fwrite(p, 1, n, f) == n);
        sum = fnv1a(static_cast<const unsigned char *>(p), n, sum);
        payload += n;
    };
    for (const tilespmv_plan *q : all)
    {
        PlanRecord r;
        fill_record(q, r);
        put(&r, sizeof(r));
        const std::pair<const DevBuf *, uint64_t> bufs[4] = {{&q->stream, r.bytes_stream}, {&q->chunk_desc, r.bytes_desc},
                                                             {&q->head, r.bytes_head}, {&q->split_tab, r.bytes_split}};
        for (const auto &b : bufs)
        {
            host.resize((size_t)b.second);
            if (b.second)
            {
                if (cudaMemcpy(host.data(), b.first->p, (size_t)b.second, cudaMemcpyDeviceToHost) != cudaSuccess)
                {
                    fclose(f);
                    remove(tmp.c_str());
                    set_error("plan_save: D2H failed: %s", cudaGetErrorString(cudaGetLastError()));
                    return TILESPMV_ERR_CUDA;
                }
                put(host.data(), (size_t)b.second);
            }
        }
    }
    h.payload_bytes = payload;
    h.checksum = sum;
    ok = ok && fseek(f, 0, SEEK_SET) == 0 && fwrite(&h, sizeof(h), 1, f) == 1;
    ok = fclose(f) == 0 && ok;
    if (!ok || rename(tmp.c_str(), path) != 0)
    {
        remove(tmp.c_str());
        set_error("plan_save: writing %s failed", path);
        return TILESPMV_ERR_IO;
    }
    return TILESPMV_OK;
}

int plan_load(const char *path, tilespmv_plan **out)
{
    FILE *f = fopen(path, "rb");
    if (!f)
    {
        set_error("plan_load: cannot open %s", path);
        return TILESPMV_ERR_IO;
    }
    PlanFileHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "TSPPLAN1", 8) != 0 || h.version != PLAN_FILE_VERSION || h.nplans < 1 ||
        h.nplans > 1024)
    {
        fclose(f);
        set_error("plan_load: %s is not a plan file of this library version", path);
        return TILESPMV_ERR_IO;
    }
    int dev = 0, sms = 0, smem_optin = 0;
    TSP_CUDA(cudaGetDevice(&dev));
    TSP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TSP_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (sms != h.sm_count || smem_optin != h.smem_optin)
    {
        fclose(f);
        set_error("plan_load: %s was planned for %d SMs / %d B of shared memory, this GPU has %d / %d: plan again", path, h.sm_count,
                  h.smem_optin, sms, smem_optin);
        return TILESPMV_ERR_UNSUPPORTED;
    }
    // the payload size comes from the file: believe it only if the file really is that long (a damaged header must end
    // in TILESPMV_ERR_IO, not in an allocation of 2^60 bytes)
    long file_end = -1;
    if (fseek(f, 0, SEEK_END) == 0)
        file_end = ftell(f);
    std::vector<unsigned char> buf;
    bool read_ok = file_end >= (long)sizeof(h) && (uint64_t)(file_end - (long)sizeof(h)) == h.payload_bytes &&
                   fseek(f, (long)sizeof(h), SEEK_SET) == 0;
    if (read_ok)
    {
        try
        {
            buf.resize((size_t)h.payload_bytes);
        }
        catch (const std::exception &)
        {
            fclose(f);
            set_error("plan_load: out of host memory for the %llu bytes of %s", (unsigned long long)h.payload_bytes, path);
            return TILESPMV_ERR_ALLOC;
        }
        read_ok = fread(buf.data(), 1, buf.size(), f) == buf.size();
    }
    fclose(f);
    if (!read_ok || fnv1a(buf.data(), buf.size(), 1469598103934665603ull) != h.checksum)
    {
        set_error("plan_load: %s is truncated or corrupt", path);
        return TILESPMV_ERR_IO;
    }
    tilespmv_plan *root = nullptr;
    size_t pos = 0;
    auto fail = [&](int rc) {
        delete root;
        return rc;
    };
    for (uint32_t k = 0; k < h.nplans; k++)
    {
        PlanRecord r;
        if (pos + sizeof(r) > buf.size())
            return fail((set_error("plan_load: %s is truncated", path), TILESPMV_ERR_IO));
        memcpy(&r, buf.data() + pos, sizeof(r));
        pos += sizeof(r);
        const uint64_t left = (uint64_t)(buf.size() - pos); // every size is checked on its own: the sum cannot wrap
        if (r.bytes_stream > left || r.bytes_desc > left || r.bytes_head > left || r.bytes_split > left ||
            r.bytes_stream + r.bytes_desc + r.bytes_head + r.bytes_split > left || r.nslots < 0 || r.nsplit < 0 ||
            (r.precision != 4 && r.precision != 8) || r.bytes_split != (uint64_t)r.nsplit * 4 * sizeof(int))
            return fail((set_error("plan_load: %s is truncated or inconsistent", path), TILESPMV_ERR_IO));
        tilespmv_plan *Q = new (std::nothrow) tilespmv_plan();
        if (!Q)
            return fail(TILESPMV_ERR_ALLOC);
        if (k == 0)
            root = Q;
        else
            root->sub.push_back(Q);
        Q->precision = r.precision;
        Q->rowA = r.rowA;
        Q->colA = r.colA;
        Q->tilem = r.tilem;
        Q->chunk_bytes = r.chunk_bytes;
        Q->xstage_bytes = r.xstage_bytes;
        Q->head_stride = r.head_stride;
        Q->stage_stride = r.stage_stride;
        Q->grid = r.grid;
        Q->block = r.block;
        Q->smem = r.smem;
        Q->ctas_per_sm = r.ctas_per_sm;
        Q->sm_count = r.sm_count;
        Q->stages = r.stages;
        Q->max_warps = r.max_warps;
        Q->flags = r.flags;
        Q->format_mask = r.format_mask;
        Q->xpanel_bytes = r.xpanel_bytes;
        Q->accumulate = r.accumulate != 0;
        Q->keep_all_rows = r.keep_all_rows != 0;
        Q->gather_bound = r.gather_bound != 0;
        Q->nnz = r.nnz;
        Q->nchunks = r.nchunks;
        Q->stream_bytes = r.stream_bytes;
        Q->nw = r.nw;
        Q->nsplit = r.nsplit;
        Q->nsplit_small = r.nsplit_small;
        Q->nslots = r.nslots;
        Q->csr_groups = r.csr_groups;
        Q->b_alg = r.b_alg;
        Q->b_csr = r.b_csr;
        Q->xcol_lo = r.xcol_lo;
        Q->xcol_hi = r.xcol_hi;
        const std::pair<DevBuf *, uint64_t> bufs[4] = {{&Q->stream, r.bytes_stream}, {&Q->chunk_desc, r.bytes_desc}, {&Q->head, r.bytes_head},
                                                       {&Q->split_tab, r.bytes_split}};
        for (const auto &b : bufs)
        {
            int rc = b.first->alloc((size_t)b.second, false);
            if (rc != TILESPMV_OK)
                return fail(rc);
            if (b.second && cudaMemcpy(b.first->p, buf.data() + pos, (size_t)b.second, cudaMemcpyHostToDevice) != cudaSuccess)
                return fail((set_error("plan_load: H2D failed: %s", cudaGetErrorString(cudaGetLastError())), TILESPMV_ERR_CUDA));
            pos += (size_t)b.second;
        }
        int rc = Q->scratch.alloc((size_t)Q->nslots * TS * (size_t)Q->precision, true);
        if (rc == TILESPMV_OK)
            rc = spmv_set_attrs(Q);
        if (rc != TILESPMV_OK)
            return fail(rc);
    }
    if (cudaDeviceSynchronize() != cudaSuccess)
        return fail((set_error("plan_load: %s", cudaGetErrorString(cudaGetLastError())), TILESPMV_ERR_CUDA));
    *out = root;
    return TILESPMV_OK;
}

} // namespace tsp

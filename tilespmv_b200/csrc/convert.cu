// convert.cu -- GPU csr2tile: CSR -> device-resident Tile_matrix, bit-exact with the reference's
// CPU conversion (Tile_create, /root/reference/src/csr2tile.h:629-1020).
//
// The reference walks block rows with O(tilem*tilen) scratch (csr2tile.h:5-106) and a per-nnz
// linear tile search (:403-419).  Here the whole conversion is data-parallel over nonzeros and
// tiles:
//   1. key(j) = (block row, tile column, row in tile) for every nonzero; ONE stable radix sort on
//      (block row, tile column) groups the nonzeros per tile in exactly the reference's in-tile
//      order (row-major, original CSR order inside a row -- the low 4 key bits are already
//      ordered, so they ride along unsorted)                               [csr2tile.h:5-106]
//   2. head flags + prefix sum -> tile ids, tile_ptr / tile_columnidx / tile_nnz, per-row starts
//   3. per-tile format selection with the reference's exact double-precision sequence
//      (no FMA contraction: __dmul_rn/__dadd_rn/__ddiv_rn/__dsqrt_rn)      [csr2tile.h:141-326]
//   4. prefix offsets of every per-format array, 8-bit wrapped blknnznnz   [csr2tile.h:729-799]
//   5. scatter of values / local indices into the per-format layouts       [csr2tile.h:427-621]
//   6. 4-bit packing with parity by global position                        [encode.h:29-50]
//   7. side CSR of the COO tiles by stream compaction in original CSR order (+ a radix sort by
//      (row, column) only if some row is not already ascending)            [csr2tile.h:899-960]
#include "dmat.cuh"
#include "primitives.cuh"

namespace tsp
{

constexpr int CV_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// 1. keys
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CV_THREADS)
    make_keys_kernel(const int *__restrict__ rowptr, const int *__restrict__ colidx, int rowA, size_t n,
                     int tcbits, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    int row = upper_row(rowptr, rowA + 1, (int)j);
    uint64_t br = (uint64_t)(row >> 4), r = (uint64_t)(row & 15);
    uint64_t tc = (uint64_t)(colidx[j] >> 4);
    keys[j] = (((br << tcbits) | tc) << 4) | r;
    vals[j] = (uint32_t)j;
}

struct HeadFlagIn // 1 where a new (block row, tile column) group starts in the sorted order
{
    const uint64_t *k;
    size_t n;
    __device__ __forceinline__ int operator()(size_t p) const
    {
        if (p >= n)
            return 0;
        return (p == 0 || (k[p] >> 4) != (k[p - 1] >> 4)) ? 1 : 0;
    }
};

// ---------------------------------------------------------------------------------------------
// 2. tile headers
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CV_THREADS)
    tile_fill_kernel(const uint64_t *__restrict__ keys, const int *__restrict__ headscan, size_t n, int tcbits,
                     int *__restrict__ tile_columnidx, int *__restrict__ tile_nnz, int *__restrict__ tile_br)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    uint64_t k = keys[p];
    if (p == 0 || (k >> 4) != (keys[p - 1] >> 4))
    {
        int t = headscan[p];
        uint64_t g = k >> 4;
        tile_columnidx[t] = (int)(g & ((1ull << tcbits) - 1ull));
        tile_br[t] = (int)(g >> tcbits);
        tile_nnz[t] = (int)p; // exclusive prefix of true nnz == position in the grouped order
    }
}

// in-tile start of every non-empty local row (empty rows keep the 0xFF sentinel, fixed later)
__global__ void __launch_bounds__(CV_THREADS)
    row_start_kernel(const uint64_t *__restrict__ keys, const int *__restrict__ headscan, size_t n,
                     const int *__restrict__ tile_nnz, unsigned char *__restrict__ rowstart)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    uint64_t k = keys[p];
    bool tile_head = p == 0 || (k >> 4) != (keys[p - 1] >> 4);
    if (tile_head || k != keys[p - 1])
    {
        int t = headscan[p] + (tile_head ? 0 : -1); // exclusive scan: heads before p
        rowstart[(size_t)t * TS + (int)(k & 15)] = (unsigned char)((int)p - tile_nnz[t]);
    }
}

__global__ void __launch_bounds__(CV_THREADS)
    tile_ptr_kernel(const int *__restrict__ tile_br, int T, int tilem, int *__restrict__ tile_ptr)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > tilem)
        return;
    tile_ptr[b] = lower_bound_dev(tile_br, T, b); // first tile whose block row is >= b
}

// ---------------------------------------------------------------------------------------------
// 3. format selection (csr2tile.h:141-326; evaluation order of SURVEY.md A.2)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    classify_kernel(int T, const int *__restrict__ tile_nnz, const int *__restrict__ tile_br,
                    const int *__restrict__ tile_columnidx, unsigned char *__restrict__ rowstart,
                    const uint32_t *__restrict__ perm, const int *__restrict__ colidx, int tilem, int tilen,
                    int rowA, int colA, int hyb_vs, char *__restrict__ Format, int *__restrict__ slots_out,
                    char *__restrict__ width_out, int *__restrict__ nd_out)
{
    // hyb_vs: 0 = the reference default (HYB never chosen); sizeof(MAT_VAL_TYPE) = the dormant rule of
    // csr2tile.h:279-316 switched on (TILESPMV_ENABLE_HYB)
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T)
        return;
    const int p0 = tile_nnz[t];
    const int nnz = tile_nnz[t + 1] - p0;
    const int br = tile_br[t], tc = tile_columnidx[t];
    const int rowlen = br == tilem - 1 ? rowA - (tilem - 1) * TS : TS;
    const int collen = tc == tilen - 1 ? colA - (tilen - 1) * TS : TS;

    // row starts: fill the empty rows from the back, then derive the per-row counts
    int rs[TS];
    {
        const uint4 raw = *reinterpret_cast<const uint4 *>(rowstart + (size_t)t * TS);
        const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int r = 0; r < TS; r++)
            rs[r] = (int)((w[r >> 2] >> (8 * (r & 3))) & 255u);
        int next = nnz;
#pragma unroll
        for (int r = TS - 1; r >= 0; r--)
        {
            if (rs[r] == 255)
                rs[r] = next;
            else
                next = rs[r];
        }
        unsigned o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int r = 0; r < TS; r++)
            o[r >> 2] |= ((unsigned)rs[r] & 255u) << (8 * (r & 3));
        *reinterpret_cast<uint4 *>(rowstart + (size_t)t * TS) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    int cnt[TS];
#pragma unroll
    for (int r = 0; r < TS; r++)
        cnt[r] = (r == TS - 1 ? nnz : rs[r + 1]) - rs[r];

    int fmt, slots, width = 0, nd = 0;
    const int dense_th = (int)((double)(rowlen * collen) * 0.75);
    bool decided = false;
    if (nnz >= dense_th)
    {
        fmt = TILESPMV_FMT_DENSE;
        slots = rowlen * collen;
        decided = true;
    }
    else if (nnz <= TILESPMV_COO_NNZ_TH)
    {
        fmt = TILESPMV_FMT_COO;
        slots = nnz;
        decided = true;
    }
    else if (nnz % collen == 0 || nnz % rowlen == 0)
    {
        int flag = 0, num = 0;
        for (int r = 0; r < rowlen; r++)
        {
            if (cnt[r] % collen != 0)
            {
                flag = 0;
                break;
            }
            if (cnt[r] == collen)
            {
                flag = 1;
                num++;
            }
        }
        if (flag)
        {
            fmt = TILESPMV_FMT_DENSEROW;
            nd = num;
            slots = num * collen;
            decided = true;
        }
        else
        {
            // per-column counts of this tile (the reference rescans the block row, :208-216)
            unsigned long long lo = 0, hi = 0; // 16 x 8-bit counters
            for (int p = p0; p < p0 + nnz; p++)
            {
                int lc = colidx[perm[p]] & 15;
                if (lc < 8)
                    lo += 1ull << (8 * lc);
                else
                    hi += 1ull << (8 * (lc - 8));
            }
            flag = 0;
            num = 0;
            for (int c = 0; c < collen; c++)
            {
                int cc = (int)(((c < 8 ? lo : hi) >> (8 * (c & 7))) & 255ull);
                if (cc % rowlen != 0)
                {
                    flag = 0;
                    break;
                }
                if (cc == rowlen)
                {
                    flag = 1;
                    num++;
                }
            }
            if (flag)
            {
                fmt = TILESPMV_FMT_DENSECOL;
                nd = num;
                slots = num * rowlen;
                decided = true;
            }
        }
    }
    if (!decided)
    {
        int wmax = 0;
#pragma unroll
        for (int r = 0; r < TS; r++)
            if (r < rowlen && cnt[r] > wmax)
                wmax = cnt[r];
        // double-precision statistics, rounded step by step like the x86-64 reference build
        const double mean = __ddiv_rn((double)nnz, (double)rowlen);
        double var = 0.0;
        for (int r = 0; r < rowlen; r++)
        {
            double d = __dsub_rn((double)cnt[r], mean);
            var = __dadd_rn(var, __dmul_rn(d, d));
        }
        var = __ddiv_rn(var, (double)rowlen);
        const double sd = __dsqrt_rn(var);
        const double cv = __ddiv_rn(sd, mean);
        if (cv <= 0.2)
        {
            fmt = TILESPMV_FMT_ELL;
            width = wmax;
            slots = wmax * rowlen;
        }
        else
        {
            fmt = TILESPMV_FMT_CSR; // the HYB branch is commented out upstream (:308-316)
            slots = nnz;
            if (hyb_vs)
            {
                // I/O-cost walk of :279-306: shrink the ELL width while the bytes (values + nibbles of the
                // ELL part, value + index byte per spilled entry) keep going down
                int hybwidth = wmax, spill = 0;
                int ioprior = wmax * rowlen * hyb_vs + (wmax * rowlen) / 2 + ((wmax * rowlen) & 1);
                for (int wi = wmax - 1; wi > 0; wi--)
                {
                    int coonext = 0;
                    for (int r = 0; r < rowlen; r++)
                        if (cnt[r] > wi)
                            coonext += cnt[r] - wi;
                    const int ionext = wi * rowlen * hyb_vs + (wi * rowlen) / 2 + ((wi * rowlen) & 1) + coonext * (hyb_vs + 1);
                    if (ioprior <= ionext)
                    {
                        hybwidth = wi + 1;
                        break;
                    }
                    hybwidth = wi;
                    ioprior = ionext;
                    spill = coonext;
                }
                if (cv >= 1.0 && spill <= 4)
                {
                    fmt = TILESPMV_FMT_HYB;
                    width = hybwidth;
                    slots = spill + hybwidth * rowlen;
                    nd = spill;
                }
            }
        }
    }
    Format[t] = (char)fmt;
    slots_out[t] = slots;
    width_out[t] = (char)width;
    nd_out[t] = nd;
}

// pre-scan value of every per-format offset array, derived from (format, slots, nd, rowlen)
enum OffsetKind
{
    OFF_BLKNNZ,
    OFF_CSR,
    OFF_CSRPTR,
    OFF_COO,
    OFF_ELL,
    OFF_DNS,
    OFF_DNSROW,
    OFF_DNSCOL,
    OFF_DNSROWPTR,
    OFF_DNSCOLPTR,
    OFF_NEWCOO,
    OFF_HYB,
    OFF_HYBCOO,
    OFF_HYBIDX,
    OFF_ZERO
};
struct OffsetIn
{
    const char *fmt;
    const int *slots;
    const int *nd;
    const int *tile_br;
    int T, tilem, rowA, kind;
    __device__ __forceinline__ int operator()(size_t i) const
    {
        if (i >= (size_t)T)
            return 0;
        const int f = fmt[i];
        switch (kind)
        {
        case OFF_BLKNNZ:
            return slots[i];
        case OFF_CSR:
            return f == TILESPMV_FMT_CSR ? slots[i] : 0;
        case OFF_CSRPTR:
            return f == TILESPMV_FMT_CSR ? (tile_br[i] == tilem - 1 ? rowA - (tilem - 1) * TS : TS) : 0;
        case OFF_COO:
            return f == TILESPMV_FMT_COO ? slots[i] : 0;
        case OFF_NEWCOO: // COO tiles whole, HYB tiles their spilled entries (:316)
            return f == TILESPMV_FMT_COO ? slots[i] : f == TILESPMV_FMT_HYB ? nd[i] : 0;
        case OFF_HYB:
            return f == TILESPMV_FMT_HYB ? slots[i] : 0;
        case OFF_HYBCOO:
            return f == TILESPMV_FMT_HYB ? nd[i] : 0;
        case OFF_HYBIDX: // bytes of the tile in hybIdx: nibble bytes of the ELL part + one byte per spilled entry (:994-1004)
            return f == TILESPMV_FMT_HYB ? (slots[i] - nd[i] + 1) / 2 + nd[i] : 0;
        case OFF_ELL:
            return f == TILESPMV_FMT_ELL ? slots[i] : 0;
        case OFF_DNS:
            return f == TILESPMV_FMT_DENSE ? slots[i] : 0;
        case OFF_DNSROW:
            return f == TILESPMV_FMT_DENSEROW ? slots[i] : 0;
        case OFF_DNSCOL:
            return f == TILESPMV_FMT_DENSECOL ? slots[i] : 0;
        case OFF_DNSROWPTR:
            return f == TILESPMV_FMT_DENSEROW ? nd[i] : 0;
        case OFF_DNSCOLPTR:
            return f == TILESPMV_FMT_DENSECOL ? nd[i] : 0;
        default:
            return 0;
        }
    }
};

__global__ void __launch_bounds__(CV_THREADS)
    blknnznnz_kernel(const int *__restrict__ slots, int T, unsigned char *__restrict__ out, const char *__restrict__ fmt,
                     unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int h[8];
    if (threadIdx.x < 8)
        h[threadIdx.x] = 0;
    __syncthreads();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= T)
        out[t] = t < T ? (unsigned char)slots[t] : 0; // 8-bit wrap, taken before the scan (:796-797)
    if (t < T)
        atomicAdd(&h[fmt[t] & 7], 1u);
    __syncthreads();
    if (threadIdx.x < 7 && h[threadIdx.x])
        atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// 5. scatter into the per-format layouts (csr2tile.h:427-621)
// ---------------------------------------------------------------------------------------------
template <class T>
struct ScatterArgs
{
    const uint64_t *keys;
    const uint32_t *perm;
    const int *headscan;
    size_t n;
    const int *colidx;
    const T *val;
    int tilem, rowA;
    const int *tile_nnz, *tile_br;
    const char *Format;
    const unsigned char *rowstart;
    const int *csr_offset, *coo_offset, *ell_offset, *hyb_offset, *dns_offset, *dnsrow_offset, *dnscol_offset, *dnscolptr;
    const char *tilewidth;
    T *Blockcsr_Val, *Blockcoo_Val, *Blockell_Val, *Blockhyb_Val, *Blockdense_Val, *Blockdenserow_Val, *Blockdensecol_Val;
    unsigned char *csr_lc, *ell_lc, *hyb_lc, *coo_idx;
    char *densecolid;
    unsigned char *sideflag;
};

template <class T>
__global__ void __launch_bounds__(CV_THREADS) scatter_kernel(ScatterArgs<T> a)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n)
        return;
    const uint64_t key = a.keys[p];
    const bool head = p == 0 || (key >> 4) != (a.keys[p - 1] >> 4);
    const int t = a.headscan[p] + (head ? 0 : -1);
    const int fmt = a.Format[t];
    const int k = (int)p - a.tile_nnz[t];
    const int r = (int)(key & 15);
    const int kr = k - (int)a.rowstart[(size_t)t * TS + r];
    const uint32_t j = a.perm[p];
    const int lc = a.colidx[j] & 15;
    const T v = a.val[j];
    const int br = a.tile_br[t];
    const int rowlen = br == a.tilem - 1 ? a.rowA - (a.tilem - 1) * TS : TS;
    switch (fmt)
    {
    case TILESPMV_FMT_CSR:
    {
        int o = a.csr_offset[t] + k;
        a.Blockcsr_Val[o] = v;
        a.csr_lc[o] = (unsigned char)lc;
        break;
    }
    case TILESPMV_FMT_COO:
    {
        int o = a.coo_offset[t] + k;
        a.Blockcoo_Val[o] = v;
        a.coo_idx[o] = (unsigned char)((r << 4) + lc);
        a.sideflag[j] = 1;
        break;
    }
    case TILESPMV_FMT_ELL:
    {
        int o = a.ell_offset[t] + kr * rowlen + r;
        a.Blockell_Val[o] = v;
        a.ell_lc[o] = (unsigned char)lc;
        break;
    }
    case TILESPMV_FMT_HYB:
    {
        // ELL part slot-major like format 2; entries past the width follow it in row order (:519-546) and
        // are ALSO handed to the side matrix (new_coocount, :538-545)
        const int w = (int)(unsigned char)a.tilewidth[t];
        const int base = a.hyb_offset[t];
        if (kr < w)
        {
            a.Blockhyb_Val[base + kr * rowlen + r] = v;
            a.hyb_lc[base + kr * rowlen + r] = (unsigned char)lc;
        }
        else
        {
            const unsigned char *rs = a.rowstart + (size_t)t * TS;
            int before = 0;
            for (int q = 0; q < r; q++)
            {
                const int len = (int)rs[q + 1] - (int)rs[q];
                before += len > w ? len - w : 0;
            }
            const int o = base + w * rowlen + before + (kr - w);
            a.Blockhyb_Val[o] = v;
            a.hyb_lc[o] = (unsigned char)((r << 4) + lc);
            a.sideflag[j] = 1;
        }
        break;
    }
    case TILESPMV_FMT_DENSE:
        a.Blockdense_Val[a.dns_offset[t] + lc * rowlen + r] = v;
        break;
    case TILESPMV_FMT_DENSEROW:
        a.Blockdenserow_Val[a.dnsrow_offset[t] + k] = v;
        break;
    case TILESPMV_FMT_DENSECOL:
        a.Blockdensecol_Val[a.dnscol_offset[t] + kr * rowlen + r] = v;
        if (r == 0) // dense-column ids in order of appearance in local row 0 (:600-606)
            a.densecolid[a.dnscolptr[t] + kr] = (char)lc;
        break;
    default:
        break;
    }
}

// per-tile leftovers: CSR in-tile row pointers, DenseRow row ids
__global__ void __launch_bounds__(128)
    tile_post_kernel(int T, const char *__restrict__ Format, const unsigned char *__restrict__ rowstart,
                     const int *__restrict__ tile_nnz, const int *__restrict__ tile_br,
                     const int *__restrict__ tile_columnidx, int tilem, int tilen, int rowA, int colA,
                     const int *__restrict__ csrptr_offset, const int *__restrict__ dnsrowptr,
                     unsigned char *__restrict__ Blockcsr_Ptr, char *__restrict__ denserowid)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T)
        return;
    const int fmt = Format[t];
    if (fmt != TILESPMV_FMT_CSR && fmt != TILESPMV_FMT_DENSEROW)
        return;
    const int rowlen = tile_br[t] == tilem - 1 ? rowA - (tilem - 1) * TS : TS;
    const unsigned char *rs = rowstart + (size_t)t * TS;
    if (fmt == TILESPMV_FMT_CSR)
    {
        const int o = csrptr_offset[t];
        for (int r = 0; r < rowlen; r++)
            Blockcsr_Ptr[o + r] = rs[r];
    }
    else
    {
        const int nnz = tile_nnz[t + 1] - tile_nnz[t];
        const int collen = tile_columnidx[t] == tilen - 1 ? colA - (tilen - 1) * TS : TS;
        int o = dnsrowptr[t];
        for (int r = 0; r < rowlen; r++)
        {
            int end = r == rowlen - 1 ? nnz : rs[r + 1];
            if (end - rs[r] == collen)
                denserowid[o++] = (char)r;
        }
    }
}

// 6. two 4-bit indices per byte, parity by global position (encode.h:29-50 over the whole array)
__global__ void __launch_bounds__(CV_THREADS)
    pack_nibbles_kernel(const unsigned char *__restrict__ idx, int len, unsigned char *__restrict__ out)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * b >= len)
        return;
    unsigned hi = idx[2 * b];
    unsigned lo = 2 * b + 1 < len ? idx[2 * b + 1] : 0u;
    out[b] = (unsigned char)((hi << 4) + lo);
}

// HYB tiles pack their indices tile by tile (csr2tile.h:984-1008): ceil(w*rowlen/2) nibble bytes of the ELL part
// (parity by position INSIDE the tile), then one (row << 4) + col byte per spilled entry.  One thread per tile.
__global__ void __launch_bounds__(128)
    pack_hyb_kernel(int T, const char *__restrict__ Format, const int *__restrict__ hyb_offset,
                    const int *__restrict__ hyb_coocount, const int *__restrict__ hyb_idxoff,
                    const unsigned char *__restrict__ lc, unsigned char *__restrict__ out)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || Format[t] != TILESPMV_FMT_HYB)
        return;
    const int base = hyb_offset[t];
    const int spill = hyb_coocount[t + 1] - hyb_coocount[t];
    const int ell = hyb_offset[t + 1] - base - spill;
    unsigned char *o = out + hyb_idxoff[t];
    for (int b = 0; 2 * b < ell; b++)
    {
        unsigned hi = lc[base + 2 * b];
        unsigned lo = 2 * b + 1 < ell ? lc[base + 2 * b + 1] : 0u;
        o[b] = (unsigned char)((hi << 4) + lo);
    }
    o += (ell + 1) / 2;
    for (int i = 0; i < spill; i++)
        o[i] = lc[base + ell + i];
}

// ---------------------------------------------------------------------------------------------
// 7. side CSR of the COO tiles (csr2tile.h:899-960)
// ---------------------------------------------------------------------------------------------
struct SideFlagIn
{
    const unsigned char *f;
    size_t n;
    __device__ __forceinline__ int operator()(size_t j) const { return j < n ? (int)f[j] : 0; }
};

template <class T>
__global__ void __launch_bounds__(CV_THREADS)
    side_fill_kernel(const unsigned char *__restrict__ flag, const int *__restrict__ sidepos, size_t n,
                     const int *__restrict__ colidx, const T *__restrict__ val, int *__restrict__ out_col,
                     T *__restrict__ out_val)
{
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || !flag[j])
        return;
    int q = sidepos[j];
    out_col[q] = colidx[j];
    out_val[q] = val[j];
}

__global__ void __launch_bounds__(CV_THREADS)
    side_ptr_kernel(const int *__restrict__ rowptr, const int *__restrict__ sidepos, int rowA, int *__restrict__ out_ptr)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > rowA)
        return;
    out_ptr[i] = sidepos[rowptr[i]]; // sidepos has n+1 entries
}

// sets *unsorted when a side row is not ascending (equal neighbours count as sorted)
__global__ void __launch_bounds__(CV_THREADS)
    side_check_sorted_kernel(const int *__restrict__ side_ptr, const int *__restrict__ side_col, int rowA, int total,
                             int *__restrict__ unsorted)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total || q == 0)
        return;
    if (side_col[q] < side_col[q - 1])
    {
        int row = upper_row(side_ptr, rowA + 1, q);
        if (q > side_ptr[row])
            *unsorted = 1;
    }
}

__global__ void __launch_bounds__(CV_THREADS)
    side_keys_kernel(const int *__restrict__ side_ptr, const int *__restrict__ side_col, int rowA, int total, int colbits,
                     uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total)
        return;
    int row = upper_row(side_ptr, rowA + 1, q);
    keys[q] = ((uint64_t)row << colbits) | (uint64_t)side_col[q];
    vals[q] = (uint32_t)q;
}

template <class T>
__global__ void __launch_bounds__(CV_THREADS)
    side_permute_kernel(const uint32_t *__restrict__ perm, int total, const int *__restrict__ col_in,
                        const T *__restrict__ val_in, int *__restrict__ col_out, T *__restrict__ val_out)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total)
        return;
    uint32_t s = perm[q];
    col_out[q] = col_in[s];
    val_out[q] = val_in[s];
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
static int read_last_int(const DevBuf &b, int idx, int *out, cudaStream_t s)
{
    TSP_CUDA(cudaMemcpyAsync(out, b.as<int>() + idx, sizeof(int), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaStreamSynchronize(s));
    return TILESPMV_OK;
}

template <class T>
int convert_csr_to_tiles(int rowA, int colA, const int *d_rowptr, const int *d_colidx, const T *d_val,
                         tilespmv_dmat *M, cudaStream_t s, bool enable_hyb)
{
    if (rowA < 0 || colA < 0)
    {
        set_error("convert: negative dimensions");
        return TILESPMV_ERR_INVALID;
    }
    M->precision = (int)sizeof(T);
    M->rowA = rowA;
    M->colA = colA;
    const int tilem = (rowA + TS - 1) / TS, tilen = (colA + TS - 1) / TS;
    M->tilem = tilem;
    M->tilen = tilen;

    int nnz_i = 0;
    if (rowA > 0)
    {
        TSP_CUDA(cudaMemcpyAsync(&nnz_i, d_rowptr + rowA, sizeof(int), cudaMemcpyDeviceToHost, s));
        TSP_CUDA(cudaStreamSynchronize(s));
    }
    if (nnz_i < 0)
    {
        set_error("convert: rowptr[rowA] is negative (int overflow?)");
        return TILESPMV_ERR_INVALID;
    }
    const size_t n = (size_t)nnz_i;
    M->nnz = (int64_t)n;
    ScanWorkspace ws;

    // ---- 1. keys + stable sort by (block row, tile column) ----
    const int tcbits = bits_for(tilen > 0 ? (uint64_t)(tilen - 1) : 0);
    const int brbits = bits_for(tilem > 0 ? (uint64_t)(tilem - 1) : 0);
    DevBuf keys_a, keys_b, vals_a, vals_b;
    uint64_t *K = nullptr;
    uint32_t *V = nullptr;
    if (n)
    {
        TSP_TRY(keys_a.alloc(n * 8, false));
        TSP_TRY(keys_b.alloc(n * 8, false));
        TSP_TRY(vals_a.alloc(n * 4, false));
        TSP_TRY(vals_b.alloc(n * 4, false));
        TSP_LAUNCH(make_keys_kernel, grid_for(n, CV_THREADS), CV_THREADS, 0, s, d_rowptr, d_colidx, rowA, n, tcbits,
                   keys_a.as<uint64_t>(), vals_a.as<uint32_t>());
        TSP_TRY(radix_sort_pairs(keys_a.as<uint64_t>(), vals_a.as<uint32_t>(), keys_b.as<uint64_t>(),
                                 vals_b.as<uint32_t>(), n, 4, 4 + tcbits + brbits, ws, s, &K, &V));
    }

    // ---- 2. tile ids and headers ----
    DevBuf headscan; // exclusive scan of the head flags, n+1 entries
    long long T_ll = 0;
    TSP_TRY(headscan.alloc((n + 1) * sizeof(int), false));
    TSP_TRY(exclusive_scan(HeadFlagIn{K, n}, n + 1, headscan.as<int>(), ws, s, &T_ll));
    const int NT = (int)T_ll;
    M->tilenum = NT;
    // The reference sizes its scratch tile_csr_ptr with the int product tilenum*BLOCK_SIZE (csr2tile.h:668) and breaks
    // beyond 2^31 / 16 = 134 M tiles; nothing in Tile_matrix itself needs that product, so this conversion carries on
    // (R-MAT scale 24 has 217 M tiles) -- every index here that involves tilenum * 16 is 64-bit.

    TSP_TRY(M->tile_ptr.alloc((size_t)(tilem + 1) * 4, true, s));
    TSP_TRY(M->tile_columnidx.alloc((size_t)NT * 4, true, s));
    TSP_TRY(M->tile_nnz.alloc((size_t)(NT + 1) * 4, true, s));
    DevBuf tile_br, rowstart, slots, nd;
    TSP_TRY(tile_br.alloc((size_t)NT * 4, false));
    TSP_TRY(rowstart.alloc((size_t)NT * TS, false));
    TSP_TRY(slots.alloc((size_t)(NT + 1) * 4, true, s));
    TSP_TRY(nd.alloc((size_t)(NT + 1) * 4, true, s));
    TSP_CUDA(cudaMemsetAsync(rowstart.p, 0xFF, rowstart.bytes, s));
    if (n)
    {
        TSP_LAUNCH(tile_fill_kernel, grid_for(n, CV_THREADS), CV_THREADS, 0, s, K, headscan.as<int>(), n, tcbits,
                   M->tile_columnidx.as<int>(), M->tile_nnz.as<int>(), tile_br.as<int>());
        TSP_CUDA(cudaMemcpyAsync(M->tile_nnz.as<int>() + NT, &nnz_i, sizeof(int), cudaMemcpyHostToDevice, s));
        TSP_LAUNCH(row_start_kernel, grid_for(n, CV_THREADS), CV_THREADS, 0, s, K, headscan.as<int>(), n,
                   M->tile_nnz.as<int>(), rowstart.as<unsigned char>());
    }
    TSP_LAUNCH(tile_ptr_kernel, grid_for((size_t)tilem + 1, CV_THREADS), CV_THREADS, 0, s, tile_br.as<int>(), NT, tilem,
               M->tile_ptr.as<int>());

    // ---- 3. format selection ----
    TSP_TRY(M->Format.alloc((size_t)NT, true, s));
    TSP_TRY(M->tilewidth.alloc((size_t)NT, true, s));
    if (NT)
        TSP_LAUNCH(classify_kernel, grid_for((size_t)NT, 128), 128, 0, s, NT, M->tile_nnz.as<int>(), tile_br.as<int>(),
                   M->tile_columnidx.as<int>(), rowstart.as<unsigned char>(), V, d_colidx, tilem, tilen, rowA, colA,
                   enable_hyb ? (int)sizeof(T) : 0, M->Format.as<char>(), slots.as<int>(), M->tilewidth.as<char>(), nd.as<int>());

    // ---- 4. prefix offsets (NT+1 entries each, last = total), blknnznnz, format histogram ----
    TSP_TRY(M->blknnznnz.alloc((size_t)NT + 1, true, s));
    DevBuf hist;
    TSP_TRY(hist.alloc(8 * sizeof(unsigned long long), true, s));
    TSP_LAUNCH(blknnznnz_kernel, grid_for((size_t)NT + 1, CV_THREADS), CV_THREADS, 0, s, slots.as<int>(), NT,
               M->blknnznnz.as<unsigned char>(), M->Format.as<char>(), hist.as<unsigned long long>());
    DevBuf hyb_idxoff; // byte offset of every HYB tile inside hybIdx (conversion scratch)
    struct
    {
        DevBuf *buf;
        int kind;
    } offs[] = {{&M->blknnz, OFF_BLKNNZ},        {&M->csr_offset, OFF_CSR},         {&M->csrptr_offset, OFF_CSRPTR},
                {&M->coo_offset, OFF_COO},       {&M->ell_offset, OFF_ELL},         {&M->hyb_offset, enable_hyb ? OFF_HYB : OFF_ZERO},
                {&M->hyb_coocount, enable_hyb ? OFF_HYBCOO : OFF_ZERO},             {&M->dns_offset, OFF_DNS},
                {&M->dnsrow_offset, OFF_DNSROW}, {&M->dnscol_offset, OFF_DNSCOL},   {&M->dnsrowptr, OFF_DNSROWPTR},
                {&M->dnscolptr, OFF_DNSCOLPTR},  {&M->new_coocount, OFF_NEWCOO},    {&hyb_idxoff, enable_hyb ? OFF_HYBIDX : OFF_ZERO}};
    for (auto &o : offs)
    {
        TSP_TRY(o.buf->alloc((size_t)(NT + 1) * 4, true, s));
        if (o.kind == OFF_ZERO || NT == 0)
            continue;
        OffsetIn in{M->Format.as<char>(), slots.as<int>(), nd.as<int>(), tile_br.as<int>(), NT, tilem, rowA, o.kind};
        long long tot = 0;
        TSP_TRY(exclusive_scan(in, (size_t)NT + 1, static_cast<int *>(o.buf->p), ws, s, &tot)); // also guards int overflow
    }
    int hyb_idx_bytes = 0;
    if (NT)
    {
        TSP_TRY(read_last_int(M->csr_offset, NT, &M->csrsize, s));
        TSP_TRY(read_last_int(M->csrptr_offset, NT, &M->csrptrlen, s));
        TSP_TRY(read_last_int(M->coo_offset, NT, &M->coosize, s));
        TSP_TRY(read_last_int(M->ell_offset, NT, &M->ellsize, s));
        TSP_TRY(read_last_int(M->dns_offset, NT, &M->dnssize, s));
        TSP_TRY(read_last_int(M->dnsrow_offset, NT, &M->dnsrowsize, s));
        TSP_TRY(read_last_int(M->dnscol_offset, NT, &M->dnscolsize, s));
        TSP_TRY(read_last_int(M->dnsrowptr, NT, &M->ndenserowid, s));
        TSP_TRY(read_last_int(M->dnscolptr, NT, &M->ndensecolid, s));
        TSP_TRY(read_last_int(M->new_coocount, NT, &M->coototal, s));
        if (enable_hyb)
        {
            // hybsize / hybellsize / hybcoosize of the size pass (:752, :776-779)
            TSP_TRY(read_last_int(M->hyb_offset, NT, &M->hybsize, s));
            TSP_TRY(read_last_int(M->hyb_coocount, NT, &M->hybcoosize, s));
            TSP_TRY(read_last_int(hyb_idxoff, NT, &hyb_idx_bytes, s));
            M->hybellsize = M->hybsize - M->hybcoosize;
        }
    }
    {
        unsigned long long h[8];
        TSP_CUDA(cudaMemcpyAsync(h, hist.p, sizeof(h), cudaMemcpyDeviceToHost, s));
        TSP_CUDA(cudaStreamSynchronize(s));
        for (int f = 0; f < 7; f++)
            M->fmt_hist[f] = (int64_t)h[f];
    }

    // ---- 5. per-format storage + scatter ----
    const size_t vs = sizeof(T);
    TSP_TRY(M->Blockcsr_Val.alloc((size_t)M->csrsize * vs, true, s));
    TSP_TRY(M->Blockcsr_Ptr.alloc((size_t)M->csrptrlen, true, s));
    TSP_TRY(M->csr_compressedIdx.alloc((size_t)(M->csrsize + 1) / 2, true, s));
    TSP_TRY(M->Blockcoo_Val.alloc((size_t)M->coosize * vs, true, s));
    TSP_TRY(M->coo_compressed_Idx.alloc((size_t)M->coosize, true, s));
    TSP_TRY(M->Blockell_Val.alloc((size_t)M->ellsize * vs, true, s));
    TSP_TRY(M->ell_compressedIdx.alloc((size_t)(M->ellsize + 1) / 2, true, s));
    // hybIdx: the reference sizes it ceil(hybellsize/2) + hybcoosize (:840-841) but fills it tile by tile with
    // per-tile rounding (:994-1004), which is longer when several HYB tiles of the ragged last block row have an
    // odd ELL part (a heap overrun upstream): allocate the larger of the two, export the reference's length
    const size_t hyb_ref_bytes = ((size_t)M->hybellsize + 1) / 2 + (size_t)M->hybcoosize;
    TSP_TRY(M->Blockhyb_Val.alloc((size_t)M->hybsize * vs, true, s));
    TSP_TRY(M->hybIdx.alloc(std::max(hyb_ref_bytes, (size_t)hyb_idx_bytes), true, s));
    TSP_TRY(M->Blockdense_Val.alloc((size_t)M->dnssize * vs, true, s));
    TSP_TRY(M->Blockdenserow_Val.alloc((size_t)M->dnsrowsize * vs, true, s));
    TSP_TRY(M->denserowid.alloc((size_t)M->ndenserowid, true, s));
    TSP_TRY(M->Blockdensecol_Val.alloc((size_t)M->dnscolsize * vs, true, s));
    TSP_TRY(M->densecolid.alloc((size_t)M->ndensecolid, true, s));
    DevBuf csr_lc, ell_lc, hyb_lc, sideflag;
    TSP_TRY(hyb_lc.alloc((size_t)M->hybsize, true, s));
    TSP_TRY(csr_lc.alloc((size_t)M->csrsize, true, s));
    TSP_TRY(ell_lc.alloc((size_t)M->ellsize, true, s));
    TSP_TRY(sideflag.alloc(n + 1, true, s));
    if (n)
    {
        ScatterArgs<T> a;
        a.keys = K;
        a.perm = V;
        a.headscan = headscan.as<int>();
        a.n = n;
        a.colidx = d_colidx;
        a.val = d_val;
        a.tilem = tilem;
        a.rowA = rowA;
        a.tile_nnz = M->tile_nnz.as<int>();
        a.tile_br = tile_br.as<int>();
        a.Format = M->Format.as<char>();
        a.rowstart = rowstart.as<unsigned char>();
        a.csr_offset = M->csr_offset.as<int>();
        a.coo_offset = M->coo_offset.as<int>();
        a.ell_offset = M->ell_offset.as<int>();
        a.hyb_offset = M->hyb_offset.as<int>();
        a.tilewidth = M->tilewidth.as<char>();
        a.Blockhyb_Val = M->Blockhyb_Val.as<T>();
        a.hyb_lc = hyb_lc.as<unsigned char>();
        a.dns_offset = M->dns_offset.as<int>();
        a.dnsrow_offset = M->dnsrow_offset.as<int>();
        a.dnscol_offset = M->dnscol_offset.as<int>();
        a.dnscolptr = M->dnscolptr.as<int>();
        a.Blockcsr_Val = M->Blockcsr_Val.as<T>();
        a.Blockcoo_Val = M->Blockcoo_Val.as<T>();
        a.Blockell_Val = M->Blockell_Val.as<T>();
        a.Blockdense_Val = M->Blockdense_Val.as<T>();
        a.Blockdenserow_Val = M->Blockdenserow_Val.as<T>();
        a.Blockdensecol_Val = M->Blockdensecol_Val.as<T>();
        a.csr_lc = csr_lc.as<unsigned char>();
        a.ell_lc = ell_lc.as<unsigned char>();
        a.coo_idx = M->coo_compressed_Idx.as<unsigned char>();
        a.densecolid = M->densecolid.as<char>();
        a.sideflag = sideflag.as<unsigned char>();
        TSP_LAUNCH((scatter_kernel<T>), grid_for(n, CV_THREADS), CV_THREADS, 0, s, a);
        TSP_LAUNCH(tile_post_kernel, grid_for((size_t)NT, 128), 128, 0, s, NT, M->Format.as<char>(),
                   rowstart.as<unsigned char>(), M->tile_nnz.as<int>(), tile_br.as<int>(),
                   M->tile_columnidx.as<int>(), tilem, tilen, rowA, colA, M->csrptr_offset.as<int>(),
                   M->dnsrowptr.as<int>(), M->Blockcsr_Ptr.as<unsigned char>(), M->denserowid.as<char>());
    }
    // ---- 6. nibble packing ----
    if (M->csrsize)
        TSP_LAUNCH(pack_nibbles_kernel, grid_for((size_t)(M->csrsize + 1) / 2, CV_THREADS), CV_THREADS, 0, s,
                   csr_lc.as<unsigned char>(), M->csrsize, M->csr_compressedIdx.as<unsigned char>());
    if (M->ellsize)
        TSP_LAUNCH(pack_nibbles_kernel, grid_for((size_t)(M->ellsize + 1) / 2, CV_THREADS), CV_THREADS, 0, s,
                   ell_lc.as<unsigned char>(), M->ellsize, M->ell_compressedIdx.as<unsigned char>());
    if (M->hybsize)
        TSP_LAUNCH(pack_hyb_kernel, grid_for((size_t)NT, 128), 128, 0, s, NT, M->Format.as<char>(), M->hyb_offset.as<int>(),
                   M->hyb_coocount.as<int>(), hyb_idxoff.as<int>(), hyb_lc.as<unsigned char>(), M->hybIdx.as<unsigned char>());

    // ---- 7. side CSR: compaction of the COO-tile nonzeros in original CSR order ----
    TSP_TRY(M->deferredcoo_ptr.alloc((size_t)(rowA + 1) * 4, true, s));
    TSP_TRY(M->deferredcoo_colidx.alloc((size_t)M->coototal * 4, true, s));
    TSP_TRY(M->deferredcoo_val.alloc((size_t)M->coototal * vs, true, s));
    if (M->coototal)
    {
        // headscan is dead now: reuse it for the compaction offsets (n+1 entries)
        long long side_total = 0;
        TSP_TRY(exclusive_scan(SideFlagIn{sideflag.as<unsigned char>(), n}, n + 1, headscan.as<int>(), ws, s, &side_total));
        if (side_total != (long long)M->coototal)
        {
            set_error("convert: side-matrix count mismatch (%lld vs %d)", side_total, M->coototal);
            return TILESPMV_ERR_CUDA;
        }
        TSP_LAUNCH((side_fill_kernel<T>), grid_for(n, CV_THREADS), CV_THREADS, 0, s, sideflag.as<unsigned char>(),
                   headscan.as<int>(), n, d_colidx, d_val, M->deferredcoo_colidx.as<int>(), M->deferredcoo_val.as<T>());
        TSP_LAUNCH(side_ptr_kernel, grid_for((size_t)rowA + 1, CV_THREADS), CV_THREADS, 0, s, d_rowptr,
                   headscan.as<int>(), rowA, M->deferredcoo_ptr.as<int>());
        DevBuf unsorted;
        TSP_TRY(unsorted.alloc(sizeof(int), true, s));
        TSP_LAUNCH(side_check_sorted_kernel, grid_for((size_t)M->coototal, CV_THREADS), CV_THREADS, 0, s,
                   M->deferredcoo_ptr.as<int>(), M->deferredcoo_colidx.as<int>(), rowA, M->coototal, unsorted.as<int>());
        int h_unsorted = 0;
        TSP_CUDA(cudaMemcpyAsync(&h_unsorted, unsorted.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        TSP_CUDA(cudaStreamSynchronize(s));
        if (h_unsorted)
        {
            // rows must come out ascending by column (quick_sort_key_val_pair, csr2tile.h:952-960);
            // stable sort by (row, column): ties (duplicate entries) keep their input order
            const size_t q = (size_t)M->coototal;
            const int colbits = bits_for(colA > 0 ? (uint64_t)(colA - 1) : 0);
            const int rowbits = bits_for(rowA > 0 ? (uint64_t)(rowA - 1) : 0);
            // keys_a/b and vals_a/b are large enough (coototal <= n)
            TSP_LAUNCH(side_keys_kernel, grid_for(q, CV_THREADS), CV_THREADS, 0, s, M->deferredcoo_ptr.as<int>(),
                       M->deferredcoo_colidx.as<int>(), rowA, M->coototal, colbits, keys_a.as<uint64_t>(),
                       vals_a.as<uint32_t>());
            uint64_t *K2;
            uint32_t *V2;
            TSP_TRY(radix_sort_pairs(keys_a.as<uint64_t>(), vals_a.as<uint32_t>(), keys_b.as<uint64_t>(),
                                     vals_b.as<uint32_t>(), q, 0, colbits + rowbits, ws, s, &K2, &V2));
            DevBuf col2, val2;
            TSP_TRY(col2.alloc(q * 4, false));
            TSP_TRY(val2.alloc(q * vs, false));
            TSP_LAUNCH((side_permute_kernel<T>), grid_for(q, CV_THREADS), CV_THREADS, 0, s, V2, M->coototal,
                       M->deferredcoo_colidx.as<int>(), M->deferredcoo_val.as<T>(), col2.as<int>(), val2.as<T>());
            TSP_CUDA(cudaStreamSynchronize(s));
            M->deferredcoo_colidx = std::move(col2);
            M->deferredcoo_val = std::move(val2);
        }
    }
    TSP_CUDA(cudaStreamSynchronize(s));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
    {
        set_error("convert: kernel failure: %s", cudaGetErrorString(e));
        return TILESPMV_ERR_CUDA;
    }
    return TILESPMV_OK;
}

template int convert_csr_to_tiles<double>(int, int, const int *, const int *, const double *, tilespmv_dmat *, cudaStream_t, bool);
template int convert_csr_to_tiles<float>(int, int, const int *, const int *, const float *, tilespmv_dmat *, cudaStream_t, bool);

} // namespace tsp

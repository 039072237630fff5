/*
 * oracle/tilespmv_oracle.c -- CPU RESTATEMENT of the TileSpMV hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (tilespmv_b200/, include/) may include,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg use it, and there only as the checker / reported CPU baseline.
 *
 * Parity status: PINNED.  The reference has no golden vectors of its own, so this
 * restatement is pinned against the unmodified reference CPU path compiled from
 * /root/reference/src into oracle/_ref/ (oracle/ref_shim.c): tests/test_oracle_vs_ref.py compares
 * every Tile_matrix array, ptroffset1/2, the warp-chunk schedule and y byte for byte on all the
 * inputs of SURVEY.md Appendix C.2/C.3, and tests/golden/ holds small committed fixtures produced
 * by that reference build (tests/golden/make_golden.py).
 *
 * This is NOT a copy of the reference: the reference discovers tiles with O(tilem*tilen) dense
 * scratch per block row (csr2tile.h:5-106) and a per-nnz linear tile search (csr2tile.h:403-419);
 * here tiles are discovered by one stable sort of each block row's nonzeros by tile column, which
 * is O(nnz log nnz) overall and therefore also usable for the BASELINE configs the reference
 * converter cannot finish.  What is restated faithfully is the OBSERVABLE behaviour:
 *
 *   - Tile_matrix layout                        format.h:3-56       -> oracle_tile_matrix
 *   - tile discovery (ptr / columnidx / nnz)    csr2tile.h:5-106    -> group_blockrow() + header pass
 *   - per-tile format selection                 csr2tile.h:141-326  -> classify_tile()
 *   - prefix offsets, sizes, blknnznnz wrap     csr2tile.h:729-799  -> oracle_tile_create()
 *   - per-format value / index layouts          csr2tile.h:427-621  -> scatter pass of oracle_tile_create()
 *   - deferred-COO side CSR                     csr2tile.h:899-960  -> side-CSR pass + sort_side_row()
 *   - 4-bit packing (global-position parity)    encode.h:29-50, csr2tile.h:973-982 -> pack_nibbles()
 *   - CPU tile SpMV (the y oracle), ptroffset   tilespmv_cpu.h:125-272 -> oracle_tilespmv_cpu()
 *   - warp-chunk schedule                       tilespmv_cpu.h:68-118  -> oracle_build_schedule()
 *   - plain CSR SpMV (y_golden)                 main.cu:101-110        -> oracle_csr_spmv()
 *   - Matrix Market -> CSR semantics            mmio_highlevel.h:593-759 -> oracle_mtx_read()
 *
 * Build: oracle/Makefile compiles this twice (-DORACLE_VAL_TYPE=double / float); no FMA
 * contraction (-ffp-contract=off) so the floating-point statistics of classify_tile() and the
 * sums of oracle_tilespmv_cpu() round exactly like the reference's x86-64 build.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#ifndef ORACLE_VAL_TYPE
#define ORACLE_VAL_TYPE double
#endif
typedef ORACLE_VAL_TYPE val_t;

#define TS 16          /* BLOCK_SIZE,        common.h:37-39 */
#define COO_TH 12      /* COO_NNZ_TH,        common.h:45-47 */
#define CHUNK_TILES 4  /* PREFETCH_SMEM_TH,  common.h:49-51 */

/* field-for-field mirror of format.h:3-56 (same order, same C types) */
typedef struct
{
    int tilem, tilen, tilenum;
    int *tile_ptr;
    int *tile_columnidx;
    int *tile_nnz;
    char *Format;
    int *blknnz;
    unsigned char *blknnznnz;
    int *dnsrowptr;
    int *dnscolptr;
    char *tilewidth;
    int *csr_offset;
    int *csrptr_offset;
    int *coo_offset;
    int *ell_offset;
    int *hyb_offset;
    int *hyb_coocount;
    int *dns_offset;
    int *dnsrow_offset;
    int *dnscol_offset;
    int *new_coocount;
    val_t *Blockcsr_Val;
    unsigned char *Blockcsr_Ptr;
    unsigned char *csr_compressedIdx;
    int csrsize;
    int csrptrlen;
    val_t *Blockcoo_Val;
    unsigned char *coo_compressed_Idx;
    int coosize;
    val_t *Blockell_Val;
    unsigned char *ell_compressedIdx;
    int ellsize;
    val_t *Blockhyb_Val;
    unsigned char *hybIdx;
    int hybsize;
    int hybellsize;
    int hybcoosize;
    val_t *Blockdense_Val;
    int dnssize;
    val_t *Blockdenserow_Val;
    char *denserowid;
    int dnsrowsize;
    val_t *Blockdensecol_Val;
    char *densecolid;
    int dnscolsize;
    int coototal;
    val_t *deferredcoo_val;
    int *deferredcoo_colidx;
    int *deferredcoo_ptr;
} oracle_tile_matrix;

int oracle_sizeof_val(void) { return (int)sizeof(val_t); }
int oracle_sizeof_tile_matrix(void) { return (int)sizeof(oracle_tile_matrix); }
int oracle_omp_max_threads(void) { return omp_get_max_threads(); }

static void *zalloc(size_t n, size_t sz)
{
    void *p = calloc(n ? n : 1, sz);
    if (!p)
    {
        fprintf(stderr, "oracle: out of memory (%zu x %zu)\n", n, sz);
        abort();
    }
    return p;
}

/* in-place exclusive prefix sum, last slot receives the total (utils.h:34-49 semantics) */
static void prefix_excl(int *a, int len)
{
    int run = 0;
    for (int i = 0; i < len; i++)
    {
        int v = a[i];
        a[i] = run;
        run += v;
    }
}

static inline int rows_in_blockrow(int blki, int tilem, int rowA)
{
    return blki == tilem - 1 ? rowA - (tilem - 1) * TS : TS;
}
static inline int cols_in_tilecol(int tc, int tilen, int colA)
{
    return tc == tilen - 1 ? colA - (tilen - 1) * TS : TS;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y);
}

/*
 * Stable grouping of one block row's nonzeros by tile column.  The reference keeps, inside a
 * tile, row-major order and inside a row the original CSR order (csr2tile.h:403-419); sorting
 * the pairs (tile column, sequence number in the block row) reproduces exactly that order.
 * perm[] receives original nnz indices; returns the number of distinct tile columns.
 */
static int group_blockrow(int nnz_lo, int nnz_hi, const int *colidx, uint64_t *keys, int *perm)
{
    int n = nnz_hi - nnz_lo;
    int sorted = 1;
    for (int s = 0; s < n; s++)
    {
        uint64_t tc = (uint64_t)(colidx[nnz_lo + s] / TS);
        keys[s] = (tc << 32) | (uint32_t)s;
        if (s && keys[s] < keys[s - 1])
            sorted = 0;
    }
    if (!sorted)
        qsort(keys, n, sizeof(uint64_t), cmp_u64);
    int ntiles = 0;
    for (int s = 0; s < n; s++)
    {
        perm[s] = nnz_lo + (int)(uint32_t)keys[s];
        if (s == 0 || (keys[s] >> 32) != (keys[s - 1] >> 32))
            ntiles++;
    }
    return ntiles;
}

/* row (0..15) of an original nnz index inside block row blki */
static inline int local_row_of(int j, const int *rowptr, int row0, int rowlen)
{
    int lo = 0, hi = rowlen; /* find r with rowptr[row0+r] <= j < rowptr[row0+r+1] */
    while (hi - lo > 1)
    {
        int mid = (lo + hi) >> 1;
        if (rowptr[row0 + mid] <= j)
            lo = mid;
        else
            hi = mid;
    }
    /* skip empty rows that share the same pointer */
    while (rowptr[row0 + lo + 1] <= j)
        lo++;
    return lo;
}

/*
 * Per-tile format selection, csr2tile.h:141-326 (evaluation order of SURVEY.md A.2).
 * cnt[] = nnz per local row, ccnt[] = nnz per local column.
 * Outputs: format code, stored slots, ELL width, #dense rows, #dense cols.
 */
/* 0 (default) = the reference as shipped: format 3 never chosen.  1 = the dormant rule of csr2tile.h:279-316
   (commented out upstream) switched on; pinned against oracle/_ref/libtilespmv_refhyb_*.so. */
static int g_enable_hyb = 0;
void oracle_set_enable_hyb(int on) { g_enable_hyb = on; }

static void classify_tile(int nnz, const unsigned char *cnt, const unsigned char *ccnt, int rowlen,
                          int collen, int *fmt, int *slots, int *width, int *ndr, int *ndc, int *spill)
{
    *spill = 0;
    *width = 0;
    *ndr = 0;
    *ndc = 0;
    int dense_th = (int)(rowlen * collen * 0.75); /* int truncation of the product, :150 */
    if (nnz >= dense_th)
    {
        *fmt = 4;
        *slots = rowlen * collen;
        return;
    }
    if (nnz <= COO_TH)
    {
        *fmt = 1;
        *slots = nnz;
        return;
    }
    if (nnz % collen == 0 || nnz % rowlen == 0)
    {
        /* DenseRow first (:175-198): every row empty or full, the LAST decisive row wins the flag */
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
            *fmt = 5;
            *ndr = num;
            *slots = num * collen;
            return;
        }
        /* DenseCol (:201-240) */
        flag = 0;
        num = 0;
        for (int c = 0; c < collen; c++)
        {
            if (ccnt[c] % rowlen != 0)
            {
                flag = 0;
                break;
            }
            if (ccnt[c] == rowlen)
            {
                flag = 1;
                num++;
            }
        }
        if (flag)
        {
            *fmt = 6;
            *ndc = num;
            *slots = num * rowlen;
            return;
        }
    }
    /* ELL vs CSR by coefficient of variation of the row lengths (:245-276), all in double,
       accumulated in row order like the reference so the rounding is identical */
    int wmax = 0;
    for (int r = 0; r < rowlen; r++)
        if (cnt[r] > wmax)
            wmax = cnt[r];
    double mean = ((double)nnz) / rowlen;
    double var = 0.0;
    for (int r = 0; r < rowlen; r++)
    {
        int len = cnt[r];
        double d = (double)(len - mean);
        var += (d * d);
    }
    var /= rowlen;
    double sd = sqrt(var);
    double cv = sd / mean;
    if (cv <= 0.2)
    {
        *fmt = 2;
        *width = wmax;
        *slots = wmax * rowlen;
        return;
    }
    if (g_enable_hyb)
    {
        /* I/O-cost walk of :279-306: shrink the ELL width while bytes (value + nibble per ELL slot, value +
           index byte per spilled entry) keep going down; HYB when cv >= 1.0 and <= 4 entries spill (:308) */
        const int vs = (int)sizeof(val_t);
        int hybwidth = wmax, prior_spill = 0;
        int ioprior = wmax * rowlen * vs + (wmax * rowlen) / 2 + ((wmax * rowlen) % 2);
        for (int wi = wmax - 1; wi > 0; wi--)
        {
            int next_spill = 0;
            for (int r = 0; r < rowlen; r++)
                if (cnt[r] > wi)
                    next_spill += cnt[r] - wi;
            int ionext = wi * rowlen * vs + (wi * rowlen) / 2 + ((wi * rowlen) % 2) + next_spill * (vs + 1);
            if (ioprior <= ionext)
            {
                hybwidth = wi + 1;
                break;
            }
            hybwidth = wi;
            ioprior = ionext;
            prior_spill = next_spill;
        }
        if (cv >= 1.0 && prior_spill <= 4)
        {
            *fmt = 3;
            *width = hybwidth;
            *spill = prior_spill;
            *slots = prior_spill + hybwidth * rowlen;
            return;
        }
    }
    /* the HYB branch is commented out upstream (:308-316): everything else is CSR */
    *fmt = 0;
    *slots = nnz;
}

/* Lomuto partition around the FIRST element, restated from utils.h:102-136.  Only used for
   side-CSR rows that contain duplicate column indices, where the (unstable) tie order of the
   reference must be reproduced; distinct keys are sorted by any correct sort. */
static void ref_order_sort(int *key, val_t *val, int len)
{
    while (len > 1)
    {
        int pivot = key[0];
        int kt = key[0];
        key[0] = key[len - 1];
        key[len - 1] = kt;
        val_t vt = val[0];
        val[0] = val[len - 1];
        val[len - 1] = vt;
        int small = 0;
        for (int i = 0; i < len; i++)
        {
            if (key[i] < pivot)
            {
                kt = key[i];
                key[i] = key[small];
                key[small] = kt;
                vt = val[i];
                val[i] = val[small];
                val[small] = vt;
                small++;
            }
        }
        kt = key[len - 1];
        key[len - 1] = key[small];
        key[small] = kt;
        vt = val[len - 1];
        val[len - 1] = val[small];
        val[small] = vt;
        ref_order_sort(key, val, small);
        key += small + 1;
        val += small + 1;
        len -= small + 1;
    }
}

typedef struct
{
    int k;
    val_t v;
} kv_t;
static int cmp_kv(const void *a, const void *b)
{
    int x = ((const kv_t *)a)->k, y = ((const kv_t *)b)->k;
    return x < y ? -1 : (x > y);
}

static void sort_side_row(int *key, val_t *val, int len)
{
    int ascending = 1, distinct = 1;
    for (int i = 1; i < len; i++)
    {
        if (key[i] < key[i - 1])
            ascending = 0;
        if (key[i] == key[i - 1])
            distinct = 0;
    }
    if (ascending && distinct)
        return;
    if (distinct)
    {
        /* check for non-adjacent duplicates after a cheap sort of a copy */
        kv_t *tmp = (kv_t *)malloc(sizeof(kv_t) * len);
        for (int i = 0; i < len; i++)
        {
            tmp[i].k = key[i];
            tmp[i].v = val[i];
        }
        qsort(tmp, len, sizeof(kv_t), cmp_kv);
        int dup = 0;
        for (int i = 1; i < len; i++)
            if (tmp[i].k == tmp[i - 1].k)
                dup = 1;
        if (!dup)
        {
            for (int i = 0; i < len; i++)
            {
                key[i] = tmp[i].k;
                val[i] = tmp[i].v;
            }
            free(tmp);
            return;
        }
        free(tmp);
    }
    ref_order_sort(key, val, len);
}

/* two 4-bit indices per byte, parity by GLOBAL position in the per-format array
   (encode.h:29-50 called once over the whole array, csr2tile.h:973, :982) */
static void pack_nibbles(const unsigned char *idx, unsigned char *out, int len)
{
    for (int p = 0; p + 1 < len; p += 2)
        out[p >> 1] = (unsigned char)((idx[p] << 4) + idx[p + 1]);
    if (len & 1)
        out[len >> 1] = (unsigned char)(idx[len - 1] << 4);
}

/*
 * Tile_create restatement (csr2tile.h:629-1020).  Same contract: caller allocates the struct and
 * keeps the CSR arrays; every array in the struct is malloc'ed here.
 */
void oracle_tile_create(oracle_tile_matrix *M, int rowA, int colA, int nnzA, const int *rowptr,
                        const int *colidx, const val_t *val)
{
    (void)nnzA;
    memset(M, 0, sizeof(*M));
    const int tilem = (rowA + TS - 1) / TS, tilen = (colA + TS - 1) / TS;
    M->tilem = tilem;
    M->tilen = tilen;
    const int nnz_used = rowA > 0 ? rowptr[rowA] : 0; /* rows past rowA are ignored (main.cu:71) */

    /* ---- tile discovery: stable grouping per block row ---- */
    int *perm = (int *)zalloc(nnz_used, sizeof(int));
    M->tile_ptr = (int *)zalloc(tilem + 1, sizeof(int));
    int max_br_nnz = 0;
    for (int b = 0; b < tilem; b++)
    {
        int hi = b == tilem - 1 ? rowA : (b + 1) * TS;
        int n = rowptr[hi] - rowptr[b * TS];
        if (n > max_br_nnz)
            max_br_nnz = n;
    }
    const int nthreads = omp_get_max_threads();
    uint64_t *keys_all = (uint64_t *)zalloc((size_t)nthreads * (max_br_nnz + 1), sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < tilem; b++)
    {
        uint64_t *keys = keys_all + (size_t)omp_get_thread_num() * (max_br_nnz + 1);
        int lo = rowptr[b * TS];
        int hi = rowptr[b == tilem - 1 ? rowA : (b + 1) * TS];
        M->tile_ptr[b] = group_blockrow(lo, hi, colidx, keys, perm + lo);
    }
    free(keys_all);
    prefix_excl(M->tile_ptr, tilem + 1);
    const int T = M->tile_ptr[tilem];
    M->tilenum = T;

    M->tile_columnidx = (int *)zalloc(T, sizeof(int));
    M->tile_nnz = (int *)zalloc(T + 1, sizeof(int));
    M->Format = (char *)zalloc(T, 1);
    M->blknnz = (int *)zalloc(T + 1, sizeof(int));
    M->blknnznnz = (unsigned char *)zalloc(T + 1, 1);
    M->dnsrowptr = (int *)zalloc(T + 1, sizeof(int));
    M->dnscolptr = (int *)zalloc(T + 1, sizeof(int));
    M->tilewidth = (char *)zalloc(T, 1);
    M->csr_offset = (int *)zalloc(T + 1, sizeof(int));
    M->csrptr_offset = (int *)zalloc(T + 1, sizeof(int));
    M->coo_offset = (int *)zalloc(T + 1, sizeof(int));
    M->ell_offset = (int *)zalloc(T + 1, sizeof(int));
    M->hyb_offset = (int *)zalloc(T + 1, sizeof(int));
    M->hyb_coocount = (int *)zalloc(T + 1, sizeof(int));
    M->dns_offset = (int *)zalloc(T + 1, sizeof(int));
    M->dnsrow_offset = (int *)zalloc(T + 1, sizeof(int));
    M->dnscol_offset = (int *)zalloc(T + 1, sizeof(int));
    M->new_coocount = (int *)zalloc(T + 1, sizeof(int));
    /* per tile: 16 row counts, later turned into exclusive row starts */
    unsigned char *rowcnt = (unsigned char *)zalloc((size_t)T * TS, 1);

    /* ---- tile headers, per-row counts, format selection (one pass over the grouped order) ---- */
#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < tilem; b++)
    {
        const int rowlen = rows_in_blockrow(b, tilem, rowA);
        const int lo = rowptr[b * TS];
        const int hi = rowptr[b == tilem - 1 ? rowA : (b + 1) * TS];
        int t = M->tile_ptr[b] - 1;
        int prev_tc = -1;
        for (int p = lo; p < hi; p++)
        {
            int j = perm[p];
            int tc = colidx[j] / TS;
            if (p == lo || tc != prev_tc)
            {
                t++;
                M->tile_columnidx[t] = tc;
                M->tile_nnz[t] = p; /* exclusive prefix of true nnz == position in grouped order */
                prev_tc = tc;
            }
            rowcnt[(size_t)t * TS + local_row_of(j, rowptr, b * TS, rowlen)]++;
        }
    }
    M->tile_nnz[T] = nnz_used;

#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < tilem; b++)
    {
        const int rowlen = rows_in_blockrow(b, tilem, rowA);
        for (int t = M->tile_ptr[b]; t < M->tile_ptr[b + 1]; t++)
        {
            const int collen = cols_in_tilecol(M->tile_columnidx[t], tilen, colA);
            const int nnz = M->tile_nnz[t + 1] - M->tile_nnz[t];
            unsigned char ccnt[TS];
            memset(ccnt, 0, sizeof(ccnt));
            for (int p = M->tile_nnz[t]; p < M->tile_nnz[t + 1]; p++)
                ccnt[colidx[perm[p]] - M->tile_columnidx[t] * TS]++;
            int fmt, slots, width, ndr, ndc, spill;
            classify_tile(nnz, rowcnt + (size_t)t * TS, ccnt, rowlen, collen, &fmt, &slots, &width,
                          &ndr, &ndc, &spill);
            M->Format[t] = (char)fmt;
            M->blknnz[t] = slots;
            M->tilewidth[t] = (char)width;
            switch (fmt)
            {
            case 0:
                M->csr_offset[t] = slots;
                M->csrptr_offset[t] = rowlen;
                break;
            case 1:
                M->coo_offset[t] = slots;
                M->new_coocount[t] = slots;
                break;
            case 2:
                M->ell_offset[t] = slots;
                break;
            case 3: /* :309-316 */
                M->hyb_offset[t] = slots;
                M->hyb_coocount[t] = spill;
                M->new_coocount[t] = spill;
                break;
            case 4:
                M->dns_offset[t] = slots;
                break;
            case 5:
                M->dnsrow_offset[t] = slots;
                M->dnsrowptr[t] = ndr;
                break;
            case 6:
                M->dnscol_offset[t] = slots;
                M->dnscolptr[t] = ndc;
                break;
            }
        }
    }

    /* ---- totals, 8-bit wrapped slot counts (taken BEFORE the scan, :796-797), prefix offsets ---- */
    for (int b = 0; b < tilem; b++)
    {
        const int rowlen = rows_in_blockrow(b, tilem, rowA);
        for (int t = M->tile_ptr[b]; t < M->tile_ptr[b + 1]; t++)
        {
            switch (M->Format[t])
            {
            case 0:
                M->csrsize += M->blknnz[t];
                M->csrptrlen += rowlen;
                break;
            case 1:
                M->coosize += M->blknnz[t];
                break;
            case 2:
                M->ellsize += M->blknnz[t];
                break;
            case 3: /* :776-779 */
                M->hybsize += M->blknnz[t];
                M->hybellsize += M->tilewidth[t] * rowlen;
                break;
            case 4:
                M->dnssize += M->blknnz[t];
                break;
            case 5:
                M->dnsrowsize += M->blknnz[t];
                break;
            case 6:
                M->dnscolsize += M->blknnz[t];
                break;
            }
        }
    }
    for (int t = 0; t <= T; t++)
        M->blknnznnz[t] = (unsigned char)M->blknnz[t];
    int *scanned[] = {M->blknnz,       M->csr_offset,    M->csrptr_offset, M->coo_offset,
                      M->ell_offset,   M->hyb_offset,    M->dns_offset,    M->dnsrow_offset,
                      M->dnscol_offset, M->dnsrowptr,    M->dnscolptr,     M->hyb_coocount,
                      M->new_coocount};
    for (size_t a = 0; a < sizeof(scanned) / sizeof(scanned[0]); a++)
        prefix_excl(scanned[a], T + 1);
    M->hybcoosize = M->hyb_coocount[T];
    M->coototal = M->new_coocount[T];

    /* ---- per-format storage ---- */
    M->Blockcsr_Val = (val_t *)zalloc(M->csrsize, sizeof(val_t));
    M->Blockcsr_Ptr = (unsigned char *)zalloc(M->csrptrlen, 1);
    M->csr_compressedIdx = (unsigned char *)zalloc((M->csrsize + 1) / 2, 1);
    M->Blockcoo_Val = (val_t *)zalloc(M->coosize, sizeof(val_t));
    M->coo_compressed_Idx = (unsigned char *)zalloc(M->coosize, 1);
    M->Blockell_Val = (val_t *)zalloc(M->ellsize, sizeof(val_t));
    M->ell_compressedIdx = (unsigned char *)zalloc((M->ellsize + 1) / 2, 1);
    M->Blockhyb_Val = (val_t *)zalloc(M->hybellsize + M->hybcoosize, sizeof(val_t));
    /* the reference sizes hybIdx ceil(hybellsize/2) + hybcoosize (:840-841) but fills it with per-tile rounding
       (:994-1004): + one spare byte per tile row so that odd ELL parts in a ragged last block row stay in bounds */
    M->hybIdx = (unsigned char *)zalloc((M->hybellsize + 1) / 2 + M->hybcoosize + T + 1, 1);
    unsigned char *hyb_lc = (unsigned char *)zalloc(M->hybsize, 1);
    M->Blockdense_Val = (val_t *)zalloc(M->dnssize, sizeof(val_t));
    M->Blockdenserow_Val = (val_t *)zalloc(M->dnsrowsize, sizeof(val_t));
    M->denserowid = (char *)zalloc(M->dnsrowptr[T], 1);
    M->Blockdensecol_Val = (val_t *)zalloc(M->dnscolsize, sizeof(val_t));
    M->densecolid = (char *)zalloc(M->dnscolptr[T], 1);
    unsigned char *csr_lc = (unsigned char *)zalloc(M->csrsize, 1); /* unpacked local columns */
    unsigned char *ell_lc = (unsigned char *)zalloc(M->ellsize, 1);
    val_t *side_v = (val_t *)zalloc(M->coototal, sizeof(val_t)); /* triplets in tile order */
    int *side_r = (int *)zalloc(M->coototal, sizeof(int));
    int *side_c = (int *)zalloc(M->coototal, sizeof(int));

#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < tilem; b++)
    {
        const int rowlen = rows_in_blockrow(b, tilem, rowA);
        for (int t = M->tile_ptr[b]; t < M->tile_ptr[b + 1]; t++)
        {
            const int tc = M->tile_columnidx[t];
            const int collen = cols_in_tilecol(tc, tilen, colA);
            const int p0 = M->tile_nnz[t], n = M->tile_nnz[t + 1] - p0;
            unsigned char *rs = rowcnt + (size_t)t * TS; /* counts -> exclusive row starts */
            {
                int run = 0;
                for (int r = 0; r < TS; r++)
                {
                    int c = rs[r];
                    rs[r] = (unsigned char)run;
                    run += c;
                }
            }
            int r = 0;
            int ndr = 0, nspill = 0;
            for (int k = 0; k < n; k++)
            {
                while (r + 1 < rowlen && k >= rs[r + 1])
                    r++; /* row of the k-th nonzero of the tile */
                const int j = perm[p0 + k];
                const int lc = colidx[j] - tc * TS;
                const int kr = k - rs[r]; /* rank inside its row */
                switch (M->Format[t])
                {
                case 0:
                    M->Blockcsr_Val[M->csr_offset[t] + k] = val[j];
                    csr_lc[M->csr_offset[t] + k] = (unsigned char)lc;
                    break;
                case 1:
                    M->Blockcoo_Val[M->coo_offset[t] + k] = val[j];
                    M->coo_compressed_Idx[M->coo_offset[t] + k] = (unsigned char)((r << 4) + lc);
                    side_v[M->new_coocount[t] + k] = val[j];
                    side_r[M->new_coocount[t] + k] = b * TS + r;
                    side_c[M->new_coocount[t] + k] = tc * TS + lc;
                    break;
                case 2:
                    M->Blockell_Val[M->ell_offset[t] + kr * rowlen + r] = val[j];
                    ell_lc[M->ell_offset[t] + kr * rowlen + r] = (unsigned char)lc;
                    break;
                case 3: /* :505-548: ELL part slot-major, then the spilled entries in row order; the spilled
                           entries are ALSO handed to the side matrix (:538-545) */
                {
                    const int w = M->tilewidth[t], base = M->hyb_offset[t];
                    if (kr < w)
                    {
                        M->Blockhyb_Val[base + kr * rowlen + r] = val[j];
                        hyb_lc[base + kr * rowlen + r] = (unsigned char)lc;
                    }
                    else
                    {
                        M->Blockhyb_Val[base + w * rowlen + nspill] = val[j];
                        hyb_lc[base + w * rowlen + nspill] = (unsigned char)((r << 4) + lc);
                        side_v[M->new_coocount[t] + nspill] = val[j];
                        side_r[M->new_coocount[t] + nspill] = b * TS + r;
                        side_c[M->new_coocount[t] + nspill] = tc * TS + lc;
                        nspill++;
                    }
                    break;
                }
                case 4:
                    M->Blockdense_Val[M->dns_offset[t] + lc * rowlen + r] = val[j];
                    break;
                case 5:
                    M->Blockdenserow_Val[M->dnsrow_offset[t] + k] = val[j];
                    if (kr == 0)
                        M->denserowid[M->dnsrowptr[t] + ndr++] = (char)r;
                    break;
                case 6:
                    M->Blockdensecol_Val[M->dnscol_offset[t] + kr * rowlen + r] = val[j];
                    if (r == 0) /* dense-column ids in order of appearance in local row 0 */
                        M->densecolid[M->dnscolptr[t] + kr] = (char)lc;
                    break;
                }
            }
            if (M->Format[t] == 0)
                for (int rr = 0; rr < rowlen; rr++)
                    M->Blockcsr_Ptr[M->csrptr_offset[t] + rr] = rs[rr];
            (void)collen;
        }
    }

    pack_nibbles(csr_lc, M->csr_compressedIdx, M->csrsize);
    pack_nibbles(ell_lc, M->ell_compressedIdx, M->ellsize);
    /* HYB indices tile by tile (:984-1008): nibble parity restarts in every tile, spill bytes follow */
    {
        int o_idx = 0;
        for (int t = 0; t < T; t++)
        {
            if (M->Format[t] != 3)
                continue;
            const int spill = M->hyb_coocount[t + 1] - M->hyb_coocount[t];
            const int ell = M->hyb_offset[t + 1] - M->hyb_offset[t] - spill;
            pack_nibbles(hyb_lc + M->hyb_offset[t], M->hybIdx + o_idx, ell);
            o_idx += (ell + 1) / 2;
            memcpy(M->hybIdx + o_idx, hyb_lc + M->hyb_offset[t] + ell, spill);
            o_idx += spill;
        }
    }
    free(hyb_lc);

    /* ---- deferred-COO side CSR over global rows / cols (csr2tile.h:899-960) ---- */
    M->deferredcoo_val = (val_t *)zalloc(M->coototal, sizeof(val_t));
    M->deferredcoo_colidx = (int *)zalloc(M->coototal, sizeof(int));
    M->deferredcoo_ptr = (int *)zalloc(rowA + 1, sizeof(int));
    for (int q = 0; q < M->coototal; q++)
        M->deferredcoo_ptr[side_r[q]]++;
    prefix_excl(M->deferredcoo_ptr, rowA + 1);
    {
        int *fill = (int *)zalloc(rowA, sizeof(int));
        for (int q = 0; q < M->coototal; q++) /* tile order is preserved inside every row */
        {
            int dst = M->deferredcoo_ptr[side_r[q]] + fill[side_r[q]]++;
            M->deferredcoo_val[dst] = side_v[q];
            M->deferredcoo_colidx[dst] = side_c[q];
        }
        free(fill);
    }
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < rowA; i++)
        sort_side_row(M->deferredcoo_colidx + M->deferredcoo_ptr[i],
                      M->deferredcoo_val + M->deferredcoo_ptr[i],
                      M->deferredcoo_ptr[i + 1] - M->deferredcoo_ptr[i]);

    free(perm);
    free(rowcnt);
    free(csr_lc);
    free(ell_lc);
    free(side_v);
    free(side_r);
    free(side_c);
}

/* frees every array (the reference's Tile_destroy leaks four of them, format.h:58-94) */
void oracle_tile_destroy(oracle_tile_matrix *M)
{
    free(M->tile_ptr); free(M->tile_columnidx); free(M->tile_nnz); free(M->Format);
    free(M->blknnz); free(M->blknnznnz); free(M->dnsrowptr); free(M->dnscolptr);
    free(M->tilewidth); free(M->csr_offset); free(M->csrptr_offset); free(M->coo_offset);
    free(M->ell_offset); free(M->hyb_offset); free(M->hyb_coocount); free(M->dns_offset);
    free(M->dnsrow_offset); free(M->dnscol_offset); free(M->new_coocount);
    free(M->Blockcsr_Val); free(M->Blockcsr_Ptr); free(M->csr_compressedIdx);
    free(M->Blockcoo_Val); free(M->coo_compressed_Idx); free(M->Blockell_Val);
    free(M->ell_compressedIdx); free(M->Blockhyb_Val); free(M->hybIdx);
    free(M->Blockdense_Val); free(M->Blockdenserow_Val); free(M->denserowid);
    free(M->Blockdensecol_Val); free(M->densecolid); free(M->deferredcoo_val);
    free(M->deferredcoo_colidx); free(M->deferredcoo_ptr);
    memset(M, 0, sizeof(*M));
}

/*
 * Warp-chunk schedule, tilespmv_cpu.h:68-118: a block row with <=4 tiles is one chunk; longer
 * rows are cut into k=ceil(n/4) chunks of ceil(n/k) tiles, flagged by bit 31.
 * Arrays are malloc'ed here and returned through the double pointers (caller frees).
 */
void oracle_build_schedule(const oracle_tile_matrix *M, int *rowblkblock, unsigned int **rowidx,
                           int **colstart, int **colstop)
{
    int total = 0;
    for (int b = 0; b < M->tilem; b++)
    {
        int n = M->tile_ptr[b + 1] - M->tile_ptr[b];
        total += n <= CHUNK_TILES ? 1 : (n + CHUNK_TILES - 1) / CHUNK_TILES;
    }
    *rowblkblock = total;
    unsigned int *ri = (unsigned int *)zalloc(total, sizeof(unsigned int));
    int *cs = (int *)zalloc(total, sizeof(int));
    int *ce = (int *)zalloc(total, sizeof(int));
    int w = 0;
    for (int b = 0; b < M->tilem; b++)
    {
        int n = M->tile_ptr[b + 1] - M->tile_ptr[b];
        if (n <= CHUNK_TILES)
        {
            ri[w++] = (unsigned int)b;
            continue;
        }
        int k = (n + CHUNK_TILES - 1) / CHUNK_TILES;
        int len = (n + k - 1) / k;
        for (int c = 0; c < k; c++)
        {
            ri[w] = (unsigned int)b | 0x80000000u;
            cs[w] = M->tile_ptr[b] + c * len;
            ce[w] = c == k - 1 ? M->tile_ptr[b] + n : M->tile_ptr[b] + (c + 1) * len;
            w++;
        }
    }
    *rowidx = ri;
    *colstart = cs;
    *colstop = ce;
}

static inline int nib_at(const unsigned char *packed, int pos)
{
    unsigned char byte = packed[pos >> 1];
    return (pos & 1) ? (byte & 15) : (byte >> 4);
}

/*
 * CPU tile SpMV, tilespmv_cpu.h:125-272: same traversal, same per-tile / per-row accumulation
 * order (CSR/ELL/DenseRow/DenseCol: row sum then y += sum; COO and Dense add every product
 * straight into y; ELL skips stored zeros), so y is bit-identical to the reference for any data.
 * Also fills ptroffset1/2[tilenum] (running per-format offsets).
 * y must hold rowA entries; unlike the reference only rowlen (not 16) entries are zeroed in a
 * partial last block row (the reference writes past rowA there, tilespmv_cpu.h:128-131).
 */
void oracle_tilespmv_cpu(const oracle_tile_matrix *M, int *ptroffset1, int *ptroffset2, int rowA,
                         int colA, const val_t *x, val_t *y)
{
    int o_csr = 0, o_csrptr = 0, o_coo = 0, o_ell = 0, o_hyb = 0, o_hybidx = 0, o_dns = 0,
        o_dnsrow = 0, o_dnscol = 0;
    const int tilem = M->tilem, tilen = M->tilen;
    for (int b = 0; b < tilem; b++)
    {
        const int rowlen = rows_in_blockrow(b, tilem, rowA);
        val_t *yb = y + (size_t)b * TS;
        for (int r = 0; r < rowlen; r++)
            yb[r] = 0;
        for (int t = M->tile_ptr[b]; t < M->tile_ptr[b + 1]; t++)
        {
            const int collen = cols_in_tilecol(M->tile_columnidx[t], tilen, colA);
            const val_t *xt = x + (size_t)M->tile_columnidx[t] * TS;
            const int slots = M->blknnz[t + 1] - M->blknnz[t];
            switch (M->Format[t])
            {
            case 0:
            {
                ptroffset1[t] = o_csr;
                ptroffset2[t] = o_csrptr;
                const unsigned char *rp = M->Blockcsr_Ptr + o_csrptr;
                for (int r = 0; r < rowlen; r++)
                {
                    val_t sum = 0;
                    int end = r == rowlen - 1 ? slots : rp[r + 1];
                    for (int k = rp[r]; k < end; k++)
                        sum += xt[nib_at(M->csr_compressedIdx, o_csr + k)] * M->Blockcsr_Val[o_csr + k];
                    yb[r] += sum;
                }
                o_csr += slots;
                o_csrptr += rowlen;
                break;
            }
            case 1:
            {
                ptroffset1[t] = o_coo;
                for (int k = 0; k < slots; k++)
                {
                    unsigned char rc = M->coo_compressed_Idx[o_coo + k];
                    yb[rc >> 4] += M->Blockcoo_Val[o_coo + k] * xt[rc & 15];
                }
                o_coo += slots;
                break;
            }
            case 2:
            {
                ptroffset1[t] = o_ell;
                const int w = M->tilewidth[t];
                for (int r = 0; r < rowlen; r++)
                {
                    val_t sum = 0;
                    for (int s = 0; s < w; s++)
                    {
                        int pos = o_ell + s * rowlen + r;
                        if (M->Blockell_Val[pos] != 0)
                            sum += M->Blockell_Val[pos] * xt[nib_at(M->ell_compressedIdx, pos)];
                    }
                    yb[r] += sum;
                }
                o_ell += w * rowlen;
                break;
            }
            case 3:
            {
                ptroffset1[t] = o_hyb;
                ptroffset2[t] = o_hybidx;
                const int w = M->tilewidth[t];
                const unsigned char *hidx = M->hybIdx + o_hybidx;
                for (int r = 0; r < rowlen; r++)
                {
                    val_t sum = 0;
                    for (int s = 0; s < w; s++)
                    {
                        int lp = s * rowlen + r; /* HYB nibble parity is tile-local */
                        if (M->Blockhyb_Val[o_hyb + lp] != 0)
                            sum += M->Blockhyb_Val[o_hyb + lp] * xt[nib_at(hidx, lp)];
                    }
                    yb[r] += sum;
                }
                int ellslots = w * rowlen, spill = slots - ellslots;
                const unsigned char *sidx = hidx + (ellslots + 1) / 2;
                for (int k = 0; k < spill; k++)
                    yb[sidx[k] >> 4] += M->Blockhyb_Val[o_hyb + ellslots + k] * xt[sidx[k] & 15];
                o_hyb += slots;
                o_hybidx += (ellslots + 1) / 2 + spill;
                break;
            }
            case 4:
            {
                ptroffset1[t] = o_dns;
                for (int r = 0; r < rowlen; r++)
                    for (int c = 0; c < collen; c++)
                        yb[r] += xt[c] * M->Blockdense_Val[o_dns + c * rowlen + r];
                o_dns += rowlen * collen;
                break;
            }
            case 5:
            {
                ptroffset1[t] = o_dnsrow;
                for (int d = M->dnsrowptr[t]; d < M->dnsrowptr[t + 1]; d++)
                {
                    val_t sum = 0;
                    const val_t *rowv = M->Blockdenserow_Val + o_dnsrow + (d - M->dnsrowptr[t]) * collen;
                    for (int c = 0; c < collen; c++)
                        sum += xt[c] * rowv[c];
                    yb[(int)M->denserowid[d]] += sum;
                }
                o_dnsrow += slots;
                break;
            }
            case 6:
            {
                ptroffset1[t] = o_dnscol;
                for (int r = 0; r < rowlen; r++)
                {
                    val_t sum = 0;
                    for (int d = M->dnscolptr[t]; d < M->dnscolptr[t + 1]; d++)
                        sum += M->Blockdensecol_Val[o_dnscol + (d - M->dnscolptr[t]) * rowlen + r] *
                               xt[(int)M->densecolid[d]];
                    yb[r] += sum;
                }
                o_dnscol += slots;
                break;
            }
            }
        }
    }
}

/* y_golden of the reference driver, main.cu:101-110 (serial, row sum then store) */
void oracle_csr_spmv(int rowA, const int *rowptr, const int *colidx, const val_t *val,
                     const val_t *x, val_t *y)
{
    for (int i = 0; i < rowA; i++)
    {
        val_t sum = 0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
            sum += val[j] * x[colidx[j]];
        y[i] = sum;
    }
}

/* OpenMP variant of the same loop, used only to build y references for the huge configs */
void oracle_csr_spmv_omp(int rowA, const int *rowptr, const int *colidx, const val_t *val,
                         const val_t *x, val_t *y)
{
#pragma omp parallel for schedule(static, 4096)
    for (int i = 0; i < rowA; i++)
    {
        val_t sum = 0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
            sum += val[j] * x[colidx[j]];
        y[i] = sum;
    }
}

/* sum_j |a_ij| |x_j| per row: the scale of the floating-point parity tolerance */
void oracle_csr_abs_spmv(int rowA, const int *rowptr, const int *colidx, const val_t *val,
                         const val_t *x, val_t *y)
{
#pragma omp parallel for schedule(static, 4096)
    for (int i = 0; i < rowA; i++)
    {
        val_t sum = 0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
            sum += (val_t)fabs((double)val[j]) * (val_t)fabs((double)x[colidx[j]]);
        y[i] = sum;
    }
}

double oracle_time_tilespmv_cpu(const oracle_tile_matrix *M, int rowA, int colA, const val_t *x,
                                val_t *y)
{
    int *p1 = (int *)zalloc(M->tilenum, sizeof(int)), *p2 = (int *)zalloc(M->tilenum, sizeof(int));
    struct timeval t1, t2;
    gettimeofday(&t1, NULL);
    oracle_tilespmv_cpu(M, p1, p2, rowA, colA, x, y);
    gettimeofday(&t2, NULL);
    free(p1);
    free(p2);
    return (t2.tv_sec - t1.tv_sec) * 1000.0 + (t2.tv_usec - t1.tv_usec) / 1000.0;
}

/*
 * Matrix Market coordinate reader with the semantics of mmio_allinone (mmio_highlevel.h:593-759):
 * 1-based -> 0-based, real / integer / pattern (=1.0) / complex (real part), symmetric and
 * hermitian files expanded by mirroring every off-diagonal entry right after the entry itself,
 * entries bucketed by row in FILE ORDER (columns are not sorted).
 * Returns 0, or -1 (open) / -2 (banner) / -4 (size line) like the reference.
 */
int oracle_mtx_read(const char *filename, int *m, int *n, int *nnz, int *is_symmetric, int **rowptr_out,
                    int **colidx_out, val_t **val_out)
{
    FILE *f = fopen(filename, "r");
    if (!f)
        return -1;
    char line[1100];
    if (!fgets(line, sizeof(line), f))
    {
        fclose(f);
        return -2;
    }
    char banner[64], object[64], fmt[64], field[64], sym[64];
    if (sscanf(line, "%63s %63s %63s %63s %63s", banner, object, fmt, field, sym) != 5 ||
        strcmp(banner, "%%MatrixMarket") != 0)
    {
        fclose(f);
        return -2;
    }
    for (char *p = field; *p; p++)
        if (*p >= 'A' && *p <= 'Z')
            *p += 32;
    for (char *p = sym; *p; p++)
        if (*p >= 'A' && *p <= 'Z')
            *p += 32;
    int is_pattern = !strcmp(field, "pattern"), is_complex = !strcmp(field, "complex");
    int symm = !strcmp(sym, "symmetric") || !strcmp(sym, "hermitian");
    int M_, N_, NZ;
    do
    {
        if (!fgets(line, sizeof(line), f))
        {
            fclose(f);
            return -4;
        }
    } while (line[0] == '%');
    while (sscanf(line, "%d %d %d", &M_, &N_, &NZ) != 3)
    {
        if (!fgets(line, sizeof(line), f))
        {
            fclose(f);
            return -4;
        }
    }
    int *ri = (int *)zalloc(NZ, sizeof(int)), *ci = (int *)zalloc(NZ, sizeof(int));
    val_t *vv = (val_t *)zalloc(NZ, sizeof(val_t));
    int *cnt = (int *)zalloc(M_ + 1, sizeof(int));
    for (int e = 0; e < NZ; e++)
    {
        int a = 0, b = 0;
        double re = 1.0, im = 0.0;
        if (is_pattern)
        {
            if (fscanf(f, "%d %d", &a, &b) != 2)
            {
                NZ = e;
                break;
            }
        }
        else if (is_complex)
        {
            if (fscanf(f, "%d %d %lg %lg", &a, &b, &re, &im) != 4)
            {
                NZ = e;
                break;
            }
        }
        else
        {
            if (fscanf(f, "%d %d %lg", &a, &b, &re) != 3)
            {
                NZ = e;
                break;
            }
        }
        ri[e] = a - 1;
        ci[e] = b - 1;
        vv[e] = (val_t)re;
        cnt[ri[e]]++;
        if (symm && ri[e] != ci[e])
            cnt[ci[e]]++;
    }
    fclose(f);
    prefix_excl(cnt, M_ + 1);
    int total = cnt[M_];
    int *rp = (int *)zalloc(M_ + 1, sizeof(int));
    memcpy(rp, cnt, sizeof(int) * (M_ + 1));
    int *cj = (int *)zalloc(total, sizeof(int));
    val_t *cv = (val_t *)zalloc(total, sizeof(val_t));
    for (int e = 0; e < NZ; e++)
    {
        int d = cnt[ri[e]]++;
        cj[d] = ci[e];
        cv[d] = vv[e];
        if (symm && ri[e] != ci[e])
        {
            d = cnt[ci[e]]++;
            cj[d] = ri[e];
            cv[d] = vv[e];
        }
    }
    free(ri); free(ci); free(vv); free(cnt);
    *m = M_; *n = N_; *nnz = total; *is_symmetric = symm;
    *rowptr_out = rp; *colidx_out = cj; *val_out = cv;
    return 0;
}

void oracle_free(void *p) { free(p); }

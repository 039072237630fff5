/*
 * gen.c -- deterministic synthetic CSR generators for the BASELINE.json configs
 * (SURVEY.md 8(d)).  Host-only C + OpenMP, built into libtilespmv_gen.so; used by the tests,
 * bench.py and the CLI.  The reference ships no inputs at all, so these are new.
 *
 * Every generator writes a CSR with ascending column indices inside each row (what a
 * column-major-sorted .mtx gives after the reference's row bucketing, mmio_highlevel.h:734-740).
 * Two-call protocol: call with rowptr/colidx/val == NULL to get nnz (return value), then with
 * buffers of m+1 / nnz / nnz entries.  val_mode 0: uniform(-1,1) (stencils: their natural
 * coefficients); val_mode 1: val[j] = j % 10 like the reference driver (main.cu:68-69).
 * Values are always written as double; the caller narrows for fp32.
 */
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
/* counter-based stream: (seed, a, b) -> 64 random bits */
static inline uint64_t rnd3(uint64_t seed, uint64_t a, uint64_t b)
{
    return splitmix64(splitmix64(splitmix64(seed) ^ a) ^ (b * 0xD6E8FEB86659FD93ull));
}
static inline double u11(uint64_t r) /* uniform in (-1,1), never exactly 0 */
{
    double u = (double)((r >> 11) + 1) / 9007199254740994.0; /* (0,1) */
    return 2.0 * u - 1.0;
}

static void prefix_from_counts(int64_t m, int *rowptr)
{
    int run = 0;
    for (int64_t i = 0; i <= m; i++)
    {
        int c = rowptr[i];
        rowptr[i] = run;
        run += c;
    }
}

static void fill_values(int64_t m, const int *rowptr, double *val, int val_mode, uint64_t seed)
{
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < m; i++)
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
            val[j] = val_mode == 1 ? (double)(j % 10) : u11(rnd3(seed ^ 0x5151ull, (uint64_t)i, (uint64_t)j));
}

/* ---- 2-D 5-point Laplacian on a G x G grid, Dirichlet (config 1) ---- */
int64_t tsgen_lap2d(int G, int val_mode, int *rowptr, int *colidx, double *val)
{
    int64_t m = (int64_t)G * G;
    int64_t nnz = 5 * m - 4 * (int64_t)G;
    if (!rowptr)
        return nnz;
#pragma omp parallel for schedule(static, 4096)
    for (int64_t r = 0; r < m; r++)
    {
        int i = (int)(r / G), j = (int)(r % G);
        rowptr[r] = 1 + (i > 0) + (i < G - 1) + (j > 0) + (j < G - 1);
    }
    rowptr[m] = 0;
    prefix_from_counts(m, rowptr);
#pragma omp parallel for schedule(static, 4096)
    for (int64_t r = 0; r < m; r++)
    {
        int i = (int)(r / G), j = (int)(r % G);
        int p = rowptr[r];
        if (i > 0) { colidx[p] = (int)(r - G); val[p++] = -1.0; }
        if (j > 0) { colidx[p] = (int)(r - 1); val[p++] = -1.0; }
        colidx[p] = (int)r; val[p++] = 4.0;
        if (j < G - 1) { colidx[p] = (int)(r + 1); val[p++] = -1.0; }
        if (i < G - 1) { colidx[p] = (int)(r + G); val[p++] = -1.0; }
    }
    if (val_mode == 1)
        fill_values(m, rowptr, val, 1, 0);
    return nnz;
}

/* ---- 3-D 27-point Laplacian on a G^3 grid, Dirichlet (config 2) ---- */
int64_t tsgen_lap3d27(int G, int val_mode, int *rowptr, int *colidx, double *val)
{
    int64_t m = (int64_t)G * G * G;
    int64_t e = 3 * (int64_t)G - 2;
    int64_t nnz = e * e * e;
    if (!rowptr)
        return nnz;
#pragma omp parallel for schedule(static, 4096)
    for (int64_t r = 0; r < m; r++)
    {
        int k = (int)(r % G), j = (int)((r / G) % G), i = (int)(r / ((int64_t)G * G));
        int ci = 1 + (i > 0) + (i < G - 1), cj = 1 + (j > 0) + (j < G - 1), ck = 1 + (k > 0) + (k < G - 1);
        rowptr[r] = ci * cj * ck;
    }
    rowptr[m] = 0;
    prefix_from_counts(m, rowptr);
#pragma omp parallel for schedule(static, 4096)
    for (int64_t r = 0; r < m; r++)
    {
        int k = (int)(r % G), j = (int)((r / G) % G), i = (int)(r / ((int64_t)G * G));
        int p = rowptr[r];
        for (int di = -1; di <= 1; di++)
        {
            if (i + di < 0 || i + di >= G) continue;
            for (int dj = -1; dj <= 1; dj++)
            {
                if (j + dj < 0 || j + dj >= G) continue;
                for (int dk = -1; dk <= 1; dk++)
                {
                    if (k + dk < 0 || k + dk >= G) continue;
                    colidx[p] = (int)(((int64_t)(i + di) * G + (j + dj)) * G + (k + dk));
                    val[p++] = (di == 0 && dj == 0 && dk == 0) ? 26.0 : -1.0;
                }
            }
        }
    }
    if (val_mode == 1)
        fill_values(m, rowptr, val, 1, 0);
    return nnz;
}

/* ---- slab [i0,i1) of the 27-point Laplacian on an nx x ny x nz grid (row = (i*ny+j)*nz+k):
        local rows, GLOBAL columns -- the row block one GPU owns in the multi-GPU runs ---- */
int64_t tsgen_lap3d27_slab(int nx, int ny, int nz, int i0, int i1, int val_mode, int *rowptr, int *colidx,
                           double *val)
{
    const int64_t plane = (int64_t)ny * nz;
    const int64_t m = (int64_t)(i1 - i0) * plane;
    int64_t nnz = 0;
    for (int i = i0; i < i1; i++)
    {
        int ci = 1 + (i > 0) + (i < nx - 1);
        nnz += (int64_t)ci * (3 * (int64_t)ny - 2) * (3 * (int64_t)nz - 2);
    }
    if (!rowptr)
        return nnz;
#pragma omp parallel for schedule(static, 4096)
    for (int64_t r = 0; r < m; r++)
    {
        int k = (int)(r % nz), j = (int)((r / nz) % ny), i = i0 + (int)(r / plane);
        rowptr[r] = (1 + (i > 0) + (i < nx - 1)) * (1 + (j > 0) + (j < ny - 1)) * (1 + (k > 0) + (k < nz - 1));
    }
    rowptr[m] = 0;
    prefix_from_counts(m, rowptr);
#pragma omp parallel for schedule(static, 4096)
    for (int64_t r = 0; r < m; r++)
    {
        int k = (int)(r % nz), j = (int)((r / nz) % ny), i = i0 + (int)(r / plane);
        int p = rowptr[r];
        for (int di = -1; di <= 1; di++)
        {
            if (i + di < 0 || i + di >= nx) continue;
            for (int dj = -1; dj <= 1; dj++)
            {
                if (j + dj < 0 || j + dj >= ny) continue;
                for (int dk = -1; dk <= 1; dk++)
                {
                    if (k + dk < 0 || k + dk >= nz) continue;
                    colidx[p] = (int)(((int64_t)(i + di) * ny + (j + dj)) * nz + (k + dk));
                    val[p++] = (di == 0 && dj == 0 && dk == 0) ? 26.0 : -1.0;
                }
            }
        }
    }
    if (val_mode == 1)
        fill_values(m, rowptr, val, 1, 0);
    return nnz;
}

/* ---- banded FEM-like: diagonal + `per_row` distinct offsets in [-hb,hb]\{0}, clipped (config 3) ---- */
static int banded_row(int64_t N, int64_t i, int hb, int per_row, uint64_t seed, int *cols)
{
    /* partial Fisher-Yates over the 2*hb candidate offsets, keyed by (seed, i) */
    int cand[512];
    int nc = 0;
    for (int d = -hb; d <= hb; d++)
        if (d != 0)
            cand[nc++] = d;
    int take = per_row < nc ? per_row : nc;
    for (int s = 0; s < take; s++)
    {
        int pick = s + (int)(rnd3(seed, (uint64_t)i, (uint64_t)s) % (uint64_t)(nc - s));
        int tmp = cand[s];
        cand[s] = cand[pick];
        cand[pick] = tmp;
    }
    int n = 0;
    cols[n++] = 0;
    for (int s = 0; s < take; s++)
        if (i + cand[s] >= 0 && i + cand[s] < N)
            cols[n++] = cand[s];
    /* insertion sort of the kept offsets */
    for (int a = 1; a < n; a++)
    {
        int v = cols[a], b = a - 1;
        while (b >= 0 && cols[b] > v)
        {
            cols[b + 1] = cols[b];
            b--;
        }
        cols[b + 1] = v;
    }
    return n;
}

int64_t tsgen_banded(int64_t N, int hb, int per_row, uint64_t seed, int val_mode, int *rowptr,
                     int *colidx, double *val)
{
    if (hb > 255) hb = 255;
    if (!rowptr)
    {
        int64_t nnz = 0;
#pragma omp parallel for schedule(static, 4096) reduction(+ : nnz)
        for (int64_t i = 0; i < N; i++)
        {
            int cols[520];
            nnz += banded_row(N, i, hb, per_row, seed, cols);
        }
        return nnz;
    }
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < N; i++)
    {
        int cols[520];
        rowptr[i] = banded_row(N, i, hb, per_row, seed, cols);
    }
    rowptr[N] = 0;
    prefix_from_counts(N, rowptr);
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < N; i++)
    {
        int cols[520];
        int n = banded_row(N, i, hb, per_row, seed, cols);
        for (int k = 0; k < n; k++)
            colidx[rowptr[i] + k] = (int)(i + cols[k]);
    }
    fill_values(N, rowptr, val, val_mode, seed);
    return rowptr[N];
}

/* rows [row0, row0 + nrows) of the SAME banded matrix (local rowptr, global columns, identical values): a shard
 * generates only its own row block.  Rows hb <= i < N - hb are never clipped and hold take + 1 entries, so the global
 * position of the first entry needs only the two clipped edge zones. */
int64_t tsgen_banded_rows(int64_t N, int64_t row0, int64_t nrows, int hb, int per_row, uint64_t seed, int val_mode,
                          int *rowptr, int *colidx, double *val)
{
    if (hb > 255) hb = 255;
    if (row0 < 0 || nrows < 0 || row0 + nrows > N)
        return -1;
    const int take = per_row < 2 * hb ? per_row : 2 * hb;
    int64_t base = 0; /* entries of rows [0, row0) */
    for (int64_t i = 0; i < row0;)
    {
        if (i >= hb && i < N - hb)
        {
            const int64_t stop = row0 < N - hb ? row0 : N - hb;
            base += (stop - i) * (int64_t)(take + 1);
            i = stop;
            continue;
        }
        int cols[520];
        base += banded_row(N, i, hb, per_row, seed, cols);
        i++;
    }
    if (!rowptr)
    {
        int64_t nnz = 0;
#pragma omp parallel for schedule(static, 4096) reduction(+ : nnz)
        for (int64_t r = 0; r < nrows; r++)
        {
            int cols[520];
            nnz += banded_row(N, row0 + r, hb, per_row, seed, cols);
        }
        return nnz;
    }
#pragma omp parallel for schedule(static, 4096)
    for (int64_t r = 0; r < nrows; r++)
    {
        int cols[520];
        rowptr[r] = banded_row(N, row0 + r, hb, per_row, seed, cols);
    }
    rowptr[nrows] = 0;
    prefix_from_counts(nrows, rowptr);
#pragma omp parallel for schedule(static, 4096)
    for (int64_t r = 0; r < nrows; r++)
    {
        int cols[520];
        const int64_t i = row0 + r;
        int n = banded_row(N, i, hb, per_row, seed, cols);
        for (int k = 0; k < n; k++)
        {
            const int64_t p = rowptr[r] + k, pg = base + p; /* local / global position */
            colidx[p] = (int)(i + cols[k]);
            val[p] = val_mode == 1 ? (double)(pg % 10) : u11(rnd3(seed ^ 0x5151ull, (uint64_t)i, (uint64_t)pg));
        }
    }
    return rowptr[nrows];
}

/* ---- contiguous band |i-j| <= hb (config 3b: Dense + CSR + COO tile mix) ---- */
int64_t tsgen_band_contig(int64_t N, int hb, uint64_t seed, int val_mode, int *rowptr, int *colidx,
                          double *val)
{
    int64_t nnz = 0;
    for (int64_t i = 0; i < N; i++)
    {
        int64_t lo = i - hb < 0 ? 0 : i - hb, hi = i + hb >= N ? N - 1 : i + hb;
        if (rowptr) rowptr[i] = (int)(hi - lo + 1);
        nnz += hi - lo + 1;
    }
    if (!rowptr)
        return nnz;
    rowptr[N] = 0;
    prefix_from_counts(N, rowptr);
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < N; i++)
    {
        int64_t lo = i - hb < 0 ? 0 : i - hb;
        for (int k = 0; k < rowptr[i + 1] - rowptr[i]; k++)
            colidx[rowptr[i] + k] = (int)(lo + k);
    }
    fill_values(N, rowptr, val, val_mode, seed);
    return nnz;
}

/* ---- uniform random: exactly per_row distinct columns per row (config 5) ---- */
static void uniform_row(int64_t ncols, int64_t i, int per_row, uint64_t seed, int *cols)
{
    int n = 0;
    uint64_t ctr = 0;
    while (n < per_row)
    {
        int c = (int)(rnd3(seed, (uint64_t)i, ctr++) % (uint64_t)ncols);
        int dup = 0;
        for (int k = 0; k < n; k++)
            if (cols[k] == c) { dup = 1; break; }
        if (!dup)
            cols[n++] = c;
    }
    for (int a = 1; a < n; a++)
    {
        int v = cols[a], b = a - 1;
        while (b >= 0 && cols[b] > v)
        {
            cols[b + 1] = cols[b];
            b--;
        }
        cols[b + 1] = v;
    }
}

/* rows [row0, row0+nrows) of the N x ncols matrix: shards generate only their own rows */
int64_t tsgen_uniform_rows(int64_t ncols, int64_t row0, int64_t nrows, int per_row, uint64_t seed,
                           int val_mode, int *rowptr, int *colidx, double *val)
{
    if (per_row > 64) per_row = 64;
    if (per_row > ncols) per_row = (int)ncols;
    int64_t nnz = nrows * per_row;
    if (!rowptr)
        return nnz;
#pragma omp parallel for schedule(static, 4096)
    for (int64_t r = 0; r < nrows; r++)
    {
        int cols[64];
        rowptr[r] = (int)(r * per_row);
        uniform_row(ncols, row0 + r, per_row, seed, cols);
        for (int k = 0; k < per_row; k++)
        {
            int64_t p = r * per_row + k;
            colidx[p] = cols[k];
            val[p] = val_mode == 1 ? (double)(p % 10) : u11(rnd3(seed ^ 0x5151ull, (uint64_t)(row0 + r), (uint64_t)k));
        }
    }
    rowptr[nrows] = (int)nnz;
    return nnz;
}

/* ---- R-MAT power-law graph, duplicates merged (config 4) ---- */
static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return x < y ? -1 : (x > y);
}

int64_t tsgen_rmat(int scale, int edge_factor, double pa, double pb, double pc, uint64_t seed,
                   int val_mode, int *rowptr, int *colidx, double *val, int64_t cap)
{
    /* Edges are regenerated on every call (deterministic); the caller sizes colidx/val by a
       first call with rowptr == NULL.  cap = capacity of colidx/val (sanity check). */
    const int64_t n = (int64_t)1 << scale;
    const int64_t ne = n * edge_factor;
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)ne);
    if (!keys)
        return -1;
    const uint64_t ta = (uint64_t)(pa * 4294967296.0), tb = (uint64_t)((pa + pb) * 4294967296.0),
                   tc = (uint64_t)((pa + pb + pc) * 4294967296.0);
#pragma omp parallel for schedule(static, 65536)
    for (int64_t e = 0; e < ne; e++)
    {
        int64_t r = 0, c = 0;
        for (int lvl = 0; lvl < scale; lvl += 2)
        {
            uint64_t bits = rnd3(seed, (uint64_t)e, (uint64_t)lvl);
            for (int h = 0; h < 2 && lvl + h < scale; h++)
            {
                uint64_t u = (bits >> (32 * h)) & 0xFFFFFFFFull;
                int q = u < ta ? 0 : (u < tb ? 1 : (u < tc ? 2 : 3));
                r = (r << 1) | (q >> 1);
                c = (c << 1) | (q & 1);
            }
        }
        keys[e] = (r << 32) | c;
    }
    /* bucket by row range in parallel, then sort: simple parallel sample-free sort */
    const int nb = 1024;
    int64_t *bcnt = (int64_t *)calloc(nb + 1, sizeof(int64_t));
    const int shift = scale > 10 ? scale - 10 : 0;
    for (int64_t e = 0; e < ne; e++)
        bcnt[(keys[e] >> 32) >> shift]++;
    int64_t run = 0;
    for (int b = 0; b <= nb; b++)
    {
        int64_t c = bcnt[b];
        bcnt[b] = run;
        run += c;
    }
    int64_t *sorted = (int64_t *)malloc(sizeof(int64_t) * (size_t)ne);
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (nb + 1));
    memcpy(fill, bcnt, sizeof(int64_t) * (nb + 1));
    for (int64_t e = 0; e < ne; e++)
        sorted[fill[(keys[e] >> 32) >> shift]++] = keys[e];
    free(keys);
    free(fill);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nb; b++)
        qsort(sorted + bcnt[b], (size_t)(bcnt[b + 1] - bcnt[b]), sizeof(int64_t), cmp_i64);
    free(bcnt);
    /* unique */
    int64_t nnz = 0;
    for (int64_t e = 0; e < ne; e++)
        if (e == 0 || sorted[e] != sorted[e - 1])
            nnz++;
    if (!rowptr)
    {
        free(sorted);
        return nnz;
    }
    if (nnz > cap)
    {
        free(sorted);
        return -2;
    }
    memset(rowptr, 0, sizeof(int) * (size_t)(n + 1));
    int64_t p = 0;
    for (int64_t e = 0; e < ne; e++)
        if (e == 0 || sorted[e] != sorted[e - 1])
        {
            rowptr[sorted[e] >> 32]++;
            colidx[p++] = (int)(sorted[e] & 0xFFFFFFFFll);
        }
    free(sorted);
    prefix_from_counts(n, rowptr);
    fill_values(n, rowptr, val, val_mode, seed);
    return nnz;
}

/*
 * ---- the 32 x 40 seven-format fixture of SURVEY.md Appendix C.3 (385 nnz) ----
 * tile (0,0): local rows 3 and 7 full                      -> DenseRow
 * tile (0,1): local columns 1 and 5 full                    -> DenseCol
 * tile (0,2): entries (0,33) (5,39) (15,32)                 -> COO
 * tile (1,0): full 16 x 16                                  -> Dense
 * tile (1,1): row lengths 10,1,8,0,3,0,0,5,0,0,2,0,0,0,0,1  -> CSR
 * tile (1,2): row r has local columns r%8 and (r+3)%8 (collen = 8) -> ELL
 */
int64_t tsgen_seven_formats(int *rowptr, int *colidx, double *val)
{
    static const int csr_len[16] = {10, 1, 8, 0, 3, 0, 0, 5, 0, 0, 2, 0, 0, 0, 0, 1};
    int cols[64];
    int64_t p = 0;
    for (int r = 0; r < 32; r++)
    {
        int n = 0;
        if (r < 16)
        {
            if (r == 3 || r == 7)
                for (int c = 0; c < 16; c++) cols[n++] = c;
            cols[n++] = 16 + 1;
            cols[n++] = 16 + 5;
            if (r == 0) cols[n++] = 33;
            if (r == 5) cols[n++] = 39;
            if (r == 15) cols[n++] = 32;
        }
        else
        {
            int lr = r - 16;
            for (int c = 0; c < 16; c++) cols[n++] = c;
            for (int c = 0; c < csr_len[lr]; c++) cols[n++] = 16 + c;
            int a = lr % 8, b = (lr + 3) % 8;
            cols[n++] = 32 + (a < b ? a : b);
            cols[n++] = 32 + (a < b ? b : a);
        }
        if (rowptr)
        {
            rowptr[r] = (int)p;
            for (int k = 0; k < n; k++)
            {
                colidx[p + k] = cols[k];
                val[p + k] = (double)((p + k) % 10);
            }
        }
        p += n;
    }
    if (rowptr)
        rowptr[32] = (int)p;
    return p;
}

/* Fast Matrix Market writer ("coordinate real general", row-major, 1-based): inputs for the
 * reference's own driver (./test -d 0 file.mtx).  Values are written with %.17g. Returns 0 / -1. */
#include <stdio.h>
int tsgen_write_mtx(const char *path, int m, int n, const int *rowptr, const int *colidx, const double *val)
{
    FILE *f = fopen(path, "w");
    if (!f)
        return -1;
    static char buf[1 << 22];
    setvbuf(f, buf, _IOFBF, sizeof(buf));
    fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n", m, n, rowptr[m]);
    for (int i = 0; i < m; i++)
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
            fprintf(f, "%d %d %.17g\n", i + 1, colidx[j] + 1, val[j]);
    return fclose(f) == 0 ? 0 : -1;
}

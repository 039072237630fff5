/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin shim that #includes the UNMODIFIED reference CPU path where it lies under
 * /root/reference/src (nothing is copied into this repo) and exposes it through a
 * plain C ABI so tests/ and bench.py's cpu_baseline / --impl reference leg can call
 * it through ctypes.  Built by oracle/Makefile into oracle/_ref/ (git-ignored).
 *
 *   Tile_create        /root/reference/src/csr2tile.h:629-1020
 *   tilespmv_cpu       /root/reference/src/tilespmv_cpu.h:3-285
 *   mmio_allinone      /root/reference/src/mmio_highlevel.h:593-759
 *   Tile_matrix        /root/reference/src/format.h:3-56
 *
 * Precision is the reference's compile-time macro MAT_VAL_TYPE (double default,
 * -DMAT_VAL_TYPE=float for fp32): the Makefile builds the shim twice.
 */
#include "common.h"
#include "mmio_highlevel.h"
#include "utils.h"
#include "csr2tile.h"
#include "tilespmv_cpu.h"

int ref_sizeof_val(void) { return (int)sizeof(MAT_VAL_TYPE); }
int ref_sizeof_tile_matrix(void) { return (int)sizeof(Tile_matrix); }
int ref_omp_max_threads(void) { return omp_get_max_threads(); }

/* Tile_create on a caller-allocated struct (caller keeps CSR ownership). */
void ref_Tile_create(Tile_matrix *matrix, int rowA, int colA, MAT_PTR_TYPE nnzA,
                     MAT_PTR_TYPE *csrRowPtrA, int *csrColIdxA, MAT_VAL_TYPE *csrValA)
{
    Tile_create(matrix, rowA, colA, nnzA, csrRowPtrA, csrColIdxA, csrValA);
}

/* tilespmv_cpu with the reference's exact argument list. */
void ref_tilespmv_cpu(Tile_matrix *matrix, int *ptroffset1, int *ptroffset2, int *rowblkblock,
                      unsigned int **blkcoostylerowidx, int **blkcoostylerowidx_colstart,
                      int **blkcoostylerowidx_colstop, int rowA, int colA, MAT_PTR_TYPE nnzA,
                      MAT_PTR_TYPE *csrRowPtrA, int *csrColIdxA, MAT_VAL_TYPE *csrValA,
                      MAT_VAL_TYPE *x, MAT_VAL_TYPE *y, MAT_VAL_TYPE *y_golden)
{
    tilespmv_cpu(matrix, ptroffset1, ptroffset2, rowblkblock, blkcoostylerowidx,
                 blkcoostylerowidx_colstart, blkcoostylerowidx_colstop, rowA, colA, nnzA,
                 csrRowPtrA, csrColIdxA, csrValA, x, y, y_golden);
}

int ref_mmio_allinone(int *m, int *n, MAT_PTR_TYPE *nnz, int *isSymmetric,
                      MAT_PTR_TYPE **csrRowPtr, int **csrColIdx, MAT_VAL_TYPE **csrVal,
                      char *filename)
{
    return mmio_allinone(m, n, nnz, isSymmetric, csrRowPtr, csrColIdx, csrVal, filename);
}

void ref_free(void *p) { free(p); }

/* wall-clock helper so the baseline leg times exactly the reference call */
double ref_time_tilespmv_cpu(Tile_matrix *matrix, int rowA, int colA, MAT_PTR_TYPE nnzA,
                             MAT_PTR_TYPE *csrRowPtrA, int *csrColIdxA, MAT_VAL_TYPE *csrValA,
                             MAT_VAL_TYPE *x, MAT_VAL_TYPE *y, MAT_VAL_TYPE *y_golden)
{
    int tilenum = matrix->tilenum;
    int *p1 = (int *)calloc(tilenum > 0 ? tilenum : 1, sizeof(int));
    int *p2 = (int *)calloc(tilenum > 0 ? tilenum : 1, sizeof(int));
    int rbb = 0;
    unsigned int *a = NULL;
    int *b = NULL, *c = NULL;
    struct timeval t1, t2;
    gettimeofday(&t1, NULL);
    tilespmv_cpu(matrix, p1, p2, &rbb, &a, &b, &c, rowA, colA, nnzA, csrRowPtrA, csrColIdxA,
                 csrValA, x, y, y_golden);
    gettimeofday(&t2, NULL);
    free(p1); free(p2); free(a); free(b); free(c);
    return (t2.tv_sec - t1.tv_sec) * 1000.0 + (t2.tv_usec - t1.tv_usec) / 1000.0;
}

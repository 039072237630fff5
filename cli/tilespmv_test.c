/*
 * cli/tilespmv_test.c -- the reference's command line on top of the drop-in C-ABI:
 *
 *     ./tilespmv_test -d <device_id> <matrix.mtx>
 *
 * Plain C host code (no CUDA in this file): it follows the flow of the reference driver
 * (/root/reference/src/main.cu:15-205 -- argument parsing :35-59, mmio_allinone :63, values
 * overwritten by i % 10 :68-69, rowA truncated to a multiple of 16 :71, Tile_create :87-91,
 * x[i] = i % 10 :93-97, serial CSR y_golden :101-110, the tilespmv_cpu call :142-156 replaced by
 * tilespmv_prepare (bookkeeping only -- this library has no CPU SpMV), call_tilespmv_cuda :165-180,
 * the 1 % check :186-197) using the reference's own unsuffixed names, which
 * TILESPMV_REFERENCE_NAMES maps onto libtilespmv_b200.so.  Build: see cli/Makefile.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#define TILESPMV_REFERENCE_NAMES
#include "tilespmv.h"

/* the only CUDA runtime call the reference driver makes itself (main.cu:74) */
extern int cudaSetDevice(int device);

int main(int argc, char **argv)
{
    if (argc < 4 || strcmp(argv[1], "-d") != 0)
    {
        fprintf(stderr, "usage: %s -d <device_id> <matrix.mtx>\n", argv[0]);
        return 0; /* the reference exits silently with 0 as well (main.cu:35-59) */
    }
    const int device_id = atoi(argv[2]);
    char *filename = argv[3];
    printf("-------------- %s --------------\n", filename);

    int rowA, colA, isSymmetricA;
    MAT_PTR_TYPE nnzA;
    MAT_PTR_TYPE *csrRowPtrA;
    int *csrColIdxA;
    MAT_VAL_TYPE *csrValA;
    struct timeval t1, t2;
    gettimeofday(&t1, NULL);
    int rc = mmio_allinone(&rowA, &colA, &nnzA, &isSymmetricA, &csrRowPtrA, &csrColIdxA, &csrValA, filename);
    gettimeofday(&t2, NULL);
    if (rc != 0)
    {
        printf("  cannot read %s (mmio_allinone returned %d)\n", filename, rc);
        return 0;
    }
    printf("  input matrix A: ( %i, %i ) nnz = %i\n  loadfile time    = %4.5f sec\n", rowA, colA, nnzA,
           (t2.tv_sec - t1.tv_sec) + (t2.tv_usec - t1.tv_usec) / 1e6);
    for (int i = 0; i < nnzA; i++)
        csrValA[i] = i % 10;
    rowA = (rowA / BLOCK_SIZE) * BLOCK_SIZE;
    nnzA = csrRowPtrA[rowA]; /* nonzeros of the rows that are kept */
    cudaSetDevice(device_id);

    Tile_matrix *matrixA = (Tile_matrix *)calloc(1, sizeof(Tile_matrix));
    Tile_create(matrixA, rowA, colA, nnzA, csrRowPtrA, csrColIdxA, csrValA);
    if (matrixA->tilenum < 0)
    {
        printf("  Tile_create failed: %s\n", tilespmv_last_error());
        return 0;
    }

    MAT_VAL_TYPE *x = (MAT_VAL_TYPE *)malloc(sizeof(MAT_VAL_TYPE) * (colA > 0 ? colA : 1));
    for (int i = 0; i < colA; i++)
        x[i] = i % 10;
    MAT_VAL_TYPE *y_golden = (MAT_VAL_TYPE *)malloc(sizeof(MAT_VAL_TYPE) * (rowA > 0 ? rowA : 1));
    for (int i = 0; i < rowA; i++)
    {
        MAT_VAL_TYPE sum = 0;
        for (int j = csrRowPtrA[i]; j < csrRowPtrA[i + 1]; j++)
            sum += csrValA[j] * x[csrColIdxA[j]];
        y_golden[i] = sum;
    }

    /* schedule + ptroffset arrays the reference obtains from tilespmv_cpu */
    int T = matrixA->tilenum > 0 ? matrixA->tilenum : 1;
    int *ptroffset1 = (int *)calloc(T, sizeof(int)), *ptroffset2 = (int *)calloc(T, sizeof(int));
    int rowblkblock = 0;
    unsigned int *blkcoostylerowidx = NULL;
    int *blkcoostylerowidx_colstart = NULL, *blkcoostylerowidx_colstop = NULL;
    tilespmv_prepare(matrixA, ptroffset1, ptroffset2, &rowblkblock, &blkcoostylerowidx, &blkcoostylerowidx_colstart,
                     &blkcoostylerowidx_colstop, rowA);

    MAT_VAL_TYPE *y = (MAT_VAL_TYPE *)calloc(rowA > 0 ? rowA : 1, sizeof(MAT_VAL_TYPE));
    call_tilespmv_cuda(filename, matrixA, ptroffset1, ptroffset2, rowblkblock, blkcoostylerowidx, blkcoostylerowidx_colstart,
                       blkcoostylerowidx_colstop, rowA, colA, nnzA, csrRowPtrA, csrColIdxA, csrValA, (MAT_VAL_TYPE)1.0, x, y,
                       y_golden);
    if (tilespmv_last_error()[0])
        printf("  %s\n", tilespmv_last_error());

    int errcount = 0;
    for (int i = 0; i < rowA; i++)
        if (fabs((double)(y_golden[i] - y[i])) > 0.01 * fabs((double)y[i]))
            errcount++;
    printf(errcount == 0 ? "Check... PASS!\n" : "Check... NO PASS! #err = %i\n", errcount);

    Tile_destroy(matrixA);
    free(matrixA);
    free(csrRowPtrA);
    free(csrColIdxA);
    free(csrValA);
    free(x);
    free(y);
    free(y_golden);
    free(ptroffset1);
    free(ptroffset2);
    free(blkcoostylerowidx);
    free(blkcoostylerowidx_colstart);
    free(blkcoostylerowidx_colstop);
    return 0;
}

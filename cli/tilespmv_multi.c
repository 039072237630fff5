/*
 * cli/tilespmv_multi.c -- a plain C host for the multi-GPU repeated-SpMV loop of the C-ABI (no torch, no MPI):
 *
 *     ./tilespmv_multi -n <ranks> [-k <iterations>] [-x nccl|fused|pipelined] [-s] <matrix.mtx>
 *
 * The reference driver is single-GPU (/root/reference/src/main.cu:74); this is what its main() would do with the
 * row-block sharding of SURVEY.md 8e.  The parent forks one process per rank; every rank selects GPU rank % (number of
 * GPUs) (-s: all ranks share GPU 0, for boxes with one GPU; NCCL cannot run that way), reads the matrix with
 * mmio_allinone, keeps the reference driver's data conventions (values i % 10 -- here scaled by 1/64 so that K
 * iterations stay finite --, rowA truncated to a multiple of 16, main.cu:68-71), takes its row block from
 * tilespmv_partition_rows, converts it with tilespmv_convert, joins the communicator and runs
 * tilespmv_dist_iterate.  Rank 0 checks the replicated result against the serial CSR loop of main.cu:101-110 applied K
 * times and prints the reference-style runtime line.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

#include "tilespmv.h"

extern int cudaSetDevice(int device);
extern int cudaGetDeviceCount(int *count);
extern int cudaMalloc(void **p, size_t bytes);
extern int cudaFree(void *p);
extern int cudaMemcpy(void *dst, const void *src, size_t bytes, int kind); /* 1 = H2D, 2 = D2H */
extern int cudaDeviceSynchronize(void);

static double now_s(void)
{
    struct timeval t;
    gettimeofday(&t, NULL);
    return (double)t.tv_sec + 1e-6 * (double)t.tv_usec;
}

static int run_rank(int rank, int nranks, int iters, int exchange, int share, const char *name, const char *filename)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != 0 || ndev < 1)
    {
        fprintf(stderr, "rank %d: no CUDA device (this library has no CPU fallback)\n", rank);
        return 3;
    }
    cudaSetDevice(share ? 0 : rank % ndev);
    int rowA, colA, nnzA, sym;
    int *rp, *ci;
    double *val;
    if (tilespmv_mmio_allinone_f64(&rowA, &colA, &nnzA, &sym, &rp, &ci, &val, filename) != 0)
    {
        fprintf(stderr, "rank %d: cannot read %s\n", rank, filename);
        return 4;
    }
    rowA = (rowA / TILESPMV_BLOCK_SIZE) * TILESPMV_BLOCK_SIZE; /* main.cu:71 */
    if (rowA != colA)
    {
        /* x <- A*x needs a square matrix: keep the leading square block */
        if (rank == 0)
            fprintf(stderr, "  note: %d x %d matrix, iterating on the leading %d columns only is not supported\n", rowA, colA, rowA);
        return 5;
    }
    for (int i = 0; i < nnzA; i++)
        val[i] = (double)(i % 10) / 64.0; /* main.cu:68-69, scaled */
    nnzA = rp[rowA];

    int *cuts = (int *)malloc((size_t)(nranks + 1) * sizeof(int));
    int64_t *cuts64 = (int64_t *)malloc((size_t)(nranks + 1) * sizeof(int64_t));
    if (tilespmv_partition_rows(TILESPMV_F64, rowA, rp, nranks, cuts) != TILESPMV_OK)
    {
        fprintf(stderr, "rank %d: %s\n", rank, tilespmv_last_error());
        return 6;
    }
    for (int r = 0; r <= nranks; r++)
        cuts64[r] = cuts[r];
    const int r0 = cuts[rank], r1 = cuts[rank + 1], m_local = r1 - r0;
    /* the rank's row block as a self-contained CSR with GLOBAL columns */
    int *lrp = (int *)malloc((size_t)(m_local + 1) * sizeof(int));
    for (int i = 0; i <= m_local; i++)
        lrp[i] = rp[r0 + i] - rp[r0];

    tilespmv_comm *comm = NULL;
    tilespmv_dmat *dm = NULL;
    tilespmv_dist *dist = NULL;
    int rc = tilespmv_comm_create(name, rank, nranks, exchange == TILESPMV_EXCHANGE_NCCL ? TILESPMV_COMM_NCCL : 0, &comm);
    if (rc == TILESPMV_OK)
        rc = tilespmv_convert(TILESPMV_F64, m_local, colA, lrp, ci + rp[r0], val + rp[r0], 0, &dm);
    if (rc == TILESPMV_OK)
        rc = tilespmv_dist_create(comm, dm, cuts64, NULL, 0, &dist);
    if (rc != TILESPMV_OK)
    {
        fprintf(stderr, "rank %d: set-up failed (%d): %s\n", rank, rc, tilespmv_last_error());
        return 7;
    }
    double *x = (double *)malloc((size_t)colA * sizeof(double));
    for (int i = 0; i < colA; i++)
        x[i] = (double)(i % 10); /* main.cu:93-97 */
    void *d_x0 = NULL;
    cudaMalloc(&d_x0, (size_t)colA * sizeof(double));
    cudaMemcpy(d_x0, x, (size_t)colA * sizeof(double), 1);

    /* warm-up + timed loop: every call restarts from x0, so the result below is K applications of A */
    rc = tilespmv_dist_iterate(dist, d_x0, iters, exchange, NULL);
    if (rc == TILESPMV_OK)
        rc = tilespmv_dist_sync(dist, NULL);
    tilespmv_comm_barrier(comm);
    const double t0 = now_s();
    if (rc == TILESPMV_OK)
        rc = tilespmv_dist_iterate(dist, d_x0, iters, exchange, NULL);
    if (rc == TILESPMV_OK)
        rc = tilespmv_dist_sync(dist, NULL);
    tilespmv_comm_barrier(comm);
    const double ms = (now_s() - t0) * 1e3 / (iters > 0 ? iters : 1);
    if (rc != TILESPMV_OK)
    {
        fprintf(stderr, "rank %d: iterate failed (%d): %s\n", rank, rc, tilespmv_last_error());
        return 8;
    }
    double *xk = (double *)malloc((size_t)colA * sizeof(double));
    cudaMemcpy(xk, tilespmv_dist_x(dist), (size_t)colA * sizeof(double), 2);

    int status = 0;
    if (rank == 0)
    {
        /* serial CSR loop (main.cu:101-110) applied K times; the data are integers / 64, so the sums are exact in fp64
         * for small K and the 1 % check of main.cu:186-197 is generous */
        double *a = x, *b = (double *)malloc((size_t)rowA * sizeof(double));
        for (int k = 0; k < iters; k++)
        {
            for (int i = 0; i < rowA; i++)
            {
                double sum = 0;
                for (int j = rp[i]; j < rp[i + 1]; j++)
                    sum += a[ci[j]] * val[j];
                b[i] = sum;
            }
            double *t = a;
            a = b;
            b = t;
        }
        int errcount = 0;
        for (int i = 0; i < rowA; i++)
            if (fabs(xk[i] - a[i]) > 1e-9 * fabs(a[i]) + 1e-300)
                errcount++;
        printf("  %d ranks, %d iterations of x <- A*x (%s exchange): %4.3f ms per iteration, %4.2f GFlops\n", nranks, iters,
               exchange == TILESPMV_EXCHANGE_NCCL ? "nccl" : (exchange == TILESPMV_EXCHANGE_FUSED ? "fused" : "pipelined"), ms,
               2.0 * (double)nnzA * 1e-6 / ms);
        if (errcount == 0)
            printf("  Check... PASS!\n");
        else
            printf("  Check... NO PASS! error = %d\n", errcount);
        status = errcount ? 9 : 0;
    }
    tilespmv_dist_destroy(dist);
    tilespmv_dmat_destroy(dm);
    tilespmv_comm_destroy(comm);
    cudaFree(d_x0);
    return status;
}

int main(int argc, char **argv)
{
    int nranks = 2, iters = 3, exchange = TILESPMV_EXCHANGE_PIPELINED, share = 0;
    const char *filename = NULL;
    for (int i = 1; i < argc; i++)
    {
        if (!strcmp(argv[i], "-n") && i + 1 < argc)
            nranks = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-k") && i + 1 < argc)
            iters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-s"))
            share = 1;
        else if (!strcmp(argv[i], "-x") && i + 1 < argc)
        {
            const char *m = argv[++i];
            exchange = !strcmp(m, "nccl") ? TILESPMV_EXCHANGE_NCCL : (!strcmp(m, "fused") ? TILESPMV_EXCHANGE_FUSED : TILESPMV_EXCHANGE_PIPELINED);
        }
        else
            filename = argv[i];
    }
    if (!filename || nranks < 1 || nranks > 8)
    {
        fprintf(stderr, "usage: %s -n <ranks 1..8> [-k iterations] [-x nccl|fused|pipelined] [-s] <matrix.mtx>\n", argv[0]);
        return 1;
    }
    printf("-------------- %s --------------\n", filename);
    fflush(stdout);
    char name[64];
    snprintf(name, sizeof(name), "cli_%d", (int)getpid()); /* unique per job: the parent's pid */
    /* fork BEFORE any CUDA call: every rank is its own process with its own CUDA context */
    pid_t pids[8];
    for (int r = 0; r < nranks; r++)
    {
        pids[r] = fork();
        if (pids[r] == 0)
        {
            const int code = run_rank(r, nranks, iters, exchange, share, name, filename);
            fflush(stdout);
            fflush(stderr);
            _exit(code);
        }
        if (pids[r] < 0)
        {
            perror("fork");
            return 2;
        }
    }
    int worst = 0;
    for (int r = 0; r < nranks; r++)
    {
        int st = 0;
        waitpid(pids[r], &st, 0);
        const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 100;
        if (code > worst)
            worst = code;
    }
    return worst;
}

/*
 * tilespmv.h -- C-ABI of the B200-native TileSpMV (libtilespmv_b200.so).
 *
 * Drop-in boundary for the hot path of SuperScientificSoftwareLaboratory/TileSpMV.  The reference
 * has no library: its "API" is a set of C functions defined in headers and included once into
 * src/main.cu (main.cu:1-7).  This header re-declares that surface -- the Tile_matrix struct and
 * the three entry points the driver calls -- as exported symbols, and adds the plan/handle API
 * the reference lacks (repeated SpMV on device-resident data, multi-GPU row blocks).
 *
 *   reference (file:line)                              this library
 *   -------------------------------------------------  ------------------------------------------
 *   struct Tile_matrix            format.h:3-56        Tile_matrix_f64 / Tile_matrix_f32
 *   Tile_create                   csr2tile.h:629-635   Tile_create_f64 / _f32   (GPU conversion)
 *   Tile_destroy                  format.h:58-94       Tile_destroy_f64 / _f32  (frees everything)
 *   tilespmv_cpu (schedule part)  tilespmv_cpu.h:68-118, :142-257
 *                                                      tilespmv_prepare_f64 / _f32 (ptroffset1/2 +
 *                                                      warp-chunk schedule; NO CPU SpMV: the
 *                                                      product has no CPU fallback)
 *   call_tilespmv_cuda            tilespmv_cuda.h:794-809
 *                                                      call_tilespmv_cuda_f64 / _f32
 *   compile-time constants        common.h:12-63       TILESPMV_* macros below
 *
 * Precision is a compile-time macro in the reference (MAT_VAL_TYPE, common.h:12-14; float via
 * -D, Makefile:5,22), so a shared library has to export two symbol sets.  Source-level drop-in:
 * compile the caller with -DMAT_VAL_TYPE=double|float and include this header; the unsuffixed
 * reference names then map to the matching set (bottom of this file).
 *
 * All functions are thread-compatible (one plan per host thread), keep no hidden global state
 * besides a thread-local last-error string, and use the CUDA device current on the calling
 * thread (the reference selects it with cudaSetDevice in main.cu:74).
 * No torch / C++ types appear in any signature.
 */
#ifndef TILESPMV_H
#define TILESPMV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants that are part of the format contract (common.h:12-63) ---- */
#define TILESPMV_BLOCK_SIZE 16      /* BLOCK_SIZE        */
#define TILESPMV_COO_NNZ_TH 12      /* COO_NNZ_TH        */
#define TILESPMV_PREFETCH_SMEM_TH 4 /* PREFETCH_SMEM_TH  */
#define TILESPMV_NUM_F 240          /* num_f (0xF0)      */
#define TILESPMV_NUM_B 15           /* num_b (0x0F)      */
#define TILESPMV_BENCH_REPEAT 1000  /* BENCH_REPEAT      */
#define TILESPMV_WARMUP_NUM 200     /* WARMUP_NUM        */

/* tile format codes stored in Tile_matrix.Format (csr2tile.h:154,162,193,235,272,310,319) */
#define TILESPMV_FMT_CSR 0
#define TILESPMV_FMT_COO 1
#define TILESPMV_FMT_ELL 2
#define TILESPMV_FMT_HYB 3
#define TILESPMV_FMT_DENSE 4
#define TILESPMV_FMT_DENSEROW 5
#define TILESPMV_FMT_DENSECOL 6

/* ---- status codes (0 = ok, negative = error; cf. ANONYMOUSLIB_* in CSR5 detail/common.h:13-18) ---- */
#define TILESPMV_OK 0
#define TILESPMV_ERR_INVALID (-1)     /* bad argument / inconsistent sizes            */
#define TILESPMV_ERR_CUDA (-2)        /* a CUDA runtime call or kernel failed         */
#define TILESPMV_ERR_ALLOC (-3)       /* host or device allocation failed             */
#define TILESPMV_ERR_UNSUPPORTED (-4) /* e.g. sizes that overflow the int-indexed format */
#define TILESPMV_ERR_NODEVICE (-5)    /* no CUDA device: there is NO CPU fallback     */
#define TILESPMV_ERR_IO (-6)

/* ---- Tile_matrix: field-for-field the reference struct (format.h:3-56) ---- */
#define TILESPMV_DECLARE_TILE_MATRIX(NAME, VAL_T)                                               \
    typedef struct                                                                              \
    {                                                                                           \
        int tilem;                                                                              \
        int tilen;                                                                              \
        int tilenum;                                                                            \
        int *tile_ptr;                 /* [tilem+1]  level-1 CSR of tiles                     */ \
        int *tile_columnidx;           /* [tilenum]  ascending inside a block row             */ \
        int *tile_nnz;                 /* [tilenum+1] exclusive prefix of true nnz            */ \
        char *Format;                  /* [tilenum]  TILESPMV_FMT_*                           */ \
        int *blknnz;                   /* [tilenum+1] exclusive prefix of stored slots        */ \
        unsigned char *blknnznnz;      /* [tilenum+1] per-tile slots, wrapped to 8 bits       */ \
        int *dnsrowptr;                /* [tilenum+1]                                         */ \
        int *dnscolptr;                /* [tilenum+1]                                         */ \
        char *tilewidth;               /* [tilenum]  ELL / HYB width                          */ \
        int *csr_offset;               /* [tilenum+1] prefix offsets into the per-format arrays */ \
        int *csrptr_offset;                                                                     \
        int *coo_offset;                                                                        \
        int *ell_offset;                                                                        \
        int *hyb_offset;                                                                        \
        int *hyb_coocount;                                                                      \
        int *dns_offset;                                                                        \
        int *dnsrow_offset;                                                                     \
        int *dnscol_offset;                                                                     \
        int *new_coocount;                                                                      \
        VAL_T *Blockcsr_Val;                                                                    \
        unsigned char *Blockcsr_Ptr;                                                            \
        unsigned char *csr_compressedIdx;                                                       \
        int csrsize;                                                                            \
        int csrptrlen;                                                                          \
        VAL_T *Blockcoo_Val;                                                                    \
        unsigned char *coo_compressed_Idx;                                                      \
        int coosize;                                                                            \
        VAL_T *Blockell_Val;                                                                    \
        unsigned char *ell_compressedIdx;                                                       \
        int ellsize;                                                                            \
        VAL_T *Blockhyb_Val;                                                                    \
        unsigned char *hybIdx;                                                                  \
        int hybsize;                                                                            \
        int hybellsize;                                                                         \
        int hybcoosize;                                                                         \
        VAL_T *Blockdense_Val;                                                                  \
        int dnssize;                                                                            \
        VAL_T *Blockdenserow_Val;                                                               \
        char *denserowid;                                                                       \
        int dnsrowsize;                                                                         \
        VAL_T *Blockdensecol_Val;                                                               \
        char *densecolid;                                                                       \
        int dnscolsize;                                                                         \
        int coototal;                                                                           \
        VAL_T *deferredcoo_val;        /* side CSR of the very sparse (COO) tiles             */ \
        int *deferredcoo_colidx;       /* GLOBAL columns, ascending inside a row              */ \
        int *deferredcoo_ptr;          /* [rowA+1]                                            */ \
    } NAME

TILESPMV_DECLARE_TILE_MATRIX(Tile_matrix_f64, double);
TILESPMV_DECLARE_TILE_MATRIX(Tile_matrix_f32, float);

/* =============================== drop-in entry points =================================== */

/*
 * Tile_create (csr2tile.h:629-635).  Same contract: the caller allocates the struct and keeps
 * ownership of the (host) CSR arrays; every array in the struct is malloc'ed here; rowA / colA
 * need not be multiples of 16; prints "\n  The number of tile = %i\n" like the reference (:661).
 * The conversion itself runs on the GPU (tile discovery, format selection, scatter, nibble
 * packing, side-CSR extraction) and the result is bit-exact with the reference CPU conversion.
 * The reference returns void and has no error path; on failure this leaves tilenum = -1 and the
 * reason in tilespmv_last_error().
 */
void Tile_create_f64(Tile_matrix_f64 *matrix, int rowA, int colA, int nnzA, int *csrRowPtrA,
                     int *csrColIdxA, double *csrValA);
void Tile_create_f32(Tile_matrix_f32 *matrix, int rowA, int colA, int nnzA, int *csrRowPtrA,
                     int *csrColIdxA, float *csrValA);

/* Tile_destroy (format.h:58-94): frees every array (the reference leaks four) and zeroes the
 * struct; like the reference it does not free the struct itself. */
void Tile_destroy_f64(Tile_matrix_f64 *matrix);
void Tile_destroy_f32(Tile_matrix_f32 *matrix);

/*
 * The bookkeeping half of tilespmv_cpu (tilespmv_cpu.h:68-118 and the ptroffset writes at
 * :142-257): fills the caller-allocated ptroffset1/2[tilenum] and returns the warp-chunk
 * schedule through malloc'ed arrays (caller frees with free()).  The reference routes these
 * through its CPU SpMV; this library needs none of them (call_tilespmv_cuda below ignores them)
 * but they stay derivable for callers and parity tests.  Returns a status code.
 */
int tilespmv_prepare_f64(const Tile_matrix_f64 *matrix, int *ptroffset1, int *ptroffset2,
                         int *rowblkblock, unsigned int **blkcoostylerowidx,
                         int **blkcoostylerowidx_colstart, int **blkcoostylerowidx_colstop, int rowA);
int tilespmv_prepare_f32(const Tile_matrix_f32 *matrix, int *ptroffset1, int *ptroffset2,
                         int *rowblkblock, unsigned int **blkcoostylerowidx,
                         int **blkcoostylerowidx_colstart, int **blkcoostylerowidx_colstop, int rowA);

/*
 * call_tilespmv_cuda (tilespmv_cuda.h:794-809).  All pointers are HOST pointers; the callee owns
 * every device allocation for the duration of the call; y[rowA] receives A*x; alpha is accepted
 * and ignored exactly like the reference (csr5_spmv_cuda.h:22); ptroffset*, the schedule arrays,
 * csr* and y_golden are accepted for signature compatibility and unused.  Follows the reference
 * protocol: WARMUP_NUM warm-ups then BENCH_REPEAT timed SpMVs (CUDA events around the batch),
 * prints "  CUDA SpMV runtime %4.2f ms, %4.2f GFlops\n\n" (:1139) and appends
 * "filename,rowA,colA,nnzA,ms,gflops" to ./results.csv (:1142-1147).
 * TILESPMV_BENCH_REPEAT / TILESPMV_WARMUP_NUM environment variables override the counts.
 * The reference returns void; errors are reported through tilespmv_last_error() and y is left
 * untouched.
 */
void call_tilespmv_cuda_f64(char *filename, Tile_matrix_f64 *matrix, int *ptroffset1, int *ptroffset2,
                            int rowblkblock, unsigned int *blkcoostylerowidx,
                            int *blkcoostylerowidx_colstart, int *blkcoostylerowidx_colstop, int rowA,
                            int colA, int nnzA, int *csrRowPtrA, int *csrColIdxA, double *csrValA,
                            double alpha, double *x, double *y, double *y_golden);
void call_tilespmv_cuda_f32(char *filename, Tile_matrix_f32 *matrix, int *ptroffset1, int *ptroffset2,
                            int rowblkblock, unsigned int *blkcoostylerowidx,
                            int *blkcoostylerowidx_colstart, int *blkcoostylerowidx_colstop, int rowA,
                            int colA, int nnzA, int *csrRowPtrA, int *csrColIdxA, float *csrValA,
                            float alpha, float *x, float *y, float *y_golden);

/* ================================ handle / plan API ====================================== */

/* a Tile_matrix resident in device memory (opaque) */
typedef struct tilespmv_dmat tilespmv_dmat;
/* a packed, scheduled SpMV plan for one dmat on one GPU (opaque) */
typedef struct tilespmv_plan tilespmv_plan;

#define TILESPMV_F64 8
#define TILESPMV_F32 4

/* flags of tilespmv_convert */
#define TILESPMV_CSR_ON_DEVICE 1 /* rowptr / colidx / val are device pointers */
/* Non-default: switch on the reference's dormant HYB selection rule (csr2tile.h:279-316, commented out
 * upstream): a tile that would be CSR becomes HYB (format 3) when its row-length variation is >= 1.0 and the
 * I/O-cost walk leaves <= 4 spilled entries.  The result is bit-exact with the reference built with that
 * rule un-commented (oracle/Makefile, target ref_hyb).  Without this flag format 3 is never produced. */
#define TILESPMV_ENABLE_HYB 2

/*
 * GPU csr2tile: CSR (host pointers, or device pointers with TILESPMV_CSR_ON_DEVICE) -> a
 * device-resident Tile_matrix.  precision is TILESPMV_F64 or TILESPMV_F32 and tells the type
 * behind val.  Rows >= rowA present in rowptr are ignored (the reference driver truncates rowA to
 * a multiple of 16 and keeps the CSR arrays, main.cu:71).
 */
int tilespmv_convert(int precision, int rowA, int colA, const int *rowptr, const int *colidx,
                     const void *val, unsigned flags, tilespmv_dmat **out);
/* upload an existing host Tile_matrix (e.g. one produced by the reference's CPU Tile_create) */
int tilespmv_dmat_upload_f64(const Tile_matrix_f64 *matrix, int rowA, int colA, tilespmv_dmat **out);
int tilespmv_dmat_upload_f32(const Tile_matrix_f32 *matrix, int rowA, int colA, tilespmv_dmat **out);
/* download into a caller-allocated struct; arrays are malloc'ed (free with Tile_destroy_*) */
int tilespmv_dmat_export_f64(const tilespmv_dmat *dm, Tile_matrix_f64 *matrix);
int tilespmv_dmat_export_f32(const tilespmv_dmat *dm, Tile_matrix_f32 *matrix);
void tilespmv_dmat_destroy(tilespmv_dmat *dm);

typedef struct
{
    int precision;        /* TILESPMV_F64 / TILESPMV_F32 */
    int rowA, colA;
    int tilem, tilen, tilenum;
    int64_t nnz;          /* true nonzeros */
    int64_t nnz_side;     /* coototal: nonzeros served from the extracted side CSR */
    int64_t tiles_by_format[7];
    int64_t device_bytes; /* bytes of device memory held by the dmat */
    int64_t slots_by_format[7]; /* stored value slots per format (csrsize, coosize, ellsize, hybsize, dnssize,
                                   dnsrowsize, dnscolsize of format.h): slots - nonzeros = padding of the format */
} tilespmv_dmat_info;
int tilespmv_dmat_get_info(const tilespmv_dmat *dm, tilespmv_dmat_info *info);

typedef struct
{
    int chunk_bytes;  /* max bytes of one scheduler chunk of the packed stream (0 = default) */
    int xstage_bytes; /* max bytes of x staged in shared memory per chunk (0 = default)      */
    int ctas_per_sm;  /* persistent CTAs per SM (0 = default: 1, with as many warps as fit)  */
    int stages;       /* TMA pipeline depth per warp, 2..4 (0 = default: 2)                  */
    int max_warps;    /* cap on warps per CTA (0 = as many as shared memory holds, <= 20)    */
    int flags;        /* TILESPMV_PLAN_* bits (0 = default)                                  */
    int xpanel_bytes; /* bytes of x per column panel of the extracted (side) matrix: 0 = automatic
                         (panels only when x is far larger than L2 and side entries dominate),
                         > 0 = that width, < 0 = never.  With panels one SpMV is one launch per
                         panel, each gathering from an L2-resident window of x.                */
    int format_mask;  /* per-format cost profiling (cf. DEBUG_FORMATCOST / formatprofile, tilespmv_cuda.h:102-111,
                         main.cu:12): 0 = the whole matrix; else bit f (TILESPMV_FMT_*) keeps the tiles of format f, bit 1
                         (COO) the extracted side entries (COO tiles + HYB spill-over).  The plans of the seven single
                         bits partition the nonzeros: their y's add up to A*x. */
} tilespmv_plan_options;
/* keep every CSR tile an individual tile of the packed stream instead of merging the CSR tiles of a block row
 * into one jagged slot-row list (the default, faster; results agree to rounding) */
#define TILESPMV_PLAN_NO_CSR_GROUPS 1
/* pack chunks that hold only extracted (side) entries like every other chunk instead of in the flat jagged-diagonal
 * layout (the default for them, several times fewer instructions per row; results agree to rounding) */
#define TILESPMV_PLAN_NO_FLAT_SIDE 2

/* Packs the tiles into the 16-byte-aligned per-chunk stream and builds the persistent,
 * byte-balanced chunk schedule.  opts may be NULL. */
int tilespmv_plan_create(const tilespmv_dmat *dm, const tilespmv_plan_options *opts, tilespmv_plan **out);
void tilespmv_plan_destroy(tilespmv_plan *plan);
/* Binary cache of a packed plan (incl. its x-panel sub-plans): a later run loads it instead of converting and planning
 * again (the reference re-converts and re-uploads the matrix on every run, tilespmv_cuda.h:794-1056).  The file is tied
 * to the launch shape it was packed for: tilespmv_plan_load returns TILESPMV_ERR_UNSUPPORTED on a GPU with another SM
 * count / shared-memory size (plan again), TILESPMV_ERR_IO for a missing, truncated or corrupt file (checksummed). */
int tilespmv_plan_save(const tilespmv_plan *plan, const char *path);
int tilespmv_plan_load(const char *path, tilespmv_plan **out);

/* y = A*x with DEVICE pointers (16-byte aligned), asynchronous on `stream` (a cudaStream_t
 * passed as void*; NULL = default stream).  x has colA entries, y has rowA entries. */
int tilespmv_plan_spmv(tilespmv_plan *plan, const void *d_x, void *d_y, void *stream);
/* y = A*x with HOST pointers: H2D of x, the SpMV, D2H of y, synchronous (the end-to-end path). */
int tilespmv_plan_spmv_host(tilespmv_plan *plan, const void *x, void *y);
/* y[i] = A*x[i] for nvec HOST vectors (x[i] has colA entries, y[i] rowA; pinned memory for full overlap),
 * pipelined over a ring of device buffers: the H2D copy of vector i+1, the SpMV of vector i and the D2H copy of
 * vector i-1 run concurrently on three streams.  Synchronous: returns when every y[i] is complete.  The reference
 * has no counterpart (call_tilespmv_cuda, tilespmv_cuda.h:794, re-uploads the whole matrix for every call). */
int tilespmv_plan_spmv_host_batch(tilespmv_plan *plan, int nvec, const void *const *x, void *const *y);

/*
 * Repeated SpMV on ONE GPU, x <- A*x for a square matrix: iteration i reads d_xa (i even) or d_xb (i odd) and writes
 * the other buffer, so the result is in d_xa when niters is even, else in d_xb.  The niters launches are captured once
 * into a CUDA graph (re-instantiated only when the buffers or niters change) and replayed with a single graph launch on
 * `stream`: no per-iteration CPU launch cost, which matters for matrices whose SpMV takes a few microseconds.
 * Both buffers are DEVICE pointers with rowA = colA entries, 16-byte aligned.  Not available while peers are set.
 */
int tilespmv_plan_iterate(tilespmv_plan *plan, void *d_xa, void *d_xb, int niters, void *stream);

/*
 * Row-block sharding for nparts GPUs (host-only, needs no device): contiguous ranges of block rows cut where the prefix
 * of streamed bytes per block row (the weights of B_alg) crosses g / nparts of the total, always at multiples of 16 rows
 * so that tiles never straddle GPUs.  row_cuts[nparts + 1] receives the first row of every part and rowA at the end;
 * part g converts / plans rows [row_cuts[g], row_cuts[g+1]) with global column indices.  The reference is single-GPU.
 */
int tilespmv_partition_rows(int precision, int rowA, const int *rowptr, int nparts, int *row_cuts);

/*
 * Multi-GPU repeated SpMV (row-block sharding, x replicated): after computing its rows the
 * kernel also stores them straight into x_next of every peer (P2P-mapped pointers over NVLink)
 * at row_offset, so the all-gather of the next x is the kernel's own store stream.
 * peer_x[i] (i < npeers) are device pointers valid on THIS device (cudaIpcOpenMemHandle /
 * peer access enabled by the caller); pass npeers = 0 to switch the fused epilogue off.
 */
int tilespmv_plan_set_peers(tilespmv_plan *plan, int npeers, void *const *peer_x, int64_t row_offset);

/* ============================ multi-GPU repeated SpMV (one process per GPU) ============================
 *
 * The reference is single-GPU (main.cu:74 selects one device).  This part implements what SURVEY.md 8b / 8e ask for:
 * the matrix cut into contiguous row blocks (tilespmv_partition_rows), one rank = one process = one GPU of the box, x
 * replicated, and the loop x <- A*x with the y slices all-gathered into the next x every iteration.  No torch, no MPI:
 *
 *   tilespmv_comm_create   rendezvous of the ranks through a POSIX shared-memory segment called `name` (all ranks pass
 *                          the same name, unique per job; rank 0 creates it).  With TILESPMV_COMM_NCCL the library also
 *                          creates an NCCL communicator (the unique id travels through the segment).
 *   tilespmv_dist_create   collective over the communicator.  `local_rows` is THIS rank's shard (rows
 *                          [row_cuts[rank], row_cuts[rank+1]) with GLOBAL column indices, converted with
 *                          tilespmv_convert); row_cuts[nranks + 1] are the cuts of tilespmv_partition_rows.  Builds the
 *                          rank's plan and two replicated x buffers that every peer maps through CUDA IPC (NVLink P2P).
 *   tilespmv_dist_iterate  niters times x <- A*x, asynchronous on `stream`.  d_x0 = the replicated start vector (device
 *                          pointer, colA values, the same on every rank) or NULL to continue from the current x.
 *   tilespmv_dist_x        device pointer of the current replicated x (valid until the next iterate call).
 *   tilespmv_dist_sync     waits for `stream` and the library's copy stream; reports a peer that never arrived (the
 *                          device-side waits give up after TILESPMV_COMM_SPIN_TIMEOUT_S, default 30 s) as an error.
 *
 * Exchanges (all give bitwise identical x: the arithmetic is the same plan):
 *   TILESPMV_EXCHANGE_NCCL       SpMV, then one in-place ncclAllGather (equal slices) or a grouped ncclBroadcast per rank
 *   TILESPMV_EXCHANGE_FUSED      the SpMV kernel's epilogue stores y into every peer's next x over NVLink; one flag
 *                                barrier per iteration
 *   TILESPMV_EXCHANGE_PIPELINED  copy engines push the slice to the peers in the order they need it; every launch of the
 *                                next iteration waits only for the slices its columns read (x panels cut at the row blocks
 *                                of the ranks, own panel first), so the exchange runs under the next iteration's compute
 *   TILESPMV_EXCHANGE_HALO       for matrices whose row blocks read only a small window of x beyond their own slice (bands,
 *                                stencils): the kernel's epilogue stores just the rows a peer's next launch reads into that
 *                                peer's next x, a flag exchange with those neighbours orders the iterations, and the copy
 *                                engines replicate the rest of every slice in the background (still a full all-gather of x
 *                                per iteration, complete on every rank when the call ends).  Falls back to PIPELINED when
 *                                some rank would have to store more than half of its slice that way.
 * All ranks must make the same sequence of calls with the same niters / exchange.  A tilespmv_dist must be destroyed
 * before its communicator (destroy is collective too).
 */
typedef struct tilespmv_comm tilespmv_comm;
typedef struct tilespmv_dist tilespmv_dist;
#define TILESPMV_COMM_NCCL 1u          /* flag of tilespmv_comm_create: also create an NCCL communicator */
#define TILESPMV_DIST_UNIFORM_PANELS 1u /* flag of tilespmv_dist_create: keep the single-GPU x-panel cuts (no rank-aligned panels) */
#define TILESPMV_EXCHANGE_NCCL 0
#define TILESPMV_EXCHANGE_FUSED 1
#define TILESPMV_EXCHANGE_PIPELINED 2
#define TILESPMV_EXCHANGE_HALO 3
int tilespmv_comm_create(const char *name, int rank, int nranks, unsigned flags, tilespmv_comm **out);
void tilespmv_comm_destroy(tilespmv_comm *comm);
int tilespmv_comm_barrier(tilespmv_comm *comm); /* host barrier over all ranks */
int tilespmv_dist_create(tilespmv_comm *comm, const tilespmv_dmat *local_rows, const int64_t *row_cuts,
                         const tilespmv_plan_options *opts, unsigned flags, tilespmv_dist **out);
void tilespmv_dist_destroy(tilespmv_dist *dist);
int tilespmv_dist_iterate(tilespmv_dist *dist, const void *d_x0, int niters, int exchange, void *stream);
void *tilespmv_dist_x(tilespmv_dist *dist);
tilespmv_plan *tilespmv_dist_plan(tilespmv_dist *dist); /* the rank's plan (borrowed): single SpMV, info */
int tilespmv_dist_sync(tilespmv_dist *dist, void *stream);
typedef struct
{
    int rank, nranks;
    int64_t row0, rows;      /* this rank's row block                                              */
    int64_t slice_bytes;     /* bytes of y this rank contributes to every all-gather               */
    int64_t device_bytes;
    int launch_units;        /* kernel launches per SpMV that can wait for different slices of x   */
    int equal_slices;        /* all row blocks have the same length (NCCL: one ncclAllGather)      */
    uint32_t unit_deps[64];  /* per launch unit: bit mask of the ranks whose slices of x it reads  */
    int halo_eligible;       /* TILESPMV_EXCHANGE_HALO applies (else it runs PIPELINED)            */
    int64_t need_lo, need_hi; /* x columns [need_lo, need_hi) this rank's launches read             */
} tilespmv_dist_info;
int tilespmv_dist_get_info(const tilespmv_dist *dist, tilespmv_dist_info *info);

typedef struct
{
    int precision;
    int64_t nchunks;           /* scheduler chunks                                        */
    int64_t stream_bytes;      /* bytes of the packed stream read by one SpMV             */
    int64_t algorithmic_bytes; /* B_alg of SURVEY.md 8(d) for this matrix                 */
    int64_t csr_bytes;         /* B_csr (cross-format comparison figure)                  */
    int64_t split_rows;        /* block rows cut across chunks (fixed up by a 2nd launch) */
    int launches_per_spmv;     /* kernels launched by one tilespmv_plan_spmv              */
    int grid, block, smem_bytes;
    int chunk_bytes, xstage_bytes;
    int64_t device_bytes;
    int64_t csr_groups;        /* block rows whose CSR tiles were merged into one CSR group */
    int64_t xpanels;           /* column panels of the side matrix (1 = none; see xpanel_bytes) */
} tilespmv_plan_info;
int tilespmv_plan_get_info(const tilespmv_plan *plan, tilespmv_plan_info *info);

/*
 * Per-format cost profile (the reference's DEBUG_FORMATCOST build, tilespmv_cuda.h:102-111: its kernel restricted to one
 * format, formatprofile = -1 all / 0..6 / 7 none).  Builds one plan per format present in dm (+ the whole matrix + an
 * empty plan that only writes zeros) and times `iters` SpMVs of each on the device vectors d_x / d_y:
 *   ms[f], f = 0..6   the tiles of format f only (f = 1: the extracted side entries); 0 when the format is absent
 *   ms[7]             no format at all: the fixed cost of visiting every block row and writing y
 *   ms[8]             the whole matrix
 * nnz[f] (may be NULL) receives the nonzeros each of those plans multiplies (nnz[7] = 0, nnz[8] = all).
 */
int tilespmv_format_profile(const tilespmv_dmat *dm, const tilespmv_plan_options *opts, const void *d_x, void *d_y, int warmup,
                            int iters, double ms[9], int64_t nnz[9]);

/* Time `iters` back-to-back SpMVs on device buffers with CUDA events on `stream` (after
 * `warmup` untimed ones); returns mean milliseconds per SpMV in *ms_per_spmv. */
int tilespmv_plan_time(tilespmv_plan *plan, const void *d_x, void *d_y, int warmup, int iters,
                       void *stream, double *ms_per_spmv);

/* Matrix Market front end with the semantics of mmio_allinone (mmio_highlevel.h:593-759):
 * returns 0 / -1 (open) / -2 (banner) / -4 (size line); outputs malloc'ed, caller frees. */
int tilespmv_mmio_allinone_f64(int *m, int *n, int *nnz, int *isSymmetric, int **csrRowPtr,
                               int **csrColIdx, double **csrVal, const char *filename);
int tilespmv_mmio_allinone_f32(int *m, int *n, int *nnz, int *isSymmetric, int **csrRowPtr,
                               int **csrColIdx, float **csrVal, const char *filename);

/* last error message of the calling thread ("" if none) and library version */
const char *tilespmv_last_error(void);
const char *tilespmv_version(void);
/* number of CUDA kernels this library has launched in this process (instrumentation) */
int64_t tilespmv_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif

/* ---- source-level drop-in: the reference's unsuffixed names (format.h, csr2tile.h, ...) ---- */
#ifdef TILESPMV_REFERENCE_NAMES
#ifndef MAT_VAL_TYPE
#define MAT_VAL_TYPE double
#endif
#ifndef MAT_PTR_TYPE
#define MAT_PTR_TYPE int
#endif
#ifndef BLOCK_SIZE
#define BLOCK_SIZE TILESPMV_BLOCK_SIZE
#endif
#define TILESPMV_CAT_(a, b) a##b
#define TILESPMV_CAT(a, b) TILESPMV_CAT_(a, b)
#ifdef TILESPMV_USE_F32
#define TILESPMV_SUFFIX _f32
#else
#define TILESPMV_SUFFIX _f64
#endif
#define Tile_matrix TILESPMV_CAT(Tile_matrix, TILESPMV_SUFFIX)
#define Tile_create TILESPMV_CAT(Tile_create, TILESPMV_SUFFIX)
#define Tile_destroy TILESPMV_CAT(Tile_destroy, TILESPMV_SUFFIX)
#define call_tilespmv_cuda TILESPMV_CAT(call_tilespmv_cuda, TILESPMV_SUFFIX)
#define tilespmv_prepare TILESPMV_CAT(tilespmv_prepare, TILESPMV_SUFFIX)
#define mmio_allinone TILESPMV_CAT(tilespmv_mmio_allinone, TILESPMV_SUFFIX)
#endif

#endif /* TILESPMV_H */

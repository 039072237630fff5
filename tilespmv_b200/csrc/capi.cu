// capi.cu -- the extern "C" surface of include/tilespmv.h: error state, dmat upload / export,
// the drop-in entry points (Tile_create, Tile_destroy, tilespmv_prepare, call_tilespmv_cuda),
// the plan API and the Matrix Market front end.
#include <omp.h>
#include <sys/stat.h>
#include <sys/time.h>

#include <cctype>
#include <cmath>
#include <charconv>
#include <new>

#include "plan.cuh"

namespace tsp
{

static thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}
const char *last_error() { return g_error.c_str(); }
void clear_error() { g_error.clear(); }

int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
    {
        cudaGetLastError();
        set_error("no CUDA device available (%s): tilespmv_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return TILESPMV_ERR_NODEVICE;
    }
    return TILESPMV_OK;
}

// ---------------------------------------------------------------------------------------------
// dmat <-> host Tile_matrix
// ---------------------------------------------------------------------------------------------
template <class TM>
static void zero_tile_matrix(TM *m)
{
    memset(m, 0, sizeof(TM));
}

template <class U>
static int download(const DevBuf &b, size_t count, U **out)
{
    size_t bytes = count * sizeof(U);
    U *h = static_cast<U *>(malloc(bytes ? bytes : 1));
    if (!h)
    {
        set_error("malloc(%zu) failed", bytes);
        return TILESPMV_ERR_ALLOC;
    }
    if (bytes)
    {
        cudaError_t e = cudaMemcpy(h, b.p, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess)
        {
            free(h);
            set_error("cudaMemcpy D2H failed: %s", cudaGetErrorString(e));
            return TILESPMV_ERR_CUDA;
        }
    }
    *out = h;
    return TILESPMV_OK;
}

template <class TM, class T>
static int dmat_export(const tilespmv_dmat *d, TM *m)
{
    if (!d || !m)
    {
        set_error("export: null argument");
        return TILESPMV_ERR_INVALID;
    }
    if (d->precision != (int)sizeof(T))
    {
        set_error("export: precision mismatch");
        return TILESPMV_ERR_INVALID;
    }
    zero_tile_matrix(m);
    const size_t T_ = (size_t)d->tilenum;
    m->tilem = d->tilem;
    m->tilen = d->tilen;
    m->tilenum = d->tilenum;
    m->csrsize = d->csrsize;
    m->csrptrlen = d->csrptrlen;
    m->coosize = d->coosize;
    m->ellsize = d->ellsize;
    m->hybsize = d->hybsize;
    m->hybellsize = d->hybellsize;
    m->hybcoosize = d->hybcoosize;
    m->dnssize = d->dnssize;
    m->dnsrowsize = d->dnsrowsize;
    m->dnscolsize = d->dnscolsize;
    m->coototal = d->coototal;
    TSP_TRY(download(d->tile_ptr, (size_t)d->tilem + 1, &m->tile_ptr));
    TSP_TRY(download(d->tile_columnidx, T_, &m->tile_columnidx));
    TSP_TRY(download(d->tile_nnz, T_ + 1, &m->tile_nnz));
    TSP_TRY(download(d->Format, T_, &m->Format));
    TSP_TRY(download(d->blknnz, T_ + 1, &m->blknnz));
    TSP_TRY(download(d->blknnznnz, T_ + 1, &m->blknnznnz));
    TSP_TRY(download(d->dnsrowptr, T_ + 1, &m->dnsrowptr));
    TSP_TRY(download(d->dnscolptr, T_ + 1, &m->dnscolptr));
    TSP_TRY(download(d->tilewidth, T_, &m->tilewidth));
    TSP_TRY(download(d->csr_offset, T_ + 1, &m->csr_offset));
    TSP_TRY(download(d->csrptr_offset, T_ + 1, &m->csrptr_offset));
    TSP_TRY(download(d->coo_offset, T_ + 1, &m->coo_offset));
    TSP_TRY(download(d->ell_offset, T_ + 1, &m->ell_offset));
    TSP_TRY(download(d->hyb_offset, T_ + 1, &m->hyb_offset));
    TSP_TRY(download(d->hyb_coocount, T_ + 1, &m->hyb_coocount));
    TSP_TRY(download(d->dns_offset, T_ + 1, &m->dns_offset));
    TSP_TRY(download(d->dnsrow_offset, T_ + 1, &m->dnsrow_offset));
    TSP_TRY(download(d->dnscol_offset, T_ + 1, &m->dnscol_offset));
    TSP_TRY(download(d->new_coocount, T_ + 1, &m->new_coocount));
    TSP_TRY(download(d->Blockcsr_Val, (size_t)d->csrsize, &m->Blockcsr_Val));
    TSP_TRY(download(d->Blockcsr_Ptr, (size_t)d->csrptrlen, &m->Blockcsr_Ptr));
    TSP_TRY(download(d->csr_compressedIdx, ((size_t)d->csrsize + 1) / 2, &m->csr_compressedIdx));
    TSP_TRY(download(d->Blockcoo_Val, (size_t)d->coosize, &m->Blockcoo_Val));
    TSP_TRY(download(d->coo_compressed_Idx, (size_t)d->coosize, &m->coo_compressed_Idx));
    TSP_TRY(download(d->Blockell_Val, (size_t)d->ellsize, &m->Blockell_Val));
    TSP_TRY(download(d->ell_compressedIdx, ((size_t)d->ellsize + 1) / 2, &m->ell_compressedIdx));
    TSP_TRY(download(d->Blockhyb_Val, (size_t)d->hybellsize + d->hybcoosize, &m->Blockhyb_Val));
    TSP_TRY(download(d->hybIdx, ((size_t)d->hybellsize + 1) / 2 + d->hybcoosize, &m->hybIdx));
    TSP_TRY(download(d->Blockdense_Val, (size_t)d->dnssize, &m->Blockdense_Val));
    TSP_TRY(download(d->Blockdenserow_Val, (size_t)d->dnsrowsize, &m->Blockdenserow_Val));
    TSP_TRY(download(d->denserowid, (size_t)d->ndenserowid, &m->denserowid));
    TSP_TRY(download(d->Blockdensecol_Val, (size_t)d->dnscolsize, &m->Blockdensecol_Val));
    TSP_TRY(download(d->densecolid, (size_t)d->ndensecolid, &m->densecolid));
    TSP_TRY(download(d->deferredcoo_val, (size_t)d->coototal, &m->deferredcoo_val));
    TSP_TRY(download(d->deferredcoo_colidx, (size_t)d->coototal, &m->deferredcoo_colidx));
    TSP_TRY(download(d->deferredcoo_ptr, (size_t)d->rowA + 1, &m->deferredcoo_ptr));
    return TILESPMV_OK;
}

static int upload(DevBuf &b, const void *src, size_t bytes)
{
    TSP_TRY(b.alloc(bytes, bytes == 0));
    if (bytes)
    {
        if (!src)
        {
            set_error("upload: null array in Tile_matrix");
            return TILESPMV_ERR_INVALID;
        }
        TSP_CUDA(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    }
    return TILESPMV_OK;
}

template <class TM, class T>
static int dmat_upload(const TM *m, int rowA, int colA, tilespmv_dmat **out)
{
    if (!m || !out || m->tilenum < 0)
    {
        set_error("upload: invalid Tile_matrix");
        return TILESPMV_ERR_INVALID;
    }
    TSP_TRY(require_device());
    if (m->tilem != (rowA + TS - 1) / TS || m->tilen != (colA + TS - 1) / TS)
    {
        set_error("upload: rowA/colA do not match tilem/tilen of the Tile_matrix");
        return TILESPMV_ERR_INVALID;
    }
    tilespmv_dmat *d = new (std::nothrow) tilespmv_dmat();
    if (!d)
        return TILESPMV_ERR_ALLOC;
    const size_t T_ = (size_t)m->tilenum, vs = sizeof(T);
    d->precision = (int)vs;
    d->rowA = rowA;
    d->colA = colA;
    d->tilem = m->tilem;
    d->tilen = m->tilen;
    d->tilenum = m->tilenum;
    d->csrsize = m->csrsize;
    d->csrptrlen = m->csrptrlen;
    d->coosize = m->coosize;
    d->ellsize = m->ellsize;
    d->hybsize = m->hybsize;
    d->hybellsize = m->hybellsize;
    d->hybcoosize = m->hybcoosize;
    d->dnssize = m->dnssize;
    d->dnsrowsize = m->dnsrowsize;
    d->dnscolsize = m->dnscolsize;
    d->coototal = m->coototal;
    d->nnz = T_ ? m->tile_nnz[T_] : 0;
    d->ndenserowid = T_ ? m->dnsrowptr[T_] : 0;
    d->ndensecolid = T_ ? m->dnscolptr[T_] : 0;
    for (size_t t = 0; t < T_; t++)
        if (m->Format[t] >= 0 && m->Format[t] < 7)
            d->fmt_hist[(int)m->Format[t]]++;
    int rc = TILESPMV_OK;
#define UP(field, count, type)                                                              \
    if (rc == TILESPMV_OK)                                                                  \
    rc = upload(d->field, m->field, (size_t)(count) * sizeof(type))
    UP(tile_ptr, d->tilem + 1, int);
    UP(tile_columnidx, T_, int);
    UP(tile_nnz, T_ + 1, int);
    UP(Format, T_, char);
    UP(blknnz, T_ + 1, int);
    UP(blknnznnz, T_ + 1, unsigned char);
    UP(dnsrowptr, T_ + 1, int);
    UP(dnscolptr, T_ + 1, int);
    UP(tilewidth, T_, char);
    UP(csr_offset, T_ + 1, int);
    UP(csrptr_offset, T_ + 1, int);
    UP(coo_offset, T_ + 1, int);
    UP(ell_offset, T_ + 1, int);
    UP(hyb_offset, T_ + 1, int);
    UP(hyb_coocount, T_ + 1, int);
    UP(dns_offset, T_ + 1, int);
    UP(dnsrow_offset, T_ + 1, int);
    UP(dnscol_offset, T_ + 1, int);
    UP(new_coocount, T_ + 1, int);
    UP(Blockcsr_Val, m->csrsize, T);
    UP(Blockcsr_Ptr, m->csrptrlen, unsigned char);
    UP(csr_compressedIdx, (m->csrsize + 1) / 2, unsigned char);
    UP(Blockcoo_Val, m->coosize, T);
    UP(coo_compressed_Idx, m->coosize, unsigned char);
    UP(Blockell_Val, m->ellsize, T);
    UP(ell_compressedIdx, (m->ellsize + 1) / 2, unsigned char);
    UP(Blockhyb_Val, m->hybellsize + m->hybcoosize, T);
    // hybIdx is indexed tile by tile with per-tile rounding (convert.cu, plan.cu TC_HYB_IDXBYTES), which can be up to
    // one byte per HYB tile longer than the reference's length ceil(hybellsize/2) + hybcoosize (csr2tile.h:840-841;
    // the reference overruns its own array there): allocate that slack zero-filled so the packer never reads past
    // the buffer of an exported / reference-built struct
    if (rc == TILESPMV_OK)
    {
        const size_t ref_len = ((size_t)m->hybellsize + 1) / 2 + (size_t)m->hybcoosize;
        rc = d->hybIdx.alloc(ref_len + (size_t)d->fmt_hist[TILESPMV_FMT_HYB] + 16, true);
        if (rc == TILESPMV_OK && ref_len)
        {
            if (!m->hybIdx)
            {
                set_error("upload: null array in Tile_matrix");
                rc = TILESPMV_ERR_INVALID;
            }
            else if (cudaMemcpy(d->hybIdx.p, m->hybIdx, ref_len, cudaMemcpyHostToDevice) != cudaSuccess)
            {
                set_error("upload: H2D of hybIdx failed");
                rc = TILESPMV_ERR_CUDA;
            }
        }
    }
    UP(Blockdense_Val, m->dnssize, T);
    UP(Blockdenserow_Val, m->dnsrowsize, T);
    UP(denserowid, d->ndenserowid, char);
    UP(Blockdensecol_Val, m->dnscolsize, T);
    UP(densecolid, d->ndensecolid, char);
    UP(deferredcoo_val, m->coototal, T);
    UP(deferredcoo_colidx, m->coototal, int);
    UP(deferredcoo_ptr, rowA + 1, int);
#undef UP
    if (rc != TILESPMV_OK)
    {
        delete d;
        return rc;
    }
    *out = d;
    return TILESPMV_OK;
}

template <class TM>
static void tile_destroy(TM *m)
{
    if (!m)
        return;
    void *all[] = {m->tile_ptr, m->tile_columnidx, m->tile_nnz, m->Format, m->blknnz, m->blknnznnz, m->dnsrowptr,
                   m->dnscolptr, m->tilewidth, m->csr_offset, m->csrptr_offset, m->coo_offset, m->ell_offset,
                   m->hyb_offset, m->hyb_coocount, m->dns_offset, m->dnsrow_offset, m->dnscol_offset, m->new_coocount,
                   m->Blockcsr_Val, m->Blockcsr_Ptr, m->csr_compressedIdx, m->Blockcoo_Val, m->coo_compressed_Idx,
                   m->Blockell_Val, m->ell_compressedIdx, m->Blockhyb_Val, m->hybIdx, m->Blockdense_Val,
                   m->Blockdenserow_Val, m->denserowid, m->Blockdensecol_Val, m->densecolid, m->deferredcoo_val,
                   m->deferredcoo_colidx, m->deferredcoo_ptr};
    for (void *p : all)
        free(p);
    memset(m, 0, sizeof(TM));
}

// ---------------------------------------------------------------------------------------------
// conversion entry
// ---------------------------------------------------------------------------------------------
static int convert_entry(int precision, int rowA, int colA, const int *rowptr, const int *colidx, const void *val,
                         unsigned flags, tilespmv_dmat **out)
{
    if (!out || (precision != TILESPMV_F64 && precision != TILESPMV_F32) || rowA < 0 || colA < 0 || (rowA > 0 && !rowptr))
    {
        set_error("convert: invalid argument");
        return TILESPMV_ERR_INVALID;
    }
    TSP_TRY(require_device());
    const size_t vs = (size_t)precision;
    const int *d_rowptr = rowptr, *d_colidx = colidx;
    const void *d_val = val;
    DevBuf b_rowptr, b_colidx, b_val;
    if (!(flags & TILESPMV_CSR_ON_DEVICE))
    {
        const size_t nnz = rowA > 0 ? (size_t)rowptr[rowA] : 0;
        TSP_TRY(b_rowptr.alloc((size_t)(rowA + 1) * 4, false));
        TSP_TRY(b_colidx.alloc(nnz * 4, false));
        TSP_TRY(b_val.alloc(nnz * vs, false));
        if (rowA > 0)
            TSP_CUDA(cudaMemcpy(b_rowptr.p, rowptr, (size_t)(rowA + 1) * 4, cudaMemcpyHostToDevice));
        else
            TSP_CUDA(cudaMemset(b_rowptr.p, 0, 4));
        if (nnz)
        {
            TSP_CUDA(cudaMemcpy(b_colidx.p, colidx, nnz * 4, cudaMemcpyHostToDevice));
            TSP_CUDA(cudaMemcpy(b_val.p, val, nnz * vs, cudaMemcpyHostToDevice));
        }
        d_rowptr = b_rowptr.as<int>();
        d_colidx = b_colidx.as<int>();
        d_val = b_val.p;
    }
    tilespmv_dmat *d = new (std::nothrow) tilespmv_dmat();
    if (!d)
        return TILESPMV_ERR_ALLOC;
    int rc = precision == TILESPMV_F64
                 ? convert_csr_to_tiles<double>(rowA, colA, d_rowptr, d_colidx, static_cast<const double *>(d_val), d, 0,
                                                (flags & TILESPMV_ENABLE_HYB) != 0)
                 : convert_csr_to_tiles<float>(rowA, colA, d_rowptr, d_colidx, static_cast<const float *>(d_val), d, 0,
                                               (flags & TILESPMV_ENABLE_HYB) != 0);
    if (rc != TILESPMV_OK)
    {
        delete d;
        return rc;
    }
    *out = d;
    return TILESPMV_OK;
}

template <class TM, class T>
static void tile_create_entry(TM *matrix, int rowA, int colA, int *rowptr, int *colidx, T *val)
{
    if (!matrix)
        return;
    zero_tile_matrix(matrix);
    matrix->tilenum = -1;
    tilespmv_dmat *d = nullptr;
    int rc = convert_entry((int)sizeof(T), rowA, colA, rowptr, colidx, val, 0, &d);
    if (rc != TILESPMV_OK)
    {
        fprintf(stderr, "Tile_create failed: %s\n", last_error());
        return;
    }
    rc = dmat_export<TM, T>(d, matrix);
    delete d;
    if (rc != TILESPMV_OK)
    {
        fprintf(stderr, "Tile_create failed: %s\n", last_error());
        tile_destroy(matrix);
        matrix->tilenum = -1;
        return;
    }
    printf("\n  The number of tile = %i\n", matrix->tilenum); // csr2tile.h:661
}

// ---------------------------------------------------------------------------------------------
// bookkeeping half of tilespmv_cpu: ptroffset1/2 + warp-chunk schedule
// ---------------------------------------------------------------------------------------------
template <class TM>
static int prepare_entry(const TM *m, int *ptroffset1, int *ptroffset2, int *rowblkblock, unsigned int **rowidx,
                         int **colstart, int **colstop, int rowA)
{
    if (!m || m->tilenum < 0 || !rowblkblock || !rowidx || !colstart || !colstop)
    {
        set_error("prepare: invalid argument");
        return TILESPMV_ERR_INVALID;
    }
    const int th = TILESPMV_PREFETCH_SMEM_TH;
    int total = 0;
    for (int b = 0; b < m->tilem; b++)
    {
        int n = m->tile_ptr[b + 1] - m->tile_ptr[b];
        total += n <= th ? 1 : (n + th - 1) / th;
    }
    unsigned int *ri = static_cast<unsigned int *>(calloc(total ? total : 1, sizeof(unsigned int)));
    int *cs = static_cast<int *>(calloc(total ? total : 1, sizeof(int)));
    int *ce = static_cast<int *>(calloc(total ? total : 1, sizeof(int)));
    if (!ri || !cs || !ce)
    {
        free(ri);
        free(cs);
        free(ce);
        set_error("prepare: out of memory");
        return TILESPMV_ERR_ALLOC;
    }
    int w = 0;
    for (int b = 0; b < m->tilem; b++)
    {
        const int n = m->tile_ptr[b + 1] - m->tile_ptr[b];
        if (n <= th)
        {
            ri[w++] = (unsigned int)b;
            continue;
        }
        const int k = (n + th - 1) / th, len = (n + k - 1) / k;
        for (int c = 0; c < k; c++, w++)
        {
            ri[w] = (unsigned int)b | 0x80000000u;
            cs[w] = m->tile_ptr[b] + c * len;
            ce[w] = c == k - 1 ? m->tile_ptr[b] + n : m->tile_ptr[b] + (c + 1) * len;
        }
    }
    *rowblkblock = total;
    *rowidx = ri;
    *colstart = cs;
    *colstop = ce;
    // ptroffset1 = the tile's own format prefix; ptroffset2 = csrptr_offset for CSR tiles.  HYB's
    // index-byte offset (tilespmv_cpu.h:196) is a running sum over HYB tiles only.
    if (ptroffset1 && ptroffset2)
    {
        int hybidx = 0;
        for (int b = 0; b < m->tilem; b++)
        {
            const int rowlen = b == m->tilem - 1 ? rowA - (m->tilem - 1) * TS : TS;
            for (int t = m->tile_ptr[b]; t < m->tile_ptr[b + 1]; t++)
            {
                switch (m->Format[t])
                {
                case TILESPMV_FMT_CSR:
                    ptroffset1[t] = m->csr_offset[t];
                    ptroffset2[t] = m->csrptr_offset[t];
                    break;
                case TILESPMV_FMT_COO:
                    ptroffset1[t] = m->coo_offset[t];
                    break;
                case TILESPMV_FMT_ELL:
                    ptroffset1[t] = m->ell_offset[t];
                    break;
                case TILESPMV_FMT_HYB:
                {
                    ptroffset1[t] = m->hyb_offset[t];
                    ptroffset2[t] = hybidx;
                    const int slots = m->blknnz[t + 1] - m->blknnz[t];
                    const int ell = (int)m->tilewidth[t] * rowlen;
                    hybidx += (ell + 1) / 2 + (slots - ell);
                    break;
                }
                case TILESPMV_FMT_DENSE:
                    ptroffset1[t] = m->dns_offset[t];
                    break;
                case TILESPMV_FMT_DENSEROW:
                    ptroffset1[t] = m->dnsrow_offset[t];
                    break;
                case TILESPMV_FMT_DENSECOL:
                    ptroffset1[t] = m->dnscol_offset[t];
                    break;
                default:
                    break;
                }
            }
        }
    }
    return TILESPMV_OK;
}

// ---------------------------------------------------------------------------------------------
// plan helpers
// ---------------------------------------------------------------------------------------------
static int plan_time_impl(tilespmv_plan *P, const void *d_x, void *d_y, int warmup, int iters, cudaStream_t s, double *ms)
{
    if (iters < 1)
        iters = 1;
    for (int i = 0; i < warmup; i++)
        TSP_TRY(plan_launch(P, d_x, d_y, s));
    cudaEvent_t e0, e1;
    TSP_CUDA(cudaEventCreate(&e0));
    TSP_CUDA(cudaEventCreate(&e1));
    TSP_CUDA(cudaStreamSynchronize(s));
    TSP_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; i++)
        TSP_TRY(plan_launch(P, d_x, d_y, s));
    TSP_CUDA(cudaEventRecord(e1, s));
    TSP_CUDA(cudaEventSynchronize(e1));
    float t = 0;
    TSP_CUDA(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms = (double)t / iters;
    return TILESPMV_OK;
}

__global__ void __launch_bounds__(256)
    nnz_by_format_kernel(int T, const char *__restrict__ fmt, const int *__restrict__ tile_nnz, unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long acc[7];
    if (threadIdx.x < 7)
        acc[threadIdx.x] = 0;
    __syncthreads();
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x)
    {
        const int f = fmt[t];
        if (f >= 0 && f < 7)
            atomicAdd(&acc[f], (unsigned long long)(tile_nnz[t + 1] - tile_nnz[t]));
    }
    __syncthreads();
    if (threadIdx.x < 7 && acc[threadIdx.x])
        atomicAdd(&out[threadIdx.x], acc[threadIdx.x]);
}

static int format_profile_impl(const tilespmv_dmat *dm, const tilespmv_plan_options *opts, const void *d_x, void *d_y, int warmup,
                               int iters, double ms[9], int64_t nnz[9])
{
    DevBuf d_n;
    TSP_TRY(d_n.alloc(7 * sizeof(unsigned long long), true));
    if (dm->tilenum > 0)
        TSP_LAUNCH(nnz_by_format_kernel, std::min(grid_for((size_t)dm->tilenum, 256), 1024u), 256, 0, 0, dm->tilenum, dm->Format.as<char>(),
                   dm->tile_nnz.as<int>(), d_n.as<unsigned long long>());
    unsigned long long h[7];
    TSP_CUDA(cudaMemcpy(h, d_n.p, sizeof(h), cudaMemcpyDeviceToHost));
    // the extracted side entries are the COO tiles' nonzeros plus what HYB tiles spill (csr2tile.h:316, :538-545)
    const long long spill = (long long)dm->coototal - (long long)h[TILESPMV_FMT_COO];
    if (nnz)
    {
        for (int f = 0; f < 7; f++)
            nnz[f] = (int64_t)h[f];
        nnz[TILESPMV_FMT_COO] = dm->coototal;
        nnz[TILESPMV_FMT_HYB] -= spill;
        nnz[7] = 0;
        nnz[8] = dm->nnz;
    }
    for (int k = 0; k < 9; k++)
    {
        ms[k] = 0.0;
        const bool present = k >= 7 || dm->fmt_hist[k] > 0 || (k == TILESPMV_FMT_COO && dm->coototal > 0);
        if (!present)
            continue;
        tilespmv_plan_options o;
        memset(&o, 0, sizeof(o));
        if (opts)
            o = *opts;
        // k = 7: a mask with no format bit (bit 7 only keeps it non-zero, i.e. "restricted"); k = 8: everything
        o.format_mask = k < 7 ? (1 << k) : (k == 7 ? 0x80 : 0);
        tilespmv_plan *P = new (std::nothrow) tilespmv_plan();
        if (!P)
            return TILESPMV_ERR_ALLOC;
        int rc = plan_build(dm, &o, P, 0);
        if (rc == TILESPMV_OK)
            rc = plan_time_impl(P, d_x, d_y, warmup, iters, 0, &ms[k]);
        delete P;
        TSP_TRY(rc);
    }
    return TILESPMV_OK;
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    if (!v || !*v)
        return dflt;
    int x = atoi(v);
    return x > 0 ? x : dflt;
}

template <class TM, class T>
static void call_entry(char *filename, TM *matrix, int rowA, int colA, int nnzA, T *x, T *y)
{
    if (!matrix || !x || !y || matrix->tilenum < 0)
    {
        set_error("call_tilespmv_cuda: invalid argument");
        fprintf(stderr, "call_tilespmv_cuda failed: %s\n", last_error());
        return;
    }
    tilespmv_dmat *d = nullptr;
    tilespmv_plan *P = nullptr;
    auto fail = [&]() {
        // whatever failed underneath (upload, plan, allocation), the caller sees it under this entry's name
        const std::string why = last_error();
        if (why.rfind("call_tilespmv_cuda", 0) != 0)
            set_error("call_tilespmv_cuda: %s", why.c_str());
        fprintf(stderr, "call_tilespmv_cuda failed: %s\n", last_error());
        if (P)
            tilespmv_plan_destroy(P);
        if (d)
            tilespmv_dmat_destroy(d);
    };
    if (dmat_upload<TM, T>(matrix, rowA, colA, &d) != TILESPMV_OK)
        return fail();
    if (tilespmv_plan_create(d, nullptr, &P) != TILESPMV_OK)
        return fail();
    DevBuf dx, dy;
    if (dx.alloc((size_t)colA * sizeof(T), false) != TILESPMV_OK || dy.alloc((size_t)rowA * sizeof(T), true) != TILESPMV_OK)
        return fail();
    if (colA && cudaMemcpy(dx.p, x, (size_t)colA * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
    {
        set_error("call_tilespmv_cuda: H2D of x failed");
        return fail();
    }
    // reference protocol (tilespmv_cuda.h:1059-1139): WARMUP_NUM warm-ups, BENCH_REPEAT timed
    // SpMVs; here the batch is timed with CUDA events and includes everything an SpMV needs
    // (the reference excludes its per-iteration cudaMemset(d_y), :1116-1118)
    const int warm = env_int("TILESPMV_WARMUP_NUM", TILESPMV_WARMUP_NUM);
    const int reps = env_int("TILESPMV_BENCH_REPEAT", TILESPMV_BENCH_REPEAT);
    double ms = 0;
    if (plan_time_impl(P, dx.p, dy.p, warm, reps, 0, &ms) != TILESPMV_OK)
        return fail();
    const double gflops = ms > 0 ? 2.0 * (double)nnzA * 1e-6 / ms : 0.0;
    printf("  CUDA SpMV runtime %4.2f ms, %4.2f GFlops\n\n", ms, gflops);
    if (FILE *f = fopen("results.csv", "a"))
    {
        // the reference's six columns (tilespmv_cuda.h:1142-1147) stay the default so that existing sweep scripts keep
        // parsing the file; TILESPMV_CSV_EXTENDED=1 appends the roofline columns SURVEY.md 8f-4 asks for:
        // algorithmic bytes (8d), achieved GB/s on them, fraction of the HBM peak (TILESPMV_HBM_PEAK_GBS, default the
        // nominal 8000 GB/s of BASELINE.json)
        if (env_int("TILESPMV_CSV_EXTENDED", 0))
        {
            double peak = 8000.0;
            if (const char *e = getenv("TILESPMV_HBM_PEAK_GBS"))
                if (atof(e) > 0)
                    peak = atof(e);
            const double gbps = ms > 0 ? (double)P->b_alg * 1e-6 / ms : 0.0;
            fprintf(f, "%s,%i,%i,%i,%f,%f,%lld,%f,%f\n", filename ? filename : "", rowA, colA, nnzA, ms, gflops, (long long)P->b_alg, gbps,
                    gbps / peak);
        }
        else
            fprintf(f, "%s,%i,%i,%i,%f,%f\n", filename ? filename : "", rowA, colA, nnzA, ms, gflops);
        fclose(f);
    }
    if (rowA && cudaMemcpy(y, dy.p, (size_t)rowA * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess)
    {
        set_error("call_tilespmv_cuda: D2H of y failed");
        return fail();
    }
    tilespmv_plan_destroy(P);
    tilespmv_dmat_destroy(d);
}

// ---------------------------------------------------------------------------------------------
// Matrix Market front end (semantics of mmio_allinone, mmio_highlevel.h:593-759)
// ---------------------------------------------------------------------------------------------
// binary CSR cache of the Matrix Market front end: 8-byte magic, {sizeof(T), m, n, nnz, symmetric} as int32, then the
// three arrays raw.  A file that does not match in every header field is ignored.
static const char MTX_CACHE_MAGIC[8] = {'T', 'S', 'P', 'C', 'S', 'R', '1', 0};
template <class T>
static bool mtx_cache_load(const char *path, int *m, int *n, int *nnz, int *sym, int **rp, int **cj, T **cv)
{
    FILE *f = fopen(path, "rb");
    if (!f)
        return false;
    char magic[8];
    int h[5];
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, MTX_CACHE_MAGIC, 8) == 0 && fread(h, sizeof(int), 5, f) == 5 &&
              h[0] == (int)sizeof(T) && h[1] >= 0 && h[2] >= 0 && h[3] >= 0;
    int *a = nullptr, *b = nullptr;
    T *c = nullptr;
    if (ok)
    {
        const size_t nz = (size_t)h[3];
        a = static_cast<int *>(malloc(((size_t)h[1] + 1) * sizeof(int)));
        b = static_cast<int *>(malloc((nz ? nz : 1) * sizeof(int)));
        c = static_cast<T *>(malloc((nz ? nz : 1) * sizeof(T)));
        ok = a && b && c && fread(a, sizeof(int), (size_t)h[1] + 1, f) == (size_t)h[1] + 1 && fread(b, sizeof(int), nz, f) == nz &&
             fread(c, sizeof(T), nz, f) == nz && a[h[1]] == h[3];
    }
    fclose(f);
    if (!ok)
    {
        free(a);
        free(b);
        free(c);
        return false;
    }
    *m = h[1];
    *n = h[2];
    *nnz = h[3];
    *sym = h[4];
    *rp = a;
    *cj = b;
    *cv = c;
    return true;
}
template <class T>
static void mtx_cache_store(const char *path, int m, int n, int nnz, int sym, const int *rp, const int *cj, const T *cv)
{
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f)
        return; // the cache is best effort
    const int h[5] = {(int)sizeof(T), m, n, nnz, sym};
    bool ok = fwrite(MTX_CACHE_MAGIC, 1, 8, f) == 8 && fwrite(h, sizeof(int), 5, f) == 5 &&
              fwrite(rp, sizeof(int), (size_t)m + 1, f) == (size_t)m + 1 && fwrite(cj, sizeof(int), (size_t)nnz, f) == (size_t)nnz &&
              fwrite(cv, sizeof(T), (size_t)nnz, f) == (size_t)nnz;
    ok = fclose(f) == 0 && ok;
    if (ok)
        rename(tmp.c_str(), path);
    else
        remove(tmp.c_str());
}

template <class T>
static int mmio_entry(int *m, int *n, int *nnz, int *isSymmetric, int **csrRowPtr, int **csrColIdx, T **csrVal,
                      const char *filename)
{
    FILE *f = fopen(filename, "rb");
    if (!f)
        return -1;
    // slurp the file: one pass, hand-rolled number parsing (the reference's fscanf loop parses
    // ~2.6 M entries/s)
    fseek(f, 0, SEEK_END);
    long fsize = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)fsize + 1);
    size_t got = fread(buf.data(), 1, (size_t)fsize, f);
    fclose(f);
    buf[got] = 0;
    const char *p = buf.data(), *end = buf.data() + got;
    auto next_line = [&](const char *q) {
        while (q < end && *q != '\n')
            q++;
        return q < end ? q + 1 : end;
    };
    // banner: %%MatrixMarket matrix coordinate <field> <symmetry>
    char tok[5][64] = {{0}};
    {
        const char *q = p;
        for (int k = 0; k < 5; k++)
        {
            while (q < end && (*q == ' ' || *q == '\t'))
                q++;
            int len = 0;
            while (q < end && !isspace((unsigned char)*q) && len < 63)
                tok[k][len++] = (char)tolower((unsigned char)*q++);
            tok[k][len] = 0;
        }
    }
    if (strcmp(tok[0], "%%matrixmarket") != 0 || strcmp(tok[1], "matrix") != 0 || tok[4][0] == 0)
    {
        printf("Could not process Matrix Market banner.\n");
        return -2;
    }
    const bool is_pattern = !strcmp(tok[3], "pattern"), is_complex = !strcmp(tok[3], "complex");
    const bool symm = !strcmp(tok[4], "symmetric") || !strcmp(tok[4], "hermitian");
    p = next_line(p);
    while (p < end && *p == '%')
        p = next_line(p);
    long M_ = 0, N_ = 0, NZ = 0;
    for (;;)
    {
        // the size line: three integers on ONE line (mmio.h:600-612 reads line by line until sscanf yields 3 items);
        // blank or malformed lines before it are skipped, and p always ends up right behind the size line itself
        if (p >= end)
            return -4;
        const char *le = next_line(p);
        const char *q = p;
        long v3[3] = {0, 0, 0};
        int got3 = 0;
        for (; got3 < 3; got3++)
        {
            while (q < le && (*q == ' ' || *q == '\t'))
                q++;
            if (q < le && *q == '+')
                q++;
            auto r = std::from_chars(q, le, v3[got3]);
            if (r.ec != std::errc())
                break;
            q = r.ptr;
        }
        p = le;
        if (got3 == 3)
        {
            M_ = v3[0];
            N_ = v3[1];
            NZ = v3[2];
            break;
        }
    }
    if (M_ < 0 || N_ < 0 || NZ < 0 || M_ > 0x7ffffff0l || N_ > 0x7ffffff0l || NZ > 0x7ffffff0l)
        return -4;
    // ---- binary cache of the parsed CSR (opt-in: TILESPMV_MTX_CACHE=<directory>), keyed by file name, size, mtime
    std::string cache_path;
    if (const char *dir = getenv("TILESPMV_MTX_CACHE"))
    {
        struct stat st;
        if (dir[0] && stat(filename, &st) == 0)
        {
            const char *base = strrchr(filename, '/');
            base = base ? base + 1 : filename;
            char key[96];
            snprintf(key, sizeof(key), ".%lld.%lld.f%d.tspcsr", (long long)st.st_size, (long long)st.st_mtime, (int)sizeof(T) * 8);
            cache_path = std::string(dir) + "/" + base + key;
            if (mtx_cache_load<T>(cache_path.c_str(), m, n, nnz, isSymmetric, csrRowPtr, csrColIdx, csrVal))
                return 0;
        }
    }
    std::vector<int> ri((size_t)NZ), ci((size_t)NZ);
    std::vector<T> vv((size_t)NZ);
    std::vector<int> cnt((size_t)M_ + 1, 0);
    long e = 0;
    // ---- parallel parse: the body is cut at line ends into one piece per thread; pass 1 counts the entry lines of
    //      every piece, pass 2 parses them straight into place.  Anything unusual (a malformed or out-of-range
    //      entry, fewer lines than the size line promises) falls back to the serial loop below, which reproduces
    //      the reference's fscanf behaviour token by token.
    bool parsed = false;
    if (NZ >= 4096)
    {
        const int nt = std::max(1, std::min(omp_get_max_threads(), 64));
        std::vector<const char *> cut((size_t)nt + 1);
        cut[0] = p;
        cut[nt] = end;
        for (int t = 1; t < nt; t++)
        {
            const char *q = p + (size_t)(end - p) * (size_t)t / (size_t)nt;
            cut[t] = std::max(cut[t - 1], next_line(q));
        }
        std::vector<long> lines((size_t)nt + 1, 0);
        int bad = 0;
#pragma omp parallel num_threads(nt)
        {
            const int t = omp_get_thread_num();
            long c = 0;
            for (const char *q = cut[t]; q < cut[t + 1];)
            {
                const char *l = q;
                q = next_line(q);
                while (l < q && isspace((unsigned char)*l))
                    l++;
                c += l < q ? 1 : 0;
            }
            lines[(size_t)t + 1] = c;
#pragma omp barrier
#pragma omp single
            for (int k = 0; k < nt; k++)
                lines[(size_t)k + 1] += lines[(size_t)k];
            long idx = lines[(size_t)t];
            int mybad = 0;
            for (const char *q = cut[t]; q < cut[t + 1] && idx < NZ && !mybad;)
            {
                const char *l = q;
                q = next_line(q);
                const char *le = q; // one past the line (incl. its newline)
                while (l < le && isspace((unsigned char)*l))
                    l++;
                if (l >= le)
                    continue;
                long a = 0, b = 0;
                double re = 1.0;
                auto rint = [&](long &out) {
                    while (l < le && (*l == ' ' || *l == '\t'))
                        l++;
                    if (l < le && *l == '+')
                        l++;
                    auto r = std::from_chars(l, le, out);
                    if (r.ec != std::errc())
                        return false;
                    l = r.ptr;
                    return true;
                };
                auto rdbl = [&](double &out) {
                    while (l < le && (*l == ' ' || *l == '\t'))
                        l++;
                    if (l < le && *l == '+')
                        l++;
                    auto r = std::from_chars(l, le, out);
                    if (r.ec != std::errc())
                        return false;
                    l = r.ptr;
                    return true;
                };
                bool ok = rint(a) && rint(b);
                if (ok && !is_pattern)
                {
                    ok = rdbl(re);
                    double im;
                    if (ok && is_complex)
                        ok = rdbl(im);
                }
                while (ok && l < le && isspace((unsigned char)*l))
                    l++;
                if (!ok || l < le || a < 1 || a > M_ || b < 1 || b > N_)
                {
                    mybad = 1;
                    break;
                }
                ri[(size_t)idx] = (int)a - 1;
                ci[(size_t)idx] = (int)b - 1;
                vv[(size_t)idx] = (T)re;
                idx++;
            }
            if (mybad)
            {
#pragma omp atomic write
                bad = 1;
            }
        }
        if (!bad && lines[(size_t)nt] >= NZ)
        {
            parsed = true;
            e = NZ;
            for (long k = 0; k < NZ; k++)
            {
                cnt[ri[(size_t)k]]++;
                if (symm && ri[(size_t)k] != ci[(size_t)k])
                    cnt[ci[(size_t)k]]++;
            }
        }
    }
    for (; !parsed && e < NZ; e++)
    {
        char *q;
        long a = strtol(p, &q, 10);
        if (q == p)
            break;
        p = q;
        long b = strtol(p, &q, 10);
        if (q == p)
            break;
        p = q;
        double re = 1.0;
        if (!is_pattern)
        {
            re = strtod(p, &q);
            if (q == p)
                break;
            p = q;
            if (is_complex)
            {
                strtod(p, &q);
                p = q;
            }
        }
        if (a < 1 || a > M_ || b < 1 || b > N_)
            break;
        ri[e] = (int)a - 1;
        ci[e] = (int)b - 1;
        vv[e] = (T)re;
        cnt[ri[e]]++;
        if (symm && ri[e] != ci[e])
            cnt[ci[e]]++;
    }
    NZ = e;
    long run = 0;
    for (long i = 0; i <= M_; i++)
    {
        long c = cnt[i];
        cnt[i] = (int)run;
        run += c;
    }
    if (run > 0x7ffffff0l)
        return -4;
    int *rp = static_cast<int *>(malloc(((size_t)M_ + 1) * sizeof(int)));
    int *cj = static_cast<int *>(malloc((run ? (size_t)run : 1) * sizeof(int)));
    T *cv = static_cast<T *>(malloc((run ? (size_t)run : 1) * sizeof(T)));
    if (!rp || !cj || !cv)
    {
        free(rp);
        free(cj);
        free(cv);
        return -4;
    }
    memcpy(rp, cnt.data(), ((size_t)M_ + 1) * sizeof(int));
    for (long k = 0; k < NZ; k++) // file order inside every row; mirrored entry right after its source
    {
        int d = cnt[ri[k]]++;
        cj[d] = ci[k];
        cv[d] = vv[k];
        if (symm && ri[k] != ci[k])
        {
            d = cnt[ci[k]]++;
            cj[d] = ri[k];
            cv[d] = vv[k];
        }
    }
    *m = (int)M_;
    *n = (int)N_;
    *nnz = (int)run;
    *isSymmetric = symm ? 1 : 0;
    *csrRowPtr = rp;
    *csrColIdx = cj;
    *csrVal = cv;
    if (!cache_path.empty())
        mtx_cache_store<T>(cache_path.c_str(), *m, *n, *nnz, *isSymmetric, rp, cj, cv);
    return 0;
}

} // namespace tsp

// =============================================================================================
// extern "C"
// =============================================================================================
using namespace tsp;

extern "C"
{

const char *tilespmv_last_error(void) { return last_error(); }
const char *tilespmv_version(void) { return "tilespmv_b200 0.1.0 (sm_100a)"; }
int64_t tilespmv_kernel_launch_count(void) { return g_launches.load(); }

void Tile_create_f64(Tile_matrix_f64 *matrix, int rowA, int colA, int nnzA, int *csrRowPtrA, int *csrColIdxA, double *csrValA)
try
{
    clear_error();
    (void)nnzA;
    tile_create_entry<Tile_matrix_f64, double>(matrix, rowA, colA, csrRowPtrA, csrColIdxA, csrValA);
}
TSP_CATCH_VOID("Tile_create_f64")
void Tile_create_f32(Tile_matrix_f32 *matrix, int rowA, int colA, int nnzA, int *csrRowPtrA, int *csrColIdxA, float *csrValA)
try
{
    clear_error();
    (void)nnzA;
    tile_create_entry<Tile_matrix_f32, float>(matrix, rowA, colA, csrRowPtrA, csrColIdxA, csrValA);
}
TSP_CATCH_VOID("Tile_create_f32")
void Tile_destroy_f64(Tile_matrix_f64 *matrix) { tile_destroy(matrix); }
void Tile_destroy_f32(Tile_matrix_f32 *matrix) { tile_destroy(matrix); }

int tilespmv_prepare_f64(const Tile_matrix_f64 *matrix, int *ptroffset1, int *ptroffset2, int *rowblkblock,
                         unsigned int **blkcoostylerowidx, int **blkcoostylerowidx_colstart,
                         int **blkcoostylerowidx_colstop, int rowA)
try
{
    clear_error();
    return prepare_entry(matrix, ptroffset1, ptroffset2, rowblkblock, blkcoostylerowidx, blkcoostylerowidx_colstart,
                         blkcoostylerowidx_colstop, rowA);
}
TSP_CATCH_INT("tilespmv_prepare_f64")
int tilespmv_prepare_f32(const Tile_matrix_f32 *matrix, int *ptroffset1, int *ptroffset2, int *rowblkblock,
                         unsigned int **blkcoostylerowidx, int **blkcoostylerowidx_colstart,
                         int **blkcoostylerowidx_colstop, int rowA)
try
{
    clear_error();
    return prepare_entry(matrix, ptroffset1, ptroffset2, rowblkblock, blkcoostylerowidx, blkcoostylerowidx_colstart,
                         blkcoostylerowidx_colstop, rowA);
}
TSP_CATCH_INT("tilespmv_prepare_f32")

void call_tilespmv_cuda_f64(char *filename, Tile_matrix_f64 *matrix, int *, int *, int, unsigned int *, int *, int *,
                            int rowA, int colA, int nnzA, int *, int *, double *, double, double *x, double *y, double *)
try
{
    clear_error();
    call_entry<Tile_matrix_f64, double>(filename, matrix, rowA, colA, nnzA, x, y);
}
TSP_CATCH_VOID("call_tilespmv_cuda_f64")
void call_tilespmv_cuda_f32(char *filename, Tile_matrix_f32 *matrix, int *, int *, int, unsigned int *, int *, int *,
                            int rowA, int colA, int nnzA, int *, int *, float *, float, float *x, float *y, float *)
try
{
    clear_error();
    call_entry<Tile_matrix_f32, float>(filename, matrix, rowA, colA, nnzA, x, y);
}
TSP_CATCH_VOID("call_tilespmv_cuda_f32")

int tilespmv_convert(int precision, int rowA, int colA, const int *rowptr, const int *colidx, const void *val,
                     unsigned flags, tilespmv_dmat **out)
try
{
    clear_error();
    return convert_entry(precision, rowA, colA, rowptr, colidx, val, flags, out);
}
TSP_CATCH_INT("tilespmv_convert")
int tilespmv_dmat_upload_f64(const Tile_matrix_f64 *matrix, int rowA, int colA, tilespmv_dmat **out)
try
{
    clear_error();
    return dmat_upload<Tile_matrix_f64, double>(matrix, rowA, colA, out);
}
TSP_CATCH_INT("tilespmv_dmat_upload_f64")
int tilespmv_dmat_upload_f32(const Tile_matrix_f32 *matrix, int rowA, int colA, tilespmv_dmat **out)
try
{
    clear_error();
    return dmat_upload<Tile_matrix_f32, float>(matrix, rowA, colA, out);
}
TSP_CATCH_INT("tilespmv_dmat_upload_f32")
int tilespmv_dmat_export_f64(const tilespmv_dmat *dm, Tile_matrix_f64 *matrix)
try
{
    clear_error();
    return dmat_export<Tile_matrix_f64, double>(dm, matrix);
}
TSP_CATCH_INT("tilespmv_dmat_export_f64")
int tilespmv_dmat_export_f32(const tilespmv_dmat *dm, Tile_matrix_f32 *matrix)
try
{
    clear_error();
    return dmat_export<Tile_matrix_f32, float>(dm, matrix);
}
TSP_CATCH_INT("tilespmv_dmat_export_f32")
void tilespmv_dmat_destroy(tilespmv_dmat *dm) { delete dm; }

int tilespmv_dmat_get_info(const tilespmv_dmat *dm, tilespmv_dmat_info *info)
try
{
    clear_error();
    if (!dm || !info)
    {
        set_error("dmat_get_info: null argument");
        return TILESPMV_ERR_INVALID;
    }
    memset(info, 0, sizeof(*info));
    info->precision = dm->precision;
    info->rowA = dm->rowA;
    info->colA = dm->colA;
    info->tilem = dm->tilem;
    info->tilen = dm->tilen;
    info->tilenum = dm->tilenum;
    info->nnz = dm->nnz;
    info->nnz_side = dm->coototal;
    for (int f = 0; f < 7; f++)
        info->tiles_by_format[f] = dm->fmt_hist[f];
    // stored slots per format = the per-format sizes of the reference's size pass (csr2tile.h:754-793): with
    // tiles_by_format and nnz this is the padding each format carries (cf. DEBUG_FORMATCOST, tilespmv_cuda.h:102-111)
    const int slots[7] = {dm->csrsize, dm->coosize, dm->ellsize, dm->hybsize, dm->dnssize, dm->dnsrowsize, dm->dnscolsize};
    for (int f = 0; f < 7; f++)
        info->slots_by_format[f] = slots[f];
    info->device_bytes = dm->device_bytes();
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_dmat_get_info")

int tilespmv_plan_create(const tilespmv_dmat *dm, const tilespmv_plan_options *opts, tilespmv_plan **out)
try
{
    clear_error();
    if (!dm || !out)
    {
        set_error("plan_create: null argument");
        return TILESPMV_ERR_INVALID;
    }
    TSP_TRY(require_device());
    tilespmv_plan *P = new (std::nothrow) tilespmv_plan();
    if (!P)
        return TILESPMV_ERR_ALLOC;
    int rc = plan_build(dm, opts, P, 0);
    if (rc != TILESPMV_OK)
    {
        delete P;
        return rc;
    }
    *out = P;
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_plan_create")
int tilespmv_plan_save(const tilespmv_plan *plan, const char *path)
try
{
    clear_error();
    if (!plan || !path)
    {
        set_error("plan_save: null argument");
        return TILESPMV_ERR_INVALID;
    }
    return plan_save(plan, path);
}
TSP_CATCH_INT("tilespmv_plan_save")
int tilespmv_plan_load(const char *path, tilespmv_plan **out)
try
{
    clear_error();
    if (!path || !out)
    {
        set_error("plan_load: null argument");
        return TILESPMV_ERR_INVALID;
    }
    TSP_TRY(require_device());
    return plan_load(path, out);
}
TSP_CATCH_INT("tilespmv_plan_load")
void tilespmv_plan_destroy(tilespmv_plan *plan) { delete plan; }

int tilespmv_plan_spmv(tilespmv_plan *plan, const void *d_x, void *d_y, void *stream)
try
{
    clear_error();
    if (!plan || (!d_x && plan->colA) || (!d_y && plan->rowA))
    {
        set_error("plan_spmv: null argument");
        return TILESPMV_ERR_INVALID;
    }
    return plan_launch(plan, d_x, d_y, static_cast<cudaStream_t>(stream));
}
TSP_CATCH_INT("tilespmv_plan_spmv")

int tilespmv_plan_spmv_host(tilespmv_plan *plan, const void *x, void *y)
try
{
    clear_error();
    if (!plan || (!x && plan->colA) || (!y && plan->rowA))
    {
        set_error("plan_spmv_host: null argument");
        return TILESPMV_ERR_INVALID;
    }
    const size_t vs = (size_t)plan->precision;
    if (!plan->hx.p)
        TSP_TRY(plan->hx.alloc((size_t)plan->colA * vs, false));
    if (!plan->hy.p)
        TSP_TRY(plan->hy.alloc((size_t)plan->rowA * vs, true));
    if (plan->colA)
        TSP_CUDA(cudaMemcpyAsync(plan->hx.p, x, (size_t)plan->colA * vs, cudaMemcpyHostToDevice, 0));
    TSP_TRY(plan_launch(plan, plan->hx.p, plan->hy.p, 0));
    if (plan->rowA)
        TSP_CUDA(cudaMemcpyAsync(y, plan->hy.p, (size_t)plan->rowA * vs, cudaMemcpyDeviceToHost, 0));
    TSP_CUDA(cudaStreamSynchronize(0));
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_plan_spmv_host")

int tilespmv_plan_spmv_host_batch(tilespmv_plan *plan, int nvec, const void *const *x, void *const *y)
try
{
    clear_error();
    if (!plan || nvec < 0 || (nvec > 0 && (!x || !y)))
    {
        set_error("plan_spmv_host_batch: invalid argument");
        return TILESPMV_ERR_INVALID;
    }
    constexpr int R = tilespmv_plan::HOST_RING;
    const size_t xb = (size_t)plan->colA * (size_t)plan->precision, yb = (size_t)plan->rowA * (size_t)plan->precision;
    if (!plan->ring_ready)
    {
        // streams, events and buffers are created as a unit: a failure half-way tears everything down again so that
        // the next call starts from scratch instead of launching on null buffers
        auto setup = [&]() -> int {
            TSP_CUDA(cudaStreamCreateWithFlags(&plan->s_in, cudaStreamNonBlocking));
            TSP_CUDA(cudaStreamCreateWithFlags(&plan->s_comp, cudaStreamNonBlocking));
            TSP_CUDA(cudaStreamCreateWithFlags(&plan->s_out, cudaStreamNonBlocking));
            for (int i = 0; i < R; i++)
            {
                TSP_CUDA(cudaEventCreateWithFlags(&plan->ev_in[i], cudaEventDisableTiming));
                TSP_CUDA(cudaEventCreateWithFlags(&plan->ev_comp[i], cudaEventDisableTiming));
                TSP_CUDA(cudaEventCreateWithFlags(&plan->ev_out[i], cudaEventDisableTiming));
                TSP_TRY(plan->bx[i].alloc(xb, false));
                TSP_TRY(plan->by[i].alloc(yb, false)); // not zeroed: the plan writes every row (and a memset on the legacy stream would race with s_comp)
            }
            return TILESPMV_OK;
        };
        const int rc = setup();
        if (rc != TILESPMV_OK)
        {
            plan->release_host_ring();
            return rc;
        }
        plan->ring_ready = true;
    }
    // vector i uses ring slot i % R.  H2D stream: x_i -> bx once the kernel that last read bx is done; kernel
    // stream: SpMV once x_i has landed and the D2H that last read by is done; D2H stream: by -> y_i.  The two
    // copy directions run concurrently (PCIe is full duplex) and both overlap the kernels.
    for (int i = 0; i < nvec; i++)
    {
        const int b = i % R;
        if (!x[i] && xb)
        {
            set_error("plan_spmv_host_batch: x[%d] is null", i);
            return TILESPMV_ERR_INVALID;
        }
        if (i >= R)
            TSP_CUDA(cudaStreamWaitEvent(plan->s_in, plan->ev_comp[b], 0));
        if (xb)
            TSP_CUDA(cudaMemcpyAsync(plan->bx[b].p, x[i], xb, cudaMemcpyHostToDevice, plan->s_in));
        TSP_CUDA(cudaEventRecord(plan->ev_in[b], plan->s_in));
        TSP_CUDA(cudaStreamWaitEvent(plan->s_comp, plan->ev_in[b], 0));
        if (i >= R)
            TSP_CUDA(cudaStreamWaitEvent(plan->s_comp, plan->ev_out[b], 0));
        TSP_TRY(plan_launch(plan, plan->bx[b].p, plan->by[b].p, plan->s_comp));
        TSP_CUDA(cudaEventRecord(plan->ev_comp[b], plan->s_comp));
        TSP_CUDA(cudaStreamWaitEvent(plan->s_out, plan->ev_comp[b], 0));
        if (yb && y[i])
            TSP_CUDA(cudaMemcpyAsync(y[i], plan->by[b].p, yb, cudaMemcpyDeviceToHost, plan->s_out));
        TSP_CUDA(cudaEventRecord(plan->ev_out[b], plan->s_out));
    }
    TSP_CUDA(cudaStreamSynchronize(plan->s_in));
    TSP_CUDA(cudaStreamSynchronize(plan->s_comp));
    TSP_CUDA(cudaStreamSynchronize(plan->s_out));
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_plan_spmv_host_batch")

int tilespmv_plan_iterate(tilespmv_plan *plan, void *d_xa, void *d_xb, int niters, void *stream)
try
{
    clear_error();
    if (!plan || niters < 0 || ((!d_xa || !d_xb) && plan->rowA))
    {
        set_error("plan_iterate: invalid argument");
        return TILESPMV_ERR_INVALID;
    }
    if (plan->rowA != plan->colA)
    {
        set_error("plan_iterate: x <- A*x needs a square matrix (%d x %d)", plan->rowA, plan->colA);
        return TILESPMV_ERR_INVALID;
    }
    if (plan->npeers > 0)
    {
        set_error("plan_iterate: with peers set every iteration needs a cross-GPU barrier; drive the loop with tilespmv_plan_spmv");
        return TILESPMV_ERR_UNSUPPORTED;
    }
    if (niters == 0 || plan->rowA == 0)
        return TILESPMV_OK;
    if (!plan->iter_exec || plan->iter_xa != d_xa || plan->iter_xb != d_xb || plan->iter_n != niters)
    {
        if (plan->iter_exec)
        {
            cudaGraphExecDestroy(plan->iter_exec);
            plan->iter_exec = nullptr;
        }
        if (!plan->s_capture)
            TSP_CUDA(cudaStreamCreateWithFlags(&plan->s_capture, cudaStreamNonBlocking));
        const int64_t counted = g_launches.load(); // launches recorded while capturing are not executions
        TSP_CUDA(cudaStreamBeginCapture(plan->s_capture, cudaStreamCaptureModeThreadLocal));
        int rc = TILESPMV_OK;
        for (int i = 0; i < niters && rc == TILESPMV_OK; i++)
            rc = plan_launch(plan, (i & 1) ? d_xb : d_xa, (i & 1) ? d_xa : d_xb, plan->s_capture);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(plan->s_capture, &graph);
        g_launches.store(counted);
        if (rc != TILESPMV_OK || e != cudaSuccess || !graph)
        {
            if (graph)
                cudaGraphDestroy(graph);
            if (rc == TILESPMV_OK)
                set_error("plan_iterate: stream capture failed: %s", cudaGetErrorString(e));
            return rc != TILESPMV_OK ? rc : TILESPMV_ERR_CUDA;
        }
        e = cudaGraphInstantiate(&plan->iter_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess)
        {
            plan->iter_exec = nullptr;
            set_error("plan_iterate: cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
            return TILESPMV_ERR_CUDA;
        }
        plan->iter_xa = d_xa;
        plan->iter_xb = d_xb;
        plan->iter_n = niters;
    }
    TSP_CUDA(cudaGraphLaunch(plan->iter_exec, static_cast<cudaStream_t>(stream)));
    tilespmv_plan_info info;
    tilespmv_plan_get_info(plan, &info);
    g_launches.fetch_add((int64_t)niters * info.launches_per_spmv, std::memory_order_relaxed);
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_plan_iterate")

int tilespmv_partition_rows(int precision, int rowA, const int *rowptr, int nparts, int *row_cuts)
try
{
    clear_error();
    if ((precision != TILESPMV_F64 && precision != TILESPMV_F32) || rowA < 0 || nparts < 1 || !row_cuts || (rowA > 0 && !rowptr))
    {
        set_error("partition_rows: invalid argument");
        return TILESPMV_ERR_INVALID;
    }
    // streamed bytes per block row with the weights of B_alg: value + 4-bit index per nonzero, 16 y
    // values written, one row record; tile headers are not known before conversion and second order
    const int tilem = (rowA + TS - 1) / TS;
    std::vector<double> pre((size_t)tilem + 1, 0.0);
    for (int b = 0; b < tilem; b++)
    {
        const int r0 = b * TS, r1 = std::min(rowA, r0 + TS);
        const double nnz_br = (double)rowptr[r1] - (double)rowptr[r0];
        pre[(size_t)b + 1] = pre[(size_t)b] + nnz_br * ((double)precision + 0.5) + (double)(TS * precision) + 16.0;
    }
    const double total = pre[(size_t)tilem];
    int last = 0;
    row_cuts[0] = 0;
    for (int g = 1; g < nparts; g++)
    {
        // cut where the prefix crosses g / nparts of the total: the nearer of the two neighbouring block-row boundaries
        const double target = total * (double)g / (double)nparts;
        int b = (int)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
        if (b > 0 && std::fabs(pre[(size_t)b - 1] - target) <= std::fabs(pre[(size_t)std::min(b, tilem)] - target))
            b--;
        b = std::min(std::max(b, last), tilem);
        last = b;
        row_cuts[g] = std::min(b * TS, rowA);
    }
    row_cuts[nparts] = rowA;
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_partition_rows")

int tilespmv_plan_set_peers(tilespmv_plan *plan, int npeers, void *const *peer_x, int64_t row_offset)
try
{
    clear_error();
    if (!plan || npeers < 0 || npeers > TSP_MAX_PEERS || (npeers > 0 && !peer_x))
    {
        set_error("plan_set_peers: invalid argument (at most %d peers)", TSP_MAX_PEERS);
        return TILESPMV_ERR_INVALID;
    }
    plan->npeers = npeers;
    plan->row_offset = row_offset;
    for (int p = 0; p < TSP_MAX_PEERS; p++)
    {
        plan->peers[p] = p < npeers ? peer_x[p] : nullptr;
        plan->peer_lo[p] = 0;
        plan->peer_hi[p] = p < npeers ? (int64_t)1 << 62 : 0; // every row
    }
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_plan_set_peers")

int tilespmv_plan_get_info(const tilespmv_plan *plan, tilespmv_plan_info *info)
try
{
    clear_error();
    if (!plan || !info)
    {
        set_error("plan_get_info: null argument");
        return TILESPMV_ERR_INVALID;
    }
    memset(info, 0, sizeof(*info));
    info->precision = plan->precision;
    info->nchunks = plan->nchunks;
    info->stream_bytes = plan->stream_bytes;
    info->algorithmic_bytes = plan->b_alg;
    info->csr_bytes = plan->b_csr;
    info->split_rows = plan->nsplit;
    auto launches = [](const tilespmv_plan *q) {
        return q->nchunks == 0 ? 0 : 1 + (q->nsplit_small > 0 ? 1 : 0) + (q->nsplit > q->nsplit_small ? 1 : 0);
    };
    info->launches_per_spmv = launches(plan);
    for (const tilespmv_plan *q : plan->sub) // column-panel sub-plans: one more launch (+ fix-ups) each
    {
        info->nchunks += q->nchunks;
        info->stream_bytes += q->stream_bytes;
        info->split_rows += q->nsplit;
        info->launches_per_spmv += launches(q);
    }
    info->xpanels = 1 + (int64_t)plan->sub.size();
    info->grid = plan->grid;
    info->block = plan->block;
    info->smem_bytes = plan->smem;
    info->chunk_bytes = plan->chunk_bytes;
    info->xstage_bytes = plan->xstage_bytes;
    info->device_bytes = plan->device_bytes();
    info->csr_groups = plan->csr_groups;
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_plan_get_info")

int tilespmv_format_profile(const tilespmv_dmat *dm, const tilespmv_plan_options *opts, const void *d_x, void *d_y, int warmup, int iters,
                            double ms[9], int64_t nnz[9])
try
{
    clear_error();
    if (!dm || !ms || (!d_x && dm->colA) || (!d_y && dm->rowA))
    {
        set_error("format_profile: null argument");
        return TILESPMV_ERR_INVALID;
    }
    TSP_TRY(require_device());
    return format_profile_impl(dm, opts, d_x, d_y, warmup, iters, ms, nnz);
}
TSP_CATCH_INT("tilespmv_format_profile")

int tilespmv_plan_time(tilespmv_plan *plan, const void *d_x, void *d_y, int warmup, int iters, void *stream,
                       double *ms_per_spmv)
try
{
    clear_error();
    if (!plan || !ms_per_spmv)
    {
        set_error("plan_time: null argument");
        return TILESPMV_ERR_INVALID;
    }
    return plan_time_impl(plan, d_x, d_y, warmup, iters, static_cast<cudaStream_t>(stream), ms_per_spmv);
}
TSP_CATCH_INT("tilespmv_plan_time")

int tilespmv_mmio_allinone_f64(int *m, int *n, int *nnz, int *isSymmetric, int **csrRowPtr, int **csrColIdx,
                               double **csrVal, const char *filename)
try
{
    clear_error();
    return mmio_entry<double>(m, n, nnz, isSymmetric, csrRowPtr, csrColIdx, csrVal, filename);
}
TSP_CATCH_INT("tilespmv_mmio_allinone_f64")
int tilespmv_mmio_allinone_f32(int *m, int *n, int *nnz, int *isSymmetric, int **csrRowPtr, int **csrColIdx,
                               float **csrVal, const char *filename)
try
{
    clear_error();
    return mmio_entry<float>(m, n, nnz, isSymmetric, csrRowPtr, csrColIdx, csrVal, filename);
}
TSP_CATCH_INT("tilespmv_mmio_allinone_f32")

} // extern "C"

// spmv.cu -- the persistent tile SpMV kernel for sm_100a and its launch path.
//
// Replaces stir_spmv_cuda_kernel_v6 (/root/reference/src/tilespmv_cuda.h:394-792), the COO replay
// buffers of v5 (:5-392) and the three CSR5 kernels used for the extracted side matrix
// (external/CSR5_cuda/detail/cuda/csr5_spmv_cuda.h:275-420) with ONE kernel:
//
//   * persistent grid (ctas_per_sm x #SM CTAs of 4 warps); every warp owns a static, byte-balanced
//     round-robin slice of the chunk list (chunks are <= chunk_bytes of packed stream, stream.cuh)
//   * each warp runs its own 3-stage TMA pipeline: one lane issues cp.async.bulk (global -> shared,
//     completion on an mbarrier) for the chunk two ahead, so the HBM stream stays in flight
//     independently of the arithmetic; the matrix bytes are read exactly once, fully coalesced,
//     16-byte aligned
//   * the x operand is staged in shared memory one chunk ahead with cp.async (16 B pieces of the
//     16-element segment each tile needs, 4/8 B gathers for the extracted nonzeros)
//   * lane L works on local row L&15, half L>>4 takes every other slot / element; the 16 partial
//     y of a block row live in registers, halves are combined with one shuffle, and y is written
//     once with a coalesced 128-byte store -- no cudaMemset(d_y), no atomics (the reference needs
//     both, tilespmv_cuda.h:784-790, :1116)
//   * block rows cut across chunks write partial sums to a scratch slot; a tiny second kernel adds
//     them in a fixed order (deterministic)
//   * optional fused all-gather epilogue: the y values are also stored into every peer's x buffer
//     over NVLink (P2P pointers), so the repeated-SpMV exchange needs no separate collective
//
// Per-format arithmetic follows tilespmv_cpu.h:138-270 (ELL skips stored zeros like :182).
#include "plan.cuh"

namespace tsp
{

constexpr int SPMV_WARPS = 4;
constexpr int SPMV_THREADS = SPMV_WARPS * 32;
constexpr int SPMV_STAGES = 3;
constexpr int SPMV_BAR_BYTES = 128; // SPMV_WARPS * SPMV_STAGES mbarriers, padded

// ---------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier, TMA bulk copy, cp.async)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completes `bytes` on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase)
{
    uint32_t done;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(phase)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <class T>
__device__ __forceinline__ T fma_t(T a, T b, T c);
template <>
__device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
template <>
__device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }

template <class T>
struct SpmvArgs
{
    const unsigned char *stream;
    const unsigned long long *chunk_off;
    long long nchunks;
    const T *x;
    T *y;
    T *scratch;
    int colA;
    int chunk_bytes, xstage_bytes;
    int npeers;
    long long row_offset;
    T *peers[TSP_MAX_PEERS];
};

// ---------------------------------------------------------------------------------------------
// x staging for one chunk (issued one chunk ahead of its use)
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void stage_x(const unsigned char *st, T *xb, const T *__restrict__ x, int colA, int lane)
{
    const ChunkHeader *h = reinterpret_cast<const ChunkHeader *>(st);
    const int ntiles = h->ntiles;
    const int nside = (int)h->nside;
    const uint2 *tdesc = reinterpret_cast<const uint2 *>(st + h->off_tiledesc);
    const uint32_t *sidecol = reinterpret_cast<const uint32_t *>(st + h->off_sidecol);
    constexpr int VPP = 16 / (int)sizeof(T);  // values per 16-byte piece
    constexpr int PIECES = TS / VPP;          // pieces per 16-element segment
    const uint32_t xb_s = smem_u32(xb);
    for (int i = lane; i < ntiles * PIECES; i += 32)
    {
        const int t = i / PIECES, pc = i % PIECES;
        const long long col0 = (long long)tdesc[t].x * TS + pc * VPP;
        long long left = (long long)colA - col0; // columns of this piece that exist
        left = left < 0 ? 0 : (left > VPP ? VPP : left);
        const T *src = x + (left > 0 ? col0 : 0);
        cp_async_16(xb_s + (uint32_t)((t * TS + pc * VPP) * (int)sizeof(T)), src, (uint32_t)left * (uint32_t)sizeof(T));
    }
    const uint32_t xs_s = xb_s + (uint32_t)(ntiles * TS * (int)sizeof(T));
    for (int e = lane; e < nside; e += 32)
    {
        const T *src = x + sidecol[e];
        if (sizeof(T) == 8)
            cp_async_8(xs_s + (uint32_t)e * 8u, src);
        else
            cp_async_4(xs_s + (uint32_t)e * 4u, src);
    }
}

// ---------------------------------------------------------------------------------------------
// one chunk: all block rows (or row pieces) it holds
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void process_chunk(const unsigned char *st, const T *xb, const SpmvArgs<T> &a, int lane)
{
    const ChunkHeader h = *reinterpret_cast<const ChunkHeader *>(st);
    const uint2 *rows = reinterpret_cast<const uint2 *>(st + 32);
    const uint2 *tdesc = reinterpret_cast<const uint2 *>(st + h.off_tiledesc);
    const uint16_t *sidecnt = reinterpret_cast<const uint16_t *>(st + h.off_sidecnt);
    const T *sideval = reinterpret_cast<const T *>(st + h.off_sideval);
    const unsigned char *pay = st + h.off_payload;
    const T *xside = xb + (int)h.ntiles * TS;
    const int r = lane & 15, hsel = lane >> 4;
    const unsigned half_mask = hsel ? 0xffff0000u : 0x0000ffffu;
    int ti = 0, so = 0;

    for (int rr = 0; rr < (int)h.nrows; rr++)
    {
        const uint2 rec = rows[rr];
        const int nt = (int)(rec.y & 0xffffu);
        const int rowlen = (int)((rec.y >> 16) & 0xffu);
        T acc = 0;
        for (int t = 0; t < nt; t++, ti++)
        {
            const uint2 d = tdesc[ti];
            const int fmt = (int)(d.y & 0xffu);
            const int w = (int)((d.y >> 8) & 0xffu);
            const int aux = (int)(d.y >> 16);
            const T *xs = xb + ti * TS;
            const T *vals = reinterpret_cast<const T *>(pay);
            switch (fmt)
            {
            case TILESPMV_FMT_ELL:
            case TILESPMV_FMT_HYB:
            {
                const unsigned char *idx = pay + w * TS * (int)sizeof(T);
                for (int s = hsel; s < w; s += 2)
                {
                    const int e = s * TS + r;
                    const T v = vals[e];
                    const unsigned b = idx[e >> 1];
                    const unsigned c = (r & 1) ? (b & 15u) : (b >> 4);
                    if (v != (T)0) // stored zeros are skipped like tilespmv_cpu.h:182
                        acc = fma_t<T>(v, xs[c], acc);
                }
                pay += w * TS * (int)sizeof(T) + w * 8;
                break;
            }
            case TILESPMV_FMT_CSR:
            {
                const int nnz = aux;
                const unsigned char *ptr = pay;
                const T *cv = reinterpret_cast<const T *>(pay + 16);
                const uint32_t vbytes = pad8((uint32_t)nnz * (uint32_t)sizeof(T));
                const unsigned char *idx = pay + 16 + vbytes;
                const int start = ptr[r];
                const int end = r == TS - 1 ? nnz : (int)ptr[r + 1];
                for (int k = start + hsel; k < end; k += 2)
                {
                    const unsigned b = idx[k >> 1];
                    const unsigned c = (k & 1) ? (b & 15u) : (b >> 4);
                    acc = fma_t<T>(cv[k], xs[c], acc);
                }
                pay += 16 + vbytes + pad8(((uint32_t)nnz + 1u) / 2u);
                break;
            }
            case TILESPMV_FMT_DENSE:
            {
#pragma unroll
                for (int c = 0; c < TS; c += 2)
                    acc = fma_t<T>(vals[(c + hsel) * TS + r], xs[c + hsel], acc);
                pay += TS * TS * (int)sizeof(T);
                break;
            }
            case TILESPMV_FMT_DENSECOL:
            {
                const unsigned long long ids = *reinterpret_cast<const unsigned long long *>(pay + w * TS * (int)sizeof(T));
                for (int k = hsel; k < w; k += 2)
                {
                    const unsigned c = (unsigned)(ids >> (4 * k)) & 15u;
                    acc = fma_t<T>(vals[k * TS + r], xs[c], acc);
                }
                pay += w * TS * (int)sizeof(T) + 8;
                break;
            }
            case TILESPMV_FMT_DENSEROW:
            {
                // half-warp per dense row: lane = column, 16-lane tree sum, result to the row's lane
                const unsigned mask = (unsigned)aux;
                for (int i = hsel; i < w; i += 2)
                {
                    T p = vals[i * TS + r] * xs[r];
                    p += __shfl_xor_sync(half_mask, p, 8);
                    p += __shfl_xor_sync(half_mask, p, 4);
                    p += __shfl_xor_sync(half_mask, p, 2);
                    p += __shfl_xor_sync(half_mask, p, 1);
                    const int target = (int)__fns(mask, 0, i + 1);
                    if (r == target)
                        acc += p;
                }
                pay += w * TS * (int)sizeof(T);
                break;
            }
            default:
                break;
            }
        }
        if ((rec.y >> 24) & ROWF_HAS_SIDE)
        {
            // extracted very-sparse nonzeros of this block row: 2 lanes per row
            const int cnt = (int)sidecnt[r];
            sidecnt += TS;
            int incl = cnt;
#pragma unroll
            for (int dd = 1; dd < TS; dd <<= 1)
            {
                int o = __shfl_up_sync(0xffffffffu, incl, dd, TS);
                if (r >= dd)
                    incl += o;
            }
            const int total = __shfl_sync(0xffffffffu, incl, TS - 1, TS);
            const int end = so + incl;
            for (int e = end - cnt + hsel; e < end; e += 2)
                acc = fma_t<T>(sideval[e], xside[e], acc);
            so += total;
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (lane < rowlen)
        {
            if (rec.x & ROW_PARTIAL)
                a.scratch[(size_t)(rec.x & ~ROW_PARTIAL) * TS + lane] = acc;
            else
            {
                const size_t row = (size_t)rec.x * TS + lane;
                a.y[row] = acc;
                for (int p = 0; p < a.npeers; p++) // fused all-gather: next x of every peer
                    a.peers[p][a.row_offset + (long long)row] = acc;
            }
        }
    }
}

template <class T>
__global__ void __launch_bounds__(SPMV_THREADS) tile_spmv_kernel(const SpmvArgs<T> a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t per_warp = (uint32_t)(SPMV_STAGES * a.chunk_bytes + 2 * a.xstage_bytes);
    unsigned char *wbase = smem + SPMV_BAR_BYTES + (size_t)warp * per_warp;
    unsigned char *xbase = wbase + (size_t)SPMV_STAGES * a.chunk_bytes;
    const uint32_t bar0 = smem_u32(smem) + (uint32_t)(warp * SPMV_STAGES * 8);

    const long long gw = (long long)blockIdx.x * SPMV_WARPS + warp;
    const long long nw = (long long)gridDim.x * SPMV_WARPS;
    const long long nk = gw < a.nchunks ? (a.nchunks - gw + nw - 1) / nw : 0;
    if (nk == 0)
        return;

    if (lane == 0)
    {
#pragma unroll
        for (int st = 0; st < SPMV_STAGES; st++)
            mbar_init(bar0 + 8u * st, 1);
        fence_mbar_init();
    }
    __syncwarp();

    auto issue = [&](long long k) { // lane 0 only
        const long long c = gw + k * nw;
        const unsigned long long off = a.chunk_off[c];
        const uint32_t bytes = (uint32_t)(a.chunk_off[c + 1] - off);
        const int st = (int)(k % SPMV_STAGES);
        mbar_expect_tx(bar0 + 8u * st, bytes);
        tma_load_1d(smem_u32(wbase + (size_t)st * a.chunk_bytes), a.stream + off, bytes, bar0 + 8u * st);
    };

    if (lane == 0)
        for (long long k = 0; k < SPMV_STAGES && k < nk; k++)
            issue(k);

    mbar_wait(bar0, 0);
    stage_x<T>(wbase, reinterpret_cast<T *>(xbase), a.x, a.colA, lane);
    cp_async_commit();

    for (long long k = 0; k < nk; k++)
    {
        const int st = (int)(k % SPMV_STAGES);
        if (k + 1 < nk)
        {
            const int st1 = (int)((k + 1) % SPMV_STAGES);
            mbar_wait(bar0 + 8u * st1, (uint32_t)(((k + 1) / SPMV_STAGES) & 1));
            stage_x<T>(wbase + (size_t)st1 * a.chunk_bytes,
                       reinterpret_cast<T *>(xbase + (size_t)((k + 1) & 1) * a.xstage_bytes), a.x, a.colA, lane);
        }
        cp_async_commit();
        cp_async_wait<1>(); // x of chunk k has landed (this thread's copies) ...
        __syncwarp();       // ... and everybody else's
        process_chunk<T>(wbase + (size_t)st * a.chunk_bytes,
                         reinterpret_cast<const T *>(xbase + (size_t)(k & 1) * a.xstage_bytes), a, lane);
        __syncwarp(); // all lanes are done reading stage st
        if (lane == 0 && k + SPMV_STAGES < nk)
        {
            fence_proxy_async();
            issue(k + SPMV_STAGES);
        }
    }
    cp_async_wait<0>();
}

// combines the partial sums of block rows that were cut across chunks, in slot order
template <class T>
__global__ void __launch_bounds__(128)
    split_fixup_kernel(const int4 *__restrict__ tab, long long nsplit, const T *__restrict__ scratch, T *__restrict__ y,
                       int npeers, long long row_offset, SpmvArgs<T> a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = g >> 4;
    const int r = (int)(g & 15);
    if (i >= nsplit)
        return;
    const int4 e = tab[i]; // block row, first slot, #slots, rowlen
    if (r >= e.w)
        return;
    T sum = 0;
    for (int k = 0; k < e.z; k++)
        sum += scratch[(size_t)(e.y + k) * TS + r];
    const size_t row = (size_t)e.x * TS + r;
    y[row] = sum;
    for (int p = 0; p < npeers; p++)
        a.peers[p][row_offset + (long long)row] = sum;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t cta_smem_bytes(const tilespmv_plan *P)
{
    return (size_t)SPMV_BAR_BYTES + (size_t)SPMV_WARPS * ((size_t)SPMV_STAGES * P->chunk_bytes + 2 * (size_t)P->xstage_bytes);
}

int spmv_configure(tilespmv_plan *P)
{
    int dev = 0;
    TSP_CUDA(cudaGetDevice(&dev));
    int sms = 0, smem_sm = 0, smem_optin = 0;
    TSP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TSP_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    TSP_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t smem = cta_smem_bytes(P);
    if (smem > (size_t)smem_optin)
    {
        set_error("plan: chunk_bytes/xstage_bytes need %zu B of shared memory per CTA, device allows %d", smem, smem_optin);
        return TILESPMV_ERR_INVALID;
    }
    int fit = (int)((size_t)smem_sm / (smem + 1024)); // 1 KB per CTA is reserved by the driver
    if (fit < 1)
        fit = 1;
    if (P->ctas_per_sm <= 0 || P->ctas_per_sm > fit)
        P->ctas_per_sm = fit;
    P->sm_count = sms;
    P->grid = sms * P->ctas_per_sm;
    P->block = SPMV_THREADS;
    P->smem = (int)smem;
    if (P->precision == 8)
    {
        TSP_CUDA(cudaFuncSetAttribute(tile_spmv_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TSP_CUDA(cudaFuncSetAttribute(tile_spmv_kernel<double>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    else
    {
        TSP_CUDA(cudaFuncSetAttribute(tile_spmv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TSP_CUDA(cudaFuncSetAttribute(tile_spmv_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    return TILESPMV_OK;
}

template <class T>
static int plan_launch_t(tilespmv_plan *P, const T *x, T *y, cudaStream_t s)
{
    if (P->nchunks == 0)
        return TILESPMV_OK;
    SpmvArgs<T> a;
    a.stream = P->stream.as<unsigned char>();
    a.chunk_off = P->chunk_off.as<unsigned long long>();
    a.nchunks = P->nchunks;
    a.x = x;
    a.y = y;
    a.scratch = P->scratch.as<T>();
    a.colA = P->colA;
    a.chunk_bytes = P->chunk_bytes;
    a.xstage_bytes = P->xstage_bytes;
    a.npeers = P->npeers;
    a.row_offset = P->row_offset;
    for (int p = 0; p < TSP_MAX_PEERS; p++)
        a.peers[p] = reinterpret_cast<T *>(P->peers[p]);
    long long warps_needed = P->nchunks;
    int grid = P->grid;
    const long long ctas_needed = (warps_needed + SPMV_WARPS - 1) / SPMV_WARPS;
    if (ctas_needed < grid)
        grid = (int)ctas_needed;
    TSP_LAUNCH((tile_spmv_kernel<T>), grid, SPMV_THREADS, (size_t)P->smem, s, a);
    if (P->nsplit > 0)
    {
        const long long threads = P->nsplit * TS;
        TSP_LAUNCH((split_fixup_kernel<T>), grid_for((size_t)threads, 128), 128, 0, s, P->split_tab.as<int4>(), (long long)P->nsplit,
                   P->scratch.as<T>(), y, P->npeers, (long long)P->row_offset, a);
    }
    return TILESPMV_OK;
}

int plan_launch(tilespmv_plan *P, const void *d_x, void *d_y, cudaStream_t s)
{
    if ((reinterpret_cast<uintptr_t>(d_x) & 15u) || (reinterpret_cast<uintptr_t>(d_y) & 15u))
    {
        set_error("spmv: x and y must be 16-byte aligned device pointers");
        return TILESPMV_ERR_INVALID;
    }
    if (P->precision == 8)
        return plan_launch_t<double>(P, static_cast<const double *>(d_x), static_cast<double *>(d_y), s);
    return plan_launch_t<float>(P, static_cast<const float *>(d_x), static_cast<float *>(d_y), s);
}

} // namespace tsp

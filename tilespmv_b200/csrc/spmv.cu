// spmv.cu -- the persistent tile SpMV kernel for sm_100a and its launch path.
//
// Replaces stir_spmv_cuda_kernel_v6 (/root/reference/src/tilespmv_cuda.h:394-792), the COO replay
// buffers of v5 (:5-392) and the three CSR5 kernels used for the extracted side matrix
// (external/CSR5_cuda/detail/cuda/csr5_spmv_cuda.h:275-420) with ONE kernel:
//
//   * persistent grid (one CTA per SM with as many warps as shared memory holds); every warp owns a
//     static, byte-balanced round-robin slice of the chunk list (chunks are <= chunk_bytes of packed stream, stream.cuh)
//   * each warp runs its own 4-stage TMA pipeline: one lane issues cp.async.bulk (global -> shared,
//     completion on an mbarrier) for the chunk three ahead, so the HBM stream stays in flight
//     independently of the arithmetic; the matrix bytes are read exactly once, fully coalesced,
//     16-byte aligned
//   * the x operand is staged in shared memory one chunk ahead with cp.async (16 B pieces of the
//     16-element segment each tile needs, 4/8 B gathers for the extracted nonzeros)
//   * lane L owns local rows 2(L&7), 2(L&7)+1 (one 128-bit shared-memory load fetches both values);
//     the ELL tiles of a block row are one flat list of 16-value slot-rows, so the inner loop has no
//     per-tile control flow; the 4 lane groups are combined with two shuffles and y is written once
//     with a coalesced 128-byte store -- no cudaMemset(d_y), no atomics (the reference needs
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

constexpr int SPMV_MAX_WARPS = 16;
constexpr int SPMV_MAX_STAGES = 4;
constexpr int SPMV_BARS_PER_WARP = SPMV_MAX_STAGES + 2;            // chunk stages + 2 x-staging buffers
constexpr int SPMV_BAR_BYTES = SPMV_MAX_WARPS * SPMV_BARS_PER_WARP * 8 + 256; // padded to 128 B

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
// x staging for one chunk (issued one chunk ahead of its use):
//   * the 16-element x segment of every stream tile: ONE TMA bulk copy per tile, issued by lane t
//     for tile t (a single warp instruction stages up to 32 segments), completion on `xbar`
//   * the x values of the extracted nonzeros: one 8 B / 4 B cp.async gather per nonzero
// A segment that sticks out past colA (last tile column of a matrix whose width is not a multiple
// of 16) is staged element-wise with zero fill by its lane instead.
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void stage_x(const unsigned char *st, T *xb, uint32_t xbar, const T *__restrict__ x,
                                        int colA, int lane)
{
    const ChunkHeader *h = reinterpret_cast<const ChunkHeader *>(st);
    const int ntiles = h->ntiles;
    const int nside = (int)h->nside;
    const uint32_t *tilecol = reinterpret_cast<const uint32_t *>(st + CHUNK_OFF_ROWS + 16u * h->nrows);
    const uint32_t *sidecol = reinterpret_cast<const uint32_t *>(st + h->off_sidecol);
    const uint32_t xb_s = smem_u32(xb);
    constexpr uint32_t SEG = TS * (uint32_t)sizeof(T);
    // pass 1: how many full segments (bytes the barrier has to expect)
    uint32_t nfull = 0;
#pragma unroll 1
    for (int t0 = 0; t0 < ntiles; t0 += 32)
    {
        const int t = t0 + lane;
        const bool full = t < ntiles && (int)(tilecol[t] * TS + TS) <= colA;
        nfull += __popc(__ballot_sync(0xffffffffu, full));
    }
    if (lane == 0)
        mbar_expect_tx(xbar, nfull * SEG);
    __syncwarp();
#pragma unroll 1
    for (int t = lane; t < ntiles; t += 32)
    {
        const int col0 = (int)(tilecol[t] * TS);
        if (col0 + TS <= colA)
            tma_load_1d(xb_s + (uint32_t)t * SEG, x + col0, SEG, xbar);
        else
        {
#pragma unroll 1
            for (int c = 0; c < TS; c++)
                xb[t * TS + c] = col0 + c < colA ? x[col0 + c] : (T)0;
        }
    }
    const uint32_t xs_s = xb_s + (uint32_t)ntiles * SEG;
#pragma unroll 1
    for (int e = lane; e < nside; e += 32)
    {
        const T *src = x + sidecol[e];
        if (sizeof(T) == 8)
            cp_async_8(xs_s + (uint32_t)e * 8u, src);
        else
            cp_async_4(xs_s + (uint32_t)e * 4u, src);
    }
}

template <class T>
struct Vec2;
template <>
struct Vec2<double>
{
    typedef double2 type;
};
template <>
struct Vec2<float>
{
    typedef float2 type;
};

// ---------------------------------------------------------------------------------------------
// one chunk: all block rows (or row pieces) it holds.
// Lane mapping: p = lane & 7 owns local rows 2p and 2p+1 (accumulators a0 / a1); g = lane >> 3
// selects every 4th slot-row / element, so a warp covers 4 slot-rows (64 values) per iteration
// with one 128-bit shared-memory load per lane.
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void process_chunk(const unsigned char *st, const T *xb, const SpmvArgs<T> &a, int lane)
{
    typedef typename Vec2<T>::type V2;
    const ChunkHeader h = *reinterpret_cast<const ChunkHeader *>(st);
    const uint4 *rows = reinterpret_cast<const uint4 *>(st + CHUNK_OFF_ROWS);
    const uint2 *odesc = reinterpret_cast<const uint2 *>(st + h.off_odesc);
    const unsigned char *sidehdr = st + h.off_sidehdr;
    const T *sideval = reinterpret_cast<const T *>(st + h.off_sideval);
    const unsigned char *pay = st + h.off_payload;
    const T *xside = xb + (int)h.ntiles * TS;
    const int p = lane & 7, g = lane >> 3;
    const int nrows = (int)h.nrows;

    for (int rr = 0; rr < nrows; rr++)
    {
        const uint4 rec = rows[rr];
        const int nsr = (int)(rec.y & 0xffffu);
        const int nother = (int)(rec.y >> 16);
        const int rowlen = (int)(rec.z & 0xffu);
        const unsigned flags = (rec.z >> 8) & 0xffu;
        T a0 = 0, a1 = 0;

        // ---- ELL group: one flat loop over the slot-rows of all ELL tiles of the row ----
        {
            const V2 *vals = reinterpret_cast<const V2 *>(pay);
            const unsigned char *idx = pay + nsr * TS * (int)sizeof(T);
            const unsigned char *xsel = idx + nsr * 8;
#pragma unroll 2
            for (int sr = g; sr < nsr; sr += 4)
            {
                const V2 v = vals[sr * 8 + p];
                const unsigned b = idx[sr * 8 + p];
                const T *xs = xb + (int)xsel[sr] * TS;
                // padding slots hold value 0 / column 0 and are multiplied through like in the
                // reference GPU kernel (tilespmv_cuda.h:597-598); only tilespmv_cpu.h:182 skips them
                a0 = fma_t<T>(v.x, xs[b >> 4], a0);
                a1 = fma_t<T>(v.y, xs[b & 15u], a1);
            }
            pay += nsr * TS * (int)sizeof(T) + (int)pad16((uint32_t)nsr * 9u);
        }

        // ---- the other tiles of the row ----
        for (int t = 0; t < nother; t++)
        {
            const uint2 d = *odesc++;
            const int fmt = (int)(d.x & 0xffu);
            const T *xs = xb + (int)((d.x >> 8) & 0xffu) * TS;
            const int w = (int)(d.x >> 16);
            const T *vals = reinterpret_cast<const T *>(pay);
            switch (fmt)
            {
            case TILESPMV_FMT_CSR:
            {
                const int nnz = (int)d.y;
                const unsigned char *ptr = pay;
                const T *cv = reinterpret_cast<const T *>(pay + 16);
                const uint32_t vbytes = pad8((uint32_t)nnz * (uint32_t)sizeof(T));
                const unsigned char *idx = pay + 16 + vbytes;
                const int s0 = ptr[2 * p], s1 = ptr[2 * p + 1];
                const int e1 = p == 7 ? nnz : (int)ptr[2 * p + 2];
#pragma unroll 1
                for (int k = s0 + g; k < s1; k += 4)
                {
                    const unsigned b = idx[k >> 1];
                    a0 = fma_t<T>(cv[k], xs[(k & 1) ? (b & 15u) : (b >> 4)], a0);
                }
#pragma unroll 1
                for (int k = s1 + g; k < e1; k += 4)
                {
                    const unsigned b = idx[k >> 1];
                    a1 = fma_t<T>(cv[k], xs[(k & 1) ? (b & 15u) : (b >> 4)], a1);
                }
                pay += pad16(16u + vbytes + pad8(((uint32_t)nnz + 1u) / 2u));
                break;
            }
            case TILESPMV_FMT_DENSE:
            {
                const V2 *dv = reinterpret_cast<const V2 *>(pay);
#pragma unroll
                for (int c = 0; c < TS; c += 4)
                {
                    const V2 v = dv[(c + g) * 8 + p];
                    const T xc = xs[c + g];
                    a0 = fma_t<T>(v.x, xc, a0);
                    a1 = fma_t<T>(v.y, xc, a1);
                }
                pay += TS * TS * (int)sizeof(T);
                break;
            }
            case TILESPMV_FMT_DENSECOL:
            {
                const V2 *dv = reinterpret_cast<const V2 *>(pay);
                const unsigned long long ids = *reinterpret_cast<const unsigned long long *>(pay + w * TS * (int)sizeof(T));
                for (int k = g; k < w; k += 4)
                {
                    const V2 v = dv[k * 8 + p];
                    const T xc = xs[(unsigned)(ids >> (4 * k)) & 15u];
                    a0 = fma_t<T>(v.x, xc, a0);
                    a1 = fma_t<T>(v.y, xc, a1);
                }
                pay += pad16((uint32_t)(w * TS) * (uint32_t)sizeof(T) + 8u);
                break;
            }
            case TILESPMV_FMT_DENSEROW:
            {
                // half-warp per dense row: lane = column, 16-lane tree sum, result to the row's owner
                const unsigned mask = d.y;
                const int hsel = lane >> 4, c = lane & 15;
                const unsigned half_mask = hsel ? 0xffff0000u : 0x0000ffffu;
                for (int i = hsel; i < w; i += 2)
                {
                    T pr = vals[i * TS + c] * xs[c];
                    pr += __shfl_xor_sync(half_mask, pr, 8);
                    pr += __shfl_xor_sync(half_mask, pr, 4);
                    pr += __shfl_xor_sync(half_mask, pr, 2);
                    pr += __shfl_xor_sync(half_mask, pr, 1);
                    const int target = (int)__fns(mask, 0, i + 1);
                    if (c == (target >> 1)) // lanes (p = target/2, g = 2*hsel): g is 0 or 2 for c < 8
                    {
                        if (target & 1)
                            a1 += pr;
                        else
                            a0 += pr;
                    }
                }
                pay += w * TS * (int)sizeof(T);
                break;
            }
            default:
                break;
            }
        }

        // ---- extracted very-sparse nonzeros of this block row: 4 lanes per row pair ----
        if (flags & ROWF_HAS_SIDE)
        {
            const uint16_t *sh = reinterpret_cast<const uint16_t *>(sidehdr);
            sidehdr += SIDEHDR_BYTES;
            const int s0 = sh[2 * p], s1 = sh[2 * p + 1], e1 = sh[2 * p + 2];
#pragma unroll 1
            for (int e = s0 + g; e < s1; e += 4)
                a0 = fma_t<T>(sideval[e], xside[e], a0);
#pragma unroll 1
            for (int e = s1 + g; e < e1; e += 4)
                a1 = fma_t<T>(sideval[e], xside[e], a1);
            const int total = sh[16];
            sideval += total;
            xside += total;
        }

        // ---- combine the 4 lane groups, lanes 0..7 store rows (2p, 2p+1) with one 16-byte store ----
        a0 += __shfl_xor_sync(0xffffffffu, a0, 8);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 8);
        a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
        if (lane < 8 && 2 * lane < rowlen)
        {
            const bool partial = (rec.x & ROW_PARTIAL) != 0;
            const size_t row = (size_t)(rec.x & ~ROW_PARTIAL) * TS + 2 * lane;
            T *dst = (partial ? a.scratch : a.y) + row;
            if (2 * lane + 1 < rowlen)
            {
                V2 o;
                o.x = a0;
                o.y = a1;
                *reinterpret_cast<V2 *>(dst) = o;
            }
            else
                dst[0] = a0;
            if (!partial)
                for (int q = 0; q < a.npeers; q++) // fused all-gather: next x of every peer
                {
                    T *px = a.peers[q] + a.row_offset + (long long)row;
                    px[0] = a0;
                    if (2 * lane + 1 < rowlen)
                        px[1] = a1;
                }
        }
    }
}

template <class T, int SPMV_STAGES>
__global__ void __launch_bounds__(SPMV_MAX_WARPS * 32, 1) tile_spmv_kernel(const SpmvArgs<T> a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps_cta = blockDim.x >> 5;
    const uint32_t per_warp = (uint32_t)(SPMV_STAGES * a.chunk_bytes + 2 * a.xstage_bytes);
    unsigned char *wbase = smem + SPMV_BAR_BYTES + (size_t)warp * per_warp;
    unsigned char *xbase = wbase + (size_t)SPMV_STAGES * a.chunk_bytes;
    // per warp: SPMV_STAGES barriers for the chunk stream + 2 for the staged x segments
    const uint32_t bar0 = smem_u32(smem) + (uint32_t)(warp * SPMV_BARS_PER_WARP * 8);
    const uint32_t xbar0 = bar0 + 8u * SPMV_MAX_STAGES;

    // warp w of CTA b takes chunks gw, gw + nw, ...: neighbouring warps stream neighbouring chunks
    const long long gw = (long long)blockIdx.x * nwarps_cta + warp;
    const long long nw = (long long)gridDim.x * nwarps_cta;
    const int nk = gw < a.nchunks ? (int)((a.nchunks - gw + nw - 1) / nw) : 0;
    if (nk == 0)
        return;

    if (lane == 0)
    {
#pragma unroll
        for (int i = 0; i < SPMV_BARS_PER_WARP; i++)
            mbar_init(bar0 + 8u * i, 1);
        fence_mbar_init();
    }
    __syncwarp();

    const uint32_t stage0 = smem_u32(wbase);
    auto issue = [&](int k, unsigned long long off, unsigned long long end) { // lane 0 only
        const int st = k % SPMV_STAGES;
        const uint32_t bytes = (uint32_t)(end - off);
        mbar_expect_tx(bar0 + 8u * st, bytes);
        tma_load_1d(stage0 + (uint32_t)st * (uint32_t)a.chunk_bytes, a.stream + off, bytes, bar0 + 8u * st);
    };
    // lane 0 keeps the offsets of the NEXT chunk to issue in registers, loaded one iteration early
    unsigned long long nxt_off = 0, nxt_end = 0;
    if (lane == 0)
    {
        for (int k = 0; k < SPMV_STAGES && k < nk; k++)
        {
            const long long c = gw + (long long)k * nw;
            issue(k, a.chunk_off[c], a.chunk_off[c + 1]);
        }
        if (SPMV_STAGES < nk)
        {
            const long long c = gw + (long long)SPMV_STAGES * nw;
            nxt_off = a.chunk_off[c];
            nxt_end = a.chunk_off[c + 1];
        }
    }

    mbar_wait(bar0, 0);
    stage_x<T>(wbase, reinterpret_cast<T *>(xbase), xbar0, a.x, a.colA, lane);
    cp_async_commit();

#pragma unroll 1
    for (int k = 0; k < nk; k++)
    {
        const int st = k % SPMV_STAGES;
        if (k + 1 < nk)
        {
            const int st1 = (k + 1) % SPMV_STAGES;
            mbar_wait(bar0 + 8u * st1, (uint32_t)(((k + 1) / SPMV_STAGES) & 1));
            stage_x<T>(wbase + (size_t)st1 * a.chunk_bytes,
                       reinterpret_cast<T *>(xbase + (size_t)((k + 1) & 1) * a.xstage_bytes), xbar0 + 8u * ((k + 1) & 1),
                       a.x, a.colA, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();                                          // gathers of chunk k (this thread's) ...
        mbar_wait(xbar0 + 8u * (k & 1), (uint32_t)((k >> 1) & 1));  // ... its x segments ...
        __syncwarp();                                                // ... and everybody else's copies
        process_chunk<T>(wbase + (size_t)st * a.chunk_bytes,
                         reinterpret_cast<const T *>(xbase + (size_t)(k & 1) * a.xstage_bytes), a, lane);
        __syncwarp(); // all lanes are done reading stage st and x buffer k&1
        if (lane == 0 && k + SPMV_STAGES < nk)
        {
            fence_proxy_async();
            issue(k + SPMV_STAGES, nxt_off, nxt_end);
            if (k + SPMV_STAGES + 1 < nk)
            {
                const long long c = gw + (long long)(k + SPMV_STAGES + 1) * nw;
                nxt_off = a.chunk_off[c];
                nxt_end = a.chunk_off[c + 1];
            }
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
static size_t warp_smem_bytes(const tilespmv_plan *P)
{
    return (size_t)P->stages * P->chunk_bytes + 2 * (size_t)P->xstage_bytes;
}

template <class T>
static int set_kernel_attrs(int stages, int smem)
{
    const void *fn = stages == 2   ? (const void *)tile_spmv_kernel<T, 2>
                     : stages == 3 ? (const void *)tile_spmv_kernel<T, 3>
                                   : (const void *)tile_spmv_kernel<T, 4>;
    TSP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TSP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return TILESPMV_OK;
}

// One persistent CTA per SM (ctas_per_sm can raise it); the CTA gets as many independent warps as
// its shared memory holds -- every warp owns SPMV_STAGES chunk buffers + 2 x-staging buffers.
int spmv_configure(tilespmv_plan *P)
{
    int dev = 0;
    TSP_CUDA(cudaGetDevice(&dev));
    int sms = 0, smem_optin = 0;
    TSP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TSP_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (P->ctas_per_sm <= 0)
        P->ctas_per_sm = 1;
    if (P->stages < 2 || P->stages > SPMV_MAX_STAGES)
        P->stages = 4;
    if (P->ctas_per_sm > 8)
        P->ctas_per_sm = 8;
    const size_t budget = ((size_t)smem_optin + 1024) / P->ctas_per_sm - 1024; // 1 KB per CTA is reserved
    const size_t per_warp = warp_smem_bytes(P);
    if (budget < SPMV_BAR_BYTES + per_warp)
    {
        set_error("plan: chunk_bytes/xstage_bytes need %zu B of shared memory per warp, only %zu available",
                  per_warp, budget - SPMV_BAR_BYTES);
        return TILESPMV_ERR_INVALID;
    }
    int warps = (int)((budget - SPMV_BAR_BYTES) / per_warp);
    if (warps > SPMV_MAX_WARPS / P->ctas_per_sm)
        warps = SPMV_MAX_WARPS / P->ctas_per_sm;
    if (warps < 1)
        warps = 1;
    const size_t smem = SPMV_BAR_BYTES + (size_t)warps * per_warp;
    P->sm_count = sms;
    P->grid = sms * P->ctas_per_sm;
    P->block = warps * 32;
    P->smem = (int)smem;
    if (P->precision == 8)
        TSP_TRY(set_kernel_attrs<double>(P->stages, (int)smem));
    else
        TSP_TRY(set_kernel_attrs<float>(P->stages, (int)smem));
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
    const int warps = P->block / 32;
    int grid = P->grid;
    const long long ctas_needed = (P->nchunks + warps - 1) / warps;
    if (ctas_needed < grid)
        grid = (int)ctas_needed;
    if (P->stages == 2)
        TSP_LAUNCH((tile_spmv_kernel<T, 2>), grid, P->block, (size_t)P->smem, s, a);
    else if (P->stages == 3)
        TSP_LAUNCH((tile_spmv_kernel<T, 3>), grid, P->block, (size_t)P->smem, s, a);
    else
        TSP_LAUNCH((tile_spmv_kernel<T, 4>), grid, P->block, (size_t)P->smem, s, a);
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

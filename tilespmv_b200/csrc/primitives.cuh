// primitives.cuh -- hand-written device-wide primitives used by the conversion and the planner:
//   * exclusive prefix sum over a functor input (reduce / single-CTA scan / apply)
//   * stable LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass
// Plain SIMT + warp shuffles / match_any; no library (CUB/Thrust) calls.
#pragma once
#include "common.cuh"

namespace tsp
{

// =============================================================================================
// exclusive scan
// =============================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long warp_incl_scan(long long v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        long long o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d)
            v += o;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix
// and leaves the block total in *total (valid for all threads)
__device__ __forceinline__ long long block_excl_scan(long long v, long long *total)
{
    __shared__ long long warp_sums[SCAN_THREADS / 32];
    __shared__ long long block_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long incl = warp_incl_scan(v, lane);
    if (lane == 31)
        warp_sums[w] = incl;
    __syncthreads();
    if (w == 0)
    {
        long long s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        long long si = warp_incl_scan(s, lane);
        if (lane < SCAN_THREADS / 32)
            warp_sums[lane] = si - s;
        if (lane == SCAN_THREADS / 32 - 1)
            block_total = si;
    }
    __syncthreads();
    long long r = warp_sums[w] + incl - v;
    *total = block_total;
    __syncthreads(); // shared scratch is reused by the next call
    return r;
}

template <class InF>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(InF in, size_t n, long long *block_sums)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    long long s = 0;
#pragma unroll 4
    for (int k = 0; k < SCAN_ITEMS; k++)
    {
        size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x; // strided: coalesced reads
        if (i < n)
            s += in(i);
    }
    long long total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0)
        block_sums[blockIdx.x] = total;
}

// in-place exclusive scan of the per-block sums by ONE CTA; writes the grand total to *total
static __global__ void __launch_bounds__(SCAN_THREADS) scan_blocksums_kernel(long long *sums, size_t nb, long long *total)
{
    long long carry = 0;
    for (size_t base = 0; base < nb; base += SCAN_THREADS)
    {
        size_t i = base + threadIdx.x;
        long long v = i < nb ? sums[i] : 0;
        long long t;
        long long e = block_excl_scan(v, &t);
        if (i < nb)
            sums[i] = carry + e;
        carry += t;
    }
    if (threadIdx.x == 0)
        *total = carry;
}

template <class InF, class OutT>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(InF in, size_t n, const long long *block_offsets, OutT *out)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
    {
        size_t i = base + k;
        v[k] = i < n ? in(i) : 0;
        s += v[k];
    }
    long long total;
    long long run = block_excl_scan(s, &total) + block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
    {
        size_t i = base + k;
        if (i < n)
            out[i] = (OutT)run;
        run += v[k];
    }
}

struct ScanWorkspace
{
    DevBuf sums;  // per-block sums
    DevBuf total; // one long long
    int reserve(size_t n)
    {
        size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE + 1;
        if (sums.bytes < nb * sizeof(long long))
            TSP_TRY(sums.alloc(nb * sizeof(long long), false));
        if (!total.p)
            TSP_TRY(total.alloc(sizeof(long long), false));
        return TILESPMV_OK;
    }
};

// out[i] = sum_{k<i} in(k) for i < n (out may alias the array `in` reads: every thread loads its
// items before storing).  If total_host != nullptr the grand total is copied back (synchronises).
template <class InF, class OutT>
int exclusive_scan(InF in, size_t n, OutT *out, ScanWorkspace &ws, cudaStream_t s, long long *total_host)
{
    if (n == 0)
    {
        if (total_host)
            *total_host = 0;
        return TILESPMV_OK;
    }
    TSP_TRY(ws.reserve(n));
    const unsigned nb = (unsigned)((n + SCAN_TILE - 1) / SCAN_TILE);
    TSP_LAUNCH((scan_reduce_kernel<InF>), nb, SCAN_THREADS, 0, s, in, n, ws.sums.as<long long>());
    TSP_LAUNCH(scan_blocksums_kernel, 1, SCAN_THREADS, 0, s, ws.sums.as<long long>(), (size_t)nb,
               ws.total.as<long long>());
    TSP_LAUNCH((scan_apply_kernel<InF, OutT>), nb, SCAN_THREADS, 0, s, in, n, ws.sums.as<long long>(), out);
    if (total_host)
    {
        TSP_CUDA(cudaMemcpyAsync(total_host, ws.total.p, sizeof(long long), cudaMemcpyDeviceToHost, s));
        TSP_CUDA(cudaStreamSynchronize(s));
        if (sizeof(OutT) == 4 && *total_host > 0x7fffffffll)
        {
            set_error("prefix sum total %lld overflows the reference's int indexing", *total_host);
            return TILESPMV_ERR_UNSUPPORTED;
        }
    }
    return TILESPMV_OK;
}

struct IntArrayIn
{
    const int *a;
    __device__ __forceinline__ int operator()(size_t i) const { return a[i]; }
};

// =============================================================================================
// stable LSD radix sort (u64 keys, u32 values), 8-bit digits
// =============================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS; // 4096 keys per CTA
constexpr int RS_WARP_SPAN = 32 * RS_ITEMS;    // 512 consecutive keys per warp

static __global__ void __launch_bounds__(RS_THREADS)
    radix_hist_kernel(const uint64_t *__restrict__ keys, size_t n, int shift, int *__restrict__ ghist, unsigned nblocks)
{
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; k++)
    {
        size_t i = base + (size_t)k * RS_THREADS + threadIdx.x;
        if (i < n)
            atomicAdd(&hist[(unsigned)(keys[i] >> shift) & 255u], 1);
    }
    __syncthreads();
    ghist[(size_t)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x]; // digit-major for the scan
}

// Stable scatter: inside a CTA the keys are visited in index order (warp w owns 512 consecutive
// keys, 32 at a time); the rank of a key among equal digits is
//   global base of (digit, CTA) + #equal digits in earlier warps + #equal digits earlier in my warp.
static __global__ void __launch_bounds__(RS_THREADS)
    radix_scatter_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
                         uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, size_t n, int shift,
                         const int *__restrict__ ghist_scanned, unsigned nblocks)
{
    __shared__ int cnt[RS_THREADS / 32][256];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (RS_THREADS / 32) * 256; i += RS_THREADS)
        (&cnt[0][0])[i] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE + (size_t)w * RS_WARP_SPAN + lane;
    uint64_t k[RS_ITEMS];
    uint32_t v[RS_ITEMS];
    int rk[RS_ITEMS];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++)
    {
        size_t i = base + (size_t)it * 32;
        bool valid = i < n;
        k[it] = valid ? kin[i] : 0;
        v[it] = valid ? vin[i] : 0;
    }
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++)
    {
        bool valid = base + (size_t)it * 32 < n;
        unsigned d = (unsigned)(k[it] >> shift) & 255u;
        unsigned m = __match_any_sync(0xffffffffu, valid ? d : 256u + lane);
        int prev = valid ? cnt[w][d] : 0;
        __syncwarp();
        int r = __popc(m & lt);
        if (valid && r == 0)
            cnt[w][d] = prev + __popc(m);
        __syncwarp();
        rk[it] = prev + r;
    }
    __syncthreads();
    {
        int run = ghist_scanned[(size_t)tid * nblocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_THREADS / 32; ww++)
        {
            int t = cnt[ww][tid];
            cnt[ww][tid] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++)
    {
        if (base + (size_t)it * 32 < n)
        {
            unsigned d = (unsigned)(k[it] >> shift) & 255u;
            size_t pos = (size_t)(cnt[w][d] + rk[it]);
            kout[pos] = k[it];
            vout[pos] = v[it];
        }
    }
}

// Sorts n pairs by key bits [begin_bit, end_bit).  keys/vals and keys_alt/vals_alt are
// ping-pong buffers of n entries; on return *keys_out / *vals_out point at the sorted data
// (one of the two buffers).  Needs n < 2^31.
inline int radix_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t *keys_alt, uint32_t *vals_alt, size_t n,
                            int begin_bit, int end_bit, ScanWorkspace &ws, cudaStream_t s,
                            uint64_t **keys_out, uint32_t **vals_out)
{
    *keys_out = keys;
    *vals_out = vals;
    if (n == 0 || end_bit <= begin_bit)
        return TILESPMV_OK;
    const unsigned nblocks = (unsigned)((n + RS_TILE - 1) / RS_TILE);
    DevBuf ghist;
    TSP_TRY(ghist.alloc((size_t)256 * nblocks * sizeof(int), false));
    uint64_t *ka = keys, *kb = keys_alt;
    uint32_t *va = vals, *vb = vals_alt;
    for (int shift = begin_bit; shift < end_bit; shift += 8)
    {
        TSP_LAUNCH(radix_hist_kernel, nblocks, RS_THREADS, 0, s, ka, n, shift, ghist.as<int>(), nblocks);
        TSP_TRY(exclusive_scan(IntArrayIn{ghist.as<int>()}, (size_t)256 * nblocks, ghist.as<int>(), ws, s, nullptr));
        TSP_LAUNCH(radix_scatter_kernel, nblocks, RS_THREADS, 0, s, ka, va, kb, vb, n, shift, ghist.as<int>(), nblocks);
        uint64_t *tk = ka;
        ka = kb;
        kb = tk;
        uint32_t *tv = va;
        va = vb;
        vb = tv;
    }
    TSP_CUDA(cudaStreamSynchronize(s)); // ghist is freed on return
    *keys_out = ka;
    *vals_out = va;
    return TILESPMV_OK;
}

// binary searches used by several kernels --------------------------------------------------
// largest i in [0, n) with a[i] <= v, given a[0] <= v (a non-decreasing)
__device__ __forceinline__ int upper_row(const int *__restrict__ a, int n, int v)
{
    int lo = 0, hi = n; // invariant: a[lo] <= v, (hi == n or a[hi] > v)
    while (hi - lo > 1)
    {
        int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
        if (a[mid] <= v)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}
// first i in [0, n] with a[i] >= v (a non-decreasing)
__device__ __forceinline__ int lower_bound_dev(const int *__restrict__ a, int n, int v)
{
    int lo = 0, hi = n;
    while (lo < hi)
    {
        int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
        if (a[mid] < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

} // namespace tsp

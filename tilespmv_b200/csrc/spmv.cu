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

constexpr int SPMV_MAX_WARPS = 24; // sizes the barrier area
// warps per CTA by pipeline depth: what 227 KB of shared memory hold with 4 KB stages, and the
// register budget that goes with it (65536 / threads): 20 warps -> 96 registers (register files are handed out 4 warps at a time), 16 -> 128
__host__ __device__ constexpr int spmv_max_warps(int stages) { return stages == 2 ? 21 : (stages == 3 ? 16 : 12); }
// the 2-stage kernel exists in two register budgets: 80 (up to 21 warps; register files are handed
// out 4 warps at a time, so 21 warps count as 24) and 96 (up to 20 warps)
constexpr int SPMV_REGS_LO = 80, SPMV_REGS_HI = 96;
constexpr int SPMV_MAX_STAGES = 4;
constexpr int SPMV_BARS_PER_WARP = SPMV_MAX_STAGES;                // one mbarrier per chunk stage
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
// the same with an L2 eviction-priority hint (the packed stream is read exactly once per SpMV: evict_first keeps it
// from pushing the re-used x lines out of L2)
__device__ __forceinline__ void tma_load_1d_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
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
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
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

// explicit shared-state-space loads on 32-bit addresses (the hot loops do their own address
// arithmetic; ptxas folds constant offsets into the instruction's immediate)
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
// a * b + c in one integer multiply-add (keeps ptxas from expanding nibble * size into shift+mask+add)
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// p + a * b as one 64-bit multiply-add (global address of a tile's x segment / a side column)
__device__ __forceinline__ const void *mad_wide(uint32_t a, uint32_t b, const void *p)
{
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(reinterpret_cast<unsigned long long>(p)));
    return reinterpret_cast<const void *>(r);
}
template <class T>
struct SL;
template <>
struct SL<double>
{
    typedef double2 V2;
    static __device__ __forceinline__ double ld(uint32_t a)
    {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ double2 ld2(uint32_t a)
    {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
        return v;
    }
};
template <>
struct SL<float>
{
    typedef float2 V2;
    static __device__ __forceinline__ float ld(uint32_t a)
    {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ float2 ld2(uint32_t a)
    {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
        return v;
    }
};

template <class T>
struct SpmvArgs
{
    const unsigned char *stream;
    const uint2 *chunk_desc; // {offset / 16, bytes} per chunk
    unsigned nchunks;
    const T *x;
    T *y;
    T *scratch;
    int colA;
    const unsigned char *head; // x-staging lists of every warp's first chunk
    int head_stride;
    int stage_stride, xstage_bytes;
    int npeers;
    int accumulate; // y += A*x (column-panel sub-plans) instead of y = A*x
    long long row_offset;
    T *peers[TSP_MAX_PEERS];
    // peer q receives the local rows [peer_lo[q], peer_hi[q]) only: everything for the plain fused exchange, just the rows
    // its next launch reads (the halo) when the copy engines replicate the rest in the background (comm.cu)
    long long peer_lo[TSP_MAX_PEERS], peer_hi[TSP_MAX_PEERS];
    // gather-bound plans: the window of x this launch reads, pulled into L2 by bulk prefetches when the launch starts
    // (first touches through scattered 8-byte gathers run at the DRAM gather rate, 52 G/s against 246 G/s from L2)
    const unsigned char *pf_base;
    unsigned long long pf_bytes;
};

// ---------------------------------------------------------------------------------------------
// x staging for one chunk (issued one chunk ahead of its use), all by cp.async (SASS LDGSTS):
//   * the 16-element x segment of every stream tile as 16-byte pieces: 8 (fp64) / 4 (fp32) lanes
//     per segment, so one warp instruction stages 4 / 8 segments, fully coalesced per segment
//   * the x values of the extracted nonzeros: one 8 B / 4 B gather per nonzero
// A segment that sticks out past colA (last tile column of a matrix whose width is not a multiple
// of 16) is zero-filled through the src-size operand (chunks flagged CHF_PARTIAL_X only).
// Completion is tracked by the cp.async group of the calling iteration (no mbarrier).
// ---------------------------------------------------------------------------------------------
// list access: shared memory (lists carried by the previous chunk) or global (head array)
template <bool GLOBAL>
__device__ __forceinline__ uint32_t list_u32(uint32_t s_addr, const uint32_t *g_ptr, uint32_t byte_off)
{
    if (GLOBAL)
        return __ldg(g_ptr + byte_off / 4u);
    return lds_u32(s_addr + byte_off);
}

template <class T, bool GLOBAL>
__device__ __forceinline__ void stage_x(uint32_t list_s, const uint32_t *list_g, int ntiles, int nside, unsigned flags,
                                        uint32_t xb_s, const T *__restrict__ x, const T *__restrict__ xpiece, int colA,
                                        int lane)
{
    constexpr int PPT = TS * (int)sizeof(T) / 16; // 16-byte pieces per segment
    constexpr int TPI = 32 / PPT;                 // segments per warp instruction
    constexpr int EPP = 16 / (int)sizeof(T);      // elements per piece
    const uint32_t tl = 4u * (uint32_t)(lane / PPT);                                      // this lane's tile column ...
    const uint32_t sl = ((4u * (uint32_t)ntiles + 15u) & ~15u) + 4u * (uint32_t)lane;     // ... and side column
    uint32_t dst = xb_s + (uint32_t)lane * 16u;
    if (!(flags & CHF_PARTIAL_X))
    {
        // xpiece = x + (lane % PPT) * EPP: one 64-bit multiply-add per copy
#pragma unroll 1
        for (int t0 = 0; t0 < ntiles; t0 += 2 * TPI)
        {
            const int t = t0 + lane / PPT;
            const bool p0 = t < ntiles, p1 = t + TPI < ntiles;
            uint32_t c0 = 0, c1 = 0;
            if (p0)
                c0 = list_u32<GLOBAL>(list_s, list_g, tl + 4u * (uint32_t)t0);
            if (p1)
                c1 = list_u32<GLOBAL>(list_s, list_g, tl + 4u * (uint32_t)(t0 + TPI));
            if (p0)
                cp_async_16(dst, mad_wide(c0, TS * (uint32_t)sizeof(T), xpiece));
            if (p1)
                cp_async_16(dst + 512u, mad_wide(c1, TS * (uint32_t)sizeof(T), xpiece));
            dst += 1024u;
        }
    }
    else
    {
        const int piece = lane % PPT;
#pragma unroll 1
        for (int t = lane / PPT; t < ntiles; t += TPI, dst += 512u)
        {
            const int col0 = (int)list_u32<GLOBAL>(list_s, list_g, 4u * (uint32_t)t) * TS + piece * EPP;
            int valid = colA - col0;
            valid = valid < 0 ? 0 : (valid > EPP ? EPP : valid);
            cp_async_16_zfill(dst, valid ? x + col0 : x, (uint32_t)valid * (uint32_t)sizeof(T));
        }
    }
    // gathers of the extracted nonzeros' x values, 4 x 32 per trip so that the column loads and the
    // copies of a trip are all in flight together
    uint32_t dst2 = xb_s + (uint32_t)ntiles * (uint32_t)(TS * sizeof(T)) + (uint32_t)lane * (uint32_t)sizeof(T);
#pragma unroll 1
    for (int e0 = lane; e0 < nside + lane; e0 += 128, dst2 += 128u * (uint32_t)sizeof(T))
    {
        uint32_t c[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            c[j] = e0 + 32 * j < nside ? list_u32<GLOBAL>(list_s, list_g, sl + 4u * (uint32_t)(e0 - lane + 32 * j)) : 0u;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (e0 + 32 * j < nside)
            {
                const void *src = mad_wide(c[j], (uint32_t)sizeof(T), x);
                if (sizeof(T) == 8)
                    cp_async_8(dst2 + (uint32_t)j * 32u * (uint32_t)sizeof(T), src);
                else
                    cp_async_4(dst2 + (uint32_t)j * 32u * (uint32_t)sizeof(T), src);
            }
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
// PLAIN = the launch has no peers and does not accumulate (every single-GPU y = A*x): the epilogue is one select of the
// destination (y / partial-sum scratch) and one store on 32-bit row arithmetic instead of the general path's chain of
// warp-uniform branches (47 -> ~12 warp instructions per block row)
template <class T, bool PLAIN>
__device__ __forceinline__ void process_chunk(const unsigned char *st, uint32_t st_s, const T *xb, uint32_t xb_s,
                                              uint32_t zero_s, const SpmvArgs<T> &a, int lane)
{
    typedef typename Vec2<T>::type V2;
    constexpr uint32_t VS = (uint32_t)sizeof(T);
    const uint4 ha = lds_v4(st_s);
    const int nrows = (int)(ha.x & 0xffffu), ntiles = (int)(ha.x >> 16);
    const int p = lane & 7, g = lane >> 3;
    const uint2 *odesc = reinterpret_cast<const uint2 *>(st + (ha.z >> 16));
    uint32_t sidehdr_s = st_s + (ha.w & 0xffffu) + 4u * (uint32_t)p;
    const T *sideval = reinterpret_cast<const T *>(st + (ha.w >> 16)); // values of the extracted nonzeros ...
    const T *xside = xb + ntiles * TS;                                  // ... and their staged x operands
    uint32_t pay_s = st_s + lds_u32(st_s + 16u);
    uint32_t rows_s = st_s + CHUNK_OFF_ROWS;

#pragma unroll 1
    for (int rr = 0; rr < nrows; rr++, rows_s += 16u)
    {
        const uint4 rec = lds_v4(rows_s);
        const int nsr = (int)(rec.y & 0xffffu);
        const int nother = (int)(rec.y >> 16);
        const bool has_side = (rec.z & (ROWF_HAS_SIDE << 8)) != 0;
        uint32_t w0 = 0, w1 = 0, wt = 0;
        if (has_side) // issued early: independent of the ELL loop
        {
            w0 = lds_u32(sidehdr_s);
            w1 = lds_u32(sidehdr_s + 4u);
            wt = lds_u32(sidehdr_s + 32u - 4u * (uint32_t)p);
        }
        T a0 = 0, a1 = 0;

        // ---- ELL group: one flat loop over the slot-rows of all ELL tiles of the row, 4 slot-rows
        //      per lane and trip, every load of a trip in flight at once (predicated tail).
        //      Padding slots hold value 0 / column 0 and are multiplied through like in the
        //      reference GPU kernel (tilespmv_cuda.h:597-598); only tilespmv_cpu.h:182 skips them
        {
            uint32_t va = pay_s + (uint32_t)lane * (2u * VS);                  // V2 of (slot-row g, pair p)
            uint32_t ia = pay_s + (uint32_t)nsr * (TS * VS) + (uint32_t)lane;   // its nibble byte
            uint32_t sa = pay_s + (uint32_t)nsr * (TS * VS + 8u) + (uint32_t)g; // its x-segment selector
            int left = nsr;
            auto slot = [&](uint32_t j) { // slot-row g + 4*j of the current position, unpredicated
                const V2 v = SL<T>::ld2(va + j * (64u * VS));
                const uint32_t b = lds_u8(ia + 32u * j);
                const uint32_t xo = xb_s + lds_u8(sa + 4u * j) * (TS * VS);
                a0 = fma_t<T>(v.x, SL<T>::ld(mad_u32(b >> 4, VS, xo)), a0);
                a1 = fma_t<T>(v.y, SL<T>::ld(mad_u32(b & 15u, VS, xo)), a1);
            };
            auto slot_if = [&](uint32_t j, bool ok) { // predicated-off lanes read the CTA's zero block
                const V2 v = SL<T>::ld2(ok ? va + j * (64u * VS) : zero_s);
                const uint32_t b = lds_u8(ia + 32u * j);
                const uint32_t xo = xb_s + lds_u8(sa + 4u * j) * (TS * VS);
                a0 = fma_t<T>(v.x, SL<T>::ld(ok ? mad_u32(b >> 4, VS, xo) : zero_s), a0);
                a1 = fma_t<T>(v.y, SL<T>::ld(ok ? mad_u32(b & 15u, VS, xo) : zero_s), a1);
            };
#pragma unroll 1
            for (; left >= 16; left -= 16) // 16 slot-rows: every lane has 4, all loads of the trip in flight
            {
                slot(0);
                slot(1);
                slot(2);
                slot(3);
                va += 256u * VS;
                ia += 128u;
                sa += 16u;
            }
            if (left >= 8)
            {
                slot(0);
                slot(1);
                va += 128u * VS;
                ia += 64u;
                sa += 8u;
                left -= 8;
            }
            if (left > 0) // 1..7 slot-rows left
            {
                if (left > 4)
                {
                    slot(0);
                    slot_if(1, g + 4 < left);
                }
                else
                    slot_if(0, g < left);
            }
            pay_s += (rec.w & 0xffffu) * 16u;
        }
        const unsigned char *pay = st + (pay_s - st_s);

        // ---- the other tiles of the row ----
#pragma unroll 1
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
                const uint32_t vbytes = pad8((uint32_t)nnz * (uint32_t)sizeof(T));
                // rows 2p / 2p+1 = entries [k0, s1) / [s1, e1); the 4 lanes of a row pair take every 4th
                // entry.  w = max over the 16 rows of ceil(row length / 4): one loop for both rows, loads
                // unconditional (a lane past its row's end reads neighbouring bytes of the chunk that
                // are discarded), only the FMA is predicated.
                const uint32_t pay_a = st_s + (uint32_t)(pay - st);
                const uint32_t s1 = lds_u8(pay_a + 2u * p + 1u);
                const uint32_t e1 = p == 7 ? (uint32_t)nnz : lds_u8(pay_a + 2u * p + 2u);
                uint32_t k0 = lds_u8(pay_a + 2u * p) + (uint32_t)g, k1 = s1 + (uint32_t)g;
                const uint32_t cv_a = pay_a + 16u, ix_a = cv_a + vbytes;
                const uint32_t xs_a = xb_s + ((d.x >> 8) & 0xffu) * (TS * VS);
#pragma unroll 1
                for (int i = 0; i < w; i += 2)
                {
                    // parity of k0 / k1 is the same for both trips of the unrolled body (stride 4)
                    const uint32_t sh0 = (k0 & 1u) ? 0u : 4u, sh1 = (k1 & 1u) ? 0u : 4u;
                    const uint32_t va0 = cv_a + k0 * VS, va1 = cv_a + k1 * VS;
                    const uint32_t ia0 = ix_a + (k0 >> 1), ia1 = ix_a + (k1 >> 1);
#pragma unroll
                    for (uint32_t j = 0; j < 2; j++)
                    {
                        const T v0 = SL<T>::ld(va0 + j * (4u * VS)), v1 = SL<T>::ld(va1 + j * (4u * VS));
                        const uint32_t n0 = (lds_u8(ia0 + 2u * j) >> sh0) & 15u, n1 = (lds_u8(ia1 + 2u * j) >> sh1) & 15u;
                        const T x0 = SL<T>::ld(mad_u32(n0, VS, xs_a)), x1 = SL<T>::ld(mad_u32(n1, VS, xs_a));
                        if (k0 + 4u * j < s1)
                            a0 = fma_t<T>(v0, x0, a0);
                        if (k1 + 4u * j < e1)
                            a1 = fma_t<T>(v1, x1, a1);
                    }
                    k0 += 8u;
                    k1 += 8u;
                }
                pay += pad16(16u + vbytes + pad8(((uint32_t)nnz + 1u) / 2u));
                break;
            }
            case TSP_FMT_CSRGROUP:
            {
                // all CSR tiles of the block row as one jagged list (stream.cuh): slot-row s = the s-th entry of
                // every local row that has one.  Lane group g takes slot-rows g, g+4, ...; lane p finds the entries
                // of rows 2p / 2p+1 from the slot-row's 16-bit row mask (popc of the bits below 2p) and its start
                // offset.  Loads are unconditional (position <= n is always inside the payload, idx[n] = 0), only
                // the FMAs are predicated; the trip count is warp-uniform.
                const uint32_t nsrg = (uint32_t)w, n = d.y & 0xffffu, nfull = d.y >> 16;
                const uint32_t hdr_a = st_s + (uint32_t)(pay - st) + 4u * (uint32_t)g;
                const uint32_t val_a = hdr_a - 4u * (uint32_t)g + pad16(4u * nsrg);
                const uint32_t idx_a = val_a + pad16(n * VS);
                const uint32_t xs_a = xb_s + ((d.x >> 8) & 0xffu) * (TS * VS);
                // ---- leading slot-rows with all 16 rows present: entry (s, r) is at position 16 s + r, so lane
                //      (g, p) reads rows 2p / 2p+1 of slot-row s0 + g with one 128-bit load + one 16-bit index load
                {
                    uint32_t va = val_a + (uint32_t)lane * (2u * VS);
                    uint32_t ia = idx_a + 2u * (uint32_t)lane;
                    int left = (int)nfull;
                    auto slot = [&](uint32_t j) {
                        const V2 v = SL<T>::ld2(va + j * (64u * VS));
                        const uint32_t u = lds_u16(ia + 64u * j);
                        a0 = fma_t<T>(v.x, SL<T>::ld(mad_u32(u & 0xffu, VS, xs_a)), a0);
                        a1 = fma_t<T>(v.y, SL<T>::ld(mad_u32(u >> 8, VS, xs_a)), a1);
                    };
                    auto slot_if = [&](uint32_t j, bool ok) { // predicated-off lanes read the CTA's zero block
                        const V2 v = SL<T>::ld2(ok ? va + j * (64u * VS) : zero_s);
                        const uint32_t u = lds_u16(ok ? ia + 64u * j : zero_s);
                        a0 = fma_t<T>(v.x, SL<T>::ld(ok ? mad_u32(u & 0xffu, VS, xs_a) : zero_s), a0);
                        a1 = fma_t<T>(v.y, SL<T>::ld(ok ? mad_u32(u >> 8, VS, xs_a) : zero_s), a1);
                    };
#pragma unroll 1
                    for (; left >= 16; left -= 16)
                    {
                        slot(0);
                        slot(1);
                        slot(2);
                        slot(3);
                        va += 256u * VS;
                        ia += 256u;
                    }
                    if (left >= 8)
                    {
                        slot(0);
                        slot(1);
                        va += 128u * VS;
                        ia += 128u;
                        left -= 8;
                    }
                    if (left > 0)
                    {
                        if (left > 4)
                        {
                            slot(0);
                            slot_if(1, g + 4 < left);
                        }
                        else
                            slot_if(0, g < left);
                    }
                }
                // ---- ragged rest: positions from the slot-row's row mask and start offset
                const uint32_t sh = 2u * (uint32_t)p, below = (1u << sh) - 1u;
#pragma unroll 1
                for (uint32_t s0 = nfull; s0 < nsrg; s0 += 8u)
                {
                    const uint32_t h0 = s0 + (uint32_t)g < nsrg ? lds_u32(hdr_a + 4u * s0) : 0u;
                    const uint32_t h1 = s0 + 4u + (uint32_t)g < nsrg ? lds_u32(hdr_a + 4u * s0 + 16u) : 0u;
                    const uint32_t q00 = (h0 >> 16) + (uint32_t)__popc(h0 & below), q01 = q00 + ((h0 >> sh) & 1u);
                    const uint32_t q10 = (h1 >> 16) + (uint32_t)__popc(h1 & below), q11 = q10 + ((h1 >> sh) & 1u);
                    const T v00 = SL<T>::ld(mad_u32(q00, VS, val_a)), v01 = SL<T>::ld(mad_u32(q01, VS, val_a));
                    const T v10 = SL<T>::ld(mad_u32(q10, VS, val_a)), v11 = SL<T>::ld(mad_u32(q11, VS, val_a));
                    const uint32_t i00 = lds_u8(idx_a + q00), i01 = lds_u8(idx_a + q01);
                    const uint32_t i10 = lds_u8(idx_a + q10), i11 = lds_u8(idx_a + q11);
                    const T x00 = SL<T>::ld(mad_u32(i00, VS, xs_a)), x01 = SL<T>::ld(mad_u32(i01, VS, xs_a));
                    const T x10 = SL<T>::ld(mad_u32(i10, VS, xs_a)), x11 = SL<T>::ld(mad_u32(i11, VS, xs_a));
                    if ((h0 >> sh) & 1u)
                        a0 = fma_t<T>(v00, x00, a0);
                    if ((h0 >> sh) & 2u)
                        a1 = fma_t<T>(v01, x01, a1);
                    if ((h1 >> sh) & 1u)
                        a0 = fma_t<T>(v10, x10, a0);
                    if ((h1 >> sh) & 2u)
                        a1 = fma_t<T>(v11, x11, a1);
                }
                pay += csr_group_bytes(nsrg, n, VS);
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

        // ---- extracted very-sparse nonzeros of this block row: 4 lanes per row pair, one
        //      predicated loop with a warp-uniform trip count, 4 trips' loads in flight at once ----
        pay_s = st_s + (uint32_t)(pay - st);
        if (has_side)
        {
            const uint32_t s1 = w0 >> 16, e1 = w1 & 0xffffu;
            uint32_t k0 = (w0 & 0xffffu) + (uint32_t)g, k1 = s1 + (uint32_t)g;
            const int nit = (int)(rec.z >> 16);
            // rows with many entries (pieces of hub rows of power-law matrices) would keep 4 lanes busy for hundreds
            // of trips: the whole warp sums them instead (below), this loop skips them
            const uint32_t longmask = wt >> 16;
            if (longmask)
            {
                if ((longmask >> (2 * p)) & 1u)
                    k0 = s1;
                if ((longmask >> (2 * p + 1)) & 1u)
                    k1 = e1;
            }
#pragma unroll 1
            for (int i = 0; i < nit; i += 2) // trips past a lane's last entry are predicated off
            {
                if (k0 < s1)
                    a0 = fma_t<T>(sideval[k0], xside[k0], a0);
                if (k1 < e1)
                    a1 = fma_t<T>(sideval[k1], xside[k1], a1);
                if (k0 + 4u < s1)
                    a0 = fma_t<T>(sideval[k0 + 4u], xside[k0 + 4u], a0);
                if (k1 + 4u < e1)
                    a1 = fma_t<T>(sideval[k1 + 4u], xside[k1 + 4u], a1);
                k0 += 8u;
                k1 += 8u;
            }
            if (longmask) // warp-uniform
            {
                const uint32_t hdr0_s = sidehdr_s - 4u * (uint32_t)p;
                uint32_t lm = longmask;
#pragma unroll 1
                while (lm)
                {
                    const int r = __ffs((int)lm) - 1;
                    lm &= lm - 1u;
                    const uint32_t se = lds_u32(hdr0_s + 2u * (uint32_t)(r & ~1));         // starts of rows r&~1, (r&~1)+1
                    const uint32_t nx = lds_u16(hdr0_s + 2u * (uint32_t)(r & ~1) + 4u);    // start of row (r&~1)+2
                    const uint32_t s = (r & 1) ? se >> 16 : se & 0xffffu, e = (r & 1) ? nx : se >> 16;
                    T c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 1
                    for (uint32_t q = s + (uint32_t)lane; q < e; q += 128u)
                    {
                        c0 = fma_t<T>(sideval[q], xside[q], c0);
                        if (q + 32u < e)
                            c1 = fma_t<T>(sideval[q + 32u], xside[q + 32u], c1);
                        if (q + 64u < e)
                            c2 = fma_t<T>(sideval[q + 64u], xside[q + 64u], c2);
                        if (q + 96u < e)
                            c3 = fma_t<T>(sideval[q + 96u], xside[q + 96u], c3);
                    }
                    T c = (c0 + c1) + (c2 + c3);
                    c += __shfl_xor_sync(0xffffffffu, c, 16);
                    c += __shfl_xor_sync(0xffffffffu, c, 8);
                    c += __shfl_xor_sync(0xffffffffu, c, 4);
                    c += __shfl_xor_sync(0xffffffffu, c, 2);
                    c += __shfl_xor_sync(0xffffffffu, c, 1);
                    if (lane == (r >> 1)) // lane (g = 0, p = r / 2) carries the sum into the row's accumulator
                    {
                        if (r & 1)
                            a1 += c;
                        else
                            a0 += c;
                    }
                }
            }
            const uint32_t total = wt & 0xffffu;
            sideval += total;
            xside += total;
            sidehdr_s += SIDEHDR_BYTES;
        }

        // ---- combine the 4 lane groups: after the exchange lanes with even g hold row 2p, lanes
        //      with odd g hold row 2p+1; lanes 0..15 write the 16 y values as one 128-byte store ----
        {
            const bool odd = (g & 1) != 0;
            const T send = odd ? a0 : a1;
            T acc = odd ? a1 : a0;
            acc += __shfl_xor_sync(0xffffffffu, send, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            const int rowlen = (int)(rec.z & 0xffu);
            const int r = 2 * p + (g & 1);
            if (PLAIN)
            {
                if (lane < 16 && r < rowlen)
                {
                    T *base = (rec.x & ROW_PARTIAL) ? a.scratch : a.y;
                    base[(rec.x & ~ROW_PARTIAL) * (uint32_t)TS + (uint32_t)r] = acc; // rows / slots < 2^32 / 16 (rowA is an int)
                }
            }
            else if (lane < 16 && r < rowlen)
            {
                const bool partial = (rec.x & ROW_PARTIAL) != 0;
                const size_t row = (size_t)(rec.x & ~ROW_PARTIAL) * TS + r;
                if (a.accumulate && !partial && a.npeers == 0)
                {
                    // column-panel sub-plan: exactly one add per row and launch, launches are ordered by the stream,
                    // so the sum is deterministic; the reduction is fire-and-forget (RED), no load to wait for
                    atomicAdd(a.y + row, acc);
                }
                else
                {
                    if (a.accumulate && !partial)
                        acc += a.y[row];
                    (partial ? a.scratch : a.y)[row] = acc;
                    if (!partial)
                        for (int q = 0; q < a.npeers; q++) // fused all-gather: next x of every peer that wants this row
                            if ((long long)row >= a.peer_lo[q] && (long long)row < a.peer_hi[q])
                                a.peers[q][a.row_offset + (long long)row] = acc;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// one FLAT chunk (stream.cuh): block rows that hold only extracted (side) entries.  32 local rows per round, lane =
// row; inside a round the entries are stored slot-major without padding, so slot j of every row that has one is a
// contiguous run: one ballot + popc gives the lane its position, the value and staged-x loads of a slot are
// conflict-free, every lane sums its own row in input order and the round's 32 y values leave as two 128-byte stores.
// ~30 + 12 * (longest row) warp instructions per round instead of ~450 per block row in process_chunk.
// ---------------------------------------------------------------------------------------------
template <class T, bool PLAIN>
__device__ __forceinline__ void store_row(const SpmvArgs<T> &a, uint32_t dest, int r, T acc)
{
    if (PLAIN)
    {
        T *base = (dest & ROW_PARTIAL) ? a.scratch : a.y;
        base[(dest & ~ROW_PARTIAL) * (uint32_t)TS + (uint32_t)r] = acc;
        return;
    }
    const bool partial = (dest & ROW_PARTIAL) != 0;
    const size_t row = (size_t)(dest & ~ROW_PARTIAL) * TS + r;
    if (a.accumulate && !partial && a.npeers == 0)
        atomicAdd(a.y + row, acc); // one add per row and launch, launches ordered by the stream: deterministic
    else
    {
        if (a.accumulate && !partial)
            acc += a.y[row];
        (partial ? a.scratch : a.y)[row] = acc;
        if (!partial)
            for (int q = 0; q < a.npeers; q++)
                if ((long long)row >= a.peer_lo[q] && (long long)row < a.peer_hi[q])
                    a.peers[q][a.row_offset + (long long)row] = acc;
    }
}

// SPLIT = false: 32 local rows per round, one lane per row.  SPLIT = true (chunks of ONE block row: 16 rows with many
// entries each, e.g. 20 per row in a uniform random matrix): two lanes per row -- lane (row, h) takes the slots j with
// j mod 2 == h -- so that all 32 lanes work, and the two halves are added with one shuffle.
template <class T, bool SPLIT, bool PLAIN>
__device__ __forceinline__ void process_chunk_flat_t(uint32_t st_s, uint32_t xb_s, const SpmvArgs<T> &a, int lane)
{
    constexpr uint32_t VS = (uint32_t)sizeof(T);
    constexpr int RPR = SPLIT ? 16 : 32; // rows per round
    const uint4 ha = lds_v4(st_s);
    const int nrows16 = (int)(ha.x & 0x7fffu) * TS;
    uint32_t nlong = ha.z >> 16;                       // off_odesc: number of FlatLong records
    const uint32_t lens_s = st_s + (ha.w & 0xffffu);   // off_sidehdr: len[16 nrows]
    const uint32_t vals_s = st_s + (ha.w >> 16);       // off_sideval: val[nside]
    uint32_t long_s = st_s + lds_u32(st_s + 16u);      // off_payload: FlatLong[]
    const int rl = SPLIT ? (lane & 15) : lane;         // row of the round this lane works on
    const uint32_t h = SPLIT ? (uint32_t)(lane >> 4) : 0u;
    const uint32_t lt = (1u << rl) - 1u, rowbits = SPLIT ? 0xffffu : 0xffffffffu;
    uint32_t pos = 0;
#pragma unroll 1
    for (int rd = 0; rd * RPR < nrows16; rd++)
    {
        const int rho = rd * RPR + rl;
        const bool live = rho < nrows16;
        const uint32_t len = live ? lds_u8(lens_s + (uint32_t)rho) : 0u;
        const uint4 rec = lds_v4(st_s + CHUNK_OFF_ROWS + 16u * (uint32_t)((live ? rho : nrows16 - 1) >> 4));
        T acc = 0;
        if (!SPLIT)
        {
#pragma unroll 1
            for (uint32_t j = 0;; j += 2u) // two slots per trip: four loads in flight
            {
                const unsigned m0 = __ballot_sync(0xffffffffu, len > j);
                if (!m0)
                    break;
                const unsigned m1 = __ballot_sync(0xffffffffu, len > j + 1u);
                const uint32_t n0 = (uint32_t)__popc(m0);
                const uint32_t p0 = pos + (uint32_t)__popc(m0 & lt), p1 = pos + n0 + (uint32_t)__popc(m1 & lt);
                T v0 = 0, x0 = 0, v1 = 0, x1 = 0;
                if (len > j)
                {
                    v0 = SL<T>::ld(mad_u32(p0, VS, vals_s));
                    x0 = SL<T>::ld(mad_u32(p0, VS, xb_s));
                }
                if (len > j + 1u)
                {
                    v1 = SL<T>::ld(mad_u32(p1, VS, vals_s));
                    x1 = SL<T>::ld(mad_u32(p1, VS, xb_s));
                }
                acc = fma_t<T>(v0, x0, acc);
                acc = fma_t<T>(v1, x1, acc);
                pos += n0 + (uint32_t)__popc(m1);
            }
        }
        else
        {
#pragma unroll 1
            for (uint32_t j = 0;; j += 4u) // four slots per trip, two per half: four loads in flight per lane
            {
                const unsigned m0 = __ballot_sync(0xffffffffu, len > j) & rowbits;
                if (!m0)
                    break;
                const unsigned m1 = __ballot_sync(0xffffffffu, len > j + 1u) & rowbits;
                const unsigned m2 = __ballot_sync(0xffffffffu, len > j + 2u) & rowbits;
                const unsigned m3 = __ballot_sync(0xffffffffu, len > j + 3u) & rowbits;
                const uint32_t n0 = (uint32_t)__popc(m0), n1 = (uint32_t)__popc(m1), n2 = (uint32_t)__popc(m2);
                // this lane's two slots: j + h and j + 2 + h
                const unsigned ma = h ? m1 : m0, mb = h ? m3 : m2;
                const uint32_t pa = pos + (h ? n0 : 0u) + (uint32_t)__popc(ma & lt);
                const uint32_t pb = pos + n0 + n1 + (h ? n2 : 0u) + (uint32_t)__popc(mb & lt);
                T v0 = 0, x0 = 0, v1 = 0, x1 = 0;
                if ((ma >> rl) & 1u)
                {
                    v0 = SL<T>::ld(mad_u32(pa, VS, vals_s));
                    x0 = SL<T>::ld(mad_u32(pa, VS, xb_s));
                }
                if ((mb >> rl) & 1u)
                {
                    v1 = SL<T>::ld(mad_u32(pb, VS, vals_s));
                    x1 = SL<T>::ld(mad_u32(pb, VS, xb_s));
                }
                acc = fma_t<T>(v0, x0, acc);
                acc = fma_t<T>(v1, x1, acc);
                pos += n0 + n1 + n2 + (uint32_t)__popc(m3);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 16); // the two halves of every row
        }
        // rows of this round with >= FLAT_LONG_ROW entries (pieces of hub rows): the whole warp sums them
#pragma unroll 1
        while (nlong) // warp-uniform
        {
            const uint32_t w0 = lds_u32(long_s), cnt = lds_u16(long_s + 4u);
            const uint32_t lrow = w0 & 0xffffu, s = w0 >> 16;
            if (lrow >= (uint32_t)(rd + 1) * (uint32_t)RPR)
                break;
            T c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 1
            for (uint32_t q = (uint32_t)lane; q < cnt; q += 128u)
            {
                const uint32_t e = s + q;
                c0 = fma_t<T>(SL<T>::ld(mad_u32(e, VS, vals_s)), SL<T>::ld(mad_u32(e, VS, xb_s)), c0);
                if (q + 32u < cnt)
                    c1 = fma_t<T>(SL<T>::ld(mad_u32(e + 32u, VS, vals_s)), SL<T>::ld(mad_u32(e + 32u, VS, xb_s)), c1);
                if (q + 64u < cnt)
                    c2 = fma_t<T>(SL<T>::ld(mad_u32(e + 64u, VS, vals_s)), SL<T>::ld(mad_u32(e + 64u, VS, xb_s)), c2);
                if (q + 96u < cnt)
                    c3 = fma_t<T>(SL<T>::ld(mad_u32(e + 96u, VS, vals_s)), SL<T>::ld(mad_u32(e + 96u, VS, xb_s)), c3);
            }
            T c = (c0 + c1) + (c2 + c3);
            c += __shfl_xor_sync(0xffffffffu, c, 16);
            c += __shfl_xor_sync(0xffffffffu, c, 8);
            c += __shfl_xor_sync(0xffffffffu, c, 4);
            c += __shfl_xor_sync(0xffffffffu, c, 2);
            c += __shfl_xor_sync(0xffffffffu, c, 1);
            if ((uint32_t)lane == (lrow & (uint32_t)(RPR - 1)))
                acc += c;
            long_s += 8u;
            nlong--;
        }
        if (live && h == 0u && (rho & 15) < (int)(rec.z & 0xffu))
            store_row<T, PLAIN>(a, rec.x, rho & 15, acc);
    }
}

template <class T, bool PLAIN>
__device__ __forceinline__ void process_chunk_flat(uint32_t st_s, uint32_t xb_s, const SpmvArgs<T> &a, int lane)
{
    if ((lds_u16(st_s) & 0x7fffu) == 1u) // one block row: two lanes per row
        process_chunk_flat_t<T, true, PLAIN>(st_s, xb_s, a, lane);
    else
        process_chunk_flat_t<T, false, PLAIN>(st_s, xb_s, a, lane);
}

template <class T, int SPMV_STAGES, int MAXREG, bool PLAIN>
__global__ void __maxnreg__(MAXREG) tile_spmv_kernel(const SpmvArgs<T> a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps_cta = blockDim.x >> 5;
    const uint32_t cb = (uint32_t)a.stage_stride, xsb = (uint32_t)a.xstage_bytes;
    const uint32_t per_warp = SPMV_STAGES * cb + 2u * xsb;
    unsigned char *wbase = smem + SPMV_BAR_BYTES + (size_t)warp * per_warp;
    unsigned char *xbase = wbase + (size_t)SPMV_STAGES * cb;
    // per warp: SPMV_STAGES barriers for the chunk stream
    const uint32_t bar0 = smem_u32(smem) + (uint32_t)(warp * SPMV_BARS_PER_WARP * 8);

    // warp w of CTA b takes chunks gw, gw + nw, ...: neighbouring warps stream neighbouring chunks
    const unsigned gw = blockIdx.x * nwarps_cta + warp;
    const unsigned nw = gridDim.x * nwarps_cta;
    const int nk = gw < a.nchunks ? (int)((a.nchunks - gw + nw - 1) / nw) : 0;
    // 64 bytes of zeros behind the barriers: where predicated-off loads of the tail loops land
    const uint32_t zero_s = smem_u32(smem) + (uint32_t)(SPMV_BAR_BYTES - 64);
    // Programmatic dependent launch: the NEXT kernel of the stream (launched with the programmatic-serialization
    // attribute, plan_launch_one) may be scheduled as soon as every CTA of this grid has passed this point.  Its CTAs
    // become resident when ours exit, and everything they do before their own griddepcontrol.wait below -- barrier
    // set-up, the TMA fetches of their first chunks of the (immutable) packed stream, the head lists -- runs under this
    // grid's tail instead of after it.  A no-op when nothing depends on this launch.
    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x < 16)
        reinterpret_cast<uint32_t *>(smem + SPMV_BAR_BYTES - 64)[threadIdx.x] = 0u;
    __syncthreads();
    // every warp of the grid prefetches its share of the x window (4 KB pieces, fire and forget); x may be the previous
    // kernel's output, so only after the grid dependency is resolved
    auto prefetch_x_window = [&]() {
        if (a.pf_bytes && lane == 0)
        {
            const unsigned long long per = ((a.pf_bytes + nw - 1) / nw + 4095ull) & ~4095ull;
            unsigned long long o = (unsigned long long)gw * per;
            const unsigned long long end = o + per < a.pf_bytes ? o + per : a.pf_bytes;
            for (; o < end; o += 4096ull)
            {
                const unsigned n = (unsigned)((end - o < 4096ull ? end - o : 4096ull) & ~15ull);
                if (n)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.pf_base + o), "r"(n) : "memory");
            }
        }
    };
    if (nk == 0)
    {
        if (a.pf_bytes)
        {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            prefetch_x_window();
        }
        return;
    }

    if (lane == 0)
    {
#pragma unroll
        for (int i = 0; i < SPMV_STAGES; i++)
            mbar_init(bar0 + 8u * i, 1);
        fence_mbar_init();
    }
    __syncwarp();

    const uint32_t stage0 = smem_u32(wbase);
    // the packed stream is read exactly once per SpMV: fetch it with L2 evict_first so that it does not push the
    // re-used x lines out of L2 (uniform random 1 M x 6.25 M, x = 50 MB: 320 -> 184 us)
    const uint64_t stream_policy = l2_policy_evict_first();
    // the first SPMV_STAGES chunks are fetched from the descriptor table; every later fetch takes
    // its descriptor from the header of the chunk whose stage it re-uses (no global load in the loop)
    if (lane == 0)
    {
        for (int k = 0; k < SPMV_STAGES && k < nk; k++)
        {
            const uint2 d = a.chunk_desc[gw + (unsigned)k * nw];
            mbar_expect_tx(bar0 + 8u * k, d.y);
            tma_load_1d_hint(stage0 + (uint32_t)k * cb, a.stream + (size_t)d.x * 16u, d.y, bar0 + 8u * k, stream_policy);
        }
    }

    // per-lane source of the x pieces: x + (lane % pieces-per-segment) * elements-per-piece
    const T *xpiece = a.x + (lane % (TS * (int)sizeof(T) / 16)) * (16 / (int)sizeof(T));
    const uint32_t xb0 = smem_u32(xbase);

    // x operand of the warp's first chunk: its lists come from the head array
    {
        const uint32_t *h = reinterpret_cast<const uint32_t *>(a.head + (size_t)gw * (size_t)a.head_stride);
        const uint32_t cnt = __ldg(h), fl = __ldg(h + 1);
        // nothing above read x or wrote y / scratch (the stream, the chunk table and the head lists never change): from
        // here on the launch needs the results of everything before it in the stream
        asm volatile("griddepcontrol.wait;" ::: "memory");
        prefetch_x_window();
        stage_x<T, true>(0u, h + HEAD_HDR_BYTES / 4, (int)(cnt & 0xffffu), (int)(cnt >> 16), fl, xb0, a.x, xpiece, a.colA, lane);
        cp_async_commit();
    }

    // st / ph: stage and mbarrier phase of chunk k
    uint32_t st = 0, ph = 0;
#pragma unroll 1
    for (int k = 0; k < nk; k++)
    {
        const uint32_t st_s = stage0 + st * cb;
        const uint32_t xcur = xb0 + (uint32_t)(k & 1) * xsb;
        mbar_wait(bar0 + 8u * st, ph);
        const uint2 issue = make_uint2(lds_u32(st_s + 24u), lds_u32(st_s + 28u)); // what to fetch into this stage next
        if (k + 1 < nk) // stage the x operand of chunk k+1 from the lists chunk k carries
        {
            const uint4 ha = lds_v4(st_s);
            const uint32_t fl = lds_u32(st_s + 20u);
            stage_x<T, false>(st_s + (ha.z & 0xffffu), nullptr, (int)(ha.y & 0xffffu), (int)(ha.y >> 16), fl & 0xffffu,
                              xb0 + (uint32_t)((k + 1) & 1) * xsb, a.x, xpiece, a.colA, lane);
        }
        cp_async_commit();
        cp_async_wait<1>(); // this lane's x copies of chunk k have landed ...
        __syncwarp();       // ... and so have everybody else's
        if (lds_u16(st_s) & CHF_FLAT) // warp-uniform: a chunk of extracted entries only
            process_chunk_flat<T, PLAIN>(st_s, xcur, a, lane);
        else
            process_chunk<T, PLAIN>(wbase + (size_t)st * cb, st_s, reinterpret_cast<const T *>(xbase + (size_t)(k & 1) * xsb), xcur,
                             zero_s, a, lane);
        __syncwarp(); // all lanes are done reading stage st and x buffer k&1
        if (lane == 0 && issue.y != 0u)
        {
            // the reads above were consumed by the arithmetic before the y store was issued, so
            // the stage can be handed back to the async proxy (same hand-over as a consumer
            // release -> producer TMA in a warp-specialised pipeline)
            mbar_expect_tx(bar0 + 8u * st, issue.y);
            tma_load_1d_hint(st_s, a.stream + (size_t)issue.x * 16u, issue.y, bar0 + 8u * st, stream_policy);
        }
        if (++st == SPMV_STAGES)
        {
            st = 0;
            ph ^= 1u;
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
    // a link of the programmatic-launch chain like the SpMV kernel itself: releases the next launch of the stream at once
    // and waits for the kernel before it (whose partial sums it adds) -- no-ops when launched without the attribute
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
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
    if (a.accumulate)
        sum += y[row];
    y[row] = sum;
    for (int p = 0; p < npeers; p++)
        if ((long long)row >= a.peer_lo[p] && (long long)row < a.peer_hi[p])
            a.peers[p][row_offset + (long long)row] = sum;
}

// the same for rows cut into many pieces: one CTA per row, 16 slot lanes x 16 rows, fixed-order
// shared-memory tree (deterministic)
template <class T>
__global__ void __launch_bounds__(256)
    split_fixup_big_kernel(const int4 *__restrict__ tab, const T *__restrict__ scratch, T *__restrict__ y, int npeers,
                           long long row_offset, SpmvArgs<T> a)
{
    __shared__ T part[16][TS + 1];
    asm volatile("griddepcontrol.launch_dependents;"); // see split_fixup_kernel
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int4 e = tab[blockIdx.x]; // block row, first slot, #slots, rowlen
    const int r = threadIdx.x & 15, q = threadIdx.x >> 4;
    T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int k = q;
    for (; k + 48 < e.z; k += 64)
    {
        s0 += scratch[(size_t)(e.y + k) * TS + r];
        s1 += scratch[(size_t)(e.y + k + 16) * TS + r];
        s2 += scratch[(size_t)(e.y + k + 32) * TS + r];
        s3 += scratch[(size_t)(e.y + k + 48) * TS + r];
    }
    for (; k < e.z; k += 16)
        s0 += scratch[(size_t)(e.y + k) * TS + r];
    part[q][r] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (q == 0 && r < e.w)
    {
        T sum = 0;
#pragma unroll
        for (int i = 0; i < 16; i++)
            sum += part[i][r];
        const size_t row = (size_t)e.x * TS + r;
        if (a.accumulate)
            sum += y[row];
        y[row] = sum;
        for (int p = 0; p < npeers; p++)
            if ((long long)row >= a.peer_lo[p] && (long long)row < a.peer_hi[p])
                a.peers[p][row_offset + (long long)row] = sum;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t warp_smem_bytes(const tilespmv_plan *P)
{
    return (size_t)P->stages * P->stage_stride + 2 * (size_t)P->xstage_bytes;
}

template <class T, bool PLAIN>
static const void *kernel_for_t(int stages, int warps)
{
    if (stages == 2)
        return warps > 20 ? (const void *)tile_spmv_kernel<T, 2, SPMV_REGS_LO, PLAIN> : (const void *)tile_spmv_kernel<T, 2, SPMV_REGS_HI, PLAIN>;
    if (stages == 3)
        return (const void *)tile_spmv_kernel<T, 3, 128, PLAIN>;
    return (const void *)tile_spmv_kernel<T, 4, 128, PLAIN>;
}
// plain = no peers, no accumulation: the specialised epilogue
template <class T>
static const void *kernel_for(int stages, int warps, bool plain)
{
    return plain ? kernel_for_t<T, true>(stages, warps) : kernel_for_t<T, false>(stages, warps);
}

template <class T>
static int set_kernel_attrs(int stages, int warps, int smem_optin)
{
    // the attribute is per kernel, not per plan: always raise it to the device limit so that plans
    // with different shared-memory footprints can coexist in one process
    int carve = 0;
    if (const char *e = getenv("TILESPMV_CARVEOUT")) // experiments only
        carve = atoi(e);
    for (int plain = 0; plain < 2; plain++)
    {
        const void *fn = kernel_for<T>(stages, warps, plain != 0);
        TSP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
        // prefer L1: the driver still has to provide the dynamic shared memory a launch asks for, so streaming plans
        // (221 KB) get the 228 KB carve-out as before, while gather-bound plans (<= gather_smem_cap) leave >= 92 KB of L1
        TSP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    }
    return TILESPMV_OK;
}

// One persistent CTA per SM (ctas_per_sm can raise it); the CTA gets as many independent warps as
// its shared memory holds -- every warp owns `stages` chunk buffers of stage_stride bytes + 2
// x-staging buffers.  Needs P->nchunks and P->stage_stride; sets grid / block / smem and the
// lookahead distance nw = grid * warps that the packer bakes into the stream.
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
        P->stages = 2;
    if (P->ctas_per_sm > 8)
        P->ctas_per_sm = 8;
    size_t budget = ((size_t)smem_optin + 1024) / P->ctas_per_sm - 1024; // 1 KB per CTA is reserved
    if (P->gather_bound && P->max_warps <= 0 && budget > (size_t)P->gather_smem_cap)
        budget = (size_t)P->gather_smem_cap; // leave the rest of the SM's 256 KB to L1 (plan.cuh)
    const size_t per_warp = warp_smem_bytes(P);
    if (budget < SPMV_BAR_BYTES + per_warp)
    {
        set_error("plan: chunk_bytes/xstage_bytes need %zu B of shared memory per warp, only %zu available",
                  per_warp, budget - SPMV_BAR_BYTES);
        return TILESPMV_ERR_INVALID;
    }
    int warps = (int)((budget - SPMV_BAR_BYTES) / per_warp);
    // default for 2 stages: 20 warps on the 96-register kernel (measured faster than 21 on 80)
    int cap = spmv_max_warps(P->stages);
    if (P->stages == 2 && P->max_warps <= 0)
        cap = 20;
    if (P->max_warps > 0 && cap > P->max_warps)
        cap = P->max_warps;
    if (warps > cap / P->ctas_per_sm)
        warps = cap / P->ctas_per_sm;
    if (warps < 1)
        warps = 1;
    const size_t smem = SPMV_BAR_BYTES + (size_t)warps * per_warp;
    int grid = sms * P->ctas_per_sm;
    const long long ctas_needed = (P->nchunks + warps - 1) / warps;
    if (ctas_needed < grid)
        grid = (int)(ctas_needed > 0 ? ctas_needed : 1);
    P->sm_count = sms;
    P->grid = grid;
    P->block = warps * 32;
    P->smem = (int)smem;
    P->nw = (long long)grid * warps;
    if (P->precision == 8)
        TSP_TRY(set_kernel_attrs<double>(P->stages, warps, smem_optin));
    else
        TSP_TRY(set_kernel_attrs<float>(P->stages, warps, smem_optin));
    return TILESPMV_OK;
}

int spmv_set_attrs(tilespmv_plan *P)
{
    int dev = 0, smem_optin = 0;
    TSP_CUDA(cudaGetDevice(&dev));
    TSP_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (P->smem > smem_optin || P->block < 32 || P->block > 1024 || P->stages < 2 || P->stages > SPMV_MAX_STAGES)
    {
        set_error("plan: launch shape (%d threads, %d B of shared memory, %d stages) does not fit this GPU", P->block, P->smem, P->stages);
        return TILESPMV_ERR_UNSUPPORTED;
    }
    if (P->precision == 8)
        return set_kernel_attrs<double>(P->stages, P->block / 32, smem_optin);
    return set_kernel_attrs<float>(P->stages, P->block / 32, smem_optin);
}

template <class T>
static int plan_launch_one(tilespmv_plan *P, const T *x, T *y, cudaStream_t s, int npeers, void *const *peers, int64_t row_offset,
                           const int64_t *peer_lo, const int64_t *peer_hi)
{
    if (P->nchunks == 0)
        return TILESPMV_OK;
    SpmvArgs<T> a;
    a.stream = P->stream.as<unsigned char>();
    a.chunk_desc = P->chunk_desc.as<uint2>();
    a.nchunks = (unsigned)P->nchunks;
    a.x = x;
    a.y = y;
    a.scratch = P->scratch.as<T>();
    a.colA = P->colA;
    a.head = P->head.as<unsigned char>();
    a.head_stride = P->head_stride;
    a.stage_stride = P->stage_stride;
    a.xstage_bytes = P->xstage_bytes;
    a.npeers = npeers;
    a.pf_base = nullptr;
    a.pf_bytes = 0;
    if (P->gather_bound && P->xcol_hi > P->xcol_lo && !getenv("TILESPMV_NO_X_PREFETCH"))
    {
        // 16-byte aligned window [xcol_lo, xcol_hi) of x, at most 96 MB (what L2 can hold next to the stream)
        const unsigned long long lo = ((unsigned long long)P->xcol_lo * sizeof(T)) & ~15ull;
        const unsigned long long hi = ((unsigned long long)P->xcol_hi * sizeof(T)) & ~15ull;
        if (hi > lo && hi - lo <= (96ull << 20))
        {
            a.pf_base = reinterpret_cast<const unsigned char *>(x) + lo;
            a.pf_bytes = hi - lo;
        }
    }
    a.accumulate = P->accumulate ? 1 : 0;
    a.row_offset = row_offset;
    for (int p = 0; p < TSP_MAX_PEERS; p++)
    {
        a.peers[p] = p < npeers ? reinterpret_cast<T *>(peers[p]) : nullptr;
        a.peer_lo[p] = p < npeers ? peer_lo[p] : 0;
        a.peer_hi[p] = p < npeers ? peer_hi[p] : 0;
    }
    const int grid = P->grid; // fixed at plan time: the stream's lookahead lists depend on it
    static const bool no_plain = getenv("TILESPMV_NO_PLAIN_EPILOGUE") != nullptr; // A/B switch, read once
    const bool plain = npeers == 0 && !a.accumulate && !no_plain;
    bool pdl = false;
    {
        void *args[] = {(void *)&a};
        // programmatic dependent launch (see the top of the kernel): this launch may begin while the previous kernel of
        // the stream drains; the kernel itself waits (griddepcontrol.wait) before it touches x, y or the scratch.  Only a
        // kernel-to-kernel edge of the same stream is relaxed; copies, events and kernels launched without the attribute
        // (the flag kernels of comm.cu, the caller's own kernels) keep the full stream order.  The fix-up kernels below are
        // links of the same chain: released by this kernel's first instruction, they wait for it before they read a sum.
        // Stream captures keep the attribute (the graph gets a programmatic edge: tilespmv_plan_iterate 21.1 -> 19.8 us per
        // iteration on config 1).  A/B switches: TILESPMV_NO_PDL=1 (never), TILESPMV_NO_PDL_IN_GRAPHS=1 (not while capturing).
        static const int pdl_mode = getenv("TILESPMV_NO_PDL") ? 0 : (getenv("TILESPMV_NO_PDL_IN_GRAPHS") ? 1 : 2);
        pdl = pdl_mode != 0;
        if (pdl_mode == 1)
        {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
                pdl = false;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(P->block);
        cfg.dynamicSmemBytes = (size_t)P->smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
        cudaError_t err = cudaLaunchKernelExC(&cfg, kernel_for<T>(P->stages, P->block / 32, plain), args);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (err != cudaSuccess)
        {
            set_error("launch of tile_spmv_kernel failed: %s", cudaGetErrorString(err));
            return TILESPMV_ERR_CUDA;
        }
    }
    // the fix-up kernels are links of the same chain (TILESPMV_NO_PDL_FIXUP=1: plain launches, the A/B switch)
    static const bool chain_fixup = getenv("TILESPMV_NO_PDL_FIXUP") == nullptr;
    auto launch_fixup = [&](auto kernel, unsigned fgrid, unsigned fblock, auto... kargs) -> int {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(fgrid);
        cfg.blockDim = dim3(fblock);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (pdl && chain_fixup) ? 1 : 0;
        cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, kargs...);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (err != cudaSuccess)
        {
            set_error("launch of a split fix-up kernel failed: %s", cudaGetErrorString(err));
            return TILESPMV_ERR_CUDA;
        }
        return TILESPMV_OK;
    };
    if (P->nsplit > P->nsplit_small)
        TSP_TRY(launch_fixup(split_fixup_big_kernel<T>, (unsigned)(P->nsplit - P->nsplit_small), 256u,
                             (const int4 *)(P->split_tab.as<int4>() + P->nsplit_small), (const T *)P->scratch.as<T>(), y, npeers,
                             (long long)row_offset, a));
    if (P->nsplit_small > 0)
    {
        const long long threads = P->nsplit_small * TS;
        TSP_TRY(launch_fixup(split_fixup_kernel<T>, grid_for((size_t)threads, 128), 128u, (const int4 *)P->split_tab.as<int4>(),
                             (long long)P->nsplit_small, (const T *)P->scratch.as<T>(), y, npeers, (long long)row_offset, a));
    }
    return TILESPMV_OK;
}

// the plan itself writes y; its column-panel sub-plans (plan.cuh) then accumulate in a fixed order.  The fused
// peer stores ride on the LAST launch (which visits every row), when y is final.
template <class T>
static int plan_launch_t(tilespmv_plan *P, const T *x, T *y, cudaStream_t s)
{
    const bool alone = P->sub.empty();
    TSP_TRY(plan_launch_one<T>(P, x, y, s, alone ? P->npeers : 0, P->peers, P->row_offset, P->peer_lo, P->peer_hi));
    for (size_t i = 0; i < P->sub.size(); i++)
    {
        const bool last = i + 1 == P->sub.size();
        TSP_TRY(plan_launch_one<T>(P->sub[i], x, y, s, last ? P->npeers : 0, P->peers, P->row_offset, P->peer_lo, P->peer_hi));
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

int plan_launch_unit(tilespmv_plan *P, int unit, const void *d_x, void *d_y, cudaStream_t s, bool with_peers)
{
    if (unit < 0 || unit > (int)P->sub.size() || (reinterpret_cast<uintptr_t>(d_x) & 15u) || (reinterpret_cast<uintptr_t>(d_y) & 15u))
    {
        set_error("spmv: bad launch unit or unaligned x / y");
        return TILESPMV_ERR_INVALID;
    }
    tilespmv_plan *Q = unit == 0 ? P : P->sub[(size_t)unit - 1];
    const int np = with_peers ? P->npeers : 0;
    if (P->precision == 8)
        return plan_launch_one<double>(Q, static_cast<const double *>(d_x), static_cast<double *>(d_y), s, np, P->peers, P->row_offset, P->peer_lo,
                                       P->peer_hi);
    return plan_launch_one<float>(Q, static_cast<const float *>(d_x), static_cast<float *>(d_y), s, np, P->peers, P->row_offset, P->peer_lo, P->peer_hi);
}

} // namespace tsp

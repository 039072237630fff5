// membw.cu -- read-bandwidth probes for the roofline denominator of a READ-dominated stream:
//   (1) grid-stride LDG.128 sum   (2) per-warp TMA bulk ring (cp.async.bulk 4 KB, no compute)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/membw.cu -o tools/bin/membw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void ldg_sum(const uint4 *p, size_t n, unsigned long long *out)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (; i + 3 * stride < n; i += 4 * stride)
    {
        uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
        acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    for (; i < n; i += stride)
        acc += __ldcs(p + i).x;
    if (acc == 0x12345678u)
        *out = acc;
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int STAGES>
__global__ void tma_ring(const unsigned char *src, size_t nchunks, int chunk, unsigned long long *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwc = blockDim.x >> 5;
    unsigned char *base = smem + 1024 + (size_t)warp * STAGES * chunk;
    const uint32_t bar0 = s32(smem) + warp * STAGES * 8;
    const size_t gw = blockIdx.x * (size_t)nwc + warp, nw = (size_t)gridDim.x * nwc;
    if (lane == 0)
    {
        for (int i = 0; i < STAGES; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](size_t c, int st) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * st), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(base + (size_t)st * chunk)),
                     "l"(src + c * (size_t)chunk), "r"(chunk), "r"(bar0 + 8 * st)
                     : "memory");
    };
    size_t c = gw;
    if (lane == 0)
        for (int i = 0; i < STAGES && c + i * nw < nchunks; i++)
            issue(c + i * nw, i);
    unsigned acc = 0;
    int st = 0, ph = 0;
    for (; c < nchunks; c += nw)
    {
        uint32_t done;
        do
        {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(bar0 + 8 * st), "r"(ph)
                         : "memory");
        } while (!done);
        acc += *reinterpret_cast<unsigned *>(base + (size_t)st * chunk + lane * 4);
        __syncwarp();
        if (lane == 0 && c + STAGES * nw < nchunks)
            issue(c + STAGES * nw, st);
        if (++st == STAGES)
        {
            st = 0;
            ph ^= 1;
        }
    }
    if (acc == 0x12345678u)
        *out = acc;
}

template <class F>
static float time_ms(F f, int iters)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 3; i++)
        f();
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++)
        f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main()
{
    const size_t bytes = (size_t)1 << 30;
    unsigned char *d;
    unsigned long long *out;
    cudaMalloc(&d, bytes);
    cudaMalloc(&out, 8);
    cudaMemset(d, 1, bytes);
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int bpsm : {4, 8, 16})
    {
        float ms = time_ms([&] { ldg_sum<<<sms * bpsm, 256>>>((const uint4 *)d, bytes / 16, out); }, 20);
        printf("ldg128 grid=%dx256: %.1f us  %.0f GB/s\n", sms * bpsm, ms * 1e3, bytes / ms / 1e6);
    }
    cudaFuncSetAttribute(tma_ring<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tma_ring<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tma_ring<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int chunk : {2048, 4096, 8192, 16384})
        for (int stages : {2, 3, 4})
            for (int warps : {8, 16, 21, 27})
            {
                size_t smem = 1024 + (size_t)warps * stages * chunk;
                if (smem > 227 * 1024)
                    continue;
                size_t nch = bytes / chunk;
                float ms = stages == 2   ? time_ms([&] { tma_ring<2><<<sms, warps * 32, smem>>>(d, nch, chunk, out); }, 20)
                           : stages == 3 ? time_ms([&] { tma_ring<3><<<sms, warps * 32, smem>>>(d, nch, chunk, out); }, 20)
                                         : time_ms([&] { tma_ring<4><<<sms, warps * 32, smem>>>(d, nch, chunk, out); }, 20);
                printf("tma chunk=%5d stages=%d warps=%2d inflight/SM=%3zu KB: %.1f us  %.0f GB/s\n", chunk, stages, warps,
                       (size_t)warps * stages * chunk / 1024, ms * 1e3, bytes / ms / 1e6);
            }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}

// gather_probe.cu -- how fast can one B200 gather scattered 8-byte x values?  The bound of the extracted
// (side) part of the SpMV when every nonzero needs its own random x element (uniform random / R-MAT).
//   ldg    : ld.global.f64 per lane, 4 independent loads in flight per lane
//   ldgsts : cp.async 8 B per lane into shared memory (what the SpMV kernel does), 4 x 32 per commit group
//   bulk16 : cp.async.bulk 16 B per lane (TMA path, bypasses the L1 tag stage), mbarrier completion
// over x windows of 8 MB (L2-resident), 50 MB and 400 MB.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gather_probe.cu -o tools/bin/gather_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_ldg(const double *x, const uint32_t *idx, size_t n, double *out)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    double acc = 0;
    for (; i + 3 * stride < n; i += 4 * stride)
    {
        const uint32_t a = idx[i], b = idx[i + stride], c = idx[i + 2 * stride], d = idx[i + 3 * stride];
        acc += x[a] + x[b] + x[c] + x[d];
    }
    if (acc == 0.12345)
        *out = acc;
}

__global__ void k_ldgsts(const double *x, const uint32_t *idx, size_t n, double *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *buf = reinterpret_cast<double *>(smem) + warp * 256; // 2 x 128 slots
    const size_t gw = blockIdx.x * (size_t)(blockDim.x >> 5) + warp, nw = (size_t)gridDim.x * (blockDim.x >> 5);
    double acc = 0;
    int par = 0;
    for (size_t base = gw * 128; base + 128 <= n; base += nw * 128, par ^= 1)
    {
        uint32_t c[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            c[j] = idx[base + 32 * j + lane];
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s32(buf + par * 128 + 32 * j + lane)), "l"(x + c[j]) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        acc += buf[(par ^ 1) * 128 + lane]; // previous batch
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (acc == 0.12345)
        *out = acc;
}

__global__ void k_bulk16(const double *x, const uint32_t *idx, size_t n, double *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwc = blockDim.x >> 5;
    unsigned char *buf = smem + 1024 + (size_t)warp * 2 * 128 * 16; // 2 x 128 slots of 16 B
    const uint32_t bar0 = s32(smem) + warp * 16;
    const size_t gw = blockIdx.x * (size_t)nwc + warp, nw = (size_t)gridDim.x * nwc;
    if (lane == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    double acc = 0;
    int par = 0, ph[2] = {0, 0};
    bool first = true;
    for (size_t base = gw * 128; base + 128 <= n; base += nw * 128, par ^= 1)
    {
        uint32_t c[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            c[j] = idx[base + 32 * j + lane] & ~1u; // 16-byte aligned pair
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * par), "r"(128 * 16) : "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(
                             s32(buf + ((size_t)par * 128 + 32 * j + lane) * 16)),
                         "l"(x + c[j]), "r"(bar0 + 8 * par)
                         : "memory");
        if (!first)
        {
            const int q = par ^ 1;
            uint32_t done;
            do
            {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done)
                             : "r"(bar0 + 8 * q), "r"(ph[q])
                             : "memory");
            } while (!done);
            ph[q] ^= 1;
            acc += *reinterpret_cast<double *>(buf + ((size_t)q * 128 + lane) * 16);
        }
        first = false;
        __syncwarp();
    }
    if (acc == 0.12345)
        *out = acc;
}

template <class F>
static float time_ms(F f, int iters)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 2; i++)
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

__global__ void fill_idx(uint32_t *idx, size_t n, uint32_t window)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    uint64_t z = i * 0x9E3779B97F4A7C15ull + 0x1234567ull; // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    idx[i] = (uint32_t)(z % window);
}

int main()
{
    const size_t n = (size_t)1 << 25; // 32 M gathers
    const size_t xmax = (size_t)50 * 1000 * 1000;
    double *x, *out;
    uint32_t *idx;
    cudaMalloc(&x, xmax * 8);
    cudaMalloc(&out, 8);
    cudaMalloc(&idx, n * 4);
    cudaMemset(x, 0, xmax * 8);
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_bulk16, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (size_t window : {(size_t)1 << 20, (size_t)6250000, xmax})
    {
        fill_idx<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, (uint32_t)window);
        for (int warps : {8, 16, 20, 32})
        {
            float a = time_ms([&] { k_ldg<<<sms * 2, warps * 32>>>(x, idx, n, out); }, 5);
            float b = time_ms([&] { k_ldgsts<<<sms, warps * 32, warps * 2048>>>(x, idx, n, out); }, 5);
            float c = time_ms([&] { k_bulk16<<<sms, warps * 32, 1024 + warps * 4096>>>(x, idx, n, out); }, 5);
            printf("window %4zu MB warps/CTA %2d: ldg %.1f us %.1f G/s | ldgsts %.1f us %.1f G/s | bulk16 %.1f us %.1f G/s\n",
                   window * 8 >> 20, warps, a * 1e3, n / a / 1e6, b * 1e3, n / b / 1e6, c * 1e3, n / c / 1e6);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}

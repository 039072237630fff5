// common.cuh -- shared host/device helpers of libtilespmv_b200 (sm_100a only).
#pragma once
#include <exception>
#include <new>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tilespmv.h"

namespace tsp
{

constexpr int TS = TILESPMV_BLOCK_SIZE; // 16 x 16 tiles (common.h:37-39 of the reference)

// ---------------------------------------------------------------------------------------------
// error reporting: thread-local message + status codes; no exceptions cross the C-ABI
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *last_error();
void clear_error(); // every extern "C" entry starts with a clean error string (a stale message must not outlive a later success)
extern std::atomic<int64_t> g_launches;

struct Status
{
    int code = TILESPMV_OK;
    bool ok() const { return code == TILESPMV_OK; }
};

#define TSP_CUDA(call)                                                                          \
    do                                                                                          \
    {                                                                                           \
        cudaError_t err__ = (call);                                                             \
        if (err__ != cudaSuccess)                                                               \
        {                                                                                       \
            ::tsp::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                 \
                             cudaGetErrorString(err__));                                        \
            return TILESPMV_ERR_CUDA;                                                           \
        }                                                                                       \
    } while (0)

#define TSP_TRY(expr)                                                                           \
    do                                                                                          \
    {                                                                                           \
        int rc__ = (expr);                                                                      \
        if (rc__ != TILESPMV_OK)                                                                \
            return rc__;                                                                        \
    } while (0)

// every kernel launch of the library goes through this macro: counts launches (the bench's
// gpu_launches claim) and surfaces launch-configuration errors immediately
#define TSP_LAUNCH(kernel, grid, block, smem, stream, ...)                                      \
    do                                                                                          \
    {                                                                                           \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                             \
        ::tsp::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
        cudaError_t err__ = cudaGetLastError();                                                 \
        if (err__ != cudaSuccess)                                                               \
        {                                                                                       \
            ::tsp::set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__, #kernel,     \
                             cudaGetErrorString(err__));                                        \
            return TILESPMV_ERR_CUDA;                                                           \
        }                                                                                       \
    } while (0)

// No C++ exception may cross the C-ABI: every extern "C" entry with a body of its own is a function-try-block that ends in
// one of these handlers (std::bad_alloc from a host vector sized by the matrix, std::length_error, ...).  A void entry
// reports through tilespmv_last_error() like its other failures.
#define TSP_CATCH_INT(name)                                                                     \
    catch (const std::bad_alloc &)                                                              \
    {                                                                                           \
        ::tsp::set_error("%s: out of host memory", name);                                       \
        return TILESPMV_ERR_ALLOC;                                                              \
    }                                                                                           \
    catch (const std::exception &e__)                                                           \
    {                                                                                           \
        ::tsp::set_error("%s: %s", name, e__.what());                                           \
        return TILESPMV_ERR_INVALID;                                                            \
    }                                                                                           \
    catch (...)                                                                                 \
    {                                                                                           \
        ::tsp::set_error("%s: unknown C++ exception", name);                                    \
        return TILESPMV_ERR_INVALID;                                                            \
    }
#define TSP_CATCH_VOID(name)                                                                    \
    catch (const std::exception &e__)                                                           \
    {                                                                                           \
        ::tsp::set_error("%s: %s", name, e__.what());                                           \
    }                                                                                           \
    catch (...)                                                                                 \
    {                                                                                           \
        ::tsp::set_error("%s: unknown C++ exception", name);                                    \
    }

// The product has no CPU fallback: every entry point that computes fails loudly without a GPU.
int require_device();

// ---------------------------------------------------------------------------------------------
// device memory: one owner object per allocation, sizes tracked for the *_info structs
// ---------------------------------------------------------------------------------------------
struct DevBuf
{
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes)
    {
        o.p = nullptr;
        o.bytes = 0;
    }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o)
        {
            release();
            p = o.p;
            bytes = o.bytes;
            o.p = nullptr;
            o.bytes = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    // allocates max(n,1) bytes (so empty arrays still have a valid pointer), optionally zeroed
    int alloc(size_t n, bool zero, cudaStream_t s = 0)
    {
        release();
        size_t want = n ? n : 16;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess)
        {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            return TILESPMV_ERR_ALLOC;
        }
        bytes = want;
        if (zero)
        {
            e = cudaMemsetAsync(p, 0, want, s);
            if (e != cudaSuccess)
            {
                set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
                return TILESPMV_ERR_CUDA;
            }
        }
        return TILESPMV_OK;
    }
    template <class T>
    T *as() const
    {
        return reinterpret_cast<T *>(p);
    }
};

inline unsigned grid_for(size_t n, unsigned block)
{
    size_t g = (n + block - 1) / block;
    if (g < 1)
        g = 1;
    if (g > 0x7fffffffull)
        g = 0x7fffffffull;
    return (unsigned)g;
}

inline int bits_for(uint64_t maxval) // number of bits needed to represent values 0..maxval
{
    int b = 0;
    while (maxval)
    {
        b++;
        maxval >>= 1;
    }
    return b;
}

} // namespace tsp

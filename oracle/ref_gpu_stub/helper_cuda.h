/* oracle/ref_gpu_stub/helper_cuda.h -- TEST/BENCH INFRASTRUCTURE ONLY.
 * The reference's vendored CSR5 includes the CUDA-samples headers <helper_cuda.h> and
 * <helper_functions.h> (src/external/CSR5_cuda/detail/cuda/common_cuda.h:5-6), which are not part
 * of the reference tree nor of this image.  The only symbol it uses from them is the error-check
 * macro; this stub supplies it so that the UNMODIFIED src/main.cu builds (oracle/Makefile, target
 * ref_gpu). */
#ifndef REF_GPU_STUB_HELPER_CUDA_H
#define REF_GPU_STUB_HELPER_CUDA_H
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define checkCudaErrors(call)                                                                         \
    do                                                                                                \
    {                                                                                                 \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
        {                                                                                             \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            exit(1);                                                                                  \
        }                                                                                             \
    } while (0)
#endif

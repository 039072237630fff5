/* empty: see helper_cuda.h */

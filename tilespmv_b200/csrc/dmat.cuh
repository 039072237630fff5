// dmat.cuh -- the device-resident Tile_matrix (tilespmv_dmat of include/tilespmv.h).
// Same arrays, names and lengths as the reference struct (format.h:3-56), held in
// device memory; produced by the GPU conversion (convert.cu) or uploaded from a host struct.
#pragma once
#include "common.cuh"

struct tilespmv_dmat
{
    int precision = TILESPMV_F64; // sizeof(value)
    int rowA = 0, colA = 0;
    int tilem = 0, tilen = 0, tilenum = 0;
    int64_t nnz = 0; // true nonzeros in rows < rowA

    int csrsize = 0, csrptrlen = 0, coosize = 0, ellsize = 0, hybsize = 0, hybellsize = 0, hybcoosize = 0;
    int dnssize = 0, dnsrowsize = 0, dnscolsize = 0, coototal = 0;
    int ndenserowid = 0, ndensecolid = 0;
    int64_t fmt_hist[7] = {0, 0, 0, 0, 0, 0, 0};

    tsp::DevBuf tile_ptr, tile_columnidx, tile_nnz, Format, blknnz, blknnznnz, dnsrowptr, dnscolptr, tilewidth;
    tsp::DevBuf csr_offset, csrptr_offset, coo_offset, ell_offset, hyb_offset, hyb_coocount, dns_offset,
        dnsrow_offset, dnscol_offset, new_coocount;
    tsp::DevBuf Blockcsr_Val, Blockcsr_Ptr, csr_compressedIdx;
    tsp::DevBuf Blockcoo_Val, coo_compressed_Idx;
    tsp::DevBuf Blockell_Val, ell_compressedIdx;
    tsp::DevBuf Blockhyb_Val, hybIdx;
    tsp::DevBuf Blockdense_Val;
    tsp::DevBuf Blockdenserow_Val, denserowid;
    tsp::DevBuf Blockdensecol_Val, densecolid;
    tsp::DevBuf deferredcoo_val, deferredcoo_colidx, deferredcoo_ptr;

    int64_t device_bytes() const
    {
        const tsp::DevBuf *all[] = {&tile_ptr, &tile_columnidx, &tile_nnz, &Format, &blknnz, &blknnznnz,
                                    &dnsrowptr, &dnscolptr, &tilewidth, &csr_offset, &csrptr_offset,
                                    &coo_offset, &ell_offset, &hyb_offset, &hyb_coocount, &dns_offset,
                                    &dnsrow_offset, &dnscol_offset, &new_coocount, &Blockcsr_Val,
                                    &Blockcsr_Ptr, &csr_compressedIdx, &Blockcoo_Val, &coo_compressed_Idx,
                                    &Blockell_Val, &ell_compressedIdx, &Blockhyb_Val, &hybIdx,
                                    &Blockdense_Val, &Blockdenserow_Val, &denserowid, &Blockdensecol_Val,
                                    &densecolid, &deferredcoo_val, &deferredcoo_colidx, &deferredcoo_ptr};
        int64_t b = 0;
        for (const tsp::DevBuf *d : all)
            b += (int64_t)d->bytes;
        return b;
    }
};

namespace tsp
{
// convert.cu
template <class T>
int convert_csr_to_tiles(int rowA, int colA, const int *d_rowptr, const int *d_colidx, const T *d_val,
                         tilespmv_dmat *out, cudaStream_t s, bool enable_hyb = false);
} // namespace tsp

// plan.cuh -- tilespmv_plan: packed stream + chunk schedule + launch configuration.
#pragma once
#include "dmat.cuh"
#include "stream.cuh"

constexpr int TSP_MAX_PEERS = 8;
constexpr int SPLIT_BIG_SLOTS = 8; // split rows with more pieces are combined by a whole CTA

struct tilespmv_plan
{
    int precision = TILESPMV_F64;
    int rowA = 0, colA = 0, tilem = 0;
    int64_t nnz = 0;

    // packed stream
    int chunk_bytes = 0, xstage_bytes = 0;
    int64_t nchunks = 0;
    int64_t stream_bytes = 0;
    tsp::DevBuf stream;    // the packed bytes
    tsp::DevBuf chunk_off; // uint64[nchunks+1], 16-byte aligned offsets into stream (planner / packer)
    tsp::DevBuf chunk_desc; // uint2[nchunks+1] {offset / 16, bytes}: what the kernel's TMA issue reads
    tsp::DevBuf head;       // x-staging lists of every warp's first chunk (stream.cuh), head_stride apart
    int head_stride = 0;
    int stage_stride = 0;   // bytes of one TMA stage in shared memory (>= the largest chunk)
    long long nw = 0;       // warps of the persistent grid = lookahead distance of the lists

    // block rows that were cut across chunks: partial sums land in scratch and are combined by the
    // fix-up kernel in a fixed order (deterministic, no atomics)
    int64_t nsplit = 0, nsplit_small = 0, nslots = 0; // rows with <= SPLIT_BIG_SLOTS pieces come first in split_tab
    tsp::DevBuf scratch;   // T[nslots*16]
    tsp::DevBuf split_tab; // int4 {block row, first slot, #slots, rowlen} per split row

    // launch configuration of the persistent kernel
    int grid = 0, block = 0, smem = 0, ctas_per_sm = 0, sm_count = 0, stages = 0, max_warps = 0;
    int flags = 0;            // TILESPMV_PLAN_*
    int format_mask = 0;      // formats this plan covers (bit f = TILESPMV_FMT_*, bit 1 = the extracted side entries); 0 = all
    // plans made mostly of extracted (side) entries are bound by scattered x gathers, and every gather in flight
    // pins a line of L1: such plans keep their shared memory under gather_smem_cap so that the SM's 256 KB are
    // carved into a larger L1 (uniform 1 M x 20: 176 us with 223 KB of shared memory, 122 us with 158 KB)
    bool gather_bound = false;
    int gather_smem_cap = 163 * 1024;
    int64_t csr_groups = 0;   // block rows whose CSR tiles were merged into a group

    // x panels: when x is much larger than L2 and most nonzeros are extracted (side) entries, the side matrix is
    // cut into column panels of xpanel_bytes of x; this plan covers the tiles + panel 0 and writes y, sub[p-1]
    // covers panel p and ACCUMULATES into y, launched in order, so that the random gathers of one launch stay
    // inside an L2-resident window of x
    int xpanel_bytes = 0;      // option: 0 = automatic, < 0 = off
    bool accumulate = false;   // this (sub-)plan adds to y instead of writing it
    bool keep_all_rows = true; // emit an item for every block row, also empty ones
    std::vector<tilespmv_plan *> sub;

    // the part of x this (sub-)plan's launch reads: columns [xcol_lo, xcol_hi)
    long long xcol_lo = 0, xcol_hi = 0;

    // roofline accounting
    int64_t b_alg = 0, b_csr = 0;

    // fused all-gather epilogue
    int npeers = 0;
    void *peers[TSP_MAX_PEERS] = {nullptr};
    int64_t row_offset = 0;
    int64_t peer_lo[TSP_MAX_PEERS] = {0}, peer_hi[TSP_MAX_PEERS] = {0}; // local rows peer q receives (set_peers: all)

    // device staging for the host-pointer entry point
    tsp::DevBuf hx, hy;
    // ... and for the pipelined batch entry point (tilespmv_plan_spmv_host_batch): a ring of x / y buffers,
    // one stream per direction + one for the kernel, events that hand the buffers from stage to stage
    static constexpr int HOST_RING = 3;
    tsp::DevBuf bx[HOST_RING], by[HOST_RING];
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[HOST_RING] = {nullptr}, ev_comp[HOST_RING] = {nullptr}, ev_out[HOST_RING] = {nullptr};
    bool ring_ready = false; // set only once every stream, event and buffer of the ring exists
    void release_host_ring()
    {
        for (int i = 0; i < HOST_RING; i++)
        {
            for (cudaEvent_t *e : {&ev_in[i], &ev_comp[i], &ev_out[i]})
                if (*e)
                {
                    cudaEventDestroy(*e);
                    *e = nullptr;
                }
            bx[i].release();
            by[i].release();
        }
        for (cudaStream_t *st : {&s_in, &s_comp, &s_out})
            if (*st)
            {
                cudaStreamDestroy(*st);
                *st = nullptr;
            }
        ring_ready = false;
    }
    // tilespmv_plan_iterate: the niters ping-pong launches captured once into a CUDA graph, re-instantiated only
    // when the buffers or the count change
    cudaStream_t s_capture = nullptr;
    cudaGraphExec_t iter_exec = nullptr;
    void *iter_xa = nullptr, *iter_xb = nullptr;
    int iter_n = 0;
    ~tilespmv_plan()
    {
        if (iter_exec)
            cudaGraphExecDestroy(iter_exec);
        if (s_capture)
            cudaStreamDestroy(s_capture);
        for (tilespmv_plan *q : sub)
            delete q;
        release_host_ring();
    }

    int64_t device_bytes() const
    {
        int64_t subs = 0;
        for (const tilespmv_plan *q : sub)
            subs += q->device_bytes();
        return subs + (int64_t)(stream.bytes + chunk_off.bytes + chunk_desc.bytes + head.bytes + scratch.bytes + split_tab.bytes + hx.bytes + hy.bytes + bx[0].bytes * HOST_RING +
                         by[0].bytes * HOST_RING);
    }
};

namespace tsp
{
// explicit column panels of the side matrix: ranges [cuts[k], cuts[k+1]); the launch order starts at range first_range
struct PanelSpec
{
    std::vector<long long> cuts;
    int first_range = 0;
};
int plan_build(const tilespmv_dmat *dm, const tilespmv_plan_options *opts, tilespmv_plan *plan, cudaStream_t s,
               const PanelSpec *spec = nullptr);
int plan_panel_bytes(const tilespmv_dmat *dm, int xpanel_bytes, long long *out);
int plan_launch(tilespmv_plan *plan, const void *d_x, void *d_y, cudaStream_t s);
// one launch unit of a plan: unit 0 = the plan itself (writes y), unit i = column-panel sub-plan i-1 (accumulates);
// fused peer stores ride on the unit flagged `with_peers`
int plan_launch_unit(tilespmv_plan *plan, int unit, const void *d_x, void *d_y, cudaStream_t s, bool with_peers);
int spmv_configure(tilespmv_plan *plan); // picks grid / smem, sets the kernel attributes
int spmv_set_attrs(tilespmv_plan *plan); // kernel attributes only (a plan loaded from a file keeps its launch shape)
int plan_save(const tilespmv_plan *plan, const char *path);
int plan_load(const char *path, tilespmv_plan **out);
} // namespace tsp

// comm.cu -- multi-GPU repeated SpMV x <- A*x behind the C-ABI.
//
// The reference is single-GPU (src/main.cu:74 selects one device); the north star adds row-block sharding over
// the GPUs of one box with x replicated by a per-iteration all-gather.  Everything that loop needs lives here, in
// the C library, so that a plain C host (one process per GPU) can run it:
//
//   tilespmv_comm   rendezvous of the ranks through a POSIX shared-memory segment (no MPI, no torch): host
//                   barrier, exchange of CUDA IPC handles, and -- on request -- an NCCL communicator whose unique
//                   id travels through the same segment
//   tilespmv_dist   one rank's shard: its plan, two replicated x buffers in ONE cudaMalloc block that every peer
//                   maps through CUDA IPC (NVLink P2P), and two rows of 32-bit flags in the same block
//
// Three exchanges of the y slices (= the slices of the next x):
//   NCCL       baseline: SpMV, then one in-place ncclAllGather (equal slices) or one grouped broadcast per rank
//   FUSED      the SpMV kernel's epilogue stores every y value into the next-x buffer of every peer (spmv.cu),
//              one flag barrier per iteration
//   PIPELINED  the copy engines push the slice to the peers in the order the peers need it, while every launch of
//              the next iteration waits only for the slices of x its columns read (column panels aligned with the
//              row blocks of the ranks, first the panel the rank owns): the exchange of iteration k runs under the
//              compute of iteration k+1 and costs no SM.
//
// Flags (uint32, monotonically increasing epochs, written with st.release.sys over NVLink, read by spin kernels
// with ld.acquire.sys and a time-out):  D[me][src] = e  <=> src's slice of the x consumed in epoch e has landed
// in my buffer;  A[me][dst] = e  <=> dst has finished epoch e - 1 (so the buffer epoch e's slice goes to is free).
// On every rank the operations of epoch e depend only on remote operations of epoch e - 1, so the protocol
// cannot deadlock even if all streams of a process were serialised.
#include <fcntl.h>
#include <nccl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <new>

#include "plan.cuh"

namespace tsp
{

constexpr uint32_t SHM_MAGIC = 0x43505354u; // "TSPC"
constexpr int COMM_MAX_RANKS = 16;
constexpr size_t DIST_FLAG_BYTES = 1024; // D[16] | A[16] | error word, then the two x buffers
constexpr size_t DIST_OFF_D = 0, DIST_OFF_A = 64, DIST_OFF_ERR = 128, DIST_OFF_EPOCH = 192;

struct ShmSlot
{
    cudaIpcMemHandle_t handle;
    int device;
    int pid;
    long long aux[4];
};
struct ShmSeg
{
    std::atomic<uint32_t> magic;
    uint32_t nranks;
    std::atomic<uint32_t> bar_count, bar_gen;
    std::atomic<uint32_t> abort_flag;
    char nccl_id[128];
    ShmSlot slot[COMM_MAX_RANKS];
};

static double now_s()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

} // namespace tsp

struct tilespmv_comm
{
    int rank = 0, nranks = 1, device = 0;
    std::string shm_name;
    tsp::ShmSeg *seg = nullptr;
    ncclComm_t nccl = nullptr;
    double timeout_s = 120.0;
};

struct tilespmv_dist
{
    tilespmv_comm *comm = nullptr;
    tilespmv_plan *plan = nullptr; // owned
    int rank = 0, nranks = 1, vs = 8;
    long long n = 0, r0 = 0, m_local = 0;
    std::vector<long long> cuts;
    bool equal_slices = false;
    tsp::DevBuf block;
    size_t x_off[2] = {0, 0}, xbytes = 0;
    unsigned char *peer_block[tsp::COMM_MAX_RANKS] = {nullptr};
    bool peer_mapped[tsp::COMM_MAX_RANKS] = {false};
    int peer_device[tsp::COMM_MAX_RANKS] = {0};
    cudaStream_t s_comm = nullptr, s_main = nullptr; // copy stream / the stream the loop itself runs on
    cudaStream_t s_peer[tsp::COMM_MAX_RANKS] = {nullptr}; // halo exchange: one copy stream per peer (the copies have no order)
    cudaEvent_t ev_peer[tsp::COMM_MAX_RANKS] = {nullptr};
    cudaEvent_t ev_user = nullptr, ev_done = nullptr; // hand-over between the caller's stream and s_main
    cudaEvent_t ev_fork = nullptr;                     // halo exchange: s_comm -> the per-peer copy streams
    // the enqueued work of one call (all iterations, both streams) captured into a CUDA graph, keyed by its shape
    struct GraphEntry
    {
        int exchange, niters, cur;
        cudaGraphExec_t exec;
        int64_t launches;
    };
    std::vector<GraphEntry> graphs;
    bool use_graph = true;
    cudaEvent_t ev_kernel[2] = {nullptr, nullptr}, ev_push[2] = {nullptr, nullptr};
    bool ev_push_valid[2] = {false, false};
    uint32_t epoch = 1; // next unused epoch (flags start at 0)
    int cur = 0;        // x buffer holding the current x
    std::vector<uint32_t> deps; // per launch unit: bit mask of the source ranks whose slices it reads (self excluded)
    // halo exchange: need[r] = the x columns rank r's launches read; rank p fuse-stores its rows inside need[q] to q and the
    // copy engines replicate the rest in the background.  Eligible when no rank sends more than half of its slice that way.
    std::vector<long long> need_lo, need_hi;
    bool halo_ok = false;
    uint32_t halo_in = 0, halo_out = 0; // peers whose rows I read / peers that read my rows
    unsigned long long spin_timeout_ns = 30ull * 1000000000ull;
    // TILESPMV_DIST_DEBUG=1: CUDA-event timing of the LAST push of every pipelined call, printed by tilespmv_dist_sync
    bool debug = false, dbg_armed = false;
    cudaEvent_t dbg0 = nullptr, dbg1 = nullptr;
    size_t dbg_bytes = 0;
    int dbg_dst = 0;
    ~tilespmv_dist()
    {
        for (int r = 0; r < tsp::COMM_MAX_RANKS; r++)
            if (peer_mapped[r])
                cudaIpcCloseMemHandle(peer_block[r]);
        for (GraphEntry &g : graphs)
            cudaGraphExecDestroy(g.exec);
        for (cudaEvent_t e : {ev_kernel[0], ev_kernel[1], ev_push[0], ev_push[1], ev_user, ev_done, ev_fork})
            if (e)
                cudaEventDestroy(e);
        for (cudaStream_t st : {s_comm, s_main})
            if (st)
                cudaStreamDestroy(st);
        for (int r = 0; r < tsp::COMM_MAX_RANKS; r++)
        {
            if (ev_peer[r])
                cudaEventDestroy(ev_peer[r]);
            if (s_peer[r])
                cudaStreamDestroy(s_peer[r]);
        }
        delete plan;
    }
};

namespace tsp
{

#define TSP_NCCL(call)                                                                          \
    do                                                                                          \
    {                                                                                           \
        ncclResult_t r__ = (call);                                                              \
        if (r__ != ncclSuccess)                                                                 \
        {                                                                                       \
            ::tsp::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                 \
                             ncclGetErrorString(r__));                                          \
            return TILESPMV_ERR_CUDA;                                                           \
        }                                                                                       \
    } while (0)

// ---------------------------------------------------------------------------------------------
// host barrier over the shared segment (sense by generation counter; time-out instead of a hang)
// ---------------------------------------------------------------------------------------------
static int comm_barrier(tilespmv_comm *c)
{
    if (c->nranks == 1)
        return TILESPMV_OK;
    ShmSeg *g = c->seg;
    const uint32_t gen = g->bar_gen.load(std::memory_order_acquire);
    if (g->bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)c->nranks)
    {
        g->bar_count.store(0, std::memory_order_relaxed);
        g->bar_gen.fetch_add(1, std::memory_order_acq_rel);
        return TILESPMV_OK;
    }
    const double t0 = now_s();
    int spins = 0;
    while (g->bar_gen.load(std::memory_order_acquire) == gen)
    {
        if (g->abort_flag.load(std::memory_order_relaxed))
        {
            set_error("comm: another rank aborted");
            return TILESPMV_ERR_CUDA;
        }
        if (++spins > 2000)
        {
            usleep(200);
            if (now_s() - t0 > c->timeout_s)
            {
                g->abort_flag.store(1);
                set_error("comm: barrier timed out after %.0f s (rank %d of %d)", c->timeout_s, c->rank, c->nranks);
                return TILESPMV_ERR_CUDA;
            }
        }
        else
            sched_yield();
    }
    return TILESPMV_OK;
}

static int comm_create(const char *name, int rank, int nranks, unsigned flags, tilespmv_comm **out)
{
    if (!name || !out || nranks < 1 || nranks > COMM_MAX_RANKS || rank < 0 || rank >= nranks)
    {
        set_error("comm_create: invalid argument (1 <= nranks <= %d)", COMM_MAX_RANKS);
        return TILESPMV_ERR_INVALID;
    }
    TSP_TRY(require_device());
    tilespmv_comm *c = new (std::nothrow) tilespmv_comm();
    if (!c)
        return TILESPMV_ERR_ALLOC;
    c->rank = rank;
    c->nranks = nranks;
    if (const char *e = getenv("TILESPMV_COMM_TIMEOUT_S"))
        c->timeout_s = atof(e) > 0 ? atof(e) : c->timeout_s;
    if (cudaGetDevice(&c->device) != cudaSuccess)
    {
        delete c;
        set_error("comm_create: cudaGetDevice failed");
        return TILESPMV_ERR_CUDA;
    }
    c->shm_name = std::string("/tilespmv_") + name;
    for (char &ch : c->shm_name)
        if (&ch != &c->shm_name[0] && ch == '/')
            ch = '_';
    auto fail = [&](int rc) {
        if (c->seg)
            munmap(c->seg, sizeof(ShmSeg));
        if (rank == 0)
            shm_unlink(c->shm_name.c_str());
        delete c;
        return rc;
    };
    int fd = -1;
    const double t0 = now_s();
    if (rank == 0)
    {
        shm_unlink(c->shm_name.c_str()); // a stale segment of a crashed run with the same name
        fd = shm_open(c->shm_name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)sizeof(ShmSeg)) != 0)
        {
            if (fd >= 0)
                close(fd);
            set_error("comm_create: cannot create shared memory %s", c->shm_name.c_str());
            return fail(TILESPMV_ERR_IO);
        }
    }
    else
    {
        for (;;)
        {
            fd = shm_open(c->shm_name.c_str(), O_RDWR, 0600);
            struct stat st;
            if (fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size >= sizeof(ShmSeg))
                break;
            if (fd >= 0)
                close(fd);
            if (now_s() - t0 > c->timeout_s)
            {
                set_error("comm_create: rank %d timed out waiting for %s", rank, c->shm_name.c_str());
                return fail(TILESPMV_ERR_IO);
            }
            usleep(1000);
        }
    }
    void *p = mmap(nullptr, sizeof(ShmSeg), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED)
    {
        set_error("comm_create: mmap of %s failed", c->shm_name.c_str());
        return fail(TILESPMV_ERR_IO);
    }
    c->seg = static_cast<ShmSeg *>(p);
    if (rank == 0)
    {
        memset(p, 0, sizeof(ShmSeg));
        c->seg->nranks = (uint32_t)nranks;
        if (flags & TILESPMV_COMM_NCCL)
        {
            ncclUniqueId id;
            static_assert(sizeof(ncclUniqueId) <= sizeof(ShmSeg::nccl_id), "ncclUniqueId does not fit");
            ncclResult_t r = ncclGetUniqueId(&id);
            if (r != ncclSuccess)
            {
                set_error("comm_create: ncclGetUniqueId failed: %s", ncclGetErrorString(r));
                return fail(TILESPMV_ERR_CUDA);
            }
            memcpy(c->seg->nccl_id, &id, sizeof(id));
        }
        c->seg->magic.store(SHM_MAGIC, std::memory_order_release);
    }
    else
    {
        while (c->seg->magic.load(std::memory_order_acquire) != SHM_MAGIC)
        {
            if (now_s() - t0 > c->timeout_s)
            {
                set_error("comm_create: rank %d timed out waiting for rank 0", rank);
                return fail(TILESPMV_ERR_IO);
            }
            usleep(200);
        }
        if (c->seg->nranks != (uint32_t)nranks)
        {
            set_error("comm_create: segment %s was created for %u ranks, not %d", c->shm_name.c_str(), c->seg->nranks, nranks);
            return fail(TILESPMV_ERR_INVALID);
        }
    }
    int rc = comm_barrier(c);
    if (rc != TILESPMV_OK)
        return fail(rc);
    if (flags & TILESPMV_COMM_NCCL)
    {
        ncclUniqueId id;
        memcpy(&id, c->seg->nccl_id, sizeof(id));
        ncclResult_t r = ncclCommInitRank(&c->nccl, nranks, id, rank);
        if (r != ncclSuccess)
        {
            c->nccl = nullptr;
            c->seg->abort_flag.store(1);
            set_error("comm_create: ncclCommInitRank failed: %s", ncclGetErrorString(r));
            return fail(TILESPMV_ERR_CUDA);
        }
        rc = comm_barrier(c);
        if (rc != TILESPMV_OK)
            return fail(rc);
    }
    *out = c;
    return TILESPMV_OK;
}

static void comm_destroy(tilespmv_comm *c)
{
    if (!c)
        return;
    if (c->nccl)
        ncclCommDestroy(c->nccl);
    if (c->seg)
    {
        comm_barrier(c); // nobody unmaps while a peer is still inside a collective step (best effort: times out)
        munmap(c->seg, sizeof(ShmSeg));
        if (c->rank == 0)
            shm_unlink(c->shm_name.c_str());
    }
    delete c;
}

// ---------------------------------------------------------------------------------------------
// flag kernels
// ---------------------------------------------------------------------------------------------
struct SignalArgs
{
    uint32_t *ptr[COMM_MAX_RANKS];
    int n;
    uint32_t delta;
    const uint32_t *base;
};

// All flag values are RELATIVE to the epoch of the running call, which lives in a device word (set once per call):
// the kernels below take (base pointer, delta).  The enqueued work of a call therefore does not depend on the absolute
// epoch and can be captured into a CUDA graph once and replayed for every later call of the same shape.
__global__ void __launch_bounds__(32) set_epoch_kernel(uint32_t *word, uint32_t value) { *word = value; }

// thread i stores *base + delta into remote (or local) flag ptr[i]; everything the stream did before is visible first
__global__ void __launch_bounds__(32) flag_signal_kernel(SignalArgs a)
{
    const int i = (int)threadIdx.x;
    if (i < a.n)
    {
        const uint32_t value = *a.base + a.delta;
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.ptr[i]), "r"(value) : "memory");
    }
}

// thread i (for every bit i of mask) spins until flags[i] >= *base + delta (wrap-safe); gives up after timeout_ns and
// records 1 + i in *err so that a dead peer becomes an error instead of a hung GPU
__global__ void __launch_bounds__(32)
    flag_wait_kernel(const uint32_t *flags, uint32_t mask, const uint32_t *base, uint32_t delta, uint32_t *err, unsigned long long timeout_ns)
{
    const int i = (int)threadIdx.x;
    if (!((mask >> i) & 1u))
        return;
    const uint32_t target = *base + delta;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned spins = 0;
    for (;;)
    {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
        if ((int32_t)(v - target) >= 0)
            break;
        __nanosleep(64);
        if ((++spins & 1023u) == 0)
        {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns)
            {
                atomicExch(err, 1u + (uint32_t)i);
                break;
            }
        }
    }
}

// wait until flags[off][r] >= (epoch of the call) + delta for every rank r of mask
static int flag_wait(tilespmv_dist *d, size_t off, uint32_t mask, uint32_t delta, cudaStream_t s)
{
    if (!mask)
        return TILESPMV_OK;
    unsigned char *b = d->block.as<unsigned char>();
    TSP_LAUNCH(flag_wait_kernel, 1, 32, 0, s, reinterpret_cast<const uint32_t *>(b + off), mask, reinterpret_cast<const uint32_t *>(b + DIST_OFF_EPOCH),
               delta, reinterpret_cast<uint32_t *>(b + DIST_OFF_ERR), d->spin_timeout_ns);
    return TILESPMV_OK;
}

// store (epoch of the call) + delta into flag row `off`, entry `d->rank`, on every rank of mask
static int flag_signal(tilespmv_dist *d, size_t off, uint32_t mask, uint32_t delta, cudaStream_t s)
{
    SignalArgs a;
    a.n = 0;
    a.delta = delta;
    a.base = reinterpret_cast<const uint32_t *>(d->block.as<unsigned char>() + DIST_OFF_EPOCH);
    for (int r = 0; r < d->nranks; r++)
        if ((mask >> r) & 1u)
            a.ptr[a.n++] = reinterpret_cast<uint32_t *>(d->peer_block[r] + off) + d->rank;
    if (a.n == 0)
        return TILESPMV_OK;
    TSP_LAUNCH(flag_signal_kernel, 1, 32, 0, s, a);
    return TILESPMV_OK;
}

static uint32_t all_peers_mask(const tilespmv_dist *d) { return ((1u << d->nranks) - 1u) & ~(1u << d->rank); }

static unsigned char *xbuf(const tilespmv_dist *d, int rank, int which) { return d->peer_block[rank] + d->x_off[which]; }

// ---------------------------------------------------------------------------------------------
// tilespmv_dist
// ---------------------------------------------------------------------------------------------
static int dist_create(tilespmv_comm *c, const tilespmv_dmat *dm, const int64_t *row_cuts, const tilespmv_plan_options *opts,
                       unsigned flags, tilespmv_dist **out)
{
    if (!c || !dm || !row_cuts || !out)
    {
        set_error("dist_create: null argument");
        return TILESPMV_ERR_INVALID;
    }
    const int R = c->nranks, me = c->rank;
    for (int r = 0; r < R; r++)
        if (row_cuts[r] > row_cuts[r + 1] || row_cuts[0] != 0 || (r > 0 && row_cuts[r] % TS != 0))
        {
            set_error("dist_create: row_cuts must ascend from 0 in multiples of %d", TS);
            return TILESPMV_ERR_INVALID;
        }
    if (row_cuts[R] != dm->colA || row_cuts[me + 1] - row_cuts[me] != dm->rowA)
    {
        set_error("dist_create: rank %d owns rows [%lld, %lld) of a square %lld x %lld matrix, but its shard is %d x %d", me,
                  (long long)row_cuts[me], (long long)row_cuts[me + 1], (long long)row_cuts[R], (long long)row_cuts[R], dm->rowA, dm->colA);
        return TILESPMV_ERR_INVALID;
    }
    if (R - 1 > TSP_MAX_PEERS)
    {
        set_error("dist_create: at most %d ranks", TSP_MAX_PEERS + 1);
        return TILESPMV_ERR_INVALID;
    }
    tilespmv_dist *d = new (std::nothrow) tilespmv_dist();
    if (!d)
        return TILESPMV_ERR_ALLOC;
    auto fail = [&](int rc) {
        c->seg ? c->seg->abort_flag.store(1) : (void)0;
        delete d;
        return rc;
    };
    d->comm = c;
    d->rank = me;
    d->nranks = R;
    d->vs = dm->precision;
    d->n = dm->colA;
    d->r0 = row_cuts[me];
    d->m_local = dm->rowA;
    d->cuts.assign(row_cuts, row_cuts + R + 1);
    d->equal_slices = true;
    for (int r = 0; r < R; r++)
        d->equal_slices = d->equal_slices && (row_cuts[r + 1] - row_cuts[r] == row_cuts[1] - row_cuts[0]);
    if (const char *e = getenv("TILESPMV_COMM_SPIN_TIMEOUT_S"))
        if (atof(e) > 0)
            d->spin_timeout_ns = (unsigned long long)(atof(e) * 1e9);

    // ---- the plan: when the side matrix wants x panels, cut them at the row blocks of the ranks and start with the
    //      panel this rank owns, so that launch u needs only the slice of rank (me + u) % R (pipelined exchange)
    d->plan = new (std::nothrow) tilespmv_plan();
    if (!d->plan)
        return fail(TILESPMV_ERR_ALLOC);
    long long panel_bytes = 0;
    int rc = plan_panel_bytes(dm, opts ? opts->xpanel_bytes : 0, &panel_bytes);
    if (rc != TILESPMV_OK)
        return fail(rc);
    PanelSpec spec;
    const bool rank_panels = panel_bytes > 0 && R > 1 && !(flags & TILESPMV_DIST_UNIFORM_PANELS);
    if (rank_panels)
    {
        spec.cuts.assign(row_cuts, row_cuts + R + 1);
        spec.first_range = me;
    }
    rc = plan_build(dm, opts, d->plan, 0, rank_panels ? &spec : nullptr);
    if (rc != TILESPMV_OK)
        return fail(rc);
    const int nunits = 1 + (int)d->plan->sub.size();
    d->deps.assign((size_t)nunits, 0u);
    for (int u = 0; u < nunits; u++)
    {
        const tilespmv_plan *Q = u == 0 ? d->plan : d->plan->sub[(size_t)u - 1];
        for (int r = 0; r < R; r++)
            if (r != me && row_cuts[r] < Q->xcol_hi && row_cuts[r + 1] > Q->xcol_lo)
                d->deps[(size_t)u] |= 1u << r;
    }

    d->need_lo.assign((size_t)R, 0);
    d->need_hi.assign((size_t)R, 0);
    {
        long long lo = d->n, hi = 0;
        for (int u = 0; u < nunits; u++)
        {
            const tilespmv_plan *Q = u == 0 ? d->plan : d->plan->sub[(size_t)u - 1];
            if (Q->xcol_hi > Q->xcol_lo)
            {
                lo = std::min(lo, Q->xcol_lo);
                hi = std::max(hi, Q->xcol_hi);
            }
        }
        d->need_lo[(size_t)me] = hi > lo ? lo : 0;
        d->need_hi[(size_t)me] = hi > lo ? hi : 0;
    }

    // ---- one block: flags | x buffer 0 | x buffer 1, mapped by every peer through CUDA IPC ----
    d->xbytes = ((size_t)d->n * (size_t)d->vs + 255u) & ~(size_t)255u;
    d->x_off[0] = DIST_FLAG_BYTES;
    d->x_off[1] = DIST_FLAG_BYTES + d->xbytes;
    rc = d->block.alloc(DIST_FLAG_BYTES + 2 * d->xbytes, false);
    if (rc != TILESPMV_OK)
        return fail(rc);
    if (cudaMemset(d->block.p, 0, DIST_FLAG_BYTES) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
    {
        set_error("dist_create: cudaMemset failed");
        return fail(TILESPMV_ERR_CUDA);
    }
    d->peer_block[me] = d->block.as<unsigned char>();
    d->peer_device[me] = c->device;
    if (R > 1)
    {
        ShmSlot &mine = c->seg->slot[me];
        if (cudaIpcGetMemHandle(&mine.handle, d->block.p) != cudaSuccess)
        {
            set_error("dist_create: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(TILESPMV_ERR_CUDA);
        }
        mine.device = c->device;
        mine.pid = (int)getpid();
        mine.aux[0] = (long long)(DIST_FLAG_BYTES + 2 * d->xbytes);
        mine.aux[1] = d->need_lo[(size_t)me];
        mine.aux[2] = d->need_hi[(size_t)me];
        rc = comm_barrier(c);
        if (rc != TILESPMV_OK)
            return fail(rc);
        for (int r = 0; r < R; r++)
        {
            if (r == me)
                continue;
            const ShmSlot &o = c->seg->slot[r];
            if (o.aux[0] != mine.aux[0])
            {
                set_error("dist_create: rank %d allocated %lld bytes, rank %d %lld: the shards disagree about the matrix size", r, o.aux[0], me,
                          mine.aux[0]);
                return fail(TILESPMV_ERR_INVALID);
            }
            if (o.pid == mine.pid)
            {
                set_error("dist_create: ranks %d and %d share a process; use one process per rank (CUDA IPC)", r, me);
                return fail(TILESPMV_ERR_UNSUPPORTED);
            }
            // peer access must be enabled explicitly for the COPY path: cudaIpcMemLazyEnablePeerAccess alone lets
            // kernels store through the mapping, but cudaMemcpyAsync between the two devices then stages through the
            // host (measured: 25-30 GB/s instead of NVLink speed)
            if (o.device != c->device)
            {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, c->device, o.device);
                cudaError_t pe = can ? cudaDeviceEnablePeerAccess(o.device, 0) : cudaErrorPeerAccessUnsupported;
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                {
                    cudaGetLastError();
                    set_error("dist_create: no peer access from device %d to device %d (%s)", c->device, o.device, cudaGetErrorString(pe));
                    return fail(TILESPMV_ERR_CUDA);
                }
                cudaGetLastError();
            }
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, o.handle, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
            {
                cudaGetLastError();
                set_error("dist_create: cudaIpcOpenMemHandle of rank %d (device %d) failed: %s", r, o.device, cudaGetErrorString(e));
                return fail(TILESPMV_ERR_CUDA);
            }
            d->peer_block[r] = static_cast<unsigned char *>(p);
            d->peer_mapped[r] = true;
            d->peer_device[r] = o.device;
            d->need_lo[(size_t)r] = o.aux[1];
            d->need_hi[(size_t)r] = o.aux[2];
        }
        rc = comm_barrier(c); // every rank has read every slot: the slots may be re-used by the next dist_create
        if (rc != TILESPMV_OK)
            return fail(rc);
    }
    // halo eligibility (the same verdict on every rank: it only uses the published ranges and the cuts)
    d->halo_ok = R > 1;
    for (int p = 0; p < R && d->halo_ok; p++)
    {
        long long sent = 0;
        for (int q = 0; q < R; q++)
            if (q != p)
                sent += std::max(0ll, std::min(d->need_hi[(size_t)q], (long long)row_cuts[p + 1]) - std::max(d->need_lo[(size_t)q], (long long)row_cuts[p]));
        if (2 * sent > (long long)(row_cuts[p + 1] - row_cuts[p]))
            d->halo_ok = false;
    }
    for (int q = 0; q < R; q++)
    {
        if (q == me)
            continue;
        if (std::min(d->need_hi[(size_t)q], (long long)row_cuts[me + 1]) > std::max(d->need_lo[(size_t)q], (long long)row_cuts[me]))
            d->halo_out |= 1u << q;
        if (std::min(d->need_hi[(size_t)me], (long long)row_cuts[q + 1]) > std::max(d->need_lo[(size_t)me], (long long)row_cuts[q]))
            d->halo_in |= 1u << q;
    }
    if (const char *e = getenv("TILESPMV_DIST_DEBUG"))
        d->debug = atoi(e) != 0;
    if (d->debug)
    {
        cudaEventCreate(&d->dbg0);
        cudaEventCreate(&d->dbg1);
    }
    if (const char *e = getenv("TILESPMV_DIST_NO_GRAPH"))
        d->use_graph = atoi(e) == 0;
    if (d->debug)
        d->use_graph = false; // the push timing events cannot be read back from a graph
    if (cudaStreamCreateWithFlags(&d->s_main, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&d->s_comm, cudaStreamNonBlocking) != cudaSuccess)
    {
        set_error("dist_create: cudaStreamCreate failed");
        return fail(TILESPMV_ERR_CUDA);
    }
    for (cudaEvent_t *e : {&d->ev_kernel[0], &d->ev_kernel[1], &d->ev_push[0], &d->ev_push[1], &d->ev_user, &d->ev_done, &d->ev_fork})
        if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess)
        {
            set_error("dist_create: cudaEventCreate failed");
            return fail(TILESPMV_ERR_CUDA);
        }
    for (int r = 0; r < R; r++)
        if (r != me && (cudaStreamCreateWithFlags(&d->s_peer[r], cudaStreamNonBlocking) != cudaSuccess ||
                        cudaEventCreateWithFlags(&d->ev_peer[r], cudaEventDisableTiming) != cudaSuccess))
        {
            set_error("dist_create: cudaStreamCreate failed");
            return fail(TILESPMV_ERR_CUDA);
        }
    *out = d;
    return TILESPMV_OK;
}

// every launch unit of the plan for one iteration: src -> dst + r0 (unit 0 writes, the others accumulate)
static int launch_units(tilespmv_dist *d, const unsigned char *src, unsigned char *dst, cudaStream_t s, bool with_peers)
{
    const int nunits = 1 + (int)d->plan->sub.size();
    for (int u = 0; u < nunits; u++)
        TSP_TRY(plan_launch_unit(d->plan, u, src, dst + (size_t)d->r0 * (size_t)d->vs, s, with_peers && u == nunits - 1));
    return TILESPMV_OK;
}

static int iterate_nccl(tilespmv_dist *d, int niters, cudaStream_t s)
{
    tilespmv_comm *c = d->comm;
    if (d->nranks > 1 && !c->nccl)
    {
        set_error("dist_iterate: the communicator was created without TILESPMV_COMM_NCCL");
        return TILESPMV_ERR_INVALID;
    }
    const ncclDataType_t dt = d->vs == 8 ? ncclDouble : ncclFloat;
    for (int i = 0; i < niters; i++)
    {
        const unsigned char *src = xbuf(d, d->rank, (d->cur + i) & 1);
        unsigned char *dst = xbuf(d, d->rank, (d->cur + i + 1) & 1);
        TSP_TRY(launch_units(d, src, dst, s, false));
        if (d->nranks == 1)
            continue;
        if (d->equal_slices)
            TSP_NCCL(ncclAllGather(dst + (size_t)d->r0 * d->vs, dst, (size_t)d->m_local, dt, c->nccl, s)); // in place
        else
        {
            TSP_NCCL(ncclGroupStart());
            for (int r = 0; r < d->nranks; r++)
            {
                unsigned char *p = dst + (size_t)d->cuts[(size_t)r] * d->vs;
                const size_t cnt = (size_t)(d->cuts[(size_t)r + 1] - d->cuts[(size_t)r]);
                if (cnt)
                    TSP_NCCL(ncclBroadcast(p, p, cnt, dt, r, c->nccl, s));
            }
            TSP_NCCL(ncclGroupEnd());
        }
    }
    return TILESPMV_OK;
}

static int iterate_fused(tilespmv_dist *d, int niters, cudaStream_t s)
{
    const uint32_t peers = all_peers_mask(d), E0 = 0; // flag values are deltas to the call's epoch
    TSP_TRY(flag_signal(d, DIST_OFF_A, peers, E0, s)); // entered the call: my buffers are free for epoch E0
    for (int i = 0; i < niters; i++)
    {
        const uint32_t e = E0 + (uint32_t)i;
        const int sb = (d->cur + i) & 1, db = sb ^ 1;
        void *pp[TSP_MAX_PEERS];
        int np = 0;
        for (int r = 0; r < d->nranks; r++)
            if (r != d->rank)
                pp[np++] = xbuf(d, r, db);
        TSP_TRY(tilespmv_plan_set_peers(d->plan, np, pp, d->r0));
        // everybody has finished epoch e - 1: their stores into my src buffer have landed and nobody still reads the
        // buffer this epoch's stores go to
        TSP_TRY(flag_wait(d, DIST_OFF_A, peers, e, s));
        TSP_TRY(launch_units(d, xbuf(d, d->rank, sb), xbuf(d, d->rank, db), s, true));
        TSP_TRY(flag_signal(d, DIST_OFF_A, peers, e + 1, s));
    }
    TSP_TRY(flag_wait(d, DIST_OFF_A, peers, E0 + (uint32_t)niters, s)); // the final x is complete on this rank
    tilespmv_plan_set_peers(d->plan, 0, nullptr, 0);
    return TILESPMV_OK;
}

static int iterate_pipelined(tilespmv_dist *d, int niters, cudaStream_t s)
{
    const uint32_t peers = all_peers_mask(d), E0 = 0; // flag values are deltas to the call's epoch
    const int R = d->nranks, me = d->rank, nunits = 1 + (int)d->plan->sub.size();
    const size_t slice_off = (size_t)d->r0 * (size_t)d->vs, slice_bytes = (size_t)d->m_local * (size_t)d->vs;
    d->ev_push_valid[0] = d->ev_push_valid[1] = false;
    TSP_TRY(flag_signal(d, DIST_OFF_A, peers, E0, s)); // entered the call: peers may push epoch E0's slices into my buffer
    for (int i = 0; i < niters; i++)
    {
        const uint32_t e = E0 + (uint32_t)i;
        const int sb = (d->cur + i) & 1, db = sb ^ 1;
        // ---- compute stream: every launch waits only for the slices of x its columns read ----
        uint32_t waited = 0;
        for (int u = 0; u < nunits; u++)
        {
            if (i > 0) // iteration 0 reads the replicated x the call started with
            {
                const uint32_t need = d->deps[(size_t)u] & ~waited;
                TSP_TRY(flag_wait(d, DIST_OFF_D, need, e, s));
                waited |= need;
            }
            if (u == 0 && d->ev_push_valid[i & 1]) // the pushes of iteration i - 2 read the slice this iteration overwrites
                TSP_CUDA(cudaStreamWaitEvent(s, d->ev_push[i & 1], 0));
            TSP_TRY(plan_launch_unit(d->plan, u, xbuf(d, me, sb), xbuf(d, me, db) + slice_off, s, false));
        }
        TSP_CUDA(cudaEventRecord(d->ev_kernel[i & 1], s));
        TSP_TRY(flag_signal(d, DIST_OFF_A, peers, e + 1, s)); // finished epoch e: buffer sb is free again
        // ---- copy stream: push my slice of the next x, first to the rank that needs it first ----
        TSP_CUDA(cudaStreamWaitEvent(d->s_comm, d->ev_kernel[i & 1], 0));
        for (int k = 1; k < R; k++)
        {
            const int dst = (me - k + R) % R;
            TSP_TRY(flag_wait(d, DIST_OFF_A, 1u << dst, e, d->s_comm)); // dst finished epoch e - 1 (or entered the call)
            const bool dbg = d->debug && i == niters - 1 && k == 1 && slice_bytes;
            if (dbg)
                TSP_CUDA(cudaEventRecord(d->dbg0, d->s_comm));
            if (slice_bytes)
                TSP_CUDA(cudaMemcpyPeerAsync(xbuf(d, dst, db) + slice_off, d->peer_device[dst], xbuf(d, me, db) + slice_off, d->comm->device,
                                             slice_bytes, d->s_comm));
            if (dbg)
            {
                TSP_CUDA(cudaEventRecord(d->dbg1, d->s_comm));
                d->dbg_armed = true;
                d->dbg_bytes = slice_bytes;
                d->dbg_dst = dst;
            }
            TSP_TRY(flag_signal(d, DIST_OFF_D, 1u << dst, e + 1, d->s_comm));
        }
        TSP_CUDA(cudaEventRecord(d->ev_push[i & 1], d->s_comm));
        d->ev_push_valid[i & 1] = true;
    }
    // the final x is complete on this rank and my own pushes are done before the caller touches the buffers
    if (niters > 0)
        TSP_TRY(flag_wait(d, DIST_OFF_D, peers, E0 + (uint32_t)niters, s));
    for (int b = 0; b < 2; b++)
        if (d->ev_push_valid[b])
            TSP_CUDA(cudaStreamWaitEvent(s, d->ev_push[b], 0));
    return TILESPMV_OK;
}

// FUSED for the rows a peer reads next (its halo) + PIPELINED for everything else: the kernel's epilogue stores the
// rows inside need[q] straight into q's next x, one flag exchange with the halo neighbours orders the iterations, and
// the copy engines push the rest of the slice to every peer in the background (nobody reads it before the call ends,
// where every rank waits for all slices).  Falls back to the pipelined exchange when the halos are not small.
static int iterate_halo(tilespmv_dist *d, int niters, cudaStream_t s)
{
    if (!d->halo_ok)
        return iterate_pipelined(d, niters, s);
    const uint32_t peers = all_peers_mask(d), E0 = 0; // flag values are deltas to the call's epoch
    const int R = d->nranks, me = d->rank;
    const size_t vs = (size_t)d->vs, slice_off = (size_t)d->r0 * vs;
    d->ev_push_valid[0] = d->ev_push_valid[1] = false;
    // local rows of mine that peer q reads: [out_lo, out_hi)
    long long out_lo[COMM_MAX_RANKS], out_hi[COMM_MAX_RANKS];
    for (int q = 0; q < R; q++)
    {
        out_lo[q] = std::max(d->need_lo[(size_t)q], d->r0) - d->r0;
        out_hi[q] = std::max(out_lo[q], std::min(d->need_hi[(size_t)q], d->r0 + d->m_local) - d->r0);
        if (q == me)
            out_lo[q] = out_hi[q] = 0;
    }
    const uint32_t nb = d->halo_in | d->halo_out;
    TSP_TRY(flag_signal(d, DIST_OFF_A, peers, E0, s)); // entered the call
    for (int i = 0; i < niters; i++)
    {
        const uint32_t e = E0 + (uint32_t)i;
        const int sb = (d->cur + i) & 1, db = sb ^ 1;
        void *pp[TSP_MAX_PEERS];
        int np = 0;
        tilespmv_plan *P = d->plan;
        for (int q = 0; q < R; q++)
            if ((d->halo_out >> q) & 1u)
            {
                pp[np] = xbuf(d, q, db);
                P->peer_lo[np] = out_lo[q];
                P->peer_hi[np] = out_hi[q];
                np++;
            }
        P->npeers = np;
        P->row_offset = d->r0;
        for (int k = 0; k < TSP_MAX_PEERS; k++)
            P->peers[k] = k < np ? pp[k] : nullptr;
        // the halo neighbours have finished epoch e - 1: their halo stores into my src buffer have landed and they no
        // longer read the buffer this epoch's stores go to
        TSP_TRY(flag_wait(d, DIST_OFF_A, nb, e, s));
        if (d->ev_push_valid[i & 1]) // the background pushes of iteration i - 2 read the slice this iteration overwrites
            TSP_CUDA(cudaStreamWaitEvent(s, d->ev_push[i & 1], 0));
        TSP_TRY(launch_units(d, xbuf(d, me, sb), xbuf(d, me, db), s, true));
        TSP_CUDA(cudaEventRecord(d->ev_kernel[i & 1], s));
        TSP_TRY(flag_signal(d, DIST_OFF_A, peers, e + 1, s));
        // ---- background replication of everything the kernel did not store itself.  ONE wait on the copy stream (all
        //      peers have finished epoch e - 1, so the buffers this epoch's slices go to are free), then the copies to
        //      the different peers -- which have no order among them -- fork onto one stream per peer so that several copy
        //      engines run at once.  Spin kernels only ever sit on s_main and s_comm: with one spinning kernel per peer
        //      stream, 9+ streams alias onto the 8 hardware queues of the default CUDA_DEVICE_MAX_CONNECTIONS and block
        //      each other (measured: 2.06 ms per iteration instead of 0.2 on 8 GPUs). ----
        TSP_CUDA(cudaStreamWaitEvent(d->s_comm, d->ev_kernel[i & 1], 0));
        TSP_TRY(flag_wait(d, DIST_OFF_A, peers, e, d->s_comm));
        TSP_CUDA(cudaEventRecord(d->ev_fork, d->s_comm));
        for (int k = 1; k < R; k++)
        {
            const int dst = (me - k + R) % R;
            cudaStream_t sp = d->s_peer[dst];
            TSP_CUDA(cudaStreamWaitEvent(sp, d->ev_fork, 0));
            const long long seg[2][2] = {{0, out_lo[dst]}, {out_hi[dst], d->m_local}};
            for (int g = 0; g < 2; g++)
                if (seg[g][1] > seg[g][0])
                    TSP_CUDA(cudaMemcpyPeerAsync(xbuf(d, dst, db) + slice_off + (size_t)seg[g][0] * vs, d->peer_device[dst],
                                                 xbuf(d, me, db) + slice_off + (size_t)seg[g][0] * vs, d->comm->device,
                                                 (size_t)(seg[g][1] - seg[g][0]) * vs, sp));
            TSP_TRY(flag_signal(d, DIST_OFF_D, 1u << dst, e + 1, sp));
            TSP_CUDA(cudaEventRecord(d->ev_peer[dst], sp));
            TSP_CUDA(cudaStreamWaitEvent(d->s_comm, d->ev_peer[dst], 0));
        }
        TSP_CUDA(cudaEventRecord(d->ev_push[i & 1], d->s_comm));
        d->ev_push_valid[i & 1] = true;
    }
    // the final x is complete on this rank: halo stores (neighbours finished the last epoch) + every background slice
    if (niters > 0)
    {
        TSP_TRY(flag_wait(d, DIST_OFF_A, nb, E0 + (uint32_t)niters, s));
        TSP_TRY(flag_wait(d, DIST_OFF_D, peers, E0 + (uint32_t)niters, s));
    }
    for (int b = 0; b < 2; b++)
        if (d->ev_push_valid[b])
            TSP_CUDA(cudaStreamWaitEvent(s, d->ev_push[b], 0));
    tilespmv_plan_set_peers(d->plan, 0, nullptr, 0);
    return TILESPMV_OK;
}

} // namespace tsp

using namespace tsp;

extern "C"
{

int tilespmv_comm_create(const char *name, int rank, int nranks, unsigned flags, tilespmv_comm **out)
try
{
    clear_error();
    return comm_create(name, rank, nranks, flags, out);
}
TSP_CATCH_INT("tilespmv_comm_create")
void tilespmv_comm_destroy(tilespmv_comm *comm)
try
{
    clear_error();
    comm_destroy(comm);
}
TSP_CATCH_VOID("tilespmv_comm_destroy")
int tilespmv_comm_barrier(tilespmv_comm *comm)
try
{
    clear_error();
    if (!comm)
    {
        set_error("comm_barrier: null argument");
        return TILESPMV_ERR_INVALID;
    }
    return comm_barrier(comm);
}
TSP_CATCH_INT("tilespmv_comm_barrier")

int tilespmv_dist_create(tilespmv_comm *comm, const tilespmv_dmat *local_rows, const int64_t *row_cuts, const tilespmv_plan_options *opts,
                         unsigned flags, tilespmv_dist **out)
try
{
    clear_error();
    return dist_create(comm, local_rows, row_cuts, opts, flags, out);
}
TSP_CATCH_INT("tilespmv_dist_create")
void tilespmv_dist_destroy(tilespmv_dist *dist)
try
{
    clear_error();
    if (!dist)
        return;
    cudaDeviceSynchronize();
    if (dist->comm && dist->nranks > 1)
        comm_barrier(dist->comm); // no peer still copies into (or out of) the block that is about to be freed
    delete dist;
}
TSP_CATCH_VOID("tilespmv_dist_destroy")

// enqueue all iterations of one call on s_main (+ s_comm)
static int iterate_enqueue(tilespmv_dist *d, int niters, int exchange)
{
    switch (exchange)
    {
    case TILESPMV_EXCHANGE_FUSED:
        return iterate_fused(d, niters, d->s_main);
    case TILESPMV_EXCHANGE_PIPELINED:
        return iterate_pipelined(d, niters, d->s_main);
    case TILESPMV_EXCHANGE_HALO:
        return iterate_halo(d, niters, d->s_main);
    default:
        return iterate_nccl(d, niters, d->s_main);
    }
}

int tilespmv_dist_iterate(tilespmv_dist *dist, const void *d_x0, int niters, int exchange, void *stream)
try
{
    clear_error();
    if (!dist || niters < 0 || exchange < TILESPMV_EXCHANGE_NCCL || exchange > TILESPMV_EXCHANGE_HALO)
    {
        set_error("dist_iterate: invalid argument");
        return TILESPMV_ERR_INVALID;
    }
    tilespmv_dist *d = dist;
    cudaStream_t su = static_cast<cudaStream_t>(stream);
    if (d_x0)
    {
        d->cur = 0;
        TSP_CUDA(cudaMemcpyAsync(xbuf(d, d->rank, 0), d_x0, (size_t)d->n * (size_t)d->vs, cudaMemcpyDeviceToDevice, su));
    }
    if (niters == 0)
        return TILESPMV_OK;
    if (d->nranks == 1)
        exchange = TILESPMV_EXCHANGE_NCCL; // nothing to exchange: plain ping-pong
    // the loop runs on the library's own stream (the caller's may be the legacy default stream, which cannot be captured)
    TSP_CUDA(cudaEventRecord(d->ev_user, su));
    TSP_CUDA(cudaStreamWaitEvent(d->s_main, d->ev_user, 0));
    TSP_LAUNCH(set_epoch_kernel, 1, 1, 0, d->s_main, reinterpret_cast<uint32_t *>(d->block.as<unsigned char>() + DIST_OFF_EPOCH), d->epoch);
    int rc = TILESPMV_OK;
    const bool graphable = d->use_graph && exchange != TILESPMV_EXCHANGE_NCCL;
    if (graphable)
    {
        tilespmv_dist::GraphEntry *g = nullptr;
        for (auto &e : d->graphs)
            if (e.exchange == exchange && e.niters == niters && e.cur == d->cur)
                g = &e;
        if (!g)
        {
            const int64_t counted = g_launches.load();
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamBeginCapture(d->s_main, cudaStreamCaptureModeThreadLocal);
            if (ce == cudaSuccess)
            {
                rc = iterate_enqueue(d, niters, exchange);
                ce = cudaStreamEndCapture(d->s_main, &graph);
            }
            const int64_t captured = g_launches.load() - counted;
            g_launches.store(counted); // launches recorded while capturing are not executions
            cudaGraphExec_t exec = nullptr;
            if (rc == TILESPMV_OK && ce == cudaSuccess && graph)
                ce = cudaGraphInstantiate(&exec, graph, 0);
            if (graph)
                cudaGraphDestroy(graph);
            if (rc != TILESPMV_OK || ce != cudaSuccess || !exec)
            {
                // e.g. a driver that cannot capture peer copies: enqueue directly from now on (a real error shows up again there)
                cudaGetLastError();
                if (exec)
                    cudaGraphExecDestroy(exec);
                d->use_graph = false;
                rc = TILESPMV_OK;
                clear_error();
            }
            else
            {
                if (d->graphs.size() >= 8)
                {
                    cudaGraphExecDestroy(d->graphs.front().exec);
                    d->graphs.erase(d->graphs.begin());
                }
                d->graphs.push_back({exchange, niters, d->cur, exec, captured});
                g = &d->graphs.back();
            }
        }
        if (g)
        {
            TSP_CUDA(cudaGraphLaunch(g->exec, d->s_main));
            g_launches.fetch_add(g->launches, std::memory_order_relaxed);
        }
        else
            rc = iterate_enqueue(d, niters, exchange);
    }
    else
        rc = iterate_enqueue(d, niters, exchange);
    TSP_TRY(rc);
    if (exchange != TILESPMV_EXCHANGE_NCCL)
        d->epoch += (uint32_t)niters + 1u;
    d->cur = (d->cur + niters) & 1;
    TSP_CUDA(cudaEventRecord(d->ev_done, d->s_main));
    TSP_CUDA(cudaStreamWaitEvent(su, d->ev_done, 0));
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_dist_iterate")

void *tilespmv_dist_x(tilespmv_dist *dist)
{
    clear_error();
    return dist ? xbuf(dist, dist->rank, dist->cur) : nullptr;
}

tilespmv_plan *tilespmv_dist_plan(tilespmv_dist *dist)
{
    clear_error();
    return dist ? dist->plan : nullptr;
}

int tilespmv_dist_sync(tilespmv_dist *dist, void *stream)
try
{
    clear_error();
    if (!dist)
    {
        set_error("dist_sync: null argument");
        return TILESPMV_ERR_INVALID;
    }
    TSP_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    TSP_CUDA(cudaStreamSynchronize(dist->s_main));
    TSP_CUDA(cudaStreamSynchronize(dist->s_comm));
    if (dist->dbg_armed)
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, dist->dbg0, dist->dbg1) == cudaSuccess)
            fprintf(stderr, "[tilespmv_dist rank %d] push of %zu bytes to rank %d: %.3f ms = %.1f GB/s\n", dist->rank, dist->dbg_bytes,
                    dist->dbg_dst, ms, ms > 0 ? dist->dbg_bytes * 1e-6 / ms : 0.0);
        dist->dbg_armed = false;
    }
    uint32_t err = 0;
    TSP_CUDA(cudaMemcpy(&err, dist->block.as<unsigned char>() + DIST_OFF_ERR, sizeof(err), cudaMemcpyDeviceToHost));
    if (err)
    {
        set_error("dist: rank %d timed out waiting for a flag of rank %u (a peer died or left the loop early)", dist->rank, err - 1u);
        return TILESPMV_ERR_CUDA;
    }
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_dist_sync")

int tilespmv_dist_get_info(const tilespmv_dist *dist, tilespmv_dist_info *info)
try
{
    clear_error();
    if (!dist || !info)
    {
        set_error("dist_get_info: null argument");
        return TILESPMV_ERR_INVALID;
    }
    memset(info, 0, sizeof(*info));
    info->rank = dist->rank;
    info->nranks = dist->nranks;
    info->row0 = dist->r0;
    info->rows = dist->m_local;
    info->launch_units = 1 + (int)dist->plan->sub.size();
    info->equal_slices = dist->equal_slices ? 1 : 0;
    for (int u = 0; u < info->launch_units && u < 64; u++)
        info->unit_deps[u] = dist->deps[(size_t)u];
    info->slice_bytes = dist->m_local * dist->vs;
    info->halo_eligible = dist->halo_ok ? 1 : 0;
    info->need_lo = dist->need_lo[(size_t)dist->rank];
    info->need_hi = dist->need_hi[(size_t)dist->rank];
    info->device_bytes = (int64_t)dist->block.bytes + dist->plan->device_bytes();
    return TILESPMV_OK;
}
TSP_CATCH_INT("tilespmv_dist_get_info")

} // extern "C"

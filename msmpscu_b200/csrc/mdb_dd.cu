// mdb_dd.cu -- ONE box over several GPUs: z-slab decomposition driven from inside the library.
//
// Replaces the reference's multi-GPU scheme for a single box -- REPLICATED positions, host-staged copies: XP gathered
// through the host after the predictor (CommonGPU/MD_Globle_Variables_GPU.F90:2026-2040), DEN gathered between the
// passes (CommonGPU/MD_EAM_ForceTable_GPU.F90:617-642), every device sorting all atoms on the host
// (CommonGPU/MD_NeighborsList_GPU.F90:1421-1695) -- by owned + ghost atoms per rank:
//
//   per step     predictor (owned) -> boundary-layer {x,y,z,den} records to the two neighbour ranks, their boundary layers
//                into the ghost ranges, together with the neighbours' max displacement since the rebuild (the guard of
//                the distance-class shortcut needs it for a rank's own halo only: no global reduction)
//                -> density pass (owned tiles) -> the same exchange again (the records now carry dF/drho)
//                -> force pass -> EPC + corrector (owned; fused in front of the next predictor inside a block)
//   per rebuild  velocities / displacements of the ghost layers are fetched, then every rank re-sorts ONLY the atoms that
//                end up in its own z-layers of cells (candidates: its owned range and its two ghost layers -- an atom moves
//                far less than a cell between rebuilds), one 16-byte all-gather gives every rank its first global slot,
//                the new boundary layers (records, types, status, ids, cell counts) are exchanged, lists are built for
//                the rank's own tiles.  Nothing is broadcast and no rank touches atoms outside its slab + ghosts.
//
// All per-atom arrays keep the GLOBAL slot numbering of the common cell-sorted order (cells ascending, atoms of a cell by
// descending original id), so slots, tiles and list entries mean the same on every rank and an owned slice is
// bit-identical to the same slice of a single-GPU run; arrays are therefore allocated full length and only the owned +
// ghost ranges are kept current (sharding the allocation is an index offset away and not done here).
// Every exchange is enqueued on the context's stream from C: ncclSend / ncclRecv in one group (NCCL is bound with dlopen,
// so the library loads without it), or -- backend for single-GPU tests -- device-to-device copies between the contexts
// of all ranks living in one process on one stream.  The host waits twice per rebuild (sizes of the exchanged ranges).
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <cstring>
#include "mdb_internal.cuh"

// ------------------------------------------------------------------------------------
// NCCL, bound at run time
// ------------------------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    bool tried = false;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi &nccl()
{
    static NcclApi a;
    if (a.tried) return a;
    a.tried = true;
    // the copy a host program already loaded (torch bundles one) is reused; else the system library
    a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!a.lib) a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) return a;
#define BIND(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, name))
    BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommDestroy, "ncclCommDestroy");
    BIND(Send, "ncclSend"); BIND(Recv, "ncclRecv"); BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd");
    BIND(AllGather, "ncclAllGather"); BIND(AllReduce, "ncclAllReduce"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    if (!a.GetUniqueId || !a.CommInitRank || !a.Send || !a.Recv || !a.GroupStart || !a.GroupEnd || !a.AllGather || !a.AllReduce) {
        dlclose(a.lib);
        a.lib = nullptr;
    }
    return a;
}
#define NCCL_TRY(c, call)                                                                                          \
    do {                                                                                                           \
        ncclResult_t r__ = (call);                                                                                 \
        if (r__ != ncclSuccess)                                                                                    \
            return mdb_fail((c), MDB_ERR_CUDA, "%s failed: %s", #call, nccl().GetErrorString ? nccl().GetErrorString(r__) : "?"); \
    } while (0)

extern "C" int mdb_dd_nccl_id(void *id128)
{
    if (!id128) return MDB_ERR_ARG;
    if (!nccl().lib) return MDB_ERR_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    return nccl().GetUniqueId(reinterpret_cast<ncclUniqueId *>(id128)) == ncclSuccess ? MDB_OK : MDB_ERR_CUDA;
}

extern "C" int mdb_dd_nccl_init(mdb_ctx *c, const void *id128)
{
    if (!c || !id128) return mdb_fail(c, MDB_ERR_ARG, "mdb_dd_nccl_init: null argument");
    if (!c->dd_on) return mdb_fail(c, MDB_ERR_STATE, "mdb_dd_nccl_init: mdb_dd_set(rank, nranks > 1) first");
    if (!nccl().lib) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_dd_nccl_init: libnccl.so.2 not found");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    NCCL_TRY(c, nccl().CommInitRank(&comm, c->dd_n, id, c->dd_rank));
    if (c->dd_comm) nccl().CommDestroy((ncclComm_t)c->dd_comm);
    c->dd_comm = comm;
    c->dd_peers.clear();
    return MDB_OK;
}

// In-process backend: the contexts of ALL ranks live in this process on ONE device and share ONE stream; an exchange is a
// device-to-device copy out of the neighbour's arrays.  Every collective entry point below, called on any of the contexts,
// then drives all ranks in lock step.  (Single-GPU parity tests of the decomposition; the data path is the same.)
extern "C" int mdb_dd_local_attach(mdb_ctx **ctxs, int nranks)
{
    if (!ctxs || nranks < 2) return MDB_ERR_ARG;
    for (int r = 0; r < nranks; r++) {
        mdb_ctx *c = ctxs[r];
        if (!c) return MDB_ERR_ARG;
        if (c->dev != ctxs[0]->dev || c->n != ctxs[0]->n) return mdb_fail(c, MDB_ERR_ARG, "mdb_dd_local_attach: contexts must share device and box");
        int rc = mdb_dd_set(c, r, nranks);
        if (rc < 0) return rc;
        c->stream = ctxs[0]->stream;
    }
    for (int r = 0; r < nranks; r++) ctxs[r]->dd_peers.assign(ctxs, ctxs + nranks);
    return MDB_OK;
}

static void p2p_free(mdb_ctx *c);
void mdb_dd_free(mdb_ctx *c)
{
    p2p_free(c);
    if (c->dd_xs) { cudaStreamSynchronize(c->dd_xs); cudaStreamDestroy(c->dd_xs); c->dd_xs = nullptr; }
    for (int k = 0; k < 4; k++) if (c->dd_ev[k]) { cudaEventDestroy(c->dd_ev[k]); c->dd_ev[k] = nullptr; }
    if (c->dd_comm && nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)c->dd_comm);
    c->dd_comm = nullptr;
    if (c->dd_dev) cudaFree(c->dd_dev);
    c->dd_dev = nullptr;
    // in-process backend: the surviving ranks forget this context and leave its stream
    for (mdb_ctx *p : c->dd_peers) {
        if (p == c) continue;
        if (p->stream == c->own_stream) { cudaStreamSynchronize(p->stream); p->stream = p->own_stream; }
        p->dd_peers.clear();
        p->dd_built = false; p->list_valid = false;
    }
    c->dd_peers.clear();
}

// ------------------------------------------------------------------------------------
// exchanges
// ------------------------------------------------------------------------------------
enum { DF_POS = 0, DF_XP1X, DF_XP1Y, DF_XP1Z, DF_DISX, DF_DISY, DF_DISZ, DF_ITYP, DF_STATU, DF_GID, DF_NAC, DF_NAAC };
static char *dd_field(mdb_ctx *c, int f, size_t &elem)
{
    const size_t n = c->n;
    elem = sizeof(double);
    switch (f) {
    case DF_POS: elem = sizeof(double4); return (char *)c->pos;
    case DF_XP1X: return (char *)c->xp1;
    case DF_XP1Y: return (char *)(c->xp1 + n);
    case DF_XP1Z: return (char *)(c->xp1 + 2 * n);
    case DF_DISX: return (char *)c->dis;
    case DF_DISY: return (char *)(c->dis + n);
    case DF_DISZ: return (char *)(c->dis + 2 * n);
    case DF_ITYP: elem = sizeof(int); return (char *)c->ityp;
    case DF_STATU: elem = sizeof(int); return (char *)c->statu;
    case DF_GID: elem = sizeof(int); return (char *)c->gid;
    case DF_NAC: elem = sizeof(int); return (char *)c->nac;
    case DF_NAAC: elem = sizeof(int); return (char *)c->naac;
    }
    return nullptr;
}
static int dd_scratch(mdb_ctx *c)
{
    if (c->dd_dev) return MDB_OK;
    CUDA_TRY(c, cudaMalloc(&c->dd_dev, sizeof(int) * (5 * (size_t)c->dd_n + 64)));   // [4 R] table, [64] scratch, [R] time-step masks
    CUDA_TRY(c, cudaMemsetAsync(c->dd_dev, 0, sizeof(int) * (5 * (size_t)c->dd_n + 64), c->stream));
    return MDB_OK;
}

// max displacement^2 since the rebuild of the two neighbour ranks merged into this rank's guard value
__global__ void k_d2max_merge(int *counters, const int *a, const int *b)
{
    atomicMax(&counters[CNT_D2MAX], max(*a, *b)); // non-negative floats order like ints
}

// ranges (element indices, identical numbering on every rank): what I send down / up, what I receive from above / below
struct XRanges { int sb0, sb1, st0, st1, ga0, ga1, gb0, gb1; };
static XRanges atom_ranges(const mdb_ctx *c)
{
    const int *d = c->dd_info;
    return XRanges{d[6], d[7], d[8], d[9], d[4], d[5], d[2], d[3]};
}
static XRanges cell_ranges(const mdb_ctx *c)
{
    const int cl = c->ncell[0] * c->ncell[1], ncz = c->ncell[2];
    const int zl0 = c->dd_info[12] / cl, zl1 = c->dd_info[13] / cl;
    const int zb = (zl0 - 1 + ncz) % ncz, za = zl1 % ncz;
    return XRanges{zl0 * cl, (zl0 + 1) * cl, (zl1 - 1) * cl, zl1 * cl, za * cl, (za + 1) * cl, zb * cl, (zb + 1) * cl};
}

static int dd_exchange(mdb_ctx *c, const int *fields, int nf, const XRanges &R, bool with_d2max)
{
    const int below = c->dd_info[10], above = c->dd_info[11];
    ProfScope ps(c, MDB_K_EXCHANGE, 1);
    if (!c->dd_peers.empty()) {
        mdb_ctx *pa = c->dd_peers[above], *pb = c->dd_peers[below];
        for (int i = 0; i < nf; i++) {
            size_t e, e2;
            char *mine = dd_field(c, fields[i], e);
            const char *fa = dd_field(pa, fields[i], e2), *fb = dd_field(pb, fields[i], e2);
            if (R.ga1 > R.ga0) CUDA_TRY(c, cudaMemcpyAsync(mine + e * R.ga0, fa + e * R.ga0, e * (size_t)(R.ga1 - R.ga0), cudaMemcpyDeviceToDevice, c->stream));
            if (R.gb1 > R.gb0) CUDA_TRY(c, cudaMemcpyAsync(mine + e * R.gb0, fb + e * R.gb0, e * (size_t)(R.gb1 - R.gb0), cudaMemcpyDeviceToDevice, c->stream));
        }
        if (with_d2max) {
            k_d2max_merge<<<1, 1, 0, c->stream>>>(c->counters, pa->dd_dev + 4 * c->dd_n + 2, pb->dd_dev + 4 * c->dd_n + 2);
            CUDA_TRY(c, cudaGetLastError());
        }
        return MDB_OK;
    }
    if (!c->dd_comm) return mdb_fail(c, MDB_ERR_STATE, "slab decomposition: no communicator (mdb_dd_nccl_init or mdb_dd_local_attach)");
    NcclApi &N = nccl();
    ncclComm_t comm = (ncclComm_t)c->dd_comm;
    int *scr = c->dd_dev + 4 * c->dd_n; // [0] from above, [1] from below
    NCCL_TRY(c, N.GroupStart());
    for (int i = 0; i < nf; i++) {
        size_t e;
        char *p = dd_field(c, fields[i], e);
        // a rank's first send is its bottom layer: with two ranks both neighbours are the same peer and its first receive
        // must therefore be its ghost layer ABOVE
        if (R.sb1 > R.sb0) NCCL_TRY(c, N.Send(p + e * R.sb0, e * (size_t)(R.sb1 - R.sb0), ncclChar, below, comm, c->stream));
        if (R.st1 > R.st0) NCCL_TRY(c, N.Send(p + e * R.st0, e * (size_t)(R.st1 - R.st0), ncclChar, above, comm, c->stream));
        if (R.ga1 > R.ga0) NCCL_TRY(c, N.Recv(p + e * R.ga0, e * (size_t)(R.ga1 - R.ga0), ncclChar, above, comm, c->stream));
        if (R.gb1 > R.gb0) NCCL_TRY(c, N.Recv(p + e * R.gb0, e * (size_t)(R.gb1 - R.gb0), ncclChar, below, comm, c->stream));
    }
    if (with_d2max) {
        NCCL_TRY(c, N.Send(c->counters + CNT_D2MAX, sizeof(int), ncclChar, below, comm, c->stream));
        NCCL_TRY(c, N.Send(c->counters + CNT_D2MAX, sizeof(int), ncclChar, above, comm, c->stream));
        NCCL_TRY(c, N.Recv(scr + 0, sizeof(int), ncclChar, above, comm, c->stream));
        NCCL_TRY(c, N.Recv(scr + 1, sizeof(int), ncclChar, below, comm, c->stream));
    }
    NCCL_TRY(c, N.GroupEnd());
    if (with_d2max) {
        k_d2max_merge<<<1, 1, 0, c->stream>>>(c->counters, scr, scr + 1);
        CUDA_TRY(c, cudaGetLastError());
    }
    return MDB_OK;
}

// in-process backend: a rank's guard value has to be frozen before its neighbours read it (they merge into their own)
__global__ void k_copy_int(int *dst, const int *src) { *dst = *src; }
static int dd_publish_d2max(mdb_ctx *c)
{
    if (c->dd_peers.empty()) return MDB_OK;
    k_copy_int<<<1, 1, 0, c->stream>>>(c->dd_dev + 4 * c->dd_n + 2, c->counters + CNT_D2MAX);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}


// ------------------------------------------------------------------------------------
// Peer-to-peer ghost exchange of the step loop (NCCL backend, same node): the boundary layers are written straight into
// the neighbours' position arrays over NVLink (CUDA IPC mappings, copy engines), ordered by flags in device memory.
// An ncclSend/ncclRecv group costs ~0.12 ms per exchange here (launch + rendezvous latency, 20 exchanges per list period);
// the direct form is two copies and two one-thread handshake kernels.  Protocol of exchange number e (all ranks issue the
// same sequence of exchanges):
//   handshake 1  tell both neighbours "everything I read from you up to exchange e-1 is consumed" (all earlier kernels of
//                this stream have finished), wait for the same word from both      -> their ghost ranges may be overwritten
//   push         cudaMemcpyAsync of my bottom / top layer into the neighbour's array (same global slots there)
//   handshake 2  publish "exchange e has landed" (+ my max displacement since the rebuild) in both neighbours' flags, wait
//                for theirs, merge their displacement bound into mine
// Flags are monotonic counters, written with system-scope fences; a wait gives up after ~seconds and raises an error
// counter instead of hanging the device.  The rebuild's own exchanges (sizes known to the host only then) stay on NCCL.
// MDB_DD_P2P=0 in the environment keeps everything on NCCL.
// ------------------------------------------------------------------------------------
struct P2PState {
    bool on = false;
    int epoch = 0, buf = 0;            // buf: which of the two position buffers is the current one (toggles at a local rebuild)
    double4 *my_pos[2] = {nullptr, nullptr};
    double4 *peer_pos[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; // [0 = rank below, 1 = rank above][buffer]
    int *peer_flags[2] = {nullptr, nullptr};
    int *flags = nullptr;              // mine: [0,1] arrive from below / above, [2,3] ack from below / above, [4,5] d2max of below / above, [6] errors
    void *opened[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
static P2PState *p2p_of(mdb_ctx *c) { return reinterpret_cast<P2PState *>(c->dd_p2p); }

#define P2P_SPIN_LIMIT 20000000LL
// word `slot` of both neighbours := val (and optionally my displacement bound next to it), then wait until both of MY
// words `mine0`, `mine1` have reached val
__global__ void k_p2p_handshake(int *peer_lo, int *peer_hi, int slot_lo, int slot_hi, int val, volatile int *flags, int mine0, int mine1,
                                int *counters, int with_d2)
{
    if (with_d2) { peer_lo[5] = counters[CNT_D2MAX]; peer_hi[4] = counters[CNT_D2MAX]; } // I am "above" the rank below me, "below" the rank above
    __threadfence_system();
    *(volatile int *)(peer_lo + slot_lo) = val;
    *(volatile int *)(peer_hi + slot_hi) = val;
    long long spins = 0;
    while ((flags[mine0] < val || flags[mine1] < val) && ++spins < P2P_SPIN_LIMIT) { }
    if (spins >= P2P_SPIN_LIMIT) atomicAdd((int *)flags + 6, 1);
    __threadfence_system();
    if (with_d2) atomicMax(&counters[CNT_D2MAX], max(flags[4], flags[5]));
}

static void p2p_free(mdb_ctx *c)
{
    P2PState *S = p2p_of(c);
    if (!S) return;
    for (void *p : S->opened) if (p) cudaIpcCloseMemHandle(p);
    if (S->flags) cudaFree(S->flags);
    delete S;
    c->dd_p2p = nullptr;
}

// after the first build: export my two position buffers and my flags, import the neighbours'
static int p2p_init(mdb_ctx *c)
{
    p2p_free(c);
    const char *env = getenv("MDB_DD_P2P");
    if (env && atoi(env) == 0) return MDB_OK;
    if (!c->dd_comm) return MDB_OK;
    P2PState *S = new P2PState();
    c->dd_p2p = S;
    auto give_up = [&](const char *why) { (void)why; p2p_free(c); cudaGetLastError(); return (int)MDB_OK; };
    if (cudaMalloc(&S->flags, 64 * sizeof(int)) != cudaSuccess) return give_up("flags");
    cudaMemsetAsync(S->flags, 0, 64 * sizeof(int), c->stream);
    S->my_pos[0] = c->pos; S->my_pos[1] = c->pos_alt;
    cudaIpcMemHandle_t mine[3], theirs[2][3];
    int ok = 1;
    if (cudaIpcGetMemHandle(&mine[0], c->pos) != cudaSuccess || cudaIpcGetMemHandle(&mine[1], c->pos_alt) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine[2], S->flags) != cudaSuccess) { ok = 0; cudaGetLastError(); memset(mine, 0, sizeof(mine)); }
    // handles travel over the communicator that already exists (device staging: NCCL moves device memory)
    char *d = nullptr;
    const size_t hb = sizeof(mine);
    if (cudaMalloc(&d, 3 * hb + 16) != cudaSuccess) return give_up("staging");
    int *dok = reinterpret_cast<int *>(d + 3 * hb);
    cudaMemcpyAsync(d, mine, hb, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(dok, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream);
    NcclApi &N = nccl();
    ncclComm_t comm = (ncclComm_t)c->dd_comm;
    const int below = c->dd_info[10], above = c->dd_info[11];
    NCCL_TRY(c, N.GroupStart());
    NCCL_TRY(c, N.Send(d, hb, ncclChar, below, comm, c->stream));
    NCCL_TRY(c, N.Send(d, hb, ncclChar, above, comm, c->stream));
    NCCL_TRY(c, N.Recv(d + 2 * hb, hb, ncclChar, above, comm, c->stream));   // (first receive = from above: see dd_exchange)
    NCCL_TRY(c, N.Recv(d + hb, hb, ncclChar, below, comm, c->stream));
    NCCL_TRY(c, N.GroupEnd());
    // every rank must take the same decision: all-reduce (min) of "my exports worked"
    NCCL_TRY(c, N.AllReduce(dok, dok, 1, ncclInt, ncclMin, comm, c->stream));
    cudaMemcpyAsync(theirs[0], d + hb, hb, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(theirs[1], d + 2 * hb, hb, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(&ok, dok, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (!ok) return give_up("export");
    int opened_ok = 1;
    for (int dir = 0; dir < 2 && opened_ok; dir++) {
        if (dir == 1 && below == above) { // two ranks: both neighbours are the same process, a handle may be opened once only
            S->peer_pos[1][0] = S->peer_pos[0][0]; S->peer_pos[1][1] = S->peer_pos[0][1]; S->peer_flags[1] = S->peer_flags[0];
            break;
        }
        void *p[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < 3; k++) {
            if (cudaIpcOpenMemHandle(&p[k], theirs[dir][k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened_ok = 0; cudaGetLastError(); break; }
            S->opened[3 * dir + k] = p[k];
        }
        S->peer_pos[dir][0] = (double4 *)p[0]; S->peer_pos[dir][1] = (double4 *)p[1]; S->peer_flags[dir] = (int *)p[2];
    }
    // again a common decision
    int *d2 = nullptr;
    if (cudaMalloc(&d2, sizeof(int)) != cudaSuccess) return give_up("staging");
    cudaMemcpyAsync(d2, &opened_ok, sizeof(int), cudaMemcpyHostToDevice, c->stream);
    NCCL_TRY(c, N.AllReduce(d2, d2, 1, ncclInt, ncclMin, comm, c->stream));
    cudaMemcpyAsync(&opened_ok, d2, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(d2);
    if (!opened_ok) return give_up("open");
    S->on = true; S->epoch = 0; S->buf = 0;
    return MDB_OK;
}

static int p2p_exchange(mdb_ctx *c, bool with_d2max, cudaStream_t xs = nullptr)
{
    cudaStream_t keep = c->stream;
    if (xs) c->stream = xs; // (ProfScope and the launches below follow c->stream)
    struct Restore { mdb_ctx *c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{c, keep};
    P2PState *S = p2p_of(c);
    const XRanges R = atom_ranges(c);
    const int e = ++S->epoch;
    ProfScope ps(c, MDB_K_EXCHANGE, 2);
    // handshake 1: acks of exchange e-1 (word 3 of the rank below = "the rank above me consumed", word 2 of the rank above)
    k_p2p_handshake<<<1, 1, 0, c->stream>>>(S->peer_flags[0], S->peer_flags[1], 3, 2, e - 1, S->flags, 2, 3, c->counters, 0);
    double4 *lo = S->peer_pos[0][S->buf], *hi = S->peer_pos[1][S->buf];
    if (R.sb1 > R.sb0) CUDA_TRY(c, cudaMemcpyAsync(lo + R.sb0, c->pos + R.sb0, sizeof(double4) * (size_t)(R.sb1 - R.sb0), cudaMemcpyDefault, c->stream));
    if (R.st1 > R.st0) CUDA_TRY(c, cudaMemcpyAsync(hi + R.st0, c->pos + R.st0, sizeof(double4) * (size_t)(R.st1 - R.st0), cudaMemcpyDefault, c->stream));
    // handshake 2: arrival of exchange e (word 1 of the rank below = "from above", word 0 of the rank above = "from below")
    k_p2p_handshake<<<1, 1, 0, c->stream>>>(S->peer_flags[0], S->peer_flags[1], 1, 0, e, S->flags, 0, 1, c->counters, with_d2max ? 1 : 0);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}
static int p2p_check(mdb_ctx *c)
{
    P2PState *S = p2p_of(c);
    if (!S || !S->on) return MDB_OK;
    int err = 0;
    CUDA_TRY(c, cudaMemcpy(&err, S->flags + 6, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return mdb_fail(c, MDB_ERR_STATE, "slab decomposition: a neighbour rank did not answer a ghost exchange (%d time-outs)", err);
    return MDB_OK;
}

// runs one phase on this rank (NCCL backend) or on every rank of the process (in-process backend)
template <class F>
static int all_ranks(mdb_ctx *c, F f)
{
    if (c->dd_peers.empty()) return f(c);
    for (mdb_ctx *p : c->dd_peers) {
        int rc = f(p);
        if (rc < 0) { if (p != c) c->err = p->err; return rc; }
    }
    return MDB_OK;
}

static int x_pos(mdb_ctx *c, bool with_d2max)
{
    const int f[1] = {DF_POS};
    int rc = MDB_OK;
    if (P2PState *S = p2p_of(c); S && S->on) return p2p_exchange(c, with_d2max);
    if (with_d2max && (rc = all_ranks(c, [](mdb_ctx *p) { return dd_publish_d2max(p); })) < 0) return rc;
    return all_ranks(c, [&](mdb_ctx *p) { return dd_exchange(p, f, 1, atom_ranges(p), with_d2max); });
}

// ------------------------------------------------------------------------------------
// rebuild
// ------------------------------------------------------------------------------------
static int dd_check_layers(mdb_ctx *c)
{
    const int ncz = c->ncell[2], R = c->dd_n;
    if (ncz / R < 2) return mdb_fail(c, MDB_ERR_ARG, "slab decomposition: %d z-layers of cells over %d ranks (every rank needs two)", ncz, R);
    for (int d = 0; d < 3; d++)
        if (!c->box.pd[d]) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "slab decomposition: the box must be periodic in x, y and z");
    return MDB_OK;
}

// the first build: every rank holds the whole (identical) initial state, sorts it and builds the lists of its own tiles
static int dd_first_build(mdb_ctx *c)
{
    int rc = dd_check_layers(c);
    if (rc < 0) return rc;
    if ((rc = dd_scratch(c)) < 0) return rc;
    c->dd_built = false;
    if ((rc = mdb_list_rebuild_checked(c)) < 0) return rc;
    if (!c->tiled.active) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "slab decomposition needs the tiled path");
    if ((rc = mdb_dd_update(c)) < 0) return rc;
    c->dd_built = true;
    return c->h_counters[CNT_OOB] > 0 ? mdb_fail(c, MDB_ERR_UNSUPPORTED, "slab decomposition: atoms outside the box") : MDB_OK;
}

static int dd_local_rebuild(mdb_ctx *c)
{
    int rc;
    const int R = c->dd_n;
    // (1) velocities, displacements and status of the ghost layers: candidates for my cells
    {
        const int f[7] = {DF_XP1X, DF_XP1Y, DF_XP1Z, DF_DISX, DF_DISY, DF_DISZ, DF_STATU};
        if ((rc = all_ranks(c, [&](mdb_ctx *p) { return dd_exchange(p, f, 7, atom_ranges(p), false); })) < 0) return rc;
    }
    // (2) bin the candidates into my layers, scan
    auto cand_of = [](const mdb_ctx *p, int cand[6]) {
        const int *d = p->dd_info;
        cand[0] = d[2]; cand[1] = d[3]; cand[2] = d[0]; cand[3] = d[1]; cand[4] = d[4]; cand[5] = d[5];
    };
    const int cl = c->ncell[0] * c->ncell[1];
    if ((rc = all_ranks(c, [&](mdb_ctx *p) {
            int cand[6];
            cand_of(p, cand);
            return mdb_cells_dd_count(p, cand, p->dd_info[12] / cl, p->dd_info[13] / cl, p->dd_dev + 4 * p->dd_rank);
        })) < 0) return rc;
    // (3) every rank learns every rank's {owned, bottom layer, top layer, max per cell}: 16 bytes per rank
    if (c->dd_peers.empty()) {
        NCCL_TRY(c, nccl().AllGather(c->dd_dev + 4 * c->dd_rank, c->dd_dev, 4, ncclInt, (ncclComm_t)c->dd_comm, c->stream));
        c->dd_tab.resize(4 * R);
        CUDA_TRY(c, cudaMemcpyAsync(c->h_dd, c->counters, sizeof(int) * CNT__N, cudaMemcpyDeviceToHost, c->stream));
        std::vector<int> tab(4 * R);
        CUDA_TRY(c, cudaMemcpyAsync(tab.data(), c->dd_dev, sizeof(int) * 4 * R, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        c->dd_tab = tab;
        if (c->h_dd[CNT_OOB] > 0) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "slab decomposition: %d atoms left the box", c->h_dd[CNT_OOB]);
    } else {
        std::vector<int> tab(4 * R);
        for (mdb_ctx *p : c->dd_peers) {
            CUDA_TRY(c, cudaMemcpyAsync(tab.data() + 4 * p->dd_rank, p->dd_dev + 4 * p->dd_rank, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaMemcpyAsync(p->h_dd, p->counters, sizeof(int) * CNT__N, cudaMemcpyDeviceToHost, c->stream));
        }
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        for (mdb_ctx *p : c->dd_peers) {
            p->dd_tab = tab;
            if (p->h_dd[CNT_OOB] > 0) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "slab decomposition: %d atoms left the box", p->h_dd[CNT_OOB]);
        }
    }
    // (4) global slots: rank r owns [base_r, base_r + owned_r); place and permute my atoms, new ranges
    if ((rc = all_ranks(c, [&](mdb_ctx *p) {
            const std::vector<int> &t = p->dd_tab;
            long long tot = 0;
            std::vector<int> base(R + 1, 0);
            for (int r = 0; r < R; r++) { base[r + 1] = base[r] + t[4 * r]; tot += t[4 * r]; }
            if (tot != p->n) return mdb_fail(p, MDB_ERR_STATE, "slab decomposition: %lld of %d atoms found by the ranks (an atom moved more than a "
                                                               "cell layer between two rebuilds)", tot, p->n);
            int cand[6];
            cand_of(p, cand);
            const int r = p->dd_rank, below = (r - 1 + R) % R, above = (r + 1) % R;
            int rc2 = mdb_cells_dd_place(p, cand, p->dd_info[12] / cl, p->dd_info[13] / cl, base[r], t[4 * r]);
            if (rc2 < 0) return rc2;
            if (P2PState *S = p2p_of(p)) S->buf ^= 1; // the position buffers were swapped
            int *d = p->dd_info;
            d[0] = base[r]; d[1] = base[r + 1];
            d[2] = base[below + 1] - t[4 * below + 2]; d[3] = base[below + 1];   // top layer of the rank below
            d[4] = base[above]; d[5] = base[above] + t[4 * above + 1];           // bottom layer of the rank above
            d[6] = base[r]; d[7] = base[r] + t[4 * r + 1];                       // my bottom layer
            d[8] = base[r + 1] - t[4 * r + 2]; d[9] = base[r + 1];               // my top layer
            int mx = 0;
            for (int q = 0; q < R; q++) mx = std::max(mx, t[4 * q + 3]);
            p->mxnac = mx;
            return (int)MDB_OK;
        })) < 0) return rc;
    // (5) cell counts and atoms of the new boundary layers -> the neighbours' ghost layers
    {
        const int fc[2] = {DF_NAC, DF_NAAC};
        if ((rc = all_ranks(c, [&](mdb_ctx *p) { return dd_exchange(p, fc, 2, cell_ranges(p), false); })) < 0) return rc;
        if ((rc = all_ranks(c, [&](mdb_ctx *p) {
                const XRanges cr = cell_ranges(p);
                int rc2 = mdb_cells_dd_ghost_layer(p, cr.ga0, p->dd_info[4]);
                if (rc2 < 0) return rc2;
                return mdb_cells_dd_ghost_layer(p, cr.gb0, p->dd_info[2]);
            })) < 0) return rc;
        const int fa[4] = {DF_POS, DF_ITYP, DF_STATU, DF_GID};
        if ((rc = all_ranks(c, [&](mdb_ctx *p) { return dd_exchange(p, fa, 4, atom_ranges(p), false); })) < 0) return rc;
    }
    // (6) lists of my tiles; a capacity overflow cannot fall back to the generic path here
    if ((rc = all_ranks(c, [&](mdb_ctx *p) {
            p->indi_stale = true; p->list_reordered = false; p->list_gen++;
            int rc2 = mdb_tiled_nlist(p);
            if (rc2 < 0) return rc2;
            p->list_valid = true;
            return (int)cudaMemcpyAsync(p->h_counters, p->counters, sizeof(int) * CNT__N, cudaMemcpyDeviceToHost, p->stream) == cudaSuccess ? MDB_OK : MDB_ERR_CUDA;
        })) < 0) return rc;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return all_ranks(c, [&](mdb_ctx *p) {
        if (p->h_counters[CNT_TILE_OVERFLOW] > 0)
            return mdb_fail(p, MDB_ERR_UNSUPPORTED, "slab decomposition: %d tiles / cells exceed the halo, list or mxKVOIS capacity of the tiled path",
                            p->h_counters[CNT_TILE_OVERFLOW]);
        return (int)MDB_OK;
    });
}

// collective: (re)build cells, ranges, ghost layers and lists
extern "C" int mdb_dd_build(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->dd_on) return mdb_fail(c, MDB_ERR_STATE, "mdb_dd_build: mdb_dd_set first");
    if (!c->has_nlist || !c->has_tables) return mdb_fail(c, MDB_ERR_STATE, "mdb_dd_build: tables and list must be set");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    if (!c->dd_built || !c->list_valid) {
        int rc = all_ranks(c, [](mdb_ctx *p) { return dd_first_build(p); });
        if (rc < 0) return rc;
        return p2p_init(c); // (NCCL backend only; falls back to NCCL exchanges when IPC mappings are not available)
    }
    return dd_local_rebuild(c);
}

// ------------------------------------------------------------------------------------
// force, step loop, reductions
// ------------------------------------------------------------------------------------
static int dd_reduce_sum(mdb_ctx *c, double *v, int nv)
{
    // sum of nv host doubles over the ranks (virial tensor, kinetic energy): one all-reduce of <= 9 numbers
    if (!c->dd_peers.empty()) return MDB_OK; // in-process backend: the caller adds the ranks' values itself
    double *d = reinterpret_cast<double *>(c->dd_dev + 4 * c->dd_n + 8);
    CUDA_TRY(c, cudaMemcpyAsync(d, v, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(c, nccl().AllReduce(d, d, nv, ncclDouble, ncclSum, (ncclComm_t)c->dd_comm, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(v, d, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MDB_OK;
}

// pCalForce / pCalPTensor on the decomposed box: density pass, DEN ghost exchange, force pass (+ virial: partial tensors
// of the owned tiles added over the ranks, as the reference adds its per-device tensors on the host, :1434-1466)
extern "C" int mdb_dd_force(mdb_ctx *c, unsigned flags, double vtensor[9])
{
    if (!c) return MDB_ERR_ARG;
    if (!c->dd_on || !c->dd_built || !c->list_valid) return mdb_fail(c, MDB_ERR_STATE, "mdb_dd_force: mdb_dd_build first");
    if ((flags & MDB_VIRIAL) && !vtensor) return mdb_fail(c, MDB_ERR_ARG, "mdb_dd_force: MDB_VIRIAL needs vtensor");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    int rc;
    if ((rc = all_ranks(c, [](mdb_ctx *p) { return mdb_force_tiled(p, MDB_DEN); })) < 0) return rc;
    if ((rc = x_pos(c, false)) < 0) return rc;
    double sum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if ((rc = all_ranks(c, [&](mdb_ctx *p) {
            double vt[9];
            int rc2 = mdb_force(p, (flags & ~MDB_DEN) | MDB_NOPASS1, (flags & MDB_VIRIAL) ? vt : nullptr);
            if (rc2 < 0) return rc2;
            if (flags & MDB_VIRIAL) for (int i = 0; i < 9; i++) sum[i] += vt[i];
            return (int)MDB_OK;
        })) < 0) return rc;
    if (flags & MDB_VIRIAL) {
        if ((rc = dd_reduce_sum(c, sum, 9)) < 0) return rc;
        for (int i = 0; i < 9; i++) vtensor[i] = sum[i];
    }
    return MDB_OK;
}

// Predictor_DEV's halving loop on the decomposed box: every rank's mask of objecting trial steps, OR-ed over the ranks
int mdb_dd_timestep(mdb_ctx *c, double hmx, double dmx, double *h)
{
    int rc;
    const int R = c->dd_n;
    if (!c->dd_built) return mdb_fail(c, MDB_ERR_STATE, "mdb_timestep_limit: mdb_dd_build first");
    if ((rc = all_ranks(c, [&](mdb_ctx *p) { return mdb_timestep_mask_launch(p, hmx, dmx * dmx); })) < 0) return rc;
    unsigned mask = 0u;
    if (!c->dd_peers.empty()) {
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        for (mdb_ctx *p : c->dd_peers) mask |= (unsigned)p->h_counters[CNT_SCRATCH];
    } else {
        int *all = c->dd_dev + 4 * R + 64;
        NCCL_TRY(c, nccl().AllGather(c->counters + CNT_SCRATCH, all, 1, ncclInt, (ncclComm_t)c->dd_comm, c->stream));
        std::vector<int> tab(R);
        CUDA_TRY(c, cudaMemcpyAsync(tab.data(), all, sizeof(int) * R, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        for (int r = 0; r < R; r++) mask |= (unsigned)tab[r];
    }
    return mdb_timestep_from_mask(c, mask, hmx, h);
}

// nsteps x For_One_Step on the decomposed box, enqueued from here (no host round trip between the kernels and the exchanges
// of a step; the host waits only inside a rebuild and at a time-step check).  sch == NULL: fixed step and list period.
static int dd_run_impl(mdb_ctx *c, int itime0, int nsteps, int it0, int nb_fixed, double h_fixed, const mdb_sched *sch, double *h_io,
                       double *time_io)
{
    int rc;
    // Overlap (peer-to-peer backend, >= 3 layers per rank): the exchange runs on a side stream while the INTERIOR tiles --
    // whose halos hold owned atoms only -- are computed; the two boundary layers follow once the ghost layers have landed.
    P2PState *PS = p2p_of(c);
    const int cl = c->ncell[0] * c->ncell[1];
    const int nlay = (c->dd_info[13] - c->dd_info[12]) / cl, tpl = c->ncell[1] * c->tiled.ntx;
    const char *env = getenv("MDB_DD_OVERLAP");
    const bool overlap = PS && PS->on && nlay >= 3 && !(env && atoi(env) == 0);
    if (overlap && !c->dd_xs) {
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->dd_xs, cudaStreamNonBlocking));
        for (int k = 0; k < 4; k++) CUDA_TRY(c, cudaEventCreateWithFlags(&c->dd_ev[k], cudaEventDisableTiming));
    }
    // steps that close themselves (EPC friction, electronic stopping, corrector at their end): whenever the step size may
    // change from one step to the next or stopping acts between friction and corrector; else the two ride in front of the
    // next predictor
    const bool closed = sch != nullptr || mdb_stopping_on(c);
    all_ranks(c, [&](mdb_ctx *p) { p->var_step = sch && sch->ihdup < 0; return (int)MDB_OK; });
    const bool guard = mdb_tile_guard_wanted(c);
    struct Leave { mdb_ctx *c; ~Leave() { all_ranks(c, [](mdb_ctx *p) { p->var_step = false; p->tile_guard_fresh = false; return (int)MDB_OK; }); } } leave{c};
    const int t0 = c->dd_info[14], t1 = c->dd_info[15];
    auto select = [&](int a, int b, int a2, int b2) { c->tile_sel[0] = a; c->tile_sel[1] = b; c->tile_sel[2] = a2; c->tile_sel[3] = b2; };
    auto pass_split = [&](unsigned flags, cudaEvent_t landed, bool bounds) -> int { // interior, wait for the ghosts, the two boundary layers
        int r = MDB_OK;
        if (bounds) r = mdb_tile_guard_launch(c, t0 + tpl, t1 - tpl);
        select(t0 + tpl, t1 - tpl, 0, 0);
        if (r >= 0) r = mdb_force_tiled(c, flags);
        if (r >= 0) {
            if (cudaStreamWaitEvent(c->stream, landed, 0) != cudaSuccess) r = mdb_fail(c, MDB_ERR_CUDA, "cudaStreamWaitEvent failed");
            // (the neighbours' displacement bounds are merged by now: the boundary tiles read them)
            if (r >= 0 && bounds) r = mdb_tile_guard_launch(c, t0, t0 + tpl);
            if (r >= 0 && bounds) r = mdb_tile_guard_launch(c, t1 - tpl, t1);
            select(t0, t0 + tpl, t1 - tpl, t1);
            if (r >= 0) r = mdb_force_tiled(c, flags);
        }
        c->tile_sel[0] = -1;
        return r;
    };
    double h = h_io ? *h_io : h_fixed, t = time_io ? *time_io : 0.0;
    for (int s = 0; s < nsteps; s++) {
        const int itime = itime0 + s;
        int nb_uptab = nb_fixed;
        if (sch) {
            h = mdb_sched_h1(sch, itime, it0, h);
            if (mdb_sched_check_due(sch, itime, it0) && (rc = mdb_dd_timestep(c, sch->hmx, sch->dmx, &h)) < 0) return rc;
            nb_uptab = mdb_sched_nb_uptab(sch, itime, it0);
        }
        const int pre = (s == 0 || closed) ? 0 : 3; // EPC friction + corrector of the previous step ride in front of this predictor
        if ((rc = all_ranks(c, [&](mdb_ctx *p) { return mdb_predict_launch(p, h, pre); })) < 0) return rc;
        const bool rebuild = nb_uptab > 0 && (itime - it0) % nb_uptab == 0;
        if (overlap && !rebuild) {
            CUDA_TRY(c, cudaEventRecord(c->dd_ev[0], c->stream));
            CUDA_TRY(c, cudaStreamWaitEvent(c->dd_xs, c->dd_ev[0], 0));
            if ((rc = p2p_exchange(c, true, c->dd_xs)) < 0) return rc;
            CUDA_TRY(c, cudaEventRecord(c->dd_ev[1], c->dd_xs));
            if ((rc = pass_split(MDB_DEN, c->dd_ev[1], guard)) < 0) return rc;
            CUDA_TRY(c, cudaEventRecord(c->dd_ev[2], c->stream));
            CUDA_TRY(c, cudaStreamWaitEvent(c->dd_xs, c->dd_ev[2], 0));
            if ((rc = p2p_exchange(c, false, c->dd_xs)) < 0) return rc;
            CUDA_TRY(c, cudaEventRecord(c->dd_ev[3], c->dd_xs));
            if ((rc = pass_split(MDB_FORCE | MDB_NOPASS1, c->dd_ev[3], false)) < 0) return rc;
        } else {
            if ((rc = x_pos(c, !rebuild)) < 0) return rc;
            if (rebuild && (rc = dd_local_rebuild(c)) < 0) return rc;
            if (guard && !rebuild && (rc = all_ranks(c, [](mdb_ctx *p) { return mdb_tile_guard_launch(p, p->dd_info[14], p->dd_info[15]); })) < 0) return rc;
            if ((rc = all_ranks(c, [](mdb_ctx *p) { return mdb_force_tiled(p, MDB_DEN); })) < 0) return rc;
            if ((rc = x_pos(c, false)) < 0) return rc;
            if ((rc = all_ranks(c, [](mdb_ctx *p) { return mdb_force_tiled(p, MDB_FORCE | MDB_NOPASS1); })) < 0) return rc;
        }
        all_ranks(c, [](mdb_ctx *p) { p->tile_guard_fresh = false; return (int)MDB_OK; }); // the bounds were this step's
        if (closed && (rc = all_ranks(c, [&](mdb_ctx *p) { return mdb_step_close_launch(p, h); })) < 0) return rc;
        t += h;
    }
    if (!closed && nsteps > 0 && (rc = all_ranks(c, [&](mdb_ctx *p) { return mdb_epc_correct_launch(p, h); })) < 0) return rc;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (h_io) *h_io = h;
    if (time_io) *time_io = t;
    return p2p_check(c);
}

extern "C" int mdb_dd_run(mdb_ctx *c, int itime0, int nsteps, int it0, int nb_uptab, double h)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->dd_on || !c->dd_built || !c->list_valid) return mdb_fail(c, MDB_ERR_STATE, "mdb_dd_run: mdb_dd_build first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    return dd_run_impl(c, itime0, nsteps, it0, nb_uptab, h, nullptr, nullptr, nullptr);
}

extern "C" int mdb_dd_run_sched(mdb_ctx *c, int itime0, int nsteps, int it0, const mdb_sched *s, double *h, double *time_s)
{
    if (!c || !s || !h) return mdb_fail(c, MDB_ERR_ARG, "mdb_dd_run_sched: null argument");
    if (!c->dd_on || !c->dd_built || !c->list_valid) return mdb_fail(c, MDB_ERR_STATE, "mdb_dd_run_sched: mdb_dd_build first");
    if (s->nb_uptabmi < 1 || (s->ihdup != 0 && !(s->hmx > 0.0))) return mdb_fail(c, MDB_ERR_ARG, "mdb_dd_run_sched: bad schedule");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    return dd_run_impl(c, itime0, nsteps, it0, 0, *h, s, h, time_s);
}

// Cal_GlobalT_DEV (CommonGPU/MD_DiffScheme_GPU.F90:1042-1064) on the decomposed box: EKIN of the owned atoms of every rank
__global__ void k_dd_ekin_owned(int n, int a0, int a1, const double *__restrict__ xp1, const int *__restrict__ statu,
                                const int *__restrict__ ityp, MassParams M, double *__restrict__ out2)
{
    __shared__ double sh[32];
    __shared__ double shc[32];
    double s = 0.0, cn = 0.0;
    for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x) {
        const int st = statu[i];
        if ((st & ST_ACTIVE) == ST_ACTIVE && (st & ST_FIXPOS) == 0) {
            const double cm0 = M.cm[ityp[i] - 1];
            const double vx = xp1[i], vy = xp1[i + (size_t)n], vz = xp1[i + 2 * (size_t)n];
            const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
            s += __dmul_rn(__dmul_rn(0.5, cm0), v2);
            cn += 1.0;
        }
    }
    for (int off = 16; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); cn += __shfl_xor_sync(0xffffffffu, cn, off); }
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = s; shc[threadIdx.x >> 5] = cn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, q = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { t += sh[w]; q += shc[w]; }
        out2[0] = t; out2[1] = q;
    }
}
extern "C" int mdb_dd_global_t(mdb_ctx *c, double *curt)
{
    if (!c || !curt) return mdb_fail(c, MDB_ERR_ARG, "mdb_dd_global_t: null argument");
    if (!c->dd_on || !c->dd_built) return mdb_fail(c, MDB_ERR_STATE, "mdb_dd_global_t: mdb_dd_build first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    double acc[2] = {0.0, 0.0};
    int rc = all_ranks(c, [&](mdb_ctx *p) {
        double *d = reinterpret_cast<double *>(p->dd_dev + 4 * p->dd_n + 32);
        double h2[2];
        k_dd_ekin_owned<<<1, 1024, 0, p->stream>>>(p->n, p->dd_info[0], p->dd_info[1], p->xp1, p->statu, p->ityp, p->mass, d);
        if (cudaMemcpyAsync(h2, d, sizeof(h2), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess || cudaStreamSynchronize(p->stream) != cudaSuccess)
            return mdb_fail(p, MDB_ERR_CUDA, "mdb_dd_global_t: copy failed");
        acc[0] += h2[0]; acc[1] += h2[1];
        return (int)MDB_OK;
    });
    if (rc < 0) return rc;
    if ((rc = dd_reduce_sum(c, acc, 2)) < 0) return rc;
    *curt = 2.0 * acc[0] / acc[1] / (3.0 * KB_CGS); // C_TWO*sum(hm_EKIN, mask)/count/(C_THR*CP_KB) :1062
    return MDB_OK;
}

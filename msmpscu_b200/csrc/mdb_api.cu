// mdb_api.cu -- context, device box, host<->device field transfer, tables, profiling.
// C ABI declared in include/mdpscu_b200.h (each entry cites the reference interface it replaces).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include "mdb_internal.cuh"

// ------------------------------------------------------------------------------------
// errors / profiling
// ------------------------------------------------------------------------------------
int mdb_fail(mdb_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

ProfScope::ProfScope(mdb_ctx *ctx, int klass, int nlaunch) : c(ctx), k(klass)
{
    c->launches_total += nlaunch;
    c->prof_launches[k] += nlaunch;
    if (!c->prof) return;
    auto get = [&]() {
        cudaEvent_t e;
        if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    };
    a = get();
    b = get();
    cudaEventRecord(a, c->stream);
}
ProfScope::~ProfScope()
{
    if (!a) return;
    cudaEventRecord(b, c->stream);
    c->ev_pending.push_back({a, b, k});
    if (c->ev_pending.size() > 4096) mdb_prof_collect(c);
}
void mdb_prof_collect(mdb_ctx *c)
{
    if (c->ev_pending.empty()) return;
    cudaEventSynchronize(c->ev_pending.back().b);
    for (auto &e : c->ev_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) c->prof_ms[e.k] += ms;
        c->ev_pool.push_back(e.a);
        c->ev_pool.push_back(e.b);
    }
    c->ev_pending.clear();
}

extern "C" int mdb_prof_enable(mdb_ctx *c, int on)
{
    if (!c) return MDB_ERR_ARG;
    mdb_prof_collect(c);
    c->prof = on != 0;
    return MDB_OK;
}
extern "C" int mdb_prof_reset(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    mdb_prof_collect(c);
    for (int k = 0; k < MDB_K__COUNT; k++) { c->prof_launches[k] = 0; c->prof_ms[k] = 0.0; }
    c->launches_total = 0;
    return MDB_OK;
}
extern "C" int mdb_prof_get(mdb_ctx *c, long long launches[MDB_K__COUNT], double ms[MDB_K__COUNT])
{
    if (!c) return MDB_ERR_ARG;
    mdb_prof_collect(c);
    for (int k = 0; k < MDB_K__COUNT; k++) { launches[k] = c->prof_launches[k]; ms[k] = c->prof_ms[k]; }
    return MDB_OK;
}
extern "C" long long mdb_launch_count(const mdb_ctx *c) { return c ? c->launches_total : 0; }

// ------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------
extern "C" const char *mdb_version(void) { return "mdpscu_b200 0.1 (sm_100a)"; }

extern "C" int mdb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int mdb_ctx_create(int device_id, mdb_ctx **out)
{
    if (!out) return MDB_ERR_ARG;
    *out = nullptr;
    int n = mdb_device_count();
    if (n <= 0) return MDB_ERR_NOGPU; // no CPU fallback, by design
    if (device_id < 0 || device_id >= n) return MDB_ERR_ARG;
    if (cudaSetDevice(device_id) != cudaSuccess) return MDB_ERR_CUDA;
    mdb_ctx *c = new mdb_ctx();
    c->dev = device_id;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return MDB_ERR_CUDA; }
    c->stream = c->own_stream;
    memset(&c->epc, 0, sizeof(c->epc));
    memset(&c->tab, 0, sizeof(c->tab));
    for (int k = 0; k < MDB_K__COUNT; k++) { c->prof_launches[k] = 0; c->prof_ms[k] = 0.0; }
    if (cudaMalloc(&c->counters, sizeof(int) * CNT__N) != cudaSuccess ||
        cudaMallocHost(&c->h_counters, sizeof(int) * CNT__N) != cudaSuccess) { delete c; return MDB_ERR_NOMEM; }
    cudaMemset(c->counters, 0, sizeof(int) * CNT__N);
    *out = c;
    return MDB_OK;
}

template <class T> static void dfree(T *&p) { if (p) { cudaFree(p); p = nullptr; } }

static void free_state(mdb_ctx *c)
{
    dfree(c->pos); dfree(c->pos_alt); dfree(c->xp1); dfree(c->xp1_alt); dfree(c->fp); dfree(c->fp_alt);
    dfree(c->dis); dfree(c->dis_alt); dfree(c->epot); dfree(c->ekin); dfree(c->ityp); dfree(c->ityp_alt);
    dfree(c->statu); dfree(c->statu_alt); dfree(c->gid); dfree(c->gid_alt); dfree(c->gidinv);
    dfree(c->ic); dfree(c->ic_alt); dfree(c->xp_view); dfree(c->den_view);
    dfree(c->slot); dfree(c->srcof); dfree(c->tmp_orig); dfree(c->oob); dfree(c->vpart);
    dfree(c->dsr); c->dsr_bytes = 0;
    dfree(c->dmax_blk); c->dmax_blk_bytes = 0;
    dfree(c->tile_d2); c->tile_d2_bytes = 0; c->tile_guard_fresh = false;
    dfree(c->pos_snap); c->pos_snap_bytes = 0;
    if (c->stage) { cudaFree(c->stage); c->stage = nullptr; c->stage_bytes = 0; }
    c->has_box = false;
}
static void free_nlist(mdb_ctx *c)
{
    mdb_tiled_free(c);
    dfree(c->nac); dfree(c->naac); dfree(c->ia1th); dfree(c->kvois); dfree(c->indi);
    dfree(c->scan_tmp); c->scan_n = 0;
    c->has_nlist = false; c->list_valid = false;
}
static void free_tables(mdb_ctx *c)
{
    for (void *p : c->tab_allocs) cudaFree(p);
    c->tab_allocs.clear();
    c->has_tables = false;
}

extern "C" void mdb_ctx_destroy(mdb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->dev);
    cudaStreamSynchronize(c->stream);
    mdb_dd_free(c);
    mdb_stopping_free(c);
    mdb_save_free(c);
    if (c->scratch) { cudaFree(c->scratch); c->scratch = nullptr; c->scratch_bytes = 0; }
    mdb_prof_collect(c);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    free_state(c); free_nlist(c); free_tables(c);
    if (c->counters) cudaFree(c->counters);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_dd) cudaFreeHost(c->h_dd);
    if (c->hstage) cudaFreeHost(c->hstage);
    if (c->avp) cudaFree(c->avp);
    if (c->lb_buf) cudaFree(c->lb_buf);
    if (c->lb_mask) cudaFree(c->lb_mask);
    if (c->lb_host) cudaFreeHost(c->lb_host);
    if (c->q_buf) cudaFree(c->q_buf);
    if (c->q_host) cudaFreeHost(c->q_host);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int mdb_set_option(mdb_ctx *c, int option, int value)
{
    if (!c) return MDB_ERR_ARG;
    if (option == MDB_OPT_TILED_LANES && (value == 0 || value == 2 || value == 4 || value == 8)) {
        c->tiled.G_opt = value; if (value) c->tiled.G = value;
        c->tiled.dirty = true; c->list_valid = false;
        return MDB_OK;
    }
    if (option == MDB_OPT_TILED_THREADS && (value == 512 || value == 768)) {
        c->tiled.threads_opt = value; c->tiled.dirty = true; c->list_valid = false;
        return MDB_OK;
    }
    if (option == MDB_OPT_TILED_STAGES && (value == 2 || value == 3)) {
        c->tiled.stages_opt = value; c->tiled.dirty = true; c->list_valid = false;
        return MDB_OK;
    }
    if (option == MDB_OPT_FUSE_EPILOGUE && (value == 0 || value == 1)) {
        c->opt_fuse_epilogue = value;
        return MDB_OK;
    }
    if (option == MDB_OPT_TILED_CLASSES && (value == 0 || value == 1)) {
        c->tiled.use_classes = value != 0;
        return MDB_OK;
    }
    if (option == MDB_OPT_TILED_BANKORDER && value >= -1 && value <= 1) {
        c->tiled.bank_order_opt = value; c->tiled.dirty = true; c->list_valid = false;
        return MDB_OK;
    }
    if (option == MDB_OPT_PDL && (value == 0 || value == 1)) {
        c->opt_pdl = value;
        return MDB_OK;
    }
    if (option == MDB_OPT_TILE_GUARD && value >= -1 && value <= 1) {
        c->opt_tile_guard = value;
        return MDB_OK;
    }
    if (option == MDB_OPT_FORCE_PATH && value >= MDB_FORCE_PATH_AUTO && value <= MDB_FORCE_PATH_TILED) {
        c->opt_force_path = value;
        c->list_valid = false; // the two paths keep different list formats
        c->tiled.dirty = true;
        return MDB_OK;
    }
    return mdb_fail(c, MDB_ERR_ARG, "mdb_set_option: unknown option %d / value %d", option, value);
}
extern "C" int mdb_get_option(const mdb_ctx *c, int option)
{
    if (!c) return MDB_ERR_ARG;
    if (option == MDB_OPT_FORCE_PATH) return c->opt_force_path;
    if (option == MDB_OPT_TILED_LANES) return c->tiled.G;
    if (option == MDB_OPT_TILED_THREADS) return c->tiled.threads_opt;
    if (option == MDB_OPT_TILED_STAGES) return c->tiled.stages_opt;
    if (option == MDB_OPT_FUSE_EPILOGUE) return c->opt_fuse_epilogue;
    if (option == MDB_OPT_TILED_CLASSES) return c->tiled.use_classes ? 1 : 0;
    if (option == MDB_OPT_TILED_BANKORDER) return c->tiled.bank_order ? 1 : 0;
    if (option == MDB_OPT_TILE_GUARD) return c->opt_tile_guard;
    if (option == MDB_OPT_PDL) return c->opt_pdl;
    if (option == MDB_OPT_ACTIVE_PATH) return c->tiled.active ? MDB_FORCE_PATH_TILED : MDB_FORCE_PATH_GENERIC;
    return MDB_ERR_ARG;
}

extern "C" const char *mdb_last_error(const mdb_ctx *c) { return c ? c->err.c_str() : "null context"; }
extern "C" int mdb_ctx_set_stream(mdb_ctx *c, void *s)
{
    if (!c) return MDB_ERR_ARG;
    cudaStreamSynchronize(c->stream);
    mdb_prof_collect(c);
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return MDB_OK;
}
extern "C" void *mdb_ctx_stream(mdb_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int mdb_sync(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->dev));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->stage_off = 0; // every staged transfer has landed
    if (c->run_pending) { // mdb_run_async: the counters were copied at the end of the block
        c->run_pending = false;
        return c->h_counters[CNT_OOB_TOTAL];
    }
    return MDB_OK;
}

// ------------------------------------------------------------------------------------
// box
// ------------------------------------------------------------------------------------
__global__ void k_set_int(int *p, int v) { *p = v; }

// positions were changed by the host (upload): the displacement-since-rebuild bound is unknown,
// so the tiled passes must scan whole lists until the next rebuild
void mdb_mark_positions_dirty(mdb_ctx *c)
{
    if (c->counters) k_set_int<<<1, 1, 0, c->stream>>>(c->counters + CNT_D2MAX, 0x7f800000);
}

__global__ void k_iota(int n, int *a, int *b)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = i + 1; b[i] = i + 1; }
}

extern "C" int mdb_box_set(mdb_ctx *c, int nbox, int napb, const double boxlow[3], const double boxsize[3],
                           const double boxshape[9], const int ifpd[3], int ngroup, const double mass[])
{
    if (!c || nbox < 1 || napb < 1 || ngroup < 1 || ngroup > MDB_MXGROUP) return mdb_fail(c, MDB_ERR_ARG, "mdb_box_set: bad argument");
    if ((long long)nbox * napb > 2000000000LL) return mdb_fail(c, MDB_ERR_ARG, "mdb_box_set: too many atoms for int32 ids");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    int n = nbox * napb;
    bool realloc = !c->has_box || n != c->n;
    c->nbox = nbox; c->napb = napb; c->ng = ngroup;
    for (int d = 0; d < 3; d++) {
        c->box.lo[d] = boxlow[d];
        c->box.size[d] = boxsize[d];
        c->box.up[d] = boxlow[d] + boxsize[d]; // BOXUP = BOXLOW + ZL, Common/MD_Gvar.F90:927
        c->box.half[d] = boxsize[d] * 0.5;     // HBX = BX*C_HALF
        c->box.pd[d] = ifpd[d];
    }
    c->shape_identity = true;
    for (int i = 0; i < 9; i++) {
        c->boxshape[i] = boxshape ? boxshape[i] : ((i % 4 == 0) ? 1.0 : 0.0);
        if (c->boxshape[i] != ((i % 4 == 0) ? 1.0 : 0.0)) c->shape_identity = false;
    }
    for (int g = 0; g < ngroup; g++) c->mass.cm[g] = mass[g];
    if (realloc) {
        free_state(c);
        free_nlist(c);
        c->n = n;
        size_t n3 = (size_t)n * 3;
#define ALLOC(p, T, cnt) CUDA_TRY(c, cudaMalloc(&c->p, sizeof(T) * (cnt)))
        ALLOC(pos, double4, n); ALLOC(pos_alt, double4, n);
        ALLOC(xp1, double, n3); ALLOC(xp1_alt, double, n3);
        ALLOC(fp, double, n3); ALLOC(fp_alt, double, n3);
        ALLOC(dis, double, n3); ALLOC(dis_alt, double, n3);
        ALLOC(epot, double, n); ALLOC(ekin, double, n);
        ALLOC(ityp, int, n + 8); ALLOC(ityp_alt, int, n + 8); // +8: the tiled passes read ITYP in aligned int4 windows
        ALLOC(statu, int, n + 8); ALLOC(statu_alt, int, n + 8); // +8: 16-byte aligned TMA windows may overshoot
        ALLOC(gid, int, n); ALLOC(gid_alt, int, n); ALLOC(gidinv, int, n);
        ALLOC(ic, int, n); ALLOC(ic_alt, int, n);
        ALLOC(slot, int, n); ALLOC(srcof, int, n); ALLOC(tmp_orig, int, n); ALLOC(oob, int, n);
#undef ALLOC
        CUDA_TRY(c, cudaMemsetAsync(c->pos, 0, sizeof(double4) * n, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->xp1, 0, sizeof(double) * n3, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->fp, 0, sizeof(double) * n3, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->dis, 0, sizeof(double) * n3, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->epot, 0, sizeof(double) * n, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->ekin, 0, sizeof(double) * n, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->ityp, 0, sizeof(int) * n, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->statu, 0, sizeof(int) * n, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->ic, 0, sizeof(int) * n, c->stream));
        k_iota<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->gid, c->gidinv);
        CUDA_TRY(c, cudaGetLastError());
        c->oob_total = 0;
    }
    c->has_box = true;
    c->list_valid = false;
    c->tiled.dirty = true;
    return MDB_OK;
}

extern "C" int mdb_natom(const mdb_ctx *c) { return c ? c->n : 0; }

// ------------------------------------------------------------------------------------
// field transfer
// ------------------------------------------------------------------------------------
// map == nullptr : CELL order, identity.  map = gidinv : host index o (ORIGINAL) <-> device map[o]-1
__global__ void k_up_d(int n, int ncol, const double *__restrict__ src, double *__restrict__ dst,
                       const int *__restrict__ map)
{
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    int d = map ? map[o] - 1 : o;
    for (int cc = 0; cc < ncol; cc++) dst[d + (size_t)cc * n] = src[o + (size_t)cc * n];
}
__global__ void k_down_d(int n, int ncol, const double *__restrict__ src, double *__restrict__ dst,
                         const int *__restrict__ map)
{
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    int d = map ? map[o] - 1 : o;
    for (int cc = 0; cc < ncol; cc++) dst[o + (size_t)cc * n] = src[d + (size_t)cc * n];
}
__global__ void k_up_i(int n, const int *__restrict__ src, int *__restrict__ dst, const int *__restrict__ map)
{
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    dst[map ? map[o] - 1 : o] = src[o];
}
__global__ void k_down_i(int n, const int *__restrict__ src, int *__restrict__ dst, const int *__restrict__ map)
{
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    dst[o] = src[map ? map[o] - 1 : o];
}
// pos {x,y,z,den}: which = 0 -> xyz (3 cols), 1 -> den (1 col)
__global__ void k_up_pos(int n, int which, const double *__restrict__ src, double4 *__restrict__ pos,
                         const int *__restrict__ map)
{
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    int d = map ? map[o] - 1 : o;
    double4 p = pos[d];
    if (which == 0) { p.x = src[o]; p.y = src[o + (size_t)n]; p.z = src[o + 2 * (size_t)n]; }
    else p.w = src[o];
    pos[d] = p;
}
__global__ void k_down_pos(int n, int which, const double4 *__restrict__ pos, double *__restrict__ dst,
                           const int *__restrict__ map)
{
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    double4 p = pos[map ? map[o] - 1 : o];
    if (which == 0) { dst[o] = p.x; dst[o + (size_t)n] = p.y; dst[o + 2 * (size_t)n] = p.z; }
    else dst[o] = p.w;
}

// Device staging for field transfers: a bump-allocated pool, so that several asynchronous transfers can be in flight on
// the stream at once (mdb_state_download_async).  A slot stays valid until the next mdb_sync / synchronous transfer,
// which rewinds the pool.  Growing the pool synchronises the stream first (outstanding copies finish), once.
static int stage_slot(mdb_ctx *c, size_t bytes, void **out)
{
    bytes = (bytes + 255) & ~(size_t)255;
    if (c->stage_off + bytes > c->stage_bytes) {
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        c->stage_off = 0;
        if (bytes > c->stage_bytes) {
            if (c->stage) cudaFree(c->stage);
            c->stage = nullptr; c->stage_bytes = 0;
            CUDA_TRY(c, cudaMalloc(&c->stage, bytes));
            c->stage_bytes = bytes;
        }
    }
    *out = (char *)c->stage + c->stage_off;
    c->stage_off += bytes;
    return MDB_OK;
}
// room for `count` more slots of `bytes` each without a later regrow (one synchronisation now instead of one per call)
static int stage_reserve(mdb_ctx *c, size_t bytes)
{
    if (c->stage_bytes >= bytes) return MDB_OK;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stage_bytes = 0; c->stage_off = 0;
    CUDA_TRY(c, cudaMalloc(&c->stage, bytes));
    c->stage_bytes = bytes;
    return MDB_OK;
}

struct FieldInfo { int ncol; bool is_int; bool per_cell; };
static bool field_info(int f, FieldInfo &fi)
{
    switch (f) {
    case MDB_F_XP: case MDB_F_XP1: case MDB_F_FP: case MDB_F_DIS: fi = {3, false, false}; return true;
    case MDB_F_EPOT: case MDB_F_EKIN: case MDB_F_DEN: fi = {1, false, false}; return true;
    case MDB_F_ITYP: case MDB_F_STATU: case MDB_F_GID: case MDB_F_GIDINV: case MDB_F_IC: case MDB_F_KVOIS:
        fi = {1, true, false}; return true;
    case MDB_F_NAC: case MDB_F_NAAC: case MDB_F_IA1TH: fi = {1, true, true}; return true;
    }
    return false;
}

static double *dptr_d(mdb_ctx *c, int f)
{
    switch (f) {
    case MDB_F_XP1: return c->xp1; case MDB_F_FP: return c->fp; case MDB_F_DIS: return c->dis;
    case MDB_F_EPOT: return c->epot; case MDB_F_EKIN: return c->ekin;
    }
    return nullptr;
}
static int *dptr_i(mdb_ctx *c, int f)
{
    switch (f) {
    case MDB_F_ITYP: return c->ityp; case MDB_F_STATU: return c->statu; case MDB_F_GID: return c->gid;
    case MDB_F_GIDINV: return c->gidinv; case MDB_F_IC: return c->ic; case MDB_F_KVOIS: return c->kvois;
    case MDB_F_NAC: return c->nac; case MDB_F_NAAC: return c->naac; case MDB_F_IA1TH: return c->ia1th;
    }
    return nullptr;
}

extern "C" int mdb_state_upload(mdb_ctx *c, int field, const void *host, int order)
{
    if (!c || !host) return mdb_fail(c, MDB_ERR_ARG, "mdb_state_upload: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_state_upload: mdb_box_set first");
    FieldInfo fi;
    if (!field_info(field, fi) || fi.per_cell || field == MDB_F_GID || field == MDB_F_GIDINV || field == MDB_F_KVOIS ||
        field == MDB_F_IC)
        return mdb_fail(c, MDB_ERR_ARG, "mdb_state_upload: field %d is not uploadable", field);
    CUDA_TRY(c, cudaSetDevice(c->dev));
    int n = c->n;
    size_t bytes = (size_t)n * fi.ncol * (fi.is_int ? sizeof(int) : sizeof(double));
    // room for a whole CopyIn/CopyOut set (XP, XP1, FP, DIS + integers) without regrowing between the calls
    int rc = stage_reserve(c, (size_t)n * 128);
    if (rc) return rc;
    void *stg = nullptr;
    if ((rc = stage_slot(c, bytes, &stg))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(stg, host, bytes, cudaMemcpyHostToDevice, c->stream));
    const int *map = (order == MDB_ORDER_ORIGINAL) ? c->gidinv : nullptr;
    int nb = cdiv(n, 256);
    ProfScope ps(c, MDB_K_OTHER);
    if (field == MDB_F_XP) k_up_pos<<<nb, 256, 0, c->stream>>>(n, 0, (const double *)stg, c->pos, map);
    else if (field == MDB_F_DEN) k_up_pos<<<nb, 256, 0, c->stream>>>(n, 1, (const double *)stg, c->pos, map);
    else if (fi.is_int) k_up_i<<<nb, 256, 0, c->stream>>>(n, (const int *)stg, dptr_i(c, field), map);
    else k_up_d<<<nb, 256, 0, c->stream>>>(n, fi.ncol, (const double *)stg, dptr_d(c, field), map);
    CUDA_TRY(c, cudaGetLastError());
    if (field == MDB_F_XP) mdb_mark_positions_dirty(c);
    if (field == MDB_F_ITYP) c->list_valid = false;
    return MDB_OK;
}

static int state_download(mdb_ctx *c, int field, void *host, int order, bool sync)
{
    if (!c || !host) return mdb_fail(c, MDB_ERR_ARG, "mdb_state_download: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_state_download: mdb_box_set first");
    FieldInfo fi;
    if (!field_info(field, fi)) return mdb_fail(c, MDB_ERR_ARG, "mdb_state_download: bad field %d", field);
    CUDA_TRY(c, cudaSetDevice(c->dev));
    if (fi.per_cell) {
        if (!c->has_nlist) return mdb_fail(c, MDB_ERR_STATE, "mdb_state_download: no cell data yet");
        CUDA_TRY(c, cudaMemcpyAsync(host, dptr_i(c, field), sizeof(int) * (size_t)c->nc, cudaMemcpyDeviceToHost, c->stream));
        if (sync) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        return MDB_OK;
    }
    if (field == MDB_F_KVOIS && !c->kvois) return mdb_fail(c, MDB_ERR_STATE, "mdb_state_download: no list yet");
    int n = c->n;
    size_t bytes = (size_t)n * fi.ncol * (fi.is_int ? sizeof(int) : sizeof(double));
    int rc = stage_reserve(c, (size_t)n * 128);
    if (rc) return rc;
    void *stg = nullptr;
    if ((rc = stage_slot(c, bytes, &stg))) return rc;
    bool id_only = (field == MDB_F_GID || field == MDB_F_GIDINV);
    const int *map = (order == MDB_ORDER_ORIGINAL && !id_only) ? c->gidinv : nullptr;
    int nb = cdiv(n, 256);
    {
        ProfScope ps(c, MDB_K_OTHER);
        if (field == MDB_F_XP) k_down_pos<<<nb, 256, 0, c->stream>>>(n, 0, c->pos, (double *)stg, map);
        else if (field == MDB_F_DEN) k_down_pos<<<nb, 256, 0, c->stream>>>(n, 1, c->pos, (double *)stg, map);
        else if (fi.is_int) k_down_i<<<nb, 256, 0, c->stream>>>(n, dptr_i(c, field), (int *)stg, map);
        else k_down_d<<<nb, 256, 0, c->stream>>>(n, fi.ncol, dptr_d(c, field), (double *)stg, map);
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(host, stg, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (sync) {
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        c->stage_off = 0;
    }
    return MDB_OK;
}

extern "C" int mdb_state_download(mdb_ctx *c, int field, void *host, int order)
{
    return state_download(c, field, host, order, true);
}
// CopyOut without the host synchronisation: the permutation kernel and the device-to-host copy are enqueued on the
// context's stream; `host` (page-locked memory for a truly asynchronous copy) holds the field after the next mdb_sync.
extern "C" int mdb_state_download_async(mdb_ctx *c, int field, void *host, int order)
{
    return state_download(c, field, host, order, false);
}

// host-side convenience over mdb_atomic_stress: AP(N,9) column-major in the requested order
extern "C" int mdb_atomic_stress_host(mdb_ctx *c, double *h_avp, int order)
{
    if (!c || !h_avp) return mdb_fail(c, MDB_ERR_ARG, "mdb_atomic_stress_host: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_atomic_stress_host: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    const int n = c->n;
    if (c->avp_n != n) {
        if (c->avp) cudaFree(c->avp);
        c->avp = nullptr; c->avp_n = 0;
        CUDA_TRY(c, cudaMalloc(&c->avp, sizeof(double) * 9 * (size_t)n));
        c->avp_n = n;
    }
    int rc = mdb_atomic_stress(c, c->avp);
    if (rc < 0) return rc;
    const size_t bytes = sizeof(double) * 9 * (size_t)n;
    void *stg = nullptr;
    if ((rc = stage_slot(c, bytes, &stg))) return rc;
    {
        ProfScope ps(c, MDB_K_OTHER);
        k_down_d<<<cdiv(n, 256), 256, 0, c->stream>>>(n, 9, c->avp, (double *)stg, order == MDB_ORDER_ORIGINAL ? c->gidinv : nullptr);
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(h_avp, stg, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->stage_off = 0;
    return MDB_OK;
}

extern "C" void *mdb_devptr(mdb_ctx *c, int field)
{
    if (!c || !c->has_box) return nullptr;
    cudaSetDevice(c->dev);
    int n = c->n, nb = cdiv(n, 256);
    switch (field) {
    case MDB_F_XP:
        if (!c->xp_view && cudaMalloc(&c->xp_view, sizeof(double) * 3 * (size_t)n) != cudaSuccess) return nullptr;
        k_down_pos<<<nb, 256, 0, c->stream>>>(n, 0, c->pos, c->xp_view, nullptr);
        c->launches_total++;
        return c->xp_view;
    case MDB_F_DEN:
        if (!c->den_view && cudaMalloc(&c->den_view, sizeof(double) * (size_t)n) != cudaSuccess) return nullptr;
        k_down_pos<<<nb, 256, 0, c->stream>>>(n, 1, c->pos, c->den_view, nullptr);
        c->launches_total++;
        return c->den_view;
    case MDB_F_INDI: return mdb_indi_ensure(c) == MDB_OK ? c->indi : nullptr;
    case MDB_F_POS4: return c->pos;                      // internal {x,y,z,den} records (ghost exchange)
    case MDB_F_D2MAX: return c->counters + CNT_D2MAX;    // max displacement^2 since the rebuild (float bits)
    }
    if (double *p = dptr_d(c, field)) return p;
    return dptr_i(c, field);
}

// ------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------
// pack T(NKIND,NTAB) Fortran layout -> per kind (ntab+2) entries {T[KK], T[KK+1]-T[KK]}, KK = 0..ntab+1,
// with T[0] = T[ntab+1] = 0 : the reference reads those two out of bounds
// (CommonGPU/MD_EAM_ForceTable_GPU.F90:522-530 at r == Rmax); the oracle defines them as 0 and so do we.
static int pack_table(mdb_ctx *c, const double *t, int nkind, int ntab, double2 **out)
{
    size_t cnt = (size_t)nkind * (ntab + 2);
    std::vector<double2> h(cnt);
    for (int k = 0; k < nkind; k++)
        for (int kk = 0; kk <= ntab + 1; kk++) {
            double a = (kk >= 1 && kk <= ntab) ? t[(size_t)(kk - 1) * nkind + k] : 0.0;
            double b = (kk + 1 >= 1 && kk + 1 <= ntab) ? t[(size_t)kk * nkind + k] : 0.0;
            h[(size_t)k * (ntab + 2) + kk] = make_double2(a, b - a);
        }
    void *d = nullptr;
    CUDA_TRY(c, cudaMalloc(&d, sizeof(double2) * cnt));
    c->tab_allocs.push_back(d);
    CUDA_TRY(c, cudaMemcpy(d, h.data(), sizeof(double2) * cnt, cudaMemcpyHostToDevice));
    *out = (double2 *)d;
    return MDB_OK;
}

extern "C" int mdb_tables_set(mdb_ctx *c, int pot_type, int nkind, int ntab, double csi, const double *potr,
                              const double *fpotr, const double *potb, const double *fpotb, int nkind1, int nembd,
                              double rhod, const double *fembd, const double *dfembd, const int *kpair,
                              const int *kembd, double ru2max)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_tables_set: mdb_box_set first");
    if (nkind < 1 || ntab < 2 || !potr || !fpotr || !potb || !fpotb || !kpair)
        return mdb_fail(c, MDB_ERR_ARG, "mdb_tables_set: bad pair-table argument");
    if (pot_type == MDB_POT_EAM && (nkind1 < 1 || nembd < 2 || !fembd || !dfembd || !kembd || !(rhod > 0.0)))
        return mdb_fail(c, MDB_ERR_ARG, "mdb_tables_set: bad embedding-table argument");
    if (pot_type != MDB_POT_EAM && pot_type != MDB_POT_FS) return mdb_fail(c, MDB_ERR_ARG, "mdb_tables_set: pot_type");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    free_tables(c);
    TableSet &t = c->tab;
    memset(&t, 0, sizeof(t));
    t.pot_type = pot_type; t.ng = c->ng; t.nkind = nkind; t.ntab = ntab; t.csi = csi; t.ru2max = ru2max;
    t.nkind1 = nkind1; t.nembd = nembd; t.rhod = rhod;
    int rc;
    if ((rc = pack_table(c, potr, nkind, ntab, &t.potr))) return rc;
    if ((rc = pack_table(c, fpotr, nkind, ntab, &t.fpotr))) return rc;
    if ((rc = pack_table(c, potb, nkind, ntab, &t.potb))) return rc;
    if ((rc = pack_table(c, fpotb, nkind, ntab, &t.fpotb))) return rc;
    if (pot_type == MDB_POT_EAM) {
        if ((rc = pack_table(c, fembd, nkind1, nembd, &t.fembd))) return rc;
        if ((rc = pack_table(c, dfembd, nkind1, nembd, &t.dfembd))) return rc;
    }
    for (int i = 0; i < c->ng; i++) {
        for (int j = 0; j < c->ng; j++) {
            int k = kpair[i + c->ng * j];
            if (k < 1 || k > nkind) return mdb_fail(c, MDB_ERR_ARG, "mdb_tables_set: KPAIR(%d,%d)=%d out of range", i + 1, j + 1, k);
            t.kpair[i + c->ng * j] = k - 1;
        }
        if (pot_type == MDB_POT_EAM) {
            int k = kembd[i];
            if (k < 1 || k > nkind1) return mdb_fail(c, MDB_ERR_ARG, "mdb_tables_set: KEMBD(%d)=%d out of range", i + 1, k);
            t.kembd[i] = k - 1;
        }
    }
    c->h_potb.assign(potb, potb + (size_t)nkind * ntab);
    c->h_potr.assign(potr, potr + (size_t)nkind * ntab);
    c->h_fpotr.assign(fpotr, fpotr + (size_t)nkind * ntab);
    c->h_fpotb.assign(fpotb, fpotb + (size_t)nkind * ntab);
    c->has_tables = true;
    c->tiled.dirty = true;
    return MDB_OK;
}

extern "C" int mdb_tables_clear(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    cudaSetDevice(c->dev);
    cudaStreamSynchronize(c->stream);
    free_tables(c);
    return MDB_OK;
}

// ------------------------------------------------------------------------------------
// neighbour-list container management (kernels live in mdb_cells.cu / mdb_nlist.cu)
// ------------------------------------------------------------------------------------
extern "C" int mdb_nlist_init(mdb_ctx *c, const double *nb_rm, int mxkvois)
{
    if (!c || !nb_rm || mxkvois < 1) return mdb_fail(c, MDB_ERR_ARG, "mdb_nlist_init: bad argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_nlist_init: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    free_nlist(c);
    double rmmax = 0.0;
    for (int i = 0; i < c->ng * c->ng; i++) {
        c->nb_rm[i] = nb_rm[i];
        // m_RM2 = NB_RM*NB_RM (double) assigned to real(KINDSF) RM2, CommonGPU/MD_NeighborsList_GPU.F90:1377,1387,1394
        c->rm2f[i] = (float)(nb_rm[i] * nb_rm[i]);
        if (nb_rm[i] > rmmax) rmmax = nb_rm[i];
    }
    if (!(rmmax > 0.0)) return mdb_fail(c, MDB_ERR_ARG, "mdb_nlist_init: NB_RM must be positive");
    const double eps = (double)0.0001f; // real(KINDDF),parameter::EPS=0.0001 (a REAL literal) :260
    for (int k = 0; k < 3; k++) {       // :289-298
        c->ncell[k] = (int)(c->box.size[k] / (1.0 * rmmax) - eps);
        if (c->ncell[k] < 3) c->ncell[k] = 3;
    }
    long long nc0 = (long long)c->ncell[0] * c->ncell[1] * c->ncell[2];
    if (nc0 * c->nbox > 2000000000LL) return mdb_fail(c, MDB_ERR_ARG, "mdb_nlist_init: too many cells");
    c->nc0 = (int)nc0;
    c->nc = c->nc0 * c->nbox;
    c->mxkvois = mxkvois;
    CUDA_TRY(c, cudaMalloc(&c->nac, sizeof(int) * (size_t)c->nc));
    CUDA_TRY(c, cudaMalloc(&c->naac, sizeof(int) * (size_t)c->nc));
    CUDA_TRY(c, cudaMalloc(&c->ia1th, sizeof(int) * (size_t)c->nc));
    CUDA_TRY(c, cudaMalloc(&c->kvois, sizeof(int) * ((size_t)c->n + 8))); // +8: aligned TMA windows
    CUDA_TRY(c, cudaMalloc(&c->indi, sizeof(int) * (size_t)c->n * mxkvois));
    CUDA_TRY(c, cudaMemsetAsync(c->kvois, 0, sizeof(int) * (size_t)c->n, c->stream)); // DevSet(KVOIS,0) :285
    CUDA_TRY(c, cudaMemsetAsync(c->indi, 0, sizeof(int) * (size_t)c->n * mxkvois, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->nac, 0, sizeof(int) * (size_t)c->nc, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->naac, 0, sizeof(int) * (size_t)c->nc, c->stream));
    c->has_nlist = true;
    c->list_valid = false;
    c->tiled.dirty = true;
    return MDB_OK;
}

extern "C" int mdb_nlist_clear(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    cudaSetDevice(c->dev);
    cudaStreamSynchronize(c->stream);
    free_nlist(c);
    return MDB_OK;
}

extern "C" int mdb_nlist_cellinfo(const mdb_ctx *c, int ncell[3], int *nc_total, int *mxnac)
{
    if (!c || !c->has_nlist) return MDB_ERR_STATE;
    for (int d = 0; d < 3; d++) ncell[d] = c->ncell[d];
    if (nc_total) *nc_total = c->nc;
    if (mxnac) *mxnac = c->mxnac;
    return MDB_OK;
}

// Rebuild + capacity check (one small device-to-host copy and a stream synchronisation): when a tile's halo, an atom's
// list or mxKVOIS overflowed on the tiled path, the same positions are rebuilt at once on the generic path (AUTO), so no
// step ever runs on an incomplete list.  Used by mdb_nlist_build and at every rebuild inside mdb_run.
void *mdb_scratch(mdb_ctx *c, size_t bytes)
{
    if (c->scratch_bytes >= bytes) return c->scratch;
    if (c->scratch) { cudaStreamSynchronize(c->stream); cudaFree(c->scratch); }
    c->scratch = nullptr; c->scratch_bytes = 0;
    if (cudaMalloc(&c->scratch, bytes + bytes / 4) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    c->scratch_bytes = bytes + bytes / 4;
    return c->scratch;
}

int mdb_list_rebuild_checked(mdb_ctx *c)
{
    int rc = mdb_list_rebuild(c);
    if (rc < 0) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(c->h_counters, c->counters, sizeof(int) * CNT__N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->tiled.active && c->h_counters[CNT_TILE_OVERFLOW] > 0) {
        // a tile's halo / an atom's list did not fit its shared-memory budget (strongly non-uniform density), or an
        // atom has more than mxKVOIS neighbours (the reference truncates in ITS scan order, which only the generic
        // kernel reproduces)
        if (c->opt_force_path == MDB_FORCE_PATH_TILED)
            return mdb_fail(c, MDB_ERR_UNSUPPORTED, "tiled path: %d tiles / cells exceed the halo, list or mxKVOIS capacity (longest list %d)",
                            c->h_counters[CNT_TILE_OVERFLOW], c->h_counters[CNT_NNMAX]);
        c->tiled.ok = false; // AUTO: fall back to the generic path and rebuild
        rc = mdb_list_rebuild(c);
        if (rc < 0) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->h_counters, c->counters, sizeof(int) * CNT__N, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        c->fallbacks++;
    }
    c->mxnac = c->h_counters[CNT_MXNAC];
    return MDB_OK;
}

extern "C" int mdb_nlist_build(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_nlist) return mdb_fail(c, MDB_ERR_STATE, "mdb_nlist_build: mdb_nlist_init first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    int rc = mdb_list_rebuild_checked(c);
    if (rc < 0) return rc;
    if (c->dd_on) {
        if (!c->tiled.active) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "slab decomposition needs the tiled path");
        if ((rc = mdb_dd_update(c)) < 0) return rc;
    }
    return c->h_counters[CNT_OOB];
}

// Reorder_NeighBoreList_Nearest_Dev(Nearest), CommonGPU/MD_NeighborsList_GPU.F90:2016-2066
extern "C" int mdb_nlist_reorder_nearest(mdb_ctx *c, int nearest)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_nlist || !c->list_valid) return mdb_fail(c, MDB_ERR_STATE, "mdb_nlist_reorder_nearest: no valid list");
    if (nearest < 1 || nearest > 512) // the reference prints and stops above mp_MXNEAREST = 512 (:2024-2029)
        return mdb_fail(c, MDB_ERR_ARG, "mdb_nlist_reorder_nearest: NEAREST = %d outside 1..512 (MXNEAREST)", nearest);
    if (c->mxkvois > 512) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_nlist_reorder_nearest: lists longer than 512 entries");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_nlist_reorder_nearest: not available in slab-decomposed runs yet");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    int rc = mdb_indi_ensure(c);
    if (rc < 0) return rc;
    if ((rc = mdb_nlist_nearest(c, nearest)) < 0) return rc;
    c->list_reordered = true; // forces now follow the truncated list (reference behaviour) until the next rebuild
    return MDB_OK;
}

// After a rebuild on the tiled path only the slot lists exist.  The reference-format INDI(N,mxKVOIS) -- same
// members, reference order -- is produced here, on demand, by the generic list kernel from the positions saved at
// that rebuild (cells, types and the sort order are frozen between rebuilds).
int mdb_indi_ensure(mdb_ctx *c)
{
    if (!c->has_nlist || !c->list_valid) return mdb_fail(c, MDB_ERR_STATE, "no valid neighbour list");
    if (!c->tiled.active || !c->indi_stale) return MDB_OK;
    int rc = mdb_nlist_kernel(c, c->pos_snap);
    if (rc < 0) return rc;
    c->indi_stale = false;
    return MDB_OK;
}

// cell sort + list kernel of the active path, no host synchronisation
int mdb_list_rebuild(mdb_ctx *c)
{
    if (c->tiled.dirty) {
        c->tiled.ok = false;
        if (c->opt_force_path != MDB_FORCE_PATH_GENERIC) {
            int rc = mdb_tiled_plan(c);
            if (rc < 0) return rc;
        }
        c->tiled.dirty = false;
    }
    if (c->opt_force_path == MDB_FORCE_PATH_TILED && !c->tiled.ok)
        return mdb_fail(c, MDB_ERR_UNSUPPORTED, "tiled path not available for this configuration (tables, BOXSHAPE or density)");
    c->tiled.active = c->tiled.ok && c->opt_force_path != MDB_FORCE_PATH_GENERIC;
    c->list_gen++;
    int rc = mdb_cells_build(c);
    if (rc < 0) return rc;
    c->indi_stale = false;
    c->list_reordered = false;
    rc = c->tiled.active ? mdb_tiled_nlist(c) : mdb_nlist_kernel(c);
    if (rc < 0) return rc;
    c->list_valid = true;
    return MDB_OK;
}

// owned / ghost atom ranges of this rank after a rebuild (slab decomposition along z, cell-sorted order makes every
// z-layer of cells one contiguous atom range)
int mdb_dd_update(mdb_ctx *c)
{
    if (!c->dd_on) return MDB_OK;
    const int cl = c->ncell[0] * c->ncell[1], ncz = c->ncell[2], R = c->dd_n, r = c->dd_rank;
    const int zl0 = (int)(((long long)r * ncz) / R), zl1 = (int)(((long long)(r + 1) * ncz) / R);
    if (zl1 - zl0 < 1) return mdb_fail(c, MDB_ERR_ARG, "slab decomposition: more ranks than z-layers of cells");
    const int zb = (zl0 - 1 + ncz) % ncz, za = zl1 % ncz; // ghost layers below / above (periodic wrap)
    if (!c->h_dd) CUDA_TRY(c, cudaMallocHost(&c->h_dd, sizeof(int) * 16));
    // first atom (0-based) of a z-layer = IA1th of its first cell - 1; layer ncz -> number of atoms in cells
    const int layers[8] = {zl0, zl1, zb, zb + 1, za, za + 1, zl0 + 1, zl1 - 1};
    for (int i = 0; i < 8; i++) {
        if (layers[i] >= ncz) CUDA_TRY(c, cudaMemcpyAsync(c->h_dd + i, c->counters + CNT_INCELL, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        else CUDA_TRY(c, cudaMemcpyAsync(c->h_dd + i, c->ia1th + (size_t)layers[i] * cl, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    auto first = [&](int i) { return layers[i] >= ncz ? c->h_dd[i] : c->h_dd[i] - 1; };
    int *d = c->dd_info;
    d[0] = first(0); d[1] = first(1);            // owned atoms [a0, a1)
    d[2] = first(2); d[3] = first(3);            // ghost layer below
    d[4] = first(4); d[5] = first(5);            // ghost layer above
    d[6] = first(0); d[7] = first(6);            // my bottom layer (sent to the rank below)
    d[8] = first(7); d[9] = first(1);            // my top layer (sent to the rank above)
    d[10] = (r - 1 + R) % R; d[11] = (r + 1) % R;
    d[12] = zl0 * cl; d[13] = zl1 * cl;
    const int tiles_per_layer = c->ncell[1] * c->tiled.ntx;
    d[14] = zl0 * tiles_per_layer; d[15] = zl1 * tiles_per_layer;
    return MDB_OK;
}

extern "C" int mdb_dd_set(mdb_ctx *c, int rank, int nranks)
{
    if (!c || nranks < 1 || rank < 0 || rank >= nranks) return mdb_fail(c, MDB_ERR_ARG, "mdb_dd_set: bad rank/nranks");
    if (nranks > 1 && c->nbox != 1) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_dd_set: slab decomposition is for a single box");
    c->dd_on = nranks > 1;
    c->dd_rank = rank; c->dd_n = nranks;
    c->dd_built = false;
    c->list_valid = false;
    return MDB_OK;
}

extern "C" int mdb_dd_info(const mdb_ctx *c, int out[16])
{
    if (!c || !c->dd_on) return MDB_ERR_STATE;
    for (int i = 0; i < 16; i++) out[i] = c->dd_info[i];
    return MDB_OK;
}

// rho > RHOMX events since the last call (the embedding table ends at RHOMX = max(POTB)*RHOSCAL; the reference reads past
// DFEMBD there, here the row is clamped to the zero pad, i.e. dF/drho = 0 for that atom)
extern "C" int mdb_embed_overruns(mdb_ctx *c)
{
    if (!c || !c->counters) return MDB_ERR_STATE;
    CUDA_TRY(c, cudaSetDevice(c->dev));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_counters + CNT_RHO_OVER, c->counters + CNT_RHO_OVER, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->counters + CNT_RHO_OVER, 0, sizeof(int), c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return c->h_counters[CNT_RHO_OVER];
}

extern "C" int mdb_nlist_overflow(mdb_ctx *c)
{
    if (!c || !c->has_nlist) return MDB_ERR_STATE;
    CUDA_TRY(c, cudaMemcpyAsync(c->h_counters, c->counters, sizeof(int) * CNT__N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return c->h_counters[CNT_OVERFLOW];
}

__global__ void k_indi_to_original(int n, int k, const int *__restrict__ indi, const int *__restrict__ kvois,
                                   const int *__restrict__ gid, int *__restrict__ out)
{
    // list of ORIGINAL atom o = list of CELL atom s with ids mapped through GID
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int o = gid[s] - 1, kv = kvois[s];
    for (int w = 0; w < k; w++) out[o + (size_t)w * n] = (w < kv) ? gid[indi[s + (size_t)w * n] - 1] : 0;
}

extern "C" int mdb_nlist_copyout(mdb_ctx *c, int *kvois, int *indi, int order)
{
    // Copyout_NeighboreList_DEV, CommonGPU/MD_NeighborsList_GPU.F90:560-755
    if (!c) return MDB_ERR_ARG;
    if (!c->has_nlist || !c->list_valid) return mdb_fail(c, MDB_ERR_STATE, "mdb_nlist_copyout: no valid list");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    int n = c->n;
    if (kvois) {
        int rc = mdb_state_download(c, MDB_F_KVOIS, kvois, order);
        if (rc) return rc;
    }
    if (indi) {
        int rc = mdb_indi_ensure(c);
        if (rc < 0) return rc;
        size_t bytes = sizeof(int) * (size_t)n * c->mxkvois;
        if (order == MDB_ORDER_CELL) {
            CUDA_TRY(c, cudaMemcpyAsync(indi, c->indi, bytes, cudaMemcpyDeviceToHost, c->stream));
        } else {
            void *stg = nullptr;
            int rc = stage_slot(c, bytes, &stg);
            if (rc) return rc;
            {
                ProfScope ps(c, MDB_K_OTHER);
                k_indi_to_original<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->mxkvois, c->indi, c->kvois, c->gid, (int *)stg);
            }
            CUDA_TRY(c, cudaGetLastError());
            CUDA_TRY(c, cudaMemcpyAsync(indi, stg, bytes, cudaMemcpyDeviceToHost, c->stream));
        }
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        c->stage_off = 0;
    }
    return MDB_OK;
}

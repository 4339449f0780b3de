// mdb_cascade.cu -- the cascade-physics neighbours of the hot path (BASELINE configs[4]):
//   * active region by cells   CommonGPU/MD_ActiveRegion_GPU.F90:193-1353  (ActivateRegion_DEV -> ActiveByCells1 -> ActiveByCells0)
//   * electronic stopping      LocalTempCtrlMeths/Stopping/MD_ST_Coupling_GPU.F90:361-600,787-831  (global-density model)
//   * primary knock-on atom    Deposition/MD_TypeDef_Projectile.F90 (CP_DEP_STYPE_PKA, CP_EK_STYLE_MONO, lattice / given direction)
// The force, list and integrator kernels already honour the result: inactive atoms get no force and do not move, and
// cells without an active atom are skipped by the list builders (NAAC, MD_NeighborsList_GPU.F90:981-982) at the next rebuild.
//
// Active region: the reference marks the seed cells on the device, copies the marks to the host, grows them there
// AR_Extend times over the 27-cell neighbourhood (periodic wrap as the list uses it) and copies them back
// (:1221-1287).  Here the growth runs on the device (one kernel per extension step over the cells), same marks.
#include <cstring>
#include <vector>
#include "mdb_stop.cuh"

__global__ void k_ar_deactive_all(int n, int *__restrict__ statu) // DeActive_All_Kernel :242-279
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) statu[i] &= ~ST_ACTIVE;
}
__global__ void k_ar_active_all(int n, int *__restrict__ statu)   // Active_All_Kernel :334-373
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (statu[i] & ST_OUTOFBOX) != ST_OUTOFBOX) statu[i] |= ST_ACTIVE;
}
// CreateSeedByType_Kernel :513-560 and CreateSeedByEkin_Kernel :646-700 in one pass; marks the seed's cell at once
// (MarkSeedCell_Kernel0 :1001-1045).  cent[t] > 0: atoms of type t+1 are seeds; e0 < 0: no energy criterion.
__global__ void k_ar_seed_cells(int n, const int *__restrict__ ityp, const int *__restrict__ statu, const double *__restrict__ xp1,
                                const int *__restrict__ ic, MassParams M, int use_type, const int *__restrict__ cent, int use_ekin,
                                double e0, int *__restrict__ mark)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if ((statu[i] & ST_OUTOFBOX) == ST_OUTOFBOX) return;
    const int t = ityp[i] - 1;
    bool seed = use_type && cent[t] > 0;
    if (use_ekin) {
        const double vx = xp1[i], vy = xp1[i + (size_t)n], vz = xp1[i + 2 * (size_t)n];
        const double ek = 0.5 * M.cm[t] * (vx * vx + vy * vy + vz * vz); // C_HALF*CM*(...) :688
        seed = seed || ek >= e0;
    }
    const int c = ic[i];
    if (seed && c > 0) mark[c - 1] = 1;
}
// one growth step: a marked cell marks its 27-cell neighbourhood (:1238-1284); src -> dst
__global__ void k_ar_grow(int nc0, int nbox, int ncx, int ncy, int ncz, int pdx, int pdy, int pdz, const int *__restrict__ src,
                          int *__restrict__ dst)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc0 * nbox) return;
    if (src[c] <= 0) return;
    const int ib = c / nc0, r = c - ib * nc0;
    const int iz = r / (ncx * ncy), iy = (r - iz * ncx * ncy) / ncx, ix = r - iz * ncx * ncy - iy * ncx;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int x = ix + dx, y = iy + dy, z = iz + dz;
                if (x >= ncx) { if (pdx > 0) x = 0; } else if (x < 0) { if (pdx > 0) x = ncx - 1; }
                if (y >= ncy) { if (pdy > 0) y = 0; } else if (y < 0) { if (pdy > 0) y = ncy - 1; }
                if (z >= ncz) { if (pdz > 0) z = 0; } else if (z < 0) { if (pdz > 0) z = ncz - 1; }
                if (x < 0 || x >= ncx || y < 0 || y >= ncy || z < 0 || z >= ncz) continue;
                dst[ib * nc0 + (z * ncy + y) * ncx + x] = 1;
            }
}
// Active_InCells_Kernel :1085-1127
__global__ void k_ar_activate(int n, const int *__restrict__ ic, const int *__restrict__ mark, int *__restrict__ statu, int *__restrict__ nact)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int a = 0;
    if (i < n) {
        const int c = ic[i];
        int st = statu[i];
        if (c > 0 && mark[c - 1] > 0) { st |= ST_ACTIVE; statu[i] = st; }
        a = (st & ST_ACTIVE) == ST_ACTIVE;
    }
    const unsigned b = __ballot_sync(0xffffffffu, a);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(nact, __popc(b));
}

// ---- activation through the neighbour list (ActiveByNeigbors0/1 :887-997, CP_BYNB_AR)
// per-atom seeds: the two criteria of k_ar_seed_cells on atoms instead of cells (CreateSeedByType_Kernel / CreateSeedByEkin_Kernel)
__global__ void k_ar_seed_atoms(int n, const int *__restrict__ ityp, const int *__restrict__ statu, const double *__restrict__ xp1,
                                MassParams M, int use_type, const int *__restrict__ cent, int use_ekin, double e0, int *__restrict__ seed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int sd = 0;
    if ((statu[i] & ST_OUTOFBOX) != ST_OUTOFBOX) {
        const int t = ityp[i] - 1;
        bool s1 = use_type && cent[t] > 0;
        if (use_ekin) {
            const double vx = xp1[i], vy = xp1[i + (size_t)n], vz = xp1[i + 2 * (size_t)n];
            s1 = s1 || 0.5 * M.cm[t] * (vx * vx + vy * vy + vz * vz) >= e0;
        }
        sd = s1 ? 1 : 0;
    }
    seed[i] = sd;
}
// MarkSeedNeighbore_Kernel :783-842 : a seed marks itself and the atoms of its list (the reference counts, MARK(J) = MARK(J) + 1,
// unsynchronised; only MARK > 0 is ever read)
__global__ void k_ar_mark_neighbours(int n, const int *__restrict__ seed, const int *__restrict__ kvois, const int *__restrict__ indi,
                                     int *__restrict__ mark)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || seed[i] <= 0) return;
    mark[i] = 1;
    const int kv = kvois[i];
    for (int w = 0; w < kv; w++) mark[indi[i + (size_t)w * n] - 1] = 1;
}
// Active_Marked_Kernel :428-471 (+ the count of active atoms)
__global__ void k_ar_activate_marked(int n, const int *__restrict__ mark, int *__restrict__ statu, int *__restrict__ nact)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int a = 0;
    if (i < n) {
        int st = statu[i];
        if (mark[i] > 0) { st |= ST_ACTIVE; statu[i] = st; }
        a = (st & ST_ACTIVE) == ST_ACTIVE;
    }
    const unsigned b = __ballot_sync(0xffffffffu, a);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(nact, __popc(b));
}

// ActivateRegion_DEV(SimBox, CtrlParam) for the cell method.  method: bit 0 = seeds by atom type (AR_CENTPART), bit 1 = seeds
// by kinetic energy >= ekin_erg (AR_EKIN, already in erg), bit 2 = KEEP (do not clear the active bits first; CP_KEEP_AR), bit 3 =
// grow through the neighbour list instead of the cells (CP_BYNB_AR, ActiveByNeigbors0/1 :887-997: a seed marks the atoms of its
// list, `extend` times; needs INDI, which the tiled path produces on demand).
// Uses the cell ids / the list of the last rebuild.  Returns the number of active atoms (>= 0) or < 0.
extern "C" int mdb_active_region(mdb_ctx *c, int method, const int *centpart, double ekin_erg, int extend)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box || !c->has_nlist || !c->list_valid) return mdb_fail(c, MDB_ERR_STATE, "mdb_active_region: needs a built neighbour list (cell ids)");
    if ((method & 1) && !centpart) return mdb_fail(c, MDB_ERR_ARG, "mdb_active_region: seeds by type need centpart[ngroup]");
    if (extend < 0) return mdb_fail(c, MDB_ERR_ARG, "mdb_active_region: negative extension");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_active_region: not available in slab-decomposed runs yet");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    const int n = c->n, nc = c->nc, nb = cdiv(n, 256);
    cudaStream_t st = c->stream;
    if (method & 8) {
        int rc = mdb_indi_ensure(c);
        if (rc < 0) return rc;
        int *w = reinterpret_cast<int *>(mdb_scratch(c, sizeof(int) * (2 * (size_t)n + MDB_MXGROUP + 1)));
        if (!w) return mdb_fail(c, MDB_ERR_CUDA, "mdb_active_region: out of device memory");
        int *seed = w, *mark = w + n, *cent = w + 2 * (size_t)n, *nact = cent + MDB_MXGROUP;
        CUDA_TRY(c, cudaMemsetAsync(cent, 0, sizeof(int) * (MDB_MXGROUP + 1), st));
        if (method & 1) CUDA_TRY(c, cudaMemcpyAsync(cent, centpart, sizeof(int) * c->ng, cudaMemcpyHostToDevice, st));
        ProfScope ps(c, MDB_K_OTHER, 3 + extend);
        if (!(method & 4)) k_ar_deactive_all<<<nb, 256, 0, st>>>(n, c->statu);   // ActiveByNeigbors1 :984-990
        if (method & 3) {
            k_ar_seed_atoms<<<nb, 256, 0, st>>>(n, c->ityp, c->statu, c->xp1, c->mass, method & 1, cent, (method & 2) ? 1 : 0, ekin_erg, seed);
            CUDA_TRY(c, cudaMemcpyAsync(mark, seed, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));   // DevMakeCopy :932
            for (int l = 0; l < extend; l++) {                                    // :935-951
                k_ar_mark_neighbours<<<nb, 256, 0, st>>>(n, seed, c->kvois, c->indi, mark);
                CUDA_TRY(c, cudaMemcpyAsync(seed, mark, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));
            }
            k_ar_activate_marked<<<nb, 256, 0, st>>>(n, mark, c->statu, nact);
        } else {
            CUDA_TRY(c, cudaMemsetAsync(mark, 0, sizeof(int) * (size_t)n, st));
            k_ar_activate_marked<<<nb, 256, 0, st>>>(n, mark, c->statu, nact);   // (counts what KEEP left active)
        }
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaMemcpyAsync(c->h_counters + CNT_SCRATCH, nact, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaStreamSynchronize(st));
        return c->h_counters[CNT_SCRATCH];
    }
    // (work space from the context's grow-only scratch buffer: stream-ordered pool allocations were seen to stall for up to
    // seconds after another context of the process had released a few GB)
    int *work = reinterpret_cast<int *>(mdb_scratch(c, sizeof(int) * (2 * (size_t)nc + MDB_MXGROUP + 1)));
    if (!work) return mdb_fail(c, MDB_ERR_CUDA, "mdb_active_region: out of device memory");
    int *m0 = work, *m1 = work + nc, *cent = work + 2 * (size_t)nc, *nact = cent + MDB_MXGROUP;
    CUDA_TRY(c, cudaMemsetAsync(work, 0, sizeof(int) * (2 * (size_t)nc + MDB_MXGROUP + 1), st));
    if (method & 1) CUDA_TRY(c, cudaMemcpyAsync(cent, centpart, sizeof(int) * c->ng, cudaMemcpyHostToDevice, st));
    ProfScope ps(c, MDB_K_OTHER, 3 + extend);
    if (!(method & 4)) k_ar_deactive_all<<<nb, 256, 0, st>>>(n, c->statu);       // ActiveByCells1 :1313-1319
    if (method & 3) {                                                            // no intrinsic seed method: nothing is activated (:1207-1209)
        k_ar_seed_cells<<<nb, 256, 0, st>>>(n, c->ityp, c->statu, c->xp1, c->ic, c->mass, method & 1, cent, (method & 2) ? 1 : 0, ekin_erg, m0);
        for (int l = 0; l < extend; l++) {
            CUDA_TRY(c, cudaMemsetAsync(m1, 0, sizeof(int) * (size_t)nc, st));
            k_ar_grow<<<cdiv(nc, 256), 256, 0, st>>>(c->nc0, c->nbox, c->ncell[0], c->ncell[1], c->ncell[2], c->box.pd[0], c->box.pd[1],
                                                    c->box.pd[2], m0, m1);
            int *t = m0; m0 = m1; m1 = t;
        }
    }
    k_ar_activate<<<nb, 256, 0, st>>>(n, c->ic, m0, c->statu, nact);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(c->h_counters + CNT_SCRATCH, nact, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    return c->h_counters[CNT_SCRATCH];
}

extern "C" int mdb_active_all(mdb_ctx *c, int on)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_active_all: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    ProfScope ps(c, MDB_K_OTHER);
    if (on) k_ar_active_all<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->statu);
    else k_ar_deactive_all<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->statu);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

// ------------------------------------------------------------------------------------
// electronic stopping, global-density model: ST_MOD_GDEN_KERNEL :431-536
// ------------------------------------------------------------------------------------
__global__ void k_stopping(int n, StopParams S, const double *__restrict__ etab, const double *__restrict__ stab,
                           const int *__restrict__ ityp, const int *__restrict__ statu, const double *__restrict__ xp1,
                           double *__restrict__ fp, int a0, int a1, const int *__restrict__ kvois, const int *__restrict__ nbc,
                           double dt, double *__restrict__ eloss, const int *__restrict__ gid)
{
    const int i = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a1) return;
    if ((statu[i] & ST_ACTIVE) != ST_ACTIVE) return;
    double fx = fp[i], fy = fp[i + (size_t)n], fz = fp[i + 2 * (size_t)n], loss;
    if (stop_force(S, etab, stab, ityp[i] - 1, xp1[i], xp1[i + (size_t)n], xp1[i + 2 * (size_t)n], fx, fy, fz, kvois, nbc, i, n, dt,
                   eloss ? &loss : nullptr)) {
        fp[i] = fx; fp[i + (size_t)n] = fy; fp[i + 2 * (size_t)n] = fz;
        if (eloss) eloss[gid[i] - 1] += loss;     // (one thread per atom: no race)
    }
}
// DEN(IG) of the local-density model for boxes with several types: list neighbours of every type (:690-694), from INDI
__global__ void k_stop_nbcount(int n, int ng, const int *__restrict__ kvois, const int *__restrict__ indi, const int *__restrict__ ityp,
                               int *__restrict__ nbc, int a0, int a1)
{
    const int i = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a1) return;
    int cnt[MDB_MXGROUP];
#pragma unroll
    for (int g = 0; g < MDB_MXGROUP; g++) cnt[g] = 0;
    const int kv = kvois[i];
    for (int w = 0; w < kv; w++) {
        const int t = ityp[indi[i + (size_t)w * n] - 1] - 1;
#pragma unroll
        for (int g = 0; g < MDB_MXGROUP; g++) cnt[g] += (t == g) ? 1 : 0;
    }
    for (int g = 0; g < ng; g++) nbc[(size_t)g * n + i] = cnt[g];
}

static StopState *stop_of(mdb_ctx *c) { return reinterpret_cast<StopState *>(c->stop_state); }

// Initialize_STMOD_DEV + Reset_STMOD_DEV (:334-427): the E-S tables of the stopping library (ETAB(NE) in erg, STAB(NE,NK) in
// erg*cm^2, as Stop_Srim / Stop_Z85 / Stop_Z95 / Stop_B fill them), KPAIR(NG,NG) column-major 1-based, per-type switch and the
// number density of every medium type (MDEN, atoms per cm^3; <= 0: atoms of the box / box volume * NA/NPRT as :379-383)
extern "C" int mdb_stopping_set(mdb_ctx *c, int ne, int nk, const double *etab, const double *stab, const int *kpair, const int *enable,
                                const double *mden)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_stopping_set: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    StopState *S = stop_of(c);
    if (S) mdb_stopping_free(c);
    if (ne < 2 || nk < 1 || !etab || !stab || !kpair || !enable) return MDB_OK; // switched off
    S = new StopState();
    memset(&S->P, 0, sizeof(S->P));
    S->P.ne = ne; S->P.nk = nk; S->P.ng = c->ng;
    for (int g = 0; g < c->ng; g++) {
        S->P.enable[g] = enable[g];
        S->P.cm2[g] = 0.5 * c->mass.cm[g];
        S->P.mden[g] = mden ? mden[g] : 0.0;
        if (enable[g] > 0) S->P.on = 1;
        for (int h = 0; h < c->ng; h++) {
            const int k = kpair[g + c->ng * h];
            if (k < 1 || k > nk) { delete S; return mdb_fail(c, MDB_ERR_ARG, "mdb_stopping_set: KPAIR(%d,%d) = %d outside 1..%d", g + 1, h + 1, k, nk); }
            S->P.kpair[g + c->ng * h] = k;
        }
    }
    CUDA_TRY(c, cudaMalloc(&S->etab, sizeof(double) * ne));
    CUDA_TRY(c, cudaMalloc(&S->stab, sizeof(double) * (size_t)ne * nk));
    CUDA_TRY(c, cudaMemcpyAsync(S->etab, etab, sizeof(double) * ne, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(S->stab, stab, sizeof(double) * (size_t)ne * nk, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->stop_state = S;
    return MDB_OK;
}

// model and bookkeeping switches of Reset_STMOD_DEV (:361-427): ST_CTRL%MDEN < 0 selects the local-density model (hm_STMOD =
// mp_STMOD_L, LVOL = 4 pi/3 NB_RM^3 from the list cut-offs), ST_CTRL%SaveEloss the per-atom energy-loss accumulation
extern "C" int mdb_stopping_options(mdb_ctx *c, int local_density, int save_eloss)
{
    if (!c) return MDB_ERR_ARG;
    StopState *S = stop_of(c);
    if (!S) return mdb_fail(c, MDB_ERR_STATE, "mdb_stopping_options: mdb_stopping_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    if (local_density) {
        if (!c->has_nlist) return mdb_fail(c, MDB_ERR_STATE, "mdb_stopping_options: the local-density model needs the list cut-offs (mdb_nlist_init)");
        if (c->dd_on && c->ng > 1)
            return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_stopping_options: local-density model with several atom types is not available in slab-decomposed runs");
        for (int i = 0; i < c->ng * c->ng; i++) {
            const double r = c->nb_rm[i];
            S->P.lv[i] = 4.0 * 3.14159265358979323846 / 3.0 * (r * r * r);   // CP_4PI3*(NB_RM**3)
        }
    }
    S->P.local = local_density ? 1 : 0;
    S->nbc_gen = -1;
    S->save_eloss = save_eloss ? 1 : 0;
    if (S->save_eloss && !S->eloss) {
        CUDA_TRY(c, cudaMalloc(&S->eloss, sizeof(double) * (size_t)c->n));
        CUDA_TRY(c, cudaMemsetAsync(S->eloss, 0, sizeof(double) * (size_t)c->n, c->stream));
    }
    return MDB_OK;
}
// accumulated inelastic energy loss per atom [erg], ORIGINAL order ("Eloss (ev)" data pad of the reference, :1248-1262, in erg);
// reset != 0 clears the accumulators afterwards.  In a decomposed run every rank holds the losses of the atoms while it owned them.
extern "C" int mdb_stopping_eloss(mdb_ctx *c, double *eloss_host, int reset)
{
    if (!c || !eloss_host) return mdb_fail(c, MDB_ERR_ARG, "mdb_stopping_eloss: null argument");
    StopState *S = stop_of(c);
    if (!S || !S->eloss) return mdb_fail(c, MDB_ERR_STATE, "mdb_stopping_eloss: energy-loss bookkeeping is not switched on (mdb_stopping_options)");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    CUDA_TRY(c, cudaMemcpyAsync(eloss_host, S->eloss, sizeof(double) * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
    if (reset) CUDA_TRY(c, cudaMemsetAsync(S->eloss, 0, sizeof(double) * (size_t)c->n, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MDB_OK;
}
// per-type neighbour counts of the local model (several types): refreshed when the list has been rebuilt since the last count
int mdb_stopping_prepare(mdb_ctx *c)
{
    StopState *S = stop_of(c);
    if (!S || !S->P.on || !S->P.local || c->ng == 1 || S->nbc_gen == c->list_gen) return MDB_OK;
    int rc = mdb_indi_ensure(c);
    if (rc < 0) return rc;
    if (!S->nbc) CUDA_TRY(c, cudaMalloc(&S->nbc, sizeof(int) * (size_t)c->ng * c->n));
    ProfScope ps(c, MDB_K_OTHER);
    k_stop_nbcount<<<cdiv(own_a1(c) - own_a0(c), 256), 256, 0, c->stream>>>(c->n, c->ng, c->kvois, c->indi, c->ityp, S->nbc, own_a0(c), own_a1(c));
    CUDA_TRY(c, cudaGetLastError());
    S->nbc_gen = c->list_gen;
    return MDB_OK;
}

// Do_STMOD_Force_DEV / Do_STMOD_Force_Eloss_DEV (:787-831, :1206-1262): to be called after the force and the EPC friction
// (Do_EPCForce_DEV :135-136); dt = CtrlParam%H of the step (energy-loss bookkeeping only)
int mdb_stopping_launch(mdb_ctx *c, double dt)
{
    StopState *S = stop_of(c);
    if (!S || !S->P.on) return MDB_OK; // hm_NEEDDO = .false.
    int rc = mdb_stopping_prepare(c);
    if (rc < 0) return rc;
    ProfScope ps(c, MDB_K_CORRECT);
    k_stopping<<<cdiv(own_a1(c) - own_a0(c), 256), 256, 0, c->stream>>>(c->n, S->P, S->etab, S->stab, c->ityp, c->statu, c->xp1, c->fp,
                                                                      own_a0(c), own_a1(c), c->kvois, S->nbc, dt,
                                                                      S->save_eloss ? S->eloss : nullptr, c->gid);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}
bool mdb_stopping_on(const mdb_ctx *c)
{
    const StopState *S = reinterpret_cast<const StopState *>(c->stop_state);
    return S && S->P.on;
}
void mdb_stopping_free(mdb_ctx *c)
{
    StopState *S = stop_of(c);
    if (S) { cudaFree(S->etab); cudaFree(S->stab); if (S->nbc) cudaFree(S->nbc); if (S->eloss) cudaFree(S->eloss); delete S; }
    c->stop_state = nullptr;
}
extern "C" int mdb_stopping_apply(mdb_ctx *c, double dt)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_stopping_apply: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    return mdb_stopping_launch(c, dt);
}

// ------------------------------------------------------------------------------------
// primary knock-on atom: the velocity of ONE atom is replaced by sqrt(2 EK / m) along a direction
// (internal deposition of MD_TypeDef_Projectile.F90: DEPSTYLE = CP_DEP_STYPE_PKA, EKSTYPE = CP_EK_STYLE_MONO)
// ------------------------------------------------------------------------------------
__global__ void k_pka(int n, int orig, const int *__restrict__ gidinv, const int *__restrict__ gid, int a0, int a1,
                      const int *__restrict__ ityp, MassParams M, double ek,
                      double dx, double dy, double dz, double *__restrict__ xp1, int *__restrict__ statu)
{
    const int s = gidinv[orig - 1] - 1;
    // slab-decomposed runs: only the rank that owns the atom acts (GIDINV is current for owned atoms only; the slot must
    // hold this very atom)
    if (s < a0 || s >= a1 || gid[s] != orig) return;
    const double v = sqrt(2.0 * ek / M.cm[ityp[s] - 1]);
    xp1[s] = v * dx; xp1[s + (size_t)n] = v * dy; xp1[s + 2 * (size_t)n] = v * dz;
    statu[s] |= ST_ACTIVE;
}
extern "C" int mdb_pka_insert(mdb_ctx *c, int orig_id, double ekin_erg, const double dir[3])
{
    if (!c || !dir) return mdb_fail(c, MDB_ERR_ARG, "mdb_pka_insert: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_pka_insert: mdb_box_set first");
    if (orig_id < 1 || orig_id > c->n || !(ekin_erg >= 0.0)) return mdb_fail(c, MDB_ERR_ARG, "mdb_pka_insert: bad atom id / energy");
    const double d = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    if (!(d > 0.0)) return mdb_fail(c, MDB_ERR_ARG, "mdb_pka_insert: zero direction");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    ProfScope ps(c, MDB_K_OTHER);
    k_pka<<<1, 1, 0, c->stream>>>(c->n, orig_id, c->gidinv, c->gid, own_a0(c), own_a1(c), c->ityp, c->mass, ekin_erg, dir[0] / d, dir[1] / d, dir[2] / d, c->xp1, c->statu);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

// ------------------------------------------------------------------------------------
// PARREP event detection on the device: Do_ChangeDetect, Appshell/MD_Method_ParRep_GPU.F90:1094-1167 =
//   copy the replicas aside -> Do_Damp (quench) -> Do_Compare (:1241-1297) -> restore the replicas and their list.
// The reference copies every replica to host SwapBoxes and compares on the host; here the state is saved in device memory
// (ORIGINAL order, so it survives the re-sorts of the quench), the comparison is one kernel, and only NB flags come back.
// ------------------------------------------------------------------------------------
struct SaveState { double *xp = nullptr, *xp1 = nullptr, *dis = nullptr, *fp = nullptr; int *statu = nullptr; int n = 0; };
static SaveState *save_of(mdb_ctx *c) { return reinterpret_cast<SaveState *>(c->save_state); }
void mdb_save_free(mdb_ctx *c)
{
    SaveState *S = save_of(c);
    if (!S) return;
    cudaFree(S->xp); cudaFree(S->xp1); cudaFree(S->dis); cudaFree(S->fp); cudaFree(S->statu);
    delete S;
    c->save_state = nullptr;
}
__global__ void k_save(int n, const int *__restrict__ gid, const double4 *__restrict__ pos, const double *__restrict__ xp1,
                       const double *__restrict__ dis, const double *__restrict__ fp, const int *__restrict__ statu,
                       double *__restrict__ sx, double *__restrict__ sv, double *__restrict__ sd, double *__restrict__ sf, int *__restrict__ ss)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int o = gid[s] - 1;
    const double4 p = pos[s];
    sx[o] = p.x; sx[o + (size_t)n] = p.y; sx[o + 2 * (size_t)n] = p.z;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        sv[o + (size_t)d * n] = xp1[s + (size_t)d * n];
        sd[o + (size_t)d * n] = dis[s + (size_t)d * n];
        sf[o + (size_t)d * n] = fp[s + (size_t)d * n];
    }
    ss[o] = statu[s];
}
__global__ void k_restore(int n, const int *__restrict__ gid, double4 *__restrict__ pos, double *__restrict__ xp1, double *__restrict__ dis,
                          double *__restrict__ fp, int *__restrict__ statu, const double *__restrict__ sx, const double *__restrict__ sv,
                          const double *__restrict__ sd, const double *__restrict__ sf, const int *__restrict__ ss)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int o = gid[s] - 1;
    double4 p = pos[s];
    p.x = sx[o]; p.y = sx[o + (size_t)n]; p.z = sx[o + 2 * (size_t)n];
    pos[s] = p;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        xp1[s + (size_t)d * n] = sv[o + (size_t)d * n];
        dis[s + (size_t)d * n] = sd[o + (size_t)d * n];
        fp[s + (size_t)d * n] = sf[o + (size_t)d * n];
    }
    statu[s] = ss[o];
}
// Copy_SimMDBox(SimBox(IB), SwapBox(IB)) for all replicas (:1131-1133), on the device
extern "C" int mdb_state_save(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_state_save: mdb_box_set first");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_state_save: not available in slab-decomposed runs");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    SaveState *S = save_of(c);
    const int n = c->n;
    if (S && S->n != n) { mdb_save_free(c); S = nullptr; }
    if (!S) {
        S = new SaveState();
        S->n = n;
        c->save_state = S;
        CUDA_TRY(c, cudaMalloc(&S->xp, sizeof(double) * 3 * (size_t)n)); CUDA_TRY(c, cudaMalloc(&S->xp1, sizeof(double) * 3 * (size_t)n));
        CUDA_TRY(c, cudaMalloc(&S->dis, sizeof(double) * 3 * (size_t)n)); CUDA_TRY(c, cudaMalloc(&S->fp, sizeof(double) * 3 * (size_t)n));
        CUDA_TRY(c, cudaMalloc(&S->statu, sizeof(int) * (size_t)n));
    }
    ProfScope ps(c, MDB_K_OTHER);
    k_save<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->gid, c->pos, c->xp1, c->dis, c->fp, c->statu, S->xp, S->xp1, S->dis, S->fp, S->statu);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}
// CopyIn_SimBox_DEV(SimBox) + Cal_NeighBoreList_DEV (:1158-1159): the saved replicas come back and the list is rebuilt
extern "C" int mdb_state_restore(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    SaveState *S = save_of(c);
    if (!S || S->n != c->n) return mdb_fail(c, MDB_ERR_STATE, "mdb_state_restore: nothing saved (mdb_state_save)");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    {
        ProfScope ps(c, MDB_K_OTHER);
        k_restore<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->gid, c->pos, c->xp1, c->dis, c->fp, c->statu, S->xp, S->xp1, S->dis, S->fp, S->statu);
    }
    CUDA_TRY(c, cudaGetLastError());
    mdb_mark_positions_dirty(c);
    c->list_valid = false;
    return c->has_nlist ? mdb_nlist_build(c) : (int)MDB_OK;
}

// Do_Compare (:1241-1297): Flag = 1 where atom I of replica IB sits further than DRTOL from its place in SimBoxIni (minimum image
// on periodic axes, strict '>' on the squared distance, MASK(I) <= 0 skipped); a replica with any flag is a transition.
__global__ void k_compare(int n, int napb, const int *__restrict__ gid, const double4 *__restrict__ pos, const double *__restrict__ xini,
                          const int *__restrict__ mask, BoxParams box, double rc2, int *__restrict__ flag_atom, int *__restrict__ flag_box)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int o = gid[s] - 1, ib = o / napb, i = o - ib * napb;
    int f = 0;
    if (!mask || mask[i] > 0) {
        const double4 p = pos[s];
        double sep[3] = {__dsub_rn(xini[i], p.x), __dsub_rn(xini[i + (size_t)napb], p.y), __dsub_rn(xini[i + 2 * (size_t)napb], p.z)};
#pragma unroll
        for (int d = 0; d < 3; d++)
            if (box.pd[d] > 0 && fabs(sep[d]) > box.half[d]) sep[d] = __dsub_rn(sep[d], copysign(box.size[d], sep[d]));
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(sep[0], sep[0]), __dmul_rn(sep[1], sep[1])), __dmul_rn(sep[2], sep[2]));
        f = r2 > rc2 ? 1 : 0;
    }
    if (flag_atom) flag_atom[o] = f;
    if (f) flag_box[ib] = 1;
}
// xp_ini: SimBoxIni%XP(NPRT,3) column-major (host); mask: NPRT ints or NULL; drtol in cm; flag_box[nbox] (host out);
// flag_atom: nbox*NPRT ints, replica-major, or NULL.  IBT = last replica with a flag (1-based, 0 = none), NCB = how many.
extern "C" int mdb_compare(mdb_ctx *c, const double *xp_ini, const int *mask, double drtol, int *flag_box, int *flag_atom, int *ibt, int *ncb)
{
    if (!c || !xp_ini || !flag_box) return mdb_fail(c, MDB_ERR_ARG, "mdb_compare: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_compare: mdb_box_set first");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_compare: not available in slab-decomposed runs");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    const int n = c->n, napb = c->napb, nb = c->nbox;
    cudaStream_t st = c->stream;
    const size_t bx = sizeof(double) * 3 * (size_t)napb, bm = sizeof(int) * (size_t)napb, bf = sizeof(int) * (size_t)nb, ba = sizeof(int) * (size_t)n;
    char *w = reinterpret_cast<char *>(mdb_scratch(c, bx + bm + bf + ba + 64));
    if (!w) return mdb_fail(c, MDB_ERR_CUDA, "mdb_compare: out of device memory");
    double *dx = (double *)w; int *dm = (int *)(w + bx), *dfb = dm + napb, *dfa = dfb + nb;
    CUDA_TRY(c, cudaMemcpyAsync(dx, xp_ini, bx, cudaMemcpyHostToDevice, st));
    if (mask) CUDA_TRY(c, cudaMemcpyAsync(dm, mask, bm, cudaMemcpyHostToDevice, st));
    CUDA_TRY(c, cudaMemsetAsync(dfb, 0, bf, st));
    {
        ProfScope ps(c, MDB_K_OTHER);
        k_compare<<<cdiv(n, 256), 256, 0, st>>>(n, napb, c->gid, c->pos, dx, mask ? dm : nullptr, c->box, drtol * drtol, flag_atom ? dfa : nullptr, dfb);
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(flag_box, dfb, bf, cudaMemcpyDeviceToHost, st));
    if (flag_atom) CUDA_TRY(c, cudaMemcpyAsync(flag_atom, dfa, ba, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    int last = 0, cnt = 0;
    for (int b = 0; b < nb; b++) if (flag_box[b] > 0) { last = b + 1; cnt++; } // :1146-1156
    if (ibt) *ibt = last;
    if (ncb) *ncb = cnt;
    return MDB_OK;
}

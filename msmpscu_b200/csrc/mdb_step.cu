// mdb_step.cu -- velocity-Verlet (Swope) predictor / corrector, kinetic energy, electron-phonon
// coupling, and the whole-step driver.
//
// Reference: CommonGPU/MD_DiffScheme_GPU.F90:242-384 (Predictor_KERNEL0), :686-758
// (Correction_KERNEL), :851-902 (CALEKIN_KERNEL); LocalTempCtrlMeths/EPC/MD_EP_Coupling_GPU.F90
// :370-417 (parameters) and :421-493 (EPC_MOD_KERNEL); step order
// Appshell/MD_Method_GenericMD_GPU.F90:596-627.
//
// Positions feed the bit-exact cell assignment, so the predictor's position update is written
// with un-fused intrinsics in the reference's source order; velocities likewise so that long
// trajectories track the oracle as closely as the force sums allow.
#include <cmath>
#include <vector>
#include "mdb_stop.cuh"

// one atom of the predictor; returns |displacement since the last rebuild|^2 (0 when not tracked).
// pre != 0: the EPC friction (bit 0) and the corrector half-kick (bit 1) of the PREVIOUS step are applied first, on
// the values already in registers -- the same operations k_epc_correct performs, one pass over XP1/FP saved.
__device__ __forceinline__ float predict_atom(int i, int n, double4 *__restrict__ pos, double *__restrict__ xp1,
                                              double *__restrict__ fp, double *__restrict__ dis,
                                              int *__restrict__ statu, const int *__restrict__ ityp, const MassParams &M,
                                              const BoxParams &box, double th, double h2s2, double hs2,
                                              float *__restrict__ dsr, int pre, const EpcParams &E)
{
    const int stat = statu[i];
    const int kk = ityp[i] - 1;
    const double cm0 = M.cm[kk];
    double4 p = pos[i];
    double x[3] = {p.x, p.y, p.z};
    const int fixp[3] = {ST_FIXPOSX, ST_FIXPOSY, ST_FIXPOSZ};
    const int fixv[3] = {ST_FIXVELX, ST_FIXVELY, ST_FIXVELZ};
    // every load of the atom is issued before the first dependent instruction (the kernel is latency-bound on HBM)
    double v[3], f[3], ds[3];
    float dr[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        v[d] = xp1[i + (size_t)d * n]; f[d] = fp[i + (size_t)d * n]; ds[d] = dis[i + (size_t)d * n];
        if (dsr) dr[d] = dsr[i + (size_t)d * n];
    }
    if ((pre & 1) && E.enable[kk] > 0) { // EPC_MOD_KERNEL, MD_EP_Coupling_GPU.F90:473-490
        const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2]));
        if (v2 <= E.eup[kk]) {
            const double tm = __dmul_rn(v2, E.v2ti[kk]);
            const double mu = __ddiv_rn(__dmul_rn(E.epa[kk], __dsub_rn(tm, E.te[kk])), fmax(tm, E.tcut[kk]));
#pragma unroll
            for (int d = 0; d < 3; d++) { f[d] = __dsub_rn(f[d], __dmul_rn(mu, v[d])); fp[i + (size_t)d * n] = f[d]; }
        }
    }
    float d2 = 0.f;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const size_t o = i + (size_t)d * n;
        const double a = __ddiv_rn(f[d], cm0);                                  // FP/CM0 :309-311
        const bool freev = (stat & fixv[d]) == 0 && (stat & fixp[d]) == 0;
        if ((pre & 2) && freev) v[d] = __dadd_rn(v[d], __dmul_rn(hs2, a));      // Correction_KERNEL :735-753
        double dd = __dadd_rn(__dmul_rn(th, v[d]), __dmul_rn(h2s2, a));        // TH*XP1 + H2S2*FP :314
        if ((stat & fixp[d]) == fixp[d]) dd = 0.0;
        double xx = __dadd_rn(x[d], dd);
        if (box.pd[d]) {                                                         // :317-325
            if (xx > box.up[d]) xx = __dsub_rn(xx, box.size[d]);
            else if (xx < box.lo[d]) xx = __dadd_rn(xx, box.size[d]);
        }
        x[d] = xx;
        if (freev) xp1[o] = __dadd_rn(v[d], __dmul_rn(hs2, a));                 // :353-361
        dis[o] = __dadd_rn(ds[d], dd);                                           // :371-373
        if (dsr) { // un-wrapped displacement accumulated since the last neighbour rebuild (fp32 is ample)
            const float t = dr[d] + (float)dd;
            dsr[o] = t;
            d2 += t * t;
        }
    }
    p.x = x[0]; p.y = x[1]; p.z = x[2];
    pos[i] = p;
    // :375-379 (the PASSBOUND bit set at :320 is never stored by the reference)
    if (x[0] > box.up[0] || x[1] > box.up[1] || x[2] > box.up[2]) statu[i] = ST_OUTOFBOX | ST_REFLECT;
    else if (x[0] < box.lo[0] || x[1] < box.lo[1] || x[2] < box.lo[2]) statu[i] = ST_OUTOFBOX | ST_TRANSMIT;
    return d2;
}

__global__ void __launch_bounds__(256, 4) k_predict(int n, double4 *__restrict__ pos, double *__restrict__ xp1, double *__restrict__ fp,
                          double *__restrict__ dis, int *__restrict__ statu, const int *__restrict__ ityp,
                          MassParams M, BoxParams box, double th, double h2s2, double hs2,
                          float *__restrict__ dsr, int *__restrict__ counters, int a0, int a1, int pre, EpcParams E,
                          const int *__restrict__ skip, float *__restrict__ dmax_blk)
{
    cudaTriggerProgrammaticLaunchCompletion();   // (programmatic dependent launch: see k_tile_pass)
    cudaGridDependencySynchronize();
    if (skip && *skip) return; // converged quench iteration (mdb_dyndamp)
    const int i = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    float d2 = 0.f;
    if (i < a1 && (statu[i] & ST_ACTIVE) == ST_ACTIVE) // :295
        d2 = predict_atom(i, n, pos, xp1, fp, dis, statu, ityp, M, box, th, h2s2, hs2, dsr, pre, E);
    if (dsr) { // block maximum -> one atomic per block; the tiled passes compare it with their class margin
        for (int off = 16; off > 0; off >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, off));
        __shared__ float smax[8];
        if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = d2;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = smax[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = fmaxf(m, smax[w]);
            if (m > 0.f) atomicMax(&counters[CNT_D2MAX], __float_as_int(m)); // non-negative floats order like ints
            if (dmax_blk) dmax_blk[blockIdx.x] = m; // per-block bound for the per-tile decision of cascade runs (k_tile_d2)
        }
    }
}

// EPC friction on FP followed by the second half kick: one read of XP1/FP instead of the
// reference's two kernels (EPC_MOD_KERNEL then Correction_KERNEL).  do_epc / do_corr select stages.
__global__ void k_epc_correct(int n, double *__restrict__ xp1, double *__restrict__ fp, const int *__restrict__ statu,
                              const int *__restrict__ ityp, MassParams M, EpcParams E, double hs2, int do_epc, int do_corr,
                              int a0, int a1, const int *__restrict__ skip = nullptr)
{
    if (skip && *skip) return;
    const int i = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a1) return;
    const int stat = statu[i];
    if ((stat & ST_ACTIVE) != ST_ACTIVE) return;
    const int kk = ityp[i] - 1;
    const size_t n1 = n, n2 = 2 * (size_t)n;
    double vx = xp1[i], vy = xp1[i + n1], vz = xp1[i + n2];
    double fx = fp[i], fy = fp[i + n1], fz = fp[i + n2];
    if (do_epc && E.enable[kk] > 0) { // EPC_MOD_KERNEL :473-490
        const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
        if (v2 <= E.eup[kk]) {
            const double tm = __dmul_rn(v2, E.v2ti[kk]);
            const double mu = __ddiv_rn(__dmul_rn(E.epa[kk], __dsub_rn(tm, E.te[kk])), fmax(tm, E.tcut[kk]));
            fx = __dsub_rn(fx, __dmul_rn(mu, vx));
            fy = __dsub_rn(fy, __dmul_rn(mu, vy));
            fz = __dsub_rn(fz, __dmul_rn(mu, vz));
            fp[i] = fx; fp[i + n1] = fy; fp[i + n2] = fz;
        }
    }
    if (do_corr) { // Correction_KERNEL :735-753
        const double cm0 = M.cm[kk];
        if ((stat & ST_FIXVELX) == 0 && (stat & ST_FIXPOSX) == 0) xp1[i] = __dadd_rn(vx, __dmul_rn(hs2, __ddiv_rn(fx, cm0)));
        if ((stat & ST_FIXVELY) == 0 && (stat & ST_FIXPOSY) == 0) xp1[i + n1] = __dadd_rn(vy, __dmul_rn(hs2, __ddiv_rn(fy, cm0)));
        if ((stat & ST_FIXVELZ) == 0 && (stat & ST_FIXPOSZ) == 0) xp1[i + n2] = __dadd_rn(vz, __dmul_rn(hs2, __ddiv_rn(fz, cm0)));
    }
}

// The end of a step with electronic stopping in ONE pass over XP1 / FP: EPC friction (EPC_MOD_KERNEL), stopping
// (ST_MOD_GDEN_KERNEL) and the corrector half-kick, each with the arithmetic of its own kernel, on the values in registers.
__global__ void k_step_close(int n, double *__restrict__ xp1, double *__restrict__ fp, const int *__restrict__ statu,
                             const int *__restrict__ ityp, MassParams M, EpcParams E, int do_epc, StopParams S,
                             const double *__restrict__ etab, const double *__restrict__ stab, double hs2, int a0, int a1,
                             const int *__restrict__ kvois, const int *__restrict__ nbc, double *__restrict__ eloss,
                             const int *__restrict__ gid)
{
    const int i = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a1) return;
    const int stat = statu[i];
    if ((stat & ST_ACTIVE) != ST_ACTIVE) return;
    const int kk = ityp[i] - 1;
    const size_t n1 = n, n2 = 2 * (size_t)n;
    const double vx = xp1[i], vy = xp1[i + n1], vz = xp1[i + n2];
    double fx = fp[i], fy = fp[i + n1], fz = fp[i + n2];
    bool wr = false;
    if (do_epc && E.enable[kk] > 0) { // EPC_MOD_KERNEL :473-490
        const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
        if (v2 <= E.eup[kk]) {
            const double tm = __dmul_rn(v2, E.v2ti[kk]);
            const double mu = __ddiv_rn(__dmul_rn(E.epa[kk], __dsub_rn(tm, E.te[kk])), fmax(tm, E.tcut[kk]));
            fx = __dsub_rn(fx, __dmul_rn(mu, vx));
            fy = __dsub_rn(fy, __dmul_rn(mu, vy));
            fz = __dsub_rn(fz, __dmul_rn(mu, vz));
            wr = true;
        }
    }
    double loss;
    const bool stopped = stop_force(S, etab, stab, kk, vx, vy, vz, fx, fy, fz, kvois, nbc, i, n, __dadd_rn(hs2, hs2), eloss ? &loss : nullptr);
    if (stopped && eloss) eloss[gid[i] - 1] += loss;
    wr = stopped || wr;
    if (wr) { fp[i] = fx; fp[i + n1] = fy; fp[i + n2] = fz; }
    const double cm0 = M.cm[kk]; // Correction_KERNEL :735-753
    if ((stat & ST_FIXVELX) == 0 && (stat & ST_FIXPOSX) == 0) xp1[i] = __dadd_rn(vx, __dmul_rn(hs2, __ddiv_rn(fx, cm0)));
    if ((stat & ST_FIXVELY) == 0 && (stat & ST_FIXPOSY) == 0) xp1[i + n1] = __dadd_rn(vy, __dmul_rn(hs2, __ddiv_rn(fy, cm0)));
    if ((stat & ST_FIXVELZ) == 0 && (stat & ST_FIXPOSZ) == 0) xp1[i + n2] = __dadd_rn(vz, __dmul_rn(hs2, __ddiv_rn(fz, cm0)));
}

__global__ void k_ekin(int n, const double *__restrict__ xp1, const int *__restrict__ statu, const int *__restrict__ ityp,
                       MassParams M, double *__restrict__ ekin)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ek = -1.0e32; // :886
    const int st = statu[i];
    if ((st & ST_ACTIVE) == ST_ACTIVE && (st & ST_FIXPOS) == 0) {
        const double cm0 = M.cm[ityp[i] - 1];
        const double vx = xp1[i], vy = xp1[i + (size_t)n], vz = xp1[i + 2 * (size_t)n];
        const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
        ek = __dmul_rn(__dmul_rn(0.5, cm0), v2); // C_HALF*CM0*(...) :896
    }
    ekin[i] = ek;
}

// ------------------------------------------------------------------------------------
// pre: bit 0 = EPC friction, bit 1 = corrector of the previous step, fused in front of this predictor (mdb_run)
static int predict_launch(mdb_ctx *c, double h, int pre)
{
    // Predictor_DEV :660-662 : TH = H, HS2 = TH/2, H2S2 = TH*TH/2
    const double th = h, hs2 = th * 0.5, h2s2 = th * th * 0.5;
    if (!c->epc.on) pre &= ~1;
    ProfScope ps(c, MDB_K_PREDICT);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cdiv(own_a1(c) - own_a0(c), 256)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = (c->opt_pdl && !c->dd_on) ? 1 : 0; // (decomposed steps order their kernels with events)
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_TRY(c, cudaLaunchKernelEx(&cfg, k_predict, c->n, c->pos, c->xp1, c->fp, c->dis, c->statu, (const int *)c->ityp, c->mass, c->box, th, h2s2,
                                       hs2, c->dsr, c->counters, own_a0(c), own_a1(c), pre, c->epc, c->skip_flag,
                                       c->dsr ? c->dmax_blk : (float *)nullptr));
    }
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

// DAMPING_KERNEL, CommonGPU/MD_DiffScheme_GPU.F90:125-185 (called by Predictor_DEV while DAMPTIME0 <= ITIME < DAMPTIME0+DAMPTIME1,
// :611-617, and by Do_DynDamp_Forsteps_DEV): a velocity component opposing its force component, or with a fixed position, is zeroed
__global__ void k_damping(int n, double *__restrict__ xp1, const double *__restrict__ fp, const int *__restrict__ statu, int a0, int a1,
                          const int *__restrict__ skip)
{
    if (skip && *skip) return;
    const int i = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a1) return;
    const int st = statu[i];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const size_t o = i + (size_t)d * n;
        double v = xp1[o];
        if (__dmul_rn(v, fp[o]) < 0.0) v = 0.0;
        if ((st & (ST_FIXPOSX << d)) == (ST_FIXPOSX << d)) v = 0.0;
        xp1[o] = v;
    }
}
extern "C" int mdb_damping(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_damping: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    ProfScope ps(c, MDB_K_PREDICT);
    k_damping<<<cdiv(own_a1(c) - own_a0(c), 256), 256, 0, c->stream>>>(c->n, c->xp1, c->fp, c->statu, own_a0(c), own_a1(c), c->skip_flag);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

int mdb_predict_launch(mdb_ctx *c, double h, int pre) { return predict_launch(c, h, pre); }
// the end of a step that is not fused into the next one: Do_EPCForce_DEV = EPC friction, then electronic stopping
// (MD_LocalTempMethod_GPU.F90:135-136), then Correction_DEV; one kernel when stopping is off
int mdb_step_close_launch(mdb_ctx *c, double h)
{
    const int a0 = own_a0(c), a1 = own_a1(c);
    if (!mdb_stopping_on(c)) return mdb_epc_correct_launch(c, h);
    const StopState *S = reinterpret_cast<const StopState *>(c->stop_state);
    int rc = mdb_stopping_prepare(c);
    if (rc < 0) return rc;
    ProfScope ps(c, MDB_K_CORRECT);
    k_step_close<<<cdiv(a1 - a0, 256), 256, 0, c->stream>>>(c->n, c->xp1, c->fp, c->statu, c->ityp, c->mass, c->epc, c->epc.on, S->P,
                                                          S->etab, S->stab, h * 0.5, a0, a1, c->kvois, S->nbc,
                                                          S->save_eloss ? S->eloss : nullptr, c->gid);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

int mdb_epc_correct_launch(mdb_ctx *c, double h)
{
    ProfScope ps(c, MDB_K_CORRECT);
    k_epc_correct<<<cdiv(own_a1(c) - own_a0(c), 256), 256, 0, c->stream>>>(c->n, c->xp1, c->fp, c->statu, c->ityp, c->mass, c->epc,
                                                                         h * 0.5, c->epc.on, 1, own_a0(c), own_a1(c));
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

extern "C" int mdb_predict(mdb_ctx *c, double h)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_predict: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    return predict_launch(c, h, 0);
}

extern "C" int mdb_correct(mdb_ctx *c, double h)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_correct: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    ProfScope ps(c, MDB_K_CORRECT);
    k_epc_correct<<<cdiv(own_a1(c) - own_a0(c), 256), 256, 0, c->stream>>>(c->n, c->xp1, c->fp, c->statu, c->ityp, c->mass, c->epc,
                                                                         h * 0.5, 0, 1, own_a0(c), own_a1(c), c->skip_flag);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

// Do_EPCForce_DEV + Correction_DEV in one pass over XP1/FP (the same fused kernel mdb_run uses at the end of a block)
extern "C" int mdb_epc_correct(mdb_ctx *c, double h)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_epc_correct: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    ProfScope ps(c, MDB_K_CORRECT);
    k_epc_correct<<<cdiv(own_a1(c) - own_a0(c), 256), 256, 0, c->stream>>>(c->n, c->xp1, c->fp, c->statu, c->ityp, c->mass, c->epc,
                                                                         h * 0.5, c->epc.on, 1, own_a0(c), own_a1(c));
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

extern "C" int mdb_ekin(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_ekin: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    ProfScope ps(c, MDB_K_OTHER);
    k_ekin<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->xp1, c->statu, c->ityp, c->mass, c->ekin);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

extern "C" int mdb_epc_set(mdb_ctx *c, const int *enable, const double *te, const double *alpha, const double *cut,
                           const double *he)
{
    if (!c || !enable || !te || !alpha || !cut || !he) return mdb_fail(c, MDB_ERR_ARG, "mdb_epc_set: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_epc_set: mdb_box_set first");
    EpcParams &E = c->epc;
    memset(&E, 0, sizeof(E));
    for (int g = 0; g < c->ng; g++) { // Reset_EPCMOD_DEV :387-394
        E.enable[g] = enable[g];
        E.te[g] = te[g];
        E.v2ti[g] = c->mass.cm[g] * (1.0 / 3.0) / KB_CGS; // CM*C_UTH/CP_KB
        E.epa[g] = c->mass.cm[g] / alpha[g];
        E.tcut[g] = te[g] * cut[g];
        E.eup[g] = 2.0 * he[g] / c->mass.cm[g];
        if (enable[g] > 0) E.on = 1;
    }
    return MDB_OK;
}

extern "C" int mdb_epc_apply(mdb_ctx *c)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_epc_apply: mdb_box_set first");
    if (!c->epc.on) return MDB_OK; // hm_NEEDDO = .false. :403-405
    CUDA_TRY(c, cudaSetDevice(c->dev));
    ProfScope ps(c, MDB_K_CORRECT);
    k_epc_correct<<<cdiv(own_a1(c) - own_a0(c), 256), 256, 0, c->stream>>>(c->n, c->xp1, c->fp, c->statu, c->ityp, c->mass, c->epc,
                                                                         0.0, 1, 0, own_a0(c), own_a1(c));
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}


// ------------------------------------------------------------------------------------
// The other procedures of the integrator module the step loops call: Cal_GlobalT_DEV (MD_DiffScheme_GPU.F90:1042-1064),
// VelScaling_DEV (:1262-1446) and CheckTimestep_DEV (:1066-1258).  The reference copies EKIN to the host and sums
// there; here the sums are deterministic two-stage device reductions and only the scalars come back.
// ------------------------------------------------------------------------------------
#define RT 256
// per box b (block b): sum and count of EKIN >= 0 over the box's atoms in ORIGINAL order (hm_EKIN(hm_GIDINV(...)) :1418)
__global__ void __launch_bounds__(RT) k_box_ekin(int napb, const double *__restrict__ ekin, const int *__restrict__ gidinv,
                                                 double *__restrict__ bsum, int *__restrict__ bcnt)
{
    __shared__ double sh[RT / 32];
    __shared__ int shc[RT / 32];
    const int b = blockIdx.x;
    double s = 0.0;
    int cn = 0;
    for (int o = threadIdx.x; o < napb; o += RT) {
        const double e = ekin[gidinv[(size_t)b * napb + o] - 1];
        if (e >= 0.0) { s += e; cn++; }
    }
    for (int off = 16; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); cn += __shfl_xor_sync(0xffffffffu, cn, off); }
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = s; shc[threadIdx.x >> 5] = cn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0; int c = 0;
        for (int w = 0; w < RT / 32; w++) { t += sh[w]; c += shc[w]; }
        bsum[b] = t; bcnt[b] = c;
    }
}
// VelScaling_KERNEL :1262-1320 : box of an atom = (sorted index)/NPRTPB; fixed components are zeroed
__global__ void k_vel_scale(int n, int napb, double *__restrict__ xp1, const int *__restrict__ statu, const double *__restrict__ scal)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int st = statu[i];
    if ((st & ST_ACTIVE) != ST_ACTIVE) return;
    const double sc = scal[i / napb];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const bool fr = (st & (ST_FIXPOSX << d)) == 0 && (st & (ST_FIXVELX << d)) == 0;
        const size_t o = i + (size_t)d * n;
        xp1[o] = fr ? __dmul_rn(xp1[o], sc) : 0.0;
    }
}
// CheckTimestep_KERNEL :1066-1140
__global__ void k_check_timestep(int n, const double *__restrict__ xp1, const double *__restrict__ fp, const int *__restrict__ statu,
                                 const int *__restrict__ ityp, MassParams M, double th, double h2s2, double mxd2, int *__restrict__ flag,
                                 int a0, int a1)
{
    const int i = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a1) return;
    const int st = statu[i];
    if ((st & ST_ACTIVE) != ST_ACTIVE) return;
    const double cm0 = M.cm[ityp[i] - 1];
    double d2 = 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double dd = 0.0;
        if ((st & (ST_FIXPOSX << d)) == 0)
            dd = __dadd_rn(__dmul_rn(th, xp1[i + (size_t)d * n]), __dmul_rn(h2s2, __ddiv_rn(fp[i + (size_t)d * n], cm0)));
        d2 = __dadd_rn(d2, __dmul_rn(dd, dd));
    }
    if (d2 > mxd2) *flag = 1;
}

// The halving loop of Predictor_DEV (:633-655) in one pass: bit k of *mask is set when CheckTimestep_KERNEL would raise its
// flag for the trial step TH_k = HMX 2^-k (H2S2_k = TH_k TH_k / 2, both exact scalings of the reference's sequence).  The
// reference halves until no atom objects: its result is TH_k for the lowest clear bit.
#define TS_MAXHALVE 31
__global__ void k_timestep_mask(int n, const double *__restrict__ xp1, const double *__restrict__ fp, const int *__restrict__ statu,
                                const int *__restrict__ ityp, MassParams M, double hmx, double mxd2, unsigned *__restrict__ mask,
                                int a0, int a1)
{
    const int i = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned m = 0u;
    if (i < a1) {
        const int st = statu[i];
        if ((st & ST_ACTIVE) == ST_ACTIVE) {
            const double cm0 = M.cm[ityp[i] - 1];
            double v[3], a[3];
            bool fr[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                fr[d] = (st & (ST_FIXPOSX << d)) == 0;
                v[d] = xp1[i + (size_t)d * n];
                a[d] = __ddiv_rn(fp[i + (size_t)d * n], cm0);
            }
            // |TH v + H2S2 a| <= TH |v| + H2S2 |a|, which grows with TH: an atom that passes this bound at HMX (with a margin for
            // the roundings) passes every trial step -- almost all atoms, which then skip the loop
            const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            const double ub = hmx * vn + 0.5 * hmx * hmx * an;
            double th = hmx;
            for (int k = 0; k < ((ub * ub * 1.000001 <= mxd2) ? 0 : TS_MAXHALVE); k++) {
                const double h2s2 = __dmul_rn(__dmul_rn(th, th), 0.5);
                double d2 = 0.0;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double dd = fr[d] ? __dadd_rn(__dmul_rn(th, v[d]), __dmul_rn(h2s2, a[d])) : 0.0;
                    d2 = __dadd_rn(d2, __dmul_rn(dd, dd));
                }
                if (d2 > mxd2) m |= 1u << k;
                th = __dmul_rn(th, 0.5);
            }
        }
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(mask, m);
}
// enqueues the mask of this rank's atoms and its copy to h_counters[CNT_SCRATCH] (the caller synchronises)
int mdb_timestep_mask_launch(mdb_ctx *c, double hmx, double dmx2)
{
    unsigned *mask = reinterpret_cast<unsigned *>(c->counters + CNT_SCRATCH);
    CUDA_TRY(c, cudaMemsetAsync(mask, 0, sizeof(int), c->stream));
    {
        ProfScope ps(c, MDB_K_OTHER);
        k_timestep_mask<<<cdiv(own_a1(c) - own_a0(c), 256), 256, 0, c->stream>>>(c->n, c->xp1, c->fp, c->statu, c->ityp, c->mass, hmx, dmx2,
                                                                                mask, own_a0(c), own_a1(c));
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(c->h_counters + CNT_SCRATCH, mask, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return MDB_OK;
}
// TH for the OR of the ranks' masks; < 0 when even HMX 2^-30 moves an atom further than DMX
int mdb_timestep_from_mask(mdb_ctx *c, unsigned mask, double hmx, double *h)
{
    int k = 0;
    while (k < TS_MAXHALVE && (mask >> k & 1u)) k++;
    if (k >= TS_MAXHALVE) return mdb_fail(c, MDB_ERR_STATE, "variable time step: an atom moves more than DMX even with HMX/2^%d", TS_MAXHALVE - 1);
    double th = hmx;
    for (int j = 0; j < k; j++) th = th * 0.5; // m_TH = m_TH*C_HALF :648
    *h = th;
    return MDB_OK;
}

// fills per-box sums on the host (pinned staging); returns nbox or <0
static int box_ekin_host(mdb_ctx *c, std::vector<double> &sum, std::vector<int> &cnt)
{
    int rc = mdb_ekin(c);
    if (rc < 0) return rc;
    const int nb = c->nbox;
    if (c->vpart_n < nb + 2) {
        if (c->vpart) cudaFree(c->vpart);
        c->vpart = nullptr;
        CUDA_TRY(c, cudaMalloc(&c->vpart, sizeof(double) * 9 * (size_t)(nb + 2)));
        c->vpart_n = nb + 2;
    }
    double *bsum = c->vpart;
    int *bcnt = reinterpret_cast<int *>(c->vpart + nb);
    {
        ProfScope ps(c, MDB_K_OTHER);
        k_box_ekin<<<nb, RT, 0, c->stream>>>(c->napb, c->ekin, c->gidinv, bsum, bcnt);
    }
    CUDA_TRY(c, cudaGetLastError());
    sum.resize(nb); cnt.resize(nb);
    CUDA_TRY(c, cudaMemcpyAsync(sum.data(), bsum, sizeof(double) * nb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(cnt.data(), bcnt, sizeof(int) * nb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return nb;
}

extern "C" int mdb_global_t(mdb_ctx *c, double *curt)
{
    if (!c || !curt) return mdb_fail(c, MDB_ERR_ARG, "mdb_global_t: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_global_t: mdb_box_set first");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_global_t: reduce the per-rank kinetic energies in slab-decomposed runs");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    std::vector<double> sum; std::vector<int> cnt;
    int nb = box_ekin_host(c, sum, cnt);
    if (nb < 0) return nb;
    double s = 0.0; long long n = 0;
    for (int b = 0; b < nb; b++) { s += sum[b]; n += cnt[b]; }
    *curt = 2.0 * s / (double)n / (3.0 * KB_CGS); // C_TWO*sum(hm_EKIN, mask)/count/(C_THR*CP_KB) :1062
    return MDB_OK;
}

// per-box temperatures of a MULTIBOX run (the per-box sums VelScaling_DEV forms, :1405-1432, as temperatures): what the
// multi-box dispatcher gathers over the ranks at an output interval
extern "C" int mdb_box_temperatures(mdb_ctx *c, double *t_box)
{
    if (!c || !t_box) return mdb_fail(c, MDB_ERR_ARG, "mdb_box_temperatures: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_box_temperatures: mdb_box_set first");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_box_temperatures: not available in slab-decomposed runs");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    std::vector<double> sum; std::vector<int> cnt;
    int nb = box_ekin_host(c, sum, cnt);
    if (nb < 0) return nb;
    for (int b = 0; b < nb; b++) t_box[b] = cnt[b] > 0 ? 2.0 * sum[b] / (double)cnt[b] / (3.0 * KB_CGS) : 0.0;
    return MDB_OK;
}

extern "C" int mdb_vel_scaling(mdb_ctx *c, double dt)
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_vel_scaling: mdb_box_set first");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_vel_scaling: not available in slab-decomposed runs");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    std::vector<double> sum; std::vector<int> cnt;
    int nb = box_ekin_host(c, sum, cnt);
    if (nb < 0) return nb;
    for (int b = 0; b < nb; b++) {
        const double ce = sum[b] / (double)cnt[b];
        if (!(ce > 0.0)) // the reference prints and stops :1424-1430
            return mdb_fail(c, MDB_ERR_STATE, "mdb_vel_scaling: current temperature of box %d is zero in scaling velocity", b + 1);
        sum[b] = sqrt(dt * 3.0 * KB_CGS * 0.5 / ce); // dsqrt(DT*C_THR*CP_KB*C_HALF/cEKIN) :1432
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->vpart, sum.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, c->stream));
    {
        ProfScope ps(c, MDB_K_OTHER);
        k_vel_scale<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->napb, c->xp1, c->statu, c->vpart);
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream)); // sum[] is pageable host memory
    return MDB_OK;
}

extern "C" int mdb_check_timestep(mdb_ctx *c, double th, double h2s2, double dmx2, int *iflag)
{
    if (!c || !iflag) return mdb_fail(c, MDB_ERR_ARG, "mdb_check_timestep: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_check_timestep: mdb_box_set first");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    int *flag = c->counters + CNT_SCRATCH;
    CUDA_TRY(c, cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
    {
        ProfScope ps(c, MDB_K_OTHER);
        k_check_timestep<<<cdiv(own_a1(c) - own_a0(c), 256), 256, 0, c->stream>>>(c->n, c->xp1, c->fp, c->statu, c->ityp, c->mass, th, h2s2,
                                                                                 dmx2, flag, own_a0(c), own_a1(c));
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(c->h_counters + CNT_SCRATCH, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *iflag = c->h_counters[CNT_SCRATCH] ? 1 : 0;
    return MDB_OK;
}

// ------------------------------------------------------------------------------------
extern "C" int mdb_force(mdb_ctx *c, unsigned flags, double vtensor[9])
{
    if (!c) return MDB_ERR_ARG;
    if (!c->has_box || !c->has_tables) return mdb_fail(c, MDB_ERR_STATE, "mdb_force: box and tables must be set");
    if (!c->has_nlist || !c->list_valid) return mdb_fail(c, MDB_ERR_STATE, "mdb_force: no valid neighbour list (mdb_nlist_build)");
    if ((flags & MDB_VIRIAL) && !vtensor) return mdb_fail(c, MDB_ERR_ARG, "mdb_force: MDB_VIRIAL needs vtensor");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    if (c->dd_on && !c->tiled.active)
        return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_force: slab decomposition needs the tiled path");
    // slab-decomposed runs: vtensor is this rank's PARTIAL sum (every directed pair whose first atom it owns, each halved
    // as in CALPTENSOR_KERNEL :1222); the caller adds the partial tensors of all ranks (SlabDomain.force_virial)
    if (c->tiled.active && !c->list_reordered) {
        // density, force (+ virial) and per-atom energy passes all run on the tiled path
        unsigned fast = flags & (MDB_FORCE | MDB_DEN | MDB_NOPASS1 | MDB_EPOT | MDB_VIRIAL);
        int rc = mdb_force_tiled(c, fast);
        if (rc < 0) return rc;
        if (flags & MDB_VIRIAL) return mdb_virial_finish(c, c->tiled.grid * (c->tiled.threads / 32), vtensor); // one partial tensor per warp
        return MDB_OK;
    }
    return mdb_force_generic(c, flags, vtensor);
}

// pCalAVStress -> Cal_EAM_AtomicStressTensor_DEV(IDEV, dAVP), CommonGPU/MD_EAM_ForceTable_GPU.F90:1973-1990.
// The reference uses whatever DEN the last force call left; here the density pass runs first, so the result is the
// same whenever DEN was current and well defined otherwise.
extern "C" int mdb_atomic_stress(mdb_ctx *c, double *d_avp)
{
    if (!c || !d_avp) return mdb_fail(c, MDB_ERR_ARG, "mdb_atomic_stress: null argument");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_atomic_stress: not available in slab-decomposed runs yet");
    int rc = mdb_force(c, MDB_DEN, nullptr);
    if (rc < 0) return rc;
    if ((rc = mdb_indi_ensure(c)) < 0) return rc;
    return mdb_avstress_generic(c, d_avp);
}

// One For_One_Step without host synchronisation.  first / last: position inside an mdb_run block -- between two
// steps of a block the EPC friction + corrector of step s run fused in front of the predictor of step s+1.
static int step_nosync(mdb_ctx *c, int itime, int it0, int nb_uptab, double h, bool first, bool last)
{
    int rc;
    // electronic stopping (Do_STMOD_DEV) acts on FP between the EPC friction and the corrector (Do_EPCForce_DEV,
    // MD_LocalTempMethod_GPU.F90:135-136): with it the three stages run unfused at the end of every step
    const bool stopping = mdb_stopping_on(c);
    const bool fused_epilogue = !stopping && c->opt_fuse_epilogue && c->tiled.active && c->has_tables && c->shape_identity;
    if ((rc = predict_launch(c, h, (first || fused_epilogue || stopping) ? 0 : 3)) < 0) return rc;
    if (nb_uptab > 0 && (itime - it0) % nb_uptab == 0) { // MOD(ITIME-IT0,NB_UPTAB)==0, GenericMD:599-601
        // the capacity counters are read back at once (a 48-byte copy and a synchronisation per list period): an overflow of
        // the tiled path is redone on the generic path before any force is evaluated on the new list
        if ((rc = mdb_list_rebuild_checked(c)) < 0) return rc;
    } else if (mdb_tile_guard_wanted(c) && c->list_valid && !c->list_reordered) {
        // cascade runs: one fast atom must not send every tile to the full list -- per-tile displacement bounds for this step
        if ((rc = mdb_tile_guard_launch(c, 0, c->tiled.P.ntiles)) < 0) return rc;
    }
    struct GuardOff { mdb_ctx *c; ~GuardOff() { c->tile_guard_fresh = false; } } guard_off{c}; // the bounds are this step's only
    if (fused_epilogue && c->list_valid && !c->list_reordered) {
        // EPC friction and the corrector are fused into the epilogue of the force pass
        return mdb_force_tiled(c, MDB_FORCE, 3, h * 0.5);
    }
    if ((rc = mdb_force(c, MDB_FORCE, nullptr)) < 0) return rc;
    if (stopping) return mdb_step_close_launch(c, h);
    if (!last) return MDB_OK; // applied by the next step's predictor kernel
    ProfScope ps(c, MDB_K_CORRECT);
    k_epc_correct<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->xp1, c->fp, c->statu, c->ityp, c->mass, c->epc,
                                                          h * 0.5, c->epc.on, 1, 0, c->n);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

// nsteps x For_One_Step, enqueued on the context's stream.  The host only waits inside a list rebuild (capacity check);
// the out-of-box count of the block is read by mdb_sync.
extern "C" int mdb_run_async(mdb_ctx *c, int itime0, int nsteps, int it0, int nb_uptab, double h)
{
    if (!c) return MDB_ERR_ARG;
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_run: a slab-decomposed step needs the ghost exchanges between its "
                                                           "kernels; use mdb_dd_run (one process per GPU, NCCL) or drive it with "
                                                           "mdb_predict / mdb_force / mdb_correct (msmpscu_b200/domain.py)");
    if (!c->has_box || !c->has_tables || !c->has_nlist) return mdb_fail(c, MDB_ERR_STATE, "mdb_run: box, tables and list must be set");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    CUDA_TRY(c, cudaMemsetAsync(c->counters + CNT_OOB_TOTAL, 0, sizeof(int), c->stream));
    for (int s = 0; s < nsteps; s++) {
        int rc = step_nosync(c, itime0 + s, it0, nb_uptab, h, s == 0, s == nsteps - 1);
        if (rc < 0) return rc;
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->h_counters, c->counters, sizeof(int) * CNT__N, cudaMemcpyDeviceToHost, c->stream));
    c->run_pending = true;
    return MDB_OK;
}

extern "C" int mdb_run(mdb_ctx *c, int itime0, int nsteps, int it0, int nb_uptab, double h)
{
    int rc = mdb_run_async(c, itime0, nsteps, it0, nb_uptab, h);
    if (rc < 0) return rc;
    return mdb_sync(c);
}

extern "C" int mdb_step(mdb_ctx *c, int itime, int it0, int nb_uptab, double h)
{
    return mdb_run(c, itime, 1, it0, nb_uptab, h);
}

// ------------------------------------------------------------------------------------
// The time loop of the GMD method with its step-size and list-period schedules (Appshell/MD_Method_GenericMD_GPU.F90:351-361)
// and the displacement-limited step of Predictor_DEV (CommonGPU/MD_DiffScheme_GPU.F90:633-655).
// ------------------------------------------------------------------------------------
bool mdb_tile_guard_wanted(const mdb_ctx *c)
{
    if (!c->tiled.active || !c->tiled.use_classes || !c->dsr) return false;
    if (c->opt_tile_guard >= 0) return c->opt_tile_guard == 1;
    return mdb_stopping_on(c) || c->var_step;
}
int mdb_sched_nb_uptab(const mdb_sched *s, int itime, int it0)
{
    if (s->nb_dbitab <= 0) return s->nb_uptabmi;
    const int v = s->nb_uptabmi * ((itime - it0 + 1) / s->nb_dbitab + 1); // NB_UPTABMI*(int((ITIME-IT0+1)/NB_DBITAB)+1) :359
    return v > s->nb_uptabmx ? s->nb_uptabmx : v;
}
bool mdb_sched_check_due(const mdb_sched *s, int itime, int it0)
{
    return s->ihdup < 0 && (itime - it0 + 1) % (-s->ihdup) == 0; // MOD(ITIME-IT0+1,IABS(IHDUP)) == 0 :637
}
double mdb_sched_h1(const mdb_sched *s, int itime, int it0, double h)
{
    if (s->ihdup <= 0) return h;
    const double v = s->hmi * (double)((itime - it0 + 1) / s->ihdup + 1);  // HMI*(int((ITIME-IT0+1)/IHDUP)+1) :354
    return v > s->hmx ? s->hmx : v;
}

// the halving loop of Predictor_DEV (:633-655) as one call: the largest TH = HMX 2^-k for which no active atom would move
// further than DMX in the predictor step (collective in slab-decomposed runs)
extern "C" int mdb_timestep_limit(mdb_ctx *c, double hmx, double dmx, double *h)
{
    if (!c || !h) return mdb_fail(c, MDB_ERR_ARG, "mdb_timestep_limit: null argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_timestep_limit: mdb_box_set first");
    if (!(hmx > 0.0)) return mdb_fail(c, MDB_ERR_ARG, "mdb_timestep_limit: HMX must be positive");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    if (c->dd_on) return mdb_dd_timestep(c, hmx, dmx, h);
    int rc = mdb_timestep_mask_launch(c, hmx, dmx * dmx);
    if (rc < 0) return rc;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return mdb_timestep_from_mask(c, (unsigned)c->h_counters[CNT_SCRATCH], hmx, h);
}

extern "C" int mdb_run_sched(mdb_ctx *c, int itime0, int nsteps, int it0, const mdb_sched *s, double *h, double *time_s)
{
    if (!c || !s || !h) return mdb_fail(c, MDB_ERR_ARG, "mdb_run_sched: null argument");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_run_sched: use mdb_dd_run_sched in slab-decomposed runs");
    if (!c->has_box || !c->has_tables || !c->has_nlist) return mdb_fail(c, MDB_ERR_STATE, "mdb_run_sched: box, tables and list must be set");
    if (s->nb_uptabmi < 1 || (s->ihdup != 0 && !(s->hmx > 0.0))) return mdb_fail(c, MDB_ERR_ARG, "mdb_run_sched: bad schedule");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    CUDA_TRY(c, cudaMemsetAsync(c->counters + CNT_OOB_TOTAL, 0, sizeof(int), c->stream));
    c->var_step = s->ihdup < 0;
    double hh = *h, t = time_s ? *time_s : 0.0;
    int rc = MDB_OK;
    for (int k = 0; k < nsteps && rc >= 0; k++) {
        const int itime = itime0 + k;
        hh = mdb_sched_h1(s, itime, it0, hh);
        if (mdb_sched_check_due(s, itime, it0)) {
            if ((rc = mdb_timestep_limit(c, s->hmx, s->dmx, &hh)) < 0) break;
        }
        // the step may differ from its neighbours': EPC friction and the corrector close every step (nothing rides in front of
        // the next predictor), and CheckTimestep sees the corrected velocities as in the reference
        rc = step_nosync(c, itime, it0, mdb_sched_nb_uptab(s, itime, it0), hh, true, true);
        t += hh; // TIME = TIME + H :367 (here in seconds)
    }
    c->var_step = false;
    if (rc < 0) return rc;
    *h = hh;
    if (time_s) *time_s = t;
    CUDA_TRY(c, cudaMemcpyAsync(c->h_counters, c->counters, sizeof(int) * CNT__N, cudaMemcpyDeviceToHost, c->stream));
    c->run_pending = true;
    return mdb_sync(c);
}

// mdb_stop.cuh -- electronic stopping, global- and local-density models (ST_MOD_GDEN_KERNEL / ST_MOD_LDEN_KERNEL and their
// ELOSS twins, LocalTempCtrlMeths/Stopping/MD_ST_Coupling_GPU.F90:431-536,604-717,835-1132): parameters and the per-atom arithmetic, shared by the stand-alone
// kernel (mdb_cascade.cu) and the fused end-of-step kernel (mdb_step.cu).
#pragma once
#include "mdb_internal.cuh"

struct StopParams {
    int on, ne, nk, ng;
    int local;             // 0: global-density model (ST_MOD_GDEN_KERNEL :431-536), 1: local-density model (ST_MOD_LDEN_KERNEL :604-717)
    int enable[MDB_MXGROUP];
    double mden[MDB_MXGROUP], cm2[MDB_MXGROUP];
    double lv[MDB_MXGROUP * MDB_MXGROUP];  // local model: LVOL(i,j) = 4 pi / 3 NB_RM(i,j)^3 at i + ng*j (Reset_STMOD_DEV :412)
    int kpair[MDB_MXGROUP * MDB_MXGROUP]; // 1-based table index for (moving type, medium type) at i + ng*j
};
// nbc: local model with several types: neighbours of every type per atom, [g * n + i], recounted from INDI after every rebuild;
// eloss: inelastic energy loss per atom accumulated over the steps, ORIGINAL order (hm_SAVEELOSS, :1248-1262)
struct StopState {
    StopParams P;
    double *etab = nullptr, *stab = nullptr;
    int *nbc = nullptr; long long nbc_gen = -1;
    double *eloss = nullptr; int save_eloss = 0;
};

// FP -= FF V/|V| for one active atom i of type kk (0-based) with velocity v; returns false when nothing is to be done.
// kvois / nbc: the atom's neighbour count(s) for the local-density model; *loss (optional) receives FF |V| DT (:943, :1130)
__device__ __forceinline__ bool stop_force(const StopParams &S, const double *__restrict__ etab, const double *__restrict__ stab, int kk,
                                           double vx, double vy, double vz, double &fx, double &fy, double &fz,
                                           const int *__restrict__ kvois, const int *__restrict__ nbc, int i, int n, double dt, double *loss)
{
    if (loss) *loss = 0.0;
    if (S.enable[kk] <= 0) return false;
    double vv = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
    const double ek = __dmul_rn(S.cm2[kk], vv);                                           // EK = CM2(KK)*VV :509
    const double emin = etab[0], emax = etab[S.ne - 1];
    if (!(ek >= emin && ek <= emax)) return false;                                        // :511
    const int iiw = S.local ? kvois[i] : 1;
    if (iiw <= 0) return false;                                                           // IIW .gt. 0 :688
    const double de = __dsub_rn(etab[1], etab[0]);
    const double deinv = __ddiv_rn(1.0, de);                                               // DEINV :482
    const int ik = (int)__dmul_rn(__dsub_rn(ek, emin), deinv);                             // 0-based IK-1 :512
    double ff = 0.0;
    for (int ig = 0; ig < S.ng; ig++) {                                                   // :516-520 / :701-705
        const int kp = S.kpair[kk + S.ng * ig] - 1;
        double w;
        if (S.local) {
            // DEN(IG)*ILV(KK1+IG), ILV = 1/LV(kk,ig)/(ETAB(2)-ETAB(1)) :669; DEN(IG) = list neighbours of type IG :690-694
            const double den = (double)(S.ng == 1 ? iiw : nbc[(size_t)ig * n + i]);
            w = __dmul_rn(den, __ddiv_rn(__ddiv_rn(1.0, S.lv[kk + S.ng * ig]), de));
        } else {
            w = __ddiv_rn(S.mden[ig], de);                                                // SK(IG) = MDEN/(ETAB(2)-ETAB(1)) :487
        }
        const double s0 = stab[ik + (size_t)S.ne * kp], s1 = stab[ik + 1 + (size_t)S.ne * kp];
        const double lin = __dadd_rn(__dmul_rn(__dsub_rn(ek, etab[ik]), s1), __dmul_rn(__dsub_rn(etab[ik + 1], ek), s0));
        ff = __dadd_rn(ff, __dmul_rn(w, lin));
    }
    vv = sqrt(vv);
    fx = __dsub_rn(fx, __ddiv_rn(__dmul_rn(ff, vx), vv));                                  // FP = FP - FF*V/|V| :523-525
    fy = __dsub_rn(fy, __ddiv_rn(__dmul_rn(ff, vy), vv));
    fz = __dsub_rn(fz, __ddiv_rn(__dmul_rn(ff, vz), vv));
    if (loss) *loss = __dmul_rn(__dmul_rn(ff, vv), dt);                                    // ELOSS = FF*VV*DT :943
    return true;
}

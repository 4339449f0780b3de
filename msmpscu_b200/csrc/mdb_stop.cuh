// mdb_stop.cuh -- electronic stopping, global-density model (ST_MOD_GDEN_KERNEL,
// LocalTempCtrlMeths/Stopping/MD_ST_Coupling_GPU.F90:431-536): parameters and the per-atom arithmetic, shared by the stand-alone
// kernel (mdb_cascade.cu) and the fused end-of-step kernel (mdb_step.cu).
#pragma once
#include "mdb_internal.cuh"

struct StopParams {
    int on, ne, nk, ng;
    int enable[MDB_MXGROUP];
    double mden[MDB_MXGROUP], cm2[MDB_MXGROUP];
    int kpair[MDB_MXGROUP * MDB_MXGROUP]; // 1-based table index for (moving type, medium type) at i + ng*j
};
struct StopState { StopParams P; double *etab = nullptr, *stab = nullptr; };

// FP -= FF V/|V| for one active atom of type kk (0-based) with velocity v: returns false when nothing is to be done
__device__ __forceinline__ bool stop_force(const StopParams &S, const double *__restrict__ etab, const double *__restrict__ stab, int kk,
                                           double vx, double vy, double vz, double &fx, double &fy, double &fz)
{
    if (S.enable[kk] <= 0) return false;
    double vv = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
    const double ek = __dmul_rn(S.cm2[kk], vv);                                           // EK = CM2(KK)*VV :509
    const double emin = etab[0], emax = etab[S.ne - 1];
    if (!(ek >= emin && ek <= emax)) return false;                                        // :511
    const double deinv = __ddiv_rn(1.0, __dsub_rn(etab[1], etab[0]));                      // DEINV :482
    const int ik = (int)__dmul_rn(__dsub_rn(ek, emin), deinv);                             // 0-based IK-1 :512
    double ff = 0.0;
    for (int ig = 0; ig < S.ng; ig++) {                                                   // :516-520
        const int kp = S.kpair[kk + S.ng * ig] - 1;
        const double sk = __ddiv_rn(S.mden[ig], __dsub_rn(etab[1], etab[0]));              // SK(IG) = MDEN/(ETAB(2)-ETAB(1)) :487
        const double s0 = stab[ik + (size_t)S.ne * kp], s1 = stab[ik + 1 + (size_t)S.ne * kp];
        const double lin = __dadd_rn(__dmul_rn(__dsub_rn(ek, etab[ik]), s1), __dmul_rn(__dsub_rn(etab[ik + 1], ek), s0));
        ff = __dadd_rn(ff, __dmul_rn(sk, lin));
    }
    vv = sqrt(vv);
    fx = __dsub_rn(fx, __ddiv_rn(__dmul_rn(ff, vx), vv));                                  // FP = FP - FF*V/|V| :523-525
    fy = __dsub_rn(fy, __ddiv_rn(__dmul_rn(ff, vy), vv));
    fz = __dsub_rn(fz, __ddiv_rn(__dmul_rn(ff, vz), vv));
    return true;
}

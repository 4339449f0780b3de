// replay_shims.cpp -- a C++ host that calls libmdpscu_b200.so exactly as fortran/mdb_shims.F90 does, procedure by
// procedure and in the order the unchanged MDPSCU shell would (no Fortran compiler exists in the build image, so this
// is the executed stand-in for the shim layer: same entry points, same argument meaning, same sequence):
//
//   Initialize_DEVICES                       mdb_device_count, mdb_ctx_create
//   Initialize_Globle_Variables_DEV          mdb_box_set + CopyAllFrom_Host_to_Devices (mdb_state_upload x 6)
//   Register_ForceClass / Init_Forcetable_Dev  pIniForcetable -> mdb_tables_set      (MD_ForceClass_Register_GPU.F90:363-442,483-559)
//   Initialize_NeighboreList_DEV             mdb_nlist_init
//   Cal_NeighBoreList_DEV                    mdb_nlist_build
//   pCalForce, pCalEpot0, pCalPTensor, pCalEDen, pCalAVStress   mdb_force(flags), mdb_atomic_stress_host
//   Do_ResetParam_DEV                        mdb_epc_set
//   For_One_Step x nsteps                    Predictor_DEV -> [Cal_NeighBoreList_DEV] -> pCalForce -> Do_EPCForce_DEV ->
//                                            Correction_DEV      (Appshell/MD_Method_GenericMD_GPU.F90:596-627)
//   CalEKin_DEV, Cal_GlobalT_DEV             mdb_ekin, mdb_global_t
//   CopyOut_SimBox_DEV                       mdb_state_download x 6 (original order)
//   Copyout_NeighboreList_DEV, GetCellInform mdb_nlist_copyout, mdb_nlist_cellinfo
//   pClrForcetable, Clear_NeighboreList_DEV, End_DEVICES   mdb_tables_clear, mdb_nlist_clear, mdb_ctx_destroy
//
// usage: replay_shims <cfg.bin> <out.bin> <nsteps>     cfg.bin: int n, double boxlow[3], zl[3], mass, ru, nbfac, then XP(n,3), XP1(n,3)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../include/mdpscu_b200.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int rc__ = (call);                                                                       \
        if (rc__ < 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, mdb_last_error(ctx)); return 2; } \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 4) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    int n = 0;
    double hdr[9];
    if (fread(&n, sizeof(int), 1, f) != 1 || fread(hdr, sizeof(double), 9, f) != 9) return 1;
    std::vector<double> xp(3 * (size_t)n), xp1(3 * (size_t)n), dis(3 * (size_t)n, 0.0), fp(3 * (size_t)n, 0.0);
    if (fread(xp.data(), sizeof(double), xp.size(), f) != xp.size() || fread(xp1.data(), sizeof(double), xp1.size(), f) != xp1.size()) return 1;
    fclose(f);
    const int nsteps = atoi(argv[3]);
    const double *boxlow = hdr, *zl = hdr + 3, mass = hdr[6], ru = hdr[7], nbfac = hdr[8];
    std::vector<int> ityp(n, 1), statu(n, 1);
    mdb_ctx *ctx = nullptr;

    // ---- Initialize_DEVICES(FIRSTDEV = 0, NDEV = 1)
    if (mdb_device_count() < 1) { fprintf(stderr, "no CUDA device\n"); return 3; }
    if (mdb_ctx_create(0, &ctx) < 0) return 3;
    // ---- Initialize_Globle_Variables_DEV(SimBox, CtrlParam)
    const int ifpd[3] = {1, 1, 1};
    CHECK(mdb_box_set(ctx, 1, n, boxlow, zl, nullptr, ifpd, 1, &mass));
    CHECK(mdb_state_upload(ctx, MDB_F_XP, xp.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_upload(ctx, MDB_F_XP1, xp1.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_upload(ctx, MDB_F_DIS, dis.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_upload(ctx, MDB_F_FP, fp.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_upload(ctx, MDB_F_ITYP, ityp.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_upload(ctx, MDB_F_STATU, statu.data(), MDB_ORDER_ORIGINAL));
    // ---- Register_ForceClass("EAM_TYPE") + Init_Forcetable_Dev: the potential library fills MDForceTable, pIniForcetable uploads it
    const int ntab = 10000, nembd = 10000, ptype[1] = {1};
    int nkind = 0, nkind1 = 0, kpair[1], kembd[1];
    std::vector<double> potr(ntab), fpotr(ntab), potb(ntab), fpotb(ntab), fembd(nembd), dfembd(nembd);
    double csi = 0.0, rhod = 0.0;
    CHECK(mdb_host_ftable_create(MDB_LIB_MARINICA_EAM2, 1, ptype, ntab, nembd, 20.0, ru, &nkind, &nkind1, kpair, kembd, potr.data(),
                                 fpotr.data(), potb.data(), fpotb.data(), fembd.data(), dfembd.data(), &csi, &rhod));
    CHECK(mdb_tables_set(ctx, MDB_POT_EAM, nkind, ntab, csi, potr.data(), fpotr.data(), potb.data(), fpotb.data(), nkind1, nembd, rhod,
                         fembd.data(), dfembd.data(), kpair, kembd, ru * ru));
    // ---- Initialize_NeighboreList_DEV, Cal_NeighBoreList_DEV
    const double nb_rm = nbfac * ru;
    const int mxkvois = 256;
    CHECK(mdb_nlist_init(ctx, &nb_rm, mxkvois));
    CHECK(mdb_nlist_build(ctx));
    // ---- the force-class slots
    double vt[9], vt0[9];
    CHECK(mdb_force(ctx, MDB_FORCE, nullptr));                 // pCalForce
    CHECK(mdb_force(ctx, MDB_EPOT, nullptr));                  // pCalEpot0
    CHECK(mdb_force(ctx, MDB_FORCE | MDB_VIRIAL, vt0));        // pCalPTensor
    CHECK(mdb_force(ctx, MDB_DEN, nullptr));                   // pCalEDen
    std::vector<double> avp(9 * (size_t)n);
    CHECK(mdb_atomic_stress_host(ctx, avp.data(), MDB_ORDER_ORIGINAL)); // pCalAVStress
    // ---- Do_ResetParam_DEV: EPC on group 1, reference defaults (Common/MD_TypeDef_EPCCtrl.F90:27-37)
    const int enable[1] = {1};
    const double te[1] = {300.0}, alpha[1] = {1.0e-12}, cut[1] = {0.1}, he[1] = {100.0 * 1.60219e-12};
    CHECK(mdb_epc_set(ctx, enable, te, alpha, cut, he));
    // ---- For_One_Step x nsteps, kernel by kernel as the shell calls them
    double h = 0.5e-15;
    const double hmx = 0.5e-15, dmx = 0.05e-8;                                      // &STEPSIZE flag -1, hmx 0.5 fs, dmx 0.05 A
    const int it0 = 1, nb_uptab = 10, ihdup = -2;
    for (int itime = 0; itime < nsteps; itime++) {
        if ((itime - it0 + 1) % (-ihdup) == 0) CHECK(mdb_timestep_limit(ctx, hmx, dmx, &h)); // Predictor_DEV, IHDUP < 0 (:633-655)
        CHECK(mdb_predict(ctx, h));                                               // Predictor_DEV
        if ((itime - it0) % nb_uptab == 0) CHECK(mdb_nlist_build(ctx));           // Cal_NeighBoreList_DEV
        CHECK(mdb_force(ctx, MDB_FORCE, nullptr));                                // CalForce_ForceClass
        CHECK(mdb_epc_apply(ctx));                                                // Do_EPCForce_DEV
        CHECK(mdb_correct(ctx, h));                                               // Correction_DEV
    }
    // ---- CalEKin_DEV, Cal_GlobalT_DEV, energies, pressure tensor
    double curt = 0.0;
    CHECK(mdb_ekin(ctx));
    CHECK(mdb_global_t(ctx, &curt));
    CHECK(mdb_force(ctx, MDB_FORCE | MDB_EPOT | MDB_VIRIAL, vt));
    // ---- CopyOut_SimBox_DEV
    std::vector<double> epot(n), ekin(n);
    CHECK(mdb_state_download(ctx, MDB_F_XP, xp.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_download(ctx, MDB_F_XP1, xp1.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_download(ctx, MDB_F_FP, fp.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_download(ctx, MDB_F_DIS, dis.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_download(ctx, MDB_F_EPOT, epot.data(), MDB_ORDER_ORIGINAL));
    CHECK(mdb_state_download(ctx, MDB_F_EKIN, ekin.data(), MDB_ORDER_ORIGINAL));
    // ---- Copyout_NeighboreList_DEV, GetCellInform
    std::vector<int> kvois(n), indi((size_t)n * mxkvois);
    CHECK(mdb_nlist_copyout(ctx, kvois.data(), indi.data(), MDB_ORDER_ORIGINAL));
    int ncell[3], tnc = 0, mxnac = 0;
    CHECK(mdb_nlist_cellinfo(ctx, ncell, &tnc, &mxnac));
    long long ksum = 0;
    for (int i = 0; i < n; i++) ksum += kvois[i];
    // ---- results
    f = fopen(argv[2], "wb");
    if (!f) return 1;
    const double scal[4] = {curt, (double)ksum, (double)tnc, (double)mxnac};
    fwrite(scal, sizeof(double), 4, f);
    fwrite(vt, sizeof(double), 9, f);
    fwrite(xp.data(), sizeof(double), xp.size(), f);
    fwrite(xp1.data(), sizeof(double), xp1.size(), f);
    fwrite(fp.data(), sizeof(double), fp.size(), f);
    fwrite(epot.data(), sizeof(double), epot.size(), f);
    fclose(f);
    // ---- pClrForcetable, Clear_NeighboreList_DEV, End_DEVICES
    CHECK(mdb_tables_clear(ctx));
    CHECK(mdb_nlist_clear(ctx));
    mdb_ctx_destroy(ctx);
    printf("replay ok: n=%d T=%.6f K, %lld list entries, %d cells, launches not counted here\n", n, curt, ksum, tnc);
    return 0;
}

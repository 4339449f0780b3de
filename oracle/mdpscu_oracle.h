/*
 * mdpscu_oracle.h -- CPU restatement of the MDPSCU tabulated EAM/FS hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and there only as the checker or the CPU arm being timed.
 *
 * The reference (CUDA-Fortran, PGI) cannot be compiled in this environment (no
 * Fortran compiler), so this is a scalar C restatement of its arithmetic.  Every
 * function cites the reference file:line it follows (paths relative to the
 * reference root).  It is PINNED against the reference's own known-answer files
 * (examples/NEB_Test/GMD/{React,Product}P0000_0001.0000, 2001 atoms, forces and
 * per-atom potential to 9 digits) by tests/test_oracle_golden.py.
 *
 * Conventions (same as the reference): CGS units, Fortran column-major arrays,
 * XP(N,3) means x[0..N), y[N..2N), z[2N..3N); atom / table indices inside INDI
 * and KPAIR/KEMBD are 1-based; ITYP is 1-based.
 */
#ifndef MDPSCU_ORACLE_H
#define MDPSCU_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MXGROUP 10

/* physical constants: MSMLIB/sor/Common/MSM_Const.F90:67,74,82-83 */
#define ORC_A2CM  1.0e-8
#define ORC_CM2A  1.0e8
#define ORC_AU2G  1.66053e-24
#define ORC_KB    1.38054e-16
#define ORC_EVERG 1.60219e-12

/* STATU bits: MDLIB/sor/Common/MD_Const.F90:79-99 */
#define ORC_STATU_ACTIVE    1
#define ORC_STATU_FIXPOSX   2
#define ORC_STATU_FIXPOSY   4
#define ORC_STATU_FIXPOSZ   8
#define ORC_STATU_FIXPOS    14
#define ORC_STATU_FIXVELX   16
#define ORC_STATU_FIXVELY   32
#define ORC_STATU_FIXVELZ   64
#define ORC_STATU_OUTOFBOX  65536
#define ORC_STATU_REFLECT   131072
#define ORC_STATU_TRANSMIT  262144
#define ORC_STATU_PASSBOUND 524288

/* potential libraries (table generators) */
#define ORC_LIB_MARINICA_EAM2 1 /* Potentials/EAM_WW_Marinica_JPCM25_2013, "EAM2": id 1 = W-W     */
#define ORC_LIB_BONNY_EAM1    2 /* Potentials/EAM_WHeH_Bonny_JPCM26_2014, "EAM1": ids 1..9         */
#define ORC_LIB_MARINICA_EAM3 3
#define ORC_LIB_MARINICA_EAM4 4
#define ORC_LIB_ACKLAND_FS_W  5 /* Potentials/EM_TB_WangJun_W-HE_2010, FS_TYPE: id 1 = W-W (Ackland, Thetford 1987) */

#define ORC_POT_EAM 0
#define ORC_POT_FS  1

void orc_set_threads(int n);
int  orc_get_max_threads(void);

/* ---- potential callbacks (NNFORCE / NEFORCE / EMBDF of MD_TypeDef_ForceTable.F90:46-78) */
int  orc_pot_nn  (int lib, int id, double r_cm, double *half_v, double *minus_dv);
int  orc_pot_rho (int lib, int id, double r_cm, double *rho, double *minus_drho); /* returns 0 if id has no NEFORCE */
int  orc_pot_embd(int lib, int id, double rho, double *f, double *df);            /* returns 0 if id has no EMBDF   */

/* ---- table generation: MD_TypeDef_ForceTable.F90:530-622,890-1056,1151-1230
 * ptype[(i-1) + ng*(j-1)] = PTYPE(i,j).  Tables come back in Fortran layout
 * T(NKIND,NTAB): element (k,i) at [(i-1)*nkind + (k-1)]; the caller allocates
 * ng*ng*ntab doubles per pair table and ng*nembd per embedding table.
 * kpair[(i-1)+ng*(j-1)] = KPAIR(i,j) (1-based), kembd[i-1] = KEMBD(i).        */
int  orc_ftable_build(int lib, int ng, const int *ptype, int ntab, int nembd,
                      double rhoscal, double rmax,
                      int *nkind, int *nkind1, int *kpair, int *kembd,
                      double *potr, double *fpotr, double *potb, double *fpotb,
                      double *fembd, double *dfembd,
                      double *csi, double *rhod);

/* ---- cells + neighbour list
 * orc_ncell: MD_NeighborsList_GPU.F90:289-298 */
void orc_ncell(const double zl[3], double nb_rm_max, int ncell[3]);

/* The device rule (cell id kernel :759-819, host linked-cell sort :1490-1570,
 * out-of-box atoms :1624-1637, fp32 list kernel :907-1199).
 * Input arrays are in ORIGINAL order; outputs describe the cell-sorted order.
 *   inc[N]   cell id per ORIGINAL atom (>0, -1 flagged out-of-box, -2 outside)
 *   gid[N]   sorted position -> original 1-based id
 *   nac/naac/ia1th[nc_total]
 *   kvois[N], indi[N*mxkvois] (column-major INDI(N,mxkvois), sorted 1-based ids)
 *   statu is updated in place (atoms found outside get OUTOFBOX), still ORIGINAL order.
 * returns the number of out-of-box atoms; *nn_max = largest un-truncated count */
int  orc_nlist_build_dev(int nbox, int napb,
                         const double *xp, const int *ityp, int *statu,
                         const double boxlow[3], const double zl[3], const int ifpd[3],
                         const double boxshape[9], int ng, const double *nb_rm /* ng*ng, cm */,
                         int mxkvois,
                         int ncell[3], int *inc, int *gid, int *nac, int *naac, int *ia1th,
                         int *kvois, int *indi, int *nn_max);

/* The host rule Cal_NeighboreList2C (Common/MD_NeighborsList.F90:396-633): fp64, '<',
 * single box, list in ORIGINAL indices, INDI(N,mxkvois) column-major.
 * returns <0 on overflow (the reference stops), else 0. */
int  orc_nlist_build_cpu(int n, const double *xp, const int *ityp, const int *statu,
                         const double boxlow[3], const double zl[3], const int ifpd[3],
                         const double boxshape[9], int ng, const double *nb_rm,
                         int mxkvois, int *kvois, int *indi);

/* ---- forces on a given list (arrays and list indices in the same order).
 * Table struct is passed flat (Fortran layout as produced by orc_ftable_build). */
typedef struct orc_tables {
    int pot_type;              /* ORC_POT_EAM | ORC_POT_FS                     */
    int ng, nkind, ntab, nkind1, nembd;
    double csi, rhod, ru2max;  /* ru2max = maxval(RU*RU), EAM_GPU:261-262       */
    const int *kpair, *kembd;
    const double *potr, *fpotr, *potb, *fpotb, *fembd, *dfembd;
} orc_tables;

/* pass 1, PRECALFOR_KERNEL (MD_EAM_ForceTable_GPU.F90:474-545; FS twin MD_FS_ForceTable_GPU.F90:341-516) */
void orc_force_pass1(int n, int ia0, int npart, const double *xp, const int *ityp, const int *statu,
                     const int *kvois, const int *indi, int ldindi,
                     const double zl[3], const int ifpd[3], const double boxshape[9],
                     const orc_tables *t, double *den);
/* pass 2, CALFORCE_KERNEL (:735-825) when vtensor==NULL, CALPTENSOR_KERNEL (:1128-1239) otherwise
 * (vtensor[9] column-major 3x3, NOT yet divided by the number of boxes). */
void orc_force_pass2(int n, int ia0, int npart, const double *xp, const int *ityp, const int *statu,
                     const int *kvois, const int *indi, int ldindi,
                     const double zl[3], const int ifpd[3], const double boxshape[9],
                     const orc_tables *t, const double *den, double *fp, int ldfp, double *vtensor);
/* CAL_EAM_AtomicStress_KERNEL (:1775-1925): AP(.,9), q = 3*(a-1)+(b-1) */
void orc_force_avstress(int n, int ia0, int npart, const double *xp, const int *ityp, const int *statu,
                        const int *kvois, const int *indi, int ldindi,
                        const double zl[3], const int ifpd[3], const double boxshape[9],
                        const orc_tables *t, const double *den, double *ap, int ldap);
/* CALEPOT_KERNEL (:1563-1634) */
void orc_force_epot(int n, int ia0, int npart, const double *xp, const int *ityp, const int *statu,
                    const int *kvois, const int *indi, int ldindi,
                    const double zl[3], const int ifpd[3], const double boxshape[9],
                    const orc_tables *t, double *epot);

/* ---- integrator: MD_DiffScheme_GPU.F90:293-381 (predictor), :724-756 (corrector), :885-899 (ekin) */
void orc_predictor(int n, double *xp, double *xp1, const double *fp, double *dis, int *statu,
                   const int *ityp, const double *cm, double h,
                   const double boxlow[3], const double boxup[3], const double zl[3], const int ifpd[3]);
void orc_corrector(int n, double *xp1, const double *fp, const int *statu,
                   const int *ityp, const double *cm, double h);
void orc_ekin(int n, const double *xp1, const int *statu, const int *ityp, const double *cm, double *ekin);

/* ---- electron-phonon coupling: MD_EP_Coupling_GPU.F90:384-393 (params), :468-490 (kernel) */
void orc_epc(int n, const double *xp1, double *fp, const int *statu, const int *ityp, int ng,
             const int *enable, const double *cm, const double *te, const double *alpha,
             const double *cut, const double *he);

/* ---- whole-step driver (Appshell/MD_Method_GenericMD_GPU.F90:496-659): predictor ->
 * rebuild if MOD(ITIME-IT0,NB_UPTAB)==0 -> force -> EPC -> corrector, in cell-sorted order. */
typedef struct orc_md orc_md;
orc_md *orc_md_create(int nbox, int napb, const double *xp, const double *xp1, const int *ityp,
                      const int *statu, int ng, const double *cm,
                      const double boxlow[3], const double zl[3], const int ifpd[3],
                      const double *nb_rm, int mxkvois, const orc_tables *t);
void  orc_md_destroy(orc_md *m);
void  orc_md_set_epc(orc_md *m, const int *enable, const double *te, const double *alpha,
                     const double *cut, const double *he);
int   orc_md_rebuild(orc_md *m);                 /* neighbour list on current positions        */
void  orc_md_force(orc_md *m, int with_virial);  /* pass1 + pass2 (+virial)                     */
void  orc_md_reorder_nearest(orc_md *m, int nearest); /* Reorder_NeighBoreList_Nearest_Dev, in place */
void  orc_md_epot(orc_md *m);
void  orc_md_avstress(orc_md *m, double *ap);    /* ap[n*9], ORIGINAL order, column-major           */
void  orc_damping(int n, double *xp1, const double *fp, const int *statu); /* DAMPING_KERNEL, MD_DiffScheme_GPU.F90:125-185 */
void  orc_md_damping(orc_md *m);
int   orc_md_dyndamp(orc_md *m, int mxnumsteps, double h, double minepot, double *delepot_out); /* :1809-1860 */
int   orc_md_step(orc_md *m, int itime, int it0, int nb_uptab, double h); /* returns 1 if rebuilt */
/* copy out in ORIGINAL order; any pointer may be NULL */
void  orc_md_get(orc_md *m, double *xp, double *xp1, double *fp, double *epot, double *ekin,
                 double *dis, int *statu, int *gid, double *vtensor);
/* steepest-descent quench, Do_Steepest0_Forsteps_DEV (CommonGPU/MD_SteepestScheme_GPU.F90:20-153): lengths in cm,
 * minepot in erg; returns IFLAG (0 not converged, >0 iteration of convergence, -1 converged at the first step) */
int   orc_md_steepest0(orc_md *m, int mxnumsteps, double alpha0, double maxdis, double mindis, double minepot,
                       double *maxmove_out, double *delepot_out);
/* Do_CG0/CG1_Forsteps_DEV (CommonGPU/MD_CGScheme_GPU.F90:16-276) and Do_Steepest1_Forsteps_DEV
 * (CommonGPU/MD_SteepestScheme_GPU.F90:157-260); return values as documented at the definitions */
int   orc_md_cg(orc_md *m, int mxnumsteps, int lsearch, double maxdis, double mindis, double minepot, double *delepot_out);
int   orc_md_steepest1(orc_md *m, int mxnumsteps, double maxdis, double mindis, double *delepot_out);
/* Cal_GlobalT_DEV :1042-1064, VelScaling_DEV :1262-1446, CheckTimestep_DEV :1066-1258 (MD_DiffScheme_GPU.F90) */
/* DO_LBFGSB_FORSTEPS_DEV (CommonGPU/MD_LBFGSScheme_GPU.F90:177-388) = SETULB without bounds; see the definition */
int   orc_md_lbfgsb(orc_md *m, int mxnumsteps, int msave, double factr, double pgtol, int *nfg_out, int *niter_out);
/* Thermalizing_MC_DEV :1608-1805 with Philox4x32-10 uniforms (see the definition) */
void  orc_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4]);
void  orc_md_thermalize(orc_md *m, double ti, unsigned long long seed, unsigned draw);
double orc_md_global_t(orc_md *m);
int   orc_md_vel_scaling(orc_md *m, double dt);          /* -1: a box with zero kinetic energy (the reference stops) */
int   orc_md_check_timestep(orc_md *m, double th, double h2s2, double mxd2);
int   orc_md_natom(orc_md *m);
const int *orc_md_kvois(orc_md *m);
const int *orc_md_indi(orc_md *m);
void  orc_md_ncell(orc_md *m, int ncell[3]);


/* ---- the reference's CPU path (see the definitions): CPU lists, Newton-3 force, step loop in C */
int   orc_nlist_build_cpu_half(int n, const double *xp, const int *ityp, const int *statu,
                               const double boxlow[3], const double zl[3], const int ifpd[3],
                               const double boxshape[9], int ng, const double *nb_rm, int mxkvois,
                               int *kvois, int *indi);
void  orc_force_newton3(int n, const double *xp, const int *ityp, const int *kvois, const int *indi, int ldindi,
                        const double zl[3], const int ifpd[3], const double bs[9], const orc_tables *t,
                        double *den, double *er, double *fp, double *epot, double *vtensor);
typedef struct orc_cpu orc_cpu;
orc_cpu *orc_cpu_create(int n, const double *xp, const double *xp1, const int *ityp, const int *statu, int ng,
                        const double *cm, const double boxlow[3], const double zl[3], const int ifpd[3],
                        const double *nb_rm, int mxkvois, const orc_tables *t, int half);
void  orc_cpu_destroy(orc_cpu *m);
void  orc_cpu_set_epc(orc_cpu *m, const int *enable, const double *te, const double *alpha, const double *cut, const double *he);
int   orc_cpu_rebuild(orc_cpu *m);
void  orc_cpu_force(orc_cpu *m, int with_epot);
int   orc_cpu_run(orc_cpu *m, int itime0, int nsteps, int it0, int nb_uptab, double h);
double orc_cpu_harmil(orc_cpu *m);
void  orc_cpu_get(orc_cpu *m, double *xp, double *xp1, double *fp, double *epot);
const int *orc_cpu_kvois(orc_cpu *m);

#ifdef __cplusplus
}
#endif
#endif

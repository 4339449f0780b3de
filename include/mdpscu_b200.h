/*
 * mdpscu_b200.h -- C ABI of the B200-native MDPSCU hot path (libmdpscu_b200.so).
 *
 * One context = one GPU = one host process/rank.  The context owns the device "box"
 * (the reference's dm_WorkSpace), the neighbour list (dm_Neighbors), the force tables
 * (dm_FTableWS) and a CUDA stream on which every kernel of the path is launched.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference root, MDLIB/sor/ prefix dropped).  All calls return an int status:
 *   0 = ok, >0 = a count the caller may act on (out-of-box atoms, list overflow),
 *   <0 = MDB_ERR_*.  Nothing here prints, reads stdin or stops the process
 * (the reference does all three: CommonGPU/MD_NeighborsList_GPU.F90:1506-1524).
 *
 * Array conventions are the reference's: Fortran column-major, XP(N,3) = x[0..N) y z,
 * 1-based atom ids / table ids / ITYP, CGS units.  "ORIGINAL" order = the order the
 * atoms had when first uploaded (SimMDBox order); "CELL" order = current device order
 * (cell-sorted; changes at every neighbour rebuild; GID maps CELL -> ORIGINAL).
 */
#ifndef MDPSCU_B200_H
#define MDPSCU_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MDB_MXGROUP 10 /* mp_MXGROUP, Common/MD_Const.F90 */

typedef struct mdb_ctx mdb_ctx;

/* ---- status codes */
#define MDB_OK               0
#define MDB_ERR_CUDA        -1
#define MDB_ERR_ARG         -2
#define MDB_ERR_STATE       -3  /* call order: box/tables/nlist not set                       */
#define MDB_ERR_UNSUPPORTED -4  /* e.g. non-identity BOXSHAPE in the fast force path           */
#define MDB_ERR_NOMEM       -5
#define MDB_ERR_NOGPU       -6  /* no CUDA device: the product path has no CPU fallback        */

/* ---- orders */
#define MDB_ORDER_ORIGINAL 0
#define MDB_ORDER_CELL     1

/* ---- fields (dm_WorkSpace / dm_Neighbors members, CommonGPU/MD_Globle_Variables_GPU.F90:158-194,
 *      CommonGPU/MD_NeighborsList_GPU.F90:75-85).  Shapes as the reference declares them. */
#define MDB_F_XP      0  /* double (N,3)  positions                                          */
#define MDB_F_XP1     1  /* double (N,3)  velocities                                         */
#define MDB_F_FP      2  /* double (N,3)  forces                                             */
#define MDB_F_DIS     3  /* double (N,3)  accumulated displacement                           */
#define MDB_F_EPOT    4  /* double (N)                                                       */
#define MDB_F_EKIN    5  /* double (N)                                                       */
#define MDB_F_DEN     6  /* double (N)    dF/drho (EAM) or -1/(2 sqrt(rho)) (FS)             */
#define MDB_F_ITYP    7  /* int (N)                                                          */
#define MDB_F_STATU   8  /* int (N)                                                          */
#define MDB_F_GID     9  /* int (N)       CELL position -> ORIGINAL id (1-based)             */
#define MDB_F_GIDINV 10  /* int (N)       ORIGINAL id -> CELL position (1-based)             */
#define MDB_F_IC     11  /* int (N)       cell id of each atom (CELL order), -1/-2 if outside */
#define MDB_F_KVOIS  12  /* int (N)                                                          */
#define MDB_F_INDI   13  /* int (N,mxKVOIS) neighbour ids, 1-based CELL-order ids            */
#define MDB_F_NAC    14  /* int (NC)      atoms per cell                                     */
#define MDB_F_NAAC   15  /* int (NC)      ACTIVE atoms per cell                              */
#define MDB_F_IA1TH  16  /* int (NC)      1-based id of the first atom of each cell          */
#define MDB_F_POS4   17  /* mdb_devptr only: the internal packed records double {x,y,z,den} per atom        */
#define MDB_F_D2MAX  18  /* mdb_devptr only: int holding the float bits of max |displacement since rebuild|^2 */
#define MDB_F__COUNT 19

/* ---- potential types: Register_ForceClass("EAM_TYPE"|"FS_TYPE"), CommonGPU/MD_ForceClass_Register_GPU.F90:363-442 */
#define MDB_POT_EAM 0
#define MDB_POT_FS  1

/* ---- mdb_force flags */
#define MDB_FORCE    1u  /* pCalForce  -> CALFORCE_EAM_Force_Table2A_DEV, CommonGPU/MD_EAM_ForceTable_GPU.F90:950      */
#define MDB_VIRIAL   2u  /* pCalPTensor-> CALPTENSOR_EAM_Force_Table2A_DEV, :1366 (forces + virial)                     */
#define MDB_EPOT     4u  /* pCalEpot0  -> UpdateEPOT_EAM_Force_Table2A_DEV, :1710                                       */
#define MDB_DEN      8u  /* pCalEDen   -> CALDEN_EAM_Force_Table2A_DEV, :891 (pass 1 only)                              */
#define MDB_NOPASS1 16u  /* with MDB_FORCE: DEN is already current (slab runs exchange it between the passes)          */

/* ------------------------------------------------------------------------------------
 * context / device selection
 * replaces Initialize_DEVICES + m_DEVICES (MSMLIB/sor/CommonGPU/MSM_MultiGPU_Basic.F90:571-597):
 * one rank per GPU instead of one host thread looping cudaSetDevice.
 * ---------------------------------------------------------------------------------- */
int         mdb_device_count(void);
int         mdb_ctx_create(int device_id, mdb_ctx **out);
void        mdb_ctx_destroy(mdb_ctx *ctx);
const char *mdb_last_error(const mdb_ctx *ctx);
const char *mdb_version(void);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL = the context's own */
int         mdb_ctx_set_stream(mdb_ctx *ctx, void *cuda_stream);
void       *mdb_ctx_stream(mdb_ctx *ctx);
/* SynchronizeDevices.  Also completes what mdb_run_async / mdb_state_download_async enqueued: returns the out-of-box
 * count of a pending mdb_run_async block (>= 0), else 0 */
int         mdb_sync(mdb_ctx *ctx);

/* ------------------------------------------------------------------------------------
 * device box: Initialize_Globle_Variables_DEV / Allocate_Working_Variables
 * (CommonGPU/MD_Globle_Variables_GPU.F90:585,767-857).  nbox boxes of natom_per_box atoms are
 * concatenated (MULTIBOX, :601-606).  boxshape is column-major 3x3.
 * ---------------------------------------------------------------------------------- */
int mdb_box_set(mdb_ctx *ctx, int nbox, int natom_per_box, const double boxlow[3], const double boxsize[3],
                const double boxshape[9], const int ifpd[3], int ngroup, const double mass[]);
int mdb_natom(const mdb_ctx *ctx);

/* CopyIn_SimBox_DEV / CopyOut_SimBox_DEV (CommonGPU/MD_SimBoxArray_GPU.F90:91-318) and the
 * CopyXXFrom_Devices_to_Host family (MD_Globle_Variables_GPU.F90:196-510), one field at a time.
 * host buffers have the reference shape of the field. */
int mdb_state_upload(mdb_ctx *ctx, int field, const void *host, int order);
int mdb_state_download(mdb_ctx *ctx, int field, void *host, int order);
/* the same CopyOut enqueued on the context's stream without waiting (the reference's CopyOut blocks the host thread,
 * MD_SimBoxArray_GPU.F90:225-318): `host` should be page-locked and holds the field after the next mdb_sync.  Uploads
 * never wait.  With two contexts on two streams, the transfers of one box overlap the steps of the other. */
int mdb_state_download_async(mdb_ctx *ctx, int field, void *host, int order);
/* device pointer with the reference shape (CELL order), for CUDA(-Fortran) code that reads
 * dm_WorkSpace / dm_Neighbors arrays directly (Analysis/, BoostMeths/ ...).  XP, DEN and INDI are
 * materialised from the packed internal layout on request and stay valid until the next
 * mdb_* call that changes them. */
void *mdb_devptr(mdb_ctx *ctx, int field);

/* ------------------------------------------------------------------------------------
 * force tables: pIniForcetable -> INITIALIZE_EAM_Force_Table_DEV
 * (CommonGPU/MD_EAM_ForceTable_GPU.F90:192-278; FS twin MD_FS_ForceTable_GPU.F90).
 * Tables exactly as type MDForceTable holds them (Common/MD_TypeDef_ForceTable.F90:117-155):
 * POTR,FPOTR,POTB,FPOTB(NKIND,NTAB), FEMBD,DFEMBD(NKIND1,NEMBD) column-major,
 * KPAIR(NG,NG) column-major, KEMBD(NG), CSI, RHOD; ru2max = maxval(RU*RU) (:261-262).
 * ---------------------------------------------------------------------------------- */
int mdb_tables_set(mdb_ctx *ctx, int pot_type, int nkind, int ntab, double csi,
                   const double *potr, const double *fpotr, const double *potb, const double *fpotb,
                   int nkind1, int nembd, double rhod, const double *fembd, const double *dfembd,
                   const int *kpair, const int *kembd, double ru2max);
int mdb_tables_clear(mdb_ctx *ctx); /* pClrForcetable -> Clear_EAM_Force_Table_DEV :343 */
/* Density-pass evaluations since the last call whose rho exceeded the embedding table (RHOMX = max(POTB)*RHOSCAL,
 * Common/MD_TypeDef_ForceTable.F90:1043-1048).  The reference reads past DFEMBD there (:535-541 has no bound test); this
 * library clamps to the zero pad (dF/drho = 0 for that atom) and counts the event, e.g. for close collisions in cascades:
 * raise RHOSCAL when it is not 0. */
int mdb_embed_overruns(mdb_ctx *ctx);

/* ------------------------------------------------------------------------------------
 * neighbour list: Initialize_NeighboreList_DEV / Cal_NeighBoreList_DEV / Copyout_NeighboreList_DEV /
 * GetCellInform / Clear_NeighboreList_DEV (CommonGPU/MD_NeighborsList_GPU.F90:116-202,241-396,1347-1732)
 * nb_rm(NG,NG) column-major in cm.  mdb_nlist_build returns the number of out-of-box atoms (they
 * are marked OUTOFBOX and parked at the end of the CELL order, :1624-1637) or <0.
 * mdb_nlist_overflow: atoms whose un-truncated count exceeded mxkvois (the reference truncates
 * silently and warns once, :1126-1131,1709-1718).
 * ---------------------------------------------------------------------------------- */
int mdb_nlist_init(mdb_ctx *ctx, const double *nb_rm, int mxkvois);
int mdb_nlist_build(mdb_ctx *ctx);
int mdb_nlist_copyout(mdb_ctx *ctx, int *kvois, int *indi, int order);
int mdb_nlist_cellinfo(const mdb_ctx *ctx, int ncell[3], int *nc_total, int *mxnac);
/* Reorder_NeighBoreList_Nearest_Dev(Nearest), MD_NeighborsList_GPU.F90:2016-2066 (kernel :1805-1946): in place, every atom
 * keeps its `nearest` closest listed neighbours in order of increasing distance (1..512 = mp_MXNEAREST).  As in the
 * reference the force procedures then see the truncated list until the next mdb_nlist_build. */
int mdb_nlist_reorder_nearest(mdb_ctx *ctx, int nearest);
int mdb_nlist_overflow(mdb_ctx *ctx);
int mdb_nlist_clear(mdb_ctx *ctx);

/* ------------------------------------------------------------------------------------
 * force class slots (type MDForceClassGPU, CommonGPU/MD_ForceClass_Register_GPU.F90:248-259)
 * vtensor (3,3) column-major, already divided by the number of boxes (COPYOUT_VIRIALTENSOR :1462);
 * may be NULL unless MDB_VIRIAL is set.  In a slab-decomposed run (mdb_dd_*) it is the partial sum over the pairs whose
 * first atom this rank owns: the sum over ranks is the tensor of the box.
 * ---------------------------------------------------------------------------------- */
int mdb_force(mdb_ctx *ctx, unsigned flags, double vtensor[9]);

/* ------------------------------------------------------------------------------------
 * integrator: Predictor_DEV / Correction_DEV / CalEKin_DEV (CommonGPU/MD_DiffScheme_GPU.F90:604,821,1000)
 * ---------------------------------------------------------------------------------- */
int mdb_predict(mdb_ctx *ctx, double h);
int mdb_correct(mdb_ctx *ctx, double h);
int mdb_ekin(mdb_ctx *ctx);

/* ------------------------------------------------------------------------------------
 * electron-phonon coupling: Do_ResetParam_DEV / Do_EPCForce_DEV
 * (LocalTempCtrlMeths/MD_LocalTempMethod_GPU.F90:95-138 -> EPC/MD_EP_Coupling_GPU.F90:370-417,421-493)
 * per-group arrays of length ngroup: enable flag, electron temperature TI [K], EPC_ALPHA [s],
 * EPC_CUT, EPC_HE [erg] (Common/MD_TypeDef_EPCCtrl.F90:27-37).
 * ---------------------------------------------------------------------------------- */
int mdb_epc_set(mdb_ctx *ctx, const int *enable, const double *te, const double *alpha, const double *cut,
                const double *he);
int mdb_epc_apply(mdb_ctx *ctx);
/* Do_EPCForce_DEV followed by Correction_DEV as ONE kernel (identical arithmetic; what mdb_run does inside a block) */
int mdb_epc_correct(mdb_ctx *ctx, double h);

/* ------------------------------------------------------------------------------------
 * cascade physics around the path (BASELINE configs[4]; csrc/mdb_cascade.cu)
 *   mdb_active_region  ActivateRegion_DEV -> ActiveByCells1 (CommonGPU/MD_ActiveRegion_GPU.F90:1165-1353): seeds by atom type
 *                      (method bit 0, centpart[ngroup] > 0) and / or kinetic energy >= ekin_erg (bit 1); bit 2 = CP_KEEP_AR
 *                      (earlier activations are kept); the seeds' cells grown `extend` times over the 27-cell
 *                      neighbourhood; atoms of marked cells become active.  bit 3 = CP_BYNB_AR: ActiveByNeigbors1 (:887-997)
 *                      instead -- a seed marks the atoms of its neighbour list, `extend` times, marked atoms become active.
 *                      Needs the cell ids / the list of a built list.  Inactive
 *                      atoms get no force and do not move; cells without active atoms are skipped by the next list build
 *                      (NAAC, MD_NeighborsList_GPU.F90:981-982).  Returns the number of active atoms.
 *   mdb_active_all     Active_All_ActiveRegion_DEV (on != 0) / DeActive_All_ActiveRegion_DEV (on == 0), :242-425
 *   mdb_stopping_set   Initialize_STMOD_DEV + Reset_STMOD_DEV (LocalTempCtrlMeths/Stopping/MD_ST_Coupling_GPU.F90:334-427):
 *                      stopping tables ETAB(NE) [erg], STAB(NE,NK) [erg cm^2] as the stopping libraries (Stop_Srim, Stop_Z85,
 *                      Stop_Z95, Stop_B) fill them, KPAIR(NG,NG) column-major 1-based (moving type, medium type), the
 *                      per-type switch (LT_CTRL%METH and CP_TICTRL_METH_ST) and the number densities MDEN [1/cm^3] of the
 *                      medium types (global-density model, ST_MOD_GDEN_KERNEL :431-536).  ne < 2 switches it off.
 *   mdb_stopping_options  the two switches Reset_STMOD_DEV derives from the control file (:361-427): local_density != 0 = the
 *                      local-density model (ST_CTRL%MDEN < 0 -> mp_STMOD_L, ST_MOD_LDEN_KERNEL :604-717: the density of medium type
 *                      g around an atom is its number of LIST neighbours of that type over LVOL = 4 pi/3 NB_RM(i,g)^3);
 *                      save_eloss != 0 = per-atom inelastic energy loss FF |V| DT accumulated over the steps
 *                      (ST_CTRL%SaveEloss, ST_MOD_ELOSS_*_KERNEL :835-1132, the "Eloss (ev)" data pad of Do_STMOD_DEV :1248-1262)
 *   mdb_stopping_eloss the accumulators [erg], ORIGINAL order; reset != 0 clears them
 *   mdb_stopping_apply Do_STMOD_Force_DEV / Do_STMOD_Force_Eloss_DEV (:787-831, :1206-1262): FP -= sum_g n_g S_kg(E) v/|v|; dt =
 *                      CtrlParam%H (used by the energy-loss bookkeeping only).  mdb_run applies it every step between the EPC
 *                      friction and the corrector once tables are set.
 *   mdb_pka_insert     a primary knock-on atom (Deposition/MD_TypeDef_Projectile.F90, CP_DEP_STYPE_PKA, mono-energetic): the
 *                      velocity of the atom with ORIGINAL id orig_id becomes sqrt(2 EK / m) along dir
 * ---------------------------------------------------------------------------------- */
int mdb_active_region(mdb_ctx *ctx, int method, const int *centpart, double ekin_erg, int extend);
int mdb_active_all(mdb_ctx *ctx, int on);
int mdb_stopping_set(mdb_ctx *ctx, int ne, int nk, const double *etab, const double *stab, const int *kpair, const int *enable,
                     const double *mden);
int mdb_stopping_options(mdb_ctx *ctx, int local_density, int save_eloss);
int mdb_stopping_eloss(mdb_ctx *ctx, double *eloss_host, int reset);
int mdb_stopping_apply(mdb_ctx *ctx, double dt);
int mdb_pka_insert(mdb_ctx *ctx, int orig_id, double ekin_erg, const double dir[3]);

/* ------------------------------------------------------------------------------------
 * PARREP event detection on the device: Do_ChangeDetect (Appshell/MD_Method_ParRep_GPU.F90:1094-1167) =
 *   mdb_state_save    Copy_SimMDBox(SimBox(IB), SwapBox(IB)) for all replicas, kept in device memory (ORIGINAL order)
 *   (quench)          Do_Damp -> mdb_steepest / mdb_cg / mdb_lbfgs / mdb_dyndamp
 *   mdb_compare       Do_Compare (:1241-1297): replicas against SimBoxIni%XP(NPRT,3) (host, column-major), optional MASK(NPRT),
 *                     drtol = STRCUT_DRTol in cm; flag_box[nbox] (host), optional per-atom Flag(nbox*NPRT); IBT / NCB as :1146-1156
 *   mdb_state_restore CopyIn_SimBox_DEV + Cal_NeighBoreList_DEV (:1158-1159): the saved replicas and their list come back
 * ---------------------------------------------------------------------------------- */
int mdb_state_save(mdb_ctx *ctx);
int mdb_state_restore(mdb_ctx *ctx);
int mdb_compare(mdb_ctx *ctx, const double *xp_ini, const int *mask, double drtol, int *flag_box, int *flag_atom, int *ibt, int *ncb);

/* ------------------------------------------------------------------------------------
 * one whole MD step, For_One_Step (Appshell/MD_Method_GenericMD_GPU.F90:496-659):
 * predictor -> [rebuild if MOD(itime-it0,nb_uptab)==0] -> force -> EPC -> corrector,
 * with the element-wise stages fused into the force passes where no rebuild intervenes.
 * mdb_run repeats it nsteps times without returning to the host (itime = itime0 .. itime0+nsteps-1).
 * Both return the accumulated out-of-box count (>=0) or <0.
 * ---------------------------------------------------------------------------------- */
int mdb_step(mdb_ctx *ctx, int itime, int it0, int nb_uptab, double h);
int mdb_run(mdb_ctx *ctx, int itime0, int nsteps, int it0, int nb_uptab, double h);
/* mdb_run without the final synchronisation: returns MDB_OK once the block is enqueued (the host only waits inside a
 * list rebuild, for its capacity check); mdb_sync returns the block's out-of-box count */
int mdb_run_async(mdb_ctx *ctx, int itime0, int nsteps, int it0, int nb_uptab, double h);

/* The time loop of the GMD method with its schedules, nsteps x { step size, list period, For_One_Step, TIME += H }:
 *   scheme I  (IHDUP > 0)  H = min(HMX, HMI*(int((ITIME-IT0+1)/IHDUP)+1))            Appshell/MD_Method_GenericMD_GPU.F90:353-356
 *   scheme II (IHDUP < 0)  every |IHDUP| steps (MOD(ITIME-IT0+1,|IHDUP|) == 0) Predictor_DEV starts from TH = HMX and halves it
 *                          until CheckTimestep_DEV finds no active atom that would move more than DMX in the predictor step;
 *                          that TH becomes CtrlParam%H                                 CommonGPU/MD_DiffScheme_GPU.F90:633-655
 *   IHDUP = 0              fixed H (the value passed in *h)
 *   list period            NB_UPTAB = min(NB_UPTABMX, NB_UPTABMI*(int((ITIME-IT0+1)/NB_DBITAB)+1))                  :358-360
 * The halving loop runs as ONE kernel that tests all trial steps HMX 2^-k at once and one 4-byte read-back per check.  Electronic
 * stopping, when switched on (mdb_stopping_set), acts between the EPC friction and the corrector of every step.  With stopping or
 * scheme II the tiled passes decide the distance-class shortcut PER TILE (MDB_OPT_TILE_GUARD): one fast atom sends only the tiles
 * around it to the full list.  hmi / hmx in s, dmx in cm (the control file gives fs and Angstrom, Common/MD_Gvar.F90:944-947).
 * In/out: *h = CtrlParam%H, *time_s (may be NULL) += the sum of the steps taken [s].  Returns like mdb_run. */
typedef struct mdb_sched {
    int ihdup;
    double hmi, hmx, dmx;
    int nb_uptabmi, nb_uptabmx, nb_dbitab;
} mdb_sched;
int mdb_run_sched(mdb_ctx *ctx, int itime0, int nsteps, int it0, const mdb_sched *sched, double *h, double *time_s);

/* ------------------------------------------------------------------------------------
 * single huge box over several GPUs: slab decomposition along z (whole z-layers of cells per rank).
 * Upgrades the reference's scheme -- contiguous cell ranges per device with REPLICATED positions and
 * host-staged copies (CommonGPU/MD_NeighborsList_GPU.F90:1576-1605, MD_Globle_Variables_GPU.F90:2026-2040,
 * Synchroniz_DEN_on_Devices MD_EAM_ForceTable_GPU.F90:617-642) -- to owned + ghost layers: every rank keeps
 * full-size arrays in the common cell-sorted order but only its owned layers and one ghost layer on each
 * side are kept current.  Per step the caller exchanges the boundary layers' packed records
 * (mdb_devptr(MDB_F_POS4), contiguous ranges from mdb_dd_info) after mdb_predict and again (DEN) between
 * mdb_force(MDB_DEN) and mdb_force(MDB_FORCE|MDB_NOPASS1); before a rebuild every rank's owned ranges are
 * gathered so that all ranks perform the same deterministic cell sort.  msmpscu_b200/domain.py does this
 * over torch.distributed (NCCL send/recv).
 * info: a0,a1 owned atoms | gb0,gb1 ghost below | ga0,ga1 ghost above | sb0,sb1 my bottom layer |
 *       st0,st1 my top layer | rank below, rank above | cell_lo,cell_hi | tile_lo,tile_hi   (0-based, CELL order)
 * ---------------------------------------------------------------------------------- */
int mdb_dd_set(mdb_ctx *ctx, int rank, int nranks);
int mdb_dd_info(const mdb_ctx *ctx, int info[16]);
/* The decomposed run driven from inside the library (mdb_dd.cu): one process per GPU, exchanges as ncclSend / ncclRecv
 * enqueued on the context's stream.  Replaces Synchroniz_XP_on_Devices (MD_Globle_Variables_GPU.F90:2026-2040) and
 * Synchroniz_DEN_on_Devices (MD_EAM_ForceTable_GPU.F90:617-642), and the all-atoms host sort of every device
 * (MD_NeighborsList_GPU.F90:1421-1695): a rank re-sorts only its slab.
 *   mdb_dd_nccl_id    rank 0 obtains the 128-byte NCCL unique id; the caller distributes it (MPI, torch.distributed, a file ...)
 *   mdb_dd_nccl_init  after mdb_dd_set: creates the communicator of the nranks processes
 *   mdb_dd_local_attach  backend for single-GPU tests: the contexts of ALL ranks in one process on one device; a call of
 *                     the entry points below on any of them then drives every rank in lock step
 *   mdb_dd_build      collective.  First call: every rank holds the whole initial state (uploaded as usual), sorts it and
 *                     builds the lists of its own tiles.  Later calls: the local rebuild.
 *   mdb_dd_force      collective pCalForce / pCalPTensor (flags as mdb_force; the virial is summed over the ranks)
 *   mdb_dd_run        collective nsteps x For_One_Step (rebuild cadence as mdb_run)
 *   mdb_dd_global_t   Cal_GlobalT_DEV over the owned atoms of all ranks
 * State is read back per rank with mdb_state_download(..., MDB_ORDER_CELL): the slice [info[0], info[1]) is this rank's. */
int mdb_dd_nccl_id(void *id128);
int mdb_dd_nccl_init(mdb_ctx *ctx, const void *id128);
int mdb_dd_local_attach(mdb_ctx **ctxs, int nranks);
int mdb_dd_build(mdb_ctx *ctx);
int mdb_dd_force(mdb_ctx *ctx, unsigned flags, double vtensor[9]);
int mdb_dd_run(mdb_ctx *ctx, int itime0, int nsteps, int it0, int nb_uptab, double h);
int mdb_dd_global_t(mdb_ctx *ctx, double *curt);
/* mdb_run_sched on the decomposed box: the time-step masks of the ranks are OR-ed (4 bytes per rank all-gathered), electronic
 * stopping acts on the owned atoms, EPC friction and the corrector close every step. */
int mdb_dd_run_sched(mdb_ctx *ctx, int itime0, int nsteps, int it0, const mdb_sched *sched, double *h, double *time_s);

/* ------------------------------------------------------------------------------------
 * The other integrator-module procedures the step loops call (CommonGPU/MD_DiffScheme_GPU.F90):
 *   mdb_global_t        Cal_GlobalT_DEV(SimBox, CtrlParam, CURT) :1042-1064 : EKIN kernel, then
 *                       CURT = 2*sum(EKIN >= 0)/count(EKIN >= 0)/(3 k_B) (inactive and FIXPOS atoms carry -1e32)
 *   mdb_vel_scaling     VelScaling_DEV(SimBox, CtrlParam, DT) :1262-1446 : per box, velocities times
 *                       sqrt(DT*3*k_B/2 / <EKIN>_box); fixed components are zeroed; MDB_ERR_STATE where the
 *                       reference stops (a box without kinetic energy)
 *   mdb_check_timestep  CheckTimestep_DEV(ITIME, SimBox, CtrlParam, TH, H2S2, DMX2, IFLAG) :1066-1258 : IFLAG = 1 if any
 *                       active atom would move more than sqrt(DMX2) in the next predictor step
 * ---------------------------------------------------------------------------------- */
int mdb_global_t(mdb_ctx *ctx, double *curt);
/* the same per box of a MULTIBOX context (t_box[nbox]): the per-box scalars a multi-box run gathers at output intervals */
int mdb_box_temperatures(mdb_ctx *ctx, double *t_box);
int mdb_vel_scaling(mdb_ctx *ctx, double dt);
int mdb_check_timestep(mdb_ctx *ctx, double th, double h2s2, double dmx2, int *iflag);
/* The halving loop Predictor_DEV runs around CheckTimestep_DEV when IHDUP < 0 (:633-655), as one call: *h = the first
 * TH = HMX, HMX/2, HMX/4, ... for which no active atom would move further than DMX [cm] in the predictor step (one kernel tests
 * all trial steps, one 4-byte read-back).  Collective in slab-decomposed runs (the ranks' findings are combined). */
int mdb_timestep_limit(mdb_ctx *ctx, double hmx, double dmx, double *h);

/* ------------------------------------------------------------------------------------
 * Quench.  Do_Steepest_Forsteps_DEV(SimBox, CtrlParam, ForceClass, MXNUMSTEPS, METH),
 * CommonGPU/MD_SteepestScheme_GPU.F90:263-290 -> Do_Steepest0_Forsteps_DEV :20-153: steepest descent with a
 * Barzilai-Borwein step on the current neighbour list (no rebuild inside, as in the reference), called by
 * For_One_Step's QUICKDAMP "ST" branch (Appshell/MD_Method_GenericMD_GPU.F90:540-546) and by the PARREP event
 * quench.  alpha = STEEPEST_Alpha; maxdis / mindis = STEEPEST_MxStep / STEEPEST_MiStep * RR [cm];
 * minepot = STEEPEST_MiDelE * CP_EV2ERG [erg].  meth = CtrlParam%DAMPSCHEME: with MDB_QUENCH_LSEARCH set the
 * reference runs the line-search variant Do_Steepest1_Forsteps_DEV (:157-260): normalised force direction, repeated
 * secant steps until |STEPSIZE| <= mindis, energy criterion fixed at 0.001 eV (alpha and minepot unused, as there).
 * Outputs (may be NULL): iflag = iteration at which a criterion was met (0: ran out of steps, -1: converged at
 * the first step), maxmove [cm], delepot [erg] as the reference prints them.  The scalars and the stop flag live on
 * the device; the host synchronises once per 8 iterations instead of five times per iteration.
 * ---------------------------------------------------------------------------------- */
#define MDB_QUENCH_LSEARCH 65536 /* CP_DAMPSCHEME_LSEARCH, Common/MD_Const.F90:27 */
int mdb_steepest(mdb_ctx *ctx, int mxnumsteps, int meth, double alpha, double maxdis, double mindis, double minepot,
                 int *iflag, double *maxmove, double *delepot);

/* pCalAVStress(IDEV, dAVP) -> Cal_EAM_AtomicStressTensor_DEV, CommonGPU/MD_EAM_ForceTable_GPU.F90:1973-1990 (kernel
 * :1775-1925; FS twin in MD_FS_ForceTable_GPU.F90): per-atom virial tensor AP(N,9) = sum_j DXYZ_a*DXYZ_b*FORTOT in the
 * order 11,12,13,21,..,33, column-major, CELL order, written to the caller's DEVICE array like the reference's dAVP.
 * The density pass is run first (the reference relies on the DEN of the last force call).  _host: the same into a
 * host array in the requested order. */
int mdb_atomic_stress(mdb_ctx *ctx, double *d_avp);
int mdb_atomic_stress_host(mdb_ctx *ctx, double *h_avp, int order);

/* ------------------------------------------------------------------------------------
 * Thermalisation.  Thermalizing_MC_DEV(SimBox, CtrlParam, TI), CommonGPU/MD_DiffScheme_GPU.F90:1746-1805
 * (kernel :1608-1672): every free velocity component of an ACTIVE atom is redrawn from the Maxwell distribution at TI
 * (V0*sqrt(-ln Z1)*cos(2 pi Z2)), inactive atoms get zero, then the mass-weighted mean velocity of each box is removed
 * from all its atoms -- on the device here, on the host in the reference (:1782-1802).
 * Random source: the reference uses per-thread cuRAND XORWOW states (Initialize_Rand_DEVICES, MSMLIB/sor/CommonGPU/
 * MSM_MultiGPU_Basic.F90:661-750), so its numbers depend on launch geometry, device count and cell order.  Here Z is a
 * pure function of (seed, draw, ORIGINAL atom id, component) through Philox4x32-10: the same call gives the same
 * velocities on any number of GPUs and in any sort order.  `draw` numbers the calls (the shim increments it).
 * mdb_thermalize_bits / mdb_philox4x32_10 expose the integers for known-answer tests (host code, no device needed).
 * ---------------------------------------------------------------------------------- */
int mdb_thermalize(mdb_ctx *ctx, double ti, unsigned long long seed, unsigned draw);
int mdb_thermalize_bits(unsigned long long seed, unsigned draw, unsigned orig_id, unsigned *out8);
int mdb_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4]);

/* Do_CG_Forsteps_DEV(SimBox, CtrlParam, ForceClass, MXNUMSTEPS, METH), CommonGPU/MD_CGScheme_GPU.F90:280-296:
 * Polak-Ribiere conjugate gradient on the current list; meth without MDB_QUENCH_LSEARCH -> Do_CG0_Forsteps_DEV (:16-133,
 * one secant step per direction; scalars and stop flag device-resident, host looks once per 4 iterations), with it ->
 * Do_CG1_Forsteps_DEV (:137-276, repeated secant steps until |STEPSIZE| <= mindis; one 128-byte readback per force
 * evaluation).  iflag: the reference's ITER when DELEPOT <= minepot or F0NORM <= 1e-64 fired, 0 when the budget ran
 * out, -1 when F0NORM <= 1e-64 before the first step. */
/* DO_LBFGSB_FORSTEPS_DEV(SimBox, CtrlParam, ForceClass, MXNUMSTEPS, IFLAG), CommonGPU/MD_LBFGSScheme_GPU.F90:177-388: the
 * limited-memory BFGS quench every method class uses (Do_Damp of PARREP / ART / BST / TAD, QUICKDAMP "LBFGS" of GMD).  The
 * reference packs X and G on the host and calls the serial SETULB (L-BFGS-B, NBD = 0: no bounds) each iteration; here the
 * vectors stay on the device and only scalars reach the host (csrc/mdb_lbfgs.cu).  msave = LBFGS_MSave (<= 16), factr =
 * LBFGS_Factr, pgtol = LBFGS_PGtol (max |gradient component|, erg/cm).  mxnumsteps counts SETULB calls as the reference does
 * (one per force evaluation, one per accepted step).  iflag: 0 finished, 1 out of steps.  Velocities are zeroed at the end. */
int mdb_lbfgs(mdb_ctx *ctx, int mxnumsteps, int msave, double factr, double pgtol, int *iflag, int *nfg, int *niter);
/* DAMPING_KERNEL (CommonGPU/MD_DiffScheme_GPU.F90:125-185): what Predictor_DEV runs in front of the predictor while
 * DAMPTIME0 <= ITIME < DAMPTIME0 + DAMPTIME1 (:611-617).  mdb_dyndamp = Do_DynDamp_Forsteps_DEV (:1809-1860), the
 * CP_DAMPSCHEME_DYN branch of Do_Damp: damped dynamics with step h until max|EPOT - EPOT0| <= minepot [erg]; iflag as for
 * mdb_steepest (0: ran out of steps). */
int mdb_damping(mdb_ctx *ctx);
int mdb_dyndamp(mdb_ctx *ctx, int mxnumsteps, double h, double minepot, int *iflag, double *delepot);
int mdb_cg(mdb_ctx *ctx, int mxnumsteps, int meth, double maxdis, double mindis, double minepot, int *iflag, double *delepot);

/* ------------------------------------------------------------------------------------
 * options.  MDB_OPT_FORCE_PATH selects the force/list implementation:
 *   AUTO    the tiled fast path when the configuration fits it, else the generic one
 *   GENERIC thread-per-atom over INDI, un-fused arithmetic in reference order (bit-faithful)
 *   TILED   shared-memory staged halo tiles (TMA producer warp + consumer warps), 16-bit slot lists
 * ---------------------------------------------------------------------------------- */
#define MDB_OPT_FORCE_PATH    0
#define MDB_OPT_TILED_LANES   1  /* lanes sharing one atom in the tiled kernels: 2, 4 or 8; 0 (default): 4, or 8 where a tile owns fewer than 96 atoms */
#define MDB_OPT_TILED_CLASSES 2  /* 1 (default): scan only the distance classes a pass needs while safe; 0: all */
#define MDB_OPT_ACTIVE_PATH   3  /* read-only: the path the last list build selected                            */
#define MDB_OPT_TILED_THREADS 4  /* threads of the persistent pass CTA (one producer warp + consumers): 512 or 768 (default) */
#define MDB_OPT_FUSE_EPILOGUE 5  /* mdb_run on the tiled path: EPC friction + corrector inside the force-pass   */
                                 /* epilogue (1) or as one separate element-wise kernel (0, default: faster)    */
#define MDB_OPT_TILED_STAGES  6  /* shared-memory pipeline stages of the pass kernel: 2 (default) or 3           */
#define MDB_OPT_TILED_BANKORDER 7 /* a kernel after the list build orders each scanned class of the stored lists so that the    */
                                 /* record gathers of a half-warp spread over the shared-memory bank groups: 1 on, 0 off,     */
                                 /* -1 (default) on for boxes of >= 12 cells per edge                                         */
#define MDB_OPT_TILE_GUARD    8  /* distance-class shortcut decided per tile from per-block displacement maxima: 1 on, 0 off,  */
                                 /* -1 (default) on with electronic stopping or the displacement-limited time step            */
#define MDB_OPT_PDL           9  /* programmatic dependent launch of the predictor and the tiled passes: 1 (default) / 0            */
#define MDB_FORCE_PATH_AUTO    0
#define MDB_FORCE_PATH_GENERIC 1
#define MDB_FORCE_PATH_TILED   2
int mdb_set_option(mdb_ctx *ctx, int option, int value);
int mdb_get_option(const mdb_ctx *ctx, int option);

/* ------------------------------------------------------------------------------------
 * measurement support (not in the reference, which only has CPU_TIME stopwatches):
 * CUDA-event timing per kernel class on the context's stream.  Class ids below.
 * ---------------------------------------------------------------------------------- */
#define MDB_K_CELLSORT 0
#define MDB_K_NLIST    1
#define MDB_K_PASS1    2
#define MDB_K_PASS2    3
#define MDB_K_EPOT     4
#define MDB_K_PREDICT  5
#define MDB_K_CORRECT  6
#define MDB_K_OTHER    7
#define MDB_K_EXCHANGE 8  /* ghost-layer exchanges of a slab-decomposed run (time between enqueue and completion on the stream) */
#define MDB_K__COUNT   9
int mdb_prof_enable(mdb_ctx *ctx, int on);
int mdb_prof_reset(mdb_ctx *ctx);
/* launches[k] = kernel launches recorded for class k, ms[k] = summed device time */
int mdb_prof_get(mdb_ctx *ctx, long long launches[MDB_K__COUNT], double ms[MDB_K__COUNT]);
long long mdb_launch_count(const mdb_ctx *ctx); /* all kernel launches since creation/reset */

/* ------------------------------------------------------------------------------------
 * host-side table generators mirroring the potential libraries' Register_Interaction_Table
 * (Potentials/EAM_WW_Marinica_JPCM25_2013/EAM_ForceTable_Marinica_JPCM25_2013.F90:17-83,
 *  Potentials/EAM_WHeH_Bonny_JPCM26_2014/EAM_ForceTable_Bonny_JPCM26_2014.F90:17-140) ->
 * Create_Interaction_ForceTable (Common/MD_TypeDef_ForceTable.F90:1151-1230).
 * In a Fortran deployment the unchanged potential modules fill MDForceTable and call
 * mdb_tables_set; these exist so that C/C++/Python hosts can run the path standalone.
 * ptype(NG,NG) column-major; rmax: table range (the shipped source uses max(RU), :591).
 * Output arrays are caller-allocated: pair tables ng*ng*ntab doubles, embedding ng*nembd.
 * ---------------------------------------------------------------------------------- */
#define MDB_LIB_MARINICA_EAM2 1
#define MDB_LIB_BONNY_EAM1    2
#define MDB_LIB_ACKLAND_FS_W  3 /* FS_TYPE: Potentials/EM_TB_WangJun_W-HE_2010/FS_Ackland_WW.F90 (id 1, W-W) */
int mdb_host_ftable_create(int lib, int ng, const int *ptype, int ntab, int nembd, double rhoscal, double rmax,
                           int *nkind, int *nkind1, int *kpair, int *kembd,
                           double *potr, double *fpotr, double *potb, double *fpotb,
                           double *fembd, double *dfembd, double *csi, double *rhod);

/* ------------------------------------------------------------------------------------
 * external force tables (csrc/host_tables_io.cpp)
 *
 * mdb_host_setfl_*: the NIST "setfl" importer = Register_ForceTableProc_Setfl + Generate_NIST_ForceTalbe
 *   (Potentials/EAM_NIST/Filedatas_Func_Setfl.F90:153-296,320-462,505-541; NIST_ForceTable.F90:332-398):
 *   cubic splines with zero end curvature through the file's F(rho), rho(r), r*V(r); one table per id
 *   it = (I-1)*NE+J ("I <- J"), kind index = id; table range and RHOMX from the file (rmax <= 0) .
 *   Outputs caller-allocated: pair tables NE*NE*ntab doubles, embedding NE*NE*nembd, T(NKIND,N) column-major.
 * mdb_host_ftable_export: Export_ForceTable (Common/MD_TypeDef_ForceTable.F90:1315-1459) -> fname.pair/.embd
 * mdb_host_ftable_file_info / _import: Import_ForceTable (:1461-1591) + Register_Imported_ForceTable
 *   (:1595-1855): re-grid the tables PTYPE(ng,ng) names onto the run's grid (SPLID1 IOP=5 / SPLID2).
 * All return MDB_OK or MDB_ERR_ARG (unreadable / malformed file, id not in the file); nothing prints or stops.
 * ---------------------------------------------------------------------------------- */
int mdb_host_setfl_info(const char *path, int *nelem, int *nrho, int *nr, double *cutoff_cm, double *rhomx,
                        char *names, int names_stride, int *z, double *mass, double *alat);
int mdb_host_setfl_ftable(const char *path, int ntab, int nembd, double rmax, int *nkind,
                          double *potr, double *fpotr, double *potb, double *fpotb, double *fembd, double *dfembd,
                          double *csi, double *rhod, double *rmax_out);
/* the ".lspt" twin (Register_ForceTableProc_SPT, Potentials/EAM_NIST/Filedatas_Func_Lspt.F90:79-541): an index file naming one
 * two-column (x, f) file per function -- F(rho), rho(r) per element ("NA" = none), V(r) per pair.  Restated as written,
 * including the reader's use of the FIRST file name of a VR line for every pair of that line (:225). */
int mdb_host_lspt_info(const char *path, int *nelem, double *cutoff_cm, double *rhomx, char *names, int names_stride);
int mdb_host_lspt_ftable(const char *path, int ntab, int nembd, double rmax, int *nkind,
                         double *potr, double *fpotr, double *potb, double *fpotb, double *fembd, double *dfembd,
                         double *csi, double *rhod, double *rmax_out);
/* the ".moldy" twin (Register_ForceTableProc_Moldy, Potentials/EAM_NIST/Filedatas_Func_Moldy.F90:21-153): one element, cubic-knot
 * V and rho with the lattice constant as the unit of the knots, F = -sqrt(rho); tables generated like a built-in library. */
int mdb_host_moldy_ftable(const char *path, int ntab, int nembd, double rhoscal, double rmax,
                          double *potr, double *fpotr, double *potb, double *fpotb, double *fembd, double *dfembd,
                          double *csi, double *rhod);
int mdb_host_ftable_export(const char *fname, int pot_type, int nkind, const int *ids, int ntab, double csi,
                           const double *potr, const double *fpotr, const double *potb, const double *fpotb,
                           int nkind1, const int *ids1, int nembd, double rhod, const double *fembd, const double *dfembd);
int mdb_host_ftable_file_info(const char *fname, int *pot_type, int *nkind, int *ids, int *ntab, int *nkind1, int *ids1,
                              int *nembd, double *rmax_cm, double *rhomx);
int mdb_host_ftable_import(const char *fname, int ng, const int *ptype, int ntab, int nembd, double rmax,
                           int *pot_type, int *nkind, int *nkind1, int *kpair, int *kembd,
                           double *potr, double *fpotr, double *potb, double *fpotb, double *fembd, double *dfembd,
                           double *csi, double *rhod);

#ifdef __cplusplus
}
#endif
#endif

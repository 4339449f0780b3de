// mdb_internal.cuh -- context layout and helpers shared by the translation units of
// libmdpscu_b200.so.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/mdpscu_b200.h"

// STATU bits, Common/MD_Const.F90:79-99
#define ST_ACTIVE    1
#define ST_FIXPOSX   2
#define ST_FIXPOSY   4
#define ST_FIXPOSZ   8
#define ST_FIXPOS    14
#define ST_FIXVELX   16
#define ST_FIXVELY   32
#define ST_FIXVELZ   64
#define ST_OUTOFBOX  65536
#define ST_REFLECT   131072
#define ST_TRANSMIT  262144

#define KB_CGS 1.38054e-16 // CP_KB, MSMLIB/sor/Common/MSM_Const.F90:82

// device counters block (one int each)
// the first CNT_PERBUILD_N are reset at every rebuild
enum { CNT_OOB = 0, CNT_OVERFLOW, CNT_NNMAX, CNT_MXNAC, CNT_INCELL, CNT_TILE_OVERFLOW,
       CNT_D2MAX,       // float bits: max |displacement since the rebuild|^2 over all atoms (predictor)
       CNT_PERBUILD_N, CNT_OOB_TOTAL = CNT_PERBUILD_N, CNT_SCRATCH,
       CNT_RHO_OVER,    // density-pass evaluations whose rho lay beyond the embedding table (RHOMX): sticky, read by mdb_embed_overruns
       CNT__N = 12 };

struct BoxParams { // passed by value to kernels
    double lo[3], up[3], size[3], half[3];
    int pd[3];
};

struct TableSet { // device tables in the packed layouts the kernels read
    int pot_type, ng, nkind, ntab, nkind1, nembd;
    double csi, rhod, inv_rhod_unused, ru2max;
    // pair tables, per kind k: entry KK (0..ntab+1, Fortran index; 0 and ntab+1 are zero pads)
    //   v2[(k*(ntab+2)+KK)] = {T[KK], T[KK+1]-T[KK]}   (value, forward difference)
    double2 *potr, *fpotr, *potb, *fpotb;
    double2 *fembd, *dfembd; // same packing over nembd
    int kpair[MDB_MXGROUP * MDB_MXGROUP]; // 0-based kind for (i,j) at i+ng*j ; -1 if none
    int kembd[MDB_MXGROUP];
};

struct EpcParams {
    int on;
    int enable[MDB_MXGROUP];
    double te[MDB_MXGROUP], epa[MDB_MXGROUP], v2ti[MDB_MXGROUP], tcut[MDB_MXGROUP], eup[MDB_MXGROUP];
};

struct MassParams { double cm[MDB_MXGROUP]; };

// geometry of the tiled path, passed by value to its kernels (see mdb_tiled.cuh)
struct TileParams {
    int n, nbox, ncx, ncy, ncz, nc0;
    int ntx, nrows, ntiles;        // tiles per x-row, rows, total tiles
    int hcap;                      // halo capacity in atoms (shared-memory budget)
    int ocap;                      // owned-atom capacity of a tile
    int pd[3];
    double lo[3], size[3], cell[3];
    double half[3];                // BOXSIZE/2 on periodic axes, 1e300 otherwise (minimum-image test of the passes)
    float fbs[3];                  // (float)BOXSIZE: the fp32 image shift of the list kernel
    int ng, mxkvois, nrow4;        // nrow4 = 4-entry index groups per (atom, lane)
    size_t npad;                   // padded atom count of the slot list
};

// state of the tiled fast path (mdb_force_tiled.cu)
struct TiledState {
    TileParams P;
    bool ok = false, dirty = true, active = false;
    int G = 4;                 // lanes per atom (set by mdb_tiled_plan)
    int G_opt = 0;             // MDB_OPT_TILED_LANES: 2 / 4 / 8, 0 = the plan decides (4, or 8 for narrow tiles)
    int ntx = 0, hcap = 0, ocap = 0, wmax = 0, lpairs = 0, lcap = 0, threads = 0, threads_opt = 768, grid = 0, nbuf = 2, stages_opt = 2;
    int ktab[2] = {0, 0}, kmin[2] = {0, 0}, khi[2] = {0, 0};
    double r2eff[2] = {0.0, 0.0};
    double r2eff_epot = 0.0; bool epot_in_class1 = false; // per-atom energy pass (PASS 3 of the tiled kernel)
    size_t smem_pass[2] = {0, 0}, smem_list = 0;
    double margin = 0.0, margin0 = 0.0; // class margins (length): a class holds while every atom moved < margin/2
    float rc2f[2] = {0.f, 0.f}, safe_d2[3] = {0.f, 0.f, 0.f};
    bool use_classes = true;
    // k_tile_deal orders the scanned classes of the stored lists for conflict-free record reads: passes -3..8 %, +0.17 ms per rebuild
    // and million atoms.  -1 = auto: on for boxes of at least 12 cells per edge (most tiles interior: the passes are bound by the
    // shared-memory pipe), off for small boxes whose tiles all take the minimum-image variant (fp64-bound: measured -1 % there)
    int bank_order_opt = -1;
    bool bank_order = false;  // resolved by mdb_tiled_plan
    unsigned short *nbl = nullptr; size_t nbl_elems = 0;   // slot list (bytes in nbl_elems)
    unsigned short *ncls = nullptr; size_t ncls_bytes = 0; // per-atom class counts [2][npad]
    void *desc = nullptr; size_t desc_bytes = 0;           // TileDesc per tile
};

struct mdb_ctx {
    int dev = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;

    // ---- box
    bool has_box = false;
    int nbox = 0, napb = 0, n = 0, ng = 0;
    double boxshape[9];
    bool shape_identity = true;
    BoxParams box;
    MassParams mass;

    // ---- state, CELL order.  pos = {x,y,z,den}
    double4 *pos = nullptr, *pos_alt = nullptr;
    double *xp1 = nullptr, *xp1_alt = nullptr;
    double *fp = nullptr, *fp_alt = nullptr;
    double *dis = nullptr, *dis_alt = nullptr;
    double *epot = nullptr, *ekin = nullptr;
    int *ityp = nullptr, *ityp_alt = nullptr;
    int *statu = nullptr, *statu_alt = nullptr;
    int *gid = nullptr, *gid_alt = nullptr, *gidinv = nullptr;
    int *ic = nullptr, *ic_alt = nullptr;
    float *dsr = nullptr; size_t dsr_bytes = 0; // displacement since the last rebuild, (N,3) fp32
    // cascade runs: per-block (256 atoms of the predictor's range) maxima of |dsr|^2 and the per-tile bounds derived from them;
    // the passes read tile_d2 only between mdb_tile_guard_launch and the end of the same step (tile_guard_fresh)
    float *dmax_blk = nullptr; size_t dmax_blk_bytes = 0;
    float *tile_d2 = nullptr; size_t tile_d2_bytes = 0;
    bool tile_guard_fresh = false;
    bool var_step = false;     // inside a run with the displacement-limited time step (mdb_run_sched, IHDUP < 0)
    int opt_tile_guard = -1;   // -1 auto (on with electronic stopping or a displacement-limited time step), 0 off, 1 on
    // materialised reference-shaped views
    double *xp_view = nullptr, *den_view = nullptr;
    // staging for up/download
    void *stage = nullptr; size_t stage_bytes = 0, stage_off = 0; // bump-allocated pool, rewound by mdb_sync
    void *hstage = nullptr; size_t hstage_bytes = 0; // pinned

    // ---- cells
    bool has_nlist = false, list_valid = false;
    bool list_reordered = false; // Reorder_NeighBoreList_Nearest_Dev truncated KVOIS/INDI in place (until the next rebuild)
    double nb_rm[MDB_MXGROUP * MDB_MXGROUP];
    float rm2f[MDB_MXGROUP * MDB_MXGROUP];
    int ncell[3] = {0, 0, 0}, nc0 = 0, nc = 0, mxnac = 0;
    int *nac = nullptr, *naac = nullptr, *ia1th = nullptr;
    int *slot = nullptr, *srcof = nullptr, *tmp_orig = nullptr, *oob = nullptr;
    int *scan_tmp = nullptr; int scan_n = 0; // block totals / maxima of the cell prefix scan
    void *scratch = nullptr; size_t scratch_bytes = 0; // grow-only work space of the occasional calls (mdb_scratch)
    int *counters = nullptr; // CNT__N ints on device
    int *h_counters = nullptr; // pinned mirror
    int mxkvois = 0;
    int *kvois = nullptr, *indi = nullptr;
    // tiled path: INDI (and the reference-ordered KVOIS) are produced on demand from the positions of the last rebuild
    bool indi_stale = false;
    double4 *pos_snap = nullptr; size_t pos_snap_bytes = 0;
    int oob_total = 0;
    long long list_gen = 0;    // counts the list rebuilds (caches derived from the list compare it)
    bool run_pending = false;  // mdb_run_async enqueued a block whose counters mdb_sync still has to read
    int fallbacks = 0;         // rebuilds that overflowed the tiled path and were redone on the generic one

    // ---- tables
    bool has_tables = false;
    TableSet tab;
    std::vector<void *> tab_allocs;

    // ---- epc
    EpcParams epc;
    void *save_state = nullptr; // device copy of the replicas taken by mdb_state_save (PARREP event detection)
    void *stop_state = nullptr; // electronic stopping tables and switches (mdb_cascade.cu)

    // ---- host copies of the pair tables (Fortran layout) for planning the tiled path
    std::vector<double> h_potb, h_fpotr, h_fpotb, h_potr;

    // ---- tiled fast path
    TiledState tiled;

    // ---- slab domain decomposition (single huge box over several ranks, z-layers of cells)
    bool dd_on = false;
    int dd_rank = 0, dd_n = 1;
    int dd_info[16] = {0}; // a0,a1, gb0,gb1, ga0,ga1, sb0,sb1, st0,st1, below,above, cell_lo,cell_hi, tile_lo,tile_hi
    int *h_dd = nullptr;   // pinned scratch
    // backend of the library-level decomposed run (mdb_dd.cu): NCCL communicator (one process per GPU) or, for single-GPU
    // tests, the contexts of all ranks in this process (same device, same stream: exchanges are device-to-device copies)
    void *dd_comm = nullptr;
    int tile_sel[4] = {-1, -1, 0, 0}; // tile ranges of the next tiled pass launches when >= 0 (mdb_dd.cu: interior / boundary split)
    cudaStream_t dd_xs = nullptr;     // side stream of the overlapped ghost exchange
    cudaEvent_t dd_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    void *dd_p2p = nullptr;     // peer-to-peer ghost exchange state (CUDA IPC mappings of the neighbours' arrays)
    std::vector<mdb_ctx *> dd_peers;
    int *dd_dev = nullptr;      // device: per rank {owned atoms, bottom-layer atoms, top-layer atoms, max atoms per cell}, then scratch
    bool dd_built = false;      // the initial replicated build is done: later rebuilds sort owned + ghost atoms only
    std::vector<int> dd_tab;    // host copy of the per-rank table of the last rebuild

    // ---- options
    int opt_force_path = MDB_FORCE_PATH_AUTO;
    int opt_pdl = 1;           // programmatic dependent launch of the predictor and the tiled passes (prologues under the previous kernel's tail)
    int opt_fuse_epilogue = 0; // measured slower than the separate 27 us kernel on B200 (profiles/r01_summary.md)

    // ---- quench (mdb_quench.cu): work arrays, pinned scalar mirror, and the device flag that turns the force
    //      kernels of already-converged iterations into no-ops
    double *q_buf = nullptr; int q_n = 0; void *q_host = nullptr;
    const int *skip_flag = nullptr;
    double *lb_buf = nullptr; size_t lb_doubles = 0; unsigned char *lb_mask = nullptr; size_t lb_mask_n = 0; void *lb_host = nullptr; // L-BFGS workspace
    double *avp = nullptr; int avp_n = 0;                      // atomic stress scratch (mdb_atomic_stress_host)

    // ---- virial partials
    double *vpart = nullptr; int vpart_n = 0;

    // ---- profiling
    bool prof = false;
    long long launches_total = 0;
    long long prof_launches[MDB_K__COUNT];
    double prof_ms[MDB_K__COUNT];
    struct Ev { cudaEvent_t a, b; int k; };
    std::vector<Ev> ev_pending;
    std::vector<cudaEvent_t> ev_pool;
};

// ---- error helpers
int mdb_fail(mdb_ctx *c, int code, const char *fmt, ...);
#define CUDA_TRY(c, call)                                                                          \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return mdb_fail((c), MDB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                                   \
    } while (0)

// ---- profiling bracket: ProfScope p(ctx, MDB_K_PASS1); launch...; (destructor records end)
struct ProfScope {
    mdb_ctx *c; int k; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(mdb_ctx *ctx, int klass, int nlaunch = 1);
    ~ProfScope();
};
void mdb_prof_collect(mdb_ctx *c);

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- internal entry points across translation units
int mdb_cells_build(mdb_ctx *c);             // mdb_cells.cu : bin, sort, permute
int mdb_nlist_kernel(mdb_ctx *c, const double4 *pos = nullptr); // mdb_nlist.cu : fill KVOIS/INDI (pos: override positions)
int mdb_nlist_nearest(mdb_ctx *c, int nearest); // mdb_nlist.cu
int mdb_indi_ensure(mdb_ctx *c);             // mdb_api.cu : materialise INDI after a tiled rebuild
int mdb_force_generic(mdb_ctx *c, unsigned flags, double *vt); // mdb_force.cu
int mdb_avstress_generic(mdb_ctx *c, double *d_ap);            // mdb_force.cu
int mdb_virial_finish(mdb_ctx *c, int nblk, double *vt);       // mdb_force.cu
int mdb_tiled_plan(mdb_ctx *c);               // mdb_force_tiled.cu
int mdb_tiled_nlist(mdb_ctx *c);
void mdb_tiled_free(mdb_ctx *c);
void mdb_mark_positions_dirty(mdb_ctx *c);    // mdb_api.cu : positions changed outside the predictor
int mdb_force_tiled(mdb_ctx *c, unsigned flags, int fuse = 0, double hs2 = 0.0);
int mdb_list_rebuild(mdb_ctx *c);
int mdb_list_rebuild_checked(mdb_ctx *c);    // mdb_api.cu : + capacity check and generic fallback (syncs)
int mdb_dd_update(mdb_ctx *c);                // mdb_api.cu : owned / ghost ranges after a rebuild (syncs)
int mdb_cells_dd_count(mdb_ctx *c, const int cand[6], int zl0, int zl1, int *d_out4);      // mdb_cells.cu
int mdb_cells_dd_place(mdb_ctx *c, const int cand[6], int zl0, int zl1, int base, int nown);
int mdb_cells_dd_ghost_layer(mdb_ctx *c, int c0, int first);
int mdb_predict_launch(mdb_ctx *c, double h, int pre);    // mdb_step.cu
int mdb_epc_correct_launch(mdb_ctx *c, double h);
int mdb_step_close_launch(mdb_ctx *c, double h);          // EPC friction [, stopping], corrector on the owned range
int mdb_timestep_mask_launch(mdb_ctx *c, double hmx, double dmx2);                 // mdb_step.cu : variable time step (scheme II)
int mdb_timestep_from_mask(mdb_ctx *c, unsigned mask, double hmx, double *h);
int mdb_dd_timestep(mdb_ctx *c, double hmx, double dmx, double *h);                // mdb_dd.cu : the same over all ranks
int mdb_sched_nb_uptab(const mdb_sched *s, int itime, int it0);
bool mdb_sched_check_due(const mdb_sched *s, int itime, int it0);
double mdb_sched_h1(const mdb_sched *s, int itime, int it0, double h);
bool mdb_tile_guard_wanted(const mdb_ctx *c);
int mdb_tile_guard_launch(mdb_ctx *c, int lo, int hi);    // mdb_force_tiled.cu : per-tile displacement bounds for this step's passes
void mdb_dd_free(mdb_ctx *c);                 // mdb_dd.cu
int mdb_stopping_launch(mdb_ctx *c, double dt); // mdb_cascade.cu : electronic stopping on FP (no-op when switched off)
int mdb_stopping_prepare(mdb_ctx *c);          // per-type neighbour counts of the local-density model, after a rebuild
bool mdb_stopping_on(const mdb_ctx *c);
void mdb_stopping_free(mdb_ctx *c);
void mdb_save_free(mdb_ctx *c);
void *mdb_scratch(mdb_ctx *c, size_t bytes);  // mdb_api.cu : grow-only device work space of the context (valid until the next call)
static inline int own_a0(const mdb_ctx *c) { return c->dd_on ? c->dd_info[0] : 0; }
static inline int own_a1(const mdb_ctx *c) { return c->dd_on ? c->dd_info[1] : c->n; }             // mdb_api.cu : cells + list kernel of the active path (no sync)

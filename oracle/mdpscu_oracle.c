/*
 * mdpscu_oracle.c -- CPU restatement of the MDPSCU tabulated EAM/FS hot path.
 * TEST INFRASTRUCTURE ONLY (see mdpscu_oracle.h).  Compile with -ffp-contract=off:
 * the restatement is defined as un-fused arithmetic in the reference's source order.
 *
 * Reference paths are relative to the reference root (MDLIB/sor/... is abbreviated
 * to its last two components where unambiguous).
 */
#include "mdpscu_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* =====================================================================================
 * Potentials
 * ===================================================================================== */

/* Cubic-knot sum  sum_k a_k (r_k - r)^3 H(r_k - r)  and  sum_k a_k (r_k - r)^2 H(r_k - r).
 * Common/MD_Pot_EAM_Utilities.F90:35-44 (H: x>=0 -> 1), :72-79 / :280-287 (the loop,
 * products in source order a*(d)*(d)*(d)). r and knots in Angstrom. */
static void knot_sum(double r, const double *a, const double *rk, int n, double *s3, double *s2)
{
    double p = 0.0, f = 0.0;
    for (int i = 0; i < n; i++) {
        double d = rk[i] - r;
        double stp = (d >= 0.0) ? 1.0 : 0.0;
        p = p + a[i] * d * d * d * stp;
        f = f + a[i] * d * d * stp;
    }
    *s3 = p;
    *s2 = f;
}

/* NN_FuncPoly3, Common/MD_Pot_EAM_Utilities.F90:48-85: returns 0.5*V [erg], -dV/dr [erg/cm] */
static void nn_poly3(double r_ang, const double *a, const double *rk, int n, double *potr, double *fpotr)
{
    double p, f;
    knot_sum(r_ang, a, rk, n, &p, &f);
    *potr = 0.5 * p * ORC_EVERG;
    *fpotr = 3.0 * f * ORC_EVERG / ORC_A2CM;
}

/* RHO_FuncPoly3, :259-290: rho (dimensionless), -drho/dr [1/cm] */
static void rho_poly3(double r_ang, const double *a, const double *rk, int n, double *potb, double *fpotb)
{
    double p, f;
    knot_sum(r_ang, a, rk, n, &p, &f);
    *potb = p;
    *fpotb = 3.0 * f / ORC_A2CM;
}

/* EMBED_FS_PLOY_Func, :332-368: F = A1 sqrt(rho) + A2 rho^2 + ... [erg] */
static void embed_fs_poly(double rho, const double *a, int n, double *frho, double *dfrho)
{
    double fr, dfr;
    if (rho <= 0.0) {
        fr = 0.0;
        dfr = 0.0;
    } else {
        double trho = rho;
        fr = a[0] * sqrt(trho);
        dfr = 0.5 * a[0] / sqrt(trho);
        for (int i = 2; i <= n; i++) {
            fr = fr + a[i - 1] * rho * trho;
            dfr = dfr + (double)i * a[i - 1] * trho;
            trho = rho * trho;
        }
    }
    *frho = fr * ORC_EVERG;
    *dfrho = dfr * ORC_EVERG;
}

/* ---- Marinica EAM2 W-W. Potentials/EAM_WW_Marinica_JPCM25_2013/EAM2_WW_Marinica_JPCM25_2013.F90:17-116.
 * Coefficients are D-exponent (double) literals; KNOTS are default-REAL literals in the
 * Fortran source (:42-56, :80-83) i.e. rounded to float32 then widened -- hence the f suffixes. */
static const double MAR2_A[15] = {
    0.960851701343041e2, -0.184410923895214e3, 0.935784079613550e2, -0.798358265041677e1,
    0.747034092936229e1, -0.152756043708453e1, 0.125205932634393e1, 0.163082162159425e1,
    -0.141854775352260e1, -0.819936046256149e0, 0.198013514305908e1, -0.696430179520267e0,
    0.304546909722160e-1, -0.163131143161660e1, 0.138409896486177e1};
static const double MAR2_R[15] = {
    2.564897500000000f, 2.629795000000000f, 2.694692500000000f, 2.866317500000000f,
    2.973045000000000f, 3.079772500000000f, 3.516472500000000f, 3.846445000000000f,
    4.176417500000000f, 4.700845000000000f, 4.895300000000000f, 5.089755000000000f,
    5.342952500000000f, 5.401695000000000f, 5.460437500000000f};
static const double MAR2_B[4] = {-0.420429107805055e1, 0.518217702261442e0, 0.562720834534370e-1,
                                 0.344164178842340e-1};
static const double MAR2_BR[4] = {2.500000000000000f, 3.100000000000000f, 3.500000000000000f,
                                  4.900000000000000f};
static const double MAR2_AF[2] = {-5.946454472402710e0, -0.049477376935239e0};

static void mar2_nn(double r, double *potr, double *fpotr)
{
    nn_poly3(r * ORC_CM2A, MAR2_A, MAR2_R, 15, potr, fpotr);
}
static void mar2_rho(double r, double *potb, double *fpotb)
{
    const double rc = 2.002970124727e0 * ORC_A2CM; /* :85 */
    if (r <= rc) {
        rho_poly3(rc * ORC_CM2A, MAR2_B, MAR2_BR, 4, potb, fpotb);
        *fpotb = 0.0;
    } else {
        rho_poly3(r * ORC_CM2A, MAR2_B, MAR2_BR, 4, potb, fpotb);
    }
}
static void mar2_embd(double rho, double *f, double *df) { embed_fs_poly(rho, MAR2_AF, 2, f, df); }

/* ---- Bonny EAM1 W-H-He. Potentials/EAM_WHeH_Bonny_JPCM26_2014/EAM1_WHeH_Bonny_JPCM26_2014.F90.
 * W-W is Marinica EAM2 under a gauge transform (:15-103). */
static void bon_ww_nn(double r, double *potr, double *fpotr)
{
    const double c = 1.848055990e0 * ORC_EVERG; /* :27 */
    double rho, frho;
    mar2_nn(r, potr, fpotr);
    mar2_rho(r, &rho, &frho);
    *potr = *potr - c * rho;           /* :35 (POTR is already 0.5 V) */
    *fpotr = *fpotr - 2.0 * c * frho;  /* :36 */
}
static void bon_ww_rho(double r, double *potb, double *fpotb)
{
    const double s = 2.232322602e-1; /* :53 */
    mar2_rho(r, potb, fpotb);
    *potb = *potb * s;
    *fpotb = *fpotb * s;
}
static void bon_ww_embd(double rho, double *frho, double *dfrho)
{
    /* :69-100 */
    const double s = 2.232322602e-01;
    const double is = 1.0e+00 / s;
    const double c = 1.848055990e+00 * ORC_EVERG;
    const double cos_ = c * is;
    const double rhoi = 1.359141225e0;
    const double a0 = -5.524855802e+00, a1 = 2.317313103e-01, a2 = -3.665345949e-02, a3 = 8.989367404e-03;
    if (rho <= rhoi) {
        double trho = rho * is, tf, tdf;
        mar2_embd(trho, &tf, &tdf);
        *frho = tf + c * trho;
        *dfrho = tdf * is + cos_;
    } else {
        double trho = rho;
        *frho = (a0 + trho * (a1 + trho * (a2 + trho * a3))) * ORC_EVERG;
        *dfrho = (a1 + trho * (2.0 * a2 + 3.0 * a3 * trho)) * ORC_EVERG;
    }
}
/* pair-only members of EAM1 (:107-330); knots here are D-exponent literals (true doubles) */
static const double BON_HEHE_R[2] = {2.0, 3.0}, BON_HEHE_A[2] = {2.106615791e+00, -2.217639348e-01};
static const double BON_HH_R[2] = {2.0, 3.0}, BON_HH_A[2] = {4.862785907e-01, 1.018797872e-01};
static const double BON_HHE_R[3] = {1.8, 2.0, 3.0}, BON_HHE_A[3] = {1.5e+01, 2.563700119e-01, -4.489510592e-02};
static const double BON_WHE_R[3] = {1.9, 2.2, 3.5}, BON_WHE_A[3] = {2.1e+01, 8.565323293e-01, 2.750099819e-01};
static const double BON_WH_R[2] = {2.0, 3.0}, BON_WH_A[2] = {1.375733214e+01, 1.296071475e-01};

/* registration order in EAM_ForceTable_Bonny_JPCM26_2014.F90:57-77 gives the numeric ids:
 * 1 W<-W, 2 W<-H, 3 W<-He, 4 H<-W, 5 H<-H, 6 H<-He, 7 He<-W, 8 He<-H, 9 He<-He */
static int bon1_nn(int id, double r, double *p, double *f)
{
    double ra = r * ORC_CM2A;
    switch (id) {
    case 1: bon_ww_nn(r, p, f); return 1;
    case 2: case 4: nn_poly3(ra, BON_WH_A, BON_WH_R, 2, p, f); return 1;
    case 3: case 7: nn_poly3(ra, BON_WHE_A, BON_WHE_R, 3, p, f); return 1;
    case 5: nn_poly3(ra, BON_HH_A, BON_HH_R, 2, p, f); return 1;
    case 6: case 8: nn_poly3(ra, BON_HHE_A, BON_HHE_R, 3, p, f); return 1;
    case 9: nn_poly3(ra, BON_HEHE_A, BON_HEHE_R, 2, p, f); return 1;
    }
    return 0;
}

/* Finnis-Sinclair W of Ackland & Thetford, Potentials/EM_TB_WangJun_W-HE_2010/FS_Ackland_WW.F90:25-90 (id 1 of the
 * library EM_TB_WANGJUN_W-HE_2010 when the box declares FS_TYPE, EM_TB_ForceTable_WangJun_W_HE_2010.F90:38-40).
 * r in cm; TB_NN: (V/2, -dV/dr) with the short-range core term below ACKB0; TB_NE: (rho = A^2 (r-d)^2, -drho/dr). */
static void ackfs_nn(double r, double *potr, double *fpotr)
{
    const double c = 3.25e-8, c0 = 47.1346499e16 * ORC_EVERG, c1 = -33.7665655e24 * ORC_EVERG, c2 = 6.2541999e32 * ORC_EVERG;
    const double ackb = 90.3e24 * ORC_EVERG, acka = 1.2e8, ackb0 = 2.7411e-8;
    double ackfs = 0.0, dackfs = 0.0;
    if (r < ackb0) {
        ackfs = ackb * pow(ackb0 - r, 3.0) * exp(-acka * r);
        dackfs = -ackb * (3.0 * pow(ackb0 - r, 2.0) * exp(-acka * r) + acka * pow(ackb0 - r, 3.0) * exp(-acka * r));
    }
    if (r <= c) {
        double v = pow(r - c, 2.0) * (c0 + c1 * r + c2 * pow(r, 2.0)) + ackfs;
        *potr = 0.5 * v;
        *fpotr = -(2.0 * (r - c) * (c0 + c1 * r + c2 * pow(r, 2.0)) + pow(r - c, 2.0) * (c1 + 2.0 * c2 * r) + dackfs);
    } else { *potr = 0.0; *fpotr = 0.0; }
}
static void ackfs_rho(double r, double *potb, double *fpotb)
{
    const double a = 1.896373e8 * ORC_EVERG, d = 4.400224e-8;
    if (r <= d) { *potb = a * a * pow(r - d, 2.0); *fpotb = -(a * a * 2.0 * (r - d)); }
    else { *potb = 0.0; *fpotb = 0.0; }
}

int orc_pot_nn(int lib, int id, double r, double *p, double *f)
{
    *p = 0.0; *f = 0.0;
    if (lib == ORC_LIB_ACKLAND_FS_W && id == 1) { ackfs_nn(r, p, f); return 1; }
    if (lib == ORC_LIB_MARINICA_EAM2 && id == 1) { mar2_nn(r, p, f); return 1; }
    if (lib == ORC_LIB_BONNY_EAM1) return bon1_nn(id, r, p, f);
    return 0;
}
int orc_pot_rho(int lib, int id, double r, double *p, double *f)
{
    *p = 0.0; *f = 0.0;
    if (lib == ORC_LIB_ACKLAND_FS_W && id == 1) { ackfs_rho(r, p, f); return 1; }
    if (lib == ORC_LIB_MARINICA_EAM2 && id == 1) { mar2_rho(r, p, f); return 1; }
    if (lib == ORC_LIB_BONNY_EAM1 && id == 1) { bon_ww_rho(r, p, f); return 1; }
    return 0;
}
int orc_pot_embd(int lib, int id, double rho, double *p, double *f)
{
    *p = 0.0; *f = 0.0;
    if (lib == ORC_LIB_MARINICA_EAM2 && id == 1) { mar2_embd(rho, p, f); return 1; }
    if (lib == ORC_LIB_BONNY_EAM1 && id == 1) { bon_ww_embd(rho, p, f); return 1; }
    return 0;
}

/* =====================================================================================
 * Table generation
 * ===================================================================================== */
int orc_ftable_build(int lib, int ng, const int *ptype, int ntab, int nembd, double rhoscal, double rmax,
                     int *nkind_out, int *nkind1_out, int *kpair, int *kembd,
                     double *potr, double *fpotr, double *potb, double *fpotb,
                     double *fembd, double *dfembd, double *csi_out, double *rhod_out)
{
    int fpair[ORC_MXGROUP * ORC_MXGROUP], fpair1[ORC_MXGROUP];
    int nkind = 0, nkind1 = 0;
    if (ng > ORC_MXGROUP) return -1;
    /* New_ForceTable, MD_TypeDef_ForceTable.F90:559-574: unique ids, I outer, J inner */
    for (int i = 0; i < ng; i++)
        for (int j = 0; j < ng; j++) {
            int id = ptype[i + ng * j], k;
            for (k = 0; k < nkind; k++)
                if (fpair[k] == id) break;
            if (k == nkind) fpair[nkind++] = id;
        }
    /* :600-611 */
    for (int i = 0; i < ng; i++) {
        int id = ptype[i + ng * i], k;
        for (k = 0; k < nkind1; k++)
            if (fpair1[k] == id) break;
        if (k == nkind1) fpair1[nkind1++] = id;
    }
    /* :591-594 */
    double rmaxsqrt = sqrt(rmax);
    double csi = (double)ntab / rmaxsqrt;
    double csiv = 1.0 / csi;
    double rhomx = 0.0, rhod = 0.0;

    /* Create_Interaction_ForceTable :1192-1205 -> Create_Pairwise_ForceTable :949-976 */
    for (int k = 0; k < nkind; k++) {
        int id = fpair[k];
        double p, f;
        if (!orc_pot_nn(lib, id, 1.0e-8, &p, &f)) return -2; /* unregistered id: reference stops (:930-940) */
        for (int i = 1; i <= ntab; i++) {
            double t = (double)i * csiv;
            double r = t * t;
            orc_pot_nn(lib, id, r, &p, &f);
            potr[(size_t)(i - 1) * nkind + k] = p * r;
            fpotr[(size_t)(i - 1) * nkind + k] = f * r;
        }
        if (orc_pot_rho(lib, id, 1.0e-8, &p, &f)) {
            for (int i = 1; i <= ntab; i++) {
                double t = (double)i * csiv;
                double r = t * t;
                orc_pot_rho(lib, id, r, &p, &f);
                potb[(size_t)(i - 1) * nkind + k] = p;
                fpotb[(size_t)(i - 1) * nkind + k] = f;
                if (rhomx < p * rhoscal) rhomx = p * rhoscal; /* :965 */
            }
        } else {
            for (int i = 1; i <= ntab; i++) {
                potb[(size_t)(i - 1) * nkind + k] = 0.0;
                fpotb[(size_t)(i - 1) * nkind + k] = 0.0;
            }
        }
        rhod = rhomx / (double)nembd; /* :976 */
    }
    /* NOTE: the reference creates pair tables in (I,J) loop order, which visits ids in the
     * same first-appearance order as fpair[]; RHOMX is a running max so the order is immaterial. */
    for (int i = 0; i < ng; i++)
        for (int j = 0; j < ng; j++) {
            kpair[i + ng * j] = -1;
            for (int k = 0; k < nkind; k++)
                if (fpair[k] == ptype[i + ng * j]) { kpair[i + ng * j] = k + 1; break; }
        }
    /* :1207-1224 -> Create_EMBDFUNTable :1043-1052 */
    for (int k = 0; k < nkind1; k++) {
        int id = fpair1[k];
        double p, f;
        int has = orc_pot_embd(lib, id, 1.0, &p, &f);
        for (int i = 1; i <= nembd; i++) {
            double r = (double)(i - 1) * rhod;
            if (has) orc_pot_embd(lib, id, r, &p, &f);
            else { p = 0.0; f = 0.0; }
            fembd[(size_t)(i - 1) * nkind1 + k] = p;
            dfembd[(size_t)(i - 1) * nkind1 + k] = f;
        }
    }
    for (int i = 0; i < ng; i++) {
        kembd[i] = -1;
        for (int k = 0; k < nkind1; k++)
            if (fpair1[k] == ptype[i + ng * i]) { kembd[i] = k + 1; break; }
    }
    if (rhod <= 1.0e-64) rhod = 1.0; /* :1223 */
    *nkind_out = nkind;
    *nkind1_out = nkind1;
    *csi_out = csi;
    *rhod_out = rhod;
    return 0;
}

/* =====================================================================================
 * Cells and neighbour lists
 * ===================================================================================== */
static const double CELL_EPS = (double)0.0001f; /* real(KINDDF),parameter::EPS=0.0001 : a REAL literal */

/* scan order of the 27 cells, MD_NeighborsList_GPU.F90:218-220 (same in Common/MD_NeighborsList.F90:427-429) */
static const int NIX[27] = {0, -1, -1, -1, 0, 0, -1, 1, -1, 0, 1, -1, 0, 1, 1, 1, 1, 0, 0, -1, 1, -1, 0, 1, -1, 0, 1};
static const int NIY[27] = {0, 0, -1, 1, 1, 0, 0, 0, -1, -1, -1, 1, 1, 1, 0, 1, -1, -1, 0, 0, 0, -1, -1, -1, 1, 1, 1};
static const int NIZ[27] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1};

void orc_ncell(const double zl[3], double nb_rm_max, int ncell[3])
{
    /* MD_NeighborsList_GPU.F90:289-298 */
    for (int k = 0; k < 3; k++) {
        ncell[k] = (int)(zl[k] / (1.0 * nb_rm_max) - CELL_EPS);
        if (ncell[k] < 3) ncell[k] = 3;
    }
}

int orc_nlist_build_dev(int nbox, int napb, const double *xp, const int *ityp, int *statu,
                        const double boxlow[3], const double zl[3], const int ifpd[3],
                        const double boxshape[9], int ng, const double *nb_rm, int mxkvois,
                        int ncell[3], int *inc, int *gid, int *nac, int *naac, int *ia1th,
                        int *kvois, int *indi, int *nn_max)
{
    const int n = nbox * napb;
    double rmmax = 0.0;
    for (int i = 0; i < ng * ng; i++)
        if (nb_rm[i] > rmmax) rmmax = nb_rm[i];
    orc_ncell(zl, rmmax, ncell);
    const int ncx = ncell[0], ncy = ncell[1], ncz = ncell[2];
    const int nc0 = ncx * ncy * ncz, nc = nc0 * nbox;

    /* NeighboreList_IC_KERNEL :798-815 (first partition: GID = identity, :1455-1457) */
    for (int i = 0; i < n; i++) {
        int ib = i / napb;
        if ((statu[i] & ORC_STATU_OUTOFBOX) != ORC_STATU_OUTOFBOX) {
            int ix = (int)((xp[i] - boxlow[0]) / zl[0] * (double)ncx - CELL_EPS);
            int iy = (int)((xp[i + n] - boxlow[1]) / zl[1] * (double)ncy - CELL_EPS);
            int iz = (int)((xp[i + 2 * n] - boxlow[2]) / zl[2] * (double)ncz - CELL_EPS);
            if (ix < 0 || ix >= ncx || iy < 0 || iy >= ncy || iz < 0 || iz >= ncz)
                inc[i] = -2;
            else
                inc[i] = 1 + (ix + ncx * (iy + ncy * iz)) + ib * nc0;
        } else {
            inc[i] = -1;
        }
    }
    /* host linked-cell build :1490-1528 */
    int *head = (int *)calloc((size_t)nc + 1, sizeof(int));
    int *link = (int *)calloc((size_t)n + 1, sizeof(int));
    memset(nac, 0, sizeof(int) * (size_t)nc);
    memset(naac, 0, sizeof(int) * (size_t)nc);
    int numout = 0;
    for (int i = 1; i <= n; i++) {
        int ic = inc[i - 1];
        if (ic > 0) {
            link[i] = head[ic];
            head[ic] = i;
            nac[ic - 1]++;
            if ((statu[i - 1] & ORC_STATU_ACTIVE) == ORC_STATU_ACTIVE) naac[ic - 1]++;
        } else {
            /* the reference prompts on stdin here (:1506-1523); the 'C'ontinue branch: */
            if (ic < -1) statu[i - 1] = ORC_STATU_OUTOFBOX;
            numout++;
        }
    }
    /* cell-ordered gather :1541-1570 : cells ascending, chain order = descending original id */
    int ip = 0;
    for (int ic = 1; ic <= nc; ic++) {
        ia1th[ic - 1] = (ic > 1) ? ia1th[ic - 2] + nac[ic - 2] : 1;
        for (int id = head[ic]; id > 0; id = link[id]) gid[ip++] = id;
    }
    /* out-of-box atoms go to the end, backward :1627-1637 */
    if (numout > 0) {
        ip = n;
        for (int i = 1; i <= n; i++)
            if (inc[i - 1] < 0) gid[--ip] = i;
    }
    free(head);
    free(link);

    /* sorted copies used by the list kernel (hm_XP, hm_ITYP) */
    float *px = (float *)malloc(sizeof(float) * (size_t)n * 3); /* (float)XP of the sorted atoms  */
    double *sx = (double *)malloc(sizeof(double) * (size_t)n * 3);
    int *sty = (int *)malloc(sizeof(int) * (size_t)n);
    for (int s = 0; s < n; s++) {
        int o = gid[s] - 1;
        for (int d = 0; d < 3; d++) {
            sx[s + d * n] = xp[o + d * n];
            px[s + d * n] = (float)xp[o + d * n];
        }
        sty[s] = ityp[o];
    }
    /* fp32 constants :1376-1377,1393-1394,1414-1415 */
    float bs[9], rm2[ORC_MXGROUP * ORC_MXGROUP];
    for (int i = 0; i < 9; i++) bs[i] = (float)boxshape[i];
    for (int i = 0; i < ng * ng; i++) rm2[i] = (float)(nb_rm[i] * nb_rm[i]);
    const float fbs[3] = {(float)zl[0], (float)zl[1], (float)zl[2]}; /* CXYZ is real(KINDSF) :955,1034 */

    memset(kvois, 0, sizeof(int) * (size_t)n); /* DevSet(KVOIS,0) at init :285 */
    int nnmx = 0;
    /* Cal_NeighboreList_Kernel2C :963-1196, one "block" per cell */
#pragma omp parallel for schedule(dynamic, 64) reduction(max : nnmx)
    for (int ic0 = 0; ic0 < nc; ic0++) {
        if (naac[ic0] <= 0) continue; /* :982 */
        if (nac[ic0] <= 0) continue;  /* :1018 */
        int is0 = ic0 / nc0, icl = ic0 - is0 * nc0;
        int iz0 = icl / (ncx * ncy), iy0 = (icl - iz0 * ncx * ncy) / ncx, ix0 = icl - iz0 * ncx * ncy - iy0 * ncx;
        int cid[27];
        float cxyz[27][3];
        for (int k = 0; k < 27; k++) {
            int c[3] = {ix0 + NIX[k], iy0 + NIY[k], iz0 + NIZ[k]};
            const int ncs[3] = {ncx, ncy, ncz};
            int out = 0;
            for (int d = 0; d < 3; d++) {
                cxyz[k][d] = 0.0f;
                if (ifpd[d] && k > 0) { /* :1031-1059 */
                    if (c[d] >= ncs[d]) { c[d] = 0; cxyz[k][d] = fbs[d]; }
                    else if (c[d] < 0) { c[d] = ncs[d] - 1; cxyz[k][d] = -fbs[d]; }
                }
                if (c[d] >= ncs[d] || c[d] < 0) out = 1;
            }
            cid[k] = out ? -1 : (ncx * ncy * c[2] + ncx * c[1] + c[0] + is0 * nc0);
        }
        int a0 = ia1th[ic0] - 1, na = nac[ic0];
        for (int ia = a0; ia < a0 + na; ia++) {
            float p1 = px[ia], p2 = px[ia + n], p3 = px[ia + 2 * n];
            int ity = sty[ia], nn = 0;
            for (int k = 0; k < 27; k++) {
                if (cid[k] < 0) continue;
                int j0 = ia1th[cid[k]] - 1, nj = nac[cid[k]];
                for (int ja = j0; ja < j0 + nj; ja++) {
                    /* SPOS = XP + CXYZ : double + float -> double, stored to float :1100-1102 */
                    float s1 = (float)(sx[ja] + (double)cxyz[k][0]);
                    float s2 = (float)(sx[ja + n] + (double)cxyz[k][1]);
                    float s3 = (float)(sx[ja + 2 * n] + (double)cxyz[k][2]);
                    float e1 = p1 - s1, e2 = p2 - s2, e3 = p3 - s3;
                    float d1 = bs[0] * e1 + bs[3] * e2 + bs[6] * e3; /* BOXSHAPE(1,1:3), column-major */
                    float d2 = bs[1] * e1 + bs[4] * e2 + bs[7] * e3;
                    float d3 = bs[2] * e1 + bs[5] * e2 + bs[8] * e3;
                    if (d1 * d1 + d2 * d2 + d3 * d3 <= rm2[(ity - 1) + ng * (sty[ja] - 1)]) { /* :1123 */
                        if (k == 0 && ja == ia) continue; /* :1124 */
                        nn++;
                        if (nn <= mxkvois) indi[ia + (size_t)(nn - 1) * n] = ja + 1;
                    }
                }
            }
            kvois[ia] = nn < mxkvois ? nn : mxkvois; /* :1195 */
            if (nn > nnmx) nnmx = nn;
        }
    }
    if (nn_max) *nn_max = nnmx;
    free(px);
    free(sx);
    free(sty);
    return numout;
}

/* half = 0: Cal_NeighboreList2C (every directed pair), Common/MD_NeighborsList.F90:396-633
 * half = 1: Cal_NeighboreListC (Newton's third law: neighbour cells 2..14 of the NIX/NIY/NIZ table and J = I+1..N inside
 *           the gathered sub-box, so every pair is stored once), :152-391 */
static int nlist_build_cpu(int half, int n, const double *xp, const int *ityp, const int *statu,
                           const double boxlow[3], const double zl[3], const int ifpd[3],
                           const double boxshape[9], int ng, const double *nb_rm, int mxkvois,
                           int *kvois, int *indi)
{
    /* Common/MD_NeighborsList.F90:432-633 */
    const int kcell_hi = half ? 14 : 27;
    double rmmax = 0.0, rcut2[ORC_MXGROUP * ORC_MXGROUP];
    for (int i = 0; i < ng * ng; i++) {
        if (nb_rm[i] > rmmax) rmmax = nb_rm[i];
        rcut2[i] = nb_rm[i] * nb_rm[i];
    }
    int ncell[3];
    for (int k = 0; k < 3; k++) { /* :474-483 */
        ncell[k] = (int)(zl[k] / (1.0 * rmmax) - CELL_EPS);
        if (ncell[k] < 3) ncell[k] = 3;
    }
    const int ncx = ncell[0], ncy = ncell[1], ncz = ncell[2], nc = ncx * ncy * ncz;
    int *head = (int *)calloc((size_t)nc + 1, sizeof(int));
    int *link = (int *)calloc((size_t)n + 1, sizeof(int));
    int err = 0;
    for (int i = 1; i <= n; i++) { /* :495-526 */
        if ((statu[i - 1] & ORC_STATU_OUTOFBOX) != ORC_STATU_OUTOFBOX) {
            int c[3];
            for (int k = 0; k < 3; k++)
                c[k] = (int)((xp[(i - 1) + k * n] - boxlow[k]) / zl[k] * (double)ncell[k] - CELL_EPS);
            if (c[0] < 0 || c[0] > ncx || c[1] < 0 || c[1] > ncy || c[2] < 0 || c[2] > ncz) { err = -2; break; }
            if (c[0] >= ncx || c[1] >= ncy || c[2] >= ncz) { err = -2; break; } /* would index past HEAD */
            int ic = 1 + (c[0] + ncx * (c[1] + ncy * c[2]));
            link[i] = head[ic];
            head[ic] = i;
        } else {
            kvois[i - 1] = 0;
        }
    }
    if (err) { free(head); free(link); return err; }

#pragma omp parallel
    {
        int cap = 1024;
        int *ident = (int *)malloc(sizeof(int) * cap);
        int *typ = (int *)malloc(sizeof(int) * cap);
        double *xpt = (double *)malloc(sizeof(double) * 3 * cap);
#pragma omp for schedule(dynamic, 16)
        for (int ic0 = 1; ic0 <= nc; ic0++) {
            if (head[ic0] == 0 || err) continue;
            int iz = (ic0 - 1) / (ncx * ncy) + 1, iy = ((ic0 - 1) - (iz - 1) * ncx * ncy) / ncx + 1,
                ix = (ic0 - 1) - (iz - 1) * ncx * ncy - (iy - 1) * ncx + 1;
            int nloc = 0, n0;
#define PUSH(ID, CX, CY, CZ)                                                            \
    do {                                                                                \
        if (nloc == cap) {                                                              \
            cap *= 2;                                                                   \
            ident = (int *)realloc(ident, sizeof(int) * cap);                           \
            typ = (int *)realloc(typ, sizeof(int) * cap);                               \
            xpt = (double *)realloc(xpt, sizeof(double) * 3 * cap);                     \
        }                                                                               \
        ident[nloc] = (ID);                                                             \
        xpt[3 * nloc] = xp[(ID)-1] + (CX);                                              \
        xpt[3 * nloc + 1] = xp[(ID)-1 + n] + (CY);                                      \
        xpt[3 * nloc + 2] = xp[(ID)-1 + 2 * n] + (CZ);                                  \
        typ[nloc] = ityp[(ID)-1];                                                       \
        nloc++;                                                                         \
    } while (0)
            for (int id = head[ic0]; id > 0; id = link[id]) PUSH(id, 0.0, 0.0, 0.0);
            n0 = nloc;
            for (int k = 1; k < kcell_hi; k++) { /* :555-588 (2C) / :317-349 (C: IC = 2,14) */
                int j[3] = {ix + NIX[k], iy + NIY[k], iz + NIZ[k]};
                double cx[3] = {0.0, 0.0, 0.0};
                int out = 0;
                for (int d = 0; d < 3; d++) {
                    if (ifpd[d]) {
                        if (j[d] > ncell[d]) { j[d] = 1; cx[d] = zl[d]; }
                        else if (j[d] < 1) { j[d] = ncell[d]; cx[d] = -zl[d]; }
                    }
                    if (j[d] > ncell[d] || j[d] < 1) out = 1;
                }
                if (out) continue;
                for (int id = head[j[0] + ncx * ((j[1] - 1) + ncy * (j[2] - 1))]; id > 0; id = link[id])
                    PUSH(id, cx[0], cx[1], cx[2]);
            }
#undef PUSH
            for (int i = 0; i < n0; i++) { /* :594-621 */
                int id = ident[i], itp = typ[i], nn = 0;
                for (int jj = half ? i + 1 : 0; jj < nloc; jj++) { /* 2C: J = 1..N, J /= I (:597-598); C: J = I+1..N (:356) */
                    if (jj == i) continue;
                    double c1 = xpt[3 * i] - xpt[3 * jj], c2 = xpt[3 * i + 1] - xpt[3 * jj + 1],
                           c3 = xpt[3 * i + 2] - xpt[3 * jj + 2];
                    double d1 = boxshape[0] * c1 + boxshape[3] * c2 + boxshape[6] * c3;
                    double d2 = boxshape[1] * c1 + boxshape[4] * c2 + boxshape[7] * c3;
                    double d3 = boxshape[2] * c1 + boxshape[5] * c2 + boxshape[8] * c3;
                    if (d1 * d1 + d2 * d2 + d3 * d3 < rcut2[(itp - 1) + ng * (typ[jj] - 1)]) { /* :605 */
                        nn++;
                        if (nn > mxkvois) { err = -1; break; } /* reference stops :608-612 */
                        indi[(id - 1) + (size_t)(nn - 1) * n] = ident[jj];
                    }
                }
                if (err) break;
                kvois[id - 1] = nn;
            }
        }
        free(ident);
        free(typ);
        free(xpt);
    }
    free(head);
    free(link);
    return err;
}

int orc_nlist_build_cpu(int n, const double *xp, const int *ityp, const int *statu,
                        const double boxlow[3], const double zl[3], const int ifpd[3],
                        const double boxshape[9], int ng, const double *nb_rm, int mxkvois,
                        int *kvois, int *indi)
{
    return nlist_build_cpu(0, n, xp, ityp, statu, boxlow, zl, ifpd, boxshape, ng, nb_rm, mxkvois, kvois, indi);
}
int orc_nlist_build_cpu_half(int n, const double *xp, const int *ityp, const int *statu,
                             const double boxlow[3], const double zl[3], const int ifpd[3],
                             const double boxshape[9], int ng, const double *nb_rm, int mxkvois,
                             int *kvois, int *indi)
{
    return nlist_build_cpu(1, n, xp, ityp, statu, boxlow, zl, ifpd, boxshape, ng, nb_rm, mxkvois, kvois, indi);
}

/* =====================================================================================
 * Forces
 * ===================================================================================== */
/* T(k,KK) in Fortran layout; KK outside 1..nt is out of bounds in the reference
 * (r == Rmax reads KK+1 = NTAB+1, MD_EAM_ForceTable_GPU.F90:522-530); defined as 0 here. */
static inline double tab(const double *t, int nk, int nt, int k, int kk)
{
    if (kk < 1 || kk > nt) return 0.0;
    return t[(size_t)(kk - 1) * nk + (k - 1)];
}

/* separation with the pairwise minimum image, :497-510 */
static inline void min_image(double *s, const double zl[3], const int ifpd[3])
{
    for (int d = 0; d < 3; d++)
        if (ifpd[d] > 0 && fabs(s[d]) > zl[d] * 0.5) s[d] = s[d] - copysign(zl[d], s[d]);
}

void orc_force_pass1(int n, int ia0, int npart, const double *xp, const int *ityp, const int *statu,
                     const int *kvois, const int *indi, int ldindi, const double zl[3], const int ifpd[3],
                     const double bs[9], const orc_tables *t, double *den)
{
#pragma omp parallel for schedule(static)
    for (int ic = 1; ic <= npart; ic++) {
        double den0 = 0.0;
        if ((statu[ic - 1] & ORC_STATU_ACTIVE) == ORC_STATU_ACTIVE) {
            int gi = ic + ia0 - 1;
            double px = xp[gi], py = xp[gi + n], pz = xp[gi + 2 * n];
            int ti = ityp[gi], iiw = kvois[ic - 1];
            for (int iw = 0; iw < iiw; iw++) {
                int j = indi[(ic - 1) + (size_t)iw * ldindi] - 1;
                double s[3] = {px - xp[j], py - xp[j + n], pz - xp[j + 2 * n]};
                min_image(s, zl, ifpd);
                double dx = bs[0] * s[0] + bs[3] * s[1] + bs[6] * s[2]; /* :515-517 */
                double dy = bs[1] * s[0] + bs[4] * s[1] + bs[7] * s[2];
                double dz = bs[2] * s[0] + bs[5] * s[1] + bs[8] * s[2];
                double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 <= t->ru2max) { /* :522 */
                    int ktab = t->kpair[(ti - 1) + t->ng * (ityp[j] - 1)];
                    double r = sqrt(r2);
                    double sk = sqrt(r) * t->csi;
                    int kk = (int)sk;
                    double a = tab(t->potb, t->nkind, t->ntab, ktab, kk);
                    den0 = den0 + (a + (sk - (double)kk) * (tab(t->potb, t->nkind, t->ntab, ktab, kk + 1) - a)); /* :530 */
                }
            }
            if (t->pot_type == ORC_POT_FS) {
                /* MD_FS_ForceTable_GPU.F90:497-507: DEN = -0.5/sqrt(rho), guarded rho>0 */
                if (den0 > 0.0) den0 = -0.5 / sqrt(den0);
            } else if (den0 > 0.0) { /* :535-541 */
                int ktab = t->kembd[ti - 1];
                double sk = den0 / t->rhod + 1.0;
                int kk = (int)(sk + 0.000001);
                double a = tab(t->dfembd, t->nkind1, t->nembd, ktab, kk);
                den0 = a + (sk - (double)kk) * (tab(t->dfembd, t->nkind1, t->nembd, ktab, kk + 1) - a);
            }
        }
        den[ic + ia0 - 1] = den0; /* :544 */
    }
}

void orc_force_pass2(int n, int ia0, int npart, const double *xp, const int *ityp, const int *statu,
                     const int *kvois, const int *indi, int ldindi, const double zl[3], const int ifpd[3],
                     const double bs[9], const orc_tables *t, const double *den, double *fp, int ldfp,
                     double *vtensor)
{
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp parallel for schedule(static) reduction(+ : v[:9])
    for (int ic = 1; ic <= npart; ic++) {
        double fx = 0.0, fy = 0.0, fz = 0.0;
        if ((statu[ic - 1] & ORC_STATU_ACTIVE) == ORC_STATU_ACTIVE) {
            int gi = ic + ia0 - 1;
            double px = xp[gi], py = xp[gi + n], pz = xp[gi + 2 * n];
            int ti = ityp[gi], iiw = kvois[ic - 1];
            double denki = den[gi];
            for (int iw = 0; iw < iiw; iw++) {
                int j = indi[(ic - 1) + (size_t)iw * ldindi] - 1;
                double s[3] = {px - xp[j], py - xp[j + n], pz - xp[j + 2 * n]};
                min_image(s, zl, ifpd);
                double dx, dy, dz, r2;
                if (vtensor) { /* CALPTENSOR_KERNEL :1171-1174 applies BOXSHAPE */
                    dx = bs[0] * s[0] + bs[3] * s[1] + bs[6] * s[2];
                    dy = bs[1] * s[0] + bs[4] * s[1] + bs[7] * s[2];
                    dz = bs[2] * s[0] + bs[5] * s[1] + bs[8] * s[2];
                    r2 = dx * dx + dy * dy + dz * dz;
                } else { /* CALFORCE_KERNEL :775 does not */
                    dx = s[0]; dy = s[1]; dz = s[2];
                    r2 = s[0] * s[0] + s[1] * s[1] + s[2] * s[2];
                }
                if (r2 <= t->ru2max) {
                    int tj = ityp[j];
                    int k0 = t->kpair[(ti - 1) + t->ng * (tj - 1)];
                    int k1 = t->kpair[(tj - 1) + t->ng * (ti - 1)];
                    double r = sqrt(r2);
                    double sk = sqrt(r) * t->csi;
                    int kk = (int)sk;
                    double dk = sk - (double)kk;
                    double denkj = den[j];
                    double a = tab(t->fpotr, t->nkind, t->ntab, k0, kk);
                    double b = tab(t->fpotb, t->nkind, t->ntab, k0, kk);
                    double c = tab(t->fpotb, t->nkind, t->ntab, k1, kk);
                    /* :811-813 */
                    double fortot = (a + dk * (tab(t->fpotr, t->nkind, t->ntab, k0, kk + 1) - a)) / r2 +
                                    ((b + dk * (tab(t->fpotb, t->nkind, t->ntab, k0, kk + 1) - b)) * denki +
                                     (c + dk * (tab(t->fpotb, t->nkind, t->ntab, k1, kk + 1) - c)) * denkj) / r;
                    fx = fx + fortot * s[0];
                    fy = fy + fortot * s[1];
                    fz = fz + fortot * s[2];
                    if (vtensor) { /* :1222-1232 ; P(a,b) stored column-major v[a+3b] */
                        fortot = fortot * 0.5;
                        v[0] += dx * dx * fortot; v[3] += dx * dy * fortot; v[6] += dx * dz * fortot;
                        v[1] += dy * dx * fortot; v[4] += dy * dy * fortot; v[7] += dy * dz * fortot;
                        v[2] += dz * dx * fortot; v[5] += dz * dy * fortot; v[8] += dz * dz * fortot;
                    }
                }
            }
        }
        fp[ic - 1] = fx;
        fp[ic - 1 + ldfp] = fy;
        fp[ic - 1 + 2 * ldfp] = fz;
    }
    if (vtensor)
        for (int i = 0; i < 9; i++) vtensor[i] = v[i];
}

/* CAL_EAM_AtomicStress_KERNEL, CommonGPU/MD_EAM_ForceTable_GPU.F90:1775-1925: AP(NAPDEV,9), column q = 3*(a-1)+b,
 * the whole pair term to atom i; BOXSHAPE applied to the separation (:1861-1864); DEN as left by pass 1. */
void orc_force_avstress(int n, int ia0, int npart, const double *xp, const int *ityp, const int *statu,
                        const int *kvois, const int *indi, int ldindi, const double zl[3], const int ifpd[3],
                        const double bs[9], const orc_tables *t, const double *den, double *ap, int ldap)
{
#pragma omp parallel for schedule(static)
    for (int ic = 1; ic <= npart; ic++) {
        double p[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if ((statu[ic - 1] & ORC_STATU_ACTIVE) == ORC_STATU_ACTIVE) {
            int gi = ic + ia0 - 1;
            double px = xp[gi], py = xp[gi + n], pz = xp[gi + 2 * n];
            int ti = ityp[gi], iiw = kvois[ic - 1];
            double denki = den[gi];
            for (int iw = 0; iw < iiw; iw++) {
                int j = indi[(ic - 1) + (size_t)iw * ldindi] - 1;
                double s[3] = {px - xp[j], py - xp[j + n], pz - xp[j + 2 * n]};
                min_image(s, zl, ifpd);
                double d[3];
                d[0] = bs[0] * s[0] + bs[3] * s[1] + bs[6] * s[2];
                d[1] = bs[1] * s[0] + bs[4] * s[1] + bs[7] * s[2];
                d[2] = bs[2] * s[0] + bs[5] * s[1] + bs[8] * s[2];
                double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (r2 <= t->ru2max) {
                    int tj = ityp[j];
                    int k0 = t->kpair[(ti - 1) + t->ng * (tj - 1)];
                    int k1 = t->kpair[(tj - 1) + t->ng * (ti - 1)];
                    double r = sqrt(r2);
                    double sk = sqrt(r) * t->csi;
                    int kk = (int)sk;
                    double dk = sk - (double)kk;
                    double a = tab(t->fpotr, t->nkind, t->ntab, k0, kk);
                    double b = tab(t->fpotb, t->nkind, t->ntab, k0, kk);
                    double c = tab(t->fpotb, t->nkind, t->ntab, k1, kk);
                    double fortot = (a + dk * (tab(t->fpotr, t->nkind, t->ntab, k0, kk + 1) - a)) / r2 +
                                    ((b + dk * (tab(t->fpotb, t->nkind, t->ntab, k0, kk + 1) - b)) * denki +
                                     (c + dk * (tab(t->fpotb, t->nkind, t->ntab, k1, kk + 1) - c)) * den[j]) / r;
                    for (int x = 0; x < 3; x++)
                        for (int y = 0; y < 3; y++) p[3 * x + y] = p[3 * x + y] + d[x] * d[y] * fortot; /* :1882-1890 */
                }
            }
        }
        for (int q = 0; q < 9; q++) ap[(ic - 1) + (size_t)q * ldap] = p[q];
    }
}

void orc_force_epot(int n, int ia0, int npart, const double *xp, const int *ityp, const int *statu,
                    const int *kvois, const int *indi, int ldindi, const double zl[3], const int ifpd[3],
                    const double bs[9], const orc_tables *t, double *epot)
{
#pragma omp parallel for schedule(static)
    for (int ic = 1; ic <= npart; ic++) {
        double er0 = 0.0, den0 = 0.0;
        if ((statu[ic - 1] & ORC_STATU_ACTIVE) == ORC_STATU_ACTIVE) {
            int gi = ic + ia0 - 1;
            double px = xp[gi], py = xp[gi + n], pz = xp[gi + 2 * n];
            int ti = ityp[gi], iiw = kvois[ic - 1];
            for (int iw = 0; iw < iiw; iw++) {
                int j = indi[(ic - 1) + (size_t)iw * ldindi] - 1;
                double s[3] = {px - xp[j], py - xp[j + n], pz - xp[j + 2 * n]};
                min_image(s, zl, ifpd);
                double dx = bs[0] * s[0] + bs[3] * s[1] + bs[6] * s[2];
                double dy = bs[1] * s[0] + bs[4] * s[1] + bs[7] * s[2];
                double dz = bs[2] * s[0] + bs[5] * s[1] + bs[8] * s[2];
                double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 <= t->ru2max) {
                    int ktab = t->kpair[(ti - 1) + t->ng * (ityp[j] - 1)];
                    double r = sqrt(r2);
                    double sk = sqrt(r) * t->csi;
                    int kk = (int)sk;
                    double dk = sk - (double)kk;
                    double a = tab(t->potr, t->nkind, t->ntab, ktab, kk);
                    double b = tab(t->potb, t->nkind, t->ntab, ktab, kk);
                    er0 = er0 + (a + dk * (tab(t->potr, t->nkind, t->ntab, ktab, kk + 1) - a)) / r; /* :1622 */
                    den0 = den0 + (b + dk * (tab(t->potb, t->nkind, t->ntab, ktab, kk + 1) - b));  /* :1623 */
                }
            }
            if (t->pot_type == ORC_POT_FS) {
                /* MD_FS_ForceTable_GPU.F90:1606 : E = sum 0.5V - sqrt(rho) */
                den0 = -sqrt(den0);
            } else { /* :1628-1631, no rho>0 guard here */
                int ktab = t->kembd[ti - 1];
                double sk = den0 / t->rhod + 1.0;
                int kk = (int)(sk + 0.000001);
                double a = tab(t->fembd, t->nkind1, t->nembd, ktab, kk);
                den0 = a + (sk - (double)kk) * (tab(t->fembd, t->nkind1, t->nembd, ktab, kk + 1) - a);
            }
        }
        epot[ic - 1] = er0 + den0; /* :1633 */
    }
}

/* =====================================================================================
 * Integrator, EPC
 * ===================================================================================== */
void orc_predictor(int n, double *xp, double *xp1, const double *fp, double *dis, int *statu,
                   const int *ityp, const double *cm, double h, const double lb[3], const double ub[3],
                   const double zl[3], const int ifpd[3])
{
    /* Predictor_DEV, MD_DiffScheme_GPU.F90:660-662: TH = H, HS2 = H/2, H2S2 = H*H/2 */
    const double th = h, hs2 = th * 0.5, h2s2 = th * th * 0.5;
    static const int FIXP[3] = {ORC_STATU_FIXPOSX, ORC_STATU_FIXPOSY, ORC_STATU_FIXPOSZ};
    static const int FIXV[3] = {ORC_STATU_FIXVELX, ORC_STATU_FIXVELY, ORC_STATU_FIXVELZ};
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        int stat = statu[i];
        if ((stat & ORC_STATU_ACTIVE) != ORC_STATU_ACTIVE) continue;
        double cm0 = cm[ityp[i] - 1];
        double x[3], f[3];
        for (int d = 0; d < 3; d++) {
            f[d] = fp[i + d * n] / cm0;                                /* :309-311 */
            double dd = th * xp1[i + d * n] + h2s2 * f[d];             /* :314 */
            if ((stat & FIXP[d]) == FIXP[d]) dd = 0.0;
            x[d] = xp[i + d * n] + dd;
            if (ifpd[d]) {                                             /* :317-325 */
                if (x[d] > ub[d]) x[d] = x[d] - zl[d];
                else if (x[d] < lb[d]) x[d] = x[d] + zl[d];
            }
            dis[i + d * n] = dis[i + d * n] + dd;                      /* :371-373 */
        }
        for (int d = 0; d < 3; d++) {
            if ((stat & FIXV[d]) == 0 && (stat & FIXP[d]) == 0)       /* :353-361 */
                xp1[i + d * n] = xp1[i + d * n] + hs2 * f[d];
            xp[i + d * n] = x[d];
        }
        /* :375-379 ; the PASSBOUND bit set at :320 is never stored */
        if (x[0] > ub[0] || x[1] > ub[1] || x[2] > ub[2]) statu[i] = ORC_STATU_OUTOFBOX | ORC_STATU_REFLECT;
        else if (x[0] < lb[0] || x[1] < lb[1] || x[2] < lb[2]) statu[i] = ORC_STATU_OUTOFBOX | ORC_STATU_TRANSMIT;
    }
}

void orc_corrector(int n, double *xp1, const double *fp, const int *statu, const int *ityp, const double *cm,
                   double h)
{
    const double hs2 = h * 0.5;
    static const int FIXP[3] = {ORC_STATU_FIXPOSX, ORC_STATU_FIXPOSY, ORC_STATU_FIXPOSZ};
    static const int FIXV[3] = {ORC_STATU_FIXVELX, ORC_STATU_FIXVELY, ORC_STATU_FIXVELZ};
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        int stat = statu[i];
        if ((stat & ORC_STATU_ACTIVE) != ORC_STATU_ACTIVE) continue;
        double cm0 = cm[ityp[i] - 1];
        for (int d = 0; d < 3; d++) {
            double f = fp[i + d * n] / cm0; /* :735-737 */
            if ((stat & FIXV[d]) == 0 && (stat & FIXP[d]) == 0) xp1[i + d * n] = xp1[i + d * n] + hs2 * f;
        }
    }
}

void orc_ekin(int n, const double *xp1, const int *statu, const int *ityp, const double *cm, double *ekin)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        double ek = -1.0e32; /* :886 */
        if ((statu[i] & ORC_STATU_ACTIVE) == ORC_STATU_ACTIVE && (statu[i] & ORC_STATU_FIXPOS) == 0) {
            double cm0 = cm[ityp[i] - 1];
            double vx = xp1[i], vy = xp1[i + n], vz = xp1[i + 2 * n];
            ek = 0.5 * cm0 * (vx * vx + vy * vy + vz * vz); /* :896 */
        }
        ekin[i] = ek;
    }
}

void orc_epc(int n, const double *xp1, double *fp, const int *statu, const int *ityp, int ng, const int *enable,
             const double *cm, const double *te, const double *alpha, const double *cut, const double *he)
{
    double v2ti[ORC_MXGROUP], epa[ORC_MXGROUP], tcut[ORC_MXGROUP], eup[ORC_MXGROUP];
    for (int g = 0; g < ng; g++) { /* Reset_EPCMOD_DEV, MD_EP_Coupling_GPU.F90:387-394 */
        v2ti[g] = cm[g] * (1.0 / 3.0) / ORC_KB;
        epa[g] = cm[g] / alpha[g];
        tcut[g] = te[g] * cut[g];
        eup[g] = 2.0 * he[g] / cm[g];
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) { /* EPC_MOD_KERNEL :471-491 */
        int kk = ityp[i] - 1;
        if (enable[kk] <= 0) continue;
        if ((statu[i] & ORC_STATU_ACTIVE) != ORC_STATU_ACTIVE) continue;
        double vx = xp1[i], vy = xp1[i + n], vz = xp1[i + 2 * n];
        double v2 = vx * vx + vy * vy + vz * vz;
        if (v2 <= eup[kk]) {
            double tm = v2 * v2ti[kk];
            double mu = epa[kk] * (tm - te[kk]) / (tm > tcut[kk] ? tm : tcut[kk]);
            fp[i] = fp[i] - mu * vx;
            fp[i + n] = fp[i + n] - mu * vy;
            fp[i + 2 * n] = fp[i + 2 * n] - mu * vz;
        }
    }
}

/* =====================================================================================
 * Whole-step driver in cell-sorted order
 * ===================================================================================== */
struct orc_md {
    int nbox, napb, n, ng, mxkvois, haspart;
    double boxlow[3], boxup[3], zl[3], bs[9], cm[ORC_MXGROUP], nb_rm[ORC_MXGROUP * ORC_MXGROUP];
    int ifpd[3], ncell[3];
    orc_tables t;
    /* sorted-order state */
    double *xp, *xp1, *fp, *dis, *den, *epot, *ekin;
    int *ityp, *statu, *gid, *kvois, *indi;
    /* original-order scratch */
    double *oxp, *tmp;
    int *oityp, *ostatu, *inc, *ngid, *nac, *naac, *ia1th, *itmp;
    double vtensor[9];
    int epc_on, epc_enable[ORC_MXGROUP];
    double epc_te[ORC_MXGROUP], epc_alpha[ORC_MXGROUP], epc_cut[ORC_MXGROUP], epc_he[ORC_MXGROUP];
};

orc_md *orc_md_create(int nbox, int napb, const double *xp, const double *xp1, const int *ityp, const int *statu,
                      int ng, const double *cm, const double boxlow[3], const double zl[3], const int ifpd[3],
                      const double *nb_rm, int mxkvois, const orc_tables *t)
{
    orc_md *m = (orc_md *)calloc(1, sizeof(orc_md));
    int n = nbox * napb;
    m->nbox = nbox; m->napb = napb; m->n = n; m->ng = ng; m->mxkvois = mxkvois;
    for (int d = 0; d < 3; d++) {
        m->boxlow[d] = boxlow[d];
        m->zl[d] = zl[d];
        m->boxup[d] = boxlow[d] + zl[d];
        m->ifpd[d] = ifpd[d];
    }
    m->bs[0] = m->bs[4] = m->bs[8] = 1.0;
    memcpy(m->cm, cm, sizeof(double) * ng);
    memcpy(m->nb_rm, nb_rm, sizeof(double) * ng * ng);
    m->t = *t;
    size_t n3 = (size_t)n * 3;
    m->xp = (double *)malloc(sizeof(double) * n3);
    m->xp1 = (double *)malloc(sizeof(double) * n3);
    m->fp = (double *)calloc(n3, sizeof(double));
    m->dis = (double *)calloc(n3, sizeof(double));
    m->den = (double *)calloc(n, sizeof(double));
    m->epot = (double *)calloc(n, sizeof(double));
    m->ekin = (double *)calloc(n, sizeof(double));
    m->ityp = (int *)malloc(sizeof(int) * n);
    m->statu = (int *)malloc(sizeof(int) * n);
    m->gid = (int *)malloc(sizeof(int) * n);
    m->kvois = (int *)calloc(n, sizeof(int));
    m->indi = (int *)malloc(sizeof(int) * (size_t)n * mxkvois);
    m->oxp = (double *)malloc(sizeof(double) * n3);
    m->tmp = (double *)malloc(sizeof(double) * n3);
    m->oityp = (int *)malloc(sizeof(int) * n);
    m->ostatu = (int *)malloc(sizeof(int) * n);
    m->inc = (int *)malloc(sizeof(int) * n);
    m->ngid = (int *)malloc(sizeof(int) * n);
    m->itmp = (int *)malloc(sizeof(int) * n);
    memcpy(m->xp, xp, sizeof(double) * n3);
    memcpy(m->xp1, xp1, sizeof(double) * n3);
    memcpy(m->ityp, ityp, sizeof(int) * n);
    memcpy(m->statu, statu, sizeof(int) * n);
    for (int i = 0; i < n; i++) m->gid[i] = i + 1;
    double rmmax = 0.0;
    for (int i = 0; i < ng * ng; i++)
        if (nb_rm[i] > rmmax) rmmax = nb_rm[i];
    orc_ncell(zl, rmmax, m->ncell);
    size_t nc = (size_t)m->ncell[0] * m->ncell[1] * m->ncell[2] * nbox;
    m->nac = (int *)malloc(sizeof(int) * nc);
    m->naac = (int *)malloc(sizeof(int) * nc);
    m->ia1th = (int *)malloc(sizeof(int) * nc);
    return m;
}

void orc_md_destroy(orc_md *m)
{
    if (!m) return;
    free(m->xp); free(m->xp1); free(m->fp); free(m->dis); free(m->den); free(m->epot); free(m->ekin);
    free(m->ityp); free(m->statu); free(m->gid); free(m->kvois); free(m->indi);
    free(m->oxp); free(m->tmp); free(m->oityp); free(m->ostatu); free(m->inc); free(m->ngid); free(m->itmp);
    free(m->nac); free(m->naac); free(m->ia1th);
    free(m);
}

void orc_md_set_epc(orc_md *m, const int *enable, const double *te, const double *alpha, const double *cut,
                    const double *he)
{
    m->epc_on = 0;
    for (int g = 0; g < m->ng; g++) {
        m->epc_enable[g] = enable[g];
        m->epc_te[g] = te[g];
        m->epc_alpha[g] = alpha[g];
        m->epc_cut[g] = cut[g];
        m->epc_he[g] = he[g];
        if (enable[g] > 0) m->epc_on = 1;
    }
}

/* permute a (n,ld) column-major array: dst[new] = src[perm[new]] */
static void permute_d(double *a, double *tmp, const int *perm, int n, int ncol)
{
    for (int c = 0; c < ncol; c++) {
        for (int i = 0; i < n; i++) tmp[i] = a[perm[i] + (size_t)c * n];
        memcpy(a + (size_t)c * n, tmp, sizeof(double) * n);
    }
}

int orc_md_rebuild(orc_md *m)
{
    /* Cal_NeighBoreList2C_0_DEV, MD_NeighborsList_GPU.F90:1421-1695: un-permute to the
     * original order, re-bin, re-sort every per-atom array (FP is not re-sorted by the
     * reference because it is recomputed right after; we carry it so that getters stay valid). */
    int n = m->n;
    for (int s = 0; s < n; s++) {
        int o = m->gid[s] - 1;
        for (int d = 0; d < 3; d++) m->oxp[o + (size_t)d * n] = m->xp[s + (size_t)d * n];
        m->oityp[o] = m->ityp[s];
        m->ostatu[o] = m->statu[s];
    }
    int nnmax = 0;
    int nout = orc_nlist_build_dev(m->nbox, m->napb, m->oxp, m->oityp, m->ostatu, m->boxlow, m->zl, m->ifpd,
                                   m->bs, m->ng, m->nb_rm, m->mxkvois, m->ncell, m->inc, m->ngid, m->nac,
                                   m->naac, m->ia1th, m->kvois, m->indi, &nnmax);
    /* perm[new sorted pos] = old sorted pos */
    int *ginv = m->inc; /* reuse: original id -> old sorted pos */
    for (int s = 0; s < n; s++) ginv[m->gid[s] - 1] = s;
    int *perm = m->itmp;
    for (int s = 0; s < n; s++) perm[s] = ginv[m->ngid[s] - 1];
    permute_d(m->xp, m->tmp, perm, n, 3);
    permute_d(m->xp1, m->tmp, perm, n, 3);
    permute_d(m->fp, m->tmp, perm, n, 3);
    permute_d(m->dis, m->tmp, perm, n, 3);
    permute_d(m->epot, m->tmp, perm, n, 1);
    permute_d(m->ekin, m->tmp, perm, n, 1);
    for (int s = 0; s < n; s++) {
        m->ityp[s] = m->oityp[m->ngid[s] - 1];
        m->statu[s] = m->ostatu[m->ngid[s] - 1];
    }
    memcpy(m->gid, m->ngid, sizeof(int) * n);
    m->haspart++;
    (void)nnmax;
    return nout;
}

void orc_md_force(orc_md *m, int with_virial)
{
    int n = m->n;
    orc_force_pass1(n, 0, n, m->xp, m->ityp, m->statu, m->kvois, m->indi, n, m->zl, m->ifpd, m->bs, &m->t, m->den);
    orc_force_pass2(n, 0, n, m->xp, m->ityp, m->statu, m->kvois, m->indi, n, m->zl, m->ifpd, m->bs, &m->t, m->den,
                    m->fp, n, with_virial ? m->vtensor : NULL);
    if (with_virial) /* COPYOUT_VIRIALTENSOR, MD_EAM_ForceTable_GPU.F90:1462-1463 */
        for (int i = 0; i < 9; i++) m->vtensor[i] = m->vtensor[i] / (double)(m->n / m->napb);
}

/* pCalAVStress on the driver's state: density pass, then the per-atom tensor; ap[n*9] in ORIGINAL order, column-major */
void orc_md_avstress(orc_md *m, double *ap)
{
    int n = m->n;
    double *tmp = (double *)malloc(sizeof(double) * 9 * (size_t)n);
    orc_force_pass1(n, 0, n, m->xp, m->ityp, m->statu, m->kvois, m->indi, n, m->zl, m->ifpd, m->bs, &m->t, m->den);
    orc_force_avstress(n, 0, n, m->xp, m->ityp, m->statu, m->kvois, m->indi, n, m->zl, m->ifpd, m->bs, &m->t, m->den, tmp, n);
    for (int s = 0; s < n; s++)
        for (int q = 0; q < 9; q++) ap[(m->gid[s] - 1) + (size_t)q * n] = tmp[s + (size_t)q * n];
    free(tmp);
}

/* Cal_NearestNeighbor_Kernel, CommonGPU/MD_NeighborsList_GPU.F90:1805-1946: insertion of every listed neighbour into
 * a distance-ordered buffer of at most NEAREST entries (strict '<': ties keep list order), written back in place. */
void orc_md_reorder_nearest(orc_md *m, int nearest)
{
    const int n = m->n;
    double *r2s = (double *)malloc(sizeof(double) * (nearest + 1));
    int *ns = (int *)malloc(sizeof(int) * (nearest + 1));
    for (int ic = 0; ic < n; ic++) {
        int nn = 0;
        for (int k = 0; k <= nearest; k++) r2s[k] = 1.0e32;
        for (int iw = 0; iw < m->kvois[ic]; iw++) {
            const int j = m->indi[ic + (size_t)iw * n];
            double s[3] = {m->xp[ic] - m->xp[j - 1], m->xp[ic + n] - m->xp[j - 1 + n], m->xp[ic + 2 * n] - m->xp[j - 1 + 2 * n]};
            min_image(s, m->zl, m->ifpd);
            const double r2 = s[0] * s[0] + s[1] * s[1] + s[2] * s[2];
            const int n1 = (nn + 1 < nearest) ? nn + 1 : nearest;
            if (r2 < r2s[n1 - 1]) {
                int i;
                for (i = 1; i <= n1; i++) if (r2 < r2s[i - 1]) break;
                for (int k = n1; k >= i + 1; k--) { r2s[k - 1] = r2s[k - 2]; ns[k - 1] = ns[k - 2]; }
                r2s[i - 1] = r2; ns[i - 1] = j;
                nn = nn + 1;
                if (nn > nearest) nn = nearest;
            }
        }
        m->kvois[ic] = nn;
        for (int k = 0; k < nn; k++) m->indi[ic + (size_t)k * n] = ns[k];
    }
    free(r2s); free(ns);
}

void orc_md_epot(orc_md *m)
{
    int n = m->n;
    orc_force_epot(n, 0, n, m->xp, m->ityp, m->statu, m->kvois, m->indi, n, m->zl, m->ifpd, m->bs, &m->t, m->epot);
}

/* DAMPING_KERNEL, CommonGPU/MD_DiffScheme_GPU.F90:125-185: a velocity component that opposes its force component is
 * zeroed, as is any component whose position is fixed.  (All atoms of the range, no ACTIVE test, as in the reference.) */
void orc_damping(int n, double *xp1, const double *fp, const int *statu)
{
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            double v = xp1[i + (size_t)d * n];
            if (v * fp[i + (size_t)d * n] < 0.0) v = 0.0;
            if ((statu[i] & (ORC_STATU_FIXPOSX << d)) == (ORC_STATU_FIXPOSX << d)) v = 0.0;
            xp1[i + (size_t)d * n] = v;
        }
}
void orc_md_damping(orc_md *m) { orc_damping(m->n, m->xp1, m->fp, m->statu); }

/* Do_DynDamp_Forsteps_DEV, CommonGPU/MD_DiffScheme_GPU.F90:1809-1860.  Returns ITER at exit, 0 when out of steps. */
int orc_md_dyndamp(orc_md *m, int mxnumsteps, double h, double minepot, double *delepot_out)
{
    const int n = m->n;
    double *epot0 = (double *)malloc(sizeof(double) * n), delepot = 0.0;
    int iflag = 0;
    orc_md_epot(m);
    memcpy(epot0, m->epot, sizeof(double) * n);
    for (int iter = 1; iter <= mxnumsteps; iter++) {
        orc_damping(n, m->xp1, m->fp, m->statu);
        orc_predictor(n, m->xp, m->xp1, m->fp, m->dis, m->statu, m->ityp, m->cm, h, m->boxlow, m->boxup, m->zl, m->ifpd);
        orc_md_force(m, 0);
        orc_md_epot(m);
        delepot = 0.0;
        for (int i = 0; i < n; i++) { const double v = fabs(m->epot[i] - epot0[i]); if (v > delepot) delepot = v; }
        if (delepot <= minepot) { iflag = iter; break; }
        memcpy(epot0, m->epot, sizeof(double) * n);
        orc_corrector(n, m->xp1, m->fp, m->statu, m->ityp, m->cm, h);
    }
    if (delepot_out) *delepot_out = delepot;
    free(epot0);
    return iflag;
}

int orc_md_step(orc_md *m, int itime, int it0, int nb_uptab, double h)
{
    /* For_One_Step, Appshell/MD_Method_GenericMD_GPU.F90:596-627 */
    int n = m->n, rebuilt = 0;
    orc_predictor(n, m->xp, m->xp1, m->fp, m->dis, m->statu, m->ityp, m->cm, h, m->boxlow, m->boxup, m->zl, m->ifpd);
    if ((itime - it0) % nb_uptab == 0) { /* Fortran MOD keeps the sign, as C % does */
        orc_md_rebuild(m);
        rebuilt = 1;
    }
    orc_md_force(m, 0);
    if (m->epc_on)
        orc_epc(n, m->xp1, m->fp, m->statu, m->ityp, m->ng, m->epc_enable, m->cm, m->epc_te, m->epc_alpha,
                m->epc_cut, m->epc_he);
    orc_corrector(n, m->xp1, m->fp, m->statu, m->ityp, m->cm, h);
    return rebuilt;
}

void orc_md_get(orc_md *m, double *xp, double *xp1, double *fp, double *epot, double *ekin, double *dis, int *statu,
                int *gid, double *vtensor)
{
    int n = m->n;
    if (ekin) orc_ekin(n, m->xp1, m->statu, m->ityp, m->cm, m->ekin);
    for (int s = 0; s < n; s++) {
        int o = m->gid[s] - 1;
        for (int d = 0; d < 3; d++) {
            if (xp) xp[o + (size_t)d * n] = m->xp[s + (size_t)d * n];
            if (xp1) xp1[o + (size_t)d * n] = m->xp1[s + (size_t)d * n];
            if (fp) fp[o + (size_t)d * n] = m->fp[s + (size_t)d * n];
            if (dis) dis[o + (size_t)d * n] = m->dis[s + (size_t)d * n];
        }
        if (epot) epot[o] = m->epot[s];
        if (ekin) ekin[o] = m->ekin[s];
        if (statu) statu[o] = m->statu[s];
        if (gid) gid[s] = m->gid[s];
    }
    if (vtensor) memcpy(vtensor, m->vtensor, sizeof(double) * 9);
}

/* ------------------------------------------------------------------------------------
 * Quench by steepest descent with a Barzilai-Borwein step: Do_Steepest0_Forsteps_DEV,
 * CommonGPU/MD_SteepestScheme_GPU.F90:20-153 (the "ST" QUICKDAMP scheme of For_One_Step,
 * Appshell/MD_Method_GenericMD_GPU.F90:540-546; PARREP uses it for its event quench).
 * Vector helpers restated from MSMLIB/sor/CommonGPU/MSM_MultiGPU_Basic.F90: DevMultiply,
 * DevMaxAbsval (max |x| over all components of all atoms), DevDot, DevMinus and
 * AddBD_DevVec_DF_KERNEL0 (:5444-5476: RT = V1+V2; RT > HB -> RT - (HB-LB); RT < LB -> RT + (HB-LB)).
 * No neighbour-list rebuild happens inside the loop (the reference does none either).
 * returns IFLAG (0 = ran out of steps, >0 = converged at that iteration, -1 = converged at
 * the first step); *maxmove, *delepot as the reference prints them (cm, erg). */
static double maxabs(const double *a, size_t n)
{
    double m = 0.0;
    for (size_t i = 0; i < n; i++) { double v = fabs(a[i]); if (v > m) m = v; }
    return m;
}
static void add_shift(int n, const double lb[3], const double hb[3], const double *dx, double *x)
{
    for (int d = 0; d < 3; d++)
        for (int i = 0; i < n; i++) {
            double rt = dx[i + (size_t)d * n] + x[i + (size_t)d * n];
            if (rt > hb[d]) rt = rt - (hb[d] - lb[d]);
            else if (rt < lb[d]) rt = rt + (hb[d] - lb[d]);
            x[i + (size_t)d * n] = rt;
        }
}
int orc_md_steepest0(orc_md *m, int mxnumsteps, double alpha0, double maxdis, double mindis, double minepot,
                     double *maxmove_out, double *delepot_out)
{
    const int n = m->n;
    const size_t n3 = (size_t)n * 3;
    double *prefp = (double *)malloc(sizeof(double) * n3), *dxp = (double *)malloc(sizeof(double) * n3);
    double *epot0 = (double *)malloc(sizeof(double) * n);
    double lb[3], hb[3], maxmove = 0.0, delepot = 0.0, alpha = alpha0;
    int iflag = 0;
    for (int d = 0; d < 3; d++) { /* :52-59 */
        lb[d] = m->ifpd[d] ? m->boxlow[d] : -1.0e108;
        hb[d] = m->ifpd[d] ? m->boxup[d] : 1.0e108;
    }
    orc_md_force(m, 0);                                                    /* :70 */
    memcpy(prefp, m->fp, sizeof(double) * n3);
    for (size_t i = 0; i < n3; i++) dxp[i] = alpha * prefp[i];             /* :72 */
    maxmove = maxabs(dxp, n3);
    if (maxmove > maxdis) { const double sc = maxdis / maxmove; for (size_t i = 0; i < n3; i++) dxp[i] = sc * dxp[i]; }
    else if (maxmove <= mindis) { iflag = -1; goto done; }                 /* :76-86 */
    orc_md_epot(m);
    memcpy(epot0, m->epot, sizeof(double) * n);
    add_shift(n, lb, hb, dxp, m->xp);                                      /* :90 */
    for (int it = 1; it <= mxnumsteps; it++) {
        orc_md_force(m, 0);
        double dotdxdf = 0.0, dotdf = 0.0;
        for (size_t i = 0; i < n3; i++) {                                  /* :97-100 */
            const double dfp = prefp[i] - m->fp[i];
            dotdxdf += dxp[i] * dfp;
            dotdf += dfp * dfp;
        }
        alpha = dotdxdf / dotdf;
        if (alpha < 0.0) alpha = alpha0;                                   /* :102-104 */
        for (size_t i = 0; i < n3; i++) dxp[i] = alpha * m->fp[i];
        maxmove = maxabs(dxp, n3);
        if (maxmove > maxdis) { const double sc = maxdis / maxmove; for (size_t i = 0; i < n3; i++) dxp[i] = sc * dxp[i]; }
        add_shift(n, lb, hb, dxp, m->xp);                                  /* :112 */
        if (maxmove <= mindis) { iflag = it; break; }                      /* :117-120 */
        orc_md_epot(m);
        delepot = 0.0;
        for (int i = 0; i < n; i++) { const double v = fabs(m->epot[i] - epot0[i]); if (v > delepot) delepot = v; }
        if (delepot <= minepot) { iflag = it; break; }                     /* :125-128 */
        memcpy(epot0, m->epot, sizeof(double) * n);
        memcpy(prefp, m->fp, sizeof(double) * n3);
    }
done:
    if (maxmove_out) *maxmove_out = maxmove;
    if (delepot_out) *delepot_out = delepot;
    free(prefp); free(dxp); free(epot0);
    return iflag;
}

/* ---- the other quench schemes of the PARREP / GMD "QUICKDAMP" section.
 * Vector helpers follow MSMLIB/sor/CommonGPU/MSM_MultiGPU_Basic.F90: Minus(V1,V2): V2 = V1 - V2 (:5246-5270),
 * Multiply(a,V): V = a*V, Add(V1,V2): V2 = V1 + V2, Dot over all 3N components, MaxAbsval. */
static double vdot(const double *a, const double *b, size_t n) { double s = 0.0; for (size_t i = 0; i < n; i++) s += a[i] * b[i]; return s; }
static void quench_bounds(const orc_md *m, double lb[3], double hb[3])
{
    for (int d = 0; d < 3; d++) {
        lb[d] = m->ifpd[d] ? m->boxlow[d] : -1.0e108;
        hb[d] = m->ifpd[d] ? m->boxup[d] : 1.0e108;
    }
}
static double max_depot(const orc_md *m, const double *epot0)
{
    double v = 0.0;
    for (int i = 0; i < m->n; i++) { const double t = fabs(m->epot[i] - epot0[i]); if (t > v) v = t; }
    return v;
}
/* Do_CG0_Forsteps_DEV (lsearch = 0, CommonGPU/MD_CGScheme_GPU.F90:16-133) and Do_CG1_Forsteps_DEV (lsearch = 1, :137-276):
 * Polak-Ribiere conjugate gradient; the step along a direction comes from one secant estimate (CG0) or from repeated
 * secant estimates until |STEPSIZE| <= MINDIS (CG1).  Returns the reference's ITER at exit if a stop criterion fired
 * (DELEPOT <= MINEPOT or F0NORM <= EPS), 0 if the step budget ran out, -1 if F0NORM <= EPS before the first step.
 * Restated as written, including its mixed units: the secant STEPSIZE (a coefficient on the un-normalised direction)
 * is compared with MAXDIS (a length) and then divided by sqrt(F0NORM) a second time, so in CGS units the cap nearly
 * always binds and the second move of CG0 is +/- the trial move (a backward one returns the atoms to where they were,
 * DELEPOT becomes ~0 and the loop ends). */
int orc_md_cg(orc_md *m, int mxnumsteps, int lsearch, double maxdis, double mindis, double minepot, double *delepot_out)
{
    const int n = m->n;
    const size_t n3 = (size_t)n * 3;
    const double eps = 1.0e-64;
    double *dir = (double *)malloc(sizeof(double) * n3), *f0 = (double *)malloc(sizeof(double) * n3);
    double *dxp = (double *)malloc(sizeof(double) * n3), *epot0 = (double *)malloc(sizeof(double) * n);
    double lb[3], hb[3], delepot = 0.0, f0norm, pf0, pf1, stepsize;
    int iflag = 0, iter;
    quench_bounds(m, lb, hb);
    orc_md_force(m, 0);
    orc_md_epot(m);
    memcpy(epot0, m->epot, sizeof(double) * n);
    memcpy(dir, m->fp, sizeof(double) * n3);
    memcpy(f0, m->fp, sizeof(double) * n3);
    f0norm = vdot(f0, f0, n3);
    pf0 = vdot(m->fp, dir, n3);
    if (f0norm <= eps) { iflag = -1; goto done; }
    iter = lsearch ? 0 : 1;
    while (iter <= mxnumsteps) {
        stepsize = maxdis / sqrt(f0norm);
        for (size_t i = 0; i < n3; i++) dxp[i] = stepsize * dir[i];
        add_shift(n, lb, hb, dxp, m->xp);
        orc_md_force(m, 0);
        pf1 = vdot(m->fp, dir, n3);
        if (!lsearch) {
            stepsize = -stepsize * pf1 / (pf1 - pf0);
            if (fabs(stepsize) > maxdis) stepsize = maxdis * fabs(stepsize) / stepsize;
            { const double c = stepsize / sqrt(f0norm); for (size_t i = 0; i < n3; i++) dxp[i] = c * dir[i]; }
            add_shift(n, lb, hb, dxp, m->xp);
        } else {
            iter++;
            while (iter <= mxnumsteps) {
                stepsize = -stepsize * pf1 / (pf1 - pf0);
                if (fabs(stepsize) > maxdis) stepsize = maxdis * fabs(stepsize) / stepsize;
                { const double c = stepsize / sqrt(f0norm); for (size_t i = 0; i < n3; i++) dxp[i] = c * dir[i]; }
                add_shift(n, lb, hb, dxp, m->xp);
                iter++;
                if (fabs(stepsize) <= mindis) break;
                pf0 = pf1;
                orc_md_force(m, 0);
                pf1 = vdot(m->fp, dir, n3);
            }
        }
        orc_md_epot(m);
        delepot = max_depot(m, epot0);
        if (delepot <= minepot) { iflag = iter; break; }
        memcpy(epot0, m->epot, sizeof(double) * n);
        orc_md_force(m, 0);
        for (size_t i = 0; i < n3; i++) f0[i] = m->fp[i] - f0[i];
        {
            const double mf1norm = vdot(m->fp, f0, n3), gama = mf1norm / f0norm;
            for (size_t i = 0; i < n3; i++) dir[i] = gama * dir[i];
            for (size_t i = 0; i < n3; i++) dir[i] = m->fp[i] + dir[i];
        }
        memcpy(f0, m->fp, sizeof(double) * n3);
        f0norm = vdot(f0, f0, n3);
        pf0 = vdot(m->fp, dir, n3);
        if (f0norm <= eps) { iflag = iter; break; }
        if (!lsearch) iter++;
    }
done:
    if (delepot_out) *delepot_out = delepot;
    free(dir); free(f0); free(dxp); free(epot0);
    return iflag;
}
/* Do_Steepest1_Forsteps_DEV (CommonGPU/MD_SteepestScheme_GPU.F90:157-260): steepest descent along the normalised force
 * with the same repeated-secant line search; MINEPOT is the literal 0.001 eV of :178.  Returns ITER at exit when the
 * energy criterion fired, 0 when the step budget ran out. */
int orc_md_steepest1(orc_md *m, int mxnumsteps, double maxdis, double mindis, double *delepot_out)
{
    const int n = m->n;
    const size_t n3 = (size_t)n * 3;
    const double minepot = 0.001 * 1.60219e-12;
    double *dir = (double *)malloc(sizeof(double) * n3), *dxp = (double *)malloc(sizeof(double) * n3);
    double *epot0 = (double *)malloc(sizeof(double) * n);
    double lb[3], hb[3], delepot = 0.0, pf0, pf1, stepsize;
    int iter = 0, iflag = 0;
    quench_bounds(m, lb, hb);
    while (iter <= mxnumsteps) {
        orc_md_force(m, 0);
        orc_md_epot(m);
        memcpy(epot0, m->epot, sizeof(double) * n);
        { /* DevNormalize: V = (1/dsqrt(V.V)) * V (CommonGPU/MD_Globle_Variables_GPU.F90:2775-2790) */
            const double sc = 1.0 / sqrt(vdot(m->fp, m->fp, n3));
            for (size_t i = 0; i < n3; i++) dir[i] = sc * m->fp[i];
        }
        stepsize = maxdis;
        for (size_t i = 0; i < n3; i++) dxp[i] = stepsize * dir[i];
        pf0 = vdot(m->fp, dir, n3);
        while (iter <= mxnumsteps) {
            add_shift(n, lb, hb, dxp, m->xp);
            orc_md_force(m, 0);
            pf1 = vdot(m->fp, dir, n3);
            stepsize = -stepsize * pf1 / (pf1 - pf0);
            if (fabs(stepsize) > maxdis) stepsize = maxdis * fabs(stepsize) / stepsize;
            for (size_t i = 0; i < n3; i++) dxp[i] = stepsize * dir[i];
            iter++;
            if (fabs(stepsize) <= mindis) break;
            pf0 = pf1;
        }
        orc_md_epot(m);
        delepot = max_depot(m, epot0);
        if (delepot <= minepot) { iflag = iter; break; }
    }
    if (delepot_out) *delepot_out = delepot;
    free(dir); free(dxp); free(epot0);
    return iflag;
}

/* ------------------------------------------------------------------------------------
 * Remaining procedures of the integrator module the step loop calls (MD_DiffScheme_GPU.F90):
 * Cal_GlobalT_DEV :1042-1064, CheckTimestep_KERNEL/_DEV :1066-1258, VelScaling_KERNEL/_DEV :1262-1446.
 * Arrays in any common order; gid[s] = 1-based original id of sorted atom s (box of an atom for the
 * per-box kinetic energy = (original id)/NPRT as the reference's hm_GIDINV slices; box of an atom
 * for the scaling factor = (sorted index)/NPRT as VelScaling_KERNEL's IB0). */
double orc_global_t(int n, const double *xp1, const int *statu, const int *ityp, const double *cm)
{
    double *ek = (double *)malloc(sizeof(double) * n), sum = 0.0;
    long cnt = 0;
    orc_ekin(n, xp1, statu, ityp, cm, ek);
    for (int i = 0; i < n; i++) if (ek[i] >= 0.0) { sum += ek[i]; cnt++; }
    free(ek);
    return 2.0 * sum / (double)cnt / (3.0 * 1.38054e-16); /* C_TWO*sum/count/(C_THR*CP_KB) :1062 */
}
int orc_check_timestep(int n, const double *xp1, const double *fp, const int *statu, const int *ityp, const double *cm,
                       double th, double h2s2, double mxd2)
{
    for (int i = 0; i < n; i++) {
        if ((statu[i] & 1) != 1) continue;
        const double cm0 = cm[ityp[i] - 1];
        double d[3] = {0.0, 0.0, 0.0};
        for (int k = 0; k < 3; k++)
            if ((statu[i] & (2 << k)) == 0) d[k] = th * xp1[i + (size_t)k * n] + h2s2 * (fp[i + (size_t)k * n] / cm0);
        if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] > mxd2) return 1;
    }
    return 0;
}
int orc_vel_scaling(int n, int napb, double *xp1, const int *statu, const int *ityp, const int *gid, const double *cm, double dt)
{
    const int nbox = n / napb;
    double *ek = (double *)malloc(sizeof(double) * n), *sum = (double *)calloc(nbox, sizeof(double));
    long *cnt = (long *)calloc(nbox, sizeof(long));
    int *pos_of = (int *)malloc(sizeof(int) * n);
    orc_ekin(n, xp1, statu, ityp, cm, ek);
    for (int s = 0; s < n; s++) pos_of[gid[s] - 1] = s;
    for (int o = 0; o < n; o++) { /* original order, as sum(hm_EKIN(hm_GIDINV(IPA+1:IPA+NPRT))) :1418-1422 */
        const double e = ek[pos_of[o]];
        if (e >= 0.0) { sum[o / napb] += e; cnt[o / napb]++; }
    }
    int bad = 0;
    for (int b = 0; b < nbox; b++) {
        const double ce = sum[b] / (double)cnt[b];
        if (!(ce > 0.0)) bad = 1;
        sum[b] = sqrt(dt * 3.0 * 1.38054e-16 * 0.5 / ce); /* dsqrt(DT*C_THR*CP_KB*C_HALF/cEKIN) :1432 */
    }
    if (!bad)
        for (int s = 0; s < n; s++) {
            if ((statu[s] & 1) != 1) continue;
            const double sc = sum[s / napb];
            for (int k = 0; k < 3; k++) {
                const int free_k = (statu[s] & (2 << k)) == 0 && (statu[s] & (16 << k)) == 0;
                xp1[s + (size_t)k * n] = free_k ? xp1[s + (size_t)k * n] * sc : 0.0; /* fixed components are zeroed :1299-1315 */
            }
        }
    free(ek); free(sum); free(cnt); free(pos_of);
    return bad ? -1 : 0;
}
double orc_md_global_t(orc_md *m) { return orc_global_t(m->n, m->xp1, m->statu, m->ityp, m->cm); }
int orc_md_vel_scaling(orc_md *m, double dt) { return orc_vel_scaling(m->n, m->napb, m->xp1, m->statu, m->ityp, m->gid, m->cm, dt); }
int orc_md_check_timestep(orc_md *m, double th, double h2s2, double mxd2)
{
    return orc_check_timestep(m->n, m->xp1, m->fp, m->statu, m->ityp, m->cm, th, h2s2, mxd2);
}

int orc_md_natom(orc_md *m) { return m->n; }
const int *orc_md_kvois(orc_md *m) { return m->kvois; }
const int *orc_md_indi(orc_md *m) { return m->indi; }
void orc_md_ncell(orc_md *m, int ncell[3]) { for (int d = 0; d < 3; d++) ncell[d] = m->ncell[d]; }


/* ------------------------------------------------------------------------------------
 * Thermalizing_MC_KERNEL / Thermalizing_MC_DEV, CommonGPU/MD_DiffScheme_GPU.F90:1608-1805, with the product's
 * documented substitution of the random source: uniforms from Philox4x32-10 (Salmon et al., SC'11) keyed by the
 * seed, counter = (original 1-based atom id, draw, block 0|1, 0), Z = (x + 1/2) 2^-32.  Everything after the
 * uniforms is the reference's arithmetic: V0*DSQRT(-DLOG(Z1))*DCOS(CP_TWOPI*Z2), zero for inactive atoms,
 * then per box WT = sum m, VT = sum m*v/WT over the box in original order, v -= VT for every atom. */
void orc_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4])
{
    unsigned c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void orc_md_thermalize(orc_md *m, double ti, unsigned long long seed, unsigned draw)
{
    const int n = m->n, napb = m->napb;
    const unsigned key[2] = {(unsigned)seed, (unsigned)(seed >> 32)};
    for (int s = 0; s < n; s++) {
        const int st = m->statu[s];
        if ((st & ORC_STATU_ACTIVE) != ORC_STATU_ACTIVE) {
            for (int d = 0; d < 3; d++) m->xp1[s + (size_t)d * n] = 0.0;
            continue;
        }
        unsigned z[8];
        const unsigned c0[4] = {(unsigned)m->gid[s], draw, 0u, 0u}, c1[4] = {(unsigned)m->gid[s], draw, 1u, 0u};
        orc_philox4x32_10(c0, key, z);
        orc_philox4x32_10(c1, key, z + 4);
        const double v0 = sqrt(2.0 * ti * ORC_KB / m->cm[m->ityp[s] - 1]);
        for (int d = 0; d < 3; d++) {
            if ((st & (ORC_STATU_FIXVELX << d)) == 0 && (st & (ORC_STATU_FIXPOSX << d)) == 0) {
                const double z1 = ((double)z[2 * d] + 0.5) * 2.3283064365386963e-10, z2 = ((double)z[2 * d + 1] + 0.5) * 2.3283064365386963e-10;
                m->xp1[s + (size_t)d * n] = v0 * sqrt(-log(z1)) * cos(6.283185307179586 * z2);
            }
        }
    }
    int *pos_of = (int *)malloc(sizeof(int) * n);
    for (int s = 0; s < n; s++) pos_of[m->gid[s] - 1] = s;
    for (int b = 0; b < m->nbox; b++) { /* :1782-1798 */
        double wt = 0.0, vt[3] = {0.0, 0.0, 0.0};
        for (int o = 0; o < napb; o++) wt += m->cm[m->ityp[pos_of[b * napb + o]] - 1];
        for (int o = 0; o < napb; o++) {
            const int s = pos_of[b * napb + o];
            for (int d = 0; d < 3; d++) vt[d] += m->cm[m->ityp[s] - 1] * m->xp1[s + (size_t)d * n] / wt;
        }
        for (int o = 0; o < napb; o++) {
            const int s = pos_of[b * napb + o];
            for (int d = 0; d < 3; d++) m->xp1[s + (size_t)d * n] -= vt[d];
        }
    }
    free(pos_of);
}


/* ------------------------------------------------------------------------------------
 * DO_LBFGSB_FORSTEPS_DEV, CommonGPU/MD_LBFGSScheme_GPU.F90:177-388: limited-memory BFGS quench driven through the
 * reverse-communication routine SETULB (LIB/sor/f/LBFGSB/lbfgsb.f, L-BFGS-B of Zhu, Byrd, Lu, Nocedal) with NBD = 0 for
 * every variable, i.e. WITHOUT bounds.  What SETULB does in that case is restated here:
 *   - variables: the free position components (FIXPOS bit clear, not OUTOFBOX, ACTIVE), :213-243; objective = sum of
 *     EPOT over all atoms, gradient = -FP, :332-337; the positions given to the force routine are X wrapped into the
 *     periodic box, X itself stays unwrapped (:268-330);
 *   - mainlb (lbfgsb.f:287-1000) without bounds: the generalized Cauchy point is x (:222 block "if (.not.cnstnd .and.
 *     col.gt.0)") and the subspace minimisation (formk/cmprlb/subsm) returns the quasi-Newton step -H g of the compact
 *     limited-memory matrix with B0 = theta*I, theta = y'y/s'y (matupd :2461); with an empty memory the step is -g/theta
 *     (cauchy with no breakpoints).  That step is computed here by the two-loop recursion, which is the same vector in
 *     exact arithmetic -- NOT the same floating-point operations as formk's LEL^T factorisation, so agreement with a
 *     run of the reference binary is to round-off amplified by the iteration, not bit for bit ("parity unpinned": no
 *     reference output of an L-BFGS quench ships);
 *   - lnsrlb (:2285-2396): first step min(1/|d|, 1e10) at iteration 0, else 1; More'-Thuente search dcsrch/dcstep
 *     (:3280-3530, :3534-3760) with ftol 1e-3, gtol 0.9, xtol 0.1, stpmin 0, stpmax 1e10; 20 evaluations at most;
 *   - tests (:777 block): max|g_i| <= pgtol; (fold - f) <= factr*epsmch*max(|fold|,|f|,1); update skipped when
 *     s'y <= epsmch*(-gdold*stp); memory reset and restart when the search fails with a non-empty memory.
 * Every SETULB call counts against MXNUMSTEPS (:296): one per function evaluation and one per accepted step.
 * Returns IFLAG (0 converged / abnormal line search, 1 out of steps); *nfg = force evaluations, *niter = accepted steps. */
typedef struct { int brackt, stage; double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1; } orc_ls;
static void orc_dcstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy, double *stp, double fp, double dp,
                       int *brackt, double stpmin, double stpmax)
{
    double gamma, p, q, r, s, sgnd, stpc, stpf, stpq, theta;
    sgnd = dp * (*dx / fabs(*dx));
    if (fp > *fx) { /* case 1: higher value -> bracketed; cubic vs quadratic, :3620-3640 */
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp < *stx) gamma = -gamma;
        p = (gamma - *dx) + theta; q = ((gamma - *dx) + gamma) + dp; r = p / q;
        stpc = *stx + r * (*stp - *stx);
        stpq = *stx + ((*dx / ((*fx - fp) / (*stp - *stx) + *dx)) / 2.0) * (*stp - *stx);
        stpf = (fabs(stpc - *stx) < fabs(stpq - *stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
        *brackt = 1;
    } else if (sgnd < 0.0) { /* case 2: derivatives of opposite sign, :3648-3665 */
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp > *stx) gamma = -gamma;
        p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + *dx; r = p / q;
        stpc = *stp + r * (*stx - *stp);
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        stpf = (fabs(stpc - *stp) > fabs(stpq - *stp)) ? stpc : stpq;
        *brackt = 1;
    } else if (fabs(dp) < fabs(*dx)) { /* case 3: derivative decreases in magnitude, :3673-3715 */
        double t;
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
        t = (theta / s) * (theta / s) - (*dx / s) * (dp / s);
        gamma = s * sqrt(t > 0.0 ? t : 0.0);
        if (*stp > *stx) gamma = -gamma;
        p = (gamma - dp) + theta; q = (gamma + (*dx - dp)) + gamma; r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = *stp + r * (*stx - *stp);
        else if (*stp > *stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (*brackt) {
            stpf = (fabs(stpc - *stp) < fabs(stpq - *stp)) ? stpc : stpq;
            if (*stp > *stx) stpf = fmin(*stp + 0.66 * (*sty - *stp), stpf);
            else stpf = fmax(*stp + 0.66 * (*sty - *stp), stpf);
        } else {
            stpf = (fabs(stpc - *stp) > fabs(stpq - *stp)) ? stpc : stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else { /* case 4, :3722-3738 */
        if (*brackt) {
            theta = 3.0 * (fp - *fy) / (*sty - *stp) + *dy + dp;
            s = fmax(fabs(theta), fmax(fabs(*dy), fabs(dp)));
            gamma = s * sqrt((theta / s) * (theta / s) - (*dy / s) * (dp / s));
            if (*stp > *sty) gamma = -gamma;
            p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + *dy; r = p / q;
            stpf = *stp + r * (*sty - *stp);
        } else stpf = (*stp > *stx) ? stpmax : stpmin;
    }
    if (fp > *fx) { *sty = *stp; *fy = fp; *dy = dp; }
    else {
        if (sgnd < 0.0) { *sty = *stx; *fy = *fx; *dy = *dx; }
        *stx = *stp; *fx = fp; *dx = dp;
    }
    *stp = stpf;
}
/* dcsrch: returns 0 = evaluate at *stp, 1 = converged, 2 = warning (search ends), -1 = error */
static int orc_dcsrch(orc_ls *L, int start, double f, double g, double *stp, double ftol, double gtol, double xtol, double stpmin, double stpmax)
{
    if (start) {
        if (*stp < stpmin || *stp > stpmax || g >= 0.0) return -1;
        L->brackt = 0; L->stage = 1; L->finit = f; L->ginit = g; L->gtest = ftol * g;
        L->width = stpmax - stpmin; L->width1 = L->width / 0.5;
        L->stx = 0.0; L->fx = f; L->gx = g; L->sty = 0.0; L->fy = f; L->gy = g;
        L->stmin = 0.0; L->stmax = *stp + 4.0 * *stp;
        return 0;
    }
    {
        const double ftest = L->finit + *stp * L->gtest;
        int warn = 0, conv = 0;
        if (L->stage == 1 && f <= ftest && g >= 0.0) L->stage = 2;
        if (L->brackt && (*stp <= L->stmin || *stp >= L->stmax)) warn = 1;
        if (L->brackt && L->stmax - L->stmin <= xtol * L->stmax) warn = 1;
        if (*stp == stpmax && f <= ftest && g <= L->gtest) warn = 1;
        if (*stp == stpmin && (f > ftest || g >= L->gtest)) warn = 1;
        if (f <= ftest && fabs(g) <= gtol * (-L->ginit)) conv = 1;
        if (conv) return 1;
        if (warn) return 2;
        if (L->stage == 1 && f <= L->fx && f > ftest) { /* modified function, :3438-3456 */
            double fm = f - *stp * L->gtest, fxm = L->fx - L->stx * L->gtest, fym = L->fy - L->sty * L->gtest;
            double gm = g - L->gtest, gxm = L->gx - L->gtest, gym = L->gy - L->gtest;
            orc_dcstep(&L->stx, &fxm, &gxm, &L->sty, &fym, &gym, stp, fm, gm, &L->brackt, L->stmin, L->stmax);
            L->fx = fxm + L->stx * L->gtest; L->fy = fym + L->sty * L->gtest; L->gx = gxm + L->gtest; L->gy = gym + L->gtest;
        } else
            orc_dcstep(&L->stx, &L->fx, &L->gx, &L->sty, &L->fy, &L->gy, stp, f, g, &L->brackt, L->stmin, L->stmax);
        if (L->brackt) {
            if (fabs(L->sty - L->stx) >= 0.66 * L->width1) *stp = L->stx + 0.5 * (L->sty - L->stx);
            L->width1 = L->width; L->width = fabs(L->sty - L->stx);
        }
        if (L->brackt) { L->stmin = fmin(L->stx, L->sty); L->stmax = fmax(L->stx, L->sty); }
        else { L->stmin = *stp + 1.1 * (*stp - L->stx); L->stmax = *stp + 4.0 * (*stp - L->stx); }
        *stp = fmax(*stp, stpmin);
        *stp = fmin(*stp, stpmax);
        if ((L->brackt && (*stp <= L->stmin || *stp >= L->stmax)) || (L->brackt && L->stmax - L->stmin <= xtol * L->stmax)) *stp = L->stx;
    }
    return 0;
}

int orc_md_lbfgsb(orc_md *m, int mxnumsteps, int msave, double factr, double pgtol, int *nfg_out, int *niter_out)
{
    const int n = m->n;
    const size_t n3 = (size_t)n * 3;
    const double epsmch = 2.220446049250313e-16, big = 1.0e10;
    unsigned char *fre = (unsigned char *)malloc(n3);
    double *x = (double *)malloc(sizeof(double) * n3), *g = (double *)malloc(sizeof(double) * n3), *d = (double *)malloc(sizeof(double) * n3);
    double *t = (double *)malloc(sizeof(double) * n3), *r = (double *)malloc(sizeof(double) * n3);
    double *ws = (double *)malloc(sizeof(double) * n3 * msave), *wy = (double *)malloc(sizeof(double) * n3 * msave);
    double *rho = (double *)malloc(sizeof(double) * msave), *alpha = (double *)malloc(sizeof(double) * msave);
    double f = 0.0, fold = 0.0, theta = 1.0, stp = 0.0, gd = 0.0, gdold = 0.0, dtd = 0.0, dnorm = 0.0, sbgnrm = 0.0;
    const double tol = factr * epsmch;
    int col = 0, head = 0, iter = 0, nfg = 0, calls = 0, iflag = 0, ifun = 0, done = 0;
    orc_ls L;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            const int st = m->statu[i];
            fre[i + (size_t)k * n] = ((st & (ORC_STATU_FIXPOSX << k)) == 0 && (st & ORC_STATU_OUTOFBOX) == 0 &&
                                      (st & ORC_STATU_ACTIVE) == ORC_STATU_ACTIVE) ? 1 : 0;
        }
    memcpy(x, m->xp, sizeof(double) * n3);
#define LB_EVAL()                                                                                                          \
    do { /* positions = X wrapped into the box (:268-330), force + energy, G = -FP on the free components (:332-337) */     \
        for (int i_ = 0; i_ < n; i_++)                                                                                     \
            for (int k_ = 0; k_ < 3; k_++) {                                                                               \
                const size_t o_ = i_ + (size_t)k_ * n;                                                                     \
                if (!fre[o_]) continue;                                                                                    \
                double v_ = x[o_];                                                                                         \
                if (m->ifpd[k_]) { if (v_ < m->boxlow[k_]) v_ = v_ + m->zl[k_]; else if (v_ > m->boxup[k_]) v_ = v_ - m->zl[k_]; } \
                m->xp[o_] = v_;                                                                                            \
            }                                                                                                              \
        orc_md_force(m, 0); orc_md_epot(m);                                                                                \
        f = 0.0; for (int i_ = 0; i_ < n; i_++) f += m->epot[i_];                                                          \
        for (size_t o_ = 0; o_ < n3; o_++) g[o_] = fre[o_] ? -m->fp[o_] : 0.0;                                            \
        nfg++;                                                                                                             \
    } while (0)
    /* call 1: 'START' -> 'FG_START' */
    calls = 1;
    if (calls > mxnumsteps) { iflag = 1; goto out; }
    LB_EVAL();
    sbgnrm = 0.0; for (size_t o = 0; o < n3; o++) if (fabs(g[o]) > sbgnrm) sbgnrm = fabs(g[o]);
    if (sbgnrm <= pgtol) goto out;
    while (!done) {
        /* ---- direction: -H g by the two-loop recursion, H0 = I/theta */
        memcpy(d, g, sizeof(double) * n3);
        for (int j = col - 1; j >= 0; j--) {
            const int p = (head + j) % msave;
            alpha[j] = rho[p] * vdot(ws + (size_t)p * n3, d, n3);
            for (size_t o = 0; o < n3; o++) d[o] = d[o] - alpha[j] * wy[(size_t)p * n3 + o];
        }
        for (size_t o = 0; o < n3; o++) d[o] = d[o] / theta;
        for (int j = 0; j < col; j++) {
            const int p = (head + j) % msave;
            const double beta = rho[p] * vdot(wy + (size_t)p * n3, d, n3);
            for (size_t o = 0; o < n3; o++) d[o] = d[o] + ws[(size_t)p * n3 + o] * (alpha[j] - beta);
        }
        for (size_t o = 0; o < n3; o++) d[o] = -d[o];
        /* ---- lnsrlb :2313-2350 */
        dtd = vdot(d, d, n3); dnorm = sqrt(dtd);
        stp = (iter == 0) ? fmin(1.0 / dnorm, big) : 1.0;
        memcpy(t, x, sizeof(double) * n3); memcpy(r, g, sizeof(double) * n3);
        fold = f; ifun = 0;
        {
            int start = 1, failed = 0;
            for (;;) {
                int rc;
                gd = vdot(g, d, n3);
                if (ifun == 0) { gdold = gd; if (gd >= 0.0) { failed = 1; break; } } /* info = -4 */
                rc = orc_dcsrch(&L, start, f, gd, &stp, 1.0e-3, 0.9, 0.1, 0.0, big);
                start = 0;
                if (rc != 0) break;                       /* CONVERGENCE or WARNING: task = NEW_X */
                ifun++;
                if (ifun - 1 >= 20) { failed = 1; break; } /* iback >= 20 (:906) */
                for (size_t o = 0; o < n3; o++) x[o] = stp * d[o] + t[o];
                calls++;
                if (calls > mxnumsteps) { iflag = 1; goto out; }
                LB_EVAL();
            }
            if (failed) { /* :906-935: restore, then either give up (empty memory) or refresh the memory and restart */
                memcpy(x, t, sizeof(double) * n3); memcpy(g, r, sizeof(double) * n3); f = fold;
                if (col == 0) { iter++; done = 1; break; }  /* ABNORMAL_TERMINATION_IN_LNSRCH */
                col = 0; head = 0; theta = 1.0;
                continue;
            }
        }
        /* ---- NEW_X returned: one more call */
        iter++;
        calls++;
        sbgnrm = 0.0; for (size_t o = 0; o < n3; o++) if (fabs(g[o]) > sbgnrm) sbgnrm = fabs(g[o]);
        if (calls > mxnumsteps || calls + 1 > mxnumsteps) { iflag = 1; goto out; } /* the tests below belong to the NEXT call */
        if (sbgnrm <= pgtol) break;
        { const double dd = fmax(fabs(fold), fmax(fabs(f), 1.0)); if ((fold - f) <= tol * dd) break; }
        /* ---- update :822-862 */
        {
            double rr, dr, ddum;
            for (size_t o = 0; o < n3; o++) r[o] = g[o] - r[o];
            rr = vdot(r, r, n3);
            if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
            else { dr = (gd - gdold) * stp; for (size_t o = 0; o < n3; o++) d[o] = stp * d[o]; ddum = -gdold * stp; }
            if (dr > epsmch * ddum) {
                int slot;
                if (col < msave) { slot = (head + col) % msave; col++; }
                else { slot = head; head = (head + 1) % msave; }
                memcpy(ws + (size_t)slot * n3, d, sizeof(double) * n3);
                memcpy(wy + (size_t)slot * n3, r, sizeof(double) * n3);
                rho[slot] = 1.0 / dr;
                theta = rr / dr;
            }
        }
    }
out:
    /* the last evaluated configuration stays on the device state; velocities are zeroed (:363-364) */
    for (size_t o = 0; o < n3; o++) m->xp1[o] = 0.0;
    if (nfg_out) *nfg_out = nfg;
    if (niter_out) *niter_out = iter;
#undef LB_EVAL
    free(fre); free(x); free(g); free(d); free(t); free(r); free(ws); free(wy); free(rho); free(alpha);
    return iflag;
}


/* =====================================================================================
 * The reference's CPU path -- what bench.py times as the CPU baseline.
 *   lists   Cal_NeighboreList2C (every directed pair)        Common/MD_NeighborsList.F90:396-633  orc_nlist_build_cpu
 *           Cal_NeighboreListC  (Newton's third law)          :152-391                             orc_nlist_build_cpu_half
 *   forces  CALFORCE_FS_Force_Table2 (preCALFOR2 / CALFOR2)   Common/MD_FS_ForceTable.F90:400-640  orc_force_pass1/2
 *           CALFORCE_FS_Force_Table  (preCALFOR / CALFOR)     :150-395                             orc_force_newton3
 *   step    Predictor / Correction (Swope form of velocity Verlet)  Common/MD_SwopeScheme.F90:22-283: the arithmetic of
 *           the device scheme (h*v + h^2/2*F/m, wrap, half kicks), restated once as orc_predictor / orc_corrector
 * The reference has no CPU EAM force (Common/MD_ForceClass_Register.F90:155-160 refuses EAM_TYPE on the CPU): as SURVEY.md
 * 8(d) prescribes, the FS code path is generalised with the embedding lookup of the device kernels (the FS form -1/2/sqrt(rho)
 * is kept for ORC_POT_FS).  The CPU path never sorts atoms: state stays in the ORIGINAL order.  The in-range test uses
 * maxval(RU^2) as the device kernels do (the CPU twin tests m_RCUT2 per pair of types: identical for one atom type).
 * ===================================================================================== */
static double embed_df(const orc_tables *t, int ti, double rho)
{
    if (t->pot_type == ORC_POT_FS) return rho > 0.0 ? -0.5 / sqrt(rho) : 0.0;
    if (!(rho > 0.0)) return 0.0;
    int ktab = t->kembd[ti - 1];
    double sk = rho / t->rhod + 1.0;
    int kk = (int)(sk + 0.000001);
    double a = tab(t->dfembd, t->nkind1, t->nembd, ktab, kk);
    return a + (sk - (double)kk) * (tab(t->dfembd, t->nkind1, t->nembd, ktab, kk + 1) - a);
}
static double embed_f(const orc_tables *t, int ti, double rho)
{
    if (t->pot_type == ORC_POT_FS) return rho > 0.0 ? -sqrt(rho) : 0.0; /* EPOT = m_ER - DSQRT(m_DEN) :180 */
    if (!(rho > 0.0)) return 0.0;
    int ktab = t->kembd[ti - 1];
    double sk = rho / t->rhod + 1.0;
    int kk = (int)(sk + 0.000001);
    double a = tab(t->fembd, t->nkind1, t->nembd, ktab, kk);
    return a + (sk - (double)kk) * (tab(t->fembd, t->nkind1, t->nembd, ktab, kk + 1) - a);
}

/* preCALFOR + CALFOR on a HALF list (Common/MD_FS_ForceTable.F90:188-395): every stored pair adds to both atoms.
 * Serial, as the reference runs it.  den: work array (rho, then dF/drho); er: pair energies; epot may be NULL.
 * vtensor (9, column-major) gets sum over pairs of FORTOT*d_a*d_b (:384-388). */
void orc_force_newton3(int n, const double *xp, const int *ityp, const int *kvois, const int *indi, int ldindi,
                       const double zl[3], const int ifpd[3], const double bs[9], const orc_tables *t,
                       double *den, double *er, double *fp, double *epot, double *vtensor)
{
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; i++) { den[i] = 0.0; er[i] = 0.0; }
    for (int i = 0; i < n; i++) { /* preCALFOR :219-266 */
        int ti = ityp[i], iiw = kvois[i];
        for (int iw = 0; iw < iiw; iw++) {
            int j = indi[i + (size_t)iw * ldindi] - 1;
            double s[3] = {xp[i] - xp[j], xp[i + n] - xp[j + n], xp[i + 2 * n] - xp[j + 2 * n]};
            min_image(s, zl, ifpd);
            double dx = bs[0] * s[0] + bs[3] * s[1] + bs[6] * s[2];
            double dy = bs[1] * s[0] + bs[4] * s[1] + bs[7] * s[2];
            double dz = bs[2] * s[0] + bs[5] * s[1] + bs[8] * s[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 <= t->ru2max) {
                int tj = ityp[j];
                int k0 = t->kpair[(ti - 1) + t->ng * (tj - 1)], k1 = t->kpair[(tj - 1) + t->ng * (ti - 1)];
                double r = sqrt(r2);
                double sk = sqrt(r) * t->csi;
                int kk = (int)sk;
                double dk = sk - (double)kk;
                double a = tab(t->potr, t->nkind, t->ntab, k0, kk);
                double exp1 = (a + dk * (tab(t->potr, t->nkind, t->ntab, k0, kk + 1) - a)) / r; /* :256 */
                er[i] = er[i] + exp1;
                er[j] = er[j] + exp1;
                double b = tab(t->potb, t->nkind, t->ntab, k0, kk), c = tab(t->potb, t->nkind, t->ntab, k1, kk);
                den[i] = den[i] + (b + dk * (tab(t->potb, t->nkind, t->ntab, k0, kk + 1) - b)); /* :260 */
                den[j] = den[j] + (c + dk * (tab(t->potb, t->nkind, t->ntab, k1, kk + 1) - c)); /* :261 */
            }
        }
    }
    for (int i = 0; i < n; i++) { /* :269-271 */
        if (epot) epot[i] = er[i] + embed_f(t, ityp[i], den[i]);
        den[i] = embed_df(t, ityp[i], den[i]);
    }
    if (n > 0) memset(fp, 0, sizeof(double) * 3 * (size_t)n); /* FP = C_ZERO :303 */
    for (int i = 0; i < n; i++) { /* CALFOR :305-391 */
        int ti = ityp[i], iiw = kvois[i];
        double denki = den[i];
        for (int iw = 0; iw < iiw; iw++) {
            int j = indi[i + (size_t)iw * ldindi] - 1;
            double s[3] = {xp[i] - xp[j], xp[i + n] - xp[j + n], xp[i + 2 * n] - xp[j + 2 * n]};
            min_image(s, zl, ifpd);
            double dx = bs[0] * s[0] + bs[3] * s[1] + bs[6] * s[2];
            double dy = bs[1] * s[0] + bs[4] * s[1] + bs[7] * s[2];
            double dz = bs[2] * s[0] + bs[5] * s[1] + bs[8] * s[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 <= t->ru2max) {
                int tj = ityp[j];
                int k0 = t->kpair[(ti - 1) + t->ng * (tj - 1)], k1 = t->kpair[(tj - 1) + t->ng * (ti - 1)];
                double r = sqrt(r2);
                double sk = sqrt(r) * t->csi;
                int kk = (int)sk;
                double dk = sk - (double)kk;
                double a = tab(t->fpotr, t->nkind, t->ntab, k0, kk);
                double b = tab(t->fpotb, t->nkind, t->ntab, k0, kk);
                double c = tab(t->fpotb, t->nkind, t->ntab, k1, kk);
                /* :364-367 */
                double fortot = (a + dk * (tab(t->fpotr, t->nkind, t->ntab, k0, kk + 1) - a)) / r2 +
                                ((c + dk * (tab(t->fpotb, t->nkind, t->ntab, k1, kk + 1) - c)) * den[j] +
                                 (b + dk * (tab(t->fpotb, t->nkind, t->ntab, k0, kk + 1) - b)) * denki) / r;
                for (int k = 0; k < 3; k++) { /* :369-373 */
                    double f = fortot * s[k];
                    fp[i + (size_t)k * n] = fp[i + (size_t)k * n] + f;
                    fp[j + (size_t)k * n] = fp[j + (size_t)k * n] - f;
                }
                if (vtensor) {
                    double d[3] = {dx, dy, dz};
                    for (int k = 0; k < 3; k++)
                        for (int k1_ = 0; k1_ < 3; k1_++) v[k + 3 * k1_] = v[k + 3 * k1_] + d[k1_] * d[k] * fortot;
                }
            }
        }
    }
    if (vtensor) memcpy(vtensor, v, sizeof(v));
}

struct orc_cpu {
    int n, ng, mxkvois, half, epc_on;
    double boxlow[3], boxup[3], zl[3], bs[9], cm[ORC_MXGROUP], nb_rm[ORC_MXGROUP * ORC_MXGROUP];
    int ifpd[3];
    orc_tables t;
    double *xp, *xp1, *fp, *dis, *den, *er, *epot, *ekin;
    int *ityp, *statu, *kvois, *indi;
    int epc_enable[ORC_MXGROUP];
    double epc_te[ORC_MXGROUP], epc_alpha[ORC_MXGROUP], epc_cut[ORC_MXGROUP], epc_he[ORC_MXGROUP];
};

orc_cpu *orc_cpu_create(int n, const double *xp, const double *xp1, const int *ityp, const int *statu, int ng,
                        const double *cm, const double boxlow[3], const double zl[3], const int ifpd[3],
                        const double *nb_rm, int mxkvois, const orc_tables *t, int half)
{
    orc_cpu *m = (orc_cpu *)calloc(1, sizeof(orc_cpu));
    size_t n3 = (size_t)n * 3;
    m->n = n; m->ng = ng; m->mxkvois = mxkvois; m->half = half;
    for (int d = 0; d < 3; d++) {
        m->boxlow[d] = boxlow[d]; m->zl[d] = zl[d]; m->boxup[d] = boxlow[d] + zl[d]; m->ifpd[d] = ifpd[d];
    }
    m->bs[0] = m->bs[4] = m->bs[8] = 1.0;
    memcpy(m->cm, cm, sizeof(double) * ng);
    memcpy(m->nb_rm, nb_rm, sizeof(double) * ng * ng);
    m->t = *t;
    m->xp = (double *)malloc(sizeof(double) * n3); m->xp1 = (double *)malloc(sizeof(double) * n3);
    m->fp = (double *)calloc(n3, sizeof(double)); m->dis = (double *)calloc(n3, sizeof(double));
    m->den = (double *)calloc(n, sizeof(double)); m->er = (double *)calloc(n, sizeof(double));
    m->epot = (double *)calloc(n, sizeof(double)); m->ekin = (double *)calloc(n, sizeof(double));
    m->ityp = (int *)malloc(sizeof(int) * n); m->statu = (int *)malloc(sizeof(int) * n);
    m->kvois = (int *)calloc(n, sizeof(int)); m->indi = (int *)malloc(sizeof(int) * (size_t)n * mxkvois);
    memcpy(m->xp, xp, sizeof(double) * n3); memcpy(m->xp1, xp1, sizeof(double) * n3);
    memcpy(m->ityp, ityp, sizeof(int) * n); memcpy(m->statu, statu, sizeof(int) * n);
    return m;
}
void orc_cpu_destroy(orc_cpu *m)
{
    if (!m) return;
    free(m->xp); free(m->xp1); free(m->fp); free(m->dis); free(m->den); free(m->er); free(m->epot); free(m->ekin);
    free(m->ityp); free(m->statu); free(m->kvois); free(m->indi);
    free(m);
}
void orc_cpu_set_epc(orc_cpu *m, const int *enable, const double *te, const double *alpha, const double *cut, const double *he)
{
    m->epc_on = 0;
    for (int g = 0; g < m->ng; g++) {
        m->epc_enable[g] = enable[g]; m->epc_te[g] = te[g]; m->epc_alpha[g] = alpha[g]; m->epc_cut[g] = cut[g];
        m->epc_he[g] = he[g];
        if (enable[g] > 0) m->epc_on = 1;
    }
}
int orc_cpu_rebuild(orc_cpu *m)
{
    return nlist_build_cpu(m->half, m->n, m->xp, m->ityp, m->statu, m->boxlow, m->zl, m->ifpd, m->bs, m->ng, m->nb_rm,
                           m->mxkvois, m->kvois, m->indi);
}
void orc_cpu_force(orc_cpu *m, int with_epot)
{
    int n = m->n;
    if (m->half) {
        orc_force_newton3(n, m->xp, m->ityp, m->kvois, m->indi, n, m->zl, m->ifpd, m->bs, &m->t, m->den, m->er, m->fp,
                          with_epot ? m->epot : NULL, NULL);
        return;
    }
    orc_force_pass1(n, 0, n, m->xp, m->ityp, m->statu, m->kvois, m->indi, n, m->zl, m->ifpd, m->bs, &m->t, m->den);
    orc_force_pass2(n, 0, n, m->xp, m->ityp, m->statu, m->kvois, m->indi, n, m->zl, m->ifpd, m->bs, &m->t, m->den, m->fp, n, NULL);
    if (with_epot)
        orc_force_epot(n, 0, n, m->xp, m->ityp, m->statu, m->kvois, m->indi, n, m->zl, m->ifpd, m->bs, &m->t, m->epot);
}
/* nsteps x For_One_Step on the CPU path; the loop is here, not in the caller.  Returns the number of rebuilds, <0 on a
 * list error (more than mxKVOIS neighbours / atom outside the cells: the reference stops). */
int orc_cpu_run(orc_cpu *m, int itime0, int nsteps, int it0, int nb_uptab, double h)
{
    int n = m->n, nreb = 0;
    for (int s = 0; s < nsteps; s++) {
        int itime = itime0 + s;
        orc_predictor(n, m->xp, m->xp1, m->fp, m->dis, m->statu, m->ityp, m->cm, h, m->boxlow, m->boxup, m->zl, m->ifpd);
        if (nb_uptab > 0 && (itime - it0) % nb_uptab == 0) {
            int rc = orc_cpu_rebuild(m);
            if (rc < 0) return rc;
            nreb++;
        }
        orc_cpu_force(m, 0);
        if (m->epc_on)
            orc_epc(n, m->xp1, m->fp, m->statu, m->ityp, m->ng, m->epc_enable, m->cm, m->epc_te, m->epc_alpha, m->epc_cut,
                    m->epc_he);
        orc_corrector(n, m->xp1, m->fp, m->statu, m->ityp, m->cm, h);
    }
    return nreb;
}
/* HARMIL = (sum EPOT + sum EKIN) / N over active atoms, Common/MD_TypeDef_SimBox.F90:5155-5163 (energies recomputed here) */
double orc_cpu_harmil(orc_cpu *m)
{
    int n = m->n;
    orc_cpu_force(m, 1);
    orc_ekin(n, m->xp1, m->statu, m->ityp, m->cm, m->ekin);
    double se = 0.0, sk = 0.0;
    int na = 0;
    for (int i = 0; i < n; i++)
        if ((m->statu[i] & ORC_STATU_ACTIVE) == ORC_STATU_ACTIVE) { se += m->epot[i]; if (m->ekin[i] >= 0.0) sk += m->ekin[i]; na++; }
    return na ? (se + sk) / (double)na : 0.0;
}
void orc_cpu_get(orc_cpu *m, double *xp, double *xp1, double *fp, double *epot)
{
    size_t n3 = (size_t)m->n * 3;
    if (xp) memcpy(xp, m->xp, sizeof(double) * n3);
    if (xp1) memcpy(xp1, m->xp1, sizeof(double) * n3);
    if (fp) memcpy(fp, m->fp, sizeof(double) * n3);
    if (epot) memcpy(epot, m->epot, sizeof(double) * m->n);
}
const int *orc_cpu_kvois(orc_cpu *m) { return m->kvois; }

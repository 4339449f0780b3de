// host_potentials.cpp -- host-side force-table generation (no CUDA).
//
// Mirrors, for C/C++/Python hosts, what the reference's Fortran potential libraries do through
// Register_Interaction_Table -> Register_ForceTableProc -> Create_Interaction_ForceTable
// (Common/MD_TypeDef_ForceTable.F90:530-622, 890-1056, 1151-1230).  In a Fortran deployment the
// unchanged potential modules fill type(MDForceTable) and hand it to mdb_tables_set; this file
// lets the path run standalone (bench, tests, C++ drivers).
//
// Design: a potential library is a registry {id -> (pair fn, density fn?, embedding fn?)} built
// from small composable function objects (cubic-knot sums, gauge transforms); the table builder
// is generic over the registry.  Arithmetic follows the reference expression order so the tables
// are bit-identical to the oracle's (tests/test_tables.py checks that).
#include <cmath>
#include <cstddef>
#include <functional>
#include <map>
#include <vector>
#include "../../include/mdpscu_b200.h"

namespace {

constexpr double kEvErg = 1.60219e-12; // CP_EVERG  MSMLIB/sor/Common/MSM_Const.F90:83
constexpr double kA2Cm = 1.0e-8;       // CP_A2CM   :67
constexpr double kCm2A = 1.0e8;        // CP_CM2A

struct Val { double f, df; };
using Fn = std::function<Val(double)>;

struct Entry { Fn pair, rho, embed; }; // rho / embed may be empty (no NEFORCE / EMBDF registered)
using Registry = std::map<int, Entry>;

// sum_k a_k (r_k - r)^3 H(r_k - r) and the same with squares; Common/MD_Pot_EAM_Utilities.F90:72-79, 280-287
struct KnotPoly {
    std::vector<double> a, rk;
    void eval(double r, double &s3, double &s2) const
    {
        s3 = 0.0; s2 = 0.0;
        for (size_t i = 0; i < a.size(); ++i) {
            const double d = rk[i] - r;
            const double h = (d >= 0.0) ? 1.0 : 0.0; // :35-44
            s3 = s3 + a[i] * d * d * d * h;
            s2 = s2 + a[i] * d * d * h;
        }
    }
};

// NN_FuncPoly3 (:48-85): r in cm -> (V/2 [erg], -dV/dr [erg/cm])
Fn make_pair_poly3(KnotPoly kp)
{
    return [kp](double r) {
        double s3, s2;
        kp.eval(r * kCm2A, s3, s2);
        return Val{0.5 * s3 * kEvErg, 3.0 * s2 * kEvErg / kA2Cm};
    };
}

// RHO_FuncPoly3 (:259-290) with the Marinica inner clamp (EAM2_WW_Marinica_JPCM25_2013.F90:85-95)
Fn make_rho_poly3_clamped(KnotPoly kp, double rc_cm)
{
    return [kp, rc_cm](double r) {
        double s3, s2;
        if (r <= rc_cm) {
            kp.eval(rc_cm * kCm2A, s3, s2);
            return Val{s3, 0.0};
        }
        kp.eval(r * kCm2A, s3, s2);
        return Val{s3, 3.0 * s2 / kA2Cm};
    };
}

// EMBED_FS_PLOY_Func (:332-368): F = A1 sqrt(rho) + A2 rho^2 + A3 rho^3 ...
Fn make_embed_fs_poly(std::vector<double> a)
{
    return [a](double rho) {
        double f = 0.0, df = 0.0;
        if (rho > 0.0) {
            double t = rho;
            f = a[0] * std::sqrt(t);
            df = 0.5 * a[0] / std::sqrt(t);
            for (size_t i = 2; i <= a.size(); ++i) {
                f = f + a[i - 1] * rho * t;
                df = df + (double)i * a[i - 1] * t;
                t = rho * t;
            }
        }
        return Val{f * kEvErg, df * kEvErg};
    };
}

// ---- Marinica EAM2 W-W; coefficients are double literals in the Fortran source, the KNOTS are
// default-REAL literals (EAM2_WW_Marinica_JPCM25_2013.F90:42-56,80-83) => float32-rounded values.
KnotPoly marinica2_pair_knots()
{
    return KnotPoly{
        {0.960851701343041e2, -0.184410923895214e3, 0.935784079613550e2, -0.798358265041677e1, 0.747034092936229e1,
         -0.152756043708453e1, 0.125205932634393e1, 0.163082162159425e1, -0.141854775352260e1, -0.819936046256149e0,
         0.198013514305908e1, -0.696430179520267e0, 0.304546909722160e-1, -0.163131143161660e1, 0.138409896486177e1},
        {2.564897500000000f, 2.629795000000000f, 2.694692500000000f, 2.866317500000000f, 2.973045000000000f,
         3.079772500000000f, 3.516472500000000f, 3.846445000000000f, 4.176417500000000f, 4.700845000000000f,
         4.895300000000000f, 5.089755000000000f, 5.342952500000000f, 5.401695000000000f, 5.460437500000000f}};
}
KnotPoly marinica2_rho_knots()
{
    return KnotPoly{{-0.420429107805055e1, 0.518217702261442e0, 0.562720834534370e-1, 0.344164178842340e-1},
                    {2.500000000000000f, 3.100000000000000f, 3.500000000000000f, 4.900000000000000f}};
}
const double kMarinicaRc = 2.002970124727e0 * kA2Cm;

Entry marinica2()
{
    return Entry{make_pair_poly3(marinica2_pair_knots()), make_rho_poly3_clamped(marinica2_rho_knots(), kMarinicaRc),
                 make_embed_fs_poly({-5.946454472402710e0, -0.049477376935239e0})};
}

// ---- Bonny EAM1 (EAM1_WHeH_Bonny_JPCM26_2014.F90): W-W = gauge transform of Marinica EAM2 (:15-103)
Entry bonny1_ww()
{
    const Entry m = marinica2();
    const double c = 1.848055990e0 * kEvErg, s = 2.232322602e-1;
    Entry e;
    e.pair = [m, c](double r) {
        Val v = m.pair(r), q = m.rho(r);
        return Val{v.f - c * q.f, v.df - 2.0 * c * q.df}; // :35-36
    };
    e.rho = [m, s](double r) {
        Val q = m.rho(r);
        return Val{q.f * s, q.df * s};
    };
    e.embed = [m, c, s](double rho) {
        const double is = 1.0e+00 / s, cos_ = c * is, rhoi = 1.359141225e0;
        const double a0 = -5.524855802e+00, a1 = 2.317313103e-01, a2 = -3.665345949e-02, a3 = 8.989367404e-03;
        if (rho <= rhoi) { // :88-93
            const double t = rho * is;
            Val f = m.embed(t);
            return Val{f.f + c * t, f.df * is + cos_};
        }
        const double t = rho; // :95-97
        return Val{(a0 + t * (a1 + t * (a2 + t * a3))) * kEvErg, (a1 + t * (2.0 * a2 + 3.0 * a3 * t)) * kEvErg};
    };
    return e;
}

// Finnis-Sinclair W of Ackland & Thetford: Potentials/EM_TB_WangJun_W-HE_2010/FS_Ackland_WW.F90:25-90, registered as id 1 of
// the library EM_TB_WANGJUN_W-HE_2010 for FS_TYPE boxes (EM_TB_ForceTable_WangJun_W_HE_2010.F90:38-40).  No EMBDF: the
// FS kernels take -sqrt(rho) themselves.
Entry ackland_fs_ww()
{
    Entry e;
    e.pair = [](double r) {
        const double c = 3.25e-8, c0 = 47.1346499e16 * kEvErg, c1 = -33.7665655e24 * kEvErg, c2 = 6.2541999e32 * kEvErg;
        const double ackb = 90.3e24 * kEvErg, acka = 1.2e8, ackb0 = 2.7411e-8;
        double core = 0.0, dcore = 0.0;
        if (r < ackb0) { // :60-66
            const double t = ackb0 - r, ex = std::exp(-acka * r);
            core = ackb * (t * t * t) * ex;
            dcore = -ackb * (3.0 * (t * t) * ex + acka * (t * t * t) * ex);
        }
        if (r <= c) { // :68-76
            const double q = c0 + c1 * r + c2 * (r * r), u = r - c;
            return Val{0.5 * ((u * u) * q + core), -(2.0 * u * q + (u * u) * (c1 + 2.0 * c2 * r) + dcore)};
        }
        return Val{0.0, 0.0};
    };
    e.rho = [](double r) {
        const double a = 1.896373e8 * kEvErg, d = 4.400224e-8;
        if (r <= d) return Val{a * a * ((r - d) * (r - d)), -(a * a * 2.0 * (r - d))}; // :84-90
        return Val{0.0, 0.0};
    };
    return e;
}

Registry make_registry(int lib)
{
    Registry r;
    if (lib == MDB_LIB_ACKLAND_FS_W) {
        r[1] = ackland_fs_ww();
        return r;
    }
    if (lib == MDB_LIB_MARINICA_EAM2) {
        r[1] = marinica2(); // Register_ForceTableProc("W","W",...) EAM_ForceTable_Marinica_JPCM25_2013.F90:56-63
    } else if (lib == MDB_LIB_BONNY_EAM1) {
        // ids follow the registration order, EAM_ForceTable_Bonny_JPCM26_2014.F90:57-77
        Fn wh = make_pair_poly3({{1.375733214e+01, 1.296071475e-01}, {2.0, 3.0}});
        Fn whe = make_pair_poly3({{2.1e+01, 8.565323293e-01, 2.750099819e-01}, {1.9, 2.2, 3.5}});
        Fn hh = make_pair_poly3({{4.862785907e-01, 1.018797872e-01}, {2.0, 3.0}});
        Fn hhe = make_pair_poly3({{1.5e+01, 2.563700119e-01, -4.489510592e-02}, {1.8, 2.0, 3.0}});
        Fn hehe = make_pair_poly3({{2.106615791e+00, -2.217639348e-01}, {2.0, 3.0}});
        r[1] = bonny1_ww();
        r[2] = Entry{wh, nullptr, nullptr};   // W <- H
        r[3] = Entry{whe, nullptr, nullptr};  // W <- He
        r[4] = Entry{wh, nullptr, nullptr};   // H <- W
        r[5] = Entry{hh, nullptr, nullptr};   // H <- H
        r[6] = Entry{hhe, nullptr, nullptr};  // H <- He
        r[7] = Entry{whe, nullptr, nullptr};  // He <- W
        r[8] = Entry{hhe, nullptr, nullptr};  // He <- H
        r[9] = Entry{hehe, nullptr, nullptr}; // He <- He
    }
    return r;
}

// first-appearance list of ids (New_ForceTable :559-574, :600-611)
std::vector<int> unique_in_order(const std::vector<int> &ids)
{
    std::vector<int> u;
    for (int id : ids) {
        bool seen = false;
        for (int v : u) seen = seen || (v == id);
        if (!seen) u.push_back(id);
    }
    return u;
}

} // namespace

static int build_tables(const Registry &reg, int ng, const int *ptype, int ntab, int nembd, double rhoscal, double rmax,
                        int *nkind_out, int *nkind1_out, int *kpair, int *kembd, double *potr, double *fpotr,
                        double *potb, double *fpotb, double *fembd, double *dfembd, double *csi_out, double *rhod_out)
{
    if (ng < 1 || ng > MDB_MXGROUP || !ptype || ntab < 2 || nembd < 2 || !(rmax > 0.0)) return MDB_ERR_ARG;
    if (reg.empty()) return MDB_ERR_ARG;

    std::vector<int> all, diag;
    for (int i = 0; i < ng; ++i)
        for (int j = 0; j < ng; ++j) all.push_back(ptype[i + ng * j]); // I outer, J inner
    for (int i = 0; i < ng; ++i) diag.push_back(ptype[i + ng * i]);
    const std::vector<int> fpair = unique_in_order(all), fpair1 = unique_in_order(diag);
    const int nkind = (int)fpair.size(), nkind1 = (int)fpair1.size();
    for (int id : fpair)
        if (!reg.count(id)) return MDB_ERR_ARG; // "potential type # is not available" :930-940

    const double csi = (double)ntab / std::sqrt(rmax); // :591-594
    const double csiv = 1.0 / csi;
    double rhomx = 0.0;
    for (int k = 0; k < nkind; ++k) { // Create_Pairwise_ForceTable :949-976
        const Entry &e = reg.at(fpair[k]);
        for (int i = 1; i <= ntab; ++i) {
            const double t = (double)i * csiv, r = t * t;
            const Val v = e.pair(r);
            const size_t o = (size_t)(i - 1) * nkind + k; // T(NKIND,NTAB) column-major
            potr[o] = v.f * r;
            fpotr[o] = v.df * r;
            if (e.rho) {
                const Val q = e.rho(r);
                potb[o] = q.f;
                fpotb[o] = q.df;
                if (rhomx < q.f * rhoscal) rhomx = q.f * rhoscal; // :965
            } else {
                potb[o] = 0.0;
                fpotb[o] = 0.0;
            }
        }
    }
    double rhod = rhomx / (double)nembd; // :976
    for (int k = 0; k < nkind1; ++k) {   // Create_EMBDFUNTable :1043-1052
        const Entry &e = reg.at(fpair1[k]);
        for (int i = 1; i <= nembd; ++i) {
            const size_t o = (size_t)(i - 1) * nkind1 + k;
            if (e.embed) {
                const Val f = e.embed((double)(i - 1) * rhod);
                fembd[o] = f.f;
                dfembd[o] = f.df;
            } else {
                fembd[o] = 0.0;
                dfembd[o] = 0.0;
            }
        }
    }
    if (rhod <= 1.0e-64) rhod = 1.0; // :1223
    for (int i = 0; i < ng; ++i) {
        for (int j = 0; j < ng; ++j)
            for (int k = 0; k < nkind; ++k)
                if (fpair[k] == ptype[i + ng * j]) kpair[i + ng * j] = k + 1; // :1197-1202
        for (int k = 0; k < nkind1; ++k)
            if (fpair1[k] == ptype[i + ng * i]) kembd[i] = k + 1; // :1213-1218
    }
    *nkind_out = nkind; *nkind1_out = nkind1; *csi_out = csi; *rhod_out = rhod;
    return MDB_OK;
}

extern "C" int mdb_host_ftable_create(int lib, int ng, const int *ptype, int ntab, int nembd, double rhoscal, double rmax,
                                      int *nkind_out, int *nkind1_out, int *kpair, int *kembd, double *potr, double *fpotr,
                                      double *potb, double *fpotb, double *fembd, double *dfembd, double *csi_out,
                                      double *rhod_out)
{
    return build_tables(make_registry(lib), ng, ptype, ntab, nembd, rhoscal, rmax, nkind_out, nkind1_out, kpair, kembd, potr, fpotr,
                        potb, fpotb, fembd, dfembd, csi_out, rhod_out);
}

// ".moldy" library: Register_ForceTableProc_Moldy, Potentials/EAM_NIST/Filedatas_Func_Moldy.F90:21-153.  Six lines: symbol;
// a_k of V; r_k of V; A_k of rho; R_k of rho; constants (the first is the lattice constant, which scales the knots and the
// coefficients, :74-77).  One table id (1): V and rho cubic-knot sums, F = -sqrt(rho) (EMBED_FS_PLOY_Func with A = (-1)).
// After registration the tables are generated like any built-in library (RHOMX = max rho * RHOSCAL).
#include <fstream>
#include <sstream>
#include <string>
extern "C" int mdb_host_moldy_ftable(const char *path, int ntab, int nembd, double rhoscal, double rmax, double *potr, double *fpotr,
                                     double *potb, double *fpotb, double *fembd, double *dfembd, double *csi_out, double *rhod_out)
{
    if (!path) return MDB_ERR_ARG;
    std::ifstream in(path);
    if (!in) return MDB_ERR_ARG;
    std::string line;
    if (!std::getline(in, line)) return MDB_ERR_ARG; // symbol
    std::vector<double> v[5];
    for (int k = 0; k < 5; ++k) {
        if (!std::getline(in, line)) return MDB_ERR_ARG;
        for (char &ch : line)
            if (ch == ',' || ch == ';') ch = ' ';
        std::istringstream ls(line);
        std::string tok;
        while (ls >> tok) { // Extract_Numb: the numbers of the line; Fortran D exponents accepted
            for (char &ch : tok)
                if (ch == 'D' || ch == 'd') ch = 'E';
            char *end = nullptr;
            const double x = std::strtod(tok.c_str(), &end);
            if (end && end != tok.c_str() && *end == '\0') v[k].push_back(x);
        }
    }
    if (v[0].empty() || v[0].size() != v[1].size() || v[2].empty() || v[2].size() != v[3].size() || v[4].empty()) return MDB_ERR_ARG;
    const double a0 = v[4][0];
    KnotPoly pv, pr;
    for (size_t i = 0; i < v[0].size(); ++i) { pv.a.push_back(v[0][i] / std::pow(a0, 3.0)); pv.rk.push_back(v[1][i] * a0); }
    for (size_t i = 0; i < v[2].size(); ++i) { pr.a.push_back(v[2][i] / std::pow(a0, 6.0)); pr.rk.push_back(v[3][i] * a0); }
    Registry reg;
    Entry e;
    e.pair = make_pair_poly3(pv);
    e.rho = [pr](double r) { // RHO_FuncPoly3, Common/MD_Pot_EAM_Utilities.F90:259-290 (no inner clamp)
        double s3, s2;
        pr.eval(r * kCm2A, s3, s2);
        return Val{s3, 3.0 * s2 / kA2Cm};
    };
    e.embed = make_embed_fs_poly({-1.0});
    reg[1] = e;
    const int ptype = 1;
    int nkind = 0, nkind1 = 0, kpair = 0, kembd = 0;
    return build_tables(reg, 1, &ptype, ntab, nembd, rhoscal, rmax, &nkind, &nkind1, &kpair, &kembd, potr, fpotr, potb, fpotb, fembd,
                        dfembd, csi_out, rhod_out);
}

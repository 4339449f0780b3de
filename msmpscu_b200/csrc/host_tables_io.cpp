// host_tables_io.cpp -- external force tables (no CUDA): the NIST "setfl" and "lspt" importers and the
// MDPSCU .pair/.embd table files.
//
// Reference behaviour restated here (paths relative to MDLIB/sor):
//   * Potentials/EAM_NIST/Filedatas_Func_Setfl.F90:153-296   reader (header, F(rho), rho(r), r*V(r) blocks,
//                                                            the "INF"/"NAN" first-token patch)
//   * :297-312 + LIB/sor/f/MATH90A/DBINT4.F, DBVALU.F        cubic spline through the file's points with zero
//                                                            second derivative at both ends (IBCL = IBCR = 2,
//                                                            FBCL = FBCR = 0)
//   * :320-462                                               range rules of Vr_/Rhor_/Frho_Spline and the unit
//                                                            conversions of NN_/RHO_/EMBED_Spline
//   * Potentials/EAM_NIST/NIST_ForceTable.F90:332-398        Generate_NIST_ForceTalbe: one table per (I<-J) id,
//                                                            Rmax and RHOMX taken from the file
//   * Common/MD_TypeDef_ForceTable.F90:1315-1459             Export_ForceTable (.pair/.embd, 10 significant digits)
//   * :1461-1591                                             Import_ForceTable
//   * :1595-1855 + LIB/sor/f/MiniUtilities/DINTF2.F:17-232   Register_Imported_ForceTable: re-grid onto the run's
//                                                            sqrt(r) grid with SPLID1 (IOP = 5: end slopes from the
//                                                            cubic through the four end points) and SPLID2
//
// The interpolants are defined mathematically (natural / end-slope cubic splines); they are computed here by one
// tridiagonal solve for the knot second derivatives instead of the B-spline (DBINT4) or the two-vector
// elimination (SPLID1) forms, so values agree with the reference to round-off, not bit for bit.  Pinned by
// tests/test_tables_io.py on the reference's own exported tables for examples/NIST_Potentials/Cu_EAM.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "../../include/mdpscu_b200.h"

namespace {

constexpr double kEvErg = 1.60219e-12; // CP_EVERG  MSMLIB/sor/Common/MSM_Const.F90:83
constexpr double kErgEv = 1.0 / kEvErg; // CP_ERGEV
constexpr double kA2Cm = 1.0e-8;
constexpr double kCm2A = 1.0e8;

// Cubic spline held as knot second derivatives m[i]; evaluation in the SPLID2 form (DINTF2.F:202-217).
struct Spline {
    std::vector<double> x, y, m;

    // end conditions: natural (m0 = mn = 0) or prescribed first derivatives s0, sn
    void fit(bool natural, double s0 = 0.0, double sn = 0.0)
    {
        const int n = (int)x.size();
        m.assign(n, 0.0);
        if (n < 3) return;
        std::vector<double> a(n), b(n), c(n), d(n);
        for (int i = 1; i < n - 1; ++i) {
            const double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
            a[i] = h0 / 6.0;
            b[i] = (h0 + h1) / 3.0;
            c[i] = h1 / 6.0;
            d[i] = (y[i + 1] - y[i]) / h1 - (y[i] - y[i - 1]) / h0;
        }
        if (natural) {
            a[0] = 0.0; b[0] = 1.0; c[0] = 0.0; d[0] = 0.0;
            a[n - 1] = 0.0; b[n - 1] = 1.0; c[n - 1] = 0.0; d[n - 1] = 0.0;
        } else {
            const double h0 = x[1] - x[0], hn = x[n - 1] - x[n - 2];
            a[0] = 0.0; b[0] = h0 / 3.0; c[0] = h0 / 6.0; d[0] = (y[1] - y[0]) / h0 - s0;
            a[n - 1] = hn / 6.0; b[n - 1] = hn / 3.0; c[n - 1] = 0.0; d[n - 1] = sn - (y[n - 1] - y[n - 2]) / hn;
        }
        for (int i = 1; i < n; ++i) { // Thomas
            const double w = a[i] / b[i - 1];
            b[i] -= w * c[i - 1];
            d[i] -= w * d[i - 1];
        }
        m[n - 1] = d[n - 1] / b[n - 1];
        for (int i = n - 2; i >= 0; --i) m[i] = (d[i] - c[i] * m[i + 1]) / b[i];
    }

    // SPLID1 with IOP = (5,5): slopes of the cubic through the first / last four points (DINTF2.F:68-88)
    void fit_endslope4()
    {
        const int n = (int)x.size();
        if (n < 4) { fit(true); return; }
        const double a1 = x[0] - x[1], a2 = x[0] - x[2], a3 = x[0] - x[3], a4 = x[1] - x[2], a5 = x[1] - x[3], a6 = x[2] - x[3];
        const double s0 = y[0] * (1.0 / a1 + 1.0 / a2 + 1.0 / a3) - a2 * a3 * y[1] / (a1 * a4 * a5) + a1 * a3 * y[2] / (a2 * a4 * a6) -
                          a1 * a2 * y[3] / (a3 * a5 * a6);
        const double b1 = x[n - 1] - x[n - 4], b2 = x[n - 1] - x[n - 3], b3 = x[n - 1] - x[n - 2], b4 = x[n - 2] - x[n - 4],
                     b5 = x[n - 2] - x[n - 3], b6 = x[n - 3] - x[n - 4];
        const double sn = -b2 * b3 * y[n - 4] / (b6 * b4 * b1) + b1 * b3 * y[n - 3] / (b6 * b5 * b2) - b1 * b2 * y[n - 2] / (b4 * b5 * b3) +
                          y[n - 1] * (1.0 / b1 + 1.0 / b2 + 1.0 / b3);
        fit(false, s0, sn);
    }

    int interval(double t) const // SPLID2 :190-196 -- the end intervals also serve outside the range
    {
        const int n = (int)x.size();
        if (t <= x[0]) return 0;
        if (t >= x[n - 1]) return n - 2;
        int lo = 0, hi = n - 1; // x[lo] <= t < x[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) / 2;
            if (t >= x[mid]) lo = mid; else hi = mid;
        }
        return lo;
    }

    void eval(double t, double &f, double &df) const
    {
        const int i = interval(t);
        const double h = x[i + 1] - x[i], u = x[i + 1] - t, v = t - x[i];
        f = (m[i] * u * u * u + m[i + 1] * v * v * v) / (6.0 * h) + (y[i + 1] / h - m[i + 1] * h / 6.0) * v + (y[i] / h - h * m[i] / 6.0) * u;
        df = (m[i + 1] * v * v - m[i] * u * u) / (2.0 * h) + (y[i + 1] - y[i]) / h + h * (m[i] - m[i + 1]) / 6.0;
    }
};

struct SetflElement {
    std::string name, lattice;
    int z = 0;
    double mass = 0.0, alat = 0.0;
    bool frho_zero = false; // m_FRHO_ZERO, Filedatas_Func_Setfl.F90:209
    bool rho_na = false;    // lspt: RHO_R file "NA" -> no density from this element
    Spline frho, rhor;
    std::vector<Spline> vr; // V_r(J), J <= I
};

struct Setfl {
    int ne = 0, nrho = 0, nr = 0;
    double drho = 0.0, dr = 0.0, cutoff = 0.0; // Angstrom
    double rhomx = 0.0;                        // FTable%RHOMX the importer sets
    bool v_is_rv = true;                       // setfl files hold r*V(r); lspt files hold V(r)
    std::vector<SetflElement> el;
};

// one block of n numbers; a leading "INF"/"NAN" token is replaced by the value that follows it (:213-231, :249-268)
bool read_block(std::istream &in, int n, std::vector<double> &out)
{
    out.resize(n);
    for (int k = 0; k < n; ++k) {
        std::string tok;
        if (!(in >> tok)) return false;
        char *end = nullptr;
        const double v = std::strtod(tok.c_str(), &end);
        const bool numeric = end && *end == '\0' && std::isfinite(v);
        if (!numeric) {
            if (k != 0) return false;
            out[0] = NAN;
            continue;
        }
        out[k] = v;
    }
    if (n > 1 && std::isnan(out[0])) out[0] = out[1];
    return true;
}

int load_setfl(const char *path, Setfl &s)
{
    std::ifstream in(path);
    if (!in) return MDB_ERR_ARG;
    std::string line;
    for (int i = 0; i < 3; ++i)
        if (!std::getline(in, line)) return MDB_ERR_ARG; // three comment lines
    if (!std::getline(in, line)) return MDB_ERR_ARG;
    {
        std::istringstream ls(line);
        if (!(ls >> s.ne) || s.ne < 1 || s.ne > MDB_MXGROUP) return MDB_ERR_ARG;
        s.el.resize(s.ne);
        for (int i = 0; i < s.ne; ++i)
            if (!(ls >> s.el[i].name)) return MDB_ERR_ARG;
    }
    if (!(in >> s.nrho >> s.drho >> s.nr >> s.dr >> s.cutoff)) return MDB_ERR_ARG;
    if (s.nrho < 4 || s.nr < 4 || !(s.cutoff > 0.0)) return MDB_ERR_ARG;
    s.rhomx = (double)s.nrho * s.drho; // :178
    std::vector<double> rho(s.nrho), r(s.nr);
    for (int i = 0; i < s.nrho; ++i) rho[i] = (double)i * s.drho; // :181-183
    const double minr = s.cutoff / (double)s.nr;                    // :184-187 (cutoff/Nr, not the file's dr)
    for (int i = 0; i < s.nr; ++i) r[i] = (double)i * minr;
    for (int i = 0; i < s.ne; ++i) {
        SetflElement &e = s.el[i];
        if (!(in >> e.z >> e.mass >> e.alat >> e.lattice)) return MDB_ERR_ARG;
        e.frho.x = rho;
        if (!read_block(in, s.nrho, e.frho.y)) return MDB_ERR_ARG;
        double mx = e.frho.y[0], mn = e.frho.y[0];
        for (double v : e.frho.y) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
        e.frho_zero = (mx == 0.0 && mn == 0.0);
        e.frho.fit(true);
        e.rhor.x = r;
        if (!read_block(in, s.nr, e.rhor.y)) return MDB_ERR_ARG;
        e.rhor.fit(true);
    }
    for (int i = 0; i < s.ne; ++i) { // :236-272: V blocks in the order (1,1), (2,1), (2,2), ...
        s.el[i].vr.resize(i + 1);
        for (int j = 0; j <= i; ++j) {
            Spline &v = s.el[i].vr[j];
            v.x = r;
            if (!read_block(in, s.nr, v.y)) return MDB_ERR_ARG;
            v.fit(true);
        }
    }
    return MDB_OK;
}

void trim(std::string &s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    s = (a == std::string::npos) ? std::string() : s.substr(a, b - a + 1);
}

// ---- "lspt": an index file naming one two-column (x, f) file per function.  Potentials/EAM_NIST/Filedatas_Func_Lspt.F90:79-300.
std::vector<std::string> quoted(const std::string &line)
{
    std::vector<std::string> out;
    size_t i = 0;
    while ((i = line.find('"', i)) != std::string::npos) {
        const size_t j = line.find('"', i + 1);
        if (j == std::string::npos) break;
        std::string t = line.substr(i + 1, j - i - 1);
        trim(t);
        out.push_back(t);
        i = j + 1;
    }
    return out;
}
bool next_line(std::istream &in, std::string &line, char comment)
{
    while (std::getline(in, line)) {
        std::string t = line;
        trim(t);
        if (t.empty() || t[0] == comment) continue;
        line = t;
        return true;
    }
    return false;
}
bool is_na(std::string s)
{
    for (char &ch : s) ch = (char)std::toupper((unsigned char)ch);
    return s == "NA";
}
bool read_spt(const std::string &path, Spline &sp)
{
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    sp.x.clear(); sp.y.clear();
    while (std::getline(in, line)) {
        std::string t = line;
        trim(t);
        if (t.empty() || t[0] == '#') continue;
        std::istringstream ls(t);
        double a, b;
        if (!(ls >> a >> b)) return false;
        sp.x.push_back(a); sp.y.push_back(b);
    }
    if (sp.x.size() < 4) return false;
    sp.fit(true);
    return true;
}
int load_lspt(const char *path, Setfl &s)
{
    std::ifstream in(path);
    if (!in) return MDB_ERR_ARG;
    std::string dir(path);
    const size_t slash = dir.find_last_of("/\\");
    dir = (slash == std::string::npos) ? std::string() : dir.substr(0, slash + 1);
    std::string line;
    if (!next_line(in, line, '!')) return MDB_ERR_ARG;
    {
        std::string kw = line.substr(0, 5);
        for (char &ch : kw) ch = (char)std::toupper((unsigned char)ch);
        if (kw != "&LSPT") return MDB_ERR_ARG; // :105-107
    }
    if (!next_line(in, line, '!')) return MDB_ERR_ARG;
    {
        size_t i = line.find_first_of("0123456789");
        if (i == std::string::npos) return MDB_ERR_ARG;
        s.ne = std::atoi(line.c_str() + i);
        if (s.ne < 1 || s.ne > MDB_MXGROUP) return MDB_ERR_ARG;
    }
    if (!next_line(in, line, '!')) return MDB_ERR_ARG;
    const std::vector<std::string> names = quoted(line);
    if ((int)names.size() != s.ne) return MDB_ERR_ARG; // :119-123
    s.el.resize(s.ne);
    s.v_is_rv = false;
    double rhomax = 0.0, rmax = 0.0;
    for (int i = 0; i < s.ne; ++i) {
        SetflElement &e = s.el[i];
        e.name = names[i];
        if (!next_line(in, line, '!')) return MDB_ERR_ARG;
        std::vector<std::string> q = quoted(line);
        if (q.empty()) return MDB_ERR_ARG;
        if (is_na(q[0])) e.frho_zero = true; // :135-137
        else {
            if (!read_spt(dir + q[0], e.frho)) return MDB_ERR_ARG;
            double hi = e.frho.x[0];
            for (double v : e.frho.x) hi = v > hi ? v : hi;
            rhomax = hi > rhomax ? hi : rhomax; // :160
        }
        if (!next_line(in, line, '!')) return MDB_ERR_ARG;
        q = quoted(line);
        if (q.empty()) return MDB_ERR_ARG;
        if (is_na(q[0])) e.rho_na = true;
        else if (!read_spt(dir + q[0], e.rhor)) return MDB_ERR_ARG;
    }
    for (int i = 0; i < s.ne; ++i) {
        if (!next_line(in, line, '!')) return MDB_ERR_ARG;
        const std::vector<std::string> q = quoted(line);
        if ((int)q.size() != i + 1) return MDB_ERR_ARG;
        s.el[i].vr.resize(i + 1);
        for (int j = 0; j <= i; ++j) {
            // the reference opens SUBSTR(1) for every J (:225): all V(I,J) of a line come from its FIRST file
            if (!read_spt(dir + q[0], s.el[i].vr[j])) return MDB_ERR_ARG;
            double hi = s.el[i].vr[j].x[0];
            for (double v : s.el[i].vr[j].x) hi = v > hi ? v : hi;
            rmax = hi > rmax ? hi : rmax; // :243
        }
    }
    s.rhomx = rhomax;   // :250
    s.cutoff = rmax;    // FTable%RMAX = RMAX*CP_A2CM :251
    return MDB_OK;
}

// value and derivative with the reference's range rule: inside [x0, xn] the spline; below x0 the end value with
// zero slope; above xn zero -- or, for the embedding function, the end value (Filedatas_Func_Setfl.F90:334-352, :418-440)
void ranged(const Spline &sp, double t, bool hold_above, double &f, double &df)
{
    const double lo = sp.x.front(), hi = sp.x.back();
    if (t >= lo && t <= hi) { sp.eval(t, f, df); return; }
    double dummy;
    if (t < lo) { sp.eval(lo, f, dummy); df = 0.0; return; }
    if (hold_above) { sp.eval(hi, f, dummy); df = 0.0; return; }
    f = 0.0; df = 0.0;
}


struct TableFile {
    std::string pottype;
    std::vector<int> ids;
    int npoint = 0;
    std::vector<double> x;              // Rij or RHO column
    std::vector<std::vector<double>> col; // per table: 4 (pair) or 2 (embd) columns
};

// header keywords of Import_ForceTable (:1488-1514); '!' starts a comment line
int read_table_file(const std::string &path, int ncol, TableFile &t)
{
    std::ifstream in(path);
    if (!in) return MDB_ERR_ARG;
    std::string line;
    int nc = -1;
    bool data = false;
    while (std::getline(in, line)) {
        std::string s = line;
        trim(s);
        if (s.empty() || s[0] == '!') continue;
        if (s[0] != '&') continue;
        std::istringstream ls(s);
        std::string kw;
        ls >> kw;
        for (char &ch : kw) ch = (char)std::toupper((unsigned char)ch);
        if (kw == "&NUMTABLE") {
            ls >> nc;
            if (nc < 0 || nc > MDB_MXGROUP * MDB_MXGROUP) return MDB_ERR_ARG;
            std::string rest;
            std::getline(ls, rest);
            for (char &ch : rest)
                if (!std::isdigit((unsigned char)ch) && ch != '-') ch = ' ';
            std::istringstream rs(rest);
            int id;
            while ((int)t.ids.size() < nc && (rs >> id)) t.ids.push_back(id);
        } else if (kw == "&NUMPOINT") {
            ls >> t.npoint;
        } else if (kw == "&POTTYPE") {
            std::string v;
            ls >> v;
            std::string q;
            for (char ch : v)
                if (ch != '"' && ch != '\'') q.push_back(ch);
            t.pottype = q;
        } else if (kw.rfind("&#", 0) == 0) {
            data = true;
            break;
        }
    }
    if (!data || nc < 0 || t.npoint < 4 || (int)t.ids.size() != nc) return MDB_ERR_ARG;
    t.x.resize(t.npoint);
    t.col.assign((size_t)nc * ncol, std::vector<double>(t.npoint));
    for (int j = 0; j < t.npoint; ++j) {
        int it;
        if (!(in >> it >> t.x[j])) return MDB_ERR_ARG;
        for (int k = 0; k < nc * ncol; ++k)
            if (!(in >> t.col[k][j])) return MDB_ERR_ARG;
    }
    return MDB_OK;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------
extern "C" int mdb_host_setfl_info(const char *path, int *nelem, int *nrho, int *nr, double *cutoff_cm, double *rhomx, char *names,
                                   int names_stride, int *z, double *mass, double *alat)
{
    if (!path) return MDB_ERR_ARG;
    Setfl s;
    const int rc = load_setfl(path, s);
    if (rc != MDB_OK) return rc;
    if (nelem) *nelem = s.ne;
    if (nrho) *nrho = s.nrho;
    if (nr) *nr = s.nr;
    if (cutoff_cm) *cutoff_cm = s.cutoff * kA2Cm;      // FTable%RMAX  :179
    if (rhomx) *rhomx = s.rhomx;                       // FTable%RHOMX :178
    for (int i = 0; i < s.ne; ++i) {
        if (names && names_stride > 1) {
            std::strncpy(names + (size_t)i * names_stride, s.el[i].name.c_str(), names_stride - 1);
            names[(size_t)i * names_stride + names_stride - 1] = '\0';
        }
        if (z) z[i] = s.el[i].z;
        if (mass) mass[i] = s.el[i].mass;
        if (alat) alat[i] = s.el[i].alat;
    }
    return MDB_OK;
}

namespace {
// Generate_NIST_ForceTalbe (NIST_ForceTable.F90:332-398) over the spline callbacks of either importer
int nist_tables(const Setfl &s, int ntab, int nembd, double rmax, int *nkind_out, double *potr, double *fpotr, double *potb,
                double *fpotb, double *fembd, double *dfembd, double *csi_out, double *rhod_out, double *rmax_out)
{
    const int ne = s.ne, nkind = ne * ne;
    if (!(rmax > 0.0)) rmax = s.cutoff * kA2Cm; // the importer overrides the run's range with the file's
    const double csi = (double)ntab / std::sqrt(rmax), csiv = 1.0 / csi; // :352-354
    const double rhod = s.rhomx / (double)nembd;                         // :339, restored at :381-384
    for (int it = 1; it <= nkind; ++it) {
        int i = (it - 1) / ne + 1, j = it - (i - 1) * ne; // table id it = "I <- J"
        const int iv = (j > i) ? j : i, jv = (j > i) ? i : j;
        const Spline &v = s.el[iv - 1].vr[jv - 1];
        const Spline &q = s.el[j - 1].rhor; // density contributed by the neighbour element J
        const bool rho_off = s.el[i - 1].frho_zero || s.el[j - 1].rho_na;
        const int k = it - 1; // FPAIR(IFORCE) = IFORCE (:369-372): kind index = id
        for (int n = 1; n <= ntab; ++n) {
            const double t = (double)n * csiv, r = t * t, ra = r * kCm2A;
            double f, df, pot, fpot;
            ranged(v, ra, false, f, df);
            if (s.v_is_rv) { // setfl NN_Spline (Filedatas_Func_Setfl.F90:443-452): the file holds r*V
                pot = f / ra;
                fpot = (df - pot) / ra;
            } else {         // lspt NN_Spline (Filedatas_Func_Lspt.F90:488-496): the file holds V
                pot = f;
                fpot = df;
            }
            pot = 0.5 * pot * kEvErg;
            fpot = -1.0 * fpot * kEvErg * kCm2A;
            const size_t o = (size_t)(n - 1) * nkind + k;
            potr[o] = pot * r; // Create_Pairwise_ForceTable :949-955
            fpotr[o] = fpot * r;
            if (rho_off) {
                potb[o] = 0.0;
                fpotb[o] = 0.0;
            } else {
                ranged(q, ra, false, f, df);
                potb[o] = f;
                fpotb[o] = -1.0 * df * kCm2A; // RHO_Spline
            }
        }
        for (int n = 1; n <= nembd; ++n) { // Create_EMBDFUNTable :1043-1047 through Frho_Spline / EMBED_Spline
            const size_t o = (size_t)(n - 1) * nkind + k;
            if (i == j && !s.el[i - 1].frho.x.empty()) { // lspt "NA": no embedding function (Frho_Spline returns 0)
                double f = 0.0, df = 0.0;
                ranged(s.el[i - 1].frho, (double)(n - 1) * rhod, true, f, df);
                fembd[o] = f * kEvErg;
                dfembd[o] = df * kEvErg;
            } else {
                fembd[o] = 0.0;
                dfembd[o] = 0.0;
            }
        }
    }
    if (nkind_out) *nkind_out = nkind;
    if (csi_out) *csi_out = csi;
    if (rhod_out) *rhod_out = (rhod <= 1.0e-64) ? 1.0 : rhod;
    if (rmax_out) *rmax_out = rmax;
    return MDB_OK;
}
} // namespace

extern "C" int mdb_host_setfl_ftable(const char *path, int ntab, int nembd, double rmax, int *nkind_out, double *potr, double *fpotr,
                                     double *potb, double *fpotb, double *fembd, double *dfembd, double *csi_out, double *rhod_out,
                                     double *rmax_out)
{
    if (!path || ntab < 2 || nembd < 2 || !potr || !fpotr || !potb || !fpotb || !fembd || !dfembd) return MDB_ERR_ARG;
    Setfl s;
    const int rc = load_setfl(path, s);
    if (rc != MDB_OK) return rc;
    return nist_tables(s, ntab, nembd, rmax, nkind_out, potr, fpotr, potb, fpotb, fembd, dfembd, csi_out, rhod_out, rmax_out);
}

// Register_ForceTableProc_SPT + Generate_NIST_ForceTalbe for a ".lspt" library (Filedatas_Func_Lspt.F90:79-300, 488-541)
extern "C" int mdb_host_lspt_info(const char *path, int *nelem, double *cutoff_cm, double *rhomx, char *names, int names_stride)
{
    if (!path) return MDB_ERR_ARG;
    Setfl s;
    const int rc = load_lspt(path, s);
    if (rc != MDB_OK) return rc;
    if (nelem) *nelem = s.ne;
    if (cutoff_cm) *cutoff_cm = s.cutoff * kA2Cm;
    if (rhomx) *rhomx = s.rhomx;
    for (int i = 0; i < s.ne && names && names_stride > 1; ++i) {
        std::strncpy(names + (size_t)i * names_stride, s.el[i].name.c_str(), names_stride - 1);
        names[(size_t)i * names_stride + names_stride - 1] = '\0';
    }
    return MDB_OK;
}
extern "C" int mdb_host_lspt_ftable(const char *path, int ntab, int nembd, double rmax, int *nkind_out, double *potr, double *fpotr,
                                    double *potb, double *fpotb, double *fembd, double *dfembd, double *csi_out, double *rhod_out,
                                    double *rmax_out)
{
    if (!path || ntab < 2 || nembd < 2 || !potr || !fpotr || !potb || !fpotb || !fembd || !dfembd) return MDB_ERR_ARG;
    Setfl s;
    const int rc = load_lspt(path, s);
    if (rc != MDB_OK) return rc;
    return nist_tables(s, ntab, nembd, rmax, nkind_out, potr, fpotr, potb, fpotb, fembd, dfembd, csi_out, rhod_out, rmax_out);
}

// Export_ForceTable, Common/MD_TypeDef_ForceTable.F90:1315-1459.  ids[k] = FPAIR(k) of table row k (written in increasing
// id order like the reference); tables in the T(NKIND,NTAB) column-major layout of mdb_tables_set.
extern "C" int mdb_host_ftable_export(const char *fname, int pot_type, int nkind, const int *ids, int ntab, double csi, const double *potr,
                                      const double *fpotr, const double *potb, const double *fpotb, int nkind1, const int *ids1, int nembd,
                                      double rhod, const double *fembd, const double *dfembd)
{
    if (!fname || nkind < 1 || !ids || ntab < 1 || nembd < 1 || nkind1 < 0 || (nkind1 > 0 && !ids1)) return MDB_ERR_ARG;
    const bool fs = (pot_type == MDB_POT_FS);
    const char *ptname = fs ? "FS_TYPE" : "EAM_TYPE";
    const double rhounit = fs ? kErgEv * kErgEv : 1.0;
    auto order = [](int n, const int *id) { // "reorder the table ID in an incremental order" :1349-1354
        std::vector<int> tid(n);
        for (int i = 0; i < n; ++i) {
            int rank = 0;
            for (int k = 0; k < n; ++k) rank += (id[k] <= id[i] && id[k] > 0) ? 1 : 0;
            tid[rank - 1] = i;
        }
        return tid;
    };
    {
        const std::string path = std::string(fname) + ".pair";
        FILE *f = std::fopen(path.c_str(), "w");
        if (!f) return MDB_ERR_ARG;
        const std::vector<int> tid = order(nkind, ids);
        std::fprintf(f, "&MDPSCU_POTTAB.Pair\n");
        std::fprintf(f, "!     written by mdb_host_ftable_export (format of Export_ForceTable)\n");
        std::fprintf(f, "!     NOTE: For FS potential, RHO is in eV^2, the potential calculated by -sqrt(RHO)+V(r)\n");
        std::fprintf(f, "!           is in unit eV, and d(RHO)/dr in eV^2/A.  For EAM potential no specific unit is\n");
        std::fprintf(f, "!           assigned to RHO; the potential F(RHO)+V(r) is in eV.\n");
        std::fprintf(f, "&POTTYPE \"%s\"\n", ptname);
        std::fprintf(f, "&NUMTABLE %7d table IDs: ", nkind);
        for (int j = 0; j < nkind; ++j) std::fprintf(f, "%5d", ids[tid[j]]);
        std::fprintf(f, "\n&NUMPOINT %7d\n", ntab);
        std::fprintf(f, "&#           Rij(A)     ");
        for (int j = 0; j < nkind; ++j) {
            const int id = ids[tid[j]];
            std::fprintf(f, "      r*V%d(r)[eV*A]        -r*dV%d/dr[eV]            RHO%d(r)            -dRHO%d/dr[/A] ", id, id, id, id);
        }
        std::fprintf(f, "\n");
        const double csiv = 1.0 / csi;
        for (int it = 1; it <= ntab; ++it) {
            const double t = (double)it * csiv, r = (t * t) * kCm2A;
            std::fprintf(f, " %6d %21.9E", it, r);
            for (int j = 0; j < nkind; ++j) {
                const size_t o = (size_t)(it - 1) * nkind + tid[j];
                std::fprintf(f, "%21.9E%21.9E%21.9E%21.9E", potr[o] * 2.0 * kErgEv * kCm2A, fpotr[o] * kErgEv, potb[o] * rhounit,
                             fpotb[o] * rhounit * kA2Cm);
            }
            std::fprintf(f, "\n");
        }
        std::fclose(f);
    }
    {
        const std::string path = std::string(fname) + ".embd";
        FILE *f = std::fopen(path.c_str(), "w");
        if (!f) return MDB_ERR_ARG;
        int nc = nkind1;
        std::vector<int> tid = nc > 0 ? order(nc, ids1) : std::vector<int>();
        const bool fs_formula = fs;                // FS files carry -sqrt(RHO) itself (:1431-1440)
        if (fs && nc == 0) nc = 1;                 // :1409
        std::fprintf(f, "&MDPSCU_POTTAB.Embd\n");
        std::fprintf(f, "!     written by mdb_host_ftable_export (format of Export_ForceTable)\n");
        std::fprintf(f, "&POTTYPE \"%s\"\n", ptname);
        std::fprintf(f, "&NUMTABLE %7d table IDs: ", nc);
        for (int j = 0; j < nkind1; ++j) std::fprintf(f, "%5d", ids1[tid[j]]);
        std::fprintf(f, "\n&NUMPOINT %7d\n", nembd);
        std::fprintf(f, "&#           RHO       ");
        for (int j = 0; j < nkind1; ++j) std::fprintf(f, "         F%d(RHO)      dF%d/d(RHO)", ids1[tid[j]], ids1[tid[j]]);
        std::fprintf(f, "\n");
        for (int it = 1; it <= nembd; ++it) {
            const double r = (double)(it - 1) * rhod * rhounit;
            std::fprintf(f, " %6d %16.8E", it, r);
            for (int j = 0; j < nc; ++j) {
                double den, deni;
                if (fs_formula) {
                    if (r > 0.0) { den = -std::sqrt(r) / kErgEv; deni = 0.5 / den; } else { den = 0.0; deni = 0.0; }
                } else {
                    const size_t o = (size_t)(it - 1) * nkind1 + tid[j];
                    den = fembd[o];
                    deni = dfembd[o];
                }
                std::fprintf(f, "%16.8E%16.8E", den * kErgEv, deni);
            }
            std::fprintf(f, "\n");
        }
        std::fclose(f);
    }
    return MDB_OK;
}

// header of a .pair/.embd couple: table ids and point counts (Import_ForceTable :1488-1514, :1543-1569)
extern "C" int mdb_host_ftable_file_info(const char *fname, int *pot_type, int *nkind, int *ids, int *ntab, int *nkind1, int *ids1,
                                         int *nembd, double *rmax_cm, double *rhomx)
{
    if (!fname) return MDB_ERR_ARG;
    TableFile p, e;
    int rc = read_table_file(std::string(fname) + ".pair", 4, p);
    if (rc != MDB_OK) return rc;
    rc = read_table_file(std::string(fname) + ".embd", 2, e);
    if (rc != MDB_OK) return rc;
    const bool fs = (p.pottype == "FS_TYPE");
    if (pot_type) *pot_type = fs ? MDB_POT_FS : MDB_POT_EAM;
    if (nkind) *nkind = (int)p.ids.size();
    if (nkind1) *nkind1 = (int)e.ids.size();
    if (ntab) *ntab = p.npoint;
    if (nembd) *nembd = e.npoint;
    if (ids) for (size_t k = 0; k < p.ids.size(); ++k) ids[k] = p.ids[k];
    if (ids1) for (size_t k = 0; k < e.ids.size(); ++k) ids1[k] = e.ids[k];
    if (rmax_cm) *rmax_cm = p.x.back() * kA2Cm;
    const double rhounit = fs ? kErgEv * kErgEv : 1.0;
    if (rhomx) *rhomx = (e.x.back() / rhounit / (double)(e.npoint - 1)) * (double)e.npoint; // :1757
    return MDB_OK;
}

// Import_ForceTable + Register_Imported_ForceTable: read fname.pair / fname.embd and re-grid the tables the box's
// PTYPE(ng,ng) (column-major) refers to onto the run's grid r_k = (k*sqrt(rmax)/ntab)^2, k = 1..ntab, and
// rho_k = (k-1)*RHOD.  Kinds are numbered in first-appearance order of PTYPE (New_ForceTable :559-611).
extern "C" int mdb_host_ftable_import(const char *fname, int ng, const int *ptype, int ntab, int nembd, double rmax, int *pot_type_out,
                                      int *nkind_out, int *nkind1_out, int *kpair, int *kembd, double *potr, double *fpotr, double *potb,
                                      double *fpotb, double *fembd, double *dfembd, double *csi_out, double *rhod_out)
{
    if (!fname || ng < 1 || ng > MDB_MXGROUP || !ptype || ntab < 2 || nembd < 2 || !(rmax > 0.0)) return MDB_ERR_ARG;
    TableFile p, e;
    int rc = read_table_file(std::string(fname) + ".pair", 4, p);
    if (rc != MDB_OK) return rc;
    rc = read_table_file(std::string(fname) + ".embd", 2, e);
    if (rc != MDB_OK) return rc;
    const bool fs = (p.pottype == "FS_TYPE");
    const double rhounit = fs ? kErgEv * kErgEv : 1.0;
    // file units -> CGS (:1527-1531, :1585-1586)
    const int nt = p.npoint, ne = e.npoint;
    std::vector<double> R(nt), RHO(ne);
    for (int j = 0; j < nt; ++j) R[j] = p.x[j] * kA2Cm;
    for (int j = 0; j < ne; ++j) RHO[j] = e.x[j] / rhounit;
    const double scale[4] = {0.5 * kEvErg * kA2Cm, kEvErg, 1.0 / rhounit, kCm2A / rhounit};

    auto find = [](const std::vector<int> &v, int id) {
        for (size_t k = 0; k < v.size(); ++k)
            if (v[k] == id) return (int)k;
        return -1;
    };
    std::vector<int> fpair, fpair1;
    for (int i = 0; i < ng; ++i)
        for (int j = 0; j < ng; ++j) {
            const int id = ptype[i + ng * j];
            if (find(p.ids, id) < 0) return MDB_ERR_ARG; // "cannot find force table #" :1616-1626
            if (find(fpair, id) < 0) fpair.push_back(id);
        }
    for (int i = 0; i < ng; ++i) {
        const int id = ptype[i + ng * i];
        if (find(e.ids, id) < 0) return MDB_ERR_ARG;
        if (find(fpair1, id) < 0) fpair1.push_back(id);
    }
    const int nkind = (int)fpair.size(), nkind1 = (int)fpair1.size();
    const double csi = (double)ntab / std::sqrt(rmax), csiv = 1.0 / csi;
    double trmax = R[0];
    for (double v : R) trmax = v > trmax ? v : trmax;
    double *out4[4] = {potr, fpotr, potb, fpotb};
    for (int k = 0; k < nkind; ++k) {
        const int k0 = find(p.ids, fpair[k]);
        for (int c = 0; c < 4; ++c) {
            Spline sp;
            sp.x = R;
            sp.y.resize(nt);
            for (int j = 0; j < nt; ++j) sp.y[j] = p.col[(size_t)k0 * 4 + c][j] * scale[c];
            sp.fit_endslope4();
            for (int n = 1; n <= ntab; ++n) {
                const double t = (double)n * csiv, tt = t * t;
                double f = 0.0, df;
                if (tt <= trmax) sp.eval(tt, f, df); // :1664-1671
                out4[c][(size_t)(n - 1) * nkind + k] = f;
            }
        }
    }
    const double rhomx = (RHO[ne - 1] / (double)(ne - 1)) * (double)ne; // :1757
    double rhod = rhomx / (double)nembd;
    if (rhod <= 1.0e-64) rhod = 1.0;
    double *out2[2] = {fembd, dfembd};
    const double scale2[2] = {kEvErg, 1.0};
    for (int k = 0; k < nkind1; ++k) {
        const int k0 = find(e.ids, fpair1[k]);
        for (int c = 0; c < 2; ++c) {
            Spline sp;
            sp.x = RHO;
            sp.y.resize(ne);
            for (int j = 0; j < ne; ++j) sp.y[j] = e.col[(size_t)k0 * 2 + c][j] * scale2[c];
            sp.fit_endslope4();
            for (int n = 1; n <= nembd; ++n) {
                double f, df;
                sp.eval((double)(n - 1) * rhod, f, df);
                out2[c][(size_t)(n - 1) * nkind1 + k] = f;
            }
        }
    }
    for (int i = 0; i < ng; ++i) {
        for (int j = 0; j < ng; ++j) kpair[i + ng * j] = find(fpair, ptype[i + ng * j]) + 1;
        kembd[i] = find(fpair1, ptype[i + ng * i]) + 1;
    }
    if (pot_type_out) *pot_type_out = fs ? MDB_POT_FS : MDB_POT_EAM;
    if (nkind_out) *nkind_out = nkind;
    if (nkind1_out) *nkind1_out = nkind1;
    if (csi_out) *csi_out = csi;
    if (rhod_out) *rhod_out = rhod;
    return MDB_OK;
}

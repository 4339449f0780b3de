// mdb_lbfgs.cu -- limited-memory BFGS quench (BASELINE configs[3] "LBFGS quench").
//
// Reference: DO_LBFGSB_FORSTEPS_DEV, CommonGPU/MD_LBFGSScheme_GPU.F90:177-388.  There every iteration copies XP, FP and EPOT
// to the host, packs the free components into X and G, calls the serial Fortran routine SETULB (LIB/sor/f/LBFGSB/lbfgsb.f,
// L-BFGS-B with NBD = 0, i.e. without bounds) on ONE core over all 3N variables, wraps X into the box and copies it back.
// Here X, G, the search direction and the 2*M correction vectors never leave the device:
//   * one fused kernel per force evaluation forms G = -FP on the free components and reduces F = sum(EPOT), G.D and max|G|;
//   * the quasi-Newton direction -H G is formed in the span of {G, S_k, Y_k}: the inner products of the pairs (S_i.Y_k,
//     Y_i.Y_k, computed once when a pair enters the memory) and the 2*col products S_k.G, Y_k.G (one fused kernel) let the host
//     run the two-loop recursion on (2M+1) coefficients, and one more kernel combines the vectors -- two kernels and one
//     64-byte readback per direction instead of 2*col dependent dot products;
//   * the host keeps only scalars: the More'-Thuente line search (lnsrlb / dcsrch / dcstep of lbfgsb.f with its constants
//     ftol 1e-3, gtol 0.9, xtol 0.1, at most 20 evaluations), the convergence tests of mainlb and the SETULB call count the
//     reference's MXNUMSTEPS limits.
// For a problem without bounds SETULB's generalized Cauchy point is x and its subspace minimisation returns -H g of the
// compact L-BFGS matrix with B0 = theta*I; the two-loop recursion gives the same vector in exact arithmetic.  Parity is
// defined against the oracle's restatement (oracle/mdpscu_oracle.c, orc_md_lbfgsb); no reference L-BFGS output ships.
#include "mdb_internal.cuh"

#define LT 256
#define LB_MAXM 16

namespace {
struct LbScal { double f, gd, gmax, dtd, rr; double dots[3 * LB_MAXM]; int ticket, pad; };
struct LbCoef { double cg; double a[LB_MAXM], b[LB_MAXM]; int col, slot[LB_MAXM]; };

__device__ __forceinline__ double blk_sum(double v, double *sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < LT / 32; w++) r += sh[w];
    return r;
}
__device__ __forceinline__ double blk_max(double v, double *sh)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < LT / 32; w++) r = fmax(r, sh[w]);
    return r;
}
__device__ __forceinline__ bool lb_last(LbScal *S) // block-uniform: the last block to arrive finishes the reduction in parallel
{
    __shared__ int is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int t = atomicAdd(&S->ticket, 1);
        is_last = (t == (int)gridDim.x - 1) ? 1 : 0;
        if (is_last) { S->ticket = 0; __threadfence(); }
    }
    __syncthreads();
    return is_last != 0;
}
__device__ __forceinline__ double fin_sum(const double *part, int stride, int off, double *sh)
{
    double v = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += LT) v += __ldcg(part + (size_t)k * stride + off);
    return blk_sum(v, sh);
}
__device__ __forceinline__ double fin_max(const double *part, int stride, int off, double *sh)
{
    double v = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += LT) v = fmax(v, __ldcg(part + (size_t)k * stride + off));
    return blk_max(v, sh);
}

// free components (:213-243) and the unwrapped variable vector X = XP
__global__ void k_lb_init(int n, const double4 *__restrict__ pos, const int *__restrict__ statu, unsigned char *__restrict__ fre,
                          double *__restrict__ xl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int st = statu[i];
    const double4 p = pos[i];
    const double x[3] = {p.x, p.y, p.z};
    const bool ok = (st & ST_OUTOFBOX) == 0 && (st & ST_ACTIVE) == ST_ACTIVE;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        fre[i + (size_t)k * n] = (ok && (st & (ST_FIXPOSX << k)) == 0) ? 1 : 0;
        xl[i + (size_t)k * n] = x[k];
    }
}
// X = STP*D + T on the free components (lnsrlb :2384-2390); positions = X wrapped into the box (:268-330)
__global__ void k_lb_move(int n, double stp, const double *__restrict__ d, const double *__restrict__ t, double *__restrict__ xl,
                          const unsigned char *__restrict__ fre, double4 *__restrict__ pos, BoxParams box, float *__restrict__ dsr,
                          int *__restrict__ counters)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float d2 = 0.f;
    if (i < n) {
        double4 p = pos[i];
        double x[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const size_t o = i + (size_t)k * n;
            if (!fre[o]) continue;
            const double xn = __dadd_rn(__dmul_rn(stp, d[o]), t[o]);
            if (dsr) { const float u = dsr[o] + (float)(xn - xl[o]); dsr[o] = u; d2 += u * u; }
            xl[o] = xn;
            double w = xn;
            if (box.pd[k]) { if (w < box.lo[k]) w = __dadd_rn(w, box.size[k]); else if (w > box.up[k]) w = __dsub_rn(w, box.size[k]); }
            x[k] = w;
        }
        p.x = x[0]; p.y = x[1]; p.z = x[2];
        pos[i] = p;
    }
    if (dsr) {
        for (int off = 16; off > 0; off >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, off));
        if ((threadIdx.x & 31) == 0 && d2 > 0.f) atomicMax(&counters[CNT_D2MAX], __float_as_int(d2));
    }
}
// G = -FP on the free components (:335-337); F = sum EPOT; G.D; max|G|
__global__ void __launch_bounds__(LT) k_lb_grad(int n, const double *__restrict__ fp, const double *__restrict__ epot,
                                                const unsigned char *__restrict__ fre, double *__restrict__ g, const double *__restrict__ d,
                                                double *__restrict__ part, LbScal *S)
{
    __shared__ double sh[LT / 32];
    const size_t n3 = (size_t)n * 3;
    double f = 0.0, gd = 0.0, gm = 0.0;
    for (size_t o = blockIdx.x * (size_t)LT + threadIdx.x; o < n3; o += (size_t)gridDim.x * LT) {
        const double gv = fre[o] ? -fp[o] : 0.0;
        g[o] = gv;
        gd += gv * d[o];
        gm = fmax(gm, fabs(gv));
        if (o < (size_t)n) f += epot[o];
    }
    f = blk_sum(f, sh); gd = blk_sum(gd, sh); gm = blk_max(gm, sh);
    if (threadIdx.x == 0) { part[3 * blockIdx.x] = f; part[3 * blockIdx.x + 1] = gd; part[3 * blockIdx.x + 2] = gm; }
    if (lb_last(S)) {
        const double a = fin_sum(part, 3, 0, sh), b = fin_sum(part, 3, 1, sh), c = fin_max(part, 3, 2, sh);
        if (threadIdx.x == 0) { S->f = a; S->gd = b; S->gmax = c; }
    }
}
// up to 3*col inner products of v with stored vectors: dots[j] = A_j.v, dots[M+j] = B_j.v, dots[2M+j] = C_j.v (null = skip)
__global__ void __launch_bounds__(LT) k_lb_dots(size_t n3, int col, LbCoef K, const double *__restrict__ v, const double *__restrict__ v2,
                                                const double *__restrict__ ws, const double *__restrict__ wy, int mode,
                                                double *__restrict__ part, LbScal *S)
{
    // mode 0: dots[j] = S_j.v, dots[M+j] = Y_j.v              (v = G)
    // mode 1: dots[j] = S_j.v2 (v2 = new Y), dots[M+j] = Y_j.v (v = new S), dots[2M+j] = Y_j.v2
    __shared__ double sh[LT / 32];
    double acc[3 * LB_MAXM];
#pragma unroll
    for (int j = 0; j < 3 * LB_MAXM; j++) acc[j] = 0.0;
    for (size_t o = blockIdx.x * (size_t)LT + threadIdx.x; o < n3; o += (size_t)gridDim.x * LT) {
        const double a = v[o], b = v2 ? v2[o] : 0.0;
        for (int j = 0; j < col; j++) {
            const double s = ws[(size_t)K.slot[j] * n3 + o], y = wy[(size_t)K.slot[j] * n3 + o];
            if (mode == 0) { acc[j] += s * a; acc[LB_MAXM + j] += y * a; }
            else { acc[j] += s * b; acc[LB_MAXM + j] += y * a; acc[2 * LB_MAXM + j] += y * b; }
        }
    }
    const int nacc = (mode == 0) ? 2 : 3;
    for (int q = 0; q < nacc; q++)
        for (int j = 0; j < col; j++) {
            const double r = blk_sum(acc[q * LB_MAXM + j], sh);
            if (threadIdx.x == 0) part[(size_t)blockIdx.x * 3 * LB_MAXM + q * LB_MAXM + j] = r;
        }
    if (lb_last(S)) {
        for (int q = 0; q < nacc; q++)
            for (int j = 0; j < col; j++) {
                const double a = fin_sum(part, 3 * LB_MAXM, q * LB_MAXM + j, sh);
                if (threadIdx.x == 0) S->dots[q * LB_MAXM + j] = a;
            }
    }
}
// D = -(cg*G + sum_j a_j S_j + b_j Y_j); T = X; R = G; D.D and G.D
__global__ void __launch_bounds__(LT) k_lb_dir(size_t n3, LbCoef K, const double *__restrict__ g, const double *__restrict__ ws,
                                               const double *__restrict__ wy, const double *__restrict__ xl, double *__restrict__ d,
                                               double *__restrict__ t, double *__restrict__ r, double *__restrict__ part, LbScal *S)
{
    __shared__ double sh[LT / 32];
    double dtd = 0.0, gd = 0.0;
    for (size_t o = blockIdx.x * (size_t)LT + threadIdx.x; o < n3; o += (size_t)gridDim.x * LT) {
        const double gv = g[o];
        double v = K.cg * gv;
        for (int j = 0; j < K.col; j++) v += K.a[j] * ws[(size_t)K.slot[j] * n3 + o] + K.b[j] * wy[(size_t)K.slot[j] * n3 + o];
        v = -v;
        d[o] = v; t[o] = xl[o]; r[o] = gv;
        dtd += v * v; gd += gv * v;
    }
    dtd = blk_sum(dtd, sh); gd = blk_sum(gd, sh);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = dtd; part[2 * blockIdx.x + 1] = gd; }
    if (lb_last(S)) {
        const double a = fin_sum(part, 2, 0, sh), b = fin_sum(part, 2, 1, sh);
        if (threadIdx.x == 0) { S->dtd = a; S->gd = b; }
    }
}
// new pair (mainlb :822-836): Y = G - R into wy[slot], S = STP*D into ws[slot]; Y.Y
__global__ void __launch_bounds__(LT) k_lb_pair(size_t n3, double stp, const double *__restrict__ g, const double *__restrict__ r,
                                                const double *__restrict__ d, double *__restrict__ sdst, double *__restrict__ ydst,
                                                double *__restrict__ part, LbScal *S)
{
    __shared__ double sh[LT / 32];
    double rr = 0.0;
    for (size_t o = blockIdx.x * (size_t)LT + threadIdx.x; o < n3; o += (size_t)gridDim.x * LT) {
        const double y = __dsub_rn(g[o], r[o]);
        ydst[o] = y;
        sdst[o] = (stp == 1.0) ? d[o] : __dmul_rn(stp, d[o]);
        rr += y * y;
    }
    rr = blk_sum(rr, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = rr;
    if (lb_last(S)) {
        const double a = fin_sum(part, 1, 0, sh);
        if (threadIdx.x == 0) S->rr = a;
    }
}
__global__ void k_lb_restore(size_t n3, const double *__restrict__ t, const double *__restrict__ r, double *__restrict__ xl, double *__restrict__ g)
{
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < n3; o += (size_t)gridDim.x * blockDim.x) { xl[o] = t[o]; g[o] = r[o]; }
}

// ---- host scalars: the More'-Thuente search of lbfgsb.f (dcsrch :3280-3530, dcstep :3534-3760)
struct MoreThuente {
    bool brackt = false;
    int stage = 1;
    double ginit = 0, gtest = 0, gx = 0, gy = 0, finit = 0, fx = 0, fy = 0, stx = 0, sty = 0, stmin = 0, stmax = 0, width = 0, width1 = 0;
    static constexpr double ftol = 1.0e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0, stpmax = 1.0e10;
    enum { EVAL = 0, CONVERGED = 1, WARNING = 2, ERROR = -1 };

    int start(double f, double g, double stp)
    {
        if (stp < stpmin || stp > stpmax || g >= 0.0) return ERROR;
        brackt = false; stage = 1; finit = f; ginit = g; gtest = ftol * g;
        width = stpmax - stpmin; width1 = width / 0.5;
        stx = 0.0; fx = f; gx = g; sty = 0.0; fy = f; gy = g; stmin = 0.0; stmax = stp + 4.0 * stp;
        return EVAL;
    }
    static double cubic_gamma(double theta, double da, double db, bool clamp)
    {
        const double s = std::fmax(std::fabs(theta), std::fmax(std::fabs(da), std::fabs(db)));
        double t = (theta / s) * (theta / s) - (da / s) * (db / s);
        if (clamp && t < 0.0) t = 0.0;
        return s * std::sqrt(t);
    }
    void step(double &stp, double fp, double dp, double &fxr, double &dxr, double &fyr, double &dyr)
    {
        const double sgnd = dp * (dxr / std::fabs(dxr));
        double stpf;
        if (fp > fxr) { // higher value: the minimum is bracketed
            const double theta = 3.0 * (fxr - fp) / (stp - stx) + dxr + dp;
            double gamma = cubic_gamma(theta, dxr, dp, false);
            if (stp < stx) gamma = -gamma;
            const double r = ((gamma - dxr) + theta) / (((gamma - dxr) + gamma) + dp);
            const double stpc = stx + r * (stp - stx);
            const double stpq = stx + ((dxr / ((fxr - fp) / (stp - stx) + dxr)) / 2.0) * (stp - stx);
            stpf = (std::fabs(stpc - stx) < std::fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
            brackt = true;
        } else if (sgnd < 0.0) { // derivatives of opposite sign: bracketed
            const double theta = 3.0 * (fxr - fp) / (stp - stx) + dxr + dp;
            double gamma = cubic_gamma(theta, dxr, dp, false);
            if (stp > stx) gamma = -gamma;
            const double r = ((gamma - dp) + theta) / (((gamma - dp) + gamma) + dxr);
            const double stpc = stp + r * (stx - stp), stpq = stp + (dp / (dp - dxr)) * (stx - stp);
            stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
            brackt = true;
        } else if (std::fabs(dp) < std::fabs(dxr)) { // derivative shrinks
            const double theta = 3.0 * (fxr - fp) / (stp - stx) + dxr + dp;
            double gamma = cubic_gamma(theta, dxr, dp, true);
            if (stp > stx) gamma = -gamma;
            const double r = ((gamma - dp) + theta) / ((gamma + (dxr - dp)) + gamma);
            double stpc;
            if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
            else stpc = (stp > stx) ? stmax : stmin;
            const double stpq = stp + (dp / (dp - dxr)) * (stx - stp);
            if (brackt) {
                stpf = (std::fabs(stpc - stp) < std::fabs(stpq - stp)) ? stpc : stpq;
                stpf = (stp > stx) ? std::fmin(stp + 0.66 * (sty - stp), stpf) : std::fmax(stp + 0.66 * (sty - stp), stpf);
            } else {
                stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
                stpf = std::fmax(stmin, std::fmin(stmax, stpf));
            }
        } else { // derivative does not shrink
            if (brackt) {
                const double theta = 3.0 * (fp - fyr) / (sty - stp) + dyr + dp;
                double gamma = cubic_gamma(theta, dyr, dp, false);
                if (stp > sty) gamma = -gamma;
                const double r = ((gamma - dp) + theta) / (((gamma - dp) + gamma) + dyr);
                stpf = stp + r * (sty - stp);
            } else stpf = (stp > stx) ? stmax : stmin;
        }
        if (fp > fxr) { sty = stp; fyr = fp; dyr = dp; }
        else {
            if (sgnd < 0.0) { sty = stx; fyr = fxr; dyr = dxr; }
            stx = stp; fxr = fp; dxr = dp;
        }
        stp = stpf;
    }
    int next(double f, double g, double &stp)
    {
        const double ftest = finit + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        bool warn = false;
        if (brackt && (stp <= stmin || stp >= stmax)) warn = true;
        if (brackt && stmax - stmin <= xtol * stmax) warn = true;
        if (stp == stpmax && f <= ftest && g <= gtest) warn = true;
        if (stp == stpmin && (f > ftest || g >= gtest)) warn = true;
        if (f <= ftest && std::fabs(g) <= gtol * (-ginit)) return CONVERGED;
        if (warn) return WARNING;
        if (stage == 1 && f <= fx && f > ftest) {
            double fm = f - stp * gtest, fxm = fx - stx * gtest, fym = fy - sty * gtest, gm = g - gtest, gxm = gx - gtest, gym = gy - gtest;
            step(stp, fm, gm, fxm, gxm, fym, gym);
            fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
        } else step(stp, f, g, fx, gx, fy, gy);
        if (brackt) {
            if (std::fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
            width1 = width; width = std::fabs(sty - stx);
            stmin = std::fmin(stx, sty); stmax = std::fmax(stx, sty);
        } else { stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx); }
        stp = std::fmin(std::fmax(stp, stpmin), stpmax);
        if (brackt && ((stp <= stmin || stp >= stmax) || stmax - stmin <= xtol * stmax)) stp = stx;
        return EVAL;
    }
};
} // namespace

extern "C" int mdb_lbfgs(mdb_ctx *c, int mxnumsteps, int msave, double factr, double pgtol, int *iflag, int *nfg_out, int *niter_out)
{
    if (!c || mxnumsteps < 0 || msave < 1 || msave > LB_MAXM) return mdb_fail(c, MDB_ERR_ARG, "mdb_lbfgs: bad argument (1 <= LBFGS_MSave <= 16)");
    if (!c->has_box || !c->has_tables || !c->has_nlist || !c->list_valid)
        return mdb_fail(c, MDB_ERR_STATE, "mdb_lbfgs: box, tables and a valid neighbour list are required");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_lbfgs: not available in slab-decomposed runs yet");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    const int n = c->n;
    const size_t n3 = (size_t)n * 3;
    const int nblk = std::min(592, cdiv((long long)n3, LT));
    cudaStream_t st = c->stream;
    // ---- workspace: xl, g, d, t, r, ws[m], wy[m], partials, scalars, mask
    const size_t nd = (5 + 2 * (size_t)msave) * n3 + (size_t)nblk * 3 * LB_MAXM + 64;
    // the workspace stays with the context: allocation (pinned host memory in particular) costs far more than a quench
    if (c->lb_doubles < nd) {
        if (c->lb_buf) cudaFree(c->lb_buf);
        c->lb_buf = nullptr; c->lb_doubles = 0;
        if (cudaMalloc(&c->lb_buf, sizeof(double) * nd) != cudaSuccess) { cudaGetLastError(); return mdb_fail(c, MDB_ERR_NOMEM, "mdb_lbfgs: workspace"); }
        c->lb_doubles = nd;
    }
    if (c->lb_mask_n < n3) {
        if (c->lb_mask) cudaFree(c->lb_mask);
        c->lb_mask = nullptr; c->lb_mask_n = 0;
        if (cudaMalloc(&c->lb_mask, n3) != cudaSuccess) { cudaGetLastError(); return mdb_fail(c, MDB_ERR_NOMEM, "mdb_lbfgs: workspace"); }
        c->lb_mask_n = n3;
    }
    if (!c->lb_host && cudaMallocHost(&c->lb_host, sizeof(LbScal)) != cudaSuccess) { cudaGetLastError(); return mdb_fail(c, MDB_ERR_NOMEM, "mdb_lbfgs: workspace"); }
    double *buf = c->lb_buf;
    unsigned char *fre = c->lb_mask;
    LbScal *H = reinterpret_cast<LbScal *>(c->lb_host);
    double *xl = buf, *g = xl + n3, *d = g + n3, *t = d + n3, *r = t + n3, *ws = r + n3, *wy = ws + (size_t)msave * n3;
    double *part = wy + (size_t)msave * n3;
    LbScal *S = reinterpret_cast<LbScal *>(part + (size_t)nblk * 3 * LB_MAXM);
    auto cleanup = [&](int code) { cudaStreamSynchronize(st); return code; };
    auto peek = [&]() -> int {
        if (cudaGetLastError() != cudaSuccess) return MDB_ERR_CUDA;
        if (cudaMemcpyAsync(H, S, sizeof(LbScal), cudaMemcpyDeviceToHost, st) != cudaSuccess) return MDB_ERR_CUDA;
        return cudaStreamSynchronize(st) == cudaSuccess ? MDB_OK : MDB_ERR_CUDA;
    };
    cudaMemsetAsync(buf, 0, sizeof(double) * nd, st);
    k_lb_init<<<cdiv(n, LT), LT, 0, st>>>(n, c->pos, c->statu, fre, xl);
    c->launches_total += 1;

    const double epsmch = 2.220446049250313e-16, tol = factr * epsmch;
    // Gram entries of the stored pairs, indexed by memory slot: SY[i][k] = S_i.Y_k, YY[i][k] = Y_i.Y_k
    static thread_local double SY[LB_MAXM][LB_MAXM], YY[LB_MAXM][LB_MAXM];
    int col = 0, head = 0, iter = 0, nfg = 0, calls = 0, flag = 0, rc = 0;
    double theta = 1.0, f = 0.0, fold = 0.0, gd = 0.0, gdold = 0.0, dtd = 0.0, stp = 0.0;
    auto evaluate = [&]() -> int { // force + energy at the wrapped positions, then G, F, G.D, max|G|
        int e = mdb_force(c, MDB_FORCE | MDB_EPOT, nullptr);
        if (e < 0) return e;
        k_lb_grad<<<nblk, LT, 0, st>>>(n, c->fp, c->epot, fre, g, d, part, S);
        c->launches_total += 1;
        nfg++;
        if ((e = peek()) < 0) return e;
        f = H->f; gd = H->gd;
        return MDB_OK;
    };
    auto finish = [&](int code) {
        if (code >= 0) { // hm_XP1 = 0 (:363-364)
            cudaMemsetAsync(c->xp1, 0, sizeof(double) * n3, st);
            if (iflag) *iflag = flag;
            if (nfg_out) *nfg_out = nfg;
            if (niter_out) *niter_out = iter;
        }
        return cleanup(code < 0 ? mdb_fail(c, code, "mdb_lbfgs: CUDA error or force failure inside the iteration") : MDB_OK);
    };

    calls = 1; // 'START' -> 'FG_START'
    if (calls > mxnumsteps) { flag = 1; return finish(MDB_OK); }
    if ((rc = evaluate()) < 0) return finish(rc);
    if (H->gmax <= pgtol) return finish(MDB_OK);
    for (;;) {
        // ---- direction -H g: two-loop recursion on the coefficients of {G, S_k, Y_k}
        LbCoef K;
        memset(&K, 0, sizeof(K));
        K.col = col;
        for (int j = 0; j < col; j++) K.slot[j] = (head + j) % msave;
        if (col > 0) {
            k_lb_dots<<<nblk, LT, 0, st>>>(n3, col, K, g, nullptr, ws, wy, 0, part, S);
            c->launches_total += 1;
            if ((rc = peek()) < 0) return finish(rc);
        }
        {
            // q = cq*G + sum cy[k] Y_k ;  r = q/theta + sum cs[k] S_k      (k = position in the memory, oldest first)
            double cy[LB_MAXM] = {0}, cs[LB_MAXM] = {0}, alpha[LB_MAXM];
            for (int j = col - 1; j >= 0; j--) {
                const int p = K.slot[j];
                double sq = H->dots[j];                                  // S_j.G
                for (int k = 0; k < col; k++) sq += cy[k] * SY[p][K.slot[k]];
                alpha[j] = sq / SY[p][p];                                // rho_j = 1/(S_j.Y_j)
                cy[j] -= alpha[j];
            }
            for (int j = 0; j < col; j++) {
                const int p = K.slot[j];
                double yr = H->dots[LB_MAXM + j];                        // Y_j.G
                for (int k = 0; k < col; k++) yr += cy[k] * YY[p][K.slot[k]];
                yr /= theta;
                for (int k = 0; k < col; k++) yr += cs[k] * SY[K.slot[k]][p];
                const double beta = yr / SY[p][p];
                cs[j] += alpha[j] - beta;
            }
            K.cg = 1.0 / theta;
            for (int j = 0; j < col; j++) { K.a[j] = cs[j]; K.b[j] = cy[j] / theta; }
        }
        k_lb_dir<<<nblk, LT, 0, st>>>(n3, K, g, ws, wy, xl, d, t, r, part, S);
        c->launches_total += 1;
        if ((rc = peek()) < 0) return finish(rc);
        dtd = H->dtd; gd = H->gd;
        // ---- line search (lnsrlb)
        stp = (iter == 0) ? std::fmin(1.0 / std::sqrt(dtd), 1.0e10) : 1.0;
        fold = f;
        gdold = gd;
        bool failed = (gd >= 0.0); // "ascent direction in projection", info = -4
        MoreThuente ls;
        int ifun = 0;
        if (!failed) {
            int stt = ls.start(f, gd, stp);
            while (stt == MoreThuente::EVAL) {
                ifun++;
                if (ifun - 1 >= 20) { failed = true; break; }
                k_lb_move<<<cdiv(n, LT), LT, 0, st>>>(n, stp, d, t, xl, fre, c->pos, c->box, c->dsr, c->counters);
                c->launches_total += 1;
                calls++;
                if (calls > mxnumsteps) { flag = 1; return finish(MDB_OK); }
                if ((rc = evaluate()) < 0) return finish(rc);
                stt = ls.next(f, gd, stp);
            }
            if (stt == MoreThuente::ERROR) failed = true;
        }
        if (failed) { // mainlb :906-935
            k_lb_restore<<<nblk, LT, 0, st>>>(n3, t, r, xl, g);
            c->launches_total += 1;
            f = fold;
            if (col == 0) { iter++; break; } // ABNORMAL_TERMINATION_IN_LNSRCH
            col = 0; head = 0; theta = 1.0;
            continue;
        }
        // ---- 'NEW_X'
        iter++;
        calls++;
        if (calls > mxnumsteps || calls + 1 > mxnumsteps) { flag = 1; break; }
        if (H->gmax <= pgtol) break;
        if ((fold - f) <= tol * std::fmax(std::fabs(fold), std::fmax(std::fabs(f), 1.0))) break;
        // ---- update the memory (mainlb :822-862, matupd)
        {
            const double dr = (stp == 1.0) ? (gd - gdold) : (gd - gdold) * stp;
            const double ddum = (stp == 1.0) ? -gdold : -gdold * stp;
            if (dr > epsmch * ddum) {
                int slot;
                if (col < msave) { slot = (head + col) % msave; col++; }
                else { slot = head; head = (head + 1) % msave; }
                k_lb_pair<<<nblk, LT, 0, st>>>(n3, stp, g, r, d, ws + (size_t)slot * n3, wy + (size_t)slot * n3, part, S);
                LbCoef Q;
                memset(&Q, 0, sizeof(Q));
                Q.col = col;
                for (int j = 0; j < col; j++) Q.slot[j] = (head + j) % msave;
                k_lb_dots<<<nblk, LT, 0, st>>>(n3, col, Q, ws + (size_t)slot * n3, wy + (size_t)slot * n3, ws, wy, 1, part, S);
                c->launches_total += 2;
                if ((rc = peek()) < 0) return finish(rc);
                for (int j = 0; j < col; j++) {
                    const int p = Q.slot[j];
                    SY[p][slot] = H->dots[j];                 // S_p . Y_new
                    SY[slot][p] = H->dots[LB_MAXM + j];       // S_new . Y_p
                    YY[p][slot] = YY[slot][p] = H->dots[2 * LB_MAXM + j];
                }
                SY[slot][slot] = dr;                          // sy(col,col) = dr (matupd)
                YY[slot][slot] = H->rr;
                theta = H->rr / dr;
            }
        }
    }
    return finish(MDB_OK);
}

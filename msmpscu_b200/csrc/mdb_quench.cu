// mdb_quench.cu -- steepest-descent quench on the device (the quench PARREP / GMD
// "QUICKDAMP ST" run through the hot path).
//
// Reference: Do_Steepest0_Forsteps_DEV, CommonGPU/MD_SteepestScheme_GPU.F90:20-153 -- a Barzilai-Borwein step
// ALPHA = (DXP.DFP)/(DFP.DFP), DXP = ALPHA*FP capped at STEEPEST_MxStep, positions moved with the periodic wrap of
// AddBD_DevVec_DF_KERNEL0 (MSMLIB/sor/CommonGPU/MSM_MultiGPU_Basic.F90:5444-5476), stop when the largest move falls
// below STEEPEST_MiStep or the largest per-atom energy change below STEEPEST_MiDelE.  The reference reads every dot
// product and maximum back to the host and decides there (five blocking reductions per iteration); here the scalars,
// the step size and the "done" flag stay on the device, every kernel of later iterations (the force passes included)
// returns at once when the flag is set, and the host looks at the flag once per batch of iterations.
#include "mdb_internal.cuh"

struct QuenchScal {
    double dotdxdf, dotdf, maxmove, delepot, alpha, scale;
    int done, iflag, ticket, pending, iter, pad;
    // conjugate-gradient / line-search schemes
    double f0norm, pf0, pf1, step, coef, mf1;
};
static_assert(sizeof(QuenchScal) <= 128, "QuenchScal is mirrored in a 128-byte pinned host buffer");

#define QT 256

// block-level reduction helpers (deterministic: fixed tree inside a block, block partials summed in block order)
__device__ __forceinline__ double block_sum(double v, double *sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += sh[w];
    __syncthreads();
    return r; // valid on thread 0
}
__device__ __forceinline__ double block_max(double v, double *sh)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); w++) r = fmax(r, sh[w]);
    __syncthreads();
    return r;
}
// true on EVERY thread of the last block to arrive (block-uniform): that block finishes the reduction in parallel
__device__ __forceinline__ bool last_block(QuenchScal *S)
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
// second stage of a reduction, by all threads of one block: partial k of slot `off` sits at part[k*stride + off].
// Fixed order (thread t takes k = t, t+QT, ...; then the block tree), so results are reproducible run to run.
// (A single thread walking the partials pays one L2 round trip per partial: ~0.5 ms for 1024 blocks.)
__device__ __forceinline__ double final_sum(const double *part, int stride, int off, double *sh)
{
    double v = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) v += __ldcg(part + (size_t)k * stride + off);
    return block_sum(v, sh);
}
__device__ __forceinline__ double final_max(const double *part, int stride, int off, double *sh)
{
    double v = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) v = fmax(v, __ldcg(part + (size_t)k * stride + off));
    return block_max(v, sh);
}
__device__ __forceinline__ void set_scale(QuenchScal *S, double maxmove, double maxdis)
{
    S->maxmove = maxmove;
    S->scale = (maxmove > maxdis) ? maxdis / maxmove : 1.0; // DevMultiply_noshift(MAXDIS/MAXMOVE, dDXP) :74-75,108-110
}

// first step :70-86 : PreFP = FP, DXP = ALPHA*FP, MAXMOVE = max|DXP|
__global__ void __launch_bounds__(QT) k_sd_first(size_t n3, double alpha0, double maxdis, double mindis, const double *__restrict__ fp,
                                                 double *__restrict__ prefp, double *__restrict__ dxp, double *__restrict__ part,
                                                 QuenchScal *S)
{
    __shared__ double sh[QT / 32];
    double m = 0.0;
    for (size_t i = blockIdx.x * (size_t)QT + threadIdx.x; i < n3; i += (size_t)gridDim.x * QT) {
        const double f = fp[i];
        prefp[i] = f;
        const double d = alpha0 * f;
        dxp[i] = d;
        m = fmax(m, fabs(d));
    }
    m = block_max(m, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = m;
    if (last_block(S)) {
        const double mm = final_max(part, 1, 0, sh);
        if (threadIdx.x == 0) {
            set_scale(S, mm, maxdis);
            S->alpha = alpha0; S->pending = 0; S->iter = 0;
            if (!(mm > maxdis) && mm <= mindis) { S->done = 1; S->iflag = -1; } // converged at the first step :76-86
        }
    }
}

// XP = wrap(DXP*scale + XP) :90,112 ; also keeps the displacement-since-rebuild bound of the tiled passes current
__global__ void __launch_bounds__(QT) k_sd_apply(int n, double *__restrict__ dxp, double4 *__restrict__ pos, BoxParams box,
                                                 float *__restrict__ dsr, int *__restrict__ counters, const QuenchScal *S)
{
    if (S->done) return;
    const double sc = S->scale;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float d2 = 0.f;
    if (i < n) {
        double4 p = pos[i];
        double x[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const size_t o = i + (size_t)d * n;
            double dd = dxp[o];
            if (sc != 1.0) { dd = __dmul_rn(sc, dd); dxp[o] = dd; }
            double rt = __dadd_rn(dd, x[d]);
            const double lb = box.pd[d] ? box.lo[d] : -1.0e108, hb = box.pd[d] ? box.up[d] : 1.0e108; // :52-59
            if (rt > hb) rt = __dsub_rn(rt, __dsub_rn(hb, lb));
            else if (rt < lb) rt = __dadd_rn(rt, __dsub_rn(hb, lb));
            x[d] = rt;
            if (dsr) { const float t = dsr[o] + (float)dd; dsr[o] = t; d2 += t * t; }
        }
        p.x = x[0]; p.y = x[1]; p.z = x[2];
        pos[i] = p;
    }
    if (dsr) {
        for (int off = 16; off > 0; off >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, off));
        if ((threadIdx.x & 31) == 0 && d2 > 0.f) atomicMax(&counters[CNT_D2MAX], __float_as_int(d2));
    }
}

// after the move: the distance criterion :117-120 takes effect (the move itself is still applied, as in the reference)
__global__ void k_sd_mark(QuenchScal *S, int it)
{
    if (!S->done) {
        S->iter = it;
        if (S->pending) { S->done = 1; S->iflag = it; }
    }
}

// DFP = PreFP - FP ; DOTDXDF = DXP.DFP ; DOTDF = DFP.DFP ; ALPHA :97-104
__global__ void __launch_bounds__(QT) k_sd_dots(size_t n3, double alpha0, const double *__restrict__ prefp, const double *__restrict__ fp,
                                                const double *__restrict__ dxp, double *__restrict__ part, QuenchScal *S)
{
    if (S->done) return;
    __shared__ double sh[QT / 32];
    double a = 0.0, b = 0.0;
    for (size_t i = blockIdx.x * (size_t)QT + threadIdx.x; i < n3; i += (size_t)gridDim.x * QT) {
        const double dfp = prefp[i] - fp[i];
        a += dxp[i] * dfp;
        b += dfp * dfp;
    }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b; }
    if (last_block(S)) {
        const double sa = final_sum(part, 2, 0, sh), sb = final_sum(part, 2, 1, sh);
        if (threadIdx.x == 0) {
            S->dotdxdf = sa; S->dotdf = sb;
            double alpha = sa / sb;
            if (alpha < 0.0) alpha = alpha0;
            S->alpha = alpha;
        }
    }
}

// DXP = ALPHA*FP ; MAXMOVE = max|DXP| ; scale ; distance criterion armed :106-110,117
__global__ void __launch_bounds__(QT) k_sd_step(size_t n3, double maxdis, double mindis, const double *__restrict__ fp,
                                                double *__restrict__ dxp, double *__restrict__ part, QuenchScal *S)
{
    if (S->done) return;
    __shared__ double sh[QT / 32];
    const double alpha = S->alpha;
    double m = 0.0;
    for (size_t i = blockIdx.x * (size_t)QT + threadIdx.x; i < n3; i += (size_t)gridDim.x * QT) {
        const double d = alpha * fp[i];
        dxp[i] = d;
        m = fmax(m, fabs(d));
    }
    m = block_max(m, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = m;
    if (last_block(S)) {
        const double mm = final_max(part, 1, 0, sh);
        if (threadIdx.x == 0) {
            set_scale(S, mm, maxdis);
            S->pending = (mm <= mindis) ? 1 : 0;
        }
    }
}

// DELEPOT = max|EPOT - EPOT0| ; energy criterion :122-128
__global__ void __launch_bounds__(QT) k_sd_echeck(int n, double minepot, const double *__restrict__ epot, const double *__restrict__ epot0,
                                                  double *__restrict__ part, QuenchScal *S, int it)
{
    if (S->done) return;
    __shared__ double sh[QT / 32];
    double m = 0.0;
    for (size_t i = blockIdx.x * (size_t)QT + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * QT) m = fmax(m, fabs(epot[i] - epot0[i]));
    m = block_max(m, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = m;
    if (last_block(S)) {
        const double mm = final_max(part, 1, 0, sh);
        if (threadIdx.x == 0) {
            S->delepot = mm;
            if (mm <= minepot) { S->done = 1; S->iflag = it; }
        }
    }
}

// EPOT0 = EPOT ; PreFP = FP :129-132 (also the EPOT0 copy of the first step :87-88 with n3 = 0)
__global__ void __launch_bounds__(QT) k_sd_save(size_t n3, int n, const double *__restrict__ fp, double *__restrict__ prefp,
                                                const double *__restrict__ epot, double *__restrict__ epot0, const QuenchScal *S)
{
    if (S->done) return;
    for (size_t i = blockIdx.x * (size_t)QT + threadIdx.x; i < n3; i += (size_t)gridDim.x * QT) prefp[i] = fp[i];
    for (size_t i = blockIdx.x * (size_t)QT + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * QT) epot0[i] = epot[i];
}

static int steepest1(mdb_ctx *c, int mxnumsteps, double maxdis, double mindis, int *iflag, double *delepot);

extern "C" int mdb_steepest(mdb_ctx *c, int mxnumsteps, int meth, double alpha, double maxdis, double mindis, double minepot,
                            int *iflag, double *maxmove, double *delepot)
{
    if (!c || mxnumsteps < 0) return mdb_fail(c, MDB_ERR_ARG, "mdb_steepest: bad argument");
    if (!c->has_box || !c->has_tables || !c->has_nlist || !c->list_valid)
        return mdb_fail(c, MDB_ERR_STATE, "mdb_steepest: box, tables and a valid neighbour list are required");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_steepest: not available in slab-decomposed runs yet");
    if (meth & MDB_QUENCH_LSEARCH) { // Do_Steepest_Forsteps_DEV :263-290 dispatches on CP_DAMPSCHEME_LSEARCH
        if (maxmove) *maxmove = 0.0;
        return steepest1(c, mxnumsteps, maxdis, mindis, iflag, delepot);
    }
    CUDA_TRY(c, cudaSetDevice(c->dev));
    const int n = c->n;
    const size_t n3 = (size_t)n * 3;
    const int nblk = std::min(1024, cdiv((long long)n3, QT));
    if (c->q_n != n) {
        if (c->q_buf) cudaFree(c->q_buf);
        c->q_buf = nullptr; c->q_n = 0;
        CUDA_TRY(c, cudaMalloc(&c->q_buf, sizeof(double) * (2 * n3 + n + 2 * 1024 + 16)));
        c->q_n = n;
    }
    if (!c->q_host) CUDA_TRY(c, cudaMallocHost(&c->q_host, 128));
    double *prefp = c->q_buf, *dxp = prefp + n3, *epot0 = dxp + n3, *part = epot0 + n;
    QuenchScal *S = reinterpret_cast<QuenchScal *>(part + 2 * 1024);
    QuenchScal *H = reinterpret_cast<QuenchScal *>(c->q_host);
    cudaStream_t st = c->stream;
    CUDA_TRY(c, cudaMemsetAsync(S, 0, sizeof(QuenchScal), st));
    int rc;
    // ---- first step :70-90.  The energies of a configuration and the forces the NEXT iteration starts from belong to the
    // same positions, so they are evaluated together (MDB_FORCE | MDB_EPOT: the energy pass also delivers dF/drho and no
    // separate density pass runs); PreFP is saved before that evaluation overwrites FP.
    if ((rc = mdb_force(c, MDB_FORCE | MDB_EPOT, nullptr)) < 0) return rc;
    c->launches_total += 1;
    k_sd_first<<<nblk, QT, 0, st>>>(n3, alpha, maxdis, mindis, c->fp, prefp, dxp, part, S);
    c->skip_flag = &S->done;
    auto finish = [&](int code) { c->skip_flag = nullptr; return code; };
    c->launches_total += 2;
    k_sd_save<<<nblk, QT, 0, st>>>(0, n, c->fp, prefp, c->epot, epot0, S);
    k_sd_apply<<<cdiv(n, QT), QT, 0, st>>>(n, dxp, c->pos, c->box, c->dsr, c->counters, S);
    if ((rc = mdb_force(c, MDB_FORCE, nullptr)) < 0) return finish(rc);          // forces of iteration 1
    // ---- iterations :93-133, the host looks at the flag once per batch
    const int batch = 8;
    int it = 1;
    CUDA_TRY(c, cudaMemcpyAsync(H, S, sizeof(QuenchScal), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    while (!H->done && it <= mxnumsteps) {
        for (int b = 0; b < batch && it <= mxnumsteps; b++, it++) {
            c->launches_total += 7;
            k_sd_dots<<<nblk, QT, 0, st>>>(n3, alpha, prefp, c->fp, dxp, part, S);
            k_sd_step<<<nblk, QT, 0, st>>>(n3, maxdis, mindis, c->fp, dxp, part, S);
            k_sd_save<<<nblk, QT, 0, st>>>(n3, 0, c->fp, prefp, c->epot, epot0, S);    // PreFP = FP :131
            k_sd_apply<<<cdiv(n, QT), QT, 0, st>>>(n, dxp, c->pos, c->box, c->dsr, c->counters, S);
            k_sd_mark<<<1, 1, 0, st>>>(S, it);
            if ((rc = mdb_force(c, MDB_FORCE | MDB_EPOT, nullptr)) < 0) return finish(rc); // EPOT :122 and the next FP :95
            k_sd_echeck<<<nblk, QT, 0, st>>>(n, minepot, c->epot, epot0, part, S, it);
            k_sd_save<<<nblk, QT, 0, st>>>(0, n, c->fp, prefp, c->epot, epot0, S);     // EPOT0 = EPOT :129
        }
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaMemcpyAsync(H, S, sizeof(QuenchScal), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaStreamSynchronize(st));
    }
    if (iflag) *iflag = H->done ? H->iflag : 0;
    if (maxmove) *maxmove = H->maxmove;
    if (delepot) *delepot = H->delepot;
    return finish(MDB_OK);
}


// =====================================================================================================================
// Conjugate gradient (Do_CG0/CG1_Forsteps_DEV, CommonGPU/MD_CGScheme_GPU.F90:16-276) and steepest descent with line
// search (Do_Steepest1_Forsteps_DEV, CommonGPU/MD_SteepestScheme_GPU.F90:157-260).
//
// The reference expresses these with whole-array device operations (MSM_MultiGPU_Basic.F90: DevMultiply, DevAdd_shift,
// DevMinus, DevDot, DevMaxAbsval), each dot product a blocking reduction finished on the host.  Here an iteration is a
// handful of fused kernels whose reductions end in the last block to arrive, which also does the scalar update that
// follows in the reference (secant step, clamp, Polak-Ribiere factor), so STEPSIZE, PF0/PF1, F0NORM and GAMA never leave
// the device.  CG0 has no data-dependent branch inside an iteration: it runs exactly like mdb_steepest (stop flag on the
// device, every later kernel -- force passes included -- returns at once, host looks once per batch).  The line-search
// variants decide after every force evaluation whether |STEPSIZE| <= MINDIS; the host reads one 128-byte record per
// force evaluation for that (the reference: two blocking reductions + host arithmetic per evaluation).
// =====================================================================================================================
enum { QD_PF1_CG = 0, QD_PF1_SD = 1, QD_MF1 = 2, QD_NORM = 3 };

// sum a[i] * (b[i] - bsub[i]) (bsub may be null) and the scalar update `mode` that follows it in the reference
__global__ void __launch_bounds__(QT) k_q_dot(size_t n3, const double *__restrict__ a, const double *__restrict__ b,
                                              const double *__restrict__ bsub, double *__restrict__ part, QuenchScal *S, int mode,
                                              double maxdis)
{
    if (S->done) return;
    __shared__ double sh[QT / 32];
    double v = 0.0;
    for (size_t i = blockIdx.x * (size_t)QT + threadIdx.x; i < n3; i += (size_t)gridDim.x * QT) {
        const double bb = bsub ? __dsub_rn(b[i], bsub[i]) : b[i];
        v += a[i] * bb;
    }
    v = block_sum(v, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    if (last_block(S)) {
        const double sum = final_sum(part, 1, 0, sh);
        if (threadIdx.x != 0) return;
        if (mode == QD_MF1) S->mf1 = sum;
        else if (mode == QD_NORM) S->f0norm = sum;
        else {
            // STEPSIZE = -STEPSIZE*PF1/(PF1 - PF0), capped at MAXDIS with its sign (MD_CGScheme_GPU.F90:71-74, MD_SteepestScheme_GPU.F90:221-224)
            const double pf1 = sum;
            double step = -S->step * pf1 / (pf1 - S->pf0);
            if (fabs(step) > maxdis) step = maxdis * fabs(step) / step;
            S->pf1 = pf1;
            S->step = step;
            S->coef = (mode == QD_PF1_CG) ? step / sqrt(S->f0norm) : step;
            S->pf0 = pf1; // PF0 = PF1 of the line search (:214 / :232); recomputed with the next direction otherwise
        }
    }
}

// DXP = coef*Dir ; XP = wrap(DXP + XP)   (DevMultiply_noshift + DevAdd_shift)
__global__ void __launch_bounds__(QT) k_q_move(int n, const double *__restrict__ dir, double4 *__restrict__ pos, BoxParams box,
                                               float *__restrict__ dsr, int *__restrict__ counters, const QuenchScal *S)
{
    if (S->done) return;
    const double coef = S->coef;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float d2 = 0.f;
    if (i < n) {
        double4 p = pos[i];
        double x[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const size_t o = i + (size_t)d * n;
            const double dd = __dmul_rn(coef, dir[o]);
            double rt = __dadd_rn(dd, x[d]);
            const double lb = box.pd[d] ? box.lo[d] : -1.0e108, hb = box.pd[d] ? box.up[d] : 1.0e108;
            if (rt > hb) rt = __dsub_rn(rt, __dsub_rn(hb, lb));
            else if (rt < lb) rt = __dadd_rn(rt, __dsub_rn(hb, lb));
            x[d] = rt;
            if (dsr) { const float t = dsr[o] + (float)dd; dsr[o] = t; d2 += t * t; }
        }
        p.x = x[0]; p.y = x[1]; p.z = x[2];
        pos[i] = p;
    }
    if (dsr) {
        for (int off = 16; off > 0; off >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, off));
        if ((threadIdx.x & 31) == 0 && d2 > 0.f) atomicMax(&counters[CNT_D2MAX], __float_as_int(d2));
    }
}

// new direction.  first != 0: Dir = FP (:56-57).  Otherwise GAMA = MF1NORM/F0NORM, Dir = FP + GAMA*Dir (:93-99).
// Then F0 = FP, F0NORM = F0.F0, PF0 = FP.Dir, and the trial step of the next iteration STEPSIZE = MAXDIS/sqrt(F0NORM) (:62).
__global__ void __launch_bounds__(QT) k_q_cgdir(size_t n3, int first, const double *__restrict__ fp, double *__restrict__ f0,
                                                double *__restrict__ dir, double *__restrict__ part, QuenchScal *S, double maxdis,
                                                double eps, int it)
{
    if (S->done) return;
    __shared__ double sh[QT / 32];
    const double gama = first ? 0.0 : S->mf1 / S->f0norm;
    double a = 0.0, b = 0.0;
    for (size_t i = blockIdx.x * (size_t)QT + threadIdx.x; i < n3; i += (size_t)gridDim.x * QT) {
        const double f = fp[i];
        const double dnew = first ? f : __dadd_rn(f, __dmul_rn(gama, dir[i]));
        dir[i] = dnew;
        f0[i] = f;
        a += f * f;
        b += f * dnew;
    }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b; }
    if (last_block(S)) {
        const double sa = final_sum(part, 2, 0, sh), sb = final_sum(part, 2, 1, sh);
        if (threadIdx.x == 0) {
            S->f0norm = sa; S->pf0 = sb;
            S->step = maxdis / sqrt(sa);
            S->coef = S->step;
            if (sa <= eps) { S->done = 1; S->iflag = it; } // :46-54 (it = -1), :103-105
        }
    }
}

// Dir = (1/sqrt(FP.FP))*FP (DevNormalize), PF0 = FP.Dir, STEPSIZE = MAXDIS   (MD_SteepestScheme_GPU.F90:207-212)
__global__ void __launch_bounds__(QT) k_q_sd1dir(size_t n3, const double *__restrict__ fp, double *__restrict__ dir,
                                                 double *__restrict__ part, QuenchScal *S, double maxdis)
{
    __shared__ double sh[QT / 32];
    const double sc = 1.0 / sqrt(S->f0norm);
    double b = 0.0;
    for (size_t i = blockIdx.x * (size_t)QT + threadIdx.x; i < n3; i += (size_t)gridDim.x * QT) {
        const double f = fp[i], dnew = __dmul_rn(sc, f);
        dir[i] = dnew;
        b += f * dnew;
    }
    b = block_sum(b, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = b;
    if (last_block(S)) {
        const double sb = final_sum(part, 1, 0, sh);
        if (threadIdx.x == 0) { S->pf0 = sb; S->step = maxdis; S->coef = maxdis; }
    }
}

namespace {
struct QuenchWork {
    double *f0, *dir, *epot0, *part;
    QuenchScal *S, *H;
    int nblk;
    size_t n3;
};
int quench_setup(mdb_ctx *c, const char *who, QuenchWork &w)
{
    if (!c->has_box || !c->has_tables || !c->has_nlist || !c->list_valid)
        return mdb_fail(c, MDB_ERR_STATE, "quench: box, tables and a valid neighbour list are required");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "quench: not available in slab-decomposed runs yet");
    (void)who;
    CUDA_TRY(c, cudaSetDevice(c->dev));
    const int n = c->n;
    w.n3 = (size_t)n * 3;
    w.nblk = std::min(1024, cdiv((long long)w.n3, QT));
    if (c->q_n != n) {
        if (c->q_buf) cudaFree(c->q_buf);
        c->q_buf = nullptr; c->q_n = 0;
        CUDA_TRY(c, cudaMalloc(&c->q_buf, sizeof(double) * (2 * w.n3 + n + 2 * 1024 + 16)));
        c->q_n = n;
    }
    if (!c->q_host) CUDA_TRY(c, cudaMallocHost(&c->q_host, 128));
    w.f0 = c->q_buf; w.dir = w.f0 + w.n3; w.epot0 = w.dir + w.n3; w.part = w.epot0 + n;
    w.S = reinterpret_cast<QuenchScal *>(w.part + 2 * 1024);
    w.H = reinterpret_cast<QuenchScal *>(c->q_host);
    CUDA_TRY(c, cudaMemsetAsync(c->q_buf, 0, sizeof(double) * (2 * w.n3 + n + 2 * 1024 + 16), c->stream));
    return MDB_OK;
}
int quench_peek(mdb_ctx *c, QuenchWork &w)
{
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(w.H, w.S, sizeof(QuenchScal), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MDB_OK;
}
} // namespace

extern "C" int mdb_cg(mdb_ctx *c, int mxnumsteps, int meth, double maxdis, double mindis, double minepot, int *iflag, double *delepot)
{
    if (!c || mxnumsteps < 0) return mdb_fail(c, MDB_ERR_ARG, "mdb_cg: bad argument");
    QuenchWork w;
    int rc = quench_setup(c, "mdb_cg", w);
    if (rc < 0) return rc;
    const int n = c->n;
    const double eps = 1.0e-64;
    cudaStream_t st = c->stream;
    QuenchScal *S = w.S;
    auto finish = [&](int code) { c->skip_flag = nullptr; return code; };
    auto move = [&]() { k_q_move<<<cdiv(n, QT), QT, 0, st>>>(n, w.dir, c->pos, c->box, c->dsr, c->counters, S); c->launches_total += 1; };
    auto dot = [&](const double *b, const double *bsub, int mode) {
        k_q_dot<<<w.nblk, QT, 0, st>>>(w.n3, c->fp, b, bsub, w.part, S, mode, maxdis); c->launches_total += 1; };
    // ---- :50-61
    if ((rc = mdb_force(c, MDB_FORCE | MDB_EPOT, nullptr)) < 0) return rc;
    k_sd_save<<<w.nblk, QT, 0, st>>>(0, n, c->fp, w.f0, c->epot, w.epot0, S);
    k_q_cgdir<<<w.nblk, QT, 0, st>>>(w.n3, 1, c->fp, w.f0, w.dir, w.part, S, maxdis, eps, -1);
    c->launches_total += 2;
    c->skip_flag = &S->done;
    if ((rc = quench_peek(c, w)) < 0) return finish(rc);
    const bool ls = (meth & MDB_QUENCH_LSEARCH) != 0;
    // the tail of an iteration, :80-105 / :235-260: energy criterion, then the new direction
    auto tail = [&](int it) -> int {
        int r;
        // the energies (:80) and the forces of the new direction (:88) belong to the same positions: one fused evaluation
        if ((r = mdb_force(c, MDB_FORCE | MDB_EPOT, nullptr)) < 0) return r;
        k_sd_echeck<<<w.nblk, QT, 0, st>>>(n, minepot, c->epot, w.epot0, w.part, S, it);
        k_sd_save<<<w.nblk, QT, 0, st>>>(0, n, c->fp, w.f0, c->epot, w.epot0, S);
        c->launches_total += 2;
        dot(c->fp, w.f0, QD_MF1);
        k_q_cgdir<<<w.nblk, QT, 0, st>>>(w.n3, 0, c->fp, w.f0, w.dir, w.part, S, maxdis, eps, it);
        c->launches_total += 1;
        return MDB_OK;
    };
    int iter = 0;
    if (!ls) {
        const int batch = 4;
        iter = 1;
        while (!w.H->done && iter <= mxnumsteps) {
            for (int b = 0; b < batch && iter <= mxnumsteps; b++, iter++) {
                move();                                                     // trial step :62-64
                if ((rc = mdb_force(c, MDB_FORCE, nullptr)) < 0) return finish(rc);
                dot(w.dir, nullptr, QD_PF1_CG);                             // PF1 and the secant step :67-75
                move();                                                     // :76
                if ((rc = tail(iter)) < 0) return finish(rc);
            }
            if ((rc = quench_peek(c, w)) < 0) return finish(rc);
        }
    } else {
        while (!w.H->done && iter <= mxnumsteps) {
            move();                                                         // :203-205
            if ((rc = mdb_force(c, MDB_FORCE, nullptr)) < 0) return finish(rc);
            dot(w.dir, nullptr, QD_PF1_CG);                                 // PF1 (+ the step of the first inner pass)
            iter++;
            while (iter <= mxnumsteps) {                                    // :211-233
                move();
                iter++;
                if ((rc = quench_peek(c, w)) < 0) return finish(rc);
                if (fabs(w.H->step) <= mindis) break;
                if ((rc = mdb_force(c, MDB_FORCE, nullptr)) < 0) return finish(rc);
                dot(w.dir, nullptr, QD_PF1_CG);
            }
            if ((rc = tail(iter)) < 0) return finish(rc);
            if ((rc = quench_peek(c, w)) < 0) return finish(rc);
        }
    }
    if (iflag) *iflag = w.H->done ? w.H->iflag : 0;
    if (delepot) *delepot = w.H->delepot;
    return finish(MDB_OK);
}

// Do_Steepest1_Forsteps_DEV.  MINEPOT is the reference's literal 0.001 eV (:178).
static int steepest1(mdb_ctx *c, int mxnumsteps, double maxdis, double mindis, int *iflag, double *delepot)
{
    QuenchWork w;
    int rc = quench_setup(c, "mdb_steepest", w);
    if (rc < 0) return rc;
    const int n = c->n;
    const double minepot = 0.001 * 1.60219e-12;
    cudaStream_t st = c->stream;
    QuenchScal *S = w.S;
    int iter = 0;
    if ((rc = quench_peek(c, w)) < 0) return rc;
    while (!w.H->done && iter <= mxnumsteps) {
        if ((rc = mdb_force(c, MDB_FORCE | MDB_EPOT, nullptr)) < 0) return rc;                   // :203-204
        k_sd_save<<<w.nblk, QT, 0, st>>>(0, n, c->fp, w.f0, c->epot, w.epot0, S);               // EPOT0 :205
        k_q_dot<<<w.nblk, QT, 0, st>>>(w.n3, c->fp, c->fp, nullptr, w.part, S, QD_NORM, maxdis);
        k_q_sd1dir<<<w.nblk, QT, 0, st>>>(w.n3, c->fp, w.dir, w.part, S, maxdis);               // :206-212
        c->launches_total += 3;
        while (iter <= mxnumsteps) {                                                            // :213-233
            k_q_move<<<cdiv(n, QT), QT, 0, st>>>(n, w.dir, c->pos, c->box, c->dsr, c->counters, S);
            if ((rc = mdb_force(c, MDB_FORCE, nullptr)) < 0) return rc;
            k_q_dot<<<w.nblk, QT, 0, st>>>(w.n3, c->fp, w.dir, nullptr, w.part, S, QD_PF1_SD, maxdis);
            c->launches_total += 2;
            iter++;
            if ((rc = quench_peek(c, w)) < 0) return rc;
            if (fabs(w.H->step) <= mindis) break;
        }
        if ((rc = mdb_force(c, MDB_EPOT, nullptr)) < 0) return rc;
        k_sd_echeck<<<w.nblk, QT, 0, st>>>(n, minepot, c->epot, w.epot0, w.part, S, iter);
        c->launches_total += 1;
        if ((rc = quench_peek(c, w)) < 0) return rc;
    }
    if (iflag) *iflag = w.H->done ? w.H->iflag : 0;
    if (delepot) *delepot = w.H->delepot;
    return MDB_OK;
}


// Do_DynDamp_Forsteps_DEV, CommonGPU/MD_DiffScheme_GPU.F90:1809-1860 (CP_DAMPSCHEME_DYN of Do_Damp): damped dynamics on the
// current list -- DAMPING, predictor, force, energy criterion max|EPOT - EPOT0| <= STEEPEST_MiDelE, corrector -- with the stop
// flag on the device: every kernel of a later iteration returns at once, the host looks once per 8 iterations.
extern "C" int mdb_dyndamp(mdb_ctx *c, int mxnumsteps, double h, double minepot, int *iflag, double *delepot)
{
    if (!c || mxnumsteps < 0 || !(h > 0.0)) return mdb_fail(c, MDB_ERR_ARG, "mdb_dyndamp: bad argument");
    QuenchWork w;
    int rc = quench_setup(c, "mdb_dyndamp", w);
    if (rc < 0) return rc;
    const int n = c->n;
    cudaStream_t st = c->stream;
    QuenchScal *S = w.S;
    auto finish = [&](int code) { c->skip_flag = nullptr; return code; };
    if ((rc = mdb_force(c, MDB_EPOT, nullptr)) < 0) return rc;                                   // :1826-1827
    k_sd_save<<<w.nblk, QT, 0, st>>>(0, n, c->fp, w.f0, c->epot, w.epot0, S);
    c->launches_total += 1;
    c->skip_flag = &S->done;
    if ((rc = quench_peek(c, w)) < 0) return finish(rc);
    int it = 1;
    while (!w.H->done && it <= mxnumsteps) {
        for (int b = 0; b < 8 && it <= mxnumsteps; b++, it++) {
            if ((rc = mdb_damping(c)) < 0) return finish(rc);                                    // :1830-1832
            if ((rc = mdb_predict(c, h)) < 0) return finish(rc);                                 // :1833-1836
            if ((rc = mdb_force(c, MDB_FORCE | MDB_EPOT, nullptr)) < 0) return finish(rc);        // :1839-1840
            k_sd_echeck<<<w.nblk, QT, 0, st>>>(n, minepot, c->epot, w.epot0, w.part, S, it);     // :1841-1845
            k_sd_save<<<w.nblk, QT, 0, st>>>(0, n, c->fp, w.f0, c->epot, w.epot0, S);            // :1846
            c->launches_total += 2;
            if ((rc = mdb_correct(c, h)) < 0) return finish(rc);                                 // :1847-1851
        }
        if ((rc = quench_peek(c, w)) < 0) return finish(rc);
    }
    if (iflag) *iflag = w.H->done ? w.H->iflag : 0;
    if (delepot) *delepot = w.H->delepot;
    return finish(MDB_OK);
}

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
};

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
// true on thread 0 of the last block to arrive
__device__ __forceinline__ bool last_block(QuenchScal *S)
{
    __shared__ int is_last;
    if (threadIdx.x == 0) {
        __threadfence();
        const int t = atomicAdd(&S->ticket, 1);
        is_last = (t == (int)gridDim.x - 1) ? 1 : 0;
        if (is_last) { S->ticket = 0; __threadfence(); }
    }
    __syncthreads();
    return is_last != 0 && threadIdx.x == 0;
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
        double mm = 0.0;
        for (int b = 0; b < (int)gridDim.x; b++) mm = fmax(mm, ((volatile double *)part)[b]);
        set_scale(S, mm, maxdis);
        S->alpha = alpha0; S->pending = 0; S->iter = 0;
        if (!(mm > maxdis) && mm <= mindis) { S->done = 1; S->iflag = -1; } // converged at the first step :76-86
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
        double sa = 0.0, sb = 0.0;
        for (int k = 0; k < (int)gridDim.x; k++) { sa += ((volatile double *)part)[2 * k]; sb += ((volatile double *)part)[2 * k + 1]; }
        S->dotdxdf = sa; S->dotdf = sb;
        double alpha = sa / sb;
        if (alpha < 0.0) alpha = alpha0;
        S->alpha = alpha;
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
        double mm = 0.0;
        for (int b = 0; b < (int)gridDim.x; b++) mm = fmax(mm, ((volatile double *)part)[b]);
        set_scale(S, mm, maxdis);
        S->pending = (mm <= mindis) ? 1 : 0;
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
        double mm = 0.0;
        for (int b = 0; b < (int)gridDim.x; b++) mm = fmax(mm, ((volatile double *)part)[b]);
        S->delepot = mm;
        if (mm <= minepot) { S->done = 1; S->iflag = it; }
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

extern "C" int mdb_steepest(mdb_ctx *c, int mxnumsteps, int meth, double alpha, double maxdis, double mindis, double minepot,
                            int *iflag, double *maxmove, double *delepot)
{
    if (!c || mxnumsteps < 0) return mdb_fail(c, MDB_ERR_ARG, "mdb_steepest: bad argument");
    if (!c->has_box || !c->has_tables || !c->has_nlist || !c->list_valid)
        return mdb_fail(c, MDB_ERR_STATE, "mdb_steepest: box, tables and a valid neighbour list are required");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_steepest: not available in slab-decomposed runs yet");
    if (meth & MDB_QUENCH_LSEARCH)
        return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_steepest: the line-search variant (Do_Steepest1_Forsteps_DEV) is not implemented yet");
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
    // ---- first step :70-90
    if ((rc = mdb_force(c, MDB_FORCE, nullptr)) < 0) return rc;
    c->launches_total += 1;
    k_sd_first<<<nblk, QT, 0, st>>>(n3, alpha, maxdis, mindis, c->fp, prefp, dxp, part, S);
    c->skip_flag = &S->done;
    auto finish = [&](int code) { c->skip_flag = nullptr; return code; };
    if ((rc = mdb_force(c, MDB_EPOT, nullptr)) < 0) return finish(rc);
    c->launches_total += 2;
    k_sd_save<<<nblk, QT, 0, st>>>(0, n, c->fp, prefp, c->epot, epot0, S);
    k_sd_apply<<<cdiv(n, QT), QT, 0, st>>>(n, dxp, c->pos, c->box, c->dsr, c->counters, S);
    // ---- iterations :93-133, the host looks at the flag once per batch
    const int batch = 8;
    int it = 1;
    CUDA_TRY(c, cudaMemcpyAsync(H, S, sizeof(QuenchScal), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    while (!H->done && it <= mxnumsteps) {
        for (int b = 0; b < batch && it <= mxnumsteps; b++, it++) {
            if ((rc = mdb_force(c, MDB_FORCE, nullptr)) < 0) return finish(rc);
            c->launches_total += 4;
            k_sd_dots<<<nblk, QT, 0, st>>>(n3, alpha, prefp, c->fp, dxp, part, S);
            k_sd_step<<<nblk, QT, 0, st>>>(n3, maxdis, mindis, c->fp, dxp, part, S);
            k_sd_apply<<<cdiv(n, QT), QT, 0, st>>>(n, dxp, c->pos, c->box, c->dsr, c->counters, S);
            k_sd_mark<<<1, 1, 0, st>>>(S, it);
            if ((rc = mdb_force(c, MDB_EPOT, nullptr)) < 0) return finish(rc);
            c->launches_total += 2;
            k_sd_echeck<<<nblk, QT, 0, st>>>(n, minepot, c->epot, epot0, part, S, it);
            k_sd_save<<<nblk, QT, 0, st>>>(n3, n, c->fp, prefp, c->epot, epot0, S);
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

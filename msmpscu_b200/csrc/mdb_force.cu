// mdb_force.cu -- the GENERIC two-pass tabulated EAM / Finnis-Sinclair force path.
//
// One thread per atom over the column-major INDI list, exactly the traversal of the
// reference kernels (CommonGPU/MD_EAM_ForceTable_GPU.F90:369-548 PRECALFOR, :646-829
// CALFORCE, :1014-1289 CALPTENSOR, :1470-1636 CALEPOT; FS twins in MD_FS_ForceTable_GPU.F90),
// but sized for B200 (grid covers all atoms; the reference launches a fixed 128x256 grid),
// with positions+density packed in one 32-byte record, (value, forward-difference) table
// pairs fetched with one 16-byte load, and a deterministic two-stage virial reduction.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: arithmetic is un-fused and in the
// reference's source order, IEEE sqrt and division, so results agree with the CPU oracle
// to the last bit or two.  It is the fallback for configurations the tiled fast path
// (mdb_force_tiled.cu) does not cover and the on-device cross-check for it at sizes where
// the CPU oracle is too slow.
#include "mdb_internal.cuh"

struct ForceParams {
    int n;
    BoxParams box;
    int pot_type, ng, ntab, nembd;
    double csi, rhod, ru2max;
    const double2 *potr, *fpotr, *potb, *fpotb, *fembd, *dfembd;
    const int *skip; // device flag: the kernel returns at once when it is set (converged quench iterations)
    int *counters;
    int identity;   // BOXSHAPE is the unit matrix
    double bs[9];   // BOXSHAPE, column-major
    int kpair[MDB_MXGROUP * MDB_MXGROUP];
    int kembd[MDB_MXGROUP];
};

__device__ __forceinline__ double3 ld_xyz(const double4 *__restrict__ pos, int j)
{
    // 16 B + 8 B: never touches .w (den), which pass 1 writes concurrently
    const double2 xy = __ldg(reinterpret_cast<const double2 *>(pos + j));
    const double z = __ldg(reinterpret_cast<const double *>(pos + j) + 2);
    return make_double3(xy.x, xy.y, z);
}

__device__ __forceinline__ void min_image(double &s, double size, double half, int pd)
{
    // if(IFPD.GT.0 .AND. DABS(SEP).GT.HB) SEP = SEP - DSIGN(B,SEP)   :500-510
    if (pd > 0 && fabs(s) > half) s = s - copysign(size, s);
}

// DXYZ = BOXSHAPE * SEP (MD_EAM_ForceTable_GPU.F90:515-517); column-major bs
__device__ __forceinline__ void box_shape(const ForceParams &P, double sx, double sy, double sz, double &dx, double &dy, double &dz)
{
    dx = sx; dy = sy; dz = sz;
    if (!P.identity) {
        dx = P.bs[0] * sx + P.bs[3] * sy + P.bs[6] * sz;
        dy = P.bs[1] * sx + P.bs[4] * sy + P.bs[7] * sz;
        dz = P.bs[2] * sx + P.bs[5] * sy + P.bs[8] * sz;
    }
}

__device__ __forceinline__ double lerp_tab(const double2 *__restrict__ t, int stride, int k, int kk, double dk)
{
    kk = min(max(kk, 0), stride - 1);
    const double2 e = __ldg(t + (size_t)k * stride + kk);
    return e.x + dk * e.y; // T(KK) + DK*(T(KK+1)-T(KK))
}

// rho beyond the table (KK > NEMBD: close collisions) reads past the arrays in the reference; here the row is clamped to the
// zero pad and the event is counted (mdb_embed_overruns) so that the caller can see it
__device__ __forceinline__ double embed_lookup(const double2 *__restrict__ t, int stride, int k, double rho, double rhod,
                                               int *__restrict__ counters)
{
    const double sk = rho / rhod + 1.0;  // :538
    const int kk = (int)(sk + 0.000001); // :539
    if (kk > stride - 2 && counters) atomicAdd(&counters[CNT_RHO_OVER], 1);
    return lerp_tab(t, stride, k, kk, sk - (double)kk);
}

// ---- pass 1: rho_i -> DEN(i) (stored in pos[i].w)
__global__ void __launch_bounds__(128)
k_pass1_generic(ForceParams P, double4 *__restrict__ pos, const int *__restrict__ ityp, const int *__restrict__ statu,
                const int *__restrict__ kvois, const int *__restrict__ indi)
{
    if (P.skip && *P.skip) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    double den0 = 0.0;
    if ((statu[i] & ST_ACTIVE) == ST_ACTIVE) {
        const double3 pi = ld_xyz(pos, i);
        const int ti = ityp[i] - 1, kv = kvois[i];
        for (int w = 0; w < kv; w++) {
            const int j = indi[i + (size_t)w * P.n] - 1;
            const double3 pj = ld_xyz(pos, j);
            double sx = pi.x - pj.x, sy = pi.y - pj.y, sz = pi.z - pj.z;
            min_image(sx, P.box.size[0], P.box.half[0], P.box.pd[0]);
            min_image(sy, P.box.size[1], P.box.half[1], P.box.pd[1]);
            min_image(sz, P.box.size[2], P.box.half[2], P.box.pd[2]);
            double dx, dy, dz;
            box_shape(P, sx, sy, sz, dx, dy, dz);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 <= P.ru2max) { // :522
                const int kt = (P.ng == 1) ? P.kpair[0] : P.kpair[ti + P.ng * (ityp[j] - 1)];
                const double r = sqrt(r2);
                const double sk = sqrt(r) * P.csi;
                const int kk = (int)sk;
                den0 = den0 + lerp_tab(P.potb, P.ntab + 2, kt, kk, sk - (double)kk); // :530
            }
        }
        if (P.pot_type == MDB_POT_FS) {
            if (den0 > 0.0) den0 = -0.5 / sqrt(den0); // MD_FS_ForceTable_GPU.F90:497-507
        } else if (den0 > 0.0) {                      // :535-541
            den0 = embed_lookup(P.dfembd, P.nembd + 2, P.kembd[ti], den0, P.rhod, P.counters);
        }
    }
    reinterpret_cast<double *>(pos + i)[3] = den0; // DEN(IC+IA0) = DEN0 :544
}

// ---- pass 2: forces (+ virial partials per block when VIR)
template <bool VIR>
__global__ void __launch_bounds__(128)
k_pass2_generic(ForceParams P, const double4 *__restrict__ pos, const int *__restrict__ ityp,
                const int *__restrict__ statu, const int *__restrict__ kvois, const int *__restrict__ indi,
                double *__restrict__ fp, double *__restrict__ vpart)
{
    if (P.skip && *P.skip) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    double v[9];
    if (VIR) {
#pragma unroll
        for (int q = 0; q < 9; q++) v[q] = 0.0;
    }
    if (i < P.n && (statu[i] & ST_ACTIVE) == ST_ACTIVE) {
        const double4 pi = pos[i];
        const int ti = ityp[i] - 1, kv = kvois[i];
        const double denki = pi.w;
        for (int w = 0; w < kv; w++) {
            const int j = indi[i + (size_t)w * P.n] - 1;
            const double4 pj = pos[j];
            double sx = pi.x - pj.x, sy = pi.y - pj.y, sz = pi.z - pj.z;
            min_image(sx, P.box.size[0], P.box.half[0], P.box.pd[0]);
            min_image(sy, P.box.size[1], P.box.half[1], P.box.pd[1]);
            min_image(sz, P.box.size[2], P.box.half[2], P.box.pd[2]);
            // the force-only kernel ignores BOXSHAPE (:775); the virial variant applies it (:1171-1174)
            double dx = sx, dy = sy, dz = sz;
            if (VIR) box_shape(P, sx, sy, sz, dx, dy, dz);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 <= P.ru2max) {
                int k0 = P.kpair[0], k1 = k0;
                if (P.ng > 1) {
                    const int tj = ityp[j] - 1;
                    k0 = P.kpair[ti + P.ng * tj];
                    k1 = P.kpair[tj + P.ng * ti];
                }
                const double r = sqrt(r2);
                const double sk = sqrt(r) * P.csi;
                const int kk = (int)sk;
                const double dk = sk - (double)kk;
                const double denkj = pj.w;
                // :811-813
                double fortot = lerp_tab(P.fpotr, P.ntab + 2, k0, kk, dk) / r2 +
                                (lerp_tab(P.fpotb, P.ntab + 2, k0, kk, dk) * denki +
                                 lerp_tab(P.fpotb, P.ntab + 2, k1, kk, dk) * denkj) / r;
                fx = fx + fortot * sx;
                fy = fy + fortot * sy;
                fz = fz + fortot * sz;
                if (VIR) { // :1222-1232
                    fortot = fortot * 0.5;
                    v[0] += dx * dx * fortot; v[3] += dx * dy * fortot; v[6] += dx * dz * fortot;
                    v[1] += dy * dx * fortot; v[4] += dy * dy * fortot; v[7] += dy * dz * fortot;
                    v[2] += dz * dx * fortot; v[5] += dz * dy * fortot; v[8] += dz * dz * fortot;
                }
            }
        }
    }
    if (i < P.n) {
        fp[i] = fx;
        fp[i + (size_t)P.n] = fy;
        fp[i + 2 * (size_t)P.n] = fz;
    }
    if (VIR) {
        __shared__ double sv[4][9];
        const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
#pragma unroll
        for (int q = 0; q < 9; q++) {
            double x = v[q];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
            if (lane == 0) sv[wib][q] = x;
        }
        __syncthreads();
        if (threadIdx.x < 9) {
            double x = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) x += sv[w][threadIdx.x];
            vpart[(size_t)blockIdx.x * 9 + threadIdx.x] = x;
        }
    }
}

// deterministic final reduction of the per-block virial partials; out[9] /= nbox (:1462-1463)
__global__ void k_virial_reduce(int nblk, const double *__restrict__ vpart, double *__restrict__ out, double inv_nbox)
{
    __shared__ double s[256];
    for (int q = 0; q < 9; q++) {
        double x = 0.0;
        for (int b = threadIdx.x; b < nblk; b += blockDim.x) x += vpart[(size_t)b * 9 + q];
        s[threadIdx.x] = x;
        __syncthreads();
        for (int off = blockDim.x / 2; off > 0; off >>= 1) {
            if ((int)threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[q] = s[0] * inv_nbox;
        __syncthreads();
    }
}

// ---- per-atom potential energy
__global__ void __launch_bounds__(128)
k_epot_generic(ForceParams P, const double4 *__restrict__ pos, const int *__restrict__ ityp, const int *__restrict__ statu,
               const int *__restrict__ kvois, const int *__restrict__ indi, double *__restrict__ epot)
{
    if (P.skip && *P.skip) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    double er0 = 0.0, den0 = 0.0;
    if ((statu[i] & ST_ACTIVE) == ST_ACTIVE) {
        const double3 pi = ld_xyz(pos, i);
        const int ti = ityp[i] - 1, kv = kvois[i];
        for (int w = 0; w < kv; w++) {
            const int j = indi[i + (size_t)w * P.n] - 1;
            const double3 pj = ld_xyz(pos, j);
            double sx = pi.x - pj.x, sy = pi.y - pj.y, sz = pi.z - pj.z;
            min_image(sx, P.box.size[0], P.box.half[0], P.box.pd[0]);
            min_image(sy, P.box.size[1], P.box.half[1], P.box.pd[1]);
            min_image(sz, P.box.size[2], P.box.half[2], P.box.pd[2]);
            double dx, dy, dz;
            box_shape(P, sx, sy, sz, dx, dy, dz);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 <= P.ru2max) {
                const int kt = (P.ng == 1) ? P.kpair[0] : P.kpair[ti + P.ng * (ityp[j] - 1)];
                const double r = sqrt(r2);
                const double sk = sqrt(r) * P.csi;
                const int kk = (int)sk;
                const double dk = sk - (double)kk;
                er0 = er0 + lerp_tab(P.potr, P.ntab + 2, kt, kk, dk) / r; // :1622
                den0 = den0 + lerp_tab(P.potb, P.ntab + 2, kt, kk, dk);   // :1623
            }
        }
        if (P.pot_type == MDB_POT_FS) den0 = -sqrt(den0); // MD_FS_ForceTable_GPU.F90:1606
        else den0 = embed_lookup(P.fembd, P.nembd + 2, P.kembd[ti], den0, P.rhod, nullptr); // :1628-1631 (no rho>0 guard)
    }
    epot[i] = er0 + den0; // :1633
}

// ---- per-atom virial tensor: CAL_EAM_AtomicStress_KERNEL, CommonGPU/MD_EAM_ForceTable_GPU.F90:1775-1925 (pCalAVStress).
// AP(i, 1..9) = sum_j DXYZ_a * DXYZ_b * FORTOT in the order 11,12,13,21,...,33; the FULL pair term goes to atom i (no 1/2).
__global__ void __launch_bounds__(128)
k_avstress_generic(ForceParams P, const double4 *__restrict__ pos, const int *__restrict__ ityp, const int *__restrict__ statu,
                   const int *__restrict__ kvois, const int *__restrict__ indi, double *__restrict__ ap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    double p[9];
#pragma unroll
    for (int q = 0; q < 9; q++) p[q] = 0.0;
    if ((statu[i] & ST_ACTIVE) == ST_ACTIVE) {
        const double4 pi = pos[i];
        const int ti = ityp[i] - 1, kv = kvois[i];
        const double denki = pi.w;
        for (int w = 0; w < kv; w++) {
            const int j = indi[i + (size_t)w * P.n] - 1;
            const double4 pj = pos[j];
            double sx = pi.x - pj.x, sy = pi.y - pj.y, sz = pi.z - pj.z;
            min_image(sx, P.box.size[0], P.box.half[0], P.box.pd[0]);
            min_image(sy, P.box.size[1], P.box.half[1], P.box.pd[1]);
            min_image(sz, P.box.size[2], P.box.half[2], P.box.pd[2]);
            double dx, dy, dz;
            box_shape(P, sx, sy, sz, dx, dy, dz); // :1861-1864
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 <= P.ru2max) {
                int k0 = P.kpair[0], k1 = k0;
                if (P.ng > 1) {
                    const int tj = ityp[j] - 1;
                    k0 = P.kpair[ti + P.ng * tj];
                    k1 = P.kpair[tj + P.ng * ti];
                }
                const double r = sqrt(r2);
                const double sk = sqrt(r) * P.csi;
                const int kk = (int)sk;
                const double dk = sk - (double)kk;
                const double fortot = lerp_tab(P.fpotr, P.ntab + 2, k0, kk, dk) / r2 +
                                      (lerp_tab(P.fpotb, P.ntab + 2, k0, kk, dk) * denki +
                                       lerp_tab(P.fpotb, P.ntab + 2, k1, kk, dk) * pj.w) / r; // :1877-1880
                p[0] = p[0] + dx * dx * fortot; p[1] = p[1] + dx * dy * fortot; p[2] = p[2] + dx * dz * fortot;
                p[3] = p[3] + dy * dx * fortot; p[4] = p[4] + dy * dy * fortot; p[5] = p[5] + dy * dz * fortot;
                p[6] = p[6] + dz * dx * fortot; p[7] = p[7] + dz * dy * fortot; p[8] = p[8] + dz * dz * fortot;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 9; q++) ap[i + (size_t)q * P.n] = p[q];
}

// final sum of nblk per-block virial partials (c->vpart) -> vt[9], divided by the number of boxes (:1462-1463)
int mdb_virial_finish(mdb_ctx *c, int nblk, double *vt)
{
    {
        ProfScope ps(c, MDB_K_OTHER);
        k_virial_reduce<<<1, 256, 0, c->stream>>>(nblk, c->vpart, c->vpart + (size_t)nblk * 9, 1.0 / (double)c->nbox);
    }
    CUDA_TRY(c, cudaGetLastError());
    if (vt) {
        CUDA_TRY(c, cudaMemcpyAsync(vt, c->vpart + (size_t)nblk * 9, sizeof(double) * 9, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return MDB_OK;
}

static void fill_params(mdb_ctx *c, ForceParams &P);
int mdb_avstress_generic(mdb_ctx *c, double *d_ap)
{
    ForceParams P;
    fill_params(c, P);
    P.skip = nullptr;
    ProfScope ps(c, MDB_K_OTHER);
    k_avstress_generic<<<cdiv(c->n, 128), 128, 0, c->stream>>>(P, c->pos, c->ityp, c->statu, c->kvois, c->indi, d_ap);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

static void fill_params(mdb_ctx *c, ForceParams &P)
{
    const TableSet &t = c->tab;
    P.n = c->n; P.box = c->box;
    P.pot_type = t.pot_type; P.ng = c->ng; P.ntab = t.ntab; P.nembd = t.nembd;
    P.csi = t.csi; P.rhod = t.rhod; P.ru2max = t.ru2max;
    P.skip = c->skip_flag;
    P.counters = c->counters;
    P.identity = c->shape_identity ? 1 : 0;
    for (int i = 0; i < 9; i++) P.bs[i] = c->boxshape[i];
    P.potr = t.potr; P.fpotr = t.fpotr; P.potb = t.potb; P.fpotb = t.fpotb; P.fembd = t.fembd; P.dfembd = t.dfembd;
    for (int i = 0; i < MDB_MXGROUP * MDB_MXGROUP; i++) P.kpair[i] = t.kpair[i];
    for (int i = 0; i < MDB_MXGROUP; i++) P.kembd[i] = t.kembd[i];
}

int mdb_force_generic(mdb_ctx *c, unsigned flags, double *vt)
{
    ForceParams P;
    fill_params(c, P);
    const int n = c->n, nb = cdiv(n, 128);
    cudaStream_t st = c->stream;
    if (flags & (MDB_FORCE | MDB_VIRIAL | MDB_DEN)) {
        ProfScope ps(c, MDB_K_PASS1);
        k_pass1_generic<<<nb, 128, 0, st>>>(P, c->pos, c->ityp, c->statu, c->kvois, c->indi);
    }
    if (flags & MDB_VIRIAL) {
        if (c->vpart_n < nb + 2) {
            if (c->vpart) cudaFree(c->vpart);
            c->vpart = nullptr;
            CUDA_TRY(c, cudaMalloc(&c->vpart, sizeof(double) * 9 * (size_t)(nb + 2)));
            c->vpart_n = nb + 2;
        }
        {
            ProfScope ps(c, MDB_K_PASS2);
            k_pass2_generic<true><<<nb, 128, 0, st>>>(P, c->pos, c->ityp, c->statu, c->kvois, c->indi, c->fp, c->vpart);
        }
        {
            ProfScope ps(c, MDB_K_OTHER);
            k_virial_reduce<<<1, 256, 0, st>>>(nb, c->vpart, c->vpart + (size_t)nb * 9, 1.0 / (double)c->nbox);
        }
        if (vt) {
            CUDA_TRY(c, cudaMemcpyAsync(vt, c->vpart + (size_t)nb * 9, sizeof(double) * 9, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(c, cudaStreamSynchronize(st));
        }
    } else if (flags & MDB_FORCE) {
        ProfScope ps(c, MDB_K_PASS2);
        k_pass2_generic<false><<<nb, 128, 0, st>>>(P, c->pos, c->ityp, c->statu, c->kvois, c->indi, c->fp, nullptr);
    }
    if (flags & MDB_EPOT) {
        ProfScope ps(c, MDB_K_EPOT);
        k_epot_generic<<<nb, 128, 0, st>>>(P, c->pos, c->ityp, c->statu, c->kvois, c->indi, c->epot);
    }
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

// mdb_thermal.cu -- Monte-Carlo thermalisation of the velocities.
//
// Reference: Thermalizing_MC_KERNEL / Thermalizing_MC_DEV, CommonGPU/MD_DiffScheme_GPU.F90:1608-1805.  Every free velocity
// component of an ACTIVE atom is redrawn as V0*sqrt(-ln Z1)*cos(2 pi Z2), V0 = sqrt(2 TI kB / m), Z uniform in (0,1];
// inactive atoms get zero velocity; afterwards the mass-weighted mean velocity of each box is subtracted from ALL atoms of
// the box (the reference does that on the host after copying XP1 out and back, :1782-1802).
//
// The reference draws Z from per-thread cuRAND XORWOW states seeded from the host generator (MSM_MultiGPU_Basic.F90:661-750):
// which number an atom receives depends on the launch geometry, the device count and the cell order.  Here the uniforms are
// a pure function of (seed, draw counter, ORIGINAL atom id, component) through the counter-based generator Philox4x32-10
// (Salmon, Moraes, Dror, Shaw, SC'11; the algorithm cuRAND ships as CURAND_RNG_PSEUDO_PHILOX4_32_10): a trajectory does not
// depend on the number of GPUs, on the sort order or on the grid size, and the CPU oracle reproduces the integers exactly.
// The sequence is therefore NOT the reference's XORWOW sequence -- only the distribution and the per-box momentum removal
// are common; tests check the integers against the published known-answer vector, the velocities against the oracle and
// the statistics against Maxwell-Boltzmann.
#include "mdb_internal.cuh"

namespace {
struct U4 { unsigned x, y, z, w; };
__host__ __device__ inline U4 philox4x32_10(U4 c, unsigned k0, unsigned k1)
{
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)M0 * c.x, p1 = (unsigned long long)M1 * c.z;
        const unsigned hi0 = (unsigned)(p0 >> 32), lo0 = (unsigned)p0, hi1 = (unsigned)(p1 >> 32), lo1 = (unsigned)p1;
        c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += W0; k1 += W1;
    }
    return c;
}
__device__ __forceinline__ double u01(unsigned x) { return ((double)x + 0.5) * 2.3283064365386963e-10; } // (0,1), 2^-32 grid
} // namespace

#define TT 256
__global__ void __launch_bounds__(TT) k_thermalize(int n, double *__restrict__ xp1, const int *__restrict__ statu,
                                                   const int *__restrict__ ityp, const int *__restrict__ gid, MassParams M, double ti,
                                                   unsigned seed_lo, unsigned seed_hi, unsigned draw)
{
    const int i = blockIdx.x * TT + threadIdx.x;
    if (i >= n) return;
    const int st = statu[i];
    if ((st & ST_ACTIVE) != ST_ACTIVE) { // :1664-1668
        xp1[i] = 0.0; xp1[i + (size_t)n] = 0.0; xp1[i + 2 * (size_t)n] = 0.0;
        return;
    }
    const unsigned orig = (unsigned)gid[i];
    const U4 a = philox4x32_10(U4{orig, draw, 0u, 0u}, seed_lo, seed_hi), b = philox4x32_10(U4{orig, draw, 1u, 0u}, seed_lo, seed_hi);
    const unsigned z[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
    const double v0 = sqrt(__dmul_rn(__dmul_rn(2.0, ti), KB_CGS) / M.cm[ityp[i] - 1]); // DSQRT(C_TWO*TI*CP_KB/CM) :1644
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if ((st & (ST_FIXVELX << d)) == 0 && (st & (ST_FIXPOSX << d)) == 0) {
            const double z1 = u01(z[2 * d]), z2 = u01(z[2 * d + 1]);
            xp1[i + (size_t)d * n] = __dmul_rn(__dmul_rn(v0, sqrt(-log(z1))), cos(__dmul_rn(6.283185307179586, z2)));
        }
    }
}

// block b = box b: WT = sum m, VT = sum m*v/WT over the box's atoms in ORIGINAL order (:1786-1795)
__global__ void __launch_bounds__(TT) k_box_com(int n, int napb, const double *__restrict__ xp1, const int *__restrict__ ityp,
                                                const int *__restrict__ gidinv, MassParams M, double *__restrict__ vt)
{
    __shared__ double sh[4][TT / 32];
    __shared__ double wt_s;
    const int b = blockIdx.x;
    auto reduce = [&](double v, int slot) {
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) sh[slot][threadIdx.x >> 5] = v;
    };
    double w = 0.0;
    for (int o = threadIdx.x; o < napb; o += TT) w += M.cm[ityp[gidinv[(size_t)b * napb + o] - 1] - 1];
    reduce(w, 0);
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < TT / 32; k++) t += sh[0][k]; wt_s = t; }
    __syncthreads();
    const double wt = wt_s;
    double s[3] = {0.0, 0.0, 0.0};
    for (int o = threadIdx.x; o < napb; o += TT) {
        const int i = gidinv[(size_t)b * napb + o] - 1;
        const double m = M.cm[ityp[i] - 1];
#pragma unroll
        for (int d = 0; d < 3; d++) s[d] += m * xp1[i + (size_t)d * n] / wt;
    }
    for (int d = 0; d < 3; d++) reduce(s[d], d + 1);
    __syncthreads();
    if (threadIdx.x < 3) { double t = 0.0; for (int k = 0; k < TT / 32; k++) t += sh[threadIdx.x + 1][k]; vt[3 * b + threadIdx.x] = t; }
}
__global__ void k_sub_com(int n, int napb, double *__restrict__ xp1, const int *__restrict__ gid, const double *__restrict__ vt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = (gid[i] - 1) / napb;
#pragma unroll
    for (int d = 0; d < 3; d++) xp1[i + (size_t)d * n] = __dsub_rn(xp1[i + (size_t)d * n], vt[3 * b + d]);
}

extern "C" int mdb_thermalize(mdb_ctx *c, double ti, unsigned long long seed, unsigned draw)
{
    if (!c || !(ti >= 0.0)) return mdb_fail(c, MDB_ERR_ARG, "mdb_thermalize: bad argument");
    if (!c->has_box) return mdb_fail(c, MDB_ERR_STATE, "mdb_thermalize: mdb_box_set first");
    if (c->dd_on) return mdb_fail(c, MDB_ERR_UNSUPPORTED, "mdb_thermalize: not available in slab-decomposed runs yet");
    CUDA_TRY(c, cudaSetDevice(c->dev));
    const int nb = c->nbox, n = c->n;
    if (c->vpart_n < nb + 2) {
        if (c->vpart) cudaFree(c->vpart);
        c->vpart = nullptr;
        CUDA_TRY(c, cudaMalloc(&c->vpart, sizeof(double) * 9 * (size_t)(nb + 2)));
        c->vpart_n = nb + 2;
    }
    {
        ProfScope ps(c, MDB_K_OTHER, 3);
        k_thermalize<<<cdiv(n, TT), TT, 0, c->stream>>>(n, c->xp1, c->statu, c->ityp, c->gid, c->mass, ti, (unsigned)seed,
                                                        (unsigned)(seed >> 32), draw);
        k_box_com<<<nb, TT, 0, c->stream>>>(n, c->napb, c->xp1, c->ityp, c->gidinv, c->mass, c->vpart);
        k_sub_com<<<cdiv(n, TT), TT, 0, c->stream>>>(n, c->napb, c->xp1, c->gid, c->vpart);
    }
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

// the integers behind draw `draw` of atom `orig_id` (1-based), for known-answer and oracle tests: out[0..7]
extern "C" int mdb_thermalize_bits(unsigned long long seed, unsigned draw, unsigned orig_id, unsigned *out)
{
    if (!out) return MDB_ERR_ARG;
    const U4 a = philox4x32_10(U4{orig_id, draw, 0u, 0u}, (unsigned)seed, (unsigned)(seed >> 32)),
             b = philox4x32_10(U4{orig_id, draw, 1u, 0u}, (unsigned)seed, (unsigned)(seed >> 32));
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    return MDB_OK;
}
// raw generator (counter and key as given), for the published known-answer vectors
extern "C" int mdb_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4])
{
    if (!ctr || !key || !out) return MDB_ERR_ARG;
    const U4 r = philox4x32_10(U4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    return MDB_OK;
}

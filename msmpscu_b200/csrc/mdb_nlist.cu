// mdb_nlist.cu -- neighbour-list kernels.
//
// Path A (this file, "cell-warp"): one warp per cell, the 27-cell candidate stream staged
// through a warp-private shared-memory tile.  It reproduces the reference list EXACTLY --
// same members, same order, same truncation -- because membership is decided by the same
// fp32 arithmetic in the same order as Cal_NeighboreList_Kernel2C
// (CommonGPU/MD_NeighborsList_GPU.F90:907-1199):
//     POS  = (float)XP_i                                   (:1080-1082, real(KINDSF) locals)
//     SPOS = (float)(XP_j + (double)(float)shift)          (:1100-1102, CXYZ is real(KINDSF))
//     SEP  = POS - SPOS ; D = BOXSHAPE_f * SEP ; D.D <= (float)(NB_RM^2)   (:1109-1123)
// every fp32 operation through __f*_rn intrinsics so nvcc cannot contract them into FMAs.
// Output layout is the reference's INDI(N,mxKVOIS) column-major (atom index contiguous).
#include "mdb_internal.cuh"

// scan order of the 27 cells, :218-220
__constant__ int c_nix[27] = {0, -1, -1, -1, 0, 0, -1, 1, -1, 0, 1, -1, 0, 1, 1, 1, 1, 0, 0, -1, 1, -1, 0, 1, -1, 0, 1};
__constant__ int c_niy[27] = {0, 0, -1, 1, 1, 0, 0, 0, -1, -1, -1, 1, 1, 1, 0, 1, -1, -1, 0, 0, 0, -1, -1, -1, 1, 1, 1};
__constant__ int c_niz[27] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1};

struct NlistParams {
    int n, nc, nc0, ncx, ncy, ncz, ng, mxkvois;
    int pd[3];
    float fbs[3];     // (float)BOXSIZE : CXYZ values
    float bsh[9];     // (float)BOXSHAPE, column-major
    int identity;
    float rm2[MDB_MXGROUP * MDB_MXGROUP];
};

#define NL_WARPS 4

__global__ void __launch_bounds__(NL_WARPS * 32)
k_nlist_cellwarp(NlistParams P, const double4 *__restrict__ pos, const int *__restrict__ ityp,
                 const int *__restrict__ nac, const int *__restrict__ naac, const int *__restrict__ ia1th,
                 int *__restrict__ kvois, int *__restrict__ indi, int *__restrict__ counters)
{
    __shared__ float4 tile[NL_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ic0 = blockIdx.x * NL_WARPS + wib;
    if (ic0 >= P.nc) return;
    const int na = nac[ic0];
    if (na <= 0) return;        // :1018
    if (naac[ic0] <= 0) {       // cells without ACTIVE atoms are skipped :981-982 (their atoms keep no list: KVOIS = 0 here, where the
                                // reference leaves the counts of an earlier build behind; such atoms never read their list)
        for (int a = lane; a < na; a += 32) kvois[ia1th[ic0] - 1 + a] = 0;
        return;
    }
    const int is0 = ic0 / P.nc0, icl = ic0 - is0 * P.nc0;
    const int ncxy = P.ncx * P.ncy;
    const int iz0 = icl / ncxy, iy0 = (icl - iz0 * ncxy) / P.ncx, ix0 = icl - iz0 * ncxy - iy0 * P.ncx;
    const int a0 = ia1th[ic0] - 1;

    for (int ab = 0; ab < na; ab += 32) {
        const bool valid = ab + lane < na;
        const int ia = a0 + ab + lane;
        float p1 = 0.f, p2 = 0.f, p3 = 0.f;
        int ity = 1, nn = 0;
        if (valid) {
            double4 p = pos[ia];
            p1 = __double2float_rn(p.x); p2 = __double2float_rn(p.y); p3 = __double2float_rn(p.z);
            ity = ityp[ia];
        }
        for (int k = 0; k < 27; k++) {
            int cc[3] = {ix0 + c_nix[k], iy0 + c_niy[k], iz0 + c_niz[k]};
            const int ncs[3] = {P.ncx, P.ncy, P.ncz};
            float sh[3] = {0.f, 0.f, 0.f};
            bool out = false;
#pragma unroll
            for (int d = 0; d < 3; d++) {
                if (P.pd[d] && k > 0) { // :1031-1059
                    if (cc[d] >= ncs[d]) { cc[d] = 0; sh[d] = P.fbs[d]; }
                    else if (cc[d] < 0) { cc[d] = ncs[d] - 1; sh[d] = -P.fbs[d]; }
                }
                if (cc[d] >= ncs[d] || cc[d] < 0) out = true;
            }
            if (out) continue;
            const int cid = ncxy * cc[2] + P.ncx * cc[1] + cc[0] + is0 * P.nc0;
            const int j0 = ia1th[cid] - 1, nj = nac[cid];
            for (int jb = 0; jb < nj; jb += 32) {
                if (jb + lane < nj) {
                    double4 q = pos[j0 + jb + lane];
                    float4 s;
                    s.x = __double2float_rn(__dadd_rn(q.x, (double)sh[0]));
                    s.y = __double2float_rn(__dadd_rn(q.y, (double)sh[1]));
                    s.z = __double2float_rn(__dadd_rn(q.z, (double)sh[2]));
                    s.w = __int_as_float(ityp[j0 + jb + lane]);
                    tile[wib][lane] = s;
                }
                __syncwarp();
                const int cnt = min(32, nj - jb);
                if (valid) {
                    for (int t = 0; t < cnt; t++) {
                        const float4 s = tile[wib][t];
                        const float e1 = __fsub_rn(p1, s.x), e2 = __fsub_rn(p2, s.y), e3 = __fsub_rn(p3, s.z);
                        float d1 = e1, d2 = e2, d3 = e3;
                        if (!P.identity) { // :1117-1119
                            d1 = __fadd_rn(__fadd_rn(__fmul_rn(P.bsh[0], e1), __fmul_rn(P.bsh[3], e2)), __fmul_rn(P.bsh[6], e3));
                            d2 = __fadd_rn(__fadd_rn(__fmul_rn(P.bsh[1], e1), __fmul_rn(P.bsh[4], e2)), __fmul_rn(P.bsh[7], e3));
                            d3 = __fadd_rn(__fadd_rn(__fmul_rn(P.bsh[2], e1), __fmul_rn(P.bsh[5], e2)), __fmul_rn(P.bsh[8], e3));
                        }
                        const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(d1, d1), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
                        const int jty = __float_as_int(s.w);
                        const int j = j0 + jb + t;
                        if (r2 <= P.rm2[(ity - 1) + P.ng * (jty - 1)] && !(k == 0 && j == ia)) { // :1123-1124
                            nn++;
                            if (nn <= P.mxkvois) indi[ia + (size_t)(nn - 1) * P.n] = j + 1;
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (valid) {
            kvois[ia] = min(nn, P.mxkvois); // silently truncated :1195
            if (nn > P.mxkvois) atomicAdd(&counters[CNT_OVERFLOW], 1);
            atomicMax(&counters[CNT_NNMAX], nn);
        }
    }
}

int mdb_nlist_kernel(mdb_ctx *c, const double4 *pos)
{
    NlistParams P;
    P.n = c->n; P.nc = c->nc; P.nc0 = c->nc0; P.ncx = c->ncell[0]; P.ncy = c->ncell[1]; P.ncz = c->ncell[2];
    P.ng = c->ng; P.mxkvois = c->mxkvois;
    for (int d = 0; d < 3; d++) { P.pd[d] = c->box.pd[d]; P.fbs[d] = (float)c->box.size[d]; }
    for (int i = 0; i < 9; i++) P.bsh[i] = (float)c->boxshape[i];
    P.identity = c->shape_identity ? 1 : 0;
    for (int i = 0; i < MDB_MXGROUP * MDB_MXGROUP; i++) P.rm2[i] = (i < c->ng * c->ng) ? c->rm2f[i] : 0.f;
    ProfScope ps(c, MDB_K_NLIST);
    k_nlist_cellwarp<<<cdiv(c->nc, NL_WARPS), NL_WARPS * 32, 0, c->stream>>>(P, pos ? pos : c->pos, c->ityp, c->nac, c->naac,
                                                                             c->ia1th, c->kvois, c->indi, c->counters);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}


// ---- Reorder_NeighBoreList_Nearest_Dev (CommonGPU/MD_NeighborsList_GPU.F90:2016-2066, kernel :1805-1946): every atom keeps
// its NEAREST closest listed neighbours, ordered by increasing distance (ties keep list order: strict '<'), in place.
// One warp per atom: lanes take the entries w = lane, lane+32, ..., the (distance, list position) keys are ranked by counting
// -- rank = #keys smaller -- which is the stable order the reference's insertion sort produces; no per-thread 512-entry
// scratch arrays (the reference keeps R2SWAP/NEARSWAP(512) per thread in local memory).
#define NEAREST_MAXLIST 512 // mp_MXNEAREST :59 (also the largest list the shared staging holds)
__global__ void __launch_bounds__(128)
k_nearest_reorder(int n, int mxkvois, int nearest, const double4 *__restrict__ pos, BoxParams box, int *__restrict__ kvois,
                  int *__restrict__ indi)
{
    __shared__ double s_r2[4][NEAREST_MAXLIST];
    __shared__ int s_j[4][NEAREST_MAXLIST];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + wib;
    if (i >= n) return;
    const int kv = min(kvois[i], NEAREST_MAXLIST);
    const double4 pi = pos[i];
    for (int w = lane; w < kv; w += 32) {
        const int j = indi[i + (size_t)w * n];
        const double4 pj = pos[j - 1];
        double sx = pi.x - pj.x, sy = pi.y - pj.y, sz = pi.z - pj.z;
        if (box.pd[0] > 0 && fabs(sx) > box.half[0]) sx = sx - copysign(box.size[0], sx);
        if (box.pd[1] > 0 && fabs(sy) > box.half[1]) sy = sy - copysign(box.size[1], sy);
        if (box.pd[2] > 0 && fabs(sz) > box.half[2]) sz = sz - copysign(box.size[2], sz);
        s_r2[wib][w] = __dadd_rn(__dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy)), __dmul_rn(sz, sz));
        s_j[wib][w] = j;
    }
    __syncwarp();
    const int nn = min(kv, nearest);
    for (int w = lane; w < kv; w += 32) {
        const double r2 = s_r2[wib][w];
        int rank = 0;
        for (int k = 0; k < kv; k++) {
            const double q = s_r2[wib][k];
            rank += (q < r2 || (q == r2 && k < w)) ? 1 : 0;
        }
        if (rank < nn) indi[i + (size_t)rank * n] = s_j[wib][w];
    }
    if (lane == 0) kvois[i] = nn;
}

int mdb_nlist_nearest(mdb_ctx *c, int nearest)
{
    ProfScope ps(c, MDB_K_NLIST);
    k_nearest_reorder<<<cdiv(c->n, 4), 128, 0, c->stream>>>(c->n, c->mxkvois, nearest, c->pos, c->box, c->kvois, c->indi);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

// mdb_cells.cu -- cell binning and cell-ordered re-sort, entirely on the device.
//
// Replaces the cell-id kernel + the SERIAL HOST linked-cell sort of the reference
// (CommonGPU/MD_NeighborsList_GPU.F90:759-819 and :1421-1574,1624-1695: ~10 full-array
// host<->device copies and an O(N) serial loop per rebuild).  The resulting order is
// bit-identical to the reference's: cells ascending (x fastest, then y, z, box), atoms
// inside a cell by DESCENDING original id (head/link insertion order, :1493-1495,1558-1565),
// out-of-box atoms parked at the end, smallest original id last (:1627-1637).
//
// Pipeline (all on ctx->stream, no host round trip):
//   k_cell_assign   cell id per atom + per-cell counts by warp-aggregated atomics
//   k_scan_local / k_scan_apply   exclusive prefix -> IA1th, max count (two kernels, one cell per thread)
//   k_cell_scatter  provisional placement inside the cell segment (atomic order)
//   k_cell_rank     deterministic in-cell order: rank by descending original id
//   k_oob_place     out-of-box atoms to the tail
//   k_permute       gather every per-atom array into the new order (double-buffered)
#include "mdb_internal.cuh"

// (x-LB)/BS*NC - eps, un-fused and in source order: cell assignment must be bit-exact
__device__ __forceinline__ int cell_coord(double x, double lo, double bs, int nc, double eps)
{
    double t = __ddiv_rn(__dsub_rn(x, lo), bs);
    t = __dsub_rn(__dmul_rn(t, (double)nc), eps);
    return (int)t; // truncation toward zero, like Fortran int()
}

__global__ void k_cell_assign(int n, int napb, const double4 *__restrict__ pos, const int *__restrict__ gid,
                              int *__restrict__ statu, BoxParams box, int ncx, int ncy, int ncz, int nc0,
                              int *__restrict__ ic, int *__restrict__ nac, int *__restrict__ naac,
                              int *__restrict__ slot, int *__restrict__ oob, int *__restrict__ counters)
{
    const double eps = (double)0.0001f; // real(KINDDF),parameter::eps=0.0001 :789
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    int cell = 0, active = 0, orig = 0;
    if (s < n) {
        int st = statu[s];
        orig = gid[s];
        int ib = (orig - 1) / napb; // :800
        if ((st & ST_OUTOFBOX) != ST_OUTOFBOX) {
            double4 p = pos[s];
            int ix = cell_coord(p.x, box.lo[0], box.size[0], ncx, eps);
            int iy = cell_coord(p.y, box.lo[1], box.size[1], ncy, eps);
            int iz = cell_coord(p.z, box.lo[2], box.size[2], ncz, eps);
            if (ix < 0 || ix >= ncx || iy < 0 || iy >= ncy || iz < 0 || iz >= ncz) cell = -2;
            else cell = 1 + (ix + ncx * (iy + ncy * iz)) + ib * nc0;
        } else {
            cell = -1;
        }
        ic[s] = cell;
        active = (st & ST_ACTIVE) == ST_ACTIVE;
        if (cell < 0) {
            // the reference asks on stdin here and, on 'C', marks the atom out of box (:1505-1526)
            if (cell < -1) statu[s] = ST_OUTOFBOX;
            int k = atomicAdd(&counters[CNT_OOB], 1);
            oob[k] = orig;
        }
    }
    // warp-aggregated counting: lanes of the same cell elect a leader that issues one atomic
    int key = (cell > 0) ? cell : -(lane + 1);
    unsigned grp = __match_any_sync(0xffffffffu, key);
    unsigned act = __ballot_sync(0xffffffffu, active && cell > 0);
    int leader = __ffs(grp) - 1;
    int base = 0;
    if (cell > 0 && lane == leader) {
        base = atomicAdd(&nac[cell - 1], __popc(grp));
        int na = __popc(grp & act);
        if (na) atomicAdd(&naac[cell - 1], na);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (cell > 0) slot[s] = base + __popc(grp & ((1u << lane) - 1u));
}

// Exclusive prefix over the cells [c0, c1) in two kernels (the reference scans on the host, :1542-1551; round 1 used one
// 1024-thread block, 40 us at 42 875 cells): k_scan_local -- one cell per thread, block-wide scan, block total and maximum;
// k_scan_apply -- every block adds the totals of the blocks before it (<= ~650 numbers even at 16 M atoms, summed by the block
// itself: no third kernel, no look-back chain) and writes IA1th = first + prefix + 1 (1-based as hm_IA1th).
#define SCAN_B 1024
__device__ __forceinline__ int block_excl_scan(int v, int &total, int &vmax)
{
    __shared__ int wsum[32], wmax[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v, mx = v;
    for (int off = 1; off < 32; off <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += u;
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 31) wsum[w] = inc;
    if (lane == 0) wmax[w] = mx;
    __syncthreads();
    if (w == 0) {
        const int nw = (int)(blockDim.x >> 5);
        int x = lane < nw ? wsum[lane] : 0, m = lane < nw ? wmax[lane] : 0, xi = x;
        for (int off = 1; off < 32; off <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, xi, off);
            if (lane >= off) xi += u;
            m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
        }
        wsum[lane] = xi - x;                      // exclusive prefix of the warp sums
        if (lane == 31) wmax[0] = m;
        if (lane == 31) wmax[1] = xi;             // block total
    }
    __syncthreads();
    total = wmax[1]; vmax = wmax[0];
    return wsum[w] + inc - v;
}
__global__ void __launch_bounds__(SCAN_B) k_scan_local(int c0, int c1, const int *__restrict__ nac, int *__restrict__ ia1th,
                                                       int *__restrict__ btot, int *__restrict__ bmax)
{
    const int i = c0 + blockIdx.x * SCAN_B + threadIdx.x;
    const int v = i < c1 ? nac[i] : 0;
    int total, vmax;
    const int ex = block_excl_scan(v, total, vmax);
    if (i < c1) ia1th[i] = ex;
    if (threadIdx.x == 0) { btot[blockIdx.x] = total; bmax[blockIdx.x] = vmax; }
}
// cnt_incell / cnt_mxnac (optional): device counters that take the number of atoms in [c0, c1) and the largest cell.
// out4 (optional, slab decomposition): {atoms in [c0, c1), atoms of the first `cl` cells, atoms of the last `cl` cells, max count}
__global__ void __launch_bounds__(SCAN_B) k_scan_apply(int c0, int c1, int first, int cl, const int *__restrict__ btot,
                                                       const int *__restrict__ bmax, int *__restrict__ ia1th, int *__restrict__ cnt_incell,
                                                       int *__restrict__ cnt_mxnac, int *__restrict__ out4)
{
    __shared__ int red[3][32];
    const int nb = gridDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    int before = 0, all = 0, mx = 0;
    for (int b = t; b < nb; b += SCAN_B) {
        const int v = btot[b];
        all += v;
        if (b < (int)blockIdx.x) before += v;
        mx = max(mx, bmax[b]);
    }
    for (int off = 16; off > 0; off >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, off);
        all += __shfl_xor_sync(0xffffffffu, all, off);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 0) { red[0][w] = before; red[1][w] = all; red[2][w] = mx; }
    __syncthreads();
    before = 0; all = 0; mx = 0;
    for (int k = 0; k < SCAN_B / 32; k++) { before += red[0][k]; all += red[1][k]; mx = max(mx, red[2][k]); }
    const int i = c0 + blockIdx.x * SCAN_B + t;
    if (i < c1) {
        const int pre = before + ia1th[i];               // atoms in the cells [c0, i)
        ia1th[i] = first + pre + 1;
        if (out4) {
            const int nc = c1 - c0;
            if (nc > cl && i == c0 + cl) out4[1] = pre;
            if (nc > cl && i == c1 - cl) out4[2] = all - pre;
        }
    }
    if (blockIdx.x == 0 && t == 0) {
        if (cnt_incell) *cnt_incell = all;
        if (cnt_mxnac) *cnt_mxnac = mx;
        if (out4) {
            out4[0] = all; out4[3] = mx;
            if (c1 - c0 <= cl) { out4[1] = all; out4[2] = all; }
        }
    }
}
// host: the two launches; scratch = 2 * cdiv(c1 - c0, SCAN_B) ints
static int scan_cells(mdb_ctx *c, int c0, int c1, int first, int cl, int *cnt_incell, int *cnt_mxnac, int *out4)
{
    const int nb = cdiv(c1 - c0, SCAN_B);
    if (nb <= 0) return MDB_OK;
    if (c->scan_n < 2 * nb) {
        if (c->scan_tmp) cudaFree(c->scan_tmp);
        c->scan_tmp = nullptr; c->scan_n = 0;
        CUDA_TRY(c, cudaMalloc(&c->scan_tmp, sizeof(int) * 2 * (size_t)(nb + 64)));
        c->scan_n = 2 * (nb + 64);
    }
    int *btot = c->scan_tmp, *bmax = c->scan_tmp + c->scan_n / 2;
    k_scan_local<<<nb, SCAN_B, 0, c->stream>>>(c0, c1, c->nac, c->ia1th, btot, bmax);
    k_scan_apply<<<nb, SCAN_B, 0, c->stream>>>(c0, c1, first, cl, btot, bmax, c->ia1th, cnt_incell, cnt_mxnac, out4);
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

__global__ void k_cell_scatter(int n, const int *__restrict__ ic, const int *__restrict__ slot,
                               const int *__restrict__ ia1th, const int *__restrict__ gid, int *__restrict__ tmp_orig)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int cell = ic[s];
    if (cell > 0) tmp_orig[ia1th[cell - 1] - 1 + slot[s]] = gid[s];
}

__global__ void k_cell_rank(int n, const int *__restrict__ ic, const int *__restrict__ ia1th, const int *__restrict__ nac,
                            const int *__restrict__ gid, const int *__restrict__ tmp_orig, int *__restrict__ gid_new,
                            int *__restrict__ srcof)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int cell = ic[s];
    if (cell <= 0) return;
    int base = ia1th[cell - 1] - 1, cnt = nac[cell - 1], me = gid[s], rank = 0;
    for (int k = 0; k < cnt; k++) rank += (tmp_orig[base + k] > me); // descending original id
    gid_new[base + rank] = me;
    srcof[base + rank] = s;
}

__global__ void k_oob_place(int n, int *__restrict__ counters, const int *__restrict__ oob,
                            const int *__restrict__ gidinv, int *__restrict__ gid_new, int *__restrict__ srcof)
{
    int m = counters[CNT_OOB];
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[CNT_OOB_TOTAL] += m;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < m; q += gridDim.x * blockDim.x) {
        int o = oob[q], rank = 0;
        for (int k = 0; k < m; k++) rank += (oob[k] < o);
        int dest = n - 1 - rank; // IP = NPART, NPART-1, ... over ascending original id :1628-1636
        gid_new[dest] = o;
        srcof[dest] = gidinv[o - 1] - 1;
    }
}

__global__ void k_permute(int n, const int *__restrict__ srcof, const int *__restrict__ gid_new,
                          const double4 *__restrict__ pos, double4 *__restrict__ pos_o,
                          const double *__restrict__ xp1, double *__restrict__ xp1_o,
                          const double *__restrict__ fp, double *__restrict__ fp_o,
                          const double *__restrict__ dis, double *__restrict__ dis_o,
                          const int *__restrict__ ityp, int *__restrict__ ityp_o,
                          const int *__restrict__ statu, int *__restrict__ statu_o,
                          const int *__restrict__ ic, int *__restrict__ ic_o, int *__restrict__ gidinv)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    int s = srcof[d];
    pos_o[d] = pos[s];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        xp1_o[d + (size_t)k * n] = xp1[s + (size_t)k * n];
        fp_o[d + (size_t)k * n] = fp[s + (size_t)k * n];
        dis_o[d + (size_t)k * n] = dis[s + (size_t)k * n];
    }
    ityp_o[d] = ityp[s];
    statu_o[d] = statu[s];
    ic_o[d] = ic[s];
    gidinv[gid_new[d] - 1] = d + 1;
}

template <class T> static inline void swp(T *&a, T *&b) { T *t = a; a = b; b = t; }

int mdb_cells_build(mdb_ctx *c)
{
    int n = c->n, nb = cdiv(n, 256);
    cudaStream_t st = c->stream;
    ProfScope ps(c, MDB_K_CELLSORT, 7);
    CUDA_TRY(c, cudaMemsetAsync(c->counters, 0, sizeof(int) * CNT_PERBUILD_N, st)); // CNT_OOB_TOTAL accumulates
    CUDA_TRY(c, cudaMemsetAsync(c->nac, 0, sizeof(int) * (size_t)c->nc, st));  // hm_NAC = 0 :1431
    CUDA_TRY(c, cudaMemsetAsync(c->naac, 0, sizeof(int) * (size_t)c->nc, st));
    if (c->dsr) CUDA_TRY(c, cudaMemsetAsync(c->dsr, 0, 3 * (size_t)n * sizeof(float), st)); // displacement since THIS rebuild
    k_cell_assign<<<nb, 256, 0, st>>>(n, c->napb, c->pos, c->gid, c->statu, c->box, c->ncell[0], c->ncell[1],
                                      c->ncell[2], c->nc0, c->ic, c->nac, c->naac, c->slot, c->oob, c->counters);
    {
        int rc = scan_cells(c, 0, c->nc, 0, c->nc, c->counters + CNT_INCELL, c->counters + CNT_MXNAC, nullptr);
        if (rc < 0) return rc;
    }
    k_cell_scatter<<<nb, 256, 0, st>>>(n, c->ic, c->slot, c->ia1th, c->gid, c->tmp_orig);
    k_cell_rank<<<nb, 256, 0, st>>>(n, c->ic, c->ia1th, c->nac, c->gid, c->tmp_orig, c->gid_alt, c->srcof);
    k_oob_place<<<64, 256, 0, st>>>(n, c->counters, c->oob, c->gidinv, c->gid_alt, c->srcof);
    k_permute<<<nb, 256, 0, st>>>(n, c->srcof, c->gid_alt, c->pos, c->pos_alt, c->xp1, c->xp1_alt, c->fp, c->fp_alt,
                                  c->dis, c->dis_alt, c->ityp, c->ityp_alt, c->statu, c->statu_alt, c->ic, c->ic_alt,
                                  c->gidinv);
    CUDA_TRY(c, cudaGetLastError());
    swp(c->pos, c->pos_alt); swp(c->xp1, c->xp1_alt); swp(c->fp, c->fp_alt); swp(c->dis, c->dis_alt);
    swp(c->ityp, c->ityp_alt); swp(c->statu, c->statu_alt); swp(c->ic, c->ic_alt); swp(c->gid, c->gid_alt);
    return MDB_OK;
}

// =====================================================================================
// Slab-decomposed rebuild (mdb_dd.cu): a rank re-sorts only the atoms that end up in ITS z-layers of cells.
// Candidates are the atoms of its old owned range and of its two old ghost layers (an atom moves far less than a cell
// between two rebuilds); the result is written at the GLOBAL slots of the common cell-sorted order -- cells ascending,
// atoms of a cell by descending original id -- so owned slices stay bit-identical to the single-GPU order.
// =====================================================================================
struct DDRanges { int r0[3], r1[3]; }; // old-slot ranges [r0, r1): ghost below, owned, ghost above

__device__ __forceinline__ int dd_slot_of(const DDRanges &R, int t)
{
    int s = -1;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int len = R.r1[k] - R.r0[k];
        if (s < 0 && t < len) s = R.r0[k] + t;
        t -= (s < 0) ? len : 0;
    }
    return s;
}

__global__ void k_dd_cell_assign(DDRanges R, int total, const double4 *__restrict__ pos, int *__restrict__ statu, BoxParams box,
                                 int ncx, int ncy, int ncz, int zl0, int zl1, int *__restrict__ ic, int *__restrict__ nac,
                                 int *__restrict__ naac, int *__restrict__ slot, int *__restrict__ counters)
{
    const double eps = (double)0.0001f;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int cell = 0, active = 0, s = -1;
    if (t < total) {
        s = dd_slot_of(R, t);
        const int st = statu[s];
        if ((st & ST_OUTOFBOX) != ST_OUTOFBOX) {
            const double4 p = pos[s];
            const int ix = cell_coord(p.x, box.lo[0], box.size[0], ncx, eps);
            const int iy = cell_coord(p.y, box.lo[1], box.size[1], ncy, eps);
            const int iz = cell_coord(p.z, box.lo[2], box.size[2], ncz, eps);
            if (ix < 0 || ix >= ncx || iy < 0 || iy >= ncy || iz < 0 || iz >= ncz) cell = -2;
            else if (iz >= zl0 && iz < zl1) cell = 1 + (ix + ncx * (iy + ncy * iz));
        } else cell = -1;
        if (cell < 0) { atomicAdd(&counters[CNT_OOB], 1); cell = 0; } // a decomposed box must be periodic / closed: reported as an error
        ic[s] = cell;
        active = (st & ST_ACTIVE) == ST_ACTIVE;
    }
    const int key = (cell > 0) ? cell : -(lane + 1);
    const unsigned grp = __match_any_sync(0xffffffffu, key);
    const unsigned act = __ballot_sync(0xffffffffu, active && cell > 0);
    const int leader = __ffs(grp) - 1;
    int base = 0;
    if (cell > 0 && lane == leader) {
        base = atomicAdd(&nac[cell - 1], __popc(grp));
        const int na = __popc(grp & act);
        if (na) atomicAdd(&naac[cell - 1], na);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (cell > 0) slot[s] = base + __popc(grp & ((1u << lane) - 1u));
}

__global__ void k_dd_add_base(int c0, int c1, int base, int *__restrict__ ia1th)
{
    const int i = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c1) ia1th[i] += base;
}

__global__ void k_dd_scatter(DDRanges R, int total, const int *__restrict__ ic, const int *__restrict__ slot,
                             const int *__restrict__ ia1th, const int *__restrict__ gid, int *__restrict__ tmp_orig)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int s = dd_slot_of(R, t), cell = ic[s];
    if (cell > 0) tmp_orig[ia1th[cell - 1] - 1 + slot[s]] = gid[s];
}

__global__ void k_dd_rank(DDRanges R, int total, const int *__restrict__ ic, const int *__restrict__ ia1th, const int *__restrict__ nac,
                          const int *__restrict__ gid, const int *__restrict__ tmp_orig, int *__restrict__ gid_new, int *__restrict__ srcof)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int s = dd_slot_of(R, t), cell = ic[s];
    if (cell <= 0) return;
    const int base = ia1th[cell - 1] - 1, cnt = nac[cell - 1], me = gid[s];
    int rank = 0;
    for (int k = 0; k < cnt; k++) rank += (tmp_orig[base + k] > me);
    gid_new[base + rank] = me;
    srcof[base + rank] = s;
}

__global__ void k_dd_permute(int n, int a0, int a1, const int *__restrict__ srcof, const int *__restrict__ gid_new,
                             const double4 *__restrict__ pos, double4 *__restrict__ pos_o,
                             const double *__restrict__ xp1, double *__restrict__ xp1_o,
                             const double *__restrict__ fp, double *__restrict__ fp_o,
                             const double *__restrict__ dis, double *__restrict__ dis_o,
                             const int *__restrict__ ityp, int *__restrict__ ityp_o,
                             const int *__restrict__ statu, int *__restrict__ statu_o,
                             const int *__restrict__ ic, int *__restrict__ ic_o, int *__restrict__ gidinv)
{
    const int d = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= a1) return;
    const int s = srcof[d];
    pos_o[d] = pos[s];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        xp1_o[d + (size_t)k * n] = xp1[s + (size_t)k * n];
        fp_o[d + (size_t)k * n] = fp[s + (size_t)k * n];
        dis_o[d + (size_t)k * n] = dis[s + (size_t)k * n];
    }
    ityp_o[d] = ityp[s];
    statu_o[d] = statu[s];
    ic_o[d] = ic[s];
    gidinv[gid_new[d] - 1] = d + 1;
}

// phase 1: bin the candidates into this rank's cells and scan; d_out4 = {owned atoms, bottom layer, top layer, max per cell}
int mdb_cells_dd_count(mdb_ctx *c, const int cand[6], int zl0, int zl1, int *d_out4)
{
    const int cl = c->ncell[0] * c->ncell[1], c0 = zl0 * cl, c1 = zl1 * cl;
    cudaStream_t st = c->stream;
    DDRanges R;
    int total = 0;
    for (int k = 0; k < 3; k++) { R.r0[k] = cand[2 * k]; R.r1[k] = cand[2 * k + 1]; total += R.r1[k] - R.r0[k]; }
    ProfScope ps(c, MDB_K_CELLSORT, 3);
    CUDA_TRY(c, cudaMemsetAsync(c->counters, 0, sizeof(int) * CNT_PERBUILD_N, st));
    CUDA_TRY(c, cudaMemsetAsync(c->nac + c0, 0, sizeof(int) * (size_t)(c1 - c0), st));
    CUDA_TRY(c, cudaMemsetAsync(c->naac + c0, 0, sizeof(int) * (size_t)(c1 - c0), st));
    k_dd_cell_assign<<<cdiv(total, 256), 256, 0, st>>>(R, total, c->pos, c->statu, c->box, c->ncell[0], c->ncell[1], c->ncell[2], zl0, zl1,
                                                       c->ic, c->nac, c->naac, c->slot, c->counters);
    return scan_cells(c, c0, c1, 0, cl, nullptr, nullptr, d_out4);
}

// phase 2: with the global slot `base` of this rank's first atom known, place and permute the owned range [base, base + nown)
int mdb_cells_dd_place(mdb_ctx *c, const int cand[6], int zl0, int zl1, int base, int nown)
{
    const int cl = c->ncell[0] * c->ncell[1], c0 = zl0 * cl, c1 = zl1 * cl, n = c->n;
    cudaStream_t st = c->stream;
    DDRanges R;
    int total = 0;
    for (int k = 0; k < 3; k++) { R.r0[k] = cand[2 * k]; R.r1[k] = cand[2 * k + 1]; total += R.r1[k] - R.r0[k]; }
    ProfScope ps(c, MDB_K_CELLSORT, 4);
    k_dd_add_base<<<cdiv(c1 - c0, 256), 256, 0, st>>>(c0, c1, base, c->ia1th);
    k_dd_scatter<<<cdiv(total, 256), 256, 0, st>>>(R, total, c->ic, c->slot, c->ia1th, c->gid, c->tmp_orig);
    k_dd_rank<<<cdiv(total, 256), 256, 0, st>>>(R, total, c->ic, c->ia1th, c->nac, c->gid, c->tmp_orig, c->gid_alt, c->srcof);
    if (nown > 0)
        k_dd_permute<<<cdiv(nown, 256), 256, 0, st>>>(n, base, base + nown, c->srcof, c->gid_alt, c->pos, c->pos_alt, c->xp1, c->xp1_alt,
                                                      c->fp, c->fp_alt, c->dis, c->dis_alt, c->ityp, c->ityp_alt, c->statu, c->statu_alt,
                                                      c->ic, c->ic_alt, c->gidinv);
    CUDA_TRY(c, cudaGetLastError());
    if (c->dsr && nown > 0) {
        for (int k = 0; k < 3; k++) CUDA_TRY(c, cudaMemsetAsync(c->dsr + base + (size_t)k * n, 0, sizeof(float) * (size_t)nown, st));
    }
    swp(c->pos, c->pos_alt); swp(c->xp1, c->xp1_alt); swp(c->fp, c->fp_alt); swp(c->dis, c->dis_alt);
    swp(c->ityp, c->ityp_alt); swp(c->statu, c->statu_alt); swp(c->ic, c->ic_alt); swp(c->gid, c->gid_alt);
    return MDB_OK;
}

// IA1th of one ghost layer of cells [c0, c0 + cl) from the counts received from the neighbour; its first atom sits at `first`
int mdb_cells_dd_ghost_layer(mdb_ctx *c, int c0, int first)
{
    const int cl = c->ncell[0] * c->ncell[1];
    ProfScope ps(c, MDB_K_CELLSORT, 2);
    return scan_cells(c, c0, c0 + cl, first, cl, nullptr, nullptr, nullptr);
}

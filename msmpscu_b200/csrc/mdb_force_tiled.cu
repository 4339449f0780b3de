// mdb_force_tiled.cu -- the TILED fast path: neighbour-list build, density pass, force pass (+ virial epilogue) and
// per-atom energy pass over shared-memory staged halo tiles (geometry in mdb_tiled.cuh).
//
// Why: with one thread per atom over a global-id list (the reference design, kept as the generic path) every (atom,
// neighbour) visit is a 32-byte random gather that costs a full L1 wavefront per lane and an un-fused sqrt/sqrt/div/div
// chain on the fp64 pipe; on B200 that runs at ~5 % of the HBM roofline (profiles/r01_generic_path_summary.md).
//
// Structure:
//   tile        a run of cells along x in one cell row: owned atoms contiguous, halo = (Wt+2) x 3 x 3 cells = at most 27
//               contiguous RUNS of the {x,y,z,den} record array.  k_tile_desc writes one descriptor per tile per rebuild.
//   k_tile_nlist  one CTA per tile, two warps per owned cell (lanes = atoms; several warp pairs share a cell of more than 32
//               atoms when the tile is narrow, each taking 32-atom blocks in turn); the halo is staged once as pairs of fp32
//               candidates, membership is the reference's fp32 expression (Cal_NeighboreList_Kernel2C,
//               CommonGPU/MD_NeighborsList_GPU.F90:1097-1131), accepted 14-bit halo SLOTS + a 2-bit build-distance class go to
//               a per-atom shared-memory column, which is partitioned by class (optionally dealt out by slot residue for
//               conflict-free record reads, MDB_OPT_TILED_BANKORDER) and written as a lane-interleaved 16-bit list.
//   k_tile_pass   one persistent CTA per SM, warp-specialised: a PRODUCER warp brings each tile's runs, first index group,
//               STATU and scan counts into a 2-stage shared-memory pipeline with TMA bulk copies (cp.async.bulk + mbarrier);
//               (multi-type boxes: its lanes also copy the halo atoms' types, aligned int4 windows of ITYP, 8 in flight);
//               CONSUMER warps grab chunks of 32/G owned atoms (G = 4, or 8 where a tile owns < 96 atoms) from a shared counter and evaluate every listed entry
//               branch-free: fp64 separation from the staged records, exact r2 <= r_eff^2, MUFU.RSQ64H + Halley for 1/r and
//               sqrt(r), table rows from a shared-memory window, masked accumulation, G-lane shuffle reduce.  No CTA-wide
//               barrier anywhere.  PASS 1 density -> dF/drho, PASS 2 forces (VIR: + per-warp partial virial tensors),
//               PASS 3 per-atom energies.
// Exact work reductions (results identical to evaluating every listed pair):
//   * rows beyond the last non-zero table row interpolate to exactly 0, so the in-range test uses min(RU, table support);
//   * the slot list is stored in three classes by BUILD-time distance (<= r_eff(pass1)+m0, <= r_eff(pass2)+m, rest).  A pass
//     scans only its classes while every atom has moved less than half the margin since the rebuild (tracked by the
//     predictor); otherwise it scans the whole list.
// The list's members are the reference's; its ORDER is free on this path (the passes only sum over it): the reference-ordered
// KVOIS/INDI pair is produced on demand by the generic kernel from the positions saved at the rebuild (mdb_indi_ensure).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>
#include "mdb_tiled.cuh"

__constant__ int t_nix[27] = {0, -1, -1, -1, 0, 0, -1, 1, -1, 0, 1, -1, 0, 1, 1, 1, 1, 0, 0, -1, 1, -1, 0, 1, -1, 0, 1};
__constant__ int t_niy[27] = {0, 0, -1, 1, 1, 0, 0, 0, -1, -1, -1, 1, 1, 1, 0, 1, -1, -1, 0, 0, 0, -1, -1, -1, 1, 1, 1};
__constant__ int t_niz[27] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1};

// per-tile halo description: computed once per rebuild by the list kernel, read by the force passes
struct TileDesc {
    int htot, own_start, own_slot0, own_count, edge_any, nhc, pad0, pad1;
    int slot[TILE_MAX_HC + 2]; // first slot of each halo cell (exclusive prefix), [nhc] = total
    int cnt[TILE_MAX_HC];
    int gst[TILE_MAX_HC];      // first global (cell-order, 0-based) atom of the cell
    int cid[TILE_MAX_HC];      // wrapped global cell id, -1 if absent
    signed char sh[TILE_MAX_HC][4]; // image shift per dim, [3] = edge flag
    // RUNS: maximal stretches of halo cells of one row that are contiguous in the cell-sorted arrays (a row
    // breaks only where it wraps around a periodic face) -> one TMA bulk copy each, at most 3 per row
    int nrun, rpad[3];
    int rslot[TILE_MAX_RUN + 1];   // first slot of each run, [nrun] = htot
    int rgst[TILE_MAX_RUN];        // first global atom of each run
};

// fills the halo table in shared memory; must be called by all threads of the CTA
__device__ __forceinline__ void build_halo_table(const TileParams &P, const TileGeom &g, const int *__restrict__ nac,
                                                 const int *__restrict__ ia1th, TileDesc &H)
{
    for (int hc = threadIdx.x; hc < g.nhc; hc += blockDim.x) {
        int sh[3], u[3];
        const int cid = halo_cell(P, g, hc, sh, u);
        H.cid[hc] = cid;
        H.cnt[hc] = cid >= 0 ? nac[cid] : 0;
        H.gst[hc] = cid >= 0 ? ia1th[cid] - 1 : 0;
        H.sh[hc][0] = (signed char)sh[0]; H.sh[hc][1] = (signed char)sh[1]; H.sh[hc][2] = (signed char)sh[2];
        // edge flag: a wrapped cell, or a cell on a periodic face of the box (its atoms may have been
        // wrapped by the predictor since the rebuild)
        const int nc[3] = {P.ncx, P.ncy, P.ncz};
        bool edge = false;
        for (int d = 0; d < 3; d++) edge = edge || (P.pd[d] && (u[d] <= 0 || u[d] >= nc[d] - 1));
        H.sh[hc][3] = (signed char)((cid >= 0 && edge) ? 1 : 0);
    }
    __syncthreads();
    if (threadIdx.x < 32) { // one warp scans the <= 126 counts
        int run = 0;
        bool edge = false;
        for (int b = 0; b < g.nhc; b += 32) {
            const int hc = b + threadIdx.x;
            const int v = hc < g.nhc ? H.cnt[hc] : 0;
            edge = edge || (hc < g.nhc && H.sh[hc][3] != 0);
            int inc = v;
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, off);
                if ((int)threadIdx.x >= off) inc += t;
            }
            if (hc < g.nhc) H.slot[hc] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        edge = __any_sync(0xffffffffu, edge);
        if (threadIdx.x == 0) {
            H.slot[g.nhc] = run;
            H.htot = run; H.edge_any = edge ? 1 : 0; H.nhc = g.nhc;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int hc_own0 = (1 * 3 + 1) * g.nhx + 1; // centre row (hy = hz = 1), cells hx = 1..wt
        H.own_start = H.gst[hc_own0];
        H.own_slot0 = H.slot[hc_own0];
        H.own_count = H.slot[hc_own0 + g.wt] - H.slot[hc_own0];
    }
    if (threadIdx.x < 32) { // runs (IA1th is a running prefix over ALL cells, so empty cells do not break a run)
        const int lane = threadIdx.x;
        int nrun = 0;
        for (int b = 0; b < g.nhc; b += 32) {
            const int hc = b + lane;
            bool start = false;
            if (hc < g.nhc && H.cid[hc] >= 0)
                start = !((hc % g.nhx) != 0 && H.cid[hc - 1] >= 0 && H.gst[hc - 1] + H.cnt[hc - 1] == H.gst[hc]);
            const unsigned m = __ballot_sync(0xffffffffu, start);
            if (start) {
                const int rid = nrun + __popc(m & ((1u << lane) - 1u));
                if (rid < TILE_MAX_RUN) { H.rslot[rid] = H.slot[hc]; H.rgst[rid] = H.gst[hc]; }
            }
            nrun += __popc(m);
        }
        if (lane == 0) { H.nrun = nrun; H.rslot[min(nrun, TILE_MAX_RUN)] = H.slot[g.nhc]; }
    }
    __syncthreads();
}

// =====================================================================================
// list build
// =====================================================================================
struct TileListArgs {
    const double4 *pos; const int *ityp; const int *naac;
    int *kvois; unsigned short *nbl; unsigned short *ncls; int *counters; const TileDesc *desc;
    int tile_lo;          // first tile of this rank (slab decomposition), 0 otherwise
    int lcap;             // rows of the per-atom shared-memory list
    float rm2[MDB_MXGROUP * MDB_MXGROUP];
    float rc2[2]; // class radii^2 (build-time, fp32): class 0 <= rc2[0] < class 1 <= rc2[1] < class 2
};

#define NL_THREADS 128

// tile descriptors: one small CTA per tile
__global__ void __launch_bounds__(NL_THREADS)
k_tile_desc(TileParams P, const int *__restrict__ nac, const int *__restrict__ ia1th, TileDesc *__restrict__ desc,
            int *__restrict__ counters, int tile_lo)
{
    __shared__ TileDesc H;
    const int tile = tile_lo + blockIdx.x;
    const TileGeom g = tile_geom(P, tile);
    build_halo_table(P, g, nac, ia1th, H);
    const int *src = reinterpret_cast<const int *>(&H);
    int *dst = reinterpret_cast<int *>(desc + tile);
    for (int i = threadIdx.x; i < (int)(sizeof(TileDesc) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    // a halo that does not fit the shared-memory budget of the passes: the host falls back to the generic path
    if (threadIdx.x == 0 && (H.htot > P.hcap || H.own_count > P.ocap || H.nrun > TILE_MAX_RUN)) atomicAdd(&counters[CNT_TILE_OVERFLOW], 1);
}

// Per-tile displacement bound of cascade runs: one warp per tile takes the maximum of the predictor's per-block maxima
// (blocks of 256 atoms of the owned range [a0, a1)) over the tile's halo runs.  A run that reaches into ghost layers
// (slab decomposition) or any atom outside the predictor's range takes the global maximum, which the ghost exchange has
// merged with the neighbours' by then.  One fast atom then sends only the tiles around it to the full list.
__global__ void __launch_bounds__(256)
k_tile_d2(const TileDesc *__restrict__ desc, int tile_lo, int ntl, const float *__restrict__ dmax_blk, int a0, int a1,
          const int *__restrict__ counters, float *__restrict__ tile_d2)
{
    const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= ntl) return;
    const TileDesc &D = desc[tile_lo + w];
    const int nrun = D.nrun;
    float m = 0.f;
    bool outside = nrun > TILE_MAX_RUN;
    if (lane < min(nrun, TILE_MAX_RUN)) {
        const int g0 = D.rgst[lane], g1 = g0 + D.rslot[lane + 1] - D.rslot[lane];
        if (g1 > g0) {
            outside = outside || g0 < a0 || g1 > a1;
            const int lo = max(g0, a0), hi = min(g1, a1);
            if (hi > lo)
                for (int b = (lo - a0) >> 8; b <= (hi - 1 - a0) >> 8; b++) m = fmaxf(m, dmax_blk[b]);
        }
    }
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (__any_sync(0xffffffffu, outside)) m = fmaxf(m, __int_as_float(counters[CNT_D2MAX]));
    if (lane == 0) tile_d2[tile_lo + w] = m;
}

// One CTA per tile, one warp per owned cell (lanes = atoms of the cell).  The tile's halo is staged once in shared
// memory as the reference's fp32 candidate SPOS = (float)(XP + (double)(float)shift) (:1100-1103) with the type in
// .w; every lane then walks the 27 neighbour cells as 9 contiguous slot ranges (broadcast reads), decides
// membership with the reference's fp32 expression and appends accepted halo SLOTS, tagged with their build-distance
// class, to its own column of a shared-memory list.  The column is then partitioned by class in place and written
// out in the lane-interleaved layout the passes stream (nbl_index); tails are padded with the dummy slot P.hcap.
// The ORDER of a list is free here (the passes only sum over it): the reference-ordered KVOIS/INDI pair is
// produced on demand by the generic kernel from the positions saved at the rebuild (mdb_indi_ensure).  An atom
// with more neighbours than the list capacity or than mxKVOIS (where the reference truncates in ITS order) sends
// the whole build to the generic path through CNT_TILE_OVERFLOW.
// ---- packed fp32 helpers: two candidates per instruction (FADD2 / FMUL2).  Only the subtraction and the squares
// are packed: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2, which would change the reference's rounding,
// so the two sums stay scalar add.rn.f32 (never contracted).
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi)
{
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long f2_sq(unsigned long long a)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(a));
    return d;
}
__device__ __forceinline__ float f2_lo(unsigned long long a) { return __uint_as_float((unsigned)a); }
__device__ __forceinline__ float f2_hi(unsigned long long a) { return __uint_as_float((unsigned)(a >> 32)); }
// Push of slot + 2-bit distance class onto a shared-memory column, with the membership test (r2 <= rm, and for the centre
// row: slot != own slot) and the class tests evaluated INSIDE the asm: the predicates never take a round trip through
// general registers (the C++ form costs a SEL and a second SETP per candidate).
// (step = +64 or -64 bytes: the two warps of a cell fill the same column from its two ends)
__device__ __forceinline__ void push_if_in(unsigned &addr, unsigned slot, float r2, float rm, float rc0, float rc1, int step)
{
    asm volatile("{\n.reg .pred q, a, b;\n.reg .b32 e;\nsetp.le.f32 q, %2, %3;\nsetp.gt.f32 a, %2, %4;\nsetp.gt.f32 b, %2, %5;\n"
                 "mov.b32 e, %1;\n@a add.u32 e, e, 0x4000;\n@b add.u32 e, e, 0x4000;\n"
                 "@q st.shared.u16 [%0], e;\n@q add.s32 %0, %0, %6;\n}\n"
                 : "+r"(addr) : "r"(slot), "f"(r2), "f"(rm), "f"(rc0), "f"(rc1), "r"(step) : "memory");
}
__device__ __forceinline__ void push_if_in_notself(unsigned &addr, unsigned slot, float r2, float rm, float rc0, float rc1, unsigned self,
                                                   int step)
{
    asm volatile("{\n.reg .pred q, a, b;\n.reg .b32 e;\nsetp.le.f32 q, %2, %3;\nsetp.ne.and.u32 q, %1, %6, q;\nsetp.gt.f32 a, %2, %4;\n"
                 "setp.gt.f32 b, %2, %5;\nmov.b32 e, %1;\n@a add.u32 e, e, 0x4000;\n@b add.u32 e, e, 0x4000;\n"
                 "@q st.shared.u16 [%0], e;\n@q add.s32 %0, %0, %7;\n}\n"
                 : "+r"(addr) : "r"(slot), "f"(r2), "f"(rm), "f"(rc0), "f"(rc1), "r"(self), "r"(step) : "memory");
}
// predicated push of a 16-bit entry onto a shared-memory column (32-bit shared address, row stride 64 bytes)
__device__ __forceinline__ void sts16_push(unsigned &addr, unsigned val, bool p)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.shared.u16 [%0], %1;\n@q add.u32 %0, %0, 64;\n}\n"
                 : "+r"(addr) : "h"((unsigned short)val), "r"((unsigned)p) : "memory");
}

__device__ __forceinline__ void pair_barrier(int cell) // the two warps of one owned cell
{
    __syncwarp();
    asm volatile("bar.sync %0, 64;" ::"r"(cell + 1) : "memory");
}

// ---- bank-aware order of a list class (see the write-out of k_tile_nlist).
// residue (slot mod 4) that lane lam0 + e % G of a half-warp wants at list entry e (row e / G)
template <int G>
__device__ __forceinline__ unsigned list_wanted(int lam0, int e) { return (unsigned)((((lam0 + e % G) >> 1) + e / G) & 3); }
template <int G, bool MT>
__global__ void __launch_bounds__(64 * TILE_MAX_W, 2)
k_tile_nlist(TileParams P, TileListArgs A)
{
    extern __shared__ __align__(16) unsigned char nl_smem[];
    __shared__ TileDesc H;
    // per lane: entries (10 bits) and overflow count (6 bits, saturating) of the second warp.  16-bit on purpose: two CTAs of
    // this kernel share an SM only while 2 x (dynamic + static + 1 KB) stays under 228 KB, and the margin is ~1 KB.
    __shared__ unsigned short s_half1[TILE_MAX_W][32];
    // halo as PAIRS of candidates: hxy[p] = {x(2p), x(2p+1), y(2p), y(2p+1)}, hzt[p] = {z, z, type, type}
    const int npair = (P.hcap + 2) / 2;
    float4 *hxy = reinterpret_cast<float4 *>(nl_smem);
    float4 *hzt = hxy + npair;
    float *fxy = reinterpret_cast<float *>(hxy), *fzt = reinterpret_cast<float *>(hzt);
    unsigned short *lists = reinterpret_cast<unsigned short *>(hzt + npair);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int tile = A.tile_lo + blockIdx.x;
    {
        const int *src = reinterpret_cast<const int *>(A.desc + tile);
        int *dst = reinterpret_cast<int *>(&H);
        for (int i = threadIdx.x; i < (int)(sizeof(TileDesc) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    if (H.htot > P.hcap || H.own_count > P.ocap || H.nrun > TILE_MAX_RUN) return; // counted by k_tile_desc
    const int nhc = H.nhc, nhx = nhc / 9;
    // ---- stage the halo
    for (int hc = warp; hc < nhc; hc += nwarps) {
        const int cnt = H.cnt[hc];
        if (H.cid[hc] < 0 || cnt <= 0) continue;
        const int sl = H.slot[hc], gst = H.gst[hc];
        const float s0 = H.sh[hc][0] * P.fbs[0], s1 = H.sh[hc][1] * P.fbs[1], s2 = H.sh[hc][2] * P.fbs[2];
        for (int a = lane; a < cnt; a += 32) {
            const double4 q = A.pos[gst + a];
            const int s = sl + a, o = (s >> 1) * 4 + (s & 1);
            fxy[o] = __double2float_rn(__dadd_rn(q.x, (double)s0));
            fxy[o + 2] = __double2float_rn(__dadd_rn(q.y, (double)s1));
            fzt[o] = __double2float_rn(__dadd_rn(q.z, (double)s2));
            fzt[o + 2] = __int_as_float(MT ? A.ityp[gst + a] : 1);
        }
    }
    __syncthreads();
    // TWO warps per owned cell (lanes = its atoms): warp 0 scans halo rows 0-3, warp 1 rows 4-8 (with the atom's own row).
    // Both fill the SAME shared-memory column of an atom, warp 0 from row 0 upwards, warp 1 from the last row downwards, so
    // the split of an atom's neighbours between the two row sets (anything from 25 % to 75 %, by where the atom sits in its
    // cell) needs no capacity of its own; only the sum can overflow, and that is checked after both are done.  Warp 0 then
    // moves warp 1's block down next to its own and does the class partition and the write-out.
    // (One warp per cell left 14 warps per SM: 63 % of the issue slots used.)
    // A CTA is launched with `nwarps / 2` warp pairs for tiles of up to wmax cells: where cells hold more than 32 atoms and the
    // tile is narrow (3 x 3 x 3-cell PARREP boxes: one 74-atom cell per tile), `ppc` pairs share a cell and take its atoms in
    // blocks of 32 in turn, each pair with a column block of its own.
    const int wt = nhx - 2;
    const int ppc = max((nwarps >> 1) / wt, 1);
    const int pid = warp >> 1, half = warp & 1;
    const int cw = pid / ppc, blk = pid - cw * ppc;
    if (cw >= wt) return;
    const int myhc = 4 * nhx + cw + 1;             // centre row (hy = hz = 1), cell hx = cw + 1
    const int ccnt = H.cnt[myhc];
    if (ccnt <= 0) return;                         // :1018
    const int cgst = H.gst[myhc], csl = H.slot[myhc];
    if (A.naac[H.cid[myhc]] <= 0) {                // cells without ACTIVE atoms are skipped (:981-982): empty lists
        if (!half && blk == 0)
            for (int a = lane; a < ccnt; a += 32) { A.kvois[cgst + a] = 0; A.ncls[cgst + a] = 0; A.ncls[cgst + a + P.npad] = 0; }
        return;
    }
    const float rm1 = A.rm2[0], rc0 = A.rc2[0], rc1 = A.rc2[1];
    unsigned short *col0 = lists + (size_t)pid * A.lcap * 32 + lane;   // entry k of this lane at col0[k * 32]
    unsigned short *col = col0;
    const int lcap = A.lcap;
    const unsigned col_a = (unsigned)__cvta_generic_to_shared(col0) + (half ? (unsigned)(lcap - 1) * 64u : 0u); // first entry of this warp
    const int step = half ? -64 : 64;
    const int r_lo = half ? 4 : 0, r_hi = half ? 9 : 4;

    for (int ab = blk * 32; ab < ccnt; ab += 32 * ppc) {
        const bool valid = ab + lane < ccnt;
        const int ia = cgst + ab + (valid ? lane : 0);
        const int myslot = csl + ab + (valid ? lane : 0);
        const int mo = (myslot >> 1) * 4 + (myslot & 1);
        const float mx = fxy[mo], my = fxy[mo + 2], mz = fzt[mo];   // POS = (float)XP_i :1080-1082 (own cells carry no shift)
        const int ity = MT ? __float_as_int(fzt[mo + 2]) : 1;
        const unsigned long long mx2 = f2_pack(mx, mx), my2 = f2_pack(my, my), mz2 = f2_pack(mz, mz);
        unsigned pa = col_a;                           // shared address of the next free entry of this lane
        const int room = valid ? lcap : 0;             // invalid lanes accept nothing (capacity-checked path)
        int nover = 0;                                 // accepted beyond the capacity

        // one candidate / one aligned pair of candidates; SELF: the range holds the lane's own atom (:1124);
        // CAP: the list may fill up inside this range
        auto accept = [&](auto self_tag, auto cap_tag, const int s, const float r2, const float rm) {
            constexpr bool SELF = decltype(self_tag)::value, CAP = decltype(cap_tag)::value;
            if (!CAP) {                                                     // the common case: room for the whole range
                if (SELF) push_if_in_notself(pa, (unsigned)s, r2, rm, rc0, rc1, (unsigned)myslot, step);
                else push_if_in(pa, (unsigned)s, r2, rm, rc0, rc1, step);
                return;
            }
            bool hit = r2 <= rm;                                            // :1123
            if (SELF) hit = hit && (s != myslot);
            const unsigned e = (unsigned)s + ((r2 > rc0) ? 0x4000u : 0u) + ((r2 > rc1) ? 0x4000u : 0u);
            const int have = half ? (int)((col_a - pa) >> 6) : (int)((pa - col_a) >> 6);
            const bool st = hit && (have < room);
            nover += (hit && !st) ? 1 : 0;
            if (st) {
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(pa), "h"((unsigned short)e) : "memory");
                pa = (unsigned)((int)pa + step);
            }
        };
        auto one = [&](auto self_tag, auto cap_tag, const int s) {
            const int o = (s >> 1) * 4 + (s & 1);
            const float e1 = __fsub_rn(mx, fxy[o]), e2 = __fsub_rn(my, fxy[o + 2]), e3 = __fsub_rn(mz, fzt[o]);
            const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(e1, e1), __fmul_rn(e2, e2)), __fmul_rn(e3, e3));
            const float rm = MT ? A.rm2[(ity - 1) + P.ng * (__float_as_int(fzt[o + 2]) - 1)] : rm1;
            accept(self_tag, cap_tag, s, r2, rm);
        };
        auto range = [&](auto self_tag, auto cap_tag, int s, const int s_hi) {
            if (s < s_hi && (s & 1)) { one(self_tag, cap_tag, s); s++; }
            const float4 *pxy = hxy + (s >> 1);
#pragma unroll 2
            for (; s + 1 < s_hi; s += 2, pxy++) {
                const float4 a = pxy[0], b = pxy[npair];
                const unsigned long long ex = f2_sub(mx2, f2_pack(a.x, a.y)), ey = f2_sub(my2, f2_pack(a.z, a.w)),
                                         ez = f2_sub(mz2, f2_pack(b.x, b.y));
                const unsigned long long qx = f2_sq(ex), qy = f2_sq(ey), qz = f2_sq(ez);
                const float r20 = __fadd_rn(__fadd_rn(f2_lo(qx), f2_lo(qy)), f2_lo(qz));
                const float r21 = __fadd_rn(__fadd_rn(f2_hi(qx), f2_hi(qy)), f2_hi(qz));
                const float rma = MT ? A.rm2[(ity - 1) + P.ng * (__float_as_int(b.z) - 1)] : rm1;
                const float rmb = MT ? A.rm2[(ity - 1) + P.ng * (__float_as_int(b.w) - 1)] : rm1;
                accept(self_tag, cap_tag, s, r20, rma);
                accept(self_tag, cap_tag, s + 1, r21, rmb);
            }
            if (s < s_hi) one(self_tag, cap_tag, s);
        };
#pragma unroll 1
        for (int r = r_lo; r < r_hi; r++) {
            const int hc0 = r * nhx + cw;              // cells hx = cw, cw+1, cw+2 of halo row r: contiguous slots
            const int s_lo = H.slot[hc0], s_hi = H.slot[hc0 + 3];
            // warp-uniform choice: can any lane run out of list rows inside this range?
            const int used = half ? (int)((col_a - pa) >> 6) : (int)((pa - col_a) >> 6);
            const bool tight = __any_sync(0xffffffffu, used + (s_hi - s_lo) > lcap);
            if (r == 4) {
                if (tight) range(std::true_type(), std::true_type(), s_lo, s_hi);
                else range(std::true_type(), std::false_type(), s_lo, s_hi);
            } else {
                if (tight) range(std::false_type(), std::true_type(), s_lo, s_hi);
                else range(std::false_type(), std::false_type(), s_lo, s_hi);
            }
        }
        int nn = half ? (int)((col_a - pa) >> 6) : (int)((pa - col_a) >> 6);
        if (!valid) { nn = 0; nover = 0; }
        if (half) s_half1[pid][lane] = (unsigned short)(min(nn, 1023) | (min(nover, 63) << 10));
        pair_barrier(pid);                             // warp 1's entries and counts are visible to warp 0
        if (half) { pair_barrier(pid); continue; }     // ... and warp 1 waits until warp 0 is done with the columns
        {
            const int h1 = s_half1[pid][lane];
            int n1 = h1 & 1023;
            nover += h1 >> 10;
            if (nn + n1 > lcap) { nover += nn + n1 - lcap; n1 = lcap - nn; }   // the two ends met: the build is discarded
            // warp 1's block [lcap - n1, lcap) moves down to [nn, nn + n1): ascending copy to lower rows, overlap-safe
            for (int k = 0; k < n1; k++) col0[(nn + k) * 32] = col0[(lcap - n1 + k) * 32];
            nn += n1;
        }
        const int nall = nn + nover;
        if (__any_sync(0xffffffffu, nover > 0 || nall > P.mxkvois)) {
            if (lane == 0) atomicAdd(&A.counters[CNT_TILE_OVERFLOW], 1);
        }
        {
            int m = nall;
            for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
            if (lane == 0) atomicMax(&A.counters[CNT_NNMAX], m);
        }
        if (valid) { // (no early exit of single lanes: the pair barrier below needs the whole warp)
        A.kvois[ia] = min(nn, P.mxkvois);
        // ---- three-way partition of the column by class tag (order inside a class is free)
        int lo = 0, mid = 0, hi = nn - 1;
        while (mid <= hi) { // branch-free: the entry at mid swaps with lo (class 0), itself (class 1) or hi (class 2)
            const unsigned e = col[mid * 32];
            const unsigned t = e >> 14;
            const int j = (t == 0u) ? lo : ((t == 1u) ? mid : hi);
            const unsigned f = col[j * 32];
            col[j * 32] = (unsigned short)e;
            col[mid * 32] = (unsigned short)f;
            lo += (t == 0u) ? 1 : 0;
            mid += (t <= 1u) ? 1 : 0;
            hi -= (t >= 2u) ? 1 : 0;
        }
        A.ncls[ia] = (unsigned short)lo;               // class 0
        A.ncls[ia + P.npad] = (unsigned short)mid;     // classes 0+1
        // ---- write out: 4G consecutive entries form one contiguous 8G-byte block [gl][m%4] of the layout
        const unsigned pad = (unsigned)P.hcap;         // list tails point at the dummy record (k_tile_deal re-aims them by residue)
        for (int q = 0; q * 4 * G < nn; q++) {
            unsigned short blk[4 * G];
#pragma unroll
            for (int j = 0; j < 4 * G; j++) {
                const int k = q * 4 * G + j;
                blk[(j % G) * 4 + j / G] = (k < nn) ? (unsigned short)(col[k * 32] & 0x3fffu) : (unsigned short)pad;
            }
            uint4 *dst = reinterpret_cast<uint4 *>(A.nbl + ((((size_t)q * P.npad + (size_t)ia) * G) << 2));
#pragma unroll
            for (int v = 0; v < (4 * G) / 8; v++) {
                uint4 w;
                w.x = blk[8 * v] | ((unsigned)blk[8 * v + 1] << 16);
                w.y = blk[8 * v + 2] | ((unsigned)blk[8 * v + 3] << 16);
                w.z = blk[8 * v + 4] | ((unsigned)blk[8 * v + 5] << 16);
                w.w = blk[8 * v + 6] | ((unsigned)blk[8 * v + 7] << 16);
                dst[v] = w;
            }
        }
        }
        pair_barrier(pid);
    }
}

// ---- bank-aware order of the stored lists (MDB_OPT_TILED_BANKORDER), one thread per owned atom.
// A pass reads the staged 32-byte record of a slot as two LDS.128 (even lanes {x,y} first, odd lanes {z,den} first); the
// hardware serves an LDS.128 per HALF-warp and takes max(2, lanes per 16-byte bank group) cycles for it
// (tools/micro/lds128_patterns.cu), so a list position is conflict-free when the 8 even and the 8 odd lanes of a half-warp
// each hold every slot residue (mod 4) exactly twice.  Lane lam of the half-warp therefore WANTS residue ((lam >> 1) + row) & 3
// at list row `row`.  Each of the two class segments a pass scans ([0, n0) and [n0, n1)) is counting-sorted by residue and dealt
// back: entry position e takes the next entry of the residue it wants; positions that find none are filled with the leftovers
// afterwards (so every residue serves as many of its positions as it can).  The order
// inside a class is free, so nothing but the order changes.  (Round 2 first did this inside the list kernel, on one of the two
// warps of a cell with strided 2-byte shared-memory accesses: +0.58 ms per rebuild.  Here every lane works, the lists are read
// and written as whole 32-byte blocks, and an atom's entries sit in its own conflict-free shared-memory column.)
// 8 * list_wanted<G>(lam0, e) as a shift count; for G = 4 the residue is (lam0/2 + ((e + 2) >> 2)) & 3
template <int G>
__device__ __forceinline__ unsigned wanted_sh(int lam0, int e)
{
    if (G == 4) return (unsigned)(((lam0 >> 1) + ((e + 2) >> 2)) & 3) << 3;
    return list_wanted<G>(lam0, e) << 3;
}
#define DEAL_T 128     // threads per CTA
#define DEAL_E 80      // entries per atom that can be ordered: an atom whose classes 0+1 hold more keeps the built order
template <int G>
__global__ void __launch_bounds__(DEAL_T)
k_tile_deal(TileParams P, const TileDesc *__restrict__ desc, int a0, int a1, int tile_lo, int tile_hi, const int *__restrict__ ic,
            unsigned short *__restrict__ nbl, const unsigned short *__restrict__ ncls, const int *__restrict__ kvois)
{
    __shared__ unsigned short sa[DEAL_E * DEAL_T], sb[DEAL_E * DEAL_T];
    const int t = threadIdx.x;
    unsigned short *A0 = sa + t, *B0 = sb + t;        // entry e of this thread at A0[e * DEAL_T]
    {
        const int ia = a0 + blockIdx.x * DEAL_T + t;
        if (ia >= a1) return;
        const int cell = ic[ia] - 1;                  // atoms parked outside the cells have no list
        if (cell < 0) return;
        const int n0 = ncls[ia], n1 = ncls[ia + P.npad], nn = min(kvois[ia], P.mxkvois);
        if (n1 <= 0 || ((n1 + 4 * G - 1) / (4 * G)) * 4 * G > DEAL_E) return;   // whole blocks are staged
        // the tile of the atom -> its index among the tile's owned atoms -> the lanes it gets in a pass
        const int box = cell / P.nc0, rem = cell - box * P.nc0;
        const int ix = rem % P.ncx, iyz = rem / P.ncx, iy = iyz % P.ncy, iz = iyz / P.ncy;
        const int tx = ((ix + 1) * P.ntx - 1) / P.ncx;          // the largest tx with (tx * ncx) / ntx <= ix (tile_geom)
        const int tile = ((box * P.ncz + iz) * P.ncy + iy) * P.ntx + tx;
        if (tile < tile_lo || tile >= tile_hi) return;          // (slab decomposition: lists exist for this rank's tiles only)
        const int o = ia - desc[tile].own_start;
        const int nblk = (n1 + 4 * G - 1) / (4 * G);
        const int lam0 = (o % (16 / G)) * G;          // first lane of this atom inside its half-warp (as the passes map atoms)
        // ---- load the blocks that hold classes 0 + 1 (entry j of a block sits at [(j % G)][j / G])
        for (int q = 0; q < nblk; q++) {
            const uint4 *src = reinterpret_cast<const uint4 *>(nbl + ((((size_t)q * P.npad + (size_t)ia) * G) << 2));
#pragma unroll
            for (int v = 0; v < (4 * G) / 8; v++) {
                const uint4 w = src[v];
                const unsigned u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int h = 0; h < 8; h++) {
                    const int p = 8 * v + h, j = (p % 4) * G + p / 4;
                    A0[(q * 4 * G + j) * DEAL_T] = (unsigned short)((u[h >> 1] >> (16 * (h & 1))) & 0xffffu);
                }
            }
        }
        // ---- the two class segments
#pragma unroll 1
        for (int seg = 0; seg < 2; seg++) {
            const int a = seg ? n0 : 0, b = seg ? n1 : n0;
            if (b - a < 2) continue;
            unsigned cnt = 0u;                        // four 8-bit counters (b - a <= DEAL_E < 256)
#pragma unroll 4
            for (int e = a; e < b; e++) cnt += 1u << ((A0[e * DEAL_T] & 3u) << 3);
            const unsigned c0 = cnt & 255u, c1 = (cnt >> 8) & 255u, c2 = (cnt >> 16) & 255u;
            const unsigned start = (unsigned)a * 0x01010101u + ((c0 << 8) | ((c0 + c1) << 16) | ((c0 + c1 + c2) << 24));
            unsigned cur = start;
#pragma unroll 4
            for (int e = a; e < b; e++) {             // counting sort by residue into B
                const unsigned v = A0[e * DEAL_T], sh = (v & 3u) << 3;
                B0[((cur >> sh) & 255u) * DEAL_T] = (unsigned short)v;
                cur += 1u << sh;
            }
            cur = start;
            unsigned have = cnt;
            int holes = 0;
#pragma unroll 2
            for (int e = a; e < b; e++) {             // deal: every position takes the residue it wants while that lasts ...
                const unsigned tw = wanted_sh<G>(lam0, e);
                const bool got = ((have >> tw) & 255u) != 0u;
                const unsigned v = B0[min((cur >> tw) & 255u, (unsigned)(DEAL_E - 1)) * DEAL_T]; // (an exhausted residue's cursor may sit at row b)
                A0[e * DEAL_T] = got ? (unsigned short)v : (unsigned short)0xffffu;
                cur += got ? 1u << tw : 0u;
                have -= got ? 1u << tw : 0u;
                holes += got ? 0 : 1;
            }
            for (int e = a; holes > 0 && e < b; e++) { // ... and the positions left empty take what is left over, lowest residue first
                if (A0[e * DEAL_T] != 0xffffu) continue;
                const unsigned r = (unsigned)(__ffs((int)have) - 1) & ~7u;
                A0[e * DEAL_T] = B0[((cur >> r) & 255u) * DEAL_T];
                cur += 1u << r; have -= 1u << r;
                holes--;
            }
        }
        // list tails inside these blocks: one dummy record per residue (hcap .. hcap + 3)
        for (int e = nn; e < nblk * 4 * G; e++) A0[e * DEAL_T] = (unsigned short)(P.hcap + (int)list_wanted<G>(lam0, e));
        // ---- store
        for (int q = 0; q < nblk; q++) {
            uint4 *dst = reinterpret_cast<uint4 *>(nbl + ((((size_t)q * P.npad + (size_t)ia) * G) << 2));
#pragma unroll
            for (int v = 0; v < (4 * G) / 8; v++) {
                unsigned u[4];
#pragma unroll
                for (int h = 0; h < 8; h += 2) {
                    const int p = 8 * v + h, j0 = (p % 4) * G + p / 4, p1 = p + 1, j1 = (p1 % 4) * G + p1 / 4;
                    u[h >> 1] = (unsigned)A0[(q * 4 * G + j0) * DEAL_T] | ((unsigned)A0[(q * 4 * G + j1) * DEAL_T] << 16);
                }
                dst[v] = make_uint4(u[0], u[1], u[2], u[3]);
            }
        }
    }
}

// =====================================================================================
// force passes
// =====================================================================================
// One persistent CTA per SM: NT/32 - 1 CONSUMER warps and one PRODUCER warp around a two-stage
// shared-memory pipeline of halo tiles.
//   producer  per tile: waits until the stage is free (mbarrier "empty"), issues one TMA bulk copy per halo
//             cell (contiguous run of {x,y,z,den} records) plus one for the first index group of the owned
//             atoms, fetches the per-atom scan counts, moves cells on a periodic face into the tile frame,
//             then publishes the stage (mbarrier "ready").  It runs one tile ahead of the consumers.
//   consumers grab chunks of 32/G owned atoms from a shared counter (no CTA-wide barrier anywhere: a warp that
//             runs out of chunks moves on to the next stage by itself) and stream each atom's class-ordered
//             slot list: first 4-entry group from shared memory, later groups from global memory one group
//             ahead.  Every listed entry is evaluated -- fp64 separation from the staged records, exact
//             r2 <= r_eff^2, one MUFU.RSQ64H + Halley step each for 1/r and 1/sqrt(r), table rows from
//             shared memory -- and masked by the range test.  List tails are padded with a dummy slot whose
//             record sits at 1e30, so no per-entry validity logic exists.
struct TilePassArgs {
    double4 *pos; const int *ityp; const int *statu; const int *kvois; const unsigned short *ncls;
    const unsigned short *nbl; double *fp; int *counters; const TileDesc *desc;
    // tables: global packed {T[kk], T[kk+1]-T[kk]} (stride ntab+2 per kind) for the rare fall-backs
    const double2 *g_potb, *g_fpotr, *g_fpotb, *g_dfembd;
    const double2 *g_potr, *g_fembd;   // PASS 3 (per-atom energy): pair term and embedding VALUE tables
    double *epot;
    double *vpart;         // VIR: per-warp partial virial tensors (9 doubles each), summed by mdb_virial_finish
    int den_too;           // PASS 3 also stores dF/drho (what pass 1 produces): its density sum is the same sum
    int ntab, nembd, pot_type;
    double csi, rhod;
    double r2eff;          // min(RU2, table support) for this pass
    int kmin, ktab;        // shared-memory table window: rows kmin .. kmin+ktab (row kk and kk+1 are read)
    int kind0;             // the kind held in shared memory = KPAIR(1,1)
    // which per-atom scan count a pass uses, by the max |displacement since the rebuild|^2 (device counter): class-count row
    // row_a while d2 <= safe_a, else row row_b while d2 <= safe_b, else the full list (KVOIS)
    float safe_a, safe_b;
    int row_a, row_b;
    // per-tile displacement bound (k_tile_d2, cascade runs): when set, a tile decides on ITS halo's maximum instead of the global one
    const float *tile_d2;
    int fuse;              // fused epilogue of pass 2 (mdb_run): bit 0 EPC friction, bit 1 corrector half-kick
    int tile_lo, tile_hi;  // tiles of this rank (slab decomposition), [0, ntiles) otherwise
    int tile_lo2, tile_hi2; // an optional second range (boundary layers of a slab: first and last layer in one launch)
    int zero_parked;       // block 0 zeroes the outputs of atoms parked outside the cells
    int nbuf;              // pipeline stages (2 or 3)
    const int *skip;       // device flag: return at once when set (converged quench iterations)
    double hs2;            // H/2
    double *xp1;
    EpcParams epc;
    MassParams mass;
    int kpair[MDB_MXGROUP * MDB_MXGROUP];
    int kembd[MDB_MXGROUP];
};

__device__ __forceinline__ double rsqrt_fast(double a)
{
    // MUFU.RSQ64H seed (>= 20 good bits) + one Halley step: y = y0 (1 + e/2 + 3e^2/8), e = 1 - a y0^2.
    // Residual 5/16 e^3 < 2^-60: the result is within an ulp of the correctly rounded value.
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
    const double t = a * y0;
    const double e = fma(-t, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y0 * e, p, y0);
}

__device__ __forceinline__ double lerp_g(const double2 *__restrict__ t, int stride, int k, int kk, double dk)
{
    kk = min(max(kk, 0), stride - 1);
    const double2 e = __ldg(t + (size_t)k * stride + kk);
    return fma(dk, e.y, e.x);
}
// rare out-of-line path of the passes: a pair whose table row lies outside the shared-memory window, or whose
// kinds are not KPAIR(1,1).  Same arithmetic as the in-line path, tables read from global memory.
// PASS 1: ta = POTB, returns the density term.  PASS 2: ta = FPOTR, tb = FPOTB, returns FORTOT.
template <int PASS>
__device__ __noinline__ double pair_slow(double4 me, double4 pj, double csi, const double2 *__restrict__ ta,
                                         const double2 *__restrict__ tb, int stride, int k0, int k1)
{
    const double sx = me.x - pj.x, sy = me.y - pj.y, sz = me.z - pj.z;
    const double r2 = fma(sz, sz, fma(sy, sy, sx * sx));
    const double y = rsqrt_fast(r2);
    const double r = r2 * y;
    const double z = rsqrt_fast(r);
    const double sk = (r * z) * csi;
    const double tk = __dadd_rd(sk, 4503599627370496.0);
    const int kk = __double2loint(tk);
    const double dk = sk - (tk - 4503599627370496.0);
    if (PASS == 1) return lerp_g(ta, stride, k0, kk, dk);
    const double fr = lerp_g(ta, stride, k0, kk, dk);
    const double fb0 = lerp_g(tb, stride, k0, kk, dk);
    const double fb1 = (k1 == k0) ? fb0 : lerp_g(tb, stride, k1, kk, dk);
    return y * fma(fr, y, fma(fb0, me.w, fb1 * pj.w));
}

// PASS 3 twin: (POTR/r, POTB) of one pair from the global tables, kind kt = KPAIR(ITYP_i, ITYP_j) for both (:1620-1623)
__device__ __noinline__ double2 pair_slow_epot(double4 me, double4 pj, double csi, const double2 *__restrict__ tr,
                                               const double2 *__restrict__ tb, int stride, int kt)
{
    const double sx = me.x - pj.x, sy = me.y - pj.y, sz = me.z - pj.z;
    const double r2 = fma(sz, sz, fma(sy, sy, sx * sx));
    const double y = rsqrt_fast(r2);
    const double r = r2 * y;
    const double z = rsqrt_fast(r);
    const double sk = (r * z) * csi;
    const double tk = __dadd_rd(sk, 4503599627370496.0);
    const int kk = __double2loint(tk);
    const double dk = sk - (tk - 4503599627370496.0);
    return make_double2(lerp_g(tr, stride, kt, kk, dk) * y, lerp_g(tb, stride, kt, kk, dk));
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- shared-memory plan of the pass kernel (host and device agree through these)
#define TP_MAXBUF 3
#define TP_HDR_BYTES 256 // mbarriers [full, ready, empty] x stages, stage info, chunk counters
__host__ __device__ __forceinline__ size_t al128(size_t x) { return (x + 127) & ~(size_t)127; }
__host__ __device__ __forceinline__ size_t tp_pos_bytes(int hcap) { return al128(sizeof(double4) * (size_t)(hcap + 4)); } // +4: dummy records (one per slot residue)
__host__ __device__ __forceinline__ size_t tp_idx_bytes(int ocap, int G) { return al128(sizeof(uint2) * (size_t)ocap * G); }
// 16-byte aligned windows of STATU / KVOIS (int32) and of the class counts (uint16) around the owned range
__host__ __device__ __forceinline__ size_t tp_i32_bytes(int ocap) { return al128(sizeof(int) * (size_t)(ocap + 8)); }
__host__ __device__ __forceinline__ size_t tp_u16_bytes(int ocap) { return al128(sizeof(unsigned short) * (size_t)(ocap + 16)); }
__host__ __device__ __forceinline__ size_t tp_buf_bytes(int hcap, int ocap, int G, bool mt)
{
    return tp_pos_bytes(hcap) + tp_idx_bytes(ocap, G) + 2 * tp_i32_bytes(ocap) + tp_u16_bytes(ocap) + (mt ? al128((size_t)hcap + 4) : 0);
}
__host__ __device__ __forceinline__ size_t tp_tab_bytes(int ktab) { return al128(sizeof(double2) * (size_t)(ktab + 2)); }

// PASS 1: rho -> DEN.  PASS 2: forces.  G lanes per atom.  MT: more than one atom type.  FUSE: EPC + corrector epilogue.
template <int PASS, int G, bool MT, bool FUSE, int NT, bool VIR = false>
__global__ void __launch_bounds__(NT, 1)
k_tile_pass(TileParams P, TilePassArgs A)
{
    extern __shared__ __align__(128) unsigned char smem[];
    // Programmatic dependent launch: the kernel that follows in the stream may start its prologue while this grid drains, and
    // this grid's own prologue (table window, barriers: constant data only) runs under the tail of the kernel before it;
    // nothing that a previous kernel produced is touched before cudaGridDependencySynchronize() below.
    cudaTriggerProgrammaticLaunchCompletion();
    constexpr int NW = NT / 32, NCW = NW - 1, APW = 32 / G;
    unsigned long long *bar_full = reinterpret_cast<unsigned long long *>(smem);  // TMA bytes landed
    unsigned long long *bar_ready = bar_full + TP_MAXBUF;                        // producer finished the stage
    unsigned long long *bar_empty = bar_ready + TP_MAXBUF;                       // every consumer warp left the stage
    int *s_info = reinterpret_cast<int *>(smem + 96);                             // [stage][own_start, own_slot0, own_count, edge, scan mode]
    int *s_ctr = reinterpret_cast<int *>(smem + 160);                             // [stage] next owned atom to hand out
    double2 *s_tab = reinterpret_cast<double2 *>(smem + TP_HDR_BYTES);
    unsigned char *buf0 = smem + TP_HDR_BYTES + tp_tab_bytes(A.ktab);
    const size_t bufb = tp_buf_bytes(P.hcap, P.ocap, G, MT);
    const size_t o_idx = tp_pos_bytes(P.hcap), o_stat = o_idx + tp_idx_bytes(P.ocap, G), o_kvo = o_stat + tp_i32_bytes(P.ocap),
                 o_ncl = o_kvo + tp_i32_bytes(P.ocap), o_typ = o_ncl + tp_u16_bytes(P.ocap);
    const int nbuf = A.nbuf;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G;      // lane within the atom's group
    // VIR (CALPTENSOR_KERNEL, MD_EAM_ForceTable_GPU.F90:1222-1232): per-lane partial sums of 1/2 FORTOT s_a s_b over every pair
    // this lane evaluates in the whole launch (xx, xy, xz, yy, yz, zz: the tensor is symmetric); reduced once at the end
    double vxx = 0.0, vxy = 0.0, vxz = 0.0, vyy = 0.0, vyz = 0.0, vzz = 0.0;

    // ---- tables for kind0, rows kmin..kmin+ktab : staged once per (persistent) CTA
    //      pass 1: {POTB[kk], POTB[kk+1]}         (one 16-byte read per pair)
    //      pass 2: {FPOTR[kk], FPOTB[kk]}         (rows kk and kk+1: two 16-byte reads)
    {
        const int stride = A.ntab + 2;
        for (int r = threadIdx.x; r <= A.ktab + 1; r += NT) {
            const int kk = min(A.kmin + r, stride - 1);
            if (PASS == 1) {
                const int k1 = min(kk + 1, stride - 1);
                s_tab[r] = make_double2(A.g_potb[(size_t)A.kind0 * stride + kk].x, A.g_potb[(size_t)A.kind0 * stride + k1].x);
            } else if (PASS == 2) {
                s_tab[r] = make_double2(A.g_fpotr[(size_t)A.kind0 * stride + kk].x, A.g_fpotb[(size_t)A.kind0 * stride + kk].x);
            } else { // PASS 3: {POTR[kk], POTB[kk]}, rows kk and kk+1 are read
                s_tab[r] = make_double2(A.g_potr[(size_t)A.kind0 * stride + kk].x, A.g_potb[(size_t)A.kind0 * stride + kk].x);
            }
        }
        if (threadIdx.x == 0) {
            for (int b = 0; b < nbuf; b++) {
                mbar_init(&bar_full[b], 1);
                mbar_init(&bar_ready[b], 32);
                mbar_init(&bar_empty[b], NCW);
                // the dummy record list tails point at: far from everything, finite
                for (int q = 0; q < 4; q++) { // hcap is a multiple of 4: dummy q has slot residue q (bank-aware list padding)
                    reinterpret_cast<double4 *>(buf0 + b * bufb)[P.hcap + q] = make_double4(1.0e30, 1.0e30, 1.0e30, 0.0);
                    if (MT) (buf0 + b * bufb + o_typ)[P.hcap + q] = 0;
                }
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    cudaGridDependencySynchronize();
    __syncthreads();
    if (A.skip && *A.skip) return;   // converged quench iteration (uniform over the grid)
    // distance classes are usable while no atom (of the tile's halo, when the per-tile bounds are given) has moved more than half
    // the class margin since the rebuild.  mode: class-count row 0 / 1, or 2 = the full list (KVOIS)
    const float d2glob = __int_as_float(A.counters[CNT_D2MAX]);
    auto mode_of = [&](const float d2) -> int { return d2 <= A.safe_a ? A.row_a : (d2 <= A.safe_b ? A.row_b : 2); };
    const uint2 *nbl2 = reinterpret_cast<const uint2 *>(A.nbl);

    const int nt1 = A.tile_hi - A.tile_lo, nvt = nt1 + (A.tile_hi2 - A.tile_lo2); // tiles of this launch: range 1, then range 2
    if (warp == NCW) {
        // =========================== producer warp ===========================
        // Everything a stage needs arrives by TMA: halo runs, the first index group, and 16-byte aligned windows of
        // STATU and of the scan counts around the owned range.  The descriptor of the NEXT tile is loaded one tile
        // ahead and only used in the next iteration, so no global-memory round trip sits on the per-tile path.
        struct Pre { int htot, nrun, edge_any, own_start, own_slot0, own_count, rs0, rs1, rgst; float d2; };
        auto prefetch = [&](int v) {
            Pre q;
            q.htot = 0; q.nrun = 0; q.edge_any = 0; q.own_start = 0; q.own_slot0 = 0; q.own_count = 0; q.rs0 = 0; q.rs1 = 0; q.rgst = 0;
            q.d2 = d2glob;
            if (v < nvt) {
                const int tile = v < nt1 ? A.tile_lo + v : A.tile_lo2 + (v - nt1);
                const TileDesc &D = A.desc[tile];
                if (A.tile_d2) q.d2 = A.tile_d2[tile];
                q.htot = D.htot; q.nrun = D.nrun; q.edge_any = D.edge_any;
                q.own_start = D.own_start; q.own_slot0 = D.own_slot0; q.own_count = D.own_count;
                q.rs0 = D.rslot[lane]; q.rs1 = D.rslot[lane + 1]; q.rgst = D.rgst[lane]; // lane r: run r (entries past nrun unused)
            }
            return q;
        };
        int b = 0;
        unsigned u = 0;
        Pre nx = prefetch(blockIdx.x);
        for (int v = blockIdx.x; v < nvt; v += gridDim.x) {
            const int tile = v < nt1 ? A.tile_lo + v : A.tile_lo2 + (v - nt1);
            const Pre cur = nx;
            int own_count = cur.own_count;
            const bool fits = cur.htot <= P.hcap && own_count <= P.ocap && cur.nrun <= TILE_MAX_RUN;
            const int mode = mode_of(cur.d2);
            const bool safe = mode < 2;
            const size_t ncl_base = (size_t)(mode & 1) * P.npad; // flat index of the class count of owned atom 0 (uint16 [2][npad]) = ncl_base + own_start
            mbar_wait(&bar_empty[b], (u & 1u) ^ 1u);
            unsigned char *bp = buf0 + b * bufb;
            double4 *sp = reinterpret_cast<double4 *>(bp);
            if (fits) {
                // aligned windows: int32 arrays in units of 4 atoms, the uint16 counts in units of 8
                const int a4 = cur.own_start & ~3, n4b = (((cur.own_start + own_count + 3) & ~3) - a4) * 4;
                const size_t c0 = ncl_base + (size_t)cur.own_start, c8 = c0 & ~(size_t)7;
                const int n8b = (int)(((c0 + own_count + 7) & ~(size_t)7) - c8) * 2;
                if (lane == 0) {
                    unsigned bytes = (unsigned)cur.htot * 32u;
                    if (own_count > 0) bytes += (unsigned)own_count * (unsigned)(G * 8) + (unsigned)n4b + (unsigned)(safe ? n8b : n4b);
                    mbar_expect_tx(&bar_full[b], bytes);
                }
                __syncwarp();
                fence_proxy_async(); // earlier generic accesses to the stage are ordered before the bulk copies
                const int rcnt = lane < cur.nrun ? cur.rs1 - cur.rs0 : 0;
                if (rcnt > 0) bulk_g2s(sp + cur.rs0, A.pos + cur.rgst, (unsigned)rcnt * 32u, &bar_full[b]);
                if (own_count > 0) {
                    if (lane == 0) bulk_g2s(bp + o_idx, nbl2 + (size_t)cur.own_start * G, (unsigned)own_count * (unsigned)(G * 8), &bar_full[b]);
                    if (lane == 1) bulk_g2s(bp + o_stat, A.statu + a4, (unsigned)n4b, &bar_full[b]);
                    if (lane == 2) {
                        if (safe) bulk_g2s(bp + o_ncl, A.ncls + c8, (unsigned)n8b, &bar_full[b]);
                        else bulk_g2s(bp + o_kvo, A.kvois + a4, (unsigned)n4b, &bar_full[b]);
                    }
                }
            }
            // next tile's descriptor: in flight while this tile's copies land, first used in the next iteration
            nx = prefetch(v + gridDim.x);
            if (fits) {
                if (MT) {
                    // Types of the halo atoms, one lane per halo cell: the cell's atoms are read as aligned int4 windows of ITYP
                    // (a window may start up to 3 atoms before the cell and end up to 3 after it: ITYP is allocated with padding),
                    // TYB windows in flight before the first value is stored.  A single warp under 23 busy consumer warps pays
                    // about as much per instruction as per L2 round trip / 50, so both counts matter: walking the atoms one by one,
                    // four loads in flight (the first version), took atoms-per-cell/4 round trips -- with the 74-atom cells of
                    // the PARREP boxes the producer was the bottleneck of both passes (0.15 ms for 192 k atoms); a warp-wide
                    // walk over the runs needs one round trip per tile but ~70 instructions per 128 atoms and was slower still.
                    constexpr int TYB = 8;
                    const TileDesc &D = A.desc[tile];
                    unsigned char *styp = bp + o_typ;
                    const int nhc = D.nhc;
                    for (int hc = lane; hc < nhc; hc += 32) {
                        const int cnt = D.cnt[hc], gst = D.gst[hc];
                        const int off = gst & 3, nw = cnt > 0 ? (off + cnt + 3) >> 2 : 0;
                        const int4 *src = reinterpret_cast<const int4 *>(A.ityp) + (gst >> 2);
                        unsigned char *dst = styp + (D.slot[hc] - off);   // element e of window w is atom 4w + e - off of the cell
                        for (int w0 = 0; w0 < nw; w0 += TYB) {
                            int4 ty[TYB];
#pragma unroll
                            for (int j = 0; j < TYB; j++) {
                                ty[j] = make_int4(1, 1, 1, 1);
                                if (w0 + j < nw)
                                    asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];"
                                                 : "=r"(ty[j].x), "=r"(ty[j].y), "=r"(ty[j].z), "=r"(ty[j].w) : "l"(src + w0 + j));
                            }
                            // every load of the batch is issued before the first value is used (the compiler otherwise converts
                            // each value as soon as it arrives: one round trip per window)
                            static_assert(TYB == 8, "the barriers below name 4 * TYB registers");
#define TY4(j) "+r"(ty[j].x), "+r"(ty[j].y), "+r"(ty[j].z), "+r"(ty[j].w)
                            asm volatile("" : TY4(0), TY4(1), TY4(2), TY4(3));
                            asm volatile("" : TY4(4), TY4(5), TY4(6), TY4(7));
#undef TY4
                            const unsigned da = smem_u32(dst) + ((unsigned)w0 << 2);
#pragma unroll
                            for (int j = 0; j < TYB; j++) {
                                const int rel = ((w0 + j) << 2) - off;   // first atom of the window; windows past the cell fail every test
                                const int t0 = ty[j].x - 1, t1 = ty[j].y - 1, t2 = ty[j].z - 1, t3 = ty[j].w - 1;
                                if (rel >= 0 && rel + 3 < cnt) {         // inside the cell: the common case
                                    asm volatile("st.shared.u8 [%0], %1;\nst.shared.u8 [%0+1], %2;\nst.shared.u8 [%0+2], %3;\nst.shared.u8 [%0+3], %4;"
                                                 :: "r"(da + 4u * j), "r"(t0), "r"(t1), "r"(t2), "r"(t3) : "memory");
                                } else {                                 // first / last window of the cell: atom by atom
                                    asm volatile("{\n.reg .pred q0, q1, q2, q3;\n.reg .u32 r1, r2, r3;\n"
                                                 "add.u32 r1, %5, 1;\nadd.u32 r2, %5, 2;\nadd.u32 r3, %5, 3;\n"
                                                 "setp.lt.u32 q0, %5, %6;\nsetp.lt.u32 q1, r1, %6;\nsetp.lt.u32 q2, r2, %6;\nsetp.lt.u32 q3, r3, %6;\n"
                                                 "@q0 st.shared.u8 [%0], %1;\n@q1 st.shared.u8 [%0+1], %2;\n@q2 st.shared.u8 [%0+2], %3;\n"
                                                 "@q3 st.shared.u8 [%0+3], %4;\n}\n"
                                                 :: "r"(da + 4u * j), "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"((unsigned)rel), "r"((unsigned)cnt) : "memory");
                                }
                            }
                        }
                    }
                }
            } else {
                // counted by the list kernel; the host falls back to the generic path
                own_count = 0;
                if (lane == 0) mbar_arrive(&bar_full[b]);
            }
            if (lane == 0) {
                s_info[5 * b + 0] = cur.own_start; s_info[5 * b + 1] = cur.own_slot0; s_info[5 * b + 2] = own_count;
                s_info[5 * b + 3] = cur.edge_any; s_info[5 * b + 4] = mode;
                s_ctr[b] = 0;
            }
            mbar_arrive(&bar_ready[b]); // all 32 lanes: their writes to the stage are released to the consumers
            if (++b == nbuf) { b = 0; u++; }
        }
    } else {
        // =========================== consumer warps ===========================
        int b = 0;
        unsigned u = 0;
        for (int v = blockIdx.x; v < nvt; v += gridDim.x) {
            mbar_wait(&bar_ready[b], u & 1u);
            mbar_wait(&bar_full[b], u & 1u);
            const int own_start = s_info[5 * b + 0], own_slot0 = s_info[5 * b + 1], own_count = s_info[5 * b + 2];
            const bool edge_tile = s_info[5 * b + 3] != 0;
            const int mode = s_info[5 * b + 4];
            const bool safe = mode < 2;
            const size_t ncl_base = (size_t)(mode & 1) * P.npad;
            const unsigned char *bp = buf0 + b * bufb;
            const double4 *sp = reinterpret_cast<const double4 *>(bp);
            const uint2 *sidx = reinterpret_cast<const uint2 *>(bp + o_idx);
            const int *sstat = reinterpret_cast<const int *>(bp + o_stat) + (own_start & 3);
            const int *skvo = reinterpret_cast<const int *>(bp + o_kvo) + (own_start & 3);
            const unsigned short *sncl = reinterpret_cast<const unsigned short *>(bp + o_ncl) + (int)((ncl_base + (size_t)own_start) & 7);
            const unsigned char *styp = bp + o_typ;

            // one chunk of APW owned atoms; EDGE: the tile touches a periodic face, separations take the minimum image
            // exactly as the reference does (|SEP| > HBS -> SEP -= sign(BS, SEP), MD_EAM_ForceTable_GPU.F90:500-512)
            auto chunk = [&](auto edge_tag, const int c0) {
                constexpr bool EDGE = decltype(edge_tag)::value;
                const int o = c0 + lane / G;
                const bool have = o < own_count;
                const int ia = own_start + (have ? o : 0);
                const int myslot = have ? own_slot0 + o : P.hcap;
                const double4 me = sp[myslot];
                const int stat = have ? sstat[o] : 0;
                const bool active = (stat & ST_ACTIVE) == ST_ACTIVE;
                const int kv = active ? (safe ? (int)sncl[o] : skvo[o]) : 0;
                const int n4 = (kv + 4 * G - 1) / (4 * G);   // 4-entry index groups per lane (same on the G lanes)
                const int ti = MT ? (int)styp[myslot] : 0;
                uint2 nxt = make_uint2(0u, 0u);
                if (n4 > 0) nxt = sidx[o * G + gl];
                const size_t gstride = P.npad * G;
                const uint2 *gp = nbl2 + (gstride + (size_t)ia * G + gl);
                double vpre = 0.0;
                if (PASS == 2 && FUSE) {
                    // lane gl < 3 of the atom owns velocity/force component gl; its velocity is requested now
                    if (active && gl < 3) vpre = A.xp1[ia + (size_t)gl * P.n];
                }
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;

                // One listed entry, branch-free: the rare cases (table row outside the staged window, a pair of kinds
                // other than KPAIR(1,1)) contribute nothing here and are flagged for pair_slow below, so the four
                // entries of a group are independent straight-line code the scheduler can interleave.
                auto eval = [&](const unsigned s) -> bool {
                    // the record is read as two 16-byte halves; odd lanes fetch {z,den} first, so that each LDS.128 of the
                    // warp spreads over all eight 16-byte bank groups instead of the four a 32-byte stride allows
                    const double2 *rec = reinterpret_cast<const double2 *>(sp + s);
                    const double2 ha = rec[lane & 1], hb = rec[(lane & 1) ^ 1];
                    const double2 pxy = (lane & 1) ? hb : ha, pzw = (lane & 1) ? ha : hb;
                    double sx = me.x - pxy.x, sy = me.y - pxy.y, sz = me.z - pzw.x;
                    if (EDGE) { // P.half = BOXSIZE/2 on periodic axes, huge otherwise
                        if (fabs(sx) > P.half[0]) sx -= copysign(P.size[0], sx);
                        if (fabs(sy) > P.half[1]) sy -= copysign(P.size[1], sy);
                        if (fabs(sz) > P.half[2]) sz -= copysign(P.size[2], sz);
                    }
                    const double r2 = fma(sz, sz, fma(sy, sy, sx * sx));
                    const bool in = r2 <= A.r2eff;                // rows beyond the table support interpolate to exactly 0
                    const double y = rsqrt_fast(r2);              // 1/r
                    const double sk = rsqrt_fast(y) * A.csi;      // SK = sqrt(r)*CSI, sqrt(r) = (1/r)^(-1/2)
                    const double tk = __dadd_rd(sk, 4503599627370496.0);
                    const int kk = __double2loint(tk);            // KK = int(SK)
                    const double dk = sk - (tk - 4503599627370496.0);
                    const unsigned rr = (unsigned)(kk - A.kmin);
                    bool fast = in && rr < (unsigned)A.ktab;
                    if (MT) {
                        const int tj = (int)styp[s];
                        fast = fast && A.kpair[ti + P.ng * tj] == A.kind0 && A.kpair[tj + P.ng * ti] == A.kind0;
                    }
                    const unsigned rs = fast ? rr : 0u;
                    if (PASS == 1) {
                        const double2 t0 = s_tab[rs];
                        const double val = fma(dk, t0.y - t0.x, t0.x);
                        if (fast) acc0 += val;
                    } else if (PASS == 3) {
                        // ER0 += POTR/R ; DEN0 += POTB   (CALEPOT_KERNEL, MD_EAM_ForceTable_GPU.F90:1620-1623)
                        const double2 t0 = s_tab[rs], t1 = s_tab[rs + 1];
                        const double er = fma(dk, t1.x - t0.x, t0.x) * y;
                        const double rb = fma(dk, t1.y - t0.y, t0.y);
                        if (fast) { acc0 += er; acc1 += rb; }
                    } else {
                        const double2 t0 = s_tab[rs], t1 = s_tab[rs + 1];
                        const double fr = fma(dk, t1.x - t0.x, t0.x);
                        const double fb = fma(dk, t1.y - t0.y, t0.y);
                        // FORTOT = FPOTR/R2 + (FPOTB_ij*DEN_i + FPOTB_ji*DEN_j)/R     (:811-813)
                        const double ft = y * fma(fr, y, fma(fb, me.w, fb * pzw.y));
                        if (fast) { // predicated accumulation (no select of the 64-bit factor)
                            acc0 = fma(ft, sx, acc0);
                            acc1 = fma(ft, sy, acc1);
                            acc2 = fma(ft, sz, acc2);
                        }
                        if (VIR && fast) {
                            const double hx = 0.5 * ft * sx, hy = 0.5 * ft * sy, hz = 0.5 * ft * sz;
                            vxx = fma(hx, sx, vxx); vxy = fma(hx, sy, vxy); vxz = fma(hx, sz, vxz);
                            vyy = fma(hy, sy, vyy); vyz = fma(hy, sz, vyz); vzz = fma(hz, sz, vzz);
                        }
                    }
                    return in && !fast;
                };
                auto redo = [&](const unsigned s) {
                    double4 pj = sp[s];
                    if (EDGE) { // the image of j nearest to i
                        if (fabs(me.x - pj.x) > P.half[0]) pj.x += copysign(P.size[0], me.x - pj.x);
                        if (fabs(me.y - pj.y) > P.half[1]) pj.y += copysign(P.size[1], me.y - pj.y);
                        if (fabs(me.z - pj.z) > P.half[2]) pj.z += copysign(P.size[2], me.z - pj.z);
                    }
                    const int tj = MT ? (int)styp[s] : 0;
                    const int k0 = MT ? A.kpair[ti + P.ng * tj] : A.kind0, k1 = MT ? A.kpair[tj + P.ng * ti] : A.kind0;
                    if (PASS == 3) {
                        const double2 e = pair_slow_epot(me, pj, A.csi, A.g_potr, A.g_potb, A.ntab + 2, k0);
                        acc0 += e.x; acc1 += e.y;
                        return;
                    }
                    const double f = pair_slow<PASS == 3 ? 1 : PASS>(me, pj, A.csi, PASS == 1 ? A.g_potb : A.g_fpotr, A.g_fpotb, A.ntab + 2, k0, k1);
                    if (PASS == 1) acc0 += f;
                    else {
                        const double sx = me.x - pj.x, sy = me.y - pj.y, sz = me.z - pj.z;
                        acc0 = fma(f, sx, acc0);
                        acc1 = fma(f, sy, acc1);
                        acc2 = fma(f, sz, acc2);
                        if (VIR) {
                            const double hx = 0.5 * f * sx, hy = 0.5 * f * sy, hz = 0.5 * f * sz;
                            vxx = fma(hx, sx, vxx); vxy = fma(hx, sy, vxy); vxz = fma(hx, sz, vxz);
                            vyy = fma(hy, sy, vyy); vyz = fma(hy, sz, vyz); vzz = fma(hz, sz, vzz);
                        }
                    }
                };

                // Entry e of a list sits at group e/(4G), position (e%(4G))/G, lane e%G: position p of group `it` holds the
                // entries 4G*it + p*G .. +G-1 of every atom of the warp, so once 4G*it + p*G reaches the longest scan count
                // in the warp that position is padding for all lanes and is skipped (a warp-uniform test; in a crystal all
                // counts are equal and the tail of the last group -- 6 of 32 entries in pass 1 -- costs nothing).
                const int kvw = __reduce_max_sync(0xffffffffu, kv);
#pragma unroll 1
                for (int it = 0; it < n4; it++) {
                    const uint2 raw = nxt;
                    if (it + 1 < n4) { nxt = __ldcs(gp); gp += gstride; } // next group in flight during this one's arithmetic
                    const unsigned s0 = raw.x & 0xffffu, s1 = raw.x >> 16, s2 = raw.y & 0xffffu, s3 = raw.y >> 16;
                    const int left = kvw - it * 4 * G;   // > 0 for some lane of the warp when this lane iterates at all
                    bool w0 = false, w1 = false, w2 = false, w3 = false;
                    if (left > 3 * G) { w0 = eval(s0); w1 = eval(s1); w2 = eval(s2); w3 = eval(s3); }
                    else if (left > 2 * G) { w0 = eval(s0); w1 = eval(s1); w2 = eval(s2); }
                    else if (left > G) { w0 = eval(s0); w1 = eval(s1); }
                    else if (left > 0) { w0 = eval(s0); }
                    if (w0 | w1 | w2 | w3) {
                        if (w0) redo(s0);
                        if (w1) redo(s1);
                        if (w2) redo(s2);
                        if (w3) redo(s3);
                    }
                }
                // ---------------- reduce the G partial sums of the atom and write
#pragma unroll
                for (int w = 1; w < G; w <<= 1) {
                    acc0 += __shfl_xor_sync(0xffffffffu, acc0, w);
                    if (PASS >= 2) acc1 += __shfl_xor_sync(0xffffffffu, acc1, w);
                    if (PASS == 2) acc2 += __shfl_xor_sync(0xffffffffu, acc2, w);
                }
                if (PASS == 3) {
                    if (have && gl == 0) {
                        double den0 = 0.0;
                        if (active) { // :1625-1631; FS twin MD_FS_ForceTable_GPU.F90:1606
                            if (A.pot_type == MDB_POT_FS) den0 = -sqrt(acc1);
                            else {
                                const double sk = acc1 / A.rhod + 1.0;
                                const int kk = (int)(sk + 0.000001);
                                den0 = lerp_g(A.g_fembd, A.nembd + 2, A.kembd[ti], kk, sk - (double)kk);
                            }
                        }
                        A.epot[ia] = active ? acc0 + den0 : 0.0;
                        if (A.den_too) { // the epilogue of pass 1 on the same density sum (rows past the support add exactly 0)
                            double dd = acc1;
                            if (A.pot_type == MDB_POT_FS) {
                                if (dd > 0.0) dd = -0.5 / sqrt(dd);
                            } else if (dd > 0.0) {
                                const double sk = dd / A.rhod + 1.0;
                                const int kk = (int)(sk + 0.000001);
                                dd = lerp_g(A.g_dfembd, A.nembd + 2, A.kembd[ti], kk, sk - (double)kk);
                            }
                            reinterpret_cast<double *>(A.pos + ia)[3] = active ? dd : 0.0;
                        }
                    }
                } else if (PASS == 1) {
                    if (have && gl == 0) {
                        double den0 = acc0; // kv = 0 for inactive atoms: 0
                        if (A.pot_type == MDB_POT_FS) {
                            if (den0 > 0.0) den0 = -0.5 / sqrt(den0);
                        } else if (den0 > 0.0) {
                            const double sk = den0 / A.rhod + 1.0;
                            const int kk = (int)(sk + 0.000001);
                            if (kk > A.nembd) atomicAdd(&A.counters[CNT_RHO_OVER], 1); // rho beyond RHOMX (mdb_embed_overruns)
                            den0 = lerp_g(A.g_dfembd, A.nembd + 2, A.kembd[ti], kk, sk - (double)kk);
                        }
                        reinterpret_cast<double *>(A.pos + ia)[3] = den0;
                    }
                } else if (!FUSE) {
                    if (have && gl == 0) {
                        A.fp[ia] = acc0;
                        A.fp[ia + (size_t)P.n] = acc1;
                        A.fp[ia + 2 * (size_t)P.n] = acc2;
                    }
                } else {
                    // EPC_MOD_KERNEL (MD_EP_Coupling_GPU.F90:468-490) + Correction_KERNEL (MD_DiffScheme_GPU.F90:735-753)
                    // on the fresh force; every lane of the group holds the reduced sums, lane d < 3 handles component d
                    const int l0 = lane & ~(G - 1);
                    const double vx = __shfl_sync(0xffffffffu, vpre, l0), vy = __shfl_sync(0xffffffffu, vpre, l0 + 1),
                                 vz = __shfl_sync(0xffffffffu, vpre, l0 + 2);
                    if (have && gl < 3) {
                        double f = gl == 0 ? acc0 : (gl == 1 ? acc1 : acc2);
                        if (active) {
                            if ((A.fuse & 1) && A.epc.enable[ti] > 0) {
                                const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
                                if (v2 <= A.epc.eup[ti]) {
                                    const double tm = __dmul_rn(v2, A.epc.v2ti[ti]);
                                    const double mu = __ddiv_rn(__dmul_rn(A.epc.epa[ti], __dsub_rn(tm, A.epc.te[ti])), fmax(tm, A.epc.tcut[ti]));
                                    f = __dsub_rn(f, __dmul_rn(mu, vpre));
                                }
                            }
                            const int fixbits = (ST_FIXVELX << gl) | (ST_FIXPOSX << gl);
                            if ((A.fuse & 2) && (stat & fixbits) == 0)
                                A.xp1[ia + (size_t)gl * P.n] = __dadd_rn(vpre, __dmul_rn(A.hs2, __ddiv_rn(f, A.mass.cm[ti])));
                        }
                        A.fp[ia + (size_t)gl * P.n] = f;
                    }
                }
            };

            while (true) {
                int c0 = 0;
                if (lane == 0) c0 = atomicAdd(&s_ctr[b], APW);
                c0 = __shfl_sync(0xffffffffu, c0, 0);
                if (c0 >= own_count) break;
                if (edge_tile) chunk(std::true_type(), c0);
                else chunk(std::false_type(), c0);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[b]);
            if (++b == nbuf) { b = 0; u++; }
        }
    }
    if (VIR) { // one partial tensor per warp -> vpart[blockIdx.x * NW + warp][9] (column-major 3x3), summed by k_virial_reduce
               // (no shared memory: the dynamic window already takes the whole per-CTA budget)
        double v6[6] = {vxx, vxy, vxz, vyy, vyz, vzz};
#pragma unroll
        for (int q = 0; q < 6; q++)
            for (int off = 16; off > 0; off >>= 1) v6[q] += __shfl_xor_sync(0xffffffffu, v6[q], off);
        if (lane == 0) {
            double *o = A.vpart + ((size_t)blockIdx.x * NW + warp) * 9;
            o[0] = v6[0]; o[1] = v6[1]; o[2] = v6[2];
            o[3] = v6[1]; o[4] = v6[3]; o[5] = v6[4];
            o[6] = v6[2]; o[7] = v6[4]; o[8] = v6[5];
        }
    }
    // atoms parked outside the cells (out of box, inactive): zero outputs, as the generic path does
    if (blockIdx.x == 0 && A.zero_parked) {
        const int n_in = A.counters[CNT_INCELL];
        for (int i = n_in + threadIdx.x; i < P.n; i += NT) {
            if (PASS == 3) { A.epot[i] = 0.0; if (A.den_too) reinterpret_cast<double *>(A.pos + i)[3] = 0.0; }
            else if (PASS == 1) reinterpret_cast<double *>(A.pos + i)[3] = 0.0;
            else { A.fp[i] = 0.0; A.fp[i + (size_t)P.n] = 0.0; A.fp[i + 2 * (size_t)P.n] = 0.0; }
        }
    }
}

// =====================================================================================
// host side: planning and launch
// =====================================================================================
static const int SMEM_BUDGET = 227 * 1024;

// largest Fortran row index with a non-zero entry in the kind-major host copy kept by the context
static int last_nonzero_row(const std::vector<double> &t, int nkind, int ntab)
{
    for (int i = ntab; i >= 1; i--)
        for (int k = 0; k < nkind; k++)
            if (t[(size_t)(i - 1) * nkind + k] != 0.0) return i;
    return 0;
}

int mdb_tiled_plan(mdb_ctx *c)
{
    TiledState &S = c->tiled;
    S.ok = false;
    if (!c->has_tables || !c->has_nlist || !c->shape_identity) return MDB_OK;
    if (c->mxkvois > 4096 || c->nc <= 0) return MDB_OK;
    const TableSet &t = c->tab;
    // ---- table support -> effective cut-offs (rows >= kz+1 interpolate to exactly 0)
    const double csi = t.csi;
    auto r2_of_row = [&](int kz) { const double r = ((double)(kz + 2) / csi) * ((double)(kz + 2) / csi); return r * r; };
    const int kz1 = last_nonzero_row(c->h_potb, t.nkind, t.ntab);
    const int kz2 = std::max(last_nonzero_row(c->h_fpotr, t.nkind, t.ntab), last_nonzero_row(c->h_fpotb, t.nkind, t.ntab));
    S.r2eff[0] = std::min(t.ru2max, r2_of_row(kz1));
    S.r2eff[1] = std::min(t.ru2max, r2_of_row(kz2));
    // per-atom energy pass: POTR and POTB; it may scan the class pass 2 scans when its tables end no later than pass 2's
    const int kz3 = std::max(last_nonzero_row(c->h_potr, t.nkind, t.ntab), kz1);
    S.r2eff_epot = std::min(t.ru2max, r2_of_row(kz3));
    S.epot_in_class1 = S.r2eff_epot <= S.r2eff[1];
    const int kru = std::min(t.ntab + 1, (int)(std::sqrt(std::sqrt(t.ru2max)) * csi) + 1);
    S.khi[0] = std::min(kru, kz1 + 2);
    S.khi[1] = std::min(kru, kz2 + 2);

    // ---- class margins (see header): class 1 (what pass 2 scans) keeps a quarter of the list skin (NB_RM - RU); class 0
    //      (pass 1) a tighter 0.12 of it, so that a neighbour shell just outside the density range does not leak into the
    //      class and push the scan to another 4-entry group per lane; when atoms have moved more than half of that,
    //      pass 1 falls back to classes 0+1, and only beyond half the wider margin to the full list
    double rmmax = 0.0;
    for (int i = 0; i < c->ng * c->ng; i++) rmmax = std::max(rmmax, c->nb_rm[i]);
    const double ru = std::sqrt(t.ru2max);
    S.margin = std::max(0.0, 0.25 * (rmmax - ru));
    S.margin0 = std::max(0.0, 0.12 * (rmmax - ru));
    // ---- tile geometry: the widest tile (fewest halo atoms staged per owned atom) whose two pipeline stages
    //      leave room for a table window that reaches down to 0.6 of the support edge's row (r ~ 0.36 r_eff;
    //      closer pairs read the tables from global memory)
    const double rho_cell = (double)c->n / (double)c->nc;
    const bool mt = c->ng > 1;
    const int need = std::max(S.khi[0], S.khi[1]) + 1;
    const int minrows = std::min(need, std::max(512, (int)(0.4 * need)));
    const int nbuf = (S.stages_opt == 3) ? 3 : 2;
    auto fit_tiles = [&](const int G) -> int { // tile width for G lanes per atom, 0: nothing fits
        for (int ntx = 1; ntx <= c->ncell[0]; ntx++) {
            const int wt = (c->ncell[0] + ntx - 1) / ntx;
            if (wt > TILE_MAX_W) continue;
            int hcap = ((int)(9.0 * (wt + 2) * rho_cell * 1.15) + 64 + 3) & ~3; // multiple of 4: see the dummy records of the passes
            {   // a halo lies in one box and holds a cell at most once per wrap: never more atoms than the box has (small periodic
                // boxes: the 27 halo cells of a 3 x 3 x 3-cell box ARE the box, 2001 atoms instead of the 2362 of the density estimate)
                const long long wraps = (long long)((wt + 2 + c->ncell[0] - 1) / c->ncell[0]) * ((3 + c->ncell[1] - 1) / c->ncell[1]) *
                                        ((3 + c->ncell[2] - 1) / c->ncell[2]);
                const long long bound = (((long long)c->napb * wraps) + 3) & ~3LL;
                if (bound < hcap) hcap = (int)bound;
            }
            const int ocap = (((int)(wt * rho_cell * 1.25) + 32) + 7) & ~7;
            if (hcap >= 16000) continue; // slots carry a 2-bit class tag while the list is built
            const size_t fixed = TP_HDR_BYTES + nbuf * tp_buf_bytes(hcap, ocap, G, mt);
            if (fixed + tp_tab_bytes(minrows) + 1024 > (size_t)SMEM_BUDGET) continue;
            const int maxrows = (int)((SMEM_BUDGET - fixed - 1024) / 16) - 2;
            S.ntx = ntx; S.hcap = hcap; S.ocap = ocap; S.nbuf = nbuf;
            for (int p = 0; p < 2; p++) {
                S.ktab[p] = std::min(S.khi[p] + 1, maxrows);
                S.kmin[p] = S.khi[p] + 1 - S.ktab[p];
            }
            return wt;
        }
        return 0;
    };
    // lanes per atom: the option, or 4 -- and 8 where a tile owns so few atoms that chunks of 8 atoms (4 lanes each) leave most
    // of the CTA's consumer warps without work (PARREP boxes: one 74-atom cell per tile = 10 chunks for 23 warps; ncu: 44 % of
    // the stall samples were warps waiting at the next stage's barrier; 8 lanes: passes 0.112 / 0.142 -> 0.094 / 0.120 ms)
    int G = S.G_opt ? S.G_opt : 4;
    int best_w = fit_tiles(G);
    if (!S.G_opt && best_w && best_w * rho_cell < 96.0) { // measured: 74 atoms per tile +12 % with 8 lanes, 125 atoms (512 x 16 000-atom boxes) -6 %
        const int w8 = fit_tiles(8);
        if (w8) { G = 8; best_w = w8; }
        else best_w = fit_tiles(4);
    }
    S.G = G;
    if (!best_w) return MDB_OK;
    if (S.threads_opt != 512 && S.threads_opt != 768) S.threads_opt = 768;
    S.threads = S.threads_opt;

    TileParams &P = S.P;
    memset(&P, 0, sizeof(P));
    P.n = c->n; P.nbox = c->nbox; P.ncx = c->ncell[0]; P.ncy = c->ncell[1]; P.ncz = c->ncell[2]; P.nc0 = c->nc0;
    P.ntx = S.ntx; P.nrows = c->nbox * c->ncell[1] * c->ncell[2]; P.ntiles = P.nrows * P.ntx; P.hcap = S.hcap; P.ocap = S.ocap;
    for (int d = 0; d < 3; d++) {
        P.pd[d] = c->box.pd[d]; P.lo[d] = c->box.lo[d]; P.size[d] = c->box.size[d];
        P.cell[d] = c->box.size[d] / (double)c->ncell[d];
        P.fbs[d] = (float)c->box.size[d];
        P.half[d] = c->box.pd[d] ? 0.5 * c->box.size[d] : 1.0e300;
    }
    P.ng = c->ng; P.mxkvois = c->mxkvois;
    const int rows_per_lane = (c->mxkvois + G - 1) / G;
    P.nrow4 = (rows_per_lane + 3) / 4;
    P.npad = (size_t)c->n;
    {
        const double re0 = std::sqrt(S.r2eff[0]), re1 = std::sqrt(S.r2eff[1]);
        const double rc0 = re0 + S.margin0, rc1 = std::max(re1 + S.margin, rc0);
        S.rc2f[0] = (float)(rc0 * rc0);
        S.rc2f[1] = (float)(rc1 * rc1);
        // a class holds for a pass while 2*d_max <= 0.98*(class radius - the pass's range)
        S.safe_d2[0] = (float)(0.49 * 0.49 * S.margin0 * S.margin0);                  // pass 1 on class 0
        S.safe_d2[1] = (float)(0.49 * 0.49 * (rc1 - re1) * (rc1 - re1));              // pass 2 on classes 0+1
        S.safe_d2[2] = (float)(0.49 * 0.49 * (rc1 - re0) * (rc1 - re0));              // pass 1 on classes 0+1
    }
    // ---- storage: slot list, raw list, class counts, tile descriptors
    auto ensure = [&](void **ptr, size_t &have, size_t want) -> bool {
        if (have >= want) return true;
        if (*ptr) cudaFree(*ptr);
        *ptr = nullptr; have = 0;
        if (cudaMalloc(ptr, want) != cudaSuccess) { cudaGetLastError(); return false; }
        have = want;
        return true;
    };
    const size_t nbl_bytes = (size_t)P.nrow4 * P.npad * G * 4 * sizeof(unsigned short);
    if (!ensure((void **)&S.nbl, S.nbl_elems, nbl_bytes)) return MDB_OK;
    if (!ensure((void **)&S.ncls, S.ncls_bytes, 2 * P.npad * sizeof(unsigned short) + 32)) return MDB_OK; // +32: aligned TMA windows
    if (!ensure((void **)&c->pos_snap, c->pos_snap_bytes, sizeof(double4) * (size_t)c->n)) return MDB_OK;
    if (!ensure((void **)&S.desc, S.desc_bytes, (size_t)P.ntiles * sizeof(TileDesc))) return MDB_OK;
    if (!ensure((void **)&c->dsr, c->dsr_bytes, 3 * (size_t)c->n * sizeof(float))) return MDB_OK;
    cudaMemsetAsync(c->dsr, 0, 3 * (size_t)c->n * sizeof(float), c->stream);
    if (!ensure((void **)&c->dmax_blk, c->dmax_blk_bytes, sizeof(float) * ((size_t)c->n / 256 + 2))) return MDB_OK;
    if (!ensure((void **)&c->tile_d2, c->tile_d2_bytes, sizeof(float) * (size_t)P.ntiles)) return MDB_OK;
    cudaMemsetAsync(c->dmax_blk, 0, c->dmax_blk_bytes, c->stream);
    c->tile_guard_fresh = false;

    // ---- launch configuration: one persistent CTA per SM
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->dev);
    S.grid = std::min(P.ntiles, nsm);
    // list kernel: one warp per owned cell; per-atom shared-memory list of lcap rows (expected neighbours + 45 %)
    {
        const double dens = (double)c->n / ((double)c->nbox * c->box.size[0] * c->box.size[1] * c->box.size[2]);
        const int expect = (int)(4.18879 * rmmax * rmmax * rmmax * dens);
        S.wmax = best_w;
        // two CTAs of the list kernel share an SM when 2 x (dynamic + ~3.7 KB static + 1 KB reserved) <= 228 KB; when a
        // slightly shorter column (never below expected + 25 % + 16) gets there, take it: it doubles the resident warps
        const size_t two_cta = (233472 / 2) - 1024 - 3712 - 512;
        const size_t halo_b = sizeof(float4) * 2 * (size_t)((S.hcap + 2) / 2);
        auto size_for = [&](int pairs, int &lcap, size_t &bytes) -> bool { // false: does not fit at all
            lcap = std::min(c->mxkvois, ((int)(1.45 * expect) + 24 + 7) & ~7);
            bytes = halo_b + sizeof(unsigned short) * 32 * (size_t)lcap * pairs;
            while (bytes > (size_t)SMEM_BUDGET - 4096 && lcap > 32) {
                lcap -= 8;
                bytes = halo_b + sizeof(unsigned short) * 32 * (size_t)lcap * pairs;
            }
            if (bytes > (size_t)SMEM_BUDGET - 4096) return false;
            if (bytes > two_cta && halo_b < two_cta) {
                const int fit = (int)((two_cta - halo_b) / (sizeof(unsigned short) * 32 * (size_t)pairs)) & ~7;
                if (fit >= (((int)(1.25 * expect) + 16 + 7) & ~7)) {
                    lcap = std::min(lcap, fit);
                    bytes = halo_b + sizeof(unsigned short) * 32 * (size_t)lcap * pairs;
                }
            }
            return true;
        };
        // warp pairs of a list CTA: one per cell of the widest tile, times the 32-atom blocks of a cell when the tile is narrow
        // (k_tile_nlist deals a cell's blocks to the pairs that share it).  A pair costs a column block of lcap rows: extra pairs
        // are taken only while two CTAs still share an SM with columns of the full expected length + 45 %.
        const int lcap_full = std::min(c->mxkvois, ((int)(1.45 * expect) + 24 + 7) & ~7);
        int ppc = std::max(1, std::min(TILE_MAX_W / best_w, (int)(1.3 * rho_cell + 31.0) / 32));
        for (;; ppc--) {
            if (!size_for(best_w * ppc, S.lcap, S.smem_list)) {
                if (ppc > 1) continue;
                return MDB_OK;
            }
            if (ppc == 1 || (S.smem_list <= two_cta && S.lcap >= std::min(lcap_full, ((int)(1.25 * expect) + 16 + 7) & ~7))) break;
        }
        S.lpairs = best_w * ppc;
    }
    for (int p = 0; p < 2; p++) S.smem_pass[p] = TP_HDR_BYTES + tp_tab_bytes(S.ktab[p]) + S.nbuf * tp_buf_bytes(S.hcap, S.ocap, G, mt);
    S.bank_order = S.bank_order_opt == 1 || (S.bank_order_opt < 0 && std::min(c->ncell[0], std::min(c->ncell[1], c->ncell[2])) >= 12);
    S.ok = true;
    return MDB_OK;
}

void mdb_tiled_free(mdb_ctx *c)
{
    TiledState &S = c->tiled;
    if (S.nbl) cudaFree(S.nbl);
    if (S.ncls) cudaFree(S.ncls);
    if (S.desc) cudaFree(S.desc);
    S.nbl = nullptr; S.ncls = nullptr; S.desc = nullptr;
    S.nbl_elems = S.ncls_bytes = S.desc_bytes = 0;
    S.ok = false; S.dirty = true; S.active = false;
}

template <int G, bool MT>
static int launch_list(mdb_ctx *c)
{
    TiledState &S = c->tiled;
    TileListArgs A;
    A.pos = c->pos; A.ityp = c->ityp; A.naac = c->naac;
    A.kvois = c->kvois; A.nbl = S.nbl; A.ncls = S.ncls; A.counters = c->counters;
    A.desc = (const TileDesc *)S.desc;
    for (int i = 0; i < MDB_MXGROUP * MDB_MXGROUP; i++) A.rm2[i] = (i < c->ng * c->ng) ? c->rm2f[i] : 0.f;
    A.rc2[0] = S.rc2f[0]; A.rc2[1] = S.rc2f[1];
    A.lcap = S.lcap;
    int tile_lo = 0, tile_hi = S.P.ntiles;
    if (c->dd_on) { // owned z-layers of cells: the descriptors are cheap and built for every tile
        const int tiles_per_layer = c->ncell[1] * S.ntx;
        tile_lo = (int)(((long long)c->dd_rank * c->ncell[2]) / c->dd_n) * tiles_per_layer;
        tile_hi = (int)(((long long)(c->dd_rank + 1) * c->ncell[2]) / c->dd_n) * tiles_per_layer;
    }
    A.tile_lo = tile_lo;
    auto kern = k_tile_nlist<G, MT>;
    CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.smem_list));
    ProfScope ps(c, MDB_K_NLIST, S.bank_order ? 3 : 2);
    // descriptors: of this rank's tiles only in a decomposed run (cell counts of other slabs are not kept current there)
    if (tile_hi > tile_lo) {
        k_tile_desc<<<tile_hi - tile_lo, NL_THREADS, 0, c->stream>>>(S.P, c->nac, c->ia1th, (TileDesc *)S.desc, c->counters, tile_lo);
        kern<<<tile_hi - tile_lo, 64 * S.lpairs, S.smem_list, c->stream>>>(S.P, A);
        if (S.bank_order) {
            // the atoms of this rank's tiles: its owned range once the decomposition knows it, every atom (filtered by tile) before
            const bool ranged = c->dd_on && c->dd_built;
            const int a0 = ranged ? own_a0(c) : 0, a1 = ranged ? own_a1(c) : c->n;
            if (a1 > a0)
                k_tile_deal<G><<<cdiv(a1 - a0, DEAL_T), DEAL_T, 0, c->stream>>>(S.P, (const TileDesc *)S.desc, a0, a1, tile_lo, tile_hi, c->ic, S.nbl,
                                                                               S.ncls, c->kvois);
        }
    }
    CUDA_TRY(c, cudaGetLastError());
    // the reference-format KVOIS/INDI pair is rebuilt on demand from the positions of this moment
    if (!c->dd_built)
        CUDA_TRY(c, cudaMemcpyAsync(c->pos_snap, c->pos, sizeof(double4) * (size_t)c->n, cudaMemcpyDeviceToDevice, c->stream));
    c->indi_stale = true;
    return MDB_OK;
}

int mdb_tiled_nlist(mdb_ctx *c)
{
    const bool mt = c->ng > 1;
    switch (c->tiled.G) {
    case 2: return mt ? launch_list<2, true>(c) : launch_list<2, false>(c);
    case 4: return mt ? launch_list<4, true>(c) : launch_list<4, false>(c);
    case 8: return mt ? launch_list<8, true>(c) : launch_list<8, false>(c);
    }
    return mdb_fail(c, MDB_ERR_ARG, "tiled path: unsupported lane-group size %d", c->tiled.G);
}

// per-tile displacement bounds for the tiles [lo, hi) (the passes of this step use them until mdb_tile_guard_clear)
int mdb_tile_guard_launch(mdb_ctx *c, int lo, int hi)
{
    TiledState &S = c->tiled;
    if (!S.active || !c->dmax_blk || !c->tile_d2 || hi <= lo) return MDB_OK;
    ProfScope ps(c, MDB_K_OTHER);
    k_tile_d2<<<cdiv(hi - lo, 8), 256, 0, c->stream>>>((const TileDesc *)S.desc, lo, hi - lo, c->dmax_blk, own_a0(c), own_a1(c), c->counters,
                                                     c->tile_d2);
    CUDA_TRY(c, cudaGetLastError());
    c->tile_guard_fresh = true;
    return MDB_OK;
}

template <int PASS, int G, bool MT, bool FUSE, int NT, bool VIR = false>
static int launch_pass(mdb_ctx *c, int fuse, double hs2)
{
    TiledState &S = c->tiled;
    const TableSet &t = c->tab;
    TilePassArgs A;
    A.pos = c->pos; A.ityp = c->ityp; A.statu = c->statu; A.kvois = c->kvois; A.ncls = S.ncls;
    A.nbl = S.nbl; A.fp = c->fp; A.counters = c->counters; A.desc = (const TileDesc *)S.desc;
    A.g_potb = t.potb; A.g_fpotr = t.fpotr; A.g_fpotb = t.fpotb; A.g_dfembd = t.dfembd;
    A.g_potr = t.potr; A.g_fembd = t.fembd; A.epot = c->epot;
    A.vpart = nullptr;
    if (VIR) {
        const int npart = S.grid * (NT / 32) + 2; // one partial per warp, + the reduced tensor
        if (c->vpart_n < npart) {
            if (c->vpart) cudaFree(c->vpart);
            c->vpart = nullptr;
            CUDA_TRY(c, cudaMalloc(&c->vpart, sizeof(double) * 9 * (size_t)npart));
            c->vpart_n = npart;
        }
        A.vpart = c->vpart;
    }
    A.den_too = (PASS == 3 && fuse == -1) ? 1 : 0;
    constexpr int PI = (PASS == 3) ? 1 : PASS - 1; // the energy pass shares the plan (window, classes, shared memory) of pass 2
    A.ntab = t.ntab; A.nembd = t.nembd; A.pot_type = t.pot_type; A.csi = t.csi; A.rhod = t.rhod;
    A.r2eff = (PASS == 3) ? S.r2eff_epot : S.r2eff[PI]; A.kmin = S.kmin[PI]; A.ktab = S.ktab[PI];
    A.kind0 = t.kpair[0];
    if (PASS == 1) { A.safe_a = S.safe_d2[0]; A.row_a = 0; A.safe_b = S.safe_d2[2]; A.row_b = 1; }
    else { A.safe_a = S.safe_d2[1]; A.row_a = 1; A.safe_b = -1.0f; A.row_b = 1; }
    if (!S.use_classes || (PASS == 3 && !S.epot_in_class1)) { A.safe_a = -1.0f; A.safe_b = -1.0f; }
    A.tile_d2 = c->tile_guard_fresh ? c->tile_d2 : nullptr;
    A.fuse = fuse; A.hs2 = hs2; A.xp1 = c->xp1; A.epc = c->epc; A.mass = c->mass;
    A.tile_lo = c->dd_on ? c->dd_info[14] : 0;
    A.tile_hi = c->dd_on ? c->dd_info[15] : S.P.ntiles;
    A.tile_lo2 = A.tile_hi2 = 0;
    if (c->tile_sel[0] >= 0) { // a sub-range launch of a decomposed step (interior tiles / the two boundary layers)
        A.tile_lo = c->tile_sel[0]; A.tile_hi = c->tile_sel[1]; A.tile_lo2 = c->tile_sel[2]; A.tile_hi2 = c->tile_sel[3];
    }
    A.zero_parked = c->dd_on ? 0 : 1;
    A.nbuf = S.nbuf;
    A.skip = c->skip_flag;
    if (PASS == 3) A.fuse = 0;
    if (!c->epc.on) A.fuse &= ~1;
    for (int i = 0; i < MDB_MXGROUP * MDB_MXGROUP; i++) A.kpair[i] = t.kpair[i];
    for (int i = 0; i < MDB_MXGROUP; i++) A.kembd[i] = t.kembd[i];
    auto kern = k_tile_pass<PASS, G, MT, FUSE, NT, VIR>;
    CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.smem_pass[PI]));
    ProfScope ps(c, PASS == 1 ? MDB_K_PASS1 : (PASS == 2 ? MDB_K_PASS2 : MDB_K_EPOT));
    const int ntl = (A.tile_hi - A.tile_lo) + (A.tile_hi2 - A.tile_lo2);
    if (ntl <= 0) return MDB_OK;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(std::min(S.grid, ntl)); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = S.smem_pass[PI]; cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = (c->opt_pdl && !c->dd_on) ? 1 : 0; // (decomposed steps order their kernels with events)
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_TRY(c, cudaLaunchKernelEx(&cfg, kern, S.P, A));
    }
    CUDA_TRY(c, cudaGetLastError());
    return MDB_OK;
}

template <int G, bool MT, int NT>
static int launch_force(mdb_ctx *c, unsigned flags, int fuse, double hs2)
{
    int rc = MDB_OK;
    const bool need_den = (flags & (MDB_FORCE | MDB_DEN | MDB_VIRIAL)) && !(flags & MDB_NOPASS1);
    // energies and forces together (every quench iteration): the energy pass forms the same density sum as pass 1, so it
    // also stores dF/drho and pass 1 is not launched
    const bool den_in_epot = need_den && (flags & MDB_EPOT);
    if (need_den && !den_in_epot) rc = launch_pass<1, G, MT, false, NT>(c, 0, 0.0);
    if (rc < 0) return rc;
    if (den_in_epot) rc = launch_pass<3, G, MT, false, NT>(c, -1, 0.0);
    if (rc < 0) return rc;
    if (flags & MDB_VIRIAL) rc = launch_pass<2, G, MT, false, NT, true>(c, 0, 0.0);   // forces + virial (CALPTENSOR)
    else if (flags & MDB_FORCE) rc = fuse ? launch_pass<2, G, MT, true, NT>(c, fuse, hs2) : launch_pass<2, G, MT, false, NT>(c, 0, 0.0);
    if (rc < 0) return rc;
    if ((flags & MDB_EPOT) && !den_in_epot) rc = launch_pass<3, G, MT, false, NT>(c, 0, 0.0);
    return rc;
}

template <int G>
static int launch_force_g(mdb_ctx *c, unsigned flags, int fuse, double hs2)
{
    const bool mt = c->ng > 1;
    if (c->tiled.threads == 512)
        return mt ? launch_force<G, true, 512>(c, flags, fuse, hs2) : launch_force<G, false, 512>(c, flags, fuse, hs2);
    return mt ? launch_force<G, true, 768>(c, flags, fuse, hs2) : launch_force<G, false, 768>(c, flags, fuse, hs2);
}

// fuse: bit 0 = EPC friction, bit 1 = corrector half-kick with hs2 = H/2, applied in the epilogue of pass 2
int mdb_force_tiled(mdb_ctx *c, unsigned flags, int fuse, double hs2)
{
    switch (c->tiled.G) {
    case 2: return launch_force_g<2>(c, flags, fuse, hs2);
    case 4: return launch_force_g<4>(c, flags, fuse, hs2);
    case 8: return launch_force_g<8>(c, flags, fuse, hs2);
    }
    return mdb_fail(c, MDB_ERR_ARG, "tiled path: unsupported lane-group size %d", c->tiled.G);
}

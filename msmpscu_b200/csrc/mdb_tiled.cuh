// mdb_tiled.cuh -- geometry of the TILED path (shared by the list-build and the force kernels).
//
// A tile is a run of consecutive cells along x inside one cell row (box, cz, cy).  Cells are
// ordered x-fastest (reference order, CommonGPU/MD_NeighborsList_GPU.F90:1543-1548), so
//   * the tile's OWNED atoms are one contiguous range of the cell-sorted arrays, and
//   * its HALO -- the (Wt+2) x 3 x 3 block of cells around it, periodic images included -- is at
//     most 27 contiguous runs of the position array.
// The halo is staged once per tile into shared memory; the neighbour list of the tiled path stores
// 16-bit SLOTS into that staged halo instead of 32-bit global atom ids, so per-pair gathers are
// shared-memory reads and the index stream is half the bytes.  Slot numbering is a pure function of
// (tile, NAC, IA1th), all frozen between rebuilds, so the list-build kernel and the force kernels
// agree on it without storing any per-tile descriptor.
#pragma once
#include "mdb_internal.cuh"

#define TILE_MAX_W      12                         // cells per tile along x
#define TILE_MAX_HC     ((TILE_MAX_W + 2) * 9)     // halo cells per tile
#define TILE_MAX_RUN    32                         // contiguous runs of halo cells per tile (one per producer lane)
#define NBL_UNROLL      4                          // list entries per 8-byte index load

struct TileGeom { // per-tile values, computed by every thread from blockIdx-independent tile id
    int box, cy, cz, cx0, wt;      // wt = cells in the tile
    int nhx, nhc;                  // halo row length wt+2, halo cells nhx*9
};

__device__ __forceinline__ TileGeom tile_geom(const TileParams &P, int tile)
{
    TileGeom g;
    const int row = tile / P.ntx, tx = tile - row * P.ntx;
    const int nyz = P.ncy * P.ncz;
    g.box = row / nyz;
    const int rem = row - g.box * nyz;
    g.cz = rem / P.ncy;
    g.cy = rem - g.cz * P.ncy;
    g.cx0 = (int)(((long long)tx * P.ncx) / P.ntx);
    const int cx1 = (int)(((long long)(tx + 1) * P.ncx) / P.ntx);
    g.wt = cx1 - g.cx0;
    g.nhx = g.wt + 2;
    g.nhc = g.nhx * 9;
    return g;
}

// halo cell hc = (hz*3 + hy)*nhx + hx  ->  wrapped global cell id (or -1 if outside a non-periodic
// box) and the image shift code sh[d] in {-1,0,+1} (candidate position += sh*BOXSIZE)
__device__ __forceinline__ int halo_cell(const TileParams &P, const TileGeom &g, int hc, int sh[3], int u[3])
{
    const int hx = hc % g.nhx, hyz = hc / g.nhx, hy = hyz % 3, hz = hyz / 3;
    u[0] = g.cx0 - 1 + hx; u[1] = g.cy - 1 + hy; u[2] = g.cz - 1 + hz; // unwrapped cell coordinates
    const int nc[3] = {P.ncx, P.ncy, P.ncz};
    int w[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        w[d] = u[d]; sh[d] = 0;
        if (u[d] < 0) { if (!P.pd[d]) return -1; w[d] = u[d] + nc[d]; sh[d] = -1; }
        else if (u[d] >= nc[d]) { if (!P.pd[d]) return -1; w[d] = u[d] - nc[d]; sh[d] = +1; }
    }
    return w[0] + P.ncx * (w[1] + P.ncy * w[2]) + g.box * P.nc0;
}

// index of list entry k of atom a (global cell-order index) for a G-lane group:
// k = G*m + gl ; entries are stored [m/4][atom][gl][m%4] so that one lane loads 4 consecutive m
// with a single 8-byte read and a warp reads 256 contiguous bytes.
template <int G>
__device__ __forceinline__ size_t nbl_index(const TileParams &P, size_t a, int k)
{
    const int gl = k % G, m = k / G;
    return ((((size_t)(m >> 2) * P.npad + a) * G + gl) << 2) + (m & 3);
}

// Micro-benchmark: which lanes of a warp share one shared-memory wavefront of an LDS.128?
// Every lane reads 16 bytes at a random 128-byte line; the 16-byte bank group inside the line is a fixed function of the
// lane (pattern).  A pattern is conflict-free exactly when the lanes the hardware serves together hit distinct groups.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds128_patterns lds128_patterns.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define LINES 1024   // 128 KB
#define ITERS 2048

__constant__ int c_pat[32];

__global__ void __launch_bounds__(512, 1) k(int randomize, const int *idx, double *out, long long *cyc)
{
    extern __shared__ double2 s[];
    for (int i = threadIdx.x; i < LINES * 8; i += blockDim.x) s[i] = make_double2(i, 0.5 * i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned g = c_pat[lane];
    unsigned r = idx[blockIdx.x * blockDim.x + threadIdx.x];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        unsigned r0 = r % LINES, r1 = (r * 7u + 13u) % LINES, r2 = (r * 31u + 5u) % LINES, r3 = (r * 127u + 1u) % LINES;
        unsigned g0 = g, g1 = g, g2 = g, g3 = g;
        if (randomize) { g0 = (r >> 3) & 7; g1 = (r >> 7) & 7; g2 = (r >> 11) & 7; g3 = (r >> 15) & 7; }
        const double2 v0 = s[r0 * 8 + g0], v1 = s[r1 * 8 + g1], v2 = s[r2 * 8 + g2], v3 = s[r3 * 8 + g3];
        a0 += v0.x; a1 += v1.y; a2 += v2.x; a3 += v3.y;
        r = r * 1664525u + 1013904223u + (unsigned)(a0 > 1e300);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    const int nb = 148, nt = 512;
    std::vector<int> hi(nb * nt);
    for (auto &x : hi) x = rand();
    int *di; double *o; long long *c;
    cudaMalloc(&di, sizeof(int) * nb * nt); cudaMalloc(&o, sizeof(double) * nb * nt); cudaMalloc(&c, sizeof(long long) * nb);
    cudaMemcpy(di, hi.data(), sizeof(int) * nb * nt, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, LINES * 128);
    struct Pat { const char *name; int (*f)(int); int rnd; };
    Pat pats[] = {
        {"random group", [](int l) { return 0; }, 1},
        {"lane % 8            (distinct within lanes 8q..8q+7)", [](int l) { return l % 8; }, 0},
        {"lane / 4            (distinct within lanes l, l+4, l+8, ..)", [](int l) { return l / 4; }, 0},
        {"l%4 + 4*((l/16)&1)  (distinct within {0-3,16-19})", [](int l) { return l % 4 + 4 * ((l / 16) & 1); }, 0},
        {"(l/2) % 8           (lanes 2m,2m+1 share a group)", [](int l) { return (l / 2) % 8; }, 0},
        {"(l%2)*4 + l/8       (distinct within {0,1,8,9,16,17,24,25})", [](int l) { return (l % 2) * 4 + l / 8; }, 0},
        {"l%4                 (4 groups only, distinct within 4 consecutive)", [](int l) { return l % 4; }, 0},
        {"(l%4)*2             (even groups only)", [](int l) { return (l % 4) * 2; }, 0},
        {"all lanes group 0   (32 lines, one group)", [](int l) { return 0; }, 0},
        {"l/8                 (8 consecutive lanes share a group)", [](int l) { return l / 8; }, 0},
        {"l/16                (16 consecutive lanes share a group)", [](int l) { return l / 16; }, 0},
    };
    for (auto &p : pats) {
        int h[32];
        for (int l = 0; l < 32; l++) h[l] = p.f(l);
        cudaMemcpyToSymbol(c_pat, h, sizeof(h));
        k<<<nb, nt, LINES * 128>>>(p.rnd, di, o, c);
        cudaDeviceSynchronize();
        k<<<nb, nt, LINES * 128>>>(p.rnd, di, o, c);
        cudaError_t e = cudaDeviceSynchronize();
        long long hc[148]; cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
        double warp_reads = (double)ITERS * 4 * (nt / 32);
        printf("%-70s : %6.2f cycles per warp-level LDS.128  [%s]\n", p.name, hc[0] / warp_reads, cudaGetErrorString(e));
    }
    return 0;
}

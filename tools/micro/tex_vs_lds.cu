// Micro-benchmark: random 16-byte table reads per lane, (a) from shared memory (LDS.128), (b) through the texture
// path (tex1Dfetch<int4> on linear memory, L1TEX-cached), (c) both interleaved.  Prints cycles per warp-level read.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_vs_lds tex_vs_lds.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define ROWS 3072   // 48 KB of 16-byte rows
#define ITERS 2048

__global__ void __launch_bounds__(512, 1) k(int mode, cudaTextureObject_t tex, const double2 *tab, const int *idx, double *out, long long *cyc)
{
    extern __shared__ double2 s_tab[];
    for (int i = threadIdx.x; i < ROWS; i += blockDim.x) s_tab[i] = tab[i];
    __syncthreads();
    unsigned r = idx[blockIdx.x * blockDim.x + threadIdx.x];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        // four independent reads per iteration (like the four entries of an index group)
        unsigned r0 = r % ROWS, r1 = (r * 7u + 13u) % ROWS, r2 = (r * 31u + 5u) % ROWS, r3 = (r * 127u + 1u) % ROWS;
        double2 v0, v1, v2, v3;
        if (mode == 0) { v0 = s_tab[r0]; v1 = s_tab[r1]; v2 = s_tab[r2]; v3 = s_tab[r3]; }
        else if (mode == 1) {
            int4 t;
            t = tex1Dfetch<int4>(tex, r0); v0 = make_double2(__hiloint2double(t.y, t.x), __hiloint2double(t.w, t.z));
            t = tex1Dfetch<int4>(tex, r1); v1 = make_double2(__hiloint2double(t.y, t.x), __hiloint2double(t.w, t.z));
            t = tex1Dfetch<int4>(tex, r2); v2 = make_double2(__hiloint2double(t.y, t.x), __hiloint2double(t.w, t.z));
            t = tex1Dfetch<int4>(tex, r3); v3 = make_double2(__hiloint2double(t.y, t.x), __hiloint2double(t.w, t.z));
        } else if (mode == 2) {
            int4 t;
            v0 = s_tab[r0]; v1 = s_tab[r1];
            t = tex1Dfetch<int4>(tex, r2); v2 = make_double2(__hiloint2double(t.y, t.x), __hiloint2double(t.w, t.z));
            t = tex1Dfetch<int4>(tex, r3); v3 = make_double2(__hiloint2double(t.y, t.x), __hiloint2double(t.w, t.z));
        } else {
            v0 = __ldg(tab + r0); v1 = __ldg(tab + r1); v2 = __ldg(tab + r2); v3 = __ldg(tab + r3);
        }
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
    std::vector<double2> h(ROWS);
    for (int i = 0; i < ROWS; i++) h[i] = make_double2(i, 0.5 * i);
    std::vector<int> hi(nb * nt);
    for (auto &x : hi) x = rand();
    double2 *d; int *di; double *o; long long *c;
    cudaMalloc(&d, sizeof(double2) * ROWS); cudaMalloc(&di, sizeof(int) * nb * nt); cudaMalloc(&o, sizeof(double) * nb * nt); cudaMalloc(&c, sizeof(long long) * nb);
    cudaMemcpy(d, h.data(), sizeof(double2) * ROWS, cudaMemcpyHostToDevice);
    cudaMemcpy(di, hi.data(), sizeof(int) * nb * nt, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = d; rd.res.linear.desc = cudaCreateChannelDesc<int4>();
    rd.res.linear.sizeInBytes = sizeof(double2) * ROWS;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    const char *names[4] = {"LDS.128 (shared)", "tex1Dfetch<int4>", "2 LDS + 2 TEX", "LDG.128 (__ldg)"};
    for (int smem_kb : {48, 160}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
        for (int mode = 0; mode < 4; mode++) {
            k<<<nb, nt, smem_kb * 1024>>>(mode, tex, d, di, o, c);
            cudaDeviceSynchronize();
            k<<<nb, nt, smem_kb * 1024>>>(mode, tex, d, di, o, c);
            cudaError_t e = cudaDeviceSynchronize();
            long long hc[148]; cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
            double warp_reads = (double)ITERS * 4 * (nt / 32);
            printf("smem %3d KB  %-18s : %.2f cycles per warp-level 16-byte gather (per SM)  [%s]\n", smem_kb, names[mode], hc[0] / warp_reads, cudaGetErrorString(e));
        }
    }
    return 0;
}

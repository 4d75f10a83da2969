// Microbenchmarks of primitives used by the densification kernels (dev tool, not shipped).
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <int MODE>
__global__ void __launch_bounds__(1024) k_chain(double* out, float* outf, int iters, double a, double b) {
    // dependent chain per thread: latency test with 1 warp, throughput test with many warps
    double x = a + threadIdx.x; double y = b; float xf = (float)a + threadIdx.x, yf = (float)b;
    long long t0 = clock64();
    if (MODE == 0) { for (int i = 0; i < iters; ++i) { x = x + y; } }                      // DADD dependent
    if (MODE == 1) { for (int i = 0; i < iters; ++i) { x = fma(x, y, y); } }               // DFMA dependent
    if (MODE == 2) { for (int i = 0; i < iters; ++i) { xf = xf + yf; } }                   // FADD dependent
    if (MODE == 3) { double x2 = x + 1, x3 = x + 2, x4 = x + 3;                             // 4 independent DADD chains
        for (int i = 0; i < iters; ++i) { x = x + y; x2 = x2 + y; x3 = x3 + y; x4 = x4 + y; } x += x2 + x3 + x4; }
    if (MODE == 4) { for (int i = 0; i < iters; ++i) { x = x / y; } }                      // DDIV dependent
    if (MODE == 5) { for (int i = 0; i < iters; ++i) { x = __shfl_xor_sync(0xffffffffu, x, 1) + y; } }  // SHFL(double)+DADD
    if (MODE == 6) { for (int i = 0; i < iters; ++i) { xf = __shfl_xor_sync(0xffffffffu, xf, 1) + yf; } } // SHFL(float)+FADD
    if (MODE == 7) { for (int i = 0; i < iters; ++i) { x = (x > y) ? x - y : x + y; } }    // DSETP+select+DADD
    long long t1 = clock64();
    if (x == 12345.678 || xf == 1234.5f) out[0] = x + xf;
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[1] = (double)(t1 - t0); }
}

__global__ void __launch_bounds__(1024) k_blockscan(double* out, int iters) {
    __shared__ double scratch[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double v = threadIdx.x * 0.5, acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double inc = v;
        for (int o = 1; o < 32; o <<= 1) { double n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
        __syncthreads();
        if (lane == 31) scratch[warp] = inc;
        __syncthreads();
        double wv = (lane < nw) ? scratch[lane] : 0.0, winc = wv;
        for (int o = 1; o < 32; o <<= 1) { double n = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += n; }
        acc += __shfl_sync(0xffffffffu, winc - wv, warp) + inc - v;
        v += 1.0;
    }
    long long t1 = clock64();
    if (acc == 1.2345) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0) / iters;
}

__global__ void __launch_bounds__(1024) k_sync(double* out, int iters) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0) / iters;
}

__global__ void __launch_bounds__(1024) k_csync(double* out, int iters) {
    cg::cluster_group c = cg::this_cluster();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) c.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0) / iters;
}

__global__ void k_clockrate(double* out) {
    long long c0 = clock64(); unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    double x = threadIdx.x; for (int i = 0; i < 200000; ++i) x = x * 1.0000001f + 1e-9;
    long long c1 = clock64(); unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (x == 1.23) out[0] = x;
    out[1] = (double)(c1 - c0) / (double)(g1 - g0);   // clock64 ticks per ns
}

int main() {
    double* d; float* f; cudaMalloc(&d, 64); cudaMalloc(&f, 64);
    double h[2];
    k_clockrate<<<1, 1>>>(d); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("clock64 ticks per ns: %.3f\n", h[1]);
    const int iters = 4096;
    const char* names[] = {"DADD dep", "DFMA dep", "FADD dep", "4x DADD indep", "DDIV dep", "SHFL64+DADD", "SHFL32+FADD", "DSETP+SEL+DADD"};
    for (int threads : {32, 1024}) {
        for (int m = 0; m < 8; ++m) {
            void (*fn)(double*, float*, int, double, double) = nullptr;
            switch (m) { case 0: fn = k_chain<0>; break; case 1: fn = k_chain<1>; break; case 2: fn = k_chain<2>; break; case 3: fn = k_chain<3>; break;
                         case 4: fn = k_chain<4>; break; case 5: fn = k_chain<5>; break; case 6: fn = k_chain<6>; break; case 7: fn = k_chain<7>; break; }
            fn<<<1, threads>>>(d, f, iters, 1.0, 1.000001); fn<<<1, threads>>>(d, f, iters, 1.0, 1.000001);
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("threads %4d  %-16s %8.1f ticks / iter\n", threads, names[m], h[1] / iters);
        }
    }
    k_blockscan<<<1, 1024>>>(d, 64); k_blockscan<<<1, 1024>>>(d, 64); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("block_exclusive_scan<double> 1024 thr: %.0f ticks\n", h[1]);
    k_blockscan<<<1, 256>>>(d, 64); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("block_exclusive_scan<double>  256 thr: %.0f ticks\n", h[1]);
    k_sync<<<1, 1024>>>(d, 256); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("__syncthreads 1024 thr: %.0f ticks\n", h[1]);
    for (int cs : {1, 2, 3, 4, 8}) {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs * 16); cfg.blockDim = dim3(1024);
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_csync, d, 256);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("cluster.sync size %d (1024 thr): %.0f ticks  (%s)\n", cs, h[1], cudaGetErrorString(e));
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); printf("SMs %d, clock %d kHz, L2 %d MB\n", p.multiProcessorCount, p.clockRate, p.l2CacheSize >> 20);
    return 0;
}

// Dev microbenchmark: what does a 4-plane max-combine streaming read cost on this B200, with / without the 5 B/px writes?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int THREADS, int QPT>
__global__ void __launch_bounds__(THREADS) k(const float* __restrict__ c, float* __restrict__ w, unsigned* __restrict__ bk, float* out, size_t n_per_plane, size_t plane_stride) {
    // grid.y = view, grid.x = span; each CTA handles THREADS*4*QPT pixels
    const size_t view = blockIdx.y;
    const float* p0 = c + view * 4 * plane_stride;
    const size_t base = (size_t)blockIdx.x * THREADS * 4 * QPT;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const size_t px = base + ((size_t)q * THREADS + threadIdx.x) * 4;
        if (px < n_per_plane) {
            float4 a = __ldcs((const float4*)(p0 + px));
            float4 b = __ldcs((const float4*)(p0 + plane_stride + px));
            float4 d = __ldcs((const float4*)(p0 + 2 * plane_stride + px));
            float4 e = __ldcs((const float4*)(p0 + 3 * plane_stride + px));
            float4 m = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                                   fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
            acc += m.x + m.y + m.z + m.w;
            if (MODE >= 1) *(float4*)(w + view * n_per_plane + px) = m;
            if (MODE >= 2) bk[(view * n_per_plane + px) >> 2] = (a.x > b.x) | ((d.y > e.y) << 8);
        }
    }
    if (acc == 1.2345f) out[0] = acc;
}
int main() {
    const size_t N = 512 * 512, R = 46;
    float *c, *w, *out; unsigned* bk;
    cudaMalloc(&c, R * 4 * N * 4); cudaMalloc(&w, R * N * 4); cudaMalloc(&bk, R * N); cudaMalloc(&out, 16);
    cudaMemset(c, 0, R * 4 * N * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto fn, dim3 grid, int threads) {
        for (int i = 0; i < 5; ++i) fn<<<grid, threads>>>(c, w, bk, out, N, N);
        cudaEventRecord(e0);
        for (int i = 0; i < 50; ++i) fn<<<grid, threads>>>(c, w, bk, out, N, N);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-40s %7.2f us   read %.0f GB/s\n", name, ms / 50 * 1e3, R * 4 * N * 4 / (ms / 50 * 1e-3) / 1e9);
    };
    run("read only, 256thr x 8 quads", k<0, 256, 8>, dim3(N / (256 * 4 * 8), R), 256);
    run("read only, 256thr x 2 quads", k<0, 256, 2>, dim3(N / (256 * 4 * 2), R), 256);
    run("read only, 512thr x 4 quads", k<0, 512, 4>, dim3(N / (512 * 4 * 4), R), 512);
    run("read + w write, 256thr x 8", k<1, 256, 8>, dim3(N / (256 * 4 * 8), R), 256);
    run("read + w + bestk write, 256thr x 8", k<2, 256, 8>, dim3(N / (256 * 4 * 8), R), 256);
    run("read + w + bestk write, 256thr x 2", k<2, 256, 2>, dim3(N / (256 * 4 * 2), R), 256);
    run("read + w + bestk write, 512thr x 4", k<2, 512, 4>, dim3(N / (512 * 4 * 4), R), 512);
    run("read + w + bestk write, 1024thr x 1", k<2, 1024, 1>, dim3(N / (1024 * 4 * 1), R), 1024);
    return 0;
}

// Dev microbenchmark: cost of the prep pass's memory pattern alone -- after the 193 MB streaming read + 60 MB workspace
// write of the stream kernel, re-read the 48 MB weight rows and write them back in place (x * const), 46 views x 512^2.
#include <cstdio>
#include <cuda_runtime.h>
template <int THREADS, int QPT, int PRELOAD>
__global__ void __launch_bounds__(THREADS) scale_inplace(float* __restrict__ w, size_t n, float f) {
    const size_t base = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * THREADS * 4 * QPT;
    float4 v[QPT];
    if (PRELOAD) {
#pragma unroll
        for (int q = 0; q < QPT; ++q) v[q] = __ldcg((const float4*)(w + base + ((size_t)q * THREADS + threadIdx.x) * 4));
#pragma unroll
        for (int q = 0; q < QPT; ++q) { float4 a = v[q]; a.x *= f; a.y *= f; a.z *= f; a.w *= f; *(float4*)(w + base + ((size_t)q * THREADS + threadIdx.x) * 4) = a; }
    } else {
#pragma unroll
        for (int q = 0; q < QPT; ++q) { float4 a = __ldcg((const float4*)(w + base + ((size_t)q * THREADS + threadIdx.x) * 4)); a.x *= f; a.y *= f; a.z *= f; a.w *= f; *(float4*)(w + base + ((size_t)q * THREADS + threadIdx.x) * 4) = a; }
    }
}
template <int THREADS, int QPT>
__global__ void __launch_bounds__(THREADS) stream4(const float* __restrict__ c, float* __restrict__ w, unsigned* __restrict__ bk, size_t n_per_plane, size_t plane_stride) {
    const size_t view = blockIdx.y;
    const float* p0 = c + view * 4 * plane_stride;
    const size_t base = (size_t)blockIdx.x * THREADS * 4 * QPT;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const size_t px = base + ((size_t)q * THREADS + threadIdx.x) * 4;
        float4 a = __ldcs((const float4*)(p0 + px)), b = __ldcs((const float4*)(p0 + plane_stride + px));
        float4 d = __ldcs((const float4*)(p0 + 2 * plane_stride + px)), e = __ldcs((const float4*)(p0 + 3 * plane_stride + px));
        float4 m = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                               fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
        *(float4*)(w + view * n_per_plane + px) = m;
        bk[(view * n_per_plane + px) >> 2] = (a.x > b.x) | ((d.y > e.y) << 8);
    }
}
int main() {
    const size_t N = 512 * 512, R = 46;
    float *c, *w; unsigned* bk;
    cudaMalloc(&c, R * 4 * N * 4); cudaMalloc(&w, R * N * 4); cudaMalloc(&bk, R * N);
    cudaMemset(c, 0, R * 4 * N * 4); cudaMemset(w, 0, R * N * 4);
    cudaEvent_t e[4]; for (auto& x : e) cudaEventCreate(&x);
    auto run = [&](const char* name, auto fn, dim3 grid, int threads) {
        float t_stream = 0, t_scale = 0;
        for (int i = 0; i < 25; ++i) {
            cudaEventRecord(e[0]);
            stream4<256, 8><<<dim3(N / (256 * 4 * 8), R), 256>>>(c, w, bk, N, N);
            cudaEventRecord(e[1]);
            fn<<<grid, threads>>>(w, R * N, 1.0001f);
            cudaEventRecord(e[2]); cudaEventSynchronize(e[2]);
            float a, b; cudaEventElapsedTime(&a, e[0], e[1]); cudaEventElapsedTime(&b, e[1], e[2]);
            if (i >= 5) { t_stream += a; t_scale += b; }
        }
        printf("%-44s stream %6.2f us   in-place rw %6.2f us  (%.0f GB/s r+w)\n", name, t_stream / 20 * 1e3, t_scale / 20 * 1e3, 2 * R * N * 4 / (t_scale / 20 * 1e-3) / 1e9);
    };
    run("256thr x 8 quads, load-use-store", scale_inplace<256, 8, 0>, dim3(N / (256 * 4 * 8), R), 256);
    run("256thr x 8 quads, preload all", scale_inplace<256, 8, 1>, dim3(N / (256 * 4 * 8), R), 256);
    run("256thr x 2 quads, preload", scale_inplace<256, 2, 1>, dim3(N / (256 * 4 * 2), R), 256);
    run("512thr x 4 quads, preload", scale_inplace<512, 4, 1>, dim3(N / (512 * 4 * 4), R), 512);
    run("1024thr x 1 quad", scale_inplace<1024, 1, 1>, dim3(N / (1024 * 4 * 1), R), 1024);
    run("256thr x 1 quad", scale_inplace<256, 1, 1>, dim3(N / (256 * 4 * 1), R), 256);
    return 0;
}

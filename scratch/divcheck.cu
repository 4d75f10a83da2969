// dev tool: div_by (hoisted-reciprocal division, ldp_sample.cu) against __fdiv_rn over many operand pairs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float div_by(float a, float b, float y) {
    const float q0 = __fmul_rn(a, y);
    const float r0 = __fmaf_rn(-b, q0, a);
    const float q1 = __fmaf_rn(r0, y, q0);
    const float r1 = __fmaf_rn(-b, q1, a);
    return __fmaf_rn(r1, y, q1);
}
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// mode 0: a mantissa+exponent random in [2^-90, 2], b in [1, 2^28); mode 1: exhaustive a mantissas for random b; mode 2: b with all-ones / near-power-of-two significands
__global__ void k(int mode, uint32_t seed, unsigned long long* bad, unsigned long long* first) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nthreads = gridDim.x * blockDim.x;
    for (uint32_t it = 0; it < 4096; ++it) {
        const uint32_t h1 = mix(gid * 4096u + it + seed), h2 = mix(h1 ^ 0x9e3779b9u);
        uint32_t bb = (127u + (h2 >> 27) % 28u) << 23 | (h2 & 0x7fffffu);
        if (mode == 2) { const uint32_t sel = (h2 >> 23) & 3u; bb = (bb & 0xff800000u) | (sel == 0 ? 0x7fffffu : sel == 1 ? 0x7ffffeu : sel == 2 ? 0x000001u : 0u); }
        uint32_t ab;
        if (mode == 1) ab = (126u << 23) | ((gid + it * nthreads) & 0x7fffffu);
        else ab = ((37u + (h1 >> 23) % 91u) << 23) | (h1 & 0x7fffffu);
        const float a = __uint_as_float(ab), b = __uint_as_float(bb);
        const float want = __fdiv_rn(a, b), got = div_by(a, b, __frcp_rn(b));
        if (__float_as_uint(want) != __float_as_uint(got)) {
            atomicAdd(bad, 1ull);
            atomicCAS(first, 0ull, ((unsigned long long)ab << 32) | bb);
        }
    }
}
int main() {
    unsigned long long *bad, *first;
    cudaMallocManaged(&bad, 8); cudaMallocManaged(&first, 8);
    for (int mode = 0; mode < 3; ++mode) {
        *bad = 0; *first = 0;
        for (int rep = 0; rep < 4; ++rep) k<<<148 * 16, 256>>>(mode, 12345u + rep * 977u, bad, first);
        cudaDeviceSynchronize();
        printf("mode %d: %llu mismatches of %llu  first a=%08llx b=%08llx\n", mode, *bad, 4ull * 148 * 16 * 256 * 4096, *first >> 32, *first & 0xffffffffull);
    }
    return 0;
}

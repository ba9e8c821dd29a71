// Developer check (not product code): the packed sqrt/div fast path of cpic_push2.cuh against the
// IEEE intrinsics, bit for bit, on random operands in the ranges the push kernel guarantees.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../cabanapic_b200/csrc -o divsqrt_test divsqrt_test.cu
#include <cstdio>
#include "cpic_push2.cuh"
using namespace cpic;

__device__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// random float with sign 0/any, exponent in [elo, ehi], random mantissa
__device__ float rnd(unsigned seed, int elo, int ehi, bool neg) {
    const unsigned h = hash(seed);
    const unsigned e = 127 + elo + hash(seed ^ 0x9e3779b9u) % (unsigned)(ehi - elo + 1);
    unsigned bits = (e << 23) | (h & 0x7fffffu);
    if (neg && (h >> 31)) bits |= 0x80000000u;
    return __uint_as_float(bits);
}
__global__ void k(unsigned long long n, unsigned long long* bad) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long b0 = 0, b1 = 0;
    for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned s = (unsigned)i * 4u;
        const float2 x = make_float2(rnd(s, 0, 59, false), rnd(s + 1, -59, 59, false));            // sqrt arg
        const float2 r = sqrt2_fast(x);
        b0 += (__float_as_uint(r.x) != __float_as_uint(__fsqrt_rn(x.x))) + (__float_as_uint(r.y) != __float_as_uint(__fsqrt_rn(x.y)));
        const float2 den = make_float2(rnd(s + 2, 0, 29, false), rnd(s + 3, -29, 29, false));      // divisor
        float2 num = make_float2(rnd(s + 5, -69, 69, true), rnd(s + 7, -69, 69, true));            // dividend
        if ((i & 1023) == 0) num.x = 0.f;
        if ((i & 1023) == 1) num.y = 1.f;
        const float2 q = div2_fast(num, den);
        b1 += (__float_as_uint(q.x) != __float_as_uint(__fdiv_rn(num.x, den.x))) + (__float_as_uint(q.y) != __float_as_uint(__fdiv_rn(num.y, den.y)));
    }
    if (b0) atomicAdd(bad, b0);
    if (b1) atomicAdd(bad + 1, b1);
}
int main() {
    unsigned long long* bad; cudaMallocManaged(&bad, 16); bad[0] = bad[1] = 0;
    const unsigned long long n = 1ull << 28;
    k<<<148 * 16, 256>>>(n, bad);
    cudaDeviceSynchronize();
    printf("pairs tested %llu: sqrt mismatches %llu, div mismatches %llu  (%s)\n", n, bad[0], bad[1], cudaGetErrorString(cudaGetLastError()));
    return (bad[0] || bad[1]) ? 1 : 0;
}

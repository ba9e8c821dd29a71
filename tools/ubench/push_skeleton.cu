// Micro-benchmark: the MEMORY SKELETON of the reordering push, one transaction type at a time.
// What does each kind of memory traffic of k_push2 cost on a B200 when the arithmetic is taken away?
// A persistent grid walks 64-record tiles (32-byte records, lane = two consecutive records) exactly like
// the kernel; a bit mask switches the transaction types on:
//    1  record loads (2 x LDG.256 per lane)                 2  record stores (2 x STG.256 per lane, out of place)
//    4  home-cell deposit (24 lanes x RED.v4 per tile)      8  movers: 13 % of the particles do 2 streaks x 3 RED.v4
//   16  slot claims (2 returning atomics per tile)             (home row, then a x/y/z neighbour row), a 16-byte
//   32  interpolator gathers (10 x LDG.128, L1 broadcast)      position store and a histogram atomic
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o push_skeleton push_skeleton.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

struct __align__(32) Rec { float v[8]; };
__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int MASK>
__global__ void __launch_bounds__(256, 3) k_skel(const Rec* __restrict__ src, Rec* __restrict__ dst, long long n,
                                                 float* __restrict__ acc, const float4* __restrict__ interp,
                                                 unsigned* __restrict__ cursor, unsigned* __restrict__ hist, int gx, int gy,
                                                 long long ncell, float* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ntiles = n / 64, stride = (long long)gridDim.x * 8;
    float keep = 0.f;
    Rec a, b; a = Rec{}; b = Rec{};
    long long tile = (long long)blockIdx.x * 8 + warp;
    if ((MASK & 1) && tile < ntiles) { a = src[tile * 64 + 2 * lane]; b = src[tile * 64 + 2 * lane + 1]; }
    for (; tile < ntiles; tile += stride) {
        Rec an = Rec{}, bn = Rec{};
        if ((MASK & 1) && tile + stride < ntiles) { an = src[(tile + stride) * 64 + 2 * lane]; bn = src[(tile + stride) * 64 + 2 * lane + 1]; }
        const long long cell = tile % ncell;                       // 64 particles per cell: tile == cell
        unsigned base = 0;
        if (MASK & 16) {
            unsigned b0 = 0, b1 = 0;
            if (lane == 0) { b0 = atomicAdd(cursor + cell, 32u); b1 = atomicAdd(cursor + cell, 32u); }
            base = __shfl_sync(0xffffffffu, b0 + b1, 0);
        }
        if (MASK & 32) {
            const float4* r = interp + cell * 5;
#pragma unroll
            for (int k = 0; k < 5; ++k) { const float4 f = __ldg(r + k); keep += f.x + f.y + f.z + f.w; }
            const float4* r2 = interp + ((cell + (lane == 7 ? gx : 0)) % ncell) * 5;
#pragma unroll
            for (int k = 0; k < 5; ++k) { const float4 f = __ldg(r2 + k); keep += f.x * f.y + f.z * f.w; }
        }
        // a little dependent arithmetic so that the loads are really consumed
        float s = keep;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a.v[k] = a.v[k] * 1.0001f + 0.5f; b.v[k] = b.v[k] * 0.9999f - 0.5f; s += a.v[k] + b.v[k]; }
        if (MASK & 2) {
            const long long o = tile * 64 + ((base & 1u) ? 1 : 0) * 0;   // same place in the other buffer
            dst[o + 2 * lane] = a; dst[o + 2 * lane + 1] = b;
        }
        if ((MASK & 4) && lane < 24) red4(acc + cell * 12 + (lane % 3) * 4, s, s, s, s);
        if (MASK & 8) {
            const unsigned h = hash32((unsigned)(tile * 64 + 2 * lane));
            for (int half = 0; half < 2; ++half) {
                const unsigned hh = half ? (h >> 16) : (h & 0xffffu);
                if ((hh % 100u) < 13u) {
                    const int dir = (hh >> 8) % 6;
                    const long long d = dir == 0 ? 1 : dir == 1 ? -1 : dir == 2 ? gx : dir == 3 ? -gx : dir == 4 ? (long long)gx * gy : -(long long)gx * gy;
                    const long long c2 = (cell + d + ncell) % ncell;
                    red4(acc + cell * 12, s, s, s, s); red4(acc + cell * 12 + 4, s, s, s, s); red4(acc + cell * 12 + 8, s, s, s, s);
                    red4(acc + c2 * 12, s, s, s, s); red4(acc + c2 * 12 + 4, s, s, s, s); red4(acc + c2 * 12 + 8, s, s, s, s);
                    atomicAdd(hist + c2, 1u);
                    if (MASK & 2) *reinterpret_cast<float4*>(&dst[tile * 64 + 2 * lane + half]) = make_float4(s, s, s, s);
                }
            }
        }
        keep = s * 1e-30f;
        a = an; b = bn;
    }
    if (keep == 123.456f) *sink = keep;
}

template <int MASK>
float run(const Rec* s, Rec* d, long long n, float* acc, const float4* ip, unsigned* cur, unsigned* hist, int gx, int gy, long long nc, float* sink) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_skel<MASK><<<148 * 3, 256>>>(s, d, n, acc, ip, cur, hist, gx, gy, nc, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int gx = 258, gy = 258, gz = 66;
    const long long nc = (long long)gx * gy * gz;         // 4.39 M cells: interpolators 351 MB, accumulators 211 MB (> L2)
    const long long n = nc * 64 / 64 * 64;                 // 64 particles per cell
    Rec *a, *b; float *acc, *sink; float4* ip; unsigned *cur, *hist;
    cudaMalloc(&a, n * 32); cudaMalloc(&b, n * 32); cudaMalloc(&acc, nc * 48); cudaMalloc(&ip, nc * 80);
    cudaMalloc(&cur, nc * 4); cudaMalloc(&hist, nc * 4); cudaMalloc(&sink, 4);
    cudaMemset(a, 0, n * 32); cudaMemset(b, 0, n * 32); cudaMemset(acc, 0, nc * 48); cudaMemset(ip, 0, nc * 80);
    cudaMemset(cur, 0, nc * 4); cudaMemset(hist, 0, nc * 4);
    printf("%lld particles, %lld cells\n", n, nc);
#define RUN(M, what) { const float ms = run<M>(a, b, n, acc, ip, cur, hist, gx, gy, nc, sink); \
        printf("mask %2d  %-58s %7.3f ms  %6.1f G particles/s  %7.1f GB/s at 56 B\n", M, what, ms, n / ms / 1e6, 56.0 * n / ms / 1e6); }
    RUN(1, "loads")
    RUN(3, "loads + stores")
    RUN(7, "loads + stores + home deposit")
    RUN(11, "loads + stores + movers")
    RUN(19, "loads + stores + claims")
    RUN(35, "loads + stores + gathers")
    RUN(15, "loads + stores + home deposit + movers")
    RUN(63, "everything")
    RUN(61, "everything but the stores")
    RUN(55, "everything but the movers")
    RUN(12, "home deposit + movers only")
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

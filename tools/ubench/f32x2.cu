// Developer micro-benchmark (not product code): issue/pipe throughput of packed FP32x2 math
// (FADD2/FMUL2/FFMA2, new on sm_100) against scalar FADD/FMUL/FFMA, with and without
// interleaved integer work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float2 a[8];
    int n[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = make_float2(seed + j + threadIdx.x, seed - j); n[j] = j + threadIdx.x; }
    const float2 c = make_float2(1.0001f, 0.9999f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) { a[j].x = a[j].x * c.x + c.y; a[j].y = a[j].y * c.x + c.y; }        // 2 scalar FFMA
            if (MODE == 1) { a[j] = __ffma2_rn(a[j], c, c); }                                    // 1 FFMA2
            if (MODE == 2) { a[j].x = a[j].x + c.x; a[j].y = a[j].y + c.y; }                     // 2 FADD
            if (MODE == 3) { a[j] = __fadd2_rn(a[j], c); }                                       // 1 FADD2
            if (MODE == 4) { a[j] = __ffma2_rn(a[j], c, c); n[j] = (n[j] ^ i) + j; }             // FFMA2 + 2 int ops
            if (MODE == 5) { a[j].x = a[j].x * c.x + c.y; a[j].y = a[j].y * c.x + c.y; n[j] = (n[j] ^ i) + j; }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j].x + a[j].y + n[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* d) {
    const int iters = 4096, blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, iters, 1.f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 1.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop_pairs = (double)blocks * 256 * iters * 8;   // pair-operations
    printf("%-28s %8.3f ms  %8.2f G pair-ops/s  (%.2f pair-ops/clk/SM at 1.965 GHz)\n", name, ms,
           flop_pairs / ms / 1e6, flop_pairs / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("2x scalar FFMA", d);
    run<1>("1x FFMA2", d);
    run<2>("2x scalar FADD", d);
    run<3>("1x FADD2", d);
    run<4>("FFMA2 + 2 int", d);
    run<5>("2x FFMA + 2 int", d);
    return 0;
}

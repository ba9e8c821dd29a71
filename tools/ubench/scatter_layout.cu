// Micro-benchmark behind the particle-record layout decision: how expensive is a permuting
// copy of 32-byte particles when ~13 % of them land far from their neighbours (the cell-order maintenance
// pattern: movers go to the x / y / z neighbour cell's segment), as struct-of-arrays (eight scattered 4-byte
// writes per moved particle, each a PARTIAL 32-byte sector) versus array-of-structs (one full sector)?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_layout scatter_layout.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
// an involution: element i swaps place with i ^ D when the pair's hash says so (13 % of all elements move)
__device__ __forceinline__ unsigned dest_of(unsigned i, unsigned pct) {
    const unsigned Ds[3] = {64u, 16384u, 4194304u};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned lo = i & ~Ds[k];
        const unsigned h = hash32(lo * 3u + k);
        // pick at most one D per element: D_k applies when hash of the *pair* selects k and the percentage
        if ((h % 3u) == (unsigned)k && (hash32(lo ^ 0x9e3779b9u) % 100u) < pct) {
            // the partner must take the same decision: lo is common to both, but an element could also be
            // selected by another k through a different lo; keep the first k that fires for BOTH ends
            bool clean = true;
            for (int j = 0; j < k; ++j) {
                const unsigned a = i & ~Ds[j], b = (i ^ Ds[k]) & ~Ds[j];
                if (((hash32(a * 3u + j) % 3u) == (unsigned)j && (hash32(a ^ 0x9e3779b9u) % 100u) < pct) ||
                    ((hash32(b * 3u + j) % 3u) == (unsigned)j && (hash32(b ^ 0x9e3779b9u) % 100u) < pct)) clean = false;
            }
            if (clean) return i ^ Ds[k];
        }
    }
    return i;
}
struct SoA { float* m[8]; };
__global__ void k_soa(SoA s, SoA d, unsigned n, unsigned pct) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned j = pct ? dest_of(i, pct) : i;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = s.m[k][i];
#pragma unroll
    for (int k = 0; k < 8; ++k) d.m[k][j] = v[k];
}
__global__ void k_aos(const float4* __restrict__ s, float4* __restrict__ d, unsigned n, unsigned pct) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned j = pct ? dest_of(i, pct) : i;
    const float4 a = s[2 * (size_t)i], b = s[2 * (size_t)i + 1];
    d[2 * (size_t)j] = a; d[2 * (size_t)j + 1] = b;
}
__global__ void k_aos256(const float* __restrict__ s, float* __restrict__ d, unsigned n, unsigned pct) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned j = pct ? dest_of(i, pct) : i;
    float v0, v1, v2, v3, v4, v5, v6, v7;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3), "=f"(v4), "=f"(v5), "=f"(v6), "=f"(v7) : "l"(s + 8 * (size_t)i));
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(d + 8 * (size_t)j), "f"(v0), "f"(v1), "f"(v2), "f"(v3), "f"(v4), "f"(v5), "f"(v6), "f"(v7) : "memory");
}
__global__ void k_check(unsigned n, unsigned pct, unsigned long long* out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned j = dest_of(i, pct);
    if (dest_of(j, pct) != i) atomicAdd(out, 1ull);
    if (j != i) atomicAdd(out + 1, 1ull);
}
int main() {
    const unsigned n = 1u << 27;
    float *a, *b;
    cudaMalloc(&a, (size_t)n * 32); cudaMalloc(&b, (size_t)n * 32);
    cudaMemset(a, 0, (size_t)n * 32); cudaMemset(b, 0, (size_t)n * 32);
    SoA s, d;
    for (int k = 0; k < 8; ++k) { s.m[k] = a + (size_t)k * n; d.m[k] = b + (size_t)k * n; }
    unsigned long long* chk; cudaMalloc(&chk, 16); cudaMemset(chk, 0, 16);
    k_check<<<n / 256, 256>>>(n, 13, chk);
    unsigned long long h[2]; cudaMemcpy(h, chk, 16, cudaMemcpyDeviceToHost);
    printf("permutation check: %llu violations, %.2f %% of elements move\n", h[0], 100.0 * h[1] / n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (unsigned pct : {0u, 13u}) {
        for (int which = 0; which < 3; ++which) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_soa<<<n / 256, 256>>>(s, d, n, pct);
                else if (which == 1) k_aos<<<n / 256, 256>>>((const float4*)a, (float4*)b, n, pct);
                else k_aos256<<<n / 256, 256>>>(a, b, n, pct);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("%-22s moved %2u %%: %7.3f ms  %7.1f GB/s (64 B/particle)\n",
                   which == 0 ? "SoA 8 x 4 B" : which == 1 ? "AoS 2 x 128-bit" : "AoS 1 x 256-bit", pct, best, 64.0 * n / best / 1e6);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

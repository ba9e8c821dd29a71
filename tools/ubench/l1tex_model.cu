// Micro-benchmark (end of round 1; results: profiles/r03_ubench_l1tex_model.log): what does ONE warp-wide memory
// instruction of each kind used by k_push2 cost the SM, as a function of how many distinct 128-byte lines its 32 lanes
// touch?  The kernel's knock-outs add up linearly and neither occupancy nor an L1-resident gather moves the total
// (DESIGN.md 9.1), while the busiest unit ncu reports is the L1TEX data pipe at 70 % -- this measures the cost model
// directly instead of guessing it:
//    pattern 0  LDG.128, read-only path      lanes spread over D distinct lines of an L2-resident table
//    pattern 1  RED.ADD.V4.F32 (global)      lanes spread over D distinct lines
//    pattern 2  ATOMG.ADD.U32 with its result consumed right away (the slot claim), D distinct words,
//               optionally (bit 8 of the pattern) with pattern-1 reductions in flight from the same warp
//    pattern 3  STG.256 (whole record)       D distinct lines (D = 8: a contiguous 1 KB tile)
//    pattern 4  LDS.128 / STS.128 pair, conflict-free (the deposit rows)
//    pattern 5  LDG.256 (whole record)       6  the record store as two STG.128
// Persistent grid of 148 x 3 blocks x 8 warps like k_push2; every warp issues ITER instructions of the pattern on
// addresses that change per iteration (hash), all inside a table that fits L2 (default 32 MB).  Output: ns per
// warp-instruction per SM (= SM-cycles at the measured clock) for D = 1, 2, 4, 8, 16, 32.
//    nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l1tex_model l1tex_model.cu && ./l1tex_model
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
struct __align__(32) Rec { float v[8]; };

// lane -> byte offset inside the table: D distinct 128-byte lines per instruction, lanes of a line 16 B apart
__device__ __forceinline__ size_t lane_offset(unsigned it, int lane, int D, size_t nlines, unsigned salt) {
    const int group = lane % D;                                  // which of the D lines
    const size_t line = hash32(it * 64u + group + salt) % nlines;
    return line * 128 + (size_t)(lane / D % 8) * 16;
}

template <int PATTERN>
__global__ void __launch_bounds__(256, 3) k_model(char* table, size_t nlines, int D, int iters, float* sink) {
    __shared__ float4 rows[8][96];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned salt = (blockIdx.x * 8 + warp) * 2654435761u;
    float keep = 0.f;
    for (int it = 0; it < iters; ++it) {
        const size_t off = lane_offset(it, lane, D, nlines, salt);
        if (PATTERN == 0) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(table + off));
            keep += v.x + v.w;
        } else if (PATTERN == 1) {
            red4(reinterpret_cast<float*>(table + off), 1.f, 2.f, 3.f, 4.f);
        } else if (PATTERN == 2 || PATTERN == 2 + 256) {
            if (PATTERN & 256) red4(reinterpret_cast<float*>(table + ((off + 4096) % (nlines * 128))), 1.f, 2.f, 3.f, 4.f);
            const unsigned r = atomicAdd(reinterpret_cast<unsigned*>(table + off), 1u);
            keep += __shfl_sync(0xffffffffu, (float)r, (lane + 1) & 31);      // consumed at once, like claimed_slot()
        } else if (PATTERN == 3) {
            Rec r;
#pragma unroll
            for (int k = 0; k < 8; ++k) r.v[k] = keep + k;
            // D lines: D = 8 -> one contiguous KB; larger D -> every record in its own line
            const size_t line = hash32(it * 64u + (D <= 8 ? 0 : lane) + salt) % (nlines - 8);
            *reinterpret_cast<Rec*>(table + line * 128 + (D <= 8 ? (size_t)lane * 32 : 0)) = r;
        } else if (PATTERN == 5) {            // LDG.256: the record load (D <= 8: a contiguous KB, else one line per record)
            const size_t line = hash32(it * 64u + (D <= 8 ? 0 : lane) + salt) % (nlines - 8);
            const Rec r = *reinterpret_cast<const Rec*>(table + line * 128 + (D <= 8 ? (size_t)lane * 32 : 0));
            keep += r.v[0] + r.v[7];
        } else if (PATTERN == 6) {            // the same record store as two STG.128
            const size_t line = hash32(it * 64u + (D <= 8 ? 0 : lane) + salt) % (nlines - 8);
            float4* q = reinterpret_cast<float4*>(table + line * 128 + (D <= 8 ? (size_t)lane * 32 : 0));
            q[0] = make_float4(keep, 1.f, 2.f, 3.f);
            q[1] = make_float4(keep, 5.f, 6.f, 7.f);
        } else {
            rows[warp][lane * 3 + 0] = make_float4(keep, 1.f, 2.f, 3.f);
            __syncwarp();
            const float4 v = rows[warp][((lane + it) & 31) * 3];
            keep += v.x;
            __syncwarp();
        }
    }
    if (keep == -1.2345f) *sink = keep;
}

template <int PATTERN>
static void run(const char* name, char* table, size_t nlines, int iters, float* sink, double ghz) {
    printf("%-46s", name);
    for (int D = 1; D <= 32; D *= 2) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_model<PATTERN><<<148 * 3, 256>>>(table, nlines, D, iters / 8, sink);      // warm-up
        cudaEventRecord(e0);
        k_model<PATTERN><<<148 * 3, 256>>>(table, nlines, D, iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // 24 warps per SM issue `iters` instructions each: time per warp-instruction as seen by one SM
        const double ns = ms * 1e6 / ((double)iters * 24);
        printf("  D=%-2d %6.2f ns (%5.1f clk)", D, ns, ns * ghz);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    printf("\n");
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? (size_t)atoi(argv[1]) : 32;
    const int iters = argc > 2 ? atoi(argv[2]) : 20000;
    const size_t nlines = mb * 1024 * 1024 / 128;
    char* table; float* sink;
    cudaMalloc(&table, nlines * 128);
    cudaMemset(table, 0, nlines * 128);
    cudaMalloc(&sink, 4);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("table %zu MB (L2-resident when <= ~100), %d instructions per warp, 24 warps per SM, SM clock %.2f GHz\n", mb, iters, ghz);
    run<0>("LDG.128 read-only, D distinct lines", table, nlines, iters, sink, ghz);
    run<1>("RED.ADD.V4.F32, D distinct lines", table, nlines, iters, sink, ghz);
    run<2>("ATOMG.ADD.U32 + immediate use", table, nlines, iters, sink, ghz);
    run<2 + 256>("ATOMG.ADD.U32 + immediate use, REDs in flight", table, nlines, iters, sink, ghz);
    run<3>("STG.256 record store (D<=8: contiguous KB)", table, nlines, iters, sink, ghz);
    run<4>("STS.128 + LDS.128 + 2 syncwarp", table, nlines, iters, sink, ghz);
    run<5>("LDG.256 record load (D<=8: contiguous KB)", table, nlines, iters, sink, ghz);
    run<6>("2 x STG.128 record store (D<=8: contiguous KB)", table, nlines, iters, sink, ghz);
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

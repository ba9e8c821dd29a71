#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
// variants of "d = a*b (rounded) ; e = d + c (rounded)" with packed f32x2
__device__ __forceinline__ float2 v_plain(float2 a, float2 b, float2 c) { return __fadd2_rn(__fmul2_rn(a, b), c); }
__device__ __forceinline__ float2 v_fma0(float2 a, float2 b, float2 c) {
    float2 d = __ffma2_rn(a, b, make_float2(-0.0f, -0.0f));
    return __fadd2_rn(d, c);
}
__device__ __forceinline__ float2 v_asm(float2 a, float2 b, float2 c) {
    unsigned long long ua, ub, uc, ud, ue;
    ua = *reinterpret_cast<unsigned long long*>(&a); ub = *reinterpret_cast<unsigned long long*>(&b); uc = *reinterpret_cast<unsigned long long*>(&c);
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(ue) : "l"(ud), "l"(uc));
    return *reinterpret_cast<float2*>(&ue);
}
__device__ __forceinline__ float2 v_one(float2 a, float2 b, float2 c, float one) {
    return __ffma2_rn(__fmul2_rn(a, b), make_float2(one, one), c);
}
template <int V>
__global__ void k(const float2* a, const float2* b, const float2* c, float2* o, int n, float one = 1.0f) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (V == 0) o[i] = v_plain(a[i], b[i], c[i]);
    if (V == 1) o[i] = v_fma0(a[i], b[i], c[i]);
    if (V == 2) o[i] = v_asm(a[i], b[i], c[i]);
    if (V == 3) { float2 x = a[i], y = b[i], z = c[i]; o[i] = make_float2(__fadd_rn(__fmul_rn(x.x, y.x), z.x), __fadd_rn(__fmul_rn(x.y, y.y), z.y)); }
    if (V == 5) o[i] = v_one(a[i], b[i], c[i], one);
    if (V == 4) { float2 x = a[i], y = b[i], z = c[i]; o[i] = make_float2(fmaf(x.x, y.x, z.x), fmaf(x.y, y.y, z.y)); }
}
int main() {
    const int n = 1 << 20;
    float2 *a, *b, *c, *o[6];
    cudaMallocManaged(&a, n * 8); cudaMallocManaged(&b, n * 8); cudaMallocManaged(&c, n * 8);
    for (int v = 0; v < 6; ++v) cudaMallocManaged(&o[v], n * 8);
    srand(1);
    for (int i = 0; i < n; ++i) {
        a[i] = make_float2(rand() / (float)RAND_MAX + 0.5f, rand() / (float)RAND_MAX - 0.5f);
        b[i] = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX + 0.5f);
        c[i] = make_float2(-a[i].x * b[i].x * (1.f + 1e-6f), rand() / (float)RAND_MAX);
    }
    k<0><<<n / 256, 256>>>(a, b, c, o[0], n); k<1><<<n / 256, 256>>>(a, b, c, o[1], n); k<2><<<n / 256, 256>>>(a, b, c, o[2], n);
    k<5><<<n / 256, 256>>>(a, b, c, o[5], n, 1.0f); k<3><<<n / 256, 256>>>(a, b, c, o[3], n); k<4><<<n / 256, 256>>>(a, b, c, o[4], n);
    cudaDeviceSynchronize();
    const char* nm[6] = {"plain intrinsics", "fma(a,b,-0)+add", "inline asm .rn", "scalar _rn (truth unfused)", "scalar fmaf (truth fused)", "FMUL2 + FFMA2(p, runtime 1, c)"};
    for (int v = 0; v < 6; ++v) {
        long du = 0, df = 0;
        for (int i = 0; i < n; ++i) {
            du += (o[v][i].x != o[3][i].x) + (o[v][i].y != o[3][i].y);
            df += (o[v][i].x != o[4][i].x) + (o[v][i].y != o[4][i].y);
        }
        printf("%-28s differs from unfused: %8ld   from fused: %8ld\n", nm[v], du, df);
    }
    return 0;
}

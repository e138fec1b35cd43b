// alu_rate.cu — issue cost (cycles per warp instruction per scheduler) of the softmax's instruction kinds, with enough
// independent chains and warps that dependencies do not matter: FFMA2, FADD2, FFMA, FMNMX3-like max, cvt.bf16x2, MUFU.EX2.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int iters, float* out, long long* cyc) {
    float x[32];
    for (int i = 0; i < 32; ++i) x[i] = 1e-3f * (threadIdx.x + i) + 0.5f;
    unsigned acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            if (MODE == 0) asm volatile("{.reg .b64 a; mov.b64 a, {%0, %1}; fma.rn.f32x2 a, a, a, a; mov.b64 {%0, %1}, a;}" : "+f"(x[i]), "+f"(x[i + 1]));
            if (MODE == 1) asm volatile("{.reg .b64 a; mov.b64 a, {%0, %1}; add.rn.f32x2 a, a, a; mov.b64 {%0, %1}, a;}" : "+f"(x[i]), "+f"(x[i + 1]));
            if (MODE == 2) { asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i + 1])); }
            if (MODE == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[i + 1]));
            if (MODE == 4) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[i + 1])); acc ^= r; }
            if (MODE == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 32; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int per_iter, float* out, long long* cyc) {
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(500, out, cyc); cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(500, out, cyc); cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-22s %2d warps per scheduler: %.2f cycles per warp instruction per scheduler\n", name, warps / 4, (double)c / (500.0 * per_iter * (warps / 4)));
    }
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    run<0>("FFMA2 (fma.f32x2)", 16, out, cyc);
    run<1>("FADD2 (add.f32x2)", 16, out, cyc);
    run<2>("FFMA", 32, out, cyc);
    run<3>("FMNMX", 16, out, cyc);
    run<4>("F2FP.BF16 pack (+LOP)", 16, out, cyc);
    run<5>("MUFU.EX2", 16, out, cyc);
    return 0;
}

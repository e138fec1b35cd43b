// mufu_rate.cu — MUFU.EX2 / packed-FMA issue rates per SM sub-partition (cycles per warp instruction), 1-4 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int mode, int iters, float* out, long long* cyc) {
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = 1e-3f * (threadIdx.x + i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (mode == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            else if (mode == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
            else if (mode == 2) asm volatile("{.reg .b64 a; mov.b64 a, {%0, %1}; fma.rn.f32x2 a, a, a, a; mov.b64 {%0, %1}, a;}" : "+f"(x[i]), "+f"(x[(i + 1) & 15]));
            else if (mode == 3) asm volatile("{.reg .b32 a; mov.b32 a, %0; ex2.approx.ftz.bf16x2 a, a; mov.b32 %0, a;}" : "+f"(x[i]));   // two bf16 results per op
            else asm volatile("{.reg .b32 a; mov.b32 a, %0; ex2.approx.f16x2 a, a; mov.b32 %0, a;}" : "+f"(x[i]));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const char* names[] = {"MUFU.EX2", "FFMA", "FFMA2", "EX2.bf16x2", "EX2.f16x2"};
    for (int mode = 0; mode < 5; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {      // warps per CTA = warps per SM; 4 schedulers
            k<<<148, warps * 32>>>(mode, 1000, out, cyc); cudaDeviceSynchronize();
            k<<<148, warps * 32>>>(mode, 1000, out, cyc); cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-9s %2d warps/SM (%d per scheduler): %.2f cycles per warp instruction per scheduler\n", names[mode], warps, warps / 4,
                   (double)c / (1000.0 * 16 * (warps / 4)));
        }
    return 0;
}

// umma_rate.cu — what the tensor pipe charges for the MMA sequences of the prefill kernel, measured in isolation.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I physics_llm_inference_b200/csrc \
//        tools/probes/umma_rate.cu -o physics_llm_inference_b200/build/umma_rate && .../umma_rate
//
// One CTA per SM (grid 148, so clocks and power look like the real kernel), one issuing warp (elect_one), the same
// descriptors, instruction shapes and TMEM/smem layout as prefill_tcgen05.cu.  Every mode issues R repetitions of a
// pattern back to back, commits once, and reports cycles per repetition (clock64 of the issuing thread, CTA 0).
// Optional "noise": eight other warps run the softmax's TMEM traffic (tcgen05.ld 2 x 32 columns, tcgen05.st 2 x 16)
// in a loop, to see whether TMEM port contention slows the MMAs.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

using namespace pli;

constexpr int kSub = 128 * 128;          // [128 rows][64 el] bf16 sub-tile
constexpr int kTile = 2 * kSub;          // [128 x 128] tile
constexpr int kSmem = 6 * kTile + 1024 + 256;

struct Result {
    long long cycles;
    int reps;
};

__global__ void __launch_bounds__(320, 1) umma_rate_kernel(int mode, int reps, int noise, Result* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                       // 2 tiles
    uint8_t* sKV = smem + 2 * kTile;          // 4 tiles
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 6 * kTile);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 6 * kTile + 64);
    volatile int* stop = reinterpret_cast<volatile int*>(smem + 6 * kTile + 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 6 * kTile / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_init(bar + 2, 1);
        mbar_init(bar + 3, 1);
        *stop = 0;
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t kLboK = 1u << 16;
    constexpr uint32_t kLboV = (uint32_t)(kSub >> 4) << 16;
    constexpr uint32_t kIdescS64 = make_idesc_f16(128, 64, true, false, false);
    constexpr uint32_t kIdescS128 = make_idesc_f16(128, 128, true, false, false);
    constexpr uint32_t kIdescS256 = make_idesc_f16(128, 256, true, false, false);
    constexpr uint32_t kIdescO = make_idesc_f16(128, 128, true, false, true);

    if (warp == 0) {
        const uint32_t k_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboK;
        const uint32_t v_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboV;
        auto q_lo = [&](int t) { return (((smem_u32(sQ) + t * kTile) >> 4) & 0x3FFFu) | kLboK; };
        auto issue_S = [&](int t, int h, int slot, uint32_t idesc, int nk16) {      // S_t buffer h <- Q_t K^T
            const uint32_t ka = k_lo + slot * (kTile >> 4) + h * ((64 * 128) >> 4);
            for (int ks = 0; ks < nk16; ++ks) {
                const uint32_t koff = ((ks >> 2) * kSub + (ks & 3) * 32) >> 4;
                umma_ss_lohi(tmem_base + t * 128 + h * 64, q_lo(t) + koff, ka + koff, kDescHi, idesc, ks > 0);
            }
        };
        auto issue_PV = [&](int t, int h, int slot, int nk16) {                   // O_t += P_t(h) V
            const uint32_t va = v_lo + slot * (kTile >> 4) + h * ((64 * 128) >> 4);
            for (int ks = 0; ks < nk16; ++ks)
                umma_ts_lohi(tmem_base + 256 + t * 128, tmem_base + t * 128 + h * 64 + ks * 8, va + ks * (2048 >> 4),
                             kDescHi, kIdescO, 1u);
        };
        long long t0 = 0, t1 = 0;
        __syncwarp();
        t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < reps; ++r) {
                const int h = r & 1, ks = (r >> 1) & 1;
                switch (mode) {
                    case 0: issue_S(0, h, ks, kIdescS64, 8); break;                   // 8 x SS N64
                    case 1: issue_S(0, 0, ks, kIdescS128, 8); break;                  // 8 x SS N128
                    case 2: issue_PV(0, h, 1 + ks * 2, 4); break;                     // 4 x TS N128
                    case 3:                                                           // the kernel's half-step pair
                        issue_PV(0, h, 1 + ks * 2, 4);
                        issue_S(0, h, ks * 2, kIdescS64, 8);
                        issue_PV(1, h, 1 + ks * 2, 4);
                        issue_S(1, h, ks * 2, kIdescS64, 8);
                        break;
                    case 4:                                                           // same FLOPs, S as N128 (every 2nd rep)
                        issue_PV(0, h, 1 + ks * 2, 4);
                        if (h) issue_S(0, 0, ks * 2, kIdescS128, 8);
                        issue_PV(1, h, 1 + ks * 2, 4);
                        if (h) issue_S(1, 0, ks * 2, kIdescS128, 8);
                        break;
                    case 5: issue_S(0, 0, ks, kIdescS256, 8); break;                  // 8 x SS N256 (S0+S1 columns)
                    case 6:                                                           // tiles alternate per MMA group of 2
                        for (int g = 0; g < 4; ++g) {
                            issue_PV(g & 1, h, 1 + ks * 2, 2);
                        }
                        issue_S(0, h, ks * 2, kIdescS64, 8);
                        issue_S(1, h, ks * 2, kIdescS64, 8);
                        break;
                    case 7:                                                           // PV only, two accumulators
                        issue_PV(0, h, 1 + ks * 2, 4);
                        issue_PV(1, h, 1 + ks * 2, 4);
                        break;
                    case 8:                                                           // S only, two tiles
                        issue_S(0, h, ks * 2, kIdescS64, 8);
                        issue_S(1, h, ks * 2, kIdescS64, 8);
                        break;
                    case 9:                                                           // pair with the kernel's commits
                        issue_PV(0, h, 1 + ks * 2, 4);
                        umma_commit(bar + 1);
                        issue_S(0, h, ks * 2, kIdescS64, 8);
                        umma_commit(bar + 2);
                        umma_commit(bar + 3);
                        issue_PV(1, h, 1 + ks * 2, 4);
                        umma_commit(bar + 1);
                        issue_S(1, h, ks * 2, kIdescS64, 8);
                        umma_commit(bar + 2);
                        umma_commit(bar + 3);
                        break;
                    case 10:                                                          // commits merged behind each batch
                        issue_PV(0, h, 1 + ks * 2, 4);
                        issue_S(0, h, ks * 2, kIdescS64, 8);
                        umma_commit(bar + 1);
                        umma_commit(bar + 2);
                        umma_commit(bar + 3);
                        issue_PV(1, h, 1 + ks * 2, 4);
                        issue_S(1, h, ks * 2, kIdescS64, 8);
                        umma_commit(bar + 1);
                        umma_commit(bar + 2);
                        umma_commit(bar + 3);
                        break;
                    case 11:                                                          // one commit per batch
                        issue_PV(0, h, 1 + ks * 2, 4);
                        issue_S(0, h, ks * 2, kIdescS64, 8);
                        umma_commit(bar + 1);
                        issue_PV(1, h, 1 + ks * 2, 4);
                        issue_S(1, h, ks * 2, kIdescS64, 8);
                        umma_commit(bar + 2);
                        break;
                    default: break;
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        t1 = clock64();
        tc_fence_after();
        *stop = 1;
        if (blockIdx.x == 0 && lane == 0) {
            out->cycles = t1 - t0;
            out->reps = reps;
        }
    } else if (warp >= 2 && noise >= 3) {
        // softmax-like ALU work from 8 warps (two per scheduler), no memory traffic at all: per 64 "elements" 32 packed
        // FMAs, 48 MUFU.EX2, 32 packed adds, 32 bf16x2 packs; noise 4 adds the TMEM loads / stores of noise 2
        const int t = (warp - 2) >> 2;
        const uint32_t lane_addr = (uint32_t)(((warp - 2) & 3) * 32) << 16;
        float2 acc = make_float2(0.f, 0.f);
        uint32_t sink = 0;
        float seed = 1e-3f * (float)threadIdx.x;
        while (!*stop) {
            uint32_t a[32], b[32];
            if (noise == 4) {
                tmem_ld_x32(tmem_base + t * 128 + lane_addr, a);
                tmem_ld_x32(tmem_base + t * 128 + 32 + lane_addr, b);
                tc_wait_ld();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) { a[i] = __float_as_uint(seed + i); b[i] = __float_as_uint(seed - i); }
            }
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float2 x = ffma2(make_float2(__uint_as_float(a[i]) * 1e-30f, __uint_as_float(b[i]) * 1e-30f),
                                 make_float2(0.5f, 0.5f), make_float2(-1.f, -1.f));
                float2 e;
                if (i < 8) e = exp2_poly2(x);
                else { e.x = ex2_approx(x.x); e.y = ex2_approx(x.y); }
                acc = fadd2(acc, e);
                pk[i] = pack2<true>(e.x, e.y);
            }
            if (noise == 4) {
                tmem_st_x16(tmem_base + t * 128 + 64 + lane_addr, pk);
                tmem_st_x16(tmem_base + t * 128 + 80 + lane_addr, pk + 16);
                tc_wait_st();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) sink ^= pk[i];
            }
            seed += acc.x * 1e-30f;
        }
        if (sink == 0x12345678u || acc.y == 123.f) out->reps = -1;
    } else if (warp >= 2 && noise) {
        // softmax-like TMEM traffic from 8 warps (two per lane quarter): read 64 columns, write 32
        const int t = (warp - 2) >> 2;
        const uint32_t lane_addr = (uint32_t)(((warp - 2) & 3) * 32) << 16;
        uint32_t acc = 0;
        while (!*stop) {
            uint32_t a[32], b[32];
            tmem_ld_x32(tmem_base + t * 128 + lane_addr, a);
            tmem_ld_x32(tmem_base + t * 128 + 32 + lane_addr, b);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= a[i] + b[i];
            if (noise > 1) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0x3c003c00u;
                // write into columns the MMAs of this probe treat as P (harmless: all values are finite)
                tmem_st_x16(tmem_base + t * 128 + 64 + lane_addr, pk);
                tmem_st_x16(tmem_base + t * 128 + 80 + lane_addr, pk);
                tc_wait_st();
            }
        }
        if (acc == 0x12345678u) out->reps = -1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- the same patterns as CTA-pair MMAs (cta_group::2, M = 256: 128 rows from each CTA, B split between the CTAs) ----
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
umma_rate2_kernel(int mode, int reps, int noise, Result* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + 2 * kTile;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 6 * kTile);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 6 * kTile + 64);
    volatile int* stop = reinterpret_cast<volatile int*>(smem + 6 * kTile + 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 6 * kTile / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        *stop = 0;
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t kLboK = 1u << 16;
    constexpr uint32_t kLboV = (uint32_t)(kSub >> 4) << 16;
    constexpr uint32_t kIdescS64 = make_idesc_f16(256, 64, true, false, false);
    constexpr uint32_t kIdescS128 = make_idesc_f16(256, 128, true, false, false);
    constexpr uint32_t kIdescO = make_idesc_f16(256, 128, true, false, true);
    if (warp == 0 && rank == 0) {
        const uint32_t k_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboK;
        const uint32_t v_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboV;
        auto q_lo = [&](int t) { return (((smem_u32(sQ) + t * kTile) >> 4) & 0x3FFFu) | kLboK; };
        auto issue_S = [&](int t, int h, int slot, uint32_t idesc, int nk16) {
            const uint32_t ka = k_lo + slot * (kTile >> 4) + h * ((32 * 128) >> 4);     // 32 keys of the half-step live here
            for (int ks = 0; ks < nk16; ++ks) {
                const uint32_t koff = ((ks >> 2) * kSub + (ks & 3) * 32) >> 4;
                umma2_ss_lohi(tmem_base + t * 128 + h * 64, q_lo(t) + koff, ka + koff, kDescHi, idesc, ks > 0);
            }
        };
        auto issue_PV = [&](int t, int h, int slot, int nk16) {                      // this CTA holds 64 of V's 128 columns
            const uint32_t va = v_lo + slot * (kTile >> 4) + h * ((64 * 128) >> 4);
            for (int ks = 0; ks < nk16; ++ks)
                umma2_ts_lohi(tmem_base + 256 + t * 128, tmem_base + t * 128 + h * 64 + ks * 8, va + ks * (2048 >> 4),
                              kDescHi, kIdescO, 1u);
        };
        __syncwarp();
        const long long t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < reps; ++r) {
                const int h = r & 1, ks = (r >> 1) & 1;
                switch (mode) {
                    case 0: issue_S(0, h, ks, kIdescS64, 8); break;
                    case 1: issue_S(0, 0, ks, kIdescS128, 8); break;
                    case 2: issue_PV(0, h, 1 + ks * 2, 4); break;
                    case 3:
                        issue_PV(0, h, 1 + ks * 2, 4);
                        issue_S(0, h, ks * 2, kIdescS64, 8);
                        issue_PV(1, h, 1 + ks * 2, 4);
                        issue_S(1, h, ks * 2, kIdescS64, 8);
                        break;
                    default: break;
                }
            }
            asm volatile(
                "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                    smem_u32(bar)),
                "h"((uint16_t)1)
                : "memory");
        }
        __syncwarp();
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        tc_fence_after();
        if (blockIdx.x == 0 && lane == 0) {
            out->cycles = t1 - t0;
            out->reps = reps;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

int main(int argc, char** argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 400;
    Result* d;
    cudaMalloc(&d, sizeof(Result));
    cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    const char* names[] = {"8 x SS M128 N64 K16   (S half-step, floor 256)",
                           "8 x SS M128 N128 K16  (S full tile, floor 512)",
                           "4 x TS M128 N128 K16  (PV half-step, floor 256)",
                           "kernel pair: PV0 S0 PV1 S1 (N64)   (floor 1024)",
                           "pair with S as N128 every 2nd rep   (floor 1024)",
                           "8 x SS M128 N256 K16  (floor 1024)",
                           "PV0 PV1 interleaved by 2, S0 S1     (floor 1024)",
                           "PV0 PV1                             (floor 512)",
                           "S0 S1 (N64)                         (floor 512)",
                           "kernel pair + commits where the kernel has them",
                           "kernel pair + 3 commits behind each batch",
                           "kernel pair + 1 commit behind each batch"};
    for (int noise = 0; noise <= 4; ++noise) {
        if (noise == 1) continue;
        printf("--- noise level %d (0 none, 2 TMEM loads + stores, 3 softmax-like ALU work, 4 both) ---\n", noise);
        for (int mode = 0; mode < 12; ++mode) {
            for (int warm = 0; warm < 2; ++warm) {
                umma_rate_kernel<<<148, 320, kSmem>>>(mode, reps, noise, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
                    return 1;
                }
            }
            Result h;
            cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("mode %d  %-52s %8.1f cycles / rep\n", mode, names[mode], (double)h.cycles / h.reps);
        }
    }
    cudaFuncSetAttribute(umma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    printf("--- CTA pairs: cta_group::2, M = 256 (per-SM floors are the same) ---\n");
    for (int mode = 0; mode < 4; ++mode) {
        for (int warm = 0; warm < 2; ++warm) {
            umma_rate2_kernel<<<148, 320, kSmem>>>(mode, reps, 0, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("pair mode %d: %s\n", mode, cudaGetErrorString(e));
                return 1;
            }
        }
        Result h;
        cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("pair mode %d  %-47s %8.1f cycles / rep\n", mode, names[mode], (double)h.cycles / h.reps);
    }
    return 0;
}

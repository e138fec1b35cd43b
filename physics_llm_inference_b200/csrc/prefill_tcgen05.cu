// prefill_tcgen05.cu — FlashAttention forward for sm_100a: TMA + tcgen05.mma + TMEM.
//
// Replaces ch06/flash_attention.py:14-74 for bf16/f16, head_dim 64/128 (with the ch01/ch02 causal
// rule and GQA map).  One persistent CTA per SM; a work item is (batch, q head, pair of 128-row Q
// tiles).  16 warps, specialised:
//
//   warps 0-3   softmax for Q tile 0   one thread per row: S (TMEM) -> registers -> row max ->
//   warps 4-7   softmax for Q tile 1   exp2 -> P (bf16, written back over S in TMEM) -> row sum
//   warps 8-11  correction + epilogue  rescales O in TMEM when the running max moved (lazy, only
//                                      when it grew by > 2^8), final O/d -> smem -> TMA store, LSE
//   warp 12     MMA issuer (1 thread)  S_t = Q_t K_j^T (SS), O_t += P_t V_j (A = P from TMEM)
//   warp 13     TMA producer (1 thread) Q tiles, K/V ring
//
// TMEM (512 columns): S0 [0,128)  S1 [128,256)  O0 [256,256+D)  O1 [384,384+D); P_t aliases the
// first 64 columns of S_t.  The tensor pipe alternates between the two Q tiles
// (PV0_j, S0_{j+1}, PV1_j, S1_{j+1}), so each tile's softmax overlaps the other tile's MMAs.
//
// smem: Q 2 x [128 x D], K/V ring of 4 (D=128) / 8 (D=64) [128 x D] tiles, one [128 x 64] O staging
// sub-tile; every tile is stored as D/64 sub-tiles of [128 rows][64 el] with the 128-byte swizzle
// that TMA and UMMA share.
#include <type_traits>

#include "common.cuh"

namespace pli {
namespace {

constexpr int kBM = 128;               // rows per Q tile
constexpr int kBN = 128;               // keys per KV tile
constexpr int kThreads = 512;
constexpr int kSubTileBytes = 128 * 128;  // [128 rows][64 el] bf16
constexpr float kRescaleThreshold = 8.f;  // log2 units: rescale O only when the max grew by more
#ifndef PLI_POLY_PAIRS
#define PLI_POLY_PAIRS 0
#endif
constexpr int kPolyPairs = PLI_POLY_PAIRS;
#ifndef PLI_PROFILE
#define PLI_PROFILE 0                     // 1: compile the in-kernel timeline / phase counters (tuning builds only)
#endif
constexpr bool kProfile = PLI_PROFILE != 0;  // of every 16 element pairs, how many take the polynomial exp2

// named barrier ids (0 is __syncthreads)
constexpr int kBarEpilogue = 1;

template <int kD>
struct SmemLayout {
    static constexpr int kTileBytes = (kD / 64) * kSubTileBytes;   // one [128 x kD] tile
    static constexpr int kKVStages = kD == 128 ? 4 : 8;
    static constexpr int kQOff = 0;
    static constexpr int kKVOff = 2 * kTileBytes;
    static constexpr int kOOff = kKVOff + kKVStages * kTileBytes;  // one [128 x 64] staging sub-tile
    static constexpr int kScaleOff = kOOff + kSubTileBytes;        // float [2][128]
    static constexpr int kSumOff = kScaleOff + 2 * 128 * 4;        // float [2][128]
    static constexpr int kMaxOff = kSumOff + 2 * 128 * 4;          // float [2][128]
    static constexpr int kBarOff = kMaxOff + 2 * 128 * 4;
    static constexpr int kNumBars = 4 + 2 * kKVStages + 10;
    static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
    static constexpr int kTotal = kTmemPtrOff + 16;
};

struct PrefillParams {
    float* lse;
    int B, Hq, Hkv, Nq, Nk;
    int num_pairs;        // ceil(Nq / 256)
    int total_items;      // B * Hq * num_pairs
    int causal;
    float scale_log2;     // scale * log2(e)
    float scale;
    // debug aids (pli_debug_prefill_trace): both null / 0 in normal use
    unsigned long long* trace;   // [0] = record count, then (tag, clock) pairs written by CTA 0
    int trace_cap;
    int debug_flags;             // bit 0: skip the exp2 / P computation (timing experiments only)
};

// CTA 0 timeline: each tracing warp owns region `region` of the buffer and keeps its own cursor in a
// register (no atomics, stores are fire-and-forget); tag = event | tile << 8 | step << 16; SM-local clock.
__device__ __forceinline__ void trace_event(const PrefillParams& p, int lane, int region, int& cursor, int event, int t,
                                            int j) {
    if (kProfile && p.trace != nullptr && blockIdx.x == 0 && lane == 0) {
        if (cursor < p.trace_cap) {
            unsigned long long* dst = p.trace + ((size_t)region * p.trace_cap + cursor) * 2;
            dst[0] = (unsigned long long)(event | (t << 8) | (j << 16)) | (1ull << 40);
            dst[1] = clock64();
        }
        ++cursor;
    }
}

__device__ __forceinline__ int kv_tiles_for(int q0_tile, const PrefillParams& p) {
    // number of KV tiles a Q tile starting at row q0_tile attends to (>= 1)
    int kmax = p.Nk;
    if (p.causal) kmax = min(p.Nk, q0_tile + kBM + (p.Nk - p.Nq));
    kmax = max(kmax, 1);
    return (kmax + kBN - 1) / kBN;
}

struct WorkItem {
    int b, h, hk, q0;     // q0 = first row of Q tile 0; tile 1 starts at q0 + 128
    int n[2];             // KV tiles per Q tile
};

__device__ __forceinline__ WorkItem decode_item(int w, const PrefillParams& p) {
    // longest first: pair index descends as w grows; q heads of a KV group are adjacent in w
    WorkItem it;
    const int bh = p.B * p.Hq;
    const int pair = p.num_pairs - 1 - w / bh;
    const int r = w % bh;
    it.b = r / p.Hq;
    it.h = r % p.Hq;
    it.hk = it.h / (p.Hq / p.Hkv);
    it.q0 = pair * 2 * kBM;
    it.n[0] = kv_tiles_for(it.q0, p);
    it.n[1] = kv_tiles_for(it.q0 + kBM, p);
    if (it.n[1] < it.n[0]) it.n[1] = it.n[0];
    return it;
}

template <int kD, bool kBf16>
__global__ void __launch_bounds__(kThreads, 1)
prefill_tcgen05_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                       const PrefillParams p) {
    using L = SmemLayout<kD>;
    constexpr int kStages = L::kKVStages;
    constexpr int kTileBytes = L::kTileBytes;
    constexpr int kHalves = kD / 64;
    constexpr uint32_t kIdescS = make_idesc_f16(kBM, kBN, kBf16, false, false);  // Q K^T: both K-major
    constexpr uint32_t kIdescO = make_idesc_f16(kBM, kD, kBf16, false, true);    // P V: B (V) is MN-major

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + L::kQOff;
    uint8_t* sKV = smem + L::kKVOff;
    uint8_t* sO = smem + L::kOOff;
    float* sScale = reinterpret_cast<float*>(smem + L::kScaleOff);
    float* sSum = reinterpret_cast<float*>(smem + L::kSumOff);
    float* sMax = reinterpret_cast<float*>(smem + L::kMaxOff);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* q_full = bars;                  // [2]   TMA -> MMA
    uint64_t* q_empty = bars + 2;             // [2]   MMA (commit) -> TMA
    uint64_t* kv_full = bars + 4;             // [kStages]
    uint64_t* kv_empty = kv_full + kStages;   // [kStages]
    uint64_t* s_full = kv_empty + kStages;    // [2]   MMA (commit) -> softmax
    // pv_ok[t]: PV_t(j) may be issued.  Eight arrivals per tile-step: the four softmax warps once P_t(j) is
    // in TMEM, and the four correction warps once O_t can be accumulated into (rescaled for j >= 1; for
    // j == 0 drained by the previous item's epilogue, or free at kernel start).  One wait for the MMA warp
    // instead of three: every satisfied wait costs it ~200 cycles of tensor-pipe idle time.
    uint64_t* pv_ok = s_full + 2;             // [2]   softmax (4) + correction (4) -> MMA
    uint64_t* sc_full = pv_ok + 2;            // [2]   softmax (4 warps) -> correction: scale factor posted
    uint64_t* o_final = sc_full + 2;          // [2]   MMA (commit) -> correction: last PV done
    uint64_t* stats_full = o_final + 2;       // [2]   softmax (4 warps) -> correction: row sum / max posted
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&pv_ok[i], 8);           // four softmax warps + four correction warps
            mbar_init(&sc_full[i], 4);
            mbar_init(&o_final[i], 1);
            mbar_init(&stats_full[i], 4);
        }
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S[2] = {tmem_base, tmem_base + 128};
    const uint32_t tmem_O[2] = {tmem_base + 256, tmem_base + 384};

    if (warp < 8) {
        // =========================== softmax warpgroups ===========================
        // Warpgroup t owns Q tile t; one thread per row (no cross-thread reductions).
        reg_alloc<176>();
        const int t = warp >> 2;                          // Q tile of this warpgroup
        const int row = (warp & 3) * 32 + lane;           // row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t s_addr = tmem_S[t] + lane_addr;
        const float c = p.scale_log2;
        const int off = p.Nk - p.Nq;
        uint32_t step = 0;
        int trace_cur = 0;
        const bool prof = kProfile && p.trace != nullptr && blockIdx.x == 0;
        long long ph[6] = {0, 0, 0, 0, 0, 0};             // wait S, ld, max, exp+store, post, steps
        long long tp = 0;
        for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
            const WorkItem it = decode_item(w, p);
            const int q_row = it.q0 + t * kBM + row;
            float m_ref = -INFINITY;                      // reference max (raw score units)
            float d = 0.f;                                // running row sum relative to m_ref
            for (int j = 0; j < it.n[t]; ++j, ++step) {
                if (prof) tp = clock64();
                mbar_wait(&s_full[t], step & 1);
                tc_fence_after();
                if (prof) { const long long now = clock64(); ph[0] += now - tp; tp = now; }
                if ((warp & 3) == 0) trace_event(p, lane, t, trace_cur, 1, t, j);       // S ready
                float s[128];
                tmem_ld_x32(s_addr + 0, s + 0);
                tmem_ld_x32(s_addr + 32, s + 32);
                tmem_ld_x32(s_addr + 64, s + 64);
                tmem_ld_x32(s_addr + 96, s + 96);
                tc_wait_ld();
                if (prof) { const long long now = clock64(); ph[1] += now - tp; tp = now; }
                // mask: key padding and the causal diagonal (warp-uniform test, per-row limit)
                const int k0 = j * kBN;
                const bool need_mask = (k0 + kBN > p.Nk) || (p.causal && (k0 + kBN - 1 > it.q0 + t * kBM + off));
                if (need_mask) {
                    int vis = p.Nk - 1 - k0;
                    if (p.causal) vis = min(vis, q_row + off - k0);
#pragma unroll
                    for (int i = 0; i < 128; ++i) s[i] = (i <= vis) ? s[i] : -INFINITY;
                }
                float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
                for (int i = 4; i < 128; i += 4) {
                    mx0 = fmaxf(mx0, s[i]);
                    mx1 = fmaxf(mx1, s[i + 1]);
                    mx2 = fmaxf(mx2, s[i + 2]);
                    mx3 = fmaxf(mx3, s[i + 3]);
                }
                const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_ref);
                if ((warp & 3) == 0) trace_event(p, lane, t, trace_cur, 2, t, j);       // row max done
                if (prof) { const long long now = clock64(); ph[2] += now - tp; tp = now; }
                float alpha = 1.f;
                if (j == 0) {
                    m_ref = m_new;                        // first tile: nothing accumulated yet
                } else if ((m_new - m_ref) * c > kRescaleThreshold) {
                    alpha = ex2_approx((m_ref - m_new) * c);
                    m_ref = m_new;
                    d *= alpha;
                }
                if (j > 0) {
                    sScale[t * 128 + row] = alpha;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sc_full[t]);
                }
                // P = exp2(S*c - m*c): packed FFMA2 for the argument, MUFU.EX2 for most elements and a
                // degree-3 polynomial on the FMA pipe for kPolyPairs of every 16 pairs (the SFU, at 16
                // exp2/clk/SM, is as scarce as the tensor pipe here); row sums in packed FADD2 chains.
                const float2 c2 = make_float2(c, c);
                const float2 nmc2 = make_float2(-m_ref * c, -m_ref * c);
                float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
                if (kProfile && (p.debug_flags & 1)) {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = __float_as_uint(s[ch * 32 + 2 * i]);
                        tmem_st_x16(s_addr + ch * 16, pk);
                    }
                    acc0.x = 1.f;
                } else
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float2 x = ffma2(make_float2(s[ch * 32 + 2 * i], s[ch * 32 + 2 * i + 1]), c2, nmc2);
                        float2 pv;
                        if (i < kPolyPairs) {
                            pv = exp2_poly2(x);
                        } else {
                            pv.x = ex2_approx(x.x);
                            pv.y = ex2_approx(x.y);
                        }
                        if (i & 1) acc1 = fadd2(acc1, pv); else acc0 = fadd2(acc0, pv);
                        pk[i] = pack2<kBf16>(pv.x, pv.y);
                    }
                    tmem_st_x16(s_addr + ch * 16, pk);    // P aliases S columns [0,64)
                }
                acc0 = fadd2(acc0, acc1);
                d += acc0.x + acc0.y;
                if (prof) { const long long now = clock64(); ph[3] += now - tp; tp = now; }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pv_ok[t]);
                if (prof) { const long long now = clock64(); ph[4] += now - tp; tp = now; ph[5] += 1; }
                if ((warp & 3) == 0) trace_event(p, lane, t, trace_cur, 3, t, j);       // P posted
            }
            sSum[t * 128 + row] = d;
            sMax[t * 128 + row] = m_ref * c;              // log2 units
            __syncwarp();
            if (lane == 0) mbar_arrive(&stats_full[t]);
        }
        if (prof && lane == 0 && (warp & 3) == 0) {           // per-phase cycle totals: region 3 of the trace buffer
            unsigned long long* dst = p.trace + ((size_t)3 * p.trace_cap) * 2 + t * 8;
            for (int i = 0; i < 6; ++i) dst[i] = (unsigned long long)ph[i];
        }
    } else if (warp < 12) {
        // =========================== correction + epilogue warpgroup ===========================
        reg_dealloc<96>();
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        uint32_t corr_cnt[2] = {0, 0}, item_cnt = 0;
        if (lane == 0) {                                  // first item: O_0 / O_1 are free
            mbar_arrive(&pv_ok[0]);
            mbar_arrive(&pv_ok[1]);
        }
        for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_cnt) {
            const WorkItem it = decode_item(w, p);
            for (int j = 1; j < it.n[1]; ++j) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (j >= it.n[t]) continue;
                    mbar_wait(&sc_full[t], corr_cnt[t] & 1);
                    ++corr_cnt[t];
                    const float alpha = sScale[t * 128 + row];
                    if (__any_sync(0xffffffffu, alpha != 1.f)) {
                        tc_fence_after();
#pragma unroll
                        for (int ch = 0; ch < kD / 32; ++ch) {
                            float orr[32];
                            tmem_ld_x32(tmem_O[t] + lane_addr + ch * 32, orr);
                            tc_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) orr[i] *= alpha;
                            tmem_st_x32(tmem_O[t] + lane_addr + ch * 32, orr);
                        }
                        tc_wait_st();
                        tc_fence_before();
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&pv_ok[t]);
                }
            }
            // ---- epilogue: O_t / d -> bf16 -> swizzled smem -> TMA store; LSE ----
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                mbar_wait(&stats_full[t], item_cnt & 1);
                mbar_wait(&o_final[t], item_cnt & 1);
                tc_fence_after();
                const float dsum = sSum[t * 128 + row];
                const float mlog2 = sMax[t * 128 + row];
                const float inv = 1.f / dsum;
                const int q_tile0 = it.q0 + t * kBM;
#pragma unroll
                for (int hf = 0; hf < kHalves; ++hf) {
                    // the previous TMA store must have finished reading sO before it is overwritten
                    if (warp == 8 && lane == 0) tma_store_wait_read<0>();
                    named_bar_sync(kBarEpilogue, 128);
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int ch = hf * 2 + c2;                   // 32-column chunk of O_t
                        float orr[32];
                        tmem_ld_x32(tmem_O[t] + lane_addr + ch * 32, orr);
                        tc_wait_ld();
                        if (ch == kD / 32 - 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&pv_ok[t]);    // O_t is in registers: next item's PV_t(0) may overwrite it
                        }
                        uint8_t* srow = sO + row * 128;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 val;
                            val.x = pack2<kBf16>(orr[8 * i + 0] * inv, orr[8 * i + 1] * inv);
                            val.y = pack2<kBf16>(orr[8 * i + 2] * inv, orr[8 * i + 3] * inv);
                            val.z = pack2<kBf16>(orr[8 * i + 4] * inv, orr[8 * i + 5] * inv);
                            val.w = pack2<kBf16>(orr[8 * i + 6] * inv, orr[8 * i + 7] * inv);
                            const int chunk = c2 * 4 + i;             // 16-byte chunk inside the 128-byte row
                            *reinterpret_cast<uint4*>(srow + ((chunk ^ (row & 7)) << 4)) = val;
                        }
                    }
                    fence_proxy_async();
                    named_bar_sync(kBarEpilogue, 128);
                    if (warp == 8 && lane == 0 && q_tile0 < p.Nq) {
                        tma_store_4d(&map_o, sO, hf * 64, q_tile0, it.h, it.b);
                        tma_store_commit();
                    }
                }
                if (p.lse != nullptr && q_tile0 + row < p.Nq)
                    p.lse[((int64_t)it.b * p.Hq + it.h) * p.Nq + q_tile0 + row] = (mlog2 + log2f(dsum)) * kLn2;
            }
        }
        if (warp == 8 && lane == 0) tma_store_wait_all<0>();
    } else {
        reg_dealloc<64>();
        if (warp == 12) {
            // =========================== MMA issuer ===========================
            // All 32 lanes run the (warp-uniform) control flow and the waits so the address arithmetic stays
            // on the uniform datapath; one elected lane issues tcgen05.mma / tcgen05.commit.
            // Descriptor high word is constant: SBO = 1024 B (8 rows x 128 B), version 1, SWIZZLE_128B.
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t kLboK = 1u << 16;                           // K-major: LBO unused (1)
            constexpr uint32_t kLboV = (uint32_t)(kSubTileBytes >> 4) << 16;  // MN-major: LBO = one sub-tile
            const uint32_t q_lo = ((smem_u32(sQ) >> 4) & 0x3FFFu) | kLboK;
            const uint32_t k_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboK;
            const uint32_t v_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboV;
            uint32_t kv_cnt = 0, item_par = 0;
            uint32_t p_par = 0;                                   // bit t: phase parity of pv_ok[t]
            int trace_cur = 0;
            auto issue_S = [&](int t, uint32_t kslot) {
                // S_t = Q_t K^T : K-major operands, 16 elements (32 bytes) of head_dim per instruction
                const uint32_t qa = q_lo + t * (kTileBytes >> 4), ka = k_lo + kslot * (kTileBytes >> 4);
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks) {
                    const uint32_t koff = ((ks >> 2) * kSubTileBytes + (ks & 3) * 32) >> 4;
                    umma_ss_lohi(tmem_base + t * 128, qa + koff, ka + koff, kDescHi, kIdescS, ks > 0 ? 1u : 0u);
                }
            };
            auto issue_PV = [&](int t, uint32_t vslot, uint32_t accumulate) {
                // O_t += P_t V : A = P from TMEM (8 columns per 16 keys), B = V MN-major (16 keys = 2048 B)
                const uint32_t va = v_lo + vslot * (kTileBytes >> 4);
#pragma unroll
                for (int ks = 0; ks < kBN / 16; ++ks)
                    umma_ts_lohi(tmem_base + 256 + t * 128, tmem_base + t * 128 + ks * 8, va + ks * (2048 >> 4), kDescHi,
                                 kIdescO, ks > 0 ? 1u : accumulate);
            };
            for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, item_par ^= 1) {
                const WorkItem it = decode_item(w, p);
                const int n0 = it.n[0], n1 = it.n[1];
                auto slot_of = [&](uint32_t idx) -> uint32_t { return (kv_cnt + idx) % kStages; };
                auto wait_kv = [&](uint32_t idx) {
                    mbar_wait(&kv_full[(kv_cnt + idx) % kStages], ((kv_cnt + idx) / kStages) & 1);
                };
                // ---- first S for both tiles ----
                wait_kv(0);
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    mbar_wait(&q_full[t], item_par);
                    tc_fence_after();
                    if (elect_one()) {
                        issue_S(t, slot_of(0));
                        umma_commit(&s_full[t]);
                        if ((t ? n1 : n0) == 1) umma_commit(&q_empty[t]);
                        if (t == 1) umma_commit(&kv_empty[slot_of(0)]);
                    }
                    __syncwarp();
                }
                for (int j = 0; j < n1; ++j) {
                    // K/V of this step: normally landed long ago (the ring runs ~1.5 steps ahead)
                    wait_kv(2 * j + 1);
                    if (j + 1 < n1) wait_kv(2 * j + 2);
#pragma unroll 1
                    for (int t = 0; t < 2; ++t) {
                        const int nt = t ? n1 : n0;
                        if (j >= nt) continue;
                        mbar_wait(&pv_ok[t], (p_par >> t) & 1);
                        p_par ^= 1u << t;
                        const bool more = j + 1 < nt;
                        tc_fence_after();
                        trace_event(p, lane, 2, trace_cur, 4, t, j);                        // inputs of PV_t(j) ready
                        if (elect_one()) {
                            issue_PV(t, slot_of(2 * j + 1), j > 0 ? 1u : 0u);              // O_t += P_t(j) V_j
                            if (!more) umma_commit(&o_final[t]);
                            if (t == 1) umma_commit(&kv_empty[slot_of(2 * j + 1)]);
                            if (more) {
                                issue_S(t, slot_of(2 * j + 2));                          // S_t(j+1) = Q_t K_{j+1}^T
                                umma_commit(&s_full[t]);
                                if (j + 2 == nt) umma_commit(&q_empty[t]);
                                if (t == 1) umma_commit(&kv_empty[slot_of(2 * j + 2)]);
                            }
                        }
                        __syncwarp();
                        trace_event(p, lane, 2, trace_cur, 5, t, j);                        // issued
                    }
                }
                kv_cnt += 2 * n1;
            }
        } else if (warp == 13 && lane == 0) {
            // =========================== TMA producer ===========================
            prefetch_tensormap(&map_q);
            prefetch_tensormap(&map_k);
            prefetch_tensormap(&map_v);
            prefetch_tensormap(&map_o);
            uint32_t kv_cnt = 0, item_par = 0;
            for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, item_par ^= 1) {
                const WorkItem it = decode_item(w, p);
                const int n_max = it.n[1];
                auto load_q = [&](int t) {
                    mbar_wait(&q_empty[t], item_par ^ 1);
                    mbar_arrive_expect_tx(&q_full[t], kTileBytes);
#pragma unroll
                    for (int hf = 0; hf < kHalves; ++hf)
                        tma_load_4d(sQ + t * kTileBytes + hf * kSubTileBytes, &map_q, &q_full[t], hf * 64,
                                    it.q0 + t * kBM, it.h, it.b);
                };
                auto load_kv = [&](const CUtensorMap* map, int j) {
                    const uint32_t slot = kv_cnt % kStages;
                    mbar_wait(&kv_empty[slot], ((kv_cnt / kStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&kv_full[slot], kTileBytes);
#pragma unroll
                    for (int hf = 0; hf < kHalves; ++hf)
                        tma_load_4d(sKV + slot * kTileBytes + hf * kSubTileBytes, map, &kv_full[slot], hf * 64, j * kBN,
                                    it.hk, it.b);
                    ++kv_cnt;
                };
                load_q(0);
                load_kv(&map_k, 0);
                load_q(1);
                load_kv(&map_v, 0);
                for (int j = 1; j < n_max; ++j) {
                    load_kv(&map_k, j);
                    load_kv(&map_v, j);
                }
            }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// UMMA self-test (debug aid, exported as pli_debug_umma_selftest): one 128x128xD tile through the
// exact descriptor / TMEM paths the prefill kernel uses.  S = A B^T (SS, K-major), then
// O = bf16(S * 1/64) C (A from TMEM, B MN-major).  Dumps S and O so a wrong descriptor bit can be
// localised from Python.
// ------------------------------------------------------------------------------------------------
template <int kD, bool kBf16>
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_c, float* __restrict__ s_out, float* __restrict__ o_out) {
    constexpr int kTileBytes = (kD / 64) * kSubTileBytes;
    constexpr int kHalves = kD / 64;
    constexpr uint32_t kIdescS = make_idesc_f16(128, 128, kBf16, false, false);
    constexpr uint32_t kIdescO = make_idesc_f16(128, kD, kBf16, false, true);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + kTileBytes;
    uint8_t* sC = smem + 2 * kTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTileBytes);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_ptr;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars[0], 3 * kTileBytes);
        for (int hf = 0; hf < kHalves; ++hf) {
            tma_load_4d(sA + hf * kSubTileBytes, &map_a, &bars[0], hf * 64, 0, 0, 0);
            tma_load_4d(sB + hf * kSubTileBytes, &map_b, &bars[0], hf * 64, 0, 0, 0);
            tma_load_4d(sC + hf * kSubTileBytes, &map_c, &bars[0], hf * 64, 0, 0, 0);
        }
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        for (int ks = 0; ks < kD / 16; ++ks) {
            const uint32_t koff = (ks >> 2) * kSubTileBytes + (ks & 3) * 32;
            umma_ss(tb, make_smem_desc_sw128(smem_u32(sA) + koff, 16, 1024),
                    make_smem_desc_sw128(smem_u32(sB) + koff, 16, 1024), kIdescS, ks > 0);
        }
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float sr[128];
    tmem_ld_x32(tb + lane_addr + 0, sr + 0);
    tmem_ld_x32(tb + lane_addr + 32, sr + 32);
    tmem_ld_x32(tb + lane_addr + 64, sr + 64);
    tmem_ld_x32(tb + lane_addr + 96, sr + 96);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 128; ++i) s_out[row * 128 + i] = sr[i];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
            pk[i] = pack2<kBf16>(sr[ch * 32 + 2 * i] * 0.015625f, sr[ch * 32 + 2 * i + 1] * 0.015625f);
        tmem_st_x16(tb + lane_addr + ch * 16, pk);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        for (int ks = 0; ks < 8; ++ks)
            umma_ts(tb + 256, tb + ks * 8, make_smem_desc_sw128(smem_u32(sC) + ks * 2048, kSubTileBytes, 1024), kIdescO,
                    ks > 0);
        umma_commit(&bars[2]);
    }
    mbar_wait(&bars[2], 0);
    tc_fence_after();
    for (int ch = 0; ch < kD / 32; ++ch) {
        float orr[32];
        tmem_ld_x32(tb + 256 + lane_addr + ch * 32, orr);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_out[row * kD + ch * 32 + i] = orr[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

unsigned long long* g_trace_buf = nullptr;
int g_trace_cap = 0;
int g_debug_flags = 0;

int make_map_4d(CUtensorMap* map, const void* base, int dtype, int D, int N, int H, int B, const int64_t* st) {
    EncodeTiledFn enc = get_encode_tiled();
    if (enc == nullptr) return set_error(PLI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const CUtensorMapDataType dt = dtype == PLI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    // dims fastest first: head_dim, token, head, batch; a broadcast (zero) stride on a size-1 dim is replaced
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
    auto fix = [&](int64_t s) -> cuuint64_t { return (cuuint64_t)(s > 0 ? s : (int64_t)D) * 2; };
    cuuint64_t strides[3] = {fix(st[2]), fix(st[1]), fix(st[0])};
    cuuint32_t box[4] = {64, 128, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PLI_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PLI_OK;
}

template <int kD, bool kBf16>
int launch_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo,
             const PrefillParams& p, cudaStream_t stream) {
    auto kern = prefill_tcgen05_kernel<kD, kBf16>;
    const int smem = SmemLayout<kD>::kTotal + 1024;
    PLI_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid = sm_count();
    if (grid <= 0) grid = 148;
    if (grid > p.total_items) grid = p.total_items;
    kern<<<grid, kThreads, smem, stream>>>(mq, mk, mv, mo, p);
    PLI_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return PLI_OK;
}

}  // namespace

int launch_prefill_tcgen05(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv,
                           int Nq, int Nk, int D, const int64_t* qs, const int64_t* ks, const int64_t* vs,
                           const int64_t* os, float scale, int causal, int dtype, cudaStream_t stream) {
    CUtensorMap mq, mk, mv, mo;
    int rc;
    if ((rc = make_map_4d(&mq, q, dtype, D, Nq, Hq, B, qs))) return rc;
    if ((rc = make_map_4d(&mk, k, dtype, D, Nk, Hkv, B, ks))) return rc;
    if ((rc = make_map_4d(&mv, v, dtype, D, Nk, Hkv, B, vs))) return rc;
    if ((rc = make_map_4d(&mo, o, dtype, D, Nq, Hq, B, os))) return rc;
    PrefillParams p;
    p.lse = lse;
    p.B = B;
    p.Hq = Hq;
    p.Hkv = Hkv;
    p.Nq = Nq;
    p.Nk = Nk;
    p.num_pairs = (Nq + 2 * kBM - 1) / (2 * kBM);
    const int64_t total = (int64_t)B * Hq * p.num_pairs;
    if (total > 0x7fffffff) return set_error(PLI_ERR_UNSUPPORTED, "too many work items");
    p.total_items = (int)total;
    p.causal = causal;
    p.scale = scale;
    p.scale_log2 = scale * kLog2e;
    p.trace = g_trace_buf;
    p.trace_cap = g_trace_cap;
    p.debug_flags = g_debug_flags;
    const bool bf16 = dtype == PLI_BF16;
    if (D == 128) return bf16 ? launch_t<128, true>(mq, mk, mv, mo, p, stream) : launch_t<128, false>(mq, mk, mv, mo, p, stream);
    return bf16 ? launch_t<64, true>(mq, mk, mv, mo, p, stream) : launch_t<64, false>(mq, mk, mv, mo, p, stream);
}

}  // namespace pli

using namespace pli;

// Debug aid: record CTA 0's pipeline timeline of the next prefill launches into `buf` (device memory of
// 3 regions x capacity x 16 bytes, zeroed by the caller); buf = NULL switches it off.  flags bit 0 skips exp2.
extern "C" int pli_debug_prefill_trace(void* buf, int capacity, int flags) {
    g_trace_buf = static_cast<unsigned long long*>(buf);
    g_trace_cap = buf ? capacity : 0;
    g_debug_flags = flags;
    return PLI_OK;
}

// Debug aid (not part of the reference-facing surface): a (128 x D), b (128 x D), c (128 x D) row-major
// device tensors; s_out (128 x 128) f32 = a b^T; o_out (128 x D) f32 = cast(s_out / 64) c.
extern "C" int pli_debug_umma_selftest(const void* a, const void* b, const void* c, float* s_out, float* o_out, int D,
                                       int dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if ((D != 64 && D != 128) || (dtype != PLI_BF16 && dtype != PLI_F16)) return set_error(PLI_ERR_UNSUPPORTED, "selftest: D in {64,128}, bf16/f16");
    CUtensorMap ma, mb, mc;
    const int64_t st[3] = {(int64_t)128 * D, (int64_t)128 * D, D};
    int rc;
    if ((rc = make_map_4d(&ma, a, dtype, D, 128, 1, 1, st))) return rc;
    if ((rc = make_map_4d(&mb, b, dtype, D, 128, 1, 1, st))) return rc;
    if ((rc = make_map_4d(&mc, c, dtype, D, 128, 1, 1, st))) return rc;
    const int smem = 3 * (D / 64) * kSubTileBytes + 64 + 1024;
#define PLI_ST(DD, BF)                                                                                   \
    do {                                                                                                 \
        auto kern = umma_selftest_kernel<DD, BF>;                                                        \
        PLI_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   \
        kern<<<1, 128, smem, stream>>>(ma, mb, mc, s_out, o_out);                                        \
    } while (0)
    if (D == 128) { if (dtype == PLI_BF16) PLI_ST(128, true); else PLI_ST(128, false); }
    else          { if (dtype == PLI_BF16) PLI_ST(64, true); else PLI_ST(64, false); }
#undef PLI_ST
    PLI_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return PLI_OK;
}

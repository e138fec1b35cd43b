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
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"

namespace pli {
namespace {

constexpr int kBM = 128;               // rows per Q tile
constexpr int kBN = 128;               // keys per KV tile
constexpr int kThreads = 512;
constexpr int kSubTileBytes = 128 * 128;  // [128 rows][64 el] bf16
constexpr float kRescaleThreshold = 8.f;  // log2 units: rescale O only when the max grew by more
// Measured with the correction warps sharing the exponentials (round 2, C2 causal, same box): 0 pairs 1268-1276 TFLOP/s,
// 2 pairs 1256-1274, 4 pairs 1230, 6 pairs 1211: with three busy warps per scheduler the issue slots are the scarce
// resource, and a polynomial exp2 costs ~8 issue cycles per element against 1 for MUFU.EX2.
#ifndef PLI_POLY_PAIRS
#define PLI_POLY_PAIRS 0
#endif
constexpr int kPolyPairs = PLI_POLY_PAIRS;   // of every 16 element pairs, how many take the polynomial exp2
// Of every 64-key half-step s >= 1, the LAST kCorrCols score columns of a row are exponentiated by the correction warp
// of the same TMEM lane quadrant instead of the softmax warp (which still reads all 64 columns for the row maximum):
// the softmax warps' serial time per half-step is what bounds the kernel, the four correction warps
// are otherwise idle, and three warps per scheduler hide each other's TMEM / MUFU latencies better than two.
#ifndef PLI_CORR_COLS
#define PLI_CORR_COLS 16
#endif
constexpr int kCorrCols = PLI_CORR_COLS;
// L2 eviction hints on the TMA traffic of the product kernel: Q and O tiles evict-first (touched once), K/V evict-last
// (re-read by every Q tile of the KV group).  Round 1 measured 1.09 GB of DRAM traffic per C2 launch against 0.67 GB
// algorithmic: the streamed 0.5 GB of Q and O pushed K/V out of L2.
#ifndef PLI_L2_HINTS
#define PLI_L2_HINTS 1
#endif
// Option (off): software-pipelined softmax (pipelined_step in the kernel; needs the 48 / 16 column split): the thread
// loads S(s+1) and folds it into the next row maximum between the exponentials of S(s).  Correct (all parity tests),
// 14 % SLOWER on C2 (1146-1158 against 1332 TFLOP/s, profiles/r02_ab_softmax_pipeline.log): S(s+1) overwrites P(s-1), so it
// is issued behind PV(s-1) and completes ~800-1400 cycles after P(s-1) was posted -- i.e. near the END of step s; a wait
// for it inside step s stalls the exponentials.
#ifndef PLI_SOFTMAX_PIPELINE
#define PLI_SOFTMAX_PIPELINE 0
#endif
// Columns of a hot half-step whose exponentials are issued before the step's row maximum is known (see half_step).
#ifndef PLI_SPEC_COLS
#define PLI_SPEC_COLS 0
#endif
constexpr int kSpecCols = PLI_SPEC_COLS;
static_assert(kSpecCols == 0 || kSpecCols == 8 || kSpecCols == 16 || kSpecCols == 24, "kSpecCols: 0, 8, 16 or 24");
// Option (off): epilogue without CTA-wide barriers -- every correction warp converts its own 32 rows in 32-column chunks
// into a private, double-buffered 2 KiB staging slot and issues its own TMA stores (32 x 32 boxes, 64-byte swizzle).
// Correct (all parity tests), not faster: N 512 547-556 against 561-564 TFLOP/s, N 128 202-204 against 209-210, C2 equal
// (profiles/r02_ab_warp_epilogue.log).  The in-kernel timeline shows why: the 2200-2900 cycles between "epilogue start" and
// "epilogue done" of a tile are mostly the wait for the item's last PV to complete, not the conversion and the stores.
// Default: the round-1 epilogue (a 64-column half of the tile at a time through one 16 KiB buffer).
// CTA-pair MMAs for the paged instance of the kernel (K/V halves assembled from page boxes, reported to the leader's barrier).
#ifndef PLI_PAGED_PAIR_MMA
#define PLI_PAGED_PAIR_MMA 1
#endif
constexpr bool kPagedPairMma = PLI_PAGED_PAIR_MMA != 0;
#ifndef PLI_KV_SUSPEND_NS
#define PLI_KV_SUSPEND_NS 1000         // suspend-time hint of the TMA producers' wait for a free K/V ring slot
#endif
#ifndef PLI_LD64
#define PLI_LD64 0                     // the softmax warps read a half-step's 64 scores with one tcgen05.ld.x64 instead of two x32
#endif
#ifndef PLI_CORR_SUSPEND_NS
#define PLI_CORR_SUSPEND_NS 1000       // suspend-time hint of the correction warps' wait for the posted scale factor
#endif
#ifndef PLI_WARP_EPILOGUE
#define PLI_WARP_EPILOGUE 0
#endif
#ifndef PLI_TILE1_DELAY
#define PLI_TILE1_DELAY 0
#endif
#ifndef PLI_SMEM_KEEP_SPACE
#define PLI_SMEM_KEEP_SPACE 1
#endif
#ifndef PLI_DIRECT_EPILOGUE
#define PLI_DIRECT_EPILOGUE 0
#endif
constexpr bool kDirectEpilogue = PLI_DIRECT_EPILOGUE != 0;
constexpr uint64_t kHintQ = PLI_L2_HINTS ? kL2EvictFirst : 0x1000000000000000ull;    // else: evict-normal
constexpr uint64_t kHintKV = PLI_L2_HINTS ? kL2EvictLast : 0x1000000000000000ull;
constexpr uint64_t kHintO = PLI_L2_HINTS ? kL2EvictFirst : 0x1000000000000000ull;
static_assert(kCorrCols >= 0 && kCorrCols <= 32 && kCorrCols % 8 == 0, "kCorrCols: 0, 8, 16, 24 or 32 of the 64 columns");
// Option (off): the correction warp loads its share of S(s) as soon as S(s) is complete (its own wait on s_full), i.e.
// BEFORE the softmax warp has posted the row maximum, to hide the ~200-cycle tensor-memory load behind that wait.
// Measured 4 % SLOWER on C2 (1245-1253 against 1293-1295 TFLOP/s, profiles/r02_ab_corr_prefetch.log): the extra barrier
// poll and the earlier TMEM traffic cost more than the latency they hide.
#ifndef PLI_CORR_PREFETCH
#define PLI_CORR_PREFETCH 0
#endif
constexpr bool kCorrPrefetch = PLI_CORR_PREFETCH != 0;
#ifndef PLI_PROFILE
#define PLI_PROFILE 0                     // 1: compile the in-kernel timeline / phase counters (tuning builds only)
#endif
// PLI_TUNING=1 (build --variant=tuning -DPLI_TUNING=1) compiles the experiment switches: the environment variables
// PLI_WIDE / PLI_PAIR_MMA / PLI_NO_CLUSTER / PLI_MAX_CTAS, the flag word of pli_debug_prefill_trace and the opt-in wide
// kernel.  The product library has none of them: its kernel choice depends on the problem alone.
#ifndef PLI_TUNING
#define PLI_TUNING (PLI_PROFILE != 0)
#endif
constexpr bool kProfile = PLI_PROFILE != 0;

// named barrier ids (0 is __syncthreads)
constexpr int kBarEpilogue = 1;

template <int kD>
struct SmemLayout {
    static constexpr int kTileBytes = (kD / 64) * kSubTileBytes;   // one [128 x kD] tile
    static constexpr int kKVStages = kD == 128 ? 4 : 8;
    static constexpr int kQOff = 0;
    static constexpr int kKVOff = 2 * kTileBytes;
    static constexpr int kOOff = kKVOff + kKVStages * kTileBytes;  // one [128 x 64] staging sub-tile
    static constexpr int kScaleOff = kOOff + kSubTileBytes;        // float [2 tiles][2 buffers][128]
    static constexpr int kSumOff = kScaleOff + 4 * 128 * 4;        // float [2][128]
    static constexpr int kMaxOff = kSumOff + 2 * 128 * 4;          // float [2][128]
    static constexpr int kNmcOff = kMaxOff + 2 * 128 * 4;          // float [2 tiles][2 buffers][128]: -m_ref * c of the step
    static constexpr int kBarOff = kNmcOff + 4 * 128 * 4;
    static constexpr int kNumBars = 4 + 4 * kKVStages + 22;       // the K/V ring has 2 x kKVStages half-size entries with pair MMAs
    static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
    static constexpr int kTotal = kTmemPtrOff + 16;
};

// Division by a run-time constant as multiply-high + shift (the multiplier is found on the host: ceil(2^p / d) with
// p = 31 + ceil(log2 d); exact for dividends below 2^31).  decode_item runs once per work item in EVERY warp role, and
// its six integer divisions, each a ~150-cycle dependent chain on a GPU without a divider, were ~1500 serial cycles per
// item and role (in-kernel timeline at N = 512: every role idled that long between items).
struct PrefillParams {
    float* lse;
    int B, Hq, Hkv, Nq, Nk;
    int num_pairs;        // row slots per head: ceil(Nq / 256) (row-pair items) or ceil(Nq / 128) (head-pair items)
    int total_items;      // B * Hq * num_pairs (row pairs) or B * Hq / 2 * num_pairs (head pairs)
    int head_pairs;       // 1: the two Q tiles of an item are the SAME 128 rows of two adjacent q heads of one KV
                          //    group (even group size): equal trip counts, no dead second tile for short Nq;
                          // 0: two consecutive 128-row blocks of one head
    int causal;
    float scale_log2;     // scale * log2(e)
    float scale;
    // debug aids (pli_debug_prefill_trace): both null / 0 in normal use
    unsigned long long* trace;   // [0] = record count, then (tag, clock) pairs written by CTA 0
    int trace_cap;
    int debug_flags;             // bit 0: skip the exp2 / P computation (timing experiments only)
    int pair_block;              // scheduling block (see decode_item)
    // decode_item's divisors: items per block, per (group x slots) of a full / of the last (short) block, per (group, slot),
    // kv heads; and G = Hq / Hkv
    FastDiv fd_block, fd_group_full, fd_group_last, fd_gi, fd_hkv;
    int group, cnt_last;
    // paged K/V (kPaged kernels): keys of sequence b come from the pools through its block-table row; Nk is then
    // per sequence (seq_lens[b]) and the mask is bottom-right aligned per sequence
    const int32_t* table;
    const int32_t* seq_lens;
    int table_stride, page_size, page_shift, layer, box_rows;       // page_size and box_rows are powers of two
    int box_rows_v;                   // pair MMAs over paged K/V: rows of a V box (box_rows: of a K box, <= 32)
    // ragged query lengths (paged kernels only): q / o are packed (total_q, Hq, D), rows [cu_q[b], cu_q[b+1]) belong
    // to sequence b; Nq is then the host's upper bound of the per-sequence lengths (it sizes the schedule)
    const int32_t* cu_q;
    uint8_t* o_base;                  // raw o pointer + element strides for the predicated store of ragged tiles
    int64_t o_st_tok, o_st_head, o_st_batch;
    int64_t lse_sb, lse_sh;           // lse index = b * lse_sb + h * lse_sh + (row inside the q tensor)
};

// Fused all-gather of O over NVLink peer memory (pli_prefill_fwd_scatter): every rank holds the FULL
// (B_total, Hq_total, Nq, D) output twice (double-buffered by step parity) in peer-mapped memory; the epilogue's TMA
// store of each finished O sub-tile is issued once per rank, into the buffer (*epoch + 1) & 1 of that rank, so the
// transfer rides behind the MMAs of the following tiles.  n == 0: plain local store through map_o.
struct PeerMaps {
    CUtensorMap maps[2][PLI_MAX_PEERS];   // [buffer][rank]: 4-D maps over the full output of that rank
    // the local output once more with 32-column x 32-row boxes (64-byte swizzle): the per-warp epilogue stores
    CUtensorMap o32;
    const uint32_t* epoch;
    int n, head_offset, batch_offset;
    int warp_store;                       // o32 is valid: n == 0 and the output is the caller's own tensor
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

constexpr int kHN = 64;                // keys per softmax / MMA half-step (half of a KV tile)

// O_t row *= alpha in tensor memory (the rare half-steps in which the running maximum grew by more than 2^8).  Out of
// line on purpose: inlined, its four LDTM / 32 FMUL / STTM groups sat between the hot blocks of the correction loop and
// every pass jumped over ~2.6 KB of cold code twice (instruction-fetch stalls after the taken branches).
template <int kD>
__device__ __noinline__ void rescale_o_rows(uint32_t o_addr, float alpha) {
#pragma unroll
    for (int ch = 0; ch < kD / 32; ++ch) {
        float orr[32];
        tmem_ld_x32(o_addr + ch * 32, orr);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) orr[i] *= alpha;
        tmem_st_x32(o_addr + ch * 32, orr);
    }
}

// P = exp2(S * c - m * c) for kCols scores of a row: packed FFMA2 for the argument, MUFU.EX2 (an FMA-pipe polynomial for
// kPolyPairs of every 16 pairs), row sums in two packed FADD2 chains, bf16/f16 pairs packed into pk[0, kCols / 2).
template <bool kBf16, int kCols>
__device__ __forceinline__ void exp_cols(const float* sv, float2 c2, float2 nmc2, float2& acc0, float2& acc1, uint32_t* pk) {
#pragma unroll
    for (int i = 0; i < kCols / 2; ++i) {
        const float2 x = ffma2(make_float2(sv[2 * i], sv[2 * i + 1]), c2, nmc2);
        float2 pv;
        if ((i & 7) < kPolyPairs / 2) {
            pv = exp2_poly2(x);
        } else {
            pv.x = ex2_approx(x.x);
            pv.y = ex2_approx(x.y);
        }
        if (i & 1) acc1 = fadd2(acc1, pv); else acc0 = fadd2(acc0, pv);
        pk[i] = pack2<kBf16>(pv.x, pv.y);
    }
}

__device__ __forceinline__ int half_steps_for(int q0_tile, const PrefillParams& p, int nk, int nq) {
    // number of 64-key half-steps a Q tile starting at row q0_tile attends to (>= 1; a tile past the last query
    // row still runs one half-step so that every role sees the same barrier phases, and stores nothing)
    int kmax = nk;
    if (p.causal) kmax = min(nk, q0_tile + kBM + (nk - nq));
    if (q0_tile >= nq) kmax = 1;
    kmax = max(kmax, 1);
    return (kmax + kHN - 1) / kHN;
}

struct WorkItem {
    int b, hk;            // batch, kv head
    int q0[2], h[2];      // first row and q head of Q tile 0 / 1
    int n[2];             // 64-key half-steps per Q tile (n[1] >= n[0])
    int n_kv;             // 128-key K/V tiles to load
    int nk;               // keys of this item's sequence (p.Nk, or seq_lens[b] with paged K/V)
    int nq;               // query rows of this item's sequence (p.Nq, or cu_q[b+1] - cu_q[b])
    int qbase, bq;        // TMA coordinates of the sequence's first q row: (token qbase, batch bq)
};

// Static persistent schedule, longest item first.  Round i hands items [i*G, (i+1)*G) to the G CTAs, in
// forward order on even rounds and reversed on odd rounds ("snake"), which cancels the within-round size
// gradient: C2 load imbalance 0.3 % instead of 3.4 % for plain round-robin.  Returns -1 when the CTA is done.
__device__ __forceinline__ int item_of_round(int i, const PrefillParams& p) {
    const int G = (int)gridDim.x, c = (int)blockIdx.x;
    const long long w = (long long)i * G + ((i & 1) ? G - 1 - c : c);
    return w < p.total_items ? (int)w : -1;
}

__device__ __forceinline__ WorkItem decode_item(int w, const PrefillParams& p) {
    // Longest first, in blocks of p.pair_block consecutive row slots.  Inside a block the order is
    // KV group (batch, kv head) -> slot -> item of the group (q head, or pair of q heads), so the ~148 items in
    // flight share few KV groups and their K/V stay L2-resident (modelled DRAM K/V traffic on C2: 1.9 GB for
    // blocks of 1, 1.2 GB for 2, 0.65 GB for 4; measured +2 %); item sizes inside a block differ by
    // < pair_block slots, so the snake schedule still balances.  Consecutive items (even, odd) are the same
    // slot of the same KV group whenever the per-group item count is even: that is what CTA pairs rely on.
    // All divisions are by launch constants (FastDiv).
    WorkItem it;
    const int G = p.group;
    uint32_t blk, r, g, sl, gi, b, hk;
    p.fd_block.divmod((uint32_t)w, blk, r);                        // block of pair_block slots, index inside it
    const int top = p.num_pairs - 1 - (int)blk * p.pair_block;     // largest slot index of this block
    if (top + 1 >= p.pair_block) p.fd_group_full.divmod(r, g, r);  // KV group; (slot, item) inside the group
    else p.fd_group_last.divmod(r, g, r);                          // the last block may be short (cnt_last slots)
    p.fd_gi.divmod(r, sl, gi);
    const int slot = top - (int)sl;
    p.fd_hkv.divmod(g, b, hk);
    it.b = (int)b;
    it.hk = (int)hk;
    if (p.head_pairs) {
        it.h[0] = it.hk * G + (int)gi * 2;
        it.h[1] = it.h[0] + 1;
        it.q0[0] = it.q0[1] = slot * kBM;
    } else {
        it.h[0] = it.h[1] = it.hk * G + (int)gi;
        it.q0[0] = slot * 2 * kBM;
        it.q0[1] = it.q0[0] + kBM;
    }
    it.nk = p.seq_lens != nullptr ? max(min(p.seq_lens[it.b], p.Nk), 0) : p.Nk;   // never past the host's bound (table width)
    it.nq = p.Nq;
    it.qbase = 0;
    it.bq = it.b;
    if (p.cu_q != nullptr) {
        it.qbase = p.cu_q[it.b];
        it.nq = p.cu_q[it.b + 1] - it.qbase;
        it.bq = 0;
    }
    it.n[0] = half_steps_for(it.q0[0], p, it.nk, it.nq);
    it.n[1] = half_steps_for(it.q0[1], p, it.nk, it.nq);
    if (it.n[1] < it.n[0]) it.n[1] = it.n[0];
    it.n_kv = (it.n[1] + 1) >> 1;
    return it;
}

// kCluster = 2: CTA pairs (thread-block clusters of two, same TPC) work on two q heads of the SAME KV group at
// the same Q-tile pair, so they need the same K/V tiles: each CTA TMA-loads half of every tile and MULTICASTS it
// into both CTAs' shared memory (L2 -> SM traffic of K/V halves), and a ring slot is refilled once the MMA warps
// of both CTAs have released it (tcgen05.commit multicast onto both CTAs' kv_empty barriers).
// kPaged: K/V tiles are assembled from block-table pages (one TMA box per page / half page, issued by the lanes of
// the producer warp) instead of one box per tile; everything downstream of shared memory is identical.
// kPairMma (cluster of 2, contiguous K/V, D = 128): the pair's MMAs are CTA-pair instructions (tcgen05.mma.cta_group::2,
// M = 256 = the 128 rows of tile t in each CTA), issued by the leader CTA (rank 0) for both.  Each CTA then holds only
// HALF of every B operand in shared memory — for S = Q K^T the 32 keys [64 h + 32 rank, +32) of each 64-key half-step,
// for O += P V the 64 head_dim columns [64 rank, +64) of all 128 keys — so the operand reads of the QK^T MMAs drop
// from 6 KiB to 5 KiB per 32-cycle MMA slot (they are shared-memory-bandwidth bound: tools/probes/umma_rate.cu), the
// K/V TMA writes into each CTA's shared memory halve, and no multicast is needed.  Barriers that gate the MMAs
// (q_full, kv_full, pv_ok) live in the leader CTA and collect both CTAs' TMA bytes / arrivals; everything the MMAs
// signal (s_full, kv_empty, q_empty, pv_tail, o_final) is committed to both CTAs.
template <int kD, bool kBf16, int kCluster, bool kPaged, bool kPairMma = false>
__global__ void __launch_bounds__(kThreads, 1)
prefill_tcgen05_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                       const PrefillParams p, const __grid_constant__ PeerMaps peers) {
    using L = SmemLayout<kD>;
    constexpr int kStages = L::kKVStages;
    constexpr int kTileBytes = L::kTileBytes;
    constexpr int kHalves = kD / 64;
    static_assert(!kPairMma || (kCluster == 2 && kD == 128), "pair MMAs: cluster of 2, D = 128");
    constexpr int kRing = kPairMma ? 2 * kStages : kStages;              // K/V ring entries ...
    constexpr int kEntryBytes = kPairMma ? kTileBytes / 2 : kTileBytes;  // ... of this size (same total)
    constexpr int kMmaM = kPairMma ? 2 * kBM : kBM;
    constexpr uint32_t kIdescS = make_idesc_f16(kMmaM, kHN, kBf16, false, false);  // Q K^T (64 keys): both K-major
    constexpr uint32_t kIdescO = make_idesc_f16(kMmaM, kD, kBf16, false, true);    // P V: B (V) is MN-major

    extern __shared__ uint8_t smem_raw[];
#if PLI_SMEM_KEEP_SPACE
    // an offset added to the __shared__ array (not a pointer rebuilt from an integer) keeps the address space: the scalar
    // traffic through shared memory (scale factors, row statistics) compiles to LDS / STS instead of generic LD / ST
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#else
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
#endif
    uint8_t* sQ = smem + L::kQOff;
    uint8_t* sKV = smem + L::kKVOff;
    uint8_t* sO = smem + L::kOOff;
    float* sScale = reinterpret_cast<float*>(smem + L::kScaleOff);   // [2 tiles][2 buffers][128]
    float* sSum = reinterpret_cast<float*>(smem + L::kSumOff);       // [2][128]
    float* sMax = reinterpret_cast<float*>(smem + L::kMaxOff);       // [2][128]
    float* sNmc = reinterpret_cast<float*>(smem + L::kNmcOff);       // [2 tiles][2 buffers][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* q_full = bars;                  // [2]      TMA -> MMA warp t
    uint64_t* q_empty = bars + 2;             // [2]      MMA warp t (commit) -> TMA
    uint64_t* kv_full = bars + 4;             // [kRing] TMA -> both MMA warps
    uint64_t* kv_empty = kv_full + kRing;     // [kRing] both MMA warps (commit, count 2) -> TMA
    // Index [t * 2 + h]: Q tile t, S/P buffer h (= half-step parity).  Each barrier has at most one phase
    // in flight, which is what lets S run two half-steps ahead of the softmax.
    uint64_t* s_full = kv_empty + kRing;      // [4]  MMA (commit) -> softmax: S_t(s) is in buffer h
    // pv_ok: PV_t(s) may be issued: four softmax-warp arrivals (P_t(s) written over S in buffer h) + four
    // correction-warp arrivals (O_t rescaled for s >= 1; drained by the previous item / free at start for s == 0).
    uint64_t* pv_ok = s_full + 4;             // [4]
    uint64_t* sc_full = pv_ok + 4;            // [4]  softmax (4 warps) -> correction: scale factor of step s posted
    // Rescaling O_t at half-step s needs PV_t(s-1) complete, and with S running two half-steps ahead the
    // softmax of step s does not imply that.  But S_t(s+1) is issued right behind PV_t(s-1), so the commit of
    // s_full for step s+1 covers it: the correction warps wait on that phase (without consuming it) in the rare
    // steps that rescale.  The last step has no S behind its predecessor PV: pv_tail[t] is committed once per
    // item right after PV_t(n-2).
    uint64_t* pv_tail = sc_full + 4;          // [2]  MMA (commit) -> correction (+2 spare slots)
    uint64_t* pv_done = pv_tail;              // (alias kept for the barrier-array layout below)
    uint64_t* o_final = pv_done + 4;          // [2]  MMA (commit) -> correction: last PV of the item done
    uint64_t* stats_full = o_final + 2;       // [2]  softmax (4 warps) -> correction: row sum / max posted
    uint64_t* stats_free = stats_full + 2;    // [2]  correction (4 warps) -> softmax: the stats slots were read (a short
                                              //      next item could otherwise finish before this one's epilogue ran)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
            mbar_init(&o_final[i], 1);
            mbar_init(&stats_full[i], 4);
            mbar_init(&stats_free[i], 4);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&pv_ok[i], kPairMma ? 16 : 8);      // pair MMAs: both CTAs' warps arrive at the leader
            mbar_init(&sc_full[i], 4);
            mbar_init(&pv_tail[i], 1);        // [0..1] used
        }
        for (int i = 0; i < kRing; ++i) {
            mbar_init(&kv_full[i], 1);
            // both MMA warps of every CTA in the cluster; with pair MMAs only the leader's two issue (and commit to both)
            mbar_init(&kv_empty[i], kPairMma ? 2 : 2 * kCluster);
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        if constexpr (kPairMma) {
            tmem_alloc_pair(tmem_ptr, 512);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_ptr, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (kCluster > 1) cluster_sync_all();   // peers' barriers are initialised before anything targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t cta_rank = kCluster > 1 ? cluster_ctarank() : 0;
    // pv_ok of the CTA whose MMA warps wait on it (the pair's leader with pair MMAs, else this CTA); lane 0 of a warp
    const uint32_t pv_ok_addr = kPairMma ? mapa_u32(smem_u32(pv_ok), 0) : smem_u32(pv_ok);
    auto arrive_pv_ok = [&](int bi) {
        if constexpr (kPairMma) mbar_arrive_cluster(pv_ok_addr + bi * 8);
        else mbar_arrive(&pv_ok[bi]);
    };
    // TMEM: S_t buffer h at column t*128 + h*64 (P aliases its first 32 columns); O_t at 256 + t*128.

    if (warp < 8) {
        // =========================== softmax warpgroups ===========================
        // Warpgroup t owns Q tile t; one thread per row (no cross-thread reductions).  Half-step s handles
        // keys [64 s, 64 s + 64) out of S buffer s & 1.
        reg_alloc<176>();
        const int t = warp >> 2;
        const int row = (warp & 3) * 32 + lane;           // row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const float c = p.scale_log2;
        uint32_t sf_par = 0;                              // bit h: phase parity of s_full[t*2+h]
        int trace_cur = 0;
        for (int rnd = 0, w; (w = item_of_round(rnd, p)) >= 0; ++rnd) {
            const uint32_t item_par = rnd & 1;
            const WorkItem it = decode_item(w, p);
            const int q_tile0 = it.q0[t];
            const int q_row = q_tile0 + row;
            const int nk = it.nk, off = it.nk - it.nq;
            const int nt = it.n[t];
            float m_ref = -INFINITY;                      // reference max (raw score units)
            float d = 0.f;                                // running row sum relative to m_ref
            // One half-step.  kInterior = true is the hot instance: s >= 1 and no key of the half-step is masked for any row
            // of the tile, so neither the mask code nor the first-step special cases are in its instruction stream (they
            // used to sit in the middle of the loop and cost instruction-fetch stalls: ncu showed ~5 % no_inst + ~5 %
            // branch_resolving samples in these warps); the cold instance serves step 0 and the masked steps, which are
            // always the LAST ones of an item (key padding, causal diagonal).
            auto half_step = [&](const int s, auto interior_tag) {
                constexpr bool kInterior = decltype(interior_tag)::value;
                const int h = s & 1;
                const uint32_t s_addr = tmem_base + t * 128 + h * 64 + lane_addr;
                mbar_wait(&s_full[t * 2 + h], (sf_par >> h) & 1);
                sf_par ^= 1u << h;
                tc_fence_after();
                if ((warp & 3) == 0) trace_event(p, lane, t, trace_cur, 1, t, s);       // S ready
                float sv[64];
#if PLI_LD64
                tmem_ld_x64(s_addr, sv);
#else
                tmem_ld_x32(s_addr + 0, sv + 0);
                tmem_ld_x32(s_addr + 32, sv + 32);
#endif
                tc_wait_ld();
                if ((warp & 3) == 0) trace_event(p, lane, t, trace_cur, 6, t, s);       // S in registers
                if constexpr (!kInterior) {
                    // mask: key padding and the causal diagonal (warp-uniform test, per-row limit)
                    const int k0 = s * kHN;
                    const bool need_mask = (k0 + kHN > nk) || (p.causal && (k0 + kHN - 1 > q_tile0 + off));
                    if (need_mask) {
                        int vis = nk - 1 - k0;
                        if (p.causal) vis = min(vis, q_row + off - k0);
#pragma unroll
                        for (int i = 0; i < 64; ++i) sv[i] = (i <= vis) ? sv[i] : -INFINITY;
                    }
                }
                // Speculative start (hot instance): the first kSpecCols exponentials are issued against the reference maximum the
                // row ALREADY has, before the maximum of this half-step is known -- it changes the reference only when it
                // exceeds it by more than the lazy-rescale threshold (rare after the first steps; the cold branch below then
                // recomputes those columns).  The MUFU work overlaps the serial max / post prefix of the step.
                [[maybe_unused]] uint32_t spk[kSpecCols > 0 ? kSpecCols / 2 : 1];
                [[maybe_unused]] float2 sacc0 = make_float2(0.f, 0.f), sacc1 = make_float2(0.f, 0.f);
                if constexpr (kInterior && kSpecCols > 0) {
                    const float nmc_old = m_ref == -INFINITY ? 0.f : -m_ref * c;
                    exp_cols<kBf16, kSpecCols>(sv, make_float2(c, c), make_float2(nmc_old, nmc_old), sacc0, sacc1, spk);
                }
                float mx0 = sv[0], mx1 = sv[1], mx2 = sv[2], mx3 = sv[3];
#pragma unroll
                for (int i = 4; i < 64; i += 4) {
                    mx0 = fmaxf(mx0, sv[i]);
                    mx1 = fmaxf(mx1, sv[i + 1]);
                    mx2 = fmaxf(mx2, sv[i + 2]);
                    mx3 = fmaxf(mx3, sv[i + 3]);
                }
                const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_ref);
                float alpha = 1.f;
                [[maybe_unused]] bool spec_ok = kInterior && kSpecCols > 0;
                if (!kInterior && s == 0) {
                    m_ref = m_new;                        // first half-step: nothing accumulated yet
                } else if ((m_new - m_ref) * c > kRescaleThreshold) {
                    alpha = ex2_approx((m_ref - m_new) * c);
                    m_ref = m_new;
                    d *= alpha;
                    spec_ok = false;
                }
                // a row that has seen no visible key yet (seq_len < its query's position: a caller error the kernel
                // survives) keeps m_ref = -inf: exponentiate against 0 so that P = 0 instead of NaN
                const float nmc = m_ref == -INFINITY ? 0.f : -m_ref * c;
                if (kInterior || s > 0) {
                    sScale[(t * 2 + h) * 128 + row] = alpha;
                    if constexpr (kCorrCols > 0) sNmc[(t * 2 + h) * 128 + row] = nmc;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sc_full[t * 2 + h]);
                }
                if ((warp & 3) == 0) trace_event(p, lane, t, trace_cur, 2, t, s);       // max known, scale factor posted
                // P = exp2(S*c - m*c) in 16-column chunks; P aliases the first 32 columns of its S buffer (8 per chunk).
                // From half-step 1 on, the last kCorrCols columns are left to the correction warp of this lane quadrant.
                const float2 c2 = make_float2(c, c);
                const float2 nmc2 = make_float2(nmc, nmc);
                float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
                constexpr int kSoft = kHN - kCorrCols;                // this warp's columns of a step s >= 1
                {
                    uint32_t pk[16];
                    if constexpr (kInterior && kSpecCols > 0) {
                        if (spec_ok) {
#pragma unroll
                            for (int i = 0; i < kSpecCols / 2; ++i) pk[i] = spk[i];
                            acc0 = sacc0;
                            acc1 = sacc1;
                        } else {
                            exp_cols<kBf16, kSpecCols>(sv, c2, nmc2, acc0, acc1, pk);
                        }
                        exp_cols<kBf16, 32 - kSpecCols>(sv + kSpecCols, c2, nmc2, acc0, acc1, pk + kSpecCols / 2);
                    } else {
                        exp_cols<kBf16, 32>(sv, c2, nmc2, acc0, acc1, pk);
                    }
                    tmem_st_x16(s_addr, pk);
                }
                if constexpr (kSoft > 32) {
                    uint32_t pk[(kSoft - 32) / 2];
                    exp_cols<kBf16, kSoft - 32>(sv + 32, c2, nmc2, acc0, acc1, pk);
                    tmem_st_cols<(kSoft - 32) / 2>(s_addr + 16, pk);
                }
                if constexpr (!kInterior && kCorrCols > 0) {
                    if (s == 0) {                                     // step 0: the whole row (warp-uniform)
                        uint32_t pk[kCorrCols / 2];
                        exp_cols<kBf16, kCorrCols>(sv + kSoft, c2, nmc2, acc0, acc1, pk);
                        tmem_st_cols<kCorrCols / 2>(s_addr + kSoft / 2, pk);
                    }
                }
                acc0 = fadd2(acc0, acc1);
                d += acc0.x + acc0.y;
                if ((warp & 3) == 0) trace_event(p, lane, t, trace_cur, 7, t, s);       // exp2 / pack / TMEM stores issued
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_pv_ok(t * 2 + h);
                if ((warp & 3) == 0) trace_event(p, lane, t, trace_cur, 3, t, s);       // P posted
            };
            // half-steps [1, n_plain) are interior: 64 (s + 1) <= nk and, when causal, 64 s + 63 <= q_tile0 + off
            int n_plain = min(nt, nk / kHN);
            if (p.causal) n_plain = min(n_plain, q_tile0 + off - (kHN - 1) >= 0 ? (q_tile0 + off - (kHN - 1)) / kHN + 1 : 0);
#if PLI_SOFTMAX_PIPELINE
            // ---- interior half-steps, software-pipelined across steps ----
            // While the exponentials of S(s) run (MUFU / issue bound), the thread loads S(s+1) 16 columns at a time and
            // folds them into the next row maximum, so the ~230-cycle tensor-memory load and the ~500-cycle maximum of
            // step s+1 hide inside step s instead of heading its chain (S runs two half-steps ahead, so S(s+1) is
            // normally complete).  cur holds S(s) with its row maximum known; nxt receives S(s+1).  Two copies of the
            // step with the arrays swapped, so nothing is moved between steps.
            auto max16 = [](const float* v, float m) -> float {
                float a = fmaxf(v[0], v[1]), b = fmaxf(v[2], v[3]);
#pragma unroll
                for (int i = 4; i < 16; i += 4) {
                    a = fmaxf(a, fmaxf(v[i], v[i + 1]));
                    b = fmaxf(b, fmaxf(v[i + 2], v[i + 3]));
                }
                return fmaxf(m, fmaxf(a, b));
            };
            auto pipelined_step = [&](const int s, float (&cur)[64], float (&nxt)[64], const float mx_cur, float& mx_next,
                                      const bool has_next) {
                constexpr int kSoft = kHN - kCorrCols;
                static_assert(kSoft == 48, "the pipelined step is written for the 48 / 16 column split");
                const int h = s & 1;
                const uint32_t s_addr = tmem_base + t * 128 + h * 64 + lane_addr;
                const uint32_t n_addr = tmem_base + t * 128 + (h ^ 1) * 64 + lane_addr;
                const float m_new = fmaxf(mx_cur, m_ref);
                float alpha = 1.f;
                if ((m_new - m_ref) * c > kRescaleThreshold) {
                    alpha = ex2_approx((m_ref - m_new) * c);
                    m_ref = m_new;
                    d *= alpha;
                }
                const float nmc = m_ref == -INFINITY ? 0.f : -m_ref * c;
                sScale[(t * 2 + h) * 128 + row] = alpha;
                sNmc[(t * 2 + h) * 128 + row] = nmc;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sc_full[t * 2 + h]);
                if (has_next) {
                    mbar_wait(&s_full[t * 2 + (h ^ 1)], (sf_par >> (h ^ 1)) & 1);
                    sf_par ^= 1u << (h ^ 1);
                    tc_fence_after();
                    tmem_ld_x16(n_addr + 48, nxt + 48);               // the correction warp's columns: needed for the maximum only
                }
                const float2 c2 = make_float2(c, c);
                const float2 nmc2 = make_float2(nmc, nmc);
                float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
                float mx = -INFINITY;
                uint32_t pk[16];
                exp_cols<kBf16, 16>(cur, c2, nmc2, acc0, acc1, pk);
                if (has_next) {
                    tc_wait_ld();
                    mx = max16(nxt + 48, mx);
                    tmem_ld_x16(n_addr, nxt);
                }
                exp_cols<kBf16, 16>(cur + 16, c2, nmc2, acc0, acc1, pk + 8);
                tmem_st_x16(s_addr, pk);
                if (has_next) {
                    tc_wait_ld();
                    mx = max16(nxt, mx);
                    tmem_ld_x16(n_addr + 16, nxt + 16);
                }
                exp_cols<kBf16, 16>(cur + 32, c2, nmc2, acc0, acc1, pk);
                tmem_st_x8(s_addr + 16, pk);
                if (has_next) {
                    tc_wait_ld();
                    mx = max16(nxt + 16, mx);
                    tmem_ld_x16(n_addr + 32, nxt + 32);
                }
                acc0 = fadd2(acc0, acc1);
                d += acc0.x + acc0.y;
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_pv_ok(t * 2 + h);
                if (has_next) {
                    tc_wait_ld();
                    mx = max16(nxt + 32, mx);
                }
                mx_next = mx;
            };
            half_step(0, std::false_type{});
            int s = 1;
            if (s < n_plain) {
                float sa[64], sb[64];
                float mxa = -INFINITY, mxb = -INFINITY;
                {
                    const uint32_t a_addr = tmem_base + t * 128 + 64 + lane_addr;      // step 1: buffer 1
                    mbar_wait(&s_full[t * 2 + 1], (sf_par >> 1) & 1);
                    sf_par ^= 2u;
                    tc_fence_after();
                    tmem_ld_x32(a_addr, sa);
                    tmem_ld_x32(a_addr + 32, sa + 32);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 64; i += 16) mxa = max16(sa + i, mxa);
                }
                for (;;) {
                    pipelined_step(s, sa, sb, mxa, mxb, s + 1 < n_plain);
                    if (++s >= n_plain) break;
                    pipelined_step(s, sb, sa, mxb, mxa, s + 1 < n_plain);
                    if (++s >= n_plain) break;
                }
            }
            for (; s < nt; ++s) half_step(s, std::false_type{});
#else
            for (int s = 0; s < nt; ++s) {
#if PLI_TILE1_DELAY > 0
                // A/B option: put the two Q tiles' softmax warpgroups in ANTI-phase.  Left alone they run in phase (both
                // start when Q and the first K tile land), so their exponentials and the correction warps' share collide on
                // the one MUFU pipe of each scheduler and are idle together in the load / max phases.
                if (t == 1 && s == 1 && nt > 8) {
                    const long long t0 = clock64();
                    while (clock64() - t0 < PLI_TILE1_DELAY) __nanosleep(64);
                }
#endif
                if (s >= 1 && s < n_plain) half_step(s, std::true_type{});
                else half_step(s, std::false_type{});
            }
#endif
            mbar_wait(&stats_free[t], item_par ^ 1);      // previous item's epilogue has read the slots
            sSum[t * 128 + row] = d;
            sMax[t * 128 + row] = m_ref * c;              // log2 units
            __syncwarp();
            if (lane == 0) mbar_arrive(&stats_full[t]);
        }
    } else if (warp < 12) {
        // =========================== correction + epilogue warpgroup ===========================
        reg_dealloc<96>();
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        uint32_t sc_par = 0, item_cnt = 0;                // sc_par bit t*2+h: phase parity of sc_full[t*2+h]
        int peer_buf = 0;
        if constexpr (!kPaged) {
            if (peers.n > 0) peer_buf = (int)((*peers.epoch + 1u) & 1u);
        }
        int trace_cur = 0;
        uint32_t sf_base = 0;                             // bit t*2+h: parity of s_full[t*2+h]'s first phase in this item
        if (lane == 0) {                                  // first item: O_0 / O_1 are free
            arrive_pv_ok(0);
            arrive_pv_ok(2);
        }
        for (int rnd = 0, w; (w = item_of_round(rnd, p)) >= 0; ++rnd, ++item_cnt) {
            const WorkItem it = decode_item(w, p);
            float d_corr[2] = {0.f, 0.f};                 // row sums of the columns this warp exponentiates (per Q tile)
            for (int s = 1; s < it.n[1]; ++s) {
                const int h = s & 1;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (s >= it.n[t]) continue;
                    const int bi = t * 2 + h;
                    [[maybe_unused]] float sv[kCorrCols > 0 ? kCorrCols : 1];
                    [[maybe_unused]] const uint32_t s_addr = tmem_base + t * 128 + h * 64 + lane_addr;
                    if constexpr (kCorrCols > 0 && kCorrPrefetch) {
                        // S_t(s) complete?  (This phase of s_full cannot be overtaken: S_t(s+2) is only issued behind
                        // PV_t(s), which needs this warp's arrival for step s.)  Then fetch the share now.
                        mbar_wait(&s_full[bi], ((sf_base >> bi) ^ (uint32_t)(s >> 1)) & 1u);
                        tc_fence_after();
                        tmem_ld_cols<kCorrCols>(s_addr + (kHN - kCorrCols), sv);
                    }
#if defined(PLI_CORR_SPIN) && PLI_CORR_SPIN
                    mbar_wait(&sc_full[bi], (sc_par >> bi) & 1);
#else
                    mbar_wait_relaxed<PLI_CORR_SUSPEND_NS>(&sc_full[bi], (sc_par >> bi) & 1);
#endif
                    sc_par ^= 1u << bi;
                    if (wq == 0) trace_event(p, lane, 4, trace_cur, 8, t, s);             // scale factor seen
                    const float alpha = sScale[bi * 128 + row];
                    if constexpr (kCorrCols > 0) {
                        // this warp's share of P_t(s): columns [64 - kCorrCols, 64) of the S buffer (the softmax warp has
                        // read the whole row and posted -m_ref * c next to the scale factor)
                        const float nmc = sNmc[bi * 128 + row];
                        const float cs = p.scale_log2;
                        if constexpr (!kCorrPrefetch) {
                            tc_fence_after();
                            tmem_ld_cols<kCorrCols>(s_addr + (kHN - kCorrCols), sv);
                        }
                        tc_wait_ld();
                        const int k0 = s * kHN, off = it.nk - it.nq;
                        const bool need_mask = (k0 + kHN > it.nk) || (p.causal && (k0 + kHN - 1 > it.q0[t] + off));
                        if (need_mask) {
                            int vis = it.nk - 1 - k0;
                            if (p.causal) vis = min(vis, it.q0[t] + row + off - k0);
                            vis -= kHN - kCorrCols;
#pragma unroll
                            for (int i = 0; i < kCorrCols; ++i) sv[i] = (i <= vis) ? sv[i] : -INFINITY;
                        }
                        const float2 c2 = make_float2(cs, cs), nmc2 = make_float2(nmc, nmc);
                        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
                        {
                            uint32_t pk[kCorrCols / 2];
                            exp_cols<kBf16, kCorrCols>(sv, c2, nmc2, acc0, acc1, pk);
                            tmem_st_cols<kCorrCols / 2>(s_addr + (kHN - kCorrCols) / 2, pk);
                        }
                        acc0 = fadd2(acc0, acc1);
                        d_corr[t] = d_corr[t] * alpha + (acc0.x + acc0.y);
                    }
                    const bool rescale = __any_sync(0xffffffffu, alpha != 1.f);
                    if (rescale) {
                        // PV_t(s-1) must have completed: covered by the commit behind S_t(s+1), or by pv_tail
                        if (s + 1 < it.n[t]) {
                            const int bj = t * 2 + (h ^ 1);
                            mbar_wait(&s_full[bj], ((sf_base >> bj) ^ ((s + 1) >> 1)) & 1);
                        } else {
                            mbar_wait(&pv_tail[t], item_cnt & 1);
                        }
                        tc_fence_after();
                        rescale_o_rows<kD>(tmem_base + 256 + t * 128 + lane_addr, alpha);
                        tc_wait_st();
                        tc_fence_before();
                    } else if constexpr (kCorrCols > 0) {
                        tc_wait_st();
                        tc_fence_before();
                    }
                    __syncwarp();
                    if (lane == 0) arrive_pv_ok(bi);
                    if (wq == 0) trace_event(p, lane, 4, trace_cur, 9, t, s);             // this warp's share of P posted
                }
            }
            // ---- epilogue: O_t / d -> bf16 -> swizzled smem -> TMA store; LSE ----
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (wq == 0) trace_event(p, lane, 4, trace_cur, 10, t, 0);                // epilogue of tile t starts
                mbar_wait_relaxed(&stats_full[t], item_cnt & 1);
                mbar_wait_relaxed(&o_final[t], item_cnt & 1);
                mbar_wait_relaxed(&pv_tail[t], item_cnt & 1);         // one phase per item: keeps its parity in step
                tc_fence_after();
                const float dsum = sSum[t * 128 + row] + d_corr[t];
                const float mlog2 = sMax[t * 128 + row];
                __syncwarp();
                if (lane == 0) mbar_arrive(&stats_free[t]);
                const float inv = dsum > 0.f ? 1.f / dsum : 0.f;      // no visible key: O = 0, LSE = -inf
                const int q_tile0 = it.q0[t];
                // ragged tile of a packed q tensor: a TMA box would spill into the next sequence's rows, so the rows
                // go from registers to global memory under a predicate (uniform over the 128 epilogue threads)
                bool direct = false;
                if constexpr (kPaged) direct = p.cu_q != nullptr && q_tile0 + kBM > it.nq;
                // (Sending EVERY tile this way — no staging, no TMA store, no CTA-wide barriers — was measured and is slower:
                // a warp's 16-byte stores to 32 different rows cost more than the barriers they save; the per-tile epilogue
                // went from ~2500 to ~3500 cycles and N = 512 from 560 to 500 TFLOP/s.  -DPLI_DIRECT_EPILOGUE=1 rebuilds it.)
                if (kDirectEpilogue && p.o_base != nullptr && peers.n == 0) direct = true;
                const bool warp_store = PLI_WARP_EPILOGUE && peers.warp_store && !direct;
                if (warp_store) {
                    // ---- per-warp epilogue: this warp's 32 rows, 32 columns at a time, two private 2 KiB slots ----
                    // (the round-1 epilogue below costs ~2400 cycles per tile: per 64-column half it waits for the previous
                    // TMA store to release the one staging buffer and crosses two 128-thread barriers; a short item pays
                    // that twice, one tile after the other, in front of the next item's first PV)
                    uint8_t* wbuf = sO + wq * 4096;
#pragma unroll
                    for (int ch = 0; ch < kD / 32; ++ch) {
                        uint8_t* buf = wbuf + (ch & 1) * 2048;
                        if (lane == 0) tma_store_wait_read<1>();      // the store that read this slot two chunks ago is done
                        __syncwarp();
                        float orr[32];
                        tmem_ld_x32(tmem_base + 256 + t * 128 + lane_addr + ch * 32, orr);
                        tc_wait_ld();
                        if (ch == kD / 32 - 1) {
                            tc_fence_before();
                            __syncwarp();
                            // O_t is in registers: the next item's PV_t(0) (buffer 0) may overwrite it
                            if (lane == 0) arrive_pv_ok(t * 2);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 val;
                            val.x = pack2<kBf16>(orr[8 * i + 0] * inv, orr[8 * i + 1] * inv);
                            val.y = pack2<kBf16>(orr[8 * i + 2] * inv, orr[8 * i + 3] * inv);
                            val.z = pack2<kBf16>(orr[8 * i + 4] * inv, orr[8 * i + 5] * inv);
                            val.w = pack2<kBf16>(orr[8 * i + 6] * inv, orr[8 * i + 7] * inv);
                            // 64-byte rows, 64-byte swizzle: 16-byte chunk i of row r sits at chunk i ^ ((r >> 1) & 3)
                            *reinterpret_cast<uint4*>(buf + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = val;
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            if (q_tile0 + wq * 32 < it.nq)
                                tma_store_4d_hint(&peers.o32, buf, ch * 32, it.qbase + q_tile0 + wq * 32, it.h[t], it.bq, kHintO);
                            tma_store_commit();                       // (an empty group keeps the slot accounting uniform)
                        }
                    }
                }
#pragma unroll
                for (int hf = 0; hf < (warp_store ? 0 : kHalves); ++hf) {
                    // the previous TMA store must have finished reading sO before it is overwritten
                    if (!direct) {
                        if (warp == 8 && lane == 0) tma_store_wait_read<0>();
                        named_bar_sync(kBarEpilogue, 128);
                    }
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int ch = hf * 2 + c2;                   // 32-column chunk of O_t
                        float orr[32];
                        tmem_ld_x32(tmem_base + 256 + t * 128 + lane_addr + ch * 32, orr);
                        tc_wait_ld();
                        if (ch == kD / 32 - 1) {
                            tc_fence_before();
                            __syncwarp();
                            // O_t is in registers: the next item's PV_t(0) (buffer 0) may overwrite it
                            if (lane == 0) arrive_pv_ok(t * 2);
                        }
                        uint8_t* srow = sO + row * 128;
                        uint8_t* grow = nullptr;
                        if constexpr (kPaged || kDirectEpilogue) {
                            if (direct && q_tile0 + row < it.nq)
                                grow = p.o_base + ((int64_t)it.bq * p.o_st_batch + (int64_t)(it.qbase + q_tile0 + row) * p.o_st_tok +
                                                   (int64_t)it.h[t] * p.o_st_head + ch * 32) * 2;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 val;
                            val.x = pack2<kBf16>(orr[8 * i + 0] * inv, orr[8 * i + 1] * inv);
                            val.y = pack2<kBf16>(orr[8 * i + 2] * inv, orr[8 * i + 3] * inv);
                            val.z = pack2<kBf16>(orr[8 * i + 4] * inv, orr[8 * i + 5] * inv);
                            val.w = pack2<kBf16>(orr[8 * i + 6] * inv, orr[8 * i + 7] * inv);
                            const int chunk = c2 * 4 + i;             // 16-byte chunk inside the 128-byte row
                            if (!direct) *reinterpret_cast<uint4*>(srow + ((chunk ^ (row & 7)) << 4)) = val;
                            else if (grow != nullptr) reinterpret_cast<uint4*>(grow)[i] = val;
                        }
                    }
                    if (!direct) {
                        fence_proxy_async();
                        named_bar_sync(kBarEpilogue, 128);
                        if (warp == 8 && lane == 0 && q_tile0 < it.nq) {
                            if (kPaged || peers.n == 0) {
                                tma_store_4d_hint(&map_o, sO, hf * 64, it.qbase + q_tile0, it.h[t], it.bq, kHintO);
                            } else {
                                for (int r = 0; r < peers.n; ++r)
                                    tma_store_4d(&peers.maps[peer_buf][r], sO, hf * 64, q_tile0, it.h[t] + peers.head_offset,
                                                 it.b + peers.batch_offset);
                            }
                            tma_store_commit();
                        }
                    }
                }
                if (wq == 0) trace_event(p, lane, 4, trace_cur, 11, t, 0);                // epilogue of tile t done
                if (p.lse != nullptr && q_tile0 + row < it.nq)
                    p.lse[it.b * p.lse_sb + it.h[t] * p.lse_sh + it.qbase + q_tile0 + row] = (mlog2 + log2f(dsum)) * kLn2;
                // s_full[t*2+h] completed ceil((n_t - h) / 2) phases in this item
                sf_base ^= (uint32_t)(((it.n[t] + 1) >> 1) & 1) << (t * 2);
                sf_base ^= (uint32_t)((it.n[t] >> 1) & 1) << (t * 2 + 1);
            }
        }
        if (lane == 0) tma_store_wait_all<0>();           // every correction warp's own stores (warp 8's in the staged form)
    } else {
        reg_dealloc<64>();
        if ((warp == 12 || warp == 13) && !(kPairMma && cta_rank != 0)) {
            // =========================== MMA issuers: warp 12 -> Q tile 0, warp 13 -> Q tile 1 ===========================
            // (pair MMAs: only in the leader CTA; tile t is then the 256 rows of both CTAs' Q tile t)
            // tcgen05.mma issue is nearly synchronous (the queue holds ~2 instructions), so each tile gets its own
            // issuing warp: while one waits on a barrier the other keeps the tensor pipe busy.  All 32 lanes run the
            // warp-uniform control flow (addresses stay in uniform registers); one elected lane issues.
            // Per half-step s: O_t += P_t(s) V[64 s .. 64 s + 64), then S_t(s+2) = Q_t K[64 (s+2) ..]^T into the
            // buffer P_t(s) just vacated, so S always runs two half-steps ahead of the softmax.
            const int t = warp - 12;
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B
            constexpr uint32_t kLboK = 1u << 16;                              // K-major: LBO unused (1)
            constexpr uint32_t kLboV = (uint32_t)(kSubTileBytes >> 4) << 16;  // MN-major: LBO = one sub-tile
            const uint32_t q_lo = (((smem_u32(sQ) + t * kTileBytes) >> 4) & 0x3FFFu) | kLboK;
            const uint32_t k_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboK;
            const uint32_t v_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboV;
            const uint32_t tmem_s = tmem_base + t * 128, tmem_o = tmem_base + 256 + t * 128;
            uint32_t kv_cnt = 0, item_par = 0, pv_par = 0;        // pv_par bit h: phase parity of pv_ok[t*2+h]
            auto release_kv = [&](uint64_t* bar) {                // this Q tile is done with a ring slot
                if constexpr (kPairMma) umma2_commit_multicast(bar, (uint16_t)3);
                else if constexpr (kCluster > 1) umma_commit_multicast(bar, (uint16_t)((1u << kCluster) - 1));
                else umma_commit(bar);
            };
            auto commit = [&](uint64_t* bar) {                    // to the waiters of this tile (in both CTAs of a pair)
                if constexpr (kPairMma) umma2_commit_multicast(bar, (uint16_t)3);
                else umma_commit(bar);
            };
            int trace_cur = 0;
            auto issue_S = [&](int s, uint32_t kslot) {
                // S_t(s) = Q_t K[rows 64 (s&1) .. +64 of the tile]^T : K-major, 32 bytes of head_dim per MMA
                // (pair MMAs: this CTA's ring entry holds the half-step's 32 keys [64 h + 32 rank, +32) as rows
                // [32 h, 32 h + 32) of two [64 rows][64 el] sub-tiles)
                constexpr int kRowsS = kPairMma ? kHN / 2 : kHN;              // B rows per half-step in this CTA
                constexpr int kSubK = kPairMma ? kSubTileBytes / 2 : kSubTileBytes;
                const uint32_t ka = k_lo + kslot * (kEntryBytes >> 4) + (s & 1) * ((kRowsS * 128) >> 4);
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks) {
                    const uint32_t qoff = ((ks >> 2) * kSubTileBytes + (ks & 3) * 32) >> 4;
                    const uint32_t koff = ((ks >> 2) * kSubK + (ks & 3) * 32) >> 4;
                    if constexpr (kPairMma)
                        umma2_ss_lohi(tmem_s + (s & 1) * 64, q_lo + qoff, ka + koff, kDescHi, kIdescS, ks > 0 ? 1u : 0u);
                    else
                        umma_ss_lohi(tmem_s + (s & 1) * 64, q_lo + qoff, ka + koff, kDescHi, kIdescS, ks > 0 ? 1u : 0u);
                }
            };
            auto issue_PV = [&](int s, uint32_t vslot) {
                // O_t += P_t(s) V[rows 64 (s&1) .. +64] : A = P from TMEM (8 columns per 16 keys), B = V MN-major
                // (pair MMAs: this CTA's ring entry holds head_dim columns [64 rank, +64) of all 128 keys: one sub-tile)
                const uint32_t va = v_lo + vslot * (kEntryBytes >> 4) + (s & 1) * ((kHN * 128) >> 4);
#pragma unroll
                for (int ks = 0; ks < kHN / 16; ++ks) {
                    if constexpr (kPairMma)
                        umma2_ts_lohi(tmem_o, tmem_s + (s & 1) * 64 + ks * 8, va + ks * (2048 >> 4), kDescHi, kIdescO,
                                      (s > 0 || ks > 0) ? 1u : 0u);
                    else
                        umma_ts_lohi(tmem_o, tmem_s + (s & 1) * 64 + ks * 8, va + ks * (2048 >> 4), kDescHi, kIdescO,
                                     (s > 0 || ks > 0) ? 1u : 0u);
                }
            };
            for (int rnd = 0, w; (w = item_of_round(rnd, p)) >= 0; ++rnd, item_par ^= 1) {
                const WorkItem it = decode_item(w, p);
                const int nt = it.n[t];
                auto slot_of = [&](uint32_t idx) -> uint32_t { return (kv_cnt + idx) % kRing; };
                auto wait_kv = [&](uint32_t idx) {
                    mbar_wait(&kv_full[(kv_cnt + idx) % kRing], ((kv_cnt + idx) / kRing) & 1);
                };
                // K tile j is ring entry 2j, V tile j is ring entry 2j+1.  after_S(s): bookkeeping once S_t(s) is issued
                auto after_S = [&](int s) {
                    commit(&s_full[t * 2 + (s & 1)]);
                    if ((s & 1) || s == nt - 1) release_kv(&kv_empty[slot_of(2 * (s >> 1))]);   // K tile done (this Q tile)
                    if (s == nt - 1) commit(&q_empty[t]);
                };
                // ---- prologue: S_t(0), S_t(1) from K tile 0 ----
                mbar_wait(&q_full[t], item_par);
                wait_kv(0);
                tc_fence_after();
                trace_event(p, lane, 2 + t, trace_cur, 12, t, 0);                           // Q and the first K tile landed
                if (elect_one()) {
                    issue_S(0, slot_of(0));
                    after_S(0);
                    if (nt > 1) {
                        issue_S(1, slot_of(0));
                        after_S(1);
                    }
                }
                __syncwarp();
                for (int s = 0; s < nt; ++s) {
                    const int h = s & 1;
                    if (h == 0) wait_kv(2 * (s >> 1) + 1);                           // V tile of this pair of half-steps
                    if constexpr (kPaged) {
                        // Rows of the V tile at or beyond the sequence end hold whatever the (recycled, never cleared:
                        // ch07/paged_memory.py:100-110) page holds, and 0 x NaN would poison the row: zero them in shared
                        // memory before the first PV that reads the tile.  Both MMA warps do it (same zeros), each
                        // before its own MMAs.  K needs nothing: scores of those keys are replaced by select.
                        const int r0 = max(it.nk - (s >> 1) * kBN, 0);
                        if constexpr (kPairMma) {
                            // pair MMAs: this (leader) CTA and its peer each hold [128 keys][64 head_dim columns] of the tile;
                            // the leader's MMA warp cleans both halves, the peer's through distributed shared memory
                            if (h == 0 && r0 < kBN) {
                                uint8_t* vt = sKV + slot_of(2 * (s >> 1) + 1) * kEntryBytes;
                                const uint32_t vt_peer = mapa_u32(smem_u32(vt), 1);
                                const int rows = kBN - r0;
                                for (int idx = lane; idx < rows * 8; idx += 32) {
                                    const int off = (r0 + (idx >> 3)) * 128 + (idx & 7) * 16;
                                    *reinterpret_cast<uint4*>(vt + off) = make_uint4(0u, 0u, 0u, 0u);
                                    st_shared_cluster_zero16(vt_peer + off);
                                }
                                asm volatile("fence.acq_rel.cluster;\n" ::: "memory");    // the peer's zeros are performed ...
                                fence_proxy_async_all();                                   // ... before the MMAs read them
                                __syncwarp();
                            }
                        } else
                        if (h == 0 && r0 < kBN) {
                            uint8_t* vt = sKV + slot_of(2 * (s >> 1) + 1) * kTileBytes;
                            const int rows = kBN - r0;
                            for (int idx = lane; idx < rows * 8 * kHalves; idx += 32) {
                                const int chunk = idx & 7, rr = idx >> 3;
                                const int hf = rr / rows, r = r0 + rr - hf * rows;
                                *reinterpret_cast<uint4*>(vt + hf * kSubTileBytes + r * 128 + chunk * 16) = make_uint4(0u, 0u, 0u, 0u);
                            }
                            fence_proxy_async();
                            __syncwarp();
                        }
                    }
                    const bool more = s + 2 < nt;
                    if (more && h == 0) wait_kv(2 * ((s + 2) >> 1));                 // K tile of the S two half-steps ahead
                    mbar_wait(&pv_ok[t * 2 + h], (pv_par >> h) & 1);
                    pv_par ^= 1u << h;
                    tc_fence_after();
                    trace_event(p, lane, 2 + t, trace_cur, 4, t, s);                        // inputs of PV_t(s) ready
                    if (elect_one()) {
                        issue_PV(s, slot_of(2 * (s >> 1) + 1));
                        if (s + 2 == nt || nt == 1) commit(&pv_tail[t]);                // no S follows this PV
                        if (s == nt - 1) commit(&o_final[t]);
                        if (h || s == nt - 1) release_kv(&kv_empty[slot_of(2 * (s >> 1) + 1)]);   // V tile done (this Q tile)
                        if (more) {
                            issue_S(s + 2, slot_of(2 * ((s + 2) >> 1)));
                            after_S(s + 2);
                        }
                    }
                    __syncwarp();
                    trace_event(p, lane, 2 + t, trace_cur, 5, t, s);                        // issued
                }
                // K/V tiles this Q tile never touches (tile 0 under the causal mask) still need its release.  Wait
                // for each to land first: an arrival for a ring entry that is not loaded yet would be counted in
                // the slot's previous phase and free it under the other tile.
                for (int j = (nt + 1) >> 1; j < it.n_kv; ++j) {
                    wait_kv(2 * j);
                    wait_kv(2 * j + 1);
                    if (elect_one()) {
                        release_kv(&kv_empty[slot_of(2 * j)]);
                        release_kv(&kv_empty[slot_of(2 * j + 1)]);
                    }
                    __syncwarp();
                }
                kv_cnt += 2 * it.n_kv;
            }
        } else if ((warp == 14 && (kPaged || lane == 0)) || (kPaged && warp == 15)) {
            // paged K/V: two producer warps (14: Q and the K tiles, 15: the V tiles), every lane issuing one page box,
            // because a tile is many small TMA operations there; otherwise one thread of warp 14 issues everything
            const bool do_k = !kPaged || warp == 14, do_v = !kPaged || warp == 15;
            // =========================== TMA producer ===========================
            if (lane == 0) {
                prefetch_tensormap(&map_q);
                prefetch_tensormap(&map_k);
                prefetch_tensormap(&map_v);
                prefetch_tensormap(&map_o);
            }
            uint32_t kv_cnt = 0, item_par = 0;
            const int rank = (int)cta_rank;
            const uint32_t q_full_leader = kPairMma ? mapa_u32(smem_u32(q_full), 0) : 0;
            const uint32_t kv_full_leader = kPairMma ? mapa_u32(smem_u32(kv_full), 0) : 0;
            const uint16_t cmask = (uint16_t)((1u << kCluster) - 1);
            for (int rnd = 0, w; (w = item_of_round(rnd, p)) >= 0; ++rnd, item_par ^= 1) {
                const WorkItem it = decode_item(w, p);
                auto load_q = [&](int t) {
                    if (lane == 0 && do_k) {
                        mbar_wait_relaxed(&q_empty[t], item_par ^ 1);   // (a spinning wait here measured neutral)
                        if constexpr (kPairMma) {
                            // both CTAs' Q tiles are counted by the leader's barrier (its MMA warp issues for the pair)
                            if (rank == 0) mbar_arrive_expect_tx(&q_full[t], 2 * kTileBytes);
#pragma unroll
                            for (int hf = 0; hf < kHalves; ++hf)
                                tma_load_4d_pair_hint(sQ + t * kTileBytes + hf * kSubTileBytes, &map_q, q_full_leader + t * 8,
                                                      hf * 64, it.qbase + it.q0[t], it.h[t], it.bq, kHintQ);
                        } else {
                            mbar_arrive_expect_tx(&q_full[t], kTileBytes);
#pragma unroll
                            for (int hf = 0; hf < kHalves; ++hf)
                                tma_load_4d_hint(sQ + t * kTileBytes + hf * kSubTileBytes, &map_q, &q_full[t], hf * 64,
                                                 it.qbase + it.q0[t], it.h[t], it.bq, kHintQ);
                        }
                    }
                };
                // paged: this lane's box of every K/V tile covers tile rows [row0, row0 + box_rows); its page id is
                // looked up once per tile j (K and V share it).  Boxes past the sequence end re-read the last valid
                // page (the byte count per tile stays fixed; those keys are masked).
                const int rows_cta = kBN / kCluster;
                // (pair MMAs: warp 14's lanes own the K boxes of this CTA's 32 keys of each half-step, warp 15's lanes the V
                // boxes of all 128 keys -- this CTA's 64 head_dim columns of them)
                const int n_boxes = !kPaged ? 0 : !kPairMma ? rows_cta / p.box_rows : do_k ? 64 / p.box_rows : kBN / p.box_rows_v;
                int row0 = rank * rows_cta + lane * p.box_rows;
                if constexpr (kPaged && kPairMma) {
                    const int per_hh = 32 / p.box_rows;            // K boxes per half-step in this CTA
                    row0 = do_k ? (lane / per_hh) * kHN + rank * 32 + (lane % per_hh) * p.box_rows : lane * p.box_rows_v;
                }
                auto page_of = [&](int j) -> int {
                    if (!kPaged || lane >= n_boxes) return 0;
                    const int key = max(min(j * kBN + row0, it.nk - 1), 0);
                    return p.table[(int64_t)it.b * p.table_stride + (key >> p.page_shift)];
                };
                auto load_kv = [&](const CUtensorMap* map, int j, int page, bool mine) {
                    if (!mine) {                          // the other producer warp's ring entry
                        ++kv_cnt;
                        return;
                    }
                    const uint32_t slot = kv_cnt % kRing;
                    mbar_wait_relaxed<PLI_KV_SUSPEND_NS>(&kv_empty[slot], ((kv_cnt / kRing) & 1) ^ 1);
                    if constexpr (kPairMma && kPaged) {
                        // pair MMAs over paged K/V: the same halves of the B operand as below, assembled from page boxes by
                        // the lanes of this warp and reported to the leader's barrier
                        if (rank == 0 && lane == 0) mbar_arrive_expect_tx(&kv_full[slot], 2 * kEntryBytes);
                        __syncwarp();
                        if (lane < n_boxes) {
                            const uint32_t bar = kv_full_leader + slot * 8;
                            const int brows = map == &map_k ? p.box_rows : p.box_rows_v;
                            const int in_page = max(min(j * kBN + row0, it.nk - 1), 0) & (p.page_size - 1);
                            const int slot0 = in_page & ~(brows - 1);                  // box-aligned slot inside the page
                            uint8_t* dst = sKV + slot * kEntryBytes;
                            if (map == &map_k) {
                                const int per_hh = 32 / p.box_rows;
                                const int hh = lane / per_hh, bi = lane % per_hh;
#pragma unroll
                                for (int hf = 0; hf < kHalves; ++hf)
                                    tma_load_5d_pair_hint(dst + hf * (kSubTileBytes / 2) + (hh * 32 + bi * p.box_rows) * 128, map, bar,
                                                          hf * 64, it.hk, slot0, p.layer, page, kHintKV);
                            } else {
                                tma_load_5d_pair_hint(dst + lane * p.box_rows_v * 128, map, bar, rank * 64, it.hk, slot0, p.layer, page,
                                                      kHintKV);
                            }
                        }
                        ++kv_cnt;
                        return;
                    } else if constexpr (kPairMma) {
                        // this CTA's half of the B operand, reported to the leader's barrier (lane 0 only runs this):
                        // K: keys [64 h + 32 rank, +32) of both half-steps h (map_k carries 32-row boxes);
                        // V: head_dim columns [64 rank, +64) of all 128 keys (map_v carries 128-row boxes)
                        if (rank == 0) mbar_arrive_expect_tx(&kv_full[slot], 2 * kEntryBytes);
                        uint8_t* dst = sKV + slot * kEntryBytes;
                        const uint32_t bar = kv_full_leader + slot * 8;
                        if (map == &map_k) {
#pragma unroll
                            for (int hf = 0; hf < kHalves; ++hf)
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh)
                                    tma_load_4d_pair_hint(dst + hf * (kSubTileBytes / 2) + hh * (32 * 128), map, bar, hf * 64,
                                                          j * kBN + hh * kHN + rank * 32, it.hk, it.b, kHintKV);
                        } else {
                            tma_load_4d_pair_hint(dst, map, bar, rank * 64, j * kBN, it.hk, it.b, kHintKV);
                        }
                        ++kv_cnt;
                        return;
                    }
                    if (lane == 0) mbar_arrive_expect_tx(&kv_full[slot], kTileBytes);
                    if constexpr (kPaged) {
                        __syncwarp();
                        if (lane < n_boxes) {
                            const int in_page = max(min(j * kBN + row0, it.nk - 1), 0) & (p.page_size - 1);
                            const int slot0 = in_page & ~(p.box_rows - 1);         // box-aligned slot inside the page
#pragma unroll
                            for (int hf = 0; hf < kHalves; ++hf) {
                                uint8_t* dst = sKV + slot * kTileBytes + hf * kSubTileBytes + row0 * 128;
                                if constexpr (kCluster > 1)
                                    tma_load_5d_multicast_hint(dst, map, &kv_full[slot], hf * 64, it.hk, slot0, p.layer, page, cmask,
                                                               kHintKV);
                                else
                                    tma_load_5d_hint(dst, map, &kv_full[slot], hf * 64, it.hk, slot0, p.layer, page, kHintKV);
                            }
                        }
                    } else if constexpr (kCluster > 1) {
                        // this CTA fetches rows [64 rank, 64 rank + 64) of the tile for every CTA of the cluster
                        // (map_k / map_v carry 64-row boxes in this mode)
#pragma unroll
                        for (int hf = 0; hf < kHalves; ++hf)
                            tma_load_4d_multicast_hint(sKV + slot * kTileBytes + hf * kSubTileBytes + rank * (kHN * 128), map,
                                                       &kv_full[slot], hf * 64, j * kBN + rank * kHN, it.hk, it.b, cmask, kHintKV);
                    } else {
#pragma unroll
                        for (int hf = 0; hf < kHalves; ++hf)
                            tma_load_4d_hint(sKV + slot * kTileBytes + hf * kSubTileBytes, map, &kv_full[slot], hf * 64, j * kBN,
                                             it.hk, it.b, kHintKV);
                    }
                    ++kv_cnt;
                };
                int page = page_of(0);
                int page1 = page_of(it.n_kv > 1 ? 1 : 0);
                // the first K tile goes out BEFORE the wait for the Q buffers (q_empty fires two half-steps before the
                // previous item ends; K/V ring slots are usually free earlier), so it is in flight while Q is waited for
                load_kv(&map_k, 0, page, do_k);
                load_q(0);
                load_q(1);
                load_kv(&map_v, 0, page, do_v);
                for (int j = 1; j < it.n_kv; ++j) {
                    // the table entry of tile j was requested one tile ago: its latency is not on the refill path
                    const int next = page_of(j + 1 < it.n_kv ? j + 1 : j);
                    load_kv(&map_k, j, page1, do_k);
                    load_kv(&map_v, j, page1, do_v);
                    page1 = next;
                }

            }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if constexpr (kCluster > 1) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
    if (warp == 0) {
        if constexpr (kPairMma) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

#if PLI_TUNING
// ================================================================================================
// prefill_wide_kernel (D = 128, CTA pairs, contiguous K/V): ONE 128-row Q tile per CTA, S tiles 128 keys wide in THREE
// TMEM buffers, CTA-pair MMAs.
//
// Why: with two Q tiles per CTA the S buffers can only be 64 keys wide (2 x (2 x 64 + 128 O) = 512 TMEM columns), and a
// Q K^T MMA of N = 64 re-reads its A operand for every 64 keys: 6 KiB of shared memory per 32-cycle slot against a
// 128 B/clk port (1283 instead of 1024 cycles per 8.4 MFLOP).  One Q tile leaves room for three 128-key
// S buffers (3 x 128 + 128 O = 512): S MMAs run at N = 128 (at the tensor floor), S runs up to two 128-key steps ahead
// of the softmax, and the pair's MMAs (M = 256: this CTA's q head and its neighbour's, which share the KV head) halve
// the B-operand bytes each CTA holds and reads.
//
//   warps 0-7   softmax: warpgroup (w >> 2) takes every other 128-key step, one thread per row (see below)
//   warps 8-11  correction (lazy O rescale) + epilogue, as in prefill_tcgen05_kernel
//   warp 12     MMA issuer (leader CTA only)      warp 14  TMA producer (each CTA: its Q tile and its half of K / V)
//
// TMEM: S/P buffer b at column 128 b (P(j) = bf16 of 128 keys in its LAST 64 columns), O at 384.
// ================================================================================================
struct WideLayout {
    static constexpr int kTileBytes = 2 * kSubTileBytes;       // Q tile [128 x 128]
    static constexpr int kEntryBytes = kTileBytes / 2;         // this CTA's half of a K or V tile
    static constexpr int kRing = 8;
    static constexpr int kQOff = 0;
    static constexpr int kKVOff = kTileBytes;
    static constexpr int kOOff = kKVOff + kRing * kEntryBytes;   // one [128 x 64] staging sub-tile
    static constexpr int kScaleOff = kOOff + kSubTileBytes;      // float [3][128]
    static constexpr int kMrefOff = kScaleOff + 3 * 128 * 4;     // float [3][128]: reference maximum after step j (slot j % 3)
    static constexpr int kSumOff = kMrefOff + 3 * 128 * 4;       // float [<= 3 warpgroups][128]
    static constexpr int kMaxOff = kSumOff + 3 * 128 * 4;        // float [<= 3 warpgroups][128]
    static constexpr int kBarOff = kMaxOff + 3 * 128 * 4;
    static constexpr int kNumBars = 2 + 2 * kRing + 3 + 3 + 3 + 3 + 3 + 3;
    static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
    static constexpr int kTotal = kTmemPtrOff + 16;
};

struct WideItem {
    int b, hk, h, q0, n, nk, nq;      // batch, kv head, q head, first row, 128-key steps, keys, query rows
};

// Same block order as decode_item (blocks of pair_block 128-row slots, longest first, KV-group-major inside a block);
// consecutive items (2k, 2k + 1) are q heads (2i, 2i + 1) of one KV group at the same rows: the two CTAs of a pair.
__device__ __forceinline__ WideItem decode_wide(int w, const PrefillParams& p) {
    WideItem it;
    const int G = p.Hq / p.Hkv;
    const int per_slot = p.B * p.Hkv * G;
    const int kPB = p.pair_block;
    const int blk = w / (kPB * per_slot);
    int r = w - blk * kPB * per_slot;
    const int top = p.num_pairs - 1 - blk * kPB;
    const int cnt = min(kPB, top + 1);
    const int g = r / (cnt * G);
    r -= g * cnt * G;
    const int slot = top - r / G;
    it.b = g / p.Hkv;
    it.hk = g % p.Hkv;
    it.h = it.hk * G + r % G;
    it.q0 = slot * kBM;
    it.nk = p.Nk;
    it.nq = p.Nq;
    int kmax = it.nk;
    if (p.causal) kmax = min(it.nk, it.q0 + kBM + (it.nk - it.nq));
    kmax = max(kmax, 1);
    it.n = (kmax + kBN - 1) / kBN;
    return it;
}

// kGroups softmax warpgroups (2: 512 threads; 3: 576 threads, one warpgroup per S buffer, 112 registers per thread)
template <int kGroups>
struct WideRoles {
    static constexpr int kSoftWarps = 4 * kGroups;
    static constexpr int kCorrWarp0 = kSoftWarps;          // four correction / epilogue warps
    static constexpr int kMmaWarp = kSoftWarps + 4;
    static constexpr int kTmaWarp = kSoftWarps + 5;
    static constexpr int kThreadsWide = kGroups == 2 ? 512 : (kSoftWarps + 6) * 32;
};

template <bool kBf16, int kGroups>
__global__ void __launch_bounds__(WideRoles<kGroups>::kThreadsWide, 1)
prefill_wide_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                    const PrefillParams p, const __grid_constant__ PeerMaps peers) {
    using L = WideLayout;
    using R = WideRoles<kGroups>;
    constexpr int kD = 128;
    constexpr int kRing = L::kRing;
    constexpr int kEntryBytes = L::kEntryBytes;
    constexpr int kTileBytes = L::kTileBytes;
    constexpr uint32_t kIdescS = make_idesc_f16(2 * kBM, kBN, kBf16, false, false);   // Q K^T, 128 keys
    constexpr uint32_t kIdescO = make_idesc_f16(2 * kBM, kD, kBf16, false, true);     // P V

    extern __shared__ uint8_t smem_raw[];
#if PLI_SMEM_KEEP_SPACE
    // an offset added to the __shared__ array (not a pointer rebuilt from an integer) keeps the address space: the scalar
    // traffic through shared memory (scale factors, row statistics) compiles to LDS / STS instead of generic LD / ST
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#else
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
#endif
    uint8_t* sQ = smem + L::kQOff;
    uint8_t* sKV = smem + L::kKVOff;
    uint8_t* sO = smem + L::kOOff;
    float* sScale = reinterpret_cast<float*>(smem + L::kScaleOff);   // [3][128]
    float* sMref = reinterpret_cast<float*>(smem + L::kMrefOff);     // [3][128]
    float* sSum = reinterpret_cast<float*>(smem + L::kSumOff);       // [2][128]
    float* sMax = reinterpret_cast<float*>(smem + L::kMaxOff);       // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* q_full = bars;                   // leader: both CTAs' Q tiles (TMA bytes)
    uint64_t* q_empty = bars + 1;              // MMA commit (both CTAs) -> TMA
    uint64_t* kv_full = bars + 2;              // [kRing] leader: both CTAs' halves
    uint64_t* kv_empty = kv_full + kRing;      // [kRing] MMA commit (both CTAs) -> TMA
    uint64_t* s_full = kv_empty + kRing;       // [3] MMA commit (both CTAs) -> softmax: S(j) is in buffer j % 3
    uint64_t* pv_ok = s_full + 3;              // [3] leader: 2 x (4 softmax warps "P written" + 4 correction warps "O ready")
    uint64_t* sc_full = pv_ok + 3;             // [3] softmax (4 warps) -> correction: scale factor of step j
    uint64_t* mref_full = sc_full + 3;         // [3] softmax (4 warps) -> the other warpgroup: reference maximum after step j
    uint64_t* pv_done = mref_full + 3;         // [3] MMA commit (both CTAs) after PV(j), slot j % 3 -> correction (waited on when rescaling)
    uint64_t* o_final = pv_done + 3;           // MMA commit (both CTAs) after the last PV of an item -> epilogue
    uint64_t* stats_full = o_final + 1;        // softmax (8 warps) -> epilogue
    uint64_t* stats_free = stats_full + 1;     // epilogue (4 warps) -> softmax
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < kRing; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&pv_ok[i], 16);
            mbar_init(&sc_full[i], 4);
            mbar_init(&mref_full[i], 4);
        }
        for (int i = 0; i < 3; ++i) mbar_init(&pv_done[i], 1);
        mbar_init(o_final, 1);
        mbar_init(stats_full, R::kSoftWarps);
        mbar_init(stats_free, 4);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc_pair(tmem_ptr, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t cta_rank = cluster_ctarank();
    const uint32_t pv_ok_leader = mapa_u32(smem_u32(pv_ok), 0);
    const uint32_t tmem_o = tmem_base + 384;

    if (warp < R::kSoftWarps) {
        // =========================== softmax ===========================
        // Warpgroup g = warp >> 2 takes the 128-key steps j = g, g + kGroups, ...: one thread per row, the whole row of a step.
        // A row of 128 scores does not fit in registers next to its P, so the thread walks it in two 64-key halves:
        // load half 0 -> max, load half 1 -> max, (row max known) exp2 half 1 -> P, reload half 0 -> exp2 -> P.
        // P(j) goes into columns [64, 128) of its S buffer, so the reload of half 0 (columns [0, 64)) still sees S.
        // The two warpgroups run one step apart, so a step may take two MMA periods; the running reference maximum
        // of a row is handed from step to step through shared memory (sMref, mref_full).
        if constexpr (kGroups == 2) reg_alloc<176>();
        const int g = warp >> 2;
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const float cs = p.scale_log2;
        uint32_t ph_base = 0;                              // bit b: parity of the phases buffer b completed in earlier items
        for (int rnd = 0, w; (w = item_of_round(rnd, p)) >= 0; ++rnd) {
            const uint32_t item_par = rnd & 1;
            const WideItem it = decode_wide(w, p);
            const int q_row = it.q0 + row;
            const int nk = it.nk, off = it.nk - it.nq;
            float m_own = -INFINITY, d = 0.f;              // this thread's reference maximum and partial row sum
            for (int j = g; j < it.n; j += kGroups) {
                const int b = j % 3;
                const uint32_t par = ((ph_base >> b) ^ (uint32_t)(j / 3)) & 1u;
                const uint32_t s_addr = tmem_base + b * 128 + lane_addr;
                // A warpgroup waits on every other phase of a barrier, and a parity wait is only meaningful while the
                // barrier is at most one phase behind: S(j-1) complete (MMAs complete in order) implies that the phase
                // before the one waited for has completed on every buffer.
                // (With three warpgroups each one waits on every phase of its own buffer and none of this is needed.)
                if (kGroups == 2 && j > 0) {
                    const int bq = (j - 1) % 3;
                    mbar_wait(&s_full[bq], ((ph_base >> bq) ^ (uint32_t)((j - 1) / 3)) & 1u);
                }
                mbar_wait(&s_full[b], par);
                tc_fence_after();
                if (kProfile && (p.debug_flags & 8)) {          // tuning builds: barrier hand-offs only (wrong results)
                    sMref[b * 128 + row] = 0.f;
                    if (j > 0) sScale[b * 128 + row] = 1.f;
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&mref_full[b]);
                        if (j > 0) mbar_arrive(&sc_full[b]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(pv_ok_leader + b * 8);
                    continue;
                }
                float sv[64];
                auto load_half = [&](int hh) {
                    tmem_ld_x32(s_addr + hh * 64 + 0, sv + 0);
                    tmem_ld_x32(s_addr + hh * 64 + 32, sv + 32);
                    tc_wait_ld();
                    const int k0 = j * kBN + hh * 64;
                    const bool need_mask = (k0 + 64 > nk) || (p.causal && (k0 + 63 > it.q0 + off));
                    if (need_mask) {
                        int vis = nk - 1 - k0;
                        if (p.causal) vis = min(vis, q_row + off - k0);
#pragma unroll
                        for (int i = 0; i < 64; ++i) sv[i] = (i <= vis) ? sv[i] : -INFINITY;
                    }
                };
                auto max_of = [&]() -> float {
                    float mx0 = sv[0], mx1 = sv[1], mx2 = sv[2], mx3 = sv[3];
#pragma unroll
                    for (int i = 4; i < 64; i += 4) {
                        mx0 = fmaxf(mx0, sv[i]);
                        mx1 = fmaxf(mx1, sv[i + 1]);
                        mx2 = fmaxf(mx2, sv[i + 2]);
                        mx3 = fmaxf(mx3, sv[i + 3]);
                    }
                    return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                };
                load_half(0);
                const float mxa = max_of();
                load_half(1);
                const float row_max = fmaxf(mxa, max_of());
                // reference maximum after step j - 1 (posted by the other warpgroup); bring this thread's sum onto it
                float m_ref = row_max;
                float alpha = 1.f;
                if (j > 0) {
                    const int bp = (j - 1) % 3;
                    mbar_wait(&mref_full[bp], ((ph_base >> bp) ^ (uint32_t)((j - 1) / 3)) & 1u);
                    const float m_prev = sMref[bp * 128 + row];
                    if (m_prev != m_own) d *= ex2_approx((m_own - m_prev) * cs);
                    m_ref = m_prev;
                    if ((row_max - m_prev) * cs > kRescaleThreshold) {
                        alpha = ex2_approx((m_prev - row_max) * cs);
                        m_ref = row_max;
                        d *= alpha;
                    }
                }
                m_own = m_ref;
                sMref[b * 128 + row] = m_ref;
                if (j > 0) sScale[b * 128 + row] = alpha;
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&mref_full[b]);
                    if (j > 0) mbar_arrive(&sc_full[b]);
                }
                const float2 c2 = make_float2(cs, cs);
                const float2 nmc2 = make_float2(-m_ref * cs, -m_ref * cs);
                float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
                auto exp_half = [&](int hh) {                   // P of keys [64 hh, +64) -> columns 64 + 32 hh .. of the buffer
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float2 x = ffma2(make_float2(sv[ch * 32 + 2 * i], sv[ch * 32 + 2 * i + 1]), c2, nmc2);
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
                        tmem_st_x16(s_addr + 64 + hh * 32 + ch * 16, pk);
                    }
                };
                exp_half(1);
                load_half(0);
                exp_half(0);
                acc0 = fadd2(acc0, acc1);
                d += acc0.x + acc0.y;
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(pv_ok_leader + b * 8);
            }
            if (kGroups == 2 && ((it.n - 1) & 1) != g) {
                // the item's last step belongs to the other warpgroup: see its S and its reference maximum complete
                // before moving on, so that no barrier is ever more than one phase behind this warpgroup's next wait
                const int bl = (it.n - 1) % 3;
                const uint32_t pl = ((ph_base >> bl) ^ (uint32_t)((it.n - 1) / 3)) & 1u;
                mbar_wait(&s_full[bl], pl);
                mbar_wait(&mref_full[bl], pl);
            }
            mbar_wait(stats_free, item_par ^ 1);
            sSum[g * 128 + row] = d;
            sMax[g * 128 + row] = m_own * cs;                   // log2 units; -inf if this warpgroup had no step
            __syncwarp();
            if (lane == 0) mbar_arrive(stats_full);
#pragma unroll
            for (int bb = 0; bb < 3; ++bb)
                if (bb < it.n) ph_base ^= (uint32_t)(((it.n - bb + 2) / 3) & 1) << bb;
        }
    } else if (warp < R::kCorrWarp0 + 4) {
        // =========================== correction + epilogue ===========================
        if constexpr (kGroups == 2) reg_dealloc<96>();
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        uint32_t sc_par = 0, item_cnt = 0, pvd_base = 0;
        int peer_buf = 0;
        if (peers.n > 0) peer_buf = (int)((*peers.epoch + 1u) & 1u);
        if (lane == 0) mbar_arrive_cluster(pv_ok_leader);               // first item: O is free
        for (int rnd = 0, w; (w = item_of_round(rnd, p)) >= 0; ++rnd, ++item_cnt) {
            const WideItem it = decode_wide(w, p);
            for (int j = 1; j < it.n; ++j) {
                const int b = j % 3;
                mbar_wait_relaxed(&sc_full[b], (sc_par >> b) & 1);
                sc_par ^= 1u << b;
                const float alpha = sScale[b * 128 + row];
                const bool rescale = __any_sync(0xffffffffu, alpha != 1.f);
                if (rescale) {
                    // PV(j-1) must have completed before O is touched.  S runs up to two steps ahead of the PVs here, so a
                    // single barrier could be two phases ahead of, or behind, a parity wait:
                    // one barrier per step mod 3: PV(j-4) is known to be complete (S(j) complete implies PV(j-3) complete)
                    // and PV(j+2) cannot have been issued, so the barrier of step j-1 is exactly in, or just past, the
                    // phase waited for
                    {
                        const int bd = (j - 1) % 3;
                        mbar_wait(&pv_done[bd], ((pvd_base >> bd) ^ (uint32_t)((j - 1) / 3)) & 1u);
                    }
                    tc_fence_after();
#pragma unroll
                    for (int ch = 0; ch < kD / 32; ++ch) {
                        float orr[32];
                        tmem_ld_x32(tmem_o + lane_addr + ch * 32, orr);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) orr[i] *= alpha;
                        tmem_st_x32(tmem_o + lane_addr + ch * 32, orr);
                    }
                    tc_wait_st();
                    tc_fence_before();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(pv_ok_leader + b * 8);
            }
            // ---- epilogue ----
            mbar_wait_relaxed(stats_full, item_cnt & 1);
            mbar_wait_relaxed(o_final, item_cnt & 1);
            tc_fence_after();
            // the two warpgroups' partial sums, each relative to its own last reference maximum (log2 units)
            float mlog2 = sMax[row];
#pragma unroll
            for (int gg = 1; gg < kGroups; ++gg) mlog2 = fmaxf(mlog2, sMax[gg * 128 + row]);
            float dsum = 0.f;
#pragma unroll
            for (int gg = 0; gg < kGroups; ++gg) dsum += sSum[gg * 128 + row] * ex2_approx(sMax[gg * 128 + row] - mlog2);
            __syncwarp();
            if (lane == 0) mbar_arrive(stats_free);
            const float inv = 1.f / dsum;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (warp == R::kCorrWarp0 && lane == 0) tma_store_wait_read<0>();
                named_bar_sync(kBarEpilogue, 128);
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const int ch = hf * 2 + c2;
                    float orr[32];
                    tmem_ld_x32(tmem_o + lane_addr + ch * 32, orr);
                    tc_wait_ld();
                    if (ch == 3) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(pv_ok_leader);   // O is in registers: the next item's PV(0) may run
                    }
                    uint8_t* srow = sO + row * 128;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 val;
                        val.x = pack2<kBf16>(orr[8 * i + 0] * inv, orr[8 * i + 1] * inv);
                        val.y = pack2<kBf16>(orr[8 * i + 2] * inv, orr[8 * i + 3] * inv);
                        val.z = pack2<kBf16>(orr[8 * i + 4] * inv, orr[8 * i + 5] * inv);
                        val.w = pack2<kBf16>(orr[8 * i + 6] * inv, orr[8 * i + 7] * inv);
                        const int chunk = c2 * 4 + i;
                        *reinterpret_cast<uint4*>(srow + ((chunk ^ (row & 7)) << 4)) = val;
                    }
                }
                fence_proxy_async();
                named_bar_sync(kBarEpilogue, 128);
                if (warp == R::kCorrWarp0 && lane == 0) {
                    if (peers.n == 0) {
                        tma_store_4d(&map_o, sO, hf * 64, it.q0, it.h, it.b);
                    } else {
                        for (int r = 0; r < peers.n; ++r)
                            tma_store_4d(&peers.maps[peer_buf][r], sO, hf * 64, it.q0, it.h + peers.head_offset,
                                         it.b + peers.batch_offset);
                    }
                    tma_store_commit();
                }
            }
            if (p.lse != nullptr && it.q0 + row < it.nq)
                p.lse[it.b * p.lse_sb + it.h * p.lse_sh + it.q0 + row] = (mlog2 + log2f(dsum)) * kLn2;
#pragma unroll
            for (int bb = 0; bb < 3; ++bb)
                if (bb < it.n) pvd_base ^= (uint32_t)(((it.n - bb + 2) / 3) & 1) << bb;
        }
        if (warp == R::kCorrWarp0 && lane == 0) tma_store_wait_all<0>();
    } else {
        reg_dealloc<64>();
        if (warp == R::kMmaWarp && cta_rank == 0) {
            // =========================== MMA issuer (leader CTA) ===========================
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t kLboK = 1u << 16;
            constexpr uint32_t kLboV = (uint32_t)(kSubTileBytes >> 4) << 16;
            const uint32_t q_lo = ((smem_u32(sQ) >> 4) & 0x3FFFu) | kLboK;
            const uint32_t k_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboK;
            const uint32_t v_lo = ((smem_u32(sKV) >> 4) & 0x3FFFu) | kLboV;
            uint32_t kv_cnt = 0, item_par = 0, pv_par = 0;
            auto commit = [&](uint64_t* bar) { umma2_commit_multicast(bar, (uint16_t)3); };
            auto wait_entry = [&](uint32_t e) { mbar_wait(&kv_full[e % kRing], (e / kRing) & 1); };
            auto issue_S = [&](int buf, uint32_t e) {
                // this CTA's entry: keys [64 rank, +64) of the tile as two [64 rows][64 el] sub-tiles
                const uint32_t ka = k_lo + (e % kRing) * (kEntryBytes >> 4);
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks) {
                    const uint32_t qoff = ((ks >> 2) * kSubTileBytes + (ks & 3) * 32) >> 4;
                    const uint32_t koff = ((ks >> 2) * (kEntryBytes / 2) + (ks & 3) * 32) >> 4;
                    umma2_ss_lohi(tmem_base + buf * 128, q_lo + qoff, ka + koff, kDescHi, kIdescS, ks > 0 ? 1u : 0u);
                }
            };
            auto issue_PV = [&](int j, uint32_t e) {
                // this CTA's entry: head_dim columns [64 rank, +64) of the tile's 128 keys (one sub-tile)
                const uint32_t va = v_lo + (e % kRing) * (kEntryBytes >> 4);
#pragma unroll
                for (int ks = 0; ks < kBN / 16; ++ks)
                    umma2_ts_lohi(tmem_o, tmem_base + (j % 3) * 128 + 64 + ks * 8, va + ks * (2048 >> 4), kDescHi, kIdescO,
                                  (j > 0 || ks > 0) ? 1u : 0u);
            };
            for (int rnd = 0, w; (w = item_of_round(rnd, p)) >= 0; ++rnd, item_par ^= 1) {
                const WideItem it = decode_wide(w, p);
                const int n = it.n;
                mbar_wait(q_full, item_par);
                const int n_pro = n < 3 ? n : 3;
                for (int i = 0; i < n_pro; ++i) {
                    const uint32_t e = kv_cnt++;
                    wait_entry(e);
                    tc_fence_after();
                    if (elect_one()) {
                        issue_S(i, e);
                        commit(&s_full[i]);
                        commit(&kv_empty[e % kRing]);
                        if (i == n - 1) commit(q_empty);
                    }
                    __syncwarp();
                }
                for (int j = 0; j < n; ++j) {
                    const int b = j % 3;
                    const uint32_t ev = kv_cnt++;
                    wait_entry(ev);
                    const bool more = j + 3 < n;
                    uint32_t ek = 0;
                    if (more) {
                        ek = kv_cnt++;
                        wait_entry(ek);
                    }
                    mbar_wait(&pv_ok[b], (pv_par >> b) & 1);
                    pv_par ^= 1u << b;
                    tc_fence_after();
                    if (elect_one()) {
                        issue_PV(j, ev);
                        commit(&pv_done[b]);
                        if (j == n - 1) commit(o_final);
                        commit(&kv_empty[ev % kRing]);
                        if (more) {
                            issue_S(b, ek);
                            commit(&s_full[b]);
                            commit(&kv_empty[ek % kRing]);
                            if (j + 3 == n - 1) commit(q_empty);
                        }
                    }
                    __syncwarp();
                }
            }
        } else if (warp == R::kTmaWarp && lane == 0) {
            // =========================== TMA producer ===========================
            prefetch_tensormap(&map_q);
            prefetch_tensormap(&map_k);
            prefetch_tensormap(&map_v);
            prefetch_tensormap(&map_o);
            const int rank = (int)cta_rank;
            const uint32_t q_full_leader = mapa_u32(smem_u32(q_full), 0);
            const uint32_t kv_full_leader = mapa_u32(smem_u32(kv_full), 0);
            uint32_t kv_cnt = 0, item_par = 0;
            for (int rnd = 0, w; (w = item_of_round(rnd, p)) >= 0; ++rnd, item_par ^= 1) {
                const WideItem it = decode_wide(w, p);
                auto acquire = [&]() -> uint32_t {
                    const uint32_t e = kv_cnt++;
                    mbar_wait_relaxed(&kv_empty[e % kRing], ((e / kRing) & 1) ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(&kv_full[e % kRing], 2 * kEntryBytes);
                    return e % kRing;
                };
                auto load_k = [&](int i) {          // keys [128 i + 64 rank, +64): map_k carries 64-row boxes
                    const uint32_t s = acquire();
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf)
                        tma_load_4d_pair(sKV + s * kEntryBytes + hf * (kEntryBytes / 2), &map_k, kv_full_leader + s * 8, hf * 64,
                                         i * kBN + rank * 64, it.hk, it.b);
                };
                auto load_v = [&](int j) {          // head_dim columns [64 rank, +64) of keys [128 j, +128): 128-row boxes
                    const uint32_t s = acquire();
                    tma_load_4d_pair(sKV + s * kEntryBytes, &map_v, kv_full_leader + s * 8, rank * 64, j * kBN, it.hk, it.b);
                };
                mbar_wait_relaxed(q_empty, item_par ^ 1);
                if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_4d_pair(sQ + hf * kSubTileBytes, &map_q, q_full_leader, hf * 64, it.q0, it.h, it.b);
                const int n_pro = it.n < 3 ? it.n : 3;
                for (int i = 0; i < n_pro; ++i) load_k(i);
                for (int j = 0; j < it.n; ++j) {
                    load_v(j);
                    if (j + 3 < it.n) load_k(j + 3);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair(tmem_base, 512);
}

#endif  // PLI_TUNING (wide kernel)

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
#if PLI_SMEM_KEEP_SPACE
    // an offset added to the __shared__ array (not a pointer rebuilt from an integer) keeps the address space: the scalar
    // traffic through shared memory (scale factors, row statistics) compiles to LDS / STS instead of generic LD / ST
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#else
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
#endif
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

// Tuning builds only: environment switches for A/B measurements (see PLI_TUNING above).
#if PLI_TUNING
bool env_is(const char* name, char value) {
    const char* e = getenv(name);
    return e && e[0] == value;
}
bool wide_env() { static const bool on = env_is("PLI_WIDE", '1'); return on; }               // the one-tile, 128-key-step kernel
bool pair_mma_off_env() { static const bool off = env_is("PLI_PAIR_MMA", '0'); return off; }  // per-CTA MMAs + TMA multicast
bool cluster_mode_enabled() { static const bool off = env_is("PLI_NO_CLUSTER", '1'); return !off; }
int max_ctas_env() {                                    // PLI_MAX_CTAS: experiments on a part of the chip
    static const int n = [] { const char* e = getenv("PLI_MAX_CTAS"); return e ? atoi(e) : 0; }();
    return n;
}
#else
constexpr bool cluster_mode_enabled() { return true; }
constexpr int max_ctas_env() { return 0; }
#endif

#if PLI_TUNING
unsigned long long* g_trace_buf = nullptr;   // process-wide, unsynchronised: tuning builds are driven from one thread
int g_trace_cap = 0;
int g_debug_flags = 0;
#endif

int make_map_4d(CUtensorMap* map, const void* base, int dtype, int D, int N, int H, int B, const int64_t* st,
                int box_rows = 128, int box_cols = 64) {
    EncodeTiledFn enc = get_encode_tiled();
    if (enc == nullptr) return set_error(PLI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const CUtensorMapDataType dt = dtype == PLI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    // dims fastest first: head_dim, token, head, batch; a broadcast (zero) stride on a size-1 dim is replaced
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
    // TMA cannot broadcast: a zero (expanded) stride is only acceptable on a dimension of size 1, where it is never used
    const int sizes[3] = {B, H, N};
    for (int i = 0; i < 3; ++i)
        if (st[i] <= 0 && sizes[i] > 1)
            return set_error(PLI_ERR_UNSUPPORTED, "stride %lld on a dimension of size %d cannot be described by a tensor map",
                             (long long)st[i], sizes[i]);
    auto fix = [&](int64_t s) -> cuuint64_t { return (cuuint64_t)(s > 0 ? s : (int64_t)D) * 2; };
    cuuint64_t strides[3] = {fix(st[2]), fix(st[1]), fix(st[0])};
    cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PLI_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PLI_OK;
}

const PeerMaps& no_peers() {
    static const PeerMaps none = [] {
        PeerMaps m;
        memset(&m, 0, sizeof(m));
        return m;
    }();
    return none;
}

template <int kD, bool kBf16, int kCluster, bool kPaged = false, bool kPairMma = false>
int launch_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo,
             const PrefillParams& p, cudaStream_t stream, const PeerMaps& peers = no_peers()) {
    auto kern = prefill_tcgen05_kernel<kD, kBf16, kCluster, kPaged, kPairMma>;
    const int smem = SmemLayout<kD>::kTotal + 1024;
    PLI_CUDA_CHECK(ensure_dynamic_smem(kern, smem));
    PLI_CUDA_CHECK(bind_status_symbol());
    int grid = sm_count();
    if (grid <= 0) grid = 148;
    if (max_ctas_env() > 0 && grid > max_ctas_env()) grid = max_ctas_env();
    if (grid > p.total_items) grid = p.total_items;
    if (kCluster > 1) grid -= grid % kCluster;          // total_items is a multiple of kCluster in this mode
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PLI_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, mo, p, peers));
    count_launch();
    return PLI_OK;
}

#if PLI_TUNING
template <bool kBf16, int kGroups>
int launch_wide(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo,
                const PrefillParams& p, cudaStream_t stream, const PeerMaps& peers) {
    auto kern = prefill_wide_kernel<kBf16, kGroups>;
    const int smem = WideLayout::kTotal + 1024;
    PLI_CUDA_CHECK(ensure_dynamic_smem(kern, smem));
    PLI_CUDA_CHECK(bind_status_symbol());
    int grid = sm_count();
    if (grid <= 0) grid = 148;
    if (max_ctas_env() > 0 && grid > max_ctas_env()) grid = max_ctas_env();
    if (grid > p.total_items) grid = p.total_items;
    grid -= grid % 2;                                   // total_items is even (even group size)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(WideRoles<kGroups>::kThreadsWide);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PLI_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, mo, p, peers));
    count_launch();
    return PLI_OK;
}

#endif

// 5-D map over a paged pool (num_pages, layers, page_size, Hkv, D): dims fastest first d, head, slot, layer, page
int make_pool_map(CUtensorMap* map, const void* base, int dtype, int D, int Hkv, int page_size, int layers, int64_t pages,
                  const int64_t* st, int box_rows) {
    EncodeTiledFn enc = get_encode_tiled();
    if (enc == nullptr) return set_error(PLI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const CUtensorMapDataType dt = dtype == PLI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t dims[5] = {(cuuint64_t)D, (cuuint64_t)Hkv, (cuuint64_t)page_size, (cuuint64_t)layers, (cuuint64_t)pages};
    const int64_t layer_stride = st[1] > 0 ? st[1] : st[0];
    cuuint64_t strides[4] = {(cuuint64_t)st[3] * 2, (cuuint64_t)st[2] * 2, (cuuint64_t)layer_stride * 2, (cuuint64_t)st[0] * 2};
    cuuint32_t box[5] = {64, 1, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, dt, 5, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PLI_ERR_CUDA, "cuTensorMapEncodeTiled(pool) failed with CUresult %d", (int)r);
    return PLI_OK;
}

void fill_schedule(PrefillParams& p, int B, int Hq, int Hkv, int Nq) {
    const int group = Hq / Hkv;
    p.head_pairs = group % 2 == 0 ? 1 : 0;
    p.num_pairs = p.head_pairs ? (Nq + kBM - 1) / kBM : (Nq + 2 * kBM - 1) / (2 * kBM);
    const int64_t total = (int64_t)B * (p.head_pairs ? Hq / 2 : Hq) * p.num_pairs;
    p.total_items = total > 0x7fffffff ? -1 : (int)total;
    // Scheduling block (decode_item), in 128- or 256-row slots.  Inside a block the items are KV-group-major, so the larger
    // the block, the fewer KV groups the ~148 items in flight touch and the more of the K/V re-reads stay in L2 (C2, ncu:
    // 809 MB of DRAM reads with blocks of 8 slots, 593 with 16, 405 = the algorithmic 403 with one block of 64); but the
    // schedule is longest-first only ACROSS blocks, and one block per launch costs 2 % in load balance.  So: the smallest
    // power of two that gives every CTA ~12 items per block, at least the round-1 value (8 / 4 slots), at most half of
    // the slots (two blocks).  C2: 32 slots, 1.05x the algorithmic DRAM traffic at an unchanged time.
    int grid = sm_count() > 0 ? sm_count() : 148;
    p.pair_block = (total >= (int64_t)16 * grid ? 4 : 2) * (p.head_pairs ? 2 : 1);
    {
        const int64_t per_slot = (int64_t)B * Hkv * (p.head_pairs ? group / 2 : group);
        while ((int64_t)p.pair_block * per_slot < (int64_t)12 * grid && p.pair_block * 2 <= p.num_pairs / 2) p.pair_block *= 2;
    }
#ifdef PLI_PAIR_BLOCK
    p.pair_block = PLI_PAIR_BLOCK;                                        // A/B builds: scheduling block in row slots
#endif
    if (p.pair_block > p.num_pairs) p.pair_block = p.num_pairs;
    {
        const int gi = p.head_pairs ? group / 2 : group;                  // items per (KV group, slot)
        const int per_slot = B * Hkv * gi;
        p.group = group;
        p.cnt_last = p.num_pairs % p.pair_block ? p.num_pairs % p.pair_block : p.pair_block;
        p.fd_block = make_fastdiv((uint32_t)(p.pair_block * per_slot));
        p.fd_group_full = make_fastdiv((uint32_t)(p.pair_block * gi));
        p.fd_group_last = make_fastdiv((uint32_t)(p.cnt_last * gi));
        p.fd_gi = make_fastdiv((uint32_t)gi);
        p.fd_hkv = make_fastdiv((uint32_t)Hkv);
    }
#if PLI_TUNING
    p.trace = g_trace_buf;
    p.trace_cap = g_trace_cap;
    p.debug_flags = g_debug_flags;
#else
    p.trace = nullptr;
    p.trace_cap = 0;
    p.debug_flags = 0;
#endif
    p.table = nullptr;
    p.seq_lens = nullptr;
    p.table_stride = p.page_size = p.page_shift = p.layer = p.box_rows = p.box_rows_v = 0;
    p.cu_q = nullptr;
    p.o_base = nullptr;
    p.o_st_tok = p.o_st_head = p.o_st_batch = 0;
    p.lse_sb = (int64_t)Hq * Nq;
    p.lse_sh = Nq;
}

}  // namespace

// cu_seqlens_q == nullptr: q / o are (B, Hq, Nq, D) with strides {batch, head, token}.  Otherwise they are packed
// (total_q, Hq, D) with strides {unused, head, token}, Nq is the upper bound of the per-sequence query lengths.
int launch_prefill_tcgen05_paged(const void* q, const void* k_pool, const void* v_pool, const int32_t* block_table,
                                 const int32_t* seq_lens, const int32_t* cu_seqlens_q, int64_t total_q, void* o, float* lse,
                                 int B, int Hq, int Hkv, int Nq, int D, int max_seq_len, int block_size, int table_stride,
                                 int layer, int64_t num_pages, const int64_t* qs, const int64_t* kvs, const int64_t* os,
                                 float scale, int dtype, cudaStream_t stream) {
    const bool pairs = (Hq / Hkv) % 4 == 0 && cluster_mode_enabled();
    // CTA-pair MMAs as in the contiguous kernel (D = 128): K boxes of at most 32 keys, V boxes of up to 128
    const bool pair_mma = pairs && D == 128 && kPagedPairMma;
    const int rows_cta = pair_mma ? 32 : pairs ? kHN : kBN;
    const int box_rows = block_size < rows_cta ? block_size : rows_cta;
    const int box_rows_v = pair_mma ? (block_size < kBN ? block_size : kBN) : box_rows;
    CUtensorMap mq, mk, mv, mo;
    int rc;
    const bool packed = cu_seqlens_q != nullptr;
    if ((rc = make_map_4d(&mq, q, dtype, D, packed ? (int)total_q : Nq, Hq, packed ? 1 : B, qs))) return rc;
    if ((rc = make_pool_map(&mk, k_pool, dtype, D, Hkv, block_size, layer + 1, num_pages, kvs, box_rows))) return rc;
    if ((rc = make_pool_map(&mv, v_pool, dtype, D, Hkv, block_size, layer + 1, num_pages, kvs, box_rows_v))) return rc;
    if ((rc = make_map_4d(&mo, o, dtype, D, packed ? (int)total_q : Nq, Hq, packed ? 1 : B, os))) return rc;
    PeerMaps pm = no_peers();
    if ((rc = make_map_4d(&pm.o32, o, dtype, D, packed ? (int)total_q : Nq, Hq, packed ? 1 : B, os, 32, 32))) return rc;
    pm.warp_store = 1;
    PrefillParams p;
    p.lse = lse;
    p.B = B;
    p.Hq = Hq;
    p.Hkv = Hkv;
    p.Nq = Nq;
    p.Nk = max_seq_len;
    p.causal = 1;
    p.scale = scale;
    p.scale_log2 = scale * kLog2e;
    fill_schedule(p, B, Hq, Hkv, Nq);
    if (p.total_items < 0) return set_error(PLI_ERR_UNSUPPORTED, "too many work items");
    p.table = block_table;
    p.seq_lens = seq_lens;
    p.table_stride = table_stride;
    p.page_size = block_size;
    p.page_shift = block_size == 16 ? 4 : block_size == 32 ? 5 : block_size == 64 ? 6 : 7;
    p.layer = layer;
    p.box_rows = box_rows;
    p.box_rows_v = box_rows_v;
    p.o_base = static_cast<uint8_t*>(o);
    p.o_st_batch = packed ? 0 : os[0];
    p.o_st_head = os[1];
    p.o_st_tok = os[2];
    if (packed) {
        p.cu_q = cu_seqlens_q;
        p.lse_sb = 0;               // lse is (Hq, total_q)
        p.lse_sh = total_q;
    }
    const bool bf16 = dtype == PLI_BF16;
#define PLI_GO(DD, BF)                                                                                  \
    return pairs ? launch_t<DD, BF, 2, true>(mq, mk, mv, mo, p, stream, pm) : launch_t<DD, BF, 1, true>(mq, mk, mv, mo, p, stream, pm)
    if (D == 128) {
        if (pair_mma) {
            if (bf16) return launch_t<128, true, 2, true, true>(mq, mk, mv, mo, p, stream, pm);
            return launch_t<128, false, 2, true, true>(mq, mk, mv, mo, p, stream, pm);
        }
        if (bf16) PLI_GO(128, true);
        PLI_GO(128, false);
    }
    if (bf16) PLI_GO(64, true);
    PLI_GO(64, false);
#undef PLI_GO
}

int launch_prefill_tcgen05(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv,
                           int Nq, int Nk, int D, const int64_t* qs, const int64_t* ks, const int64_t* vs,
                           const int64_t* os, float scale, int causal, int dtype, cudaStream_t stream,
                           const PrefillPeerInfo* peer) {
    CUtensorMap mq, mk, mv, mo;
    int rc;
    PeerMaps pm = no_peers();
    if (peer != nullptr) {
        // os = strides of the FULL output {batch, head, token}; o = this rank's buffer 0 (only used for the local map)
        for (int buf = 0; buf < 2; ++buf)
            for (int r = 0; r < peer->n; ++r) {
                const char* base = static_cast<const char*>(peer->o[r]) + (size_t)buf * peer->buffer_stride * 2;
                if ((rc = make_map_4d(&pm.maps[buf][r], base, dtype, D, Nq, peer->Hq_total, peer->B_total, os))) return rc;
            }
        pm.epoch = peer->epoch;
        pm.n = peer->n;
        pm.head_offset = peer->head_offset;
        pm.batch_offset = peer->batch_offset;
    }
    if ((rc = make_map_4d(&mq, q, dtype, D, Nq, Hq, B, qs))) return rc;
    // CTA pairs share K/V when consecutive items are two q heads of one KV group (even group size): each CTA
    // then loads 64-row halves of the K/V tiles and multicasts them
    // (even, odd) items share their K/V when the per-group item count is even: group size divisible by 4 with
    // head-pair items (group size odd -> row-pair items of single heads -> never)
    const bool pairs = (Hq / Hkv) % 4 == 0 && cluster_mode_enabled();
    // CTA-pair MMAs (D = 128, cta_group::2): each CTA loads 32-key boxes of K and full-height, 64-column boxes of V (its
    // half of B).  The default whenever a pair of CTAs shares its K/V: with the correction warps sharing the softmax
    // work the tensor side matters again, and the pair instructions run the kernel's MMA sequence in 1183 instead of
    // 1283 cycles: measured +2 % on C2 (1298-1306 against 1274-1283 TFLOP/s, same box), bit-identical.
    bool pair_mma = pairs && D == 128;
#if PLI_TUNING
    if (pair_mma_off_env() || (g_debug_flags & 64)) pair_mma = false;      // flags bit 6: per-CTA MMAs + TMA multicast
    // one Q tile per CTA, 128-key S tiles, pair MMAs (any even group size): opt-in experiment
    const bool wide = D == 128 && (Hq / Hkv) % 2 == 0 && cluster_mode_enabled() && (wide_env() || (g_debug_flags & 4));
    if (wide) {
        if ((rc = make_map_4d(&mk, k, dtype, D, Nk, Hkv, B, ks, kHN))) return rc;
        if ((rc = make_map_4d(&mv, v, dtype, D, Nk, Hkv, B, vs, kBN))) return rc;
        if (peer != nullptr) mo = pm.maps[0][0];
        else if ((rc = make_map_4d(&mo, o, dtype, D, Nq, Hq, B, os))) return rc;
        PrefillParams p;
        p.lse = lse;
        p.B = B;
        p.Hq = Hq;
        p.Hkv = Hkv;
        p.Nq = Nq;
        p.Nk = Nk;
        p.causal = causal;
        p.scale = scale;
        p.scale_log2 = scale * kLog2e;
        fill_schedule(p, B, Hq, Hkv, Nq);
        p.head_pairs = 0;
        p.num_pairs = (Nq + kBM - 1) / kBM;                       // 128-row slots; one item per (q head, slot)
        const int64_t total = (int64_t)B * Hq * p.num_pairs;
        if (total > 0x7fffffff) return set_error(PLI_ERR_UNSUPPORTED, "too many work items");
        p.total_items = (int)total;
        const int sms = sm_count() > 0 ? sm_count() : 148;
        p.pair_block = total >= (int64_t)32 * sms ? 8 : 4;
        if (p.pair_block > p.num_pairs) p.pair_block = p.num_pairs;
        if (g_debug_flags & 32)              // two softmax warpgroups (512 threads) instead of three
            return dtype == PLI_BF16 ? launch_wide<true, 2>(mq, mk, mv, mo, p, stream, pm)
                                     : launch_wide<false, 2>(mq, mk, mv, mo, p, stream, pm);
        return dtype == PLI_BF16 ? launch_wide<true, 3>(mq, mk, mv, mo, p, stream, pm)
                                 : launch_wide<false, 3>(mq, mk, mv, mo, p, stream, pm);
    }
#endif  // PLI_TUNING
    if ((rc = make_map_4d(&mk, k, dtype, D, Nk, Hkv, B, ks, pair_mma ? kHN / 2 : pairs ? kHN : kBN))) return rc;
    if ((rc = make_map_4d(&mv, v, dtype, D, Nk, Hkv, B, vs, pair_mma ? kBN : pairs ? kHN : kBN))) return rc;
    if (peer != nullptr) mo = pm.maps[0][0];
    else if ((rc = make_map_4d(&mo, o, dtype, D, Nq, Hq, B, os))) return rc;
    if (peer == nullptr) {
        if ((rc = make_map_4d(&pm.o32, o, dtype, D, Nq, Hq, B, os, 32, 32))) return rc;
        pm.warp_store = 1;
    }
    PrefillParams p;
    p.lse = lse;
    p.B = B;
    p.Hq = Hq;
    p.Hkv = Hkv;
    p.Nq = Nq;
    p.Nk = Nk;
    p.causal = causal;
    p.scale = scale;
    p.scale_log2 = scale * kLog2e;
    fill_schedule(p, B, Hq, Hkv, Nq);
    if (p.total_items < 0) return set_error(PLI_ERR_UNSUPPORTED, "too many work items");
    if (peer == nullptr) {                    // raw output address for the register -> global epilogue
        p.o_base = static_cast<uint8_t*>(o);
        p.o_st_batch = os[0];
        p.o_st_head = os[1];
        p.o_st_tok = os[2];
    }
    const bool bf16 = dtype == PLI_BF16;
#define PLI_GO(DD, BF)                                                                            \
    return pairs ? launch_t<DD, BF, 2>(mq, mk, mv, mo, p, stream, pm) : launch_t<DD, BF, 1>(mq, mk, mv, mo, p, stream, pm)
    if (D == 128) {
        if (pair_mma) {
            if (bf16) return launch_t<128, true, 2, false, true>(mq, mk, mv, mo, p, stream, pm);
            return launch_t<128, false, 2, false, true>(mq, mk, mv, mo, p, stream, pm);
        }
        if (bf16) PLI_GO(128, true);
        PLI_GO(128, false);
    }
    if (bf16) PLI_GO(64, true);
    PLI_GO(64, false);
#undef PLI_GO
}

cudaError_t bind_status_prefill() { return bind_status_symbol(); }

}  // namespace pli

using namespace pli;

// Debug aid: record CTA 0's pipeline timeline of the next prefill launches into `buf` (device memory of
// 5 regions x capacity x 16 bytes, zeroed by the caller); buf = NULL switches it off.  flags (tuning builds): bit 2 wide kernel (bit 5: with two softmax
// warpgroups), bit 6 per-CTA MMAs instead of CTA-pair MMAs; formerly bit 1: CTA-pair MMAs
// (tcgen05.mma.cta_group::2) for the following launches (tuning / tests).
extern "C" int pli_debug_prefill_trace(void* buf, int capacity, int flags) {
#if PLI_TUNING
    g_trace_buf = static_cast<unsigned long long*>(buf);
    g_trace_cap = buf ? capacity : 0;
    g_debug_flags = flags;
    return PLI_OK;
#else
    (void)buf; (void)capacity;
    if (flags == 0) return PLI_OK;           // "switch everything off" is always satisfiable
    return set_error(PLI_ERR_UNSUPPORTED, "this is the product build: kernel-selection flags and the timeline exist only in "
                                          "tuning builds (python -m physics_llm_inference_b200.build --variant=tuning -DPLI_TUNING=1)");
#endif
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

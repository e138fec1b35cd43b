// pli_capi.cu — C-ABI entry points (include/pli_attention.h): argument validation, kernel
// selection, error text.  Kernels live in prefill_tcgen05.cu, prefill_simt.cu and decode.cu.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace pli {

static thread_local char g_err[512] = "";
static thread_local uint64_t g_launches = 0;

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches += (uint64_t)n; }

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// Fault record in host-mapped pinned memory (see common.cuh): readable by the host even after a kernel trapped.
unsigned long long* status_words() {
    static unsigned long long* words = []() -> unsigned long long* {
        void* p = nullptr;
        if (cudaHostAlloc(&p, 8 * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        memset(p, 0, 8 * sizeof(unsigned long long));
        return static_cast<unsigned long long*>(p);
    }();
    return words;
}
static uint64_t g_peer_timeout_ns = 60ull * 1000 * 1000 * 1000;      // 0 = wait forever
uint64_t peer_timeout_ns() { return g_peer_timeout_ns; }

static thread_local int g_device = -1;
int current_device() {
    if (g_device < 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) g_device = dev;
    }
    return g_device;
}

cudaError_t ensure_dynamic_smem_impl(const void* kern, int bytes) {
    struct Entry { const void* fn; uint64_t devices; };
    static thread_local Entry table[64] = {};
    const int dev = current_device();
    Entry* slot = nullptr;
    for (Entry& e : table) {
        if (e.fn == kern || e.fn == nullptr) { slot = &e; break; }
    }
    if (slot && slot->fn == kern && dev >= 0 && dev < 64 && ((slot->devices >> dev) & 1)) return cudaSuccess;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (err == cudaSuccess && slot && dev >= 0 && dev < 64) {
        slot->fn = kern;
        slot->devices |= 1ull << dev;
    }
    return err;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
        cached = prop.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

static int device_is_sm100() {
    static thread_local int cached_dev = -1, major = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
        cached_dev = dev;
    }
    return major == 10;
}

static bool tcgen05_eligible(int D, int dtype, const int64_t* qs, const int64_t* ks, const int64_t* vs,
                             const int64_t* os, const void* q, const void* k, const void* v, const void* o) {
    if (dtype != PLI_BF16 && dtype != PLI_F16) return false;
    if (D != 64 && D != 128) return false;
    const uintptr_t ptrs = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                           reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o);
    if (ptrs & 15) return false;
    for (int i = 0; i < 3; ++i) {
        if (qs[i] % 8 || ks[i] % 8 || vs[i] % 8 || os[i] % 8) return false;
        if (qs[i] < 0 || ks[i] < 0 || vs[i] < 0 || os[i] < 0) return false;
    }
    // every stride must be real: TMA cannot broadcast, so expanded (zero-stride) views such as k.expand(B, ...) or an
    // MQA k[:, :1].expand(-1, H, -1, -1) go to the SIMT kernel, which indexes with the strides as given
    for (int i = 0; i < 3; ++i)
        if (qs[i] == 0 || ks[i] == 0 || vs[i] == 0 || os[i] == 0) return false;
    return true;
}

}  // namespace pli

using namespace pli;

extern "C" int pli_abi_version(void) { return PLI_ABI_VERSION; }
extern "C" const char* pli_last_error(void) { return g_err; }
extern "C" int pli_set_device(int device) {
    PLI_CUDA_CHECK(cudaSetDevice(device));
    g_device = device;
    if (!device_is_sm100()) return set_error(PLI_ERR_DEVICE, "CUDA device %d is not sm_100 (B200)", device);
    PLI_CUDA_CHECK(bind_status_prefill());       // the fault record's device-side pointer, per translation unit
    PLI_CUDA_CHECK(bind_status_decode());
    return PLI_OK;
}
extern "C" int pli_device_status(uint64_t out[8], int clear) {
    unsigned long long* w = status_words();
    if (!out) return set_error(PLI_ERR_INVALID, "null output");
    for (int i = 0; i < 8; ++i) out[i] = w ? w[i] : 0;
    if (w && clear) memset(w, 0, 8 * sizeof(unsigned long long));
    return w ? PLI_OK : set_error(PLI_ERR_CUDA, "the host-mapped fault record could not be allocated");
}
extern "C" int pli_set_peer_timeout_ms(int64_t ms) {
    if (ms < 0) return set_error(PLI_ERR_INVALID, "negative timeout");
    g_peer_timeout_ns = (uint64_t)ms * 1000000ull;
    return PLI_OK;
}
extern "C" uint64_t pli_launch_count(void) { return g_launches; }
extern "C" void pli_reset_launch_count(void) { g_launches = 0; }

extern "C" int pli_prefill_kernel_kind(int D, int dtype, const int64_t q_strides[3], const int64_t k_strides[3],
                                       const int64_t v_strides[3], const int64_t o_strides[3], const void* q,
                                       const void* k, const void* v, const void* o) {
    if (dtype != PLI_BF16 && dtype != PLI_F16 && dtype != PLI_F32) return PLI_KIND_NONE;
    if (D < 1 || D > 256) return PLI_KIND_NONE;
    return tcgen05_eligible(D, dtype, q_strides, k_strides, v_strides, o_strides, q, k, v, o) ? PLI_KIND_TCGEN05
                                                                                              : PLI_KIND_SIMT;
}

extern "C" int pli_prefill_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv,
                               int Nq, int Nk, int D, const int64_t q_strides[3], const int64_t k_strides[3],
                               const int64_t v_strides[3], const int64_t o_strides[3], float scale, int causal,
                               int dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!q || !k || !v || !o) return set_error(PLI_ERR_INVALID, "null pointer argument");
    if (!q_strides || !k_strides || !v_strides || !o_strides) return set_error(PLI_ERR_INVALID, "null stride array");
    if (B <= 0 || Hq <= 0 || Hkv <= 0 || Nq <= 0 || Nk <= 0 || D <= 0)
        return set_error(PLI_ERR_INVALID, "non-positive dimension (B=%d Hq=%d Hkv=%d Nq=%d Nk=%d D=%d)", B, Hq, Hkv, Nq, Nk, D);
    if (Hq % Hkv != 0) return set_error(PLI_ERR_INVALID, "Hq (%d) must be a multiple of Hkv (%d)", Hq, Hkv);
    if (causal && Nq > Nk)
        return set_error(PLI_ERR_INVALID, "causal attention needs Nq <= Nk (got %d > %d): rows without a visible key", Nq, Nk);
    if (dtype != PLI_BF16 && dtype != PLI_F16 && dtype != PLI_F32) return set_error(PLI_ERR_INVALID, "unknown dtype %d", dtype);
    if (!device_is_sm100()) return set_error(PLI_ERR_DEVICE, "current CUDA device is not sm_100 (B200)");
    if (scale > 0.f && tcgen05_eligible(D, dtype, q_strides, k_strides, v_strides, o_strides, q, k, v, o))
        return launch_prefill_tcgen05(q, k, v, o, lse, B, Hq, Hkv, Nq, Nk, D, q_strides, k_strides, v_strides, o_strides,
                                      scale, causal, dtype, stream);
    return launch_prefill_simt(q, k, v, o, lse, B, Hq, Hkv, Nq, Nk, D, q_strides, k_strides, v_strides, o_strides, scale,
                               causal, dtype, stream);
}

extern "C" int pli_prefill_fwd_scatter(const void* q, const void* k, const void* v, float* lse, int B, int Hq, int Hkv,
                                       int Nq, int Nk, int D, const int64_t q_strides[3], const int64_t k_strides[3],
                                       const int64_t v_strides[3], const int64_t o_strides[3], float scale, int causal,
                                       int dtype, int B_total, int Hq_total, int batch_offset, int head_offset,
                                       const pli_peer_scatter* ps, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!q || !k || !v || !ps) return set_error(PLI_ERR_INVALID, "null pointer argument");
    if (!q_strides || !k_strides || !v_strides || !o_strides) return set_error(PLI_ERR_INVALID, "null stride array");
    if (B <= 0 || Hq <= 0 || Hkv <= 0 || Nq <= 0 || Nk <= 0 || D <= 0)
        return set_error(PLI_ERR_INVALID, "non-positive dimension (B=%d Hq=%d Hkv=%d Nq=%d Nk=%d D=%d)", B, Hq, Hkv, Nq, Nk, D);
    if (Hq % Hkv != 0) return set_error(PLI_ERR_INVALID, "Hq (%d) must be a multiple of Hkv (%d)", Hq, Hkv);
    if (causal && Nq > Nk) return set_error(PLI_ERR_INVALID, "causal attention needs Nq <= Nk (got %d > %d)", Nq, Nk);
    if (ps->n_peers < 1 || ps->n_peers > PLI_MAX_PEERS) return set_error(PLI_ERR_INVALID, "n_peers must be in [1, %d]", PLI_MAX_PEERS);
    if (ps->rank < 0 || ps->rank >= ps->n_peers) return set_error(PLI_ERR_INVALID, "rank outside [0, n_peers)");
    if (!ps->epoch) return set_error(PLI_ERR_INVALID, "null epoch word");
    if (batch_offset < 0 || head_offset < 0 || batch_offset + B > B_total || head_offset + Hq > Hq_total)
        return set_error(PLI_ERR_INVALID, "the local (batch, head) slice does not fit the full output");
    if (!device_is_sm100()) return set_error(PLI_ERR_DEVICE, "current CUDA device is not sm_100 (B200)");
    if (!(scale > 0.f)) return set_error(PLI_ERR_UNSUPPORTED, "the scatter entry needs scale > 0");
    PrefillPeerInfo info{};
    for (int r = 0; r < ps->n_peers; ++r) {
        if (!ps->peer_o[r]) return set_error(PLI_ERR_INVALID, "null peer pointer for rank %d", r);
        if ((reinterpret_cast<uintptr_t>(ps->peer_o[r]) & 15) || (ps->buffer_stride % 8))
            return set_error(PLI_ERR_INVALID, "peer buffers must be 16-byte aligned");
        info.o[r] = ps->peer_o[r];
    }
    if (!tcgen05_eligible(D, dtype, q_strides, k_strides, v_strides, o_strides, q, k, v, ps->peer_o[ps->rank]))
        return set_error(PLI_ERR_UNSUPPORTED, "the scatter entry serves the tcgen05 kernel only (bf16/f16, head_dim 64/128, "
                                               "16-byte aligned, strides multiple of 8)");
    info.epoch = ps->epoch;
    info.buffer_stride = ps->buffer_stride;
    info.n = ps->n_peers;
    info.Hq_total = Hq_total;
    info.B_total = B_total;
    info.head_offset = head_offset;
    info.batch_offset = batch_offset;
    return launch_prefill_tcgen05(q, k, v, ps->peer_o[ps->rank], lse, B, Hq, Hkv, Nq, Nk, D, q_strides, k_strides,
                                  v_strides, o_strides, scale, causal, dtype, stream, &info);
}

extern "C" int pli_prefill_paged_fwd(const void* q, const void* k_pool, const void* v_pool, const int32_t* block_table,
                                     const int32_t* seq_lens, void* o, float* lse, int B, int Hq, int Hkv, int Nq, int D,
                                     int max_seq_len, int block_size, int table_stride, int layer, int64_t num_pages,
                                     const int64_t q_strides[3], const int64_t kv_strides[4], const int64_t o_strides[3],
                                     float scale, int dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!q || !k_pool || !v_pool || !block_table || !seq_lens || !o) return set_error(PLI_ERR_INVALID, "null pointer argument");
    if (!q_strides || !kv_strides || !o_strides) return set_error(PLI_ERR_INVALID, "null stride array");
    if (B <= 0 || Hq <= 0 || Hkv <= 0 || Nq <= 0 || D <= 0 || max_seq_len <= 0 || num_pages <= 0)
        return set_error(PLI_ERR_INVALID, "non-positive dimension");
    if (Hq % Hkv != 0) return set_error(PLI_ERR_INVALID, "Hq (%d) must be a multiple of Hkv (%d)", Hq, Hkv);
    if (layer < 0) return set_error(PLI_ERR_INVALID, "negative layer");
    if (!device_is_sm100()) return set_error(PLI_ERR_DEVICE, "current CUDA device is not sm_100 (B200)");
    if (dtype != PLI_BF16 && dtype != PLI_F16)
        return set_error(PLI_ERR_UNSUPPORTED, "paged prefill takes bf16/f16 (gather the pages and call pli_prefill_fwd for f32)");
    if (D != 64 && D != 128) return set_error(PLI_ERR_UNSUPPORTED, "paged prefill supports head_dim 64 and 128, got %d", D);
    if (block_size != 16 && block_size != 32 && block_size != 64 && block_size != 128)
        return set_error(PLI_ERR_UNSUPPORTED, "paged prefill supports page sizes 16, 32, 64, 128, got %d", block_size);
    if (!(scale > 0.f)) return set_error(PLI_ERR_UNSUPPORTED, "paged prefill needs scale > 0");
    const uintptr_t ptrs = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k_pool) |
                           reinterpret_cast<uintptr_t>(v_pool) | reinterpret_cast<uintptr_t>(o);
    if (ptrs & 15) return set_error(PLI_ERR_INVALID, "q, o and the pools must be 16-byte aligned");
    for (int i = 0; i < 3; ++i)
        if (q_strides[i] % 8 || o_strides[i] % 8 || q_strides[i] < 0 || o_strides[i] < 0)
            return set_error(PLI_ERR_INVALID, "q/o strides must be non-negative multiples of 8 elements");
    {
        const int sizes[3] = {B, Hq, Nq};
        for (int i = 0; i < 3; ++i)
            if (sizes[i] > 1 && (q_strides[i] == 0 || o_strides[i] == 0))
                return set_error(PLI_ERR_UNSUPPORTED, "paged prefill cannot read an expanded (zero-stride) q; make it contiguous");
    }
    for (int i = 0; i < 4; ++i)
        if (kv_strides[i] % 8 || kv_strides[i] < 0) return set_error(PLI_ERR_INVALID, "pool strides must be multiples of 8 elements");
    return launch_prefill_tcgen05_paged(q, k_pool, v_pool, block_table, seq_lens, nullptr, 0, o, lse, B, Hq, Hkv, Nq, D,
                                        max_seq_len, block_size, table_stride, layer, num_pages, q_strides, kv_strides,
                                        o_strides, scale, dtype, stream);
}

extern "C" int pli_prefill_varlen_paged_fwd(const void* q, const void* k_pool, const void* v_pool,
                                            const int32_t* block_table, const int32_t* seq_lens,
                                            const int32_t* cu_seqlens_q, void* o, float* lse, int B, int Hq, int Hkv,
                                            int64_t total_q, int max_q_len, int D, int max_seq_len, int block_size,
                                            int table_stride, int layer, int64_t num_pages, const int64_t q_strides[2],
                                            const int64_t kv_strides[4], const int64_t o_strides[2], float scale,
                                            int dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!q || !k_pool || !v_pool || !block_table || !seq_lens || !cu_seqlens_q || !o)
        return set_error(PLI_ERR_INVALID, "null pointer argument");
    if (!q_strides || !kv_strides || !o_strides) return set_error(PLI_ERR_INVALID, "null stride array");
    if (B <= 0 || Hq <= 0 || Hkv <= 0 || total_q <= 0 || max_q_len <= 0 || D <= 0 || max_seq_len <= 0 || num_pages <= 0)
        return set_error(PLI_ERR_INVALID, "non-positive dimension");
    if (total_q > 0x7fffffff) return set_error(PLI_ERR_UNSUPPORTED, "total_q exceeds 2^31-1");
    if (Hq % Hkv != 0) return set_error(PLI_ERR_INVALID, "Hq (%d) must be a multiple of Hkv (%d)", Hq, Hkv);
    if (layer < 0) return set_error(PLI_ERR_INVALID, "negative layer");
    if (!device_is_sm100()) return set_error(PLI_ERR_DEVICE, "current CUDA device is not sm_100 (B200)");
    if (dtype != PLI_BF16 && dtype != PLI_F16) return set_error(PLI_ERR_UNSUPPORTED, "varlen paged prefill takes bf16/f16");
    if (D != 64 && D != 128) return set_error(PLI_ERR_UNSUPPORTED, "varlen paged prefill supports head_dim 64 and 128, got %d", D);
    if (block_size != 16 && block_size != 32 && block_size != 64 && block_size != 128)
        return set_error(PLI_ERR_UNSUPPORTED, "varlen paged prefill supports page sizes 16, 32, 64, 128, got %d", block_size);
    if (!(scale > 0.f)) return set_error(PLI_ERR_UNSUPPORTED, "varlen paged prefill needs scale > 0");
    const uintptr_t ptrs = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k_pool) |
                           reinterpret_cast<uintptr_t>(v_pool) | reinterpret_cast<uintptr_t>(o);
    if (ptrs & 15) return set_error(PLI_ERR_INVALID, "q, o and the pools must be 16-byte aligned");
    for (int i = 0; i < 2; ++i)
        if (q_strides[i] % 8 || o_strides[i] % 8 || q_strides[i] <= 0 || o_strides[i] <= 0)
            return set_error(PLI_ERR_INVALID, "q/o strides must be positive multiples of 8 elements");
    for (int i = 0; i < 4; ++i)
        if (kv_strides[i] % 8 || kv_strides[i] < 0) return set_error(PLI_ERR_INVALID, "pool strides must be multiples of 8 elements");
    // the launcher takes {batch, head, token}: the packed tensors have no batch dimension
    const int64_t qs[3] = {0, q_strides[1], q_strides[0]}, os[3] = {0, o_strides[1], o_strides[0]};
    return launch_prefill_tcgen05_paged(q, k_pool, v_pool, block_table, seq_lens, cu_seqlens_q, total_q, o, lse, B, Hq, Hkv,
                                        max_q_len, D, max_seq_len, block_size, table_stride, layer, num_pages, qs, kv_strides,
                                        os, scale, dtype, stream);
}

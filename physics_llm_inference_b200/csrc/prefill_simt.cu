// prefill_simt.cu — CUDA-core attention forward, fp32 accumulate, any head_dim <= 256.
//
// Serves f32 inputs (parity at 1e-3 against the fp32 oracle) and bf16/f16
// problems whose head_dim the tcgen05 kernel does not cover.  It is the ch06 recurrence
// (ch06/flash_attention.py:38-72) with one CTA per 32 query rows: S = QK^T*scale (:55), running
// max / sum (:57-62), O accumulation (:64-65, normalisation deferred to the end — SURVEY D12),
// plus the ch01/ch02 causal rule and GQA head map.
#include "common.cuh"

namespace pli {
namespace {

constexpr int kBQ = 32;       // query rows per CTA (8 per warp)
constexpr int kBK = 32;       // keys per step (one per lane)
constexpr int kThreads = 128;

template <typename T, int kDC>  // kDC = ceil(D / 32)
__global__ void __launch_bounds__(kThreads) prefill_simt_kernel(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
    float* __restrict__ lse, int Hq, int Hkv, int Nq, int Nk, int D, int64_t qsb, int64_t qsh, int64_t qsn,
    int64_t ksb, int64_t ksh, int64_t ksn, int64_t vsb, int64_t vsh, int64_t vsn, int64_t osb, int64_t osh,
    int64_t osn, float scale, int causal) {
    extern __shared__ float smem[];
    float* Qs = smem;                  // [kBQ][D]
    float* Ks = Qs + kBQ * D;          // [kBK][D+1]
    float* Vs = Ks + kBK * (D + 1);    // [kBK][D]
    float* Ps = Vs + kBK * D;          // [kBQ][kBK+1]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kBQ;
    const int hk = h / (Hq / Hkv);     // ch01/gqa.py:30-31
    const int off = Nk - Nq;           // ch02/cached_generation.py:87-90
    const T* qp = q + b * qsb + h * qsh;
    const T* kp = k + b * ksb + hk * ksh;
    const T* vp = v + b * vsb + hk * vsh;

    for (int i = tid; i < kBQ * D; i += kThreads) {
        int r = i / D, d = i - r * D;
        Qs[i] = (q0 + r < Nq) ? to_f32<T>(qp[(int64_t)(q0 + r) * qsn + d]) : 0.f;
    }

    float m[8], dsum[8], acc[8][kDC];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        m[r] = -INFINITY;
        dsum[r] = 0.f;
#pragma unroll
        for (int c = 0; c < kDC; ++c) acc[r][c] = 0.f;
    }

    int kend = Nk;
    if (causal) kend = min(Nk, q0 + kBQ + off);
    for (int k0 = 0; k0 < kend; k0 += kBK) {
        __syncthreads();
        for (int i = tid; i < kBK * D; i += kThreads) {
            int r = i / D, d = i - r * D;
            bool ok = k0 + r < Nk;
            Ks[r * (D + 1) + d] = ok ? to_f32<T>(kp[(int64_t)(k0 + r) * ksn + d]) : 0.f;
            Vs[i] = ok ? to_f32<T>(vp[(int64_t)(k0 + r) * vsn + d]) : 0.f;
        }
        __syncthreads();

        float s[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) s[r] = 0.f;
        const float* krow = Ks + lane * (D + 1);
        const float* qrow = Qs + (warp * 8) * D;
        for (int d = 0; d < D; ++d) {
            float kv = krow[d];
#pragma unroll
            for (int r = 0; r < 8; ++r) s[r] = fmaf(qrow[r * D + d], kv, s[r]);
        }
        const int key = k0 + lane;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int row = q0 + warp * 8 + r;
            float sv = s[r] * scale;
            if (key >= Nk || (causal && key > row + off)) sv = -INFINITY;
            float mx = sv;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
            const float m_new = fmaxf(m[r], mx);
            const float p = (m_new == -INFINITY) ? 0.f : expf(sv - m_new);
            const float alpha = (m[r] == -INFINITY) ? 0.f : expf(m[r] - m_new);
            float ps = p;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o2);
            dsum[r] = dsum[r] * alpha + ps;
#pragma unroll
            for (int c = 0; c < kDC; ++c) acc[r][c] *= alpha;
            m[r] = m_new;
            Ps[(warp * 8 + r) * (kBK + 1) + lane] = p;
        }
        __syncwarp();
        for (int kk = 0; kk < kBK; ++kk) {
            float vv[kDC];
#pragma unroll
            for (int c = 0; c < kDC; ++c) vv[c] = (lane + 32 * c < D) ? Vs[kk * D + lane + 32 * c] : 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float pr = Ps[(warp * 8 + r) * (kBK + 1) + kk];
#pragma unroll
                for (int c = 0; c < kDC; ++c) acc[r][c] = fmaf(pr, vv[c], acc[r][c]);
            }
        }
    }

#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = q0 + warp * 8 + r;
        if (row >= Nq) continue;
        const float inv = 1.f / dsum[r];
        T* op = o + b * osb + h * osh + (int64_t)row * osn;
#pragma unroll
        for (int c = 0; c < kDC; ++c)
            if (lane + 32 * c < D) op[lane + 32 * c] = from_f32<T>(acc[r][c] * inv);
        if (lse != nullptr && lane == 0) lse[((int64_t)b * Hq + h) * Nq + row] = m[r] + logf(dsum[r]);
    }
}

template <typename T, int kDC>
int launch_t(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv, int Nq,
             int Nk, int D, const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os,
             float scale, int causal, cudaStream_t stream) {
    auto kern = prefill_simt_kernel<T, kDC>;
    size_t smem = sizeof(float) * (size_t)(kBQ * D + kBK * (D + 1) + kBK * D + kBQ * (kBK + 1));
    PLI_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((Nq + kBQ - 1) / kBQ, Hq, B);
    kern<<<grid, kThreads, smem, stream>>>((const T*)q, (const T*)k, (const T*)v, (T*)o, lse, Hq, Hkv, Nq, Nk, D,
                                           qs[0], qs[1], qs[2], ks[0], ks[1], ks[2], vs[0], vs[1], vs[2], os[0],
                                           os[1], os[2], scale, causal);
    PLI_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return PLI_OK;
}

template <typename T>
int launch_d(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv, int Nq,
             int Nk, int D, const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os,
             float scale, int causal, cudaStream_t stream) {
    const int dc = (D + 31) / 32;
#define PLI_GO(N) return launch_t<T, N>(q, k, v, o, lse, B, Hq, Hkv, Nq, Nk, D, qs, ks, vs, os, scale, causal, stream)
    if (dc <= 1) PLI_GO(1);
    if (dc <= 2) PLI_GO(2);
    if (dc <= 4) PLI_GO(4);
    PLI_GO(8);
#undef PLI_GO
}

}  // namespace

int launch_prefill_simt(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv,
                        int Nq, int Nk, int D, const int64_t* qs, const int64_t* ks, const int64_t* vs,
                        const int64_t* os, float scale, int causal, int dtype, cudaStream_t stream) {
    if (D < 1 || D > 256) return set_error(PLI_ERR_UNSUPPORTED, "SIMT prefill supports head_dim 1..256, got %d", D);
    if (Hq > 65535 || B > 65535) return set_error(PLI_ERR_UNSUPPORTED, "B and Hq must be <= 65535");
    switch (dtype) {
        case PLI_F32: return launch_d<float>(q, k, v, o, lse, B, Hq, Hkv, Nq, Nk, D, qs, ks, vs, os, scale, causal, stream);
        case PLI_BF16: return launch_d<__nv_bfloat16>(q, k, v, o, lse, B, Hq, Hkv, Nq, Nk, D, qs, ks, vs, os, scale, causal, stream);
        case PLI_F16: return launch_d<__half>(q, k, v, o, lse, B, Hq, Hkv, Nq, Nk, D, qs, ks, vs, os, scale, causal, stream);
    }
    return set_error(PLI_ERR_INVALID, "unknown dtype %d", dtype);
}

}  // namespace pli

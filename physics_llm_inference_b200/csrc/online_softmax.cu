// online_softmax.cu — the scalar online-softmax recurrence of ch06/online_softmax.py:13-53 as row kernels.
//
//   online_softmax(x)                 m' = max(m, x_i);  d' = d e^{m-m'} + e^{x_i-m'};  result = e^{x-m} / d      (:13-25)
//   online_softmax_with_output(x, v)  the same with the running weighted sum o' = (o d e^{m-m'} + v_i e^{x_i-m'}) / d'
//                                     returns (o, d)                                                              (:28-53)
//
// This is the recurrence the prefill kernel's softmax warps run per 64-key half-step and the decode kernel runs per
// 16-token slice; here it is exposed on its own, element by element, for the reference's ch06 surface.  One warp per
// row: lane l walks elements l, l + 32, ... with the reference's update rule (un-normalised o, as flash kernels keep
// it), then the 32 lane states are merged with the same rule applied to whole partials.  fp32 arithmetic, accurate
// expf; inputs f32 / bf16 / f16, arbitrary row strides, unit inner stride.
#include "common.cuh"

namespace pli {
namespace {

struct RowState {
    float m, d;
};

__device__ __forceinline__ void merge(RowState& a, float& scale_a, float& scale_b, const RowState& b) {
    const float m = fmaxf(a.m, b.m);
    scale_a = a.m == -INFINITY ? 0.f : expf(a.m - m);
    scale_b = b.m == -INFINITY ? 0.f : expf(b.m - m);
    a.d = a.d * scale_a + b.d * scale_b;
    a.m = m;
}

template <typename T>
__global__ void __launch_bounds__(128) online_softmax_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t rows,
                                                             int n, int64_t x_row_stride, int64_t o_row_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* xr = x + row * x_row_stride;
    RowState st{-INFINITY, 0.f};
    for (int i = lane; i < n; i += 32) {
        const float xi = to_f32<T>(xr[i]);
        const float m_new = fmaxf(st.m, xi);
        st.d = st.d * (st.m == -INFINITY ? 0.f : expf(st.m - m_new)) + expf(xi - m_new);
        st.m = m_new;
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
        RowState other{__shfl_xor_sync(0xffffffffu, st.m, o2), __shfl_xor_sync(0xffffffffu, st.d, o2)};
        float sa, sb;
        merge(st, sa, sb, other);
    }
    const float inv = 1.f / st.d;
    T* orow = out + row * o_row_stride;
    for (int i = lane; i < n; i += 32) orow[i] = from_f32<T>(expf(to_f32<T>(xr[i]) - st.m) * inv);
}

// One warp per row; lane l owns value columns l, l + 32, ... (kMaxDvPerLane of them) of the running output and every
// lane runs the scalar recurrence over ALL n elements (the statistics are redundant per lane, the v reads are not).
constexpr int kMaxDvPerLane = 8;   // d_v <= 256

template <typename T>
__global__ void __launch_bounds__(128) online_softmax_output_kernel(const T* __restrict__ x, const T* __restrict__ v,
                                                                    T* __restrict__ o, float* __restrict__ d_out,
                                                                    int64_t rows, int n, int dv, int64_t x_row_stride,
                                                                    int64_t v_row_stride, int64_t v_elem_stride,
                                                                    int64_t o_row_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* xr = x + row * x_row_stride;
    const T* vr = v + row * v_row_stride;
    float m = -INFINITY, d = 0.f;
    float acc[kMaxDvPerLane];
#pragma unroll
    for (int c = 0; c < kMaxDvPerLane; ++c) acc[c] = 0.f;
    for (int i = 0; i < n; ++i) {
        const float xi = to_f32<T>(xr[i]);
        const float m_new = fmaxf(m, xi);
        const float scale_old = m == -INFINITY ? 0.f : expf(m - m_new);
        const float scale_new = expf(xi - m_new);
        d = d * scale_old + scale_new;
#pragma unroll
        for (int c = 0; c < kMaxDvPerLane; ++c) {
            const int col = lane + 32 * c;
            if (col < dv) acc[c] = acc[c] * scale_old + to_f32<T>(vr[i * v_elem_stride + col]) * scale_new;
        }
        m = m_new;
    }
    const float inv = 1.f / d;
#pragma unroll
    for (int c = 0; c < kMaxDvPerLane; ++c) {
        const int col = lane + 32 * c;
        if (col < dv) o[row * o_row_stride + col] = from_f32<T>(acc[c] * inv);
    }
    if (lane == 0) d_out[row] = d;
}

}  // namespace
}  // namespace pli

using namespace pli;

extern "C" int pli_online_softmax(const void* x, void* out, int64_t rows, int n, int64_t x_row_stride,
                                  int64_t out_row_stride, int dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !out) return set_error(PLI_ERR_INVALID, "null pointer argument");
    if (rows <= 0 || n <= 0) return set_error(PLI_ERR_INVALID, "online_softmax needs at least one row and one element");
    const int64_t blocks = (rows + 3) / 4;
    if (blocks > 0x7fffffff) return set_error(PLI_ERR_UNSUPPORTED, "too many rows");
#define PLI_OS(T) online_softmax_kernel<T><<<(unsigned)blocks, 128, 0, stream>>>((const T*)x, (T*)out, rows, n, x_row_stride, out_row_stride)
    if (dtype == PLI_F32) PLI_OS(float);
    else if (dtype == PLI_BF16) PLI_OS(__nv_bfloat16);
    else if (dtype == PLI_F16) PLI_OS(__half);
    else return set_error(PLI_ERR_INVALID, "unknown dtype %d", dtype);
#undef PLI_OS
    PLI_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return PLI_OK;
}

extern "C" int pli_online_softmax_with_output(const void* x, const void* v, void* o, float* d, int64_t rows, int n, int dv,
                                              int64_t x_row_stride, int64_t v_row_stride, int64_t v_elem_stride,
                                              int64_t o_row_stride, int dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !v || !o || !d) return set_error(PLI_ERR_INVALID, "null pointer argument");
    if (rows <= 0 || n <= 0 || dv <= 0) return set_error(PLI_ERR_INVALID, "non-positive dimension");
    if (dv > 32 * kMaxDvPerLane) return set_error(PLI_ERR_UNSUPPORTED, "value dimension %d > %d", dv, 32 * kMaxDvPerLane);
    const int64_t blocks = (rows + 3) / 4;
    if (blocks > 0x7fffffff) return set_error(PLI_ERR_UNSUPPORTED, "too many rows");
#define PLI_OS(T)                                                                                                    \
    online_softmax_output_kernel<T><<<(unsigned)blocks, 128, 0, stream>>>((const T*)x, (const T*)v, (T*)o, d, rows, n, dv, \
                                                                         x_row_stride, v_row_stride, v_elem_stride,       \
                                                                         o_row_stride)
    if (dtype == PLI_F32) PLI_OS(float);
    else if (dtype == PLI_BF16) PLI_OS(__nv_bfloat16);
    else if (dtype == PLI_F16) PLI_OS(__half);
    else return set_error(PLI_ERR_INVALID, "unknown dtype %d", dtype);
#undef PLI_OS
    PLI_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return PLI_OK;
}

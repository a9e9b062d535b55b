// conv1d.cu -- depthwise causal conv1d (+ bias) + SiLU on channels-last tokens, forward and backward.
//
// Replaces the x-branch prologue of MambaBlock.forward (models/mamba.py:176-180):
//     x = x.transpose(1, 2); x = conv1d(x)[:, :, :L]; x = x.transpose(1, 2); x = F.silu(x)
// where conv1d = nn.Conv1d(ED, ED, kernel_size=K, groups=ED, padding=K-1) (models/mamba.py:125-128), i.e.
//     pre[t, d] = bias[d] + sum_{j<K} w[d, j] * x[t - (K-1) + j, d]        (x[<0] = 0),     y = silu(pre).
// The reference pays two transposes (+ their copies) around a cuDNN depthwise kernel working on (B, ED, L); here the
// tokens stay (B, L, ED): a thread owns VEC adjacent channels and walks a time segment with the K-row window in
// registers, so every row is one coalesced 128-bit access per lane and x is read once (+ K-1 halo rows per segment).
// Backward recomputes pre, forms dpre = dy * silu'(pre), emits dx[t] = sum_j w[j] dpre[t + K-1 - j] from a sliding window
// of dpre, and accumulates dw / dbias per thread -> shared-memory block reduction -> one fp32 atomic per block and
// element (so dw / dbias are summed in a run-dependent order; they agree with the reference to fp32 rounding).
#include "../../include/mmidet_b200.h"
#include "common.cuh"

namespace mmi {

constexpr int kConvMaxK = 4;
constexpr int kConvSeg = 32;    // timesteps per thread segment
constexpr int kConvWarps = 4;   // time segments per block

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    using type = float4;
    using raw = float4;
    static __device__ __forceinline__ raw ldraw(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ void unpack(const raw &q, float (&v)[4]) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) { unpack(ldraw(p), v); }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
};
template <> struct Vec4<__nv_bfloat16> {
    using raw = uint2;
    static __device__ __forceinline__ raw ldraw(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ void unpack(const raw &q, float (&v)[4]) {
        v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
        v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[4]) { unpack(ldraw(p), v); }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 q;
        q.x = *reinterpret_cast<const uint32_t *>(&a);
        q.y = *reinterpret_cast<const uint32_t *>(&b);
        __stcs(reinterpret_cast<uint2 *>(p), q);
    }
};
template <> struct Vec4<__half> {
    using raw = uint2;
    static __device__ __forceinline__ raw ldraw(const __half *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ void unpack(const raw &q, float (&v)[4]) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&q.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&q.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void load(const __half *p, float (&v)[4]) { unpack(ldraw(p), v); }
    static __device__ __forceinline__ void store(__half *p, const float (&v)[4]) {
        const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        uint2 q;
        q.x = *reinterpret_cast<const uint32_t *>(&a);
        q.y = *reinterpret_cast<const uint32_t *>(&b);
        __stcs(reinterpret_cast<uint2 *>(p), q);
    }
};

struct ConvParams {
    const void *x, *dy;
    const float *w, *bias;
    void *y, *dx;
    float *dw, *dbias;
    int B, L, ED, K, silu;
    int64_t x_ld, y_ld, dy_ld, dx_ld;
};

// grid (ceil(ED / 128), ceil(L / (kConvSeg * kConvWarps)), B); block 32 x kConvWarps
template <typename T, int K> __global__ void __launch_bounds__(32 * kConvWarps) causal_conv1d_fwd_kernel(const ConvParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4, b = blockIdx.z;
    const int t0 = (blockIdx.y * kConvWarps + warp) * kConvSeg;
    if (c >= p.ED || t0 >= p.L) return;
    float w[K][4], bs[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        bs[v] = p.bias ? p.bias[c + v] : 0.f;
#pragma unroll
        for (int j = 0; j < K; ++j) w[j][v] = p.w[(c + v) * K + j];
    }
    const T *x = static_cast<const T *>(p.x) + int64_t(b) * p.L * p.x_ld + c;
    T *y = static_cast<T *>(p.y) + int64_t(b) * p.L * p.y_ld + c;
    // x of the last four steps lives in a ring indexed by (t - t0) & 3; the step loop is unrolled in multiples of four,
    // so every ring index is a compile-time constant and nothing is shifted between steps
    float ring[4][4];
#pragma unroll
    for (int m = 1; m < K; ++m) {
        const int t = t0 - m;
        if (t >= 0) Vec4<T>::load(x + int64_t(t) * p.x_ld, ring[(4 - m) & 3]);
        else
#pragma unroll
            for (int v = 0; v < 4; ++v) ring[(4 - m) & 3][v] = 0.f;
    }
    const int t1 = min(t0 + kConvSeg, p.L);
    constexpr int PB = sizeof(T) == 2 ? 8 : 4;  // rows requested ahead of their use (packed), see the backward kernel
    for (int tb = t0; tb < t1; tb += PB) {
        typename Vec4<T>::raw xq[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u)
            if (tb + u < t1) xq[u] = Vec4<T>::ldraw(x + int64_t(tb + u) * p.x_ld);
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int t = tb + u;
            if (t < t1) {
                Vec4<T>::unpack(xq[u], ring[u & 3]);
                float o[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float acc = bs[v];
#pragma unroll
                    for (int j = 0; j < K; ++j) acc = fmaf(w[j][v], ring[(u - (K - 1) + j) & 3][v], acc);  // x[t-(K-1)+j]
                    o[v] = p.silu ? acc * sigmoidf_fast(acc) : acc;
                }
                Vec4<T>::store(y + int64_t(t) * p.y_ld, o);
            }
        }
    }
}

template <typename T, int K> __global__ void __launch_bounds__(32 * kConvWarps) causal_conv1d_bwd_kernel(const ConvParams p) {
    __shared__ float red[kConvWarps][K + 1][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4, b = blockIdx.z;
    const int t0 = (blockIdx.y * kConvWarps + warp) * kConvSeg;
    const bool live = c < p.ED && t0 < p.L;
    float dwa[K][4], dba[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        dba[v] = 0.f;
#pragma unroll
        for (int j = 0; j < K; ++j) dwa[j][v] = 0.f;
    }
    if (live) {
        float w[K][4], bs[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            bs[v] = p.bias ? p.bias[c + v] : 0.f;
#pragma unroll
            for (int j = 0; j < K; ++j) w[j][v] = p.w[(c + v) * K + j];
        }
        const T *x = static_cast<const T *>(p.x) + int64_t(b) * p.L * p.x_ld + c;
        const T *dy = static_cast<const T *>(p.dy) + int64_t(b) * p.L * p.dy_ld + c;
        T *dx = static_cast<T *>(p.dx) + int64_t(b) * p.L * p.dx_ld + c;
        // walk t = t0 .. t1 + K - 2: dpre[t] needs x[t-K+1 .. t]; dx[s] (s = t - K + 1) needs dpre[s .. s+K-1]
        // x and dpre of the last four steps live in rings indexed by (t - t0) & 3 (compile-time constants after unrolling
        // in multiples of four: no register shifting between steps)
        float xr[4][4], dr[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) xr[j][v] = dr[j][v] = 0.f;
#pragma unroll
        for (int m = 1; m < K; ++m) {
            const int t = t0 - m;
            if (t >= 0) Vec4<T>::load(x + int64_t(t) * p.x_ld, xr[(4 - m) & 3]);
        }
        const int t1 = min(t0 + kConvSeg, p.L), tend = t1 + K - 1;
        // the walk is serial per thread: the x / dy rows of the next PB steps are requested (packed) before the first of
        // them is used, otherwise every step waits out a full memory latency
        constexpr int PB = sizeof(T) == 2 ? 8 : 4;
        for (int tb = t0; tb < tend; tb += PB) {
            typename Vec4<T>::raw xq[PB], gq[PB];
#pragma unroll
            for (int u = 0; u < PB; ++u)
                if (tb + u < tend && tb + u < p.L) {
                    xq[u] = Vec4<T>::ldraw(x + int64_t(tb + u) * p.x_ld);
                    gq[u] = Vec4<T>::ldraw(dy + int64_t(tb + u) * p.dy_ld);
                }
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int t = tb + u;
                if (t < tend) {
                    if (t < p.L) {
                        float g[4];
                        Vec4<T>::unpack(xq[u], xr[u & 3]);
                        Vec4<T>::unpack(gq[u], g);
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            float pre = bs[v];
#pragma unroll
                            for (int j = 0; j < K; ++j) pre = fmaf(w[j][v], xr[(u - (K - 1) + j) & 3][v], pre);
                            float d = g[v];
                            if (p.silu) {
                                const float sg = sigmoidf_fast(pre);
                                d *= sg * fmaf(pre, 1.f - sg, 1.f);
                            }
                            dr[u & 3][v] = d;
                            if (t < t1) {  // parameter gradients: own segment only (halo steps belong to the next segment)
                                dba[v] += d;
#pragma unroll
                                for (int j = 0; j < K; ++j) dwa[j][v] = fmaf(d, xr[(u - (K - 1) + j) & 3][v], dwa[j][v]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; ++v) xr[u & 3][v] = dr[u & 3][v] = 0.f;
                    }
                    const int sx = t - (K - 1);
                    if (sx >= t0) {  // dx[s] = sum_j w[j] * dpre[s + K-1 - j] = sum_j w[j] * dpre[t - j]
                        float o[4];
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            float acc = 0.f;
#pragma unroll
                            for (int j = 0; j < K; ++j) acc = fmaf(w[j][v], dr[(u - j) & 3][v], acc);
                            o[v] = acc;
                        }
                        Vec4<T>::store(dx + int64_t(sx) * p.dx_ld, o);
                    }
                }
            }
        }
    }
    // block reduction over the kConvWarps time segments, then one atomic per (channel, tap)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
        for (int j = 0; j < K; ++j) red[warp][j][lane * 4 + v] = dwa[j][v];
        red[warp][K][lane * 4 + v] = dba[v];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (K + 1) * 128; i += 32 * kConvWarps) {
        const int j = i / 128, cc = i % 128, ch = blockIdx.x * 128 + cc;
        if (ch >= p.ED) continue;
        float s = 0.f;
#pragma unroll
        for (int wq = 0; wq < kConvWarps; ++wq) s += red[wq][j][cc];
        if (j < K) atomicAdd(p.dw + ch * K + j, s);
        else if (p.dbias) atomicAdd(p.dbias + ch, s);
    }
}

template <typename T> static int conv_launch_t(const ConvParams &p, bool bwd, cudaStream_t st) {
    dim3 grid((p.ED + 127) / 128, (p.L + kConvSeg * kConvWarps - 1) / (kConvSeg * kConvWarps), p.B), block(32 * kConvWarps);
    if (bwd) {
        if (int e = check_cuda(cudaMemsetAsync(p.dw, 0, size_t(p.ED) * p.K * 4, st), "conv1d dw memset")) return e;
        if (p.dbias)
            if (int e = check_cuda(cudaMemsetAsync(p.dbias, 0, size_t(p.ED) * 4, st), "conv1d dbias memset")) return e;
    }
#define MMI_CONV_K(KK)                                                          \
    case KK:                                                                    \
        if (bwd) causal_conv1d_bwd_kernel<T, KK><<<grid, block, 0, st>>>(p);    \
        else causal_conv1d_fwd_kernel<T, KK><<<grid, block, 0, st>>>(p);        \
        break;
    switch (p.K) {
        MMI_CONV_K(1)
        MMI_CONV_K(2)
        MMI_CONV_K(3)
        MMI_CONV_K(4)
        default: set_error("causal_conv1d: kernel size %d unsupported (1..%d)", p.K, kConvMaxK); return MMI_ERR_UNSUPPORTED;
    }
#undef MMI_CONV_K
    return check_cuda(cudaGetLastError(), bwd ? "causal_conv1d_bwd launch" : "causal_conv1d_fwd launch");
}

int conv1d_launch(const ConvParams &p, int dtype, bool bwd, cudaStream_t st) {
    switch (dtype) {
        case MMI_F32: return conv_launch_t<float>(p, bwd, st);
        case MMI_BF16: return conv_launch_t<__nv_bfloat16>(p, bwd, st);
        case MMI_F16: return conv_launch_t<__half>(p, bwd, st);
    }
    set_error("causal_conv1d: unknown dtype %d", dtype);
    return MMI_ERR_ARG;
}

}  // namespace mmi

using namespace mmi;

extern "C" {

int mmi_causal_conv1d_fwd(const void *x, const float *w, const float *bias, void *y, int B, int L, int ED, int K, int64_t x_ld,
                          int64_t y_ld, int dtype, int silu, void *stream) {
    if (!x || !w || !y) { set_error("mmi_causal_conv1d_fwd: null required pointer"); return MMI_ERR_ARG; }
    if (B <= 0 || L <= 0 || ED <= 0 || ED % 8 || B > 65535) { set_error("mmi_causal_conv1d_fwd: bad shape (B=%d L=%d ED=%d; ED must be a multiple of 8)", B, L, ED); return MMI_ERR_ARG; }
    const int es = dtype == MMI_F32 ? 4 : 2;
    if ((x_ld * es) % 16 || (y_ld * es) % 16 || (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) {
        set_error("mmi_causal_conv1d_fwd: rows must be 16-byte aligned");
        return MMI_ERR_ARG;
    }
    int dev = 0, major = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) { set_error("libmmidet_b200 is built for sm_100a only"); return MMI_ERR_UNSUPPORTED; }
    ConvParams p{};
    p.x = x; p.w = w; p.bias = bias; p.y = y;
    p.B = B; p.L = L; p.ED = ED; p.K = K; p.silu = silu;
    p.x_ld = x_ld; p.y_ld = y_ld;
    return conv1d_launch(p, dtype, false, static_cast<cudaStream_t>(stream));
}

int mmi_causal_conv1d_bwd(const void *x, const float *w, const float *bias, const void *dy, void *dx, float *dw, float *dbias,
                          int B, int L, int ED, int K, int64_t x_ld, int64_t dy_ld, int64_t dx_ld, int dtype, int silu,
                          void *stream) {
    if (!x || !w || !dy || !dx || !dw) { set_error("mmi_causal_conv1d_bwd: null required pointer"); return MMI_ERR_ARG; }
    if (B <= 0 || L <= 0 || ED <= 0 || ED % 8 || B > 65535) { set_error("mmi_causal_conv1d_bwd: bad shape"); return MMI_ERR_ARG; }
    const int es = dtype == MMI_F32 ? 4 : 2;
    if ((x_ld * es) % 16 || (dy_ld * es) % 16 || (dx_ld * es) % 16 ||
        (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) {
        set_error("mmi_causal_conv1d_bwd: rows must be 16-byte aligned");
        return MMI_ERR_ARG;
    }
    int dev = 0, major = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) { set_error("libmmidet_b200 is built for sm_100a only"); return MMI_ERR_UNSUPPORTED; }
    ConvParams p{};
    p.x = x; p.w = w; p.bias = bias; p.dy = dy; p.dx = dx; p.dw = dw; p.dbias = dbias;
    p.B = B; p.L = L; p.ED = ED; p.K = K; p.silu = silu;
    p.x_ld = x_ld; p.dy_ld = dy_ld; p.dx_ld = dx_ld;
    return conv1d_launch(p, dtype, true, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

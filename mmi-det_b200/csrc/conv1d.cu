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

// The arithmetic runs on packed pairs (FFMA2 / FMUL2 / FADD2): a thread's 4 channels are two float2 lanes.  At 6 bytes per
// element and direction (16-bit I/O) the scalar form of these kernels was bound by instruction issue, not by HBM.
struct F4 {
    float2 lo, hi;  // channels (c, c+1) | (c+2, c+3)
};
__device__ __forceinline__ F4 f4_zero() { return {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}; }
__device__ __forceinline__ F4 f4_from(const float (&v)[4]) { return {make_float2(v[0], v[1]), make_float2(v[2], v[3])}; }
__device__ __forceinline__ F4 f4_fma(const F4 &a, const F4 &b, const F4 &c) { return {fma2(a.lo, b.lo, c.lo), fma2(a.hi, b.hi, c.hi)}; }
__device__ __forceinline__ F4 f4_mul(const F4 &a, const F4 &b) { return {mul2(a.lo, b.lo), mul2(a.hi, b.hi)}; }
__device__ __forceinline__ F4 f4_add(const F4 &a, const F4 &b) { return {add2(a.lo, b.lo), add2(a.hi, b.hi)}; }
__device__ __forceinline__ float2 sigmoid2(float2 z) {
    const float2 e = mul2(z, splat2(-kLog2e));
    const float2 d = add2(make_float2(ex2(e.x), ex2(e.y)), splat2(1.f));
    return make_float2(rcp(d.x), rcp(d.y));
}
__device__ __forceinline__ F4 f4_sigmoid(const F4 &z) { return {sigmoid2(z.lo), sigmoid2(z.hi)}; }
template <typename T> __device__ __forceinline__ F4 f4_unpack(const typename Vec4<T>::raw &q) {
    float v[4];
    Vec4<T>::unpack(q, v);
    return f4_from(v);
}
template <typename T> __device__ __forceinline__ void f4_store(T *p, const F4 &o) {
    const float v[4] = {o.lo.x, o.lo.y, o.hi.x, o.hi.y};
    Vec4<T>::store(p, v);
}

// grid (ceil(ED / 128), ceil(L / (kConvSeg * kConvWarps)), B); block 32 x kConvWarps
template <typename T, int K> __global__ void __launch_bounds__(32 * kConvWarps) causal_conv1d_fwd_kernel(const ConvParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4, b = blockIdx.z;
    const int t0 = (blockIdx.y * kConvWarps + warp) * kConvSeg;
    if (c >= p.ED || t0 >= p.L) return;
    F4 w[K], bs;
    {
        float bv[4], wv[K][4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            bv[v] = p.bias ? p.bias[c + v] : 0.f;
#pragma unroll
            for (int j = 0; j < K; ++j) wv[j][v] = p.w[(c + v) * K + j];
        }
        bs = f4_from(bv);
#pragma unroll
        for (int j = 0; j < K; ++j) w[j] = f4_from(wv[j]);
    }
    const T *x = static_cast<const T *>(p.x) + int64_t(b) * p.L * p.x_ld + c;
    T *y = static_cast<T *>(p.y) + int64_t(b) * p.L * p.y_ld + c;
    // x of the last four steps lives in a ring indexed by (t - t0) & 3; the step loop is unrolled in multiples of four,
    // so every ring index is a compile-time constant and nothing is shifted between steps
    F4 ring[4];
#pragma unroll
    for (int m = 1; m < K; ++m) {
        const int t = t0 - m;
        ring[(4 - m) & 3] = t >= 0 ? f4_unpack<T>(Vec4<T>::ldraw(x + int64_t(t) * p.x_ld)) : f4_zero();
    }
    const int t1 = min(t0 + kConvSeg, p.L);
    constexpr int PB = sizeof(T) == 2 ? 8 : 4;  // rows requested ahead of their use (packed), see the backward kernel
    const T *xr = x + int64_t(t0) * p.x_ld;
    T *yr = y + int64_t(t0) * p.y_ld;
    for (int tb = t0; tb < t1; tb += PB, xr += PB * p.x_ld, yr += PB * p.y_ld) {
        typename Vec4<T>::raw xq[PB];
        const bool whole = tb + PB <= t1;  // warp-uniform
#pragma unroll
        for (int u = 0; u < PB; ++u)
            if (whole || tb + u < t1) xq[u] = Vec4<T>::ldraw(xr + u * p.x_ld);
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            if (whole || tb + u < t1) {
                ring[u & 3] = f4_unpack<T>(xq[u]);
                F4 acc = bs;
#pragma unroll
                for (int j = 0; j < K; ++j) acc = f4_fma(w[j], ring[(u - (K - 1) + j) & 3], acc);  // x[t-(K-1)+j]
                if (p.silu) acc = f4_mul(acc, f4_sigmoid(acc));
                f4_store<T>(yr + u * p.y_ld, acc);
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
    F4 dwa[K], dba = f4_zero();
#pragma unroll
    for (int j = 0; j < K; ++j) dwa[j] = f4_zero();
    if (live) {
        F4 w[K], bs;
        {
            float bv[4], wv[K][4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                bv[v] = p.bias ? p.bias[c + v] : 0.f;
#pragma unroll
                for (int j = 0; j < K; ++j) wv[j][v] = p.w[(c + v) * K + j];
            }
            bs = f4_from(bv);
#pragma unroll
            for (int j = 0; j < K; ++j) w[j] = f4_from(wv[j]);
        }
        const T *x = static_cast<const T *>(p.x) + int64_t(b) * p.L * p.x_ld + c;
        const T *dy = static_cast<const T *>(p.dy) + int64_t(b) * p.L * p.dy_ld + c;
        T *dx = static_cast<T *>(p.dx) + int64_t(b) * p.L * p.dx_ld + c;
        // walk t = t0 .. t1 + K - 2: dpre[t] needs x[t-K+1 .. t]; dx[s] (s = t - K + 1) needs dpre[s .. s+K-1]
        // x and dpre of the last four steps live in rings indexed by (t - t0) & 3 (compile-time constants after unrolling
        // in multiples of four: no register shifting between steps)
        F4 xr[4], dr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xr[j] = dr[j] = f4_zero();
#pragma unroll
        for (int m = 1; m < K; ++m) {
            const int t = t0 - m;
            if (t >= 0) xr[(4 - m) & 3] = f4_unpack<T>(Vec4<T>::ldraw(x + int64_t(t) * p.x_ld));
        }
        const int t1 = min(t0 + kConvSeg, p.L), tend = t1 + K - 1;
        // the walk is serial per thread: the x / dy rows of the next PB steps are requested (packed) before the first of
        // them is used, otherwise every step waits out a full memory latency
        constexpr int PB = sizeof(T) == 2 ? 8 : 4;
        const T *xp = x + int64_t(t0) * p.x_ld, *gp = dy + int64_t(t0) * p.dy_ld;
        T *op = dx + int64_t(t0 - (K - 1)) * p.dx_ld;  // row of dx written at step t: t - (K - 1)
        for (int tb = t0; tb < tend; tb += PB, xp += PB * p.x_ld, gp += PB * p.dy_ld, op += PB * p.dx_ld) {
            typename Vec4<T>::raw xq[PB], gq[PB];
            // warp-uniform fast path: all PB steps are inside the thread's own segment and inside L, and (past the first
            // block) every dx row they complete belongs to the segment -- no per-step predicates
            const bool whole = tb + PB <= t1 && tb > t0;
#pragma unroll
            for (int u = 0; u < PB; ++u)
                if (whole || (tb + u < tend && tb + u < p.L)) {
                    xq[u] = Vec4<T>::ldraw(xp + u * p.x_ld);
                    gq[u] = Vec4<T>::ldraw(gp + u * p.dy_ld);
                }
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int t = tb + u;
                if (whole || t < tend) {
                    if (whole || t < p.L) {
                        xr[u & 3] = f4_unpack<T>(xq[u]);
                        F4 pre = bs;
#pragma unroll
                        for (int j = 0; j < K; ++j) pre = f4_fma(w[j], xr[(u - (K - 1) + j) & 3], pre);
                        F4 d = f4_unpack<T>(gq[u]);
                        if (p.silu) {  // d silu(pre) / d pre = sg (1 + pre (1 - sg))
                            const F4 sg = f4_sigmoid(pre);
                            const F4 one = {splat2(1.f), splat2(1.f)}, om = {add2(splat2(1.f), mul2(sg.lo, splat2(-1.f))), add2(splat2(1.f), mul2(sg.hi, splat2(-1.f)))};
                            d = f4_mul(d, f4_mul(sg, f4_fma(pre, om, one)));
                        }
                        dr[u & 3] = d;
                        if (whole || t < t1) {  // parameter gradients: own segment only (halo steps belong to the next segment)
                            dba = f4_add(dba, d);
#pragma unroll
                            for (int j = 0; j < K; ++j) dwa[j] = f4_fma(d, xr[(u - (K - 1) + j) & 3], dwa[j]);
                        }
                    } else {
                        xr[u & 3] = dr[u & 3] = f4_zero();
                    }
                    if (whole || t - (K - 1) >= t0) {  // dx[s] = sum_j w[j] * dpre[s + K-1 - j] = sum_j w[j] * dpre[t - j]
                        F4 acc = f4_mul(w[0], dr[u & 3]);
#pragma unroll
                        for (int j = 1; j < K; ++j) acc = f4_fma(w[j], dr[(u - j) & 3], acc);
                        f4_store<T>(op + u * p.dx_ld, acc);
                    }
                }
            }
        }
    }
    // block reduction over the kConvWarps time segments, then one atomic per (channel, tap)
#pragma unroll
    for (int j = 0; j < K; ++j) *reinterpret_cast<float4 *>(&red[warp][j][lane * 4]) = make_float4(dwa[j].lo.x, dwa[j].lo.y, dwa[j].hi.x, dwa[j].hi.y);
    *reinterpret_cast<float4 *>(&red[warp][K][lane * 4]) = make_float4(dba.lo.x, dba.lo.y, dba.hi.x, dba.hi.y);
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

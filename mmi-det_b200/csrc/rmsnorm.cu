// rmsnorm.cu -- RMSNorm over the channel axis of (rows, C) tokens, forward and backward, one warp per row.
//
// Replaces RMSNorm.forward (models/mamba.py:356-366): y = x * rsqrt(mean(x^2) + eps) * weight, which the reference (and
// stock PyTorch) evaluates as five elementwise / reduction kernels per direction.  HBM-bound: the forward reads x and
// writes y once (the row stays in registers between the reduction and the scaling); the backward reads x and dy once,
//     dx = r * (w*dy - x * r^2 * mean(w*dy*x)),   r = rsqrt(mean(x^2) + eps),   dw = sum_rows dy * x * r,
// with dw accumulated per lane over a strip of rows and finished with one fp32 atomic per (block, channel).
#include "../../include/mmidet_b200.h"
#include "common.cuh"

namespace mmi {

// rows up to 1024 channels: one warp per row, the row in registers (32 lanes x 4 x 8 float4 groups); wider rows: one CTA per row
constexpr int kRmsRowsPerWarp = 32;       // rows per warp strip in the backward (dw is reduced over the strip, then over the
                                          // block's four warps in shared memory: the C addresses of dw sit in a handful of L2
                                          // slices, so every atomic saved matters)

template <typename T> __device__ __forceinline__ void rms_load4(const T *p, float (&v)[4]);
template <> __device__ __forceinline__ void rms_load4<float>(const float *p, float (&v)[4]) {
    const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void rms_load4<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[4]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
    v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void rms_load4<__half>(const __half *p, float (&v)[4]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&q.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&q.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
// packed row fragments: requested several rows ahead and unpacked only when consumed (2 registers per 4 16-bit elements)
template <typename T> struct RmsRaw { using type = uint2; };
template <> struct RmsRaw<float> { using type = float4; };
template <typename T> __device__ __forceinline__ typename RmsRaw<T>::type rms_ldraw(const T *p) {
    return __ldg(reinterpret_cast<const typename RmsRaw<T>::type *>(p));
}
template <typename T> __device__ __forceinline__ void rms_unpack(const typename RmsRaw<T>::type &q, float (&v)[4]);
template <> __device__ __forceinline__ void rms_unpack<float>(const float4 &q, float (&v)[4]) {
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void rms_unpack<__nv_bfloat16>(const uint2 &q, float (&v)[4]) {
    v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void rms_unpack<__half>(const uint2 &q, float (&v)[4]) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&q.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&q.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ void rms_store4(T *p, const float (&v)[4]);
template <> __device__ __forceinline__ void rms_store4<float>(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void rms_store4<__nv_bfloat16>(__nv_bfloat16 *p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 q;
    q.x = *reinterpret_cast<const uint32_t *>(&a);
    q.y = *reinterpret_cast<const uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = q;
}
template <> __device__ __forceinline__ void rms_store4<__half>(__half *p, const float (&v)[4]) {
    const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    uint2 q;
    q.x = *reinterpret_cast<const uint32_t *>(&a);
    q.y = *reinterpret_cast<const uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = q;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// NV = number of float4 groups per lane actually used (C <= 128 * NV); rows = B * L, x/y row pitches in elements
template <typename T, typename TY, int NV>
__global__ void __launch_bounds__(128) rmsnorm_fwd_kernel(const T *__restrict__ x, const float *__restrict__ w, TY *__restrict__ y,
                                                          float *__restrict__ rstd, int64_t rows, int C, int64_t x_ld, int64_t y_ld,
                                                          float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = int64_t(blockIdx.x) * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    float v[NV][4];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) {
            rms_load4<T>(x + row * x_ld + c, v[i]);
#pragma unroll
            for (int k = 0; k < 4; ++k) ss = fmaf(v[i][k], v[i][k], ss);
        }
    }
    const float r = rsqrtf(warp_sum(ss) / float(C) + eps);
    if (rstd && lane == 0) rstd[row] = r;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) {
            const float4 ww = __ldg(reinterpret_cast<const float4 *>(w + c));
            float o[4] = {v[i][0] * r * ww.x, v[i][1] * r * ww.y, v[i][2] * r * ww.z, v[i][3] * r * ww.w};
            rms_store4<TY>(y + row * y_ld + c, o);
        }
    }
}

// Backward, rows up to 1024 channels.  A row is owned by a group of LPR = 8 / 16 / 32 lanes (C <= 256 / 512 / 1024), so a warp
// works on 32 / LPR rows at once: the two row reductions (sum x^2, sum w dy x) take log2(LPR) shuffle rounds for ALL of the
// warp's rows together instead of five rounds per row, the arithmetic runs on packed pairs (FFMA2 / FMUL2), and a lane keeps
// its <= 32 channels of w and of the dw accumulator in registers over the whole strip of rows.  NCH = chunks of 4 channels per
// lane; chunk i of sub-lane s covers channels (i * LPR + s) * 4 (coalesced across the group).
template <typename T, typename TY, int LPR, int NCH>
__global__ void __launch_bounds__(128) rmsnorm_bwd_kernel(const T *__restrict__ x, const float *__restrict__ w, const TY *__restrict__ dy,
                                                          T *__restrict__ dx, float *__restrict__ dw, int64_t rows, int C, int64_t x_ld,
                                                          int64_t dy_ld, int64_t dx_ld, float eps) {
    constexpr int RW = 32 / LPR;  // rows per warp step
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LPR, rw = lane / LPR;
    const int64_t row0 = (int64_t(blockIdx.x) * 4 + warp) * kRmsRowsPerWarp;
    float2 wv[NCH][2], dwa[NCH][2];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (i * LPR + sub) * 4;
        dwa[i][0] = dwa[i][1] = make_float2(0.f, 0.f);
        wv[i][0] = wv[i][1] = make_float2(0.f, 0.f);
        if (c < C) {
            const float4 ww = __ldg(reinterpret_cast<const float4 *>(w + c));
            wv[i][0] = make_float2(ww.x, ww.y), wv[i][1] = make_float2(ww.z, ww.w);
        }
    }
    using Raw = typename RmsRaw<T>::type;
    using RawY = typename RmsRaw<TY>::type;
    const float invC = 1.f / float(C);
    for (int r0 = 0; r0 < kRmsRowsPerWarp; r0 += RW) {
        const int64_t row = row0 + r0 + rw;
        const bool live = row < rows;
        if (row0 + r0 >= rows) break;  // warp-uniform
        Raw xq[NCH];
        RawY gq[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (i * LPR + sub) * 4;
            if (live && c < C) {
                xq[i] = rms_ldraw<T>(x + row * x_ld + c);
                gq[i] = rms_ldraw<TY>(dy + row * dy_ld + c);
            }
        }
        float2 xv[NCH][2], gw[NCH][2], gv[NCH][2];
        float2 ss2 = make_float2(0.f, 0.f), sg2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (i * LPR + sub) * 4;
            float a[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
            if (live && c < C) {
                rms_unpack<T>(xq[i], a);
                rms_unpack<TY>(gq[i], g);
            }
            xv[i][0] = make_float2(a[0], a[1]), xv[i][1] = make_float2(a[2], a[3]);
            gv[i][0] = make_float2(g[0], g[1]), gv[i][1] = make_float2(g[2], g[3]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                gw[i][h] = mul2(gv[i][h], wv[i][h]);
                ss2 = fma2(xv[i][h], xv[i][h], ss2);
                sg2 = fma2(gw[i][h], xv[i][h], sg2);
            }
        }
        float ss = ss2.x + ss2.y, sg = sg2.x + sg2.y;
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) {  // within the row's lane group; all rows of the warp at once
            ss += __shfl_xor_sync(0xffffffffu, ss, m);
            sg += __shfl_xor_sync(0xffffffffu, sg, m);
        }
        const float r = rsqrtf(ss * invC + eps);
        const float2 r2 = splat2(r), ncoef = splat2(-(r * r * r * sg * invC));
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (i * LPR + sub) * 4;
            float2 o[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                o[h] = fma2(xv[i][h], ncoef, mul2(gw[i][h], r2));                 // r w dy - x r^3 mean(w dy x)
                dwa[i][h] = fma2(mul2(gv[i][h], r2), xv[i][h], dwa[i][h]);        // dw += dy x r
            }
            if (live && c < C) {
                const float ov[4] = {o[0].x, o[0].y, o[1].x, o[1].y};
                rms_store4<T>(dx + row * dx_ld + c, ov);
            }
        }
    }
    // dw: sum over the warp's row groups and the block's four warps in shared memory, one atomic per channel and block
    __shared__ float4 red[4][RW][NCH * LPR];
#pragma unroll
    for (int i = 0; i < NCH; ++i) red[warp][rw][i * LPR + sub] = make_float4(dwa[i][0].x, dwa[i][0].y, dwa[i][1].x, dwa[i][1].y);
    __syncthreads();
    for (int q = threadIdx.x; q < NCH * LPR; q += 128) {
        const int c = q * 4;
        if (c >= C) continue;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int wq = 0; wq < 4; ++wq)
#pragma unroll
            for (int g = 0; g < RW; ++g) {
                const float4 v = red[wq][g][q];
                s4.x += v.x, s4.y += v.y, s4.z += v.z, s4.w += v.w;
            }
        atomicAdd(dw + c, s4.x), atomicAdd(dw + c + 1, s4.y), atomicAdd(dw + c + 2, s4.z), atomicAdd(dw + c + 3, s4.w);
    }
}

// ---- wide rows (1024 < C <= 4096, e.g. d_model = 1280 at P5 of YOLOv5x): one 256-thread CTA per row strip -------------
constexpr int kRmsWideThreads = 256;
constexpr int kRmsWideMaxC = 4096;
constexpr int kRmsWideRows = 8;  // rows per CTA strip in the backward

__device__ __forceinline__ float block_sum256(float v, float *red) {  // red: 8 floats; every thread gets the total
    v = warp_sum(v);
    __syncthreads();  // protects `red` against the previous call's readers
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kRmsWideThreads / 32; ++i) t += red[i];
    return t;
}

template <typename T, typename TY, int NV>
__global__ void __launch_bounds__(kRmsWideThreads) rmsnorm_wide_fwd_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                                           TY *__restrict__ y, int64_t rows, int C, int64_t x_ld,
                                                                           int64_t y_ld, float eps) {
    __shared__ float red[8];
    const int64_t row = blockIdx.x;
    float v[NV][4];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * kRmsWideThreads + threadIdx.x) * 4;
        if (c < C) {
            rms_load4<T>(x + row * x_ld + c, v[i]);
#pragma unroll
            for (int k = 0; k < 4; ++k) ss = fmaf(v[i][k], v[i][k], ss);
        }
    }
    const float r = rsqrtf(block_sum256(ss, red) / float(C) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * kRmsWideThreads + threadIdx.x) * 4;
        if (c < C) {
            const float4 ww = __ldg(reinterpret_cast<const float4 *>(w + c));
            float o[4] = {v[i][0] * r * ww.x, v[i][1] * r * ww.y, v[i][2] * r * ww.z, v[i][3] * r * ww.w};
            rms_store4<TY>(y + row * y_ld + c, o);
        }
    }
}

template <typename T, typename TY, int NV>
__global__ void __launch_bounds__(kRmsWideThreads) rmsnorm_wide_bwd_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                                           const TY *__restrict__ dy, T *__restrict__ dx,
                                                                           float *__restrict__ dw, int64_t rows, int C, int64_t x_ld,
                                                                           int64_t dy_ld, int64_t dx_ld, float eps) {
    __shared__ float red[8];
    const int64_t row0 = int64_t(blockIdx.x) * kRmsWideRows;
    const int nrow = int(min(int64_t(kRmsWideRows), rows - row0));
    float wv[NV][4], dwa[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * kRmsWideThreads + threadIdx.x) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) dwa[i][k] = 0.f, wv[i][k] = 0.f;
        if (c < C) {
            const float4 ww = __ldg(reinterpret_cast<const float4 *>(w + c));
            wv[i][0] = ww.x; wv[i][1] = ww.y; wv[i][2] = ww.z; wv[i][3] = ww.w;
        }
    }
    for (int rr = 0; rr < nrow; ++rr) {
        const int64_t row = row0 + rr;
        float xv[NV][4], gv[NV][4];
        float ss = 0.f, sg = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * kRmsWideThreads + threadIdx.x) * 4;
            if (c < C) {
                rms_load4<T>(x + row * x_ld + c, xv[i]);
                rms_load4<TY>(dy + row * dy_ld + c, gv[i]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ss = fmaf(xv[i][k], xv[i][k], ss);
                    sg = fmaf(gv[i][k] * wv[i][k], xv[i][k], sg);
                }
            }
        }
        ss = block_sum256(ss, red);
        sg = block_sum256(sg, red);
        const float r = rsqrtf(ss / float(C) + eps);
        const float coef = r * r * r * sg / float(C);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * kRmsWideThreads + threadIdx.x) * 4;
            if (c < C) {
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    o[k] = fmaf(gv[i][k] * wv[i][k], r, -xv[i][k] * coef);
                    dwa[i][k] = fmaf(gv[i][k] * r, xv[i][k], dwa[i][k]);
                }
                rms_store4<T>(dx + row * dx_ld + c, o);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * kRmsWideThreads + threadIdx.x) * 4;
        if (c < C)
#pragma unroll
            for (int k = 0; k < 4; ++k) atomicAdd(dw + c + k, dwa[i][k]);
    }
}

// T: dtype of x / dx; TY: dtype of y (forward) and dy (backward) -- equal to T, or 16-bit with fp32 x (autocast: the
// norm runs in fp32 and its output is consumed by a 16-bit GEMM, so the rounding happens in the kernel's store)
template <typename T, typename TY> static int rms_launch_t(bool bwd, const void *x, const float *w, const void *dy, void *out, float *dw,
                                                           int64_t rows, int C, int64_t x_ld, int64_t dy_ld, int64_t out_ld, float eps,
                                                           cudaStream_t st) {
    const T *xp = static_cast<const T *>(x);
    if (bwd)
        if (int e = check_cuda(cudaMemsetAsync(dw, 0, size_t(C) * 4, st), "rmsnorm dw memset")) return e;
    const unsigned gf = unsigned((rows + 3) / 4), gb = unsigned((rows + 4 * kRmsRowsPerWarp - 1) / (4 * kRmsRowsPerWarp));
#define MMI_RMS(NVV)                                                                                                   \
    if (!bwd && C <= 128 * NVV) {                                                                                      \
        rmsnorm_fwd_kernel<T, TY, NVV><<<gf, 128, 0, st>>>(xp, w, static_cast<TY *>(out), nullptr, rows, C, x_ld, out_ld, eps); \
        return check_cuda(cudaGetLastError(), "rmsnorm launch");                                                      \
    }
    MMI_RMS(1)
    MMI_RMS(2)
    MMI_RMS(4)
    MMI_RMS(8)
#undef MMI_RMS
#define MMI_RMS_B(LPR, NCH)                                                                                            \
    if (bwd && C <= 4 * LPR * NCH) {                                                                                   \
        rmsnorm_bwd_kernel<T, TY, LPR, NCH><<<gb, 128, 0, st>>>(xp, w, static_cast<const TY *>(dy), static_cast<T *>(out), dw, rows, C, \
                                                                x_ld, dy_ld, out_ld, eps);                             \
        return check_cuda(cudaGetLastError(), "rmsnorm backward launch");                                             \
    }
    MMI_RMS_B(8, 2)   // C <= 64
    MMI_RMS_B(8, 4)   // C <= 128
    MMI_RMS_B(8, 8)   // C <= 256
    MMI_RMS_B(16, 8)  // C <= 512
    MMI_RMS_B(32, 8)  // C <= 1024
#undef MMI_RMS_B
#define MMI_RMS_WIDE(NVV)                                                                                              \
    if (C <= kRmsWideThreads * 4 * NVV) {                                                                              \
        if (bwd)                                                                                                       \
            rmsnorm_wide_bwd_kernel<T, TY, NVV><<<unsigned((rows + kRmsWideRows - 1) / kRmsWideRows), kRmsWideThreads, 0, st>>>( \
                xp, w, static_cast<const TY *>(dy), static_cast<T *>(out), dw, rows, C, x_ld, dy_ld, out_ld, eps);     \
        else                                                                                                           \
            rmsnorm_wide_fwd_kernel<T, TY, NVV><<<unsigned(rows), kRmsWideThreads, 0, st>>>(xp, w, static_cast<TY *>(out), rows, C, \
                                                                                            x_ld, out_ld, eps);       \
        return check_cuda(cudaGetLastError(), "rmsnorm (wide rows) launch");                                           \
    }
    MMI_RMS_WIDE(2)
    MMI_RMS_WIDE(4)
#undef MMI_RMS_WIDE
    set_error("rmsnorm: C=%d exceeds %d", C, kRmsWideMaxC);
    return MMI_ERR_UNSUPPORTED;
}

static int rms_dispatch(bool bwd, const void *x, const float *w, const void *dy, void *out, float *dw, int64_t rows, int C,
                        int64_t x_ld, int64_t dy_ld, int64_t out_ld, float eps, int dtype, int y_dtype, cudaStream_t st) {
#define MMI_RMS_GO(T, TY) return rms_launch_t<T, TY>(bwd, x, w, dy, out, dw, rows, C, x_ld, dy_ld, out_ld, eps, st)
    if (y_dtype == dtype) {
        if (dtype == MMI_F32) MMI_RMS_GO(float, float);
        if (dtype == MMI_BF16) MMI_RMS_GO(__nv_bfloat16, __nv_bfloat16);
        MMI_RMS_GO(__half, __half);
    }
    if (dtype == MMI_F32 && y_dtype == MMI_BF16) MMI_RMS_GO(float, __nv_bfloat16);
    if (dtype == MMI_F32 && y_dtype == MMI_F16) MMI_RMS_GO(float, __half);
#undef MMI_RMS_GO
    set_error("rmsnorm: unsupported dtype pair (x %d, y/dy %d): y/dy must match x, or be 16-bit with fp32 x", dtype, y_dtype);
    return MMI_ERR_UNSUPPORTED;
}

}  // namespace mmi

using namespace mmi;

extern "C" {

static int rms_esize(int dtype) { return dtype == MMI_F32 ? 4 : (dtype == MMI_BF16 || dtype == MMI_F16) ? 2 : 0; }

static int rms_check(const char *who, int64_t rows, int C, int dtype, int y_dtype, const void *a, const void *b, int64_t lda,
                     int64_t ldb) {
    if (rows <= 0 || C <= 0 || C % 8) { set_error("%s: bad shape (rows=%lld C=%d; C must be a multiple of 8)", who, (long long)rows, C); return MMI_ERR_ARG; }
    const int es = rms_esize(dtype), eb = rms_esize(y_dtype);
    if (!es || !eb) { set_error("%s: unknown dtype %d / %d", who, dtype, y_dtype); return MMI_ERR_ARG; }
    if ((lda * es) % 16 || (ldb * eb) % (eb == 2 ? 8 : 16) || (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15)) {
        set_error("%s: rows must be 16-byte aligned", who);
        return MMI_ERR_ARG;
    }
    int dev = 0, major = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) { set_error("libmmidet_b200 is built for sm_100a only"); return MMI_ERR_UNSUPPORTED; }
    return MMI_OK;
}

int mmi_rmsnorm_fwd(const void *x, const float *w, void *y, int64_t rows, int C, int64_t x_ld, int64_t y_ld, float eps, int dtype,
                    int y_dtype, void *stream) {
    if (!x || !w || !y) { set_error("mmi_rmsnorm_fwd: null pointer"); return MMI_ERR_ARG; }
    if (y_dtype < 0) y_dtype = dtype;
    if (int e = rms_check("mmi_rmsnorm_fwd", rows, C, dtype, y_dtype, x, y, x_ld, y_ld)) return e;
    return rms_dispatch(false, x, w, nullptr, y, nullptr, rows, C, x_ld, 0, y_ld, eps, dtype, y_dtype, static_cast<cudaStream_t>(stream));
}

int mmi_rmsnorm_bwd(const void *x, const float *w, const void *dy, void *dx, float *dw, int64_t rows, int C, int64_t x_ld,
                    int64_t dy_ld, int64_t dx_ld, float eps, int dtype, int dy_dtype, void *stream) {
    if (!x || !w || !dy || !dx || !dw) { set_error("mmi_rmsnorm_bwd: null pointer"); return MMI_ERR_ARG; }
    if (dy_dtype < 0) dy_dtype = dtype;
    if (int e = rms_check("mmi_rmsnorm_bwd", rows, C, dtype, dtype, x, dx, x_ld, dx_ld)) return e;
    if (int e = rms_check("mmi_rmsnorm_bwd", rows, C, dtype, dy_dtype, x, dy, x_ld, dy_ld)) return e;
    return rms_dispatch(true, x, w, dy, dx, dw, rows, C, x_ld, dy_ld, dx_ld, eps, dtype, dy_dtype, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ffm.cu -- Fusion Focus Module Fourier step + separation loss.
//
// extract_frequency2 (models/common.py:37-69) is fftn -> fftshift -> box masks -> ifftshift -> ifftn -> real -> fp16,
// ~18 tiny launches per modality on an (B, C, 8, 8) tensor: launch-latency bound.  The box masks are separable
// (kept rows x kept columns of the shifted spectrum, including the negative-slice wrap of common.py:44-56), so the
// low-pass image is the exact projection
//     low = Re( Pr . x . Pc^T ),   Pr[h,h'] = 1/H sum_{k in Kr} e^{2 pi i k (h-h')/H}   (circulant, likewise Pc)
// evaluated directly in one kernel, one CTA per (b, c) image; high = x - low because the reference's high-pass
// zeroes exactly the block its low-pass keeps.  Outputs are rounded to fp16 as the reference's .half() does
// (common.py:66-67); the product high * fea of common.py:440-441 is emitted from the same kernel.
#include "../../include/mmidet_b200.h"
#include "common.cuh"

namespace mmi {

constexpr int kFfmMax = 64;

template <typename T>
__global__ void __launch_bounds__(128) ffm_extract_kernel(const T *__restrict__ img, __half *__restrict__ low,
                                                          __half *__restrict__ high, float *__restrict__ high_mul, int H,
                                                          int W, int r0, int r1, int c0, int c1,
                                                          const T *__restrict__ img2 = nullptr,
                                                          float *__restrict__ high_mul2 = nullptr) {
    if (blockIdx.y) {  // second modality of the pattern path (product only)
        img = img2;
        high_mul = high_mul2;
    }
    extern __shared__ float sm[];
    float *x = sm;               // [H][W]
    float *tre = x + H * W;      // [H][W]  T = x . Pc^T (complex)
    float *tim = tre + H * W;
    float *pr = tim + H * W;     // [H] complex circulant generators: pr[2*d], pr[2*d+1]
    float *pc = pr + 2 * H;      // [W]
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t off = int64_t(blockIdx.x) * H * W;

    for (int i = tid; i < H * W; i += nt) x[i] = to_f32<T>(img[off + i]);
    // generators: p[d] = 1/n sum_{s in [s0,s1)} exp(2 pi i k d / n), k = (s - n/2) mod n  (fftshift index -> frequency)
    for (int i = tid; i < H + W; i += nt) {
        const bool row = i < H;
        const int n = row ? H : W, d = row ? i : i - H, s0 = row ? r0 : c0, s1 = row ? r1 : c1;
        float re = 0.f, im = 0.f;
        for (int s = s0; s < s1; ++s) {
            int k = (s - n / 2) % n;
            if (k < 0) k += n;
            const int kd = (k * d) % n;  // exact phase reduction
            float sn, cs;
            sincospif(2.0f * float(kd) / float(n), &sn, &cs);
            re += cs;
            im += sn;
        }
        float *p = row ? pr : pc;
        p[2 * d] = re / float(n);
        p[2 * d + 1] = im / float(n);
    }
    __syncthreads();
    // T[h,w] = sum_w' x[h,w'] pc[(w - w') mod W]
    for (int i = tid; i < H * W; i += nt) {
        const int h = i / W, w = i % W;
        float re = 0.f, im = 0.f;
        for (int w2 = 0; w2 < W; ++w2) {
            int d = w - w2;
            if (d < 0) d += W;
            const float v = x[h * W + w2];
            re = fmaf(v, pc[2 * d], re);
            im = fmaf(v, pc[2 * d + 1], im);
        }
        tre[i] = re;
        tim[i] = im;
    }
    __syncthreads();
    // low[h,w] = Re sum_h' pr[(h - h') mod H] T[h',w]
    for (int i = tid; i < H * W; i += nt) {
        const int h = i / W, w = i % W;
        float lo = 0.f;
        for (int h2 = 0; h2 < H; ++h2) {
            int d = h - h2;
            if (d < 0) d += H;
            lo = fmaf(pr[2 * d], tre[h2 * W + w], lo);
            lo = fmaf(-pr[2 * d + 1], tim[h2 * W + w], lo);
        }
        const float xv = x[i];
        const __half l16 = __float2half_rn(lo), h16 = __float2half_rn(xv - lo);
        if (low) low[off + i] = l16;
        if (high) high[off + i] = h16;
        if (high_mul) high_mul[off + i] = __half2float(h16) * xv;  // torch.mul(high.half(), fea) -> fp32
    }
}

// Maps larger than 64 x 64 (the reference only ever calls extract_frequency2 on the 8 x 8 pooled maps, but the helper is
// public and size-generic): the kept block is nr x nc = about (W/8) x (W/8) bins of the shifted spectrum, so the
// projection is evaluated through those bins only --
//     T1[h,j] = sum_w x[h,w] e^{-i kc_j w},  X[i,j] = sum_h T1[h,j] e^{-i kr_i h},
//     U[h,j]  = sum_i X[i,j] e^{+i kr_i h},  low[h,w] = Re sum_j U[h,j] e^{+i kc_j w} / (H W)
// one CTA per (b, c) image, T1 / U / X and the two twiddle tables in shared memory, x read from global memory twice.
constexpr int kFfmWideThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kFfmWideThreads) ffm_extract_wide_kernel(const T *__restrict__ img, __half *__restrict__ low,
                                                                           __half *__restrict__ high, float *__restrict__ high_mul,
                                                                           int H, int W, int r0, int r1, int c0, int c1) {
    extern __shared__ float sm[];
    const int nr = r1 - r0, nc = c1 - c0;
    float2 *twH = reinterpret_cast<float2 *>(sm);  // [H] (cos, sin)(2 pi k / H)
    float2 *twW = twH + H;                         // [W]
    float2 *T1 = twW + W;                          // [H][nc]
    float2 *U = T1 + size_t(H) * nc;               // [H][nc]
    float2 *X = U + size_t(H) * nc;                // [nr][nc]
    const int tid = threadIdx.x, nt = blockDim.x;
    const T *x = img + int64_t(blockIdx.x) * H * W;
    const int64_t off = int64_t(blockIdx.x) * H * W;
    for (int i = tid; i < H + W; i += nt) {
        const bool row = i < H;
        const int n = row ? H : W, k = row ? i : i - H;
        float sn, cs;
        sincospif(2.0f * float(k) / float(n), &sn, &cs);
        (row ? twH : twW)[k] = make_float2(cs, sn);
    }
    __syncthreads();
    auto freq = [](int s, int n) {  // fftshift index -> frequency
        int k = (s - n / 2) % n;
        return k < 0 ? k + n : k;
    };
    for (int it = tid; it < H * nc; it += nt) {
        const int h = it / nc, j = it % nc, kc = freq(c0 + j, W);
        float re = 0.f, im = 0.f;
        int idx = 0;
        for (int w = 0; w < W; ++w) {
            const float v = to_f32<T>(x[h * W + w]);
            const float2 t = twW[idx];
            re = fmaf(v, t.x, re);
            im = fmaf(-v, t.y, im);
            idx += kc;
            if (idx >= W) idx -= W;
        }
        T1[it] = make_float2(re, im);
    }
    __syncthreads();
    for (int it = tid; it < nr * nc; it += nt) {
        const int i = it / nc, j = it % nc, kr = freq(r0 + i, H);
        float re = 0.f, im = 0.f;
        int idx = 0;
        for (int h = 0; h < H; ++h) {
            const float2 a = T1[h * nc + j], t = twH[idx];  // a * conj(t)
            re += a.x * t.x + a.y * t.y;
            im += a.y * t.x - a.x * t.y;
            idx += kr;
            if (idx >= H) idx -= H;
        }
        X[it] = make_float2(re, im);
    }
    __syncthreads();
    for (int it = tid; it < H * nc; it += nt) {
        const int h = it / nc, j = it % nc;
        float re = 0.f, im = 0.f;
        for (int i = 0; i < nr; ++i) {
            const int kr = freq(r0 + i, H);
            const float2 a = X[i * nc + j], t = twH[int((int64_t(kr) * h) % H)];  // a * t
            re += a.x * t.x - a.y * t.y;
            im += a.x * t.y + a.y * t.x;
        }
        U[it] = make_float2(re, im);
    }
    __syncthreads();
    const float scale = 1.0f / (float(H) * float(W));
    for (int it = tid; it < H * W; it += nt) {
        const int h = it / W, w = it % W;
        float lo = 0.f;
        for (int j = 0; j < nc; ++j) {
            const int kc = freq(c0 + j, W);
            const float2 a = U[h * nc + j], t = twW[int((int64_t(kc) * w) % W)];
            lo += a.x * t.x - a.y * t.y;
        }
        lo *= scale;
        const float xv = to_f32<T>(x[it]);
        const __half l16 = __float2half_rn(lo), h16 = __float2half_rn(xv - lo);
        if (low) low[off + it] = l16;
        if (high) high[off + it] = h16;
        if (high_mul) high_mul[off + it] = __half2float(h16) * xv;
    }
}

// models/common.py:128-139 in closed form; one CTA.
__global__ void __launch_bounds__(256) separation_loss_kernel(const float *__restrict__ M, float *__restrict__ loss, int l,
                                                              int K) {
    __shared__ float red[256];
    __shared__ float part[2][256];
    float acc = 0.f;
    if (K <= 128) {  // the pattern path: K = 64 columns, up to 18 * B rows -- split the rows over 256 / K thread groups
        const int G = 256 / K, g = threadIdx.x / K, k = threadIdx.x % K;
        float s = 0.f, q = 0.f;
        if (g < G)
            for (int i = g; i < l; i += G) {
                const float v = M[int64_t(i) * K + k];
                s += v;
                q = fmaf(v, v, q);
            }
        part[0][threadIdx.x] = s;
        part[1][threadIdx.x] = q;
        __syncthreads();
        if (threadIdx.x < K) {
            s = 0.f, q = 0.f;
            for (int gg = 0; gg < G; ++gg) {
                s += part[0][gg * K + k];
                q += part[1][gg * K + k];
            }
            acc = s * s - q;
        }
    } else {
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            float s = 0.f, q = 0.f;
            for (int i = 0; i < l; ++i) {
                const float v = M[int64_t(i) * K + k];
                s += v;
                q = fmaf(v, v, q);
            }
            acc += s * s - q;  // column k of |sum_i M_i|^2 - sum_i |M_i|^2
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = red[0] / (2.0f * float(l) * float(l - 1));
}

void ffm_kept_range(int H, int W, int *r0, int *r1, int *c0, int *c1) {
    // common.py:41-56: threshold = crow + ccol // 4; slice(crow - thr, crow + thr) with Python semantics
    const int crow = H / 2, ccol = W / 2, thr = crow + ccol / 4;
    auto clampi = [](int v, int n) {
        if (v < 0) v += n;
        if (v < 0) v = 0;
        if (v > n) v = n;
        return v;
    };
    *r0 = clampi(crow - thr, H);
    *r1 = clampi(crow + thr, H);
    *c0 = clampi(ccol - thr, W);
    *c1 = clampi(ccol + thr, W);
    if (*r1 < *r0) *r1 = *r0;
    if (*c1 < *c0) *c1 = *c0;
}

int ffm_extract_launch(const void *img, void *low, void *high, float *high_mul, int BC, int H, int W, int dtype,
                       cudaStream_t st) {
    if (H < 1 || W < 1) {
        set_error("mmi_ffm_extract: bad map size %dx%d", H, W);
        return MMI_ERR_ARG;
    }
    int r0, r1, c0, c1;
    ffm_kept_range(H, W, &r0, &r1, &c0, &c1);
    __half *lo = static_cast<__half *>(low), *hi = static_cast<__half *>(high);
    if (H > kFfmMax || W > kFfmMax) {  // large maps: projection through the kept bins only
        const size_t nr = size_t(r1 - r0), nc = size_t(c1 - c0);
        const size_t smem_w = (size_t(H) + W + 2 * size_t(H) * nc + nr * nc) * sizeof(float2);
        if (smem_w > 200 * 1024) {
            set_error("mmi_ffm_extract: %dx%d map keeps %zux%zu bins, which needs %zu bytes of shared memory (limit 200 KiB)", H, W,
                      nr, nc, smem_w);
            return MMI_ERR_UNSUPPORTED;
        }
#define MMI_FFM_WIDE(T)                                                                                                  \
    do {                                                                                                                 \
        auto kern = ffm_extract_wide_kernel<T>;                                                                          \
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_w)),     \
                               "ffm smem attribute"))                                                                    \
            return e;                                                                                                    \
        kern<<<BC, kFfmWideThreads, smem_w, st>>>(static_cast<const T *>(img), lo, hi, high_mul, H, W, r0, r1, c0, c1);  \
    } while (0)
        switch (dtype) {
            case MMI_F32: MMI_FFM_WIDE(float); break;
            case MMI_BF16: MMI_FFM_WIDE(__nv_bfloat16); break;
            case MMI_F16: MMI_FFM_WIDE(__half); break;
            default: set_error("mmi_ffm_extract: unknown dtype %d", dtype); return MMI_ERR_ARG;
        }
#undef MMI_FFM_WIDE
        return check_cuda(cudaGetLastError(), "ffm_extract (wide) launch");
    }
    const size_t smem = (size_t(3) * H * W + 2 * (H + W)) * sizeof(float);
#define MMI_FFM_LAUNCH(T)                                                                                               \
    do {                                                                                                                \
        auto kern = ffm_extract_kernel<T>;                                                                              \
        if (smem > 48 * 1024)                                                                                           \
            if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)), \
                                   "ffm smem attribute"))                                                               \
                return e;                                                                                               \
        kern<<<BC, 128, smem, st>>>(static_cast<const T *>(img), lo, hi, high_mul, H, W, r0, r1, c0, c1, nullptr, nullptr); \
    } while (0)
    switch (dtype) {
        case MMI_F32: MMI_FFM_LAUNCH(float); break;
        case MMI_BF16: MMI_FFM_LAUNCH(__nv_bfloat16); break;
        case MMI_F16: MMI_FFM_LAUNCH(__half); break;
        default: set_error("mmi_ffm_extract: unknown dtype %d", dtype); return MMI_ERR_ARG;
    }
#undef MMI_FFM_LAUNCH
    return check_cuda(cudaGetLastError(), "ffm_extract launch");
}

// high * fea (fp32) of both modalities in one launch: the only part of the Fourier split the pattern path consumes.
int ffm_highmul_pair_launch(const void *vis, const void *ir, float *hm_vis, float *hm_ir, int BC, int H, int W, int dtype,
                            cudaStream_t st) {
    if (H > kFfmMax || W > kFfmMax || H < 1 || W < 1) {
        set_error("mmi_ffm_pattern_fwd: pooled map must be within [1, %d]^2 (got %dx%d)", kFfmMax, H, W);
        return MMI_ERR_UNSUPPORTED;
    }
    int r0, r1, c0, c1;
    ffm_kept_range(H, W, &r0, &r1, &c0, &c1);
    const size_t smem = (size_t(3) * H * W + 2 * (H + W)) * sizeof(float);  // <= 48 KB whenever H*W <= 128
    const dim3 grid(BC, 2);
#define MMI_FFM_PAIR(T)                                                                                                 \
    ffm_extract_kernel<T><<<grid, 128, smem, st>>>(static_cast<const T *>(vis), nullptr, nullptr, hm_vis, H, W, r0, r1, c0, \
                                                  c1, static_cast<const T *>(ir), hm_ir)
    switch (dtype) {
        case MMI_F32: MMI_FFM_PAIR(float); break;
        case MMI_BF16: MMI_FFM_PAIR(__nv_bfloat16); break;
        case MMI_F16: MMI_FFM_PAIR(__half); break;
        default: set_error("mmi_ffm_pattern_fwd: unknown dtype %d", dtype); return MMI_ERR_ARG;
    }
#undef MMI_FFM_PAIR
    return check_cuda(cudaGetLastError(), "ffm high-pass product launch");
}

int separation_loss_launch(const float *M, float *loss, int l, int K, cudaStream_t st) {
    separation_loss_kernel<<<1, 256, 0, st>>>(M, loss, l, K);
    return check_cuda(cudaGetLastError(), "separation_loss launch");
}

}  // namespace mmi

// resample.cu -- the two resampling steps either side of the FFM token path:
//   AdaptiveAvgPool2d((vert_anchors, horz_anchors)) of the (B, C, H, W) maps   (models/common.py:324-325, :396-397)
//   F.interpolate(size=(H, W), mode='bilinear') of the (B, C, 8, 8) outputs     (models/common.py:540-543)
// Both are separable linear maps between a big grid (H, W) and a small anchor grid (hs, ws) in which every big-grid
// index touches at most two small-grid indices (a pooling window pair / the two bilinear taps):
//     reduce:  small[i, j] = sum_{y, x} Wy[i, y] Wx[j, x] big[y, x]     (pool forward, upsample backward)
//     expand:  big[y, x]   = sum_{i, j} Wy[i, y] Wx[j, x] small[i, j]   (upsample forward, pool backward)
// so two kernels cover the four directions; each streams the big map through HBM exactly once (a CTA works on one
// (b, c) image at a time, fp32 accumulation, fixed summation order -- no atomics).  The stock backward of the bilinear upsample
// scatters 25600 gradients per image onto 64 addresses with global atomics (12 ms at B=16, C=128, 160x160 fp32 on
// B200, profiles/r01_ffm_module.txt); the reduce kernel reads the gradient once.
#include "../../include/mmidet_b200.h"
#include "common.cuh"

namespace mmi {

int sm_count();

constexpr int kRsThreads = 256;
constexpr int kRsRows = 16;        // big-grid rows per tile in the reduce kernel
constexpr int kRsMaxSmall = 256;   // hs * ws
constexpr int kRsMaxBig = 2048;    // H, W

enum { RS_POOL = 0, RS_BILINEAR = 1 };

// taps of big-grid index `p` (extent n) on the small grid (extent ns): indices i0 <= i1 and weights w0, w1.
__device__ __forceinline__ void taps(int mode, int p, int n, int ns, int &i0, int &i1, float &w0, float &w1) {
    if (mode == RS_POOL) {
        // adaptive windows [floor(j n / ns), ceil((j + 1) n / ns)): p lies in windows lo..hi, hi - lo <= 1 for n >= ns
        // (32-bit arithmetic: n <= 2048 and ns <= 256 keep every product below 2^20)
        const unsigned un = n, uns = ns, up = p;
        const unsigned lo = (up * uns) / un, hi = ((up + 1) * uns - 1) / un;
        auto inv_len = [&](unsigned j) {
            const unsigned s = (j * un) / uns, e = ((j + 1) * un + uns - 1) / uns;
            return 1.0f / float(e - s);
        };
        i0 = int(lo);
        i1 = int(hi);
        w0 = inv_len(lo);
        w1 = hi != lo ? inv_len(hi) : 0.f;
    } else {
        // align_corners=False: src = (p + 0.5) ns / n - 0.5, clamped at 0 (ATen area_pixel_compute_source_index)
        float src = (float(p) + 0.5f) * (float(ns) / float(n)) - 0.5f;
        src = src < 0.f ? 0.f : src;
        i0 = min(int(src), ns - 1);
        i1 = i0 + (i0 < ns - 1 ? 1 : 0);
        w1 = src - float(i0);
        w0 = 1.0f - w1;
    }
}

template <typename T> __device__ __forceinline__ float4 load4(const T *p);
template <> __device__ __forceinline__ float4 load4<float>(const float *p) { return *reinterpret_cast<const float4 *>(p); }
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const uint2 r = *reinterpret_cast<const uint2 *>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 load4<__half>(const __half *p) {
    const uint2 r = *reinterpret_cast<const uint2 *>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void store4(T *p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
    uint2 r;
    *reinterpret_cast<__nv_bfloat162 *>(&r.x) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162 *>(&r.y) = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2 *>(p) = r;
}
template <> __device__ __forceinline__ void store4<__half>(__half *p, float4 v) {
    uint2 r;
    *reinterpret_cast<__half2 *>(&r.x) = __floats2half2_rn(v.x, v.y);
    *reinterpret_cast<__half2 *>(&r.y) = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2 *>(p) = r;
}

struct __align__(16) Tap {
    int i0, i1;
    float w0, w1;
};

// per-index taps in shared memory, one 16-byte entry each (a single broadcast load per row / column)
struct TapTable {
    Tap *t;
    __device__ void carve(float *&p, int n) {
        t = reinterpret_cast<Tap *>(p);
        p += 4 * n;
    }
    __device__ void fill(int mode, int n, int ns) {
        for (int p = threadIdx.x; p < n; p += kRsThreads) {
            Tap e;
            taps(mode, p, n, ns, e.i0, e.i1, e.w0, e.w1);
            t[p] = e;
        }
    }
    __device__ __forceinline__ Tap at(int p) const { return t[p]; }
    // an index at the border has both bilinear taps on one cell: the two weights add up
    __device__ __forceinline__ float weight(int p, int j) const {
        const Tap e = t[p];
        return (e.i0 == j ? e.w0 : 0.f) + (e.i1 == j ? e.w1 : 0.f);
    }
};

// small[i, j] = sum Wy[i, y] Wx[j, x] big[y, x]; one CTA per image.  General widths: row tiles staged in shared memory.
template <typename T>
__global__ void __launch_bounds__(kRsThreads)
    resample_reduce_kernel(const T *__restrict__ big, T *__restrict__ small, int H, int W, int hs, int ws, int mode) {
    extern __shared__ float sm[];
    float *p = sm;
    TapTable tx, ty;
    tx.carve(p, W);
    ty.carve(p, H);
    int *xlo = reinterpret_cast<int *>(p);  // support [xlo[j], xhi[j]] of column bin j
    int *xhi = xlo + ws;
    p += 2 * ws;
    float *acc = p;        // [hs][ws]
    p += hs * ws;
    float *Tr = p;         // [kRsRows][ws]
    p += kRsRows * ws;
    p = sm + (((p - sm) + 3) & ~3);  // 16-byte aligned tile rows
    const int WP = ((W + 3) & ~3) + 4;
    float *tile = p;       // [kRsRows][WP]
    const int tid = threadIdx.x;
    const T *img = big + int64_t(blockIdx.x) * H * W;

    tx.fill(mode, W, ws);
    ty.fill(mode, H, hs);
    for (int i = tid; i < hs * ws; i += kRsThreads) acc[i] = 0.f;
    __syncthreads();
    for (int j = tid; j < ws; j += kRsThreads) {
        int lo = W, hi = -1;
        for (int x = 0; x < W; ++x)
            if (tx.at(x).i0 == j || tx.at(x).i1 == j) {
                lo = min(lo, x);
                hi = max(hi, x);
            }
        xlo[j] = lo;
        xhi[j] = hi;
    }
    for (int y0 = 0; y0 < H; y0 += kRsRows) {
        const int nr = min(kRsRows, H - y0);
        const T *src = img + int64_t(y0) * W;
        if ((W & 3) == 0) {  // images start 4-element aligned whenever W % 4 == 0
            const int W4 = W >> 2;
            for (int i = tid; i < nr * W4; i += kRsThreads)
                *reinterpret_cast<float4 *>(tile + (i / W4) * WP + 4 * (i % W4)) = load4<T>(src + 4 * i);
        } else {
            for (int i = tid; i < nr * W; i += kRsThreads) tile[(i / W) * WP + i % W] = to_f32<T>(src[i]);
        }
        __syncthreads();
        for (int i = tid; i < nr * ws; i += kRsThreads) {
            const int r = i / ws, j = i % ws;
            float s = 0.f;
            for (int x = xlo[j]; x <= xhi[j]; ++x) s = fmaf(tx.weight(x, j), tile[r * WP + x], s);
            Tr[r * ws + j] = s;
        }
        __syncthreads();
        for (int i = tid; i < hs * ws; i += kRsThreads) {
            const int a = i / ws, j = i % ws;
            float s = acc[i];
            for (int r = 0; r < nr; ++r) s = fmaf(ty.weight(y0 + r, a), Tr[r * ws + j], s);
            acc[i] = s;
        }
        // the next tile's loads do not touch Tr / acc; its first barrier orders them against this tile's readers
    }
    __syncthreads();
    T *dst = small + int64_t(blockIdx.x) * hs * ws;
    for (int i = tid; i < hs * ws; i += kRsThreads) dst[i] = from_f32<T>(acc[i]);
}

// Same map for W % 4 == 0, W <= 1024: vertical pass first, in registers.  Thread (rg, q) owns the four columns 4q..4q+3
// of the row range rg and walks down them with 16-byte loads; the anchor rows a row touches are i0(y) <= i1(y) <=
// i0(y) + 1 and never decrease, so two running sums per column suffice, flushed to Vp[rg][a][x] when i0 advances.
// Then V = sum_rg Vp and the horizontal taps, four threads per output cell.
constexpr int kRsUnroll = 8;

template <typename T>
__global__ void __launch_bounds__(kRsThreads, 3)
    resample_reduce_rows_kernel(const T *__restrict__ big, T *__restrict__ small, int BC, int H, int W, int hs, int ws,
                                int mode, int RG) {
    extern __shared__ float sm[];
    float *p = sm;
    TapTable tx, ty;
    tx.carve(p, W);
    ty.carve(p, H);
    int *xlo = reinterpret_cast<int *>(p);
    int *xhi = xlo + ws;
    p += 2 * ws;
    p = sm + (((p - sm) + 3) & ~3);
    float *Vp = p;  // [RG][hs][W]; slice 0 ends up holding the sum over rg
    const int tid = threadIdx.x, W4 = W >> 2;

    tx.fill(mode, W, ws);
    ty.fill(mode, H, hs);
    for (int j = tid; j < ws; j += kRsThreads) xlo[j] = 0, xhi[j] = -1;  // a bin no column touches stays empty
    __syncthreads();
    // support [xlo[j], xhi[j]] of column bin j: the columns touching a bin are contiguous, so each end has one writer
    for (int x = tid; x < W; x += kRsThreads) {
        const Tap e = tx.at(x);
        const int bins[2] = {e.i0, e.i1};
        for (int k = 0; k < 2; ++k) {
            const int bn = bins[k];
            if (x == 0 || (tx.at(x - 1).i0 != bn && tx.at(x - 1).i1 != bn)) xlo[bn] = x;
            if (x == W - 1 || (tx.at(x + 1).i0 != bn && tx.at(x + 1).i1 != bn)) xhi[bn] = x;
        }
    }
    // tables are image-independent: a CTA builds them once and walks over images
    for (int im = blockIdx.x; im < BC; im += gridDim.x) {
    const T *img = big + int64_t(im) * H * W;
    if (tid < RG * W4) {
        const int rg = tid / W4, q = tid % W4;
        const int ya = int(unsigned(rg) * unsigned(H) / unsigned(RG)), yb = int(unsigned(rg + 1) * unsigned(H) / unsigned(RG));
        float *vp = Vp + int64_t(rg) * hs * W + 4 * q;
        for (int a = 0; a < hs; ++a) *reinterpret_cast<float4 *>(vp + a * W) = make_float4(0.f, 0.f, 0.f, 0.f);
        auto flush = [&](int a, const float4 &v) {
            if (a >= 0 && a < hs) {
                float4 *d = reinterpret_cast<float4 *>(vp + a * W);
                float4 o = *d;
                o.x += v.x, o.y += v.y, o.z += v.z, o.w += v.w;
                *d = o;
            }
        };
        auto axpy = [](float4 &acc, float w, const float4 &v) {
            acc.x = fmaf(w, v.x, acc.x), acc.y = fmaf(w, v.y, acc.y), acc.z = fmaf(w, v.z, acc.z), acc.w = fmaf(w, v.w, acc.w);
        };
        int cur = -1;
        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
        const float4 zero = lo;
        for (int y = ya; y < yb; y += kRsUnroll) {
            float4 v[kRsUnroll];
#pragma unroll
            for (int u = 0; u < kRsUnroll; ++u)
                if (y + u < yb) v[u] = load4<T>(img + int64_t(y + u) * W + 4 * q);
#pragma unroll
            for (int u = 0; u < kRsUnroll; ++u) {
                if (y + u >= yb) break;
                const Tap e = ty.at(y + u);
                const int i0 = e.i0, i1 = e.i1;
                if (i0 != cur) {  // warp-uniform: depends on the row only
                    flush(cur, lo);
                    if (i0 == cur + 1) {
                        lo = hi;
                    } else {
                        flush(cur + 1, hi);
                        lo = zero;
                    }
                    hi = zero;
                    cur = i0;
                }
                axpy(lo, e.w0, v[u]);
                if (i1 == i0) axpy(lo, e.w1, v[u]);
                else axpy(hi, e.w1, v[u]);
            }
        }
        flush(cur, lo);
        flush(cur + 1, hi);
    }
    __syncthreads();
    for (int i = tid; i < hs * W4; i += kRsThreads) {
        float4 a = reinterpret_cast<const float4 *>(Vp)[i];
        for (int rg = 1; rg < RG; ++rg) {
            const float4 b = reinterpret_cast<const float4 *>(Vp + int64_t(rg) * hs * W)[i];
            a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        }
        reinterpret_cast<float4 *>(Vp)[i] = a;
    }
    __syncthreads();
    T *dst = small + int64_t(im) * hs * ws;
    for (int c0 = 0; c0 < hs * ws; c0 += kRsThreads / 4) {
        const int cell = c0 + tid / 4, part = tid % 4;
        float sacc = 0.f;
        int a = 0, j = 0;
        if (cell < hs * ws) {
            a = cell / ws, j = cell % ws;
            for (int x = xlo[j] + part; x <= xhi[j]; x += 4) sacc = fmaf(tx.weight(x, j), Vp[a * W + x], sacc);
        }
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        if (cell < hs * ws && part == 0) dst[a * ws + j] = from_f32<T>(sacc);
    }
    __syncthreads();  // slice 0 of Vp is re-zeroed by the next image
    }
}

// Same map for images of 32 KB and more: one CTA per SM streams its images, cut into row bands, through a four-slot
// shared-memory ring filled by 1-D bulk copies (cp.async.bulk, one mbarrier per slot), so three bands are in flight
// while one is reduced out of shared memory -- the per-image tail no longer idles the memory system.  Vertical taps
// first (thread = anchor row x four columns, 16-byte shared loads), then the horizontal taps as above.
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kRingThreads = 512;  // measured: 256 threads 0.060 ms (bilinear), 512 0.053 ms, 1024 with split supports 0.072 ms

constexpr int kRingSlots = 4;

// Units of work are row bands (NH per image, RBn rows each): with four slots three bands are in flight while one is
// reduced, which is what keeps one CTA per SM close to its share of the HBM bandwidth.
template <typename T>
__global__ void __launch_bounds__(kRingThreads, 1)
    resample_reduce_ring_kernel(const T *__restrict__ big, T *__restrict__ small, int BC, int H, int W, int hs, int ws,
                                int mode, int NH, int RBn, uint32_t slot_stride) {
    extern __shared__ __align__(128) float sm[];
    float *p = sm;
    TapTable tx, ty;
    tx.carve(p, W);
    ty.carve(p, H);
    int *xlo = reinterpret_cast<int *>(p), *xhi = xlo + ws, *ylo = xhi + ws, *yhi = ylo + hs;
    p += 2 * ws + 2 * hs;
    p = sm + (((p - sm) + 3) & ~3);
    float *V = p;  // [hs][W]: vertical sums of the current image, accumulated band by band
    p += hs * W;
    float *WY = p;  // [hs][H] and [ws][W]: dense tap weights (image-independent, built once), so that the inner loops
    p += hs * H;    // are one broadcast weight load + one data load + FMAs
    float *WX = p;
    p += ws * W;
    p = sm + (((p - sm) + 3) & ~3);
    uint64_t *full = reinterpret_cast<uint64_t *>(p);
    p += 2 * kRingSlots;
    p = sm + (((p - sm) + 31) & ~31);  // 128-byte aligned slots
    unsigned char *slot0 = reinterpret_cast<unsigned char *>(p);
    const int tid = threadIdx.x, W4 = W >> 2;

    for (int i = tid; i < W; i += kRingThreads) taps(mode, i, W, ws, tx.t[i].i0, tx.t[i].i1, tx.t[i].w0, tx.t[i].w1);
    for (int i = tid; i < H; i += kRingThreads) taps(mode, i, H, hs, ty.t[i].i0, ty.t[i].i1, ty.t[i].w0, ty.t[i].w1);
    for (int j = tid; j < ws; j += kRingThreads) xlo[j] = 0, xhi[j] = -1;
    for (int a = tid; a < hs; a += kRingThreads) ylo[a] = 0, yhi[a] = -1;
    if (tid == 0) {
        for (int i = 0; i < kRingSlots; ++i) mbar_init(&full[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto ends = [](const TapTable &t, int n, int *lo, int *hi, int i) {  // contiguous supports: one writer per end
        const Tap e = t.at(i);
        const int bins[2] = {e.i0, e.i1};
        for (int k = 0; k < 2; ++k) {
            const int bn = bins[k];
            if (i == 0 || (t.at(i - 1).i0 != bn && t.at(i - 1).i1 != bn)) lo[bn] = i;
            if (i == n - 1 || (t.at(i + 1).i0 != bn && t.at(i + 1).i1 != bn)) hi[bn] = i;
        }
    };
    for (int x = tid; x < W; x += kRingThreads) ends(tx, W, xlo, xhi, x);
    for (int y = tid; y < H; y += kRingThreads) ends(ty, H, ylo, yhi, y);
    for (int i = tid; i < hs * H; i += kRingThreads) WY[i] = ty.weight(i % H, i / H);
    for (int i = tid; i < ws * W; i += kRingThreads) WX[i] = tx.weight(i % W, i / W);

    const int nimg = int(blockIdx.x) < BC ? (BC - 1 - int(blockIdx.x)) / int(gridDim.x) + 1 : 0;  // images of this CTA
    const int nunit = nimg * NH;
    auto issue = [&](int u) {  // unit u = band (u % NH) of this CTA's image (u / NH), into slot u % kRingSlots
        const int im = blockIdx.x + (u / NH) * gridDim.x, y0 = (u % NH) * RBn;
        const uint32_t bytes = uint32_t(min(RBn, H - y0)) * W * sizeof(T);
        uint64_t *bar = &full[u % kRingSlots];
        mbar_arrive_expect_tx(bar, bytes);
        bulk_load_1d(slot0 + size_t(u % kRingSlots) * slot_stride, big + (int64_t(im) * H + y0) * W, bytes, bar);
    };
    if (tid == 0)
        for (int u = 0; u < kRingSlots - 1 && u < nunit; ++u) issue(u);
    __syncthreads();
    const int parts = kRingThreads / (hs * ws) >= 32 ? 32 : (kRingThreads / (hs * ws) >= 16 ? 16 : (kRingThreads / (hs * ws) >= 8 ? 8 : 4));
    for (int u = 0; u < nunit; ++u) {
        // slot (u - 1) % kRingSlots was drained before the barrier that ended unit u - 1
        if (tid == 0 && u + kRingSlots - 1 < nunit) issue(u + kRingSlots - 1);
        const int band = u % NH, y0 = band * RBn, y1 = min(H, y0 + RBn);
        mbar_wait(&full[u % kRingSlots], (u / kRingSlots) & 1);
        const T *rowsp = reinterpret_cast<const T *>(slot0 + size_t(u % kRingSlots) * slot_stride) - int64_t(y0) * W;
        for (int task = tid; task < hs * W4; task += kRingThreads) {  // the same thread owns V[a][4q..] in every band
            const int a = task / W4, q = task % W4;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (band) acc = *reinterpret_cast<const float4 *>(V + a * W + 4 * q);
            const int ya = max(ylo[a], y0), yb = min(yhi[a] + 1, y1);
            const T *col = rowsp + 4 * q;
            const float *wy = WY + a * H;
            int y = ya;
            for (; y + 4 <= yb; y += 4) {  // four rows at a time: the shared-memory loads of a group are independent
                float w[4];
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    w[i] = wy[y + i];
                    v[i] = load4<T>(col + int64_t(y + i) * W);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    acc.x = fmaf(w[i], v[i].x, acc.x), acc.y = fmaf(w[i], v[i].y, acc.y), acc.z = fmaf(w[i], v[i].z, acc.z),
                    acc.w = fmaf(w[i], v[i].w, acc.w);
            }
            for (; y < yb; ++y) {
                const float w = wy[y];
                const float4 v = load4<T>(col + int64_t(y) * W);
                acc.x = fmaf(w, v.x, acc.x), acc.y = fmaf(w, v.y, acc.y), acc.z = fmaf(w, v.z, acc.z), acc.w = fmaf(w, v.w, acc.w);
            }
            *reinterpret_cast<float4 *>(V + a * W + 4 * q) = acc;
        }
        __syncthreads();  // the slot is free again; after the last band V is complete
        if (band == NH - 1) {
            const int im = blockIdx.x + (u / NH) * gridDim.x;
            T *dst = small + int64_t(im) * hs * ws;
            for (int c0 = 0; c0 < hs * ws; c0 += kRingThreads / parts) {
                const int cell = c0 + tid / parts, part = tid % parts;
                float sacc = 0.f;
                int a = 0, j = 0;
                if (cell < hs * ws) {
                    a = cell / ws, j = cell % ws;
                    for (int x = xlo[j] + part; x <= xhi[j]; x += parts) sacc = fmaf(WX[j * W + x], V[a * W + x], sacc);
                }
                for (int o = 1; o < parts; o <<= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
                if (cell < hs * ws && part == 0) dst[a * ws + j] = from_f32<T>(sacc);
            }
            __syncthreads();  // V is rewritten by the next image's first band
        }
    }
}

// big[y, x] = sum Wy[i, y] Wx[j, x] small[i, j]; one CTA per image: horizontal pass into shared memory, then rows.
template <typename T>
__global__ void __launch_bounds__(kRsThreads)
    resample_expand_kernel(const T *__restrict__ small, T *__restrict__ big, int BC, int H, int W, int hs, int ws, int mode) {
    extern __shared__ float sm[];
    float *p = sm;
    TapTable tx, ty;
    tx.carve(p, W);
    ty.carve(p, H);
    float *s = p;   // [hs][ws]
    p += hs * ws;
    p = sm + (((p - sm) + 3) & ~3);
    float *hx = p;  // [hs][W]: rows of the small grid expanded along x
    const int tid = threadIdx.x;
    tx.fill(mode, W, ws);
    ty.fill(mode, H, hs);
    for (int im = blockIdx.x; im < BC; im += gridDim.x) {
    const T *src = small + int64_t(im) * hs * ws;
    T *img = big + int64_t(im) * H * W;
    for (int i = tid; i < hs * ws; i += kRsThreads) s[i] = to_f32<T>(src[i]);
    __syncthreads();
    for (int i = tid; i < hs * W; i += kRsThreads) {
        const int a = i / W, x = i % W;
        const Tap e = tx.at(x);
        hx[i] = e.w0 * s[a * ws + e.i0] + e.w1 * s[a * ws + e.i1];
    }
    __syncthreads();
    if ((W & 3) == 0 && (W >> 2) <= kRsThreads) {
        // thread (rg, q) owns columns 4q..4q+3 of a row range; the two anchor rows stay in registers while i0(y) holds
        const int W4 = W >> 2, RG = min(kRsThreads / W4, H);
        if (tid < RG * W4) {
            const int rg = tid / W4, q = tid % W4;
            const int ya = int(unsigned(rg) * unsigned(H) / unsigned(RG)), yb = int(unsigned(rg + 1) * unsigned(H) / unsigned(RG));
            int c0 = -1, c1 = -1;
            float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
            const float *hq = hx + 4 * q;
            T *dst = img + int64_t(ya) * W + 4 * q;
#pragma unroll 4
            for (int y = ya; y < yb; ++y, dst += W) {
                const Tap e = ty.at(y);
                const int i0 = e.i0, i1 = e.i1;
                if (i0 != c0 || i1 != c1) {
                    u = *reinterpret_cast<const float4 *>(hq + i0 * W);
                    v = *reinterpret_cast<const float4 *>(hq + i1 * W);
                    c0 = i0, c1 = i1;
                }
                const float a = e.w0, b = e.w1;
                store4<T>(dst, make_float4(a * u.x + b * v.x, a * u.y + b * v.y, a * u.z + b * v.z, a * u.w + b * v.w));
            }
        }
    } else {
        for (int i = tid; i < H * W; i += kRsThreads) {
            const int y = i / W, x = i % W;
            const Tap e = ty.at(y);
            img[i] = from_f32<T>(e.w0 * hx[e.i0 * W + x] + e.w1 * hx[e.i1 * W + x]);
        }
    }
    __syncthreads();  // s / hx are rewritten for the next image
    }
}

static int check_resample(const char *fn, int BC, int H, int W, int hs, int ws, int mode) {
    if (BC < 1 || H < 1 || W < 1 || hs < 1 || ws < 1) { set_error("%s: sizes must be positive", fn); return MMI_ERR_ARG; }
    if (H > kRsMaxBig || W > kRsMaxBig || hs * ws > kRsMaxSmall) {
        set_error("%s: map up to %dx%d and anchor grid up to %d cells supported (got %dx%d, %dx%d)", fn, kRsMaxBig, kRsMaxBig,
                  kRsMaxSmall, H, W, hs, ws);
        return MMI_ERR_UNSUPPORTED;
    }
    if (mode == RS_POOL && (H < hs || W < ws)) {
        set_error("%s: adaptive pooling needs the map (%dx%d) at least as large as the anchor grid (%dx%d)", fn, H, W, hs, ws);
        return MMI_ERR_UNSUPPORTED;
    }
    return MMI_OK;
}

template <typename K>
static int opt_in_smem(K kern, size_t smem) {
    // Streaming kernels with no L1 reuse: give the carve-out to shared memory so the CTA count is register-bound.
    // Attributes are per device and sticky: set once per (kernel, device), and again only for a larger request.
    struct Seen { const void *fn; int dev; size_t bytes; };
    static thread_local Seen seen[32] = {};  // 2 kernels... x 3 dtypes x devices touched by this thread; overflow just re-sets
    int dev = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    const void *fn = reinterpret_cast<const void *>(kern);
    Seen *slot = &seen[0];
    for (Seen &c : seen) {
        if ((c.fn == fn && c.dev == dev) || c.fn == nullptr) { slot = &c; break; }
    }
    if (slot->fn != fn || slot->dev != dev) *slot = Seen{fn, dev, 0};
    size_t &have = slot->bytes;
    if (have == 0)
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared),
                               "resample carve-out attribute"))
            return e;
    if (smem > 48 * 1024 && smem > have)
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)), "resample smem attribute"))
            return e;
    if (smem > have || have == 0) have = smem > 0 ? smem : 1;
    return MMI_OK;
}

int resample_reduce_launch(const void *big, void *small, int BC, int H, int W, int hs, int ws, int mode, int dtype,
                           cudaStream_t st, const char *fn) {
    if (int e = check_resample(fn, BC, H, W, hs, ws, mode)) return e;
    const int W4 = W >> 2;
    int RG = W4 > 0 ? kRsThreads / W4 : 0;
    RG = RG > H ? H : RG;
    const size_t smem_rows = (size_t(4) * (H + W) + 2 * ws + 4 + size_t(RG) * hs * W) * sizeof(float);
    const bool rows = (W & 3) == 0 && RG >= 1 && smem_rows <= 160 * 1024;
    const size_t smem = rows ? smem_rows
                             : (size_t(4) * (H + W) + 2 * ws + hs * ws + kRsRows * ws + 4 + size_t(kRsRows) * (((W + 3) & ~3) + 4)) * sizeof(float);
    // ring variant: two whole images + tables in shared memory, one CTA per SM walking over its images
    const size_t esz = dtype == MMI_F32 ? 4 : 2;
    const size_t img_b = (size_t(H) * W * esz + 127) & ~size_t(127);
    // row bands: the fewest per image such that kRingSlots slots fit next to the tables (band bytes a multiple of 16)
    const size_t ring_fixed = (size_t(4) * (H + W) + 2 * ws + 2 * hs + 4 + size_t(hs) * W + size_t(hs) * H + size_t(ws) * W + 4 +
                               2 * kRingSlots + 32) * sizeof(float);
    const size_t slot_budget = ring_fixed < 227 * 1024 ? (227 * 1024 - ring_fixed) / kRingSlots : 0;
    int NH = 1, RBn = H;
    while (NH < H && ((size_t(RBn) * W * esz + 127) & ~size_t(127)) > slot_budget) {
        ++NH;
        RBn = (H + NH - 1) / NH;
        if (esz == 2 && (RBn & 1)) ++RBn;  // W % 4 == 0 only guarantees 8-byte rows
    }
    const size_t slot_stride = (size_t(RBn) * W * esz + 127) & ~size_t(127);
    NH = (H + RBn - 1) / RBn;
    const size_t smem_ring = ring_fixed + kRingSlots * slot_stride;
    const bool ring = (W & 3) == 0 && (size_t(RBn) * W * esz) % 16 == 0 && slot_stride <= slot_budget && img_b >= 32 * 1024 &&
                      BC >= 2 * sm_count();
#define MMI_RS_REDUCE(T)                                                                                                \
    do {                                                                                                                \
        if (ring) {                                                                                                     \
            auto kern = resample_reduce_ring_kernel<T>;                                                                 \
            if (int e = opt_in_smem(kern, smem_ring)) return e;                                                         \
            kern<<<min(BC, sm_count()), kRingThreads, smem_ring, st>>>(static_cast<const T *>(big), static_cast<T *>(small), BC, \
                                                                        H, W, hs, ws, mode, NH, RBn, uint32_t(slot_stride)); \
        } else if (rows) {                                                                                              \
            auto kern = resample_reduce_rows_kernel<T>;                                                                 \
            if (int e = opt_in_smem(kern, smem)) return e;                                                              \
            kern<<<min(BC, 3 * sm_count()), kRsThreads, smem, st>>>(static_cast<const T *>(big), static_cast<T *>(small), BC, H, W, hs, ws, mode, RG); \
        } else {                                                                                                        \
            auto kern = resample_reduce_kernel<T>;                                                                      \
            if (int e = opt_in_smem(kern, smem)) return e;                                                              \
            kern<<<BC, kRsThreads, smem, st>>>(static_cast<const T *>(big), static_cast<T *>(small), H, W, hs, ws, mode); \
        }                                                                                                               \
    } while (0)
    switch (dtype) {
        case MMI_F32: MMI_RS_REDUCE(float); break;
        case MMI_BF16: MMI_RS_REDUCE(__nv_bfloat16); break;
        case MMI_F16: MMI_RS_REDUCE(__half); break;
        default: set_error("%s: unknown dtype %d", fn, dtype); return MMI_ERR_ARG;
    }
#undef MMI_RS_REDUCE
    return check_cuda(cudaGetLastError(), fn);
}

int resample_expand_launch(const void *small, void *big, int BC, int H, int W, int hs, int ws, int mode, int dtype,
                           cudaStream_t st, const char *fn) {
    if (int e = check_resample(fn, BC, H, W, hs, ws, mode)) return e;
    const size_t smem = (size_t(4) * (H + W) + hs * ws + 4 + size_t(hs) * W) * sizeof(float);
    if (smem > 200 * 1024) { set_error("%s: anchor rows x map width too large for shared memory (%d x %d)", fn, hs, W); return MMI_ERR_UNSUPPORTED; }
#define MMI_RS_EXPAND(T)                                                                                                \
    do {                                                                                                                \
        auto kern = resample_expand_kernel<T>;                                                                          \
        if (int e = opt_in_smem(kern, smem)) return e;                                                                  \
        kern<<<BC, kRsThreads, smem, st>>>(static_cast<const T *>(small), static_cast<T *>(big), BC, H, W, hs, ws, mode);  \
    } while (0)
    switch (dtype) {
        case MMI_F32: MMI_RS_EXPAND(float); break;
        case MMI_BF16: MMI_RS_EXPAND(__nv_bfloat16); break;
        case MMI_F16: MMI_RS_EXPAND(__half); break;
        default: set_error("%s: unknown dtype %d", fn, dtype); return MMI_ERR_ARG;
    }
#undef MMI_RS_EXPAND
    return check_cuda(cudaGetLastError(), fn);
}

}  // namespace mmi

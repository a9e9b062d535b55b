// ffm_pattern.cu -- the pattern path of the Fusion Focus Module between the 8x8 pooling and the transformer
// (GPT1_fourier.forward, models/common.py:440-503).
//
// Per modality m in {VIS, IR} and pooled map fea (B, C, P), P = vert_anchors * horz_anchors:
//     M   = sigmoid(conv1(fea))          conv1: 1x1, C -> 8, no bias      (common.py:476-480)
//     Mh  = sigmoid(conv1(high * fea))   high-pass branch                 (common.py:440-455)
//     tok[b, m*P + p, c] = conv2(M)[b, c, p] * fea[b, c, p]               (common.py:496-516, conv2: 1x1, 8 -> C)
//     rows = [M_vis.view(-1,P); M_ir.view(-1,P); Mh_vis.view(-1,P)[:B]; Mh_ir.view(-1,P)[:B]]   (common.py:487-490)
// `rows` (18B, P) is what Seperation_loss consumes (separation_loss_kernel in ffm.cu).  The reference spends two
// dozen small launches per modality here; this is one launch (grid B x 2) forward and one + a finish backward.
// Only the first B of the 8B high-pass rows are ever used (len // 8, common.py:487), i.e. batch entries b < ceil(B/8):
// the launcher runs the Fourier split on those alone (ffm_highmul_pair_launch, ffm.cu) into the workspace hm (2, nbh, C, P).
// The backward differentiates the token path only: the pattern loss is detached by the reference
// (models/yolo_test.py:230, :268) and carries no gradient.
#include "../../include/mmidet_b200.h"
#include "common.cuh"

namespace mmi {

constexpr int kPatJ = 8;      // conv1 output channels, fixed by the reference (common.py:330)
constexpr int kPatMaxP = 128; // pooled positions per map (8 x 8 = 64 in the reference)
constexpr int kPatCT = 32;    // channel tile
constexpr int kPatThreads = 256;

// dst[j][p] = sigmoid(sum_c W1[j, c] src[c, p]); red: [G][8][P] scratch, G = 256 / P channel groups.
template <typename S>
__device__ __forceinline__ void conv1_sigmoid(const S *__restrict__ src, const float *__restrict__ W1, int C, int P,
                                              float *red, float *dst) {
    const int tid = threadIdx.x, G = kPatThreads / P;
    if (tid < G * P) {
        const int g = tid / P, p = tid % P;
        float acc[kPatJ];
#pragma unroll
        for (int j = 0; j < kPatJ; ++j) acc[j] = 0.f;
        for (int c = g; c < C; c += G) {
            const float v = to_f32<S>(src[int64_t(c) * P + p]);
#pragma unroll
            for (int j = 0; j < kPatJ; ++j) acc[j] = fmaf(__ldg(W1 + int64_t(j) * C + c), v, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < kPatJ; ++j) red[(g * kPatJ + j) * P + p] = acc[j];
    }
    __syncthreads();
    for (int i = tid; i < kPatJ * P; i += kPatThreads) {
        const int j = i / P, p = i % P;
        float s = 0.f;
        for (int g = 0; g < G; ++g) s += red[(g * kPatJ + j) * P + p];
        dst[j * (P + 1) + p] = 1.0f / (1.0f + expf(-s));
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kPatThreads)
    ffm_pattern_fwd_kernel(const T *__restrict__ fea_vis, const T *__restrict__ fea_ir, const float *__restrict__ hm_vis,
                           const float *__restrict__ hm_ir, const float *__restrict__ W1, const float *__restrict__ W2,
                           T *__restrict__ tok, float *__restrict__ rows, int B, int C, int P, int nbh) {
    extern __shared__ float sm[];
    float *Ms = sm;                          // [8][P+1]
    float *red = Ms + kPatJ * (P + 1);       // [G][8][P]
    float *ft = red + kPatJ * kPatThreads;   // [CT][P+1]
    const int b = blockIdx.x, m = blockIdx.y, tid = threadIdx.x;
    const T *fea = (m ? fea_ir : fea_vis) + int64_t(b) * C * P;

    if (b < nbh) {  // high-pass rows r = b*8 + j < B
        const float *hm = (m ? hm_ir : hm_vis) + int64_t(b) * C * P;
        conv1_sigmoid<float>(hm, W1, C, P, red, Ms);
        for (int i = tid; i < kPatJ * P; i += kPatThreads) {
            const int j = i / P, p = i % P, r = b * kPatJ + j;
            if (r < B) rows[(int64_t(16) * B + int64_t(m) * B + r) * P + p] = Ms[j * (P + 1) + p];
        }
        __syncthreads();
    }
    conv1_sigmoid<T>(fea, W1, C, P, red, Ms);
    for (int i = tid; i < kPatJ * P; i += kPatThreads) {
        const int j = i / P, p = i % P;
        rows[(int64_t(m) * kPatJ * B + int64_t(b) * kPatJ + j) * P + p] = Ms[j * (P + 1) + p];
    }
    T *tk = tok + (int64_t(b) * 2 * P + int64_t(m) * P) * C;
    for (int c0 = 0; c0 < C; c0 += kPatCT) {
        for (int i = tid; i < kPatCT * P; i += kPatThreads) {
            const int c = i / P, p = i % P;
            ft[c * (P + 1) + p] = c0 + c < C ? to_f32<T>(fea[int64_t(c0 + c) * P + p]) : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < kPatCT * P; i += kPatThreads) {
            const int c = i % kPatCT, p = i / kPatCT;
            if (c0 + c < C) {
                const float *w2 = W2 + int64_t(c0 + c) * kPatJ;
                float pt = 0.f;
#pragma unroll
                for (int j = 0; j < kPatJ; ++j) pt = fmaf(__ldg(w2 + j), Ms[j * (P + 1) + p], pt);
                tk[int64_t(p) * C + c0 + c] = from_f32<T>(pt * ft[c * (P + 1) + p]);
            }
        }
        __syncthreads();
    }
}

// Token-path backward.  ws: per CTA (b, m) partial sums [8][C] of dW1 then [C][8] of dW2.
template <typename T>
__global__ void __launch_bounds__(kPatThreads)
    ffm_pattern_bwd_kernel(const T *__restrict__ fea_vis, const T *__restrict__ fea_ir, const T *__restrict__ dtok,
                           const float *__restrict__ rows, const float *__restrict__ W1, const float *__restrict__ W2,
                           T *__restrict__ dfea_vis, T *__restrict__ dfea_ir, float *__restrict__ ws, int B, int C, int P) {
    extern __shared__ float sm[];
    const int PP = P + 1;
    float *Ms = sm;                   // [8][P+1]
    float *dMs = Ms + kPatJ * PP;     // [8][P+1]  dM, then dpre = dM M (1 - M)
    float *ft = dMs + kPatJ * PP;     // [CT][P+1] fea tile
    float *dt = ft + kPatCT * PP;     // [CT][P+1] dtok tile (pass 1: dtok * fea)
    const int b = blockIdx.x, m = blockIdx.y, tid = threadIdx.x;
    const T *fea = (m ? fea_ir : fea_vis) + int64_t(b) * C * P;
    T *dfea = (m ? dfea_ir : dfea_vis) + int64_t(b) * C * P;
    const T *dtk = dtok + (int64_t(b) * 2 * P + int64_t(m) * P) * C;
    float *wsb = ws + (int64_t(b) * 2 + m) * 2 * kPatJ * C;

    for (int i = tid; i < kPatJ * P; i += kPatThreads) {
        const int j = i / P, p = i % P;
        Ms[j * PP + p] = rows[(int64_t(m) * kPatJ * B + int64_t(b) * kPatJ + j) * P + p];
        dMs[j * PP + p] = 0.f;
    }
    __syncthreads();
    // pass 1: dPT = dtok * fea;  dM += W2^T dPT;  dW2 partial = dPT M^T
    for (int c0 = 0; c0 < C; c0 += kPatCT) {
        const int nc = min(kPatCT, C - c0);
        for (int i = tid; i < kPatCT * P; i += kPatThreads) {
            const int c = i / P, p = i % P;
            ft[c * PP + p] = c < nc ? to_f32<T>(fea[int64_t(c0 + c) * P + p]) : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < kPatCT * P; i += kPatThreads) {
            const int c = i % kPatCT, p = i / kPatCT;
            dt[c * PP + p] = c < nc ? to_f32<T>(dtk[int64_t(p) * C + c0 + c]) * ft[c * PP + p] : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < kPatJ * P; i += kPatThreads) {
            const int j = i / P, p = i % P;
            float s = 0.f;
            for (int c = 0; c < nc; ++c) s = fmaf(__ldg(W2 + int64_t(c0 + c) * kPatJ + j), dt[c * PP + p], s);
            dMs[j * PP + p] += s;
        }
        {
            const int c = tid / kPatJ, j = tid % kPatJ;
            if (c < nc) {
                float s = 0.f;
                for (int p = 0; p < P; ++p) s = fmaf(dt[c * PP + p], Ms[j * PP + p], s);
                wsb[int64_t(kPatJ) * C + int64_t(c0 + c) * kPatJ + j] = s;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < kPatJ * P; i += kPatThreads) {
        const int j = i / P, p = i % P;
        const float mv = Ms[j * PP + p];
        dMs[j * PP + p] *= mv * (1.0f - mv);
    }
    __syncthreads();
    // pass 2: dfea = dtok * conv2(M) + W1^T dpre;  dW1 partial = dpre fea^T
    for (int c0 = 0; c0 < C; c0 += kPatCT) {
        const int nc = min(kPatCT, C - c0);
        for (int i = tid; i < kPatCT * P; i += kPatThreads) {
            const int c = i / P, p = i % P;
            ft[c * PP + p] = c < nc ? to_f32<T>(fea[int64_t(c0 + c) * P + p]) : 0.f;
        }
        for (int i = tid; i < kPatCT * P; i += kPatThreads) {
            const int c = i % kPatCT, p = i / kPatCT;
            dt[c * PP + p] = c < nc ? to_f32<T>(dtk[int64_t(p) * C + c0 + c]) : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < kPatCT * P; i += kPatThreads) {
            const int c = i / P, p = i % P;
            if (c < nc) {
                const float *w2 = W2 + int64_t(c0 + c) * kPatJ;
                float pt = 0.f, back = 0.f;
#pragma unroll
                for (int j = 0; j < kPatJ; ++j) {
                    pt = fmaf(__ldg(w2 + j), Ms[j * PP + p], pt);
                    back = fmaf(__ldg(W1 + int64_t(j) * C + c0 + c), dMs[j * PP + p], back);
                }
                dfea[int64_t(c0 + c) * P + p] = from_f32<T>(fmaf(dt[c * PP + p], pt, back));
            }
        }
        {
            const int c = tid % kPatCT, j = tid / kPatCT;
            if (c < nc) {
                float s = 0.f;
                for (int p = 0; p < P; ++p) s = fmaf(dMs[j * PP + p], ft[c * PP + p], s);
                wsb[int64_t(j) * C + c0 + c] = s;
            }
        }
        __syncthreads();
    }
}

// dW1 (8, C) and dW2 (C, 8): fixed-order sum of the 2B per-CTA partials.
__global__ void __launch_bounds__(256) ffm_pattern_finish_kernel(const float *__restrict__ ws, float *__restrict__ dW1,
                                                                 float *__restrict__ dW2, int nparts, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, n = kPatJ * C;
    if (i >= 2 * n) return;
    float s = 0.f;
    for (int k = 0; k < nparts; ++k) s += ws[int64_t(k) * 2 * n + i];
    if (i < n) dW1[i] = s;
    else dW2[i - n] = s;
}

static int check_pattern(const char *fn, int B, int C, int P) {
    if (B < 1 || C < 1) { set_error("%s: B and C must be positive (B=%d C=%d)", fn, B, C); return MMI_ERR_ARG; }
    if (P < 1 || P > kPatMaxP) { set_error("%s: P = vert_anchors*horz_anchors must be in [1, %d] (got %d)", fn, kPatMaxP, P); return MMI_ERR_UNSUPPORTED; }
    return MMI_OK;
}

int ffm_highmul_pair_launch(const void *vis, const void *ir, float *hm_vis, float *hm_ir, int BC, int H, int W, int dtype,
                            cudaStream_t st);
int separation_loss_launch(const float *M, float *loss, int l, int K, cudaStream_t st);

// backward: per-CTA weight-gradient partials; forward: high * fea of the ceil(B/8) leading batch entries, both modalities
int64_t ffm_pattern_ws_bytes(int B, int C, int P) {
    const int64_t bwd = int64_t(B) * 2 * 2 * kPatJ * C, fwd = int64_t(2) * ((B + 7) / 8) * C * P;
    return (bwd > fwd ? bwd : fwd) * int64_t(sizeof(float));
}

int ffm_pattern_fwd_launch(const void *fea_vis, const void *fea_ir, const float *W1, const float *W2, void *tok, float *rows,
                           float *loss, float *ws, int B, int C, int H, int W, int dtype, cudaStream_t st) {
    // ws == nullptr: no high-pass branch (GPT1.forward, models/common.py:218-239: rows = [M_vis; M_ir], 16B of them)
    const int P = H * W, nbh = ws ? (B + 7) / 8 : 0;
    if (int e = check_pattern("mmi_ffm_pattern_fwd", B, C, P)) return e;
    float *hm_vis = ws, *hm_ir = ws ? ws + int64_t(nbh) * C * P : nullptr;
    if (nbh)
        if (int e = ffm_highmul_pair_launch(fea_vis, fea_ir, hm_vis, hm_ir, nbh * C, H, W, dtype, st)) return e;
    const size_t smem = (size_t(kPatJ + kPatCT) * (P + 1) + kPatJ * kPatThreads) * sizeof(float);
    const dim3 grid(B, 2);
#define MMI_PAT_FWD(T)                                                                                                  \
    ffm_pattern_fwd_kernel<T><<<grid, kPatThreads, smem, st>>>(static_cast<const T *>(fea_vis), static_cast<const T *>(fea_ir), \
                                                              hm_vis, hm_ir, W1, W2, static_cast<T *>(tok), rows, B, C, P, nbh)
    switch (dtype) {
        case MMI_F32: MMI_PAT_FWD(float); break;
        case MMI_BF16: MMI_PAT_FWD(__nv_bfloat16); break;
        case MMI_F16: MMI_PAT_FWD(__half); break;
        default: set_error("mmi_ffm_pattern_fwd: unknown dtype %d", dtype); return MMI_ERR_ARG;
    }
#undef MMI_PAT_FWD
    if (int e = check_cuda(cudaGetLastError(), "ffm_pattern_fwd launch")) return e;
    return loss ? separation_loss_launch(rows, loss, (ws ? 18 : 16) * B, P, st) : MMI_OK;
}

int ffm_pattern_bwd_launch(const void *fea_vis, const void *fea_ir, const void *dtok, const float *rows, const float *W1,
                           const float *W2, void *dfea_vis, void *dfea_ir, float *dW1, float *dW2, float *ws, int B, int C,
                           int P, int dtype, cudaStream_t st) {
    if (int e = check_pattern("mmi_ffm_pattern_bwd", B, C, P)) return e;
    const size_t smem = size_t(2 * kPatJ + 2 * kPatCT) * (P + 1) * sizeof(float);
    const dim3 grid(B, 2);
#define MMI_PAT_BWD(T)                                                                                                  \
    ffm_pattern_bwd_kernel<T><<<grid, kPatThreads, smem, st>>>(static_cast<const T *>(fea_vis), static_cast<const T *>(fea_ir), \
                                                              static_cast<const T *>(dtok), rows, W1, W2,               \
                                                              static_cast<T *>(dfea_vis), static_cast<T *>(dfea_ir), ws, B, C, P)
    switch (dtype) {
        case MMI_F32: MMI_PAT_BWD(float); break;
        case MMI_BF16: MMI_PAT_BWD(__nv_bfloat16); break;
        case MMI_F16: MMI_PAT_BWD(__half); break;
        default: set_error("mmi_ffm_pattern_bwd: unknown dtype %d", dtype); return MMI_ERR_ARG;
    }
#undef MMI_PAT_BWD
    if (int e = check_cuda(cudaGetLastError(), "ffm_pattern_bwd launch")) return e;
    const int n2 = 2 * kPatJ * C;
    ffm_pattern_finish_kernel<<<(n2 + 255) / 256, 256, 0, st>>>(ws, dW1, dW2, 2 * B, C);
    return check_cuda(cudaGetLastError(), "ffm_pattern_finish launch");
}

}  // namespace mmi

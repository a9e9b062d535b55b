// detect.cu -- input and post-processing either side of the detector.
//
//   u8_split_normalize   train.py:743-745 / detect_twostream.py:74-85: the loader's uint8 batch (B, 6, H, W) [RGB | IR
//                        stacked on the channel axis] -> imgs.float() / 255 -> two (B, 3, H, W) streams.  The reference
//                        spends a cast pass, a divide pass and (once a consumer wants contiguous maps) two slice copies;
//                        here one pass reads 1 byte and writes one element per pixel-channel.
//   detect_decode        Detect.forward, inference branch (models/yolo_test.py:47-68): per level, the 1x1-conv output
//                        (bs, na*no, ny, nx) -> raw (bs, na, ny, nx, no) [view + permute + contiguous] and the decoded
//                        rows y = sigmoid(x); xy = (2y - 0.5 + grid) * stride; wh = (2y)^2 * anchor, written straight
//                        into the concatenated (bs, sum na*ny*nx, no) prediction (torch.cat(z, 1)).
//   nms                  non_max_suppression (utils/general.py:486-580) for the whole batch in three launches instead of
//                        a Python loop over images: candidate filter + conf = obj * cls + best class + xywh -> xyxy +
//                        the class offset of :563-565; then, on candidates sorted by (image, -conf), the classic
//                        64-wide suppression bit matrix and one sequential sweep per image.  The IoU test is evaluated
//                        in the same fp32 arithmetic as torchvision.ops.nms (inter / (a + b - inter) > thr on the
//                        class-offset boxes), so the kept set is the reference's.
// All HBM-bound elementwise / tiny kernels; nothing here allocates.
#include <cstdint>

#include "../../include/mmidet_b200.h"
#include "common.cuh"

namespace mmi {

// ---- uint8 (B, 6, H, W) -> two float (B, 3, H, W) ----------------------------------------------------------------
template <typename T> __device__ __forceinline__ void st16(T *p, const float (&v)[16]);
template <> __device__ __forceinline__ void st16<float>(float *p, const float (&v)[16]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) reinterpret_cast<float4 *>(p)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
template <> __device__ __forceinline__ void st16<__half>(__half *p, const float (&v)[16]) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<const uint32_t *>(&h);
    }
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
template <> __device__ __forceinline__ void st16<__nv_bfloat16>(__nv_bfloat16 *p, const float (&v)[16]) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<const uint32_t *>(&h);
    }
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// one thread = 16 consecutive bytes of one (b, stream) block of 3*HW bytes; n3 = 3*HW must be a multiple of 16
template <typename T>
__global__ void __launch_bounds__(256) u8_split_kernel(const uint8_t *__restrict__ src, T *__restrict__ rgb, T *__restrict__ ir,
                                                       int64_t n3, int64_t nvec_total) {
    const int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nvec_total) return;
    const int64_t per = n3 / 16;           // vectors per (b, stream) block
    const int64_t blk = v / per, off = (v % per) * 16;
    const int64_t b = blk >> 1;
    const int m = int(blk & 1);
    const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(src + blk * n3 + off));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    float f[16];
#pragma unroll
    // torch divides a tensor by a Python scalar as a multiplication by the fp32 reciprocal (ATen div_true_kernel): same bits
    constexpr float kInv255 = 1.0f / 255.0f;
    for (int k = 0; k < 16; ++k) f[k] = float((w[k >> 2] >> (8 * (k & 3))) & 0xffu) * kInv255;
    st16<T>((m ? ir : rgb) + b * n3 + off, f);
}
template <typename T>
__global__ void __launch_bounds__(256) u8_split_scalar_kernel(const uint8_t *__restrict__ src, T *__restrict__ rgb,
                                                              T *__restrict__ ir, int64_t n3, int64_t total) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t blk = i / n3, off = i % n3;
    ((blk & 1) ? ir : rgb)[(blk >> 1) * n3 + off] = from_f32<T>(float(src[i]) * (1.0f / 255.0f));
}

int u8_split_launch(const void *src, void *rgb, void *ir, int B, int64_t HW, int dtype, cudaStream_t st) {
    const int64_t n3 = 3 * HW, total = int64_t(B) * 2 * n3;
    const bool vec = (n3 % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(rgb) |
                                         reinterpret_cast<uintptr_t>(ir)) & 15u) == 0;
#define MMI_U8(T)                                                                                                          \
    if (vec)                                                                                                                \
        u8_split_kernel<T><<<unsigned((total / 16 + 255) / 256), 256, 0, st>>>(static_cast<const uint8_t *>(src),            \
                                                                              static_cast<T *>(rgb), static_cast<T *>(ir), n3, \
                                                                              total / 16);                                   \
    else                                                                                                                    \
        u8_split_scalar_kernel<T><<<unsigned((total + 255) / 256), 256, 0, st>>>(static_cast<const uint8_t *>(src),          \
                                                                                static_cast<T *>(rgb), static_cast<T *>(ir), \
                                                                                n3, total)
    switch (dtype) {
        case MMI_F32: MMI_U8(float); break;
        case MMI_F16: MMI_U8(__half); break;
        case MMI_BF16: MMI_U8(__nv_bfloat16); break;
        default: set_error("mmi_u8_split_normalize: unknown dtype %d", dtype); return MMI_ERR_ARG;
    }
#undef MMI_U8
    return check_cuda(cudaGetLastError(), "u8_split_normalize launch");
}

// ---- Detect decode -------------------------------------------------------------------------------------------------
// x (bs, na*no, ny, nx) -> raw (bs, na, ny, nx, no) and z rows [row_off + (a*ny + y)*nx + x] of pred (bs, rows_total, no).
// A CTA handles 32 cells of one (b, a); the (no x 32) block crosses shared memory so both sides are coalesced.
template <typename T>
__global__ void __launch_bounds__(128) detect_decode_kernel(const T *__restrict__ x, T *__restrict__ raw, T *__restrict__ pred,
                                                            int na, int no, int ny, int nx, float stride, const float *__restrict__ anchor_wh,
                                                            int64_t rows_total, int64_t row_off) {
    extern __shared__ float tile[];  // [no][33]
    const int cells = ny * nx, c0 = blockIdx.x * 32, a = blockIdx.y, b = blockIdx.z;
    const T *xin = x + (int64_t(b) * na + a) * no * cells;
    for (int i = threadIdx.x; i < no * 32; i += blockDim.x) {
        const int o = i >> 5, c = c0 + (i & 31);
        if (c < cells) tile[o * 33 + (i & 31)] = to_f32<T>(xin[int64_t(o) * cells + c]);
    }
    __syncthreads();
    const float aw = anchor_wh[2 * a], ah = anchor_wh[2 * a + 1];
    T *r = raw ? raw + ((int64_t(b) * na + a) * cells + c0) * no : nullptr;
    T *z = pred + ((int64_t(b) * rows_total + row_off + int64_t(a) * cells + c0) * no);
    const int ncell = min(32, cells - c0);
    for (int i = threadIdx.x; i < ncell * no; i += blockDim.x) {
        const int cl = i / no, o = i % no, c = c0 + cl;
        const float v = tile[o * 33 + cl];
        if (r) r[i] = from_f32<T>(v);
        // the reference evaluates this chain in the tensor's dtype with fp32 grid / anchor tensors promoting the last two
        // steps (models/yolo_test.py:60-65): every intermediate is rounded to T exactly where torch rounds it
        auto rnd = [](float q) { return to_f32<T>(from_f32<T>(q)); };
        const float y = rnd(1.0f / (1.0f + __expf(-v)));
        float out = y;
        if (o < 2) out = (rnd(rnd(y * 2.0f) - 0.5f) + float(o == 0 ? c % nx : c / nx)) * stride;
        else if (o < 4) {
            const float t2 = rnd(y * 2.0f);
            out = rnd(t2 * t2) * (o == 2 ? aw : ah);
        }
        z[i] = from_f32<T>(out);
    }
}

int detect_decode_launch(const void *x, void *raw, void *pred, int bs, int na, int no, int ny, int nx, float stride,
                         const float *anchor_wh, int64_t rows_total, int64_t row_off, int dtype, cudaStream_t st) {
    const dim3 grid((ny * nx + 31) / 32, na, bs);
    const size_t smem = size_t(no) * 33 * sizeof(float);
#define MMI_DEC(T)                                                                                                      \
    detect_decode_kernel<T><<<grid, 128, smem, st>>>(static_cast<const T *>(x), static_cast<T *>(raw), static_cast<T *>(pred), \
                                                    na, no, ny, nx, stride, anchor_wh, rows_total, row_off)
    switch (dtype) {
        case MMI_F32: MMI_DEC(float); break;
        case MMI_F16: MMI_DEC(__half); break;
        case MMI_BF16: MMI_DEC(__nv_bfloat16); break;
        default: set_error("mmi_detect_decode: unknown dtype %d", dtype); return MMI_ERR_ARG;
    }
#undef MMI_DEC
    return check_cuda(cudaGetLastError(), "detect_decode launch");
}

// ---- batched NMS ---------------------------------------------------------------------------------------------------
// Stage 1: per prediction row -> candidate record (utils/general.py:494, :517-534): obj > conf_thres, conf = obj * max cls
// (first maximum, as torch.max), conf > conf_thres.  Rejected rows get key = +inf-like so that the sort puts them last.
//   det (rows, 6) = x1, y1, x2, y2, conf, cls;  key (rows) = image index - conf/2 for candidates (conf in (0, 1]: sorts by
//   image, then by descending confidence; fp64 so that no two confidences collapse), 1e30 otherwise.
template <typename T>
__global__ void __launch_bounds__(256) nms_candidates_kernel(const T *__restrict__ pred, float *__restrict__ det, double *__restrict__ key,
                                                             int *__restrict__ count, int64_t rows_per_img, int64_t rows, int no,
                                                             float conf_thres) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const T *p = pred + i * no;
    const float obj = to_f32<T>(p[4]);
    double k = 1e30;
    if (obj > conf_thres) {
        float best = -1.f;
        int bj = 0;
        for (int j = 5; j < no; ++j) {
            // x[:, 5:] *= x[:, 4:5] happens in the prediction's dtype (utils/general.py:517)
            const float c = to_f32<T>(from_f32<T>(to_f32<T>(p[j]) * obj));
            if (c > best) best = c, bj = j - 5;
        }
        if (best > conf_thres) {
            const float cx = to_f32<T>(p[0]), cy = to_f32<T>(p[1]), w = to_f32<T>(p[2]), h = to_f32<T>(p[3]);
            // xywh2xyxy in the prediction's dtype (utils/general.py:337-344), then the fp32 detections matrix
            float *d = det + i * 6;
            d[0] = to_f32<T>(from_f32<T>(cx - to_f32<T>(from_f32<T>(w / 2.f))));
            d[1] = to_f32<T>(from_f32<T>(cy - to_f32<T>(from_f32<T>(h / 2.f))));
            d[2] = to_f32<T>(from_f32<T>(cx + to_f32<T>(from_f32<T>(w / 2.f))));
            d[3] = to_f32<T>(from_f32<T>(cy + to_f32<T>(from_f32<T>(h / 2.f))));
            d[4] = best;
            d[5] = float(bj);
            const int img = int(i / rows_per_img);
            k = double(img) + 0.5 - double(best) * 0.5;  // in (img, img + 0.5): image-major, descending confidence (exact in fp64)
            atomicAdd(count + img, 1);
        }
    }
    key[i] = k;
}

// Stage 2: suppression bit matrix within one image.  Candidates of image `img` are order[start .. start + n) (sorted by
// descending confidence).  mask[(start + i) * words + w] bit j: box (w*64 + j) is suppressed by box i (j > i).
__device__ __forceinline__ bool nms_iou_gt(const float *a, const float *b, float thr) {
    const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    const float inter = width * height;
    const float sa = (a[2] - a[0]) * (a[3] - a[1]), sb = (b[2] - b[0]) * (b[3] - b[1]);
    return inter / (sa + sb - inter) > thr;  // torchvision's devIoU
}

__global__ void __launch_bounds__(64) nms_mask_kernel(const float *__restrict__ det, const int64_t *__restrict__ order,
                                                      const int *__restrict__ start, const int *__restrict__ count,
                                                      unsigned long long *__restrict__ mask, int words, int max_nms, float iou_thres,
                                                      float max_wh) {
    const int img = blockIdx.z, n = min(count[img], max_nms), s0 = start[img];
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (rb * 64 >= n || cb * 64 >= n || cb < rb) return;
    __shared__ float cbox[64][4];
    const int ncol = min(64, n - cb * 64);
    if (int(threadIdx.x) < ncol) {
        const float *d = det + order[s0 + cb * 64 + threadIdx.x] * 6;
        const float c = d[5] * max_wh;  // boxes + class offset (utils/general.py:563-565)
        cbox[threadIdx.x][0] = d[0] + c, cbox[threadIdx.x][1] = d[1] + c, cbox[threadIdx.x][2] = d[2] + c, cbox[threadIdx.x][3] = d[3] + c;
    }
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i >= n) return;
    const float *d = det + order[s0 + i] * 6;
    const float c = d[5] * max_wh;
    const float me[4] = {d[0] + c, d[1] + c, d[2] + c, d[3] + c};
    unsigned long long bits = 0;
    const int j0 = (rb == cb) ? int(threadIdx.x) + 1 : 0;
    for (int j = j0; j < ncol; ++j)
        if (nms_iou_gt(me, cbox[j], iou_thres)) bits |= 1ull << j;
    mask[(int64_t(s0) + i) * words + cb] = bits;
}

// Stage 3: one CTA per image sweeps its candidates in order; keep[] gets the indices (into det rows) of the kept boxes.
__global__ void __launch_bounds__(64) nms_sweep_kernel(const int64_t *__restrict__ order, const int *__restrict__ start,
                                                       const int *__restrict__ count, const unsigned long long *__restrict__ mask,
                                                       int words, int max_nms, int max_det, int64_t *__restrict__ keep,
                                                       int *__restrict__ nkeep) {
    extern __shared__ unsigned long long removed[];  // [words]
    const int img = blockIdx.x, n = min(count[img], max_nms), s0 = start[img];
    const int nw = (n + 63) / 64;
    for (int w = threadIdx.x; w < nw; w += blockDim.x) removed[w] = 0;
    __syncthreads();
    int kept = 0;
    for (int i = 0; i < n && kept < max_det; ++i) {
        const bool alive = !((removed[i >> 6] >> (i & 63)) & 1ull);  // uniform: every thread reads the same word
        if (alive) {
            if (threadIdx.x == 0) keep[int64_t(img) * max_det + kept] = order[s0 + i];
            ++kept;
            const unsigned long long *m = mask + (int64_t(s0) + i) * words;
            for (int w = (i >> 6) + threadIdx.x; w < nw; w += blockDim.x) removed[w] |= m[w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) nkeep[img] = kept;
}

int nms_candidates_launch(const void *pred, float *det, double *key, int *count, int bs, int64_t rows_per_img, int no,
                          float conf_thres, int dtype, cudaStream_t st) {
    const int64_t rows = int64_t(bs) * rows_per_img;
    if (int e = check_cuda(cudaMemsetAsync(count, 0, size_t(bs) * sizeof(int), st), "nms count memset")) return e;
    const unsigned grid = unsigned((rows + 255) / 256);
    switch (dtype) {
        case MMI_F32: nms_candidates_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(pred), det, key, count, rows_per_img, rows, no, conf_thres); break;
        case MMI_F16: nms_candidates_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half *>(pred), det, key, count, rows_per_img, rows, no, conf_thres); break;
        case MMI_BF16: nms_candidates_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(pred), det, key, count, rows_per_img, rows, no, conf_thres); break;
        default: set_error("mmi_nms_candidates: unknown dtype %d", dtype); return MMI_ERR_ARG;
    }
    return check_cuda(cudaGetLastError(), "nms candidates launch");
}

int nms_suppress_launch(const float *det, const int64_t *order, const int *start, const int *count, unsigned long long *mask,
                        int64_t *keep, int *nkeep, int bs, int max_count, int max_nms, int max_det, float iou_thres, float max_wh,
                        cudaStream_t st) {
    const int n = max_count < max_nms ? max_count : max_nms;
    const int words = (n + 63) / 64;
    if (words == 0) return check_cuda(cudaMemsetAsync(nkeep, 0, size_t(bs) * sizeof(int), st), "nms nkeep memset");
    nms_mask_kernel<<<dim3(words, words, bs), 64, 0, st>>>(det, order, start, count, mask, words, max_nms, iou_thres, max_wh);
    if (int e = check_cuda(cudaGetLastError(), "nms mask launch")) return e;
    nms_sweep_kernel<<<bs, 64, size_t(words) * 8, st>>>(order, start, count, mask, words, max_nms, max_det, keep, nkeep);
    return check_cuda(cudaGetLastError(), "nms sweep launch");
}

}  // namespace mmi

using namespace mmi;

extern "C" {

int mmi_u8_split_normalize(const void *imgs_u8, void *rgb, void *ir, int B, int H, int W, int dtype, void *stream) {
    if (!imgs_u8 || !rgb || !ir) { set_error("mmi_u8_split_normalize: null pointer"); return MMI_ERR_ARG; }
    if (B <= 0 || H <= 0 || W <= 0) { set_error("mmi_u8_split_normalize: bad shape B=%d H=%d W=%d", B, H, W); return MMI_ERR_ARG; }
    return u8_split_launch(imgs_u8, rgb, ir, B, int64_t(H) * W, dtype, static_cast<cudaStream_t>(stream));
}

int mmi_detect_decode(const void *x, void *raw, void *pred, int bs, int na, int no, int ny, int nx, float stride,
                      const float *anchor_wh, int64_t rows_total, int64_t row_off, int dtype, void *stream) {
    if (!x || !pred || !anchor_wh) { set_error("mmi_detect_decode: null pointer"); return MMI_ERR_ARG; }
    if (bs <= 0 || na <= 0 || no < 5 || ny <= 0 || nx <= 0 || bs > 65535 || na > 65535 || no > 1024) {
        set_error("mmi_detect_decode: bad shape bs=%d na=%d no=%d ny=%d nx=%d", bs, na, no, ny, nx);
        return MMI_ERR_ARG;
    }
    if (row_off < 0 || row_off + int64_t(na) * ny * nx > rows_total) { set_error("mmi_detect_decode: level does not fit in pred"); return MMI_ERR_ARG; }
    return detect_decode_launch(x, raw, pred, bs, na, no, ny, nx, stride, anchor_wh, rows_total, row_off, dtype,
                                static_cast<cudaStream_t>(stream));
}

int mmi_nms_candidates(const void *pred, float *det, double *key, int *count, int bs, int64_t rows_per_img, int no,
                       float conf_thres, int dtype, void *stream) {
    if (!pred || !det || !key || !count) { set_error("mmi_nms_candidates: null pointer"); return MMI_ERR_ARG; }
    if (bs <= 0 || rows_per_img <= 0 || no < 6) { set_error("mmi_nms_candidates: bad shape"); return MMI_ERR_ARG; }
    return nms_candidates_launch(pred, det, key, count, bs, rows_per_img, no, conf_thres, dtype, static_cast<cudaStream_t>(stream));
}

int mmi_nms_suppress(const float *det, const int64_t *order, const int *start, const int *count, void *mask, int64_t *keep,
                     int *nkeep, int bs, int max_count, int max_nms, int max_det, float iou_thres, float max_wh, void *stream) {
    if (!det || !order || !start || !count || !keep || !nkeep || (max_count > 0 && !mask)) { set_error("mmi_nms_suppress: null pointer"); return MMI_ERR_ARG; }
    if (bs <= 0 || bs > 65535 || max_det <= 0 || max_nms <= 0) { set_error("mmi_nms_suppress: bad arguments"); return MMI_ERR_ARG; }
    return nms_suppress_launch(det, order, start, count, static_cast<unsigned long long *>(mask), keep, nkeep, bs, max_count, max_nms,
                               max_det, iou_thres, max_wh, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

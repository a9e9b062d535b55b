/*
 * include/mmidet_b200.h -- C ABI of libmmidet_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for MMI-Det's cross-modal fusion hot path.  The reference is pure Python/PyTorch and has
 * no FFI of its own; every entry point below names the reference symbol it replaces.  The
 * Python host side (mmi-det_b200/*.py) binds these with ctypes and mirrors the reference's operator interface.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers valid on the device that owns `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream) unless the name ends in `_host`.
 *   - The caller (PyTorch) owns every buffer; the library never allocates or frees device memory, except the
 *     `_host` entry points which own a private, reusable staging workspace.
 *   - Return value: 0 = success; non-zero = error (MMI_ERR_*), message via mmi_last_error() (thread-local).
 *   - dtype: MMI_F32 / MMI_BF16 / MMI_F16 is the I/O element type of x, delta, z, B, C, out and their gradients;
 *     A, D, states and all accumulation are fp32.
 *   - Row pitches (`*_ld`) are in ELEMENTS; tensors (B, L, ED) are addressed as row (b*L + t), column d, so a
 *     chunk()/slice view along the channel axis is passed without a copy.  Pitch*sizeof(elem) and the base
 *     address must be multiples of 16 bytes (TMA bulk-copy requirement), ED a multiple of 8.
 *   - No CPU fallback exists: on a machine without an sm_100 device every compute entry returns MMI_ERR_CUDA.
 */
#ifndef MMIDET_B200_H
#define MMIDET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { MMI_F32 = 0, MMI_BF16 = 1, MMI_F16 = 2 };
enum { MMI_OK = 0, MMI_ERR_ARG = 1, MMI_ERR_CUDA = 2, MMI_ERR_UNSUPPORTED = 3 };

/* flags for the selective-scan entry points */
enum {
    MMI_FLAG_NO_GEOM = 1,      /* disable the geometric-A fast path (A[d,n] == (n+1)*A[d,0], the S4D-real init) */
    MMI_FLAG_DELTA_SOFTPLUS = 2, /* `delta` holds the pre-activation dt_proj(.) (models/mamba.py:203): the kernels apply softplus
                                  on load and the backward returns the gradient w.r.t. the pre-activation */
    MMI_FLAG_CFG_SHIFT = 4,    /* bits 4..7: kernel selection for tuning runs and tests; 0 = default dispatch (results do not   */
    MMI_FLAG_CFG_MASK = 0xF0,  /*            depend on it).  1..6: first-generation forward CTA shapes; 8: second-generation      */
                               /*            kernels (persistent grid over chained L segments) for forward and backward; 9: first  */
                               /*            generation for both; 10: as 8 with the 4-warp / two-CTA forward; 11: default forward, */
                               /*            16-warp second-generation backward (selscan_bwd3.cu).                                 */
    MMI_FLAG_NSEG_SHIFT = 8,   /* bits 8..15: force the number of L segments per sequence (first generation: published summaries, */
    MMI_FLAG_NSEG_MASK = 0xFF00 /*           1..32; second generation: chained hand-off, 1..16); 0 = heuristic                     */
};

const char *mmi_last_error(void);
int mmi_version(void);
/* number of SMs / compute capability of the current device; returns MMI_ERR_CUDA when no device is usable */
int mmi_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---------------------------------------------------------------------------------------------------------
 * Fused selective scan (+ optional SiLU gate).
 * Replaces MambaBlock.selective_scan (models/mamba.py:212-233: exp(delta*A), delta*B*x, pscan, hs@C, +D*x) and,
 * when z != NULL, the gate `y * silu(z)` of MambaBlock.forward (models/mamba.py:184-186).
 *   x, delta, z, out : (B, L, ED) dtype, row pitches x_ld, delta_ld, z_ld, out_ld
 *   A (ED, N) fp32 [= -exp(A_log), models/mamba.py:196], Bm, Cm (B, L, N) dtype contiguous, D (ED) fp32
 *   h0   (nullable) : (B, ED, N) fp32 initial state (reference: zeros, models/mamba.py:252)
 *   hT   (nullable) : (B, ED, N) fp32 final state h[L-1]
 *   chk  (nullable) : (B, ceil(L/chunk), ED, N) fp32; chk[b,j] = state entering step j*chunk.  Written for the
 *                     backward pass, which recomputes states chunk by chunk instead of materialising (B,L,ED,N).
 *   chunk           : checkpoint interval, must equal mmi_selscan_chunk() when chk != NULL
 *   ws   (nullable) : device workspace of mmi_selscan_fwd_ws_bytes() bytes.  With it, when B * ED alone cannot fill the
 *                     GPU (small batches, inference), L is cut into segments scanned by different CTAs of the same
 *                     launch: every segment first publishes (end state from zero, sum of delta), later segments chain
 *                     those summaries (decoupled look-back) and then scan for real.  Without it one CTA walks all of L.
 *   N must be 16 (the reference default d_state, models/mamba.py:35).
 * --------------------------------------------------------------------------------------------------------- */
int mmi_selscan_chunk(void);
int64_t mmi_selscan_fwd_ws_bytes(int B, int L, int ED, int N);
int mmi_selscan_fwd(const void *x, const void *delta, const void *z, const float *A, const void *Bm, const void *Cm,
                    const float *D, const float *h0, void *out, float *hT, float *chk, void *ws, int B, int L, int ED,
                    int N, int64_t x_ld, int64_t delta_ld, int64_t z_ld, int64_t out_ld, int chunk, int dtype, int flags,
                    void *stream);

/* Backward of the above.  Replaces autograd through models/mamba.py:222-231 and PScan.backward
 * (models/pscan.py:189-224: reverse scan with A shifted left, gradA = H[t-1]*G[t], gradX = G).
 *   dout, dx, ddelta, dz : (B, L, ED) dtype (pitches dout_ld, and ED-contiguous outputs), dz nullable iff z is
 *   dBm, dCm : (B, L, N) dtype; dA (ED, N) fp32; dD (ED) fp32   -- all OVERWRITTEN (not accumulated)
 *   chk      : checkpoints written by mmi_selscan_fwd with the same chunk
 *   ws       : device workspace of at least mmi_selscan_bwd_ws_bytes(B, L, ED, N) bytes (partial reductions) */
int64_t mmi_selscan_bwd_ws_bytes(int B, int L, int ED, int N);
int mmi_selscan_bwd(const void *x, const void *delta, const void *z, const float *A, const void *Bm, const void *Cm,
                    const float *D, const void *dout, const float *chk, void *dx, void *ddelta, void *dz, float *dA,
                    void *dBm, void *dCm, float *dD, void *ws, int B, int L, int ED, int N, int64_t x_ld,
                    int64_t delta_ld, int64_t z_ld, int64_t dout_ld, int chunk, int dtype, int flags, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Materialised linear recurrence  H[t] = A[t]*H[t-1] + X[t]  -- parity API for `pscan` (models/pscan.py:226).
 *   A, X, H : (B, L, D, N) fp32 contiguous.  Inputs are not modified (reference clones, pscan.py:167-174).
 * mmi_pscan_bwd replaces PScan.backward (models/pscan.py:189-224): gA[t] = H[t-1]*G[t] (gA[0]=0), gX = G with
 *   G[t] = gH[t] + A[t+1]*G[t+1].
 * One pass over the tensors: L is cut into 32-step segments scanned in parallel and chained by a decoupled look-back;
 * `ws` (>= mmi_pscan_ws_bytes bytes, contents arbitrary) holds the item ticket and the per-(segment, column) records
 * {product, local end state, inclusive state, status}.  Results are bit-reproducible.  Any D, N (only D*N matters).
 * --------------------------------------------------------------------------------------------------------- */
int64_t mmi_pscan_ws_bytes(int B, int L, int D, int N);
int mmi_pscan_fwd(const float *A, const float *X, float *H, void *ws, int B, int L, int D, int N, void *stream);
int mmi_pscan_bwd(const float *A, const float *H, const float *gH, float *gA, float *gX, void *ws, int B, int L, int D,
                  int N, void *stream);
/* fp64 variants of the two above (models/pscan.py is dtype-generic; fp64 is what its own gradcheck-style use needs) */
int mmi_pscan_fwd_f64(const double *A, const double *X, double *H, void *ws, int B, int L, int D, int N, void *stream);
int mmi_pscan_bwd_f64(const double *A, const double *H, const double *gH, double *gA, double *gX, void *ws, int B, int L,
                      int D, int N, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fusion Focus Module Fourier step.  Replaces extract_frequency2 (models/common.py:37-69):
 *   fftn -> fftshift -> rectangular masks (with the negative-slice wrap of :44-56) -> ifftshift -> ifftn ->
 *   real part -> fp16, for the low- and the high-pass image in one pass, plus the product high*fea used at
 *   models/common.py:440-441.  The box masks are separable, so the low-pass is evaluated directly as
 *   low = Re(Pr . x . Pc) with the small real/complex projection matrices of the kept bins; high = x - low
 *   whenever the two masks are complementary (always true for the reference's masks).
 *   img (BC, H, W) dtype contiguous; low, high (BC, H, W) fp16; high_mul (nullable) (BC, H, W) fp32 = high*img.
 *   H, W <= 64.
 * --------------------------------------------------------------------------------------------------------- */
int mmi_ffm_extract(const void *img, void *low, void *high, float *high_mul, int BC, int H, int W, int dtype,
                    void *stream);
/* Host helper: the [start, stop) row/column ranges of the SHIFTED spectrum that the reference's low-pass keeps
 * (== the block its high-pass zeroes), reproducing Python slice semantics for negative starts. */
void mmi_ffm_kept_range(int H, int W, int *r0, int *r1, int *c0, int *c1);

/* Separation loss (models/common.py:128-139) in closed form:
 *   sum_{i<j} M_i.M_j / (l(l-1)) = (|sum_i M_i|^2 - sum_i |M_i|^2) / (2 l (l-1)).   M (l, K) fp32 -> loss[0]. */
int mmi_separation_loss(const float *M, float *loss, int l, int K, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Pattern path of the Fusion Focus Module between the pooling and the transformer.  Replaces, in
 * GPT1_fourier.forward, models/common.py:440-455 (conv1 + sigmoid of high*fea), :476-480 (conv1 + sigmoid of the
 * pooled maps), :487-490 (the row set handed to Seperation_loss) and :496-516 (conv2(M) * fea, flatten, concat,
 * permute to tokens), for both modalities in one launch:
 *     M = sigmoid(W1 fea),  tok[b, m*P + p, c] = (W2 M)[c, p] * fea[b, c, p],  m = 0 (VIS), 1 (IR)
 *     rows = [M_vis (8B rows); M_ir (8B); sigmoid(W1 hm_vis).view(-1, P)[:B]; sigmoid(W1 hm_ir).view(-1, P)[:B]]
 *   fea_vis, fea_ir (B, C, H, W) dtype contiguous (the pooled maps, P = H*W = vert_anchors * horz_anchors <= 128);
 *   W1 (8, C) = conv1.weight, W2 (C, 8) = conv2.weight, fp32; tok (B, 2P, C) dtype; rows (18B, P) fp32;
 *   loss (nullable) fp32[1] = Seperation_loss(rows) (models/common.py:494).  hm = high * fea is needed only for the
 *   first ceil(B / 8) batch entries (len // 8 rows reach the loss, common.py:487); the call computes it into ws.
 *   Three launches: high-pass product (both modalities), pattern kernel (grid B x 2), separation loss.
 *   ws == NULL selects the sibling module GPT1.forward (models/common.py:196-239), which has no Fourier branch: the two
 *   high-pass row blocks are not produced and the loss runs over rows[:16B] = [M_vis; M_ir].
 * The backward differentiates the token path (the reference detaches the pattern loss, models/yolo_test.py:230):
 *   dtok (B, 2P, C) dtype -> dfea_vis, dfea_ir (B, C, P) dtype, dW1 (8, C), dW2 (C, 8) fp32 (overwritten; summed in a
 *   fixed order from per-CTA partials in ws).  rows is the forward's output.
 *   ws: mmi_ffm_pattern_ws_bytes(B, C, P) bytes, either direction.
 * --------------------------------------------------------------------------------------------------------- */
int64_t mmi_ffm_pattern_ws_bytes(int B, int C, int P);
int mmi_ffm_pattern_fwd(const void *fea_vis, const void *fea_ir, const float *W1, const float *W2, void *tok, float *rows,
                        float *loss, void *ws, int B, int C, int H, int W, int dtype, void *stream);
int mmi_ffm_pattern_bwd(const void *fea_vis, const void *fea_ir, const void *dtok, const float *rows, const float *W1,
                        const float *W2, void *dfea_vis, void *dfea_ir, float *dW1, float *dW2, void *ws, int B, int C,
                        int P, int dtype, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Resampling either side of the FFM token path.  Replace nn.AdaptiveAvgPool2d((vert_anchors, horz_anchors))
 * (models/common.py:324-325, applied at :396-397) and F.interpolate(size=(H, W), mode='bilinear') with the default
 * align_corners=False (models/common.py:540-543), forward and backward:
 *   big (BC, H, W), small (BC, hs, ws), both dtype contiguous; hs * ws <= 256, H, W <= 2048; pooling needs H >= hs, W >= ws.
 *   mmi_avgpool_fwd: big -> small;  mmi_avgpool_bwd: d small -> d big;
 *   mmi_upsample_bilinear_fwd: small -> big;  mmi_upsample_bilinear_bwd: d big -> d small.
 * One CTA per (b, c) image, the big map crosses HBM once, fp32 accumulation in a fixed order (no atomics).
 * --------------------------------------------------------------------------------------------------------- */
int mmi_avgpool_fwd(const void *big, void *small, int BC, int H, int W, int hs, int ws, int dtype, void *stream);
int mmi_avgpool_bwd(const void *dsmall, void *dbig, int BC, int H, int W, int hs, int ws, int dtype, void *stream);
int mmi_upsample_bilinear_fwd(const void *small, void *big, int BC, int H, int W, int hs, int ws, int dtype, void *stream);
int mmi_upsample_bilinear_bwd(const void *dbig, void *dsmall, int BC, int H, int W, int hs, int ws, int dtype, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Depthwise causal conv1d (+ bias) + SiLU on channels-last tokens.  Replaces the x-branch prologue of
 * MambaBlock.forward (models/mamba.py:176-180): transpose -> nn.Conv1d(ED, ED, K, groups=ED, padding=K-1)[..., :L]
 * -> transpose -> F.silu, without the transposes:
 *     y[t, d] = act(bias[d] + sum_{j<K} w[d, j] * x[t - (K-1) + j, d]),   x[t<0] = 0,   act = SiLU if silu else identity.
 *   x, y, dy, dx : (B, L, ED) dtype with row pitches in elements (x may be the first half of in_proj's output);
 *   w (ED, K) fp32 contiguous (= conv1d.weight (ED, 1, K), models/mamba.py:125); bias (ED) fp32, nullable; K in 1..4.
 * The backward overwrites dx, dw (ED, K) fp32 and dbias (ED) fp32 (nullable); dw / dbias are summed with fp32 atomics.
 * --------------------------------------------------------------------------------------------------------- */
int mmi_causal_conv1d_fwd(const void *x, const float *w, const float *bias, void *y, int B, int L, int ED, int K, int64_t x_ld,
                          int64_t y_ld, int dtype, int silu, void *stream);
int mmi_causal_conv1d_bwd(const void *x, const float *w, const float *bias, const void *dy, void *dx, float *dw, float *dbias,
                          int B, int L, int ED, int K, int64_t x_ld, int64_t dy_ld, int64_t dx_ld, int dtype, int silu,
                          void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * RMSNorm over the last axis of (rows, C) tokens.  Replaces RMSNorm.forward (models/mamba.py:356-366):
 *     y = x * rsqrt(mean(x^2, -1) + eps) * w          x, y, dy, dx : (rows, C) dtype with row pitches in elements; w (C) fp32
 * The backward overwrites dx and dw (C) fp32 (summed with fp32 atomics).  C a multiple of 8, C <= 1024.
 * y_dtype / dy_dtype: element type of y (forward) and dy (backward); -1 or `dtype` for the same type as x, or a 16-bit
 * type with fp32 x -- the autocast case, where the fp32 norm feeds a 16-bit GEMM: the kernel rounds once in its store
 * (identical to an fp32 result followed by a cast) and reads the 16-bit gradient directly, saving a pass each way.
 * --------------------------------------------------------------------------------------------------------- */
int mmi_rmsnorm_fwd(const void *x, const float *w, void *y, int64_t rows, int C, int64_t x_ld, int64_t y_ld, float eps, int dtype,
                    int y_dtype, void *stream);
int mmi_rmsnorm_bwd(const void *x, const float *w, const void *dy, void *dx, float *dw, int64_t rows, int C, int64_t x_ld,
                    int64_t dy_ld, int64_t dx_ld, float eps, int dtype, int dy_dtype, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Token layout either side of the fusion block.  Replaces flatten / cat / permute / contiguous of
 * models/common.py:1338-1343 and the split / reshape back of :1352-1366:
 *     gather : tok[b, m*HW + p, c] = (m == 0 ? rgb : ir)[b, c, p]      rgb, ir (B, C, HW) contiguous (NCHW maps)
 *     scatter: the inverse (and the adjoint: each is the other's backward)        tok (B, 2*HW, C) contiguous
 * Any 2- or 4-byte dtype; pure data movement through a shared-memory tile transpose.
 * --------------------------------------------------------------------------------------------------------- */
int mmi_tokens_gather(const void *rgb, const void *ir, void *tok, int B, int C, int HW, int dtype, void *stream);
int mmi_tokens_scatter(const void *tok, void *rgb, void *ir, int B, int C, int HW, int dtype, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Host-buffer entry (end-to-end path used by bench.py `e2e`): same maths as mmi_selscan_fwd followed by
 * mmi_selscan_bwd, with every pointer a HOST pointer (pinned memory recommended).  Copies inputs H2D, runs
 * forward + backward on an internal stream, copies out/dx/ddelta/dz/dB/dC/dA/dD back and synchronises.
 * The staging workspace is owned by the library and reused across calls; mmi_host_workspace_free releases it.
 * --------------------------------------------------------------------------------------------------------- */
int mmi_selscan_fwd_bwd_host(const void *x, const void *delta, const void *z, const float *A, const void *Bm,
                             const void *Cm, const float *D, const void *dout, void *out, void *dx, void *ddelta,
                             void *dz, float *dA, void *dBm, void *dCm, float *dD, int B, int L, int ED, int N,
                             int dtype, int flags);
void mmi_host_workspace_free(void);

/* ---------------------------------------------------------------------------------------------------------
 * Detector input / post-processing.
 * mmi_u8_split_normalize replaces `imgs.float() / 255.0; imgs[:, :3]; imgs[:, 3:]` (train.py:743-745,
 *   detect_twostream.py:74-85): imgs_u8 (B, 6, H, W) uint8 -> rgb, ir (B, 3, H, W) dtype, one pass.
 * mmi_detect_decode replaces the inference branch of Detect.forward for one level (models/yolo_test.py:47-68):
 *   x (bs, na*no, ny, nx) dtype [the 1x1 conv output] -> raw (bs, na, ny, nx, no) (nullable) and the decoded rows
 *   [row_off, row_off + na*ny*nx) of pred (bs, rows_total, no); anchor_wh (na, 2) fp32 = anchor_grid of the level.
 * mmi_nms_candidates + (caller sorts `key` ascending, stable -> order; start = exclusive prefix sum of count) +
 * mmi_nms_suppress replace non_max_suppression (utils/general.py:486-580, best-class branch) for the whole batch:
 *   pred (bs, rows_per_img, no) dtype; det (bs*rows_per_img, 6) fp32 = xyxy, conf, cls of the candidate rows;
 *   key (bs*rows_per_img) fp64; count (bs) int32; order int64 row indices sorted by key; mask scratch of
 *   total_candidates * ceil(min(max_count, max_nms) / 64) * 8 bytes; keep (bs, max_det) int64 row indices into det,
 *   nkeep (bs) int32.  max_wh is the class offset of :563 (4096). */
int mmi_u8_split_normalize(const void *imgs_u8, void *rgb, void *ir, int B, int H, int W, int dtype, void *stream);
int mmi_detect_decode(const void *x, void *raw, void *pred, int bs, int na, int no, int ny, int nx, float stride,
                      const float *anchor_wh, int64_t rows_total, int64_t row_off, int dtype, void *stream);
int mmi_nms_candidates(const void *pred, float *det, double *key, int *count, int bs, int64_t rows_per_img, int no,
                       float conf_thres, int dtype, void *stream);
int mmi_nms_suppress(const float *det, const int64_t *order, const int *start, const int *count, void *mask, int64_t *keep,
                     int *nkeep, int bs, int max_count, int max_nms, int max_det, float iou_thres, float max_wh, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MMIDET_B200_H */

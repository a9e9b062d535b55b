// capi.cu -- the extern "C" boundary declared in include/mmidet_b200.h: argument validation, error strings,
// dispatch to the kernel launchers.  No torch types, no allocation of caller-visible memory.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/mmidet_b200.h"
#include "common.cuh"
#include "selscan.h"

namespace mmi {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return MMI_OK;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return MMI_ERR_CUDA;
}

int sm_count() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0) {
        (void)cudaGetLastError();
        return 148;  // B200
    }
    cached = n;
    return n;
}

static int elem_size(int dtype) { return dtype == MMI_F32 ? 4 : (dtype == MMI_BF16 || dtype == MMI_F16) ? 2 : 0; }

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// every launch path funnels through here first: refuses to run on anything but sm_100 (no fallback of any kind)
static int require_device() {
    static int state = 0;  // 0 unknown, 1 ok
    if (state == 1) return MMI_OK;
    int dev = 0, major = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    if (int e = check_cuda(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev), "device attribute"))
        return e;
    if (major != 10) {
        set_error("libmmidet_b200 is built for sm_100a only; current device has compute capability major %d", major);
        return MMI_ERR_UNSUPPORTED;
    }
    state = 1;
    return MMI_OK;
}

static int check_scan_args(const char *who, int B, int L, int ED, int N, int dtype, const void *const *ptrs,
                           const int64_t *lds, int nptr) {
    const int es = elem_size(dtype);
    if (!es) { set_error("%s: unknown dtype %d", who, dtype); return MMI_ERR_ARG; }
    if (B <= 0 || L <= 0 || ED <= 0) { set_error("%s: B, L, ED must be positive (B=%d L=%d ED=%d)", who, B, L, ED); return MMI_ERR_ARG; }
    if (B > 65535) { set_error("%s: B=%d exceeds 65535", who, B); return MMI_ERR_ARG; }
    if (N != kN) { set_error("%s: d_state N=%d unsupported (this build keeps N=%d states in registers)", who, N, kN); return MMI_ERR_UNSUPPORTED; }
    if (ED % 8) { set_error("%s: ED=%d must be a multiple of 8", who, ED); return MMI_ERR_ARG; }
    for (int i = 0; i < nptr; ++i) {
        if (!ptrs[i]) continue;
        if (!aligned16(ptrs[i])) { set_error("%s: tensor %d is not 16-byte aligned", who, i); return MMI_ERR_ARG; }
        if (lds && lds[i] && ((lds[i] * es) % 16 || lds[i] < ED)) {
            set_error("%s: row pitch %lld of tensor %d must be >= ED and a multiple of 16 bytes", who, (long long)lds[i], i);
            return MMI_ERR_ARG;
        }
    }
    return MMI_OK;
}

int pscan_fwd_launch(const float *A, const float *X, float *H, float *ws, int B, int L, int DN, cudaStream_t st);
int pscan_bwd_launch(const float *A, const float *H, const float *gH, float *gA, float *gX, float *ws, int B, int L, int DN,
                     cudaStream_t st);
int pscan_fwd_launch_f64(const double *A, const double *X, double *H, double *ws, int B, int L, int DN, cudaStream_t st);
int pscan_bwd_launch_f64(const double *A, const double *H, const double *gH, double *gA, double *gX, double *ws, int B, int L,
                         int DN, cudaStream_t st);
int64_t pscan_ws_bytes(int B, int L, int D, int N);
int ffm_extract_launch(const void *img, void *low, void *high, float *high_mul, int BC, int H, int W, int dtype,
                       cudaStream_t st);
int separation_loss_launch(const float *M, float *loss, int l, int K, cudaStream_t st);
void ffm_kept_range(int H, int W, int *r0, int *r1, int *c0, int *c1);
int resample_reduce_launch(const void *big, void *small, int BC, int H, int W, int hs, int ws, int mode, int dtype,
                           cudaStream_t st, const char *fn);
int resample_expand_launch(const void *small, void *big, int BC, int H, int W, int hs, int ws, int mode, int dtype,
                           cudaStream_t st, const char *fn);
int64_t ffm_pattern_ws_bytes(int B, int C, int P);
int ffm_pattern_fwd_launch(const void *fea_vis, const void *fea_ir, const float *W1, const float *W2, void *tok, float *rows,
                           float *loss, float *ws, int B, int C, int H, int W, int dtype, cudaStream_t st);
int ffm_pattern_bwd_launch(const void *fea_vis, const void *fea_ir, const void *dtok, const float *rows, const float *W1,
                           const float *W2, void *dfea_vis, void *dfea_ir, float *dW1, float *dW2, float *ws, int B, int C,
                           int P, int dtype, cudaStream_t st);

}  // namespace mmi

using namespace mmi;

extern "C" {

const char *mmi_last_error(void) { return g_err; }
int mmi_version(void) { return 100; }

int mmi_device_info(int *sms, int *maj, int *min) {
    int dev = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    int a = 0, b = 0, c = 0;
    if (int e = check_cuda(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev), "device attribute")) return e;
    cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev);
    if (sms) *sms = a;
    if (maj) *maj = b;
    if (min) *min = c;
    return MMI_OK;
}

int mmi_selscan_chunk(void) { return kChunk; }

int64_t mmi_selscan_fwd_ws_bytes(int B, int L, int ED, int N) {
    (void)L; (void)N;
    if (B <= 0 || ED <= 0) return 0;
    return selscan_fwd_ws_bytes(B, ED);
}

int mmi_selscan_fwd(const void *x, const void *delta, const void *z, const float *A, const void *Bm, const void *Cm,
                    const float *D, const float *h0, void *out, float *hT, float *chk, void *ws, int B, int L, int ED,
                    int N, int64_t x_ld, int64_t delta_ld, int64_t z_ld, int64_t out_ld, int chunk, int dtype, int flags,
                    void *stream) {
    if (!x || !delta || !A || !Bm || !Cm || !D || !out) { set_error("mmi_selscan_fwd: null required pointer"); return MMI_ERR_ARG; }
    const void *ptrs[] = {x, delta, z, out, Bm, Cm, chk, hT, h0, ws};
    const int64_t lds[] = {x_ld, delta_ld, z ? z_ld : 0, out_ld, 0, 0, 0, 0, 0, 0};
    if (int e = check_scan_args("mmi_selscan_fwd", B, L, ED, N, dtype, ptrs, lds, 10)) return e;
    if (chk && chunk != kChunk) { set_error("mmi_selscan_fwd: chunk=%d, expected mmi_selscan_chunk()=%d", chunk, kChunk); return MMI_ERR_ARG; }
    if (int e = require_device()) return e;
    FwdParams p{};
    p.x = x; p.delta = delta; p.z = z; p.Bm = Bm; p.Cm = Cm; p.A = A; p.D = D; p.h0 = h0;
    p.out = out; p.hT = hT; p.chk = chk;
    p.B = B; p.L = L; p.ED = ED;
    p.x_ld = x_ld; p.d_ld = delta_ld; p.z_ld = z_ld; p.o_ld = out_ld;
    p.flags = flags;
    return selscan_fwd_launch(p, dtype, ws, static_cast<cudaStream_t>(stream));
}

int64_t mmi_selscan_bwd_ws_bytes(int B, int L, int ED, int N) {
    (void)N;
    if (B <= 0 || L <= 0 || ED <= 0) return 0;
    return selscan_bwd_ws_bytes(B, L, ED);
}

int mmi_selscan_bwd(const void *x, const void *delta, const void *z, const float *A, const void *Bm, const void *Cm,
                    const float *D, const void *dout, const float *chk, void *dx, void *ddelta, void *dz, float *dA,
                    void *dBm, void *dCm, float *dD, void *ws, int B, int L, int ED, int N, int64_t x_ld,
                    int64_t delta_ld, int64_t z_ld, int64_t dout_ld, int chunk, int dtype, int flags, void *stream) {
    if (!x || !delta || !A || !Bm || !Cm || !D || !dout || !chk || !dx || !ddelta || !dA || !dBm || !dCm || !dD || !ws) {
        set_error("mmi_selscan_bwd: null required pointer");
        return MMI_ERR_ARG;
    }
    if ((z == nullptr) != (dz == nullptr)) { set_error("mmi_selscan_bwd: z and dz must both be given or both be NULL"); return MMI_ERR_ARG; }
    const void *ptrs[] = {x, delta, z, dout, Bm, Cm, dx, ddelta, dz, dBm, dCm, chk, ws};
    const int64_t lds[] = {x_ld, delta_ld, z ? z_ld : 0, dout_ld, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (int e = check_scan_args("mmi_selscan_bwd", B, L, ED, N, dtype, ptrs, lds, 13)) return e;
    if (chunk != kChunk) { set_error("mmi_selscan_bwd: chunk=%d, expected mmi_selscan_chunk()=%d", chunk, kChunk); return MMI_ERR_ARG; }
    if (int e = require_device()) return e;
    BwdParams p{};
    p.x = x; p.delta = delta; p.z = z; p.Bm = Bm; p.Cm = Cm; p.dout = dout; p.A = A; p.D = D; p.chk = chk;
    p.dx = dx; p.ddelta = ddelta; p.dz = dz; p.dBm = dBm; p.dCm = dCm; p.dA = dA; p.dD = dD;
    p.B = B; p.L = L; p.ED = ED;
    p.x_ld = x_ld; p.d_ld = delta_ld; p.z_ld = z_ld; p.g_ld = dout_ld;
    p.flags = flags;
    return selscan_bwd_launch(p, dtype, ws, static_cast<cudaStream_t>(stream));
}

int64_t mmi_pscan_ws_bytes(int B, int L, int D, int N) {
    if (B <= 0 || L <= 0 || D <= 0 || N <= 0) return 0;
    return pscan_ws_bytes(B, L, D, N);
}

static int check_pscan(const char *who, const void *ws, int B, int L, int D, int N) {
    if (B <= 0 || L <= 0 || D <= 0 || N <= 0) { set_error("%s: B, L, D, N must be positive", who); return MMI_ERR_ARG; }
    if (B > 65535) { set_error("%s: B=%d exceeds 65535", who, B); return MMI_ERR_ARG; }
    if (int64_t(D) * N > (int64_t(1) << 30)) { set_error("%s: D*N too large", who); return MMI_ERR_ARG; }
    // one CTA per (batch, 32-step segment, 128 columns): the grid is one-dimensional
    const int64_t items = int64_t(B) * ((L + 31) / 32) * ((int64_t(D) * N + 127) / 128);
    if (items > 0x7fffffffLL) { set_error("%s: B * ceil(L/32) * ceil(D*N/128) = %lld exceeds the grid limit", who, (long long)items); return MMI_ERR_ARG; }
    if (reinterpret_cast<uintptr_t>(ws) & 15) { set_error("%s: the workspace must be 16-byte aligned (128-bit records)", who); return MMI_ERR_ARG; }
    return require_device();
}

int mmi_pscan_fwd(const float *A, const float *X, float *H, void *ws, int B, int L, int D, int N, void *stream) {
    if (!A || !X || !H || !ws) { set_error("mmi_pscan_fwd: null pointer"); return MMI_ERR_ARG; }
    if (int e = check_pscan("mmi_pscan_fwd", ws, B, L, D, N)) return e;
    return pscan_fwd_launch(A, X, H, static_cast<float *>(ws), B, L, D * N, static_cast<cudaStream_t>(stream));
}

int mmi_pscan_bwd(const float *A, const float *H, const float *gH, float *gA, float *gX, void *ws, int B, int L, int D,
                  int N, void *stream) {
    if (!A || !H || !gH || !gA || !gX || !ws) { set_error("mmi_pscan_bwd: null pointer"); return MMI_ERR_ARG; }
    if (int e = check_pscan("mmi_pscan_bwd", ws, B, L, D, N)) return e;
    return pscan_bwd_launch(A, H, gH, gA, gX, static_cast<float *>(ws), B, L, D * N, static_cast<cudaStream_t>(stream));
}

int mmi_pscan_fwd_f64(const double *A, const double *X, double *H, void *ws, int B, int L, int D, int N, void *stream) {
    if (!A || !X || !H || !ws) { set_error("mmi_pscan_fwd_f64: null pointer"); return MMI_ERR_ARG; }
    if (int e = check_pscan("mmi_pscan_fwd_f64", ws, B, L, D, N)) return e;
    return pscan_fwd_launch_f64(A, X, H, static_cast<double *>(ws), B, L, D * N, static_cast<cudaStream_t>(stream));
}

int mmi_pscan_bwd_f64(const double *A, const double *H, const double *gH, double *gA, double *gX, void *ws, int B, int L, int D,
                      int N, void *stream) {
    if (!A || !H || !gH || !gA || !gX || !ws) { set_error("mmi_pscan_bwd_f64: null pointer"); return MMI_ERR_ARG; }
    if (int e = check_pscan("mmi_pscan_bwd_f64", ws, B, L, D, N)) return e;
    return pscan_bwd_launch_f64(A, H, gH, gA, gX, static_cast<double *>(ws), B, L, D * N, static_cast<cudaStream_t>(stream));
}

int mmi_ffm_extract(const void *img, void *low, void *high, float *high_mul, int BC, int H, int W, int dtype,
                    void *stream) {
    if (!img || !low || !high) { set_error("mmi_ffm_extract: null pointer"); return MMI_ERR_ARG; }
    if (BC <= 0) { set_error("mmi_ffm_extract: BC must be positive"); return MMI_ERR_ARG; }
    if (!elem_size(dtype)) { set_error("mmi_ffm_extract: unknown dtype %d", dtype); return MMI_ERR_ARG; }
    if (int e = require_device()) return e;
    return ffm_extract_launch(img, low, high, high_mul, BC, H, W, dtype, static_cast<cudaStream_t>(stream));
}

void mmi_ffm_kept_range(int H, int W, int *r0, int *r1, int *c0, int *c1) { ffm_kept_range(H, W, r0, r1, c0, c1); }

int mmi_separation_loss(const float *M, float *loss, int l, int K, void *stream) {
    if (!M || !loss) { set_error("mmi_separation_loss: null pointer"); return MMI_ERR_ARG; }
    if (l < 2 || K < 1) { set_error("mmi_separation_loss: need l >= 2 rows and K >= 1 columns (l=%d K=%d)", l, K); return MMI_ERR_ARG; }
    if (int e = require_device()) return e;
    return separation_loss_launch(M, loss, l, K, static_cast<cudaStream_t>(stream));
}

#define MMI_RESAMPLE_ENTRY(NAME, SRC, DST, LAUNCH, MODE)                                                             \
    int NAME(const void *SRC, void *DST, int BC, int H, int W, int hs, int ws, int dtype, void *stream) {              \
        if (!SRC || !DST) { set_error(#NAME ": null pointer"); return MMI_ERR_ARG; }                                    \
        if (!elem_size(dtype)) { set_error(#NAME ": unknown dtype %d", dtype); return MMI_ERR_ARG; }                    \
        if (int e = require_device()) return e;                                                                         \
        return LAUNCH(SRC, DST, BC, H, W, hs, ws, MODE, dtype, static_cast<cudaStream_t>(stream), #NAME);               \
    }
MMI_RESAMPLE_ENTRY(mmi_avgpool_fwd, big, small, resample_reduce_launch, 0)
MMI_RESAMPLE_ENTRY(mmi_avgpool_bwd, dsmall, dbig, resample_expand_launch, 0)
MMI_RESAMPLE_ENTRY(mmi_upsample_bilinear_fwd, small, big, resample_expand_launch, 1)
MMI_RESAMPLE_ENTRY(mmi_upsample_bilinear_bwd, dbig, dsmall, resample_reduce_launch, 1)
#undef MMI_RESAMPLE_ENTRY

int64_t mmi_ffm_pattern_ws_bytes(int B, int C, int P) { return B > 0 && C > 0 && P > 0 ? ffm_pattern_ws_bytes(B, C, P) : 0; }

int mmi_ffm_pattern_fwd(const void *fea_vis, const void *fea_ir, const float *W1, const float *W2, void *tok, float *rows,
                        float *loss, void *ws, int B, int C, int H, int W, int dtype, void *stream) {
    if (!fea_vis || !fea_ir || !W1 || !W2 || !tok || !rows) { set_error("mmi_ffm_pattern_fwd: null pointer"); return MMI_ERR_ARG; }
    if (H < 1 || W < 1) { set_error("mmi_ffm_pattern_fwd: pooled map %dx%d", H, W); return MMI_ERR_ARG; }
    if (!elem_size(dtype)) { set_error("mmi_ffm_pattern_fwd: unknown dtype %d", dtype); return MMI_ERR_ARG; }
    if (int e = require_device()) return e;
    return ffm_pattern_fwd_launch(fea_vis, fea_ir, W1, W2, tok, rows, loss, static_cast<float *>(ws), B, C, H, W, dtype, static_cast<cudaStream_t>(stream));
}

int mmi_ffm_pattern_bwd(const void *fea_vis, const void *fea_ir, const void *dtok, const float *rows, const float *W1,
                        const float *W2, void *dfea_vis, void *dfea_ir, float *dW1, float *dW2, void *ws, int B, int C,
                        int P, int dtype, void *stream) {
    if (!fea_vis || !fea_ir || !dtok || !rows || !W1 || !W2 || !dfea_vis || !dfea_ir || !dW1 || !dW2 || !ws) { set_error("mmi_ffm_pattern_bwd: null pointer"); return MMI_ERR_ARG; }
    if (!elem_size(dtype)) { set_error("mmi_ffm_pattern_bwd: unknown dtype %d", dtype); return MMI_ERR_ARG; }
    if (int e = require_device()) return e;
    return ffm_pattern_bwd_launch(fea_vis, fea_ir, dtok, rows, W1, W2, dfea_vis, dfea_ir, dW1, dW2, static_cast<float *>(ws), B, C, P, dtype, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

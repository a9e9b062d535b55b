// pscan.cu -- materialised linear recurrence H[t] = A[t] H[t-1] + X[t] on (B, L, D, N) tensors: the parity API for
// `pscan = PScan.apply` (models/pscan.py:226).  The reference runs a Blelloch up/down sweep over a power-of-two padded
// copy (models/pscan.py:37-92, :152-186: ~76 strided launches at L=6400); here L is cut into segments, pass 1 reduces
// every segment to its (product, local end state) pair and pass 2 chains the pairs of the preceding segments
// (a few dozen FMAs per thread) and rescans its own segment, so both passes are fully parallel over B x segments x D*N
// and every load is coalesced along the contiguous D*N axis.  No padding, inputs untouched.
// Backward (models/pscan.py:189-224) is the same machinery run right-to-left on the left-shifted A.
#include "../../include/mmidet_b200.h"
#include "common.cuh"
#include "selscan.h"

namespace mmi {

constexpr int kPscanSeg = 64;  // minimum segment length

static int pscan_nseg(int B, int L, int DN) {
    const long rows = long(B) * DN;
    long want = (long(sm_count()) * 2048 * 2 + rows - 1) / rows;  // ~2 waves of threads
    long maxseg = (L + kPscanSeg - 1) / kPscanSeg;
    if (want < 1) want = 1;
    if (want > maxseg) want = maxseg;
    return int(want);
}

// REV = false: forward scan of (A, X).  REV = true: reverse scan with coefficient A[t+1] (0 at t = L-1).
template <bool REV, typename S>
__global__ void pscan_summary_kernel(const S *__restrict__ A, const S *__restrict__ X, S *__restrict__ ws,
                                     int L, int DN, int nseg, int seglen) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= DN) return;
    const int seg = blockIdx.y, b = blockIdx.z;
    const int t0 = seg * seglen, t1 = min(L, t0 + seglen);
    const S *a = A + (int64_t(b) * L) * DN + col, *x = X + (int64_t(b) * L) * DN + col;
    S P = S(1), Sum = S(0);
    if (!REV) {
#pragma unroll 8
        for (int t = t0; t < t1; ++t) {
            const S av = __ldg(a + int64_t(t) * DN), xv = __ldg(x + int64_t(t) * DN);
            Sum = fma(av, Sum, xv);
            P *= av;
        }
    } else {
#pragma unroll 8
        for (int t = t1 - 1; t >= t0; --t) {
            const S av = (t + 1 < L) ? __ldg(a + int64_t(t + 1) * DN) : S(0), xv = __ldg(x + int64_t(t) * DN);
            Sum = fma(av, Sum, xv);
            P *= av;
        }
    }
    S *w = ws + ((int64_t(b) * nseg + seg) * 2) * DN + col;
    w[0] = P;
    w[DN] = Sum;
}

template <bool REV, typename S>
__global__ void pscan_apply_kernel(const S *__restrict__ A, const S *__restrict__ X, const S *__restrict__ Hin,
                                   const S *__restrict__ ws, S *__restrict__ out, S *__restrict__ gA, int L,
                                   int DN, int nseg, int seglen) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= DN) return;
    const int seg = blockIdx.y, b = blockIdx.z;
    const int t0 = seg * seglen, t1 = min(L, t0 + seglen);
    const int64_t base = (int64_t(b) * L) * DN + col;
    const S *a = A + base, *x = X + base;
    S h = S(0);  // state entering this segment = chain of the summaries before (after, if REV) it
    if (!REV) {
        for (int s = 0; s < seg; ++s) {
            const S *w = ws + ((int64_t(b) * nseg + s) * 2) * DN + col;
            h = fma(w[0], h, w[DN]);
        }
#pragma unroll 8
        for (int t = t0; t < t1; ++t) {
            h = fma(__ldg(a + int64_t(t) * DN), h, __ldg(x + int64_t(t) * DN));
            __stcs(out + base + int64_t(t) * DN, h);
        }
    } else {
        for (int s = nseg - 1; s > seg; --s) {
            const S *w = ws + ((int64_t(b) * nseg + s) * 2) * DN + col;
            h = fma(w[0], h, w[DN]);
        }
#pragma unroll 8
        for (int t = t1 - 1; t >= t0; --t) {
            const S av = (t + 1 < L) ? __ldg(a + int64_t(t + 1) * DN) : S(0);
            h = fma(av, h, __ldg(x + int64_t(t) * DN));                          // G[t]   (pscan.py:216-219)
            __stcs(out + base + int64_t(t) * DN, h);                              // gradX  (pscan.py:224)
            const S hp = (t > 0) ? __ldg(Hin + base + int64_t(t - 1) * DN) : S(0);
            __stcs(gA + base + int64_t(t) * DN, hp * h);                          // gradA  (pscan.py:221-222)
        }
    }
}

int64_t pscan_ws_bytes(int B, int L, int D, int N) {
    const int64_t DN = int64_t(D) * N;
    const int64_t maxseg = (L + kPscanSeg - 1) / kPscanSeg;
    return int64_t(B) * maxseg * 2 * DN * 8;  // sized for the fp64 entry points
}

template <bool REV, typename S>
static int pscan_run(const S *A, const S *X, const S *Hin, S *out, S *gA, S *ws, int B, int L,
                     int DN, cudaStream_t st) {
    const int nseg = pscan_nseg(B, L, DN);
    const int seglen = (L + nseg - 1) / nseg;
    const int nseg_eff = (L + seglen - 1) / seglen;
    dim3 block(128), grid((DN + 127) / 128, nseg_eff, B);
    if (nseg_eff > 1) {
        pscan_summary_kernel<REV, S><<<grid, block, 0, st>>>(A, X, ws, L, DN, nseg_eff, seglen);
        if (int e = check_cuda(cudaGetLastError(), "pscan summary launch")) return e;
    }
    pscan_apply_kernel<REV, S><<<grid, block, 0, st>>>(A, X, Hin, ws, out, gA, L, DN, nseg_eff, seglen);
    return check_cuda(cudaGetLastError(), "pscan apply launch");
}

int pscan_fwd_launch(const float *A, const float *X, float *H, float *ws, int B, int L, int DN, cudaStream_t st) {
    return pscan_run<false, float>(A, X, nullptr, H, nullptr, ws, B, L, DN, st);
}
int pscan_bwd_launch(const float *A, const float *H, const float *gH, float *gA, float *gX, float *ws, int B, int L, int DN,
                     cudaStream_t st) {
    return pscan_run<true, float>(A, gH, H, gX, gA, ws, B, L, DN, st);
}
// fp64 variants: models/pscan.py is dtype-generic ("fp64 works"), so the drop-in is too
int pscan_fwd_launch_f64(const double *A, const double *X, double *H, double *ws, int B, int L, int DN, cudaStream_t st) {
    return pscan_run<false, double>(A, X, nullptr, H, nullptr, ws, B, L, DN, st);
}
int pscan_bwd_launch_f64(const double *A, const double *H, const double *gH, double *gA, double *gX, double *ws, int B, int L,
                         int DN, cudaStream_t st) {
    return pscan_run<true, double>(A, gH, H, gX, gA, ws, B, L, DN, st);
}

}  // namespace mmi

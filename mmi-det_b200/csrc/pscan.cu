// pscan.cu -- materialised linear recurrence H[t] = A[t] H[t-1] + X[t] on (B, L, D, N) tensors: the parity API for
// `pscan = PScan.apply` (models/pscan.py:226).  The reference runs a Blelloch up/down sweep over a power-of-two padded
// copy (models/pscan.py:37-92, :152-186: ~76 strided launches at L=6400).  Here the scan is ONE pass over the tensors
// (read A, X once, write H once -- the algorithmic minimum; round 1 read A and X twice):
//   item     (batch element, 128 adjacent columns of the contiguous D*N axis, segment of 32 steps); one CTA per item, a thread
//            owns one column and holds the segment's (A, X) in registers -- every load of the segment is issued before the first
//            use, coalesced along D*N
//   chain    decoupled look-back along L, per column: the thread publishes a record {product of A, local end state, inclusive
//            state, status} for its segment (fp32: one 128-bit store, read back with one 128-bit load, so no fence and no CTA
//            barrier is involved), walks back two records at a time to the nearest inclusive state, chains the aggregates in
//            between onto it oldest first (so the value is the same whatever the timing was), publishes its own inclusive state
//            and only then rescans its registers from the right entry state and writes H.  Items are handed out by a ticket in dependency order, so a thread only ever waits on
//            CTAs that already run.
// No padding, inputs untouched.  Backward (models/pscan.py:189-224) is the same machinery run right-to-left on the
// left-shifted A, with gradA = H[t-1] G[t] written in the same pass.
#include "../../include/mmidet_b200.h"
#include "common.cuh"
#include "selscan.h"

namespace mmi {

constexpr int kPsCols = 128;  // columns (threads) per CTA
constexpr int kPsT = 32;      // steps per segment
constexpr int kPsDepth = 8;   // look-back depth limit (parked aggregates live in registers)
constexpr int kPsWin = 2;     // predecessors examined per look-back round (measured: 2..4 equal, 8+ costs registers)

// ---- per-(batch, segment, column) records ------------------------------------------------------------------------------
// status 0: nothing yet (the launcher clears it), 1: (P, Sum) valid, 2: inclusive state valid as well
template <typename S> struct PsRec;
template <> struct PsRec<float> {
    static constexpr size_t BYTES = 16, CLEAR = 16;  // {P, Sum, incl, status}: one 128-bit transaction either way
    float4 *r;
    __device__ void put(int64_t i, float P, float Sum, float incl, int status) const {
        asm volatile("st.volatile.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(r + i), "f"(P), "f"(Sum), "f"(incl),
                     "f"(__int_as_float(status))
                     : "memory");
    }
    __device__ void get(int64_t i, float &P, float &Sum, float &incl, int &status) const {
        float s;
        asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(P), "=f"(Sum), "=f"(incl), "=f"(s) : "l"(r + i) : "memory");
        status = __float_as_int(s);
    }
};
template <> struct PsRec<double> {
    static constexpr size_t BYTES = 32, CLEAR = 8;  // status words first (cleared), then {P, Sum, incl, pad} per record
    unsigned long long *st;
    double *d;
    __device__ void put(int64_t i, double P, double Sum, double incl, int status) const {
        __stcg(d + 4 * i, P), __stcg(d + 4 * i + 1, Sum), __stcg(d + 4 * i + 2, incl);
        __threadfence();
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(st + i), "l"((unsigned long long)status) : "memory");
    }
    __device__ void get(int64_t i, double &P, double &Sum, double &incl, int &status) const {
        unsigned long long s;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(s) : "l"(st + i) : "memory");
        status = int(s);
        P = __ldcg(d + 4 * i), Sum = __ldcg(d + 4 * i + 1), incl = __ldcg(d + 4 * i + 2);
    }
};
template <typename S> static PsRec<S> ps_carve(void *ws, int64_t nrec);
template <> PsRec<float> ps_carve<float>(void *ws, int64_t) { return {reinterpret_cast<float4 *>(static_cast<char *>(ws) + 256)}; }
template <> PsRec<double> ps_carve<double>(void *ws, int64_t nrec) {
    unsigned long long *st = reinterpret_cast<unsigned long long *>(static_cast<char *>(ws) + 256);
    return {st, reinterpret_cast<double *>(st + nrec)};
}
// workspace: [ticket, 256 B] [records]; sized for the fp64 entry points
int64_t pscan_ws_bytes(int B, int L, int D, int N) {
    const int64_t nseg = (L + kPsT - 1) / kPsT;
    return 256 + int64_t(B) * nseg * D * N * (PsRec<double>::BYTES + PsRec<double>::CLEAR);
}

// REV = false: forward scan of (A, X) -> out.  REV = true: reverse scan with coefficient A[t+1] (0 at t = L-1) of X = gradH
// -> out = gradX = G (pscan.py:216-219, :224) and gA[t] = H[t-1] G[t] (pscan.py:221-222).
template <bool REV, typename S>
__global__ void __launch_bounds__(kPsCols, sizeof(S) == 4 ? 4 : 2)
    pscan_chain_kernel(const S *__restrict__ A, const S *__restrict__ X, const S *__restrict__ Hin, S *__restrict__ out,
                       S *__restrict__ gA, unsigned *ticket, PsRec<S> rec, int B, int L, int DN, int nseg, int ncb) {
    constexpr int T = kPsT, W = kPsWin, D = kPsDepth;
    __shared__ unsigned s_id;
    if (threadIdx.x == 0) s_id = atomicAdd(ticket, 1u);
    __syncthreads();
    const int id = int(s_id), nchain = B * ncb;
    const int k = id / nchain, r = id % nchain, b = r / ncb, cb = r % ncb;  // k = position of the segment in dependency order
    const int seg = REV ? nseg - 1 - k : k, col = cb * kPsCols + threadIdx.x, t0 = seg * T;
    if (col >= DN) return;
    const int64_t base = int64_t(b) * L * DN + col;
    const S *pa = A + base + int64_t(REV ? t0 + 1 : t0) * DN, *px = X + base + int64_t(t0) * DN;

    S a[T], x[T];
    const bool whole = t0 + T + (REV ? 1 : 0) <= L;  // CTA-uniform: every row of the segment (and of the shifted A) exists
    if (whole) {                                      // running pointers, no predicates: two 64-bit adds per load
#pragma unroll
        for (int i = 0; i < T; ++i) {
            a[i] = __ldcs(pa), x[i] = __ldcs(px);
            pa += DN, px += DN;
        }
    } else {
#pragma unroll
        for (int i = 0; i < T; ++i) {  // rows past L are identity steps
            const int t = t0 + i;
            a[i] = ((REV ? t + 1 : t) < L) ? __ldcs(pa + int64_t(i) * DN) : ((REV && t < L) ? S(0) : S(1));
            x[i] = t < L ? __ldcs(px + int64_t(i) * DN) : S(0);
        }
    }
    S P = S(1), Sum = S(0);
#pragma unroll
    for (int ii = 0; ii < T; ++ii) {
        const int i = REV ? T - 1 - ii : ii;
        Sum = fma(a[i], Sum, x[i]);
        P *= a[i];
    }
    // record index of the segment at position j of the dependency order
    auto rix = [&](int j) { return (int64_t(b) * nseg + (REV ? nseg - 1 - j : j)) * DN + col; };
    if (k == 0) {
        if (nseg > 1) rec.put(rix(0), P, Sum, Sum, 2);  // enters with state 0: its local end state is the inclusive one
    } else if (k + 1 < nseg) {
        rec.put(rix(k), P, Sum, S(0), 1);
    }
    S hp[T];  // reverse only: H[t-1] of the segment's rows, in flight during the look-back
    if constexpr (REV) {
        const S *ph = Hin + base + int64_t(t0 - 1) * DN;
        if (whole && t0 > 0) {
#pragma unroll
            for (int i = 0; i < T; ++i) {
                hp[i] = __ldcs(ph);
                ph += DN;
            }
        } else {
#pragma unroll
            for (int i = 0; i < T; ++i) {
                const int t = t0 + i;
                hp[i] = (t < L && t > 0) ? __ldcs(ph + int64_t(i) * DN) : S(0);
            }
        }
    }

    // ---- look-back: state entering the segment ----------------------------------------------------------------------
    // Walk back over at most kPsDepth segments, two records per round trip, parking the aggregates met on the way, until a
    // record carries an inclusive state (the one at depth kPsDepth is waited for); then apply the parked aggregates oldest
    // first.  Whatever depth the inclusive state is found at, the value is the segment-by-segment chain
    // h <- fma(P_j, h, Sum_j) from state 0, so the result does not depend on timing: bit-reproducible run to run.
    S h = S(0);
    if (k > 0) {
        S bp[D - 1], bs[D - 1];
        int n = -1;  // depth at which the inclusive state was found
        const long long tw = clock64();
#pragma unroll
        for (int r0 = 0; r0 < D; r0 += W) {
            if (n < 0) {
                bool retry;
                do {
                    S vP[W], vS[W], vI[W];
                    int f[W];
#pragma unroll
                    for (int u = 0; u < W; ++u) {
                        const int j = k - 1 - (r0 + u);
                        if (j >= 0) rec.get(rix(j), vP[u], vS[u], vI[u], f[u]);
                        else vP[u] = S(1), vS[u] = S(0), vI[u] = S(0), f[u] = 2;  // before the first segment: state 0
                    }
                    retry = false;
#pragma unroll
                    for (int u = 0; u < W; ++u) {
                        if (n < 0 && !retry) {
                            if (f[u] == 2) {
                                h = vI[u];
                                n = r0 + u;
                            } else if (f[u] == 0 || r0 + u == D - 1) {
                                retry = true;  // not published yet, or the walk is at its depth limit: wait for this record
                            } else {
                                bp[r0 + u < D - 1 ? r0 + u : 0] = vP[u];
                                bs[r0 + u < D - 1 ? r0 + u : 0] = vS[u];
                            }
                        }
                    }
                    if (retry) {
                        __nanosleep(64);
                        if (clock64() - tw > 20000000000LL) __trap();  // ~10 s: a lost predecessor traps instead of hanging
                    }
                } while (retry);
            }
        }
#pragma unroll
        for (int u = D - 2; u >= 0; --u)
            if (u < n) h = fma(bp[u], h, bs[u]);
        if (k + 1 < nseg) rec.put(rix(k), P, Sum, fma(P, h, Sum), 2);  // before the output traffic: others wait on it
    }

    // ---- rescan from the entry state, write ---------------------------------------------------------------------------
    S *po = out + base + int64_t(t0) * DN, *pg = REV ? gA + base + int64_t(t0) * DN : nullptr;
    if (whole) {
        if (REV) po += int64_t(T - 1) * DN, pg += int64_t(T - 1) * DN;
#pragma unroll
        for (int ii = 0; ii < T; ++ii) {
            const int i = REV ? T - 1 - ii : ii;
            h = fma(a[i], h, x[i]);
            __stcs(po, h);
            if constexpr (REV) {
                __stcs(pg, hp[i] * h);
                po -= DN, pg -= DN;
            } else {
                po += DN;
            }
        }
    } else {
#pragma unroll
        for (int ii = 0; ii < T; ++ii) {
            const int i = REV ? T - 1 - ii : ii, t = t0 + i;
            h = fma(a[i], h, x[i]);
            if (t < L) {
                __stcs(po + int64_t(i) * DN, h);
                if constexpr (REV) __stcs(pg + int64_t(i) * DN, hp[i] * h);
            }
        }
    }
}

template <bool REV, typename S>
static int pscan_run(const S *A, const S *X, const S *Hin, S *out, S *gA, void *ws, int B, int L, int DN, cudaStream_t st) {
    const int nseg = (L + kPsT - 1) / kPsT, ncb = (DN + kPsCols - 1) / kPsCols;
    const int64_t nrec = int64_t(B) * nseg * DN, items = int64_t(B) * nseg * ncb;
    if (int e = check_cuda(cudaMemsetAsync(ws, 0, 256 + size_t(nrec) * PsRec<S>::CLEAR, st), "pscan record reset")) return e;
    pscan_chain_kernel<REV, S><<<unsigned(items), kPsCols, 0, st>>>(A, X, Hin, out, gA, static_cast<unsigned *>(ws),
                                                                   ps_carve<S>(ws, nrec), B, L, DN, nseg, ncb);
    return check_cuda(cudaGetLastError(), "pscan launch");
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

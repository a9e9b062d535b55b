/*
 * oracle/selscan_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * CPU restatement of the reference's selective-scan hot path, used as the checker in tests/,
 * __graft_entry__.smoke() and as the `cpu_baseline` / `--impl reference` leg of bench.py.
 *
 * Pinned against the reference itself: tests/golden/*.npz are outputs of the UNMODIFIED reference
 * (models/pscan.py, models/mamba.py imported from /root/reference by tests/golden/make_golden.py);
 * tests/test_oracle.py checks every function below against them.
 *
 * Reference lines restated:
 *   - recurrence H[t] = A[t]*H[t-1] + X[t], H[-1] = 0          models/pscan.py:41-43 (what PScan.pscan computes)
 *   - pscan backward: G[t] = g[t] + A[t+1]*G[t+1]; gradA[t] = H[t-1]*G[t] (gradA[0]=0); gradX = G
 *                                                               models/pscan.py:206-224
 *   - selective scan: deltaA = exp(delta*A); BX = delta*B*x; hs = scan; y = hs@C + D*x
 *                                                               models/mamba.py:235-265 (selective_scan_seq)
 *   - gate: output = y * silu(z)                                models/mamba.py:184-186
 *
 * Arithmetic: `real` is float (default, the reference's dtype) or double (-DORACLE_F64 build, the
 * high-precision anchor). OpenMP over independent (batch, channel) rows.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_F64
typedef double real;
#define EXP exp
#define SFX(n) n##_f64
#else
typedef float real;
#define EXP expf
#define SFX(n) n##_f32
#endif

/* models/pscan.py:41-43 -- sequential statement of the scan PScan.pscan evaluates in parallel.
 * A, X, H: (B, L, D, N) contiguous. */
void SFX(oracle_pscan_fwd)(const real *A, const real *X, real *H, int B, int L, int D, int N) {
    const size_t DN = (size_t)D * N;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            for (int n = 0; n < N; ++n) {
                real h = 0;
                for (int t = 0; t < L; ++t) {
                    size_t i = ((size_t)b * L + t) * DN + (size_t)d * N + n;
                    h = A[i] * h + X[i];
                    H[i] = h;
                }
            }
}

/* models/pscan.py:206-224 -- reverse scan with A shifted one step left, then gradA = H[t-1]*G[t]. */
void SFX(oracle_pscan_bwd)(const real *A, const real *H, const real *gH, real *gA, real *gX, int B, int L, int D,
                           int N) {
    const size_t DN = (size_t)D * N;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            for (int n = 0; n < N; ++n) {
                real g = 0;
                for (int t = L - 1; t >= 0; --t) {
                    size_t i = ((size_t)b * L + t) * DN + (size_t)d * N + n;
                    real anext = (t + 1 < L) ? A[i + DN] : (real)0; /* pscan.py:216 */
                    g = gH[i] + anext * g;                          /* pscan.py:219 */
                    gX[i] = g;                                      /* pscan.py:224 */
                    gA[i] = (t > 0) ? H[i - DN] * g : (real)0;      /* pscan.py:221-222 */
                }
            }
}

static inline real silu_(real z) { return z / ((real)1 + EXP(-z)); }

/* models/mamba.py:235-265 (+ :184-186 when z != NULL).
 * x, delta, z, out: (B, L, ED); A: (ED, N); Bm, Cm: (B, L, N); D: (ED).
 * h0 (nullable): (B, ED, N) initial state; hT (nullable): final state. */
void SFX(oracle_selscan_fwd)(const real *x, const real *delta, const real *z, const real *A, const real *Bm,
                             const real *Cm, const real *D, const real *h0, real *out, real *hT, int B, int L, int ED,
                             int N) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < ED; ++d) {
            real h[64];
            for (int n = 0; n < N; ++n) h[n] = h0 ? h0[((size_t)b * ED + d) * N + n] : (real)0;
            for (int t = 0; t < L; ++t) {
                size_t i = ((size_t)b * L + t) * ED + d;
                const real *Bt = Bm + ((size_t)b * L + t) * N;
                const real *Ct = Cm + ((size_t)b * L + t) * N;
                real xv = x[i], dv = delta[i], y = 0;
                for (int n = 0; n < N; ++n) {
                    real a = EXP(dv * A[(size_t)d * N + n]); /* mamba.py:247 */
                    real u = (dv * Bt[n]) * xv;              /* mamba.py:248-250 */
                    h[n] = a * h[n] + u;                     /* mamba.py:255-257 */
                    y += h[n] * Ct[n];                       /* mamba.py:261 */
                }
                y += D[d] * xv; /* mamba.py:263 */
                out[i] = z ? y * silu_(z[i]) : y; /* mamba.py:184-186 */
            }
            if (hT)
                for (int n = 0; n < N; ++n) hT[((size_t)b * ED + d) * N + n] = h[n];
        }
}

/* Analytic gradient of the function above (autograd of mamba.py:222-231 through pscan.py:189-224).
 * Gradients w.r.t. x, delta, z (B,L,ED); Bm, Cm (B,L,N); A (ED,N); D (ED). dB/dC/dA/dD are ACCUMULATED
 * in double regardless of `real`, then cast. Not parallel over d for dB/dC (uses per-thread buffers). */
void SFX(oracle_selscan_bwd)(const real *x, const real *delta, const real *z, const real *A, const real *Bm,
                             const real *Cm, const real *D, const real *dout, real *dx, real *ddelta, real *dz,
                             real *dA, real *dB, real *dC, real *dD, int B, int L, int ED, int N) {
    double *accB = (double *)calloc((size_t)B * L * N, sizeof(double));
    double *accC = (double *)calloc((size_t)B * L * N, sizeof(double));
    double *accA = (double *)calloc((size_t)ED * N, sizeof(double));
    double *accD = (double *)calloc((size_t)ED, sizeof(double));
#pragma omp parallel
    {
        real *hs = (real *)malloc((size_t)(L + 1) * N * sizeof(real));
        double *locB = (double *)calloc((size_t)L * N, sizeof(double));
        double *locC = (double *)calloc((size_t)L * N, sizeof(double));
#pragma omp for collapse(2) schedule(static)
        for (int b = 0; b < B; ++b)
            for (int d = 0; d < ED; ++d) {
                /* forward recompute, keeping every state */
                for (int n = 0; n < N; ++n) hs[n] = 0;
                for (int t = 0; t < L; ++t) {
                    size_t i = ((size_t)b * L + t) * ED + d;
                    const real *Bt = Bm + ((size_t)b * L + t) * N;
                    for (int n = 0; n < N; ++n) {
                        real a = EXP(delta[i] * A[(size_t)d * N + n]);
                        hs[(size_t)(t + 1) * N + n] = a * hs[(size_t)t * N + n] + (delta[i] * Bt[n]) * x[i];
                    }
                }
                real g[64];
                real anext[64];
                for (int n = 0; n < N; ++n) { g[n] = 0; anext[n] = 0; }
                double aA[64];
                for (int n = 0; n < N; ++n) aA[n] = 0;
                double aD = 0;
                for (int t = L - 1; t >= 0; --t) {
                    size_t i = ((size_t)b * L + t) * ED + d;
                    const real *Bt = Bm + ((size_t)b * L + t) * N;
                    const real *Ct = Cm + ((size_t)b * L + t) * N;
                    real xv = x[i], dv = delta[i];
                    real y = 0;
                    for (int n = 0; n < N; ++n) y += hs[(size_t)(t + 1) * N + n] * Ct[n];
                    y += D[d] * xv;
                    real dy = dout[i];
                    if (z) {
                        real zv = z[i], s = (real)1 / ((real)1 + EXP(-zv));
                        dz[i] = dout[i] * y * s * ((real)1 + zv * ((real)1 - s));
                        dy = dout[i] * zv * s;
                    }
                    real ddv = 0, dxv = D[d] * dy;
                    for (int n = 0; n < N; ++n) {
                        real An = A[(size_t)d * N + n];
                        real a = EXP(dv * An);
                        g[n] = Ct[n] * dy + anext[n] * g[n]; /* pscan.py:216-219 */
                        real da = hs[(size_t)t * N + n] * g[n]; /* pscan.py:221-222 (hs[t] = H[t-1]) */
                        real du = g[n];                         /* pscan.py:224 */
                        ddv += da * a * An + du * Bt[n] * xv;
                        dxv += du * dv * Bt[n];
                        locB[(size_t)t * N + n] = (double)du * dv * xv;
                        locC[(size_t)t * N + n] = (double)dy * hs[(size_t)(t + 1) * N + n];
                        aA[n] += (double)da * a * dv;
                        anext[n] = a;
                    }
                    aD += (double)dy * xv;
                    dx[i] = dxv;
                    ddelta[i] = ddv;
                }
#pragma omp critical
                {
                    for (int t = 0; t < L; ++t)
                        for (int n = 0; n < N; ++n) {
                            accB[((size_t)b * L + t) * N + n] += locB[(size_t)t * N + n];
                            accC[((size_t)b * L + t) * N + n] += locC[(size_t)t * N + n];
                        }
                    for (int n = 0; n < N; ++n) accA[(size_t)d * N + n] += aA[n];
                    accD[d] += aD;
                }
            }
        free(hs);
        free(locB);
        free(locC);
    }
    for (size_t i = 0; i < (size_t)B * L * N; ++i) { dB[i] = (real)accB[i]; dC[i] = (real)accC[i]; }
    for (size_t i = 0; i < (size_t)ED * N; ++i) dA[i] = (real)accA[i];
    for (int i = 0; i < ED; ++i) dD[i] = (real)accD[i];
    free(accB); free(accC); free(accA); free(accD);
}

/* thread control for the timed CPU-baseline legs of bench.py (torchrun exports OMP_NUM_THREADS=1 to its workers) */
#ifndef ORACLE_F64 /* this file is compiled twice (f32, f64); define the helpers once */
#ifdef _OPENMP
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_max_threads(void) { return omp_get_max_threads(); }
#else
void oracle_set_threads(int n) { (void)n; }
int oracle_max_threads(void) { return 1; }
#endif
#endif

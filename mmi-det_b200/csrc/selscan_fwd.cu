// selscan_fwd.cu -- fused selective scan forward for sm_100a.
//
// Replaces MambaBlock.selective_scan (+ the SiLU gate) of the reference, models/mamba.py:212-233 and :184-186:
//     a[t,n] = exp(delta[t,d] * A[d,n]);  u[t,n] = delta[t,d] * B[t,n] * x[t,d]
//     h[t,n] = a[t,n] * h[t-1,n] + u[t,n]
//     y[t,d] = sum_n C[t,n] h[t,n] + D[d] x[t,d];   out = y * silu(z)
// The reference materialises four (B,L,ED,N) tensors and runs ~76 strided kernels of a Blelloch scan
// (models/pscan.py:37-92); here every input is read from HBM exactly once and nothing of size N is written except one
// state checkpoint per kChunk steps (for the backward pass).
//
// Mapping.  One lane owns one channel d and its N = 16 states in registers (channels-last layout: a warp reads 32
// adjacent channels of a timestep).  The only thing that is serial in t is h = a*h + u, and B*ED rows alone
// do not fill 148 SMs, so the time axis is parallelised INSIDE the CTA: a CTA owns CH = 32*WC channels of one batch
// element and walks L in super-tiles of ST = WT*TC steps (TC = 16); warp (wc, wt) owns chunk wt of the super-tile.  Tiles
// [ST x CH] of x / delta / z and [ST x N] of B / C are staged by the TMA engine (3-D tensor maps, one mbarrier per
// stage, a STAGES-deep ring), so HBM latency is covered by bytes in flight rather than by thread count.  Per super-tile:
//   sweep A  each warp reduces its chunk to (local end state E, sum of delta) by direct evaluation
//            E[n] = sum_t exp(A[n] * (delta summed after t)) * delta[t] x[t] B[t,n]      (no dependence on h)
//   fold     after one CTA barrier every warp chains the summaries of the chunks before it onto the carried state:
//            h <- exp(A * sum delta) * h + E  (the product of a chunk's decays is exp(A * sum delta)); the last warp
//            also publishes the carry for the next super-tile
//   sweep B  the real scan from the correct entry state: decay, recurrence, C.h readout, D skip, SiLU gate, in
//            branch-free software-pipelined blocks of kBlk steps; the output tile overwrites the z tile in shared
//            memory and leaves through one TMA tile store.
// exp() is MUFU ex2 on delta*A*log2(e); when a channel's A row is geometric, A[d,n] = (n+1) A[d,0] (the S4D-real init
// of models/mamba.py:158-159, which the reference training loop never updates), a[t,n] = r^(n+1) needs
// one ex2 per step instead of N (detected on the device, per CTA; MMI_FLAG_NO_GEOM forces the general path).
#include <algorithm>
#include <cstring>
#include <type_traits>

#include "../../include/mmidet_b200.h"
#include "common.cuh"
#include "selscan.h"

namespace mmi {

struct FwdMaps {
    CUtensorMap x, d, z, B, C, o;
};

constexpr int kBlk = 8;  // steps per branch-free block of sweep B
static_assert(kChunk % kBlk == 0, "sweep B writes a checkpoint at block starts");

template <typename T, int WC, int WT, int STAGES, int TCH> struct FwdLayout {
    static constexpr int N = kN, TC = TCH, ST = WT * TC, CH = 32 * WC, NW = WC * WT;
    static constexpr size_t TILE_BYTES = size_t(ST) * CH * sizeof(T);
    static constexpr size_t BCT_BYTES = size_t(ST) * N * sizeof(T);
    static constexpr size_t STAGE_BYTES = 3 * TILE_BYTES + 2 * BCT_BYTES;  // x | delta | z (-> out) | B | C
    static constexpr size_t SUM_OFF = STAGES * STAGE_BYTES;                // chunk end states  [WT][WC][4][32] float4
    static constexpr size_t SUMD_OFF = SUM_OFF + size_t(NW) * 4 * 32 * 16;  // chunk sum(delta)  [WT][WC][32] float
    static constexpr size_t CARRY_OFF = SUMD_OFF + size_t(NW) * 32 * 4;     // carried state [2][WC][4][32] float4
    static constexpr size_t BC32_OFF = CARRY_OFF + size_t(2) * WC * 4 * 32 * 16;
    static constexpr size_t BC32_BYTES = sizeof(T) == 2 ? size_t(NW) * 2 * TC * N * 4 : 0;  // per-warp widened B | C
    static constexpr size_t BAR_OFF = BC32_OFF + BC32_BYTES;
    static constexpr size_t SMEM = BAR_OFF + STAGES * sizeof(uint64_t);
};

// a[n] = exp(dsum * A[n]) for this lane's 16 states, as 8 packed pairs
template <bool GEOM> __device__ __forceinline__ void decay16(float dsum, float A2base, const float2 (&A2p)[8], float2 (&a2)[8]) {
    if constexpr (GEOM) {
        const float r = ex2(dsum * A2base), r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
        a2[0] = make_float2(r, r2);
        a2[1] = mul2(a2[0], splat2(r2));
        a2[2] = mul2(a2[0], splat2(r4));
        a2[3] = mul2(a2[1], splat2(r4));
#pragma unroll
        for (int k = 4; k < 8; ++k) a2[k] = mul2(a2[k - 4], splat2(r8));
    } else {
        const float2 d2 = splat2(dsum);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 e = mul2(d2, A2p[k]);
            a2[k] = make_float2(ex2(e.x), ex2(e.y));
        }
    }
}

__device__ __forceinline__ void load16(const float *p, float2 (&v)[8]) {  // 16 consecutive floats (LDS.128 x4)
    const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 w = q[k];
        v[2 * k] = make_float2(w.x, w.y);
        v[2 * k + 1] = make_float2(w.z, w.w);
    }
}

template <typename T, int WC, int WT, int STAGES, int TCH, bool SPLIT, bool GEOM, bool HAS_Z>
__device__ __forceinline__ void fwd_body(const FwdParams &p, const FwdMaps &tm, unsigned char *smem, const float2 (&A2p)[8],
                                         float A2base, float Dd, int c0, int b, int seg, int ctile, int wc, int wt, int lane, int c,
                                         bool active) {
    using Lay = FwdLayout<T, WC, WT, STAGES, TCH>;
    static_assert(TCH % kBlk == 0, "chunks are scanned in blocks of kBlk steps");
    constexpr int N = kN, TC = Lay::TC, ST = Lay::ST, CH = Lay::CH, NW = Lay::NW;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    float4 *sumE = reinterpret_cast<float4 *>(smem + Lay::SUM_OFF) + (wt * WC + wc) * 4 * 32 + lane;  // [k4 * 32]
    float *sumD = reinterpret_cast<float *>(smem + Lay::SUMD_OFF);
    float4 *carry = reinterpret_cast<float4 *>(smem + Lay::CARRY_OFF) + wc * 4 * 32 + lane;  // + buf * WC*4*32 + k4 * 32
    const int warp = wt * WC + wc;
    float *bc32 = reinterpret_cast<float *>(smem + Lay::BC32_OFF) + warp * 2 * TC * N;

    const int L = p.L, ED = p.ED;
    const int ntiles_all = (L + ST - 1) / ST, nchk = (L + kChunk - 1) / kChunk;
    const int tile_lo = seg * p.seg_tiles, ntiles = min(p.seg_tiles, ntiles_all - tile_lo);  // this CTA's segment of L
    const int cl = wc * 32 + lane, tb = wt * TC;

    // one elected thread: the tile loads of a super-tile arrive on full[s] (the summary pass needs x, delta and B only)
    auto issue = [&](int s, int ti, bool full_pass) {
        unsigned char *st = smem + size_t(s) * Lay::STAGE_BYTES;
        const uint32_t total = full_pass ? uint32_t(Lay::TILE_BYTES) * (HAS_Z ? 3u : 2u) + 2u * uint32_t(Lay::BCT_BYTES)
                                         : uint32_t(Lay::TILE_BYTES) * 2u + uint32_t(Lay::BCT_BYTES);
        mbar_arrive_expect_tx(&full[s], total);
        tma_load_3d(st, &tm.x, c0, ti * ST, b, &full[s]);
        tma_load_3d(st + Lay::TILE_BYTES, &tm.d, c0, ti * ST, b, &full[s]);
        if (HAS_Z && full_pass) tma_load_3d(st + 2 * Lay::TILE_BYTES, &tm.z, c0, ti * ST, b, &full[s]);
        tma_load_3d(st + 3 * Lay::TILE_BYTES, &tm.B, 0, ti * ST, b, &full[s]);
        if (full_pass) tma_load_3d(st + 3 * Lay::TILE_BYTES + Lay::BCT_BYTES, &tm.C, 0, ti * ST, b, &full[s]);
    };

    float2 hlast[8];  // state after the newest super-tile (kept by the last time-warp: segment summary, carry, hT)
#pragma unroll
    for (int k = 0; k < 8; ++k) hlast[k] = make_float2(0.f, 0.f);
    float sdseg = 0.f;  // sum of delta over the segment so far (summary pass, last time-warp)
    int g = 0;          // super-tiles processed so far over both passes: stage / mbarrier-parity / carry-buffer bookkeeping

    // When L is split over several CTAs (p.nseg > 1), every CTA but the one owning the last segment first reduces its
    // segment to (end state from zero, sum of delta) -- pass 0: sweep A + fold only, no C / z loads, no output -- and
    // publishes it; pass 1 then waits for the summaries of the earlier segments (decoupled look-back through global
    // memory), chains them onto h0 and scans the segment for real.
    for (int pass = (SPLIT && seg < p.nseg - 1) ? 0 : 1; pass < 2; ++pass) {
    const bool full_pass = !SPLIT || pass == 1;
    if (SPLIT && full_pass && seg > 0) {  // look-back: wait until every earlier segment of this (batch, channel tile) is published
        if (threadIdx.x == 0) {
            for (int sp = 0; sp < seg; ++sp) {
                const volatile unsigned *f = p.seg_flags + (int64_t(b) * p.nseg + sp) * p.ntile_c + ctile;
                const long long tw = clock64();
                while (*f < unsigned(WC)) {
                    if (clock64() - tw > 20000000000LL) __trap();  // ~10 s: a lost predecessor traps instead of hanging
                }
            }
            __threadfence();
        }
        __syncthreads();
    }
    if (wt == 0) {  // entry state of the pass: zeros (summary) | h0 (reference: zeros, models/mamba.py:252) chained through
        float2 hin[8];  //                                         the earlier segments' summaries
        const float4 *h0 = (full_pass && p.h0 && active) ? reinterpret_cast<const float4 *>(p.h0 + (int64_t(b) * ED + c) * N) : nullptr;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 w = h0 ? h0[k] : make_float4(0.f, 0.f, 0.f, 0.f);
            hin[2 * k] = make_float2(w.x, w.y);
            hin[2 * k + 1] = make_float2(w.z, w.w);
        }
        if (SPLIT && full_pass && active) {
            for (int sp = 0; sp < seg; ++sp) {
                const float *sw = p.seg_ws + ((int64_t(b) * p.nseg + sp) * ED + c) * (N + 1);
                float2 a2[8];
                decay16<GEOM>(__ldcg(sw + N), A2base, A2p, a2);
#pragma unroll
                for (int k = 0; k < 8; ++k) hin[k] = fma2(a2[k], hin[k], make_float2(__ldcg(sw + 2 * k), __ldcg(sw + 2 * k + 1)));
            }
        }
        float4 *cin = carry + (g & 1) * WC * 4 * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) cin[k * 32] = make_float4(hin[2 * k].x, hin[2 * k].y, hin[2 * k + 1].x, hin[2 * k + 1].y);
    }
    if (threadIdx.x == 0)
        for (int i = 0; i < STAGES && i < ntiles; ++i) issue((g + i) % STAGES, tile_lo + i, full_pass);

    for (int it = 0; it < ntiles; ++it, ++g) {
        const int s = g % STAGES, t0 = (tile_lo + it) * ST;
        unsigned char *st = smem + size_t(s) * Lay::STAGE_BYTES;
        const T *sx = reinterpret_cast<const T *>(st) + tb * CH + cl, *sz = sx + 2 * ST * CH;
        T *sd = const_cast<T *>(sx) + ST * CH;
        T *so = const_cast<T *>(sz);
        mbar_wait(&full[s], (g / STAGES) & 1);

        const float *fB, *fC;
        if constexpr (sizeof(T) == 2) {  // widen this chunk's B / C rows once per warp instead of once per use
            const T *gB = reinterpret_cast<const T *>(st + 3 * Lay::TILE_BYTES) + tb * N;
            const T *gC = reinterpret_cast<const T *>(st + 3 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N;
            for (int i = lane; i < TC * N; i += 32) {
                bc32[i] = to_f32<T>(gB[i]);
                if (full_pass) bc32[TC * N + i] = to_f32<T>(gC[i]);
            }
            __syncwarp();
            fB = bc32;
            fC = bc32 + TC * N;
        } else {
            fB = reinterpret_cast<const float *>(st + 3 * Lay::TILE_BYTES) + tb * N;
            fC = reinterpret_cast<const float *>(st + 3 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N;
        }

        if (p.flags & MMI_FLAG_DELTA_SOFTPLUS) {  // fused softplus(dt_proj(.)), models/mamba.py:203: activate this chunk's delta
#pragma unroll                              // once, in place, rounded to the I/O type exactly as the unfused path does
            for (int u = 0; u < TC; ++u)  // rows past L stay 0 (TMA zero fill): identity steps, so hT is h[L-1]
                sd[u * CH] = from_f32<T>(t0 + tb + u < L ? softplus_fast(to_f32<T>(sd[u * CH])) : 0.f);
        }
        // ---- sweep A: chunk summary by direct evaluation (walk t backwards, S = sum of delta after t) ----------
        float2 acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = make_float2(0.f, 0.f);
        float S = 0.f;
#pragma unroll
        for (int u = TC - 1; u >= 0; --u) {
            const float xv = to_f32<T>(sx[u * CH]), dv = to_f32<T>(sd[u * CH]);
            float2 Bv[8], pw[8];
            load16(fB + u * N, Bv);
            const float dx = dv * xv;
            if constexpr (GEOM) {
                const float r = ex2(S * A2base), r2 = r * r, r4 = r2 * r2, r8 = r4 * r4, p0 = dx * r;
                pw[0] = make_float2(p0, p0 * r);
                pw[1] = mul2(pw[0], splat2(r2));
                pw[2] = mul2(pw[0], splat2(r4));
                pw[3] = mul2(pw[1], splat2(r4));
#pragma unroll
                for (int k = 4; k < 8; ++k) pw[k] = mul2(pw[k - 4], splat2(r8));
            } else {
                const float2 S2 = splat2(S), dx2 = splat2(dx);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float2 e = mul2(S2, A2p[k]);
                    pw[k] = mul2(dx2, make_float2(ex2(e.x), ex2(e.y)));
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fma2(pw[k], Bv[k], acc[k]);
            S += dv;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sumE[k * 32] = make_float4(acc[2 * k].x, acc[2 * k].y, acc[2 * k + 1].x, acc[2 * k + 1].y);
        sumD[warp * 32 + lane] = S;
        __syncthreads();  // summaries of this super-tile (and the carry written during the previous one) are visible

        // the previous super-tile's output store has had a whole sweep to drain; its stage can be refilled
        if (threadIdx.x == 0 && it >= 1 && it - 1 + STAGES < ntiles) {
            bulk_wait_read<0>();
            issue((g - 1) % STAGES, tile_lo + it - 1 + STAGES, full_pass);
        }
        if (SPLIT && !full_pass) {  // summary pass: only the running end state of the segment is needed
            if (wt == WT - 1) {
                float2 hs[8];
                const float4 *cin = carry + (g & 1) * WC * 4 * 32;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 w = cin[k * 32];
                    hs[2 * k] = make_float2(w.x, w.y);
                    hs[2 * k + 1] = make_float2(w.z, w.w);
                }
#pragma unroll
                for (int v = 0; v < WT - 1; ++v) {
                    float2 a2[8];
                    const float sv = sumD[(v * WC + wc) * 32 + lane];
                    sdseg += sv;
                    decay16<GEOM>(sv, A2base, A2p, a2);
                    const float4 *e = reinterpret_cast<const float4 *>(smem + Lay::SUM_OFF) + (v * WC + wc) * 4 * 32 + lane;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 w = e[k * 32];
                        hs[2 * k] = fma2(a2[2 * k], hs[2 * k], make_float2(w.x, w.y));
                        hs[2 * k + 1] = fma2(a2[2 * k + 1], hs[2 * k + 1], make_float2(w.z, w.w));
                    }
                }
                float2 a2[8];
                decay16<GEOM>(S, A2base, A2p, a2);
                sdseg += S;
                float4 *cout = carry + ((g + 1) & 1) * WC * 4 * 32;
#pragma unroll
                for (int k = 0; k < 8; ++k) hlast[k] = fma2(a2[k], hs[k], acc[k]);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    cout[k * 32] = make_float4(hlast[2 * k].x, hlast[2 * k].y, hlast[2 * k + 1].x, hlast[2 * k + 1].y);
            }
            __syncthreads();  // summaries and stage s are free again
            continue;
        }

        // ---- fold: state entering this warp's chunk ---------------------------------------------------------
        float2 h2[8];
        {
            const float4 *cin = carry + (g & 1) * WC * 4 * 32;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 w = cin[k * 32];
                h2[2 * k] = make_float2(w.x, w.y);
                h2[2 * k + 1] = make_float2(w.z, w.w);
            }
        }
#pragma unroll
        for (int v = 0; v < WT - 1; ++v) {
            if (v < wt) {
                float2 a2[8];
                decay16<GEOM>(sumD[(v * WC + wc) * 32 + lane], A2base, A2p, a2);
                const float4 *e = reinterpret_cast<const float4 *>(smem + Lay::SUM_OFF) + (v * WC + wc) * 4 * 32 + lane;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 w = e[k * 32];
                    h2[2 * k] = fma2(a2[2 * k], h2[2 * k], make_float2(w.x, w.y));
                    h2[2 * k + 1] = fma2(a2[2 * k + 1], h2[2 * k + 1], make_float2(w.z, w.w));
                }
            }
        }
        if (wt == WT - 1) {  // carry for the next super-tile = this chunk's entry state pushed through its own summary
            float2 a2[8];
            decay16<GEOM>(S, A2base, A2p, a2);
            float4 *cout = carry + ((g + 1) & 1) * WC * 4 * 32;
#pragma unroll
            for (int k = 0; k < 8; ++k) hlast[k] = fma2(a2[k], h2[k], acc[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                cout[k * 32] = make_float4(hlast[2 * k].x, hlast[2 * k].y, hlast[2 * k + 1].x, hlast[2 * k + 1].y);
        }
        // ---- sweep B: the scan proper.  kBlk consecutive timesteps, branch-free and software-pipelined in three
        // phases so that independent work of neighbouring steps (LDS, MUFU, gate) overlaps the serial h chain:
        //   1  loads + per-step scalars (decay base r; delta*x; gate factor z*sigmoid(z))
        //   2  decay powers, h = a h + u, C.h readout                    (the only phase that is serial in t)
        //   3  D skip, gate, store into the output tile
#pragma unroll 1
        for (int ub = 0; ub < TC; ub += kBlk) {
            if (p.chk && active && (tb + ub) % kChunk == 0 && t0 + tb + ub < L) {  // checkpoint = state entering step t0 + tb + ub
                float4 *ck = reinterpret_cast<float4 *>(p.chk + ((int64_t(b) * nchk + (t0 + tb + ub) / kChunk) * ED + c) * N);
#pragma unroll
                for (int k = 0; k < 4; ++k) __stcs(ck + k, make_float4(h2[2 * k].x, h2[2 * k].y, h2[2 * k + 1].x, h2[2 * k + 1].y));
            }
            float xv[kBlk], dvv[kBlk], gz[kBlk], yv[kBlk];
#pragma unroll
            for (int u = 0; u < kBlk; ++u) {
                xv[u] = to_f32<T>(sx[(ub + u) * CH]);
                dvv[u] = to_f32<T>(sd[(ub + u) * CH]);
                if constexpr (HAS_Z) {
                    const float zv = to_f32<T>(sz[(ub + u) * CH]);
                    gz[u] = zv * sigmoidf_fast(zv);
                }
            }
#pragma unroll
            for (int u = 0; u < kBlk; ++u) {
                float2 Bv[8], Cv[8], a2[8];
                load16(fB + (ub + u) * N, Bv);
                load16(fC + (ub + u) * N, Cv);
                decay16<GEOM>(dvv[u], A2base, A2p, a2);
                const float2 dx2 = splat2(dvv[u] * xv[u]);
                float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    h2[k] = fma2(a2[k], h2[k], mul2(dx2, Bv[k]));
                    if (k & 1) yb = fma2(Cv[k], h2[k], yb);
                    else ya = fma2(Cv[k], h2[k], ya);
                }
                ya = add2(ya, yb);
                yv[u] = ya.x + ya.y;
            }
#pragma unroll
            for (int u = 0; u < kBlk; ++u) {
                float y = fmaf(Dd, xv[u], yv[u]);
                if constexpr (HAS_Z) y *= gz[u];
                so[(ub + u) * CH] = from_f32<T>(y);
            }
        }
        fence_proxy_async();  // make the generic-proxy writes of the output tile visible to the TMA engine
        __syncthreads();      // every warp is done with stage s
        if (threadIdx.x == 0) {
            tma_store_3d(&tm.o, c0, t0, b, st + 2 * Lay::TILE_BYTES);  // rows past L / columns past ED are clipped
            bulk_commit();
        }
    }
    if (SPLIT && !full_pass) {  // publish the segment summary, then raise the flag the later segments spin on
        if (wt == WT - 1) {
            if (active) {
                float *sw = p.seg_ws + ((int64_t(b) * p.nseg + seg) * ED + c) * (N + 1);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    __stcg(sw + 2 * k, hlast[k].x);
                    __stcg(sw + 2 * k + 1, hlast[k].y);
                }
                __stcg(sw + N, sdseg);
            }
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(p.seg_flags + (int64_t(b) * p.nseg + seg) * p.ntile_c + ctile, 1u);
        }
    }
    }  // pass
    if (threadIdx.x == 0) bulk_wait_read<0>();  // shared memory must outlive the last tile store's reads

    if (p.hT && active && wt == WT - 1 && seg == p.nseg - 1) {  // steps past L are identities (zero-filled delta), so this is h[L-1]
        float *hT = p.hT + (int64_t(b) * ED + c) * N;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            hT[2 * k] = hlast[k].x;
            hT[2 * k + 1] = hlast[k].y;
        }
    }
}

template <typename T, int WC, int WT, int STAGES, int TCH, bool SPLIT>
__global__ void __launch_bounds__(WC *WT * 32) selscan_fwd_kernel(const FwdParams p, const __grid_constant__ FwdMaps tm) {
    using Lay = FwdLayout<T, WC, WT, STAGES, TCH>;
    constexpr int N = kN, CH = Lay::CH;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wc = warp % WC, wt = warp / WC;
    int b = blockIdx.y, ctile = blockIdx.x, seg = 0;
    if constexpr (SPLIT) {  // L split over CTAs: take a ticket so that segments start in dependency order (earlier ones first)
        __shared__ unsigned ticket;
        if (tid == 0) ticket = atomicAdd(p.seg_ticket, 1u);
        __syncthreads();
        const int v = int(ticket);
        ctile = v % p.ntile_c;
        b = (v / p.ntile_c) % p.B;
        seg = v / (p.ntile_c * p.B);
    }
    const int c0 = ctile * CH;
    const int c = c0 + wc * 32 + lane;
    const bool active = c < p.ED;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }

    // A row of this lane's channel, pre-scaled by log2(e); geometric-row test (block-uniform decision)
    const int cc = active ? c : p.ED - 1;
    float2 A2p[8];
    const float A2base = p.A[int64_t(cc) * N] * kLog2e;
    bool ok = !(p.flags & MMI_FLAG_NO_GEOM);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        A2p[k] = make_float2(p.A[int64_t(cc) * N + 2 * k] * kLog2e, p.A[int64_t(cc) * N + 2 * k + 1] * kLog2e);
        const float w0 = float(2 * k + 1) * A2base, w1 = float(2 * k + 2) * A2base;
        ok = ok && (fabsf(A2p[k].x - w0) <= 2e-6f * fabsf(w0)) && (fabsf(A2p[k].y - w1) <= 2e-6f * fabsf(w1));
    }
    const float Dd = p.D[cc];
    const bool geom = __syncthreads_and(ok);  // also publishes the mbarrier inits
    const bool has_z = p.z != nullptr;

#define MMI_FWD_BODY(G, Z) \
    fwd_body<T, WC, WT, STAGES, TCH, SPLIT, G, Z>(p, tm, smem, A2p, A2base, Dd, c0, b, seg, ctile, wc, wt, lane, c, active)
    if (geom) {
        if (has_z) MMI_FWD_BODY(true, true);
        else MMI_FWD_BODY(true, false);
    } else {
        if (has_z) MMI_FWD_BODY(false, true);
        else MMI_FWD_BODY(false, false);
    }
#undef MMI_FWD_BODY
}

constexpr int kMaxSeg = 32;  // upper bound on the number of L segments (sizes the workspace)

// workspace of the L split: [ticket | flags (B, nseg, ntile_c)] [summaries (B, nseg, ED, N + 1) fp32]
static size_t seg_header_bytes(int B, int ntile_c) { return (size_t(16) + size_t(B) * kMaxSeg * ntile_c * 4 + 255) & ~size_t(255); }
// L is only ever split when the unsplit grid leaves most SMs idle (the heuristic below, also applied to a forced count), so
// the summaries' workspace is sized for kMaxSeg segments only in that case
static int fwd_seg_cap(int B, int ntile_c) { return int64_t(B) * ntile_c * 4 <= 2 * sm_count() ? kMaxSeg : 1; }
int64_t selscan_fwd1_ws_bytes(int B, int ED) {
    const int gx = (ED + 31) / 32;
    return int64_t(seg_header_bytes(B, gx)) + int64_t(B) * fwd_seg_cap(B, gx) * ED * (kN + 1) * 4;
}

template <typename T, int WC, int WT, int STAGES, int TCH = kFwdChunk>
static int launch_fwd_t(FwdParams p, int dtype, void *ws, cudaStream_t st) {
    using Lay = FwdLayout<T, WC, WT, STAGES, TCH>;
    auto kern = selscan_fwd_kernel<T, WC, WT, STAGES, TCH, false>;
    auto kern_split = selscan_fwd_kernel<T, WC, WT, STAGES, TCH, true>;
    FwdMaps tm;
    memset(&tm, 0, sizeof(tm));
    const uint64_t nb = p.B, L = p.L;
    if (int e = make_tmap_3d(&tm.x, p.x, dtype, nb, L, p.ED, p.x_ld * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (int e = make_tmap_3d(&tm.d, p.delta, dtype, nb, L, p.ED, p.d_ld * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (p.z)
        if (int e = make_tmap_3d(&tm.z, p.z, dtype, nb, L, p.ED, p.z_ld * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (int e = make_tmap_3d(&tm.B, p.Bm, dtype, nb, L, kN, kN * sizeof(T), Lay::ST, kN)) return e;
    if (int e = make_tmap_3d(&tm.C, p.Cm, dtype, nb, L, kN, kN * sizeof(T), Lay::ST, kN)) return e;
    if (int e = make_tmap_3d(&tm.o, p.out, dtype, nb, L, p.ED, p.o_ld * sizeof(T), Lay::ST, Lay::CH)) return e;
    static thread_local int attr_dev = -1;  // the opt-in is per device and sticky: set it once, not on every launch
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Lay::SMEM)),
                               "selscan_fwd smem attribute"))
            return e;
        if (int e = check_cuda(cudaFuncSetAttribute(kern_split, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Lay::SMEM)),
                               "selscan_fwd smem attribute"))
            return e;
        attr_dev = dev;
    }
    const int gx = (p.ED + Lay::CH - 1) / Lay::CH, ntiles = (p.L + Lay::ST - 1) / Lay::ST;
    // L split: when B * (ED / CH) CTAs cannot fill the GPU (small batches, inference), cut L into segments scanned by
    // different CTAs; each segment costs one extra summary pass over x / delta / B, so only split when SMs would idle
    int nseg = 1;
    if (ws) {
        const int forced = (p.flags & MMI_FLAG_NSEG_MASK) >> MMI_FLAG_NSEG_SHIFT;
        const int ctas = gx * p.B, slots = 2 * sm_count();
        nseg = forced ? forced : (ctas * 4 <= slots ? slots / ctas : 1);  // measured: splitting 128 CTAs in two loses
        nseg = std::max(1, std::min({nseg, fwd_seg_cap(p.B, (p.ED + 31) / 32), forced ? ntiles : ntiles / 2}));
    }
    p.seg_tiles = (ntiles + nseg - 1) / nseg;
    p.nseg = (ntiles + p.seg_tiles - 1) / p.seg_tiles;
    p.ntile_c = gx;
    if (p.nseg > 1) {
        const size_t hdr = seg_header_bytes(p.B, gx);
        p.seg_ticket = static_cast<unsigned *>(ws);
        p.seg_flags = p.seg_ticket + 4;
        p.seg_ws = reinterpret_cast<float *>(static_cast<char *>(ws) + hdr);
        if (int e = check_cuda(cudaMemsetAsync(ws, 0, hdr, st), "selscan_fwd segment flags memset")) return e;
        kern_split<<<dim3(unsigned(gx) * p.B * p.nseg), Lay::NW * 32, Lay::SMEM, st>>>(p, tm);
    } else {
        kern<<<dim3(gx, p.B), Lay::NW * 32, Lay::SMEM, st>>>(p, tm);
    }
    return check_cuda(cudaGetLastError(), "selscan_fwd launch");
}

int selscan_fwd1_launch(const FwdParams &p, int dtype, void *ws, cudaStream_t st) {
    const int cfg = (p.flags & MMI_FLAG_CFG_MASK) >> MMI_FLAG_CFG_SHIFT;
#define MMI_FWD_DISPATCH(T)                                            \
    switch (std::is_same<T, float>::value ? cfg : 0) {                 \
        case 1: return launch_fwd_t<float, 2, 4, 3>(p, dtype, ws, st);     \
        case 2: return launch_fwd_t<float, 1, 8, 2>(p, dtype, ws, st);     \
        case 3: return launch_fwd_t<float, 1, 2, 4>(p, dtype, ws, st);     \
        case 4: return launch_fwd_t<float, 1, 8, 3, 8>(p, dtype, ws, st);  \
        case 6: return launch_fwd_t<float, 2, 8, 2, 8>(p, dtype, ws, st);  \
        default: return launch_fwd_t<T, 1, 4, 3>(p, dtype, ws, st);        \
    }
    // (the alternative CTA shapes exist for tuning runs and shape tests: fp32 only)
    switch (dtype) {
        case MMI_F32: MMI_FWD_DISPATCH(float)
        case MMI_BF16: MMI_FWD_DISPATCH(__nv_bfloat16)
        case MMI_F16: MMI_FWD_DISPATCH(__half)
    }
#undef MMI_FWD_DISPATCH
    set_error("selscan_fwd: unknown dtype %d", dtype);
    return MMI_ERR_ARG;
}

}  // namespace mmi

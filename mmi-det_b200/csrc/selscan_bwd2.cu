// selscan_bwd2.cu -- fused selective scan backward, second generation (sm_100a).
//
// Same mathematics as selscan_bwd.cu (models/mamba.py:222-231 + :184-186 differentiated, PScan.backward of
// models/pscan.py:189-224: reverse scan on A shifted left, gradA = H[t-1] G[t], gradX = G; states recomputed from one
// checkpoint per 16 steps instead of the reference's five (B, Lp, ED, N) tensors).  What changed is the decomposition
// (selscan2.cuh): chains of 32 channels, 8 chunk-warps per CTA, lanes split the 16 states in two halves and pack the two
// channels of a pair into one register pair, a persistent grid takes (chain, L segment) items in dependency order.
//
// Per super-tile (128 steps) and warp (16 steps, walking L backwards across super-tiles):
//   prologue  dy = dout * silu(z) and the dz factor e = dout * sigma(z)(1 + z(1 - sigma(z))), parked over the dout / z tiles
//   phase 1   forward in time from the chunk's checkpoint: h is recomputed and every state is parked in TENSOR MEMORY
//             (tcgen05.st, thread-private 16-column slots: 256 KB hold the 64 B/(t,d) history of 128 steps x 32 channels);
//             y -> dz; dC partials (dy h summed over the warp's channels through a shared-memory transposition); and, in
//             the same sweep, the chunk's reverse-scan summary: with P_t = prod_{tau <= t} a_tau the a*g leaving the chunk
//             when nothing enters it is  G = sum_t P_t dy_t C_t,  and an incoming carry passes through as P_end * carry --
//             one running product and one FMA per state instead of a separate pass with its own exponentials
//   barrier   summaries of chunks 1..7 visible; fold: the warp chains the LATER chunks' (P_end, G) onto the carried a*g
//   phase 2   reverse in time: h[t-1] back from tensor memory (tcgen05.ld), g recurrence, ddelta / dx / dA / dD, dB partials.
//             Warp 0's a*g after its last step IS the carry of the next (earlier) super-tile -- chunk 0 needs no summary.
// dB / dC partials per (t, channel tile) and dA / dD partials per (batch, segment) go to the workspace; a finishing kernel
// reduces them in a fixed order (no atomics anywhere: results are bit-reproducible run to run).
#include <algorithm>
#include <cstring>

#include "../../include/mmidet_b200.h"
#include "selscan.h"
#include "selscan2.cuh"

namespace mmi {

using namespace v2;

constexpr int kRB2 = 4;  // steps per cross-channel reduction block

template <typename T> struct Bwd2Layout {
    static constexpr int N = kN, STAGES = 2;
    static constexpr size_t TILE_BYTES = size_t(kST) * kCH * sizeof(T);
    static constexpr size_t BCT_BYTES = size_t(kST) * N * sizeof(T);
    static constexpr size_t STAGE_BYTES = 4 * TILE_BYTES + 2 * BCT_BYTES;  // x | delta | dout | z | B | C
    static constexpr size_t DYE_OFF = STAGES * STAGE_BYTES;                // fp32 dy | e when T is 16 bit
    static constexpr size_t DYE_BYTES = sizeof(T) == 2 ? size_t(2) * kST * kCH * 4 : 0;
    static constexpr size_t BC32_OFF = DYE_OFF + DYE_BYTES;                // per-warp widened B | C rows when T is 16 bit
    static constexpr size_t BC32_BYTES = sizeof(T) == 2 ? size_t(kNW) * 2 * kTC * N * 4 : 0;
    static constexpr size_t SCR_OFF = BC32_OFF + BC32_BYTES;               // per warp [kRB2][2][32] float4
    static constexpr size_t SCR_WARP = size_t(kRB2) * 2 * 32 * 16;
    static constexpr size_t SUM_OFF = SCR_OFF + kNW * SCR_WARP;            // chunk summaries [kNW - 1][8][32] float4 (G | P_end)
    static constexpr size_t SUM_WARP = size_t(8) * 32 * 16;
    static constexpr size_t CARRY_OFF = SUM_OFF + (kNW - 1) * SUM_WARP;    // carried a*g [2][4][32] float4
    static constexpr size_t BAR_OFF = CARRY_OFF + size_t(2) * 4 * 32 * 16;
    static constexpr size_t SMEM = BAR_OFF + STAGES * sizeof(uint64_t) + 16;
    static_assert(SMEM <= 232448, "shared memory budget of one CTA per SM");
    static_assert(kNW * SCR_WARP >= size_t(kNW) * 8 * 32 * 8 + size_t(kNW) * 32 * 8, "dA / dD reduction aliases the scratch");
};

template <typename T, bool GEOM, bool HAS_Z>
__device__ __forceinline__ void bwd2_item(const Bwd2Params &pp, const Bwd2Maps &tm, unsigned char *smem, uint32_t tmem_base,
                                          const float2 (&A2p)[8], float2 A2b, float2 Dd, const float (&kw)[8], int c0, int b,
                                          int seg, int segi, int chain, int ctile, int wt, int lane, bool active, int &g) {
    using Lay = Bwd2Layout<T>;
    constexpr int N = kN, TC = kTC, ST = kST, CH = kCH;
    constexpr float kLn2 = 0.6931471805599453f;
    const BwdParams &p = pp.b;
    const SegSched &sc = pp.s;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    float4 *scr = reinterpret_cast<float4 *>(smem + Lay::SCR_OFF + wt * Lay::SCR_WARP);  // [kRB2][2][32]
    float4 *sums = reinterpret_cast<float4 *>(smem + Lay::SUM_OFF);                       // [(v - 1) * 8 + q][32]
    float4 *carry = reinterpret_cast<float4 *>(smem + Lay::CARRY_OFF) + lane;             // + (buf * 4 + q) * 32
    float *bc32 = reinterpret_cast<float *>(smem + Lay::BC32_OFF) + wt * 2 * TC * N;
    const int pr = lane & 15, hs = lane >> 4;
    const bool up = hs != 0;
    const uint32_t tslot = tmem_base + (uint32_t(wt & 3) * 32u << 16) + uint32_t(wt >> 2) * (TC * 16);

    const int L = p.L, ED = p.ED;
    const int ntiles_all = (L + ST - 1) / ST, nchk = (L + TC - 1) / TC;
    const int tile_lo = seg * sc.seg_tiles, ntiles = min(sc.seg_tiles, ntiles_all - tile_lo);
    const int tb = wt * TC, cl = 2 * pr, c = c0 + cl;
    const int64_t row_b = int64_t(b) * L;
    const float2 ddscale = GEOM ? make_float2(A2b.x * kLn2, A2b.y * kLn2) : make_float2(kLn2, kLn2);
    const bool softplus = (p.flags & MMI_FLAG_DELTA_SOFTPLUS) != 0;

    auto issue = [&](int s, int tj) {  // one elected thread: the six tiles of super-tile tj arrive on full[s]
        unsigned char *st = smem + size_t(s) * Lay::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], uint32_t(Lay::TILE_BYTES) * (HAS_Z ? 4u : 3u) + 2u * uint32_t(Lay::BCT_BYTES));
        tma_load_3d(st, &tm.x, c0, tj * ST, b, &full[s]);
        tma_load_3d(st + Lay::TILE_BYTES, &tm.d, c0, tj * ST, b, &full[s]);
        tma_load_3d(st + 2 * Lay::TILE_BYTES, &tm.g, c0, tj * ST, b, &full[s]);
        if (HAS_Z) tma_load_3d(st + 3 * Lay::TILE_BYTES, &tm.z, c0, tj * ST, b, &full[s]);
        tma_load_3d(st + 4 * Lay::TILE_BYTES, &tm.B, 0, tj * ST, b, &full[s]);
        tma_load_3d(st + 4 * Lay::TILE_BYTES + Lay::BCT_BYTES, &tm.C, 0, tj * ST, b, &full[s]);
    };
    if (threadIdx.x == 0) {
        bulk_wait_read<0>();  // output tiles of the previous item have left shared memory
        for (int i = 0; i < Lay::STAGES && i < ntiles; ++i) issue((g + i) % Lay::STAGES, tile_lo + ntiles - 1 - i);
    }

    // a*g entering the segment: zero for the last segment of L, else what the later segment of this chain left behind
    if (wt == 0) {
        float2 cin[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) cin[k] = make_float2(0.f, 0.f);
        if (segi > 0) {
            if (lane == 0) {
                const long long tw = clock64();
                while (ld_acquire(sc.done + chain) < unsigned(segi)) {
                    __nanosleep(64);
                    if (clock64() - tw > 20000000000LL) __trap();  // ~10 s: a lost predecessor traps instead of hanging
                }
            }
            __syncwarp();
            const float2 *gc = reinterpret_cast<const float2 *>(sc.carry) + int64_t(chain) * 8 * 32 + lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) cin[k] = __ldcg(gc + k * 32);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            carry[((g & 1) * 4 + q) * 32] = make_float4(cin[2 * q].x, cin[2 * q].y, cin[2 * q + 1].x, cin[2 * q + 1].y);
    }

    float2 dA[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) dA[k] = make_float2(0.f, 0.f);
    float2 dDacc = make_float2(0.f, 0.f);
    float2 ga[8];  // a*g entering the current step from the later ones (after the last super-tile: the segment's carry out)

    // sum over the warp's 16 channel pairs of the per-lane vectors parked in `scr` ([kRB2][2][32] float4: step, state quad,
    // lane).  16 items (step, state half, quad) x 2 source halves of 8 lanes; rotated reads are conflict-free.
    // which = 0: dB, 1: dC.  Partials go to ws_bc[((row) * ntile_c + ctile) * 32 + which * 16 + n].
    auto reduce_block = [&](int tblk, int which) {
        __syncwarp();
        const int item = lane & 15, part = lane >> 4;
        const int uu = item >> 2, ihs = (item >> 1) & 1, q = item & 1;
        const float4 *src = scr + (uu * 2 + q) * 32 + ihs * 16 + part * 8;
        float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f), u0 = make_float2(0.f, 0.f), u1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float4 v = src[(i + lane) & 7], w = src[(i + 1 + lane) & 7];
            s0 = add2(s0, make_float2(v.x, v.y));
            s1 = add2(s1, make_float2(v.z, v.w));
            u0 = add2(u0, make_float2(w.x, w.y));
            u1 = add2(u1, make_float2(w.z, w.w));
        }
        s0 = add2(s0, u0);
        s1 = add2(s1, u1);
        s0.x += __shfl_xor_sync(0xffffffffu, s0.x, 16);
        s0.y += __shfl_xor_sync(0xffffffffu, s0.y, 16);
        s1.x += __shfl_xor_sync(0xffffffffu, s1.x, 16);
        s1.y += __shfl_xor_sync(0xffffffffu, s1.y, 16);
        const int t = tblk + uu;
        if (part == 0 && t < L)
            __stcs(reinterpret_cast<float4 *>(p.ws_bc + ((row_b + t) * sc.ntile_c + ctile) * (2 * N) + which * N + ihs * 8 + q * 4),
                   make_float4(s0.x, s0.y, s1.x, s1.y));
        __syncwarp();
    };

    for (int it = 0; it < ntiles; ++it, ++g) {
        const int s = g % Lay::STAGES, tj = tile_lo + ntiles - 1 - it, t0 = tj * ST;
        unsigned char *st = smem + size_t(s) * Lay::STAGE_BYTES;
        T *sx = reinterpret_cast<T *>(st) + tb * CH + cl, *sd = sx + ST * CH, *sg = sd + ST * CH, *sz = sg + ST * CH;
        float *sdy, *se;  // fp32 dy and dz factor: in place over dout / z for fp32 I/O, separate arrays for 16-bit I/O
        if constexpr (sizeof(T) == 2) {
            sdy = reinterpret_cast<float *>(smem + Lay::DYE_OFF) + tb * CH + cl;
            se = sdy + ST * CH;
        } else {
            sdy = reinterpret_cast<float *>(sg);
            se = reinterpret_cast<float *>(sz);
        }
        mbar_wait(&full[s], (g / Lay::STAGES) & 1);

        const float *fB, *fC;  // this warp's 16 rows of B / C in fp32, offset to the lane's state half
        if constexpr (sizeof(T) == 2) {
            const T *gB = reinterpret_cast<const T *>(st + 4 * Lay::TILE_BYTES) + tb * N;
            const T *gC = reinterpret_cast<const T *>(st + 4 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N;
            for (int i = lane; i < TC * N; i += 32) {
                bc32[i] = to_f32<T>(gB[i]);
                bc32[TC * N + i] = to_f32<T>(gC[i]);
            }
            fB = bc32 + 8 * hs;
            fC = bc32 + TC * N + 8 * hs;
        } else {
            fB = reinterpret_cast<const float *>(st + 4 * Lay::TILE_BYTES) + tb * N + 8 * hs;
            fC = reinterpret_cast<const float *>(st + 4 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N + 8 * hs;
        }

        // chunk checkpoint (state entering step t0 + tb), this lane's 8 states of both channels: issued now, used in phase 1
        // (scalar loads on purpose: a 128-bit load hands ptxas four consecutive registers of ONE channel, which it then keeps as
        // the home of the loop-carried state and re-pairs with MOVs on every step)
        float2 ck[8];
        {
            const bool inb = active && t0 + tb < L;
            const float *cp = p.chk + ((int64_t(b) * nchk + (inb ? (t0 + tb) / TC : 0)) * ED + (inb ? c : 0)) * N + 8 * hs;
#pragma unroll
            for (int k = 0; k < 8; ++k) ck[k] = inb ? make_float2(__ldcs(cp + k), __ldcs(cp + N + k)) : make_float2(0.f, 0.f);
        }

        // ---- prologue: the lane handles its channel pair at steps 2 i + hs ------------------------------------------
#pragma unroll
        for (int i = 0; i < TC / 2; ++i) {
            const int u = 2 * i + hs;
            if (softplus) {  // fused softplus(dt_proj(.)), models/mamba.py:203; rows past L stay 0 (identity steps)
                const float2 r = ld2<T>(sd + u * CH);
                const bool in = t0 + tb + u < L;
                st2<T>(sd + u * CH, make_float2(in ? softplus_fast(r.x) : 0.f, in ? softplus_fast(r.y) : 0.f));
            }
            const float2 gv = ld2<T>(sg + u * CH);
            if constexpr (HAS_Z) {
                const float2 zv = ld2<T>(sz + u * CH);
                const float2 sgm = make_float2(sigmoidf_fast(zv.x), sigmoidf_fast(zv.y));
                const float2 gs = mul2(gv, sgm);
                *reinterpret_cast<float2 *>(sdy + u * CH) = mul2(gs, zv);
                *reinterpret_cast<float2 *>(se + u * CH) = mul2(gs, fma2(zv, make_float2(1.f - sgm.x, 1.f - sgm.y), make_float2(1.f, 1.f)));
            } else if constexpr (sizeof(T) == 2) {
                *reinterpret_cast<float2 *>(sdy + u * CH) = gv;
            }
        }
        __syncwarp();

        // ---- phase 1: recompute h, park it, y -> dz, dC partials, reverse-scan summary -----------------------------
        float2 h[8], P[8], acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            h[k] = ck[k];
            P[k] = make_float2(1.f, 1.f);
            acc[k] = make_float2(0.f, 0.f);
        }
#pragma unroll 1
        for (int ub = 0; ub < TC; ub += kRB2) {
            float2 yp[kRB2];              // partial y (this lane's 8 states) of the block's steps
#pragma unroll
            for (int uu = 0; uu < kRB2; ++uu) {
                const int u = ub + uu;
                const float2 xv = ld2<T>(sx + u * CH), dv = ld2<T>(sd + u * CH);
                const float2 dy = *reinterpret_cast<const float2 *>(sdy + u * CH);
                float Bv[8], Cv[8], pc[8];
                ld8(fB + u * N, Bv);
                ld8(fC + u * N, Cv);
                tmem_st16(tslot + uint32_t(u * 16), h);  // slot u = state ENTERING step u
                float2 a[8];
                decay8<GEOM>(dv, A2b, A2p, up, a);
                const float2 dx = mul2(dv, xv);
                float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    h[k] = fma2(a[k], h[k], mul2(dx, splat2(Bv[k])));
                    if (k & 1) yb = fma2(h[k], splat2(Cv[k]), yb);
                    else ya = fma2(h[k], splat2(Cv[k]), ya);
                    pc[k] = fmaf(dy.y, h[k].y, dy.x * h[k].x);  // dC partial, pre-added over the lane's two channels
                    P[k] = mul2(P[k], a[k]);
                    acc[k] = fma2(mul2(P[k], dy), splat2(Cv[k]), acc[k]);
                }
                yp[uu] = add2(ya, yb);
                scr[(uu * 2) * 32 + lane] = make_float4(pc[0], pc[1], pc[2], pc[3]);
                scr[(uu * 2 + 1) * 32 + lane] = make_float4(pc[4], pc[5], pc[6], pc[7]);
            }
            if constexpr (HAS_Z) {
                // the lane finishes steps ub + 2 j + hs: it keeps its own partial of those and receives the other state
                // half's; the exchange is batched per block so that no shuffle latency sits on the per-step critical path
#pragma unroll
                for (int j = 0; j < kRB2 / 2; ++j) {
                    const float2 mine = hs ? yp[2 * j + 1] : yp[2 * j], theirs = hs ? yp[2 * j] : yp[2 * j + 1];
                    const float ox = __shfl_xor_sync(0xffffffffu, theirs.x, 16), oy = __shfl_xor_sync(0xffffffffu, theirs.y, 16);
                    const int u = ub + 2 * j + hs;
                    const float2 xv = ld2<T>(sx + u * CH), ee = *reinterpret_cast<const float2 *>(se + u * CH);
                    const float2 y = fma2(Dd, xv, make_float2(mine.x + ox, mine.y + oy));
                    st2<T>(sz + u * CH, mul2(y, ee));  // dz, in place over z
                }
            }
            reduce_block(t0 + tb + ub, 1);
        }
        tmem_wait_st();
        if (wt > 0) {  // chunk 0's summary is never used: its a*g after phase 2 is the carry itself
            float4 *o = sums + (wt - 1) * 8 * 32 + lane;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                o[q * 32] = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
                o[(4 + q) * 32] = make_float4(P[2 * q].x, P[2 * q].y, P[2 * q + 1].x, P[2 * q + 1].y);
            }
        }
        __syncthreads();  // summaries of this super-tile and the carry written at the end of the previous one are visible

        // the previous super-tile's output stores have had a whole phase to drain; its stage can be refilled
        if (threadIdx.x == 0 && it >= 1 && it - 1 + Lay::STAGES < ntiles) {
            bulk_wait_read<0>();
            issue((g - 1) % Lay::STAGES, tile_lo + ntiles - 1 - (it - 1 + Lay::STAGES));
        }

        // ---- fold: a*g entering this chunk = carry chained through the later chunks of the super-tile ---------------
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w = carry[((g & 1) * 4 + q) * 32];
            ga[2 * q] = make_float2(w.x, w.y);
            ga[2 * q + 1] = make_float2(w.z, w.w);
        }
#pragma unroll
        for (int v = kNW - 1; v >= 1; --v) {  // unrolled with a warp-uniform guard: the loads of all summaries go out at once
            if (v > wt) {
                const float4 *o = sums + (v - 1) * 8 * 32 + lane;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 G = o[q * 32], Pe = o[(4 + q) * 32];
                    ga[2 * q] = fma2(make_float2(Pe.x, Pe.y), ga[2 * q], make_float2(G.x, G.y));
                    ga[2 * q + 1] = fma2(make_float2(Pe.z, Pe.w), ga[2 * q + 1], make_float2(G.z, G.w));
                }
            }
        }

        // ---- phase 2: reverse scan --------------------------------------------------------------------------------
#pragma unroll 1
        for (int ub = TC - kRB2; ub >= 0; ub -= kRB2) {
            float2 ddp[kRB2], gBp[kRB2];  // partial sums over this lane's 8 states for the block's steps
#pragma unroll
            for (int uu = kRB2 - 1; uu >= 0; --uu) {
                const int u = ub + uu;
                float2 hp[8];
                tmem_ld16(tslot + uint32_t(u * 16), hp);
                const float2 xv = ld2<T>(sx + u * CH), dv = ld2<T>(sd + u * CH);
                const float2 dy = *reinterpret_cast<const float2 *>(sdy + u * CH);
                float Bv[8], Cv[8], pb[8];
                ld8(fB + u * N, Bv);
                ld8(fC + u * N, Cv);
                float2 a[8];
                decay8<GEOM>(dv, A2b, A2p, up, a);
                const float2 dxw = mul2(dv, xv);
                float2 dda = make_float2(0.f, 0.f), ddb = make_float2(0.f, 0.f);
                float2 gBa = make_float2(0.f, 0.f), gBb = make_float2(0.f, 0.f);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float2 gk = fma2(dy, splat2(Cv[k]), ga[k]);  // g[t] = C dy + a[t+1] g[t+1]
                    const float2 ag = mul2(a[k], gk);                 // a[t] g[t]   (carried to step t-1)
                    const float2 w = mul2(hp[k], ag);                 // h[t-1] a g
                    const float2 Aw = GEOM ? splat2(kw[k]) : A2p[k];
                    if (k & 1) {
                        ddb = fma2(w, Aw, ddb);
                        gBb = fma2(gk, splat2(Bv[k]), gBb);
                    } else {
                        dda = fma2(w, Aw, dda);
                        gBa = fma2(gk, splat2(Bv[k]), gBa);
                    }
                    dA[k] = fma2(w, dv, dA[k]);
                    pb[k] = fmaf(dxw.y, gk.y, dxw.x * gk.x);  // dB partial, pre-added over the lane's two channels
                    ga[k] = ag;
                }
                ddp[uu] = add2(dda, ddb);
                gBp[uu] = add2(gBa, gBb);
                scr[(uu * 2) * 32 + lane] = make_float4(pb[0], pb[1], pb[2], pb[3]);
                scr[(uu * 2 + 1) * 32 + lane] = make_float4(pb[4], pb[5], pb[6], pb[7]);
            }
            // the lane finishes steps ub + 2 j + hs (dx, ddelta, dD of its channel pair): one batched exchange per block
#pragma unroll
            for (int j = 0; j < kRB2 / 2; ++j) {
                const float2 md = hs ? ddp[2 * j + 1] : ddp[2 * j], td = hs ? ddp[2 * j] : ddp[2 * j + 1];
                const float2 mg = hs ? gBp[2 * j + 1] : gBp[2 * j], tg = hs ? gBp[2 * j] : gBp[2 * j + 1];
                const float o0 = __shfl_xor_sync(0xffffffffu, td.x, 16), o1 = __shfl_xor_sync(0xffffffffu, td.y, 16);
                const float o2 = __shfl_xor_sync(0xffffffffu, tg.x, 16), o3 = __shfl_xor_sync(0xffffffffu, tg.y, 16);
                const int u = ub + 2 * j + hs;
                const float2 xv = ld2<T>(sx + u * CH), dv = ld2<T>(sd + u * CH);
                const float2 dy = *reinterpret_cast<const float2 *>(sdy + u * CH);
                const float2 dd = mul2(make_float2(md.x + o0, md.y + o1), ddscale), gB = make_float2(mg.x + o2, mg.y + o3);
                float2 odd = fma2(gB, xv, dd);
                if (softplus) {  // gradient w.r.t. the pre-activation
                    odd.x *= softplus_grad_from_value(dv.x);
                    odd.y *= softplus_grad_from_value(dv.y);
                }
                dDacc = fma2(dy, xv, dDacc);
                st2<T>(sx + u * CH, fma2(gB, dv, mul2(Dd, dy)));  // dx, in place over x
                st2<T>(sd + u * CH, odd);                         // ddelta, in place over delta
            }
            reduce_block(t0 + tb + ub, 0);
        }
        if (wt == 0) {  // carry for the next (earlier) super-tile
#pragma unroll
            for (int q = 0; q < 4; ++q)
                carry[(((g + 1) & 1) * 4 + q) * 32] = make_float4(ga[2 * q].x, ga[2 * q].y, ga[2 * q + 1].x, ga[2 * q + 1].y);
        }
        fence_proxy_async();  // generic-proxy writes of the in-place output tiles -> visible to the TMA engine
        __syncthreads();      // every warp is done with stage s
        if (threadIdx.x == 0) {
            tma_store_3d(&tm.odx, c0, t0, b, st);
            tma_store_3d(&tm.odd, c0, t0, b, st + Lay::TILE_BYTES);
            if (HAS_Z) tma_store_3d(&tm.odz, c0, t0, b, st + 3 * Lay::TILE_BYTES);
            bulk_commit();
        }
    }

    // hand the carry to the earlier segment of this chain, then raise its flag
    if (wt == 0 && seg > 0) {
        float2 *gc = reinterpret_cast<float2 *>(sc.carry) + int64_t(chain) * 8 * 32 + lane;
#pragma unroll
        for (int k = 0; k < 8; ++k) __stcg(gc + k * 32, ga[k]);
        __threadfence();
        __syncwarp();
        if (lane == 0) red_release_add(sc.done + chain, 1u);
    }

    // dA (A2 is A log2 e: dA = sum w delta needs no rescale) and dD of the item: summed over the 8 chunk-warps, one partial
    // per (batch, segment)
    float2 *red = reinterpret_cast<float2 *>(smem + Lay::SCR_OFF);  // [kNW][8][32] | [kNW][16]
    float2 *redD = red + kNW * 8 * 32;  // [kNW][32]: both state halves hold dD of their own steps
#pragma unroll
    for (int k = 0; k < 8; ++k) red[(wt * 8 + k) * 32 + lane] = dA[k];
    redD[wt * 32 + lane] = dDacc;
    __syncthreads();
    {
        const int k = threadIdx.x >> 5, l = threadIdx.x & 31;
        float2 sA = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < kNW; ++w) sA = add2(sA, red[(w * 8 + k) * 32 + l]);
        const int cc = c0 + 2 * (l & 15), n = 8 * (l >> 4) + k;
        if (cc < ED) {
            float *o = p.ws_ad + ((int64_t(b) * sc.nseg + seg) * ED + cc) * (N + 1);
            o[n] = sA.x;
            o[N + 1 + n] = sA.y;
        }
        if (threadIdx.x < 16) {
            float2 sD = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < kNW; ++w) sD = add2(sD, add2(redD[w * 32 + threadIdx.x], redD[w * 32 + 16 + threadIdx.x]));
            const int cd = c0 + 2 * int(threadIdx.x);
            if (cd < ED) {
                float *o = p.ws_ad + ((int64_t(b) * sc.nseg + seg) * ED + cd) * (N + 1);
                o[N] = sD.x;
                o[N + 1 + N] = sD.y;
            }
        }
    }
    __syncthreads();  // the scratch is free for the next item
}

template <typename T, bool HAS_Z>
__global__ void __launch_bounds__(kNW * 32, 1) selscan_bwd2_kernel(const Bwd2Params pp, const __grid_constant__ Bwd2Maps tm) {
    using Lay = Bwd2Layout<T>;
    constexpr int N = kN;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Lay::BAR_OFF + Lay::STAGES * sizeof(uint64_t));
    unsigned *ticket_s = reinterpret_cast<unsigned *>(tmem_slot + 1);
    const BwdParams &p = pp.b;
    const SegSched &sc = pp.s;
    const int tid = threadIdx.x, wt = tid >> 5, lane = tid & 31, pr = lane & 15, hs = lane >> 4;

    if (wt == 0) {  // all 512 tensor-memory columns: the state history of 128 steps x 32 channels
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < Lay::STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    int g = 0;  // super-tiles processed so far by this CTA: stage / mbarrier-parity / carry-buffer bookkeeping
    for (;;) {
        if (tid == 0) *ticket_s = atomicAdd(sc.ticket, 1u);
        __syncthreads();
        const int v = int(*ticket_s);
        if (v >= sc.nitems) break;
        // dependency order: every chain's LAST segment of L first (segi = 0), then the one before it, ...
        const int segi = v / sc.nchains, chain = v % sc.nchains;
        const int b = chain / sc.ntile_c, ctile = chain % sc.ntile_c, seg = sc.nseg - 1 - segi;
        const int c0 = ctile * kCH, c = c0 + 2 * pr;
        const bool active = c < p.ED;
        const int cc = active ? c : p.ED - 2;
        const float *Ar = p.A + int64_t(cc) * N;
        const float2 A2b = make_float2(Ar[0] * kLog2e, Ar[N] * kLog2e);
        const float2 Dd = make_float2(p.D[cc], p.D[cc + 1]);
        float2 A2p[8];
        float kw[8];
        bool ok = !(p.flags & MMI_FLAG_NO_GEOM);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int n = 8 * hs + k;
            A2p[k] = make_float2(Ar[n] * kLog2e, Ar[N + n] * kLog2e);
            kw[k] = float(n + 1);
            const float w0 = kw[k] * A2b.x, w1 = kw[k] * A2b.y;
            ok = ok && (fabsf(A2p[k].x - w0) <= 2e-6f * fabsf(w0)) && (fabsf(A2p[k].y - w1) <= 2e-6f * fabsf(w1));
        }
        const bool geom = __syncthreads_and(ok);  // also: everyone has read the ticket before thread 0 takes the next one
        if (geom) bwd2_item<T, true, HAS_Z>(pp, tm, smem, tmem_base, A2p, A2b, Dd, kw, c0, b, seg, segi, chain, ctile, wt, lane, active, g);
        else bwd2_item<T, false, HAS_Z>(pp, tm, smem, tmem_base, A2p, A2b, Dd, kw, c0, b, seg, segi, chain, ctile, wt, lane, active, g);
    }
    if (tid == 0) bulk_wait_read<0>();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wt == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// Deterministic reduction of the workspace partials: dB / dC over channel tiles (one thread per float4 of a 32-float row),
// dA / dD over (batch, segment).
template <typename T>
__global__ void __launch_bounds__(256) selscan_bwd2_finish_kernel(const float *__restrict__ ws_bc, const float *__restrict__ ws_ad,
                                                                  T *dBm, T *dCm, float *dA, float *dD, int64_t rows, int ntile,
                                                                  int nparts, int ED) {
    constexpr int N = kN;
    const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t n_bc = rows * 8;  // float4 groups: 2 * N / 4 per row
    if (gid < n_bc) {
        const int64_t row = gid >> 3;
        const int q = int(gid & 7);
        const float4 *src = reinterpret_cast<const float4 *>(ws_bc + row * ntile * (2 * N)) + q;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int tI = 0; tI < ntile; ++tI) {
            const float4 v = __ldcs(src + tI * (2 * N / 4));
            s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
        }
        T *dst = (q < 4 ? dBm : dCm) + row * N + (q & 3) * 4;
        dst[0] = from_f32<T>(s.x), dst[1] = from_f32<T>(s.y), dst[2] = from_f32<T>(s.z), dst[3] = from_f32<T>(s.w);
        return;
    }
    const int64_t g2 = gid - n_bc;
    if (g2 < int64_t(ED) * (N + 1)) {
        float v = 0.f;
        for (int bI = 0; bI < nparts; ++bI) v += ws_ad[int64_t(bI) * ED * (N + 1) + g2];
        const int c = int(g2 / (N + 1)), n = int(g2 % (N + 1));
        if (n < N) dA[int64_t(c) * N + n] = v;
        else dD[c] = v;
    }
}

constexpr int kMaxSeg2 = 16;
static size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

// L segments per chain: as few as fill the SMs (each extra segment costs one carry hand-off through global memory)
int seg_sched_plan(int B, int L, int ED, int flags, SegSched *s) {
    const int ntile_c = (ED + kCH - 1) / kCH, ntiles = (L + kST - 1) / kST;
    const int64_t nchains = int64_t(B) * ntile_c;
    const int slots = sm_count();
    int nseg = (flags & MMI_FLAG_NSEG_MASK) >> MMI_FLAG_NSEG_SHIFT;
    if (!nseg) {
        nseg = 1;
        if (nchains > slots) {  // more chains than SMs: segments smooth the last wave
            double best = 0.0;
            for (int cand = 1; cand <= 8; ++cand) {
                if (ntiles / cand < 4 && cand > 1) break;
                const int64_t items = nchains * cand, waves = (items + slots - 1) / slots;
                const double eff = double(items) / double(waves * slots);
                if (eff > best + 0.02) best = eff, nseg = cand;
            }
        }
    }
    nseg = std::max(1, std::min({nseg, kMaxSeg2, ntiles}));
    s->seg_tiles = (ntiles + nseg - 1) / nseg;
    s->nseg = (ntiles + s->seg_tiles - 1) / s->seg_tiles;
    s->ntile_c = ntile_c;
    s->nchains = int(nchains);
    s->nitems = int(nchains * s->nseg);
    return MMI_OK;
}

// workspace: [dB/dC partials (B, L, ntile_c, 2N)] [dA/dD partials (B, nseg, ED, N+1)] [ticket | done (nchains)] [carry (nchains, 8, 32) float2]
int64_t selscan_bwd2_ws_bytes(int B, int L, int ED) {
    const int64_t ntile = (ED + kCH - 1) / kCH, nch = int64_t(B) * ntile;
    return int64_t(al256(size_t(B) * L * ntile * 2 * kN * 4)) + int64_t(al256(size_t(B) * kMaxSeg2 * ED * (kN + 1) * 4)) +
           int64_t(al256(16 + size_t(nch) * 4)) + nch * 8 * 32 * 8;
}

// Host side shared by the 8-warp and the 16-warp kernel (selscan_bwd3.cu): plan the segments, carve the workspace, encode
// the tensor maps, clear the ticket / flags.
int bwd2_prepare(Bwd2Params &pp, Bwd2Maps &tm, int dtype, void *ws, cudaStream_t st) {
    BwdParams &p = pp.b;
    const size_t es = dtype == MMI_F32 ? 4 : 2;
    const bool has_z = p.z != nullptr;
    const uint64_t rows = uint64_t(p.B) * p.L, nb = p.B, L = p.L;
    if (int e = seg_sched_plan(p.B, p.L, p.ED, p.flags, &pp.s)) return e;
    SegSched &sc = pp.s;
    p.ntile_c = sc.ntile_c;
    char *w = static_cast<char *>(ws);
    p.ws_bc = reinterpret_cast<float *>(w);
    w += al256(size_t(rows) * sc.ntile_c * 2 * kN * 4);
    p.ws_ad = reinterpret_cast<float *>(w);
    w += al256(size_t(p.B) * kMaxSeg2 * p.ED * (kN + 1) * 4);
    sc.ticket = reinterpret_cast<unsigned *>(w);
    sc.done = sc.ticket + 4;
    const size_t hdr = al256(16 + size_t(sc.nchains) * 4);
    sc.carry = reinterpret_cast<float *>(w + hdr);
    memset(&tm, 0, sizeof(tm));
    if (int e = make_tmap_3d(&tm.x, p.x, dtype, nb, L, p.ED, p.x_ld * es, kST, kCH)) return e;
    if (int e = make_tmap_3d(&tm.d, p.delta, dtype, nb, L, p.ED, p.d_ld * es, kST, kCH)) return e;
    if (int e = make_tmap_3d(&tm.g, p.dout, dtype, nb, L, p.ED, p.g_ld * es, kST, kCH)) return e;
    if (has_z)
        if (int e = make_tmap_3d(&tm.z, p.z, dtype, nb, L, p.ED, p.z_ld * es, kST, kCH)) return e;
    if (int e = make_tmap_3d(&tm.B, p.Bm, dtype, nb, L, kN, kN * es, kST, kN)) return e;
    if (int e = make_tmap_3d(&tm.C, p.Cm, dtype, nb, L, kN, kN * es, kST, kN)) return e;
    if (int e = make_tmap_3d(&tm.odx, p.dx, dtype, nb, L, p.ED, p.ED * es, kST, kCH)) return e;
    if (int e = make_tmap_3d(&tm.odd, p.ddelta, dtype, nb, L, p.ED, p.ED * es, kST, kCH)) return e;
    if (has_z)
        if (int e = make_tmap_3d(&tm.odz, p.dz, dtype, nb, L, p.ED, p.ED * es, kST, kCH)) return e;
    return check_cuda(cudaMemsetAsync(sc.ticket, 0, hdr, st), "selscan_bwd2 ticket memset");
}

template <typename T> static int finish_t(const Bwd2Params &pp, cudaStream_t st) {
    const BwdParams &p = pp.b;
    const int64_t rows = int64_t(p.B) * p.L, work = rows * 8 + int64_t(p.ED) * (kN + 1);
    selscan_bwd2_finish_kernel<T><<<unsigned((work + 255) / 256), 256, 0, st>>>(
        p.ws_bc, p.ws_ad, static_cast<T *>(p.dBm), static_cast<T *>(p.dCm), p.dA, p.dD, rows, pp.s.ntile_c, p.B * pp.s.nseg, p.ED);
    return check_cuda(cudaGetLastError(), "selscan_bwd2 finish launch");
}
int bwd2_finish(const Bwd2Params &pp, int dtype, cudaStream_t st) {
    switch (dtype) {
        case MMI_F32: return finish_t<float>(pp, st);
        case MMI_BF16: return finish_t<__nv_bfloat16>(pp, st);
        case MMI_F16: return finish_t<__half>(pp, st);
    }
    set_error("selscan_bwd2: unknown dtype %d", dtype);
    return MMI_ERR_ARG;
}

template <typename T, bool HAS_Z> static int launch_bwd2_t(Bwd2Params pp, int dtype, void *ws, cudaStream_t st) {
    using Lay = Bwd2Layout<T>;
    auto kern = selscan_bwd2_kernel<T, HAS_Z>;
    static thread_local int attr_dev = -1;  // the opt-in is per device and sticky: set it once, not on every launch
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Lay::SMEM)),
                               "selscan_bwd2 smem attribute"))
            return e;
        attr_dev = dev;
    }
    Bwd2Maps tm;
    if (int e = bwd2_prepare(pp, tm, dtype, ws, st)) return e;
    const int grid = std::min(pp.s.nitems, sm_count());
    kern<<<grid, kNW * 32, Lay::SMEM, st>>>(pp, tm);
    if (int e = check_cuda(cudaGetLastError(), "selscan_bwd2 launch")) return e;
    return bwd2_finish(pp, dtype, st);
}

int selscan_bwd2_launch(const BwdParams &p, int dtype, void *ws, cudaStream_t st) {
    Bwd2Params pp;
    memset(&pp, 0, sizeof(pp));
    pp.b = p;
    const bool z = p.z != nullptr;
    switch (dtype) {
        case MMI_F32: return z ? launch_bwd2_t<float, true>(pp, dtype, ws, st) : launch_bwd2_t<float, false>(pp, dtype, ws, st);
        case MMI_BF16:
            return z ? launch_bwd2_t<__nv_bfloat16, true>(pp, dtype, ws, st) : launch_bwd2_t<__nv_bfloat16, false>(pp, dtype, ws, st);
        case MMI_F16: return z ? launch_bwd2_t<__half, true>(pp, dtype, ws, st) : launch_bwd2_t<__half, false>(pp, dtype, ws, st);
    }
    set_error("selscan_bwd2: unknown dtype %d", dtype);
    return MMI_ERR_ARG;
}

}  // namespace mmi

// selscan_bwd3.cu -- fused selective scan backward, 16-warp form of the second generation (sm_100a).
//
// Same mathematics, same chains / super-tiles / persistent ticket schedule / workspace / finishing kernel as selscan_bwd2.cu
// (see there and selscan2.cuh).  What changes is the lane layout: selscan_bwd2.cu keeps 8 states x 2 channels per lane, which
// needs 255 registers and leaves two warps per scheduler -- every warp then runs at its own dependent-issue pace (4.8 cycles
// per instruction, profiles/r02_scan_generations.txt: time follows the instruction count, no pipe is saturated).  Here a lane
// keeps 4 states x 2 channels in <= 128 registers, so that 16 warps (4 per scheduler) are resident:
//   warp      w = 2 * wt + ch:  chunk wt (16 steps of the 128-step super-tile) x channel half ch (16 of the chain's 32 channels)
//   lane      pr = lane & 7 -> channel pair (c0 + 16 ch + 2 pr, + 1);  q = lane >> 3 -> states 4 q .. 4 q + 3
//   block     4 steps; lane q finishes step ub + q of its pair (y -> dz in phase 1; ddelta, dx, dD in phase 2) from the four
//             quarters' partial sums, exchanged through the warp's own rows of the chunk scratch
//   dB / dC   pre-added over the lane's two channels, parked in the chunk scratch [step][ch][lane] and summed over the 16 pairs
//             of BOTH warps of the chunk (two 64-thread named barriers per block); warp ch reduces steps 2 ch, 2 ch + 1
// Tensor memory: warp w owns lanes 32 (w & 3) .. + 31, columns 128 (w >> 2) .. + 127 (16 steps x 8 fp32).
#include <algorithm>
#include <cstring>

#include "../../include/mmidet_b200.h"
#include "selscan.h"
#include "selscan2.cuh"

namespace mmi {

using namespace v2;

constexpr int kNW3 = 2 * kNW;  // warps per CTA
constexpr int kRB3 = 4;        // steps per block == state quarters

template <typename T> struct Bwd3Layout {
    static constexpr int N = kN, STAGES = 2;
    static constexpr size_t TILE_BYTES = size_t(kST) * kCH * sizeof(T);
    static constexpr size_t BCT_BYTES = size_t(kST) * N * sizeof(T);
    static constexpr size_t STAGE_BYTES = 4 * TILE_BYTES + 2 * BCT_BYTES;  // x | delta | dout | z | B | C
    static constexpr size_t DYE_OFF = STAGES * STAGE_BYTES;                // fp32 dy | e when T is 16 bit
    static constexpr size_t DYE_BYTES = sizeof(T) == 2 ? size_t(2) * kST * kCH * 4 : 0;
    static constexpr size_t BC32_OFF = DYE_OFF + DYE_BYTES;                // per-chunk widened B | C rows when T is 16 bit
    static constexpr size_t BC32_BYTES = sizeof(T) == 2 ? size_t(kNW) * 2 * kTC * N * 4 : 0;
    static constexpr size_t SCR_OFF = BC32_OFF + BC32_BYTES;               // per chunk [kRB3][2][32] float4
    static constexpr size_t SCR_CHUNK = size_t(kRB3) * 2 * 32 * 16;
    static constexpr size_t SUM_OFF = SCR_OFF + kNW * SCR_CHUNK;           // chunk summaries [kNW - 1][4][64] float4 (G | P_end)
    static constexpr size_t SUM_CHUNK = size_t(4) * 64 * 16;
    static constexpr size_t CARRY_OFF = SUM_OFF + (kNW - 1) * SUM_CHUNK;   // carried a*g [2][2][64] float4
    static constexpr size_t BAR_OFF = CARRY_OFF + size_t(2) * 2 * 64 * 16;
    static constexpr size_t SMEM = BAR_OFF + STAGES * sizeof(uint64_t) + 16;
    static_assert(SMEM <= 232448, "shared memory budget of one CTA per SM");
    static_assert(kNW * SCR_CHUNK >= size_t(kNW3) * 4 * 32 * 8 + size_t(kNW3) * 32 * 8, "dA / dD reduction aliases the scratch");
};

// the two warps of a chunk (64 threads) meet on named barrier 1 + wt
__device__ __forceinline__ void pair_bar(int wt) { asm volatile("bar.sync %0, 64;" ::"r"(wt + 1) : "memory"); }

template <typename T, bool GEOM, bool HAS_Z>
__device__ __forceinline__ void bwd3_item(const Bwd2Params &pp, const Bwd2Maps &tm, unsigned char *smem, uint32_t tmem_base,
                                          const float2 (&A2p)[4], float2 A2b, float2 A2q, float2 Dd, const float (&kw)[4], int c0,
                                          int b, int seg, int segi, int chain, int ctile, int warp, int lane, bool active, int &g) {
    using Lay = Bwd3Layout<T>;
    constexpr int N = kN, TC = kTC, ST = kST, CH = kCH;
    constexpr float kLn2 = 0.6931471805599453f;
    const BwdParams &p = pp.b;
    const SegSched &sc = pp.s;
    const int wt = warp >> 1, ch = warp & 1, pr = lane & 7, q = lane >> 3, l64 = ch * 32 + lane;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    float4 *cscr = reinterpret_cast<float4 *>(smem + Lay::SCR_OFF + wt * Lay::SCR_CHUNK);  // [kRB3][2][32]
    float4 *wrow = cscr + ch * 32 + lane;                                                  // this warp's row of step uu: + uu * 64
    float4 *sums = reinterpret_cast<float4 *>(smem + Lay::SUM_OFF) + l64;                  // + ((v - 1) * 4 + j) * 64
    float4 *carry = reinterpret_cast<float4 *>(smem + Lay::CARRY_OFF) + l64;               // + (buf * 2 + j) * 64
    float *bc32 = reinterpret_cast<float *>(smem + Lay::BC32_OFF) + wt * 2 * TC * N;
    const uint32_t tslot = tmem_base + (uint32_t(warp & 3) * 32u << 16) + uint32_t(warp >> 2) * (TC * 8);

    const int L = p.L, ED = p.ED;
    const int ntiles_all = (L + ST - 1) / ST, nchk = (L + TC - 1) / TC;
    const int tile_lo = seg * sc.seg_tiles, ntiles = min(sc.seg_tiles, ntiles_all - tile_lo);
    const int tb = wt * TC, cl = ch * 16 + 2 * pr, c = c0 + cl;
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
        float2 cin[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) cin[k] = make_float2(0.f, 0.f);
        if (segi > 0) {
            if (lane == 0) {
                const long long tw = clock64();
                while (ld_acquire(sc.done + chain) < unsigned(segi)) {
                    __nanosleep(64);
                    if (clock64() - tw > 20000000000LL) __trap();  // ~10 s: a lost predecessor traps instead of hanging
                }
            }
            __syncwarp();
            const float2 *gc = reinterpret_cast<const float2 *>(sc.carry) + int64_t(chain) * 4 * 64 + l64;
#pragma unroll
            for (int k = 0; k < 4; ++k) cin[k] = __ldcg(gc + k * 64);
        }
        carry[((g & 1) * 2) * 64] = make_float4(cin[0].x, cin[0].y, cin[1].x, cin[1].y);
        carry[((g & 1) * 2 + 1) * 64] = make_float4(cin[2].x, cin[2].y, cin[3].x, cin[3].y);
    }

    float2 dA[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) dA[k] = make_float2(0.f, 0.f);
    float2 dDacc = make_float2(0.f, 0.f);
    float2 ga[4];  // a*g entering the current step from the later ones (after the last super-tile: the segment's carry out)

    // Sum of the chunk scratch over the 16 channel pairs of both warps.  This warp reduces steps 2 ch + u' of the block;
    // lane = (source warp sg, u', state quarter qq, half jh of the quarter's float4): 8 LDS.64 (rotated: conflict-free), one
    // exchange between the source halves, 16 lanes store 64 contiguous bytes per step.  which = 0: dB, 1: dC.
    auto reduce_block = [&](int tblk, int which) {
        pair_bar(wt);  // both warps' partials of the block are in the scratch
        const int sg = lane >> 4, up = (lane >> 3) & 1, qq = (lane >> 1) & 3, jh = lane & 1;
        const int us = 2 * ch + up;
        const float2 *src = reinterpret_cast<const float2 *>(cscr + (us * 2 + sg) * 32 + qq * 8) + jh;
        const int rot = qq + 4 * up;
        float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            s0 = add2(s0, src[2 * ((i + rot) & 7)]);
            s1 = add2(s1, src[2 * ((i + 1 + rot) & 7)]);
        }
        s0 = add2(s0, s1);
        s0.x += __shfl_xor_sync(0xffffffffu, s0.x, 16);
        s0.y += __shfl_xor_sync(0xffffffffu, s0.y, 16);
        const int t = tblk + us;
        if (sg == 0 && t < L)
            __stcs(reinterpret_cast<float2 *>(p.ws_bc + ((row_b + t) * sc.ntile_c + ctile) * (2 * N) + which * N + qq * 4 + jh * 2), s0);
        pair_bar(wt);  // the scratch may be overwritten
    };
    // partial sums of the block's 4 steps (this lane's 4 states) -> the full sum of step ub + q, for the lane's channel pair:
    // parked as [2][32] float4 in the warp's own scratch rows (row j = steps 2 j, 2 j + 1), read back across the 4 quarters
    auto park4 = [&](int j0, const float2 (&v)[kRB3]) {
        wrow[j0 * 64] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        wrow[(j0 + 1) * 64] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
    };
    auto gather4 = [&](int j0) {
        const float2 *src = reinterpret_cast<const float2 *>(cscr + ((j0 + (q >> 1)) * 2 + ch) * 32 + pr) + (q & 1);
        return add2(add2(src[0], src[16]), add2(src[32], src[48]));  // lanes pr, 8 + pr, 16 + pr, 24 + pr
    };

    for (int it = 0; it < ntiles; ++it, ++g) {
        const int s = g % Lay::STAGES, tj = tile_lo + ntiles - 1 - it, t0 = tj * ST;
        unsigned char *st = smem + size_t(s) * Lay::STAGE_BYTES;
        T *sx = reinterpret_cast<T *>(st) + tb * CH + cl, *sd = sx + ST * CH, *sg_ = sd + ST * CH, *sz = sg_ + ST * CH;
        float *sdy, *se;  // fp32 dy and dz factor: in place over dout / z for fp32 I/O, separate arrays for 16-bit I/O
        if constexpr (sizeof(T) == 2) {
            sdy = reinterpret_cast<float *>(smem + Lay::DYE_OFF) + tb * CH + cl;
            se = sdy + ST * CH;
        } else {
            sdy = reinterpret_cast<float *>(sg_);
            se = reinterpret_cast<float *>(sz);
        }
        mbar_wait(&full[s], (g / Lay::STAGES) & 1);

        const float *fB, *fC;  // this chunk's 16 rows of B / C in fp32, offset to the lane's state quarter
        if constexpr (sizeof(T) == 2) {
            const T *gB = reinterpret_cast<const T *>(st + 4 * Lay::TILE_BYTES) + tb * N;
            const T *gC = reinterpret_cast<const T *>(st + 4 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N;
            for (int i = ch * (TC * N / 2) + lane; i < (ch + 1) * (TC * N / 2); i += 32) {  // each warp widens half of the rows
                bc32[i] = to_f32<T>(gB[i]);
                bc32[TC * N + i] = to_f32<T>(gC[i]);
            }
            pair_bar(wt);
            fB = bc32 + 4 * q;
            fC = bc32 + TC * N + 4 * q;
        } else {
            fB = reinterpret_cast<const float *>(st + 4 * Lay::TILE_BYTES) + tb * N + 4 * q;
            fC = reinterpret_cast<const float *>(st + 4 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N + 4 * q;
        }

        // chunk checkpoint (state entering step t0 + tb), this lane's 4 states of both channels: issued now, used in phase 1
        float2 h[4];
        {
            const bool inb = active && t0 + tb < L;
            const float *cp = p.chk + ((int64_t(b) * nchk + (inb ? (t0 + tb) / TC : 0)) * ED + (inb ? c : 0)) * N + 4 * q;
#pragma unroll
            for (int k = 0; k < 4; ++k) h[k] = inb ? make_float2(__ldcs(cp + k), __ldcs(cp + N + k)) : make_float2(0.f, 0.f);
        }

        // ---- prologue: the lane handles its channel pair at steps 4 i + q ------------------------------------------
#pragma unroll
        for (int i = 0; i < TC / 4; ++i) {
            const int u = 4 * i + q;
            if (softplus) {  // fused softplus(dt_proj(.)), models/mamba.py:203; rows past L stay 0 (identity steps)
                const float2 r = ld2<T>(sd + u * CH);
                const bool in = t0 + tb + u < L;
                st2<T>(sd + u * CH, make_float2(in ? softplus_fast(r.x) : 0.f, in ? softplus_fast(r.y) : 0.f));
            }
            const float2 gv = ld2<T>(sg_ + u * CH);
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
        float2 P[4], acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            P[k] = make_float2(1.f, 1.f);
            acc[k] = make_float2(0.f, 0.f);
        }
#pragma unroll 1
        for (int ub = 0; ub < TC; ub += kRB3) {
            float2 yp[kRB3];  // partial y (this lane's 4 states) of the block's steps
#pragma unroll
            for (int uu = 0; uu < kRB3; ++uu) {
                const int u = ub + uu;
                const float2 xv = ld2<T>(sx + u * CH), dv = ld2<T>(sd + u * CH);
                const float2 dy = *reinterpret_cast<const float2 *>(sdy + u * CH);
                const float4 Bv = *reinterpret_cast<const float4 *>(fB + u * N), Cv = *reinterpret_cast<const float4 *>(fC + u * N);
                const float Bk[4] = {Bv.x, Bv.y, Bv.z, Bv.w}, Ck[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
                tmem_st8(tslot + uint32_t(u * 8), h);  // slot u = state ENTERING step u
                float2 a[4];
                decay4<GEOM>(dv, A2b, A2q, A2p, a);
                const float2 dx = mul2(dv, xv);
                float pc[4];
                float2 y;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    h[k] = fma2(a[k], h[k], mul2(dx, splat2(Bk[k])));
                    y = k ? fma2(h[k], splat2(Ck[k]), y) : mul2(h[k], splat2(Ck[k]));
                    pc[k] = fmaf(dy.y, h[k].y, dy.x * h[k].x);  // dC partial, pre-added over the lane's two channels
                    P[k] = mul2(P[k], a[k]);
                    acc[k] = fma2(mul2(P[k], dy), splat2(Ck[k]), acc[k]);
                }
                yp[uu] = y;
                wrow[uu * 64] = make_float4(pc[0], pc[1], pc[2], pc[3]);
            }
            reduce_block(t0 + tb + ub, 1);
            if constexpr (HAS_Z) {
                park4(0, yp);
                __syncwarp();
                const int u = ub + q;
                const float2 xv = ld2<T>(sx + u * CH), ee = *reinterpret_cast<const float2 *>(se + u * CH);
                const float2 y = fma2(Dd, xv, gather4(0));
                st2<T>(sz + u * CH, mul2(y, ee));  // dz, in place over z
                __syncwarp();
            }
        }
        tmem_wait_st();
        if (wt > 0) {  // chunk 0's summary is never used: its a*g after phase 2 is the carry itself
            float4 *o = sums + (wt - 1) * 4 * 64;
            o[0] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
            o[64] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
            o[128] = make_float4(P[0].x, P[0].y, P[1].x, P[1].y);
            o[192] = make_float4(P[2].x, P[2].y, P[3].x, P[3].y);
        }
        __syncthreads();  // summaries of this super-tile and the carry written at the end of the previous one are visible

        // the previous super-tile's output stores have had a whole phase to drain; its stage can be refilled
        if (threadIdx.x == 0 && it >= 1 && it - 1 + Lay::STAGES < ntiles) {
            bulk_wait_read<0>();
            issue((g - 1) % Lay::STAGES, tile_lo + ntiles - 1 - (it - 1 + Lay::STAGES));
        }

        // ---- fold: a*g entering this chunk = carry chained through the later chunks of the super-tile ---------------
        {
            const float4 w0 = carry[((g & 1) * 2) * 64], w1 = carry[((g & 1) * 2 + 1) * 64];
            ga[0] = make_float2(w0.x, w0.y), ga[1] = make_float2(w0.z, w0.w);
            ga[2] = make_float2(w1.x, w1.y), ga[3] = make_float2(w1.z, w1.w);
        }
#pragma unroll
        for (int v = kNW - 1; v >= 1; --v) {  // unrolled with a warp-uniform guard: the loads of all summaries go out at once
            if (v > wt) {
                const float4 *o = sums + (v - 1) * 4 * 64;
                const float4 G0 = o[0], G1 = o[64], P0 = o[128], P1 = o[192];
                ga[0] = fma2(make_float2(P0.x, P0.y), ga[0], make_float2(G0.x, G0.y));
                ga[1] = fma2(make_float2(P0.z, P0.w), ga[1], make_float2(G0.z, G0.w));
                ga[2] = fma2(make_float2(P1.x, P1.y), ga[2], make_float2(G1.x, G1.y));
                ga[3] = fma2(make_float2(P1.z, P1.w), ga[3], make_float2(G1.z, G1.w));
            }
        }

        // ---- phase 2: reverse scan --------------------------------------------------------------------------------
#pragma unroll 1
        for (int ub = TC - kRB3; ub >= 0; ub -= kRB3) {
            float2 ddp[kRB3], gBp[kRB3];  // partial sums over this lane's 4 states for the block's steps
#pragma unroll
            for (int uu = kRB3 - 1; uu >= 0; --uu) {
                const int u = ub + uu;
                float2 hp[4];
                tmem_ld8(tslot + uint32_t(u * 8), hp);
                const float2 xv = ld2<T>(sx + u * CH), dv = ld2<T>(sd + u * CH);
                const float2 dy = *reinterpret_cast<const float2 *>(sdy + u * CH);
                const float4 Bv = *reinterpret_cast<const float4 *>(fB + u * N), Cv = *reinterpret_cast<const float4 *>(fC + u * N);
                const float Bk[4] = {Bv.x, Bv.y, Bv.z, Bv.w}, Ck[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
                float2 a[4];
                decay4<GEOM>(dv, A2b, A2q, A2p, a);
                const float2 dxw = mul2(dv, xv);
                float2 dd, gB;
                float pb[4];
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 gk = fma2(dy, splat2(Ck[k]), ga[k]);  // g[t] = C dy + a[t+1] g[t+1]
                    const float2 ag = mul2(a[k], gk);                  // a[t] g[t]   (carried to step t-1)
                    const float2 w = mul2(hp[k], ag);                  // h[t-1] a g
                    const float2 Aw = GEOM ? splat2(kw[k]) : A2p[k];
                    dd = k ? fma2(w, Aw, dd) : mul2(w, Aw);
                    gB = k ? fma2(gk, splat2(Bk[k]), gB) : mul2(gk, splat2(Bk[k]));
                    dA[k] = fma2(w, dv, dA[k]);
                    pb[k] = fmaf(dxw.y, gk.y, dxw.x * gk.x);  // dB partial, pre-added over the lane's two channels
                    ga[k] = ag;
                }
                ddp[uu] = dd;
                gBp[uu] = gB;
                wrow[uu * 64] = make_float4(pb[0], pb[1], pb[2], pb[3]);
            }
            reduce_block(t0 + tb + ub, 0);
            // the lane finishes step ub + q (dx, ddelta, dD of its channel pair)
            park4(0, ddp);
            park4(2, gBp);
            __syncwarp();
            {
                const int u = ub + q;
                const float2 xv = ld2<T>(sx + u * CH), dv = ld2<T>(sd + u * CH);
                const float2 dy = *reinterpret_cast<const float2 *>(sdy + u * CH);
                const float2 dd = mul2(gather4(0), ddscale), gB = gather4(2);
                float2 odd = fma2(gB, xv, dd);
                if (softplus) {  // gradient w.r.t. the pre-activation
                    odd.x *= softplus_grad_from_value(dv.x);
                    odd.y *= softplus_grad_from_value(dv.y);
                }
                dDacc = fma2(dy, xv, dDacc);
                st2<T>(sx + u * CH, fma2(gB, dv, mul2(Dd, dy)));  // dx, in place over x
                st2<T>(sd + u * CH, odd);                         // ddelta, in place over delta
            }
            __syncwarp();
        }
        if (wt == 0) {  // carry for the next (earlier) super-tile
            carry[(((g + 1) & 1) * 2) * 64] = make_float4(ga[0].x, ga[0].y, ga[1].x, ga[1].y);
            carry[(((g + 1) & 1) * 2 + 1) * 64] = make_float4(ga[2].x, ga[2].y, ga[3].x, ga[3].y);
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
        float2 *gc = reinterpret_cast<float2 *>(sc.carry) + int64_t(chain) * 4 * 64 + l64;
#pragma unroll
        for (int k = 0; k < 4; ++k) __stcg(gc + k * 64, ga[k]);
        __threadfence();
        pair_bar(0);  // both channel halves have written
        if (ch == 0 && lane == 0) red_release_add(sc.done + chain, 1u);
    }

    // dA (A2 is A log2 e: dA = sum w delta needs no rescale) and dD of the item: summed over the 8 chunks, one partial per
    // (batch, segment)
    float2 *red = reinterpret_cast<float2 *>(smem + Lay::SCR_OFF);  // [kNW3][4][32]
    float2 *redD = red + kNW3 * 4 * 32;                             // [kNW3][32]: every quarter holds dD of its own steps
#pragma unroll
    for (int k = 0; k < 4; ++k) red[(warp * 4 + k) * 32 + lane] = dA[k];
    redD[warp * 32 + lane] = dDacc;
    __syncthreads();
    if (threadIdx.x < 256) {
        const int l = threadIdx.x & 31, k = (threadIdx.x >> 5) & 3, hh = threadIdx.x >> 7;
        float2 sA = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < kNW; ++w) sA = add2(sA, red[((2 * w + hh) * 4 + k) * 32 + l]);
        const int cc = c0 + hh * 16 + 2 * (l & 7), n = 4 * (l >> 3) + k;
        if (cc < ED) {
            float *o = p.ws_ad + ((int64_t(b) * sc.nseg + seg) * ED + cc) * (N + 1);
            o[n] = sA.x;
            o[N + 1 + n] = sA.y;
        }
    } else if (threadIdx.x < 256 + 16) {
        const int i = threadIdx.x - 256, hh = i >> 3, pp_ = i & 7;
        float2 sD = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < kNW; ++w)
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) sD = add2(sD, redD[(2 * w + hh) * 32 + qq * 8 + pp_]);
        const int cd = c0 + hh * 16 + 2 * pp_;
        if (cd < ED) {
            float *o = p.ws_ad + ((int64_t(b) * sc.nseg + seg) * ED + cd) * (N + 1);
            o[N] = sD.x;
            o[N + 1 + N] = sD.y;
        }
    }
    __syncthreads();  // the scratch is free for the next item
}

template <typename T, bool HAS_Z>
__global__ void __launch_bounds__(kNW3 * 32, 1) selscan_bwd3_kernel(const Bwd2Params pp, const __grid_constant__ Bwd2Maps tm) {
    using Lay = Bwd3Layout<T>;
    constexpr int N = kN;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Lay::BAR_OFF + Lay::STAGES * sizeof(uint64_t));
    unsigned *ticket_s = reinterpret_cast<unsigned *>(tmem_slot + 1);
    const BwdParams &p = pp.b;
    const SegSched &sc = pp.s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, pr = lane & 7, q = lane >> 3, ch = warp & 1;

    if (warp == 0) {  // all 512 tensor-memory columns: the state history of 128 steps x 32 channels
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
        const int c0 = ctile * kCH, c = c0 + ch * 16 + 2 * pr;
        const bool active = c < p.ED;
        const int cc = active ? c : p.ED - 2;
        const float *Ar = p.A + int64_t(cc) * N;
        const float2 A2b = make_float2(Ar[0] * kLog2e, Ar[N] * kLog2e);
        const float kq1 = float(4 * q + 1);
        const float2 A2q = make_float2(A2b.x * kq1, A2b.y * kq1);
        const float2 Dd = make_float2(p.D[cc], p.D[cc + 1]);
        float2 A2p[4];
        float kw[4];
        bool ok = !(p.flags & MMI_FLAG_NO_GEOM);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int n = 4 * q + k;
            A2p[k] = make_float2(Ar[n] * kLog2e, Ar[N + n] * kLog2e);
            kw[k] = float(n + 1);
            const float w0 = kw[k] * A2b.x, w1 = kw[k] * A2b.y;
            ok = ok && (fabsf(A2p[k].x - w0) <= 2e-6f * fabsf(w0)) && (fabsf(A2p[k].y - w1) <= 2e-6f * fabsf(w1));
        }
        const bool geom = __syncthreads_and(ok);  // also: everyone has read the ticket before thread 0 takes the next one
        if (geom) bwd3_item<T, true, HAS_Z>(pp, tm, smem, tmem_base, A2p, A2b, A2q, Dd, kw, c0, b, seg, segi, chain, ctile, warp, lane, active, g);
        else bwd3_item<T, false, HAS_Z>(pp, tm, smem, tmem_base, A2p, A2b, A2q, Dd, kw, c0, b, seg, segi, chain, ctile, warp, lane, active, g);
    }
    if (tid == 0) bulk_wait_read<0>();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <typename T, bool HAS_Z> static int launch_bwd3_t(Bwd2Params pp, int dtype, void *ws, cudaStream_t st) {
    using Lay = Bwd3Layout<T>;
    auto kern = selscan_bwd3_kernel<T, HAS_Z>;
    static thread_local int attr_dev = -1;  // the opt-in is per device and sticky: set it once, not on every launch
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Lay::SMEM)),
                               "selscan_bwd3 smem attribute"))
            return e;
        attr_dev = dev;
    }
    Bwd2Maps tm;
    if (int e = bwd2_prepare(pp, tm, dtype, ws, st)) return e;
    const int grid = std::min(pp.s.nitems, sm_count());
    kern<<<grid, kNW3 * 32, Lay::SMEM, st>>>(pp, tm);
    if (int e = check_cuda(cudaGetLastError(), "selscan_bwd3 launch")) return e;
    return bwd2_finish(pp, dtype, st);
}

int selscan_bwd3_launch(const BwdParams &p, int dtype, void *ws, cudaStream_t st) {
    Bwd2Params pp;
    memset(&pp, 0, sizeof(pp));
    pp.b = p;
    const bool z = p.z != nullptr;
    switch (dtype) {
        case MMI_F32: return z ? launch_bwd3_t<float, true>(pp, dtype, ws, st) : launch_bwd3_t<float, false>(pp, dtype, ws, st);
        case MMI_BF16:
            return z ? launch_bwd3_t<__nv_bfloat16, true>(pp, dtype, ws, st) : launch_bwd3_t<__nv_bfloat16, false>(pp, dtype, ws, st);
        case MMI_F16: return z ? launch_bwd3_t<__half, true>(pp, dtype, ws, st) : launch_bwd3_t<__half, false>(pp, dtype, ws, st);
    }
    set_error("selscan_bwd3: unknown dtype %d", dtype);
    return MMI_ERR_ARG;
}

}  // namespace mmi

// selscan_fwd2.cu -- fused selective scan forward, second generation (sm_100a).
//
// Same mathematics as selscan_fwd.cu (MambaBlock.selective_scan + gate, models/mamba.py:212-233, :184-186):
//     a[t,n] = exp(delta[t,d] A[d,n]);  h[t,n] = a[t,n] h[t-1,n] + delta[t,d] B[t,n] x[t,d];  y = sum_n C h + D x;  out = y silu(z)
// on the decomposition of selscan2.cuh: chains of 32 channels, 8 chunk-warps per CTA (128-step super-tiles through a
// 3-stage TMA ring), lanes split the 16 states in two halves and pack the two channels of a pair into one register pair,
// a persistent grid takes (chain, L segment) items in dependency order, the state crossing a segment boundary travels
// through global memory.
// Per super-tile and warp:
//   sweep A   (warps 0..6) chunk summary by direct evaluation, walking t backwards with S = sum of delta after t:
//             E[n] = sum_t exp(A[n] S_t) delta_t x_t B[t,n];  P[n] = exp(A[n] S_chunk)            (no dependence on h)
//   barrier, fold   h entering chunk wt = carry chained through the summaries of chunks 0..wt-1 (one FMA per state each)
//   sweep B   the scan proper: checkpoint (state entering the chunk, for the backward pass), decay, recurrence, C.h; every
//             four steps the two state halves exchange their partial y in one batch and each lane finishes two of the four
//             steps: D skip, SiLU gate, store into the output tile (which overwrites the z tile and leaves by one TMA store).
//             Warp 7's state after its last step is the carry of the next super-tile (its chunk needs no summary).
#include <algorithm>
#include <cstring>

#include "../../include/mmidet_b200.h"
#include "selscan.h"
#include "selscan2.cuh"

namespace mmi {

using namespace v2;

struct Fwd2Maps {
    CUtensorMap x, d, z, B, C, o;
};

constexpr int kFB2 = 4;  // steps per exchange block of sweep B
static_assert(kTC == kChunk, "a chunk starts at a checkpoint");

template <typename T, int NW> struct Fwd2Layout {
    static constexpr int N = kN, STAGES = 3, kNW = NW, kST = NW * kTC;  // (shadow the 8-warp constants of selscan2.cuh)
    static constexpr size_t TILE_BYTES = size_t(kST) * kCH * sizeof(T);
    static constexpr size_t BCT_BYTES = size_t(kST) * N * sizeof(T);
    static constexpr size_t STAGE_BYTES = 3 * TILE_BYTES + 2 * BCT_BYTES;  // x | delta | z (-> out) | B | C
    static constexpr size_t BC32_OFF = STAGES * STAGE_BYTES;               // per-warp widened B | C rows when T is 16 bit
    static constexpr size_t BC32_BYTES = sizeof(T) == 2 ? size_t(kNW) * 2 * kTC * N * 4 : 0;
    static constexpr size_t SUM_OFF = BC32_OFF + BC32_BYTES;               // chunk summaries [kNW - 1][8][32] float4 (E | P)
    static constexpr size_t SUM_WARP = size_t(8) * 32 * 16;
    static constexpr size_t CARRY_OFF = SUM_OFF + (kNW - 1) * SUM_WARP;    // carried state [2][4][32] float4
    static constexpr size_t BAR_OFF = CARRY_OFF + size_t(2) * 4 * 32 * 16;
    static constexpr size_t SMEM = BAR_OFF + STAGES * sizeof(uint64_t) + 16;
    static_assert(SMEM <= 232448, "shared memory budget of one CTA per SM");
};

template <typename T, int NW, bool GEOM, bool HAS_Z>
__device__ __forceinline__ void fwd2_item(const Fwd2Params &pp, const Fwd2Maps &tm, unsigned char *smem, const float2 (&A2p)[8],
                                          float2 A2b, float2 Dd, int c0, int b, int seg, int chain, int wt, int lane, bool active,
                                          int &g) {
    using Lay = Fwd2Layout<T, NW>;
    constexpr int kNW = NW, kST = NW * kTC;
    constexpr int N = kN, TC = kTC, ST = kST, CH = kCH;
    const FwdParams &p = pp.f;
    const SegSched &sc = pp.s;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    float4 *sums = reinterpret_cast<float4 *>(smem + Lay::SUM_OFF);            // [v * 8 + q][32], v = 0 .. kNW - 2
    float4 *carry = reinterpret_cast<float4 *>(smem + Lay::CARRY_OFF) + lane;  // + (buf * 4 + q) * 32
    float *bc32 = reinterpret_cast<float *>(smem + Lay::BC32_OFF) + wt * 2 * TC * N;
    const int pr = lane & 15, hs = lane >> 4;
    const bool up = hs != 0;

    const int L = p.L, ED = p.ED;
    const int ntiles_all = (L + ST - 1) / ST, nchk = (L + kChunk - 1) / kChunk;
    const int tile_lo = seg * sc.seg_tiles, ntiles = min(sc.seg_tiles, ntiles_all - tile_lo);
    const int tb = wt * TC, cl = 2 * pr, c = c0 + cl;
    const bool softplus = (p.flags & MMI_FLAG_DELTA_SOFTPLUS) != 0;

    auto issue = [&](int s, int ti) {  // one elected thread: the tiles of super-tile ti arrive on full[s]
        unsigned char *st = smem + size_t(s) * Lay::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], uint32_t(Lay::TILE_BYTES) * (HAS_Z ? 3u : 2u) + 2u * uint32_t(Lay::BCT_BYTES));
        tma_load_3d(st, &tm.x, c0, ti * ST, b, &full[s]);
        tma_load_3d(st + Lay::TILE_BYTES, &tm.d, c0, ti * ST, b, &full[s]);
        if (HAS_Z) tma_load_3d(st + 2 * Lay::TILE_BYTES, &tm.z, c0, ti * ST, b, &full[s]);
        tma_load_3d(st + 3 * Lay::TILE_BYTES, &tm.B, 0, ti * ST, b, &full[s]);
        tma_load_3d(st + 3 * Lay::TILE_BYTES + Lay::BCT_BYTES, &tm.C, 0, ti * ST, b, &full[s]);
    };
    if (threadIdx.x == 0) {
        bulk_wait_read<0>();  // output tiles of the previous item have left shared memory
        for (int i = 0; i < Lay::STAGES && i < ntiles; ++i) issue((g + i) % Lay::STAGES, tile_lo + i);
    }

    // state entering the segment: h0 (reference: zeros, models/mamba.py:252) for the first segment of L, else what the
    // earlier segment of this chain left behind
    if (wt == kNW - 1) {
        float2 hin[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hin[k] = make_float2(0.f, 0.f);
        if (seg > 0) {
            if (lane == 0) {
                const long long tw = clock64();
                while (ld_acquire(sc.done + chain) < unsigned(seg)) {
                    __nanosleep(64);
                    if (clock64() - tw > 20000000000LL) __trap();  // ~10 s: a lost predecessor traps instead of hanging
                }
            }
            __syncwarp();
            const float2 *gc = reinterpret_cast<const float2 *>(sc.carry) + int64_t(chain) * 8 * 32 + lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) hin[k] = __ldcg(gc + k * 32);
        } else if (p.h0 && active) {
            const float *h0 = p.h0 + (int64_t(b) * ED + c) * N + 8 * hs;
#pragma unroll
            for (int k = 0; k < 8; ++k) hin[k] = make_float2(h0[k], h0[N + k]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            carry[((g & 1) * 4 + q) * 32] = make_float4(hin[2 * q].x, hin[2 * q].y, hin[2 * q + 1].x, hin[2 * q + 1].y);
    }

    float2 h[8];  // after the last super-tile, in warp kNW - 1: the state leaving the segment
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = make_float2(0.f, 0.f);

    for (int it = 0; it < ntiles; ++it, ++g) {
        const int s = g % Lay::STAGES, t0 = (tile_lo + it) * ST;
        unsigned char *st = smem + size_t(s) * Lay::STAGE_BYTES;
        T *sx = reinterpret_cast<T *>(st) + tb * CH + cl, *sd = sx + ST * CH, *sz = sd + ST * CH;
        T *so = sz;
        mbar_wait(&full[s], (g / Lay::STAGES) & 1);

        const float *fB, *fC;  // this warp's 16 rows of B / C in fp32, offset to the lane's state half
        if constexpr (sizeof(T) == 2) {
            const T *gB = reinterpret_cast<const T *>(st + 3 * Lay::TILE_BYTES) + tb * N;
            const T *gC = reinterpret_cast<const T *>(st + 3 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N;
            for (int i = lane; i < TC * N; i += 32) {
                bc32[i] = to_f32<T>(gB[i]);
                bc32[TC * N + i] = to_f32<T>(gC[i]);
            }
            fB = bc32 + 8 * hs;
            fC = bc32 + TC * N + 8 * hs;
        } else {
            fB = reinterpret_cast<const float *>(st + 3 * Lay::TILE_BYTES) + tb * N + 8 * hs;
            fC = reinterpret_cast<const float *>(st + 3 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N + 8 * hs;
        }
        if (softplus) {  // fused softplus(dt_proj(.)), models/mamba.py:203: the lane activates its pair at steps 2 i + hs,
#pragma unroll       // in place, rounded to the I/O type as the unfused path does; rows past L stay 0 (identity steps)
            for (int i = 0; i < TC / 2; ++i) {
                const int u = 2 * i + hs;
                const float2 r = ld2<T>(sd + u * CH);
                const bool in = t0 + tb + u < L;
                st2<T>(sd + u * CH, make_float2(in ? softplus_fast(r.x) : 0.f, in ? softplus_fast(r.y) : 0.f));
            }
        }
        if (softplus || sizeof(T) == 2) __syncwarp();

        // ---- sweep A: chunk summary by direct evaluation (chunk kNW - 1 needs none) ------------------------------------
        if (wt < kNW - 1) {
            float2 acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = make_float2(0.f, 0.f);
            float2 S = make_float2(0.f, 0.f);
#pragma unroll
            for (int u = TC - 1; u >= 0; --u) {
                const float2 xv = ld2<T>(sx + u * CH), dv = ld2<T>(sd + u * CH);
                float Bv[8];
                ld8(fB + u * N, Bv);
                float2 pw[8];
                decay8f<GEOM>(S, A2b, A2p, up, mul2(dv, xv), pw);  // exp(A S) delta x
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fma2(pw[k], splat2(Bv[k]), acc[k]);
                S = add2(S, dv);
            }
            float2 Pe[8];
            decay8<GEOM>(S, A2b, A2p, up, Pe);
            float4 *o = sums + wt * 8 * 32 + lane;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                o[q * 32] = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
                o[(4 + q) * 32] = make_float4(Pe[2 * q].x, Pe[2 * q].y, Pe[2 * q + 1].x, Pe[2 * q + 1].y);
            }
        }
        __syncthreads();  // summaries of this super-tile and the carry written at the end of the previous one are visible

        // the previous super-tile's output store has had a whole sweep to drain; its stage can be refilled
        if (threadIdx.x == 0 && it >= 1 && it - 1 + Lay::STAGES < ntiles) {
            bulk_wait_read<0>();
            issue((g - 1) % Lay::STAGES, tile_lo + it - 1 + Lay::STAGES);
        }

        // ---- fold: state entering this warp's chunk -------------------------------------------------------------------
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w = carry[((g & 1) * 4 + q) * 32];
            h[2 * q] = make_float2(w.x, w.y);
            h[2 * q + 1] = make_float2(w.z, w.w);
        }
#pragma unroll
        for (int v = 0; v < kNW - 1; ++v) {  // unrolled with a warp-uniform guard: the loads of all summaries go out at once
            if (v < wt) {
                const float4 *o = sums + v * 8 * 32 + lane;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 E = o[q * 32], Pe = o[(4 + q) * 32];
                    h[2 * q] = fma2(make_float2(Pe.x, Pe.y), h[2 * q], make_float2(E.x, E.y));
                    h[2 * q + 1] = fma2(make_float2(Pe.z, Pe.w), h[2 * q + 1], make_float2(E.z, E.w));
                }
            }
        }
        if (p.chk && active && t0 + tb < L) {  // checkpoint = state entering step t0 + tb (kTC == kChunk): 128-bit stores
            float4 *ck = reinterpret_cast<float4 *>(p.chk + ((int64_t(b) * nchk + (t0 + tb) / kChunk) * ED + c) * N + 8 * hs);
            __stcs(ck, make_float4(h[0].x, h[1].x, h[2].x, h[3].x));
            __stcs(ck + 1, make_float4(h[4].x, h[5].x, h[6].x, h[7].x));
            __stcs(ck + N / 4, make_float4(h[0].y, h[1].y, h[2].y, h[3].y));
            __stcs(ck + N / 4 + 1, make_float4(h[4].y, h[5].y, h[6].y, h[7].y));
        }

        // ---- sweep B: the scan proper -------------------------------------------------------------------------------------
#pragma unroll 1
        for (int ub = 0; ub < TC; ub += kFB2) {
            float2 yp[kFB2];  // partial y (this lane's 8 states) of the block's steps
#pragma unroll
            for (int uu = 0; uu < kFB2; ++uu) {
                const int u = ub + uu;
                const float2 xv = ld2<T>(sx + u * CH), dv = ld2<T>(sd + u * CH);
                float Bv[8], Cv[8];
                ld8(fB + u * N, Bv);
                ld8(fC + u * N, Cv);
                float2 a[8];
                decay8<GEOM>(dv, A2b, A2p, up, a);
                const float2 dx = mul2(dv, xv);
                float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    h[k] = fma2(a[k], h[k], mul2(dx, splat2(Bv[k])));
                    if (k & 1) yb = fma2(h[k], splat2(Cv[k]), yb);
                    else ya = fma2(h[k], splat2(Cv[k]), ya);
                }
                yp[uu] = add2(ya, yb);
            }
            // the lane finishes steps ub + 2 j + hs: it keeps its own partial of those and receives the other half's
#pragma unroll
            for (int j = 0; j < kFB2 / 2; ++j) {
                const float2 mine = hs ? yp[2 * j + 1] : yp[2 * j], theirs = hs ? yp[2 * j] : yp[2 * j + 1];
                const float ox = __shfl_xor_sync(0xffffffffu, theirs.x, 16), oy = __shfl_xor_sync(0xffffffffu, theirs.y, 16);
                const int u = ub + 2 * j + hs;
                const float2 xv = ld2<T>(sx + u * CH);
                float2 y = fma2(Dd, xv, make_float2(mine.x + ox, mine.y + oy));
                if constexpr (HAS_Z) {
                    const float2 zv = ld2<T>(sz + u * CH);
                    y = mul2(y, make_float2(zv.x * sigmoidf_fast(zv.x), zv.y * sigmoidf_fast(zv.y)));
                }
                st2<T>(so + u * CH, y);
            }
        }
        if (wt == kNW - 1) {  // carry for the next super-tile
#pragma unroll
            for (int q = 0; q < 4; ++q)
                carry[(((g + 1) & 1) * 4 + q) * 32] = make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y);
        }
        fence_proxy_async();  // make the generic-proxy writes of the output tile visible to the TMA engine
        __syncthreads();      // every warp is done with stage s
        if (threadIdx.x == 0) {
            tma_store_3d(&tm.o, c0, t0, b, st + 2 * Lay::TILE_BYTES);  // rows past L / columns past ED are clipped
            bulk_commit();
        }
    }

    if (wt == kNW - 1) {
        if (seg < sc.nseg - 1) {  // hand the state to the next segment of this chain, then raise its flag
            float2 *gc = reinterpret_cast<float2 *>(sc.carry) + int64_t(chain) * 8 * 32 + lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) __stcg(gc + k * 32, h[k]);
            __threadfence();
            __syncwarp();
            if (lane == 0) red_release_add(sc.done + chain, 1u);
        } else if (p.hT && active) {  // steps past L are identities (zero-filled delta), so this is h[L-1]
            float *hT = p.hT + (int64_t(b) * ED + c) * N + 8 * hs;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                hT[k] = h[k].x;
                hT[N + k] = h[k].y;
            }
        }
    }
}

template <typename T, int NW, bool HAS_Z>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 1 : 2) selscan_fwd2_kernel(const Fwd2Params pp, const __grid_constant__ Fwd2Maps tm) {
    using Lay = Fwd2Layout<T, NW>;
    constexpr int N = kN;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    unsigned *ticket_s = reinterpret_cast<unsigned *>(smem + Lay::BAR_OFF + Lay::STAGES * sizeof(uint64_t));
    const FwdParams &p = pp.f;
    const SegSched &sc = pp.s;
    const int tid = threadIdx.x, wt = tid >> 5, lane = tid & 31, pr = lane & 15, hs = lane >> 4;

    if (tid == 0) {
        for (int s = 0; s < Lay::STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    int g = 0;  // super-tiles processed so far by this CTA: stage / mbarrier-parity / carry-buffer bookkeeping
    for (;;) {
        if (tid == 0) *ticket_s = atomicAdd(sc.ticket, 1u);
        __syncthreads();
        const int v = int(*ticket_s);
        if (v >= sc.nitems) break;
        // dependency order: every chain's FIRST segment of L, then every chain's second segment, ...
        const int seg = v / sc.nchains, chain = v % sc.nchains;
        const int b = chain / sc.ntile_c, ctile = chain % sc.ntile_c;
        const int c0 = ctile * kCH, c = c0 + 2 * pr;
        const bool active = c < p.ED;
        const int cc = active ? c : p.ED - 2;
        const float *Ar = p.A + int64_t(cc) * N;
        const float2 A2b = make_float2(Ar[0] * kLog2e, Ar[N] * kLog2e);
        const float2 Dd = make_float2(p.D[cc], p.D[cc + 1]);
        float2 A2p[8];
        bool ok = !(p.flags & MMI_FLAG_NO_GEOM);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int n = 8 * hs + k;
            A2p[k] = make_float2(Ar[n] * kLog2e, Ar[N + n] * kLog2e);
            const float w0 = float(n + 1) * A2b.x, w1 = float(n + 1) * A2b.y;
            ok = ok && (fabsf(A2p[k].x - w0) <= 2e-6f * fabsf(w0)) && (fabsf(A2p[k].y - w1) <= 2e-6f * fabsf(w1));
        }
        const bool geom = __syncthreads_and(ok);  // also: everyone has read the ticket before thread 0 takes the next one
        if (geom) fwd2_item<T, NW, true, HAS_Z>(pp, tm, smem, A2p, A2b, Dd, c0, b, seg, chain, wt, lane, active, g);
        else fwd2_item<T, NW, false, HAS_Z>(pp, tm, smem, A2p, A2b, Dd, c0, b, seg, chain, wt, lane, active, g);
    }
    if (tid == 0) bulk_wait_read<0>();  // shared memory must outlive the last tile store's reads
}

static size_t al256f(size_t v) { return (v + 255) & ~size_t(255); }

// workspace: [ticket | done (nchains)] [carry (nchains, 8, 32) float2]
int64_t selscan_fwd2_ws_bytes(int B, int L, int ED) {
    (void)L;
    const int64_t nch = int64_t(B) * ((ED + kCH - 1) / kCH);
    return int64_t(al256f(16 + size_t(nch) * 4)) + nch * 8 * 32 * 8;
}

template <typename T, int NW, bool HAS_Z> static int launch_fwd2_t(Fwd2Params pp, int dtype, void *ws, cudaStream_t st) {
    using Lay = Fwd2Layout<T, NW>;
    constexpr int kNW = NW, kST = NW * kTC;
    FwdParams &p = pp.f;
    auto kern = selscan_fwd2_kernel<T, NW, HAS_Z>;
    static thread_local int attr_dev = -1;  // the opt-in is per device and sticky: set it once, not on every launch
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Lay::SMEM)),
                               "selscan_fwd2 smem attribute"))
            return e;
        attr_dev = dev;
    }
    if (!ws) {
        set_error("selscan_fwd2: a workspace of mmi_selscan_fwd_ws_bytes() bytes is required");
        return MMI_ERR_ARG;
    }
    if (int e = seg_sched_plan(p.B, p.L, p.ED, p.flags, &pp.s)) return e;
    SegSched &sc = pp.s;
    if (NW != 8) {  // two CTAs per SM: more resident CTAs than chains, so chaining segments cannot add parallelism; re-plan
        const int ntiles = (p.L + kST - 1) / kST;  // on this variant's super-tile with the forced segment count only
        int nseg = (p.flags & MMI_FLAG_NSEG_MASK) >> MMI_FLAG_NSEG_SHIFT;
        nseg = std::max(1, std::min({nseg ? nseg : 1, 16, ntiles}));
        sc.seg_tiles = (ntiles + nseg - 1) / nseg;
        sc.nseg = (ntiles + sc.seg_tiles - 1) / sc.seg_tiles;
        sc.nitems = sc.nchains * sc.nseg;
    }
    sc.ticket = static_cast<unsigned *>(ws);
    sc.done = sc.ticket + 4;
    const size_t hdr = al256f(16 + size_t(sc.nchains) * 4);
    sc.carry = reinterpret_cast<float *>(static_cast<char *>(ws) + hdr);
    Fwd2Maps tm;
    memset(&tm, 0, sizeof(tm));
    const uint64_t nb = p.B, L = p.L;
    if (int e = make_tmap_3d(&tm.x, p.x, dtype, nb, L, p.ED, p.x_ld * sizeof(T), kST, kCH)) return e;
    if (int e = make_tmap_3d(&tm.d, p.delta, dtype, nb, L, p.ED, p.d_ld * sizeof(T), kST, kCH)) return e;
    if (HAS_Z)
        if (int e = make_tmap_3d(&tm.z, p.z, dtype, nb, L, p.ED, p.z_ld * sizeof(T), kST, kCH)) return e;
    if (int e = make_tmap_3d(&tm.B, p.Bm, dtype, nb, L, kN, kN * sizeof(T), kST, kN)) return e;
    if (int e = make_tmap_3d(&tm.C, p.Cm, dtype, nb, L, kN, kN * sizeof(T), kST, kN)) return e;
    if (int e = make_tmap_3d(&tm.o, p.out, dtype, nb, L, p.ED, p.o_ld * sizeof(T), kST, kCH)) return e;
    if (int e = check_cuda(cudaMemsetAsync(ws, 0, hdr, st), "selscan_fwd2 ticket memset")) return e;
    const int grid = std::min(sc.nitems, sm_count() * (NW == 8 ? 1 : 2));
    kern<<<grid, kNW * 32, Lay::SMEM, st>>>(pp, tm);
    return check_cuda(cudaGetLastError(), "selscan_fwd2 launch");
}

int selscan_fwd2_launch(const FwdParams &p, int dtype, void *ws, cudaStream_t st) {
    Fwd2Params pp;
    memset(&pp, 0, sizeof(pp));
    pp.f = p;
    const bool z = p.z != nullptr;
    const bool four = ((p.flags & MMI_FLAG_CFG_MASK) >> MMI_FLAG_CFG_SHIFT) == 10;  // 4 chunk-warps, two CTAs per SM
#define MMI_FWD2(T)                                                                                           \
    return four ? (z ? launch_fwd2_t<T, 4, true>(pp, dtype, ws, st) : launch_fwd2_t<T, 4, false>(pp, dtype, ws, st)) \
                : (z ? launch_fwd2_t<T, 8, true>(pp, dtype, ws, st) : launch_fwd2_t<T, 8, false>(pp, dtype, ws, st))
    switch (dtype) {
        case MMI_F32: MMI_FWD2(float);
        case MMI_BF16: MMI_FWD2(__nv_bfloat16);
        case MMI_F16: MMI_FWD2(__half);
    }
#undef MMI_FWD2
    set_error("selscan_fwd2: unknown dtype %d", dtype);
    return MMI_ERR_ARG;
}

}  // namespace mmi

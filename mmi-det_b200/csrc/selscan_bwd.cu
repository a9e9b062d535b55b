// selscan_bwd.cu -- fused selective scan backward for sm_100a (reverse scan with state recompute).
//
// Replaces autograd through MambaBlock.selective_scan + gate (models/mamba.py:222-231, :184-186) and PScan.backward
// (models/pscan.py:189-224).  With a = exp(delta*A), u = delta*B*x, dy = dout*silu(z):
//     g[t,n]  = C[t,n] dy[t] + a[t+1,n] g[t+1,n]            (pscan.py:216-219: A shifted left, reverse scan)
//     gradX   = g,  gradA[t] = h[t-1] g[t]                  (pscan.py:221-224)
//     ddelta  = sum_n (h[t-1] g a A + g B x);  dx = delta sum_n g B + D dy
//     dB[t,n] = sum_d g delta x;  dC[t,n] = sum_d dy h[t,n];  dA[d,n] = sum_{b,t} h[t-1] g a delta;  dD = sum dy x
//     dz      = dout y sigma(z) (1 + z (1 - sigma(z)))
// The reference keeps five padded (B,Lp,ED,N) tensors alive for this; here the forward pass leaves one state
// checkpoint per kChunk steps and each chunk's states are recomputed on chip.
//
// Mapping.  A CTA owns 64 adjacent channels of one batch element and walks L backwards in super-tiles of
// ST = WT*kChunk steps; warp wt owns chunk wt of the super-tile and each lane owns TWO adjacent channels with all
// N = 16 states of both in registers (the B / C rows, which every lane needs, are then fetched once per two channels,
// and the cross-channel sums for dB / dC start with an in-thread add).  Tiles of x / delta / dout / z / B / C arrive
// through a TMA ring; dx / ddelta / dz are written in place over the x / delta / z tiles and leave through TMA stores.
// Per super-tile, per warp:
//   A   dy = dout*silu(z) and the dz factor e = dout*sigma(z)(1 + z(1-sigma(z))) (parked over the dout / z tiles), and the
//       chunk's reverse-scan summary G[n] = sum_t exp(A[n]*(delta summed up to and including t)) dy[t] C[t,n] by direct
//       evaluation -- the time axis is scanned by WT warps at once, exactly as in the forward kernel
//   fold  after one CTA barrier each warp chains the summaries of the LATER chunks onto the carried a*g
//   P1  forward recompute of h from the chunk's checkpoint; every state is parked in TENSOR MEMORY (tcgen05.st, one
//       16-column slot per channel and step -- thread-private, so it costs no shared-memory bandwidth); y -> dz; the
//       per-step products dy*h are pre-added over the lane's two channels and summed over the warp's 64 channels by
//       a shared-memory transposition -> dC partial
//   P3  reverse scan: h[t-1] comes back from tensor memory (tcgen05.ld), g recurrence, ddelta / dx / dA / dD, and the
//       products g*delta*x are reduced the same way -> dB partial.
// Per-CTA dB/dC partials and per-batch dA/dD partials go to a workspace; selscan_bwd_finish_kernel reduces them
// deterministically (no atomics).
#include <algorithm>
#include <cstring>

#include "../../include/mmidet_b200.h"
#include "common.cuh"
#include "selscan.h"

namespace mmi {

struct BwdMaps {
    CUtensorMap x, d, z, g, B, C, odx, odd, odz;
};

constexpr int kRB = 4;  // steps per cross-channel reduction block

template <typename T, int WT, int STAGES> struct BwdLayout {
    static constexpr int N = kN, TC = kChunk, ST = WT * TC, CH = 64, NW = WT;
    static constexpr size_t TILE_BYTES = size_t(ST) * CH * sizeof(T);
    static constexpr size_t BCT_BYTES = size_t(ST) * N * sizeof(T);
    static constexpr size_t STAGE_BYTES = 4 * TILE_BYTES + 2 * BCT_BYTES;  // x | delta | dout | z | B | C
    static constexpr size_t DYE_OFF = STAGES * STAGE_BYTES;                // fp32 dy | e when T is 16 bit
    static constexpr size_t DYE_BYTES = sizeof(T) == 2 ? size_t(2) * ST * CH * 4 : 0;
    static constexpr size_t SCR_OFF = DYE_OFF + DYE_BYTES;                   // per warp [kRB][4][32] float4
    static constexpr size_t SCR_WARP = size_t(kRB) * 4 * 32 * 16;
    static constexpr size_t DA_OFF = SCR_OFF + NW * SCR_WARP;                // parked dA: [8][NW*32] float4
    static constexpr size_t SUM_OFF = DA_OFF + size_t(8) * NW * 32 * 16;     // chunk summaries [WT][2][4][32] float4
    static constexpr size_t SUMD_OFF = SUM_OFF + size_t(WT) * 2 * 4 * 32 * 16;  // chunk sum(delta) [WT][32] float2
    static constexpr size_t CARRY_OFF = SUMD_OFF + size_t(WT) * 32 * 8;      // carried a*g [2][2][4][32] float4
    static constexpr size_t BC32_OFF = CARRY_OFF + size_t(2) * 2 * 4 * 32 * 16;
    static constexpr size_t BC32_BYTES = sizeof(T) == 2 ? size_t(NW) * 2 * TC * N * 4 : 0;
    static constexpr size_t BAR_OFF = BC32_OFF + BC32_BYTES;
    static constexpr size_t SMEM = BAR_OFF + STAGES * sizeof(uint64_t) + 16;
};

// ---- tensor memory as thread-private scratch: 16 fp32 per thread per op (lane = TMEM lane, 16 columns) -----------
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float2 (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y), "f"(v[4].x),
        "f"(v[4].y), "f"(v[5].x), "f"(v[5].y), "f"(v[6].x), "f"(v[6].y), "f"(v[7].x), "f"(v[7].y)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float2 (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
          "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <bool GEOM> __device__ __forceinline__ void bdecay16(float dsum, float A2base, const float2 (&A2p)[8], float2 (&a2)[8]) {
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

__device__ __forceinline__ void bload16(const float *p, float2 (&v)[8]) {
    const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 w = q[k];
        v[2 * k] = make_float2(w.x, w.y);
        v[2 * k + 1] = make_float2(w.z, w.w);
    }
}

// two adjacent elements of a tile row (the lane's two channels)
template <typename T> __device__ __forceinline__ float2 ld_pair(const T *p);
template <> __device__ __forceinline__ float2 ld_pair<float>(const float *p) { return *reinterpret_cast<const float2 *>(p); }
template <> __device__ __forceinline__ float2 ld_pair<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p));
}
template <> __device__ __forceinline__ float2 ld_pair<__half>(const __half *p) {
    return __half22float2(*reinterpret_cast<const __half2 *>(p));
}
template <typename T> __device__ __forceinline__ void st_pair(T *p, float2 v);
template <> __device__ __forceinline__ void st_pair<float>(float *p, float2 v) { *reinterpret_cast<float2 *>(p) = v; }
template <> __device__ __forceinline__ void st_pair<__nv_bfloat16>(__nv_bfloat16 *p, float2 v) {
    *reinterpret_cast<__nv_bfloat162 *>(p) = __float22bfloat162_rn(v);
}
template <> __device__ __forceinline__ void st_pair<__half>(__half *p, float2 v) {
    *reinterpret_cast<__half2 *>(p) = __float22half2_rn(v);
}
__device__ __forceinline__ float pick(float2 v, int j) { return j ? v.y : v.x; }

template <typename T, int WT, int STAGES, bool SPLIT, bool GEOM, bool HAS_Z>
__device__ __forceinline__ void bwd_body(const BwdParams &p, const BwdMaps &tm, unsigned char *smem, uint32_t tmem_base,
                                         const float2 (&A2p)[2][8], const float (&A2base)[2], const float (&Dd)[2], int c0,
                                         int b, int seg, int ctile, int wt, int lane, int c, const bool (&active)[2]) {
    using Lay = BwdLayout<T, WT, STAGES>;
    constexpr int N = kN, TC = Lay::TC, ST = Lay::ST, CH = Lay::CH, NW = Lay::NW;
    constexpr float kLn2 = 0.6931471805599453f;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    float4 *scr = reinterpret_cast<float4 *>(smem + Lay::SCR_OFF + wt * Lay::SCR_WARP);           // [kRB][4][32]
    float4 *dApark = reinterpret_cast<float4 *>(smem + Lay::DA_OFF) + wt * 32 + lane;             // + k * NW*32
    float4 *sumG = reinterpret_cast<float4 *>(smem + Lay::SUM_OFF);                                // [WT][2][4][32]
    float2 *sumD = reinterpret_cast<float2 *>(smem + Lay::SUMD_OFF);                               // [WT][32]
    float4 *carry = reinterpret_cast<float4 *>(smem + Lay::CARRY_OFF) + lane;                      // [2][2][4][32]
    float *bc32 = reinterpret_cast<float *>(smem + Lay::BC32_OFF) + wt * 2 * TC * N;
    const uint32_t tslot = tmem_base + (uint32_t(wt & 3) * 32u << 16) + uint32_t(wt >> 2) * (2 * TC * N);

    const int L = p.L, ED = p.ED;
    const int ntiles_all = (L + ST - 1) / ST, nchk = (L + TC - 1) / TC;
    const int tile_lo = seg * p.seg_tiles, ntiles = min(p.seg_tiles, ntiles_all - tile_lo);  // this CTA's segment of L
    const int tb = wt * TC, cl = 2 * lane;
    const int64_t row_b = int64_t(b) * L;

    // one elected thread: the tile loads of super-tile tj arrive on full[s] (the summary pass needs delta, dout, z, C only)
    auto issue = [&](int s, int tj, bool full_pass) {
        unsigned char *st = smem + size_t(s) * Lay::STAGE_BYTES;
        const uint32_t total = uint32_t(Lay::TILE_BYTES) * ((HAS_Z ? 3u : 2u) + (full_pass ? 1u : 0u)) +
                               (full_pass ? 2u : 1u) * uint32_t(Lay::BCT_BYTES);
        mbar_arrive_expect_tx(&full[s], total);
        if (full_pass) tma_load_3d(st, &tm.x, c0, tj * ST, b, &full[s]);
        tma_load_3d(st + Lay::TILE_BYTES, &tm.d, c0, tj * ST, b, &full[s]);
        tma_load_3d(st + 2 * Lay::TILE_BYTES, &tm.g, c0, tj * ST, b, &full[s]);
        if (HAS_Z) tma_load_3d(st + 3 * Lay::TILE_BYTES, &tm.z, c0, tj * ST, b, &full[s]);
        if (full_pass) tma_load_3d(st + 4 * Lay::TILE_BYTES, &tm.B, 0, tj * ST, b, &full[s]);
        tma_load_3d(st + 4 * Lay::TILE_BYTES + Lay::BCT_BYTES, &tm.C, 0, tj * ST, b, &full[s]);
    };

    // zero the parked dA
#pragma unroll
    for (int k = 0; k < 8; ++k) dApark[k * NW * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    float dDacc[2] = {0.f, 0.f};
    float sdseg[2] = {0.f, 0.f};  // sum of delta over the segment so far (summary pass, warp 0)
    float2 glast[2][8];           // a*g leaving the segment so far (summary pass, warp 0)
    int g = 0;                    // super-tiles processed over both passes: stage / parity / carry-buffer bookkeeping

    // sum over the warp's 64 channels of the per-lane vectors parked in `scr` ([kRB][4][32] float4, one float4 per
    // (step, state quad, lane)); lane l sums 16 source lanes of item l & 15, halves combined by one shuffle.
    // which = 0: dB, 1: dC.  Partials go to ws_bc[(row * ntile_c + blockIdx.x) * 32 + which * 16 + n].
    auto reduce_block = [&](int tblk, int which) {
        __syncwarp();
        const int item = lane & 15, half = lane >> 4;
        const float4 *src = scr + item * 32 + half * 16;
        // four independent accumulator chains: the sum is latency-bound, not throughput-bound
        float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f), u0 = make_float2(0.f, 0.f), u1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const float4 v = src[(i + lane) & 15], w = src[(i + 1 + lane) & 15];
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
        const int t = tblk + (item >> 2), k4 = item & 3;
        if (half == 0 && t < L)
            __stcs(reinterpret_cast<float4 *>(p.ws_bc + ((row_b + t) * p.ntile_c + ctile) * (2 * N) + which * N + 4 * k4),
                   make_float4(s0.x, s0.y, s1.x, s1.y));
        __syncwarp();
    };

    // When L is split over several CTAs (SPLIT), every CTA but the one owning the earliest segment first reduces its
    // segment to (a*g leaving it when nothing enters, sum of delta) -- pass 0: sweep A + fold only -- and publishes it;
    // pass 1 waits for the summaries of the LATER segments (decoupled look-back through global memory), chains them and
    // then runs the segment for real.
    for (int pass = (SPLIT && seg > 0) ? 0 : 1; pass < 2; ++pass) {
    const bool full_pass = !SPLIT || pass == 1;
    if (SPLIT && full_pass && seg < p.nseg - 1) {
        if (threadIdx.x == 0) {
            for (int sp = p.nseg - 1; sp > seg; --sp) {
                const volatile unsigned *f = p.seg_flags + (int64_t(b) * p.nseg + sp) * p.ntile_c + ctile;
                const long long tw = clock64();
                while (*f == 0u) {
                    if (clock64() - tw > 20000000000LL) __trap();  // ~10 s: a lost predecessor traps instead of hanging
                }
            }
            __threadfence();
        }
        __syncthreads();
    }
    if (wt == 0) {  // a*g entering the segment: zero, or the later segments' summaries chained from the end of L
        float2 gin[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) gin[j][k] = make_float2(0.f, 0.f);
        if (SPLIT && full_pass) {
            for (int sp = p.nseg - 1; sp > seg; --sp) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (!active[j]) continue;
                    const float *sw = p.seg_ws + ((int64_t(b) * p.nseg + sp) * ED + c + j) * (N + 1);
                    float2 a2[8];
                    bdecay16<GEOM>(__ldcg(sw + N), A2base[j], A2p[j], a2);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        gin[j][k] = fma2(a2[k], gin[j][k], make_float2(__ldcg(sw + 2 * k), __ldcg(sw + 2 * k + 1)));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                carry[(((g & 1) * 2 + j) * 4 + k) * 32] =
                    make_float4(gin[j][2 * k].x, gin[j][2 * k].y, gin[j][2 * k + 1].x, gin[j][2 * k + 1].y);
    }
    if (threadIdx.x == 0)
        for (int i = 0; i < STAGES && i < ntiles; ++i) issue((g + i) % STAGES, tile_lo + ntiles - 1 - i, full_pass);

    for (int it = 0; it < ntiles; ++it, ++g) {
        const int s = g % STAGES, tj = tile_lo + ntiles - 1 - it, t0 = tj * ST;
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
        mbar_wait(&full[s], (g / STAGES) & 1);

        const float *fB, *fC;
        if constexpr (sizeof(T) == 2) {
            const T *gB = reinterpret_cast<const T *>(st + 4 * Lay::TILE_BYTES) + tb * N;
            const T *gC = reinterpret_cast<const T *>(st + 4 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N;
            for (int i = lane; i < TC * N; i += 32) {
                if (full_pass) bc32[i] = to_f32<T>(gB[i]);
                bc32[TC * N + i] = to_f32<T>(gC[i]);
            }
            __syncwarp();
            fB = bc32;
            fC = bc32 + TC * N;
        } else {
            fB = reinterpret_cast<const float *>(st + 4 * Lay::TILE_BYTES) + tb * N;
            fC = reinterpret_cast<const float *>(st + 4 * Lay::TILE_BYTES + Lay::BCT_BYTES) + tb * N;
        }

        // chunk checkpoint (state entering step t0 + tb): issued now, consumed in P1
        float4 ckv[2][4];
        {
            const bool inb = full_pass && t0 + tb < L;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float4 *ck = reinterpret_cast<const float4 *>(
                    p.chk + ((int64_t(b) * nchk + (inb ? (t0 + tb) / TC : 0)) * ED + (active[j] ? c + j : 0)) * N);
#pragma unroll
                for (int k = 0; k < 4; ++k) ckv[j][k] = (inb && active[j]) ? __ldcs(ck + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }

        if (p.flags & MMI_FLAG_DELTA_SOFTPLUS) {  // fused softplus(dt_proj(.)), models/mamba.py:203: activate this chunk's delta
#pragma unroll                              // once, in place (sweep A, P1 and P3 all read the activated value)
            for (int u = 0; u < TC; ++u) {
                const float2 r = ld_pair<T>(sd + u * CH);
                const bool in = t0 + tb + u < L;  // rows past L stay 0 (TMA zero fill), as in the forward kernel
                st_pair<T>(sd + u * CH, make_float2(in ? softplus_fast(r.x) : 0.f, in ? softplus_fast(r.y) : 0.f));
            }
        }
        // ---- A: dy, dz factor, reverse-scan summary of the chunk ------------------------------------------------
        float2 acc[2][8];
        float S[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[j][k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < TC; ++u) {
            const float2 dv = ld_pair<T>(sd + u * CH), gv = ld_pair<T>(sg + u * CH);
            float2 dy = gv, ee = make_float2(0.f, 0.f);
            if constexpr (HAS_Z) {
                const float2 zv = ld_pair<T>(sz + u * CH);
                const float sx_ = sigmoidf_fast(zv.x), sy_ = sigmoidf_fast(zv.y);
                const float gsx = gv.x * sx_, gsy = gv.y * sy_;
                dy = make_float2(gsx * zv.x, gsy * zv.y);
                ee = make_float2(gsx * fmaf(zv.x, 1.f - sx_, 1.f), gsy * fmaf(zv.y, 1.f - sy_, 1.f));
                *reinterpret_cast<float2 *>(se + u * CH) = ee;
            }
            if constexpr (HAS_Z || sizeof(T) == 2) *reinterpret_cast<float2 *>(sdy + u * CH) = dy;
            float2 Cv[8];
            bload16(fC + u * N, Cv);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                S[j] += pick(dv, j);
                float2 pw[8];
                const float dyj = pick(dy, j);
                if constexpr (GEOM) {
                    const float r = ex2(S[j] * A2base[j]), r2 = r * r, r4 = r2 * r2, r8 = r4 * r4, p0 = dyj * r;
                    pw[0] = make_float2(p0, p0 * r);
                    pw[1] = mul2(pw[0], splat2(r2));
                    pw[2] = mul2(pw[0], splat2(r4));
                    pw[3] = mul2(pw[1], splat2(r4));
#pragma unroll
                    for (int k = 4; k < 8; ++k) pw[k] = mul2(pw[k - 4], splat2(r8));
                } else {
                    const float2 S2 = splat2(S[j]), dy2 = splat2(dyj);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float2 e = mul2(S2, A2p[j][k]);
                        pw[k] = mul2(dy2, make_float2(ex2(e.x), ex2(e.y)));
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[j][k] = fma2(pw[k], Cv[k], acc[j][k]);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                sumG[((wt * 2 + j) * 4 + k) * 32 + lane] =
                    make_float4(acc[j][2 * k].x, acc[j][2 * k].y, acc[j][2 * k + 1].x, acc[j][2 * k + 1].y);
        sumD[wt * 32 + lane] = make_float2(S[0], S[1]);
        __syncthreads();  // summaries of this super-tile (and the carry written during the previous one) are visible

        // the previous super-tile's output stores have had a whole sweep to drain; its stage can be refilled
        if (threadIdx.x == 0 && it >= 1 && it - 1 + STAGES < ntiles) {
            bulk_wait_read<0>();
            issue((g - 1) % STAGES, tile_lo + ntiles - 1 - (it - 1 + STAGES), full_pass);
        }

        // ---- fold: a*g entering this chunk from the later ones --------------------------------------------------
        float2 ga[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float4 *cin = carry + ((g & 1) * 2 + j) * 4 * 32;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 w = cin[k * 32];
                ga[j][2 * k] = make_float2(w.x, w.y);
                ga[j][2 * k + 1] = make_float2(w.z, w.w);
            }
        }
#pragma unroll
        for (int v = WT - 1; v >= 1; --v) {
            if (v > wt) {
                const float2 sdv = sumD[v * 32 + lane];
                if (SPLIT && !full_pass) {
                    sdseg[0] += sdv.x;
                    sdseg[1] += sdv.y;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float2 a2[8];
                    bdecay16<GEOM>(pick(sdv, j), A2base[j], A2p[j], a2);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 w = sumG[((v * 2 + j) * 4 + k) * 32 + lane];
                        ga[j][2 * k] = fma2(a2[2 * k], ga[j][2 * k], make_float2(w.x, w.y));
                        ga[j][2 * k + 1] = fma2(a2[2 * k + 1], ga[j][2 * k + 1], make_float2(w.z, w.w));
                    }
                }
            }
        }
        if (wt == 0) {  // carry for the next (earlier) super-tile = this chunk's entry value pushed through its summary
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float2 a2[8];
                bdecay16<GEOM>(S[j], A2base[j], A2p[j], a2);
                float4 *cout = carry + (((g + 1) & 1) * 2 + j) * 4 * 32;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 v0 = fma2(a2[2 * k], ga[j][2 * k], acc[j][2 * k]);
                    const float2 v1 = fma2(a2[2 * k + 1], ga[j][2 * k + 1], acc[j][2 * k + 1]);
                    cout[k * 32] = make_float4(v0.x, v0.y, v1.x, v1.y);
                    if (SPLIT && !full_pass) {
                        glast[j][2 * k] = v0;
                        glast[j][2 * k + 1] = v1;
                    }
                }
                if (SPLIT && !full_pass) sdseg[j] += S[j];
            }
        }
        if (SPLIT && !full_pass) {  // summary pass: nothing else to do for this super-tile
            __syncthreads();        // summaries and stage s are free again
            continue;
        }

        // ---- P1: recompute the chunk's states; park them in tensor memory; y -> dz; dC -------------------------
        float2 h[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                h[j][2 * k] = make_float2(ckv[j][k].x, ckv[j][k].y);
                h[j][2 * k + 1] = make_float2(ckv[j][k].z, ckv[j][k].w);
            }
#pragma unroll 1
        for (int ub = 0; ub < TC; ub += kRB) {
#pragma unroll
            for (int uu = 0; uu < kRB; ++uu) {
                const int u = ub + uu;
                const float2 xv = ld_pair<T>(sx + u * CH), dv = ld_pair<T>(sd + u * CH);
                const float2 dy = *reinterpret_cast<const float2 *>(sdy + u * CH);
                float2 Bv[8], Cv[8], pc[8];
                bload16(fB + u * N, Bv);
                bload16(fC + u * N, Cv);
                float yy[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    tmem_st16(tslot + uint32_t((u * 2 + j) * N), h[j]);  // slot u = state ENTERING step u
                    float2 a2[8];
                    bdecay16<GEOM>(pick(dv, j), A2base[j], A2p[j], a2);
                    const float2 dx2 = splat2(pick(dv, j) * pick(xv, j)), dy2 = splat2(pick(dy, j));
                    float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        h[j][k] = fma2(a2[k], h[j][k], mul2(dx2, Bv[k]));
                        if (k & 1) yb = fma2(Cv[k], h[j][k], yb);
                        else ya = fma2(Cv[k], h[j][k], ya);
                        pc[k] = j ? fma2(dy2, h[j][k], pc[k]) : mul2(dy2, h[j][k]);
                    }
                    ya = add2(ya, yb);
                    yy[j] = fmaf(Dd[j], pick(xv, j), ya.x + ya.y);
                }
                if constexpr (HAS_Z) {
                    const float2 ee = *reinterpret_cast<const float2 *>(se + u * CH);
                    st_pair<T>(sz + u * CH, make_float2(yy[0] * ee.x, yy[1] * ee.y));  // dz, in place over z
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    scr[(uu * 4 + k) * 32 + lane] = make_float4(pc[2 * k].x, pc[2 * k].y, pc[2 * k + 1].x, pc[2 * k + 1].y);
            }
            reduce_block(t0 + tb + ub, 1);
        }
        tmem_wait_st();

        // ---- P3: reverse scan ---------------------------------------------------------------------------------
        float2 dA2[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 w = dApark[(j * 4 + k) * NW * 32];
                dA2[j][2 * k] = make_float2(w.x, w.y);
                dA2[j][2 * k + 1] = make_float2(w.z, w.w);
            }
#pragma unroll 1
        for (int ub = TC - kRB; ub >= 0; ub -= kRB) {
#pragma unroll
            for (int uu = kRB - 1; uu >= 0; --uu) {
                const int u = ub + uu;
                float2 hp[2][8];
                tmem_ld16(tslot + uint32_t((u * 2) * N), hp[0]);
                tmem_ld16(tslot + uint32_t((u * 2 + 1) * N), hp[1]);
                const float2 xv = ld_pair<T>(sx + u * CH), dv = ld_pair<T>(sd + u * CH);
                const float2 dy = *reinterpret_cast<const float2 *>(sdy + u * CH);
                float2 Bv[8], Cv[8], pb[8];
                bload16(fB + u * N, Bv);
                bload16(fC + u * N, Cv);
                tmem_wait_ld();
                float odx[2], odd[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float2 a2[8];
                    const float dvj = pick(dv, j), xvj = pick(xv, j), dyj = pick(dy, j);
                    bdecay16<GEOM>(dvj, A2base[j], A2p[j], a2);
                    const float2 dy2 = splat2(dyj), dv2 = splat2(dvj), dxw = splat2(dvj * xvj);
                    float2 dda = make_float2(0.f, 0.f), ddb = make_float2(0.f, 0.f);
                    float2 gBa = make_float2(0.f, 0.f), gBb = make_float2(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float2 g = fma2(Cv[k], dy2, ga[j][k]);  // g[t] = C dy + a[t+1] g[t+1]
                        const float2 ag = mul2(a2[k], g);             // a[t] g[t]   (carried to step t-1)
                        const float2 w = mul2(hp[j][k], ag);          // h[t-1] a g
                        float2 Aw;
                        if constexpr (GEOM) Aw = make_float2(float(2 * k + 1), float(2 * k + 2));
                        else Aw = A2p[j][k];
                        if (k & 1) {
                            ddb = fma2(w, Aw, ddb);
                            gBb = fma2(g, Bv[k], gBb);
                        } else {
                            dda = fma2(w, Aw, dda);
                            gBa = fma2(g, Bv[k], gBa);
                        }
                        dA2[j][k] = fma2(w, dv2, dA2[j][k]);
                        pb[k] = j ? fma2(dxw, g, pb[k]) : mul2(dxw, g);
                        ga[j][k] = ag;
                    }
                    dda = add2(dda, ddb);
                    gBa = add2(gBa, gBb);
                    float dd = (dda.x + dda.y) * kLn2;
                    if constexpr (GEOM) dd *= A2base[j];
                    const float gB = gBa.x + gBa.y;
                    odx[j] = fmaf(gB, dvj, Dd[j] * dyj);
                    odd[j] = fmaf(gB, xvj, dd);
                    if (p.flags & MMI_FLAG_DELTA_SOFTPLUS) odd[j] *= softplus_grad_from_value(dvj);  // d/d(pre-activation)
                    dDacc[j] = fmaf(dyj, xvj, dDacc[j]);
                }
                st_pair<T>(sx + u * CH, make_float2(odx[0], odx[1]));  // dx, in place over x
                st_pair<T>(sd + u * CH, make_float2(odd[0], odd[1]));  // ddelta, in place over delta
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    scr[(uu * 4 + k) * 32 + lane] = make_float4(pb[2 * k].x, pb[2 * k].y, pb[2 * k + 1].x, pb[2 * k + 1].y);
            }
            reduce_block(t0 + tb + ub, 0);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                dApark[(j * 4 + k) * NW * 32] = make_float4(dA2[j][2 * k].x, dA2[j][2 * k].y, dA2[j][2 * k + 1].x, dA2[j][2 * k + 1].y);

        fence_proxy_async();  // generic-proxy writes of the in-place output tiles -> visible to the TMA engine
        __syncthreads();      // every warp is done with stage s
        if (threadIdx.x == 0) {
            tma_store_3d(&tm.odx, c0, t0, b, st);
            tma_store_3d(&tm.odd, c0, t0, b, st + Lay::TILE_BYTES);
            if (HAS_Z) tma_store_3d(&tm.odz, c0, t0, b, st + 3 * Lay::TILE_BYTES);
            bulk_commit();
        }
    }
    if (SPLIT && !full_pass && wt == 0) {  // publish the segment summary, then raise the flag the earlier segments spin on
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!active[j]) continue;
            float *sw = p.seg_ws + ((int64_t(b) * p.nseg + seg) * ED + c + j) * (N + 1);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                __stcg(sw + 2 * k, glast[j][k].x);
                __stcg(sw + 2 * k + 1, glast[j][k].y);
            }
            __stcg(sw + N, sdseg[j]);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(p.seg_flags + (int64_t(b) * p.nseg + seg) * p.ntile_c + ctile, 1u);
    }
    }  // pass
    if (threadIdx.x == 0) bulk_wait_read<0>();
    __syncthreads();

    // per-(batch, segment, time-warp) partials of dA (A2 is A*log2e: dA = sum w delta, no rescale needed) and dD
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (active[j]) {
            float *o = p.ws_ad + (((int64_t(b) * p.nseg + seg) * NW + wt) * ED + c + j) * (N + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 w = dApark[(j * 4 + k) * NW * 32];
                o[4 * k] = w.x;
                o[4 * k + 1] = w.y;
                o[4 * k + 2] = w.z;
                o[4 * k + 3] = w.w;
            }
            o[N] = dDacc[j];
        }
    }
}

template <typename T, int WT, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(WT * 32, 1) selscan_bwd_kernel(const BwdParams p, const __grid_constant__ BwdMaps tm) {
    using Lay = BwdLayout<T, WT, STAGES>;
    constexpr int N = kN, CH = Lay::CH;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Lay::BAR_OFF + STAGES * sizeof(uint64_t));

    const int tid = threadIdx.x, wt = tid >> 5, lane = tid & 31;
    int b = blockIdx.y, ctile = blockIdx.x, seg = 0;
    if constexpr (SPLIT) {  // L split over CTAs: tickets hand out the LATER segments first (they are the predecessors)
        __shared__ unsigned ticket;
        if (tid == 0) ticket = atomicAdd(p.seg_ticket, 1u);
        __syncthreads();
        const int v = int(ticket);
        ctile = v % p.ntile_c;
        b = (v / p.ntile_c) % p.B;
        seg = p.nseg - 1 - v / (p.ntile_c * p.B);
    }
    const int c0 = ctile * CH;
    const int c = c0 + 2 * lane;
    const bool active[2] = {c < p.ED, c + 1 < p.ED};

    if (wt == 0) {  // all 512 tensor-memory columns: the state history of this CTA's 64 channels x WT chunks
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }

    float2 A2p[2][8];
    float A2base[2], Dd[2];
    bool ok = !(p.flags & MMI_FLAG_NO_GEOM);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int cc = active[j] ? c + j : p.ED - 1;
        A2base[j] = p.A[int64_t(cc) * N] * kLog2e;
        Dd[j] = p.D[cc];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            A2p[j][k] = make_float2(p.A[int64_t(cc) * N + 2 * k] * kLog2e, p.A[int64_t(cc) * N + 2 * k + 1] * kLog2e);
            const float w0 = float(2 * k + 1) * A2base[j], w1 = float(2 * k + 2) * A2base[j];
            ok = ok && (fabsf(A2p[j][k].x - w0) <= 2e-6f * fabsf(w0)) && (fabsf(A2p[j][k].y - w1) <= 2e-6f * fabsf(w1));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    const bool geom = __syncthreads_and(ok);  // also publishes the mbarrier inits and the tensor-memory base address
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const bool has_z = p.z != nullptr;
#define MMI_BWD_BODY(G, Z) \
    bwd_body<T, WT, STAGES, SPLIT, G, Z>(p, tm, smem, tmem_base, A2p, A2base, Dd, c0, b, seg, ctile, wt, lane, c, active)
    if (geom) {
        if (has_z) MMI_BWD_BODY(true, true);
        else MMI_BWD_BODY(true, false);
    } else {
        if (has_z) MMI_BWD_BODY(false, true);
        else MMI_BWD_BODY(false, false);
    }
#undef MMI_BWD_BODY
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wt == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// Deterministic reduction of the workspace partials: dB/dC over channel tiles, dA/dD over (batch, time-warp).
template <typename T>
__global__ void selscan_bwd_finish_kernel(const float *__restrict__ ws_bc, const float *__restrict__ ws_ad, T *dBm, T *dCm,
                                          float *dA, float *dD, int64_t rows, int ntile, int nparts, int ED) {
    constexpr int N = kN;
    const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t n_bc = rows * 2 * N;
    if (gid < n_bc) {
        const int64_t row = gid / (2 * N);
        const int r = int(gid % (2 * N));
        const float *src = ws_bc + row * ntile * (2 * N) + r;
        float v = 0.f;
        for (int tI = 0; tI < ntile; ++tI) v += src[int64_t(tI) * 2 * N];
        if (r < N) dBm[row * N + r] = from_f32<T>(v);
        else dCm[row * N + (r - N)] = from_f32<T>(v);
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

constexpr int kBwdWT = 4;  // time warps per CTA: their state history (2 channels x 16 states x 16 steps per lane) fills tensor memory
constexpr int kBwdMaxSeg = 32;
static size_t bwd_seg_header_bytes(int B, int ntile_c) { return (size_t(16) + size_t(B) * kBwdMaxSeg * ntile_c * 4 + 255) & ~size_t(255); }
static size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

// workspace: [dB/dC partials (B, L, ntile_c, 2N)] [dA/dD partials (B, nseg, WT, ED, N+1)] [ticket | flags] [summaries]
// L is only split when the unsplit grid leaves at least half of the SMs idle (also applied to a forced count), so the
// per-segment workspaces are sized for kBwdMaxSeg segments only in that case
static int bwd_seg_cap(int B, int ntile_c) { return int64_t(B) * ntile_c * 2 <= sm_count() ? kBwdMaxSeg : 1; }
int64_t selscan_bwd1_ws_bytes(int B, int L, int ED) {
    const int64_t ntile = (ED + 63) / 64;
    const int cap = bwd_seg_cap(B, int(ntile));
    return int64_t(al256(size_t(B) * L * ntile * 2 * kN * 4)) + int64_t(al256(size_t(B) * cap * kBwdWT * ED * (kN + 1) * 4)) +
           int64_t(bwd_seg_header_bytes(B, int(ntile))) + int64_t(B) * cap * ED * (kN + 1) * 4;
}

template <typename T> static int launch_bwd_t(BwdParams p, int dtype, void *ws, cudaStream_t st) {
    constexpr int WT = kBwdWT, STAGES = 2;
    using Lay = BwdLayout<T, WT, STAGES>;
    auto kern = selscan_bwd_kernel<T, WT, STAGES, false>;
    auto kern_split = selscan_bwd_kernel<T, WT, STAGES, true>;
    static thread_local int attr_dev = -1;  // the opt-in is per device and sticky: set it once, not on every launch
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Lay::SMEM)),
                               "selscan_bwd smem attribute"))
            return e;
        if (int e = check_cuda(cudaFuncSetAttribute(kern_split, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Lay::SMEM)),
                               "selscan_bwd smem attribute"))
            return e;
        attr_dev = dev;
    }
    const uint64_t rows = uint64_t(p.B) * p.L, nb = p.B, L = p.L;
    p.ntile_c = (p.ED + Lay::CH - 1) / Lay::CH;
    const int ntiles = (p.L + Lay::ST - 1) / Lay::ST;
    // L split (see the forward launcher): only when B * (ED / 64) CTAs leave SMs idle; one CTA per SM here
    const int forced = (p.flags & MMI_FLAG_NSEG_MASK) >> MMI_FLAG_NSEG_SHIFT;
    const int ctas = p.ntile_c * p.B, slots = sm_count();
    int nseg = forced ? forced : (ctas * 2 <= slots ? slots / ctas : 1);
    const int cap = bwd_seg_cap(p.B, p.ntile_c);
    nseg = std::max(1, std::min({nseg, cap, forced ? ntiles : ntiles / 2}));
    p.seg_tiles = (ntiles + nseg - 1) / nseg;
    p.nseg = (ntiles + p.seg_tiles - 1) / p.seg_tiles;
    char *w = static_cast<char *>(ws);
    p.ws_bc = reinterpret_cast<float *>(w);
    w += al256(size_t(rows) * p.ntile_c * 2 * kN * 4);
    p.ws_ad = reinterpret_cast<float *>(w);
    w += al256(size_t(p.B) * cap * WT * p.ED * (kN + 1) * 4);
    p.seg_ticket = reinterpret_cast<unsigned *>(w);
    p.seg_flags = p.seg_ticket + 4;
    const size_t hdr = bwd_seg_header_bytes(p.B, p.ntile_c);
    p.seg_ws = reinterpret_cast<float *>(w + hdr);
    BwdMaps tm;
    memset(&tm, 0, sizeof(tm));
    if (int e = make_tmap_3d(&tm.x, p.x, dtype, nb, L, p.ED, p.x_ld * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (int e = make_tmap_3d(&tm.d, p.delta, dtype, nb, L, p.ED, p.d_ld * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (int e = make_tmap_3d(&tm.g, p.dout, dtype, nb, L, p.ED, p.g_ld * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (p.z)
        if (int e = make_tmap_3d(&tm.z, p.z, dtype, nb, L, p.ED, p.z_ld * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (int e = make_tmap_3d(&tm.B, p.Bm, dtype, nb, L, kN, kN * sizeof(T), Lay::ST, kN)) return e;
    if (int e = make_tmap_3d(&tm.C, p.Cm, dtype, nb, L, kN, kN * sizeof(T), Lay::ST, kN)) return e;
    if (int e = make_tmap_3d(&tm.odx, p.dx, dtype, nb, L, p.ED, p.ED * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (int e = make_tmap_3d(&tm.odd, p.ddelta, dtype, nb, L, p.ED, p.ED * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (p.dz)
        if (int e = make_tmap_3d(&tm.odz, p.dz, dtype, nb, L, p.ED, p.ED * sizeof(T), Lay::ST, Lay::CH)) return e;
    if (p.nseg > 1) {
        if (int e = check_cuda(cudaMemsetAsync(p.seg_ticket, 0, hdr, st), "selscan_bwd segment flags memset")) return e;
        kern_split<<<dim3(unsigned(p.ntile_c) * p.B * p.nseg), WT * 32, Lay::SMEM, st>>>(p, tm);
    } else {
        kern<<<dim3(p.ntile_c, p.B), WT * 32, Lay::SMEM, st>>>(p, tm);
    }
    if (int e = check_cuda(cudaGetLastError(), "selscan_bwd launch")) return e;
    const int64_t work = int64_t(rows) * 2 * kN + int64_t(p.ED) * (kN + 1);
    selscan_bwd_finish_kernel<T><<<unsigned((work + 255) / 256), 256, 0, st>>>(
        p.ws_bc, p.ws_ad, static_cast<T *>(p.dBm), static_cast<T *>(p.dCm), p.dA, p.dD, int64_t(rows), p.ntile_c,
        p.B * p.nseg * WT, p.ED);
    return check_cuda(cudaGetLastError(), "selscan_bwd finish launch");
}

int selscan_bwd1_launch(BwdParams p, int dtype, void *ws, cudaStream_t st) {
    switch (dtype) {
        case MMI_F32: return launch_bwd_t<float>(p, dtype, ws, st);
        case MMI_BF16: return launch_bwd_t<__nv_bfloat16>(p, dtype, ws, st);
        case MMI_F16: return launch_bwd_t<__half>(p, dtype, ws, st);
    }
    set_error("selscan_bwd: unknown dtype %d", dtype);
    return MMI_ERR_ARG;
}

}  // namespace mmi

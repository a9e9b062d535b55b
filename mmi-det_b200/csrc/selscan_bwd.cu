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
// checkpoint per kChunk steps and each chunk's states are recomputed into shared memory.
//
// Per chunk (processed last to first), each warp runs four phases on its own 32/LPC channels:
//   P1 (owner lanes, forward)  recompute h[t], park it in the shared history, compute dy
//   P2 (transposed mapping)    dC[t,:] += sum over the warp's channels of dy * h[t]      (reads the history)
//   P3 (owner lanes, reverse)  g recurrence, ddelta / dx / dz / dA / dD, overwrite history with g*delta*x
//   P4 (transposed mapping)    dB[t,:] += sum over the warp's channels of the overwritten history
// so the cross-channel reductions cost one conflict-free LDS.128 + FFMA2 pair per 4 states instead of a shuffle
// butterfly.  Per-CTA dB/dC partials and per-batch dA/dD partials go to a workspace; selscan_bwd_finish_kernel
// reduces them deterministically (no atomics).  Inputs arrive through the same TMA ring as the forward kernel.
#include <cstring>

#include "../../include/mmidet_b200.h"
#include "common.cuh"
#include "selscan.h"

namespace mmi {

struct BwdMaps {
    CUtensorMap x, d, z, g, B, C, odx, odd, odz;
};

template <typename T, int LPC, int NW, int TC, int STAGES> struct BwdLayout {
    static constexpr int N = kN, NS = N / LPC, K4 = NS / 4, CPW = 32 / LPC, CH = NW * CPW;
    static constexpr size_t TILE_BYTES = size_t(TC) * CH * sizeof(T);
    static constexpr size_t BCT_BYTES = size_t(TC) * N * sizeof(T);
    static constexpr size_t CHK_BYTES = size_t(CH) * N * 4;
    static constexpr size_t STAGE_BYTES = 4 * TILE_BYTES + 2 * BCT_BYTES + CHK_BYTES;
    static constexpr size_t HIST_OFF = STAGES * STAGE_BYTES;
    static constexpr size_t HIST_WARP_BYTES = size_t(TC + 1) * K4 * 512;
    static constexpr size_t DYS_OFF = HIST_OFF + NW * HIST_WARP_BYTES;
    static constexpr int DYS_LD = CPW + 1;
    static constexpr size_t DBC_OFF = DYS_OFF + size_t(NW) * TC * DYS_LD * 4;
    static constexpr size_t OUT_OFF = DBC_OFF + size_t(2) * NW * TC * N * 4;  // [2 buffers][dx, ddelta, dz] tiles
    static constexpr size_t BC32_OFF = OUT_OFF + 6 * TILE_BYTES;
    static constexpr size_t BAR_OFF = BC32_OFF + (sizeof(T) == 2 ? size_t(2) * TC * N * 4 : 0);
    static constexpr size_t SMEM = BAR_OFF + STAGES * sizeof(uint64_t);
};

template <typename T, int LPC, int NW, int TC, int STAGES, bool GEOM, bool HAS_Z>
__device__ __forceinline__ void bwd_body(const BwdParams &p, const BwdMaps &tm, unsigned char *smem,
                                         const float (&A2)[kN / LPC], float A2base, float Dd, int c0, int chw, int b,
                                         int cl, int c, bool active, int sub, int warp, int lane) {
    using Lay = BwdLayout<T, LPC, NW, TC, STAGES>;
    constexpr int N = kN, NS = Lay::NS, K4 = Lay::K4, CH = Lay::CH, NP = NS / 2;
    constexpr float kLn2 = 0.6931471805599453f;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);
    float4 *hist = reinterpret_cast<float4 *>(smem + Lay::HIST_OFF + warp * Lay::HIST_WARP_BYTES);  // [TC+1][K4][32]
    float *dys = reinterpret_cast<float *>(smem + Lay::DYS_OFF) + warp * TC * Lay::DYS_LD;          // [TC][CPW+1]
    float *dbc = reinterpret_cast<float *>(smem + Lay::DBC_OFF);                                    // [2][NW][TC][N]
    float *bc32 = reinterpret_cast<float *>(smem + Lay::BC32_OFF);

    const int L = p.L, ED = p.ED;
    constexpr bool has_z = HAS_Z;
    const int nch = (L + TC - 1) / TC;
    const int64_t row_b = int64_t(b) * L;
    const int chl = lane / LPC;  // channel within the warp
    T *gdx = static_cast<T *>(p.dx), *gdd = static_cast<T *>(p.ddelta), *gdz = static_cast<T *>(p.dz);

    auto stage_ptr = [&](int s) { return smem + size_t(s) * Lay::STAGE_BYTES; };
    auto issue = [&](int s, int j) {  // elected thread: 6 TMA tiles + the chunk's state checkpoint
        unsigned char *st = stage_ptr(s);
        const int row0 = int(row_b) + j * TC;
        const uint32_t total = uint32_t(Lay::TILE_BYTES) * (has_z ? 4u : 3u) + 2u * uint32_t(Lay::BCT_BYTES) +
                               uint32_t(chw) * N * 4u;
        mbar_arrive_expect_tx(&full[s], total);
        tma_load_2d(st, &tm.x, c0, row0, &full[s]);
        tma_load_2d(st + Lay::TILE_BYTES, &tm.d, c0, row0, &full[s]);
        tma_load_2d(st + 2 * Lay::TILE_BYTES, &tm.g, c0, row0, &full[s]);
        if (has_z) tma_load_2d(st + 3 * Lay::TILE_BYTES, &tm.z, c0, row0, &full[s]);
        tma_load_2d(st + 4 * Lay::TILE_BYTES, &tm.B, 0, row0, &full[s]);
        tma_load_2d(st + 4 * Lay::TILE_BYTES + Lay::BCT_BYTES, &tm.C, 0, row0, &full[s]);
        bulk_g2s(st + 4 * Lay::TILE_BYTES + 2 * Lay::BCT_BYTES, p.chk + ((int64_t(b) * nch + j) * ED + c0) * N,
                 uint32_t(chw) * N * 4u, &full[s]);
    };

    float2 A2p[NP], g2[NP], an2[NP], dA2[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        A2p[k] = make_float2(A2[2 * k], A2[2 * k + 1]);
        g2[k] = an2[k] = dA2[k] = make_float2(0.f, 0.f);
    }
    float dDacc = 0.f;

    auto decay = [&](float dv, float2(&a2)[NP]) {  // a[t,n] for this lane's states
        if constexpr (GEOM) {
            const float r = ex2(dv * A2base);
            const float q = (LPC == 1) ? r : ex2(dv * A2[0]);
            const float r2 = r * r;
            a2[0] = make_float2(q, q * r);
            if constexpr (NP >= 2) a2[1] = mul2(a2[0], splat2(r2));
            if constexpr (NP >= 4) {
                const float2 r4 = splat2(r2 * r2);
                a2[2] = mul2(a2[0], r4);
                a2[3] = mul2(a2[1], r4);
            }
            if constexpr (NP >= 8) {
                const float r4s = r2 * r2;
                const float2 r8 = splat2(r4s * r4s);
#pragma unroll
                for (int k = 4; k < 8; ++k) a2[k] = mul2(a2[k - 4], r8);
            }
        } else {
            const float2 dv2 = splat2(dv);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const float2 e = mul2(dv2, A2p[k]);
                a2[k] = make_float2(ex2(e.x), ex2(e.y));
            }
        }
    };

    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES && s < nch; ++s) issue(s, nch - 1 - s);

    for (int i = 0; i < nch; ++i) {
        const int s = i % STAGES, j = nch - 1 - i;
        const int t0 = j * TC, tl = min(TC, L - t0);
        unsigned char *st = stage_ptr(s);
        const T *sx = reinterpret_cast<const T *>(st), *sd = sx + TC * CH, *sg = sd + TC * CH, *sz = sg + TC * CH;
        const T *sB = reinterpret_cast<const T *>(st + 4 * Lay::TILE_BYTES), *sC = sB + TC * N;
        const float *sck = reinterpret_cast<const float *>(st + 4 * Lay::TILE_BYTES + 2 * Lay::BCT_BYTES);
        mbar_wait(&full[s], (i / STAGES) & 1);

        const float *fB, *fC;
        if constexpr (sizeof(T) == 2) {
            for (int q = threadIdx.x; q < 2 * TC * N; q += NW * 32) bc32[q] = to_f32<T>(sB[q]);
            __syncthreads();
            fB = bc32;
            fC = bc32 + TC * N;
        } else {
            fB = reinterpret_cast<const float *>(sB);
            fC = reinterpret_cast<const float *>(sC);
        }
        auto loadBC = [&](const float *base, int t, float2(&v)[NP]) {
            const float4 *q = reinterpret_cast<const float4 *>(base + t * N + sub * NS);
#pragma unroll
            for (int k = 0; k < K4; ++k) {
                const float4 w = q[k];
                v[2 * k] = make_float2(w.x, w.y);
                v[2 * k + 1] = make_float2(w.z, w.w);
            }
        };

        // ---- P1: recompute the chunk's states from its checkpoint -----------------------------------------
        float2 h2[NP];
        {
            const float4 *q = reinterpret_cast<const float4 *>(sck + cl * N + sub * NS);
#pragma unroll
            for (int k = 0; k < K4; ++k) {
                const float4 w = active ? q[k] : make_float4(0.f, 0.f, 0.f, 0.f);
                h2[2 * k] = make_float2(w.x, w.y);
                h2[2 * k + 1] = make_float2(w.z, w.w);
                hist[k * 32 + lane] = w;
            }
        }
        auto p1_step = [&](int t) {
            const float xv = to_f32<T>(sx[t * CH + cl]), dv = to_f32<T>(sd[t * CH + cl]);
            float2 Bv[NP], a2[NP];
            loadBC(fB, t, Bv);
            decay(dv, a2);
            const float2 dx2 = splat2(dv * xv);
#pragma unroll
            for (int k = 0; k < NP; ++k) h2[k] = fma2(a2[k], h2[k], mul2(dx2, Bv[k]));
#pragma unroll
            for (int k = 0; k < K4; ++k)
                hist[((t + 1) * K4 + k) * 32 + lane] = make_float4(h2[2 * k].x, h2[2 * k].y, h2[2 * k + 1].x, h2[2 * k + 1].y);
            float dy = to_f32<T>(sg[t * CH + cl]);
            if constexpr (HAS_Z) {
                const float zv = to_f32<T>(sz[t * CH + cl]);
                dy *= zv * sigmoidf_fast(zv);
            }
            dys[t * Lay::DYS_LD + chl] = dy;  // the LPC lanes of a channel write the same value
        };
        if (tl == TC) {
#pragma unroll
            for (int t = 0; t < TC; ++t) p1_step(t);
        } else {
            for (int t = 0; t < tl; ++t) p1_step(t);
        }
        __syncwarp();

        // ---- P2 / P4: sum the history over this warp's channels (lane l walks entries (i + l) & 31) ---------
        auto reduce_hist = [&](float *dst, bool weighted) {
#pragma unroll
            for (int jj = 0; jj < (TC * K4 + 31) / 32; ++jj) {
                const int id = lane + 32 * jj;
                const int t = id / K4, k = id % K4;
                if (id < TC * K4 && t < tl) {
                    float2 acc[LPC][2];
#pragma unroll
                    for (int q = 0; q < LPC; ++q) acc[q][0] = acc[q][1] = make_float2(0.f, 0.f);
                    const float4 *src = hist + ((t + 1) * K4 + k) * 32;
                    const float *dyr = dys + t * Lay::DYS_LD;
#pragma unroll 4
                    for (int i0 = 0; i0 < 32; i0 += LPC) {
#pragma unroll
                        for (int q = 0; q < LPC; ++q) {
                            const int e = (i0 + q + lane) & 31;
                            const float4 v = src[e];
                            const float2 m = splat2(weighted ? dyr[e / LPC] : 1.0f);
                            acc[q][0] = fma2(m, make_float2(v.x, v.y), acc[q][0]);
                            acc[q][1] = fma2(m, make_float2(v.z, v.w), acc[q][1]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < LPC; ++q) {
                        const int se = (q + lane) % LPC;  // which state slice entry e belonged to
                        *reinterpret_cast<float4 *>(dst + t * N + se * NS + 4 * k) =
                            make_float4(acc[q][0].x, acc[q][0].y, acc[q][1].x, acc[q][1].y);
                    }
                }
            }
        };
        reduce_hist(dbc + (NW + warp) * TC * N, true);  // dC partial of this warp
        __syncwarp();

        // ---- P3: reverse scan -----------------------------------------------------------------------------
        float2 hc2[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) hc2[k] = h2[k];
        T *so = reinterpret_cast<T *>(smem + Lay::OUT_OFF + (i & 1) * 3 * Lay::TILE_BYTES) + cl;  // dx | ddelta | dz tiles
        auto p3_step = [&](int t, bool to_smem) {
            const float xv = to_f32<T>(sx[t * CH + cl]), dv = to_f32<T>(sd[t * CH + cl]);
            const float gv = to_f32<T>(sg[t * CH + cl]);
            float2 Bv[NP], Cv[NP], a2[NP], hp2[NP];
            loadBC(fB, t, Bv);
            loadBC(fC, t, Cv);
            decay(dv, a2);
#pragma unroll
            for (int k = 0; k < K4; ++k) {
                const float4 w = hist[(t * K4 + k) * 32 + lane];
                hp2[2 * k] = make_float2(w.x, w.y);
                hp2[2 * k + 1] = make_float2(w.z, w.w);
            }
            float zv = 0.f, sig = 1.f, dy = gv;
            if constexpr (HAS_Z) {
                zv = to_f32<T>(sz[t * CH + cl]);
                sig = sigmoidf_fast(zv);
                dy = gv * zv * sig;
            }
            const float2 dy2 = splat2(dy), dv2 = splat2(dv), dx2 = splat2(dv * xv);
            float2 ya = make_float2(0.f, 0.f), dda = make_float2(0.f, 0.f), gBa = make_float2(0.f, 0.f);
            float2 gd[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                ya = fma2(Cv[k], hc2[k], ya);                         // y[t] readout (for dz)
                g2[k] = fma2(an2[k], g2[k], mul2(dy2, Cv[k]));        // g[t]
                const float2 tmp = mul2(mul2(hp2[k], g2[k]), a2[k]);  // h[t-1] g a
                dda = fma2(tmp, A2p[k], dda);
                dA2[k] = fma2(tmp, dv2, dA2[k]);
                gBa = fma2(g2[k], Bv[k], gBa);
                gd[k] = mul2(g2[k], dx2);
                an2[k] = a2[k];
                hc2[k] = hp2[k];
            }
#pragma unroll
            for (int k = 0; k < K4; ++k)
                hist[((t + 1) * K4 + k) * 32 + lane] = make_float4(gd[2 * k].x, gd[2 * k].y, gd[2 * k + 1].x, gd[2 * k + 1].y);
            float y = ya.x + ya.y, dd = (dda.x + dda.y) * kLn2, gB = gBa.x + gBa.y;
            if constexpr (LPC >= 2) {
                y += __shfl_xor_sync(0xffffffffu, y, 1);
                dd += __shfl_xor_sync(0xffffffffu, dd, 1);
                gB += __shfl_xor_sync(0xffffffffu, gB, 1);
            }
            if constexpr (LPC >= 4) {
                y += __shfl_xor_sync(0xffffffffu, y, 2);
                dd += __shfl_xor_sync(0xffffffffu, dd, 2);
                gB += __shfl_xor_sync(0xffffffffu, gB, 2);
            }
            y = fmaf(Dd, xv, y);
            dDacc = fmaf(dy, xv, dDacc);
            const T odx = from_f32<T>(fmaf(gB, dv, Dd * dy)), odd = from_f32<T>(fmaf(gB, xv, dd));
            const T odz = from_f32<T>(gv * y * sig * fmaf(zv, 1.f - sig, 1.f));
            if (to_smem) {  // the LPC lanes of a channel write identical values
                so[t * CH] = odx;
                so[TC * CH + t * CH] = odd;
                if constexpr (HAS_Z) so[2 * TC * CH + t * CH] = odz;
            } else if (active && sub == 0) {
                const int64_t o = (row_b + t0 + t) * ED + c;
                gdx[o] = odx;
                gdd[o] = odd;
                if constexpr (HAS_Z) gdz[o] = odz;
            }
        };
        if (tl == TC) {
#pragma unroll
            for (int t = TC - 1; t >= 0; --t) p3_step(t, true);
            fence_proxy_async();
        } else {
            for (int t = tl - 1; t >= 0; --t) p3_step(t, false);
        }
        __syncwarp();
        reduce_hist(dbc + warp * TC * N, false);  // dB partial of this warp
        if (threadIdx.x == 0) bulk_wait_read<0>();  // the previous chunk's tile stores have drained their buffers
        __syncthreads();

        // ---- combine the warps' partials, one 128-byte row [dB(16) | dC(16)] per timestep --------------------
        for (int q = threadIdx.x; q < tl * 2 * N; q += NW * 32) {
            const int t = q / (2 * N), r = q % (2 * N), which = r / N, n = r % N;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += dbc[((which * NW + w) * TC + t) * N + n];
            p.ws_bc[((row_b + t0 + t) * p.ntile_c + blockIdx.x) * (2 * N) + r] = v;
        }
        __syncthreads();  // stage s, history and dbc are free again
        if (threadIdx.x == 0) {
            if (tl == TC) {
                const T *ob = so - cl;
                tma_store_2d(&tm.odx, c0, int(row_b) + t0, ob);
                tma_store_2d(&tm.odd, c0, int(row_b) + t0, ob + TC * CH);
                if (HAS_Z) tma_store_2d(&tm.odz, c0, int(row_b) + t0, ob + 2 * TC * CH);
                bulk_commit();
            }
            if (i + STAGES < nch) issue(s, nch - 1 - (i + STAGES));
        }
    }
    if (threadIdx.x == 0) bulk_wait_read<0>();

    if (active) {  // per-batch partials of dA (pre-scaled A2 -> A handled above), dD
        float *o = p.ws_ad + (int64_t(b) * ED + c) * (N + 1);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            o[sub * NS + 2 * k] = dA2[k].x;
            o[sub * NS + 2 * k + 1] = dA2[k].y;
        }
        if (sub == 0) o[N] = dDacc;
    }
}

template <typename T, int LPC, int NW, int TC, int STAGES>
__global__ void __launch_bounds__(NW * 32) selscan_bwd_kernel(const BwdParams p, const __grid_constant__ BwdMaps tm) {
    using Lay = BwdLayout<T, LPC, NW, TC, STAGES>;
    constexpr int N = kN, NS = Lay::NS, CPW = Lay::CPW, CH = Lay::CH;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Lay::BAR_OFF);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, c0 = blockIdx.x * CH;
    const int chw = min(CH, p.ED - c0);
    const int sub = lane % LPC;
    const int cl = warp * CPW + lane / LPC;
    const int c = c0 + cl;
    const bool active = c < p.ED;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    const int cc = active ? c : p.ED - 1;
    float A2[NS];
    const float A2base = p.A[int64_t(cc) * N] * kLog2e;
    bool ok = !(p.flags & MMI_FLAG_NO_GEOM);
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        A2[k] = p.A[int64_t(cc) * N + sub * NS + k] * kLog2e;
        const float want = float(sub * NS + k + 1) * A2base;
        ok = ok && (fabsf(A2[k] - want) <= 2e-6f * fabsf(want));
    }
    const float Dd = p.D[cc];
    const bool geom = __syncthreads_and(ok);
    const bool has_z = p.z != nullptr;
#define MMI_BWD_BODY(G, Z) \
    bwd_body<T, LPC, NW, TC, STAGES, G, Z>(p, tm, smem, A2, A2base, Dd, c0, chw, b, cl, c, active, sub, warp, lane)
    if (geom) {
        if (has_z) MMI_BWD_BODY(true, true);
        else MMI_BWD_BODY(true, false);
    } else {
        if (has_z) MMI_BWD_BODY(false, true);
        else MMI_BWD_BODY(false, false);
    }
#undef MMI_BWD_BODY
}

// Deterministic reduction of the workspace partials: dB/dC over channel tiles, dA/dD over the batch.
template <typename T>
__global__ void selscan_bwd_finish_kernel(const float *__restrict__ ws_bc, const float *__restrict__ ws_ad, T *dBm, T *dCm,
                                          float *dA, float *dD, int64_t rows, int ntile, int B, int ED) {
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
        for (int bI = 0; bI < B; ++bI) v += ws_ad[int64_t(bI) * ED * (N + 1) + g2];
        const int c = int(g2 / (N + 1)), n = int(g2 % (N + 1));
        if (n < N) dA[int64_t(c) * N + n] = v;
        else dD[c] = v;
    }
}

static int pick_lpc(int B, int ED, int flags) {
    (void)B; (void)flags;
    return ED >= 64 ? 1 : ED >= 32 ? 2 : 4;
}

template <typename T, int LPC> static int launch_bwd_t(BwdParams p, int dtype, void *ws, cudaStream_t st) {
    constexpr int NW = 2, TC = kChunk, STAGES = 3;
    using Lay = BwdLayout<T, LPC, NW, TC, STAGES>;
    auto kern = selscan_bwd_kernel<T, LPC, NW, TC, STAGES>;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Lay::SMEM)),
                           "selscan_bwd smem attribute"))
        return e;
    const uint64_t rows = uint64_t(p.B) * p.L;
    p.ntile_c = (p.ED + Lay::CH - 1) / Lay::CH;
    p.ws_bc = static_cast<float *>(ws);
    p.ws_ad = p.ws_bc + rows * p.ntile_c * 2 * kN;
    BwdMaps tm;
    memset(&tm, 0, sizeof(tm));
    if (int e = make_tmap_2d(&tm.x, p.x, dtype, rows, p.ED, p.x_ld * sizeof(T), TC, Lay::CH)) return e;
    if (int e = make_tmap_2d(&tm.d, p.delta, dtype, rows, p.ED, p.d_ld * sizeof(T), TC, Lay::CH)) return e;
    if (int e = make_tmap_2d(&tm.g, p.dout, dtype, rows, p.ED, p.g_ld * sizeof(T), TC, Lay::CH)) return e;
    if (p.z)
        if (int e = make_tmap_2d(&tm.z, p.z, dtype, rows, p.ED, p.z_ld * sizeof(T), TC, Lay::CH)) return e;
    if (int e = make_tmap_2d(&tm.B, p.Bm, dtype, rows, kN, kN * sizeof(T), TC, kN)) return e;
    if (int e = make_tmap_2d(&tm.C, p.Cm, dtype, rows, kN, kN * sizeof(T), TC, kN)) return e;
    if (int e = make_tmap_2d(&tm.odx, p.dx, dtype, rows, p.ED, p.ED * sizeof(T), TC, Lay::CH)) return e;
    if (int e = make_tmap_2d(&tm.odd, p.ddelta, dtype, rows, p.ED, p.ED * sizeof(T), TC, Lay::CH)) return e;
    if (p.dz)
        if (int e = make_tmap_2d(&tm.odz, p.dz, dtype, rows, p.ED, p.ED * sizeof(T), TC, Lay::CH)) return e;
    dim3 grid(p.ntile_c, p.B);
    kern<<<grid, NW * 32, Lay::SMEM, st>>>(p, tm);
    if (int e = check_cuda(cudaGetLastError(), "selscan_bwd launch")) return e;
    const int64_t work = int64_t(rows) * 2 * kN + int64_t(p.ED) * (kN + 1);
    selscan_bwd_finish_kernel<T><<<unsigned((work + 255) / 256), 256, 0, st>>>(
        p.ws_bc, p.ws_ad, static_cast<T *>(p.dBm), static_cast<T *>(p.dCm), p.dA, p.dD, int64_t(rows), p.ntile_c, p.B, p.ED);
    return check_cuda(cudaGetLastError(), "selscan_bwd finish launch");
}

template <typename T> static int launch_bwd_lpc(const BwdParams &p, int dtype, int lpc, void *ws, cudaStream_t st) {
    switch (lpc) {
        case 1: return launch_bwd_t<T, 1>(p, dtype, ws, st);
        case 2: return launch_bwd_t<T, 2>(p, dtype, ws, st);
        case 4: return launch_bwd_t<T, 4>(p, dtype, ws, st);
    }
    set_error("selscan_bwd: lanes-per-channel must be 1, 2 or 4 (got %d)", lpc);
    return MMI_ERR_ARG;
}

// worst case over the LPC choices: smallest channel tile = NW*32/4 = 16 channels
int64_t selscan_bwd_ws_bytes(int B, int L, int ED) {
    const int64_t ntile_max = (ED + 15) / 16;
    return (int64_t(B) * L * ntile_max * 2 * kN + int64_t(B) * ED * (kN + 1)) * 4;
}

int selscan_bwd_launch(BwdParams p, int dtype, void *ws, cudaStream_t st) {
    const int lpc = pick_lpc(p.B, p.ED, p.flags);
    switch (dtype) {
        case MMI_F32: return launch_bwd_lpc<float>(p, dtype, lpc, ws, st);
        case MMI_BF16: return launch_bwd_lpc<__nv_bfloat16>(p, dtype, lpc, ws, st);
        case MMI_F16: return launch_bwd_lpc<__half>(p, dtype, lpc, ws, st);
    }
    set_error("selscan_bwd: unknown dtype %d", dtype);
    return MMI_ERR_ARG;
}

}  // namespace mmi

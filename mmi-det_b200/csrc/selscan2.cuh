// selscan2.cuh -- device helpers shared by the second-generation scan kernels (selscan_fwd2.cu, selscan_bwd2.cu).
//
// Work decomposition of both kernels:
//   chain     one batch element x 32 adjacent channels; B * ceil(ED / 32) independent chains (256 at the bench shape, more
//             than the 148 SMs -- the 64-channel tiles of the first generation gave only 128)
//   CTA       8 warps; warp wt owns chunk wt (16 steps) of a 128-step super-tile of the chain the CTA is working on
//   lane      pr = lane & 15 -> channel pair (c0 + 2 pr, c0 + 2 pr + 1);  hs = lane >> 4 -> states 8 hs .. 8 hs + 7.
//             A register pair (float2) holds ONE state of BOTH channels, so every per-(t, d) quantity (delta, x, dy, the
//             decay base, y, ddelta, dx ...) is a packed pair straight out of shared memory, and the per-(t, n) quantities
//             B[t, n], C[t, n], which are shared by all channels, enter FFMA2 as its 32-bit broadcast operand -- no
//             register shuffling between packed and scalar form anywhere in the step loop.
//   item      (chain, L segment).  A persistent grid of one CTA per SM takes items from a ticket counter in dependency
//             order (all chains' first segment, then all chains' second segment, ...); a segment starts from the carry its
//             predecessor left in global memory (flag = number of finished segments of the chain).  No summary pass is
//             needed for this: the predecessor's ticket is always older, so it is running or done.  With 4 segments the
//             1024 items of the bench shape fill 148 SMs to 98.8 % (256 whole chains: 86.5 %).
#pragma once
#include "common.cuh"

#include "selscan.h"

namespace mmi {

// host side shared by selscan_bwd2.cu and selscan_bwd3.cu (defined in selscan_bwd2.cu)
struct Bwd2Maps {
    CUtensorMap x, d, z, g, B, C, odx, odd, odz;
};
int bwd2_prepare(Bwd2Params &pp, Bwd2Maps &tm, int dtype, void *ws, cudaStream_t st);  // plan, workspace, tensor maps, ticket reset
int bwd2_finish(const Bwd2Params &pp, int dtype, cudaStream_t st);                     // deterministic reduction of the partials

namespace v2 {

constexpr int kCH = 32;    // channels per chain
constexpr int kNW = 8;     // chunk-warps per CTA
constexpr int kTC = 16;    // steps per chunk (= checkpoint interval, selscan.h kChunk)
constexpr int kST = kNW * kTC;  // steps per super-tile
constexpr int kHS = 8;     // states per lane

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
// 8 fp32 per thread per op (the 16-warp backward: 4 states x 2 channels per lane)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float2 (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "f"(v[0].x), "f"(v[0].y),
                 "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float2 (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- pairs of adjacent tile elements (the lane's two channels) --------------------------------------------------------
template <typename T> __device__ __forceinline__ float2 ld2(const T *p);
template <> __device__ __forceinline__ float2 ld2<float>(const float *p) { return *reinterpret_cast<const float2 *>(p); }
template <> __device__ __forceinline__ float2 ld2<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p));
}
template <> __device__ __forceinline__ float2 ld2<__half>(const __half *p) { return __half22float2(*reinterpret_cast<const __half2 *>(p)); }
template <typename T> __device__ __forceinline__ void st2(T *p, float2 v);
template <> __device__ __forceinline__ void st2<float>(float *p, float2 v) { *reinterpret_cast<float2 *>(p) = v; }
template <> __device__ __forceinline__ void st2<__nv_bfloat16>(__nv_bfloat16 *p, float2 v) {
    *reinterpret_cast<__nv_bfloat162 *>(p) = __float22bfloat162_rn(v);
}
template <> __device__ __forceinline__ void st2<__half>(__half *p, float2 v) { *reinterpret_cast<__half2 *>(p) = __float22half2_rn(v); }

// 8 consecutive floats of a B / C row (this lane's states) -- two LDS.128
__device__ __forceinline__ void ld8(const float *p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4 *>(p)[0], b = reinterpret_cast<const float4 *>(p)[1];
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}

// a[k] = exp(dv * A[., 8 hs + k]) for both channels.  GEOM: A[c, n] = (n + 1) A[c, 0] -> powers of R = exp(dv * A[c, 0]);
// `up` = (hs == 1) selects the factor R^8 of the upper half.  A2b / A2p are A * log2(e).
template <bool GEOM>
__device__ __forceinline__ void decay8(float2 dv, float2 A2b, const float2 (&A2p)[8], bool up, float2 (&a)[8]) {
    if constexpr (GEOM) {
        const float2 e = mul2(dv, A2b);
        const float2 R = make_float2(ex2(e.x), ex2(e.y));
        const float2 R2 = mul2(R, R), R4 = mul2(R2, R2), R8 = mul2(R4, R4);
        const float2 F = up ? R8 : make_float2(1.f, 1.f);
        a[0] = mul2(R, F);
        a[1] = mul2(R2, F);
        a[2] = mul2(a[0], R2);
        a[3] = mul2(a[1], R2);
        a[4] = mul2(a[0], R4);
        a[5] = mul2(a[1], R4);
        a[6] = mul2(a[2], R4);
        a[7] = mul2(a[3], R4);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 e = mul2(dv, A2p[k]);
            a[k] = make_float2(ex2(e.x), ex2(e.y));
        }
    }
}

// the same with every element multiplied by `fac` (one extra multiply in the geometric form: the factor rides on F)
template <bool GEOM>
__device__ __forceinline__ void decay8f(float2 dv, float2 A2b, const float2 (&A2p)[8], bool up, float2 fac, float2 (&a)[8]) {
    if constexpr (GEOM) {
        const float2 e = mul2(dv, A2b);
        const float2 R = make_float2(ex2(e.x), ex2(e.y));
        const float2 R2 = mul2(R, R), R4 = mul2(R2, R2), R8 = mul2(R4, R4);
        const float2 F = mul2(up ? R8 : make_float2(1.f, 1.f), fac);
        a[0] = mul2(R, F);
        a[1] = mul2(R2, F);
        a[2] = mul2(a[0], R2);
        a[3] = mul2(a[1], R2);
        a[4] = mul2(a[0], R4);
        a[5] = mul2(a[1], R4);
        a[6] = mul2(a[2], R4);
        a[7] = mul2(a[3], R4);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 e = mul2(dv, A2p[k]);
            a[k] = mul2(make_float2(ex2(e.x), ex2(e.y)), fac);
        }
    }
}

// 4 states per lane (state quarter q): a[k] = exp(dv * A[., 4 q + k]).  GEOM: a[0] = exp2(dv * A2q) with A2q = (4 q + 1) A2b
// evaluated directly, the other three by powers of R = exp2(dv * A2b) -- 4 exponentials + 5 packed multiplies per step.
template <bool GEOM>
__device__ __forceinline__ void decay4(float2 dv, float2 A2b, float2 A2q, const float2 (&A2p)[4], float2 (&a)[4]) {
    if constexpr (GEOM) {
        const float2 e = mul2(dv, A2b), e1 = mul2(dv, A2q);
        const float2 R = make_float2(ex2(e.x), ex2(e.y));
        a[0] = make_float2(ex2(e1.x), ex2(e1.y));
        const float2 R2 = mul2(R, R);
        a[1] = mul2(a[0], R);
        a[2] = mul2(a[0], R2);
        a[3] = mul2(a[1], R2);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 e = mul2(dv, A2p[k]);
            a[k] = make_float2(ex2(e.x), ex2(e.y));
        }
    }
}

// sum of a packed pair over the two state halves (lanes l and l ^ 16): both lanes get the total
__device__ __forceinline__ float2 xhalf_sum(float2 v) {
    const float ox = __shfl_xor_sync(0xffffffffu, v.x, 16), oy = __shfl_xor_sync(0xffffffffu, v.y, 16);
    return make_float2(v.x + ox, v.y + oy);
}

// acquire / release on a global flag
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace v2
}  // namespace mmi

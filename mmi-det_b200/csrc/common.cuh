// common.cuh -- device helpers shared by the sm_100a kernels (PTX wrappers for mbarrier / TMA bulk copies,
// packed f32x2 math, MUFU ex2, I/O type conversion).  Everything here is Blackwell-only by design.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmi {

constexpr float kLog2e = 1.4426950408889634f;
constexpr int kN = 16;  // d_state (models/mamba.py:35); kernels keep the N states of a channel in registers

// ---- error plumbing (capi.cu) --------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);

// ---- I/O element conversion ----------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// ---- math ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2(float x) {  // MUFU.EX2
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp(float x) {  // MUFU.RCP
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoidf_fast(float z) { return rcp(1.0f + ex2(-kLog2e * z)); }

__device__ __forceinline__ float lg2(float x) {  // MUFU.LG2
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softplus with torch's semantics (F.softplus, beta = 1, threshold = 20; models/mamba.py:203): x above the threshold passes
// through; log1p(e^x) is evaluated as a short series when e^x is small (lg2(1 + e) would lose it to rounding)
__device__ __forceinline__ float softplus_fast(float x) {
    const float e = ex2(kLog2e * x);
    const float big = 0.6931471805599453f * lg2(1.0f + e);
    const float small = e * fmaf(e, fmaf(e, 0.33333334f, -0.5f), 1.0f);
    return x > 20.0f ? x : (e < 0.0078125f ? small : big);
}
// d softplus(x) / dx = sigmoid(x), recovered from s = softplus(x):  1 - e^(-s)
__device__ __forceinline__ float softplus_grad_from_value(float s) { return 1.0f - ex2(-kLog2e * s); }

// packed fp32x2 (FFMA2 / FMUL2 / FADD2 on sm_100): two lanes of work per issue slot
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

// ---- shared-memory addressing, mbarrier, TMA bulk copy --------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost TMA completion traps (~2 s) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// TMA engine, 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA engine, 1-D bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// Shared-memory loads the compiler may hoist above earlier plain stores.  The scan kernels update their tiles in place
// (row t is read, then overwritten with an output); nvcc cannot prove that the store to row t does not alias the load of
// row t+1, so with plain loads every step starts with a load burst it then waits on.  `asm volatile` without a memory
// clobber keeps each load (no CSE across the in-place update, never deleted, ordered against the other volatile asm) but
// lets it move across plain stores; real producer -> consumer hand-offs between phases all sit behind __syncthreads /
// __syncwarp / tcgen05.wait, which are compiler barriers.
__device__ __forceinline__ float4 lds_f4(const void *p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ float2 lds_f2(const void *p) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ float lds_f1(const void *p) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(const void *p) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ unsigned short lds_u16(const void *p) {
    unsigned short v;
    asm volatile("ld.shared.b16 %0, [%1];" : "=h"(v) : "r"(smem_u32(p)));
    return v;
}
// one tile element / two adjacent tile elements as fp32
template <typename T> __device__ __forceinline__ float lds_elem(const T *p);
template <> __device__ __forceinline__ float lds_elem<float>(const float *p) { return lds_f1(p); }
template <> __device__ __forceinline__ float lds_elem<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __uint_as_float(uint32_t(lds_u16(p)) << 16);
}
template <> __device__ __forceinline__ float lds_elem<__half>(const __half *p) { return __half2float(__ushort_as_half(lds_u16(p))); }
template <typename T> __device__ __forceinline__ float2 lds_pair(const T *p);
template <> __device__ __forceinline__ float2 lds_pair<float>(const float *p) { return lds_f2(p); }
template <> __device__ __forceinline__ float2 lds_pair<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const uint32_t w = lds_u32(p);
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 lds_pair<__half>(const __half *p) {
    const uint32_t w = lds_u32(p);
    return __half22float2(*reinterpret_cast<const __half2 *>(&w));
}

// streaming global stores (outputs are never re-read by the kernel that writes them)
__device__ __forceinline__ void st_cs(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(__nv_bfloat16 *p, __nv_bfloat16 v) {
    __stcs(reinterpret_cast<unsigned short *>(p), __bfloat16_as_ushort(v));
}
__device__ __forceinline__ void st_cs(__half *p, __half v) {
    __stcs(reinterpret_cast<unsigned short *>(p), __half_as_ushort(v));
}

}  // namespace mmi

// ---- TMA tensor maps (2-D tiles) -------------------------------------------------------------------------
#include <cuda.h>

namespace mmi {

// Encode a row-major 2-D tensor [rows x cols] of `esize`-byte elements with a row pitch of `pitch_bytes`
// and a box of [box_rows x box_cols].  Out-of-bounds elements are zero-filled by the TMA engine.
// dtype: MMI_F32 / MMI_BF16 / MMI_F16.  Returns 0 or an MMI_ERR_* code (host side, tmap.cu).
int make_tmap_2d(CUtensorMap *map, const void *base, int dtype, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                 uint32_t box_rows, uint32_t box_cols);

int make_tmap_3d(CUtensorMap *map, const void *base, int dtype, uint64_t nb, uint64_t L, uint64_t cols, uint64_t pitch_bytes,
                 uint32_t box_rows, uint32_t box_cols);

// TMA engine, 3-D tile global -> shared; coordinates are {column, t, b} of a (B, L, cols) tensor.
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int col, int t, int b, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(col), "r"(t), "r"(b), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int col, int t, int b, const void *src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(col),
                 "r"(t), "r"(b), "r"(smem_u32(src))
                 : "memory");
}

// TMA engine, 2-D tile global -> shared (SASS: UTMALDG); coordinates are {column, row}.
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int col, int row, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(col), "r"(row), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace mmi

namespace mmi {
// TMA engine, 2-D tile shared -> global (SASS: UTMASTG); out-of-bounds parts of the box are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int col, int row, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col),
                 "r"(row), "r"(smem_u32(src))
                 : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
}  // namespace mmi

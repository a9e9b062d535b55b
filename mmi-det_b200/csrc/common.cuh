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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

}  // namespace mmi

// ---- TMA tensor maps (3-D tiles) -------------------------------------------------------------------------
#include <cuda.h>

namespace mmi {

// Encode a (B, L, cols) tensor of `dtype` elements (MMI_F32 / MMI_BF16 / MMI_F16) with a row pitch of `pitch_bytes` and a box of
// [box_rows x box_cols] within one batch element.  Out-of-bounds elements are zero-filled by the TMA engine on loads and
// clipped on stores.  Returns 0 or an MMI_ERR_* code (host side, tmap.cu; descriptors are memoised per thread).
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

}  // namespace mmi

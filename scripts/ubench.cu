// ubench.cu -- sm_100a issue-rate microbenchmarks used to size the selective-scan kernels (DESIGN.md "pipe budget").
// Every test: 1 CTA per SM, W warps, a long unrolled loop of independent ops; reports warp-instructions per
// clock per SM from clock64() of CTA 0.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench.bin ubench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x)                                                                  \
    do {                                                                       \
        cudaError_t e = (x);                                                   \
        if (e != cudaSuccess) {                                                \
            printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); \
            return 1;                                                          \
        }                                                                      \
    } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }


__device__ __forceinline__ float4 lds128(const void *p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ void sts128(void *p, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds32(const void *p) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

template <int MODE> __global__ void k_alu(float *out, long long *cyc, float s) {
    // MODE 0: FFMA (3-reg) x16 chains; 1: FFMA2 x8 chains; 2: FMUL2; 3: MUFU.EX2 x8; 4: FFMA2 + EX2 interleaved
    // 5: SHFL.BFLY x8; 6: FFMA2 x8 + SHFL x4 ; 7: FADD2
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
    const float a = s, b = s * 0.5f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = fma2(make_float2(r[2 * i], r[2 * i + 1]), make_float2(a, b), make_float2(b, a));
                r[2 * i] = v.x;
                r[2 * i + 1] = v.y;
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = __fmul2_rn(make_float2(r[2 * i], r[2 * i + 1]), make_float2(a, b));
                r[2 * i] = v.x;
                r[2 * i + 1] = v.y;
            }
        } else if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = __fadd2_rn(make_float2(r[2 * i], r[2 * i + 1]), make_float2(a, b));
                r[2 * i] = v.x;
                r[2 * i + 1] = v.y;
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = fma2(make_float2(r[2 * i], r[2 * i + 1]), make_float2(a, b), make_float2(b, a));
                r[2 * i] = v.x;
                r[2 * i + 1] = v.y;
                if (i < 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[2 * i]));
            }
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __shfl_xor_sync(0xffffffffu, r[i], 1 << (i % 5));
        } else if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = fma2(make_float2(r[2 * i], r[2 * i + 1]), make_float2(a, b), make_float2(b, a));
                r[2 * i] = v.x;
                r[2 * i + 1] = v.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) r[2 * i] = __shfl_xor_sync(0xffffffffu, r[2 * i], 1 << i);
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared memory.  MODE 0: LDS.128 lane-private (conflict-free, 512 B/instr); 1: LDS.128 rotated by lane (conflict-free);
// 2: STS.128; 3: LDS.32 conflict-free; 4: LDS.128 broadcast (all lanes one address); 5: LDS.64 broadcast; 6: LDS.32
// broadcast; 7: FFMA2 x8 + LDS.128 broadcast x4; 8: FFMA2 x8 + LDS.128 private x2 + STS.128 x2
template <int MODE> __global__ void k_smem(float *out, long long *cyc, float s) {
    extern __shared__ float4 sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 4 * nt; i += nt) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float acc = 0.f;
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = tid * 0.001f + i;
    const float a = s, b = 0.5f * s;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 7 || MODE == 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = fma2(make_float2(r[2 * i], r[2 * i + 1]), make_float2(a, b), make_float2(b, a));
                r[2 * i] = v.x;
                r[2 * i + 1] = v.y;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t addr;
            if (MODE == 0 || MODE == 2) addr = base + 16u * ((i & 3) * nt + tid);
            else if (MODE == 1) addr = base + 16u * ((i & 3) * nt + (tid & ~31) + ((tid + it + i) & 31));
            else if (MODE == 3) addr = base + 4u * ((i & 3) * nt + tid);
            else if (MODE == 8) addr = base + 16u * ((i & 1) * nt + tid);
            else addr = base + 16u * ((i & 3) * nt + (tid & ~31) + ((it + i) & 31));
            if (MODE == 0 || MODE == 1 || MODE == 4 || (MODE == 7 && i < 4) || (MODE == 8 && i < 2)) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
                acc += (v.x + v.y) + (v.z + v.w);
            } else if (MODE == 2 || (MODE == 8 && i >= 2 && i < 4)) {
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr + (MODE == 8 ? 32u * nt : 0u)), "f"(r[0]), "f"(r[1]),
                             "f"(r[2]), "f"(acc)
                             : "memory");
            } else if (MODE == 5) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
                acc += v.x + v.y;
            } else if (MODE == 3 || MODE == 6) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
                acc += v;
            }
        }
    }
    const long long t1 = clock64();
    float o = acc;
#pragma unroll
    for (int i = 0; i < 16; ++i) o += r[i];
    out[blockIdx.x * nt + tid] = o;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

// tensor memory as a thread-private spill space: tcgen05.st / tcgen05.ld 32x32b.x16 (16 fp32 per thread per op)
// MODE 0: st only; 1: ld only; 2: st+ld pairs; 3: FFMA2 x8 + st + ld per iter
template <int MODE> __global__ void k_tmem(float *out, long long *cyc, float s) {
    __shared__ uint32_t tbase_s;
    __shared__ float4 scr[4 * 512];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4 * 512; i += blockDim.x) scr[i] = make_float4(i, 1, 2, 3);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            (uint32_t)__cvta_generic_to_shared(&tbase_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tbase_s;
    // warp w may touch lanes 32*(w%4)..+31; warps sharing a lane quarter get disjoint column ranges
    const int nw = blockDim.x >> 5;
    const int colspan = 512 / ((nw + 3) / 4);
    const uint32_t taddr0 = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * colspan);
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = tid * 16 + i;
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = tid * 0.001f + i;
    const float a = s, b = 0.5f * s;
    const int nslot = colspan / 16;
    // initialise all columns this warp owns
    for (int j = 0; j < nslot; ++j) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                taddr0 + j * 16),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
            "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const uint32_t ta = taddr0 + (uint32_t)((it % nslot) * 16);
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = fma2(make_float2(f[2 * i], f[2 * i + 1]), make_float2(a, b), make_float2(b, a));
                f[2 * i] = v.x;
                f[2 * i + 1] = v.y;
            }
        }
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(scr) + 16u * (i * blockDim.x + tid);
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa));
                f[0] += (v.x + v.y) + (v.z + v.w);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sa + 32u * blockDim.x), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[0]) : "memory");
            }
        }
        if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4) {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta),
                "r"(r[0] + it), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
        }
        if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 4) {
            uint32_t q[16];
            const uint32_t tb = taddr0 + (uint32_t)(((it + 1) % nslot) * 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
                  "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
                : "r"(tb));
            if ((it & 3) == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += q[0] + q[15];
        }
        if ((MODE == 0) && (it & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    // round-trip check: write a pattern, read it back
    uint32_t chk[16];
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr0),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(chk[0]), "=r"(chk[1]), "=r"(chk[2]), "=r"(chk[3]), "=r"(chk[4]), "=r"(chk[5]), "=r"(chk[6]), "=r"(chk[7]),
          "=r"(chk[8]), "=r"(chk[9]), "=r"(chk[10]), "=r"(chk[11]), "=r"(chk[12]), "=r"(chk[13]), "=r"(chk[14]), "=r"(chk[15])
        : "r"(taddr0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    int bad = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) bad += (chk[i] != r[i]);
    float o = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) o += f[i];
    out[blockIdx.x * blockDim.x + tid] = o + acc + 1e6f * bad;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    if (bad && blockIdx.x == 0 && (tid & 31) == 0) printf("tmem round trip mismatch warp %d bad=%d\n", warp, bad);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

template <typename K> static int run(const char *name, K kern, int warps, double ops_per_iter, size_t smem, float *out,
                                     long long *cyc, int sms) {
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<sms, warps * 32, smem>>>(out, cyc, 1.0001f);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<sms, warps * 32, smem>>>(out, cyc, 1.0001f);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c;
    CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    const double winstr = double(warps) * ITERS * ops_per_iter;
    printf("%-44s warps=%2d  %8.3f warp-instr/clk/SM  (%lld cyc, %.3f ms)\n", name, warps, winstr / double(c), c, ms);
    return 0;
}

int main() {
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, cc %d.%d\n", pr.name, sms, pr.major, pr.minor);
    float *out;
    long long *cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * 1024));
    CK(cudaMalloc(&cyc, 8 * sms));
    for (int w : {4, 8, 16}) {
        run("FFMA 3-reg x16", k_alu<0>, w, 16, 0, out, cyc, sms);
        run("FFMA2 x8", k_alu<1>, w, 8, 0, out, cyc, sms);
        run("FMUL2 x8", k_alu<2>, w, 8, 0, out, cyc, sms);
        run("FADD2 x8", k_alu<7>, w, 8, 0, out, cyc, sms);
        run("MUFU.EX2 x8", k_alu<3>, w, 8, 0, out, cyc, sms);
        run("FFMA2 x8 + EX2 x2 (count 10)", k_alu<4>, w, 10, 0, out, cyc, sms);
        run("SHFL.BFLY x8", k_alu<5>, w, 8, 0, out, cyc, sms);
        run("FFMA2 x8 + SHFL x4 (count 12)", k_alu<6>, w, 12, 0, out, cyc, sms);
        const size_t sm = size_t(4) * w * 32 * 16;
        run("LDS.128 lane-private x8", k_smem<0>, w, 8, sm, out, cyc, sms);
        run("LDS.128 rotated x8", k_smem<1>, w, 8, sm, out, cyc, sms);
        run("STS.128 x8", k_smem<2>, w, 8, sm, out, cyc, sms);
        run("LDS.32 private x8", k_smem<3>, w, 8, sm, out, cyc, sms);
        run("LDS.128 broadcast x8", k_smem<4>, w, 8, sm, out, cyc, sms);
        run("LDS.64 broadcast x8", k_smem<5>, w, 8, sm, out, cyc, sms);
        run("LDS.32 broadcast x8", k_smem<6>, w, 8, sm, out, cyc, sms);
        run("FFMA2 x8 + LDS.128 bcast x4 (count 12)", k_smem<7>, w, 12, sm, out, cyc, sms);
        run("FFMA2 x8 + LDS.128 x2 + STS.128 x2 (12)", k_smem<8>, w, 12, sm, out, cyc, sms);
        run("tcgen05.st x16 (count 1)", k_tmem<0>, w, 1, 0, out, cyc, sms);
        run("tcgen05.ld x16 (count 1)", k_tmem<1>, w, 1, 0, out, cyc, sms);
        run("tcgen05.st+ld x16 (count 2)", k_tmem<2>, w, 2, 0, out, cyc, sms);
        run("FFMA2 x8 + tcgen05 st+ld (count 10)", k_tmem<3>, w, 10, 0, out, cyc, sms);
        run("tcgen05 st+ld + LDS.128x2 + STS.128x2 (6)", k_tmem<4>, w, 6, 0, out, cyc, sms);
    }
    return 0;
}

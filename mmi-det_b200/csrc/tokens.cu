// tokens.cu -- token layout either side of the fusion block.
//
// The detector hands the fusion block two NCHW feature maps (VIS, IR); the scan wants channels-last tokens
// (B, 2*HW, C) with the VIS tokens first and the IR tokens after them (models/common.py:1338-1343 builds the same order
// with flatten / cat / permute / contiguous, i.e. three passes over the data), and the block's output goes back to two
// NCHW maps (models/common.py:1352-1366).  Both directions are one tiled transpose each: a 32x32 tile goes through
// shared memory so that the NCHW side is read / written along the pixel axis and the token side along the channel axis,
// both fully coalesced.  Pure data movement (any 2- or 4-byte element type); each kernel is the other's adjoint.
#include "../../include/mmidet_b200.h"
#include "common.cuh"

namespace mmi {

// grid (ceil(HW / 32), ceil(C / 32), 2 * B), block (32, 8)
template <typename E, bool GATHER>
__global__ void __launch_bounds__(256) tokens_kernel(E *__restrict__ rgb, E *__restrict__ ir, E *__restrict__ tok, int C, int HW) {
    __shared__ E tile[32][33];
    const int m = blockIdx.z & 1, b = blockIdx.z >> 1;
    E *map = (m ? ir : rgb) + int64_t(b) * C * HW;                     // (C, HW)
    E *tk = tok + (int64_t(b) * 2 + m) * int64_t(HW) * C;               // (HW, C)
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    if (GATHER) {
#pragma unroll
        for (int i = threadIdx.y; i < 32; i += 8) {
            const int c = c0 + i, pp = p0 + threadIdx.x;
            if (c < C && pp < HW) tile[i][threadIdx.x] = map[int64_t(c) * HW + pp];
        }
        __syncthreads();
#pragma unroll
        for (int i = threadIdx.y; i < 32; i += 8) {
            const int pp = p0 + i, c = c0 + threadIdx.x;
            if (c < C && pp < HW) tk[int64_t(pp) * C + c] = tile[threadIdx.x][i];
        }
    } else {
#pragma unroll
        for (int i = threadIdx.y; i < 32; i += 8) {
            const int pp = p0 + i, c = c0 + threadIdx.x;
            if (c < C && pp < HW) tile[i][threadIdx.x] = tk[int64_t(pp) * C + c];
        }
        __syncthreads();
#pragma unroll
        for (int i = threadIdx.y; i < 32; i += 8) {
            const int c = c0 + i, pp = p0 + threadIdx.x;
            if (c < C && pp < HW) map[int64_t(c) * HW + pp] = tile[threadIdx.x][i];
        }
    }
}

static int tokens_launch(bool gather, void *rgb, void *ir, void *tok, int B, int C, int HW, int dtype, cudaStream_t st) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32, 2 * B), block(32, 8);
    if (dtype == MMI_F32) {
        if (gather) tokens_kernel<uint32_t, true><<<grid, block, 0, st>>>((uint32_t *)rgb, (uint32_t *)ir, (uint32_t *)tok, C, HW);
        else tokens_kernel<uint32_t, false><<<grid, block, 0, st>>>((uint32_t *)rgb, (uint32_t *)ir, (uint32_t *)tok, C, HW);
    } else {
        if (gather) tokens_kernel<uint16_t, true><<<grid, block, 0, st>>>((uint16_t *)rgb, (uint16_t *)ir, (uint16_t *)tok, C, HW);
        else tokens_kernel<uint16_t, false><<<grid, block, 0, st>>>((uint16_t *)rgb, (uint16_t *)ir, (uint16_t *)tok, C, HW);
    }
    return check_cuda(cudaGetLastError(), gather ? "tokens_gather launch" : "tokens_scatter launch");
}

}  // namespace mmi

using namespace mmi;

extern "C" {

static int tokens_check(const char *who, const void *a, const void *b, const void *c, int B, int C, int HW, int dtype) {
    if (!a || !b || !c) { set_error("%s: null pointer", who); return MMI_ERR_ARG; }
    if (B <= 0 || C <= 0 || HW <= 0 || B > 32767) { set_error("%s: bad shape (B=%d C=%d HW=%d)", who, B, C, HW); return MMI_ERR_ARG; }
    if (dtype != MMI_F32 && dtype != MMI_BF16 && dtype != MMI_F16) { set_error("%s: unknown dtype %d", who, dtype); return MMI_ERR_ARG; }
    int dev = 0, major = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) { set_error("libmmidet_b200 is built for sm_100a only"); return MMI_ERR_UNSUPPORTED; }
    return MMI_OK;
}

int mmi_tokens_gather(const void *rgb, const void *ir, void *tok, int B, int C, int HW, int dtype, void *stream) {
    if (int e = tokens_check("mmi_tokens_gather", rgb, ir, tok, B, C, HW, dtype)) return e;
    return tokens_launch(true, const_cast<void *>(rgb), const_cast<void *>(ir), tok, B, C, HW, dtype, static_cast<cudaStream_t>(stream));
}

int mmi_tokens_scatter(const void *tok, void *rgb, void *ir, int B, int C, int HW, int dtype, void *stream) {
    if (int e = tokens_check("mmi_tokens_scatter", tok, rgb, ir, B, C, HW, dtype)) return e;
    return tokens_launch(false, rgb, ir, const_cast<void *>(tok), B, C, HW, dtype, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

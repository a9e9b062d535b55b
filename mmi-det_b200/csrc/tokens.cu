// tokens.cu -- token layout either side of the fusion block.
//
// The detector hands the fusion block two NCHW feature maps (VIS, IR); the scan wants channels-last tokens
// (B, 2*HW, C) with the VIS tokens first and the IR tokens after them (models/common.py:1338-1343 builds the same order
// with flatten / cat / permute / contiguous, i.e. three passes over the data), and the block's output goes back to two
// NCHW maps (models/common.py:1352-1366).  Both directions are one tiled transpose each: a 32x32 tile goes through
// shared memory so that the NCHW side is read / written along the pixel axis and the token side along the channel axis,
// both fully coalesced.  Pure data movement (any 2- or 4-byte element type); each kernel is the other's adjoint.
// Shapes whose rows are whole 16-byte vectors take tokens_vec_kernel below; the 32x32 scalar kernel covers the rest.
#include <cstdint>

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

// Vector variant (C and HW multiples of 16 bytes' worth of elements): a 64-channel x 64-word tile (one 32-bit word = one
// fp32 pixel or two adjacent 16-bit pixels) crosses shared memory with 16-byte global accesses on both sides -- 256 B
// (fp32) / 128 B (16-bit) contiguous per tile row.  For 16-bit elements the 2 x 2 (pixel pair x channel pair) blocks are
// transposed in registers with byte permutes.  grid (ceil(HW / (64 * PU)), ceil(C / 64), 2 * B), block 256.
template <int ES, bool GATHER>
__global__ void __launch_bounds__(256) tokens_vec_kernel(unsigned char *__restrict__ rgb, unsigned char *__restrict__ ir,
                                                         unsigned char *__restrict__ tok, int C, int HW) {
    constexpr int PU = 4 / ES;  // pixels per 32-bit word
    __shared__ uint32_t tile[64][65];
    const int m = blockIdx.z & 1, b = blockIdx.z >> 1, t = threadIdx.x;
    unsigned char *map = (m ? ir : rgb) + int64_t(b) * C * HW * ES;  // (C, HW)
    unsigned char *tk = tok + (int64_t(b) * 2 + m) * int64_t(HW) * C * ES;  // (HW, C)
    const int p0 = blockIdx.x * 64 * PU, c0 = blockIdx.y * 64;
    // NCHW side: thread (channel row, group of four words)
    auto map_ptr = [&](int c, int wg) { return reinterpret_cast<uint4 *>(map + (int64_t(c0 + c) * HW + p0 + 4 * wg * PU) * ES); };
    auto map_ok = [&](int c, int wg) { return c0 + c < C && p0 + 4 * wg * PU < HW; };
    if (GATHER) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = (t >> 4) + 16 * i, wg = t & 15;
            if (map_ok(c, wg)) {
                const uint4 v = *map_ptr(c, wg);
                tile[c][4 * wg] = v.x, tile[c][4 * wg + 1] = v.y, tile[c][4 * wg + 2] = v.z, tile[c][4 * wg + 3] = v.w;
            }
        }
        __syncthreads();
        if constexpr (ES == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pp = (t >> 4) + 16 * i, cg = t & 15;
                if (p0 + pp < HW && c0 + 4 * cg < C)
                    *reinterpret_cast<uint4 *>(tk + (int64_t(p0 + pp) * C + c0 + 4 * cg) * 4) =
                        make_uint4(tile[4 * cg][pp], tile[4 * cg + 1][pp], tile[4 * cg + 2][pp], tile[4 * cg + 3][pp]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int pp = (t >> 3) + 32 * i, cg = t & 7;
                if (p0 + 2 * pp < HW && c0 + 8 * cg < C) {
                    uint32_t w[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) w[k] = tile[8 * cg + k][pp];
                    unsigned char *dst = tk + (int64_t(p0 + 2 * pp) * C + c0 + 8 * cg) * 2;
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(__byte_perm(w[0], w[1], 0x5410), __byte_perm(w[2], w[3], 0x5410),
                                                                 __byte_perm(w[4], w[5], 0x5410), __byte_perm(w[6], w[7], 0x5410));
                    *reinterpret_cast<uint4 *>(dst + int64_t(C) * 2) =
                        make_uint4(__byte_perm(w[0], w[1], 0x7632), __byte_perm(w[2], w[3], 0x7632),
                                   __byte_perm(w[4], w[5], 0x7632), __byte_perm(w[6], w[7], 0x7632));
                }
            }
        }
    } else {
        if constexpr (ES == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pp = (t >> 4) + 16 * i, cg = t & 15;
                if (p0 + pp < HW && c0 + 4 * cg < C) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(tk + (int64_t(p0 + pp) * C + c0 + 4 * cg) * 4);
                    tile[4 * cg][pp] = v.x, tile[4 * cg + 1][pp] = v.y, tile[4 * cg + 2][pp] = v.z, tile[4 * cg + 3][pp] = v.w;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int pp = (t >> 3) + 32 * i, cg = t & 7;
                if (p0 + 2 * pp < HW && c0 + 8 * cg < C) {
                    const unsigned char *src = tk + (int64_t(p0 + 2 * pp) * C + c0 + 8 * cg) * 2;
                    const uint4 a = *reinterpret_cast<const uint4 *>(src), bb = *reinterpret_cast<const uint4 *>(src + int64_t(C) * 2);
                    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        tile[8 * cg + 2 * j][pp] = __byte_perm(av[j], bv[j], 0x5410);
                        tile[8 * cg + 2 * j + 1][pp] = __byte_perm(av[j], bv[j], 0x7632);
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = (t >> 4) + 16 * i, wg = t & 15;
            if (map_ok(c, wg))
                *map_ptr(c, wg) = make_uint4(tile[c][4 * wg], tile[c][4 * wg + 1], tile[c][4 * wg + 2], tile[c][4 * wg + 3]);
        }
    }
}

static int tokens_launch(bool gather, void *rgb, void *ir, void *tok, int B, int C, int HW, int dtype, cudaStream_t st) {
    const int es = dtype == MMI_F32 ? 4 : 2, vec = 16 / es;
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (C % vec == 0 && HW % vec == 0 && al16(rgb) && al16(ir) && al16(tok)) {
        unsigned char *r8 = static_cast<unsigned char *>(rgb), *i8 = static_cast<unsigned char *>(ir), *t8 = static_cast<unsigned char *>(tok);
        const dim3 grid((HW + 64 * (4 / es) - 1) / (64 * (4 / es)), (C + 63) / 64, 2 * B);
        if (es == 4) {
            if (gather) tokens_vec_kernel<4, true><<<grid, 256, 0, st>>>(r8, i8, t8, C, HW);
            else tokens_vec_kernel<4, false><<<grid, 256, 0, st>>>(r8, i8, t8, C, HW);
        } else {
            if (gather) tokens_vec_kernel<2, true><<<grid, 256, 0, st>>>(r8, i8, t8, C, HW);
            else tokens_vec_kernel<2, false><<<grid, 256, 0, st>>>(r8, i8, t8, C, HW);
        }
        return check_cuda(cudaGetLastError(), gather ? "tokens_gather launch" : "tokens_scatter launch");
    }
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

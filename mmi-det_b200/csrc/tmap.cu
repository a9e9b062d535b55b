// tmap.cu -- host-side construction of TMA tensor maps.  cuTensorMapEncodeTiled is resolved through the runtime's
// driver entry-point query, so the library does not link libcuda directly.
#include <cudaTypedefs.h>

#include "../../include/mmidet_b200.h"
#include "common.cuh"

namespace mmi {

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !p) {
        (void)cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    return fn;
}

int encode_tmap_3d(CUtensorMap *map, const void *base, int dtype, uint64_t nb, uint64_t L, uint64_t cols, uint64_t pitch_bytes,
                   uint32_t box_rows, uint32_t box_cols);

// (B, L, cols) tensor addressed as {column, t, b}: a box never crosses a batch boundary, so a tile that runs past L is
// zero-filled on load and clipped on store -- ragged sequence tails need no special case in the kernels.
// Encoding a tensor map is a driver call; a training loop presents the same few (pointer, shape, box) tuples every
// step, so the encoded descriptors are memoised (direct-mapped, per thread -> no locking, no cross-thread state).
struct TmapKey {
    const void *base;
    uint64_t nb, L, cols, pitch;
    uint32_t box_rows, box_cols;
    int dtype;
    bool operator==(const TmapKey &o) const {
        return base == o.base && nb == o.nb && L == o.L && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows &&
               box_cols == o.box_cols && dtype == o.dtype;
    }
};
struct TmapSlot {
    TmapKey key;
    CUtensorMap map;
    bool valid;
};
static thread_local TmapSlot g_tmap_cache[256];

int make_tmap_3d(CUtensorMap *map, const void *base, int dtype, uint64_t nb, uint64_t L, uint64_t cols, uint64_t pitch_bytes,
                 uint32_t box_rows, uint32_t box_cols) {
    const TmapKey key{base, nb, L, cols, pitch_bytes, box_rows, box_cols, dtype};
    uint64_t hsh = reinterpret_cast<uint64_t>(base) * 0x9E3779B97F4A7C15ull ^ (L * 0xC2B2AE3D27D4EB4Full) ^ (cols << 7) ^
                   (uint64_t(box_rows) << 17) ^ (uint64_t(box_cols) << 29) ^ (pitch_bytes * 31) ^ (nb << 3) ^ uint64_t(dtype);
    TmapSlot &slot = g_tmap_cache[(hsh >> 32 ^ hsh) & 255];
    if (slot.valid && slot.key == key) {
        *map = slot.map;
        return MMI_OK;
    }
    if (int e = encode_tmap_3d(map, base, dtype, nb, L, cols, pitch_bytes, box_rows, box_cols)) return e;
    slot.key = key;
    slot.map = *map;
    slot.valid = true;
    return MMI_OK;
}

int encode_tmap_3d(CUtensorMap *map, const void *base, int dtype, uint64_t nb, uint64_t L, uint64_t cols, uint64_t pitch_bytes,
                   uint32_t box_rows, uint32_t box_cols) {
    auto fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the installed driver");
        return MMI_ERR_CUDA;
    }
    CUtensorMapDataType dt = dtype == MMI_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                             : dtype == MMI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const cuuint64_t dims[3] = {cols, L, nb};
    const cuuint64_t strides[2] = {pitch_bytes, pitch_bytes * L};
    const cuuint32_t box[3] = {box_cols, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed with CUresult %d (B=%llu L=%llu cols=%llu pitch=%llu box=%ux%u)", int(r),
                  (unsigned long long)nb, (unsigned long long)L, (unsigned long long)cols, (unsigned long long)pitch_bytes,
                  box_rows, box_cols);
        return MMI_ERR_CUDA;
    }
    return MMI_OK;
}

}  // namespace mmi

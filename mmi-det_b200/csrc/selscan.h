// selscan.h -- parameter blocks and launch entry points of the fused selective-scan kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmi {

constexpr int kChunk = 16;     // state-checkpoint interval in timesteps == the chunk one backward warp scans per super-tile
constexpr int kFwdChunk = 16;  // default chunk one forward warp scans per super-tile (shape-test configs also use 8)

struct FwdParams {
    const void *x, *delta, *z, *Bm, *Cm;
    const float *A, *D, *h0;
    void *out;
    float *hT, *chk;
    int B, L, ED;
    int64_t x_ld, d_ld, z_ld, o_ld;
    int flags;
    // L split over CTAs (filled by the launcher): segment s covers super-tiles [s*seg_tiles, (s+1)*seg_tiles)
    int nseg, seg_tiles, ntile_c;
    unsigned *seg_ticket, *seg_flags;  // start-order ticket; per (b, segment, channel tile) publication counters
    float *seg_ws;                     // (B, nseg, ED, N + 1): end state of the segment from zero | sum of delta
};

struct BwdParams {
    const void *x, *delta, *z, *Bm, *Cm, *dout;
    const float *A, *D, *chk;
    void *dx, *ddelta, *dz, *dBm, *dCm;
    float *dA, *dD;
    float *ws_bc;  // (B, L, ntile_c, 2, N) fp32 per-CTA partial dB/dC
    float *ws_ad;  // (B, ED, N + 1) fp32 per-batch partial dA / dD
    int B, L, ED;
    int ntile_c;   // channel tiles per row (grid.x)
    int64_t x_ld, d_ld, z_ld, g_ld;
    int flags;
    // L split over CTAs (filled by the launcher), as in FwdParams
    int nseg, seg_tiles;
    unsigned *seg_ticket, *seg_flags;
    float *seg_ws;  // (B, nseg, ED, N + 1): a*g leaving the segment when nothing enters it | sum of delta
};

// second-generation kernels (selscan_fwd2.cu / selscan_bwd2.cu): persistent grid, items = (chain, L segment) handed out by
// ticket in dependency order; a chain = one batch element x 32 channels
struct SegSched {
    int nseg, seg_tiles;   // L segments per chain, super-tiles per segment
    int ntile_c, nchains;  // channel tiles per batch element, B * ntile_c
    int nitems;            // nchains * nseg
    unsigned *ticket;      // [1] item counter
    unsigned *done;        // [nchains] number of finished segments of the chain
    float *carry;          // [nchains][8][32] float2: state (forward) / a*g (backward) handed to the next segment
};
struct Bwd2Params {
    BwdParams b;
    SegSched s;
};
struct Fwd2Params {
    FwdParams f;
    SegSched s;
};

// first generation (look-back L split for small grids) and second generation entry points; selscan_*_launch dispatch
int selscan_fwd1_launch(const FwdParams &p, int dtype, void *ws, cudaStream_t st);
int64_t selscan_fwd1_ws_bytes(int B, int ED);
int selscan_bwd1_launch(BwdParams p, int dtype, void *ws, cudaStream_t st);
int64_t selscan_bwd1_ws_bytes(int B, int L, int ED);
int selscan_fwd2_launch(const FwdParams &p, int dtype, void *ws, cudaStream_t st);
int64_t selscan_fwd2_ws_bytes(int B, int L, int ED);
int selscan_bwd2_launch(const BwdParams &p, int dtype, void *ws, cudaStream_t st);
int64_t selscan_bwd2_ws_bytes(int B, int L, int ED);
int selscan_bwd3_launch(const BwdParams &p, int dtype, void *ws, cudaStream_t st);  // 16-warp form, same workspace as bwd2
bool selscan_use_v2(int B, int L, int ED, int flags);  // which generation a call takes (shape heuristic, MMI_FLAG_CFG override)
int seg_sched_plan(int B, int L, int ED, int flags, SegSched *s);  // fills nseg / seg_tiles / ntile_c / nchains / nitems

int selscan_fwd_launch(const FwdParams &p, int dtype, void *ws, cudaStream_t st);
int64_t selscan_fwd_ws_bytes(int B, int ED);
int selscan_bwd_launch(BwdParams p, int dtype, void *ws, cudaStream_t st);
int64_t selscan_bwd_ws_bytes(int B, int L, int ED);
int sm_count();

}  // namespace mmi

// selscan_dispatch.cu -- which generation of the scan kernels a call takes.
//
//   second generation (selscan_fwd2.cu / selscan_bwd2.cu): persistent grid over (32-channel chain, L segment) items; needs
//       enough independent chains to occupy the GPU -- the training shapes (B * ceil(ED / 32) >= two thirds of the SMs)
//   first generation (selscan_fwd.cu / selscan_bwd.cu): one CTA per 32 / 64-channel tile, L split across CTAs with published
//       segment summaries (decoupled look-back) -- small batches and inference, where only splitting L can fill the GPU
// MMI_FLAG_CFG bits: 8 forces the second generation, 9 the first, 1..6 are first-generation CTA shapes, 10 the 4-warp second-
// generation forward, 11 the 16-warp backward (selscan_bwd3.cu: measured slower, kept as the tested counter-example) -- tuning / tests.
#include <algorithm>

#include "../../include/mmidet_b200.h"
#include "common.cuh"
#include "selscan.h"

namespace mmi {

bool selscan_use_v2(int B, int L, int ED, int flags) {
    const int cfg = (flags & MMI_FLAG_CFG_MASK) >> MMI_FLAG_CFG_SHIFT;
    if (cfg == 8 || cfg == 10 || cfg == 11) return true;
    if (cfg != 0) return false;
    (void)L;
    // measured crossover (profiles/r02_scan_generations.txt): a second-generation CTA takes the same time per chain however
    // few chains there are, the first generation splits L to fill the GPU -- 64 chains: 0.405 vs 0.543 ms, 128: 0.771 vs 0.557 ms
    const int64_t chains = int64_t(B) * ((ED + 31) / 32);
    return chains * 3 >= 2 * int64_t(sm_count());
}

int64_t selscan_bwd_ws_bytes(int B, int L, int ED) { return std::max(selscan_bwd1_ws_bytes(B, L, ED), selscan_bwd2_ws_bytes(B, L, ED)); }
int64_t selscan_fwd_ws_bytes(int B, int ED) { return std::max(selscan_fwd1_ws_bytes(B, ED), selscan_fwd2_ws_bytes(B, 0, ED)); }

int selscan_bwd_launch(BwdParams p, int dtype, void *ws, cudaStream_t st) {
    const int cfg = (p.flags & MMI_FLAG_CFG_MASK) >> MMI_FLAG_CFG_SHIFT;
    if (cfg == 11) return selscan_bwd3_launch(p, dtype, ws, st);
    if (selscan_use_v2(p.B, p.L, p.ED, p.flags)) return selscan_bwd2_launch(p, dtype, ws, st);
    return selscan_bwd1_launch(p, dtype, ws, st);
}

// The forward stays on the first generation unless forced: its two independent 4-warp CTAs per SM hide each other's
// barrier stalls, which the 8-warp persistent CTA cannot (measured at the bench shape, profiles/r02_scan_generations.txt:
// 0.372 ms vs 0.426 ms with checkpoints); the backward, whose 255-register / tensor-memory footprint allows one CTA per SM
// either way, gains from the second generation (1.224 -> 1.01 ms).
int selscan_fwd_launch(const FwdParams &p, int dtype, void *ws, cudaStream_t st) {
    const int cfg = (p.flags & MMI_FLAG_CFG_MASK) >> MMI_FLAG_CFG_SHIFT;
    if (ws && (cfg == 8 || cfg == 10)) return selscan_fwd2_launch(p, dtype, ws, st);
    return selscan_fwd1_launch(p, dtype, ws, st);
}

}  // namespace mmi

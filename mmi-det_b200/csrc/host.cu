// host.cu -- host-buffer entry point (bench.py `e2e`): H2D of every input, forward + backward on a private stream,
// D2H of every output, one synchronisation at the end.  Owns a reusable device staging workspace.
#include <cstdlib>

#include "../../include/mmidet_b200.h"
#include "common.cuh"
#include "selscan.h"

namespace mmi {

struct HostWs {
    void *dev = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    int device = -1;
};
static HostWs g_hws;

static int hws_reserve(size_t bytes) {
    int dev = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    if (g_hws.device != dev && g_hws.dev) {
        cudaFree(g_hws.dev);
        g_hws.dev = nullptr;
        g_hws.bytes = 0;
    }
    g_hws.device = dev;
    if (!g_hws.stream)
        if (int e = check_cuda(cudaStreamCreateWithFlags(&g_hws.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return e;
    if (bytes > g_hws.bytes) {
        if (g_hws.dev) cudaFree(g_hws.dev);
        g_hws.dev = nullptr;
        g_hws.bytes = 0;
        if (int e = check_cuda(cudaMalloc(&g_hws.dev, bytes), "cudaMalloc(host-entry workspace)")) return e;
        g_hws.bytes = bytes;
    }
    return MMI_OK;
}

static size_t al(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace mmi

using namespace mmi;

extern "C" {

void mmi_host_workspace_free(void) {
    if (g_hws.dev) cudaFree(g_hws.dev);
    if (g_hws.stream) cudaStreamDestroy(g_hws.stream);
    g_hws = HostWs{};
}

int mmi_selscan_fwd_bwd_host(const void *x, const void *delta, const void *z, const float *A, const void *Bm,
                             const void *Cm, const float *D, const void *dout, void *out, void *dx, void *ddelta,
                             void *dz, float *dA, void *dBm, void *dCm, float *dD, int B, int L, int ED, int N,
                             int dtype, int flags) {
    if (!x || !delta || !A || !Bm || !Cm || !D || !dout || !out || !dx || !ddelta || !dA || !dBm || !dCm || !dD) {
        set_error("mmi_selscan_fwd_bwd_host: null required pointer");
        return MMI_ERR_ARG;
    }
    if ((z == nullptr) != (dz == nullptr)) { set_error("mmi_selscan_fwd_bwd_host: z and dz go together"); return MMI_ERR_ARG; }
    if (N != kN || ED % 8 || B <= 0 || L <= 0 || ED <= 0) { set_error("mmi_selscan_fwd_bwd_host: bad shape"); return MMI_ERR_ARG; }
    const size_t es = dtype == MMI_F32 ? 4 : 2;
    const size_t big = al(size_t(B) * L * ED * es), bc = al(size_t(B) * L * N * es), an = al(size_t(ED) * N * 4), dn = al(size_t(ED) * 4);
    const int nchk = (L + kChunk - 1) / kChunk;
    const size_t chkb = al(size_t(B) * nchk * ED * N * 4), wsb = al(size_t(mmi_selscan_bwd_ws_bytes(B, L, ED, N)));
    // x delta z dout out dx ddelta dz | B C dB dC | A dA | D dD | chk | ws
    const size_t total = 8 * big + 4 * bc + 2 * an + 2 * dn + chkb + wsb;
    if (int e = hws_reserve(total)) return e;
    char *p = static_cast<char *>(g_hws.dev);
    auto take = [&](size_t n) { char *r = p; p += n; return r; };
    char *d_x = take(big), *d_d = take(big), *d_z = take(big), *d_g = take(big), *d_o = take(big), *d_dx = take(big),
         *d_dd = take(big), *d_dz = take(big);
    char *d_B = take(bc), *d_C = take(bc), *d_dB = take(bc), *d_dC = take(bc);
    char *d_A = take(an), *d_dA = take(an), *d_D = take(dn), *d_dD = take(dn), *d_chk = take(chkb), *d_ws = take(wsb);
    cudaStream_t st = g_hws.stream;
    const size_t nbig = size_t(B) * L * ED * es, nbc = size_t(B) * L * N * es;
#define MMI_CP(dst, src, n, kind) \
    if (int e = check_cuda(cudaMemcpyAsync(dst, src, n, kind, st), "cudaMemcpyAsync")) return e
    MMI_CP(d_x, x, nbig, cudaMemcpyHostToDevice);
    MMI_CP(d_d, delta, nbig, cudaMemcpyHostToDevice);
    if (z) MMI_CP(d_z, z, nbig, cudaMemcpyHostToDevice);
    MMI_CP(d_g, dout, nbig, cudaMemcpyHostToDevice);
    MMI_CP(d_B, Bm, nbc, cudaMemcpyHostToDevice);
    MMI_CP(d_C, Cm, nbc, cudaMemcpyHostToDevice);
    MMI_CP(d_A, A, size_t(ED) * N * 4, cudaMemcpyHostToDevice);
    MMI_CP(d_D, D, size_t(ED) * 4, cudaMemcpyHostToDevice);
    if (int e = mmi_selscan_fwd(d_x, d_d, z ? d_z : nullptr, (const float *)d_A, d_B, d_C, (const float *)d_D, nullptr, d_o,
                                nullptr, (float *)d_chk, B, L, ED, N, ED, ED, ED, ED, kChunk, dtype, flags, st))
        return e;
    if (int e = mmi_selscan_bwd(d_x, d_d, z ? d_z : nullptr, (const float *)d_A, d_B, d_C, (const float *)d_D, d_g,
                                (const float *)d_chk, d_dx, d_dd, z ? d_dz : nullptr, (float *)d_dA, d_dB, d_dC,
                                (float *)d_dD, d_ws, B, L, ED, N, ED, ED, ED, ED, kChunk, dtype, flags, st))
        return e;
    MMI_CP(out, d_o, nbig, cudaMemcpyDeviceToHost);
    MMI_CP(dx, d_dx, nbig, cudaMemcpyDeviceToHost);
    MMI_CP(ddelta, d_dd, nbig, cudaMemcpyDeviceToHost);
    if (z) MMI_CP(dz, d_dz, nbig, cudaMemcpyDeviceToHost);
    MMI_CP(dBm, d_dB, nbc, cudaMemcpyDeviceToHost);
    MMI_CP(dCm, d_dC, nbc, cudaMemcpyDeviceToHost);
    MMI_CP(dA, d_dA, size_t(ED) * N * 4, cudaMemcpyDeviceToHost);
    MMI_CP(dD, d_dD, size_t(ED) * 4, cudaMemcpyDeviceToHost);
#undef MMI_CP
    return check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
}

}  // extern "C"

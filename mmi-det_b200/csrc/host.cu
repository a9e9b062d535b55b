// host.cu -- host-buffer entry point (bench.py `e2e`): the batch flows in chunks through H2D -> forward + backward -> D2H
// on three private streams so both PCIe directions and the kernels overlap.  Owns a reusable device staging workspace.
#include <cstdlib>

#include "../../include/mmidet_b200.h"
#include "common.cuh"
#include "selscan.h"

namespace mmi {

struct HostWs {
    void *dev = nullptr;
    size_t bytes = 0;
    float *hsum = nullptr;  // pinned host staging for the per-chunk dA / dD partials
    size_t hsum_bytes = 0;
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    cudaEvent_t in_done[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
    int device = -1;
};
static HostWs g_hws;

static int hws_reserve(size_t bytes, size_t hsum_bytes) {
    int dev = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    if (g_hws.device != dev && g_hws.dev) {
        cudaFree(g_hws.dev);
        g_hws.dev = nullptr;
        g_hws.bytes = 0;
    }
    g_hws.device = dev;
    if (!g_hws.s_in) {
        cudaStream_t *ss[3] = {&g_hws.s_in, &g_hws.s_comp, &g_hws.s_out};
        for (auto s : ss)
            if (int e = check_cuda(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking), "cudaStreamCreate")) return e;
        for (int i = 0; i < 2; ++i) {
            cudaEvent_t *ev[3] = {&g_hws.in_done[i], &g_hws.comp_done[i], &g_hws.out_done[i]};
            for (auto e_ : ev)
                if (int e = check_cuda(cudaEventCreateWithFlags(e_, cudaEventDisableTiming), "cudaEventCreate")) return e;
        }
    }
    if (bytes > g_hws.bytes) {
        if (g_hws.dev) cudaFree(g_hws.dev);
        g_hws.dev = nullptr;
        g_hws.bytes = 0;
        if (int e = check_cuda(cudaMalloc(&g_hws.dev, bytes), "cudaMalloc(host-entry workspace)")) return e;
        g_hws.bytes = bytes;
    }
    if (hsum_bytes > g_hws.hsum_bytes) {
        if (g_hws.hsum) cudaFreeHost(g_hws.hsum);
        g_hws.hsum = nullptr;
        g_hws.hsum_bytes = 0;
        if (int e = check_cuda(cudaMallocHost(reinterpret_cast<void **>(&g_hws.hsum), hsum_bytes), "cudaMallocHost")) return e;
        g_hws.hsum_bytes = hsum_bytes;
    }
    return MMI_OK;
}

static size_t al(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace mmi

using namespace mmi;

extern "C" {

void mmi_host_workspace_free(void) {
    if (g_hws.dev) cudaFree(g_hws.dev);
    if (g_hws.hsum) cudaFreeHost(g_hws.hsum);
    cudaStream_t ss[3] = {g_hws.s_in, g_hws.s_comp, g_hws.s_out};
    for (auto s : ss)
        if (s) cudaStreamDestroy(s);
    for (int i = 0; i < 2; ++i) {
        cudaEvent_t ev[3] = {g_hws.in_done[i], g_hws.comp_done[i], g_hws.out_done[i]};
        for (auto e : ev)
            if (e) cudaEventDestroy(e);
    }
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
    // The batch is cut into chunks that flow through a 3-stage pipeline on three streams -- H2D of chunk i+1, forward +
    // backward of chunk i and D2H of chunk i-1 run concurrently (PCIe is full duplex) -- over two device staging slots.
    const size_t es = dtype == MMI_F32 ? 4 : 2;
    // 8 chunks by default (MMI_HOST_CHUNKS overrides, 1..64): fill and drain of the pipeline cost 2 / (nchunk + 2) of the step,
    // smaller chunks cost kernel efficiency and per-copy overhead -- measured at B=16 L=6400 ED=512 fp32: 4 / 6 / 8 / 12 / 16
    // chunks = 25.0 / 22.1 / 22.8 / 21.8 / 23.5 ms per step, i.e. PCIe-bound and flat (profiles/r02_scan_generations.txt)
    static const int want = [] {
        const char *e = getenv("MMI_HOST_CHUNKS");
        const int v = e ? atoi(e) : 8;
        return v < 1 ? 1 : (v > 64 ? 64 : v);
    }();
    const int nchunk = B < want ? B : want, cb = (B + nchunk - 1) / nchunk;
    const size_t big = al(size_t(cb) * L * ED * es), bc = al(size_t(cb) * L * N * es), an = al(size_t(ED) * N * 4), dn = al(size_t(ED) * 4);
    const int nchk = (L + kChunk - 1) / kChunk;
    const size_t chkb = al(size_t(cb) * nchk * ED * N * 4), wsb = al(size_t(mmi_selscan_bwd_ws_bytes(cb, L, ED, N)));
    // per slot: x delta z dout out dx ddelta dz | B C dB dC ; shared: A D | per chunk dA dD | chk | ws
    const size_t slot = 8 * big + 4 * bc;
    const size_t total = 2 * slot + an + dn + size_t(nchunk) * (an + dn) + chkb + wsb;
    const size_t hs = size_t(nchunk) * (size_t(ED) * N + ED) * 4;
    if (int e = hws_reserve(total, hs)) return e;
    char *p = static_cast<char *>(g_hws.dev);
    auto take = [&](size_t n) { char *r = p; p += n; return r; };
    char *slots[2] = {take(slot), take(slot)};
    char *d_A = take(an), *d_D = take(dn), *d_dAD = take(size_t(nchunk) * (an + dn)), *d_chk = take(chkb), *d_ws = take(wsb);
    cudaStream_t s_in = g_hws.s_in, s_comp = g_hws.s_comp, s_out = g_hws.s_out;
#define MMI_CK(call, what)     if (int e = check_cuda(call, what)) return e
    MMI_CK(cudaMemcpyAsync(d_A, A, size_t(ED) * N * 4, cudaMemcpyHostToDevice, s_in), "cudaMemcpyAsync(A)");
    MMI_CK(cudaMemcpyAsync(d_D, D, size_t(ED) * 4, cudaMemcpyHostToDevice, s_in), "cudaMemcpyAsync(D)");
    const size_t rowb = size_t(L) * ED * es, rowbc = size_t(L) * N * es;  // bytes per batch element
    for (int i = 0; i < nchunk; ++i) {
        const int b0 = i * cb, nb = (B - b0) < cb ? (B - b0) : cb;
        if (nb <= 0) break;
        const int sl = i & 1;
        char *q = slots[sl];
        char *d_x = q, *d_d = q + big, *d_z = q + 2 * big, *d_g = q + 3 * big, *d_o = q + 4 * big, *d_dx = q + 5 * big,
             *d_dd = q + 6 * big, *d_dz = q + 7 * big;
        char *d_B = q + 8 * big, *d_C = d_B + bc, *d_dB = d_C + bc, *d_dC = d_dB + bc;
        float *d_dA = reinterpret_cast<float *>(d_dAD + size_t(i) * (an + dn)), *d_dD = reinterpret_cast<float *>(reinterpret_cast<char *>(d_dA) + an);
        const size_t nbig = size_t(nb) * rowb, nbc = size_t(nb) * rowbc, ob = size_t(b0) * rowb, obc = size_t(b0) * rowbc;
        const char *hx = static_cast<const char *>(x) + ob, *hd = static_cast<const char *>(delta) + ob,
                   *hg = static_cast<const char *>(dout) + ob;
        // stage 1: host -> device (after the slot's previous occupant has been copied out)
        if (i >= 2) MMI_CK(cudaStreamWaitEvent(s_in, g_hws.out_done[sl], 0), "cudaStreamWaitEvent");
        MMI_CK(cudaMemcpyAsync(d_x, hx, nbig, cudaMemcpyHostToDevice, s_in), "cudaMemcpyAsync(x)");
        MMI_CK(cudaMemcpyAsync(d_d, hd, nbig, cudaMemcpyHostToDevice, s_in), "cudaMemcpyAsync(delta)");
        if (z) MMI_CK(cudaMemcpyAsync(d_z, static_cast<const char *>(z) + ob, nbig, cudaMemcpyHostToDevice, s_in), "cudaMemcpyAsync(z)");
        MMI_CK(cudaMemcpyAsync(d_g, hg, nbig, cudaMemcpyHostToDevice, s_in), "cudaMemcpyAsync(dout)");
        MMI_CK(cudaMemcpyAsync(d_B, static_cast<const char *>(Bm) + obc, nbc, cudaMemcpyHostToDevice, s_in), "cudaMemcpyAsync(B)");
        MMI_CK(cudaMemcpyAsync(d_C, static_cast<const char *>(Cm) + obc, nbc, cudaMemcpyHostToDevice, s_in), "cudaMemcpyAsync(C)");
        MMI_CK(cudaEventRecord(g_hws.in_done[sl], s_in), "cudaEventRecord");
        // stage 2: forward + backward
        MMI_CK(cudaStreamWaitEvent(s_comp, g_hws.in_done[sl], 0), "cudaStreamWaitEvent");
        if (int e = mmi_selscan_fwd(d_x, d_d, z ? d_z : nullptr, (const float *)d_A, d_B, d_C, (const float *)d_D, nullptr, d_o,
                                    nullptr, (float *)d_chk, nullptr, nb, L, ED, N, ED, ED, ED, ED, kChunk, dtype, flags, s_comp))
            return e;
        if (int e = mmi_selscan_bwd(d_x, d_d, z ? d_z : nullptr, (const float *)d_A, d_B, d_C, (const float *)d_D, d_g,
                                    (const float *)d_chk, d_dx, d_dd, z ? d_dz : nullptr, d_dA, d_dB, d_dC, d_dD, d_ws, nb, L,
                                    ED, N, ED, ED, ED, ED, kChunk, dtype, flags, s_comp))
            return e;
        MMI_CK(cudaEventRecord(g_hws.comp_done[sl], s_comp), "cudaEventRecord");
        // stage 3: device -> host
        MMI_CK(cudaStreamWaitEvent(s_out, g_hws.comp_done[sl], 0), "cudaStreamWaitEvent");
        MMI_CK(cudaMemcpyAsync(static_cast<char *>(out) + ob, d_o, nbig, cudaMemcpyDeviceToHost, s_out), "cudaMemcpyAsync(out)");
        MMI_CK(cudaMemcpyAsync(static_cast<char *>(dx) + ob, d_dx, nbig, cudaMemcpyDeviceToHost, s_out), "cudaMemcpyAsync(dx)");
        MMI_CK(cudaMemcpyAsync(static_cast<char *>(ddelta) + ob, d_dd, nbig, cudaMemcpyDeviceToHost, s_out), "cudaMemcpyAsync(ddelta)");
        if (z) MMI_CK(cudaMemcpyAsync(static_cast<char *>(dz) + ob, d_dz, nbig, cudaMemcpyDeviceToHost, s_out), "cudaMemcpyAsync(dz)");
        MMI_CK(cudaMemcpyAsync(static_cast<char *>(dBm) + obc, d_dB, nbc, cudaMemcpyDeviceToHost, s_out), "cudaMemcpyAsync(dB)");
        MMI_CK(cudaMemcpyAsync(static_cast<char *>(dCm) + obc, d_dC, nbc, cudaMemcpyDeviceToHost, s_out), "cudaMemcpyAsync(dC)");
        float *h_part = g_hws.hsum + size_t(i) * (size_t(ED) * N + ED);
        MMI_CK(cudaMemcpyAsync(h_part, d_dA, size_t(ED) * N * 4, cudaMemcpyDeviceToHost, s_out), "cudaMemcpyAsync(dA)");
        MMI_CK(cudaMemcpyAsync(h_part + size_t(ED) * N, d_dD, size_t(ED) * 4, cudaMemcpyDeviceToHost, s_out), "cudaMemcpyAsync(dD)");
        MMI_CK(cudaEventRecord(g_hws.out_done[sl], s_out), "cudaEventRecord");
    }
    MMI_CK(cudaStreamSynchronize(s_out), "cudaStreamSynchronize");
#undef MMI_CK
    // dA / dD are sums over the batch: add the chunks' partials (ED*(N+1) floats each) on the host
    const size_t na = size_t(ED) * N, nd = size_t(ED), stride = na + nd;
    const int used = (B + cb - 1) / cb;
    for (size_t k = 0; k < na; ++k) {
        float v = 0.f;
        for (int i = 0; i < used; ++i) v += g_hws.hsum[size_t(i) * stride + k];
        dA[k] = v;
    }
    for (size_t k = 0; k < nd; ++k) {
        float v = 0.f;
        for (int i = 0; i < used; ++i) v += g_hws.hsum[size_t(i) * stride + na + k];
        dD[k] = v;
    }
    return MMI_OK;
}

}  // extern "C"

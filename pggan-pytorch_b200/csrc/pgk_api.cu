// pgk_api.cu -- shape dispatch for the two GEMM-shaped entry points, and the per-launch timing hooks bench.py uses
// for the roofline.  The tcgen05 path (pgk_conv_tc.cu) serves the tensor-core friendly layers; everything else runs
// on the CUDA-core implicit GEMM.  Both are this library's own sm_100a kernels -- there is no library or CPU
// fallback.
#include <stdlib.h>

#include <vector>

#include "pgk_common.cuh"

extern "C" int pgk_conv_simt(const void* x, int P, long long x_ps, int N, int H, int W, int Cin, int Cout, int KS,
                             int ups, const float* wf, const float* bias, const float* posT, const float* pos_s,
                             int act, const void* mask_ref, long long mask_ps, float out_scale, void* out,
                             long long out_ps, pgk_stream_t stream);
extern "C" int pgk_wgrad_simt(const void* x, long long x_ps, const void* g, long long g_ps, int P, int H, int W,
                              int Cin, int Cout, int KS, int ups, int ngroups, int group_n, const int* xoff,
                              const int* goff, float* dwp, pgk_stream_t stream);

extern "C" int pgk_conv_tc_supported(int N, int H, int W, int Cin, int Cout, int KS, int ups);
extern "C" int pgk_conv_tc(const void* x, int P, int Pr, long long x_ps, int N, int H, int W, int Cin, int Cout, int KS,
                           const void* wt, long long wt_ps, const float* bias, const float* posT, const float* pos_s,
                           int act, const void* mask_ref, long long mask_ps, float out_scale, void* out,
                           long long out_ps, pgk_stream_t stream, int fp16_x, int fp16_w, float acc_scale, float* pn_r);
extern "C" int pgk_conv_tc_fuses_pixelnorm(int Cout, int split_acc);
extern "C" int pgk_wgrad_tc_supported(int H, int W, int Cin, int Cout, int KS, int ups, int ngroups, int group_n);
extern "C" int pgk_wgrad_tc(const void* x, long long x_ps, const void* g, long long g_ps, int P, int Pr, int H, int W,
                            int Cin, int Cout, int KS, int ngroups, int group_n, const int* xoff, const int* goff,
                            float* dwp, float* db, unsigned bias_mask, int* bias_fused, pgk_stream_t stream);

extern "C" int pgk_conv_thin_supported(int N, int H, int W, int Cin, int Cout, int KS, int ups);
extern "C" int pgk_conv_thin(const void* x, int P, int Pr, long long x_ps, int N, int H, int W, int Cin, int Cout,
                             const void* wpack, long long wpack_ps, const float* bias, int act, const void* mask_ref,
                             long long mask_ps, float out_scale, void* out, long long out_ps, float* pn_r,
                             pgk_stream_t stream);
extern "C" int pgk_conv_thin_fuses_pixelnorm(int Cout);

extern "C" int pgk_wgrad_thin_supported(int H, int W, int Cin, int Cout, int KS, int ups, int ngroups, int group_n,
                                        int Pr);
int pgk_wgrad_thin_direct(const void* x, const void* g, int H, int W, int Cin, int cin_total, int c0, int Cout,
                          int cout_total, int co0, int ngroups, int group_n, const int* xoff, const int* goff, float* dwp,
                          float* db, unsigned bias_mask, pgk_stream_t stream);
extern "C" int pgk_wgrad_thin(const void* x, long long x_ps, const void* g, long long g_ps, int P, int Pr, int H, int W,
                              int Cin, int cin_total, int c0, int Cout, int ngroups, int group_n, const int* xoff,
                              const int* goff, float* dwp, float* db, unsigned bias_mask, pgk_stream_t stream);

// PGK_TC=0 in the environment (or pgk_set_tc(0)) routes every shape to the CUDA-core kernels (A/B comparisons)
static int g_tc = -1;
static bool tc_enabled() {
    if (g_tc < 0) {
        const char* e = getenv("PGK_TC");
        g_tc = (e && e[0] == '0') ? 0 : 1;
    }
    return g_tc != 0;
}
extern "C" void pgk_set_tc(int on) { g_tc = on ? 1 : 0; }

// ---- per-launch timing (off by default; bench.py's roofline leg turns it on) --------------------------------
namespace {
struct ProfRec {
    cudaEvent_t e0, e1;
    double flops, bytes;
    double products;   // bf16 tensor-core FLOPs actually issued: flops x (products of planes i + j < Pr); 0 on CUDA cores
    int family;
};
bool g_prof = false;
std::vector<ProfRec> g_recs;

struct ProfScope {
    bool on;
    cudaStream_t s;
    ProfRec r;
    ProfScope(int family, double flops, double bytes, pgk_stream_t stream, int Pr = 0)
        : on(g_prof), s((cudaStream_t)stream) {
        if (!on) return;
        r.family = family, r.flops = flops, r.bytes = bytes;
        r.products = flops * (Pr * (Pr + 1) / 2);
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, s);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(r.e1, s);
        g_recs.push_back(r);
    }
};
}  // namespace

extern "C" void pgk_prof_enable(int on) { g_prof = on != 0; }

extern "C" int pgk_prof_read(int family, double* flops, double* bytes, double* ms, long long* launches) {
    double f = 0.0, t = 0.0, b = 0.0;
    long long n = 0;
    for (auto& r : g_recs) {
        if (r.family != family) continue;
        cudaError_t e = cudaEventSynchronize(r.e1);
        float dt = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&dt, r.e0, r.e1);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_prof_read: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
        f += r.flops, b += r.bytes, t += dt, ++n;
    }
    if (bytes) *bytes = b;
    if (flops) *flops = f;
    if (ms) *ms = t;
    if (launches) *launches = n;
    return PGK_OK;
}

extern "C" int pgk_prof_read_products(int family, double* product_flops) {
    double f = 0.0;
    for (auto& r : g_recs)
        if (r.family == family) f += r.products;
    if (product_flops) *product_flops = f;
    return PGK_OK;
}

extern "C" void pgk_prof_reset(void) {
    for (auto& r : g_recs) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_recs.clear();
}

extern "C" int pgk_conv(const void* x, int P, int Pr, long long x_ps, int N, int H, int W, int Cin, int Cout, int KS,
                        int ups,
                        const float* wf, const void* wt, long long wt_ps, const float* bias, const float* posT,
                        const float* pos_s, int act, const void* mask_ref, long long mask_ps, float out_scale, void* out,
                        long long out_ps, float* pn_r, pgk_stream_t stream) {
    PGK_REQUIRE(Pr >= 1 && Pr <= P, "pgk_conv: need 1 <= Pr <= P");
    PGK_REQUIRE(!pn_r || (!mask_ref && out_scale == 1.0f), "pgk_conv: the pixel norm needs mask_ref == NULL and out_scale == 1");
    const double flops = 2.0 * N * H * W * (double)Cout * KS * KS * Cin;
    // algorithmic HBM bytes: read the planes of x that are used, write all planes of out, read one mask plane
    const double bytes = 2.0 * N * H * W * ((double)Cin * Pr / (ups ? 4 : 1) + (double)Cout * P + (mask_ref ? Cout : 0));
    int rc;
    bool pn_done = false;
    if (wt && !posT && tc_enabled() && pgk_conv_thin_supported(N, H, W, Cin, Cout, KS, ups)) {
        ProfScope prof(PGK_PROF_CONV_THIN, flops, bytes, stream, Pr);
        pn_done = pn_r && pgk_conv_thin_fuses_pixelnorm(Cout);
        rc = pgk_conv_thin(x, P, Pr, x_ps, N, H, W, Cin, Cout, wt, wt_ps, bias, act, mask_ref, mask_ps, out_scale, out,
                           out_ps, pn_done ? pn_r : nullptr, stream);
    } else if (wt && tc_enabled() && pgk_conv_tc_supported(N, H, W, Cin, Cout, KS, ups)) {
        ProfScope prof(PGK_PROF_CONV, flops, bytes, stream, Pr);
        pn_done = pn_r && !posT && pgk_conv_tc_fuses_pixelnorm(Cout, Pr == 3);
        rc = pgk_conv_tc(x, P, Pr, x_ps, N, H, W, Cin, Cout, KS, wt, wt_ps, bias, posT, pos_s, act, mask_ref, mask_ps,
                         out_scale, out, out_ps, stream, 0, 0, 1.0f, pn_done ? pn_r : nullptr);
    } else {
        PGK_REQUIRE(wf != nullptr, "pgk_conv: this shape runs on the CUDA-core kernel, which needs the fp32 operand wf");
        ProfScope prof(PGK_PROF_CONV_SIMT, flops, bytes, stream);
        rc = pgk_conv_simt(x, P, x_ps, N, H, W, Cin, Cout, KS, ups, wf, bias, posT, pos_s, act, mask_ref, mask_ps,
                           out_scale, out, out_ps, stream);
    }
    if (rc || !pn_r || pn_done) return rc;
    // kernels whose tile does not hold every channel of a pixel: the pixel norm as a second pass, in place
    return pgk_pixelnorm(out, out_ps, P, (long long)N * H * W, Cout, out, out_ps, pn_r, nullptr, 0, stream);
}

// forward convolution on IEEE-half operand planes (see include/pgk.h): the wide tensor-core kernel with fp16 A and B
// formats, three products (hi*hi, hi*lo, lo*hi) into one accumulator, scaled back by 2^-PGK_FP16_WSHIFT
extern "C" int pgk_conv_fp16(const void* xh, long long xh_ps, int N, int H, int W, int Cin, int Cout, int KS,
                             const void* wth, long long wth_ps, const float* bias, const float* posT,
                             const float* pos_s, int act, void* out, int P, long long out_ps, pgk_stream_t stream) {
    PGK_REQUIRE(P >= 2 && P <= 3, "pgk_conv_fp16: out has 2 or 3 bf16 planes (the fp32-faithful modes)");
    PGK_REQUIRE(xh && wth && out, "pgk_conv_fp16: null operand");
    PGK_REQUIRE(tc_enabled() && pgk_conv_tc_supported(N, H, W, Cin, Cout, KS, 0),
                "pgk_conv_fp16: only shapes of the wide tensor-core kernel (Cin %% 64 == 0, Cout %% 16 == 0, power-of-two H, W)");
    const double flops = 2.0 * N * H * W * (double)Cout * KS * KS * Cin;
    const double bytes = 2.0 * N * H * W * ((double)Cin * 2 + (double)Cout * P);
    ProfScope prof(PGK_PROF_CONV, flops, bytes, stream, 2);
    // the kernel reads Pr = 2 planes of x and wt (its tensor maps are declared P >= 2 planes deep) and writes P planes
    return pgk_conv_tc(xh, P, 2, xh_ps, N, H, W, Cin, Cout, KS, wth, wth_ps, bias, posT, pos_s, act, nullptr, 0, 1.0f,
                       out, out_ps, stream, 1, 1, 1.0f / (float)(1 << PGK_FP16_WSHIFT), nullptr);
}

extern "C" int pgk_wgrad(const void* x, long long x_ps, const void* g, long long g_ps, int P, int Pr, int H, int W,
                         int Cin, int Cout, int KS, int ups, int ngroups, int group_n, const int* xoff, const int* goff,
                         float* dwp, float* db, unsigned bias_groups, pgk_stream_t stream) {
    PGK_REQUIRE(ngroups >= 1 && ngroups <= 4, "pgk_wgrad: 1..4 groups");
    const double flops = 2.0 * ngroups * group_n * H * W * (double)Cout * KS * KS * Cin;
    const double bytes = 2.0 * ngroups * group_n * H * W * ((double)Cin / (ups ? 4 : 1) + Cout) * Pr;
    // one-plane mode, 64 input channels at W >= 128: the transposer-free thin kernel reads every X row once, where the
    // wide kernel re-loads a shifted box per tap (9 x the X traffic: 0.35-0.40 of the tensor peak on these shapes, bound by
    // L2 bandwidth).  Output channels in windows of 64 (PGK_WGRAD64=0: the wide kernel, for A/B runs).
    {
        static int w64 = -1;
        if (w64 < 0) {
            const char* e = getenv("PGK_WGRAD64");
            w64 = e ? atoi(e) != 0 : 1;
        }
        // (128 input channels -- the generator's first conv of a level reads the previous level's width: two windows of 64)
        const bool in_ok = Cin == 64 || (Cin == 128 && Cout <= 64);
        if (w64 && tc_enabled() && in_ok && P == 1 && Pr == 1 && KS == 3 && !ups && (Cout == 32 || Cout == 64 || Cout == 128) &&
            W % 128 == 0 && H >= 8 && H % 8 == 0 && (((uintptr_t)x | (uintptr_t)g) & 15) == 0) {
            ProfScope prof(PGK_PROF_WGRAD_THIN, flops, bytes, stream, Pr);
            const int cw = Cout == 32 ? 32 : 64;
            for (int c0 = 0; c0 < Cin; c0 += 64) {
                for (int co0 = 0; co0 < Cout; co0 += cw) {
                    int rc = pgk_wgrad_thin_direct(x, g, H, W, 64, Cin, c0, cw, Cout, co0, ngroups, group_n, xoff, goff, dwp,
                                                   c0 == 0 ? db : nullptr, bias_groups, stream);
                    if (rc) return rc;
                    PGK_LAUNCH_CHECK("pgk_wgrad(thin tcgen05, direct, 64-channel windows)");
                }
            }
            return PGK_OK;
        }
    }
    // thin layers: the bias gradient rides along as one more accumulator row (a 64-channel input = two launches)
    const int thin_cin = Cin == 64 ? 32 : Cin;
    if (tc_enabled() && (Cin != 64 || Cout < 64) &&
        pgk_wgrad_thin_supported(H, W, thin_cin, Cout, KS, ups, ngroups, group_n, Pr)) {
        ProfScope prof(PGK_PROF_WGRAD_THIN, flops, bytes, stream, Pr);
        for (int c0 = 0; c0 < Cin; c0 += thin_cin) {
            int rc = pgk_wgrad_thin(x, x_ps, g, g_ps, P, Pr, H, W, thin_cin, Cin, c0, Cout, ngroups, group_n, xoff, goff,
                                    dwp, c0 == 0 ? db : nullptr, bias_groups, stream);
            if (rc) return rc;
        }
        return PGK_OK;
    }
    int rc;
    if (tc_enabled() && pgk_wgrad_tc_supported(H, W, Cin, Cout, KS, ups, ngroups, group_n)) {
        ProfScope prof(PGK_PROF_WGRAD, flops, bytes, stream, Pr);
        // the bias gradient rides along in the same launch (PGK_WGRAD_BIAS_FUSE=0: the separate pass below, for A/B runs)
        static int fuse = -1;
        if (fuse < 0) {
            const char* e = getenv("PGK_WGRAD_BIAS_FUSE");
            fuse = e ? atoi(e) != 0 : 1;
        }
        int fused = 0;
        rc = pgk_wgrad_tc(x, x_ps, g, g_ps, P, Pr, H, W, Cin, Cout, KS, ngroups, group_n, xoff, goff, dwp,
                          (fuse && P == Pr) ? db : nullptr, bias_groups, &fused, stream);
        if (rc || fused) return rc;
    } else {
        ProfScope prof(PGK_PROF_WGRAD_SIMT, flops, bytes, stream);
        rc = pgk_wgrad_simt(x, x_ps, g, g_ps, P, H, W, Cin, Cout, KS, ups, ngroups, group_n, xoff, goff, dwp, stream);
    }
    if (rc || !db) return rc;
    int boff[4], nb = 0;
    for (int i = 0; i < ngroups; ++i)
        if ((bias_groups >> i) & 1) boff[nb++] = goff[i];
    if (nb == 0) return PGK_OK;
    return pgk_bias_grad(g, g_ps, P, H * W, Cout, nb, group_n, boff, 1.0f, db, 1, stream);
}

// pgk_api.cu -- shape dispatch for the two GEMM-shaped entry points.  The tcgen05 path (pgk_conv_tc.cu) serves
// the tensor-core friendly layers; everything else runs on the CUDA-core implicit GEMM.  Both are this library's
// own sm_100a kernels -- there is no library or CPU fallback.
#include <stdlib.h>

#include "pgk_common.cuh"

extern "C" int pgk_conv_simt(const void* x, int P, long long x_ps, int N, int H, int W, int Cin, int Cout, int KS,
                             int ups, const float* wf, const float* bias, const float* posT, const float* pos_s,
                             int act, const void* mask_ref, long long mask_ps, float out_scale, void* out,
                             long long out_ps, pgk_stream_t stream);
extern "C" int pgk_wgrad_simt(const void* x, long long x_ps, const void* g, long long g_ps, int P, int H, int W,
                              int Cin, int Cout, int KS, int ups, int ngroups, int group_n, const int* xoff,
                              const int* goff, float* dwp, pgk_stream_t stream);

extern "C" int pgk_conv(const void* x, int P, long long x_ps, int N, int H, int W, int Cin, int Cout, int KS, int ups,
                        const float* wf, const float* bias, const float* posT, const float* pos_s, int act,
                        const void* mask_ref, long long mask_ps, float out_scale, void* out, long long out_ps,
                        pgk_stream_t stream) {
    return pgk_conv_simt(x, P, x_ps, N, H, W, Cin, Cout, KS, ups, wf, bias, posT, pos_s, act, mask_ref, mask_ps,
                         out_scale, out, out_ps, stream);
}

extern "C" int pgk_wgrad(const void* x, long long x_ps, const void* g, long long g_ps, int P, int H, int W, int Cin,
                         int Cout, int KS, int ups, int ngroups, int group_n, const int* xoff, const int* goff,
                         float* dwp, pgk_stream_t stream) {
    return pgk_wgrad_simt(x, x_ps, g, g_ps, P, H, W, Cin, Cout, KS, ups, ngroups, group_n, xoff, goff, dwp, stream);
}

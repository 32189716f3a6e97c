// pgk_relayout.cuh -- index logic of the weight re-layout (pgk_prep_weight / pgk_unprep_grad).
//
// PyTorch keeps a conv weight as w[co][ci][ky][kx] (network.py:16).  The kernels read
//   wf[fi]  forward operand  [K][Cout]  (output channel fastest)
//   wb[bi]  data-gradient operand, taps flipped, channel roles swapped (input channel fastest)
// so the re-layout is a transpose between co and (ci, taps).
#pragma once
#include "../../include/pgk.h"

#if defined(__CUDACC__)
#define PGK_HD __host__ __device__ __forceinline__
#else
#define PGK_HD inline
#endif

// element (co, ci, ky, kx) of the PyTorch weight -> its index in wf (fi) and in wb (bi)
PGK_HD void weight_index(int kind, int cin, int cout, int ks, int co, int ci, int ky, int kx, long long& fi,
                         long long& bi) {
    if (kind == PGK_W_CONV) {
        int tap = ky * ks + kx, tapf = (ks - 1 - ky) * ks + (ks - 1 - kx);
        fi = ((long long)tap * cin + ci) * cout + co;
        bi = ((long long)tapf * cout + co) * cin + ci;
    } else if (kind == PGK_W_GFIRST) {  // out pixel (y,x) = (3-ky, 3-kx)
        int p = (3 - ky) * 4 + (3 - kx);
        fi = (long long)ci * (16 * cout) + (long long)p * cout + co;
        bi = ((long long)p * cout + co) * cin + ci;
    } else {  // PGK_W_DLAST: in pixel (y,x) = (ky,kx)
        int p = ky * 4 + kx;
        fi = ((long long)p * cin + ci) * cout + co;
        bi = (long long)co * (16 * cin) + (long long)p * cin + ci;
    }
}

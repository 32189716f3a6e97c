// pgk_relayout.cuh -- index logic of the weight re-layout (pgk_prep_weight / pgk_unprep_grad), written so that the
// same functions run on the device and, thread by thread, on the host (tests/relayout_host_check.cpp emulates the
// kernels' two phases with them and compares every element with the per-element mapping below).
//
// PyTorch keeps a conv weight as w[co][ci][ky][kx] (network.py:16).  The kernels read
//   wf[fi]  forward operand  [K][Cout]  (output channel fastest)
//   wb[bi]  data-gradient operand, taps flipped, channel roles swapped (input channel fastest)
// so the re-layout is a transpose between co and (ci, taps): one thread per element (the first version) has either
// its reads or its writes 4 bytes apart in different 32-byte sectors.  The tiled version moves a tile of
// kTileCo output channels x TCI input channels x all taps through shared memory: rows of TCI * taps contiguous
// floats on the PyTorch side, 128-byte (64-byte for the 4x4 layers) runs of co / ci on the operand side.
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

constexpr int kTileCo = 32;

// one tile of the tiled re-layout: output channels [co0, co0 + kTileCo) x input channels [ci0, ci0 + tci) x taps
struct RelayoutTile {
    int kind, cin, cin_stride, cout, ks, taps;
    int tci;      // input channels per tile: 32 (ks <= 3) or 16 (ks = 4, to stay inside 48 KB of shared memory)
    int tp;       // shared-memory stride of one input channel: taps | 1 (odd: ci-fastest reads are conflict free)
    int row;      // shared-memory stride of one output channel: tci * tp rounded to 1 mod 32 (co-fastest reads ditto)
    int co0, ci0;
};

PGK_HD int relayout_tci(int ks) { return ks == 4 ? 16 : 32; }
PGK_HD int relayout_tile_floats(int ks) {
    const int taps = ks * ks, tp = taps | 1, body = relayout_tci(ks) * tp;
    return kTileCo * (((body + 31) / 32) * 32 + 1);
}
PGK_HD RelayoutTile relayout_tile(int kind, int cin, int cin_stride, int cout, int ks, int bx, int by) {
    RelayoutTile t;
    t.kind = kind, t.cin = cin, t.cin_stride = cin_stride, t.cout = cout, t.ks = ks, t.taps = ks * ks;
    t.tci = relayout_tci(ks);
    t.tp = t.taps | 1;
    t.row = ((t.tci * t.tp + 31) / 32) * 32 + 1;
    t.ci0 = bx * t.tci, t.co0 = by * kTileCo;
    return t;
}

// phase 1 of pgk_prep_weight: tile[co_l][ci_l * tp + tap] = c * w[co][ci][tap]; consecutive threads read consecutive
// floats of the tci * taps run that one output channel contributes
PGK_HD void relayout_load_w(float* tile, const float* w, float c, const RelayoutTile& t, int tid, int nthreads) {
    const int run = t.tci * t.taps;
    for (int idx = tid; idx < kTileCo * run; idx += nthreads) {
        const int co_l = idx / run, rem = idx - co_l * run;
        const int ci_l = rem / t.taps, tap = rem - ci_l * t.taps;
        const int co = t.co0 + co_l, ci = t.ci0 + ci_l;
        if (co < t.cout && ci < t.cin)
            tile[co_l * t.row + ci_l * t.tp + tap] = c * w[((long long)co * t.cin_stride + ci) * t.taps + tap];
    }
}

// phase 2 of pgk_prep_weight: wf with the output channel fastest across threads, wb with the input channel fastest
PGK_HD void relayout_store_fb(const float* tile, float* wf, float* wb, const RelayoutTile& t, int tid, int nthreads) {
    if (wf) {
        for (int idx = tid; idx < kTileCo * t.tci * t.taps; idx += nthreads) {
            const int co_l = idx % kTileCo, r = idx / kTileCo;
            const int ci_l = r % t.tci, tap = r / t.tci;
            const int co = t.co0 + co_l, ci = t.ci0 + ci_l;
            if (co < t.cout && ci < t.cin) {
                long long fi, bi;
                weight_index(t.kind, t.cin, t.cout, t.ks, co, ci, tap / t.ks, tap % t.ks, fi, bi);
                wf[fi] = tile[co_l * t.row + ci_l * t.tp + tap];
            }
        }
    }
    if (wb) {
        for (int idx = tid; idx < kTileCo * t.tci * t.taps; idx += nthreads) {
            const int ci_l = idx % t.tci, r = idx / t.tci;
            const int co_l = r % kTileCo, tap = r / kTileCo;
            const int co = t.co0 + co_l, ci = t.ci0 + ci_l;
            if (co < t.cout && ci < t.cin) {
                long long fi, bi;
                weight_index(t.kind, t.cin, t.cout, t.ks, co, ci, tap / t.ks, tap % t.ks, fi, bi);
                wb[bi] = tile[co_l * t.row + ci_l * t.tp + tap];
            }
        }
    }
}

// phase 1 of pgk_unprep_grad: tile <- dwp (wf layout), output channel fastest across threads
PGK_HD void relayout_load_dwp(float* tile, const float* dwp, const RelayoutTile& t, int tid, int nthreads) {
    for (int idx = tid; idx < kTileCo * t.tci * t.taps; idx += nthreads) {
        const int co_l = idx % kTileCo, r = idx / kTileCo;
        const int ci_l = r % t.tci, tap = r / t.tci;
        const int co = t.co0 + co_l, ci = t.ci0 + ci_l;
        if (co < t.cout && ci < t.cin) {
            long long fi, bi;
            weight_index(t.kind, t.cin, t.cout, t.ks, co, ci, tap / t.ks, tap % t.ks, fi, bi);
            tile[co_l * t.row + ci_l * t.tp + tap] = dwp[fi];
        }
    }
}

// phase 2 of pgk_unprep_grad: dw[co][ci][tap] (PyTorch layout) = (+=) c * tile, contiguous runs per output channel
PGK_HD void relayout_store_dw(const float* tile, float* dw, float c, int accumulate, const RelayoutTile& t, int tid,
                              int nthreads) {
    const int run = t.tci * t.taps;
    for (int idx = tid; idx < kTileCo * run; idx += nthreads) {
        const int co_l = idx / run, rem = idx - co_l * run;
        const int ci_l = rem / t.taps, tap = rem - ci_l * t.taps;
        const int co = t.co0 + co_l, ci = t.ci0 + ci_l;
        if (co < t.cout && ci < t.cin) {
            const long long o = ((long long)co * t.cin_stride + ci) * t.taps + tap;
            const float v = c * tile[co_l * t.row + ci_l * t.tp + tap];
            dw[o] = accumulate ? dw[o] + v : v;
        }
    }
}

// pgk_prep.cu -- every kernel-side operand of the equalised-LR convs of ONE network in ONE launch (pgk_prep_multi).
//
// After each optimizer step every active conv layer needs its weight re-laid: c folded in (network.py:33 scales the
// input; folding it into the weight is the same product), the forward operand [Cout][K] and the tap-flipped /
// transposed data-gradient operand [Cin][K'] as K-major bf16 planes for the tcgen05 kernels (or the thin-layer
// packing of pgk_conv_thin.cu), optionally the fp32 [K][Cout] operands of the CUDA-core kernels and the IEEE-half
// packing of the forward operand.  Layer by layer that was three to five small launches per layer (pgk_prep_weight,
// pgk_pack_operand x2, pgk_pack_thin x2): ~100 launches and 9 % of a depth-8 / batch-4 iteration, each kernel either
// reading or writing 4 bytes per 32-byte sector.  Here one CTA moves a tile of 32 output channels x 32 (16 for the 4x4
// layers) input channels x all taps through shared memory: the PyTorch side is read in runs of 1152 bytes, every
// operand is written in runs of 64 bytes or more (ci-fastest for the forward operand, co-fastest for the data-gradient
// one), and all layers of the table share the launch.
#include "pgk_common.cuh"
#include "pgk_relayout.cuh"

#include <cuda_fp16.h>

namespace {

constexpr int kTileCo = 32;
constexpr int kMaxLayers = 24;

struct PrepTable {
    PgkPrepLayer l[kMaxLayers];
    int tile0[kMaxLayers + 1];   // first tile of each layer; tile0[n] = number of tiles
    int n;
};

__device__ __forceinline__ int tci_of(int ks) { return ks == 4 ? 16 : 32; }

// index of (k, n) in pgk_pack_thin's layout: out[step][khalf][n][e] (csrc/pgk_conv_thin.cu pack_thin_kernel)
__device__ __forceinline__ long long thin_index(int k, int n, int Cin, int Npad) {
    int st, h, e;
    if (Cin == 8) {
        const int tap = k >> 3, dy = tap / 3, dx = tap - dy * 3;
        st = dy * 2 + (dx >> 1), h = dx & 1, e = k & 7;
    } else {
        const int tap = k / Cin, r = k - tap * Cin;
        st = tap * (Cin / 16) + (r >> 4), h = (r >> 3) & 1, e = r & 7;
    }
    return ((long long)(st * 2 + h) * Npad + n) * 8 + e;
}

__device__ __forceinline__ void store_planes(bf16* out, long long ps, int P, long long idx, float v) {
    for (int p = 0; p < P; ++p) {
        const bf16 h = __float2bfloat16_rn(v);
        out[p * ps + idx] = h;
        v -= __bfloat162float(h);
    }
}

__global__ void __launch_bounds__(256) prep_multi_kernel(const __grid_constant__ PrepTable t) {
    extern __shared__ float tile[];
    pgk_pdl_enter();
    // which layer does this tile belong to?
    int li = 0;
    while (li + 1 < t.n && (int)blockIdx.x >= t.tile0[li + 1]) ++li;
    const PgkPrepLayer& L = t.l[li];
    const int taps = L.ks * L.ks, tci = tci_of(L.ks), tp = taps | 1;
    const int row = ((tci * tp) | 1);                 // odd strides: both store orders read shared memory conflict free
    const int tiles_ci = (L.cin + tci - 1) / tci;
    const int tl = (int)blockIdx.x - t.tile0[li];
    const int co0 = (tl / tiles_ci) * kTileCo, ci0 = (tl % tiles_ci) * tci;
    const int nco = min(kTileCo, L.cout - co0), nci = min(tci, L.cin - ci0);
    // ---- load: for every output channel of the tile one contiguous run of nci * taps floats
    // (element r of a run sits at ci_l * tp + tap = r + ci_l * (tp - taps): the padding is 0 for 1 / 9 taps, 1 for 16)
    const int run = nci * taps;
    for (int co_l = threadIdx.x >> 5; co_l < nco; co_l += blockDim.x >> 5) {
        const float* src = L.w + ((long long)(co0 + co_l) * L.cin_stride + ci0) * taps;
        for (int r = threadIdx.x & 31; r < run; r += 32)
            tile[co_l * row + r + (taps == 16 ? (r >> 4) : 0)] = L.c * src[r];
    }
    __syncthreads();
    // operand geometry (engine.ConvW): forward [nf][kf], data gradient [nb][kb]
    int kf, kb;
    if (L.kind == PGK_W_CONV) kf = taps * L.cin, kb = taps * L.cout;
    else if (L.kind == PGK_W_GFIRST) kf = L.cin, kb = 16 * L.cout;
    else kf = 16 * L.cin, kb = L.cout;
    const int npad_f = L.cout < 16 ? 16 : L.cout, npad_b = L.cin < 16 ? 16 : L.cin;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // ---- pass 1, a warp per (output channel, tap) row, lanes = input channels: forward operand (+ half packing), wb.
    // Every destination is contiguous in ci: per row one base index, no division per element.
    // (a warp walks the taps of its output channels: no division per row -- the per-row index arithmetic was the larger
    // part of this kernel's instructions)
    for (int co_l = warp; co_l < nco; co_l += nwarps)
    for (int tap = 0; tap < taps; ++tap) {
        const int co = co0 + co_l;
        long long fbase, bbase;      // index of (co, ci0, tap) in F (standard layout) and in wb
        if (L.kind == PGK_W_CONV) {
            const int tapf = taps - 1 - tap;
            fbase = (long long)co * kf + (long long)tap * L.cin + ci0;
            bbase = ((long long)tapf * L.cout + co) * L.cin + ci0;
        } else if (L.kind == PGK_W_GFIRST) {
            fbase = ((long long)(15 - tap) * L.cout + co) * L.cin + ci0;
            bbase = fbase;
        } else {
            fbase = (long long)co * kf + (long long)tap * L.cin + ci0;
            bbase = fbase;
        }
        if (lane < nci) {
            const float v = tile[co_l * row + lane * tp + tap];
            if (L.F) {
                const long long idx = L.thinF ? thin_index(tap * L.cin + ci0 + lane, co, L.cin, npad_f) : fbase + lane;
                store_planes((bf16*)L.F, L.F_ps, L.planes, idx, v);
            }
            if (L.F16) {
                __half* o = (__half*)L.F16;
                float sv = v * (float)(1 << PGK_FP16_WSHIFT);
                for (int p = 0; p < 2; ++p) {
                    const __half h = __float2half_rn(sv);
                    o[p * L.F16_ps + fbase + lane] = h;
                    sv -= __half2float(h);
                }
            }
            if (L.wb) L.wb[bbase + lane] = v;
        }
    }
    // ---- pass 2, a warp per (input channel, tap) row, lanes = output channels: data-gradient operand, wf
    if (L.B || L.wf) {
        for (int ci_l = warp; ci_l < nci; ci_l += nwarps)
        for (int tap = 0; tap < taps; ++tap) {
            const int ci = ci0 + ci_l;
            long long fibase, bidx;     // index of (co0, ci, tap) in wf and in B (standard layout)
            if (L.kind == PGK_W_CONV) {
                const int tapf = taps - 1 - tap;
                fibase = ((long long)tap * L.cin + ci) * L.cout + co0;
                bidx = (long long)ci * kb + (long long)tapf * L.cout + co0;
            } else if (L.kind == PGK_W_GFIRST) {
                fibase = (long long)ci * (16 * L.cout) + (long long)(15 - tap) * L.cout + co0;
                bidx = 0;               // (the first generator layer has no data-gradient operand)
            } else {
                fibase = ((long long)tap * L.cin + ci) * L.cout + co0;
                bidx = fibase;
            }
            if (lane < nco) {
                const float v = tile[lane * row + ci_l * tp + tap];
                if (L.B) {
                    const long long idx = L.thinB ? thin_index((taps - 1 - tap) * L.cout + co0 + lane, ci, L.cout, npad_b)
                                                  : bidx + lane;
                    store_planes((bf16*)L.B, L.B_ps, L.planes, idx, v);
                }
                if (L.wf) L.wf[fibase + lane] = v;
            }
        }
    }
}

// ---- the inverse map for weight gradients, every layer of a network in one launch -------------------------------------
struct UnprepTable {
    PgkUnprepLayer l[kMaxLayers * 2];
    int tile0[kMaxLayers * 2 + 1];
    int n;
};

__global__ void __launch_bounds__(256) unprep_multi_kernel(const __grid_constant__ UnprepTable t) {
    extern __shared__ float tile[];
    pgk_pdl_enter();
    int li = 0;
    while (li + 1 < t.n && (int)blockIdx.x >= t.tile0[li + 1]) ++li;
    const PgkUnprepLayer& L = t.l[li];
    const int taps = L.ks * L.ks, tci = tci_of(L.ks), tp = taps | 1;
    const int row = ((tci * tp) | 1);
    const int tiles_ci = (L.cin + tci - 1) / tci;
    const int tl = (int)blockIdx.x - t.tile0[li];
    const int co0 = (tl / tiles_ci) * kTileCo, ci0 = (tl % tiles_ci) * tci;
    const int nco = min(kTileCo, L.cout - co0), nci = min(tci, L.cin - ci0);
    // ---- load, a warp per (input channel, tap) row, lanes = output channels: dwp is [K][Cout] (wf's layout)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int ci_l = warp; ci_l < nci; ci_l += nwarps)
    for (int tap = 0; tap < taps; ++tap) {
        const int ci = ci0 + ci_l;
        long long fibase;
        if (L.kind == PGK_W_GFIRST) fibase = (long long)ci * (16 * L.cout) + (long long)(15 - tap) * L.cout + co0;
        else fibase = ((long long)tap * L.cin + ci) * L.cout + co0;
        if (lane < nco) tile[lane * row + ci_l * tp + tap] = L.c * L.dwp[fibase + lane];
    }
    __syncthreads();
    // ---- store: for every output channel one contiguous run of nci * taps floats of the PyTorch layout
    const int run = nci * taps;
    for (int co_l = warp; co_l < nco; co_l += nwarps) {
        float* dst = L.dw + ((long long)(co0 + co_l) * L.cin_stride + ci0) * taps;
        for (int r = lane; r < run; r += 32) {
            const float v = tile[co_l * row + r + (taps == 16 ? (r >> 4) : 0)];
            dst[r] = L.accumulate ? dst[r] + v : v;
        }
    }
}

}  // namespace

extern "C" int pgk_unprep_multi(const PgkUnprepLayer* layers, int n, pgk_stream_t stream) {
    PGK_REQUIRE(layers && n > 0, "pgk_unprep_multi: empty table");
    const int smem_max = (int)sizeof(float) * kTileCo * ((32 * 9) | 1);
    for (int base = 0; base < n; base += kMaxLayers * 2) {
        UnprepTable t;
        t.n = n - base < kMaxLayers * 2 ? n - base : kMaxLayers * 2;
        int tiles = 0;
        for (int i = 0; i < t.n; ++i) {
            const PgkUnprepLayer& L = layers[base + i];
            PGK_REQUIRE(L.dwp && L.dw && L.kind >= 0 && L.kind <= 2 && L.cin > 0 && L.cout > 0 && L.cin_stride >= L.cin,
                        "pgk_unprep_multi: bad layer %d", base + i);
            PGK_REQUIRE(L.kind == PGK_W_CONV ? (L.ks == 1 || L.ks == 3) : L.ks == 4, "pgk_unprep_multi: bad ks %d for kind %d",
                        L.ks, L.kind);
            t.l[i] = L;
            t.tile0[i] = tiles;
            const int tci = L.ks == 4 ? 16 : 32;
            tiles += ((L.cout + kTileCo - 1) / kTileCo) * ((L.cin + tci - 1) / tci);
        }
        t.tile0[t.n] = tiles;
        pgk_launch(unprep_multi_kernel, dim3((unsigned)tiles), dim3(256), (size_t)smem_max, (cudaStream_t)stream, t);
        PGK_LAUNCH_CHECK("pgk_unprep_multi");
    }
    return PGK_OK;
}

extern "C" int pgk_prep_multi(const PgkPrepLayer* layers, int n, pgk_stream_t stream) {
    PGK_REQUIRE(layers && n > 0, "pgk_prep_multi: empty table");
    static bool attr = false;
    const int smem_max = (int)sizeof(float) * kTileCo * ((32 * 9) | 1);
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(prep_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_prep_multi: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
        attr = true;
    }
    for (int base = 0; base < n; base += kMaxLayers) {
        PrepTable t;
        t.n = n - base < kMaxLayers ? n - base : kMaxLayers;
        int tiles = 0;
        for (int i = 0; i < t.n; ++i) {
            const PgkPrepLayer& L = layers[base + i];
            PGK_REQUIRE(L.w && L.kind >= 0 && L.kind <= 2 && L.cin > 0 && L.cout > 0 && L.cin_stride >= L.cin,
                        "pgk_prep_multi: bad layer %d", base + i);
            PGK_REQUIRE(L.kind == PGK_W_CONV ? (L.ks == 1 || L.ks == 3) : L.ks == 4, "pgk_prep_multi: bad ks %d for kind %d",
                        L.ks, L.kind);
            PGK_REQUIRE(L.planes >= 1 && L.planes <= 3, "pgk_prep_multi: planes must be 1..3");
            PGK_REQUIRE(!(L.thinF || L.thinB) || (L.kind == PGK_W_CONV && L.ks == 3), "pgk_prep_multi: thin packing is for 3x3 layers");
            PGK_REQUIRE(!L.thinF || ((L.cin == 8 || L.cin == 16 || L.cin == 32) && L.cout % 8 == 0 && L.cout <= 64),
                        "pgk_prep_multi: thin forward packing needs Cin in {8,16,32}, Cout <= 64");
            PGK_REQUIRE(!L.thinB || ((L.cout == 8 || L.cout == 16 || L.cout == 32) && L.cin % 8 == 0 && L.cin <= 64),
                        "pgk_prep_multi: thin data-gradient packing needs Cout in {8,16,32}, Cin <= 64");
            t.l[i] = L;
            t.tile0[i] = tiles;
            const int tci = L.ks == 4 ? 16 : 32;
            tiles += ((L.cout + kTileCo - 1) / kTileCo) * ((L.cin + tci - 1) / tci);
        }
        t.tile0[t.n] = tiles;
        pgk_launch(prep_multi_kernel, dim3((unsigned)tiles), dim3(256), (size_t)smem_max, (cudaStream_t)stream, t);
        PGK_LAUNCH_CHECK("pgk_prep_multi");
    }
    return PGK_OK;
}

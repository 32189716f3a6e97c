// pgk_conv_tc.cu -- the tensor-core path of the two GEMM-shaped operations (sm_100a: TMA -> shared memory ->
// tcgen05.mma -> TMEM -> fused epilogue).
//
//  conv_tc_kernel   out[pix][co] = E( sum_{tap,ci} X[pix + tap][ci] * Wt[co][tap*Cin + ci] )         (network.py:34)
//      im2col-free: for every (tap, 64-channel slice) ONE TMA box load of the shifted [TN x TH x TW] pixel block;
//      TMA's out-of-bounds zero fill IS the conv's zero padding.  A (pixels x K) and B (Cout x K) are K-major,
//      128-byte swizzled; the accumulator tile 128 pixels x NT channels lives in TMEM.
//  wgrad_tc_kernel  dW[tap*Cin + ci][co] += sum_pix X[pix + tap][ci] * G[pix][co]    (cuDNN convolution_backward, weight)
//      the same boxes, read as MN-major operands (the reduction runs over pixels); split over pixel ranges, fp32
//      atomics into the [K][Cout] gradient.
//
// Planes (see include/pgk.h): with P planes per operand the kernels issue the products of planes (i, j), i + j < P,
// into the same fp32 accumulator: 1 product (bf16), 3 (16 mantissa bits) or 6 (24 bits, the fp32-faithful mode).
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 TMA producer, warp 5 TMEM
// allocation + MMA issue.  smem ring of `stages` {A planes, B planes}; full/empty mbarriers; tcgen05.commit releases
// a stage back to the producer and, after the last k-step, hands the accumulator to the epilogue.
#include "pgk_tc.cuh"

using namespace tc;

// ------------------------------------------------------------------------------------------------------------
// host helpers shared by both kernels
// ------------------------------------------------------------------------------------------------------------
pgk_encode_tiled_fn pgk_get_encode_tiled() {
    static pgk_encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        pgk_set_error("cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
        return nullptr;
    }
    fn = (pgk_encode_tiled_fn)p;
    return fn;
}

int pgk_make_tmap(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
                  const unsigned long long* strides, const unsigned* box, int swizzle_bytes, const char* what) {
    pgk_encode_tiled_fn enc = pgk_get_encode_tiled();
    if (!enc) return PGK_ERR_CUDA;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = 1;
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
    CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        pgk_set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu %llu %llu box %u %u %u)", what,
                      (int)r, rank, dims[0], dims[1], rank > 2 ? dims[2] : 0ull, box[0], box[1], rank > 2 ? box[2] : 0u);
        return PGK_ERR_CUDA;
    }
    return PGK_OK;
}

namespace {

constexpr int kThreads = 192;
constexpr int kSmemLimit = 227 * 1024;

inline int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
__host__ __device__ inline unsigned tmem_cols(int n) {
    unsigned c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

// ------------------------------------------------------------------------------------------------------------
// forward conv / data gradient
// ------------------------------------------------------------------------------------------------------------
struct ConvTcArgs {
    int N, H, W, Cin, Cout, KS, P;
    int lw, lh, TN;        // pixel block = TN samples x 2^lh rows x 2^lw columns = 128 pixels
    int tiles_x, tiles_y;
    int NT, stages, bkb;   // channels per CTA, ring depth, bytes per K row of a stage (128 or 64)
    const float* bias;
    const float* posT;
    const float* pos_s;
    int act, has_mask;
    Planes mask;
    float out_scale;
    Planes out;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    const int P = a.P;
    const uint32_t a_bytes = 128u * a.bkb, b_bytes = (uint32_t)a.NT * a.bkb;
    const uint32_t stage_bytes = P * (a_bytes + b_bytes);
    const uint32_t bars = sbase + a.stages * stage_bytes;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (a.stages + s); };
    const uint32_t tfull = bars + 16u * a.stages, tptr = tfull + 8u;

    int t = blockIdx.x;
    const int tx = t % a.tiles_x;
    t /= a.tiles_x;
    const int ty = t % a.tiles_y;
    const int tn = t / a.tiles_y;
    const int x0 = tx << a.lw, y0 = ty << a.lh, n0 = tn * a.TN;
    const int co0 = blockIdx.y * a.NT;
    const int kelems = a.bkb >> 1;
    const int kchunks = a.Cin / kelems;
    const int nk = a.KS * a.KS * kchunks;
    const int pad = a.KS >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    // With P > 1 the plane-0 x plane-0 products go to one accumulator and all correction products (2^-8 and smaller)
    // to a second one, added in the epilogue: the tensor core truncates the accumulator on every add, so keeping
    // the adds into the large accumulator to K/16 (instead of 3x / 6x that) is what holds fp32-level accuracy.
    const unsigned ncols = tmem_cols(P > 1 ? 2 * a.NT : a.NT);
    if (warp == 5) tmem_alloc(tptr, ncols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));

    if (warp == 4) {
        if (lane == 0) {
            for (int kc = 0; kc < nk; ++kc) {
                const int s = kc % a.stages;
                mbar_wait(empty(s), ((kc / a.stages) & 1) ^ 1);
                mbar_expect_tx(full(s), stage_bytes);
                const int tap = kc / kchunks, cc = kc - tap * kchunks;
                const int ky = tap / a.KS, kx = tap - ky * a.KS;
                const uint32_t dst = sbase + s * stage_bytes;
                for (int p = 0; p < P; ++p)
                    tma_load_5d(dst + p * a_bytes, &tmA, full(s), cc * kelems, x0 + kx - pad, y0 + ky - pad, n0, p);
                for (int p = 0; p < P; ++p)
                    tma_load_3d(dst + P * a_bytes + p * b_bytes, &tmB, full(s), kc * kelems, co0, p);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16(a.NT, 0, 0);
            const uint32_t layout = a.bkb == 128 ? 2u : 4u;
            const uint32_t sbo = 8u * a.bkb;
            const int ksteps = a.bkb / 32;
            uint32_t acc = 0, acc_corr = 0;
            for (int kc = 0; kc < nk; ++kc) {
                const int s = kc % a.stages;
                mbar_wait(full(s), (kc / a.stages) & 1);
                fence_after();
                const uint32_t abase = sbase + s * stage_bytes, bbase = abase + P * a_bytes;
                for (int ks = 0; ks < ksteps; ++ks) {
                    for (int i = 0; i < P; ++i) {
                        const uint64_t ad = smem_desc(abase + i * a_bytes + ks * 32, 16, sbo, layout);
                        for (int j = 0; i + j < P; ++j) {
                            const uint64_t bd = smem_desc(bbase + j * b_bytes + ks * 32, 16, sbo, layout);
                            if (i + j == 0) {
                                mma_bf16(tmem, ad, bd, idesc, acc);
                                acc = 1;
                            } else {
                                mma_bf16(tmem + a.NT, ad, bd, idesc, acc_corr);
                                acc_corr = 1;
                            }
                        }
                    }
                }
                mma_commit(empty(s));
            }
            mma_commit(tfull);
        }
        __syncwarp();
    } else {
        // ---- epilogue: one accumulator row (= one pixel) per thread
        mbar_wait(tfull, 0);
        fence_after();
        const int r = warp * 32 + lane;
        const int px = r & ((1 << a.lw) - 1);
        const int py = (r >> a.lw) & ((1 << a.lh) - 1);
        const int n = n0 + (r >> (a.lw + a.lh));
        const int x = x0 + px, y = y0 + py;
        const bool valid = n < a.N;
        const long long pix = ((long long)n * a.H + y) * a.W + x;
        const float ps = (a.posT && valid) ? __ldg(a.pos_s + n) : 0.f;
        const float* posrow = a.posT ? a.posT + (long long)(y * a.W + x) * a.Cout : nullptr;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < a.NT; c += 16) {
            float v[16];
            tmem_ld16(trow + c, v);
            if (P > 1) {
                float w[16];
                tmem_ld16(trow + a.NT + c, w);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += w[j];
            }
            if (!valid) continue;
            const int co = co0 + c;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float* f = v + 8 * h;
                const int cc = co + 8 * h;
                if (a.bias) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + cc));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias + cc + 4));
                    f[0] += b0.x, f[1] += b0.y, f[2] += b0.z, f[3] += b0.w;
                    f[4] += b1.x, f[5] += b1.y, f[6] += b1.z, f[7] += b1.w;
                }
                if (posrow) {
                    const float4 p0 = __ldg(reinterpret_cast<const float4*>(posrow + cc));
                    const float4 p1 = __ldg(reinterpret_cast<const float4*>(posrow + cc + 4));
                    f[0] = fmaf(ps, p0.x, f[0]), f[1] = fmaf(ps, p0.y, f[1]), f[2] = fmaf(ps, p0.z, f[2]);
                    f[3] = fmaf(ps, p0.w, f[3]), f[4] = fmaf(ps, p1.x, f[4]), f[5] = fmaf(ps, p1.y, f[5]);
                    f[6] = fmaf(ps, p1.z, f[6]), f[7] = fmaf(ps, p1.w, f[7]);
                }
                if (a.act) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = lrelu(f[j]);
                }
                const long long o = pix * a.Cout + cc;
                if (a.has_mask) {
                    float m[8];
                    ld8(a.mask, o, m);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] *= lrelu_grad(m[j]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] *= a.out_scale;
                split_store8(a.out, o, f);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------------------
struct WgradTcArgs {
    int H, W, Cin, Cout, KS, P;
    int lw, lh, TN;       // pixel block of one stage: TN samples x 2^lh rows x 2^lw columns = PXS pixels
    int PXS, tiles_x, tiles_per_sample;   // tiles_per_sample = 0 when one block spans TN > 1 samples
    int ngroups, group_n;
    int xoff[4], goff[4];
    int NT, S, stages;    // channels per CTA, 128-row slabs of the [K][Cout] gradient per CTA, ring depth
    int RG;               // 64-row groups in K = KS*KS*Cin / 64
    long long tiles_per_group, tiles_total, tiles_per_cta;
    float* dwp;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                const WgradTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    const int P = a.P, S = a.S;
    const uint32_t box_bytes = (uint32_t)a.PXS * 128u;
    const int gboxes = a.NT / 64;
    const uint32_t plane_bytes = (2 * S + gboxes) * box_bytes;   // [2S boxes of X | NT/64 boxes of G]
    const uint32_t stage_bytes = P * plane_bytes;
    const uint32_t bars = sbase + a.stages * stage_bytes;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (a.stages + s); };
    const uint32_t tfull = bars + 16u * a.stages, tptr = tfull + 8u;

    const int slab0 = blockIdx.x * S;          // first 128-row slab of this CTA
    const int co0 = blockIdx.y * a.NT;
    const long long t_begin = (long long)blockIdx.z * a.tiles_per_cta;
    long long t_end = t_begin + a.tiles_per_cta;
    if (t_end > a.tiles_total) t_end = a.tiles_total;
    const int ntiles = (int)(t_end - t_begin);   // >= 1 by construction of the grid
    const int pad = a.KS >> 1;
    const int cchunks = a.Cin / 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmG);
    }
    const unsigned ncols = tmem_cols(S * a.NT);
    if (warp == 5) tmem_alloc(tptr, ncols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));

    if (warp == 4) {
        if (lane == 0) {
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % a.stages;
                mbar_wait(empty(s), ((it / a.stages) & 1) ^ 1);
                mbar_expect_tx(full(s), stage_bytes);
                const long long t = t_begin + it;
                const int grp = (int)(t / a.tiles_per_group);
                const int u = (int)(t - grp * a.tiles_per_group);
                int smp, y0 = 0, x0 = 0;
                if (a.tiles_per_sample > 0) {
                    smp = u / a.tiles_per_sample;
                    const int v = u - smp * a.tiles_per_sample;
                    y0 = (v / a.tiles_x) << a.lh;
                    x0 = (v % a.tiles_x) << a.lw;
                } else {
                    smp = u * a.TN;
                }
                const int xn = a.xoff[grp] + smp, gn = a.goff[grp] + smp;
                const uint32_t dst = sbase + s * stage_bytes;
                for (int p = 0; p < P; ++p) {
                    const uint32_t pd = dst + p * plane_bytes;
                    for (int b = 0; b < 2 * S; ++b) {
                        int rg = slab0 * 2 + b;
                        if (rg >= a.RG) rg = a.RG - 1;   // padding rows of the last slab: loaded, never stored
                        const int tap = rg / cchunks, cc = rg - tap * cchunks;
                        const int ky = tap / a.KS, kx = tap - ky * a.KS;
                        tma_load_5d(pd + b * box_bytes, &tmX, full(s), cc * 64, x0 + kx - pad, y0 + ky - pad, xn, p);
                    }
                    for (int b = 0; b < gboxes; ++b)
                        tma_load_5d(pd + (2 * S + b) * box_bytes, &tmG, full(s), co0 + b * 64, x0, y0, gn, p);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16(a.NT, 1, 1);
            const int ksteps = a.PXS / 16;
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % a.stages;
                mbar_wait(full(s), (it / a.stages) & 1);
                fence_after();
                const uint32_t st = sbase + s * stage_bytes;
                for (int ks = 0; ks < ksteps; ++ks) {
                    for (int sl = 0; sl < S; ++sl) {
                        uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                        for (int i = 0; i < P; ++i) {
                            const uint64_t ad =
                                smem_desc(st + i * plane_bytes + (2 * sl) * box_bytes + ks * 2048, box_bytes, 1024, 2);
                            for (int j = 0; i + j < P; ++j) {
                                const uint64_t bd = smem_desc(st + j * plane_bytes + (2 * S) * box_bytes + ks * 2048,
                                                              box_bytes, 1024, 2);
                                mma_bf16(tmem + sl * a.NT, ad, bd, idesc, acc);
                                acc = 1;
                            }
                        }
                    }
                }
                mma_commit(empty(s));
            }
            mma_commit(tfull);
        }
        __syncwarp();
    } else {
        mbar_wait(tfull, 0);
        fence_after();
        const int K = a.KS * a.KS * a.Cin;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        for (int sl = 0; sl < S; ++sl) {
            const int k = (slab0 + sl) * 128 + warp * 32 + lane;
            float* drow = a.dwp + (long long)k * a.Cout + co0;
            for (int c = 0; c < a.NT; c += 16) {
                float v[16];
                tmem_ld16(trow + sl * a.NT + c, v);
                if (k < K) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(drow + c + j, v[j]);
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, ncols);
}

// out[p][n][k] = plane p of w[k][n]
__global__ void pack_operand_kernel(const float* __restrict__ w, int K, int Nn, Planes out) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int k = k0 + i, n = n0 + threadIdx.x;
        tile[i][threadIdx.x] = (k < K && n < Nn) ? w[(long long)k * Nn + n] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int n = n0 + i, k = k0 + threadIdx.x;
        if (n < Nn && k < K) st1(out, (long long)n * K + k, tile[threadIdx.x][i]);
    }
}

bool g_attr_conv = false, g_attr_wgrad = false;

}  // namespace

// can the tensor-core conv take this shape?
extern "C" int pgk_conv_tc_supported(int N, int H, int W, int Cin, int Cout, int KS, int ups) {
    if (ups || (KS != 1 && KS != 3)) return 0;
    if (Cin % 64 || Cout % 16) return 0;
    if (Cout > 256 && Cout % 256) return 0;
    if (!is_pow2(H) || !is_pow2(W)) return 0;
    if (Cout < 256 && !is_pow2(Cout)) return 0;
    return N > 0;
}

extern "C" int pgk_conv_tc(const void* x, int P, int Pr, long long x_ps, int N, int H, int W, int Cin, int Cout, int KS,
                           const void* wt, long long wt_ps, const float* bias, const float* posT, const float* pos_s,
                           int act, const void* mask_ref, long long mask_ps, float out_scale, void* out,
                           long long out_ps, pgk_stream_t stream) {
    PGK_REQUIRE(pgk_conv_tc_supported(N, H, W, Cin, Cout, KS, 0), "pgk_conv_tc: unsupported shape");
    PGK_REQUIRE(P >= 1 && P <= 3 && Pr >= 1 && Pr <= P, "pgk_conv_tc: need 1 <= Pr <= P <= 3");
    ConvTcArgs a;
    a.N = N, a.H = H, a.W = W, a.Cin = Cin, a.Cout = Cout, a.KS = KS, a.P = Pr;   // the kernel's P = planes READ
    const int TW = W < 128 ? W : 128;
    const int TH = H < 128 / TW ? H : 128 / TW;
    a.TN = 128 / (TW * TH);
    a.lw = ilog2(TW), a.lh = ilog2(TH);
    a.tiles_x = W / TW, a.tiles_y = H / TH;
    const int tiles_n = (N + a.TN - 1) / a.TN;
    a.NT = Cout < 256 ? Cout : 256;
    a.bkb = 128;
    int stage_bytes = Pr * (128 * a.bkb + a.NT * a.bkb);
    if (2 * stage_bytes + 2048 > kSmemLimit && Cin % 32 == 0) {   // three planes of a 256-wide tile: halve the K slice
        a.bkb = 64;
        stage_bytes = Pr * (128 * a.bkb + a.NT * a.bkb);
    }
    // several CTAs per SM (bounded by TMEM columns and by >= 3 smem stages each): one tile's prologue / epilogue
    // overlaps another tile's main loop
    int ctas = 512 / (int)tmem_cols(Pr > 1 ? 2 * a.NT : a.NT);
    if (ctas > 4) ctas = 4;
    while (ctas > 1 && (kSmemLimit / ctas - 2048) / stage_bytes < 3) --ctas;
    a.stages = (kSmemLimit / ctas - 2048) / stage_bytes;
    if (a.stages > 8) a.stages = 8;
    PGK_REQUIRE(a.stages >= 1, "pgk_conv_tc: tile does not fit in shared memory");
    a.bias = bias, a.posT = posT, a.pos_s = pos_s, a.act = act;
    a.has_mask = mask_ref != nullptr;
    a.mask = make_planes(mask_ref, mask_ps, P);
    a.out_scale = out_scale;
    a.out = make_planes(out, out_ps, P);

    CUtensorMap tmA, tmB;
    {
        unsigned long long dims[5] = {(unsigned long long)Cin, (unsigned long long)W, (unsigned long long)H,
                                      (unsigned long long)N, (unsigned long long)P};
        unsigned long long str[4] = {2ull * Cin, 2ull * Cin * W, 2ull * Cin * W * H,
                                     P > 1 ? 2ull * x_ps : 2ull * Cin * W * H * N};
        unsigned box[5] = {(unsigned)(a.bkb / 2), (unsigned)TW, (unsigned)TH, (unsigned)a.TN, 1u};
        int rc = pgk_make_tmap(&tmA, x, 5, dims, str, box, a.bkb, "pgk_conv_tc(x)");
        if (rc) return rc;
    }
    {
        const unsigned long long K = (unsigned long long)KS * KS * Cin;
        unsigned long long dims[3] = {K, (unsigned long long)Cout, (unsigned long long)P};
        unsigned long long str[2] = {2ull * K, P > 1 ? 2ull * wt_ps : 2ull * K * Cout};
        unsigned box[3] = {(unsigned)(a.bkb / 2), (unsigned)a.NT, 1u};
        int rc = pgk_make_tmap(&tmB, wt, 3, dims, str, box, a.bkb, "pgk_conv_tc(w)");
        if (rc) return rc;
    }
    const int smem = a.stages * stage_bytes + 1024 + 256;
    if (!g_attr_conv) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
        g_attr_conv = true;
    }
    dim3 grid((unsigned)(a.tiles_x * a.tiles_y * tiles_n), (unsigned)(Cout / a.NT));
    conv_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmA, tmB, a);
    PGK_LAUNCH_CHECK("pgk_conv(tcgen05)");
    return PGK_OK;
}

extern "C" int pgk_wgrad_tc_supported(int H, int W, int Cin, int Cout, int KS, int ups, int ngroups, int group_n) {
    if (ups || (KS != 1 && KS != 3)) return 0;
    if (Cin % 64 || Cout % 64) return 0;
    if (Cout > 256 && Cout % 256) return 0;
    if (Cout < 256 && !is_pow2(Cout)) return 0;
    if (!is_pow2(H) || !is_pow2(W)) return 0;
    const int PXS = 32;
    if (H * W < PXS && (group_n % (PXS / (H * W)))) return 0;
    if ((long long)ngroups * group_n * H * W < 4096) return 0;   // tiny reductions stay on the CUDA-core kernel
    return 1;
}

extern "C" int pgk_wgrad_tc(const void* x, long long x_ps, const void* g, long long g_ps, int P, int Pr, int H, int W,
                            int Cin, int Cout, int KS, int ngroups, int group_n, const int* xoff, const int* goff,
                            float* dwp, pgk_stream_t stream) {
    PGK_REQUIRE(pgk_wgrad_tc_supported(H, W, Cin, Cout, KS, 0, ngroups, group_n), "pgk_wgrad_tc: unsupported shape");
    PGK_REQUIRE(P >= 1 && P <= 3 && Pr >= 1 && Pr <= P, "pgk_wgrad_tc: need 1 <= Pr <= P <= 3");
    PGK_REQUIRE(ngroups >= 1 && ngroups <= 4, "pgk_wgrad_tc: 1..4 groups");
    WgradTcArgs a;
    a.H = H, a.W = W, a.Cin = Cin, a.Cout = Cout, a.KS = KS, a.P = Pr;   // the kernel's P = planes READ
    a.PXS = 32;
    const int TW = W < a.PXS ? W : a.PXS;
    const int TH = H < a.PXS / TW ? H : a.PXS / TW;
    a.TN = a.PXS / (TW * TH);
    a.lw = ilog2(TW), a.lh = ilog2(TH);
    a.tiles_x = W / TW;
    a.tiles_per_sample = a.TN > 1 ? 0 : (W / TW) * (H / TH);
    a.ngroups = ngroups, a.group_n = group_n;
    int xmax = 0, gmax = 0;
    for (int i = 0; i < 4; ++i) {
        a.xoff[i] = i < ngroups ? xoff[i] : 0;
        a.goff[i] = i < ngroups ? goff[i] : 0;
        if (a.xoff[i] > xmax) xmax = a.xoff[i];
        if (a.goff[i] > gmax) gmax = a.goff[i];
    }
    a.NT = Cout < 256 ? Cout : 256;
    a.RG = KS * KS * Cin / 64;
    const int slabs = (a.RG + 1) / 2;
    const int smax = 512 / a.NT;
    const int sgroups = (slabs + smax - 1) / smax;
    a.S = (slabs + sgroups - 1) / sgroups;
    const int box_bytes = a.PXS * 128;
    const int stage_bytes = Pr * (2 * a.S + a.NT / 64) * box_bytes;
    int ctas = 512 / (int)tmem_cols(a.S * a.NT);
    if (ctas > 4) ctas = 4;
    while (ctas > 1 && (kSmemLimit / ctas - 2048) / stage_bytes < 3) --ctas;
    a.stages = (kSmemLimit / ctas - 2048) / stage_bytes;
    if (a.stages > 8) a.stages = 8;
    PGK_REQUIRE(a.stages >= 1, "pgk_wgrad_tc: stage does not fit in shared memory");
    a.tiles_per_group = (long long)group_n * H * W / a.PXS;
    a.tiles_total = a.tiles_per_group * ngroups;
    // split the pixel range so that the grid is (just under) one or two full waves of the SMs
    const int base = sgroups * (Cout / a.NT);
    const int sms = pgk_num_sms();
    long long max_split = (a.tiles_total + 15) / 16;
    long long split = 1;
    double best = -1.0;
    for (int w = 1; w <= 2; ++w) {
        long long sp = (long long)w * sms / base;
        if (sp < 1) sp = 1;
        if (sp > max_split) sp = max_split;
        const long long ctas = sp * base;
        const double util = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
        if (util > best + 0.02) best = util, split = sp;
    }
    if (split > 65535) split = 65535;
    a.tiles_per_cta = (a.tiles_total + split - 1) / split;
    split = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
    a.dwp = dwp;

    CUtensorMap tmX, tmG;
    {
        unsigned long long dims[5] = {(unsigned long long)Cin, (unsigned long long)W, (unsigned long long)H,
                                      (unsigned long long)(xmax + group_n), (unsigned long long)P};
        unsigned long long str[4] = {2ull * Cin, 2ull * Cin * W, 2ull * Cin * W * H,
                                     P > 1 ? 2ull * x_ps : 2ull * Cin * W * H * (xmax + group_n)};
        unsigned box[5] = {64u, (unsigned)TW, (unsigned)TH, (unsigned)a.TN, 1u};
        int rc = pgk_make_tmap(&tmX, x, 5, dims, str, box, 128, "pgk_wgrad_tc(x)");
        if (rc) return rc;
    }
    {
        unsigned long long dims[5] = {(unsigned long long)Cout, (unsigned long long)W, (unsigned long long)H,
                                      (unsigned long long)(gmax + group_n), (unsigned long long)P};
        unsigned long long str[4] = {2ull * Cout, 2ull * Cout * W, 2ull * Cout * W * H,
                                     P > 1 ? 2ull * g_ps : 2ull * Cout * W * H * (gmax + group_n)};
        unsigned box[5] = {64u, (unsigned)TW, (unsigned)TH, (unsigned)a.TN, 1u};
        int rc = pgk_make_tmap(&tmG, g, 5, dims, str, box, 128, "pgk_wgrad_tc(g)");
        if (rc) return rc;
    }
    const int smem = a.stages * stage_bytes + 1024 + 256;
    if (!g_attr_wgrad) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
        g_attr_wgrad = true;
    }
    dim3 grid((unsigned)sgroups, (unsigned)(Cout / a.NT), (unsigned)split);
    wgrad_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmX, tmG, a);
    PGK_LAUNCH_CHECK("pgk_wgrad(tcgen05)");
    return PGK_OK;
}

extern "C" int pgk_pack_operand(const float* w, int K, int Nn, void* out, long long out_ps, int P,
                                pgk_stream_t stream) {
    PGK_REQUIRE(P >= 1 && P <= 3 && K > 0 && Nn > 0, "pgk_pack_operand: bad arguments");
    dim3 grid((unsigned)((K + 31) / 32), (unsigned)((Nn + 31) / 32));
    pack_operand_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(w, K, Nn, make_planes(out, out_ps, P));
    PGK_LAUNCH_CHECK("pgk_pack_operand");
    return PGK_OK;
}

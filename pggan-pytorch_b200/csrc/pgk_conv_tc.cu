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
#include <stdlib.h>

#include <cuda_fp16.h>

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
// forward conv / data gradient: persistent CTAs (one per SM), tiles = (128-pixel block, NT-channel slice).
// Warps 0-7 epilogue (TMEM lane quarter = warp % 4, column half = warp / 4), warp 8 TMA producer, warp 9 MMA issue.
// Two TMEM accumulator buffers: the epilogue of tile i overlaps the main loop of tile i + 1.
// ------------------------------------------------------------------------------------------------------------
constexpr int kConvThreads = 320;
constexpr int kEpiWarps = 8;

struct ConvTcArgs {
    int N, H, W, Cin, Cout, KS, P;   // P = planes READ (products of planes i + j < P)
    int lw, lh, TN;        // pixel block = TN samples x 2^lh rows x 2^lw columns = 128 pixels
    int tiles_x, tiles_y;
    int ntiles_n;          // Cout / NT
    int total_tiles;
    int NT, stages, bkb;   // channels per tile, ring depth, bytes per K row of a stage (128 or 64)
    int split_acc;         // 1: plane0 x plane0 products and correction products in separate accumulators
    int Pout;              // planes written
    int slab;              // channels per epilogue slab (32 or 16): one TMA store box = 128 pixels x slab channels
    int nmask;             // mask staging buffers per warpgroup (2 with a mask, else 0)
    const float* bias;
    const float* posT;
    const float* pos_s;
    int act, has_mask;
    Planes mask;
    float out_scale;
    Planes out;
    float* pn_r;           // pixel norm after the activation (the tile holds every channel of a pixel: NT == Cout): the
                           // per-pixel factor rsqrt(mean_c(h^2) + 1e-8) is stored here; NULL = no pixel norm
    int fp16_a, fp16_b;    // operand formats of this launch: 1 = IEEE half planes (pgk_conv_fp16), 0 = bf16 planes
    float acc_scale;       // applied to the accumulator before the bias: 2^-PGK_FP16_WSHIFT with fp16 weights, else 1
};

// P (planes read), KSTEPS (= bkb / 32) and SPLIT (separate correction accumulator) are compile-time: the MMA-issue
// sequence of a k-slice unrolls to one descriptor add + one UTCHMMA per product.  (With run-time loop bounds the
// dependent uniform-datapath address arithmetic cost ~350 cycles per MMA -- three times the MMA itself.)
// F16 = 1 (pgk_conv_fp16): IEEE-half operand planes -- the operand formats come from the arguments and the accumulator
// is scaled back before the bias.  F16 = 0 instances carry none of that (their machine code is the GPU-verified one).
template <int P, int KSTEPS, int SPLIT, int F16>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmM, const ConvTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    constexpr uint32_t kBkb = 32u * KSTEPS;
    constexpr uint32_t a_bytes = 128u * kBkb;
    const uint32_t b_bytes = (uint32_t)a.NT * kBkb;
    const uint32_t stage_bytes = P * (a_bytes + b_bytes);
    const uint32_t bars = sbase + a.stages * stage_bytes;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (a.stages + s); };
    const uint32_t tfull0 = bars + 16u * a.stages;   // tfull[2], tempty[2]
    auto tfull = [&](int b) { return tfull0 + 8u * b; };
    auto tempty = [&](int b) { return tfull0 + 16u + 8u * b; };
    auto mfull = [&](int wg, int b) { return tfull0 + 32u + 16u * wg + 8u * b; };   // mask slab landed (per warpgroup)
    const uint32_t tptr = tfull0 + 64u;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (tptr + 16u - raw));   // [2][NT]
    // epilogue staging, per warpgroup: 2 mask slabs + Pout output slabs of 128 rows x slab channels (1024-aligned)
    const uint32_t slab_bytes = 128u * a.slab * 2u;
    const uint32_t epi_wg_bytes = (uint32_t)(a.nmask + a.Pout) * slab_bytes;
    float* ss_s = bias_s + 2 * a.NT;                                            // [2 warpgroups][128 pixels]
    const uint32_t epi_base = (tptr + 16u + 2u * 4u * a.NT + 1024u + 1023u) & ~1023u;

    constexpr int kelems = kBkb >> 1;
    const int kchunks = a.Cin / kelems;
    const int nk = a.KS * a.KS * kchunks;
    const int pad = a.KS >> 1;
    const int acc_cols = SPLIT ? 2 * a.NT : a.NT;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull(b), 1);
            mbar_init(tempty(b), kEpiWarps);
            mbar_init(mfull(0, b), 1);
            mbar_init(mfull(1, b), 1);
        }
        fence_barrier_init();
    }
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        tma_prefetch_desc(&tmM);
    }
    const unsigned ncols = tmem_cols(2 * acc_cols);
    if (warp == 9) tmem_alloc(tptr, ncols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
    pgk_pdl_enter();   // everything above touched shared / tensor memory and the kernel parameters only

    auto tile_coords = [&](int t, int& x0, int& y0, int& n0, int& co0) {
        const int nt = t % a.ntiles_n;
        int pt = t / a.ntiles_n;
        const int tx = pt % a.tiles_x;
        pt /= a.tiles_x;
        const int ty = pt % a.tiles_y;
        const int tn = pt / a.tiles_y;
        x0 = tx << a.lw, y0 = ty << a.lh, n0 = tn * a.TN, co0 = nt * a.NT;
    };

    // The producer and the MMA issuer are single threads: their loops carry only counters (no divisions, no
    // descriptor rebuilds) -- a dependent-instruction chain of a few hundred cycles per k-slice would otherwise
    // bound the whole kernel.
    if (warp == 8) {
        int s = 0;
        uint32_t ph = 1;   // parity to wait for on the empty barrier of stage s (first pass: passes immediately)
        constexpr uint32_t a_all = P * a_bytes;
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            int x0, y0, n0, co0;
            tile_coords(t, x0, y0, n0, co0);
            int kcol = 0;   // K coordinate of the weight slice
            for (int ky = 0; ky < a.KS; ++ky) {
                for (int kx = 0; kx < a.KS; ++kx) {
                    const int xs = x0 + kx - pad, ys = y0 + ky - pad;
                    for (int c = 0; c < a.Cin; c += kelems, kcol += kelems) {
                        mbar_wait_spin(empty(s), ph);
                        const uint32_t fb = full(s);
                        const uint32_t dst = sbase + s * stage_bytes;
                        if (elect_one()) {
                            mbar_expect_tx(fb, stage_bytes);
#pragma unroll
                            for (int p = 0; p < P; ++p) tma_load_5d(dst + p * a_bytes, &tmA, fb, c, xs, ys, n0, p);
#pragma unroll
                            for (int p = 0; p < P; ++p) tma_load_3d(dst + a_all + p * b_bytes, &tmB, fb, kcol, co0, p);
                        }
                        __syncwarp();
                        if (++s == a.stages) s = 0, ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 9) {
        const uint32_t idesc = F16 ? idesc_f16(a.NT, 0, 0, a.fp16_a, a.fp16_b) : idesc_bf16(a.NT, 0, 0);
        // descriptor = constant high part | (address >> 4) in the low 14 bits: per MMA only an add
        const uint64_t dbase = smem_desc(0, 16, 8u * kBkb, KSTEPS == 4 ? 2u : 4u);
        constexpr uint32_t a16 = a_bytes >> 4;
        const uint32_t b16 = b_bytes >> 4;
        int s = 0, ti = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++ti) {
            const int b = ti & 1;
            mbar_wait_spin(tempty(b), ((ti >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator buffer
            fence_after();
            const uint32_t d_main = tmem + b * acc_cols, d_corr = d_main + a.NT;
            for (int kc = 0; kc < nk; ++kc) {
                mbar_wait_spin(full(s), ph);
                fence_after();
                const uint32_t abase = (sbase + s * stage_bytes) >> 4;
                const uint64_t ad0 = dbase | abase, bd0 = dbase | (abase + P * a16);
                if (elect_one()) {
                    const uint32_t later = kc > 0 ? 1u : 0u;   // 0 only for the first products of a tile
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
                        for (int i = 0; i < P; ++i) {
#pragma unroll
                            for (int j = 0; j < P - i; ++j) {
                                const uint64_t ad = ad0 + (uint32_t)(i * a16 + ks * 2);
                                const uint64_t bd = bd0 + (uint32_t)(j * b16 + ks * 2);
                                if (i + j == 0 || !SPLIT) {
                                    mma_bf16(d_main, ad, bd, idesc, (ks == 0 && i + j == 0) ? later : 1u);
                                } else {
                                    mma_bf16(d_corr, ad, bd, idesc, (ks == 0 && i == 0 && j == 1) ? later : 1u);
                                }
                            }
                        }
                    }
                    mma_commit(empty(s));
                }
                __syncwarp();
                if (++s == a.stages) s = 0, ph ^= 1;
            }
            if (elect_one()) mma_commit(tfull(b));
            __syncwarp();
        }
    } else {
        // ---- epilogue: one accumulator row (= one pixel) per thread; warpgroup wg = warp / 4 owns half of the tile's
        // channels and walks them in slabs: TMEM -> registers -> (bias, stddev channel, LeakyReLU, mask, scale,
        // plane split) -> swizzled shared memory -> one TMA tensor store per plane.  Masks arrive the same way (TMA
        // load of plane 0, whose sign is the value's sign), double buffered one slab ahead.
        const int q = warp & 3, wg = warp >> 2;
        const int wcols = a.NT >= 32 ? a.NT / 2 : a.NT;
        const bool active = wg == 0 || a.NT >= 32;
        const int nslabs = wcols / a.slab;
        const int r = q * 32 + lane;
        const int px = r & ((1 << a.lw) - 1);
        const int py = (r >> a.lw) & ((1 << a.lh) - 1);
        const int pn = r >> (a.lw + a.lh);
        const bool lead_warp = (warp & 3) == 0;   // warp-uniform: its elected lane issues the warpgroup's TMA traffic
        const uint32_t row_bytes = a.slab * 2u;
        const uint32_t swz = a.slab == 32 ? 3u : 1u;
        const uint32_t wg_base = epi_base + wg * epi_wg_bytes;
        auto sm_mask = [&](int b) { return wg_base + b * slab_bytes; };
        auto sm_out = [&](int p) { return wg_base + (a.nmask + p) * slab_bytes; };
        auto chunk_addr = [&](uint32_t buf, int j) {   // 16-byte chunk j of this thread's row, TMA swizzle applied
            const uint32_t off = r * row_bytes + j * 16u;
            return buf + (off ^ (((off >> 7) & swz) << 4));
        };
        const int barid = 2 + wg;
        int sc = 0;   // slabs processed by this warpgroup (mask buffer / phase bookkeeping)
        if (active && a.has_mask && lead_warp && blockIdx.x < a.total_tiles) {
            int x0, y0, n0, co0;
            tile_coords(blockIdx.x, x0, y0, n0, co0);
            if (elect_one()) {
                mbar_expect_tx(mfull(wg, 0), slab_bytes);
                tma_load_5d(sm_mask(0), &tmM, mfull(wg, 0), co0 + wg * wcols, x0, y0, n0, 0);
            }
            __syncwarp();
        }
        int ti = 0;
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++ti) {
            const int b = ti & 1;
            int x0, y0, n0, co0;
            tile_coords(t, x0, y0, n0, co0);
            // stage this tile's bias slice while the main loop runs (buffer b: the other one may still be in use)
            float* bs = bias_s + b * a.NT;
            if (a.bias) {
                for (int i = threadIdx.x; i < a.NT; i += kEpiWarps * 32) bs[i] = __ldg(a.bias + co0 + i);
            }
            named_bar_sync(1, kEpiWarps * 32);
            const int n = n0 + pn, x = x0 + px, y = y0 + py;
            const bool valid = n < a.N;
            const float ps = (a.posT && valid) ? __ldg(a.pos_s + n) : 0.f;
            const float* posrow = a.posT ? a.posT + (long long)(y * a.W + x) * a.Cout : nullptr;
            mbar_wait(tfull(b), (ti >> 1) & 1);
            fence_after();
            const uint32_t trow = tmem + b * acc_cols + ((uint32_t)(q * 32) << 16);
            float pn_rr = 1.f;
            if (a.pn_r) {
                // pixel norm (network.py:37-40), fused: the tile holds all Cout channels of its 128 pixels, half of them
                // per warpgroup.  First pass over the accumulator: bias + LeakyReLU, sum of squares of this thread's
                // half; the two halves of a pixel meet in shared memory; the slab loop below reads the accumulator
                // again and scales what it stores -- one rounding of the finished value, as in the thin kernel.
                float ss = 0.f;
                if (active) {
                    for (int c = wg * wcols; c < (wg + 1) * wcols; c += 16) {
                        float v[16];
                        tmem_ld16(trow + c, v);
                        if (SPLIT) {
                            float w[16];
                            tmem_ld16(trow + a.NT + c, w);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] += w[j];
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float f = F16 ? v[j] * a.acc_scale : v[j];
                            if (a.bias) f += bs[c + j];
                            if (a.act) f = lrelu(f);
                            ss = fmaf(f, f, ss);
                        }
                    }
                }
                ss_s[wg * 128 + r] = ss;
                named_bar_sync(4, kEpiWarps * 32);
                pn_rr = rsqrtf((ss_s[r] + ss_s[128 + r]) / (float)a.Cout + 1e-8f);
                if (wg == 0 && valid) a.pn_r[((long long)n * a.H + y) * a.W + x] = pn_rr;
            }
            if (active) {
                for (int sl = 0; sl < nslabs; ++sl, ++sc) {
                    const int c0 = wg * wcols + sl * a.slab;   // first channel of the slab within the tile
                    if (a.has_mask) {
                        if (lead_warp) {   // prefetch the next slab's mask (next tile's first slab at the end of a tile)
                            int nx0 = x0, ny0 = y0, nn0 = n0, nco0 = co0, nsl = sl + 1;
                            bool more = true;
                            if (nsl == nslabs) {
                                nsl = 0;
                                more = t + (int)gridDim.x < a.total_tiles;
                                if (more) tile_coords(t + gridDim.x, nx0, ny0, nn0, nco0);
                            }
                            if (more && elect_one()) {
                                const int nb = (sc + 1) & 1;
                                mbar_expect_tx(mfull(wg, nb), slab_bytes);
                                tma_load_5d(sm_mask(nb), &tmM, mfull(wg, nb), nco0 + wg * wcols + nsl * a.slab, nx0, ny0,
                                            nn0, 0);
                            }
                            __syncwarp();
                        }
                        mbar_wait(mfull(wg, sc & 1), (sc >> 1) & 1);
                    }
                    // the previous slab's stores no longer read the staging tiles (bulk groups belong to the issuing
                    // thread: elect.sync picks the same lane every time)
                    if (lead_warp) {
                        if (elect_one()) tma_store_wait_read();
                        __syncwarp();
                    }
                    named_bar_sync(barid, 128);
                    for (int cc = 0; cc < a.slab; cc += 16) {
                        float v[16];
                        tmem_ld16(trow + c0 + cc, v);
                        if (SPLIT) {
                            float w[16];
                            tmem_ld16(trow + a.NT + c0 + cc, w);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] += w[j];
                        }
                        if (F16) {   // fp16 weight operands are packed times a power of two
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] *= a.acc_scale;
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float* f = v + 8 * h;
                            const int cl = c0 + cc + 8 * h;          // channel within the tile
                            const int jc = (cc >> 3) + h;             // 16-byte chunk within the slab row
                            if (a.bias) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] += bs[cl + j];
                            }
                            if (posrow) {
                                const float4 p0 = __ldg(reinterpret_cast<const float4*>(posrow + co0 + cl));
                                const float4 p1 = __ldg(reinterpret_cast<const float4*>(posrow + co0 + cl + 4));
                                f[0] = fmaf(ps, p0.x, f[0]), f[1] = fmaf(ps, p0.y, f[1]), f[2] = fmaf(ps, p0.z, f[2]);
                                f[3] = fmaf(ps, p0.w, f[3]), f[4] = fmaf(ps, p1.x, f[4]), f[5] = fmaf(ps, p1.y, f[5]);
                                f[6] = fmaf(ps, p1.z, f[6]), f[7] = fmaf(ps, p1.w, f[7]);
                            }
                            if (a.act) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] = lrelu(f[j]);
                            }
                            if (a.pn_r) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] *= pn_rr;
                            }
                            if (a.has_mask) {
                                float m[8];
                                unpack8(ld_shared_v4(chunk_addr(sm_mask(sc & 1), jc)), m);
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] *= lrelu_grad(m[j]);
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] *= a.out_scale;
                            for (int p = 0; p < a.Pout; ++p) {
                                uint4 qv;
                                qv.x = pack2(f[0], f[1]), qv.y = pack2(f[2], f[3]);
                                qv.z = pack2(f[4], f[5]), qv.w = pack2(f[6], f[7]);
                                st_shared_v4(chunk_addr(sm_out(p), jc), qv);
                                if (p + 1 < a.Pout) {
                                    float hv[8];
                                    unpack8(qv, hv);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) f[j] -= hv[j];
                                }
                            }
                        }
                    }
                    if (sl == nslabs - 1) {   // all TMEM reads of this buffer are complete: hand it back to the MMA warp
                        fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty(b));
                    }
                    fence_proxy_async();
                    named_bar_sync(barid, 128);
                    if (lead_warp) {
                        if (elect_one()) {
                            for (int p = 0; p < a.Pout; ++p) tma_store_5d(&tmO, sm_out(p), co0 + c0, x0, y0, n0, p);
                            tma_store_commit();
                        }
                        __syncwarp();
                    }
                }
            } else {
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty(b));
            }
        }
        if (lead_warp) {
            if (elect_one()) tma_store_wait_all();
            __syncwarp();
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem, ncols);
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
    float* db;            // optional fused bias gradient: db[co] += sum over the pixels of the groups in bias_mask of G
    unsigned bias_mask;
};

// four consecutive floats added to global memory in one reduction (sm_90+; the address must be 16-byte aligned)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// The flush into dwp uses 16-byte vector reductions (SASS REDG.E.ADD.F32x4), or a plain read-add-write where the pixel
// range is not split (this thread is then the only writer of its row slice): at batch 4 the flush of up to 147 CTAs x
// 36 864 scalar atomics was the larger part of a 50-120 us launch (measured: -5 % of the depth-8 step, -3.5 % at depth 6).
template <int P>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                const WgradTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    const int S = a.S;
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
    // the bias gradient (column sums of G) rides along in the CTAs of the first slab group: warps 0-3, idle until the
    // flush, add up every G tile from the stage ring while the MMAs of that stage run
    const bool do_bias = a.db != nullptr && blockIdx.x == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), do_bias ? 5 : 1);   // the MMA commit (+ the four bias-gradient warps)
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
    pgk_pdl_enter();   // everything above touched shared / tensor memory and the kernel parameters only

    if (warp == 4) {
        {
            // warp-uniform loop with counters only (see conv_tc_kernel); per-box tap offsets are computed once
            int bc[16], bdx[16], bdy[16];
            for (int b = 0; b < 2 * S; ++b) {
                int rg = slab0 * 2 + b;
                if (rg >= a.RG) rg = a.RG - 1;   // padding rows of the last slab: loaded, never stored
                const int tap = rg / cchunks, cc = rg - tap * cchunks;
                const int ky = tap / a.KS, kx = tap - ky * a.KS;
                bc[b] = cc * 64, bdx[b] = kx - pad, bdy[b] = ky - pad;
            }
            // position of the first pixel block of this CTA
            int grp = (int)(t_begin / a.tiles_per_group);
            int u = (int)(t_begin - grp * a.tiles_per_group);
            int smp, v = 0;
            if (a.tiles_per_sample > 0) {
                smp = u / a.tiles_per_sample;
                v = u - smp * a.tiles_per_sample;
            } else {
                smp = u * a.TN;
            }
            int vy = v / a.tiles_x, vx = v - vy * a.tiles_x;
            int s = 0;
            uint32_t ph = 1;
            for (int it = 0; it < ntiles; ++it) {
                mbar_wait_spin(empty(s), ph);
                const uint32_t fb = full(s);
                const int y0 = vy << a.lh, x0 = vx << a.lw;
                const int xn = a.xoff[grp] + smp, gn = a.goff[grp] + smp;
                const uint32_t dst = sbase + s * stage_bytes;
                if (elect_one()) {
                    mbar_expect_tx(fb, stage_bytes);
                    for (int p = 0; p < P; ++p) {
                        const uint32_t pd = dst + p * plane_bytes;
                        for (int b = 0; b < 2 * S; ++b)
                            tma_load_5d(pd + b * box_bytes, &tmX, fb, bc[b], x0 + bdx[b], y0 + bdy[b], xn, p);
                        for (int b = 0; b < gboxes; ++b)
                            tma_load_5d(pd + (2 * S + b) * box_bytes, &tmG, fb, co0 + b * 64, x0, y0, gn, p);
                    }
                }
                __syncwarp();
                if (++s == a.stages) s = 0, ph ^= 1;
                // next pixel block
                if (a.tiles_per_sample > 0) {
                    if (++vx == a.tiles_x) {
                        vx = 0;
                        if (++vy == a.tiles_per_sample / a.tiles_x) vy = 0, ++smp;
                    }
                } else {
                    smp += a.TN;
                }
                if (smp >= a.group_n) smp = 0, ++grp;
            }
        }
    } else if (warp == 5) {
        const uint32_t idesc = idesc_bf16(a.NT, 1, 1);
        const uint64_t dbase = smem_desc(0, box_bytes, 1024, 2);
        const uint32_t pl16 = plane_bytes >> 4, bx16 = box_bytes >> 4;
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < ntiles; ++it) {
            mbar_wait_spin(full(s), ph);
            fence_after();
            const uint32_t st = (sbase + s * stage_bytes) >> 4;
            const uint64_t bd0 = dbase | (st + 2 * S * bx16);
            if (elect_one()) {
                const uint32_t later = it > 0 ? 1u : 0u;
                uint64_t ad_sl = dbase | st;
                uint32_t d = tmem;
                for (int sl = 0; sl < S; ++sl, ad_sl += 2 * bx16, d += a.NT) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {   // PXS = 32 pixels = two K = 16 steps
#pragma unroll
                        for (int i = 0; i < P; ++i) {
#pragma unroll
                            for (int j = 0; j < P - i; ++j)
                                mma_bf16(d, ad_sl + (uint32_t)(i * pl16 + ks * 128), bd0 + (uint32_t)(j * pl16 + ks * 128),
                                         idesc, (ks == 0 && i + j == 0) ? later : 1u);
                        }
                    }
                }
                mma_commit(empty(s));
            }
            __syncwarp();
            if (++s == a.stages) s = 0, ph ^= 1;
        }
        if (elect_one()) mma_commit(tfull);
        __syncwarp();
    } else {
        if (do_bias) {
            // thread = (16-byte chunk c of the 128-byte pixel row, pixel group pg): pixels pg and pg + 16 of every box;
            // SWIZZLE_128B: chunk c of pixel row px sits at chunk position c ^ (px & 7)
            const int t = warp * 32 + lane, c = t & 7, pg = t >> 3;
            float acc[4][8];
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[b][j] = 0.f;
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < ntiles; ++it) {
                mbar_wait(full(s), ph);
                const int grp = (int)((t_begin + it) / a.tiles_per_group);
                if ((a.bias_mask >> grp) & 1) {
                    const uint32_t g0 = sbase + s * stage_bytes + 2 * S * box_bytes;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            if (b < gboxes) {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const uint32_t px = (uint32_t)(pg + 16 * h);
                                    float f[8];
                                    unpack8(ld_shared_v4(g0 + p * plane_bytes + b * box_bytes + px * 128u +
                                                         ((uint32_t)(c ^ (int)(px & 7u)) << 4)), f);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) acc[b][j] += f[j];
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty(s));
                if (++s == a.stages) s = 0, ph ^= 1;
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (b < gboxes) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v = acc[b][j];
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (lane < 8 && v != 0.f) atomicAdd(a.db + co0 + b * 64 + c * 8 + j, v);
                    }
                }
            }
        }
        mbar_wait(tfull, 0);
        fence_after();
        const int K = a.KS * a.KS * a.Cin;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        // A thread owns one accumulator row (lane = row), i.e. one row of dW: written directly, a warp's 16-byte access
        // touches 32 rows -- 32 partial sectors per instruction, and this flush was the larger part of the small
        // launches (4 x 32^2 pixels, 512 -> 256: 6 us of MMAs in a 144 us launch).  The stage ring is idle by now: each
        // warp turns its 32 rows x 32 columns through it (row pitch 36 floats: both directions conflict free) so that
        // every global access covers 4 rows x 128 contiguous bytes.
        constexpr uint32_t kPitch = 144u;
        const bool staged = (uint32_t)a.stages * stage_bytes >= 4u * 32u * kPitch;
        const uint32_t fl = sbase + (uint32_t)warp * (32u * kPitch);
        for (int sl = 0; sl < S; ++sl) {
            const int kw = (slab0 + sl) * 128 + warp * 32;     // first row of this warp
            for (int c = 0; c < a.NT; c += 32) {
                float v[32];
                tmem_ld16(trow + sl * a.NT + c, v);
                tmem_ld16(trow + sl * a.NT + c + 16, v + 16);
                if (staged) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(fl + (uint32_t)lane * kPitch + 16u * j,
                                     make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
                    __syncwarp();
                    const int q = lane & 7, r0 = lane >> 3;
                    float* dcol = a.dwp + co0 + c + 4 * q;
                    if (gridDim.z == 1) {
                        // the pixel range is not split: this CTA is the only writer of its tile, no atomics
                        float4 o[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int k = kw + it * 4 + r0;
                            if (k < K) o[it] = *reinterpret_cast<const float4*>(dcol + (long long)k * a.Cout);
                        }
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int k = kw + it * 4 + r0;
                            if (k < K) {
                                const uint4 x = ld_shared_v4(fl + (uint32_t)(it * 4 + r0) * kPitch + 16u * q);
                                o[it].x += __uint_as_float(x.x), o[it].y += __uint_as_float(x.y);
                                o[it].z += __uint_as_float(x.z), o[it].w += __uint_as_float(x.w);
                                *reinterpret_cast<float4*>(dcol + (long long)k * a.Cout) = o[it];
                            }
                        }
                    } else {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int k = kw + it * 4 + r0;
                            if (k < K) {
                                const uint4 x = ld_shared_v4(fl + (uint32_t)(it * 4 + r0) * kPitch + 16u * q);
                                red_add_v4(dcol + (long long)k * a.Cout, __uint_as_float(x.x), __uint_as_float(x.y),
                                           __uint_as_float(x.z), __uint_as_float(x.w));
                            }
                        }
                    }
                    __syncwarp();
                } else {
                    const int k = kw + lane;
                    float* drow = a.dwp + (long long)k * a.Cout + co0 + c;
                    if (k < K) {
                        if (gridDim.z == 1) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4* p4 = reinterpret_cast<float4*>(drow + j);
                                float4 o = *p4;
                                o.x += v[j], o.y += v[j + 1], o.z += v[j + 2], o.w += v[j + 3];
                                *p4 = o;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) red_add_v4(drow + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
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
    pgk_pdl_enter();
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

// out[p][n][k] = IEEE-half plane p of w[k][n] * scale (scale = 2^PGK_FP16_WSHIFT keeps the low plane of the small
// equalised-LR weights out of the fp16 subnormals; the conv scales its accumulator back)
__global__ void pack_operand_h_kernel(const float* __restrict__ w, int K, int Nn, float scale, __half* out,
                                      long long out_ps, int P) {
    pgk_pdl_enter();
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int k = k0 + i, n = n0 + threadIdx.x;
        tile[i][threadIdx.x] = (k < K && n < Nn) ? w[(long long)k * Nn + n] * scale : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int n = n0 + i, k = k0 + threadIdx.x;
        if (n < Nn && k < K) {
            float v = tile[threadIdx.x][i];
            for (int pl = 0; pl < P; ++pl) {
                const __half h = __float2half_rn(v);
                out[pl * out_ps + (long long)n * K + k] = h;
                v -= __half2float(h);
            }
        }
    }
}

}  // namespace

// does the wide kernel apply the pixel norm itself?  Its tile must hold every channel of a pixel: Cout <= 256 channels
// of accumulator, 128 where the correction products have an accumulator of their own (three planes / half operands)
extern "C" int pgk_conv_tc_fuses_pixelnorm(int Cout, int split_acc) { return Cout <= (split_acc ? 128 : 256); }

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
                           long long out_ps, pgk_stream_t stream, int fp16_x, int fp16_w, float acc_scale, float* pn_r) {
    PGK_REQUIRE(pgk_conv_tc_supported(N, H, W, Cin, Cout, KS, 0), "pgk_conv_tc: unsupported shape");
    PGK_REQUIRE(!pn_r || (pgk_conv_tc_fuses_pixelnorm(Cout, Pr == 3 || (fp16_x && fp16_w)) && !mask_ref && !posT &&
                          out_scale == 1.0f),
                "pgk_conv_tc: the fused pixel norm needs all channels of a pixel in one tile, no mask, no stddev channel");
    PGK_REQUIRE(P >= 1 && P <= 3 && Pr >= 1 && Pr <= P, "pgk_conv_tc: need 1 <= Pr <= P <= 3");
    ConvTcArgs a;
    a.N = N, a.H = H, a.W = W, a.Cin = Cin, a.Cout = Cout, a.KS = KS, a.P = Pr;   // the kernel's P = planes READ
    const int TW = W < 128 ? W : 128;
    const int TH = H < 128 / TW ? H : 128 / TW;
    a.TN = 128 / (TW * TH);
    a.lw = ilog2(TW), a.lh = ilog2(TH);
    a.tiles_x = W / TW, a.tiles_y = H / TH;
    const int tiles_n = (N + a.TN - 1) / a.TN;
    // Full-precision passes (three planes read) keep the plane0 x plane0 products in their own accumulator; with two
    // accumulator buffers in the 512 TMEM columns that allows 128 channels per tile, otherwise 256.  The fp16 forward
    // pass (pgk_conv_fp16: two half planes, 22 bits) is a full-precision pass too: its hi x hi products get the same
    // treatment, or three times as many truncating accumulator adds would land on the main sum.
    a.split_acc = (Pr == 3 || (Pr == 2 && fp16_x && fp16_w)) ? 1 : 0;
    int nt_max = a.split_acc ? 128 : 256;
    if (const char* e = getenv("PGK_CONV_NT")) nt_max = atoi(e) < nt_max ? atoi(e) : nt_max;   // tuning knob
    a.NT = Cout < nt_max ? Cout : nt_max;
    a.pn_r = pn_r;
    // few pixel blocks (the 4x4 ... 16x16 levels at small batch): the launch is bound by streaming the weights, so
    // spread them over more CTAs with narrower channel slices
    {
        const int pixel_tiles = a.tiles_x * a.tiles_y * tiles_n, sms = pgk_num_sms();
        while (!pn_r && a.NT > 32 && pixel_tiles * (Cout / a.NT) < sms) a.NT >>= 1;   // (pixel norm: NT stays Cout)
    }
    a.ntiles_n = Cout / a.NT;
    a.total_tiles = a.tiles_x * a.tiles_y * tiles_n * a.ntiles_n;
    a.Pout = P;
    a.nmask = mask_ref ? 2 : 0;
    const int wcols = a.NT >= 32 ? a.NT / 2 : a.NT;
    a.slab = (wcols >= 32 && !(P == 3 && mask_ref)) ? 32 : 16;
    const int epi_bytes = 2 * (a.nmask + a.Pout) * 128 * a.slab * 2 + 1024;
    const int fixed = 1024 + 512 + 2 * 4 * a.NT + 1024 + epi_bytes;   // (+1024: the pixel norm's exchange buffer)
    const int budget = kSmemLimit - fixed;
    a.bkb = 128;
    int stage_bytes = Pr * (128 * a.bkb + a.NT * a.bkb);
    if (budget / stage_bytes < 3 && Cin % 32 == 0) {   // keep the ring at least three deep: halve the K slice
        a.bkb = 64;
        stage_bytes = Pr * (128 * a.bkb + a.NT * a.bkb);
    }
    if (const char* e = getenv("PGK_CONV_BKB")) {   // tuning knob
        a.bkb = atoi(e);
        stage_bytes = Pr * (128 * a.bkb + a.NT * a.bkb);
    }
    a.stages = budget / stage_bytes;
    if (a.stages > 8) a.stages = 8;
    PGK_REQUIRE(a.stages >= 1, "pgk_conv_tc: tile does not fit in shared memory");
    a.bias = bias, a.posT = posT, a.pos_s = pos_s, a.act = act;
    a.has_mask = mask_ref != nullptr;
    a.mask = make_planes(mask_ref, mask_ps, P);
    a.out_scale = out_scale;
    a.out = make_planes(out, out_ps, P);
    a.fp16_a = fp16_x ? 1 : 0, a.fp16_b = fp16_w ? 1 : 0;
    a.acc_scale = acc_scale;

    CUtensorMap tmA, tmB;
    // (half operands come as exactly two planes: their tensor maps must not declare a third one behind the allocation)
    const unsigned long long planes_x = fp16_x ? 2ull : (unsigned long long)P, planes_w = fp16_w ? 2ull : 3ull;
    {
        unsigned long long dims[5] = {(unsigned long long)Cin, (unsigned long long)W, (unsigned long long)H,
                                      (unsigned long long)N, planes_x};
        unsigned long long str[4] = {2ull * Cin, 2ull * Cin * W, 2ull * Cin * W * H,
                                     P > 1 ? 2ull * x_ps : 2ull * Cin * W * H * N};
        unsigned box[5] = {(unsigned)(a.bkb / 2), (unsigned)TW, (unsigned)TH, (unsigned)a.TN, 1u};
        int rc = pgk_make_tmap(&tmA, x, 5, dims, str, box, a.bkb, "pgk_conv_tc(x)");
        if (rc) return rc;
    }
    {
        const unsigned long long K = (unsigned long long)KS * KS * Cin;
        unsigned long long dims[3] = {K, (unsigned long long)Cout, planes_w};
        unsigned long long str[2] = {2ull * K, 2ull * wt_ps};
        unsigned box[3] = {(unsigned)(a.bkb / 2), (unsigned)a.NT, 1u};
        int rc = pgk_make_tmap(&tmB, wt, 3, dims, str, box, a.bkb, "pgk_conv_tc(w)");
        if (rc) return rc;
    }
    CUtensorMap tmO, tmM;
    {
        unsigned long long dims[5] = {(unsigned long long)Cout, (unsigned long long)W, (unsigned long long)H,
                                      (unsigned long long)N, (unsigned long long)P};
        unsigned long long str[4] = {2ull * Cout, 2ull * Cout * W, 2ull * Cout * W * H,
                                     P > 1 ? 2ull * out_ps : 2ull * Cout * W * H * N};
        unsigned box[5] = {(unsigned)a.slab, (unsigned)TW, (unsigned)TH, (unsigned)a.TN, 1u};
        int rc = pgk_make_tmap(&tmO, out, 5, dims, str, box, a.slab * 2, "pgk_conv_tc(out)");
        if (rc) return rc;
        tmM = tmO;
        if (mask_ref) {
            str[3] = P > 1 ? 2ull * mask_ps : 2ull * Cout * W * H * N;
            rc = pgk_make_tmap(&tmM, mask_ref, 5, dims, str, box, a.slab * 2, "pgk_conv_tc(mask)");
            if (rc) return rc;
        }
    }
    const int smem = a.stages * stage_bytes + fixed;
    int grid = pgk_num_sms();
    if (grid > a.total_tiles) grid = a.total_tiles;
    typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvTcArgs);
    kern_t kern = nullptr;
    const int ks = a.bkb / 32;
    const int f16 = (a.fp16_a || a.fp16_b) ? 1 : 0;
    PGK_REQUIRE(!f16 || (a.fp16_a && a.fp16_b && Pr == 2), "pgk_conv_tc: half operands come as two planes of both x and w");
#define PGK_CONV_CASE(P_, K_, S_, F_) \
    if (Pr == P_ && ks == K_ && a.split_acc == S_ && f16 == F_) kern = conv_tc_kernel<P_, K_, S_, F_>;
    PGK_CONV_CASE(1, 4, 0, 0) PGK_CONV_CASE(1, 2, 0, 0)
    PGK_CONV_CASE(2, 4, 0, 0) PGK_CONV_CASE(2, 2, 0, 0)
    PGK_CONV_CASE(2, 4, 1, 1) PGK_CONV_CASE(2, 2, 1, 1)
    PGK_CONV_CASE(3, 4, 1, 0) PGK_CONV_CASE(3, 2, 1, 0)
#undef PGK_CONV_CASE
    PGK_REQUIRE(kern != nullptr, "pgk_conv_tc: no kernel instance for Pr %d ksteps %d split %d", Pr, ks, a.split_acc);
    static bool attr_done[3][2][2] = {};
    if (!attr_done[Pr - 1][ks == 4][a.split_acc]) {
        cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
        attr_done[Pr - 1][ks == 4][a.split_acc] = true;
    }
    pgk_launch(kern, grid, kConvThreads, smem, (cudaStream_t)stream, tmA, tmB, tmO, tmM, a);
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
    // very small reductions stay on the CUDA-core kernel (PGK_WGRAD_TC_MIN pixels, default 256: measured on the
    // 4x4 .. 16x16 levels at batch 4 the tensor-core kernel is ~8x faster from 256 pixels up)
    static long long min_px = -1;
    if (min_px < 0) {
        const char* e = getenv("PGK_WGRAD_TC_MIN");
        min_px = e ? atoll(e) : 256;
        if (min_px < 32) min_px = 32;
    }
    if ((long long)ngroups * group_n * H * W < min_px) return 0;
    return 1;
}

extern "C" int pgk_wgrad_tc(const void* x, long long x_ps, const void* g, long long g_ps, int P, int Pr, int H, int W,
                            int Cin, int Cout, int KS, int ngroups, int group_n, const int* xoff, const int* goff,
                            float* dwp, float* db, unsigned bias_mask, int* bias_fused, pgk_stream_t stream) {
    PGK_REQUIRE(pgk_wgrad_tc_supported(H, W, Cin, Cout, KS, 0, ngroups, group_n), "pgk_wgrad_tc: unsupported shape");
    PGK_REQUIRE((((uintptr_t)dwp) & 15) == 0, "pgk_wgrad_tc: dwp must be 16-byte aligned (vector reductions)");
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
    a.RG = KS * KS * Cin / 64;
    const int slabs = (a.RG + 1) / 2;
    const int box_bytes = a.PXS * 128;
    a.tiles_per_group = (long long)group_n * H * W / a.PXS;
    a.tiles_total = a.tiles_per_group * ngroups;
    const int sms = pgk_num_sms();
    const long long max_split = (a.tiles_total + 15) / 16;
    // ---- plan: widest channel tile, S * NT = 512 accumulator columns, the pixel range split to one or two waves.
    // (A cost model that avoided splitting short reductions was tried in round 2b and measured slower twice -- c4
    // wgrad_tc 0.66 -> 0.87 / 0.80 ms per iteration: un-split plans re-load X per tap from L2 -- and was removed; the
    // staged flush below is what helped those shapes.)
    a.NT = Cout < 256 ? Cout : 256;
    const int smax = 512 / a.NT;
    const int sgroups = (slabs + smax - 1) / smax;
    a.S = (slabs + sgroups - 1) / sgroups;
    const int stage_bytes = Pr * (2 * a.S + a.NT / 64) * box_bytes;
    int ctas = 512 / (int)tmem_cols(a.S * a.NT);
    if (ctas > 4) ctas = 4;
    while (ctas > 1 && (kSmemLimit / ctas - 2048) / stage_bytes < 3) --ctas;
    a.stages = (kSmemLimit / ctas - 2048) / stage_bytes;
    if (a.stages > 8) a.stages = 8;
    PGK_REQUIRE(a.stages >= 1, "pgk_wgrad_tc: stage does not fit in shared memory");
    // split the pixel range so that the grid is (just under) one or two full waves of the SMs
    const int base = sgroups * (Cout / a.NT);
    long long split = 1;
    {
        double best = -1.0;
        for (int w = 1; w <= 2; ++w) {
            long long sp = (long long)w * sms / base;
            if (sp < 1) sp = 1;
            if (sp > max_split) sp = max_split;
            const long long nctas = sp * base;
            const double util = (double)nctas / (double)(((nctas + sms - 1) / sms) * sms);
            if (util > best + 0.02) best = util, split = sp;
        }
    }
    if (split > 65535) split = 65535;
    a.tiles_per_cta = (a.tiles_total + split - 1) / split;
    split = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
    a.dwp = dwp;
    // The bias gradient rides along (idle warps sum the G tiles of the first slab group's CTAs) where that measured a
    // gain: one-plane mode with long reductions (c3 19.90 -> 19.63 ms, c5 25.74 -> 25.34).  In the fp32-faithful mode
    // the extra shared-memory reads slow those CTAs' MMAs by what the separate pass cost (c2: +-0), and on the short
    // reductions of depth 8 / batch 4 it lost 0.4 ms per iteration: there the caller runs pgk_bias_grad.
    const bool fuse_bias = db && bias_mask && Pr == 1 && a.tiles_per_cta >= 64;
    a.db = fuse_bias ? db : nullptr, a.bias_mask = bias_mask;
    if (bias_fused) *bias_fused = fuse_bias ? 1 : 0;

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
    typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, const WgradTcArgs);
    kern_t kern = Pr == 1 ? wgrad_tc_kernel<1> : Pr == 2 ? wgrad_tc_kernel<2> : wgrad_tc_kernel<3>;
    static bool attr_done[3] = {};
    const int ai = Pr - 1;
    if (!attr_done[ai]) {
        cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
        attr_done[ai] = true;
    }
    dim3 grid((unsigned)sgroups, (unsigned)(Cout / a.NT), (unsigned)split);
    pgk_launch(kern, grid, kThreads, smem, (cudaStream_t)stream, tmX, tmG, a);
    PGK_LAUNCH_CHECK("pgk_wgrad(tcgen05)");
    return PGK_OK;
}

extern "C" int pgk_pack_operand_fp16(const float* w, int K, int Nn, void* out, long long out_ps, int P,
                                     pgk_stream_t stream) {
    PGK_REQUIRE(P >= 1 && P <= 2 && K > 0 && Nn > 0, "pgk_pack_operand_fp16: bad arguments");
    dim3 grid((unsigned)((K + 31) / 32), (unsigned)((Nn + 31) / 32));
    pgk_launch(pack_operand_h_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, w, K, Nn, (float)(1 << PGK_FP16_WSHIFT),
                                                                          (__half*)out, out_ps, P);
    PGK_LAUNCH_CHECK("pgk_pack_operand_fp16");
    return PGK_OK;
}

extern "C" int pgk_pack_operand(const float* w, int K, int Nn, void* out, long long out_ps, int P,
                                pgk_stream_t stream) {
    PGK_REQUIRE(P >= 1 && P <= 3 && K > 0 && Nn > 0, "pgk_pack_operand: bad arguments");
    dim3 grid((unsigned)((K + 31) / 32), (unsigned)((Nn + 31) / 32));
    pgk_launch(pack_operand_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, w, K, Nn, make_planes(out, out_ps, P));
    PGK_LAUNCH_CHECK("pgk_pack_operand");
    return PGK_OK;
}

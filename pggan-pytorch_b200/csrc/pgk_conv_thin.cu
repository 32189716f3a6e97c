// pgk_conv_thin.cu -- 3x3 convolution (forward and data gradient) for the thin, high-resolution layers:
// Cin in {8, 16, 32}, Cout in {8, 16, 32, 64}, W a multiple of 128 (the 256^2 ... 1024^2 levels, network.py:94-95 with
// fmap_base 4096).  These layers are HBM-bound (9*Cin*Cout/(Cin+Cout) = 36 ... 190 flop per byte), so the kernel is
// organised around touching every activation byte once:
//
//   * persistent CTAs stream image rows: a work unit is (sample, 128-pixel column strip, RC consecutive rows); the
//     three input rows of an output row live in a shared-memory ring, and moving down one row loads ONE new row
//     (TMA boxes of [130 pixels][8 channels]: out-of-range rows and the +-1 pixel halo are zero filled by the
//     hardware = the conv's padding.  A producer warp of 16-byte cp.async with the same shared-memory layout is kept
//     as a switchable flavour, PGK_THIN_TMA=0: measured equal at one CTA per SM, 5-15 % slower at two);
//   * two CTAs per SM: the tile chain load -> MMA -> commit -> epilogue is latency-bound, and every tcgen05.mma
//     streams its 128-row A tile from shared memory whatever N is ("what bounds the thin kernels");
//   * a row buffer is stored channel-group planar, [Cin/8][130 + pad pixels][8 channels]: one pixel of one channel
//     group is 16 bytes, so 8 consecutive pixels are exactly one un-swizzled K-major UMMA core matrix, and the 9 taps
//     are nothing but 9 start addresses ((dy row buffer) + (1 + dx) * 16 bytes) into the same bytes -- no im2col,
//     no per-tap reload;  K = 16 per MMA is two channel groups (LBO = plane stride) or, for Cin = 8, two
//     neighbouring taps (LBO = 16 bytes);
//   * the whole weight tensor (<= 9*32*64 bf16 per plane) is staged once per CTA in UMMA layout;
//   * accumulator 128 pixels x Npad channels in TMEM, four buffers; two epilogue warpgroups (thread = pixel) apply
//     bias / LeakyReLU / backward mask (prefetched one tile ahead) / scale or the generator's pixel norm, split into
//     planes and store straight to global memory (a thread's pixel row is contiguous with its neighbours': fully
//     coalesced without staging).
#include <stdlib.h>
#include <string.h>

#include "pgk_tc.cuh"

using namespace tc;

namespace {

constexpr int kThinThreads = 320;   // warps 0-3 / 4-7: two epilogue warpgroups (even / odd tiles), 8 producer, 9 MMA issue
constexpr int kAcc = 4;             // TMEM accumulator buffers: hides the MMA -> epilogue -> MMA round trip
constexpr int kMaxRing = 16;        // row buffers: 16, 8 or 4 (power of two), as many as fit
constexpr int kRowPix = 136;        // 130 loaded pixels, padded so that every channel-group plane is 128-byte aligned
constexpr int kCgBytes = kRowPix * 16;
constexpr int kSmemLimit = 227 * 1024;

struct ThinArgs {
    const bf16* x;              // input planes, N,H,W,Cin each, x_ps elements apart
    long long x_ps;
    int N, H, W, Cout, Npad;
    int RC, chunks_y, strips;   // rows per unit, units per column strip, W / 128
    int ring, ring_log2, look;  // row buffers; rows the producer keeps in flight before it hands the oldest over
    int pair;                   // MMA warp interleaves the K steps of two output rows (two accumulators)
    int tma;                    // producer: 1 = TMA boxes (16-byte inner extent), 0 = cp.async chunks
    int spin;                   // producer / MMA warps poll their barriers (1) or suspend in try_wait (0)
    int tstore;                 // epilogue stores through shared memory + TMA (one plane out, Cout >= 16): see the epilogue
    int mload;                  // backward masks loaded warp-coalesced and redistributed through the staging tile
    int dbg;                    // PGK_THIN_DBG knock-outs for stage timing (results are wrong): 1 no MMAs, 2 no stores /
                                // mask loads, 8 no loads
    float* pn_r;                // pixel norm after the activation (NPAD <= 32): per-pixel factor stored here, or NULL
    int total_units;
    int Pout, split_acc;
    const bf16* wpack;          // [P][STEPS][2][Npad][8]
    const float* bias;
    int act, has_mask;
    Planes mask;
    float out_scale;
    Planes out;
};

template <int CIN>
struct Steps {
    static constexpr int CG = CIN / 8;
    static constexpr int N = CIN == 8 ? 6 : 9 * (CIN / 16);
};

// 16-byte read-only load that stays where it is written (volatile asm is not moved across the mbarrier waits)
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// STK = 1 (the one-plane mode): INPUT-ROW-STATIONARY issue.  A tcgen05.mma with M = 128, K = 16 costs ~66 cycles for any
// N <= 64 (tools/probes/umma_probe.cu on B200: 66.5 / 70.0 / 78.2 / 104.6 / 168.6 cycles at N = 16 / 32 / 64 / 128 /
// 256, swizzled or not -- the 128 x 32-byte A tile streams through the shared-memory port whatever N is; 42.8 with A
// in tensor memory), so with one output row per accumulator the 8 -> 8 layer pays 6 x 66 cycles per 128-pixel row
// against ~180 cycles of HBM time.  Here the three filter rows are stacked along N instead: an input row r is
// multiplied ONCE per (dx, channel-group pair) by B = [ky = 2 | ky = 1 | ky = 0] (N = 3 * Npad) and accumulates into the
// three neighbouring accumulator blocks of the output rows r-1, r, r+1 -- a ring of kStkBlocks blocks of Npad columns;
// a window that wraps around the ring is issued as two instructions.  Three times fewer instructions (2 / 3 / 6 per
// row for Cin 8 / 16 / 32), each input row is consumed by its own MMAs alone (released at once), and output row i is
// complete when input row i+2 has been issued.  Every MMA accumulates: the epilogue warps, who own the lanes, zero a
// block (tcgen05.st) after reading it and before handing it back.
template <int NPAD>
struct StkRing {
    static constexpr int BLOCKS = NPAD == 64 ? 4 : 8;
};
constexpr int kAccSlots = 8;        // accumulator barrier pairs laid out in shared memory (kAcc or StkRing::BLOCKS used)
template <int CIN, int NPAD, int SPLIT, int STK>
struct ThinCols {
    static constexpr int ACC = SPLIT ? 2 * NPAD : NPAD;
    static constexpr int NEED = STK ? StkRing<NPAD>::BLOCKS * NPAD : kAcc * ACC;
    static constexpr unsigned N = NEED <= 64 ? 64u : NEED <= 128 ? 128u : NEED <= 256 ? 256u : 512u;
};

template <int CIN, int P, int SPLIT, int NPAD, int STK>
__global__ void __launch_bounds__(kThinThreads, CIN == 64 ? 1 : 2) conv_thin_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmO,
                                                                    const ThinArgs a) {
    static_assert(!STK || (P == 1 && SPLIT == 0), "the input-row-stationary flavour exists for the one-plane mode only");
    constexpr int NACC = STK ? StkRing<NPAD>::BLOCKS : kAcc;   // accumulator blocks in flight
    constexpr int CG = Steps<CIN>::CG, STEPS = Steps<CIN>::N;
    constexpr uint32_t plane_bytes = CG * kCgBytes;
    constexpr uint32_t row_bytes = P * plane_bytes;
    extern __shared__ uint8_t smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 127u) & ~127u;
    const int kRing = a.ring, kRingLog = a.ring_log2;
    const uint32_t rows0 = sbase;                                 // kRing row buffers
    const uint32_t w0 = rows0 + kRing * row_bytes;                // weights [P][STEPS][2][Npad][8]
    constexpr uint32_t wstep = (uint32_t)NPAD * 32u, wplane = STEPS * wstep;
    const uint32_t bars = (w0 + P * wplane + 15u) & ~15u;
    auto rfull = [&](int s) { return bars + 8u * s; };
    auto rempty = [&](int s) { return bars + 8u * (kMaxRing + s); };
    auto afull = [&](int b) { return bars + 16u * kMaxRing + 8u * b; };
    auto aempty = [&](int b) { return bars + 16u * kMaxRing + 8u * kAccSlots + 8u * b; };
    const uint32_t tptr = bars + 16u * kMaxRing + 16u * kAccSlots;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (tptr + 16u - raw));
    const uint32_t stage0 = (tptr + 16u + 4u * NPAD + 1023u) & ~1023u;   // store staging: 8 warps x 32 pixels x Cout bf16

    // ---- one-time setup: zero the row ring (padding pixels must be finite), stage weights and bias
    for (uint32_t o = threadIdx.x * 16u; o < kRing * row_bytes; o += kThinThreads * 16u)
        st_shared_v4(rows0 + o, make_uint4(0, 0, 0, 0));
    pgk_pdl_enter();   // the weight operand and the bias are written by earlier launches of the stream
    if constexpr (STK != 0) {
        // the packed operand is [step = (ky, dx or dx-pair, channel-group pair)][K half][Npad][8]; staged here as
        // [k step = (dx..., channel-group pair)][K half][ky = 2, 1, 0][Npad][8], i.e. the three filter rows side by side
        // along N in the order of the output rows they feed (r-1, r, r+1)
        const uint4* src = reinterpret_cast<const uint4*>(a.wpack);
        for (uint32_t i = threadIdx.x; i < (uint32_t)(STEPS * 2 * NPAD); i += kThinThreads) {
            const uint32_t n = i % NPAD, rr = i / NPAD, h = rr & 1u, st = rr >> 1;
            uint32_t dy, ks;
            if constexpr (CIN == 8) {
                dy = st >> 1, ks = st & 1u;
            } else {
                const uint32_t tap = st / (CIN / 16), cgp = st % (CIN / 16);
                dy = tap / 3, ks = (tap % 3) * (CIN / 16) + cgp;
            }
            st_shared_v4(w0 + ((((ks * 2 + h) * 3 + (2 - dy)) * NPAD) + n) * 16u, __ldg(src + i));
        }
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(a.wpack);
        const uint32_t n16 = P * wplane / 16u;
        for (uint32_t i = threadIdx.x; i < n16; i += kThinThreads) st_shared_v4(w0 + i * 16u, __ldg(src + i));
    }
    for (int i = threadIdx.x; i < NPAD; i += kThinThreads) bias_s[i] = (a.bias && i < a.Cout) ? __ldg(a.bias + i) : 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kRing; ++s) {
            mbar_init(rfull(s), 1);
            mbar_init(rempty(s), 1);
        }
        for (int b = 0; b < NACC; ++b) {
            mbar_init(afull(b), 1);
            mbar_init(aempty(b), 4);
        }
        fence_barrier_init();
    }
    constexpr int acc_cols = ThinCols<CIN, NPAD, SPLIT, STK>::ACC;
    constexpr unsigned ncols = ThinCols<CIN, NPAD, SPLIT, STK>::N;
    if (warp == 9) tmem_alloc(tptr, ncols);
    fence_proxy_async();   // generic-proxy writes (weights, zeroed ring) -> visible to the tensor core / TMA
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
    if constexpr (STK != 0) {
        // every MMA of this flavour accumulates: the accumulator ring starts out as zeros (warps 0-3 own the four
        // lane quarters) and each block is zeroed again by the epilogue that drains it
        if (warp < 4) {
            const uint32_t t0 = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int c = 0; c < NACC * NPAD; c += 8) tmem_zero8(t0 + c);
            tmem_st_wait();
        }
        fence_before();
        __syncthreads();
        fence_after();
    }

    // The two single-warp roles share their SM sub-partitions with epilogue warps: busy polling (test_wait) costs
    // those warps issue slots (12 % of all issued instructions in the ncu capture), try_wait suspends instead.
    auto wait_bar = [&](uint32_t bar, uint32_t parity) {
        if (a.spin) mbar_wait_spin(bar, parity);
        else mbar_wait(bar, parity);
    };
    auto unit_coords = [&](int u, int& n, int& x0, int& ya) {
        const int cy = u % a.chunks_y;
        int r = u / a.chunks_y;
        const int st = r % a.strips;
        n = r / a.strips;
        x0 = st * 128;
        ya = cy * a.RC;
    };

    if (warp == 8 && a.tma) {
        // ---- producer, TMA flavour: per row and plane CG boxes of [130 pixels][8 channels]; halo and out-of-image
        // rows are zero filled by the hardware
        constexpr uint32_t tx_bytes = P * CG * 130 * 16;
        if (lane == 0) tma_prefetch_desc(&tmA);
        uint32_t g = 0;
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            int n, x0, ya;
            unit_coords(u, n, x0, ya);
            for (int j = 0; j < a.RC + 2; ++j, ++g) {
                const int s = g & (kRing - 1);
                wait_bar(rempty(s), ((g >> kRingLog) & 1) ^ 1);
                if (elect_one()) {
                    const uint32_t fb = rfull(s);
                    if (a.dbg & 8) mbar_arrive(fb);
                    else {
                    mbar_expect_tx(fb, tx_bytes);
                    const uint32_t dst = rows0 + s * row_bytes;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
#pragma unroll
                        for (int cg = 0; cg < CG; ++cg)
                            tma_load_5d(dst + p * plane_bytes + cg * kCgBytes, &tmA, fb, cg * 8, x0 - 1, ya - 1 + j, n, p);
                    }
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 8) {
        // ---- producer: one input row per step of the ring, as 130 * CG 16-byte chunks (pixel, channel group) per
        // plane.  Chunk q of a row sits at byte 16 * q of the global row segment, so a warp instruction reads 512
        // contiguous bytes; its shared-memory home is [channel group][pixel].  Up to `look` rows stay in flight
        // (one cp.async group per row); a row is handed to the MMA warp once every lane's copies of it have landed
        // and been fenced towards the async proxy.
        uint32_t g = 0, signalled = 0;   // rows requested / rows handed over (warp-uniform)
        auto hand_over = [&](uint32_t upto) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
                for (uint32_t r = signalled; r < upto; ++r) mbar_arrive(rfull(r & (kRing - 1)));
            signalled = upto;
        };
        const int look = a.look;
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            int n, x0, ya;
            unit_coords(u, n, x0, ya);
            for (int j = 0; j < a.RC + 2; ++j, ++g) {
                const int s = g & (kRing - 1);
                const uint32_t par = ((g >> kRingLog) & 1) ^ 1;
                if (!__all_sync(0xffffffffu, mbar_test(rempty(s), par))) {
                    // about to sleep on a busy slot: hand over everything requested so far first -- the MMA warp may
                    // need exactly those rows to finish the tile that frees this slot
                    cp_async_wait<0>();
                    hand_over(g);
                    wait_bar(rempty(s), par);
                }
                const int y = ya - 1 + j;
                const bool row_ok = y >= 0 && y < a.H;
                const bf16* rowp = a.x + ((long long)n * a.H + (row_ok ? y : 0)) * a.W * CIN;
                const uint32_t dst = rows0 + s * row_bytes;
#pragma unroll
                for (int p = 0; p < P; ++p) {
#pragma unroll
                    for (int it = 0; it < (130 * CG + 31) / 32; ++it) {
                        const int q = it * 32 + lane;
                        if (q < 130 * CG) {
                            const int px = q / CG, cg = q % CG;
                            const int xx = x0 - 1 + px;
                            const bool ok = row_ok && xx >= 0 && xx < a.W;
                            const bf16* src = rowp + (long long)p * a.x_ps + (long long)(ok ? xx : 0) * CIN + cg * 8;
                            cp_async16(dst + p * plane_bytes + cg * kCgBytes + px * 16, src, ok ? 16u : 0u);
                        }
                    }
                }
                cp_async_commit();
                if ((int)(g + 1 - signalled) > look) {   // rows <= g - look have landed for this lane
                    if (look >= 12) cp_async_wait<12>();
                    else if (look >= 4) cp_async_wait<4>();
                    else if (look >= 1) cp_async_wait<1>();
                    else cp_async_wait<0>();
                    const int lk = look >= 12 ? 12 : look >= 4 ? 4 : look >= 1 ? 1 : 0;
                    hand_over(g + 1 - lk);
                }
            }
        }
        cp_async_wait<0>();
        hand_over(g);
    } else if (warp == 9 && STK != 0) {
        // ---- MMA issue, input-row stationary (see the kernel header)
        constexpr int KSTEPS = STEPS / 3;
        constexpr uint32_t kstep16 = (uint32_t)(6 * NPAD);            // one k step of the staged weights, in 16-byte units
        const uint64_t adesc_hi = smem_desc(0, CIN == 8 ? 16u : (uint32_t)kCgBytes, 128, 0);
        const uint64_t bdesc0 = smem_desc(w0, (uint32_t)(3 * NPAD) * 16u, 128, 0);
        const uint32_t idesc0 = idesc_bf16(0, 0, 0);                  // N is filled in per instruction: (N >> 3) << 17
        auto idesc_n = [&](int nb) { return idesc0 | ((uint32_t)(nb * NPAD >> 3) << 17); };
        uint32_t g = 0, ti = 0;
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            for (int j = 0; j < a.RC + 2; ++j, ++g) {
                const int s = g & (kRing - 1);
                wait_bar(rfull(s), (g >> kRingLog) & 1);
                if (j < a.RC) {      // the block of output row j is touched for the first time: its last reader is done
                    const uint32_t tn = ti + (uint32_t)j;
                    wait_bar(aempty(tn % NACC), ((tn / NACC) & 1) ^ 1);
                }
                fence_after();
                // output rows [i_lo, i_hi] take this input row through ky = j - i; their blocks are neighbours in the ring
                const int i_lo = j >= 2 ? j - 2 : 0, i_hi = j < a.RC ? j : a.RC - 1;
                const int nb = i_hi - i_lo + 1;
                const uint32_t kb0 = (uint32_t)(2 - j + i_lo);                   // first filter-row block of B
                const uint32_t blk0 = (ti + (uint32_t)i_lo) % NACC;
                const int nb_a = (int)blk0 + nb <= NACC ? nb : NACC - (int)blk0;  // blocks before the ring wraps
                if (elect_one()) {
                    const uint32_t rowa = (rows0 + s * row_bytes) >> 4;
#pragma unroll
                    for (int ks = 0; ks < KSTEPS && !(a.dbg & 1); ++ks) {
                        uint32_t aoff;
                        if constexpr (CIN == 8) {
                            aoff = (uint32_t)(ks * 2 * 16) / 16u;                 // (dx -1, dx 0) | (dx +1, zero weights)
                        } else {
                            const int dx = ks / (CIN / 16), cgp = ks % (CIN / 16);
                            aoff = (uint32_t)(2 * cgp * kCgBytes + dx * 16) / 16u;
                        }
                        const uint64_t ad = adesc_hi | (uint64_t)(rowa + aoff);
                        const uint64_t bd = bdesc0 + (uint32_t)(ks * kstep16 + kb0 * NPAD);
                        mma_bf16(tmem + blk0 * NPAD, ad, bd, idesc_n(nb_a), 1u);
                        if (nb_a < nb)
                            mma_bf16(tmem, ad, bd + (uint32_t)(nb_a * NPAD), idesc_n(nb - nb_a), 1u);
                    }
                    mma_commit(rempty(s));                                        // this input row is consumed
                    if (j >= 2) mma_commit(afull((ti + (uint32_t)(j - 2)) % NACC));   // output row j - 2 is complete
                }
                __syncwarp();
            }
            ti += (uint32_t)a.RC;
        }
    } else if (warp == 9) {
        // ---- MMA issue: per output row, STEPS x (products of planes) MMAs of 128 pixels x Npad x 16
        const uint32_t idesc = idesc_bf16(NPAD, 0, 0);
        // A: 8 pixels x 16 B core matrices, SBO = 128 B to the next 8 pixels; LBO = distance between the two K halves
        const uint64_t adesc_hi = smem_desc(0, CIN == 8 ? 16u : (uint32_t)kCgBytes, 128, 0);
        const uint64_t bdesc0 = smem_desc(w0, (uint32_t)NPAD * 16u, 128, 0);
        constexpr uint32_t wstep16 = wstep >> 4, wplane16 = wplane >> 4;
        uint32_t g = 0, ti = 0;
        // one K = 16 step of output row `row` (window rows rb[0..2]) into accumulator buffer b
        auto issue_step = [&](int st, const uint32_t* rb, int b) {
            const uint32_t d_main = tmem + b * acc_cols, d_corr = d_main + NPAD;
            int dy, xoff, cgp;
            if (CIN == 8) {
                dy = st >> 1, xoff = (st & 1) * 2, cgp = 0;      // (dx -1, dx 0) | (dx +1, zero weights)
            } else {
                const int tap = st / (CIN / 16);
                cgp = st % (CIN / 16);
                dy = tap / 3, xoff = tap % 3;
            }
#pragma unroll
            for (int pi = 0; pi < P; ++pi) {
#pragma unroll
                for (int pj = 0; pj < P - pi; ++pj) {
                    const uint64_t ad = adesc_hi | (uint64_t)(rb[dy] + (pi * plane_bytes + 2 * cgp * kCgBytes + xoff * 16) / 16);
                    const uint64_t bd = bdesc0 + (uint32_t)(pj * wplane16 + st * wstep16);
                    if (pi + pj == 0 || !SPLIT)
                        mma_bf16(d_main, ad, bd, idesc, (st == 0 && pi + pj == 0) ? 0u : 1u);
                    else
                        mma_bf16(d_corr, ad, bd, idesc, (st == 0 && pi == 0 && pj == 1) ? 0u : 1u);
                }
            }
        };
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            wait_bar(rfull(g & (kRing - 1)), (g >> kRingLog) & 1);
            wait_bar(rfull((g + 1) & (kRing - 1)), ((g + 1) >> kRingLog) & 1);
            if (a.pair) {
                // two output rows at a time: their MMA chains accumulate into different TMEM buffers, so the tensor
                // core can overlap the (latency-bound, N <= 64) steps of one row with those of the other
                for (int i = 0; i < a.RC; i += 2, g += 2, ti += 2) {
                    wait_bar(rfull((g + 2) & (kRing - 1)), ((g + 2) >> kRingLog) & 1);
                    wait_bar(rfull((g + 3) & (kRing - 1)), ((g + 3) >> kRingLog) & 1);
                    const int b0 = ti & (kAcc - 1), b1 = (ti + 1) & (kAcc - 1);
                    wait_bar(aempty(b0), ((ti / kAcc) & 1) ^ 1);
                    wait_bar(aempty(b1), (((ti + 1) / kAcc) & 1) ^ 1);
                    fence_after();
                    uint32_t rb[4];
#pragma unroll
                    for (int dy = 0; dy < 4; ++dy) rb[dy] = (rows0 + ((g + dy) & (kRing - 1)) * row_bytes) >> 4;
                    if (elect_one()) {
#pragma unroll
                        for (int st = 0; st < STEPS; ++st) {
                            issue_step(st, rb, b0);
                            issue_step(st, rb + 1, b1);
                        }
                        mma_commit(afull(b0));
                        mma_commit(afull(b1));
                        mma_commit(rempty(g & (kRing - 1)));
                        mma_commit(rempty((g + 1) & (kRing - 1)));
                        if (i == a.RC - 2) {
                            mma_commit(rempty((g + 2) & (kRing - 1)));
                            mma_commit(rempty((g + 3) & (kRing - 1)));
                        }
                    }
                    __syncwarp();
                }
            } else {
                for (int i = 0; i < a.RC; ++i, ++g, ++ti) {
                    wait_bar(rfull((g + 2) & (kRing - 1)), ((g + 2) >> kRingLog) & 1);
                    const int b = ti & (kAcc - 1);
                    wait_bar(aempty(b), ((ti / kAcc) & 1) ^ 1);
                    fence_after();
                    uint32_t rb[3];
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) rb[dy] = (rows0 + ((g + dy) & (kRing - 1)) * row_bytes) >> 4;
                    if (elect_one()) {
#pragma unroll
                        for (int st = 0; st < STEPS; ++st) issue_step(st, rb, b);
                        mma_commit(afull(b));
                        mma_commit(rempty(g & (kRing - 1)));   // the top row of this window is done
                        if (i == a.RC - 1) {
                            mma_commit(rempty((g + 1) & (kRing - 1)));
                            mma_commit(rempty((g + 2) & (kRing - 1)));
                        }
                    }
                    __syncwarp();
                }
            }
            g += 2;
        }
    } else {
        // ---- epilogue: thread = pixel; warpgroup wg takes the tiles with (tile index & 1) == wg.  Tiles are numbered
        // over the whole life of the CTA (tile = unit index * RC + row), exactly as the MMA warp counts them.  The
        // backward-mask vectors of a thread's NEXT tile are requested before it waits for the current accumulator, so
        // their HBM latency is hidden behind a whole tile of work.
        constexpr int NV = NPAD / 8;
        const int wg = warp >> 2, q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t trow_off = (uint32_t)(q * 32) << 16;
        const int my_units = (int)blockIdx.x < a.total_units ? (a.total_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const uint32_t ntiles = (uint32_t)my_units * (uint32_t)a.RC;
        // pixel index of this thread in tile (unit k of this CTA, row i); the walk below advances (k, i) two rows at
        // a time (RC is even or the warpgroups alternate units), so the divisions happen once per unit, not per tile
        auto unit_pix0 = [&](uint32_t k) {
            int n, x0, ya;
            unit_coords((int)(blockIdx.x + k * gridDim.x), n, x0, ya);
            return ((long long)n * a.H + ya) * a.W + x0 + r;
        };
        uint4 mk[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) mk[k] = make_uint4(0, 0, 0, 0);
        auto load_mask = [&](long long pix) {   // plane 0 only: its sign is the sign of the value
            if (a.mload) {
                // the warp's 32 pixels x Cout channels are 32 * Cout * 2 contiguous bytes: lane takes the 16-byte chunks
                // lane, lane + 32, ... (full lines per instruction); they reach their pixel's thread through the
                // staging tile when the tile is processed (below)
                const uint4* mp = reinterpret_cast<const uint4*>(a.mask.p + (pix - lane) * a.Cout) + lane;
#pragma unroll
                for (int k = 0; k < NV; ++k)
                    if (!(a.dbg & 2)) mk[k] = ldg_nc_v4(mp + k * 32);
                return;
            }
            const uint4* mp = reinterpret_cast<const uint4*>(a.mask.p + pix * a.Cout);
#pragma unroll
            for (int k = 0; k < NV; ++k)
                if (k * 8 < a.Cout && !(a.dbg & 2)) mk[k] = ldg_nc_v4(mp + k);
        };
        const float s_pos = a.out_scale, s_neg = a.act ? PGK_LRELU * a.out_scale : a.out_scale;
        // tstore (one output plane, Cout >= 16): a lane's Cout * 2 bytes are not contiguous with its neighbour's in one
        // 16-byte store instruction (stride Cout * 2): every such instruction occupies the LSU for 32 partial sectors,
        // measured as about half of the kernel at Cout >= 32 (PGK_THIN_DBG=2).  Instead the warp's 32 pixels x Cout
        // channels go to a staging tile in shared memory, written in the 32 / 64 / 128-byte swizzle pattern of the
        // output tensor map (conflict-free 16-byte stores), and leave as ONE bulk tensor store per warp and tile.
        const uint32_t cb = (uint32_t)a.Cout * 2u;
        const uint32_t stg = stage0 + (uint32_t)warp * 32u * cb, my_row = stg + (uint32_t)lane * cb;
        const uint32_t my_sw = (my_row >> 7) & (cb / 16u - 1u);
        auto emit8 = [&](long long o, int h, const float* f) {   // channels 8h .. 8h+7 of this thread's pixel
            if (a.tstore) {
                uint4 q;
                q.x = pack2(f[0], f[1]), q.y = pack2(f[2], f[3]), q.z = pack2(f[4], f[5]), q.w = pack2(f[6], f[7]);
                st_shared_v4(my_row + (((uint32_t)h ^ my_sw) << 4), q);
            } else if (!(a.dbg & 2)) {
                split_store8(a.out, o + 8 * h, f);
            }
        };
        uint32_t ti = (uint32_t)wg;
        // (uk, ui): unit and row of the NEXT tile this warpgroup will prefetch for; base = pixel of row 0 of unit uk
        uint32_t uk = ti / (uint32_t)a.RC, ui = ti - uk * (uint32_t)a.RC;
        long long base = ti < ntiles ? unit_pix0(uk) : 0;
        long long pix = base + (long long)ui * a.W;
        if (ti < ntiles && a.has_mask) load_mask(pix);
        for (; ti < ntiles; ti += 2) {
            uint4 mc[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) mc[k] = mk[k];
            const long long pcur = pix, o = pix * a.Cout;
            if (ti + 2 < ntiles) {
                ui += 2;
                if (ui >= (uint32_t)a.RC) {
                    ui -= (uint32_t)a.RC;
                    ++uk;
                    base = unit_pix0(uk);
                }
                pix = base + (long long)ui * a.W;
                if (a.has_mask) load_mask(pix);
            }
            const int b = (int)(ti % NACC);
            mbar_wait(afull(b), (ti / NACC) & 1);
            fence_after();
            const uint32_t trow = tmem + b * acc_cols + trow_off;
            if (a.tstore) {   // the previous tile's bulk store has finished reading the staging tile
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
            }
            if (a.mload) {
                // chunk c = 32 k + lane of the warp's block belongs to pixel c / NV, channels 8 (c % NV) ..: written in
                // the tile's swizzle (512 contiguous bytes per instruction), read back by the pixel's own thread
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const uint32_t c = (uint32_t)(k * 32 + lane), row = stg + (c / NV) * cb;
                    st_shared_v4(row + (((c % NV) ^ ((row >> 7) & (cb / 16u - 1u))) << 4), mc[k]);
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < NV; ++k) mc[k] = ld_shared_v4(my_row + (((uint32_t)k ^ my_sw) << 4));
                __syncwarp();
            }
            bool done_pn = false;
            if constexpr (NPAD <= 32) {
                if (a.pn_r) {
                    // generator layers: bias, LeakyReLU, then the pixel norm over this pixel's Cout channels
                    // (network.py:37-40) -- the thread holds all of them
                    float v[NPAD];
#pragma unroll
                    for (int c = 0; c < NPAD; c += 16) {
                        tmem_ld16(trow + c, v + c);
                        if (SPLIT) {
                            float w[16];
                            tmem_ld16(trow + NPAD + c, w);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[c + j] += w[j];
                        }
                    }
                    float ss = 0.f;
#pragma unroll
                    for (int j = 0; j < NPAD; ++j) {
                        float f = v[j] + bias_s[j];
                        f *= f > 0.f ? s_pos : s_neg;
                        f = j < a.Cout ? f : 0.f;
                        v[j] = f;
                        ss = fmaf(f, f, ss);
                    }
                    const float rr = rsqrtf(ss / (float)a.Cout + 1e-8f);
                    a.pn_r[pcur] = rr;
#pragma unroll
                    for (int h = 0; h < NPAD / 8; ++h) {
                        if (8 * h < a.Cout) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[8 * h + j] *= rr;
                            emit8(o, h, v + 8 * h);
                        }
                    }
                    done_pn = true;
                }
            }
            if (!done_pn) {
#pragma unroll
            for (int c = 0; c < NPAD; c += 16) {
                float v[16];
                tmem_ld16(trow + c, v);
                if (SPLIT) {
                    float w[16];
                    tmem_ld16(trow + NPAD + c, w);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += w[j];
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (c + 8 * h < a.Cout) {
                        float* f = v + 8 * h;
                        {
                            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c + 8 * h);
                            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c + 8 * h + 4);
                            f[0] += b0.x, f[1] += b0.y, f[2] += b0.z, f[3] += b0.w;
                            f[4] += b1.x, f[5] += b1.y, f[6] += b1.z, f[7] += b1.w;
                        }
                        // LeakyReLU and the output scale as one select + multiply
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] *= f[j] > 0.f ? s_pos : s_neg;
                        if (a.has_mask) {
                            float m[8];
                            unpack8(mc[c / 8 + h], m);
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] *= lrelu_grad(m[j]);
                        }
                        emit8(o, c / 8 + h, f);
                    }
                }
            }
            }
            if (a.tstore) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && !(a.dbg & 2)) {
                    tma_store_2d(&tmO, stg, 0, (int)(pcur - lane));
                    tma_store_commit();
                }
            }
            if constexpr (STK != 0) {
                // hand the block back zeroed: the next output row that lands here accumulates from its first MMA on
#pragma unroll
                for (int c = 0; c < NPAD; c += 8) tmem_zero8(trow + c);
                tmem_st_wait();
            }
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(aempty(b));
        }
        if (a.tstore && lane == 0) tma_store_wait_all();
    }
    fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem, ncols);
}

// out[p][step][khalf][n][e] for the K = 16 steps of conv_thin_kernel; w = fp32 [9*Cin][Cout] (pgk_prep_weight's wf / wb)
__global__ void pack_thin_kernel(const float* __restrict__ w, int Cin, int Cout, int Npad, int steps, Planes out) {
    pgk_pdl_enter();
    const int total = steps * 2 * Npad * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i & 7;
        int r = i >> 3;
        const int n = r % Npad;
        r /= Npad;
        const int h = r & 1, st = r >> 1;
        int k = -1;
        if (Cin == 8) {
            const int dy = st >> 1, q = st & 1;
            const int dx = q * 2 + h;
            if (dx < 3) k = (dy * 3 + dx) * 8 + e;
        } else {
            const int per_tap = Cin / 16;
            const int tap = st / per_tap, cgp = st % per_tap;
            k = tap * Cin + cgp * 16 + h * 8 + e;
        }
        const float v = (k >= 0 && n < Cout) ? w[(long long)k * Cout + n] : 0.f;
        st1(out, i, v);
    }
}

}  // namespace

extern "C" int pgk_conv_thin_supported(int N, int H, int W, int Cin, int Cout, int KS, int ups) {
    if (ups || KS != 3) return 0;
    if (Cin != 8 && Cin != 16 && Cin != 32) return 0;
    if (Cout != 8 && Cout != 16 && Cout != 32 && Cout != 64) return 0;
    if (W % 128 || H < 8 || H % 8) return 0;
    return N > 0;
}

// number of bf16 elements of one plane of the packed operand
extern "C" long long pgk_pack_thin_plane_elems(int Cin, int Cout) {
    const int npad = Cout < 16 ? 16 : Cout;
    const int steps = Cin == 8 ? 6 : 9 * (Cin / 16);
    return (long long)steps * 2 * npad * 8;
}

extern "C" int pgk_pack_thin(const float* w, int Cin, int Cout, void* out, long long out_ps, int P,
                             pgk_stream_t stream) {
    PGK_REQUIRE((Cin == 8 || Cin == 16 || Cin == 32 || Cin == 64) && Cout % 8 == 0 && Cout >= 8 && Cout <= 64 && P >= 1 && P <= 3,
                "pgk_pack_thin: unsupported shape (Cin %d Cout %d)", Cin, Cout);
    const int npad = Cout < 16 ? 16 : Cout;
    const int steps = Cin == 8 ? 6 : 9 * (Cin / 16);
    const int total = steps * 2 * npad * 8;
    pgk_launch(pack_thin_kernel, dim3((total + 255) / 256), 256, 0, (cudaStream_t)stream, w, Cin, Cout, npad, steps,
                                                                           make_planes(out, out_ps, P));
    PGK_LAUNCH_CHECK("pgk_pack_thin");
    return PGK_OK;
}

// CTAs per SM: the kernel is bound by memory latency (one row of 128 pixels per pipeline step), so co-resident CTAs
// are what fills the HBM pipe.  Limits: shared memory + registers (asked of the runtime), 512 TMEM columns per SM,
// and the PGK_THIN_OCC knob (default 2).
static int thin_occ_cap() {
    static int cap = 0;
    if (cap == 0) {
        const char* e = getenv("PGK_THIN_OCC");
        cap = e ? atoi(e) : 2;
        if (cap < 1) cap = 1;
        if (cap > 4) cap = 4;
    }
    return cap;
}

struct ThinPlan {
    int occ, ring, smem;
};

template <int CIN, int P, int SPLIT, int NPAD, int STK>
static int launch_thin(const CUtensorMap& tmA, const CUtensorMap& tmO, ThinArgs& a, cudaStream_t stream) {
    static bool attr = false;
    static ThinPlan plan;   // CTAs per SM, ring depth and shared memory of this instance
    auto kern = conv_thin_kernel<CIN, P, SPLIT, NPAD, STK>;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_conv_thin: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
        constexpr int steps = Steps<CIN>::N;
        constexpr int ncols = (int)ThinCols<CIN, NPAD, SPLIT, STK>::N;
        const int fixed = 128 + P * steps * NPAD * 32 + 16 + 16 * kMaxRing + 16 * kAccSlots + 32 + 4 * NPAD + 64 +
                          (P == 1 ? 1024 + 8 * 32 * NPAD * 2 : 0);   // (+ the store staging tiles of the one-plane mode)
        const int row = P * (CIN / 8) * kCgBytes;
        ThinPlan pl = {0, 0, 0};
        const int cap = thin_occ_cap();
        // preference: CTAs per SM first (two tiles in the epilogue / MMA pipe per SM), then ring depth (rows in flight).
        // Residency is computed here -- shared memory (+1 KB reserved per CTA) against the 228 KB of an SM, 512 TMEM
        // columns, and registers through __launch_bounds__(kThinThreads, 2).
        const int occ_max = cap < 512 / ncols ? cap : 512 / ncols;
        // (a ring of 4 rows with two CTAs per SM keeps as many rows in flight per SM as 8 rows with one, and the second
        // CTA overlaps epilogues with MMAs; the output-row-stationary flavours need 8 rows for their MMA window + look-ahead)
        const int min_ring2 = STK ? 4 : 8;
        for (int pass = 0; pass < 2 && pl.occ == 0; ++pass) {
            for (int occ = occ_max > 2 ? 2 : occ_max; occ >= 1 && pl.occ == 0; --occ) {
                const int lo = pass == 0 ? (occ >= 2 ? min_ring2 : 8) : 4;
                for (int ring = pass == 0 ? 16 : 4; ring >= lo && pl.occ == 0; ring >>= 1) {
                    const int smem = fixed + ring * row;
                    if (smem > kSmemLimit || occ * (smem + 1024) > 228 * 1024) continue;
                    pl.occ = occ, pl.ring = ring, pl.smem = smem;
                }
            }
        }
        if (getenv("PGK_THIN_DEBUG")) {
            int got = -1;
            cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&got, kern, kThinThreads, pl.smem);
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, kern);
            fprintf(stderr, "pgk_conv_thin<%d,%d,%d,%d,%d>: plan occ %d ring %d smem %d | runtime says %d blocks/SM (%s), regs %d, "
                            "static smem %zu, max dyn %d\n", CIN, P, SPLIT, NPAD, STK, pl.occ, pl.ring, pl.smem, got,
                    cudaGetErrorString(oe), fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
            cudaGetLastError();
        }
        if (pl.occ == 0) {
            cudaGetLastError();
            pgk_set_error("pgk_conv_thin: no shared-memory plan for Cin %d Npad %d P %d", CIN, NPAD, P);
            return PGK_ERR_ARG;
        }
        plan = pl;
        attr = true;
    }
    // rows per unit: minimise the rows of the busiest CTA, ceil(units / grid) * (RC + 2 halo rows), over the chunk
    // sizes that divide H
    const ThinPlan best_pl = plan;
    int best_rc = 0, best_grid = 0;
    long long best_cost = -1;
    for (int rc = 64; rc >= 8; rc >>= 1) {
        if (a.H % rc) continue;
        const long long units = (long long)a.N * a.strips * (a.H / rc);
        long long grid = (long long)best_pl.occ * pgk_num_sms();
        if (grid > units) grid = units;
        const long long cost = (units + grid - 1) / grid * (rc + 2);
        if (best_cost < 0 || cost < best_cost) best_cost = cost, best_rc = rc, best_grid = (int)grid;
    }
    if (best_rc == 0) {
        pgk_set_error("pgk_conv_thin: no row chunk divides H = %d", a.H);
        return PGK_ERR_ARG;
    }
    a.RC = best_rc;
    a.chunks_y = a.H / a.RC;
    a.total_units = a.N * a.strips * a.chunks_y;
    a.ring = best_pl.ring;
    a.ring_log2 = best_pl.ring == 16 ? 4 : best_pl.ring == 8 ? 3 : 2;
    a.look = best_pl.ring == 16 ? 12 : best_pl.ring == 8 ? 4 : 1;   // <= ring - 4: the MMA window holds up to four rows
    {
        static int pair = -1;
        if (pair < 0) {
            const char* e = getenv("PGK_THIN_PAIR");
            pair = e ? atoi(e) != 0 : 1;
        }
        a.pair = !STK && pair && best_pl.ring >= 8 && best_rc % 2 == 0;
    }
    if (P != 1) a.tstore = a.mload = 0;   // (staging tiles are planned for the one-plane instances only)
    pgk_launch(kern, best_grid, kThinThreads, best_pl.smem, stream, tmA, tmO, a);
    return PGK_OK;
}

// wpack: pgk_pack_thin output with 3 planes, wpack_ps elements apart
// the epilogue applies the pixel norm itself when one thread holds every channel of its pixel in registers
extern "C" int pgk_conv_thin_fuses_pixelnorm(int Cout) { return Cout <= 32; }

extern "C" int pgk_conv_thin(const void* x, int P, int Pr, long long x_ps, int N, int H, int W, int Cin, int Cout,
                             const void* wpack, long long wpack_ps, const float* bias, int act, const void* mask_ref,
                             long long mask_ps, float out_scale, void* out, long long out_ps, float* pn_r,
                             pgk_stream_t stream) {
    PGK_REQUIRE(!pn_r || (pgk_conv_thin_fuses_pixelnorm(Cout) && !mask_ref && out_scale == 1.0f),
                "pgk_conv_thin: the fused pixel norm needs Cout <= 32, no mask and out_scale 1");
    // 64 input channels (one-plane mode, Cout 32 / 64): not part of pgk_conv's dispatch -- those layers belong to the wide
    // kernel's channel counts, where a tcgen05.mma of N = 32 ... 64 runs at a quarter of the tensor rate (c3: 0.26-0.44 of
    // the peak); the row-streaming kernel stacks the three filter rows along N and is bound by HBM instead.  The caller
    // asks for it explicitly with an operand packed by pgk_pack_thin.
    const bool wide64 = Cin == 64 && P == 1 && Pr == 1 && (Cout == 32 || Cout == 64) && W % 128 == 0 && H >= 8 && H % 8 == 0 && N > 0;
    PGK_REQUIRE(wide64 || pgk_conv_thin_supported(N, H, W, Cin, Cout, 3, 0), "pgk_conv_thin: unsupported shape");
    PGK_REQUIRE(P >= 1 && P <= 3 && Pr >= 1 && Pr <= P, "pgk_conv_thin: need 1 <= Pr <= P <= 3");
    ThinArgs a;
    a.N = N, a.H = H, a.W = W, a.Cout = Cout;
    a.Npad = Cout < 16 ? 16 : Cout;
    a.strips = W / 128;
    a.RC = a.chunks_y = a.total_units = a.ring = a.ring_log2 = a.look = 0;   // chosen by launch_thin
    a.Pout = P;
    a.split_acc = Pr == 3 ? 1 : 0;
    PGK_REQUIRE(wpack_ps == pgk_pack_thin_plane_elems(Cin, Cout), "pgk_conv_thin: wpack plane stride mismatch");
    a.wpack = (const bf16*)wpack;
    a.bias = bias, a.act = act;
    a.has_mask = mask_ref != nullptr;
    a.mask = make_planes(mask_ref, mask_ps, P);
    a.out_scale = out_scale;
    a.out = make_planes(out, out_ps, P);
    a.pn_r = pn_r;
    a.x = (const bf16*)x;
    a.x_ps = x_ps;
    PGK_REQUIRE((((uintptr_t)x) & 15) == 0 && (P == 1 || (x_ps * 2) % 16 == 0), "pgk_conv_thin: x must be 16-byte aligned");
    static int use_tma = -1;
    if (use_tma < 0) {
        const char* e = getenv("PGK_THIN_TMA");
        use_tma = e ? atoi(e) != 0 : 1;
    }
    a.tma = use_tma;
    {
        const char* e = getenv("PGK_THIN_DBG");
        a.dbg = e ? atoi(e) : 0;
    }
    static int spin = -1;
    if (spin < 0) {
        const char* e = getenv("PGK_THIN_SPIN");
        spin = e ? atoi(e) != 0 : 0;
    }
    a.spin = spin;
    CUtensorMap tmA;
    memset(&tmA, 0, sizeof(tmA));
    if (a.tma) {
        unsigned long long dims[5] = {(unsigned long long)Cin, (unsigned long long)W, (unsigned long long)H,
                                      (unsigned long long)N, (unsigned long long)P};
        unsigned long long str[4] = {2ull * Cin, 2ull * Cin * W, 2ull * Cin * W * H,
                                     P > 1 ? 2ull * x_ps : 2ull * Cin * W * H * N};
        unsigned box[5] = {8u, 130u, 1u, 1u, 1u};
        int rc = pgk_make_tmap(&tmA, x, 5, dims, str, box, 0, "pgk_conv_thin(x)");
        if (rc) return rc;
    }
    // the output as a 2-D tensor [pixel][Cout] for the epilogue's bulk stores (box = one warp's 32 pixels)
    CUtensorMap tmO;
    memset(&tmO, 0, sizeof(tmO));
    static int use_tstore = -1;
    if (use_tstore < 0) {
        const char* e = getenv("PGK_THIN_TSTORE");
        use_tstore = e ? atoi(e) != 0 : 1;
    }
    static int use_mload = -1;
    if (use_mload < 0) {
        const char* e = getenv("PGK_THIN_MLOAD");
        use_mload = e ? atoi(e) != 0 : 1;
    }
    a.tstore = use_tstore && P == 1 && Pr == 1 && Cout >= 16 && (long long)N * H * W < (1ll << 31) && (((uintptr_t)out) & 15) == 0;
    a.mload = a.tstore && use_mload && mask_ref != nullptr && (((uintptr_t)mask_ref) & 15) == 0;
    if (a.tstore) {
        unsigned long long dims[2] = {(unsigned long long)Cout, (unsigned long long)N * H * W};
        unsigned long long str[1] = {2ull * Cout};
        unsigned box[2] = {(unsigned)Cout, 32u};
        int rc = pgk_make_tmap(&tmO, out, 2, dims, str, box, 2 * Cout, "pgk_conv_thin(out)");
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = PGK_ERR_ARG;
    // one-plane mode: the input-row-stationary flavour (see the kernel header); PGK_THIN_STK=0 keeps the
    // output-row-stationary issue order for A/B runs
    static int stk = -1;
    if (stk < 0) {
        const char* e = getenv("PGK_THIN_STK");
        stk = e ? atoi(e) != 0 : 1;
    }
#define PGK_THIN_STK_CASE(C_, N_) \
    if (stk && Pr == 1 && P == 1 && Cin == C_ && a.Npad == N_) rc = launch_thin<C_, 1, 0, N_, 1>(tmA, tmO, a, st);
    PGK_THIN_STK_CASE(8, 16) PGK_THIN_STK_CASE(8, 32) PGK_THIN_STK_CASE(8, 64)
    PGK_THIN_STK_CASE(16, 16) PGK_THIN_STK_CASE(16, 32) PGK_THIN_STK_CASE(16, 64)
    PGK_THIN_STK_CASE(32, 16) PGK_THIN_STK_CASE(32, 32) PGK_THIN_STK_CASE(32, 64)
    PGK_THIN_STK_CASE(64, 32) PGK_THIN_STK_CASE(64, 64)
#undef PGK_THIN_STK_CASE
#define PGK_THIN_CASE_N(C_, P_, S_, N_) \
    if (rc == PGK_ERR_ARG && Cin == C_ && Pr == P_ && a.split_acc == S_ && a.Npad == N_) rc = launch_thin<C_, P_, S_, N_, 0>(tmA, tmO, a, st);
#define PGK_THIN_CASE(C_, P_, S_) PGK_THIN_CASE_N(C_, P_, S_, 16) PGK_THIN_CASE_N(C_, P_, S_, 32) PGK_THIN_CASE_N(C_, P_, S_, 64)
    PGK_THIN_CASE(8, 1, 0) PGK_THIN_CASE(16, 1, 0) PGK_THIN_CASE(32, 1, 0)
    PGK_THIN_CASE(8, 2, 0) PGK_THIN_CASE(16, 2, 0) PGK_THIN_CASE(32, 2, 0)
    PGK_THIN_CASE(8, 3, 1) PGK_THIN_CASE(16, 3, 1) PGK_THIN_CASE(32, 3, 1)
#undef PGK_THIN_CASE
#undef PGK_THIN_CASE_N
    if (rc) {
        if (rc == PGK_ERR_ARG) pgk_set_error("pgk_conv_thin: no kernel instance for Cin %d Pr %d Npad %d", Cin, Pr, a.Npad);
        return rc;
    }
    PGK_LAUNCH_CHECK("pgk_conv(thin tcgen05)");
    return PGK_OK;
}

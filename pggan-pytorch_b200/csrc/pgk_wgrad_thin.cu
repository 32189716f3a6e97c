// pgk_wgrad_thin.cu -- weight gradient of the thin, high-resolution 3x3 layers (Cin in {8,16,32}, Cout in {8..64},
// W a multiple of 128) on the tensor cores.
//
//   dW[ky][kx][ci][co] = sum over samples, y, x of  X[y+ky-1][x+kx-1][ci] * G[y][x][co]
//
// The reduction runs over pixels, so both operands must be K-major along x -- the transpose of how activations are
// stored (channels innermost).  Per CTA (persistent, streaming rows like pgk_conv_thin.cu):
//   warp 4      producer: halo rows of X ([Cin/8][130+ px][8 ch], zero filled outside the image) and rows of G into raw
//               rings of 2-8 rows (TMA boxes; a 16-byte cp.async producer is kept as the PGK_THIN_TMA=0 flavour);
//   warps 0-3   transpose in shared memory with ldmatrix.trans + stmatrix (8x8 bf16 blocks): every X row once into
//               XT3 = three copies shifted by kx, stacked along M (row m = kx*Cin + ci), every G row once into GT;
//   warp 5      per output row y: D[ky] (3*Cin x Cout, fp32 in TMEM) += XT3[row y+ky-1] * GT[row y]^T, K = 128 pixels
//               (8 MMAs of K = 16 per ky); the accumulators persist over ALL rows this CTA processes;
//   warps 0-3   at the very end: TMEM -> fp32 atomics into dW[(ky*3+kx)*Cin + ci][co] (one flush per CTA).
// Cin = 8: the four ring slots of transposed rows are exactly the 16 row groups one M = 128 MMA reads, so ONE chain
// per output row covers ky = 0, 1, 2 (four rotating accumulators, see STACK below) -- a third of the MMAs.
// Every activation byte is read from HBM once; two CTAs per SM where shared memory allows (Cin <= 16).  The kernel is
// bound by the MMA rate (each MMA streams a 128-row A tile from shared memory) and the load -> transpose -> MMA chain.
#include <stdlib.h>
#include <string.h>

#include "pgk_tc.cuh"

using namespace tc;

namespace {

constexpr int kThreads = 192;
constexpr int kRowPix = 136;
constexpr int kCgBytes = kRowPix * 16;   // one channel group of a raw X row
constexpr int kGrp = 2048;               // one transposed 8-channel group: 16 pixel chunks x (8 ch x 8 px)
constexpr int kSmemLimit = 227 * 1024;

constexpr int kMaxRaw = 8;               // raw ring depth: 8, 4 or 2 rows of X and of G

struct WThinArgs {
    const bf16* x;            // X planes (N,H,W,cin_total), x_ps elements apart
    const bf16* g;            // G planes (N,H,W,Cout), g_ps elements apart
    long long x_ps, g_ps;
    int raw, raw_log2, look;  // raw ring depth and the rows (X or G) the producer keeps in flight
    int tma;                  // producer: 1 = TMA boxes, 0 = cp.async chunks
    int spin;                 // producer / MMA warps poll their barriers (1) or suspend in try_wait (0)
    int ks_major;             // MMA issue order: 1 = the three ky accumulators interleaved per K step
    int dbg;                  // PGK_WTHIN_DBG knock-outs for stage timing (results are wrong): 1 no MMAs, 2 no X
                              // transposition, 4 no G transposition, 8 no loads
    int H, W, Cout, Npad, CGO;
    int RC, chunks_y, strips;
    int ngroups, group_n;
    int xoff[4], goff[4];
    int total_units;
    uint32_t off_gt, off_rawx, off_rawg, off_bars;   // byte offsets from the 1024-aligned base
    int cin_total, c0;        // this launch handles input channels [c0, c0 + CIN) of cin_total
    float* dwp;
    float* db;                // optional fused bias gradient: db[co] += sum over pixels of G (groups in bias_mask)
    unsigned bias_mask;
};

// ATM = 1 (one-plane mode, 8 / 32 input channels; PGK_WTHIN_ATM=0 switches it off): the A operand lives in TENSOR MEMORY.  Cin = 8 (the stacked chain):  The four ring
// slots of 32 accumulator rows are the four 32-lane quarters of tensor memory, and a quarter belongs to one warp: the
// transposer warp (input row & 3) gathers its row straight from the raw [pixel][8 channels] buffer -- lane = (kx, ci)
// reads X[p + kx][ci] for p = 0..127, conflict-free 16-bit loads -- packs pixel pairs and writes them with tcgen05.st
// (64 columns per plane); the MMAs take [a_tmem] and no longer stream a 128-row A tile through the shared-memory port
// per K = 16 step (~75 cycles each measured, 8 per image row, against ~180 cycles of HBM time per row).  Same rings,
// barriers, accumulators and flush; no ldmatrix / stmatrix and no transposed tiles for X.  A slot is rewritten while
// the MMAs of the current row still read its lanes as the don't-care ky -- those accumulator rows are never flushed.
template <int CIN, int P, int ATM>
__global__ void __launch_bounds__(kThreads, 2)
wgrad_thin_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WThinArgs a) {
    // Cin = 16 / 32 with ATM (one plane only): every input row is one 128-lane A tile of its own (four of them, 64
    // columns each, after the three accumulators), with the rows ordered (channel group, kx, channel): lane quarter cg
    // = warp cg gathers from channel-group plane cg alone, conflict-free like the 8-channel case, and the flush maps
    // lane -> (cg, kx, ci) back.  Lanes 24 (ones row, quarter 0) .. 31 of a quarter and the quarters beyond Cin / 8 hold
    // zeros (written once at start-up).
    static_assert(!ATM || (CIN == 8 || P == 1), "tensor-memory A: one plane for Cin >= 16");
    // row groups of XT3: 3 * CG shifted copies + one group whose first row is all ones (the bias-gradient row)
    constexpr int CG = CIN / 8, XG = 3 * CG;
    constexpr uint32_t xt_plane = (XG + 1) * kGrp, xt_buf = P * xt_plane;
    // CIN == 8: one transposed row is 4 row groups (3 kx copies + the ones group) = 32 accumulator rows, so the FOUR
    // ring slots of one plane, laid out back to back, are exactly the 16 row groups (M = 128) one tcgen05.mma reads:
    // a single MMA chain per output row covers ky = 0, 1, 2 (plus one slot of don't-care rows) instead of three
    // chains -- a third of the MMAs, and this kernel is bound by the MMA rate (every MMA streams its 128-row A tile
    // from shared memory, whatever N is).  Slot j then holds ky = (j - row index) & 3, which changes from row to row,
    // so there are four accumulators, selected by (row index & 3): inside accumulator c slot j is always ky = (j-c)&3.
    constexpr bool STACK = CIN == 8;
    constexpr uint32_t rawx_plane = CG * kCgBytes, rawx_slot = P * rawx_plane;
    extern __shared__ uint8_t smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    const uint32_t gt_plane = (uint32_t)a.CGO * kGrp, gt_slot = P * gt_plane;
    const uint32_t rawg_plane = gt_plane, rawg_slot = gt_slot;
    const uint32_t xt0 = sbase, gt0 = sbase + a.off_gt, rawx0 = sbase + a.off_rawx, rawg0 = sbase + a.off_rawg;
    // transposed X row buffers: [slot][plane] (three chains) or [plane][slot] (stacked)
    auto xt_at = [&](uint32_t buf, uint32_t p) {
        return STACK ? xt0 + p * (4u * xt_plane) + buf * xt_plane : xt0 + buf * xt_buf + p * xt_plane;
    };
    const uint32_t bars = sbase + a.off_bars;
    const int kRaw = a.raw, kRawLog = a.raw_log2;
    auto xfull = [&](int s) { return bars + 8u * s; };                       // [kMaxRaw]
    auto xempty = [&](int s) { return bars + 8u * (kMaxRaw + s); };          // [kMaxRaw]
    auto gfull = [&](int s) { return bars + 8u * (2 * kMaxRaw + s); };       // [kMaxRaw]
    auto gempty = [&](int s) { return bars + 8u * (3 * kMaxRaw + s); };      // [kMaxRaw]
    const uint32_t bars2 = bars + 32u * kMaxRaw;
    auto xtfull = [&](int s) { return bars2 + 8u * s; };           // [4]
    auto xtempty = [&](int s) { return bars2 + 32u + 8u * s; };    // [4]
    auto gtfull = [&](int s) { return bars2 + 64u + 8u * s; };     // [2]
    auto gtempty = [&](int s) { return bars2 + 80u + 8u * s; };    // [2]
    const uint32_t done = bars2 + 96u, tptr = bars2 + 104u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kRaw; ++s) {
            mbar_init(xfull(s), 1), mbar_init(xempty(s), 4);
            mbar_init(gfull(s), 1), mbar_init(gempty(s), 4);
        }
        for (int s = 0; s < 2; ++s) mbar_init(gtfull(s), 4), mbar_init(gtempty(s), 1);
        for (int s = 0; s < 4; ++s) mbar_init(xtfull(s), 4), mbar_init(xtempty(s), 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    // the group holding the ones row: everything but row 0 of plane 0 stays zero for the whole kernel
    for (uint32_t o = threadIdx.x * 16u; o < 4u * P * kGrp; o += kThreads * 16u) {
        const uint32_t buf = o / (P * kGrp), rem = o - buf * (P * kGrp);
        const uint32_t p = rem / kGrp, w = rem - p * kGrp;
        st_shared_v4(xt_at(buf, p) + XG * kGrp + w, make_uint4(0, 0, 0, 0));
    }
    fence_proxy_async();
    const unsigned nacc = STACK ? 4u : 3u;
    // ATM: the A operand (128 pixels = 64 columns per plane) sits after the accumulators
    constexpr unsigned kAtmCols = STACK ? 64u * P : 256u;   // stacked: one tile per plane; else four tiles (ring slots)
    const unsigned ncols = ATM ? (nacc * a.Npad + kAtmCols <= 128 ? 128u : nacc * a.Npad + kAtmCols <= 256 ? 256u : 512u)
                               : (nacc * a.Npad <= 64 ? 64u : nacc * a.Npad <= 128 ? 128u : 256u);
    if (warp == 5) tmem_alloc(tptr, ncols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
    pgk_pdl_enter();   // everything above touched shared / tensor memory and the kernel parameters only
    if constexpr (ATM != 0 && !STACK) {
        // the four A tiles start out as zeros: lanes that never receive data must still be finite for the MMAs
        if (warp < 4) {
            const uint32_t t0 = tmem + 3u * a.Npad + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int c = 0; c < 256; c += 8) tmem_zero8(t0 + c);
            tmem_st_wait();
        }
        fence_before();
        __syncthreads();
        fence_after();
    }

    // busy polling by the two single-warp roles takes issue slots from the transposer warps of the same SM
    // sub-partitions; try_wait suspends instead
    auto wait_bar = [&](uint32_t bar, uint32_t parity) {
        if (a.spin) mbar_wait_spin(bar, parity);
        else mbar_wait(bar, parity);
    };
    auto unit_coords = [&](int u, int& xn, int& gn, int& x0, int& ya) {
        const int cy = u % a.chunks_y;
        int r = u / a.chunks_y;
        const int st = r % a.strips;
        r /= a.strips;
        const int smp = r % a.group_n, grp = r / a.group_n;
        xn = a.xoff[grp] + smp, gn = a.goff[grp] + smp;
        x0 = st * 128, ya = cy * a.RC;
        return grp;
    };

    if (warp == 4 && a.tma) {
        // ---- producer, TMA flavour: X row j of the unit, then (from j = 2) G row j - 2
        if (lane == 0) {
            tma_prefetch_desc(&tmX);
            tma_prefetch_desc(&tmG);
        }
        uint32_t gx = 0, gg = 0;
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            int xn, gn, x0, ya;
            unit_coords(u, xn, gn, x0, ya);
            for (int j = 0; j < a.RC + 2; ++j) {
                {
                    const int s = gx & (kRaw - 1);
                    wait_bar(xempty(s), ((gx >> kRawLog) & 1) ^ 1);
                    if (elect_one()) {
                        const uint32_t fb = xfull(s);
                        if (a.dbg & 8) mbar_arrive(fb);
                        else {
                        mbar_expect_tx(fb, P * CG * 130 * 16);
                        const uint32_t dst = rawx0 + s * rawx_slot;
#pragma unroll
                        for (int p = 0; p < P; ++p) {
#pragma unroll
                            for (int cg = 0; cg < CG; ++cg)
                                tma_load_5d(dst + p * rawx_plane + cg * kCgBytes, &tmX, fb, a.c0 + cg * 8, x0 - 1, ya - 1 + j, xn, p);
                        }
                        }
                    }
                    __syncwarp();
                    ++gx;
                }
                if (j >= 2) {
                    const int s = gg & (kRaw - 1);
                    wait_bar(gempty(s), ((gg >> kRawLog) & 1) ^ 1);
                    if (elect_one()) {
                        const uint32_t fb = gfull(s);
                        if (a.dbg & 8) mbar_arrive(fb);
                        else {
                        mbar_expect_tx(fb, P * a.CGO * 128 * 16);
                        const uint32_t dst = rawg0 + s * rawg_slot;
#pragma unroll
                        for (int p = 0; p < P; ++p)
                            for (int cg = 0; cg < a.CGO; ++cg)
                                tma_load_5d(dst + p * rawg_plane + cg * kGrp, &tmG, fb, cg * 8, x0, ya + j - 2, gn, p);
                        }
                    }
                    __syncwarp();
                    ++gg;
                }
            }
        }
    } else if (warp == 4) {
        // ---- producer: X row j of the unit, then (from j = 2) G row j - 2, each one cp.async group of 16-byte chunks
        // (chunk q of a row = byte 16 * q of the global row segment; home [channel group][pixel]).  `look` rows stay
        // in flight; bit (k & 31) of `kinds` remembers whether the k-th requested row was a G row, so that rows are
        // handed to the transposers in request order on the right barrier.
        uint32_t req = 0, signalled = 0, kinds = 0, gx = 0, gg = 0, sx = 0, sg = 0;
        auto hand_over = [&](uint32_t upto) {
            fence_proxy_async();
            __syncwarp();
            for (uint32_t r = signalled; r < upto; ++r) {
                if ((kinds >> (r & 31)) & 1) {
                    if (lane == 0) mbar_arrive(gfull(sg & (kRaw - 1)));
                    ++sg;
                } else {
                    if (lane == 0) mbar_arrive(xfull(sx & (kRaw - 1)));
                    ++sx;
                }
            }
            signalled = upto;
        };
        const int look = a.look;
        auto acquire = [&](uint32_t bar, uint32_t par) {
            if (!__all_sync(0xffffffffu, mbar_test(bar, par))) {
                cp_async_wait<0>();   // never sleep on a busy slot while holding rows back (the consumers may need them)
                hand_over(req);
                wait_bar(bar, par);
            }
        };
        auto after_request = [&](bool is_g) {
            cp_async_commit();
            kinds = (kinds & ~(1u << (req & 31))) | ((is_g ? 1u : 0u) << (req & 31));
            ++req;
            if ((int)(req - signalled) > look) {
                if (look >= 8) cp_async_wait<8>();
                else if (look >= 4) cp_async_wait<4>();
                else if (look >= 1) cp_async_wait<1>();
                else cp_async_wait<0>();
                const int lk = look >= 8 ? 8 : look >= 4 ? 4 : look >= 1 ? 1 : 0;
                hand_over(req - lk);
            }
        };
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            int xn, gn, x0, ya;
            unit_coords(u, xn, gn, x0, ya);
            for (int j = 0; j < a.RC + 2; ++j) {
                {
                    const int s = gx & (kRaw - 1);
                    acquire(xempty(s), ((gx >> kRawLog) & 1) ^ 1);
                    const int y = ya - 1 + j;
                    const bool row_ok = y >= 0 && y < a.H;
                    const bf16* rowp = a.x + (((long long)xn * a.H + (row_ok ? y : 0)) * a.W) * a.cin_total + a.c0;
                    const uint32_t dst = rawx0 + s * rawx_slot;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
#pragma unroll
                        for (int it = 0; it < (130 * CG + 31) / 32; ++it) {
                            const int q = it * 32 + lane;
                            if (q < 130 * CG) {
                                const int px = q / CG, cg = q % CG;
                                const int xx = x0 - 1 + px;
                                const bool ok = row_ok && xx >= 0 && xx < a.W;
                                const bf16* src = rowp + (long long)p * a.x_ps + (long long)(ok ? xx : 0) * a.cin_total + cg * 8;
                                cp_async16(dst + p * rawx_plane + cg * kCgBytes + px * 16, src, ok ? 16u : 0u);
                            }
                        }
                    }
                    ++gx;
                    after_request(false);
                }
                if (j >= 2) {
                    const int s = gg & (kRaw - 1);
                    acquire(gempty(s), ((gg >> kRawLog) & 1) ^ 1);
                    const bf16* rowp = a.g + (((long long)gn * a.H + (ya + j - 2)) * a.W + x0) * a.Cout;
                    const uint32_t dst = rawg0 + s * rawg_slot;
                    const int nchunk = 128 * a.CGO;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        for (int q = lane; q < nchunk; q += 32) {
                            const int px = q / a.CGO, cg = q - px * a.CGO;
                            cp_async16(dst + p * rawg_plane + cg * kGrp + px * 16, rowp + (long long)p * a.g_ps + (long long)q * 8, 16u);
                        }
                    }
                    ++gg;
                    after_request(true);
                }
            }
        }
        cp_async_wait<0>();
        hand_over(req);
    } else if (warp == 5) {
        // ---- MMA issue
        const uint32_t idesc = idesc_bf16(a.Npad, 0, 0);
        // K-major, no swizzle: LBO = next 8 pixels, SBO = next 8 rows
        const uint64_t dhi = smem_desc(0, 128, kGrp, 0);
        const uint32_t gtp16 = gt_plane >> 4;
        uint32_t gx = 0, gg = 0, rows_done = 0, used = 0;
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            wait_bar(xtfull(gx & 3), (gx >> 2) & 1);
            wait_bar(xtfull((gx + 1) & 3), ((gx + 1) >> 2) & 1);
            for (int i = 0; i < a.RC; ++i, ++gx, ++gg, ++rows_done) {
                wait_bar(xtfull((gx + 2) & 3), ((gx + 2) >> 2) & 1);
                wait_bar(gtfull(gg & 1), (gg >> 1) & 1);
                fence_after();
                const uint64_t bd0 = dhi | ((gt0 + (gg & 1) * gt_slot) >> 4);
                uint32_t xb[3];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) xb[ky] = (xt0 + ((gx + ky) & 3) * xt_buf) >> 4;
                const uint32_t c4 = gx & 3;
                const uint32_t later4 = (used >> c4) & 1u;   // stacked flavour: has accumulator c4 been started?
                used |= 1u << c4;
                if (elect_one()) {
                    const uint32_t later = rows_done > 0 ? 1u : 0u;
                    if (!(a.dbg & 1)) {
                    auto issue = [&](int ky, int ks) {
                        const uint32_t d = tmem + ky * a.Npad;
                        if constexpr (ATM != 0) {   // (Cin >= 16: one plane) A tile of input row gx + ky
                            mma_bf16_ts(d, tmem + 3u * a.Npad + ((gx + ky) & 3u) * 64u + (uint32_t)(ks * 8),
                                        bd0 + (uint32_t)(ks * 16), idesc, ks == 0 ? later : 1u);
                            return;
                        }
#pragma unroll
                        for (int pi = 0; pi < P; ++pi) {
#pragma unroll
                            for (int pj = 0; pj < P - pi; ++pj)
                                mma_bf16(d, dhi | (uint64_t)(xb[ky] + (pi * xt_plane + ks * 256) / 16),
                                         bd0 + (uint32_t)(pj * gtp16 + ks * 16), idesc,
                                         (ks == 0 && pi + pj == 0) ? later : 1u);
                        }
                    };
                    if (STACK && ATM) {
                        const uint32_t d = tmem + c4 * a.Npad, ta = tmem + 4u * a.Npad;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                            for (int pi = 0; pi < P; ++pi) {
#pragma unroll
                                for (int pj = 0; pj < P - pi; ++pj)
                                    mma_bf16_ts(d, ta + (uint32_t)(pi * 64 + ks * 8),
                                                bd0 + (uint32_t)(pj * gtp16 + ks * 16), idesc,
                                                (ks == 0 && pi + pj == 0) ? later4 : 1u);
                            }
                        }
                    } else if (STACK) {
                        const uint32_t d = tmem + c4 * a.Npad;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                            for (int pi = 0; pi < P; ++pi) {
#pragma unroll
                                for (int pj = 0; pj < P - pi; ++pj)
                                    mma_bf16(d, dhi | (uint64_t)((xt0 + pi * (4u * xt_plane) + ks * 256) >> 4),
                                             bd0 + (uint32_t)(pj * gtp16 + ks * 16), idesc,
                                             (ks == 0 && pi + pj == 0) ? later4 : 1u);
                            }
                        }
                    } else if (a.ks_major) {
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky) issue(ky, ks);
                        }
                    } else {
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks) issue(ky, ks);
                        }
                    }
                    }
                    mma_commit(gtempty(gg & 1));
                    mma_commit(xtempty(gx & 3));
                    if (i == a.RC - 1) {
                        mma_commit(xtempty((gx + 1) & 3));
                        mma_commit(xtempty((gx + 2) & 3));
                    }
                }
                __syncwarp();
            }
            gx += 2;
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
    } else {
        // ---- transposers (warps 0-3): lane l addresses row l & 7 of matrix l >> 3
        const int li = lane & 7, lj = lane >> 3;
        uint32_t gx = 0, gg = 0;
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            int xn_, gn_, x0_, ya_;
            const int grp = unit_coords(u, xn_, gn_, x0_, ya_);
            const uint32_t one16 = (a.db && ((a.bias_mask >> grp) & 1)) ? 0x3F803F80u : 0u;   // bf16 1.0 pairs
            for (int j = 0; j < a.RC + 2; ++j) {
                {
                    const int s = gx & (kRaw - 1), b = gx & 3;
                    mbar_wait(xfull(s), (gx >> kRawLog) & 1);
                    mbar_wait(xtempty(b), ((gx >> 2) & 1) ^ 1);
                    const uint32_t src0 = rawx0 + s * rawx_slot;
                    if (a.dbg & 2) {
                    } else if (ATM) {
                        // the warp that owns ring slot b (= lane quarter b of tensor memory) writes the whole row:
                        // lane = kx * 8 + ci gathers its channel of the kx-shifted pixels, lane 24 is the ones row
                        // stacked (Cin = 8): the warp that owns ring slot b writes the row into its lane quarter;
                        // Cin >= 16: warp cg writes channel group cg of the row into tile b (columns b * 64 ..)
                        if (STACK ? warp == b : warp < CG) {
                            const uint32_t ta = STACK ? tmem + 4u * a.Npad + ((uint32_t)(warp * 32) << 16)
                                                      : tmem + 3u * a.Npad + (uint32_t)b * 64u + ((uint32_t)(warp * 32) << 16);
                            const uint32_t lsrc = src0 + (STACK ? 0u : (uint32_t)warp * kCgBytes) + (uint32_t)(lane >> 3) * 16u +
                                                  (uint32_t)(lane & 7) * 2u;
                            for (int p = 0; p < P; ++p) {
#pragma unroll
                                for (int c = 0; c < 64; c += 8) {   // 8 columns = 16 pixels
                                    uint32_t v[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        uint32_t lo = 0, hi = 0;
                                        if (lane < 24) {
                                            const uint32_t ad = lsrc + p * rawx_plane + (uint32_t)(2 * (c + j)) * 16u;
                                            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(lo) : "r"(ad));
                                            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(hi) : "r"(ad + 16u));
                                        }
                                        v[j] = lane < 24 ? (lo | (hi << 16))
                                                         : (lane == 24 && p == 0 && (STACK || warp == 0)) ? one16 : 0u;
                                    }
                                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::
                                                 "r"(ta + (uint32_t)(p * 64 + c)), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                                                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
                                }
                            }
                            tmem_st_wait();
                            fence_before();
                        }
                    } else
                    // ops: (plane, kx, channel group, 32-pixel block)
                    for (int o = warp; o < P * 3 * CG * 4; o += 4) {
                        const int blk = o & 3;
                        int r = o >> 2;
                        const int cg = r % CG;
                        r /= CG;
                        const int kx = r % 3, p = r / 3;
                        uint32_t v[4];
                        ldmatrix_x4_trans(src0 + p * rawx_plane + cg * kCgBytes + (blk * 32 + lj * 8 + li + kx) * 16, v);
                        stmatrix_x4(xt_at(b, p) + (kx * CG + cg) * kGrp + (blk * 4 + lj) * 128 + li * 16, v);
                    }
                    if (!ATM && warp == 0 && lane < 16) {   // the ones row (plane 0, row 0 of group XG), 16 pixel chunks
                        const uint32_t dst = xt_at(b, 0) + XG * kGrp + lane * 128;
                        st_shared_v4(dst, make_uint4(one16, one16, one16, one16));
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(xempty(s));
                        mbar_arrive(xtfull(b));
                    }
                    ++gx;
                }
                if (j >= 2) {
                    const int s = gg & (kRaw - 1), t = gg & 1;
                    mbar_wait(gfull(s), (gg >> kRawLog) & 1);
                    mbar_wait(gtempty(t), ((gg >> 1) & 1) ^ 1);
                    const uint32_t src0 = rawg0 + s * rawg_slot, dst0 = gt0 + t * gt_slot;
                    for (int o = warp; o < P * a.CGO * 4 && !(a.dbg & 4); o += 4) {
                        const int blk = o & 3;
                        const int r = o >> 2;
                        const int cg = r % a.CGO, p = r / a.CGO;
                        uint32_t v[4];
                        ldmatrix_x4_trans(src0 + p * rawg_plane + cg * kGrp + (blk * 32 + lj * 8 + li) * 16, v);
                        stmatrix_x4(dst0 + p * gt_plane + cg * kGrp + (blk * 4 + lj) * 128 + li * 16, v);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(gempty(s));
                        mbar_arrive(gtfull(t));
                    }
                    ++gg;
                }
            }
        }
        // ---- flush: accumulator row m = kx*Cin + ci of D[ky] -> dW[(ky*3 + kx)*Cin + ci][:]
        mbar_wait(done, 0);
        fence_after();
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        if (STACK) {
            // accumulator c, rows 32*j .. 32*j+31 (this warp: j = warp) = slot j = ky (j - c) & 3; row kx*8 + ci of
            // the block, row 24 = the ones row.  Every CTA processes at least one unit of >= 8 rows, so all four
            // accumulators have been written.
            const int kx = lane >> 3, ci = lane & 7;
            for (int c4 = 0; c4 < 4; ++c4) {
                const int ky = (warp - c4) & 3;
                float* drow = a.dwp + (long long)((ky * 3 + kx) * a.cin_total + a.c0 + ci) * a.Cout;
                for (int c = 0; c < a.Npad; c += 16) {
                    float v[16];
                    tmem_ld16(trow + c4 * a.Npad + c, v);
                    if (ky < 3 && lane < 24) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c + j < a.Cout) atomicAdd(drow + c + j, v[j]);
                    } else if (ky == 1 && lane == 24 && a.db) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c + j < a.Cout) atomicAdd(a.db + c + j, v[j]);
                    }
                }
            }
        } else {
            const int m = warp * 32 + lane;
            // ATM: rows are ordered (channel group = warp, kx, channel), 24 per lane quarter, row 24 of quarter 0 = ones
            const bool valid = ATM ? (warp < CG && lane < 24) : m < 3 * CIN;
            const int kx = ATM ? lane >> 3 : m / CIN, ci = ATM ? warp * 8 + (lane & 7) : m - kx * CIN;
            for (int ky = 0; ky < 3; ++ky) {
                float* drow = a.dwp + (long long)((ky * 3 + kx) * a.cin_total + a.c0 + ci) * a.Cout;
                const bool bias_row = a.db && ky == 1 && (ATM ? m == 24 : m == 3 * CIN);   // the ones row against G of the same row
                for (int c = 0; c < a.Npad; c += 16) {
                    float v[16];
                    tmem_ld16(trow + ky * a.Npad + c, v);
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c + j < a.Cout) atomicAdd(drow + c + j, v[j]);
                    } else if (bias_row) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c + j < a.Cout) atomicAdd(a.db + c + j, v[j]);
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, ncols);
}

}  // namespace

static size_t wthin_layout(int Cin, int Cout, int Pr, int R, WThinArgs* a) {
    const int CG = Cin / 8, CGO = Cout / 8, npad = Cout < 16 ? 16 : Cout;
    const size_t xt_buf = (size_t)Pr * (3 * CG + 1) * kGrp, gt_slot = (size_t)Pr * CGO * kGrp;
    const size_t rawx_slot = (size_t)Pr * CG * kCgBytes, rawg_slot = gt_slot;
    size_t off = 4 * xt_buf;
    const size_t off_gt = off;
    off += 2 * gt_slot;
    const size_t off_rawx = off;
    off += (size_t)R * rawx_slot;
    off = (off + 127) & ~(size_t)127;
    const size_t off_rawg = off;
    off += (size_t)R * rawg_slot;
    // the MMAs read 16 row groups of A (3*Cin/8 are real) and Npad/8 of B: keep those reads inside the allocation
    size_t need = 3 * xt_buf + (size_t)(Pr - 1) * (3 * CG + 1) * kGrp + 16 * kGrp;
    const size_t need_b = off_gt + gt_slot + (size_t)(Pr - 1) * CGO * kGrp + (size_t)(npad / 8) * kGrp;
    if (need_b > need) need = need_b;
    if (off < need) off = need;
    off = (off + 15) & ~(size_t)15;
    if (a) {
        a->off_gt = (uint32_t)off_gt, a->off_rawx = (uint32_t)off_rawx, a->off_rawg = (uint32_t)off_rawg;
        a->off_bars = (uint32_t)off;
    }
    return off + 512 + 1024;
}

extern "C" int pgk_wgrad_thin_supported(int H, int W, int Cin, int Cout, int KS, int ups, int ngroups, int group_n,
                                        int Pr) {
    if (ups || KS != 3) return 0;
    if (Cin != 8 && Cin != 16 && Cin != 32) return 0;
    if (Cout != 8 && Cout != 16 && Cout != 32 && Cout != 64) return 0;
    if (W % 128 || H < 8 || H % 8) return 0;
    if (Pr < 1 || Pr > 3) return 0;
    return wthin_layout(Cin, Cout, Pr, 2, nullptr) <= (size_t)kSmemLimit;
}

struct WThinPlan {
    int occ, raw, smem;
};

template <int CIN, int P, int ATM>
static int launch_wthin(const CUtensorMap& tmX, const CUtensorMap& tmG, WThinArgs& a, cudaStream_t stream) {
    auto kern = wgrad_thin_kernel<CIN, P, ATM>;
    static bool attr = false;
    static WThinPlan plans[4];   // by Cout / 8 -> index 0..3 (8, 16, 32, 64)
    static bool have[4] = {};
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_wgrad_thin: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
        attr = true;
    }
    const int pi = a.Cout == 8 ? 0 : a.Cout == 16 ? 1 : a.Cout == 32 ? 2 : 3;
    if (!have[pi]) {
        // CTAs per SM first (co-resident CTAs overlap one CTA's transposes with the other's loads; PGK_THIN_OCC caps
        // it, default 2), then raw ring depth.  Limits: shared memory (asked of the runtime), 512 TMEM columns per SM.
        const char* e = getenv("PGK_THIN_OCC");
        int cap = e ? atoi(e) : 2;
        cap = cap < 1 ? 1 : cap > 2 ? 2 : cap;
        const int nacc = CIN == 8 ? 4 : 3;
        const int need = nacc * a.Npad + (ATM ? (CIN == 8 ? 64 * P : 256) : 0);
        const int ncols = need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
        if (cap > 512 / ncols) cap = 512 / ncols;
        WThinPlan pl = {0, 0, 0};
        for (int occ = cap; occ >= 1 && pl.occ == 0; --occ) {
            for (int R = 8; R >= 2 && pl.occ == 0; R >>= 1) {
                const size_t smem = wthin_layout(CIN, a.Cout, P, R, nullptr);
                // residency computed here: shared memory (+1 KB reserved per CTA) against the 228 KB of an SM
                // (registers are bounded by __launch_bounds__(kThreads, 2), TMEM columns by `cap` above)
                if (smem > (size_t)kSmemLimit || (size_t)occ * (smem + 1024) > (size_t)228 * 1024) continue;
                pl.occ = occ, pl.raw = R, pl.smem = (int)smem;
            }
        }
        if (getenv("PGK_THIN_DEBUG")) {
            int got = -1;
            cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&got, kern, kThreads, pl.smem);
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, kern);
            fprintf(stderr, "pgk_wgrad_thin<%d,%d,%d> Cout %d: plan occ %d raw %d smem %d | runtime says %d blocks/SM (%s), regs %d\n",
                    CIN, P, ATM, a.Cout, pl.occ, pl.raw, pl.smem, got, cudaGetErrorString(oe), fa.numRegs);
            cudaGetLastError();
        }
        if (pl.occ == 0) {
            pgk_set_error("pgk_wgrad_thin: no shared-memory plan for Cin %d Cout %d P %d", CIN, a.Cout, P);
            return PGK_ERR_ARG;
        }
        plans[pi] = pl, have[pi] = true;
    }
    const WThinPlan pl = plans[pi];
    wthin_layout(CIN, a.Cout, P, pl.raw, &a);
    a.raw = pl.raw;
    a.raw_log2 = pl.raw == 8 ? 3 : pl.raw == 4 ? 2 : 1;
    a.look = pl.raw == 8 ? 8 : pl.raw == 4 ? 4 : 1;
    int grid = pl.occ * pgk_num_sms();
    if (grid > a.total_units) grid = a.total_units;
    pgk_launch(kern, grid, kThreads, pl.smem, stream, tmX, tmG, a);
    return PGK_OK;
}

// the transposer-free one-plane flavour (pgk_wgrad_direct.cu)
int pgk_wgrad_thin_direct(const void* x, const void* g, int H, int W, int Cin, int cin_total, int c0, int Cout,
                          int cout_total, int co0, int ngroups, int group_n, const int* xoff, const int* goff, float* dwp,
                          float* db, unsigned bias_mask, pgk_stream_t stream);

// Cin: channels handled by this launch (8, 16 or 32), starting at channel c0 of an x tensor with cin_total channels
// (a 64-channel input is two launches).  db (optional): fused bias gradient over the groups in bias_mask; db and dwp
// are accumulated into (the caller zeroes them).
extern "C" int pgk_wgrad_thin(const void* x, long long x_ps, const void* g, long long g_ps, int P, int Pr, int H, int W,
                              int Cin, int cin_total, int c0, int Cout, int ngroups, int group_n, const int* xoff,
                              const int* goff, float* dwp, float* db, unsigned bias_mask, pgk_stream_t stream) {
    PGK_REQUIRE(pgk_wgrad_thin_supported(H, W, Cin, Cout, 3, 0, ngroups, group_n, Pr), "pgk_wgrad_thin: unsupported shape");
    PGK_REQUIRE(c0 >= 0 && c0 % 8 == 0 && c0 + Cin <= cin_total, "pgk_wgrad_thin: bad channel window");
    PGK_REQUIRE(P >= Pr && P <= 3 && ngroups >= 1 && ngroups <= 4, "pgk_wgrad_thin: bad planes / groups");
    PGK_REQUIRE((((uintptr_t)x | (uintptr_t)g) & 15) == 0 && (Pr == 1 || ((x_ps | g_ps) * 2) % 16 == 0),
                "pgk_wgrad_thin: x and g must be 16-byte aligned");
    {
        // one-plane mode: rows are read as MN-major operands where TMA put them (no transposition stage);
        // PGK_WTHIN_DIRECT=0 keeps the transposing kernel below for A/B runs
        static int direct = -1;
        if (direct < 0) {
            const char* e = getenv("PGK_WTHIN_DIRECT");
            direct = e ? atoi(e) != 0 : 1;
        }
        if (direct && P == 1 && Pr == 1) {
            int rc = pgk_wgrad_thin_direct(x, g, H, W, Cin, cin_total, c0, Cout, Cout, 0, ngroups, group_n, xoff, goff, dwp,
                                           db, bias_mask, stream);
            if (rc) return rc;
            PGK_LAUNCH_CHECK("pgk_wgrad(thin tcgen05, direct)");
            return PGK_OK;
        }
    }
    WThinArgs a;
    a.H = H, a.W = W, a.Cout = Cout, a.Npad = Cout < 16 ? 16 : Cout, a.CGO = Cout / 8;
    a.RC = H < 32 ? H : H % 32 == 0 ? 32 : H % 16 == 0 ? 16 : 8;   // (H is a multiple of 8)
    a.chunks_y = H / a.RC;
    a.strips = W / 128;
    a.ngroups = ngroups, a.group_n = group_n;
    for (int i = 0; i < 4; ++i) {
        a.xoff[i] = i < ngroups ? xoff[i] : 0;
        a.goff[i] = i < ngroups ? goff[i] : 0;
    }
    a.total_units = ngroups * group_n * a.strips * a.chunks_y;
    a.dwp = dwp, a.db = db, a.bias_mask = bias_mask;
    a.cin_total = cin_total, a.c0 = c0;
    a.x = (const bf16*)x, a.g = (const bf16*)g;
    a.x_ps = x_ps, a.g_ps = g_ps;
    PGK_REQUIRE((((uintptr_t)x | (uintptr_t)g) & 15) == 0 && (Pr == 1 || ((x_ps | g_ps) * 2) % 16 == 0),
                "pgk_wgrad_thin: x and g must be 16-byte aligned");
    static int use_tma = -1, ks_major = -1;
    if (use_tma < 0) {
        const char* e = getenv("PGK_THIN_TMA");
        use_tma = e ? atoi(e) != 0 : 1;
        e = getenv("PGK_WTHIN_KS_MAJOR");
        ks_major = e ? atoi(e) != 0 : 0;
    }
    a.tma = use_tma, a.ks_major = ks_major;
    {
        const char* e = getenv("PGK_WTHIN_DBG");
        a.dbg = e ? atoi(e) : 0;
    }
    {
        static int spin = -1;
        if (spin < 0) {
            const char* e = getenv("PGK_THIN_SPIN");
            spin = e ? atoi(e) != 0 : 0;
        }
        a.spin = spin;
    }
    CUtensorMap tmX, tmG;
    memset(&tmX, 0, sizeof(tmX));
    memset(&tmG, 0, sizeof(tmG));
    if (a.tma) {
        int xmax = 0, gmax = 0;
        for (int i = 0; i < ngroups; ++i) {
            if (xoff[i] > xmax) xmax = xoff[i];
            if (goff[i] > gmax) gmax = goff[i];
        }
        {
            const unsigned long long Ct = (unsigned long long)cin_total;
            unsigned long long dims[5] = {Ct, (unsigned long long)W, (unsigned long long)H,
                                          (unsigned long long)(xmax + group_n), (unsigned long long)P};
            unsigned long long str[4] = {2ull * Ct, 2ull * Ct * W, 2ull * Ct * W * H,
                                         P > 1 ? 2ull * x_ps : 2ull * Ct * W * H * (xmax + group_n)};
            unsigned box[5] = {8u, 130u, 1u, 1u, 1u};
            int rc = pgk_make_tmap(&tmX, x, 5, dims, str, box, 0, "pgk_wgrad_thin(x)");
            if (rc) return rc;
        }
        {
            unsigned long long dims[5] = {(unsigned long long)Cout, (unsigned long long)W, (unsigned long long)H,
                                          (unsigned long long)(gmax + group_n), (unsigned long long)P};
            unsigned long long str[4] = {2ull * Cout, 2ull * Cout * W, 2ull * Cout * W * H,
                                         P > 1 ? 2ull * g_ps : 2ull * Cout * W * H * (gmax + group_n)};
            unsigned box[5] = {8u, 128u, 1u, 1u, 1u};
            int rc = pgk_make_tmap(&tmG, g, 5, dims, str, box, 0, "pgk_wgrad_thin(g)");
            if (rc) return rc;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = PGK_ERR_ARG;
    // one-plane mode, 8 / 32 input channels: the A operand in tensor memory (see the kernel header).  Measured per
    // shape at batch 12 (tools/thin_bench.py): 8 -> 8 0.345 -> 0.278 ms, 8 -> 16 0.346 -> 0.299, 32 -> 16 0.234 -> 0.169,
    // 32 -> 32 0.108 -> 0.084, 32 -> 64 0.128 -> 0.112; the 16-channel flavour was slower (0.139 -> 0.160) and stays on
    // shared-memory operands.  PGK_WTHIN_ATM=0 switches it off for A/B runs.
    static int watm = -1;
    if (watm < 0) {
        const char* e = getenv("PGK_WTHIN_ATM");
        watm = e ? atoi(e) != 0 : 1;
    }
    if (watm && Cin == 8 && Pr == 1) rc = launch_wthin<8, 1, 1>(tmX, tmG, a, st);
    if (watm && Cin == 32 && Pr == 1) rc = launch_wthin<32, 1, 1>(tmX, tmG, a, st);
#define PGK_WTHIN_CASE(C_, P_) \
    if (rc == PGK_ERR_ARG && Cin == C_ && Pr == P_) rc = launch_wthin<C_, P_, 0>(tmX, tmG, a, st);
    PGK_WTHIN_CASE(8, 1) PGK_WTHIN_CASE(16, 1) PGK_WTHIN_CASE(32, 1)
    PGK_WTHIN_CASE(8, 2) PGK_WTHIN_CASE(16, 2) PGK_WTHIN_CASE(32, 2)
    PGK_WTHIN_CASE(8, 3) PGK_WTHIN_CASE(16, 3) PGK_WTHIN_CASE(32, 3)
#undef PGK_WTHIN_CASE
    if (rc) {
        if (rc == PGK_ERR_ARG) pgk_set_error("pgk_wgrad_thin: no kernel instance for Cin %d Pr %d", Cin, Pr);
        return rc;
    }
    PGK_LAUNCH_CHECK("pgk_wgrad(thin tcgen05)");
    return PGK_OK;
}

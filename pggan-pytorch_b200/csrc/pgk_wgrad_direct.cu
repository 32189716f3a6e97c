// pgk_wgrad_direct.cu -- weight (+ bias) gradient of the thin, high-resolution 3x3 layers in the one-plane mode,
// WITHOUT transposing anything: image rows stay in shared memory exactly as TMA writes them -- natural [pixel][channel]
// rows in the 32 / 64 / 128-byte swizzle that matches their width -- and tcgen05.mma reads them as MN-major operands
// of a reduction over pixels.
//
//   dW[ky][kx][ci][co] = sum over samples, y, x of  X[y+ky-1][x+kx-1][ci] * G[y][x][co]
//
// Per input row r of X (one 128-pixel strip) ONE chain of 8 MMAs (K = 16 pixels each):
//   A = the G ring, rows r-1, r, r+1: three consecutive ring slots are three M groups of Cout channels, a slot apart
//       (M = 64 or 128; rows beyond the third slot are don't-care rows that are never flushed);
//   B = the X row read at the three pixel shifts kx = 0, 1, 2: N groups of Cin channels ONE PIXEL apart -- the 8-pixel
//       core matrices of neighbouring groups overlap in memory, which an MN-major descriptor is free to describe
//       (tools/probes/mnmajor_probe.cu: exact for every channel-count pair, swizzled or not; 69-130 cycles per MMA);
//   D[(slot, co)][(kx, ci)] accumulates in tensor memory over ALL rows the CTA ever processes: slot j of the window is
//       G row r-1+j, i.e. ky = 2 - j, whatever r is -- one flush of 9 * Cin * Cout atomics per CTA.
// So that a window never wraps, the ring has two mirror slots behind it: rows that land in slots 0 and 1 are loaded
// into slots R and R+1 as well (a second TMA box from L2).  Out-of-image G rows and the x halo of X are zero-filled by
// the tensor maps.  Replaces the load -> ldmatrix/stmatrix (or tcgen05.st gather) -> MMA chain of pgk_wgrad_thin.cu,
// whose two-deep hand-over between transposer and MMA warps bounded it (PGK_WTHIN_DBG knock-outs: the barrier skeleton
// alone cost 35-60 % of the kernel); here a row is TMA -> MMA -> slot free, with rings of 4-8 rows.
//   warp 4      producer (TMA);   warp 5  MMA issue;
//   warps 0-3   bias gradient (column sums of the G rows, read from the same ring) and the final flush.
#include <stdlib.h>
#include <string.h>

#include "pgk_tc.cuh"

using namespace tc;

namespace {

constexpr int kThreads = 192;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxRing = 8;

struct WDirArgs {
    int H, W, Cout;
    int RC, chunks_y, strips;
    int ngroups, group_n;
    int xoff[4], goff[4];
    int total_units;
    int rx, rx_log2, rg, rg_log2;   // ring depths (powers of two); the G ring has two mirror slots behind it
    uint32_t off_g, off_bars;       // byte offsets from the 1024-aligned base (the X ring starts at 0)
    int cin_total, c0;              // this launch handles input channels [c0, c0 + CIN) of cin_total
    int cout_total, co0;            // ... and output channels [co0, co0 + COUT) of cout_total (a.Cout = cout_total)
    float* dwp;
    float* db;                      // optional fused bias gradient over the groups in bias_mask
    unsigned bias_mask;
    int spin, dbg;                  // dbg (PGK_WTHIN_DBG): 1 no MMAs, 8 no loads (stage timing; results are wrong)
};

// natural [pixel][C channels] rows of C * 2 bytes: swizzle span = row size (none for 16-byte rows)
template <int C>
struct Lay {
    static constexpr uint32_t cb = 2u * C;
    static constexpr uint32_t code = C == 8 ? 0u : C == 16 ? 6u : C == 32 ? 4u : 2u;   // smem_desc layout field
    static constexpr uint32_t swmask = cb / 16u - 1u;
};

// MN-major operand over such rows: K = 16 pixels = two 8-pixel core matrices 8 rows apart; MN groups `mn` bytes apart.
// Un-swizzled descriptors keep the MN-group stride in SBO and the K-group stride in LBO, swizzled ones the other way
// round (cute's canonical MN-major layouts; both readings verified by the probe).
template <int C>
__device__ __forceinline__ uint64_t mn_desc(uint32_t start, uint32_t mn) {
    if (C == 8) return smem_desc(start, 128u, mn, 0u);
    return smem_desc(start, mn, 8u * Lay<C>::cb, Lay<C>::code);
}

__host__ __device__ inline uint32_t idesc_mn(int M, int N) {
    return (idesc_bf16(N, 1, 1) & ~(0x1Fu << 24)) | ((uint32_t)(M >> 4) << 24);
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 2)
wgrad_direct_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WDirArgs a) {
    constexpr uint32_t cbx = Lay<CIN>::cb, cbg = Lay<COUT>::cb;
    constexpr uint32_t xslot = 136u * cbx, gslot = 128u * cbg;
    constexpr int N = 3 * CIN;
    constexpr int M0 = COUT <= 16 ? 64 : 128;      // first MMA: 8 / 4 / 4 / 2 slots of 8 / 16 / 32 / 64 channels
    constexpr bool TWO = COUT == 64;               // second MMA (M = 64): the third slot of the window
    constexpr uint32_t acc_cols = N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : 256u;
    constexpr uint32_t ncols = TWO ? 2u * acc_cols : acc_cols;
    extern __shared__ uint8_t smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    const uint32_t xr0 = sbase, gr0 = sbase + a.off_g, bars = sbase + a.off_bars;
    const int RX = a.rx, RXL = a.rx_log2, RG = a.rg, RGL = a.rg_log2;
    auto xfull = [&](int s) { return bars + 8u * s; };
    auto xempty = [&](int s) { return bars + 8u * (kMaxRing + s); };
    auto gfull = [&](int s) { return bars + 8u * (2 * kMaxRing + s); };
    auto gempty = [&](int s) { return bars + 8u * (3 * kMaxRing + s); };
    const uint32_t done = bars + 32u * kMaxRing, tptr = done + 8u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RX; ++s) mbar_init(xfull(s), 1), mbar_init(xempty(s), 1);
        for (int s = 0; s < RG; ++s) mbar_init(gfull(s), 1), mbar_init(gempty(s), 5);   // MMA commit + warps 0-3
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tptr, ncols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
    pgk_pdl_enter();   // everything above touched shared / tensor memory and the kernel parameters only

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

    if (warp == 4) {
        // ---- producer: G rows ya-1 .. ya+RC of the unit (sequence numbers j = 0 .. RC+1), X row ya+j-2 after G row j
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
                    const int s = gg & (RG - 1);
                    wait_bar(gempty(s), ((gg >> RGL) & 1) ^ 1);
                    if (elect_one()) {
                        const uint32_t fb = gfull(s);
                        if (a.dbg & 8) {
                            mbar_arrive(fb);
                        } else {
                            mbar_expect_tx(fb, (s < 2 ? 2u : 1u) * gslot);
                            tma_load_5d(gr0 + s * gslot, &tmG, fb, a.co0, x0, ya - 1 + j, gn, 0);
                            if (s < 2) tma_load_5d(gr0 + (RG + s) * gslot, &tmG, fb, a.co0, x0, ya - 1 + j, gn, 0);
                        }
                    }
                    __syncwarp();
                    ++gg;
                }
                if (j >= 2) {
                    const int s = gx & (RX - 1);
                    wait_bar(xempty(s), ((gx >> RXL) & 1) ^ 1);
                    if (elect_one()) {
                        const uint32_t fb = xfull(s);
                        if (a.dbg & 8) {
                            mbar_arrive(fb);
                        } else {
                            mbar_expect_tx(fb, 130u * cbx);
                            tma_load_5d(xr0 + s * xslot, &tmX, fb, a.c0, x0 - 1, ya + j - 2, xn, 0);
                        }
                    }
                    __syncwarp();
                    ++gx;
                }
            }
        }
    } else if (warp == 5) {
        // ---- MMA issue: input row j of the unit against the window of G sequence numbers j, j+1, j+2
        const uint32_t idesc0 = idesc_mn(M0, N), idesc1 = idesc_mn(64, N);
        uint32_t gx = 0, gg = 0, started = 0;
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            wait_bar(gfull(gg & (RG - 1)), (gg >> RGL) & 1);
            wait_bar(gfull((gg + 1) & (RG - 1)), ((gg + 1) >> RGL) & 1);
            for (int j = 0; j < a.RC; ++j, ++gx, ++gg) {
                wait_bar(gfull((gg + 2) & (RG - 1)), ((gg + 2) >> RGL) & 1);
                const int xs = gx & (RX - 1), ws = gg & (RG - 1);
                wait_bar(xfull(xs), (gx >> RXL) & 1);
                fence_after();
                if (elect_one()) {
                    const uint32_t ab = gr0 + ws * gslot, bb = xr0 + xs * xslot;   // (mirror slots keep ws .. ws+2 contiguous)
                    if (!(a.dbg & 1)) {
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            const uint32_t acc = (started | (uint32_t)ks) ? 1u : 0u;
                            const uint64_t bd = mn_desc<CIN>(bb + ks * 16u * cbx, cbx);
                            mma_bf16(tmem, mn_desc<COUT>(ab + ks * 16u * cbg, gslot), bd, idesc0, acc);
                            if (TWO) mma_bf16(tmem + acc_cols, mn_desc<COUT>(ab + 2u * gslot + ks * 16u * cbg, gslot), bd, idesc1, acc);
                        }
                    }
                    mma_commit(xempty(xs));
                    mma_commit(gempty(ws));                  // G row j is not part of any later window
                    if (j == a.RC - 1) {
                        mma_commit(gempty((gg + 1) & (RG - 1)));
                        mma_commit(gempty((gg + 2) & (RG - 1)));
                    }
                }
                started = 1;
                __syncwarp();
            }
            gg += 2;
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
    } else {
        // ---- warps 0-3: bias gradient (thread = pixel of the strip: sums of its Cout channels over the G rows of
        // this CTA's units), then the flush
        float bsum[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) bsum[c] = 0.f;
        const uint32_t px = (uint32_t)(warp * 32 + lane);
        uint32_t gg = 0;
        for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
            int xn_, gn_, x0_, ya_;
            const int grp = unit_coords(u, xn_, gn_, x0_, ya_);
            const bool on = a.db && ((a.bias_mask >> grp) & 1);
            for (int t = 0; t < a.RC + 2; ++t, ++gg) {
                const int s = gg & (RG - 1);
                mbar_wait(gfull(s), (gg >> RGL) & 1);
                if (on && t >= 1 && t <= a.RC) {
                    const uint32_t row = gr0 + s * gslot + px * cbg, sw = (row >> 7) & Lay<COUT>::swmask;
#pragma unroll
                    for (int h = 0; h < COUT / 8; ++h) {
                        float f[8];
                        unpack8(ld_shared_v4(row + (((uint32_t)h ^ sw) << 4)), f);
#pragma unroll
                        for (int e = 0; e < 8; ++e) bsum[8 * h + e] += f[e];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(gempty(s));
            }
        }
        if (a.db) {
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                float v = bsum[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && v != 0.f) atomicAdd(a.db + a.co0 + c, v);
            }
        }
        // ---- flush: accumulator row m = slot * Cout + co (ky = 2 - slot), column n = kx * Cin + ci.  An M = 64
        // accumulator keeps row m in lane (m & 15) + 32 * (m >> 4).
        mbar_wait(done, 0);
        fence_after();
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        auto flush = [&](uint32_t col0, int M, int slot0) {
            const bool lane_ok = M == 128 || lane < 16;
            const int m = M == 128 ? warp * 32 + lane : warp * 16 + lane;
            const int slot = slot0 + m / COUT, co = m % COUT;
            const bool ok = lane_ok && slot < 3;
            const int ky = 2 - slot;
#pragma unroll 1
            for (int c = 0; c < N; c += 16) {
                float v[16];
                tmem_ld16(trow + col0 + c, v);
                if (ok) {
#pragma unroll
                    for (int jn = 0; jn < 16; ++jn) {
                        const int n = c + jn;
                        if (n < N) {
                            const int kx = n / CIN, ci = n % CIN;
                            atomicAdd(a.dwp + (long long)((ky * 3 + kx) * a.cin_total + a.c0 + ci) * a.Cout + a.co0 + co, v[jn]);
                        }
                    }
                }
            }
        };
        flush(0u, M0, 0);
        if (TWO) flush(acc_cols, 64, 2);
    }
    fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, ncols);
}

struct WDirPlan {
    int occ, rx, rg, smem;
    uint32_t off_g, off_bars;
};

// shared memory of one CTA: X ring | G ring + 2 mirror slots | barriers; the MMAs read up to 8 slots from a window
// start (Cout = 8: M = 64 rows = 8 groups of 8 channels), which must stay inside the allocation
static size_t wdir_layout(int Cin, int Cout, int rx, int rg, WDirPlan* pl) {
    const size_t xslot = 136 * 2 * (size_t)Cin, gslot = 128 * 2 * (size_t)Cout;
    size_t off = ((size_t)rx * xslot + 1023) & ~(size_t)1023;
    const size_t off_g = off;
    off += (size_t)(rg + 2) * gslot;
    const size_t groups = Cout <= 16 ? 64 / (size_t)Cout : Cout == 32 ? 4 : 3;
    const size_t reach = off_g + (size_t)(rg - 1 + groups) * gslot;
    if (off < reach) off = reach;
    off = (off + 15) & ~(size_t)15;
    if (pl) pl->off_g = (uint32_t)off_g, pl->off_bars = (uint32_t)off;
    return off + 32 * kMaxRing + 64 + 1024;
}

template <int CIN, int COUT>
static int launch_wdir(const CUtensorMap& tmX, const CUtensorMap& tmG, WDirArgs& a, cudaStream_t stream) {
    auto kern = wgrad_direct_kernel<CIN, COUT>;
    static bool have = false;
    static WDirPlan plan;
    if (!have) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_wgrad_thin(direct): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
        // two CTAs per SM first (one CTA's MMAs run under the other's loads), then ring depth
        WDirPlan pl = {0, 0, 0, 0, 0, 0};
        for (int occ = 2; occ >= 1 && pl.occ == 0; --occ) {
            for (int r = 8; r >= 4 && pl.occ == 0; r >>= 1) {
                const size_t smem = wdir_layout(CIN, COUT, r, r, &pl);
                if (smem > (size_t)kSmemLimit || (size_t)occ * (smem + 1024) > (size_t)228 * 1024) continue;
                pl.occ = occ, pl.rx = r, pl.rg = r, pl.smem = (int)smem;
            }
        }
        if (getenv("PGK_THIN_DEBUG")) {
            int got = -1;
            cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&got, kern, kThreads, pl.smem);
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, kern);
            fprintf(stderr, "pgk_wgrad_thin(direct)<%d,%d>: plan occ %d rings %d smem %d | runtime says %d blocks/SM (%s), regs %d\n",
                    CIN, COUT, pl.occ, pl.rx, pl.smem, got, cudaGetErrorString(oe), fa.numRegs);
            cudaGetLastError();
        }
        if (pl.occ == 0) {
            pgk_set_error("pgk_wgrad_thin(direct): no shared-memory plan for Cin %d Cout %d", CIN, COUT);
            return PGK_ERR_ARG;
        }
        plan = pl, have = true;
    }
    a.rx = plan.rx, a.rg = plan.rg;
    a.rx_log2 = plan.rx == 8 ? 3 : 2, a.rg_log2 = plan.rg == 8 ? 3 : 2;
    a.off_g = plan.off_g, a.off_bars = plan.off_bars;
    int grid = plan.occ * pgk_num_sms();
    if (grid > a.total_units) grid = a.total_units;
    pgk_launch(kern, grid, kThreads, plan.smem, stream, tmX, tmG, a);
    return PGK_OK;
}

}  // namespace

// The one-plane (Pr = 1) flavour of pgk_wgrad_thin: same arguments and meaning (pgk_wgrad_thin.cu), called from there.
int pgk_wgrad_thin_direct(const void* x, const void* g, int H, int W, int Cin, int cin_total, int c0, int Cout,
                          int cout_total, int co0, int ngroups, int group_n, const int* xoff, const int* goff, float* dwp,
                          float* db, unsigned bias_mask, pgk_stream_t stream) {
    WDirArgs a;
    a.H = H, a.W = W, a.Cout = cout_total;
    a.cout_total = cout_total, a.co0 = co0;
    a.RC = H < 32 ? H : H % 32 == 0 ? 32 : H % 16 == 0 ? 16 : 8;   // (H is a multiple of 8)
    a.chunks_y = H / a.RC;
    a.strips = W / 128;
    a.ngroups = ngroups, a.group_n = group_n;
    int xmax = 0, gmax = 0;
    for (int i = 0; i < 4; ++i) {
        a.xoff[i] = i < ngroups ? xoff[i] : 0;
        a.goff[i] = i < ngroups ? goff[i] : 0;
        if (a.xoff[i] > xmax) xmax = a.xoff[i];
        if (a.goff[i] > gmax) gmax = a.goff[i];
    }
    a.total_units = ngroups * group_n * a.strips * a.chunks_y;
    a.dwp = dwp, a.db = db, a.bias_mask = bias_mask;
    a.cin_total = cin_total, a.c0 = c0;
    {
        static int spin = -1;
        if (spin < 0) {
            const char* e = getenv("PGK_THIN_SPIN");
            spin = e ? atoi(e) != 0 : 0;
        }
        a.spin = spin;
        const char* e = getenv("PGK_WTHIN_DBG");
        a.dbg = e ? atoi(e) : 0;
    }
    CUtensorMap tmX, tmG;
    {
        const unsigned long long Ct = (unsigned long long)cin_total;
        unsigned long long dims[5] = {Ct, (unsigned long long)W, (unsigned long long)H, (unsigned long long)(xmax + group_n), 1ull};
        unsigned long long str[4] = {2ull * Ct, 2ull * Ct * W, 2ull * Ct * W * H, 2ull * Ct * W * H * (xmax + group_n)};
        unsigned box[5] = {(unsigned)Cin, 130u, 1u, 1u, 1u};
        int rc = pgk_make_tmap(&tmX, x, 5, dims, str, box, Cin == 8 ? 0 : 2 * Cin, "pgk_wgrad_thin(direct, x)");
        if (rc) return rc;
    }
    {
        const unsigned long long Ct = (unsigned long long)cout_total;
        unsigned long long dims[5] = {Ct, (unsigned long long)W, (unsigned long long)H,
                                      (unsigned long long)(gmax + group_n), 1ull};
        unsigned long long str[4] = {2ull * Ct, 2ull * Ct * W, 2ull * Ct * W * H, 2ull * Ct * W * H * (gmax + group_n)};
        unsigned box[5] = {(unsigned)Cout, 128u, 1u, 1u, 1u};
        int rc = pgk_make_tmap(&tmG, g, 5, dims, str, box, Cout == 8 ? 0 : 2 * Cout, "pgk_wgrad_thin(direct, g)");
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = PGK_ERR_ARG;
    bool matched = false;
#define PGK_WDIR_CASE(C_, O_) \
    if (Cin == C_ && Cout == O_) matched = true, rc = launch_wdir<C_, O_>(tmX, tmG, a, st);
    PGK_WDIR_CASE(8, 8) PGK_WDIR_CASE(8, 16) PGK_WDIR_CASE(8, 32) PGK_WDIR_CASE(8, 64)
    PGK_WDIR_CASE(16, 8) PGK_WDIR_CASE(16, 16) PGK_WDIR_CASE(16, 32) PGK_WDIR_CASE(16, 64)
    PGK_WDIR_CASE(32, 8) PGK_WDIR_CASE(32, 16) PGK_WDIR_CASE(32, 32) PGK_WDIR_CASE(32, 64)
    PGK_WDIR_CASE(64, 32) PGK_WDIR_CASE(64, 64)
#undef PGK_WDIR_CASE
    if (!matched) pgk_set_error("pgk_wgrad_thin(direct): no kernel instance for Cin %d Cout %d", Cin, Cout);
    return rc;
}

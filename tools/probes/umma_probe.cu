// umma_probe.cu -- standalone hardware probe (development aid, not part of libpgk): three questions about tcgen05.mma
// operands that decide how the thin-layer kernels and a row-streaming wide conv should lay out their tiles
// (DESIGN.md 7c).
//
//  (1) ROW-SHIFTED STARTS IN A 128-BYTE-SWIZZLED K-MAJOR TILE.  A conv tap (dy, dx) over an image row kept in shared
//      memory as [pixel][64 channels] is the same bytes read from a start address dx rows (dx * 128 bytes) further on.
//      With SWIZZLE_128B the XOR pattern repeats every 8 rows (1024 bytes); the descriptor has a 3-bit "matrix base
//      offset" for starts that are not 1024-byte aligned.  The probe multiplies rows [r0, r0+128) of a 256-row tile
//      for r0 = 0..9 with base_offset = 0 and with base_offset = (start >> 7) & 7 and compares with the CPU: whichever
//      variant is exact for every r0 is how taps-as-descriptor-offsets work with swizzled operands.
//  (2) COST PER MMA vs N AND OPERAND LAYOUT at M = 128, K = 16: back-to-back accumulating MMAs from un-swizzled
//      K-major core matrices (what pgk_conv_thin.cu / pgk_wgrad_thin.cu use today: ~75 cycles at N = 16 measured
//      inside the kernels) against SWIZZLE_128B K-major tiles (what pgk_conv_tc.cu uses), N = 16 ... 256.
//  (3) THE SAME WITH A IN TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): B300_MICROARCH.md gives a floor of
//      128 * N / 256 cycles per MMA with A in TMEM (8 cycles at N = 16) and says the shared-memory A read is what an
//      SS-mode MMA exposes (128 rows x 32 bytes = 32 cycles of the 128 B/cycle port).  If the table confirms it, the
//      thin weight gradient (one MMA per 16 pixels, N = 8..64: bound by exactly this) should write its transposed X
//      rows to TMEM (tcgen05.st from the transposer warps) instead of shared memory.
//
//  (4) A FILLED BY tcgen05.cp.  The thin conv keeps image rows in shared memory as [pixel][8 channels] (16 bytes per
//      pixel), and a tap's dx shift is a 16-byte shift of the start address.  tcgen05.cp .128x128b copies 128 rows x
//      16 bytes -- exactly 128 pixels of one channel group, from any such start -- into four tensor-memory columns.
//      Two neighbouring four-column slots are one K = 16 A operand, so an input row copied once per dx (3 copies of
//      2 KB per channel group) serves the three output rows that use it, and the nine taps become MMAs whose A never
//      touches the shared-memory port (8 cycles per MMA at N = 16 instead of 32+).  The probe checks the copy +
//      MMA numerics for dx = 0..2 against the CPU (which also settles the ordering of tcgen05.cp -> tcgen05.mma issued
//      by one thread) and times copy / MMA mixes per iteration.
//
// Build + run on the GPU box (nvcc is in the image):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/umma_probe tools/probes/umma_probe.cu && /tmp/umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../pggan-pytorch_b200/csrc/pgk_tc.cuh"

using namespace tc;

namespace {

constexpr int kRows = 256;   // rows of the A and B tiles held in shared memory
constexpr int kK = 64;       // K elements per row (128 bytes)
constexpr uint32_t kTile = kRows * 128;

enum Layout { SW128 = 0, NOSWZ = 1, ATMEM = 2 };   // ATMEM: A read from tensor memory, B from a SW128 tile

// byte offset of element (r, k) of a kRows x 64 tile
__host__ __device__ inline uint32_t elem_off(int layout, int r, int k) {
    if (layout != NOSWZ) return (uint32_t)r * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
    // un-swizzled K-major core matrices (8 rows x 16 bytes, contiguous): K group kg at kg * (kRows * 16), rows 16 bytes apart
    return ((uint32_t)k >> 3) * (kRows * 16u) + (uint32_t)r * 16u + ((uint32_t)k & 7u) * 2u;
}

// (mma_bf16_ts and tmem_cp_128x128b come from pgk_tc.cuh)
// 32 lanes x 8 consecutive 32-bit columns from registers (lane = TMEM lane of this warp's quarter)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint64_t make_desc(int layout, uint32_t tile_base, int r0, int kstep, int base_off_mode) {
    if (layout != NOSWZ) {
        const uint32_t start = tile_base + (uint32_t)r0 * 128u + (uint32_t)kstep * 32u;
        uint64_t d = smem_desc(start, 16, 1024, 2);
        if (base_off_mode) d |= (uint64_t)(((tile_base + (uint32_t)r0 * 128u) >> 7) & 7u) << 49;
        return d;
    }
    // K = 16 = two K groups: LBO = distance between them, SBO = 128 bytes to the next 8 rows
    const uint32_t start = tile_base + (uint32_t)(2 * kstep) * (kRows * 16u) + (uint32_t)r0 * 16u;
    return smem_desc(start, kRows * 16u, 128, 0);
}

struct ProbeArgs {
    const bf16* A;   // [kRows][64] row major
    const bf16* B;   // [kRows][64] row major
    float* D;        // [128][256]
    long long* cycles;
    int layout, N, r0, base_off_mode, iters, two_acc;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const ProbeArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t tA = base, tB = base + kTile, bar = base + 2 * kTile, tptr = bar + 16;
    uint8_t* gen = smem_raw + (base - raw);
    for (int i = threadIdx.x; i < kRows * kK; i += blockDim.x) {
        const int r = i / kK, k = i % kK;
        *reinterpret_cast<bf16*>(gen + elem_off(a.layout, r, k)) = a.A[i];
        *reinterpret_cast<bf16*>(gen + kTile + elem_off(a.layout, r, k)) = a.B[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tptr, 512);
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
    const uint32_t idesc = idesc_bf16(a.N, 0, 0);
    constexpr uint32_t kAcol = 480;   // A in tensor memory: 128 lanes x 32 columns (K = 64 bf16, two per column)
    if (a.layout == ATMEM) {
        const int row = a.r0 + warp * 32 + lane;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.A + (size_t)row * kK);
        for (int c = 0; c < 32; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = src[c + j];
            tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + kAcol + c, v);
        }
        fence_before();
        __syncthreads();
        fence_after();
    }
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        uint64_t ad[4], bd[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            ad[ks] = make_desc(a.layout, tA, a.r0, ks, a.base_off_mode);
            bd[ks] = make_desc(a.layout, tB, 0, ks, 0);
        }
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
            const uint32_t d = tmem + ((a.two_acc && (it & 1)) ? 256u : 0u);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t acc = (it < (a.two_acc ? 2 : 1) && ks == 0) ? 0u : 1u;
                    if (a.layout == ATMEM) mma_bf16_ts(d, tmem + kAcol + ks * 8, bd[ks], idesc, acc);
                    else mma_bf16(d, ad[ks], bd[ks], idesc, acc);
                }
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(bar);
        __syncwarp();
    }
    mbar_wait(bar, 0);
    fence_after();
    if (warp == 0) {
        t1 = clock64();
        if (lane == 0) a.cycles[0] = t1 - t0;
    }
    // read the accumulator back: lane = row within the warp's 32-lane quarter
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    for (int c = 0; c < a.N; c += 16) {
        float v[16];
        tmem_ld16(trow + c, v);
        for (int j = 0; j < 16; ++j) a.D[row * 256 + c + j] = v[j];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- (4): pixel rows [256 pixels][8 channels] x 2 "image rows" in shared memory; slot j <- pixels dx .. dx+127 of row j
struct CpArgs {
    const bf16* A;   // [kRows][64]: row = pixel, k = 8 * j + c is channel c of image row j (only k < 16 is used)
    const bf16* B;   // [kRows][64]: row = output channel n
    float* D;
    long long* cycles;
    int N, dx, iters, ncp, nmma, variant;
};

__global__ void __launch_bounds__(128, 1) cp_probe_kernel(const CpArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t tR = base, tB = base + kTile, bar = base + 2 * kTile, tptr = bar + 16;
    uint8_t* gen = smem_raw + (base - raw);
    constexpr uint32_t kRowBuf = kRows * 16u;   // one image row: 256 pixels x 16 bytes
    for (int i = threadIdx.x; i < kRows * 16; i += blockDim.x) {
        const int p = i / 16, k = i % 16, j = k >> 3, c = k & 7;
        *reinterpret_cast<bf16*>(gen + j * kRowBuf + p * 16 + c * 2) = a.A[p * kK + k];
    }
    for (int i = threadIdx.x; i < kRows * kK; i += blockDim.x) {
        const int r = i / kK, k = i % kK;
        *reinterpret_cast<bf16*>(gen + kTile + elem_off(SW128, r, k)) = a.B[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tptr, 512);
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
    const uint32_t idesc = idesc_bf16(a.N, 0, 0);
    constexpr uint32_t kAcol = 448;   // A slots: 4 columns each, 16 slots
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        // no swizzle, K-major: one 16-byte column, so LBO is unused; SBO = 128 bytes between 8-pixel groups
        uint64_t rd[2];
        // (variants 1, 2 only matter if variant 0 turns out wrong: other readings of the two offsets for this shape)
        const uint32_t lbo = a.variant == 0 ? 16u : 128u, sbo = a.variant == 1 ? 16u : 128u;
        for (int j = 0; j < 2; ++j) rd[j] = smem_desc(tR + j * kRowBuf + (uint32_t)a.dx * 16u, lbo, sbo, 0);
        const uint64_t bd = make_desc(SW128, tB, 0, 0, 0);
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
            if (elect_one()) {
                for (int c = 0; c < a.ncp; ++c) tmem_cp_128x128b(tmem + kAcol + 4u * (c & 15), rd[c & 1]);
                for (int m = 0; m < a.nmma; ++m)
                    mma_bf16_ts(tmem, tmem + kAcol, bd, idesc, (it == 0 && m == 0) ? 0u : 1u);
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(bar);
        __syncwarp();
    }
    mbar_wait(bar, 0);
    fence_after();
    if (warp == 0) {
        t1 = clock64();
        if (lane == 0) a.cycles[0] = t1 - t0;
    }
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    for (int c = 0; c < a.N; c += 16) {
        float v[16];
        tmem_ld16(trow + c, v);
        for (int j = 0; j < 16; ++j) a.D[row * 256 + c + j] = v[j];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e__ = (x);                                                            \
        if (e__ != cudaSuccess) {                                                         \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

}  // namespace

int main() {
    std::vector<float> hA(kRows * kK), hB(kRows * kK);
    std::vector<bf16> bA(kRows * kK), bB(kRows * kK);
    unsigned s = 12345u;
    auto rnd = [&]() {
        s = s * 1664525u + 1013904223u;
        return (float)((int)((s >> 16) % 5) - 2);   // -2 .. 2: every product and sum is exact in fp32
    };
    for (int i = 0; i < kRows * kK; ++i) {
        hA[i] = rnd(), hB[i] = rnd();
        bA[i] = __float2bfloat16(hA[i]), bB[i] = __float2bfloat16(hB[i]);
    }
    bf16 *dA, *dB;
    float* dD;
    long long* dC;
    CK(cudaMalloc(&dA, sizeof(bf16) * kRows * kK));
    CK(cudaMalloc(&dB, sizeof(bf16) * kRows * kK));
    CK(cudaMalloc(&dD, sizeof(float) * 128 * 256));
    CK(cudaMalloc(&dC, sizeof(long long)));
    CK(cudaMemcpy(dA, bA.data(), sizeof(bf16) * kRows * kK, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, bB.data(), sizeof(bf16) * kRows * kK, cudaMemcpyHostToDevice));
    const int smem = 2 * kTile + 1024 + 64;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> hD(128 * 256);

    auto run = [&](int layout, int N, int r0, int bo, int iters, int two_acc, long long* cyc) {
        ProbeArgs a{dA, dB, dD, dC, layout, N, r0, bo, iters, two_acc};
        CK(cudaMemset(dD, 0, sizeof(float) * 128 * 256));
        probe_kernel<<<1, 128, smem>>>(a);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, sizeof(float) * 128 * 256, cudaMemcpyDeviceToHost));
        if (cyc) CK(cudaMemcpy(cyc, dC, sizeof(long long), cudaMemcpyDeviceToHost));
    };
    auto max_err = [&](int N, int r0, int iters) {
        double worst = 0.0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0.0;
                for (int k = 0; k < kK; ++k) ref += (double)hA[(r0 + m) * kK + k] * hB[n * kK + k];
                ref *= iters;
                const double e = fabs(ref - hD[m * 256 + n]);
                if (e > worst) worst = e;
            }
        return worst;
    };

    printf("== (1) row-shifted start of the A tile, M = 128, N = 64, K = 64 (4 MMAs); max |error| vs CPU (0 = exact)\n");
    printf("%-10s %4s %14s %22s\n", "layout", "r0", "base_offset=0", "base_offset=(addr>>7)&7");
    for (int r0 = 0; r0 <= 9; ++r0) {
        run(SW128, 64, r0, 0, 1, 0, nullptr);
        const double e0 = max_err(64, r0, 1);
        run(SW128, 64, r0, 1, 1, 0, nullptr);
        const double e1 = max_err(64, r0, 1);
        printf("%-10s %4d %14.1f %22.1f\n", "SW128", r0, e0, e1);
    }
    for (int r0 = 0; r0 <= 9; ++r0) {
        run(NOSWZ, 64, r0, 0, 1, 0, nullptr);
        printf("%-10s %4d %14.1f %22s\n", "no-swizzle", r0, max_err(64, r0, 1), "-");
    }

    printf("== (2) cycles per tcgen05.mma (M = 128, K = 16), %d back-to-back accumulating MMAs, one CTA\n", 4 * 512);
    printf("%-10s %5s %12s %12s %10s\n", "layout", "N", "one acc", "two accs", "exact?");
    const char* names[3] = {"SW128", "no-swizzle", "A in TMEM"};
    for (int layout = 0; layout < 3; ++layout) {
        for (int N = 16; N <= 256; N *= 2) {
            long long c1 = 0, c2 = 0;
            run(layout, N, 0, 0, 512, 0, &c1);
            const double e = max_err(N, 0, 512);
            const bool two_ok = !(layout == ATMEM && N == 256);   // the second accumulator would overlap the A columns
            if (two_ok) run(layout, N, 0, 0, 512, 1, &c2);
            printf("%-10s %5d %12.1f %12.1f %10s\n", names[layout], N, c1 / 2048.0, two_ok ? c2 / 2048.0 : -1.0,
                   e == 0.0 ? "yes" : "NO");
        }
    }

    printf("== (4) A filled by tcgen05.cp.128x128b from [pixel][8 channel] rows (start shifted by dx pixels), K = 16 = two slots\n");
    CK(cudaFuncSetAttribute(cp_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    auto run_cp = [&](int N, int dx, int iters, int ncp, int nmma, long long* cyc, int variant = 0) {
        CpArgs a{dA, dB, dD, dC, N, dx, iters, ncp, nmma, variant};
        CK(cudaMemset(dD, 0, sizeof(float) * 128 * 256));
        cp_probe_kernel<<<1, 128, smem>>>(a);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, sizeof(float) * 128 * 256, cudaMemcpyDeviceToHost));
        if (cyc) CK(cudaMemcpy(cyc, dC, sizeof(long long), cudaMemcpyDeviceToHost));
    };
    for (int variant = 0; variant < 3; ++variant)
    for (int dx = 0; dx <= 2; ++dx) {
        run_cp(16, dx, 1, 2, 1, nullptr, variant);   // slots 0, 1 <- image rows 0, 1; one MMA over both
        double worst = 0.0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
                double ref = 0.0;
                for (int k = 0; k < 16; ++k) ref += (double)hA[(m + dx) * kK + k] * hB[n * kK + k];
                const double e = fabs(ref - hD[m * 256 + n]);
                if (e > worst) worst = e;
            }
        printf("   descriptor (LBO, SBO) = %s, dx %d: max |error| vs CPU %.1f %s\n",
               variant == 0 ? "(16, 128)" : variant == 1 ? "(128, 16)" : "(128, 128)", dx, worst, worst == 0.0 ? "(exact)" : "(WRONG)");
    }
    printf("%5s %5s %6s %16s\n", "N", "ncp", "nmma", "cycles/iteration");
    const int mixes[5][2] = {{0, 6}, {3, 0}, {3, 6}, {6, 6}, {3, 9}};
    for (int N = 16; N <= 64; N *= 2)
        for (int i = 0; i < 5; ++i) {
            long long c = 0;
            run_cp(N, 0, 512, mixes[i][0], mixes[i][1], &c);
            printf("%5d %5d %6d %16.1f\n", N, mixes[i][0], mixes[i][1], c / 512.0);
        }
    return 0;
}

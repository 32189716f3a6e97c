// mnmajor_probe.cu -- standalone hardware probe (development aid, not part of libpgk) for the transposer-free thin
// weight gradient: can tcgen05.mma read image rows kept as [pixel][8 channels] (16 bytes per pixel, the
// layout TMA writes) DIRECTLY as MN-major operands of a reduction over pixels?
//
//   dW[(slot, co)][(kx, ci)] += sum over 128 pixels p of  G[slot][p][co] * X[p + kx][ci]
//
//   A = G ring: M groups of 8 channels = (slot, channel group), `gp` bytes apart        -> MN-group stride = gp
//   B = X row:  N groups of 8 channels = the row shifted by kx pixels, 16 bytes apart   -> MN-group stride = 16 (the
//       8-pixel core matrices of neighbouring groups OVERLAP in memory; un-swizzled descriptors only compute addresses)
//   both: K = 16 pixels per instruction = two 8-pixel core matrices 128 bytes apart     -> K-group stride = 128
//
// Questions: (a) which of the descriptor's two offsets is the MN-group stride for un-swizzled MN-major operands
// (variant 0: SBO = MN stride, LBO = K stride, as cute's canonical INTERLEAVE layout reads; variant 1: swapped), and are
// overlapping groups read correctly; (b) where do the rows of an M = 64 accumulator live in tensor memory; (c) cycles
// per instruction at M = 64 / 128 and N = 24 ... 64 with both operands streamed from shared memory.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mnmajor_probe tools/probes/mnmajor_probe.cu && /tmp/mnmajor_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../pggan-pytorch_b200/csrc/pgk_tc.cuh"

using namespace tc;

namespace {

constexpr int kGPix = 128, kXPix = 160;
constexpr uint32_t kGMax = 4 * 128 * 128 + 8 * 128 * 16, kXMax = kXPix * 128u;   // bytes reserved for the G ring / the X row

struct Args {
    const bf16* G;   // [slots][128 px][Cg]
    const bf16* X;   // [160 px][Cx]
    float* D;        // [128 lanes][256 columns]
    long long* cycles;
    int Cg, Cx, M, shifts, variant, iters;
};

__host__ __device__ inline int swz_bits(int C) { return C == 8 ? 0 : C == 16 ? 1 : C == 32 ? 2 : 3; }
__host__ __device__ inline uint32_t layout_code(int C) { return C == 8 ? 0u : C == 16 ? 6u : C == 32 ? 4u : 2u; }
// the hardware pattern Swizzle<B,4,3> on the absolute shared-memory byte address
__device__ __forceinline__ uint32_t swz(uint32_t addr, int B) { return addr ^ (((addr >> 7) & ((1u << B) - 1u)) << 4); }

__global__ void __launch_bounds__(128, 1) probe(const Args a) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t tG = base, tX = base + kGMax, bar = tX + kXMax, tptr = bar + 16;
    const int slots = a.M / a.Cg;
    const uint32_t cbg = 2u * a.Cg, cbx = 2u * a.Cx, slot_pitch = kGPix * cbg;
    for (int i = threadIdx.x; i < slots * kGPix * a.Cg; i += blockDim.x) {
        const uint32_t ad = swz(tG + 2u * i, swz_bits(a.Cg));     // natural [slot][px][Cg], swizzled on the address
        *reinterpret_cast<bf16*>(smem_raw + (ad - raw)) = a.G[i];
    }
    for (int i = threadIdx.x; i < kXPix * a.Cx; i += blockDim.x) {
        const uint32_t ad = swz(tX + 2u * i, swz_bits(a.Cx));
        *reinterpret_cast<bf16*>(smem_raw + (ad - raw)) = a.X[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tptr, 256);
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
    const int N = a.shifts * a.Cx;
    // both operands MN-major (bits 15, 16); M in bits 24..28 as M >> 4
    const uint32_t idesc = (idesc_bf16(N, 1, 1) & ~(0x1Fu << 24)) | ((uint32_t)(a.M >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        // MN-group stride / K-group (8 pixels) stride of each operand
        const uint32_t a_mn = a.Cg == 8 ? slot_pitch : slot_pitch, a_k = 8u * cbg;
        const uint32_t b_mn = cbx, b_k = 8u * cbx;
        // un-swizzled (8 channels): variant 0 reads SBO = MN stride, LBO = K stride; swizzled: variant 0 reads LBO = MN
        // stride, SBO = K stride (cute's canonical MN-major layouts); variant 1 swaps the two
        auto desc = [&](uint32_t start, uint32_t mn, uint32_t k, int C) {
            const bool lbo_is_mn = (C == 8) == (a.variant == 1);
            return smem_desc(start, lbo_is_mn ? mn : k, lbo_is_mn ? k : mn, layout_code(C));
        };
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    mma_bf16(tmem, desc(tG + ks * 16u * cbg, a_mn, a_k, a.Cg), desc(tX + ks * 16u * cbx, b_mn, b_k, a.Cx),
                             idesc, (it == 0 && ks == 0) ? 0u : 1u);
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
    for (int c = 0; c < 256; c += 16) {
        float v[16];
        tmem_ld16(trow + c, v);
        for (int j = 0; j < 16; ++j) a.D[row * 256 + c + j] = v[j];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
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
    const size_t nG = 16 * kGPix * 64, nX = kXPix * 64;
    std::vector<float> hG(nG), hX(nX);
    std::vector<bf16> bG(nG), bX(nX);
    unsigned s = 2024u;
    auto rnd = [&]() {
        s = s * 1664525u + 1013904223u;
        return (float)((int)((s >> 16) % 5) - 2);
    };
    for (size_t i = 0; i < nG; ++i) hG[i] = rnd(), bG[i] = __float2bfloat16(hG[i]);
    for (size_t i = 0; i < nX; ++i) hX[i] = rnd(), bX[i] = __float2bfloat16(hX[i]);
    bf16 *dG, *dX;
    float* dD;
    long long* dC;
    CK(cudaMalloc(&dG, sizeof(bf16) * nG));
    CK(cudaMalloc(&dX, sizeof(bf16) * nX));
    CK(cudaMalloc(&dD, sizeof(float) * 128 * 256));
    CK(cudaMalloc(&dC, sizeof(long long)));
    CK(cudaMemcpy(dG, bG.data(), sizeof(bf16) * nG, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dX, bX.data(), sizeof(bf16) * nX, cudaMemcpyHostToDevice));
    const int smem = kGMax + kXMax + 1024 + 1024 + 64;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> hD(128 * 256);
    auto run = [&](int Cg, int Cx, int M, int shifts, int variant, int iters, long long* cyc) {
        Args a{dG, dX, dD, dC, Cg, Cx, M, shifts, variant, iters};
        CK(cudaMemset(dD, 0, sizeof(float) * 128 * 256));
        probe<<<1, 128, smem>>>(a);
        CK(cudaGetLastError());
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("   Cg %d Cx %d M %d variant %d: %s\n", Cg, Cx, M, variant, cudaGetErrorString(e));
            exit(1);
        }
        CK(cudaMemcpy(hD.data(), dD, sizeof(float) * 128 * 256, cudaMemcpyDeviceToHost));
        if (cyc) CK(cudaMemcpy(cyc, dC, sizeof(long long), cudaMemcpyDeviceToHost));
    };
    printf("== (a, b) numerics, one pass over 128 pixels (8 MMAs of K = 16); A = G ring [slot][px][Cg], B = X row [px][Cx] read at 3 pixel shifts\n");
    const int cases[][2] = {{8, 8}, {16, 8}, {16, 16}, {32, 16}, {16, 32}, {32, 32}, {64, 32}, {32, 64}, {64, 64}};
    for (int variant = 0; variant < 2; ++variant)
        for (auto& cs : cases) {
            const int Cg = cs[0], Cx = cs[1], M = 4 * Cg > 128 ? 128 : (4 * Cg < 64 ? 64 : 4 * Cg), N = 3 * Cx;
            if (M == 128 && N % 16) continue;
            run(Cg, Cx, M, 3, variant, 1, nullptr);
            auto expect = [&](int m, int n) {
                const int slot = m / Cg, co = m % Cg, sh = n / Cx, ci = n % Cx;
                double r = 0.0;
                for (int p = 0; p < 128; ++p) r += (double)hG[((size_t)slot * kGPix + p) * Cg + co] * hX[(size_t)(p + sh) * Cx + ci];
                return (float)r;
            };
            int found = 0, ident = 0;
            int lane_of[128];
            for (int m = 0; m < M; ++m) {
                lane_of[m] = -1;
                for (int l = 0; l < 128 && lane_of[m] < 0; ++l) {
                    bool ok = true;
                    for (int n = 0; n < N && ok; ++n) ok = hD[l * 256 + n] == expect(m, n);
                    if (ok) lane_of[m] = l;
                }
                found += lane_of[m] >= 0;
                ident += lane_of[m] == m;
            }
            printf("   variant %d  Cg %2d Cx %2d  M %3d N %3d: %3d of %3d rows exact, %3d at lane = row;  rows 0,15,16,31,32,47,48,63 -> lanes %d %d %d %d %d %d %d %d\n",
                   variant, Cg, Cx, M, N, found, M, ident, lane_of[0], lane_of[15], lane_of[16], lane_of[31], lane_of[32],
                   lane_of[47], lane_of[48], lane_of[63]);
        }
    printf("== (c) cycles per tcgen05.mma, both operands MN-major from shared memory, K = 16, 2048 back to back\n");
    for (auto& cs : cases) {
        const int Cg = cs[0], Cx = cs[1], M = 4 * Cg > 128 ? 128 : (4 * Cg < 64 ? 64 : 4 * Cg), N = 3 * Cx;
        if (M == 128 && N % 16) continue;
        long long c = 0;
        run(Cg, Cx, M, 3, 0, 256, &c);
        printf("   Cg %2d Cx %2d  M %3d N %3d: %.1f cycles\n", Cg, Cx, M, N, c / 2048.0);
    }
    return 0;
}

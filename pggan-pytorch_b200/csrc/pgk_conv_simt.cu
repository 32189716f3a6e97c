// pgk_conv_simt.cu -- fp32-accumulate implicit-GEMM convolution, weight gradient and bias gradient on CUDA
// cores.  This is the shape-agnostic path: every layer can run through it (any H, W incl. 1x1 and 4x4, any
// C % 8 == 0); the tcgen05 path (pgk_conv_tc.cu) takes over the tensor-core friendly layers.
//
// GEMM view (network.py:34):  out[m][co] = sum_k A[m][k] * wf[k][co]
//    m = (n, y, x) linearised, k = (tap, ci), A gathered on the fly from the planes tensor (never materialised),
//    nearest-upsample folded into the gather (network.py:127,129).
#include "pgk_common.cuh"

namespace {

constexpr int BM = 128;  // pixels per CTA
constexpr int BN = 64;   // output channels per CTA
constexpr int BK = 16;   // reduction slice (two 8-channel chunks)

struct ConvArgs {
    Planes x;
    int N, H, W, Cin, Cout, KS, ups;
    const float* wf;
    const float* bias;
    const float* posT;
    const float* pos_s;
    int act;
    Planes mask;
    int has_mask;
    float out_scale;
    Planes out;
    int M, K;
};

__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
    pgk_pdl_enter();
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int t = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int HW = a.H * a.W;
    const int Hi = a.ups ? a.H >> 1 : a.H, Wi = a.ups ? a.W >> 1 : a.W;
    const int pad = a.KS >> 1;

    // ---- A loader role: one (pixel row, 8-channel chunk) per thread
    const int arow = t & (BM - 1), ach = t >> 7;
    const int am = m0 + arow;
    const bool am_ok = am < a.M;
    int an = 0, ay = 0, ax = 0;
    if (am_ok) {
        an = am / HW;
        int r = am - an * HW;
        ay = r / a.W;
        ax = r - ay * a.W;
    }
    const long long abase = (long long)an * Hi * Wi * a.Cin;
    // ---- B loader role: one float4 of wf per thread
    const int bk = t >> 4, bc = (t & 15) * 4;

    float areg[8];
    float4 breg;
    auto load_tile = [&](int kt) {
        int k = kt * BK + ach * 8;
        bool ok = am_ok && k < a.K;
        if (ok) {
            int tap = k / a.Cin;
            int ci = k - tap * a.Cin;
            int ky = tap / a.KS;
            int iy = ay + ky - pad, ix = ax + (tap - ky * a.KS) - pad;
            ok = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
            if (ok) {
                if (a.ups) {
                    iy >>= 1;
                    ix >>= 1;
                }
                ld8(a.x, abase + ((long long)iy * Wi + ix) * a.Cin + ci, areg);
            }
        }
        if (!ok) {
#pragma unroll
            for (int j = 0; j < 8; ++j) areg[j] = 0.f;
        }
        int kb = kt * BK + bk, co = n0 + bc;
        if (kb < a.K && co < a.Cout)
            breg = __ldg(reinterpret_cast<const float4*>(a.wf + (long long)kb * a.Cout + co));
        else
            breg = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 8; ++j) As[buf][ach * 8 + j][arow] = areg[j];
        *reinterpret_cast<float4*>(&Bs[buf][bk][bc]) = breg;
    };

    const int tx = t & 15, ty = t >> 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk = (a.K + BK - 1) / BK;
    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tile(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tile(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue
    const int co = n0 + tx * 4;
    if (co >= a.Cout) return;
    float bias4[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.bias) {
        float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + co));
        bias4[0] = b.x, bias4[1] = b.y, bias4[2] = b.z, bias4[3] = b.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + ty * 8 + i;
        if (m >= a.M) break;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias4[j];
        if (a.posT) {
            int n = m / HW;
            int r = m - n * HW;
            float s = __ldg(a.pos_s + n);
            float4 p = __ldg(reinterpret_cast<const float4*>(a.posT + (long long)r * a.Cout + co));
            v[0] = fmaf(s, p.x, v[0]), v[1] = fmaf(s, p.y, v[1]), v[2] = fmaf(s, p.z, v[2]), v[3] = fmaf(s, p.w, v[3]);
        }
        if (a.act) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = lrelu(v[j]);
        }
        long long o = (long long)m * a.Cout + co;
        if (a.has_mask) {
            float r4[4];
            Planes m0 = a.mask;   // the sign of plane 0 is the sign of the value
            m0.P = 1;
            ld4(m0, o, r4);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] *= lrelu_grad(r4[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] *= a.out_scale;
        split_store4(a.out, o, v);
    }
}

// ------------------------------------------------------------------------------------------
// weight gradient: dwp[k][co] += sum_r A[r][k] * g[r][co]   (r runs over pixels of the listed groups)
// ------------------------------------------------------------------------------------------
constexpr int WK = 128;  // k rows per CTA
constexpr int WN = 64;   // co per CTA
constexpr int WR = 16;   // pixels per reduction step

struct WgradArgs {
    Planes x, g;
    int H, W, Cin, Cout, KS, ups;
    int ngroups, group_n;
    int xoff[4], goff[4];
    float* dwp;
    int K;
    long long R;        // total pixels in the reduction
    long long r_per_cta;  // multiple of WR
};

__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradArgs a) {
    pgk_pdl_enter();
    __shared__ __align__(16) float As[2][WR][WK];
    __shared__ __align__(16) float Gs[2][WR][WN];
    const int t = threadIdx.x;
    const int k0 = blockIdx.x * WK, n0 = blockIdx.y * WN;
    const long long r_begin = (long long)blockIdx.z * a.r_per_cta;
    long long r_end = r_begin + a.r_per_cta;
    if (r_end > a.R) r_end = a.R;
    if (r_begin >= r_end) return;
    const int HW = a.H * a.W;
    const long long per_group = (long long)a.group_n * HW;
    const int Hi = a.ups ? a.H >> 1 : a.H, Wi = a.ups ? a.W >> 1 : a.W;
    const int pad = a.KS >> 1;

    // A loader: pixel = t >> 4, chunk = t & 15 ; tap/ci of this thread's chunk are fixed for the whole loop
    const int apix = t >> 4, achunk = t & 15;
    const int ak = k0 + achunk * 8;
    const bool ak_ok = ak < a.K;
    int atap = 0, aci = 0, ady = 0, adx = 0;
    if (ak_ok) {
        atap = ak / a.Cin;
        aci = ak - atap * a.Cin;
        int ky = atap / a.KS;
        ady = ky - pad;
        adx = atap - ky * a.KS - pad;
    }
    // G loader (threads 0..127): pixel = t >> 3, chunk = t & 7
    const int gpix = t >> 3, gchunk = t & 7;
    const int gco = n0 + gchunk * 8;
    const bool g_ok = t < 128 && gco < a.Cout;

    float areg[8], greg[8];
    auto load_tile = [&](long long r0) {
        {
            long long r = r0 + apix;
            bool ok = ak_ok && r < r_end;
            if (ok) {
                int grp = (int)(r / per_group);
                long long w = r - grp * per_group;
                int s = (int)(w / HW);
                int pix = (int)(w - (long long)s * HW);
                int y = pix / a.W, x = pix - y * a.W;
                int iy = y + ady, ix = x + adx;
                ok = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
                if (ok) {
                    if (a.ups) {
                        iy >>= 1;
                        ix >>= 1;
                    }
                    long long n = a.xoff[grp] + s;
                    ld8(a.x, ((n * Hi + iy) * Wi + ix) * a.Cin + aci, areg);
                }
            }
            if (!ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) areg[j] = 0.f;
            }
        }
        if (t < 128) {
            long long r = r0 + gpix;
            bool ok = g_ok && r < r_end;
            if (ok) {
                int grp = (int)(r / per_group);
                long long w = r - grp * per_group;
                long long n = a.goff[grp];
                ld8(a.g, (n * HW + w) * a.Cout + gco, greg);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) greg[j] = 0.f;
            }
        }
    };
    auto store_tile = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][apix][achunk * 8]) = make_float4(areg[0], areg[1], areg[2], areg[3]);
        *reinterpret_cast<float4*>(&As[buf][apix][achunk * 8 + 4]) = make_float4(areg[4], areg[5], areg[6], areg[7]);
        if (t < 128) {
            *reinterpret_cast<float4*>(&Gs[buf][gpix][gchunk * 8]) = make_float4(greg[0], greg[1], greg[2], greg[3]);
            *reinterpret_cast<float4*>(&Gs[buf][gpix][gchunk * 8 + 4]) = make_float4(greg[4], greg[5], greg[6], greg[7]);
        }
    };

    const int tc = t & 15, tk = t >> 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nsteps = (int)((r_end - r_begin + WR - 1) / WR);
    load_tile(r_begin);
    store_tile(0);
    __syncthreads();
    for (int s = 0; s < nsteps; ++s) {
        const int buf = s & 1;
        if (s + 1 < nsteps) load_tile(r_begin + (long long)(s + 1) * WR);
#pragma unroll
        for (int r = 0; r < WR; ++r) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][r][tk * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][r][tk * 8 + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Gs[buf][r][tc * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (s + 1 < nsteps) store_tile(buf ^ 1);
        __syncthreads();
    }
    const int co = n0 + tc * 4;
    if (co >= a.Cout) return;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int k = k0 + tk * 8 + i;
        if (k >= a.K) break;
        float* d = a.dwp + (long long)k * a.Cout + co;
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(d + j, acc[i][j]);
    }
}

// ------------------------------------------------------------------------------------------
// bias gradient: db[co] (+)= scale * sum_pixels g[pix][co]
// ------------------------------------------------------------------------------------------
struct BiasArgs {
    Planes g;
    int HW, Cout, ngroups, group_n;
    int goff[4];
    float scale;
    float* db;
    long long R, r_per_cta;
};

__global__ void __launch_bounds__(256) bias_grad_kernel(BiasArgs a) {
    pgk_pdl_enter();
    __shared__ float red[256][9];
    const int nch = a.Cout >> 3;  // <= 256
    const int t = threadIdx.x;
    const int lanes = 256 / nch;  // pixel lanes
    const int ch = t % nch, pl = t / nch;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // R < 2^31 (checked by the launcher): 32-bit index arithmetic keeps the division out of the load loop's way
    const unsigned per_group = (unsigned)a.group_n * (unsigned)a.HW;
    unsigned r_begin = (unsigned)((long long)blockIdx.x * a.r_per_cta);
    unsigned r_end = (unsigned)(((long long)blockIdx.x + 1) * a.r_per_cta < a.R ? ((long long)blockIdx.x + 1) * a.r_per_cta : a.R);
    if (pl < lanes) {
#pragma unroll 4
        for (unsigned r = r_begin + pl; r < r_end; r += lanes) {
            const unsigned grp = r / per_group;
            const unsigned w = r - grp * per_group;
            float f[8];
            ld8(a.g, ((long long)a.goff[grp] * a.HW + w) * a.Cout + ch * 8, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += f[j];
        }
    }
    // nch a power of two <= 32: lanes l, l + nch, ... of a warp share a channel chunk -> xor shuffles, then the 8 warp
    // partials through shared memory; other widths: one thread per chunk walks its lanes
    const bool fast = nch <= 32 && (nch & (nch - 1)) == 0;
    const int lane = t & 31;
    if (fast) {
        for (int o = nch; o < 32; o <<= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[t][j] = s[j];
    __syncthreads();
    if (fast) {
        if (t < nch * 8) {
            const int chn = t >> 3, j = t & 7;
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += red[w * 32 + chn][j];
            atomicAdd(a.db + chn * 8 + j, a.scale * tot);
        }
    } else if (t < nch) {
        float tot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int l = 0; l < lanes; ++l)
#pragma unroll
            for (int j = 0; j < 8; ++j) tot[j] += red[l * nch + t][j];
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(a.db + t * 8 + j, a.scale * tot[j]);
    }
}

}  // namespace

extern "C" int pgk_conv_simt(const void* x, int P, long long x_ps, int N, int H, int W, int Cin, int Cout, int KS,
                             int ups, const float* wf, const float* bias, const float* posT, const float* pos_s,
                             int act, const void* mask_ref, long long mask_ps, float out_scale, void* out,
                             long long out_ps, pgk_stream_t stream) {
    PGK_REQUIRE(P >= 1 && P <= 3, "pgk_conv: P must be 1, 2 or 3 (got %d)", P);
    PGK_REQUIRE(KS == 1 || KS == 3, "pgk_conv: KS must be 1 or 3 (got %d)", KS);
    PGK_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, "pgk_conv: channels must be multiples of 8 (Cin %d Cout %d)", Cin, Cout);
    PGK_REQUIRE(!ups || (H % 2 == 0 && W % 2 == 0), "pgk_conv: ups needs even H, W");
    PGK_REQUIRE(N > 0 && H > 0 && W > 0, "pgk_conv: empty tensor");
    PGK_REQUIRE((posT == nullptr) == (pos_s == nullptr), "pgk_conv: posT and pos_s go together");
    long long M = (long long)N * H * W;
    PGK_REQUIRE(M < (1ll << 31), "pgk_conv: N*H*W too large");
    ConvArgs a;
    a.x = make_planes(x, x_ps, P);
    a.N = N, a.H = H, a.W = W, a.Cin = Cin, a.Cout = Cout, a.KS = KS, a.ups = ups;
    a.wf = wf, a.bias = bias, a.posT = posT, a.pos_s = pos_s, a.act = act;
    a.mask = make_planes(mask_ref, mask_ps, P);
    a.has_mask = mask_ref != nullptr;
    a.out_scale = out_scale;
    a.out = make_planes(out, out_ps, P);
    a.M = (int)M;
    a.K = KS * KS * Cin;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((Cout + BN - 1) / BN));
    pgk_launch(conv_simt_kernel, grid, 256, 0, (cudaStream_t)stream, a);
    PGK_LAUNCH_CHECK("pgk_conv(simt)");
    return PGK_OK;
}

extern "C" int pgk_wgrad_simt(const void* x, long long x_ps, const void* g, long long g_ps, int P, int H, int W,
                              int Cin, int Cout, int KS, int ups, int ngroups, int group_n, const int* xoff,
                              const int* goff, float* dwp, pgk_stream_t stream) {
    PGK_REQUIRE(P >= 1 && P <= 3, "pgk_wgrad: P must be 1, 2 or 3");
    PGK_REQUIRE(KS == 1 || KS == 3, "pgk_wgrad: KS must be 1 or 3");
    PGK_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, "pgk_wgrad: channels must be multiples of 8");
    PGK_REQUIRE(ngroups >= 1 && ngroups <= 4 && group_n > 0, "pgk_wgrad: 1..4 groups");
    WgradArgs a;
    a.x = make_planes(x, x_ps, P);
    a.g = make_planes(g, g_ps, P);
    a.H = H, a.W = W, a.Cin = Cin, a.Cout = Cout, a.KS = KS, a.ups = ups;
    a.ngroups = ngroups, a.group_n = group_n;
    for (int i = 0; i < 4; ++i) {
        a.xoff[i] = i < ngroups ? xoff[i] : 0;
        a.goff[i] = i < ngroups ? goff[i] : 0;
    }
    a.dwp = dwp;
    a.K = KS * KS * Cin;
    a.R = (long long)ngroups * group_n * H * W;
    int gx = (a.K + WK - 1) / WK, gy = (Cout + WN - 1) / WN;
    // split the pixel reduction so the grid covers ~4 waves, but keep >= 256 pixels per CTA
    long long want = (4ll * pgk_num_sms() + gx * gy - 1) / (gx * gy);
    long long max_split = (a.R + 255) / 256;
    if (want > max_split) want = max_split;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    long long per = (a.R + want - 1) / want;
    per = (per + WR - 1) / WR * WR;
    int gz = (int)((a.R + per - 1) / per);
    a.r_per_cta = per;
    dim3 grid(gx, gy, gz);
    pgk_launch(wgrad_simt_kernel, grid, 256, 0, (cudaStream_t)stream, a);
    PGK_LAUNCH_CHECK("pgk_wgrad(simt)");
    return PGK_OK;
}

extern "C" int pgk_bias_grad(const void* g, long long g_ps, int P, int HW, int Cout, int ngroups, int group_n,
                             const int* goff, float scale, float* db, int accumulate, pgk_stream_t stream) {
    PGK_REQUIRE(P >= 1 && P <= 3, "pgk_bias_grad: P must be 1, 2 or 3");
    PGK_REQUIRE(Cout % 8 == 0 && Cout / 8 <= 256, "pgk_bias_grad: Cout must be a multiple of 8 and <= 2048");
    PGK_REQUIRE(ngroups >= 1 && ngroups <= 4 && group_n > 0, "pgk_bias_grad: 1..4 groups");
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * Cout, (cudaStream_t)stream);
        if (e != cudaSuccess) {
            pgk_set_error("pgk_bias_grad: memset failed: %s", cudaGetErrorString(e));
            return PGK_ERR_CUDA;
        }
    }
    BiasArgs a;
    a.g = make_planes(g, g_ps, P > 2 ? 2 : P);   // gradients: 16 mantissa bits suffice (see pgk_conv: Pr)
    a.HW = HW, a.Cout = Cout, a.ngroups = ngroups, a.group_n = group_n;
    for (int i = 0; i < 4; ++i) a.goff[i] = i < ngroups ? goff[i] : 0;
    a.scale = scale;
    a.db = db;
    a.R = (long long)ngroups * group_n * HW;
    PGK_REQUIRE(a.R < (1ll << 31), "pgk_bias_grad: more than 2^31 pixels");
    // every thread walks r_per_cta / lanes pixels one dependent 16-byte load at a time, so small tensors need many
    // short CTAs (a 512-channel 4x4 .. 16x16 level at batch 4 took 75 us on two CTAs): ~8 loads per thread
    const int nch = Cout >> 3, lanes = 256 / nch > 0 ? 256 / nch : 1;
    long long ctas = (a.R + 8ll * lanes - 1) / (8ll * lanes);
    long long cap = 4ll * pgk_num_sms();
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    a.r_per_cta = (a.R + ctas - 1) / ctas;
    pgk_launch(bias_grad_kernel, dim3((unsigned)ctas), 256, 0, (cudaStream_t)stream, a);
    PGK_LAUNCH_CHECK("pgk_bias_grad");
    return PGK_OK;
}

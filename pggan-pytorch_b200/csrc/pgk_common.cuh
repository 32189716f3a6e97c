// pgk_common.cuh -- shared device helpers for libpgk (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pgk.h"

#define PGK_LRELU 0.2f

extern "C" void pgk_set_error(const char* fmt, ...);
extern "C" void pgk_count_launch(int n);

#define PGK_REQUIRE(cond, ...)                                  \
    do {                                                        \
        if (!(cond)) {                                          \
            pgk_set_error(__VA_ARGS__);                         \
            return PGK_ERR_ARG;                                 \
        }                                                       \
    } while (0)

#define PGK_LAUNCH_CHECK(name)                                                          \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            pgk_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
            return PGK_ERR_CUDA;                                                        \
        }                                                                               \
        pgk_count_launch(1);                                                            \
    } while (0)

typedef __nv_bfloat16 bf16;

// A planes tensor: value = sum over k < P of p[i + k * ps]   (P = 1, 2 or 3)
struct Planes {
    bf16* p;
    long long ps;
    int P;
};
static inline Planes make_planes(const void* p, long long ps, int P) {
    Planes t;
    t.p = (bf16*)p;
    t.ps = ps;
    t.P = P;
    return t;
}

__device__ __forceinline__ float bf16_bits_to_f(uint32_t b) { return __uint_as_float(b << 16); }

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
    f[0] = __uint_as_float(q.x << 16);
    f[1] = __uint_as_float(q.x & 0xffff0000u);
    f[2] = __uint_as_float(q.y << 16);
    f[3] = __uint_as_float(q.y & 0xffff0000u);
    f[4] = __uint_as_float(q.z << 16);
    f[5] = __uint_as_float(q.z & 0xffff0000u);
    f[6] = __uint_as_float(q.w << 16);
    f[7] = __uint_as_float(q.w & 0xffff0000u);
}

// load 8 consecutive channels starting at element index i (i % 8 == 0); planes are summed smallest first
__device__ __forceinline__ void ld8(const Planes& t, long long i, float* f) {
    uint4 q = __ldg(reinterpret_cast<const uint4*>(t.p + i));
    if (t.P == 1) {
        unpack8(q, f);
        return;
    }
    float h[8], g[8];
    unpack8(q, h);
    uint4 r = __ldg(reinterpret_cast<const uint4*>(t.p + t.ps + i));
    unpack8(r, g);
    if (t.P == 3) {
        float e[8];
        uint4 u = __ldg(reinterpret_cast<const uint4*>(t.p + 2 * t.ps + i));
        unpack8(u, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] += e[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = h[j] + g[j];
}

__device__ __forceinline__ void ld4(const Planes& t, long long i, float* f) {
    uint2 q = __ldg(reinterpret_cast<const uint2*>(t.p + i));
    f[0] = __uint_as_float(q.x << 16);
    f[1] = __uint_as_float(q.x & 0xffff0000u);
    f[2] = __uint_as_float(q.y << 16);
    f[3] = __uint_as_float(q.y & 0xffff0000u);
    if (t.P == 1) return;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    for (int pl = t.P - 1; pl >= 1; --pl) {
        uint2 r = __ldg(reinterpret_cast<const uint2*>(t.p + pl * t.ps + i));
        g[0] += __uint_as_float(r.x << 16);
        g[1] += __uint_as_float(r.x & 0xffff0000u);
        g[2] += __uint_as_float(r.y << 16);
        g[3] += __uint_as_float(r.y & 0xffff0000u);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) f[j] += g[j];
}

__device__ __forceinline__ float ld1(const Planes& t, long long i) {
    float v = __bfloat162float(t.p[i]);
    if (t.P == 1) return v;
    float lo = __bfloat162float(t.p[t.ps + i]);
    if (t.P == 3) lo += __bfloat162float(t.p[2 * t.ps + i]);
    return v + lo;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// split v into P bf16 planes: plane k = bf16_rn(v - sum of the earlier planes)   (8 / 16 / 24 mantissa bits)
__device__ __forceinline__ void split_store8(const Planes& t, long long i, const float* f) {
    float res[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) res[j] = f[j];
    for (int pl = 0; pl < t.P; ++pl) {
        uint4 q;
        q.x = pack2(res[0], res[1]);
        q.y = pack2(res[2], res[3]);
        q.z = pack2(res[4], res[5]);
        q.w = pack2(res[6], res[7]);
        *reinterpret_cast<uint4*>(t.p + pl * t.ps + i) = q;
        if (pl + 1 < t.P) {
            float h[8];
            unpack8(q, h);
#pragma unroll
            for (int j = 0; j < 8; ++j) res[j] -= h[j];
        }
    }
}

__device__ __forceinline__ void split_store4(const Planes& t, long long i, const float* f) {
    float res[4] = {f[0], f[1], f[2], f[3]};
    for (int pl = 0; pl < t.P; ++pl) {
        uint2 q;
        q.x = pack2(res[0], res[1]);
        q.y = pack2(res[2], res[3]);
        *reinterpret_cast<uint2*>(t.p + pl * t.ps + i) = q;
        if (pl + 1 < t.P) {
            res[0] -= __uint_as_float(q.x << 16);
            res[1] -= __uint_as_float(q.x & 0xffff0000u);
            res[2] -= __uint_as_float(q.y << 16);
            res[3] -= __uint_as_float(q.y & 0xffff0000u);
        }
    }
}

__device__ __forceinline__ void st1(const Planes& t, long long i, float v) {
    for (int pl = 0; pl < t.P; ++pl) {
        bf16 h = __float2bfloat16_rn(v);
        t.p[pl * t.ps + i] = h;
        v -= __bfloat162float(h);
    }
}

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : PGK_LRELU * v; }
__device__ __forceinline__ float lrelu_grad(float v) { return v > 0.f ? 1.f : PGK_LRELU; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; result valid in thread 0 (and broadcast through smem to all when bcast)
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 33 floats */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = lane < nw ? sh[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

// ---- programmatic dependent launch ----------------------------------------------------------------------------------
// A training iteration at small batch is several hundred short launches in one stream (depth 8, batch 4: ~620 per
// 15 ms).  With the attribute below a kernel may be scheduled while its predecessor in the stream is still draining:
// its blocks run their prologue (barrier init, tensor-memory allocation, descriptor prefetch) and then block in
// griddepcontrol.wait until the predecessor has completed and its writes are visible -- the launch latency and the
// prologue leave the critical path.  Every kernel of the library executes pgk_pdl_enter() in every thread before its
// first read of global memory (and before any early return), so a chain of such launches stays transitively ordered.
// Measured on B200 (bench.py, round 2): depth 8 / batch 4 15.12 -> 14.18 ms per iteration, depth 6 / batch 32
// 24.33 -> 23.94, depth 0 2.90 -> 2.85; the GPU test-suite is green with it.  On by default; PGK_PDL=0 (or
// pgk_set_pdl(0): CUDA-graph capture turns it off, where it measured slower) launches plainly.
__device__ __forceinline__ void pgk_pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

extern "C" int pgk_pdl_state(int set);   // set < 0: query; 0 / 1: switch off / on; returns the state (pgk_elem.cu)
static inline bool pgk_pdl_enabled() { return pgk_pdl_state(-1) != 0; }

// kern<<<grid, block, smem, stream>>>(args...), or the same launch with programmatic stream serialization allowed
template <typename... KArgs, typename... Args>
static inline void pgk_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    if (pgk_pdl_enabled()) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
        return;
    }
    kern<<<grid, block, smem, stream>>>(static_cast<KArgs>(args)...);
}

static inline int pgk_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Shared helpers for libdtraj (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <utility>

#include "../../include/dtraj.h"

namespace dtraj {

// ---- error plumbing: C ABI never throws, it records a thread-local message -----------
inline char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
#define DTRAJ_CUDA(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::dtraj::fail(DTRAJ_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,  \
                                 cudaGetErrorString(_e));                                  \
    } while (0)
#define DTRAJ_LAUNCH_CHECK()                                                               \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess)                                                             \
            return ::dtraj::fail(DTRAJ_ECUDA, "%s:%d launch -> %s", __FILE__, __LINE__,     \
                                 cudaGetErrorString(_e));                                  \
    } while (0)
#define DTRAJ_TRY(expr)                                                                    \
    do {                                                                                   \
        int _r = (expr);                                                                   \
        if (_r != 0) return _r;                                                            \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr int kCPad = 32;     // channel padding unit of NHWC feature maps (= one 128-byte TMA/UMMA K block)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int64_t round_up64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// activation post-processing chosen by the model's precision mode
enum ActMode : int {
    ACT_PLAIN = 0,   // store y
    ACT_ROUND = 1,   // store rna_tf32(y): operands of the single-pass TF32 convs are exact tf32
    ACT_SPLIT = 2    // store y and, lo_off floats further, y - trunc_tf32(y) for the 3xTF32 convs
};

// ---- tf32 helpers (device + host) --------------------------------------------------------
__host__ __device__ inline float tf32_trunc(float v) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__float_as_uint(v) & 0xffffe000u);
#else
    uint32_t u;
    memcpy(&u, &v, 4);
    u &= 0xffffe000u;
    memcpy(&v, &u, 4);
    return v;
#endif
}
__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
inline float tf32_rna_host(float v) {  // round-to-nearest, ties away (matches cvt.rna)
    uint32_t u;
    memcpy(&u, &v, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return v;
    u += 0x00001000u;
    u &= 0xffffe000u;
    memcpy(&v, &u, 4);
    return v;
}

// y -> what goes to memory under `mode`; writes the low plane when splitting
__device__ __forceinline__ float act_store_value(float y, int mode) {
    return mode == ACT_ROUND ? tf32_rna(y) : y;
}
__device__ __forceinline__ float4 act_round4(float4 v, int mode) {
    if (mode == ACT_ROUND) {
        v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
    }
    return v;
}
__device__ __forceinline__ float4 act_lo4(float4 v) {
    return make_float4(v.x - tf32_trunc(v.x), v.y - tf32_trunc(v.y),
                       v.z - tf32_trunc(v.z), v.w - tf32_trunc(v.w));
}

// ---- programmatic dependent launch (PDL): the kernels of the sampler loop are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, signal `launch_dependents` at their top and `wait` after their
// prologue (barrier init, TMEM allocation, constant staging), so that the next kernel's launch latency and prologue
// overlap the tail of the current one -- also inside the captured CUDA graph.  Both instructions are no-ops for a
// kernel launched without the attribute.  Measured on B200 (bench, fp16 mode, CUDA graph): 132.2 ms per step with the
// attribute against 132.7 ms without -- the step sits at the board's power cap, so closing launch gaps buys almost
// nothing for full-size launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Round 2: small launches are a different case -- a 2-row forward is 19 launches of 1..4 CTAs, ~10 us each of which half is
// prologue and launch latency (batch-1 trajectory 17.3 -> 16.2 ms with the attribute).  Default now: ON for launches of at most
// 64 CTAs, off above; DTRAJ_PDL=0 / 1 force it off / on everywhere.
inline bool use_pdl(unsigned grid) {
    static int v = -2;
    if (v == -2) v = getenv("DTRAJ_PDL") ? (atoi(getenv("DTRAJ_PDL")) != 0) : -1;
    return v == 1 || (v == -1 && grid <= 64u);
}
// <<<grid, block, smem, st>>> with optional cluster width and the PDL attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, int cluster,
                             bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl && use_pdl(grid)) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit loads/stores that do not pollute L1
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

}  // namespace dtraj

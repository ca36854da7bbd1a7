// tcgen05 / TMA implicit-GEMM convolution for sm_100a (DTRAJ_PREC_TF32, DTRAJ_PREC_TF32X3).
//
// GEMM view of one conv layer (models.py:48-57):  D[M, N] = A[M, K] * B[K, N]
//   M = n_img*H*W output pixels, N = coutp, K = ntaps * (c0p + c1p).
// A is never materialised: with NHWC feature maps a 3x3 tap is a SHIFTED 4-d box of the
// input tensor, so one TMA box load {32 ch, W, Hb, Nb} at coordinates
// {c0, dx, y0+dy, img0} delivers the 128 x 32 im2col tile of tap (dy,dx) straight into the
// 128-byte-swizzled K-major layout tcgen05.mma reads; out-of-bounds rows/columns (the
// conv's zero padding) are zero-filled by the TMA unit.  The channel concat of the decoder
// blocks (models.py:206,211,216) is two tensor maps feeding consecutive K blocks.
//
// Roles (192 threads): warp 0 = TMA producer + TMEM allocator, warp 1 = MMA issuer (one
// elected lane), warps 2-5 = epilogue (TMEM -> registers -> bias / ReLU / time-bias /
// residual -> global).  A `stages`-deep mbarrier ring connects producer and issuer; the
// fp32 accumulator (128 lanes x N columns) lives in TMEM.
//
// 3xTF32 (DTRAJ_PREC_TF32X3): the K loop runs three passes  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi
// into the same accumulator, where x_lo = x - trunc_tf32(x) is kept in a second plane by
// every producer of an activation (ACT_SPLIT) and the weights are split on the host.
#pragma once
#include "common.cuh"
#include "conv_simt.cuh"

namespace dtraj {

// set by any role that timed out on an mbarrier (would otherwise hang the GPU)
__device__ unsigned int g_umma_error = 0;

struct UmmaConv {
    ConvLayer L;               // epilogue parameters + shapes (wpk unused here)
    int npass;                 // 1 (TF32) or 3 (TF32X3)
    int stages;
    int tmem_cols;             // power of two >= coutp (x2 for 3 passes: main + correction accumulator)
    int corr_col;              // first TMEM column of the correction accumulator (3 passes), else 0
    int box_h, box_n;          // A box = {32, W, box_h, box_n}
    int tiles_per_img;         // >= 1
    int b_lo_row;              // row offset of the low-plane weights inside the B tensor map
};

struct UmmaMaps {              // 64-byte aligned tensor maps, passed as __grid_constant__
    CUtensorMap a[4];          // [src0 hi, src1 hi, src0 lo, src1 lo]
    CUtensorMap b;
};

namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug must surface as an error flag, never as a hung GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 22); ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    atomicOr(&g_umma_error, 1u);
    return false;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
}  // namespace ptx

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 = 1 [16,30) | SBO>>4 = 1024B/16 [32,46) | version 1 [46,48) | SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, M=128
__host__ __device__ inline uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int kUmmaThreads = 192;
constexpr int kATileBytes = 128 * 128;   // 128 rows x 32 fp32

__global__ void __launch_bounds__(kUmmaThreads, 1)
k_conv_umma(const __grid_constant__ UmmaMaps maps, const UmmaConv p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x (A 16 KB | B coutp*128 B)] then barriers
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int coutp = p.L.coutp;
    const uint32_t b_bytes = (uint32_t)coutp * 128u;
    const uint32_t stage_bytes = kATileBytes + b_bytes;        // multiple of 1024 (coutp % 32 == 0 -> b_bytes % 4096 == 0)
    const uint32_t bar_base = base + p.stages * stage_bytes;   // full[stages], empty[stages], accum, tmem slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
    const uint32_t accum_bar = bar_base + 16u * p.stages;
    const uint32_t tmem_slot = accum_bar + 8u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
        smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nch0 = p.L.c0p / 32, nch = nch0 + p.L.c1p / 32;
    const int iters_per_pass = p.L.ntaps * nch;
    const int n_iters = p.npass * iters_per_pass;

    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::prefetch_tmap(&maps.a[0]);
            ptx::prefetch_tmap(&maps.b);
            if (p.L.c1p) ptx::prefetch_tmap(&maps.a[1]);
            for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
            ptx::mbar_init(accum_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int tile = blockIdx.x;
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            int img0, y0;
            if (p.tiles_per_img > 1) { img0 = tile / p.tiles_per_img; y0 = (tile % p.tiles_per_img) * p.box_h; }
            else { img0 = tile * p.box_n; y0 = 0; }
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                if (!ptx::mbar_wait(empty_bar(s), ph ^ 1u)) break;
                const int pass = it / iters_per_pass, kb = it % iters_per_pass;
                const int tap = kb / nch, chunk = kb % nch;
                int dy = 0, dx = 0;
                if (p.L.ntaps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
                const int src = chunk < nch0 ? 0 : 1;
                const int c0 = (src == 0 ? chunk : chunk - nch0) * 32;
                const CUtensorMap* amap = &maps.a[src + (pass == 2 ? 2 : 0)];
                const uint32_t a_dst = base + s * stage_bytes, b_dst = a_dst + kATileBytes;
                ptx::mbar_expect_tx(full_bar(s), kATileBytes + b_bytes);
                ptx::tma_load_4d(a_dst, amap, full_bar(s), c0, dx, y0 + dy, img0);
                ptx::tma_load_2d(b_dst, &maps.b, full_bar(s), 0, (pass == 1 ? p.b_lo_row : 0) + kb * coutp);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (ptx::elect_one()) {
            const uint32_t idesc = umma_idesc_tf32(coutp);
            bool ok = true;
            for (int it = 0; it < n_iters && ok; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                ok = ptx::mbar_wait(full_bar(s), ph);
                ptx::tc_fence_after();
                const uint32_t a_src = base + s * stage_bytes, b_src = a_src + kATileBytes;
                const uint64_t adesc = umma_desc_sw128(a_src), bdesc = umma_desc_sw128(b_src);
                // 3xTF32: the two small cross terms go to their own accumulator.  The tensor core adds
                // into the accumulator with round-toward-zero; adding 2^-11-sized terms to a full-size
                // accumulator for 2/3 of the K loop costs ~K/16 ulps of systematic shrink (measured 6e-5
                // at K = 4608), a separate accumulator keeps that at the single-pass level.
                const bool corr = it >= iters_per_pass;
                const uint32_t d_tmem = tmem_base + (corr ? (uint32_t)p.corr_col : 0u);
                const int first_it = corr ? iters_per_pass : 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)   // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle atom
                    ptx::mma_tf32(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (it != first_it) || (k != 0));
                ptx::tc_commit(empty_bar(s));   // frees the smem slot when these MMAs retire
            }
            ptx::tc_commit(accum_bar);          // accumulator complete
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int64_t m = (int64_t)tile * 128 + q * 32 + lane;
        ptx::mbar_wait(accum_bar, 0);
        ptx::tc_fence_after();
        const bool valid = m < p.L.M;
        const int HW = p.L.H * p.L.W;
        const float* tb = nullptr;
        if ((p.L.flags & CONV_TBIAS) && valid) {
            const int var = p.L.row_variant ? p.L.row_variant[m / HW] : 0;
            tb = p.L.tbias + (size_t)var * p.L.tb_var_stride;
        }
        for (int c0 = 0; c0 < coutp; c0 += 32) {
            uint32_t raw[32];
            ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, raw);
            if (p.npass == 3) {
                uint32_t raw2[32];
                ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.corr_col + c0), raw2);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[j]));
            } else {
                ptx::tmem_ld_wait();
            }
            if (!valid) continue;
            float* dst = p.L.out + m * coutp + c0;
            const float* res = (p.L.flags & CONV_RESID) ? p.L.resid + m * coutp + c0 : nullptr;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 b4 = __ldg(reinterpret_cast<const float4*>(p.L.bias + c0 + j));
                float4 v = make_float4(__uint_as_float(raw[j]) + b4.x, __uint_as_float(raw[j + 1]) + b4.y,
                                       __uint_as_float(raw[j + 2]) + b4.z, __uint_as_float(raw[j + 3]) + b4.w);
                if (p.L.flags & CONV_RELU) {
                    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                }
                if (tb) {
                    float4 t4 = __ldg(reinterpret_cast<const float4*>(tb + c0 + j));
                    v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w;
                }
                if (res) {
                    float4 r4 = *reinterpret_cast<const float4*>(res + j);
                    v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
                }
                v = act_round4(v, p.L.act_mode);
                *reinterpret_cast<float4*>(dst + j) = v;
                if (p.L.act_mode == ACT_SPLIT) *reinterpret_cast<float4*>(dst + p.L.lo_off + j) = act_lo4(v);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------- host side: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// NHWC activation map {cp, W, H, n_img} with box {32, W, box_h, box_n}
inline int make_act_map(CUtensorMap* m, const float* base, int cp, int W, int H, int64_t n_img, int box_h, int box_n) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(DTRAJ_ECUDA, "cuTensorMapEncodeTiled not available");
    cuuint64_t dims[4] = {(cuuint64_t)cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)cp * 4, (cuuint64_t)W * cp * 4, (cuuint64_t)H * W * cp * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)W, (cuuint32_t)box_h, (cuuint32_t)box_n};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DTRAJ_ECUDA, "cuTensorMapEncodeTiled(act cp=%d W=%d H=%d n=%lld) -> %d", cp, W, H, (long long)n_img, (int)r);
    return 0;
}
// packed weights as a 2-d map {32, rows} with box {32, coutp}
inline int make_w_map(CUtensorMap* m, const float* base, int64_t rows, int coutp) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(DTRAJ_ECUDA, "cuTensorMapEncodeTiled not available");
    cuuint64_t dims[2] = {32, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {32, (cuuint32_t)coutp};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DTRAJ_ECUDA, "cuTensorMapEncodeTiled(weights rows=%lld) -> %d", (long long)rows, (int)r);
    return 0;
}

struct UmmaLaunch {            // everything a launch needs, built once per (layer, batch)
    UmmaMaps maps;
    UmmaConv conv;
    unsigned grid;
    size_t smem;
};

// `w_rows` = rows of the packed weight matrix (hi planes then, for 3 passes, lo planes)
inline int build_umma_launch(UmmaLaunch* U, const ConvLayer& L, int npass, const float* wpk, int64_t w_rows) {
    memset(U, 0, sizeof(*U));
    const int HW = L.H * L.W;
    if (L.W > 32 || (HW < 128 && 128 % HW != 0) || (HW >= 128 && (128 % L.W != 0 || HW % 128 != 0)))
        return fail(DTRAJ_EINVAL, "umma conv: unsupported spatial size %dx%d", L.H, L.W);
    if (L.coutp % 32 || L.coutp > 256 || L.c0p % 32 || L.c1p % 32) return fail(DTRAJ_EINVAL, "umma conv: bad channel padding");
    UmmaConv& c = U->conv;
    c.L = L;
    c.npass = npass;
    c.box_h = HW >= 128 ? 128 / L.W : L.H;
    c.box_n = HW >= 128 ? 1 : 128 / HW;
    c.tiles_per_img = HW >= 128 ? HW / 128 : 1;
    c.tmem_cols = 32;
    while (c.tmem_cols < L.coutp) c.tmem_cols *= 2;
    c.corr_col = 0;
    if (npass == 3) { c.corr_col = c.tmem_cols; c.tmem_cols *= 2; }
    const int nkb = L.ntaps * (L.c0p + L.c1p) / 32;
    c.b_lo_row = nkb * L.coutp;
    const size_t stage = kATileBytes + (size_t)L.coutp * 128;
    // <= ~100 KB so two CTAs share an SM: one runs its epilogue while the other's MMAs run
    int stages = (int)((100 * 1024) / stage);
    if (stages < 2) stages = 2;
    if (stages > 6) stages = 6;
    c.stages = stages;
    U->smem = 1024 + stages * stage + 16 * stages + 16 + 16;
    U->grid = (unsigned)((L.M + 127) / 128);
    const int64_t n_img = L.M / HW;
    DTRAJ_TRY(make_act_map(&U->maps.a[0], L.src0, L.c0p, L.W, L.H, n_img, c.box_h, c.box_n));
    if (L.c1p) DTRAJ_TRY(make_act_map(&U->maps.a[1], L.src1, L.c1p, L.W, L.H, n_img, c.box_h, c.box_n));
    if (npass == 3) {
        DTRAJ_TRY(make_act_map(&U->maps.a[2], L.src0_lo, L.c0p, L.W, L.H, n_img, c.box_h, c.box_n));
        if (L.c1p) DTRAJ_TRY(make_act_map(&U->maps.a[3], L.src1_lo, L.c1p, L.W, L.H, n_img, c.box_h, c.box_n));
    }
    DTRAJ_TRY(make_w_map(&U->maps.b, wpk, w_rows, L.coutp));
    return 0;
}

inline int launch_conv_umma(const UmmaLaunch& U, cudaStream_t st) {
    k_conv_umma<<<U.grid, kUmmaThreads, U.smem, st>>>(U.maps, U.conv);
    DTRAJ_LAUNCH_CHECK();
    return 0;
}

}  // namespace dtraj

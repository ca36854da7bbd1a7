// Fused enc1 block for sm_100a, fp16 mode (DTRAJ_PREC_F16): the same computation as enc1_umma.cuh,
//     h  = relu(bn1(conv1_3x3(x))) + relu(time_mlp(temb))          (models.py:62-77, C = config.channels <= 4)
//     y  = relu(bn2(conv2_3x3(h))) + residual_conv_1x1(x)           (models.py:79-83)
//     p1 = MaxPool2d(2)(y)                                          (models.py:191)
// with two structural differences that fp16 operands make possible:
//   * conv2's packed weights ([tap][64-channel chunk][coutp][64] halfs) are RESIDENT in shared memory: 288 KB for
//     the teacher, 144 KB per CTA of a pair (each CTA holds its half of the output channels, tcgen05.mma.cta_group::2)
//     -- they are loaded once per CTA, nothing streams from L2 inside the tile loop (the tf32 kernel re-read 295 KB
//     of weights per 128-pixel tile, which paced it);
//   * a halo chunk buffer holds 64 channels per 128-byte row, so the teacher's K loop is 2 chunks x 9 taps x 4 MMAs.
// Output tile = 16 image rows x 8 columns; GENERATOR warps compute conv1 on the 18 x 10 halo with CUDA cores and write it
// rounded to fp16 into the 128-byte-swizzled K-major layout; each conv2 tap is a descriptor VIEW of the halo tile
// (start shifted by (dy*10 + dx) rows, 8-row groups 1280 bytes apart; profiles/r01_umma_view_probe.txt).
// Epilogue: bias, ReLU, 1x1 residual recomputed from x, fp16 rounding, 2x2 max-pool, 16-byte stores.
#pragma once
#include "enc1_umma.cuh"

namespace dtraj {

struct Enc1hParams {
    int C, H, W, coutp;
    int n_chunks;                // coutp / 64: K chunks of conv2
    int n_tiles, tiles_x, tiles_per_img;
    int n_hbuf;                  // halo chunk buffers in the ring (>= 1; 2 x n_chunks when they fit)
    int acc_cols;                // TMEM columns per accumulator
    int w_rows;                  // weight rows (output channels) this CTA holds per (tap, chunk): coutp, or coutp / 2 in pair mode
    const float* x; int64_t x_stride; const int32_t* row_sample; const int32_t* row_variant;
    const float* w3; const float* b3;          // conv1, BN folded: [9*C][coutp] tap-major then cin; [coutp]   (fp32)
    const float* tbias; int tb_var_stride;
    const float* bias2;                        // conv2 folded bias [coutp]
    const float* rw1; const float* rb1;        // residual 1x1: [C][coutp], [coutp]
    __half* pool_out;                          // [R, H/2, W/2, coutp]
    int debug;                   // timing experiments (DTRAJ_E1_DEBUG): 1 generators store zeros, 2 generators only signal, 4 no MMAs
};

template <bool kPair>
__global__ void __launch_bounds__(kE1Threads, 1)
k_enc1_f16(const __grid_constant__ Enc1Maps maps, const Enc1hParams p) {
    extern __shared__ __align__(1024) uint8_t e1_smem[];
    const uint32_t base = (ptx::smem_u32(e1_smem) + 1023u) & ~1023u;
    uint8_t* gbase = e1_smem + (base - ptx::smem_u32(e1_smem));
    const int coutp = p.coutp, C = p.C;
    const uint32_t wblk_bytes = (uint32_t)p.w_rows * 128u;                 // one (tap, chunk) weight block in this CTA
    const uint32_t w_bytes = 9u * (uint32_t)p.n_chunks * wblk_bytes;
    // carve: [resident weights][halo ring][epilogue ring 4 x 2 KB][x patches 2 x 4 ch x 240][constants][barriers]
    const uint32_t wres0 = base;
    const uint32_t halo0 = wres0 + w_bytes;                                // w_bytes is a multiple of 1024 (w_rows % 8 == 0, 9 blocks)
    const uint32_t ring0 = halo0 + (uint32_t)p.n_hbuf * kE1HaloBytes;
    const uint32_t xp0 = ring0 + kE1Epi * 2048u;
    const uint32_t cst0 = xp0 + 2u * 4u * 240u * 4u;
    // constants (floats): w3 [9C][coutp] | b3 | bias2 | rb1 | rw1 [C][coutp]
    const int n_cst = (9 * C + 3 + C) * coutp;
    const uint32_t bar0 = (cst0 + (uint32_t)n_cst * 4u + 15u) & ~15u;
    auto hfull = [&](int b) { return bar0 + 8u * b; };
    auto hempty = [&](int b) { return bar0 + 64u + 8u * b; };
    const uint32_t acc_full0 = bar0 + 128u, acc_empty0 = bar0 + 144u, wbar = bar0 + 160u, tmem_slot = bar0 + 168u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
    float* xpatch = reinterpret_cast<float*>(gbase + (xp0 - base));       // [2][4][20][12]
    float* cst = reinterpret_cast<float*>(gbase + (cst0 - base));
    const float* w3s = cst;
    const float* b3s = cst + 9 * C * coutp;
    const float* bias2s = b3s + coutp;
    const float* rb1s = bias2s + coutp;
    const float* rw1s = rb1s + coutp;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = kPair ? (int)ptx::cluster_ctarank() : 0;
    const uint16_t cmask = kPair ? 3 : 1;
    const int work0 = (int)blockIdx.x - crank;

    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::prefetch_tmap(&maps.w);
            for (int b = 0; b < p.n_hbuf; ++b) { ptx::mbar_init(hfull(b), kE1Gen * (kPair ? 2 : 1)); ptx::mbar_init(hempty(b), 1); }
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(acc_full0 + 8u * i, 1);
                ptx::mbar_init(acc_empty0 + 8u * i, kE1Epi * (kPair ? 2 : 1));
            }
            ptx::mbar_init(wbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if constexpr (!kPair) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * p.acc_cols)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * p.acc_cols)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    // constants into shared memory (all threads)
    for (int i = threadIdx.x; i < 9 * C * coutp; i += blockDim.x) cst[i] = p.w3[i];
    for (int i = threadIdx.x; i < coutp; i += blockDim.x) {
        cst[9 * C * coutp + i] = p.b3[i];
        cst[(9 * C + 1) * coutp + i] = p.bias2[i];
        cst[(9 * C + 2) * coutp + i] = p.rb1[i];
    }
    for (int i = threadIdx.x; i < C * coutp; i += blockDim.x) cst[(9 * C + 3) * coutp + i] = p.rw1[i];
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (kPair) ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    auto tile_geom = [&](int tile, int& img, int& y0, int& x0) {
        img = tile / p.tiles_per_img;
        const int r = tile - img * p.tiles_per_img;
        y0 = (r / p.tiles_x) * 16;
        x0 = (r % p.tiles_x) * 8;
    };

    if (warp == 0) {
        // ------------------------------------------------------------ one-time TMA load of the resident conv2 weights
        if (ptx::elect_one()) {
            uint32_t fb = wbar;
            if constexpr (kPair) fb = ptx::map_to_cta(fb, 0);              // both CTAs' bytes complete on the leader's barrier
            if (!kPair || crank == 0) ptx::mbar_expect_tx(wbar, w_bytes * (kPair ? 2u : 1u));
            for (int b = 0; b < 9 * p.n_chunks; ++b) {                     // block b = tap * n_chunks + chunk
                const int row = b * coutp + crank * p.w_rows;
                if constexpr (kPair) ptx::tma_load_2d_2sm(wres0 + b * wblk_bytes, &maps.w, fb, 0, row);
                else ptx::tma_load_2d(wres0 + b * wblk_bytes, &maps.w, fb, 0, row);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (pair: leader CTA only)
        if (ptx::elect_one() && (!kPair || crank == 0)) {
            const uint32_t idesc = umma_idesc_f16(coutp) + (kPair ? ((uint32_t)(128 >> 4) << 24) : 0u);
            // halo view: K-major SWIZZLE_128B, 8-row groups one halo row (10 pixels = 1280 B) apart
            const uint64_t hdesc0 = ((uint64_t)1 << 16) | ((uint64_t)(1280 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            int hb = 0, acc = 0;
            uint32_t hph = 0, acc_ph = 0;
            bool ok = ptx::mbar_wait(wbar, 0u);
            ptx::tc_fence_after();
            for (int wk = work0; wk < p.n_tiles && ok; wk += gridDim.x) {
                ok = ptx::mbar_wait(acc_empty0 + 8u * acc, acc_ph ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_cols);
                uint32_t accum = 0u;
                for (int c = 0; c < p.n_chunks && ok; ++c) {
                    ok = ptx::mbar_wait(hfull(hb), hph);            // this chunk's halo tile is in shared memory (both CTAs)
                    ptx::tc_fence_after();
                    const uint32_t hbuf = halo0 + (uint32_t)hb * kE1HaloBytes;
                    int dy = 0, dx = 0;
                    for (int t = 0; t < 9; ++t) {
                        const uint32_t a_addr = hbuf + (uint32_t)(dy * 10 + dx) * 128u;
                        const uint64_t ad = hdesc0 | (uint64_t)((a_addr >> 4) & 0x3fffu);
                        const uint64_t bd = umma_desc_sw128(wres0 + (uint32_t)(t * p.n_chunks + c) * wblk_bytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (p.debug & 4) break;
                            if constexpr (!kPair) ptx::mma_f16(d_tmem, ad + 2u * k, bd + 2u * k, idesc, accum);
                            else ptx::mma_f16_2sm(d_tmem, ad + 2u * k, bd + 2u * k, idesc, accum);
                            accum = 1u;
                        }
                        if (++dx == 3) { dx = 0; ++dy; }
                    }
                    if constexpr (kPair) ptx::tc_commit_2sm(hempty(hb), cmask); else ptx::tc_commit(hempty(hb));
                    if (++hb == p.n_hbuf) { hb = 0; hph ^= 1u; }
                }
                if constexpr (kPair) ptx::tc_commit_2sm(acc_full0 + 8u * acc, cmask); else ptx::tc_commit(acc_full0 + 8u * acc);
                if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
            }
        }
    } else if (warp >= 2 + kE1Epi) {
        // ------------------------------------------------------------ generators: conv1 + BN + ReLU + time bias -> halo tiles
        const int gt = threadIdx.x - 32 * (2 + kE1Epi);            // 0..255
        const int g = gt & 7, pl = gt >> 3;                         // 16-byte cell (8 channels) of the 64-channel chunk, pixel lane (0..31)
        int hb = 0;
        uint32_t hph = 0;
        int it = 0;
        for (int wk = work0; wk < p.n_tiles; wk += gridDim.x, ++it) {
            const int tile = wk + crank;
            int img, y0, x0;
            tile_geom(tile, img, y0, x0);
            const bool real = tile < p.n_tiles;
            // x patch (zero outside the image): rows y0-2 .. y0+17, cols x0-2 .. x0+9, double-buffered across tiles
            float* xp = xpatch + (it & 1) * 4 * 240;
            const float* xs = p.x + (size_t)(real ? (p.row_sample ? p.row_sample[img] : img) : 0) * p.x_stride;
            for (int i = gt; i < C * 240; i += 32 * kE1Gen) {
                const int c = i / 240, r = i - c * 240, yy = y0 - 2 + r / 12, xx = x0 - 2 + r % 12;
                xp[i] = (real && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) ? __ldg(xs + ((size_t)c * p.H + yy) * p.W + xx) : 0.f;
            }
            asm volatile("bar.sync 9, 256;" ::: "memory");
            const int var = (real && p.row_variant) ? p.row_variant[img] : 0;
            const float* tb = p.tbias + (size_t)var * p.tb_var_stride;
            for (int c = 0; c < p.n_chunks; ++c) {
                const int ch = 64 * c + 8 * g;
                const float4 b3a = *reinterpret_cast<const float4*>(b3s + ch), b3b = *reinterpret_cast<const float4*>(b3s + ch + 4);
                const float4 t4a = __ldg(reinterpret_cast<const float4*>(tb + ch)), t4b = __ldg(reinterpret_cast<const float4*>(tb + ch + 4));
                float4 wa[9], wb[9];                                // C == 1: the nine taps of this thread's 8 channels stay in registers
                if (C == 1) {
#pragma unroll
                    for (int t9 = 0; t9 < 9; ++t9) {
                        wa[t9] = *reinterpret_cast<const float4*>(w3s + (size_t)t9 * coutp + ch);
                        wb[t9] = *reinterpret_cast<const float4*>(w3s + (size_t)t9 * coutp + ch + 4);
                    }
                }
                ptx::mbar_wait(hempty(hb), hph ^ 1u);               // the MMAs that read this buffer have retired
                uint8_t* hbuf = gbase + (halo0 - base) + (size_t)hb * kE1HaloBytes;
                auto conv1_at = [&](int px) -> uint4 {
                    const int ry = px / 10, rx = px - ry * 10;
                    const int yy = y0 - 1 + ry, xx = x0 - 1 + rx;
                    if (!(yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)) return make_uint4(0u, 0u, 0u, 0u);   // conv2's zero padding
                    float4 a0 = b3a, a1 = b3b;
                    if (C == 1) {
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const float v = xp[(ry + ky) * 12 + rx + kx];
                                const float4 u0 = wa[ky * 3 + kx], u1 = wb[ky * 3 + kx];
                                a0.x = fmaf(v, u0.x, a0.x); a0.y = fmaf(v, u0.y, a0.y); a0.z = fmaf(v, u0.z, a0.z); a0.w = fmaf(v, u0.w, a0.w);
                                a1.x = fmaf(v, u1.x, a1.x); a1.y = fmaf(v, u1.y, a1.y); a1.z = fmaf(v, u1.z, a1.z); a1.w = fmaf(v, u1.w, a1.w);
                            }
                    } else {
                        for (int ci = 0; ci < C; ++ci)
#pragma unroll
                            for (int t9 = 0; t9 < 9; ++t9) {
                                const float v = xp[ci * 240 + (ry + t9 / 3) * 12 + rx + t9 % 3];
                                const float4 u0 = *reinterpret_cast<const float4*>(w3s + (size_t)(t9 * C + ci) * coutp + ch);
                                const float4 u1 = *reinterpret_cast<const float4*>(w3s + (size_t)(t9 * C + ci) * coutp + ch + 4);
                                a0.x = fmaf(v, u0.x, a0.x); a0.y = fmaf(v, u0.y, a0.y); a0.z = fmaf(v, u0.z, a0.z); a0.w = fmaf(v, u0.w, a0.w);
                                a1.x = fmaf(v, u1.x, a1.x); a1.y = fmaf(v, u1.y, a1.y); a1.z = fmaf(v, u1.z, a1.z); a1.w = fmaf(v, u1.w, a1.w);
                            }
                    }
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
                    oh[0] = __floats2half2_rn(fmaxf(a0.x, 0.f) + t4a.x, fmaxf(a0.y, 0.f) + t4a.y);
                    oh[1] = __floats2half2_rn(fmaxf(a0.z, 0.f) + t4a.z, fmaxf(a0.w, 0.f) + t4a.w);
                    oh[2] = __floats2half2_rn(fmaxf(a1.x, 0.f) + t4b.x, fmaxf(a1.y, 0.f) + t4b.y);
                    oh[3] = __floats2half2_rn(fmaxf(a1.z, 0.f) + t4b.z, fmaxf(a1.w, 0.f) + t4b.w);
                    return o;
                };
                // 32 pixel lanes x 6 rounds cover the 180 halo pixels; two pixels (px, px + 96) per trip
                for (int px = pl; px < 96 && !(p.debug & 2); px += 32) {
                    const int px1 = px + 96;
                    uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
                    if (!(p.debug & 1)) {
                        o0 = conv1_at(px);
                        if (px1 < kE1HaloRows) o1 = conv1_at(px1);
                    }
                    *reinterpret_cast<uint4*>(hbuf + px * 128 + (((uint32_t)g ^ (uint32_t)(px & 7)) << 4)) = o0;
                    if (px1 < kE1HaloRows) *reinterpret_cast<uint4*>(hbuf + px1 * 128 + (((uint32_t)g ^ (uint32_t)(px1 & 7)) << 4)) = o1;
                }
                ptx::fence_proxy_async();                           // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    if constexpr (!kPair) ptx::mbar_arrive(hfull(hb));
                    else ptx::mbar_arrive_cluster(ptx::map_to_cta(hfull(hb), 0));
                }
                if (++hb == p.n_hbuf) { hb = 0; hph ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5): thread = output pixel
        const int q = warp & 3, ew = warp - 2;
        const int nchunk = coutp >> 5;                              // 32-column steps
        uint8_t* bufp = gbase + (ring0 - base) + (size_t)ew * 2048; // [32 rows][32 halfs], 64-byte swizzle
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);
        const int Wh = p.W >> 1;
        int acc = 0;
        uint32_t acc_ph = 0;
        float amax = 0.f;
        auto arrive_acc_empty = [&]() {
            if constexpr (!kPair) ptx::mbar_arrive(acc_empty0 + 8u * acc);
            else ptx::mbar_arrive_cluster(ptx::map_to_cta(acc_empty0 + 8u * acc, 0));
        };
        for (int wk = work0; wk < p.n_tiles; wk += gridDim.x) {
            const int tile = wk + crank;
            int img, y0, x0;
            tile_geom(tile, img, y0, x0);
            const bool real = tile < p.n_tiles;
            const int r = 32 * q + lane, yl = r >> 3, xl = r & 7;          // row of the 16 x 8 tile
            float xv[4] = {0.f, 0.f, 0.f, 0.f};
            if (real) {
                const float* xs = p.x + (size_t)(p.row_sample ? p.row_sample[img] : img) * p.x_stride + (size_t)(y0 + yl) * p.W + x0 + xl;
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) if (ci < C) xv[ci] = __ldg(xs + (size_t)ci * p.H * p.W);
            }
            ptx::mbar_wait(acc_full0 + 8u * acc, acc_ph);
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_cols);
            for (int c = 0; c < nchunk; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld32(t_acc + (uint32_t)(32 * c), raw);
                ptx::tmem_ld_wait();
                if (c == nchunk - 1) { ptx::tc_fence_before(); __syncwarp(); if (lane == 0) arrive_acc_empty(); }
                uint8_t* rowp = bufp + lane * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col = 32 * c + 8 * j;
                    float v[8];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias2s + col + 4 * hh);
                        float4 r4 = *reinterpret_cast<const float4*>(rb1s + col + 4 * hh);
#pragma unroll
                        for (int ci = 0; ci < 4; ++ci) {
                            if (ci >= C) break;
                            const float4 w4 = *reinterpret_cast<const float4*>(rw1s + (size_t)ci * coutp + col + 4 * hh);
                            r4.x = fmaf(xv[ci], w4.x, r4.x); r4.y = fmaf(xv[ci], w4.y, r4.y);
                            r4.z = fmaf(xv[ci], w4.z, r4.z); r4.w = fmaf(xv[ci], w4.w, r4.w);
                        }
                        v[4 * hh] = fmaxf(__uint_as_float(raw[8 * j + 4 * hh]) + b4.x, 0.f) + r4.x;
                        v[4 * hh + 1] = fmaxf(__uint_as_float(raw[8 * j + 4 * hh + 1]) + b4.y, 0.f) + r4.y;
                        v[4 * hh + 2] = fmaxf(__uint_as_float(raw[8 * j + 4 * hh + 2]) + b4.z, 0.f) + r4.z;
                        v[4 * hh + 3] = fmaxf(__uint_as_float(raw[8 * j + 4 * hh + 3]) + b4.w, 0.f) + r4.w;
                    }
                    uint4 pk;
                    __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        amax = fmaxf(amax, fmaxf(fabsf(v[2 * i]), fabsf(v[2 * i + 1])));
                        ph2[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                    }
                    *reinterpret_cast<uint4*>(rowp + (((uint32_t)j ^ swz) << 4)) = pk;
                }
                __syncwarp();
                if (real) {
                    // the warp's 4 x 8 pixel patch holds 2 x 4 complete 2x2 windows: lane -> (window, 16-byte cell)
                    const int pr = lane >> 2, wy = pr >> 2, wx = pr & 3;
                    const int r00 = (2 * wy) * 8 + 2 * wx;
                    const int py = (y0 + 4 * q + 2 * wy) >> 1, pxx = (x0 + 2 * wx) >> 1;
                    const uint32_t jj = (uint32_t)(lane & 3);
                    auto at = [&](int rr) { return *reinterpret_cast<const uint4*>(bufp + rr * 64 + ((jj ^ (((uint32_t)rr >> 1) & 3u)) << 4)); };
                    const uint4 a = at(r00), b = at(r00 + 1), cq = at(r00 + 8), d = at(r00 + 9);
                    const __half2* ah = reinterpret_cast<const __half2*>(&a);
                    const __half2* bh = reinterpret_cast<const __half2*>(&b);
                    const __half2* ch2 = reinterpret_cast<const __half2*>(&cq);
                    const __half2* dh = reinterpret_cast<const __half2*>(&d);
                    uint4 o4;
                    __half2* oh = reinterpret_cast<__half2*>(&o4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) oh[i] = __hmax2(__hmax2(ah[i], bh[i]), __hmax2(ch2[i], dh[i]));
                    *reinterpret_cast<uint4*>(p.pool_out + (((size_t)img * (p.H >> 1) + py) * Wh + pxx) * coutp + 32 * c + 8 * (int)jj) = o4;
                }
                __syncwarp();
            }
            if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
        if (!(amax <= 65504.f)) atomicOr(&g_umma_error, 2u);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (kPair) ptx::cluster_sync_all();
    if (warp == 0) {
        ptx::tc_fence_after();
        if constexpr (!kPair) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.acc_cols)) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.acc_cols)) : "memory");
    }
}

struct Enc1hLaunch {
    Enc1Maps maps;
    Enc1hParams p;
    unsigned grid;
    size_t smem;
    int pair;
    double flops;                  // conv2's tensor-core flops (real channels)
};

// `w2` = conv2 weights packed by pack_conv in DTRAJ_PREC_F16 ([tap][chunk][coutp][64] halfs), `w2_rows` its 128-byte rows
inline int build_enc1h_launch(Enc1hLaunch* E, int C, int H, int coutp, int cout_real, int64_t R, const float* w2, int64_t w2_rows) {
    memset(E, 0, sizeof(*E));
    if (C < 1 || C > 4 || H % 16 || H > 32 || coutp % 64 || coutp > 256) return fail(DTRAJ_EINVAL, "enc1(f16): unsupported geometry");
    Enc1hParams& p = E->p;
    p.C = C; p.H = H; p.W = H; p.coutp = coutp;
    p.n_chunks = coutp / 64;
    p.tiles_x = H / 8;
    p.tiles_per_img = (H / 16) * p.tiles_x;
    const int64_t nt = R * p.tiles_per_img;
    if (nt >= ((int64_t)1 << 30)) return fail(DTRAJ_EINVAL, "enc1(f16): batch too large");
    p.n_tiles = (int)nt;
    p.acc_cols = 32;
    while (p.acc_cols < coutp) p.acc_cols *= 2;
    auto fixed_for = [&](int pair) {
        return (size_t)1024 + (size_t)9 * p.n_chunks * (coutp / (pair ? 2 : 1)) * 128 + kE1Epi * 2048 + 2 * 4 * 240 * 4 +
               (size_t)(9 * C + 3 + C) * coutp * 4 + 16 + 256;
    };
    // pairs halve the resident weights per CTA; without them the weights must still fit next to one halo buffer
    E->pair = (p.n_tiles >= 2 * kNumSMs && !getenv("DTRAJ_NO_PAIR")) ? 1 : 0;
    if (!E->pair && fixed_for(0) + kE1HaloBytes > 227 * 1024) {
        if (p.n_tiles < 2) return fail(DTRAJ_EINVAL, "enc1(f16): weights do not fit without a CTA pair");
        E->pair = 1;
    }
    p.w_rows = coutp / (E->pair ? 2 : 1);
    const size_t fixed = fixed_for(E->pair);
    int nh = (int)((227 * 1024 - fixed) / kE1HaloBytes);
    if (nh < 1) return fail(DTRAJ_EINVAL, "enc1(f16): shared memory does not fit (coutp=%d pair=%d)", coutp, E->pair);
    if (nh > 2 * p.n_chunks) nh = 2 * p.n_chunks;
    if (nh > 8) nh = 8;
    p.n_hbuf = nh;
    E->smem = fixed + (size_t)nh * kE1HaloBytes;
    E->grid = (unsigned)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
    if (E->pair) E->grid = (E->grid + 1) / 2 * 2;
    DTRAJ_TRY(make_w_map(&E->maps.w, w2, w2_rows, p.w_rows, 1));
    E->flops = 2.0 * (double)R * H * H * cout_real * (double)cout_real * 9.0;
    p.debug = getenv("DTRAJ_E1_DEBUG") ? atoi(getenv("DTRAJ_E1_DEBUG")) : 0;
    return 0;
}

inline cudaError_t enc1h_set_smem_attr() {
    cudaError_t e = cudaFuncSetAttribute(k_enc1_f16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_enc1_f16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

inline int launch_enc1h(const Enc1hLaunch& E, cudaStream_t st) {
    if (!E.pair) {
        k_enc1_f16<false><<<E.grid, kE1Threads, E.smem, st>>>(E.maps, E.p);
        DTRAJ_LAUNCH_CHECK();
        return 0;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(E.grid);
    cfg.blockDim = dim3(kE1Threads);
    cfg.dynamicSmemBytes = E.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DTRAJ_CUDA(cudaLaunchKernelEx(&cfg, k_enc1_f16<true>, E.maps, E.p));
    return 0;
}

}  // namespace dtraj

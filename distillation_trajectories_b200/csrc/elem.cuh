// HBM/L2-bound kernels around the convolutions: time-embedding table, first conv (Cin<=4),
// 2x2 max-pool, bilinear x2 upsample (align_corners), final 1x1 conv, fused sampler step.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace dtraj {

// =====================================================================================
// Time-embedding bias table.
// Reference: models.py:15-39 (sinusoidal), :120-131,:175-185 (time / cond MLPs),
//            :66-67 (per block relu(time_mlp(temb))).
// Inside one sampler step every row shares t and cond takes <= 3 values, so these are
// functions of (t, variant) only: table[(t*3 + variant) * tb_stride + block_off + c].
// =====================================================================================
struct TimeTableParams {
    int temb;            // time_emb_dim
    int half;            // temb / 2
    float freq_scale;    // -(log(10000) / (half - 1 + 1e-8)) as fp32
    const float* w1; const float* b1;      // time_mlp.1    [temb,temb],[temb]
    const float* cw0; const float* cb0;    // cond_emb.0    [temb,1],[temb]
    const float* cw2; const float* cb2;    // cond_emb.2    [temb,temb],[temb]
    const float* bw[8]; const float* bb[8]; // block.time_mlp [cout,temb],[cout]
    int bcout[8];        // real cout per block
    int boff[8];         // offset of the block inside one table row
    int tb_stride;       // floats per (t, variant) row
    float* table;        // [T*3*tb_stride], zero-initialised (pad channels stay 0)
};

__global__ void __launch_bounds__(256) k_time_table(TimeTableParams p) {
    extern __shared__ float sm[];
    float* emb = sm;             // [temb]
    float* h = sm + p.temb;      // [temb]
    float* g = h + p.temb;       // [temb]
    const int t = blockIdx.x, variant = blockIdx.y;
    for (int j = threadIdx.x; j < p.temb; j += blockDim.x) {
        float v = 0.f;
        if (j < 2 * p.half) {
            int k = j < p.half ? j : j - p.half;
            float freq = expf(__fmul_rn((float)k, p.freq_scale));
            float arg = __fmul_rn((float)t, freq);
            v = j < p.half ? sinf(arg) : cosf(arg);
        }
        emb[j] = v;
        if (variant != DTRAJ_VAR_NONE) {
            float c = variant == DTRAJ_VAR_COND1 ? 1.f : 0.f;
            g[j] = fmaxf(fmaf(p.cw0[j], c, p.cb0[j]), 0.f);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.temb; i += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < p.temb; ++j) acc = fmaf(p.w1[(size_t)i * p.temb + j], emb[j], acc);
        float v = fmaxf(acc + p.b1[i], 0.f);
        if (variant != DTRAJ_VAR_NONE) {
            float a2 = 0.f;
            for (int j = 0; j < p.temb; ++j) a2 = fmaf(p.cw2[(size_t)i * p.temb + j], g[j], a2);
            v += a2 + p.cb2[i];
        }
        h[i] = v;
    }
    __syncthreads();
    float* row = p.table + (size_t)(t * 3 + variant) * p.tb_stride;
    for (int b = 0; b < 8; ++b) {
        for (int n = threadIdx.x; n < p.bcout[b]; n += blockDim.x) {
            float acc = 0.f;
            const float* w = p.bw[b] + (size_t)n * p.temb;
            for (int j = 0; j < p.temb; ++j) acc = fmaf(w[j], h[j], acc);
            row[p.boff[b] + n] = fmaxf(acc + p.bb[b][n], 0.f);
        }
    }
}

// =====================================================================================
// enc1.conv1 (3x3, Cin = C <= 4) + folded BN + ReLU + time bias, and enc1.residual_conv
// (1x1) in the same pass over x.  Reference: models.py:59-77 with in_ch = config.channels.
// K = 9*C is far too small for the tensor cores; this layer is output-bandwidth bound.
// One CTA per forward row (image); x is read straight from the trajectory frame.
// =====================================================================================
struct FirstConvParams {
    const float* x;          // frame base; sample s at x + s * x_stride, layout [C,H,W]
    int64_t x_stride;
    const int32_t* row_sample;   // [R] or null (row == sample)
    const int32_t* row_variant;  // [R] or null (variant NONE)
    int C, H, W, coutp;
    const float* w3;         // [9*C][coutp]  BN-folded, tap-major then cin
    const float* b3;         // [coutp]
    const float* w1;         // [C][coutp]    residual 1x1
    const float* b1;         // [coutp]
    const float* tbias;      // table row base for this t: + variant*tb_var_stride + block_off
    int tb_var_stride;
    float* h;                // [R,H,W,coutp]   relu(bn(conv1 x)) + tbias
    float* r;                // [R,H,W,coutp]   residual_conv(x); null when the consumer recomputes it (CONV_RESX)
    int64_t lo_off;          // ACT_SPLIT: h low plane offset
    int act_mode;
    int f16;                 // DTRAJ_PREC_F16: h is a __half map (act_mode must be ACT_PLAIN, r null)
};

// four consecutive channels of a feature map at element index `idx`: fp32, or rounded to fp16 (rn)
__device__ __forceinline__ void store_act4(float* base, size_t idx, float4 o, int f16) {
    if (f16) {
        __half2* d = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(base) + idx);
        d[0] = __floats2half2_rn(o.x, o.y);
        d[1] = __floats2half2_rn(o.z, o.w);
    } else {
        *reinterpret_cast<float4*>(base + idx) = o;
    }
}

__global__ void __launch_bounds__(256) k_conv_first(FirstConvParams p) {
    extern __shared__ float sm[];
    const int C = p.C, H = p.H, W = p.W, cp = p.coutp;
    const int PW = W + 2, PH = H + 2;
    float* img = sm;                       // [C][PH][PW] zero-padded
    float* w3 = img + ((C * PH * PW + 3) & ~3);   // [9C][cp], 16-byte aligned
    float* w1 = w3 + 9 * C * cp;           // [C][cp]
    const int row = blockIdx.x;
    const int sample = p.row_sample ? p.row_sample[row] : row;
    const int variant = p.row_variant ? p.row_variant[row] : 0;
    const float* x = p.x + (size_t)sample * p.x_stride;
    for (int i = threadIdx.x; i < C * PH * PW; i += blockDim.x) {
        int c = i / (PH * PW), rem = i % (PH * PW), yy = rem / PW - 1, xx = rem % PW - 1;
        img[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? x[(c * H + yy) * W + xx] : 0.f;
    }
    for (int i = threadIdx.x; i < 9 * C * cp; i += blockDim.x) w3[i] = p.w3[i];
    for (int i = threadIdx.x; i < C * cp; i += blockDim.x) w1[i] = p.w1[i];
    __syncthreads();
    const int groups = cp / 4;
    const float* tb = p.tbias + (size_t)variant * p.tb_var_stride;
    const size_t hbase = (size_t)row * H * W * cp;       // element offset of this row's map
    float* hout = p.h + hbase;                           // (fp32 maps only)
    float* rout = p.r + hbase;
    // 256 threads and cp/4 <= 64 channel groups: a thread's group is the same for every item it owns when
    // 256 % groups == 0 (all padded widths except 96/160/192/224-style ones), so its 9*C weight vectors,
    // bias and time bias live in registers and the pixel loop is 9*C broadcast loads + 36*C FMAs.
    if (256 % groups == 0 && C == 1) {
        const int g = threadIdx.x % groups;
        float4 w[9];
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) w[t9] = *reinterpret_cast<const float4*>(w3 + t9 * cp + g * 4);
        const float4 wr = *reinterpret_cast<const float4*>(w1 + g * 4);
        const float4 b3 = *reinterpret_cast<const float4*>(p.b3 + g * 4), b1 = *reinterpret_cast<const float4*>(p.b1 + g * 4);
        const float4 t4 = *reinterpret_cast<const float4*>(tb + g * 4);
        for (int pix = threadIdx.x / groups; pix < H * W; pix += 256 / groups) {
            const int y = pix / W, xq = pix % W;
            float4 acc = b3;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float v = img[(y + ky) * PW + xq + kx];
                    const float4 ww = w[ky * 3 + kx];
                    acc.x = fmaf(v, ww.x, acc.x); acc.y = fmaf(v, ww.y, acc.y);
                    acc.z = fmaf(v, ww.z, acc.z); acc.w = fmaf(v, ww.w, acc.w);
                }
            float4 o = make_float4(fmaxf(acc.x, 0.f) + t4.x, fmaxf(acc.y, 0.f) + t4.y,
                                   fmaxf(acc.z, 0.f) + t4.z, fmaxf(acc.w, 0.f) + t4.w);
            o = act_round4(o, p.act_mode);
            float* dst = hout + (size_t)pix * cp + g * 4;
            store_act4(p.h, hbase + (size_t)pix * cp + g * 4, o, p.f16);
            if (p.act_mode == ACT_SPLIT) *reinterpret_cast<float4*>(dst + p.lo_off) = act_lo4(o);
            if (p.r) {
                const float v = img[(y + 1) * PW + xq + 1];
                *reinterpret_cast<float4*>(rout + (size_t)pix * cp + g * 4) =
                    make_float4(fmaf(v, wr.x, b1.x), fmaf(v, wr.y, b1.y), fmaf(v, wr.z, b1.z), fmaf(v, wr.w, b1.w));
            }
        }
        return;
    }
    for (int item = threadIdx.x; item < H * W * groups; item += blockDim.x) {
        const int pix = item / groups, g = item % groups, y = pix / W, xq = pix % W;
        float4 acc = *reinterpret_cast<const float4*>(p.b3 + g * 4);
        float4 racc = *reinterpret_cast<const float4*>(p.b1 + g * 4);
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float v = img[(c * PH + y + ky) * PW + xq + kx];
                    float4 w = *reinterpret_cast<const float4*>(w3 + ((ky * 3 + kx) * C + c) * cp + g * 4);
                    acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y);
                    acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
                }
            float v = img[(c * PH + y + 1) * PW + xq + 1];
            float4 w = *reinterpret_cast<const float4*>(w1 + c * cp + g * 4);
            racc.x = fmaf(v, w.x, racc.x); racc.y = fmaf(v, w.y, racc.y);
            racc.z = fmaf(v, w.z, racc.z); racc.w = fmaf(v, w.w, racc.w);
        }
        float4 t4 = *reinterpret_cast<const float4*>(tb + g * 4);
        float4 o = make_float4(fmaxf(acc.x, 0.f) + t4.x, fmaxf(acc.y, 0.f) + t4.y,
                               fmaxf(acc.z, 0.f) + t4.z, fmaxf(acc.w, 0.f) + t4.w);
        o = act_round4(o, p.act_mode);
        float* dst = hout + (size_t)pix * cp + g * 4;
        store_act4(p.h, hbase + (size_t)pix * cp + g * 4, o, p.f16);
        if (p.act_mode == ACT_SPLIT) *reinterpret_cast<float4*>(dst + p.lo_off) = act_lo4(o);
        if (p.r) *reinterpret_cast<float4*>(rout + (size_t)pix * cp + g * 4) = racc;
    }
}

// =====================================================================================
// MaxPool2d(2) on NHWC (models.py:134,191-201).
// =====================================================================================
__global__ void __launch_bounds__(256) k_pool2(const float* __restrict__ in, float* __restrict__ out,
                                               int64_t n_out4, int Ho, int Wo, int cp4,
                                               int64_t lo_off, int act_mode) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out4) return;
    int c4 = (int)(i % cp4);
    int64_t pix = i / cp4;
    int xo = (int)(pix % Wo);
    int64_t t = pix / Wo;
    int yo = (int)(t % Ho);
    int64_t n = t / Ho;
    const int Wi = Wo * 2;
    const float4* src = reinterpret_cast<const float4*>(in) +
                        (((n * (Ho * 2) + yo * 2) * Wi) + xo * 2) * cp4 + c4;
    float4 a = src[0], b = src[cp4], c = src[(int64_t)Wi * cp4], d = src[(int64_t)Wi * cp4 + cp4];
    float4 o = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)),
                           fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w)));
    reinterpret_cast<float4*>(out)[i] = o;
    if (act_mode == ACT_SPLIT) reinterpret_cast<float4*>(out + lo_off)[i] = act_lo4(o);
}

// fp16 maps (DTRAJ_PREC_F16): one thread = 8 channels (16 bytes) of one output pixel
__global__ void __launch_bounds__(256) k_pool2_h(const __half* __restrict__ in, __half* __restrict__ out,
                                                 int64_t n_out8, int Ho, int Wo, int cp8) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out8) return;
    int c8 = (int)(i % cp8);
    int64_t pix = i / cp8;
    int xo = (int)(pix % Wo);
    int64_t t = pix / Wo;
    int yo = (int)(t % Ho);
    int64_t n = t / Ho;
    const int Wi = Wo * 2;
    const uint4* src = reinterpret_cast<const uint4*>(in) + (((n * (Ho * 2) + yo * 2) * Wi) + xo * 2) * cp8 + c8;
    const uint4 a = src[0], b = src[cp8], c = src[(int64_t)Wi * cp8], d = src[(int64_t)Wi * cp8 + cp8];
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
    const __half2* ch = reinterpret_cast<const __half2*>(&c);
    const __half2* dh = reinterpret_cast<const __half2*>(&d);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) oh[k] = __hmax2(__hmax2(ah[k], bh[k]), __hmax2(ch[k], dh[k]));
    reinterpret_cast<uint4*>(out)[i] = o;
}

// =====================================================================================
// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) on NHWC
// (models.py:135,205-221).  Source coordinate = dst * (in-1)/(out-1) computed in fp32 as
// ATen does (area_pixel_compute_scale with align_corners).
// =====================================================================================
__device__ __forceinline__ void up2_coord(int dst, int in, float scale, int& i0, int& i1, float& l1) {
    float s = scale * (float)dst;
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0;
}

__global__ void __launch_bounds__(256) k_upsample2(const float* __restrict__ in, float* __restrict__ out,
                                                   int64_t n_out4, int Hi, int Wi, int cp4,
                                                   int64_t lo_off, int act_mode) {
    // one thread = 4 channels of one output pixel; 32-bit index arithmetic (the host checks n_out4 < 2^31),
    // Ho and Wo are powers of two
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint32_t)n_out4) return;
    const int Ho = Hi * 2, Wo = Wi * 2;
    const uint32_t pix = i / (uint32_t)cp4, c4 = i - pix * (uint32_t)cp4;
    const int lw = 31 - __clz(Wo), lh = 31 - __clz(Ho);
    const int xo = (int)(pix & (uint32_t)(Wo - 1)), yo = (int)((pix >> lw) & (uint32_t)(Ho - 1));
    const uint32_t n = pix >> (lw + lh);
    const float sh = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
    const float sw = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    int y0, y1, x0, x1; float ly, lx;
    up2_coord(yo, Hi, sh, y0, y1, ly);
    up2_coord(xo, Wi, sw, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float4* base = reinterpret_cast<const float4*>(in) + (size_t)n * Hi * Wi * cp4 + c4;
    const float4 v00 = __ldg(base + (y0 * Wi + x0) * cp4), v01 = __ldg(base + (y0 * Wi + x1) * cp4);
    const float4 v10 = __ldg(base + (y1 * Wi + x0) * cp4), v11 = __ldg(base + (y1 * Wi + x1) * cp4);
    float4 o;
    o.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    o.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    o.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    o.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    o = act_round4(o, act_mode);
    reinterpret_cast<float4*>(out)[i] = o;
    if (act_mode == ACT_SPLIT) reinterpret_cast<float4*>(out + lo_off)[i] = act_lo4(o);
}

// fp16 maps (DTRAJ_PREC_F16): one thread = 8 channels of one output pixel; interpolation in fp32, one rounding
__global__ void __launch_bounds__(256) k_upsample2_h(const __half* __restrict__ in, __half* __restrict__ out,
                                                     int64_t n_out8, int Hi, int Wi, int cp8) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint32_t)n_out8) return;
    const int Ho = Hi * 2, Wo = Wi * 2;
    const uint32_t pix = i / (uint32_t)cp8, c8 = i - pix * (uint32_t)cp8;
    const int lw = 31 - __clz(Wo), lh = 31 - __clz(Ho);
    const int xo = (int)(pix & (uint32_t)(Wo - 1)), yo = (int)((pix >> lw) & (uint32_t)(Ho - 1));
    const uint32_t n = pix >> (lw + lh);
    const float sh = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
    const float sw = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    int y0, y1, x0, x1; float ly, lx;
    up2_coord(yo, Hi, sh, y0, y1, ly);
    up2_coord(xo, Wi, sw, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const uint4* base = reinterpret_cast<const uint4*>(in) + (size_t)n * Hi * Wi * cp8 + c8;
    const uint4 v00 = __ldg(base + (y0 * Wi + x0) * cp8), v01 = __ldg(base + (y0 * Wi + x1) * cp8);
    const uint4 v10 = __ldg(base + (y1 * Wi + x0) * cp8), v11 = __ldg(base + (y1 * Wi + x1) * cp8);
    const __half2* a = reinterpret_cast<const __half2*>(&v00);
    const __half2* b = reinterpret_cast<const __half2*>(&v01);
    const __half2* c = reinterpret_cast<const __half2*>(&v10);
    const __half2* d = reinterpret_cast<const __half2*>(&v11);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 fa = __half22float2(a[k]), fb = __half22float2(b[k]), fc = __half22float2(c[k]), fd = __half22float2(d[k]);
        oh[k] = __floats2half2_rn(hy * (hx * fa.x + lx * fb.x) + ly * (hx * fc.x + lx * fd.x),
                                  hy * (hx * fa.y + lx * fb.y) + ly * (hx * fc.y + lx * fd.y));
    }
    reinterpret_cast<uint4*>(out)[i] = o;
}

// Block form of the same upsample: one thread = 8 channels of a 2 x 2 OUTPUT block.  With scale (in-1)/(2*in-1) the even
// output row 2*yb interpolates source rows (yb-1, yb) and the odd row 2*yb+1 rows (yb, yb+1) (clamped) -- checked on the
// host against ATen's fp32 index formula for the actual size (up2_static_ok) -- so the block needs a 3 x 3 source patch:
// 9 loads and conversions for 4 outputs instead of 16, and the horizontal partial sums of the middle row are shared.
// The generic kernel above was issue-bound (226 instructions per output vector, 88 % issue-active in ncu).
struct Up2Coef { float ly_even[16], ly_odd[16]; };     // vertical (= horizontal) weights per block row, ATen's l1 = s - floor(s)

__global__ void __launch_bounds__(256) k_upsample2_h4(const __half* __restrict__ in, __half* __restrict__ out,
                                                      uint32_t n_blk, int Hi, int lg_hi, int cp8, int lg_cp8,
                                                      const __grid_constant__ Up2Coef cf) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;        // (n, yb, xb, c8), c8 fastest
    if (i >= n_blk) return;
    const int c8 = (int)(i & (uint32_t)(cp8 - 1));
    const uint32_t pix = i >> lg_cp8;
    const int xb = (int)(pix & (uint32_t)(Hi - 1)), yb = (int)((pix >> lg_hi) & (uint32_t)(Hi - 1));
    const uint32_t n = pix >> (2 * lg_hi);
    const int ym = yb > 0 ? yb - 1 : 0, yp = yb < Hi - 1 ? yb + 1 : Hi - 1;
    const int xm = xb > 0 ? xb - 1 : 0, xp = xb < Hi - 1 ? xb + 1 : Hi - 1;
    const uint4* base = reinterpret_cast<const uint4*>(in) + (size_t)n * Hi * Hi * cp8 + c8;
    const int rows[3] = {ym, yb, yp}, cols[3] = {xm, xb, xp};
    const float lxe = cf.ly_even[xb], lxo = cf.ly_odd[xb], hxe = 1.f - lxe, hxo = 1.f - lxo;
    const float lye = cf.ly_even[yb], lyo = cf.ly_odd[yb], hye = 1.f - lye, hyo = 1.f - lyo;
    // horizontal partial sums per source row: he = hx*v[xm] + lx*v[xb] (even output column), ho = hx*v[xb] + lx*v[xp] (odd)
    float he[3][8], ho[3][8];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const uint4 a = __ldg(base + (rows[r] * Hi + cols[0]) * cp8);
        const uint4 b = __ldg(base + (rows[r] * Hi + cols[1]) * cp8);
        const uint4 c = __ldg(base + (rows[r] * Hi + cols[2]) * cp8);
        const __half2* ah = reinterpret_cast<const __half2*>(&a);
        const __half2* bh = reinterpret_cast<const __half2*>(&b);
        const __half2* ch = reinterpret_cast<const __half2*>(&c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fa = __half22float2(ah[k]), fb = __half22float2(bh[k]), fc = __half22float2(ch[k]);
            he[r][2 * k] = hxe * fa.x + lxe * fb.x; he[r][2 * k + 1] = hxe * fa.y + lxe * fb.y;
            ho[r][2 * k] = hxo * fb.x + lxo * fc.x; ho[r][2 * k + 1] = hxo * fb.y + lxo * fc.y;
        }
    }
    const int Wo = 2 * Hi;
    uint4* obase = reinterpret_cast<uint4*>(out) + ((size_t)n * Wo * Wo + (size_t)(2 * yb) * Wo + 2 * xb) * cp8 + c8;
    auto emit = [&](const float (&top)[8], const float (&bot)[8], float hy, float ly, uint4* dst) {
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(hy * top[2 * k] + ly * bot[2 * k], hy * top[2 * k + 1] + ly * bot[2 * k + 1]);
        *dst = o;
    };
    emit(he[0], he[1], hye, lye, obase);                                  // (2yb,   2xb)
    emit(ho[0], ho[1], hye, lye, obase + cp8);                            // (2yb,   2xb+1)
    emit(he[1], he[2], hyo, lyo, obase + (size_t)Wo * cp8);               // (2yb+1, 2xb)
    emit(ho[1], ho[2], hyo, lyo, obase + (size_t)Wo * cp8 + cp8);         // (2yb+1, 2xb+1)
}

// host: does ATen's fp32 coordinate formula pick (yb-1, yb) / (yb, yb+1) for every output row of this size?  Fills the weights.
inline bool up2_static_ok(int Hi, Up2Coef* cf) {
    if (Hi < 1 || Hi > 16 || (Hi & (Hi - 1))) return false;
    const int Ho = 2 * Hi;
    const float sc = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
    for (int yb = 0; yb < Hi; ++yb)
        for (int par = 0; par < 2; ++par) {
            const float s = sc * (float)(2 * yb + par);
            int i0 = (int)s;
            if (i0 > Hi - 1) i0 = Hi - 1;
            const int i1 = i0 + (i0 < Hi - 1 ? 1 : 0);
            const float l1 = s - (float)i0;
            const int top = par == 0 ? (yb > 0 ? yb - 1 : 0) : yb;
            const int bot = par == 0 ? yb : (yb < Hi - 1 ? yb + 1 : Hi - 1);
            // the static pair must give the same value: identical rows, or all the weight on a shared first row
            if (!((i0 == top && i1 == bot) || (l1 == 0.f && i0 == top))) return false;
            (par == 0 ? cf->ly_even : cf->ly_odd)[yb] = l1;
        }
    return true;
}

// =====================================================================================
// final 1x1 conv (models.py:157,224) evaluated at HALF resolution: a 1x1 conv commutes
// with the per-channel bilinear resize (lerp weights sum to 1), so
// final(upsample(y)) == upsample(final(y)) and the 128-channel full-resolution tensor of
// models.py:221 is never materialised (SURVEY.md note C).  One warp per pixel.
// =====================================================================================
__global__ void __launch_bounds__(256) k_final1x1(const float* __restrict__ y, const float* __restrict__ w,
                                                  const float* __restrict__ bias, float* __restrict__ elow,
                                                  int64_t n_pix, int cp, int C) {
    int64_t pix = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (pix >= n_pix) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* src = y + pix * cp;
    for (int c = lane * 4; c < cp; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(src + c);
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (o < C) {
                float4 ww = *reinterpret_cast<const float4*>(w + o * cp + c);
                acc[o] = fmaf(v.x, ww.x, fmaf(v.y, ww.y, fmaf(v.z, ww.z, fmaf(v.w, ww.w, acc[o]))));
            }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
        if (o < C) {
            float s = warp_sum(acc[o]);
            if (lane == 0) elow[pix * C + o] = s + bias[o];
        }
}

// bilinear x2 (align_corners) sample of the half-resolution eps map [h,w,C] at (c, yo, xo)
__device__ __forceinline__ float eps_up2(const float* __restrict__ e, int C, int c, int hi, int wi,
                                         int y0, int y1, float ly, int x0, int x1, float lx) {
    float v00 = e[(y0 * wi + x0) * C + c], v01 = e[(y0 * wi + x1) * C + c];
    float v10 = e[(y1 * wi + x0) * C + c], v11 = e[(y1 * wi + x1) * C + c];
    float hy = 1.f - ly, hx = 1.f - lx;
    return hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

// DiffusionUNet.forward output for the stand-alone forward API: eps [R,C,H,W] from e_low
__global__ void __launch_bounds__(256) k_eps_out(const float* __restrict__ elow, float* __restrict__ eps,
                                                 int64_t n, int C, int H, int W) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int xo = (int)(i % W);
    int64_t t = i / W;
    int yo = (int)(t % H); t /= H;
    int c = (int)(t % C);
    int64_t row = t / C;
    const int hi = H / 2, wi = W / 2;
    const float sh = (float)(hi - 1) / (float)(H - 1), sw = (float)(wi - 1) / (float)(W - 1);
    int y0, y1, x0, x1; float ly, lx;
    up2_coord(yo, hi, sh, y0, y1, ly);
    up2_coord(xo, wi, sw, x0, x1, lx);
    eps[i] = eps_up2(elow + row * hi * wi * C, C, c, hi, wi, y0, y1, ly, x0, x1, lx);
}

// =====================================================================================
// Fused sampler step: final upsample of eps, CFG combine, update rule, trajectory store.
//   CFG   utils/diffusion.py:126, analysis/trajectory_engine.py:80
//   S1    utils/diffusion.py:149-158       x' = k0*(x - k1*eps) + z*k2
//   S2    trajectory_engine.py:98-110      x' = (k0*x - k1*eps) + k2*z
//   S3    trajectory_manager.py:194-203    x' = (x - k0*eps)/k1 + k2*z
// Explicit *_rn intrinsics keep the reference's separately rounded mul/add (no FMA
// contraction).  Algorithmic traffic: read eps_u, eps_c, x, z, write x' = 20 B/element.
// =====================================================================================
__device__ __forceinline__ float step_rule(int rule, float k0, float k1, float k2, float x, float eps, float z) {
    if (rule == DTRAJ_RULE_S1)
        return __fadd_rn(__fmul_rn(k0, __fsub_rn(x, __fmul_rn(k1, eps))), __fmul_rn(z, k2));
    if (rule == DTRAJ_RULE_S2)
        return __fadd_rn(__fsub_rn(__fmul_rn(k0, x), __fmul_rn(k1, eps)), __fmul_rn(k2, z));
    return __fadd_rn(__fdiv_rn(__fsub_rn(x, __fmul_rn(k0, eps)), k1), __fmul_rn(k2, z));
}
__device__ __forceinline__ float cfg_mix(float eu, float ec, float w) {
    return __fadd_rn(eu, __fmul_rn(w, __fsub_rn(ec, eu)));
}

struct StepParams {
    int rule; float k0, k1, k2;
    const float* elow;           // [R, H/2, W/2, C]
    const int32_t* sample_row_u; // [B]
    const int32_t* sample_row_c; // [B] (-1: none) or null
    const float* guidance;       // [B] or null
    const float* z_bank;         // [*, D]
    const int32_t* z_index;      // [B] for this step (-1: none) or null
    const float* x_in;           // frame f   : x_in  + b*frame_stride
    float* x_out;                // frame f+1 : x_out + b*frame_stride
    int64_t frame_stride;        // L*D
    int B, C, H, W;
};

// one thread = 4 consecutive pixels along W of one (sample, channel, y)
__global__ void __launch_bounds__(256) k_step(StepParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int W4 = p.W / 4;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)p.B * p.C * p.H * W4;
    if (i >= total) return;
    int xq = (int)(i % W4);
    int64_t t = i / W4;
    int yo = (int)(t % p.H); t /= p.H;
    int c = (int)(t % p.C);
    int b = (int)(t / p.C);
    const int hi = p.H / 2, wi = p.W / 2, D = p.C * p.H * p.W;
    const float sh = (float)(hi - 1) / (float)(p.H - 1), sw = (float)(wi - 1) / (float)(p.W - 1);
    int y0, y1; float ly;
    up2_coord(yo, hi, sh, y0, y1, ly);
    const int ru = p.sample_row_u[b];
    const int rc = p.sample_row_c ? p.sample_row_c[b] : -1;
    const float w = (rc >= 0) ? p.guidance[b] : 0.f;
    const float* eu = p.elow + (size_t)ru * hi * wi * p.C;
    const float* ec = p.elow + (size_t)(rc >= 0 ? rc : ru) * hi * wi * p.C;
    const int64_t off = (int64_t)(c * p.H + yo) * p.W + xq * 4;
    float4 x = *reinterpret_cast<const float4*>(p.x_in + (int64_t)b * p.frame_stride + off);
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int zi = p.z_index ? p.z_index[b] : -1;
    if (zi >= 0) z = ld_stream4(p.z_bank + (int64_t)zi * D + off);
    float xv[4] = {x.x, x.y, x.z, x.w}, zv[4] = {z.x, z.y, z.z, z.w}, o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int x0, x1; float lx;
        up2_coord(xq * 4 + j, wi, sw, x0, x1, lx);
        float e = eps_up2(eu, p.C, c, hi, wi, y0, y1, ly, x0, x1, lx);
        if (rc >= 0) e = cfg_mix(e, eps_up2(ec, p.C, c, hi, wi, y0, y1, ly, x0, x1, lx), w);
        o[j] = step_rule(p.rule, p.k0, p.k1, p.k2, xv[j], e, zv[j]);
    }
    *reinterpret_cast<float4*>(p.x_out + (int64_t)b * p.frame_stride + off) = make_float4(o[0], o[1], o[2], o[3]);
}

// stand-alone form on full-resolution eps (dtraj_step_fused)
__global__ void __launch_bounds__(256) k_step_plain(int rule, float k0, float k1, float k2,
                                                    const float* __restrict__ eu, const float* __restrict__ ec,
                                                    const float* __restrict__ w,
                                                    const float* __restrict__ x, int64_t xs,
                                                    const float* __restrict__ z, int64_t zs,
                                                    float* __restrict__ out, int64_t os, int64_t B, int64_t D) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    int64_t b = i / D, e = i % D;
    float eps = eu[i];
    if (ec) eps = cfg_mix(eps, ec[i], w[b]);
    out[b * os + e] = step_rule(rule, k0, k1, k2, x[b * xs + e], eps, z ? z[b * zs + e] : 0.f);
}

// duplicate frame (S2 at t == 0 records x unchanged, trajectory_engine.py:86,113)
__global__ void __launch_bounds__(256) k_copy_frame(const float* __restrict__ in, float* __restrict__ out,
                                                    int64_t frame_stride, int B, int D4) {
    pdl_launch_dependents();
    pdl_wait();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * D4) return;
    int64_t b = i / D4, e = i % D4;
    reinterpret_cast<float4*>(out + b * frame_stride)[e] = reinterpret_cast<const float4*>(in + b * frame_stride)[e];
}

}  // namespace dtraj

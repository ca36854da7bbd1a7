// fp32 CUDA-core implicit-GEMM convolution (3x3 pad 1, or 1x1) on NHWC feature maps.
// This is the exact-arithmetic mode (DTRAJ_PREC_FP32): the reference's conv layers
// (models.py:48-57) with folded BatchNorm, ReLU, time-bias add and residual add fused in
// the epilogue (models.py:59-83).  It shares the packed weight layout and the layer
// descriptor with the tcgen05 kernel in conv_umma.cuh.
#pragma once
#include "common.cuh"

namespace dtraj {

enum ConvFlags : int {
    CONV_RELU = 1, CONV_TBIAS = 2, CONV_RESID = 4,
    // fused tails, tcgen05 kernel only (conv_umma.cuh)
    CONV_POOL = 8, CONV_NOSTORE = 16, CONV_RESX = 32, CONV_FINAL = 64,
    CONV_RESACC = 128   // the block's 1x1 residual conv runs as extra MMAs of THIS kernel into a second accumulator
};

// One convolution layer over up to two NHWC sources (implicit channel concat,
// models.py:206,211,216).  M = n_img*H*W output pixels, N = coutp, K = ntaps * (c0p + c1p).
struct ConvLayer {
    const float* src0; const float* src1;   // [n_img,H,W,c0p] / [n_img,H,W,c1p] (src1 may be null)
    const float* src0_lo; const float* src1_lo;   // low planes of the sources (3xTF32 only)
    int c0p, c1p;
    int H, W;
    int64_t M;
    int ntaps;                 // 9 (3x3) or 1 (1x1, or 3x3 at 1x1 spatial where only the centre tap sees data)
    const float* wpk;          // [ntaps * (c0p+c1p)/32][coutp][32]  K-major blocks
    const float* bias;         // [coutp]
    int coutp;
    const float* tbias;        // + variant*tb_var_stride + c  (already offset to t and block)
    int tb_var_stride;
    int tb_rows;               // 1: row_variant indexes the WHOLE [T][3] table (a timestep per row, dtraj_unet_forward_rows)
    const int32_t* row_variant;  // per image; null = variant 0
    const float* resid;        // [M, coutp] or null
    float* out;                // [M, coutp]
    int64_t lo_off;            // ACT_SPLIT low plane offset (floats)
    int act_mode;
    int flags;
    // fused tails (tcgen05 kernel only)
    float* pool_out;           // CONV_POOL: [M/4, coutp] max-pooled output
    const float* xraw;         // CONV_RESX: raw input frame base; sample s at xraw + s*x_stride, [xC,H,W]
    int64_t x_stride;
    const int32_t* row_sample; //            per image: sample index (null = identity)
    const float* rw1;          //            residual 1x1 weights [xC][coutp]
    const float* rb1;          //            residual 1x1 bias [coutp]
    int xC;
    int finC;                  // CONV_FINAL: output channels of the final 1x1
    const float* finw;         //            [finC][coutp]
    const float* finb;         //            [finC]
    float* elow;               //            [M, finC]
    const float* rsrc0;        // CONV_RESACC: the block's input maps (sources of residual_conv, models.py:54-60)
    const float* rsrc1;
    int rc0p, rc1p;
    const float* rbias;        //            residual_conv bias [coutp]
    int f16;                   // DTRAJ_PREC_F16 (tcgen05 kernel only): src*/resid/out/pool_out/rsrc* point at __half maps,
                               // channel counts are padded to 64
    unsigned int* err;         // the owning handle's device error word (null: the library-wide word), tcgen05 kernels only
};

template <int BN>
__global__ void __launch_bounds__(256) k_conv_simt(ConvLayer p) {
    constexpr int BM = 128, BK = 32, TN = BN / 16;     // thread tile 8 x TN
    constexpr int AP = BM + 4, BP = BN + 4;
    __shared__ __align__(16) float As[BK][AP];
    __shared__ __align__(16) float Bs[BK][BP];

    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int HW = p.H * p.W;

    // A loader: thread -> (row, 16-channel half)
    const int lrow = tid & 127, lhalf = tid >> 7;
    const int64_t lm = m0 + lrow;
    const bool lvalid = lm < p.M;
    int64_t limg = 0; int ly = 0, lx = 0;
    if (lvalid) { limg = lm / HW; int r = (int)(lm % HW); ly = r / p.W; lx = r % p.W; }
    // B loader: thread -> (cout n, 8-wide k quarter)
    const int bn = tid % BN, bq = tid / BN;     // BN=64: 4 quarters of 8; BN=32: 8 parts of 4
    constexpr int BKQ = BK / (256 / BN);        // k values per thread

    const int nch0 = p.c0p / 32, nch = nch0 + p.c1p / 32;
    const int nkb = p.ntaps * nch;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int tm = tid / 16, tn = tid % 16;
    float4 areg[4];
    float breg[BKQ];

    auto load_kb = [&](int kb) {
        const int tap = kb / nch, chunk = kb % nch;
        int dy = 0, dx = 0;
        if (p.ntaps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
        const float* src; int cp, c0;
        if (chunk < nch0) { src = p.src0; cp = p.c0p; c0 = chunk * 32; }
        else { src = p.src1; cp = p.c1p; c0 = (chunk - nch0) * 32; }
        const int yy = ly + dy, xx = lx + dx;
        if (lvalid && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
            const float4* g = reinterpret_cast<const float4*>(
                src + ((limg * p.H + yy) * p.W + xx) * cp + c0 + lhalf * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) areg[j] = g[j];
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) areg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float* wb = p.wpk + ((size_t)kb * p.coutp + n0 + bn) * 32 + bq * BKQ;
#pragma unroll
        for (int j = 0; j < BKQ; j += 4) {
            float4 v = *reinterpret_cast<const float4*>(wb + j);
            breg[j] = v.x; breg[j + 1] = v.y; breg[j + 2] = v.z; breg[j + 3] = v.w;
        }
    };
    auto store_kb = [&]() {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = lhalf * 16 + j * 4;
            As[c][lrow] = areg[j].x; As[c + 1][lrow] = areg[j].y;
            As[c + 2][lrow] = areg[j].z; As[c + 3][lrow] = areg[j].w;
        }
#pragma unroll
        for (int j = 0; j < BKQ; ++j) Bs[bq * BKQ + j][bn] = breg[j];
    };

    load_kb(0);
    for (int kb = 0; kb < nkb; ++kb) {
        __syncthreads();
        store_kb();
        __syncthreads();
        if (kb + 1 < nkb) load_kb(kb + 1);
#pragma unroll 8
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[TN];
            if constexpr (TN == 4) {
                float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
                b[0] = b4.x; b[1] = b4.y; b[2] = b4.z; b[3] = b4.w;
            } else {
                float2 b2 = *reinterpret_cast<const float2*>(&Bs[k][tn * 2]);
                b[0] = b2.x; b[1] = b2.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }

    // epilogue: bias, relu, + time bias, + residual, precision-mode store
    const int nc = n0 + tn * TN;
    float bias[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bias[j] = p.bias[nc + j];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + tm * 8 + i;
        if (m >= p.M) continue;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            v[j] = acc[i][j] + bias[j];
            if (p.flags & CONV_RELU) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.flags & CONV_TBIAS) {
            const int var = p.row_variant ? p.row_variant[m / HW] : 0;
            const float* tb = p.tbias + (size_t)var * p.tb_var_stride + nc;
#pragma unroll
            for (int j = 0; j < TN; ++j) v[j] += tb[j];
        }
        if (p.flags & CONV_RESID) {
            const float* r = p.resid + m * p.coutp + nc;
#pragma unroll
            for (int j = 0; j < TN; ++j) v[j] += r[j];
        }
        float* dst = p.out + m * p.coutp + nc;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float s = act_store_value(v[j], p.act_mode);
            dst[j] = s;
            if (p.act_mode == ACT_SPLIT) dst[p.lo_off + j] = s - tf32_trunc(s);
        }
    }
}

inline int launch_conv_simt(const ConvLayer& L, cudaStream_t st) {
    const unsigned gm = (unsigned)((L.M + 127) / 128);
    if (L.coutp % 64 == 0) {
        k_conv_simt<64><<<dim3(gm, L.coutp / 64), 256, 0, st>>>(L);
    } else {
        k_conv_simt<32><<<dim3(gm, L.coutp / 32), 256, 0, st>>>(L);
    }
    DTRAJ_LAUNCH_CHECK();
    return 0;
}

}  // namespace dtraj

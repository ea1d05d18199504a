// HBM-bound kernels of the UAVSal path: layout conversion, stem conv, depthwise 3x3, bilinear (align_corners),
// temporal differences, context-prior sum, readout dot+sigmoid, uint8 post-processing.
// All activations are NHWC split-bf16 planes (common.cuh).  sm_100a only.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <cstring>
#include "common.cuh"

namespace uavsal {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------------------------------
// layout conversion
// ---------------------------------------------------------------------------------------------------
__global__ void pack_nchw_kernel(const float* __restrict__ src, int n, int c, int hw, ActW dst, int cpad) {
    pdl_trigger();
    pdl_wait();
    const int groups = cpad >> 3;
    const int64_t total = (int64_t)n * hw * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % hw);
        const int64_t r = i / hw;
        const int g = (int)(r % groups);
        const int img = (int)(r / groups);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = g * 8 + j;
            v[j] = ch < c ? __ldg(src + ((int64_t)img * c + ch) * hw + p) : 0.f;
        }
        store8(dst.p + ((int64_t)img * hw + p) * dst.ld + g * 8, dst.plane, v);
    }
}

__global__ void unpack_nchw_kernel(Act src, int n, int c, int hw, float* __restrict__ dst) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = (int64_t)n * c * hw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % hw);
        const int64_t r = i / hw;
        const int ch = (int)(r % c);
        const int img = (int)(r / c);
        dst[i] = load1(src.p + ((int64_t)img * hw + p) * src.ld + ch, src.plane);
    }
}

// ---------------------------------------------------------------------------------------------------
// stem: (normalise) + conv3x3 s2 p1 (3->32) + BN + ReLU6      [utils_data.py:43-65, torchvision features[0]]
// one thread per output pixel, 32 accumulators, weights [27][32] in shared memory
// ---------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ float stem_fetch(const void* x, int img, int ch, int y, int xx, int h, int w) {
    if (KIND == 0) {
        return __ldg(reinterpret_cast<const float*>(x) + (((int64_t)img * 3 + ch) * h + y) * w + xx);
    } else {
        uint8_t u;
        if (KIND == 1) u = __ldg(reinterpret_cast<const uint8_t*>(x) + (((int64_t)img * 3 + ch) * h + y) * w + xx);
        else           u = __ldg(reinterpret_cast<const uint8_t*>(x) + (((int64_t)img * h + y) * w + xx) * 3 + ch);
        const float mean = ch == 0 ? 0.485f : (ch == 1 ? 0.456f : 0.406f);
        const float sd = ch == 0 ? 0.229f : (ch == 1 ? 0.224f : 0.225f);
        const float f = __fdiv_rn((float)u, 255.0f);
        return __fdiv_rn(__fsub_rn(f, mean), sd);
    }
}

// PW = true: weights [27][32] + bias [32] travel in the kernel parameters (constant bank): the FMAs take their weights as
// warp-uniform operands (pairs through uniform registers, packed fma.rn.f32x2) instead of 216 shared-memory loads per pixel
// (uavsal_stem_conv3x3s2_hw: host weight pointers)
struct StemW {
    float w[27 * 32 + 32];
};

template <int KIND, bool PW>
__global__ void __launch_bounds__(128) stem_kernel(const void* __restrict__ x, int n, int h, int w, int ho, int wo,
                                                   const float* __restrict__ wgt, const float* __restrict__ bias,
                                                   ActW out, const __grid_constant__ StemW cw) {
    pdl_trigger();
    pdl_wait();
    __shared__ float4 sw[PW ? 1 : 27 * 8];
    __shared__ float sb[32];
    __shared__ float lut[3][256];      // uint8 -> normalised fp32 with the reference's exact expression (utils_data.py:56-60)
    if (!PW) {
        for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) sw[i] = reinterpret_cast<const float4*>(wgt)[i];
        if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
    }
    if (KIND != 0) {
        for (int i = threadIdx.x; i < 768; i += blockDim.x) {
            const int ch = i >> 8, u = i & 255;
            const float mean = ch == 0 ? 0.485f : (ch == 1 ? 0.456f : 0.406f);
            const float sd = ch == 0 ? 0.229f : (ch == 1 ? 0.224f : 0.225f);
            lut[ch][u] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), mean), sd);
        }
    }
    __syncthreads();
    __shared__ float st[128 * 33];                                         // output tile, row pitch 33 floats (conflict-free)
    // persistent blocks over tiles of 128 consecutive output pixels: the weight / LUT set-up above (two IEEE divisions per
    // LUT entry) is paid once per block, not once per 128 pixels; 32-bit index arithmetic
    const unsigned total = (unsigned)n * ho * wo;                          // host-checked < 2^31
    for (unsigned i0 = blockIdx.x * 128u; i0 < total; i0 += gridDim.x * 128u) {
        const unsigned i = i0 + threadIdx.x;
        if (i < total) {
            const unsigned row = i / (unsigned)wo;
            const int ox = (int)(i - row * (unsigned)wo);
            const int img = (int)(row / (unsigned)ho);
            const int oy = (int)(row - (unsigned)img * (unsigned)ho);
            float acc[32];
            float2 acc2[PW ? 16 : 1];                                          // PW: channel pairs (packed fma.rn.f32x2, weights through uniform registers)
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = PW ? 0.f : sb[j];
            if (PW) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc2[j] = make_float2(cw.w[27 * 32 + 2 * j], cw.w[27 * 32 + 2 * j + 1]);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int y = oy * 2 - 1 + ky;
                if (y < 0 || y >= h) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = ox * 2 - 1 + kx;
                    if (xx < 0 || xx >= w) continue;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        float v;
                        if (KIND == 0) v = __ldg(reinterpret_cast<const float*>(x) + (((int64_t)img * 3 + ch) * h + y) * w + xx);
                        else if (KIND == 1) v = lut[ch][__ldg(reinterpret_cast<const uint8_t*>(x) + (((int64_t)img * 3 + ch) * h + y) * w + xx)];
                        else v = lut[ch][__ldg(reinterpret_cast<const uint8_t*>(x) + (((int64_t)img * h + y) * w + xx) * 3 + ch)];
                        if (PW) {
                            const float* wk = cw.w + ((ky * 3 + kx) * 3 + ch) * 32;
#pragma unroll
                            for (int j = 0; j < 16; ++j) acc2[j] = __ffma2_rn(make_float2(v, v), make_float2(wk[2 * j], wk[2 * j + 1]), acc2[j]);
                            continue;
                        }
                        const float4* wr = sw + ((ky * 3 + kx) * 3 + ch) * 8;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 ww = wr[q];
                            acc[q * 4 + 0] = fmaf(v, ww.x, acc[q * 4 + 0]);
                            acc[q * 4 + 1] = fmaf(v, ww.y, acc[q * 4 + 1]);
                            acc[q * 4 + 2] = fmaf(v, ww.z, acc[q * 4 + 2]);
                            acc[q * 4 + 3] = fmaf(v, ww.w, acc[q * 4 + 3]);
                        }
                    }
                }
            }
            if (PW) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) st[threadIdx.x * 33 + j] = relu6f(acc[j]);
        }
        __syncthreads();
        // cooperative copy-out: consecutive threads write consecutive 16-byte pieces of the tile's pixels (the per-thread
        // version wrote one 128-byte row per lane, 32 lines per store instruction)
        if (out.plane == UAVSAL_PLANE_F32) {                               // fp32 rows (input of features.1's depthwise conv)
            float* of = reinterpret_cast<float*>(out.p);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = k * 128 + threadIdx.x, px = idx >> 3, q = idx & 7;
                if (i0 + px < total) {
                    const float* sp = st + px * 33 + q * 4;
                    *reinterpret_cast<float4*>(of + (int64_t)(i0 + px) * out.ld + q * 4) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = k * 128 + threadIdx.x, px = idx >> 2, q = idx & 3;
                if (i0 + px < total) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = st[px * 33 + q * 8 + j];
                    store8(out.p + (int64_t)(i0 + px) * out.ld + q * 8, out.plane, v);
                }
            }
        }
        __syncthreads();                                                   // st is rewritten by the next tile
    }
}

// ---------------------------------------------------------------------------------------------------
// depthwise 3x3 + BN + ReLU6, stride 1|2, dilation d, pad d     [model.py:92, torchvision InvertedResidual]
// one thread per (output pixel, 8-channel group); taps falling in the padding are skipped
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dw3x3_kernel(Act in, int n, int h, int w, int c, int ho, int wo, int stride,
                                                    int dil, const float* __restrict__ wgt,
                                                    const float* __restrict__ bias, int relu6, ActW out) {
    pdl_trigger();
    pdl_wait();
    const int groups = c >> 3;
    const int64_t total = (int64_t)n * ho * wo * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;
    const int ox = (int)(pix % wo);
    const int oy = (int)((pix / wo) % ho);
    const int img = (int)(pix / ((int64_t)wo * ho));
    const int c0 = g * 8;
    float acc[8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
        acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
        acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int y = oy * stride + (ky - 1) * dil;
        if (y < 0 || y >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int xx = ox * stride + (kx - 1) * dil;
            if (xx < 0 || xx >= w) continue;
            float v[8];
            load8(in.p + (((int64_t)img * h + y) * w + xx) * in.ld + c0, in.plane, v);
            const float* wr = wgt + (ky * 3 + kx) * c + c0;
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr + 4));
            acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]);
            acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
            acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]);
            acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
        }
    }
    if (relu6) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = relu6f(acc[j]);
    }
    store8(out.p + pix * out.ld + c0, out.plane, acc);
}

// ---------------------------------------------------------------------------------------------------
// depthwise 3x3 fast path (dilation 1, stride 1|2): sliding 3x3 register window down a strip of RB output rows.
// thread = (8-channel group, output column); a block covers 64 channels x 32 columns x RB rows, so every warp-level
// load is four full 128-byte lines per plane and each input row is fetched once per strip (3 loads per output instead
// of 9); the two neighbouring columns come from L1.  BN-folded weights of the block's 64 channels sit in shared memory.
// ---------------------------------------------------------------------------------------------------
constexpr int kDwRB = 9;

template <int STRIDE>
__global__ void __launch_bounds__(256, 2) dw3x3_rows_kernel(Act in, int h, int w, int c, int ho, int wo,
                                                         const float* __restrict__ wgt, const float* __restrict__ bias,
                                                         int relu6, ActW out, int strips) {
    __shared__ float4 sw[9][8][2];      // [tap][group in block][half]
    __shared__ float4 sb[8][2];
    const int gl = threadIdx.x & 7;
    const int cbase = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 9 * 16; i += 256) {
        const int tap = i / 16, q = i % 16;       // q: float4 index within the 64 channels
        const int ch = cbase + q * 4;
        sw[tap][q >> 1][q & 1] = ch < c ? __ldg(reinterpret_cast<const float4*>(wgt + tap * c + ch)) : make_float4(0, 0, 0, 0);
    }
    if (threadIdx.x < 16) {
        const int ch = cbase + threadIdx.x * 4;
        sb[threadIdx.x >> 1][threadIdx.x & 1] = ch < c ? __ldg(reinterpret_cast<const float4*>(bias + ch)) : make_float4(0, 0, 0, 0);
    }
    __syncthreads();
    const int c0 = cbase + gl * 8;
    const int ox = blockIdx.y * 32 + (threadIdx.x >> 3);
    const int strip = blockIdx.z % strips, img = blockIdx.z / strips;
    if (c0 >= c || ox >= wo) return;
    const int oy0 = strip * kDwRB;
    const int xc = ox * STRIDE;                   // centre input column
    const bool xl = xc - 1 >= 0, xr = xc + 1 < w;
    const uint16_t* ibase = in.p + (int64_t)img * h * w * in.ld + c0;

    float win[3][3][8];                           // [row slot][dx][channel]
    auto load_row = [&](int slot, int y) {
        if (y < 0 || y >= h) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int j = 0; j < 8; ++j) win[slot][d][j] = 0.f;
            return;
        }
        const uint16_t* rp = ibase + ((int64_t)y * w + xc) * in.ld;
        if (xl) load8(rp - in.ld, in.plane, win[slot][0]);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) win[slot][0][j] = 0.f;
        }
        load8(rp, in.plane, win[slot][1]);
        if (xr) load8(rp + in.ld, in.plane, win[slot][2]);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) win[slot][2][j] = 0.f;
        }
    };
    // slots hold input rows: stride 1 -> rows (oy-1, oy, oy+1); stride 2 -> rows (2oy-1, 2oy, 2oy+1)
    if (STRIDE == 1) { load_row(0, oy0 - 1); load_row(1, oy0); }
    else             { load_row(0, 2 * oy0 - 1); }
#pragma unroll
    for (int i = 0; i < kDwRB; ++i) {
        const int oy = oy0 + i;
        if (oy >= ho) break;
        // slot rotation is static because the loop is fully unrolled
        int s0, s1, s2;
        if (STRIDE == 1) {
            s0 = i % 3; s1 = (i + 1) % 3; s2 = (i + 2) % 3;
            load_row(s2, oy + 1);
        } else {
            s0 = (2 * i) % 3; s1 = (2 * i + 1) % 3; s2 = (2 * i + 2) % 3;
            load_row(s1, 2 * oy);
            load_row(s2, 2 * oy + 1);
        }
        float acc[8];
        {
            const float4 b0 = sb[gl][0], b1 = sb[gl][1];
            acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
        }
        const int slots[3] = {s0, s1, s2};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 w0 = sw[ky * 3 + kx][gl][0], w1 = sw[ky * 3 + kx][gl][1];
                const float* v = win[slots[ky]][kx];
                acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]);
                acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
                acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]);
                acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
            }
        if (relu6) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = relu6f(acc[j]);
        }
        store8(out.p + (((int64_t)img * ho + oy) * wo + ox) * out.ld + c0, out.plane, acc);
    }
}

// ---------------------------------------------------------------------------------------------------
// bilinear, align_corners=True, into a concat slot; output frame i reads source frame i % n_src
// [model.py:152-153, 360-361; ATen upsample_bilinear2d: ratio=(in-1)/(out-1), src=ratio*dst, l1=src-floor]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bilinear_ac_kernel(Act in, int n_src, int hs, int ws, int c, ActW out,
                                                          int n_dst, int hd, int wd, float ry, float rx, int src_group, int dst_group,
                                                          int rows_per_cta) {
    pdl_trigger();
    pdl_wait();
    // thread = (output column, 8-channel group) of one output frame (blockIdx.z); it walks down rows_per_cta output rows.  The
    // horizontal interpolation of a source row (lx0*a + lx1*b) is computed once and reused by every output row between two
    // source rows (3.75 of them when 12x20 maps go to 45x80); the first version recomputed all four taps per output pixel and
    // spent ~200 instructions per 8 outputs, which - not memory - was what the kernel was bound by.
    const unsigned groups = (unsigned)c >> 3;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)wd * groups) return;
    const int ox = (int)(idx / groups);
    const int g = (int)(idx - (unsigned)ox * groups);
    const int img = blockIdx.z;
    const int oy_begin = blockIdx.y * rows_per_cta, oy_end = min(hd, oy_begin + rows_per_cta);
    // output frame i = (call g, local j) reads source frame g*src_group + j % (sources of call g): with one group this is the
    // reference's repeat(T) interleave i % n_src (quirk Q3); several groups = several reference calls batched in one launch
    const int grp = img / dst_group, j = img - grp * dst_group;
    const int nsg = min(src_group, n_src - grp * src_group);
    const int simg = grp * src_group + j % nsg;
    const float sx = __fmul_rn(rx, (float)ox);
    const int x0 = (int)sx;
    const int x1 = x0 + (x0 < ws - 1 ? 1 : 0);
    const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
    const uint16_t* base = in.p + (int64_t)simg * hs * ws * in.ld + g * 8;
    float h0[8], h1[8];
    int cy0 = -1, cy1 = -1;                                                    // source rows held in h0 / h1
    auto hrow = [&](int y, float (&h)[8]) {
        float a[8], b[8];
        load8(base + ((int64_t)y * ws + x0) * in.ld, in.plane, a);
        load8(base + ((int64_t)y * ws + x1) * in.ld, in.plane, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) h[k] = lx0 * a[k] + lx1 * b[k];
    };
    uint16_t* dst = out.p + (((int64_t)img * hd + oy_begin) * wd + ox) * out.ld + g * 8;
    for (int oy = oy_begin; oy < oy_end; ++oy, dst += (int64_t)wd * out.ld) {
        const float sy = __fmul_rn(ry, (float)oy);
        const int y0 = (int)sy;
        const int y1 = y0 + (y0 < hs - 1 ? 1 : 0);
        const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
        if (y0 != cy0) {
            if (y0 == cy1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) h0[k] = h1[k];
            } else {
                hrow(y0, h0);
            }
            cy0 = y0;
        }
        if (y1 != cy1) {
            if (y1 == y0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) h1[k] = h0[k];
            } else {
                hrow(y1, h1);
            }
            cy1 = y1;
        }
        float r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = ly0 * h0[k] + ly1 * h1[k];
        store8(dst, out.plane, r);
    }
}

// ---------------------------------------------------------------------------------------------------
// teConv_sub neighbour differences [model.py:194-200]
//   first half : x[i]-x[i-1]   (i=0: x[1]-x[0])
//   second half: x[i]-x[i+1]   (i=n-1: x[n-2]-x[n-1])
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tdiff_cat_kernel(Act in, int n, int hw, int c, ActW out, int group) {
    pdl_trigger();
    pdl_wait();
    const int groups = c >> 3;
    const int64_t total = (int64_t)n * hw * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;          // img*hw + p
    const int img = (int)(pix / hw);
    const int64_t fs = (int64_t)hw * in.ld;  // frame stride
    const uint16_t* cur = in.p + pix * in.ld + g * 8;
    float x[8], pv[8], nx[8], d0[8], d1[8];
    load8(cur, in.plane, x);
    // the mirrored edges sit at the boundaries of each reference call (`group` frames; the last group may be shorter)
    const int g0 = (img / group) * group, g1 = min(n, g0 + group);
    const bool has_prev = img > g0, has_next = img < g1 - 1;
    if (has_prev) load8(cur - fs, in.plane, pv);
    if (has_next) load8(cur + fs, in.plane, nx);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        d0[j] = has_prev ? x[j] - pv[j] : nx[j] - x[j];
        d1[j] = has_next ? x[j] - nx[j] : pv[j] - x[j];
    }
    uint16_t* o = out.p + pix * out.ld + g * 8;
    store8(o, out.plane, d0);
    store8(o + c, out.plane, d1);
}

// context prior: sum over the T frames of each chunk [model.py:357-358]
__global__ void __launch_bounds__(256) ctx_sum_kernel(Act in, int b, int t, int hw, int c, ActW out) {
    pdl_trigger();
    pdl_wait();
    const int groups = c >> 3;
    const int64_t total = (int64_t)b * hw * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;          // chunk*hw + p
    const int chunk = (int)(pix / hw);
    const int p = (int)(pix % hw);
    float s[8], v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (int k = 0; k < t; ++k) {
        load8(in.p + (((int64_t)chunk * t + k) * hw + p) * in.ld + g * 8, in.plane, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
    store8(out.p + pix * out.ld + g * 8, out.plane, s);
}

__global__ void __launch_bounds__(256) add_kernel(Act a, Act b, int64_t rows, int c, ActW out) {
    const int groups = c >> 3;
    const int64_t total = rows * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    const int64_t r = i / groups;
    float x[8], y[8];
    load8(a.p + r * a.ld + g * 8, a.plane, x);
    load8(b.p + r * b.ld + g * 8, b.plane, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    store8(out.p + r * out.ld + g * 8, out.plane, x);
}

// readout project (k -> 1) + BN + sigmoid: one warp per row [conv_out_st.conv.2/.3, model.py:373]
__global__ void __launch_bounds__(256) dot_sigmoid_kernel(Act a, int64_t rows, int k, const float* __restrict__ wgt,
                                                          float bias, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const uint16_t* ar = a.p + row * a.ld;
    float s = 0.f;
    for (int k0 = lane * 8; k0 < k; k0 += 256) {
        float v[8];
        load8(ar + k0, a.plane, v);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wgt + k0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wgt + k0 + 4));
        s += v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w + v[4] * w1.x + v[5] * w1.y + v[6] * w1.z + v[7] * w1.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = sigmoid_acc(s + bias);
}

// ---------------------------------------------------------------------------------------------------
// postprocess_predictions + np2mat [utils_data.py:289-303, 68-82]; cv2.resize INTER_LINEAR semantics
// ---------------------------------------------------------------------------------------------------
struct PostGeom {
    int hs, ws;        // source map
    int rh, rw;        // resized (before crop)
    int oy, ox;        // crop offset
    int hd, wd;        // output
    double sy, sx;     // cv2 scale = src/dst
};

__device__ __forceinline__ void cv_tap(int d, double scale, int srcn, int& s0, int& s1, float& w1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= srcn - 1) { s = srcn - 1; f = 0.f; }
    s0 = s;
    s1 = min(s + 1, srcn - 1);
    w1 = f;
}

__device__ __forceinline__ float post_value(const float* __restrict__ m, const PostGeom& g, int y, int x) {
    int y0, y1, x0, x1;
    float wy, wx;
    cv_tap(y + g.oy, g.sy, g.hs, y0, y1, wy);
    cv_tap(x + g.ox, g.sx, g.ws, x0, x1, wx);
    const float ax = 1.f - wx, ay = 1.f - wy;
    const float r0 = __fadd_rn(__fmul_rn(m[y0 * g.ws + x0], ax), __fmul_rn(m[y0 * g.ws + x1], wx));
    const float r1 = __fadd_rn(__fmul_rn(m[y1 * g.ws + x0], ax), __fmul_rn(m[y1 * g.ws + x1], wx));
    return __fadd_rn(__fmul_rn(r0, ay), __fmul_rn(r1, wy));
}

// The u8 path works on quads of 4 output pixels of one row: the row taps (fp64 coordinate arithmetic) are evaluated once
// per quad and the column taps come from a per-block table in shared memory, with exactly the arithmetic of post_value()
// (the first version recomputed both fp64 tap pairs for every pixel in both kernels).
constexpr int kPostMaxW = 2048;

__device__ __forceinline__ void post_xtaps(const PostGeom& g, uint2* xt) {
    for (int x = threadIdx.x; x < g.wd; x += blockDim.x) {
        int x0, x1;
        float wx;
        cv_tap(x + g.ox, g.sx, g.ws, x0, x1, wx);
        xt[x] = make_uint2((uint32_t)x0 | ((uint32_t)x1 << 16), __float_as_uint(wx));
    }
    __syncthreads();
}

__device__ __forceinline__ void post_quad(const float* __restrict__ m, const PostGeom& g, const uint2* xt, int y, int x, float v[4]) {
    int y0, y1;
    float wy;
    cv_tap(y + g.oy, g.sy, g.hs, y0, y1, wy);
    const float ay = 1.f - wy;
    const float* r0p = m + y0 * g.ws;
    const float* r1p = m + y1 * g.ws;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint2 t = xt[x + j];
        const int x0 = t.x & 0xFFFF, x1 = t.x >> 16;
        const float wx = __uint_as_float(t.y), ax = 1.f - wx;
        const float r0 = __fadd_rn(__fmul_rn(r0p[x0], ax), __fmul_rn(r0p[x1], wx));
        const float r1 = __fadd_rn(__fmul_rn(r1p[x0], ax), __fmul_rn(r1p[x1], wx));
        v[j] = __fadd_rn(__fmul_rn(r0, ay), __fmul_rn(r1, wy));
    }
}

__global__ void __launch_bounds__(256) post_max_kernel(const float* __restrict__ maps, PostGeom g,
                                                       float* __restrict__ frame_max) {
    __shared__ uint2 xt[kPostMaxW];
    __shared__ float s[8];
    pdl_trigger();
    post_xtaps(g, xt);
    pdl_wait();
    const int img = blockIdx.y;
    const float* m = maps + (int64_t)img * g.hs * g.ws;
    const int qpr = g.wd >> 2, total4 = g.hd * qpr;     // wd % 4 == 0 is required by the host wrapper
    float mx = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const int y = i / qpr, x = (i - y * qpr) << 2;
        float v[4];
        post_quad(m, g, xt, y, x, v);
        mx = fmaxf(fmaxf(mx, fmaxf(v[0], v[1])), fmaxf(v[2], v[3]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmaxf(mx, s[k]);
        atomicMax(reinterpret_cast<int*>(frame_max) + img, __float_as_int(mx));   // maps are positive (sigmoid)
    }
}

__global__ void __launch_bounds__(256) post_write_kernel(const float* __restrict__ maps, PostGeom g,
                                                         const float* __restrict__ frame_max,
                                                         uint8_t* __restrict__ out) {
    __shared__ uint2 xt[kPostMaxW];
    pdl_trigger();
    post_xtaps(g, xt);
    pdl_wait();
    const int img = blockIdx.y;
    const float* m = maps + (int64_t)img * g.hs * g.ws;
    const float mx = frame_max[img];
    const int qpr = g.wd >> 2, total4 = g.hd * qpr;
    uint32_t* o = reinterpret_cast<uint32_t*>(out + (int64_t)img * g.hd * g.wd);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const int y = i / qpr, x = (i - y * qpr) << 2;
        float q[4];
        post_quad(m, g, xt, y, x, q);
        uint32_t pk = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = __fmul_rn(__fdiv_rn(q[j], mx), 255.f);
            v = fminf(fmaxf(v, 0.f), 255.f);
            pk |= ((uint32_t)__float2int_rn(v) & 0xFFu) << (8 * j);   // rint = round half to even
        }
        o[i] = pk;
    }
}

// fp32 output (tests / any width): one pixel per thread
__global__ void __launch_bounds__(256) post_max_px_kernel(const float* __restrict__ maps, PostGeom g,
                                                          float* __restrict__ frame_max) {
    pdl_trigger();
    pdl_wait();
    const int img = blockIdx.y;
    const float* m = maps + (int64_t)img * g.hs * g.ws;
    const int total = g.hd * g.wd;
    float mx = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        mx = fmaxf(mx, post_value(m, g, i / g.wd, i % g.wd));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmaxf(mx, s[k]);
        atomicMax(reinterpret_cast<int*>(frame_max) + img, __float_as_int(mx));
    }
}

__global__ void __launch_bounds__(256) post_write_f32_kernel(const float* __restrict__ maps, PostGeom g,
                                                             const float* __restrict__ frame_max,
                                                             float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int img = blockIdx.y;
    const float* m = maps + (int64_t)img * g.hs * g.ws;
    const float mx = frame_max[img];
    const int total = g.hd * g.wd;
    float* o = out + (int64_t)img * total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        o[i] = __fmul_rn(__fdiv_rn(post_value(m, g, i / g.wd, i % g.wd), mx), 255.f);
}

}  // namespace uavsal

// ===================================================================================================
// C ABI
// ===================================================================================================
using namespace uavsal;

namespace uavsal { int g_pdl = 1; }
int g_dw_fast = 2;     // uavsal_set_option key 2: 2 = TMA-staged depthwise kernel (default), 1 = sliding window, 0 = generic
namespace uavsal {
int dw3x3_tma(Act in, int n, int h, int w, int c, int stride, const float* wgt, const float* bias, int relu6, ActW out,
              cudaStream_t s);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool act_ok(const void* p, int64_t plane, int ld) {
    return p != nullptr && aligned16(p) && (ld % 8) == 0 && (plane % 8) == 0 && plane >= 0;
}

namespace uavsal {
bool dw3x3_img_fits(int64_t in_plane, int h, int w);
int dw3x3_img(Act in, int n, int h, int w, int c, int dil, const float* wgt, const float* bias, int relu6, ActW out, cudaStream_t s);
int dw3x3_dot_tma(const void* in, bool q16, int in_ld, int n, int h, int w, int c, const float* wgt, const float* bias, const float* wproj,
                  float bias_proj, float* partial, float* out, cudaStream_t s);
}

template <bool PW>
static int stem_launch(const void* x, int x_kind, int n, int h, int w, const float* wgt, const float* bias, const StemW* hw,
                       uint16_t* out, int64_t out_plane, int out_ld, void* stream) {
    UAVSAL_REQUIRE(x && (PW ? hw != nullptr : (wgt && bias && aligned16(wgt))) &&
                       (out_plane == UAVSAL_PLANE_F32 ? (out && aligned16(out) && out_ld % 4 == 0) : act_ok(out, out_plane, out_ld)) && out_ld >= 32 && n > 0 &&
                       h >= 2 && w >= 2 && x_kind >= 0 && x_kind <= 2,
                   UAVSAL_EINVAL, "stem_conv3x3s2: bad arguments");
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    const int64_t total = (int64_t)n * ho * wo;
    UAVSAL_REQUIRE(total < (1LL << 31) - (1 << 20), UAVSAL_ENOTSUP, "stem_conv3x3s2: more than 2^31 output pixels");
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    static int bps[3] = {0, 0, 0};                                                        // resident blocks per SM (persistent grid)
    if (!bps[x_kind]) {
        int b = 0;
        cudaError_t e = x_kind == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, stem_kernel<0, PW>, 128, 0)
                      : x_kind == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, stem_kernel<1, PW>, 128, 0)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, stem_kernel<2, PW>, 128, 0);
        if (e != cudaSuccess || b <= 0) { b = 4; cudaGetLastError(); }
        bps[x_kind] = b;
    }
    const dim3 grid((unsigned)std::min<int64_t>(div_up(total, 128), (int64_t)sms * bps[x_kind]));
    ActW o{out, out_plane, out_ld};
    cudaStream_t s = (cudaStream_t)stream;
    static const StemW none{};
    const StemW& cw = PW ? *hw : none;
    if (x_kind == 0) launch_k(stem_kernel<0, PW>, dim3(grid), dim3(128), 0, s, 1, x, n, h, w, ho, wo, wgt, bias, o, cw);
    else if (x_kind == 1) launch_k(stem_kernel<1, PW>, dim3(grid), dim3(128), 0, s, 1, x, n, h, w, ho, wo, wgt, bias, o, cw);
    else launch_k(stem_kernel<2, PW>, dim3(grid), dim3(128), 0, s, 1, x, n, h, w, ho, wo, wgt, bias, o, cw);
    return check_launch("stem_conv3x3s2");
}

extern "C" {

int uavsal_version(void) { return 1; }
const char* uavsal_arch(void) { return "sm_100a"; }
const char* uavsal_last_error(void) { return g_err; }

int uavsal_device_ok(int device) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return (int)e; }
    UAVSAL_REQUIRE(p.major == 10, UAVSAL_ENOTSUP, "device %d is sm_%d%d, this library is sm_100a only", device, p.major, p.minor);
    return 0;
}

int uavsal_pack_nchw_f32(const float* src, int n, int c, int h, int w, uint16_t* dst, int64_t plane, int ld, int cpad,
                         void* stream) {
    UAVSAL_REQUIRE(src && act_ok(dst, plane, ld) && cpad % 8 == 0 && cpad >= c && ld >= cpad && n > 0, UAVSAL_EINVAL,
                   "pack_nchw_f32: bad arguments");
    const int64_t total = (int64_t)n * h * w * (cpad / 8);
    launch_k(pack_nchw_kernel, dim3(div_up(total, 256)), dim3(256), 0, (cudaStream_t)stream, 1, src, n, c, h * w, ActW{dst, plane, ld}, cpad);
    return check_launch("pack_nchw_f32");
}

int uavsal_unpack_nchw_f32(const uint16_t* src, int64_t plane, int ld, int n, int c, int h, int w, float* dst,
                           void* stream) {
    UAVSAL_REQUIRE(dst && src && ld >= c && n > 0, UAVSAL_EINVAL, "unpack_nchw_f32: bad arguments");
    const int64_t total = (int64_t)n * c * h * w;
    launch_k(unpack_nchw_kernel, dim3(div_up(total, 256)), dim3(256), 0, (cudaStream_t)stream, 1, Act{src, plane, ld}, n, c, h * w, dst);
    return check_launch("unpack_nchw_f32");
}

int uavsal_stem_conv3x3s2(const void* x, int x_kind, int n, int h, int w, const float* wgt, const float* bias,
                          uint16_t* out, int64_t out_plane, int out_ld, void* stream) {
    return stem_launch<false>(x, x_kind, n, h, w, wgt, bias, nullptr, out, out_plane, out_ld, stream);
}

int uavsal_stem_conv3x3s2_hw(const void* x, int x_kind, int n, int h, int w, const float* wgt_host, const float* bias_host,
                             uint16_t* out, int64_t out_plane, int out_ld, void* stream) {
    UAVSAL_REQUIRE(wgt_host && bias_host, UAVSAL_EINVAL, "stem_conv3x3s2_hw: host weight pointers required");
    StemW hw;
    memcpy(hw.w, wgt_host, 27 * 32 * sizeof(float));
    memcpy(hw.w + 27 * 32, bias_host, 32 * sizeof(float));
    return stem_launch<true>(x, x_kind, n, h, w, nullptr, nullptr, &hw, out, out_plane, out_ld, stream);
}

int uavsal_dw3x3(const uint16_t* in, int64_t in_plane, int in_ld, int n, int h, int w, int c, int stride, int dilation,
                 const float* wgt, const float* bias, int relu6, uint16_t* out, int64_t out_plane, int out_ld,
                 void* stream) {
    const bool q16in = in_plane == UAVSAL_PLANE_Q16;
    const bool f32in = in_plane == UAVSAL_PLANE_F32 || q16in;              // plain rows (fp32 | q16): TMA kernel only
    UAVSAL_REQUIRE((f32in ? (in && aligned16(in) && in_ld % (q16in ? 8 : 4) == 0) : act_ok(in, in_plane, in_ld)) && act_ok(out, out_plane, out_ld) &&
                       wgt && bias && aligned16(wgt) && aligned16(bias) && c % 8 == 0 && c > 0 && in_ld >= c && out_ld >= c && n > 0,
                   UAVSAL_EINVAL, "dw3x3: bad arguments (c=%d)", c);
    UAVSAL_REQUIRE((stride == 1 || stride == 2) && dilation >= 1 && (stride == 1 || dilation == 1), UAVSAL_ENOTSUP,
                   "dw3x3: stride %d dilation %d unsupported", stride, dilation);   // model.py:78 assert stride in [1,2]
    const int ho = stride == 1 ? h : (h - 1) / 2 + 1, wo = stride == 1 ? w : (w - 1) / 2 + 1;
    if (dilation == 1 && (g_dw_fast == 2 || f32in) && in_plane != 0 && out_plane != 0 && ((int64_t)n * ho * wo >= 512 || f32in)) {
        return dw3x3_tma(Act{in, in_plane, in_ld}, n, h, w, c, stride, wgt, bias, relu6, ActW{out, out_plane, out_ld},
                         (cudaStream_t)stream);
    }
    if (dilation > 1 && stride == 1 && f32in && out_plane != 0 && dw3x3_img_fits(in_plane, h, w)) {
        // small maps with dilation (the ASPP branches at 12x20), plain-row input: the whole image of a channel block staged by TMA
        // (split-bf16 input stays with the generic kernel below: two 61 KB images per CTA leave one CTA per SM - measured slower)
        return dw3x3_img(Act{in, in_plane, in_ld}, n, h, w, c, dilation, wgt, bias, relu6, ActW{out, out_plane, out_ld}, (cudaStream_t)stream);
    }
    UAVSAL_REQUIRE(!f32in, UAVSAL_ENOTSUP, "dw3x3: fp32 / q16 row input is only implemented by the TMA kernels (dilation 1, or small maps)");
    if (dilation == 1 && g_dw_fast == 1) {
        const int strips = div_up(ho, kDwRB);
        const dim3 grid(div_up(c, 64), div_up(wo, 32), strips * n);
        if (grid.y <= 65535 && grid.z <= 65535) {
            if (stride == 1)
                dw3x3_rows_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(Act{in, in_plane, in_ld}, h, w, c, ho, wo, wgt, bias,
                                                                           relu6, ActW{out, out_plane, out_ld}, strips);
            else
                dw3x3_rows_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(Act{in, in_plane, in_ld}, h, w, c, ho, wo, wgt, bias,
                                                                           relu6, ActW{out, out_plane, out_ld}, strips);
            return check_launch("dw3x3(rows)");
        }
    }
    const int64_t total = (int64_t)n * ho * wo * (c / 8);
    launch_k(dw3x3_kernel, dim3(div_up(total, 256)), dim3(256), 0, (cudaStream_t)stream, 1, Act{in, in_plane, in_ld}, n, h, w, c, ho, wo,
                                                                      stride, dilation, wgt, bias, relu6,
                                                                      ActW{out, out_plane, out_ld});
    return check_launch("dw3x3");
}

int uavsal_bilinear_ac(const uint16_t* in, int64_t in_plane, int in_ld, int n_src, int hs, int ws, int c, uint16_t* out,
                       int64_t out_plane, int out_ld, int n_dst, int hd, int wd, int src_group, int dst_group, void* stream) {
    UAVSAL_REQUIRE(act_ok(in, in_plane, in_ld) && act_ok(out, out_plane, out_ld) && c % 8 == 0 && c > 0 && n_src > 0 &&
                       n_dst > 0 && hs > 0 && ws > 0 && hd > 0 && wd > 0,
                   UAVSAL_EINVAL, "bilinear_ac: bad arguments");
    const float ry = hd > 1 ? (float)(hs - 1) / (float)(hd - 1) : 0.f;
    const float rx = wd > 1 ? (float)(ws - 1) / (float)(wd - 1) : 0.f;
    if (src_group <= 0 || dst_group <= 0) { src_group = n_src; dst_group = n_dst; }
    UAVSAL_REQUIRE((int64_t)div_up(n_dst, dst_group) * src_group >= n_src && div_up(n_dst, dst_group) == div_up(n_src, src_group), UAVSAL_EINVAL,
                   "bilinear_ac: %d sources in groups of %d do not match %d outputs in groups of %d", n_src, src_group, n_dst, dst_group);
    UAVSAL_REQUIRE(hd <= 65535 && n_dst <= 65535, UAVSAL_ENOTSUP, "bilinear_ac: more than 65535 output rows / frames");
    // a thread walks down a strip of output rows (the source-row interpolation is reused); the strips are as long as still leaves
    // about four CTAs per SM
    const int64_t ctas_full = (int64_t)div_up((int64_t)wd * (c / 8), 256) * n_dst;
    int chunks = (int)((4LL * 148 + ctas_full - 1) / ctas_full);
    if (chunks < 1) chunks = 1;
    if (chunks > hd) chunks = hd;
    const int rows_per_cta = div_up(hd, chunks);
    launch_k(bilinear_ac_kernel, dim3(div_up((int64_t)wd * (c / 8), 256), div_up(hd, rows_per_cta), n_dst), dim3(256), 0, (cudaStream_t)stream, 1,
             Act{in, in_plane, in_ld}, n_src, hs, ws, c, ActW{out, out_plane, out_ld}, n_dst, hd, wd, ry, rx, src_group, dst_group, rows_per_cta);
    return check_launch("bilinear_ac");
}

int uavsal_tdiff_cat(const uint16_t* in, int64_t in_plane, int in_ld, int n, int hw, int c, uint16_t* out,
                     int64_t out_plane, int out_ld, int group, void* stream) {
    UAVSAL_REQUIRE(act_ok(in, in_plane, in_ld) && act_ok(out, out_plane, out_ld) && c % 8 == 0 && c > 0 && out_ld >= 2 * c,
                   UAVSAL_EINVAL, "tdiff_cat: bad arguments");
    if (group <= 0 || group > n) group = n;
    UAVSAL_REQUIRE(n >= 2 && group >= 2 && (n % group == 0 || n % group >= 2), UAVSAL_EINVAL,
                   "tdiff_cat: needs at least 2 frames per call (model.py:194 indexes x1[1])");
    const int64_t total = (int64_t)n * hw * (c / 8);
    launch_k(tdiff_cat_kernel, dim3(div_up(total, 256)), dim3(256), 0, (cudaStream_t)stream, 1, Act{in, in_plane, in_ld}, n, hw, c,
                                                                          ActW{out, out_plane, out_ld}, group);
    return check_launch("tdiff_cat");
}

int uavsal_ctx_sum(const uint16_t* in, int64_t in_plane, int in_ld, int b, int t, int hw, int c, uint16_t* out,
                   int64_t out_plane, int out_ld, void* stream) {
    UAVSAL_REQUIRE(act_ok(in, in_plane, in_ld) && act_ok(out, out_plane, out_ld) && c % 8 == 0 && c > 0 && b > 0 && t > 0,
                   UAVSAL_EINVAL, "ctx_sum: bad arguments");
    const int64_t total = (int64_t)b * hw * (c / 8);
    launch_k(ctx_sum_kernel, dim3(div_up(total, 256)), dim3(256), 0, (cudaStream_t)stream, 1, Act{in, in_plane, in_ld}, b, t, hw, c,
                                                                        ActW{out, out_plane, out_ld});
    return check_launch("ctx_sum");
}

int uavsal_add(const uint16_t* a, int64_t a_plane, int a_ld, const uint16_t* b, int64_t b_plane, int b_ld, int64_t rows,
               int c, uint16_t* out, int64_t out_plane, int out_ld, void* stream) {
    UAVSAL_REQUIRE(act_ok(a, a_plane, a_ld) && act_ok(b, b_plane, b_ld) && act_ok(out, out_plane, out_ld) && c % 8 == 0 &&
                       c > 0 && rows > 0,
                   UAVSAL_EINVAL, "add: bad arguments");
    add_kernel<<<div_up(rows * (c / 8), 256), 256, 0, (cudaStream_t)stream>>>(Act{a, a_plane, a_ld}, Act{b, b_plane, b_ld},
                                                                             rows, c, ActW{out, out_plane, out_ld});
    return check_launch("add");
}

int uavsal_dw3x3_dot_sigmoid(const float* in, int in_ld, int n, int h, int w, int c, const float* wd, const float* bd,
                             const float* wproj, float bias_proj, float* partial_ws, float* out_f32, void* stream) {
    UAVSAL_REQUIRE(in && wd && bd && wproj && partial_ws && out_f32 && aligned16(in) && aligned16(wd) && aligned16(bd) && aligned16(wproj) &&
                       n > 0 && h > 0 && w > 0 && c > 0 && in_ld % 4 == 0 && in_ld >= c,
                   UAVSAL_EINVAL, "dw3x3_dot_sigmoid: bad arguments");
    UAVSAL_REQUIRE(c % 4 == 0, UAVSAL_ENOTSUP, "dw3x3_dot_sigmoid: channels must be a multiple of 4");
    return dw3x3_dot_tma(in, false, in_ld, n, h, w, c, wd, bd, wproj, bias_proj, partial_ws, out_f32, (cudaStream_t)stream);
}

int uavsal_dw3x3_dot_sigmoid_q16(const uint16_t* in, int in_ld, int n, int h, int w, int c, const float* wd, const float* bd,
                                 const float* wproj, float bias_proj, float* partial_ws, float* out_f32, void* stream) {
    UAVSAL_REQUIRE(in && wd && bd && wproj && partial_ws && out_f32 && aligned16(in) && aligned16(wd) && aligned16(bd) && aligned16(wproj) &&
                       n > 0 && h > 0 && w > 0 && c > 0 && in_ld % 8 == 0 && in_ld >= c,
                   UAVSAL_EINVAL, "dw3x3_dot_sigmoid_q16: bad arguments");
    UAVSAL_REQUIRE(c % 4 == 0, UAVSAL_ENOTSUP, "dw3x3_dot_sigmoid_q16: channels must be a multiple of 4");
    return dw3x3_dot_tma(in, true, in_ld, n, h, w, c, wd, bd, wproj, bias_proj, partial_ws, out_f32, (cudaStream_t)stream);
}

int uavsal_dot_sigmoid(const uint16_t* a, int64_t a_plane, int a_ld, int64_t rows, int k, const float* wgt, float bias,
                       float* out_f32, void* stream) {
    UAVSAL_REQUIRE(act_ok(a, a_plane, a_ld) && wgt && aligned16(wgt) && out_f32 && k % 8 == 0 && k > 0 && rows > 0,
                   UAVSAL_EINVAL, "dot_sigmoid: bad arguments");
    launch_k(dot_sigmoid_kernel, dim3(div_up(rows * 32, 256)), dim3(256), 0, (cudaStream_t)stream, 1, Act{a, a_plane, a_ld}, rows, k, wgt, bias,
                                                                                out_f32);
    return check_launch("dot_sigmoid");
}

static int post_common(const float* maps, int n, int hs, int ws, int hd, int wd, float* frame_max, uint8_t* out_u8,
                       float* out_f32, void* stream) {
    UAVSAL_REQUIRE(maps && frame_max && (out_u8 || out_f32) && n > 0 && hs > 0 && ws > 0 && hd > 0 && wd > 0, UAVSAL_EINVAL,
                   "post: bad arguments");
    UAVSAL_REQUIRE(out_f32 || (wd % 4 == 0 && wd <= kPostMaxW && ws < 65536 && (reinterpret_cast<uintptr_t>(out_u8) & 3) == 0), UAVSAL_ENOTSUP,
                   "post_u8: output width must be a multiple of 4 and <= 2048");
    PostGeom g;
    g.hs = hs; g.ws = ws; g.hd = hd; g.wd = wd;
    // utils_data.py:291-301: compare rates, resize keeping aspect, centre-crop
    const double rows_rate = (double)hd / hs, cols_rate = (double)wd / ws;
    if (rows_rate > cols_rate) {
        g.rh = hd; g.rw = (int)(((int64_t)ws * hd) / hs);
        g.oy = 0; g.ox = (g.rw - wd) / 2;
    } else {
        g.rw = wd; g.rh = (int)(((int64_t)hs * wd) / ws);
        g.ox = 0; g.oy = (g.rh - hd) / 2;
    }
    UAVSAL_REQUIRE(g.rh >= hd && g.rw >= wd, UAVSAL_ENOTSUP, "post_u8: letterbox geometry not supported");
    g.sy = (double)hs / g.rh;
    g.sx = (double)ws / g.rw;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(frame_max, 0, sizeof(float) * n, s);
    if (e != cudaSuccess) { set_error("post_u8 memset: %s", cudaGetErrorString(e)); return (int)e; }
    const int bpf = max(1, min(64, div_up((int64_t)hd * wd, 256 * 16)));
    const bool quads = wd % 4 == 0 && wd <= kPostMaxW && ws < 65536;
    if (quads) launch_k(post_max_kernel, dim3(dim3(bpf, n)), dim3(256), 0, s, 1, maps, g, frame_max);
    else       launch_k(post_max_px_kernel, dim3(dim3(bpf, n)), dim3(256), 0, s, 1, maps, g, frame_max);
    if (out_u8) launch_k(post_write_kernel, dim3(dim3(bpf, n)), dim3(256), 0, s, 1, maps, g, frame_max, out_u8);
    else        launch_k(post_write_f32_kernel, dim3(dim3(bpf, n)), dim3(256), 0, s, 1, maps, g, frame_max, out_f32);
    return check_launch("post");
}

int uavsal_post_u8(const float* maps, int n, int hs, int ws, int hd, int wd, float* frame_max, uint8_t* out_u8,
                   void* stream) {
    return post_common(maps, n, hs, ws, hd, wd, frame_max, out_u8, nullptr, stream);
}

int uavsal_post_f32(const float* maps, int n, int hs, int ws, int hd, int wd, float* frame_max, float* out_f32,
                    void* stream) {
    return post_common(maps, n, hs, ws, hd, wd, frame_max, nullptr, out_f32, stream);
}

}  // extern "C"

// Video front-end after decode: utils_data.padding (:321-343) as used by preprocess_videos (:255-287) - aspect-preserving
// cv2.resize (uint8, INTER_LINEAR) into a zero-padded (shape_r, shape_c) canvas, plus the BGR -> RGB swap of :270 - on the
// device, so decoded frames go host -> HBM once at their native size and come out as the (n, H, W, 3) uint8 tensor the stem
// kernel consumes (x_kind 2).
//
// Bit-exact with OpenCV's 8-bit bilinear path (resize.cpp; pinned against cv2 in tests/): per axis
//   f = (float)((d + 0.5) * scale - 0.5), s = floor(f), weights round-half-even((1-f)*2048), round-half-even(f*2048)
//   (x taps clamp with a zeroed fraction at the border, y taps clamp their row indices),
//   horizontal: S = p[s]*a0 + p[s+1]*a1 (int),  vertical: (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2,
// the exact-2x case runs OpenCV's INTER_AREA 2x2 average (a+b+c+d+2)>>2 (resize() switches to it), equal sizes copy.
// HBM-bound byte work: thread = one output pixel (3 bytes); per-block tap tables for the columns in shared memory.
#include "common.cuh"

namespace uavsal {

struct LbArgs {
    const uint8_t* src;     // (n, sh, sw, 3)
    uint8_t* dst;           // (n, dh, dw, 3)
    int n, sh, sw, dh, dw;
    int nw, nh, ox, oy;     // resized size and its offset inside the canvas
    int swap_rb;            // 1: write channel 2-c (BGR -> RGB)
    int mode;               // 0 bilinear, 1 exact 2x (area), 2 copy
    double scale_x, scale_y;
};

__device__ __forceinline__ void lb_tap(int d, double scale, int srcn, bool zero_at_border, int& s, int& a0, int& a1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int si = (int)floorf(f);
    f -= (float)si;
    if (zero_at_border) {
        if (si < 0) { f = 0.f; si = 0; }
        if (si >= srcn - 1) { f = 0.f; si = srcn - 1; }
    }
    s = si;
    a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

constexpr int kLbMaxW = 4096;

__global__ void __launch_bounds__(256) letterbox_kernel(const LbArgs g) {
    __shared__ int xs[kLbMaxW];          // source column of the left tap
    __shared__ short xa[kLbMaxW][2];     // its two weights
    if (g.mode == 0)
        for (int x = threadIdx.x; x < g.nw; x += blockDim.x) {
            int s, a0, a1;
            lb_tap(x, g.scale_x, g.sw, true, s, a0, a1);
            xs[x] = s; xa[x][0] = (short)a0; xa[x][1] = (short)a1;
        }
    __syncthreads();
    const int y = blockIdx.y, img = blockIdx.z;
    uint8_t* drow = g.dst + ((int64_t)img * g.dh + y) * g.dw * 3;
    const int ry = y - g.oy;                                       // row inside the resized image
    const bool row_in = ry >= 0 && ry < g.nh;
    const uint8_t* simg = g.src + (int64_t)img * g.sh * g.sw * 3;
    int sy = 0, b0 = 0, b1 = 0;
    if (row_in && g.mode == 0) lb_tap(ry, g.scale_y, g.sh, false, sy, b0, b1);
    const int y0 = min(max(sy, 0), g.sh - 1), y1 = min(max(sy + 1, 0), g.sh - 1);
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < g.dw; x += gridDim.x * blockDim.x) {
        const int rx = x - g.ox;
        int v[3] = {0, 0, 0};
        if (row_in && rx >= 0 && rx < g.nw) {
            if (g.mode == 2) {
                const uint8_t* p = simg + ((int64_t)ry * g.sw + rx) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = p[c];
            } else if (g.mode == 1) {
                const uint8_t* p = simg + ((int64_t)(2 * ry) * g.sw + 2 * rx) * 3;
                const uint8_t* q = p + (int64_t)g.sw * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = (p[c] + p[3 + c] + q[c] + q[3 + c] + 2) >> 2;
            } else {
                const int s0 = xs[rx], s1 = min(s0 + 1, g.sw - 1), a0 = xa[rx][0], a1 = xa[rx][1];
                const uint8_t* r0 = simg + (int64_t)y0 * g.sw * 3;
                const uint8_t* r1 = simg + (int64_t)y1 * g.sw * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int h0 = r0[s0 * 3 + c] * a0 + r0[s1 * 3 + c] * a1;
                    const int h1 = r1[s0 * 3 + c] * a0 + r1[s1 * 3 + c] * a1;
                    const int o = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                    v[c] = min(max(o, 0), 255);
                }
            }
        }
        if (g.swap_rb) { const int t = v[0]; v[0] = v[2]; v[2] = t; }
        drow[x * 3 + 0] = (uint8_t)v[0];
        drow[x * 3 + 1] = (uint8_t)v[1];
        drow[x * 3 + 2] = (uint8_t)v[2];
    }
}

}  // namespace uavsal

using namespace uavsal;

extern "C" int uavsal_letterbox_u8(const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw, int swap_rb, void* stream) {
    UAVSAL_REQUIRE(src && dst && n > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0, UAVSAL_EINVAL, "letterbox_u8: bad arguments");
    UAVSAL_REQUIRE(dw <= kLbMaxW && dh <= 65535 && n <= 65535, UAVSAL_ENOTSUP, "letterbox_u8: output wider than %d or more than 65535 rows / frames", kLbMaxW);
    LbArgs g{};
    g.src = src; g.dst = dst; g.n = n; g.sh = sh; g.sw = sw; g.dh = dh; g.dw = dw; g.swap_rb = swap_rb ? 1 : 0;
    // utils_data.py:329-341 (Python float rates, floor divisions)
    if ((double)sh / dh > (double)sw / dw) {
        g.nw = (int)(((int64_t)sw * dh) / sh); g.nh = dh;
        UAVSAL_REQUIRE(g.nw >= 1 && g.nw <= dw, UAVSAL_ENOTSUP, "letterbox_u8: degenerate geometry");
        g.ox = (dw - g.nw) / 2; g.oy = 0;
    } else {
        g.nw = dw; g.nh = (int)(((int64_t)sh * dw) / sw);
        UAVSAL_REQUIRE(g.nh >= 1 && g.nh <= dh, UAVSAL_ENOTSUP, "letterbox_u8: degenerate geometry");
        g.ox = 0; g.oy = (dh - g.nh) / 2;
    }
    g.mode = (g.nw == sw && g.nh == sh) ? 2 : (sw == 2 * g.nw && sh == 2 * g.nh) ? 1 : 0;
    g.scale_x = (double)sw / g.nw;
    g.scale_y = (double)sh / g.nh;
    letterbox_kernel<<<dim3(div_up(dw, 256), dh, n), 256, 0, (cudaStream_t)stream>>>(g);
    return check_launch("letterbox_u8");
}

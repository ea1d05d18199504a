// Depthwise 3x3 + BN + ReLU6 (dilation 1, stride 1|2) with TMA-staged input tiles.
//
// A CTA owns a tile of 16x8 (stride 1) or 8x8 (stride 2) output pixels x 64 channels.  One thread issues two
// cp.async.bulk.tensor loads (hi and lo plane) of the haloed input box into shared memory; out-of-bounds zero fill
// provides the convolution padding, so there is no per-thread address arithmetic or predication on the input side and
// every input element is fetched from L2/HBM once per tile (1.4x halo factor) instead of nine times.  Threads then
// slide a 3x3 register window down their column of the tile (3 shared-memory reads per output), apply the folded
// BN bias + ReLU6 and write both bf16 planes with 128-byte-contiguous stores.  Several CTAs per SM overlap each other's
// TMA latency and compute.
#include "tc_common.cuh"

namespace uavsal {

struct DwArgs {
    int n, h, w, c, ho, wo;
    int tiles_x, tiles_y, cblocks, num_tiles;
    const float* wgt;      // [9][c]
    const float* bias;     // [c]
    int relu6;
    ActW out;
};

template <int STRIDE>
struct DwGeom {
    static constexpr int TW = STRIDE == 1 ? 16 : 8;       // output tile
    static constexpr int TH = 8;
    static constexpr int IW = (TW - 1) * STRIDE + 3;      // haloed input box
    static constexpr int IH = (TH - 1) * STRIDE + 3;
    static constexpr int PIX = IW * IH;
    static constexpr int RPT = TW * TH * 16 / 256;        // output rows per thread: 8 (stride 1) or 4 (stride 2)
    static constexpr uint32_t PLANE_BYTES = PIX * 128;    // 64 channels x bf16
};

// 4 channels (8 bytes per plane) from the staged tile -> fp32
__device__ __forceinline__ void lds4(const uint8_t* hi, const uint8_t* lo, float v[4]) {
    const uint2 a = *reinterpret_cast<const uint2*>(hi);
    const uint2 b = *reinterpret_cast<const uint2*>(lo);
    float t[4];
    unpack2(a.x, v[0], v[1]); unpack2(a.y, v[2], v[3]);
    unpack2(b.x, t[0], t[1]); unpack2(b.y, t[2], t[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += t[i];
}

// thread = (4-channel quad, output column); it keeps its 36 folded weights in registers for the whole tile and slides a
// 3x3x4 register window down RPT output rows, so the shared-memory/LSU pipe (the measured limiter of the first version,
// whose weights were re-read from smem for every row) only carries 3 narrow reads and one write per output.
template <int STRIDE>
__global__ void __launch_bounds__(256, 2) dw3x3_tma_kernel(const __grid_constant__ CUtensorMap tmIn, const DwArgs g) {
    using G = DwGeom<STRIDE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* s_hi = smem;
    uint8_t* s_lo = smem + G::PLANE_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_lo + G::PLANE_BYTES);

    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();

    const int quad = tid & 15;                                                 // 4-channel quad inside the 64-channel block
    const int col = (tid >> 4) % G::TW;
    const int rgrp = (tid >> 4) / G::TW;
    int last_cblk = -1;
    uint32_t parity = 0;
    float wr[9][4], br[4];
    for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
        int r = t;
        const int cblk = r % g.cblocks; r /= g.cblocks;
        const int tx = r % g.tiles_x;   r /= g.tiles_x;
        const int ty = r % g.tiles_y;
        const int img = r / g.tiles_y;
        const int x0 = tx * G::TW, y0 = ty * G::TH;
        if (tid == 0) {
            fence_async_smem();                                                // order the tile's generic reads before the async overwrite
            mbar_expect_tx(bar, 2 * G::PLANE_BYTES);
            tma_load_5d(&tmIn, bar, s_hi, cblk * 64, x0 * STRIDE - 1, y0 * STRIDE - 1, img, 0);
            tma_load_5d(&tmIn, bar, s_lo, cblk * 64, x0 * STRIDE - 1, y0 * STRIDE - 1, img, 1);
        }
        const int c0 = cblk * 64 + quad * 4;
        if (cblk != last_cblk) {                                               // block-uniform; overlaps the TMA flight time
            if (c0 < g.c) {
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(g.wgt + k * g.c + c0));
                    wr[k][0] = w4.x; wr[k][1] = w4.y; wr[k][2] = w4.z; wr[k][3] = w4.w;
                }
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + c0));
                br[0] = b4.x; br[1] = b4.y; br[2] = b4.z; br[3] = b4.w;
            }
            last_cblk = cblk;
        }
        mbar_wait(bar, parity);
        parity ^= 1;

        const int ox = x0 + col;
        if (c0 < g.c && ox < g.wo) {
            const uint8_t* bh = s_hi + quad * 8;
            const uint8_t* bl = s_lo + quad * 8;
            float win[3][3][4];
            auto load_row = [&](int slot, int iy) {                            // iy: row inside the input box
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int pix = iy * G::IW + col * STRIDE + d;
                    lds4(bh + pix * 128, bl + pix * 128, win[slot][d]);
                }
            };
            const int oyl0 = rgrp * G::RPT;
            if (STRIDE == 1) { load_row(0, oyl0); load_row(1, oyl0 + 1); }
            else             { load_row(0, oyl0 * 2); }
#pragma unroll
            for (int i = 0; i < G::RPT; ++i) {
                const int oyl = oyl0 + i;
                int s0, s1, s2;
                if (STRIDE == 1) {
                    s0 = i % 3; s1 = (i + 1) % 3; s2 = (i + 2) % 3;
                    load_row(s2, oyl + 2);
                } else {
                    s0 = (2 * i) % 3; s1 = (2 * i + 1) % 3; s2 = (2 * i + 2) % 3;
                    load_row(s1, oyl * 2 + 1);
                    load_row(s2, oyl * 2 + 2);
                }
                const int oy = y0 + oyl;
                if (oy >= g.ho) break;
                float acc[4] = {br[0], br[1], br[2], br[3]};
                const int slots[3] = {s0, s1, s2};
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float* v = win[slots[ky]][kx];
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[j] = fmaf(v[j], wr[ky * 3 + kx][j], acc[j]);
                    }
                if (g.relu6) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] = relu6f(acc[j]);
                }
                store4(g.out.p + (((int64_t)img * g.ho + oy) * g.wo + ox) * g.out.ld + c0, g.out.plane, acc);
            }
        }
        __syncthreads();                                                       // tile consumed: the next TMA may overwrite it
    }
}

template <int STRIDE>
static int launch_dw_tma(const CUtensorMap& tm, DwArgs& g, cudaStream_t s) {
    using G = DwGeom<STRIDE>;
    g.tiles_x = div_up(g.wo, G::TW);
    g.tiles_y = div_up(g.ho, G::TH);
    g.cblocks = div_up(g.c, 64);
    g.num_tiles = g.n * g.tiles_x * g.tiles_y * g.cblocks;
    const size_t smem = 2 * G::PLANE_BYTES + 64 + 128;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(dw3x3_tma_kernel<STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("dw3x3(tma): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int per_sm = (int)((220u * 1024u) / smem) < 4 ? (int)((220u * 1024u) / smem) : 4;
    const int grid = g.num_tiles < sms * per_sm ? g.num_tiles : sms * per_sm;
    dw3x3_tma_kernel<STRIDE><<<grid, 256, smem, s>>>(tm, g);
    return check_launch("dw3x3(tma)");
}

int dw3x3_tma(Act in, int n, int h, int w, int c, int stride, const float* wgt, const float* bias, int relu6, ActW out,
              cudaStream_t s) {
    DwArgs g{};
    g.n = n; g.h = h; g.w = w; g.c = c;
    g.ho = stride == 1 ? h : (h - 1) / 2 + 1;
    g.wo = stride == 1 ? w : (w - 1) / 2 + 1;
    g.wgt = wgt; g.bias = bias; g.relu6 = relu6; g.out = out;
    const uint64_t dims[5] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n, in.plane ? 2u : 1u};
    const uint64_t row = (uint64_t)in.ld * 2;
    const uint64_t str[4] = {row, row * w, row * w * h, in.plane ? (uint64_t)in.plane * 2 : row * w * h * (uint64_t)n};
    CUtensorMap tm;
    int rc;
    if (stride == 1) {
        const uint32_t box[5] = {64, (uint32_t)DwGeom<1>::IW, (uint32_t)DwGeom<1>::IH, 1, 1};
        rc = tc_encode(&tm, in.p, 5, dims, str, box, "dw input", 0);
        if (rc) return rc;
        return launch_dw_tma<1>(tm, g, s);
    }
    const uint32_t box[5] = {64, (uint32_t)DwGeom<2>::IW, (uint32_t)DwGeom<2>::IH, 1, 1};
    rc = tc_encode(&tm, in.p, 5, dims, str, box, "dw input", 0);
    if (rc) return rc;
    return launch_dw_tma<2>(tm, g, s);
}

}  // namespace uavsal

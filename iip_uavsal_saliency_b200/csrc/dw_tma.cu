// Depthwise 3x3 + BN + ReLU6 (dilation 1, stride 1|2) with TMA-staged, double-buffered input tiles.
//
// A CTA walks a static list of tiles of 16x8 (stride 1) or 8x4 (stride 2) output pixels x 64 channels.  One thread
// issues the cp.async.bulk.tensor load of the haloed input box of tile i+1 while the CTA computes tile i from the other
// buffer; out-of-bounds zero fill provides the convolution padding, so there is no per-thread address arithmetic or
// predication on the input side and every input element crosses L2 -> SM once per tile (1.4x halo factor) instead of
// nine times.  Threads slide a 3x3 register window down their column of the tile (3 shared-memory reads per output),
// apply the folded BN bias + ReLU6 and write both bf16 planes of the output with 128-byte-contiguous stores.
//
// Input formats (FMT): 0 = the arena's split-bf16 planes (two boxes per tile); 1 = plain fp32 rows (the "hidden" tensor
// between an expand GEMM and its depthwise conv is kept in fp32: same bytes, no unpack/re-split work on either side);
// 2 = q16 rows (16-bit fixed point of a ReLU6 output, common.cuh: half the bytes, used for the widest hidden tensors).
#include "tc_common.cuh"

namespace uavsal {

struct DwArgs {
    int n, h, w, c, ho, wo;
    int tiles_x, tiles_y, cblocks, num_tiles;
    const float* wgt;      // [9][c]
    const float* bias;     // [c]
    int relu6;
    ActW out;
    const float* wproj;    // DOT mode: weights of a following 1-output pointwise conv [c]
    float* partial;        // DOT mode: per-(pixel, part) partial dot products [n*ho*wo][parts]
    int parts, cpp;        // work unit = (spatial tile, part); a unit covers cpp consecutive 64-channel blocks (1 unless DOT)
};

template <int STRIDE>
struct DwGeom {
    static constexpr int TW = STRIDE == 1 ? 16 : 8;       // output tile
    static constexpr int TH = STRIDE == 1 ? 8 : 4;
    static constexpr int IW = (TW - 1) * STRIDE + 3;      // haloed input box
    static constexpr int IH = (TH - 1) * STRIDE + 3;
    static constexpr int PIX = IW * IH;
    static constexpr int RGRPS = 256 / (16 * TW);         // row groups: 1 (stride 1) or 2 (stride 2)
    static constexpr int RPT = TH / RGRPS;                // output rows per thread: 8 (stride 1) or 2 (stride 2)
    static constexpr uint32_t TILE_BYTES = PIX * 256;     // 64 channels x (2 bf16 planes | fp32)
    static constexpr uint32_t TILE_BYTES_Q16 = PIX * 128; // 64 channels x uint16
};

// 4 channels from the staged tile -> fp32 (explicit shared-space loads: `tile` is a 32-bit shared address)
// (as two channel pairs: the 9-tap dot products run as packed fma.rn.f32x2 - the same IEEE fma per lane, half the instructions)
template <int FMT>
__device__ __forceinline__ void lds4(uint32_t tile, uint32_t plane_bytes, int pix, int quad, float2 v[2]) {
    if (FMT == 1) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y) : "r"(tile + pix * 256 + quad * 16));
    } else if (FMT == 2) {
        uint2 a;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a.x), "=r"(a.y) : "r"(tile + pix * 128 + quad * 8));
        v[0] = q16_unpack2(a.x); v[1] = q16_unpack2(a.y);
    } else {
        uint2 a, b;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a.x), "=r"(a.y) : "r"(tile + pix * 128 + quad * 8));
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(b.x), "=r"(b.y) : "r"(tile + plane_bytes + pix * 128 + quad * 8));
        float2 t[2];
        unpack2(a.x, v[0].x, v[0].y); unpack2(a.y, v[1].x, v[1].y);
        unpack2(b.x, t[0].x, t[0].y); unpack2(b.y, t[1].x, t[1].y);
        v[0] = __fadd2_rn(v[0], t[0]); v[1] = __fadd2_rn(v[1], t[1]);
    }
}

// thread = (4-channel quad, output column[, row group]); it keeps its 36 folded weights in registers for the tile and
// slides a 3x3x4 register window down RPT output rows.
// DOT: instead of storing the depthwise output, every thread folds its 4 channels into the dot product with `wproj` (the
// dwBlock's project conv when it has ONE output channel: the readout, model.py:372).  A CTA walks the cpp channel blocks of a
// (spatial tile, part) unit back to back and keeps the per-row sums in registers; after the unit's last block the 16 threads
// sharing a pixel reduce by shuffles and one partial per (pixel, part) is written (until round 2: shuffles + a store per channel
// block - the kernel is instruction-issue bound); dot_finish_kernel adds the parts in a fixed order.
template <int STRIDE, int FMT, bool DOT>
__global__ void __launch_bounds__(256, 2) dw3x3_tma_kernel(const __grid_constant__ CUtensorMap tmIn, const DwArgs g) {
    using G = DwGeom<STRIDE>;
    constexpr uint32_t kTile = FMT == 2 ? G::TILE_BYTES_Q16 : G::TILE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kTile);              // [2]

    const int tid = threadIdx.x;
    pdl_trigger();
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
    __syncthreads();
    pdl_wait();                                                                // the input tensor is the previous kernel's output

    // iteration j of this CTA: channel block j % cpp of its (j / cpp)-th unit; unit = (image, tile y, tile x, part), part fastest
    auto decode = [&](int j, int& cblk, int& x0, int& y0, int& img) {
        const int cb = DOT ? j % g.cpp : 0;
        int r = (int)blockIdx.x + (DOT ? j / g.cpp : j) * (int)gridDim.x;
        cblk = (r % g.parts) * g.cpp + cb; r /= g.parts;
        x0 = (r % g.tiles_x) * G::TW; r /= g.tiles_x;
        y0 = (r % g.tiles_y) * G::TH;
        img = r / g.tiles_y;
    };
    auto issue = [&](int j, int b) {                                           // one thread
        int cblk, x0, y0, img;
        decode(j, cblk, x0, y0, img);
        uint8_t* dst = smem + b * kTile;
        fence_async_smem();                                                    // order the buffer's generic reads before the async overwrite
        mbar_expect_tx(bar + b, kTile);
        if (FMT != 0) {
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(smem_u32(dst)), "l"(&tmIn), "r"(smem_u32(bar + b)), "r"(cblk * 64), "r"(x0 * STRIDE - 1), "r"(y0 * STRIDE - 1), "r"(img)
                : "memory");
        } else {
            tma_load_5d(&tmIn, bar + b, dst, cblk * 64, x0 * STRIDE - 1, y0 * STRIDE - 1, img, 0);
            tma_load_5d(&tmIn, bar + b, dst + kTile / 2, cblk * 64, x0 * STRIDE - 1, y0 * STRIDE - 1, img, 1);
        }
    };

    const int quad = tid & 15;                                                 // 4-channel quad inside the 64-channel block
    const int col = (tid >> 4) % G::TW;
    const int rgrp = (tid >> 4) / G::TW;
    const int units_mine = (int)blockIdx.x < g.num_tiles ? (g.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int iters = units_mine * (DOT ? g.cpp : 1);
    if (tid == 0 && iters > 0) issue(0, 0);
    float dsum[DOT ? G::RPT : 1];                                              // DOT: this thread's 4-channel share of the row sums of the unit
    for (int it = 0; it < iters; ++it) {
        const int b = it & 1;
        if (tid == 0 && it + 1 < iters) issue(it + 1, b ^ 1);                  // prefetch the next tile
        int cblk, x0, y0, img;
        decode(it, cblk, x0, y0, img);
        const int cb = DOT ? it % g.cpp : 0;
        const int c0 = cblk * 64 + quad * 4;
        const bool cvalid = c0 < g.c;
        float2 wr[9][2], br[2], wp[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (cvalid) {                                                          // overlaps the TMA flight time (L1/L2 hits)
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(g.wgt + k * g.c + c0));
                wr[k][0] = make_float2(w4.x, w4.y); wr[k][1] = make_float2(w4.z, w4.w);
            }
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + c0));
            br[0] = make_float2(b4.x, b4.y); br[1] = make_float2(b4.z, b4.w);
            if (DOT) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(g.wproj + c0));
                wp[0] = make_float2(p4.x, p4.y); wp[1] = make_float2(p4.z, p4.w);
            }
        } else if (DOT) {
#pragma unroll
            for (int k = 0; k < 9; ++k) { wr[k][0] = wr[k][1] = make_float2(0.f, 0.f); }
            br[0] = br[1] = make_float2(0.f, 0.f);
        }
        if (DOT && cb == 0) {
#pragma unroll
            for (int i = 0; i < (DOT ? G::RPT : 1); ++i) dsum[i] = 0.f;
        }
        mbar_wait(bar + b, (it >> 1) & 1);

        const int ox = x0 + col;
        if (DOT || (cvalid && ox < g.wo)) {                                    // DOT: all lanes stay for the shuffles
            const uint32_t tile = smem_u32(smem) + b * kTile;
            float2 win[3][3][2];
            auto load_row = [&](int slot, int iy) {                            // iy: row inside the input box
#pragma unroll
                for (int d = 0; d < 3; ++d) lds4<FMT>(tile, kTile / 2, iy * G::IW + col * STRIDE + d, quad, win[slot][d]);
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
                float2 a2[2] = {br[0], br[1]};
                const int slots[3] = {s0, s1, s2};
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float2* v = win[slots[ky]][kx];
#pragma unroll
                        for (int j = 0; j < 2; ++j) a2[j] = __ffma2_rn(v[j], wr[ky * 3 + kx][j], a2[j]);
                    }
                float acc[4] = {a2[0].x, a2[0].y, a2[1].x, a2[1].y};
                if (g.relu6) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] = relu6f(acc[j]);
                }
                if (DOT) {
                    dsum[i] += acc[0] * wp[0].x + acc[1] * wp[0].y + acc[2] * wp[1].x + acc[3] * wp[1].y;
                    continue;
                }
                store4(g.out.p + (((int64_t)img * g.ho + oy) * g.wo + ox) * g.out.ld + c0, g.out.plane, acc);
            }
            if (DOT && cb == g.cpp - 1) {                                      // unit complete: one partial per (pixel, part)
                const int part = cblk / g.cpp;
#pragma unroll
                for (int i = 0; i < (DOT ? G::RPT : 1); ++i) {
                    float d = dsum[i];
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);      // the 16 quads of this pixel
                    const int oy = y0 + oyl0 + i;
                    if (quad == 0 && ox < g.wo && oy < g.ho)
                        g.partial[(((int64_t)img * g.ho + oy) * g.wo + ox) * g.parts + part] = d;
                }
            }
        }
        __syncthreads();                                                       // tile consumed: its buffer may be refilled
    }
}

template <int STRIDE, int FMT, bool DOT = false>
static int launch_dw_tma(const CUtensorMap& tm, DwArgs& g, cudaStream_t s) {
    using G = DwGeom<STRIDE>;
    g.tiles_x = div_up(g.wo, G::TW);
    g.tiles_y = div_up(g.ho, G::TH);
    g.cblocks = div_up(g.c, 64);
    // DOT: two parts per spatial tile when there are enough channel blocks (units / CTA slots stays fine-grained enough), else one
    g.parts = DOT ? ((g.cblocks % 2 == 0 && g.cblocks >= 8) ? 2 : 1) : g.cblocks;
    g.cpp = g.cblocks / g.parts;
    g.num_tiles = g.n * g.tiles_x * g.tiles_y * g.parts;                       // work units
    const size_t smem = 2 * (FMT == 2 ? G::TILE_BYTES_Q16 : G::TILE_BYTES) + 64 + 128;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(dw3x3_tma_kernel<STRIDE, FMT, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("dw3x3(tma): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int grid = g.num_tiles < sms * 2 ? g.num_tiles : sms * 2;           // 2 resident CTAs per SM (registers, 2 x 92 KiB smem)
    cudaError_t e = launch_k(dw3x3_tma_kernel<STRIDE, FMT, DOT>, dim3(grid), dim3(256), smem, s, 1, tm, g);
    if (e != cudaSuccess) { set_error("dw3x3(tma): launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("dw3x3(tma)");
}

// in.plane == UAVSAL_PLANE_F32: `in.p` is a float* to fp32 rows [n*h*w][in.ld]; UAVSAL_PLANE_Q16: uint16 fixed-point rows
int dw3x3_tma(Act in, int n, int h, int w, int c, int stride, const float* wgt, const float* bias, int relu6, ActW out,
              cudaStream_t s) {
    DwArgs g{};
    g.n = n; g.h = h; g.w = w; g.c = c;
    g.ho = stride == 1 ? h : (h - 1) / 2 + 1;
    g.wo = stride == 1 ? w : (w - 1) / 2 + 1;
    g.wgt = wgt; g.bias = bias; g.relu6 = relu6; g.out = out;
    CUtensorMap tm;
    int rc;
    if (in.plane == UAVSAL_PLANE_F32 || in.plane == UAVSAL_PLANE_Q16) {
        const bool q16 = in.plane == UAVSAL_PLANE_Q16;
        const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        const uint64_t row = (uint64_t)in.ld * (q16 ? 2 : 4);
        const uint64_t str[3] = {row, row * w, row * w * h};
        if (stride == 1) {
            const uint32_t box[4] = {64, (uint32_t)DwGeom<1>::IW, (uint32_t)DwGeom<1>::IH, 1};
            rc = tc_encode(&tm, in.p, 4, dims, str, box, q16 ? "dw input (q16)" : "dw input (f32)", q16 ? 0 : 2);
            if (rc) return rc;
            return q16 ? launch_dw_tma<1, 2>(tm, g, s) : launch_dw_tma<1, 1>(tm, g, s);
        }
        const uint32_t box[4] = {64, (uint32_t)DwGeom<2>::IW, (uint32_t)DwGeom<2>::IH, 1};
        rc = tc_encode(&tm, in.p, 4, dims, str, box, q16 ? "dw input (q16)" : "dw input (f32)", q16 ? 0 : 2);
        if (rc) return rc;
        return q16 ? launch_dw_tma<2, 2>(tm, g, s) : launch_dw_tma<2, 1>(tm, g, s);
    }
    const uint64_t dims[5] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n, in.plane ? 2u : 1u};
    const uint64_t row = (uint64_t)in.ld * 2;
    const uint64_t str[4] = {row, row * w, row * w * h, in.plane ? (uint64_t)in.plane * 2 : row * w * h * (uint64_t)n};
    if (stride == 1) {
        const uint32_t box[5] = {64, (uint32_t)DwGeom<1>::IW, (uint32_t)DwGeom<1>::IH, 1, 1};
        rc = tc_encode(&tm, in.p, 5, dims, str, box, "dw input", 0);
        if (rc) return rc;
        return launch_dw_tma<1, 0>(tm, g, s);
    }
    const uint32_t box[5] = {64, (uint32_t)DwGeom<2>::IW, (uint32_t)DwGeom<2>::IH, 1, 1};
    rc = tc_encode(&tm, in.p, 5, dims, str, box, "dw input", 0);
    if (rc) return rc;
    return launch_dw_tma<2, 0>(tm, g, s);
}

// ---- small maps, any dilation (the ASPP branches: 12x20 maps, dilation 6 / 12 / 18, model.py:117-128) ------------------------
// A CTA stages the WHOLE image of one 64-channel block by TMA (double-buffered across work items), so every input byte crosses
// HBM once whatever the dilation, and out-of-image taps are skipped.  thread = (4-channel quad, every 16th pixel).
struct DwImgArgs {
    int n, h, w, c, dil, cblocks, num_items;
    const float* wgt;      // [9][c]
    const float* bias;     // [c]
    int relu6;
    ActW out;
};

template <int FMT>
__global__ void __launch_bounds__(256) dw3x3_img_kernel(const __grid_constant__ CUtensorMap tmIn, const DwImgArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int hw = g.h * g.w;
    const uint32_t tile_bytes = (uint32_t)hw * (FMT == 2 ? 128u : 256u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * tile_bytes);         // [2]
    const int tid = threadIdx.x;
    pdl_trigger();
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
    __syncthreads();
    pdl_wait();
    auto issue = [&](int t, int b) {                                           // one thread; item t = (image, channel block), block fastest
        const int cblk = t % g.cblocks, img = t / g.cblocks;
        uint8_t* dst = smem + b * tile_bytes;
        fence_async_smem();
        mbar_expect_tx(bar + b, tile_bytes);
        if (FMT != 0) {
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(smem_u32(dst)), "l"(&tmIn), "r"(smem_u32(bar + b)), "r"(cblk * 64), "r"(0), "r"(0), "r"(img)
                : "memory");
        } else {
            tma_load_5d(&tmIn, bar + b, dst, cblk * 64, 0, 0, img, 0);
            tma_load_5d(&tmIn, bar + b, dst + tile_bytes / 2, cblk * 64, 0, 0, img, 1);
        }
    };
    const int quad = tid & 15, pl = tid >> 4;
    if (tid == 0 && (int)blockIdx.x < g.num_items) issue(blockIdx.x, 0);
    int it = 0;
    for (int t = blockIdx.x; t < g.num_items; t += gridDim.x, ++it) {
        const int b = it & 1;
        if (tid == 0 && t + (int)gridDim.x < g.num_items) issue(t + gridDim.x, b ^ 1);
        const int cblk = t % g.cblocks, img = t / g.cblocks;
        const int c0 = cblk * 64 + quad * 4;
        const bool cvalid = c0 < g.c;
        float2 wr[9][2], br[2];
        if (cvalid) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(g.wgt + k * g.c + c0));
                wr[k][0] = make_float2(w4.x, w4.y); wr[k][1] = make_float2(w4.z, w4.w);
            }
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + c0));
            br[0] = make_float2(b4.x, b4.y); br[1] = make_float2(b4.z, b4.w);
        }
        mbar_wait(bar + b, (it >> 1) & 1);
        if (cvalid) {
            const uint32_t tile = smem_u32(smem) + b * tile_bytes;
            for (int p = pl; p < hw; p += 16) {
                const int y = p / g.w, x = p - y * g.w;
                float2 a2[2] = {br[0], br[1]};
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int yy = y + (ky - 1) * g.dil;
                    if (yy < 0 || yy >= g.h) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int xx = x + (kx - 1) * g.dil;
                        if (xx < 0 || xx >= g.w) continue;
                        float2 v[2];
                        lds4<FMT>(tile, tile_bytes / 2, yy * g.w + xx, quad, v);
                        a2[0] = __ffma2_rn(v[0], wr[ky * 3 + kx][0], a2[0]);
                        a2[1] = __ffma2_rn(v[1], wr[ky * 3 + kx][1], a2[1]);
                    }
                }
                float acc[4] = {a2[0].x, a2[0].y, a2[1].x, a2[1].y};
                if (g.relu6) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] = relu6f(acc[j]);
                }
                store4(g.out.p + ((int64_t)img * hw + p) * g.out.ld + c0, g.out.plane, acc);
            }
        }
        __syncthreads();                                                       // image consumed: its buffer may be refilled
    }
}

template <int FMT>
static int launch_dw_img(const CUtensorMap& tm, DwImgArgs& g, cudaStream_t s) {
    const size_t smem = 2 * (size_t)g.h * g.w * (FMT == 2 ? 128 : 256) + 64 + 128;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(dw3x3_img_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("dw3x3(img): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr = smem;
    }
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    int per_sm = (int)((220 * 1024) / smem);
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    const int grid = g.num_items < sms * per_sm ? g.num_items : sms * per_sm;
    cudaError_t e = launch_k(dw3x3_img_kernel<FMT>, dim3(grid), dim3(256), smem, s, 1, tm, g);
    if (e != cudaSuccess) { set_error("dw3x3(img): launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("dw3x3(img)");
}

// whole-image staging fits: 2 buffers of h*w pixels x 64 channels (256 B per pixel split-bf16 / fp32, 128 B q16)
bool dw3x3_img_fits(int64_t in_plane, int h, int w) {
    const int64_t px = (int64_t)h * w;
    return w <= 256 && h <= 256 && 2 * px * (in_plane == UAVSAL_PLANE_Q16 ? 128 : 256) + 192 <= 200 * 1024;
}

// stride 1, dilation >= 1 (padding = dilation) on small maps; in.plane as dw3x3_tma
int dw3x3_img(Act in, int n, int h, int w, int c, int dil, const float* wgt, const float* bias, int relu6, ActW out, cudaStream_t s) {
    DwImgArgs g{};
    g.n = n; g.h = h; g.w = w; g.c = c; g.dil = dil; g.cblocks = div_up(c, 64); g.num_items = n * g.cblocks;
    g.wgt = wgt; g.bias = bias; g.relu6 = relu6; g.out = out;
    CUtensorMap tm;
    if (in.plane == UAVSAL_PLANE_F32 || in.plane == UAVSAL_PLANE_Q16) {
        const bool q16 = in.plane == UAVSAL_PLANE_Q16;
        const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        const uint64_t row = (uint64_t)in.ld * (q16 ? 2 : 4);
        const uint64_t str[3] = {row, row * w, row * w * h};
        const uint32_t box[4] = {64, (uint32_t)w, (uint32_t)h, 1};
        int rc = tc_encode(&tm, in.p, 4, dims, str, box, q16 ? "dw(img) input (q16)" : "dw(img) input (f32)", q16 ? 0 : 2);
        if (rc) return rc;
        return q16 ? launch_dw_img<2>(tm, g, s) : launch_dw_img<1>(tm, g, s);
    }
    set_error("dw3x3(img): plain-row input (fp32 / q16) only");
    return UAVSAL_ENOTSUP;
}

// readout tail: sum the channel-block partials of a pixel in a fixed order, add the folded BN bias, sigmoid (model.py:373)
__global__ void __launch_bounds__(256) dot_finish_kernel(const float* __restrict__ partial, int64_t rows, int cblocks, float bias,
                                                         float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float s = 0.f;
    for (int k = 0; k < cblocks; ++k) s += partial[r * cblocks + k];
    out[r] = 1.f / (1.f + expf(-(s + bias)));
}

int dw3x3_dot_tma(const void* in, bool q16, int in_ld, int n, int h, int w, int c, const float* wgt, const float* bias, const float* wproj,
                  float bias_proj, float* partial, float* out, cudaStream_t s) {
    DwArgs g{};
    g.n = n; g.h = h; g.w = w; g.c = c; g.ho = h; g.wo = w;
    g.wgt = wgt; g.bias = bias; g.relu6 = 1; g.wproj = wproj; g.partial = partial;
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t row = (uint64_t)in_ld * (q16 ? 2 : 4);
    const uint64_t str[3] = {row, row * w, row * w * h};
    const uint32_t box[4] = {64, (uint32_t)DwGeom<1>::IW, (uint32_t)DwGeom<1>::IH, 1};
    CUtensorMap tm;
    int rc = tc_encode(&tm, in, 4, dims, str, box, q16 ? "dw_dot input (q16)" : "dw_dot input (f32)", q16 ? 0 : 2);
    if (rc) return rc;
    rc = q16 ? launch_dw_tma<1, 2, true>(tm, g, s) : launch_dw_tma<1, 1, true>(tm, g, s);
    if (rc) return rc;
    const int64_t rows = (int64_t)n * h * w;
    cudaError_t e = launch_k(dot_finish_kernel, dim3(div_up(rows, 256)), dim3(256), 0, s, 1, (const float*)partial, rows, g.parts, bias_proj, out);
    if (e != cudaSuccess) { set_error("dw_dot: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("dw_dot(finish)");
}

}  // namespace uavsal

// Fused 1x1 expand (+BN+ReLU6) -> depthwise 3x3 (+BN+ReLU6) for dwBlocks with few input channels (cin <= 32):
// model.py:90-92 / torchvision InvertedResidual conv[0], conv[1].
//
// At 180x320 .. 45x80 the 6x-expanded hidden tensor of MobileNetV2's first blocks is the whole cost of the block (a
// 16 -> 96 expand at 180x320 writes 442 MB that the depthwise conv immediately re-reads).  This kernel never materialises
// it: a CTA loads the haloed INPUT tile (cin channels, split-bf16), expands 64 hidden channels of it on the tensor cores
// into shared memory (fp32, zero outside the image = the depthwise conv's padding), and runs the sliding-window depthwise
// stage of dw_tma.cu straight from that buffer.  HBM traffic per block drops from (1 + 6 + 6 + 6s) x to (1 + 6s) x
// activations (s = 1 or 1/4 for stride 2); the halo recompute (1.4x of a K <= 64 GEMM) is free.
//
// The expand is a [<=304 pixels] x [64 channels] x [K = cin <= 32] product per tile: too small and too irregular (haloed
// pixel rows, per-tile weight block) for a TMA/tcgen05 pipeline to pay off, and the kernel is HBM/latency-bound, so it
// uses warp-level mma.sync (m16n8k16, bf16 hi/lo 3-term split, fp32 accumulate) with ldmatrix operands.
#include "tc_common.cuh"

namespace uavsal {

struct ExpDwArgs {
    Act x;                 // input NHWC, split-bf16
    int cin, kp;           // input channels (multiple of 8) and K padded to a multiple of 16
    const uint16_t* w1;    // [2 planes][cblocks*64][kp] bf16, K-major, zero padded
    const float* b1;       // [cblocks*64]
    const float* wd;       // [9][hidden]
    const float* bd;       // [hidden]
    int n, h, w, hidden, ho, wo;
    int tiles_x, tiles_y, cblocks, num_tiles;
    int cpg, cgroups;      // channel blocks per work item (they share one staged input tile) and groups per spatial tile
    ActW out;              // depthwise output, split-bf16
};

template <int STRIDE>
struct EdGeom {
    static constexpr int TW = STRIDE == 1 ? 16 : 8;       // output tile
    static constexpr int TH = STRIDE == 1 ? 8 : 4;
    static constexpr int IW = (TW - 1) * STRIDE + 3;      // haloed input box
    static constexpr int IH = (TH - 1) * STRIDE + 3;
    static constexpr int PIX = IW * IH;                   // 180 | 153
    static constexpr int MT = (PIX + 15) / 16;            // m16 tiles: 12 | 10
    static constexpr int ROWS = MT * 16;
    static constexpr int RGRPS = 256 / (16 * TW);
    static constexpr int RPT = TH / RGRPS;
    static constexpr int HPITCH = 272;                    // bytes per pixel of the hidden tile: 64 fp32 + 16 (bank spread)
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& a, uint32_t& b) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int STRIDE>
__global__ void __launch_bounds__(256, 2) expdw_kernel(const ExpDwArgs g) {
    using G = EdGeom<STRIDE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 127) & ~127u;
    const int kpitch = (g.kp + 8) * 2;                                        // bytes per operand row (16 B pad: conflict-free ldmatrix)
    const uint32_t xs_hi = sbase;                                             // [ROWS][kp+8] bf16
    const uint32_t xs_lo = xs_hi + G::ROWS * kpitch;
    const uint32_t ws_hi = xs_lo + G::ROWS * kpitch;                          // [64][kp+8] bf16
    const uint32_t ws_lo = ws_hi + 64 * kpitch;
    const uint32_t hid = (ws_lo + 64 * kpitch + 15) & ~15u;                   // [PIX][HPITCH] fp32

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cpp = g.cin >> 3;                                               // 16-byte chunks per pixel per plane
    const int kchunks = g.kp >> 3;
    // staging map: 16 consecutive threads serve one row (pixel / weight row): chunk lc of plane lp
    const int lrow = tid >> 4, lp = tid & 1, lc = (tid & 15) >> 1;
    pdl_trigger();
    // zero the K padding columns once (they are multiplied by zero weights, but must not hold NaN bit patterns)
    if (lc >= cpp && lc < kchunks)
        for (int r = lrow; r < G::ROWS; r += 16) sts128((lp ? xs_lo : xs_hi) + r * kpitch + lc * 16, 0, 0, 0, 0);

    const int quad = tid & 15;
    const int col = (tid >> 4) % G::TW;
    const int rgrp = (tid >> 4) / G::TW;
    pdl_wait();
    for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
        int r = t;
        const int cg = r % g.cgroups; r /= g.cgroups;
        const int x0 = (r % g.tiles_x) * G::TW; r /= g.tiles_x;
        const int y0 = (r % g.tiles_y) * G::TH;
        const int img = r / g.tiles_y;
        const int gx0 = x0 * STRIDE - 1, gy0 = y0 * STRIDE - 1;

        // ---- stage 1: haloed input tile (both planes) -> shared memory, once for all channel blocks of this work item ----
        // chunk i -> (pixel, plane, 16-byte chunk); every thread issues all of its loads before the first (ordered, asm) store,
        // so one L2 round trip covers the tile
        {
            const int cpp2 = 2 * cpp, total = G::PIX * cpp2;
            const uint32_t inv = (1u << 20) / (uint32_t)cpp2 + 1;             // exact i / cpp2 for i < 8192
            constexpr int kMaxU = (G::PIX * 8 + 255) / 256;                   // cin <= 32: at most 8 chunks per pixel
            uint4 v[kMaxU];
            uint32_t dst[kMaxU];
#pragma unroll
            for (int u = 0; u < kMaxU; ++u) {
                const int i = tid + u * 256;
                dst[u] = 0;
                v[u] = make_uint4(0, 0, 0, 0);
                if (i < total) {
                    const int p = (int)(((uint32_t)i * inv) >> 20), rem = i - p * cpp2;
                    const int pl = rem & 1, c = rem >> 1;
                    const int iy = p / G::IW, ix = p - iy * G::IW;
                    const int gy = gy0 + iy, gx = gx0 + ix;
                    if (gy >= 0 && gy < g.h && gx >= 0 && gx < g.w)
                        v[u] = __ldg(reinterpret_cast<const uint4*>(g.x.p + (pl ? g.x.plane : 0) + c * 8 + (((int64_t)img * g.h + gy) * g.w + gx) * g.x.ld));
                    dst[u] = (pl ? xs_lo : xs_hi) + p * kpitch + c * 16;
                }
            }
#pragma unroll
            for (int u = 0; u < kMaxU; ++u)
                if (dst[u]) sts128(dst[u], v[u].x, v[u].y, v[u].z, v[u].w);
        }

        const int cb_end = min(g.cblocks, (cg + 1) * g.cpg);
        for (int cblk = cg * g.cpg; cblk < cb_end; ++cblk) {
            // expand weights of this channel block -> shared memory; this lane's 16 expand biases -> registers
            if (lc < kchunks) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = __ldg(reinterpret_cast<const uint4*>(g.w1 + ((int64_t)lp * g.cblocks * 64 + cblk * 64 + u * 16 + lrow) * g.kp + lc * 8));
#pragma unroll
                for (int u = 0; u < 4; ++u) sts128((lp ? ws_lo : ws_hi) + (u * 16 + lrow) * kpitch + lc * 16, v[u].x, v[u].y, v[u].z, v[u].w);
            }
            float2 eb[8];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) eb[nt] = __ldg(reinterpret_cast<const float2*>(g.b1 + cblk * 64 + nt * 8 + 2 * (lane & 3)));
            __syncthreads();                                                   // x tile + weights staged; previous block's depthwise stage done

            // ---- stage 2: hidden[pix][64] = ReLU6(x . W1^T + b1), zero outside the image (tensor cores, 3-term split) ----
            // task = (m16 tile, 32-channel half): 2*MT tasks round-robin over the 8 warps; the three split products of the
            // four n8 tiles are issued term-major so that consecutive HMMAs never depend on each other
            for (int task = warp; task < 2 * G::MT; task += 8) {
                const int mt = task >> 1, nh = task & 1;
                float acc[4][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
                const uint32_t a_off = (mt * 16 + (lane & 15)) * kpitch + (lane >> 4) * 16;
                const uint32_t b_off = (nh * 32 + (lane & 7)) * kpitch + ((lane >> 3) & 1) * 16;
                for (int ks = 0; ks < (g.kp >> 4); ++ks) {
                    uint32_t ah[4], al[4], bh[4][2], bl[4][2];
                    ldsm_x4(xs_hi + a_off + ks * 32, ah[0], ah[1], ah[2], ah[3]);
                    ldsm_x4(xs_lo + a_off + ks * 32, al[0], al[1], al[2], al[3]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        ldsm_x2(ws_hi + b_off + nt * 8 * kpitch + ks * 32, bh[nt][0], bh[nt][1]);
                        ldsm_x2(ws_lo + b_off + nt * 8 * kpitch + ks * 32, bl[nt][0], bl[nt][1]);
                    }
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], ah, bh[nt][0], bh[nt][1]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], ah, bl[nt][0], bl[nt][1]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], al, bh[nt][0], bh[nt][1]);
                }
                // accumulator fragment: rows lane/4 and lane/4 + 8 of the m-tile, columns nh*32 + nt*8 + 2*(lane%4) + {0,1}
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int p = mt * 16 + (lane >> 2) + half * 8;
                    if (p < G::PIX) {
                        const int iy = p / G::IW, ix = p - iy * G::IW;
                        const int gy = gy0 + iy, gx = gx0 + ix;
                        const bool inb = gy >= 0 && gy < g.h && gx >= 0 && gx < g.w;
                        const uint32_t dst = hid + p * G::HPITCH + nh * 128 + 8 * (lane & 3);
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const float2 b2 = nh ? eb[4 + nt] : eb[nt];
                            float v0 = relu6f(acc[nt][half * 2 + 0] + b2.x), v1 = relu6f(acc[nt][half * 2 + 1] + b2.y);
                            if (!inb) { v0 = 0.f; v1 = 0.f; }
                            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(dst + nt * 32), "f"(v0), "f"(v1) : "memory");
                        }
                    }
                }
            }
            // depthwise weights of this thread's 4 channels (global/L1; loaded here to keep them out of the MMA stage's registers)
            const int c0 = cblk * 64 + quad * 4;
            const bool cvalid = c0 < g.hidden;
            float wr[9][4], br[4];
            if (cvalid) {
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(g.wd + k * g.hidden + c0));
                    wr[k][0] = w4.x; wr[k][1] = w4.y; wr[k][2] = w4.z; wr[k][3] = w4.w;
                }
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bd + c0));
                br[0] = b4.x; br[1] = b4.y; br[2] = b4.z; br[3] = b4.w;
            }
            __syncthreads();                                                   // hidden tile complete

            // ---- stage 3: depthwise 3x3 + bias + ReLU6 from the hidden tile (sliding 3x3x4 register window, as dw_tma.cu) ----
            const int ox = x0 + col;
            if (cvalid && ox < g.wo) {
                float win[3][3][4];
                auto load_row = [&](int slot, int iy) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        float* v = win[slot][d];
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                                     : "r"(hid + (iy * G::IW + col * STRIDE + d) * G::HPITCH + quad * 16));
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
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] = relu6f(acc[j]);
                    store4(g.out.p + (((int64_t)img * g.ho + oy) * g.wo + ox) * g.out.ld + c0, g.out.plane, acc);
                }
            }
            // no barrier here: the next block's weight stores touch only ws (idle since stage 2) and its first barrier orders
            // this depthwise stage before the hidden tile is overwritten
        }
        __syncthreads();                                                       // the x tile may be refilled
    }
}

template <int STRIDE>
static int launch_expdw(ExpDwArgs& g, cudaStream_t s) {
    using G = EdGeom<STRIDE>;
    g.tiles_x = div_up(g.wo, G::TW);
    g.tiles_y = div_up(g.ho, G::TH);
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    // as many channel blocks per work item as keeps >= 4 work items per resident CTA (input tile reuse vs load balance)
    const int spatial = g.n * g.tiles_x * g.tiles_y;
    g.cpg = g.cblocks;
    while (g.cpg > 1 && (int64_t)spatial * div_up(g.cblocks, g.cpg) < 8LL * sms) --g.cpg;
    g.cgroups = div_up(g.cblocks, g.cpg);
    g.num_tiles = spatial * g.cgroups;
    const int kpitch = (g.kp + 8) * 2;
    const size_t smem = 2 * (size_t)G::ROWS * kpitch + 2 * 64 * (size_t)kpitch + (size_t)G::PIX * G::HPITCH + 256;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(expdw_kernel<STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("expand_dw3x3: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr = smem;
    }
    const int grid = g.num_tiles < sms * 2 ? g.num_tiles : sms * 2;
    cudaError_t e = launch_k(expdw_kernel<STRIDE>, dim3(grid), dim3(256), smem, s, 1, g);
    if (e != cudaSuccess) { set_error("expand_dw3x3: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("expand_dw3x3");
}

}  // namespace uavsal

using namespace uavsal;

extern "C" int uavsal_expand_dw3x3(const uint16_t* x, int64_t x_plane, int x_ld, int n, int h, int w, int cin, const uint16_t* w1, int kp,
                                   const float* b1, int hidden, int stride, const float* wd, const float* bd, uint16_t* out,
                                   int64_t out_plane, int out_ld, void* stream) {
    const bool ok16 = x && w1 && b1 && wd && bd && out && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w1) |
                                                            reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(wd) |
                                                            reinterpret_cast<uintptr_t>(bd) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    UAVSAL_REQUIRE(ok16 && x_plane > 0 && out_plane > 0 && x_plane % 8 == 0 && out_plane % 8 == 0 && x_ld % 8 == 0 && out_ld % 8 == 0 &&
                       n > 0 && h > 0 && w > 0 && cin > 0 && cin % 8 == 0 && hidden > 0 && hidden % 8 == 0 && x_ld >= cin && out_ld >= hidden,
                   UAVSAL_EINVAL, "expand_dw3x3: bad arguments (cin=%d hidden=%d)", cin, hidden);
    UAVSAL_REQUIRE(cin <= 32 && kp % 16 == 0 && kp >= cin && kp <= 32 && (stride == 1 || stride == 2), UAVSAL_ENOTSUP,
                   "expand_dw3x3: cin %d / kp %d / stride %d unsupported (cin <= 32, stride 1|2)", cin, kp, stride);
    ExpDwArgs g{};
    g.x = Act{x, x_plane, x_ld};
    g.cin = cin; g.kp = kp; g.w1 = w1; g.b1 = b1; g.wd = wd; g.bd = bd;
    g.n = n; g.h = h; g.w = w; g.hidden = hidden;
    g.ho = stride == 1 ? h : (h - 1) / 2 + 1;
    g.wo = stride == 1 ? w : (w - 1) / 2 + 1;
    g.cblocks = div_up(hidden, 64);
    g.out = ActW{out, out_plane, out_ld};
    return stride == 1 ? launch_expdw<1>(g, (cudaStream_t)stream) : launch_expdw<2>(g, (cudaStream_t)stream);
}

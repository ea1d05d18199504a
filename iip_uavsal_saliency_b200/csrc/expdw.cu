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
// The expand is a [<=192 pixels] x [64 channels] x [K = cin <= 32] product per tile: too small and too irregular (haloed
// pixel rows, per-tile weight block) for a TMA/tcgen05 pipeline to pay off, so it uses warp-level mma.sync (m16n8k16,
// bf16 hi/lo 3-term split, fp32 accumulate, bias as the initial accumulator) with ldmatrix operands.  The kernel is
// instruction/latency-bound (ncu: ~100 thread instructions per output value in the first version, FFMA 9 of them), so
// the structure is about instruction count: weights resident in shared memory for the whole (persistent) CTA, the next
// tile's input fetched by cp.async under the depthwise stage, ReLU6 + out-of-image mask as ONE mul.sat per hidden value
// (the factor 6 moves into the depthwise taps), warp-uniform control flow (no WARPSYNC around ldmatrix/mma).
// Measured (20 frames, 180x320, 16 -> 96, stride 2): 275 us -> 180 us.
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
    static constexpr int MT = (PIX + 31) / 32 * 2;        // m16 tiles (an even number: two per round): 12 | 10
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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {      // src_bytes 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Persistent CTA, work item = (spatial tile, group of 64-channel blocks).  All weights of the block (expand hi/lo, both
// bias vectors, depthwise taps) stay in shared memory for the whole kernel; the haloed input tile of the NEXT work item is
// fetched with cp.async into the (single) x buffer while the depthwise stage of the last channel block runs - the buffer
// is dead from the moment that block's expand stage has finished.  Two barriers per channel block.
template <int STRIDE>
__global__ void __launch_bounds__(256, 2) expdw_kernel(const ExpDwArgs g) {
    using G = EdGeom<STRIDE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t sbase = (sraw + 127) & ~127u;
    uint8_t* const sgen = smem_raw + (sbase - sraw);                          // generic-space view of sbase
    const int kpitch = (g.kp + 8) * 2;                                        // bytes per operand row (16 B pad: conflict-free ldmatrix)
    const int cb64 = g.cblocks * 64;
    const uint32_t xs_hi = sbase;                                             // [ROWS][kp+8] bf16
    const uint32_t xs_lo = xs_hi + G::ROWS * kpitch;
    const uint32_t ws_hi = xs_lo + G::ROWS * kpitch;                          // [cb64][kp+8] bf16, every channel block
    const uint32_t ws_lo = ws_hi + cb64 * kpitch;
    const uint32_t b1s = ws_lo + cb64 * kpitch;                               // [cb64] expand bias
    const uint32_t wds = b1s + cb64 * 4;                                      // [9][cb64] depthwise taps
    const uint32_t bds = wds + 9 * cb64 * 4;                                  // [cb64] depthwise bias
    const uint32_t hid = bds + cb64 * 4;                                      // [ROWS][HPITCH] fp32

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);                   // tells the compiler the index is warp-uniform
    const int cpp = g.cin >> 3;                                               // 16-byte chunks per pixel per plane
    const int kchunks = g.kp >> 3;
    pdl_trigger();
    // ---- once: weights -> shared memory (constants: no need to wait for the producer kernel), K padding of the x tile = 0 ----
    for (int i = tid; i < 2 * cb64 * kchunks; i += 256) {
        const int c = i % kchunks, r = (i / kchunks) % cb64, pl = i / (kchunks * cb64);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(g.w1 + ((int64_t)pl * cb64 + r) * g.kp + c * 8));
        *reinterpret_cast<uint4*>(sgen + ((pl ? ws_lo : ws_hi) - sbase) + r * kpitch + c * 16) = v;
    }
    for (int i = tid; i < cb64; i += 256) {
        const bool v = i < g.hidden;
        reinterpret_cast<float*>(sgen + (b1s - sbase))[i] = __ldg(g.b1 + i);
        reinterpret_cast<float*>(sgen + (bds - sbase))[i] = v ? __ldg(g.bd + i) : 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) reinterpret_cast<float*>(sgen + (wds - sbase))[k * cb64 + i] = v ? 6.f * __ldg(g.wd + k * g.hidden + i) : 0.f;   // x6: see the expand stage
    }
    for (int i = tid; i < 2 * G::ROWS * (kchunks - cpp); i += 256) {
        const int c = cpp + i % (kchunks - cpp), r = (i / (kchunks - cpp)) % G::ROWS, pl = i / ((kchunks - cpp) * G::ROWS);
        *reinterpret_cast<uint4*>(sgen + ((pl ? xs_lo : xs_hi) - sbase) + r * kpitch + c * 16) = make_uint4(0, 0, 0, 0);
    }

    auto decode = [&](int t, int& img, int& y0, int& x0, int& cg) {
        int r = t;
        cg = r % g.cgroups; r /= g.cgroups;
        x0 = (r % g.tiles_x) * G::TW; r /= g.tiles_x;
        y0 = (r % g.tiles_y) * G::TH;
        img = r / g.tiles_y;
    };
    // haloed input tile (both planes) -> shared memory: 8 consecutive threads serve one pixel (slot = plane x 16-byte chunk)
    auto issue_x = [&](int img, int y0, int x0) {
        const int slot = tid & 7;
        const int pl = slot >= cpp ? 1 : 0, c = slot - (pl ? cpp : 0);
        const int gx0 = x0 * STRIDE - 1, gy0 = y0 * STRIDE - 1;
        const uint16_t* base = g.x.p + (pl ? g.x.plane : 0) + c * 8;
        const uint32_t dst0 = (pl ? xs_lo : xs_hi) + c * 16;
        if (slot < 2 * cpp) {
#pragma unroll
            for (int u = 0; u < (G::PIX + 31) / 32; ++u) {
                const int p = u * 32 + (tid >> 3);
                if (p < G::PIX) {
                    const int iy = p / G::IW, ix = p - iy * G::IW;
                    const int gy = gy0 + iy, gx = gx0 + ix;
                    const bool inb = gy >= 0 && gy < g.h && gx >= 0 && gx < g.w;
                    const uint16_t* src = inb ? base + (((int64_t)img * g.h + gy) * g.w + gx) * g.x.ld : g.x.p;
                    cp_async16(dst0 + p * kpitch, src, inb ? 16u : 0u);
                }
            }
        }
        cp_async_commit();
    };

    int img, y0, x0, cg;
    decode(blockIdx.x, img, y0, x0, cg);
    pdl_wait();
    issue_x(img, y0, x0);
    for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        int nimg = 0, ny0 = 0, nx0 = 0, ncg = 0;
        if (tn < g.num_tiles) decode(tn, nimg, ny0, nx0, ncg);
        const int gx0 = x0 * STRIDE - 1, gy0 = y0 * STRIDE - 1;
        const bool border = gx0 < 0 || gy0 < 0 || gx0 + G::IW > g.w || gy0 + G::IH > g.h;   // some haloed pixels lie outside the image

        const int cb_begin = cg * g.cpg, cb_end = min(g.cblocks, (cg + 1) * g.cpg);
        for (int cblk = cb_begin; cblk < cb_end; ++cblk) {
            if (cblk == cb_begin) cp_async_wait_all();
            __syncthreads();                                                   // x tile staged; previous block's depthwise stage done

            // ---- expand: hidden[pix][64] = ReLU6(x . W1^T + b1) / 6, zero outside the image (tensor cores, 3-term split) ----
            // warp -> one 16-channel quarter of the block (B fragments and bias stay in registers) x every (8/nqd)-th pair of m16
            // tiles; a partially filled last block spreads its valid quarters over all warps.  ReLU6(v)/6 = sat(v * 1/6) is ONE
            // instruction (mul.sat) and the out-of-image mask rides on its multiplier; the depthwise taps carry the factor 6.
            {
                const int nq_valid = (min(64, g.hidden - cblk * 64) + 15) >> 4;
                const int sh = nq_valid <= 1 ? 0 : nq_valid <= 2 ? 1 : 2;     // log2 of the quarters the warps are dealt over
                const int nq = warp & ((1 << sh) - 1);
                if (nq < nq_valid) {
                    uint32_t bh[2][4], bl[2][4];
                    const uint32_t b_off = (cblk * 64 + nq * 16 + (lane >> 4) * 8 + (lane & 7)) * kpitch + ((lane >> 3) & 1) * 16;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        if (ks < (g.kp >> 4)) {
                            ldsm_x4(ws_hi + b_off + ks * 32, bh[ks][0], bh[ks][1], bh[ks][2], bh[ks][3]);
                            ldsm_x4(ws_lo + b_off + ks * 32, bl[ks][0], bl[ks][1], bl[ks][2], bl[ks][3]);
                        }
                    float2 eb[2];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(eb[nt].x), "=f"(eb[nt].y)
                                     : "r"(b1s + (cblk * 64 + nq * 16 + nt * 8 + 2 * (lane & 3)) * 4));
                    const uint32_t a_lane = (lane & 15) * kpitch + (lane >> 4) * 16;
                    const uint32_t h_lane = hid + (lane >> 2) * G::HPITCH + nq * 64 + 8 * (lane & 3);
                    for (int mb = warp >> sh; mb < G::MT / 2; mb += 8 >> sh) {
                        // two adjacent m-tiles per round: four independent accumulator chains, bias as the initial accumulator
                        float acc[2][2][4];
#pragma unroll
                        for (int m2 = 0; m2 < 2; ++m2)
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) { acc[m2][nt][0] = acc[m2][nt][2] = eb[nt].x; acc[m2][nt][1] = acc[m2][nt][3] = eb[nt].y; }
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            if (ks < (g.kp >> 4)) {
                                uint32_t ah[2][4], al[2][4];
#pragma unroll
                                for (int m2 = 0; m2 < 2; ++m2) {
                                    const uint32_t a_off = (mb * 2 + m2) * 16 * kpitch + a_lane + ks * 32;
                                    ldsm_x4(xs_hi + a_off, ah[m2][0], ah[m2][1], ah[m2][2], ah[m2][3]);
                                    ldsm_x4(xs_lo + a_off, al[m2][0], al[m2][1], al[m2][2], al[m2][3]);
                                }
#pragma unroll
                                for (int term = 0; term < 3; ++term)
#pragma unroll
                                    for (int m2 = 0; m2 < 2; ++m2)
#pragma unroll
                                        for (int nt = 0; nt < 2; ++nt) {
                                            const uint32_t* bb = term == 1 ? bl[ks] : bh[ks];
                                            mma_bf16(acc[m2][nt], term == 2 ? al[m2] : ah[m2], bb[2 * nt], bb[2 * nt + 1]);
                                        }
                            }
                        // accumulator fragment: rows lane/4 and lane/4 + 8 of the m-tile, columns nq*16 + nt*8 + 2*(lane%4) + {0,1};
                        // rows >= PIX of the last tile land in the tile's padding rows
#pragma unroll
                        for (int m2 = 0; m2 < 2; ++m2)
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                const int pr = (mb * 2 + m2) * 16 + half * 8;
                                float mul = 1.f / 6.f;
                                if (border) {
                                    const int p = pr + (lane >> 2);
                                    const int iy = p / G::IW, ix = p - iy * G::IW;
                                    const int gy = gy0 + iy, gx = gx0 + ix;
                                    if (!(gy >= 0 && gy < g.h && gx >= 0 && gx < g.w)) mul = 0.f;
                                }
#pragma unroll
                                for (int nt = 0; nt < 2; ++nt) {
                                    float v0, v1;
                                    asm("mul.sat.f32 %0, %1, %2;" : "=f"(v0) : "f"(acc[m2][nt][half * 2 + 0]), "f"(mul));
                                    asm("mul.sat.f32 %0, %1, %2;" : "=f"(v1) : "f"(acc[m2][nt][half * 2 + 1]), "f"(mul));
                                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(h_lane + pr * G::HPITCH + nt * 32), "f"(v0), "f"(v1) : "memory");
                                }
                            }
                    }
                }
            }
            // depthwise stage mapping: thread = (4-channel quad, output column, row group).  A last block with at most 32 valid channels
            // (hidden = 96: features.2) has only 8 quads: the threads are dealt over twice as many row groups instead of leaving half
            // of them idle, and the stage takes half as long
            const bool half = g.hidden - cblk * 64 <= 32;                      // warp-uniform
            const int quad = half ? (tid & 7) : (tid & 15);
            const int pcol = half ? (tid >> 3) : (tid >> 4);
            const int col = pcol % G::TW, rgrp = pcol / G::TW;
            const int rpt = half ? G::RPT / 2 : G::RPT;
            // depthwise taps / bias of this thread's 4 channels
            const int c0 = cblk * 64 + quad * 4;
            const bool cvalid = c0 < g.hidden;
            // (channel pairs: the 9-tap dot products run as packed fma.rn.f32x2 - the same IEEE fma per lane, half the instructions)
            float2 wr[9][2], br[2];
#pragma unroll
            for (int k = 0; k < 9; ++k)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(wr[k][0].x), "=f"(wr[k][0].y), "=f"(wr[k][1].x), "=f"(wr[k][1].y)
                             : "r"(wds + (k * cb64 + c0) * 4));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(br[0].x), "=f"(br[0].y), "=f"(br[1].x), "=f"(br[1].y) : "r"(bds + c0 * 4));
            __syncthreads();                                                   // hidden tile complete; the x tile is dead
            if (cblk == cb_end - 1 && tn < g.num_tiles) issue_x(nimg, ny0, nx0);

            // ---- depthwise 3x3 + bias + ReLU6 from the hidden tile (sliding 3x3x4 register window, as dw_tma.cu) ----
            const int ox = x0 + col;
            if (cvalid && ox < g.wo) {
                float2 win[3][3][2];
                auto load_row = [&](int slot, int iy) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        float2* v = win[slot][d];
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y)
                                     : "r"(hid + (iy * G::IW + col * STRIDE + d) * G::HPITCH + quad * 16));
                    }
                };
                const int oyl0 = rgrp * rpt;
                if (STRIDE == 1) { load_row(0, oyl0); load_row(1, oyl0 + 1); }
                else             { load_row(0, oyl0 * 2); }
                uint16_t* orow = g.out.p + (((int64_t)img * g.ho + y0 + oyl0) * g.wo + ox) * g.out.ld + c0;
#pragma unroll
                for (int i = 0; i < G::RPT; ++i) {
                    if (i >= rpt) break;
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
                    if (y0 + oyl >= g.ho) break;
                    float2 acc[2] = {br[0], br[1]};
                    const int slots[3] = {s0, s1, s2};
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const float2* v = win[slots[ky]][kx];
#pragma unroll
                            for (int j = 0; j < 2; ++j) acc[j] = __ffma2_rn(v[j], wr[ky * 3 + kx][j], acc[j]);
                        }
                    uint32_t h0, l0, h1, l1;
                    split2(relu6f(acc[0].x), relu6f(acc[0].y), h0, l0);
                    split2(relu6f(acc[1].x), relu6f(acc[1].y), h1, l1);
                    uint16_t* dst = orow + (int64_t)i * g.wo * g.out.ld;
                    *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
                    if (g.out.plane) *reinterpret_cast<uint2*>(dst + g.out.plane) = make_uint2(l0, l1);
                }
            }
        }
        img = nimg; y0 = ny0; x0 = nx0; cg = ncg;
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
    const int kpitch = (g.kp + 8) * 2, cb64 = g.cblocks * 64;
    const size_t smem = 2 * (size_t)G::ROWS * kpitch + 2 * (size_t)cb64 * kpitch + 11 * (size_t)cb64 * 4 + (size_t)G::ROWS * G::HPITCH + 256;
    if (smem > 227 * 1024) { set_error("expand_dw3x3: hidden %d needs %zu bytes of shared memory", g.hidden, smem); return UAVSAL_ENOTSUP; }
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

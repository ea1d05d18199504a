// Whole inverted-residual block in ONE kernel: 1x1 expand (+BN+ReLU6) -> depthwise 3x3 (+BN+ReLU6) -> 1x1 project (+BN, +x)
// (model.py:74-103 dwBlock / torchvision InvertedResidual) for the narrow stride-1 blocks: cin <= 64, cout <= 64,
// hidden = 6 cin a multiple of 64.  MobileNetV2 features.5/6/8/9/10, the temporal branch of the ST blocks (teConv_sub,
// 64 -> 384 -> 32 at 45x80) and the prior branches.
//
// These blocks are HBM-bound when run as three kernels: the 6x hidden tensor is written by the expand GEMM, read and
// re-written by the depthwise kernel and read again by the project GEMM (te.sub: 2.6 GB of a 3.2 GB block per 120 frames).
// Here it never leaves the SM.  Per tile of 16x8 output pixels a CTA
//   * loads the haloed input box (18x10 pixels x cin, bf16 hi/lo, K-major 128-B swizzle) ONCE          [TMA, warp 0]
//   * per 64-channel chunk of the hidden tensor:
//       expand   : hidden_halo[180 px (two M=128 tiles)] x [64] = X_halo . W1_chunk^T   (tcgen05, 3-term split, TMEM acc1[c&1])
//       epi 1    : TMEM -> + bias -> ReLU6 -> 0 outside the image (= the depthwise conv's padding) -> fp32 tile in shared memory
//       depthwise: sliding 3x3 window (as dwproj.cu) -> bias -> ReLU6 -> hi/lo split -> A2[c&1] in the UMMA K-major layout
//       project  : acc2 += A2 . W2_chunk^T                                              (tcgen05, TMEM acc2)
//     the expand MMAs of chunk c+1 run under the depthwise stage of chunk c (two acc1 buffers)
//   * drains acc2: + bias (+ residual) -> hi/lo split -> coalesced stores.
// The halo recompute costs 180/128 of a K <= 64 expand GEMM - a few percent of the tile's time.  Arithmetic and accumulation
// order are those of the separate kernels (gemm_tc2 -> dw3x3_tma -> gemm_tc2), so results are bit-identical to them.
#include <cstdio>

#include "tc_common.cuh"
#include "gemm_tc2.cuh"

namespace uavsal {

extern int g_tc_debug;

struct MbArgs {
    int n, H, W, cin, hidden, N;     // images, map size, channels in / hidden / out (N: cout rounded up to 16, the UMMA N)
    int Nv;                          // valid output channels (a multiple of 8, <= N): stores and residual loads stop there
    int ksteps1;                     // expand k-steps of 16 (K padded)
    int tiles_x, tiles_y, num_tiles, nchunks;
    const float* b1;                 // expand bias [hidden]
    const float* wd;                 // depthwise weights [9][hidden]
    const float* bd;                 // depthwise bias [hidden]
    const float* b2;                 // project bias [N]
    int flags;                       // UAVSAL_F_RESIDUAL
    Act res;
    ActW out;
    int dbg;                         // 1 << 22: CTA 0 prints a per-chunk phase trace of its third tile (development)
};

__device__ __forceinline__ unsigned long long mb_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int kMbTW = 16, kMbTH = 8, kMbIW = kMbTW + 2, kMbIH = kMbTH + 2, kMbPix = kMbIW * kMbIH;   // 180 haloed pixels
constexpr uint32_t kMbXPlane = 24 * 1024;                  // 192 rows x 128 B (the second M tile's rows 192..255 read whatever follows)
constexpr uint32_t kMbXBytes = kMbPix * 128;               // bytes landed per plane
constexpr uint32_t kMbHid = kMbPix * 256;                  // 46 080: fp32 hidden tile, 64 channels per pixel
constexpr uint32_t kMbA2Plane = 128 * 128;                 // 16 KiB
constexpr uint32_t kMbW1Plane = 64 * 128;                  // 64 hidden rows x 64 k

// 16-byte chunk k16 (0..15) of pixel p's 256-byte row sits at chunk (k16 & 8) | ((k16 ^ p) & 7): row-per-lane writes (epi 1) and
// channel-per-lane reads (depthwise) are both conflict-free
__device__ __forceinline__ uint32_t mb_hid_off(int p, int k16) { return (uint32_t)p * 256u + (uint32_t)(((k16 & 8) | ((k16 ^ p) & 7)) << 4); }

template <int TERMS>
__global__ void __launch_bounds__(kThreads2, 1) mbconv_kernel(const __grid_constant__ CUtensorMap tmX,
                                                             const __grid_constant__ CUtensorMap tmW1,
                                                             const __grid_constant__ CUtensorMap tmW2, const MbArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int NPL = TERMS == 3 ? 2 : 1;
    const uint32_t w2_plane = (uint32_t)g.N * 128;
    const uint32_t w_stage = NPL * (kMbW1Plane + w2_plane);                   // [W1 hi | W1 lo | W2 hi | W2 lo]
    uint8_t* xbuf = smem;                                                     // [NPL][kMbXPlane]
    uint8_t* a2buf = xbuf + NPL * kMbXPlane;                                  // [2][NPL][kMbA2Plane]
    uint8_t* wbuf = a2buf + 2 * NPL * kMbA2Plane;                             // [2][w_stage]
    uint8_t* hid = wbuf + 2 * w_stage;                                        // [kMbHid]
    uint64_t* bars = reinterpret_cast<uint64_t*>(hid + kMbHid);
    uint64_t* x_full = bars;            // TMA -> issuer
    uint64_t* x_empty = bars + 1;       // expand MMAs of the tile retired -> producer
    // The two weight operands of a chunk have rings of their own: W1(c) is released by the expand MMAs of chunk c, W2(c) only by its
    // project MMAs a chunk and a half later.  (With one ring of [W1|W2] stages W1(c) could not be fetched before project(c-2) had
    // retired, and the expand of chunk c - TMA latency + 24 MMAs - no longer fitted under the depthwise stage of chunk c-1.)
    uint64_t* w1_full = bars + 2;       // [2] TMA -> issuer
    uint64_t* w1_empty = bars + 4;      // [2] expand MMAs of the chunk retired -> producer
    uint64_t* acc1_full = bars + 6;     // [2] expand MMAs retired -> workers
    uint64_t* acc1_empty = bars + 8;    // [2] workers (16 warps) -> issuer
    uint64_t* a2_full = bars + 10;      // [2] workers (one arrive after the CTA-wide barrier) -> issuer
    uint64_t* a2_empty = bars + 12;     // [2] project MMAs retired -> workers
    uint64_t* acc2_full = bars + 14;
    uint64_t* acc2_empty = bars + 15;   // workers (16 warps) -> issuer
    uint64_t* w2_full = bars + 16;      // [2] TMA -> issuer
    uint64_t* w2_empty = bars + 18;     // [2] project MMAs of the chunk retired -> producer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    unsigned long long* trace = reinterpret_cast<unsigned long long*>(bars + 22);           // [8 chunks][8]
    const bool tracing = (g.dbg & (1 << 22)) && blockIdx.x == 0;
#define MB_TRACE(itv, c, slot) do { if (tracing && (itv) == 2 && (c) < 8) trace[(c) * 8 + (slot)] = mb_gtime(); } while (0)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(x_full, 1); mbar_init(x_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(w1_full + s, 1); mbar_init(w1_empty + s, 1);
            mbar_init(w2_full + s, 1); mbar_init(w2_empty + s, 1);
            mbar_init(acc1_full + s, 1); mbar_init(acc1_empty + s, kEpiWarps);
            mbar_init(a2_full + s, 1); mbar_init(a2_empty + s, 1);
        }
        mbar_init(acc2_full, 1); mbar_init(acc2_empty, kEpiWarps);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc2_col = 256;
    pdl_wait();

    auto tile_coords = [&](int t, int& img, int& y0, int& x0) {
        const int per = g.tiles_x * g.tiles_y;
        img = t / per;
        const int r = t - img * per;
        y0 = (r / g.tiles_x) * kMbTH;
        x0 = (r % g.tiles_x) * kMbTW;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t wc = 0;
            int it = 0;
            for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x, ++it) {
                int img, y0, x0;
                tile_coords(t, img, y0, x0);
                mbar_wait(x_empty, (it & 1) ^ 1);
                mbar_expect_tx(x_full, NPL * kMbXBytes);
#pragma unroll
                for (int p = 0; p < NPL; ++p) tma_load_5d(&tmX, x_full, xbuf + p * kMbXPlane, 0, x0 - 1, y0 - 1, img, p);
                for (int c = 0; c < g.nchunks; ++c, ++wc) {
                    const int s = wc & 1;
                    const uint32_t par = ((wc >> 1) & 1) ^ 1;
                    uint8_t* ws = wbuf + s * w_stage;
                    mbar_wait(w1_empty + s, par);
                    mbar_expect_tx(w1_full + s, NPL * kMbW1Plane);
#pragma unroll
                    for (int p = 0; p < NPL; ++p) tma_load_3d(&tmW1, w1_full + s, ws + p * kMbW1Plane, 0, c * 64, p);
                    mbar_wait(w2_empty + s, par);
                    mbar_expect_tx(w2_full + s, NPL * w2_plane);
#pragma unroll
                    for (int p = 0; p < NPL; ++p) tma_load_3d(&tmW2, w2_full + s, ws + NPL * kMbW1Plane + p * w2_plane, c * 64, 0, p);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc1 = umma_idesc(64), idesc2 = umma_idesc(g.N);
        uint32_t wc = 0;
        int it = 0;
        auto project = [&](uint32_t pc, bool first, bool last) {              // chunk counter pc: A2[pc & 1] x W2 of stage pc & 1
            const int s = pc & 1;
            mbar_wait(w2_full + s, (pc >> 1) & 1);
            mbar_wait(a2_full + s, (pc >> 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_hi = smem_u32(a2buf + s * NPL * kMbA2Plane);
                const uint32_t b_hi = smem_u32(wbuf + s * w_stage + NPL * kMbW1Plane);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    const uint64_t dah = umma_desc(a_hi + k * 32);
                    const uint64_t dbh = umma_desc(b_hi + k * 32);
                    umma_bf16(tmem_base + acc2_col, dah, dbh, idesc2, (!first || k) ? 1u : 0u);
                    if (TERMS == 3) {
                        const uint64_t dal = umma_desc(a_hi + kMbA2Plane + k * 32);
                        const uint64_t dbl = umma_desc(b_hi + w2_plane + k * 32);
                        umma_bf16(tmem_base + acc2_col, dah, dbl, idesc2, 1u);
                        umma_bf16(tmem_base + acc2_col, dal, dbh, idesc2, 1u);
                    }
                }
                umma_commit(a2_empty + s);
                umma_commit(w2_empty + s);
                if (last) umma_commit(acc2_full);
            }
            __syncwarp();
        };
        for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x, ++it) {
            mbar_wait(x_full, it & 1);
            for (int c = 0; c < g.nchunks; ++c, ++wc) {
                const int s = wc & 1;
                mbar_wait(w1_full + s, (wc >> 1) & 1);
                mbar_wait(acc1_empty + s, ((wc >> 1) & 1) ^ 1);               // epi 1 of two chunks ago has drained this buffer
                tc_fence_after();
                if (lane == 0) {
                    MB_TRACE(it, c, 6);
                    const uint32_t x_hi = smem_u32(xbuf);
                    const uint32_t b_hi = smem_u32(wbuf + s * w_stage);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint32_t d = tmem_base + (uint32_t)(s * 128 + mt * 64);
                        for (int k = 0; k < g.ksteps1; ++k) {
                            const uint64_t dah = umma_desc(x_hi + mt * (128 * 128) + k * 32);
                            const uint64_t dbh = umma_desc(b_hi + k * 32);
                            umma_bf16(d, dah, dbh, idesc1, k ? 1u : 0u);
                            if (TERMS == 3) {
                                const uint64_t dal = umma_desc(x_hi + kMbXPlane + mt * (128 * 128) + k * 32);
                                const uint64_t dbl = umma_desc(b_hi + kMbW1Plane + k * 32);
                                umma_bf16(d, dah, dbl, idesc1, 1u);
                                umma_bf16(d, dal, dbh, idesc1, 1u);
                            }
                        }
                    }
                    umma_commit(acc1_full + s);
                    umma_commit(w1_empty + s);
                    if (c == g.nchunks - 1) umma_commit(x_empty);
                }
                __syncwarp();
                if (c == 0) { mbar_wait(acc2_empty, (it & 1) ^ 1); tc_fence_after(); }   // the previous tile's output has left acc2
                if (c > 0) project(wc - 1, c == 1, false);
            }
            project(wc - 1, g.nchunks == 1, true);
        }
    } else {
        // ===================== workers (warps 2..17, 512 threads): epi 1 -> depthwise -> (end of tile) epilogue =====================
        const int et = threadIdx.x - 64;
        const int ew = warp - 2, q = warp & 3, j16 = ew >> 2;                  // TMEM lane quarter, 16-column piece
        const int quad = et & 15;                                             // depthwise: 4 channels of the chunk
        const int cp = (et >> 4) & 7;                                         //            output columns 2cp, 2cp+1
        const int rp = et >> 7;                                               //            output rows 2rp, 2rp+1
        const uint32_t hid_s = smem_u32(hid);
        const uint32_t wst = smem_u32(a2buf) + ew * 2048;                     // epilogue staging (both A2 buffers are idle then)
        uint32_t wc = 0;
        int it = 0;
        for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x, ++it) {
            int img, y0, x0;
            tile_coords(t, img, y0, x0);
            // this lane's haloed pixels in the two M tiles and whether they lie inside the image
            const int p0 = q * 32 + lane, p1 = 128 + p0;
            bool in0, in1;
            {
                const int iy = p0 / kMbIW, ix = p0 - iy * kMbIW;
                const int gy = y0 - 1 + iy, gx = x0 - 1 + ix;
                in0 = gy >= 0 && gy < g.H && gx >= 0 && gx < g.W;
            }
            {
                const int iy = p1 / kMbIW, ix = p1 - iy * kMbIW;
                const int gy = y0 - 1 + iy, gx = x0 - 1 + ix;
                in1 = p1 < kMbPix && gy >= 0 && gy < g.H && gx >= 0 && gx < g.W;
            }
            for (int c = 0; c < g.nchunks; ++c, ++wc) {
                const int s = wc & 1;
                const uint32_t par = (wc >> 1) & 1;
                // ---- epi 1: hidden tile of this chunk -> shared memory ----
                float b1v[16];
                {
                    const float4* bp = reinterpret_cast<const float4*>(g.b1 + c * 64 + j16 * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b4 = __ldg(bp + i);
                        b1v[4 * i] = b4.x; b1v[4 * i + 1] = b4.y; b1v[4 * i + 2] = b4.z; b1v[4 * i + 3] = b4.w;
                    }
                }
                if (et == 0) MB_TRACE(it, c, 0);
                mbar_wait(acc1_full + s, par);
                tc_fence_after();
                if (et == 0) MB_TRACE(it, c, 1);
                {
                    // both M tiles' TMEM reads are issued before the first is waited for (rows 192..255 of the second do not exist)
                    uint32_t raw[2][16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 128 + j16 * 16);
                    __syncwarp();
                    tmem_ld16_issue(taddr, raw[0]);
                    if (q < 2) tmem_ld16_issue(taddr + 64, raw[1]);
                    tmem_ld16_wait(raw[0]);
                    if (q < 2) tmem_ld16_wait(raw[1]);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        if (mt == 1 && q >= 2) break;
                        const int p = mt ? p1 : p0;
                        const bool inside = mt ? in1 : in0;
                        if (p < kMbPix) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float v[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) v[k] = inside ? relu6f(__uint_as_float(raw[mt][4 * i + k]) + b1v[4 * i + k]) : 0.f;
                                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hid_s + mb_hid_off(p, j16 * 4 + i)), "f"(v[0]), "f"(v[1]),
                                             "f"(v[2]), "f"(v[3]) : "memory");
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc1_empty + s);
                // depthwise taps / bias of this thread's 4 channels
                const int c0 = c * 64 + quad * 4;
                // (channel pairs: the 9-tap dot products run as packed fma.rn.f32x2 - the same IEEE fma per lane, half the instructions)
                float2 wr[9][2], br[2];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(g.wd + k * g.hidden + c0));
                    wr[k][0] = make_float2(w4.x, w4.y); wr[k][1] = make_float2(w4.z, w4.w);
                }
                {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bd + c0));
                    br[0] = make_float2(b4.x, b4.y); br[1] = make_float2(b4.z, b4.w);
                }
                if (et == 0) MB_TRACE(it, c, 2);
                named_bar_sync(1, kEpiThreads);                               // hidden tile complete
                mbar_wait(a2_empty + s, par ^ 1);                             // the project MMAs that read A2[s] two chunks ago retired
                if (et == 0) MB_TRACE(it, c, 3);
                // ---- depthwise 3x3 + bias + ReLU6 -> A2[s] (K-major, 128-B swizzle) ----
                {
                    const uint32_t a_hi = smem_u32(a2buf + s * NPL * kMbA2Plane);
                    float2 win[3][4][2];                                      // [row slot][column 2cp-1 .. 2cp+2][channel pair]
                    auto load_row = [&](int slot, int iy) {
#pragma unroll
                        for (int d = 0; d < 4; ++d) {
                            float2* v = win[slot][d];
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y)
                                         : "r"(hid_s + mb_hid_off(iy * kMbIW + 2 * cp + d, quad)));
                        }
                    };
                    const int oyl0 = rp * 2;
                    load_row(0, oyl0); load_row(1, oyl0 + 1);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int s0 = i % 3, s1 = (i + 1) % 3, s2 = (i + 2) % 3;
                        load_row(s2, oyl0 + i + 2);
                        const int slots[3] = {s0, s1, s2};
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            float2 acc[2] = {br[0], br[1]};
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx) {
                                    const float2* v = win[slots[ky]][cc + kx];
#pragma unroll
                                    for (int k = 0; k < 2; ++k) acc[k] = __ffma2_rn(v[k], wr[ky * 3 + kx][k], acc[k]);
                                }
                            const int r = (oyl0 + i) * kMbTW + 2 * cp + cc;   // A row = pixel index inside the tile
                            uint32_t h0, h1, l0, l1;
                            split2(relu6f(acc[0].x), relu6f(acc[0].y), h0, l0);
                            split2(relu6f(acc[1].x), relu6f(acc[1].y), h1, l1);
                            const uint32_t off = r * 128 + ((((uint32_t)quad >> 1) ^ (r & 7)) << 4) + (quad & 1) * 8;
                            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_hi + off), "r"(h0), "r"(h1) : "memory");
                            if (TERMS == 3) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_hi + kMbA2Plane + off), "r"(l0), "r"(l1) : "memory");
                        }
                    }
                }
                fence_async_smem();                                           // generic-proxy writes -> visible to the tensor core's async-proxy reads
                if (et == 0) MB_TRACE(it, c, 4);
                named_bar_sync(1, kEpiThreads);                               // A2[s] complete; the hidden tile may be overwritten
                if (et == 0) { mbar_arrive(a2_full + s); MB_TRACE(it, c, 5); }
            }

            // ---- epilogue of the tile: acc2 -> + bias (+ residual) -> split -> staging -> coalesced stores ----
            mbar_wait(acc2_full, it & 1);
            tc_fence_after();
            auto grow_of = [&](int rr) -> int64_t {
                const int y = y0 + rr / kMbTW, x = x0 + rr % kMbTW;
                if (y >= g.H || x >= g.W) return -1;
                return ((int64_t)img * g.H + y) * g.W + x;
            };
            const int64_t orow = grow_of(q * 32 + lane);
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc2_col;
            const int nch = (g.N + 63) >> 6;
            for (int ch = 0; ch < nch; ++ch) {
                const int n = ch * 64 + j16 * 16;
                if (n < g.N) {
                    uint32_t raw[16];
                    __syncwarp();
                    tmem_ld16(trow + n, raw);
                    float v[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(raw[k]);
                    if (orow >= 0) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.b2 + n) + j4);
                            v[j4 * 4 + 0] += b4.x; v[j4 * 4 + 1] += b4.y; v[j4 * 4 + 2] += b4.z; v[j4 * 4 + 3] += b4.w;
                        }
                        if (g.flags & UAVSAL_F_RESIDUAL) {
                            float rr[8];
                            load8(g.res.p + orow * g.res.ld + n, g.res.plane, rr);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] += rr[k];
                            if (n + 8 < g.Nv) {
                                load8(g.res.p + orow * g.res.ld + n + 8, g.res.plane, rr);
#pragma unroll
                                for (int k = 0; k < 8; ++k) v[8 + k] += rr[k];
                            }
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) split2(v[half * 8 + 2 * k], v[half * 8 + 2 * k + 1], h[k], l[k]);
                        const int off = lane * 32 + ((half ^ ((lane >> 2) & 1)) << 4);
                        sts128(wst + off, h[0], h[1], h[2], h[3]);
                        sts128(wst + 1024 + off, l[0], l[1], l[2], l[3]);
                    }
                    __syncwarp();
                    uint4 hv4[2], lv4[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int row = 16 * i + (lane >> 1), cc = lane & 1;
                        const int off = row * 32 + ((cc ^ ((row >> 2) & 1)) << 4);
                        hv4[i] = lds128(wst + off);
                        lv4[i] = lds128(wst + 1024 + off);
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int row = 16 * i + (lane >> 1), cc = lane & 1;
                        const int64_t gr = grow_of(q * 32 + row);
                        if (gr >= 0 && n + cc * 8 < g.Nv) {
                            uint16_t* dst = g.out.p + gr * g.out.ld + n + cc * 8;
                            *reinterpret_cast<uint4*>(dst) = hv4[i];
                            if (g.out.plane) *reinterpret_cast<uint4*>(dst + g.out.plane) = lv4[i];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc2_empty);
            // (the staging reads are ordered before the next tile's A2 writes by the barrier that follows its first epi 1)
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    if (tracing && threadIdx.x == 0) {
        const unsigned long long b = trace[0];
        for (int c = 0; c < min(g.nchunks, 8); ++c)
            printf("chunk %d: start %6lld | acc1 ready %6lld | epi1 done %6lld | hidden barrier + A2 free %6lld | dw done %6lld | A2 published %6lld | expand issued %6lld ns\n",
                   c, (long long)(trace[c * 8 + 0] - b), (long long)(trace[c * 8 + 1] - b), (long long)(trace[c * 8 + 2] - b), (long long)(trace[c * 8 + 3] - b),
                   (long long)(trace[c * 8 + 4] - b), (long long)(trace[c * 8 + 5] - b), (long long)(trace[c * 8 + 6] - b));
    }
#undef MB_TRACE
}

}  // namespace uavsal

using namespace uavsal;

extern "C" int uavsal_mbconv_fused(const uint16_t* x, int64_t x_plane, int x_ld, int n, int h, int w, int cin, const uint16_t* w1, int kp1,
                                   const float* b1, int hidden, const float* wd, const float* bd, const uint16_t* w2, int cout,
                                   const float* b2, int flags, int terms, const uint16_t* res, int64_t res_plane, int res_ld, uint16_t* out,
                                   int64_t out_plane, int out_ld, void* stream) {
    auto al16 = [](const void* p) { return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    UAVSAL_REQUIRE(al16(x) && al16(w1) && al16(b1) && al16(wd) && al16(bd) && al16(w2) && al16(b2) && al16(out) && n > 0 && h > 0 && w > 0 &&
                       x_ld % 8 == 0 && x_ld >= cin && x_plane % 8 == 0 && out_ld % 8 == 0 && out_ld >= cout && out_plane > 0 &&
                       out_plane % 8 == 0 && kp1 % 8 == 0 && kp1 >= cin,
                   UAVSAL_EINVAL, "mbconv_fused: bad arguments");
    UAVSAL_REQUIRE(cin % 8 == 0 && cin <= 64 && hidden % 64 == 0 && hidden >= 64 && cout % 8 == 0 && cout > 0 && cout <= 64 && (terms == 1 || terms == 3),
                   UAVSAL_ENOTSUP, "mbconv_fused: cin %d (<= 64), hidden %d (multiple of 64), cout %d (multiple of 8, <= 64)", cin, hidden, cout);
    const int cout16 = (cout + 15) & ~15;                         // the project MMA's N; w2 / b2 carry cout16 rows (zero beyond cout)
    UAVSAL_REQUIRE(terms == 1 || x_plane > 0, UAVSAL_EINVAL, "mbconv_fused: terms = 3 needs the lo plane of x");
    UAVSAL_REQUIRE(!(flags & UAVSAL_F_RESIDUAL) || (al16(res) && res_ld % 8 == 0 && res_plane % 8 == 0 && res_plane > 0), UAVSAL_EINVAL,
                   "mbconv_fused: residual requested without a residual tensor");
    UAVSAL_REQUIRE(!(flags & ~UAVSAL_F_RESIDUAL), UAVSAL_ENOTSUP, "mbconv_fused: only the residual flag is supported (the project conv is linear)");
    cudaStream_t s = (cudaStream_t)stream;
    MbArgs g{};
    g.n = n; g.H = h; g.W = w; g.cin = cin; g.hidden = hidden; g.N = cout16; g.Nv = cout;
    g.ksteps1 = (cin + 15) / 16;
    g.tiles_x = div_up(w, kMbTW); g.tiles_y = div_up(h, kMbTH);
    g.num_tiles = n * g.tiles_x * g.tiles_y;
    g.nchunks = hidden / 64;
    g.b1 = b1; g.wd = wd; g.bd = bd; g.b2 = b2; g.flags = flags; g.dbg = g_tc_debug;
    g.res = Act{res, res_plane, res_ld};
    g.out = ActW{out, out_plane, out_ld};
    const uint32_t npl = terms == 3 ? 2 : 1;
    CUtensorMap tX, tW1, tW2;
    {
        const uint64_t dims[5] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n, x_plane ? 2u : 1u};
        const uint64_t row = (uint64_t)x_ld * 2;
        const uint64_t str[4] = {row, row * w, row * w * h, x_plane ? (uint64_t)x_plane * 2 : row * w * h * (uint64_t)n};
        const uint32_t box[5] = {64, (uint32_t)kMbIW, (uint32_t)kMbIH, 1, 1};
        int rc = tc_encode(&tX, x, 5, dims, str, box, "mbconv_fused x (haloed tile)", 1);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)kp1, (uint64_t)hidden, 2};
        const uint64_t str[2] = {(uint64_t)kp1 * 2, (uint64_t)kp1 * 2 * (uint64_t)hidden};
        const uint32_t box[3] = {kBK, 64, 1};
        int rc = tc_encode(&tW1, w1, 3, dims, str, box, "mbconv_fused expand weights", 1);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)hidden, (uint64_t)cout16, 2};
        const uint64_t str[2] = {(uint64_t)hidden * 2, (uint64_t)hidden * 2 * (uint64_t)cout16};
        const uint32_t box[3] = {kBK, (uint32_t)cout16, 1};
        int rc = tc_encode(&tW2, w2, 3, dims, str, box, "mbconv_fused project weights", 1);
        if (rc) return rc;
    }
    const size_t smem = (size_t)npl * kMbXPlane + 2 * (size_t)npl * kMbA2Plane + 2 * (size_t)npl * (kMbW1Plane + (size_t)cout16 * 128) + kMbHid + 256 + 512 + 1024;
    UAVSAL_REQUIRE(smem <= 227 * 1024, UAVSAL_ENOTSUP, "mbconv_fused: tile does not fit shared memory");
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int grid = g.num_tiles < sms ? g.num_tiles : sms;
    cudaError_t e;
    if (terms == 3) {
        static bool attr = false;
        if (!attr) {
            e = cudaFuncSetAttribute(mbconv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("mbconv_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            attr = true;
        }
        e = launch_k(mbconv_kernel<3>, dim3(grid), dim3(kThreads2), smem, s, 1, tX, tW1, tW2, g);
    } else {
        static bool attr = false;
        if (!attr) {
            e = cudaFuncSetAttribute(mbconv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("mbconv_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            attr = true;
        }
        e = launch_k(mbconv_kernel<1>, dim3(grid), dim3(kThreads2), smem, s, 1, tX, tW1, tW2, g);
    }
    if (e != cudaSuccess) { set_error("mbconv_fused: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("mbconv_fused");
}

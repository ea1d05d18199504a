// Fused depthwise 3x3 (+BN+ReLU6) -> 1x1 project (+BN, +residual) of a dwBlock: model.py:92-101, conv[1] .. conv[3].
//
// The depthwise output of the 45x80 blocks (72 000 x 1536 x 4 B = 442 MB per 20-frame call) was written by one kernel and
// read back by the project GEMM; here it only ever exists as the A operand tile of a tcgen05 GEMM in shared memory:
//
//   warp 0        TMA producer : haloed hidden tile (18 x 10 pixels x 64 channels, fp32 or q16; OOB zero fill = conv padding)
//                                of k-block kb+1 and this CTA's half of the project-weight k-block (bf16 hi/lo, 128-B swizzle)
//   warps 2..17   depthwise    : sliding 3x3 window over the hidden tile -> bias -> ReLU6 -> hi/lo split -> A tile
//                                [128 pixels x 64 channels] written in the K-major 128-B-swizzled layout UMMA expects
//   warp 1        MMA issuer   : cta_group::2 MMAs (M = 256: the 16x8-pixel tiles of BOTH CTAs of the pair), fp32 in TMEM
//   warps 2..17   epilogue     : at the end of the tile the same warps drain TMEM (bias, residual, split, coalesced stores),
//                                reusing the A buffers as their transpose staging
//
// Per k-block a CTA ingests 46 KB (hidden) + 32 KB (weights) instead of the GEMM's 64 KB, and the 2 x 442 MB round trip of
// the depthwise output through HBM plus the separate depthwise launch disappear.
//
// Q16 (UAVSAL_F_HID_Q16): the hidden tensor arrives as 16-bit fixed point rows (common.cuh) - 23 KB per k-block, one LDS.64
// instead of one LDS.128 per 4 channels (the kernel is bound by the shared-memory data path) and half the HBM read.
#include "tc_common.cuh"
#include "gemm_tc2.cuh"

namespace uavsal {

struct DwProjArgs {
    int n, H, W, hidden, N;          // images, map size, hidden channels (= K), output channels
    int tiles_x, tiles_y, num_tiles, num_kb;
    const float* wd;                 // depthwise weights [9][hidden] (BN folded)
    const float* bd;                 // depthwise bias [hidden]
    const float* bias;               // project bias [N]
    int flags;                       // UAVSAL_F_RESIDUAL
    Act res;
    ActW out;
    int tmem_cols;
};

constexpr int kDpTW = 16, kDpTH = 8, kDpIW = kDpTW + 2, kDpIH = kDpTH + 2;
constexpr uint32_t kDpHBytesF32 = kDpIW * kDpIH * 256;     // 46 080: haloed tile, 64 fp32 channels per pixel
constexpr uint32_t kDpHBytesQ16 = kDpIW * kDpIH * 128;     // 23 040: 64 q16 channels per pixel
constexpr uint32_t kDpAPlane = 128 * 128;                  // 16 KiB: 128 rows x 64 bf16

template <int TERMS, bool Q16>
__global__ void __launch_bounds__(kThreads2, 1) dwproj_kernel(const __grid_constant__ CUtensorMap tmH,
                                                             const __grid_constant__ CUtensorMap tmB, const DwProjArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int NPL = TERMS == 3 ? 2 : 1;
    constexpr uint32_t kDpHBytes = Q16 ? kDpHBytesQ16 : kDpHBytesF32;
    constexpr uint32_t kPixB = Q16 ? 128 : 256;                               // bytes per pixel of the staged hidden tile
    const uint32_t b_plane = (uint32_t)(g.N / 2) * 128;                       // this CTA's half of the weight k-block, one plane
    const uint32_t a_stage = NPL * kDpAPlane, b_stage = NPL * b_plane;
    uint8_t* abuf = smem;                                                     // [2][a_stage]   (1024-aligned)
    uint8_t* bbuf = abuf + 2 * a_stage;                                       // [2][b_stage]   (b_plane % 1024 == 0)
    uint8_t* hbuf = bbuf + 2 * b_stage;                                       // [2][kDpHBytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(hbuf + 2 * kDpHBytes);
    uint64_t* h_full = bars;          // [2] TMA -> depthwise warps (local)
    uint64_t* h_empty = bars + 2;     // [2] depthwise warps -> producer (local, 16 warps)
    uint64_t* a_full = bars + 4;      // [2] depthwise warps of BOTH CTAs -> issuer (even CTA's copy is waited on)
    uint64_t* b_full = bars + 6;      // [2] TMA bytes of both CTAs -> issuer (even CTA's copy)
    uint64_t* ab_empty = bars + 8;    // [2] MMAs retired -> producer + depthwise warps (multicast commit, both CTAs)
    uint64_t* acc_full = bars + 10;   // accumulator complete (multicast commit, both CTAs)
    uint64_t* acc_empty = bars + 11;  // epilogue warps of both CTAs -> issuer (even CTA's copy)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const int num_pairs = (g.num_tiles + 1) / 2;
    const int first = (int)(blockIdx.x >> 1), step = (int)(gridDim.x >> 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(h_full + s, 1); mbar_init(h_empty + s, kEpiWarps / 2);
            mbar_init(a_full + s, kEpiWarps); mbar_init(b_full + s, 1); mbar_init(ab_empty + s, 1);
        }
        mbar_init(acc_full, 1); mbar_init(acc_empty, 2 * kEpiWarps);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    auto tile_coords = [&](int p, int& img, int& y0, int& x0) -> bool {
        const int t = 2 * p + (int)crank;
        const int per = g.tiles_x * g.tiles_y;
        img = t / per;
        const int r = t - img * per;
        y0 = (r / g.tiles_x) * kDpTH;
        x0 = (r % g.tiles_x) * kDpTW;
        return t < g.num_tiles;                                               // the odd tile out: img == n -> TMA zero fill, stores masked
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t kc = 0;
            for (int p = first; p < num_pairs; p += step) {
                int img, y0, x0;
                tile_coords(p, img, y0, x0);
                for (int kb = 0; kb < g.num_kb; ++kb, ++kc) {
                    const int s = kc & 1;
                    const uint32_t par = ((kc >> 1) & 1) ^ 1;
                    mbar_wait(h_empty + s, par);
                    mbar_expect_tx(h_full + s, kDpHBytes);
                    asm volatile(
                        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                        ::"r"(smem_u32(hbuf + s * kDpHBytes)), "l"(&tmH), "r"(smem_u32(h_full + s)), "r"(kb * 64), "r"(x0 - 1), "r"(y0 - 1), "r"(img)
                        : "memory");
                    mbar_wait(ab_empty + s, par);
                    if (crank == 0) mbar_expect_tx(b_full + s, 2 * b_stage);
                    const uint32_t fbar = mapa_u32(smem_u32(b_full + s), 0);
#pragma unroll
                    for (int pl = 0; pl < NPL; ++pl)
                        tma_load_3d_pair(&tmB, fbar, bbuf + s * b_stage + pl * b_plane, kb * kBK, (int)crank * (g.N / 2), pl);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (even CTA of the pair) =====================
        if (crank == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            uint32_t kc = 0;
            int it = 0;
            for (int p = first; p < num_pairs; p += step, ++it) {
                mbar_wait_cluster(acc_empty, (it & 1) ^ 1);                   // both CTAs' epilogues have drained the accumulator
                tc_fence_after();
                for (int kb = 0; kb < g.num_kb; ++kb, ++kc) {
                    const int s = kc & 1;
                    const uint32_t par = (kc >> 1) & 1;
                    mbar_wait_cluster(a_full + s, par);                       // A tiles written by the depthwise warps of both CTAs
                    mbar_wait(b_full + s, par);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_hi = smem_u32(abuf + s * a_stage);
                        const uint32_t b_hi = smem_u32(bbuf + s * b_stage);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            const uint64_t dah = umma_desc(a_hi + k * 32);
                            const uint64_t dbh = umma_desc(b_hi + k * 32);
                            umma_bf16_pair(tmem_base, dah, dbh, idesc, (kb | k) ? 1u : 0u);
                            if (TERMS == 3) {
                                const uint64_t dal = umma_desc(a_hi + kDpAPlane + k * 32);
                                const uint64_t dbl = umma_desc(b_hi + b_plane + k * 32);
                                umma_bf16_pair(tmem_base, dah, dbl, idesc, 1u);
                                umma_bf16_pair(tmem_base, dal, dbh, idesc, 1u);
                            }
                        }
                        umma_commit_pair(ab_empty + s);
                        if (kb == g.num_kb - 1) umma_commit_pair(acc_full);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================== depthwise producer + epilogue (warps 2..17, 512 threads) =====================
        // The 16 warps form two groups of 8; group gsel owns the k-blocks kb == gsel (mod 2) and therefore A/hidden buffer gsel
        // (num_kb is even).  Inside a group a thread owns 4 channels x 2 adjacent columns x 4 rows: 24 LDS.128 and 40 weight
        // floats per 8 output pixels (the 1-column x 4-row mapping of the first version needed 36 and 80 - the shared-memory
        // / L1 data path, not the FMA pipe, limits this kernel).
        const int et = threadIdx.x - 64;
        const int gsel = et >> 8;
        const int gt = et & 255;
        const int quad = gt & 15;                                             // 4 channels of the 64-channel k-block
        const int cp = (gt >> 4) & 7;                                         // output columns 2cp, 2cp+1 of the 16x8 tile
        const int rgrp = gt >> 7;                                             // output rows rgrp*4 .. +3
        const int ew = warp - 2, q = warp & 3, sub = (ew >> 2) * 16;
        const uint32_t wst = smem_u32(abuf) + ew * 2048;                      // epilogue staging (A buffers are idle then)
        const uint32_t a_full_leader0 = mapa_u32(smem_u32(a_full), 0);
        const uint32_t acc_empty_leader = mapa_u32(smem_u32(acc_empty), 0);
        int it = 0;
        for (int p = first; p < num_pairs; p += step, ++it) {
            int img, y0, x0;
            const bool tvalid = tile_coords(p, img, y0, x0);
            for (int kb = gsel; kb < g.num_kb; kb += 2) {
                const int s = gsel;
                const uint32_t par = (uint32_t)((it * (g.num_kb >> 1) + (kb >> 1)) & 1);   // use count of buffer s so far
                const int c0 = kb * 64 + quad * 4;
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
                mbar_wait(h_full + s, par);                                   // hidden tile landed
                mbar_wait(ab_empty + s, par ^ 1);                             // the MMAs that read this A buffer two k-blocks ago retired
                const uint32_t tile = smem_u32(hbuf + s * kDpHBytes);
                const uint32_t a_hi = smem_u32(abuf + s * a_stage);
                float2 win[3][4][2];                                          // [row slot][column 2cp-1 .. 2cp+2][channel pair]
                auto load_row = [&](int slot, int iy) {
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        float2* v = win[slot][d];
                        if (Q16) {
                            uint32_t w0, w1;
                            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1)
                                         : "r"(tile + (iy * kDpIW + 2 * cp + d) * kPixB + quad * 8));
                            v[0] = q16_unpack2(w0); v[1] = q16_unpack2(w1);
                        } else {
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y)
                                         : "r"(tile + (iy * kDpIW + 2 * cp + d) * kPixB + quad * 16));
                        }
                    }
                };
                const int oyl0 = rgrp * 4;
                load_row(0, oyl0); load_row(1, oyl0 + 1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int s0 = i % 3, s1 = (i + 1) % 3, s2 = (i + 2) % 3;
                    load_row(s2, oyl0 + i + 2);
                    const int slots[3] = {s0, s1, s2};
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        float2 acc[2] = {br[0], br[1]};
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const float2* v = win[slots[ky]][c + kx];
#pragma unroll
                                for (int j = 0; j < 2; ++j) acc[j] = __ffma2_rn(v[j], wr[ky * 3 + kx][j], acc[j]);
                            }
                        // A row = pixel index inside the tile; 16-byte chunk (quad >> 1) of the 128-byte row sits at chunk ^ (row & 7)
                        const int r = (oyl0 + i) * kDpTW + 2 * cp + c;
                        uint32_t h0, h1, l0, l1;
                        split2(relu6f(acc[0].x), relu6f(acc[0].y), h0, l0);
                        split2(relu6f(acc[1].x), relu6f(acc[1].y), h1, l1);
                        const uint32_t off = r * 128 + ((((uint32_t)quad >> 1) ^ (r & 7)) << 4) + (quad & 1) * 8;
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_hi + off), "r"(h0), "r"(h1) : "memory");
                        if (TERMS == 3) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_hi + kDpAPlane + off), "r"(l0), "r"(l1) : "memory");
                    }
                }
                fence_async_smem();                                           // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster(a_full_leader0 + s * 8);
                    mbar_arrive(h_empty + s);
                }
            }

            // ---- epilogue of this tile (same warps): TMEM -> bias (+ residual) -> split -> staging -> coalesced stores ----
            mbar_wait(acc_full, it & 1);
            tc_fence_after();
            auto grow_of = [&](int rr) -> int64_t {
                const int y = y0 + rr / kDpTW, x = x0 + rr % kDpTW;
                if (!tvalid || y >= g.H || x >= g.W) return -1;
                return ((int64_t)img * g.H + y) * g.W + x;
            };
            const int r = q * 32 + lane;
            const int64_t orow = grow_of(r);
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
            const int nchunks = g.N >> 6;
            for (int ch = 0; ch < nchunks; ++ch) {
                uint32_t raw[16];
                __syncwarp();
                tmem_ld16(trow + ch * 64 + sub, raw);
                const int n = ch * 64 + sub;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
                if (orow >= 0) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n) + j4);
                        v[j4 * 4 + 0] += b4.x; v[j4 * 4 + 1] += b4.y; v[j4 * 4 + 2] += b4.z; v[j4 * 4 + 3] += b4.w;
                    }
                    if (g.flags & UAVSAL_F_RESIDUAL) {
                        float rr[8];
                        load8(g.res.p + orow * g.res.ld + n, g.res.plane, rr);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += rr[j];
                        load8(g.res.p + orow * g.res.ld + n + 8, g.res.plane, rr);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[8 + j] += rr[j];
                    }
                }
                if (ch == nchunks - 1) {                                      // accumulator fully read by this warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc_empty_leader);
                }
                __syncwarp();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split2(v[half * 8 + 2 * j], v[half * 8 + 2 * j + 1], h[j], l[j]);
                    const int off = lane * 32 + ((half ^ ((lane >> 2) & 1)) << 4);
                    sts128(wst + off, h[0], h[1], h[2], h[3]);
                    sts128(wst + 1024 + off, l[0], l[1], l[2], l[3]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int row = 16 * i + (lane >> 1), c = lane & 1;
                    const int off = row * 32 + ((c ^ ((row >> 2) & 1)) << 4);
                    const uint4 hv4 = lds128(wst + off);
                    const uint4 lv4 = lds128(wst + 1024 + off);
                    const int64_t gr = grow_of(q * 32 + row);
                    if (gr >= 0) {
                        uint16_t* dst = g.out.p + gr * g.out.ld + n + c * 8;
                        *reinterpret_cast<uint4*>(dst) = hv4;
                        if (g.out.plane) *reinterpret_cast<uint4*>(dst + g.out.plane) = lv4;
                    }
                }
            }
            named_bar_sync(1, kEpiThreads);                                   // staging reads done before the next tile's A writes
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
}

template <int TERMS, bool Q16>
static int launch_dwproj(const CUtensorMap& tH, const CUtensorMap& tB, DwProjArgs& g, cudaStream_t s) {
    const uint32_t npl = TERMS == 3 ? 2 : 1;
    const size_t smem = 2 * (size_t)npl * kDpAPlane + 2 * (size_t)npl * (g.N / 2) * 128 + 2 * (size_t)(Q16 ? kDpHBytesQ16 : kDpHBytesF32) + 256 + 1024;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(dwproj_kernel<TERMS, Q16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("dw_project: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    UAVSAL_REQUIRE(smem <= 227 * 1024, UAVSAL_ENOTSUP, "dw_project: tile does not fit shared memory");
    static int sms = 0, max_clusters = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    if (!max_clusters) {
        cudaLaunchConfig_t q{};
        q.gridDim = dim3((sms / 2) * 2); q.blockDim = dim3(kThreads2); q.dynamicSmemBytes = 227 * 1024 - 1024; q.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        q.attrs = at; q.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, dwproj_kernel<TERMS, Q16>, &q) != cudaSuccess || n <= 0) { n = sms / 2 - 2; cudaGetLastError(); }
        max_clusters = n;
    }
    const int pairs = (g.num_tiles + 1) / 2;
    const int grid = 2 * (pairs < max_clusters ? pairs : max_clusters);
    cudaError_t e = launch_k(dwproj_kernel<TERMS, Q16>, dim3(grid), dim3(kThreads2), smem, s, 2, tH, tB, g);
    if (e != cudaSuccess) { set_error("dw_project: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("dw_project");
}

}  // namespace uavsal

using namespace uavsal;

namespace uavsal {
int dw_project32(const float* hid, int hid_ld, int n, int h, int w, const float* wd, const float* bd, const uint16_t* wgt, int kpad,
                 const float* bias, ActW out, cudaStream_t s);
}

extern "C" int uavsal_dw_project(const float* hid, int hid_ld, int n, int h, int w, int hidden, const float* wd, const float* bd,
                                 const uint16_t* wgt, int kpad, int cout, const float* bias, int flags, int terms,
                                 const uint16_t* res, int64_t res_plane, int res_ld, uint16_t* out, int64_t out_plane, int out_ld,
                                 void* stream) {
    auto al16 = [](const void* p) { return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool q16 = flags & UAVSAL_F_HID_Q16;                    // `hid` holds uint16 fixed-point rows (uavsal_pw_gemm with UAVSAL_F_OUT_Q16)
    flags &= ~UAVSAL_F_HID_Q16;
    UAVSAL_REQUIRE(al16(hid) && al16(wd) && al16(bd) && al16(wgt) && al16(bias) && al16(out) && n > 0 && h > 0 && w > 0 &&
                       hid_ld % (q16 ? 8 : 4) == 0 && hid_ld >= hidden && out_ld % 8 == 0 && out_ld >= cout && out_plane > 0 && out_plane % 8 == 0 &&
                       kpad >= hidden && kpad % 8 == 0,
                   UAVSAL_EINVAL, "dw_project: bad arguments");
    if (hidden == 32 && cout == 16) {                             // features.1: fp32 FFMA kernel (dwproj32.cu), exact in both precision modes
        UAVSAL_REQUIRE(!flags && !q16, UAVSAL_ENOTSUP, "dw_project: the 32 -> 16 kernel takes fp32 rows and has no residual input");
        return dw_project32(hid, hid_ld, n, h, w, wd, bd, wgt, kpad, bias, ActW{out, out_plane, out_ld}, (cudaStream_t)stream);
    }
    UAVSAL_REQUIRE(hidden % 128 == 0 && cout % 64 == 0 && cout <= 256 && (terms == 1 || terms == 3), UAVSAL_ENOTSUP,
                   "dw_project: hidden %d must be a multiple of 128 and cout %d a multiple of 64 up to 256 (or 32 -> 16)", hidden, cout);
    UAVSAL_REQUIRE(!(flags & UAVSAL_F_RESIDUAL) || (al16(res) && res_ld % 8 == 0 && res_plane % 8 == 0 && res_plane > 0), UAVSAL_EINVAL,
                   "dw_project: residual requested without a residual tensor");
    UAVSAL_REQUIRE(!(flags & ~UAVSAL_F_RESIDUAL), UAVSAL_ENOTSUP, "dw_project: only the residual flag is supported (the project conv is linear)");
    DwProjArgs g{};
    g.n = n; g.H = h; g.W = w; g.hidden = hidden; g.N = cout;
    g.tiles_x = div_up(w, kDpTW); g.tiles_y = div_up(h, kDpTH);
    g.num_tiles = n * g.tiles_x * g.tiles_y;
    g.num_kb = hidden / 64;
    g.wd = wd; g.bd = bd; g.bias = bias; g.flags = flags;
    g.res = Act{res, res_plane, res_ld};
    g.out = ActW{out, out_plane, out_ld};
    int cols = 32;
    while (cols < cout) cols <<= 1;
    g.tmem_cols = cols;
    CUtensorMap tH, tB;
    {
        const uint64_t dims[4] = {(uint64_t)hidden, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        const uint64_t row = (uint64_t)hid_ld * (q16 ? 2 : 4);
        const uint64_t str[3] = {row, row * w, row * w * h};
        const uint32_t box[4] = {64, (uint32_t)kDpIW, (uint32_t)kDpIH, 1};
        int rc = q16 ? tc_encode(&tH, hid, 4, dims, str, box, "dw_project hidden (q16)", 0)      // 2-byte elements, zero fill decodes to 0.0
                     : tc_encode(&tH, hid, 4, dims, str, box, "dw_project hidden (f32)", 2);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)kpad, (uint64_t)cout, 2};
        const uint64_t str[2] = {(uint64_t)kpad * 2, (uint64_t)kpad * 2 * (uint64_t)cout};
        const uint32_t box[3] = {kBK, (uint32_t)(cout / 2), 1};
        int rc = tc_encode(&tB, wgt, 3, dims, str, box, "dw_project weights", 1);
        if (rc) return rc;
    }
    if (q16) {
        if (terms == 3) return launch_dwproj<3, true>(tH, tB, g, (cudaStream_t)stream);
        return launch_dwproj<1, true>(tH, tB, g, (cudaStream_t)stream);
    }
    if (terms == 3) return launch_dwproj<3, false>(tH, tB, g, (cudaStream_t)stream);
    return launch_dwproj<1, false>(tH, tB, g, (cudaStream_t)stream);
}

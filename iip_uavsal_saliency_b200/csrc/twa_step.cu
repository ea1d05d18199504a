// One ConvTWA step (model_convlstm.py:276-292) as an implicit GEMM whose A operand is RESIDENT: the haloed h_{t-1} tile of a
// channel block is loaded once and the nine filter taps read shifted views of it.
//
// The generic implicit-GEMM kernel (gemm_tc2.cuh, MODE_CONV) re-fetches the 128-pixel A tile for every tap: 36 k-blocks x
// 48 KB = 1.7 MB per CTA per step, and the step (41 us) is bound by per-SM TMA ingest.  Here a CTA owns an 8-wide x 16-high
// pixel tile; the haloed (10 x 18 pixel) x 64-channel box of h_{t-1} lands in shared memory as 180 rows of 128 swizzled
// bytes, and tap (dy, dx) is the UMMA descriptor that starts ((1+dy)*10 + (1+dx)) rows further down with a stride of ten
// rows (1280 B) between the 8-row groups: TMEM lane 8g + i = pixel (y0 + g, x0 + i).  A-side traffic drops 9x (184 KB per
// CTA per step); the recurrent weights stream through a 7-stage ring.  Epilogue = ConvTWA gate + blend as in gemm_tc2.cuh.
#include "tc_common.cuh"
#include "gemm_tc2.cuh"

namespace uavsal {

struct TwaStepArgs {
    int H, W, C, bn;                  // map size, channels (input = hidden = output), N tile
    int tiles_x, tiles_y, ncb;        // ncb = C / 64
    int a_img, out_img;               // image index of h_{t-1} in its tensor, of h_t / x_t in theirs (sequence 0)
    int a_stride, out_stride;         // image-index stride between the sequences of a batch (blockIdx.z)
    int bk_tap_stride, bk_off;        // weight K coordinate of (tap, cb) = tap * bk_tap_stride + bk_off + cb * 64
    const float* gx;                  // hoisted W_x * x_t pre-activations [rows][C] (fp32)
    Act x, hprev;
    ActW out;
    int tmem_cols;
    int base_off_mode;                // descriptor base-offset convention for row-shifted views (see umma_desc_shift)
    int dbg;                          // DBG_* timing-ablation bits
    int bstages;                      // depth of the weight ring
};

constexpr int kTwTW = 8, kTwTH = 16, kTwIW = kTwTW + 2, kTwIH = kTwTH + 2;
constexpr uint32_t kTwAPlaneBytes = kTwIW * kTwIH * 128;              // 23 040 bytes landed per plane
constexpr uint32_t kTwAPlane = 23 * 1024;                             // plane pitch (1024-aligned for the 128-B swizzle)
constexpr int kTwBStagesMax = 7;

// K-major SW128 descriptor with an arbitrary stride between 8-row groups and an optional base offset
__device__ __forceinline__ uint64_t umma_desc_shift(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) |
           (2ull << 61);
}

template <int TERMS>
__global__ void __launch_bounds__(kThreads2, 1) twa_step_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB, const TwaStepArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int NPL = TERMS == 3 ? 2 : 1;
    const uint32_t a_stage = NPL * kTwAPlane;
    const uint32_t b_plane = (uint32_t)g.bn * 128, b_stage = NPL * b_plane;
    uint8_t* abuf = smem;                                                     // [2][a_stage]
    uint8_t* bbuf = abuf + 2 * a_stage;                                       // [kTwBStages][b_stage]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bbuf + g.bstages * b_stage);
    uint64_t* a_full = bars;                   // [2]
    uint64_t* a_empty = bars + 2;              // [2]
    uint64_t* b_full = bars + 4;               // [kTwBStages]
    uint64_t* b_empty = b_full + kTwBStagesMax;
    uint64_t* acc_full = b_empty + kTwBStagesMax;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, n0 = blockIdx.y * g.bn;
    const int a_img = g.a_img + (int)blockIdx.z * g.a_stride, out_img = g.out_img + (int)blockIdx.z * g.out_stride;
    const int y0 = (tile / g.tiles_x) * kTwTH, x0 = (tile % g.tiles_x) * kTwTW;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < g.bstages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                                               // h_{t-1} is the previous step's output

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int kbB = 0;
            for (int cb = 0; cb < g.ncb; ++cb) {
                const int sa = cb & 1;
                mbar_wait(a_empty + sa, ((cb >> 1) & 1) ^ 1);
                mbar_expect_tx(a_full + sa, NPL * kTwAPlaneBytes);
#pragma unroll
                for (int p = 0; p < NPL; ++p)
                    tma_load_5d(&tmA, a_full + sa, abuf + sa * a_stage + p * kTwAPlane, cb * 64, x0 - 1, y0 - 1, a_img, p);
                for (int tap = 0; tap < 9; ++tap, ++kbB) {
                    const int s = kbB % g.bstages;
                    mbar_wait(b_empty + s, ((kbB / g.bstages) & 1) ^ 1);
                    if (g.dbg & DBG_NO_B) { mbar_arrive(b_full + s); continue; }
                    mbar_expect_tx(b_full + s, b_stage);
#pragma unroll
                    for (int p = 0; p < NPL; ++p)
                        tma_load_3d(&tmB, b_full + s, bbuf + s * b_stage + p * b_plane, tap * g.bk_tap_stride + g.bk_off + cb * 64, n0, p);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc(g.bn);
        int kbB = 0;
        for (int cb = 0; cb < g.ncb; ++cb) {
            const int sa = cb & 1;
            mbar_wait(a_full + sa, (cb >> 1) & 1);
            for (int tap = 0; tap < 9; ++tap, ++kbB) {
                const int s = kbB % g.bstages;
                mbar_wait(b_full + s, (kbB / g.bstages) & 1);
                tc_fence_after();
                if (lane == 0) {
                    // rows of the tap's view: haloed pixel ((g + tap/3) * 10 + i + tap%3), g = 0..15 (stride 1280 B), i = 0..7
                    const uint32_t a_hi = smem_u32(abuf + sa * a_stage) + (uint32_t)((tap / 3) * kTwIW + tap % 3) * 128;
                    const uint32_t b_hi = smem_u32(bbuf + s * b_stage);
                    const uint32_t boff = g.base_off_mode ? ((a_hi >> 7) & 7) : 0;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        if (g.dbg & DBG_NO_MMA) break;
                        const uint64_t dah = umma_desc_shift(a_hi + k * 32, kTwIW * 128, boff);
                        const uint64_t dbh = umma_desc(b_hi + k * 32);
                        umma_bf16(tmem_base, dah, dbh, idesc, (cb | tap | k) ? 1u : 0u);
                        if (TERMS == 3) {
                            const uint64_t dal = umma_desc_shift(a_hi + kTwAPlane + k * 32, kTwIW * 128, boff);
                            const uint64_t dbl = umma_desc(b_hi + b_plane + k * 32);
                            umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                            umma_bf16(tmem_base, dal, dbh, idesc, 1u);
                        }
                    }
                    umma_commit(b_empty + s);
                    if (tap == 8) umma_commit(a_empty + sa);
                    if (tap == 8 && cb == g.ncb - 1) umma_commit(acc_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: gate + blend (model_convlstm.py:283,290) =====================
        const int ew = warp - 2, q = warp & 3, sub = (ew >> 2) * 16;
        const uint32_t wst = smem_u32(bbuf) + ew * 2048;                      // the weight ring is idle once the accumulator is complete
        auto pix_of = [&](int rr) -> int64_t {                                // TMEM lane 8g + i = pixel (y0 + g, x0 + i)
            const int y = y0 + (rr >> 3), x = x0 + (rr & 7);
            return (y < g.H && x < g.W) ? (int64_t)y * g.W + x : -1;
        };
        const int r = q * 32 + lane;
        const int64_t pix = pix_of(r);
        const int64_t hw = (int64_t)g.H * g.W;
        const int64_t orow = (int64_t)out_img * hw + pix, hrow = (int64_t)a_img * hw + pix;
        // the blend operands and the hoisted W_x*x_t term do not depend on the accumulator: those of the first 64-column chunk
        // are fetched while the MMAs run (issued after the wait they cost 8 us of a 33 us step: three dependent global round
        // trips per thread); wider N tiles load the later chunks' operands in the loop
        float gxv[16], xv[16], hv[16];
        const bool rowlive = pix >= 0 && !(g.dbg & DBG_NO_STORE);
        auto fetch = [&](int n) {
            const float4* gp = reinterpret_cast<const float4*>(g.gx + orow * g.C + n);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = __ldg(gp + j4);
                gxv[j4 * 4 + 0] = b4.x; gxv[j4 * 4 + 1] = b4.y; gxv[j4 * 4 + 2] = b4.z; gxv[j4 * 4 + 3] = b4.w;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                load8(g.x.p + orow * g.x.ld + n + half * 8, g.x.plane, xv + half * 8);
                load8(g.hprev.p + hrow * g.hprev.ld + n + half * 8, g.hprev.plane, hv + half * 8);
            }
        };
        if (rowlive) fetch(n0 + sub);
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int ch = 0; ch < (g.bn >> 6); ++ch) {
            const int n = n0 + ch * 64 + sub;
            if (ch > 0 && rowlive) fetch(n);
            uint32_t raw[16];
            __syncwarp();
            tmem_ld16(trow + ch * 64 + sub, raw);
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
            if (rowlive) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float gi = sigmoid_acc(v[j] + gxv[j]);
                    v[j] = gi * xv[j] + (1.f - gi) * hv[j];
                }
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
                const int64_t px = pix_of(q * 32 + row);
                if (px >= 0 && !(g.dbg & DBG_NO_STORE)) {
                    uint16_t* dst = g.out.p + ((int64_t)out_img * hw + px) * g.out.ld + n + c * 8;
                    *reinterpret_cast<uint4*>(dst) = hv4;
                    if (g.out.plane) *reinterpret_cast<uint4*>(dst + g.out.plane) = lv4;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
}

int g_twa_bn = 64;          // uavsal_set_option key 8 (dev): N tile of the resident-A step kernel (64 | 128)
int g_twa_resident = 1;     // uavsal_set_option key 7: 0 = generic implicit GEMM per step, 1 = resident-A kernel (2 = with descriptor base
                            // offsets: WRONG results - kept as the record of the experiment that settled the swizzle convention)

// one step: seq[out_img] = blend(sigmoid(gx[out_img] + conv3x3(hsrc[a_img]; W_h)), x[out_img], hsrc[a_img])
// batch sequences per launch: sequence b reads image a_img + b*a_stride and writes image out_img + b*out_stride
int twa_step_resident(Act hsrc, int hsrc_nimg, int a_img, int a_stride, Act x, ActW seq, int out_img, int out_stride, int batch, int H, int W, int c,
                      const uint16_t* wgt, int wk_total, int wk_off, const float* gx, int terms, cudaStream_t s, int dbg) {
    TwaStepArgs g{};
    const uint32_t npl = terms == 3 ? 2 : 1;
    g.H = H; g.W = W; g.C = c;
    // N tile: as narrow as keeps the whole step on one wave of SMs (the step is latency-bound; measured equal time at 64 / 128)
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int tiles = div_up(W, kTwTW) * div_up(H, kTwTH);
    g.bn = 64;
    while (g.bn < 256 && c % (2 * g.bn) == 0 && (int64_t)tiles * batch * (c / g.bn) > sms) g.bn *= 2;
    if (g_twa_bn == 128 && c % 128 == 0 && g.bn < 128) g.bn = 128;
    g.tiles_x = div_up(W, kTwTW); g.tiles_y = div_up(H, kTwTH); g.ncb = c / 64;
    g.a_img = a_img; g.out_img = out_img; g.a_stride = a_stride; g.out_stride = out_stride;
    g.bk_tap_stride = wk_total; g.bk_off = wk_off;
    g.gx = gx; g.x = x; g.hprev = hsrc; g.out = seq;
    g.tmem_cols = g.bn;
    g.base_off_mode = g_twa_resident == 2;
    g.dbg = dbg;
    CUtensorMap tA, tB;
    {
        const uint64_t dims[5] = {(uint64_t)c, (uint64_t)W, (uint64_t)H, (uint64_t)hsrc_nimg, hsrc.plane ? 2u : 1u};
        const uint64_t row = (uint64_t)hsrc.ld * 2;
        const uint64_t str[4] = {row, row * W, row * W * H, hsrc.plane ? (uint64_t)hsrc.plane * 2 : row * W * H * (uint64_t)hsrc_nimg};
        const uint32_t box[5] = {64, (uint32_t)kTwIW, (uint32_t)kTwIH, 1, 1};
        int rc = tc_encode(&tA, hsrc.p, 5, dims, str, box, "twa_step A (haloed tile)", 1);
        if (rc) return rc;
    }
    {
        const uint64_t kpad = 9ull * wk_total;
        const uint64_t dims[3] = {kpad, (uint64_t)c, 2};
        const uint64_t str[2] = {kpad * 2, kpad * 2 * (uint64_t)c};
        const uint32_t box[3] = {kBK, (uint32_t)g.bn, 1};
        int rc = tc_encode(&tB, wgt, 3, dims, str, box, "twa_step B (weights)", 1);
        if (rc) return rc;
    }
    g.bstages = (int)((220u * 1024u - 2u * npl * kTwAPlane) / (npl * (uint32_t)g.bn * 128u));
    if (g.bstages > kTwBStagesMax) g.bstages = kTwBStagesMax;
    const size_t smem = 2 * (size_t)npl * kTwAPlane + (size_t)g.bstages * npl * g.bn * 128 + 256 + 1024;
    cudaError_t e;
    if (terms == 3) {
        static bool attr = false;
        if (!attr) {
            e = cudaFuncSetAttribute(twa_step_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("twa_step: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            attr = true;
        }
        e = launch_k(twa_step_kernel<3>, dim3(g.tiles_x * g.tiles_y, c / g.bn, batch), dim3(kThreads2), smem, s, 1, tA, tB, g);
    } else {
        static bool attr = false;
        if (!attr) {
            e = cudaFuncSetAttribute(twa_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("twa_step: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            attr = true;
        }
        e = launch_k(twa_step_kernel<1>, dim3(g.tiles_x * g.tiles_y, c / g.bn, batch), dim3(kThreads2), smem, s, 1, tA, tB, g);
    }
    if (e != cudaSuccess) { set_error("twa_step: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("twa_step");
}

}  // namespace uavsal

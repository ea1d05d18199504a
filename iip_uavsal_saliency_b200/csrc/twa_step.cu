// One ConvTWA step (model_convlstm.py:276-292) as an implicit GEMM whose A operand is RESIDENT: the haloed h_{t-1} tile of a
// channel block is loaded once and the nine filter taps read shifted views of it.
//
// The generic implicit-GEMM kernel (gemm_tc2.cuh, MODE_CONV) re-fetches the 128-pixel A tile for every tap: 36 k-blocks x
// 48 KB = 1.7 MB per CTA per step, and the step (41 us) is bound by per-SM TMA ingest.  Here a CTA owns an 8-wide x 16-high
// pixel tile; the haloed (10 x 18 pixel) x 64-channel box of h_{t-1} lands in shared memory as 180 rows of 128 swizzled
// bytes, and tap (dy, dx) is the UMMA descriptor that starts ((1+dy)*10 + (1+dx)) rows further down with a stride of ten
// rows (1280 B) between the 8-row groups: TMEM lane 8g + i = pixel (y0 + g, x0 + i).  A-side traffic drops 9x (184 KB per
// CTA per step); the recurrent weights stream through a 7-stage ring.  Epilogue = ConvTWA gate + blend as in gemm_tc2.cuh.
#include <cstdio>

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
int g_twa_resident = 1;     // uavsal_set_option key 7: 0 = generic implicit GEMM per step, 1 = resident-A kernel per step (default), 3 = one launch
                            // for the whole recurrence when a sync workspace is given (measured slower, see twa_seq_kernel) (2 = per step with descriptor base
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

// =====================================================================================================================
// The whole recurrence in ONE launch (model_convlstm.py:364-377) - uavsal_set_option(7, 3); NOT the default, see below.
// The grid of a step (tile x N slice x sequence, at most one CTA per SM so that all of them are resident) stays on the SMs
// for all t_steps.  A CTA may start step t as soon as the tiles its haloed box touches have published h_{t-1}: every CTA
// that finishes a step bumps the `ready` counter of each of its (at most nine) neighbour tiles, itself included, with a
// gpu-scope release; the thread that issues the TMA loads of A polls its own tile's counter (one address) with acquire
// loads.  There is no grid-wide barrier.  Barriers, TMEM, tensor maps are set up once; the epilogue stages its transposes in
// the A buffers (idle between acc_full and the next step's loads); the weight ring is refilled for step t+1 once step t's tile
// is published (earlier, the 4-stage burst of every CTA stretched the epilogue's loads, stores and the release by 5-8 us).
// Results are bit-identical to the per-step kernel (same MMA order, same epilogue arithmetic).
// No griddepcontrol.launch_dependents: a dependent kernel spinning on SMs this grid still needs would deadlock it.
//
// Measured (two sequences, 45x80, 256 ch; per-phase globaltimer trace of one CTA, DBG_TRACE; profiles/r02_twa_trace.txt):
//   flag seen -> A landed 0.6 us | 432 MMAs 25.2 us (= 94 % of the measured dense bf16 peak per SM: the step's floor on 120 SMs)
//   | epilogue 10 us (row-per-lane operand loads: 32 cache lines per load instruction) | release 2.5 us | counter seen 1.5-4.5 us
// = 43 us per step against 36 us for one launch per step: with programmatic dependent launch the kernel boundary (drain +
// griddepcontrol.wait) costs ~1 us, LESS than the in-kernel release/acquire chain, so the per-step launches stay the default.
struct TwaSeqArgs {
    int H, W, C, bn;
    int tiles_x, tiles_y, ncb, nsplit;
    int t_steps;
    int bk_tap_stride, bk_off;
    const float* gx;
    Act x, h0;
    ActW seq;
    int tmem_cols, bstages, dbg;
    int* ready;                       // [batch][tiles] zeroed before the launch
};

constexpr int DBG_TRACE = 1 << 22;     // development: CTA (11,0,0) prints per-step phase timestamps (globaltimer, ns)
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// 8 consecutive channels written earlier in THIS kernel (possibly by other threads of the CTA): L2-coherent loads, not .nc
__device__ __forceinline__ void load8_cg(const uint16_t* hi, int64_t plane, float v[8]) {
    uint4 h, l;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(h.x), "=r"(h.y), "=r"(h.z), "=r"(h.w) : "l"(hi) : "memory");
    unpack2(h.x, v[0], v[1]); unpack2(h.y, v[2], v[3]); unpack2(h.z, v[4], v[5]); unpack2(h.w, v[6], v[7]);
    if (plane) {
        asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(l.x), "=r"(l.y), "=r"(l.z), "=r"(l.w) : "l"(hi + plane) : "memory");
        float t[8];
        unpack2(l.x, t[0], t[1]); unpack2(l.y, t[2], t[3]); unpack2(l.z, t[4], t[5]); unpack2(l.w, t[6], t[7]);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
}

constexpr int kSeqThreads = kThreads2 + 32;        // + one warp that gates and issues the h_{t-1} loads
constexpr int kSeqAWarp = kThreads2 / 32;

template <int TERMS>
__global__ void __launch_bounds__(kSeqThreads, 1) twa_seq_kernel(const __grid_constant__ CUtensorMap tmA0,    // h0   [batch] images
                                                                const __grid_constant__ CUtensorMap tmA,     // seq  [batch * t_steps] images
                                                                const __grid_constant__ CUtensorMap tmB, const TwaSeqArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int NPL = TERMS == 3 ? 2 : 1;
    const uint32_t a_stage = NPL * kTwAPlane;
    const uint32_t b_plane = (uint32_t)g.bn * 128, b_stage = NPL * b_plane;
    uint8_t* abuf = smem;                                                     // [2][a_stage]; the epilogue's staging area between steps
    uint8_t* bbuf = abuf + 2 * a_stage;                                       // [bstages][b_stage]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bbuf + g.bstages * b_stage);
    uint64_t* a_full = bars;                   // [2]
    uint64_t* a_empty = bars + 2;              // [2]
    uint64_t* b_full = bars + 4;               // [kTwBStagesMax]
    uint64_t* b_empty = b_full + kTwBStagesMax;
    uint64_t* acc_full = b_empty + kTwBStagesMax;
    uint64_t* epi_done = acc_full + 1;         // the step's tile is published: the weight ring may start on the next step
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 1);
    unsigned long long* trace = reinterpret_cast<unsigned long long*>(tmem_slot + 2);      // [16 steps][16]
    const bool tracing = (g.dbg & DBG_TRACE) && blockIdx.x == 11 && blockIdx.y == 0 && blockIdx.z == 0;
#define TWA_TRACE(step, slot) do { if (tracing && (step) < 16) trace[(step) * 16 + (slot)] = gtime(); } while (0)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, n0 = blockIdx.y * g.bn, seq_b = blockIdx.z;
    const int ty = tile / g.tiles_x, tx = tile % g.tiles_x;
    const int y0 = ty * kTwTH, x0 = tx * kTwTW;
    const int ntiles = g.tiles_x * g.tiles_y;
    int* const ready = g.ready + seq_b * ntiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < g.bstages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_init(epi_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                                               // gx / x / h0 come from earlier kernels

    if (warp == 0) {
        if (lane == 0) {
            // ===================== weights =====================
            const int total = g.t_steps * g.ncb * 9, per_step = g.ncb * 9;
            int cb = 0, tap = 0, t = 0;
            for (int kbB = 0; kbB < total; ++kbB) {
                const int s = kbB % g.bstages;
                if (kbB == (t + 1) * per_step) { mbar_wait(epi_done, t & 1); ++t; }
                mbar_wait(b_empty + s, ((kbB / g.bstages) & 1) ^ 1);
                mbar_expect_tx(b_full + s, b_stage);
#pragma unroll
                for (int p = 0; p < NPL; ++p)
                    tma_load_3d(&tmB, b_full + s, bbuf + s * b_stage + p * b_plane, tap * g.bk_tap_stride + g.bk_off + cb * 64, n0, p);
                if (++tap == 9) { tap = 0; if (++cb == g.ncb) cb = 0; }
            }
        }
    } else if (warp == kSeqAWarp) {
        if (lane == 0) {
            // ===================== h_{t-1}: haloed boxes, gated by the neighbour tiles' progress =====================
            int nnb = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    nnb += (ty + dy >= 0 && ty + dy < g.tiles_y && tx + dx >= 0 && tx + dx < g.tiles_x) ? 1 : 0;
            int acount = 0;
            for (int t = 0; t < g.t_steps; ++t) {
                if (t > 0) {
                    const int want = t * g.nsplit * nnb;
                    for (uint32_t it = 0; ld_acquire_gpu(ready + tile) < want; ++it) {
                        __nanosleep(40);
                        if (it > (1u << 24)) __trap();                        // a protocol bug must not hang the GPU
                    }
                    fence_proxy_async_all();                                  // the boxes were written with generic-proxy stores
                }
                TWA_TRACE(t, 0);
                for (int cb = 0; cb < g.ncb; ++cb, ++acount) {
                    const int sa = acount & 1;
                    mbar_wait(a_empty + sa, ((acount >> 1) & 1) ^ 1);
                    mbar_expect_tx(a_full + sa, NPL * kTwAPlaneBytes);
#pragma unroll
                    for (int p = 0; p < NPL; ++p) {
                        if (t == 0) tma_load_5d(&tmA0, a_full + sa, abuf + sa * a_stage + p * kTwAPlane, cb * 64, x0 - 1, y0 - 1, seq_b, p);
                        else        tma_load_5d(&tmA, a_full + sa, abuf + sa * a_stage + p * kTwAPlane, cb * 64, x0 - 1, y0 - 1,
                                                seq_b * g.t_steps + t - 1, p);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc(g.bn);
        int kbB = 0, acount = 0;
        for (int t = 0; t < g.t_steps; ++t) {
            for (int cb = 0; cb < g.ncb; ++cb, ++acount) {
                const int sa = acount & 1;
                mbar_wait(a_full + sa, (acount >> 1) & 1);
                if (lane == 0 && cb == 0) TWA_TRACE(t, 1);
                for (int tap = 0; tap < 9; ++tap, ++kbB) {
                    const int s = kbB % g.bstages;
                    mbar_wait(b_full + s, (kbB / g.bstages) & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        if (cb == 0 && tap == 0) TWA_TRACE(t, 2);
                        if (cb == g.ncb - 1 && tap == 8) TWA_TRACE(t, 3);
                        const uint32_t a_hi = smem_u32(abuf + sa * a_stage) + (uint32_t)((tap / 3) * kTwIW + tap % 3) * 128;
                        const uint32_t b_hi = smem_u32(bbuf + s * b_stage);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            if (g.dbg & DBG_NO_MMA) break;
                            const uint64_t dah = umma_desc_shift(a_hi + k * 32, kTwIW * 128, 0);
                            const uint64_t dbh = umma_desc(b_hi + k * 32);
                            umma_bf16(tmem_base, dah, dbh, idesc, (cb | tap | k) ? 1u : 0u);
                            if (TERMS == 3) {
                                const uint64_t dal = umma_desc_shift(a_hi + kTwAPlane + k * 32, kTwIW * 128, 0);
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
        }
    } else {
        // ===================== epilogue: gate + blend (model_convlstm.py:283,290), then publish the tile =====================
        const int ew = warp - 2, q = warp & 3, sub = (ew >> 2) * 16;
        const uint32_t wst = smem_u32(abuf) + ew * 2048;                      // both A buffers are idle between acc_full and the next step's loads
        auto pix_of = [&](int rr) -> int64_t {
            const int y = y0 + (rr >> 3), x = x0 + (rr & 7);
            return (y < g.H && x < g.W) ? (int64_t)y * g.W + x : -1;
        };
        const int r = q * 32 + lane;
        const int64_t pix = pix_of(r);
        const int64_t hw = (int64_t)g.H * g.W;
        const bool rowlive = pix >= 0 && !(g.dbg & DBG_NO_STORE);
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const int et = threadIdx.x - 64;                                      // 0 .. 511
        const int nch = g.bn >> 6;
        for (int t = 0; t < g.t_steps; ++t) {
            const int64_t oimg = (int64_t)seq_b * g.t_steps + t;
            const int64_t orow = oimg * hw + pix;
            const uint16_t* hsrc = t == 0 ? g.h0.p : g.seq.p;
            const int64_t hplane = t == 0 ? g.h0.plane : g.seq.plane;
            const int hld = t == 0 ? g.h0.ld : g.seq.ld;
            const int64_t hrow = (t == 0 ? (int64_t)seq_b : oimg - 1) * hw + pix;
            float gxv[16], xv[16], hv[16];
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
                    load8_cg(hsrc + hrow * hld + n + half * 8, hplane, hv + half * 8);
                }
            };
            if (rowlive) {
                fetch(n0 + sub);
                for (int ch = 1; ch < nch; ++ch) {                           // later chunks: DRAM -> L2 while the MMAs run
                    const int n = n0 + ch * 64 + sub;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(g.gx + orow * g.C + n));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(g.x.p + orow * g.x.ld + n));
                    if (g.x.plane) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.x.p + g.x.plane + orow * g.x.ld + n));
                }
            }
            mbar_wait(acc_full, t & 1);
            tc_fence_after();
            if (et == 0) TWA_TRACE(t, 4);
            for (int ch = 0; ch < nch; ++ch) {
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
                if (et == 0) TWA_TRACE(t, 8 + ch * 3);
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
                if (et == 0) TWA_TRACE(t, 9 + ch * 3);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int row = 16 * i + (lane >> 1), c = lane & 1;
                    const int off = row * 32 + ((c ^ ((row >> 2) & 1)) << 4);
                    const uint4 hv4 = lds128(wst + off);
                    const uint4 lv4 = lds128(wst + 1024 + off);
                    const int64_t px = pix_of(q * 32 + row);
                    if (px >= 0 && !(g.dbg & DBG_NO_STORE)) {
                        uint16_t* dst = g.seq.p + (oimg * hw + px) * g.seq.ld + n + c * 8;
                        *reinterpret_cast<uint4*>(dst) = hv4;
                        if (g.seq.plane) *reinterpret_cast<uint4*>(dst + g.seq.plane) = lv4;
                    }
                }
                if (et == 0) TWA_TRACE(t, 10 + ch * 3);
            }
            // the tile's slice of h_t is written: order the TMEM reads / staging accesses before the next step's MMAs and TMA
            // writes, then tell every tile whose haloed box overlaps this one
            tc_fence_before();
            fence_async_smem();
            named_bar_sync(1, kEpiThreads);
            if (et == 0) TWA_TRACE(t, 6);
            if (et < 9) {
                const int ny = ty + et / 3 - 1, nx = tx + et % 3 - 1;
                if (ny >= 0 && ny < g.tiles_y && nx >= 0 && nx < g.tiles_x)
                    red_release_gpu_add(ready + ny * g.tiles_x + nx, 1);     // release: cumulative over the stores ordered by the barrier
            }
            if (et == 0) { mbar_arrive(epi_done); TWA_TRACE(t, 7); }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    if (tracing && threadIdx.x == 0) {
        // slots: 0 counter seen, 1 A(cb0) landed, 2 first MMA issued, 3 last k-block issued, 4 acc_full seen, 8/11 chunk math done,
        // 9/12 staged, 10/13 stores issued, 6 after the epilogue barrier, 7 counters bumped
        for (int t = 1; t < min(g.t_steps, 16); ++t) {
            const unsigned long long b = trace[t * 16 + 0];
            printf("step %2d: A %5lld | mma0 %5lld | mmaN %5lld | acc %5lld | c0 math %5lld staged %5lld stored %5lld | c1 math %5lld staged %5lld stored %5lld"
                   " | bar %5lld | signal %5lld | prev signal->flag %5lld ns\n", t,
                   (long long)(trace[t * 16 + 1] - b), (long long)(trace[t * 16 + 2] - b), (long long)(trace[t * 16 + 3] - b), (long long)(trace[t * 16 + 4] - b),
                   (long long)(trace[t * 16 + 8] - b), (long long)(trace[t * 16 + 9] - b), (long long)(trace[t * 16 + 10] - b),
                   (long long)(trace[t * 16 + 11] - b), (long long)(trace[t * 16 + 12] - b), (long long)(trace[t * 16 + 13] - b),
                   (long long)(trace[t * 16 + 6] - b), (long long)(trace[t * 16 + 7] - b), (long long)(b - trace[(t - 1) * 16 + 7]));
        }
    }
#undef TWA_TRACE
}

size_t twa_sync_bytes(int batch, int H, int W) {
    return (size_t)batch * div_up(W, kTwTW) * div_up(H, kTwTH) * sizeof(int);
}

// all t_steps of `batch` sequences in one launch; returns UAVSAL_ENOTSUP (nothing launched) when the step grid cannot be fully resident
int twa_sequence_persistent(Act x, Act h0, ActW seq, int batch, int t_steps, int H, int W, int c, const uint16_t* wgt, int wk_total, int wk_off,
                            const float* gx, int terms, int* ready, cudaStream_t s, int dbg) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    TwaSeqArgs g{};
    const uint32_t npl = terms == 3 ? 2 : 1;
    g.H = H; g.W = W; g.C = c; g.t_steps = t_steps;
    g.tiles_x = div_up(W, kTwTW); g.tiles_y = div_up(H, kTwTH); g.ncb = c / 64;
    const int tiles = g.tiles_x * g.tiles_y;
    g.bn = 64;
    while (g.bn < 256 && c % (2 * g.bn) == 0 && (int64_t)tiles * batch * (c / g.bn) > sms) g.bn *= 2;
    g.nsplit = c / g.bn;
    if (c % g.bn || (int64_t)tiles * batch * g.nsplit > sms) return UAVSAL_ENOTSUP;
    g.bk_tap_stride = wk_total; g.bk_off = wk_off;
    g.gx = gx; g.x = x; g.h0 = h0; g.seq = seq;
    g.tmem_cols = g.bn < 32 ? 32 : g.bn;
    g.dbg = dbg;
    g.ready = ready;
    CUtensorMap tA0, tA, tB;
    {
        const uint64_t dims[5] = {(uint64_t)c, (uint64_t)W, (uint64_t)H, (uint64_t)batch, h0.plane ? 2u : 1u};
        const uint64_t row = (uint64_t)h0.ld * 2;
        const uint64_t str[4] = {row, row * W, row * W * H, h0.plane ? (uint64_t)h0.plane * 2 : row * W * H * (uint64_t)batch};
        const uint32_t box[5] = {64, (uint32_t)kTwIW, (uint32_t)kTwIH, 1, 1};
        int rc = tc_encode(&tA0, h0.p, 5, dims, str, box, "twa_sequence A (h0)", 1);
        if (rc) return rc;
    }
    {
        const uint64_t nimg = (uint64_t)batch * t_steps;
        const uint64_t dims[5] = {(uint64_t)c, (uint64_t)W, (uint64_t)H, nimg, seq.plane ? 2u : 1u};
        const uint64_t row = (uint64_t)seq.ld * 2;
        const uint64_t str[4] = {row, row * W, row * W * H, seq.plane ? (uint64_t)seq.plane * 2 : row * W * H * nimg};
        const uint32_t box[5] = {64, (uint32_t)kTwIW, (uint32_t)kTwIH, 1, 1};
        int rc = tc_encode(&tA, seq.p, 5, dims, str, box, "twa_sequence A (h_{t-1})", 1);
        if (rc) return rc;
    }
    {
        const uint64_t kpad = 9ull * wk_total;
        const uint64_t dims[3] = {kpad, (uint64_t)c, 2};
        const uint64_t str[2] = {kpad * 2, kpad * 2 * (uint64_t)c};
        const uint32_t box[3] = {kBK, (uint32_t)g.bn, 1};
        int rc = tc_encode(&tB, wgt, 3, dims, str, box, "twa_sequence B (weights)", 1);
        if (rc) return rc;
    }
    g.bstages = (int)((220u * 1024u - 2u * npl * kTwAPlane) / (npl * (uint32_t)g.bn * 128u));
    if (g.bstages > kTwBStagesMax) g.bstages = kTwBStagesMax;
    const size_t smem = 2 * (size_t)npl * kTwAPlane + (size_t)g.bstages * npl * g.bn * 128 + 256 + 1024 + 2048;     // + the trace slots
    cudaError_t e = cudaMemsetAsync(ready, 0, twa_sync_bytes(batch, H, W), s);
    if (e != cudaSuccess) { set_error("twa_sequence: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    const dim3 grid(tiles, g.nsplit, batch);
    if (terms == 3) {
        static bool attr = false;
        if (!attr) {
            e = cudaFuncSetAttribute(twa_seq_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("twa_sequence: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            attr = true;
        }
        e = launch_k(twa_seq_kernel<3>, grid, dim3(kSeqThreads), smem, s, 1, tA0, tA, tB, g);
    } else {
        static bool attr = false;
        if (!attr) {
            e = cudaFuncSetAttribute(twa_seq_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("twa_sequence: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            attr = true;
        }
        e = launch_k(twa_seq_kernel<1>, grid, dim3(kSeqThreads), smem, s, 1, tA0, tA, tB, g);
    }
    if (e != cudaSuccess) { set_error("twa_sequence: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("twa_sequence(persistent)");
}

}  // namespace uavsal

// Persistent tcgen05 GEMM / implicit-GEMM 3x3 (version 2 of gemm_tc_kernel).
//
// One CTA per SM (or one CTA pair per TPC) walks a static round-robin list of output tiles (n fastest, so CTAs running at
// the same time share A tiles in L2).  Three asynchronous pipelines overlap across tiles:
//   TMA producer  -> smem ring (full/empty mbarriers, runs ahead into the next tile's k-blocks, L2 prefetch of A)
//   MMA issuer    -> TWO TMEM accumulators (acc_full/acc_empty mbarriers): tile i+1 accumulates while
//   epilogue      -> tile i is drained: tcgen05.ld -> bias/ReLU6/residual/sigmoid | TWA blend | LSTM cell -> hi/lo split (or fp32)
//                    -> per-warp transpose buffer -> coalesced global stores (clipped at M/N/image edges).
// (The output tensor map parameter tmO is unused: TMA stores queued behind the producer's loads and were replaced by the
//  LSU copy-out; the parameter is kept so the launch signature stays stable.)
#pragma once
#include "tc_common.cuh"

namespace uavsal {

constexpr uint32_t kOutPlaneBytes = kBM * 64 * 2;          // one 128 x 64 bf16 box
constexpr uint32_t kOutStageBytes = 2 * kOutPlaneBytes;    // hi + lo
constexpr int kEpiWarps = 16;
constexpr int kPF = 6;                                     // L2 prefetch distance of the A operand, in k-blocks
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads2 = 64 + kEpiThreads;                // producer warp + MMA warp + epilogue warps

// CL = CTAs per cluster (1 | 2).  With CL = 2 the two CTAs of a cluster (the two SMs of a TPC) run cta_group::2 MMAs on
// a 256-row tile: each CTA stages its own 128 rows of A and HALF of the B tile, keeps its 128 accumulator lanes in its own
// TMEM and drains them with its own epilogue warps; the even CTA issues the MMAs and its commits arrive on both CTAs'
// barriers.  Measured motivation: a single-CTA 128x256 tile with the 3-term split reads 36 KB of operands from shared
// memory per k-step and fills 96 KB per k-block - more than the 128 B/clk/SM the shared memory delivers - and only two
// 96 KB stages fit.  The pair reads 24 KB per k-step per SM, fills 64 KB per k-block and gets three stages.
template <int MODE, int EPI, int TERMS, int CL>
__global__ void __launch_bounds__(kThreads2, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                               const __grid_constant__ CUtensorMap tmA1,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmO, const TcArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int NPL = TERMS == 3 ? 2 : 1;
    constexpr bool PAIR = CL == 2;
    const uint32_t b_bytes = (uint32_t)(g.bn / CL) * kBK * 2;                 // B rows staged by THIS CTA
    const uint32_t stage_bytes = NPL * (kABytes + b_bytes);
    uint8_t* ostage = smem + (size_t)g.stages * stage_bytes;                  // 1024-aligned (stage_bytes % 1024 == 0)
    uint64_t* full = reinterpret_cast<uint64_t*>(ostage + kOutStageBytes);
    uint64_t* empty = full + g.stages;
    uint64_t* acc_full = empty + g.stages;                                    // [2]
    uint64_t* acc_empty = acc_full + 2;                                       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    const int cta_first = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;   // first (pair-)tile of this CTA
    const int cta_step = (int)(gridDim.x / CL);
    const int num_work = CL > 1 ? ((g.tiles_m + CL - 1) / CL) * g.tiles_n : g.num_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_empty + b, CL * kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                                           // peers' barriers are initialised
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                               // everything above overlapped the previous kernel's tail; operands are its outputs

    // tile decomposition shared by all roles (work item t -> 128-row tile of this CTA; may be == tiles_m for the odd one out)
    auto tile_coords = [&](int t, int& tile_m, int& n0, int& bidx, int& y0, int& x0) {
        const int wm = t / g.tiles_n;
        tile_m = CL > 1 ? wm * CL + (int)crank : wm;
        n0 = (t - wm * g.tiles_n) * g.bn;
        bidx = 0; y0 = 0; x0 = 0;
        if (MODE == MODE_CONV) {
            const int per_img = g.tiles_x * g.tiles_y;
            bidx = tile_m / per_img;
            const int r = tile_m - bidx * per_img;
            y0 = (r / g.tiles_x) * g.TH;
            x0 = (r % g.tiles_x) * g.TW;
        }
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t kc = 0;                                                  // k-block counter across tiles
            for (int t = cta_first; t < num_work; t += cta_step) {
                int tile_m, n0, bidx, y0, x0;
                tile_coords(t, tile_m, n0, bidx, y0, x0);
                for (int kb = 0; kb < g.num_kb; ++kb, ++kc) {
                    if (MODE == MODE_PW && g.num_kb >= 8) {
                        // long-K GEMMs stream A from HBM with only 2-3 smem stages in flight: pull the tile kPF k-blocks
                        // ahead (this tile's, then the next tile's) into L2 so the staged loads see L2 latency
                        int pk = kb + kPF, pt = tile_m;
                        if (pk >= g.num_kb) {
                            int nm, nn, nb, ny, nx;
                            pk -= g.num_kb;
                            if (t + cta_step < num_work) { tile_coords(t + cta_step, nm, nn, nb, ny, nx); pt = nm; } else pt = -1;
                        }
                        if (pt >= 0) {
#pragma unroll
                            for (int p = 0; p < NPL; ++p) tma_prefetch_3d(&tmA0, pk * kBK, pt * kBM, p);
                        }
                    }
                    const int s = kc % g.stages;
                    mbar_wait(empty + s, ((kc / g.stages) & 1) ^ 1);
                    uint8_t* sa = smem + (size_t)s * stage_bytes;
                    uint8_t* sb = sa + NPL * kABytes;
                    const bool no_a = g.flags & DBG_NO_A, no_b = g.flags & DBG_NO_B;
                    const uint32_t tx = stage_bytes - (no_a ? NPL * kABytes : 0) - (no_b ? NPL * b_bytes : 0);
                    // pair: all bytes of both CTAs complete on the even CTA's barrier, which alone is waited on (by the issuer)
                    if (!PAIR || crank == 0) mbar_expect_tx(full + s, CL * tx);
                    const uint32_t fbar = PAIR ? mapa_u32(smem_u32(full + s), 0) : 0;
                    int bk = kb * kBK;
                    if (MODE == MODE_CONV) {
                        const int tap = kb / g.kb_per_tap;
                        bk = tap * g.bk_tap_stride + g.bk_off + (kb - tap * g.kb_per_tap) * kBK;
                    }
                    if (no_a) {
                    } else if (MODE == MODE_PW) {
#pragma unroll
                        for (int p = 0; p < NPL; ++p) {
                            if (PAIR) tma_load_3d_pair(&tmA0, fbar, sa + p * kABytes, kb * kBK, tile_m * kBM, p);
                            else tma_load_3d(&tmA0, full + s, sa + p * kABytes, kb * kBK, tile_m * kBM, p);
                        }
                    } else {
                        const int tap = kb / g.kb_per_tap;
                        const int r = kb - tap * g.kb_per_tap;
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        const bool src0 = r < g.kb_src0;
                        const CUtensorMap* tm = src0 ? &tmA0 : &tmA1;
                        const int cch = (src0 ? r : r - g.kb_src0) * kBK;
                        const int img = src0 ? bidx * g.a0_mul + g.a0_off : bidx * g.a1_mul + g.a1_off;
#pragma unroll
                        for (int p = 0; p < NPL; ++p) {
                            if (PAIR) tma_load_5d_pair(tm, fbar, sa + p * kABytes, cch, x0 + dx, y0 + dy, img, p);
                            else tma_load_5d(tm, full + s, sa + p * kABytes, cch, x0 + dx, y0 + dy, img, p);
                        }
                    }
                    if (!no_b) {
#pragma unroll
                        for (int p = 0; p < NPL; ++p) {
                            if (PAIR) tma_load_3d_pair(&tmB, fbar, sb + p * b_bytes, bk, n0 + (int)crank * (g.bn / CL), p);   // my half of the B tile
                            else tma_load_3d(&tmB, full + s, sb + p * b_bytes, bk, n0, p);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair: the even CTA only) =====================
        if (!PAIR || crank == 0) {
        const uint32_t idesc = PAIR ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24))
                                    : umma_idesc(g.bn);
        uint32_t kc = 0;
        int it = 0;
        for (int t = cta_first; t < num_work; t += cta_step, ++it) {
            const int buf = it & 1;
            // the epilogue warps (of both CTAs) have drained this accumulator
            if (PAIR) mbar_wait_cluster(acc_empty + buf, ((it >> 1) & 1) ^ 1);
            else mbar_wait(acc_empty + buf, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * g.bn);
            for (int kb = 0; kb < g.num_kb; ++kb, ++kc) {
                const int s = kc % g.stages;
                mbar_wait(full + s, (kc / g.stages) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t b_hi = a_hi + NPL * kABytes;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        if (g.flags & DBG_NO_MMA) break;
                        const uint64_t dah = umma_desc(a_hi + k * 32);
                        const uint64_t dbh = umma_desc(b_hi + k * 32);
                        if (PAIR) umma_bf16_pair(d_tmem, dah, dbh, idesc, (kb | k) ? 1u : 0u);
                        else umma_bf16(d_tmem, dah, dbh, idesc, (kb | k) ? 1u : 0u);
                        if (TERMS == 3) {
                            const uint64_t dal = umma_desc(a_hi + kABytes + k * 32);
                            const uint64_t dbl = umma_desc(b_hi + b_bytes + k * 32);
                            if (PAIR) { umma_bf16_pair(d_tmem, dah, dbl, idesc, 1u); umma_bf16_pair(d_tmem, dal, dbh, idesc, 1u); }
                            else { umma_bf16(d_tmem, dah, dbl, idesc, 1u); umma_bf16(d_tmem, dal, dbh, idesc, 1u); }
                        }
                    }
                    if (PAIR) umma_commit_pair(empty + s); else umma_commit(empty + s);
                    if (kb == g.num_kb - 1) { if (PAIR) umma_commit_pair(acc_full + buf); else umma_commit(acc_full + buf); }
                }
                __syncwarp();
            }
        }
        }
    } else {
        // ===================== epilogue (warps 2..17) =====================
        // A warp may only touch TMEM lanes 32*(warp%4)..+31, so the four warps sharing a quadrant split every 64-column
        // chunk into 16-column pieces.  Each warp works on its own: tcgen05.ld -> bias/ReLU6/residual/sigmoid | TWA blend ->
        // (hi/lo split) -> a PRIVATE 2 KiB transpose buffer -> coalesced global stores.  There is no block-wide barrier
        // in the loop (the first version's two 512-thread barriers per chunk left the epilogue latency-bound: 80 us of a
        // 180 us 256->1536 GEMM with MMA, loads and stores all disabled), so the warps of one scheduler hide each other's
        // TMEM / shared-memory / store latency.
        const int ew = warp - 2;
        const int q = warp & 3;
        const int sub = (ew >> 2) * 16;                                       // column offset inside a 64-column chunk
        const int r = q * 32 + lane;                                          // tile row = TMEM lane
        const uint32_t wst = smem_u32(ostage) + ew * 2048;                    // 32 rows x 16 columns: fp32, or hi | lo bf16
        constexpr bool kStd = EPI == EPI_STD || EPI == EPI_RES || EPI == EPI_Q16;
        constexpr bool kRes = EPI == EPI_RES;                                  // residual: bf16 hi/lo output, no sigmoid (host-checked)
        constexpr bool kQ16 = EPI == EPI_Q16;                                  // ReLU6 output as 16-bit fixed point rows (host-checked flags)
        const bool f32out = EPI == EPI_STD && (g.flags & UAVSAL_F_OUT_F32);
        const bool do_store = !(g.flags & DBG_NO_STORE);
        int it = 0;
        for (int t = cta_first; t < num_work; t += cta_step, ++it) {
            int tile_m, n0, bidx, y0, x0;
            tile_coords(t, tile_m, n0, bidx, y0, x0);
            const int buf = it & 1;
            int img_out = 0;
            if (MODE == MODE_CONV) img_out = bidx * g.out_mul + g.out_off;
            // global row of tile row rr (-1: outside the tensor)
            auto grow_of = [&](int rr) -> int64_t {
                if (MODE == MODE_PW) {
                    const int64_t gr = (int64_t)tile_m * kBM + rr;
                    return gr < g.M ? gr : -1;
                }
                const int y = y0 + rr / g.TW, x = x0 + rr % g.TW;
                if (y >= g.H || x >= g.W || tile_m >= g.tiles_m) return -1;
                return ((int64_t)img_out * g.H + y) * g.W + x;
            };
            const int64_t orow = grow_of(r);
            const bool rvalid = orow >= 0;
            int64_t hrow = 0;
            if (MODE == MODE_CONV && rvalid)
                hrow = orow + (int64_t)((bidx * g.a1_mul + g.a1_off) - img_out) * g.H * g.W;
            const int nchunks = (g.bn + 63) >> 6;
            // EPI_RES: this lane's two (row, 8-column) cells of a piece; the residual loads of piece ch+1 are issued before piece
            // ch is processed (those of piece 0 before the accumulator is waited for), so their latency is hidden
            uint4 res_h[2], res_l[2], nres_h[2], nres_l[2];
            int64_t gro[2] = {-1, -1};
            auto issue_res = [&](int ch, uint4 (&h)[2], uint4 (&l)[2]) {
                const int col = n0 + ch * 64 + sub + (lane & 1) * 8;
                if (sub < min(64, g.bn - ch * 64) && col < g.N) {
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (gro[i] >= 0) {
                            const uint16_t* a = g.res.p + gro[i] * g.res.ld + col;
                            h[i] = __ldg(reinterpret_cast<const uint4*>(a));
                            if (g.res.plane) l[i] = __ldg(reinterpret_cast<const uint4*>(a + g.res.plane));
                        }
                }
            };
            if (kRes) {
                gro[0] = grow_of(q * 32 + (lane >> 1));
                gro[1] = grow_of(q * 32 + 16 + (lane >> 1));
                issue_res(0, nres_h, nres_l);
            }
            mbar_wait(acc_full + buf, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * g.bn);
            // the TMEM read of piece ch+1 is issued as soon as piece ch's values have left `raw` and waited for one piece later, so its
            // latency hides under the copy-out
            // (not in the residual instantiation: its prefetched residual cells already fill the register budget)
            constexpr bool kPipe = EPI != EPI_RES;
            uint32_t raw[16];
            if (kPipe && sub < min(64, g.bn)) { __syncwarp(); tmem_ld16_issue(trow + sub, raw); }
            for (int ch = 0; ch < nchunks; ++ch) {
                const int ncols = min(64, g.bn - ch * 64);                    // multiple of 16
                if (kRes) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) { res_h[i] = nres_h[i]; res_l[i] = nres_l[i]; }
                    if (ch + 1 < nchunks) issue_res(ch + 1, nres_h, nres_l);
                }
                if (sub < ncols) {
                    const int n = n0 + ch * 64 + sub;
                    const bool live = rvalid && n < g.N;
                    const bool second = n + 8 < g.N;
                    // the bias of this 16-column piece is fetched BEFORE the TMEM load is waited for: issued after it, the adds sat on
                    // the global-load latency once per piece (ncu, round 2: 13.7 % of the epilogue warps' stall samples)
                    float4 bv[4] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
                    if (EPI != EPI_RAW && g.bias && live) {
                        const float4* bp = reinterpret_cast<const float4*>(g.bias + n);
                        bv[0] = __ldg(bp); bv[1] = __ldg(bp + 1);
                        if (second) { bv[2] = __ldg(bp + 2); bv[3] = __ldg(bp + 3); }
                    }
                    if (!kPipe) { __syncwarp(); tmem_ld16_issue(trow + ch * 64 + sub, raw); }
                    tmem_ld16_wait(raw);
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
                    if (kPipe && ch + 1 < nchunks && sub < min(64, g.bn - (ch + 1) * 64)) { __syncwarp(); tmem_ld16_issue(trow + (ch + 1) * 64 + sub, raw); }
                    if (EPI == EPI_RAW) {
                        if (live) {
                            float4* o = reinterpret_cast<float4*>(g.raw_out + orow * g.N + n);
                            o[0] = make_float4(v[0], v[1], v[2], v[3]);
                            o[1] = make_float4(v[4], v[5], v[6], v[7]);
                            if (second) {
                                o[2] = make_float4(v[8], v[9], v[10], v[11]);
                                o[3] = make_float4(v[12], v[13], v[14], v[15]);
                            }
                        }
                    } else if (EPI == EPI_LSTM) {
                        // interleaved gates (i,f,o,g) x 4 channels per 16-column piece; c' = s(f)c + s(i)tanh(g); h' = s(o)tanh(c')
                        // (model_convlstm.py:117-124).  c state fp32 in place; h goes straight to the output sequence.
                        if (live) {
                            if (g.bias) {
#pragma unroll
                                for (int j4 = 0; j4 < 4; ++j4) {
                                    v[j4 * 4 + 0] += bv[j4].x; v[j4 * 4 + 1] += bv[j4].y; v[j4 * 4 + 2] += bv[j4].z; v[j4 * 4 + 3] += bv[j4].w;
                                }
                            }
                            const int nch = g.N >> 2, chn = n >> 2;
                            const int64_t crow = orow - (int64_t)(img_out - bidx) * g.H * g.W;
                            float* cs = g.c_state + crow * nch + chn;
                            const int cnt = second ? 4 : 2;
                            float hout[4];
                            for (int j = 0; j < cnt; ++j) {
                                const float gi = sigmoid_acc(v[4 * j + 0]), gf = sigmoid_acc(v[4 * j + 1]);
                                const float go = sigmoid_acc(v[4 * j + 2]), gg = tanhf(v[4 * j + 3]);
                                const float cn = gf * cs[j] + gi * gg;
                                cs[j] = cn;
                                hout[j] = go * tanhf(cn);
                            }
                            if (do_store) {
                                if (second) store4(g.out.p + orow * g.out.ld + chn, g.out.plane, hout);
                                else { store1(g.out.p + orow * g.out.ld + chn, g.out.plane, hout[0]); store1(g.out.p + orow * g.out.ld + chn + 1, g.out.plane, hout[1]); }
                            }
                        }
                    } else {
                        if (live) {
                            if (g.bias) {
#pragma unroll
                                for (int j4 = 0; j4 < 4; ++j4) {
                                    v[j4 * 4 + 0] += bv[j4].x; v[j4 * 4 + 1] += bv[j4].y; v[j4 * 4 + 2] += bv[j4].z; v[j4 * 4 + 3] += bv[j4].w;
                                }
                            }
                            if (kStd) {
                                if (g.flags & (UAVSAL_F_RELU6 | UAVSAL_F_RELU)) {
                                    const float cap = (g.flags & UAVSAL_F_RELU6) ? 6.f : 3.0e38f;      // ReLU6 (model.py:71) | ReLU (ResNet / VGG backbones)
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[j] = fminf(fmaxf(v[j], 0.f), cap);
                                }
                                if (!kRes && (g.flags & UAVSAL_F_RESIDUAL)) {
                                    float rr[8];
                                    load8(g.res.p + orow * g.res.ld + n, g.res.plane, rr);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[j] += rr[j];
                                    if (second) {
                                        load8(g.res.p + orow * g.res.ld + n + 8, g.res.plane, rr);
#pragma unroll
                                        for (int j = 0; j < 8; ++j) v[8 + j] += rr[j];
                                    }
                                }
                                if (g.flags & UAVSAL_F_SIGMOID) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[j] = sigmoid_acc(v[j]);
                                }
                            } else {   // EPI_TWA: h = i*x_t + (1-i)*h_{t-1}  (model_convlstm.py:283,290)
                                if (g.gx) {                                   // hoisted W_x * x_t half of the gate conv
                                    const float4* gp = reinterpret_cast<const float4*>(g.gx + orow * g.N + n);
#pragma unroll
                                    for (int j4 = 0; j4 < 4; ++j4) {
                                        if (j4 >= 2 && !second) break;
                                        const float4 b4 = __ldg(gp + j4);
                                        v[j4 * 4 + 0] += b4.x; v[j4 * 4 + 1] += b4.y; v[j4 * 4 + 2] += b4.z; v[j4 * 4 + 3] += b4.w;
                                    }
                                }
#pragma unroll
                                for (int half = 0; half < 2; ++half) {
                                    if (half == 1 && !second) break;
                                    float xv[8], hv[8];
                                    load8(g.x.p + orow * g.x.ld + n + half * 8, g.x.plane, xv);
                                    load8(g.hprev.p + hrow * g.hprev.ld + n + half * 8, g.hprev.plane, hv);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const float gi = sigmoid_acc(v[half * 8 + j]);
                                        v[half * 8 + j] = gi * xv[j] + (1.f - gi) * hv[j];
                                    }
                                }
                            }
                        }
                        __syncwarp();                                         // the previous piece has been copied out of wst
                        if (kQ16) {
                            // 32 rows x 16 columns of uint16: rows of 32 bytes at wst, chunk j of a row at j ^ ((row >> 2) & 1) (as the hi plane below)
#pragma unroll
                            for (int half = 0; half < 2; ++half)
                                sts128(wst + lane * 32 + ((half ^ ((lane >> 2) & 1)) << 4), q16_pack2(v[half * 8 + 0], v[half * 8 + 1]),
                                       q16_pack2(v[half * 8 + 2], v[half * 8 + 3]), q16_pack2(v[half * 8 + 4], v[half * 8 + 5]),
                                       q16_pack2(v[half * 8 + 6], v[half * 8 + 7]));
                            __syncwarp();
                            uint4 val[2];
#pragma unroll
                            for (int i = 0; i < 2; ++i) {                     // 16 rows x 32 contiguous bytes per instruction
                                const int row = 16 * i + (lane >> 1), c = lane & 1;
                                val[i] = lds128(wst + row * 32 + ((c ^ ((row >> 2) & 1)) << 4));
                            }
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const int row = 16 * i + (lane >> 1), c = lane & 1;
                                const int64_t gr = grow_of(q * 32 + row);
                                const int col = n + c * 8;
                                if (gr >= 0 && col < g.N && do_store) *reinterpret_cast<uint4*>(g.out.p + gr * g.out.ld + col) = val[i];
                            }
                        } else if (f32out) {
                            // (four 16-byte stores per lane straight from registers were measured: 897 -> 1346 us for 256 -> 1536, uncoalesced)
                            // row = lane: 64 bytes = four 16-byte chunks, chunk j stored at j ^ ((lane >> 1) & 3) (conflict-free)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                sts128(wst + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                       __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                            __syncwarp();
                            float* outf = reinterpret_cast<float*>(g.out.p);
                            // all four reads first, into registers of their own: with one register quad per read -> store pair the
                            // compiler chained them (ncu: LDS waiting for the previous STG to release its source, 22 % of the samples)
                            uint4 val[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {                     // 8 rows x 64 contiguous bytes per instruction
                                const int row = 8 * i + (lane >> 2), c = lane & 3;
                                val[i] = lds128(wst + row * 64 + ((c ^ ((row >> 1) & 3)) << 4));
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int row = 8 * i + (lane >> 2), c = lane & 3;
                                const int64_t gr = grow_of(q * 32 + row);
                                const int col = n + c * 4;
                                if (gr >= 0 && col < g.N && do_store) *reinterpret_cast<uint4*>(outf + gr * g.out.ld + col) = val[i];
                            }
                        } else if (kRes) {
                            // stage fp32 (layout as above); each lane then owns 8 columns of two rows: residual hi/lo arrive as
                            // 16 rows x 32 contiguous bytes per plane per instruction, like the stores
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                sts128(wst + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                       __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                            __syncwarp();
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const int row = 16 * i + (lane >> 1), c = lane & 1;
                                const int64_t gr = gro[i];
                                const int col = n + c * 8;
                                if (gr >= 0 && col < g.N) {
                                    const uint4 a0 = lds128(wst + row * 64 + (((2 * c) ^ ((row >> 1) & 3)) << 4));
                                    const uint4 a1 = lds128(wst + row * 64 + (((2 * c + 1) ^ ((row >> 1) & 3)) << 4));
                                    float rr[8], rl[8];
                                    unpack2(res_h[i].x, rr[0], rr[1]); unpack2(res_h[i].y, rr[2], rr[3]);
                                    unpack2(res_h[i].z, rr[4], rr[5]); unpack2(res_h[i].w, rr[6], rr[7]);
                                    if (g.res.plane) {
                                        unpack2(res_l[i].x, rl[0], rl[1]); unpack2(res_l[i].y, rl[2], rl[3]);
                                        unpack2(res_l[i].z, rl[4], rl[5]); unpack2(res_l[i].w, rl[6], rl[7]);
#pragma unroll
                                        for (int j = 0; j < 8; ++j) rr[j] += rl[j];
                                    }
                                    const float o[8] = {__uint_as_float(a0.x) + rr[0], __uint_as_float(a0.y) + rr[1], __uint_as_float(a0.z) + rr[2],
                                                        __uint_as_float(a0.w) + rr[3], __uint_as_float(a1.x) + rr[4], __uint_as_float(a1.y) + rr[5],
                                                        __uint_as_float(a1.z) + rr[6], __uint_as_float(a1.w) + rr[7]};
                                    uint32_t h[4], l[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) split2(o[2 * j], o[2 * j + 1], h[j], l[j]);
                                    if (do_store) {
                                        uint16_t* dst = g.out.p + gr * g.out.ld + col;
                                        *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                                        if (g.out.plane) *reinterpret_cast<uint4*>(dst + g.out.plane) = make_uint4(l[0], l[1], l[2], l[3]);
                                    }
                                }
                            }
                        } else {
                            // hi plane rows of 32 bytes at wst, lo plane at wst + 1024; chunk j of row at j ^ ((lane >> 2) & 1)
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
                            uint4 hv4[2], lv4[2];
#pragma unroll
                            for (int i = 0; i < 2; ++i) {                     // 16 rows x 32 contiguous bytes per plane per instruction
                                const int row = 16 * i + (lane >> 1), c = lane & 1;
                                const int off = row * 32 + ((c ^ ((row >> 2) & 1)) << 4);
                                hv4[i] = lds128(wst + off);
                                lv4[i] = lds128(wst + 1024 + off);
                            }
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const int row = 16 * i + (lane >> 1), c = lane & 1;
                                const int64_t gr = grow_of(q * 32 + row);
                                const int col = n + c * 8;
                                if (gr >= 0 && col < g.N && do_store) {
                                    uint16_t* dst = g.out.p + gr * g.out.ld + col;
                                    *reinterpret_cast<uint4*>(dst) = hv4[i];
                                    if (g.out.plane) *reinterpret_cast<uint4*>(dst + g.out.plane) = lv4[i];
                                }
                            }
                        }
                    }
                }
                if (ch == nchunks - 1) {                                      // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(acc_empty + buf), 0));
                        else mbar_arrive(acc_empty + buf);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // the peer may still multicast into this CTA's smem / arrive on its barriers
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    }
}

}  // namespace uavsal

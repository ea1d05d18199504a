// tcgen05 / TMA / TMEM GEMM and implicit-GEMM 3x3 for sm_100a.
//
//   D[128 x bn] (fp32, TMEM)  =  sum over k-blocks of  A[128 x 64] (smem, K-major, SW128) * B[bn x 64]^T
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> fused bias/ReLU6/residual/sigmoid | TWA blend | LSTM cell ->
// split-bf16 stores).  Operands are the arena's bf16 hi/lo planes; TERMS=3 issues the error-compensated
// product Ahi*Bhi + Ahi*Blo + Alo*Bhi into the same fp32 accumulator, TERMS=1 issues Ahi*Bhi only.
//
// MODE_PW  : A is a 2-D [M][K] matrix (NHWC rows x channels), tensor map (K, M, plane).
// MODE_CONV: A is gathered by TMA from NHWC images with a (64ch, TW, TH) box per filter tap; padding comes
//            from TMA out-of-bounds zero fill; two sources form the virtual concat [x, h] of the recurrences.
#include "tc_common.cuh"
#include "gemm_tc2.cuh"

namespace uavsal {

// ---- kernel ---------------------------------------------------------------------------------------
template <int MODE, int EPI, int TERMS>
__global__ void __launch_bounds__(kThreads) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                           const __grid_constant__ CUtensorMap tmA1,
                                                           const __grid_constant__ CUtensorMap tmB, const TcArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int NPL = TERMS == 3 ? 2 : 1;                 // planes staged per operand
    const uint32_t b_bytes = (uint32_t)g.bn * kBK * 2;
    const uint32_t stage_bytes = NPL * (kABytes + b_bytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)g.stages * stage_bytes);
    uint64_t* empty = full + g.stages;
    uint64_t* acc_full = empty + g.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x, n0 = blockIdx.y * g.bn;

    // tile -> image / origin (conv)
    int bidx = 0, y0 = 0, x0 = 0;
    if (MODE == MODE_CONV) {
        const int per_img = g.tiles_x * g.tiles_y;
        bidx = tile_m / per_img;
        const int t = tile_m - bidx * per_img;
        y0 = (t / g.tiles_x) * g.TH;
        x0 = (t % g.tiles_x) * g.TW;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
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

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < g.num_kb; ++kb) {
                const int s = kb % g.stages;
                const uint32_t par = ((kb / g.stages) & 1) ^ 1;
                mbar_wait(empty + s, par);
                uint8_t* sa = smem + (size_t)s * stage_bytes;
                uint8_t* sb = sa + NPL * kABytes;
                mbar_expect_tx(full + s, stage_bytes);
                if (MODE == MODE_PW) {
#pragma unroll
                    for (int p = 0; p < NPL; ++p) tma_load_3d(&tmA0, full + s, sa + p * kABytes, kb * kBK, tile_m * kBM, p);
                } else {
                    const int tap = kb / g.kb_per_tap;
                    const int r = kb - tap * g.kb_per_tap;
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    const bool src0 = r < g.kb_src0;
                    const CUtensorMap* tm = src0 ? &tmA0 : &tmA1;
                    const int cch = (src0 ? r : r - g.kb_src0) * kBK;
                    const int img = src0 ? bidx * g.a0_mul + g.a0_off : bidx * g.a1_mul + g.a1_off;
#pragma unroll
                    for (int p = 0; p < NPL; ++p) tma_load_5d(tm, full + s, sa + p * kABytes, cch, x0 + dx, y0 + dy, img, p);
                }
                int bk = kb * kBK;
                if (MODE == MODE_CONV) {
                    const int tap = kb / g.kb_per_tap;
                    bk = tap * g.bk_tap_stride + g.bk_off + (kb - tap * g.kb_per_tap) * kBK;
                }
#pragma unroll
                for (int p = 0; p < NPL; ++p) tma_load_3d(&tmB, full + s, sb + p * b_bytes, bk, n0, p);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc(g.bn);
        for (int kb = 0; kb < g.num_kb; ++kb) {
            const int s = kb % g.stages;
            const uint32_t par = (kb / g.stages) & 1;
            mbar_wait(full + s, par);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t b_hi = a_hi + NPL * kABytes;
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    const uint64_t dah = umma_desc(a_hi + k * 32);
                    const uint64_t dbh = umma_desc(b_hi + k * 32);
                    umma_bf16(tmem_base, dah, dbh, idesc, (kb | k) ? 1u : 0u);
                    if (TERMS == 3) {
                        const uint64_t dal = umma_desc(a_hi + kABytes + k * 32);
                        const uint64_t dbl = umma_desc(b_hi + b_bytes + k * 32);
                        umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                        umma_bf16(tmem_base, dal, dbh, idesc, 1u);
                    }
                }
                umma_commit(empty + s);                       // frees the smem stage when these MMAs retire
                if (kb == g.num_kb - 1) umma_commit(acc_full); // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;                               // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;                          // row of the tile
        int64_t orow, hrow = 0, crow = 0;
        bool rvalid;
        if (MODE == MODE_PW) {
            orow = (int64_t)tile_m * kBM + r;
            rvalid = orow < g.M;
        } else {
            const int y = y0 + r / g.TW, x = x0 + r % g.TW;
            rvalid = y < g.H && x < g.W;
            const int64_t pix = (int64_t)y * g.W + x;
            const int64_t hw = (int64_t)g.H * g.W;
            orow = (int64_t)(bidx * g.out_mul + g.out_off) * hw + pix;
            hrow = (int64_t)(bidx * g.a1_mul + g.a1_off) * hw + pix;
            crow = (int64_t)bidx * hw + pix;
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < g.bn; c0 += 16) {
            uint32_t raw[16];
            __syncwarp();
            tmem_ld16(trow + c0, raw);                         // warp-collective: executed by all lanes
            const int n = n0 + c0;
            if (!rvalid || n >= g.N) continue;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
            const bool second = n + 8 < g.N;                   // N is a multiple of 8
            if (g.bias) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    if (j4 >= 2 && !second) break;
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n) + j4);
                    v[j4 * 4 + 0] += b4.x; v[j4 * 4 + 1] += b4.y; v[j4 * 4 + 2] += b4.z; v[j4 * 4 + 3] += b4.w;
                }
            }
            if (EPI == EPI_STD) {
                if (g.flags & (UAVSAL_F_RELU6 | UAVSAL_F_RELU)) {
                    const float cap = (g.flags & UAVSAL_F_RELU6) ? 6.f : 3.0e38f;      // ReLU6 (model.py:71) | ReLU (ResNet / VGG backbones)
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fminf(fmaxf(v[j], 0.f), cap);
                }
                if (g.flags & UAVSAL_F_RESIDUAL) {
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
                store8(g.out.p + orow * g.out.ld + n, g.out.plane, v);
                if (second) store8(g.out.p + orow * g.out.ld + n + 8, g.out.plane, v + 8);
            } else if (EPI == EPI_TWA) {
                // h = i*x_t + (1-i)*h_{t-1}, i = sigmoid(conv)      (model_convlstm.py:283,290)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half == 1 && !second) break;
                    float xv[8], hv[8];
                    load8(g.x.p + orow * g.x.ld + n + half * 8, g.x.plane, xv);
                    load8(g.hprev.p + hrow * g.hprev.ld + n + half * 8, g.hprev.plane, hv);
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float gi = sigmoid_acc(v[half * 8 + j]);
                        o[j] = gi * xv[j] + (1.f - gi) * hv[j];
                    }
                    store8(g.out.p + orow * g.out.ld + n + half * 8, g.out.plane, o);
                }
            } else {
                // interleaved gates (i,f,o,g) x 4 channels     (model_convlstm.py:117-124)
                const int nch = g.N >> 2, ch = n >> 2;
                float* cs = g.c_state + crow * nch + ch;
                const int cnt = second ? 4 : 2;
                float hout[4];
                for (int j = 0; j < cnt; ++j) {
                    const float gi = sigmoid_acc(v[4 * j + 0]), gf = sigmoid_acc(v[4 * j + 1]);
                    const float go = sigmoid_acc(v[4 * j + 2]), gg = tanhf(v[4 * j + 3]);
                    const float cn = gf * cs[j] + gi * gg;
                    cs[j] = cn;
                    hout[j] = go * tanhf(cn);
                }
                if (second) store4(g.out.p + orow * g.out.ld + ch, g.out.plane, hout);
                else { store1(g.out.p + orow * g.out.ld + ch, g.out.plane, hout[0]); store1(g.out.p + orow * g.out.ld + ch + 1, g.out.plane, hout[1]); }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int tc_encode(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, const char* what, int swizzle128) {
    EncodeTiledFn fn = get_encode();
    UAVSAL_REQUIRE(fn != nullptr, UAVSAL_EDRIVER, "cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    // swizzle128: bit 0 = 128-byte swizzle, bit 1 = fp32 elements (else bf16)
    CUresult r = fn(tm, (swizzle128 & 2) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                    reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                    reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    (swizzle128 & 1) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAVSAL_REQUIRE(r == CUDA_SUCCESS, UAVSAL_EINVAL, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return 0;
}
static int encode(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const char* what) {
    return tc_encode(tm, base, rank, dims, strides_bytes, box, what, 1);
}

// 2-D activation matrix [rows][k] with hi/lo planes -> (k, rows, plane)
static int map_pw(CUtensorMap* tm, Act a, int64_t rows, int k) {
    const uint64_t dims[3] = {(uint64_t)k, (uint64_t)rows, a.plane ? 2u : 1u};
    const uint64_t str[2] = {(uint64_t)a.ld * 2, a.plane ? (uint64_t)a.plane * 2 : (uint64_t)a.ld * 2 * (uint64_t)rows};
    const uint32_t box[3] = {kBK, kBM, 1};
    return encode(tm, a.p, 3, dims, str, box, "A(pw)");
}
// weights [n][kpad] with hi/lo planes -> (kpad, n, 2)
static int map_w(CUtensorMap* tm, const uint16_t* w, int n, int kpad, int bn) {
    const uint64_t dims[3] = {(uint64_t)kpad, (uint64_t)n, 2};
    const uint64_t str[2] = {(uint64_t)kpad * 2, (uint64_t)kpad * 2 * (uint64_t)n};
    const uint32_t box[3] = {kBK, (uint32_t)bn, 1};
    return encode(tm, w, 3, dims, str, box, "B(weights)");
}
// NHWC images with hi/lo planes -> (c, w, h, nimg, plane)
static int map_img(CUtensorMap* tm, Act a, int nimg, int h, int w, int c, int tw, int th) {
    const uint64_t dims[5] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)nimg, a.plane ? 2u : 1u};
    const uint64_t row = (uint64_t)a.ld * 2;
    const uint64_t str[4] = {row, row * w, row * w * h, a.plane ? (uint64_t)a.plane * 2 : row * w * h * (uint64_t)nimg};
    const uint32_t box[5] = {kBK, (uint32_t)tw, (uint32_t)th, 1, 1};
    return encode(tm, a.p, 5, dims, str, box, "A(conv)");
}

static int pick_bn(int n) {
    // N tile: whole N when it fits one UMMA (<= 256), else 128/192/256 tiles chosen to minimise padding
    const int n16 = (n + 15) / 16 * 16;
    if (n16 <= 256) return n16;
    int best = 256, waste = 1 << 30;
    for (int bn = 256; bn >= 128; bn -= 64) {
        const int w = (n + bn - 1) / bn * bn - n;
        if (w < waste) { waste = w; best = bn; }
    }
    return best;
}

static int tmem_cols_for(int bn) {
    int c = 32;
    while (c < bn) c <<= 1;
    return c;
}

template <int MODE, int EPI>
static int launch_tc(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, TcArgs& g, int terms, int tiles_m,
                     cudaStream_t s, const char* what) {
    const int npl = terms == 3 ? 2 : 1;
    const uint32_t stage_bytes = npl * (kABytes + (uint32_t)g.bn * kBK * 2);
    int stages = (int)((200u * 1024u) / stage_bytes);
    if (stages > 6) stages = 6;
    if (stages > g.num_kb) stages = g.num_kb;
    if (stages < 1) stages = 1;
    g.stages = stages;
    g.tmem_cols = tmem_cols_for(g.bn);
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    dim3 grid(tiles_m, div_up(g.N, g.bn));
    cudaError_t e;
    if (terms == 3) {
        static bool attr3 = false;
        if (!attr3) {
            e = cudaFuncSetAttribute(gemm_tc_kernel<MODE, EPI, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return (int)e; }
            attr3 = true;
        }
        gemm_tc_kernel<MODE, EPI, 3><<<grid, kThreads, smem, s>>>(a0, a1, b, g);
    } else {
        static bool attr1 = false;
        if (!attr1) {
            e = cudaFuncSetAttribute(gemm_tc_kernel<MODE, EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return (int)e; }
            attr1 = true;
        }
        gemm_tc_kernel<MODE, EPI, 1><<<grid, kThreads, smem, s>>>(a0, a1, b, g);
    }
    return check_launch(what);
}

// ---- version 2: persistent kernel (gemm_tc2.cuh) ------------------------------------------------------
static int g_tc_version = 2;
int g_tc_debug = 0;     // DBG_* ablation bits ORed into the kernel flags

static int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static int g_tc_max_stages = 8;   // uavsal_set_option key 5 (dev): cap on the smem pipeline depth
static int g_tc_cluster = 2;   // uavsal_set_option key 4: CTAs per cluster of the persistent GEMM (1 = no multicast)

template <int MODE, int EPI, int TERMS, int CL>
static int launch_tc2_inst(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o, const TcArgs& g,
                           int grid, size_t smem, cudaStream_t s, const char* what) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<MODE, EPI, TERMS, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    int grid_x = grid;
    if (CL > 1) {
        // the persistent tile walk assumes every cluster is resident at once: clamp the grid to what the GPCs can co-schedule
        static int max_clusters = 0;
        if (!max_clusters) {
            cudaLaunchConfig_t q{};
            q.gridDim = dim3((num_sms() / CL) * CL); q.blockDim = dim3(kThreads2); q.dynamicSmemBytes = 227 * 1024 - 1024; q.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            q.attrs = at; q.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, gemm_tc2_kernel<MODE, EPI, TERMS, CL>, &q) != cudaSuccess || n <= 0) { n = num_sms() / CL - 2; cudaGetLastError(); }
            max_clusters = n;
        }
        if (grid_x > max_clusters * CL) grid_x = max_clusters * CL;
    }
    cudaError_t e = launch_k(gemm_tc2_kernel<MODE, EPI, TERMS, CL>, dim3(grid_x), dim3(kThreads2), smem, s, CL, a0, a1, b, o, g);
    if (e != cudaSuccess) { set_error("%s: cudaLaunchKernelEx: %s", what, cudaGetErrorString(e)); return (int)e; }
    return check_launch(what);
}

// the B tensor map of a cluster launch has a box of bn / CL rows (each CTA loads and multicasts its share)
static bool want_cluster(const TcArgs& g, int tiles_m) {
    // pair mode pays off when the tensor pipe / shared memory is the limiter (wide N tile, several k-blocks); the tiny-K
    // high-resolution layers are HBM-bound and lose to the extra cross-CTA handshakes
    return g_tc_cluster == 2 && tiles_m >= 2 && (int64_t)tiles_m * div_up(g.N, g.bn) >= num_sms() && g.bn % 32 == 0 && g.bn >= 128 &&
           g.num_kb >= 3;
}

template <int MODE, int EPI>
static int launch_tc2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o, TcArgs& g,
                      int terms, int tiles_m, bool cluster, cudaStream_t s, const char* what) {
    const int npl = terms == 3 ? 2 : 1;
    const uint32_t stage_bytes = npl * (kABytes + (uint32_t)(cluster ? g.bn / 2 : g.bn) * kBK * 2);   // a pair CTA stages half of B
    const uint32_t budget = 227u * 1024u - 1024u - kOutStageBytes - 512u;
    int stages = (int)(budget / stage_bytes);
    if (stages > g_tc_max_stages) stages = g_tc_max_stages;
    UAVSAL_REQUIRE(stages >= 1, UAVSAL_ENOTSUP, "%s: tile does not fit shared memory", what);
    g.stages = stages;
    g.tiles_m = tiles_m;
    g.tiles_n = div_up(g.N, g.bn);
    g.num_tiles = tiles_m * g.tiles_n;
    g.flags |= g_tc_debug;
    int cols = 32;
    while (cols < 2 * g.bn) cols <<= 1;
    g.tmem_cols = cols;
    const size_t smem = (size_t)stages * stage_bytes + kOutStageBytes + 1024 + 512;
    if (cluster) {
        const int work = div_up(tiles_m, 2) * g.tiles_n;
        int grid = 2 * work < num_sms() ? 2 * work : (num_sms() / 2) * 2;
        if (terms == 3) return launch_tc2_inst<MODE, EPI, 3, 2>(a0, a1, b, o, g, grid, smem, s, what);
        return launch_tc2_inst<MODE, EPI, 1, 2>(a0, a1, b, o, g, grid, smem, s, what);
    }
    const int grid = g.num_tiles < num_sms() ? g.num_tiles : num_sms();
    if (terms == 3) return launch_tc2_inst<MODE, EPI, 3, 1>(a0, a1, b, o, g, grid, smem, s, what);
    return launch_tc2_inst<MODE, EPI, 1, 1>(a0, a1, b, o, g, grid, smem, s, what);
}

// output maps: same geometry as the loads, 64-column boxes, extent = VALID channels so stores clip at N and never touch
// the neighbouring slot of a concat buffer
static int map_out_pw(CUtensorMap* tm, ActW o, int64_t rows, int n) {
    const uint64_t dims[3] = {(uint64_t)n, (uint64_t)rows, o.plane ? 2u : 1u};
    const uint64_t str[2] = {(uint64_t)o.ld * 2, o.plane ? (uint64_t)o.plane * 2 : (uint64_t)o.ld * 2 * (uint64_t)rows};
    const uint32_t box[3] = {64, kBM, 1};
    return encode(tm, o.p, 3, dims, str, box, "O(pw)");
}
static int map_out_img(CUtensorMap* tm, ActW o, int nimg, int h, int w, int c, int tw, int th) {
    return map_img(tm, Act{o.p, o.plane, o.ld}, nimg, h, w, c, tw, th);
}

// conv tiling: 128 pixels per tile as TW x TH with TW*TH = 128; pick the shape wasting the fewest pixels
static void pick_tile(int H, int W, int& tw, int& th) {
    int best = 1 << 30;
    tw = 16; th = 8;
    for (int cand = 8; cand <= 128; cand <<= 1) {
        const int t_h = 128 / cand;
        const int cover = div_up(W, cand) * cand * div_up(H, t_h) * t_h;
        if (cover < best) { best = cover; tw = cand; th = t_h; }
    }
}

int conv_tc(Act a0, int n0img, int a0_mul, int a0_off, int c0, Act a1, int n1img, int a1_mul, int a1_off, int c1,
            int batch, int H, int W, const uint16_t* wgt, int cout, const float* bias, int flags, int terms, int epi,
            Act x, Act hprev, float* c_state, ActW out, int out_nimg, int out_mul, int out_off, cudaStream_t s, const char* what,
            int wk_total = 0, int wk_off = 0, const float* gx = nullptr, float* raw_out = nullptr) {
    const bool v2 = !(terms & UAVSAL_TERMS_GEN1) && g_tc_version == 2;      // per-call kernel generation (cross-check engine "tc1")
    terms &= 0xFF;
    UAVSAL_REQUIRE(c0 % kBK == 0 && c1 % kBK == 0 && c0 > 0, UAVSAL_ENOTSUP, "%s: channels must be multiples of 64", what);
    UAVSAL_REQUIRE(cout % 8 == 0 && (terms == 1 || terms == 3), UAVSAL_EINVAL, "%s: bad cout/terms", what);
    TcArgs g{};
    int tw, th;
    pick_tile(H, W, tw, th);
    g.H = H; g.W = W; g.TW = tw; g.TH = th;
    g.tiles_x = div_up(W, tw); g.tiles_y = div_up(H, th);
    g.kb_src0 = c0 / kBK; g.kb_per_tap = (c0 + c1) / kBK; g.num_kb = 9 * g.kb_per_tap;
    g.a0_mul = a0_mul; g.a0_off = a0_off; g.a1_mul = a1_mul; g.a1_off = a1_off; g.out_mul = out_mul; g.out_off = out_off;
    g.M = batch * H * W; g.N = cout; g.bn = pick_bn(cout);
    const int tiles_m_ = batch * g.tiles_x * g.tiles_y;
    UAVSAL_REQUIRE(v2 || epi != EPI_RAW, UAVSAL_ENOTSUP, "%s: raw output needs the persistent kernel", what);
    if (v2) {   // small problems (one image per step in the recurrence): narrower N tiles so that more SMs get a tile
        while (g.bn > 64 && g.bn % 128 == 0 && tiles_m_ * div_up(cout, g.bn) < 100) g.bn /= 2;
    }
    g.bias = bias; g.flags = flags; g.out = out; g.x = x; g.hprev = hprev; g.c_state = c_state;
    const int kpad = wk_total ? 9 * wk_total : 9 * (c0 + c1);        // weight rows may hold more input channels than this launch reads
    g.bk_tap_stride = wk_total ? wk_total : c0 + c1;
    g.bk_off = wk_off;
    g.gx = gx; g.raw_out = raw_out;
    CUtensorMap tA0, tA1, tB;
    int rc = map_img(&tA0, a0, n0img, H, W, c0, tw, th);
    if (rc) return rc;
    if (c1) rc = map_img(&tA1, a1, n1img, H, W, c1, tw, th); else tA1 = tA0;
    if (rc) return rc;
    const int tiles_m = batch * g.tiles_x * g.tiles_y;
    const bool cl = v2 && want_cluster(g, tiles_m);
    rc = map_w(&tB, wgt, cout, kpad, cl ? g.bn / 2 : g.bn);
    if (rc) return rc;
    if (v2) {
        CUtensorMap tO;
        if (epi == EPI_RAW || epi == EPI_LSTM) tO = tA0;          // (the output map is only kept for TMA-store experiments)
        else rc = map_out_img(&tO, out, out_nimg, H, W, cout, tw, th);
        if (rc) return rc;
        if (epi == EPI_STD) return launch_tc2<MODE_CONV, EPI_STD>(tA0, tA1, tB, tO, g, terms, tiles_m, cl, s, what);
        if (epi == EPI_RAW) return launch_tc2<MODE_CONV, EPI_RAW>(tA0, tA1, tB, tO, g, terms, tiles_m, cl, s, what);
        if (epi == EPI_LSTM) return launch_tc2<MODE_CONV, EPI_LSTM>(tA0, tA1, tB, tO, g, terms, tiles_m, cl, s, what);
        return launch_tc2<MODE_CONV, EPI_TWA>(tA0, tA1, tB, tO, g, terms, tiles_m, cl, s, what);
    }
    if (epi == EPI_STD) return launch_tc<MODE_CONV, EPI_STD>(tA0, tA1, tB, g, terms, tiles_m, s, what);
    if (epi == EPI_TWA) return launch_tc<MODE_CONV, EPI_TWA>(tA0, tA1, tB, g, terms, tiles_m, s, what);
    return launch_tc<MODE_CONV, EPI_LSTM>(tA0, tA1, tB, g, terms, tiles_m, s, what);
}

// SIMT fall-backs for the recurrences live in gemm_simt.cu
int twa_sequence_simt(Act x, Act h0, int t_steps, int H, int W, int c, const float* w, ActW seq, cudaStream_t s);
int lstm_sequence_simt(Act x, Act h0, float* c_state, int b, int t_steps, int H, int W, int cin, int ch, const float* w,
                       const float* bias, ActW seq, cudaStream_t s);

}  // namespace uavsal

using namespace uavsal;

static inline bool act_ok16(const void* p, int64_t plane, int ld) {
    return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 8) == 0 && (plane % 8) == 0 && plane >= 0;
}

extern int g_dw_fast;
namespace uavsal {
extern int g_twa_resident, g_twa_bn, g_metrics_stream, g_metrics_stages;
size_t twa_sync_bytes(int batch, int H, int W);
int twa_sequence_persistent(Act x, Act h0, ActW seq, int batch, int t_steps, int H, int W, int c, const uint16_t* wgt, int wk_total, int wk_off,
                            const float* gx, int terms, int* ready, cudaStream_t s, int dbg);
int twa_step_resident(Act hsrc, int hsrc_nimg, int a_img, int a_stride, Act x, ActW seq, int out_img, int out_stride, int batch, int H, int W, int c,
                      const uint16_t* wgt, int wk_total, int wk_off, const float* gx, int terms, cudaStream_t s, int dbg);
}

extern "C" {

int uavsal_set_option(int key, int value) {
    if (key == 1 && (value == 1 || value == 2)) { g_tc_version = value; return 0; }
    if (key == 2 && value >= 0 && value <= 2) { g_dw_fast = value; return 0; }
    if (key == 3) { g_tc_debug = value & 0x5F0000; return 0; }
    if (key == 4 && (value == 1 || value == 2)) { g_tc_cluster = value; return 0; }
    if (key == 5 && value >= 1 && value <= 8) { g_tc_max_stages = value; return 0; }
    if (key == 6 && (value == 0 || value == 1)) { g_pdl = value; return 0; }
    if (key == 7 && value >= 0 && value <= 3) { g_twa_resident = value; return 0; }
    if (key == 8 && (value == 64 || value == 128)) { g_twa_bn = value; return 0; }
    if (key == 9 && value >= 0 && value <= 3) { g_metrics_stream = value; return 0; }
    if (key == 10 && (value == 3 || value == 4 || value == 5 || value == 7)) { g_metrics_stages = value; return 0; }
    set_error("set_option: unknown key %d / value %d", key, value);
    return UAVSAL_EINVAL;
}

int uavsal_pw_gemm(const uint16_t* a, int64_t a_plane, int a_ld, int m, int k, const uint16_t* wgt, int kpad, int n,
                   const float* bias, int flags, int terms, const uint16_t* res, int64_t res_plane, int res_ld,
                   uint16_t* out, int64_t out_plane, int out_ld, void* stream) {
    const bool f32out = flags & UAVSAL_F_OUT_F32;
    const bool q16out = flags & UAVSAL_F_OUT_Q16;
    const bool v2 = !(terms & UAVSAL_TERMS_GEN1) && g_tc_version == 2;
    terms &= 0xFF;
    UAVSAL_REQUIRE(!f32out || (v2 && !(flags & UAVSAL_F_SIGMOID)), UAVSAL_ENOTSUP, "pw_gemm: fp32 output needs the persistent kernel");
    UAVSAL_REQUIRE(!q16out || (v2 && flags == (UAVSAL_F_OUT_Q16 | UAVSAL_F_RELU6)), UAVSAL_ENOTSUP,
                   "pw_gemm: q16 output is the ReLU6 output of the persistent kernel (flags = RELU6 | OUT_Q16 only)");
    UAVSAL_REQUIRE(act_ok16(a, a_plane, a_ld) && (f32out ? (out && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && out_ld % 4 == 0) :
                                                  q16out ? act_ok16(out, 0, out_ld) : act_ok16(out, out_plane, out_ld)) && wgt &&
                       (reinterpret_cast<uintptr_t>(wgt) & 15) == 0 && m > 0 && k > 0 && n > 0 && k % 8 == 0 &&
                       kpad % 8 == 0 && kpad >= k && n % 8 == 0 && a_ld >= k && out_ld >= n,
                   UAVSAL_EINVAL, "pw_gemm: bad arguments (m=%d k=%d kpad=%d n=%d)", m, k, kpad, n);
    UAVSAL_REQUIRE(terms == 1 || (terms == 3 && a_plane != 0), UAVSAL_EINVAL, "pw_gemm: terms must be 1, or 3 with a lo plane");
    UAVSAL_REQUIRE(!(flags & UAVSAL_F_RESIDUAL) || act_ok16(res, res_plane, res_ld), UAVSAL_EINVAL,
                   "pw_gemm: residual requested without a residual tensor");
    UAVSAL_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, UAVSAL_EINVAL, "pw_gemm: bias must be 16-byte aligned");
    TcArgs g{};
    g.M = m; g.N = n; g.bn = pick_bn(n); g.num_kb = div_up(k, kBK);
    g.bias = bias; g.flags = flags;
    g.res = Act{res, res_plane, res_ld};
    g.out = ActW{out, out_plane, out_ld};
    CUtensorMap tA, tB;
    int rc = map_pw(&tA, Act{a, a_plane, a_ld}, m, k);
    if (rc) return rc;
    const bool cl = v2 && want_cluster(g, div_up(m, kBM));
    rc = map_w(&tB, wgt, n, kpad, cl ? g.bn / 2 : g.bn);
    if (rc) return rc;
    if (v2) {
        CUtensorMap tO = tA;                                   // (the copy-out uses the LSU path; the map is kept for TMA-store experiments)
        if (!f32out && !q16out) rc = map_out_pw(&tO, g.out, m, n);
        if (rc) return rc;
        if (q16out) return launch_tc2<MODE_PW, EPI_Q16>(tA, tA, tB, tO, g, terms, div_up(m, kBM), cl, (cudaStream_t)stream, "pw_gemm");
        // residual blocks with wide outputs: the residual is added in the coalesced copy-out phase (measured 318 -> 272 us for
        // 432000 x 32 -> 256); narrow outputs (N < 64) keep the row-strided loads, which are as fast there
        const bool res_coal = (flags & UAVSAL_F_RESIDUAL) && !(flags & UAVSAL_F_SIGMOID) && !f32out && n >= 64 && !(g_tc_debug & DBG_ROW_RES);
        if (res_coal) return launch_tc2<MODE_PW, EPI_RES>(tA, tA, tB, tO, g, terms, div_up(m, kBM), cl, (cudaStream_t)stream, "pw_gemm");
        return launch_tc2<MODE_PW, EPI_STD>(tA, tA, tB, tO, g, terms, div_up(m, kBM), cl, (cudaStream_t)stream, "pw_gemm");
    }
    return launch_tc<MODE_PW, EPI_STD>(tA, tA, tB, g, terms, div_up(m, kBM), (cudaStream_t)stream, "pw_gemm");
}

int uavsal_conv3x3(const uint16_t* in, int64_t in_plane, int in_ld, int n, int h, int w, int c, const uint16_t* wgt,
                   int cout, const float* bias, int flags, int terms, uint16_t* out, int64_t out_plane, int out_ld,
                   void* stream) {
    UAVSAL_REQUIRE(act_ok16(in, in_plane, in_ld) && act_ok16(out, out_plane, out_ld) && wgt && n > 0 && h > 0 && w > 0 &&
                       in_ld >= c && out_ld >= cout,
                   UAVSAL_EINVAL, "conv3x3: bad arguments");
    UAVSAL_REQUIRE((terms & 0xFF) == 1 || ((terms & 0xFF) == 3 && in_plane != 0), UAVSAL_EINVAL, "conv3x3: terms must be 1, or 3 with a lo plane");
    Act a{in, in_plane, in_ld};
    return conv_tc(a, n, 1, 0, c, a, n, 1, 0, 0, n, h, w, wgt, cout, bias, flags, terms, EPI_STD, Act{}, Act{}, nullptr,
                   ActW{out, out_plane, out_ld}, n, 1, 0, (cudaStream_t)stream, "conv3x3");
}

// one sequence (batch element): x (t_steps images), h0 (1 image), seq (t_steps images)
static int twa_sequence_one(Act X, Act H0, ActW S, int t_steps, int h, int w, int c, const uint16_t* wgt, int terms, float* gx_workspace,
                            cudaStream_t s) {
    Act SA{S.p, S.plane, S.ld};
    int rc;
    if (gx_workspace && !(terms & UAVSAL_TERMS_GEN1) && g_tc_version == 2) {
        // hoisted: G_x = W_x * x_t for ALL steps in one batched implicit GEMM (fp32 pre-activations), then per step only the
        // recurrent half W_h * h_{t-1} (K = 9c instead of 18c) with G_x[t] added in the epilogue before the gate
        rc = conv_tc(X, t_steps, 1, 0, c, X, t_steps, 1, 0, 0, t_steps, h, w, wgt, c, nullptr, 0, terms, EPI_RAW, Act{}, Act{},
                     nullptr, ActW{}, t_steps, 1, 0, s, "twa_sequence(x half)", 2 * c, 0, nullptr, gx_workspace);
        if (rc) return rc;
        for (int t = 0; t < t_steps; ++t) {
            if (t == 0)
                rc = conv_tc(H0, 1, 0, 0, c, H0, 1, 0, 0, 0, 1, h, w, wgt, c, nullptr, 0, terms, EPI_TWA, X, H0, nullptr, S, t_steps, 0, 0,
                             s, "twa_sequence(h half)", 2 * c, c, gx_workspace, nullptr);
            else
                rc = conv_tc(SA, t_steps, 0, t - 1, c, SA, t_steps, 0, t - 1, 0, 1, h, w, wgt, c, nullptr, 0, terms, EPI_TWA, X, SA, nullptr,
                             S, t_steps, 0, t, s, "twa_sequence(h half)", 2 * c, c, gx_workspace, nullptr);
            if (rc) return rc;
        }
        return 0;
    }
    for (int t = 0; t < t_steps; ++t) {
        // step t: A = [x_t, h_{t-1}], out = seq[t]
        if (t == 0)
            rc = conv_tc(X, t_steps, 0, 0, c, H0, 1, 0, 0, c, 1, h, w, wgt, c, nullptr, 0, terms, EPI_TWA, X, H0, nullptr, S,
                         t_steps, 0, 0, s, "twa_sequence");
        else
            rc = conv_tc(X, t_steps, 0, t, c, SA, t_steps, 0, t - 1, c, 1, h, w, wgt, c, nullptr, 0, terms, EPI_TWA, X, SA,
                         nullptr, S, t_steps, 0, t, s, "twa_sequence");
        if (rc) return rc;
    }
    return 0;
}

int uavsal_twa_sequence(const uint16_t* x, int64_t x_plane, int x_ld, const uint16_t* h0, int64_t h0_plane, int h0_ld,
                        int t_steps, int h, int w, int c, const uint16_t* wgt, const float* wgt_f32, int terms,
                        float* gx_workspace, uint16_t* seq_out, int64_t seq_plane, int seq_ld, int batch, void* sync_workspace,
                        void* stream) {
    UAVSAL_REQUIRE(act_ok16(x, x_plane, x_ld) && act_ok16(h0, h0_plane, h0_ld) && act_ok16(seq_out, seq_plane, seq_ld) &&
                       t_steps > 0 && batch > 0 && c % 8 == 0 && x_ld >= c && h0_ld >= c && seq_ld >= c,
                   UAVSAL_EINVAL, "twa_sequence: bad arguments");
    UAVSAL_REQUIRE((wgt != nullptr) != (wgt_f32 != nullptr), UAVSAL_EINVAL,
                   "twa_sequence: pass exactly one of wgt (tcgen05) / wgt_f32 (SIMT)");
    UAVSAL_REQUIRE(wgt_f32 || (terms & 0xFF) == 1 || ((terms & 0xFF) == 3 && x_plane && h0_plane && seq_plane), UAVSAL_EINVAL,
                   "twa_sequence: terms=3 needs lo planes");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t hw = (int64_t)h * w;
    if (!wgt_f32 && gx_workspace && !(terms & UAVSAL_TERMS_GEN1) && g_tc_version == 2 && g_twa_resident && c % 64 == 0 && c <= 512) {
        // all sequences of the batch advance together: W_x * x hoisted over batch*t_steps frames, then per step one launch of the
        // resident-A kernel (twa_step.cu) with the batch in blockIdx.z
        Act X{x, x_plane, x_ld}, H0{h0, h0_plane, h0_ld}, SA{seq_out, seq_plane, seq_ld};
        ActW S{seq_out, seq_plane, seq_ld};
        const int nimg = batch * t_steps;
        int rc = conv_tc(X, nimg, 1, 0, c, X, nimg, 1, 0, 0, nimg, h, w, wgt, c, nullptr, 0, terms, EPI_RAW, Act{}, Act{},
                         nullptr, ActW{}, nimg, 1, 0, s, "twa_sequence(x half)", 2 * c, 0, nullptr, gx_workspace);
        if (rc) return rc;
        if (sync_workspace && g_twa_resident == 3 && !(reinterpret_cast<uintptr_t>(sync_workspace) & 3)) {
            // the recurrence as ONE launch whenever the step grid fits the SMs (twa_seq_kernel); otherwise one launch per step
            rc = twa_sequence_persistent(X, H0, S, batch, t_steps, h, w, c, wgt, 2 * c, c, gx_workspace, terms & 0xFF,
                                         static_cast<int*>(sync_workspace), s, g_tc_debug);
            if (rc != UAVSAL_ENOTSUP) return rc;
        }
        for (int t = 0; t < t_steps; ++t) {
            rc = t == 0 ? twa_step_resident(H0, batch, 0, 1, X, S, 0, t_steps, batch, h, w, c, wgt, 2 * c, c, gx_workspace, terms, s, g_tc_debug)
                        : twa_step_resident(SA, nimg, t - 1, t_steps, X, S, t, t_steps, batch, h, w, c, wgt, 2 * c, c, gx_workspace, terms, s, g_tc_debug);
            if (rc) return rc;
        }
        return 0;
    }
    for (int b = 0; b < batch; ++b) {
        Act X{x + b * t_steps * hw * x_ld, x_plane, x_ld}, H0{h0 + b * hw * h0_ld, h0_plane, h0_ld};
        ActW S{seq_out + b * t_steps * hw * seq_ld, seq_plane, seq_ld};
        int rc = wgt_f32 ? twa_sequence_simt(X, H0, t_steps, h, w, c, wgt_f32, S, s)
                         : twa_sequence_one(X, H0, S, t_steps, h, w, c, wgt, terms, gx_workspace ? gx_workspace + b * t_steps * hw * c : nullptr, s);
        if (rc) return rc;
    }
    return 0;
}

size_t uavsal_twa_sync_bytes(int batch, int h, int w) {
    return batch > 0 && h > 0 && w > 0 ? twa_sync_bytes(batch, h, w) : 0;
}

int uavsal_convlstm_sequence(const uint16_t* x, int64_t x_plane, int x_ld, const uint16_t* h0, int64_t h0_plane,
                             int h0_ld, float* c_state, int b, int t_steps, int h, int w, int cin, int ch,
                             const uint16_t* wgt, const float* wgt_f32, const float* bias, int terms, uint16_t* seq_out,
                             int64_t seq_plane, int seq_ld, void* stream) {
    UAVSAL_REQUIRE(act_ok16(x, x_plane, x_ld) && act_ok16(h0, h0_plane, h0_ld) && act_ok16(seq_out, seq_plane, seq_ld) &&
                       c_state && b > 0 && t_steps > 0 && cin % 8 == 0 && ch % 8 == 0 && x_ld >= cin && h0_ld >= ch &&
                       seq_ld >= ch,
                   UAVSAL_EINVAL, "convlstm_sequence: bad arguments");
    UAVSAL_REQUIRE((wgt != nullptr) != (wgt_f32 != nullptr), UAVSAL_EINVAL,
                   "convlstm_sequence: pass exactly one of wgt (tcgen05) / wgt_f32 (SIMT)");
    Act X{x, x_plane, x_ld}, H0{h0, h0_plane, h0_ld};
    ActW S{seq_out, seq_plane, seq_ld};
    cudaStream_t s = (cudaStream_t)stream;
    if (wgt_f32) return lstm_sequence_simt(X, H0, c_state, b, t_steps, h, w, cin, ch, wgt_f32, bias, S, s);
    UAVSAL_REQUIRE((terms & 0xFF) == 1 || ((terms & 0xFF) == 3 && x_plane && h0_plane && seq_plane), UAVSAL_EINVAL,
                   "convlstm_sequence: terms=3 needs lo planes");
    Act SA{seq_out, seq_plane, seq_ld};
    for (int t = 0; t < t_steps; ++t) {
        // image index of batch element bi: x -> bi*T + t ; h_{t-1} -> h0[bi] or seq[bi*T + t-1] ; out -> seq[bi*T + t]
        int rc;
        if (t == 0)
            rc = conv_tc(X, b * t_steps, t_steps, 0, cin, H0, b, 1, 0, ch, b, h, w, wgt, 4 * ch, bias, 0, terms, EPI_LSTM,
                         Act{}, Act{}, c_state, S, b * t_steps, t_steps, 0, s, "convlstm_sequence");
        else
            rc = conv_tc(X, b * t_steps, t_steps, t, cin, SA, b * t_steps, t_steps, t - 1, ch, b, h, w, wgt, 4 * ch, bias, 0,
                         terms, EPI_LSTM, Act{}, Act{}, c_state, S, b * t_steps, t_steps, t, s, "convlstm_sequence");
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"

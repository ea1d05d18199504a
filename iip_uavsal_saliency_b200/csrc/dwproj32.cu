// Fused depthwise 3x3 (+BN+ReLU6) -> 1x1 project (+BN) for the one dwBlock whose project conv is too narrow for the tensor
// pipe: features.1 (32 channels -> 16, stride 1, 180x320; torchvision InvertedResidual with expand_ratio 1, used by
// model_feature.py:63).  As two kernels it cost a depthwise pass (885 MB in, 885 MB out) plus a 32 -> 16 GEMM whose
// 128x16 tiles leave 12 of the 16 epilogue warps idle and half of every TMA box zero-filled (400 us for 1.3 GB).
//
// Here the whole block is fp32 FFMA work on a TMA-staged tile (the effective project weights are the same hi + lo bf16
// pair the GEMM uses): a CTA double-buffers haloed 18x10-pixel x 32-channel fp32 boxes (128 bytes per pixel, 128-byte
// swizzle: 8 consecutive pixels put a given 16-byte channel chunk into 8 different bank groups).  Thread = (output
// pixel, 16-channel half): for each of its four 4-channel chunks it runs the 3x3 depthwise taps, ReLU6, and adds the
// chunk's contribution to all 16 project outputs (project weights and taps are read from shared memory as warp-wide
// broadcasts); the two halves of a pixel sit in adjacent lanes and swap 8 partial sums each, so every lane ends up
// with 8 finished output channels = one 16-byte store per bf16 plane.  The depthwise output never exists in memory.
#include <cstring>

#include "tc_common.cuh"

namespace uavsal {

constexpr int kD3TW = 16, kD3TH = 8, kD3IW = kD3TW + 2, kD3IH = kD3TH + 2;
constexpr int kD3C = 32, kD3N = 16;
constexpr uint32_t kD3TileBytes = ((kD3IW * kD3IH * 128 + 1023) / 1024) * 1024;       // 23 552: buffers stay 1024-byte aligned (swizzle atom)

struct DwProj32Args {
    int n, h, w;
    int tiles_x, tiles_y, num_tiles;
    const float* wd;        // [9][32] depthwise taps (BN folded)
    const float* bd;        // [32]
    const uint16_t* wp;     // [2 planes][16][kpad] project weights, bf16 hi | lo
    int kpad;
    const float* bias;      // [16]
    ActW out;
};

__global__ void __launch_bounds__(256, 3) dwproj32_kernel(const __grid_constant__ CUtensorMap tmIn, const DwProj32Args g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* wp_s = reinterpret_cast<float*>(smem + 2 * kD3TileBytes);          // [32 c][16 o]; channels 16..31 shifted by 16 bytes:
    float* wd_s = wp_s + kD3C * kD3N + 4;                                      //   the two halves of a warp read different banks; [9][32]
    float* bd_s = wd_s + 9 * kD3C;                                             // [32]
    float* bo_s = bd_s + kD3C;                                                 // [16]
    uint64_t* bar = reinterpret_cast<uint64_t*>(bo_s + kD3N);                  // [2]

    const int tid = threadIdx.x;
    pdl_trigger();
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
    // constants (weights): no dependency on the producer kernel
    for (int i = tid; i < kD3C * kD3N; i += 256) {
        const int c = i >> 4, o = i & 15;
        wp_s[i + (c >> 4) * 4] = bf16_bits_to_f32(__ldg(g.wp + o * g.kpad + c)) + bf16_bits_to_f32(__ldg(g.wp + (int64_t)kD3N * g.kpad + o * g.kpad + c));
    }
    for (int i = tid; i < 9 * kD3C; i += 256) wd_s[i] = __ldg(g.wd + i);
    if (tid < kD3C) bd_s[tid] = __ldg(g.bd + tid);
    if (tid < kD3N) bo_s[tid] = __ldg(g.bias + tid);
    __syncthreads();
    pdl_wait();                                                                // the input tensor is the previous kernel's output

    auto decode = [&](int t, int& x0, int& y0, int& img) {
        int r = t;
        x0 = (r % g.tiles_x) * kD3TW; r /= g.tiles_x;
        y0 = (r % g.tiles_y) * kD3TH;
        img = r / g.tiles_y;
    };
    auto issue = [&](int t, int b) {                                           // one thread
        int x0, y0, img;
        decode(t, x0, y0, img);
        fence_async_smem();                                                    // order the buffer's generic reads before the async overwrite
        mbar_expect_tx(bar + b, kD3IW * kD3IH * 128);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_u32(smem + b * kD3TileBytes)), "l"(&tmIn), "r"(smem_u32(bar + b)), "r"(0), "r"(x0 - 1), "r"(y0 - 1), "r"(img)
            : "memory");
    };

    const int half = tid & 1, col = (tid >> 1) & 15, row = tid >> 5;           // warp = one output row of the tile
    const uint32_t wp_a = smem_u32(wp_s), wd_a = smem_u32(wd_s), bd_a = smem_u32(bd_s);
    if (tid == 0 && (int)blockIdx.x < g.num_tiles) issue(blockIdx.x, 0);
    int it = 0;
    for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x, ++it) {
        const int b = it & 1;
        if (tid == 0 && t + (int)gridDim.x < g.num_tiles) issue(t + gridDim.x, b ^ 1);   // prefetch the next tile
        int x0, y0, img;
        decode(t, x0, y0, img);
        mbar_wait(bar + b, (it >> 1) & 1);
        const uint32_t tile = smem_u32(smem) + b * kD3TileBytes;

        float p[kD3N];
#pragma unroll
        for (int o = 0; o < kD3N; ++o) p[o] = 0.f;
#pragma unroll
        for (int chunk = 0; chunk < 4; ++chunk) {
            const int q = half * 4 + chunk;                                    // 4-channel chunk of the 32 channels
            float acc[4];
            {
                const uint4 b4 = lds128(bd_a + q * 16);
                acc[0] = __uint_as_float(b4.x); acc[1] = __uint_as_float(b4.y); acc[2] = __uint_as_float(b4.z); acc[3] = __uint_as_float(b4.w);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int pp = (row + ky) * kD3IW + col + kx;              // pixel inside the haloed box = 128-byte row
                    const uint4 v = lds128(tile + pp * 128 + ((q ^ (pp & 7)) << 4));
                    const uint4 w4 = lds128(wd_a + ((ky * 3 + kx) * kD3C + q * 4) * 4);
                    acc[0] = fmaf(__uint_as_float(v.x), __uint_as_float(w4.x), acc[0]);
                    acc[1] = fmaf(__uint_as_float(v.y), __uint_as_float(w4.y), acc[1]);
                    acc[2] = fmaf(__uint_as_float(v.z), __uint_as_float(w4.z), acc[2]);
                    acc[3] = fmaf(__uint_as_float(v.w), __uint_as_float(w4.w), acc[3]);
                }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = relu6f(acc[j]);
#pragma unroll
                for (int o4 = 0; o4 < 4; ++o4) {
                    const uint4 w4 = lds128(wp_a + ((q * 4 + j) * kD3N + o4 * 4 + half * 4) * 4);
                    p[o4 * 4 + 0] = fmaf(a, __uint_as_float(w4.x), p[o4 * 4 + 0]);
                    p[o4 * 4 + 1] = fmaf(a, __uint_as_float(w4.y), p[o4 * 4 + 1]);
                    p[o4 * 4 + 2] = fmaf(a, __uint_as_float(w4.z), p[o4 * 4 + 2]);
                    p[o4 * 4 + 3] = fmaf(a, __uint_as_float(w4.w), p[o4 * 4 + 3]);
                }
            }
        }
        // the two halves of a pixel (adjacent lanes) swap 8 partial sums: half h keeps outputs 8h..8h+7
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = half ? p[i] : p[8 + i];
            const float keep = half ? p[8 + i] : p[i];
            r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
        const int ox = x0 + col, oy = y0 + row;
        if (ox < g.w && oy < g.h) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float a = r[2 * i] + bo_s[half * 8 + 2 * i], c = r[2 * i + 1] + bo_s[half * 8 + 2 * i + 1];
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, c);
                const float2 hf = __bfloat1622float2(h2);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, c - hf.y);
                hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
                lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            uint16_t* dst = g.out.p + (((int64_t)img * g.h + oy) * g.w + ox) * g.out.ld + half * 8;
            *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (g.out.plane) *reinterpret_cast<uint4*>(dst + g.out.plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        __syncthreads();                                                       // tile consumed: its buffer may be refilled
    }
}

// ---------------------------------------------------------------------------------------------------
// Parameter-block variant (uavsal_dw_project32_hw, the engine's default): the block's 848 weights (taps, both biases, project
// matrix) are HOST arrays copied into the kernel parameters (constant bank), so the weights reach the FMAs as warp-uniform operands
// (pairs of them through uniform registers, feeding packed fma.rn.f32x2) and shared memory only carries the activations.  Thread = one output pixel with all 32 channels (no exchange
// between threads); tile 16 x 16 output pixels, 18 x 18 haloed box (1.27x re-read instead of 1.41x).
// ---------------------------------------------------------------------------------------------------
constexpr int kP3T = 16, kP3I = kP3T + 2;
constexpr uint32_t kP3TileBytes = ((kP3I * kP3I * 128 + 1023) / 1024) * 1024;          // 41 984

struct DwProj32W {
    float wd[9][kD3C];
    float bd[kD3C];
    float wp[kD3C][kD3N];      // [hidden channel][output channel]
    float bo[kD3N];
};

struct DwProj32PArgs {
    int n, h, w;
    int tiles_x, tiles_y, num_tiles;
    ActW out;
};

__global__ void __launch_bounds__(256, 2) dwproj32p_kernel(const __grid_constant__ CUtensorMap tmIn, const DwProj32PArgs g,
                                                           const __grid_constant__ DwProj32W W) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kP3TileBytes);      // [2]
    const int tid = threadIdx.x;
    pdl_trigger();
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
    __syncthreads();
    pdl_wait();

    auto decode = [&](int t, int& x0, int& y0, int& img) {
        int r = t;
        x0 = (r % g.tiles_x) * kP3T; r /= g.tiles_x;
        y0 = (r % g.tiles_y) * kP3T;
        img = r / g.tiles_y;
    };
    auto issue = [&](int t, int b) {                                           // one thread
        int x0, y0, img;
        decode(t, x0, y0, img);
        fence_async_smem();
        mbar_expect_tx(bar + b, kP3I * kP3I * 128);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_u32(smem + b * kP3TileBytes)), "l"(&tmIn), "r"(smem_u32(bar + b)), "r"(0), "r"(x0 - 1), "r"(y0 - 1), "r"(img)
            : "memory");
    };

    const int col = tid & 15, row = tid >> 4;
    if (tid == 0 && (int)blockIdx.x < g.num_tiles) issue(blockIdx.x, 0);
    int it = 0;
    for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x, ++it) {
        const int b = it & 1;
        if (tid == 0 && t + (int)gridDim.x < g.num_tiles) issue(t + gridDim.x, b ^ 1);
        int x0, y0, img;
        decode(t, x0, y0, img);
        mbar_wait(bar + b, (it >> 1) & 1);
        const uint32_t tile = smem_u32(smem) + b * kP3TileBytes;
        // channel / output pairs as packed fma.rn.f32x2 (the same IEEE fma per lane): the weight pairs come through uniform
        // registers (one 16-byte constant load per two packed FMAs) - 400 FFMA2 + 200 LDCU.128 per pixel instead of 800 FFMA
        float2 p[kD3N / 2];
#pragma unroll
        for (int o = 0; o < kD3N / 2; ++o) p[o] = make_float2(W.bo[2 * o], W.bo[2 * o + 1]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {                                          // 4-channel chunk
            float2 acc[2] = {make_float2(W.bd[q * 4 + 0], W.bd[q * 4 + 1]), make_float2(W.bd[q * 4 + 2], W.bd[q * 4 + 3])};
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int pp = (row + ky) * kP3I + col + kx;               // pixel inside the haloed box = 128-byte row
                    const uint4 v = lds128(tile + pp * 128 + ((q ^ (pp & 7)) << 4));
                    const float* wk = W.wd[ky * 3 + kx] + q * 4;
                    acc[0] = __ffma2_rn(make_float2(__uint_as_float(v.x), __uint_as_float(v.y)), make_float2(wk[0], wk[1]), acc[0]);
                    acc[1] = __ffma2_rn(make_float2(__uint_as_float(v.z), __uint_as_float(v.w)), make_float2(wk[2], wk[3]), acc[1]);
                }
            const float a4[4] = {relu6f(acc[0].x), relu6f(acc[0].y), relu6f(acc[1].x), relu6f(acc[1].y)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* wo = W.wp[q * 4 + j];
#pragma unroll
                for (int o = 0; o < kD3N / 2; ++o) p[o] = __ffma2_rn(make_float2(a4[j], a4[j]), make_float2(wo[2 * o], wo[2 * o + 1]), p[o]);
            }
        }
        const int ox = x0 + col, oy = y0 + row;
        if (ox < g.w && oy < g.h) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split2(p[i].x, p[i].y, hi[i], lo[i]);
            uint16_t* dst = g.out.p + (((int64_t)img * g.h + oy) * g.w + ox) * g.out.ld;
            reinterpret_cast<uint4*>(dst)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            reinterpret_cast<uint4*>(dst)[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            if (g.out.plane) {
                reinterpret_cast<uint4*>(dst + g.out.plane)[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                reinterpret_cast<uint4*>(dst + g.out.plane)[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
        }
        __syncthreads();                                                       // tile consumed: its buffer may be refilled
    }
}

int dw_project32_hw(const float* hid, int hid_ld, int n, int h, int w, const float* wd_host, const float* bd_host, const float* wp_host,
                    const float* bias_host, ActW out, cudaStream_t s) {
    DwProj32W W;
    memcpy(W.wd, wd_host, sizeof(W.wd));
    memcpy(W.bd, bd_host, sizeof(W.bd));
    for (int o = 0; o < kD3N; ++o)
        for (int c = 0; c < kD3C; ++c) W.wp[c][o] = wp_host[o * kD3C + c];     // (cout, hidden) row-major in
    memcpy(W.bo, bias_host, sizeof(W.bo));
    DwProj32PArgs g{};
    g.n = n; g.h = h; g.w = w;
    g.tiles_x = div_up(w, kP3T); g.tiles_y = div_up(h, kP3T);
    g.num_tiles = n * g.tiles_x * g.tiles_y;
    g.out = out;
    CUtensorMap tm;
    const uint64_t dims[4] = {(uint64_t)kD3C, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t rowb = (uint64_t)hid_ld * 4;
    const uint64_t str[3] = {rowb, rowb * w, rowb * w * h};
    const uint32_t box[4] = {kD3C, (uint32_t)kP3I, (uint32_t)kP3I, 1};
    int rc = tc_encode(&tm, hid, 4, dims, str, box, "dw_project32 input (f32)", 3);
    if (rc) return rc;
    const size_t smem = 2 * kP3TileBytes + 16 + 1024;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(dwproj32p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("dw_project32_hw: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    static int sms = 0, bps = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, dwproj32p_kernel, 256, smem) != cudaSuccess || bps <= 0) { bps = 2; cudaGetLastError(); }
    }
    const int grid = g.num_tiles < sms * bps ? g.num_tiles : sms * bps;
    cudaError_t e = launch_k(dwproj32p_kernel, dim3(grid), dim3(256), smem, s, 1, tm, g, W);
    if (e != cudaSuccess) { set_error("dw_project32_hw: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("dw_project32_hw");
}

// hidden == 32, cout == 16 variant of uavsal_dw_project (no residual: the block changes the channel count)
int dw_project32(const float* hid, int hid_ld, int n, int h, int w, const float* wd, const float* bd, const uint16_t* wgt, int kpad,
                 const float* bias, ActW out, cudaStream_t s) {
    DwProj32Args g{};
    g.n = n; g.h = h; g.w = w;
    g.tiles_x = div_up(w, kD3TW); g.tiles_y = div_up(h, kD3TH);
    g.num_tiles = n * g.tiles_x * g.tiles_y;
    g.wd = wd; g.bd = bd; g.wp = wgt; g.kpad = kpad; g.bias = bias; g.out = out;
    CUtensorMap tm;
    const uint64_t dims[4] = {(uint64_t)kD3C, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t rowb = (uint64_t)hid_ld * 4;
    const uint64_t str[3] = {rowb, rowb * w, rowb * w * h};
    const uint32_t box[4] = {kD3C, (uint32_t)kD3IW, (uint32_t)kD3IH, 1};
    int rc = tc_encode(&tm, hid, 4, dims, str, box, "dw_project32 input (f32)", 3);
    if (rc) return rc;
    const size_t smem = 2 * kD3TileBytes + (kD3C * kD3N + 4 + 9 * kD3C + kD3C + kD3N) * 4 + 16 + 1024;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(dwproj32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("dw_project32: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    static int sms = 0, bps = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, dwproj32_kernel, 256, smem) != cudaSuccess || bps <= 0) { bps = 2; cudaGetLastError(); }
    }
    const int grid = g.num_tiles < sms * bps ? g.num_tiles : sms * bps;
    cudaError_t e = launch_k(dwproj32_kernel, dim3(grid), dim3(256), smem, s, 1, tm, g);
    if (e != cudaSuccess) { set_error("dw_project32: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("dw_project32");
}

}  // namespace uavsal

extern "C" int uavsal_dw_project32_hw(const float* hid, int hid_ld, int n, int h, int w, const float* wd_host, const float* bd_host,
                                      const float* wp_host, const float* bias_host, uint16_t* out, int64_t out_plane, int out_ld, void* stream) {
    UAVSAL_REQUIRE(hid && (reinterpret_cast<uintptr_t>(hid) & 15) == 0 && wd_host && bd_host && wp_host && bias_host && out &&
                       (reinterpret_cast<uintptr_t>(out) & 15) == 0 && n > 0 && h > 0 && w > 0 && hid_ld % 4 == 0 && hid_ld >= 32 &&
                       out_ld % 8 == 0 && out_ld >= 16 && out_plane >= 0 && out_plane % 8 == 0,
                   UAVSAL_EINVAL, "dw_project32_hw: bad arguments");
    return uavsal::dw_project32_hw(hid, hid_ld, n, h, w, wd_host, bd_host, wp_host, bias_host, uavsal::ActW{out, out_plane, out_ld}, (cudaStream_t)stream);
}

// Kernels only the alternative backbones need (model_feature.ReResNet / ReVGG, model_feature.py:72-128; torchvision resnet.py /
// vgg.py): the first conv from the raw frame (ResNet conv1 7x7 stride 2, VGG features.0 3x3 stride 1; 3 input channels, with the
// uint8 normalisation of utils_data.normalize_data fused as in the MobileNetV2 stem), max pooling (k = 3 s2 p1 for ResNet, k = 2
// s2 for VGG; k = 1 s2 is the row subsampling in front of a stride-2 1x1 conv and behind a stride-1 evaluation of a stride-2 3x3
// conv), and the residual add + ReLU that closes a ResNet block.  All the other convs of those backbones run on the tcgen05 GEMM /
// implicit-GEMM kernels with the UAVSAL_F_RELU epilogue flag.  These are HBM-bound glue kernels, not on the UAVSal hot path.
#include "common.cuh"

namespace uavsal {

// out[n][oy][ox][co] = act(bias[co] + sum_{ky,kx,ci} w[(ky*K+kx)*3+ci][co] * x[n][oy*s-p+ky][ox*s-p+kx][ci]),  co < 64
// block = 64 output pixels x 4 channel quarters (thread = one pixel x 16 output channels); weights staged in shared memory
template <int KIND>
__global__ void __launch_bounds__(256) conv_first_kernel(const void* __restrict__ x, int n, int h, int w, int ho, int wo, int K, int stride, int pad,
                                                         const float* __restrict__ wgt, const float* __restrict__ bias, int relu, ActW out) {
    extern __shared__ float cf_smem[];
    float* sw = cf_smem;                       // [K*K*3][64]
    float* lut = sw + K * K * 3 * 64;          // [3][256]
    for (int i = threadIdx.x; i < K * K * 3 * 64; i += blockDim.x) sw[i] = wgt[i];
    if (KIND != 0) {
        for (int i = threadIdx.x; i < 768; i += blockDim.x) {
            const int ch = i >> 8, u = i & 255;
            const float mean = ch == 0 ? 0.485f : (ch == 1 ? 0.456f : 0.406f);
            const float sd = ch == 0 ? 0.229f : (ch == 1 ? 0.224f : 0.225f);
            lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), mean), sd);          // utils_data.py:56-60
        }
    }
    __syncthreads();
    const int64_t total = (int64_t)n * ho * wo;
    const int q = threadIdx.x >> 6;            // channel quarter
    for (int64_t p0 = (int64_t)blockIdx.x * 64; p0 < total; p0 += (int64_t)gridDim.x * 64) {
        const int64_t p = p0 + (threadIdx.x & 63);
        if (p >= total) continue;
        const int ox = (int)(p % wo), oy = (int)((p / wo) % ho), img = (int)(p / ((int64_t)wo * ho));
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = bias ? __ldg(bias + q * 16 + j) : 0.f;
        for (int ky = 0; ky < K; ++ky) {
            const int y = oy * stride - pad + ky;
            if (y < 0 || y >= h) continue;
            for (int kx = 0; kx < K; ++kx) {
                const int xx = ox * stride - pad + kx;
                if (xx < 0 || xx >= w) continue;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float v;
                    if (KIND == 0) v = __ldg(reinterpret_cast<const float*>(x) + (((int64_t)img * 3 + ch) * h + y) * w + xx);
                    else if (KIND == 1) v = lut[ch * 256 + __ldg(reinterpret_cast<const uint8_t*>(x) + (((int64_t)img * 3 + ch) * h + y) * w + xx)];
                    else v = lut[ch * 256 + __ldg(reinterpret_cast<const uint8_t*>(x) + (((int64_t)img * h + y) * w + xx) * 3 + ch)];
                    const float4* wr = reinterpret_cast<const float4*>(sw + ((ky * K + kx) * 3 + ch) * 64 + q * 16);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 ww = wr[j4];
                        acc[j4 * 4 + 0] = fmaf(v, ww.x, acc[j4 * 4 + 0]); acc[j4 * 4 + 1] = fmaf(v, ww.y, acc[j4 * 4 + 1]);
                        acc[j4 * 4 + 2] = fmaf(v, ww.z, acc[j4 * 4 + 2]); acc[j4 * 4 + 3] = fmaf(v, ww.w, acc[j4 * 4 + 3]);
                    }
                }
            }
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = fmaxf(acc[j], 0.f);
        }
        uint16_t* o = out.p + p * out.ld + q * 16;
        store8(o, out.plane, acc);
        store8(o + 8, out.plane, acc + 8);
    }
}

// max over the k x k window (windows are clipped at the border: the padding never wins, as -inf padding in torch)
__global__ void __launch_bounds__(256) pool_kernel(Act in, int n, int h, int w, int c, int k, int stride, int pad, int ho, int wo, ActW out) {
    const int c8 = c >> 3;
    const int64_t total = (int64_t)n * ho * wo * c8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cb = (int)(i % c8);
        const int64_t p = i / c8;
        const int ox = (int)(p % wo), oy = (int)((p / wo) % ho), img = (int)(p / ((int64_t)wo * ho));
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -3.4e38f;
        for (int ky = 0; ky < k; ++ky) {
            const int y = oy * stride - pad + ky;
            if (y < 0 || y >= h) continue;
            for (int kx = 0; kx < k; ++kx) {
                const int xx = ox * stride - pad + kx;
                if (xx < 0 || xx >= w) continue;
                float v[8];
                load8(in.p + (((int64_t)img * h + y) * w + xx) * in.ld + cb * 8, in.plane, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
            }
        }
        store8(out.p + p * out.ld + cb * 8, out.plane, m);
    }
}

__global__ void __launch_bounds__(256) add_act_kernel(Act a, Act b, int64_t rows, int c, int relu, ActW out) {
    const int c8 = c >> 3;
    const int64_t total = rows * c8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c8;
        const int cb = (int)(i % c8) * 8;
        float x[8], y[8];
        load8(a.p + r * a.ld + cb, a.plane, x);
        load8(b.p + r * b.ld + cb, b.plane, y);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x[j] += y[j];
            if (relu) x[j] = fmaxf(x[j], 0.f);
        }
        store8(out.p + r * out.ld + cb, out.plane, x);
    }
}

static inline bool bb_act_ok(const void* p, int64_t plane, int ld) {
    return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 8) == 0 && (plane % 8) == 0 && plane >= 0;
}

}  // namespace uavsal

using namespace uavsal;

extern "C" int uavsal_conv_first(const void* x, int x_kind, int n, int h, int w, int k, int stride, const float* wgt, const float* bias,
                                 int flags, uint16_t* out, int64_t out_plane, int out_ld, void* stream) {
    UAVSAL_REQUIRE(x && wgt && bb_act_ok(out, out_plane, out_ld) && out_ld >= 64 && n > 0 && h > 0 && w > 0 && x_kind >= 0 && x_kind <= 2,
                   UAVSAL_EINVAL, "conv_first: bad arguments");
    UAVSAL_REQUIRE((k == 3 || k == 7) && (stride == 1 || stride == 2), UAVSAL_ENOTSUP, "conv_first: k in {3,7}, stride in {1,2} (3 -> 64 channels)");
    UAVSAL_REQUIRE(!(flags & ~(UAVSAL_F_RELU)), UAVSAL_ENOTSUP, "conv_first: only the ReLU flag is supported");
    const int pad = k / 2;
    const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
    const int64_t total = (int64_t)n * ho * wo;
    const size_t smem = ((size_t)k * k * 3 * 64 + 768) * sizeof(float);
    const int grid = (int)((total + 63) / 64 < 148 * 8 ? (total + 63) / 64 : 148 * 8);
    ActW o{out, out_plane, out_ld};
    cudaStream_t s = (cudaStream_t)stream;
    const int relu = (flags & UAVSAL_F_RELU) ? 1 : 0;
    if (x_kind == 0) conv_first_kernel<0><<<grid, 256, smem, s>>>(x, n, h, w, ho, wo, k, stride, pad, wgt, bias, relu, o);
    else if (x_kind == 1) conv_first_kernel<1><<<grid, 256, smem, s>>>(x, n, h, w, ho, wo, k, stride, pad, wgt, bias, relu, o);
    else conv_first_kernel<2><<<grid, 256, smem, s>>>(x, n, h, w, ho, wo, k, stride, pad, wgt, bias, relu, o);
    return check_launch("conv_first");
}

extern "C" int uavsal_maxpool(const uint16_t* in, int64_t in_plane, int in_ld, int n, int h, int w, int c, int k, int stride, int pad,
                              uint16_t* out, int64_t out_plane, int out_ld, void* stream) {
    UAVSAL_REQUIRE(bb_act_ok(in, in_plane, in_ld) && bb_act_ok(out, out_plane, out_ld) && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 &&
                       in_ld >= c && out_ld >= c && k >= 1 && k <= 3 && stride >= 1 && stride <= 2 && pad >= 0 && 2 * pad < k + 1,
                   UAVSAL_EINVAL, "maxpool: bad arguments");
    const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;           // floor mode (torchvision resnet / vgg)
    UAVSAL_REQUIRE(ho > 0 && wo > 0, UAVSAL_EINVAL, "maxpool: empty output");
    const int64_t total = (int64_t)n * ho * wo * (c / 8);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    pool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Act{in, in_plane, in_ld}, n, h, w, c, k, stride, pad, ho, wo, ActW{out, out_plane, out_ld});
    return check_launch("maxpool");
}

extern "C" int uavsal_add_act(const uint16_t* a, int64_t a_plane, int a_ld, const uint16_t* b, int64_t b_plane, int b_ld, int64_t rows, int c,
                              int flags, uint16_t* out, int64_t out_plane, int out_ld, void* stream) {
    UAVSAL_REQUIRE(bb_act_ok(a, a_plane, a_ld) && bb_act_ok(b, b_plane, b_ld) && bb_act_ok(out, out_plane, out_ld) && c % 8 == 0 && c > 0 && rows > 0,
                   UAVSAL_EINVAL, "add_act: bad arguments");
    UAVSAL_REQUIRE(!(flags & ~(UAVSAL_F_RELU)), UAVSAL_ENOTSUP, "add_act: only the ReLU flag is supported");
    const int64_t total = rows * (c / 8);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    add_act_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Act{a, a_plane, a_ld}, Act{b, b_plane, b_ld}, rows, c, (flags & UAVSAL_F_RELU) ? 1 : 0,
                                                           ActW{out, out_plane, out_ld});
    return check_launch("add_act");
}

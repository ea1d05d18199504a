// Weight preparation on the device: BatchNorm (eval) folding, the bf16 hi/lo split and the layout change every conv kernel of
// this library expects, one launch per conv layer (plus its folded bias).  Replaces what the reference leaves to cuDNN at every
// forward (conv -> batch_norm as separate ATen ops, model.py:69-71, 94-95) and what the first version of this library did with
// a dozen eager tensor ops per layer at plan build.
//
//   w'[n][ci][tap] = w[n][ci][tap] * s[n],  s[n] = gamma[n] / sqrt(var[n] + eps),  b'[n] = beta[n] - mean[n] * s[n]
//
// (fp32, IEEE division / square root, no FMA contraction: bit-identical to the torch expressions).  Destination element
// (row r, k = tap * cin + ci): rows past cout and k past taps * cin are zero padding.  `gates` > 1 re-orders the rows of a
// ConvLSTM gate conv from g * ch + c (i, f, o, g blocks, model_convlstm.py:117) to c * gates + g, so that one epilogue thread
// owns the four gates of a (pixel, channel).
#include "common.cuh"

namespace uavsal {

template <int LAYOUT>
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int taps,
                                                           const float* __restrict__ g, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                                           const float* __restrict__ cbias, int gates, int n_pad, int k_pad,
                                                           void* __restrict__ out_w, float* __restrict__ out_bias) {
    const int64_t total = (int64_t)n_pad * k_pad;
    const int ch = gates > 1 ? cout / gates : cout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r, k;
        if (LAYOUT == UAVSAL_W_COLS_F32) { k = (int)(i / n_pad); r = (int)(i % n_pad); }      // dest [k][r]: r fastest
        else { r = (int)(i / k_pad); k = (int)(i % k_pad); }                                 // dest [r][k]: k fastest
        float v = 0.f;
        if (r < cout && k < taps * cin) {
            const int n = gates > 1 ? (r % gates) * ch + r / gates : r;
            const int tap = k / cin, ci = k % cin;
            v = __ldg(w + ((int64_t)n * cin + ci) * taps + tap);
            if (g) v = __fmul_rn(v, __fdiv_rn(__ldg(g + n), __fsqrt_rn(__fadd_rn(__ldg(var + n), eps))));
        }
        if (LAYOUT == UAVSAL_W_ROWS_SPLIT) {
            uint32_t hi, lo;
            split1(v, hi, lo);
            uint16_t* o = reinterpret_cast<uint16_t*>(out_w);
            o[i] = (uint16_t)hi;
            o[total + i] = (uint16_t)lo;
        } else {
            reinterpret_cast<float*>(out_w)[i] = v;
        }
    }
    if (out_bias && blockIdx.x == 0) {
        for (int r = threadIdx.x; r < n_pad; r += blockDim.x) {
            float b = 0.f;
            if (r < cout) {
                const int n = gates > 1 ? (r % gates) * ch + r / gates : r;
                b = cbias ? __ldg(cbias + n) : 0.f;
                if (g) {
                    const float s = __fdiv_rn(__ldg(g + n), __fsqrt_rn(__fadd_rn(__ldg(var + n), eps)));
                    b = __fadd_rn(__ldg(beta + n), __fmul_rn(__fsub_rn(b, __ldg(mean + n)), s));   // beta - mean * s when the conv has no bias
                }
            }
            out_bias[r] = b;
        }
    }
}

}  // namespace uavsal

using namespace uavsal;

extern "C" int uavsal_pack_weights(const float* w, int cout, int cin, int taps, const float* bn_weight, const float* bn_bias,
                                   const float* bn_mean, const float* bn_var, float bn_eps, const float* conv_bias, int gates,
                                   int layout, int n_pad, int k_pad, void* out_w, float* out_bias, void* stream) {
    UAVSAL_REQUIRE(w && out_w && cout > 0 && cin > 0 && taps > 0 && n_pad >= cout && k_pad >= cin * taps && gates >= 1 && cout % gates == 0,
                   UAVSAL_EINVAL, "pack_weights: bad arguments (cout=%d cin=%d taps=%d n_pad=%d k_pad=%d gates=%d)", cout, cin, taps, n_pad, k_pad, gates);
    const bool bn = bn_weight != nullptr;
    UAVSAL_REQUIRE(bn == (bn_bias != nullptr) && bn == (bn_mean != nullptr) && bn == (bn_var != nullptr), UAVSAL_EINVAL,
                   "pack_weights: pass all four BatchNorm tensors or none");
    UAVSAL_REQUIRE(layout == UAVSAL_W_ROWS_SPLIT || layout == UAVSAL_W_ROWS_F32 || layout == UAVSAL_W_COLS_F32, UAVSAL_EINVAL,
                   "pack_weights: unknown layout %d", layout);
    const int64_t total = (int64_t)n_pad * k_pad;
    const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    cudaStream_t s = (cudaStream_t)stream;
    if (layout == UAVSAL_W_ROWS_SPLIT)
        pack_weights_kernel<UAVSAL_W_ROWS_SPLIT><<<grid, 256, 0, s>>>(w, cout, cin, taps, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, conv_bias, gates, n_pad, k_pad, out_w, out_bias);
    else if (layout == UAVSAL_W_ROWS_F32)
        pack_weights_kernel<UAVSAL_W_ROWS_F32><<<grid, 256, 0, s>>>(w, cout, cin, taps, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, conv_bias, gates, n_pad, k_pad, out_w, out_bias);
    else
        pack_weights_kernel<UAVSAL_W_COLS_F32><<<grid, 256, 0, s>>>(w, cout, cin, taps, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, conv_bias, gates, n_pad, k_pad, out_w, out_bias);
    return check_launch("pack_weights");
}

// CUDA-core (FFMA) GEMM / implicit-GEMM 3x3 used for the HBM-bound small-K / small-N layers of the backbone
// and as the bring-up path the tcgen05 kernels are validated against.  fp32 weights, fp32 accumulation,
// split-bf16 activations in and out.  64x64 tile, BK=16, 256 threads, 4x4 micro-tile.
#include "common.cuh"

namespace uavsal {

enum { EPI_STD = 0, EPI_TWA = 1, EPI_LSTM = 2 };

struct SimtArgs {
    Act a0, a1;          // A sources (a1 only for the recurrences' virtual concat [x, h])
    int c0, c1;          // channels of each source (c1 = 0 when unused)
    int H, W;            // conv geometry (rows = nimg*H*W); unused for pointwise
    const float* w;      // [Ktot][N] fp32
    const float* bias;   // [N] or null
    int M, N, K;         // K = c0 (pointwise) or 9*(c0+c1) (conv)
    int flags;
    Act res;
    ActW out;
    // recurrences
    Act x;               // TWA: x_t rows
    Act hprev;           // TWA: h_{t-1} rows
    const float* c_in;   // LSTM: c_{t-1} [M][N/4]
    float* c_out;        // LSTM: c_t
};

template <int CONV, int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtArgs g) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int ctot = g.c0 + g.c1;

    // A loader: thread -> (row = tid/4, 4 consecutive k = (tid%4)*4)
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int gm = m0 + lrow;
    int py = 0, px = 0, pimg = 0;
    if (CONV) {
        const int hw = g.H * g.W;
        pimg = gm / hw;
        const int p = gm - pimg * hw;
        py = p / g.W;
        px = p - py * g.W;
    }
    // B loader: thread -> (k = tid/16, 4 consecutive n = (tid%16)*4)
    const int bk = tid >> 4, bn = (tid & 15) * 4;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += BK) {
        // ---- stage A ----
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        const int kk = k0 + lk;
        if (gm < g.M && kk < g.K) {
            if (!CONV) {
                load4(g.a0.p + (int64_t)gm * g.a0.ld + kk, g.a0.plane, av);
            } else {
                const int tap = kk / ctot;
                const int cc = kk - tap * ctot;
                const int y = py + tap / 3 - 1, x = px + tap % 3 - 1;
                if (y >= 0 && y < g.H && x >= 0 && x < g.W) {
                    const int64_t r = ((int64_t)pimg * g.H + y) * g.W + x;
                    if (cc < g.c0) load4(g.a0.p + r * g.a0.ld + cc, g.a0.plane, av);
                    else           load4(g.a1.p + r * g.a1.ld + (cc - g.c0), g.a1.plane, av);
                }
            }
        }
        // ---- stage B ----
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            const int k = k0 + bk, n = n0 + bn;
            if (k < g.K && n < g.N) bv = __ldg(reinterpret_cast<const float4*>(g.w + (int64_t)k * g.N + n));
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) As[lk + j][lrow] = av[j];
        *reinterpret_cast<float4*>(&Bs[bk][bn]) = bv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }

    // ---- epilogue ----
    const int n = n0 + tx * 4;
    if (n >= g.N) return;
    float bb[4] = {0.f, 0.f, 0.f, 0.f};
    if (g.bias) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
        bb[0] = b4.x; bb[1] = b4.y; bb[2] = b4.z; bb[3] = b4.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bb[j];
        if (EPI == EPI_STD) {
            if (g.flags & (UAVSAL_F_RELU6 | UAVSAL_F_RELU)) {
                const float cap = (g.flags & UAVSAL_F_RELU6) ? 6.f : 3.0e38f;      // ReLU6 (model.py:71) | ReLU (ResNet / VGG backbones)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = fminf(fmaxf(v[j], 0.f), cap);
            }
            if (g.flags & UAVSAL_F_RESIDUAL) {
                float r[4];
                load4(g.res.p + (int64_t)m * g.res.ld + n, g.res.plane, r);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] += r[j];
            }
            if (g.flags & UAVSAL_F_SIGMOID) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = sigmoid_acc(v[j]);
            }
            store4(g.out.p + (int64_t)m * g.out.ld + n, g.out.plane, v);
        } else if (EPI == EPI_TWA) {
            // h = i*x + (1-i)*h   (model_convlstm.py:283,290)
            float xv[4], hv[4];
            load4(g.x.p + (int64_t)m * g.x.ld + n, g.x.plane, xv);
            load4(g.hprev.p + (int64_t)m * g.hprev.ld + n, g.hprev.plane, hv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float gi = sigmoid_acc(v[j]);
                v[j] = gi * xv[j] + (1.f - gi) * hv[j];
            }
            store4(g.out.p + (int64_t)m * g.out.ld + n, g.out.plane, v);
        } else {
            // four interleaved gates i,f,o,g of channel n/4   (model_convlstm.py:117-124)
            const int ch = n >> 2, nch = g.N >> 2;
            const float gi = sigmoid_acc(v[0]), gf = sigmoid_acc(v[1]), go = sigmoid_acc(v[2]), gg = tanhf(v[3]);
            const float cn = gf * g.c_in[(int64_t)m * nch + ch] + gi * gg;
            g.c_out[(int64_t)m * nch + ch] = cn;
            store1(g.out.p + (int64_t)m * g.out.ld + ch, g.out.plane, go * tanhf(cn));
        }
    }
}

template <int CONV, int EPI>
static int launch_simt(const SimtArgs& g, cudaStream_t s, const char* what) {
    dim3 grid(div_up(g.M, 64), div_up(g.N, 64));
    gemm_simt_kernel<CONV, EPI><<<grid, 256, 0, s>>>(g);
    return check_launch(what);
}

// sequence helpers shared with the tcgen05 front-end (api in gemm_tc.cu decides which path runs)
int twa_sequence_simt(Act x, Act h0, int t_steps, int H, int W, int c, const float* w, ActW seq, cudaStream_t s) {
    const int64_t fr = (int64_t)H * W;
    for (int t = 0; t < t_steps; ++t) {
        SimtArgs g{};
        Act xt{x.p + t * fr * x.ld, x.plane, x.ld};
        Act hp = t == 0 ? h0 : Act{seq.p + (t - 1) * fr * seq.ld, seq.plane, seq.ld};
        g.a0 = xt; g.a1 = hp; g.c0 = c; g.c1 = c; g.H = H; g.W = W;
        g.w = w; g.bias = nullptr; g.M = (int)fr; g.N = c; g.K = 9 * 2 * c; g.flags = 0;
        g.x = xt; g.hprev = hp;
        g.out = ActW{seq.p + t * fr * seq.ld, seq.plane, seq.ld};
        int rc = launch_simt<1, EPI_TWA>(g, s, "twa_sequence(simt)");
        if (rc) return rc;
    }
    return 0;
}

int lstm_sequence_simt(Act x, Act h0, float* c_state, int b, int t_steps, int H, int W, int cin, int ch, const float* w,
                       const float* bias, ActW seq, cudaStream_t s) {
    // x rows are ordered (b, t, h, w); one launch per (t, b) keeps each image's rows contiguous
    const int64_t fr = (int64_t)H * W;
    for (int t = 0; t < t_steps; ++t) {
        for (int bi = 0; bi < b; ++bi) {
            SimtArgs g{};
            const int64_t row = ((int64_t)bi * t_steps + t) * fr;
            Act xt{x.p + row * x.ld, x.plane, x.ld};
            Act hp = t == 0 ? Act{h0.p + bi * fr * h0.ld, h0.plane, h0.ld}
                            : Act{seq.p + (row - fr) * seq.ld, seq.plane, seq.ld};
            g.a0 = xt; g.a1 = hp; g.c0 = cin; g.c1 = ch; g.H = H; g.W = W;
            g.w = w; g.bias = bias; g.M = (int)fr; g.N = 4 * ch; g.K = 9 * (cin + ch); g.flags = 0;
            g.c_in = c_state + bi * fr * ch; g.c_out = c_state + bi * fr * ch;
            g.out = ActW{seq.p + row * seq.ld, seq.plane, seq.ld};
            int rc = launch_simt<1, EPI_LSTM>(g, s, "convlstm_sequence(simt)");
            if (rc) return rc;
        }
    }
    return 0;
}

}  // namespace uavsal

using namespace uavsal;

static inline bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool act_ok(const void* p, int64_t plane, int ld) {
    return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 7) == 0 && (ld % 4) == 0 && (plane % 4) == 0 && plane >= 0;
}

extern "C" {

int uavsal_pw_gemm_simt(const uint16_t* a, int64_t a_plane, int a_ld, int m, int k, const float* wgt_f32, int n,
                        const float* bias, int flags, const uint16_t* res, int64_t res_plane, int res_ld, uint16_t* out,
                        int64_t out_plane, int out_ld, void* stream) {
    UAVSAL_REQUIRE(act_ok(a, a_plane, a_ld) && act_ok(out, out_plane, out_ld) && wgt_f32 && a16(wgt_f32) && m > 0 &&
                       k > 0 && n > 0 && k % 4 == 0 && n % 4 == 0 && a_ld >= k && out_ld >= n && (!bias || a16(bias)),
                   UAVSAL_EINVAL, "pw_gemm_simt: bad arguments (m=%d k=%d n=%d)", m, k, n);
    UAVSAL_REQUIRE(!(flags & UAVSAL_F_RESIDUAL) || act_ok(res, res_plane, res_ld), UAVSAL_EINVAL,
                   "pw_gemm_simt: residual requested without a residual tensor");
    SimtArgs g{};
    g.a0 = Act{a, a_plane, a_ld}; g.c0 = k; g.c1 = 0;
    g.w = wgt_f32; g.bias = bias; g.M = m; g.N = n; g.K = k; g.flags = flags;
    g.res = Act{res, res_plane, res_ld};
    g.out = ActW{out, out_plane, out_ld};
    return launch_simt<0, EPI_STD>(g, (cudaStream_t)stream, "pw_gemm_simt");
}

int uavsal_conv3x3_simt(const uint16_t* in, int64_t in_plane, int in_ld, int n, int h, int w, int c, const float* wgt_f32,
                        int cout, const float* bias, int flags, uint16_t* out, int64_t out_plane, int out_ld,
                        void* stream) {
    UAVSAL_REQUIRE(act_ok(in, in_plane, in_ld) && act_ok(out, out_plane, out_ld) && wgt_f32 && a16(wgt_f32) && n > 0 &&
                       c % 16 == 0 && cout % 4 == 0 && in_ld >= c && out_ld >= cout && (!bias || a16(bias)),
                   UAVSAL_EINVAL, "conv3x3_simt: bad arguments (c=%d cout=%d)", c, cout);
    SimtArgs g{};
    g.a0 = Act{in, in_plane, in_ld}; g.c0 = c; g.c1 = 0; g.H = h; g.W = w;
    g.w = wgt_f32; g.bias = bias; g.M = n * h * w; g.N = cout; g.K = 9 * c; g.flags = flags;
    g.out = ActW{out, out_plane, out_ld};
    return launch_simt<1, EPI_STD>(g, (cudaStream_t)stream, "conv3x3_simt");
}

}  // extern "C"

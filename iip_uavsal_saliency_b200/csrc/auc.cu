// AUC-Judd and the sampled AUCs (Borji / shuffled) of utils_score_torch.py:53-177, one CTA per (prediction, fixation map) pair.
//
// All three start like the reference: S = (pred - min) / (max - min + EPS) in fp32 (IEEE division, so the >= comparisons
// below see exactly the reference's values), S_fix = S at the pixels whose fixation channel is > 0.5.
//
// AUC-Judd (auc_j, :53-74): the reference loops over the sorted fixation values and counts `S >= thresh` over the whole map
// for each (n_fix passes over 230 400 pixels).  Here the thresholds are sorted once in shared memory (bitonic), every pixel
// finds by binary search the first threshold it reaches and bumps one histogram bin, and a prefix sum turns the bins into
// all the counts: one pass over the map.  tp / fp are built with the reference's fp32 expressions, the trapezoid sum runs in
// fp64.  Frames with more than kAucMaxFix fixations run the same steps over the caller's global-memory workspace.
//
// Sampled AUCs (auc_b :91-120, auc_s :135-159): the random pixel indices are drawn by the CALLER exactly as the reference
// draws them (np.random.randint under the caller's seed) and passed in; thread `rep` walks its column of samples, bins them
// against the thresholds k*step (k*step < max, fp64 as numpy's arange), and integrates its ROC; the mean of the n_rep areas
// is the score.  Degenerate pairs (all-zero map, no fixation) give NaN as in the reference.
#include <cmath>

#include "common.cuh"

namespace uavsal {

constexpr int kAucMaxFix = 4096;     // fixations per frame held in shared memory
constexpr int kAucBins = 64;         // thresholds k*step representable (S <= 1: step >= 1/62)
constexpr int kAucMaxRep = 128;
constexpr float kEpsF = 2.2204e-16f; // utils_score_torch.py:13

struct MinMax {
    float mn, mx;
};

// block-wide min / max of a map (all threads get the result)
__device__ __forceinline__ MinMax block_minmax(const float* __restrict__ p, int hw, float* red /* >= 64 floats */) {
    float mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
        const float v = __ldg(p + i);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[warp] = mn; red[32 + warp] = mx; }
    __syncthreads();
    mn = red[0]; mx = red[32];
    for (int w = 1; w < nw; ++w) { mn = fminf(mn, red[w]); mx = fmaxf(mx, red[32 + w]); }
    __syncthreads();
    return MinMax{mn, mx};
}

// Storage of the thresholds / histogram bins: shared memory up to kAucMaxFix fixations per frame, the caller's global workspace
// above that (dense fixation maps, mouse-click datasets): same algorithm, the bitonic passes then run over global memory (one CTA,
// __syncthreads between passes), which costs milliseconds per such frame instead of microseconds - the reference (:53-74) has no
// cap, so neither has this kernel.  ws layout per pair: float thr[np2(hw)] | unsigned hist[hw + 1].
__global__ void __launch_bounds__(1024) auc_judd_kernel(const float* __restrict__ pred, const float* __restrict__ truth, int hw,
                                                        float* __restrict__ ws, int64_t ws_pair_elems, int np2_hw,
                                                        float* __restrict__ out) {
    __shared__ float thr_s[kAucMaxFix];
    __shared__ unsigned hist_s[kAucMaxFix + 1];
    __shared__ float red[64];
    __shared__ double dred[32];
    __shared__ unsigned scan_base[32];
    __shared__ unsigned carry_s;
    __shared__ int nfix_s;
    const int pair = blockIdx.x, tid = threadIdx.x;
    const float* P = pred + (int64_t)pair * hw;
    const float* F = truth + ((int64_t)pair * 2 + 1) * hw;
    float* thr_g = ws ? ws + (int64_t)pair * ws_pair_elems : nullptr;
    unsigned* hist_g = ws ? reinterpret_cast<unsigned*>(thr_g + np2_hw) : nullptr;
    if (tid == 0) nfix_s = 0;
    const MinMax mm = block_minmax(P, hw, red);                 // (also orders the nfix_s store)
    const float d = __fadd_rn(__fsub_rn(mm.mx, mm.mn), kEpsF);
    for (int i = tid; i < hw; i += blockDim.x)
        if (__ldg(F + i) > 0.5f) {
            const int k = atomicAdd(&nfix_s, 1);
            const float s = __fdiv_rn(__fsub_rn(__ldg(P + i), mm.mn), d);
            if (k < kAucMaxFix) thr_s[k] = s;
            else if (thr_g) thr_g[k] = s;
        }
    __syncthreads();
    const int n_fix = nfix_s;
    const bool any_s = __fdiv_rn(__fsub_rn(mm.mx, mm.mn), d) > 0.f;
    // :54 NaN cases; n_fix == hw divides by n_pixels - n_fix = 0 in the reference (NaN / inf there as well).  The launcher
    // refuses maps above kAucMaxFix pixels without a workspace, so `big && !thr_g` cannot happen.
    const bool big = n_fix > kAucMaxFix;
    if (!any_s || n_fix == 0 || n_fix >= hw || (big && !thr_g)) {
        if (tid == 0) out[pair] = NAN;
        return;
    }
    float* thr = thr_s;
    unsigned* hist = hist_s;
    if (big) {
        thr = thr_g;
        hist = hist_g;
        for (int i = tid; i < kAucMaxFix; i += blockDim.x) thr_g[i] = thr_s[i];
    }
    // ---- sort the thresholds, descending (bitonic; padding = -inf sinks to the end) ----
    int np2 = 1;
    while (np2 < n_fix) np2 <<= 1;
    for (int i = n_fix + tid; i < np2; i += blockDim.x) thr[i] = -INFINITY;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const float a = thr[i], b = thr[l];
                    const bool desc = (i & k) == 0;
                    if (desc ? a < b : a > b) { thr[i] = b; thr[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i <= n_fix; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    // ---- one pass over the map: pixel with value s counts for every threshold <= s, i.e. for the suffix starting at the
    //      first index j with thr[j] <= s ----
    for (int i = tid; i < hw; i += blockDim.x) {
        const float s = __fdiv_rn(__fsub_rn(__ldg(P + i), mm.mn), d);
        int lo = 0, hi = n_fix;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (thr[mid] <= s) hi = mid; else lo = mid + 1;
        }
        atomicAdd(&hist[lo], 1u);
    }
    if (tid == 0) carry_s = 0;
    __syncthreads();
    // ---- inclusive prefix sum: above[i] = #{S >= thr[i]}; chunks of 4096 bins (4 consecutive bins per thread) with a carry ----
    for (int c0 = 0; c0 < n_fix; c0 += kAucMaxFix) {
        unsigned loc[4], run = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = c0 + tid * 4 + u;
            run += i < n_fix ? hist[i] : 0u;
            loc[u] = run;
        }
        unsigned inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
            if ((tid & 31) >= o) inc += v;
        }
        if ((tid & 31) == 31) scan_base[tid >> 5] = inc;
        __syncthreads();
        unsigned base = carry_s + inc - run;
        for (int w = 0; w < (tid >> 5); ++w) base += scan_base[w];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = c0 + tid * 4 + u;
            if (i < n_fix) hist[i] = base + loc[u];
        }
        if (tid == 1023) carry_s = base + run;
        __syncthreads();
    }
    // ---- ROC points with the reference's fp32 expressions (:66-72), trapezoid in fp64 ----
    const float fn = (float)n_fix, fneg = (float)(hw - n_fix);
    auto tp_at = [&](int i) -> float { return i == 0 ? 0.f : (i == n_fix + 1 ? 1.f : __fdiv_rn((float)i, fn)); };
    auto fp_at = [&](int i) -> float {
        if (i == 0) return 0.f;
        if (i == n_fix + 1) return 1.f;
        return __fdiv_rn((float)((long long)hist[i - 1] - (long long)i), fneg);   // above_th[i-1] - (i-1) - 1
    };
    double acc = 0.0;
    for (int i = tid; i <= n_fix; i += blockDim.x)
        acc += (double)__fmul_rn(__fsub_rn(fp_at(i + 1), fp_at(i)), __fadd_rn(tp_at(i + 1), tp_at(i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) dred[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += dred[w];
        out[pair] = (float)(s * 0.5);
    }
}

// bin of a value: the number of thresholds k*step (k < kAucBins) that it reaches
__device__ __forceinline__ int auc_bin(float v, double step) {
    int c = (int)floor((double)v / step);                      // estimate, then settle the comparison exactly as `v >= k*step`
    c = c < 0 ? 0 : (c > kAucBins - 1 ? kAucBins - 1 : c);
    while (c > 0 && !((double)v >= (double)c * step)) --c;
    while (c < kAucBins && (double)v >= (double)c * step) ++c;
    return c;                                                  // thresholds 0 .. c-1 are <= v
}

__global__ void __launch_bounds__(kAucMaxRep) auc_sampled_kernel(const float* __restrict__ pred, const float* __restrict__ truth, int hw,
                                                                 const int32_t* __restrict__ rand_idx, const int32_t* __restrict__ n_k,
                                                                 int max_k, int n_rep, double step, float* __restrict__ out) {
    __shared__ unsigned hist_f[kAucBins + 1];
    __shared__ unsigned short hist_r[kAucMaxRep][kAucBins + 2];
    __shared__ float red[64];
    __shared__ double area[kAucMaxRep];
    __shared__ int nfix_s;
    __shared__ unsigned maxf_bits;
    const int pair = blockIdx.x, tid = threadIdx.x;
    const float* P = pred + (int64_t)pair * hw;
    const float* F = truth + ((int64_t)pair * 2 + 1) * hw;
    if (tid == 0) { nfix_s = 0; maxf_bits = 0; }
    for (int i = tid; i <= kAucBins; i += blockDim.x) hist_f[i] = 0;
    const MinMax mm = block_minmax(P, hw, red);
    const float d = __fadd_rn(__fsub_rn(mm.mx, mm.mn), kEpsF);
    for (int i = tid; i < hw; i += blockDim.x)
        if (__ldg(F + i) > 0.5f) {
            const float s = __fdiv_rn(__fsub_rn(__ldg(P + i), mm.mn), d);    // >= 0
            atomicAdd(&nfix_s, 1);
            atomicAdd(&hist_f[auc_bin(s, step)], 1u);
            atomicMax(&maxf_bits, __float_as_uint(s));                       // non-negative floats order like their bit patterns
        }
    __syncthreads();
    const int n_fix = nfix_s, nk = n_k[pair];
    const bool any_s = __fdiv_rn(__fsub_rn(mm.mx, mm.mn), d) > 0.f;
    if (!any_s || n_fix == 0 || nk <= 0 || nk > max_k || nk > 65535) {       // :92 / :136 (NaN), and samples the caller did not provide
        if (tid == 0) out[pair] = NAN;
        return;
    }
    if (tid < n_rep) {
        unsigned short* h = hist_r[tid];
        for (int c = 0; c <= kAucBins; ++c) h[c] = 0;
        float mxr = __uint_as_float(maxf_bits);
        const int32_t* idx = rand_idx + (int64_t)pair * max_k * n_rep + tid;
        for (int k = 0; k < nk; ++k) {
            const int px = __ldg(idx + (int64_t)k * n_rep);
            const float s = __fdiv_rn(__fsub_rn(__ldg(P + px), mm.mn), d);
            mxr = fmaxf(mxr, s);
            ++h[auc_bin(s, step)];
        }
        // thresholds np.r_[0:mxr:step]: count = ceil(mxr / step) in fp64, values k*step; visited in descending order (:107, :147)
        int len = (int)ceil((double)mxr / step);
        len = len > kAucBins ? kAucBins : len;
        // suffix counts: #{v >= k*step} = sum of the bins above k
        unsigned cf = 0, cr = 0;
        for (int c = kAucBins; c > len; --c) { cf += hist_f[c]; cr += h[c]; }
        double x0 = 0.0, y0 = 0.0, a = 0.0;
        for (int k = len - 1; k >= 0; --k) {
            cf += hist_f[k + 1];
            cr += h[k + 1];
            const double y1 = (double)cf / (double)n_fix, x1 = (double)cr / (double)nk;
            a += (x1 - x0) * (y1 + y0) / 2.0;
            x0 = x1; y0 = y1;
        }
        a += (1.0 - x0) * (1.0 + y0) / 2.0;
        area[tid] = a;
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int r = 0; r < n_rep; ++r) s += area[r];
        out[pair] = (float)(s / (double)n_rep);
    }
}

}  // namespace uavsal

using namespace uavsal;

static int64_t auc_np2(int64_t v) {
    int64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

extern "C" int64_t uavsal_auc_judd_workspace(int n, int h, int w) {
    const int64_t hw = (int64_t)h * w;
    if (n <= 0 || hw <= kAucMaxFix) return 0;                      // every frame fits the shared-memory path
    return (int64_t)n * (auc_np2(hw) + ((hw + 1 + 3) & ~(int64_t)3)) * 4;
}

extern "C" int uavsal_auc_judd(const float* pred, const float* truth, int n, int h, int w, float* out, void* workspace,
                               int64_t workspace_bytes, void* stream) {
    UAVSAL_REQUIRE(pred && truth && out && n > 0 && h > 0 && w > 0 && (int64_t)h * w < (1 << 24), UAVSAL_EINVAL, "auc_judd: bad arguments");
    const int64_t need = uavsal_auc_judd_workspace(n, h, w);
    UAVSAL_REQUIRE(need == 0 || (workspace && workspace_bytes >= need && ((uintptr_t)workspace & 15) == 0), UAVSAL_EINVAL,
                   "auc_judd: maps of more than %d pixels need uavsal_auc_judd_workspace(n,h,w) = %lld bytes of 16-byte-aligned workspace "
                   "(a frame may hold that many fixations)", kAucMaxFix, (long long)need);
    const int hw = h * w;
    auc_judd_kernel<<<n, 1024, 0, (cudaStream_t)stream>>>(pred, truth, hw, need ? (float*)workspace : nullptr, need ? need / 4 / n : 0,
                                                           (int)auc_np2(hw), out);
    return check_launch("auc_judd");
}

extern "C" int uavsal_auc_sampled(const float* pred, const float* truth, int n, int h, int w, const int32_t* rand_idx, const int32_t* n_k,
                                  int max_k, int n_rep, double step, float* out, void* stream) {
    UAVSAL_REQUIRE(pred && truth && out && rand_idx && n_k && n > 0 && h > 0 && w > 0 && max_k > 0, UAVSAL_EINVAL, "auc_sampled: bad arguments");
    UAVSAL_REQUIRE(n_rep > 0 && n_rep <= kAucMaxRep && step * (kAucBins - 2) > 1.0 && max_k <= 65535, UAVSAL_ENOTSUP,
                   "auc_sampled: at most %d repetitions, %d samples per repetition, and a step of at least 1/%d", kAucMaxRep, 65535, kAucBins - 2);
    auc_sampled_kernel<<<n, kAucMaxRep, 0, (cudaStream_t)stream>>>(pred, truth, h * w, rand_idx, n_k, max_k, n_rep, step, out);
    return check_launch("auc_sampled");
}

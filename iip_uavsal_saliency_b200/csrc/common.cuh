// Shared helpers for the uavsal-b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/uavsal_b200.h"

namespace uavsal {

// ---- status / error reporting ---------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define UAVSAL_REQUIRE(cond, code, ...)                 \
    do {                                                \
        if (!(cond)) {                                  \
            ::uavsal::set_error(__VA_ARGS__);           \
            return (code);                              \
        }                                               \
    } while (0)

static inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// ---- the arena activation format: two bf16 planes (hi, lo) -----------------------------------------
// x ~= float(hi) + float(lo) with hi = bf16_rn(x), lo = bf16_rn(x - hi)  (16 significant bits).
struct Act {
    const uint16_t* p;   // hi plane, offset to the first channel of the slot
    int64_t plane;       // element offset hi -> lo (0: hi only)
    int ld;              // row pitch in elements
};
struct ActW {
    uint16_t* p;
    int64_t plane;
    int ld;
};

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t b) { return __uint_as_float(b << 16); }

__device__ __forceinline__ uint32_t f32_to_bf16_bits(float x) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
}

// split one fp32 into (hi, lo) bf16 bit patterns
__device__ __forceinline__ void split1(float x, uint32_t& hi, uint32_t& lo) {
    hi = f32_to_bf16_bits(x);
    lo = f32_to_bf16_bits(x - bf16_bits_to_f32(hi));
}

// unpack a 32-bit word holding two bf16 (little endian: element 0 in the low half)
__device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b) {
    a = __uint_as_float(w << 16);
    b = __uint_as_float(w & 0xFFFF0000u);
}

// load 8 consecutive channels (16-byte aligned) as fp32
__device__ __forceinline__ void load8(const uint16_t* hi, int64_t plane, float v[8]) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi));
    unpack2(h.x, v[0], v[1]);
    unpack2(h.y, v[2], v[3]);
    unpack2(h.z, v[4], v[5]);
    unpack2(h.w, v[6], v[7]);
    if (plane) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(hi + plane));
        float t[8];
        unpack2(l.x, t[0], t[1]);
        unpack2(l.y, t[2], t[3]);
        unpack2(l.z, t[4], t[5]);
        unpack2(l.w, t[6], t[7]);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
}

// load 4 consecutive channels (8-byte aligned) as fp32
__device__ __forceinline__ void load4(const uint16_t* hi, int64_t plane, float v[4]) {
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(hi));
    unpack2(h.x, v[0], v[1]);
    unpack2(h.y, v[2], v[3]);
    if (plane) {
        const uint2 l = __ldg(reinterpret_cast<const uint2*>(hi + plane));
        float t[4];
        unpack2(l.x, t[0], t[1]);
        unpack2(l.y, t[2], t[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += t[i];
    }
}

__device__ __forceinline__ float load1(const uint16_t* hi, int64_t plane) {
    float v = bf16_bits_to_f32(__ldg(hi));
    if (plane) v += bf16_bits_to_f32(__ldg(hi + plane));
    return v;
}

// split two fp32 into packed (hi, hi) / (lo, lo) bf16 pairs: one packed conversion per plane (F2FP) instead of two scalar F2F
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void store8(uint16_t* hi, int64_t plane, const float v[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
    if (plane) *reinterpret_cast<uint4*>(hi + plane) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void store4(uint16_t* hi, int64_t plane, const float v[4]) {
    uint32_t h[2], l[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    *reinterpret_cast<uint2*>(hi) = make_uint2(h[0], h[1]);
    if (plane) *reinterpret_cast<uint2*>(hi + plane) = make_uint2(l[0], l[1]);
}

__device__ __forceinline__ void store1(uint16_t* hi, int64_t plane, float v) {
    uint32_t h, l;
    split1(v, h, l);
    *hi = (uint16_t)h;
    if (plane) hi[plane] = (uint16_t)l;
}

// ---- "q16" rows: the ReLU6 output of a wide expand conv as 16-bit fixed point ------------------------------------------
// A value v in [0, 6] is stored as q = rne(v * 65535 / 6) (uint16) and read back as fp32(q * 6 / 65535): absolute error
// <= 4.6e-5, half the bytes of the fp32 rows.  Both directions avoid the conversion unit: adding 2^23 leaves rne(v * s) in the
// low mantissa bits, and fma(2^23 + q, 6/65535, -2^23 * 6/65535) is the single correctly rounded product q * 6/65535.
constexpr float kQ16Enc = 10922.5f;               // 65535 / 6 (exact)
constexpr float kQ16Dec = 6.0f / 65535.0f;
__device__ __forceinline__ uint32_t q16_pack2(float a, float b) {     // a, b already clamped to [0, 6]
    return __byte_perm(__float_as_uint(fmaf(a, kQ16Enc, 8388608.f)), __float_as_uint(fmaf(b, kQ16Enc, 8388608.f)), 0x5410);
}
__device__ __forceinline__ float2 q16_unpack2(uint32_t w) {
    const float2 f = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432)));
    return __ffma2_rn(f, make_float2(kQ16Dec, kQ16Dec), make_float2(-8388608.f * kQ16Dec, -8388608.f * kQ16Dec));
}

__device__ __forceinline__ float relu6f(float x) { return fminf(fmaxf(x, 0.f), 6.f); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// accurate variants used where the reference's sigmoid/tanh feed a recurrence
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch ------------------------------------------------------------------
// Every product-path kernel is launched with programmatic stream serialisation: it may start (block scheduling, barrier /
// TMEM / descriptor setup) while its predecessor drains, and calls pdl_wait() before its first access to global memory that
// another kernel produced or still reads.  pdl_trigger() lets the NEXT kernel begin its own prologue early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_pdl;      // uavsal_set_option key 6 (1 = on, default)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (g_pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cluster > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace uavsal

// Dev probe: the register layout of tcgen05.ld.16x256b (run on a B200: nvcc -gencode arch=compute_100a,code=sm_100a -o probe probe_tmem_ld.cu).
// Fills 16 TMEM columns with value = lane * 100 + column through the 32x32b shape (thread i <-> lane i), reads them back through
// 16x256b.x2 at lane offsets 0 and 16 and prints what every thread of warp 0 / warp 1 received.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void probe(float* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t trow = base + ((uint32_t)(warp * 32) << 16);
    float v[16];
    for (int c = 0; c < 16; ++c) v[c] = (float)((warp * 32 + lane) * 100 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(trow), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
                   "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int half = 0; half < 2; ++half) {
        float r[8];
        const uint32_t ta = base + ((uint32_t)(warp * 32 + half * 16) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "r"(ta) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) out[((warp * 2 + half) * 32 + lane) * 8 + j] = r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u) : "memory");
}

int main() {
    float* d;
    cudaMalloc(&d, 4 * 2 * 32 * 8 * sizeof(float));
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    static float h[4 * 2 * 32 * 8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int warp = 0; warp < 2; ++warp)
        for (int half = 0; half < 2; ++half) {
            printf("warp %d, lane offset %d: thread -> 8 registers (value = lane*100 + column)\n", warp, half * 16);
            for (int t = 0; t < 32; ++t) {
                printf("  t%02d:", t);
                for (int j = 0; j < 8; ++j) printf(" %6.0f", h[((warp * 2 + half) * 32 + t) * 8 + j]);
                printf("\n");
            }
        }
    return 0;
}

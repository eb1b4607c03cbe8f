// Micro-benchmark 4: which instruction classes of a lone warp get slower as soon as a second warp is
// active on another scheduler partition of the same SM?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t lds(uint32_t a){uint32_t v; asm volatile("ld.volatile.shared.u32 %0,[%1];":"=r"(v):"r"(a):"memory"); return v;}
__device__ __forceinline__ void sts(uint32_t a, uint32_t v){asm volatile("st.volatile.shared.u32 [%0],%1;"::"r"(a),"r"(v):"memory");}
template <int TEST, int NOISE> __global__ void k(long long *out, int n) {
    __shared__ uint32_t sm[1024];
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 4 + 4) & 4095;
    if (threadIdx.x == 0) sm[0] = 0;
    __syncthreads();
    const uint32_t ctrl = base;
    if ((threadIdx.x >> 5) != 15) {
        if ((int)(threadIdx.x >> 5) >= NOISE) return;
        uint32_t acc = threadIdx.x;
        while ((int)lds(ctrl) >= 0) { for (int q = 0; q < 64; q++) acc = acc * 1664525u + lds(base + 64 + ((acc >> 8) & 0x7f0)); }
        if (acc == 0x12345) out[5] = acc;
        return;
    }
    uint32_t x = threadIdx.x + 12345u, y = 7u, z = 1u;
    long long t0 = clock64();
    if (TEST == 0) {        // dependent IMAD chain, 16 per iteration
        for (int i = 0; i < n; i++) {
#pragma unroll
            for (int q = 0; q < 16; q++) { x = x * 1664525u + 1013904223u; asm volatile("" : "+r"(x)); }
        }
    } else if (TEST == 1) { // dependent IADD/LOP chain (alu pipe), 16 per iteration
        for (int i = 0; i < n; i++) {
#pragma unroll
            for (int q = 0; q < 16; q++) { x = (x + y) ^ z; asm volatile("" : "+r"(x)); }
        }
    } else if (TEST == 2) { // uniform data-dependent branch per 4 IMADs (never taken at run time)
        for (int i = 0; i < n; i++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                x = x * 1664525u + 1013904223u; x = x * 1664525u + 1013904223u; x = x * 1664525u + 1013904223u; x = x * 1664525u + 1013904223u;
                if (x == 0xdeadbeefu) { sts(ctrl + 8, x); y++; }
            }
        }
    } else if (TEST == 3) { // dependent LDS chain (pointer chasing in shared memory), 8 per iteration
        uint32_t a = 4;
        for (int i = 0; i < n; i++) {
#pragma unroll
            for (int q = 0; q < 8; q++) a = lds(base + a);
        }
        x = a;
    } else if (TEST == 4) { // 64-bit multiply / subtract / compare chain like the coder, no branch
        uint64_t R = 0xfedcba9876543210ull, D = 0x123456789ull;
        for (int i = 0; i < n; i++) {
#pragma unroll
            for (int q = 0; q < 4; q++) { uint64_t s = R >> 24; uint64_t rn = s * 16769000u; D = D - s * (uint64_t)(q + 1); R = rn | (1ull << 60); }
        }
        x = (uint32_t)(R ^ D);
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = x + y + z; }
    sts(ctrl, 0x80000000u);
}
template <int TEST, int NOISE> void run(long long *d, const char *name, int per) {
    long long h[2]; const int n = 100000;
    k<TEST, NOISE><<<1, 512>>>(d, n); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-34s noise %d: %.2f cycles per op\n", name, NOISE, (double)h[0] / n / per);
}
int main() {
    long long *d; cudaMalloc(&d, 128);
    run<0, 0>(d, "dependent IMAD", 16); run<0, 1>(d, "dependent IMAD", 16);
    run<1, 0>(d, "dependent IADD+LOP (2 ops)", 16); run<1, 1>(d, "dependent IADD+LOP (2 ops)", 16);
    run<2, 0>(d, "4 IMAD + untaken branch", 4); run<2, 1>(d, "4 IMAD + untaken branch", 4);
    run<3, 0>(d, "dependent LDS", 8); run<3, 1>(d, "dependent LDS", 8);
    run<4, 0>(d, "64-bit coder-like step", 4); run<4, 1>(d, "64-bit coder-like step", 4);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

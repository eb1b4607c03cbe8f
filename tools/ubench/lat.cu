// Micro-benchmark: per-iteration latency (SM cycles) of dependent chains executed by ONE warp,
// to calibrate the design of the range-coder warp.  nvcc -gencode arch=compute_100a,code=sm_100a lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lds(uint32_t a){uint32_t v; asm volatile("ld.volatile.shared.u32 %0,[%1];":"=r"(v):"r"(a):"memory"); return v;}
__device__ __forceinline__ void sts(uint32_t a, uint32_t v){asm volatile("st.volatile.shared.u32 [%0],%1;"::"r"(a),"r"(v):"memory");}
__device__ __forceinline__ uint4 lds4(uint32_t a){uint4 v; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3},[%4];":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"r"(a):"memory"); return v;}

template<int V> __global__ void k(long long* out, int n, uint32_t seed) {
    __shared__ __align__(16) uint32_t sm[64];
    const int lane = threadIdx.x;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    asm volatile("" : "+r"(base));
    sm[lane] = lane; sm[lane+32] = 0; __syncwarp();
    uint32_t x = seed + lane;
    uint64_t D = 0x123456789abcull + seed, R = 0xfedcba9876543210ull;
    const uint32_t L0 = lane * 500000u, L1 = (lane + 1) * 500000u;   // 32 intervals covering [0, 16M)
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (V == 0) { x = x * 3u + 1u; }                                             // IMAD chain
        if (V == 1) { x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31) + 1u; }       // SHFL chain
        if (V == 2) { x = __ballot_sync(0xffffffffu, (x & 1u) != 0) + lane; }        // VOTE chain
        if (V == 3) { uint32_t r; asm volatile("bfind.u32 %0,%1;":"=r"(r):"r"(x|1u)); x = r + x; } // BFIND
        if (V == 4) { sts(base + 4 * lane, x); x = lds(base + 4 * ((lane + 1) & 31)) + 1u; } // STS->LDS
        if (V == 5) { x = lds(base + 4 * (x & 31)) + 1u; }                          // LDS chain
        if (V == 6) { uint64_t s = R >> 24; uint64_t P = s * (x | 1u); R = (P <= D) ? R - P : R + P; x = (uint32_t)(R >> 13); } // 64-bit mul/cmp/sel
        if (V == 7) { // full shuffle-based step
            uint64_t s = R >> 24; uint64_t P0 = s * L0, P1 = s * L1; bool w = (P0 <= D) && (D < P1);
            uint32_t b = __ballot_sync(0xffffffffu, w); uint32_t src; asm("bfind.u32 %0,%1;":"=r"(src):"r"(b));
            uint64_t Dn = D - P0, Rn = P1 - P0;
            uint32_t dl = __shfl_sync(0xffffffffu,(uint32_t)Dn,src), dh = __shfl_sync(0xffffffffu,(uint32_t)(Dn>>32),src);
            uint32_t rl = __shfl_sync(0xffffffffu,(uint32_t)Rn,src), rh = __shfl_sync(0xffffffffu,(uint32_t)(Rn>>32),src);
            D = ((uint64_t)dh<<32)|dl; R = ((uint64_t)rh<<32)|rl;
            if ((R>>32)==0) { R <<= 32; D = (D<<32) | x; x = x*1664525u+1013904223u; }
        }
        if (V == 8) { // full smem-broadcast step
            uint64_t s = R >> 24; uint64_t P0 = s * L0, P1 = s * L1; bool w = (P0 <= D) && (D < P1);
            uint64_t Dn = D - P0, Rn = P1 - P0;
            asm volatile("{\n .reg .pred p;\n setp.ne.s32 p,%0,0;\n @p st.volatile.shared.v4.u32 [%1],{%2,%3,%4,%5};\n}\n"::"r"((int)w),"r"(base+128),"r"((uint32_t)Dn),"r"((uint32_t)(Dn>>32)),"r"((uint32_t)Rn),"r"((uint32_t)(Rn>>32)):"memory");
            uint4 st = lds4(base + 128);
            D = ((uint64_t)st.y<<32)|st.x; R = ((uint64_t)st.w<<32)|st.z;
            if ((R>>32)==0) { R <<= 32; D = (D<<32) | x; x = x*1664525u+1013904223u; }
        }
        if (V == 9) { // smem-broadcast step + ballot + branch + lane-0 store (as in coder_step)
            uint64_t s = R >> 24; uint64_t P0 = s * L0, P1 = s * L1; bool w = (P0 <= D) && (D < P1);
            uint64_t Dn = D - P0, Rn = P1 - P0;
            asm volatile("{\n .reg .pred p;\n setp.ne.s32 p,%0,0;\n @p st.volatile.shared.v4.u32 [%1],{%2,%3,%4,%5};\n}\n"::"r"((int)w),"r"(base+128),"r"((uint32_t)Dn),"r"((uint32_t)(Dn>>32)),"r"((uint32_t)Rn),"r"((uint32_t)(Rn>>32)):"memory");
            uint32_t b = __ballot_sync(0xffffffffu, w);
            if (b != 0) { uint4 st = lds4(base + 128); D = ((uint64_t)st.y<<32)|st.x; R = ((uint64_t)st.w<<32)|st.z; }
            else { D += 7; R |= 1ull<<40; }
            uint32_t src; asm("bfind.u32 %0,%1;":"=r"(src):"r"(b));
            asm volatile("{\n .reg .pred p;\n setp.eq.s32 p,%0,0;\n @p st.volatile.shared.u32 [%1],%2;\n}\n"::"r"(lane),"r"(base+160+4*(i&7)),"r"(src|(i<<8)):"memory");
            if ((R>>32)==0) { R <<= 32; D = (D<<32) | x; x = x*1664525u+1013904223u; }
        }
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = (long long)x + (long long)D + (long long)R; }
}

int main() {
    long long* d; cudaMalloc(&d, 16); long long h[2]; const int n = 200000;
    const char* names[] = {"IMAD chain","SHFL chain","VOTE chain","BFIND chain","STS->LDS","LDS chain","64b mul/cmp/sel","step: ballot+bfind+4shfl","step: smem broadcast","step: smem bcast+ballot+res store"};
#define RUN(V) k<V><<<1,32>>>(d,n,1); k<V><<<1,32>>>(d,n,2); cudaMemcpy(h,d,16,cudaMemcpyDeviceToHost); printf("%-36s %7.1f cycles/iter\n", names[V], (double)h[0]/n);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

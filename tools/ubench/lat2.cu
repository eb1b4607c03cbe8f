// Micro-benchmark 2: the coder loop as written in ccd_entropy.cu, in isolation and with
// "noise" warps (spinning on shared memory like waiting producers) on the other schedulers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t lds(uint32_t a){uint32_t v; asm volatile("ld.volatile.shared.u32 %0,[%1];":"=r"(v):"r"(a):"memory"); return v;}
__device__ __forceinline__ void sts(uint32_t a, uint32_t v){asm volatile("st.volatile.shared.u32 [%0],%1;"::"r"(a),"r"(v):"memory");}
__device__ __forceinline__ uint4 lds4(uint32_t a){uint4 v; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3},[%4];":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"r"(a):"memory"); return v;}
__device__ __forceinline__ uint2 lds2(uint32_t a){uint2 v; asm volatile("ld.volatile.shared.v2.u32 {%0,%1},[%2];":"=r"(v.x),"=r"(v.y):"r"(a):"memory"); return v;}

__device__ __noinline__ void advance(uint32_t& x, uint32_t& w) { x = x * 1664525u + 1013904223u; w = __shfl_sync(0xffffffffu, x, 3); }

template<int NOISE, int SLEEP> __global__ void k(long long* out, int n) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(smraw);
    asm volatile("" : "+r"(base));
    const uint32_t RING = 256, win = base + 1024, res = win + RING * 256, bc = base + 64, ctrl = base;
    for (int i = threadIdx.x; i < RING * 32; i += blockDim.x) {
        int l = i & 31; ((uint2*)(smraw + 1024))[i] = make_uint2(l * 500000u, (l + 1) * 500000u);
    }
    if (threadIdx.x == 0) { ((volatile uint32_t*)smraw)[0] = 0; ((volatile uint32_t*)smraw)[1] = 0x7fffffff; }
    __syncthreads();
    if (warp != 15) {
        if (warp >= NOISE) return;
        // noise: spin on the control word like a waiting producer
        if (SLEEP >= 0) { while ((int)lds(ctrl) >= 0) { if (SLEEP) __nanosleep(SLEEP); } }
        else { volatile long long* g = out + 8; while (g[0] == 0) { } }   // spin on GLOBAL memory instead
        return;
    }
    uint64_t D = 0x123456789abcull, R = 0xfedcba9876543210ull;
    uint32_t x = 12345 + lane, wnext = 7;
    const uint32_t lane8 = lane * 8, mask = RING - 1;
    uint32_t j = 0, limit = 0, end = n;
    long long t0 = clock64();
    while (j != end) {
        if ((int)(limit - j) <= 0) { do { uint32_t r = lds(ctrl + 4); limit = ((int)(r - end) > 0) ? end : r; } while ((int)(limit - j) <= 0); }
        uint2 a0 = lds2(win + (j & mask) * 256 + lane8), a1 = lds2(win + ((j + 1) & mask) * 256 + lane8);
        while (true) {
            if ((int)(limit - j) < 3) { uint32_t r = lds(ctrl + 4); limit = ((int)(r - end) > 0) ? end : r; if ((int)(limit - j) < 3) break; }
            const uint2 a2 = lds2(win + ((j + 2) & mask) * 256 + lane8);
            {
                uint64_t s = R >> 24; uint64_t P0 = s * a0.x, P1 = s * a0.y; bool w = (P0 <= D) && (D < P1);
                uint64_t Dn = D - P0, Rn = P1 - P0;
                asm volatile("{\n .reg .pred p;\n setp.ne.s32 p,%0,0;\n @p st.volatile.shared.v4.u32 [%1],{%2,%3,%4,%5};\n}\n"::"r"((int)w),"r"(bc),"r"((uint32_t)Dn),"r"((uint32_t)(Dn>>32)),"r"((uint32_t)Rn),"r"((uint32_t)(Rn>>32)):"memory");
                uint32_t b = __ballot_sync(0xffffffffu, w); uint32_t rw;
                if (b != 0) { uint4 st = lds4(bc); D = ((uint64_t)st.y<<32)|st.x; R = ((uint64_t)st.w<<32)|st.z; uint32_t src; asm("bfind.u32 %0,%1;":"=r"(src):"r"(b)); rw = src | ((j + 1) << 8); }
                else { D += 7; R |= 1ull<<40; rw = 0x80000000u; }
                asm volatile("{\n .reg .pred p;\n setp.eq.s32 p,%0,0;\n @p st.volatile.shared.u32 [%1],%2;\n}\n"::"r"(lane),"r"(res + (j & mask) * 4),"r"(rw):"memory");
                if ((R>>32)==0) { R <<= 32; D = (D<<32) | wnext; advance(x, wnext); }
            }
            j++; a0 = a1; a1 = a2;
        }
        j += 2;
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = (long long)x + (long long)D + (long long)R; sts(ctrl, 0x80000000u); out[8] = 1; __threadfence(); }
}
int main() {
    long long* d; cudaMalloc(&d, 128); cudaMemset(d, 0, 128); long long h[2]; const int n = 300000; size_t smem = 1024 + 256*256 + 256*4 + 64;
#define RUN(NOISE,SLEEP,name) cudaMemset(d, 0, 128); cudaFuncSetAttribute(k<NOISE,SLEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<NOISE,SLEEP><<<1,512,smem>>>(d,n); cudaDeviceSynchronize(); cudaMemset(d, 0, 128); k<NOISE,SLEEP><<<1,512,smem>>>(d,n); cudaMemcpy(h,d,16,cudaMemcpyDeviceToHost); printf("%-44s %7.1f cycles/symbol  (%s)\n", name, (double)h[0]/n, cudaGetErrorString(cudaGetLastError()));
    RUN(0,0,"coder loop alone")
    RUN(3,0,"+ 3 spinning warps (SMSP 0,1,2)")
    RUN(15,0,"+ 15 spinning warps (3 share the coder's SMSP)")
    RUN(15,200,"+ 15 warps spinning with nanosleep(200)")
    RUN(15,2000,"+ 15 warps spinning with nanosleep(2000)")
    RUN(15,20000,"+ 15 warps spinning with nanosleep(20000)")
    RUN(15,-1,"+ 15 warps spinning on GLOBAL memory")
    RUN(1,0,"+ 1 spinning warp (warp 0)")
    RUN(12,2000,"+ 12 warps, nanosleep(2000)")
    return 0;
}

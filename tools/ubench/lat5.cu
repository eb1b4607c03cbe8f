// Micro-benchmark 5 (thread-block cluster): the coder loop of lat3 (variant C) on its OWN SM while the
// producers of the stream (here: noise warps that also rewrite the hot ring through distributed shared
// memory) run on the other SM of a 2-CTA cluster.  Derived from micro-benchmark 3: candidate inner loops of the range-coder warp in isolation (one warp on an idle SM,
// every symbol is the mode): A = one branch per symbol (software pipelined), B = one branch per 3 symbols.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t lds(uint32_t a){uint32_t v; asm volatile("ld.volatile.shared.u32 %0,[%1];":"=r"(v):"r"(a):"memory"); return v;}
__device__ __forceinline__ void sts(uint32_t a, uint32_t v){asm volatile("st.volatile.shared.u32 [%0],%1;"::"r"(a),"r"(v):"memory");}
__device__ __forceinline__ uint4 lds4(uint32_t a){uint4 v; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3},[%4];":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"r"(a):"memory"); return v;}
__device__ __noinline__ uint4 slow(uint64_t D, uint64_t R, uint4 h, uint32_t j) {
    // stand-in for coder_slow: exact decision + renormalisation
    uint64_t s = R >> 24, lo = s * h.x, rn = s * h.y, dn = D - lo;
    if (!(dn < rn)) { dn = D % rn; }
    if ((rn >> 32) == 0) { rn = (rn << 32) | 0x9e3779b9u; dn = (dn << 32) | (j * 2654435761u); }
    return make_uint4((uint32_t)dn, (uint32_t)(dn >> 32), (uint32_t)rn, (uint32_t)(rn >> 32));
}
__device__ __forceinline__ bool spec(uint64_t D, uint64_t lo, uint64_t rn, uint32_t nL, uint32_t nP, uint64_t &dn, uint64_t &lo2, uint64_t &rn2) {
    uint32_t dn_lo, dn_hi, l2_lo, l2_hi, r2_lo, r2_hi, ok;
    asm volatile("{\n .reg .u32 s_lo, s_hi, t;\n .reg .pred p;\n sub.cc.u32 %0, %7, %9;\n subc.u32 %1, %8, %10;\n shf.r.clamp.b32 s_lo, %11, %12, 24;\n shr.u32 s_hi, %12, 24;\n"
                 " mul.lo.u32 %2, s_lo, %13;\n mul.hi.u32 t, s_lo, %13;\n mad.lo.u32 %3, s_hi, %13, t;\n mul.lo.u32 %4, s_lo, %14;\n mul.hi.u32 t, s_lo, %14;\n mad.lo.u32 %5, s_hi, %14, t;\n"
                 " setp.lt.u32 p, %1, %12;\n selp.u32 %6, 1, 0, p;\n}\n"
                 : "=r"(dn_lo), "=r"(dn_hi), "=r"(l2_lo), "=r"(l2_hi), "=r"(r2_lo), "=r"(r2_hi), "=r"(ok)
                 : "r"((uint32_t)D), "r"((uint32_t)(D >> 32)), "r"((uint32_t)lo), "r"((uint32_t)(lo >> 32)), "r"((uint32_t)rn), "r"((uint32_t)(rn >> 32)), "r"(nL), "r"(nP));
    dn = ((uint64_t)dn_hi << 32) | dn_lo; lo2 = ((uint64_t)l2_hi << 32) | l2_lo; rn2 = ((uint64_t)r2_hi << 32) | r2_lo;
    return ok != 0u;
}
#define PRE(H) do { const uint64_t s_ = R >> 24; lo = s_ * (H).x; rn = s_ * (H).y; } while (0)
#define STEP(JJ, H, HN) do { uint64_t dn_, lo2_, rn2_; if (spec(D, lo, rn, (HN).x, (HN).y, dn_, lo2_, rn2_)) { D = dn_; R = rn; lo = lo2_; rn = rn2_; } else { SLOW(JJ, H); PRE(HN); } } while (0)
#define SLOW(JJ, H) do { uint4 r_ = slow(D, R, (H), (JJ)); D = ((uint64_t)r_.y << 32) | r_.x; R = ((uint64_t)r_.w << 32) | r_.z; nslow++; } while (0)

#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank){uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;":"=r"(r):"r"(a),"r"(rank)); return r;}
__device__ __forceinline__ void st_cluster_v4(uint32_t a, uint4 v){asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};"::"r"(a),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}
__device__ __forceinline__ void st_cluster_u32(uint32_t a, uint32_t v){asm volatile("st.shared::cluster.u32 [%0], %1;"::"r"(a),"r"(v):"memory");}
// CODER_RANK: cluster rank that runs the coder warp (0 = same SM as the noise warps, 1 = its own SM)
template <int VAR, int NOISE, int CODER_RANK> __global__ void __cluster_dims__(2, 1, 1) k(long long *out, int n) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    extern __shared__ __align__(16) unsigned char smraw[];
    uint32_t base = (uint32_t)__cvta_generic_to_shared(smraw);
    asm volatile("" : "+r"(base));
    const uint32_t RING = 256, hot = base + 64, ctrl = base, mask = RING - 1;
    // mode with probability ~ 0.999: left = 2^23 - 2^22.., p = 2^24 * 0.9995
    for (int i = threadIdx.x; i < 512; i += blockDim.x) ((uint4 *)(smraw + 64))[i] = make_uint4(0u, 16777216u - 8000u - (i & 7), 0u, 16777216u);
    if (threadIdx.x == 0) { ((volatile uint32_t *)smraw)[1] = 0x7fffffff; }
    __syncthreads();
    cluster.sync();
    if ((threadIdx.x >> 5) != 15 || rank != CODER_RANK) {
        if (rank == CODER_RANK && rank != 0) return;  // the coder's SM keeps ONE resident warp (exited threads do not take part in cluster barriers)
        if (rank != 0 || (threadIdx.x >> 5) == 15 || (int)(threadIdx.x >> 5) >= NOISE) { cluster.sync(); return; }
        const uint32_t rhot = mapa(hot, CODER_RANK), rctrl = mapa(ctrl, CODER_RANK);
        (void)rctrl;
        // noise: work like a producer (shared-memory loads + integer math) until the coder is done
        uint32_t acc = threadIdx.x;
        uint32_t kk = threadIdx.x;
        while ((int)lds(ctrl) >= 0) {
            for (int q = 0; q < 64; q++) acc = acc * 1664525u + lds(hot + ((acc >> 8) & 0x0ff0));
            // one warp's worth of remote hot entries per round, like a producer chunk (8 symbols)
            if ((threadIdx.x & 31) < 8) { kk = (kk + 8) & 255; st_cluster_v4(rhot + kk * 16u, make_uint4(0u, 16777216u - 8000u - (kk & 7), 0u, 16777216u)); }
        }
        if (acc == 0x12345) out[5] = acc;
        cluster.sync();
        return;
    }
    uint64_t D = 0x0000123456789abcull, R = 0xfedcba9876543210ull, lo = 0, rn = 0;
    uint32_t j = 0, limit = 0, end = n, nslow = 0;
    auto hot_of = [&](uint32_t jj) { return lds4(hot + (jj & mask) * 16u); };
    auto refresh = [&]() { uint32_t r = lds(ctrl + 4); limit = ((int)(r - end) > 0) ? end : r; };
    long long t0 = clock64();
    while (j != end) {
        if ((int)(limit - j) <= 0) { do { refresh(); } while ((int)(limit - j) <= 0); }
        if ((int)(limit - j) >= 6) {
            uint4 a0 = hot_of(j), a1 = hot_of(j + 1), a2 = hot_of(j + 2);
            PRE(a0);
            if (VAR == 2) {
                // 6 symbols per iteration, roles of (a, b) swap instead of being copied; one address per half
                while (true) {
                    if ((int)(limit - j) < 9) { refresh(); if ((int)(limit - j) < 9) break; }
                    uint32_t o = hot + ((j + 3) & mask) * 16u;   // ring is mirrored: o + 32 never wraps
                    const uint4 b0 = lds4(o), b1 = lds4(o + 16), b2 = lds4(o + 32);
                    STEP(j, a0, a1); STEP(j + 1, a1, a2); STEP(j + 2, a2, b0);
                    o = hot + ((j + 6) & mask) * 16u;
                    a0 = lds4(o); a1 = lds4(o + 16); a2 = lds4(o + 32);
                    STEP(j + 3, b0, b1); STEP(j + 4, b1, b2); STEP(j + 5, b2, a0);
                    j += 6; sts(ctrl + 8, j);
                }
            } else
            while (true) {
                if ((int)(limit - j) < 6) { refresh(); if ((int)(limit - j) < 6) break; }
                const uint4 b0 = hot_of(j + 3), b1 = hot_of(j + 4), b2 = hot_of(j + 5);
                if (VAR == 0) {
                    STEP(j, a0, a1); STEP(j + 1, a1, a2); STEP(j + 2, a2, b0);
                } else {
                    uint64_t dn0, lo1, rn1, dn1, lo2, rn2, dn2, lo3, rn3;
                    const bool ok0 = spec(D, lo, rn, a1.x, a1.y, dn0, lo1, rn1);
                    const bool ok1 = spec(dn0, lo1, rn1, a2.x, a2.y, dn1, lo2, rn2);
                    const bool ok2 = spec(dn1, lo2, rn2, b0.x, b0.y, dn2, lo3, rn3);
                    if (ok0 & ok1 & ok2) { D = dn2; R = rn2; lo = lo3; rn = rn3; }
                    else {
                        if (ok0) { D = dn0; R = rn; } else { SLOW(j, a0); }
                        if (ok0 & ok1) { D = dn1; R = rn1; }
                        else { if (!ok0) { PRE(a1); lo1 = lo; rn1 = rn; } const uint64_t d_ = D - lo1; if ((uint32_t)(d_ >> 32) < (uint32_t)(rn1 >> 32)) { D = d_; R = rn1; } else { SLOW(j + 1, a1); } }
                        if (ok0 & ok1 & ok2) { } else { PRE(a2); const uint64_t d_ = D - lo; if ((uint32_t)(d_ >> 32) < (uint32_t)(rn >> 32)) { D = d_; R = rn; } else { SLOW(j + 2, a2); } }
                        PRE(b0);
                    }
                }
                j += 3; sts(ctrl + 8, j);
                a0 = b0; a1 = b1; a2 = b2;
            }
            // tail (not timed separately)
            { uint64_t d_ = D - lo; if ((uint32_t)(d_ >> 32) < (uint32_t)(rn >> 32)) { D = d_; R = rn; } else SLOW(j, a0); PRE(a1);
              d_ = D - lo; if ((uint32_t)(d_ >> 32) < (uint32_t)(rn >> 32)) { D = d_; R = rn; } else SLOW(j + 1, a1); PRE(a2);
              d_ = D - lo; if ((uint32_t)(d_ >> 32) < (uint32_t)(rn >> 32)) { D = d_; R = rn; } else SLOW(j + 2, a2); }
            j += 3;
        } else { uint4 a0 = hot_of(j); PRE(a0); uint64_t d_ = D - lo; if ((uint32_t)(d_ >> 32) < (uint32_t)(rn >> 32)) { D = d_; R = rn; } else SLOW(j, a0); j++; }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = (long long)D + (long long)R; out[2] = nslow; }
    st_cluster_u32(mapa(ctrl, 0), 0x80000000u);   // stop the noise warps (they poll their own CTA's control word)
    sts(ctrl, 0x80000000u);
    cluster.sync();
}
template <int VAR, int NOISE, int CODER_RANK> void run(long long *d, const char *name) {
    long long h[3]; const int n = 600000; size_t smem = 64 + 512 * 16;
    k<VAR, NOISE, CODER_RANK><<<2, 512, smem>>>(d, n); cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("%s coder on %s, noise warps %2d: %.1f cycles/symbol (slow %lld, chk %llx) %s\n", name, CODER_RANK ? "its own SM " : "the noise SM", NOISE, (double)h[0] / n, h[2], h[1], cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long long *d; cudaMalloc(&d, 128);
    run<2, 0, 0>(d, "C"); run<2, 11, 0>(d, "C"); run<2, 11, 1>(d, "C"); run<2, 3, 1>(d, "C");
    run<0, 11, 0>(d, "A"); run<0, 11, 1>(d, "A");
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

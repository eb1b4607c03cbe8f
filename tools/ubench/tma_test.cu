// Development check: cp.async.bulk.tensor with the descriptor (a) as a __grid_constant__ parameter, (b) in global memory,
// (c) in global memory + fence.proxy.tensormap; uint8 2-D box with arbitrary (negative / unaligned) coordinates.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                        CUtensorMapFloatOOBfill);
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap pm, const CUtensorMap *gm, int cx, int cy, int bw, int bh, unsigned *out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *bar = (uint64_t *)sm;
    unsigned char *dst = sm + 128;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(bw * bh)) : "memory");
        const uint64_t tm = MODE == 0 ? (uint64_t)&pm : (uint64_t)gm;
        if (MODE == 2) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tm), "r"(cx), "r"(cy), "r"(bar_a) : "memory");
    }
    uint32_t done = 0;
    while (!done) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar_a) : "memory");
    unsigned s = 0;
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) s += dst[i];
    atomicAdd(out, s);
}
int main() {
    const int W = 768, H = 512, BW = 48, BH = 26;
    unsigned char *h = new unsigned char[W * H];
    for (int i = 0; i < W * H; i++) h[i] = (unsigned char)((i * 7 + i / W) & 0xff);
    unsigned char *d; cudaMalloc(&d, W * H + 64); cudaMemcpy(d + 0, h, W * H, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN enc = (PFN)p;
    CUtensorMap tm;
    cuuint64_t gd[2] = {W, H}, gs[1] = {W}; cuuint32_t bx[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d\n", (int)r);
    CUtensorMap *gtm; cudaMalloc(&gtm, 256); cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    unsigned *out; cudaMalloc(&out, 4);
    const int coords[4][2] = {{0, 0}, {-16, -5}, {16, 11}, {176, 3}};
    for (int mode = 0; mode < 3; mode++)
        for (int c = 0; c < 4; c++) {
            cudaMemset(out, 0, 4);
            const int cx = coords[c][0], cy = coords[c][1];
            if (mode == 0) k<0><<<1, 128, 128 + BW * BH>>>(tm, gtm, cx, cy, BW, BH, out);
            if (mode == 1) k<1><<<1, 128, 128 + BW * BH>>>(tm, gtm, cx, cy, BW, BH, out);
            if (mode == 2) k<2><<<1, 128, 128 + BW * BH>>>(tm, gtm, cx, cy, BW, BH, out);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned got = 0; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
            unsigned want = 0;
            for (int y = 0; y < BH; y++) for (int x = 0; x < BW; x++) { int gy = cy + y, gx = cx + x; if (gy >= 0 && gy < H && gx >= 0 && gx < W) want += h[gy * W + gx]; }
            printf("mode %d coord (%d,%d): %s got %u want %u\n", mode, cx, cy, cudaGetErrorString(e), got, want);
            if (e != cudaSuccess) return 0;
        }
    return 0;
}

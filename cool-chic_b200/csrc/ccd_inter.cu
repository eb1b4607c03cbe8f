// ccd_inter.cu -- P/B frame reconstruction on sm_100a.
//
// Replaces the P/B branch of decode_frame (bitstream/decode.py:156-189):
//   apply_global_translation  globalmotion.py:151-160   integer shift, border clamp
//   Warper.forward            warp.py:294-397           TRAINING branch (no 1/64-pel flow rounding),
//                                                       filter_size >= 6: windowed sinc (warp.py:226-268)
//   alpha / beta blending     decode.py:171-189
// One thread per pixel; the global shift is folded into the gather indices, so the shifted
// references are never materialised.  Same operation order as oracle/ccoracle.c::warp_sinc
// (coefficients in double -> fp32, fp32 products and sequential sums, no fusion).
#include <cuda_runtime.h>
#include <stdint.h>

#include "ccd_internal.h"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <int N>
__device__ __forceinline__ void sinc_coeffs(float s, float (&c)[N]) {
    const float PIf = 3.14159265358979323846f;
    constexpr int lt = -(N / 2) + 1;
#pragma unroll
    for (int k = 0; k < N; k++) {
        const float arg = __fsub_rn(s, (float)(lt + k));
        const float pa = __fmul_rn(PIf, arg);
        const double win = cos((double)__fdiv_rn(pa, (float)N));
        const double sc = (arg == 0.0f) ? 1.0 : sin((double)pa) / (double)pa;
        c[k] = __fmul_rn((float)win, (float)sc);
    }
}

template <int N>
__device__ __forceinline__ void warp_pixel(const float *__restrict__ ref, int h, int w, int gx, int gy, float fx,
                                           float fy, int x, int y, float (&out)[3]) {
    constexpr int lt = -(N / 2) + 1;
    const float rx = floorf(fx), ry = floorf(fy);
    float cx[N], cy[N];
    sinc_coeffs<N>(__fsub_rn(fx, rx), cx);
    sinc_coeffs<N>(__fsub_rn(fy, ry), cy);
    int xs[N], ys[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        float nx = __fadd_rn(__fadd_rn((float)x, (float)(lt + k)), rx);
        float ny = __fadd_rn(__fadd_rn((float)y, (float)(lt + k)), ry);
        nx = nx < 0.0f ? 0.0f : (nx > (float)(w - 1) ? (float)(w - 1) : nx);
        ny = ny < 0.0f ? 0.0f : (ny > (float)(h - 1) ? (float)(h - 1) : ny);
        xs[k] = clampi((int)nx + gx, 0, w - 1);
        ys[k] = clampi((int)ny + gy, 0, h - 1);
    }
    const size_t plane = (size_t)h * w;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *p = ref + (size_t)c * plane;
        float col = 0.0f;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const float *row = p + (size_t)ys[i] * w;
            float line = 0.0f;
#pragma unroll
            for (int j = 0; j < N; j++) {
                const float t = __fmul_rn(__ldg(row + xs[j]), cx[j]);
                line = (j == 0) ? t : __fadd_rn(line, t);
            }
            const float t2 = __fmul_rn(line, cy[i]);
            col = (i == 0) ? t2 : __fadd_rn(col, t2);
        }
        out[c] = col;
    }
}

template <int N>
__global__ void k_inter_predict(const float *__restrict__ residue, const float *__restrict__ motion,
                                const float *__restrict__ ref0, const float *__restrict__ ref1, int h, int w,
                                int is_b, int g0x, int g0y, int g1x, int g1y, float *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t plane = (size_t)h * w, i = (size_t)y * w + x;
    float p0[3], p1[3] = {0.0f, 0.0f, 0.0f};
    warp_pixel<N>(ref0, h, w, g0x, g0y, motion[i], motion[plane + i], x, y, p0);
    float beta = 0.0f;
    if (is_b) {
        warp_pixel<N>(ref1, h, w, g1x, g1y, motion[2 * plane + i], motion[3 * plane + i], x, y, p1);
        beta = __fadd_rn(residue[4 * plane + i], 0.5f);
        beta = beta < 0.0f ? 0.0f : (beta > 1.0f ? 1.0f : beta);
    }
    float alpha = __fadd_rn(residue[3 * plane + i], 0.5f);
    alpha = alpha < 0.0f ? 0.0f : (alpha > 1.0f ? 1.0f : alpha);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float pred = p0[c];
        if (is_b) pred = __fadd_rn(__fmul_rn(beta, pred), __fmul_rn(__fsub_rn(1.0f, beta), p1[c]));
        out[c * plane + i] = __fadd_rn(__fmul_rn(alpha, pred), residue[c * plane + i]);
    }
}

}  // namespace

int ccd_inter_launch(const float *d_residue, const float *d_motion, const float *d_ref0, const float *d_ref1, int h,
                     int w, int is_b, const int32_t *gf, int filter_size, float *d_out, cudaStream_t st) {
    const dim3 block(32, 8, 1), grid((w + 31) / 32, (h + 7) / 8, 1);
#define LAUNCH(N)                                                                                              \
    k_inter_predict<N><<<grid, block, 0, st>>>(d_residue, d_motion, d_ref0, d_ref1, h, w, is_b, gf[0], gf[1],   \
                                               gf[2], gf[3], d_out)
    switch (filter_size) {
        case 6: LAUNCH(6); break;
        case 8: LAUNCH(8); break;
        case 10: LAUNCH(10); break;
        case 12: LAUNCH(12); break;
        default: return -1;
    }
#undef LAUNCH
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

// ccd_inter.cu -- P/B frame reconstruction on sm_100a.
//
// Replaces the P/B branch of decode_frame (bitstream/decode.py:156-189):
//   apply_global_translation  globalmotion.py:151-160   integer shift, border clamp
//   Warper.forward            warp.py:294-397           TRAINING branch (no 1/64-pel flow rounding),
//                                                       filter_size >= 6: windowed sinc (warp.py:226-268)
//                                                       filter_size 2 / 4: F.grid_sample bilinear / bicubic,
//                                                       border padding, align_corners (warp.py:50-56,314-334)
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

// torch.linspace(-1, 1, n)[i] as PyTorch's CPU kernel evaluates it (see oracle/ccoracle.c::lin_coord)
__device__ __forceinline__ float lin_coord(int n, int i) {
    const float step = __fdiv_rn(2.0f, (float)(n - 1));
    return (i < n / 2) ? __fmaf_rn(step, (float)i, -1.0f) : __fmaf_rn(-step, (float)(n - 1 - i), 1.0f);
}
__device__ __forceinline__ float cubic_near(float x) {
    const float a = __fsub_rn(__fmul_rn(1.25f, x), 2.25f);
    return __fmaf_rn(__fmul_rn(a, x), x, 1.0f);
}
__device__ __forceinline__ float cubic_far(float x) {
    const float a = __fadd_rn(__fmul_rn(-0.75f, x), 3.75f);
    const float b = __fadd_rn(__fmul_rn(a, x), -6.0f);
    return __fadd_rn(__fmul_rn(b, x), 3.0f);
}

// grid_sample(border, align_corners=True) of the globally shifted reference: N == 2 bilinear, N == 4 bicubic.
// Same operation order as oracle/ccoracle.c::warp_grid.
template <int N>
__device__ __forceinline__ void grid_pixel(const float *__restrict__ ref, int h, int w, int gx, int gy, float fx,
                                           float fy, int x, int y, float (&out)[3]) {
    const float sx = (float)((w - 1.0) / 2.0), sy = (float)((h - 1.0) / 2.0);
    const float g0 = __fadd_rn(lin_coord(w, x), __fdiv_rn(fx, sx)), g1 = __fadd_rn(lin_coord(h, y), __fdiv_rn(fy, sy));
    float ix = __fmul_rn(__fadd_rn(g0, 1.0f), sx), iy = __fmul_rn(__fadd_rn(g1, 1.0f), sy);
    const size_t plane = (size_t)h * w;
    if (N == 2) {
        ix = fminf((float)(w - 1), fmaxf(ix, 0.0f));
        iy = fminf((float)(h - 1), fmaxf(iy, 0.0f));
        const float xw = floorf(ix), yn = floorf(iy);
        const float ww = __fsub_rn(ix, xw), e = __fsub_rn(1.0f, ww), nn = __fsub_rn(iy, yn), s = __fsub_rn(1.0f, nn);
        const float k_nw = __fmul_rn(s, e), k_ne = __fmul_rn(s, ww), k_sw = __fmul_rn(nn, e), k_se = __fmul_rn(nn, ww);
        const int x0 = (int)xw, y0 = (int)yn, x1 = x0 + 1, y1 = y0 + 1;
        const int xa = clampi(x0 + gx, 0, w - 1), xb = clampi((x1 < w ? x1 : w - 1) + gx, 0, w - 1);
        const int ya = clampi(y0 + gy, 0, h - 1), yb = clampi((y1 < h ? y1 : h - 1) + gy, 0, h - 1);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float *p = ref + (size_t)c * plane;
            const float v00 = __ldg(p + (size_t)ya * w + xa);
            const float v01 = x1 < w ? __ldg(p + (size_t)ya * w + xb) : 0.0f;
            const float v10 = y1 < h ? __ldg(p + (size_t)yb * w + xa) : 0.0f;
            const float v11 = (x1 < w && y1 < h) ? __ldg(p + (size_t)yb * w + xb) : 0.0f;
            float r = __fmul_rn(v00, k_nw);
            r = __fmaf_rn(v01, k_ne, r);
            r = __fmaf_rn(v10, k_sw, r);
            r = __fmaf_rn(v11, k_se, r);
            out[c] = r;
        }
    } else {
        const float fxx = floorf(ix), fyy = floorf(iy);
        const float tx = __fsub_rn(ix, fxx), ty = __fsub_rn(iy, fyy);
        const float cx[4] = {cubic_far(__fadd_rn(tx, 1.0f)), cubic_near(tx), cubic_near(__fsub_rn(1.0f, tx)),
                             cubic_far(__fsub_rn(2.0f, tx))};
        const float cy[4] = {cubic_far(__fadd_rn(ty, 1.0f)), cubic_near(ty), cubic_near(__fsub_rn(1.0f, ty)),
                             cubic_far(__fsub_rn(2.0f, ty))};
        int xs[4], ys[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float nx = fminf((float)(w - 1), fmaxf(__fadd_rn(fxx, (float)(k - 1)), 0.0f));
            const float ny = fminf((float)(h - 1), fmaxf(__fadd_rn(fyy, (float)(k - 1)), 0.0f));
            xs[k] = clampi((int)nx + gx, 0, w - 1);
            ys[k] = clampi((int)ny + gy, 0, h - 1);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float *p = ref + (size_t)c * plane;
            float rows[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float *q = p + (size_t)ys[i] * w;
                float r = __fmul_rn(cx[0], __ldg(q + xs[0]));
                r = __fmaf_rn(cx[1], __ldg(q + xs[1]), r);
                r = __fmaf_rn(cx[2], __ldg(q + xs[2]), r);
                r = __fmaf_rn(cx[3], __ldg(q + xs[3]), r);
                rows[i] = r;
            }
            float a = __fmul_rn(cy[0], rows[0]);
            a = __fmaf_rn(cy[1], rows[1], a);
            a = __fmaf_rn(cy[2], rows[2], a);
            a = __fmaf_rn(cy[3], rows[3], a);
            out[c] = a;
        }
    }
}

template <int N>
__device__ __forceinline__ void predict_pixel(const float *__restrict__ ref, int h, int w, int gx, int gy, float fx,
                                              float fy, int x, int y, float (&out)[3]) {
    if constexpr (N <= 4)
        grid_pixel<N>(ref, h, w, gx, gy, fx, fy, x, y, out);
    else
        warp_pixel<N>(ref, h, w, gx, gy, fx, fy, x, y, out);
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
    predict_pixel<N>(ref0, h, w, g0x, g0y, motion[i], motion[plane + i], x, y, p0);
    float beta = 0.0f;
    if (is_b) {
        predict_pixel<N>(ref1, h, w, g1x, g1y, motion[2 * plane + i], motion[3 * plane + i], x, y, p1);
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
        case 2: if (h < 2 || w < 2) return -1; LAUNCH(2); break;
        case 4: if (h < 2 || w < 2) return -1; LAUNCH(4); break;
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

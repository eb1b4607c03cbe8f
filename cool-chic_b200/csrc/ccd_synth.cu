// ccd_synth.cu -- float tail of the Cool-chic decoder on sm_100a: learned upsampling,
// synthesis, final resize, frame quantisation.
//
// Replaces (reference, fp32 PyTorch CPU):
//   core/upsampling.py:189-196   pre-concat conv, TRAIN-mode kron form (zero pad, + x)
//   core/upsampling.py:306-325   transposed conv, TRAIN-mode kron form (replicate pad 4, crop 11)
//   core/upsampling.py:463-500   Upsampling.forward cascade
//   core/synthesis.py:61-76      SynthesisConv2d.forward (replicate pad, conv+bias, +x, ReLU)
//   core/synthesis.py:272-294    Synthesis.forward (trunk + stabiliser, output_transform)
//   component/coolchic.py:187-192 final F.interpolate + crop
//   bitstream/decode.py:191-206  round / 420 average / clamp / round
//
// Canonical fp32 order (identical to oracle/ccoracle.c, so GPU == oracle bit for bit):
// every output is acc = init; for ci, for ky, for kx: acc = fmaf(w, x, acc).
#include <cuda_runtime.h>
#include <stdint.h>

#include "ccd_internal.h"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct K1d {
    float w[16];
};

__global__ void k_ups_first(const int8_t *__restrict__ lat, size_t n, float *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)lat[i];
}

// hi = conv2d(x, kron(w, w), zero padding k/2) + x      (x = int8 latent as float)
__global__ void k_ups_pre(const int8_t *__restrict__ lat, int h, int w, K1d kw, int k, float *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int p = k / 2;
    float acc = 0.0f;
    for (int a = 0; a < k; a++) {
        const int yy = y + a - p;
        if (yy < 0 || yy >= h) continue;
        for (int b = 0; b < k; b++) {
            const int xx = x + b - p;
            if (xx < 0 || xx >= w) continue;
            const float kk = __fmul_rn(kw.w[a], kw.w[b]);
            acc = __fmaf_rn(kk, (float)lat[(size_t)yy * w + xx], acc);
        }
    }
    out[(size_t)y * w + x] = __fadd_rn(acc, (float)lat[(size_t)y * w + x]);
}

// transposed conv stride 2 on the replicate-padded input, cropped (see oracle convt_kron)
__global__ void k_ups_convt(const float *__restrict__ in, int h, int w, K1d kw, int k, float *__restrict__ out,
                            int ht, int wt) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int u = blockIdx.y * blockDim.y + threadIdx.y;
    const int c = blockIdx.z;
    if (v >= wt || u >= ht) return;
    const int P0 = k / 2, C = 2 * P0 - 1 + k / 2;
    const float *src = in + (size_t)c * h * w;
    const int o1 = u + C, o2 = v + C;
    const int i1_lo = (o1 - (k - 1) < 0) ? 0 : (o1 - (k - 1) + 1) / 2, i1_hi = o1 / 2;
    const int i2_lo = (o2 - (k - 1) < 0) ? 0 : (o2 - (k - 1) + 1) / 2, i2_hi = o2 / 2;
    float acc = 0.0f;
    for (int i1 = i1_lo; i1 <= i1_hi; i1++) {
        const int a = o1 - 2 * i1;
        const int r = clampi(i1 - P0, 0, h - 1);
        for (int i2 = i2_lo; i2 <= i2_hi; i2++) {
            const int b = o2 - 2 * i2;
            const int cc = clampi(i2 - P0, 0, w - 1);
            const float kk = __fmul_rn(kw.w[a], kw.w[b]);
            acc = __fmaf_rn(kk, src[(size_t)r * w + cc], acc);
        }
    }
    out[((size_t)c * ht + u) * wt + v] = acc;
}

// generic SynthesisConv2d: one thread = one pixel, loops over output channels
__global__ void k_syn_layer(const float *__restrict__ in, int h, int w, int cin, int cout, int k, int residual,
                            int relu, const float *__restrict__ wt, const float *__restrict__ bias,
                            float *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int p = (k - 1) / 2;
    const size_t plane = (size_t)h * w;
    for (int co = 0; co < cout; co++) {
        float acc = __ldg(bias + co);
        for (int ci = 0; ci < cin; ci++) {
            for (int ky = 0; ky < k; ky++) {
                const int yy = clampi(y + ky - p, 0, h - 1);
                for (int kx = 0; kx < k; kx++) {
                    const int xx = clampi(x + kx - p, 0, w - 1);
                    acc = __fmaf_rn(__ldg(wt + (((size_t)co * cin + ci) * k + ky) * k + kx),
                                    in[ci * plane + (size_t)yy * w + xx], acc);
                }
            }
        }
        if (residual) acc = __fadd_rn(acc, in[co * plane + (size_t)y * w + x]);
        if (relu) acc = fmaxf(acc, 0.0f);
        out[co * plane + (size_t)y * w + x] = acc;
    }
}

// two fused 1x1 layers (cin -> chid -> cout), hidden activations stay in registers.
// Same summation order as two sequential generic layers.
template <int CIN_MAX, int COUT_MAX>
__global__ void k_syn_pw2(const float *__restrict__ in, size_t plane, int cin, int chid, int cout, int relu0,
                          int relu1, const float *__restrict__ w0, const float *__restrict__ b0,
                          const float *__restrict__ w1, const float *__restrict__ b1, float *__restrict__ out) {
    extern __shared__ float s_w[];
    float *sw0 = s_w;                   // [chid][cin]
    float *sb0 = sw0 + chid * cin;      // [chid]
    float *sw1 = sb0 + chid;            // [cout][chid]
    float *sb1 = sw1 + cout * chid;     // [cout]
    for (int i = threadIdx.x; i < chid * cin; i += blockDim.x) sw0[i] = w0[i];
    for (int i = threadIdx.x; i < chid; i += blockDim.x) sb0[i] = b0[i];
    for (int i = threadIdx.x; i < cout * chid; i += blockDim.x) sw1[i] = w1[i];
    for (int i = threadIdx.x; i < cout; i += blockDim.x) sb1[i] = b1[i];
    __syncthreads();
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane) return;
    float x[CIN_MAX], o[COUT_MAX];
#pragma unroll
    for (int i = 0; i < CIN_MAX; i++) x[i] = (i < cin) ? in[(size_t)i * plane + p] : 0.0f;
#pragma unroll
    for (int c = 0; c < COUT_MAX; c++) o[c] = (c < cout) ? sb1[c] : 0.0f;
    for (int hcn = 0; hcn < chid; hcn++) {
        float a = sb0[hcn];
#pragma unroll
        for (int i = 0; i < CIN_MAX; i++)
            if (i < cin) a = __fmaf_rn(sw0[hcn * cin + i], x[i], a);
        if (relu0) a = fmaxf(a, 0.0f);
#pragma unroll
        for (int c = 0; c < COUT_MAX; c++)
            if (c < cout) o[c] = __fmaf_rn(sw1[c * chid + hcn], a, o[c]);
    }
#pragma unroll
    for (int c = 0; c < COUT_MAX; c++)
        if (c < cout) out[(size_t)c * plane + p] = relu1 ? fmaxf(o[c], 0.0f) : o[c];
}

__global__ void k_add(float *__restrict__ a, const float *__restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = __fadd_rn(a[i], b[i]);
}

// legacy "nearest": src = min(floor(dst * (in/out as fp32)), in - 1)
__global__ void k_resize_nearest(const float *__restrict__ in, int h, int w, float *__restrict__ out, int H, int W,
                                 float sy, float sx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int c = blockIdx.z;
    if (x >= W || y >= H) return;
    int yy = (int)floorf(__fmul_rn((float)y, sy));
    int xx = (int)floorf(__fmul_rn((float)x, sx));
    yy = yy > h - 1 ? h - 1 : yy;
    xx = xx > w - 1 ? w - 1 : xx;
    out[((size_t)c * H + y) * W + x] = in[((size_t)c * h + yy) * w + xx];
}

// F.interpolate(mode = bilinear | bicubic, align_corners=False, antialias=False), same operation order as
// oracle/ccoracle.c::resize_torch (PyTorch's separable CPU kernel): per axis src = fma(scale, dst + 0.5, -0.5),
// taps clamped to the grid, A = -0.75 cubic coefficients; value = sum_i wy[i] * (sum_j wx[j] * v[i][j]) as fma chains.
__device__ __forceinline__ float cubic_near(float x) {
    const float a = __fsub_rn(__fmul_rn(1.25f, x), 2.25f);
    return __fmaf_rn(__fmul_rn(a, x), x, 1.0f);
}
__device__ __forceinline__ float cubic_far(float x) {
    const float a = __fadd_rn(__fmul_rn(-0.75f, x), 3.75f);
    const float b = __fadd_rn(__fmul_rn(a, x), -6.0f);
    return __fadd_rn(__fmul_rn(b, x), 3.0f);
}
template <int MODE>  // 1 bilinear (2 taps), 2 bicubic (4 taps)
__device__ __forceinline__ void resize_taps(int n_in, int i, float scale, int (&idx)[4], float (&wt)[4]) {
    float src = __fmaf_rn(scale, __fadd_rn((float)i, 0.5f), -0.5f);
    if (MODE == 1 && src < 0.0f) src = 0.0f;
    int i0 = (int)floorf(src);
    if (i0 > n_in - 1) i0 = n_in - 1;
    float lam = __fsub_rn(src, (float)i0);
    lam = lam < 0.0f ? 0.0f : (lam > 1.0f ? 1.0f : lam);
    if (MODE == 1) {
        idx[0] = i0;
        idx[1] = i0 + (i0 < n_in - 1 ? 1 : 0);
        wt[0] = __fsub_rn(1.0f, lam);
        wt[1] = lam;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) idx[j] = clampi(i0 + j - 1, 0, n_in - 1);
        const float x2 = __fsub_rn(1.0f, lam);
        wt[0] = cubic_far(__fadd_rn(lam, 1.0f));
        wt[1] = cubic_near(lam);
        wt[2] = cubic_near(x2);
        wt[3] = cubic_far(__fadd_rn(x2, 1.0f));
    }
}
template <int MODE>
__global__ void k_resize_torch(const float *__restrict__ in, int h, int w, float *__restrict__ out, int H, int W,
                               float sy, float sx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int c = blockIdx.z;
    if (x >= W || y >= H) return;
    constexpr int N = MODE == 1 ? 2 : 4;
    int iy[4], ix[4];
    float wy[4], wx[4];
    resize_taps<MODE>(h, y, sy, iy, wy);
    resize_taps<MODE>(w, x, sx, ix, wx);
    const float *p = in + (size_t)c * h * w;
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const float *row = p + (size_t)iy[i] * w;
        float r = __fmul_rn(wx[0], __ldg(row + ix[0]));
#pragma unroll
        for (int j = 1; j < N; j++) r = __fmaf_rn(wx[j], __ldg(row + ix[j]), r);
        acc = (i == 0) ? __fmul_rn(wy[0], r) : __fmaf_rn(wy[i], r, acc);
    }
    out[((size_t)c * H + y) * W + x] = acc;
}

// Common randomness (core/noise.py:18-55): sample k of the Park-Miller sequence, Box-Muller in f64.
// seed_j = a^j * seed_0 mod m is evaluated directly (square-and-multiply) instead of serially.
__device__ __forceinline__ uint64_t lcg_pow(uint64_t e) {
    const uint64_t m = 2147483647ULL;
    uint64_t r = 1, b = 16807ULL;
    while (e) {
        if (e & 1) r = (r * b) % m;
        b = (b * b) % m;
        e >>= 1;
    }
    return r;
}
__global__ void k_cr_noise(float *__restrict__ out, size_t first, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t m = 2147483647ULL;
    const size_t k = first + i;
    const uint64_t s1 = (lcg_pow(2 * k + 1) * 18101995ULL) % m;
    const uint64_t s2 = (s1 * 16807ULL) % m;
    const double u1 = (double)s1 / (double)m, u2 = (double)s2 / (double)m;
    out[i] = (float)(sqrt(-2 * log(u1)) * cos(2 * 3.14159265359 * u2));
}

__device__ __forceinline__ float quant(float v, float M) { return __fdiv_rn(rintf(__fmul_rn(M, v)), M); }
__device__ __forceinline__ float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

__global__ void k_finish_444(const float *__restrict__ in, size_t n, float M, float *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = quant(clamp01(quant(in[i], M)), M);
}

__global__ void k_finish_420_uv(const float *__restrict__ in, int h, int w, float M, float *__restrict__ ou,
                                float *__restrict__ ov) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int c = blockIdx.z;  // 0: u, 1: v
    const int h2 = h / 2, w2 = w / 2;
    if (x >= w2 || y >= h2) return;
    const float *p = in + (size_t)(c + 1) * h * w;
    float s = 0.0f;
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) s = __fadd_rn(s, quant(p[(size_t)(2 * y + dy) * w + 2 * x + dx], M));
    float v = __fdiv_rn(s, 4.0f);
    (c == 0 ? ou : ov)[(size_t)y * w2 + x] = quant(clamp01(v), M);
}

inline dim3 grid2(int w, int h, int z = 1) { return dim3((w + 31) / 32, (h + 7) / 8, z); }
const dim3 kBlock2(32, 8, 1);

K1d make_k1d(const float *w1d, int k) {
    K1d r;
    for (int i = 0; i < 16; i++) r.w[i] = i < k ? w1d[i] : 0.0f;
    return r;
}

}  // namespace

int ccd_ups_first(const int8_t *d_lat, int h, int w, float *d_out, cudaStream_t st) {
    size_t n = (size_t)h * w;
    k_ups_first<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_lat, n, d_out);
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

int ccd_ups_pre(const int8_t *d_lat, int h, int w, const float *w1d, int k, float *d_out, cudaStream_t st) {
    k_ups_pre<<<grid2(w, h), kBlock2, 0, st>>>(d_lat, h, w, make_k1d(w1d, k), k, d_out);
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

int ccd_ups_convt(const float *d_in, int c, int h, int w, const float *w1d, int k, float *d_out, int ht,
                  int wt, cudaStream_t st) {
    k_ups_convt<<<grid2(wt, ht, c), kBlock2, 0, st>>>(d_in, h, w, make_k1d(w1d, k), k, d_out, ht, wt);
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

int ccd_syn_layer(const float *d_in, int h, int w, const SynLayerDev &L, float *d_out, cudaStream_t st) {
    k_syn_layer<<<grid2(w, h), kBlock2, 0, st>>>(d_in, h, w, L.cin, L.cout, L.k, L.residual, L.relu, L.w, L.b,
                                                 d_out);
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

int ccd_syn_pointwise2(const float *d_in, int h, int w, const SynLayerDev &L0, const SynLayerDev &L1,
                       float *d_out, cudaStream_t st) {
    // preconditions checked by the caller: k == 1, no residual, cin <= 16, cout(L1) <= 8
    size_t plane = (size_t)h * w;
    size_t smem = ((size_t)L0.cout * L0.cin + L0.cout + (size_t)L1.cout * L0.cout + L1.cout) * sizeof(float);
    k_syn_pw2<16, 8><<<(unsigned)((plane + 255) / 256), 256, smem, st>>>(
        d_in, plane, L0.cin, L0.cout, L1.cout, L0.relu, L1.relu, L0.w, L0.b, L1.w, L1.b, d_out);
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

int ccd_syn_add(float *d_a, const float *d_b, size_t n, cudaStream_t st) {
    k_add<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a, d_b, n);
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

int ccd_resize_nearest(const float *d_in, int c, int h, int w, float *d_out, int H, int W, cudaStream_t st) {
    k_resize_nearest<<<grid2(W, H, c), kBlock2, 0, st>>>(d_in, h, w, d_out, H, W, (float)h / (float)H,
                                                         (float)w / (float)W);
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

int ccd_resize_torch(const float *d_in, int c, int h, int w, float *d_out, int H, int W, int mode, float sy, float sx,
                     cudaStream_t st) {
    if (mode == 1)
        k_resize_torch<1><<<grid2(W, H, c), kBlock2, 0, st>>>(d_in, h, w, d_out, H, W, sy, sx);
    else
        k_resize_torch<2><<<grid2(W, H, c), kBlock2, 0, st>>>(d_in, h, w, d_out, H, W, sy, sx);
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

int ccd_cr_noise(float *d_out, size_t first, size_t n, cudaStream_t st) {
    k_cr_noise<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_out, first, n);
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

int ccd_finish(const float *d_in, int h, int w, int bitdepth, int data_type, float *a, float *b, float *c,
               cudaStream_t st) {
    const float M = (float)((1 << bitdepth) - 1);
    const size_t n = (size_t)h * w;
    if (data_type != 1) {
        k_finish_444<<<(unsigned)((3 * n + 255) / 256), 256, 0, st>>>(d_in, 3 * n, M, a);
    } else {
        k_finish_444<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, n, M, a);
        if (h / 2 > 0 && w / 2 > 0) {
            k_finish_420_uv<<<grid2(w / 2, h / 2, 2), kBlock2, 0, st>>>(d_in, h, w, M, b, c);
            g_ccd_launches++;
        }
    }
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

// ccd_entropy.cu -- entropy stage of the Cool-chic decoder on sm_100a.
//
// Replaces, for one Cool-chic stream, the reference's hot loops A/B/C
//   component/coolchic.py:89-166   per-grid loop, IFCE context (coarse -> fine)
//   latent.py:142-173              wavefront loop: gather context, ARM, range decode, scatter
//   armint.py:180-203              int64 fixed-point ARM MLP
//   rangecoder.py:87-94            (mu, scale) table lookup + constriction RangeDecoder.decode
// with ONE persistent CTA per stream (a stream is a strict serial chain):
//
//   * 15 producer warps, one thread per symbol: wait until the symbol's left neighbour is
//     decoded (shared-memory progress counter), gather the causal neighbourhood from a
//     shared-memory row ring, evaluate IFCE + ARM in integer arithmetic (IMAD.WIDE, int32
//     operands proven safe by the host, int64 accumulators), then fetch the symbol's
//     32-entry cumulative window from the device-resident quantised-Laplace table and
//     publish it in a shared-memory ring.
//   * 1 range-coder warp: each lane owns one candidate symbol of the window; the lane whose
//     [scale*left, scale*left') interval contains (point - lower) wins (no division), the
//     new state is broadcast with shuffles.  ~1 ballot + 4 shuffles per symbol.
//
// The same kernel runs in "encode" / "sample" mode (range ENcoder, rangecoder.py:46-78) to
// fabricate self-consistent synthetic streams on the device.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ccd_internal.h"

namespace {

// context pattern: core/arm.py:496-562 (priority order over the 9x9 causal mask)
__constant__ int8_t c_ctx_dy[40] = {0,  -1, -1, -1, 0,  -2, -3, 0,  -1, -2, -2, -1, -2, -1,
                                    -2, -3, 0,  -1, -2, -2, -3, -3, -3, -4, -1, -4, -1, -2,
                                    -3, -3, -4, -4, -2, -3, -3, -4, -4, -4, -4, -4};
__constant__ int8_t c_ctx_dx[40] = {-1, 0,  -1, 1, -2, 0,  0,  -3, -2, 1,  -1, 2,  -2, -3,
                                    2,  1,  -4, 3, -3, 3,  -1, -2, 2,  0,  -4, -1, 4,  4,
                                    -3, 3,  -2, 1, -4, -4, 4,  -3, 2,  3,  -4, 4};

constexpr double kFreeWeight = 16777215.0 - 127.0;  // (2^24 - 1) - (max - min)
constexpr int kSymMin = -64, kSymMax = 63;

// trunc(FW * cdf(d / b)) with cdf the Laplace CDF of constriction's QuantizedLaplace
// (SURVEY Appendix C.1); d = (s - 0.5) - mu is exact in f64.
__device__ __forceinline__ uint32_t laplace_nonleaky(double d, double b) {
    double c;
    if (d <= 0.0)
        c = __dmul_rn(0.5, exp(__ddiv_rn(d, b)));
    else
        c = __dsub_rn(1.0, __dmul_rn(0.5, exp(__ddiv_rn(-d, b))));
    return (uint32_t)__double2ll_rz(__dmul_rn(kFreeWeight, c));
}

__device__ __forceinline__ uint32_t laplace_left_exact(int s, double mu, double b) {
    if (s <= kSymMin) return 0u;
    if (s > kSymMax) return 1u << 24;
    return laplace_nonleaky(((double)s - 0.5) - mu, b) + (uint32_t)(s - kSymMin);
}

// Table: NL[sc][f][t] = trunc(FW*cdf((t - 15.5) - (f-128)/256)), sc < 2561, f < 256, t < 32.
__global__ void k_cdf_table(uint32_t *__restrict__ tab, const float *__restrict__ scale) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)CCD_N_SCALE * 256 * CCD_WIN;
    if (idx >= total) return;
    int t = (int)(idx & 31);
    int f = (int)((idx >> 5) & 255);
    int sc = (int)(idx >> 13);
    double b = (double)scale[sc];
    double d = ((double)t - 15.5) - (double)(f - 128) * (1.0 / 256.0);
    tab[idx] = laplace_nonleaky(d, b);
}

__global__ void k_laplace_domain(const float *__restrict__ scale, int sc_lo, int sc_hi, uint32_t *lo,
                                 uint32_t *hi) {
    const int ND = 32641;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)(sc_hi - sc_lo) * ND;
    if (idx >= total) return;
    int n = (int)(idx % ND);
    int sc = sc_lo + (int)(idx / ND);
    double b = (double)scale[sc];
    double d = (double)n * (1.0 / 256.0);
    lo[idx] = laplace_nonleaky(-d, b);
    hi[idx] = (n == 0) ? laplace_nonleaky(0.0, b) : laplace_nonleaky(d, b);
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    return *reinterpret_cast<const volatile uint32_t *>(p);
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
    *reinterpret_cast<volatile uint32_t *>(p) = v;
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// per-stream scalars, copied to registers once (EntStream lives in global memory)
struct SLoc {
    int ring, rows, n_hidden, n_ctx, cf, mode;
    int8_t *latents;
    const uint32_t *words;
    int64_t n_words;
    uint32_t *out_words;
    int64_t out_cap;
};

__device__ __forceinline__ void st_shared_v4(uint4 *p, uint4 v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"((uint32_t)__cvta_generic_to_shared(p)),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p))
                 : "memory");
    return v;
}

// shared memory carve-up -----------------------------------------------------------------
struct SmemLayout {
    uint32_t *ctrl;        // [0] progress (symbols decoded), [1] abort flag
    EntGrid *grid;         // current grid
    unsigned char *arm;    // ARM blob
    unsigned char *ifce;   // IFCE blob of the current grid
    uint4 *meta;           // [ring]
    uint32_t *win;         // [ring][32], 16B chunks XOR-swizzled by (slot & 7)
    int8_t *rows;          // [rows][64]
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

__device__ __forceinline__ SmemLayout carve(unsigned char *base, int ring, int rows, int arm_bytes,
                                            int ifce_bytes) {
    SmemLayout L;
    size_t p = 0;
    L.ctrl = reinterpret_cast<uint32_t *>(base + p);
    p += 64;
    L.grid = reinterpret_cast<EntGrid *>(base + p);
    p += align16(sizeof(EntGrid));
    L.arm = base + p;
    p += align16((size_t)arm_bytes);
    L.ifce = base + p;
    p += align16((size_t)ifce_bytes);
    L.meta = reinterpret_cast<uint4 *>(base + p);
    p += (size_t)ring * 16;
    L.win = reinterpret_cast<uint32_t *>(base + p);
    p += (size_t)ring * CCD_WIN * 4;
    L.rows = reinterpret_cast<int8_t *>(base + p);
    (void)rows;
    return L;
}

// ---------------------------------------------------------------------------------------
// ARM / IFCE evaluation (armint.py:180-203).  FAST: int32 operands, IMAD.WIDE.
template <int NCTX, int CF>
struct FastArm {
    static constexpr int DIM = NCTX + CF;
    static constexpr int DIMP = (DIM + 3) & ~3;
    static constexpr int CFP = (CF + 3) & ~3;

    // IFCE features (component/coolchic.py:105-146) for the pixel (yy, xx) of the previous grid
    static __device__ __forceinline__ void ifce(const EntGrid *g, const unsigned char *blob,
                                                const int8_t *lat, int yy, int xx, int32_t *xf) {
        if constexpr (CF > 0) {
            const int n_in = g->ifce_in;
            if (n_in == 0) {
#pragma unroll
                for (int f = 0; f < CF; f++) xf[f] = 0;
                return;
            }
            const int32_t *W = reinterpret_cast<const int32_t *>(blob);
            const long long *B =
                reinterpret_cast<const long long *>(blob + (((size_t)n_in * CFP * 4 + 7) & ~(size_t)7));
            long long acc[CF];
#pragma unroll
            for (int f = 0; f < CF; f++) acc[f] = B[f];
            for (int c = 0; c < n_in; c++) {
                int sh = g->ch_sh[c];
                int v = 0;
                if (sh >= 0) v = lat[g->ch_off[c] + (long long)(yy >> sh) * g->ch_w[c] + (xx >> sh)];
                int xi = v << 16;
#pragma unroll
                for (int f4 = 0; f4 < CFP / 4; f4++) {
                    int4 w = *reinterpret_cast<const int4 *>(W + c * CFP + f4 * 4);
                    if (f4 * 4 + 0 < CF) acc[f4 * 4 + 0] += (long long)w.x * xi;
                    if (f4 * 4 + 1 < CF) acc[f4 * 4 + 1] += (long long)w.y * xi;
                    if (f4 * 4 + 2 < CF) acc[f4 * 4 + 2] += (long long)w.z * xi;
                    if (f4 * 4 + 3 < CF) acc[f4 * 4 + 3] += (long long)w.w * xi;
                }
            }
#pragma unroll
            for (int f = 0; f < CF; f++) {
                long long o = acc[f] >> 24;
                // F.interpolate(ctx.to(torch.float)).to(int64): fp32 round trip (coolchic.py:142-144)
                float fl = __ll2float_rn(o);
                xf[f] = (int32_t)__float2ll_rz(fl);
            }
        }
    }

    // x: DIM context integers (latents, IFCE features with 8 fractional bits)
    static __device__ __forceinline__ void arm(const unsigned char *blob, int n_hidden, int32_t (&x)[DIM],
                                               long long &o0, long long &o1) {
        const int32_t *Wh = reinterpret_cast<const int32_t *>(blob);
        const int32_t *Wl = Wh + (size_t)n_hidden * DIM * DIMP;
        const int32_t *Ws = Wl + DIM * 2;
        size_t wbytes = ((size_t)(n_hidden * DIM * DIMP + DIM * 4) * 4 + 7) & ~(size_t)7;
        const long long *Bh = reinterpret_cast<const long long *>(blob + wbytes);
        const long long *Bl = Bh + (size_t)n_hidden * DIM;
        const long long *Bs = Bl + 2;
#pragma unroll
        for (int i = 0; i < DIM; i++) x[i] <<= 16;
        long long s0 = Bs[0], s1 = Bs[1];
#pragma unroll
        for (int i = 0; i < DIM; i++) {
            int2 w = *reinterpret_cast<const int2 *>(Ws + 2 * i);
            s0 += (long long)w.x * x[i];
            s1 += (long long)w.y * x[i];
        }
        for (int l = 0; l < n_hidden; l++) {
            long long acc[DIMP];
            const long long *B = Bh + (size_t)l * DIM;
            const int32_t *W = Wh + (size_t)l * DIM * DIMP;
#pragma unroll
            for (int o = 0; o < DIMP; o++) acc[o] = (o < DIM) ? B[o] : 0;
#pragma unroll
            for (int i = 0; i < DIM; i++) {
                const int xi = x[i];
#pragma unroll
                for (int o4 = 0; o4 < DIMP / 4; o4++) {
                    int4 w = *reinterpret_cast<const int4 *>(W + i * DIMP + o4 * 4);
                    acc[o4 * 4 + 0] += (long long)w.x * xi;
                    acc[o4 * 4 + 1] += (long long)w.y * xi;
                    acc[o4 * 4 + 2] += (long long)w.z * xi;
                    acc[o4 * 4 + 3] += (long long)w.w * xi;
                }
            }
#pragma unroll
            for (int o = 0; o < DIM; o++) {
                long long a = acc[o];
                a = a < 0 ? 0 : a;
                x[o] = (int32_t)(a >> 16);
            }
        }
        long long a0 = Bl[0], a1 = Bl[1];
#pragma unroll
        for (int i = 0; i < DIM; i++) {
            int2 w = *reinterpret_cast<const int2 *>(Wl + 2 * i);
            a0 += (long long)w.x * x[i];
            a1 += (long long)w.y * x[i];
        }
        o0 = (a0 + s0) >> 24;
        o1 = (a1 + s1) >> 24;
    }
};

// GENERIC: everything int64, runtime sizes, arrays in local memory.  Always correct, slow.
struct GenericArm {
    static __device__ void ifce(const EntGrid *g, const unsigned char *blob, const int8_t *lat, int cf,
                                int yy, int xx, long long *xf) {
        const int n_in = g->ifce_in;
        if (n_in == 0) {
            for (int f = 0; f < cf; f++) xf[f] = 0;
            return;
        }
        const long long *W = reinterpret_cast<const long long *>(blob);
        const long long *B = W + (size_t)n_in * cf;
        for (int f = 0; f < cf; f++) xf[f] = B[f];
        for (int c = 0; c < n_in; c++) {
            int sh = g->ch_sh[c];
            long long v = 0;
            if (sh >= 0) v = lat[g->ch_off[c] + (long long)(yy >> sh) * g->ch_w[c] + (xx >> sh)];
            long long xi = v << 16;
            for (int f = 0; f < cf; f++) xf[f] += W[(size_t)c * cf + f] * xi;
        }
        for (int f = 0; f < cf; f++) {
            long long o = xf[f] >> 24;
            float fl = __ll2float_rn(o);
            xf[f] = __float2ll_rz(fl);
        }
    }
    static __device__ void arm(const unsigned char *blob, int dim, int n_hidden, long long *x, long long &o0,
                               long long &o1) {
        const long long *Wh = reinterpret_cast<const long long *>(blob);
        const long long *Wl = Wh + (size_t)n_hidden * dim * dim;
        const long long *Ws = Wl + (size_t)dim * 2;
        const long long *Bh = Ws + (size_t)dim * 2;
        const long long *Bl = Bh + (size_t)n_hidden * dim;
        const long long *Bs = Bl + 2;
        long long y[CCD_MAX_DIM];
        for (int i = 0; i < dim; i++) x[i] <<= 16;
        long long s0 = Bs[0], s1 = Bs[1];
        for (int i = 0; i < dim; i++) {
            s0 += Ws[2 * i] * x[i];
            s1 += Ws[2 * i + 1] * x[i];
        }
        for (int l = 0; l < n_hidden; l++) {
            const long long *W = Wh + (size_t)l * dim * dim;
            for (int o = 0; o < dim; o++) y[o] = Bh[(size_t)l * dim + o];
            for (int i = 0; i < dim; i++) {
                long long xi = x[i];
                for (int o = 0; o < dim; o++) y[o] += W[(size_t)i * dim + o] * xi;
            }
            for (int o = 0; o < dim; o++) {
                long long a = y[o] < 0 ? 0 : y[o];
                x[o] = a >> 16;
            }
        }
        long long a0 = Bl[0], a1 = Bl[1];
        for (int i = 0; i < dim; i++) {
            a0 += Wl[2 * i] * x[i];
            a1 += Wl[2 * i + 1] * x[i];
        }
        o0 = (a0 + s0) >> 24;
        o1 = (a1 + s1) >> 24;
    }
};

// ---------------------------------------------------------------------------------------
// Producer: symbols [c0, c0+32) of diagonal k.
template <int NCTX, int CF, bool FAST>
__device__ __forceinline__ void produce_chunk(const SLoc &S, const SmemLayout &sm,
                                              const uint32_t *__restrict__ cdf, int lane, int y0, int x0,
                                              int n_k, int c0, uint32_t ord_diag, uint32_t ord_prev,
                                              int y0_prev) {
    const EntGrid *g = sm.grid;
    const int i = c0 + lane;
    const bool valid = i < n_k;
    const int w = g->w;
    const int y = y0 + i;
    const int x = g->raster ? x0 : x0 - CCD_MASK_STRIDE * i;
    const uint32_t ord = ord_diag + (uint32_t)i;
    const uint32_t ring_mask = (uint32_t)S.ring - 1u;
    const uint32_t row_mask = (uint32_t)S.rows - 1u;

    // ---- dependencies: left neighbour decoded (and everything older), ring slot free
    uint32_t need;
    if (g->raster) need = ord;                                  // everything before me
    else if (x > 0) need = ord_prev + (uint32_t)(y - y0_prev) + 1u; // (y, x-1) sits on diagonal k-1
    else need = ord_prev;                                        // first pixel of a row
    uint32_t need_ring = ord + 1u - (uint32_t)S.ring;            // slot reuse: ord - ring consumed
    if ((int32_t)(need_ring - need) > 0) need = need_ring;
    // warp-wide maximum (needs are increasing with the lane): one wait for the whole warp
    int32_t rel = valid ? (int32_t)(need - ord_diag) : INT32_MIN;
    rel = __reduce_max_sync(0xffffffffu, rel);
    need = ord_diag + (uint32_t)rel;
    while ((int32_t)(ld_volatile_u32(&sm.ctrl[0]) - need) < 0) __nanosleep(32);
    if (!valid) return;

    // ---- gather the causal neighbourhood (latent.py:148-153), zero outside the grid
    const int n_ctx = FAST ? NCTX : S.n_ctx;
    const int cf = FAST ? CF : S.cf;
    long long o0, o1;
    if constexpr (FAST) {
        int32_t xin[NCTX + CF];
#pragma unroll
        for (int t = 0; t < NCTX; t++) {
            const int yy = y + c_ctx_dy[t], xx = x + c_ctx_dx[t];
            int v = 0;
            if (yy >= 0 && xx >= 0 && xx < w)
                v = *reinterpret_cast<const volatile int8_t *>(
                    &sm.rows[(((uint32_t)yy & row_mask) << 6) | ((uint32_t)xx & (CCD_ROW_COLS - 1))]);
            xin[t] = v;
        }
        if constexpr (CF > 0) FastArm<NCTX, CF>::ifce(g, sm.ifce, S.latents, y >> 1, x >> 1, xin + NCTX);
        FastArm<NCTX, CF>::arm(sm.arm, S.n_hidden, xin, o0, o1);
    } else {
        long long xin[CCD_MAX_DIM];
        for (int t = 0; t < n_ctx; t++) {
            const int yy = y + c_ctx_dy[t], xx = x + c_ctx_dx[t];
            int v = 0;
            if (yy >= 0 && xx >= 0 && xx < w)
                v = *reinterpret_cast<const volatile int8_t *>(
                    &sm.rows[(((uint32_t)yy & row_mask) << 6) | ((uint32_t)xx & (CCD_ROW_COLS - 1))]);
            xin[t] = v;
        }
        if (cf > 0) GenericArm::ifce(g, sm.ifce, S.latents, cf, y >> 1, x >> 1, xin + n_ctx);
        GenericArm::arm(sm.arm, n_ctx + cf, S.n_hidden, xin, o0, o1);
    }
    // latent.py:165 + rangecoder.py:89-91 (np.take(..., mode="clip"))
    long long im = o0 + 16384, is = o1 + 1280;
    im = im < 0 ? 0 : (im > 32767 ? 32767 : im);
    is = is < 0 ? 0 : (is > 2560 ? 2560 : is);
    const int mu_idx = (int)im, sc_idx = (int)is;
    const int mu_int = ((mu_idx + 128) >> 8) - 64;
    const int fr = (mu_idx + 128) & 255;
    const int s_lo = mu_int - CCD_WIN_HALF;

    // ---- cumulative window: table row -> leak term + clamps -> shared ring
    const uint4 *row = reinterpret_cast<const uint4 *>(cdf + (((size_t)sc_idx << 8 | (size_t)fr) << 5));
    uint4 v[8];
#pragma unroll
    for (int c = 0; c < 8; c++) v[c] = __ldg(row + c);
    const uint32_t slot = ord & ring_mask;
    uint4 *wdst = reinterpret_cast<uint4 *>(sm.win + (size_t)slot * CCD_WIN);
#pragma unroll
    for (int c = 0; c < 8; c++) {
        uint32_t e[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int s = s_lo + c * 4 + q;
            uint32_t l = e[q] + (uint32_t)(s - kSymMin);
            l = (s <= kSymMin) ? 0u : l;
            l = (s > kSymMax) ? (1u << 24) : l;
            e[q] = l;
        }
        wdst[c ^ (slot & 7u)] = make_uint4(e[0], e[1], e[2], e[3]);
    }
    __threadfence_block();
    const uint32_t out_off = (uint32_t)(g->lat_off + (long long)y * w + x);
    const uint32_t row_idx = (((uint32_t)y & row_mask) << 6) | ((uint32_t)x & (CCD_ROW_COLS - 1));
    uint4 m = make_uint4(out_off, row_idx, (uint32_t)mu_idx | ((uint32_t)sc_idx << 16), ord + 1u);
    st_shared_v4(&sm.meta[slot], m);
}

template <int NCTX, int CF, bool FAST>
__device__ void producer_grid(const SLoc &S, const SmemLayout &sm, const uint32_t *__restrict__ cdf,
                              int pwarp, int lane, uint32_t ord_grid, int &chunk_ctr) {
    const EntGrid *g = sm.grid;
    const int h = g->h, w = g->w, n_diag = g->n_diag, raster = g->raster;
    uint32_t ord = ord_grid, ord_prev = ord_grid;
    int y0_prev = 0;
    for (int k = 0; k < n_diag; k++) {
        int y0, x0, n_k;
        if (raster) {
            y0 = k / w;
            x0 = k - y0 * w;
            n_k = 1;
        } else if (k < w) {
            y0 = 0;
            x0 = k;
            n_k = min(h, x0 / CCD_MASK_STRIDE + 1);
        } else {
            const int r = k - w;
            y0 = r / CCD_MASK_STRIDE + 1;
            x0 = w - CCD_MASK_STRIDE + (r - (y0 - 1) * CCD_MASK_STRIDE);
            n_k = min(h - y0, x0 / CCD_MASK_STRIDE + 1);
        }
        for (int c0 = 0; c0 < n_k; c0 += 32) {
            if (chunk_ctr == pwarp)
                produce_chunk<NCTX, CF, FAST>(S, sm, cdf, lane, y0, x0, n_k, c0, ord, ord_prev, y0_prev);
            chunk_ctr = (chunk_ctr + 1 == CCD_ENT_PRODUCERS) ? 0 : chunk_ctr + 1;
        }
        ord_prev = ord;
        y0_prev = y0;
        ord += (uint32_t)n_k;
    }
}

// ---------------------------------------------------------------------------------------
// Range-coder warp.  State: D = point - lower (mod 2^64), R = range.
struct Coder {
    uint64_t D, R;       // decoder: D = point - lower ; encoder: D = lower
    int64_t wpos;        // next word index
    uint32_t wcur, wnxt; // lane l holds word (chunk*32 + l) of the current / next 32-word chunk
    uint32_t wnext;      // word[wpos], broadcast
    uint64_t prng;
    int64_t nout;        // encoder: words emitted
    uint32_t slow;       // slow-path count
    int err;
};

__device__ __forceinline__ uint32_t load_word(const SLoc &S, int64_t i) {
    return (i < S.n_words) ? __ldg(S.words + i) : 0u;
}

__device__ __forceinline__ void coder_advance_word(const SLoc &S, Coder &c, int lane) {
    c.wpos++;
    if ((c.wpos & 31) == 0) {
        c.wcur = c.wnxt;
        c.wnxt = load_word(S, c.wpos + 32 + lane);
    }
    c.wnext = __shfl_sync(0xffffffffu, c.wcur, (int)(c.wpos & 31));
}

__device__ void encoder_emit(const SLoc &S, Coder &c, uint32_t L0, uint32_t L1, int lane) {
    const uint64_t scale = c.R >> 24;
    const uint64_t nl = c.D + scale * L0;
    const bool carry = nl < c.D;
    c.D = nl;
    c.R = scale * (uint64_t)(L1 - L0);
    if (lane == 0 && carry) {
        int64_t i = c.nout;
        while (i > 0) {
            i--;
            if (i < S.out_cap) {
                uint32_t v = S.out_words[i] + 1u;
                S.out_words[i] = v;
                if (v != 0u) break;
            }
        }
    }
    if (c.R < (1ull << 32)) {
        if (lane == 0 && c.nout < S.out_cap) S.out_words[c.nout] = (uint32_t)(c.D >> 32);
        c.nout++;
        c.D <<= 32;
        c.R <<= 32;
    }
}

__device__ void coder_grid(const SLoc &S, const SmemLayout &sm, const float *__restrict__ scale_tab,
                           int lane, uint32_t ord_begin, uint32_t ord_end, Coder &c) {
    const uint32_t ring_mask = (uint32_t)S.ring - 1u;
    const int mode = S.mode;
    for (uint32_t j = ord_begin; j != ord_end; j++) {
        const uint32_t slot = j & ring_mask;
        uint4 m;
        do {
            m = ld_shared_v4(&sm.meta[slot]);
        } while (m.w != j + 1u);
        const uint32_t *wrow = sm.win + (size_t)slot * CCD_WIN;
        const uint32_t sw = slot & 7u;
        const int e1 = lane < 31 ? lane + 1 : lane;
        const uint32_t L0 = ld_volatile_u32(wrow + ((((uint32_t)lane >> 2) ^ sw) << 2) + (lane & 3));
        const uint32_t L1 = ld_volatile_u32(wrow + ((((uint32_t)e1 >> 2) ^ sw) << 2) + (e1 & 3));
        const int mu_idx = (int)(m.z & 0xffffu), sc_idx = (int)(m.z >> 16);
        const int s_lo = (((mu_idx + 128) >> 8) - 64) - CCD_WIN_HALF;
        int sym;
        if (mode == 0) {
            // ------------------------------------------------------------ decode
            const uint64_t scale = c.R >> 24;
            uint64_t P0 = scale * L0, P1 = scale * L1;
            bool win = (P0 <= c.D) && (c.D < P1);
            uint32_t ballot = __ballot_sync(0xffffffffu, win);
            int src;
            if (ballot == 0u) {
                // slow path: symbol outside the 31-symbol window (or corrupt stream)
                c.slow++;
                uint64_t q = c.D / scale;
                if (q >= (1ull << 24)) {
                    c.err = CCD_ERR_DESYNC;
                    q = (1ull << 24) - 1;
                }
                const double mu = (double)(mu_idx - 16384) * (1.0 / 256.0);
                const double b = (double)scale_tab[sc_idx];
                uint32_t l0 = 0, l1 = 0;
                int s = 0;
                for (int r = 0; r < 4; r++) {
                    s = kSymMin + r * 32 + lane;
                    l0 = laplace_left_exact(s, mu, b);
                    l1 = laplace_left_exact(s + 1, mu, b);
                    ballot = __ballot_sync(0xffffffffu, l0 <= (uint32_t)q && (uint32_t)q < l1);
                    if (ballot) break;
                }
                if (ballot == 0u) {  // cannot happen: left(-64)=0, left(64)=2^24 > q
                    c.err = CCD_ERR_DESYNC;
                    ballot = 1u;
                }
                src = __ffs(ballot) - 1;
                sym = __shfl_sync(0xffffffffu, s, src);
                P0 = scale * l0;
                P1 = scale * l1;
            } else {
                src = __ffs(ballot) - 1;
                sym = s_lo + src;
            }
            uint64_t Dn = c.D - P0, Rn = P1 - P0;
            const bool rn = (Rn >> 32) == 0;
            if (rn) {
                Rn <<= 32;
                Dn = (Dn << 32) | c.wnext;
            }
            const uint32_t rmask = __ballot_sync(0xffffffffu, rn);
            c.D = shfl_u64(Dn, src);
            c.R = shfl_u64(Rn, src);
            if ((rmask >> src) & 1u) coder_advance_word(S, c, lane);
        } else {
            // ------------------------------------------------------------ encode / sample
            uint32_t l0, l1;
            int src = -1;
            if (mode == 2) {
                const uint32_t q = (uint32_t)(splitmix64(c.prng) >> 40);
                uint32_t ballot = __ballot_sync(0xffffffffu, L0 <= q && q < L1);
                if (ballot) {
                    src = __ffs(ballot) - 1;
                    sym = s_lo + src;
                } else {
                    const double mu = (double)(mu_idx - 16384) * (1.0 / 256.0);
                    const double b = (double)scale_tab[sc_idx];
                    int s = 0;
                    for (int r = 0; r < 4; r++) {
                        s = kSymMin + r * 32 + lane;
                        ballot = __ballot_sync(0xffffffffu, laplace_left_exact(s, mu, b) <= q &&
                                                                q < laplace_left_exact(s + 1, mu, b));
                        if (ballot) break;
                    }
                    sym = __shfl_sync(0xffffffffu, s, __ffs(ballot) - 1);
                }
            } else {
                sym = S.latents[m.x];
                if (sym >= s_lo && sym < s_lo + 31) src = sym - s_lo;
            }
            if (src >= 0) {
                l0 = __shfl_sync(0xffffffffu, L0, src);
                l1 = __shfl_sync(0xffffffffu, L1, src);
            } else {
                const double mu = (double)(mu_idx - 16384) * (1.0 / 256.0);
                const double b = (double)scale_tab[sc_idx];
                l0 = laplace_left_exact(sym, mu, b);
                l1 = laplace_left_exact(sym + 1, mu, b);
                c.slow++;
            }
            encoder_emit(S, c, l0, l1, lane);
        }
        if (lane == 0) {
            *reinterpret_cast<volatile int8_t *>(&sm.rows[m.y]) = (int8_t)sym;
            S.latents[m.x] = (int8_t)sym;
            st_volatile_u32(&sm.ctrl[0], j + 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------
template <int NCTX, int CF, bool FAST>
__global__ void __launch_bounds__(CCD_ENT_THREADS, 1)
    k_entropy(const EntStream *__restrict__ streams, const uint32_t *__restrict__ cdf,
              const float *__restrict__ scale_tab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EntStream &G = streams[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SmemLayout sm = carve(smem_raw, G.ring, G.rows, G.arm_blob_bytes, G.ifce_blob_max);
    SLoc S;
    S.ring = G.ring;
    S.rows = G.rows;
    S.n_hidden = G.n_hidden;
    S.n_ctx = G.n_ctx;
    S.cf = G.cf;
    S.mode = G.mode;
    S.latents = G.latents;
    S.words = G.words;
    S.n_words = G.n_words;
    S.out_words = G.out_words;
    S.out_cap = G.out_cap;

    // one-time: control words, meta tags, ARM parameters
    if (tid < 16) sm.ctrl[tid] = 0u;
    for (int i = tid; i < S.ring; i += CCD_ENT_THREADS) sm.meta[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid * 4; i < G.arm_blob_bytes; i += CCD_ENT_THREADS * 4)
        *reinterpret_cast<uint32_t *>(sm.arm + i) = *reinterpret_cast<const uint32_t *>(G.blob + i);

    const bool is_coder = (warp == CCD_ENT_WARPS - 1);
    Coder cd;
    cd.err = 0;
    cd.slow = 0;
    cd.nout = 0;
    cd.prng = G.seed;
    cd.wpos = 2;
    cd.wcur = cd.wnxt = cd.wnext = 0;
    if (is_coder) {
        if (S.mode == 0) {
            cd.wcur = load_word(S, lane);
            cd.wnxt = load_word(S, 32 + lane);
            const uint32_t w0 = __shfl_sync(0xffffffffu, cd.wcur, 0);
            const uint32_t w1 = __shfl_sync(0xffffffffu, cd.wcur, 1);
            cd.D = ((uint64_t)w0 << 32) | w1;
            cd.wnext = __shfl_sync(0xffffffffu, cd.wcur, 2);
        } else {
            cd.D = 0;
        }
        cd.R = ~0ull;
    }
    uint32_t ord = 0;
    int chunk_ctr = 0;
    const int n_grids = G.n_grids;
    for (int gi = 0; gi < n_grids; gi++) {
        __syncthreads();  // previous grid fully decoded, its latents visible CTA-wide
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&G.grid[gi]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(sm.grid);
            for (int i = tid; i < (int)(sizeof(EntGrid) / 4); i += CCD_ENT_THREADS) dst[i] = src[i];
            const EntGrid &Gg = G.grid[gi];
            for (int i = tid * 4; i < Gg.ifce_blob_bytes; i += CCD_ENT_THREADS * 4)
                *reinterpret_cast<uint32_t *>(sm.ifce + i) =
                    *reinterpret_cast<const uint32_t *>(G.blob + Gg.ifce_blob_off + i);
        }
        __syncthreads();
        const uint32_t n_sym = (uint32_t)sm.grid->h * (uint32_t)sm.grid->w;
        if (is_coder)
            coder_grid(S, sm, scale_tab, lane, ord, ord + n_sym, cd);
        else
            producer_grid<NCTX, CF, FAST>(S, sm, cdf, warp, lane, ord, chunk_ctr);
        ord += n_sym;
    }
    if (is_coder) {
        if (S.mode != 0 && ord > 0) {
            // seal: point = lower + 2^32 - 1, emit its high word
            const uint64_t point = cd.D + ((1ull << 32) - 1);
            if (lane == 0) {
                if (point < cd.D) {
                    int64_t i = cd.nout;
                    while (i > 0) {
                        i--;
                        if (i < S.out_cap) {
                            uint32_t v = S.out_words[i] + 1u;
                            S.out_words[i] = v;
                            if (v != 0u) break;
                        }
                    }
                }
                if (cd.nout < S.out_cap) S.out_words[cd.nout] = (uint32_t)(point >> 32);
            }
            cd.nout++;
        }
        if (lane == 0) {
            G.status[0] = cd.err;
            G.status[1] = (int32_t)cd.wpos;
            G.status[2] = (int32_t)cd.slow;
            G.status[3] = (int32_t)cd.nout;
        }
    }
}

template <int NCTX, int CF, bool FAST>
int launch_t(const EntStream *d_streams, int n, size_t smem, const uint32_t *cdf, const float *scale,
             cudaStream_t st) {
    auto kern = k_entropy<NCTX, CF, FAST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<n, CCD_ENT_THREADS, smem, st>>>(d_streams, cdf, scale);
    return (int)cudaGetLastError();
}

}  // namespace

size_t ccd_entropy_smem_bytes(int ring, int rows, int arm_blob_bytes, int ifce_blob_max) {
    size_t p = 64 + align16(sizeof(EntGrid)) + align16((size_t)arm_blob_bytes) + align16((size_t)ifce_blob_max);
    p += (size_t)ring * 16 + (size_t)ring * CCD_WIN * 4 + (size_t)rows * CCD_ROW_COLS;
    return p;
}

bool ccd_entropy_has_fast(int n_ctx, int cf) {
    return (n_ctx == 6 && cf == 2) || (n_ctx == 10 && cf == 2) || (n_ctx == 10 && cf == 4) ||
           (n_ctx == 14 && cf == 6) || (n_ctx == 20 && cf == 6);
}

int ccd_entropy_launch(const EntStream *d_streams, int n_streams, const EntLaunchCfg &cfg,
                       const uint32_t *d_cdf, const float *d_scale, cudaStream_t st) {
    if (cfg.fast) {
        if (cfg.n_ctx == 6 && cfg.cf == 2) return launch_t<6, 2, true>(d_streams, n_streams, cfg.smem_bytes, d_cdf, d_scale, st);
        if (cfg.n_ctx == 10 && cfg.cf == 2) return launch_t<10, 2, true>(d_streams, n_streams, cfg.smem_bytes, d_cdf, d_scale, st);
        if (cfg.n_ctx == 10 && cfg.cf == 4) return launch_t<10, 4, true>(d_streams, n_streams, cfg.smem_bytes, d_cdf, d_scale, st);
        if (cfg.n_ctx == 14 && cfg.cf == 6) return launch_t<14, 6, true>(d_streams, n_streams, cfg.smem_bytes, d_cdf, d_scale, st);
        if (cfg.n_ctx == 20 && cfg.cf == 6) return launch_t<20, 6, true>(d_streams, n_streams, cfg.smem_bytes, d_cdf, d_scale, st);
    }
    return launch_t<0, 0, false>(d_streams, n_streams, cfg.smem_bytes, d_cdf, d_scale, st);
}

int ccd_cdf_table_build(uint32_t *d_cdf, const float *d_scale, cudaStream_t st) {
    size_t total = (size_t)CCD_N_SCALE * 256 * CCD_WIN;
    k_cdf_table<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_cdf, d_scale);
    return (int)cudaGetLastError();
}

int ccd_laplace_domain(const float *d_scale, int sc_lo, int sc_hi, uint32_t *d_lo, uint32_t *d_hi,
                       cudaStream_t st) {
    size_t total = (size_t)(sc_hi - sc_lo) * 32641;
    k_laplace_domain<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_scale, sc_lo, sc_hi, d_lo, d_hi);
    return (int)cudaGetLastError();
}

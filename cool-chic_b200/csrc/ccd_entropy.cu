// ccd_entropy.cu -- entropy stage of the Cool-chic decoder on sm_100a.
//
// Replaces, for one Cool-chic stream, the reference's hot loops A/B/C
//   component/coolchic.py:89-166   per-grid loop, IFCE context (coarse -> fine)
//   latent.py:142-173              wavefront loop: gather context, ARM, range decode, scatter
//   armint.py:180-203              int64 fixed-point ARM MLP
//   rangecoder.py:87-94            (mu, scale) table lookup + constriction RangeDecoder.decode
// with ONE persistent CTA per stream (a stream is a strict serial chain):
//
//   * producer warps, a QUAD of lanes per symbol (8 symbols per warp): wait until the
//     symbol's left neighbour is decoded (shared-memory progress counter), gather the causal
//     neighbourhood from a shared-memory row ring, evaluate IFCE + ARM in integer arithmetic
//     (IMAD.WIDE, int32 operands proven safe by the host, int64 accumulators, activations
//     exchanged with quad shuffles), then fetch the symbol's 32-entry cumulative window from
//     the device-resident quantised-Laplace table and publish it in a shared-memory ring.
//   * 1 range-coder warp: each lane owns one candidate symbol of the window; the lane whose
//     [scale*left, scale*left') interval contains (point - lower) wins (no division), the
//     new state is broadcast with shuffles.  ~1 ballot + 4 shuffles per symbol.
//
// The same kernel runs in "encode" / "sample" mode (range ENcoder, rangecoder.py:46-78) to
// fabricate self-consistent synthetic streams on the device.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ccd_internal.h"

#ifdef CCD_PROFILE
#define PROF_T(var) long long var = clock64()
#define PROF_ADD(acc, t0) acc += clock64() - (t0)
#else
#define PROF_T(var)
#define PROF_ADD(acc, t0)
#endif

namespace {

struct ProfCounters {
    long long wait = 0, arm = 0, win = 0, total = 0;
    long long seg[6] = {0, 0, 0, 0, 0, 0};  // coder: per-segment cycles of the decode step
};

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

// Table: NL[sc][f][t] = trunc(FW*cdf((t - CCD_WIN_HALF - 0.5) - (f-128)/256)), sc < 2561, f < 256, t < 32.
__global__ void k_cdf_table(uint32_t *__restrict__ tab, const float *__restrict__ scale) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)CCD_N_SCALE * 256 * CCD_WIN;
    if (idx >= total) return;
    int t = (int)(idx & 31);
    int f = (int)((idx >> 5) & 255);
    int sc = (int)(idx >> 13);
    double b = (double)scale[sc];
    double d = ((double)t - ((double)CCD_WIN_HALF + 0.5)) - (double)(f - 128) * (1.0 / 256.0);
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

__device__ __forceinline__ uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, mask);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), mask);
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
    int32_t *status;
    int64_t n_symbols;
};

// Shared-memory accessors by 32-bit shared-space address (explicit LDS/STS, never generic).
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t a, uint2 v) {
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(a)
                 : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int lds_s8(uint32_t a) {
    int v;
    asm volatile("ld.volatile.shared.s8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, int v) {
    asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// shared memory carve-up -----------------------------------------------------------------
struct SmemLayout {
    uint32_t ctrl;         // shared address: [0] progress (symbols decoded)
    EntGrid *grid;         // current grid
    unsigned char *arm;    // ARM blob
    unsigned char *ifce;   // IFCE blob of the current grid
    uint32_t meta;         // shared address: uint4 [ring]
    uint32_t win;          // shared address: u32 [ring][32]: left(s_lo + t), t = 0..31 (mode at t = 15)
    uint32_t hot;          // shared address: uint4 [ring]: left(M), left(M+1)-left(M), left(M-1), left(M+2), M = mode
    uint32_t res;          // shared address: u32 [ring]: result word of each symbol that is NOT the mode (coder -> helper)
    uint32_t rows;         // shared address: int8 [rows][64]
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

__device__ __forceinline__ SmemLayout carve(unsigned char *base, int ring, int rows, int arm_bytes,
                                            int ifce_bytes) {
    SmemLayout L;
    uint32_t base_a = (uint32_t)__cvta_generic_to_shared(base);
    asm volatile("" : "+r"(base_a));  // opaque: keep it in a register instead of re-deriving it
    size_t p = 0;
    L.ctrl = base_a + (uint32_t)p;
    p += 64;
    L.grid = reinterpret_cast<EntGrid *>(base + p);
    p += align16(sizeof(EntGrid));
    L.arm = base + p;
    p += align16((size_t)arm_bytes);
    L.ifce = base + p;
    p += align16((size_t)ifce_bytes);
    p += 16;  // (spare line)
    L.meta = base_a + (uint32_t)p;
    p += (size_t)ring * 16;
    L.win = base_a + (uint32_t)p;
    p += (size_t)ring * CCD_WIN * 4;
    L.hot = base_a + (uint32_t)p;
    p += (size_t)(ring + CCD_HOT_MIRROR) * 16;  // entries 0..7 are mirrored after the end: +32 B never wraps
    L.res = base_a + (uint32_t)p;
    p += (size_t)ring * 4;
    L.rows = base_a + (uint32_t)p;
    (void)rows;
    return L;
}

// ---------------------------------------------------------------------------------------
// ARM / IFCE evaluation (armint.py:180-203).
// FAST path: int32 operands (proven safe by the host), int64 accumulators (IMAD.WIDE), and
// each symbol is spread over a QUAD of lanes: member m owns activations [m*OPM, m*OPM+OPM).
// One warp = 8 symbols.  This cuts the ARM latency ~8x w.r.t. one thread per symbol, which
// is what bounds the short diagonals (the next diagonal cannot start before its ARM is done).
template <int NCTX, int CF>
struct QuadArm {
    static constexpr int DIM = NCTX + CF;
    static constexpr int OPM = (DIM + 3) / 4;                       // activations per member
    static constexpr int OPMP = OPM <= 2 ? 2 : (OPM <= 4 ? 4 : 8);  // padded for vector LDS
    static constexpr int DIMP = 4 * OPM;
    static constexpr int CFP = (CF + 3) & ~3;
    static_assert(OPM <= 8, "ARM too wide for the quad layout");

    static __device__ __forceinline__ void ld_w(const int32_t *p, int32_t (&w)[8]) {
        if constexpr (OPMP == 2) {
            int2 a = *reinterpret_cast<const int2 *>(p);
            w[0] = a.x; w[1] = a.y;
        } else if constexpr (OPMP == 4) {
            int4 a = *reinterpret_cast<const int4 *>(p);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        } else {
            int4 a = *reinterpret_cast<const int4 *>(p);
            int4 b = *reinterpret_cast<const int4 *>(p + 4);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
            w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        }
    }

    // IFCE inputs of the pixel (y, x): the already decoded coarser grids at (y>>1 >> sh, x>>1 >> sh)
    // (component/coolchic.py:95-100, core/upsampling.py:575-593).  All loads are issued back to back,
    // BEFORE the producer waits for its dependencies: their L2 latency is hidden by that wait.
    static __device__ __forceinline__ void ifce_prefetch(const EntGrid *g, const int8_t *lat, int y, int x, int m,
                                                         int (&ifv)[CCD_IFCE_FAST_MAX]) {
#pragma unroll
        for (int c = 0; c < CCD_IFCE_FAST_MAX; c++) ifv[c] = 0;
        if constexpr (CF > 0) {
            const int n_in = g->ifce_in;
            if (n_in > 0 && (m + 1) * OPM > NCTX) {
                const int yy = y >> 1, xx = x >> 1;
#pragma unroll
                for (int c = 0; c < CCD_IFCE_FAST_MAX; c++) {
                    if (c < n_in) {
                        const int sh = g->ch_sh[c];
                        if (sh >= 0) ifv[c] = lat[g->ch_off[c] + (long long)(yy >> sh) * g->ch_w[c] + (xx >> sh)];
                    }
                }
            }
        }
    }

    // returns (mu, log-scale) in 1/256 units, identical on the 4 lanes of the quad
    static __device__ __forceinline__ void run(const EntGrid *g, const unsigned char *arm_blob,
                                               const unsigned char *ifce_blob, const int (&ifv)[CCD_IFCE_FAST_MAX],
                                               uint32_t rows, uint32_t row_mask, int n_hidden, int y, int x,
                                               int m, long long &o0, long long &o1) {
        const int w = g->w;
        int32_t x0[OPM];
        // ---- my share of the context: causal neighbours (latent.py:148-153) ...
#pragma unroll
        for (int o = 0; o < OPM; o++) {
            const int i = m * OPM + o;
            int v = 0;
            if (i < NCTX) {
                const int yy = y + c_ctx_dy[i], xx = x + c_ctx_dx[i];
                if (yy >= 0 && xx >= 0 && xx < w)
                    v = lds_s8(rows + ((((uint32_t)yy & row_mask) << 6) | ((uint32_t)xx & (CCD_ROW_COLS - 1))));
            }
            x0[o] = v;
        }
        // ---- ... and IFCE features (component/coolchic.py:105-146) from the prefetched inputs
        if constexpr (CF > 0) {
            const int n_in = g->ifce_in;
            if (n_in > 0 && (m + 1) * OPM > NCTX) {
                const int32_t *W = reinterpret_cast<const int32_t *>(ifce_blob);
                const long long *B =
                    reinterpret_cast<const long long *>(ifce_blob + (((size_t)n_in * CFP * 4 + 7) & ~(size_t)7));
                long long acc[OPM];
#pragma unroll
                for (int o = 0; o < OPM; o++) {
                    const int f = m * OPM + o - NCTX;
                    acc[o] = (f >= 0 && f < CF) ? B[f] : 0;
                }
#pragma unroll
                for (int c = 0; c < CCD_IFCE_FAST_MAX; c++) {
                    if (c < n_in) {
                        const int xi = ifv[c] << 16;
#pragma unroll
                        for (int o = 0; o < OPM; o++) {
                            const int f = m * OPM + o - NCTX;
                            if (f >= 0 && f < CF) acc[o] += (long long)W[c * CFP + f] * xi;
                        }
                    }
                }
#pragma unroll
                for (int o = 0; o < OPM; o++) {
                    const int f = m * OPM + o - NCTX;
                    if (f >= 0 && f < CF) {
                        // F.interpolate(ctx.to(torch.float)).to(int64): fp32 round trip (coolchic.py:142-144)
                        x0[o] = (int32_t)__float2ll_rz(__ll2float_rn(acc[o] >> 24));
                    }
                }
            }
        }
        // ---- MLP
        const int32_t *Wh = reinterpret_cast<const int32_t *>(arm_blob);
        const int32_t *Wl = Wh + (size_t)n_hidden * DIM * 4 * OPMP;
        const int32_t *Ws = Wl + DIMP * 2;
        const size_t wbytes = ((size_t)(n_hidden * DIM * 4 * OPMP + DIMP * 4) * 4 + 7) & ~(size_t)7;
        const long long *Bh = reinterpret_cast<const long long *>(arm_blob + wbytes);
        const long long *Bl = Bh + (size_t)n_hidden * DIMP;
        int32_t xo[OPM];
#pragma unroll
        for (int o = 0; o < OPM; o++) {
            x0[o] <<= 16;
            xo[o] = x0[o];
        }
        for (int l = 0; l < n_hidden; l++) {
            long long acc[OPM];
            const long long *B = Bh + (size_t)l * DIMP + m * OPM;
#pragma unroll
            for (int o = 0; o < OPM; o++) acc[o] = B[o];
            const int32_t *W = Wh + ((size_t)l * DIM * 4 + m) * OPMP;
#pragma unroll
            for (int i = 0; i < DIM; i++) {
                const int xi = __shfl_sync(0xffffffffu, xo[i % OPM], i / OPM, 4);
                int32_t wv[8];
                ld_w(W + (size_t)i * 4 * OPMP, wv);
#pragma unroll
                for (int o = 0; o < OPM; o++) acc[o] += (long long)wv[o] * xi;
            }
#pragma unroll
            for (int o = 0; o < OPM; o++) {
                long long a = acc[o];
                a = a < 0 ? 0 : a;
                xo[o] = (int32_t)(a >> 16);
            }
        }
        // last layer (on the hidden state) + stabiliser (on the input): partial sums over my inputs
        long long t0 = 0, t1 = 0;
#pragma unroll
        for (int o = 0; o < OPM; o++) {
            const int i = m * OPM + o;
            const int2 wl = *reinterpret_cast<const int2 *>(Wl + 2 * i);
            const int2 ws = *reinterpret_cast<const int2 *>(Ws + 2 * i);
            t0 += (long long)wl.x * xo[o] + (long long)ws.x * x0[o];
            t1 += (long long)wl.y * xo[o] + (long long)ws.y * x0[o];
        }
#pragma unroll
        for (int d = 1; d < 4; d <<= 1) {
            t0 += (long long)shfl_xor_u64((uint64_t)t0, d);
            t1 += (long long)shfl_xor_u64((uint64_t)t1, d);
        }
        o0 = (t0 + Bl[0] + Bl[2]) >> 24;
        o1 = (t1 + Bl[1] + Bl[3]) >> 24;
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
        for (int f = 0; f < cf; f++) xf[f] = __float2ll_rz(__ll2float_rn(xf[f] >> 24));
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

// cumulative of symbol s from its table entry: leak term + clamps
__device__ __forceinline__ uint32_t fix_left(uint32_t nl, int s) {
    uint32_t l = nl + (uint32_t)(s - kSymMin);
    l = (s <= kSymMin) ? 0u : l;
    l = (s > kSymMax) ? (1u << 24) : l;
    return l;
}

// ---------------------------------------------------------------------------------------
// Producer: symbols [c0, c0 + CHUNK) of diagonal k.  FAST: CHUNK = 8 (quad per symbol),
// GENERIC: CHUNK = 32 (thread per symbol).
template <int NCTX, int CF, bool FAST>
__device__ __forceinline__ void produce_chunk(const SLoc &S, const SmemLayout &sm,
                                              const uint32_t *__restrict__ cdf, int lane, int y0, int x0,
                                              int n_k, int c0, uint32_t ord_diag, uint32_t ord_prev,
                                              int y0_prev, ProfCounters &pc) {
    const EntGrid *g = sm.grid;
    const int member = FAST ? (lane & 3) : 0;
    const int i = c0 + (FAST ? (lane >> 2) : lane);
    const bool valid = i < n_k;
    const int w = g->w;
    const int y = y0 + i;
    const int x = g->raster ? x0 : x0 - CCD_MASK_STRIDE * i;
    const uint32_t ord = ord_diag + (uint32_t)i;
    const uint32_t ring_mask = (uint32_t)S.ring - 1u;
    const uint32_t row_mask = (uint32_t)S.rows - 1u;

    // ---- dependencies: left neighbour decoded (and everything older), ring slot free
    uint32_t need;
    if (g->raster) need = ord;                                      // everything before me
    else if (x > 0) need = ord_prev + (uint32_t)(y - y0_prev) + 1u; // (y, x-1) sits on diagonal k-1
    else need = ord_prev;                                            // first pixel of a row
    const uint32_t need_ring = ord + 1u - (uint32_t)S.ring;          // slot reuse: ord - ring consumed
    if ((int32_t)(need_ring - need) > 0) need = need_ring;
    int ifv[CCD_IFCE_FAST_MAX];
    if constexpr (FAST) QuadArm<NCTX, CF>::ifce_prefetch(g, S.latents, valid ? y : y0, valid ? x : x0, member, ifv);
    // needs grow with the lane: one warp-wide wait on the maximum
    int32_t rel = valid ? (int32_t)(need - ord_diag) : INT32_MIN;
    rel = __reduce_max_sync(0xffffffffu, rel);
    need = ord_diag + (uint32_t)rel;
    PROF_T(t0);
    while ((int32_t)(lds_u32(sm.ctrl) - need) < 0) {
#ifdef CCD_SPIN_SLEEP
        __nanosleep(CCD_SPIN_SLEEP);
#endif
    }
    PROF_ADD(pc.wait, t0);
    PROF_T(t1);
    long long o0 = 0, o1 = 0;
    if constexpr (FAST) {
        // all 32 lanes take part (quad shuffles); out-of-range symbols compute on clamped coordinates
        const int yc = valid ? y : y0, xc = valid ? x : x0;
        QuadArm<NCTX, CF>::run(g, sm.arm, sm.ifce, ifv, sm.rows, row_mask, S.n_hidden, yc, xc, member, o0, o1);
    } else {
        if (valid) {
            const int n_ctx = S.n_ctx, cf = S.cf;
            long long xin[CCD_MAX_DIM];
            for (int t = 0; t < n_ctx; t++) {
                const int yy = y + c_ctx_dy[t], xx = x + c_ctx_dx[t];
                int v = 0;
                if (yy >= 0 && xx >= 0 && xx < w)
                    v = lds_s8(sm.rows + ((((uint32_t)yy & row_mask) << 6) | ((uint32_t)xx & (CCD_ROW_COLS - 1))));
                xin[t] = v;
            }
            if (cf > 0) GenericArm::ifce(g, sm.ifce, S.latents, cf, y >> 1, x >> 1, xin + n_ctx);
            GenericArm::arm(sm.arm, n_ctx + cf, S.n_hidden, xin, o0, o1);
        }
    }
    PROF_ADD(pc.arm, t1);
    PROF_T(t2);
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
    const uint32_t slot = ord & ring_mask;
    const uint32_t wdst = sm.win + slot * (CCD_WIN * 4);
    if constexpr (FAST) {
        if (valid) {
            // member m owns entries [8m, 8m+8)
            const uint4 va = __ldg(row + 2 * member), vb = __ldg(row + 2 * member + 1);
            const int s0 = s_lo + 8 * member;
            sts_v4(wdst + 32u * member, make_uint4(fix_left(va.x, s0), fix_left(va.y, s0 + 1), fix_left(va.z, s0 + 2),
                                                   fix_left(va.w, s0 + 3)));
            const uint4 fb = make_uint4(fix_left(vb.x, s0 + 4), fix_left(vb.y, s0 + 5), fix_left(vb.z, s0 + 6),
                                        fix_left(vb.w, s0 + 7));
            sts_v4(wdst + 32u * member + 16u, fb);
            // hot entry (the coder's steady state reads nothing else): window entries 13 .. 16
            static_assert(CCD_WIN_HALF == 14, "hot entry layout");
            const uint32_t hdst = sm.hot + slot * 16u;
            const bool mirror = slot < CCD_HOT_MIRROR;
            const uint32_t hmir = hdst + (uint32_t)S.ring * 16u;
            if (member == 1) {
                sts_v2(hdst, make_uint2(fb.z, fb.w - fb.z));
                sts_u32(hdst + 8u, fb.y);
                if (mirror) {
                    sts_v2(hmir, make_uint2(fb.z, fb.w - fb.z));
                    sts_u32(hmir + 8u, fb.y);
                }
            } else if (member == 2) {
                const uint32_t l16 = fix_left(va.x, s0);
                sts_u32(hdst + 12u, l16);
                if (mirror) sts_u32(hmir + 12u, l16);
            }
        }
        __threadfence_block();
        __syncwarp();
    } else {
        if (valid) {
            uint4 v[8];
#pragma unroll
            for (int c = 0; c < 8; c++) v[c] = __ldg(row + c);
#pragma unroll
            for (int c = 0; c < 8; c++)
                sts_v4(wdst + 16u * c, make_uint4(fix_left(v[c].x, s_lo + 4 * c), fix_left(v[c].y, s_lo + 4 * c + 1),
                                                  fix_left(v[c].z, s_lo + 4 * c + 2), fix_left(v[c].w, s_lo + 4 * c + 3)));
            {
                const uint32_t l13 = fix_left(v[3].y, s_lo + 13), l14 = fix_left(v[3].z, s_lo + 14);
                const uint32_t l15 = fix_left(v[3].w, s_lo + 15), l16 = fix_left(v[4].x, s_lo + 16);
                sts_v4(sm.hot + slot * 16u, make_uint4(l14, l15 - l14, l13, l16));
                if (slot < CCD_HOT_MIRROR) sts_v4(sm.hot + (slot + (uint32_t)S.ring) * 16u, make_uint4(l14, l15 - l14, l13, l16));
            }
            __threadfence_block();
        }
    }
    if (valid && member == 0) {
        const uint32_t out_off = (uint32_t)(g->lat_off + (long long)y * w + x);
        const uint32_t row_idx = (((uint32_t)y & row_mask) << 6) | ((uint32_t)x & (CCD_ROW_COLS - 1));
        // meta: x = output offset, y = row-ring index | (s_lo + 128) << 16, z = mu_idx | sc_idx << 16, w = tag
        sts_v4(sm.meta + slot * 16u, make_uint4(out_off, row_idx | ((uint32_t)(s_lo + 128) << 16),
                                                (uint32_t)mu_idx | ((uint32_t)sc_idx << 16), ord + 1u));
    }
    PROF_ADD(pc.win, t2);
}

template <int NCTX, int CF, bool FAST>
__device__ __forceinline__ void producer_grid(const SLoc &S, const SmemLayout &sm, const uint32_t *__restrict__ cdf,
                                           int prank, int n_prod, int lane, uint32_t ord_grid, int &chunk_ctr,
                                           ProfCounters &pc) {
    constexpr int CHUNK = FAST ? 8 : 32;
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
        for (int c0 = 0; c0 < n_k; c0 += CHUNK) {
            if (chunk_ctr == prank)
                produce_chunk<NCTX, CF, FAST>(S, sm, cdf, lane, y0, x0, n_k, c0, ord, ord_prev, y0_prev, pc);
            chunk_ctr = (chunk_ctr + 1 == n_prod) ? 0 : chunk_ctr + 1;
        }
        ord_prev = ord;
        y0_prev = y0;
        ord += (uint32_t)n_k;
    }
}

// ---------------------------------------------------------------------------------------
// Range-coder warp.  State: D = point - lower (mod 2^64), R = range.
// (lower and point only ever appear through their difference, so one u64 replaces two.)
struct Coder {          // range ENcoder state
    uint64_t D, R;       // D = lower, R = range
    uint64_t prng;
    int64_t nout;        // words emitted
    uint32_t slow;       // symbols outside the window
};

__device__ __forceinline__ uint32_t load_word(const SLoc &S, int64_t i) {
    return (i < S.n_words) ? __ldg(S.words + i) : 0u;
}

// Exhaustive search with the exact f64 model for a symbol outside the 31-symbol window.
// q: quantile (uniform).  Returns {l0, l1, src, sym}: per-lane cumulatives of the lane's
// candidate in the winning round, the winning lane and the symbol.  (Returned in registers:
// reference outputs would force the hot loop's variables into local memory.)
__device__ __noinline__ uint4 slow_search(uint32_t q, int mu_idx, int sc_idx, const float *scale_tab, int lane) {
    const double mu = (double)(mu_idx - 16384) * (1.0 / 256.0);
    const double b = (double)scale_tab[sc_idx];
    uint32_t ballot = 0, l0 = 0, l1 = 0;
    int s = 0;
    for (int r = 0; r < 4; r++) {
        s = kSymMin + r * 32 + lane;
        l0 = laplace_left_exact(s, mu, b);
        l1 = laplace_left_exact(s + 1, mu, b);
        ballot = __ballot_sync(0xffffffffu, l0 <= q && q < l1);
        if (ballot) break;
    }
    if (ballot == 0u) ballot = 1u;  // unreachable: left(-64) = 0 <= q < 2^24 = left(64)
    const int src = __ffs(ballot) - 1;
    const int sym = __shfl_sync(0xffffffffu, s, src);
    return make_uint4(l0, l1, (uint32_t)src, (uint32_t)sym);
}

__device__ __noinline__ uint2 exact_bounds(int sym, int mu_idx, int sc_idx, const float *scale_tab) {
    const double mu = (double)(mu_idx - 16384) * (1.0 / 256.0);
    const double b = (double)scale_tab[sc_idx];
    return make_uint2(laplace_left_exact(sym, mu, b), laplace_left_exact(sym + 1, mu, b));
}

__device__ __noinline__ void encoder_emit(const SLoc &S, Coder &c, uint32_t L0, uint32_t L1, int lane) {
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

// =======================================================================================
// Range DEcoder = two warps.
//   coder  (warp 15): only the (D, R) recursion, executed identically by all its lanes: a
//          scalar most-probable-first search over the 8 cumulatives around the mode (two
//          broadcast LDS.128, prefetched), pure 64-bit integer ALU on the serial chain.
//          (Measured on B200: every vote / shuffle / shared-memory round trip costs 27-35
//          cycles, a dependent ALU op ~5: lane-parallel candidate evaluation loses.)
//   helper (warp 14): everything that is not on that recursion, 32 symbols at a time:
//          finds how far the producers have got (contiguous valid tags -> `ready`), turns the
//          coder's per-symbol results into symbols, writes the row ring / latent array and
//          advances `progress` for the producers.
// ctrl words: [0] progress (symbols published), [1] ready (symbols whose window is in the ring)
// =======================================================================================
struct DecState {
    uint64_t D, R;        // D = point - lower (mod 2^64), R = range   
    int64_t wpos;         // index of the next unread word
    uint32_t wcur, wnxt;  // lane l holds word (32*chunk + l) of the current / next chunk
    uint32_t wnext;       // word[wpos], broadcast
    uint32_t slow;
    uint32_t n_far;       // symbols that were not the mode (instrumented build)
    int err;
};

__device__ __noinline__ void dec_advance_word(const SLoc &S, DecState &c, int lane) {
    c.wpos++;
    if ((c.wpos & 31) == 0) {
        c.wcur = c.wnxt;
        c.wnxt = load_word(S, c.wpos + 32 + lane);
    }
    c.wnext = __shfl_sync(0xffffffffu, c.wcur, (int)(c.wpos & 31));
}

// result word of a symbol j that is not the mode: [31] value is the symbol itself (else the window index t),
// [30] valid (an all-zero word never matches), [29:8] tag = (j + 1) mod 2^22, [7:0] value
__device__ __forceinline__ uint32_t res_tag(uint32_t j) { return 0x40000000u | (((j + 1u) & 0x3fffffu) << 8); }
__device__ __forceinline__ uint32_t res_word(uint32_t j, uint32_t value, bool is_symbol) {
    return ((uint32_t)is_symbol << 31) | res_tag(j) | (value & 0xffu);
}

// Out-of-line parts of the coder take and return everything BY VALUE (registers), never through a
// reference into the kernel's local-memory state.
struct FarOut {
    uint64_t lo, hi;
    uint32_t rw;
    uint32_t flags;  // 2: outside the window (exact search), 4: desynchronised
};

// Symbol is neither the mode nor one of its two neighbours: the whole 32-entry window at once (lane per
// entry), then the exact f64 model (warp-cooperative).
__device__ __noinline__ FarOut coder_far(uint32_t wrow, uint32_t meta_slot, const float *__restrict__ scale_tab, int lane,
                                         uint32_t j, uint64_t scale, uint64_t D) {
    FarOut o;
    o.flags = 0;
    // lane t evaluates window entry t: one conflict-free LDS, one product, one vote -- the cost does not
    // depend on how far from the mode the symbol is (wide distributions of the coarse grids)
    const uint32_t Lt = lds_u32(wrow + 4u * (uint32_t)lane);
    const uint32_t b = __ballot_sync(0xffffffffu, scale * Lt <= D);  // lefts are non-decreasing: bits 0..t
    if (b != 0u && b != 0xffffffffu) {
        const int t = 31 - __clz((int)b);
        o.lo = scale * (uint64_t)__shfl_sync(0xffffffffu, Lt, t);
        o.hi = scale * (uint64_t)__shfl_sync(0xffffffffu, Lt, t + 1);
        o.rw = res_word(j, (uint32_t)t, false);
        return o;
    }
    // outside the window, or corrupt stream
    o.flags = 2;
    uint64_t q = D / scale;
    if (q >= (1ull << 24)) {
        o.flags |= 4;
        q = (1ull << 24) - 1;
    }
    const uint4 m = lds_v4(meta_slot);
    const uint4 r = slow_search((uint32_t)q, (int)(m.z & 0xffffu), (int)(m.z >> 16), scale_tab, lane);
    o.lo = scale * (uint64_t)__shfl_sync(0xffffffffu, r.x, (int)r.z);
    o.hi = scale * (uint64_t)__shfl_sync(0xffffffffu, r.y, (int)r.z);
    o.rw = res_word(j, r.w, true);
    return o;
}

// Everything that is not "the mode, no renormalisation": the two neighbours of the mode from the hot
// entry, any other symbol through coder_far, the result word for the helper, the renormalisation.
// Out of line (about one symbol in seven): the steady loop stays ~20 instructions per symbol.
struct SlowOut {
    uint32_t d_lo, d_hi, r_lo, r_hi;  // (D, R) after the symbol
    uint32_t flags;                   // 1: consumed the next word, 2 / 4: see FarOut, 8: not the mode
};
__device__ __noinline__ SlowOut coder_slow(uint32_t res_a, uint32_t win_a, uint32_t meta_a,
                                           const float *__restrict__ scale_tab, uint32_t ring_mask, int lane, uint32_t j,
                                           uint4 h, uint64_t D, uint64_t R, uint32_t wnext) {
    // Branches are what costs on this warp (~30 cycles each: predicate wait + fetch redirect, measured), so
    // the mode and its two neighbours are decided with selects; only "none of the three" and the
    // renormalisation bookkeeping branch (both rare).
    constexpr uint32_t M = CCD_WIN_HALF;
    const uint64_t scale = R >> 24;
    const uint64_t lo = scale * h.x, rn = scale * h.y;
    const uint64_t hiM = lo + rn;  // = scale * left(M + 1)
    // which neighbour can it be?  D < lo: only the left one (needs left(M-1)), else only the right one (left(M+2))
    const bool is_m = (D - lo) < rn;
    const bool below = D < lo;
    const uint64_t pn = scale * (below ? h.z : h.w);
    const bool is_l = below & (pn <= D);
    const bool is_r = (!below) & (hiM <= D) & (D < pn);
    uint64_t nlo = is_m ? lo : (is_l ? pn : hiM);
    uint64_t nhi = is_m ? hiM : (is_l ? lo : pn);
    uint32_t rw = res_tag(j) | (is_l ? (M - 1u) : (M + 1u));
    const uint32_t slot = j & ring_mask;
    uint32_t flags = is_m ? 0u : 8u;
    if (!(is_m | is_l | is_r)) {
        const FarOut f = coder_far(win_a + slot * (CCD_WIN * 4), meta_a + slot * 16u, scale_tab, lane, j, scale, D);
        nlo = f.lo;
        nhi = f.hi;
        rw = f.rw;
        flags |= f.flags;
    }
    uint64_t Dn = D - nlo, Rn = nhi - nlo;
    // shared-memory stores of one warp are performed in program order: this word is visible before the
    // `done` store that follows it
    if (!is_m && lane == 0) sts_u32(res_a + slot * 4u, rw);
    const bool renorm = (Rn >> 32) == 0;  // at most one renormalisation per symbol
    Dn = renorm ? ((Dn << 32) | wnext) : Dn;
    Rn = renorm ? (Rn << 32) : Rn;
    flags |= renorm ? 1u : 0u;
    SlowOut o;
    o.d_lo = (uint32_t)Dn;
    o.d_hi = (uint32_t)(Dn >> 32);
    o.r_lo = (uint32_t)Rn;
    o.r_hi = (uint32_t)(Rn >> 32);
    o.flags = flags;
    return o;
}

// The rare follow-ups of coder_slow that touch the (local-memory) decoder state.
__device__ __noinline__ uint32_t coder_bookkeeping(const SLoc &S, DecState &c, int lane, uint32_t flags) {
    if (flags & 1u) dec_advance_word(S, c, lane);
    if (flags & 2u) c.slow++;
    if (flags & 4u) c.err = CCD_ERR_DESYNC;
    return c.wnext;
}

// One symbol of the recursion, executed IDENTICALLY by every lane of the coder warp: no vote, no
// shuffle, no shared-memory round trip on the serial chain -- two 64-bit products, one subtraction and
// one unsigned compare.  h = (left(M), left(M+1) - left(M), ..) of the most probable symbol M:
// D - scale*left(M) < scale*p(M) (wrapping) <=> the symbol is M;  new R = scale*p(M).
// Nothing is written for a mode symbol: the helper learns it from `done` and the absence of a result word.
#ifdef CCD_PROFILE
#define CODER_COUNT_FAR(F) c.n_far += ((F) >> 3) & 1u
#define CODER_SLOW_T0 const long long ts_ = clock64()
#define CODER_SLOW_T1 pc.seg[2] += clock64() - ts_; pc.seg[0]++
#else
#define CODER_COUNT_FAR(F)
#define CODER_SLOW_T0
#define CODER_SLOW_T1
#endif
// The test: with Dn = D - scale*left(M) (wrapping) and Rn = scale*p(M), "symbol is M and no
// renormalisation" <=> Dn < Rn and Rn >= 2^32.  hi32(Dn) < hi32(Rn) implies both; the converse fails only
// when the two high words are equal (probability ~ 1 / hi32(Rn)): those go to coder_slow too, which decides
// exactly.  One 32-bit compare per symbol instead of a 64-bit compare chain plus a range test.
//
// Software pipelining: the branch of symbol j has to wait for that compare, so the products of symbol
// j+1 are issued BEFORE it, from Rn (the state if j is the mode, the common case), and redone after
// coder_slow otherwise.  The loop-carried chain is R -> shift -> multiply -> R.
#define CODER_SLOW(JJ, H)                                                                                   \
    do {                                                                                                    \
        CODER_SLOW_T0;                                                                                      \
        const SlowOut r_ = coder_slow(sm.res, sm.win, sm.meta, scale_tab, ring_mask, lane, (JJ), (H), D, R, wnext); \
        D = ((uint64_t)r_.d_hi << 32) | r_.d_lo;                                                            \
        R = ((uint64_t)r_.r_hi << 32) | r_.r_lo;                                                            \
        CODER_COUNT_FAR(r_.flags);                                                                          \
        if (r_.flags & 7u) wnext = coder_bookkeeping(S, c, lane, r_.flags);                                 \
        CODER_SLOW_T1;                                                                                      \
    } while (0)
// products of a symbol from the current R
#define CODER_PRE(H)                                                                                        \
    do {                                                                                                    \
        const uint64_t scale_ = R >> 24;                                                                    \
        lo = scale_ * (H).x;                                                                                \
        rn = scale_ * (H).y;                                                                                \
    } while (0)
// symbol JJ (products in lo, rn), then the products of the next symbol (hot entry HN).  Written in PTX
// and volatile so that the look-ahead products stay ABOVE the branch (the compiler otherwise proves
// them equal to CODER_PRE after the join and sinks them below it, back onto the serial chain).
__device__ __forceinline__ uint32_t coder_spec(uint64_t D, uint64_t lo, uint64_t rn, uint32_t nL, uint32_t nP, uint64_t &dn,
                                           uint64_t &lo2, uint64_t &rn2) {
    uint32_t dn_lo, dn_hi, l2_lo, l2_hi, r2_lo, r2_hi, ok;
    asm volatile(
        "{\n"
        " .reg .u32 s_lo, s_hi, t;\n"
        " .reg .pred p;\n"
        " sub.cc.u32 %0, %7, %9;\n"
        " subc.u32 %1, %8, %10;\n"
        " shf.r.clamp.b32 s_lo, %11, %12, 24;\n"
        " shr.u32 s_hi, %12, 24;\n"
        " mul.lo.u32 %2, s_lo, %13;\n"
        " mul.hi.u32 t, s_lo, %13;\n"
        " mad.lo.u32 %3, s_hi, %13, t;\n"
        " mul.lo.u32 %4, s_lo, %14;\n"
        " mul.hi.u32 t, s_lo, %14;\n"
        " mad.lo.u32 %5, s_hi, %14, t;\n"
        " setp.lt.u32 p, %1, %12;\n"
        " selp.u32 %6, 1, 0, p;\n"
        "}\n"
        : "=r"(dn_lo), "=r"(dn_hi), "=r"(l2_lo), "=r"(l2_hi), "=r"(r2_lo), "=r"(r2_hi), "=r"(ok)
        : "r"((uint32_t)D), "r"((uint32_t)(D >> 32)), "r"((uint32_t)lo), "r"((uint32_t)(lo >> 32)), "r"((uint32_t)rn),
          "r"((uint32_t)(rn >> 32)), "r"(nL), "r"(nP));
    dn = ((uint64_t)dn_hi << 32) | dn_lo;
    lo2 = ((uint64_t)l2_hi << 32) | l2_lo;
    rn2 = ((uint64_t)r2_hi << 32) | r2_lo;
    return ok;
}
#define CODER_STEP_SPEC(JJ, H, HN)                                                                          \
    do {                                                                                                    \
        uint64_t dn_, lo2_, rn2_;                                                                           \
        if (__builtin_expect(coder_spec(D, lo, rn, (HN).x, (HN).y, dn_, lo2_, rn2_) != 0u, 1)) {                                       \
            D = dn_;                                                                                        \
            R = rn;                                                                                         \
            lo = lo2_;                                                                                      \
            rn = rn2_;                                                                                      \
        } else {                                                                                            \
            CODER_SLOW(JJ, H);                                                                              \
            CODER_PRE(HN);                                                                                  \
        }                                                                                                   \
    } while (0)
// symbol JJ (products in lo, rn) without look-ahead
#define CODER_STEP_LAST(JJ, H)                                                                              \
    do {                                                                                                    \
        const uint64_t dn_ = D - lo;                                                                        \
        if ((uint32_t)(dn_ >> 32) < (uint32_t)(rn >> 32)) {                                                 \
            D = dn_;                                                                                        \
            R = rn;                                                                                         \
        } else {                                                                                            \
            CODER_SLOW(JJ, H);                                                                              \
        }                                                                                                   \
    } while (0)

__device__ __forceinline__ void coder_grid(const SLoc &S, const SmemLayout &sm, const float *__restrict__ scale_tab,
                                           int lane, uint32_t ord_begin, uint32_t ord_end, DecState &c,
                                           ProfCounters &pc) {
    const uint32_t ring_mask = (uint32_t)S.ring - 1u;
    const uint32_t ready_a = sm.ctrl + 4u, done_a = sm.ctrl + 8u;
    uint64_t D = c.D, R = c.R;
    uint64_t lo = 0, rn = 0;
    uint32_t wnext = c.wnext;
    uint32_t j = ord_begin;
    uint32_t limit = ord_begin;  // symbols < limit have their window in the ring
    auto refresh = [&]() {
        const uint32_t r = lds_u32(ready_a);
        limit = ((int32_t)(r - ord_end) > 0) ? ord_end : r;
    };
    auto hot_of = [&](uint32_t jj) { return lds_v4(sm.hot + (jj & ring_mask) * 16u); };
    while (j != ord_end) {
        if ((int32_t)(limit - j) <= 0) {
            PROF_T(t0);
            do {
                refresh();
            } while ((int32_t)(limit - j) <= 0);
            PROF_ADD(pc.wait, t0);
        }
        if ((int32_t)(limit - j) >= 3) {
            // at least a triple is ready.  Steady state (9 or more ready): six symbols per round trip through
            // the loop: the hot entries of the next triple
            // are requested while a triple is decoded, and the two register sets swap roles (no copies);
            // the ring's first entries are mirrored behind its end, so one address serves a triple
            uint32_t o = sm.hot + (j & ring_mask) * 16u;
            uint4 a0 = lds_v4(o), a1 = lds_v4(o + 16u), a2 = lds_v4(o + 32u);
            CODER_PRE(a0);
            while ((int32_t)(limit - j) >= 9) {
                o = sm.hot + ((j + 3u) & ring_mask) * 16u;
                const uint4 b0 = lds_v4(o), b1 = lds_v4(o + 16u), b2 = lds_v4(o + 32u);
                CODER_STEP_SPEC(j, a0, a1);
                CODER_STEP_SPEC(j + 1u, a1, a2);
                CODER_STEP_SPEC(j + 2u, a2, b0);
                sts_u32(done_a, j + 3u);  // every lane stores the same word: no predicate on the hot path
                o = sm.hot + ((j + 6u) & ring_mask) * 16u;
                a0 = lds_v4(o);
                a1 = lds_v4(o + 16u);
                a2 = lds_v4(o + 32u);
                CODER_STEP_SPEC(j + 3u, b0, b1);
                CODER_STEP_SPEC(j + 4u, b1, b2);
                CODER_STEP_SPEC(j + 5u, b2, a0);
                j += 6u;
                sts_u32(done_a, j);
#ifdef CCD_PROFILE
                pc.seg[3] += 6;
#endif
                if ((int32_t)(limit - j) < 9) refresh();
            }
            // a0..a2 are valid (limit - j >= 3 here), lo / rn belong to a0
            CODER_STEP_SPEC(j, a0, a1);
            CODER_STEP_SPEC(j + 1u, a1, a2);
            CODER_STEP_LAST(j + 2u, a2);
            j += 3u;
            sts_u32(done_a, j);
        } else {
            const uint4 a0 = hot_of(j);
            CODER_PRE(a0);
            CODER_STEP_LAST(j, a0);
            j++;
            sts_u32(done_a, j);
#ifdef CCD_PROFILE
            pc.seg[4]++;
#endif
        }
    }
    c.D = D;
    c.R = R;
    c.wnext = wnext;
}

// Helper warp: readiness scan + publication, 32 symbols per round (lane = symbol).
__device__ __forceinline__ void helper_grid(const SLoc &S, const SmemLayout &sm, int lane, uint32_t ord_begin,
                                            uint32_t ord_end, ProfCounters &pc) {
    const uint32_t ring_mask = (uint32_t)S.ring - 1u;
    uint32_t r = ord_begin, p = ord_begin;
    while (p != ord_end) {
#ifdef CCD_PROFILE
        pc.seg[0]++;  // rounds
#endif
        if (r != ord_end) {
            const uint32_t jj = r + (uint32_t)lane;
            bool ok = false;
            if ((int32_t)(ord_end - jj) > 0) ok = lds_u32(sm.meta + (jj & ring_mask) * 16u + 12u) == jj + 1u;
            const uint32_t b = __ballot_sync(0xffffffffu, ok);
            const uint32_t cnt = (b == 0xffffffffu) ? 32u : (uint32_t)(__ffs(~b) - 1);
            if (cnt) {
                r += cnt;
                if (lane == 0) sts_u32(sm.ctrl + 4u, r);
            }
#ifdef CCD_PROFILE
            pc.seg[1] += cnt;
            if (cnt == 32u) pc.seg[2]++;
            pc.seg[3] += (int32_t)(r - p);  // ready - published
#endif
        }
        {
            // symbols [p, done) are decoded: mode symbols left no trace, the others a tagged result word
            const uint32_t d = lds_u32(sm.ctrl + 8u);
            const int32_t avail = (int32_t)(d - p);
            const uint32_t cnt = avail <= 0 ? 0u : (avail > 32 ? 32u : (uint32_t)avail);
            if ((uint32_t)lane < cnt) {
                const uint32_t jj = p + (uint32_t)lane;
                const uint32_t slot = jj & ring_mask;
                const uint32_t w = lds_u32(sm.res + slot * 4u);
                const uint4 m = lds_v4(sm.meta + slot * 16u);
                const bool other = (w & 0x7fffff00u) == res_tag(jj);
                const int base = (int)(m.y >> 16) - 128;  // s_lo
                int sym = base + CCD_WIN_HALF;
                if (other) {
                    const int v = (int)(w & 0xffu);
                    sym = (w >> 31) ? (int)(int8_t)v : base + v;
                    sts_u32(sm.res + slot * 4u, 0u);  // no stale word survives a trip around the ring
                }
                sts_u8(sm.rows + (m.y & 0xffffu), sym);
                S.latents[m.x] = (int8_t)sym;
            }
            __syncwarp();
            if (cnt) {
                p += cnt;
                if (lane == 0) sts_u32(sm.ctrl, p);
            }
        }
    }
}

// ---- encode (MODE 1: the latents given in S.latents; MODE 2: draw them from the model).
// Not performance critical: used to fabricate synthetic streams.  Runs in the coder warp and
// publishes by itself (the helper warp idles).
template <int MODE>
__device__ __noinline__ void encode_grid(const SLoc &S, const SmemLayout &sm, const float *__restrict__ scale_tab,
                                         int lane, uint32_t ord_begin, uint32_t ord_end, Coder &c) {
    const uint32_t ring_mask = (uint32_t)S.ring - 1u;
    for (uint32_t j = ord_begin; j != ord_end; j++) {
        const uint32_t slot = j & ring_mask;
        uint4 m;
        do {
            m = lds_v4(sm.meta + slot * 16u);
        } while (!__all_sync(0xffffffffu, m.w == j + 1u));
        const uint32_t L0 = lds_u32(sm.win + slot * (CCD_WIN * 4) + (uint32_t)lane * 4u);
        const uint32_t L1 = __shfl_down_sync(0xffffffffu, L0, 1);  // lane 31 keeps L0: empty interval
        const int mu_idx = (int)(m.z & 0xffffu), sc_idx = (int)(m.z >> 16);
        const int s_lo = (int)(m.y >> 16) - 128;
        uint32_t l0, l1;
        int src = -1, sym;
        if constexpr (MODE == 2) {
            const uint32_t q = (uint32_t)(splitmix64(c.prng) >> 40);
            const uint32_t ballot = __ballot_sync(0xffffffffu, L0 <= q && q < L1);
            if (ballot) {
                src = __ffs(ballot) - 1;
                sym = s_lo + src;
                l0 = __shfl_sync(0xffffffffu, L0, src);
                l1 = __shfl_sync(0xffffffffu, L1, src);
            } else {
                c.slow++;
                const uint4 r = slow_search(q, mu_idx, sc_idx, scale_tab, lane);
                src = (int)r.z;
                sym = (int)r.w;
                l0 = __shfl_sync(0xffffffffu, r.x, src);
                l1 = __shfl_sync(0xffffffffu, r.y, src);
            }
        } else {
            sym = S.latents[m.x];
            if (sym >= s_lo && sym < s_lo + 31) {
                src = sym - s_lo;
                l0 = __shfl_sync(0xffffffffu, L0, src);
                l1 = __shfl_sync(0xffffffffu, L1, src);
            } else {
                c.slow++;
                const uint2 r = exact_bounds(sym, mu_idx, sc_idx, scale_tab);
                l0 = r.x;
                l1 = r.y;
            }
        }
        encoder_emit(S, c, l0, l1, lane);
        if (lane == 0) {
            sts_u8(sm.rows + (m.y & 0xffffu), sym);
            S.latents[m.x] = (int8_t)sym;
            sts_u32(sm.ctrl, j + 1u);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void named_barrier(int nthreads) {
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------
template <int NCTX, int CF, bool FAST>
__global__ void __launch_bounds__(CCD_ENT_THREADS, 1)
    k_entropy(const EntStream *__restrict__ streams, const uint32_t *__restrict__ cdf,
              const float *__restrict__ scale_tab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EntStream &G = streams[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // role assignment: warp 15 = range coder (recursion), warp 14 = its helper; producer warps are
    // those enabled in prod_mask (by default the coder keeps its scheduler partition for itself)
    const uint32_t prod_mask = G.prod_mask & ((1u << (CCD_ENT_WARPS - 2)) - 1u);
    const bool is_coder = (warp == CCD_ENT_WARPS - 1);
    const bool is_helper = (warp == CCD_ENT_WARPS - 2);
    const bool is_prod = !is_coder && !is_helper && ((prod_mask >> warp) & 1u);
    if (!is_coder && !is_helper && !is_prod) return;
    const int n_prod = __popc(prod_mask);
    const int prank = __popc(prod_mask & ((1u << warp) - 1u));
    const int n_active = (n_prod + 2) * 32;
    // dense index among active threads
    const int atid = is_coder ? (n_prod + 1) * 32 + lane : (is_helper ? n_prod * 32 + lane : prank * 32 + lane);

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
    S.status = G.status;
    S.n_symbols = G.n_symbols;

    // one-time: control words, meta tags, ARM parameters
    if (atid < 16) sts_u32(sm.ctrl + 4u * atid, 0u);
    for (int i = atid; i < S.ring; i += n_active) {
        sts_v4(sm.meta + 16u * i, make_uint4(0, 0, 0, 0));
        sts_u32(sm.res + 4u * i, 0u);
    }
    for (int i = atid * 4; i < G.arm_blob_bytes; i += n_active * 4)
        *reinterpret_cast<uint32_t *>(sm.arm + i) = *reinterpret_cast<const uint32_t *>(G.blob + i);

    Coder cd;
    cd.slow = 0;
    cd.nout = 0;
    cd.prng = G.seed;
    cd.D = 0;
    cd.R = ~0ull;
    DecState ds;
    ds.err = 0;
    ds.slow = 0;
    ds.n_far = 0;
    ds.wpos = 2;
    ds.wcur = ds.wnxt = ds.wnext = 0;
    ds.D = 0;
    ds.R = ~0ull;
    if (is_coder && S.mode == 0) {
        ds.wcur = load_word(S, lane);
        ds.wnxt = load_word(S, 32 + lane);
        const uint32_t w0 = __shfl_sync(0xffffffffu, ds.wcur, 0);
        const uint32_t w1 = __shfl_sync(0xffffffffu, ds.wcur, 1);
        ds.D = ((uint64_t)w0 << 32) | w1;
        ds.wnext = __shfl_sync(0xffffffffu, ds.wcur, 2);
    }
    uint32_t ord = 0;
    int chunk_ctr = 0;
    ProfCounters pc;
    const int n_grids = G.n_grids;
    for (int gi = 0; gi < n_grids; gi++) {
        named_barrier(n_active);  // previous grid fully decoded, its latents visible CTA-wide
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&G.grid[gi]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(sm.grid);
            for (int i = atid; i < (int)(sizeof(EntGrid) / 4); i += n_active) dst[i] = src[i];
            const EntGrid &Gg = G.grid[gi];
            for (int i = atid * 4; i < Gg.ifce_blob_bytes; i += n_active * 4)
                *reinterpret_cast<uint32_t *>(sm.ifce + i) =
                    *reinterpret_cast<const uint32_t *>(G.blob + Gg.ifce_blob_off + i);
        }
        named_barrier(n_active);
        const uint32_t n_sym = (uint32_t)sm.grid->h * (uint32_t)sm.grid->w;
        PROF_T(tg);
        if (is_coder) {
            if (S.mode == 0) coder_grid(S, sm, scale_tab, lane, ord, ord + n_sym, ds, pc);
            else if (S.mode == 1) encode_grid<1>(S, sm, scale_tab, lane, ord, ord + n_sym, cd);
            else encode_grid<2>(S, sm, scale_tab, lane, ord, ord + n_sym, cd);
        } else if (is_helper) {
            if (S.mode == 0) helper_grid(S, sm, lane, ord, ord + n_sym, pc);
        } else {
            producer_grid<NCTX, CF, FAST>(S, sm, cdf, prank, n_prod, lane, ord, chunk_ctr, pc);
        }
        PROF_ADD(pc.total, tg);
        ord += n_sym;
    }
#ifdef CCD_PROFILE
    // [4] coder wait, [5] coder total, [6] producers wait, [7] arm, [8] window, [9] total (kilo-cycles, lane 0 of each warp)
    if (lane == 0 && is_helper) {
        G.status[6] = (int)pc.seg[0];
        G.status[7] = (int)pc.seg[1];
        G.status[8] = (int)pc.seg[2];
        G.status[9] = (int)(pc.seg[3] >> 4);
        G.status[15] = (int)(pc.total >> 10);
    } else if (lane == 0) {
        if (is_coder) {
            atomicAdd(&G.status[4], (int)(pc.wait >> 10));
            atomicAdd(&G.status[5], (int)(pc.total >> 10));
            G.status[10] = (int)pc.seg[0];
            G.status[11] = (int)ds.n_far;
            G.status[12] = (int)(pc.seg[2] >> 10);
            G.status[13] = (int)pc.seg[3];  // symbols decoded in the steady loop
            G.status[14] = (int)pc.seg[4];  // symbols decoded one at a time (coder close behind the producers)
        }
    }
#endif
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
            G.status[0] = ds.err;
            G.status[1] = (int32_t)ds.wpos;
            G.status[2] = (int32_t)(ds.slow + cd.slow);
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
    g_ccd_launches++;
    return (int)cudaGetLastError();
}

}  // namespace

unsigned long long g_ccd_launches = 0;

size_t ccd_entropy_smem_bytes(int ring, int rows, int arm_blob_bytes, int ifce_blob_max) {
    size_t p = 64 + align16(sizeof(EntGrid)) + align16((size_t)arm_blob_bytes) + align16((size_t)ifce_blob_max);
    p += 16 + (size_t)ring * 16 + (size_t)ring * CCD_WIN * 4 + (size_t)(ring + CCD_HOT_MIRROR) * 16 + (size_t)ring * 4 +
         (size_t)rows * CCD_ROW_COLS;
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

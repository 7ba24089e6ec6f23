// fft_device.cuh -- in-register radix-2/4/8/16 butterflies and a Stockham
// shared-memory complex FFT for sm_100a.
//
// This is the arithmetic the reference delegates to FFTW
// (reference fft_lines/fft_calculate.c:24,29,40,45: forward, unnormalised
// complex-to-complex DFT, fp32).  FP32 CUDA cores only: the transform is not a
// dense contraction, so tcgen05 is deliberately not used (BASELINE north_star).
//
// Layout: a transform of size N lives in shared memory as N float2 with one
// pad element every 16 (index i -> i + i/16), which makes every pass's
// stride-N/R loads and stride-NS stores conflict-free for 64-bit accesses.
// Each thread holds VPT complex values per pass (VPT/R butterflies of radix R).
#pragma once
#include <cuda_runtime.h>

namespace fftl {

// Complex helpers.  FFTL_F32X2 (the fast-path translation units): Blackwell's packed FP32 instructions (FADD2 / FMUL2 /
// FFMA2 on a 64-bit register pair, with half swizzle, per-half negate and scalar-broadcast operands) do the two halves
// of a complex operation in one issue slot: complex add = 1 instruction, multiply = 2, a + w*b = 2.  Every packed form
// performs exactly the two IEEE operations of its scalar twin below (same operands, same roundings), so the results
// are bit-identical; the compiler folds the swizzles and negations into operand modifiers only without flush-to-zero
// (those units are built with -ftz=false).
#ifdef FFTL_F32X2
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    const float2 t = __fmul2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x));
    return __ffma2_rn(make_float2(a.x, a.x), b, make_float2(-t.x, t.y));
}
// a - i b, a + i b
__device__ __forceinline__ float2 csub_i(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }
__device__ __forceinline__ float2 cadd_i(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a - i b, a + i b
__device__ __forceinline__ float2 csub_i(float2 a, float2 b) { return make_float2(a.x + b.y, a.y - b.x); }
__device__ __forceinline__ float2 cadd_i(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }
#endif
// F floats (one per frame of a tile, F even) as F/2 packed pairs for the tail's sums, scales and filter taps: the same
// IEEE operations as the scalar loops, one issue slot per two frames.  PK chooses per kernel instantiation: measured
// +4.3 % at N = 8192 (512 bars), -0.7 % at N = 4096 / 200 bars, +-0.3 % elsewhere (gpurun_out/ab_tail.log).
template <int F, bool PK> __device__ __forceinline__ void vec_add(float* acc, const float* u) // acc += u
{
#ifdef FFTL_F32X2
    if constexpr (PK) {
#pragma unroll
        for (int i = 0; i < F; i += 2) {
            const float2 r = __fadd2_rn(make_float2(acc[i], acc[i + 1]), make_float2(u[i], u[i + 1]));
            acc[i] = r.x;
            acc[i + 1] = r.y;
        }
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < F; i++) acc[i] += u[i];
}
template <int F, bool PK> __device__ __forceinline__ void vec_scale(float* acc, float s) // acc *= s, round to nearest, never fused
{
#ifdef FFTL_F32X2
    if constexpr (PK) {
#pragma unroll
        for (int i = 0; i < F; i += 2) {
            const float2 r = __fmul2_rn(make_float2(acc[i], acc[i + 1]), make_float2(s, s));
            acc[i] = r.x;
            acc[i + 1] = r.y;
        }
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < F; i++) acc[i] = __fmul_rn(acc[i], s);
}
template <int F, bool PK> __device__ __forceinline__ void vec_fma(float c, const float* u, float* y) // y = c*u + y
{
#ifdef FFTL_F32X2
    if constexpr (PK) {
#pragma unroll
        for (int i = 0; i < F; i += 2) {
            const float2 r = __ffma2_rn(make_float2(c, c), make_float2(u[i], u[i + 1]), make_float2(y[i], y[i + 1]));
            y[i] = r.x;
            y[i + 1] = r.y;
        }
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < F; i++) y[i] = fmaf(c, u[i], y[i]);
}

// a * (-i)
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

__device__ __forceinline__ void dft2(float2& a, float2& b)
{
    float2 t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

// forward 4-point DFT in place, natural order
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3)
{
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = csub_i(t1, d);
    a3 = cadd_i(t1, d);
}

// sqrt.approx.ftz.f32 (one MUFU): what --use_fast_math makes of sqrtf; spelled out because the fast-path units are
// built without flush-to-zero (FFTL_F32X2), where sqrtf would become the slower denormal-aware sequence
__device__ __forceinline__ float sqrt_approx_ftz(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#define FFTL_SQRT1_2 0.70710678118654752440f
#define FFTL_COS_PI_8 0.92387953251128675613f
#define FFTL_SIN_PI_8 0.38268343236508977173f

#ifdef FFTL_F32X2
// multiply by W8^1 = (1-i)/sqrt2, W8^3 = (-1-i)/sqrt2
__device__ __forceinline__ float2 cmul_w8_1(float2 a) { return __fmul2_rn(__fadd2_rn(a, make_float2(a.y, -a.x)), make_float2(FFTL_SQRT1_2, FFTL_SQRT1_2)); }
__device__ __forceinline__ float2 cmul_w8_3(float2 a) { return __fmul2_rn(__fadd2_rn(make_float2(a.y, -a.y), make_float2(-a.x, -a.x)), make_float2(FFTL_SQRT1_2, FFTL_SQRT1_2)); }
// W16^1 = (c,-s), W16^3 = (s,-c), W16^9 = -W16^1   (c = cos(pi/8), s = sin(pi/8))
__device__ __forceinline__ float2 cmul_w16_1(float2 a) { return __ffma2_rn(a, make_float2(FFTL_COS_PI_8, FFTL_COS_PI_8), __fmul2_rn(make_float2(a.y, -a.x), make_float2(FFTL_SIN_PI_8, FFTL_SIN_PI_8))); }
__device__ __forceinline__ float2 cmul_w16_3(float2 a) { return __ffma2_rn(a, make_float2(FFTL_SIN_PI_8, FFTL_SIN_PI_8), __fmul2_rn(make_float2(a.y, -a.x), make_float2(FFTL_COS_PI_8, FFTL_COS_PI_8))); }
#else
// multiply by W8^1 = (1-i)/sqrt2, W8^3 = (-1-i)/sqrt2
__device__ __forceinline__ float2 cmul_w8_1(float2 a) { return make_float2((a.x + a.y) * FFTL_SQRT1_2, (a.y - a.x) * FFTL_SQRT1_2); }
__device__ __forceinline__ float2 cmul_w8_3(float2 a) { return make_float2((a.y - a.x) * FFTL_SQRT1_2, -(a.x + a.y) * FFTL_SQRT1_2); }
// W16^1 = (c,-s), W16^3 = (s,-c), W16^9 = -W16^1   (c = cos(pi/8), s = sin(pi/8))
__device__ __forceinline__ float2 cmul_w16_1(float2 a) { return make_float2(fmaf(a.x, FFTL_COS_PI_8, a.y * FFTL_SIN_PI_8), fmaf(a.y, FFTL_COS_PI_8, -a.x * FFTL_SIN_PI_8)); }
__device__ __forceinline__ float2 cmul_w16_3(float2 a) { return make_float2(fmaf(a.x, FFTL_SIN_PI_8, a.y * FFTL_COS_PI_8), fmaf(a.y, FFTL_SIN_PI_8, -a.x * FFTL_COS_PI_8)); }
#endif
__device__ __forceinline__ float2 cmul_w16_9(float2 a) { float2 t = cmul_w16_1(a); return make_float2(-t.x, -t.y); }

// In-register forward DFT of R points.  Output X[k] ends up in register
// v[rev<R>(k)]; use out_index<R>(r) to get k from the register number.
template <int R> struct Radix;

template <> struct Radix<2> {
    __device__ __forceinline__ static void run(float2* v) { dft2(v[0], v[1]); }
    __host__ __device__ __forceinline__ static constexpr int out_index(int r) { return r; }
};
template <> struct Radix<4> {
    __device__ __forceinline__ static void run(float2* v) { dft4(v[0], v[1], v[2], v[3]); }
    __host__ __device__ __forceinline__ static constexpr int out_index(int r) { return r; }
};
template <> struct Radix<8> {
    // n = n1 + 2*n2, k = 4*k1 + k2
    __device__ __forceinline__ static void run(float2* v)
    {
        dft4(v[0], v[2], v[4], v[6]);
        dft4(v[1], v[3], v[5], v[7]);
        v[3] = cmul_w8_1(v[3]);
        v[5] = cmul_mi(v[5]);
        v[7] = cmul_w8_3(v[7]);
        dft2(v[0], v[1]);
        dft2(v[2], v[3]);
        dft2(v[4], v[5]);
        dft2(v[6], v[7]);
    }
    __host__ __device__ __forceinline__ static constexpr int out_index(int r) { return 4 * (r & 1) + (r >> 1); }
};
template <> struct Radix<16> {
    // n = n1 + 4*n2, k = 4*k1 + k2
    __device__ __forceinline__ static void run(float2* v)
    {
#pragma unroll
        for (int n1 = 0; n1 < 4; n1++) dft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
        // v[n1 + 4*k2] *= W16^(n1*k2)
        v[5] = cmul_w16_1(v[5]);   // n1=1,k2=1
        v[9] = cmul_w8_1(v[9]);    // n1=1,k2=2 : W16^2
        v[13] = cmul_w16_3(v[13]); // n1=1,k2=3
        v[6] = cmul_w8_1(v[6]);    // n1=2,k2=1 : W16^2
        v[10] = cmul_mi(v[10]);    // n1=2,k2=2 : W16^4
        v[14] = cmul_w8_3(v[14]);  // n1=2,k2=3 : W16^6
        v[7] = cmul_w16_3(v[7]);   // n1=3,k2=1
        v[11] = cmul_w8_3(v[11]);  // n1=3,k2=2 : W16^6
        v[15] = cmul_w16_9(v[15]); // n1=3,k2=3
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) dft4(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
    }
    __host__ __device__ __forceinline__ static constexpr int out_index(int r) { return 4 * (r & 3) + (r >> 2); }
};

// ---- radix-16 with the multiplies folded into the butterflies (fast path only) ----------------------
// a + w*b with both products as fused multiply-adds (4 FFMA), and the matching difference as 2a - (a + w*b)
// (2 FFMA): a twiddled radix-2 costs 6 instructions instead of 8.
#ifdef FFTL_F32X2
__device__ __forceinline__ float2 cfma(float2 a, float2 w, float2 b) // a + w*b
{
    const float2 t = __ffma2_rn(make_float2(b.y, b.x), make_float2(-w.y, w.y), a);
    return __ffma2_rn(make_float2(w.x, w.x), b, t);
}
__device__ __forceinline__ float2 ctwice_minus(float2 a, float2 t) { return __ffma2_rn(make_float2(2.0f, 2.0f), a, make_float2(-t.x, -t.y)); }
#else
__device__ __forceinline__ float2 cfma(float2 a, float2 w, float2 b) // a + w*b
{
    return make_float2(fmaf(w.x, b.x, fmaf(-w.y, b.y, a.x)), fmaf(w.x, b.y, fmaf(w.y, b.x, a.y)));
}
__device__ __forceinline__ float2 ctwice_minus(float2 a, float2 t) { return make_float2(fmaf(2.0f, a.x, -t.x), fmaf(2.0f, a.y, -t.y)); }
#endif

// forward 4-point DFT of (a0, w1*b1, w2*b2, w3*b3), natural order; a0 comes in already multiplied
__device__ __forceinline__ void dft4_tw(float2& a0, float2& b1, float2& b2, float2& b3, float2 w1, float2 w2, float2 w3)
{
    const float2 t0 = cfma(a0, w2, b2), t1 = ctwice_minus(a0, t0);   // a0 +- w2 b2
    const float2 a1 = cmul(b1, w1);
    const float2 t2 = cfma(a1, w3, b3), d = ctwice_minus(a1, t2);    // a1 +- w3 b3
    a0 = cadd(t0, t2);
    b2 = csub(t0, t2);
    b1 = csub_i(t1, d);                                              // t1 - i d
    b3 = cadd_i(t1, d);                                              // t1 + i d
}

// Radix<16>::run with (TW) every input but v[0] first multiplied by tw[r]; same output order.
template <bool TW> __device__ __forceinline__ void radix16_fused(float2* v, const float2* tw)
{
    constexpr float2 W3 = {FFTL_SIN_PI_8, -FFTL_COS_PI_8}, W9 = {-FFTL_COS_PI_8, FFTL_SIN_PI_8}; // W16^3, W16^9
    if (TW) {
        dft4_tw(v[0], v[4], v[8], v[12], tw[4], tw[8], tw[12]);
#pragma unroll
        for (int n1 = 1; n1 < 4; n1++) {
            v[n1] = cmul(v[n1], tw[n1]);
            dft4_tw(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12], tw[n1 + 4], tw[n1 + 8], tw[n1 + 12]);
        }
    } else {
#pragma unroll
        for (int n1 = 0; n1 < 4; n1++) dft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    }
    // second stage: v[n1 + 4*k2] *= W16^(n1*k2), then 4-point DFTs over n1
    dft4(v[0], v[1], v[2], v[3]);
    {   // k2 = 1: (1, W16^1, W16^2, W16^3)
        const float2 b2 = cmul_w8_1(v[6]);
        const float2 t0 = cadd(v[4], b2), t1 = csub(v[4], b2);
        const float2 a1 = cmul_w16_1(v[5]);
        const float2 t2 = cfma(a1, W3, v[7]), d = ctwice_minus(a1, t2);
        v[4] = cadd(t0, t2);
        v[6] = csub(t0, t2);
        v[5] = csub_i(t1, d);
        v[7] = cadd_i(t1, d);
    }
    v[9] = cmul_w8_1(v[9]);    // k2 = 2: (1, W16^2, W16^4, W16^6)
    v[10] = cmul_mi(v[10]);
    v[11] = cmul_w8_3(v[11]);
    dft4(v[8], v[9], v[10], v[11]);
    {   // k2 = 3: (1, W16^3, W16^6, W16^9)
        const float2 b2 = cmul_w8_3(v[14]);
        const float2 t0 = cadd(v[12], b2), t1 = csub(v[12], b2);
        const float2 a1 = cmul_w16_3(v[13]);
        const float2 t2 = cfma(a1, W9, v[15]), d = ctwice_minus(a1, t2);
        v[12] = cadd(t0, t2);
        v[14] = csub(t0, t2);
        v[13] = csub_i(t1, d);
        v[15] = cadd_i(t1, d);
    }
}

__host__ __device__ __forceinline__ constexpr int pad_index(int i) { return i + (i >> 4); }
__host__ __device__ __forceinline__ constexpr int padded_size(int n) { return n + (n >> 4); }

// One Stockham pass of radix R with NS = product of earlier radices, over an
// N-point transform held in `s` (padded layout), executed by `G` cooperating
// threads (lane id g in [0,G)), VPT = N/G values per thread.
//   LOAD_SMEM  = false: v[] already holds this thread's inputs
//                (v[t*R + r] = x[(g + t*G) + r*N/R]) -- used by the first pass
//                when the inputs come straight from global memory.
//   STORE_SMEM = false: leave results in v[] (v[t*R + rr], X index via
//                Radix<R>::out_index) -- lets the caller fuse an epilogue.
// tw: W_N^m = exp(-2*pi*i*m/N), m < N, fp32 from double.
// `sync()` must be a barrier over all threads that share `s`.
template <int N, int R, int NS, int G, bool LOAD_SMEM, typename Sync>
__device__ __forceinline__ void stockham_pass(float2* v, float2* s, const float2* __restrict__ tw, int g, Sync sync)
{
    constexpr int VPT = N / G;
    constexpr int T = VPT / R; // butterflies per thread
    static_assert(VPT % R == 0 && T >= 1, "radix must divide values-per-thread");
    if (LOAD_SMEM) {
#pragma unroll
        for (int t = 0; t < T; t++) {
            int j = g + t * G;
#pragma unroll
            for (int r = 0; r < R; r++) v[t * R + r] = s[pad_index(j + r * (N / R))];
        }
    }
#pragma unroll
    for (int t = 0; t < T; t++) {
        int j = g + t * G;
        if (NS > 1) {
            int k = j & (NS - 1);
            int base = k * (N / (NS * R));
#pragma unroll
            for (int r = 1; r < R; r++) v[t * R + r] = cmul(v[t * R + r], __ldg(&tw[base * r]));
        }
        Radix<R>::run(&v[t * R]);
    }
    sync(); // everyone has read the old contents
#pragma unroll
    for (int t = 0; t < T; t++) {
        int j = g + t * G;
        int k = j & (NS - 1);
        int j0 = (j - k) * R + k;
#pragma unroll
        for (int rr = 0; rr < R; rr++) s[pad_index(j0 + Radix<R>::out_index(rr) * NS)] = v[t * R + rr];
    }
    sync();
}

// Radix schedule: 16,16,... then one tail pass (2,4 or 8) when log2 N is not
// a multiple of 4.  Pass 0 takes its inputs from registers.
template <int N, int G, typename Sync>
__device__ __forceinline__ void fft_from_regs(float2* v, float2* s, const float2* __restrict__ tw, int g, Sync sync)
{
    static_assert(N >= 256 && N <= 16384, "supported sizes");
    stockham_pass<N, 16, 1, G, false>(v, s, tw, g, sync);
    stockham_pass<N, 16, 16, G, true>(v, s, tw, g, sync);
    if constexpr (N == 512) stockham_pass<N, 2, 256, G, true>(v, s, tw, g, sync);
    if constexpr (N == 1024) stockham_pass<N, 4, 256, G, true>(v, s, tw, g, sync);
    if constexpr (N == 2048) stockham_pass<N, 8, 256, G, true>(v, s, tw, g, sync);
    if constexpr (N >= 4096) stockham_pass<N, 16, 256, G, true>(v, s, tw, g, sync);
    if constexpr (N == 8192) stockham_pass<N, 2, 4096, G, true>(v, s, tw, g, sync);
    if constexpr (N == 16384) stockham_pass<N, 4, 4096, G, true>(v, s, tw, g, sync);
}

} // namespace fftl

// stft_rdif.cuh -- the fast path of the fused STFT->bars kernel (sm_100a).
//
// Real-input decimation-in-frequency transform of size N = 16*M, built so that
// the FP32 pipe -- not shared-memory bandwidth -- is the limiter:
//
//   pass A  thread t (0..M-1) owns column t: the 16 samples x[t + M*q] of a frame.
//           hop = HOPQ*M, so consecutive frames of the tile reuse the same
//           registers (frame f uses xr[q + HOPQ*f]): every int16 sample is loaded
//           and converted once per tile, not once per frame.
//           Real 16-point DFT over q (66 flops, half of a complex one) gives
//           Y_t[k0], k0 = 0..8; twiddle by W_N^(t*k0) (register-resident).
//   pass B/C for each k0 a complex M-point FFT over t yields X[k0 + 16 m].
//           k0 = 1..7 are full complex sub-transforms; k0 = 0 and k0 = 8 have real
//           inputs, so frames f and f^1 share one complex sub-transform each
//           (untangled with half-warp shuffles).  8 sub-transforms per frame.
//           One sub-transform is done by M/16 lanes with warp-level syncs only.
//   bins    magnitudes -> padded per-frame array -> balanced segment sums ->
//           bars -> Savitzky-Golay -> zoom -> float/int heights, band energy.
//
// What the reference does per tick (fft_lines/fft_lines.c:190-208, :281) happens
// here for a tile of F frames per CTA iteration; CTAs are persistent.
// Shared memory per frame: 8 slots * M complex (+1/16 padding) + N/2+1 magnitudes.
#pragma once
#include "device_common.cuh"

namespace fftl {

struct RdifParams {
    const int16_t* pcm;
    long long channel_stride;
    long long sample_stride;    // 1 (planar) except in the HOPQ = 0 variant, which also takes interleaved PCM
    long long samples;          // samples per channel (loads beyond it read as 0)
    long long frame0, frame_emit, frame_end;
    long long nf_emit, nf_aux, aux_base;
    long long tiles_per_channel;
    long long total_tiles;
    int tic_full;               // tiles with tic < tic_full have all their samples inside the channel
    int tic_emit_lo, tic_emit_hi; // tiles with tic in [lo, hi) emit all their frames
    const float* window;        // n floats (all ones for RECT)
    const float2* tw;           // W_n^m
    const float2* twm;          // W_M^m  (sub-transform twiddles)
    const int* seg_packed;      // [nseg] work list of the segment sums, longest first (see rdif_segment_sums)
    const int* bar_seg0;        // [bars+1] first segment of the bar's bin range | number of segments << 16
    const float* bar_inv;       // 1 / (hi - lo)
    const float* sg;
    float* bars;
    int* bars_i32;
    int* aux_w;
    int* aux_bass;
    int hop, bars_n, nseg, sg_half, strict, bass_bin;
    int mag_rows;               // 256-bin rows of magnitudes that are kept (bins < 256*mag_rows, + bin N/2 when 8)
    int mag_bins;               // bins < mag_bins are read by a bar or as the bass bin
    int mag_stride;             // float2 per frame pair in the magnitude area
    int beat_bars[4];
    float zoom, circle_add;
    int no_tm;                  // host side only: keep the constants in registers (skip stft_rdif_tm.cuh)
    int* used_tm;               // host side only: set to 1 by the launcher when the tensor-memory variant was launched
};

// Real 16-point DFT: x[0..15] real -> Y[0..8] (Y[0], Y[8] real).  66 flops.
// yr/yi are indexed by k0; yi[0] and yi[8] are not written.
// WINDOWED: the inputs are x[q]*w[q]; the multiply is folded into the first butterfly stage
// (one FMUL + two FFMA per pair instead of two FMUL + two FADD).
template <bool WINDOWED>
__device__ __forceinline__ void real_dft16(const float* x, const float* w, float* yr, float* yi)
{
    float s[8], d[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        if (WINDOWED) {
            const float hi = x[q + 8] * w[q + 8];
            s[q] = fmaf(x[q], w[q], hi);
            d[q] = fmaf(x[q], w[q], -hi);
        } else {
            s[q] = x[q] + x[q + 8];
            d[q] = x[q] - x[q + 8];
        }
    }
    // even outputs: real 8-point DFT of s
    float a0 = s[0] + s[4], a1 = s[1] + s[5], a2 = s[2] + s[6], a3 = s[3] + s[7];
    float b0 = s[0] - s[4], b1 = s[1] - s[5], b2 = s[2] - s[6], b3 = s[3] - s[7];
    float e0 = a0 + a2, e1 = a1 + a3;
    yr[0] = e0 + e1;
    yr[8] = e0 - e1;
    yr[4] = a0 - a2;
    yi[4] = -(a1 - a3);
    float p = (b1 - b3) * FFTL_SQRT1_2, m = (b1 + b3) * FFTL_SQRT1_2;
    yr[2] = b0 + p;
    yi[2] = -(b2 + m);
    yr[6] = b0 - p;
    yi[6] = b2 - m;
    // odd outputs: c_q = d_q - i d_{q+4}; g_q = c_q W16^q; complex DFT-4 of g gives
    // Y[1], Y[5], conj(Y[7]), conj(Y[3])
    // (the 4-point DFT of g with W16^3 c3 folded into its butterflies: q1 +- W16^3 c3 = u2, 2 q1 - u2)
    const float2 c0 = make_float2(d[0], -d[4]), q2 = cmul_w8_1(make_float2(d[2], -d[6]));
    const float2 u0 = cadd(c0, q2), u1 = csub(c0, q2);
    const float2 q1 = cmul_w16_1(make_float2(d[1], -d[5]));
    const float2 u2 = cfma(q1, make_float2(FFTL_SIN_PI_8, -FFTL_COS_PI_8), make_float2(d[3], -d[7])), dd = ctwice_minus(q1, u2);
    const float2 g0 = cadd(u0, u2), g2 = csub(u0, u2);
    const float2 g1 = make_float2(u1.x + dd.y, u1.y - dd.x), g3 = make_float2(u1.x - dd.y, u1.y + dd.x);
    yr[1] = g0.x; yi[1] = g0.y;
    yr[5] = g1.x; yi[5] = g1.y;
    yr[7] = g2.x; yi[7] = -g2.y;
    yr[3] = g3.x; yi[3] = -g3.y;
}

template <int N, int F> struct RdifShape {
    static constexpr int M = N / 16;                 // sub-transform size and pass-A thread count
    static constexpr int THREADS = M > 512 ? 512 : M; // one thread per column; M = 1024: two columns per thread
    static constexpr int COLS = M / THREADS;
    static constexpr int SUBG = M / 16 > 32 ? 32 : M / 16; // lanes per sub-transform (a warp at most)
    static constexpr int SLOT = padded_size(M);      // float2 per slot
    static constexpr int SLOTS = 8 * F;              // per tile
    static constexpr int NB = N / 2 + 1;
    // magnitudes: one float2 array per frame pair (.x = even frame, .y = odd frame), padded
    // bin -> bin + bin/16: the 16 lanes of a sub-transform (bins 16 apart) then hit 16 distinct
    // even banks, the other half-warp (odd frame) the odd ones
    // Only the 256-bin rows a configuration reads are kept (mag_rows of them; all 8 plus bin N/2 for full band).
    static constexpr int FULL_ROWS = N / 512;        // 256-bin rows below N/2 (8 at N = 4096, 4 at N = 2048)
    __host__ __device__ static constexpr int mag_stride(int mag_rows) { return padded_size(mag_rows >= FULL_ROWS ? NB : 256 * mag_rows) + 3; }
    __host__ __device__ static size_t tail_bytes(int bars, int nseg) { return (size_t)(nseg + bars) * F * sizeof(float); }
    // per-CTA copies of the tail's tables (the L1 is tiny next to 105 KB of shared memory and is
    // streamed through by the PCM, so table loads from global would pay L2 latency in every tail phase)
    __host__ __device__ static size_t table_bytes(int bars, int nseg, int sg_half)
    {
        int m = 2 * sg_half + 1;
        return (size_t)nseg * sizeof(int) + (size_t)(bars + 1) * sizeof(int) + (size_t)bars * sizeof(float) + (size_t)m * m * sizeof(float);
    }
    static size_t smem_bytes(int bars, int nseg, int sg_half, int mag_rows)
    {
        // segment sums and raw bars alias the slot area (dead once pass B/C is done)
        size_t slots = (size_t)SLOTS * SLOT * sizeof(float2);
        size_t tail = tail_bytes(bars, nseg);
        return (slots > tail ? slots : tail) + (size_t)(F / 2) * mag_stride(mag_rows) * sizeof(float2) + F * 4 * sizeof(int) +
               table_bytes(bars, nseg, sg_half) + 16;
    }
};

// int16 sample -> float without the conversion unit (I2F runs at a quarter of the FP32 rate and a warp's 22
// conversions queue up in front of pass A): the zero-extended 16 bits, sign bit flipped, become the mantissa
// of 2^23 + (x + 32768); one exact subtraction gives x.
__device__ __forceinline__ float pcm_to_float(const int16_t* p)
{
#ifdef FFTL_I2F
    return (float)*p;
#else
    const unsigned u = *reinterpret_cast<const unsigned short*>(p);
    return __uint_as_float(u ^ 0x4B008000u) - 8421376.0f;
#endif
}

// cos / sin of q*pi/8, q = 0..15 (the Hann column of column t is 0.5 - 0.5*cos(theta_t + q*pi/8))
__device__ __forceinline__ constexpr float cos_pi8(int q)
{
    constexpr float c[16] = {1.0f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f,
                             0.0f, -0.38268343236508977173f, -0.70710678118654752440f, -0.92387953251128675613f,
                             -1.0f, -0.92387953251128675613f, -0.70710678118654752440f, -0.38268343236508977173f,
                             0.0f, 0.38268343236508977173f, 0.70710678118654752440f, 0.92387953251128675613f};
    return c[q];
}
__device__ __forceinline__ constexpr float sin_pi8(int q) { return cos_pi8((q + 12) & 15); } // sin(x) = cos(x - pi/2)

// F floats per tail item, moved as one vector
template <int F> struct FrameVec;
template <> struct FrameVec<2> {
    using type = float2;
    __device__ __forceinline__ static void load(const float* p, float* v) { float2 u = *reinterpret_cast<const float2*>(p); v[0] = u.x; v[1] = u.y; }
    __device__ __forceinline__ static void store(float* p, const float* v) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <> struct FrameVec<4> {
    using type = float4;
    __device__ __forceinline__ static void load(const float* p, float* v) { float4 u = *reinterpret_cast<const float4*>(p); v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; }
    __device__ __forceinline__ static void store(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

// read-only load that the compiler may not hoist out of the tile loop (keeps the
// per-thread constant tables out of the register file when RESIDENT is false)
__device__ __forceinline__ float ldg_pinned(const float* p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_pinned(const float2* p)
{
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// Savitzky-Golay over the raw bar vector for F frames at once; TAPS < 0: run-time width.
template <int F, int TAPS, bool PK = false>
__device__ __forceinline__ void sg_apply(const float* cf, const float* rw, int m, float* y)
{
#pragma unroll
    for (int f = 0; f < F; f++) y[f] = 0.0f;
    if (TAPS > 0) {
#pragma unroll
        for (int k = 0; k < TAPS; k++) {
            const float c = cf[k];
            float u[F];
            FrameVec<F>::load(rw + k * F, u);
            vec_fma<F, PK>(c, u, y);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < m; k++) {
            const float c = cf[k];
            float u[F];
            FrameVec<F>::load(rw + k * F, u);
            vec_fma<F, PK>(c, u, y);
        }
    }
}

// Persistent-CTA tile cursor: tile -> (channel, tile within the channel), the tile's first sample and its
// place in the outputs are all carried from tile to tile, so the tile loop has neither a 64-bit division nor
// 64-bit multiplies in front of its loads.  tiles_per_channel < 2^31 (checked by the launcher).
struct TileCursor {
    int left;           // tiles this CTA still has to do
    int tic, ch;        // tile within its channel; channel
    const int16_t* xp;  // first sample of the tile in its channel
    long long o0;       // element offset of (ch, first frame of the tile, bar 0) in the outputs; meaningless for warm-up tiles
    long long dx, dox;  // per-advance steps of xp and o0
    __device__ __forceinline__ void init(const RdifParams& p, long long first, long long stride, int F)
    {
        left = first < p.total_tiles ? (int)((p.total_tiles - first + stride - 1) / stride) : 0;
        ch = (int)(first / p.tiles_per_channel);
        tic = (int)(first - (long long)ch * p.tiles_per_channel);
        const long long f0 = p.frame0 + (long long)tic * F;
        xp = p.pcm + (long long)ch * p.channel_stride + f0 * p.hop * p.sample_stride;
        o0 = ((long long)ch * p.nf_emit + (f0 - p.frame_emit)) * p.bars_n;
        dx = stride * F * p.hop * p.sample_stride;
        dox = stride * F * p.bars_n;
    }
    __device__ __forceinline__ void advance(const RdifParams& p, int stride, int F)
    {
        left--;
        tic += stride;
        xp += dx;
        o0 += dox;
        while (tic >= (int)p.tiles_per_channel) { // next channel (more than once only when the grid outnumbers a channel's tiles)
            tic -= (int)p.tiles_per_channel;
            ch++;
            xp += p.channel_stride - p.tiles_per_channel * F * p.hop * p.sample_stride;
            o0 += (p.nf_emit - p.tiles_per_channel * F) * p.bars_n;
        }
    }
    __device__ __forceinline__ long long first_frame(const RdifParams& p, int F) const { return p.frame0 + (long long)tic * F; }
};

// Work-list entry of the segment sums: padded first bin (14 bits) | length 1..16 (5 bits) | segment number (13 bits).
// The list stays in bin order: neighbouring lanes then read neighbouring magnitudes, which is conflict-free
// (sorting it by length was measured: 5 % slower from bank conflicts).
#define FFTL_SEG_PACK(padded_lo, cnt, idx) ((padded_lo) | ((cnt) << 14) | ((idx) << 19))

template <int F, int MAGS, int K, bool PK> __device__ __forceinline__ void seg_steps(const float2* m0, int cnt, float* acc)
{
#pragma unroll
    for (int i = 0; i < K; i++) {
        if (i < cnt) {
#pragma unroll
            for (int pr = 0; pr < F / 2; pr++) {
                const float2 a = m0[pr * MAGS + i];
                const float u[2] = {a.x, a.y};
                vec_add<2, PK>(acc + 2 * pr, u);
            }
        }
    }
}

// Tail, first step: per-segment sums of the magnitudes of the tile's F frames.  Segments are at most 8 bins
// long and never cross a multiple of 16, so one is at most 8 consecutive padded positions.  Every thread of
// the calling group must take part (warp shuffles); nthreads is a multiple of 32.
template <int F, int MAGS, bool PK = false>
__device__ __forceinline__ void rdif_segment_sums(const int* t_seg, const float2* mags, float* segsum, int nseg, int t, int nthreads)
{
    for (int base = 0; base < nseg; base += nthreads) {
        const int sgi = base + t;
        const int sd = sgi < nseg ? t_seg[sgi] : 0;
        const int cnt = (sd >> 14) & 31;
        const float2* m0 = mags + (sd & 0x3fff);
        float acc[F];
#pragma unroll
        for (int f = 0; f < F; f++) acc[f] = 0.0f;
        // the low bars are 1-2 bins wide: their warps take the short form; segments longer than 8 bins exist only
        // where the host needed them to fit the work list into one pass
        const int longest = __reduce_max_sync(0xffffffffu, cnt);
        if (longest <= 2) seg_steps<F, MAGS, 2, PK>(m0, cnt, acc);
        else if (longest <= 8) seg_steps<F, MAGS, 8, PK>(m0, cnt, acc);
        else seg_steps<F, MAGS, 16, PK>(m0, cnt, acc);
        if (cnt) FrameVec<F>::store(segsum + (int)((unsigned)sd >> 19) * F, acc);
    }
}

// Row of the Savitzky-Golay matrix and first raw bar it applies to, for output bar b of B (half width w):
// interior bars use the centre row on [b-w, b+w]; the w bars at either edge use the edge rows of the
// polynomial fit over the first / last 2w+1 bars (scipy's mode="interp").
__device__ __forceinline__ void sg_row(int b, int B, int w, int& row, int& base)
{
    base = min(max(b - w, 0), B - (2 * w + 1));
    row = b - base;
}

// Tail, last step: Savitzky-Golay -> zoom -> heights of the tile's F frames (fft_lines.c:195-208), done by
// `nthreads` threads (this one is number t).  The float multiply and add are the explicit _rn forms so
// that the fast and the careful branch round alike (a frame can sit in an edge tile of one shard and in
// an interior tile of the full run).
template <int F, int SGW, bool PK = false>
__device__ __forceinline__ void rdif_emit_bars(const RdifParams& p, const float* raw, const float* t_sg, int t, int nthreads, bool interior, long long o0,
                                               long long f0)
{
    const int B = p.bars_n;
    const int w = SGW >= 0 ? SGW : p.sg_half, m = 2 * w + 1;
    if (interior && p.bars_i32 == nullptr) {
        // every frame of the tile is emitted, float heights only: nothing to predicate
#pragma unroll 1
        for (int b = t; b < B; b += nthreads) {
            float y[F];
            if (w == 0) FrameVec<F>::load(raw + b * F, y);
            else {
                int row, base;
                sg_row(b, B, w, row, base);
                sg_apply<F, (SGW > 0 ? 2 * SGW + 1 : -1), PK>(t_sg + row * m, raw + base * F, m, y);
            }
            float* ob = p.bars + (o0 + b);
            vec_scale<F, PK>(y, p.zoom);
#pragma unroll
            for (int f = 0; f < F; f++) ob[f * B] = __fadd_rn(y[f], p.circle_add);
        }
        return;
    }
    // frames of this tile that are emitted: [e_lo, e_hi)
    const long long dlo = p.frame_emit - f0, dhi = p.frame_end - f0;
    const int e_lo = dlo > 0 ? (dlo < F ? (int)dlo : F) : 0;
    const int e_hi = dhi < F ? (dhi > 0 ? (int)dhi : 0) : F;
    const int cadd = (int)p.circle_add;
#pragma unroll 1
    for (int b = t; b < B; b += nthreads) {
        float y[F];
        if (w == 0) FrameVec<F>::load(raw + b * F, y);
        else {
            int row, base;
            sg_row(b, B, w, row, base);
            sg_apply<F, (SGW > 0 ? 2 * SGW + 1 : -1), PK>(t_sg + row * m, raw + base * F, m, y);
        }
        float* ob = p.bars + (o0 + b);
        int* oi = p.bars_i32 ? p.bars_i32 + (o0 + b) : nullptr;
        vec_scale<F, PK>(y, p.zoom);
#pragma unroll
        for (int f = 0; f < F; f++) {
            if (f >= e_lo && f < e_hi) {
                ob[f * B] = __fadd_rn(y[f], p.circle_add);
                if (oi) oi[f * B] = wrap_add(trunc_i32(y[f], p.strict), cadd);
            }
        }
    }
}

// Band energy w = (h[b0]+h[b1]+h[b2]+h[b3])/4 (fft_lines.c:281) and bass = (int)|X[bass_bin]|
// (beatDetector.c:53) of the tile's F frames, by threads 0..4F-1 of one warp: thread (f, q) redoes the
// integer height of beat bar q for frame f with the arithmetic of rdif_emit_bars, a 4-lane shuffle sums them.
// mags: the tile's magnitude area ([F/2][mag_stride] float2).
template <int F, int SGW>
__device__ __forceinline__ void rdif_emit_beat(const RdifParams& p, const float* raw, const float* t_sg, const float2* mags, int mag_stride,
                                               int t, int ch, long long f0)
{
    if (p.aux_w == nullptr || t >= 4 * F) return;
    constexpr unsigned MASK = (4 * F >= 32) ? 0xffffffffu : ((1u << (4 * F)) - 1u);
    const int B = p.bars_n;
    const int w = SGW >= 0 ? SGW : p.sg_half, m = 2 * w + 1;
    const int f = t >> 2, q = t & 3;
    const int b = p.beat_bars[q];
    float y;
    if (w == 0) y = raw[b * F + f];
    else {
        int row, base;
        sg_row(b, B, w, row, base);
        const float* cf = t_sg + row * m;
        const float* rw = raw + base * F + f;
        y = 0.0f;
        if (SGW > 0) {
#pragma unroll
            for (int k = 0; k < 2 * SGW + 1; k++) y = fmaf(cf[k], rw[k * F], y);
        } else {
#pragma unroll 1
            for (int k = 0; k < m; k++) y = fmaf(cf[k], rw[k * F], y);
        }
    }
    y = __fmul_rn(y, p.zoom);
    unsigned h = (unsigned)wrap_add(trunc_i32(y, p.strict), (int)p.circle_add);
    h += __shfl_xor_sync(MASK, h, 1);
    h += __shfl_xor_sync(MASK, h, 2);
    const long long fr = f0 + f;
    if (q == 0 && fr < p.frame_end) {
        const long long o = (long long)ch * p.nf_aux + (fr - p.aux_base);
        p.aux_w[o] = (int)h / 4;
        if (mags) { // null: the caller has written the bass already (its magnitude buffer is gone by now)
            const float2 mb = mags[(f >> 1) * mag_stride + pad_index(p.bass_bin)];
            p.aux_bass[o] = trunc_i32((f & 1) ? mb.y : mb.x, p.strict);
        }
    }
}

// SGW: compile-time Savitzky-Golay half width (0 = off, 3 = the default) or -1 for any run-time width.
// ROWS: number of 256-bin magnitude rows kept (compile time, so unused rows cost nothing).
template <int N, int F, int HOPQ, bool WINDOWED, bool RESIDENT, int MINB, int SGW, int ROWS>
__global__ void __launch_bounds__(RdifShape<N, F>::THREADS, MINB) stft_bars_rdif_kernel(const RdifParams p)
{
    using Sh = RdifShape<N, F>;
    constexpr int M = Sh::M, SUBG = Sh::SUBG, SLOT = Sh::SLOT, THREADS = Sh::THREADS, COLS = Sh::COLS;
    constexpr int MAGS = Sh::mag_stride(ROWS);
    constexpr bool PK = N == 8192; // the tail's per-frame arithmetic on packed pairs (vec_add / vec_scale / vec_fma, fft_device.cuh)
    static_assert(N == 16384 || N == 8192 || N == 4096 || N == 2048 || N == 1024,
                  "N = 16 M with M = 1024 (radix 16, 16, 4), 512 (radix 16, 16, 2), 256 (radix 16, 16), 128 (radix 16, 8) or 64 (radix 16, 4)");
    static_assert(COLS == 1 || (F == 2 && HOPQ == 0 && !RESIDENT), "M = 1024: one frame pair per tile, own samples, constants reloaded per column");
    static_assert(F == 2 || F == 4, "frames are packed in pairs; the tail moves F floats as one vector");
    // HOPQ > 0: hop = HOPQ*M, the tile's frames share their samples (NX per column per tile).
    // HOPQ = 0: any hop; every frame loads its own 16 samples per column, one frame pair (32) at a time.
    constexpr int NX = HOPQ ? 16 + HOPQ * (F - 1) : 32;

    extern __shared__ float2 smem[];
    float2* slots = smem;                                           // [F/2][16][SLOT]
    const size_t tail_bytes = Sh::tail_bytes(p.bars_n, p.nseg);
    const size_t slot_bytes = (size_t)Sh::SLOTS * SLOT * sizeof(float2);
    // The tail's segment sums / raw bars alias the END of the slot area, i.e. slots of the LAST frame pair:
    // the next tile's pass A may then write the first pair's slots while slow warps still finish this tile's
    // tail; the barrier that protects the aliased part sits in the middle of pass A (LATE_ALIAS).
    const bool late_alias = (F == 4) && tail_bytes <= (size_t)16 * SLOT * sizeof(float2);
    float* segsum = reinterpret_cast<float*>(reinterpret_cast<char*>(smem) + (late_alias ? slot_bytes - tail_bytes : 0)); // [nseg][F]
    float* raw = segsum + p.nseg * F;                               // [bars][F]
    float2* mags = reinterpret_cast<float2*>(reinterpret_cast<char*>(smem) + (slot_bytes > tail_bytes ? slot_bytes : tail_bytes)); // [F/2][MAGS]
    int* beat_h = reinterpret_cast<int*>(mags + (F / 2) * MAGS);    // [F][4]
    int* t_seg = beat_h + F * 4;                                    // [nseg] FFTL_SEG_PACK work list
    int* t_bar0 = reinterpret_cast<int*>(t_seg + p.nseg);           // [bars+1]
    float* t_inv = reinterpret_cast<float*>(t_bar0 + p.bars_n + 1); // [bars]
    float* t_sg = t_inv + p.bars_n;                                 // [(2w+1)^2]
    const int t = threadIdx.x;
    const int B = p.bars_n;
    const int j = t & (SUBG - 1);  // lane within the sub-transform
    const int hw = t / SUBG;       // which slot of a frame pair this half-warp owns

    // ---- per-thread constants: window column, pass-A twiddles W_N^(t*k0), pass-C twiddles W_M^(j*r)
    float win[16];
    float2 twa[9];
    float2 twc[16];
    if (RESIDENT) {
        if (WINDOWED) {
#pragma unroll
            for (int q = 0; q < 16; q++) win[q] = __ldg(p.window + t + M * q);
        }
#pragma unroll
        for (int k0 = 1; k0 <= 8; k0++) twa[k0] = __ldg(p.tw + t * k0);
        if (M == 256) {
#pragma unroll
            for (int r = 1; r < 16; r++) twc[r] = __ldg(p.twm + j * r);
        } else if (M == 512) { // second sub-pass of three: W_512^(2 k r), k = j mod 16
#pragma unroll
            for (int r = 1; r < 16; r++) twc[r] = __ldg(p.twm + ((2 * (j & 15) * r) & (M - 1)));
        } else { // M = 128 / 64: second sub-pass is 16/R2 radix-R2 butterflies per lane (R2 = M/16), jj = j + SUBG tt: W_M^(jj r)
            constexpr int R2 = M >= 256 ? 8 : M / 16;
#pragma unroll
            for (int tt = 0; tt < 16 / R2; tt++)
#pragma unroll
                for (int r = 1; r < R2; r++) twc[R2 * tt + r] = __ldg(p.twm + (((j + SUBG * tt) * r) & (M - 1)));
        }
    }

    // the tail's tables (their loads overlap the per-thread constants above: a one-tile launch, i.e. a streaming
    // tick, waits for all of these once)
    for (int i = threadIdx.x; i < p.nseg; i += THREADS) t_seg[i] = __ldg(p.seg_packed + i);
    for (int i = threadIdx.x; i <= p.bars_n; i += THREADS) t_bar0[i] = __ldg(p.bar_seg0 + i);
    for (int i = threadIdx.x; i < p.bars_n; i += THREADS) t_inv[i] = __ldg(p.bar_inv + i);
    for (int i = threadIdx.x; i < (2 * p.sg_half + 1) * (2 * p.sg_half + 1); i += THREADS) t_sg[i] = p.sg ? __ldg(p.sg + i) : 0.0f;
    __syncthreads();

    TileCursor cur;
    cur.init(p, blockIdx.x, gridDim.x, F);
    for (; cur.left > 0; cur.advance(p, gridDim.x, F)) {
        const int16_t* xp = cur.xp + (HOPQ ? (long long)t : t * p.sample_stride);

        // ================= pass A =================
        float xr[NX];
        const bool all_samples = cur.tic < p.tic_full; // every sample the tile reads exists
        auto load_pair = [&](const int pr) {           // HOPQ = 0: samples of frames 2pr, 2pr+1 of the tile
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const long long off = (long long)(2 * pr + h) * p.hop;
                const int16_t* xf = xp + off * p.sample_stride;
                const long long qs = (long long)M * p.sample_stride; // column step in elements
                if (all_samples) {
#pragma unroll
                    for (int q = 0; q < 16; q++) xr[16 * h + q] = pcm_to_float(xf + q * qs);
                } else {
                    const long long s0 = cur.first_frame(p, F) * (long long)p.hop + off + t;
#pragma unroll
                    for (int q = 0; q < 16; q++) xr[16 * h + q] = (s0 + (long long)q * M < p.samples) ? pcm_to_float(xf + q * qs) : 0.0f;
                }
            }
        };
        if (COLS > 1) {
            // M = 1024: the samples are loaded column by column inside pass_a_cols
        } else if (HOPQ == 0) {
            load_pair(0);
        } else if (all_samples) {
#pragma unroll
            for (int r = 0; r < NX; r++) xr[r] = pcm_to_float(xp + r * M);
        } else {
            const long long s0 = cur.first_frame(p, F) * (long long)p.hop; // the tile's first sample
#pragma unroll
            for (int r = 0; r < NX; r++) xr[r] = (s0 + t + (long long)r * M < p.samples) ? pcm_to_float(xp + r * M) : 0.0f;
        }
        if (!late_alias) {
            __syncthreads(); // previous tile's tail is done with the aliased slots
        }
        if (!RESIDENT && COLS == 1) {
            if (WINDOWED) {
#pragma unroll
                for (int q = 0; q < 16; q++) win[q] = ldg_pinned(p.window + t + M * q);
            }
#pragma unroll
            for (int k0 = 1; k0 <= 8; k0++) twa[k0] = ldg_pinned(p.tw + t * k0);
        }
        // pass A of frame pair pr: the 16 slots of the pair
        auto pass_a = [&](const int pr) {
            float2* ps = slots + pr * 16 * SLOT + pad_index(t);
            float y0[2], y8[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int f = 2 * pr + h;
                float yr[9], yi[9];
                real_dft16<WINDOWED>(HOPQ ? xr + HOPQ * f : xr + 16 * h, win, yr, yi);
                y0[h] = yr[0];
                y8[h] = yr[8];
                // slots 2*k0 + h (k0 = 1..7): frame h of the pair, residue k0
#pragma unroll
                for (int k0 = 1; k0 < 8; k0++) ps[(2 * k0 + h) * SLOT] = cmul(make_float2(yr[k0], yi[k0]), twa[k0]);
            }
            // slot 0: residue 0 of both frames packed as one complex sequence;
            // slot 1: residue 8 likewise, twiddled by W_N^(8t)
            ps[0] = make_float2(y0[0], y0[1]);
            ps[SLOT] = cmul(make_float2(y8[0], y8[1]), twa[8]);
        };
        // pass B/C of frame pair pr: one slot per half-warp, warp-level syncs only
        auto wsync = [] { __syncwarp(); };
        auto pass_bc256 = [&](const int pr) {
            float2* s = slots + (pr * 16 + hw) * SLOT;
            float2 v[16];
            // first sub-pass (no twiddles): radix 16 over t = j + 16 r, result X1[16 j + k] back into the slot
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = s[pad_index(j + r * (M / 16))];
            radix16_fused<false>(v, nullptr);
            wsync();
#pragma unroll
            for (int rr = 0; rr < 16; rr++) s[pad_index(16 * j + Radix<16>::out_index(rr))] = v[rr];
            wsync();
            // second sub-pass; the W_M^(j r) multiplies are folded into its first butterflies, results stay in registers
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = s[pad_index(j + r * (M / 16))];
            if (!RESIDENT) {
#pragma unroll
                for (int r = 1; r < 16; r++) twc[r] = ldg_pinned(p.twm + j * r);
            }
            radix16_fused<true>(v, twc);
            // v[rr] = Z[m], m = j + 16*kk, kk = out_index(rr)
            float2* mg2 = mags + pr * MAGS;
            if (hw >= 2) {
                // regular slot: frame h = hw&1, residue k0 = hw>>1; all 256 outputs are used:
                // m < 128 directly (bin k0 + 16m), m >= 128 through the conjugate mirror
                const int k0 = hw >> 1;
                float* mg = reinterpret_cast<float*>(mg2) + (hw & 1);
                // pad_index(k0 + 16j + 256kk) = k0 + 17j + 272kk ; mirror: (16-k0) + 17(15-j) + 272(15-kk)
                float* lo_base = mg + 2 * (k0 + 17 * j);
                float* hi_base = mg + 2 * ((16 - k0) + 17 * (15 - j));
#pragma unroll
                for (int rr = 0; rr < 16; rr++) {
                    const int kk = Radix<16>::out_index(rr);
                    const int c = kk < 8 ? kk : 15 - kk; // 256-bin row of this output
                    if (c < ROWS) {                      // rows no bar reads are neither computed nor stored
                        float mag = sqrt_approx_ftz(fmaf(v[rr].x, v[rr].x, v[rr].y * v[rr].y));
                        if (kk < 8) lo_base[2 * 272 * c] = mag;
                        else hi_base[2 * 272 * c] = mag;
                    }
                }
            } else {
                // packed slot (hw 0: residue 0, partner m' = 256-m; hw 1: residue 8, partner m' = 255-m)
                const int src = hw ? (15 - j) : ((16 - j) & 15);
                const bool self = (hw == 0) && (j == 0);
#pragma unroll
                for (int kk = 0; kk < (ROWS < 8 ? ROWS : 8); kk++) {
                    // register holding output kk is rr with out_index(rr) == kk: rr = 4*(kk&3) + (kk>>2)
                    const int rr = 4 * (kk & 3) + (kk >> 2);
                    const int kp = 15 - kk, rp = 4 * (kp & 3) + (kp >> 2);
                    float pr_ = __shfl_sync(0xffffffffu, v[rp].x, src, 16);
                    float pi_ = __shfl_sync(0xffffffffu, v[rp].y, src, 16);
                    if (self) {
                        const int ks = (16 - kk) & 15, rs = 4 * (ks & 3) + (ks >> 2);
                        pr_ = v[rs].x;
                        pi_ = v[rs].y;
                    }
                    float ar = v[rr].x + pr_, ai = v[rr].y - pi_; // 2 * Xa
                    float br = v[rr].x - pr_, bi = v[rr].y + pi_; // 2i * Xb
                    int bin = 16 * (j + 16 * kk) + (hw ? 8 : 0);
                    const float2 mm = make_float2(0.5f * sqrt_approx_ftz(fmaf(ar, ar, ai * ai)), 0.5f * sqrt_approx_ftz(fmaf(br, br, bi * bi)));
                    mg2[pad_index(bin)] = mm;
                }
                if (self && ROWS >= 8) {
                    // m = 128 (bin N/2) is its own partner: Z[128] = Xa[N/2] + i Xb[N/2], both real
                    const int rr = 4 * (8 & 3) + (8 >> 2);
                    mg2[pad_index(N / 2)] = make_float2(fabsf(v[rr].x), fabsf(v[rr].y));
                }
            }
        };
        // M = 128 / 64 (N = 2048 / 1024): SUBG = M/16 lanes per slot, radix 16, then NT = 16/R2 radix-R2 butterflies
        // per lane with R2 = M/16 (8 or 4): butterfly jj = j + SUBG*tt reads X1[jj + 16 r] and leaves Z[jj + 16 kk]
        auto pass_bc_small = [&](const int pr) {
            constexpr int R2 = M >= 256 ? 8 : M / 16, NT = 16 / R2, HALF = R2 / 2; // (M >= 256: never called, placeholder)
            auto reg_of = [](int kk) { return R2 == 8 ? (((kk & 3) << 1) | (kk >> 2)) : kk; }; // Radix<R2>::out_index(reg_of(kk)) == kk
            float2* s = slots + (pr * 16 + hw) * SLOT;
            float2 v[16];
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = s[pad_index(j + r * SUBG)];
            radix16_fused<false>(v, nullptr);
            wsync();
#pragma unroll
            for (int rr = 0; rr < 16; rr++) s[pad_index(16 * j + Radix<16>::out_index(rr))] = v[rr];
            wsync();
#pragma unroll
            for (int tt = 0; tt < NT; tt++) {
                const int jj = j + SUBG * tt;
#pragma unroll
                for (int r = 0; r < R2; r++) v[R2 * tt + r] = s[pad_index(jj + 16 * r)];
#pragma unroll
                for (int r = 1; r < R2; r++) v[R2 * tt + r] = cmul(v[R2 * tt + r], RESIDENT ? twc[R2 * tt + r] : ldg_pinned(p.twm + ((jj * r) & (M - 1))));
                Radix<R2>::run(v + R2 * tt);
            }
            // v[R2 tt + rr] = Z[m], m = jj + 16 kk, kk = Radix<R2>::out_index(rr)
            float2* mg2 = mags + pr * MAGS;
            if (hw >= 2) {
                // regular slot: m < M/2 directly (bin k0 + 16 m), m >= M/2 through the conjugate mirror (16-k0) + 16 (M-1-m)
                const int k0 = hw >> 1;
                float* mg = reinterpret_cast<float*>(mg2) + (hw & 1);
#pragma unroll
                for (int tt = 0; tt < NT; tt++) {
                    const int jj = j + SUBG * tt;
                    float* lo_base = mg + 2 * (k0 + 17 * jj);
                    float* hi_base = mg + 2 * ((16 - k0) + 17 * (15 - jj));
#pragma unroll
                    for (int rr = 0; rr < R2; rr++) {
                        const int kk = Radix<R2>::out_index(rr);
                        const int c = kk < HALF ? kk : R2 - 1 - kk;
                        if (c < ROWS) {
                            const float2 z = v[R2 * tt + rr];
                            const float mag = sqrt_approx_ftz(fmaf(z.x, z.x, z.y * z.y));
                            if (kk < HALF) lo_base[2 * 272 * c] = mag;
                            else hi_base[2 * 272 * c] = mag;
                        }
                    }
                }
            } else {
                // packed slot (hw 0: residue 0, partner m' = M-m; hw 1: residue 8, partner m' = M-1-m): the partner of
                // (jj, kk) is (15-jj, R2-1-kk) for hw 1 and (16-jj, R2-1-kk) for hw 0, i.e. butterfly NT-1-tt of lane
                // SUBG-1-j / SUBG-j; lane 0 of hw 0 is its own partner: butterfly NT-tt for tt > 0 and (0, (R2-kk) mod R2)
                constexpr unsigned MASK = SUBG >= 16 ? 0xffffffffu : ((1u << ((2 * SUBG) & 31)) - 1u); // the lanes of hw 0 and 1
                const int src = hw ? (SUBG - 1 - j) : ((SUBG - j) & (SUBG - 1));
                const bool self = (hw == 0) && (j == 0);
#pragma unroll
                for (int tt = 0; tt < NT; tt++) {
#pragma unroll
                    for (int kk = 0; kk < (ROWS < HALF ? ROWS : HALF); kk++) {
                        const int rr = reg_of(kk), rp = reg_of(R2 - 1 - kk);
                        float pr_ = __shfl_sync(MASK, v[R2 * (NT - 1 - tt) + rp].x, src, SUBG);
                        float pi_ = __shfl_sync(MASK, v[R2 * (NT - 1 - tt) + rp].y, src, SUBG);
                        if (self) {
                            const int rs = reg_of((R2 - kk) & (R2 - 1));
                            const float2 own = tt ? v[R2 * ((NT - tt) & (NT - 1)) + rp] : v[rs];
                            pr_ = own.x;
                            pi_ = own.y;
                        }
                        const float2 z = v[R2 * tt + rr];
                        float ar = z.x + pr_, ai = z.y - pi_; // 2 * Xa
                        float br = z.x - pr_, bi = z.y + pi_; // 2i * Xb
                        const int bin = 16 * ((j + SUBG * tt) + 16 * kk) + (hw ? 8 : 0);
                        mg2[pad_index(bin)] = make_float2(0.5f * sqrt_approx_ftz(fmaf(ar, ar, ai * ai)), 0.5f * sqrt_approx_ftz(fmaf(br, br, bi * bi)));
                    }
                }
                if (self && ROWS >= HALF) {
                    // m = M/2 (bin N/2) is its own partner: Z[M/2] = Xa[N/2] + i Xb[N/2], both real; jj = 0, kk = HALF
                    const int rr = reg_of(HALF);
                    mg2[pad_index(N / 2)] = make_float2(fabsf(v[rr].x), fabsf(v[rr].y));
                }
            }
        };
        // M = 512 (N = 8192): one warp per slot, radix 16, radix 16, radix 2 (eight butterflies per lane); the last
        // pass leaves Z[m], m = j + 32 tt + 256 rr, in v[2 tt + rr].  All sixteen 256-bin rows are kept.
        auto pass_bc512 = [&](const int pr) {
            float2* s = slots + (pr * 16 + hw) * SLOT;
            float2 v[16];
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = s[pad_index(j + r * SUBG)];
            radix16_fused<false>(v, nullptr);
            wsync();
#pragma unroll
            for (int rr = 0; rr < 16; rr++) s[pad_index(16 * j + Radix<16>::out_index(rr))] = v[rr];
            wsync();
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = s[pad_index(j + r * SUBG)];
            if (!RESIDENT) {
#pragma unroll
                for (int r = 1; r < 16; r++) twc[r] = ldg_pinned(p.twm + ((2 * (j & 15) * r) & (M - 1)));
            }
            radix16_fused<true>(v, twc);
            wsync();
            {
                const int k = j & 15, j0 = (j - k) * 16 + k;
#pragma unroll
                for (int rr = 0; rr < 16; rr++) s[pad_index(j0 + 16 * Radix<16>::out_index(rr))] = v[rr];
            }
            wsync();
#pragma unroll
            for (int tt = 0; tt < 8; tt++) {
                const int jb = j + 32 * tt;
                const float2 a = s[pad_index(jb)], b = cmul(s[pad_index(jb + 256)], __ldg(p.twm + jb));
                v[2 * tt] = cadd(a, b);
                v[2 * tt + 1] = csub(a, b);
            }
            float2* mg2 = mags + pr * MAGS;
            if (hw >= 2) {
                // regular slot: rr = 0 is bin k0 + 16 m (m < 256), rr = 1 the conjugate mirror (16-k0) + 16 (511-m)
                const int k0 = hw >> 1;
                float* mg = reinterpret_cast<float*>(mg2) + (hw & 1);
#pragma unroll
                for (int tt = 0; tt < 8; tt++) {
                    const int m = j + 32 * tt;
                    const float2 z0 = v[2 * tt], z1 = v[2 * tt + 1];
                    mg[2 * (k0 + 17 * m)] = sqrt_approx_ftz(fmaf(z0.x, z0.x, z0.y * z0.y));
                    mg[2 * ((16 - k0) + 17 * (255 - m))] = sqrt_approx_ftz(fmaf(z1.x, z1.x, z1.y * z1.y));
                }
            } else {
                // packed slot (hw 0: residue 0, partner m' = 512-m; hw 1: residue 8, partner m' = 511-m); bins m < 256
                // are rr = 0, their partners are rr = 1 of butterfly 7-tt in lane 31-j (hw 1) or 32-j (hw 0, j > 0);
                // lane 0 of hw 0: butterfly 8-tt of its own for tt > 0, and m = 0 (DC) is its own partner
                const int src = hw ? (31 - j) : ((32 - j) & 31);
                const bool self = (hw == 0) && (j == 0);
#pragma unroll
                for (int tt = 0; tt < 8; tt++) {
                    float pr_ = __shfl_sync(0xffffffffu, v[2 * (7 - tt) + 1].x, src);
                    float pi_ = __shfl_sync(0xffffffffu, v[2 * (7 - tt) + 1].y, src);
                    if (self) {
                        const float2 own = tt ? v[2 * ((8 - tt) & 7) + 1] : v[0];
                        pr_ = own.x;
                        pi_ = own.y;
                    }
                    const float2 z = v[2 * tt];
                    float ar = z.x + pr_, ai = z.y - pi_; // 2 * Xa
                    float br = z.x - pr_, bi = z.y + pi_; // 2i * Xb
                    const int bin = 16 * (j + 32 * tt) + (hw ? 8 : 0);
                    mg2[pad_index(bin)] = make_float2(0.5f * sqrt_approx_ftz(fmaf(ar, ar, ai * ai)), 0.5f * sqrt_approx_ftz(fmaf(br, br, bi * bi)));
                }
                if (self) mg2[pad_index(N / 2)] = make_float2(fabsf(v[1].x), fabsf(v[1].y)); // m = 256: Xa[N/2] + i Xb[N/2], both real
            }
        };
        // M = 1024 (N = 16384): pass A of the tile's only frame pair, one of this thread's COLS columns after the other;
        // samples, window column and twiddles of a column are loaded when it is its turn
        // one column's share of pass A: x0 / x1 = the column's 16 samples of frame 0 / frame 1 of the pair
        auto pass_a_col = [&](const int col, const float* x0, const float* x1) {
            float wn[16];
            float2 tc[9];
            // one load per column: W_N^col = (cos th, -sin th); its powers by multiplication, the periodic Hann
            // column 0.5 - 0.5 cos(th + q pi/8) from the angle-sum identity
            tc[1] = __ldg(p.tw + col);
#pragma unroll
            for (int k0 = 2; k0 <= 8; k0++) tc[k0] = cmul(tc[k0 - 1], tc[1]);
            if (WINDOWED) {
#pragma unroll
                for (int q = 0; q < 16; q++) wn[q] = fmaf(-0.5f * cos_pi8(q), tc[1].x, fmaf(-0.5f * sin_pi8(q), tc[1].y, 0.5f));
            }
            float2* ps = slots + pad_index(col);
            float y0[2], y8[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float yr[9], yi[9];
                real_dft16<WINDOWED>(h ? x1 : x0, wn, yr, yi);
                y0[h] = yr[0];
                y8[h] = yr[8];
#pragma unroll
                for (int k0 = 1; k0 < 8; k0++) ps[(2 * k0 + h) * SLOT] = cmul(make_float2(yr[k0], yi[k0]), tc[k0]);
            }
            ps[0] = make_float2(y0[0], y0[1]);
            ps[SLOT] = cmul(make_float2(y8[0], y8[1]), tc[8]);
        };
        auto pass_a_cols = [&]() {
            if (COLS == 2 && p.hop * 2 == M && p.sample_stride == 1) {
                // hop = M/2 (configs[3]: 16384 / 512): frame 1 of column t is frame 0 of column t + M/2, and frame 1 of
                // column t + M/2 is frame 0 of column t one row further on: 33 samples serve both columns and both frames
                const int16_t* xc = cur.xp + t;
                float xa[17], xb[16]; // x[t + M q], q = 0..16 and x[t + M/2 + M q], q = 0..15
                if (all_samples) {
#pragma unroll
                    for (int q = 0; q < 17; q++) xa[q] = pcm_to_float(xc + q * M);
#pragma unroll
                    for (int q = 0; q < 16; q++) xb[q] = pcm_to_float(xc + M / 2 + q * M);
                } else {
                    const long long s0 = cur.first_frame(p, F) * (long long)p.hop + t;
#pragma unroll
                    for (int q = 0; q < 17; q++) xa[q] = (s0 + (long long)q * M < p.samples) ? pcm_to_float(xc + q * M) : 0.0f;
#pragma unroll
                    for (int q = 0; q < 16; q++) xb[q] = (s0 + M / 2 + (long long)q * M < p.samples) ? pcm_to_float(xc + M / 2 + q * M) : 0.0f;
                }
                pass_a_col(t, xa, xb);
                pass_a_col(t + THREADS, xb, xa + 1);
                return;
            }
#pragma unroll 1
            for (int c = 0; c < COLS; c++) {
                const int col = t + c * THREADS;
                const int16_t* xc = cur.xp + (long long)col * p.sample_stride;
                const long long qs = (long long)M * p.sample_stride;
                float xs[32];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const long long off = (long long)h * p.hop;
                    const int16_t* xf = xc + off * p.sample_stride;
                    if (all_samples) {
#pragma unroll
                        for (int q = 0; q < 16; q++) xs[16 * h + q] = pcm_to_float(xf + q * qs);
                    } else {
                        const long long s0 = cur.first_frame(p, F) * (long long)p.hop + off + col;
#pragma unroll
                        for (int q = 0; q < 16; q++) xs[16 * h + q] = (s0 + (long long)q * M < p.samples) ? pcm_to_float(xf + q * qs) : 0.0f;
                    }
                }
                pass_a_col(col, xs, xs + 16);
            }
        };
        // M = 1024: one warp per slot, 32 values per lane: radix 16 and radix 16 (two butterflies per lane each, jb = j + 32 u),
        // then radix 4 (eight butterflies per lane, jb = j + 32 tt) leaving Z[m], m = j + 32 tt + 256 rr, in v[4 tt + rr].
        // All thirty-two 256-bin rows are kept.
        auto pass_bc1024 = [&](const int pr) {
            float2* s = slots + (pr * 16 + hw) * SLOT;
            float2 v[32];
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int r = 0; r < 16; r++) v[16 * u + r] = s[pad_index((j + 32 * u) + 64 * r)];
            radix16_fused<false>(v, nullptr);
            radix16_fused<false>(v + 16, nullptr);
            wsync();
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int rr = 0; rr < 16; rr++) s[pad_index(16 * (j + 32 * u) + Radix<16>::out_index(rr))] = v[16 * u + rr];
            wsync();
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int r = 0; r < 16; r++) v[16 * u + r] = s[pad_index((j + 32 * u) + 64 * r)];
            {
                // second sub-pass: W_1024^(4 k r), k = jb mod 16 = j mod 16 for both butterflies
                float2 w2[16];
#pragma unroll
                for (int r = 1; r < 16; r++) w2[r] = __ldg(p.twm + ((4 * (j & 15) * r) & (M - 1)));
                radix16_fused<true>(v, w2);
                radix16_fused<true>(v + 16, w2);
            }
            wsync();
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int jb = j + 32 * u, k = jb & 15, j0 = (jb - k) * 16 + k;
#pragma unroll
                for (int rr = 0; rr < 16; rr++) s[pad_index(j0 + 16 * Radix<16>::out_index(rr))] = v[16 * u + rr];
            }
            wsync();
            const float2 wj = __ldg(p.twm + j); // W_1024^j; W_1024^(j + 32 tt) = wj * W_32^tt
#pragma unroll
            for (int tt = 0; tt < 8; tt++) {
                const int jb = j + 32 * tt;
                // W_32^tt = exp(-2 pi i tt / 32), tt = 0..7
                constexpr float C32[8] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                                          0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
                const float2 w1 = tt ? cmul(wj, make_float2(C32[tt], -C32[8 - tt])) : wj; // sin(2 pi tt/32) = cos(2 pi (8-tt)/32)
                const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                float2 a[4];
                a[0] = s[pad_index(jb)];
                a[1] = cmul(s[pad_index(jb + 256)], w1);
                a[2] = cmul(s[pad_index(jb + 512)], w2);
                a[3] = cmul(s[pad_index(jb + 768)], w3);
                dft4(a[0], a[1], a[2], a[3]);
#pragma unroll
                for (int rr = 0; rr < 4; rr++) v[4 * tt + rr] = a[rr];
            }
            float2* mg2 = mags + pr * MAGS;
            if (hw >= 2) {
                // regular slot: rr = 0, 1 are bins k0 + 16 m (m < 512), rr = 2, 3 the conjugate mirror (16-k0) + 16 (1023-m)
                const int k0 = hw >> 1;
                float* mg = reinterpret_cast<float*>(mg2) + (hw & 1);
#pragma unroll
                for (int tt = 0; tt < 8; tt++) {
#pragma unroll
                    for (int rr = 0; rr < 4; rr++) {
                        const int m = j + 32 * tt + 256 * rr;
                        const float2 z = v[4 * tt + rr];
                        const float mag = sqrt_approx_ftz(fmaf(z.x, z.x, z.y * z.y));
                        if (rr < 2) mg[2 * (k0 + 17 * m)] = mag;
                        else mg[2 * ((16 - k0) + 17 * (1023 - m))] = mag;
                    }
                }
            } else {
                // packed slot (hw 0: residue 0, partner m' = 1024-m; hw 1: residue 8, partner m' = 1023-m); bins m < 512 are
                // rr = 0, 1, their partners rr' = 3-rr of butterfly 7-tt in lane 31-j (hw 1) or 32-j (hw 0, j > 0);
                // lane 0 of hw 0: butterfly 8-tt of its own for tt > 0; tt = 0: m = 0 is its own partner, m = 256 pairs with 768
                const int src = hw ? (31 - j) : ((32 - j) & 31);
                const bool self = (hw == 0) && (j == 0);
#pragma unroll
                for (int tt = 0; tt < 8; tt++) {
#pragma unroll
                    for (int rr = 0; rr < 2; rr++) {
                        float pr_ = __shfl_sync(0xffffffffu, v[4 * (7 - tt) + (3 - rr)].x, src);
                        float pi_ = __shfl_sync(0xffffffffu, v[4 * (7 - tt) + (3 - rr)].y, src);
                        if (self) {
                            const float2 own = tt ? v[4 * ((8 - tt) & 7) + (3 - rr)] : (rr ? v[3] : v[0]);
                            pr_ = own.x;
                            pi_ = own.y;
                        }
                        const float2 z = v[4 * tt + rr];
                        float ar = z.x + pr_, ai = z.y - pi_; // 2 * Xa
                        float br = z.x - pr_, bi = z.y + pi_; // 2i * Xb
                        const int bin = 16 * (j + 32 * tt + 256 * rr) + (hw ? 8 : 0);
                        mg2[pad_index(bin)] = make_float2(0.5f * sqrt_approx_ftz(fmaf(ar, ar, ai * ai)), 0.5f * sqrt_approx_ftz(fmaf(br, br, bi * bi)));
                    }
                }
                if (self) mg2[pad_index(N / 2)] = make_float2(fabsf(v[2].x), fabsf(v[2].y)); // m = 512: Xa[N/2] + i Xb[N/2], both real
            }
        };
        auto pass_bc = [&](const int pr) {
            if constexpr (M == 1024) pass_bc1024(pr);
            else if constexpr (M == 512) pass_bc512(pr);
            else if constexpr (M == 256) pass_bc256(pr);
            else pass_bc_small(pr);
        };
        // Order of work and CTA barriers: pass B/C of a pair needs pass A of that pair from every thread.
        //   A(0) | A(1) | B/C(0) B/C(1) | tail      (F = 4)          A(0) | B/C(0) | tail      (F = 2)
        // The barrier after A(0) is the one that lets the previous tile's tail finish with the buffers it
        // aliases onto the last pair's slots (late_alias).  (Running B/C(0) right behind A(1), with the second
        // barrier after it, was measured: 1.3 % slower with the tail's seven magnitude rows, equal without.)
        if constexpr (COLS > 1) pass_a_cols();
        else pass_a(0);
        if (HOPQ == 0 && F == 4) load_pair(1); // in flight across the barrier
        __syncthreads();
        if (F == 4) {
            pass_a(1);
            __syncthreads();
            pass_bc(0);
            pass_bc(1);
        } else {
            pass_bc(0);
        }
        __syncthreads();
        // ================= bins -> bars, SG, outputs: every item carries the tile's F frames =================
        rdif_segment_sums<F, MAGS, PK>(t_seg, mags, segsum, p.nseg, t, THREADS);
        __syncthreads();
        for (int b = t; b < B; b += THREADS) {
            const int a = t_bar0[b] & 0xffff, e = a + (t_bar0[b] >> 16); // the bar's range: first segment, count
            float acc[F];
            FrameVec<F>::load(segsum + a * F, acc);
            for (int k = a + 1; k < e; k++) {
                float u[F];
                FrameVec<F>::load(segsum + k * F, u);
                vec_add<F, PK>(acc, u);
            }
            const float inv = t_inv[b];
            vec_scale<F, PK>(acc, inv);
            FrameVec<F>::store(raw + b * F, acc);
        }
        __syncthreads();
        // band energy / bass first (16 lanes of warp 0; the magnitudes are still in place), then the heights
        if (t < 4 * F) rdif_emit_beat<F, SGW>(p, raw, t_sg, mags, MAGS, t, cur.ch, cur.first_frame(p, F));
        rdif_emit_bars<F, SGW, PK>(p, raw, t_sg, t, THREADS, cur.tic >= p.tic_emit_lo && cur.tic < p.tic_emit_hi, cur.o0, cur.first_frame(p, F));
        // No barrier here: the next tile's pass A starts at once; the barrier that protects the aliased
        // buffers is the one inside (late_alias) or in front of (otherwise) its pass A.
    }
}

} // namespace fftl

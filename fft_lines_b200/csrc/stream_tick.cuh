// stream_tick.cuh -- ONE kernel per streaming tick, laid out for latency (BASELINE configs[2]).
//
// A tick of `streams` windows is a single wave of work (about 1 us of arithmetic at the roofline for 1024 streams of
// 1024 points), so what it costs is the length of the dependent chain, not throughput.  The three-launch form
// (stream_slide_kernel, the batched bars kernel with 16 points per thread, stream_beat_carry_kernel: stream.cuh) pays
// three launch latencies and a 16-point-per-thread transform built for throughput.  Here a CTA of N/4 threads owns two
// streams from the window move to the beat record:
//   * MoveArray (wasapi_audio.cpp:390-400, exact form) in place: every thread reads its 4 new window values per stream
//     (old window shifted by `count`, or the fresh samples), one CTA barrier, writes them back and keeps them in registers;
//   * the two real windows as one complex transform (fft_lines.c:190-194), 4 points per thread: radix-4 Stockham passes
//     through shared memory (five at N = 1024), every thread's dependent chain a quarter of the 16-point form;
//   * untangle + magnitudes (fft_lines.c:202), log bars by up to 8 lanes per bar with a shuffle reduction, Savitzky-Golay,
//     zoom, (int) heights, band energy and bass bin (fft_lines.c:195-208, :281; beatDetector.c:53);
//   * AddBeatSample / BassBeatDetector on the streams' rings with the carried tempo transform (stream_carry.cuh): the ring
//     state is fetched at the start of the kernel, the bins are updated between the untangling and the bars, the peak scan
//     runs on the last two warps (32 lanes per stream, warp reductions) beside the Savitzky-Golay step; the CTA that
//     arrives last at the device-resident tick counter advances it (the arrival itself is made at the start, so its round
//     trip is off the tail).
// Used for N = 1024 and 2048 with the exact window move; every other case keeps the three launches.
#pragma once
#include "kernels.cuh"

namespace fftl {

struct StreamTickParams {
    int16_t* window;          // [streams][n] sliding windows (device resident)
    const int16_t* fresh;     // [streams][fresh_stride] new samples of this tick (the newest `count`)
    long long fresh_stride;
    int count;                // 1..n new samples per stream
    int streams;
};

// Passes behind the first one (NS = product of the earlier radices): radix 4 while it divides, then one radix-2 pass.
// The twiddles depend on the thread only, so all of them are fetched before the first pass (tick_twiddles) and no pass
// waits on a global load: 3 per radix-4 pass, 2 for the radix-2 pass (two butterflies per thread).
template <int N, int NS> struct TickPass {
    static constexpr int R = (N / NS >= 4) ? 4 : 2, T = 4 / R, NTW = T * (R - 1);
};
template <int N, int G, int NS> __device__ __forceinline__ void tick_twiddles(float2* w, const float2* __restrict__ tw, int g)
{
    if constexpr (NS < N) {
        using P = TickPass<N, NS>;
#pragma unroll
        for (int t = 0; t < P::T; t++) {
            const int j = g + t * G, kk = j & (NS - 1), base = kk * (N / (NS * P::R));
#pragma unroll
            for (int r = 1; r < P::R; r++) w[t * (P::R - 1) + r - 1] = __ldg(&tw[base * r]);
        }
        tick_twiddles<N, G, NS * P::R>(w + P::NTW, tw, g);
    }
}
template <int N> __host__ __device__ constexpr int tick_twiddle_count(int ns = 4) { return ns < N ? (ns * 4 <= N ? 3 + tick_twiddle_count<N>(ns * 4) : 2) : 0; }

template <int N, int G, int NS, typename Sync>
__device__ __forceinline__ void tick_fft_rest(float2* v, float2* s, const float2* w, int g, Sync sync)
{
    if constexpr (NS < N) {
        using P = TickPass<N, NS>;
        constexpr int R = P::R, T = P::T;
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int j = g + t * G;
#pragma unroll
            for (int r = 0; r < R; r++) v[t * R + r] = s[pad_index(j + r * (N / R))];
#pragma unroll
            for (int r = 1; r < R; r++) v[t * R + r] = cmul(v[t * R + r], w[t * (R - 1) + r - 1]);
            Radix<R>::run(&v[t * R]);
        }
        sync(); // everyone has read the old contents
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int j = g + t * G, kk = j & (NS - 1), j0 = (j - kk) * R + kk;
#pragma unroll
            for (int rr = 0; rr < R; rr++) s[pad_index(j0 + Radix<R>::out_index(rr) * NS)] = v[t * R + rr];
        }
        sync();
        tick_fft_rest<N, G, NS * R>(v, s, w + P::NTW, g, sync);
    }
}

// 16 bytes global -> shared memory without a register in between (LDGSTS); completion: cp.async.wait_all by the same thread
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

template <int N>
__global__ void __launch_bounds__(N / 4, 4096 / N) stream_tick_kernel(const StftParams p, const StreamTickParams k, const StreamCarry q)
{
    constexpr int G = N / 4; // threads per CTA = per stream pair; 64 registers: 1024 streams of 1024 points are one wave (four CTAs per SM)
    extern __shared__ float2 smem[];
    __shared__ int s_h[2 * STREAM_CARRY_HROW];
    __shared__ int beat_h[8];
    float2* s = smem;
    const int B = p.bars_n;
    float* raw = reinterpret_cast<float*>(smem + padded_size(N)); // [2][B]
    float* t_sg = raw + 2 * B;                                    // [(2w+1)^2] Savitzky-Golay matrix
    auto sync = [] { __syncthreads(); };
    const int g = threadIdx.x;
    const int sa = 2 * blockIdx.x;
    const bool has_b = sa + 1 < k.streams;
    // ---- MoveArray, in place (this CTA is the only one that touches its two windows)
    int16_t* wa = k.window + (long long)sa * N;
    int16_t* wb = wa + N;
    const int16_t* fa = k.fresh + (long long)sa * k.fresh_stride;
    const int16_t* fb = fa + k.fresh_stride;
    const int keep = N - k.count;
    int16_t xa[4], xb[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int i = g + r * G;
        xa[r] = i < keep ? wa[i + k.count] : __ldg(fa + (i - keep));
        xb[r] = has_b ? (i < keep ? wb[i + k.count] : __ldg(fb + (i - keep))) : (int16_t)0;
    }
    // everything that depends on the thread only: twiddles, the first bar's bin range, the window column, the SG matrix
    float2 twv[tick_twiddle_count<N>()];
    tick_twiddles<N, G, 4>(twv, p.tw, g);
    const int L = B * 8 <= G ? 8 : B * 4 <= G ? 4 : B * 2 <= G ? 2 : 1; // lanes per bar in the bins -> bars step
    const int sub = g & (L - 1), per = G / L;
    int l0 = 0, h0 = 0;
    float inv0 = 0.0f;
    if (g / L < B) {
        l0 = __ldg(p.lo + g / L);
        h0 = __ldg(p.hi + g / L);
        inv0 = __ldg(p.bar_inv + g / L);
    }
    float wn[4];
#pragma unroll
    for (int r = 0; r < 4; r++) wn[r] = p.window ? __ldg(p.window + g + r * G) : 1.0f;

    // (the loop's branches end the first basic block: everything above has its parameter reads hoisted to the kernel's top)
    const int w = p.sg_half, m = 2 * w + 1;
    if (w > 0)
        for (int i = g; i < m * m; i += G) cp_async4(t_sg + i, p.sg + i);
    // ---- everything that depends on the tick, behind the loads that do not
    // The tick counter is read exactly once, through the L1 (one L2 request per SM instead of one per warp on one L2
    // line; the counter changes only once every CTA has its copy, and the next kernel starts with a clean L1).  The CTA
    // that arrives last at `done` advances it (stream_finish), and a CTA may arrive as soon as it has its copy: the
    // arrival carries a (zero) term of the loaded value, so the atomic cannot be issued before the load has returned,
    // and its round trip is off the kernel's tail.
    long long tick;
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(tick) : "l"(&q.state->tick) : "memory");
    unsigned arrived = 0;
    if (q.finish && g == 0) arrived = atomicAdd(&q.state->done, 1u + (unsigned)((unsigned long long)tick >> 63));
    // beat state of this tick, asked for now and used behind the transform: the tempo-ring slot that is overwritten, the
    // carried transform's bins of this thread and their twiddles, the beat ring's slot and sum (stream_carry.cuh)
    constexpr int LPS = G / 2; // lanes per stream in the beat stage
    const int half = g / LPS, lane = g - half * LPS;
    const int sc = sa + half;
    const bool valid = sc < k.streams;
    const int pos256 = (int)(tick % FFTL_TEMPO_RING), pos160 = (int)(tick % FFTL_BEAT_SAMPLES);
    int* ring = q.tempo_ring + (long long)sc * FFTL_TEMPO_RING;
    double2* xs = q.x + (long long)sc * STREAM_CARRY_BINS;
    // one bin per thread, the bins beyond LPS (bin 128 at N = 1024) on the first lanes: copied global -> shared memory
    // asynchronously (cp.async: no register is held across the transform), read back by the same thread
    constexpr int EX = STREAM_CARRY_BINS > LPS ? STREAM_CARRY_BINS - LPS : 0;
    static_assert(EX <= LPS, "at most two bins per thread");
    __shared__ double2 s_cx[G + 2 * EX], s_cw[G + 2 * EX];
    int ring_old = 0;
    if (valid) {
        ring_old = ring[pos256];
        if (lane < STREAM_CARRY_BINS) {
            cp_async16(s_cx + g, xs + lane);
            cp_async16(s_cw + g, q.tw256d + ((pos256 * lane) & (FFTL_TEMPO_RING - 1)));
        }
        if (lane < EX) {
            cp_async16(s_cx + G + half * EX + lane, xs + LPS + lane);
            cp_async16(s_cw + G + half * EX + lane, q.tw256d + ((pos256 * (LPS + lane)) & (FFTL_TEMPO_RING - 1)));
        }
    }
    // the peak scan of stream a runs on the CTA's last warp but one, that of stream b on the last warp (32 lanes, 5 bins
    // each); lane 0 of either writes the stream's beat record
    const int si = (g >> 5) - (G / 32 - 2);
    const int srec = sa + si;
    const bool rec_lane = si >= 0 && (g & 31) == 0 && srec < k.streams;
    int w_old = 0, w_sum = 0;
    if (rec_lane) {
        w_old = q.w_ring[(long long)srec * FFTL_BEAT_SAMPLES + pos160];
        w_sum = q.w_sum[srec];
    }
    __syncthreads();
    float2 v[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int i = g + r * G;
        wa[i] = xa[r];
        if (has_b) wb[i] = xb[r];
        float a = (float)xa[r], b = (float)xb[r];
        if (p.window) {
            a *= wn[r];
            b *= wn[r];
        }
        v[r] = make_float2(a, b); // z[i] = window_a[i] + i*window_b[i]
    }
    stockham_pass<N, 4, 1, G, false>(v, s, p.tw, g, sync);
    tick_fft_rest<N, G, 4>(v, s, twv, g, sync);

    // ---- untangle the two real spectra and take magnitudes (fft_lines.c:202)
    for (int kk = g; kk <= N / 2; kk += G) {
        const float2 z1 = s[pad_index(kk)], z2 = s[pad_index((N - kk) & (N - 1))];
        const float ar = z1.x + z2.x, ai = z1.y - z2.y; // 2 * Xa[k]
        const float br = z1.x - z2.x, bi = z1.y + z2.y; // 2i * Xb[k]
        s[pad_index(kk)] = make_float2(0.5f * sqrtf(fmaf(ar, ar, ai * ai)), 0.5f * sqrtf(fmaf(br, br, bi * bi)));
    }
    __syncthreads();

    // ---- tempo ring: bass = (int)|X[bass_bin]| (beatDetector.c:53) enters the carried transform, heightBuffer into s_h
    int bass = 0;
    asm volatile("cp.async.wait_all;" ::: "memory"); // this thread's copies (the SG matrix becomes visible to the others at the next barrier)
    if (valid) {
        const float2 mb = s[pad_index(p.bass_bin)];
        bass = trunc_i32(half == 0 ? mb.x : mb.y, p.strict);
        const double delta = (double)(float)bass - (double)(float)ring_old; // the reference transforms the ring as floats (beatDetector.c:61-65)
        int* hrow = s_h + half * STREAM_CARRY_HROW;
        if (lane < STREAM_CARRY_BINS) {
            double2 x = s_cx[g];
            stream_carry_bin(delta, s_cw[g], x, lane, hrow, q.strict);
            xs[lane] = x;
        }
        if (lane < EX) {
            double2 x = s_cx[G + half * EX + lane];
            stream_carry_bin(delta, s_cw[G + half * EX + lane], x, LPS + lane, hrow, q.strict);
            xs[LPS + lane] = x;
        }
        if (lane == 0) ring[pos256] = bass;
    }

    // ---- bins -> bars: mean of the magnitudes over [lo, hi), L lanes per bar (L = 8, 4, 2 or 1: as many as the CTA has)
    for (int b0 = 0; b0 < B; b0 += per) {
        const int b = b0 + g / L;
        int l = l0, h = h0;
        float inv = inv0;
        if (b0 > 0) { // more bars than the CTA has lanes: L = 1
            l = h = 0;
            if (b < B) {
                l = __ldg(p.lo + b);
                h = __ldg(p.hi + b);
                inv = __ldg(p.bar_inv + b);
            }
        }
        float su = 0.0f, sv = 0.0f;
        for (int kk = l + sub; kk < h; kk += L) {
            const float2 m = s[pad_index(kk)];
            su += m.x;
            sv += m.y;
        }
        for (int o = 1; o < L; o <<= 1) {
            su += __shfl_xor_sync(0xffffffffu, su, o);
            sv += __shfl_xor_sync(0xffffffffu, sv, o);
        }
        if (b < B && sub == 0) {
            raw[b] = su * inv;
            raw[B + b] = sv * inv;
        }
    }
    __syncthreads();

    // ---- SG smoothing, zoom, (int) cast, outputs (fft_lines.c:202-207); both streams of a bar by one thread
    for (int b = g; b < B; b += G) {
        float ya, yb;
        if (w == 0) {
            ya = raw[b];
            yb = raw[B + b];
        } else {
            int row, base;
            if (b < w) { row = b; base = 0; }
            else if (b >= B - w) { row = m - (B - b); base = B - m; }
            else { row = w; base = b - w; }
            ya = 0.0f;
            yb = 0.0f;
            for (int t = 0; t < m; t++) {
                const float c = t_sg[row * m + t];
                ya = fmaf(c, raw[base + t], ya);
                yb = fmaf(c, raw[B + base + t], yb);
            }
        }
        float ha = __fmul_rn(ya, p.zoom), hb = __fmul_rn(yb, p.zoom);
        int ia = trunc_i32(ha, p.strict), ib = trunc_i32(hb, p.strict);
        if (p.circle_add != 0.0f) {
            ha = __fadd_rn(ha, p.circle_add);
            hb = __fadd_rn(hb, p.circle_add);
            ia = wrap_add(ia, (int)p.circle_add);
            ib = wrap_add(ib, (int)p.circle_add);
        }
        const long long o = (long long)sa * B + b;
        p.bars[o] = ha;
        if (p.bars_i32) p.bars_i32[o] = ia;
        if (has_b) {
            p.bars[o + B] = hb;
            if (p.bars_i32) p.bars_i32[o + B] = ib;
        }
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (b == p.beat_bars[t]) {
                beat_h[t] = ia;
                beat_h[4 + t] = ib;
            }
    }
    // meanwhile on the last two warps: the peak scan of the two height rows (written before the previous barrier)
    int peak = 0, bass_rec = 0;
    if (si >= 0) {
        peak = stream_carry_peak<32>(s_h + si * STREAM_CARRY_HROW, g & 31, 0xffffffffu);
        const float2 mb = s[pad_index(p.bass_bin)];
        bass_rec = trunc_i32(si == 0 ? mb.x : mb.y, p.strict);
    }
    __syncthreads();

    // ---- band energy w (fft_lines.c:281), AddBeatSample and the beat record: lane 0 of the last two warps
    if (rec_lane) {
        const int* h4 = beat_h + 4 * si;
        const int wsum4 = wrap_add(wrap_add(h4[0], h4[1]), wrap_add(h4[2], h4[3]));
        stream_carry_record(q, srec, pos160, peak, wsum4 / 4, bass_rec, w_old, w_sum);
    }
    if (q.finish && g == 0 && arrived == gridDim.x - 1) { // every CTA has its copy of the tick: this one arrived last
        q.state->done = 0;
        q.state->tick = tick + 1;
    }
}

template <int N> static size_t stream_tick_smem(int bars, int sg_half)
{
    return padded_size(N) * sizeof(float2) + (2 * (size_t)bars + (size_t)(2 * sg_half + 1) * (2 * sg_half + 1)) * sizeof(float);
}

} // namespace fftl

// stream.cuh -- many-stream real-time front end (BASELINE configs[2]).
//
// The reference keeps ONE sliding window of the last N samples per channel
// (GetAudioBuffer + MoveArray, fft_lines/wasapi_audio.cpp:390-484) and runs one
// WM_TIMER tick on it (fft_lines.c:186-236, :281).  Here `streams` such windows
// and their beat state live on the device; one push = one tick for every stream:
//   stream_slide_kernel   MoveArray for all streams (exact, or with the reference's
//                         off-by-one, wasapi_audio.cpp:394)
//   stft_bars_kernel<N>   one frame per stream (kernels.cuh)
//   stream_beat_kernel    AddBeatSample + BassBeatDetector on per-stream rings
// All three take their positions from a device-resident tick counter (advanced by the last CTA
// of the push's last kernel), so the triple can be captured once in a CUDA graph and replayed every tick.
#pragma once
#include "kernels.cuh"

namespace fftl {

struct StreamState {
    long long tick;      // number of completed pushes
    unsigned done;       // CTAs of the push's last kernel that have finished (stream_finish)
};

// Called by every CTA of the LAST kernel of a push when it is done: the CTA that finishes last advances the
// tick.  Every CTA has read the tick before it gets here, so no separate one-thread launch is needed.
__device__ __forceinline__ void stream_finish(StreamState* st)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&st->done, 1u) == gridDim.x - 1) {
            st->done = 0;
            st->tick++;
        }
    }
}

// Vector form of stream_slide_kernel for shift, keep and the fresh rows all multiples of 8 samples (16 bytes).
__global__ void __launch_bounds__(256) stream_slide_vec_kernel(int16_t* __restrict__ window, const int16_t* __restrict__ fresh, int n, int count,
                                                               long long fresh_stride)
{
    const int s = blockIdx.x;
    uint4* w = reinterpret_cast<uint4*>(window + (long long)s * n);
    const uint4* f = reinterpret_cast<const uint4*>(fresh + (long long)s * fresh_stride);
    const int nv = n >> 3, shiftv = count >> 3, keepv = nv - shiftv;
    constexpr int PER = 8; // 256 threads * 8 vectors * 8 samples = 16384 samples max
    uint4 v[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int k = threadIdx.x + i * 256;
        if (k < nv) v[i] = (k < keepv) ? w[k + shiftv] : __ldg(f + (k - keepv));
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int k = threadIdx.x + i * 256;
        if (k < nv) w[k] = v[i];
    }
}

// window[s][0..n) <- shift by `count` (or count-1 in reference mode), append new[s][0..count)
// One CTA per stream; the window is staged through registers so the shift is safe in place.
__global__ void __launch_bounds__(256) stream_slide_kernel(int16_t* __restrict__ window, const int16_t* __restrict__ fresh, int n,
                                                           int count, long long fresh_stride, int reference_move)
{
    const int s = blockIdx.x;
    int16_t* w = window + (long long)s * n;
    const int16_t* f = fresh + (long long)s * fresh_stride;
    const int shift = reference_move ? count - 1 : count; // wasapi_audio.cpp:394 copies from k + addCount - 1
    const int keep = n - count;                            // samples that survive (count <= n)
    constexpr int PER = 64;                                // 256 threads * 64 = 16384 samples max
    int16_t v[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
        int k = threadIdx.x + i * 256;
        v[i] = (k < keep) ? w[k + shift] : (k < n ? f[k - keep] : (int16_t)0);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; i++) {
        int k = threadIdx.x + i * 256;
        if (k < n) w[k] = v[i];
    }
}

// Capture format -> int16 (wasapi_audio.cpp:398: (int16_t)(source[k] * 65536)).  The float product is truncated to
// int32 the way the reference's x86 binary does it (cvttss2si: NaN / out of range -> INT32_MIN) and the low 16 bits
// are kept, which gives the C-undefined cases (|x| >= 0.5) the value the reference's build produces.
__device__ __forceinline__ int16_t capture_f32_to_i16(float x)
{
    return (int16_t)(unsigned short)((unsigned)trunc_i32(__fmul_rn(x, 65536.0f), 1) & 0xffffu);
}

// out: [streams][out_row] int16 (out_row >= count)
__global__ void __launch_bounds__(256) stream_f32_to_i16_kernel(const float* __restrict__ fresh, int16_t* __restrict__ out, int streams,
                                                                int count, long long stream_stride, long long sample_stride, long long out_row)
{
    const long long total = (long long)streams * count;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / count;
        const int k = (int)(i - s * count);
        out[s * out_row + k] = capture_f32_to_i16(__ldg(fresh + s * stream_stride + k * sample_stride));
    }
}

struct StreamBeatParams {
    const int* aux_w;       // [streams] band energy of this tick
    const int* aux_bass;    // [streams] (int)|X[bass_bin]| of this tick
    int* w_ring;            // [streams][160]   beatValues   (beatDetector.c:14)
    int* w_sum;             // [streams]        sumOfBeatValues
    int* tempo_ring;        // [streams][256]   bassBeatBuffer (settingsFile.c:295)
    StreamState* state;
    fftl_beat* out;         // [streams]
    const float2* tw256;
    int streams;
    float n_times_dt;       // (float)N * timeDelta of this tick
    int strict;
    int finish;             // this is the push's last kernel: advance the tick when done
};

// 32 streams per CTA: streams 2g, 2g+1 share one complex 256-point transform of their tempo rings.
__global__ void __launch_bounds__(256) stream_beat_kernel(const StreamBeatParams p)
{
    __shared__ float2 s_tw[FFTL_TEMPO_RING];
    __shared__ float2 sf[16][padded_size(FFTL_TEMPO_RING)];
    const int tid = threadIdx.x;
    const long long tick = p.state->tick;
    const int pos256 = (int)(tick % FFTL_TEMPO_RING), pos160 = (int)(tick % FFTL_BEAT_SAMPLES);
    s_tw[tid] = __ldg(p.tw256 + tid);
    __syncthreads();
    {
        const int grp = tid >> 4, l = tid & 15;
        const int sa = blockIdx.x * 32 + 2 * grp, sb = sa + 1;
        const bool va = sa < p.streams, vb = sb < p.streams;
        const int na = va ? p.aux_bass[sa] : 0, nb = vb ? p.aux_bass[sb] : 0;
        int* ra = p.tempo_ring + (long long)sa * FFTL_TEMPO_RING;
        int* rb = p.tempo_ring + (long long)sb * FFTL_TEMPO_RING;
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int i = l + 16 * r;
            int a = va ? ra[i] : 0, b = vb ? rb[i] : 0;
            if (i == pos256) { // beatDetector.c:53: the new value goes in before the transform
                a = na;
                b = nb;
                if (va) ra[i] = na;
                if (vb) rb[i] = nb;
            }
            v[r] = make_float2((float)a, (float)b);
        }
        auto sync16 = [] { __syncwarp(); };
        float2* s = sf[grp];
        stockham_pass<256, 16, 1, 16, false>(v, s, s_tw, l, sync16);
#pragma unroll
        for (int r = 0; r < 16; r++) v[r] = s[pad_index(l + 16 * r)];
#pragma unroll
        for (int r = 1; r < 16; r++) v[r] = cmul(v[r], s_tw[l * r]);
        Radix<16>::run(v);
        __syncwarp();
        const int src = (16 - l) & 15;
        int* ha = reinterpret_cast<int*>(s);
        int* hb = ha + BEAT_HBS;
#pragma unroll
        for (int kk = 0; kk <= 10; kk++) {
            const int rr = 4 * (kk & 3) + (kk >> 2);
            const int kp = 15 - kk, rp = 4 * (kp & 3) + (kp >> 2);
            float pr_ = __shfl_sync(0xffffffffu, v[rp].x, src, 16);
            float pi_ = __shfl_sync(0xffffffffu, v[rp].y, src, 16);
            if (l == 0) {
                const int ks = (16 - kk) & 15, rs = 4 * (ks & 3) + (ks >> 2);
                pr_ = v[rs].x;
                pi_ = v[rs].y;
            }
            float ar = v[rr].x + pr_, ai = v[rr].y - pi_;
            float br = v[rr].x - pr_, bi = v[rr].y + pi_;
            const int k = l + 16 * kk;
            if (k <= FFTL_BEAT_SAMPLES) {
                ha[k] = trunc_i32(0.5f * __fsqrt_rn(fmaf(ar, ar, ai * ai)), p.strict); // _rn: these feed (int) casts
                hb[k] = trunc_i32(0.5f * __fsqrt_rn(fmaf(br, br, bi * bi)), p.strict);
            }
        }
    }
    __syncthreads();
    {
        // peak scan, 8 lanes per stream (same code shape as beat_post_kernel)
        const int si = tid >> 3, sub = tid & 7;
        const int s = blockIdx.x * 32 + si;
        const int* h = reinterpret_cast<const int*>(sf[si >> 1]) + (si & 1) * BEAT_HBS;
        const int i0 = 20 * sub;
        unsigned bits = 0;
        int prev = (i0 == 0) ? 0 : h[i0 - 1];
        int cur = h[i0];
#pragma unroll
        for (int q = 0; q < 20; q++) {
            int nxt = h[i0 + q + 1];
            int dold = (i0 + q == 0) ? 0 : wrap_sub(cur, prev);
            int dnew = wrap_sub(nxt, cur);
            if (dold >= 0 && dnew <= 0) bits |= 1u << q;
            prev = cur;
            cur = nxt;
        }
        int cnt = __popc(bits), incl = cnt;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, incl, o, 8);
            if (sub >= o) incl += n;
        }
        int rank = incl - cnt;
        int best_val = 0, best_idx = -1;
        int first = bits ? (i0 + __ffs(bits) - 1) : 1 << 30;
        unsigned bb = bits;
        while (bb) {
            int q = __ffs(bb) - 1;
            bb &= bb - 1;
            int i = i0 + q;
            int hv = h[i];
            if (rank < 30 && i > 30 && hv > best_val) {
                best_val = hv;
                best_idx = i;
            }
            rank++;
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            int ov = __shfl_xor_sync(0xffffffffu, best_val, o, 8);
            int oi = __shfl_xor_sync(0xffffffffu, best_idx, o, 8);
            int of = __shfl_xor_sync(0xffffffffu, first, o, 8);
            bool take = (oi >= 0) && (best_idx < 0 || ov > best_val || (ov == best_val && oi < best_idx));
            if (take) {
                best_val = ov;
                best_idx = oi;
            }
            first = of < first ? of : first;
        }
        if (sub == 0 && s < p.streams) {
            const int peak = best_idx >= 0 ? best_idx : (first < (1 << 30) ? first : 0);
            // AddBeatSample (beatDetector.c:30-44) on the stream's own ring
            const int w = p.aux_w[s];
            int* ring = p.w_ring + (long long)s * FFTL_BEAT_SAMPLES;
            int sum = wrap_sub(wrap_add(p.w_sum[s], w), ring[pos160]);
            ring[pos160] = w;
            p.w_sum[s] = sum;
            const int thr = sum / FFTL_BEAT_SAMPLES;
            const float qv = __fdiv_rn((float)w, (float)thr);
            fftl_beat rec;
            rec.w = w;
            rec.thr = thr;
            rec.beat_value = trunc_i32(__fmul_rn(qv, 128.0f), 1);
            rec.flag = (thr > 0 && w > thr) ? 1 : 0;
            rec.bass = p.aux_bass[s];
            rec.peak_index = peak;
            rec.movement_per_sec = __fmul_rn(p.n_times_dt, (float)peak);
            rec.reserved = 0;
            p.out[s] = rec;
        }
    }
    if (p.finish) stream_finish(p.state);
}

// The same stage with the tempo ring's 256-point transform CARRIED from tick to tick instead of recomputed
// (the streaming form of beat_post_kernel's update, kernels.cuh): a tick changes one slot of the ring
// (beatDetector.c:53-59), so
//     X[k] += (ring_new - ring_old) * W256^(pos k),   k = 0..128 (the ring is real: bins 129..160 mirror 127..96)
// in double with double twiddles (1e-16 per tick, no drift that matters in any run length; a reset zeroes X with the
// ring, which is its exact transform).  No FFT, no shuffles: one thread per bin, two streams per CTA of 256 threads,
// |X|^2 in double -> float -> correctly rounded square root -> (int) as in beat_ring_steps.  The dependent chain of a
// tick is tick -> old ring value -> fma -> sqrt -> scan instead of 16 ring loads -> 256-point transform -> untangle.
struct StreamCarryParams {
    StreamBeatParams b;
    double2* x;             // [streams][129] carried transform
    const double2* tw256d;  // W256^m in double
};

constexpr int STREAM_CARRY_BINS = FFTL_TEMPO_RING / 2 + 1; // 129

__global__ void __launch_bounds__(256) stream_beat_carry_kernel(const StreamCarryParams q)
{
    const StreamBeatParams& p = q.b;
    __shared__ int s_h[2][FFTL_BEAT_SAMPLES + 4]; // heightBuffer[0..160] of the CTA's two streams
    const int tid = threadIdx.x;
    const int half = tid >> 7, k = tid & 127;
    const int s = blockIdx.x * 2 + half;
    const bool valid = s < p.streams;
    const long long tick = p.state->tick;
    const int pos256 = (int)(tick % FFTL_TEMPO_RING), pos160 = (int)(tick % FFTL_BEAT_SAMPLES);
    int* ring = p.tempo_ring + (long long)s * FFTL_TEMPO_RING;
    int nv = 0;
    if (valid) {
        // the reference transforms the ring as floats (beatDetector.c:61-65)
        nv = p.aux_bass[s];
        const double delta = (double)(float)nv - (double)(float)ring[pos256];
        double2* xs = q.x + (long long)s * STREAM_CARRY_BINS;
        auto bin = [&](const int kk) {
            const double2 w = __ldg(q.tw256d + ((pos256 * kk) & (FFTL_TEMPO_RING - 1)));
            double2 x = xs[kk];
            x.x = fma(delta, w.x, x.x);
            x.y = fma(delta, w.y, x.y);
            xs[kk] = x;
            const float mag = __fsqrt_rn(f64_to_f32(fma(x.x, x.x, x.y * x.y))); // correctly rounded: feeds an (int) cast
            const int hv = trunc_i32(mag, p.strict);
            s_h[half][kk] = hv;
            if (kk >= 96 && kk < 128) s_h[half][FFTL_TEMPO_RING - kk] = hv; // |X[256-k]| = |X[k]|: bins 129..160
        };
        bin(k);
        if (k == 0) bin(128);
    }
    __syncthreads();
    if (valid && k == 0) ring[pos256] = nv; // every thread of the stream has read the old value
    if (tid < 16) {
        // peak scan, 8 lanes per stream (same code shape as beat_post_kernel)
        const int si = tid >> 3, sub = tid & 7;
        const int ss = blockIdx.x * 2 + si;
        const int* h = s_h[si];
        const int i0 = 20 * sub;
        unsigned bits = 0;
        int prev = (i0 == 0) ? 0 : h[i0 - 1];
        int cur = h[i0];
#pragma unroll
        for (int qq = 0; qq < 20; qq++) {
            int nxt = h[i0 + qq + 1];
            int dold = (i0 + qq == 0) ? 0 : wrap_sub(cur, prev);
            int dnew = wrap_sub(nxt, cur);
            if (dold >= 0 && dnew <= 0) bits |= 1u << qq;
            prev = cur;
            cur = nxt;
        }
        int cnt = __popc(bits), incl = cnt;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            int n = __shfl_up_sync(0xffffu, incl, o, 8);
            if (sub >= o) incl += n;
        }
        int rank = incl - cnt;
        int best_val = 0, best_idx = -1;
        int first = bits ? (i0 + __ffs(bits) - 1) : 1 << 30;
        unsigned bb = bits;
        while (bb) {
            int qq = __ffs(bb) - 1;
            bb &= bb - 1;
            int i = i0 + qq;
            int hv = h[i];
            if (rank < 30 && i > 30 && hv > best_val) {
                best_val = hv;
                best_idx = i;
            }
            rank++;
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            int ov = __shfl_xor_sync(0xffffu, best_val, o, 8);
            int oi = __shfl_xor_sync(0xffffu, best_idx, o, 8);
            int of = __shfl_xor_sync(0xffffu, first, o, 8);
            bool take = (oi >= 0) && (best_idx < 0 || ov > best_val || (ov == best_val && oi < best_idx));
            if (take) {
                best_val = ov;
                best_idx = oi;
            }
            first = of < first ? of : first;
        }
        if (sub == 0 && ss < p.streams) {
            const int peak = best_idx >= 0 ? best_idx : (first < (1 << 30) ? first : 0);
            // AddBeatSample (beatDetector.c:30-44) on the stream's own ring
            const int w = p.aux_w[ss];
            int* wr = p.w_ring + (long long)ss * FFTL_BEAT_SAMPLES;
            int sum = wrap_sub(wrap_add(p.w_sum[ss], w), wr[pos160]);
            wr[pos160] = w;
            p.w_sum[ss] = sum;
            const int thr = sum / FFTL_BEAT_SAMPLES;
            const float qv = __fdiv_rn((float)w, (float)thr);
            fftl_beat rec;
            rec.w = w;
            rec.thr = thr;
            rec.beat_value = trunc_i32(__fmul_rn(qv, 128.0f), 1);
            rec.flag = (thr > 0 && w > thr) ? 1 : 0;
            rec.bass = p.aux_bass[ss];
            rec.peak_index = peak;
            rec.movement_per_sec = __fmul_rn(p.n_times_dt, (float)peak);
            rec.reserved = 0;
            p.out[ss] = rec;
        }
    }
    if (p.finish) stream_finish(p.state);
}

// history ring: hist[s][tick % depth][b] = bars[s][b]
__global__ void __launch_bounds__(256) stream_history_push_kernel(float* __restrict__ hist, const float* __restrict__ bars,
                                                                  StreamState* state, int streams, int depth, int nbars)
{
    const int slot = (int)(state->tick % depth);
    const long long total = (long long)streams * nbars;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / nbars;
        const int b = (int)(i - s * nbars);
        hist[(s * depth + slot) * nbars + b] = bars[i];
    }
    stream_finish(state); // always the push's last kernel
}

// out[s][k][b], k = 0 oldest ... depth-1 newest; `tick` pushes have been made so far
__global__ void __launch_bounds__(256) stream_history_read_kernel(const float* __restrict__ hist, float* __restrict__ out,
                                                                  const StreamState* state, int streams, int depth, int nbars)
{
    const long long tick = state->tick;
    const long long total = (long long)streams * depth * nbars;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i % nbars);
        const long long sk = i / nbars;
        const int k = (int)(sk % depth);
        const long long s = sk / depth;
        const long long age = depth - 1 - k;         // 0 = newest frame
        const long long fr = tick - 1 - age;         // push index of that frame
        out[i] = fr >= 0 ? hist[(s * depth + (fr % depth)) * nbars + b] : 0.0f;
    }
}

// fft_lines.c:239-247: waveBar[i].height = audioBufferLeft[i] * 0.02f + offset  (float -> int assignment truncates)
__global__ void __launch_bounds__(256) stream_waveform_kernel(const int16_t* __restrict__ window, int n, int points, float scale,
                                                              float offset, int* __restrict__ heights, int streams)
{
    const long long total = (long long)streams * points;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / points;
        const int k = (int)(i - s * points);
        heights[i] = trunc_i32(__fadd_rn(__fmul_rn((float)window[s * n + k], scale), offset), 1);
    }
}

// advances the tick once per push, after every block of stream_beat_kernel has read it

} // namespace fftl

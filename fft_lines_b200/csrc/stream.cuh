// stream.cuh -- many-stream real-time front end (BASELINE configs[2]).
//
// The reference keeps ONE sliding window of the last N samples per channel
// (GetAudioBuffer + MoveArray, fft_lines/wasapi_audio.cpp:390-484) and runs one
// WM_TIMER tick on it (fft_lines.c:186-236, :281).  Here `streams` such windows
// and their beat state live on the device; one push = one tick for every stream.  For n = 1024 / 2048 with the exact
// window move a push is ONE kernel laid out for latency (stream_tick.cuh); this file holds the three-launch form
// every other case runs (and the one-kernel tick is tested against):
//   stream_slide_kernel   MoveArray for all streams (exact, or with the reference's
//                         off-by-one, wasapi_audio.cpp:394)
//   bars kernel           one frame per stream (the fast path for n = 1024 ... 16384, else stft_bars_kernel<N>)
//   stream_beat_carry_kernel  AddBeatSample + BassBeatDetector on per-stream rings
// All three take their positions from a device-resident tick counter (advanced by the last CTA
// of the push's last kernel), so the triple can be captured once in a CUDA graph and replayed every tick.
#pragma once
#include "kernels.cuh"

namespace fftl {

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

// AddBeatSample + BassBeatDetector on per-stream rings with the tempo ring's 256-point transform CARRIED from tick to
// tick instead of recomputed (stream_carry.cuh): no FFT, no shuffles, one thread per bin, two streams per CTA of 256
// threads.  The dependent chain of a tick is tick -> old ring value -> fma -> sqrt -> scan (a 256-point transform of
// two streams' rings by 16 lanes every tick, the round-1 form, cost 3.3 us more per tick).
__global__ void __launch_bounds__(256) stream_beat_carry_kernel(const StreamCarry q)
{
    __shared__ int s_h[2 * STREAM_CARRY_HROW];
    const long long tick = q.state->tick;
    stream_carry_pair<256>(q, tick, blockIdx.x * 2, s_h, threadIdx.x);
    if (q.finish) stream_finish(q.state);
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


} // namespace fftl

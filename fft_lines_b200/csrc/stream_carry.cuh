// stream_carry.cuh -- the beat stage of ONE streaming tick for a pair of streams (stream_beat_carry_kernel, stream.cuh).
//
// The 256-point transform of a stream's tempo ring (beatDetector.c:61-67) is carried from tick to tick instead of
// recomputed: a tick changes one slot of the ring (beatDetector.c:53-59), so
//     X[k] += (ring_new - ring_old) * W256^(pos k),   k = 0..128 (the ring is real: bins 129..160 mirror 127..96)
// in double with double twiddles (1e-16 per tick; a reset zeroes X with the ring, which is its exact transform),
// then heightBuffer = (int)|X| (beatDetector.c:69-75), the peak scan (:78-103) and AddBeatSample (:30-44).
#pragma once
#include "device_common.cuh"

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

__device__ __forceinline__ float f64_to_f32(double x)
{
    float r; // plain round-to-nearest: --use_fast_math would turn a C cast into a flush-to-zero sequence
    asm("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(x));
    return r;
}

// One bin of the carried transform: X[k] += delta * W256^(pos k), heightBuffer[k] = (int)|X[k]| (beatDetector.c:69-75)
__device__ __forceinline__ void stream_carry_bin(const double delta, const double2 w, double2& x, const int kk, int* hrow, const int strict)
{
    x.x = fma(delta, w.x, x.x);
    x.y = fma(delta, w.y, x.y);
    const float mag = __fsqrt_rn(f64_to_f32(fma(x.x, x.x, x.y * x.y))); // correctly rounded: feeds an (int) cast
    const int hv = trunc_i32(mag, strict);
    hrow[kk] = hv;
    if (kk >= FFTL_TEMPO_RING - FFTL_BEAT_SAMPLES && kk < FFTL_TEMPO_RING / 2) hrow[FFTL_TEMPO_RING - kk] = hv; // |X[256-k]| = |X[k]|: bins 129..160
}

struct StreamCarry {
    const int* aux_w;       // [streams] band energy of this tick
    const int* aux_bass;    // [streams] (int)|X[bass_bin]| of this tick
    int* w_ring;            // [streams][160]   beatValues   (beatDetector.c:14)
    int* w_sum;             // [streams]        sumOfBeatValues
    int* tempo_ring;        // [streams][256]   bassBeatBuffer (settingsFile.c:295)
    double2* x;             // [streams][129]   carried transform of the tempo ring
    const double2* tw256d;  // W256^m in double
    StreamState* state;
    fftl_beat* out;         // [streams]
    int streams;
    float n_times_dt;       // (float)N * timeDelta of this tick
    int strict;
    int finish;             // the calling kernel is the push's last one: advance the tick when done
};

constexpr int STREAM_CARRY_BINS = FFTL_TEMPO_RING / 2 + 1; // 129
constexpr int STREAM_CARRY_HROW = FFTL_BEAT_SAMPLES + 4;   // heightBuffer[0..160] of one stream

// Peak scan of one stream's height row h[0..160] (beatDetector.c:78-103) by LANES (8, 16 or 32) consecutive lanes of one
// warp, 160 / LANES bins each (sub = 0..LANES-1; mask = those lanes, or all the lanes of the warp that make this call
// together): returns the tempo peak, valid in the lane with sub = 0.
template <int LANES> __device__ __forceinline__ int stream_carry_peak(const int* h, const int sub, const unsigned mask)
{
    // same code shape as beat_post_kernel
    constexpr int PER = FFTL_BEAT_SAMPLES / LANES;
    static_assert(PER * LANES == FFTL_BEAT_SAMPLES && LANES <= 32, "160 bins split evenly");
    const int i0 = PER * sub;
    unsigned bits = 0;
    int prev = (i0 == 0) ? 0 : h[i0 - 1];
    int cur = h[i0];
#pragma unroll
    for (int qq = 0; qq < PER; qq++) {
        int nxt = h[i0 + qq + 1];
        int dold = (i0 + qq == 0) ? 0 : wrap_sub(cur, prev);
        int dnew = wrap_sub(nxt, cur);
        if (dold >= 0 && dnew <= 0) bits |= 1u << qq;
        prev = cur;
        cur = nxt;
    }
    int cnt = __popc(bits), incl = cnt;
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) {
        int n = __shfl_up_sync(mask, incl, o, LANES);
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
    if constexpr (LANES == 32) {
        // a whole warp per stream: three warp reductions instead of fifteen shuffles (largest value among the lanes that have
        // a candidate -- a candidate's value is > 0 --, lowest index on a tie; the earliest peak of all as the fall-back)
        const int m = __reduce_max_sync(mask, best_idx >= 0 ? best_val : 0);
        const int idx = __reduce_min_sync(mask, (best_idx >= 0 && best_val == m) ? best_idx : (1 << 30));
        first = __reduce_min_sync(mask, first);
        return m > 0 ? idx : (first < (1 << 30) ? first : 0);
    } else {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
            int ov = __shfl_xor_sync(mask, best_val, o, LANES);
            int oi = __shfl_xor_sync(mask, best_idx, o, LANES);
            int of = __shfl_xor_sync(mask, first, o, LANES);
            bool take = (oi >= 0) && (best_idx < 0 || ov > best_val || (ov == best_val && oi < best_idx));
            if (take) {
                best_val = ov;
                best_idx = oi;
            }
            first = of < first ? of : first;
        }
        return best_idx >= 0 ? best_idx : (first < (1 << 30) ? first : 0);
    }
}

// AddBeatSample (beatDetector.c:30-44) on stream ss's own ring and the tick's beat record: w / bass = this tick's band energy
// and bass value, w_old = the slot of the beat ring that is overwritten, w_sum = the ring's running sum.  One thread.
__device__ __forceinline__ void stream_carry_record(const StreamCarry& q, const int ss, const int pos160, const int peak, const int w, const int bass,
                                                    const int w_old, const int w_sum)
{
    const int sum = wrap_sub(wrap_add(w_sum, w), w_old);
    q.w_ring[(long long)ss * FFTL_BEAT_SAMPLES + pos160] = w;
    q.w_sum[ss] = sum;
    const int thr = sum / FFTL_BEAT_SAMPLES;
    const float qv = __fdiv_rn((float)w, (float)thr);
    fftl_beat rec;
    rec.w = w;
    rec.thr = thr;
    rec.beat_value = trunc_i32(__fmul_rn(qv, 128.0f), 1);
    rec.flag = (thr > 0 && w > thr) ? 1 : 0;
    rec.bass = bass;
    rec.peak_index = peak;
    rec.movement_per_sec = __fmul_rn(q.n_times_dt, (float)peak);
    rec.reserved = 0;
    q.out[ss] = rec;
}

// Streams s0 and s0 + 1 by THREADS threads (a multiple of 64; tid = 0..THREADS-1): THREADS/2 threads per stream, one or
// more bins each (the function was also run with 64 threads as an epilogue of the bars kernel: DESIGN.md section 4.4).  s_h: [2][STREAM_CARRY_HROW] ints of shared memory.  Contains two CTA barriers: every thread of the
// CTA must call it.
template <int THREADS>
__device__ __forceinline__ void stream_carry_pair(const StreamCarry& q, const long long tick, const int s0, int* s_h, const int tid)
{
    constexpr int LPS = THREADS / 2; // lanes per stream
    const int half = tid / LPS, lane = tid - half * LPS;
    const int s = s0 + half;
    const bool valid = s < q.streams;
    const int pos256 = (int)(tick % FFTL_TEMPO_RING), pos160 = (int)(tick % FFTL_BEAT_SAMPLES);
    int* ring = q.tempo_ring + (long long)s * FFTL_TEMPO_RING;
    int* hrow = s_h + half * STREAM_CARRY_HROW;
    int nv = 0;
    if (valid) {
        // the reference transforms the ring as floats (beatDetector.c:61-65)
        nv = q.aux_bass[s];
        const double delta = (double)(float)nv - (double)(float)ring[pos256];
        double2* xs = q.x + (long long)s * STREAM_CARRY_BINS;
        for (int kk = lane; kk < STREAM_CARRY_BINS; kk += LPS) {
            const double2 w = __ldg(q.tw256d + ((pos256 * kk) & (FFTL_TEMPO_RING - 1)));
            double2 x = xs[kk];
            stream_carry_bin(delta, w, x, kk, hrow, q.strict);
            xs[kk] = x;
        }
    }
    __syncthreads();
    if (valid && lane == 0) ring[pos256] = nv; // every thread of the stream has read the old value
    if (tid < 16) {
        const int peak = stream_carry_peak<8>(s_h + (tid >> 3) * STREAM_CARRY_HROW, tid & 7, 0xffffu << ((threadIdx.x - (unsigned)tid) & 31u));
        const int ss = s0 + (tid >> 3);
        if ((tid & 7) == 0 && ss < q.streams)
            stream_carry_record(q, ss, pos160, peak, q.aux_w[ss], q.aux_bass[ss], q.w_ring[(long long)ss * FFTL_BEAT_SAMPLES + pos160], q.w_sum[ss]);
    }
    __syncthreads(); // s_h may be reused by the caller
}

} // namespace fftl

// stft_rdif3.cuh -- the fast path at three CTAs per SM.
//
// Same transform as stft_rdif.cuh (real-input DIF, N = 16*256, 8 complex 256-point
// sub-transforms per frame, frames f/f^1 packed for residues 0 and 8), re-budgeted for
// occupancy: ncu showed the 2-CTA version issue-bound by warp latency (6.8 cycles per
// issued instruction with 4 warps per scheduler).  Here a CTA needs <= 80 registers and
// ~74 KB of shared memory, so 24 warps per SM are resident:
//   * one frame PAIR at a time goes through a single set of 16 slots (pass A -> barrier ->
//     pass B/C), twice per tile; magnitudes of the tile's 4 frames are kept for the tail;
//   * nothing per-thread is register resident across the loop: the Hann column is
//     rebuilt from (cos, sin)(2*pi*t/N) with two FMAs per tap, the pass-A twiddles
//     W_N^(t*k0) are running powers of W_N^t, the pass-C twiddles come from a 2 KB
//     shared table.
#pragma once
#include "stft_rdif.cuh"

namespace fftl {

struct Rdif3Shape {
    static constexpr int N = 4096, M = 256, SUBG = 16, F = 4;
    static constexpr int SLOT = padded_size(M);
    static constexpr int NB = N / 2 + 1;
    static constexpr int MAGS = padded_size(NB) + 3; // float2 per frame pair
    __host__ __device__ static size_t slot_bytes() { return (size_t)16 * SLOT * sizeof(float2); }
    __host__ __device__ static size_t tail_bytes(int bars, int nseg) { return (size_t)(nseg + bars) * F * sizeof(float); }
    __host__ __device__ static size_t table_bytes(int bars, int nseg, int sg_half)
    {
        int m = 2 * sg_half + 1;
        return (size_t)nseg * sizeof(int2) + (size_t)(bars + 1) * sizeof(int) + (size_t)bars * sizeof(float) + (size_t)m * m * sizeof(float);
    }
    static size_t smem_bytes(int bars, int nseg, int sg_half)
    {
        size_t s = slot_bytes(), t = tail_bytes(bars, nseg);
        return (s > t ? s : t) + (size_t)2 * MAGS * sizeof(float2) + (size_t)F * 4 * sizeof(int) + (size_t)15 * 16 * sizeof(float2) +
               table_bytes(bars, nseg, sg_half) + 16;
    }
};

// cos / sin of q*pi/8, q = 0..15 (the Hann column of thread t is 0.5 - 0.5*cos(theta_t + q*pi/8))
__device__ __forceinline__ constexpr float cos_pi8(int q)
{
    constexpr float c[16] = {1.0f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f,
                             0.0f, -0.38268343236508977173f, -0.70710678118654752440f, -0.92387953251128675613f,
                             -1.0f, -0.92387953251128675613f, -0.70710678118654752440f, -0.38268343236508977173f,
                             0.0f, 0.38268343236508977173f, 0.70710678118654752440f, 0.92387953251128675613f};
    return c[q];
}
__device__ __forceinline__ constexpr float sin_pi8(int q) { return cos_pi8((q + 12) & 15); } // sin(x) = cos(x - pi/2)

template <int HOPQ, bool WINDOWED, int SGW>
__global__ void __launch_bounds__(256, 3) stft_bars_rdif3_kernel(const RdifParams p)
{
    using Sh = Rdif3Shape;
    constexpr int N = Sh::N, M = Sh::M, SUBG = Sh::SUBG, SLOT = Sh::SLOT, MAGS = Sh::MAGS, F = Sh::F;
    constexpr int NX = 16 + HOPQ; // samples per column for one frame pair

    extern __shared__ float2 smem[];
    float2* slots = smem;                                                // [16][SLOT]   one frame pair
    float* segsum = reinterpret_cast<float*>(smem);                      // [nseg][F]    (aliases slots)
    float* raw = segsum + p.nseg * F;                                    // [bars][F]
    const size_t tail_bytes = Sh::tail_bytes(p.bars_n, p.nseg);
    const size_t slot_bytes = Sh::slot_bytes();
    float2* mags = reinterpret_cast<float2*>(reinterpret_cast<char*>(smem) + (slot_bytes > tail_bytes ? slot_bytes : tail_bytes)); // [2][MAGS]
    int* beat_h = reinterpret_cast<int*>(mags + 2 * MAGS);              // [F][4]
    float2* t_twc = reinterpret_cast<float2*>(beat_h + F * 4);          // [15][16]  W_256^(j*r), r = 1..15
    int2* t_seg = reinterpret_cast<int2*>(t_twc + 15 * 16);             // [nseg] (padded first bin, count)
    int* t_bar0 = reinterpret_cast<int*>(t_seg + p.nseg);               // [bars+1]
    float* t_inv = reinterpret_cast<float*>(t_bar0 + p.bars_n + 1);     // [bars]
    float* t_sg = t_inv + p.bars_n;                                     // [(2w+1)^2]

    const int t = threadIdx.x;
    const int B = p.bars_n;
    const int j = t & (SUBG - 1); // lane within the sub-transform
    const int hw = t / SUBG;      // which slot of the frame pair this half-warp owns

    for (int i = t; i < 15 * 16; i += M) t_twc[i] = __ldg(p.twm + (i & 15) * ((i >> 4) + 1));
    for (int i = t; i < p.nseg; i += M) {
        const int l = __ldg(p.seg_lo + i);
        t_seg[i] = make_int2(pad_index(l), __ldg(p.seg_hi + i) - l);
    }
    for (int i = t; i <= B; i += M) t_bar0[i] = __ldg(p.bar_seg0 + i);
    for (int i = t; i < B; i += M) t_inv[i] = __ldg(p.bar_inv + i);
    for (int i = t; i < (2 * p.sg_half + 1) * (2 * p.sg_half + 1); i += M) t_sg[i] = p.sg ? __ldg(p.sg + i) : 0.0f;

    const float2 w1 = __ldg(p.tw + t); // W_N^t = (cos theta_t, -sin theta_t)

    long long prev_f0 = -1;
    int prev_ch = 0;
    auto flush_aux = [&]() {
        if (t < F && p.aux_w && prev_f0 >= 0) {
            const long long fr = prev_f0 + t;
            if (fr < p.frame_end) {
                const int* h4 = beat_h + 4 * t;
                int sum = wrap_add(wrap_add(h4[0], h4[1]), wrap_add(h4[2], h4[3]));
                const float2 mb = mags[(t >> 1) * MAGS + pad_index(p.bass_bin)];
                long long o = (long long)prev_ch * p.nf_aux + (fr - p.aux_base);
                p.aux_w[o] = sum / 4;                                       // fft_lines.c:281
                p.aux_bass[o] = trunc_i32((t & 1) ? mb.y : mb.x, p.strict); // beatDetector.c:53
            }
        }
    };

    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int ch = (int)(tile / p.tiles_per_channel);
        const long long f0 = p.frame0 + (tile % p.tiles_per_channel) * F; // first frame of the tile (even)

#pragma unroll 1
        for (int pr = 0; pr < 2; pr++) {
            // ================= pass A for frames f0+2pr, f0+2pr+1 =================
            const long long s0 = (f0 + 2 * pr) * (long long)p.hop;
            const int16_t* xp = p.pcm + (long long)ch * p.channel_stride + s0 + t;
            float xr[NX];
            if (s0 + (long long)NX * M <= p.samples) {
#pragma unroll
                for (int r = 0; r < NX; r++) xr[r] = (float)xp[r * M];
            } else {
#pragma unroll
                for (int r = 0; r < NX; r++) xr[r] = (s0 + t + (long long)r * M < p.samples) ? (float)xp[r * M] : 0.0f;
            }
            // everyone is done with the slots (previous pair's pass B/C, or the previous tile's tail, which
            // aliases them) and the previous tile's beat_h is complete
            __syncthreads();
            if (pr == 0) flush_aux();
            float win[16];
            if (WINDOWED) {
                // periodic Hann column: 0.5 - 0.5*cos(theta_t + q*pi/8), cos/sin(theta_t) = (w1.x, -w1.y)
#pragma unroll
                for (int q = 0; q < 16; q++) win[q] = fmaf(-0.5f * cos_pi8(q), w1.x, fmaf(-0.5f * sin_pi8(q), w1.y, 0.5f));
            }
            {
                float2* ps = slots + pad_index(t);
                float y0[2], y8[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    float yr[9], yi[9];
                    real_dft16<WINDOWED>(xr + HOPQ * h, win, yr, yi);
                    y0[h] = yr[0];
                    y8[h] = yr[8];
                    float2 tw = w1; // running power W_N^(t*k0)
#pragma unroll
                    for (int k0 = 1; k0 < 8; k0++) {
                        ps[(2 * k0 + h) * SLOT] = cmul(make_float2(yr[k0], yi[k0]), tw);
                        tw = cmul(tw, w1);
                    }
                    if (h == 1) ps[SLOT] = cmul(make_float2(y8[0], y8[1]), tw); // tw = W_N^(8t) here
                }
                ps[0] = make_float2(y0[0], y0[1]);
            }
            __syncthreads();

            // ================= pass B/C: one slot per half-warp, warp-level syncs only =================
            {
                float2* s = slots + hw * SLOT;
                auto wsync = [] { __syncwarp(); };
                float2 v[16];
                stockham_pass<M, 16, 1, SUBG, true>(v, s, p.twm, j, wsync);
#pragma unroll
                for (int r = 0; r < 16; r++) v[r] = s[pad_index(j + r * (M / 16))];
#pragma unroll
                for (int r = 1; r < 16; r++) v[r] = cmul(v[r], t_twc[(r - 1) * 16 + j]);
                Radix<16>::run(v);
                float2* mg2 = mags + pr * MAGS;
                if (hw >= 2) {
                    const int k0 = hw >> 1;
                    float* mg = reinterpret_cast<float*>(mg2) + (hw & 1);
                    float* lo_base = mg + 2 * (k0 + 17 * j);
                    float* hi_base = mg + 2 * ((16 - k0) + 17 * (15 - j));
#pragma unroll
                    for (int rr = 0; rr < 16; rr++) {
                        const int kk = Radix<16>::out_index(rr);
                        float mag = sqrtf(fmaf(v[rr].x, v[rr].x, v[rr].y * v[rr].y));
                        if (kk < 8) lo_base[2 * 272 * kk] = mag;
                        else hi_base[2 * 272 * (15 - kk)] = mag;
                    }
                } else {
                    const int src = hw ? (15 - j) : ((16 - j) & 15);
                    const bool self = (hw == 0) && (j == 0);
#pragma unroll
                    for (int kk = 0; kk < 8; kk++) {
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
                        mg2[pad_index(bin)] = make_float2(0.5f * sqrtf(fmaf(ar, ar, ai * ai)), 0.5f * sqrtf(fmaf(br, br, bi * bi)));
                    }
                    if (self) {
                        const int rr = 4 * (8 & 3) + (8 >> 2);
                        mg2[pad_index(N / 2)] = make_float2(fabsf(v[rr].x), fabsf(v[rr].y));
                    }
                }
            }
        }
        __syncthreads();

        // ================= bins -> bars, SG, outputs: every item carries the tile's 4 frames =================
        for (int sgi = t; sgi < p.nseg; sgi += M) {
            const int2 sd = t_seg[sgi];
            const int cnt = sd.y;
            const float2* m0 = mags + sd.x;
            float acc[F];
#pragma unroll
            for (int f = 0; f < F; f++) acc[f] = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < cnt) {
#pragma unroll
                    for (int q = 0; q < F / 2; q++) {
                        const float2 a = m0[q * MAGS + i];
                        acc[2 * q] += a.x;
                        acc[2 * q + 1] += a.y;
                    }
                }
            }
            FrameVec<F>::store(segsum + sgi * F, acc);
        }
        __syncthreads();
        for (int b = t; b < B; b += M) {
            const int a = t_bar0[b], e = t_bar0[b + 1];
            float acc[F];
            FrameVec<F>::load(segsum + a * F, acc);
            for (int k = a + 1; k < e; k++) {
                float u[F];
                FrameVec<F>::load(segsum + k * F, u);
#pragma unroll
                for (int f = 0; f < F; f++) acc[f] += u[f];
            }
            const float inv = t_inv[b];
#pragma unroll
            for (int f = 0; f < F; f++) acc[f] *= inv;
            FrameVec<F>::store(raw + b * F, acc);
        }
        __syncthreads();
        {
            const int w = SGW >= 0 ? SGW : p.sg_half, m = 2 * w + 1;
            const long long dlo = p.frame_emit - f0, dhi = p.frame_end - f0;
            const int e_lo = dlo > 0 ? (dlo < F ? (int)dlo : F) : 0;
            const int e_hi = dhi < F ? (dhi > 0 ? (int)dhi : 0) : F;
            const long long o0 = ((long long)ch * p.nf_emit + (f0 - p.frame_emit)) * B;
            const bool want_int = (p.bars_i32 != nullptr) || (p.aux_w != nullptr);
#pragma unroll 1
            for (int b = t; b < B; b += M) {
                float y[F];
                if (w == 0) FrameVec<F>::load(raw + b * F, y);
                else {
                    int row, base;
                    if (b < w) { row = b; base = 0; }
                    else if (b >= B - w) { row = m - (B - b); base = B - m; }
                    else { row = w; base = b - w; }
                    sg_apply<F, (SGW > 0 ? 2 * SGW + 1 : -1)>(t_sg + row * m, raw + base * F, m, y);
                }
                float* ob = p.bars + (o0 + b);
#pragma unroll
                for (int f = 0; f < F; f++) {
                    y[f] *= p.zoom;
                    if (f >= e_lo && f < e_hi) ob[f * B] = y[f] + p.circle_add;
                }
                if (want_int) {
                    const bool isbeat = (b == p.beat_bars[0]) | (b == p.beat_bars[1]) | (b == p.beat_bars[2]) | (b == p.beat_bars[3]);
                    if (p.bars_i32 || isbeat) {
                        int* oi = p.bars_i32 ? p.bars_i32 + (o0 + b) : nullptr;
                        const int cadd = (int)p.circle_add;
#pragma unroll
                        for (int f = 0; f < F; f++) {
                            const int ih = wrap_add(trunc_i32(y[f], p.strict), cadd);
                            if (oi && f >= e_lo && f < e_hi) oi[f * B] = ih;
                            if (isbeat) {
#pragma unroll
                                for (int q = 0; q < 4; q++)
                                    if (b == p.beat_bars[q]) beat_h[f * 4 + q] = ih;
                            }
                        }
                    }
                }
            }
        }
        // no barrier here: the next pass A's leading barrier protects the aliased buffers and publishes beat_h
        prev_f0 = f0;
        prev_ch = ch;
    }
    __syncthreads();
    flush_aux();
}

} // namespace fftl

// stft_rdif_ws2.cuh -- fast path with the tail on its own warps (one persistent CTA per SM).
//
// ncu on stft_rdif.cuh: pass B/C issues at ~80 % of peak, pass A at ~60 %, the tail
// (segments -> bars -> SG -> outputs: 18 % of the instructions, three CTA barriers, few items
// per thread) at ~30 %, and while a CTA is in its tail only the other CTA's 8 warps feed the SM.
// Here the tail is a separate ROLE so the transform warps never execute it:
//
//   warps  0..7   transform group 0 : pass A -> pass B/C on its own tiles (its own 32 slots)
//   warps  8..15  transform group 1 : the same, on the interleaved tiles
//   warps 16..23  tail group        : magnitudes of either group -> bars -> outputs
//
// A transform group is the old 256-thread CTA without its tail: it syncs internally with a named
// barrier (id 2+g) and hands its magnitude buffer to the tail group through mbarriers
// (mags_full[g]: 256 arrivals from the group; mags_empty[g]: 256 arrivals from the tail warps
// once the segment sums have read it).  Registers are re-partitioned with setmaxnreg:
// transform warps 96, tail warps 40 (768 threads launched at 80: the CTA pool is 768*80 registers).  Nothing per-thread is register resident
// across tiles: the Hann column and the pass-A twiddles are rebuilt per tile from W_N^t, the
// pass-C twiddles are reloaded from a 2 KB shared table.
#pragma once
#include "stft_rdif.cuh"

namespace fftl {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    const unsigned a = smem_u32(bar);
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

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


template <int ROWS> struct RdifWs2Shape {
    static constexpr int N = 4096, M = 256, SUBG = 16, F = 4;
    static constexpr int G_THREADS = 256, T_THREADS = 256, THREADS = 2 * G_THREADS + T_THREADS;
    static constexpr int SLOT = padded_size(M);
    static constexpr int MAGS = RdifShape<N, F>::mag_stride(ROWS); // float2 per frame pair
    static constexpr int SLOT_TILE = 32 * SLOT;                   // float2 per group
    static constexpr int MAG_TILE = 2 * MAGS;                     // float2 per group
    static size_t smem_bytes(int bars, int nseg, int sg_half)
    {
        int m = 2 * sg_half + 1;
        return (size_t)(2 * SLOT_TILE + 2 * MAG_TILE + 15 * 16) * sizeof(float2) + (size_t)(nseg + bars) * F * sizeof(float) +
               (size_t)F * 4 * sizeof(int) + (size_t)nseg * sizeof(int) + (size_t)(bars + 1) * sizeof(int) + (size_t)bars * sizeof(float) +
               (size_t)m * m * sizeof(float) + 4 * sizeof(unsigned long long) + 64;
    }
};

template <int HOPQ, bool WINDOWED, int SGW, int ROWS>
__global__ void __launch_bounds__(RdifWs2Shape<ROWS>::THREADS, 1) stft_bars_rdif_ws2_kernel(const RdifParams p)
{
    using Sh = RdifWs2Shape<ROWS>;
    constexpr int N = Sh::N, M = Sh::M, SUBG = Sh::SUBG, SLOT = Sh::SLOT, MAGS = Sh::MAGS, F = Sh::F;
    constexpr int SLOT_TILE = Sh::SLOT_TILE, MAG_TILE = Sh::MAG_TILE;
    constexpr int NX = 16 + HOPQ * (F - 1);

    extern __shared__ float2 smem[];
    float2* const slots0 = smem;                               // [2 groups][SLOT_TILE]
    float2* const mags0 = smem + 2 * SLOT_TILE;                // [2 groups][MAG_TILE]
    float2* const t_twc = mags0 + 2 * MAG_TILE;                // [15][16]  W_256^(j*r), r = 1..15
    float* segsum = reinterpret_cast<float*>(t_twc + 15 * 16); // [nseg][F]
    float* raw = segsum + p.nseg * F;                          // [bars][F]
    int* beat_h = reinterpret_cast<int*>(raw + p.bars_n * F);  // [F][4]
    int* t_seg = beat_h + F * 4;                               // [nseg]
    int* t_bar0 = t_seg + p.nseg;                              // [bars+1]
    float* t_inv = reinterpret_cast<float*>(t_bar0 + p.bars_n + 1);
    float* t_sg = t_inv + p.bars_n;
    unsigned long long* mbar =
        reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(t_sg + (2 * p.sg_half + 1) * (2 * p.sg_half + 1)) + 15) & ~(uintptr_t)15);
    unsigned long long* mags_full = mbar;      // [2]
    unsigned long long* mags_empty = mbar + 2; // [2]

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int g = 0; g < 2; g++) {
            mbar_init(&mags_full[g], Sh::G_THREADS);
            mbar_init(&mags_empty[g], Sh::T_THREADS);
        }
    }
    for (int i = tid; i < 15 * 16; i += Sh::THREADS) t_twc[i] = __ldg(p.twm + (i & 15) * ((i >> 4) + 1));
    for (int i = tid; i < p.nseg; i += Sh::THREADS) t_seg[i] = __ldg(p.seg_packed + i);
    for (int i = tid; i <= p.bars_n; i += Sh::THREADS) t_bar0[i] = __ldg(p.bar_seg0 + i);
    for (int i = tid; i < p.bars_n; i += Sh::THREADS) t_inv[i] = __ldg(p.bar_inv + i);
    for (int i = tid; i < (2 * p.sg_half + 1) * (2 * p.sg_half + 1); i += Sh::THREADS) t_sg[i] = p.sg ? __ldg(p.sg + i) : 0.0f;
    __syncthreads();

    const long long tile_stride = 2ll * gridDim.x;
    if (tid < 2 * Sh::G_THREADS) {
        // ============================== transform group g ==============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        const int g = tid >> 8, t = tid & 255;
        const int j = t & (SUBG - 1), hw = t / SUBG;
        float2* const slots = slots0 + g * SLOT_TILE;
        float2* const mags = mags0 + g * MAG_TILE;
        const float2 w1 = __ldg(p.tw + t); // W_N^t = (cos theta_t, -sin theta_t)
        auto wsync = [] { __syncwarp(); };
        int i = 0;
#ifdef FFTL_WS2_PROF
        long long prof_wait_ = 0, prof_t0_ = clock64();
#endif
        TileCursor cur;
        cur.init(p, 2ll * blockIdx.x + g, tile_stride, F);
        for (; cur.left > 0; cur.advance(p, (int)tile_stride, F), i++) {
            const int ch = cur.ch;
            const long long f0 = cur.first_frame(p, F);
            const long long s0 = f0 * (long long)p.hop;
            const int16_t* xp = p.pcm + (long long)ch * p.channel_stride + s0 + t;
            // ---------------- pass A ----------------
            float xr[NX];
            if (s0 + (long long)NX * M <= p.samples) {
#pragma unroll
                for (int r = 0; r < NX; r++) xr[r] = pcm_to_float(xp + r * M);
            } else {
#pragma unroll
                for (int r = 0; r < NX; r++) xr[r] = (s0 + t + (long long)r * M < p.samples) ? pcm_to_float(xp + r * M) : 0.0f;
            }
            float win[16];
            if (WINDOWED) {
                // periodic Hann column: 0.5 - 0.5*cos(theta_t + q*pi/8), cos/sin(theta_t) = (w1.x, -w1.y)
#pragma unroll
                for (int q = 0; q < 16; q++) win[q] = fmaf(-0.5f * cos_pi8(q), w1.x, fmaf(-0.5f * sin_pi8(q), w1.y, 0.5f));
            }
            float2 twa[9]; // W_N^(t*k0) as running powers of W_N^t
            twa[1] = w1;
#pragma unroll
            for (int k0 = 2; k0 <= 8; k0++) twa[k0] = cmul(twa[k0 - 1], w1);
            named_bar_sync(2 + g, Sh::G_THREADS); // the group's previous pass B/C has finished reading the slots
#pragma unroll
            for (int pr = 0; pr < F / 2; pr++) {
                float2* ps = slots + pr * 16 * SLOT + pad_index(t);
                float y0[2], y8[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int f = 2 * pr + h;
                    float yr[9], yi[9];
                    real_dft16<WINDOWED>(xr + HOPQ * f, win, yr, yi);
                    y0[h] = yr[0];
                    y8[h] = yr[8];
#pragma unroll
                    for (int k0 = 1; k0 < 8; k0++) ps[(2 * k0 + h) * SLOT] = cmul(make_float2(yr[k0], yi[k0]), twa[k0]);
                }
                ps[0] = make_float2(y0[0], y0[1]);
                ps[SLOT] = cmul(make_float2(y8[0], y8[1]), twa[8]);
            }
            named_bar_sync(2 + g, Sh::G_THREADS);
            // ---------------- pass B/C ----------------
#ifdef FFTL_WS2_PROF
            long long c0_ = clock64();
#endif
            if (i >= 1) mbar_wait(&mags_empty[g], (i - 1) & 1); // the tail has read the previous tile's magnitudes
#ifdef FFTL_WS2_PROF
            prof_wait_ += clock64() - c0_;
#endif
            float2 twc[16];
#pragma unroll
            for (int r = 1; r < 16; r++) twc[r] = t_twc[(r - 1) * 16 + j];
#pragma unroll 1
            for (int pr = 0; pr < F / 2; pr++) {
                float2* s = slots + (pr * 16 + hw) * SLOT;
                float2 v[16];
                stockham_pass<M, 16, 1, SUBG, true>(v, s, p.twm, j, wsync);
#pragma unroll
                for (int r = 0; r < 16; r++) v[r] = s[pad_index(j + r * (M / 16))];
#pragma unroll
                for (int r = 1; r < 16; r++) v[r] = cmul(v[r], twc[r]);
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
                        const int c = kk < 8 ? kk : 15 - kk;
                        if (c < ROWS) {
                            float mag = sqrtf(fmaf(v[rr].x, v[rr].x, v[rr].y * v[rr].y));
                            if (kk < 8) lo_base[2 * 272 * c] = mag;
                            else hi_base[2 * 272 * c] = mag;
                        }
                    }
                } else {
                    const int src = hw ? (15 - j) : ((16 - j) & 15);
                    const bool self = (hw == 0) && (j == 0);
#pragma unroll
                    for (int kk = 0; kk < (ROWS < 8 ? ROWS : 8); kk++) {
                        const int rr = 4 * (kk & 3) + (kk >> 2);
                        const int kp = 15 - kk, rp = 4 * (kp & 3) + (kp >> 2);
                        float pr_ = __shfl_sync(0xffffffffu, v[rp].x, src, 16);
                        float pi_ = __shfl_sync(0xffffffffu, v[rp].y, src, 16);
                        if (self) {
                            const int ks = (16 - kk) & 15, rs = 4 * (ks & 3) + (ks >> 2);
                            pr_ = v[rs].x;
                            pi_ = v[rs].y;
                        }
                        float ar = v[rr].x + pr_, ai = v[rr].y - pi_;
                        float br = v[rr].x - pr_, bi = v[rr].y + pi_;
                        int bin = 16 * (j + 16 * kk) + (hw ? 8 : 0);
                        mg2[pad_index(bin)] = make_float2(0.5f * sqrtf(fmaf(ar, ar, ai * ai)), 0.5f * sqrtf(fmaf(br, br, bi * bi)));
                    }
                    if (self && ROWS >= 8) {
                        const int rr = 4 * (8 & 3) + (8 >> 2);
                        mg2[pad_index(N / 2)] = make_float2(fabsf(v[rr].x), fabsf(v[rr].y));
                    }
                }
            }
            mbar_arrive(&mags_full[g]);
        }
#ifdef FFTL_WS2_PROF
        if (blockIdx.x == 3 && t == 0) printf("group %d: tiles %d total %lld wait_empty %lld\n", g, i, clock64() - prof_t0_, prof_wait_);
#endif
    } else {
        // ============================== tail group ==============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int t = tid - 2 * Sh::G_THREADS;
        constexpr int TT = Sh::T_THREADS;
        const int B = p.bars_n;
#ifdef FFTL_WS2_PROF
        long long prof_wait_ = 0, prof_t0_ = clock64();
        long long ph_[6] = {0, 0, 0, 0, 0, 0};
#endif
        TileCursor cur2[2];
        cur2[0].init(p, 2ll * blockIdx.x, tile_stride, F);
        cur2[1].init(p, 2ll * blockIdx.x + 1, tile_stride, F);
        for (long long i = 0; cur2[0].left > 0; i++) {
#pragma unroll
            for (int g = 0; g < 2; g++) {
                TileCursor& cur = cur2[g];
                if (cur.left <= 0) break;
                const int ch = cur.ch;
                const long long f0 = cur.first_frame(p, F);
#ifdef FFTL_WS2_PROF
                long long c0_ = clock64();
#endif
                mbar_wait(&mags_full[g], (unsigned)(i & 1));
#ifdef FFTL_WS2_PROF
                prof_wait_ += clock64() - c0_;
#endif
#ifdef FFTL_WS2_SKIP_TAIL
                mbar_arrive(&mags_empty[g]);
                cur.advance(p, (int)tile_stride, F);
                continue;
#endif
                const float2* mags = mags0 + g * MAG_TILE;
#ifdef FFTL_WS2_PROF
                long long c1_ = clock64();
#endif
                rdif_segment_sums<F, MAGS>(t_seg, mags, segsum, p.nseg, t, TT);
                if (t < F && p.aux_w) { // bass bin of the tile's frames (beatDetector.c:53) before the buffer goes back
                    const long long fr = f0 + t;
                    if (fr < p.frame_end) {
                        const float2 mb = mags[(t >> 1) * MAGS + pad_index(p.bass_bin)];
                        p.aux_bass[(long long)ch * p.nf_aux + (fr - p.aux_base)] = trunc_i32((t & 1) ? mb.y : mb.x, p.strict);
                    }
                }
                mbar_arrive(&mags_empty[g]);
#ifdef FFTL_WS2_PROF
                long long c2_ = clock64();
#endif
                named_bar_sync(1, TT);
#ifdef FFTL_WS2_PROF
                long long c3_ = clock64();
#endif
                for (int bb = t; bb < B; bb += TT) {
                    const int a = t_bar0[bb], e = t_bar0[bb + 1];
                    float acc[F];
                    FrameVec<F>::load(segsum + a * F, acc);
                    for (int q = a + 1; q < e; q++) {
                        float u[F];
                        FrameVec<F>::load(segsum + q * F, u);
#pragma unroll
                        for (int f = 0; f < F; f++) acc[f] += u[f];
                    }
                    const float inv = t_inv[bb];
#pragma unroll
                    for (int f = 0; f < F; f++) acc[f] *= inv;
                    FrameVec<F>::store(raw + bb * F, acc);
                }
#ifdef FFTL_WS2_PROF
                long long c4_ = clock64();
#endif
                named_bar_sync(1, TT);
#ifdef FFTL_WS2_PROF
                long long c5_ = clock64();
#endif
                rdif_emit_beat<F, SGW>(p, raw, t_sg, nullptr, MAGS, t, ch, f0);
                rdif_emit_bars<F, SGW>(p, raw, t_sg, t, TT, cur.tic >= p.tic_emit_lo && cur.tic < p.tic_emit_hi, cur.o0, f0);
#ifdef FFTL_WS2_PROF
                long long c6_ = clock64();
                ph_[0] += c2_ - c1_; ph_[1] += c3_ - c2_; ph_[2] += c4_ - c3_; ph_[3] += c5_ - c4_; ph_[4] += c6_ - c5_;
#endif
                cur.advance(p, (int)tile_stride, F);
            }
        }
#ifdef FFTL_WS2_PROF
        if (blockIdx.x == 3 && t == 0) printf("tail: total %lld wait_full %lld seg %lld bar1 %lld bars %lld bar2 %lld sg %lld bar3 %lld\n", clock64() - prof_t0_, prof_wait_, ph_[0], ph_[1], ph_[2], ph_[3], ph_[4], ph_[5]);
        if (blockIdx.x == 3 && t == 255) printf("tail255: total %lld wait_full %lld seg %lld bar1 %lld bars %lld bar2 %lld sg %lld bar3 %lld\n", clock64() - prof_t0_, prof_wait_, ph_[0], ph_[1], ph_[2], ph_[3], ph_[4], ph_[5]);
#endif
    }
}

} // namespace fftl

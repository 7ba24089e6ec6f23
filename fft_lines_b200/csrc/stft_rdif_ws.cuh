// stft_rdif_ws.cuh -- warp-specialised version of the fast path (one persistent CTA per SM).
//
// Same arithmetic as stft_rdif.cuh (real-input DIF, N = 16*256, frames f/f^1 packed for
// residues 0 and 8, balanced segment tail).  What changes is the schedule.  ncu showed the
// phase-by-phase kernel bound by warp latency: with 16 warps per SM about one warp per
// scheduler was always parked at a CTA-wide barrier and another on a load, because every
// warp walks through every phase.  Here the three phases are ROLES that run concurrently on
// different tiles and hand buffers to each other through mbarriers, so no warp ever waits for
// a phase it is not part of:
//
//   warps  0..7   (256 threads) pass A   : PCM -> real DFT-16 -> twiddle -> slots[b]
//   warps  8..15  (256 threads) pass B/C : slots[b] -> 256-pt sub-transforms -> |X| -> mags[b]
//   warps 16..19  (128 threads) tail     : mags[b] -> segments -> bars -> SG -> outputs, w, bass
//
// slots and magnitudes are double buffered (b = tile parity): 2*69.6 KB + 2*30.5 KB of the
// 227 KB of shared memory.  full/empty mbarriers per buffer:
//   slots_full[b]  A  -> B/C  (256 arrivals)     slots_empty[b]  B/C -> A    (256)
//   mags_full[b]   B/C -> T   (256 arrivals)     mags_empty[b]   T   -> B/C  (128)
// The tail warps use a named barrier (id 1, 128 threads) between their own stages.
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

template <int ROWS> struct RdifWsShape {
    static constexpr int N = 4096, M = 256, SUBG = 16, F = 4;
    static constexpr int A_THREADS = 256, BC_THREADS = 256, T_THREADS = 128, THREADS = A_THREADS + BC_THREADS + T_THREADS;
    static constexpr int SLOT = padded_size(M);
    static constexpr int MAGS = RdifShape<N, F>::mag_stride(ROWS);                 // float2 per frame pair
    static constexpr size_t SLOT_BYTES = (size_t)32 * SLOT * sizeof(float2);      // one tile
    static constexpr size_t MAG_BYTES = (size_t)2 * MAGS * sizeof(float2);        // one tile
    static size_t smem_bytes(int bars, int nseg, int sg_half)
    {
        int m = 2 * sg_half + 1;
        return 2 * SLOT_BYTES + 2 * MAG_BYTES + (size_t)(nseg + bars) * F * sizeof(float) + (size_t)F * 4 * sizeof(int) +
               (size_t)nseg * sizeof(int) + (size_t)(bars + 1) * sizeof(int) + (size_t)bars * sizeof(float) + (size_t)m * m * sizeof(float) +
               8 * sizeof(unsigned long long) + 64;
    }
};

template <int HOPQ, bool WINDOWED, int SGW, int ROWS>
__global__ void __launch_bounds__(RdifWsShape<ROWS>::THREADS, 1) stft_bars_rdif_ws_kernel(const RdifParams p)
{
    using Sh = RdifWsShape<ROWS>;
    constexpr int N = Sh::N, M = Sh::M, SUBG = Sh::SUBG, SLOT = Sh::SLOT, MAGS = Sh::MAGS, F = Sh::F;
    constexpr int NX = 16 + HOPQ * (F - 1);

    extern __shared__ float2 smem[];
    // buffers are addressed as smem + offset so the compiler keeps them in the shared address space
    constexpr int SLOT_TILE = 32 * SLOT;          // float2 per tile of slots
    constexpr int MAG_TILE = 2 * MAGS;            // float2 per tile of magnitudes
    float2* const slots0 = smem;                  // [2][SLOT_TILE]
    float2* const mags0 = smem + 2 * SLOT_TILE;   // [2][MAG_TILE]
    float* segsum = reinterpret_cast<float*>(smem + 2 * SLOT_TILE + 2 * MAG_TILE); // [nseg][F]
    float* raw = segsum + p.nseg * F;                             // [bars][F]
    int* beat_h = reinterpret_cast<int*>(raw + p.bars_n * F);     // [F][4]
    int* t_seg = beat_h + F * 4;                                  // [nseg]
    int* t_bar0 = t_seg + p.nseg;                                 // [bars+1]
    float* t_inv = reinterpret_cast<float*>(t_bar0 + p.bars_n + 1);
    float* t_sg = t_inv + p.bars_n;
    unsigned long long* bars_mb =
        reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(t_sg + (2 * p.sg_half + 1) * (2 * p.sg_half + 1)) + 15) & ~(uintptr_t)15);
    unsigned long long* slots_full = bars_mb;      // [2]
    unsigned long long* slots_empty = bars_mb + 2; // [2]
    unsigned long long* mags_full = bars_mb + 4;   // [2]
    unsigned long long* mags_empty = bars_mb + 6;  // [2]

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < 2; b++) {
            mbar_init(&slots_full[b], Sh::A_THREADS);
            mbar_init(&slots_empty[b], Sh::BC_THREADS);
            mbar_init(&mags_full[b], Sh::BC_THREADS);
            mbar_init(&mags_empty[b], Sh::T_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // tail tables (read by the tail warps only, staged by everyone once)
    for (int i = tid; i < p.nseg; i += Sh::THREADS) {
        const int l = __ldg(p.seg_lo + i);
        t_seg[i] = pad_index(l) | ((__ldg(p.seg_hi + i) - l) << 16);
    }
    for (int i = tid; i <= p.bars_n; i += Sh::THREADS) t_bar0[i] = __ldg(p.bar_seg0 + i);
    for (int i = tid; i < p.bars_n; i += Sh::THREADS) t_inv[i] = __ldg(p.bar_inv + i);
    for (int i = tid; i < (2 * p.sg_half + 1) * (2 * p.sg_half + 1); i += Sh::THREADS) t_sg[i] = p.sg ? __ldg(p.sg + i) : 0.0f;
    __syncthreads();

    if (tid < Sh::A_THREADS) {
        // ============================== role A: pass A ==============================
        const int t = tid;
        float win[16];
        float2 twa[9];
        if (WINDOWED) {
#pragma unroll
            for (int q = 0; q < 16; q++) win[q] = __ldg(p.window + t + M * q);
        }
#pragma unroll
        for (int k0 = 1; k0 <= 8; k0++) twa[k0] = __ldg(p.tw + t * k0);
        int i = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, i++) {
            const int b = i & 1, k = i >> 1;
            const int ch = (int)(tile / p.tiles_per_channel);
            const long long f0 = p.frame0 + (tile % p.tiles_per_channel) * F;
            const long long s0 = f0 * (long long)p.hop;
            const int16_t* xp = p.pcm + (long long)ch * p.channel_stride + s0 + t;
            float xr[NX];
            if (s0 + (long long)NX * M <= p.samples) {
#pragma unroll
                for (int r = 0; r < NX; r++) xr[r] = (float)xp[r * M];
            } else {
#pragma unroll
                for (int r = 0; r < NX; r++) xr[r] = (s0 + t + (long long)r * M < p.samples) ? (float)xp[r * M] : 0.0f;
            }
            if (k >= 1) mbar_wait(&slots_empty[b], (k - 1) & 1); // pass B/C is done with this buffer's previous tile
            float2* slots = slots0 + b * SLOT_TILE;
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
            mbar_arrive(&slots_full[b]);
        }
    } else if (tid < Sh::A_THREADS + Sh::BC_THREADS) {
        // ============================== role B/C: sub-transforms ==============================
        const int t = tid - Sh::A_THREADS;
        const int j = t & (SUBG - 1), hw = t / SUBG;
        float2 twc[16];
#pragma unroll
        for (int r = 1; r < 16; r++) twc[r] = __ldg(p.twm + j * r);
        auto wsync = [] { __syncwarp(); };
        int i = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, i++) {
            const int b = i & 1, k = i >> 1;
            mbar_wait(&slots_full[b], k & 1);
            if (k >= 1) mbar_wait(&mags_empty[b], (k - 1) & 1); // the tail has read this buffer's previous magnitudes
            float2* slots = slots0 + b * SLOT_TILE;
            float2* mags = mags0 + b * MAG_TILE;
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
            mbar_arrive(&slots_empty[b]);
            mbar_arrive(&mags_full[b]);
        }
    } else {
        // ============================== role T: tail ==============================
        const int t = tid - Sh::A_THREADS - Sh::BC_THREADS;
        constexpr int TT = Sh::T_THREADS;
        const int B = p.bars_n;
        int i = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, i++) {
            const int b = i & 1, k = i >> 1;
            const int ch = (int)(tile / p.tiles_per_channel);
            const long long f0 = p.frame0 + (tile % p.tiles_per_channel) * F;
            mbar_wait(&mags_full[b], k & 1);
            const float2* mags = mags0 + b * MAG_TILE;
            for (int sgi = t; sgi < p.nseg; sgi += TT) {
                const int sd = t_seg[sgi];
                const int cnt = sd >> 16;
                const float2* m0 = mags + (sd & 0xffff);
                float acc[F];
#pragma unroll
                for (int f = 0; f < F; f++) acc[f] = 0.0f;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (q < cnt) {
#pragma unroll
                        for (int pr = 0; pr < F / 2; pr++) {
                            const float2 a = m0[pr * MAGS + q];
                            acc[2 * pr] += a.x;
                            acc[2 * pr + 1] += a.y;
                        }
                    }
                }
                FrameVec<F>::store(segsum + sgi * F, acc);
            }
            // bass bin of the tile's frames (beatDetector.c:53) before the buffer is handed back
            if (t < F && p.aux_w) {
                const long long fr = f0 + t;
                if (fr < p.frame_end) {
                    const float2 mb = mags[(t >> 1) * MAGS + pad_index(p.bass_bin)];
                    p.aux_bass[(long long)ch * p.nf_aux + (fr - p.aux_base)] = trunc_i32((t & 1) ? mb.y : mb.x, p.strict);
                }
            }
            mbar_arrive(&mags_empty[b]);
            named_bar_sync(1, TT);
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
            named_bar_sync(1, TT);
            {
                const int w = SGW >= 0 ? SGW : p.sg_half, m = 2 * w + 1;
                const long long dlo = p.frame_emit - f0, dhi = p.frame_end - f0;
                const int e_lo = dlo > 0 ? (dlo < F ? (int)dlo : F) : 0;
                const int e_hi = dhi < F ? (dhi > 0 ? (int)dhi : 0) : F;
                const long long o0 = ((long long)ch * p.nf_emit + (f0 - p.frame_emit)) * B;
                const bool want_int = (p.bars_i32 != nullptr) || (p.aux_w != nullptr);
#pragma unroll 1
                for (int bb = t; bb < B; bb += TT) {
                    float y[F];
                    if (w == 0) FrameVec<F>::load(raw + bb * F, y);
                    else {
                        int row, rb;
                        if (bb < w) { row = bb; rb = 0; }
                        else if (bb >= B - w) { row = m - (B - bb); rb = B - m; }
                        else { row = w; rb = bb - w; }
                        sg_apply<F, (SGW > 0 ? 2 * SGW + 1 : -1)>(t_sg + row * m, raw + rb * F, m, y);
                    }
                    float* ob = p.bars + (o0 + bb);
#pragma unroll
                    for (int f = 0; f < F; f++) {
                        y[f] *= p.zoom;
                        if (f >= e_lo && f < e_hi) ob[f * B] = y[f] + p.circle_add;
                    }
                    if (want_int) {
                        const bool isbeat = (bb == p.beat_bars[0]) | (bb == p.beat_bars[1]) | (bb == p.beat_bars[2]) | (bb == p.beat_bars[3]);
                        if (p.bars_i32 || isbeat) {
                            int* oi = p.bars_i32 ? p.bars_i32 + (o0 + bb) : nullptr;
                            const int cadd = (int)p.circle_add;
#pragma unroll
                            for (int f = 0; f < F; f++) {
                                const int ih = wrap_add(trunc_i32(y[f], p.strict), cadd);
                                if (oi && f >= e_lo && f < e_hi) oi[f * B] = ih;
                                if (isbeat) {
#pragma unroll
                                    for (int q = 0; q < 4; q++)
                                        if (bb == p.beat_bars[q]) beat_h[f * 4 + q] = ih;
                                }
                            }
                        }
                    }
                }
            }
            named_bar_sync(1, TT);
            if (t < F && p.aux_w) {
                const long long fr = f0 + t;
                if (fr < p.frame_end) {
                    const int* h4 = beat_h + 4 * t;
                    int sum = wrap_add(wrap_add(h4[0], h4[1]), wrap_add(h4[2], h4[3]));
                    p.aux_w[(long long)ch * p.nf_aux + (fr - p.aux_base)] = sum / 4; // fft_lines.c:281
                }
            }
            // beat_h / segsum / raw are rewritten only after the named barriers of the next tile
        }
    }
}

} // namespace fftl

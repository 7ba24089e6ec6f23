// stft_rdif_tm.cuh -- the N = 4096 / 2048 fast path with its per-thread constants parked in tensor memory (sm_100a):
// any hop, planar or interleaved PCM.
//
// stft_bars_rdif_kernel<4096, ...> (stft_rdif.cuh) keeps 62 per-thread constants in registers for the whole
// tile loop (window column, pass-A twiddles, pass-C twiddles): 127 registers, two CTAs of 256 threads per SM,
// and the kernel is bound by warp latency at 16 warps per SM, not by issue slots.  Here the same arithmetic
// runs with THREE CTAs per SM (24 warps, 80 registers; N = 2048: five CTAs of 128 threads instead of four):
//   * the constants live in TMEM (Blackwell's 256 KB per-SM tensor memory, 512 columns x 128 lanes x 32 bit):
//     every warp owns 64 columns of its 32-lane quarter, written once with tcgen05.st; pass A fetches the
//     window column and W_N^(t k0) per frame (two tcgen05.ld of 16 columns, live only while they are used),
//     pass B/C its W_M^(j r) once per tile (32 columns): a constant costs no register outside its phase and
//     no shared-memory or L1 traffic;
//   * the magnitudes overlay the first frame pair's slots (dead once every warp is past pass B/C of that
//     pair), so a CTA needs 76 KB of shared memory instead of 103 KB;
//   * the Savitzky-Golay + store step of a tile runs one tile late, in the middle of the next tile's pass B/C,
//     on the warps that do not own the packed slots; a split barrier (mbarrier arrive / wait) replaces the CTA
//     barrier in front of the next tile's pass A: four CTA barriers per tile instead of five.
// Same results bit for bit as stft_bars_rdif_kernel (same operations in the same order); tests compare both
// with each other, the generic kernel and the oracle.  What the reference does per tick:
// fft_lines/fft_lines.c:190-208, :281.
#pragma once
#include "stft_rdif.cuh"

namespace fftl {

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// COLS (64 or 128) columns of tensor memory for this CTA (one warp calls; the address lands in *holder)
template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t* holder)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_u32(holder)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_free(uint32_t taddr) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory"); }

// 32 consecutive columns of this thread's TMEM lane <- v[0..31]
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]),
        "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]),
        "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// v[0..31] <- 32 consecutive columns of this thread's TMEM lane; returns when the values have arrived
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
          "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]),
          "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]),
          "=f"(v[31])
        : "r"(taddr)
        : "memory");
}


// v[0..15] <- 16 consecutive columns of this thread's TMEM lane, asynchronously: tmem_wait16(v) before any use
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, float* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
          "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
// the wait names the registers as in/out operands so that no use of them can be scheduled in front of it
__device__ __forceinline__ void tmem_wait16(float* v)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]),
                   "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}

// TileCursor (stft_rdif.cuh) without the two 64-bit per-advance steps held in registers: they are recomputed from the
// launch parameters (constant bank) once per tile.  STRIDED: p.sample_stride may differ from 1 (interleaved PCM).
template <bool STRIDED> struct TileCursorLeanT {
    int left, tic, ch;
    const int16_t* xp;
    long long o0;
    __device__ __forceinline__ void init(const RdifParams& p, long long first, long long stride, int F)
    {
        left = first < p.total_tiles ? (int)((p.total_tiles - first + stride - 1) / stride) : 0;
        ch = (int)(first / p.tiles_per_channel);
        tic = (int)(first - (long long)ch * p.tiles_per_channel);
        const long long f0 = p.frame0 + (long long)tic * F;
        xp = p.pcm + (long long)ch * p.channel_stride + f0 * p.hop * (STRIDED ? p.sample_stride : 1);
        o0 = ((long long)ch * p.nf_emit + (f0 - p.frame_emit)) * p.bars_n;
    }
    __device__ __forceinline__ void advance(const RdifParams& p, int stride, int F)
    {
        left--;
        tic += stride;
        xp += (long long)(stride * F) * p.hop * (STRIDED ? p.sample_stride : 1);
        o0 += (long long)(stride * F) * p.bars_n;
        while (tic >= (int)p.tiles_per_channel) {
            tic -= (int)p.tiles_per_channel;
            ch++;
            xp += p.channel_stride - p.tiles_per_channel * F * p.hop * (STRIDED ? p.sample_stride : 1);
            o0 += (p.nf_emit - p.tiles_per_channel * F) * p.bars_n;
        }
    }
    __device__ __forceinline__ long long first_frame(const RdifParams& p, int F) const { return p.frame0 + (long long)tic * F; }
};

// split-phase CTA barrier on an mbarrier in shared memory (count = all threads of the CTA)
__device__ __forceinline__ void mbar_init(uint64_t* mb, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(mb)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* mb)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr_u32(mb)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mb, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(smem_addr_u32(mb)),
        "r"(parity)
        : "memory");
}

template <int N> struct RdifTmShape {
    static_assert(N == 4096 || N == 2048, "built for M = 256 (eight warps) and M = 128 (four warps)");
    static constexpr int F = 4, M = N / 16, THREADS = M, WARPS = M / 32, SUBG = M / 16;
    static constexpr int MINB = N == 4096 ? 3 : 5;                 // CTAs per SM the kernel is built for
    static constexpr int SLOT = padded_size(M);
    static constexpr int TM_COLS = WARPS > 4 ? 128 : 64;           // 64 columns per warp of a 32-lane quarter
    static constexpr bool SEG_IN_OVERLAY = N == 4096;              // the segment sums fit behind the magnitudes
    using Sh = RdifShape<N, F>;
    // float2 per frame pair in the magnitude area (bins < 256*rows, padded bin + bin/16)
    __host__ __device__ static constexpr int mags(int rows) { return N == 4096 ? Sh::mag_stride(rows) : padded_size(256 * rows); }
    // shared memory: [32-byte header] [32 slots] [raw bars] ([segment sums]) [tables]; the magnitudes overlay the first pair's slots
    static size_t smem_bytes(int bars, int nseg, int sg_half)
    {
        return (size_t)32 * SLOT * sizeof(float2) + (size_t)bars * F * sizeof(float) + (SEG_IN_OVERLAY ? 0 : (size_t)nseg * F * sizeof(float)) +
               Sh::table_bytes(bars, nseg, sg_half) + 32;
    }
    static bool fits(int bars, int nseg, int sg_half, int rows, int mag_bins)
    {
        if (mag_bins > 256 * rows) return false; // a bar (or the bass bin) reads bin N/2: no room for it in the overlay
        const size_t overlay = (size_t)2 * mags(rows) * sizeof(float2) + (SEG_IN_OVERLAY ? (size_t)nseg * F * sizeof(float) : 0);
        return overlay <= (size_t)16 * SLOT * sizeof(float2) && (smem_bytes(bars, nseg, sg_half) + 1024) * MINB <= 228 * 1024;
    }
};

template <int N, int HOPQ, bool WINDOWED, int SGW, int ROWS>
__global__ void __launch_bounds__(RdifTmShape<N>::THREADS, RdifTmShape<N>::MINB) stft_bars_rdif_tm_kernel(const RdifParams p)
{
    using Ts = RdifTmShape<N>;
    constexpr int F = Ts::F, M = Ts::M, SLOT = Ts::SLOT, THREADS = Ts::THREADS, SUBG = Ts::SUBG;
    constexpr int MAGS = Ts::mags(ROWS);
    // HOPQ = 1, 2, 4: hop = HOPQ*M, the tile's frames share their samples (NX per column and tile);
    // HOPQ = 0: any hop, planar or interleaved PCM: every frame loads its own 16 samples per column, the next frame's
    // while the current one is transformed
    static_assert(HOPQ == 0 || HOPQ == 1 || HOPQ == 2 || HOPQ == 4, "hop variants");
    static_assert(ROWS >= 1 && 256 * ROWS <= N / 2, "magnitude rows");
    constexpr int NX = HOPQ ? 16 + HOPQ * (F - 1) : 16;
    // second sub-pass: M = 256: one radix-16 butterfly per lane; M = 128: NT = 2 radix-8 butterflies per lane (jj = j + SUBG tt)
    constexpr int R2 = M == 256 ? 16 : 8, NT = 16 / R2, HALF = R2 / 2;
    constexpr int NMG = NT * 2 * (ROWS < HALF ? ROWS : HALF); // magnitudes a thread produces per sub-transform
    constexpr int RW = ROWS < HALF ? ROWS : HALF;             // rows per butterfly (= ROWS: 256*ROWS <= N/2)

    extern __shared__ float2 smem_raw[];
    // 32-byte header at a fixed place (no address registers): split barrier between the bar sums and the next tile's pass A
    // (8 bytes), TMEM address at byte 24
    uint64_t* mb = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* tm_holder = reinterpret_cast<uint32_t*>(mb + 3);
    float2* smem = smem_raw + 4;
    float2* slots = smem;                                        // [2 pairs][16][SLOT]
    float2* mags = smem;                                         // [2 pairs][MAGS], over the first pair's slots
    float* raw = reinterpret_cast<float*>(smem + 32 * SLOT);     // [bars][F]
    float* segsum = Ts::SEG_IN_OVERLAY ? reinterpret_cast<float*>(smem + 2 * MAGS) : raw + p.bars_n * F; // [nseg][F] (RdifTmShape::fits)
    int* t_seg = reinterpret_cast<int*>(raw + p.bars_n * F + (Ts::SEG_IN_OVERLAY ? 0 : p.nseg * F)); // [nseg] FFTL_SEG_PACK work list
    int* t_bar0 = t_seg + p.nseg;                                // [bars+1]
    float* t_inv = reinterpret_cast<float*>(t_bar0 + p.bars_n + 1); // [bars]
    float* t_sg = t_inv + p.bars_n;                              // [(2w+1)^2]
    const int t = threadIdx.x;
    const int B = p.bars_n;
    const int j = t & (SUBG - 1); // lane within the sub-transform
    const int hw = t / SUBG;      // which slot of a frame pair this group of SUBG lanes owns
    const int warp = t >> 5;

    if (warp == 0) tmem_alloc<Ts::TM_COLS>(tm_holder);
    if (t == 0) mbar_init(mb, THREADS);
    unsigned mb_parity = 0;
    for (int i = t; i < p.nseg; i += THREADS) t_seg[i] = __ldg(p.seg_packed + i);
    for (int i = t; i <= p.bars_n; i += THREADS) t_bar0[i] = __ldg(p.bar_seg0 + i);
    for (int i = t; i < p.bars_n; i += THREADS) t_inv[i] = __ldg(p.bar_inv + i);
    for (int i = t; i < (2 * p.sg_half + 1) * (2 * p.sg_half + 1); i += THREADS) t_sg[i] = p.sg ? __ldg(p.sg + i) : 0.0f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_base = *tm_holder;
    // this warp's 64 columns of its 32-lane quarter: [0, 16) window column, [16, 32) W_N^(t k0), [32, 64) W_M^(jj r)
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0); // tells the compiler the value is warp-uniform: the TMEM address stays in a uniform register
    const uint32_t tm = tm_base + ((uint32_t)(32 * (warp_u & 3)) << 16) + (uint32_t)(warp_u >> 2) * 64;
    {
        float c[32];
#pragma unroll
        for (int q = 0; q < 16; q++) c[q] = WINDOWED ? __ldg(p.window + t + M * q) : 1.0f;
#pragma unroll
        for (int k0 = 1; k0 <= 8; k0++) {
            const float2 w = __ldg(p.tw + t * k0);
            c[14 + 2 * k0] = w.x;
            c[15 + 2 * k0] = w.y;
        }
        tmem_st32(tm, c);
#pragma unroll
        for (int i = 0; i < 32; i++) c[i] = 0.0f;
        // twiddle of input r of butterfly tt sits at 2 * ((R2 - 1) * tt + r - 1)
#pragma unroll
        for (int tt = 0; tt < NT; tt++)
#pragma unroll
            for (int r = 1; r < R2; r++) {
                const float2 w = __ldg(p.twm + (((j + SUBG * tt) * r) & (M - 1)));
                c[2 * ((R2 - 1) * tt + r - 1)] = w.x;
                c[2 * ((R2 - 1) * tt + r - 1) + 1] = w.y;
            }
        tmem_st32(tm + 32, c);
        tmem_wait_st();
    }

    TileCursorLeanT<HOPQ == 0> cur;
    cur.init(p, blockIdx.x, gridDim.x, F);
    // Heights of a tile are written one tile late: the Savitzky-Golay + store step of tile i-1 runs in the middle of
    // pass B/C of tile i, on warps 1.. only -- warp 0 owns the two packed slots, whose untangling costs it ~70
    // instructions more per frame pair than a regular slot costs the others, so the other warps would otherwise wait
    // for it at both barriers of pass B/C; and the tile loop has one barrier and one latency-bound phase less.
    // Warp 0's share of that step is the band energy.
    bool have_prev = false;
    int prev_tic = 0, prev_ch = 0;
    long long prev_o0 = 0;
    auto emit_prev = [&](const int tid, const int nthr, const bool beat_lanes) {
        const long long pf0 = p.frame0 + (long long)prev_tic * F;
        if (beat_lanes) rdif_emit_beat<F, SGW>(p, raw, t_sg, nullptr, MAGS, t, prev_ch, pf0);
        if (tid >= 0) rdif_emit_bars<F, SGW>(p, raw, t_sg, tid, nthr, prev_tic >= p.tic_emit_lo && prev_tic < p.tic_emit_hi, prev_o0, pf0);
    };
    for (; cur.left > 0;) {
        // ================= pass A (both frame pairs) =================
        float xr[NX], xn[HOPQ ? 1 : 16]; // HOPQ = 0: the frame being transformed / the next one (roles swap per frame)
        const bool all_samples = cur.tic < p.tic_full; // every sample the tile reads exists
        // HOPQ = 0: the 16 samples of column t of frame f of the tile
        auto load_frame = [&](const int f, float* x) {
            const long long ss = p.sample_stride;
            const int16_t* xf = cur.xp + (t + (long long)f * p.hop) * ss;
            if (all_samples) {
#pragma unroll
                for (int q = 0; q < 16; q++) x[q] = pcm_to_float(xf + (long long)q * M * ss);
            } else {
                const long long s0 = (cur.first_frame(p, F) + f) * (long long)p.hop + t;
#pragma unroll
                for (int q = 0; q < 16; q++) x[q] = (s0 + (long long)q * M < p.samples) ? pcm_to_float(xf + (long long)q * M * ss) : 0.0f;
            }
        };
        if constexpr (HOPQ == 0) {
            load_frame(0, xr);
        } else {
            const int16_t* xp = cur.xp + t;
            if (all_samples) {
#pragma unroll
                for (int r = 0; r < NX; r++) xr[r] = pcm_to_float(xp + r * M);
            } else {
                const long long s0 = cur.first_frame(p, F) * (long long)p.hop; // the tile's first sample
#pragma unroll
                for (int r = 0; r < NX; r++) xr[r] = (s0 + t + (long long)r * M < p.samples) ? pcm_to_float(xp + r * M) : 0.0f;
            }
        }
        if (have_prev) { // every thread is done with the previous tile's magnitudes and segment sums
            mbar_wait(mb, mb_parity);
            mb_parity ^= 1u;
        }
        // the window column and the twiddles are fetched per frame (16 registers each, live only while they are used):
        // the window is dead after the first butterfly stage, the twiddles are asked for before it and waited for behind the DFT
#pragma unroll
        for (int pr = 0; pr < 2; pr++) {
            float2* ps = slots + pr * 16 * SLOT + pad_index(t);
            float y0[2], y8[2];
            float2 w8;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int f = 2 * pr + h;
                float yr[9], yi[9];
                float cw[16], ct[16];
                if (WINDOWED) tmem_ld16_async(tm, cw);
                tmem_ld16_async(tm + 16, ct);
                if constexpr (HOPQ == 0) {
                    // frames alternate between the two sample buffers; the next frame's loads travel during this frame's twiddle
                    // multiplies and slot stores and the other CTAs' work
                    float* xc = (f & 1) ? xn : xr;
                    float* xnext = (f & 1) ? xr : xn;
                    if (WINDOWED) tmem_wait16(cw);
                    float xcur[16];
#pragma unroll
                    for (int q = 0; q < 16; q++) xcur[q] = xc[q];
                    real_dft16<WINDOWED>(xcur, cw, yr, yi);
                    if (f + 1 < F) load_frame(f + 1, xnext); // behind the DFT: in front of it 14 registers spill (169 vs 175 M frames/s)
                } else {
                    if (WINDOWED) tmem_wait16(cw);
                    real_dft16<WINDOWED>(xr + HOPQ * f, cw, yr, yi);
                }
                tmem_wait16(ct);
                y0[h] = yr[0];
                y8[h] = yr[8];
#pragma unroll
                for (int k0 = 1; k0 < 8; k0++) ps[(2 * k0 + h) * SLOT] = cmul(make_float2(yr[k0], yi[k0]), make_float2(ct[2 * k0 - 2], ct[2 * k0 - 1]));
                w8 = make_float2(ct[14], ct[15]);
            }
            ps[0] = make_float2(y0[0], y0[1]);
            ps[SLOT] = cmul(make_float2(y8[0], y8[1]), w8);
        }
        __syncthreads(); // (also: the previous tile's raw bars are complete)

        // ================= pass B/C =================
        float cc[32]; // W_M^(jj r)
        tmem_ld32(tm + 32, cc);
        // register of butterfly output kk: Radix<R2>::out_index(reg_of(kk)) == kk
        auto reg_of = [](int kk) { return R2 == 16 ? (4 * (kk & 3) + (kk >> 2)) : (((kk & 3) << 1) | (kk >> 2)); };
        // first sub-pass of slot (pr, hw) (no twiddles): radix 16 over t = j + SUBG r through the slot, inputs of the second in v[]
        auto bc_first = [&](const int pr, float2* v) {
            float2* s = slots + (pr * 16 + hw) * SLOT;
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = s[pad_index(j + r * SUBG)];
            radix16_fused<false>(v, nullptr);
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < 16; rr++) s[pad_index(16 * j + Radix<16>::out_index(rr))] = v[rr];
            __syncwarp();
#pragma unroll
            for (int tt = 0; tt < NT; tt++)
#pragma unroll
                for (int r = 0; r < R2; r++) v[R2 * tt + r] = s[pad_index(j + SUBG * tt + 16 * r)];
        };
        // second sub-pass (twiddled) and the magnitudes this thread has to store, left in mg[]:
        // v[R2 tt + rr] = Z[m], m = jj + 16 kk, jj = j + SUBG tt, kk = Radix<R2>::out_index(rr)
        auto bc_second = [&](float2* v, float* mg) {
            if constexpr (R2 == 16) {
                float2 twc[16];
#pragma unroll
                for (int r = 1; r < 16; r++) twc[r] = make_float2(cc[2 * r - 2], cc[2 * r - 1]);
                radix16_fused<true>(v, twc);
            } else {
#pragma unroll
                for (int tt = 0; tt < NT; tt++) {
#pragma unroll
                    for (int r = 1; r < R2; r++)
                        v[R2 * tt + r] = cmul(v[R2 * tt + r], make_float2(cc[2 * ((R2 - 1) * tt + r - 1)], cc[2 * ((R2 - 1) * tt + r - 1) + 1]));
                    Radix<R2>::run(v + R2 * tt);
                }
            }
            if (hw >= 2) {
#pragma unroll
                for (int tt = 0; tt < NT; tt++)
#pragma unroll
                    for (int rr = 0; rr < R2; rr++) {
                        const int kk = Radix<R2>::out_index(rr);
                        const int c = kk < HALF ? kk : R2 - 1 - kk; // 256-bin row of this output
                        const float2 z = v[R2 * tt + rr];
                        if (c < RW) mg[2 * RW * tt + (kk < HALF ? c : RW + c)] = sqrt_approx_ftz(fmaf(z.x, z.x, z.y * z.y));
                    }
            } else {
                // packed slot (hw 0: residue 0, partner m' = M-m; hw 1: residue 8, partner m' = M-1-m): the partner of (jj, kk) is
                // butterfly NT-1-tt, output R2-1-kk of lane SUBG-1-j (hw 1) / SUBG-j (hw 0); lane 0 of hw 0 is its own partner:
                // butterfly (NT-tt) mod NT for tt > 0 and output (R2-kk) mod R2 for tt = 0
                constexpr unsigned MASK = SUBG >= 16 ? 0xffffffffu : 0xffffu; // the lanes of hw 0 and 1
                const int src = hw ? (SUBG - 1 - j) : ((SUBG - j) & (SUBG - 1));
                const bool self = (hw == 0) && (j == 0);
#pragma unroll
                for (int tt = 0; tt < NT; tt++)
#pragma unroll
                    for (int kk = 0; kk < RW; kk++) {
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
                        const float2 xa = cadd(z, make_float2(pr_, -pi_)); // 2 * Xa
                        const float2 xb = cadd(z, make_float2(-pr_, pi_)); // 2i * Xb
                        mg[2 * RW * tt + 2 * kk] = 0.5f * sqrt_approx_ftz(fmaf(xa.x, xa.x, xa.y * xa.y));
                        mg[2 * RW * tt + 2 * kk + 1] = 0.5f * sqrt_approx_ftz(fmaf(xb.x, xb.x, xb.y * xb.y));
                    }
            }
        };
        auto bc_store = [&](const int pr, const float* mg) {
            float2* mg2 = mags + pr * MAGS;
            if (hw >= 2) {
                // regular slot: frame h = hw&1, residue k0 = hw>>1; m < M/2 is bin k0 + 16m, m >= M/2 the conjugate mirror
                const int k0 = hw >> 1;
                float* mgf = reinterpret_cast<float*>(mg2) + (hw & 1);
#pragma unroll
                for (int tt = 0; tt < NT; tt++) {
                    const int jj = j + SUBG * tt;
                    float* lo_base = mgf + 2 * (k0 + 17 * jj);                 // pad_index(k0 + 16jj + 256c) = k0 + 17jj + 272c
                    float* hi_base = mgf + 2 * ((16 - k0) + 17 * (15 - jj));   // mirror: (16-k0) + 17(15-jj) + 272c
#pragma unroll
                    for (int c = 0; c < RW; c++) {
                        lo_base[2 * 272 * c] = mg[2 * RW * tt + c];
                        hi_base[2 * 272 * c] = mg[2 * RW * tt + RW + c];
                    }
                }
            } else {
#pragma unroll
                for (int tt = 0; tt < NT; tt++)
#pragma unroll
                    for (int kk = 0; kk < RW; kk++) {
                        const int bin = 16 * ((j + SUBG * tt) + 16 * kk) + (hw ? 8 : 0);
                        mg2[pad_index(bin)] = make_float2(mg[2 * RW * tt + 2 * kk], mg[2 * RW * tt + 2 * kk + 1]);
                    }
            }
        };
        {
            float mg0[NMG], mg1[NMG];
            float2 v[16];
            bc_first(0, v);
            bc_second(v, mg0);
            __syncthreads(); // every warp is done with the first pair's slots: the magnitudes may overlay them
            bc_store(0, mg0);
            if (have_prev) emit_prev(t - 32, THREADS - 32, t < 4 * F);
            bc_first(1, v);
            bc_second(v, mg1);
            bc_store(1, mg1);
        }
        __syncthreads();
        prev_tic = cur.tic;
        prev_ch = cur.ch;
        prev_o0 = cur.o0;
        const long long cur_f0 = cur.first_frame(p, F);
        cur.advance(p, gridDim.x, F);
        // ================= bins -> bars: every item carries the tile's F frames =================
        rdif_segment_sums<F, MAGS>(t_seg, mags, segsum, p.nseg, t, THREADS);
        __syncthreads();
        for (int b = t; b < B; b += THREADS) {
            const int a = t_bar0[b] & 0xffff, e = a + (t_bar0[b] >> 16); // the bar's range: first segment, count
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
        // bass = (int)|X[bass_bin]| (beatDetector.c:53) while the magnitudes are still in place: the last F threads
        if (p.aux_w != nullptr && t >= THREADS - F) {
            const int f = t - (THREADS - F);
            const long long fr = cur_f0 + f;
            if (fr < p.frame_end) {
                const float2 mb2 = mags[(f >> 1) * MAGS + pad_index(p.bass_bin)];
                p.aux_bass[(long long)prev_ch * p.nf_aux + (fr - p.aux_base)] = trunc_i32((f & 1) ? mb2.y : mb2.x, p.strict);
            }
        }
        have_prev = true;
        // Split barrier instead of a fifth CTA barrier: the next tile's pass A writes slots, and through them the
        // magnitudes (and segment sums) this phase has just read; a thread arrives here and waits only in front of its
        // first slot store, behind the next tile's sample loads.  `raw` is read behind the next tile's first two
        // barriers and rewritten behind its fourth.
        mbar_arrive(mb);
    }
    __syncthreads();
    if (have_prev) emit_prev(t, THREADS, t < 4 * F); // the CTA's last tile
    __syncthreads();
    if (warp == 0) tmem_free<Ts::TM_COLS>(tm_base);
}

} // namespace fftl

// rdif_launch.cu -- instantiations and launchers of the fast path for ONE hop variant
// (-DFFTL_RDIF_HOPQ=0|1|2|4; see rdif_launch.h).
#include <cuda_runtime.h>
#include "rdif_launch.h"
#include "stft_rdif_tm.cuh"

#ifndef FFTL_RDIF_HOPQ
#error "compile with -DFFTL_RDIF_HOPQ=0|1|2|4"
#endif

namespace fftl {

// Fast path (stft_rdif.cuh).  Returns cudaErrorNotSupported when the configuration is not covered.
template <int N, int F, int HOPQ, bool WIN, int SGW, int ROWS>
static cudaError_t launch_rdif(int num_sms, RdifParams p, int channels, cudaStream_t st)
{
    // design points: N = 4096: two CTAs of 256 threads per SM (128 registers, <= 113 KB shared memory each);
    // N = 2048: four CTAs of 128 threads (three when the tables are large); N = 1024: eight CTAs of 64 threads;
    // N = 8192 and 16384: one CTA of 512 threads
    constexpr int MINB = N >= 8192 ? 1 : N == 4096 ? 2 : (N == 2048 ? 4 : 8);
    constexpr bool RESIDENT = N != 16384; // two columns per thread at N = 16384: per-column constants are reloaded
    using Sh = RdifShape<N, F>;
#ifndef FFTL_NO_TM
    // N = 4096 / 2048 with shared samples: more CTAs per SM with the per-thread constants in tensor memory (stft_rdif_tm.cuh)
    // (N = 2048 with one magnitude row, i.e. the reference's own 100-bar configuration, is 3 % faster from registers: measured)
    if constexpr ((N == 4096 || (N == 2048 && ROWS > 1)) && F == 4 && 256 * ROWS <= N / 2) {
        using Ts = RdifTmShape<N>;
        if (!p.no_tm && Ts::fits(p.bars_n, p.nseg, p.sg_half, ROWS, p.mag_bins)) {
            if (!rdif_set_tiles(p, channels, F, HOPQ, N)) return cudaErrorNotSupported;
            if (p.total_tiles <= 0) return cudaSuccess;
            const size_t smem_tm = Ts::smem_bytes(p.bars_n, p.nseg, p.sg_half);
            auto kern_tm = stft_bars_rdif_tm_kernel<N, HOPQ, WIN, SGW, ROWS>;
            cudaError_t e = cudaFuncSetAttribute(kern_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tm);
            if (e != cudaSuccess) return e;
            long long grid = (long long)Ts::MINB * num_sms;
            if (grid > p.total_tiles) grid = p.total_tiles;
            if (p.used_tm) *p.used_tm = 1;
            kern_tm<<<(unsigned)grid, Ts::THREADS, smem_tm, st>>>(p);
            return cudaGetLastError();
        }
    }
#endif
    size_t smem = Sh::smem_bytes(p.bars_n, p.nseg, p.sg_half, ROWS);
    if ((smem + 1024) * (MINB <= 2 ? MINB : MINB - 1) > 228 * 1024) return cudaErrorNotSupported;
    if (!rdif_set_tiles(p, channels, F, HOPQ, N)) return cudaErrorNotSupported;
    if (p.total_tiles <= 0) return cudaSuccess;
    auto kern = stft_bars_rdif_kernel<N, F, HOPQ, WIN, RESIDENT, MINB, SGW, ROWS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = (long long)MINB * num_sms;
    if (grid > p.total_tiles) grid = p.total_tiles;
    kern<<<(unsigned)grid, Sh::THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

// magnitude rows kept: the smallest instantiated count that covers every bin a bar (or the bass bin) reads
template <int HOPQ, bool WIN, int SGW>
static cudaError_t launch_rdif_rows(int num_sms, const RdifParams& p, int channels, cudaStream_t st)
{
    if (p.mag_rows <= 1) return launch_rdif<4096, 4, HOPQ, WIN, SGW, 1>(num_sms, p, channels, st);
    if (p.mag_rows <= 4) return launch_rdif<4096, 4, HOPQ, WIN, SGW, 4>(num_sms, p, channels, st);
    if (p.mag_rows <= 7) return launch_rdif<4096, 4, HOPQ, WIN, SGW, 7>(num_sms, p, channels, st);
    return launch_rdif<4096, 4, HOPQ, WIN, SGW, 8>(num_sms, p, channels, st);
}

// N = 2048 (the reference's own size, global.h:3): run-time Savitzky-Golay width, one row or the full band
template <int HOPQ, bool WIN>
static cudaError_t launch_rdif_2k(int num_sms, const RdifParams& p, int channels, cudaStream_t st)
{
    if (p.mag_rows <= 1) return launch_rdif<2048, 4, HOPQ, WIN, -1, 1>(num_sms, p, channels, st);
    if (HOPQ > 0 && p.sg_half == 3) return launch_rdif<2048, 4, HOPQ, WIN, 3, 4>(num_sms, p, channels, st); // the default smoothing width, unrolled
    return launch_rdif<2048, 4, HOPQ, WIN, -1, 4>(num_sms, p, channels, st);
}

// N = 8192: one CTA of 512 threads per SM, all sixteen magnitude rows, run-time Savitzky-Golay width
template <int HOPQ, bool WIN>
static cudaError_t launch_rdif_8k(int num_sms, const RdifParams& p, int channels, cudaStream_t st)
{
    if (p.sg_half == 3) return launch_rdif<8192, 4, HOPQ, WIN, 3, 16>(num_sms, p, channels, st); // the default smoothing width, unrolled
    return launch_rdif<8192, 4, HOPQ, WIN, -1, 16>(num_sms, p, channels, st);
}

// N = 16384: one frame pair per tile, own samples, all thirty-two magnitude rows
template <bool WIN>
static cudaError_t launch_rdif_16k(int num_sms, const RdifParams& p, int channels, cudaStream_t st)
{
    if (p.sg_half == 3) return launch_rdif<16384, 2, 0, WIN, 3, 32>(num_sms, p, channels, st); // the default smoothing width, unrolled
    return launch_rdif<16384, 2, 0, WIN, -1, 32>(num_sms, p, channels, st);
}

// N = 1024 (own-samples variant only: the many-stream front end and odd hops)
template <bool WIN>
static cudaError_t launch_rdif_1k(int num_sms, const RdifParams& p, int channels, cudaStream_t st)
{
    // A job that is a single wave of CTAs anyway (one streaming tick: 1024 frames) is latency-bound: tiles of two
    // frames instead of four double the CTAs and halve the dependent work of each
    const long long frames = (p.frame_end - p.frame0) * channels;
    if (frames <= 2ll * 8 * num_sms) {
        if (p.mag_rows <= 1) return launch_rdif<1024, 2, 0, WIN, -1, 1>(num_sms, p, channels, st);
        return launch_rdif<1024, 2, 0, WIN, -1, 2>(num_sms, p, channels, st);
    }
    if (p.mag_rows <= 1) return launch_rdif<1024, 4, 0, WIN, -1, 1>(num_sms, p, channels, st);
    return launch_rdif<1024, 4, 0, WIN, -1, 2>(num_sms, p, channels, st);
}

template <int HOPQ, bool WIN>
static cudaError_t launch_rdif_sg(int num_sms, const RdifParams& p, int channels, cudaStream_t st)
{
#ifdef FFTL_GENERIC_SG
    return launch_rdif_rows<HOPQ, WIN, -1>(num_sms, p, channels, st);
#endif
    switch (p.sg_half) {
    case 0: return launch_rdif_rows<HOPQ, WIN, 0>(num_sms, p, channels, st);
    case 3: return launch_rdif_rows<HOPQ, WIN, 3>(num_sms, p, channels, st);
    default: return launch_rdif_rows<HOPQ, WIN, -1>(num_sms, p, channels, st);
    }
}


#define FFTL_CAT2(a, b) a##b
#define FFTL_CAT(a, b) FFTL_CAT2(a, b)
cudaError_t FFTL_CAT(launch_rdif_h, FFTL_RDIF_HOPQ)(int n, bool win, int num_sms, const RdifParams& p, int channels, cudaStream_t st)
{
    if (n == 16384) {
#if FFTL_RDIF_HOPQ == 0
        return win ? launch_rdif_16k<true>(num_sms, p, channels, st) : launch_rdif_16k<false>(num_sms, p, channels, st);
#else
        return cudaErrorNotSupported;
#endif
    }
    if (n == 8192) {
#if FFTL_RDIF_HOPQ == 0 || FFTL_RDIF_HOPQ == 1
        return win ? launch_rdif_8k<FFTL_RDIF_HOPQ, true>(num_sms, p, channels, st) : launch_rdif_8k<FFTL_RDIF_HOPQ, false>(num_sms, p, channels, st);
#else
        return cudaErrorNotSupported;
#endif
    }
    if (n == 1024) {
#if FFTL_RDIF_HOPQ == 0
        return win ? launch_rdif_1k<true>(num_sms, p, channels, st) : launch_rdif_1k<false>(num_sms, p, channels, st);
#else
        return cudaErrorNotSupported;
#endif
    }
    if (n == 2048) {
#if FFTL_RDIF_HOPQ == 0 || FFTL_RDIF_HOPQ == 4
        return win ? launch_rdif_2k<FFTL_RDIF_HOPQ, true>(num_sms, p, channels, st) : launch_rdif_2k<FFTL_RDIF_HOPQ, false>(num_sms, p, channels, st);
#else
        return cudaErrorNotSupported;
#endif
    }
#if FFTL_RDIF_HOPQ == 0
    // the own-samples variant is built with the default Savitzky-Golay width unrolled and with the run-time width
    if (p.sg_half == 3) return win ? launch_rdif_rows<0, true, 3>(num_sms, p, channels, st) : launch_rdif_rows<0, false, 3>(num_sms, p, channels, st);
    return win ? launch_rdif_rows<0, true, -1>(num_sms, p, channels, st) : launch_rdif_rows<0, false, -1>(num_sms, p, channels, st);
#else
    return win ? launch_rdif_sg<FFTL_RDIF_HOPQ, true>(num_sms, p, channels, st) : launch_rdif_sg<FFTL_RDIF_HOPQ, false>(num_sms, p, channels, st);
#endif
}

} // namespace fftl

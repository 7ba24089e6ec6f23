// rdif_launch.h -- host-side launchers of the fast path, one translation unit per hop variant
// (rdif_launch.cu compiled with -DFFTL_RDIF_HOPQ=0|1|2|4) so that the template instantiations build in parallel.
#pragma once
#include "stft_rdif.cuh"

namespace fftl {

// Tile bookkeeping of the fast path: tiles of F frames per channel and the tile ranges the kernel may treat
// without edge handling (TileCursor).  False when a channel has too many tiles for the 32-bit cursor.
inline bool rdif_set_tiles(RdifParams& p, int channels, int F, int hopq, int n = 4096)
{
    p.tiles_per_channel = (p.frame_end - p.frame0 + F - 1) / F;
    p.total_tiles = p.tiles_per_channel * channels;
    if (p.tiles_per_channel >= (1ll << 30)) return false;
    const long long span = hopq ? (16 + (long long)hopq * (F - 1)) * (n / 16) : (long long)(F - 1) * p.hop + n; // samples a tile reads per channel
    const long long q = p.samples - span - p.frame0 * (long long)p.hop;
    long long full = q < 0 ? 0 : q / ((long long)F * p.hop) + 1;                // tiles whose samples all exist
    p.tic_full = (int)(full > p.tiles_per_channel ? p.tiles_per_channel : full);
    p.tic_emit_lo = (int)((p.frame_emit - p.frame0 + F - 1) / F);               // first tile that starts at or after frame_emit
    p.tic_emit_hi = (int)((p.frame_end - p.frame0) / F);                        // tiles that end at or before frame_end
    return true;
}

// Launch the fast kernel for frames of n = 4096 or 2048 samples and hop = hopq*n/16 (hopq = 0: any hop /
// interleaved PCM; n = 2048 is built for hopq 0 and 4 only).  win: windowed.
// Returns cudaErrorNotSupported when the configuration does not fit (shared memory, tile count).
cudaError_t launch_rdif_h0(int n, bool win, int num_sms, const RdifParams& p, int channels, cudaStream_t st);
cudaError_t launch_rdif_h1(int n, bool win, int num_sms, const RdifParams& p, int channels, cudaStream_t st);
cudaError_t launch_rdif_h2(int n, bool win, int num_sms, const RdifParams& p, int channels, cudaStream_t st);
cudaError_t launch_rdif_h4(int n, bool win, int num_sms, const RdifParams& p, int channels, cudaStream_t st);

} // namespace fftl

/*
 * fftlines.h -- batched C-ABI of libfftlines_cuda (B200 / sm_100a).
 *
 * This is the generalisation of the reference's per-tick spectrum chain to
 * batched STFT.  The reference has no such entry points: it runs ONE frame per
 * WM_TIMER tick inside WndProc (reference fft_lines/fft_lines.c:186-236,281).
 * Each function below cites the reference lines whose work it performs for
 * many frames at once.  The single-frame drop-in entry points of the reference
 * itself live in fft_calculate.h and beatDetector.h next to this file.
 *
 * Plain C: pointers, sizes, PODs.  No torch / C++ types cross this boundary.
 * There is NO CPU fallback: every compute call fails with FFTL_ERR_CUDA (and
 * latches the error, see fftl_last_error) when no usable CUDA device exists.
 */
#ifndef FFTLINES_H
#define FFTLINES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFTL_VERSION 100

/* ---- status codes ------------------------------------------------------ */
enum {
    FFTL_OK            = 0,
    FFTL_ERR_ARG       = 1,  /* bad argument / unsupported configuration      */
    FFTL_ERR_CUDA      = 2,  /* CUDA runtime error (no device, launch, copy)  */
    FFTL_ERR_NOMEM     = 3,  /* host or device allocation failed              */
    FFTL_ERR_STATE     = 4   /* call sequence error (e.g. execute before init)*/
};

/* ---- configuration ----------------------------------------------------- */
enum { FFTL_WINDOW_RECT = 0, FFTL_WINDOW_HANN = 1 };
enum { FFTL_BINNING_LINEAR = 0, FFTL_BINNING_LOG = 1 };

/* Reference constants (fft_lines/settingsFile.h:4-13, beatDetector.c:6). */
#define FFTL_BEAT_SAMPLES   160   /* NUMBER_OF_SAMPLES, beatDetector.c:6      */
#define FFTL_TEMPO_RING     256   /* DEFAULT_BEATBBUFFERSIZE, settingsFile.h:11 */
#define FFTL_MAX_BARS       4096
#define FFTL_MAX_SG_HALF    8

typedef struct fftl_config {
    int32_t n;              /* window / FFT size: 256,512,...,16384 (reference: N=2048, global.h:3) */
    int32_t hop;            /* samples between frame starts (reference: whatever WASAPI delivered)   */
    int32_t bars;           /* bar count B (reference barCount 100..1000, settingsFile.h:4-5)       */
    int32_t window;         /* FFTL_WINDOW_RECT = reference behaviour (fft_lines.c:190-194)         */
    int32_t binning;        /* FFTL_BINNING_LINEAR = reference: bar i = bin i (fft_lines.c:199-203) */
    int32_t sg_half;        /* Savitzky-Golay half width w (0 = off = reference)                    */
    int32_t sg_degree;      /* SG polynomial degree d (< 2w+1)                                      */
    int32_t circle;         /* nonzero: +10 on every height (fft_lines.c:204-207)                   */
    float   zoom;           /* height = |X| * zoom (fft_lines.c:202; default 1e-4 settingsFile.h:13) */
    float   sample_rate;    /* Hz; only used by LOG binning and tempo (timeDelta = hop/sample_rate) */
    float   fmin;           /* LOG binning lower edge, Hz                                            */
    float   fmax;           /* LOG binning upper edge, Hz                                            */
    int32_t beat_bars[4];   /* bar indices averaged into the band energy; reference {3,4,5,6}
                               (fft_lines.c:281)                                                     */
    int32_t bass_bin;       /* FFT bin whose magnitude feeds the tempo ring; reference 4
                               (beatDetector.c:53)                                                   */
    int32_t strict_int;     /* 1: (int) casts overflow to INT32_MIN like x86 cvtt*2si (what the
                               reference binary does); 0: saturate                                   */
    int32_t reserved[3];
} fftl_config;

/* Per (channel, frame) beat record.
 *  w, thr, beat_value : AddBeatSample state after the frame (beatDetector.c:30-44)
 *  flag               : thr > 0 && w > thr (the reference has no boolean; this is ratio > 1)
 *  bass               : (int)|X[bass_bin]| written to the tempo ring (beatDetector.c:53)
 *  peak_index         : peaks[highestpeaknumber] (beatDetector.c:78-103)
 *  movement_per_sec   : N * timeDelta * peak_index (beatDetector.c:105), timeDelta = hop/fs
 */
typedef struct fftl_beat {
    int32_t w;
    int32_t thr;
    int32_t beat_value;
    int32_t flag;
    int32_t bass;
    int32_t peak_index;
    float   movement_per_sec;
    int32_t reserved;
} fftl_beat;

/* Fills *cfg with the reference's behaviour ("parity mode"): RECT window,
 * LINEAR binning, no smoothing, zoom 1e-4, beat bars 3..6, bass bin 4. */
void fftl_config_reference(fftl_config* cfg, int32_t n, int32_t hop, int32_t bars);
/* "Extended mode": Hann, LOG binning fmin..fmax, SG (w=3,d=2). Semantics are
 * defined by this repo's oracle; the reference has none of these stages. */
void fftl_config_extended(fftl_config* cfg, int32_t n, int32_t hop, int32_t bars,
                          float sample_rate, float fmin, float fmax);
/* FFTL_OK or FFTL_ERR_ARG (and latches a message). */
int  fftl_config_validate(const fftl_config* cfg);

/* Host-side tables derived from a config (also what the device uses).
 * lo/hi: bars entries; bar b = mean(|X[k]|, lo[b] <= k < hi[b]).            */
int  fftl_build_bins(const fftl_config* cfg, int32_t* lo, int32_t* hi);
/* coef: (2w+1) rows x (2w+1) taps, row p = weights that evaluate the fitted
 * polynomial at window position p; interior points use row w.               */
int  fftl_build_sg(int32_t sg_half, int32_t sg_degree, float* coef);

/* ---- batched STFT handle ----------------------------------------------- */
typedef struct fftl_stft fftl_stft;

/* device < 0: current device. Creates stream, uploads tables. */
int     fftl_stft_create(const fftl_config* cfg, int device, fftl_stft** out);
void    fftl_stft_destroy(fftl_stft* h);
/* F = floor((samples - n)/hop) + 1, or 0. */
int64_t fftl_stft_num_frames(const fftl_stft* h, int64_t samples_per_channel);
/* First frame a run over [frame_begin, frame_end) reads samples of: frame_begin itself (rounded down to even)
 * without beat output; with beat output the band-energy / bass history it recomputes in front of the range
 * (the start of frame_begin's 96-frame block minus the 256-entry tempo ring, beatDetector.c:53-67: up to
 * 95 + 256 frames back). */
int64_t fftl_stft_first_frame_needed(int64_t frame_begin, int32_t with_beat);
/* Samples [*first_sample, *end_sample) of every channel that a run over frames [frame_begin, frame_end) of a job
 * of total_samples_per_channel samples reads: from frame fftl_stft_first_frame_needed() to the end of the last
 * frame PAIR (frames 2k and 2k+1 share packed sub-transforms, so an even last frame also reads its partner),
 * clipped to the job.  This is what a frame-range shard has to hold (fftl_stft_run_device_span). */
int fftl_stft_sample_span(const fftl_stft* h, int64_t total_samples_per_channel, int64_t frame_begin, int64_t frame_end,
                          int32_t with_beat, int64_t* first_sample, int64_t* end_sample);

/* PCM addressing: sample s of channel c is pcm[c*channel_stride + s*sample_stride]
 * (planar: channel_stride = samples, sample_stride = 1; interleaved: 1, channels).
 * Frames [frame_begin, frame_end) of every channel are produced:
 *   bars     [channels][frame_end-frame_begin][bars] float   (height before the (int) cast)
 *   bars_i32 same shape, int32, the reference's BARINFO.height      (may be NULL)
 *   beat     [channels][frame_end-frame_begin]                       (may be NULL)
 * Beat outputs are those of a reference run started at frame 0: the band energy /
 * bass of up to 351 frames before frame_begin is recomputed internally
 * (fftl_stft_first_frame_needed; frame-range shards need no neighbour exchange).
 * The signal reads as zero beyond samples_per_channel (only the pair partner of
 * an odd job's last frame ever looks there).
 * ONE call per handle may be in flight at a time: the handle owns the scratch
 * between its two kernels.  Use one handle per stream / thread / GPU.
 * Restates for all frames: fft_lines.c:190-231 (bars), :281 + beatDetector.c:30-44
 * (band energy), :233-236 + beatDetector.c:51-106 (tempo).                       */
int fftl_stft_run_device(fftl_stft* h, const int16_t* pcm_dev, int32_t channels,
                         int64_t samples_per_channel, int64_t channel_stride,
                         int64_t sample_stride, int64_t frame_begin, int64_t frame_end,
                         float* bars_dev, int32_t* bars_i32_dev, fftl_beat* beat_dev,
                         void* cuda_stream /* cudaStream_t; NULL = CUDA's default stream */);

/* Frame-range shard that holds ONLY its own samples (multi-GPU sharding): pcm_dev points at sample
 * span_first_sample of every channel and holds span_samples of them; frame numbers, outputs and beat records
 * are those of the whole job of total_samples_per_channel samples (no re-basing on the caller's side, so the
 * tempo ring slot and the beat blocks of a frame are those of a single-GPU run: results are bit-identical).
 * The span must cover fftl_stft_sample_span(); FFTL_ERR_ARG otherwise. */
int fftl_stft_run_device_span(fftl_stft* h, const int16_t* pcm_dev, int32_t channels, int64_t span_first_sample,
                              int64_t span_samples, int64_t total_samples_per_channel, int64_t channel_stride,
                              int64_t sample_stride, int64_t frame_begin, int64_t frame_end, float* bars_dev,
                              int32_t* bars_i32_dev, fftl_beat* beat_dev, void* cuda_stream);

/* Same, host buffers.  Copies are chunked by frame range and overlapped with
 * the kernels on three streams; buffers from fftl_host_alloc (pinned) make the
 * copies asynchronous. Returns after all outputs are in host memory.          */
int fftl_stft_run(fftl_stft* h, const int16_t* pcm_host, int32_t channels,
                  int64_t samples_per_channel, int64_t channel_stride,
                  int64_t sample_stride, int64_t frame_begin, int64_t frame_end,
                  float* bars_host, int32_t* bars_i32_host, fftl_beat* beat_host);

/* Number of kernel launches issued by this handle since creation; how many of them were the fast path. */
int64_t fftl_stft_launch_count(const fftl_stft* h);
int64_t fftl_stft_fast_launch_count(const fftl_stft* h);
/* ... and how many of those ran the variant that parks its per-thread constants in tensor memory (n = 4096 with fewer
 * than 2048 bins read: three CTAs per SM instead of two; n = 2048 with more than 256 and fewer than 1025). */
int64_t fftl_stft_tmem_launch_count(const fftl_stft* h);
/* Which bars kernel a handle uses: AUTO = the real-input fast path wherever it is built (n = 1024..16384), else the
 * generic packed-complex kernel; GENERIC = always the latter (the correctness baseline the fast path is tested
 * against); FAST_REGISTERS = the fast path with its constants in registers even where the tensor-memory variant
 * exists (the two give the same bits; tests compare them).  All are this library's own sm_100a kernels. */
enum { FFTL_KERNEL_AUTO = 0, FFTL_KERNEL_GENERIC = 1, FFTL_KERNEL_FAST_REGISTERS = 2 };
int fftl_stft_set_kernel(fftl_stft* h, int32_t which);
/* Device time (ms, CUDA events recorded on the launching stream) summed over
 * the fftl_stft_run_device calls made since the previous fftl_stft_timing call
 * (at most the last 64): whole call, and the dominant stft_bars kernel only.
 * Waits for those calls to finish. */
int fftl_stft_timing(fftl_stft* h, int32_t* calls, float* total_ms, float* bars_kernel_ms);

/* ---- many-stream real-time front end ------------------------------------
 * `streams` sliding windows of the last n samples and their beat state live on
 * the device; one push = one tick of the reference's WM_TIMER for every stream:
 * GetAudioBuffer/MoveArray (wasapi_audio.cpp:390-484: shift the window, append
 * the new samples), then the spectrum chain and AddBeatSample/BassBeatDetector
 * on the stream's own rings (fft_lines.c:186-236, :281).  timeDelta of a tick is
 * count/sample_rate.  cfg->hop is ignored (count is the hop, it may vary).     */
enum { FFTL_MOVE_EXACT = 0,      /* window = previous window shifted by count, new samples appended           */
       FFTL_MOVE_REFERENCE = 1   /* the reference's MoveArray shifts by count-1 (wasapi_audio.cpp:394): the
                                    newest old sample is dropped on every push                                 */ };
typedef struct fftl_stream fftl_stream;
int  fftl_stream_create(const fftl_config* cfg, int32_t streams, int32_t move_mode, int device, fftl_stream** out);
void fftl_stream_destroy(fftl_stream* s);
int  fftl_stream_reset(fftl_stream* s);      /* windows, rings and tick counter back to zero */
/* fresh: [streams][count] int16 with row stride `stream_stride` (1 <= count).  Outputs: bars [streams][bars],
 * bars_i32 (may be NULL), beat [streams] (may be NULL).  Device variant is asynchronous on cuda_stream and can
 * be captured in a CUDA graph (all positions come from a device-resident tick counter).                       */
int  fftl_stream_push_device(fftl_stream* s, const int16_t* fresh_dev, int32_t count, int64_t stream_stride,
                             float* bars_dev, int32_t* bars_i32_dev, fftl_beat* beat_dev, void* cuda_stream);
int  fftl_stream_push(fftl_stream* s, const int16_t* fresh_host, int32_t count, int64_t stream_stride,
                      float* bars_host, int32_t* bars_i32_host, fftl_beat* beat_host);
/* Capture-format input (wasapi_audio.cpp:390-400, :458-466): the reference receives 32-bit float samples and
 * stores (int16_t)(x * 65536), which is undefined in C for |x*65536| >= 32768 (|x| >= 0.5).  Here the
 * conversion is DEFINED as what the reference's x86 binary does: truncate toward zero to int32
 * (cvttss2si: NaN and out-of-range give INT32_MIN), keep the low 16 bits.  fresh: [streams][count] float,
 * element (s, k) at fresh[s*stream_stride + k*sample_stride] (WASAPI stereo interleaved: sample_stride = 2). */
int  fftl_stream_push_device_f32(fftl_stream* s, const float* fresh_dev, int32_t count, int64_t stream_stride,
                                 int64_t sample_stride, float* bars_dev, int32_t* bars_i32_dev, fftl_beat* beat_dev,
                                 void* cuda_stream);
int64_t fftl_stream_launch_count(const fftl_stream* s);
/* How a push runs: AUTO = one kernel per tick where it is built (n = 1024 / 2048 with FFTL_MOVE_EXACT: window move,
 * transform with 4 points per thread, bars and beat stage of two streams per CTA -- laid out for latency), else the
 * three launches (window move, batched bars kernel, beat stage); THREE_LAUNCHES = always the latter (what the one-kernel
 * tick is tested against).  fftl_stream_fused_tick_count: pushes that ran as one kernel.                          */
enum { FFTL_TICK_AUTO = 0, FFTL_TICK_THREE_LAUNCHES = 1 };
int  fftl_stream_set_tick_kernel(fftl_stream* s, int32_t which);
int64_t fftl_stream_fused_tick_count(const fftl_stream* s);
/* Render-side history (the reference's "3D mode" keeps the last ~96 bar frames on screen by feeding the
 * bitmap back, drawBar2D.cpp:405-426,670-680).  depth > 0 keeps the last `depth` float bar frames of every
 * stream on the device; fftl_stream_history_device copies them out oldest first as [streams][depth][bars]
 * (frames before the first push read as 0).                                                              */
int  fftl_stream_set_history(fftl_stream* s, int32_t depth);
int  fftl_stream_history_device(fftl_stream* s, float* out_dev, void* cuda_stream);
/* Waveform view (fft_lines.c:239-247): height[i] = (int)(window[i] * scale + offset) for the first `points`
 * samples of every stream's current window (reference: 2048 points, scale 0.02, offset from the client rect). */
int  fftl_stream_waveform_device(fftl_stream* s, int32_t points, float scale, float offset, int32_t* heights_dev,
                                 void* cuda_stream);

/* ---- callers and sinks either side of the path -----
 * Settings line: the 11 space-separated numbers of Settings.txt, same order, same clamps and defaults as
 * readSettings/writeSettings (settingsFile.c:172-247, :368-378, limits settingsFile.h:4-25).               */
typedef struct fftl_settings {
    int32_t bar_count;      /* 100..1000, default 500 */
    float   zoom;           /* 1e-5..1e-3, default 1e-4 */
    int32_t color_sel;      /* 0..5, default 0 */
    int32_t dofft, background, gradient, ignore_serial, circle, waveform, stereo, beat_detection; /* 0/1 */
} fftl_settings;
void fftl_settings_default(fftl_settings* s);
int  fftl_settings_parse(const char* text, int64_t len, fftl_settings* out);
/* writes "%d %0.6f %lld %d %d %d %d %d %d %d %d"; returns the length or -1 if cap is too small */
int  fftl_settings_format(const fftl_settings* s, char* buf, int64_t cap);
/* reference-mode config (n, hop, sample_rate are not in Settings.txt) from a settings record */
void fftl_settings_to_config(const fftl_settings* s, int32_t n, int32_t hop, float sample_rate, fftl_config* cfg);

/* RANGE255(i, count): brush / colour-map index of bar i (drawBar2D.cpp:33).                                 */
uint32_t fftl_color_index(int32_t i, int32_t count);

/* Serial/LED sink (fft_lines.c:250-255): one byte per frame, (char)(height[led_bar] * 0.1f), sent only when
 * the height is >= 0.  heights: int32 [frames][bars]; out bytes [frames]; valid [frames] (may be NULL).     */
int fftl_led_bytes_device(const int32_t* heights_dev, int64_t frames, int32_t bars, int32_t led_bar,
                          uint8_t* bytes_dev, uint8_t* valid_dev, void* cuda_stream);

/* Beat-marker sweep (fft_lines.c:266-274): per frame the marker makes ceil(beatMovementPerSec*timeDelta)
 * unit steps, bouncing between 0 and 2048.  beat: [frames] records of ONE channel; pos_out/dir_out [frames]
 * hold beatposition / direction after each frame.  start_pos in 0..2048, start_dir -1, 0 or 1.              */
int fftl_marker_sweep_device(const fftl_beat* beat_dev, int64_t frames, float time_delta, int32_t start_pos,
                             int32_t start_dir, int32_t* pos_out_dev, int32_t* dir_out_dev, void* cuda_stream);

/* Pinned host memory helpers. */
void* fftl_host_alloc(uint64_t bytes);
void  fftl_host_free(void* p);

/* On-device synthetic PCM: sweep + Philox noise + AM bass tone.
 * Writes planar int16 [channels][samples] at pcm_dev; sample index offset
 * `first_sample` lets shards generate their own range.                        */
int fftl_synth_pcm_device(int16_t* pcm_dev, int32_t channels, int64_t samples,
                          int64_t channel_stride, int64_t first_sample,
                          float sample_rate, uint32_t seed, void* cuda_stream);

/* Measured dense FP32 FMA rate (TFLOP/s, 2 flop per FMA) of the current device:
 * the denominator of the compute roofline (not in MEASURED_PEAKS.json).      */
int fftl_measure_fp32_peak(double* tflops, int32_t reps);

/* ---- errors / info ----------------------------------------------------- */
int         fftl_last_error(void);          /* sticky until fftl_clear_error */
const char* fftl_last_error_string(void);
void        fftl_clear_error(void);
int         fftl_version(void);
int         fftl_device_count(void);        /* 0 when no CUDA device is usable */

/* Compat (single-frame) layer knobs; see fft_calculate.h / beatDetector.h. */
int  fftl_set_n(int32_t n);                 /* default 2048 = reference N (global.h:3) */
int  fftl_get_n(void);
void fftl_compat_reset_beat(void);          /* zero AddBeatSample / tempo state */
int64_t fftl_compat_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FFTLINES_H */

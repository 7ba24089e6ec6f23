/*
 * ref_harness.c -- headless driver around the reference's UNMODIFIED
 * fft_lines/fft_calculate.c and fft_lines/beatDetector.c.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into
 * oracle/_ref/libfftlines_ref.so together with those two reference files
 * (compiled where they lie under /root/reference; never copied) and
 * oracle/fftwf_shim.c as the FFT provider.
 *
 * fft_lines.c itself is a Win32 message loop and cannot be built here, so the
 * parts of WndProc that belong to the hot path are RESTATED below, each block
 * citing the lines it restates:
 *   WM_CREATE  fft_lines.c:110-116   buffers + plans
 *   WM_TIMER   fft_lines.c:190-236   load, FFT, bars, stereo, tempo
 *              fft_lines.c:281       band energy -> AddBeatSample
 *   WM_DESTROY fft_lines.c:420-433   cleanup order
 * Everything the reference computes with its own functions (plans, execute,
 * AddBeatSample, GetBeatValue, BassBeatDetector) runs the reference's code.
 *
 * Same harness, other provider: compiled with -DFFTL_HARNESS_CUDA_PROVIDER and
 * -I<repo>/include, and linked with -lfftlines_cuda INSTEAD of the two
 * reference files and the FFT shim, the very same restated WndProc drives the
 * drop-in library (tests/test_gpu_dropin_c.py replays the golden fixtures
 * through it).  Only what reaches into the reference's private state differs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>

#ifdef FFTL_HARNESS_CUDA_PROVIDER
#include "fft_calculate.h"         /* this repo's include/: the reference's prototypes, fft_calculate.h:5-11 */
#include "beatDetector.h"          /* ... beatDetector.h:3-11                                                 */
#include "fftlines.h"              /* fftl_set_n, fftl_compat_reset_beat, fftl_last_error                      */
#define N 2048                     /* global.h:3                       */
#define DEFAULT_BEATBBUFFERSIZE 256 /* settingsFile.h:11                */
int* bassBeatBuffer = NULL;        /* settingsFile.c:83: the application owns it; the library mirrors its ring here */
static int beatFramesRecorded = 0; /* the harness's own copy of beatDetector.c:23 (slot written by the next call) */
static int BeatThreshold = 0;      /* not exported by the drop-in: reported as 0 */
static void fftwf_cleanup(void) {}
#else
#include <fftw3.h>                 /* the reference's vendored header */
#include "global.h"                /* reference: N, globalhwnd         */
#include "settingsFile.h"          /* reference: constants, externs    */
#include "fft_calculate.h"         /* reference                        */
#include "beatDetector.h"          /* reference                        */

/* Symbols the two reference files import from the rest of the app. */
HWND globalhwnd = NULL;            /* fft_lines.c:46                   */
int* bassBeatBuffer = NULL;        /* settingsFile.c:83, alloc :295    */

/* Non-static state of beatDetector.c (:14-23) so a test can reset it. */
extern int beatValues[160];
extern int currentBeatValue, sumOfBeatValues, BeatThreshold, beatValue, beatFramesRecorded;

extern int fftl_shim_override_n;   /* oracle/fftwf_shim.c */
#endif

static fftwf_complex *input, *output, *beatinput, *beatoutput; /* fft_lines.c:63-67 */
static int g_n = N;

int ref_compiled_n(void) { return N; }

int ref_open(int n)
{
    /* fft_lines.c:110-116.  n != N only through the shim's size override. */
    g_n = n > 0 ? n : N;
#ifdef FFTL_HARNESS_CUDA_PROVIDER
    fftl_set_n(g_n);
#else
    fftl_shim_override_n = (g_n != N) ? g_n : 0;
#endif
    input = (fftwf_complex*)malloc(sizeof(fftwf_complex) * g_n);
    output = (fftwf_complex*)malloc(sizeof(fftwf_complex) * g_n);
    initializefft(input, output);
#ifndef FFTL_HARNESS_CUDA_PROVIDER
    fftl_shim_override_n = 0;
#endif
    beatinput = (fftwf_complex*)malloc(sizeof(fftwf_complex) * DEFAULT_BEATBBUFFERSIZE);
    beatoutput = (fftwf_complex*)malloc(sizeof(fftwf_complex) * DEFAULT_BEATBBUFFERSIZE);
    initializefft256(beatinput, beatoutput);
    /* settingsFile.c:295-301: 256 ints, zeroed */
    bassBeatBuffer = (int*)calloc(DEFAULT_BEATBBUFFERSIZE, sizeof(int));
    return (input && output && beatinput && beatoutput && bassBeatBuffer) ? 0 : -1;
}

void ref_reset_beat(void)
{
#ifdef FFTL_HARNESS_CUDA_PROVIDER
    fftl_compat_reset_beat();
    beatFramesRecorded = 0;
#else
    memset(beatValues, 0, sizeof(beatValues));
    currentBeatValue = sumOfBeatValues = BeatThreshold = beatValue = beatFramesRecorded = 0;
#endif
    beatMovementPerSec = 0; beatposition = 0; timeDelta = 0; direction = 0;
    if (bassBeatBuffer) memset(bassBeatBuffer, 0, sizeof(int) * DEFAULT_BEATBBUFFERSIZE);
}

void ref_close(void)
{
    /* fft_lines.c:420-433 */
    cleanupfft();
    cleanupfft256();
    fftwf_cleanup();
    free(input); free(output); free(beatinput); free(beatoutput); free(bassBeatBuffer);
    input = output = beatinput = beatoutput = NULL;
    bassBeatBuffer = NULL;
}

/* Direct access to the reference entry points for unit tests. */
void ref_execute(const float* in_interleaved, float* out_interleaved)
{
    memcpy(input, in_interleaved, sizeof(fftwf_complex) * g_n);
    executefft(input, output);
    memcpy(out_interleaved, output, sizeof(fftwf_complex) * g_n);
}
void ref_execute256(const float* in_interleaved, float* out_interleaved)
{
    memcpy(beatinput, in_interleaved, sizeof(fftwf_complex) * DEFAULT_BEATBBUFFERSIZE);
    executefft256(beatinput, beatoutput);
    memcpy(out_interleaved, beatoutput, sizeof(fftwf_complex) * DEFAULT_BEATBBUFFERSIZE);
}
void ref_add_beat_sample(int w) { AddBeatSample(w); }
int  ref_get_beat_value(void) { return GetBeatValue(); }
int  ref_get_threshold(void) { return BeatThreshold; }
#ifdef FFTL_HARNESS_CUDA_PROVIDER
int  ref_provider_error(void) { return fftl_last_error(); }
#endif
float ref_get_movement(void) { return beatMovementPerSec; }

typedef struct ref_frame_out {
    int32_t w;            /* argument passed to AddBeatSample           */
    int32_t thr;          /* BeatThreshold after the call               */
    int32_t beat_value;   /* GetBeatValue()                             */
    int32_t bass;         /* bassBeatBuffer slot written this frame     */
    float   movement_per_sec;
} ref_frame_out;

/* One WM_TIMER tick.  left/right: N int16 each (right NULL = mono).
 * heights_*: barCount ints = BARINFO.height.  td = timeDelta for this tick. */
void ref_tick(const int16_t* left, const int16_t* right, int barCount, float zoom, int circle,
              int beatDetection, float td, int32_t* heights_left, int32_t* heights_right,
              ref_frame_out* fo)
{
    /* fft_lines.c:190-194 */
    for (int i = 0; i < g_n; i++) {
        input[i][REAL] = (float)left[i];
        input[i][IMAG] = 0;
    }
    executefft(input, output); /* :196 */
    /* :199-208 */
    for (int i = 0; i < barCount; i++) {
        heights_left[i] = (int)(sqrt(pow(output[i][REAL], 2) + pow(output[i][IMAG], 2)) * zoom);
        if (circle) heights_left[i] += 10;
    }
    if (right) {
        /* :210-231 */
        for (int i = 0; i < g_n; i++) {
            input[i][REAL] = (float)right[i];
            input[i][IMAG] = 0;
        }
        executefft(input, output);
        for (int i = 0; i < barCount; i++) {
            heights_right[i] = (int)(sqrt(pow(output[i][REAL], 2) + pow(output[i][IMAG], 2)) * zoom);
            if (circle) heights_right[i] += 10;
        }
    }
    int slot = beatFramesRecorded;
    if (beatDetection) {
        timeDelta = td;                                      /* :265 (wall clock there) */
        BassBeatDetector(input, output, beatinput, beatoutput); /* :233-236 */
    }
    /* :281 */
    int w = (heights_left[3] + heights_left[4] + heights_left[5] + heights_left[6]) / 4;
    AddBeatSample(w);
    if (fo) {
        fo->w = w;
        fo->thr = BeatThreshold;
        fo->beat_value = GetBeatValue();
        fo->bass = beatDetection ? bassBeatBuffer[slot] : 0;
        fo->movement_per_sec = beatMovementPerSec;
    }
#ifdef FFTL_HARNESS_CUDA_PROVIDER
    if (beatDetection) beatFramesRecorded = (beatFramesRecorded + 1) % DEFAULT_BEATBBUFFERSIZE; /* beatDetector.c:54-59 */
#endif
}

/* Batched run of the reference tick over strided frames of planar PCM
 * [channels][samples] (channels 1 or 2 = the reference's mono / stereo).
 * Outputs: heights [channels][frames][barCount] int32, fo [frames].
 * Returns seconds spent in the frame loop (CLOCK_MONOTONIC). */
double ref_run(const int16_t* pcm, int channels, int64_t samples, int64_t channel_stride, int hop,
               int barCount, float zoom, int circle, int beatDetection, float sample_rate,
               int32_t* heights, ref_frame_out* fo)
{
    int64_t F = samples < g_n ? 0 : (samples - g_n) / hop + 1;
    int32_t* hl = (int32_t*)malloc(sizeof(int32_t) * barCount);
    int32_t* hr = (int32_t*)malloc(sizeof(int32_t) * barCount);
    float td = (float)hop / sample_rate;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int64_t f = 0; f < F; f++) {
        const int16_t* l = pcm + f * hop;
        const int16_t* r = channels > 1 ? pcm + channel_stride + f * hop : NULL;
        ref_frame_out o;
        ref_tick(l, r, barCount, zoom, circle, beatDetection, td, hl, hr, &o);
        if (heights) {
            memcpy(heights + f * barCount, hl, sizeof(int32_t) * barCount);
            if (r) memcpy(heights + (F + f) * barCount, hr, sizeof(int32_t) * barCount);
        }
        if (fo) fo[f] = o;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(hl); free(hr);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

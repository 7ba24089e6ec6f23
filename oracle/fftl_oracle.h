/*
 * fftl_oracle.h -- CPU oracle for the fft_lines spectrum hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fft_lines_b200/ or include/ may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg use it, as the checker.
 *
 * Parity status:
 *   - parity mode (RECT window, LINEAR bins, no SG): pinned against the
 *     UNMODIFIED reference sources fft_calculate.c + beatDetector.c compiled
 *     into oracle/_ref (see oracle/Makefile, oracle/ref_harness.c) and against
 *     closed-form known answers.  The reference ships no tests or golden
 *     vectors of its own, and its FFT arithmetic lives in a
 *     binary-only FFTW DLL (version not pinned, not in the tree), so the DFT
 *     itself is pinned by its mathematical definition.
 *   - extended mode (Hann, LOG bins, Savitzky-Golay): NOT in the reference.
 *     "parity unpinned": semantics are defined here and cross-checked against
 *     numpy/scipy in tests/test_oracle.py.
 */
#ifndef FFTL_ORACLE_H
#define FFTL_ORACLE_H
#include <stdint.h>
#include "../include/fftlines.h"   /* fftl_config / fftl_beat PODs only */

#ifdef __cplusplus
extern "C" {
#endif

/* x86 `cvttsd2si` semantics of the reference's (int) casts (strict != 0), or
 * saturating (strict == 0). */
int32_t orc_trunc_i32(double x, int strict);

/* Unnormalised forward DFT, double precision, n a power of two.
 * Definition of what fftwf_execute computes at fft_calculate.c:29,45. */
void orc_fft(int n, const double* in_re, const double* in_im, double* out_re, double* out_im);
/* O(n^2) DFT straight from the definition (pins orc_fft in the tests). */
void orc_dft_naive(int n, const double* in_re, const double* in_im, double* out_re, double* out_im);

int  orc_config_validate(const fftl_config* cfg);
int  orc_build_bins(const fftl_config* cfg, int32_t* lo, int32_t* hi);
int  orc_build_sg(int w, int d, double* coef /* (2w+1)*(2w+1) */);
void orc_window(const fftl_config* cfg, double* win /* n */);
int64_t orc_num_frames(const fftl_config* cfg, int64_t samples);

/* One frame: fft_lines.c:190-208.  pcm[i*sample_stride], i<n.
 * mag: n/2+1 magnitudes (may be NULL); height_f64/bars_i32: cfg->bars entries. */
void orc_frame(const fftl_config* cfg, const int16_t* pcm, int64_t sample_stride,
               double* mag, double* height_f64, int32_t* bars_i32);

/* AddBeatSample / GetBeatValue restated (beatDetector.c:14-49). */
typedef struct orc_beat_state {
    int32_t ring[FFTL_BEAT_SAMPLES];
    int32_t pos, sum, thr, beat_value;
    int32_t tempo_ring[FFTL_TEMPO_RING];
    int32_t tempo_pos;
    int32_t hb[FFTL_BEAT_SAMPLES + 1];   /* heightBuffer incl. the [160] the reference reads OOB */
    int32_t peak_index;
    float   movement_per_sec;
} orc_beat_state;
void orc_beat_reset(orc_beat_state* s);
void orc_add_beat_sample(orc_beat_state* s, int32_t w);
/* BassBeatDetector restated (beatDetector.c:51-106); bass = (int)|output[4]|. */
void orc_bass_beat(orc_beat_state* s, int32_t bass, int n, float time_delta, int strict);
/* The peak scan alone (beatDetector.c:78-103) on a given heightBuffer[161]. */
int32_t orc_peak_pick(const int32_t* hb);
/* hb[161] of a 256-entry ring (beatDetector.c:61-75). */
void orc_tempo_hb(const int32_t* ring, int strict, int32_t* hb);
void orc_tempo_mag(const int32_t* ring, double* mag);
void orc_tempo_mag_replay(const int32_t* bass, int64_t count, int64_t first_frame, double* mag_out, double* ring_norm);

/* Whole chain over frames [frame_begin, frame_end) of every channel, as
 * fftl_stft_run defines it.  Any output pointer may be NULL. */
int orc_stft(const fftl_config* cfg, const int16_t* pcm, int32_t channels,
             int64_t samples_per_channel, int64_t channel_stride, int64_t sample_stride,
             int64_t frame_begin, int64_t frame_end,
             double* bars_f64, float* bars_f32, int32_t* bars_i32, fftl_beat* beat);

/* Beat records recomputed from given per-frame (w, bass) sequences of ONE
 * channel starting at absolute frame `first_frame` with zeroed state -- lets a
 * test replay the integer stages on the GPU's own integers (bit-exact bar). */
void orc_beat_replay(const fftl_config* cfg, const int32_t* w, const int32_t* bass,
                     int64_t count, int64_t first_frame, fftl_beat* out, int32_t* hb_out /* count*161 or NULL */);

/* Synthetic PCM of SURVEY 8(d), identical to fftl_synth_pcm_device. */
void orc_synth_pcm(int16_t* pcm, int32_t channels, int64_t samples, int64_t channel_stride,
                   int64_t first_sample, float sample_rate, uint32_t seed);

/* Callers / sinks either side of the path, restated from settingsFile.c:172-247,
 * drawBar2D.cpp:33, fft_lines.c:250-255 and fft_lines.c:266-274. */
void orc_settings_parse(const char* text, long len, int32_t* v /*[11]*/, float* zoom);
uint32_t orc_range255(int number, int maxValueOfNumber);
void orc_led_byte(int32_t height, uint8_t* byte, uint8_t* valid);
void orc_marker_sweep(const float* movement, int64_t frames, float timeDelta, int32_t beatposition, int32_t direction,
                      int32_t* pos_out, int32_t* dir_out);

#ifdef __cplusplus
}
#endif
#endif

/*
 * fft_calculate.h -- drop-in for the reference header of the same name
 * (reference fft_lines/fft_calculate.h:1-11), served by libfftlines_cuda.
 *
 * Same six C symbols, same signatures, same contract as the reference
 * implementation (fft_lines/fft_calculate.c:22-52):
 *   - initializefft*(in, out) binds a forward, unnormalised complex-to-complex
 *     transform to the two CALLER-OWNED host buffers (N resp. 256 elements);
 *   - executefft*(in, out) runs the bound transform and, exactly like the
 *     reference (fft_calculate.c:27-30, :43-46), IGNORES its arguments and
 *     uses the pointers captured at initialise time;
 *   - cleanupfft*() releases the plan; it is idempotent here.
 * N is the reference's compile-time 2048 (global.h:3) unless fftl_set_n() from
 * fftlines.h was called before initializefft.
 * The functions return void as in the reference; failures (no CUDA device,
 * execute before initialise) leave `out` untouched and latch an error readable
 * through fftl_last_error() (fftlines.h).  There is no CPU fallback.
 */
#ifndef FFTL_FFT_CALCULATE_H
#define FFTL_FFT_CALCULATE_H
#include "fftl_complex.h"

#ifndef REAL
#define REAL 0   /* reference fft_calculate.h:2 */
#endif
#ifndef IMAG
#define IMAG 1   /* reference fft_calculate.h:3 */
#endif

#ifdef __cplusplus
extern "C" {
#endif

void initializefft(fftwf_complex* in, fftwf_complex* out);   /* ref fft_calculate.h:5  */
void executefft(fftwf_complex* in, fftwf_complex* out);      /* ref fft_calculate.h:6  */
void cleanupfft(void);                                       /* ref fft_calculate.h:7  */

void initializefft256(fftwf_complex* in, fftwf_complex* out);/* ref fft_calculate.h:9  */
void executefft256(fftwf_complex* in, fftwf_complex* out);   /* ref fft_calculate.h:10 */
void cleanupfft256(void);                                    /* ref fft_calculate.h:11 */

#ifdef __cplusplus
}
#endif
#endif

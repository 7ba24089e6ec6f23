/*
 * beatDetector.h -- drop-in for the reference header of the same name
 * (reference fft_lines/beatDetector.h:1-11), served by libfftlines_cuda.
 *
 *   AddBeatSample(w)   160-entry ring + running sum of the band energy,
 *                      threshold = sum/160, beatValue = (int)(w/threshold*128)
 *                      (reference beatDetector.c:30-44).
 *   GetBeatValue()     last beatValue (beatDetector.c:46-49).
 *   BassBeatDetector() pushes (int)|output[4]| into a 256-entry ring, runs the
 *                      256-point transform bound by initializefft256 over the
 *                      ring, picks the modulation peak and sets
 *                      beatMovementPerSec = N * timeDelta * peak
 *                      (beatDetector.c:51-106).  `input` is unused, as in the
 *                      reference.  beatinput/beatoutput are filled like the
 *                      reference fills them.
 * The state lives on the GPU; every call is one or two tiny kernel launches.
 * The four data symbols are defined by the library exactly as the reference
 * defines them in beatDetector.c:25-28; `timeDelta` is written by the caller
 * (reference fft_lines.c:265).
 * Where the reference has undefined behaviour the library is defined (see
 * DESIGN.md "UB made defined"): threshold 0 gives beatValue INT32_MIN (what the
 * x86 binary produces), heightBuffer[160] is computed rather than read out of
 * bounds, and "no peak found" yields index 0.
 */
#ifndef FFTL_BEATDETECTOR_H
#define FFTL_BEATDETECTOR_H
#include "fftl_complex.h"

#ifdef __cplusplus
extern "C" {
#endif

extern void AddBeatSample(int weightedAverage);               /* ref beatDetector.h:3 */
extern int  GetBeatValue(void);                               /* ref beatDetector.h:4 */
extern void BassBeatDetector(fftwf_complex* input, fftwf_complex* output,
                             fftwf_complex* beatinput,
                             fftwf_complex* beatoutput);      /* ref beatDetector.h:6 */

extern float beatMovementPerSec;                              /* ref beatDetector.h:8  */
extern int   beatposition;                                    /* ref beatDetector.h:9  */
extern float timeDelta;                                       /* ref beatDetector.h:10 */
extern int   direction;                                       /* ref beatDetector.h:11 */

#ifdef __cplusplus
}
#endif
#endif

/* headless_tick.c -- the reference's WM_CREATE / WM_TIMER / WM_DESTROY hot-path calls
 * (fft_lines/fft_lines.c:110-116, :190-236, :281, :420-422) against libfftlines_cuda,
 * on a synthetic sine instead of WASAPI.  Build: see INTEGRATION.md. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "fft_calculate.h"
#include "beatDetector.h"
#include "fftlines.h"

#define N 2048 /* global.h:3 */

int main(void)
{
    fftwf_complex *input = malloc(N * sizeof(fftwf_complex)), *output = malloc(N * sizeof(fftwf_complex));
    fftwf_complex *beatinput = malloc(256 * sizeof(fftwf_complex)), *beatoutput = malloc(256 * sizeof(fftwf_complex));
    int16_t audio[N];
    int barCount = 100, heights[100];
    float zoom = 0.0001f;
    initializefft(input, output);
    initializefft256(beatinput, beatoutput);
    if (fftl_last_error()) { fprintf(stderr, "%s\n", fftl_last_error_string()); return 1; }
    for (int tick = 0; tick < 5; tick++) {
        for (int i = 0; i < N; i++) audio[i] = (int16_t)lrint(10000.0 * sin(2.0 * M_PI * 5.0 * (i + 512 * tick) / N));
        for (int i = 0; i < N; i++) { input[i][REAL] = (float)audio[i]; input[i][IMAG] = 0; }
        executefft(input, output);
        for (int i = 0; i < barCount; i++)
            heights[i] = (int)(sqrt(pow(output[i][REAL], 2) + pow(output[i][IMAG], 2)) * zoom);
        timeDelta = 512.0f / 48000.0f;
        BassBeatDetector(input, output, beatinput, beatoutput);
        AddBeatSample((heights[3] + heights[4] + heights[5] + heights[6]) / 4);
        printf("tick %d: height[5]=%d beatValue=%d beatMovementPerSec=%g\n", tick, heights[5], GetBeatValue(), beatMovementPerSec);
    }
    cleanupfft();
    cleanupfft256();
    free(input); free(output); free(beatinput); free(beatoutput);
    return fftl_last_error();
}

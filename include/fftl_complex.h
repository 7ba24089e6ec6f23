/*
 * fftl_complex.h -- the one FFTW type the reference's hot-path headers use.
 *
 * The reference includes <fftw3.h> only for `fftwf_complex`, which FFTW defines
 * as `float[2]` when <complex.h> is not in play (reference
 * fft_lines/fftw/include/fftw3.h:62,359).  libfftlines_cuda does not need FFTW:
 * if the application already included fftw3.h we use its typedef, otherwise we
 * provide a layout-identical one.
 */
#ifndef FFTL_COMPLEX_H
#define FFTL_COMPLEX_H
#if !defined(FFTW3_H) && !defined(FFTL_HAVE_FFTWF_COMPLEX)
#define FFTL_HAVE_FFTWF_COMPLEX 1
typedef float fftwf_complex[2];
#endif
#endif

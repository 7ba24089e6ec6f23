/*
 * fftwf_shim.c -- from-scratch single-precision FFT that exports the four
 * FFTW symbols the reference's hot path links against
 * (reference fft_lines/fft_calculate.c:24,29,34,40,45,50; fft_lines.c:422).
 *
 * TEST INFRASTRUCTURE ONLY.  The reference links a prebuilt Windows
 * libfftw3f-3.dll that is not in its tree (version not pinned); no host FFTW
 * exists in this image.  This file is the FFT *provider* for oracle/_ref (the
 * unmodified reference C files) and for the CPU baseline; every report labels
 * it "fftwf-shim", never "fftwf".
 *
 * Algorithm: Stockham autosort, radix-4 passes plus one radix-2 pass when
 * log2(n) is odd, twiddles precomputed in double.  Forward, unnormalised,
 * out-of-place or in-place (a scratch buffer is used either way).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef float cpx[2];

struct fftwf_plan_s {
    int n;
    int sign;
    cpx* in;
    cpx* out;
    int npass;
    int radix[16];
    float* tw[16]; /* per pass: Ns * (R-1) complex twiddles, [k][r-1] */
    cpx* scratch;
};
typedef struct fftwf_plan_s* fftwf_plan;

/* Lets the harness run the unmodified reference wrappers at a transform size
 * other than the reference's compile-time N=2048 (global.h:3): when non-zero,
 * a plan requested with n == 2048 is built with this size instead. */
int fftl_shim_override_n = 0;

static void pass_r4(int n, int Ns, const float* tw, const cpx* x, cpx* y, float fwd)
{
    int q = n / 4;
    for (int blk = 0; blk < q / Ns; blk++) {
        const cpx* x0 = x + blk * Ns;
        cpx* y0 = y + blk * Ns * 4;
        for (int k = 0; k < Ns; k++) {
            float ar = x0[k][0], ai = x0[k][1];
            float br = x0[k + q][0], bi = x0[k + q][1];
            float cr = x0[k + 2 * q][0], ci = x0[k + 2 * q][1];
            float dr = x0[k + 3 * q][0], di = x0[k + 3 * q][1];
            const float* t = tw + 6 * k;
            float t1r = t[0], t1i = t[1], t2r = t[2], t2i = t[3], t3r = t[4], t3i = t[5];
            float b2r = br * t1r - bi * t1i, b2i = br * t1i + bi * t1r;
            float c2r = cr * t2r - ci * t2i, c2i = cr * t2i + ci * t2r;
            float d2r = dr * t3r - di * t3i, d2i = dr * t3i + di * t3r;
            float s0r = ar + c2r, s0i = ai + c2i, s1r = ar - c2r, s1i = ai - c2i;
            float s2r = b2r + d2r, s2i = b2i + d2i, s3r = b2r - d2r, s3i = b2i - d2i;
            /* forward (fwd=+1): multiply s3 by -i; backward: by +i */
            float m3r = fwd * s3i, m3i = -fwd * s3r;
            y0[k][0] = s0r + s2r;            y0[k][1] = s0i + s2i;
            y0[k + Ns][0] = s1r + m3r;       y0[k + Ns][1] = s1i + m3i;
            y0[k + 2 * Ns][0] = s0r - s2r;   y0[k + 2 * Ns][1] = s0i - s2i;
            y0[k + 3 * Ns][0] = s1r - m3r;   y0[k + 3 * Ns][1] = s1i - m3i;
        }
    }
}

static void pass_r2(int n, int Ns, const float* tw, const cpx* x, cpx* y)
{
    int q = n / 2;
    for (int blk = 0; blk < q / Ns; blk++) {
        const cpx* x0 = x + blk * Ns;
        cpx* y0 = y + blk * Ns * 2;
        for (int k = 0; k < Ns; k++) {
            float ar = x0[k][0], ai = x0[k][1];
            float br = x0[k + q][0], bi = x0[k + q][1];
            float tr = tw[2 * k], ti = tw[2 * k + 1];
            float b2r = br * tr - bi * ti, b2i = br * ti + bi * tr;
            y0[k][0] = ar + b2r;       y0[k][1] = ai + b2i;
            y0[k + Ns][0] = ar - b2r;  y0[k + Ns][1] = ai - b2i;
        }
    }
}

fftwf_plan fftwf_plan_dft_1d(int n, cpx* in, cpx* out, int sign, unsigned flags)
{
    (void)flags;
    if (fftl_shim_override_n > 0 && n == 2048) n = fftl_shim_override_n;
    if (n < 1 || (n & (n - 1))) return NULL;
    fftwf_plan p = (fftwf_plan)calloc(1, sizeof(*p));
    if (!p) return NULL;
    p->n = n;
    p->sign = sign;
    p->in = in;
    p->out = out;
    p->scratch = (cpx*)malloc(sizeof(cpx) * (size_t)n * 2);
    int lg = 0;
    while ((1 << lg) < n) lg++;
    int np = 0;
    for (int i = 0; i < lg / 2; i++) p->radix[np++] = 4;
    if (lg & 1) p->radix[np++] = 2;
    p->npass = np;
    int Ns = 1;
    double sg = sign < 0 ? -1.0 : 1.0;
    for (int s = 0; s < np; s++) {
        int R = p->radix[s];
        p->tw[s] = (float*)malloc(sizeof(float) * 2 * (size_t)Ns * (R - 1));
        for (int k = 0; k < Ns; k++)
            for (int r = 1; r < R; r++) {
                double a = sg * 2.0 * M_PI * (double)k * (double)r / (double)(Ns * R);
                p->tw[s][2 * (k * (R - 1) + (r - 1))] = (float)cos(a);
                p->tw[s][2 * (k * (R - 1) + (r - 1)) + 1] = (float)sin(a);
            }
        Ns *= R;
    }
    return p;
}

/* Execute on explicit buffers (used by the CPU port with per-thread plans). */
void fftl_shim_execute_dft(const fftwf_plan p, cpx* in, cpx* out)
{
    int n = p->n;
    cpx* a = p->scratch;
    cpx* b = p->scratch + n;
    const cpx* src = in;
    int Ns = 1;
    if (p->npass == 0) { out[0][0] = in[0][0]; out[0][1] = in[0][1]; return; }
    for (int s = 0; s < p->npass; s++) {
        int R = p->radix[s];
        cpx* dst = (s == p->npass - 1) ? out : ((s & 1) ? b : a);
        if (dst == (cpx*)src) dst = (dst == a) ? b : a; /* never alias */
        if (s == p->npass - 1 && out == in && src == (const cpx*)in) {
            /* single pass in place: go through scratch */
            dst = a;
        }
        if (R == 4) pass_r4(n, Ns, p->tw[s], src, dst, p->sign < 0 ? 1.0f : -1.0f);
        else pass_r2(n, Ns, p->tw[s], src, dst);
        src = dst;
        Ns *= R;
    }
    if ((const cpx*)src != (const cpx*)out) memcpy(out, src, sizeof(cpx) * (size_t)n);
}

void fftwf_execute(const fftwf_plan p)
{
    if (p) fftl_shim_execute_dft(p, p->in, p->out);
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    for (int s = 0; s < p->npass; s++) free(p->tw[s]);
    free(p->scratch);
    free(p);
}

void fftwf_cleanup(void) {}

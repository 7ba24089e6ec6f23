/*
 * cpu_port.c -- single-precision, multi-threaded CPU port of the reference's
 * per-tick chain, generalised to batched STFT with the extended stages.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (bench.py cpu_baseline and
 * `--impl reference`; tests pin it against oracle/_ref).  Never linked into
 * the product.
 *
 * It keeps the reference's arithmetic shape so that on the reference's own
 * configuration (N=2048, RECT, LINEAR, no SG) it is bit-identical to the
 * unmodified reference files run by ref_harness.c with the same FFT provider:
 *   fft_lines.c:190-194   (float)int16 into a complex buffer, imag 0
 *   fft_calculate.c:29    c2c forward FFT, fp32 (provider: oracle/fftwf_shim.c)
 *   fft_lines.c:202       (int)(sqrt(pow(re,2)+pow(im,2)) * zoom), double math
 *   fft_lines.c:281 + beatDetector.c:30-44, :51-106   beat chain
 * The reference is single-threaded with global state; here each worker owns
 * a plan and a contiguous frame range (with the 255-frame beat warm-up the
 * GPU path uses too), which is how SURVEY 8(d) says to use all host cores.
 */
#include "fftl_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

typedef float cpx[2];
typedef struct fftwf_plan_s* fftwf_plan;
fftwf_plan fftwf_plan_dft_1d(int n, cpx* in, cpx* out, int sign, unsigned flags);
void fftl_shim_execute_dft(const fftwf_plan p, cpx* in, cpx* out);
void fftwf_destroy_plan(fftwf_plan p);

int port_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct worker {
    fftwf_plan plan, plan256;
    cpx *in, *out, *bin, *bout;
    float* raw;
    int32_t* h;
} worker;

static int worker_init(worker* w, const fftl_config* c)
{
    w->in = (cpx*)malloc(sizeof(cpx) * c->n);
    w->out = (cpx*)malloc(sizeof(cpx) * c->n);
    w->bin = (cpx*)malloc(sizeof(cpx) * FFTL_TEMPO_RING);
    w->bout = (cpx*)malloc(sizeof(cpx) * FFTL_TEMPO_RING);
    w->raw = (float*)malloc(sizeof(float) * c->bars);
    w->h = (int32_t*)malloc(sizeof(int32_t) * c->bars);
    w->plan = fftwf_plan_dft_1d(c->n, w->in, w->out, -1, 0);
    w->plan256 = fftwf_plan_dft_1d(FFTL_TEMPO_RING, w->bin, w->bout, -1, 0);
    return (w->in && w->out && w->bin && w->bout && w->raw && w->h && w->plan && w->plan256) ? 0 : -1;
}
static void worker_free(worker* w)
{
    fftwf_destroy_plan(w->plan); fftwf_destroy_plan(w->plan256);
    free(w->in); free(w->out); free(w->bin); free(w->bout); free(w->raw); free(w->h);
}

typedef struct job {
    const fftl_config* c; const int16_t* pcm; int64_t cs, ss, f0, f1, nf, per, ranges, items;
    const float *win, *sg; const int32_t *lo, *hi;
    float* bars_f32; int32_t* bars_i32; fftl_beat* beat;
    worker* W; int64_t next; pthread_mutex_t mu;
} job;
typedef struct targ { job* j; int tid; } targ;

static void* port_thread(void* arg)
{
    job* J = ((targ*)arg)->j;
    worker* w = &J->W[((targ*)arg)->tid];
    const fftl_config* c = J->c;
    const int16_t* pcm = J->pcm;
    const float *win = J->win, *sg = J->sg;
    const int32_t *lo = J->lo, *hi = J->hi;
    float* bars_f32 = J->bars_f32; int32_t* bars_i32 = J->bars_i32; fftl_beat* beat = J->beat;
    int64_t cs = J->cs, ss = J->ss, f0 = J->f0, f1 = J->f1, nf = J->nf, per = J->per, ranges = J->ranges;
    int n = c->n, B = c->bars, sw = c->sg_half, m = 2 * sw + 1;
    float td = (float)c->hop / c->sample_rate;
    int rect = (c->window == FFTL_WINDOW_RECT);
    int identity = (c->binning == FFTL_BINNING_LINEAR);
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int64_t it = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (it >= J->items) break;
        {
            int ch = (int)(it / ranges);
            int64_t a = f0 + (it % ranges) * per, b = a + per;
            if (b > f1) b = f1;
            int64_t warm = a - (FFTL_TEMPO_RING - 1);
            if (warm < 0) warm = 0;
            if (!beat) warm = a;
            orc_beat_state st;
            orc_beat_reset(&st);
            st.tempo_pos = (int32_t)(warm % FFTL_TEMPO_RING);
            for (int64_t f = warm; f < b; f++) {
                const int16_t* x = pcm + ch * cs + f * c->hop * ss;
                if (rect) for (int i = 0; i < n; i++) { w->in[i][0] = (float)x[i * ss]; w->in[i][1] = 0; }
                else for (int i = 0; i < n; i++) { w->in[i][0] = (float)x[i * ss] * win[i]; w->in[i][1] = 0; }
                fftl_shim_execute_dft(w->plan, w->in, w->out);
                int emit = (f >= a);
                float* fout = (emit && bars_f32) ? bars_f32 + ((int64_t)ch * nf + (f - f0)) * B : NULL;
                if (identity && sw == 0) {
                    /* reference path, fft_lines.c:199-208, same double expression */
                    for (int i = 0; i < B; i++) {
                        double mag = sqrt(pow(w->out[i][0], 2) + pow(w->out[i][1], 2));
                        w->h[i] = orc_trunc_i32(mag * c->zoom, c->strict_int);
                        if (c->circle) w->h[i] += 10;
                        if (fout) fout[i] = (float)(mag * c->zoom + (c->circle ? 10.0 : 0.0));
                    }
                } else {
                    for (int i = 0; i < B; i++) {
                        float s = 0;
                        for (int k = lo[i]; k < hi[i]; k++) {
                            float re = w->out[k][0], im = w->out[k][1];
                            s += sqrtf(re * re + im * im);
                        }
                        w->raw[i] = s / (float)(hi[i] - lo[i]);
                    }
                    for (int i = 0; i < B; i++) {
                        float v;
                        if (sw == 0) v = w->raw[i];
                        else {
                            int row, base;
                            if (i < sw) { row = i; base = 0; }
                            else if (i >= B - sw) { row = m - (B - i); base = B - m; }
                            else { row = sw; base = i - sw; }
                            v = 0;
                            for (int k = 0; k < m; k++) v += sg[row * m + k] * w->raw[base + k];
                        }
                        double hgt = (double)v * (double)c->zoom;
                        w->h[i] = orc_trunc_i32(hgt, c->strict_int);
                        if (c->circle) { w->h[i] += 10; hgt += 10.0; }
                        if (fout) fout[i] = (float)hgt;
                    }
                }
                if (emit && bars_i32) memcpy(bars_i32 + ((int64_t)ch * nf + (f - f0)) * B, w->h, sizeof(int32_t) * B);
                if (beat) {
                    /* fft_lines.c:281 */
                    int32_t wv = (w->h[c->beat_bars[0]] + w->h[c->beat_bars[1]] + w->h[c->beat_bars[2]] + w->h[c->beat_bars[3]]) / 4;
                    double bre = w->out[c->bass_bin][0], bim = w->out[c->bass_bin][1];
                    int32_t bass = orc_trunc_i32(sqrt(pow(bre, 2) + pow(bim, 2)), c->strict_int);
                    orc_add_beat_sample(&st, wv);
                    /* beatDetector.c:53-75 with the fp32 256-pt provider */
                    st.tempo_ring[st.tempo_pos] = bass;
                    if (++st.tempo_pos >= FFTL_TEMPO_RING) st.tempo_pos = 0;
                    for (int i = 0; i < FFTL_TEMPO_RING; i++) { w->bin[i][0] = (float)st.tempo_ring[i]; w->bin[i][1] = 0; }
                    fftl_shim_execute_dft(w->plan256, w->bin, w->bout);
                    for (int i = 0; i <= FFTL_BEAT_SAMPLES; i++)
                        st.hb[i] = orc_trunc_i32(sqrt(pow(w->bout[i][0], 2) + pow(w->bout[i][1], 2)), c->strict_int);
                    st.peak_index = orc_peak_pick(st.hb);
                    st.movement_per_sec = ((float)n * td) * (float)st.peak_index;
                    if (emit) {
                        fftl_beat* o = &beat[(int64_t)ch * nf + (f - f0)];
                        memset(o, 0, sizeof(*o));
                        o->w = wv; o->thr = st.thr; o->beat_value = st.beat_value;
                        o->flag = (st.thr > 0 && wv > st.thr) ? 1 : 0;
                        o->bass = bass; o->peak_index = st.peak_index;
                        o->movement_per_sec = st.movement_per_sec;
                    }
                }
            }
        }
    }
    return NULL;
}

/* Run frames [f0,f1) of `channels` channels of planar/strided PCM on `threads`
 * workers.  Outputs as fftl_stft_run; any may be NULL.  Returns seconds spent
 * in the parallel frame loop (table setup and plan creation excluded, as the
 * reference's FFTW_MEASURE planning would be). */
double port_run(const fftl_config* c, const int16_t* pcm, int32_t channels, int64_t S, int64_t cs,
                int64_t ss, int64_t f0, int64_t f1, int threads, float* bars_f32,
                int32_t* bars_i32, fftl_beat* beat)
{
    if (orc_config_validate(c) != FFTL_OK) return -1.0;
    int64_t F = orc_num_frames(c, S);
    if (f0 < 0 || f1 < f0 || f1 > F) return -1.0;
    int n = c->n, B = c->bars, sw = c->sg_half, m = 2 * sw + 1;
    int64_t nf = f1 - f0;
    float* win = (float*)malloc(sizeof(float) * n);
    int32_t* lo = (int32_t*)malloc(sizeof(int32_t) * B);
    int32_t* hi = (int32_t*)malloc(sizeof(int32_t) * B);
    double* sgd = (double*)malloc(sizeof(double) * m * m);
    float* sg = (float*)malloc(sizeof(float) * m * m);
    double* wd = (double*)malloc(sizeof(double) * n);
    orc_window(c, wd);
    for (int i = 0; i < n; i++) win[i] = (float)wd[i];
    free(wd);
    orc_build_bins(c, lo, hi);
    if (sw > 0) {
        orc_build_sg(sw, c->sg_degree, sgd);
        for (int i = 0; i < m * m; i++) sg[i] = (float)sgd[i];
    }
    if (threads < 1) threads = 1;
    /* work items: (channel, contiguous frame range) */
    int64_t per = (nf + threads - 1) / threads;
    if (per < 1) per = 1;
    int64_t ranges = (nf + per - 1) / per;
    int64_t items = ranges * channels;
    worker* W = (worker*)calloc((size_t)threads, sizeof(worker));
    int bad = 0;
    for (int t = 0; t < threads; t++) bad |= worker_init(&W[t], c);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (!bad) {
        job J;
        memset(&J, 0, sizeof(J));
        J.c = c; J.pcm = pcm; J.cs = cs; J.ss = ss; J.f0 = f0; J.f1 = f1; J.nf = nf; J.per = per;
        J.ranges = ranges; J.items = items; J.win = win; J.sg = sg; J.lo = lo; J.hi = hi;
        J.bars_f32 = bars_f32; J.bars_i32 = bars_i32; J.beat = beat; J.W = W; J.next = 0;
        pthread_mutex_init(&J.mu, NULL);
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
        targ* ta = (targ*)malloc(sizeof(targ) * threads);
        for (int t = 0; t < threads; t++) { ta[t].j = &J; ta[t].tid = t; }
        for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, port_thread, &ta[t]);
        port_thread(&ta[0]);
        for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
        free(th); free(ta);
        pthread_mutex_destroy(&J.mu);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int t = 0; t < threads; t++) worker_free(&W[t]);
    free(W); free(win); free(lo); free(hi); free(sgd); free(sg);
    if (bad) return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/*
 * fftl_oracle.c -- CPU restatement (double precision) of the reference's
 * spectrum hot path.  TEST INFRASTRUCTURE ONLY; see fftl_oracle.h.
 *
 * Reference lines restated (all under /root/reference/fft_lines/):
 *   fft_lines.c:190-194   int16 -> complex input, imag = 0, no window
 *   fft_calculate.c:24,29 forward, unnormalised c2c DFT of size N
 *   fft_lines.c:199-208   height = (int)(sqrt(re^2+im^2) * zoom) (+10 circle)
 *   fft_lines.c:281       w = (h3+h4+h5+h6)/4  -> AddBeatSample
 *   beatDetector.c:30-44  AddBeatSample ring / threshold / beatValue
 *   beatDetector.c:51-106 BassBeatDetector ring / 256-pt DFT / peak scan
 * Extended stages (Hann, LOG bins, Savitzky-Golay) are defined here.
 */
#include "fftl_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ casts */
int32_t orc_trunc_i32(double x, int strict)
{
    if (x != x) return strict ? INT32_MIN : 0;
    if (x >= 2147483648.0) return strict ? INT32_MIN : INT32_MAX;
    if (x <= -2147483649.0) return INT32_MIN;
    return (int32_t)x; /* truncation toward zero, in range */
}

static int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

/* -------------------------------------------------------------------- DFT */
void orc_dft_naive(int n, const double* ir, const double* ii, double* or_, double* oi)
{
    for (int k = 0; k < n; k++) {
        long double sr = 0, si = 0;
        for (int t = 0; t < n; t++) {
            /* exact angle reduction: (k*t) mod n */
            long long m = ((long long)k * t) % n;
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / n;
            long double c = cosl(a), s = sinl(a);
            long double xr = ir[t], xi = ii ? ii[t] : 0.0;
            sr += xr * c - xi * s;
            si += xr * s + xi * c;
        }
        or_[k] = (double)sr;
        oi[k] = (double)si;
    }
}

void orc_fft(int n, const double* ir, const double* ii, double* or_, double* oi)
{
    int lg = 0;
    while ((1 << lg) < n) lg++;
    /* bit-reversed copy */
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < lg; b++) r |= ((i >> b) & 1) << (lg - 1 - b);
        or_[r] = ir[i];
        oi[r] = ii ? ii[i] : 0.0;
    }
    double* wr = (double*)malloc(sizeof(double) * (n / 2 + 1));
    double* wi = (double*)malloc(sizeof(double) * (n / 2 + 1));
    for (int k = 0; k < n / 2; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        wr[k] = cos(a);
        wi[k] = sin(a);
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int s = 0; s < n; s += len) {
            for (int j = 0; j < half; j++) {
                double cr = wr[j * step], ci = wi[j * step];
                int a = s + j, b = a + half;
                double tr = or_[b] * cr - oi[b] * ci;
                double ti = or_[b] * ci + oi[b] * cr;
                or_[b] = or_[a] - tr;
                oi[b] = oi[a] - ti;
                or_[a] += tr;
                oi[a] += ti;
            }
        }
    }
    free(wr);
    free(wi);
}

/* ------------------------------------------------------------- validation */
static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

int orc_config_validate(const fftl_config* c)
{
    if (!c) return FFTL_ERR_ARG;
    if (!is_pow2(c->n) || c->n < 256 || c->n > 16384) return FFTL_ERR_ARG;
    if (c->hop < 1) return FFTL_ERR_ARG;
    if (c->bars < 1 || c->bars > FFTL_MAX_BARS) return FFTL_ERR_ARG;
    if (c->window != FFTL_WINDOW_RECT && c->window != FFTL_WINDOW_HANN) return FFTL_ERR_ARG;
    if (c->binning == FFTL_BINNING_LINEAR) {
        if (c->bars > c->n / 2 + 1) return FFTL_ERR_ARG;
    } else if (c->binning == FFTL_BINNING_LOG) {
        if (!(c->sample_rate > 0) || !(c->fmin > 0) || !(c->fmax > c->fmin)) return FFTL_ERR_ARG;
        if (c->fmax > 0.5f * c->sample_rate) return FFTL_ERR_ARG;
    } else return FFTL_ERR_ARG;
    if (c->sg_half < 0 || c->sg_half > FFTL_MAX_SG_HALF) return FFTL_ERR_ARG;
    if (c->sg_half > 0) {
        if (c->sg_degree < 0 || c->sg_degree > 2 * c->sg_half) return FFTL_ERR_ARG;
        if (c->bars < 2 * c->sg_half + 1) return FFTL_ERR_ARG;
    }
    for (int i = 0; i < 4; i++)
        if (c->beat_bars[i] < 0 || c->beat_bars[i] >= c->bars) return FFTL_ERR_ARG;
    if (c->bass_bin < 0 || c->bass_bin > c->n / 2) return FFTL_ERR_ARG;
    if (!(c->sample_rate > 0)) return FFTL_ERR_ARG;
    return FFTL_OK;
}

int64_t orc_num_frames(const fftl_config* c, int64_t samples)
{
    if (samples < c->n) return 0;
    return (samples - c->n) / c->hop + 1;
}

/* ----------------------------------------------------------------- tables */
int orc_build_bins(const fftl_config* c, int32_t* lo, int32_t* hi)
{
    int nb = c->n / 2 + 1; /* usable bins 0..n/2 */
    if (c->binning == FFTL_BINNING_LINEAR) {
        /* reference: bar i = bin i (fft_lines.c:199-203) */
        for (int b = 0; b < c->bars; b++) { lo[b] = b; hi[b] = b + 1; }
        return FFTL_OK;
    }
    /* LOG: geometric edges e_b = fmin*(fmax/fmin)^(b/B); bar b averages bins
     * floor(e_b/df) .. floor(e_{b+1}/df)-1, at least one bin; bars may share
     * a bin at the low end. */
    double ratio = (double)c->fmax / (double)c->fmin;
    double scale = (double)c->n / (double)c->sample_rate;
    for (int b = 0; b < c->bars; b++) {
        double e0 = (double)c->fmin * pow(ratio, (double)b / (double)c->bars);
        double e1 = (double)c->fmin * pow(ratio, (double)(b + 1) / (double)c->bars);
        int l = (int)floor(e0 * scale);
        int h = (int)floor(e1 * scale);
        if (l < 0) l = 0;
        if (l > nb - 1) l = nb - 1;
        if (h < l + 1) h = l + 1;
        if (h > nb) h = nb;
        lo[b] = l;
        hi[b] = h;
    }
    return FFTL_OK;
}

/* Savitzky-Golay projection matrix P = A (A^T A)^-1 A^T over x = -w..w.
 * Row p evaluates the least-squares polynomial at window position p; row w is
 * the usual symmetric smoothing kernel; rows 0..w-1 / w+1..2w are used for the
 * first / last w bars (same convention as scipy savgol_filter mode='interp'). */
int orc_build_sg(int w, int d, double* coef)
{
    int m = 2 * w + 1, p = d + 1;
    if (w < 1 || d < 0 || p > m) return FFTL_ERR_ARG;
    long double A[2 * FFTL_MAX_SG_HALF + 1][2 * FFTL_MAX_SG_HALF + 1];
    long double G[2 * FFTL_MAX_SG_HALF + 1][2 * (2 * FFTL_MAX_SG_HALF + 1)];
    for (int i = 0; i < m; i++) {
        long double x = (long double)(i - w) / (long double)(w > 0 ? w : 1); /* scaled for conditioning */
        long double v = 1;
        for (int j = 0; j < p; j++) { A[i][j] = v; v *= x; }
    }
    /* G = [A^T A | I] -> Gauss-Jordan */
    for (int i = 0; i < p; i++) {
        for (int j = 0; j < p; j++) {
            long double s = 0;
            for (int k = 0; k < m; k++) s += A[k][i] * A[k][j];
            G[i][j] = s;
        }
        for (int j = 0; j < p; j++) G[i][p + j] = (i == j) ? 1 : 0;
    }
    for (int col = 0; col < p; col++) {
        int piv = col;
        for (int r = col + 1; r < p; r++)
            if (fabsl(G[r][col]) > fabsl(G[piv][col])) piv = r;
        if (fabsl(G[piv][col]) < 1e-30L) return FFTL_ERR_ARG;
        if (piv != col)
            for (int j = 0; j < 2 * p; j++) { long double t = G[col][j]; G[col][j] = G[piv][j]; G[piv][j] = t; }
        long double inv = 1 / G[col][col];
        for (int j = 0; j < 2 * p; j++) G[col][j] *= inv;
        for (int r = 0; r < p; r++) {
            if (r == col) continue;
            long double f = G[r][col];
            if (f == 0) continue;
            for (int j = 0; j < 2 * p; j++) G[r][j] -= f * G[col][j];
        }
    }
    /* P[i][k] = sum_{a,b} A[i][a] * Ginv[a][b] * A[k][b] */
    for (int i = 0; i < m; i++)
        for (int k = 0; k < m; k++) {
            long double s = 0;
            for (int a = 0; a < p; a++)
                for (int b = 0; b < p; b++) s += A[i][a] * G[a][p + b] * A[k][b];
            coef[i * m + k] = (double)s;
        }
    return FFTL_OK;
}

void orc_window(const fftl_config* c, double* win)
{
    for (int i = 0; i < c->n; i++) {
        if (c->window == FFTL_WINDOW_HANN) /* periodic Hann */
            win[i] = 0.5 - 0.5 * cos(2.0 * M_PI * (double)i / (double)c->n);
        else
            win[i] = 1.0; /* reference: no window (fft_lines.c:190-194) */
    }
}

/* ------------------------------------------------------------------ frame */
static void frame_core(const fftl_config* c, const int16_t* pcm, int64_t ss, const double* win,
                       const int32_t* lo, const int32_t* hi, const double* sg, double* scratch,
                       double* mag_out, double* height, int32_t* bars_i32, double* bass_mag)
{
    int n = c->n, nb = n / 2 + 1, B = c->bars;
    double* xr = scratch;          /* n */
    double* Xr = scratch + n;      /* n */
    double* Xi = scratch + 2 * n;  /* n */
    double* mag = scratch + 3 * n; /* nb */
    double* raw = mag + nb;        /* B */
    for (int i = 0; i < n; i++) xr[i] = (double)pcm[(int64_t)i * ss] * win[i]; /* fft_lines.c:192 */
    orc_fft(n, xr, NULL, Xr, Xi);                                             /* fft_lines.c:196 */
    for (int k = 0; k < nb; k++) mag[k] = sqrt(Xr[k] * Xr[k] + Xi[k] * Xi[k]); /* fft_lines.c:202 */
    if (mag_out) memcpy(mag_out, mag, sizeof(double) * nb);
    if (bass_mag) *bass_mag = mag[c->bass_bin];
    for (int b = 0; b < B; b++) {
        double s = 0;
        for (int k = lo[b]; k < hi[b]; k++) s += mag[k];
        raw[b] = s / (double)(hi[b] - lo[b]);
    }
    int w = c->sg_half, m = 2 * w + 1;
    double zoom = (double)c->zoom; /* float zoom promoted in the reference's double expression */
    for (int b = 0; b < B; b++) {
        double v;
        if (w == 0) v = raw[b];
        else {
            int row, base;
            if (b < w) { row = b; base = 0; }
            else if (b >= B - w) { row = m - (B - b); base = B - m; }
            else { row = w; base = b - w; }
            v = 0;
            for (int k = 0; k < m; k++) v += sg[row * m + k] * raw[base + k];
        }
        double h = v * zoom;
        int32_t hi32 = orc_trunc_i32(h, c->strict_int);
        if (c->circle) { h += 10.0; hi32 = wrap_add(hi32, 10); } /* fft_lines.c:204-207 */
        if (height) height[b] = h;
        if (bars_i32) bars_i32[b] = hi32;
    }
}

typedef struct orc_tables {
    double* win; int32_t* lo; int32_t* hi; double* sg; double* scratch;
} orc_tables;

static int tables_make(const fftl_config* c, orc_tables* t)
{
    int n = c->n, m = 2 * c->sg_half + 1;
    t->win = (double*)malloc(sizeof(double) * n);
    t->lo = (int32_t*)malloc(sizeof(int32_t) * c->bars);
    t->hi = (int32_t*)malloc(sizeof(int32_t) * c->bars);
    t->sg = (double*)malloc(sizeof(double) * m * m);
    t->scratch = (double*)malloc(sizeof(double) * (3 * n + n / 2 + 1 + c->bars));
    if (!t->win || !t->lo || !t->hi || !t->sg || !t->scratch) return FFTL_ERR_NOMEM;
    orc_window(c, t->win);
    orc_build_bins(c, t->lo, t->hi);
    if (c->sg_half > 0) {
        int rc = orc_build_sg(c->sg_half, c->sg_degree, t->sg);
        if (rc) return rc;
    }
    return FFTL_OK;
}
static void tables_free(orc_tables* t)
{
    free(t->win); free(t->lo); free(t->hi); free(t->sg); free(t->scratch);
}

void orc_frame(const fftl_config* c, const int16_t* pcm, int64_t ss, double* mag, double* height,
               int32_t* bars_i32)
{
    orc_tables t;
    if (tables_make(c, &t) == FFTL_OK)
        frame_core(c, pcm, ss, t.win, t.lo, t.hi, t.sg, t.scratch, mag, height, bars_i32, NULL);
    tables_free(&t);
}

/* ------------------------------------------------------------------- beat */
void orc_beat_reset(orc_beat_state* s) { memset(s, 0, sizeof(*s)); }

void orc_add_beat_sample(orc_beat_state* s, int32_t w)
{
    /* beatDetector.c:32-37 */
    s->sum = wrap_add(s->sum, w);
    s->sum = wrap_sub(s->sum, s->ring[s->pos]);
    s->ring[s->pos] = w;
    s->pos++;
    if (s->pos == FFTL_BEAT_SAMPLES) s->pos = 0;
    /* beatDetector.c:41 (C division truncates toward zero) */
    s->thr = s->sum / FFTL_BEAT_SAMPLES;
    /* beatDetector.c:43: float division, *128, (int).  thr == 0 gives inf/NaN
     * whose (int) is INT32_MIN on x86 -- defined here as exactly that. */
    float q = (float)w / (float)s->thr; /* IEEE: x/0 = inf, 0/0 = NaN */
    float v = q * 128.0f;
    s->beat_value = orc_trunc_i32((double)v, 1);
}

void orc_tempo_hb(const int32_t* ring, int strict, int32_t* hb)
{
    double ir[FFTL_TEMPO_RING], orr[FFTL_TEMPO_RING], oi[FFTL_TEMPO_RING];
    for (int i = 0; i < FFTL_TEMPO_RING; i++) ir[i] = (double)(float)ring[i]; /* beatDetector.c:63 */
    orc_fft(FFTL_TEMPO_RING, ir, NULL, orr, oi);                              /* beatDetector.c:67 */
    for (int i = 0; i <= FFTL_BEAT_SAMPLES; i++)                              /* :70-75, +[160]  */
        hb[i] = orc_trunc_i32(sqrt(orr[i] * orr[i] + oi[i] * oi[i]), strict);
}

/* The magnitudes of beatDetector.c:73 BEFORE the (int) cast, bins 0..160, in double: what the parity tests use
 * to tell an integer-boundary flip of the peak scan from a wrong transform. */
void orc_tempo_mag(const int32_t* ring, double* mag)
{
    double ir[FFTL_TEMPO_RING], orr[FFTL_TEMPO_RING], oi[FFTL_TEMPO_RING];
    for (int i = 0; i < FFTL_TEMPO_RING; i++) ir[i] = (double)(float)ring[i];
    orc_fft(FFTL_TEMPO_RING, ir, NULL, orr, oi);
    for (int i = 0; i <= FFTL_BEAT_SAMPLES; i++) mag[i] = sqrt(orr[i] * orr[i] + oi[i] * oi[i]);
}

/* Tempo-ring magnitudes (orc_tempo_mag, 161 doubles) after each of `count` frames whose bass values are given,
 * the ring starting zeroed at slot first_frame % 256; ring_norm[f] = Euclidean norm of the ring (may be NULL). */
void orc_tempo_mag_replay(const int32_t* bass, int64_t count, int64_t first_frame, double* mag_out, double* ring_norm)
{
    int32_t ring[FFTL_TEMPO_RING];
    memset(ring, 0, sizeof(ring));
    int pos = (int)(first_frame % FFTL_TEMPO_RING);
    for (int64_t f = 0; f < count; f++) {
        ring[pos] = bass[f];
        pos = (pos + 1) % FFTL_TEMPO_RING;
        orc_tempo_mag(ring, mag_out + f * (FFTL_BEAT_SAMPLES + 1));
        if (ring_norm) {
            double s2 = 0;
            for (int i = 0; i < FFTL_TEMPO_RING; i++) s2 += (double)ring[i] * (double)ring[i];
            ring_norm[f] = sqrt(s2);
        }
    }
}

int32_t orc_peak_pick(const int32_t* hb)
{
    /* beatDetector.c:78-103 */
    int32_t diffnew = 0, diffold = 0;
    int32_t peaks[30];
    int peaknumber = 0, highestpeakvalue = 0, highestpeaknumber = 0;
    memset(peaks, 0, sizeof(peaks)); /* reference leaves this uninitialised */
    for (int i = 0; i < FFTL_BEAT_SAMPLES; i++) {
        diffold = diffnew;
        diffnew = wrap_sub(hb[i + 1], hb[i]);
        if (diffold >= 0 && diffnew <= 0) {
            peaks[peaknumber] = i;
            peaknumber++;
            if (hb[i] > highestpeakvalue && i > 30) {
                highestpeakvalue = hb[i];
                highestpeaknumber = peaknumber - 1;
            }
            if (peaknumber >= 30) break;
        }
    }
    return peaks[highestpeaknumber];
}

void orc_bass_beat(orc_beat_state* s, int32_t bass, int n, float time_delta, int strict)
{
    s->tempo_ring[s->tempo_pos] = bass; /* beatDetector.c:53 */
    s->tempo_pos++;
    if (s->tempo_pos >= FFTL_TEMPO_RING) s->tempo_pos = 0; /* :56-59 */
    orc_tempo_hb(s->tempo_ring, strict, s->hb);
    s->peak_index = orc_peak_pick(s->hb);
    /* beatDetector.c:105: int * float * int, evaluated left to right in float */
    s->movement_per_sec = ((float)n * time_delta) * (float)s->peak_index;
}

static int32_t band_energy(const fftl_config* c, const int32_t* h)
{
    /* fft_lines.c:281 -- int sum, int division */
    int32_t s = wrap_add(wrap_add(h[c->beat_bars[0]], h[c->beat_bars[1]]),
                         wrap_add(h[c->beat_bars[2]], h[c->beat_bars[3]]));
    return s / 4;
}

void orc_beat_replay(const fftl_config* c, const int32_t* w, const int32_t* bass, int64_t count,
                     int64_t first_frame, fftl_beat* out, int32_t* hb_out)
{
    orc_beat_state s;
    orc_beat_reset(&s);
    s.tempo_pos = (int32_t)(first_frame % FFTL_TEMPO_RING); /* ring slot = absolute frame mod 256 */
    float td = (float)c->hop / c->sample_rate;
    for (int64_t f = 0; f < count; f++) {
        orc_add_beat_sample(&s, w[f]);
        orc_bass_beat(&s, bass[f], c->n, td, c->strict_int);
        if (out) {
            fftl_beat* o = &out[f];
            memset(o, 0, sizeof(*o));
            o->w = w[f];
            o->thr = s.thr;
            o->beat_value = s.beat_value;
            o->flag = (s.thr > 0 && w[f] > s.thr) ? 1 : 0;
            o->bass = bass[f];
            o->peak_index = s.peak_index;
            o->movement_per_sec = s.movement_per_sec;
        }
        if (hb_out) memcpy(hb_out + f * (FFTL_BEAT_SAMPLES + 1), s.hb, sizeof(int32_t) * (FFTL_BEAT_SAMPLES + 1));
    }
}

/* ------------------------------------------------------------------- stft */
int orc_stft(const fftl_config* c, const int16_t* pcm, int32_t channels, int64_t S, int64_t cs,
             int64_t ss, int64_t f0, int64_t f1, double* bars_f64, float* bars_f32,
             int32_t* bars_i32, fftl_beat* beat)
{
    int rc = orc_config_validate(c);
    if (rc) return rc;
    int64_t F = orc_num_frames(c, S);
    if (f0 < 0 || f1 < f0 || f1 > F || channels < 1) return FFTL_ERR_ARG;
    orc_tables t;
    rc = tables_make(c, &t);
    if (rc) { tables_free(&t); return rc; }
    int B = c->bars;
    int64_t nf = f1 - f0;
    int64_t warm0 = f0 - (FFTL_TEMPO_RING - 1);
    if (warm0 < 0) warm0 = 0;
    if (!beat) warm0 = f0;
    int64_t nw = f1 - warm0;
    double* h = (double*)malloc(sizeof(double) * B);
    int32_t* hi32 = (int32_t*)malloc(sizeof(int32_t) * B);
    int32_t* wseq = (int32_t*)malloc(sizeof(int32_t) * (nw > 0 ? nw : 1));
    int32_t* bseq = (int32_t*)malloc(sizeof(int32_t) * (nw > 0 ? nw : 1));
    fftl_beat* brec = beat ? (fftl_beat*)malloc(sizeof(fftl_beat) * (nw > 0 ? nw : 1)) : NULL;
    for (int ch = 0; ch < channels; ch++) {
        for (int64_t f = warm0; f < f1; f++) {
            double bass_mag;
            frame_core(c, pcm + ch * cs + f * c->hop * ss, ss, t.win, t.lo, t.hi, t.sg, t.scratch,
                       NULL, h, hi32, &bass_mag);
            wseq[f - warm0] = band_energy(c, hi32);
            bseq[f - warm0] = orc_trunc_i32(bass_mag, c->strict_int); /* beatDetector.c:53 */
            if (f >= f0) {
                int64_t o = ((int64_t)ch * nf + (f - f0)) * B;
                for (int b = 0; b < B; b++) {
                    if (bars_f64) bars_f64[o + b] = h[b];
                    if (bars_f32) bars_f32[o + b] = (float)h[b];
                    if (bars_i32) bars_i32[o + b] = hi32[b];
                }
            }
        }
        if (beat) {
            orc_beat_replay(c, wseq, bseq, nw, warm0, brec, NULL);
            memcpy(beat + (int64_t)ch * nf, brec + (f0 - warm0), sizeof(fftl_beat) * nf);
        }
    }
    free(h); free(hi32); free(wseq); free(bseq); free(brec);
    tables_free(&t);
    return FFTL_OK;
}

/* --------------------------------------------------------------- synthetic
 * Bit-reproducible generator shared (as a specification) with the device
 * generator: only integer arithmetic and explicitly-rounded double ops, so a
 * CPU and a GPU produce identical int16 streams.
 *   x[c][n] = clip( rint( 8000*sin(sweep) + tone_c0 + noise ) )
 *   sweep: linear 20 Hz -> 20 kHz over 10 s, repeating (odd channels sweep
 *          downwards); phase kept as a 64-bit binary fraction of a cycle.
 *   tone (channel 0 only): 6000 * (0.5+0.5*sin(2 Hz)) * sin(94 Hz)
 *   noise: Philox4x32-10(counter = n, key = seed + c) -> sum of 8 uniform
 *          16-bit halves, centred and scaled to sigma ~ 2000.
 */
static double sin_turns_u32(uint32_t p)
{
    /* sin(2*pi*p/2^32): quarter-wave reduction on integer bits, then an odd
     * polynomial on [0, pi/2] in Horner form with fma() only. */
    uint32_t q = p >> 30;
    uint32_t r = p & 0x3FFFFFFFu;
    if (q & 1u) r = 0x40000000u - r;
    double x = (double)r * (1.5707963267948966 / 1073741824.0); /* exact int * constant: one rounding */
    double x2 = x * x;
    double s = -7.6471637318198164759e-13;                       /* -1/15! */
    s = fma(s, x2, 1.6059043836821614599e-10);                   /*  1/13! */
    s = fma(s, x2, -2.5052108385441718775e-08);                  /* -1/11! */
    s = fma(s, x2, 2.7557319223985890653e-06);                   /*  1/9!  */
    s = fma(s, x2, -1.9841269841269841270e-04);                  /* -1/7!  */
    s = fma(s, x2, 8.3333333333333333333e-03);                   /*  1/5!  */
    s = fma(s, x2, -1.6666666666666666667e-01);                  /* -1/3!  */
    s = fma(s * x2, x, x);
    return (q & 2u) ? -s : s;
}

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                          uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define ORC_SWEEP_SECONDS 10.0

static uint64_t frac64(double x, double two64)
{
    /* x cycles -> 2^-64 fixed point, modulo one cycle */
    if (x < 0) return (uint64_t)0 - frac64(-x, two64);
    x -= floor(x);
    return (uint64_t)(x * two64);
}

void orc_synth_pcm(int16_t* pcm, int32_t channels, int64_t samples, int64_t cs, int64_t first,
                   float sample_rate, uint32_t seed)
{
    double fs = (double)sample_rate;
    uint64_t period = (uint64_t)(ORC_SWEEP_SECONDS * fs);
    if (period == 0) period = 1;
    /* phase(m) = A*m + B*m^2 cycles (mod 1), as 2^-64 fixed point */
    double T = (double)period / fs;
    double two64 = 18446744073709551616.0;
    for (int c = 0; c < channels; c++) {
        double f0 = (c & 1) ? 20000.0 : 20.0, f1 = (c & 1) ? 20.0 : 20000.0;
        /* B is negative for down-sweeps: two's-complement negate the magnitude */
        uint64_t A = frac64(f0 / fs, two64);
        uint64_t Bq = frac64((f1 - f0) / (2.0 * T * fs * fs), two64);
        uint64_t tone_step = frac64(94.0 / fs, two64);
        uint64_t am_step = frac64(2.0 / fs, two64);
        for (int64_t i = 0; i < samples; i++) {
            uint64_t n = (uint64_t)(first + i);
            uint64_t m = n % period;
            uint64_t ph = A * m + Bq * m * m + ((uint64_t)c << 62);
            double v = 8000.0 * sin_turns_u32((uint32_t)(ph >> 32));
            if (c == 0) {
                double tone = sin_turns_u32((uint32_t)((tone_step * n) >> 32));
                double am = 0.5 + 0.5 * sin_turns_u32((uint32_t)((am_step * n) >> 32));
                double t = 6000.0 * am;
                v = v + t * tone;
            }
            uint32_t r[4];
            philox4x32_10((uint32_t)n, (uint32_t)(n >> 32), 0u, 0u, seed + (uint32_t)c, 0x5EED0000u, r);
            int32_t acc = 0;
            for (int k = 0; k < 4; k++) acc += (int32_t)(r[k] & 0xFFFFu) + (int32_t)(r[k] >> 16);
            acc -= 262140; /* 8 * 32767.5 */
            /* sigma(sum of 8 u16) = 65536*sqrt(8/12) = 53509.9; 2000/53509.9 ~ 2449/65536 */
            int32_t noise = (acc * 2449) >> 16; /* arithmetic shift; |acc*2449| < 2^31 */
            v = v + (double)noise;
            double r_ = rint(v);
            if (r_ > 32767.0) r_ = 32767.0;
            if (r_ < -32768.0) r_ = -32768.0;
            pcm[(int64_t)c * cs + i] = (int16_t)r_;
        }
    }
}

/* ------------------------------------------------- callers / sinks
 * Restated from the reference; test infrastructure like the rest of this file. */

/* readSettings (settingsFile.c:172-247): 11 numbers, each clamped to its default when out of range.
 * v[0]=barCount v[1]=zoom (as float bits via *zoom) v[2]=colorSel v[3..10] = the eight bools. */
void orc_settings_parse(const char* text, long len, int32_t* v, float* zoom)
{
    v[0] = 500; *zoom = 0.0001f; v[2] = 0;           /* settingsFile.h:10,13,14 */
    v[3] = 1; v[4] = 1; v[5] = 1; v[6] = 1; v[7] = 0; v[8] = 0; v[9] = 0; v[10] = 1; /* settingsFile.h:15-22 */
    if (len > 1000 || len == 0) return;              /* :178-181 */
    char* buffer = (char*)malloc((size_t)len + 1);
    memcpy(buffer, text, (size_t)len);
    buffer[len] = 0;
    char* next;
    long barCount = strtol(buffer, &next, 10);       /* :200 */
    if (barCount < 100 || barCount > 1000) barCount = 500;
    v[0] = (int32_t)barCount;
    float z = strtof(next, &next);                   /* :204 */
    if (z < 0.00001f || z > 0.001f) z = 0.0001f;
    *zoom = z;
    long long colorSel = strtoll(next, &next, 10);   /* :208 */
    if (colorSel < 0 || colorSel >= 6) colorSel = 0;
    v[2] = (int32_t)colorSel;
    for (int i = 3; i <= 10; i++) {                  /* :213-243: `bool x = strtol(...)` */
        _Bool b = strtol(next, &next, 10);
        v[i] = b;
    }
    free(buffer);
}

/* RANGE255 (drawBar2D.cpp:33) */
uint32_t orc_range255(int number, int maxValueOfNumber)
{
    return (unsigned int)((float)(((float)number / ((float)maxValueOfNumber - 1.0)) * 254.0));
}

/* fft_lines.c:250-255: line[0] = (char)((float)barLeft[led_bar].height * 0.1f), sent when height >= 0 */
void orc_led_byte(int32_t height, uint8_t* byte, uint8_t* valid)
{
    float v = (float)height * 0.1f;
    *byte = (uint8_t)(orc_trunc_i32((double)v, 1) & 0xff);
    *valid = height >= 0;
}

/* fft_lines.c:266-274, verbatim loop, over a series of beatMovementPerSec values */
void orc_marker_sweep(const float* movement, int64_t frames, float timeDelta, int32_t beatposition, int32_t direction,
                      int32_t* pos_out, int32_t* dir_out)
{
    for (int64_t f = 0; f < frames; f++) {
        float beatMovementPerSec = movement[f];
        for (int i = 0; i < (float)beatMovementPerSec * timeDelta; i++) {
            if (beatposition >= 2048) direction = -1;
            if (beatposition <= 0) direction = 1;
            beatposition += direction;
        }
        pos_out[f] = beatposition;
        dir_out[f] = direction;
    }
}

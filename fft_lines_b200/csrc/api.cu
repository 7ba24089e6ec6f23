// api.cu -- host side of libfftlines_cuda: the C-ABI declared in
// include/fftlines.h, include/fft_calculate.h and include/beatDetector.h.
//
// No CPU fallback anywhere: if CUDA is unusable every compute entry point
// fails and latches the error.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>

#include "kernels.cuh"
#include "rdif_launch.h"
#include "stream.cuh"
#include "stream_tick.cuh"
#include "../../include/fft_calculate.h"
#include "../../include/beatDetector.h"

using namespace fftl;

// ------------------------------------------------------------------ errors
static int g_err = FFTL_OK;
static char g_errmsg[512] = "";
static std::mutex g_err_mu;

static int set_error(int code, const char* fmt, ...)
{
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_errmsg, sizeof(g_errmsg), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return set_error(FFTL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                                   \
    } while (0)

extern "C" int fftl_last_error(void) { return g_err; }
extern "C" const char* fftl_last_error_string(void) { return g_errmsg; }
extern "C" void fftl_clear_error(void)
{
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = FFTL_OK;
    g_errmsg[0] = 0;
}
extern "C" int fftl_version(void) { return FFTL_VERSION; }
extern "C" int fftl_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------ config
extern "C" void fftl_config_reference(fftl_config* c, int32_t n, int32_t hop, int32_t bars)
{
    memset(c, 0, sizeof(*c));
    c->n = n;
    c->hop = hop;
    c->bars = bars;
    c->window = FFTL_WINDOW_RECT;     // fft_lines.c:190-194: raw samples
    c->binning = FFTL_BINNING_LINEAR; // fft_lines.c:199-203: bar i = bin i
    c->zoom = 1e-4f;                  // DEFAULT_ZOOM, settingsFile.h:13
    c->sample_rate = 48000.0f;
    c->fmin = 20.0f;
    c->fmax = 20000.0f;
    c->beat_bars[0] = 3;              // fft_lines.c:281
    c->beat_bars[1] = 4;
    c->beat_bars[2] = 5;
    c->beat_bars[3] = 6;
    c->bass_bin = 4;                  // beatDetector.c:53
    c->strict_int = 1;
}

extern "C" void fftl_config_extended(fftl_config* c, int32_t n, int32_t hop, int32_t bars, float sample_rate, float fmin,
                                     float fmax)
{
    fftl_config_reference(c, n, hop, bars);
    c->window = FFTL_WINDOW_HANN;
    c->binning = FFTL_BINNING_LOG;
    c->sg_half = 3;
    c->sg_degree = 2;
    c->sample_rate = sample_rate;
    c->fmin = fmin;
    c->fmax = fmax;
}

static bool pow2(int n) { return n > 0 && !(n & (n - 1)); }

extern "C" int fftl_config_validate(const fftl_config* c)
{
    if (!c) return set_error(FFTL_ERR_ARG, "config is NULL");
    if (!pow2(c->n) || c->n < 256 || c->n > 16384) return set_error(FFTL_ERR_ARG, "n=%d: need a power of two in 256..16384", c->n);
    if (c->hop < 1) return set_error(FFTL_ERR_ARG, "hop=%d must be >= 1", c->hop);
    if (c->bars < 1 || c->bars > FFTL_MAX_BARS) return set_error(FFTL_ERR_ARG, "bars=%d out of 1..%d", c->bars, FFTL_MAX_BARS);
    if (c->window != FFTL_WINDOW_RECT && c->window != FFTL_WINDOW_HANN) return set_error(FFTL_ERR_ARG, "unknown window %d", c->window);
    if (!(c->sample_rate > 0)) return set_error(FFTL_ERR_ARG, "sample_rate must be positive");
    if (c->binning == FFTL_BINNING_LINEAR) {
        if (c->bars > c->n / 2 + 1) return set_error(FFTL_ERR_ARG, "LINEAR binning: bars=%d exceeds n/2+1=%d", c->bars, c->n / 2 + 1);
    } else if (c->binning == FFTL_BINNING_LOG) {
        if (!(c->fmin > 0) || !(c->fmax > c->fmin) || c->fmax > 0.5f * c->sample_rate)
            return set_error(FFTL_ERR_ARG, "LOG binning: need 0 < fmin < fmax <= sample_rate/2");
    } else
        return set_error(FFTL_ERR_ARG, "unknown binning %d", c->binning);
    if (c->sg_half < 0 || c->sg_half > FFTL_MAX_SG_HALF) return set_error(FFTL_ERR_ARG, "sg_half=%d out of 0..%d", c->sg_half, FFTL_MAX_SG_HALF);
    if (c->sg_half > 0) {
        if (c->sg_degree < 0 || c->sg_degree > 2 * c->sg_half) return set_error(FFTL_ERR_ARG, "sg_degree=%d out of 0..%d", c->sg_degree, 2 * c->sg_half);
        if (c->bars < 2 * c->sg_half + 1) return set_error(FFTL_ERR_ARG, "SG window wider than the bar vector");
    }
    for (int i = 0; i < 4; i++)
        if (c->beat_bars[i] < 0 || c->beat_bars[i] >= c->bars) return set_error(FFTL_ERR_ARG, "beat_bars[%d]=%d outside 0..bars-1", i, c->beat_bars[i]);
    if (c->bass_bin < 0 || c->bass_bin > c->n / 2) return set_error(FFTL_ERR_ARG, "bass_bin=%d outside 0..n/2", c->bass_bin);
    return FFTL_OK;
}

extern "C" int fftl_build_bins(const fftl_config* c, int32_t* lo, int32_t* hi)
{
    int rc = fftl_config_validate(c);
    if (rc) return rc;
    const int top = c->n / 2 + 1;
    if (c->binning == FFTL_BINNING_LINEAR) {
        for (int b = 0; b < c->bars; b++) {
            lo[b] = b;
            hi[b] = b + 1;
        }
        return FFTL_OK;
    }
    // geometric band edges fmin..fmax; a bar covers the bins whose index lies
    // between its two edges (at least one bin, low bars may repeat a bin)
    const double ratio = (double)c->fmax / (double)c->fmin;
    const double bins_per_hz = (double)c->n / (double)c->sample_rate;
    for (int b = 0; b < c->bars; b++) {
        double edge_lo = (double)c->fmin * pow(ratio, (double)b / (double)c->bars);
        double edge_hi = (double)c->fmin * pow(ratio, (double)(b + 1) / (double)c->bars);
        int first = (int)floor(edge_lo * bins_per_hz);
        int last = (int)floor(edge_hi * bins_per_hz);
        first = first < 0 ? 0 : (first > top - 1 ? top - 1 : first);
        if (last <= first) last = first + 1;
        if (last > top) last = top;
        lo[b] = first;
        hi[b] = last;
    }
    return FFTL_OK;
}

// Savitzky-Golay via orthonormal (Gram-Schmidt) polynomials on x=-w..w:
// P = Q Q^T with Q the orthonormal basis of degree <= d.
extern "C" int fftl_build_sg(int32_t w, int32_t d, float* coef)
{
    if (w < 1 || w > FFTL_MAX_SG_HALF || d < 0 || d > 2 * w) return set_error(FFTL_ERR_ARG, "bad SG parameters w=%d d=%d", w, d);
    const int m = 2 * w + 1;
    std::vector<std::vector<long double>> q;
    for (int j = 0; j <= d; j++) {
        std::vector<long double> col(m);
        for (int i = 0; i < m; i++) col[i] = powl((long double)(i - w) / (long double)w, j);
        for (int pass = 0; pass < 2; pass++) // re-orthogonalise once
            for (auto& prev : q) {
                long double dot = 0;
                for (int i = 0; i < m; i++) dot += prev[i] * col[i];
                for (int i = 0; i < m; i++) col[i] -= dot * prev[i];
            }
        long double nrm = 0;
        for (int i = 0; i < m; i++) nrm += col[i] * col[i];
        nrm = sqrtl(nrm);
        if (nrm < 1e-18L) return set_error(FFTL_ERR_ARG, "SG basis is degenerate");
        for (int i = 0; i < m; i++) col[i] /= nrm;
        q.push_back(col);
    }
    for (int i = 0; i < m; i++)
        for (int k = 0; k < m; k++) {
            long double s = 0;
            for (auto& col : q) s += col[i] * col[k];
            coef[i * m + k] = (float)s;
        }
    return FFTL_OK;
}

// ------------------------------------------------------------------ tables
static void fill_twiddles(std::vector<float2>& tw, int n)
{
    tw.resize(n);
    for (int m = 0; m < n; m++) {
        // exact octant symmetry is not needed; double sincos of the exact angle
        double a = -2.0 * M_PI * (double)m / (double)n;
        tw[m] = make_float2((float)cos(a), (float)sin(a));
    }
}

template <int N> static cudaError_t launch_stft(const StftParams& p, int channels, long long pairs, cudaStream_t st)
{
    using Sh = StftShape<N>;
    size_t smem = Sh::smem_bytes(p.bars_n);
    if (smem > 48 * 1024) { // opt-in is per device, so no process-wide cache
        cudaError_t e = cudaFuncSetAttribute(stft_bars_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // pairs of all channels are numbered consecutively, so many-channel / few-frame jobs still fill CTAs
    long long blocks = (pairs * channels + Sh::PAIRS - 1) / Sh::PAIRS;
    if (blocks <= 0) return cudaSuccess;
    if (blocks > 0x7fffffffll) return cudaErrorInvalidValue;
    StftParams q = p;
    q.pairs_per_channel = pairs;
    q.channels = channels;
    stft_bars_kernel<N><<<(unsigned)blocks, Sh::THREADS, smem, st>>>(q);
    return cudaGetLastError();
}

static cudaError_t dispatch_stft(int n, const StftParams& p, int channels, long long pairs, cudaStream_t st)
{
    switch (n) {
    case 256: return launch_stft<256>(p, channels, pairs, st);
    case 512: return launch_stft<512>(p, channels, pairs, st);
    case 1024: return launch_stft<1024>(p, channels, pairs, st);
    case 2048: return launch_stft<2048>(p, channels, pairs, st);
    case 4096: return launch_stft<4096>(p, channels, pairs, st);
    case 8192: return launch_stft<8192>(p, channels, pairs, st);
    case 16384: return launch_stft<16384>(p, channels, pairs, st);
    }
    return cudaErrorInvalidValue;
}

template <int N> static cudaError_t launch_c2c(const float2* in, float2* out, const float2* tw, cudaStream_t st)
{
    constexpr int VPT = (N >= 16384) ? 32 : 16;
    size_t smem = padded_size(N) * sizeof(float2);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(c2c_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    c2c_kernel<N><<<1, N / VPT, smem, st>>>(in, out, tw);
    return cudaGetLastError();
}

static cudaError_t dispatch_c2c(int n, const float2* in, float2* out, const float2* tw, cudaStream_t st)
{
    switch (n) {
    case 256: return launch_c2c<256>(in, out, tw, st);
    case 512: return launch_c2c<512>(in, out, tw, st);
    case 1024: return launch_c2c<1024>(in, out, tw, st);
    case 2048: return launch_c2c<2048>(in, out, tw, st);
    case 4096: return launch_c2c<4096>(in, out, tw, st);
    case 8192: return launch_c2c<8192>(in, out, tw, st);
    case 16384: return launch_c2c<16384>(in, out, tw, st);
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ handle
static const int TIMING_SLOTS = 64;

struct fftl_stft {
    fftl_config cfg;
    int device;
    cudaStream_t stream;            // compute
    cudaStream_t s_h2d, s_d2h;      // copy streams of the host-buffer path
    float* d_window;
    float2* d_tw;
    float2* d_tw256;
    double2* d_tw256d;
    int* d_lo;
    int* d_hi;
    float* d_sg;
    // fast path (stft_rdif.cuh): sub-transform twiddles and the balanced binning segments
    float2* d_twm;
    int* d_seg_packed;
    bool seg_packable;
    int* d_bar_seg0;
    float* d_bar_inv;
    int nseg;
    int mag_bins;                   // highest bin any bar (or the bass bin) reads, + 1
    int num_sms;
    int64_t fast_launches;
    int64_t tmem_launches;          // fast-path launches that took the tensor-memory variant (stft_rdif_tm.cuh)
    int kernel_select;              // FFTL_KERNEL_AUTO / FFTL_KERNEL_GENERIC (fftl_stft_set_kernel)
    int* d_aux;                     // [2][channels][nf_aux] (w then bass)
    size_t aux_capacity;            // ints
    int64_t launches;
    // device timing of run_device calls (ring of event triples)
    cudaEvent_t ev[TIMING_SLOTS][3];
    int timing_count;               // calls recorded since last reset (saturates at TIMING_SLOTS)
    int timing_next;
    // staging of the host-buffer path
    void* d_stage[2][4];            // per slot: pcm, bars, bars_i32, beat
    size_t stage_bytes[2][4];
    cudaEvent_t ev_h2d[2], ev_comp[2], ev_d2h[2];
};

static int ensure_bytes(void** p, size_t* cap, size_t need)
{
    if (*cap >= need) return FFTL_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) return set_error(FFTL_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
    *cap = need;
    return FFTL_OK;
}

extern "C" int fftl_stft_create(const fftl_config* cfg, int device, fftl_stft** out)
{
    if (!out) return set_error(FFTL_ERR_ARG, "out is NULL");
    *out = nullptr;
    int rc = fftl_config_validate(cfg);
    if (rc) return rc;
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return set_error(FFTL_ERR_CUDA, "no CUDA device (libfftlines_cuda has no CPU fallback)");
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= ndev) return set_error(FFTL_ERR_ARG, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    fftl_stft* h = (fftl_stft*)calloc(1, sizeof(fftl_stft));
    if (!h) return set_error(FFTL_ERR_NOMEM, "out of host memory");
    h->cfg = *cfg;
    h->device = device;
    const int n = cfg->n, B = cfg->bars;
    std::vector<float2> tw, tw256;
    fill_twiddles(tw, n);
    fill_twiddles(tw256, FFTL_TEMPO_RING);
    std::vector<int32_t> lo(B), hi(B);
    fftl_build_bins(cfg, lo.data(), hi.data());
#define H_TRY(expr)                                                                                  \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            set_error(FFTL_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));               \
            fftl_stft_destroy(h);                                                                    \
            return FFTL_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)
    H_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    H_TRY(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    H_TRY(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    H_TRY(cudaMalloc(&h->d_tw, sizeof(float2) * n));
    H_TRY(cudaMemcpy(h->d_tw, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    H_TRY(cudaMalloc(&h->d_tw256, sizeof(float2) * FFTL_TEMPO_RING));
    H_TRY(cudaMemcpy(h->d_tw256, tw256.data(), sizeof(float2) * FFTL_TEMPO_RING, cudaMemcpyHostToDevice));
    {
        std::vector<double2> twd(FFTL_TEMPO_RING);
        for (int m = 0; m < FFTL_TEMPO_RING; m++) {
            double a = -2.0 * M_PI * (double)m / (double)FFTL_TEMPO_RING;
            twd[m] = make_double2(cos(a), sin(a));
        }
        H_TRY(cudaMalloc(&h->d_tw256d, sizeof(double2) * FFTL_TEMPO_RING));
        H_TRY(cudaMemcpy(h->d_tw256d, twd.data(), sizeof(double2) * FFTL_TEMPO_RING, cudaMemcpyHostToDevice));
    }
    H_TRY(cudaMalloc(&h->d_lo, sizeof(int) * B));
    H_TRY(cudaMalloc(&h->d_hi, sizeof(int) * B));
    H_TRY(cudaMemcpy(h->d_lo, lo.data(), sizeof(int) * B, cudaMemcpyHostToDevice));
    H_TRY(cudaMemcpy(h->d_hi, hi.data(), sizeof(int) * B, cudaMemcpyHostToDevice));
    if (cfg->window == FFTL_WINDOW_HANN) {
        std::vector<float> win(n);
        for (int i = 0; i < n; i++) win[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)i / (double)n)); // periodic Hann
        H_TRY(cudaMalloc(&h->d_window, sizeof(float) * n));
        H_TRY(cudaMemcpy(h->d_window, win.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
    }
    if (cfg->sg_half > 0) {
        int m = 2 * cfg->sg_half + 1;
        std::vector<float> sg(m * m);
        if (fftl_build_sg(cfg->sg_half, cfg->sg_degree, sg.data())) {
            fftl_stft_destroy(h);
            return FFTL_ERR_ARG;
        }
        H_TRY(cudaMalloc(&h->d_sg, sizeof(float) * m * m));
        H_TRY(cudaMemcpy(h->d_sg, sg.data(), sizeof(float) * m * m, cudaMemcpyHostToDevice));
    }
    {
        // balanced binning: every DISTINCT bin range (the low log-frequency bars repeat ranges) is cut into segments of
        // at most L bins that never cross a multiple of 16 (so a segment is contiguous in the padded magnitude
        // array); a bar adds up its range's segment sums in a fixed order, which keeps the result deterministic.
        // L is the smallest of 8, 12, 16 for which the fast kernel's segment phase is ONE pass over its threads
        // (a second pass costs a full shared-memory round trip per tile: 3.6 % of the c2 kernel).
        const int fast_threads = n / 16 > 512 ? 512 : n / 16; // RdifShape::THREADS
        std::vector<int> seg_lo, seg_hi, bar_seg0(B + 1);     // bar_seg0[b] = first segment | segments << 16
        std::vector<float> bar_inv(B);
        int max_per_bar = 0;
        for (int L = 8; L <= 16; L += 4) {
            seg_lo.clear();
            seg_hi.clear();
            max_per_bar = 0;
            for (int b = 0; b < B; b++) {
                if (b > 0 && lo[b] == lo[b - 1] && hi[b] == hi[b - 1]) { // same range as the bar before: share its segments
                    bar_seg0[b] = bar_seg0[b - 1];
                    continue;
                }
                const int first = (int)seg_lo.size();
                for (int k = lo[b]; k < hi[b];) {
                    int e = k + L < hi[b] ? k + L : hi[b];
                    int block_end = (k / 16 + 1) * 16;
                    if (e > block_end) e = block_end;
                    seg_lo.push_back(k);
                    seg_hi.push_back(e);
                    k = e;
                }
                const int cnt = (int)seg_lo.size() - first;
                bar_seg0[b] = first | (cnt << 16);
                max_per_bar = cnt > max_per_bar ? cnt : max_per_bar;
            }
            if ((int)seg_lo.size() <= fast_threads) break;
        }
        for (int b = 0; b < B; b++) bar_inv[b] = 1.0f / (float)(hi[b] - lo[b]);
        bar_seg0[B] = 0;
        h->nseg = (int)seg_lo.size();
        h->mag_bins = cfg->bass_bin + 1;
        for (int b = 0; b < B; b++) h->mag_bins = hi[b] > h->mag_bins ? hi[b] : h->mag_bins;
        // work list of the segment sums in bin order, packed for the fast kernel
        std::vector<int> packed(h->nseg);
        h->seg_packable = h->nseg < (1 << 13) && max_per_bar < (1 << 15);
        for (int i = 0; i < h->nseg; i++) {
            const int k = i, pl = pad_index(seg_lo[k]);
            if (pl >= (1 << 14)) h->seg_packable = false;
            packed[i] = FFTL_SEG_PACK(pl, seg_hi[k] - seg_lo[k], k);
        }
        H_TRY(cudaMalloc(&h->d_seg_packed, sizeof(int) * h->nseg));
        H_TRY(cudaMalloc(&h->d_bar_seg0, sizeof(int) * (B + 1)));
        H_TRY(cudaMalloc(&h->d_bar_inv, sizeof(float) * B));
        H_TRY(cudaMemcpy(h->d_seg_packed, packed.data(), sizeof(int) * h->nseg, cudaMemcpyHostToDevice));
        H_TRY(cudaMemcpy(h->d_bar_seg0, bar_seg0.data(), sizeof(int) * (B + 1), cudaMemcpyHostToDevice));
        H_TRY(cudaMemcpy(h->d_bar_inv, bar_inv.data(), sizeof(float) * B, cudaMemcpyHostToDevice));
        std::vector<float2> twm;
        fill_twiddles(twm, n / 16);
        H_TRY(cudaMalloc(&h->d_twm, sizeof(float2) * (n / 16)));
        H_TRY(cudaMemcpy(h->d_twm, twm.data(), sizeof(float2) * (n / 16), cudaMemcpyHostToDevice));
        H_TRY(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    }
    for (int i = 0; i < TIMING_SLOTS; i++)
        for (int k = 0; k < 3; k++) H_TRY(cudaEventCreate(&h->ev[i][k]));
    for (int i = 0; i < 2; i++) {
        H_TRY(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
        H_TRY(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
        H_TRY(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
    }
    // The uploads above are pageable-memory copies and memsets on the legacy stream; the handle's own streams are
    // non-blocking and do not order against it, so the tables must have landed before the handle is handed out.
    H_TRY(cudaDeviceSynchronize());
#undef H_TRY
    *out = h;
    return FFTL_OK;
}

extern "C" void fftl_stft_destroy(fftl_stft* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->s_h2d) cudaStreamSynchronize(h->s_h2d);
    if (h->s_d2h) cudaStreamSynchronize(h->s_d2h);
    cudaFree(h->d_window);
    cudaFree(h->d_tw);
    cudaFree(h->d_tw256);
    cudaFree(h->d_tw256d);
    cudaFree(h->d_lo);
    cudaFree(h->d_hi);
    cudaFree(h->d_sg);
    cudaFree(h->d_aux);
    cudaFree(h->d_twm);
    cudaFree(h->d_seg_packed);
    cudaFree(h->d_bar_seg0);
    cudaFree(h->d_bar_inv);
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 4; k++) cudaFree(h->d_stage[i][k]);
    for (int i = 0; i < TIMING_SLOTS; i++)
        for (int k = 0; k < 3; k++)
            if (h->ev[i][k]) cudaEventDestroy(h->ev[i][k]);
    for (int i = 0; i < 2; i++) {
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
        if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    cudaGetLastError();
    free(h);
}

extern "C" int64_t fftl_stft_num_frames(const fftl_stft* h, int64_t samples)
{
    if (!h || samples < h->cfg.n) return 0;
    return (samples - h->cfg.n) / h->cfg.hop + 1;
}

static int64_t beat_warm_start(int64_t fb);
extern "C" int64_t fftl_stft_first_frame_needed(int64_t frame_begin, int32_t with_beat)
{
    if (frame_begin < 0) frame_begin = 0;
    return (with_beat ? beat_warm_start(frame_begin) : frame_begin) & ~(int64_t)1;
}

extern "C" int64_t fftl_stft_launch_count(const fftl_stft* h) { return h ? h->launches : 0; }
extern "C" int64_t fftl_stft_fast_launch_count(const fftl_stft* h) { return h ? h->fast_launches : 0; }
extern "C" int64_t fftl_stft_tmem_launch_count(const fftl_stft* h) { return h ? h->tmem_launches : 0; }

extern "C" int fftl_stft_set_kernel(fftl_stft* h, int32_t which)
{
    if (!h) return set_error(FFTL_ERR_ARG, "handle is NULL");
    if (which != FFTL_KERNEL_AUTO && which != FFTL_KERNEL_GENERIC && which != FFTL_KERNEL_FAST_REGISTERS) return set_error(FFTL_ERR_ARG, "unknown kernel selection %d", which);
    h->kernel_select = which;
    return FFTL_OK;
}

static cudaError_t try_rdif(const fftl_stft* h, const StftParams& q, int channels, long long samples, cudaStream_t st, int* used_tm)
{
    const fftl_config& c = h->cfg;
    if (c.n != 16384 && c.n != 8192 && c.n != 4096 && c.n != 2048 && c.n != 1024) return cudaErrorNotSupported;
    // hop = 1, 2 or 4 columns (n/16 samples) on planar PCM: the frames of a tile share their samples in registers;
    // any other hop, or interleaved PCM: the variant in which every frame loads its own (hopq = 0).
    // n = 2048 is built for hop 512 (4 columns) and the own-samples variant, n = 1024 for the own-samples variant.
    const int M = c.n / 16;
    int hopq = (q.hop % M) == 0 ? q.hop / M : 0; // q.hop: the caller's hop (the streaming front end passes n: one frame per window)
    if ((hopq != 1 && hopq != 2 && hopq != 4) || q.sample_stride != 1 || (c.n == 2048 && hopq != 4) || c.n == 1024 || c.n == 16384 || (c.n == 8192 && hopq != 1)) hopq = 0;
    if (h->kernel_select == FFTL_KERNEL_GENERIC || !h->seg_packable) return cudaErrorNotSupported;
    RdifParams p;
    memset(&p, 0, sizeof(p));
    p.pcm = q.pcm;
    p.channel_stride = q.channel_stride;
    p.sample_stride = q.sample_stride;
    p.samples = samples;
    p.frame0 = q.frame0;
    p.frame_emit = q.frame_emit;
    p.frame_end = q.frame_end;
    p.nf_emit = q.nf_emit;
    p.nf_aux = q.nf_aux;
    p.aux_base = q.aux_base;
    p.window = q.window;
    p.tw = q.tw;
    p.twm = h->d_twm;
    p.seg_packed = h->d_seg_packed;
    p.bar_seg0 = h->d_bar_seg0;
    p.bar_inv = h->d_bar_inv;
    p.sg = q.sg;
    p.bars = q.bars;
    p.bars_i32 = q.bars_i32;
    p.aux_w = q.aux_w;
    p.aux_bass = q.aux_bass;
    p.hop = q.hop;
    p.bars_n = q.bars_n;
    p.nseg = h->nseg;
    p.sg_half = q.sg_half;
    p.strict = q.strict;
    p.bass_bin = q.bass_bin;
    const int full_rows = c.n / 512;
    p.mag_bins = h->mag_bins;
    p.mag_rows = (h->mag_bins + 255) / 256;       // bins < 256*rows are kept; all rows also keep bin N/2
    if (p.mag_rows > full_rows || h->mag_bins > c.n / 2) p.mag_rows = full_rows;
    if (c.n >= 8192) p.mag_rows = full_rows; // built with all rows only
    p.mag_stride = c.n == 16384 ? RdifShape<16384, 2>::mag_stride(p.mag_rows) : c.n == 8192 ? RdifShape<8192, 4>::mag_stride(p.mag_rows) : c.n == 4096 ? RdifShape<4096, 4>::mag_stride(p.mag_rows) : c.n == 2048 ? RdifShape<2048, 4>::mag_stride(p.mag_rows) : RdifShape<1024, 4>::mag_stride(p.mag_rows);
    for (int i = 0; i < 4; i++) p.beat_bars[i] = q.beat_bars[i];
    p.zoom = q.zoom;
    p.circle_add = q.circle_add;
    p.no_tm = h->kernel_select == FFTL_KERNEL_FAST_REGISTERS;
    p.used_tm = used_tm;
    const bool win = q.window != nullptr;
    switch (hopq) {
    case 0: return launch_rdif_h0(c.n, win, h->num_sms, p, channels, st);
    case 1: return launch_rdif_h1(c.n, win, h->num_sms, p, channels, st);
    case 2: return launch_rdif_h2(c.n, win, h->num_sms, p, channels, st);
    default: return launch_rdif_h4(c.n, win, h->num_sms, p, channels, st);
    }
}

// Core: frames [fb, fe) of every channel; aux (w, bass) indexed from aux_base
// with per-channel stride nf_aux.  `compute_from` <= fb is where this call
// starts computing aux (frames before it must already be in the aux arrays).
static int run_range(fftl_stft* h, const int16_t* pcm, int channels, int64_t cstride, int64_t sstride, int64_t samples, int64_t compute_from,
                     int64_t fb, int64_t fe, int64_t out_base, int64_t nf_emit, int64_t aux_base, int64_t nf_aux, float* bars,
                     int32_t* bars_i32, bool aux, fftl_beat* beat, cudaStream_t st, cudaEvent_t* evs, bool force_generic = false, int hop_override = 0)
{
    const fftl_config& c = h->cfg;
    StftParams p;
    memset(&p, 0, sizeof(p));
    p.pcm = pcm;
    p.channel_stride = cstride;
    p.sample_stride = sstride;
    p.frame0 = compute_from;
    p.frame_emit = fb;
    p.frame_end = fe;
    p.samples = samples;
    p.nf_emit = nf_emit;
    p.nf_aux = nf_aux;
    p.aux_base = aux_base;
    p.window = h->d_window;
    p.tw = h->d_tw;
    p.lo = h->d_lo;
    p.hi = h->d_hi;
    p.bar_inv = h->d_bar_inv;
    p.sg = h->d_sg;
    // outputs are addressed relative to frame_emit inside the kernel
    p.bars = bars + (fb - out_base) * (int64_t)c.bars;
    p.bars_i32 = bars_i32 ? bars_i32 + (fb - out_base) * (int64_t)c.bars : nullptr;
    p.aux_w = aux ? h->d_aux : nullptr;
    p.aux_bass = aux ? h->d_aux + (size_t)channels * nf_aux : nullptr;
    p.hop = hop_override ? hop_override : c.hop;
    p.bars_n = c.bars;
    p.sg_half = c.sg_half;
    p.strict = c.strict_int;
    p.bass_bin = c.bass_bin;
    for (int i = 0; i < 4; i++) p.beat_bars[i] = c.beat_bars[i];
    p.zoom = c.zoom;
    p.circle_add = c.circle ? 10.0f : 0.0f; // fft_lines.c:204-207
    long long pairs = (fe - compute_from + 1) / 2;
    if (evs) CUDA_TRY(cudaEventRecord(evs[0], st));
    int used_tm = 0;
    cudaError_t fe_ = force_generic ? cudaErrorNotSupported : try_rdif(h, p, channels, samples, st, &used_tm);
    if (fe_ == cudaErrorNotSupported) {
        CUDA_TRY(dispatch_stft(c.n, p, channels, pairs, st)); // generic kernel: any n, strided PCM, any hop
    } else {
        CUDA_TRY(fe_);
        h->fast_launches++;
        h->tmem_launches += used_tm;
    }
    h->launches++;
    if (evs) CUDA_TRY(cudaEventRecord(evs[1], st));
    if (beat && fe > fb) {
        BeatParams bp;
        bp.aux_w = p.aux_w;
        bp.aux_bass = p.aux_bass;
        bp.out = beat + (fb - out_base);
        bp.tw256 = h->d_tw256;
        bp.tw256d = h->d_tw256d;
        bp.aux_base = aux_base;
        bp.nf_aux = nf_aux;
        bp.frame_emit = fb;
        bp.frame_end = fe;
        bp.nf_emit = nf_emit;
        float dt = (float)c.hop / c.sample_rate;  // timeDelta of a batch run (fft_lines.c:265 uses wall clock)
        bp.n_times_dt = (float)c.n * dt;          // beatDetector.c:105, left-to-right in float
        bp.strict = c.strict_int;
        // blocks of BEAT_FB frames aligned to absolute frame numbers
        dim3 grid((unsigned)((fe + BEAT_FB - 1) / BEAT_FB - fb / BEAT_FB), (unsigned)channels);
        beat_post_kernel<<<grid, BEAT_THREADS, 0, st>>>(bp);
        CUDA_TRY(cudaGetLastError());
        h->launches++;
    }
    if (evs) CUDA_TRY(cudaEventRecord(evs[2], st));
    return FFTL_OK;
}

// First frame whose (w, bass) the beat stage of a run emitting from frame fb needs: the tempo-ring transform
// is carried forward from the start of fb's block of BEAT_FB frames, whose own state needs 256 frames of history.
static int64_t beat_warm_start(int64_t fb)
{
    int64_t w = fb / BEAT_FB * BEAT_FB - FFTL_TEMPO_RING;
    return w < 0 ? 0 : w;
}

static int check_run_args(fftl_stft* h, const void* pcm, int channels, int64_t samples, int64_t fb, int64_t fe, const void* bars)
{
    if (!h) return set_error(FFTL_ERR_ARG, "handle is NULL");
    if (!pcm || !bars) return set_error(FFTL_ERR_ARG, "pcm and bars must not be NULL");
    if (channels < 1 || channels > 65535) return set_error(FFTL_ERR_ARG, "channels=%d out of 1..65535", channels);
    int64_t F = fftl_stft_num_frames(h, samples);
    if (fb < 0 || fe < fb || fe > F) return set_error(FFTL_ERR_ARG, "frame range [%lld,%lld) outside [0,%lld]", (long long)fb, (long long)fe, (long long)F);
    return FFTL_OK;
}

// Shared body of the device-buffer entry points.  `pcm` is addressed with ABSOLUTE sample indices of the job
// (a span-only shard passes a rebased pointer); samples at or beyond `samples_held` read as zero.
static int run_device_impl(fftl_stft* h, const int16_t* pcm, int32_t channels, int64_t samples_held, int64_t cstride, int64_t sstride,
                           int64_t fb, int64_t fe, float* bars, int32_t* bars_i32, fftl_beat* beat, cudaStream_t st)
{
    CUDA_TRY(cudaSetDevice(h->device));
    int64_t warm = fb;
    if (beat) {
        warm = beat_warm_start(fb);
    }
    warm &= ~(int64_t)1; // frame f always shares its transform with frame f^1: results do not depend on the shard start
    int64_t nf_aux = fe - warm;
    int rc;
    if (beat) {
        size_t need = sizeof(int) * 2 * (size_t)channels * (size_t)nf_aux;
        if (h->aux_capacity < need) {
            // one call per handle may be in flight (fftlines.h): nothing else can be using the old buffer once st has drained
            CUDA_TRY(cudaStreamSynchronize(st));
            void* pv = h->d_aux;
            rc = ensure_bytes(&pv, &h->aux_capacity, need);
            h->d_aux = (int*)pv;
            if (rc) return rc;
        }
    }
    // inside a CUDA-graph capture the timing events are skipped (they could not be queried afterwards)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) {
        cudaGetLastError();
        cap = cudaStreamCaptureStatusNone;
    }
    cudaEvent_t* evs = (cap == cudaStreamCaptureStatusNone) ? h->ev[h->timing_next] : nullptr;
    rc = run_range(h, pcm, channels, cstride, sstride, samples_held, warm, fb, fe, fb, fe - fb, warm, nf_aux, bars, bars_i32, beat != nullptr, beat, st, evs);
    if (rc) return rc;
    if (evs) {
        h->timing_next = (h->timing_next + 1) % TIMING_SLOTS;
        if (h->timing_count < TIMING_SLOTS) h->timing_count++;
    }
    return FFTL_OK;
}

extern "C" int fftl_stft_run_device(fftl_stft* h, const int16_t* pcm, int32_t channels, int64_t samples, int64_t cstride,
                                    int64_t sstride, int64_t fb, int64_t fe, float* bars, int32_t* bars_i32, fftl_beat* beat,
                                    void* cuda_stream)
{
    int rc = check_run_args(h, pcm, channels, samples, fb, fe, bars);
    if (rc) return rc;
    if (fe == fb) return FFTL_OK;
    // NULL = CUDA's default stream, as everywhere in CUDA
    return run_device_impl(h, pcm, channels, samples, cstride, sstride, fb, fe, bars, bars_i32, beat, (cudaStream_t)cuda_stream);
}

// Samples [*first, *end) of every channel that a run over frames [fb, fe) of a job of `total` samples reads:
// from the first frame of the beat warm-up (rounded down to even) to the end of frame fe-1 or, when that frame
// is even, of its pair partner fe (the two share packed sub-transforms), clipped to the job.
extern "C" int fftl_stft_sample_span(const fftl_stft* h, int64_t total, int64_t fb, int64_t fe, int32_t with_beat, int64_t* first,
                                     int64_t* end)
{
    if (!h || !first || !end) return set_error(FFTL_ERR_ARG, "NULL argument");
    const int64_t F = fftl_stft_num_frames(h, total);
    if (fb < 0 || fe < fb || fe > F) return set_error(FFTL_ERR_ARG, "frame range [%lld,%lld) outside [0,%lld]", (long long)fb, (long long)fe, (long long)F);
    if (fe == fb) {
        *first = *end = 0;
        return FFTL_OK;
    }
    *first = fftl_stft_first_frame_needed(fb, with_beat) * (int64_t)h->cfg.hop;
    const int64_t last = (fe - 1) | 1; // the odd frame of the last pair
    const int64_t e = last * (int64_t)h->cfg.hop + h->cfg.n;
    *end = e < total ? e : total;
    return FFTL_OK;
}

extern "C" int fftl_stft_run_device_span(fftl_stft* h, const int16_t* pcm, int32_t channels, int64_t span_first, int64_t span_samples,
                                         int64_t total, int64_t cstride, int64_t sstride, int64_t fb, int64_t fe, float* bars,
                                         int32_t* bars_i32, fftl_beat* beat, void* cuda_stream)
{
    int rc = check_run_args(h, pcm, channels, total, fb, fe, bars);
    if (rc) return rc;
    if (fe == fb) return FFTL_OK;
    int64_t need0 = 0, need1 = 0;
    if ((rc = fftl_stft_sample_span(h, total, fb, fe, beat != nullptr, &need0, &need1))) return rc;
    if (span_first < 0 || span_samples < 0 || span_first > need0 || span_first + span_samples < need1)
        return set_error(FFTL_ERR_ARG, "span [%lld,%lld) does not cover the samples frames [%lld,%lld) read: [%lld,%lld) (fftl_stft_sample_span)",
                         (long long)span_first, (long long)(span_first + span_samples), (long long)fb, (long long)fe, (long long)need0, (long long)need1);
    int64_t held = span_first + span_samples;
    if (held > total) held = total;
    // rebased pointer: element [c*cstride + s*sstride] is absolute sample s; nothing below span_first is ever read
    // (the run starts at need0 >= span_first) and everything at or beyond `held` reads as zero
    return run_device_impl(h, pcm - span_first * sstride, channels, held, cstride, sstride, fb, fe, bars, bars_i32, beat, (cudaStream_t)cuda_stream);
}

// Sum of device times over the run_device calls recorded since the last call
// of this function (at most TIMING_SLOTS).  Synchronises the events.
extern "C" int fftl_stft_timing(fftl_stft* h, int32_t* calls, float* total_ms, float* bars_kernel_ms)
{
    if (!h) return set_error(FFTL_ERR_ARG, "handle is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    float tot = 0, bars = 0;
    int n = h->timing_count;
    for (int i = 0; i < n; i++) {
        int slot = (h->timing_next - 1 - i + 2 * TIMING_SLOTS) % TIMING_SLOTS;
        CUDA_TRY(cudaEventSynchronize(h->ev[slot][2]));
        float a = 0, b = 0;
        CUDA_TRY(cudaEventElapsedTime(&a, h->ev[slot][0], h->ev[slot][2]));
        CUDA_TRY(cudaEventElapsedTime(&b, h->ev[slot][0], h->ev[slot][1]));
        tot += a;
        bars += b;
    }
    h->timing_count = 0;
    if (total_ms) *total_ms = tot;
    if (bars_kernel_ms) *bars_kernel_ms = bars;
    if (calls) *calls = n;
    return FFTL_OK;
}

// Host-buffer path: frame-range chunks, H2D / kernels / D2H overlapped on
// three streams with two staging slots.
extern "C" int fftl_stft_run(fftl_stft* h, const int16_t* pcm, int32_t channels, int64_t samples, int64_t cstride,
                             int64_t sstride, int64_t fb, int64_t fe, float* bars, int32_t* bars_i32, fftl_beat* beat)
{
    int rc = check_run_args(h, pcm, channels, samples, fb, fe, bars);
    if (rc) return rc;
    if (fe == fb) return FFTL_OK;
    if (sstride != 1 && !(sstride == channels && cstride == 1))
        return set_error(FFTL_ERR_ARG, "host path supports planar (sample_stride 1) or interleaved (sample_stride = channels, channel_stride 1) PCM");
    CUDA_TRY(cudaSetDevice(h->device));
    const fftl_config& c = h->cfg;
    const int B = c.bars, n = c.n, hop = c.hop;
    const bool interleaved = (sstride != 1);
    int64_t warm = fb;
    if (beat) {
        warm = beat_warm_start(fb);
    }
    warm &= ~(int64_t)1; // same pairing as the device path (chunks are even-sized)
    const int64_t nf_aux = fe - warm;
    if (beat) {
        size_t need = sizeof(int) * 2 * (size_t)channels * (size_t)nf_aux;
        void* pv = h->d_aux;
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        rc = ensure_bytes(&pv, &h->aux_capacity, need);
        h->d_aux = (int*)pv;
        if (rc) return rc;
    }
    // chunk size: ~32 MiB of bar output per slot (measured best on c2: 16-32 MiB tie, 8 and 64 lose 5-10 %), at least 512 frames, even
#ifdef FFTL_DEV // developer builds only (python -m fft_lines_b200.build --dev): never in the shipped library
    static const long long chunk_mib = getenv("FFTL_CHUNK_MIB") ? atoll(getenv("FFTL_CHUNK_MIB")) : 32;
    static const int dbg_skip = getenv("FFTL_E2E_SKIP") ? atoi(getenv("FFTL_E2E_SKIP")) : 0; // 1 = no H2D, 2 = no D2H
    static const bool ramp = !getenv("FFTL_CHUNK_NORAMP");
#else
    constexpr long long chunk_mib = 32;
    constexpr int dbg_skip = 0;
    constexpr bool ramp = true;
#endif
    int64_t chunk = (int64_t)(chunk_mib << 20) / ((int64_t)channels * B * 4);
    if (chunk < 512) chunk = 512;
    if (chunk > fe - warm) chunk = fe - warm;
    chunk += chunk & 1;
    const int64_t nf_total = fe - fb;
    // both staging slots get their full-chunk size up front: growing one later would free a buffer the
    // pipeline may still be using (cudaFree waits for the device, so it is safe, but it drains the pipe)
    for (int sl = 0; sl < 2; sl++) {
        if ((rc = ensure_bytes(&h->d_stage[sl][0], &h->stage_bytes[sl][0], sizeof(int16_t) * (size_t)channels * (size_t)(chunk * hop + n)))) return rc;
        if ((rc = ensure_bytes(&h->d_stage[sl][1], &h->stage_bytes[sl][1], sizeof(float) * (size_t)channels * (size_t)chunk * B))) return rc;
        if (bars_i32 && (rc = ensure_bytes(&h->d_stage[sl][2], &h->stage_bytes[sl][2], sizeof(int32_t) * (size_t)channels * (size_t)chunk * B))) return rc;
        if (beat && (rc = ensure_bytes(&h->d_stage[sl][3], &h->stage_bytes[sl][3], sizeof(fftl_beat) * (size_t)channels * (size_t)chunk))) return rc;
    }
    int slot = 0;
    int64_t nchunks = 0;
    // Chunk sizes ramp up at the start and down at the end (chunk/8, /4, /2, chunk ... chunk, half of what is left
    // until chunk/8): the first upload and the last download are the only transfers nothing overlaps with.
    const int64_t min_chunk = ((chunk / 8) + 1) & ~(int64_t)1;
    int64_t step = 0;
    for (int64_t a = warm; a < fe; a += step, slot ^= 1, nchunks++) {
        step = chunk;
        if (ramp && chunk >= 4096) {
            if (nchunks < 3) step = (chunk >> (3 - nchunks)) & ~(int64_t)1;
            const int64_t rem = fe - a;
            if (rem < 2 * step && rem > min_chunk) step = ((rem / 2) + 1) & ~(int64_t)1; // the tail halves
            if (step < 2) step = 2;
        }
        int64_t b = a + step < fe ? a + step : fe; // compute frames [a,b)
        int64_t ea = a > fb ? a : fb;                // emit frames [ea,b)
        int64_t ne = b > ea ? b - ea : 0;
        // chunks are even-sized, so b is even unless it is fe: then frame b is the partner of b-1 and its samples
        // are uploaded too, as far as the job has them (what a device-resident run reads: the signal is zero beyond)
        int64_t s0 = a * hop, s1 = ((b - 1) | 1) * hop + n; // sample span per channel
        if (s1 > samples) s1 = samples;
        int64_t span = s1 - s0;
        size_t pcm_bytes = sizeof(int16_t) * (size_t)channels * span;
        size_t bars_bytes = sizeof(float) * (size_t)channels * (size_t)(ne > 0 ? ne : 1) * B;
        size_t beat_bytes = sizeof(fftl_beat) * (size_t)channels * (size_t)(ne > 0 ? ne : 1);
        if ((rc = ensure_bytes(&h->d_stage[slot][0], &h->stage_bytes[slot][0], pcm_bytes))) return rc;
        if ((rc = ensure_bytes(&h->d_stage[slot][1], &h->stage_bytes[slot][1], bars_bytes))) return rc;
        if (bars_i32 && (rc = ensure_bytes(&h->d_stage[slot][2], &h->stage_bytes[slot][2], bars_bytes))) return rc;
        if (beat && (rc = ensure_bytes(&h->d_stage[slot][3], &h->stage_bytes[slot][3], beat_bytes))) return rc;
        int16_t* d_pcm = (int16_t*)h->d_stage[slot][0];
        // H2D (the compute that last read this slot's pcm finished before its D2H did)
        if (dbg_skip == 1) {
        } else if (interleaved) {
            CUDA_TRY(cudaMemcpyAsync(d_pcm, pcm + s0 * channels, pcm_bytes, cudaMemcpyHostToDevice, h->s_h2d));
        } else {
            // one strided copy for all channels (1024 streams must not become 1024 memcpy calls)
            CUDA_TRY(cudaMemcpy2DAsync(d_pcm, sizeof(int16_t) * span, pcm + s0, sizeof(int16_t) * cstride, sizeof(int16_t) * span,
                                       (size_t)channels, cudaMemcpyHostToDevice, h->s_h2d));
        }
        CUDA_TRY(cudaEventRecord(h->ev_h2d[slot], h->s_h2d));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_h2d[slot], 0));
        // the slot's previous D2H (two chunks back) must have drained its output buffers; waiting on the
        // device keeps the host enqueueing ahead, so the H2D engine never idles behind a D2H
        if (nchunks >= 2) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_d2h[slot], 0));
        // kernels: device pcm starts at sample s0, i.e. frame a
        {
            const int16_t* base = interleaved ? d_pcm - s0 * channels : d_pcm - s0;
            int64_t dcs = interleaved ? 1 : span;
            int64_t dss = interleaved ? channels : 1;
            float* d_bars = (float*)h->d_stage[slot][1];
            int32_t* d_i32 = bars_i32 ? (int32_t*)h->d_stage[slot][2] : nullptr;
            fftl_beat* d_beat = beat ? (fftl_beat*)h->d_stage[slot][3] : nullptr;
            if (ne > 0)
                rc = run_range(h, base, channels, dcs, dss, s1, a, ea, b, ea, ne, warm, nf_aux, d_bars, d_i32, beat != nullptr, d_beat, h->stream, nullptr);
            else // pure beat warm-up chunk: band energy / bass only, nothing emitted
                rc = run_range(h, base, channels, dcs, dss, s1, a, b, b, b, 1, warm, nf_aux, d_bars, nullptr, beat != nullptr, nullptr, h->stream, nullptr);
            if (rc) return rc;
        }
        CUDA_TRY(cudaEventRecord(h->ev_comp[slot], h->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->s_d2h, h->ev_comp[slot], 0));
        // next H2D into the other slot may start now; it must not overwrite this
        // slot's pcm until this compute is done -> h2d stream waits on ev_comp of the slot it reuses
        CUDA_TRY(cudaStreamWaitEvent(h->s_h2d, h->ev_comp[slot ^ 1], 0));
        if (ne > 0 && dbg_skip != 2) {
            // host [ch][nf_total][B] <- device [ch][ne][B]: one strided copy per output
            const size_t off_h = (size_t)(ea - fb) * B;
            CUDA_TRY(cudaMemcpy2DAsync(bars + off_h, sizeof(float) * nf_total * B, h->d_stage[slot][1], sizeof(float) * ne * B,
                                       sizeof(float) * ne * B, (size_t)channels, cudaMemcpyDeviceToHost, h->s_d2h));
            if (bars_i32)
                CUDA_TRY(cudaMemcpy2DAsync(bars_i32 + off_h, sizeof(int32_t) * nf_total * B, h->d_stage[slot][2], sizeof(int32_t) * ne * B,
                                           sizeof(int32_t) * ne * B, (size_t)channels, cudaMemcpyDeviceToHost, h->s_d2h));
            if (beat)
                CUDA_TRY(cudaMemcpy2DAsync(beat + (ea - fb), sizeof(fftl_beat) * nf_total, h->d_stage[slot][3], sizeof(fftl_beat) * ne,
                                           sizeof(fftl_beat) * ne, (size_t)channels, cudaMemcpyDeviceToHost, h->s_d2h));
        }
        CUDA_TRY(cudaEventRecord(h->ev_d2h[slot], h->s_d2h));
    }
    CUDA_TRY(cudaStreamSynchronize(h->s_d2h));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->s_h2d));
    return FFTL_OK;
}

// ------------------------------------------------------------------ streams
struct fftl_stream {
    fftl_stft* core;        // tables, launcher, aux (w, bass) buffers
    int streams, move_mode;
    int16_t* d_window;      // [streams][n]
    int* d_w_ring;          // [streams][160]
    int* d_w_sum;           // [streams]
    int* d_tempo_ring;      // [streams][256]
    double2* d_tempo_x;     // [streams][129] transform of the tempo ring, carried from tick to tick (stream_beat_carry_kernel)
    StreamState* d_state;
    // staging for the host-buffer variant
    int16_t* d_fresh;
    size_t fresh_cap;
    int16_t* d_conv;        // int16 image of a float capture push (fftl_stream_push_device_f32)
    size_t conv_cap;
    float* d_bars;
    int32_t* d_bars_i32;
    fftl_beat* d_beat;
    float* d_hist;          // [streams][hist_depth][bars] or NULL
    int hist_depth;
    int tick_kernel;        // FFTL_TICK_AUTO / FFTL_TICK_THREE_LAUNCHES
    int64_t launches, fused_ticks;
};

extern "C" void fftl_stream_destroy(fftl_stream* s)
{
    if (!s) return;
    if (s->core) {
        cudaSetDevice(s->core->device);
        cudaStreamSynchronize(s->core->stream);
    }
    cudaFree(s->d_window);
    cudaFree(s->d_w_ring);
    cudaFree(s->d_w_sum);
    cudaFree(s->d_tempo_ring);
    cudaFree(s->d_tempo_x);
    cudaFree(s->d_state);
    cudaFree(s->d_fresh);
    cudaFree(s->d_conv);
    cudaFree(s->d_bars);
    cudaFree(s->d_bars_i32);
    cudaFree(s->d_beat);
    cudaFree(s->d_hist);
    cudaGetLastError();
    fftl_stft_destroy(s->core);
    free(s);
}

extern "C" int fftl_stream_reset(fftl_stream* s)
{
    if (!s) return set_error(FFTL_ERR_ARG, "stream handle is NULL");
    CUDA_TRY(cudaSetDevice(s->core->device));
    const size_t S = (size_t)s->streams;
    CUDA_TRY(cudaDeviceSynchronize()); // pushes still in flight on other streams finish first
    CUDA_TRY(cudaMemset(s->d_window, 0, sizeof(int16_t) * S * s->core->cfg.n));
    CUDA_TRY(cudaMemset(s->d_w_ring, 0, sizeof(int) * S * FFTL_BEAT_SAMPLES));
    CUDA_TRY(cudaMemset(s->d_w_sum, 0, sizeof(int) * S));
    CUDA_TRY(cudaMemset(s->d_tempo_ring, 0, sizeof(int) * S * FFTL_TEMPO_RING));
    CUDA_TRY(cudaMemset(s->d_tempo_x, 0, sizeof(double2) * S * STREAM_CARRY_BINS)); // the exact transform of the zero ring
    CUDA_TRY(cudaMemset(s->d_state, 0, sizeof(StreamState)));
    if (s->d_hist) CUDA_TRY(cudaMemset(s->d_hist, 0, sizeof(float) * S * s->hist_depth * s->core->cfg.bars));
    // the memsets run on the legacy stream, pushes on non-blocking / caller streams: finish them before returning
    CUDA_TRY(cudaDeviceSynchronize());
    return FFTL_OK;
}

extern "C" int fftl_stream_set_history(fftl_stream* s, int32_t depth)
{
    if (!s) return set_error(FFTL_ERR_ARG, "stream handle is NULL");
    if (depth < 0 || depth > 4096) return set_error(FFTL_ERR_ARG, "history depth %d out of 0..4096", depth);
    CUDA_TRY(cudaSetDevice(s->core->device));
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(s->d_hist);
    s->d_hist = nullptr;
    s->hist_depth = 0;
    if (depth > 0) {
        size_t bytes = sizeof(float) * (size_t)s->streams * depth * s->core->cfg.bars;
        cudaError_t e = cudaMalloc(&s->d_hist, bytes);
        if (e != cudaSuccess) return set_error(FFTL_ERR_NOMEM, "history allocation failed: %s", cudaGetErrorString(e));
        CUDA_TRY(cudaMemset(s->d_hist, 0, bytes));
        CUDA_TRY(cudaDeviceSynchronize());
        s->hist_depth = depth;
    }
    return FFTL_OK;
}

extern "C" int fftl_stream_history_device(fftl_stream* s, float* out, void* cuda_stream)
{
    if (!s || !out) return set_error(FFTL_ERR_ARG, "NULL argument");
    if (!s->d_hist) return set_error(FFTL_ERR_STATE, "history is off: call fftl_stream_set_history first");
    CUDA_TRY(cudaSetDevice(s->core->device));
    long long total = (long long)s->streams * s->hist_depth * s->core->cfg.bars;
    long long blocks = (total + 255) / 256;
    if (blocks > 8192) blocks = 8192;
    stream_history_read_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(s->d_hist, out, s->d_state, s->streams, s->hist_depth,
                                                                                       s->core->cfg.bars);
    CUDA_TRY(cudaGetLastError());
    s->launches++;
    return FFTL_OK;
}

extern "C" int fftl_stream_waveform_device(fftl_stream* s, int32_t points, float scale, float offset, int32_t* heights, void* cuda_stream)
{
    if (!s || !heights) return set_error(FFTL_ERR_ARG, "NULL argument");
    if (points < 1 || points > s->core->cfg.n) return set_error(FFTL_ERR_ARG, "points=%d outside 1..n", points);
    CUDA_TRY(cudaSetDevice(s->core->device));
    long long total = (long long)s->streams * points;
    long long blocks = (total + 255) / 256;
    if (blocks > 8192) blocks = 8192;
    stream_waveform_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(s->d_window, s->core->cfg.n, points, scale, offset, heights,
                                                                                   s->streams);
    CUDA_TRY(cudaGetLastError());
    s->launches++;
    return FFTL_OK;
}

extern "C" int fftl_stream_create(const fftl_config* cfg, int32_t streams, int32_t move_mode, int device, fftl_stream** out)
{
    if (!out) return set_error(FFTL_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (streams < 1 || streams > (1 << 22)) return set_error(FFTL_ERR_ARG, "streams=%d out of range", streams);
    if (move_mode != FFTL_MOVE_EXACT && move_mode != FFTL_MOVE_REFERENCE) return set_error(FFTL_ERR_ARG, "unknown move_mode %d", move_mode);
    fftl_stream* s = (fftl_stream*)calloc(1, sizeof(fftl_stream));
    if (!s) return set_error(FFTL_ERR_NOMEM, "out of host memory");
    int rc = fftl_stft_create(cfg, device, &s->core);
    if (rc) {
        free(s);
        return rc;
    }
    s->streams = streams;
    s->move_mode = move_mode;
    const size_t S = (size_t)streams;
    const int B = cfg->bars;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&s->d_window, sizeof(int16_t) * S * cfg->n);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_w_ring, sizeof(int) * S * FFTL_BEAT_SAMPLES);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_w_sum, sizeof(int) * S);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_tempo_ring, sizeof(int) * S * FFTL_TEMPO_RING);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_tempo_x, sizeof(double2) * S * STREAM_CARRY_BINS);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_state, sizeof(StreamState));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_bars, sizeof(float) * S * B);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_bars_i32, sizeof(int32_t) * S * B);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_beat, sizeof(fftl_beat) * S);
    void* pv = s->core->d_aux;
    if (e == cudaSuccess && ensure_bytes(&pv, &s->core->aux_capacity, sizeof(int) * 2 * S)) e = cudaErrorMemoryAllocation;
    s->core->d_aux = (int*)pv;
    if (e != cudaSuccess) {
        set_error(FFTL_ERR_NOMEM, "stream state allocation failed: %s", cudaGetErrorString(e));
        fftl_stream_destroy(s);
        return FFTL_ERR_NOMEM;
    }
    rc = fftl_stream_reset(s);
    if (rc) {
        fftl_stream_destroy(s);
        return rc;
    }
    *out = s;
    return FFTL_OK;
}

extern "C" int64_t fftl_stream_launch_count(const fftl_stream* s) { return s ? s->launches : 0; }
extern "C" int64_t fftl_stream_fused_tick_count(const fftl_stream* s) { return s ? s->fused_ticks : 0; }
extern "C" int fftl_stream_set_tick_kernel(fftl_stream* s, int32_t which)
{
    if (!s) return set_error(FFTL_ERR_ARG, "stream handle is NULL");
    if (which != FFTL_TICK_AUTO && which != FFTL_TICK_THREE_LAUNCHES) return set_error(FFTL_ERR_ARG, "unknown tick kernel selection %d", which);
    s->tick_kernel = which;
    return FFTL_OK;
}

// One kernel per tick (stream_tick.cuh): n = 1024 / 2048, exact window move.
template <int N>
static cudaError_t launch_stream_tick(const StftParams& p, const StreamTickParams& k, const StreamCarry& q, int sg_half, cudaStream_t st)
{
    stream_tick_kernel<N><<<(k.streams + 1) / 2, N / 4, stream_tick_smem<N>(p.bars_n, sg_half), st>>>(p, k, q);
    return cudaGetLastError();
}

extern "C" int fftl_stream_push_device(fftl_stream* s, const int16_t* fresh, int32_t count, int64_t stride, float* bars, int32_t* bars_i32,
                                       fftl_beat* beat, void* cuda_stream)
{
    if (!s) return set_error(FFTL_ERR_ARG, "stream handle is NULL");
    if (!fresh || !bars) return set_error(FFTL_ERR_ARG, "fresh and bars must not be NULL");
    if (count < 1) return set_error(FFTL_ERR_ARG, "count=%d: a push needs at least one new sample", count);
    fftl_stft* h = s->core;
    const fftl_config& c = h->cfg;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const float dt = (float)count / c.sample_rate; // timeDelta of this tick (fft_lines.c:265 uses wall clock)
    if (count > c.n) { // only the newest n samples can matter
        if (s->move_mode == FFTL_MOVE_REFERENCE) return set_error(FFTL_ERR_ARG, "count > n is undefined for the reference's MoveArray");
        fresh += count - c.n;
        count = c.n;
    }
    if (s->tick_kernel == FFTL_TICK_AUTO && s->move_mode == FFTL_MOVE_EXACT && (c.n == 1024 || c.n == 2048) &&
        stream_tick_smem<2048>(c.bars, c.sg_half) <= 28 * 1024) {
        StftParams p;
        memset(&p, 0, sizeof(p));
        p.window = h->d_window;
        p.tw = h->d_tw;
        p.lo = h->d_lo;
        p.hi = h->d_hi;
        p.bar_inv = h->d_bar_inv;
        p.sg = h->d_sg;
        p.bars = bars;
        p.bars_i32 = bars_i32;
        p.aux_w = h->d_aux;
        p.aux_bass = h->d_aux + s->streams;
        p.bars_n = c.bars;
        p.sg_half = h->d_sg ? c.sg_half : 0;
        p.strict = c.strict_int;
        p.bass_bin = c.bass_bin;
        for (int i = 0; i < 4; i++) p.beat_bars[i] = c.beat_bars[i];
        p.zoom = c.zoom;
        p.circle_add = c.circle ? 10.0f : 0.0f; // fft_lines.c:204-207
        StreamTickParams k;
        k.window = s->d_window;
        k.fresh = fresh;
        k.fresh_stride = stride;
        k.count = count;
        k.streams = s->streams;
        StreamCarry cq;
        cq.aux_w = p.aux_w;
        cq.aux_bass = p.aux_bass;
        cq.w_ring = s->d_w_ring;
        cq.w_sum = s->d_w_sum;
        cq.tempo_ring = s->d_tempo_ring;
        cq.x = s->d_tempo_x;
        cq.tw256d = h->d_tw256d;
        cq.state = s->d_state;
        cq.out = beat ? beat : s->d_beat;
        cq.streams = s->streams;
        cq.n_times_dt = (float)c.n * dt;
        cq.strict = c.strict_int;
        cq.finish = s->d_hist ? 0 : 1; // the history push, when enabled, is the last kernel
        CUDA_TRY(c.n == 1024 ? launch_stream_tick<1024>(p, k, cq, p.sg_half, st) : launch_stream_tick<2048>(p, k, cq, p.sg_half, st));
        s->launches++;
        s->fused_ticks++;
        if (s->d_hist) {
            long long total = (long long)s->streams * c.bars;
            long long blocks = (total + 255) / 256;
            if (blocks > 4096) blocks = 4096;
            stream_history_push_kernel<<<(unsigned)blocks, 256, 0, st>>>(s->d_hist, bars, s->d_state, s->streams, s->hist_depth, c.bars);
            CUDA_TRY(cudaGetLastError());
            s->launches++;
        }
        return FFTL_OK;
    }
    if (s->move_mode != FFTL_MOVE_REFERENCE && (count & 7) == 0 && (c.n & 7) == 0 && (stride & 7) == 0 && ((uintptr_t)fresh & 15) == 0)
        stream_slide_vec_kernel<<<s->streams, 256, 0, st>>>(s->d_window, fresh, c.n, count, stride);
    else
        stream_slide_kernel<<<s->streams, 256, 0, st>>>(s->d_window, fresh, c.n, count, stride, s->move_mode == FFTL_MOVE_REFERENCE);
    CUDA_TRY(cudaGetLastError());
    // beat stage of the tick (stream_carry.cuh).  (Run as the epilogue of the bars kernel for 1024-point windows it saved
    // 0.26 us of 16.65 per tick of 1024 streams -- measured, not kept.)
    StreamCarry cp;
    cp.aux_w = h->d_aux;
    cp.aux_bass = h->d_aux + s->streams;
    cp.w_ring = s->d_w_ring;
    cp.w_sum = s->d_w_sum;
    cp.tempo_ring = s->d_tempo_ring;
    cp.x = s->d_tempo_x;
    cp.tw256d = h->d_tw256d;
    cp.state = s->d_state;
    cp.out = beat ? beat : s->d_beat;
    cp.streams = s->streams;
    cp.n_times_dt = (float)c.n * dt;
    cp.strict = c.strict_int;
    cp.finish = s->d_hist ? 0 : 1; // the history push, when enabled, is the last kernel
    int rc;
    if (c.n == 16384 || c.n == 8192 || c.n == 4096 || c.n == 2048 || c.n == 1024) {
        // The windows [streams][n] are one channel of streams*n samples framed with hop n: frame f = stream f.  The
        // fast path then packs streams 2k, 2k+1 into its shared sub-transforms (half the transform work of one
        // frame per channel); outputs [1][streams][bars] and aux [1][streams] have the layout of [streams][1][..].
        const long long S = (long long)s->streams * c.n;
        rc = run_range(h, s->d_window, 1, S, 1, S, 0, 0, s->streams, 0, s->streams, 0, s->streams, bars, bars_i32, true, nullptr, st,
                       nullptr, false, c.n);
    } else {
        // one frame per stream: channels = streams, samples = n
        rc = run_range(h, s->d_window, s->streams, c.n, 1, c.n, 0, 0, 1, 0, 1, 0, 1, bars, bars_i32, true, nullptr, st, nullptr, true);
    }
    if (rc) return rc;
    stream_beat_carry_kernel<<<(s->streams + 1) / 2, 256, 0, st>>>(cp);
    CUDA_TRY(cudaGetLastError());
    if (s->d_hist) {
        long long total = (long long)s->streams * c.bars;
        long long blocks = (total + 255) / 256;
        if (blocks > 4096) blocks = 4096;
        stream_history_push_kernel<<<(unsigned)blocks, 256, 0, st>>>(s->d_hist, bars, s->d_state, s->streams, s->hist_depth, c.bars);
        CUDA_TRY(cudaGetLastError());
        s->launches++;
    }
    s->launches += 3;
    return FFTL_OK;
}

extern "C" int fftl_stream_push_device_f32(fftl_stream* s, const float* fresh, int32_t count, int64_t stride, int64_t sstride, float* bars,
                                           int32_t* bars_i32, fftl_beat* beat, void* cuda_stream)
{
    if (!s) return set_error(FFTL_ERR_ARG, "stream handle is NULL");
    if (!fresh || !bars) return set_error(FFTL_ERR_ARG, "fresh and bars must not be NULL");
    if (count < 1 || sstride < 1) return set_error(FFTL_ERR_ARG, "count=%d, sample_stride=%lld: need at least one new sample and a positive stride", count, (long long)sstride);
    CUDA_TRY(cudaSetDevice(s->core->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t need = sizeof(int16_t) * (size_t)s->streams * (((size_t)count + 7) & ~(size_t)7);
    if (s->conv_cap < need) {
        // grows on the first push of a given size (not inside a CUDA-graph capture: push once before capturing)
        CUDA_TRY(cudaStreamSynchronize(st));
        void* pv = s->d_conv;
        int rc = ensure_bytes(&pv, &s->conv_cap, need);
        s->d_conv = (int16_t*)pv;
        if (rc) return rc;
    }
    const int64_t row = ((int64_t)count + 7) & ~(int64_t)7; // rows stay 16-byte aligned for the vector window move
    long long blocks = ((long long)s->streams * count + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    stream_f32_to_i16_kernel<<<(unsigned)blocks, 256, 0, st>>>(fresh, s->d_conv, s->streams, count, stride, sstride, row);
    CUDA_TRY(cudaGetLastError());
    s->launches++;
    return fftl_stream_push_device(s, s->d_conv, count, row, bars, bars_i32, beat, cuda_stream);
}

extern "C" int fftl_stream_push(fftl_stream* s, const int16_t* fresh, int32_t count, int64_t stride, float* bars, int32_t* bars_i32,
                                fftl_beat* beat)
{
    if (!s) return set_error(FFTL_ERR_ARG, "stream handle is NULL");
    if (!fresh || !bars) return set_error(FFTL_ERR_ARG, "fresh and bars must not be NULL");
    if (count < 1) return set_error(FFTL_ERR_ARG, "count=%d: a push needs at least one new sample", count);
    fftl_stft* h = s->core;
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t S = (size_t)s->streams;
    const int B = h->cfg.bars;
    size_t need = sizeof(int16_t) * S * count;
    void* pv = s->d_fresh;
    if (s->fresh_cap < need) CUDA_TRY(cudaStreamSynchronize(h->stream));
    int rc = ensure_bytes(&pv, &s->fresh_cap, need);
    s->d_fresh = (int16_t*)pv;
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(s->d_fresh, sizeof(int16_t) * count, fresh, sizeof(int16_t) * stride, sizeof(int16_t) * count, S,
                               cudaMemcpyHostToDevice, h->stream));
    rc = fftl_stream_push_device(s, s->d_fresh, count, count, s->d_bars, bars_i32 ? s->d_bars_i32 : nullptr, s->d_beat, h->stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(bars, s->d_bars, sizeof(float) * S * B, cudaMemcpyDeviceToHost, h->stream));
    if (bars_i32) CUDA_TRY(cudaMemcpyAsync(bars_i32, s->d_bars_i32, sizeof(int32_t) * S * B, cudaMemcpyDeviceToHost, h->stream));
    if (beat) CUDA_TRY(cudaMemcpyAsync(beat, s->d_beat, sizeof(fftl_beat) * S, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FFTL_OK;
}

// ------------------------------------------------------------------ callers / sinks either side of the path
extern "C" void fftl_settings_default(fftl_settings* s)
{
    // settingsFile.h:10-22
    s->bar_count = 500;
    s->zoom = 0.0001f;
    s->color_sel = 0;
    s->dofft = 1;
    s->background = 1;
    s->gradient = 1;
    s->ignore_serial = 1;
    s->circle = 0;
    s->waveform = 0;
    s->stereo = 0;
    s->beat_detection = 1;
}

extern "C" int fftl_settings_parse(const char* text, int64_t len, fftl_settings* out)
{
    if (!text || !out) return set_error(FFTL_ERR_ARG, "settings: NULL argument");
    fftl_settings_default(out);
    // readSettings leaves everything at its default for an empty or oversized file (settingsFile.c:178-181)
    if (len <= 0 || len > 1000) return FFTL_OK;
    std::vector<char> buf(text, text + len);
    buf.push_back(0);
    char* p = buf.data();
    long bc = strtol(p, &p, 10);
    out->bar_count = (bc < 100 || bc > 1000) ? 500 : (int32_t)bc;                 // :200-202
    float z = strtof(p, &p);
    out->zoom = (z < 0.00001f || z > 0.001f) ? 0.0001f : z;                         // :204-206
    long long cs = strtoll(p, &p, 10);
    out->color_sel = (cs < 0 || cs >= 6) ? 0 : (int32_t)cs;                        // :208-210
    // the remaining eight are C `bool`s in the reference: any non-zero number reads as true (:213-243)
    int32_t* flags[8] = {&out->dofft, &out->background, &out->gradient, &out->ignore_serial,
                         &out->circle, &out->waveform, &out->stereo, &out->beat_detection};
    for (int i = 0; i < 8; i++) *flags[i] = strtol(p, &p, 10) != 0 ? 1 : 0;
    return FFTL_OK;
}

extern "C" int fftl_settings_format(const fftl_settings* s, char* buf, int64_t cap)
{
    if (!s || !buf || cap <= 0) return -1;
    int n = snprintf(buf, (size_t)cap, "%d %0.6f %lld %d %d %d %d %d %d %d %d", s->bar_count, (double)s->zoom, (long long)s->color_sel,
                     s->dofft, s->background, s->gradient, s->ignore_serial, s->circle, s->waveform, s->stereo, s->beat_detection); // :376
    return (n < 0 || n >= cap) ? -1 : n;
}

extern "C" void fftl_settings_to_config(const fftl_settings* s, int32_t n, int32_t hop, float sample_rate, fftl_config* cfg)
{
    fftl_config_reference(cfg, n, hop, s->bar_count);
    cfg->zoom = s->zoom;
    cfg->circle = s->circle;
    cfg->sample_rate = sample_rate;
}

extern "C" uint32_t fftl_color_index(int32_t i, int32_t count)
{
    // RANGE255 (drawBar2D.cpp:33): float division by (float - 1.0) is done in double, cast to float, then to unsigned
    return (unsigned int)((float)(((float)i / ((float)count - 1.0)) * 254.0));
}

extern "C" int fftl_led_bytes_device(const int32_t* heights, int64_t frames, int32_t bars, int32_t led_bar, uint8_t* bytes,
                                     uint8_t* valid, void* cuda_stream)
{
    if (!heights || !bytes || frames < 0 || bars < 1 || led_bar < 0 || led_bar >= bars) return set_error(FFTL_ERR_ARG, "bad LED arguments");
    if (frames == 0) return FFTL_OK;
    long long blocks = (frames + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    led_bytes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(heights, frames, bars, led_bar, bytes, valid);
    CUDA_TRY(cudaGetLastError());
    return FFTL_OK;
}

extern "C" int fftl_marker_sweep_device(const fftl_beat* beat, int64_t frames, float time_delta, int32_t start_pos, int32_t start_dir,
                                        int32_t* pos_out, int32_t* dir_out, void* cuda_stream)
{
    if (!beat || !pos_out || frames < 0) return set_error(FFTL_ERR_ARG, "bad marker arguments");
    if (start_pos < 0 || start_pos > 2048 || start_dir < -1 || start_dir > 1) return set_error(FFTL_ERR_ARG, "marker start state out of range");
    if (frames == 0) return FFTL_OK;
    marker_sweep_kernel<<<1, 1024, 0, (cudaStream_t)cuda_stream>>>(beat, frames, time_delta, start_pos, start_dir, pos_out, dir_out);
    CUDA_TRY(cudaGetLastError());
    return FFTL_OK;
}

extern "C" void* fftl_host_alloc(uint64_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        set_error(FFTL_ERR_NOMEM, "cudaHostAlloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
extern "C" void fftl_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------ probes
// Dense FP32 FMA rate of the current device (2 flop per FMA), best of `reps`.
extern "C" int fftl_measure_fp32_peak(double* tflops, int32_t reps)
{
    if (!tflops) return set_error(FFTL_ERR_ARG, "tflops is NULL");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 4096, blocks = sms * 8, threads = 256;
    double best = 0;
    if (reps < 1) reps = 1;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(e0, 0);
        fp32_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(e1, 0);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaFree(d);
            return set_error(FFTL_ERR_CUDA, "fp32 probe failed: %s", cudaGetErrorString(e));
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        double tf = fl / ((double)ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf; // first launch is the warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return FFTL_OK;
}

// ------------------------------------------------------------------ synth
static unsigned long long frac64(double x)
{
    const double two64 = 18446744073709551616.0;
    if (x < 0) return 0ull - frac64(-x);
    x -= floor(x);
    return (unsigned long long)(x * two64);
}

extern "C" int fftl_synth_pcm_device(int16_t* pcm, int32_t channels, int64_t samples, int64_t cstride, int64_t first, float sample_rate,
                                     uint32_t seed, void* cuda_stream)
{
    if (!pcm || channels < 1 || channels > 65535 || samples < 0 || !(sample_rate > 0)) return set_error(FFTL_ERR_ARG, "bad synth arguments");
    if (samples == 0) return FFTL_OK;
    const double fs = (double)sample_rate;
    unsigned long long period = (unsigned long long)(10.0 * fs);
    if (period == 0) period = 1;
    const double T = (double)period / fs;
    std::vector<SynthChannel> hc(channels);
    for (int c = 0; c < channels; c++) {
        double f0 = (c & 1) ? 20000.0 : 20.0, f1 = (c & 1) ? 20.0 : 20000.0;
        hc[c].A = frac64(f0 / fs);
        hc[c].B = frac64((f1 - f0) / (2.0 * T * fs * fs));
        hc[c].tone_step = frac64(94.0 / fs);
        hc[c].am_step = frac64(2.0 / fs);
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    SynthChannel* dc = nullptr;
    CUDA_TRY(cudaMalloc(&dc, sizeof(SynthChannel) * channels));
    cudaError_t e = cudaMemcpyAsync(dc, hc.data(), sizeof(SynthChannel) * channels, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        SynthParams sp;
        sp.pcm = pcm;
        sp.samples = samples;
        sp.channel_stride = cstride;
        sp.first_sample = first;
        sp.period = period;
        sp.seed = seed;
        sp.channels = channels;
        long long blocks = (samples + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        synth_pcm_kernel<<<dim3((unsigned)blocks, (unsigned)channels), 256, 0, st>>>(sp, dc);
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaStreamSynchronize(st); // hc / dc lifetimes
    cudaFree(dc);
    if (e != cudaSuccess) return set_error(FFTL_ERR_CUDA, "synth launch failed: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return set_error(FFTL_ERR_CUDA, "synth failed: %s", cudaGetErrorString(e2));
    return FFTL_OK;
}

// ================================================================== compat
// Drop-in layer for the reference's single-frame API.  One global context,
// not thread-safe, exactly like the reference's globals (fft_calculate.c:6-7,
// beatDetector.c:14-28).
extern "C" {
float beatMovementPerSec = 0; // beatDetector.c:25
int beatposition = 0;         // beatDetector.c:26
float timeDelta = 0;          // beatDetector.c:27
int direction;                // beatDetector.c:28
// The application's tempo ring (settingsFile.c:83).  The reference writes into
// it; if the host program defines it we mirror the ring there, otherwise the
// library's device-resident ring is the only copy.
extern int* bassBeatBuffer __attribute__((weak));
}

struct CompatPlan {
    bool live;
    int n;
    fftwf_complex* in;
    fftwf_complex* out;
    float2* d_in;
    float2* d_out;
    float2* d_tw;
};
static CompatPlan g_plan = {false, 0, nullptr, nullptr, nullptr, nullptr, nullptr};
static CompatPlan g_plan256 = {false, 0, nullptr, nullptr, nullptr, nullptr, nullptr};
static int g_compat_n = 2048; // reference N, global.h:3
static cudaStream_t g_cstream = nullptr;
static CompatBeatState* g_beat_state = nullptr;
static int g_beat_value = 0;
static int64_t g_compat_launches = 0;

extern "C" int fftl_set_n(int32_t n)
{
    if (!pow2(n) || n < 256 || n > 16384) return set_error(FFTL_ERR_ARG, "fftl_set_n(%d): need a power of two in 256..16384", n);
    if (g_plan.live) return set_error(FFTL_ERR_STATE, "fftl_set_n while a plan is live; call cleanupfft first");
    g_compat_n = n;
    return FFTL_OK;
}
extern "C" int fftl_get_n(void) { return g_compat_n; }
extern "C" int64_t fftl_compat_launch_count(void) { return g_compat_launches; }

static int compat_stream()
{
    if (!g_cstream) CUDA_TRY(cudaStreamCreateWithFlags(&g_cstream, cudaStreamNonBlocking));
    return FFTL_OK;
}

static int compat_beat_state()
{
    if (g_beat_state) return FFTL_OK;
    if (compat_stream()) return FFTL_ERR_CUDA;
    CUDA_TRY(cudaMalloc(&g_beat_state, sizeof(CompatBeatState)));
    CUDA_TRY(cudaMemset(g_beat_state, 0, sizeof(CompatBeatState)));
    CUDA_TRY(cudaDeviceSynchronize()); // g_cstream is non-blocking: it does not order against the legacy stream's memset
    return FFTL_OK;
}

extern "C" void fftl_compat_reset_beat(void)
{
    if (g_beat_state) {
        cudaMemset(g_beat_state, 0, sizeof(CompatBeatState));
        cudaDeviceSynchronize();
    }
    g_beat_value = 0;
    beatMovementPerSec = 0;
    beatposition = 0;
    direction = 0;
}

static void plan_release(CompatPlan* p)
{
    if (p->d_in) cudaFree(p->d_in);
    if (p->d_out) cudaFree(p->d_out);
    if (p->d_tw) cudaFree(p->d_tw);
    memset(p, 0, sizeof(*p));
    cudaGetLastError();
}

static int plan_init(CompatPlan* p, int n, fftwf_complex* in, fftwf_complex* out)
{
    plan_release(p);
    if (!in || !out) return set_error(FFTL_ERR_ARG, "initializefft: NULL buffer");
    int ndev = fftl_device_count();
    if (ndev < 1) return set_error(FFTL_ERR_CUDA, "no CUDA device (libfftlines_cuda has no CPU fallback)");
    if (compat_stream()) return FFTL_ERR_CUDA;
    std::vector<float2> tw;
    fill_twiddles(tw, n);
    CUDA_TRY(cudaMalloc(&p->d_in, sizeof(float2) * n));
    CUDA_TRY(cudaMalloc(&p->d_out, sizeof(float2) * n));
    CUDA_TRY(cudaMalloc(&p->d_tw, sizeof(float2) * n));
    CUDA_TRY(cudaMemcpy(p->d_tw, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize()); // pageable upload on the legacy stream vs the non-blocking g_cstream
    p->n = n;
    p->in = in;   // the plan is bound to the caller's buffers (fft_calculate.c:24)
    p->out = out;
    p->live = true;
    return FFTL_OK;
}

static int plan_execute(CompatPlan* p)
{
    if (!p->live) return set_error(FFTL_ERR_STATE, "executefft before initializefft (the reference would dereference a NULL plan)");
    // fft_calculate.c:29: the arguments are ignored, the bound buffers are used
    CUDA_TRY(cudaMemcpyAsync(p->d_in, p->in, sizeof(float2) * p->n, cudaMemcpyHostToDevice, g_cstream));
    CUDA_TRY(dispatch_c2c(p->n, p->d_in, p->d_out, p->d_tw, g_cstream));
    g_compat_launches++;
    CUDA_TRY(cudaMemcpyAsync(p->out, p->d_out, sizeof(float2) * p->n, cudaMemcpyDeviceToHost, g_cstream));
    CUDA_TRY(cudaStreamSynchronize(g_cstream));
    return FFTL_OK;
}

extern "C" void initializefft(fftwf_complex* in, fftwf_complex* out) { plan_init(&g_plan, g_compat_n, in, out); }
extern "C" void executefft(fftwf_complex* in, fftwf_complex* out)
{
    (void)in;
    (void)out;
    plan_execute(&g_plan);
}
extern "C" void cleanupfft(void) { plan_release(&g_plan); }
extern "C" void initializefft256(fftwf_complex* in, fftwf_complex* out) { plan_init(&g_plan256, FFTL_TEMPO_RING, in, out); }
extern "C" void executefft256(fftwf_complex* in, fftwf_complex* out)
{
    (void)in;
    (void)out;
    plan_execute(&g_plan256);
}
extern "C" void cleanupfft256(void) { plan_release(&g_plan256); }

static int add_beat_sample_impl(int w)
{
    if (fftl_device_count() < 1) return set_error(FFTL_ERR_CUDA, "no CUDA device (libfftlines_cuda has no CPU fallback)");
    if (compat_beat_state()) return FFTL_ERR_CUDA;
    compat_add_beat_sample_kernel<<<1, 32, 0, g_cstream>>>(g_beat_state, w);
    CUDA_TRY(cudaGetLastError());
    g_compat_launches++;
    CUDA_TRY(cudaMemcpyAsync(&g_beat_value, &g_beat_state->beat_value, sizeof(int), cudaMemcpyDeviceToHost, g_cstream));
    CUDA_TRY(cudaStreamSynchronize(g_cstream));
    return FFTL_OK;
}
extern "C" void AddBeatSample(int weightedAverage) { add_beat_sample_impl(weightedAverage); }
extern "C" int GetBeatValue(void) { return g_beat_value; }

static int bass_beat_impl(fftwf_complex* output, fftwf_complex* beatinput, fftwf_complex* beatoutput)
{
    if (!output) return set_error(FFTL_ERR_ARG, "BassBeatDetector: output is NULL");
    if (fftl_device_count() < 1) return set_error(FFTL_ERR_CUDA, "no CUDA device (libfftlines_cuda has no CPU fallback)");
    if (!g_plan256.live) return set_error(FFTL_ERR_STATE, "BassBeatDetector before initializefft256");
    if (compat_beat_state()) return FFTL_ERR_CUDA;
    // beatDetector.c:105: N * timeDelta, left to right in float
    float n_dt = (float)g_compat_n * timeDelta;
    compat_bass_beat_kernel<<<1, 32, 0, g_cstream>>>(g_beat_state, output[4][REAL], output[4][IMAG], n_dt, g_plan256.d_tw, 1);
    CUDA_TRY(cudaGetLastError());
    g_compat_launches++;
    static CompatBeatState host_copy; // 6 KB; the drop-in API is single-threaded like the reference
    CUDA_TRY(cudaMemcpyAsync(&host_copy, g_beat_state, sizeof(CompatBeatState), cudaMemcpyDeviceToHost, g_cstream));
    CUDA_TRY(cudaStreamSynchronize(g_cstream));
    beatMovementPerSec = host_copy.movement_per_sec;
    // the reference runs plan256, which is bound to the buffers given to
    // initializefft256 (beatDetector.c:61-67 fill `beatinput`, the plan writes its own out)
    if (beatinput) memcpy(beatinput, host_copy.beat_in, sizeof(float2) * FFTL_TEMPO_RING);
    if (g_plan256.out) memcpy(g_plan256.out, host_copy.beat_out, sizeof(float2) * FFTL_TEMPO_RING);
    if (beatoutput && (void*)beatoutput != (void*)g_plan256.out) memcpy(beatoutput, host_copy.beat_out, sizeof(float2) * FFTL_TEMPO_RING);
    if (&bassBeatBuffer && bassBeatBuffer) memcpy(bassBeatBuffer, host_copy.tempo_ring, sizeof(int) * FFTL_TEMPO_RING);
    return FFTL_OK;
}
extern "C" void BassBeatDetector(fftwf_complex* input, fftwf_complex* output, fftwf_complex* beatinput, fftwf_complex* beatoutput)
{
    (void)input; // unused by the reference as well
    bass_beat_impl(output, beatinput, beatoutput);
}

"""Batched STFT -> bars (+ beat) through libfftlines_cuda's C-ABI.

``BatchedSpectrum`` is the many-frames generalisation of the reference's
per-tick chain (reference fft_lines/fft_lines.c:186-236,281).  Host arrays are
numpy; device arrays are torch CUDA tensors whose ``data_ptr()`` is handed to
the C-ABI (torch is only the device allocator / stream provider here).
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import FftlConfig

BEAT_DTYPE = np.dtype([("w", "<i4"), ("thr", "<i4"), ("beat_value", "<i4"), ("flag", "<i4"),
                       ("bass", "<i4"), ("peak_index", "<i4"), ("movement_per_sec", "<f4"),
                       ("reserved", "<i4")])
assert BEAT_DTYPE.itemsize == C.sizeof(_capi.FftlBeat)


def device_count():
    return _capi.lib().fftl_device_count()


def measure_fp32_peak(reps=5):
    """Measured FP32 FMA rate of the current CUDA device in TFLOP/s."""
    v = C.c_double()
    _capi.check(_capi.lib().fftl_measure_fp32_peak(C.byref(v), reps))
    return v.value


def reference_config(n=2048, hop=512, bars=100, zoom=1e-4, sample_rate=48000.0, circle=0):
    """Parity mode: what the reference computes (RECT window, bar i = bin i, no smoothing)."""
    c = FftlConfig()
    _capi.lib().fftl_config_reference(C.byref(c), n, hop, bars)
    c.zoom, c.sample_rate, c.circle = zoom, sample_rate, circle
    return c


def extended_config(n=4096, hop=512, bars=200, zoom=1e-4, sample_rate=48000.0, fmin=20.0, fmax=20000.0,
                    sg_half=3, sg_degree=2, circle=0):
    """north_star's chain: Hann window, log-frequency bins, Savitzky-Golay smoothing."""
    c = FftlConfig()
    _capi.lib().fftl_config_extended(C.byref(c), n, hop, bars, sample_rate, fmin, fmax)
    c.zoom, c.circle, c.sg_half, c.sg_degree = zoom, circle, sg_half, sg_degree
    return c


def build_bins(cfg):
    lo = np.empty(cfg.bars, np.int32)
    hi = np.empty(cfg.bars, np.int32)
    _capi.check(_capi.lib().fftl_build_bins(C.byref(cfg), lo.ctypes.data_as(C.POINTER(C.c_int32)),
                                            hi.ctypes.data_as(C.POINTER(C.c_int32))))
    return lo, hi


def build_sg(sg_half, sg_degree):
    m = 2 * sg_half + 1
    coef = np.empty((m, m), np.float32)
    _capi.check(_capi.lib().fftl_build_sg(sg_half, sg_degree, coef.ctypes.data_as(C.POINTER(C.c_float))))
    return coef


class PinnedArray:
    """numpy view over page-locked host memory from fftl_host_alloc."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = _capi.lib().fftl_host_alloc(max(nbytes, 1))
        if not self._ptr:
            raise _capi.FftLinesError(*_capi.last_error())
        buf = (C.c_char * max(nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            _capi.lib().fftl_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def synth_pcm_device(out, first_sample=0, sample_rate=48000.0, seed=0x5EED0000, stream=None):
    """Fill a torch int16 CUDA tensor [channels, samples] with the synthetic test signal."""
    import torch
    assert out.is_cuda and out.dtype == torch.int16 and out.dim() == 2 and out.stride(1) == 1
    st = stream if stream is not None else torch.cuda.current_stream(out.device)
    with torch.cuda.device(out.device):
        _capi.check(_capi.lib().fftl_synth_pcm_device(out.data_ptr(), out.shape[0], out.shape[1], out.stride(0),
                                                      first_sample, sample_rate, seed, st.cuda_stream))
    return out


class BatchedSpectrum:
    """One STFT configuration bound to one GPU (opaque ``fftl_stft`` handle)."""

    def __init__(self, cfg, device=-1):
        self.cfg = cfg.copy()
        self._h = C.c_void_p()
        _capi.check(_capi.lib().fftl_stft_create(C.byref(self.cfg), device, C.byref(self._h)))

    def close(self):
        if self._h:
            _capi.lib().fftl_stft_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def num_frames(self, samples):
        return int(_capi.lib().fftl_stft_num_frames(self._h, samples))

    @property
    def launch_count(self):
        return int(_capi.lib().fftl_stft_launch_count(self._h))

    @property
    def fast_launch_count(self):
        """How many of the bars-kernel launches took the real-input fast path."""
        return int(_capi.lib().fftl_stft_fast_launch_count(self._h))

    @property
    def tmem_launch_count(self):
        """How many fast-path launches ran the tensor-memory variant (constants parked in TMEM, three CTAs per SM)."""
        return int(_capi.lib().fftl_stft_tmem_launch_count(self._h))

    def set_kernel(self, which):
        """_capi.KERNEL_AUTO (default), _capi.KERNEL_GENERIC (the packed-complex baseline kernel) or
        _capi.KERNEL_FAST_REGISTERS (the fast path without its tensor-memory variant)."""
        _capi.check(_capi.lib().fftl_stft_set_kernel(self._h, which))

    def sample_span(self, total_samples, frame_begin, frame_end, with_beat=True):
        """Samples [first, end) a frame-range shard of a ``total_samples`` job must hold (fftl_stft_sample_span)."""
        a, b = C.c_int64(), C.c_int64()
        _capi.check(_capi.lib().fftl_stft_sample_span(self._h, total_samples, frame_begin, frame_end, 1 if with_beat else 0,
                                                      C.byref(a), C.byref(b)))
        return a.value, b.value

    def timing(self):
        """(calls, total_ms, bars_kernel_ms) summed over run_device calls since the last timing()."""
        n, t, b = C.c_int32(), C.c_float(), C.c_float()
        _capi.check(_capi.lib().fftl_stft_timing(self._h, C.byref(n), C.byref(t), C.byref(b)))
        return n.value, t.value, b.value

    # ---- host buffers --------------------------------------------------
    def run(self, pcm, frame_begin=0, frame_end=None, want_i32=True, want_beat=True, out=None,
            interleaved=False):
        """pcm: int16 numpy, planar [channels, samples] (or interleaved [samples, channels]).
        Returns dict(bars float32 [C,F,B], bars_i32, beat structured [C,F])."""
        pcm = np.asarray(pcm)
        if pcm.dtype != np.int16:
            raise TypeError("pcm must be int16")
        if pcm.ndim == 1:
            pcm = pcm[None, :]
        if not pcm.flags.c_contiguous:
            pcm = np.ascontiguousarray(pcm)
        if interleaved:
            S, ch = pcm.shape
            cs, ss = 1, ch
        else:
            ch, S = pcm.shape
            cs, ss = S, 1
        F = self.num_frames(S)
        if frame_end is None:
            frame_end = F
        nf = max(frame_end - frame_begin, 0)
        B = self.cfg.bars
        out = out or {}
        bars = out.get("bars")
        if bars is None:
            bars = np.empty((ch, nf, B), np.float32)
        i32 = out.get("bars_i32")
        if i32 is None and want_i32:
            i32 = np.empty((ch, nf, B), np.int32)
        beat = out.get("beat")
        if beat is None and want_beat:
            beat = np.zeros((ch, nf), BEAT_DTYPE)
        _capi.check(_capi.lib().fftl_stft_run(self._h, pcm.ctypes.data, ch, S, cs, ss, frame_begin, frame_end,
                                              bars.ctypes.data, i32.ctypes.data if i32 is not None else None,
                                              beat.ctypes.data if beat is not None else None))
        return {"bars": bars, "bars_i32": i32, "beat": beat}

    # ---- device buffers ------------------------------------------------
    def run_device(self, pcm, bars, bars_i32=None, beat=None, frame_begin=0, frame_end=None, stream=None):
        """pcm: torch int16 CUDA [channels, samples] (any strides); bars: float32 CUDA
        [channels, frames, B] contiguous; bars_i32 int32 same shape; beat: uint8/int32 CUDA
        buffer of channels*frames*32 bytes.  Asynchronous on ``stream`` (default: torch's current)."""
        import torch
        assert pcm.is_cuda and pcm.dtype == torch.int16 and pcm.dim() == 2
        ch, S = pcm.shape
        F = self.num_frames(S)
        if frame_end is None:
            frame_end = F
        nf = frame_end - frame_begin
        B = self.cfg.bars
        assert bars.is_cuda and bars.dtype == torch.float32 and bars.is_contiguous() and bars.numel() >= ch * nf * B
        if bars_i32 is not None:
            assert bars_i32.dtype == torch.int32 and bars_i32.is_contiguous() and bars_i32.numel() >= ch * nf * B
        if beat is not None:
            assert beat.is_contiguous() and beat.numel() * beat.element_size() >= ch * nf * BEAT_DTYPE.itemsize
        st = stream if stream is not None else torch.cuda.current_stream(pcm.device)
        _capi.check(_capi.lib().fftl_stft_run_device(
            self._h, pcm.data_ptr(), ch, S, pcm.stride(0), pcm.stride(1), frame_begin, frame_end, bars.data_ptr(),
            bars_i32.data_ptr() if bars_i32 is not None else None, beat.data_ptr() if beat is not None else None,
            st.cuda_stream))

    def run_device_span(self, pcm_span, span_first_sample, total_samples, bars, bars_i32=None, beat=None, frame_begin=0,
                        frame_end=None, stream=None):
        """Frame-range shard holding only its own samples: ``pcm_span`` (torch int16 CUDA [channels, span]) starts at
        sample ``span_first_sample`` of a job of ``total_samples`` per channel; frames are numbered as in the whole job."""
        import torch
        assert pcm_span.is_cuda and pcm_span.dtype == torch.int16 and pcm_span.dim() == 2
        ch, span = pcm_span.shape
        if frame_end is None:
            frame_end = self.num_frames(total_samples)
        nf = frame_end - frame_begin
        B = self.cfg.bars
        assert bars.is_cuda and bars.dtype == torch.float32 and bars.is_contiguous() and bars.numel() >= ch * nf * B
        if bars_i32 is not None:
            assert bars_i32.dtype == torch.int32 and bars_i32.is_contiguous() and bars_i32.numel() >= ch * nf * B
        if beat is not None:
            assert beat.is_contiguous() and beat.numel() * beat.element_size() >= ch * nf * BEAT_DTYPE.itemsize
        st = stream if stream is not None else torch.cuda.current_stream(pcm_span.device)
        _capi.check(_capi.lib().fftl_stft_run_device_span(
            self._h, pcm_span.data_ptr(), ch, span_first_sample, span, total_samples, pcm_span.stride(0), pcm_span.stride(1),
            frame_begin, frame_end, bars.data_ptr(), bars_i32.data_ptr() if bars_i32 is not None else None,
            beat.data_ptr() if beat is not None else None, st.cuda_stream))

    def alloc_device_outputs(self, channels, frames, device, want_i32=False, want_beat=True):
        import torch
        bars = torch.empty((channels, frames, self.cfg.bars), dtype=torch.float32, device=device)
        i32 = torch.empty((channels, frames, self.cfg.bars), dtype=torch.int32, device=device) if want_i32 else None
        beat = torch.empty((channels, frames, BEAT_DTYPE.itemsize // 4), dtype=torch.int32, device=device) if want_beat else None
        return bars, i32, beat

    @staticmethod
    def beat_to_numpy(beat_tensor):
        a = beat_tensor.cpu().numpy()
        return a.view(BEAT_DTYPE).reshape(a.shape[:-1])

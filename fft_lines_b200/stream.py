"""Many-stream real-time front end: ``streams`` sliding windows + beat state on the GPU,
one push = one tick of the reference's WM_TIMER per stream (``fftl_stream_*``)."""
import ctypes as C

import numpy as np

from . import _capi
from .stft import BEAT_DTYPE


class StreamingSpectrum:
    """Mirrors GetAudioBuffer/MoveArray (reference wasapi_audio.cpp:390-484) followed by the
    per-tick chain (fft_lines.c:186-236, :281) for many independent streams."""

    def __init__(self, cfg, streams, move_mode=_capi.MOVE_EXACT, device=-1):
        self.cfg = cfg.copy()
        self.streams = int(streams)
        self._h = C.c_void_p()
        _capi.check(_capi.lib().fftl_stream_create(C.byref(self.cfg), self.streams, move_mode, device, C.byref(self._h)))

    def close(self):
        if self._h:
            _capi.lib().fftl_stream_destroy(self._h)
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

    def reset(self):
        _capi.check(_capi.lib().fftl_stream_reset(self._h))

    @property
    def launch_count(self):
        return int(_capi.lib().fftl_stream_launch_count(self._h))

    @property
    def fused_tick_count(self):
        """Pushes that ran as ONE kernel (n = 1024 / 2048, exact window move: csrc/stream_tick.cuh)."""
        return int(_capi.lib().fftl_stream_fused_tick_count(self._h))

    def set_tick_kernel(self, which):
        """``TICK_AUTO`` (one kernel per tick where built) or ``TICK_THREE_LAUNCHES`` (window move, bars kernel, beat stage)."""
        _capi.check(_capi.lib().fftl_stream_set_tick_kernel(self._h, which))

    def push(self, fresh, want_i32=True, out=None):
        """fresh: int16 numpy [streams, count]. Returns dict(bars [streams,B], bars_i32, beat [streams])."""
        fresh = np.ascontiguousarray(fresh, dtype=np.int16)
        assert fresh.ndim == 2 and fresh.shape[0] == self.streams
        count = fresh.shape[1]
        out = out or {}
        bars = out.get("bars")
        if bars is None:
            bars = np.empty((self.streams, self.cfg.bars), np.float32)
        i32 = out.get("bars_i32")
        if i32 is None and want_i32:
            i32 = np.empty((self.streams, self.cfg.bars), np.int32)
        beat = out.get("beat")
        if beat is None:
            beat = np.zeros(self.streams, BEAT_DTYPE)
        _capi.check(_capi.lib().fftl_stream_push(self._h, fresh.ctypes.data, count, count, bars.ctypes.data,
                                                 i32.ctypes.data if i32 is not None else None, beat.ctypes.data))
        return {"bars": bars, "bars_i32": i32, "beat": beat}

    def push_device(self, fresh, bars, bars_i32=None, beat=None, stream=None):
        """torch CUDA tensors: fresh int16 [streams, count] (row stride free), bars float32 [streams, B]."""
        import torch
        assert fresh.is_cuda and fresh.dtype == torch.int16 and fresh.dim() == 2 and fresh.stride(1) == 1
        assert fresh.shape[0] == self.streams and bars.is_contiguous() and bars.dtype == torch.float32
        st = stream if stream is not None else torch.cuda.current_stream(fresh.device)
        _capi.check(_capi.lib().fftl_stream_push_device(
            self._h, fresh.data_ptr(), fresh.shape[1], fresh.stride(0), bars.data_ptr(),
            bars_i32.data_ptr() if bars_i32 is not None else None, beat.data_ptr() if beat is not None else None,
            st.cuda_stream))

    def push_device_f32(self, fresh, bars, bars_i32=None, beat=None, stream=None):
        """Capture-format push (wasapi_audio.cpp:398): fresh float32 CUDA [streams, count] with any strides, e.g. one
        channel of interleaved stereo; converted on the device as ``(int16_t)(x * 65536)``."""
        import torch
        assert fresh.is_cuda and fresh.dtype == torch.float32 and fresh.dim() == 2 and fresh.shape[0] == self.streams
        assert bars.is_contiguous() and bars.dtype == torch.float32
        st = stream if stream is not None else torch.cuda.current_stream(fresh.device)
        _capi.check(_capi.lib().fftl_stream_push_device_f32(
            self._h, fresh.data_ptr(), fresh.shape[1], fresh.stride(0), fresh.stride(1), bars.data_ptr(),
            bars_i32.data_ptr() if bars_i32 is not None else None, beat.data_ptr() if beat is not None else None,
            st.cuda_stream))

    def set_history(self, depth):
        """Keep the last ``depth`` bar frames of every stream on the device (the reference's 3D-mode trail)."""
        _capi.check(_capi.lib().fftl_stream_set_history(self._h, depth))
        self._depth = depth

    def history_device(self, device="cuda"):
        """torch float32 [streams, depth, bars], oldest frame first."""
        import torch
        out = torch.empty((self.streams, self._depth, self.cfg.bars), dtype=torch.float32, device=device)
        _capi.check(_capi.lib().fftl_stream_history_device(self._h, out.data_ptr(), torch.cuda.current_stream(out.device).cuda_stream))
        return out

    def waveform_device(self, points=2048, scale=0.02, offset=0.0, device="cuda"):
        """Waveform view of the current windows (fft_lines.c:239-247): int32 [streams, points]."""
        import torch
        out = torch.empty((self.streams, points), dtype=torch.int32, device=device)
        _capi.check(_capi.lib().fftl_stream_waveform_device(self._h, points, scale, offset, out.data_ptr(),
                                                            torch.cuda.current_stream(out.device).cuda_stream))
        return out

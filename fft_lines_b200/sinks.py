"""Callers and sinks either side of the hot path: the Settings.txt line, the
RANGE255 colour index, the serial/LED byte and the beat-marker sweep."""
import ctypes as C

from . import _capi
from ._capi import FftlSettings, FftlConfig


def settings_default():
    s = FftlSettings()
    _capi.lib().fftl_settings_default(C.byref(s))
    return s


def settings_parse(text):
    """Parse the 11-number line of the reference's Settings.txt (settingsFile.c:172-247)."""
    raw = text.encode() if isinstance(text, str) else bytes(text)
    s = FftlSettings()
    _capi.check(_capi.lib().fftl_settings_parse(raw, len(raw), C.byref(s)))
    return s


def settings_format(s):
    buf = C.create_string_buffer(256)
    n = _capi.lib().fftl_settings_format(C.byref(s), buf, 256)
    if n < 0:
        raise ValueError("settings line does not fit")
    return buf.value.decode()


def settings_to_config(s, n=2048, hop=512, sample_rate=48000.0):
    c = FftlConfig()
    _capi.lib().fftl_settings_to_config(C.byref(s), n, hop, sample_rate, C.byref(c))
    return c


def color_index(i, count):
    return int(_capi.lib().fftl_color_index(i, count))


def led_bytes_device(heights, led_bar=4, stream=None):
    """heights: torch int32 CUDA [frames, bars] -> (uint8 bytes [frames], uint8 valid [frames])."""
    import torch
    assert heights.is_cuda and heights.dtype == torch.int32 and heights.dim() == 2 and heights.is_contiguous()
    frames, bars = heights.shape
    out = torch.empty(frames, dtype=torch.uint8, device=heights.device)
    valid = torch.empty(frames, dtype=torch.uint8, device=heights.device)
    st = stream if stream is not None else torch.cuda.current_stream(heights.device)
    _capi.check(_capi.lib().fftl_led_bytes_device(heights.data_ptr(), frames, bars, led_bar, out.data_ptr(), valid.data_ptr(),
                                                  st.cuda_stream))
    return out, valid


def marker_sweep_device(beat, time_delta, start_pos=0, start_dir=0, stream=None):
    """beat: torch int32 CUDA [frames, 8] (fftl_beat records of one channel) -> (pos [frames], dir [frames])."""
    import torch
    assert beat.is_cuda and beat.is_contiguous() and beat.dim() == 2 and beat.shape[1] * beat.element_size() == 32
    frames = beat.shape[0]
    pos = torch.empty(frames, dtype=torch.int32, device=beat.device)
    direction = torch.empty(frames, dtype=torch.int32, device=beat.device)
    st = stream if stream is not None else torch.cuda.current_stream(beat.device)
    _capi.check(_capi.lib().fftl_marker_sweep_device(beat.data_ptr(), frames, time_delta, start_pos, start_dir, pos.data_ptr(),
                                                     direction.data_ptr(), st.cuda_stream))
    return pos, direction

"""Python mirror of the reference's ``beatDetector.h`` (reference
fft_lines/beatDetector.h:3-11), bound to libfftlines_cuda's drop-in symbols.
The state (160-sample ring, 256-entry tempo ring) lives on the GPU.
"""
import ctypes as C

import numpy as np

from . import _capi


def AddBeatSample(weightedAverage):
    _capi.lib().AddBeatSample(int(weightedAverage))


def GetBeatValue():
    return _capi.lib().GetBeatValue()


def BassBeatDetector(input, output, beatinput, beatoutput):
    """``input`` is unused (as in the reference); ``output`` is the last spectrum."""
    def p(a):
        if a is None:
            return None
        if not isinstance(a, np.ndarray) or a.dtype != np.complex64 or not a.flags.c_contiguous:
            raise TypeError("need C-contiguous complex64 arrays")
        return a.ctypes.data
    _capi.lib().BassBeatDetector(p(input), p(output), p(beatinput), p(beatoutput))


def reset():
    """Zero the beat state (the reference has no reset; its globals start zeroed)."""
    _capi.lib().fftl_compat_reset_beat()


def _sym(name, t):
    return _capi.data_symbol(name, t)


def get_beatMovementPerSec():
    return _sym("beatMovementPerSec", C.c_float).value


def get_timeDelta():
    return _sym("timeDelta", C.c_float).value


def set_timeDelta(v):
    """The caller writes timeDelta each tick (reference fft_lines.c:265)."""
    _sym("timeDelta", C.c_float).value = v


def get_beatposition():
    return _sym("beatposition", C.c_int).value


def set_beatposition(v):
    _sym("beatposition", C.c_int).value = v


def get_direction():
    return _sym("direction", C.c_int).value


def set_direction(v):
    _sym("direction", C.c_int).value = v
